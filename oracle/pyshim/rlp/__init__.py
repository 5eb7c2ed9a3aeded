"""Minimal RLP encoder (ints as big-endian minimal bytes, bytes, nested lists)."""


def _enc_len(n, off):
    if n < 56:
        return bytes([off + n])
    b = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([off + 55 + len(b)]) + b


def encode(x):
    if isinstance(x, int):
        x = b"" if x == 0 else x.to_bytes((x.bit_length() + 7) // 8, "big")
    if isinstance(x, (bytes, bytearray)):
        x = bytes(x)
        if len(x) == 1 and x[0] < 0x80:
            return x
        return _enc_len(len(x), 0x80) + x
    if isinstance(x, (list, tuple)):
        body = b"".join(encode(i) for i in x)
        return _enc_len(len(body), 0xC0) + body
    raise TypeError(type(x))
