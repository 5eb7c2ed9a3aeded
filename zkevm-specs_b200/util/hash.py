"""Keccak-256 (original 0x01 padding, as Ethereum uses) for host-side witness generation.

The reference gets it from pycryptodome (`util/hash.py:7-10`); hashlib only ships the
NIST-padded SHA3, so the sponge is written out here.  Pinned in tests by the standard
vectors keccak256(b"") and keccak256(0x80) (== EMPTY_HASH / EMPTY_TRIE_HASH)."""
from typing import Union

_MASK = (1 << 64) - 1
_ROUND_CONSTANTS = []
_ROTATIONS = [[0] * 5 for _ in range(5)]


def _init_tables():
    lfsr = 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            if lfsr & 1:
                rc |= 1 << ((1 << j) - 1)
            lfsr = ((lfsr << 1) ^ (0x71 if lfsr & 0x80 else 0)) & 0xFF
        _ROUND_CONSTANTS.append(rc)
    x, y = 1, 0
    for t in range(24):
        _ROTATIONS[x][y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5


_init_tables()


def _rotl(v: int, s: int) -> int:
    return ((v << s) | (v >> (64 - s))) & _MASK if s else v


def _permute(lanes):
    for rc in _ROUND_CONSTANTS:
        col = [lanes[x] ^ lanes[x + 5] ^ lanes[x + 10] ^ lanes[x + 15] ^ lanes[x + 20] for x in range(5)]
        for x in range(5):
            d = col[(x + 4) % 5] ^ _rotl(col[(x + 1) % 5], 1)
            for y in range(0, 25, 5):
                lanes[x + y] ^= d
        moved = [0] * 25
        for x in range(5):
            for y in range(5):
                moved[y + 5 * ((2 * x + 3 * y) % 5)] = _rotl(lanes[x + 5 * y], _ROTATIONS[x][y])
        for y in range(0, 25, 5):
            row = moved[y : y + 5]
            for x in range(5):
                lanes[x + y] = row[x] ^ (~row[(x + 1) % 5] & _MASK & row[(x + 2) % 5])
        lanes[0] ^= rc


def keccak256(data: Union[str, bytes, bytearray]) -> bytes:
    if isinstance(data, str):
        data = bytes.fromhex(data)
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    msg.extend(b"\x00" * (-len(msg) % rate))
    msg[-1] |= 0x80
    lanes = [0] * 25
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            lanes[i] ^= int.from_bytes(msg[off + 8 * i : off + 8 * i + 8], "little")
        _permute(lanes)
    return b"".join(lanes[i].to_bytes(8, "little") for i in range(4))


EMPTY_HASH = int.from_bytes(keccak256(b""), "big")
EMPTY_CODE_HASH = EMPTY_HASH
EMPTY_TRIE_HASH = int.from_bytes(keccak256("80"), "big")
