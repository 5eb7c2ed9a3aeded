from Crypto.Hash.keccak import keccak256 as _k


def keccak(primitive=None, hexstr=None, text=None):
    if primitive is None and hexstr is not None:
        primitive = bytes.fromhex(hexstr.removeprefix("0x"))
    if primitive is None and text is not None:
        primitive = text.encode()
    return _k(bytes(primitive))
