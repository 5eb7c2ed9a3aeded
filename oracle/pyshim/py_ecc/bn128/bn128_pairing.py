"""Import-only stub (pairing is out of scope)."""


def pairing(*a, **k):
    raise NotImplementedError("bn128 pairing is not part of the oracle shim")
