"""Import-only stubs: bn128 curve arithmetic is out of scope."""


def _unavailable(*a, **k):
    raise NotImplementedError("bn128 curve arithmetic is not part of the oracle shim")


is_on_curve = add = multiply = eq = double = neg = _unavailable
b = b2 = b12 = None
G1 = G2 = Z1 = Z2 = None


class FQ2:  # pragma: no cover - placeholder
    def __init__(self, *a, **k):
        _unavailable()


class FQ12(FQ2):
    pass
