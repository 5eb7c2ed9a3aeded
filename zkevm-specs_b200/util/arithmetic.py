"""Host-side field and word types with the reference's API surface.

Mirrors `FQ`, `RLC`, `Word`, `WordOrValue`, `Expression`, `linear_combine_bytes`,
`add_words` of /root/reference/src/zkevm_specs/util/arithmetic.py:9-276 (the reference
subclasses py_ecc.bn128.FQ; here the class is self-contained).  These objects only BUILD
witnesses on the host; every check over them runs in CUDA behind libzkcheck.so.
"""
from __future__ import annotations

from typing import Protocol, Sequence, Tuple, Union, runtime_checkable

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MAX_N_BYTES = 31


def _raw(x) -> int:
    if isinstance(x, FQ):
        return x.n
    if isinstance(x, int):
        return int(x)
    raise TypeError(f"Expected an int or FQ object, but got object of type {type(x)}")


class FQ:
    """BN254 scalar-field element (reference: util/arithmetic.py:41-63).

    Same observable behaviour as py_ecc's FQ: ints are reduced on construction, `==` against
    an int compares with the RAW int, `/` multiplies by the inverse and inv(0) == 0."""

    __slots__ = ("n",)
    field_modulus = FR_MODULUS

    def __init__(self, value: Union[int, "FQ"]) -> None:
        if isinstance(value, FQ):
            self.n = value.n
        elif isinstance(value, int):
            self.n = int(value) % FR_MODULUS
        else:
            raise TypeError(f"Expected an int or FQ object, but got object of type {type(value)}")

    # arithmetic -------------------------------------------------------------------------
    def __add__(self, o):
        return FQ(self.n + _raw(o))

    __radd__ = __add__

    def __sub__(self, o):
        return FQ(self.n - _raw(o))

    def __rsub__(self, o):
        return FQ(_raw(o) - self.n)

    def __mul__(self, o):
        return FQ(self.n * _raw(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return FQ(self.n * pow(_raw(o) % FR_MODULUS, -1, FR_MODULUS)) if _raw(o) % FR_MODULUS else FQ(0)

    def __rtruediv__(self, o):
        return FQ(_raw(o) * self.inv().n)

    def __pow__(self, e: int):
        return FQ(pow(self.n, e, FR_MODULUS))

    def __neg__(self):
        return FQ(-self.n)

    def __eq__(self, o) -> bool:
        return self.n == _raw(o)

    def __ne__(self, o) -> bool:
        return not self == o

    def __hash__(self) -> int:
        return hash(self.n)

    def __int__(self) -> int:
        return self.n

    def __repr__(self) -> str:
        return hex(self.n)

    def expr(self) -> "FQ":
        return FQ(self)

    def inv(self) -> "FQ":
        return FQ(pow(self.n, -1, FR_MODULUS)) if self.n else FQ(0)

    @classmethod
    def zero(cls) -> "FQ":
        return cls(0)

    @classmethod
    def one(cls) -> "FQ":
        return cls(1)


IntOrFQ = Union[int, FQ]


@runtime_checkable
class Expression(Protocol):
    def expr(self) -> FQ:
        ...


def linear_combine_bytes(seq: Sequence[IntOrFQ], base: IntOrFQ, range_check: bool = True) -> FQ:
    """Horner evaluation, little-endian (reference: util/arithmetic.py:9-24)."""
    b = _raw(base)
    acc = 0
    for limb in reversed(seq):
        v = _raw(limb)
        if range_check:
            assert 0 <= v < 256, "Each byte should fit in 8-bit"
        acc = (acc * b + v) % FR_MODULUS
    return FQ(acc)


def bytes_to_fq(value: bytes) -> FQ:
    assert len(value) <= MAX_N_BYTES
    return FQ(int.from_bytes(value, "little"))


class RLC:
    """Random linear combination of up to n_bytes little-endian bytes (util/arithmetic.py:69-96)."""

    def __init__(self, value: Union[int, bytes], randomness: FQ = FQ(0), n_bytes: int = 32) -> None:
        if isinstance(value, int):
            value = value.to_bytes(n_bytes, "little")
        if len(value) > n_bytes:
            raise ValueError(f"RLC expects to have {n_bytes} bytes, but got {len(value)} bytes")
        value = bytes(value).ljust(n_bytes, b"\x00")
        self.le_bytes = value
        self.int_value = int.from_bytes(value, "little")
        self.rlc_value = linear_combine_bytes(value, randomness)

    def expr(self) -> FQ:
        return FQ(self.rlc_value)

    def __hash__(self) -> int:
        return hash(self.rlc_value)

    def __repr__(self) -> str:
        return f"RLC({self.int_value})"


class Word:
    """256-bit word as (lo, hi) 128-bit field cells (util/arithmetic.py:99-168)."""

    def __init__(self, value, check: bool = True) -> None:
        if isinstance(value, tuple):
            self.lo, self.hi = value
            assert not check or (self.lo.expr().n < 1 << 128 and self.hi.expr().n < 1 << 128)
            return
        if isinstance(value, int):
            assert not check or value < 1 << 256
            value = int(value).to_bytes(32, "little")
        assert isinstance(value, (bytes, bytearray)) and len(value) == 32, "Word expects 32 bytes"
        self.lo = FQ(int.from_bytes(value[:16], "little"))
        self.hi = FQ(int.from_bytes(value[16:], "little"))

    @classmethod
    def from_lo(cls, lo: Expression) -> "Word":
        return cls((lo, FQ(0)))

    def int_value(self) -> int:
        return self.lo.expr().n + (self.hi.expr().n << 128)

    def to_lo_hi(self) -> Tuple[FQ, FQ]:
        return self.lo.expr(), self.hi.expr()

    def to_le_bytes(self) -> Tuple[FQ, ...]:
        raw = self.lo.expr().n.to_bytes(16, "little") + self.hi.expr().n.to_bytes(16, "little")
        return tuple(FQ(b) for b in raw)

    def to_64s(self) -> Tuple[FQ, ...]:
        lo, hi = self.lo.expr().n, self.hi.expr().n
        assert lo < 1 << 128 and hi < 1 << 128
        m = (1 << 64) - 1
        return FQ(lo & m), FQ(lo >> 64), FQ(hi & m), FQ(hi >> 64)

    def select(self, selector: FQ) -> "Word":
        return Word((selector * self.lo, selector * self.hi))

    def __add__(self, other: "Word") -> "Word":
        return Word((self.lo.expr() + other.lo.expr(), self.hi.expr() + other.hi.expr()))

    def __eq__(self, other) -> bool:
        assert isinstance(other, Word)
        return self.lo.expr() == other.lo.expr() and self.hi.expr() == other.hi.expr()

    def __hash__(self) -> int:
        return hash((self.lo, self.hi))

    def __repr__(self) -> str:
        return f"Word({hex(self.int_value())})"


class WordOrValue(Word):
    """A Word, or a plain field value with hi == 0 (util/arithmetic.py:171-195).  The
    `is_word` bit travels to the device as a per-row type flag (zk_upload_row_flags)."""

    def __init__(self, value: Union[Word, Expression]) -> None:
        if isinstance(value, Word):
            self.is_word = True
            self.lo, self.hi = value.lo, value.hi
        else:
            self.is_word = False
            self.lo, self.hi = value, FQ(0)

    def value(self) -> Expression:
        assert not self.is_word
        return self.lo

    def __repr__(self) -> str:
        return super().__repr__() if self.is_word else f"Value({hex(self.lo.expr().n)})"


IntOrFQOrWord = Union[int, FQ, Word]


def add_words(addends: Sequence[Word]) -> Tuple[Word, FQ]:
    """256-bit sum via lo/hi halves with carries (util/arithmetic.py:236-242)."""
    lo = sum(w.lo.expr().n for w in addends) % FR_MODULUS
    carry_lo, sum_lo = divmod(lo, 1 << 128)
    hi = (sum(w.hi.expr().n for w in addends) + carry_lo) % FR_MODULUS
    carry_hi, sum_hi = divmod(hi, 1 << 128)
    return Word((FQ(sum_lo), FQ(sum_hi))), FQ(carry_hi)
