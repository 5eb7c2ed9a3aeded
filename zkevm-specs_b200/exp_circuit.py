"""Exponentiation circuit, host side — API of /root/reference/src/zkevm_specs/exp_circuit.py and of
`ExpCircuit` (evm_circuit/typing.py:868-994).

`ExpCircuit.add_event(base, exponent, identifier)` unrolls exponentiation by squaring into one row
per multiplication (a * b == d mod 2^256 with the parity decomposition of the exponent), exactly
the rows the reference builds; `verify_exp_circuit(exp_circuit)` keeps the reference signature
(:88) and checks all rows in one zk_check(ZK_CIRCUIT_EXP)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from . import native, packing
from .evm_circuit.main import raise_first_failure
from .util.arithmetic import FQ, Word

POW2 = 1 << 256


@dataclass(frozen=True)
class ExpCircuitRow:
    q_usable: FQ
    is_step: FQ
    identifier: FQ
    is_last: FQ
    base: Word
    exponent: Word
    exponentiation: Word
    a: Word
    b: Word
    c: Word
    d: Word
    q: Word
    r: FQ


class ExpCircuit:
    OFFSET_INCREMENT = 7

    def __init__(self, max_exp_steps: int = 100) -> None:
        self.rows: List[ExpCircuitRow] = []
        self.max_exp_steps = max_exp_steps

    def table(self) -> Sequence[ExpCircuitRow]:
        return self.rows

    def add_event(self, base: int, exponent: int, identifier) -> "ExpCircuit":
        """multiplications of square-and-multiply, emitted from the final product backwards"""
        muls: List[Tuple[int, int, int]] = []

        def power(e: int) -> int:
            if e == 0:
                return 1
            if e == 1:
                return base
            half = power(e // 2)
            sq = half * half % POW2
            muls.append((half, half, sq))
            if e % 2 == 0:
                return sq
            odd = base * sq % POW2
            muls.append((sq, base, odd))
            return odd

        power(exponent)
        muls.reverse()
        e = exponent
        for k, (a, b, d) in enumerate(muls):
            quotient, is_odd = divmod(e, 2)
            self.rows.append(ExpCircuitRow(FQ(1), FQ(1), FQ(packing.cell_int(identifier)), FQ(int(k == len(muls) - 1)),
                                           Word(base), Word(e), Word(d), Word(a), Word(b), Word(0), Word(d),
                                           Word(quotient), FQ(is_odd)))
            e = e - 1 if is_odd else e // 2
        return self

    def fill_dummy_events(self) -> "ExpCircuit":
        for _ in range(self.max_exp_steps * self.OFFSET_INCREMENT - len(self.rows)):
            self.rows.append(ExpCircuitRow(FQ(1), FQ(0), FQ(0), FQ(0), Word(1), Word(1), Word(1), Word(1), Word(1),
                                           Word(0), Word(1), Word(0), FQ(1)))
        return self


def exp_row(r: ExpCircuitRow) -> List[int]:
    c = packing.cell_int
    out = [c(r.q_usable), c(r.is_step), c(r.identifier), c(r.is_last)]
    for w in (r.base, r.exponent, r.exponentiation, r.a, r.b, r.c, r.d, r.q):
        out += [c(w.lo), c(w.hi)]
    return out + [c(r.r)]


def check_rows(ctx: native.Context, cols, row_begin=0, row_end=None, row_base=0, flags=native.FLAG_WRAP):
    ctx.upload_columns(native.CIRCUIT_EXP, cols)
    return ctx.check(native.CIRCUIT_EXP, row_begin, cols.shape[1] if row_end is None else row_end, row_base, flags)


def verify_exp_circuit(exp_circuit: ExpCircuit, ctx: Optional[native.Context] = None) -> None:
    rows = list(exp_circuit.table())
    if not rows:
        return
    ctx = ctx or native.default_context()
    ff, _ = check_rows(ctx, packing.matrix_from_ints([exp_row(r) for r in rows], 21))
    raise_first_failure(ff, native.CIRCUIT_EXP, "exp row")
