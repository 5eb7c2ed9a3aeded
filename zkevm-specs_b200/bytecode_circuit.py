"""Bytecode circuit, host side — API of /root/reference/src/zkevm_specs/bytecode_circuit.py.

`Row`, `UnrolledBytecode`, `assign_bytecode_circuit`, `assign_push_table`,
`assign_keccak_table` build the witness exactly like the reference (:15-186).  The checks of
`check_bytecode_row` (:37-100) run on the device: `verify_bytecode_circuit` packs all rows
and makes ONE zk_check call (the reference's tests loop over rows in Python and call
check_bytecode_row per row, tests/test_bytecode_circuit.py:26-47); `check_bytecode_row`
keeps the per-row signature for drop-in use and checks a 2-row matrix.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Sequence, Set, Tuple

import numpy as np

from . import native, packing
from .evm_circuit.spec import BytecodeFieldTag, get_push_size
from .evm_circuit.table import BytecodeTableRow, KeccakTableRow
from .evm_circuit.typing import KeccakCircuit
from .util.arithmetic import FQ, Word
from .util.hash import EMPTY_HASH


@dataclass
class Row:
    q_first: FQ
    q_last: FQ
    hash: Word
    tag: FQ
    index: FQ
    value: FQ
    is_code: FQ
    push_data_left: FQ
    value_rlc: FQ
    length: FQ
    push_data_size: FQ


@dataclass
class UnrolledBytecode:
    bytes: bytes
    rows: Sequence[BytecodeTableRow]


def assign_bytecode_circuit(k: int, bytecodes: Sequence[UnrolledBytecode], keccak_randomness: FQ) -> List[Row]:
    """2^k rows: every bytecode's Header + Byte rows, then Header padding (reference :104-167)."""
    size = 1 << k
    rows: List[Row] = []

    def emit(hash_, tag, index, value, is_code, left, rlc, length, push_size):
        at = len(rows)
        rows.append(Row(FQ(int(at == 0)), FQ(int(at == size - 1)), hash_, FQ(tag), FQ(index), FQ(value),
                        FQ(is_code), FQ(left), FQ(rlc), FQ(length), FQ(push_size)))

    for bc in bytecodes:
        pending, acc = 0, FQ(0)
        for idx, trow in enumerate(bc.rows):
            left = pending
            is_code = left == 0
            push_size = 0
            if idx > 0:
                push_size = get_push_size(trow.value.expr().n)
                pending = push_size if is_code else left - 1
                acc = acc * keccak_randomness + trow.value
            emit(trow.bytecode_hash, trow.field_tag.expr(), trow.index.expr(), trow.value.expr(),
                 trow.is_code.expr(), left, acc, len(bc.bytes), push_size)
            if len(rows) == size:
                return rows
    while len(rows) < size:
        emit(Word(EMPTY_HASH), int(BytecodeFieldTag.Header), 0, 0, 0, 0, 0, 0, 0)
    return rows


def assign_push_table() -> List[Tuple[FQ, FQ]]:
    """byte -> number of pushed bytes (reference :174-178)."""
    return [(FQ(b), FQ(get_push_size(b))) for b in range(256)]


def assign_keccak_table(bytecodes: Sequence[bytes], keccak_randomness: FQ) -> Set[KeccakTableRow]:
    kc = KeccakCircuit()
    for code in bytecodes:
        kc.add(code, keccak_randomness)
    return set(kc.rows)


def pack_rows(rows: Sequence[Row]) -> np.ndarray:
    return packing.pack(rows, packing.bytecode_circuit_row, 12)


def pack_push_table(push_table: Iterable) -> np.ndarray:
    return packing.matrix_from_ints([[packing.cell_int(a), packing.cell_int(b)] for a, b in push_table], 2)


def pack_keccak_table(keccak_table: Iterable[KeccakTableRow]) -> np.ndarray:
    return packing.pack(keccak_table, packing.keccak_table_row, 5)


def check_matrices(cols: np.ndarray, push: np.ndarray, keccak: np.ndarray, keccak_randomness,
                   ctx: native.Context = None, row_begin: int = 0, row_end: int = None,
                   flags: int = native.FLAG_WRAP):
    """Run the device checker on packed matrices; returns (first_fail, fail_count)."""
    ctx = ctx or native.default_context()
    ctx.set_challenge(native.CHALLENGE_KECCAK, packing.cell_int(keccak_randomness))
    ctx.upload_table(native.TABLE_PUSH, push)
    ctx.upload_table(native.TABLE_KECCAK, keccak)
    ctx.upload_columns(native.CIRCUIT_BYTECODE, cols)
    n = cols.shape[1]
    return ctx.check(native.CIRCUIT_BYTECODE, row_begin, n if row_end is None else row_end, 0, flags)


def verify_bytecode_circuit(rows: Sequence[Row], push_table, keccak_table, keccak_randomness,
                            ctx: native.Context = None) -> None:
    """All rows in one device call; raises AssertionError naming the first failing row and
    constraint, like the reference's loop stops at its first failing assert."""
    ff, _ = check_matrices(pack_rows(rows), pack_push_table(push_table), pack_keccak_table(keccak_table),
                           keccak_randomness, ctx)
    hit = native.first_failure(ff, native.CIRCUIT_BYTECODE)
    if hit is not None:
        row, cid, cls, name = hit
        raise AssertionError(f"bytecode circuit row {row}: {name}")


def check_bytecode_row(cur: Row, next: Row, push_table, keccak_table, keccak_randomness,
                       ctx: native.Context = None) -> None:
    """Reference signature (bytecode_circuit.py:37): checks `cur` against `next` on the device."""
    cols = pack_rows([cur, next])
    ctx = ctx or native.default_context()
    ff, _ = check_matrices(cols, pack_push_table(push_table), pack_keccak_table(keccak_table),
                           keccak_randomness, ctx, 0, 1, 0)
    hit = native.first_failure(ff, native.CIRCUIT_BYTECODE)
    if hit is not None:
        raise AssertionError(f"bytecode circuit: {hit[3]}")
