"""Copy circuit, host side — API of /root/reference/src/zkevm_specs/copy_circuit.py.

`verify_copy_table(copy_circuit, tables, r)` keeps the reference signature (:92) and error
convention; the 26 gates of verify_row / verify_step (:23-89) and the per-row table lookups
(:106-130) run on the device in one zk_check(ZK_CIRCUIT_COPY) over all rows (rotations wrap,
like the reference's `(i + 1) % n`)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import native, packing
from .evm_circuit.main import raise_first_failure
from .evm_circuit.table import Tables
from .evm_circuit.typing import CopyCircuit


def upload_copy_tables(ctx: native.Context, tables: Tables) -> None:
    rws = list(tables.rw_table)
    ctx.upload_table(native.TABLE_RW, packing.pack(rws, packing.rw_table_row, 14),
                     flags=np.array([packing.rw_table_flags(r) for r in rws], dtype=np.uint8))
    ctx.upload_table(native.TABLE_BYTECODE, packing.pack(tables.bytecode_table, packing.bytecode_table_row, 6))
    txs = list(tables.tx_table)
    ctx.upload_table(native.TABLE_TX, packing.pack(txs, packing.tx_table_row, 5),
                     flags=np.array([packing.word_flag(r.value) for r in txs], dtype=np.uint8))


def check_copy_rows(ctx: native.Context, rows, r):
    """copy rows (Python objects) against the tables already on the device"""
    cols = packing.pack(rows, packing.copy_circuit_row, 20)
    flags = np.array([packing.word_flag(x.id) if hasattr(x.id, "is_word") else 1 for x in rows], dtype=np.uint8)
    ctx.set_challenge(native.CHALLENGE_KECCAK, packing.cell_int(r))
    ctx.upload_columns(native.CIRCUIT_COPY, cols, flags=flags)
    return ctx.check(native.CIRCUIT_COPY, 0, cols.shape[1], 0, native.FLAG_WRAP)


def verify_copy_table(copy_circuit: CopyCircuit, tables: Tables, r, ctx: Optional[native.Context] = None) -> None:
    ctx = ctx or native.default_context()
    rows = list(copy_circuit.table())
    if not rows:
        return
    upload_copy_tables(ctx, tables)
    ff, _ = check_copy_rows(ctx, rows, r)
    raise_first_failure(ff, native.CIRCUIT_COPY, "copy row")
