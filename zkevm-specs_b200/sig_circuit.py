"""Sig circuit, host side — the Fr parts of /root/reference/src/zkevm_specs/sig_circuit.py.

`Row` (:7-49), `Witness` (:107-109) and `verify_circuit(witness, keccak_randomness)` (:113-123)
keep the reference's names and meaning.  The ECDSA check (`ECDSAVerifyChip.verify`,
util/ec.py:109-117) is third-party curve math (eth_keys): the caller's `ecdsa_chip` supplies
`pub_key_x_bytes`, `pub_key_y_bytes`, `msg_hash_bytes`, `sig_v / sig_r / sig_s` (objects with
`.le_bytes`) and `verify() -> bool`; its verdict travels to the device as a row flag and is
compared there with the row's `is_valid`, after the constraints the reference checks first."""
from __future__ import annotations

from typing import List, NamedTuple, Optional

import numpy as np

from . import native, packing
from .tx_circuit import KeccakTable, _word_cells  # noqa: F401  (same table class as util/tables.py:10-33)
from .util.arithmetic import FQ, Word


class Row:
    def __init__(self, pub_key_hash: bytes, address: FQ, msg_hash: Word, ecdsa_chip, is_valid: bool = True) -> None:
        self.ecdsa_chip = ecdsa_chip
        self.pub_key_x_bytes = ecdsa_chip.pub_key_x_bytes
        self.pub_key_y_bytes = ecdsa_chip.pub_key_y_bytes
        self.msg_hash_bytes = ecdsa_chip.msg_hash_bytes
        self.msg_hash = msg_hash
        self.sig_v = FQ(int.from_bytes(ecdsa_chip.sig_v.le_bytes, "little"))
        self.sig_r = Word(int.from_bytes(ecdsa_chip.sig_r.le_bytes, "little"))
        self.sig_s = Word(int.from_bytes(ecdsa_chip.sig_s.le_bytes, "little"))
        self.recovered_addr = address
        self.is_valid = is_valid
        self.pub_key_hash = pub_key_hash


class Witness(NamedTuple):
    rows: List[Row]
    keccak_table: KeccakTable


def pack_witness(witness: Witness):
    """-> (rows uint64[21][n][4], flags uint8[n], keccak uint64[5][k][4])"""
    c = packing.cell_int
    cells, flags = [], []
    for row in witness.rows:
        chip = row.ecdsa_chip
        assert row.pub_key_x_bytes == chip.pub_key_x_bytes  # sig_circuit.py:67-69
        assert row.pub_key_y_bytes == chip.pub_key_y_bytes
        assert row.msg_hash_bytes == chip.msg_hash_bytes
        cells.append([c(row.sig_v), c(row.recovered_addr), *_word_cells(bytes(row.pub_key_x_bytes)),
                      *_word_cells(bytes(row.pub_key_y_bytes)), *_word_cells(bytes(row.pub_key_hash)),
                      c(row.msg_hash.lo), c(row.msg_hash.hi), *_word_cells(bytes(row.msg_hash_bytes)), int(row.is_valid),
                      c(row.sig_r.lo), c(row.sig_r.hi), c(row.sig_s.lo), c(row.sig_s.hi),
                      *_word_cells(bytes(chip.sig_r.le_bytes)), *_word_cells(bytes(chip.sig_s.le_bytes))])
        flags.append(int(bool(chip.verify())) << 1)
    keccak = packing.matrix_from_ints([[c(a), c(b), c(l), c(o.lo), c(o.hi)] for a, b, l, o in witness.keccak_table.table], 5)
    return packing.matrix_from_ints(cells, 21), np.array(flags, dtype=np.uint8), keccak


def check_matrices(ctx: native.Context, rows, flags, keccak, keccak_randomness, row_begin=0, row_end=None):
    ctx.set_challenge(native.CHALLENGE_KECCAK, packing.cell_int(keccak_randomness))
    ctx.upload_table(native.TABLE_KECCAK, keccak)
    ctx.upload_columns(native.CIRCUIT_SIG, rows, flags=flags)
    return ctx.check(native.CIRCUIT_SIG, row_begin, rows.shape[1] if row_end is None else row_end, 0, 0)


def verify_circuit(witness: Witness, keccak_randomness: FQ, ctx: Optional[native.Context] = None) -> None:
    """Reference signature (sig_circuit.py:113); AssertionError names the first failing row."""
    ctx = ctx or native.default_context()
    rows, flags, keccak = pack_witness(witness)
    ff, _ = check_matrices(ctx, rows, flags, keccak, keccak_randomness)
    hit = native.first_failure(ff, native.CIRCUIT_SIG)
    if hit is not None:
        row, cid, cls, name = hit
        raise AssertionError(f"Constraints failed at row = {row}: {name}")
