"""Tx circuit, host side — the Fr parts of /root/reference/src/zkevm_specs/tx_circuit.py.

Same names and meaning as the reference: `Row` (:21-35), `KeccakTable` (:36-61), `SignVerifyChip`
(:160-243), `Witness` (:246-249), `verify_circuit(witness, MAX_TXS, MAX_CALLDATA_BYTES,
keccak_randomness)` (:253-289).  The ECDSA signature check itself (`ECDSAVerifyChip.verify`,
:147-158) is third-party curve math (eth_keys KeyAPI): the caller supplies an `ecdsa_chip` object
with `pub_key_x_bytes`, `pub_key_y_bytes`, `msg_hash_bytes` (little-endian, :128-130) and a
`verify(assert_msg)` method; its verdict travels to the device as a row flag so that the FIRST
failing constraint is still the one the reference raises first.  Everything else — the RLC of the
64 public-key bytes, the keccak-table membership, address / msg-hash equalities and the copy
constraints to the tx-table rows — is checked in one zk_check(ZK_CIRCUIT_TX)."""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Set, Tuple, Union

import numpy as np

from . import native, packing
from .util.arithmetic import FQ, RLC, Word, WordOrValue
from .util.hash import keccak256

# TxContextFieldTag values used for row offsets (evm_circuit/table.py:256-270: CallerAddress = 4,
# TxSignHash = 12; a tx occupies TxSignHash consecutive rows, tx_circuit.py:268-270)
TAG_CALLER_ADDRESS, TAG_TX_SIGN_HASH = 4, 12


class Row:
    """Tx circuit row (tx_circuit.py:21-35)"""

    def __init__(self, tx_id: FQ, tag: FQ, index: FQ, value: Union[FQ, Word]):
        self.tx_id, self.tag, self.index, self.value = tx_id, tag, index, WordOrValue(value)


class KeccakTable:
    """columns (is_enabled, input_rlc, input_len, output) — tx_circuit.py:36-61"""

    def __init__(self) -> None:
        self.table: Set[Tuple[FQ, FQ, FQ, Word]] = {(FQ(0), FQ(0), FQ(0), Word(0))}

    def add(self, input: bytes, keccak_randomness: FQ) -> None:
        self.table.add((FQ(1), RLC(bytes(reversed(input)), keccak_randomness, n_bytes=64).expr(), FQ(len(input)),
                        Word(keccak256(input))))


class SignVerifyChip:
    """tx_circuit.py:160-203 (construction); verification happens on the device"""

    def __init__(self, pub_key_hash: bytes, address: FQ, msg_hash: Word, ecdsa_chip) -> None:
        self.pub_key_hash, self.address, self.msg_hash, self.ecdsa_chip = pub_key_hash, address, msg_hash, ecdsa_chip
        self.pub_key_x_bytes = ecdsa_chip.pub_key_x_bytes
        self.pub_key_y_bytes = ecdsa_chip.pub_key_y_bytes
        self.msg_hash_bytes = ecdsa_chip.msg_hash_bytes


class Witness(NamedTuple):
    rows: List[Row]
    keccak_table: KeccakTable
    sign_verifications: List[SignVerifyChip]


def _word_cells(b: bytes) -> Tuple[int, int]:
    assert len(b) == 32
    return int.from_bytes(b[0:16], "little"), int.from_bytes(b[16:32], "little")


def pack_witness(witness: Witness, MAX_TXS: int):
    """-> (rows uint64[14][MAX_TXS][4], flags uint8[MAX_TXS], keccak uint64[5][k][4])"""
    c = packing.cell_int
    cells, flags = [], []
    for tx_index in range(MAX_TXS):
        chip = witness.sign_verifications[tx_index]
        base = tx_index * TAG_TX_SIGN_HASH
        caller, sign_hash = witness.rows[base + TAG_CALLER_ADDRESS - 1], witness.rows[base + TAG_TX_SIGN_HASH - 1]
        # tx_circuit.py:209-211: the chip's byte fields are copies of the ECDSA chip's
        assert chip.pub_key_x_bytes == chip.ecdsa_chip.pub_key_x_bytes
        assert chip.pub_key_y_bytes == chip.ecdsa_chip.pub_key_y_bytes
        assert chip.msg_hash_bytes == chip.ecdsa_chip.msg_hash_bytes
        ecdsa_failed = 0
        try:  # the reference calls ecdsa_chip.verify for every tx, padding ones included (:242)
            chip.ecdsa_chip.verify(f"Constraints failed for tx_index = {tx_index}")
        except AssertionError:
            ecdsa_failed = 1
        cells.append([c(chip.address), *_word_cells(bytes(chip.pub_key_x_bytes)), *_word_cells(bytes(chip.pub_key_y_bytes)),
                      *_word_cells(bytes(chip.pub_key_hash)), c(chip.msg_hash.lo), c(chip.msg_hash.hi),
                      *_word_cells(bytes(chip.msg_hash_bytes)), c(caller.value.lo), c(sign_hash.value.lo),
                      c(sign_hash.value.hi)])
        flags.append(int(caller.value.is_word) | (ecdsa_failed << 1))
    keccak = packing.matrix_from_ints([[c(a), c(b), c(l), c(o.lo), c(o.hi)] for a, b, l, o in witness.keccak_table.table], 5)
    return packing.matrix_from_ints(cells, 14), np.array(flags, dtype=np.uint8), keccak


def check_matrices(ctx: native.Context, rows, flags, keccak, keccak_randomness, row_begin=0, row_end=None):
    ctx.set_challenge(native.CHALLENGE_KECCAK, packing.cell_int(keccak_randomness))
    ctx.upload_table(native.TABLE_KECCAK, keccak)
    ctx.upload_columns(native.CIRCUIT_TX, rows, flags=flags)
    return ctx.check(native.CIRCUIT_TX, row_begin, rows.shape[1] if row_end is None else row_end, 0, 0)


def verify_circuit(witness: Witness, MAX_TXS: int, MAX_CALLDATA_BYTES: int, keccak_randomness: FQ,
                   ctx: Optional[native.Context] = None) -> None:
    """Reference signature (tx_circuit.py:253); raises AssertionError naming the first failing
    tx_index and constraint, like the reference's loop stops at its first failing assert."""
    ctx = ctx or native.default_context()
    rows, flags, keccak = pack_witness(witness, MAX_TXS)
    ff, _ = check_matrices(ctx, rows, flags, keccak, keccak_randomness)
    hit = native.first_failure(ff, native.CIRCUIT_TX)
    if hit is not None:
        row, cid, cls, name = hit
        raise AssertionError(f"Constraints failed for tx_index = {row}: {name}")
