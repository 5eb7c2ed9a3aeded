"""Lookup-table rows and the `Tables` container, host side.

Mirrors the row types of /root/reference/src/zkevm_specs/evm_circuit/table.py:404-576 and
`Tables` (:578-671).  Rows are plain frozen dataclasses used to BUILD tables; membership /
lookup semantics (table.py:864-884) are evaluated on the device from the packed matrices
(see packing.py for each table's cell order)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence, Set

import numpy as np

from ..util.arithmetic import FQ, Expression, Word, WordOrValue
from .spec import ExecutionState, FixedTableTag, SPEC, valid_opcodes


class LookupUnsatFailure(Exception):
    def __init__(self, table_name: str, inputs) -> None:
        self.inputs = inputs
        self.message = f"Lookup {table_name} is unsatisfied on inputs {inputs}"
        super().__init__(self.message)


class LookupAmbiguousFailure(Exception):
    def __init__(self, table_name: str, inputs, matched_rows=()) -> None:
        self.inputs = inputs
        self.message = f"Lookup {table_name} is ambiguous on inputs {inputs}"
        super().__init__(self.message)


@dataclass(frozen=True)
class FixedTableRow:
    tag: Expression
    value0: Expression
    value1: Expression = field(default=FQ(0))
    value2: Expression = field(default=FQ(0))


@dataclass(frozen=True)
class BlockTableRow:
    field_tag: Expression
    block_number_or_zero: Expression
    value: WordOrValue


@dataclass(frozen=True)
class TxTableRow:
    tx_id: Expression
    field_tag: Expression
    call_data_index_or_zero: Expression
    value: WordOrValue


@dataclass(frozen=True)
class WithdrawalTableRow:
    id: Expression
    validator_id: Expression
    address: Expression
    amount: Expression


@dataclass(frozen=True)
class BytecodeTableRow:
    bytecode_hash: Word
    field_tag: Expression
    index: Expression
    is_code: Expression
    value: Expression


@dataclass(frozen=True)
class RWTableRow:
    rw_counter: Expression
    rw: Expression
    key0: Expression  # Target
    id: Expression = field(default=FQ(0))
    address: Expression = field(default=FQ(0))
    field_tag: Expression = field(default=FQ(0))
    storage_key: Word = field(default=Word(0))
    value: WordOrValue = field(default=WordOrValue(FQ(0)))
    value_prev: WordOrValue = field(default=WordOrValue(FQ(0)))
    aux0: Word = field(default=Word(0))


@dataclass(frozen=True)
class MPTTableRow:
    address: Expression
    proof_type: Expression
    storage_key: Word
    root: Word
    root_prev: Word
    value: Word
    value_prev: Word


@dataclass(frozen=True)
class CopyCircuitRow:
    q_step: FQ
    is_first: FQ
    is_last: FQ
    id: WordOrValue
    tag: FQ
    addr: FQ
    src_addr_end: FQ
    bytes_left: FQ
    value: FQ
    rlc_acc: FQ
    is_code: FQ
    is_pad: FQ
    rw_counter: FQ
    rwc_inc_left: FQ
    is_memory: FQ
    is_bytecode: FQ
    is_tx_calldata: FQ
    is_tx_log: FQ
    is_rlc_acc: FQ


@dataclass(frozen=True)
class CopyTableRow:
    is_first: FQ
    src_id: WordOrValue
    src_tag: FQ
    dst_id: WordOrValue
    dst_tag: FQ
    src_addr: FQ
    src_addr_end: FQ
    dst_addr: FQ
    length: FQ
    rlc_acc: FQ
    rw_counter: FQ
    rwc_inc: FQ


@dataclass(frozen=True)
class KeccakTableRow:
    state_tag: FQ
    input_rlc: FQ
    input_len: FQ
    output: Word


_FIXED_MATRIX: Optional[np.ndarray] = None


def fixed_table_matrix() -> np.ndarray:
    """The whole fixed table (224,490 rows, reference table.py:37-103) as a cell matrix
    uint64[4][n][4], generated arithmetically (every entry fits one limb)."""
    global _FIXED_MATRIX
    if _FIXED_MATRIX is not None:
        return _FIXED_MATRIX
    T = FixedTableTag
    parts: List[np.ndarray] = []

    def block(tag, v0, v1=None, v2=None):
        v0 = np.asarray(v0, dtype=np.uint64)
        z = np.zeros_like(v0)
        parts.append(np.stack([np.full_like(v0, int(tag)), v0,
                               z if v1 is None else np.asarray(v1, dtype=np.uint64),
                               z if v2 is None else np.asarray(v2, dtype=np.uint64)]))

    for tag, n in ((T.Range5, 5), (T.Range16, 16), (T.Range32, 32), (T.Range64, 64), (T.Range256, 256),
                   (T.Range512, 512), (T.Range1024, 1024), (T.Range24_576, 24576)):
        block(tag, np.arange(n))
    b = np.arange(256, dtype=np.uint64)
    block(T.SignByte, b, (b >> np.uint64(7)) * np.uint64(0xFF))
    lhs, rhs = np.repeat(b, 256), np.tile(b, 256)
    block(T.BitwiseAnd, lhs, rhs, lhs & rhs)
    block(T.BitwiseOr, lhs, rhs, lhs | rhs)
    block(T.BitwiseXor, lhs, rhs, lhs ^ rhs)
    resp = [(int(s), o, a) for s in ExecutionState for o, a in s.responsible_opcode()]
    block(T.ResponsibleOpcode, [r[0] for r in resp], [r[1] for r in resp], [r[2] for r in resp])
    gas = [(o, c[2]) for name, c in SPEC["opcode_info"].items()
           for o in [SPEC["enums"]["Opcode"][name]] if not c[3] and c[2] > 0]
    block(T.OpcodeConstantGas, [g[0] for g in gas], [g[1] for g in gas])
    pre = SPEC["precompile_info_pairs"]
    block(T.PrecompileInfo, [p[0] for p in pre], [p[1] for p in pre], [p[2] for p in pre])
    small = np.concatenate(parts, axis=1)  # [4][n] single-limb values
    n = small.shape[1] + 256
    out = np.zeros((4, n, 4), dtype=np.uint64)
    out[:, : small.shape[1], 0] = small
    # Pow2: (value, 2^value lo-part, hi-part): 1 << value spans two limbs of a 128-bit half
    base = small.shape[1]
    for v in range(256):
        out[0, base + v, 0] = int(T.Pow2)
        out[1, base + v, 0] = v
        col = 2 if v < 128 else 3
        e = v if v < 128 else v - 128
        out[col, base + v, e // 64] = np.uint64(1 << (e % 64))
    _FIXED_MATRIX = out
    return out


class Tables:
    """The lookup tables one EVM/copy-circuit check runs against (reference table.py:578-625).
    Holds Python row sets exactly like the reference; `packing.pack_tables` turns them into
    cell matrices for libzkcheck."""

    def __init__(self, block_table, tx_table, withdrawal_table, bytecode_table, rw_table,
                 copy_circuit: Optional[Sequence[CopyCircuitRow]] = None,
                 keccak_table: Optional[Sequence[KeccakTableRow]] = None,
                 exp_circuit=None, sig_table=None, ecc_table=None) -> None:
        self.block_table = set(block_table)
        self.tx_table = set(tx_table)
        self.withdrawal_table = set(withdrawal_table)
        self.bytecode_table = set(bytecode_table)
        self.rw_table = set(r if isinstance(r, RWTableRow) else RWTableRow(*r) for r in rw_table)
        self.copy_table: Set[CopyTableRow] = set()
        self.keccak_table: Set[KeccakTableRow] = set()
        if copy_circuit is not None:
            self.copy_table = self._copy_circuit_to_table(copy_circuit)
        if keccak_table is not None:
            self.keccak_table = set(keccak_table)
        self.exp_circuit = exp_circuit
        self.sig_table = set(sig_table) if sig_table is not None else set()
        self.ecc_table = set(ecc_table) if ecc_table is not None else set()

    @staticmethod
    def _copy_circuit_to_table(copy_circuit: Sequence[CopyCircuitRow]) -> Set[CopyTableRow]:
        """One table row per copy event: fields of the first row and of the row after it
        (reference table.py:627-652)."""
        out = set()
        for i, first in enumerate(copy_circuit):
            if first.is_first == 1:
                assert i + 1 < len(copy_circuit), "Not enough rows in copy circuit"
                nxt = copy_circuit[i + 1]
                assert nxt.q_step == 0, "Invalid copy circuit"
                out.add(CopyTableRow(first.is_first, first.id, first.tag, nxt.id, nxt.tag, first.addr,
                                     first.src_addr_end, nxt.addr, first.bytes_left, first.rlc_acc,
                                     first.rw_counter, first.rwc_inc_left))
        return out
