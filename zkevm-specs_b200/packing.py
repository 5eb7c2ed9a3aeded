"""Column packer: Python witness/table rows -> the cell matrices of include/zkcheck.h
(uint64[n_cols][n_rows][4], canonical little-endian limbs).

Cell order per circuit/table follows the field order of the reference's row types, with a
Word expanded to (lo, hi):
  bytecode circuit  bytecode_circuit.Row        bytecode_circuit.py:15-27   12 cells
  bytecode table    BytecodeTableRow            evm_circuit/table.py:438-443  6 cells
  keccak table      KeccakTableRow              evm_circuit/table.py:511-515  5 cells
  rw table          RWTableRow                  evm_circuit/table.py:447-457 14 cells
Large synthetic witnesses skip the Python objects entirely and are produced as numpy
matrices by synth.py.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np

_MASK = 0xFFFFFFFFFFFFFFFF


def cell_int(x) -> int:
    """canonical integer of an FQ / Expression / IntEnum / int"""
    if hasattr(x, "expr"):
        return x.expr().n
    if hasattr(x, "n"):
        return x.n
    return int(x)


def matrix_from_ints(rows: Sequence[Sequence[int]], n_cols: int) -> np.ndarray:
    """rows of python ints (row-major) -> uint64[n_cols][n_rows][4]"""
    n_rows = len(rows)
    out = np.zeros((n_cols, n_rows, 4), dtype=np.uint64)
    if n_rows == 0:
        return out
    small = True
    for r in rows:
        assert len(r) == n_cols
        for v in r:
            if v > _MASK:
                small = False
                break
        if not small:
            break
    if small:
        out[:, :, 0] = np.array(rows, dtype=np.uint64).T
        return out
    for i, r in enumerate(rows):
        for c, v in enumerate(r):
            if v <= _MASK:
                out[c, i, 0] = v
            else:
                out[c, i, 0] = v & _MASK
                out[c, i, 1] = (v >> 64) & _MASK
                out[c, i, 2] = (v >> 128) & _MASK
                out[c, i, 3] = v >> 192
    return out


def int_to_cell(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & _MASK for i in range(4)], dtype=np.uint64)


def cell_to_int(c) -> int:
    return sum(int(c[i]) << (64 * i) for i in range(4))


# ---- per-type row flatteners ----------------------------------------------------------
def bytecode_circuit_row(r) -> List[int]:
    return [cell_int(r.q_first), cell_int(r.q_last), cell_int(r.hash.lo), cell_int(r.hash.hi),
            cell_int(r.tag), cell_int(r.index), cell_int(r.value), cell_int(r.is_code),
            cell_int(r.push_data_left), cell_int(r.value_rlc), cell_int(r.length),
            cell_int(r.push_data_size)]


def bytecode_table_row(r) -> List[int]:
    return [cell_int(r.bytecode_hash.lo), cell_int(r.bytecode_hash.hi), cell_int(r.field_tag),
            cell_int(r.index), cell_int(r.is_code), cell_int(r.value)]


def keccak_table_row(r) -> List[int]:
    return [cell_int(r.state_tag), cell_int(r.input_rlc), cell_int(r.input_len),
            cell_int(r.output.lo), cell_int(r.output.hi)]


def rw_table_row(r) -> List[int]:
    return [cell_int(r.rw_counter), cell_int(r.rw), cell_int(r.key0), cell_int(r.id),
            cell_int(r.address), cell_int(r.field_tag), cell_int(r.storage_key.lo),
            cell_int(r.storage_key.hi), cell_int(r.value.lo), cell_int(r.value.hi),
            cell_int(r.value_prev.lo), cell_int(r.value_prev.hi), cell_int(r.aux0.lo),
            cell_int(r.aux0.hi)]


def rw_table_flags(r) -> int:
    """bit0: value is a Word, bit1: value_prev is a Word (WordOrValue.is_word)"""
    return int(bool(getattr(r.value, "is_word", True))) | (int(bool(getattr(r.value_prev, "is_word", True))) << 1)


def pack(rows: Iterable, flatten, n_cols: int) -> np.ndarray:
    return matrix_from_ints([flatten(r) for r in rows], n_cols)


def copy_circuit_row(r) -> List[int]:
    return [cell_int(r.q_step), cell_int(r.is_first), cell_int(r.is_last), cell_int(r.id.lo), cell_int(r.id.hi),
            cell_int(r.tag), cell_int(r.addr), cell_int(r.src_addr_end), cell_int(r.bytes_left), cell_int(r.value),
            cell_int(r.rlc_acc), cell_int(r.is_code), cell_int(r.is_pad), cell_int(r.rw_counter),
            cell_int(r.rwc_inc_left), cell_int(r.is_memory), cell_int(r.is_bytecode), cell_int(r.is_tx_calldata),
            cell_int(r.is_tx_log), cell_int(r.is_rlc_acc)]


def tx_table_row(r) -> List[int]:
    return [cell_int(r.tx_id), cell_int(r.field_tag), cell_int(r.call_data_index_or_zero),
            cell_int(r.value.lo), cell_int(r.value.hi)]


def word_flag(x) -> int:
    """the WordOrValue.is_word type bit (a plain Word counts as a word)"""
    return int(bool(getattr(x, "is_word", True)))


def copy_table_row(r) -> List[int]:
    return [cell_int(r.is_first), cell_int(r.src_id.lo), cell_int(r.src_id.hi), cell_int(r.src_tag),
            cell_int(r.dst_id.lo), cell_int(r.dst_id.hi), cell_int(r.dst_tag), cell_int(r.src_addr),
            cell_int(r.src_addr_end), cell_int(r.dst_addr), cell_int(r.length), cell_int(r.rlc_acc),
            cell_int(r.rw_counter), cell_int(r.rwc_inc)]
