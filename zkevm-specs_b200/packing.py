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


def block_table_row(r) -> List[int]:
    """BlockTableRow (evm_circuit/table.py:413-417): field_tag, block_number_or_zero, value (lo, hi)"""
    return [cell_int(r.field_tag), cell_int(r.block_number_or_zero), cell_int(r.value.lo), cell_int(r.value.hi)]


def word_flag(x) -> int:
    """the WordOrValue.is_word type bit (a plain Word counts as a word)"""
    return int(bool(getattr(x, "is_word", True)))


def copy_table_row(r) -> List[int]:
    return [cell_int(r.is_first), cell_int(r.src_id.lo), cell_int(r.src_id.hi), cell_int(r.src_tag),
            cell_int(r.dst_id.lo), cell_int(r.dst_id.hi), cell_int(r.dst_tag), cell_int(r.src_addr),
            cell_int(r.src_addr_end), cell_int(r.dst_addr), cell_int(r.length), cell_int(r.rlc_acc),
            cell_int(r.rw_counter), cell_int(r.rwc_inc)]


# ---------------------------------------------------------------------------------------------
# Packed columns (include/zkcheck.h "packed columns"): each column at the smallest width that
# holds its values.  The reference's cells are Python ints, most of them bytes, flags, tags and
# counters (evm_circuit/table.py:405-535); shipping them as 32-byte cells wastes ~5x of both the
# PCIe and the HBM bytes of a check.
# ---------------------------------------------------------------------------------------------
class PackedMatrix:
    """ONE host buffer + per-column (offset, width) — the argument of zk_upload_*_packed.

    `buf` is a uint8 array (pin it for asynchronous copies); column c holds n_rows little-endian
    unsigned integers of widths[c] bytes at byte offsets[c] (a multiple of 32), widths[c] in
    {1, 2, 4, 8, 16, 32}; width 0 = constant column stored once as one 32-byte cell."""

    def __init__(self, buf: np.ndarray, offsets: np.ndarray, widths: np.ndarray, n_rows: int) -> None:
        self.buf, self.offsets, self.widths, self.n_rows = buf, offsets, widths, n_rows

    @property
    def n_cols(self) -> int:
        return len(self.widths)

    @property
    def nbytes(self) -> int:
        return int(self.buf.nbytes)

    def column(self, c: int) -> np.ndarray:
        """column c widened back to uint64[n_rows][4] (tests)"""
        w, off = int(self.widths[c]), int(self.offsets[c])
        out = np.zeros((self.n_rows, 4), dtype=np.uint64)
        if w == 0:
            out[:] = self.buf[off:off + 32].view(np.uint64)
        elif w >= 8:
            out[:, :w // 8] = self.buf[off:off + w * self.n_rows].view(np.uint64).reshape(self.n_rows, w // 8)
        else:
            dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[w]
            out[:, 0] = self.buf[off:off + w * self.n_rows].view(dt)
        return out

    def unpack(self) -> np.ndarray:
        return np.stack([self.column(c) for c in range(self.n_cols)])


def column_width(col: np.ndarray) -> int:
    """smallest of 0/1/2/4/8/16/32 bytes that stores every cell of a uint64[n_rows][4] column"""
    if col.shape[0] == 0:
        return 32
    if (col == col[0]).all():
        return 0
    if col[:, 2].any() or col[:, 3].any():
        return 32
    if col[:, 1].any():
        return 16
    m = int(col[:, 0].max())
    return 1 if m < 1 << 8 else 2 if m < 1 << 16 else 4 if m < 1 << 32 else 8


def _value_width(cell: np.ndarray) -> int:
    if cell[2] or cell[3]:
        return 32
    if cell[1]:
        return 16
    m = int(cell[0])
    return 1 if m < 1 << 8 else 2 if m < 1 << 16 else 4 if m < 1 << 32 else 8


# Data-independent widths by column TYPE (what a packer that never looks at the values would use;
# bench.py packs with these so that the measured bytes do not depend on the synthetic trace having
# one contract or small counters).  Field order = the reference row types, Word = (lo, hi).
TYPE_WIDTHS = {
    # StepState, evm_circuit/step.py:16-44: state, rwc, call_id, is_root, is_create, code_hash lo/hi,
    # pc, sp, gas_left, memory_word_size, reversible_write_counter, log_id
    "evm_steps": [2, 8, 8, 1, 1, 16, 16, 8, 2, 8, 8, 8, 8],
    # RWTableRow, evm_circuit/table.py:447-457: rw_counter, rw, key0 (tag), key1 (id), key2 (address, up
    # to 160 bits), key3 (field tag), key4 lo/hi (storage key), value lo/hi, value_prev lo/hi, aux lo/hi
    "rw_table": [8, 1, 1, 8, 32, 1, 16, 16, 16, 16, 16, 16, 16, 16],
    # BytecodeTableRow, table.py:438-443: hash lo/hi, tag, index, is_code, value (a byte; the Header row
    # holds the code length)
    "bytecode_table": [16, 16, 1, 4, 1, 4],
}


def pack_matrix(matrix: np.ndarray, widths: Sequence[int] | None = None,
                min_widths: Sequence[int] | None = None) -> PackedMatrix:
    """uint64[n_cols][n_rows][4] -> PackedMatrix.
    widths=None, min_widths=None : every column at its measured minimal width (0 = constant column)
    min_widths                   : at least the given (type) width per column, wider if a value
                                   needs it (corrupted witnesses), never a constant column
    widths                       : exactly these; a value that does not fit is an error."""
    m = np.ascontiguousarray(matrix, dtype=np.uint64)
    n_cols, n_rows = m.shape[0], m.shape[1]
    if min_widths is not None:
        assert widths is None and len(min_widths) == n_cols
        ws = [max(int(t), column_width(m[c]) or _value_width(m[c][0])) if n_rows else int(t)
              for c, t in enumerate(min_widths)]
    else:
        ws = [column_width(m[c]) for c in range(n_cols)] if widths is None else [int(w) for w in widths]
    offsets, total = [], 0
    for w in ws:
        offsets.append(total)
        total += ((w * n_rows if w else 32) + 31) // 32 * 32
    buf = np.zeros(max(total, 32), dtype=np.uint8)
    for c, (w, off) in enumerate(zip(ws, offsets)):
        col = m[c]
        if n_rows == 0:
            continue
        if widths is not None:
            assert w in (0, 1, 2, 4, 8, 16, 32), f"bad width {w}"
            need = column_width(col[:1]) if w else column_width(col)  # w == 0 needs a constant column
            if w:
                need = max(column_width(col), _value_width(col[0]))
            assert (w == 0 and need == 0) or (w and need <= w), f"column {c} does not fit {w} bytes"
        if w == 0:
            buf[off:off + 32] = col[0].view(np.uint8)
        elif w >= 8:
            buf[off:off + w * n_rows] = np.ascontiguousarray(col[:, :w // 8]).view(np.uint8).reshape(-1)
        else:
            dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[w]
            buf[off:off + w * n_rows] = col[:, 0].astype(dt).view(np.uint8)
    return PackedMatrix(buf, np.asarray(offsets, dtype=np.uint64), np.asarray(ws, dtype=np.uint8), n_rows)


def bytecode_src_from_table(table: np.ndarray):
    """Inverse of Bytecode.table_assignments for a REGULAR bytecode-table matrix (runs
    [Header, Byte 0, Byte 1, ...] per contract): returns the arguments of
    Context.upload_bytecode_table_from_code, or None if the table is not of that form."""
    t = np.ascontiguousarray(table, dtype=np.uint64)
    n = t.shape[1]
    if n == 0:
        return None
    small = not t[2:, :, 1:].any()
    tag, index, is_code, value = t[2, :, 0], t[3, :, 0], t[4, :, 0], t[5, :, 0]
    heads = np.nonzero(tag == 1)[0]
    if not small or len(heads) == 0 or heads[0] != 0 or t[0:2, :, 2:].any():
        return None
    ends = np.append(heads[1:], n)
    code, bits, offs, hashes = [], [], [0], []
    for h, e in zip(heads, ends):
        ln = int(e - h - 1)
        body = slice(h + 1, e)
        ok = (int(value[h]) == ln and int(index[h]) == 0 and int(is_code[h]) == 0 and (tag[body] == 2).all()
              and (index[body] == np.arange(ln, dtype=np.uint64)).all() and (value[body] < 256).all()
              and (is_code[body] < 2).all() and (t[0:2, h:e, :] == t[0:2, h:h + 1, :]).all())
        if not ok:
            return None
        code.append(value[body].astype(np.uint8))
        bits.append(is_code[body].astype(np.uint8))
        offs.append(offs[-1] + ln)
        hashes.append([t[0, h, 0], t[0, h, 1], t[1, h, 0], t[1, h, 1]])
    code = np.concatenate(code) if code else np.zeros(0, dtype=np.uint8)
    bits = np.packbits(np.concatenate(bits), bitorder="little") if len(code) else np.zeros(1, dtype=np.uint8)
    return {"code": code, "is_code_bits": bits, "code_offsets": np.array(offs, dtype=np.uint64),
            "hashes": np.array(hashes, dtype=np.uint64)}
