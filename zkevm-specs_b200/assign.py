"""Witness assignment on the device — host side of include/zkcheck.h "witness assignment" (csrc/assign.cu).

The reference builds circuit rows with Python object loops (assign_bytecode_circuit, bytecode_circuit.py:104-167;
op2row / assign_state_circuit, state_circuit.py:827-889; CopyCircuit.copy, evm_circuit/typing.py:1010-1147).  The
functions here hand the device the compact data those loops start from — raw code bytes, 15 operation cells, copy
events + the copied bytes — and leave the expanded witness resident for zk_check."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import native, packing
from .evm_circuit.spec import CopyDataTypeTag
from .util.arithmetic import FQ, Word
from .util.hash import keccak256


def is_code_bits(code: bytes) -> np.ndarray:
    """one bit per byte, LSB first: the byte is an opcode, not PUSH data (Bytecode.table_assignments, typing.py:390-427)"""
    n = len(code)
    flags = np.zeros(n, dtype=np.uint8)
    i = 0
    while i < n:
        flags[i] = 1
        b = code[i]
        i += 1 + (b - 0x5F if 0x60 <= b <= 0x7F else 0)
    return flags


def bytecode_src(codes: Sequence[bytes]) -> Dict[str, np.ndarray]:
    """arguments of Context.upload_bytecode_table_from_code / Context.assign_bytecode_circuit for these contracts"""
    flags = [is_code_bits(bytes(c)) for c in codes]
    offs = np.zeros(len(codes) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(c) for c in codes])
    code = np.frombuffer(b"".join(bytes(c) for c in codes), dtype=np.uint8).copy()
    bits = np.packbits(np.concatenate(flags), bitorder="little") if len(code) else np.zeros(1, dtype=np.uint8)
    hashes = np.zeros((len(codes), 4), dtype=np.uint64)
    for k, c in enumerate(codes):
        h = Word(int.from_bytes(keccak256(bytes(c)), "big"))
        hashes[k] = [h.lo.n & 0xFFFFFFFFFFFFFFFF, h.lo.n >> 64, h.hi.n & 0xFFFFFFFFFFFFFFFF, h.hi.n >> 64]
    return {"code": code, "is_code_bits": bits, "code_offsets": offs, "hashes": hashes}


def assign_bytecode_circuit(ctx: native.Context, k: int, codes: Sequence[bytes], keccak_randomness, stream: int = 0) -> None:
    """the reference's assign_bytecode_circuit(k, [UnrolledBytecode(code, Bytecode(code).table_assignments()) ...], r)
    as one device call; the rows stay resident as ZK_CIRCUIT_BYTECODE"""
    ctx.set_challenge(native.CHALLENGE_KECCAK, packing.cell_int(keccak_randomness))
    ctx.assign_bytecode_circuit(k, stream=stream, **bytecode_src(codes))


STATE_OP_COLS = list(range(8)) + list(range(50, 57))  # the 15 cells of a 57-cell state row an operation brings


def state_ops_from_rows(rows57: np.ndarray) -> np.ndarray:
    """the operation cells of packed state rows: uint64[57][n][4] -> uint64[15][n][4]"""
    return np.ascontiguousarray(rows57[STATE_OP_COLS])


def assign_state_circuit(ctx: native.Context, ops, flags=None, stream: int = 0) -> None:
    """op2row for every operation on the device.  `ops`: a list of state_circuit.Operation (roots from the mock MPT
    updates like state_circuit.assign_state_circuit) or an operation-cell matrix uint64[15][n][4]"""
    if isinstance(ops, np.ndarray):
        m = ops
    else:
        from . import state_circuit as sc
        rows = sc.assign_state_circuit(list(ops))
        cols, flags = sc.pack_rows(rows)  # host op2row only to read the 15 operation cells and the roots
        m = state_ops_from_rows(cols)
    ctx.assign_state_circuit(packing.pack_matrix(m), flags=flags, stream=stream)


def copy_event(src_id, src_tag, dst_id, dst_tag, src_addr: int, src_addr_end: int, dst_addr: int, copy_length: int,
               rw_counter: int, log_id: int = 0) -> List[int]:
    """one row of the events array of zk_assign_copy_circuit — the arguments of CopyCircuit.copy plus the rw_counter the
    event starts at"""
    def cell(x):
        if isinstance(x, Word) and not hasattr(x, "is_word"):
            return x.lo.n, x.hi.n, 1
        if hasattr(x, "is_word"):
            return packing.cell_int(x.lo), packing.cell_int(x.hi), int(x.is_word)
        return packing.cell_int(x), 0, 0

    s_lo, s_hi, s_w = cell(src_id)
    d_lo, d_hi, d_w = cell(dst_id)
    M = 0xFFFFFFFFFFFFFFFF
    return [int(src_tag) | (s_w << 8), int(dst_tag) | (d_w << 8), int(src_addr), int(src_addr_end), int(dst_addr), int(copy_length),
            int(log_id), int(rw_counter), s_lo & M, s_lo >> 64, s_hi & M, s_hi >> 64, d_lo & M, d_lo >> 64, d_hi & M, d_hi >> 64]


def assign_copy_circuit(ctx: native.Context, r, events: Sequence[Sequence[int]], data: bytes,
                        is_code: Optional[Sequence[int]] = None, stream: int = 0) -> None:
    """CopyCircuit.copy of every event as one device call: `events` from copy_event(), `data` the copied bytes of all
    events back to back (0 where the source is out of bounds), `is_code` one flag per data byte (events that touch
    bytecode)"""
    ctx.set_challenge(native.CHALLENGE_KECCAK, packing.cell_int(r))
    ev = np.array(events, dtype=np.uint64).reshape(-1, 16)
    bits = None
    if is_code is not None:
        bits = np.packbits(np.asarray(is_code, dtype=np.uint8), bitorder="little") if len(data) else np.zeros(1, dtype=np.uint8)
    ctx.assign_copy_circuit(ev, np.frombuffer(bytes(data), dtype=np.uint8), bits, stream=stream)
