"""Device witness assignment (csrc/assign.cu), per-thread bodies run serially on the CPU (tests/emu): the rows equal,
cell for cell, what the host mirrors of the reference's builders produce (assign_bytecode_circuit, op2row,
CopyCircuit.copy).  The GPU run of the same cases is tests/test_gpu_assign.py."""
import numpy as np

import assign_cases
import emu_lib
from zkevm_specs_b200 import assign


def test_emu_assign_bytecode_circuit_equals_host_builder():
    n = 0
    for name, k, codes, exp in assign_cases.bytecode_cases():
        got = emu_lib.assign_bytecode(k, r=assign_cases.R.n, **assign.bytecode_src(codes))
        assert np.array_equal(got, exp), (name, k, np.argwhere((got != exp).any(axis=2))[:5])
        n += 1
    assert n >= 6


def test_emu_assign_state_rows_equal_op2row():
    ops, flags, exp, _ = assign_cases.state_case(4096)
    got = emu_lib.assign_state(ops)
    assert np.array_equal(got, exp), np.argwhere((got != exp).any(axis=2))[:5]


def test_emu_assign_copy_circuit_equals_host_copy():
    n = 0
    for name, events, data, code_flags, exp, exp_flags in assign_cases.copy_cases():
        bits = None if code_flags is None else np.packbits(np.asarray(code_flags, dtype=np.uint8), bitorder="little")
        got, fl = emu_lib.assign_copy(events, np.frombuffer(data, dtype=np.uint8), bits, assign_cases.R.n)
        assert got.shape == exp.shape, (name, got.shape, exp.shape)
        assert np.array_equal(got, exp), (name, np.argwhere((got != exp).any(axis=2))[:8])
        assert np.array_equal(fl, exp_flags), name
        n += 1
    assert n >= 7
