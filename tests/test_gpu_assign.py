"""Device witness assignment (csrc/assign.cu, include/zkcheck.h "witness assignment") on the GPU: the resident rows,
downloaded, equal cell for cell what the host mirrors of the reference's builders produce (assign_bytecode_circuit
bytecode_circuit.py:104-167, op2row state_circuit.py:827-857, CopyCircuit.copy typing.py:1010-1147), and the circuit
checkers accept them where they stand."""
import numpy as np
import pytest

import assign_cases
from zkevm_specs_b200 import assign, native, packing, synth
from zkevm_specs_b200 import bytecode_circuit as bc

pytestmark = pytest.mark.gpu


def test_assign_bytecode_circuit_equals_host_builder_and_checks():
    ctx = native.default_context()
    n = 0
    for name, k, codes, exp in assign_cases.bytecode_cases():
        assign.assign_bytecode_circuit(ctx, k, codes, assign_cases.R)
        got, _ = ctx.download_columns(native.CIRCUIT_BYTECODE)
        assert np.array_equal(got, exp), (name, k, np.argwhere((got != exp).any(axis=2))[:5])
        if sum(len(c) + 1 for c in codes) <= (1 << k) - 1:  # every contract fits: the circuit is satisfiable
            ctx.upload_table(native.TABLE_PUSH, bc.pack_push_table(bc.assign_push_table()))
            ctx.upload_table(native.TABLE_KECCAK, bc.pack_keccak_table(bc.assign_keccak_table([bytes(c) for c in codes], assign_cases.R)))
            ff, _ = ctx.check(native.CIRCUIT_BYTECODE, 0, 1 << k, 0, native.FLAG_WRAP)
            assert (ff == native.PASS).all(), (name, k, native.first_failure(ff, native.CIRCUIT_BYTECODE))
        n += 1
    assert n >= 6


def test_assign_state_circuit_equals_op2row_and_checks():
    ctx = native.default_context()
    ops, flags, exp, mpt = assign_cases.state_case(1 << 14)
    assign.assign_state_circuit(ctx, ops, flags)
    got, fl = ctx.download_columns(native.CIRCUIT_STATE)
    assert np.array_equal(got, exp), np.argwhere((got != exp).any(axis=2))[:5]
    assert np.array_equal(fl, flags)
    ctx.upload_table(native.TABLE_MPT, mpt)
    ff, fc = ctx.check(native.CIRCUIT_STATE, 0, exp.shape[1], 0, native.FLAG_WRAP)
    assert (ff == native.PASS).all(), native.first_failure(ff, native.CIRCUIT_STATE)
    # same verdicts as the canonical upload of the host-built rows, also on a corrupted operation
    bad = ops.copy()
    bad[4, 9000, 0] ^= np.uint64(1 << 20)  # address
    assign.assign_state_circuit(ctx, bad, flags)
    ff_a, fc_a = ctx.check(native.CIRCUIT_STATE, 0, exp.shape[1], 0, native.FLAG_WRAP)
    ctx.upload_columns(native.CIRCUIT_STATE, ctx.download_columns(native.CIRCUIT_STATE)[0], flags=flags)
    ff_c, fc_c = ctx.check(native.CIRCUIT_STATE, 0, exp.shape[1], 0, native.FLAG_WRAP)
    assert np.array_equal(ff_a, ff_c) and np.array_equal(fc_a, fc_c) and (ff_a != native.PASS).any()
    with pytest.raises(native.NativeError):  # op.address.to_bytes(20, "little") overflows in the reference
        bad[4, 5, 3] = np.uint64(1)
        ctx.assign_state_circuit(packing.pack_matrix(bad, widths=[32] * 15), flags=flags)


def test_assign_copy_circuit_equals_host_copy():
    ctx = native.default_context()
    n = 0
    for name, events, data, code_flags, exp, exp_flags in assign_cases.copy_cases():
        assign.assign_copy_circuit(ctx, assign_cases.R, events, data, code_flags)
        got, fl = ctx.download_columns(native.CIRCUIT_COPY)
        assert got.shape == exp.shape, (name, got.shape, exp.shape)
        assert np.array_equal(got, exp), (name, np.argwhere((got != exp).any(axis=2))[:8])
        assert np.array_equal(fl, exp_flags), name
        n += 1
    assert n >= 7


def test_assign_full_size_and_check():
    """cfg-sized inputs: bytecode 2^17 rows (4 contracts of 32 KB: long Horner chains), copy 2^18 rows with its rw / tx
    tables, state 2^18 rows — assigned on the device, equal to the numpy generators' rows, accepted by the checkers"""
    ctx = native.default_context()
    b = synth.bytecode_circuit_rows(17, 4)
    ctx.set_challenge(native.CHALLENGE_KECCAK, b["r_int"])
    ctx.assign_bytecode_circuit(17, **assign.bytecode_src(b["codes"]))
    got, _ = ctx.download_columns(native.CIRCUIT_BYTECODE)
    assert np.array_equal(got, b["rows"])
    ctx.upload_table(native.TABLE_PUSH, b["push"])
    ctx.upload_table(native.TABLE_KECCAK, b["keccak"])
    ff, _ = ctx.check(native.CIRCUIT_BYTECODE, 0, 1 << 17, 0, native.FLAG_WRAP)
    assert (ff == native.PASS).all(), native.first_failure(ff, native.CIRCUIT_BYTECODE)

    w = synth.copy_events(128, 1024)
    ctx.set_challenge(native.CHALLENGE_KECCAK, w["r_int"])
    ctx.assign_copy_circuit(w["events"], w["data"])
    got, fl = ctx.download_columns(native.CIRCUIT_COPY)
    assert np.array_equal(got, w["copy"]) and np.array_equal(fl, w["copy_flags"])
    ctx.upload_table(native.TABLE_RW, w["rw"], flags=w["rw_flags"])
    ctx.upload_table(native.TABLE_BYTECODE, w["bytecode"])
    ctx.upload_table(native.TABLE_TX, w["tx"], flags=w["tx_flags"])
    ff, _ = ctx.check(native.CIRCUIT_COPY, 0, got.shape[1], 0, native.FLAG_WRAP)
    assert (ff == native.PASS).all(), native.first_failure(ff, native.CIRCUIT_COPY)

    s = synth.state_rows(1 << 18, seed=3)
    ctx.assign_state_circuit(packing.pack_matrix(assign.state_ops_from_rows(s["rows"])), flags=s["flags"])
    got, fl = ctx.download_columns(native.CIRCUIT_STATE)
    assert np.array_equal(got, s["rows"]) and np.array_equal(fl, s["flags"])
    ctx.upload_table(native.TABLE_MPT, s["mpt"])
    ff, _ = ctx.check(native.CIRCUIT_STATE, 0, 1 << 18, 0, native.FLAG_WRAP)
    assert (ff == native.PASS).all(), native.first_failure(ff, native.CIRCUIT_STATE)
