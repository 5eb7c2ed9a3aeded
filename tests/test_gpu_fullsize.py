"""BASELINE.json's configurations at their FULL sizes on the device, against the oracle run at
the same size (the C oracle finishes each in seconds), with planted corruptions, plus the
size-independent properties the path offers: row-sharding linearity (min / sum over shards ==
whole), packed == canonical storage.
  cfg2: evm_circuit ADD/SUB/MUL/DIV/MOD trace, 2^16 steps     cfg3: state_circuit, 2^18 rows
  cfg4: copy_circuit, 2^20 rows (SHA3 + CALLDATACOPY events)  bench: evm trace, 2^20 steps"""
import numpy as np
import pytest

import oracle_lib
from zkevm_specs_b200 import native, packing, synth
from zkevm_specs_b200 import state_circuit as sc
from zkevm_specs_b200.evm_circuit import main as evm_main
from zkevm_specs_b200.evm_circuit.table import fixed_table_matrix

pytestmark = pytest.mark.gpu
PASS = native.PASS


def _evm_upload(ctx, S, B, R):
    evm_main.upload_fixed_table(ctx)
    ctx.upload_table(native.TABLE_COPY, np.zeros((14, 0, 4), dtype=np.uint64))
    ctx.upload_table(native.TABLE_KECCAK, np.zeros((5, 0, 4), dtype=np.uint64))
    ctx.upload_table(native.TABLE_BYTECODE, B)
    ctx.upload_table(native.TABLE_RW, R)
    ctx.upload_columns(native.CIRCUIT_EVM, S)


def _shards(ctx, circuit, n, k, flags):
    """min of first_fail / sum of fail_count over k contiguous row ranges of the resident matrix"""
    ff = np.full(ctx.n_constraints(circuit), PASS, dtype=np.uint32)
    fc = np.zeros(ctx.n_constraints(circuit), dtype=np.uint64)
    edges = [n * i // k for i in range(k + 1)]
    for a, b in zip(edges, edges[1:]):
        f, c = ctx.check(circuit, a, b, 0, flags)
        ff, fc = np.minimum(ff, f), fc + c
    return ff, fc


def test_cfg2_evm_2e16_steps_full_size_vs_oracle():
    ctx = native.default_context()
    fixed = fixed_table_matrix()
    w = synth.evm_trace(1 << 14, seed=2)
    S, B, R = w["steps"].copy(), w["bytecode"].copy(), w["rw"].copy()
    assert S.shape[1] == (1 << 16) + 1
    _evm_upload(ctx, S, B, R)
    n = S.shape[1] - 1
    ff, fc = ctx.check(native.CIRCUIT_EVM, 0, n, 0, 0)
    assert (ff == PASS).all() and fc.sum() == 0
    # the negative set of SURVEY.md 8(d) cfg2: one operand limb / one gas_left / one rwc in 64 random steps
    rng = np.random.default_rng(202)
    for t in range(64):
        i = int(rng.integers(1, n))
        if t % 3 == 0:
            R[8 + int(rng.integers(2)), int(rng.integers(R.shape[1])), int(rng.integers(2))] ^= np.uint64(1 << int(rng.integers(64)))
        elif t % 3 == 1:
            S[9, i, 0] += np.uint64(1)
        else:
            S[1, i, 0] += np.uint64(1 + int(rng.integers(3)))
    _evm_upload(ctx, S, B, R)
    ff, fc = ctx.check(native.CIRCUIT_EVM, 0, n, 0, 0)
    off, ofc = oracle_lib.check_evm(S, B, R, fixed)
    assert np.array_equal(ff, off) and np.array_equal(fc, ofc)
    assert fc.sum() >= 64
    sff, sfc = _shards(ctx, native.CIRCUIT_EVM, n, 7, 0)
    assert np.array_equal(sff, ff) and np.array_equal(sfc, fc)
    ctx.packed_uploads = "min"
    try:
        _evm_upload(ctx, S, B, R)
        pff, pfc = ctx.check(native.CIRCUIT_EVM, 0, n, 0, 0)
    finally:
        ctx.packed_uploads = None
    assert np.array_equal(pff, ff) and np.array_equal(pfc, fc)


def test_cfg3_state_2e18_rows_full_size_vs_oracle():
    ctx = native.default_context()
    w = synth.state_rows(1 << 18, seed=3)
    S, F, M = w["rows"].copy(), w["flags"], w["mpt"]
    ff, fc = sc.check_matrices(ctx, S, F, M)
    assert (ff == PASS).all() and fc.sum() == 0
    rng = np.random.default_rng(303)
    for t in range(48):
        i = int(rng.integers(S.shape[1]))
        col = [50, 0, 4, 54, 1, 52][t % 6]  # value, rw_counter, address, root, is_write, initial value
        S[col, i, 0] += np.uint64(1 + int(rng.integers(3)))
    ff, fc = sc.check_matrices(ctx, S, F, M)
    off, ofc = oracle_lib.check_state(S, F, M)
    assert np.array_equal(ff, off) and np.array_equal(fc, ofc)
    assert fc.sum() >= 40
    sff, sfc = _shards(ctx, native.CIRCUIT_STATE, S.shape[1], 5, native.FLAG_WRAP)
    assert np.array_equal(sff, ff) and np.array_equal(sfc, fc)
    ctx.packed_uploads = "min"
    try:
        pff, pfc = sc.check_matrices(ctx, S, F, M)
    finally:
        ctx.packed_uploads = None
    assert np.array_equal(pff, ff) and np.array_equal(pfc, fc)


def test_cfg4_copy_2e20_rows_full_size_vs_oracle():
    ctx = native.default_context()
    w = synth.copy_events(512, 1024, seed=4)
    assert w["copy"].shape[1] == 1 << 20
    r = w["r"]

    def run(v):
        ctx.set_challenge(native.CHALLENGE_KECCAK, packing.cell_to_int(r))
        ctx.upload_table(native.TABLE_RW, v["rw"], flags=v["rw_flags"])
        ctx.upload_table(native.TABLE_BYTECODE, v["bytecode"])
        ctx.upload_table(native.TABLE_TX, v["tx"], flags=v["tx_flags"])
        ctx.upload_columns(native.CIRCUIT_COPY, v["copy"], flags=v["copy_flags"])
        return ctx.check(native.CIRCUIT_COPY, 0, v["copy"].shape[1], 0, native.FLAG_WRAP)

    ff, fc = run(w)
    assert (ff == PASS).all() and fc.sum() == 0
    v = dict(w)
    v["copy"], v["rw"], v["tx"] = w["copy"].copy(), w["rw"].copy(), w["tx"].copy()
    rng = np.random.default_rng(404)
    for t in range(40):
        if t % 4 == 0:
            v["copy"][9, int(rng.integers(v["copy"].shape[1])), 0] ^= np.uint64(1)   # value
        elif t % 4 == 1:
            v["rw"][8, int(rng.integers(v["rw"].shape[1])), 0] ^= np.uint64(2)       # a memory byte
        elif t % 4 == 2:
            v["tx"][3, int(rng.integers(v["tx"].shape[1])), 0] ^= np.uint64(4)       # a calldata byte
        else:
            v["copy"][13, int(rng.integers(v["copy"].shape[1])), 0] += np.uint64(1)  # rwc_inc_left
    ff, fc = run(v)
    off, ofc = oracle_lib.check_copy(v, r)
    assert np.array_equal(ff, off) and np.array_equal(fc, ofc)
    assert fc.sum() >= 30
    sff, sfc = _shards(ctx, native.CIRCUIT_COPY, v["copy"].shape[1], 8, native.FLAG_WRAP)
    assert np.array_equal(sff, ff) and np.array_equal(sfc, fc)
    ctx.packed_uploads = "min"
    try:
        pff, pfc = run(v)
    finally:
        ctx.packed_uploads = None
    assert np.array_equal(pff, ff) and np.array_equal(pfc, fc)


def test_bench_workload_2e20_steps_packed_type_widths_properties():
    """bench.py's workload and storage format at full size: the valid trace passes; with planted
    corruptions the verdict is independent of storage (packed == canonical) and of sharding; every
    failing row is one of the planted steps or its predecessor (locality: rotation {cur, next})."""
    ctx = native.default_context()
    w = synth.evm_trace(1 << 18, seed=2)
    S, B, R = w["steps"].copy(), w["bytecode"], w["rw"].copy()
    n = S.shape[1] - 1
    assert n == 1 << 20

    def upload(packed):
        evm_main.upload_fixed_table(ctx)
        ctx.upload_table(native.TABLE_COPY, np.zeros((14, 0, 4), dtype=np.uint64))
        ctx.upload_table(native.TABLE_KECCAK, np.zeros((5, 0, 4), dtype=np.uint64))
        if packed:
            ctx.upload_table_packed(native.TABLE_BYTECODE, packing.pack_matrix(B, min_widths=packing.TYPE_WIDTHS["bytecode_table"]))
            ctx.upload_table_packed(native.TABLE_RW, packing.pack_matrix(R, min_widths=packing.TYPE_WIDTHS["rw_table"]))
            ctx.upload_columns_packed(native.CIRCUIT_EVM, packing.pack_matrix(S, min_widths=packing.TYPE_WIDTHS["evm_steps"]))
        else:
            ctx.upload_table(native.TABLE_BYTECODE, B)
            ctx.upload_table(native.TABLE_RW, R)
            ctx.upload_columns(native.CIRCUIT_EVM, S)

    upload(True)
    ff, fc = ctx.check(native.CIRCUIT_EVM, 0, n, 0, 0)
    assert (ff == PASS).all() and fc.sum() == 0
    rng = np.random.default_rng(505)
    planted = sorted(int(x) for x in rng.choice(np.arange(8, n - 8), size=32, replace=False))
    for t, i in enumerate(planted):
        S[9 if t % 2 else 7, i, 0] += np.uint64(1)  # gas_left / program_counter of step i
    upload(True)
    ff, fc = ctx.check(native.CIRCUIT_EVM, 0, n, 0, 0)
    bad = ff[ff != PASS]
    assert len(bad) and fc.sum() >= 32
    assert set(int(x) for x in bad) <= set(planted) | set(i - 1 for i in planted)
    assert int(bad.min()) == planted[0] - 1
    sff, sfc = _shards(ctx, native.CIRCUIT_EVM, n, 8, 0)
    assert np.array_equal(sff, ff) and np.array_equal(sfc, fc)
    upload(False)
    cff, cfc = ctx.check(native.CIRCUIT_EVM, 0, n, 0, 0)
    assert np.array_equal(cff, ff) and np.array_equal(cfc, fc)
