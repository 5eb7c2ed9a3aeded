"""GPU parity of the CUDA state-circuit checker against the reference's verdicts
(tests/golden/state.npz), the CPU oracle's per-constraint result, a 2^16-row synthetic RW table
with seeded corruptions, row sharding with halos, and the host API."""
import numpy as np
import pytest

import golden_util
import oracle_lib
from zkevm_specs_b200 import native, synth
from zkevm_specs_b200 import state_circuit as sc

pytestmark = pytest.mark.gpu


def test_state_golden_and_oracle_parity():
    ctx = native.default_context()
    n = n_domain = 0
    for name, k, s, f, m, exp_row, exp_exc in golden_util.state_vectors():
        ff, fc = sc.check_matrices(ctx, s, f, m)
        off, ofc = oracle_lib.check_state(s, f, m)
        assert np.array_equal(ff, off) and np.array_equal(fc, ofc), f"{name}[{k}] differs from oracle"
        hit = native.first_failure(ff, native.CIRCUIT_STATE)
        got = (-1, "") if hit is None else (hit[0], oracle_lib.EXC_OF_CLASS[hit[2]])
        if got[1] == "NotImplementedError" and exp_exc != got[1]:
            assert got[0] == exp_row
            n_domain += 1
            continue
        assert got == (exp_row, exp_exc), f"{name}[{k}] cuda {got} reference {(exp_row, exp_exc)}"
        n += 1
    assert n > 300 and n_domain < 5


def test_state_synthetic_2e16_rows_and_corruptions_match_oracle():
    ctx = native.default_context()
    w = synth.state_rows(1 << 16, seed=3)
    S, F, M = w["rows"], w["flags"], w["mpt"]
    ff, fc = sc.check_matrices(ctx, S, F, M)
    assert (ff == native.PASS).all() and fc.sum() == 0
    rng = np.random.default_rng(33)
    detected = 0
    for t in range(48):
        s, m = S.copy(), M
        i = int(rng.integers(S.shape[1]))
        kind = t % 6
        if kind == 0:
            s[50, i, 0] ^= np.uint64(1 << int(rng.integers(8)))  # value
        elif kind == 1:
            s[0, i, 0] += np.uint64(rng.integers(1, 5))  # rw_counter
        elif kind == 2:
            s[4, i, 0] += np.uint64(1)  # address without its limbs
        elif kind == 3:
            s[:, [i, i + 1 if i + 1 < S.shape[1] else i - 1], :] = s[:, [i + 1 if i + 1 < S.shape[1] else i - 1, i], :]
        elif kind == 4:
            s[54, i, 0] += np.uint64(5)  # root
        else:
            m = M.copy(); m[8, int(rng.integers(M.shape[1])), 0] ^= np.uint64(1)  # an MPT value
        ff, fc = sc.check_matrices(ctx, s, F, m)
        off, ofc = oracle_lib.check_state(s, F, m)
        assert np.array_equal(ff, off) and np.array_equal(fc, ofc), f"corruption {t} kind {kind}"
        detected += bool((ff != native.PASS).any())
    assert detected >= 40


def test_state_sharded_rows_match_whole():
    ctx = native.default_context()
    w = synth.state_rows(4096, seed=9)
    S, F, M = w["rows"].copy(), w["flags"], w["mpt"]
    S[50, 3000, 0] ^= np.uint64(4)
    whole, _ = sc.check_matrices(ctx, S, F, M)
    n, half = S.shape[1], 2048
    # shard 0: rows [0, half) with halos n-1 (wrap) and half ; shard 1: rows [half, n) with halos half-1 and 0
    a_rows = np.ascontiguousarray(np.concatenate([S[:, n - 1:], S[:, : half + 1]], axis=1))
    a_flags = np.concatenate([F[n - 1:], F[: half + 1]])
    a, _ = sc.check_matrices(ctx, a_rows, a_flags, M, 1, half + 1, -1 % (1 << 64), 0)
    b_rows = np.ascontiguousarray(np.concatenate([S[:, half - 1:], S[:, :1]], axis=1))
    b_flags = np.concatenate([F[half - 1:], F[:1]])
    b, _ = sc.check_matrices(ctx, b_rows, b_flags, M, 1, n - half + 1, half - 1, 0)
    assert np.array_equal(np.minimum(a, b), whole)


def test_state_host_api_like_reference_tests():
    """tests/test_state_circuit.py:41-110 (ok) and :113-124 (bad key2 limbs) on our host API"""
    from zkevm_specs_b200.evm_circuit import RW, AccountFieldTag, CallContextFieldTag
    from zkevm_specs_b200.util import FQ, Word

    R_, W_ = RW.Read, RW.Write
    ops = [sc.StartOp(1, R_, 0), sc.StartOp(2, R_), sc.MemoryOp(1, R_, 1, 0, 0), sc.MemoryOp(2, W_, 1, 0, 42),
           sc.MemoryOp(3, R_, 1, 0, 42), sc.StackOp(4, W_, 1, 1022, Word(4321)), sc.StackOp(5, W_, 1, 1023, Word(533)),
           sc.StackOp(6, R_, 1, 1023, Word(533)),
           sc.StorageOp(7, R_, 1, 0x12345678, 0x1516, Word(789), Word(789)),
           sc.StorageOp(8, W_, 1, 0x12345678, 0x4959, Word(38491), Word(98765)),
           sc.CallContextOp(9, R_, 1, CallContextFieldTag.IsStatic, FQ(0)),
           sc.AccountOp(12, W_, 0x12345678, AccountFieldTag.Nonce, FQ(1), FQ(0)),
           sc.AccountOp(13, R_, 0x12345678, AccountFieldTag.Nonce, FQ(1), FQ(0))]
    rows = sc.assign_state_circuit(ops)
    tables = sc.Tables(sc.mpt_table_from_ops(ops))
    sc.verify_state_circuit(rows, tables)
    sc.check_state_row(rows[3], rows[2], rows[4], tables)
    bad = list(rows)
    bad[2] = bad[2]._replace(key2_limbs=(FQ(1),) * 10)
    with pytest.raises(AssertionError):
        sc.verify_state_circuit(bad, tables)
    with pytest.raises(Exception) as ei:  # MPT table without the storage update
        sc.verify_state_circuit(rows, sc.Tables(set()))
    assert type(ei.value).__name__ == "LookupUnsatFailure"
