"""Pins the CPU oracle's EVM-circuit restatement (verify_step + ADD/SUB, MUL/DIV/MOD, PUSH,
POP gadgets) against vectors produced by the reference's own verify_step loop
(tests/golden/evm.npz): same first failing step, same exception class."""
import golden_util
import oracle_lib
from zkevm_specs_b200.evm_circuit.table import fixed_table_matrix


def test_oracle_evm_matches_reference_golden():
    fixed = fixed_table_matrix()
    classes = oracle_lib.constraint_classes(3)
    n = n_fail = n_unsupported = 0
    kinds = set()
    for name, k, s, b, r, flags, exp_row, exp_exc in golden_util.evm_vectors():
        ff, fc = oracle_lib.check_evm(s, b, r, fixed, flags=flags)
        row, exc = oracle_lib.first_failure(ff, classes)
        if exc == "NotImplementedError" and exp_exc != exc:
            # declared limits of this build (EV_UNSUPPORTED_STATE / EV_MUL_WITNESS_DOMAIN): reported
            # loudly at the SAME step the reference fails on, never silently passed
            assert row == exp_row, f"{name}[{k}]"
            n_unsupported += 1
            continue
        if exc == "ValueError" and exp_exc == "OverflowError":
            exc = "OverflowError"  # both are the ZK_ERR_VALUE class (Python runtime errors)
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert n > 250 and n_fail > 150 and n_unsupported < 12
    assert {"AssertionError", "LookupUnsatFailure", "LookupAmbiguousFailure"} <= kinds


def test_oracle_evm_sha3_calldatacopy_matches_reference_golden():
    fixed = fixed_table_matrix()
    classes = oracle_lib.constraint_classes(3)
    n = n_fail = n_unsupported = 0
    kinds = set()
    for name, k, w, exp_row, exp_exc in golden_util.evm2_vectors():
        ff, fc = oracle_lib.check_evm_x(w, fixed)
        row, exc = oracle_lib.first_failure(ff, classes)
        if exc == "NotImplementedError" and exp_exc != exc:
            assert row == exp_row
            n_unsupported += 1
            continue
        if exc == "ValueError" and exp_exc == "OverflowError":
            exc = "OverflowError"
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert n > 380 and n_fail > 250 and n_unsupported == 0
    assert {"AssertionError", "LookupUnsatFailure", "ConstraintUnsatFailure"} <= kinds


def test_oracle_evm_stop_matches_reference_golden():
    """STOP in the root call and in an internal call (tests/evm/test_stop.py): 12 restore-context lookups"""
    fixed = fixed_table_matrix()
    classes = oracle_lib.constraint_classes(3)
    n = n_fail = 0
    kinds = set()
    for name, k, w, exp_row, exp_exc in golden_util.evm3_vectors():
        ff, fc = oracle_lib.check_evm_x(w, fixed)
        row, exc = oracle_lib.first_failure(ff, classes)
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert n > 550 and n_fail > 400
    assert {"AssertionError", "LookupUnsatFailure", "LookupAmbiguousFailure"} <= kinds, kinds


def test_oracle_evm_memory_matches_reference_golden():
    """MLOAD / MSTORE / MSTORE8 (tests/evm/test_memory.py): 1 or 32 memory lookups, memory expansion"""
    fixed = fixed_table_matrix()
    classes = oracle_lib.constraint_classes(3)
    n = n_fail = 0
    kinds = set()
    for name, k, w, exp_row, exp_exc in golden_util.evm4_vectors():
        ff, fc = oracle_lib.check_evm_x(w, fixed)
        row, exc = oracle_lib.first_failure(ff, classes)
        if exc == "ValueError" and exp_exc == "OverflowError":
            exc = "OverflowError"
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert n > 600 and n_fail > 450
    assert {"AssertionError", "LookupUnsatFailure", "LookupAmbiguousFailure", "ConstraintUnsatFailure"} <= kinds, kinds


def test_oracle_evm_simple_gadgets_match_reference_golden():
    """MSIZE, GAS, ISZERO, CMP (LT/GT/EQ), JUMP, JUMPI (tests/evm/test_{msize,gas,iszero,comparator,jump,jumpi}.py)
    and CALLER, CALLVALUE, CALLDATASIZE, ADDRESS, RETURNDATASIZE, CODESIZE (their tests/evm files)"""
    fixed = fixed_table_matrix()
    classes = oracle_lib.constraint_classes(3)
    n = n_fail = n_unsupported = 0
    kinds = set()
    import itertools

    for name, k, w, exp_row, exp_exc in itertools.chain(golden_util.evm5_vectors(), golden_util.evm6_vectors(), golden_util.evm7_vectors(), golden_util.evm8_vectors(), golden_util.evm9_vectors(), golden_util.evm10_vectors(), golden_util.evm12_vectors(), golden_util.evm13_vectors(), golden_util.evm14_vectors(), golden_util.evm15_vectors(), golden_util.evm16_vectors(), golden_util.evm17_vectors(), golden_util.evm18_vectors(), golden_util.evm19_vectors(), golden_util.evm20_vectors(), golden_util.evm21_vectors(), golden_util.evm22_vectors(), golden_util.evm23_vectors(), golden_util.evm24_vectors(), golden_util.evm25_vectors(), golden_util.evm26_vectors(), golden_util.evm27_vectors(), golden_util.evm28_vectors()):
        ff, fc = oracle_lib.check_evm_x(w, fixed)
        row, exc = oracle_lib.first_failure(ff, classes)
        if exc == "ValueError" and exp_exc in ("OverflowError", "UnboundLocalError", "AttributeError", "TypeError"):
            exc = exp_exc  # one "Python runtime error" class (include/zkcheck.h ZK_ERR_VALUE)
        if exc == "NotImplementedError" and exp_exc != exc:
            # EV_AR_WITNESS_DOMAIN: ADDMOD / MULMOD / SDIV / SMOD with a stack word half >= 2^128, reported at the SAME
            # step the reference fails on
            assert row == exp_row, f"{name}[{k}]"
            n_unsupported += 1
            continue
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert n > 5700 and n_fail > 4200 and n_unsupported <= 150
    assert {"AssertionError", "LookupUnsatFailure", "LookupAmbiguousFailure"} <= kinds, kinds


def test_oracle_evm_begin_end_tx_end_block_match_reference_golden():
    """BeginTx / EndTx / EndBlock (tests/evm/test_begin_tx.py, test_end_tx.py, test_end_block.py scenarios)"""
    fixed = fixed_table_matrix()
    classes = oracle_lib.constraint_classes(3)
    n = n_fail = 0
    kinds = set()
    bad = []
    for name, k, w, exp_row, exp_exc in golden_util.evm11_vectors():
        ff, fc = oracle_lib.check_evm_x(w, fixed)
        row, exc = oracle_lib.first_failure(ff, classes)
        if exc == "ValueError" and exp_exc == "OverflowError":
            exc = "OverflowError"
        if (row, exc) != (exp_row, exp_exc):
            bad.append(f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}")
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert not bad, f"{len(bad)} of {n} differ: " + "; ".join(bad[:12])
    assert n > 2500 and n_fail > 1000
    assert {"AssertionError", "LookupUnsatFailure", "LookupAmbiguousFailure", "ConstraintUnsatFailure", "OverflowError"} <= kinds


def test_oracle_and_emu_accept_whole_block_trace_and_agree_on_corruptions():
    """synth.block_trace (BeginTx .. STOP, EndTx per transaction, EndBlock; validated on the reference by gen_golden.py
    synth): the oracle and the emulated gate programs (positional head + tail rw table, and hash paths) accept it with the
    first / last step flags and agree array for array on corrupted copies"""
    import numpy as np

    import emu_lib
    from zkevm_specs_b200 import synth

    fixed = fixed_table_matrix()
    w = synth.block_trace(12, 5, 4, seed=9)
    n = oracle_lib.lib().orc_n_constraints(3)
    ff, fc = oracle_lib.check_evm_x(w, fixed, row_end=w["n_steps"])
    assert (ff == 0xFFFFFFFF).all(), oracle_lib.first_failure(ff, oracle_lib.constraint_classes(3))
    rng = np.random.default_rng(3)
    for trial in range(40):
        w2 = dict(w)
        which = trial % 4
        if which == 0:
            m = w["steps"].copy(); m[int(rng.integers(1, 13)), int(rng.integers(0, w["n_steps"])), 0] += np.uint64(1); w2["steps"] = m
        elif which == 1:
            m = w["rw"].copy(); m[int(rng.choice([3, 4, 5, 8, 10])), int(rng.integers(0, m.shape[1])), 0] ^= np.uint64(1); w2["rw"] = m
        elif which == 2:
            m = w["tx"].copy(); m[3, int(rng.integers(0, m.shape[1])), 0] += np.uint64(1); w2["tx"] = m
        else:
            m = w["block"].copy(); m[2, int(rng.integers(0, m.shape[1])), 0] += np.uint64(1); w2["block"] = m
        off, ofc = oracle_lib.check_evm_x(w2, fixed, row_end=w["n_steps"])
        for positional in (True, False):
            emu_lib.set_positional(positional)
            eff, efc = emu_lib.check_evm_x(w2, fixed, n=n, row_end=w["n_steps"])
            assert np.array_equal(eff, off) and np.array_equal(efc, ofc), (trial, positional, np.nonzero(eff != off))
    emu_lib.set_positional(True)
