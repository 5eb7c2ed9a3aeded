"""Oracle (oracle/tx.c) vs the reference's own verdicts for the Fr parts of the tx circuit
(tests/golden/tx.npz: tx_circuit.verify_circuit run unmodified, ECDSA chip stood in — see
tests/golden/gen_golden.py:tx_cases), and the CPU emulation of the device row program vs the oracle."""
import numpy as np

import emu_lib
import golden_util
import oracle_lib


def test_oracle_tx_matches_reference_golden():
    classes = oracle_lib.constraint_classes(5)
    n = n_fail = 0
    for k, rows, flags, kec, r, exp_row, exp_exc in golden_util.tx_vectors():
        ff, fc = oracle_lib.check_tx(rows, flags, kec, r)
        row, exc = oracle_lib.first_failure(ff, classes)
        assert (row, exc) == (exp_row, exp_exc), f"[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
    assert n > 350 and n_fail > 300


def test_emu_tx_equals_oracle_on_goldens():
    n = oracle_lib.lib().orc_n_constraints(5)
    for packed in (False, True):
        emu_lib.set_packed(packed)
        try:
            for k, rows, flags, kec, r, exp_row, exp_exc in golden_util.tx_vectors():
                ff, fc = emu_lib.check_tx(rows, flags, kec, r)
                off, ofc = oracle_lib.check_tx(rows, flags, kec, r)
                assert np.array_equal(ff[:n], off) and np.array_equal(fc[:n], ofc), (k, packed, ff[:n], off)
        finally:
            emu_lib.set_packed(False)


def test_oracle_sig_matches_reference_golden():
    """sig_circuit.Row.verify (sig_circuit.py:64-104), tests/golden/sig.npz"""
    classes = oracle_lib.constraint_classes(6)
    n = n_fail = 0
    for k, rows, flags, kec, r, exp_row, exp_exc in golden_util.sig_vectors():
        ff, fc = oracle_lib.check_sig(rows, flags, kec, r)
        row, exc = oracle_lib.first_failure(ff, classes)
        assert (row, exc) == (exp_row, exp_exc), f"[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
    assert n > 350 and n_fail > 300


def test_emu_sig_equals_oracle_on_goldens():
    n = oracle_lib.lib().orc_n_constraints(6)
    for packed in (False, True):
        emu_lib.set_packed(packed)
        try:
            for k, rows, flags, kec, r, exp_row, exp_exc in golden_util.sig_vectors():
                ff, fc = emu_lib.check_sig(rows, flags, kec, r)
                off, ofc = oracle_lib.check_sig(rows, flags, kec, r)
                assert np.array_equal(ff[:n], off) and np.array_equal(fc[:n], ofc), (k, packed, ff[:n], off)
        finally:
            emu_lib.set_packed(False)
