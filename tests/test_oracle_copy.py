"""Pins the CPU oracle's copy-circuit restatement against vectors produced by the reference's
own verify_copy_table (tests/golden/copy.npz): same first failing row, same exception class."""
import golden_util
import oracle_lib


def test_oracle_copy_matches_reference_golden():
    classes = oracle_lib.constraint_classes(2)
    n = n_fail = 0
    kinds = set()
    for name, k, w, r, exp_row, exp_exc in golden_util.copy_vectors():
        ff, fc = oracle_lib.check_copy(w, r)
        row, exc = oracle_lib.first_failure(ff, classes)
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert n > 200 and n_fail > 120
    assert {"AssertionError", "LookupUnsatFailure", "LookupAmbiguousFailure"} <= kinds
