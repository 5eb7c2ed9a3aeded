"""Pins the CPU oracle's exp-circuit restatement against vectors produced by the reference's own
verify_step (tests/golden/exp.npz)."""
import golden_util
import oracle_lib


def test_oracle_exp_matches_reference_golden():
    classes = oracle_lib.constraint_classes(4)
    n = n_fail = 0
    kinds = set()
    for name, k, r, exp_row, exp_exc in golden_util.exp_vectors():
        ff, fc = oracle_lib.check_exp(r)
        row, exc = oracle_lib.first_failure(ff, classes)
        if exc == "ValueError" and exp_exc == "OverflowError":
            exc = "OverflowError"
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert n > 300 and n_fail > 200
    assert {"AssertionError", "ConstraintUnsatFailure"} <= kinds
