"""Pins the CPU oracle's state-circuit restatement against vectors produced by the reference's
own check_state_row loop (tests/golden/state.npz)."""
import golden_util
import oracle_lib


def test_oracle_state_matches_reference_golden():
    classes = oracle_lib.constraint_classes(1)
    n = n_fail = n_domain = 0
    kinds = set()
    for name, k, s, f, m, exp_row, exp_exc in golden_util.state_vectors():
        ff, fc = oracle_lib.check_state(s, f, m)
        row, exc = oracle_lib.first_failure(ff, classes)
        if exc == "NotImplementedError" and exp_exc != exc:
            assert row == exp_row, f"{name}[{k}]"  # declared limit (ST_WITNESS_DOMAIN), same row
            n_domain += 1
            continue
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
        kinds.add(exp_exc)
    assert n > 300 and n_fail > 200 and n_domain < 5
    assert {"AssertionError", "LookupUnsatFailure", "ValueError"} <= kinds
