"""GPU parity of the CUDA exp-circuit checker against the reference's verdicts
(tests/golden/exp.npz), the CPU oracle, and the host API (tests/evm/test_exp.py:45-80 style)."""
import dataclasses

import numpy as np
import pytest

import golden_util
import oracle_lib
from zkevm_specs_b200 import exp_circuit as xc
from zkevm_specs_b200 import native

pytestmark = pytest.mark.gpu


def test_exp_golden_and_oracle_parity():
    ctx = native.default_context()
    n = 0
    for name, k, r, exp_row, exp_exc in golden_util.exp_vectors():
        ff, fc = xc.check_rows(ctx, r)
        off, ofc = oracle_lib.check_exp(r)
        assert np.array_equal(ff, off) and np.array_equal(fc, ofc), f"{name}[{k}] differs from oracle"
        hit = native.first_failure(ff, native.CIRCUIT_EXP)
        got = (-1, "") if hit is None else (hit[0], oracle_lib.EXC_OF_CLASS[hit[2]])
        if got[1] == "ValueError" and exp_exc == "OverflowError":
            got = (got[0], exp_exc)
        assert got == (exp_row, exp_exc), f"{name}[{k}] cuda {got} reference {(exp_row, exp_exc)}"
        n += 1
    assert n > 300


def test_exp_host_api_and_scale():
    from zkevm_specs_b200.util import Word

    c = xc.ExpCircuit(max_exp_steps=40)
    rng = np.random.default_rng(8)
    for ident in range(1, 60):
        base = int.from_bytes(rng.bytes(32), "little")
        expo = int(rng.integers(2, 1 << 30))
        c.add_event(base, expo, ident * 7)
    c.max_exp_steps = (len(c.rows) + 6) // 7 + 1
    c.fill_dummy_events()
    xc.verify_exp_circuit(c)
    bad = xc.ExpCircuit()
    bad.rows = list(c.rows)
    bad.rows[5] = dataclasses.replace(bad.rows[5], d=Word((bad.rows[5].d.int_value() + 1) % (1 << 256)))
    with pytest.raises(Exception) as ei:
        xc.verify_exp_circuit(bad)
    assert type(ei.value).__name__ in ("AssertionError", "ConstraintUnsatFailure")
