"""Host mirror of Tables._convert_exp_circuit_to_table (reference table.py:654-671): the exp-table rows that
evm_circuit.main.upload_tables ships as ZK_TABLE_EXP equal the rows the reference derived for the same exponentiation
(tests/golden/evm19.npz holds the reference's exp table of every EXP scenario)."""
import os

import numpy as np

from zkevm_specs_b200.evm_circuit.main import exp_table_rows
from zkevm_specs_b200.exp_circuit import ExpCircuit

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "evm19.npz")


def _int(limbs):
    return sum(int(x) << (64 * k) for k, x in enumerate(limbs))


def test_exp_table_rows_equal_reference_rows():
    z = np.load(GOLDEN)
    checked = 0
    for name in z["names"]:
        name = str(name)
        if f"{name}/exp" not in z.files:
            continue  # exponent 0 / 1: no exp-circuit rows
        rw, ex = z[f"{name}/rw"], z[f"{name}/exp"]
        base = _int(rw[8, 0]) + (_int(rw[9, 0]) << 128)       # first stack read
        exponent = _int(rw[8, 1]) + (_int(rw[9, 1]) << 128)   # second stack read
        identifier = _int(rw[0, 2]) + 1                       # rw_counter after the three stack rows
        want = {tuple(_int(ex[c, r]) for c in range(11)) for r in range(ex.shape[1])}
        got = {tuple(r) for r in exp_table_rows(ExpCircuit().add_event(base, exponent, identifier).rows)}
        assert got and got <= want, name   # the golden table may also hold a second, unrelated event
        assert {r for r in want if r[1] == identifier} == got, name
        checked += 1
    assert checked >= 10
