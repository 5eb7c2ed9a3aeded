"""StepState.aux_data -> the step-aux side table (ZK_TABLE_STEP_AUX) that evm_circuit.main.verify_steps ships for
CREATE / CREATE2 (reference create.py:107 reads curr.aux_data as the init code's hash)."""
from zkevm_specs_b200 import native
from zkevm_specs_b200.evm_circuit.main import step_aux_rows
from zkevm_specs_b200.evm_circuit.step import StepState
from zkevm_specs_b200.evm_circuit.spec import ExecutionState
from zkevm_specs_b200.util.arithmetic import Word


def test_step_aux_rows_one_row_per_step_with_a_word():
    h = 0xC5D2460186F7233C927E7DB2DCC703C0E500B653CA82273B7BFAD8045D85A470
    steps = [
        StepState(execution_state=ExecutionState.CREATE2, rw_counter=1, aux_data=Word(h)),
        StepState(execution_state=ExecutionState.STOP, rw_counter=30),
        StepState(execution_state=ExecutionState.CREATE, rw_counter=40, aux_data=Word(7)),
    ]
    assert step_aux_rows(steps) == [[0, h & ((1 << 128) - 1), h >> 128], [2, 7, 0]]
    assert step_aux_rows(steps, row_base=100)[1][0] == 102
    assert step_aux_rows([steps[1]]) == []
    oog = StepState(execution_state=ExecutionState.ErrorOutOfGasSloadSstore, rw_counter=1, aux_data=(5 << 128) + 9)
    assert step_aux_rows([steps[1], oog]) == [[1, 9, 5]]


def test_table_ids_follow_the_header():
    import os
    import re

    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "zkcheck.h")).read()
    assert int(re.search(r"ZK_TABLE_STEP_AUX = (\d+)", hdr).group(1)) == native.TABLE_STEP_AUX == 12
    assert int(re.search(r"ZK_N_TABLES = (\d+)", hdr).group(1)) == 13


def test_step_aux_is_keyed_by_the_global_step_row():
    """A sharded check numbers its steps from row_base: the step-aux side table is keyed by that GLOBAL row (the row a
    failure is reported at), in the oracle and in the product's gate program (CPU emulation) alike"""
    import os
    import sys

    import numpy as np

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_lib
    import golden_util
    import oracle_lib
    from test_oracle_evm import fixed_table_matrix

    fixed = fixed_table_matrix()
    w = next(dict(w) for name, k, w, exp_row, exp_exc in golden_util.evm25_vectors() if exp_row == -1 and "sstore" in name)
    assert w["aux"].shape[1] == 1 and int(w["aux"][0, 0, 0]) == 0
    base = 1000
    for lib_ in (oracle_lib, emu_lib):
        ff, _ = lib_.check_evm_x(w, fixed, row_base=base)  # aux keyed by the local row 0: not found at global row 1000
        bad = np.nonzero(ff != 0xFFFFFFFF)[0]
        assert len(bad) == 1 and int(ff[bad[0]]) == base
    shifted = dict(w)
    shifted["aux"] = w["aux"].copy()
    shifted["aux"][0, 0, 0] = base
    for lib_ in (oracle_lib, emu_lib):
        ff, _ = lib_.check_evm_x(shifted, fixed, row_base=base)
        assert (ff == 0xFFFFFFFF).all()
