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
