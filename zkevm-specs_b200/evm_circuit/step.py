"""StepState — the 13 cells the EVM circuit carries from step to step (reference:
evm_circuit/step.py:6-75).  Cell order on the device = attribute order below, code_hash
expanded to (lo, hi)."""
from __future__ import annotations

from typing import Any, Optional

from ..util.arithmetic import FQ, Word
from .spec import ExecutionState


class StepState:
    def __init__(self, execution_state: ExecutionState, rw_counter: int, call_id: int = 0,
                 is_root: bool = False, is_create: bool = False, code_hash: Word = Word(0),
                 program_counter: int = 0, stack_pointer: int = 1024, gas_left: int = 0,
                 memory_word_size: int = 0, reversible_write_counter: int = 0, log_id: int = 0,
                 aux_data: Optional[Any] = None) -> None:
        self.execution_state = execution_state
        self.rw_counter = FQ(rw_counter)
        self.call_id = FQ(call_id)
        self.is_root = is_root
        self.is_create = is_create
        self.code_hash = code_hash
        self.program_counter = FQ(program_counter)
        self.stack_pointer = FQ(stack_pointer)
        self.gas_left = FQ(gas_left)
        self.memory_word_size = FQ(memory_word_size)
        self.reversible_write_counter = FQ(reversible_write_counter)
        self.log_id = FQ(log_id)
        self.aux_data = aux_data
