"""Enumerations and static EVM facts, built from spec_data.json (extracted from the reference
by tools/gen_spec_data.py; see that script for provenance).  Mirrors the names of
evm_circuit/{opcode,execution_state,table}.py so host code reads like the reference's."""
from __future__ import annotations

import enum
import json
import os
from typing import Dict, List, Tuple

from ..util.arithmetic import FQ

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "spec_data.json")) as _f:
    SPEC = json.load(_f)


class _FieldEnum(enum.IntEnum):
    def expr(self) -> FQ:
        return FQ(int(self))


def _mk(name: str, base=_FieldEnum):
    return base(name, sorted(SPEC["enums"][name].items(), key=lambda kv: kv[1]))


class _OpcodeBase(_FieldEnum):
    def hex(self) -> str:
        return f"{int(self):02x}"

    def bytes(self) -> bytes:
        return bytes([int(self)])

    def is_push(self) -> bool:
        return 0x5F <= int(self) <= 0x7F

    def is_push_with_data(self) -> bool:
        return 0x60 <= int(self) <= 0x7F

    def is_dup(self) -> bool:
        return 0x80 <= int(self) <= 0x8F

    def is_swap(self) -> bool:
        return 0x90 <= int(self) <= 0x9F

    def min_stack_pointer(self) -> int:
        return SPEC["opcode_info"][self.name][0]

    def max_stack_pointer(self) -> int:
        return SPEC["opcode_info"][self.name][1]

    def constant_gas_cost(self) -> int:
        return SPEC["opcode_info"][self.name][2]

    def has_dynamic_gas(self) -> bool:
        return SPEC["opcode_info"][self.name][3]


class _ExecutionStateBase(_FieldEnum):
    def responsible_opcode(self) -> List[Tuple[int, int]]:
        return [tuple(p) for p in SPEC["execution_state"][self.name]["responsible"]]

    def halts(self) -> bool:
        return SPEC["execution_state"][self.name]["halts"]

    def halts_in_success(self) -> bool:
        return SPEC["execution_state"][self.name]["halts_in_success"]

    def halts_in_exception(self) -> bool:
        return SPEC["execution_state"][self.name]["halts_in_exception"]


Opcode = _mk("Opcode", _OpcodeBase)
ExecutionState = _mk("ExecutionState", _ExecutionStateBase)
FixedTableTag = _mk("FixedTableTag")
BlockContextFieldTag = _mk("BlockContextFieldTag")
TxContextFieldTag = _mk("TxContextFieldTag")
BytecodeFieldTag = _mk("BytecodeFieldTag")
RW = _mk("RW")
Target = _mk("Target")
CallContextFieldTag = _mk("CallContextFieldTag")
AccountFieldTag = _mk("AccountFieldTag")
TxLogFieldTag = _mk("TxLogFieldTag")
TxReceiptFieldTag = _mk("TxReceiptFieldTag")
CopyDataTypeTag = _mk("CopyDataTypeTag")
MPTProofType = _mk("MPTProofType")
StateTag = _mk("StateTag")
PARAMS: Dict[str, int] = SPEC["params"]


def is_push_with_data(op: int) -> bool:
    return 0x60 <= int(op) <= 0x7F


def get_push_size(op: int) -> int:
    """bytes pushed by PUSH1..PUSH32, else 0 (evm_circuit/opcode.py:427-433)."""
    return int(op) - 0x5F if is_push_with_data(op) else 0


def valid_opcodes() -> List[int]:
    return [int(o) for o in Opcode]
