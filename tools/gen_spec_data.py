"""Extract the reference's ENUMERATIONS and static EVM facts into spec_data.json.

    PYTHONPATH=oracle/pyshim:/root/reference/src python tools/gen_spec_data.py

Only numbers travel: enum member values (Opcode, ExecutionState, the table tags), the
per-opcode info map (stack bounds, constant gas — evm_circuit/opcode.py:212-358), each
execution state's responsible opcodes and halting flags (execution_state.py:143-420), the
precompile info pairs and the EVM parameters of util/param.py.  The product builds its own
IntEnums and fixed table from this file; no reference code is copied.  Runs in the authoring
container only (the reference is absent on the GPU box); the JSON is committed.
"""
import enum
import json
import os

import zkevm_specs.evm_circuit as ec
import zkevm_specs.evm_circuit.table as tb
import zkevm_specs.util.param as param
from zkevm_specs.evm_circuit import opcode as op
from zkevm_specs.evm_circuit.execution_state import ExecutionState
from zkevm_specs.evm_circuit.precompile import precompile_info_pairs, Precompile

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zkevm-specs_b200",
                   "evm_circuit", "spec_data.json")


def enum_dict(e):
    return {m.name: int(m.value) for m in e}


data = {"enums": {}, "opcode_info": {}, "execution_state": {}, "params": {}}
for mod in (tb, ec):
    for name in dir(mod):
        obj = getattr(mod, name)
        if isinstance(obj, type) and issubclass(obj, enum.IntEnum) and obj is not enum.IntEnum:
            data["enums"][name] = enum_dict(obj)
import zkevm_specs.state_circuit as sc  # noqa: E402

data["enums"]["StateTag"] = enum_dict(sc.Tag)
data["enums"]["Precompile"] = enum_dict(Precompile)

for o in op.Opcode:
    info = op.OPCODE_INFO_MAP[o]
    data["opcode_info"][o.name] = [info.min_stack_pointer, info.max_stack_pointer,
                                   info.constant_gas_cost, bool(info.has_dynamic_gas)]

for s in ExecutionState:
    resp = []
    for pair in s.responsible_opcode():
        resp.append([int(pair[0]), int(pair[1])] if isinstance(pair, tuple) else [int(pair), 0])
    data["execution_state"][s.name] = {
        "value": int(s),
        "responsible": resp,
        "halts": bool(s.halts()),
        "halts_in_success": bool(s.halts_in_success()),
        "halts_in_exception": bool(s.halts_in_exception()),
        "implemented": s in ec.execution.EXECUTION_STATE_IMPL,
    }
data["precompile_info_pairs"] = [[int(a), int(b), int(c)] for a, b, c in precompile_info_pairs()]
for name in dir(param):
    v = getattr(param, name)
    if name.isupper() and isinstance(v, int):
        data["params"][name] = int(v)
data["fixed_table_rows"] = len(tb.Tables.fixed_table)
with open(OUT, "w") as f:
    json.dump(data, f, indent=0, sort_keys=True)
print("wrote", os.path.normpath(OUT), {k: len(v) if hasattr(v, "__len__") else v for k, v in data.items()})
print(sorted(data["enums"].keys()))
