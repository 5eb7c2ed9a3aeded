"""Inputs of the device witness assignment (csrc/assign.cu) and the rows the host mirrors of the reference's builders
produce for them (zkevm_specs_b200.bytecode_circuit.assign_bytecode_circuit, state rows of synth.state_rows = op2row,
evm_circuit.typing.CopyCircuit.copy — each validated against the reference by tests/golden/gen_golden.py)."""
import numpy as np

from zkevm_specs_b200 import assign, packing, synth
from zkevm_specs_b200 import bytecode_circuit as bc
from zkevm_specs_b200.evm_circuit.spec import CopyDataTypeTag as T
from zkevm_specs_b200.evm_circuit.typing import Bytecode, CopyCircuit, RWDictionary
from zkevm_specs_b200.util import FQ, Word
from zkevm_specs_b200.util.hash import keccak256

R = FQ(0x2545F4914F6CDD1D5851F42D4C957F2D14057B7EF767814F1234567)


def bytecode_cases():
    """yield (name, k, codes, expected uint64[12][2^k][4])"""
    rng = np.random.default_rng(21)
    rnd = lambda n: bytes(rng.integers(0, 256, n, dtype=np.uint8))  # noqa: E731
    pushy = bytes([0x7F]) + rnd(32) + bytes([0x60, 0x01, 0x61, 0xAA, 0xBB, 0x00, 0x7F]) + rnd(10)  # ends inside PUSH32 data
    sets = {
        "mixed": [rnd(100), b"", pushy, bytes([0x00]), rnd(300)],
        "one_long": [rnd(1000)],
        "empties": [b"", b"", rnd(3)],
    }
    for name, codes in sets.items():
        total = sum(len(c) + 1 for c in codes)
        for k in sorted({max(1, int(np.ceil(np.log2(total))) + 1), max(1, int(np.floor(np.log2(total))) - 1), 6}):
            un = [bc.UnrolledBytecode(c, list(Bytecode(bytearray(c)).table_assignments())) for c in codes]
            yield name, k, codes, bc.pack_rows(bc.assign_bytecode_circuit(k, un, R))


def state_case(n_rows=4096):
    """(operation cells uint64[15][n][4], flags, expected rows uint64[57][n][4], mpt)"""
    w = synth.state_rows(n_rows, seed=3, n_start=64)
    return assign.state_ops_from_rows(w["rows"]), w["flags"], w["rows"], w["mpt"]


def copy_cases():
    """yield (name, events uint64[n][16], data bytes, is_code flags or None, expected rows uint64[20][m][4], flags)"""
    rng = np.random.default_rng(22)
    code = bytes([0x60, 0x05, 0x7F]) + bytes(rng.integers(0, 256, 32, dtype=np.uint8)) + bytes(rng.integers(0, 256, 45, dtype=np.uint8))
    code_flags = assign.is_code_bits(code)
    code_hash = Word(int.from_bytes(keccak256(code), "big"))
    mem = {a: int(v) for a, v in enumerate(rng.integers(0, 256, 4096, dtype=np.uint8))}
    cd = {a: int(v) for a, v in enumerate(rng.integers(0, 256, 200, dtype=np.uint8))}
    code_data = {a: (code[a], int(code_flags[a])) for a in range(len(code))}
    mem_code = {a: (mem[a], 0) for a in mem}
    specs = {
        # name: list of (src_id, src_tag, dst_id, dst_tag, src_addr, src_addr_end, dst_addr, length, src_data, log_id)
        "mem_to_mem": [(3, T.Memory, 4, T.Memory, 10, 4096, 70, 77, mem, 0)],
        "calldata_pad": [(2, T.TxCalldata, 5, T.Memory, 150, 200, 0, 100, cd, 0), (2, T.TxCalldata, 5, T.Memory, 300, 200, 0, 40, cd, 0)],
        "sha3": [(7, T.Memory, 7, T.RlcAcc, 33, 4096, 0, 100, mem, 0), (7, T.Memory, 7, T.RlcAcc, 0, 4096, 0, 1, mem, 0)],
        "codecopy_pad": [(code_hash, T.Bytecode, 9, T.Memory, 60, len(code), 5, 50, code_data, 0)],
        "log": [(6, T.Memory, 11, T.TxLog, 8, 4096, 0, 45, mem, 2)],
        "return_create": [(8, T.Memory, code_hash, T.Bytecode, 0, 4096, 0, 64, mem_code, 0)],
        "mixed_with_empty": [(3, T.Memory, 4, T.Memory, 0, 4096, 0, 0, mem, 0), (7, T.Memory, 7, T.RlcAcc, 5, 4096, 0, 65, mem, 0),
                             (2, T.TxCalldata, 5, T.Memory, 0, 200, 64, 33, cd, 0)],
    }
    for name, evs in specs.items():
        rw = RWDictionary(17)
        cc = CopyCircuit()
        events, data, flags, any_code = [], bytearray(), [], False
        for (sid, st, did, dt, sa, se, da, ln, src, log_id) in evs:
            events.append(assign.copy_event(sid, st, did, dt, sa, se, da, ln, rw.rw_counter, log_id))
            for i in range(ln):
                item = src[sa + i] if sa + i < se else 0
                if isinstance(item, tuple):
                    any_code = True
                    data.append(item[0]); flags.append(item[1])
                else:
                    data.append(item); flags.append(0)
            cc.copy(R, rw, sid, st, did, dt, sa, se, da, ln, src, log_id)
        rows = list(cc.table())
        exp = packing.pack(rows, packing.copy_circuit_row, 20)
        exp_flags = np.array([packing.word_flag(x.id) if hasattr(x.id, "is_word") else 1 for x in rows], dtype=np.uint8)
        yield name, np.array(events, dtype=np.uint64).reshape(-1, 16), bytes(data), (flags if any_code else None), exp, exp_flags
