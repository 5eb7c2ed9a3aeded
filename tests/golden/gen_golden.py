"""Generate tests/golden/*.npz by running the REFERENCE itself (authoring container only).

    PYTHONPATH=oracle/pyshim:/root/reference/src python tests/golden/gen_golden.py [bytecode|state|copy|evm|all]

The reference (/root/reference, pure Python) is imported through oracle/pyshim (stand-ins
for its absent third-party deps).  For every case we store the witness matrix in the
column-major cell layout of include/zkcheck.h plus what the reference's own driver loop did
on it: pass, or (first failing row, exception class).  Negative cases are seeded single-cell
corruptions of a positive witness, the reference's own negative-test pattern
(tests/test_bytecode_circuit.py:113-313, tests/test_state_circuit.py:113-372).

/root/reference does not exist on the GPU box, so tests only ever read the .npz files.
"""
import dataclasses
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def limbs(v: int):
    assert 0 <= v < (1 << 256)
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def n_of(x) -> int:
    if hasattr(x, "expr"):
        return x.expr().n
    if hasattr(x, "n"):
        return x.n
    return int(x)


def to_matrix(rows_of_ints):
    """rows_of_ints: list of rows, each a list of python ints -> uint64[n_cols][n_rows][4]"""
    n_rows = len(rows_of_ints)
    n_cols = len(rows_of_ints[0]) if n_rows else 0
    out = np.zeros((n_cols, n_rows, 4), dtype=np.uint64)
    for i, row in enumerate(rows_of_ints):
        for c, v in enumerate(row):
            out[c, i, :] = limbs(v)
    return out


INTERESTING = [0, 1, 2, 3, 255, 256, 0x7F, 0x60, 1 << 64, (1 << 128) - 1, 1 << 128, P - 1]


def corrupt_value(rng: random.Random, old: int) -> int:
    while True:
        k = rng.randrange(6)
        if k == 0:
            v = (old + 1) % P
        elif k == 1:
            v = (old - 1) % P
        elif k == 2:
            v = rng.choice(INTERESTING)
        elif k == 3:
            v = rng.randrange(P)
        elif k == 4:
            v = old ^ (1 << rng.randrange(0, 16))
        else:
            v = rng.randrange(0, 1 << rng.choice([8, 16, 32, 64, 128]))
        if v != old and v < P:
            return v


# --------------------------------------------------------------------------- bytecode
def bytecode_cases():
    from zkevm_specs import bytecode_circuit as bc
    from zkevm_specs.evm_circuit import Bytecode, Opcode, is_push_with_data
    from zkevm_specs.util import FQ, Word

    r = FQ(0x1234567890ABCDEF1122334455667788990011223344556677889900AABBCCDD % P)

    def unroll(code: bytes):
        return bc.UnrolledBytecode(code, list(Bytecode(bytearray(code)).table_assignments()))

    def row_ints(row):
        return [n_of(row.q_first), n_of(row.q_last), n_of(row.hash.lo), n_of(row.hash.hi),
                n_of(row.tag), n_of(row.index), n_of(row.value), n_of(row.is_code),
                n_of(row.push_data_left), n_of(row.value_rlc), n_of(row.length),
                n_of(row.push_data_size)]

    FIELDS = ["q_first", "q_last", "hash.lo", "hash.hi", "tag", "index", "value", "is_code",
              "push_data_left", "value_rlc", "length", "push_data_size"]

    def set_cell(row, col, v):
        f = FIELDS[col]
        if f == "hash.lo":
            return dataclasses.replace(row, hash=Word((FQ(v), row.hash.hi), check=False))
        if f == "hash.hi":
            return dataclasses.replace(row, hash=Word((row.hash.lo, FQ(v)), check=False))
        return dataclasses.replace(row, **{f: FQ(v)})

    def run(rows, push_table, keccak_table):
        """the reference driver loop, tests/test_bytecode_circuit.py:26-47"""
        for idx, row in enumerate(rows):
            try:
                bc.check_bytecode_row(row, rows[(idx + 1) % len(rows)], push_table, keccak_table, r)
            except Exception as e:  # noqa: BLE001
                return idx, type(e).__name__
        return -1, ""

    all_bytes = []
    for b in range(256):
        if not is_push_with_data(b):
            all_bytes.append(b)
    for n in range(1, 33):
        all_bytes.append(int(Opcode.PUSH0) + n)
        all_bytes.extend([int(Opcode.PUSH32)] * n)

    rng = random.Random(1)
    cfg1_code = bytes(np.random.default_rng(1).integers(0, 256, 32, dtype=np.uint8))
    positives = {
        "unrolling_all_bytes": (10, [bytes(all_bytes)]),
        "empty": (6, [b""]),
        "full": (6, [bytes([7] * (2**6 - 2)), b""]),
        "incomplete": (6, [bytes([7] * (2**6 + 1))]),
        "multiple": (6, [b"", bytes([0x7F]), bytes([0x7F, 1]), bytes([1, 0x7F]), bytes([1, 0x7F, 1])]),
        "small": (5, [bytes([8, 2, 3, 8, 9, 7, 128])]),
        "is_code": (5, [bytes([0x01, 0x60, 0x60, 0x03, 0x66, 0x01, 0x65])]),
        "cfg1_32byte_contract": (10, [cfg1_code]),
        "truncated_push": (5, [bytes([0x61, 0xAA]), bytes([0x7F] * 5)]),
    }
    cases = []
    for name, (k, codes) in positives.items():
        unrolled = [unroll(c) for c in codes]
        rows = bc.assign_bytecode_circuit(k, unrolled, r)
        push_table = bc.assign_push_table()
        keccak_table = bc.assign_keccak_table(codes, r)
        push_m = to_matrix([[n_of(a), n_of(b)] for a, b in push_table])
        kec_m = to_matrix([[n_of(x.state_tag), n_of(x.input_rlc), n_of(x.input_len),
                            n_of(x.output.lo), n_of(x.output.hi)] for x in keccak_table])
        base = to_matrix([row_ints(x) for x in rows])
        fail_row, exc = run(rows, push_table, keccak_table)
        muts = [(-1, -1, 0, fail_row, exc)]
        n_mut = 60 if len(rows) <= 64 else 40
        live = max(2, sum(len(c) + 1 for c in codes) + 2)
        for _ in range(n_mut):
            i = rng.randrange(min(len(rows), live)) if rng.random() < 0.8 else rng.randrange(len(rows))
            c = rng.randrange(12)
            old = row_ints(rows[i])[c]
            v = corrupt_value(rng, old)
            bad = list(rows)
            bad[i] = set_cell(rows[i], c, v)
            fr_, ex_ = run(bad, push_table, keccak_table)
            muts.append((i, c, v, fr_, ex_))
        # corrupting the tables instead of the witness
        for _ in range(6):
            kt = list(keccak_table)
            if kt:
                j = rng.randrange(len(kt))
                kt[j] = dataclasses.replace(kt[j], input_len=FQ(n_of(kt[j].input_len) + 1))
                fr_, ex_ = run(rows, push_table, set(kt))
                # encode a table mutation as col = 100 + table col, row = table row (keccak col 2)
                # the stored keccak table is the ORIGINAL; the loader re-applies the mutation
                muts.append((1000 + j, 102, n_of(kt[j].input_len), fr_, ex_))
                break
        cases.append((name, base, push_m, kec_m, muts))
    out = {"r": np.array(limbs(r.n), dtype=np.uint64), "names": np.array([c[0] for c in cases])}
    for name, base, push_m, kec_m, muts in cases:
        out[f"{name}/cols"] = base
        out[f"{name}/push"] = push_m
        out[f"{name}/keccak"] = kec_m
        out[f"{name}/mut_row"] = np.array([m[0] for m in muts], dtype=np.int64)
        out[f"{name}/mut_col"] = np.array([m[1] for m in muts], dtype=np.int64)
        out[f"{name}/mut_val"] = np.array([limbs(m[2]) for m in muts], dtype=np.uint64)
        out[f"{name}/exp_row"] = np.array([m[3] for m in muts], dtype=np.int64)
        out[f"{name}/exp_exc"] = np.array([m[4] for m in muts])
    np.savez_compressed(os.path.join(HERE, "bytecode.npz"), **out)
    n_fail = sum(1 for c in cases for m in c[4] if m[3] >= 0)
    n_all = sum(len(c[4]) for c in cases)
    print(f"bytecode: {len(cases)} witnesses, {n_all} vectors ({n_fail} failing)")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    todo = {"bytecode": bytecode_cases}
    g = globals()
    for nm in ["state", "copy", "evm", "fr"]:
        if nm + "_cases" in g:
            todo[nm] = g[nm + "_cases"]
    for nm, fn in todo.items():
        if which in ("all", nm):
            fn()
