"""GPU parity: the CUDA bytecode-circuit checker (through the C-ABI) against
(1) the reference's own verdicts stored in tests/golden/bytecode.npz and
(2) the CPU oracle's full per-constraint result on the same inputs."""
import numpy as np
import pytest

import golden_util
import oracle_lib
from zkevm_specs_b200 import bytecode_circuit as bc
from zkevm_specs_b200 import native, packing

pytestmark = pytest.mark.gpu


def test_bytecode_golden_and_oracle_parity():
    ctx = native.default_context()
    n = 0
    for name, k, cols, push, kec, r, exp_row, exp_exc in golden_util.bytecode_vectors():
        ff, fc = bc.check_matrices(cols, push, kec, packing.cell_to_int(r), ctx)
        hit = native.first_failure(ff, native.CIRCUIT_BYTECODE)
        got = (-1, "") if hit is None else (hit[0], oracle_lib.EXC_OF_CLASS[hit[2]])
        assert got == (exp_row, exp_exc), f"{name}[{k}] cuda {got} reference {(exp_row, exp_exc)}"
        off, ofc = oracle_lib.check_bytecode(cols, push, kec, r)
        assert np.array_equal(ff, off), f"{name}[{k}] first_fail differs from oracle"
        assert np.array_equal(fc, ofc), f"{name}[{k}] fail_count differs from oracle"
        n += 1
    assert n > 400


def test_bytecode_host_api_roundtrip():
    """witness built by OUR host API passes; corrupting it raises AssertionError."""
    from zkevm_specs_b200.evm_circuit import Bytecode
    from zkevm_specs_b200.util import FQ

    r = FQ(0xABCDEF0123456789)
    codes = [bytes([0x60, 0x01, 0x60, 0x02, 0x01, 0x00]), b"", bytes(range(0x5F, 0x80))]
    unrolled = [bc.UnrolledBytecode(c, list(Bytecode(bytearray(c)).table_assignments())) for c in codes]
    rows = bc.assign_bytecode_circuit(8, unrolled, r)
    push, kec = bc.assign_push_table(), bc.assign_keccak_table(codes, r)
    bc.verify_bytecode_circuit(rows, push, kec, r)
    rows[3].value = FQ(rows[3].value.n + 1)
    with pytest.raises(AssertionError):
        bc.verify_bytecode_circuit(rows, push, kec, r)
    bc.check_bytecode_row(rows[10], rows[11], push, kec, r)


def test_bytecode_row_sharding_with_halo():
    """two shards with a one-row halo give the same verdict as the whole circuit."""
    ctx = native.default_context()
    for name, k, cols, push, kec, r, exp_row, exp_exc in golden_util.bytecode_vectors():
        if k != 5:
            continue
        n = cols.shape[1]
        whole, _ = bc.check_matrices(cols, push, kec, packing.cell_to_int(r), ctx)
        half = n // 2
        ctx.upload_columns(native.CIRCUIT_BYTECODE, np.ascontiguousarray(cols[:, : half + 1]))
        a, _ = ctx.check(native.CIRCUIT_BYTECODE, 0, half, 0, 0)
        shard = np.ascontiguousarray(np.concatenate([cols[:, half:], cols[:, :1]], axis=1))
        ctx.upload_columns(native.CIRCUIT_BYTECODE, shard)
        b, _ = ctx.check(native.CIRCUIT_BYTECODE, 0, n - half, half, 0)
        assert np.array_equal(np.minimum(a, b), whole)
