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


def test_keccak256_and_keccak_table_on_the_device():
    """zk_keccak256_batch / zk_assign_keccak_table (csrc/keccak.cuh, k_keccak256): digests equal the host sponge at every
    length around the 136-byte rate, and a bytecode circuit checked against the DEVICE-built keccak table gives the same
    arrays as against the host-built one (KeccakCircuit.add, typing.py:854-865)"""
    from zkevm_specs_b200.evm_circuit import Bytecode
    from zkevm_specs_b200.util import FQ
    from zkevm_specs_b200.util.hash import keccak256

    ctx = native.default_context()
    rng = np.random.default_rng(11)
    msgs = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in list(range(0, 140)) + [271, 272, 273, 1000, 24576, 50000]]
    got = ctx.keccak256_batch(msgs)
    for m, d in zip(msgs, got):
        assert d == keccak256(m), len(m)
    r = FQ(0x1234567ABCDEF)
    codes = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in (1, 33, 200, 3000)]
    rows = bc.assign_bytecode_circuit(12, [bc.UnrolledBytecode(c, list(Bytecode(bytearray(c)).table_assignments())) for c in codes], r)
    cols = bc.pack_rows(rows)
    push = bc.pack_push_table(bc.assign_push_table())
    kec = bc.pack_keccak_table(bc.assign_keccak_table(codes, r))
    ff_host, fc_host = bc.check_matrices(cols, push, kec, r, ctx)
    assert (ff_host == native.PASS).all()
    ctx.set_challenge(native.CHALLENGE_KECCAK, r.n)
    ctx.upload_table(native.TABLE_PUSH, push)
    ctx.upload_columns(native.CIRCUIT_BYTECODE, cols)
    ctx.assign_keccak_table(codes)
    ff_dev, fc_dev = ctx.check(native.CIRCUIT_BYTECODE, 0, cols.shape[1], 0, native.FLAG_WRAP)
    assert np.array_equal(ff_dev, ff_host) and np.array_equal(fc_dev, fc_host)
    ctx.assign_keccak_table(codes[:-1] + [codes[-1][:-1] + b"\\x00"])  # one wrong message: the last contract's lookup fails
    ff_bad, _ = ctx.check(native.CIRCUIT_BYTECODE, 0, cols.shape[1], 0, native.FLAG_WRAP)
    assert (ff_bad != native.PASS).sum() == 1
