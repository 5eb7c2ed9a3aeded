"""GPU parity of the CUDA copy-circuit checker against the reference's verdicts
(tests/golden/copy.npz), the CPU oracle's per-constraint result, and the host API."""
import numpy as np
import pytest

import golden_util
import oracle_lib
from zkevm_specs_b200 import native, packing

pytestmark = pytest.mark.gpu


def _device_check(ctx, w, r, row_begin=0, row_end=None, flags=native.FLAG_WRAP):
    ctx.set_challenge(native.CHALLENGE_KECCAK, packing.cell_to_int(r))
    ctx.upload_table(native.TABLE_RW, w["rw"], flags=w["rw_flags"])
    ctx.upload_table(native.TABLE_BYTECODE, w["bytecode"])
    ctx.upload_table(native.TABLE_TX, w["tx"], flags=w["tx_flags"])
    ctx.upload_columns(native.CIRCUIT_COPY, w["copy"], flags=w["copy_flags"])
    n = w["copy"].shape[1]
    return ctx.check(native.CIRCUIT_COPY, row_begin, n if row_end is None else row_end, 0, flags)


def test_copy_golden_and_oracle_parity():
    ctx = native.default_context()
    n = 0
    for name, k, w, r, exp_row, exp_exc in golden_util.copy_vectors():
        ff, fc = _device_check(ctx, w, r)
        off, ofc = oracle_lib.check_copy(w, r)
        assert np.array_equal(ff, off) and np.array_equal(fc, ofc), f"{name}[{k}] differs from oracle"
        hit = native.first_failure(ff, native.CIRCUIT_COPY)
        got = (-1, "") if hit is None else (hit[0], oracle_lib.EXC_OF_CLASS[hit[2]])
        assert got == (exp_row, exp_exc), f"{name}[{k}] cuda {got} reference {(exp_row, exp_exc)}"
        n += 1
    assert n > 200


def test_verify_copy_table_host_api():
    """SHA3-style Memory -> RlcAcc copy built with our CopyCircuit, like tests/evm/test_sha3.py:80-111"""
    from zkevm_specs_b200.copy_circuit import verify_copy_table
    from zkevm_specs_b200.evm_circuit import Block, CopyCircuit, CopyDataTypeTag, RWDictionary, Tables
    from zkevm_specs_b200.util import FQ

    r = FQ(0x1234567890ABCDEF)
    data = {100 + k: (7 * k + 3) % 256 for k in range(40)}
    rw = RWDictionary(5)
    for a, b in data.items():
        rw.memory_write(1, a, b)
    cc = CopyCircuit().copy(r, rw, 1, CopyDataTypeTag.Memory, 1, CopyDataTypeTag.RlcAcc, 100, 140, 0, 40, data)
    tables = Tables(block_table=set(Block().table_assignments()), tx_table=set(), withdrawal_table=set(),
                    bytecode_table=set(), rw_table=set(rw.rws), copy_circuit=cc.rows)
    verify_copy_table(cc, tables, r)
    assert len(tables.copy_table) == 1
    # corrupt the value of one memory READ row the copy circuit looks up (row 40 + 17 of the rw list)
    import dataclasses

    from zkevm_specs_b200.util import WordOrValue

    rws = list(rw.rws)
    rws[57] = dataclasses.replace(rws[57], value=WordOrValue(FQ((rws[57].value.lo.n + 1) % 256)))
    tables_bad = Tables(block_table=set(), tx_table=set(), withdrawal_table=set(), bytecode_table=set(),
                        rw_table=set(rws))
    with pytest.raises(AssertionError):
        verify_copy_table(cc, tables_bad, r)


def test_copy_synthetic_2e14_rows_and_corruptions_match_oracle():
    """cfg4 generator (SHA3-style + CALLDATACOPY-style events) at 2^14 rows: passes; seeded
    corruptions give identical per-constraint results on GPU and oracle"""
    from zkevm_specs_b200 import synth

    ctx = native.default_context()
    w = synth.copy_events(16, 512, seed=4)
    ff, fc = _device_check(ctx, w, w["r"])
    assert (ff == native.PASS).all() and fc.sum() == 0
    rng = np.random.default_rng(44)
    detected = 0
    for t in range(32):
        v = dict(w)
        kind = t % 4
        if kind == 0:
            v["copy"] = w["copy"].copy(); v["copy"][9, int(rng.integers(w["copy"].shape[1])), 0] ^= np.uint64(1)
        elif kind == 1:
            v["rw"] = w["rw"].copy(); v["rw"][8, int(rng.integers(w["rw"].shape[1])), 0] ^= np.uint64(2)
        elif kind == 2:
            v["tx"] = w["tx"].copy(); v["tx"][3, int(rng.integers(w["tx"].shape[1])), 0] ^= np.uint64(4)
        else:
            v["copy"] = w["copy"].copy(); v["copy"][13, int(rng.integers(w["copy"].shape[1])), 0] += np.uint64(1)
        ff, fc = _device_check(ctx, v, w["r"])
        off, ofc = oracle_lib.check_copy(v, w["r"])
        assert np.array_equal(ff, off) and np.array_equal(fc, ofc), f"corruption {t} kind {kind}"
        detected += bool((ff != native.PASS).any())
    assert detected >= 28
