"""Host-side column packer (zkevm_specs_b200/packing.py): the packed format of include/zkcheck.h
("packed columns") round-trips every matrix exactly, picks the documented widths, rejects values
that do not fit, and bytecode_src_from_table inverts Bytecode.table_assignments.  CPU only."""
import numpy as np
import pytest

from zkevm_specs_b200 import packing, synth


def _matrix(rng, n_cols, n_rows, widths):
    m = np.zeros((n_cols, n_rows, 4), dtype=np.uint64)
    for c, w in enumerate(widths):
        if w == 0:
            m[c, :, :] = rng.integers(0, 1 << 62, 4, dtype=np.uint64)
        else:
            limbs = (w + 7) // 8
            m[c, :, :limbs] = rng.integers(0, 1 << 62, (n_rows, limbs), dtype=np.uint64)
            if w < 8:
                m[c, :, 0] &= np.uint64((1 << (8 * w)) - 1)
            m[c, 0, (w - 1) // 8] |= np.uint64(1) << np.uint64(8 * ((w - 1) % 8))  # make the width tight
    return m


def test_pack_matrix_roundtrip_and_width_selection():
    rng = np.random.default_rng(1)
    widths = [0, 1, 2, 4, 8, 16, 32, 1, 0, 4]
    m = _matrix(rng, len(widths), 1000, widths)
    pm = packing.pack_matrix(m)
    assert [int(w) for w in pm.widths] == widths
    assert all(int(o) % 32 == 0 for o in pm.offsets)
    assert np.array_equal(pm.unpack(), m)
    assert pm.nbytes < m.nbytes // 4


def test_pack_matrix_min_widths_never_narrower_and_never_constant():
    rng = np.random.default_rng(2)
    m = _matrix(rng, 4, 64, [0, 1, 16, 2])
    pm = packing.pack_matrix(m, min_widths=[1, 4, 8, 2])
    assert [int(w) for w in pm.widths] == [32, 4, 16, 2]  # a constant column is stored at its value's width (here 256 bits); wide data wins
    assert np.array_equal(pm.unpack(), m)


def test_pack_matrix_explicit_widths_must_fit():
    rng = np.random.default_rng(3)
    m = _matrix(rng, 3, 32, [2, 16, 1])
    assert np.array_equal(packing.pack_matrix(m, widths=[4, 32, 1]).unpack(), m)
    with pytest.raises(AssertionError):
        packing.pack_matrix(m, widths=[1, 16, 1])
    with pytest.raises(AssertionError):
        packing.pack_matrix(m, widths=[2, 16, 0])  # not a constant column


def test_pack_matrix_empty_and_single_row():
    e = np.zeros((5, 0, 4), dtype=np.uint64)
    assert packing.pack_matrix(e).unpack().shape == (5, 0, 4)
    one = np.arange(20, dtype=np.uint64).reshape(5, 1, 4)
    assert np.array_equal(packing.pack_matrix(one).unpack(), one)


def test_type_widths_hold_the_synthetic_witness():
    w = synth.evm_trace(64, seed=9)
    for key, typ in (("steps", "evm_steps"), ("rw", "rw_table"), ("bytecode", "bytecode_table")):
        pm = packing.pack_matrix(w[key], min_widths=packing.TYPE_WIDTHS[typ])
        assert [int(x) for x in pm.widths] == packing.TYPE_WIDTHS[typ]
        assert np.array_equal(pm.unpack(), w[key])


def test_bytecode_src_from_table_inverts_table_assignments():
    from zkevm_specs_b200.evm_circuit import Bytecode

    a = Bytecode().push(0x1234, n_bytes=2).push(7, n_bytes=32).add().stop()
    b = Bytecode().push(1, n_bytes=1).pop().stop()
    rows = list(a.table_assignments()) + list(b.table_assignments())
    table = packing.pack(rows, packing.bytecode_table_row, 6)
    src = packing.bytecode_src_from_table(table)
    assert src is not None and len(src["hashes"]) == 2
    assert bytes(src["code"]) == bytes(a.code) + bytes(b.code)
    bits = np.unpackbits(src["is_code_bits"], bitorder="little")[: len(src["code"])]
    assert list(bits) == [int(x) for x in list(a.is_code) + list(b.is_code)]
    assert list(src["code_offsets"]) == [0, len(a.code), len(a.code) + len(b.code)]
    # an irregular table (a Byte row out of order) is not invertible
    bad = table.copy()
    bad[3, 2, 0] += np.uint64(5)
    assert packing.bytecode_src_from_table(bad) is None
