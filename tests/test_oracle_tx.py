"""Oracle (oracle/tx.c) vs the reference's own verdicts for the Fr parts of the tx circuit
(tests/golden/tx.npz: tx_circuit.verify_circuit run unmodified, ECDSA chip stood in — see
tests/golden/gen_golden.py:tx_cases), and the CPU emulation of the device row program vs the oracle."""
import numpy as np

import emu_lib
import golden_util
import oracle_lib


def test_oracle_tx_matches_reference_golden():
    classes = oracle_lib.constraint_classes(5)
    n = n_fail = 0
    for k, rows, flags, kec, r, exp_row, exp_exc in golden_util.tx_vectors():
        ff, fc = oracle_lib.check_tx(rows, flags, kec, r)
        row, exc = oracle_lib.first_failure(ff, classes)
        assert (row, exc) == (exp_row, exp_exc), f"[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
    assert n > 350 and n_fail > 300


def test_emu_tx_equals_oracle_on_goldens():
    n = oracle_lib.lib().orc_n_constraints(5)
    for packed in (False, True):
        emu_lib.set_packed(packed)
        try:
            for k, rows, flags, kec, r, exp_row, exp_exc in golden_util.tx_vectors():
                ff, fc = emu_lib.check_tx(rows, flags, kec, r)
                off, ofc = oracle_lib.check_tx(rows, flags, kec, r)
                assert np.array_equal(ff[:n], off) and np.array_equal(fc[:n], ofc), (k, packed, ff[:n], off)
        finally:
            emu_lib.set_packed(False)


def test_oracle_sig_matches_reference_golden():
    """sig_circuit.Row.verify (sig_circuit.py:64-104), tests/golden/sig.npz"""
    classes = oracle_lib.constraint_classes(6)
    n = n_fail = 0
    for k, rows, flags, kec, r, exp_row, exp_exc in golden_util.sig_vectors():
        ff, fc = oracle_lib.check_sig(rows, flags, kec, r)
        row, exc = oracle_lib.first_failure(ff, classes)
        assert (row, exc) == (exp_row, exp_exc), f"[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
    assert n > 350 and n_fail > 300


def test_emu_sig_equals_oracle_on_goldens():
    n = oracle_lib.lib().orc_n_constraints(6)
    for packed in (False, True):
        emu_lib.set_packed(packed)
        try:
            for k, rows, flags, kec, r, exp_row, exp_exc in golden_util.sig_vectors():
                ff, fc = emu_lib.check_sig(rows, flags, kec, r)
                off, ofc = oracle_lib.check_sig(rows, flags, kec, r)
                assert np.array_equal(ff[:n], off) and np.array_equal(fc[:n], ofc), (k, packed, ff[:n], off)
        finally:
            emu_lib.set_packed(False)


class _LE:
    def __init__(self, v):
        self.le_bytes = int(v).to_bytes(32, "little")


class _Chip:
    """stands in for the ECDSA chips (tx_circuit.py:107-158, util/ec.py:59-117): eth_keys is third party"""

    def __init__(self, pk, msg_bytes, ok=True):
        self.pub_key_x_bytes, self.pub_key_y_bytes = bytes(reversed(pk[:32])), bytes(reversed(pk[32:]))
        self.msg_hash_bytes, self.ok = msg_bytes, ok
        self.sig_v, self.sig_r, self.sig_s = _LE(1), _LE(0x1234 << 130), _LE(0x5678)

    def verify(self, assert_msg=None):
        if assert_msg is None:  # sig circuit: returns the verdict
            return self.ok
        assert self.ok, f"{assert_msg}: ecdsa_verify failed"


def test_host_tx_and_sig_packers_produce_witnesses_the_oracle_accepts():
    """zkevm_specs_b200.tx_circuit.pack_witness / sig_circuit.pack_witness (host mirror, CPU only):
    objects built like the reference's tests pack into cell matrices that the oracle accepts, and a
    corrupted object is rejected at the right row with the right constraint"""
    from zkevm_specs_b200 import sig_circuit as sc
    from zkevm_specs_b200 import tx_circuit as tc
    from zkevm_specs_b200.util import FQ, Word, keccak256

    r = FQ(0x123456789ABCDEF123456789)
    rl = np.array([(r.n >> (64 * j)) & (2**64 - 1) for j in range(4)], dtype=np.uint64)
    rng = np.random.default_rng(11)
    kt, rows, chips, srows = tc.KeccakTable(), [], [], []
    for i in range(4):
        pk, msg = bytes(rng.integers(0, 256, 64, dtype=np.uint8)), bytes(rng.integers(0, 256, 32, dtype=np.uint8))
        kt.add(pk, r)
        h = keccak256(pk)
        addr, mw = FQ(int.from_bytes(h[-20:], "big")), Word(int.from_bytes(msg, "big"))
        chips.append(tc.SignVerifyChip(h, addr, mw, _Chip(pk, bytes(reversed(msg)))))
        for tag in range(1, 13):
            rows.append(tc.Row(FQ(i + 1), FQ(tag), FQ(0), addr if tag == 4 else mw if tag == 12 else FQ(0)))
        srows.append(sc.Row(h, addr, Word(msg), _Chip(pk, msg, ok=i != 2), is_valid=i != 2))
    cells, flags, kec = tc.pack_witness(tc.Witness(rows, kt, chips), 4)
    ff, fc = oracle_lib.check_tx(cells, flags, kec, rl)
    assert (ff == 0xFFFFFFFF).all()
    chips[1].msg_hash = Word(chips[1].msg_hash.int_value() ^ 1)
    cells, flags, kec = tc.pack_witness(tc.Witness(rows, kt, chips), 4)
    ff, fc = oracle_lib.check_tx(cells, flags, kec, rl)
    assert oracle_lib.first_failure(ff, oracle_lib.constraint_classes(5)) == (1, "AssertionError") and fc.sum() == 1
    cells, flags, kec = sc.pack_witness(sc.Witness(srows, kt))
    ff, fc = oracle_lib.check_sig(cells, flags, kec, rl)
    assert (ff == 0xFFFFFFFF).all()
    srows[2].is_valid = True  # the chip says the signature is invalid
    cells, flags, kec = sc.pack_witness(sc.Witness(srows, kt))
    ff, fc = oracle_lib.check_sig(cells, flags, kec, rl)
    assert oracle_lib.first_failure(ff, oracle_lib.constraint_classes(6)) == (2, "AssertionError")
