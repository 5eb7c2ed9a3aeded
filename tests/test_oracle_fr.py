"""Pins the oracle's Fr arithmetic (oracle/fr.h) against Python big-int arithmetic mod the
BN254 scalar modulus — the semantics of py_ecc.bn128.FQ that the reference's FQ subclasses
(src/zkevm_specs/util/arithmetic.py:41-63)."""
import ctypes
import random

import numpy as np

import oracle_lib

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _vals():
    rng = random.Random(7)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 128) - 1, 1 << 128, (1 << 64) - 1, 1 << 64, 1 << 253, 255, 256]
    return edge + [rng.randrange(P) for _ in range(64)]


def _call(fn, *args):
    out = np.zeros(4, dtype=np.uint64)
    fn(*[oracle_lib.p64(oracle_lib.limbs(a)) for a in args], oracle_lib.p64(out))
    return oracle_lib.from_limbs(out)


def test_fr_add_sub_mul():
    lib = oracle_lib.lib()
    vals = _vals()
    for a in vals:
        for b in vals[:24]:
            assert _call(lib.orc_fr_mul, a, b) == a * b % P
            assert _call(lib.orc_fr_add, a, b) == (a + b) % P
            assert _call(lib.orc_fr_sub, a, b) == (a - b) % P


def test_fr_inv_zero_is_zero():
    lib = oracle_lib.lib()
    assert _call(lib.orc_fr_inv, 0) == 0
    for a in _vals()[1:20]:
        assert _call(lib.orc_fr_inv, a) * a % P == 1
    # constants the MUL/DIV/MOD gadget multiplies by (mul_div_mod.py:14-16)
    assert _call(lib.orc_fr_inv, 8) == pow(8, -1, P)
    assert _call(lib.orc_fr_inv, 4) == pow(4, -1, P)
    assert _call(lib.orc_fr_inv, 1 << 128) == pow(1 << 128, -1, P)
