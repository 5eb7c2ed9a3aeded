"""Pins the CPU oracle's bytecode-circuit restatement against vectors produced by the
reference's own check_bytecode_row loop (tests/golden/bytecode.npz)."""
import numpy as np

import golden_util
import oracle_lib

BC_CLASSES = [0] * 22


def test_oracle_bytecode_matches_reference_golden():
    n = n_fail = 0
    for name, k, cols, push, kec, r, exp_row, exp_exc in golden_util.bytecode_vectors():
        ff, fc = oracle_lib.check_bytecode(cols, push, kec, r)
        row, exc = oracle_lib.first_failure(ff, BC_CLASSES)
        assert (row, exc) == (exp_row, exp_exc), f"{name}[{k}]: oracle {(row, exc)} reference {(exp_row, exp_exc)}"
        n += 1
        n_fail += exp_row >= 0
    assert n > 400 and n_fail > 100


def test_oracle_bytecode_row_range():
    """row_begin/row_end restrict the rows checked (shard semantics)."""
    for name, k, cols, push, kec, r, exp_row, exp_exc in golden_util.bytecode_vectors():
        if exp_row < 2:
            continue
        ff, _ = oracle_lib.check_bytecode(cols, push, kec, r, 0, exp_row)
        assert (ff == 0xFFFFFFFF).all()
        ff, _ = oracle_lib.check_bytecode(cols, push, kec, r, exp_row, cols.shape[1])
        assert ff.min() == exp_row
        break
