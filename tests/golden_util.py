"""Readers for tests/golden/*.npz (vectors produced by the reference itself, see gen_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bytecode_vectors():
    """yield (case, vector idx, cols, push, keccak, r, expected_row, expected_exc)"""
    z = np.load(os.path.join(GOLDEN, "bytecode.npz"))
    r = z["r"]
    for name in z["names"]:
        name = str(name)
        base, push, kec = z[f"{name}/cols"], z[f"{name}/push"], z[f"{name}/keccak"]
        for k in range(len(z[f"{name}/mut_row"])):
            i, c = int(z[f"{name}/mut_row"][k]), int(z[f"{name}/mut_col"][k])
            cols, kk = base, kec
            if i >= 1000:
                kk = kec.copy()
                kk[c - 100, i - 1000, :] = z[f"{name}/mut_val"][k]
            elif i >= 0:
                cols = base.copy()
                cols[c, i, :] = z[f"{name}/mut_val"][k]
            yield name, k, cols, push, kk, r, int(z[f"{name}/exp_row"][k]), str(z[f"{name}/exp_exc"][k])
