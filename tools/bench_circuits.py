"""Supplementary measurements (not the bench.py contract): device time, rows/s and achieved GB/s
(algorithmic bytes) of the state-circuit (BASELINE cfg3, 2^18 rows) and copy-circuit
(cfg4, 2^20 rows) checkers on one B200, each pass including the lookup-index work.  Prints one
JSON line per circuit; results are committed under profiles/."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zkevm_specs_b200 import native, synth  # noqa: E402


def measure(ctx, circuit, n_rows, flags, reps=20):
    ctx.enable_timing(True)
    idx, chk = [], []
    for _ in range(reps + 3):
        ctx.invalidate_indexes()
        ctx.check_async(circuit, 0, n_rows, 0, flags)
        a, b = ctx.last_timing()
        idx.append(a); chk.append(b)
    ctx.enable_timing(False)
    ff, fc = ctx.fetch_result(circuit)
    assert (ff == native.PASS).all(), native.first_failure(ff, circuit)
    return float(np.mean(idx[3:])), float(np.mean(chk[3:]))


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    ctx = native.Context(0)
    # state circuit, cfg3
    w = synth.state_rows(1 << 18, seed=3)
    ctx.upload_table(native.TABLE_MPT, w["mpt"])
    ctx.upload_columns(native.CIRCUIT_STATE, w["rows"], flags=w["flags"])
    n = w["rows"].shape[1]
    i_ms, c_ms = measure(ctx, native.CIRCUIT_STATE, n, native.FLAG_WRAP)
    byt = 32 * (n * 57 + w["mpt"].shape[1] * 12)
    print(json.dumps({"circuit": "state (cfg3)", "rows": n, "mpt_rows": int(w["mpt"].shape[1]), "index_ms": i_ms,
                      "check_ms": c_ms, "rows_per_s": n / ((i_ms + c_ms) / 1e3), "algorithmic_bytes": byt,
                      "achieved_gbs": byt / (c_ms / 1e3) / 1e9, "frac_of_measured_hbm": byt / (c_ms / 1e3) / 1e9 / peak}))
    # copy circuit, cfg4: 512 events x 1024 bytes = 2^20 rows
    w = synth.copy_events(512, 1024, seed=4)
    ctx.set_challenge(native.CHALLENGE_KECCAK, sum(int(w["r"][k]) << (64 * k) for k in range(4)))
    ctx.upload_table(native.TABLE_RW, w["rw"], flags=w["rw_flags"])
    ctx.upload_table(native.TABLE_BYTECODE, w["bytecode"])
    ctx.upload_table(native.TABLE_TX, w["tx"], flags=w["tx_flags"])
    ctx.upload_columns(native.CIRCUIT_COPY, w["copy"], flags=w["copy_flags"])
    n = w["copy"].shape[1]
    i_ms, c_ms = measure(ctx, native.CIRCUIT_COPY, n, native.FLAG_WRAP)
    byt = 32 * (n * 20 + w["rw"].shape[1] * 14 + w["tx"].shape[1] * 5)
    print(json.dumps({"circuit": "copy (cfg4)", "rows": n, "rw_rows": int(w["rw"].shape[1]), "tx_rows": int(w["tx"].shape[1]),
                      "index_ms": i_ms, "check_ms": c_ms, "rows_per_s": n / ((i_ms + c_ms) / 1e3), "algorithmic_bytes": byt,
                      "achieved_gbs": byt / (c_ms / 1e3) / 1e9, "frac_of_measured_hbm": byt / (c_ms / 1e3) / 1e9 / peak}))


if __name__ == "__main__":
    main()
