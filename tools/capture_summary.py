#!/usr/bin/env python
"""profiles/current_capture.json from an ncu metrics log of the bench command (see tools/gpu_capture.sh):

    ncu --metrics <METRICS> --clock-control none -k regex:k_evm_ --csv --log-file X.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras

Takes the LAST launch of every check-phase kernel (k_evm_classify, k_evm_scatter, the gate-program
kernels) = one check, and writes per-kernel duration, DRAM bytes, warp instructions, issue-active %,
warps-active %, registers, local-load sectors plus their sums, tagged with the hash of the CUDA
sources so that bench.py only quotes it for the build it was taken from.
usage: capture_summary.py <csv> <out.json> [capture-name]"""
import collections
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_hash() -> str:
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "zkevm-specs_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "usecond": 1e-3,
         "nsecond": 1e-6, "msecond": 1}
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, per = None, collections.OrderedDict()
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("zk::", "")
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    per.setdefault((int(d["ID"]), name), {})[d["Metric Name"]] = v * SCALE.get(d["Metric Unit"], 1)
last = collections.OrderedDict()
for (i, name), m in per.items():
    if name.startswith("k_evm_"):
        last[name] = m  # later launches overwrite earlier ones
short = {"gpu__time_duration.sum": "ms", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
         "smsp__inst_executed.sum": "warp_inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
         "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
         "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum": "local_load_sectors", "launch__grid_size": "grid"}
out = {"source_hash": source_hash(), "capture": sys.argv[3] if len(sys.argv) > 3 else os.path.basename(sys.argv[1]),
       "command": "ncu --clock-control none: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras",
       "kernels": {}}
tot_ms = tot_b = 0.0
for name, m in last.items():
    k = {short[a]: b for a, b in m.items() if a in short}
    out["kernels"][name] = k
    tot_ms += k.get("ms", 0)
    tot_b += k.get("dram_read", 0) + k.get("dram_write", 0)
out["kernel_ms_sum"] = tot_ms
out["dram_bytes_per_check"] = tot_b
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({"kernel_ms_sum": tot_ms, "dram_bytes_per_check": tot_b, "kernels": list(last)}))
