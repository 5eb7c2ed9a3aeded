#!/bin/bash
# round-2 call F: all GPU tests (incl. the evm12 / evm13 error-state vectors) and the full bench line on the current build
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/f_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -6 $O/f_gpu_tests.log
timeout 900 python bench.py > $O/f_bench.json 2> $O/f_bench.err; echo "bench rc=$?"; tail -3 $O/f_bench.err
python - <<PY
import json
d=json.loads(open("$O/f_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M rows/s" % (d["value"]/1e6), "ms/step", d["ms_per_step"], "check", r["kernel_ms"], "index", r["index_build_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"]["ms_per_step"])
for c in d.get("circuits", []): print(c["circuit"], c["ms_per_pass"], c["roofline"]["kernel_ms"], c["roofline"]["frac"])
print("block", d["block_trace"]["ms_per_pass"], d["block_trace"]["check_ms"], "typed", d["typed"]["kernel_ms"], "cfg5", d["cfg5"])
print("assign", d.get("assign"))
PY
