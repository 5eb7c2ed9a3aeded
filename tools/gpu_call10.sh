#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c10_gpu_tests.log 2>&1; echo "pytest all rc=$?"; grep -n "passed\|failed" gpurun_out/c10_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " gpurun_out/c10_gpu_tests.log | head -12
for wl in bytecode state copy; do
  timeout 300 python bench.py --workload $wl --steps 20 > gpurun_out/c10_wl_$wl.json 2> gpurun_out/c10_wl_$wl.err; echo "$wl rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/c10_wl_$wl.json").read().strip().splitlines()[-1])
print("$wl", "ms/pass %.4f" % d["ms_per_pass"], "check %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"])
PY
done
timeout 600 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/c10_bench.json").read().strip().splitlines()[-1])
print("value %.1f M" % (d["value"]/1e6), "check", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"])
PY
