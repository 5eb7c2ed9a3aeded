#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c12_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" gpurun_out/c12_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " gpurun_out/c12_gpu_tests.log | head -12
for wl in copy state bytecode block; do
  timeout 300 python bench.py --workload $wl --steps 20 > gpurun_out/c12_wl_$wl.json 2> gpurun_out/c12_wl_$wl.err; echo "$wl rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/c12_wl_$wl.json").read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print("$wl ms/pass %.4f" % d["ms_per_pass"], "check", r.get("kernel_ms", d.get("check_ms")), "frac", r.get("frac"))
PY
  tail -3 gpurun_out/c12_wl_$wl.err
done
timeout 600 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/c12_bench.json").read().strip().splitlines()[-1])
print("value %.1f M" % (d["value"]/1e6), "check", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"].get("ms_per_step"))
PY
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_evm_group --launch-skip 1 -c 1 -o gpurun_out/c12_kgtx -f python bench.py --workload block --steps 3 > gpurun_out/c12_ncu_kgtx.log 2>&1; echo "ncu kgtx rc=$?"
