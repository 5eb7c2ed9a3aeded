#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c3_gpu_tests.log
timeout 600 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/c3_bench.json").read().strip().splitlines()[-1])
print("value %.1f M" % (d["value"]/1e6), "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"])
PY
tail -3 gpurun_out/c3_bench.err
for wl in bytecode copy state; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_check_ --launch-skip 8 -c 1 -o gpurun_out/c3_$wl python bench.py --workload $wl --steps 3 > /dev/null 2> gpurun_out/c3_ncu_$wl.err; echo "ncu $wl rc=$?"
done
ls -la gpurun_out/*.ncu-rep
