#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
for v in 2631b34 b51b280 head new; do
  ZKCHECK_LIB=build/variants/libzk_$v.so timeout 400 python bench.py --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/bisect_$v.json 2> gpurun_out/bisect_$v.err
  echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bisect_$v.json").read().strip().splitlines()[-1])
    print("$v", "value %.1f M rows/s" % (d["value"]/1e6), "ms/step %.3f" % d["ms_per_step"], "check %.3f" % d["roofline"]["kernel_ms"], "index %.3f" % d["roofline"]["index_build_ms"], d["clocks"])
except Exception as e: print("$v parse failed", e)
PY
done
for v in head new; do
  ZKCHECK_LIB=build/variants/libzk_$v.so timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/bisect_launches_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  echo "ncu $v rc=$?"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c1_gpu_tests.log
