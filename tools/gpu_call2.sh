#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c2_gpu_tests.log
for v in base a4p5 a4p4 g4; do
  ZKCHECK_LIB=build/variants/libzk_$v.so timeout 400 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c2_sweep_$v.json 2> gpurun_out/c2_sweep_$v.err
  echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c2_sweep_$v.json").read().strip().splitlines()[-1])
    print("$v", "value %.1f M rows/s" % (d["value"]/1e6), "ms/step %.3f" % d["ms_per_step"], "check %.3f" % d["roofline"]["kernel_ms"], "index %.3f" % d["roofline"]["index_build_ms"])
except Exception as e: print("$v parse failed", e)
PY
done
for wl in state copy bytecode; do
  timeout 300 python bench.py --workload $wl --steps 20 > gpurun_out/c2_wl_$wl.json 2> gpurun_out/c2_wl_$wl.err; echo "$wl rc=$?"; cat gpurun_out/c2_wl_$wl.json | cut -c1-600
done
timeout 900 python bench.py --steps 20 > gpurun_out/c2_bench_full.json 2> gpurun_out/c2_bench_full.err; echo "full rc=$?"; tail -c 3000 gpurun_out/c2_bench_full.json; tail -5 gpurun_out/c2_bench_full.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu rc=$?"
