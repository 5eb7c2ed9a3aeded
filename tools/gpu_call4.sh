#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bytecode.py tests/test_gpu_state.py tests/test_gpu_packed.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/c4_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c4_gpu_tests.log
for wl in bytecode state copy; do
  timeout 300 python bench.py --workload $wl --steps 20 > gpurun_out/c4_wl_$wl.json 2> gpurun_out/c4_wl_$wl.err; echo "$wl rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/c4_wl_$wl.json").read().strip().splitlines()[-1])
print("$wl", "ms/pass %.4f" % d["ms_per_pass"], "check %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_check_\|k_state --csv --log-file gpurun_out/c4_launches_state.csv python bench.py --workload state --steps 3 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_check_bytecode --launch-skip 8 -c 1 -o gpurun_out/c4_bytecode python bench.py --workload bytecode --steps 3 > /dev/null 2> gpurun_out/c4_ncu_bytecode.err; echo "ncu rc=$?"
