#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bytecode.py tests/test_gpu_state.py tests/test_gpu_packed.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/c5_gpu_tests_a.log 2>&1; echo "pytest subset rc=$?"; tail -3 gpurun_out/c5_gpu_tests_a.log; grep -n "NativeError" gpurun_out/c5_gpu_tests_a.log | head -3
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c5_gpu_tests.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/c5_gpu_tests.log
for wl in bytecode state copy; do
  timeout 300 python bench.py --workload $wl --steps 20 > gpurun_out/c5_wl_$wl.json 2> gpurun_out/c5_wl_$wl.err; echo "$wl rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/c5_wl_$wl.json").read().strip().splitlines()[-1])
print("$wl", "ms/pass %.4f" % d["ms_per_pass"], "check %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"])
PY
done
