#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bytecode.py tests/test_gpu_copy.py tests/test_gpu_evm.py -m gpu -q > gpurun_out/c11_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" gpurun_out/c11_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " gpurun_out/c11_gpu_tests.log | head -12
for wl in copy block; do
  timeout 300 python bench.py --workload $wl --steps 20 > gpurun_out/c11_wl_$wl.json 2> gpurun_out/c11_wl_$wl.err; echo "$wl rc=$?"; tail -c 900 gpurun_out/c11_wl_$wl.json; tail -3 gpurun_out/c11_wl_$wl.err
done
timeout 600 python bench.py --steps 20 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/c11_bench.json").read().strip().splitlines()[-1])
print("value %.1f M" % (d["value"]/1e6), "check", d["roofline"]["kernel_ms"])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/c11_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_evm_ --launch-skip 60 -c 60 --csv --log-file gpurun_out/c11_launches_block.csv python bench.py --workload block --steps 3 > /dev/null 2>&1; echo "ncu block rc=$?"
