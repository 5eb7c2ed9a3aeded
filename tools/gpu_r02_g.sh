#!/bin/bash
# round-2 call G: narrow instances of the hot EVM kernels (StepCtx::narrow): bench main line + the EVM / packed / full-size tests
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --steps 30 --no-extras --no-cpu-baseline --no-e2e > $O/g_evm.json 2> $O/g_evm.err; echo "bench rc=$?"; tail -2 $O/g_evm.err
python - <<PY
import json
e=json.loads(open("$O/g_evm.json").read().strip().splitlines()[-1])
print("evm check", e["roofline"]["kernel_ms"], "value", e["value"]/1e9, "launches", e["gpu_launches"])
PY
timeout 300 python bench.py --steps 30 --no-extras --no-cpu-baseline --no-e2e --storage typed > $O/g_evm_typed.json 2> $O/g_evm_typed.err; echo "typed rc=$?"
python - <<PY
import json
e=json.loads(open("$O/g_evm_typed.json").read().strip().splitlines()[-1])
print("typed: evm check", e["roofline"]["kernel_ms"], "value", e["value"]/1e9)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/g_launches.csv 2 | grep "k_evm"
timeout 1200 python -m pytest tests/test_gpu_evm.py tests/test_gpu_packed.py tests/test_gpu_fullsize.py -m gpu -q > $O/g_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 $O/g_gpu_tests.log
