#!/bin/bash
# round-2 call I: EVM GPU tests incl. evm15 / evm16 / evm17 (58 states), the EVM workload of the bench (no regression in the hot
# kernels after the new groups), memcheck of the EVM golden tests
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_evm.py -m gpu -q > $O/i_gpu_evm.log 2>&1; echo "pytest evm rc=$?"; tail -5 $O/i_gpu_evm.log
timeout 600 python bench.py --workload evm --steps 30 --warmup 3 --no-extras --no-cpu-baseline > $O/i_evm.json 2> $O/i_evm.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/i_evm.json").read().strip().splitlines()[-1])
print("evm check", d["roofline"]["kernel_ms"], "value", d["value"]/1e9, "launches", d["gpu_launches"])
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_evm.py -m gpu -q -k "golden" > $O/i_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 $O/i_memcheck.log
