#!/bin/bash
# round-2 call O: launch-bound variants of PUSH (5 / 6 blocks) and MUL (4 / 5 blocks) after the fast paths, EVM GPU tests
# incl. ErrorOutOfGasCall (67 states)
O=gpurun_out
mkdir -p $O
for v in default p5 p6 g4 g5 default; do
  if [ $v = default ]; then L=$PWD/zkevm-specs_b200/libzkcheck.so; else L=$PWD/build_tune/libzkcheck_$v.so; fi
  ZKCHECK_LIB=$L timeout 300 python bench.py --steps 40 --no-extras --no-cpu-baseline --no-e2e > $O/o_${v}_evm.json 2> $O/o_${v}_evm.err
  python - <<PY
import json
try:
    e=json.loads(open("$O/o_${v}_evm.json").read().strip().splitlines()[-1]); print("$v", "evm check", e["roofline"]["kernel_ms"], "value", e["value"]/1e9)
except Exception as ex: print("$v", "failed", ex)
PY
done
timeout 1500 python -m pytest tests/test_gpu_evm.py -m gpu -q > $O/o_gpu_evm.log 2>&1; echo "pytest rc=$?"; tail -4 $O/o_gpu_evm.log
