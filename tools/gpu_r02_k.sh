#!/bin/bash
# round-2 call K: aggregate fast paths (PUSH bytes decided as a whole, MUL / DIV / MOD as 256-bit integer statements) —
# launch-bound A/B on the EVM workload, launch list of the default build, EVM GPU tests
O=gpurun_out
mkdir -p $O
for v in default g4 m4 p5 p6 a4p4; do
  if [ $v = default ]; then L=$PWD/zkevm-specs_b200/libzkcheck.so; else L=$PWD/build_tune/libzkcheck_$v.so; fi
  ZKCHECK_LIB=$L timeout 300 python bench.py --steps 30 --no-extras --no-cpu-baseline --no-e2e > $O/k_${v}_evm.json 2> $O/k_${v}_evm.err
  python - <<PY
import json
try:
    e=json.loads(open("$O/k_${v}_evm.json").read().strip().splitlines()[-1]); print("$v", "evm check", e["roofline"]["kernel_ms"], "value", e["value"]/1e9)
except Exception as ex: print("$v", "failed", ex)
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/k_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/k_launches.csv 2 > $O/k_launch_summary.txt 2>&1; grep "k_evm" $O/k_launch_summary.txt
timeout 1500 python -m pytest tests/test_gpu_evm.py tests/test_gpu_packed.py -m gpu -q > $O/k_gpu_evm.log 2>&1; echo "pytest rc=$?"; tail -4 $O/k_gpu_evm.log
