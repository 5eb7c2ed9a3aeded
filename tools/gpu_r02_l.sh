#!/bin/bash
# round-2 call L: narrow classify (40 registers; block size A/B), EVM GPU tests incl. EXP and the code-store errors (65 states)
O=gpurun_out
mkdir -p $O
for v in default c1024 c256; do
  if [ $v = default ]; then L=$PWD/zkevm-specs_b200/libzkcheck.so; else L=$PWD/build_tune/libzkcheck_$v.so; fi
  ZKCHECK_LIB=$L timeout 300 python bench.py --steps 30 --no-extras --no-cpu-baseline --no-e2e > $O/l_${v}_evm.json 2> $O/l_${v}_evm.err
  python - <<PY
import json
try:
    e=json.loads(open("$O/l_${v}_evm.json").read().strip().splitlines()[-1]); print("$v", "evm check", e["roofline"]["kernel_ms"], "value", e["value"]/1e9)
except Exception as ex: print("$v", "failed", ex)
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/l_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/l_launches.csv 2 > $O/l_launch_summary.txt 2>&1; grep "k_evm" $O/l_launch_summary.txt
timeout 1500 python -m pytest tests/test_gpu_evm.py -m gpu -q > $O/l_gpu_evm.log 2>&1; echo "pytest rc=$?"; tail -4 $O/l_gpu_evm.log
