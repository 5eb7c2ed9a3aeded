#!/bin/bash
# round-2 call V: software pipeline of the hot loops (next iteration's step cells prefetched to L1 / L2, the index of the
# iteration after next loaded early) vs the default, EVM workload; GPU tests of the better variant
O=gpurun_out
mkdir -p $O
for v in default pf1 pf2 default pf1; do
  if [ $v = default ]; then L=$PWD/zkevm-specs_b200/libzkcheck.so; else L=$PWD/build_tune/libzkcheck_$v.so; fi
  ZKCHECK_LIB=$L timeout 300 python bench.py --steps 40 --no-extras --no-cpu-baseline --no-e2e > $O/v_${v}_evm.json 2> $O/v_${v}_evm.err
  python - <<PY
import json
try:
    e=json.loads(open("$O/v_${v}_evm.json").read().strip().splitlines()[-1]); print("$v", "evm check", e["roofline"]["kernel_ms"], "value", e["value"]/1e9)
except Exception as ex: print("$v", "failed", ex)
PY
done
for v in pf1 pf2; do
ZKCHECK_LIB=$PWD/build_tune/libzkcheck_$v.so timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/v_launches_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches $v rc=$?"
python tools/launch_summary.py $O/v_launches_$v.csv 2 2>&1 | grep k_evm
done
ZKCHECK_LIB=$PWD/build_tune/libzkcheck_pf1.so timeout 1500 python -m pytest tests/test_gpu_evm.py tests/test_gpu_packed.py tests/test_gpu_fullsize.py -m gpu -q > $O/v_gpu_tests_pf1.log 2>&1; echo "pytest pf1 rc=$?"; tail -3 $O/v_gpu_tests_pf1.log
