#!/bin/bash
# round-2 call S: block size of the transaction-level group (64 / 128 / 256 threads) on the whole-block workload, with the
# launch list of the block workload for the default
O=gpurun_out
mkdir -p $O
for v in default tx256 tx64 default; do
  if [ $v = default ]; then L=$PWD/zkevm-specs_b200/libzkcheck.so; else L=$PWD/build_tune/libzkcheck_$v.so; fi
  for ov in 1 0; do
    ZKCHECK_LIB=$L ZKCHECK_TX_OVERLAP=$ov timeout 300 python bench.py --workload block --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/s_block_${v}_ov$ov.json 2> $O/s_block_${v}_ov$ov.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/s_block_${v}_ov$ov.json").read().strip().splitlines()[-1]); print("$v overlap=$ov block check", d["check_ms"], "ms/pass", d["ms_per_pass"])
except Exception as ex: print("$v overlap=$ov failed", ex)
PY
  done
done
for v in tx256; do
ZKCHECK_LIB=$PWD/build_tune/libzkcheck_$v.so ZKCHECK_TX_OVERLAP=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/s_launches_block_$v.csv python bench.py --workload block --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu block launches rc=$?"
python tools/launch_summary.py $O/s_launches_block_$v.csv 2 2>&1 | grep "k_evm_group"
done
