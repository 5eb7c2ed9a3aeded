#!/bin/bash
# round-2 call E: after the local-array fix in lookup.cuh — launch-bound A/B (tuning builds), then all GPU tests and the
# full bench line on the default build
O=gpurun_out
mkdir -p $O
for v in default g4 a4p4 p5 p3 c5s4 c3s2; do
  if [ $v = default ]; then L=$PWD/zkevm-specs_b200/libzkcheck.so; else L=$PWD/build_tune/libzkcheck_$v.so; fi
  ZKCHECK_LIB=$L timeout 300 python bench.py --steps 30 --no-extras --no-cpu-baseline --no-e2e > $O/e_${v}_evm.json 2> $O/e_${v}_evm.err
  for wl in copy state; do ZKCHECK_LIB=$L timeout 200 python bench.py --workload $wl --steps 20 > $O/e_${v}_$wl.json 2> $O/e_${v}_$wl.err; done
  python - <<PY
import json
def last(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
e=last("$O/e_${v}_evm.json"); c=last("$O/e_${v}_copy.json"); s=last("$O/e_${v}_state.json")
print("$v", "evm check", e and e["roofline"]["kernel_ms"], "value", e and e["value"]/1e9, "| copy", c and c["roofline"]["kernel_ms"], c and c["roofline"]["frac"], "| state", s and s["roofline"]["kernel_ms"], s and s["roofline"]["frac"])
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/e_launches.csv 2 > $O/e_launch_summary.txt 2>&1; cat $O/e_launch_summary.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/e_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 $O/e_gpu_tests.log
