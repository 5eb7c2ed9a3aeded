#!/bin/bash
# round-2 bisect of the check-phase regression: same bench.py, four builds of libzkcheck.so
mkdir -p gpurun_out
for v in "$@"; do
  ZKCHECK_LIB=build/variants/libzk_$v.so timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/bisect_$v.json 2> gpurun_out/bisect_$v.err
  echo "$v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bisect_$v.json").read().strip().splitlines()[-1])
    print("$v", "value %.1f M rows/s" % (d["value"]/1e6), "ms/step %.3f" % d["ms_per_step"], "check %.3f" % d["roofline"]["kernel_ms"], "index %.3f" % d["roofline"]["index_build_ms"], d["clocks"])
except Exception as e: print("$v parse failed", e)
PY
done
for v in "$@"; do
  ZKCHECK_LIB=build/variants/libzk_$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/bisect_launches_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
