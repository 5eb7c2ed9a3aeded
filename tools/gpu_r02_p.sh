#!/bin/bash
# round-2 call P (HEAD): all GPU tests, whole-block workload with and without the auxiliary-stream overlap of the
# transaction-level group, the full bench line, launch list, the per-kernel capture tied to the source hash
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/p_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" $O/p_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " $O/p_gpu_tests.log | head -12
for ov in 1 0; do
  ZKCHECK_TX_OVERLAP=$ov timeout 300 python bench.py --workload block --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/p_block_ov$ov.json 2> $O/p_block_ov$ov.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/p_block_ov$ov.json").read().strip().splitlines()[-1]); print("overlap=$ov block ms/pass", d["ms_per_step"], "value", d["value"]/1e9, d.get("roofline",{}).get("kernel_ms"))
except Exception as ex: print("block overlap=$ov failed", ex)
PY
done
timeout 900 python bench.py > $O/p_bench.json 2> $O/p_bench.err; echo "bench rc=$?"; tail -3 $O/p_bench.err
python - <<PY
import json
d=json.loads(open("$O/p_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M rows/s" % (d["value"]/1e6), "ms/step", d["ms_per_step"], "check", r["kernel_ms"], "index", r["index_build_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"]["ms_per_step"])
for c in d.get("circuits", []): print(c["circuit"], c["ms_per_pass"], c["roofline"]["kernel_ms"], c["roofline"]["frac"])
print("block", d["block_trace"]["ms_per_pass"], d["block_trace"]["check_ms"], "typed", d["typed"]["kernel_ms"])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/p_launches.csv 2 > $O/p_launch_summary.txt 2>&1; grep k_evm $O/p_launch_summary.txt
bash tools/gpu_capture.sh p; python tools/capture_summary.py $O/p_metrics.csv $O/current_capture.json r02_p
