#!/bin/bash
# round-2 call U (HEAD): single-lane vote shortcut + lazy contract address in the transaction-level group — whole-block
# workload, EVM GPU tests, bench line, launch lists, the per-kernel capture tied to the source hash
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/u_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" $O/u_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " $O/u_gpu_tests.log | head -12
bash tools/gpu_capture.sh u; python tools/capture_summary.py $O/u_metrics.csv $O/current_capture.json r02_u
cp $O/current_capture.json profiles/current_capture.json
for ov in 1 0; do
  ZKCHECK_TX_OVERLAP=$ov timeout 300 python bench.py --workload block --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/u_block_ov$ov.json 2> $O/u_block_ov$ov.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/u_block_ov$ov.json").read().strip().splitlines()[-1]); print("overlap=$ov block check", d["check_ms"], "ms/pass", d["ms_per_pass"])
except Exception as ex: print("block overlap=$ov failed", ex)
PY
done
timeout 900 python bench.py > $O/u_bench.json 2> $O/u_bench.err; echo "bench rc=$?"; tail -3 $O/u_bench.err
python - <<PY
import json
d=json.loads(open("$O/u_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M rows/s" % (d["value"]/1e6), "ms/step", d["ms_per_step"], "check", r["kernel_ms"], "index", r["index_build_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"]["ms_per_step"])
print("traffic", r.get("traffic"), "dram_frac", r.get("dram_frac"))
for c in d.get("circuits", []): print(c["circuit"], c["ms_per_pass"], c["roofline"]["kernel_ms"], c["roofline"]["frac"])
print("block", d["block_trace"]["ms_per_pass"], d["block_trace"]["check_ms"], "typed", d["typed"]["kernel_ms"])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/u_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/u_launches.csv 2 > $O/u_launch_summary.txt 2>&1; grep k_evm $O/u_launch_summary.txt
ZKCHECK_TX_OVERLAP=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/u_launches_block.csv python bench.py --workload block --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu block launches rc=$?"
python tools/launch_summary.py $O/u_launches_block.csv 2 > $O/u_launch_summary_block.txt 2>&1; grep k_evm $O/u_launch_summary_block.txt
