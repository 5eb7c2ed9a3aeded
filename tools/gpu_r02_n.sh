#!/bin/bash
# round-2 call N (HEAD): all GPU tests (66 states), the full bench line, launch lists (EVM + whole-block workload), the
# per-kernel metric capture tied to the source hash, memcheck of the EVM golden tests
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/n_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" $O/n_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " $O/n_gpu_tests.log | head -12
timeout 900 python bench.py > $O/n_bench.json 2> $O/n_bench.err; echo "bench rc=$?"; tail -3 $O/n_bench.err
python - <<PY
import json
d=json.loads(open("$O/n_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M rows/s" % (d["value"]/1e6), "ms/step", d["ms_per_step"], "check", r["kernel_ms"], "index", r["index_build_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"]["ms_per_step"])
for c in d.get("circuits", []): print(c["circuit"], c["ms_per_pass"], c["roofline"]["kernel_ms"], c["roofline"]["frac"])
print("block", d["block_trace"]["ms_per_pass"], d["block_trace"]["check_ms"], "typed", d["typed"]["kernel_ms"])
print("traffic", r.get("traffic"), r.get("dram_frac"), r.get("frac"), r.get("stored_frac"))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/n_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/n_launches.csv 2 > $O/n_launch_summary.txt 2>&1; grep k_evm $O/n_launch_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/n_launches_block.csv python bench.py --workload block --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu block launches rc=$?"
python tools/launch_summary.py $O/n_launches_block.csv 2 > $O/n_launch_summary_block.txt 2>&1; grep k_evm $O/n_launch_summary_block.txt
bash tools/gpu_capture.sh n; python tools/capture_summary.py $O/n_metrics.csv $O/current_capture.json r02_n
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_evm.py -m gpu -q -k "golden" > $O/n_sanitizer_memcheck_evm.log 2>&1; echo "memcheck evm rc=$?"; tail -3 $O/n_sanitizer_memcheck_evm.log
