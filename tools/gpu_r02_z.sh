#!/bin/bash
# round-2 call Z (HEAD): + ErrorGasUintOverflow — GPU tests (74 states), the 8,758 step-aux / error-state
# vectors through the C-ABI (also under compute-sanitizer memcheck), the per-kernel capture tied to the new source hash,
# bench line, launch list
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/z_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" $O/z_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " $O/z_gpu_tests.log | head -12
timeout 600 python tools/gpu_create_vectors.py > $O/z_create_vectors.log 2>&1; echo "create vectors rc=$?"; tail -2 $O/z_create_vectors.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_create_vectors.py 900 > $O/z_sanitizer_memcheck_create.log 2>&1; echo "memcheck create rc=$?"; grep -n "step-aux vectors ok\|ERROR SUMMARY" $O/z_sanitizer_memcheck_create.log | tail -3
bash tools/gpu_capture.sh z; python tools/capture_summary.py $O/z_metrics.csv $O/current_capture.json r02_z
cp $O/current_capture.json profiles/current_capture.json
timeout 300 python bench.py --workload block --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/z_block.json 2> $O/z_block.err
python - <<PY
import json
try:
    d=json.loads(open("$O/z_block.json").read().strip().splitlines()[-1]); print("block check", d["check_ms"], "ms/pass", d["ms_per_pass"])
except Exception as ex: print("block failed", ex)
PY
timeout 900 python bench.py > $O/z_bench.json 2> $O/z_bench.err; echo "bench rc=$?"; tail -3 $O/z_bench.err
python - <<PY
import json
d=json.loads(open("$O/z_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M rows/s" % (d["value"]/1e6), "ms/step", d["ms_per_step"], "check", r["kernel_ms"], "index", r["index_build_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"]["ms_per_step"])
print("traffic", r.get("traffic"), "dram_frac", r.get("dram_frac"))
for c in d.get("circuits", []): print(c["circuit"], c["ms_per_pass"], c["roofline"]["kernel_ms"], c["roofline"]["frac"])
print("block", d["block_trace"]["ms_per_pass"], d["block_trace"]["check_ms"], "typed", d["typed"]["kernel_ms"])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/z_launches.csv 2 > $O/z_launch_summary.txt 2>&1; grep k_evm $O/z_launch_summary.txt
