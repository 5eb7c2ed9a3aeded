#!/bin/bash
# round-2 call R (HEAD with CALL_OP, 68 states): all GPU tests, bench line, launch list, the per-kernel capture tied to the
# source hash, memcheck of the EVM golden tests
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" $O/r_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " $O/r_gpu_tests.log | head -12
bash tools/gpu_capture.sh r; python tools/capture_summary.py $O/r_metrics.csv $O/current_capture.json r02_r
cp $O/current_capture.json profiles/current_capture.json
timeout 900 python bench.py > $O/r_bench.json 2> $O/r_bench.err; echo "bench rc=$?"; tail -3 $O/r_bench.err
python - <<PY
import json
d=json.loads(open("$O/r_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M rows/s" % (d["value"]/1e6), "ms/step", d["ms_per_step"], "check", r["kernel_ms"], "index", r["index_build_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"]["ms_per_step"])
print("traffic", r.get("traffic"), "dram_frac", r.get("dram_frac"), "frac", r.get("frac"), "stored_frac", r.get("stored_frac"))
for c in d.get("circuits", []): print(c["circuit"], c["ms_per_pass"], c["roofline"]["kernel_ms"], c["roofline"]["frac"])
print("block", d["block_trace"]["ms_per_pass"], d["block_trace"]["check_ms"], "typed", d["typed"]["kernel_ms"])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/r_launches.csv 2 > $O/r_launch_summary.txt 2>&1; grep k_evm $O/r_launch_summary.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_evm.py -m gpu -q -k "golden" > $O/r_sanitizer_memcheck_evm.log 2>&1; echo "memcheck evm rc=$?"; tail -3 $O/r_sanitizer_memcheck_evm.log
