#!/bin/bash
# round-2 evidence call A: GPU tests, full bench line, reference arm, launch list, per-kernel metric capture,
# one `--set full` capture of the two dominant kernels, compute-sanitizer on the smoke path + small tests
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $O/a_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" $O/a_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " $O/a_gpu_tests.log | head -12
timeout 900 python bench.py > $O/a_bench.json 2> $O/a_bench.err; echo "bench rc=$?"; tail -3 $O/a_bench.err
python - <<PY
import json
d=json.loads(open("$O/a_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M rows/s" % (d["value"]/1e6), "ms/step", d["ms_per_step"], "check", r["kernel_ms"], "index", r["index_build_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"]["ms_per_step"])
for c in d.get("circuits", []): print(c["circuit"], c["ms_per_pass"], c["roofline"]["kernel_ms"], c["roofline"]["frac"])
print("block", d["block_trace"]["ms_per_pass"], d["block_trace"]["check_ms"], "typed", d["typed"]["kernel_ms"], "cfg5", d["cfg5"])
print("cpu", d["cpu_baseline"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/a_bench_reference.json 2> $O/a_bench_reference.err; echo "ref rc=$?"; tail -c 600 $O/a_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/a_launches.csv 2 > $O/a_launch_summary.txt 2>&1; cat $O/a_launch_summary.txt
bash tools/gpu_capture.sh a; python tools/capture_summary.py $O/a_metrics.csv $O/current_capture.json r02_a
timeout 500 ncu --set full --import-source on --clock-control none -k regex:'k_evm_push_pos|k_evm_gadget' --launch-skip 12 -c 4 -o $O/a_top_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/a_ncu_full.log 2>&1; echo "ncu full rc=$?"
for wl in state copy bytecode; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:'k_check_|k_state_' --launch-skip 4 -c 4 --csv --log-file $O/a_rowcirc_$wl.csv python bench.py --workload $wl --steps 3 > /dev/null 2>&1; echo "ncu $wl rc=$?"
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/a_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -4 $O/a_sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_bytecode.py tests/test_gpu_copy.py tests/test_gpu_state.py tests/test_gpu_exp.py tests/test_gpu_tx.py -m gpu -q -x > $O/a_sanitizer_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 $O/a_sanitizer_memcheck_tests.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/a_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -4 $O/a_sanitizer_racecheck_smoke.log
