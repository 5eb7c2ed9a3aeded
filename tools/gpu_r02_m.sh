#!/bin/bash
# round-2 evidence call M (HEAD): all GPU tests, the full bench line, the reference arm, launch lists (EVM + whole-block
# workload), the per-kernel metric capture tied to the source hash, one `--set full` capture of the hot kernels,
# launch-bound variants of ADD / POP, compute-sanitizer memcheck (smoke + golden tests of every circuit) and racecheck
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/m_smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/m_gpu_tests.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" $O/m_gpu_tests.log | tail -2; grep -n "^FAILED\|^E   " $O/m_gpu_tests.log | head -12
timeout 900 python bench.py > $O/m_bench.json 2> $O/m_bench.err; echo "bench rc=$?"; tail -3 $O/m_bench.err
python - <<PY
import json
d=json.loads(open("$O/m_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M rows/s" % (d["value"]/1e6), "ms/step", d["ms_per_step"], "check", r["kernel_ms"], "index", r["index_build_ms"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "serial", d["e2e"]["serial"]["ms_per_step"])
for c in d.get("circuits", []): print(c["circuit"], c["ms_per_pass"], c["roofline"]["kernel_ms"], c["roofline"]["frac"])
print("block", d["block_trace"]["ms_per_pass"], d["block_trace"]["check_ms"], "typed", d["typed"]["kernel_ms"], "cfg5", d["cfg5"])
print("assign", d.get("assign"))
print("cpu", d["cpu_baseline"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/m_bench_reference.json 2> $O/m_bench_reference.err; echo "ref rc=$?"; tail -c 400 $O/m_bench_reference.json
for v in a5p5 a6p6; do
  ZKCHECK_LIB=$PWD/build_tune/libzkcheck_$v.so timeout 300 python bench.py --steps 30 --no-extras --no-cpu-baseline --no-e2e > $O/m_${v}_evm.json 2> $O/m_${v}_evm.err
  python - <<PY
import json
try:
    e=json.loads(open("$O/m_${v}_evm.json").read().strip().splitlines()[-1]); print("$v", "evm check", e["roofline"]["kernel_ms"], "value", e["value"]/1e9)
except Exception as ex: print("$v", "failed", ex)
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/m_launches.csv 2 > $O/m_launch_summary.txt 2>&1; cat $O/m_launch_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/m_launches_block.csv python bench.py --workload block --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu block launches rc=$?"
python tools/launch_summary.py $O/m_launches_block.csv 2 > $O/m_launch_summary_block.txt 2>&1; cat $O/m_launch_summary_block.txt
bash tools/gpu_capture.sh m; python tools/capture_summary.py $O/m_metrics.csv $O/current_capture.json r02_m
timeout 500 ncu --set full --import-source on --clock-control none -k regex:'k_evm_push_pos|k_evm_gadget|k_evm_classify' --launch-skip 15 -c 5 -o $O/m_top_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/m_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/m_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 $O/m_sanitizer_memcheck_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_evm.py tests/test_gpu_bytecode.py tests/test_gpu_copy.py tests/test_gpu_state.py tests/test_gpu_exp.py tests/test_gpu_tx.py tests/test_gpu_pi.py tests/test_gpu_assign.py -m gpu -q -k "golden or parity or oracle" > $O/m_sanitizer_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 $O/m_sanitizer_memcheck_tests.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/m_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 $O/m_sanitizer_racecheck_smoke.log
