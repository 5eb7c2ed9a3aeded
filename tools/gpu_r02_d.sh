#!/bin/bash
# round-2 call D: A/B of the lookup tail handling and the copy-kernel launch bounds (tuning builds under build_tune/),
# then the copy / fullsize GPU tests on the default build
O=gpurun_out
mkdir -p $O
for v in base unrollref cp4 cp5 cp6 dyn; do
  L=$PWD/build_tune/libzkcheck_$v.so
  ZKCHECK_LIB=$L timeout 300 python bench.py --steps 30 --no-extras --no-cpu-baseline --no-e2e > $O/d_${v}_evm.json 2> $O/d_${v}_evm.err
  for wl in copy state bytecode; do ZKCHECK_LIB=$L timeout 200 python bench.py --workload $wl --steps 20 > $O/d_${v}_$wl.json 2> $O/d_${v}_$wl.err; done
  python - <<PY
import json
def last(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
e=last("$O/d_${v}_evm.json"); c=last("$O/d_${v}_copy.json"); s=last("$O/d_${v}_state.json"); b=last("$O/d_${v}_bytecode.json")
print("$v", "evm check", e and e["roofline"]["kernel_ms"], "value", e and e["value"]/1e9, "| copy", c and c["roofline"]["kernel_ms"], c and c["roofline"]["frac"], "| state", s and s["roofline"]["kernel_ms"], "| bytecode", b and b["roofline"]["kernel_ms"])
PY
done
timeout 900 python -m pytest tests/test_gpu_copy.py tests/test_gpu_fullsize.py tests/test_gpu_packed.py tests/test_gpu_assign.py -m gpu -q -x > $O/d_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $O/d_gpu_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum --clock-control none -k regex:'k_check_copy' --launch-skip 8 -c 4 --csv --log-file $O/d_rowcirc_copy.csv python bench.py --workload copy --steps 3 > /dev/null 2>&1; echo "ncu copy rc=$?"
