#!/bin/bash
# round-2 call B: public-inputs circuit on the GPU (parity tests, bench line, metric capture)
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pi.py -m gpu -q -x > $O/b_gpu_pi.log 2>&1; echo "pytest rc=$?"; tail -5 $O/b_gpu_pi.log
timeout 600 python bench.py --workload pi --steps 20 > $O/b_wl_pi.json 2> $O/b_wl_pi.err; echo "pi rc=$?"; tail -c 1200 $O/b_wl_pi.json; tail -3 $O/b_wl_pi.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:'k_check_pi' --launch-skip 4 -c 2 --csv --log-file $O/b_rowcirc_pi.csv python bench.py --workload pi --steps 3 > /dev/null 2>&1; echo "ncu pi rc=$?"
