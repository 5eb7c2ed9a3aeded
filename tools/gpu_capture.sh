#!/bin/bash
# ncu metrics capture of the bench command -> gpurun_out/<tag>_metrics.csv (+ optional full report of the top kernels)
tag=${1:-cap}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none -k regex:k_evm_ --csv --log-file gpurun_out/${tag}_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2> gpurun_out/${tag}_metrics.err
echo "metrics rc=$?"
