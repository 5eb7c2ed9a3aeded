#!/bin/bash
# round-2 call J: EVM GPU tests over every golden part (61 states) and one `--set full` capture of the narrow hot kernels +
# classify at HEAD (read back here with ncu -i)
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_evm.py -m gpu -q > $O/j_gpu_evm.log 2>&1; echo "pytest evm rc=$?"; tail -5 $O/j_gpu_evm.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_evm_push_pos|k_evm_gadget|k_evm_classify' --launch-skip 15 -c 5 -o $O/j_top_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/j_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/j_top_full.ncu-rep
