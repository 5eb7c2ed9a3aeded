#!/bin/bash
# round-2 call T: one `--set full` capture of k_evm_group<TX> on the whole-block workload (read back here with ncu -i)
O=gpurun_out
mkdir -p $O
ZKCHECK_TX_OVERLAP=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_evm_group' --launch-skip 2 -c 1 -o $O/t_tx_full -f python bench.py --workload block --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > $O/t_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/t_tx_full.ncu-rep
