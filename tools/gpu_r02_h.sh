#!/bin/bash
# round-2 call H (2 GPUs): the C-ABI collective test and the bench line under torchrun (weak main line + strong scaling,
# cfg4 / cfg5 sharded), then the single-GPU EVM tests incl. evm14
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/h_gpu_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 $O/h_gpu_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 3 > $O/h_bench_2gpu.json 2> $O/h_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -3 $O/h_bench_2gpu.err
python - <<PY
import json
d=json.loads(open("$O/h_bench_2gpu.json").read().strip().splitlines()[-1])
print("2 GPUs: value %.2f G rows/s" % (d["value"]/1e9), "ms/step", d["ms_per_step"], "check", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"]/1e6)
print("strong", d["strong_scaling"]); print("cfg4", d["cfg4"]); print("cfg5", d["cfg5"])
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 3 --scaling strong --no-extras --no-cpu-baseline > $O/h_bench_2gpu_strong.json 2> $O/h_bench_2gpu_strong.err; echo "bench strong rc=$?"; tail -c 400 $O/h_bench_2gpu_strong.json | head -c 400
timeout 900 python -m pytest tests/test_gpu_evm.py -m gpu -q > $O/h_gpu_evm.log 2>&1; echo "pytest evm rc=$?"; tail -3 $O/h_gpu_evm.log
