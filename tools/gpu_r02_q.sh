#!/bin/bash
# round-2 call Q (2 GPUs, HEAD): the C-ABI collective test and the bench line under torchrun (weak main line with strong
# scaling / cfg4 / cfg5 extras, then --scaling strong)
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/q_gpu_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/q_gpu_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 30 --warmup 3 > $O/q_bench_2gpu.json 2> $O/q_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -3 $O/q_bench_2gpu.err
python - <<PY
import json
d=json.loads(open("$O/q_bench_2gpu.json").read().strip().splitlines()[-1])
print("2 GPUs: value %.2f G rows/s" % (d["value"]/1e9), "ms/step", d["ms_per_step"], "check", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"]/1e6, "traffic", d["roofline"].get("traffic"), d["roofline"].get("dram_frac"))
print("strong", d["strong_scaling"]); print("cfg4", d["cfg4"]); print("cfg5", d["cfg5"])
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 2 --steps 30 --warmup 3 --scaling strong --no-extras --no-cpu-baseline > $O/q_bench_2gpu_strong.json 2> $O/q_bench_2gpu_strong.err; echo "bench strong rc=$?"; tail -c 300 $O/q_bench_2gpu_strong.json
