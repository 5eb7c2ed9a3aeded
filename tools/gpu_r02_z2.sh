#!/bin/bash
# round-2 call Z2 (HEAD, 2 GPUs): the C-ABI collective tests and the 2-GPU bench line at the final sources (74 states)
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/z2_gpu_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/z2_gpu_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 30 --warmup 3 > $O/z2_bench_2gpu.json 2> $O/z2_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -3 $O/z2_bench_2gpu.err
python - <<PY
import json
d=json.loads(open("$O/z2_bench_2gpu.json").read().strip().splitlines()[-1])
print("2 GPUs: value %.2f G rows/s" % (d["value"]/1e9), "ms/step", d["ms_per_step"], "check", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"]/1e6, "traffic", d["roofline"].get("traffic"), d["roofline"].get("dram_frac"))
print("strong", d.get("strong_scaling")); print("cfg5", d.get("cfg5"))
PY
