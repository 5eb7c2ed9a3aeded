#!/bin/bash
# round-2 call Z4 (HEAD, 4 GPUs): the bench line as the driver launches it at N=4 (weak scaling, no data-path collective)
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > $O/z4_bench_4gpu.json 2> $O/z4_bench_4gpu.err; echo "bench 4gpu rc=$?"; tail -2 $O/z4_bench_4gpu.err
python - <<PY
import json
d=json.loads(open("$O/z4_bench_4gpu.json").read().strip().splitlines()[-1])
print("4 GPUs: value %.2f G rows/s" % (d["value"]/1e9), "ms/step", d["ms_per_step"], "check", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"]/1e6, "scaling", d["scaling"], "n_gpus", d["n_gpus"])
print("strong", d.get("strong_scaling"))
PY
