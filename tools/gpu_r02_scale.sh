#!/bin/bash
# N-GPU bench line (weak-scaling inference + strong-scaling training leg); usage: gpurun --gpus N -- 'bash tools/gpu_r02_scale.sh N'
N=$1
mkdir -p gpurun_out/r02h
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02h/bench_${N}gpu.json 2> gpurun_out/r02h/bench_${N}gpu.err
tail -2 gpurun_out/r02h/bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02h/bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"])
print("train:", {k: v for k, v in d.get("train", {}).items() if k not in ("kernels_us", "workload")})
print("configs2:", {k: v for k, v in (d.get("configs2") or {}).items() if k != "workload"})
PY
