#!/bin/bash
# 2 GPUs: NCCL multi-rank tests (sharded arg-max all-gather, data-parallel training step) + the 2-GPU bench line
mkdir -p gpurun_out/r02g
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -m gpu > gpurun_out/r02g/pytest_multirank.txt 2>&1
echo "exit $?" >> gpurun_out/r02g/pytest_multirank.txt
grep -v "^ \|^$\|^>" gpurun_out/r02g/pytest_multirank.txt | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02g/bench_2gpu.json 2> gpurun_out/r02g/bench_2gpu.err
tail -3 gpurun_out/r02g/bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02g/bench_2gpu.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"])
print("train:", {k: v for k, v in d.get("train", {}).items() if k != "kernels_us"})
print("configs2:", d.get("configs2"))
PY
