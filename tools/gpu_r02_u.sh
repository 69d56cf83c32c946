#!/bin/bash
mkdir -p gpurun_out/r02u
timeout 600 python -m pytest tests/test_gpu_tsdf.py tests/test_gpu_train_native.py -q -m gpu > gpurun_out/r02u/pytest_tsdf_train.txt 2>&1
echo "exit $?" >> gpurun_out/r02u/pytest_tsdf_train.txt
grep -v "^ \|^$\|^>" gpurun_out/r02u/pytest_tsdf_train.txt | tail -30
timeout 900 python bench.py > gpurun_out/r02u/bench.json 2> gpurun_out/r02u/bench.err
tail -3 gpurun_out/r02u/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02u/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"])
print("train:", {k: v for k, v in d.get("train", {}).items() if k != "kernels_us"})
print("kernels:", {k: v["us"] for k, v in d["kernels"].items()})
PY
