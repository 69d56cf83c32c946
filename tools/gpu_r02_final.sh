#!/bin/bash
# final single-GPU records: whole GPU suite, smoke, default bench line
mkdir -p gpurun_out/r02f
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02f/pytest_gpu.txt 2>&1
echo "exit $?" >> gpurun_out/r02f/pytest_gpu.txt
grep -v "^ \|^$\|^>" gpurun_out/r02f/pytest_gpu.txt | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02f/bench.json 2> gpurun_out/r02f/bench.err
tail -2 gpurun_out/r02f/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02f/bench_reference.json 2>> gpurun_out/r02f/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("dominant_tensor_kernel"))
print("clocks", d["clocks"])
print("train:", {k: v for k, v in d.get("train", {}).items() if k != "kernels_us"})
print("planner:", {k: d["planner"][k] for k in ("latency_ms_1_scene", "scenes_per_sec_batched")})
print("cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
r = json.loads(open("gpurun_out/r02f/bench_reference.json").read().strip().splitlines()[-1])
print("reference arm:", r.get("value"), r.get("impl"))
PY
