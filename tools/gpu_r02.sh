#!/bin/bash
# One GPU-box visit of round 2.  usage (under gpurun): bash tools/gpu_r02.sh <tag> [quick-test-expr] [skip_full]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ -n "$2" ]; then
  echo "== quick: $2" | tee $OUT/pytest_quick.txt
  timeout 600 python -m pytest tests -x -q -m gpu -k "$2" 2>&1 | tail -25 | tee -a $OUT/pytest_quick.txt
fi
if [ -z "$3" ]; then
  echo "== pytest -m gpu" | tee $OUT/pytest_gpu.txt
  timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee -a $OUT/pytest_gpu.txt
  echo "== smoke" | tee $OUT/smoke.txt
  timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee -a $OUT/smoke.txt
fi
echo "== bench"
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
    for k, v in d["kernels"].items():
        print("  %-28s %8.2f us  share %.3f  %7.2f TF/s" % (k, v["us"], v["share"], v["tflops"]))
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "clocks", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $OUT/bench.err
echo "== decode timeline"
GIGA_TIMELINE=decode timeout 300 python tools/decode_timeline.py 2>&1 | tail -20 | tee $OUT/decode_timeline.txt
if [ -n "$NCU_K" ]; then
  echo "== ncu full: $NCU_K"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-3} -c ${NCU_C:-1} -o $OUT/prof -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu.log 2>&1
  tail -2 $OUT/ncu.log
fi
