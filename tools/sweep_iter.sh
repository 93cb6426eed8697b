#!/bin/bash
# tuning sweep of the fused iteration kernel variants (run on the GPU box)
# usage: tools/sweep_iter.sh "128x2 128x4" "120 270" "16 32"
mkdir -p gpurun_out
for cfg in ${1:-128x2 128x4 192x2 256x2}; do
  for sh in ${2:-120 270}; do
    for bf in ${3:-16}; do
    FFB_ITER_CFG=$cfg FFB_ITER_SH=$sh timeout 300 python bench.py --steps 5 --warmup 2 --batch-frames $bf --no-cpu-baseline > gpurun_out/sweep_${cfg}_${sh}_${bf}.log 2>&1
    python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/sweep_${cfg}_${sh}_${bf}.log").read().strip().splitlines()[-1])
    print("${cfg} sh=${sh} bf=${bf}", "value %.0f e2e %.0f"%(l["value"], l["e2e"]["value"]), "iter frac %.3f"%l["roofline"]["frac"], l["kernel_ms_per_step"])
except Exception as e:
    print("${cfg} sh=${sh} bf=${bf} FAILED", e)
PY
    done
  done
done
