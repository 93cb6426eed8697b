#!/bin/bash
# tuning sweep of the fused iteration kernel variants (run on the GPU box)
mkdir -p gpurun_out
for cfg in 128x4 128x2 128x1 192x2 192x4 256x2 256x4; do
  for sh in 120 270; do
    FFB_ITER_CFG=$cfg FFB_ITER_SH=$sh timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/sweep_${cfg}_${sh}.log 2>&1
    python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/sweep_${cfg}_${sh}.log").read().strip().splitlines()[-1])
    print("${cfg} sh=${sh}", "value %.0f e2e %.0f"%(l["value"], l["e2e"]["value"]), "iter frac %.3f"%l["roofline"]["frac"], l["kernel_ms_per_step"])
except Exception as e:
    print("${cfg} sh=${sh} FAILED", e)
PY
  done
done
