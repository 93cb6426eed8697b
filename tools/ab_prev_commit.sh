#!/bin/bash
# A/B of the working tree against a previous commit IN ONE PROCESS on the GPU box.
#   here (no GPU):   tools/ab_prev_commit.sh build [REV]     # builds scratch/libffb_prev.so from REV (default HEAD)
#   on the GPU box:  gpurun -- 'python tools/ab_prev_commit.py > gpurun_out/ab.json'
# scratch/ is git-ignored but travels with the gpurun snapshot, like the other built libraries.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REV=${2:-HEAD}
if [ "$1" = "build" ]; then
  mkdir -p "$ROOT/scratch/x/csrc" "$ROOT/scratch/include"
  for f in ffb_api.cu ffb_kernels.cuh ffb_common.h; do git -C "$ROOT" show "$REV:funscript_flow_b200/csrc/$f" > "$ROOT/scratch/x/csrc/$f"; done
  git -C "$ROOT" show "$REV:include/ffb.h" > "$ROOT/scratch/include/ffb.h"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -shared -cudart shared \
       -I "$ROOT/scratch/include" -I "$ROOT/scratch/x/csrc" -o "$ROOT/scratch/libffb_prev.so" "$ROOT/scratch/x/csrc/ffb_api.cu"
  echo "built scratch/libffb_prev.so from $REV"
else
  echo "usage: $0 build [REV]"; exit 2
fi
