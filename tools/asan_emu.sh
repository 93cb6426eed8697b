#!/bin/bash
# Memory-safety pass over the kernel SOURCES on the CPU: the g++ emulation build (tests/emu) with AddressSanitizer,
# driven through the C ABI on awkward frame sizes (odd, non-multiple-of-4 widths, frames smaller than a tile), the
# randomized reductions, the pre-processing windows and the stage hooks.  Run here (no GPU): tools/asan_emu.sh [FFB_ITER_CFG]
# compute-sanitizer is not available on the GPU pool, so this is the out-of-bounds check of the tiling / halo / padding logic.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$(mktemp -d)
g++ -x c++ -std=c++17 -O1 -g -fPIC -shared -DFFB_EMU -mfma -ffp-contract=fast -fsanitize=address -fno-omit-frame-pointer \
    -Wno-unknown-pragmas -I "$ROOT/tests/emu" -I "$ROOT/funscript_flow_b200/csrc" -I "$ROOT/include" \
    -o "$OUT/libffb_emu_asan.so" "$ROOT/funscript_flow_b200/csrc/ffb_api.cu"
cat > "$OUT/run.py" <<PY
import sys
sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
import numpy as np
from funscript_flow_b200 import _native, api
from funscript_flow_b200.synth import make_clip
import parity_checks as pc
ctx = _native.FlowContext(0, lib_path="$OUT/libffb_emu_asan.so")
for (w, h) in [(256, 256), (150, 101), (33, 20), (264, 72), (97, 131)]:
    r = api.process_bracket(make_clip(w, h, 4, seed=w), {}, ctx=ctx, batch_frames=2)
    print(w, h, "ok", float(np.sum(r["scalar"])))
pc.check_reductions_random(ctx, n_cases=8, max_side=70)
pc.check_preprocess(ctx, sizes=((97, 131), (64, 48)))
pc.check_preprocess_window(ctx)
pc.check_native_resolution_bracket(ctx, 96, 64, 4)
pc.check_stages(ctx, 150, 101)
print("asan run complete")
PY
[ -n "$1" ] && export FFB_ITER_CFG=$1
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python "$OUT/run.py"
rm -rf "$OUT"
