#!/bin/bash
# Round evidence on the GPU box: default bench line, ncu launch list, and one `ncu --set full` capture per
# kernel of the hot path.  Usage (through gpurun): tools/profile_round.sh TAG  -> gpurun_out/*_TAG.*
# The captures run with one flow stream so that every launch covers the whole batch (128 pairs / 129 frames at 1080p).
# Afterwards, here: python tools/make_traffic_json.py gpurun_out/prof_iter_TAG.ncu-rep gpurun_out/bench_TAG.json
set -u
TAG=${1:-rX}
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo "bench failed"; tail -5 gpurun_out/bench_$TAG.err; exit 1; }
$B > gpurun_out/plain_$TAG.log 2>&1 || { echo "short bench failed"; exit 1; }     # exits 0 without ncu first
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1
export FFB_FLOW_STREAMS=1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_flow_iter -s 9 -c 3 -o gpurun_out/prof_iter_$TAG $B > gpurun_out/ncu_iter_$TAG.log 2>&1
$NCU -k regex:k_polyexp -c 4 -o gpurun_out/prof_poly_$TAG $B > gpurun_out/ncu_poly_$TAG.log 2>&1
$NCU -k regex:k_pyramid_pow2 -c 1 -o gpurun_out/prof_pyr_$TAG $B > gpurun_out/ncu_pyr_$TAG.log 2>&1
$NCU -k regex:k_divmag -c 1 -o gpurun_out/prof_div_$TAG $B > gpurun_out/ncu_div_$TAG.log 2>&1
$NCU -k "regex:^k_radial$" -s 1 -c 1 -o gpurun_out/prof_rad_$TAG $B > gpurun_out/ncu_rad_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.*
