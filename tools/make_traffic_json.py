"""Write profiles/roofline_traffic.json from an `ncu --set full` capture of k_flow_iter (tools/profile_round.sh):

    python tools/make_traffic_json.py gpurun_out/prof_iter_TAG.ncu-rep gpurun_out/bench_TAG.json

The record carries the hash of the kernel sources the capture was taken with (read from the bench line of the same
gpurun call); bench.py refuses the record when the sources have changed since."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, bench = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, k):
    v = float(r[ix[k]].replace(",", ""))
    u = units[ix[k]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3}.get(u, 1.0)
    return v * scale


# iterations 2 and 3 of level 0 (the launches without the fused up-sampling: second template flag false)
cand = [r for r in data if "k_flow_iter" in r[ix["Kernel Name"]]]
r = cand[-1]
line = json.loads(open(bench).read().strip().splitlines()[-1])
sha = line["run"]["kernels_sha"]
grid = r[ix["launch__grid_size"]]
rd, wr, ms = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
pairs = int(grid) // 32            # 1080p, 256x4x8: 8 strips x 4 row segments per pair
alg = pairs * 56.0 * 1920 * 1080
rec = {
    "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, B200, tools/profile_round.sh)",
    "kernels_sha": sha,
    "kernel": r[ix["Kernel Name"]].split("(")[0],
    "launch_shape": f"level 0 (1920x1080), one {pairs}-pair launch (capture taken with FFB_FLOW_STREAMS=1), grid size {grid}, iterations 2-3",
    "alg_bytes_per_launch": alg,
    "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr, "ncu_duration_ms": ms,
    "actual_dram_gbs": (rd + wr) / ms / 1e6, "algorithmic_gbs": alg / ms / 1e6,
    "note": "traffic is BELOW the algorithmic bytes: pair-fastest launch order lets frame j+1's expansion (R1 of pair j, R0 of pair j+1) hit L2 the second time",
}
json.dump(rec, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(rec, indent=1))
