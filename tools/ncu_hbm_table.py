"""Per-kernel HBM table from ncu --set full reports: duration, DRAM bytes, DRAM GB/s and % of the measured peak."""
import csv, json, subprocess, sys
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if len(sys.argv) < 2 or not sys.argv[1].replace('.', '').isdigit() else float(sys.argv[1])
reps = [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]
print(f"kernel                          grid      time_us  dram_read_MB dram_write_MB  dram_GB/s  %of_measured_peak({peak:.0f})  dram_active_pct  l1tex_pct  issue_pct")
for rep in reps:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, units = rows[0], rows[1]
    def col(name):
        return H.index(name)
    for r in rows[2:]:
        def val(name, scale=1.0):
            i = col(name)
            v = float(r[i].replace(",", ""))
            u = units[i]
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "%": 1}.get(u, 1)
            return v * mult
        t = val("gpu__time_duration.sum")
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        name = r[col("Kernel Name")].split("(")[0].replace("void ", "")[:30]
        grid = r[col("launch__grid_size")]
        print(f"{name:30s} {grid:>8s} {t*1e6:9.1f} {rd/1e6:12.1f} {wr/1e6:12.1f} {((rd+wr)/t)/1e9:10.0f} {100*((rd+wr)/t)/1e9/peak:12.1f}"
              f" {val('dram__cycles_active.avg.pct_of_peak_sustained_elapsed'):22.1f} {val('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):10.1f}"
              f" {val('smsp__issue_active.avg.pct_of_peak_sustained_active'):10.1f}")
