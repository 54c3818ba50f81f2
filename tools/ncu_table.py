"""Per-kernel roofline table from an ncu report: one line per distinct (kernel, grid) with duration, DRAM bytes and GB/s.
usage: ncu_table.py report.ncu-rep [hbm_peak_GBps]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6538.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {k: hdr.index(k) for k in hdr}


def val(r, k, default=0.0):
    if k not in col:
        return default
    try:
        v = float(r[col[k]].replace(",", ""))
    except ValueError:
        return default
    u = units[col[k]]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return v * scale


seen = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    key = (name, r[col["Grid Size"]])
    t = val(r, "gpu__time_duration.sum")
    by = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    d = seen.setdefault(key, {"n": 0, "t": 0.0, "by": 0.0, "regs": r[col["launch__registers_per_thread"]] if "launch__registers_per_thread" in col else "?",
                              "occ": 0.0, "sm": 0.0})
    d["n"] += 1
    d["t"] += t
    d["by"] += by
    d["occ"] += val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    d["sm"] += val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
print(f"# {rep}: per (kernel, grid): launches, avg duration, avg DRAM read+write per launch, DRAM GB/s, % of {peak:.0f} GB/s "
      "(measured copy peak), registers, achieved occupancy %, SM throughput %")
print(f"{'kernel':44s} {'grid':>16s} {'n':>3s} {'us':>8s} {'MB':>8s} {'GB/s':>7s} {'%hbm':>5s} {'regs':>4s} {'occ%':>5s} {'sm%':>5s}")
for (name, grid), d in sorted(seen.items(), key=lambda kv: -kv[1]["t"]):
    n = d["n"]
    t, by = d["t"] / n, d["by"] / n
    gbs = by / t / 1e3 if t > 0 else 0.0
    print(f"{name[:44]:44s} {grid:>16s} {n:3d} {t:8.1f} {by / 1e6:8.1f} {gbs:7.0f} {100 * gbs / peak:5.1f} {d['regs']:>4s} "
          f"{d['occ'] / n:5.1f} {d['sm'] / n:5.1f}")
