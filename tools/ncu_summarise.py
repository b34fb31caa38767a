"""Turn ncu outputs into the small text summaries committed under profiles/.

  python tools/ncu_summarise.py launches <launches.csv> "<header>"       # --metrics gpu__time_duration.sum --csv log
  python tools/ncu_summarise.py full <report.ncu-rep> "<header>" [raw.csv]  # --set full capture (reads it with ncu -i)
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

FULL_COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
             "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
             "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
             "launch__grid_size"]


def launches(path, header):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
        tot[r[ki]][0] += 1
        tot[r[ki]][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"# {header}")
    print("launches     avg_ns   share  kernel")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:8d} {t / n:10.0f}  {100 * t / total:5.1f}%  {k[:100]}")
    mine = sum(t for k, (n, t) in tot.items() if "lgcn::" in k)
    print(f"# lgcn kernels {100 * mine / total:.1f}% of all device time in the run")


def full(rep, header, raw_out=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    if raw_out:
        open(raw_out, "w").write(out)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(c) for c in FULL_COLS if c in hdr]
    ki = hdr.index("Kernel Name")
    print(f"# {header}")
    print(" | ".join(["Kernel Name"] + [hdr[c] for c in cols]))
    print(" | ".join([""] + [units[c] for c in cols]))
    for r in rows[2:]:
        print(" | ".join([r[ki][:60]] + [r[c] for c in cols]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
