"""Summaries for profiles/ from ncu outputs brought back in gpurun_out/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rN_launches_summary.txt
  python tools/ncu_summary.py full gpurun_out/match.ncu-rep   > profiles/rN_match_kernel_ncu_full.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "dram__bytes.sum"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    mu = h.index("Metric Unit")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0, "second": 1e3}.get(r[mu], 1e-6)
        tot[r[kn]] += v * scale
        cnt[r[kn]] += 1
    s = sum(tot.values())
    print(f"{'kernel':90s} {'n':>5s} {'total_ms':>10s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k[:90]:90s} {cnt[k]:5d} {v:10.3f} {100 * v / s:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    units = rows[1]
    for r in rows[2:]:
        print("Kernel Name".ljust(86), r[h.index("Kernel Name")])
        for m in sorted(KEEP):
            if m in h:
                i = h.index(m)
                print(f"{m:70s} {units[i]:>15s} {r[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
