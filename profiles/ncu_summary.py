"""Readers for the two ncu artefacts of a round (run here, no GPU needed):
  python profiles/ncu_summary.py launches gpurun_out/launches.csv        -> share of every kernel in the captured run
  python profiles/ncu_summary.py full gpurun_out/prof.ncu-rep            -> per-launch table of the --set full capture
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, per, cnt = 0.0, collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])[:70]
        v = float(r[vi].replace(",", ""))
        per[name] = per.get(name, 0.0) + v
        cnt[name] += 1
        tot += v
    print(f"{len(rows) - 1} launches, {tot / 1e3:.1f} us of GPU time (cold-cache, serialised: compare shares)")
    print(f"{'share':>7} {'avg ns':>10} {'count':>6}  kernel")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
        print(f"{v / tot * 100:6.2f}% {v / cnt[k]:10.0f} {cnt[k]:6d}  {k}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_bytes.sum", "launch__grid_size", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
    idx = [hdr.index(w) for w in want if w in hdr]
    print(" | ".join(f"{hdr[i]} [{units[i]}]" for i in idx))
    for r in rows[2:]:
        print(" | ".join(re.sub(r"\(.*", "", r[i])[:40] for i in idx))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
