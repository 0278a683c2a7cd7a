"""Per-source-line stall-sample summary from an .ncu-rep (needs -lineinfo builds).
usage: python profiles/ncu_lines.py <report.ncu-rep> <kernel-name-substring> [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
hdr = None
agg = collections.defaultdict(lambda: [0, 0])
fname = "?"
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        i_smp, i_inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and r[0].isdigit() and len(r) > i_inst:
        try:
            agg[(fname, int(r[0]), r[1].strip()[:110])][0] += int(r[i_smp])
            agg[(fname, int(r[0]), r[1].strip()[:110])][1] += int(r[i_inst])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
print(f"{kern}: {tot} samples (all launches in the report)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / tot:5.1f}%  smp={v[0]:6d} inst={v[1]:9d}  {k[0]}:{k[1]}  {k[2]}")
