"""Static evidence (no GPU needed): per kernel of libpcgnn_b200.so, the SASS instruction count and the counts of
the mnemonics that show how it works: UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier operations (arrive /
expect_tx / try_wait), LDGSTS = cp.async, UCGABAR_* = thread-block-cluster barriers (distributed
shared memory exchange), ATOMS = shared-memory atomics (selection histograms), REDUX = warp reductions,
ATOMG/RED = global atomics (tickets, peer counters), SHFL/VOTE = warp scans / ballots, FFMA = fp32 math.
Usage: python profiles/sass_summary.py [path/to/libpcgnn_b200.so]"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pc-gnn_b200",
                                                        "libpcgnn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, cnt, size = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        size[cur] += 1
        cnt[cur][m.group(1).split(".")[0]] += 1
cols = ["UBLKCP", "SYNCS", "LDGSTS", "UCGABAR", "ATOMS", "REDUX", "ATOMG+RED", "SHFL", "VOTE", "FFMA", "BAR"]
print(f"{'kernel':62s} {'instr':>6s} " + " ".join(f"{k:>9s}" for k in cols))
for f in sorted(size, key=lambda k: -size[k]):
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    if "cub" in name or not f.startswith("_Z"):
        continue
    c = cnt[f]
    vals = [c["UBLKCP"], c["SYNCS"], c["LDGSTS"], c["UCGABAR_ARV"] + c["UCGABAR_WAIT"], c["ATOMS"], c["REDUX"], c["ATOMG"] + c["REDG"] + c["RED"],
            c["SHFL"], c["VOTE"], c["FFMA"], c["BAR"]]
    print(f"{name[:60]:62s} {size[f]:6d} " + " ".join(f"{v:9d}" for v in vals))
