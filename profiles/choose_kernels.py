"""Per-kernel device times of score-table / choose / aggregate captured in CUDA graphs (torch profiler).
Usage: python profiles/choose_kernels.py [workload]"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from pcgnn_b200 import _lib  # noqa: E402
from pcgnn_b200.engine import Engine  # noqa: E402
from pcgnn_b200.synth import make_graph  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
spec, batch, embed, desc = bench.WORKLOADS[wl]
data = make_graph(spec, seed=bench.SEED)
batches = bench.make_batches(data, 3, batch, bench.SEED)
eng = Engine(data.graph, "cuda")
eng.set_features(torch.from_numpy(data.feat).cuda())
eng.set_pool(sorted(data.train_pos))
R = data.graph.n_rel
rng = np.random.default_rng(0)
w = torch.from_numpy(rng.normal(size=(2, data.feat.shape[1])).astype(np.float32) * 0.3).cuda()
b = torch.zeros(2, device="cuda")
nodes, labels = batches[0]
t, host = eng.upload_targets(nodes.astype(np.int32))
lab = torch.from_numpy(labels).cuda()
cap = eng.slots_bound(host, [0.5] * R, 0.5, True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        eng.score_table(w, b)
        sel = eng.choose(t, lab, True, [0.5] * R, 0.5, cap)
        eng.aggregate(sel, copy_dups=False)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with _lib.capture(g):
    eng.score_table(w, b)
    sel = eng.choose(t, lab, True, [0.5] * R, 0.5, cap)
    eng.aggregate(sel, copy_dups=False)
for rep in range(3):
    flush.zero_()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t0 = min(e.time_range.start for e in ev)
    print(f"replay {rep}: {len(ev)} kernels, span {(max(e.time_range.end for e in ev) - t0):.1f} us")
    for e in sorted(ev, key=lambda e: e.time_range.start):
        print(f"   +{e.time_range.start - t0:7.1f} us  {e.device_time:7.1f} us  {e.name[:80]}")
