"""Small driver for ncu: runs only the hot-path kernels (score table + pool sort, choose, aggregate) on the
bench workload, a few iterations. Usage: python profiles/prof_kernels.py [workload] [iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from pcgnn_b200.engine import Engine  # noqa: E402
from pcgnn_b200.synth import make_graph  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec, batch, embed, desc = bench.WORKLOADS[wl]
data = make_graph(spec, seed=bench.SEED)
batches = bench.make_batches(data, iters, batch, bench.SEED)
eng = Engine(data.graph, "cuda")
eng.set_features(torch.from_numpy(data.feat).cuda())
eng.set_pool(sorted(data.train_pos))
rng = np.random.default_rng(0)
w = torch.from_numpy(rng.normal(size=(2, data.feat.shape[1])).astype(np.float32) * 0.3).cuda()
b = torch.zeros(2, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for nodes, labels in batches:
    t, host = eng.upload_targets(nodes.astype(np.int32))
    lab = torch.from_numpy(labels).cuda()
    cap = eng.slots_bound(host, [0.5] * 3, 0.5, True)
    eng.score_table(w, b)
    flush.zero_()
    sel = eng.choose(t, lab, True, [0.5] * 3, 0.5, cap)
    agg = eng.aggregate(sel, copy_dups=False)
    torch.cuda.synchronize()
print("ok", float(agg.sum()), int(sel.it_m[sel.it_rep.long()].sum()))
