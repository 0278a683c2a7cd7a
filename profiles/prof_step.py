"""Driver for ncu: ONE replay of the captured step graph inside a cudaProfilerStart/Stop range (run ncu with
`--profile-from-start off`), after warm-up replays, L2 flushed in front of the profiled replay.
Usage: python profiles/prof_step.py [workload] [n_profiled_replays]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from pcgnn_b200.parallel import FusedAdam, GradAllReduce  # noqa: E402
from pcgnn_b200.runtime import GraphedTrainStep  # noqa: E402
from pcgnn_b200.synth import make_graph  # noqa: E402
from pcgnn_b200.testing import build_cuda_pcgnn  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
spec, batch, embed, desc = bench.WORKLOADS[wl]
data = make_graph(spec, seed=72)
params = bench.init_params(data.feat.shape[1], embed, 3, 72)
model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device="cuda")
reducer = GradAllReduce(model.parameters()).attach()
opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
batches = bench.make_batches(data, 8 + reps, batch, 72)
eng = model.inter1.engine()
eng.set_features(model.inter1.features.weight)
cap = max(eng.slots_bound(b[0].astype(np.int32), [0.5] * 3, 0.5, True) for b in batches)
g = GraphedTrainStep(model, opt, batch, cap, reducer=reducer, warmup_batch=batches[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n, l in batches[:8]:
    g.run(n, l)
torch.cuda.synchronize()
for n, l in batches[8:]:
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss = g.run(n, l)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", float(loss), "overflow", g.overflowed())
