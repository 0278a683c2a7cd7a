"""Timeline of ONE replay of the captured step graph (torch profiler / CUPTI): every kernel with its start
offset and duration, in start order, so that gaps (launch / dependency latency) and overlaps (forked branches)
are visible. `python profiles/anatomy.py [workload]` (default yelp = C2)."""
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
from torch.profiler import ProfilerActivity, profile  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
spec, batch, embed, desc = bench.WORKLOADS[wl]
data = make_graph(spec, seed=72)
params = bench.init_params(data.feat.shape[1], embed, 3, 72)
model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device="cuda")
reducer = GradAllReduce(model.parameters()).attach()
opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
batches = bench.make_batches(data, 6, batch, 72)
eng = model.inter1.engine()
eng.set_features(model.inter1.features.weight)
cap = max(eng.slots_bound(b[0].astype(np.int32), [0.5] * 3, 0.5, True) for b in batches)
g = GraphedTrainStep(model, opt, batch, cap, reducer=reducer, warmup_batch=batches[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rep in range(3):
    n, l = batches[1 + rep]
    g.nodes.copy_(torch.from_numpy(n.astype(np.int32)))
    g.labels.copy_(torch.from_numpy(l))
    flush.zero_()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.g_fb.replay()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    end = max(e.time_range.end for e in ev)
    print(f"--- replay {rep}: {len(ev)} kernels, wall {end - t0:.1f} us, sum {sum(e.device_time for e in ev):.1f} us")
    for e in ev:
        print(f"  +{e.time_range.start - t0:7.1f}  {e.device_time:6.1f} us  {e.name[:100]}")

# the recording that host batches replay (runtime.GraphedTrainStep.run): H2D copy of the batch and D2H copy of the loss
# are nodes of the graph
n, l = batches[1]
g.run(n, l)
torch.cuda.synchronize()
flush.zero_()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.g_host.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
end = max(e.time_range.end for e in ev)
print(f"--- host-batch recording: {len(ev)} nodes, wall {end - t0:.1f} us")
for e in ev:
    print(f"  +{e.time_range.start - t0:7.1f}  {e.device_time:6.1f} us  {e.name[:100]}")
