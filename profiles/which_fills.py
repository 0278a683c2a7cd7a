"""Debug: list the kernels inside the forward+backward graph and inside the optimizer graph."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench
from pcgnn_b200.parallel import FusedAdam, GradAllReduce
from pcgnn_b200.runtime import GraphedTrainStep
from pcgnn_b200.synth import make_graph
from pcgnn_b200.testing import build_cuda_pcgnn
from torch.profiler import profile, ProfilerActivity

spec, batch, embed, desc = bench.WORKLOADS["yelp"]
data = make_graph(spec, seed=72)
params = bench.init_params(32, embed, 3, 72)
model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device="cuda")
reducer = GradAllReduce(model.parameters()).attach()
opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)          # as bench.py: exchange + Adam kernel inside the step graph
batches = bench.make_batches(data, 2, batch, 72)
eng = model.inter1.engine(); eng.set_features(model.inter1.features.weight)
cap = eng.slots_bound(batches[0][0].astype(np.int32), [0.5]*3, 0.5, True) * 2
g = GraphedTrainStep(model, opt, batch, cap, reducer=reducer, warmup_batch=batches[0])
for name, gr in (("step graph", g.g_fb),):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        gr.replay(); torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    print(name, len(ev), "kernels")
    tot = sum(e.device_time for e in ev)
    print("   total device time %.1f us" % tot)
    for e in sorted(ev, key=lambda e: -e.device_time)[:22]:
        print("   %7.1f us  %s" % (e.device_time, e.name[:90]))
