import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from pcgnn_b200.synth import make_graph
from pcgnn_b200.testing import build_cuda_pcgnn
from pcgnn_b200.parallel import FusedAdam, GradAllReduce
from pcgnn_b200.runtime import GraphedTrainStep
spec, batch, embed, desc = bench.WORKLOADS["yelp"]
data = make_graph(spec, seed=72)
params = bench.init_params(32, embed, 3, 72)
batches = bench.make_batches(data, 40, batch, 72)
model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device="cuda")
host = [(n.tolist(), torch.from_numpy(l).cuda()) for n, l in batches]
for i in range(5):
    model.loss(*host[i]).item()
slot = next(iter(model.inter1.graphs().slots.values()))
def t(fn, n=30, sync=True):
    ts=[]
    for i in range(n):
        if sync: torch.cuda.synchronize()
        t0=time.perf_counter(); fn(i); ts.append(time.perf_counter()-t0)
    torch.cuda.synchronize()
    return np.median(ts)*1e6
print("cached graph replay host us (GPU idle):", t(lambda i: slot.graph.replay()))
print("cached graph replay host us (back to back):", t(lambda i: slot.graph.replay(), sync=False))
print("np.asarray(list):", t(lambda i: np.asarray(host[i][0], dtype=np.int32)))
cache = model.inter1.graphs()
eng = model.inter1.engine()
print("train_loss total host us:", t(lambda i: cache.train_loss(eng, host[i][0], host[i][1], model.weight, 2.0)))
print("model.loss host us:", t(lambda i: model.loss(*host[i])))
m2 = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device="cuda")
red = GraphedTrainStep  # noqa
reducer = GradAllReduce(m2.parameters()).attach()
opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
e2 = m2.inter1.engine(); e2.set_features(m2.inter1.features.weight)
cap = max(e2.slots_bound(b[0].astype(np.int32), [0.5]*3, 0.5, True) for b in batches)
g = GraphedTrainStep(m2, opt, batch, cap, reducer=reducer, warmup_batch=batches[0])
print("GraphedTrainStep graph replay host us (GPU idle):", t(lambda i: g.g_fb.replay()))
opt2 = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=0.01, weight_decay=1e-3)
def full(i):
    opt2.zero_grad(); l = model.loss(*host[i]); l.backward(); opt2.step(); return l.item()
for i in range(5): full(i)
def seg(i):
    out=[]
    t0=time.perf_counter(); opt2.zero_grad(); out.append(time.perf_counter()-t0)
    t0=time.perf_counter(); l = model.loss(*host[i]); out.append(time.perf_counter()-t0)
    t0=time.perf_counter(); l.backward(); out.append(time.perf_counter()-t0)
    t0=time.perf_counter(); opt2.step(); out.append(time.perf_counter()-t0)
    t0=time.perf_counter(); l.item(); out.append(time.perf_counter()-t0)
    return out
r=np.array([seg(i) for i in range(5,35)])
print("zero_grad / loss / backward / step / item  (median us):", np.round(np.median(r,axis=0)*1e6,1))

# the same loop after fastloop.enable() (fresh model and optimizer), then a cProfile of it
from pcgnn_b200 import fastloop
import cProfile, pstats, io
model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device="cuda")
opt2 = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=0.01, weight_decay=1e-3)
fastloop.enable()
for i in range(8): full(i)
r=np.array([seg(i) for i in range(8,38)])
print("fastloop: zero_grad / loss / backward / step / item  (median us):", np.round(np.median(r,axis=0)*1e6,1))
pr = cProfile.Profile()
pr.enable()
for rep in range(4):
    for i in range(5, 35): full(i)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print("\n".join(l[:170] for l in s.getvalue().splitlines()[:75]))
fastloop.disable()
