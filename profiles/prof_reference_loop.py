"""Host-side profile (cProfile) of the reference's own training loop on the package's modules: where the time of
`zero_grad; model.loss(list, labels); backward; Adam.step; loss.item()` goes once the GPU work is ~0.1 ms."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from pcgnn_b200.synth import make_graph  # noqa: E402
from pcgnn_b200.testing import build_cuda_pcgnn  # noqa: E402

spec, batch, embed, desc = bench.WORKLOADS["yelp"]
data = make_graph(spec, seed=72)
params = bench.init_params(32, embed, 3, 72)
model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device="cuda")
opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=0.01, weight_decay=1e-3)
batches = bench.make_batches(data, 60, batch, 72)
host = [(n.tolist(), l) for n, l in batches]


def step(i):
    opt.zero_grad()
    lab = torch.from_numpy(host[i][1]).cuda()
    loss = model.loss(host[i][0], lab)
    loss.backward()
    opt.step()
    return loss.item()


for i in range(10):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10, 60):
    step(i)
torch.cuda.synchronize()
print("wall per step: %.1f us; graph replays %d captures %d" % ((time.perf_counter() - t0) / 50 * 1e6,
      model.inter1.graphs().replays, model.inter1.graphs().captures))
pr = cProfile.Profile()
pr.enable()
for i in range(10, 60):
    step(i)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
