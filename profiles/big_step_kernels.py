"""Per-kernel device times of one C5 (workload big) step graph on ONE GPU (torch profiler).
Usage: python profiles/big_step_kernels.py [nodes_per_gpu]"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from pcgnn_b200.parallel import FusedAdam, GradAllReduce  # noqa: E402
from pcgnn_b200.runtime import GraphedTrainStep  # noqa: E402
from pcgnn_b200.synth_big import BigSpec, make_partition  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
dev = torch.device("cuda", 0)
part = make_partition(BigSpec(nodes_per_rank=n, seed=bench.SEED), 0, 1, dev)
params = bench.init_params(part.feat.shape[1], 64, 3, bench.SEED)
model = bench.build_cuda_pcgnn_device(part.feat, part.graph, part.train_pos, params, dev)
inter = model.inter1
reducer = GradAllReduce(model.parameters()).attach()
opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
drawn = part.sample_batches(4, 1024, bench.SEED)
shards = [(a.cpu().numpy().astype(np.int64), b.cpu().numpy()) for a, b in drawn]
eng = inter.engine()
eng.set_features(inter.features.weight)
cap = max(eng.slots_bound(a.astype(np.int32), inter.thresholds, 0.5, True) for a, _ in shards)
deg = np.diff(part.graph.indptr)
print(f"nodes {n}, entries {int(part.graph.indptr[-1])}, pool {int(part.train_pos.shape[0])}, max row {deg.max()}, "
      f"rows > 16384: {(deg > 16384).sum()}")
g = GraphedTrainStep(model, opt, 1024, cap, reducer=reducer, warmup_batch=shards[0])
for i in range(3):
    g.run(*shards[i])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.run(*shards[3])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in ev)
print(f"{len(ev)} kernels, span {max(e.time_range.end for e in ev) - t0:.1f} us")
for e in sorted(ev, key=lambda e: e.time_range.start):
    print(f"   +{e.time_range.start - t0:7.1f} us  {e.device_time:7.1f} us  {e.name[:90]}")
