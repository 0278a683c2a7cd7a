"""Per-rank timeline of the data-parallel step graph (run under torchrun): every kernel of one replay with start
offset and duration on every rank, and the spread of the ranks' arrival at the exchange kernel.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/anatomy_dp.py [workload] [strong]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from pcgnn_b200.parallel import FusedAdam, GradAllReduce, PeerComm  # noqa: E402
from pcgnn_b200.runtime import GraphedTrainStep  # noqa: E402
from pcgnn_b200.synth import make_graph  # noqa: E402
from pcgnn_b200.testing import build_cuda_pcgnn  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
strong = len(sys.argv) > 2 and sys.argv[2] == "strong"
spec, wl_batch, embed, desc = bench.WORKLOADS[wl]
gb = wl_batch if strong else wl_batch * world
batch = gb // world
data = make_graph(spec, seed=72)
params = bench.init_params(data.feat.shape[1], embed, 3, 72)
model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, device=dev)
reducer = GradAllReduce(model.parameters()).attach()
opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3, comm=PeerComm(reducer.flat.numel()))
weight = sum(data.graph.degrees(r).astype(np.int64) for r in range(3))
shards = [bench.deal(n, l, weight, rank, world) for n, l in bench.make_batches(data, 12, gb, 72)]
eng = model.inter1.engine()
eng.set_features(model.inter1.features.weight)
cap = max(eng.slots_bound(n.astype(np.int32), [0.5] * 3, 0.5, True) for n, _ in shards)
g = GraphedTrainStep(model, opt, batch, cap, reducer=reducer, world=world, warmup_batch=shards[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, l in shards[:6]:
    g.run(n, l)
torch.cuda.synchronize()
for rep in range(3):
    n, l = shards[6 + rep]
    g.nodes.copy_(torch.from_numpy(n.astype(np.int32)))
    g.labels.copy_(torch.from_numpy(l))
    flush.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.g_fb.replay()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    end = max(e.time_range.end for e in ev)
    lines = [f"rank {rank} replay {rep}: wall {end - t0:.1f} us"]
    for e in ev:
        lines.append(f"  +{e.time_range.start - t0:7.1f}  {e.device_time:6.1f} us  {e.name[:60]}")
    adam = [e for e in ev if "allreduce_adam" in e.name][0]
    info = torch.tensor([adam.time_range.start - t0, adam.device_time, end - t0], dtype=torch.float64, device=dev)
    allinfo = [torch.empty_like(info) for _ in range(world)]
    dist.all_gather(allinfo, info)
    for r in range(world):
        if r == rank and rep == 2:
            print("\n".join(lines), flush=True)
        dist.barrier()
    if rank == 0:
        a = torch.stack(allinfo).cpu().numpy()
        print(f"replay {rep}: exchange kernel starts at +{a[:, 0].min():.1f} .. +{a[:, 0].max():.1f} us on the ranks, lasts "
              f"{a[:, 1].min():.1f} .. {a[:, 1].max():.1f} us, step wall {a[:, 2].min():.1f} .. {a[:, 2].max():.1f} us", flush=True)
dist.destroy_process_group()
