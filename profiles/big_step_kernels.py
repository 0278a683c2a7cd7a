"""Per-kernel device times of one C5 (workload big) step graph (torch profiler, rank 0 prints).
Usage: python profiles/big_step_kernels.py [nodes_per_gpu]      (or under torchrun for several GPUs)"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
import torch.distributed as dist  # noqa: E402
from pcgnn_b200.parallel import FusedAdam, GradAllReduce, PeerComm  # noqa: E402
from pcgnn_b200.runtime import GraphedTrainStep  # noqa: E402
from pcgnn_b200.synth_big import BigSpec, make_partition  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
part = make_partition(BigSpec(nodes_per_rank=n, seed=bench.SEED), rank, world, dev)
params = bench.init_params(part.feat.shape[1], 64, 3, bench.SEED)
model = bench.build_cuda_pcgnn_device(part.feat, part.graph, part.train_pos, params, dev)
inter = model.inter1
reducer = GradAllReduce(model.parameters()).attach()
opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3, comm=PeerComm(reducer.flat.numel()))
if world > 1:
    inter.engine().set_features(inter.features.weight)
    inter.engine().enable_score_broadcast(dist.group.WORLD)
drawn = part.sample_batches(4, 1024, bench.SEED + rank)
shards = [(a.cpu().numpy().astype(np.int64), b.cpu().numpy()) for a, b in drawn]
eng = inter.engine()
eng.set_features(inter.features.weight)
cap = max(eng.slots_bound(a.astype(np.int32), inter.thresholds, 0.5, True) for a, _ in shards)
deg = np.diff(part.graph.indptr)
if rank == 0:
    print(f"nodes {n} x {world}, entries {int(part.graph.indptr[-1])}, pool {int(part.train_pos.shape[0])}, max row {deg.max()}, "
      f"rows > 16384: {(deg > 16384).sum()}")
g = GraphedTrainStep(model, opt, 1024, cap, reducer=reducer, warmup_batch=shards[0])
for i in range(3):
    g.run(*shards[i])
torch.cuda.synchronize()
if os.environ.get("PCG_NCU"):          # under `ncu --profile-from-start off`: one replay inside the profiler range
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    g.run(*shards[3])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ok")
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.run(*shards[3])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in ev)
if rank == 0:
    print(f"{len(ev)} kernels, span {max(e.time_range.end for e in ev) - t0:.1f} us")
    for e in sorted(ev, key=lambda e: e.time_range.start):
        print(f"   +{e.time_range.start - t0:7.1f} us  {e.device_time:7.1f} us  {e.name[:90]}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
