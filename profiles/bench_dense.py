"""Micro-benchmark of the fused dense kernels alone (CUDA events, warm). usage: bench_dense.py [B F E R]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pcgnn_b200.engine import Engine, padded_ld
from pcgnn_b200.graph import RelGraph, csr_from_edges

B, F_, E, R = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (1024, 32, 64, 3)
N = 50000
rng = np.random.default_rng(0)
ip, ix = csr_from_edges(N, rng.integers(0, N, 1000), rng.integers(0, N, 1000))
eng = Engine(RelGraph(N, [ip] * R, [ix] * R), "cuda")
feat = torch.rand(N, F_, device="cuda")
eng.set_features(feat)
ldf = padded_ld(F_)
targets = torch.randint(0, N, (B,), dtype=torch.int32, device="cuda")
agg = torch.rand(R * B, ldf, device="cuda")
w_intra = [torch.randn(2 * F_, E, device="cuda") * 0.1 for _ in range(R)]
w_inter = torch.randn(F_ + R * E, E, device="cuda") * 0.1
d_out = torch.randn(E, B, device="cuda")

def devtime(fn, n=10):
    """Per-kernel device time from the profiler (the eager loop below is bound by the host's launch rate)."""
    from torch.profiler import profile, ProfilerActivity
    for _ in range(3): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n): fn()
        torch.cuda.synchronize()
    import collections
    acc = collections.OrderedDict()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            acc[e.name[:60]] = acc.get(e.name[:60], 0.0) + e.device_time / n
    return acc


def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

out, cat = eng.dense_fwd(targets, agg, w_intra, w_inter, F_)
# reference
selff = feat[targets.long()]
hs = [torch.relu(torch.cat((selff, agg[r * B:(r + 1) * B, :F_]), 1) @ w_intra[r]) for r in range(R)]
ref = torch.relu(torch.cat([selff] + hs, 1) @ w_inter).t()
print("fwd max err", float((out - ref).abs().max()), "rel", float((out - ref).abs().max() / ref.abs().max()))
print("fwd  %.1f us" % timeit(lambda: eng.dense_fwd(targets, agg, w_intra, w_inter, F_)))
print("bwd  %.1f us" % timeit(lambda: eng.dense_bwd(agg, w_inter, cat, out, d_out, F_, R)))
def torch_fwd():
    selff = feat[targets.long()]
    hs = [torch.relu(torch.cat((selff, agg[r * B:(r + 1) * B, :F_]), 1) @ w_intra[r]) for r in range(R)]
    return torch.relu(torch.cat([selff] + hs, 1) @ w_inter).t()
print("torch fwd (eager, ~15 launches) %.1f us" % timeit(torch_fwd))

for name, fn in (("fwd", lambda: eng.dense_fwd(targets, agg, w_intra, w_inter, F_)),
                 ("bwd", lambda: eng.dense_bwd(agg, w_inter, cat, out, d_out, F_, R))):
    acc = devtime(fn)
    print(name, "device time per call: %.1f us" % sum(acc.values()))
    for k, v in acc.items():
        print("    %6.1f us  %s" % (v, k))
