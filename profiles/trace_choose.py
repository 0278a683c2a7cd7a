"""Debug-build tracer: per-item phase cycle counts of the choose kernels (needs a library built with
-DPCG_TRACE: make -C pc-gnn_b200/csrc clean all EXTRA=-DPCG_TRACE). Prints where the time of the slowest
items goes. Usage: python profiles/trace_choose.py [workload]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from pcgnn_b200 import _lib  # noqa: E402
from pcgnn_b200.engine import Engine  # noqa: E402
from pcgnn_b200.synth import make_graph  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
spec, batch, embed, desc = bench.WORKLOADS[wl]
data = make_graph(spec, seed=bench.SEED)
batches = bench.make_batches(data, 3, batch, bench.SEED)
eng = Engine(data.graph, "cuda")
eng.set_features(torch.from_numpy(data.feat).cuda())
eng.set_pool(sorted(data.train_pos))
L = _lib.lib()
R = data.graph.n_rel
trace = torch.zeros(batch * R * 12, dtype=torch.int64, device="cuda")
L.pcg_debug_set_trace.argtypes = [ctypes.c_void_p]
assert L.pcg_debug_set_trace(trace.data_ptr()) == 0
rng = np.random.default_rng(0)
w = torch.from_numpy(rng.normal(size=(2, data.feat.shape[1])).astype(np.float32) * 0.3).cuda()
b = torch.zeros(2, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it, (nodes, labels) in enumerate(batches):
    t, host = eng.upload_targets(nodes.astype(np.int32))
    lab = torch.from_numpy(labels).cuda()
    cap = eng.slots_bound(host, [0.5] * R, 0.5, True)
    eng.score_table(w, b)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sel = eng.choose(t, lab, True, [0.5] * R, 0.5, cap)
    e1.record()
    torch.cuda.synchronize()
    tr = trace.cpu().numpy().reshape(-1, 12)
    ts = tr[:, :8].astype(np.float64)
    d, k, o = tr[:, 8], tr[:, 9], tr[:, 10]
    start = ts[:, 0].min()
    dur = ts[:, 7] - ts[:, 0]
    names = ["slot-alloc", "load-dist", "select", "compact", "pool-search", "pool-emit", "finish"]
    print(f"iter {it}: choose {e0.elapsed_time(e1) * 1e3:.1f} us; last item ends at {ts[:, 7].max() - start:.0f} cycles "
          f"(global clock domain per SM differs; treat as approximate)")
    for tier, mask in (("warp tier (d<=512)", d <= 512), ("cta tier", d > 512)):
        if not mask.any():
            continue
        idx = np.nonzero(mask)[0]
        ph = np.diff(ts[idx], axis=1)
        print(f"  {tier}: {len(idx)} items, mean {dur[idx].mean():.0f} cyc, p99 {np.percentile(dur[idx], 99):.0f}, "
              f"max {dur[idx].max():.0f}; begin-time p50 {np.median(ts[idx, 0] - start):.0f} max {(ts[idx, 0] - start).max():.0f}")
        print("    mean per phase: " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, ph.mean(0))))
        worst = idx[np.argsort(-dur[idx])[:5]]
        for wi in worst:
            print(f"    slow item d={d[wi]} k={k[wi]} o={o[wi]} total={dur[wi]:.0f}: " +
                  ", ".join(f"{n}={v:.0f}" for n, v in zip(names, np.diff(ts[wi]))))
