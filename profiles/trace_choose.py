"""Debug-build tracer: per-item phase timestamps (ns, %globaltimer) of the choose kernels (needs a library built
with -DPCG_TRACE: make -C pc-gnn_b200/csrc clean all EXTRA=-DPCG_TRACE). Prints where the time of the slowest
items goes and when each tier starts/ends relative to the first item. Usage: python profiles/trace_choose.py [workload]"""
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
S = 40          # PCG_TRACE_SLOTS
trace = torch.zeros(batch * R * S + 16, dtype=torch.int64, device="cuda")
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
    trace.zero_()
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sel = eng.choose(t, lab, True, [0.5] * R, 0.5, cap)
    e1.record()
    torch.cuda.synchronize()
    raw = trace.cpu().numpy()
    pt = raw[batch * R * S:batch * R * S + 5].astype(np.float64)
    print("  prep kernel phases (ns): first-table=%d sizes=%d scan+queues=%d tail=%d; first item starts %+d ns after prep ends"
          % (pt[1] - pt[0], pt[2] - pt[1], pt[3] - pt[2], pt[4] - pt[3], raw[:batch * R * S].reshape(-1, S)[:, 0][raw[:batch * R * S].reshape(-1, S)[:, 7] > 0].min() - pt[4]))
    tr = raw[:batch * R * S].reshape(-1, S)
    done = tr[:, 7] > 0
    tr = tr[done]
    ts = tr[:, :8].astype(np.float64)
    d, k, o = tr[:, 8], tr[:, 9], tr[:, 10]
    fine = tr[:, 11:].astype(np.float64)        # [:, 0] steps; [:, 1 + 4 * step + j] select sub-phases; [:, 17..] compaction / pool
    start = ts[:, 0].min()
    dur = ts[:, 7] - ts[:, 0]
    names = ["header", "load-dist", "select", "compact", "pool-search", "pool-emit", "finish"]
    print(f"iter {it}: choose {e0.elapsed_time(e1) * 1e3:.1f} us (eager launch, events); {done.sum()} representative items; "
          f"last item ends {(ts[:, 7].max() - start) / 1e3:.1f} us after the first item starts")
    for tier, mask in (("warp tier (d<=256)", d <= 256), ("cta tier (<=2048)", (d > 256) & (d <= 2048)),
                       ("wide tier (<=16384)", (d > 2048) & (d <= 16384)), ("huge / big tiers", d > 16384)):
        if not mask.any():
            continue
        idx = np.nonzero(mask)[0]
        tsi = ts[idx].copy()
        # phases 5 (pool-search) is only stamped for positive items: fall back to the previous stamp
        for c in range(1, 8):
            tsi[:, c] = np.where(tsi[:, c] > 0, tsi[:, c], tsi[:, c - 1])
        ph = np.diff(tsi, axis=1)
        print(f"  {tier}: {len(idx)} items, mean {dur[idx].mean():.0f} ns, p99 {np.percentile(dur[idx], 99):.0f}, "
              f"max {dur[idx].max():.0f}; first starts +{(ts[idx, 0].min() - start) / 1e3:.1f} us, "
              f"median start +{(np.median(ts[idx, 0]) - start) / 1e3:.1f} us, last start +{(ts[idx, 0].max() - start) / 1e3:.1f} us, "
              f"last end +{(ts[idx, 7].max() - start) / 1e3:.1f} us")
        print("    mean ns per phase: " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, ph.mean(0))))
        worst = idx[np.argsort(-dur[idx])[:4]]
        for wi in worst:
            row = ts[wi].copy()
            for c in range(1, 8):
                row[c] = row[c] if row[c] > 0 else row[c - 1]
            print(f"    slow item d={d[wi]} k={k[wi]} o={o[wi]} total={dur[wi]:.0f} ns (start +{(ts[wi, 0] - start) / 1e3:.1f} us): " +
                  ", ".join(f"{n}={v:.0f}" for n, v in zip(names, np.diff(row))))
            if tier.startswith("wide") and fine[wi, 0] > 0:
                # CTA 0 of the cluster: per selection step (local histogram | merge + cluster barrier | remote reads |
                # scan + broadcast), relative to the end of the load phase; then the compaction's and the pool's stamps
                t2 = ts[wi, 2]
                steps = int(fine[wi, 0])
                parts, prev = [], t2
                for st in range(steps):
                    seg = []
                    for j in range(4):
                        v = fine[wi, 1 + 4 * st + j]
                        seg.append(f"{(v - prev):.0f}" if v > 0 else "-")
                        prev = v if v > 0 else prev
                    parts.append("/".join(seg))
                c28, c29, c30, c31 = fine[wi, 17], fine[wi, 18], fine[wi, 19], fine[wi, 20]
                print(f"      select steps={steps} [hist/merge+clsync/dsmem/scan] " + "  ".join(parts) +
                      f"; after select {ts[wi, 3] - prev:.0f}; compact: ballots+sync {c28 - ts[wi, 3]:.0f}, scan+clsync {c29 - c28:.0f}, "
                      f"stores {ts[wi, 4] - c29:.0f}, kbits clsync {c30 - ts[wi, 4]:.0f}" +
                      (f", pool search {c31 - c30:.0f}, ties {ts[wi, 5] - c31:.0f}" if c31 > 0 else ""))
