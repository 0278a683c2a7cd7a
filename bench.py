#!/usr/bin/env python
"""Benchmark of the PC-GNN pick-and-choose hot path on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload yelp|amazon|yelp100|amazon_gcn|big]
                  [--impl reference] [--torch-adam] [--nccl-scores] [--no-graph] [--no-cpu-baseline]

One "step" = one full training step of PCALayer(InterAgg3(IntraAgg x3)) on one label-balanced batch of
B target nodes per GPU: loss (score table, pool sort, choose, aggregate, relation transforms, combine, heads,
both cross-entropies) -> backward -> gradient mean over the ranks + Adam (one kernel over NVLink peer memory;
--torch-adam: NCCL all-reduce + torch.optim.Adam). Under torchrun every rank runs its shard; rank 0 prints ONE
JSON line.

  value      train target-nodes/s with the batch already resident in HBM (one CUDA-graph replay per step)
  e2e        the same with HOST inputs through runtime.GraphedTrainStep.run(list_of_ids, labels): H2D of the
             ids/labels and D2H of the loss inside the timed region (e2e_reference_api_eager: the reference's
             own call sequence model.loss(list, labels); backward; step; loss.item(), eager)
  roofline   the slower of the two hot-path kernel groups (choose / aggregate), algorithmic bytes per
             launch / CUDA-event time, against MEASURED_PEAKS.json; traffic = ncu DRAM bytes (profiles/)
  cpu_baseline  the oracle port (same per-target structure as the reference, CPU) on a 1024-target batch

Workloads (BASELINE.json configs): yelp = C2 (default, the config the metric is quoted on), amazon = C1,
yelp100 = C3, amazon_gcn = C4, big = C5 (row-partitioned CSR, 1.25M nodes / 1.26e8 entries per GPU).
`--impl reference` times the oracle port alone on the host cores (the reference is pure Python and is not
present on the GPU box; oracle/port.py is its restatement, pinned by tests/golden).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synthetic spec, batch per GPU, embed dim, description)
    "yelp": ("yelp", 1024, 64, "C2 YelpChi-shaped N=45954 F=32 R=3 ~4.0M edges, emb 64, batch 1024/GPU"),
    "amazon": ("amazon", 1024, 64, "C1 Amazon-shaped N=11944 F=25 R=3 ~4.8M edges, emb 64, batch 1024/GPU"),
    "yelp100": ("yelp100", 4096, 128, "C3 YelpChi-shaped F=100, emb 128, batch 4096/GPU"),
    "amazon_gcn": ("amazon", 1024, 64, "C4 GCN baseline on the Amazon-shaped union graph, emb 64, batch 1024/GPU"),
    "big": ("big", 1024, 64, "C5 power-law fraud graph, 1.25M nodes and ~1.25e8 CSR entries PER GPU (10M / 1e9 at 8 GPUs), "
                             "F=64, R=3, 10% positives, CSR row-partitioned by node range, emb 64, batch 1024/GPU"),
}
RHO, ALPHA, LR, WD = 0.5, 2.0, 0.01, 1e-3
SEED = 72


MY_KERNELS_PER_STEP = 16   # score+sort 2, choose 3 (prep, wide, small), aggregate 1, dense fwd 2, center/head fwd 2,
#                            head/center bwd 2, dense bwd 3, gradient exchange + Adam 1 (+2 when P > 8192, +1 big tier,
#                            +1 self copy when F > E)


def config_dict(desc, batch, world):
    return {"workload": desc, "global_batch": batch * world, "rho": RHO, "thresholds": 0.5, "optimizer": "Adam",
            "l2": "flushed between timed steps (256 MiB write)",
            "parallelism": f"dp{world} (targets sharded, grads all-reduced)" if world > 1 else "single"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_batches(data, n_batches, batch, seed):
    """Label-balanced batches: nodes drawn with pick_step's weights deg/LF (utils.py:274-278)."""
    from pcgnn_b200.utils import pick_weights

    w = pick_weights(data.idx_train, data.y_train, data.homo)
    rng = np.random.default_rng(seed)
    idx = np.asarray(data.idx_train)
    out = []
    for _ in range(n_batches):
        nodes = idx[rng.choice(len(idx), batch, p=w / w.sum())]
        out.append((nodes.astype(np.int64), data.labels[nodes].astype(np.int64)))
    return out


def init_params(feat_dim, embed, n_rel, seed):
    rng = np.random.default_rng(seed)

    def xavier(r, c):
        a = np.sqrt(6.0 / (r + c))
        return rng.uniform(-a, a, size=(r, c)).astype(np.float32)

    return dict(intra=[xavier(2 * feat_dim, embed) for _ in range(n_rel)], inter=xavier(feat_dim + n_rel * embed, embed),
                clf_w=xavier(2, feat_dim), clf_b=np.zeros(2, np.float32), head=xavier(2, embed),
                enc=xavier(embed, feat_dim))      # enc: GCNEncoder weight [E,F] (C4 only)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake": 0x80, "sync_boost": 0x10}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.02)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_port_rate(data, params, batches, sample, steps, warmup, gcn=False):
    """Oracle port (CPU restatement of the reference's structure): full train steps on `sample` targets."""
    import torch
    from oracle import port

    tp = sorted(data.train_pos)
    if gcn:
        pm = port.PortGCN(data.feat, data.homo, params["enc"], params["head"])
    else:
        pm = port.PortPCGNN(data.feat, data.graph, tp, params, rho=RHO, alpha=ALPHA)
    opt = torch.optim.Adam(pm.parameters(), lr=LR, weight_decay=WD)
    times = []
    for s in range(warmup + steps):
        nodes, labels = batches[s % len(batches)]
        t0 = time.perf_counter()
        opt.zero_grad()
        if gcn:
            loss = pm.loss(nodes[:sample].tolist(), labels[:sample])
        else:
            loss = pm.loss(nodes[:sample].tolist(), labels[:sample], True, shared_table=False)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return sample / float(np.mean(times)), float(np.mean(times))


def c_port_rate(data, batches, reps=3):
    """C restatement (OpenMP, all host cores) of choose + aggregate only, full batch."""
    from oracle import c_oracle

    c_oracle.lib()
    rng = np.random.default_rng(1)
    score = (data.feat @ rng.normal(size=data.feat.shape[1]).astype(np.float32) * 0.3).astype(np.float32)
    tp = sorted(data.train_pos)
    best = 1e9
    for i in range(reps):
        nodes, labels = batches[i % len(batches)]
        t0 = time.perf_counter()
        sp, si = c_oracle.choose(data.graph, score, nodes, labels == 1, rho=RHO, pool=tp, train=True)
        c_oracle.aggregate(data.feat, sp, si)
        best = min(best, time.perf_counter() - t0)
    return len(batches[0][0]) / best


GCN_KERNELS_PER_STEP = 2   # select-all + aggregate (the GCN encoder/head are torch library GEMMs, as in the reference)


def build_cuda_pcgnn_device(feat_dev, graph, train_pos_dev, params, dev):
    """PCALayer(InterAgg3(IntraAgg x3)) over tables that already live on the device (workload big)."""
    import torch
    import torch.nn as nn

    from pcgnn_b200.layers import InterAgg3, IntraAgg
    from pcgnn_b200.model import PCALayer

    F_, E = feat_dev.shape[1], params["inter"].shape[1]
    features = nn.Embedding(1, 1)
    features.weight = nn.Parameter(feat_dev, requires_grad=False)          # no copy of the 2.56 GB table
    features.num_embeddings, features.embedding_dim = feat_dev.shape
    intras = [IntraAgg(features, F_, E, train_pos_dev, RHO, cuda=True) for _ in range(graph.n_rel)]
    inter = InterAgg3(features, F_, E, train_pos_dev, graph, intras, cuda=True)
    model = PCALayer(2, inter, ALPHA)
    with torch.no_grad():
        for ia, w in zip(intras, params["intra"]):
            ia.weight.copy_(torch.from_numpy(np.asarray(w)))
        inter.weight.copy_(torch.from_numpy(params["inter"]))
        inter.label_clf.weight.copy_(torch.from_numpy(params["clf_w"]))
        inter.label_clf.bias.copy_(torch.from_numpy(params["clf_b"]))
        model.weight.copy_(torch.from_numpy(params["head"]))
    return model.to(dev)


def build_cuda_gcn(data, params, dev):
    """GCN(GCNEncoder(GCNAggregator)) of the product on the union graph (model_handler.py:99-101, 119-120)."""
    import torch
    import torch.nn as nn
    from pcgnn_b200 import graphsage as gs

    F_ = data.feat.shape[1]
    E = params["enc"].shape[0]
    features = nn.Embedding(*data.feat.shape)
    features.weight = nn.Parameter(torch.from_numpy(np.ascontiguousarray(data.feat)), requires_grad=False)
    features = features.to(dev)
    enc = gs.GCNEncoder(features, F_, E, data.homo, gs.GCNAggregator(features, cuda=True), cuda=True)
    model = gs.GCN(2, enc)
    with torch.no_grad():
        enc.weight.copy_(torch.from_numpy(params["enc"]))
        model.weight.copy_(torch.from_numpy(params["head"]))
    return model.to(dev)


def _ncu_traffic(workload, group):
    """DRAM bytes per launch of a kernel group from the committed ncu capture (profiles/r01_traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return t[workload][group]["bytes"]
    except Exception:
        return None


def _roof(kern, dom, extra=None, workload=None):
    peak, peak_src = measured_peaks()
    r = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
         "frac": kern[dom]["gbs"] / peak, "traffic": _ncu_traffic(workload, dom), "peak_source": peak_src,
         "algorithmic_bytes": kern[dom]["alg_bytes"]}
    r.update(extra or {})
    return r


def hot_kernels_pcgnn(eng, inter, data, shards, dev_nodes, dev_labels, cap, W, K, flush, dev, batch, workload=None):
    import torch

    R, F_ = data.graph.n_rel, data.feat.shape[1]
    t_choose = t_agg = t_score = 0.0
    alg_choose = alg_agg = 0.0
    P = eng.P
    st_nodes = dev_nodes[W].clone()
    st_labels = dev_labels[W].clone()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            eng.score_table(inter.label_clf.weight, inter.label_clf.bias)
            sel = eng.choose(st_nodes, st_labels, True, inter.thresholds, RHO, cap)
            eng.aggregate(sel, copy_dups=False)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g_score, g_choose, g_agg = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    exchange = eng.score_group is not None          # partitioned graph: the all-gather stays outside the graphs
    with torch.cuda.graph(g_score):
        if exchange:
            eng.score_local(inter.label_clf.weight, inter.label_clf.bias)
        else:
            eng.score_table(inter.label_clf.weight, inter.label_clf.bias)
    with torch.cuda.graph(g_choose):
        sel = eng.choose(st_nodes, st_labels, True, inter.thresholds, RHO, cap)
    with torch.cuda.graph(g_agg):
        eng.aggregate(sel, copy_dups=False)       # as in the train step: the dense kernels read through it_rep
    for s in range(K):
        i = W + s
        st_nodes.copy_(dev_nodes[i])
        st_labels.copy_(dev_labels[i])
        flush.zero_()
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        e0.record()
        g_score.replay()
        if exchange:
            eng.score_exchange()
            eng.resort_pool()
        e1.record()
        g_choose.replay()
        e2.record()
        g_agg.replay()
        e3.record()
        torch.cuda.synchronize()
        t_score += e0.elapsed_time(e1)
        t_choose += e1.elapsed_time(e2)
        t_agg += e2.elapsed_time(e3)
        nodes, labels = shards[i]
        n_pos = int((labels == 1).sum())
        sum_d = sum(int(data.graph.degrees(r)[nodes - data.graph.row_lo].sum()) for r in range(R))
        m_tot = int(sel.it_m[sel.it_rep.long()].sum().item())
        alg_choose += 8.0 * sum_d + R * batch * 16 + 4 * batch + 4.0 * P * n_pos * R + 4.0 * m_tot
        alg_agg += (4.0 * F_ + 4.0) * m_tot + 4.0 * F_ * R * batch
        assert not sel.overflowed()
    kern = {
        "choose": {"ms": t_choose / K, "alg_bytes": alg_choose / K, "gbs": alg_choose / t_choose / 1e6,
                   "launches_per_step": 3},
        "aggregate": {"ms": t_agg / K, "alg_bytes": alg_agg / K, "gbs": alg_agg / t_agg / 1e6,
                      "launches_per_step": 1},
        "score_table_and_pool_sort": {"ms": t_score / K, "launches_per_step": 2},
    }
    dom = "choose" if t_choose >= t_agg else "aggregate"
    return kern, _roof(kern, dom, {"filter_plus_aggregate_gbs": (alg_choose + alg_agg) / (t_choose + t_agg) / 1e6},
                       workload=workload)


def hot_kernels_gcn(eng, agg_mod, data, shards, dev_nodes, cap, W, K, flush, dev):
    """C4: select-all + aggregate (rsqrt norm, self union) alone; bytes = (4F + 4) per neighbour row + output."""
    import torch
    from pcgnn_b200 import _lib

    F_ = data.feat.shape[1]
    st_nodes = dev_nodes[W].clone()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            eng.aggregate(eng.select_all(st_nodes, True, cap, _lib.NORM_RSQRT))
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        sel = eng.select_all(st_nodes, True, cap, _lib.NORM_RSQRT)
        eng.aggregate(sel)
    t = alg = 0.0
    for s in range(K):
        i = W + s
        st_nodes.copy_(dev_nodes[i])
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
        nodes, _ = shards[i]
        sum_d = int(data.homo.degrees(0)[nodes].sum())
        alg += (4.0 * F_ + 4.0) * sum_d + 16.0 * len(nodes) + 4.0 * F_ * len(nodes)
        assert not sel.overflowed()
    kern = {"aggregate": {"ms": t / K, "alg_bytes": alg / K, "gbs": alg / t / 1e6, "launches_per_step": 2}}
    return kern, _roof(kern, "aggregate")


def run_reference(args):
    """--impl reference: the oracle port on the host cores, same config/metric/unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from pcgnn_b200.synth import make_graph

    spec, batch, embed, desc = WORKLOADS[args.workload]
    if args.workload == "big":
        print(json.dumps({"impl": "reference", "unavailable": "workload big: the graph exists only as a device-side CSR "
                          "(1e9 entries); the CPU arm is measured on the yelp / amazon workloads"}))
        return
    data = make_graph(spec, seed=SEED)
    params = init_params(data.feat.shape[1], embed, data.graph.n_rel, SEED)
    batches = make_batches(data, 4, batch, SEED)
    sample = min(args.cpu_sample, batch)
    rate, sec = cpu_port_rate(data, params, batches, sample, args.steps, args.warmup,
                              gcn=args.workload.endswith("_gcn"))
    line = {
        "impl": "reference", "metric": "train target-nodes/sec (fwd+bwd)", "value": rate, "unit": "target-nodes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(desc, batch, max(args.gpus, 1)),
        "cpu_baseline": {"value": rate, "unit": "target-nodes/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"full train step of oracle/port.py on the first {sample} targets of each "
                                   f"{batch}-target batch (Python loop per target like the reference; "
                                   f"host has {os.cpu_count()} cpus)"},
        "e2e": {"value": rate, "unit": "target-nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="yelp", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=1024, help="targets per CPU-baseline step (bounded sample of a batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not use CUDA graphs for the device-resident step")
    ap.add_argument("--torch-adam", action="store_true",
                    help="NCCL all-reduce + torch.optim.Adam instead of the fused peer-memory exchange + Adam kernel")
    ap.add_argument("--nccl-scores", action="store_true",
                    help="workload big: exchange the score slices with an NCCL all-gather instead of peer-memory stores")
    ap.add_argument("--nodes-per-gpu", type=int, default=1_250_000, help="workload big: rows of the CSR held per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from pcgnn_b200 import _lib
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce, PeerComm
    from pcgnn_b200.synth import make_graph
    from pcgnn_b200.testing import build_cuda_pcgnn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    spec, batch, embed, desc = WORKLOADS[args.workload]
    is_gcn = args.workload.endswith("_gcn")
    is_big = args.workload == "big"
    n_b = W + K
    if is_big:
        # C5: every rank generates and holds only its own row range of the CSR; features, scores and the pool
        # are global. Each rank draws its share of the global batch from its own nodes (targets live with
        # their rows), so the only exchanges per step are the score-slice all-gather and the gradient sum.
        from pcgnn_b200.synth_big import BigSpec, make_partition

        part = make_partition(BigSpec(nodes_per_rank=args.nodes_per_gpu, seed=SEED), rank, world, dev)
        data = part
        F_, R = part.feat.shape[1], part.graph.n_rel
        params = init_params(F_, embed, R, SEED)
        model = build_cuda_pcgnn_device(part.feat, part.graph, part.train_pos, params, dev)
        inter = model.inter1
        if world > 1:
            # self-check of the exchange: slice kernel + all-gather must equal the whole table computed locally
            e_ = inter.engine()
            e_.set_features(inter.features.weight)
            e_.score_table(inter.label_clf.weight, inter.label_clf.bias)
            whole = e_.score.clone()
            if args.nccl_scores:
                e_.score.zero_()
                e_.score_group = dist.group.WORLD          # slice kernel | NCCL all-gather (between two graphs)
            else:
                e_.enable_score_broadcast(dist.group.WORLD)  # slice kernel storing into every rank's table (one graph)
            e_.score_table(inter.label_clf.weight, inter.label_clf.bias)
            torch.cuda.synchronize()
            dist.barrier()
            assert torch.equal(whole, e_.score), "score exchange differs from the locally computed table"
            del whole
        drawn = part.sample_batches(n_b, batch, SEED + rank)
        shards = [(n.cpu().numpy().astype(np.int64), l.cpu().numpy()) for n, l in drawn]
        global_batches = None
        torch.cuda.synchronize()
        desc += f"; this run: {part.n_global} nodes, {int(part.graph.indptr[-1]) * world / 1e6:.0f}M CSR entries, pool {int(part.train_pos.shape[0])}"
    elif is_gcn:
        data = make_graph(spec, seed=SEED)
        F_, R = data.feat.shape[1], data.graph.n_rel
        params = init_params(F_, embed, R, SEED)
        model = build_cuda_gcn(data, params, dev)
        inter = model.enc.aggregator             # the module that owns the engine / slot capacity
    else:
        data = make_graph(spec, seed=SEED)
        F_, R = data.feat.shape[1], data.graph.n_rel
        params = init_params(F_, embed, R, SEED)
        model = build_cuda_pcgnn(data.feat, data.graph, sorted(data.train_pos), params, rho=RHO, alpha=ALPHA, device=dev)
        inter = model.inter1
    reducer = GradAllReduce(model.parameters()).attach()
    if args.torch_adam:     # NCCL all-reduce between two CUDA graphs + torch's fused Adam (the baseline arrangement)
        opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=LR, weight_decay=WD,
                               capturable=True, fused=True)
    else:                   # gradient exchange over NVLink peer memory + Adam in one kernel inside the step graph
        opt = FusedAdam(reducer, lr=LR, weight_decay=WD, comm=PeerComm(reducer.flat.numel()))
    fused = not args.torch_adam
    if not is_big:
        # global batches of batch*world targets, identical on every rank; this rank's contiguous shard
        global_batches = make_batches(data, n_b, batch * world, SEED)
        shards = [(n[rank * batch:(rank + 1) * batch], l[rank * batch:(rank + 1) * batch]) for n, l in global_batches]
    dev_nodes = [torch.from_numpy(n.astype(np.int32)).to(dev) for n, _ in shards]
    dev_labels = [torch.from_numpy(l).to(dev) for _, l in shards]
    host_nodes = [n.tolist() for n, _ in shards]
    host_labels = [l for _, l in shards]
    if is_gcn:
        eng = inter._get_engine()
        eng.set_features(inter.features.weight)
        cap = max(inter.slots_bound(n) for n, _ in shards)
    else:
        eng = inter.engine()
        eng.set_features(inter.features.weight)
        cap = max(eng.slots_bound(n.astype(np.int32), inter.thresholds, RHO, True) for n, _ in shards)
    inter.cap_slots_hint = cap
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def finish_step():
        if not fused:
            reducer()
            if world > 1:
                reducer.flat.div_(world)
        opt.step()                                               # fused: exchange + Adam + gradient clear

    def step_device_eager(i):
        if not fused:
            reducer.zero()
        loss = model.loss(dev_nodes[i], dev_labels[i])
        loss.backward()
        finish_step()
        return loss

    def step_host_eager(i):
        if not fused:
            reducer.zero()
        lab = torch.from_numpy(host_labels[i]).to(dev)          # model_handler.py:150 (cuda LongTensor of labels)
        loss = model.loss(host_nodes[i], lab)
        loss.backward()
        finish_step()
        return loss.item()                                       # D2H of the step's result

    use_graph = not args.no_graph
    if use_graph:
        from pcgnn_b200.runtime import GraphedTrainStep

        gstep = GraphedTrainStep(model, opt, batch, cap, reducer=reducer, world=world, warmup_batch=shards[0])

        def step_device(i):
            return gstep.run_device(dev_nodes[i], dev_labels[i])

        def step_host(i):
            return gstep.run(host_nodes[i], host_labels[i]).item()     # H2D ids+labels, replay, D2H loss
    else:
        step_device, step_host = step_device_eager, step_host_eager

    def timed(fn, first):
        evs = []
        for s in range(K):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(first + s)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    # ---- device-resident inputs (value) ----
    for s in range(W):
        step_device(s)
    barrier()
    sampler.start()
    ms_dev = timed(step_device, W)
    sampler.sample()
    barrier()
    ms_dev = max_over_ranks(ms_dev)
    # ---- host inputs through the public API (e2e) ----
    for s in range(W):
        step_host(s)
    barrier()
    ms_e2e = timed(step_host, W)
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    ms_api = None
    if use_graph:      # the reference-facing eager call model.loss(list, labels) for comparison
        if hasattr(inter, "scores_external"):
            inter.scores_external = False      # the eager call computes (and exchanges) the scores itself
        for s in range(W):
            step_host_eager(s)
        barrier()
        ms_api = max_over_ranks(timed(step_host_eager, W))
        barrier()
        assert not gstep.overflowed()
    sampler.stop_flag = True

    # ---- hot-path kernels alone (roofline), same batches. Each group is captured into its own CUDA graph
    # (static input buffers) so the events bracket GPU work only, not the host's launch calls. ----
    kern, roof = hot_kernels_gcn(eng, inter, data, shards, dev_nodes, cap, W, K, flush, dev) if is_gcn else \
        hot_kernels_pcgnn(eng, inter, data, shards, dev_nodes, dev_labels, cap, W, K, flush, dev, batch, args.workload)

    if rank == 0:
        total_nodes = batch * world * K
        line = {
            "metric": "train target-nodes/sec (fwd+bwd)", "value": total_nodes / (ms_dev / 1e3),
            "unit": "target-nodes/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(desc, batch, world),
            "e2e": {"value": total_nodes / (ms_e2e / 1e3), "unit": "target-nodes/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": batch * 4 + batch * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": (GCN_KERNELS_PER_STEP if is_gcn else MY_KERNELS_PER_STEP) * K,
            "mode": "cuda-graph replay (runtime.GraphedTrainStep)" if use_graph else "eager",
            "roofline": roof, "kernels": kern, "clocks": sampler.result(),
        }
        if ms_api is not None:
            line["e2e_reference_api_eager"] = {"value": total_nodes / (ms_api / 1e3), "unit": "target-nodes/s",
                                               "ms_per_step": ms_api / K,
                                               "call": "model.loss(list_of_ids, cuda_labels); backward; Adam.step; loss.item()"}
        if world == 1 and not args.no_cpu_baseline and not is_big:
            sample = min(args.cpu_sample, batch)
            rate, sec = cpu_port_rate(data, params, global_batches, sample, 12, 1, gcn=is_gcn)
            line["cpu_baseline"] = {
                "value": rate, "unit": "target-nodes/s", "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"12 full train steps of oracle/port.py on the first {sample} targets of a batch "
                          f"({sec * 1e3:.0f} ms each; host has {os.cpu_count()} cpus)",
                }
            if not is_gcn:
                line["cpu_baseline"]["c_port_choose_aggregate_nodes_per_s"] = c_port_rate(data, global_batches)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
