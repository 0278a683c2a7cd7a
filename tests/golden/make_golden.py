"""Generates tests/golden/*.npz from the LIVE reference (/root/reference, imported unmodified on CPU
and canonicalised by oracle/ref_harness.py). Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds the inputs (stacked CSR, features, labels, batch, train_pos, parameters, the shared
[N,2] score table) and the reference's outputs (kept-id sets per relation and target, distance lists,
combined embedding, center scores, logits, loss, parameter gradients) for train and eval mode.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import port, ref_harness as H  # noqa: E402
from pcgnn_b200.graph import RelGraph, csr_from_edges  # noqa: E402
from pcgnn_b200.synth import SynthSpec, make_graph  # noqa: E402


def flat_sets(sel):
    """list[R] of list[B] of sorted ids -> (ptr [R*B+1], idx)"""
    ptr, idx = [0], []
    for rel in sel:
        for ids in rel:
            idx.extend(ids)
            ptr.append(len(idx))
    return np.asarray(ptr, dtype=np.int64), np.asarray(idx, dtype=np.int32)


def flat_lists(diffs):
    ptr, val = [0], []
    for rel in diffs:
        for d in rel:
            val.extend(d)
            ptr.append(len(val))
    return np.asarray(ptr, dtype=np.int64), np.asarray(val, dtype=np.float32)


def edge_case_graph(seed):
    """One relation set with hand-made rows: degrees 1..6, a hub, duplicate scores, and train positives
    that are / are not neighbours of positive targets."""
    rng = np.random.default_rng(seed)
    n = 80
    rels = []
    for r in range(3):
        src, dst = [], []
        # node v (v < 12) gets exactly v%6 extra neighbours (so degrees 1..6 with the self loop)
        for v in range(12):
            for j in range((v + r) % 6):
                src.append(v)
                dst.append(20 + (v * 7 + j * 3 + r) % 50)
        # hub
        for u in range(15, 75):
            src.append(13 + r)
            dst.append(u)
        m = 60
        src += rng.integers(0, n, m).tolist()
        dst += rng.integers(0, n, m).tolist()
        rels.append(csr_from_edges(n, src, dst))
    g = RelGraph(n, [a for a, _ in rels], [b for _, b in rels])
    feat = rng.random((n, 7), dtype=np.float32)
    feat[30:40] = feat[40:50]          # duplicated feature rows => equal scores => ties
    feat[5] = feat[6]
    labels = (rng.random(n) < 0.3).astype(np.int64)
    labels[[0, 3, 4, 13, 14]] = 1
    train_pos = [int(v) for v in np.nonzero(labels)[0] if v % 3 != 2]
    return g, feat, labels, train_pos


def run_case(name, graph, feat, labels, train_pos, nodes, E, rho, seed):
    rng = np.random.default_rng(seed)
    F_, R = feat.shape[1], graph.n_rel
    params = dict(intra=[port.xavier(rng, 2 * F_, E) for _ in range(R)], inter=port.xavier(rng, F_ + R * E, E),
                  clf_w=port.xavier(rng, 2, F_), clf_b=rng.uniform(-.1, .1, 2).astype(np.float32),
                  head=port.xavier(rng, 2, E))
    ns = H.load(True)
    model = H.build_pcgnn(ns, feat, graph.to_adj_lists(), train_pos, E, rho, alpha=2.0, params=params)
    lab = labels[np.asarray(nodes)]
    out = {"indptr": graph.indptr, "indices": graph.indices, "n_nodes": graph.n_nodes, "n_rel": R,
           "feat": feat, "labels": labels, "train_pos": np.asarray(sorted(train_pos), dtype=np.int32),
           "nodes": np.asarray(nodes, dtype=np.int32), "E": E, "rho": rho, "alpha": 2.0,
           "intra": np.stack(params["intra"]), "inter": params["inter"], "clf_w": params["clf_w"],
           "clf_b": params["clf_b"], "head": params["head"]}
    for mode, tf in (("train", True), ("eval", False)):
        o = H.run_pcgnn(ns, model, nodes, lab, tf)
        out[f"{mode}_sel_ptr"], out[f"{mode}_sel_idx"] = flat_sets(o["sel"])
        out[f"{mode}_diff_ptr"], out[f"{mode}_diff_val"] = flat_lists(o["diffs"])
        out[f"{mode}_combined"] = o["combined"]
        out[f"{mode}_center"] = o["center"]
        out[f"{mode}_logits"] = o["logits"]
        out["score_table"] = o["score_table"]
        if tf:
            out["train_loss"] = np.float64(o["loss"])
            for k, v in o["grads"].items():
                out["grad__" + k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path), "bytes")


def run_homo_case(name, kind, data, nodes, E, seed):
    """GCN / GraphSAGE train step of the live reference on the union graph (model_handler.py:96-101, 118-120)."""
    import torch

    rng = np.random.default_rng(seed)
    F_ = data.feat.shape[1]
    enc_w = port.xavier(rng, E, F_)
    head = port.xavier(rng, 2, E)
    ns = H.load(False)
    G = ns.graphsage
    features = torch.nn.Embedding(*data.feat.shape)
    features.weight = torch.nn.Parameter(torch.from_numpy(data.feat), requires_grad=False)
    adj = data.homo.to_adj_lists()[0]
    if kind == "GCN":
        enc = G.GCNEncoder(features, F_, E, adj, G.GCNAggregator(features, cuda=False), cuda=False)
        model = G.GCN(2, enc)
    else:
        enc = G.Encoder(features, F_, E, adj, G.MeanAggregator(features, cuda=False), gcn=True, cuda=False)
        model = G.GraphSage(2, enc)
    with torch.no_grad():
        enc.weight.copy_(torch.from_numpy(enc_w))
        model.weight.copy_(torch.from_numpy(head))
    lab = torch.from_numpy(data.labels[np.asarray(nodes)])
    loss = model.loss([int(v) for v in nodes], lab)
    loss.backward()
    with torch.no_grad():
        emb = enc([int(v) for v in nodes])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), indptr=data.homo.indptr, indices=data.homo.indices,
                        n_nodes=data.homo.n_nodes, feat=data.feat, labels=data.labels,
                        nodes=np.asarray(nodes, dtype=np.int32), E=E, enc_w=enc_w, head=head,
                        loss=np.float64(loss.item()), emb=emb.numpy(), grad_enc=enc.weight.grad.numpy(),
                        grad_head=model.weight.grad.numpy())
    print(name, "loss", loss.item())


def main():
    assert H.available(), "needs /root/reference"
    # 1. hand-made edge cases: every node once + duplicates, as one batch
    g, feat, labels, tp = edge_case_graph(7)
    nodes = list(range(0, 20)) + [13, 13, 0, 3]
    run_case("edge_cases", g, feat, labels, tp, nodes, 8, 0.5, 1)
    # 2. small random graph with duplicated features (ties) and degree skew
    d = make_graph("tiny", seed=3, dup_feature_frac=0.3)
    rng = np.random.default_rng(11)
    nodes = rng.choice(d.idx_train, 96).tolist()
    run_case("tiny_dup", d.graph, d.feat, d.labels, d.train_pos, nodes, 16, 0.5, 2)
    # 3. amazon-shaped miniature (F=25 -> padded rows, unlabeled prefix, normalised features), rho 0.8
    d = make_graph("tiny_amz", seed=5)
    rng = np.random.default_rng(12)
    nodes = rng.choice(d.idx_train, 64).tolist()
    run_case("tiny_amz", d.graph, d.feat, d.labels, d.train_pos, nodes, 12, 0.8, 3)
    # 4. single relation (InterAgg1)
    d = make_graph(SynthSpec("one_rel", 300, 9, (2500,), 0.25, zipf=0.9), seed=9, dup_feature_frac=0.2)
    rng = np.random.default_rng(13)
    nodes = rng.choice(d.idx_train, 48).tolist()
    run_case("one_rel", d.graph, d.feat, d.labels, d.train_pos, nodes, 8, 0.5, 4)
    # 5./6. GCN and GraphSAGE baselines on the union graph
    d = make_graph("tiny_amz", seed=47)
    rng = np.random.default_rng(14)
    nodes = rng.choice(d.idx_train, 90).tolist()
    run_homo_case("homo_gcn", "GCN", d, nodes, 16, 5)
    run_homo_case("homo_sage", "SAGE", d, nodes, 16, 6)


if __name__ == "__main__":
    main()
