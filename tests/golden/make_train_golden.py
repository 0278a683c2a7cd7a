"""Generates tests/golden/train_run.json from the LIVE reference (/root/reference, imported unmodified on CPU and
canonicalised by oracle/ref_harness.py): full training runs of PCALayer(InterAgg3(IntraAgg x3)) with the
reference's own epoch loop (model_handler.py:128-156: pick_step -> shuffle -> mini-batches -> Adam), then the
reference's evaluation pass (utils.py:298-312: batched to_prob(train_flag=False)), over several seeds.
The GPU test (tests/test_gpu_training_run.py) repeats the same runs on the CUDA modules and must land inside the
reference's run-to-run spread. Run in the build container only (~3 min):

    python tests/golden/make_train_golden.py
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import port, ref_harness as H  # noqa: E402
from pcgnn_b200.synth import make_graph  # noqa: E402

CONFIG = dict(spec="train_sig", graph_seed=5, signal=0.25, homophily=0.3, embed=64, rho=0.5, alpha=2.0, lr=0.01,
              weight_decay=1e-3, epochs=30, batch_size=100, eval_batch=256, seeds=[1, 2, 3, 5, 72])


def init_params(seed, F_, E, R):
    rng = np.random.default_rng(1000 + seed)
    return dict(intra=[port.xavier(rng, 2 * F_, E) for _ in range(R)], inter=port.xavier(rng, F_ + R * E, E),
                clf_w=port.xavier(rng, 2, F_), clf_b=np.zeros(2, np.float32), head=port.xavier(rng, 2, E))


def metrics(labels, prob1, pred):
    """AUC of the positive-class probability and G-mean of the argmax predictions (utils.py:316-325, :452-461)."""
    from sklearn.metrics import confusion_matrix, roc_auc_score

    tn, fp, fn, tp = confusion_matrix(labels, pred, labels=[0, 1]).ravel()
    gmean = float(np.sqrt((tp / max(tp + fn, 1)) * (tn / max(tn + fp, 1))))
    return float(roc_auc_score(labels, prob1)), gmean


def train_and_eval(model, loss_fn, prob_fn, data, cfg, seed, pick_step, make_labels):
    """The reference's loop (model_handler.py:124-156), then its evaluation pass (utils.py:298-312)."""
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    random.seed(seed)
    np.random.seed(seed)
    losses = []
    for epoch in range(cfg["epochs"]):
        sampled = pick_step(data.idx_train, data.y_train, data.homo_adj, size=len(data.train_pos) * 2)
        random.shuffle(sampled)
        num_batches = int(len(sampled) / cfg["batch_size"]) + 1
        for b in range(num_batches):
            nodes = sampled[b * cfg["batch_size"]:min((b + 1) * cfg["batch_size"], len(sampled))]
            if len(nodes) < 2:        # (the reference itself fails on empty / single-target batches)
                continue
            lab = data.labels[np.array(nodes)]
            opt.zero_grad()
            loss = loss_fn(nodes, make_labels(lab))
            loss.backward()
            opt.step()
            losses.append(float(loss.item()))
    test_nodes, y = data.idx_rest, np.asarray(data.y_rest)
    prob, pred = [], []
    for s in range(0, len(test_nodes), cfg["eval_batch"]):
        bn = test_nodes[s:s + cfg["eval_batch"]]
        if len(bn) < 2:
            continue
        p = prob_fn(bn, y[s:s + cfg["eval_batch"]])
        prob.extend(p[:, 1].tolist())
        pred.extend(p.argmax(axis=1).tolist())
    auc, gmean = metrics(y[:len(prob)], np.asarray(prob), np.asarray(pred))
    return auc, gmean, losses


def main():
    cfg = CONFIG
    d = make_graph(cfg["spec"], seed=cfg["graph_seed"], signal=cfg["signal"], homophily=cfg["homophily"])
    d.homo_adj = d.homo.to_adj_lists()[0]
    adj = d.graph.to_adj_lists()
    ns = H.load(canonical=True)
    out = dict(config=cfg, auc=[], gmean=[], first_losses=[], final_loss=[], n_steps=None)
    for seed in cfg["seeds"]:
        t0 = time.time()
        params = init_params(seed, d.feat.shape[1], cfg["embed"], 3)
        model = H.build_pcgnn(ns, d.feat, adj, d.train_pos, cfg["embed"], cfg["rho"], cfg["alpha"], params,
                              shared_scores=False)

        def loss_fn(nodes, lab):
            return model.loss([int(v) for v in nodes], lab)

        def prob_fn(nodes, lab):
            with torch.no_grad():
                gnn_prob, _ = model.to_prob([int(v) for v in nodes], lab, train_flag=False)
            return gnn_prob.numpy()

        auc, gmean, losses = train_and_eval(model, loss_fn, prob_fn, d, cfg, seed, ns.utils.pick_step,
                                            lambda lab: torch.from_numpy(lab))
        out["auc"].append(auc)
        out["gmean"].append(gmean)
        out["first_losses"].append(losses[:3])
        out["final_loss"].append(float(np.mean(losses[-5:])))
        out["n_steps"] = len(losses)
        print(f"seed {seed}: AUC {auc:.4f} G-mean {gmean:.4f} loss {losses[0]:.4f} -> {np.mean(losses[-5:]):.4f} "
              f"({len(losses)} steps, {time.time() - t0:.0f} s)", flush=True)
    out["auc_mean"], out["auc_sd"] = float(np.mean(out["auc"])), float(np.std(out["auc"], ddof=1))
    out["gmean_mean"], out["gmean_sd"] = float(np.mean(out["gmean"])), float(np.std(out["gmean"], ddof=1))
    with open(os.path.join(HERE, "train_run.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("AUC %.4f +- %.4f, G-mean %.4f +- %.4f" % (out["auc_mean"], out["auc_sd"], out["gmean_mean"], out["gmean_sd"]))


if __name__ == "__main__":
    main()
