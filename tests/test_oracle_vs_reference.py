"""CPU, build container only: the oracle port against the LIVE reference on fresh seeds (skipped
where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from helpers import random_params, rel_err
from oracle import port, ref_harness as H
from pcgnn_b200.synth import make_graph

pytestmark = pytest.mark.skipif(not H.available(), reason="reference tree not present")


@pytest.mark.parametrize("seed", [0, 1])
def test_port_equals_canonical_reference(seed):
    d = make_graph("tiny", seed=20 + seed, dup_feature_frac=0.25)
    rng = np.random.default_rng(seed)
    params = random_params(rng, d.feat.shape[1], 16, 3)
    ns = H.load(True)
    model = H.build_pcgnn(ns, d.feat, d.graph.to_adj_lists(), d.train_pos, 16, 0.5, params=params)
    nodes = rng.choice(d.idx_train, 80).tolist()
    labels = d.labels[nodes]
    for tf in (True, False):
        o = H.run_pcgnn(ns, model, nodes, labels, tf)
        pm = port.PortPCGNN(d.feat, d.graph, sorted(d.train_pos), params)
        if tf:
            pm.step_loss_backward(nodes, labels)
            for k, v in pm.named_grads().items():
                assert rel_err(v, o["grads"][k]) <= 1e-5, k
        else:
            with torch.no_grad():
                pm.loss(nodes, labels, False)
        assert pm.last["sel"] == o["sel"]
        assert pm.last["diffs"] == o["diffs"]
        assert rel_err(pm.last["combined"].detach().numpy(), o["combined"]) <= 1e-5


def test_unpatched_reference_equals_canonical_without_ties():
    """With all scores distinct the un-patched reference (set order, unstable sort) selects the same
    sets as the canonical one: the canonicalisation only decides ties."""
    d = make_graph("tiny", seed=31)          # no duplicated features
    rng = np.random.default_rng(5)
    params = random_params(rng, d.feat.shape[1], 8, 3)
    nodes = rng.choice(d.idx_train, 60).tolist()
    labels = d.labels[nodes]
    outs = []
    for canonical in (True, False):
        ns = H.load(canonical)
        tp = sorted(d.train_pos)
        model = H.build_pcgnn(ns, d.feat, d.graph.to_adj_lists(), tp, 8, 0.5, params=params)
        outs.append(H.run_pcgnn(ns, model, nodes, labels, True))
    assert len(np.unique(outs[0]["score_table"][:, 0])) == d.feat.shape[0]
    assert outs[0]["sel"] == outs[1]["sel"]
    assert rel_err(outs[1]["combined"], outs[0]["combined"]) <= 1e-5


def test_pick_step_replay_matches_reference():
    import random

    d = make_graph("tiny", seed=4)
    ns = H.load(False)
    homo = d.homo.to_adj_lists()[0]
    for seed in (1, 72):
        random.seed(seed)
        want = ns.utils.pick_step(d.idx_train, d.y_train, homo, size=150)
        random.seed(seed)
        u = [random.random() for _ in range(150)]
        lf = (d.y_train.sum() - len(d.y_train)) * d.y_train + len(d.y_train)
        w = np.array([len(homo[v]) for v in d.idx_train]) / lf
        assert port.pick_step_replay(d.idx_train, w, u) == want
        random.seed(seed)
        assert port.pick_step_port(d.idx_train, d.y_train, lambda v: len(homo[v]), 150) == want
