"""CPU: the oracle (Python port + C restatement) against the golden vectors generated from the live
reference (tests/golden/make_golden.py). This is what pins the oracle."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, graph_of, load_golden, params_of, rel_err, split_sets
from oracle import c_oracle, port

TOL = 1e-5   # fp32 relative tolerance stated by BASELINE.json north_star


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_port_matches_golden(name, mode):
    g = load_golden(name)
    graph = graph_of(g)
    nodes = g["nodes"].tolist()
    B, R = len(nodes), graph.n_rel
    labels = g["labels"][g["nodes"]]
    pm = port.PortPCGNN(g["feat"], graph, g["train_pos"].tolist(), params_of(g), rho=float(g["rho"]),
                        alpha=float(g["alpha"]))
    pm.score_table = torch.from_numpy(g["score_table"])        # identical score bits
    if mode == "train":
        loss = pm.step_loss_backward(nodes, labels)
        assert abs(loss - float(g["train_loss"])) <= TOL * abs(float(g["train_loss"]))
    else:
        with torch.no_grad():
            pm.loss(nodes, labels, False)
    want = split_sets(g[f"{mode}_sel_ptr"], g[f"{mode}_sel_idx"], R, B)
    assert pm.last["sel"] == want                                # bit-exact id sets
    dptr, dval = g[f"{mode}_diff_ptr"], g[f"{mode}_diff_val"]
    got = [np.float32(x) for rel in pm.last["diffs"] for row in rel for x in row]
    assert np.array_equal(np.asarray(got, dtype=np.float32), dval)
    assert rel_err(pm.last["combined"].detach().numpy(), g[f"{mode}_combined"]) <= TOL
    assert rel_err(pm.last["logits"].detach().numpy(), g[f"{mode}_logits"]) <= TOL
    assert rel_err(pm.last["center"].detach().numpy(), g[f"{mode}_center"]) <= TOL


@pytest.mark.parametrize("name", GOLDEN)
def test_port_grads_match_golden(name):
    g = load_golden(name)
    graph = graph_of(g)
    nodes = g["nodes"].tolist()
    labels = g["labels"][g["nodes"]]
    pm = port.PortPCGNN(g["feat"], graph, g["train_pos"].tolist(), params_of(g), rho=float(g["rho"]),
                        alpha=float(g["alpha"]))
    pm.step_loss_backward(nodes, labels)        # own score table here: label_clf must get its gradient
    grads = pm.named_grads()
    for k, v in grads.items():
        assert rel_err(v, g["grad__" + k]) <= 1e-4, k


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_c_oracle_matches_golden(name, mode):
    g = load_golden(name)
    graph = graph_of(g)
    nodes = g["nodes"]
    B, R = len(nodes), graph.n_rel
    labels = g["labels"][nodes]
    sp, si = c_oracle.choose(graph, g["score_table"][:, 0], nodes, labels == 1, rho=float(g["rho"]),
                             pool=g["train_pos"], train=(mode == "train"))
    assert np.array_equal(sp, g[f"{mode}_sel_ptr"])
    assert np.array_equal(si, g[f"{mode}_sel_idx"])
    # aggregation of those sets, against the port's dense-mask mean
    agg = c_oracle.aggregate(g["feat"], sp, si)
    feat = torch.from_numpy(g["feat"])
    sets = split_sets(sp, si, R, B)
    for r in range(R):
        want = port._dense_mask_agg(feat, [set(s) for s in sets[r]], "mean").numpy()
        assert rel_err(agg[r * B:(r + 1) * B], want) <= TOL


def test_edge_case_fixture_covers_the_documented_cases():
    """d = 1, 2, 3 (keep all), 4 and 5 (first filtered sizes), a hub, ties, duplicate targets,
    oversampled ids already kept and not kept."""
    g = load_golden("edge_cases")
    graph = graph_of(g)
    nodes = g["nodes"].tolist()
    degs = {int(len(graph.row(r, v))) for r in range(3) for v in nodes}
    assert {1, 2, 3, 4, 5}.issubset(degs) and max(degs) >= 60
    assert len(nodes) != len(set(nodes))                                  # duplicate targets
    s = g["score_table"][:, 0]
    assert len(np.unique(s)) < len(s)                                     # ties exist
    B = len(nodes)
    tr = split_sets(g["train_sel_ptr"], g["train_sel_idx"], 3, B)
    ev = split_sets(g["eval_sel_ptr"], g["eval_sel_idx"], 3, B)
    labels = g["labels"][g["nodes"]]
    pool = set(g["train_pos"].tolist())
    added = kept_in_pool = 0
    for r in range(3):
        for i in range(B):
            if labels[i] == 1:
                added += len(set(tr[r][i]) - set(ev[r][i]))
                kept_in_pool += len(set(ev[r][i]) & pool)
    assert added > 0 and kept_in_pool > 0


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_homo_ports_match_golden(kind):
    """GCN / GraphSAGE restatements against the live reference's train step on the union graph."""
    from pcgnn_b200.graph import RelGraph

    g = load_golden("homo_" + kind)
    graph = RelGraph(int(g["n_nodes"]), [g["indptr"]], [g["indices"]])
    cls = port.PortGCN if kind == "gcn" else port.PortSAGE
    pm = cls(g["feat"], graph, g["enc_w"], g["head"])
    nodes = g["nodes"].tolist()
    loss = pm.loss(nodes, g["labels"][g["nodes"]])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    assert rel_err(pm.last["combined"].detach().numpy(), g["emb"]) <= TOL
    assert rel_err(pm.enc_w.grad.numpy(), g["grad_enc"]) <= 1e-4
    assert rel_err(pm.head.grad.numpy(), g["grad_head"]) <= 1e-4
