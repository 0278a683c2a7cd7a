"""GPU: the drop-in modules (InterAgg*/IntraAgg/PCALayer, GCN/SAGE, choose_step_*, pick_step)
against the oracle port and the golden vectors from the live reference."""
import random

import numpy as np
import pytest
import torch

from helpers import (GOLDEN, build_cuda_pcgnn, graph_of, load_golden, params_of, random_params, rel_err,
                     split_sets)
from oracle import port

pytestmark = pytest.mark.gpu
TOL = 1e-5          # embeddings / logits (north_star: 1e-5 relative, fp32)
GTOL = 1e-4         # parameter gradients (sums over the batch of fp32 products)


@pytest.mark.parametrize("name", GOLDEN)
def test_model_matches_golden(name):
    g = load_golden(name)
    graph = graph_of(g)
    nodes = g["nodes"].tolist()
    B, R = len(nodes), graph.n_rel
    labels = g["labels"][g["nodes"]]
    model = build_cuda_pcgnn(g["feat"], graph, g["train_pos"].tolist(), params_of(g), rho=float(g["rho"]),
                             alpha=float(g["alpha"]))
    model.inter1.score_override = torch.from_numpy(g["score_table"][:, 0].copy()).cuda()
    lab = torch.from_numpy(labels).cuda()
    # train step
    loss = model.loss(nodes, lab)
    loss.backward()
    got = model.inter1.last_selection.lists()
    want = split_sets(g["train_sel_ptr"], g["train_sel_idx"], R, B)
    for r in range(R):
        for i in range(B):
            assert got[r * B + i].tolist() == want[r][i]
    assert abs(loss.item() - float(g["train_loss"])) <= TOL * abs(float(g["train_loss"]))
    for k, p in model.named_parameters():
        if p.grad is None or "label_clf" in k:       # label_clf: the golden run differentiates through
            continue                                   # the shared table; checked in test_label_clf_grad
        assert rel_err(p.grad.cpu().numpy(), g["grad__" + k]) <= GTOL, k
    # eval pass through to_prob (utils.test calls it with numpy labels, utils.py:302-305)
    with torch.no_grad():
        emb, center = model.inter1(nodes, labels, train_flag=False)
        prob = model.to_prob(nodes, labels, train_flag=False)[0]
    assert rel_err(emb.cpu().numpy(), g["eval_combined"]) <= TOL
    assert rel_err(center.cpu().numpy(), g["eval_center"]) <= TOL
    assert rel_err(prob.cpu().numpy(), 1 / (1 + np.exp(-g["eval_logits"].astype(np.float64)))) <= TOL


def test_label_clf_grad_and_own_score_table():
    """Without the injected table: scores come from pcg_score_table; selection may only differ from the
    oracle where distances are within an ulp, embeddings/grads still agree to tolerance on a
    tie-free graph."""
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny", seed=41)
    rng = np.random.default_rng(0)
    params = random_params(rng, d.feat.shape[1], 16, 3)
    tp = sorted(d.train_pos)
    nodes = rng.choice(d.idx_train, 100).tolist()
    labels = d.labels[nodes]
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    loss = model.loss(nodes, torch.from_numpy(labels).cuda())
    loss.backward()
    table = model.inter1.engine().score.cpu().numpy()
    pm = port.PortPCGNN(d.feat, d.graph, tp, params)
    ref = pm.step_loss_backward(nodes, labels)
    want_table = pm.last["score_table"].detach().numpy()[:, 0]
    assert rel_err(table, want_table) <= TOL
    assert abs(loss.item() - ref) <= 1e-4 * abs(ref)
    grads = pm.named_grads()
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert rel_err(p.grad.cpu().numpy(), grads[k]) <= GTOL, k       # (r1: 1e-3)


def test_state_dict_keys_match_reference():
    g = load_golden("tiny_dup")
    model = build_cuda_pcgnn(g["feat"], graph_of(g), g["train_pos"].tolist(), params_of(g))
    keys = set(model.state_dict().keys())
    want = {"weight", "inter1.weight", "inter1.features.weight", "inter1.label_clf.weight", "inter1.label_clf.bias"}
    for r in (1, 2, 3):
        want |= {f"inter1.intra_agg{r}.weight", f"inter1.intra_agg{r}.features.weight"}
    assert keys == want                                     # SURVEY.md §5 (checkpoint compatibility)


def test_adj_lists_dict_of_sets_input():
    """The reference hands InterAgg a list of dict[int -> set[int]] with numpy-integer members."""
    from pcgnn_b200.layers import InterAgg3, IntraAgg
    import torch.nn as nn

    g = load_golden("tiny_dup")
    graph = graph_of(g)
    adj = graph.to_adj_lists()
    adj = [{np.int64(k): {np.int64(x) for x in v} for k, v in a.items()} for a in adj]
    feat = g["feat"]
    features = nn.Embedding(*feat.shape)
    features.weight = nn.Parameter(torch.from_numpy(feat), requires_grad=False)
    tp = g["train_pos"].tolist()
    intras = [IntraAgg(features, feat.shape[1], 16, tp, 0.5, cuda=True) for _ in range(3)]
    inter = InterAgg3(features, feat.shape[1], 16, tp, adj, intras, cuda=True).to("cuda")
    inter.score_override = torch.from_numpy(g["score_table"][:, 0].copy()).cuda()
    nodes = g["nodes"].tolist()
    inter(nodes, torch.from_numpy(g["labels"][g["nodes"]]).cuda(), True)
    got = inter.last_selection.lists()
    want = split_sets(g["train_sel_ptr"], g["train_sel_idx"], 3, len(nodes))
    assert all(got[r * len(nodes) + i].tolist() == want[r][i] for r in range(3) for i in range(len(nodes)))


def test_trainable_features_scatter_backward():
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny", seed=43)
    rng = np.random.default_rng(1)
    params = random_params(rng, d.feat.shape[1], 8, 3)
    tp = sorted(d.train_pos)
    nodes = rng.choice(d.idx_train, 50).tolist()
    labels = d.labels[nodes]
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params, trainable_features=True)
    pm = port.PortPCGNN(d.feat, d.graph, tp, params)
    pm.feat.requires_grad_(True)
    ref = pm.loss(nodes, labels, True)
    ref.backward()
    model.inter1.score_override = pm.last["score_table"].detach()[:, 0].contiguous().cuda()
    loss = model.loss(nodes, torch.from_numpy(labels).cuda())
    loss.backward()
    got = model.inter1.features.weight.grad.cpu().numpy()
    # the port differentiates the label scores w.r.t. the features through its table; the product's
    # injected table is a constant, so compare the embedding path only: rebuild the port gradient
    # with a detached table
    pm2 = port.PortPCGNN(d.feat, d.graph, tp, params)
    pm2.feat.requires_grad_(True)
    pm2.score_table = pm.last["score_table"].detach()
    lab = torch.from_numpy(labels)
    logits, center, _ = pm2.forward(nodes, lab, True)
    torch.nn.functional.cross_entropy(logits, lab).backward()
    want = pm2.feat.grad.numpy()
    g_model = build_cuda_pcgnn(d.feat, d.graph, tp, params, trainable_features=True)
    g_model.inter1.score_override = model.inter1.score_override
    lg, _ = g_model.forward(nodes, lab.cuda(), True)
    torch.nn.functional.cross_entropy(lg, lab.cuda()).backward()
    got = g_model.inter1.features.weight.grad.cpu().numpy()
    assert rel_err(got, want) <= GTOL


@pytest.mark.parametrize("train", [True, False])
def test_intra_agg_explicit_api(train):
    """IntraAgg.forward with the reference's explicit lists/scores signature (layers.py:562)."""
    from pcgnn_b200.layers import IntraAgg
    import torch.nn as nn

    g = load_golden("tiny_dup")
    graph = graph_of(g)
    nodes = g["nodes"].tolist()
    labels = g["labels"][g["nodes"]]
    table = torch.from_numpy(g["score_table"])
    feat = torch.from_numpy(g["feat"])
    tp = g["train_pos"].tolist()
    r = 1
    neighs = [graph.row(r, v).tolist() for v in nodes]
    rng = random.Random(0)
    shuffled = [rng.sample(x, len(x)) for x in neighs]               # list order must not matter
    nscores = [table[torch.tensor(x)] for x in shuffled]
    center = table[torch.tensor(nodes)]
    k_list = [int(np.ceil(len(x) * 0.5)) for x in neighs]
    w = torch.from_numpy(g["intra"][r])
    want, picked, diffs = port.intra_forward(feat, w, nodes, labels, neighs, center,
                                             [table[torch.tensor(x)] for x in neighs], table[torch.tensor(tp)],
                                             tp, k_list, float(g["rho"]), train)
    features = nn.Embedding(*g["feat"].shape)
    features.weight = nn.Parameter(feat.clone(), requires_grad=False)
    ia = IntraAgg(features, feat.shape[1], w.shape[1], tp, float(g["rho"]), cuda=True).to("cuda")
    with torch.no_grad():
        ia.weight.copy_(w)
    got, scores = ia.forward(nodes, torch.from_numpy(labels).cuda(), shuffled, center.cuda(),
                             [s.cuda() for s in nscores], table[torch.tensor(tp)].cuda(), k_list, train)
    assert rel_err(got.detach().cpu().numpy(), want.numpy()) <= TOL
    assert [list(map(np.float32, row)) for row in scores] == [list(map(np.float32, row)) for row in diffs]


def test_choose_step_functions_match_port():
    from pcgnn_b200.layers import choose_step_neighs, choose_step_test

    g = load_golden("edge_cases")
    graph = graph_of(g)
    nodes = g["nodes"].tolist()
    labels = g["labels"][g["nodes"]]
    table = torch.from_numpy(g["score_table"])
    tp = g["train_pos"].tolist()
    for r in range(3):
        neighs = [graph.row(r, v).tolist() for v in nodes]
        nscores = [table[torch.tensor(x)] for x in neighs]
        center = table[torch.tensor(nodes)]
        k_list = [int(np.ceil(len(x) * 0.5)) for x in neighs]
        want_s, want_d = port.choose_train(center, labels, nscores, neighs, table[torch.tensor(tp)], tp, k_list, 0.5)
        got_s, got_d = choose_step_neighs(center.cuda(), torch.from_numpy(labels).cuda(), [s.cuda() for s in nscores],
                                          neighs, table[torch.tensor(tp)].cuda(), tp, k_list, 0.5)
        assert got_s == want_s
        assert [list(map(np.float32, x)) for x in got_d] == [list(map(np.float32, x)) for x in want_d]
        want_s, want_d = port.choose_test(center, nscores, neighs, k_list)
        got_s, got_d = choose_step_test(center.cuda(), [s.cuda() for s in nscores], neighs, k_list)
        assert got_s == want_s
        assert [list(map(np.float32, x)) for x in got_d] == [list(map(np.float32, x)) for x in want_d]


@pytest.mark.parametrize("kind", ["GCN", "SAGE"])
def test_homo_baselines_match_port(kind):
    """GCN / GraphSAGE train step on the union graph (model_handler.py:96-101,118-120)."""
    import torch.nn as nn
    from pcgnn_b200 import graphsage as gs
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny_amz", seed=47)
    rng = np.random.default_rng(2)
    F_, E = d.feat.shape[1], 16
    enc_w = port.xavier(rng, E, F_)
    head = port.xavier(rng, 2, E)
    nodes = rng.choice(d.idx_train, 90).tolist()
    labels = d.labels[nodes]
    features = nn.Embedding(*d.feat.shape)
    features.weight = nn.Parameter(torch.from_numpy(d.feat), requires_grad=False)
    features = features.cuda()
    if kind == "GCN":
        agg = gs.GCNAggregator(features, cuda=True)
        enc = gs.GCNEncoder(features, F_, E, d.homo, agg, cuda=True)
        model = gs.GCN(2, enc)
        pm = port.PortGCN(d.feat, d.homo, enc_w, head)
    else:
        agg = gs.MeanAggregator(features, cuda=True)
        enc = gs.Encoder(features, F_, E, d.homo, agg, gcn=True, cuda=True)
        model = gs.GraphSage(2, enc)
        pm = port.PortSAGE(d.feat, d.homo, enc_w, head)
    with torch.no_grad():
        enc.weight.copy_(torch.from_numpy(enc_w))
        model.weight.copy_(torch.from_numpy(head))
    model = model.cuda()
    loss = model.loss(nodes, torch.from_numpy(labels).cuda())
    loss.backward()
    ref = pm.loss(nodes, labels)
    ref.backward()
    assert abs(loss.item() - float(ref)) <= TOL * abs(float(ref))
    assert rel_err(model.weight.grad.cpu().numpy(), pm.head.grad.numpy()) <= GTOL
    assert rel_err(enc.weight.grad.cpu().numpy(), pm.enc_w.grad.numpy()) <= GTOL
    with torch.no_grad():
        emb = enc(nodes)
    assert rel_err(emb.cpu().numpy(), pm.last["combined"].detach().numpy()) <= TOL


def test_gcn_adds_self_when_row_lacks_it():
    import torch.nn as nn
    from pcgnn_b200 import graphsage as gs
    from pcgnn_b200.graph import RelGraph, csr_from_edges

    rng = np.random.default_rng(3)
    n = 50
    ip, ix = csr_from_edges(n, rng.integers(0, n, 200), rng.integers(0, n, 200), self_loops=False)
    graph = RelGraph(n, [ip], [ix])
    feat = rng.random((n, 6), dtype=np.float32)
    features = nn.Embedding(n, 6)
    features.weight = nn.Parameter(torch.from_numpy(feat), requires_grad=False)
    features = features.cuda()
    agg = gs.GCNAggregator(features, cuda=True)
    agg.bind_graph(graph)
    nodes = list(range(n))
    got = agg.forward(nodes, gs._Rows(n)).cpu().numpy()
    for v in nodes:
        ids = sorted(set(graph.row(0, v).tolist()) | {v})
        want = feat[ids].sum(0) / np.sqrt(len(ids))
        assert np.allclose(got[v], want, rtol=1e-5, atol=1e-6)
    # stand-alone call with explicit sets (graphsage.py:200 signature)
    agg2 = gs.GCNAggregator(features, cuda=True)
    sets = [set(graph.row(0, v).tolist()) for v in nodes[:10]]
    got2 = agg2.forward(nodes[:10], sets).cpu().numpy()
    assert np.allclose(got2, got[:10], rtol=1e-5, atol=1e-6)


def test_pick_step_replays_random_choices():
    from pcgnn_b200.synth import make_graph
    from pcgnn_b200.utils import pick_step, pick_step_device, pick_weights

    d = make_graph("tiny", seed=4)
    homo = d.homo.to_adj_lists()[0]
    for seed in (1, 72):
        random.seed(seed)
        want = port.pick_step_port(d.idx_train, d.y_train, lambda v: len(homo[v]), 333)
        state_after = random.getstate()
        random.seed(seed)
        got = pick_step(d.idx_train, d.y_train, homo, 333)
        assert got == want
        assert random.getstate() == state_after            # same consumption of the global stream
        random.seed(seed)
        assert pick_step(d.idx_train, d.y_train, d.homo, 333) == want      # CSR graph input
    # Philox variant: same distribution (positives ~ half of the draws), reproducible per seed
    a = pick_step_device(d.idx_train, d.y_train, d.homo, 20000, seed=5).cpu().numpy()
    b = pick_step_device(d.idx_train, d.y_train, d.homo, 20000, seed=5).cpu().numpy()
    assert np.array_equal(a, b)
    w = pick_weights(d.idx_train, d.y_train, d.homo)
    pos_mass = w[d.y_train == 1].sum() / w.sum()
    assert abs(d.labels[a].mean() - pos_mass) < 0.02


def test_graphed_train_step_equals_eager_steps():
    """runtime.GraphedTrainStep (two CUDA graphs, static buffers) follows the same parameter trajectory as
    eager model.loss / backward / Adam.step over several batches, including duplicated targets."""
    from pcgnn_b200.parallel import GradAllReduce
    from pcgnn_b200.runtime import GraphedTrainStep
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny_amz", seed=51, dup_feature_frac=0.1)
    rng = np.random.default_rng(4)
    params = random_params(rng, d.feat.shape[1], 16, 3)
    tp = sorted(d.train_pos)
    B = 128
    batches = []
    for _ in range(5):
        nodes = rng.choice(d.idx_train, B)           # with replacement: duplicates inside a batch
        batches.append((nodes, d.labels[nodes]))

    def make():
        m = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        o = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=0.01, weight_decay=1e-3,
                             capturable=True, fused=True)
        return m, o

    eager, opt_e = make()
    losses_e = []
    for nodes, labels in batches:
        opt_e.zero_grad()
        loss = eager.loss(nodes.tolist(), torch.from_numpy(labels).cuda())
        loss.backward()
        opt_e.step()
        losses_e.append(loss.item())

    graphed, opt_g = make()
    red = GradAllReduce(graphed.parameters()).attach()
    eng = graphed.inter1.engine()
    eng.set_features(graphed.inter1.features.weight)
    cap = GraphedTrainStep.plan(eng, [n for n, _ in batches], graphed.inter1.thresholds, 0.5)
    step = GraphedTrainStep(graphed, opt_g, B, cap, reducer=red, warmup_batch=batches[0])
    losses_g = [step.run(nodes, labels).item() for nodes, labels in batches]
    assert not step.overflowed()
    assert np.allclose(losses_g, losses_e, rtol=1e-5, atol=1e-6), (losses_g, losses_e)
    for (k, a), (_, b) in zip(eager.named_parameters(), graphed.named_parameters()):
        assert rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) <= 1e-4, k


def test_duplicate_targets_share_their_representative():
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny", seed=53)
    rng = np.random.default_rng(5)
    params = random_params(rng, d.feat.shape[1], 8, 3)
    tp = sorted(d.train_pos)
    base = rng.choice(d.idx_train, 40, replace=False)
    nodes = np.concatenate([base, base[:15], base[5:10]]).tolist()     # every duplicate pattern
    labels = d.labels[nodes]
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    pm = port.PortPCGNN(d.feat, d.graph, tp, params)
    ref = pm.step_loss_backward(nodes, labels)
    model.inter1.score_override = pm.last["score_table"].detach()[:, 0].contiguous().cuda()
    loss = model.loss(nodes, torch.from_numpy(labels).cuda())
    loss.backward()
    sel = model.inter1.last_selection
    rep = sel.it_rep.cpu().numpy()
    B = len(nodes)
    assert (rep != np.arange(3 * B)).sum() == 3 * 20                    # 20 duplicates per relation
    got = sel.lists()
    for r in range(3):
        for i in range(B):
            assert got[r * B + i].tolist() == pm.last["sel"][r][i]
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
    grads = pm.named_grads()
    for k, p in model.named_parameters():
        if p.grad is not None and "label_clf" not in k:
            assert rel_err(p.grad.cpu().numpy(), grads[k]) <= GTOL, k


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_homo_baselines_match_golden(kind):
    """GCN / GraphSAGE modules against the live reference's golden train step."""
    import torch.nn as nn
    from pcgnn_b200 import graphsage as gs
    from pcgnn_b200.graph import RelGraph

    g = load_golden("homo_" + kind)
    graph = RelGraph(int(g["n_nodes"]), [g["indptr"]], [g["indices"]])
    F_, E = g["feat"].shape[1], int(g["E"])
    features = nn.Embedding(*g["feat"].shape)
    features.weight = nn.Parameter(torch.from_numpy(g["feat"]), requires_grad=False)
    features = features.cuda()
    if kind == "gcn":
        enc = gs.GCNEncoder(features, F_, E, graph, gs.GCNAggregator(features, cuda=True), cuda=True)
        model = gs.GCN(2, enc)
    else:
        enc = gs.Encoder(features, F_, E, graph, gs.MeanAggregator(features, cuda=True), gcn=True, cuda=True)
        model = gs.GraphSage(2, enc)
    with torch.no_grad():
        enc.weight.copy_(torch.from_numpy(g["enc_w"]))
        model.weight.copy_(torch.from_numpy(g["head"]))
    model = model.cuda()
    nodes = g["nodes"].tolist()
    loss = model.loss(nodes, torch.from_numpy(g["labels"][g["nodes"]]).cuda())
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    with torch.no_grad():
        assert rel_err(enc(nodes).cpu().numpy(), g["emb"]) <= TOL
    assert rel_err(enc.weight.grad.cpu().numpy(), g["grad_enc"]) <= GTOL
    assert rel_err(model.weight.grad.cpu().numpy(), g["grad_head"]) <= GTOL


def test_fused_adam_kernel_matches_torch_adam_on_the_same_gradients():
    """pcg_allreduce_adam (world 1) against torch.optim.Adam fed with identical gradients, 12 steps."""
    import torch.nn as nn
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce

    torch.manual_seed(0)
    shapes = [(217, 64), (50, 64), (2, 25), (2,), (2, 64), (7,)]          # 17,307 floats: not a multiple of 4
    ref = [nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    mine = [nn.Parameter(p.detach().clone()) for p in ref]
    opt_ref = torch.optim.Adam(ref, lr=0.01, weight_decay=1e-3)
    reducer = GradAllReduce(mine).attach()
    opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
    for step in range(12):
        grads = [torch.randn(*s, device="cuda") * (10.0 ** (step % 4 - 2)) for s in shapes]
        for p, q, g in zip(ref, mine, grads):
            p.grad = g.clone()
            q.grad.copy_(g)
        opt_ref.step()
        opt.step()
        assert float(reducer.flat.abs().max()) == 0.0                       # gradients cleared for the next backward
    assert opt.steps == 12
    for p, q in zip(ref, mine):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) <= 2e-6


def test_fused_adam_train_step_follows_torch_adam():
    """parallel.FusedAdam inside the step graph against torch.optim.Adam on the same batches: same losses
    (individual weights with vanishing gradients may differ: Adam's m / sqrt(v) amplifies rounding there)."""
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce
    from pcgnn_b200.runtime import GraphedTrainStep
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny", seed=11)
    rng = np.random.default_rng(4)
    params = random_params(rng, d.feat.shape[1], 16, 3)
    tp = sorted(d.train_pos)
    batches = [rng.choice(d.idx_train, 64) for _ in range(6)]
    results = []
    for fused in (False, True):
        model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        reducer = GradAllReduce(model.parameters()).attach()
        if fused:
            opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
        else:
            opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=0.01, weight_decay=1e-3,
                                   capturable=True)
        eng = model.inter1.engine()
        eng.set_features(model.inter1.features.weight)
        cap = max(eng.slots_bound(np.asarray(b, dtype=np.int32), [0.5] * 3, 0.5, True) for b in batches)
        g = GraphedTrainStep(model, opt, 64, cap, reducer=reducer, warmup_batch=(batches[0], d.labels[batches[0]]))
        losses = [float(g.run(b, d.labels[b]).item()) for b in batches]
        assert not g.overflowed()
        results.append((losses, {k: v.detach().cpu().numpy().copy() for k, v in model.named_parameters()
                                 if v.requires_grad}))
    (l0, p0), (l1, p1) = results
    assert np.allclose(l0, l1, rtol=5e-5, atol=0)
    for k in p0:
        assert rel_err(p1[k], p0[k]) <= 5e-3, k


@pytest.mark.parametrize("concat_self", [False, True])
def test_sage_encoder_native_kernels(concat_self):
    """GraphSAGE encoder with and without the self half (graphsage.py:145-149) through pcg_encoder_* and the
    head + cross-entropy through pcg_head_loss_* (lambda = 0): no torch / cuBLAS op on the path."""
    import torch.nn as nn
    from pcgnn_b200 import graphsage as gs
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny_amz", seed=49)
    rng = np.random.default_rng(6)
    F_, E = d.feat.shape[1], 48
    enc_w = port.xavier(rng, E, 2 * F_ if concat_self else F_)
    head = port.xavier(rng, 2, E)
    nodes = rng.choice(d.idx_train, 70).tolist()
    labels = d.labels[nodes]
    features = nn.Embedding(*d.feat.shape)
    features.weight = nn.Parameter(torch.from_numpy(d.feat), requires_grad=False)
    features = features.cuda()
    enc = gs.Encoder(features, F_, E, d.homo, gs.MeanAggregator(features, cuda=True), gcn=not concat_self, cuda=True)
    model = gs.GraphSage(2, enc)
    with torch.no_grad():
        enc.weight.copy_(torch.from_numpy(enc_w))
        model.weight.copy_(torch.from_numpy(head))
    model = model.cuda()
    loss = model.loss(nodes, torch.from_numpy(labels).cuda())
    assert type(loss.grad_fn).__name__.startswith("HeadLossFn")
    loss.backward()
    pm = port.PortSAGE(d.feat, d.homo, enc_w, head, concat_self=concat_self)
    ref = pm.loss(nodes, labels)
    ref.backward()
    assert abs(loss.item() - float(ref)) <= TOL * abs(float(ref))
    assert rel_err(model.weight.grad.cpu().numpy(), pm.head.grad.numpy()) <= GTOL
    assert rel_err(enc.weight.grad.cpu().numpy(), pm.enc_w.grad.numpy()) <= GTOL
    with torch.no_grad():
        emb = enc(nodes)
    assert rel_err(emb.cpu().numpy(), pm.last["combined"].detach().numpy()) <= TOL
