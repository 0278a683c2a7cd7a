"""GPU: one full train step at the shapes bench.py measures (BASELINE.json configs C1-C4), against the oracle
port on the same batch: bit-exact kept-id sets, loss / embeddings within 1e-5, every parameter gradient within
1e-4 (label_clf included) and the weights after one Adam step.

C1  amazon   N=11,944 F=25  E=64  B=1024     (k_gemm_p: F is not a multiple of 4)
C2  yelp     N=45,954 F=32  E=64  B=1024     (the bench line; 16-byte cp.async tiles)
C3  yelp100  F=100 E=128 B=4096              (batches > 2048 take the 64x64 GEMM kernel)
C4  amazon union graph, GCN                  (select-all + rsqrt aggregate)

The port is a Python loop per target (like the reference): 1-10 s per step at these sizes.
"""
import numpy as np
import pytest
import torch

from helpers import build_cuda_pcgnn, random_params, rel_err
from oracle import port

pytestmark = pytest.mark.gpu
TOL = 1e-5          # loss, embeddings, logits (north_star: 1e-5 relative, fp32)
GTOL = 1e-4         # parameter gradients (fp32 sums over the batch in a different order)
LR, WD = 0.01, 1e-3

SHAPES = {"C1": ("amazon", 1024, 64), "C2": ("yelp", 1024, 64), "C3": ("yelp100", 4096, 128)}
_cache = {}


def _data(spec):
    from pcgnn_b200.synth import make_graph

    if spec not in _cache:
        _cache.clear()                       # one full-size graph in memory at a time
        _cache[spec] = make_graph(spec, seed=72)
    return _cache[spec]


def _batch(d, B, seed):
    from pcgnn_b200.utils import pick_weights

    rng = np.random.default_rng(seed)
    w = pick_weights(d.idx_train, d.y_train, d.homo)
    nodes = rng.choice(d.idx_train, B, p=w / w.sum())       # degree-biased with replacement, like pick_step
    return nodes, d.labels[nodes]


def _adam_reference(pm, names_to_param):
    """torch.optim.Adam on the port's parameters with the port's gradients: {name: weights after one step}."""
    opt = torch.optim.Adam(pm.parameters(), lr=LR, weight_decay=WD)
    opt.step()
    names = ["weight", "inter1.weight"] + [f"inter1.intra_agg{r + 1}.weight" for r in range(pm.R)] \
        + ["inter1.label_clf.weight", "inter1.label_clf.bias"]
    return {n: p.detach().numpy().copy() for n, p in zip(names, pm.parameters())}


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3"])
def test_full_train_step_matches_port(cfg):
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce

    spec, B, E = SHAPES[cfg]
    d = _data(spec)
    F_ = d.feat.shape[1]
    rng = np.random.default_rng(11)
    params = random_params(rng, F_, E, 3)
    tp = sorted(d.train_pos)
    nodes, labels = _batch(d, B, 5)
    # ---- oracle
    pm = port.PortPCGNN(d.feat, d.graph, tp, params, rho=0.5, alpha=2.0)
    ref_loss = pm.step_loss_backward(nodes.tolist(), labels)
    grads = pm.named_grads()
    table = pm.last["score_table"].detach()
    want_emb = pm.last["combined"].detach().numpy()
    want_center = pm.last["center"].detach().numpy()
    want_logits = pm.last["logits"].detach().numpy()
    # ---- product, same score bits
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params, rho=0.5, alpha=2.0)
    model.inter1.score_override = table[:, 0].contiguous().cuda()
    lab = torch.from_numpy(labels).cuda()
    reducer = GradAllReduce(model.parameters()).attach()
    opt = FusedAdam(reducer, lr=LR, weight_decay=WD)
    loss = model.loss(nodes.tolist(), lab)
    loss.backward()
    torch.cuda.synchronize()
    sel = model.inter1.last_selection
    assert not sel.overflowed()
    got = sel.lists()
    for r in range(3):
        for i in range(B):
            assert got[r * B + i].tolist() == pm.last["sel"][r][i], (cfg, r, i)
    assert abs(loss.item() - ref_loss) <= TOL * abs(ref_loss), (loss.item(), ref_loss)
    got_grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.requires_grad}
    for k, g in got_grads.items():
        assert rel_err(g, grads[k]) <= GTOL, (cfg, k, rel_err(g, grads[k]))
    # embeddings / center scores / logits of the same (train-mode) forward
    with torch.no_grad():
        emb, center = model.inter1(nodes.tolist(), lab, True)
        logits = model.weight.mm(emb).t()
    assert rel_err(emb.cpu().numpy(), want_emb) <= TOL
    assert rel_err(center.cpu().numpy(), want_center) <= TOL
    assert rel_err(logits.cpu().numpy(), want_logits) <= TOL
    # ---- one Adam step (pcg_allreduce_adam) against torch.optim.Adam on the port's gradients. Adam's first
    # update is lr * g / (|g| + eps): only entries whose gradient is far above the gradient tolerance are compared
    before = {k: p.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.requires_grad}
    want_after = _adam_reference(pm, None)
    opt.step()
    torch.cuda.synchronize()
    compared = total = 0
    for k, p in model.named_parameters():
        if not p.requires_grad:
            continue
        g_eff = np.abs(grads[k] + WD * before[k])
        solid = g_eff > 100 * GTOL * np.abs(grads[k]).max()
        diff = np.abs(p.detach().cpu().numpy() - want_after[k])
        assert diff[solid].max(initial=0.0) <= 2e-2 * LR, (cfg, k, float(diff[solid].max()))
        assert diff.max() <= 2.0 * LR + 1e-6, (cfg, k)
        compared += int(solid.sum())
        total += solid.size
    assert compared >= 0.5 * total


def test_full_size_eval_forward_matches_port():
    """C2, train_flag=False (utils.test -> to_prob): kept sets, embeddings and probabilities."""
    d = _data("yelp")
    rng = np.random.default_rng(13)
    params = random_params(rng, 32, 64, 3)
    tp = sorted(d.train_pos)
    nodes = rng.choice(d.idx_rest, 1024, replace=False)
    labels = d.labels[nodes]
    pm = port.PortPCGNN(d.feat, d.graph, tp, params)
    with torch.no_grad():
        logits, center, emb = pm.forward(nodes.tolist(), torch.from_numpy(labels), False)
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    model.inter1.score_override = pm.last["score_table"].detach()[:, 0].contiguous().cuda()
    with torch.no_grad():
        gnn_prob, label_prob = model.to_prob(nodes.tolist(), labels, train_flag=False)
    got = model.inter1.last_selection.lists()
    for r in range(3):
        for i in range(1024):
            assert got[r * 1024 + i].tolist() == pm.last["sel"][r][i]
    assert rel_err(gnn_prob.cpu().numpy(), torch.sigmoid(logits).numpy()) <= TOL
    assert rel_err(label_prob.cpu().numpy(), torch.sigmoid(center).numpy()) <= TOL


def test_full_size_gcn_step_matches_port():
    """C4: GCN(GCNEncoder(GCNAggregator)) on the Amazon-shaped union graph, B = 1024."""
    import torch.nn as nn
    from pcgnn_b200 import graphsage as gs

    d = _data("amazon")
    rng = np.random.default_rng(17)
    F_, E = d.feat.shape[1], 64
    enc_w = port.xavier(rng, E, F_)
    head = port.xavier(rng, 2, E)
    nodes, labels = _batch(d, 1024, 9)
    features = nn.Embedding(*d.feat.shape)
    features.weight = nn.Parameter(torch.from_numpy(d.feat), requires_grad=False)
    features = features.cuda()
    enc = gs.GCNEncoder(features, F_, E, d.homo, gs.GCNAggregator(features, cuda=True), cuda=True)
    model = gs.GCN(2, enc)
    with torch.no_grad():
        enc.weight.copy_(torch.from_numpy(enc_w))
        model.weight.copy_(torch.from_numpy(head))
    model = model.cuda()
    loss = model.loss(nodes.tolist(), torch.from_numpy(labels).cuda())
    loss.backward()
    pm = port.PortGCN(d.feat, d.homo, enc_w, head)
    ref = pm.loss(nodes.tolist(), labels)
    ref.backward()
    assert abs(loss.item() - float(ref)) <= TOL * abs(float(ref))
    assert rel_err(model.weight.grad.cpu().numpy(), pm.head.grad.numpy()) <= GTOL
    assert rel_err(enc.weight.grad.cpu().numpy(), pm.enc_w.grad.numpy()) <= GTOL


def test_graphed_step_equals_eager_at_bench_shape():
    """C2: the captured step graph (what bench.py times) gives the same loss trajectory and weights as the
    eager reference-API calls over 4 batches, and its first loss equals the port's."""
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce
    from pcgnn_b200.runtime import GraphedTrainStep

    d = _data("yelp")
    rng = np.random.default_rng(19)
    params = random_params(rng, 32, 64, 3)
    tp = sorted(d.train_pos)
    batches = [_batch(d, 1024, 100 + s) for s in range(4)]
    out = []
    for graphed in (False, True):
        model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        reducer = GradAllReduce(model.parameters()).attach()
        opt = FusedAdam(reducer, lr=LR, weight_decay=WD)
        losses = []
        if graphed:
            eng = model.inter1.engine()
            eng.set_features(model.inter1.features.weight)
            cap = GraphedTrainStep.plan(eng, [n for n, _ in batches], model.inter1.thresholds, 0.5)
            step = GraphedTrainStep(model, opt, 1024, cap, reducer=reducer, warmup_batch=batches[0])
            for n, l in batches:
                losses.append(step.run(n.tolist(), l).item())
            assert not step.overflowed()
        else:
            for n, l in batches:
                loss = model.loss(n.tolist(), torch.from_numpy(l).cuda())
                loss.backward()
                opt.step()
                losses.append(loss.item())
        out.append((losses, [p.detach().cpu().numpy().copy() for p in model.parameters() if p.requires_grad]))
    assert np.allclose(out[0][0], out[1][0], rtol=1e-5), out
    for a, b in zip(out[0][1], out[1][1]):
        assert rel_err(b, a) <= 1e-4
    pm = port.PortPCGNN(d.feat, d.graph, tp, params)
    ref = pm.step_loss_backward(batches[0][0].tolist(), batches[0][1])
    assert abs(out[1][0][0] - ref) <= 1e-4 * abs(ref)       # own score table: selection may differ within an ulp
