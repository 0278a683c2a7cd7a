"""GPU: the fused dense kernels (csrc/pcg_tile.cu: pcg_tile_fwd / pcg_tile_train) against the oracle port on
small graphs, over the shapes that select every tile size, K-split and padding path:
F in {12, 25, 32, 100} (F % 4 != 0 pads the operand rows), E in {64, 128, 192}, R in {1, 3, 5}, batches that are
not a multiple of the tile rows, duplicated targets. The full-size shapes are in test_gpu_fullsize.py."""
import numpy as np
import pytest
import torch

from helpers import build_cuda_pcgnn, random_params, rel_err
from oracle import port

pytestmark = pytest.mark.gpu
TOL, GTOL = 1e-5, 1e-4


def _graph(F_, R, seed):
    """tiny synthetic graph with F_ features and R relations"""
    from pcgnn_b200.graph import RelGraph
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny", seed=seed)
    rng = np.random.default_rng(seed)
    feat = rng.random((d.feat.shape[0], F_), dtype=np.float32)
    rels = [d.graph.relation(r % 3) for r in range(R)]
    graph = RelGraph(d.graph.n_nodes, [np.asarray(ip) for ip, _ in rels], [np.asarray(ix) for _, ix in rels])
    return d, feat, graph


@pytest.mark.parametrize("F_,E,R,B", [(12, 64, 3, 100), (25, 64, 3, 64), (32, 64, 3, 1), (32, 128, 3, 77),
                                      (100, 128, 3, 50), (12, 192, 1, 33), (25, 64, 5, 200), (32, 256, 3, 40),
                                      (32, 64, 3, 1500), (32, 64, 3, 4200), (12, 128, 3, 2100)])
def test_tile_train_step_matches_port(F_, E, R, B):
    d, feat, graph = _graph(F_, R, 100 + F_ + E + R)
    rng = np.random.default_rng(B)
    params = random_params(rng, F_, E, R)
    tp = sorted(d.train_pos)
    nodes = rng.choice(d.idx_train, B)                       # with replacement: duplicated targets
    labels = d.labels[nodes]
    pm = port.PortPCGNN(feat, graph, tp, params)
    if B == 1:     # the port (like the reference, layers.py:243) needs B >= 2: duplicate the target for the oracle
        ref = pm.step_loss_backward(nodes.tolist() * 2, np.repeat(labels, 2))
    else:
        ref = pm.step_loss_backward(nodes.tolist(), labels)
    grads = pm.named_grads()
    model = build_cuda_pcgnn(feat, graph, tp, params)
    eng = model.inter1.engine()
    eng.set_features(model.inter1.features.weight)
    assert eng.tile_supported(B, R, E)
    model.inter1.score_override = pm.last["score_table"].detach()[:, 0].contiguous().cuda()
    lab = torch.from_numpy(labels).cuda()
    loss = model.loss(nodes.tolist(), lab)                   # TrainStepFn (pcg_tile_train)
    assert type(loss.grad_fn).__name__.startswith("TrainStepFn")
    loss.backward()
    assert abs(loss.item() - ref) <= TOL * abs(ref), (loss.item(), ref)
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert rel_err(p.grad.cpu().numpy(), grads[k]) <= GTOL, (k, rel_err(p.grad.cpu().numpy(), grads[k]))
    # the autograd composition the reference's own model.py uses: forward() (pcg_tile_fwd) + torch head / losses
    model.zero_grad()
    logits, center = model.forward(nodes.tolist(), lab, True)
    assert type(center.grad_fn).__name__.startswith("_TileFn")
    loss2 = torch.nn.functional.cross_entropy(logits, lab) + 2.0 * torch.nn.functional.cross_entropy(center, lab)
    loss2.backward()
    assert abs(loss2.item() - ref) <= TOL * abs(ref)
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert rel_err(p.grad.cpu().numpy(), grads[k]) <= GTOL, ("autograd", k)
    if B > 1:
        assert rel_err(logits.detach().cpu().numpy(), pm.last["logits"].detach().numpy()) <= TOL
        assert rel_err(center.detach().cpu().numpy(), pm.last["center"].detach().numpy()) <= TOL
    # inference mode (no cat kept)
    with torch.no_grad():
        emb, c2 = model.inter1(nodes.tolist(), lab, True)
    if B > 1:
        assert rel_err(emb.cpu().numpy(), pm.last["combined"].detach().numpy()) <= TOL
    assert torch.equal(c2, center.detach())


def test_tile_train_is_deterministic_and_accumulates_like_autograd():
    """Two runs give bit-identical gradients; a second backward accumulates into .grad like any autograd node;
    a scaled loss scales the gradients."""
    d, feat, graph = _graph(32, 3, 7)
    rng = np.random.default_rng(3)
    params = random_params(rng, 32, 64, 3)
    tp = sorted(d.train_pos)
    nodes = rng.choice(d.idx_train, 300)
    lab = torch.from_numpy(d.labels[nodes]).cuda()
    model = build_cuda_pcgnn(feat, graph, tp, params)
    model.loss(nodes.tolist(), lab).backward()
    g1 = [p.grad.clone() for p in model.parameters() if p.requires_grad]
    model.zero_grad()
    (3.0 * model.loss(nodes.tolist(), lab)).backward()
    g3 = [p.grad.clone() for p in model.parameters() if p.requires_grad]
    model.loss(nodes.tolist(), lab).backward()               # accumulates: 3g + g
    g4 = [p.grad.clone() for p in model.parameters() if p.requires_grad]
    for a, b, c in zip(g1, g3, g4):
        assert torch.equal(a * 3.0, b)
        assert torch.allclose(c, a * 4.0, rtol=1e-6, atol=0)
    model.zero_grad()
    model.loss(nodes.tolist(), lab).backward()
    for a, p in zip(g1, [p for p in model.parameters() if p.requires_grad]):
        assert torch.equal(a, p.grad)


def test_unsupported_embed_dim_takes_the_gemm_kernels():
    """E = 16 is outside the tile kernel's shapes: same results through pcg_dense_* / pcg_head_*."""
    d, feat, graph = _graph(12, 3, 9)
    rng = np.random.default_rng(5)
    params = random_params(rng, 12, 16, 3)
    tp = sorted(d.train_pos)
    nodes = rng.choice(d.idx_train, 60)
    labels = d.labels[nodes]
    pm = port.PortPCGNN(feat, graph, tp, params)
    ref = pm.step_loss_backward(nodes.tolist(), labels)
    model = build_cuda_pcgnn(feat, graph, tp, params)
    model.inter1.score_override = pm.last["score_table"].detach()[:, 0].contiguous().cuda()
    loss = model.loss(nodes.tolist(), torch.from_numpy(labels).cuda())
    assert type(loss.grad_fn).__name__.startswith("HeadLossFn")
    loss.backward()
    assert abs(loss.item() - ref) <= TOL * abs(ref)
    grads = pm.named_grads()
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert rel_err(p.grad.cpu().numpy(), grads[k]) <= GTOL, k
