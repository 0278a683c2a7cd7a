"""GPU: the CUDA-graph cache behind the reference-facing calls (pcgnn_b200/stepgraph.py) and the pinned staging
rings: model.loss(list, labels) / backward / optimizer.step in the reference's own loop shape
(model_handler.py:142-156) must give the same trajectory as the uncached eager kernels, whatever changes between
calls (batch size, slot demand, parameter storage), and back-to-back calls must not race their uploads."""
import numpy as np
import pytest
import torch

from helpers import build_cuda_pcgnn, random_params, rel_err
from oracle import port

pytestmark = pytest.mark.gpu


def _setup(seed=61, E=64, spec="tiny_amz"):
    from pcgnn_b200.synth import make_graph

    d = make_graph(spec, seed=seed, dup_feature_frac=0.1)
    rng = np.random.default_rng(seed)
    params = random_params(rng, d.feat.shape[1], E, 3)
    return d, rng, params, sorted(d.train_pos)


def _train(model, batches, labels_of, cached):
    model.inter1.graph_cache = cached
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=0.01, weight_decay=1e-3)
    losses = []
    for nodes in batches:
        opt.zero_grad()
        loss = model.loss(nodes.tolist(), torch.from_numpy(labels_of(nodes)).cuda())   # model_handler.py:150
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return losses


def test_cached_training_loop_equals_eager_loop():
    d, rng, params, tp = _setup()
    # batch sizes change (last batch of an epoch is short, model_handler.py:134), demand for slots grows (hubs later)
    deg = d.homo.degrees(0)
    order = np.asarray(d.idx_train)[np.argsort(deg[d.idx_train])]
    small, hubs = order[:200], order[-128:]
    batches = [rng.choice(small, 128), rng.choice(small, 128), hubs, rng.choice(d.idx_train, 77),
               rng.choice(d.idx_train, 128), rng.choice(d.idx_train, 77)]
    out = []
    for cached in (False, True):
        model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        losses = _train(model, batches, lambda n: d.labels[n], cached)
        out.append((losses, [p.detach().cpu().numpy().copy() for p in model.parameters() if p.requires_grad], model))
    assert np.allclose(out[0][0], out[1][0], rtol=1e-5, atol=1e-6), (out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert rel_err(b, a) <= 1e-4
    cache = out[1][2].inter1.graphs()
    assert cache.replays == len(batches)
    assert cache.captures == 2          # two batch sizes
    # a batch that needs more slots than the recorded capacity re-records (forced here by shrinking the record's)
    model = out[1][2]
    slot = next(s for k, s in cache.slots.items() if k[1] == 128)
    slot.cap = 8
    l_c = model.loss(hubs.tolist(), torch.from_numpy(d.labels[hubs]).cuda())
    assert cache.captures == 3 and not model.inter1.engine().overflow_since_reset()
    eager = out[0][2]
    l_e = eager.loss(hubs.tolist(), torch.from_numpy(d.labels[hubs]).cuda())
    assert abs(l_c.item() - l_e.item()) <= 1e-4 * abs(l_e.item())     # (weights of the two runs agree to 1e-4)
    assert out[0][2].inter1.graphs().replays == 0
    # first loss equals the oracle's (own score table: selection may differ within an ulp of a distance)
    pm = port.PortPCGNN(d.feat, d.graph, tp, params)
    ref = pm.step_loss_backward(batches[0].tolist(), d.labels[batches[0]])
    assert abs(out[1][0][0] - ref) <= 1e-4 * abs(ref)


def test_cached_graph_follows_parameter_storage_changes():
    """FusedAdam re-points every parameter into its flat buffer: the cached graph must notice (addresses change)."""
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce

    d, rng, params, tp = _setup(seed=62)
    nodes = rng.choice(d.idx_train, 96)
    lab = torch.from_numpy(d.labels[nodes]).cuda()
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    l0 = model.loss(nodes.tolist(), lab)
    l0.backward()
    g0 = [p.grad.clone() for p in model.parameters() if p.requires_grad]
    model.zero_grad()
    reducer = GradAllReduce(model.parameters()).attach()
    FusedAdam(reducer, lr=0.01)                               # parameters now live in another buffer
    with torch.no_grad():
        for p in model.parameters():
            if p.requires_grad:
                p.mul_(1.5)
    l1 = model.loss(nodes.tolist(), lab)
    l1.backward()
    ref = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    ref.inter1.graph_cache = False
    with torch.no_grad():
        for p in ref.parameters():
            if p.requires_grad:
                p.mul_(1.5)
    l2 = ref.loss(nodes.tolist(), lab)
    l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-6 * abs(l2.item())
    assert abs(l1.item() - l0.item()) > 1e-3                  # and it did see the new weights
    for p, q in zip([p for p in model.parameters() if p.requires_grad], [p for p in ref.parameters() if p.requires_grad]):
        assert rel_err(p.grad.cpu().numpy(), q.grad.cpu().numpy()) <= 1e-5
    assert model.inter1.graphs().captures == 2


def test_cached_eval_forward_equals_eager():
    d, rng, params, tp = _setup(seed=63)
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    eager = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    eager.inter1.graph_cache = False
    for B in (100, 100, 37):
        nodes = rng.choice(d.idx_rest, B, replace=False)
        labels = d.labels[nodes]                              # numpy, as utils.test passes them (utils.py:303)
        with torch.no_grad():
            a = model.to_prob(nodes.tolist(), labels, train_flag=False)
            b = eager.to_prob(nodes.tolist(), labels, train_flag=False)
        assert torch.allclose(a[0], b[0], rtol=1e-6, atol=1e-7) and torch.allclose(a[1], b[1], rtol=1e-6, atol=1e-7)
        got, want = model.inter1.last_selection.lists(), eager.inter1.last_selection.lists()
        assert all(np.array_equal(x, y) for x, y in zip(got, want))
    assert model.inter1.graphs().replays == 3 and model.inter1.graphs().captures == 2


def test_back_to_back_calls_do_not_race_their_uploads():
    """No host sync between calls: every call must train on ITS batch. Cached loss() and
    GraphedTrainStep.run() against the same steps with a sync after each."""
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce
    from pcgnn_b200.runtime import GraphedTrainStep

    d, rng, params, tp = _setup(seed=64)
    B = 128
    batches = [rng.choice(d.idx_train, B) for _ in range(12)]
    # (a) cached reference-facing call
    res = []
    for sync in (True, False):
        model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        losses = []
        for n in batches:
            losses.append(model.loss(n.tolist(), d.labels[n]).detach())
            if sync:
                torch.cuda.synchronize()
        res.append(torch.stack(losses).cpu().numpy())
    assert np.array_equal(res[0], res[1])
    # (b) GraphedTrainStep.run
    res = []
    for sync in (True, False):
        model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        reducer = GradAllReduce(model.parameters()).attach()
        opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
        eng = model.inter1.engine()
        eng.set_features(model.inter1.features.weight)
        cap = GraphedTrainStep.plan(eng, batches, model.inter1.thresholds, 0.5)
        step = GraphedTrainStep(model, opt, B, cap, reducer=reducer, warmup_batch=(batches[0], d.labels[batches[0]]))
        losses = []
        for n in batches:
            losses.append(step.run(n, d.labels[n]).clone())
            if sync:
                torch.cuda.synchronize()
        res.append(torch.stack(losses).cpu().numpy())
        assert not step.overflowed()
    assert np.array_equal(res[0], res[1])


def test_overflow_flag_is_sticky_across_replays():
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce
    from pcgnn_b200.runtime import GraphedTrainStep

    d, rng, params, tp = _setup(seed=65)
    B = 64
    deg = d.homo.degrees(0)
    order = np.asarray(d.idx_train)[np.argsort(deg[d.idx_train])]
    small, big = order[:B], order[-B:]
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    reducer = GradAllReduce(model.parameters()).attach()
    opt = FusedAdam(reducer, lr=0.01)
    eng = model.inter1.engine()
    eng.set_features(model.inter1.features.weight)
    cap = GraphedTrainStep.plan(eng, [small], model.inter1.thresholds, 0.5)       # too small for the hub batch
    step = GraphedTrainStep(model, opt, B, cap, reducer=reducer, warmup_batch=(small, d.labels[small]))
    step.run(small, d.labels[small])
    assert not step.overflowed()
    step.run(big, d.labels[big])              # overflows ...
    step.run(small, d.labels[small])          # ... and a later, fitting replay must not hide it
    assert step.overflowed()
    assert not step.overflowed()              # reading clears it


def test_host_batches_equal_device_batches():
    """GraphedTrainStep.run / run_item (the recorded step fetches the batch from pinned host memory by itself and stores
    the loss into a pinned word: pcg_pool_scores_stage / pcg_stage) against run_device on the same batches, from the
    same initial state: identical losses, bit for bit, and identical parameters afterwards."""
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce
    from pcgnn_b200.runtime import GraphedTrainStep

    d, rng, params, tp = _setup(seed=66)
    B = 128
    batches = [rng.choice(d.idx_train, B) for _ in range(8)]
    res, weights = [], []
    for mode in ("device", "host", "host_item"):
        model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        reducer = GradAllReduce(model.parameters()).attach()
        opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
        eng = model.inter1.engine()
        eng.set_features(model.inter1.features.weight)
        cap = GraphedTrainStep.plan(eng, batches, model.inter1.thresholds, 0.5)
        step = GraphedTrainStep(model, opt, B, cap, reducer=reducer, warmup_batch=(batches[0], d.labels[batches[0]]))
        losses = []
        for n in batches:
            if mode == "device":
                loss = step.run_device(torch.from_numpy(n.astype(np.int32)).cuda(), torch.from_numpy(d.labels[n]).cuda())
                losses.append(float(loss.item()))
            elif mode == "host":
                losses.append(float(step.run(n, d.labels[n]).item()))
            else:
                losses.append(step.run_item(n.tolist(), d.labels[n]))
        assert not step.overflowed()
        res.append(np.asarray(losses, dtype=np.float32))
        weights.append(reducer.flat_params.detach().cpu().numpy().copy() if hasattr(reducer, "flat_params")
                       else np.concatenate([p.detach().cpu().numpy().ravel() for p in model.parameters() if p.requires_grad]))
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
    assert np.array_equal(weights[0], weights[1]) and np.array_equal(weights[0], weights[2])


def test_stage_copies_through_pinned_host_memory():
    from pcgnn_b200 import _lib

    L = _lib.lib()
    for n in (1, 3, 1024, 3 * 1024 + 5, 100_003):
        src = torch.arange(n, dtype=torch.int32).pin_memory()
        dst = torch.zeros(n, dtype=torch.int32, device="cuda")
        _lib.check(L.pcg_stage(_lib.host_device_ptr(src), dst.data_ptr(), n * 4, _lib.stream_ptr()), "pcg_stage")
        back = torch.zeros(n, dtype=torch.int32).pin_memory()
        _lib.check(L.pcg_stage(dst.data_ptr(), _lib.host_device_ptr(back), n * 4, _lib.stream_ptr()), "pcg_stage")
        torch.cuda.synchronize()
        assert torch.equal(dst.cpu(), src) and torch.equal(back, src)
    with pytest.raises(_lib.PcgError):
        _lib.check(L.pcg_stage(dst.data_ptr(), dst.data_ptr() + 2, 4, _lib.stream_ptr()), "pcg_stage")


def test_fast_reference_loop_matches_the_plain_loop():
    """fastloop.enable(): the reference's loop unchanged (zero_grad / model.loss / backward / torch.optim.Adam.step),
    with backward() storing the replay's gradients directly and the caller's Adam served by the one-kernel Adam.
    Same gradients bit for bit after backward(); the same training trajectory as torch's own Adam within the rounding of
    the two update formulas; anything unusual (explicit gradient argument, another optimizer) takes torch's own route."""
    from pcgnn_b200 import fastloop

    d, rng, params, tp = _setup(seed=67)
    B = 128
    batches = [rng.choice(d.idx_train, B) for _ in range(12)]

    def loop(fast, optim_cls=torch.optim.Adam):
        model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        opt = optim_cls([p for p in model.parameters() if p.requires_grad], lr=0.01, weight_decay=1e-3)
        if fast:
            fastloop.enable()
        try:
            losses, first_grads = [], None
            for n in batches:
                opt.zero_grad()
                loss = model.loss(n.tolist(), torch.from_numpy(d.labels[n]).cuda())
                assert isinstance(loss, fastloop.StepLoss) == fast
                loss.backward()
                if first_grads is None:
                    first_grads = [p.grad.detach().clone() for p in model.parameters() if p.requires_grad]
                opt.step()
                losses.append(loss.item())
            weights = [p.detach().cpu().numpy().copy() for p in model.parameters() if p.requires_grad]
            taken = "_pcg_flat" in opt.__dict__
        finally:
            fastloop.disable()
        return np.asarray(losses), first_grads, weights, taken

    l0, g0, w0, t0 = loop(False)
    l1, g1, w1, t1 = loop(True)
    assert not t0 and t1
    assert all(torch.equal(a, b) for a, b in zip(g0, g1))
    assert rel_err(l1, l0) <= 1e-4
    assert max(rel_err(a, b) for a, b in zip(w1, w0)) <= 1e-3
    # another optimizer: the hook must leave it alone, the fast backward still feeds it
    l2, _, w2, t2 = loop(True, torch.optim.SGD)
    l3, _, w3, _ = loop(False, torch.optim.SGD)
    assert not t2 and np.array_equal(l2, l3) and all(np.array_equal(a, b) for a, b in zip(w2, w3))
    # explicit gradient argument: torch's own backward
    model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
    fastloop.enable()
    try:
        n = batches[0]
        loss = model.loss(n.tolist(), torch.from_numpy(d.labels[n]).cuda())
        loss.backward(torch.tensor(2.0, device="cuda"))
        got = [p.grad.detach().clone() for p in model.parameters() if p.requires_grad]
    finally:
        fastloop.disable()
    assert max(rel_err(a.cpu().numpy(), 2.0 * b.cpu().numpy()) for a, b in zip(got, g0)) <= 1e-6


def test_epoch_plan_equals_device_batches():
    """GraphedTrainStep.load_plan / run_planned (every batch of the epoch resident in HBM, a device cursor, the recorded
    step fetches its batch and files its loss by itself: pcg_pool_scores_stage / pcg_stage_indexed) against run_device
    on the same sequence of batches, two epochs (the cursor wraps): identical losses and weights, bit for bit."""
    from pcgnn_b200.parallel import FusedAdam, GradAllReduce
    from pcgnn_b200.runtime import GraphedTrainStep

    d, rng, params, tp = _setup(seed=68)
    B, n = 128, 5
    batches = [rng.choice(d.idx_train, B) for _ in range(n)]
    out = []
    for mode in ("device", "plan"):
        model = build_cuda_pcgnn(d.feat, d.graph, tp, params)
        reducer = GradAllReduce(model.parameters()).attach()
        opt = FusedAdam(reducer, lr=0.01, weight_decay=1e-3)
        eng = model.inter1.engine()
        eng.set_features(model.inter1.features.weight)
        cap = GraphedTrainStep.plan(eng, batches, model.inter1.thresholds, 0.5)
        step = GraphedTrainStep(model, opt, B, cap, reducer=reducer, warmup_batch=(batches[0], d.labels[batches[0]]))
        losses = []
        if mode == "device":
            for epoch in range(2):
                for b in batches:
                    loss = step.run_device(torch.from_numpy(b.astype(np.int32)).cuda(), torch.from_numpy(d.labels[b]).cuda())
                    losses.append(float(loss.item()))
        else:
            assert step.load_plan([(b, d.labels[b]) for b in batches]) == n
            for epoch in range(2):
                for _ in range(n):
                    step.run_planned()                      # no host data, no sync
                losses += step.plan_losses().cpu().tolist()
            assert int(step._cursor.item()) == 2 * n
        assert not step.overflowed()
        weights = np.concatenate([p.detach().cpu().numpy().ravel() for p in model.parameters() if p.requires_grad])
        out.append((np.asarray(losses, dtype=np.float32), weights))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
