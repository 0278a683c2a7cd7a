"""GPU + reference tree: the reference's OWN callers, unchanged, on top of the CUDA modules.

  * /root/reference/src/model.py  PCALayer.loss / backward / to_prob  over pcgnn_b200.layers.InterAgg3
  * /root/reference/src/model_handler.py  ModelHandler(config).train()  for two epochs, with `load_data` answering
    from the synthetic generator (the dataset pickles are not available offline) and the result bookkeeping pointed
    at a temp directory

Runs wherever BOTH a CUDA device and the reference tree (PCGNN_REFERENCE_ROOT, default /root/reference) exist. The
build container has the tree but no GPU and the GPU boxes have no tree (reference sources may not be copied into
this repository), so on the driver's boxes this file skips; tests/test_host_logic.py checks the import wiring on the
CPU, and every other GPU test exercises the same modules through the mirror of model.py.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_harness as H

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not H.available(), reason="reference tree not present on this box")]


def _install():
    import pcgnn_b200.shim as shim

    if H.REF_ROOT not in sys.path:
        sys.path.insert(0, H.REF_ROOT)
    shim.install(reference_root=H.REF_ROOT)
    return shim


def test_reference_pcalayer_runs_on_the_cuda_modules():
    from helpers import random_params, rel_err
    from oracle import port
    from pcgnn_b200.layers import InterAgg3, IntraAgg
    from pcgnn_b200.synth import make_graph

    shim = _install()
    try:
        ref_model = importlib.import_module("src.model")
        assert ref_model.__file__.startswith(H.REF_ROOT)
        d = make_graph("tiny_amz", seed=71)
        rng = np.random.default_rng(1)
        F_, E = d.feat.shape[1], 64
        params = random_params(rng, F_, E, 3)
        tp = sorted(d.train_pos)
        features = torch.nn.Embedding(*d.feat.shape)
        features.weight = torch.nn.Parameter(torch.from_numpy(d.feat), requires_grad=False)
        intras = [IntraAgg(features, F_, E, tp, 0.5, cuda=True) for _ in range(3)]
        inter = InterAgg3(features, F_, E, tp, d.graph, intras, cuda=True)
        model = ref_model.PCALayer(2, inter, 2.0)                      # the reference's class, unchanged
        with torch.no_grad():
            for ia, w in zip(intras, params["intra"]):
                ia.weight.copy_(torch.from_numpy(w))
            inter.weight.copy_(torch.from_numpy(params["inter"]))
            inter.label_clf.weight.copy_(torch.from_numpy(params["clf_w"]))
            inter.label_clf.bias.copy_(torch.from_numpy(params["clf_b"]))
            model.weight.copy_(torch.from_numpy(params["head"]))
        model = model.to("cuda")
        nodes = rng.choice(d.idx_train, 90).tolist()
        labels = d.labels[nodes]
        pm = port.PortPCGNN(d.feat, d.graph, tp, params)
        want = pm.step_loss_backward(nodes, labels)
        inter.score_override = pm.last["score_table"].detach()[:, 0].contiguous().cuda()
        loss = model.loss(nodes, torch.from_numpy(labels).cuda())      # model.py:47-61
        loss.backward()
        assert abs(loss.item() - want) <= 1e-5 * abs(want)
        grads = pm.named_grads()
        for k, p in model.named_parameters():
            if p.requires_grad:
                assert rel_err(p.grad.cpu().numpy(), grads[k]) <= 1e-4, k
        with torch.no_grad():
            gnn_prob, label_prob = model.to_prob(nodes, labels, train_flag=False)   # model.py:41-45
        assert gnn_prob.shape == (90, 2) and label_prob.shape == (90, 2)
    finally:
        shim.uninstall()


def test_reference_model_handler_trains_two_epochs(tmp_path, monkeypatch):
    from pcgnn_b200.synth import make_graph

    shim = _install()
    try:
        mh = importlib.import_module("src.model_handler")
        d = make_graph("train_sig", seed=5, signal=0.25, homophily=0.3)
        monkeypatch.setattr(mh, "load_data", lambda name: (d.homo, d.graph, d.feat, d.labels))
        monkeypatch.chdir(tmp_path)
        config = dict(seed=72, data_name="yelp", model="PCGNN", train_ratio=0.4, test_ratio=0.67, emb_size=64, lr=0.01,
                      weight_decay=1e-3, alpha=2, rho=0.5, epochs=2, valid_epochs=1, batch_size=100, patience=100, exp_num=0)
        auc, recall, f1 = mh.ModelHandler(config).train()              # model_handler.py:22-178, unchanged
        assert 0.0 <= auc <= 1.0 and 0.0 <= f1 <= 1.0
    finally:
        shim.uninstall()
