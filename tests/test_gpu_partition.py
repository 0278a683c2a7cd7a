"""GPU: row-partitioned CSR (config C5's layout). A graph generated on the device is cut into row ranges;
every partition's engine, given only its own rows (global neighbour ids, global score table and pool), must
choose exactly what the C oracle chooses on the whole graph."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _host_graph(part):
    from pcgnn_b200.graph import RelGraph

    ip, ix = part.graph.device("cuda")
    ip, ix = ip.cpu().numpy(), ix.cpu().numpy()
    n = part.graph.n_nodes
    ips, ixs = [], []
    for r in range(part.graph.n_rel):
        seg = ip[r * n:(r + 1) * n + 1]
        ips.append(seg - seg[0])
        ixs.append(ix[seg[0]:seg[-1]])
    return RelGraph(n, ips, ixs)


@pytest.mark.parametrize("world", [2, 4])
def test_partitions_choose_like_the_whole_graph(world):
    from pcgnn_b200.engine import Engine
    from pcgnn_b200.synth_big import BigSpec, make_partition

    n_global = 24000
    mk = lambda rows, rank, w: make_partition(BigSpec(nodes_per_rank=rows, feat_dim=8, rel_mean_deg=(2.0, 8.0, 20.0),
                                                      max_degree=40000, seed=5), rank, w, "cuda")
    full = mk(n_global, 0, 1)
    host = _host_graph(full)
    assert np.diff(host.indptr).max() > 1024                      # hubs reach the cluster tier
    rng = np.random.default_rng(world)
    score = rng.normal(size=n_global).astype(np.float32)
    pool = full.train_pos.cpu().numpy()
    feat = full.feat.cpu().numpy()
    labels = full.labels.cpu().numpy()
    for rank in range(world):
        part = mk(n_global // world, rank, world)
        assert torch.equal(part.train_pos, full.train_pos) and torch.equal(part.feat, full.feat)
        nodes, lab = part.sample_batches(1, 96, seed=rank)[0]
        nodes_h = nodes.cpu().numpy().astype(np.int64)
        assert nodes_h.min() >= part.row_lo and nodes_h.max() < part.row_lo + part.graph.n_nodes
        assert np.array_equal(lab.cpu().numpy(), labels[nodes_h])
        eng = Engine(part.graph, "cuda")
        eng.set_features(part.feat)
        eng.set_pool(part.train_pos)
        eng.score.copy_(torch.from_numpy(score).cuda())
        eng.resort_pool()
        cap = eng.slots_bound(nodes_h, [0.5] * 3, 0.5, True)
        sel = eng.choose(nodes, lab, True, [0.5] * 3, 0.5, cap)
        agg = eng.aggregate(sel)
        torch.cuda.synchronize()
        assert not sel.overflowed()
        sp, si = c_oracle.choose(host, score, nodes_h, labels[nodes_h] == 1, rho=0.5, pool=pool, train=True)
        got = sel.lists()
        for w in range(3 * len(nodes_h)):
            assert np.array_equal(got[w], si[sp[w]:sp[w + 1]]), (rank, w)
        assert rel_err(agg.cpu().numpy()[:, :8], c_oracle.aggregate(feat, sp, si)) <= 1e-5


def test_target_outside_the_partition_is_flagged():
    from pcgnn_b200.engine import Engine
    from pcgnn_b200.synth_big import BigSpec, make_partition

    part = make_partition(BigSpec(nodes_per_rank=3000, feat_dim=8, rel_mean_deg=(2.0, 4.0, 6.0), seed=9), 1, 2, "cuda")
    eng = Engine(part.graph, "cuda")
    eng.set_features(part.feat)
    eng.set_pool(part.train_pos)
    eng.score.zero_()
    eng.resort_pool()
    nodes = torch.tensor([3000, 3001, 17, 5999], dtype=torch.int32, device="cuda")       # 17 belongs to rank 0
    lab = torch.zeros(4, dtype=torch.int64, device="cuda")
    sel = eng.choose(nodes, lab, True, [0.5] * 3, 0.5, 4096)
    torch.cuda.synchronize()
    assert int(sel.status[3].item()) == 2
    m, _ = sel.item_sizes()
    assert m[2] == 0 and m[0] > 0
