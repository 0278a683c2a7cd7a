"""CPU: the two oracle restatements against each other on seeded random graphs (the C restatement is what the
full-size GPU parity tests use; the Python port is the one pinned to the live reference's golden vectors), and
the bench's CPU arm / JSON contract."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import random_params
from oracle import c_oracle, port

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.parametrize("seed,dup,rho,train", [(1, 0.0, 0.5, True), (2, 0.4, 0.5, True), (3, 0.9, 1.3, True),
                                                (4, 0.2, 0.5, False), (5, 0.0, 0.2, True)])
def test_c_oracle_equals_port_on_random_graphs(seed, dup, rho, train):
    from pcgnn_b200.synth import make_graph

    d = make_graph("tiny", seed=seed, dup_feature_frac=dup)
    rng = np.random.default_rng(seed)
    params = random_params(rng, d.feat.shape[1], 8, 3)
    pool = sorted(d.train_pos) if seed % 2 else list(d.train_pos)         # sorted and idx_train order (ties by position)
    nodes = rng.choice(d.idx_train, 48).tolist()
    labels = d.labels[nodes]
    pm = port.PortPCGNN(d.feat, d.graph, pool, params, rho=rho)
    if dup > 0.5:      # coarse scores: most distances tie
        pm.score_table = torch.from_numpy(np.round(rng.normal(size=(d.feat.shape[0], 2)), 1).astype(np.float32))
    with torch.no_grad():
        pm.loss(nodes, labels, train)
    score = pm.last["score_table"][:, 0].detach().numpy().astype(np.float32)
    sp, si = c_oracle.choose(d.graph, score, np.asarray(nodes), labels == 1, rho=rho, pool=pool, train=train)
    B = len(nodes)
    for r in range(3):
        for i in range(B):
            w = r * B + i
            assert sorted(pm.last["sel"][r][i]) == si[sp[w]:sp[w + 1]].tolist(), (r, i)


def test_row_partition_holds_the_same_rows():
    from pcgnn_b200.synth import make_graph

    g = make_graph("tiny", seed=8).graph
    lo, hi = 100, 350
    part = g.row_partition(lo, hi)
    assert part.partitioned and part.row_lo == lo and part.n_global == g.n_nodes and part.n_nodes == hi - lo
    for r in range(g.n_rel):
        for v in (lo, lo + 1, (lo + hi) // 2, hi - 1):
            assert np.array_equal(part.row(r, v), g.row(r, v))
        assert np.array_equal(part.degrees(r), g.degrees(r)[lo:hi])
    with pytest.raises(ValueError):
        part.row_partition(0, 10)


def _run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_bench_reference_arm_prints_the_contract_line():
    line = _run_bench("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--cpu-sample", "16")
    assert line["impl"] == "reference" and line["metric"].startswith("train target-nodes/sec")
    assert line["unit"] == "target-nodes/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "target-nodes/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_bench_reference_arm_other_ranks_stay_silent(monkeypatch):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_reference_arm_runs_the_partitioned_graph_on_batch_rows():
    """C5: the CPU arm materialises only the batch targets' rows (all the reference reads, layers.py:219)."""
    line = _run_bench("--impl", "reference", "--workload", "big", "--nodes-per-gpu", "20000", "--cpu-sample", "16",
                      "--steps", "1", "--warmup", "0")
    assert line["impl"] == "reference" and line["value"] > 0 and "unavailable" not in line
    assert "batch rows" in line["cpu_baseline"]["sample"]
