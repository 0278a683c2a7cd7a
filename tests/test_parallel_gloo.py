"""CPU, world_size 2 over gloo: the data-parallel plumbing (target sharding + flat gradient all-reduce)
reproduces the single-process gradient of the global batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcgnn_b200.parallel import GradAllReduce, global_mean_loss_scale, shard_batch, shard_range


def test_shard_range_covers_batch_without_overlap():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tiny_model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        x = torch.from_numpy(rng.normal(size=(n, 6)).astype(np.float32))
        y = torch.from_numpy(rng.integers(0, 2, n))
        model = _tiny_model(1)
        red = GradAllReduce(model.parameters()).attach()
        xs, ys = shard_batch(x, y, rank, world)
        red.zero()
        # local mean loss re-weighted so that the sum over ranks is the global-batch mean (unequal shards)
        loss = torch.nn.functional.cross_entropy(model(xs), ys) * global_mean_loss_scale(len(xs), n, world)
        loss.backward()
        red()
        torch.save(red.flat[:red.n].clone(), os.path.join(out_dir, f"g{rank}.pt"))      # (flat is padded to 4 floats)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 37])
def test_two_rank_gradient_equals_single_process(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.normal(size=(n, 6)).astype(np.float32))
    y = torch.from_numpy(rng.integers(0, 2, n))
    model = _tiny_model(1)
    torch.nn.functional.cross_entropy(model(x), y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"g{r}.pt"))
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)


def test_flat_buffer_views_alias_grads():
    model = _tiny_model(2)
    red = GradAllReduce(model.parameters()).attach()
    model(torch.ones(3, 6)).sum().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.equal(red.flat[:red.n], flat) and red.flat.numel() % 4 == 0
    red.zero()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in model.parameters())
