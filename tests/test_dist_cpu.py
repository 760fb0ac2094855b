"""Data-parallel host logic on CPU: world_size-2 gloo run of unet_convlstm_b200.dist.GradReducer against
the single-process emulation (per-shard forward/backward, gradients averaged) -- SURVEY.md section 8e."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(3)
    return nn.Sequential(nn.Conv2d(2, 6, 3, padding=1), nn.BatchNorm2d(6), nn.ReLU(), nn.Conv2d(6, 4, 3, padding=1),
                         nn.Flatten(), nn.Linear(4 * 8 * 8, 5))


def _data():
    g = torch.Generator().manual_seed(9)
    return torch.randn(8, 2, 8, 8, generator=g), torch.randn(8, 5, generator=g)


def _worker(rank, world, port, bucket_bytes, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet_convlstm_b200.dist import GradReducer
    m = _model()
    red = GradReducer(m.parameters(), bucket_bytes=bucket_bytes)
    x, y = _data()
    n = x.shape[0] // world
    xs, ys = x[rank * n:(rank + 1) * n], y[rank * n:(rank + 1) * n]
    for it in range(2):  # second iteration: gradients already live in the buckets
        for p in m.parameters():
            p.grad = None if it == 0 else p.grad.zero_()
        ((m(xs) - ys) ** 2).mean().backward()
        red.finish()
    torch.save([p.grad.clone() for p in m.parameters()], os.path.join(out_dir, f"g{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [64, 1 << 20])
def test_grad_reducer_matches_single_process_emulation(tmp_path, bucket_bytes):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), bucket_bytes, str(tmp_path)), nprocs=world, join=True)
    x, y = _data()
    n = x.shape[0] // world
    ref = None
    for r in range(world):
        m = _model()
        ((m(x[r * n:(r + 1) * n]) - y[r * n:(r + 1) * n]) ** 2).mean().backward()
        g = [p.grad for p in m.parameters()]
        ref = g if ref is None else [a + b for a, b in zip(ref, g)]
    ref = [g / world for g in ref]
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"g{r}.pt"))
        for a, b in zip(got, ref):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_grad_reducer_single_process_is_identity():
    from unet_convlstm_b200.dist import GradReducer
    m = _model()
    red = GradReducer(m.parameters(), bucket_bytes=128)
    x, y = _data()
    ((m(x) - y) ** 2).mean().backward()
    red.finish()
    m2 = _model()
    ((m2(x) - y) ** 2).mean().backward()
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.equal(a.grad, b.grad)


def test_make_loaders_shard_the_dataset_per_rank():
    """loop.make_loaders: every rank builds the same split; the per-rank samplers partition the training part
    (per-rank DistributedSampler)."""
    from torch.utils.data import TensorDataset
    from unet_convlstm_b200.loop import make_loaders
    ds = TensorDataset(torch.arange(40).float().view(40, 1))
    seen = []
    for rank in range(4):
        tr, va = make_loaders(ds, batch_size=3, rank=rank, world=4)
        tr.sampler.set_epoch(0)
        seen.append(sorted(int(v) for (b,) in tr for v in b.view(-1)))
        assert sum(len(b[0]) for b in va) == 2                      # 8 validation items over 4 ranks
    flat = sorted(v for s in seen for v in s)
    assert len(flat) == 32 and len(set(flat)) == 32                 # 80 % of 40, each item on exactly one rank
    tr1, va1 = make_loaders(ds, batch_size=3, rank=0, world=1)
    assert sorted(int(v) for (b,) in tr1 for v in b.view(-1)) == flat   # same split as the single-process loaders
