"""N-rank data-parallel parity on real GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_dp_parity.py

Every rank runs the training step of bench.py (model replica, its shard of the batch, GradReducer with the NCCL
all-reduce overlapped with backward, background weight-gradient stream, fused AdamW step) for two steps; rank 0 then
replays the same two steps as the single-process EMULATION the reference semantics define -- per-shard forward/backward
on an identical replica (BatchNorm statistics stay per shard: the reference has no SyncBN), gradients averaged over the
shards, same optimizer -- and compares averaged gradients (step 1) and updated parameters (step 2).
The tensor-core weight gradients combine their reduction splits with fp32 atomics, so two runs differ in the last
bits: the tolerance is 2e-3 tensor-relative (L2)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_convlstm_b200 as pkg  # noqa: E402
from train.unet import TemporalUNetDualView  # noqa: E402
from unet_convlstm_b200.dist import GradReducer  # noqa: E402
from unet_convlstm_b200.loss import compute_loss  # noqa: E402
from unet_convlstm_b200.optim import AdamW  # noqa: E402

TOL = 2e-3


def shard(r):
    rng = np.random.default_rng(100 + r)
    x = torch.from_numpy((rng.random((4, 5, 2, 32, 32)) * 2).astype(np.float32)).cuda()
    y = torch.from_numpy(np.clip(rng.standard_normal((4, 5, 1, 32, 32)), -1, 1).astype(np.float32)).cuda()
    return x, y, (x[:, :, 0:1] > 1.1).float()


def make_model():
    torch.manual_seed(0)
    return TemporalUNetDualView(base_ch=16, use_skip_lstm=True).cuda().train()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg.set_precision("bf16")
    model = make_model()
    opt = AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    reducer = GradReducer(model.parameters(), bucket_bytes=1 << 20)
    x, y, m = shard(rank)
    grads1 = None
    for step in range(2):
        opt.zero_grad(set_to_none=True)
        out, _ = model(x)
        compute_loss(torch.stack(out, dim=1), y, m).backward()
        reducer.finish()
        if step == 0:
            grads1 = [p.grad.detach().float().clone() for p in model.parameters()]
        opt.step(clip_max_norm=1.0)
    torch.cuda.synchronize()
    params2 = [p.detach().float().clone() for p in model.parameters()]
    # every rank must hold the same replica after the update
    for p in params2:
        lo, hi = p.clone(), p.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "replicas diverged"
    ok = True
    if rank == 0:
        reducer.remove()
        ref = make_model()
        ropt = AdamW(ref.parameters(), lr=1e-3, weight_decay=1e-4)
        names = [n for n, _ in ref.named_parameters()]
        shards = [shard(r) for r in range(world)]
        worst_g = worst_p = 0.0
        zero_grad = set()   # conv biases feeding a train-mode BatchNorm: mathematically zero gradient, Adam turns its
        # rounding noise into +-lr steps of random sign, so neither gradient nor parameter is comparable
        for step in range(2):
            acc = None
            bufs = {k: v.clone() for k, v in ref.state_dict().items() if "running" in k or "num_batches" in k}
            for r in range(world):
                # each shard sees the replica's BatchNorm buffers as they were at the start of the step (rank-local
                # buffers evolve independently; rank 0's are the ones compared below, so restore + keep shard 0's)
                ref.load_state_dict(bufs, strict=False)
                ropt.zero_grad(set_to_none=True)
                xs, ys, ms = shards[r]
                out, _ = ref(xs)
                compute_loss(torch.stack(out, dim=1), ys, ms).backward()
                g = [p.grad.detach().float().clone() for p in ref.parameters()]
                acc = g if acc is None else [a + b for a, b in zip(acc, g)]
                if r == 0:
                    bufs0 = {k: v.clone() for k, v in ref.state_dict().items() if "running" in k or "num_batches" in k}
            ref.load_state_dict(bufs0, strict=False)
            avg = [a / world for a in acc]
            if step == 0:
                for n, a, b in zip(names, grads1, avg):
                    if float(b.norm()) < 1e-6:
                        zero_grad.add(n)
                        continue
                    worst_g = max(worst_g, float((a - b).norm() / b.norm()))
            for p, gavg in zip(ref.parameters(), avg):
                p.grad = gavg.to(p.dtype)
            ropt.step(clip_max_norm=1.0)
        torch.cuda.synchronize()
        for n, a, p in zip(names, params2, ref.parameters()):
            if n in zero_grad:
                continue
            worst_p = max(worst_p, float((a - p.detach().float()).norm() / p.detach().float().norm().clamp_min(1e-20)))
        ok = worst_g < TOL and worst_p < TOL
        print(f"world {world}: averaged gradients vs per-shard emulation worst l2-rel {worst_g:.3e}, parameters after 2 steps "
              f"{worst_p:.3e} (tol {TOL:g})", "OK" if ok else "MISMATCH", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
