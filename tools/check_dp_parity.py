"""N-rank data-parallel parity on real GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_dp_parity.py

Every rank runs the training step of bench.py (model replica, its shard of the batch, GradReducer with the NCCL
all-reduce overlapped with backward, background weight-gradient stream, fused AdamW step) for two steps; rank 0 then
computes the single-process EMULATION the reference semantics define -- per-shard forward/backward on an identical
replica (BatchNorm statistics stay per shard: the reference has no SyncBN), gradients averaged over the shards -- and
compares the averaged gradients of both steps (the second from the parameters after the first update, with the
gradients already living in the reducer's buckets); all ranks must hold bit-identical parameters at the end.
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
    grads, params_after = [], []
    for step in range(2):
        opt.zero_grad(set_to_none=True)
        out, _ = model(x)
        compute_loss(torch.stack(out, dim=1), y, m).backward()
        reducer.finish()
        grads.append([p.grad.detach().float().clone() for p in model.parameters()])
        opt.step(clip_max_norm=1.0)
        params_after.append([p.detach().clone() for p in model.parameters()])
    torch.cuda.synchronize()
    # every rank must hold the same replica after the updates (identical averaged gradients -> identical parameters)
    for p in params_after[-1]:
        lo, hi = p.clone(), p.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "replicas diverged"
    ok = True
    if rank == 0:
        reducer.remove()
        ref = make_model()
        names = [n for n, _ in ref.named_parameters()]
        shards = [shard(r) for r in range(world)]
        worst = [0.0, 0.0]
        for step in range(2):
            # step 1 (the second one) starts from the parameters the data-parallel run had after ITS first update:
            # Adam's first update is lr * sign(g), so replaying the optimizer would compare rounding-noise signs
            if step == 1:
                with torch.no_grad():
                    for p, q in zip(ref.parameters(), params_after[0]):
                        p.copy_(q)
            acc = None
            for r in range(world):
                ref.zero_grad(set_to_none=True)
                xs, ys, ms = shards[r]
                out, _ = ref(xs)     # train-mode BatchNorm: batch statistics only, the running buffers do not matter
                compute_loss(torch.stack(out, dim=1), ys, ms).backward()
                g = [p.grad.detach().float().clone() for p in ref.parameters()]
                acc = g if acc is None else [a + b for a, b in zip(acc, g)]
            avg = [a / world for a in acc]
            for n, a, b in zip(names, grads[step], avg):
                if float(b.norm()) < 1e-6:   # conv biases feeding a train-mode BatchNorm: mathematically zero gradient
                    continue
                worst[step] = max(worst[step], float((a - b).norm() / b.norm()))
        torch.cuda.synchronize()
        ok = max(worst) < TOL
        print(f"world {world}: averaged gradients vs per-shard emulation, worst l2-rel: step 1 {worst[0]:.3e}, step 2 "
              f"{worst[1]:.3e} (tol {TOL:g})", "OK" if ok else "MISMATCH", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
