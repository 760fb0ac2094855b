"""torchrun --nproc-per-node 2 tools/check_dp_background.py : data-parallel gradients with the weight gradients on the
background stream (+ artificial delay) against the same step with everything on one stream."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_convlstm_b200 as pkg  # noqa: E402
from train.unet import TemporalUNetDualView  # noqa: E402
from unet_convlstm_b200 import ops  # noqa: E402
from unet_convlstm_b200.dist import GradReducer  # noqa: E402
from unet_convlstm_b200.loss import compute_loss  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg.set_precision("bf16")
torch.manual_seed(0)
model = TemporalUNetDualView(base_ch=16, use_skip_lstm=True).cuda()
rng = np.random.default_rng(rank)
x = torch.from_numpy((rng.random((4, 6, 2, 32, 32)) * 2).astype(np.float32)).cuda()
y = torch.from_numpy(np.clip(rng.standard_normal((4, 6, 1, 32, 32)), -1, 1).astype(np.float32)).cuda()
m = (x[:, :, 0:1] > 1.1).float()
reducer = GradReducer(model.parameters(), bucket_bytes=1 << 20)
res = {}
for mode in (False, True, True):
    ops.WGRAD_STREAM = mode
    ops._BG_DEBUG_DELAY = 30_000_000 if mode else 0
    ops.LSTM_WGRAD_CHUNK = 2
    model.zero_grad(set_to_none=True)
    out, _ = model(x)
    compute_loss(torch.stack(out, dim=1), y, m).backward()
    reducer.finish()
    torch.cuda.synchronize()
    res[mode] = [p.grad.detach().float().clone() for p in model.parameters()]
worst = 0.0
for a, b in zip(res[False], res[True]):
    worst = max(worst, float((a - b).norm() / b.norm().clamp_min(1e-20)))
t = torch.tensor([worst], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: worst relative gradient difference background vs in-line: {float(t):.3e}", "OK" if float(t) < 2e-3 else "MISMATCH")
dist.destroy_process_group()
sys.exit(0 if float(t) < 2e-3 else 1)
