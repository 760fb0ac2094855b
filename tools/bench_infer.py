"""Inference throughput of the UNet-ConvLSTM: model.eval() under torch.no_grad() -- BatchNorm
folded into the conv epilogues, the ConvLSTM layers as timestep-persistent fused kernels -- against the same model with
the unfused eval path (autograd enabled).  python tools/bench_infer.py [B T size base_ch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unet_convlstm_b200 as pkg  # noqa: E402
from train.unet import TemporalUNetDualView  # noqa: E402

B, T, S, b = ([int(v) for v in sys.argv[1:5]] + [256, 20, 64, 64][len(sys.argv) - 1:])[:4]
pkg.set_precision("bf16")
torch.manual_seed(0)
m = TemporalUNetDualView(base_ch=b, use_skip_lstm=True).cuda().eval()
x = torch.rand(B, T, 2, S, S, device="cuda")


def timeit(fn, iters=3):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def folded():
    with torch.no_grad():
        m(x)


def streaming():
    # online use: frames arrive one at a time, the (h, c) state of the temporal cell carries over (unet.py:185)
    with torch.no_grad():
        st = None
        for t in range(T):
            _, st = m(x[:, t:t + 1], st)


ms_f = timeit(folded)
ms_u = timeit(lambda: m(x))
ms_s = timeit(streaming, 1)
print(f"inference B{B} T{T} {S}x{S} base_ch{b}: folded-BN no_grad {ms_f:8.2f} ms = {B / ms_f * 1e3:8.1f} seq/s | "
      f"unfused eval {ms_u:8.2f} ms = {B / ms_u * 1e3:8.1f} seq/s | frame-by-frame with state carry {ms_s:8.2f} ms", flush=True)
