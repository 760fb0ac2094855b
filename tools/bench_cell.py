"""ConvLSTM cell microbench sweep (BASELINE.json configs[4]: C_hidden 32-256, HxW 64-512, T 8-32, fwd + BPTT at
1xB200).  A standalone ConvLSTM(C, C) layer (reference train/unet.py:39-60) driven through the same autograd
Function the model uses (unet_convlstm_b200.functional.ConvLSTMSeq): one timestep-persistent fused forward
launch, then BPTT (gate gradients, dgrad per step, one wgrad over the sequence).  x_t ~ N(0,1), PyTorch
default init, h_0 = c_0 = 0, loss = sum_t mean(h_t^2).

    python tools/bench_cell.py                 # the sweep
    python tools/bench_cell.py C HW T [B]      # one point

Prints ms for forward / forward+backward and TFLOP/s with the algorithmic FLOPs of SURVEY 8(d):
fwd 2*P*9*(Cin+Ch)*4Ch per step, fwd+bwd = 3x (the t=0 h-half skip is not subtracted).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unet_convlstm_b200 as pkg  # noqa: E402
from train.unet import ConvLSTMCell  # noqa: E402

dev = torch.device("cuda")


def run(C, HW, T, B, iters=3, warm=2):
    pkg.set_precision("bf16")
    torch.manual_seed(0)
    cell = ConvLSTMCell(C, C).to(dev)
    x = torch.randn(T, B, HW, HW, C, device=dev).to(torch.bfloat16).requires_grad_(True)

    def fwd():
        h_seq, _ = cell._seq(x, None, None)
        return h_seq

    def fwd_bwd():
        cell.zero_grad(set_to_none=True)
        x.grad = None
        h_seq = fwd()
        # d loss / d h_t = 2 h_t / numel for loss = sum_t mean(h_t^2): fed directly, so only our kernels run
        h_seq.backward((2.0 / h_seq[0].numel()) * h_seq.detach())

    def timeit(fn):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    with torch.no_grad():
        ms_f = timeit(fwd)
    ms_fb = timeit(fwd_bwd)
    fl = 2.0 * T * B * HW * HW * 9 * (2 * C) * (4 * C)
    print(f"C{C:<4d} {HW:>3d}x{HW:<3d} T{T:<3d} B{B:<4d} fwd {ms_f:8.3f} ms {fl / ms_f / 1e9:7.1f} TFLOP/s | "
          f"fwd+BPTT {ms_fb:8.3f} ms {3 * fl / ms_fb / 1e9:7.1f} TFLOP/s", flush=True)


def pick_batch(C, HW, T):
    # keep the saved state (h, c, gates, dz, x, dx ~ 30*C bytes per pixel-step) near 8 GB
    per_seq = T * HW * HW * C * 30
    return max(1, min(256, int(8e9 // per_seq)))


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    if a:
        C, HW, T = a[:3]
        run(C, HW, T, a[3] if len(a) > 3 else pick_batch(C, HW, T))
    else:
        for C in (32, 64, 128, 256):
            for HW in (64, 128, 256, 512):
                for T in (8, 32):
                    run(C, HW, T, pick_batch(C, HW, T))
