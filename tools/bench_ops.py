"""Micro-benchmarks of single kernels at the shapes of BASELINE.json configs[1] (developer aid, GPU box).

    python tools/bench_ops.py wgrad 64 64 64 [T B]     # Nz Csrc HW
    python tools/bench_ops.py conv 64 64 64            # K N HW
    python tools/bench_ops.py all
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unet_convlstm_b200 import ops  # noqa: E402

dev = torch.device("cuda")
bf = torch.bfloat16


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_wgrad(Nz, C, HW, T=20, B=256, ks=3):
    dz = torch.randn(T, B, HW, HW, Nz, device=dev).to(bf)
    src = torch.randn(T, B, HW, HW, C, device=dev).to(bf)
    dw = torch.zeros(ks * ks, Nz, C, device=dev)
    ms = timeit(lambda: ops.conv_wgrad(dz, src, ks, dw, 0))
    fl = 2.0 * T * B * HW * HW * ks * ks * C * Nz
    print(f"wgrad Nz{Nz} C{C} {HW}x{HW} T{T} B{B}: {ms:8.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


def bench_conv(K, N, HW, T=20, B=256, ks=3):
    x = torch.randn(T, B, HW, HW, K, device=dev).to(bf)
    wp = (torch.randn(ks * ks, N, K, device=dev) * 0.05).to(bf)
    out = torch.empty(T, B, HW, HW, N, device=dev, dtype=bf)
    ms = timeit(lambda: ops.conv_fwd(x, None, wp, None, ks, out))
    fl = 2.0 * T * B * HW * HW * ks * ks * K * N
    print(f"conv K{K} N{N} {HW}x{HW} T{T} B{B}: {ms:8.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


def bench_convbn(K, N, HW, T=20, B=256, ks=3):
    x = torch.randn(T, B, HW, HW, K, device=dev).to(bf)
    wp = (torch.randn(ks * ks, N, K, device=dev) * 0.05).to(bf)
    bias = torch.randn(N, device=dev)
    out = torch.empty(T, B, HW, HW, N, device=dev, dtype=bf)
    ws = torch.empty(2, T, N, device=dev, dtype=torch.float64)
    fl = 2.0 * T * B * HW * HW * ks * ks * K * N
    ms = timeit(lambda: ops.conv_fwd(x, None, wp, bias, ks, out))
    print(f"conv        K{K} N{N} {HW}x{HW} T{T} B{B}: {ms:8.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)
    ms = timeit(lambda: ops.conv_fwd(x, None, wp, bias, ks, out, bn_ws=ws))
    print(f"conv+bnstat K{K} N{N} {HW}x{HW} T{T} B{B}: {ms:8.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


def bench_convT(Cin, Cout, HW, T=20, B=256):
    x = torch.randn(T, B, HW, HW, Cin, device=dev).to(bf)
    w = torch.randn(Cin, Cout, 2, 2, device=dev) / Cin ** 0.5
    bias = torch.randn(Cout, device=dev)
    wf, _ = ops.pack_convT_weight(w, bf)
    fl = 2.0 * T * B * HW * HW * Cin * 4 * Cout
    by = T * B * HW * HW * (Cin + 4 * Cout) * 2
    ms = timeit(lambda: ops.convT2x2_fwd(x, wf, bias, Cout, 2 * HW, 2 * HW))
    print(f"convT Cin{Cin} Cout{Cout} {HW}x{HW} fused={int(ops.CONVT_FUSED)}: {ms:8.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s "
          f"{by / ms / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    if a[0] == "wgrad":
        bench_wgrad(*[int(v) for v in a[1:]])
    elif a[0] == "conv":
        bench_conv(*[int(v) for v in a[1:]])
    elif a[0] == "convT":
        bench_convT(*[int(v) for v in a[1:]])
    elif a[0] == "convbn":
        bench_convbn(*[int(v) for v in a[1:]])
    else:
        for Nz, C, HW in [(64, 64, 64), (64, 16, 64), (128, 64, 32), (128, 128, 32), (256, 128, 16), (256, 256, 16),
                          (512, 512, 8), (1024, 1024, 4)]:
            bench_wgrad(Nz, C, HW)
        for Nz, C, HW, T in [(4096, 1024, 4, 20), (2048, 512, 8, 20), (1024, 256, 16, 20)]:
            bench_wgrad(Nz, C, HW, T)
        for K, N, HW in [(16, 64, 64), (64, 64, 64), (128, 64, 64), (64, 128, 64), (128, 128, 32), (256, 128, 32),
                         (256, 256, 16), (512, 512, 8), (1024, 1024, 4)]:
            bench_conv(K, N, HW)
