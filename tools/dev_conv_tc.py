"""Developer bring-up check for the tcgen05 conv kernel (run on the GPU box via gpurun).

Compares b200_conv_tc_fwd / b200_convlstm_cell_fwd_tc with torch fp32 convolutions on the same
bf16-rounded inputs.  Not part of the product path; the pytest parity suite supersedes it.
"""
import ctypes
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "unet_convlstm_b200", "libb200convlstm.so"))
lib.b200_last_error.restype = ctypes.c_char_p
vp, ci, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
lib.b200_conv_tc_fwd.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, vp, vp, ci, ci, vp, ll, ci, vp, ll, ci, ci, ci, vp]
lib.b200_convlstm_cell_fwd_tc.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, ci, vp]

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")


def ptr(t):
    return None if t is None else t.data_ptr()


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: rc={rc} {lib.b200_last_error().decode()}")


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def pack_w(w):  # OIHW fp32 -> [k*k][N][C] bf16
    N, C, k, _ = w.shape
    return w.permute(2, 3, 0, 1).reshape(k * k, N, C).contiguous().to(torch.bfloat16)


def lstm_row_perm(Ch):
    cht = 64 if Ch % 64 == 0 else (32 if Ch % 32 == 0 else 16)
    idx = []
    for nt in range(Ch // cht):
        for g in range(4):
            for j in range(cht):
                idx.append(g * Ch + nt * cht + j)
    return torch.tensor(idx, device=dev)


def test_conv(T, B, H, W, C0, C1, N, k=3, out_fp32=False, relu=False, split=None):
    g = torch.Generator(device=dev).manual_seed(1)
    x0 = torch.randn(T, B, H, W, C0, device=dev, generator=g).to(torch.bfloat16)
    x1 = torch.randn(T, B, H, W, C1, device=dev, generator=g).to(torch.bfloat16) if C1 else None
    w = torch.randn(N, C0 + C1, k, k, device=dev, generator=g) / ((C0 + C1) * k * k) ** 0.5
    bias = torch.randn(N, device=dev, generator=g)
    wp = pack_w(w)
    split = N if split is None else split
    odt = torch.float32 if out_fp32 else torch.bfloat16
    d0 = torch.full((T, B, H, W, split), float("nan"), device=dev, dtype=odt)
    d1 = torch.full((T, B, H, W, N - split), float("nan"), device=dev, dtype=odt) if split < N else None
    rc = lib.b200_conv_tc_fwd(ptr(x0), C0, ptr(x1), C1, T, B, H, W, ptr(wp), ptr(bias), N, k, ptr(d0), split,
                              split, ptr(d1), N - split, int(out_fp32), int(relu), 0,
                              torch.cuda.current_stream().cuda_stream)
    check(rc, "conv_tc_fwd")
    torch.cuda.synchronize()
    xin = x0 if x1 is None else torch.cat([x0, x1], dim=-1)
    xin = xin.float().reshape(T * B, H, W, C0 + C1).permute(0, 3, 1, 2)
    ref = F.conv2d(xin, wp.float().reshape(k, k, N, C0 + C1).permute(2, 3, 0, 1), bias, padding=k // 2)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1).reshape(T, B, H, W, N)
    out = d0 if d1 is None else torch.cat([d0, d1], dim=-1)
    e = relerr(out.float(), ref)
    tol = 2e-5 if out_fp32 else 1e-2
    ok = e < tol and not torch.isnan(out.float()).any().item()
    print(f"conv T={T} B={B} H={H} W={W} C0={C0} C1={C1} N={N} k={k} fp32out={out_fp32} split={split}: "
          f"relerr={e:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def test_lstm(B, H, W, Cin, Ch, with_state=True):
    g = torch.Generator(device=dev).manual_seed(2)
    x = torch.randn(B, H, W, Cin, device=dev, generator=g).to(torch.bfloat16)
    h = torch.randn(B, H, W, Ch, device=dev, generator=g).to(torch.bfloat16) if with_state else None
    c = torch.randn(B, H, W, Ch, device=dev, generator=g) if with_state else None
    w = torch.randn(4 * Ch, Cin + Ch, 3, 3, device=dev, generator=g) / ((Cin + Ch) * 9) ** 0.5
    bias = torch.randn(4 * Ch, device=dev, generator=g) * 0.1
    perm = lstm_row_perm(Ch)
    wp = pack_w(w[perm])
    bp = bias[perm].contiguous()
    c_next = torch.full((B, H, W, Ch), float("nan"), device=dev)
    h_next = torch.full((B, H, W, Ch), float("nan"), device=dev, dtype=torch.bfloat16)
    gates = torch.full((B, H, W, 4, Ch), float("nan"), device=dev, dtype=torch.bfloat16)
    rc = lib.b200_convlstm_cell_fwd_tc(ptr(x), Cin, ptr(h), Ch, B, H, W, ptr(wp), ptr(bp), ptr(c), ptr(c_next),
                                       ptr(h_next), ptr(gates), 3, torch.cuda.current_stream().cuda_stream)
    check(rc, "convlstm_cell_fwd_tc")
    torch.cuda.synchronize()
    wq = w.to(torch.bfloat16).float()
    hh = h if h is not None else torch.zeros(B, H, W, Ch, device=dev, dtype=torch.bfloat16)
    cc = c if c is not None else torch.zeros(B, H, W, Ch, device=dev)
    xin = torch.cat([x, hh], dim=-1).float().permute(0, 3, 1, 2)
    z = F.conv2d(xin, wq, bias, padding=1)
    i, f, gg, o = torch.chunk(z, 4, dim=1)
    i, f, gg, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(gg), torch.sigmoid(o)
    cn = f * cc.permute(0, 3, 1, 2) + i * gg
    hn = o * torch.tanh(cn)
    e_c = relerr(c_next.permute(0, 3, 1, 2), cn)
    e_h = relerr(h_next.float().permute(0, 3, 1, 2), hn)
    gref = torch.stack([i, f, gg, o], dim=1).permute(0, 3, 4, 1, 2)  # B,H,W,4,Ch
    e_g = relerr(gates.float(), gref)
    ok = e_c < 5e-3 and e_h < 1e-2 and e_g < 1e-2
    print(f"lstm B={B} H={H} W={W} Cin={Cin} Ch={Ch} state={with_state}: c={e_c:.3e} h={e_h:.3e} g={e_g:.3e} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    return ok


lib.b200_wgrad_tc.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci, vp, ll, ci, vp]


def test_wgrad(T, B, H, W, Nz, C0, C1, k=3):
    g = torch.Generator(device=dev).manual_seed(3)
    dz = torch.randn(T, B, H, W, Nz, device=dev, generator=g).to(torch.bfloat16)
    s0 = torch.randn(T, B, H, W, C0, device=dev, generator=g).to(torch.bfloat16)
    s1 = torch.randn(T, B, H, W, C1, device=dev, generator=g).to(torch.bfloat16) if C1 else None
    Ct = C0 + C1
    dw = torch.zeros(k * k, Nz, Ct, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.b200_wgrad_tc(ptr(dz), Nz, ptr(s0), C0, T, B, H, W, k, ptr(dw), Ct, 0, st), "wgrad0")
    if s1 is not None:
        check(lib.b200_wgrad_tc(ptr(dz), Nz, ptr(s1), C1, T, B, H, W, k, ptr(dw), Ct, C0, st), "wgrad1")
    torch.cuda.synchronize()
    xin = s0 if s1 is None else torch.cat([s0, s1], dim=-1)
    xin = xin.float().reshape(T * B, H, W, Ct).permute(0, 3, 1, 2).contiguous()
    dzz = dz.float().reshape(T * B, H, W, Nz).permute(0, 3, 1, 2).contiguous()
    ref = torch.nn.grad.conv2d_weight(xin, (Nz, Ct, k, k), dzz, padding=k // 2)  # OIHW
    ref = ref.permute(2, 3, 0, 1).reshape(k * k, Nz, Ct)
    e = relerr(dw, ref)
    ok = e < 1e-4
    print(f"wgrad T={T} B={B} H={H} W={W} Nz={Nz} C0={C0} C1={C1} k={k}: relerr={e:.3e} {'OK' if ok else 'FAIL'}",
          flush=True)
    return ok


def bench_wgrad(T, B, H, W, Nz, C, iters=5):
    dz = torch.randn(T, B, H, W, Nz, device=dev).to(torch.bfloat16)
    s0 = torch.randn(T, B, H, W, C, device=dev).to(torch.bfloat16)
    dw = torch.zeros(9, Nz, C, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        lib.b200_wgrad_tc(ptr(dz), Nz, ptr(s0), C, T, B, H, W, 3, ptr(dw), C, 0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.b200_wgrad_tc(ptr(dz), Nz, ptr(s0), C, T, B, H, W, 3, ptr(dw), C, 0, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * T * B * H * W * 9 * C * Nz
    print(f"bench wgrad T={T} B={B} H={H} W={W} Nz={Nz} C={C}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


def bench_lstm(B, H, W, C, iters=20):
    x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    h = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    c = torch.randn(B, H, W, C, device=dev)
    wp = torch.randn(9, 4 * C, 2 * C, device=dev).to(torch.bfloat16) * 0.01
    bp = torch.zeros(4 * C, device=dev)
    c_next = torch.empty_like(c)
    h_next = torch.empty_like(h)
    gates = torch.empty(B, H, W, 4, C, device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        lib.b200_convlstm_cell_fwd_tc(ptr(x), C, ptr(h), C, B, H, W, ptr(wp), ptr(bp), ptr(c), ptr(c_next),
                                      ptr(h_next), ptr(gates), 3, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.b200_convlstm_cell_fwd_tc(ptr(x), C, ptr(h), C, B, H, W, ptr(wp), ptr(bp), ptr(c), ptr(c_next),
                                      ptr(h_next), ptr(gates), 3, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * B * H * W * 9 * 2 * C * 4 * C
    print(f"bench lstm B={B} H={H} W={W} C={C}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    ok = True
    ok &= test_conv(1, 2, 8, 16, 64, 0, 64)
    ok &= test_conv(1, 1, 16, 128, 64, 0, 128, out_fp32=True)
    ok &= test_conv(2, 4, 4, 4, 128, 0, 256)
    ok &= test_conv(2, 3, 8, 8, 64, 64, 256, relu=True)
    ok &= test_conv(1, 2, 32, 32, 64, 128, 192, split=64, out_fp32=True)
    ok &= test_conv(1, 2, 16, 16, 32, 32, 96)
    ok &= test_conv(1, 2, 16, 16, 16, 0, 32, k=1)
    ok &= test_conv(1, 1, 12, 16, 64, 0, 64)
    ok &= test_conv(1, 2, 64, 256, 64, 0, 64)
    ok &= test_lstm(2, 16, 16, 64, 64)
    ok &= test_lstm(2, 16, 16, 64, 64, with_state=False)
    ok &= test_lstm(4, 8, 8, 128, 128)
    ok &= test_lstm(2, 32, 32, 32, 32)
    ok &= test_lstm(2, 16, 16, 32, 16)
    ok &= test_wgrad(1, 2, 8, 8, 128, 64, 0)
    ok &= test_wgrad(2, 4, 4, 4, 256, 128, 128)
    ok &= test_wgrad(3, 2, 16, 16, 64, 64, 0)
    ok &= test_wgrad(2, 2, 32, 32, 128, 32, 96)
    ok &= test_wgrad(1, 1, 12, 16, 64, 16, 0)
    ok &= test_wgrad(2, 2, 8, 128, 320, 320, 0, k=1)
    ok &= test_wgrad(4, 8, 64, 64, 64, 64, 0)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    bench_wgrad(20, 256, 4, 4, 4096, 1024)
    bench_wgrad(20, 256, 16, 16, 1024, 256)
    bench_wgrad(4, 256, 64, 64, 64, 64)
    if ok:
        bench_lstm(256, 4, 4, 1024)
        bench_lstm(256, 8, 8, 512)
        bench_lstm(256, 16, 16, 256)
        bench_lstm(64, 64, 64, 64)
    sys.exit(0 if ok else 1)
