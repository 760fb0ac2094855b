"""GPU parity: the CUDA path (through train/unet.py -> C ABI) against the golden fixtures produced by
the reference itself (tests/golden/make_golden.py) and against the numpy oracle on seeded inputs.

Tolerances (north_star): 1e-5 relative in the fp32 check mode, 2e-2 relative in bf16 mode.
"Relative" is tensor-relative: max|a-b| / max|b| (fp32 mode) -- and, in bf16 mode, the L2 form
||a-b||_2 / ||b||_2 <= 2e-2 with the max form bounded by 5x that, because a single ReLU / max-pool
decision flipped by bf16 rounding moves one element by O(1) without being a numerical error.
Gradients that are mathematically zero (conv biases feeding a train-mode BatchNorm) are compared
absolutely.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "bf16": 2e-2}
# how far above the reference's own noise floor (fixture keys e32.* / e16.*, see make_golden.py) a
# result may sit: train-mode BatchNorm backward at small batch x spatial sizes is ill-conditioned and
# the reference itself misses 1e-5 (fp32) / 2e-2 (bf16 autocast) there
FLOOR_FACTOR = {"fp32": 4.0, "bf16": 4.0}


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def rel2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30)


def close(a, b, mode, floor=None, problem_floor=None):
    """The parity criterion described in the module docstring.
    floor: [max-rel, l2-rel] error of the reference itself at this precision for this tensor;
    problem_floor: the largest such error over all tensors of the same problem.  bf16 errors of a
    BatchNorm/ReLU/max-pool net are spiky -- one flipped ReLU decision moves a whole gradient entry -- so
    a tensor may also sit at 2x the noise level of the problem as a whole."""
    fmax, fl2 = (0.0, 0.0) if floor is None else (float(floor[0]), float(floor[1]))
    pmax, pl2 = (0.0, 0.0) if problem_floor is None else (float(problem_floor[0]), float(problem_floor[1]))
    k = FLOOR_FACTOR[mode]
    if mode == "fp32":
        r = rel(a, b)
        return r < max(TOL["fp32"], k * fmax), (r, fmax)
    r2, r = rel2(a, b), rel(a, b)
    ok = r2 < max(TOL["bf16"], k * fl2, 2 * pl2) and r < max(5 * TOL["bf16"], 1.5 * k * fmax, 3 * pmax)
    return ok, (r2, r, fl2, fmax, pl2, pmax)


def expect(a, b, mode, floor=None, what="", problem_floor=None):
    ok, err = close(a, b, mode, floor, problem_floor)
    assert ok, (what, mode, err)


def _problem_floor(z, mode):
    pre = "e32." if mode == "fp32" else "e16."
    worst = np.zeros(2)
    for k in z.files:
        if k.startswith(pre) and "num_batches" not in k and np.abs(z[k[4:]]).max() > 1e-6:
            worst = np.maximum(worst, z[k])
    return worst


def expect_key(a, z, key, mode):
    """Compares `a` with fixture entry z[key] using the fixture's noise floors."""
    fk = ("e32." if mode == "fp32" else "e16.") + key
    expect(a, z[key], mode, z[fk] if fk in z.files else None, key, _problem_floor(z, mode))


def _np(t):
    return t.detach().float().cpu().numpy()


@pytest.fixture(params=["fp32", "bf16"])
def mode(request):
    import unet_convlstm_b200 as pkg
    old = pkg.get_precision()
    pkg.set_precision(request.param)
    yield request.param
    pkg.set_precision(old)


def _load_sd(module, z):
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
    module.load_state_dict(sd, strict=True)
    return module.cuda()


def _cuda(a, grad=False):
    t = torch.from_numpy(np.asarray(a, dtype=np.float32)).cuda()
    return t.requires_grad_(grad)


def _check_grads(module, z, mode, zero_scale):
    params = dict(module.named_parameters())
    for k in z.files:
        if not k.startswith("g."):
            continue
        ref = z[k].astype(np.float64)
        g = params[k[2:]].grad
        assert g is not None, k
        if np.abs(ref).max() < 1e-7 * zero_scale:
            # mathematically zero (conv bias feeding a train-mode BatchNorm): absolute comparison
            assert np.abs(_np(g)).max() < 10 * TOL[mode] * zero_scale, k
        else:
            expect_key(_np(g), z, k, mode)


def _check_buffers(module, z, mode):
    bufs = dict(module.named_buffers())
    for k in z.files:
        if k.startswith("after."):
            if "num_batches_tracked" in k:
                assert int(bufs[k[6:]]) == int(z[k]), k
            else:
                expect_key(_np(bufs[k[6:]].float()), z, k, mode)


@pytest.mark.parametrize("name", ["convlstm_c8_l1_zero.npz", "convlstm_c6_12_l2_state.npz", "convlstm_c16_l1_state.npz"])
def test_convlstm_golden(golden_dir, name, mode):
    from train.unet import ConvLSTM
    z = np.load(os.path.join(golden_dir, name))
    cin, ch, L, B, T, H, W, with_state = [int(v) for v in z["meta"]]
    m = _load_sd(ConvLSTM(cin, ch, num_layers=L), z)
    xs = [_cuda(z["x"][t], True) for t in range(T)]
    state = None
    if with_state:
        state = [(_cuda(z[f"h0{l}"], True), _cuda(z[f"c0{l}"], True)) for l in range(L)]
    out, new_state = m(xs, state)
    assert isinstance(out, list) and len(out) == T and tuple(out[0].shape) == (B, ch, H, W)
    expect_key(_np(torch.stack(out)), z, "out", mode)
    for l in range(L):
        expect_key(_np(new_state[l][0]), z, f"hT{l}", mode)
        expect_key(_np(new_state[l][1]), z, f"cT{l}", mode)
    loss = sum((o * _cuda(z["dout"][t])).sum() for t, o in enumerate(out))
    loss = loss + (new_state[-1][0] * _cuda(z["dh_last"])).sum() + (new_state[-1][1] * _cuda(z["dc_last"])).sum()
    loss.backward()
    expect_key(np.stack([_np(x.grad) for x in xs]), z, "dx", mode)
    _check_grads(m, z, mode, 1.0)
    if with_state:
        for l in range(L):
            expect_key(_np(state[l][0].grad), z, f"dh0{l}", mode)
            expect_key(_np(state[l][1].grad), z, f"dc0{l}", mode)


@pytest.mark.parametrize("name,kind,cin,cout", [("double_3_8.npz", "double", 3, 8), ("down_8_16.npz", "down", 8, 16),
                                                ("up_16_8.npz", "up", 16, 8), ("up_16_8_pad.npz", "up", 16, 8)])
def test_blocks_golden(golden_dir, name, kind, cin, cout, mode):
    from train.unet import DoubleConv, Down, Up
    z = np.load(os.path.join(golden_dir, name))
    m = _load_sd({"double": DoubleConv, "down": Down, "up": Up}[kind](cin, cout), z)
    args = [_cuda(z[f"x{i}"], True) for i in range(2 if kind == "up" else 1)]
    m.train()
    y = m(*args)
    expect_key(_np(y), z, "y_train", mode)
    (y * _cuda(z["dy"])).sum().backward()
    for i, a in enumerate(args):
        expect_key(_np(a.grad), z, f"dx{i}", mode)
    _check_grads(m, z, mode, float(np.abs(z["dy"]).max()) * 10)
    _check_buffers(m, z, mode)
    m.eval()
    with torch.no_grad():
        ye = m(*[a.detach() for a in args])
    expect_key(_np(ye), z, "y_eval", mode)


@pytest.mark.parametrize("name", ["model_b4_skip.npz", "model_b2_noskip_l2.npz", "model_b4_skip_64.npz"])
def test_model_golden(golden_dir, name, mode):
    from train.unet import TemporalUNetDualView
    z = np.load(os.path.join(golden_dir, name))
    base_ch, skip, L, B, T, H, W = [int(v) for v in z["meta"]]
    m = _load_sd(TemporalUNetDualView(base_ch=base_ch, lstm_layers=L, use_skip_lstm=bool(skip)), z)
    x = _cuda(z["x"], True)
    m.train()
    out, st = m(x)
    assert isinstance(out, list) and len(out) == T and tuple(out[0].shape) == (B, 1, H, W)
    y = torch.stack(out, dim=1)
    expect_key(_np(y), z, "y_train", mode)
    for l in range(L):
        expect_key(_np(st[l][0]), z, f"hT{l}", mode)
        expect_key(_np(st[l][1]), z, f"cT{l}", mode)
    (y * _cuda(z["dy"])).sum().backward()
    expect_key(_np(x.grad), z, "dx", mode)
    _check_grads(m, z, mode, float(np.abs(z["dy"]).max()) * 100)
    _check_buffers(m, z, mode)
    # eval mode with the updated running statistics, and the state round trip (unet.py:185)
    m.eval()
    with torch.no_grad():
        oe, _ = m(x.detach())
        k = T // 2
        o1, s1 = m(x.detach()[:, :k])
        o2, _ = m(x.detach()[:, k:], s1)
    expect_key(_np(torch.stack(oe, dim=1)), z, "y_eval", mode)
    expect_key(_np(torch.stack(o1 + o2, dim=1)), z, "y_eval_split", mode)


# ------------------------------------------------------------------------------------------------
# tensor-core path (bf16) against the numpy oracle on shapes the tcgen05 kernels tile
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,ch,B,T,H,W,with_state", [(16, 16, 2, 3, 8, 8, True), (32, 32, 2, 3, 8, 16, False),
                                                       (64, 64, 4, 2, 4, 4, True)])
def test_convlstm_tc_vs_oracle(cin, ch, B, T, H, W, with_state):
    import unet_convlstm_b200 as pkg
    from oracle import unet_oracle as O
    from train.unet import ConvLSTM
    from unet_convlstm_b200 import ops
    pkg.set_precision("bf16")
    mode = "bf16"
    assert ops.lstm_tc_ok(torch.empty(B, H, W, cin, device="cuda", dtype=torch.bfloat16), ch)
    rng = np.random.default_rng(5)
    torch.manual_seed(5)
    m = ConvLSTM(cin, ch).cuda()
    x = rng.standard_normal((T, B, cin, H, W)).astype(np.float32)
    h0 = (0.5 * rng.standard_normal((B, ch, H, W))).astype(np.float32)
    c0 = (0.5 * rng.standard_normal((B, ch, H, W))).astype(np.float32)
    dout = rng.standard_normal((T, B, ch, H, W)).astype(np.float32)
    dc_last = rng.standard_normal((B, ch, H, W)).astype(np.float32)
    w = _np(m.layers[0].conv.weight).astype(np.float64)
    b = _np(m.layers[0].conv.bias).astype(np.float64)
    state = [(h0.astype(np.float64), c0.astype(np.float64))] if with_state else None
    o_ref, st_ref, caches = O.convlstm_fwd(list(x.astype(np.float64)), [(w, b)], state)
    dx_ref, wg_ref, d0_ref = O.convlstm_bwd(caches, list(dout.astype(np.float64)), [(0.0, dc_last.astype(np.float64))])

    xs = [_cuda(x[t], True) for t in range(T)]
    st = [(_cuda(h0, True), _cuda(c0, True))] if with_state else None
    out, new_state = m(xs, st)
    expect(_np(torch.stack(out)), np.stack(o_ref), "bf16")
    expect(_np(new_state[0][1]), st_ref[0][1], "bf16")
    loss = sum((o * _cuda(dout[t])).sum() for t, o in enumerate(out)) + (new_state[0][1] * _cuda(dc_last)).sum()
    loss.backward()
    expect(np.stack([_np(v.grad) for v in xs]), np.stack(dx_ref), "bf16")
    expect(_np(m.layers[0].conv.weight.grad), wg_ref[0][0], "bf16")
    expect(_np(m.layers[0].conv.bias.grad), wg_ref[0][1], "bf16")
    if with_state:
        expect(_np(st[0][0].grad), d0_ref[0][0], "bf16")
        expect(_np(st[0][1].grad), d0_ref[0][1], "bf16")


def _port_run(sd, x, dy, dtype, autocast, training):
    """oracle/torch_port.py on CPU: returns {key: float64 ndarray} of y, dx and all parameter gradients."""
    from oracle import torch_port as TP
    p = TP.params_from_state_dict(sd, dtype)
    xt = torch.from_numpy(x).to(dtype).requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out, _ = TP.temporal_unet(p, xt, None, training=training, track=False)
    y = torch.stack(out, dim=1)
    (y.to(dtype) * torch.from_numpy(dy).to(dtype)).sum().backward()
    r = {"y": y.detach().double().numpy(), "dx": xt.grad.double().numpy()}
    for k, v in p.items():
        if v.requires_grad:
            r["g." + k] = v.grad.double().numpy()
    return r


@pytest.mark.parametrize("training", [False, True])
def test_model_tc_vs_oracle(training):
    """base_ch=16 at 64x64, B=4: every layer runs on the tcgen05 path (the 2 input channels are
    zero-padded to one K chunk).  Oracle: the torch CPU port in fp64.  Noise floor: the same port under
    bf16 autocast, i.e. what the reference's own arithmetic library delivers at this precision --
    BatchNorm backward amplifies bf16 rounding through the 18-conv-deep net, so gradients are held to
    a small multiple of that floor rather than to 2e-2."""
    import unet_convlstm_b200 as pkg
    from train.unet import TemporalUNetDualView
    pkg.set_precision("bf16")
    B, T, H, W = 4, 2, 64, 64
    torch.manual_seed(7)
    m = TemporalUNetDualView(base_ch=16, use_skip_lstm=True)
    rng = np.random.default_rng(7)
    x = rng.random((B, T, 2, H, W)).astype(np.float32)
    dy = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    ref = _port_run(sd, x, dy, torch.float64, False, training)
    r16 = _port_run(sd, x, dy, torch.float32, True, training)
    m = m.cuda()
    m.train(training)
    xg = _cuda(x, True)
    out, _ = m(xg)
    y = torch.stack(out, dim=1)
    (y * _cuda(dy)).sum().backward()
    got = {"y": _np(y), "dx": _np(xg.grad)}
    got.update({"g." + k: _np(prm.grad) for k, prm in m.named_parameters()})
    keys = [k for k, v in ref.items() if np.abs(v).max() >= 1e-6 * np.abs(dy).max()]  # skip the
    # mathematically-zero conv-bias gradients of train mode
    floors = {k: np.array([rel(r16[k], ref[k]), rel2(r16[k], ref[k])]) for k in keys}
    problem = np.max(np.stack(list(floors.values())), axis=0)
    bad = []
    for k in keys:
        ok, err = close(got[k], ref[k], "bf16", floors[k], problem)
        if not ok:
            bad.append((k, err))
    assert not bad, bad


# ------------------------------------------------------------------------------------------------
# tcgen05 kernels against the CUDA-core fp32 kernels (which the fp32-mode golden tests above pin to
# the reference at 1e-5) on identical bf16-rounded operands: only accumulation order and the final
# bf16 rounding of the output differ
# ------------------------------------------------------------------------------------------------
def _simt_conv(x0, x1, wp, bias, ks, n_out):
    from unet_convlstm_b200 import _lib
    T, B, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[-1]
    out = torch.empty((T, B, H, W, n_out), device="cuda", dtype=torch.float32)
    x0f, x1f, wf = x0.float(), None if x1 is None else x1.float(), wp.float()
    _lib.call("b200_conv_simt_fwd", x0f.data_ptr(), C0, None if x1f is None else x1f.data_ptr(), C1, T * B, H, W,
              wf.data_ptr(), None if bias is None else bias.data_ptr(), n_out, ks, out.data_ptr(), n_out, n_out,
              None, 0, 1, 1, 0, torch.cuda.current_stream().cuda_stream)
    return out


@pytest.mark.parametrize("T,B,H,W,C0,C1,N,ks", [(2, 3, 8, 8, 64, 64, 256, 3), (1, 2, 32, 32, 64, 128, 192, 3),
                                                (1, 2, 16, 16, 32, 32, 96, 3), (2, 4, 4, 4, 128, 0, 256, 3),
                                                (1, 1, 16, 128, 64, 0, 128, 3), (1, 2, 16, 16, 16, 0, 32, 1),
                                                (1, 2, 64, 64, 16, 0, 64, 3), (1, 1, 12, 16, 64, 0, 64, 3),
                                                # narrow layers: the halo kernel (conv_halo.cu), resident and
                                                # streamed weights, one and two sources, odd heights
                                                (2, 4, 64, 64, 64, 0, 64, 3), (1, 4, 64, 64, 64, 64, 64, 3),
                                                (1, 4, 32, 32, 128, 0, 128, 3), (1, 2, 32, 32, 128, 128, 128, 3),
                                                (1, 3, 16, 16, 64, 0, 128, 3), (1, 2, 7, 16, 64, 0, 48, 3),
                                                (1, 2, 64, 64, 128, 0, 64, 3)])
def test_conv_tc_vs_simt(T, B, H, W, C0, C1, N, ks):
    from unet_convlstm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randn(T, B, H, W, C0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(T, B, H, W, C1, device="cuda", generator=g).bfloat16() if C1 else None
    w = torch.randn(N, C0 + C1, ks, ks, device="cuda", generator=g) / ((C0 + C1) * ks * ks) ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    assert ops.tc_conv_ok(x0, x1, N)
    out = torch.full((T, B, H, W, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv_fwd(x0, x1, wp, bias, ks, out)
    ref = _simt_conv(x0, x1, wp, bias, ks, N)
    assert not torch.isnan(out.float()).any()
    assert rel(_np(out), _np(ref)) < 6e-3  # bf16 output rounding (2^-8 relative)
    # data gradient of the same convolution: two destinations (virtual concat split)
    wd = ops.pack_conv_weight_dgrad(w, torch.bfloat16)
    dz = torch.randn(T, B, H, W, N, device="cuda", generator=g).bfloat16()
    if N % 16 == 0 and ops.tc_conv_ok(dz, None, C0 + C1):
        d0 = torch.full((T, B, H, W, C0), float("nan"), device="cuda", dtype=torch.bfloat16)
        d1 = torch.full((T, B, H, W, C1), float("nan"), device="cuda", dtype=torch.bfloat16) if C1 else None
        ops.conv_fwd(dz, None, wd, None, ks, d0, d1)
        refd = _simt_conv(dz, None, wd, None, ks, C0 + C1)
        got = d0 if d1 is None else torch.cat([d0, d1], dim=-1)
        assert rel(_np(got), _np(refd)) < 6e-3


@pytest.mark.parametrize("T,B,H,W,C0,C1,N", [(3, 4, 64, 64, 64, 0, 64),     # halo kernel, resident weights
                                             (2, 5, 32, 32, 128, 128, 128),  # halo, two sources, streamed weights
                                             (3, 6, 8, 8, 256, 0, 512),      # generic kernel, two N tiles
                                             (2, 3, 16, 16, 64, 0, 48),      # N not a multiple of the N tile
                                             (2, 7, 4, 4, 128, 0, 256),      # batch tail rows outside the tensor
                                             (2, 2, 64, 64, 16, 0, 64)])     # first layer (16-channel K chunk)
def test_conv_fused_bn_stats(T, B, H, W, C0, C1, N):
    """The BatchNorm sums produced by the conv epilogue (b200_conv_bnstats_tc_fwd) equal the sums of the
    stored bf16 output (b200_bn_stats), and the output itself is unchanged."""
    from unet_convlstm_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.randn(T, B, H, W, C0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(T, B, H, W, C1, device="cuda", generator=g).bfloat16() if C1 else None
    w = torch.randn(N, C0 + C1, 3, 3, device="cuda", generator=g) / ((C0 + C1) * 9) ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    plain = torch.empty((T, B, H, W, N), device="cuda", dtype=torch.bfloat16)
    assert ops.conv_fwd(x0, x1, wp, bias, 3, plain) is False
    out = torch.full((T, B, H, W, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ws = torch.full((2, T, N), float("nan"), device="cuda", dtype=torch.float64)
    assert ops.conv_fwd(x0, x1, wp, bias, 3, out, bn_ws=ws) is True
    # the plain call may run on the CTA-pair kernel (different accumulation order): equal up to bf16 rounding
    assert rel(_np(out), _np(plain)) < 8e-3
    ref = torch.empty((2, T, N), device="cuda", dtype=torch.float64)
    _lib.call("b200_bn_stats", out.data_ptr(), T, B * H * W, N, 0, ref[0].data_ptr(), ref[1].data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    exact = out.double().reshape(T, -1, N)
    np.testing.assert_allclose(_np(ws[0]), _np(exact.sum(1)), rtol=0, atol=2e-5 * B * H * W)
    np.testing.assert_allclose(_np(ws[1]), _np((exact * exact).sum(1)), rtol=2e-5)
    np.testing.assert_allclose(_np(ws[0]), _np(ref[0]), rtol=0, atol=2e-5 * B * H * W)
    np.testing.assert_allclose(_np(ws[1]), _np(ref[1]), rtol=2e-5)


@pytest.mark.parametrize("T,B,H,W,Nz,C0,C1,ks", [(1, 2, 8, 8, 128, 64, 0, 3), (2, 4, 4, 4, 256, 128, 128, 3),
                                                 (3, 2, 16, 16, 64, 64, 0, 3), (2, 2, 32, 32, 128, 32, 96, 3),
                                                 (1, 1, 12, 16, 64, 16, 0, 3), (2, 2, 8, 128, 320, 320, 0, 1),
                                                 (1, 2, 8, 8, 320, 64, 0, 3), (2, 2, 16, 16, 128, 64, 64, 3),
                                                 (4, 8, 64, 64, 64, 64, 0, 3)])
def test_wgrad_tc_vs_simt(T, B, H, W, Nz, C0, C1, ks):
    from unet_convlstm_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(3)
    dz = torch.randn(T, B, H, W, Nz, device="cuda", generator=g).bfloat16()
    srcs = [torch.randn(T, B, H, W, C0, device="cuda", generator=g).bfloat16()]
    if C1:
        srcs.append(torch.randn(T, B, H, W, C1, device="cuda", generator=g).bfloat16())
    Ct = C0 + C1
    assert _lib.supported("b200_wgrad_tc_supported", B, H, W, Nz, C0)
    dw = torch.zeros(ks * ks, Nz, Ct, device="cuda")
    ref = torch.zeros(ks * ks, Nz, Ct, device="cuda")
    koff = 0
    st = torch.cuda.current_stream().cuda_stream
    for s_ in srcs:
        ops.conv_wgrad(dz, s_, ks, dw, koff)
        dzf, sf = dz.float(), s_.float()
        _lib.call("b200_wgrad_simt", dzf.data_ptr(), Nz, sf.data_ptr(), s_.shape[-1], T * B, H, W, ks, ref.data_ptr(),
                  Ct, koff, 1, st)
        koff += s_.shape[-1]
    assert rel(_np(dw), _np(ref)) < 1e-4  # both accumulate exact bf16 products in fp32


@pytest.mark.parametrize("B,H,W,Cin,Ch,with_state", [(2, 16, 16, 64, 64, True), (2, 16, 16, 64, 64, False),
                                                     (4, 8, 8, 128, 128, True), (2, 32, 32, 32, 32, True),
                                                     (2, 16, 16, 32, 16, True), (8, 4, 4, 256, 256, True)])
def test_lstm_cell_fused_vs_unfused(B, H, W, Cin, Ch, with_state):
    """The fused tcgen05 cell kernel (gate math in the GEMM epilogue) against conv + gate-math kernels."""
    from unet_convlstm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    bf = torch.bfloat16
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(bf)
    h = torch.randn(B, H, W, Ch, device="cuda", generator=g).to(bf) if with_state else None
    c = torch.randn(B, H, W, Ch, device="cuda", generator=g) if with_state else None
    w = torch.randn(4 * Ch, Cin + Ch, 3, 3, device="cuda", generator=g) / ((Cin + Ch) * 9) ** 0.5
    b = torch.randn(4 * Ch, device="cuda", generator=g) * 0.1
    assert ops.lstm_tc_ok(x, Ch)
    wp_il, bp_il = ops.pack_lstm_weight(w, b, bf)
    wp = ops.pack_conv_weight(w, bf)
    mk = lambda dt, n=Ch: torch.full((B, H, W, n), float("nan"), device="cuda", dtype=dt)  # noqa: E731
    c1, h1, g1 = mk(torch.float32), mk(bf), mk(bf, 4 * Ch)
    ops.lstm_cell_fwd_fused(x, h, c, wp_il, bp_il, c1, h1, g1, 3)
    c2, h2, g2 = mk(torch.float32), mk(torch.float32), mk(torch.float32, 4 * Ch)
    hz = h if h is not None else torch.zeros(B, H, W, Ch, device="cuda", dtype=bf)
    ops.lstm_cell_fwd_unfused(x.float(), hz.float(), c, wp.float(), b, c2, h2, g2, 3,
                              torch.empty(B, H, W, 4 * Ch, device="cuda"))
    # tanh.approx (fused, bf16 mode) vs tanhf/expf: 2^-11 relative, plus bf16 rounding of h and gates
    assert rel(_np(c1), _np(c2)) < 3e-3
    assert rel(_np(h1), _np(h2)) < 8e-3
    assert rel(_np(g1), _np(g2)) < 8e-3


# ------------------------------------------------------------------------------------------------
# script level: the optimisation loop of train/overfit_check.py (reference :91-123 -- AdamW lr 1e-3 wd 1e-4,
# masked MSE on one fixed batch) driven on the CUDA model and on the CPU port of the reference with the
# same initial state_dict, data and hyper-parameters: the loss curves must match
# ------------------------------------------------------------------------------------------------
# fp32: the first steps agree to 1e-5; AdamW then amplifies rounding differences (sign-like updates for
# small gradients), as it would between any two fp32 back ends -- 1% over 25 steps
@pytest.mark.parametrize("mode_name,iters,tol", [("fp32", 25, 1e-2), ("bf16", 25, 0.15)])
def test_overfit_loss_curve_matches_reference_port(mode_name, iters, tol):
    import unet_convlstm_b200 as pkg
    from oracle import torch_port as TP
    from train.unet import TemporalUNetDualView
    pkg.set_precision(mode_name)
    B, T, H, W = 8, 4, 32, 32
    rng = np.random.default_rng(11)
    x = rng.random((B, T, 2, H, W)).astype(np.float32)
    y = np.tanh(rng.standard_normal((B, T, 1, H, W))).astype(np.float32)
    mask = (rng.random((B, T, 1, H, W)) > 0.4).astype(np.float32)
    torch.manual_seed(5)
    m = TemporalUNetDualView(base_ch=16, use_skip_lstm=True)
    sd = {k: v.clone() for k, v in m.state_dict().items()}

    # reference port on CPU (fp32, like the reference's own CPU path)
    p = TP.params_from_state_dict(sd, torch.float32)
    opt = torch.optim.AdamW([v for v in p.values() if v.requires_grad], lr=1e-3, weight_decay=1e-4)
    xt, yt, mt = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(mask)
    ref_curve = []
    for _ in range(iters):
        opt.zero_grad()
        out, _ = TP.temporal_unet(p, xt, None, training=True)
        yp = torch.stack(out, dim=1)
        loss = (((yp - yt) ** 2) * mt).sum() / mt.sum().clamp(min=1.0)
        loss.backward()
        opt.step()
        ref_curve.append(loss.item())

    m = m.cuda().train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)
    xc, yc, mc = xt.cuda(), yt.cuda(), mt.cuda()
    curve = []
    for _ in range(iters):
        opt.zero_grad()
        out, _ = m(xc)
        yp = torch.stack(out, dim=1) if isinstance(out, list) else out
        loss = (((yp - yc) ** 2) * mc).sum() / mc.sum().clamp(min=1.0)
        loss.backward()
        opt.step()
        curve.append(loss.item())
    ref_curve, curve = np.array(ref_curve), np.array(curve)
    assert ref_curve[-1] < 0.9 * ref_curve[0]  # the loop does optimise
    if mode_name == "fp32":
        assert abs(curve[0] - ref_curve[0]) / ref_curve[0] < 1e-5
        assert abs(curve[1] - ref_curve[1]) / ref_curve[1] < 5e-5
    assert np.abs(curve - ref_curve).max() / ref_curve.max() < tol, (curve, ref_curve)
    pkg.set_precision("bf16")


@pytest.mark.parametrize("T,B,H,W,C,have_h0", [(5, 8, 4, 4, 256, False), (4, 2, 16, 16, 64, True), (3, 64, 8, 8, 128, True),
                                               (6, 256, 4, 4, 64, False)])
def test_lstm_persistent_sequence_kernel_matches_stepwise(T, B, H, W, C, have_h0):
    """The timestep-persistent cooperative kernel (one launch per layer, grid-wide step counter) must be
    bit-identical to T per-step launches of the same fused cell kernel."""
    from unet_convlstm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    bf = torch.bfloat16
    x = torch.randn(T, B, H, W, C, device="cuda", generator=g).to(bf)
    w = torch.randn(4 * C, 2 * C, 3, 3, device="cuda", generator=g) / (18 * C) ** 0.5
    b = torch.randn(4 * C, device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_lstm_weight(w, b, bf)
    res = []
    for persistent in (False, True):
        h_all = torch.full((T + 1, B, H, W, C), float("nan"), device="cuda", dtype=bf)
        c_all = torch.full((T + 1, B, H, W, C), float("nan"), device="cuda")
        gates = torch.full((T, B, H, W, 4 * C), float("nan"), device="cuda", dtype=bf)
        if have_h0:
            gg = torch.Generator(device="cuda").manual_seed(8)
            h_all[0] = torch.randn(B, H, W, C, device="cuda", generator=gg).to(bf)
            c_all[0] = torch.randn(B, H, W, C, device="cuda", generator=gg)
        if persistent:
            ops.lstm_seq_fwd_fused(x, h_all, c_all, wp, bp, gates, have_h0, 3)
        else:
            for t in range(T):
                z0 = t == 0 and not have_h0
                ops.lstm_cell_fwd_fused(x[t], None if z0 else h_all[t], None if z0 else c_all[t], wp, bp, c_all[t + 1],
                                        h_all[t + 1], gates[t], 3)
        torch.cuda.synchronize()
        res.append((h_all[1:].clone(), c_all[1:].clone(), gates.clone()))
    for a, b_ in zip(res[0], res[1]):
        assert not torch.isnan(a.float()).any()
        assert torch.equal(a, b_)


def test_wgrad_tc_long_reduction_multi_producer():
    """Many reduction blocks per unit and tap-group sets of different sizes (2,2,2,2,1): the producer
    warps of wgrad_tc.cu must stay within one lap of the stage ring (a producer that skipped the stages
    of a group set it does not own once ran ahead and corrupted the pipeline)."""
    from unet_convlstm_b200 import _lib, ops
    T, B, H, W, Nz, C = 2, 64, 16, 16, 256, 256
    g = torch.Generator(device="cuda").manual_seed(11)
    dz = torch.randn(T, B, H, W, Nz, device="cuda", generator=g).bfloat16()
    src = torch.randn(T, B, H, W, C, device="cuda", generator=g).bfloat16()
    dw = torch.zeros(9, Nz, C, device="cuda")
    ops.conv_wgrad(dz, src, 3, dw, 0)
    ref = torch.zeros(9, Nz, C, device="cuda")
    dzf, sf = dz.float(), src.float()
    _lib.call("b200_wgrad_simt", dzf.data_ptr(), Nz, sf.data_ptr(), C, T * B, H, W, 3, ref.data_ptr(), C, 0, 1,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert _lib.lib().b200_device_error() == 0
    assert rel(_np(dw), _np(ref)) < 1e-4


@pytest.mark.parametrize("T,B,H,W,Cin,Ch,have_h0,need_dx,with_dh,with_dc", [
    (5, 8, 4, 4, 256, 256, False, True, True, False),    # bottleneck-like, zero initial state
    (4, 2, 16, 16, 64, 64, True, True, True, True),      # carried state: dh0 / dc0 are produced
    (3, 64, 8, 8, 128, 128, True, False, True, False),   # input does not need a gradient: dx tiles skipped
    (4, 3, 16, 16, 32, 64, False, True, True, True),     # Cin != Ch, N tile straddles the dx | dh boundary
    (2, 2, 16, 16, 64, 64, False, True, False, True),    # only the final cell state is used downstream
])
def test_lstm_fused_bptt_matches_stepwise(T, B, H, W, Cin, Ch, have_h0, need_dx, with_dh, with_dc):
    """The timestep-persistent fused BPTT kernel (dgrad conv with the gate gradients of the previous step in
    its epilogue) against the per-step path (b200_lstm_gates_bwd + b200_conv_tc_fwd per timestep): the only
    arithmetic difference is that dh_{t-1} is no longer rounded to bf16 between the two kernels."""
    from unet_convlstm_b200 import functional as Fn, ops
    g = torch.Generator(device="cuda").manual_seed(12)
    bf = torch.bfloat16
    w = (torch.randn(4 * Ch, Cin + Ch, 3, 3, device="cuda", generator=g) / (9 * (Cin + Ch)) ** 0.5).requires_grad_(True)
    b = (torch.randn(4 * Ch, device="cuda", generator=g) * 0.1).requires_grad_(True)
    x0 = torch.randn(T, B, H, W, Cin, device="cuda", generator=g).to(bf)
    h0v = torch.randn(B, H, W, Ch, device="cuda", generator=g).to(bf)
    c0v = torch.randn(B, H, W, Ch, device="cuda", generator=g)
    dh_up = torch.randn(T, B, H, W, Ch, device="cuda", generator=g).to(bf)
    dc_up = torch.randn(B, H, W, Ch, device="cuda", generator=g)
    res = []
    old = ops.FUSED_BPTT
    try:
        for persistent in (False, True):
            ops.FUSED_BPTT = persistent
            # poison the caching allocator's free blocks: an output element the kernel forgets to write must
            # not inherit the right value from the previous pass of this loop
            junk = torch.full((256 << 20,), float("nan"), device="cuda")
            del junk
            x = x0.clone().requires_grad_(need_dx)
            h0 = h0v.clone().requires_grad_(True) if have_h0 else None
            c0 = c0v.clone().requires_grad_(True) if have_h0 else None
            w.grad = b.grad = None
            h_seq, c_T = Fn.ConvLSTMSeq.apply(x, h0, c0, w, b, Fn.WeightCache())
            loss = 0.0
            if with_dh:
                loss = loss + (h_seq.float() * dh_up.float()).sum()
            if with_dc:
                loss = loss + (c_T * dc_up).sum()
            loss.backward()
            torch.cuda.synchronize()
            res.append({"dx": x.grad if need_dx else None, "dh0": h0.grad if have_h0 else None,
                        "dc0": c0.grad if have_h0 else None, "dw": w.grad.clone(), "db": b.grad.clone()})
    finally:
        ops.FUSED_BPTT = old
    for k in res[0]:
        a, f = res[0][k], res[1][k]
        assert (a is None) == (f is None), k
        if a is not None:
            assert torch.isfinite(f.float()).all(), k
            assert rel2(_np(f), _np(a)) < 6e-3, (k, rel2(_np(f), _np(a)))


def test_fused_loss_matches_reference_compute_loss(golden_dir):
    """unet_convlstm_b200.loss.compute_loss (two fused kernels) against the fixture produced by the reference's
    own main.compute_loss (tests/golden/make_golden_loss.py): loss and d loss / d y_pred, fp32 arithmetic."""
    from unet_convlstm_b200.loss import compute_loss
    z = np.load(os.path.join(golden_dir, "loss_main_compute_loss.npz"))
    for name in ("a", "b", "c", "z"):
        for tag in ("mask", "nomask", "ignored"):
            if f"{name}.{tag}.loss" not in z.files:
                continue
            yp = _cuda(z[f"{name}.yp"], True)
            mask = None if tag == "nomask" else _cuda(z[f"{name}.mask"])
            loss = compute_loss(yp, _cuda(z[f"{name}.y"]), mask, use_mask=(tag != "ignored"))
            (3.0 * loss).backward()
            ref_l, ref_g = float(z[f"{name}.{tag}.loss"]), 3.0 * z[f"{name}.{tag}.grad"]
            assert abs(float(loss) - ref_l) <= 2e-6 * max(1.0, abs(ref_l)), (name, tag, float(loss), ref_l)
            np.testing.assert_allclose(_np(yp.grad), ref_g, rtol=2e-5, atol=1e-7 * max(1.0, np.abs(ref_g).max()),
                                       err_msg=f"{name}.{tag}")


def test_fused_loss_fullsize_against_oracle_sample():
    """BASELINE.json configs[1] size (B = 256, T = 20, 64x64): the loss against the numpy oracle evaluated on the
    full maps in fp64 (seconds on the CPU) and the gradient on a sample of images."""
    from oracle import loss_oracle as LO
    from unet_convlstm_b200.loss import compute_loss
    g = torch.Generator(device="cuda").manual_seed(2)
    shape = (256, 20, 1, 64, 64)
    yp = torch.randn(shape, device="cuda", generator=g).requires_grad_(True)
    y = torch.randn(shape, device="cuda", generator=g).clamp_(-1, 1)
    mask = (torch.rand(shape, device="cuda", generator=g) < 0.3).float()
    loss = compute_loss(yp, y, mask)
    loss.backward()
    ref_l, ref_g = LO.compute_loss(_np(yp), _np(y), _np(mask))
    assert abs(float(loss) - ref_l) <= 1e-5 * abs(ref_l)
    got = _np(yp.grad)
    assert rel(got[::37], ref_g[::37]) < 1e-4


@pytest.mark.parametrize("T,B,H,W,Cin,Cout,Hd,Wd", [(2, 3, 8, 8, 64, 32, 16, 16), (1, 2, 4, 4, 256, 128, 8, 8),
                                                    (1, 2, 8, 16, 32, 16, 17, 35),   # F.pad offsets (1 / 3 extra)
                                                    (2, 4, 32, 32, 128, 64, 64, 64)])
def test_convT_fused_shuffle_matches_gemm_plus_shuffle(T, B, H, W, Cin, Cout, Hd, Wd):
    """b200_convT2x2_tc_fwd (pixel shuffle + bias in the GEMM epilogue) against the two-kernel path
    b200_conv_tc_fwd + b200_shuffle2x2: bit-identical (same accumulation, same single bf16 rounding)."""
    from unet_convlstm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    bf = torch.bfloat16
    x = torch.randn(T, B, H, W, Cin, device="cuda", generator=g).to(bf)
    w = torch.randn(Cin, Cout, 2, 2, device="cuda", generator=g) / Cin ** 0.5
    bias = torch.randn(Cout, device="cuda", generator=g)
    wf, _ = ops.pack_convT_weight(w, bf)
    junk = torch.full((64 << 20,), float("nan"), device="cuda")
    del junk
    y = ops.convT2x2_fwd(x, wf, bias, Cout, Hd, Wd)
    z = torch.empty((T, B, H, W, 4 * Cout), device="cuda", dtype=bf)
    ops.conv_fwd(x, None, wf, None, 1, z)
    ref = ops.shuffle2x2(z.float(), bias, Hd, Wd)  # fp32 shuffle of the bf16 GEMM output + bias
    assert torch.isfinite(y.float()).all()
    # the fused epilogue adds the bias before the single bf16 rounding, the two-kernel path rounds twice
    assert rel(_np(y), _np(ref)) < 6e-3
    # reference ConvTranspose2d arithmetic on the same bf16-rounded operands
    xt = x.float().permute(0, 1, 4, 2, 3).reshape(T * B, Cin, H, W)
    rt = torch.nn.functional.conv_transpose2d(xt, w.to(bf).float(), bias, stride=2)
    oy, ox = (Hd - 2 * H) // 2, (Wd - 2 * W) // 2
    full = torch.zeros(T * B, Cout, Hd, Wd, device="cuda")
    full[:, :, oy:oy + 2 * H, ox:ox + 2 * W] = rt
    assert rel(_np(y.reshape(T * B, Hd, Wd, Cout).permute(0, 3, 1, 2)), _np(full)) < 6e-3


def test_inference_folded_batchnorm_matches_eval_path_and_oracle():
    """model.eval() under torch.no_grad(): DoubleConv runs conv + BatchNorm(eval) + ReLU as ONE kernel
    (b200_conv_affine_relu_tc_fwd).  Against the unfused eval path of the same model (autograd enabled) and
    against the fp64 CPU oracle in eval mode with non-trivial running statistics."""
    import unet_convlstm_b200 as pkg
    from oracle import torch_port as TP
    from train.unet import TemporalUNetDualView
    from unet_convlstm_b200 import _lib
    pkg.set_precision("bf16")
    B, T, H, W = 3, 3, 64, 64
    torch.manual_seed(5)
    m = TemporalUNetDualView(base_ch=16, use_skip_lstm=True)
    gen = torch.Generator().manual_seed(6)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data = 0.5 + torch.rand(mod.weight.shape, generator=gen)
            mod.bias.data = 0.2 * torch.randn(mod.bias.shape, generator=gen)
            mod.running_mean.data = 0.1 * torch.randn(mod.running_mean.shape, generator=gen)
            mod.running_var.data = 0.5 + torch.rand(mod.running_var.shape, generator=gen)
    x = np.random.default_rng(5).random((B, T, 2, H, W)).astype(np.float32)
    p = TP.params_from_state_dict({k: v.clone() for k, v in m.state_dict().items()}, torch.float64)
    with torch.no_grad():
        ref, ref_state = TP.temporal_unet(p, torch.from_numpy(x).double(), None, training=False, track=False)
    ref = torch.stack(ref, dim=1).numpy()
    m = m.cuda().eval()
    xg = torch.from_numpy(x).cuda()
    out_u, _ = m(xg)                               # autograd enabled: conv, finalize, normalise/ReLU kernels
    calls0 = dict(_lib.CALLS)
    with torch.no_grad():
        out_f, state = m(xg)                       # folded
    assert _lib.CALLS.get("b200_conv_affine_relu_tc_fwd", 0) - calls0.get("b200_conv_affine_relu_tc_fwd", 0) == 18
    assert _lib.CALLS.get("b200_bn_relu_apply", 0) == calls0.get("b200_bn_relu_apply", 0)
    yu, yf = _np(torch.stack(out_u, dim=1)), _np(torch.stack(out_f, dim=1))
    assert rel2(yf, ref) < 2e-2 and rel2(yu, ref) < 2e-2
    assert rel2(yf, yu) < 2e-2
    assert rel2(_np(state[0][0]), ref_state[0][0].numpy()) < 2e-2


class _Preset(torch.nn.Module):
    """Stand-in model of tests/golden/make_golden_metrics.py: returns preset predictions as a list of T frames."""

    def __init__(self, preds):
        super().__init__()
        self.preds, self.i = preds, 0

    def forward(self, x):
        p = self.preds[self.i]
        self.i += 1
        return [p[:, t] for t in range(p.shape[1])], None


def test_epoch_loop_matches_reference_evaluate(golden_dir):
    """unet_convlstm_b200.loop.evaluate (fused loss + on-device metric accumulators, host batches prefetched on a
    side stream) against the fixture produced by the reference's own main.evaluate (main.py:150-204)."""
    import types
    from unet_convlstm_b200 import loop
    z = np.load(os.path.join(golden_dir, "metrics_main_evaluate.npz"))
    for tr in ("asinh", "signed_log", "none"):
        tmin, tmax, scale = z[f"{tr}.params"]
        ds = types.SimpleNamespace(trans_min=tmin, trans_max=tmax, y_scale=scale, y_transform=tr)
        host = [tuple(torch.from_numpy(z[f"{tr}.b{i}.{k}"]) for k in ("x", "y", "mask")) for i in range(3)]
        preds = [_cuda(z[f"{tr}.b{i}.pred"]) for i in range(3)]
        for use in (True, False):
            got = loop.evaluate(_Preset(preds), host, "cuda", ds, use_mask=use)
            np.testing.assert_allclose(got, z[f"{tr}.use{int(use)}.result"], rtol=3e-6, atol=3e-7,
                                       err_msg=f"{tr} use={use}")
    ds = types.SimpleNamespace(trans_min=z["asinh.params"][0], trans_max=z["asinh.params"][1],
                               y_scale=z["asinh.params"][2], y_transform="asinh")
    x, y, m = (_cuda(z[f"asinh.b1.{k}"]) for k in ("x", "y", "mask"))          # device batches are taken as they are
    got = loop.evaluate(_Preset([_cuda(z["asinh.b1.pred"])]), [(x, y, torch.zeros_like(m))], "cuda", ds)
    np.testing.assert_allclose(got, z["empty.result"], rtol=3e-6, atol=3e-7)


def test_train_one_epoch_matches_stepwise_host_loop():
    """loop.train_one_epoch against the reference's loop structure restated step by step (main.py:77-145:
    per-step loss.item(), host-side NumPy metric lists through the oracle) on the same model, data and seeds."""
    import copy
    import types
    import unet_convlstm_b200 as pkg
    from oracle import metrics_oracle as MO
    from train.unet import TemporalUNetDualView
    from unet_convlstm_b200 import loop
    from unet_convlstm_b200.loss import compute_loss
    pkg.set_precision("fp32")
    try:
        torch.manual_seed(11)
        m1 = TemporalUNetDualView(base_ch=4, use_skip_lstm=True).cuda()
        m2 = copy.deepcopy(m1)
        rng = np.random.default_rng(5)
        batches = []
        for b in (3, 2, 3):
            x = torch.from_numpy((rng.random((b, 3, 2, 16, 16)) * 2).astype(np.float32))
            y = torch.from_numpy(np.clip(rng.standard_normal((b, 3, 1, 16, 16)), -1, 1).astype(np.float32))
            batches.append((x, y, (x[:, :, 0:1] > 1.1).float()))
        ds = types.SimpleNamespace(trans_min=-1.05, trans_max=1.17, y_scale=6.0, y_transform="asinh")
        o1 = torch.optim.AdamW(m1.parameters(), lr=1e-3)
        o2 = torch.optim.AdamW(m2.parameters(), lr=1e-3)
        got = loop.train_one_epoch(m1, batches, o1, "cuda", ds)
        m2.train()
        tot, n, seen = 0.0, 0, []
        for x, y, mask in batches:
            x, y, mask = x.cuda(), y.cuda(), mask.cuda()
            o2.zero_grad(set_to_none=True)
            out, _ = m2(x)
            yp = torch.stack(out, dim=1)
            loss = compute_loss(yp, y, mask, True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m2.parameters(), 1.0)
            o2.step()
            tot += loss.item() * x.size(0)
            n += x.size(0)
            seen.append((_np(yp), _np(y), _np(mask)))
        ref = (tot / n, *MO.epoch_metrics(seen, ds.trans_min, ds.trans_max, ds.y_scale, "asinh"))
        np.testing.assert_allclose(got, ref, rtol=2e-4, atol=1e-6)
        for (k, a), b in zip(m1.state_dict().items(), m2.state_dict().values()):
            assert rel2(_np(a), _np(b)) < 1e-3, k
    finally:
        pkg.set_precision("bf16")


def _optim_problem(n_tensors, seed):
    rng = np.random.default_rng(seed)
    sizes = [int(s) for s in rng.integers(1, 9000, n_tensors)]
    sizes[0], sizes[1], sizes[2] = 4096 * 3, 4096 * 2 + 5, 1            # whole chunks, ragged tail, single element
    ps = [rng.standard_normal(s) for s in sizes]
    gs = [[rng.standard_normal(s) * (10.0 if it == 0 else 0.01) for s in sizes] for it in range(4)]
    return ps, gs


@pytest.mark.parametrize("n_tensors,clip", [(7, 1.0), (101, 1.0), (101, None)])
def test_multi_tensor_clip_adamw_matches_torch(n_tensors, clip):
    """unet_convlstm_b200.optim (b200_grad_sqnorm_multi / b200_grad_clip_multi / b200_adamw_multi) against
    torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW in fp64 on the CPU -- the reference's own calls
    (main.py:106, :275).  Covers > 48 tensors (several launches), unaligned storage offsets, a clipping and a
    non-clipping step, the fused clip-in-update form and the state_dict round trip into torch.optim.AdamW."""
    from unet_convlstm_b200 import optim
    ps, gs = _optim_problem(n_tensors, n_tensors)
    ref_p = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in ps]
    pool = torch.zeros(sum(p.size + 3 for p in ps), device="cuda")
    mine, fused, off = [], [], 0
    for i, p in enumerate(ps):
        off += i % 4                                                     # storage offsets that are not 16-byte aligned
        v = pool[off:off + p.size]
        v.copy_(torch.from_numpy(p))
        off += p.size
        mine.append(torch.nn.Parameter(v))
        fused.append(torch.nn.Parameter(torch.from_numpy(p).float().cuda()))
    kw = dict(lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05)
    o_ref, o_mine, o_fused = torch.optim.AdamW(ref_p, **kw), optim.AdamW(mine, **kw), optim.AdamW(fused, **kw)
    for it in range(4):
        for lst in (ref_p, mine, fused):
            for p, g in zip(lst, gs[it]):
                p.grad = torch.from_numpy(g).to(p.dtype).to(p.device)
        if clip is not None:
            n_ref = torch.nn.utils.clip_grad_norm_(ref_p, clip)
            n_mine = optim.clip_grad_norm_(mine, clip)
            assert abs(float(n_mine) - float(n_ref)) <= 2e-6 * float(n_ref)
            for a, b in zip(mine, ref_p):
                assert rel2(_np(a.grad), b.grad.numpy()) < 1e-6
        o_ref.step(), o_mine.step(), o_fused.step(clip_max_norm=clip)
        for a, f, b in zip(mine, fused, ref_p):
            ref = b.detach().numpy()
            tol = 2e-6 * (it + 1)
            assert np.abs(_np(a) - ref).max() <= tol * max(1.0, np.abs(ref).max()), it
            assert np.abs(_np(f) - ref).max() <= tol * max(1.0, np.abs(ref).max()), it
    for a, b in zip(mine, ref_p):
        assert rel2(_np(o_mine.state[a]["exp_avg"]), o_ref.state[b]["exp_avg"].numpy()) < 1e-5
        assert rel2(_np(o_mine.state[a]["exp_avg_sq"]), o_ref.state[b]["exp_avg_sq"].numpy()) < 1e-5
        assert float(o_mine.state[a]["step"]) == 4.0
    # the state moves into torch.optim.AdamW and both continue identically
    o_t = torch.optim.AdamW(fused, **kw)
    import copy
    o_t.load_state_dict(copy.deepcopy(o_fused.state_dict()))         # load_state_dict shares the tensors otherwise
    before = [p.detach().clone() for p in fused]
    for p, g in zip(fused, gs[0]):
        p.grad = torch.from_numpy(g).float().cuda()
    o_t.step()
    after_t = [p.detach().clone() for p in fused]
    with torch.no_grad():
        for p, b in zip(fused, before):
            p.copy_(b)
    o_fused.step()
    for a, b in zip(fused, after_t):
        assert np.abs(_np(a) - _np(b)).max() <= 2e-6 * max(1.0, float(b.abs().max()))


def _bg_grads(model, x, dy, background, twice=False, keep_grads=False):
    """Gradients of sum(y * dy) with the weight gradients in line or on the background stream (every background
    block delayed by ~10 ms, so that anything reading a gradient too early sees memory that is not written yet)."""
    from unet_convlstm_b200 import ops
    old = (ops.WGRAD_STREAM, ops._BG_DEBUG_DELAY, ops.LSTM_WGRAD_CHUNK)
    ops.WGRAD_STREAM, ops._BG_DEBUG_DELAY = background, (20_000_000 if background else 0)
    ops.LSTM_WGRAD_CHUNK = 2       # T = 4: one chunk [2, 4) queued during the BPTT sweep, [0, 2) after it
    try:
        if not keep_grads:
            model.zero_grad(set_to_none=True)
        poison = torch.full((256 << 20,), float("nan"), device="cuda")   # freed memory is NaN
        del poison
        n0 = ops.BG_BLOCKS[0]
        if twice:
            k = x.shape[1] // 2
            out1, st = model(x[:, :k])
            out2, _ = model(x[:, k:], st)
            y = torch.stack(list(out1) + list(out2), dim=1)
        else:
            out, _ = model(x)
            y = torch.stack(out, dim=1)
        (y * dy).sum().backward()
        used = ops.BG_BLOCKS[0] - n0
        return {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in model.named_parameters()}, used
    finally:
        ops.WGRAD_STREAM, ops._BG_DEBUG_DELAY, ops.LSTM_WGRAD_CHUNK = old


def test_background_wgrad_stream_matches_inline():
    """Weight gradients produced on the background stream (ops.background) against the single-stream backward:
    a plain step, a step whose parameters are used TWICE in one graph (state carried between two calls: autograd sums
    the two contributions on the current stream, so the blocks must fall back in line), and gradient accumulation
    into existing .grad tensors (also in line)."""
    import unet_convlstm_b200 as pkg
    from train.unet import TemporalUNetDualView
    pkg.set_precision("fp32")
    try:
        torch.manual_seed(4)
        m = TemporalUNetDualView(base_ch=4, use_skip_lstm=True).cuda()
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.rand(2, 4, 2, 32, 32, device="cuda", generator=g) * 2
        dy = torch.randn(2, 4, 1, 32, 32, device="cuda", generator=g)
        ref, used = _bg_grads(m, x, dy, False)
        assert used == 0
        got, used = _bg_grads(m, x, dy, True)
        assert used >= 20, used                                   # 18 conv + 4 convT + 3 LSTM layers
        for k in ref:
            assert torch.isfinite(got[k]).all(), k
            assert rel2(_np(got[k]), _np(ref[k])) < 1e-4, k
        ref2, _ = _bg_grads(m, x, dy, False, twice=True)
        got2, used = _bg_grads(m, x, dy, True, twice=True)
        assert used == 0, used                                    # every parameter has two uses: all in line
        for k in ref2:
            assert rel2(_np(got2[k]), _np(ref2[k])) < 1e-4, k
        got3, used = _bg_grads(m, x, dy, True, keep_grads=True)   # accumulates on top of got2's .grad tensors
        assert used == 0, used
        for k in ref:
            assert rel2(_np(got3[k]), _np(ref2[k] + ref[k])) < 1e-4, k
        _, used = _bg_grads(m, x, dy, True)                       # and the background path is taken again afterwards
        assert used >= 20, used
        frozen = [k for k, p in m.named_parameters() if k.startswith("down1") and k.endswith(".0.weight")]
        assert frozen
        for k, p in m.named_parameters():
            p.requires_grad_(k not in frozen)                     # frozen conv weights: no weight-gradient GEMM
        got4, _ = _bg_grads(m, x, dy, True)
        for k in ref:
            if k in frozen:
                assert got4[k] is None
            else:
                assert rel2(_np(got4[k]), _np(ref[k])) < 1e-4, k
    finally:
        pkg.set_precision("bf16")


@pytest.mark.parametrize("which", ["torch_fused", "torch_foreach", "b200"])
def test_packed_weights_follow_the_optimizer(which):
    """The packed bf16 weight copies must be rebuilt after EVERY optimizer step -- including torch's fused
    optimizers, whose `torch._fused_adamw_` updates the parameters without touching their version counter, and this
    package's AdamW, which writes through raw pointers.  A model trained for two steps must answer exactly like a
    fresh model (empty caches) loaded with its state_dict: in eval mode (folded BatchNorm), and in train mode."""
    import copy
    import unet_convlstm_b200 as pkg
    from train.unet import TemporalUNetDualView
    from unet_convlstm_b200 import optim
    pkg.set_precision("bf16")
    torch.manual_seed(8)
    m = TemporalUNetDualView(base_ch=16, use_skip_lstm=True).cuda()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand(2, 3, 2, 32, 32, device="cuda", generator=g) * 2
    dy = torch.randn(2, 3, 1, 32, 32, device="cuda", generator=g)
    if which == "b200":
        opt = optim.AdamW(m.parameters(), lr=3e-2)
    else:
        opt = torch.optim.AdamW(m.parameters(), lr=3e-2, fused=(which == "torch_fused"),
                                foreach=(which == "torch_foreach"))
    m.train()
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        out, _ = m(x)
        (torch.stack(out, dim=1) * dy).sum().backward()
        opt.step()
    fresh = TemporalUNetDualView(base_ch=16, use_skip_lstm=True).cuda()
    fresh.load_state_dict(copy.deepcopy(m.state_dict()))
    m.eval(), fresh.eval()
    with torch.no_grad():
        a = torch.stack(m(x)[0], dim=1)
        b = torch.stack(fresh(x)[0], dim=1)
    assert torch.equal(a, b), float((a - b).abs().max())
    m.train(), fresh.train()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    a = torch.stack(m(xa)[0], dim=1)
    b = torch.stack(fresh(xb)[0], dim=1)
    assert torch.equal(a, b), float((a - b).abs().max())
    # ... and differentiate like it: the data-gradient weight copies ([tap flipped][K][N]) follow the optimizer too
    (a * dy).sum().backward()
    (b * dy).sum().backward()
    assert rel(_np(xa.grad), _np(xb.grad)) < 1e-4, rel(_np(xa.grad), _np(xb.grad))
    if which == "b200" and optim.ADAMW_PACK:
        # the update kernel emitted the packed copies itself: no weight was re-packed by these two forward/backward passes
        assert pkg._lib.CALLS.get("b200_adamw_pack_multi", 0) > 0
        n0 = pkg._lib.CALLS.get("b200_pack_weight", 0)
        opt.zero_grad(set_to_none=True)
        (torch.stack(m(x)[0], dim=1) * dy).sum().backward()
        opt.step()
        (torch.stack(m(x)[0], dim=1) * dy).sum().backward()
        assert pkg._lib.CALLS.get("b200_pack_weight", 0) == n0


@pytest.mark.parametrize("A,B,taps", [(64, 48, 9), (40, 18, 9), (256, 32, 4), (33, 70, 1)])
def test_adamw_pack_kernel_matches_update_then_pack(A, B, taps):
    """b200_adamw_pack = b200_adamw_multi followed by b200_pack_weight, bit for bit: parameter, both moments and two
    packed destinations (forward layout in bf16, transposed + tap-flipped data-gradient layout in fp32; gate-interleaved
    rows when A = 4 * Ch), with the clip coefficient read from the device."""
    import unet_convlstm_b200 as pkg
    from unet_convlstm_b200 import ops, optim
    g = torch.Generator(device="cuda").manual_seed(A * B)
    w = torch.randn(A, B, taps, device="cuda", generator=g)
    gr = torch.randn(A, B, taps, device="cuda", generator=g)
    m = 0.1 * torch.randn(A, B, taps, device="cuda", generator=g)
    v = torch.rand(A, B, taps, device="cuda", generator=g)
    hyper = (1e-2, 0.9, 0.999, 1e-8, 1e-2)
    sq = optim.grad_sqnorm([gr])
    perm = (A // 4, 16) if (A % 64 == 0) else (0, 0)
    # reference: the multi-tensor update, then the two pack launches
    w0, m0, v0 = w.clone(), m.clone(), v.clone()
    pkg._lib.call("b200_adamw_multi", 1, optim._ptr_array([w0]), optim._ptr_array([gr]), optim._ptr_array([m0]),
                  optim._ptr_array([v0]), optim._numel_array([w0]), *hyper, 3, sq, 1.0, ops._st())
    f0 = torch.empty(taps, A, B, device="cuda", dtype=torch.bfloat16)
    d0 = torch.zeros(taps, B + 5, A, device="cuda", dtype=torch.float32)
    ops._pack(w0, A, B, taps, f0, False, False, A * B, B, *perm)
    ops._pack(w0, A, B, taps, d0, True, True, (B + 5) * A, A)
    # fused
    w1, m1, v1 = w.clone(), m.clone(), v.clone()
    f1, d1 = torch.empty_like(f0), torch.zeros_like(d0)
    pkg._lib.call("b200_adamw_pack", w1, gr, m1, v1, A, B, taps, *hyper, 3, sq, 1.0,
                  f1, 0, 0, 0, A * B, B, *perm, d1, 1, 1, 1, (B + 5) * A, A, 0, 0, ops._st())
    for a, b in ((w1, w0), (m1, m0), (v1, v0), (f1, f0), (d1, d0)):
        assert torch.equal(a, b)
    assert not torch.equal(w1, w)
    # one destination only
    w2, m2, v2, f2 = w.clone(), m.clone(), v.clone(), torch.empty_like(f0)
    pkg._lib.call("b200_adamw_pack", w2, gr, m2, v2, A, B, taps, *hyper, 3, sq, 1.0,
                  f2, 0, 0, 0, A * B, B, *perm, None, 0, 0, 0, 0, 1, 0, 0, ops._st())
    assert torch.equal(w2, w0) and torch.equal(f2, f0)


def test_adamw_pack_multi_tensor_launch():
    """b200_adamw_pack_multi on 20 weights of different shapes at once (two launches of the tile kernel: 16 + 4), one or two
    packed destinations each, against the multi-tensor update followed by one pack launch per destination: bit-identical."""
    import unet_convlstm_b200 as pkg
    from unet_convlstm_b200 import ops, optim
    g = torch.Generator(device="cuda").manual_seed(77)
    shapes = [(64, 16, 9), (16, 64, 9), (33, 5, 9), (128, 128, 4), (256, 40, 9), (8, 8, 1), (70, 33, 9), (64, 64, 9),
              (32, 96, 4), (100, 3, 9), (2, 64, 9), (192, 64, 9), (64, 2, 9), (31, 31, 9), (48, 48, 1), (512, 128, 9),
              (65, 65, 9), (128, 16, 9), (24, 200, 4), (320, 64, 9)]
    hyper = (3e-3, 0.9, 0.999, 1e-8, 1e-2)
    ws = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    grs = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    ms = [0.1 * torch.randn(s, device="cuda", generator=g) for s in shapes]
    vs = [torch.rand(s, device="cuda", generator=g) for s in shapes]
    sq = optim.grad_sqnorm(grs)

    def dests(i):
        A, B, taps = shapes[i]
        perm = (A // 4, 16) if A % 64 == 0 else (0, 0)
        f = torch.full((taps, A, B), float("nan"), device="cuda", dtype=torch.bfloat16)
        geoms = [(A, B, taps, f, 0, 0, 0, A * B, B, *perm)]
        if i % 3:   # two of three weights also have a data-gradient copy (fp32, transposed, taps flipped)
            d = torch.full((taps, B, A), float("nan"), device="cuda", dtype=torch.float32)
            geoms.append((A, B, taps, d, 1, 1, 1, B * A, A, 0, 0))
        return geoms

    # reference
    w0, m0, v0 = [w.clone() for w in ws], [m.clone() for m in ms], [v.clone() for v in vs]
    pkg._lib.call("b200_adamw_multi", len(ws), optim._ptr_array(w0), optim._ptr_array(grs), optim._ptr_array(m0),
                  optim._ptr_array(v0), optim._numel_array(w0), *hyper, 5, sq, 1.0, ops._st())
    ref = [dests(i) for i in range(len(ws))]
    for i, geoms in enumerate(ref):
        for gm in geoms:
            ops._pack(w0[i], gm[0], gm[1], gm[2], gm[3], bool(gm[5]), bool(gm[6]), gm[7], gm[8], gm[9], gm[10])
    # fused, all 20 in one call
    w1, m1, v1 = [w.clone() for w in ws], [m.clone() for m in ms], [v.clone() for v in vs]
    got = [dests(i) for i in range(len(ws))]
    n0 = pkg._lib.CALLS.get("b200_adamw_pack_multi", 0)
    optim.AdamW._update_and_pack([((w1[i], grs[i], m1[i], v1[i]), got[i]) for i in range(len(ws))], hyper, 5, sq, 1.0)
    assert pkg._lib.CALLS.get("b200_adamw_pack_multi", 0) == n0 + 1
    for i in range(len(ws)):
        assert torch.equal(w1[i], w0[i]) and torch.equal(m1[i], m0[i]) and torch.equal(v1[i], v0[i]), shapes[i]
        for a, b in zip(got[i], ref[i]):
            assert torch.equal(a[3], b[3]), shapes[i]   # (no NaN left: every element of every destination was written)


@pytest.mark.parametrize("N,K,ks,kpad", [(64, 64, 3, None), (40, 18, 3, 32), (128, 2, 3, 16), (96, 200, 1, None),
                                         (4096, 2048, 3, None)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_weight_pack_kernels_exact(N, K, ks, kpad, dtype):
    """b200_pack_weight / b200_unpack_wgrad (tiled through shared memory) against the index formulas they
    implement, written with torch views: pure data movement (+ one bf16 rounding), so bit-exact.  Ragged tiles
    (A, B not multiples of 32), zero-padded K, 1x1 and 3x3 taps, and the 75 M-parameter cell weight."""
    from unet_convlstm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N + K)
    w = torch.randn(N, K, ks, ks, device="cuda", generator=g)
    taps, Kp = ks * ks, (K if kpad is None else kpad)
    wt = w.reshape(N, K, taps)
    ref = torch.zeros(taps, N, Kp, device="cuda", dtype=dtype)
    ref[:, :, :K] = wt.permute(2, 0, 1).to(dtype)
    assert torch.equal(ops.pack_conv_weight(w, dtype, kpad), ref)
    ref = torch.zeros(taps, Kp, N, device="cuda", dtype=dtype)
    ref[:, :K, :] = wt.permute(2, 1, 0).flip(0).to(dtype)
    assert torch.equal(ops.pack_conv_weight_dgrad(w, dtype, kpad), ref)
    if N % 64 == 0 and kpad is None:
        Ch = N // 4
        cht = ops.lstm_cht(Ch)
        b = torch.randn(N, device="cuda", generator=g)
        got, gb = ops.pack_lstm_weight(w, b, dtype)
        ref = wt.reshape(4, Ch // cht, cht, K, taps).permute(4, 1, 0, 2, 3).reshape(taps, N, K).to(dtype)
        assert torch.equal(got, ref)
        assert torch.equal(gb, b.reshape(4, Ch // cht, cht).permute(1, 0, 2).reshape(N))
    dw = torch.randn(taps, N, Kp, device="cuda", generator=g)
    assert torch.equal(ops.unpack_conv_wgrad(dw, K), dw[:, :, :K].permute(1, 2, 0).reshape(N, K, ks, ks))


@pytest.mark.parametrize("Cin,Cout", [(128, 64), (40, 24), (1024, 512)])
def test_convT_weight_pack_exact(Cin, Cout):
    from unet_convlstm_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(Cin)
    w = torch.randn(Cin, Cout, 2, 2, device="cuda", generator=g)
    fwd, bwd = ops.pack_convT_weight(w, torch.bfloat16)
    wt = w.reshape(Cin, Cout, 4)
    assert torch.equal(fwd, wt.permute(2, 1, 0).reshape(1, 4 * Cout, Cin).bfloat16())
    assert torch.equal(bwd, wt.permute(0, 2, 1).reshape(1, Cin, 4 * Cout).bfloat16())


def _torch_convlstm(w, b, xs, ks):
    """Plain PyTorch fp64 reference of ConvLSTM (one layer, zero initial state): unet.py:21-36 over the t loop."""
    Ch = w.shape[0] // 4
    B, _, H, W = xs[0].shape
    h = torch.zeros(B, Ch, H, W, dtype=torch.float64)
    c = torch.zeros_like(h)
    outs = []
    for x in xs:
        gates = torch.nn.functional.conv2d(torch.cat([x, h], dim=1), w, b, padding=ks // 2)
        i, f, g, o = torch.chunk(gates, 4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return outs, (h, c)


@pytest.mark.parametrize("ks,cin,ch", [(5, 6, 8), (1, 16, 16), (5, 16, 32)])
def test_convlstm_other_kernel_sizes(ks, cin, ch):
    """kernel_size 5 (25 taps: beyond the tiled weight-pack kernel, generic strided pack) and 1, fp32 check mode
    and bf16 mode, against a plain PyTorch fp64 ConvLSTM: outputs, final state, input and parameter gradients."""
    import unet_convlstm_b200 as pkg
    from train.unet import ConvLSTM
    torch.manual_seed(ks + ch)
    m = ConvLSTM(cin, ch, num_layers=1, kernel_size=ks)
    w = m.layers[0].conv.weight.detach().double().requires_grad_(True)
    b = m.layers[0].conv.bias.detach().double().requires_grad_(True)
    B, T, H, W = 2, 3, 16, 16
    g = torch.Generator().manual_seed(ks)
    xs = [torch.randn(B, cin, H, W, generator=g) for _ in range(T)]
    dout = [torch.randn(B, ch, H, W, generator=g) for _ in range(T)]
    xr = [x.double().requires_grad_(True) for x in xs]
    outs, (hT, cT) = _torch_convlstm(w, b, xr, ks)
    (sum((o * d.double()).sum() for o, d in zip(outs, dout)) + cT.sum()).backward()
    m = m.cuda()
    try:
        for mode, tol_y, tol_g in (("fp32", 1e-5, 1e-4), ("bf16", 2e-2, 3e-2)):
            pkg.set_precision(mode)
            m.zero_grad(set_to_none=True)
            xg = [x.cuda().requires_grad_(True) for x in xs]
            out, st = m(xg)
            (sum((o * d.cuda()).sum() for o, d in zip(out, dout)) + st[0][1].sum()).backward()
            assert rel2(_np(torch.stack(out)), torch.stack(outs).detach().numpy()) < tol_y, mode
            assert rel2(_np(st[0][0]), hT.detach().numpy()) < tol_y and rel2(_np(st[0][1]), cT.detach().numpy()) < tol_y
            assert rel2(np.stack([_np(x.grad) for x in xg]), np.stack([x.grad.numpy() for x in xr])) < tol_g, mode
            assert rel2(_np(m.layers[0].conv.weight.grad), w.grad.numpy()) < tol_g, mode
            assert rel2(_np(m.layers[0].conv.bias.grad), b.grad.numpy()) < tol_g, mode
    finally:
        pkg.set_precision("bf16")


# ------------------------------------------------------------------------------------------------
# BatchNorm + ReLU + 2x2 max-pool in one pass (encoder outputs: skip connection + next Down stage)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(3, 2, 8, 12, 16), (2, 3, 6, 4, 24), (2, 2, 4, 4, 7), (1, 1, 2, 2, 64)])
@pytest.mark.parametrize("training", [True, False])
def test_bn_relu_pool_fused_matches_separate_kernels(mode, shape, training):
    """b200_bn_relu_apply_pool / b200_bn_relu_pool_bwd_reduce / _apply against b200_bn_relu_apply + b200_maxpool2_fwd and
    b200_maxpool2_bwd (accumulating into the skip gradient) + b200_bn_relu_bwd_reduce / _apply: the forward outputs and
    the gradient that reaches the BatchNorm backward are bit-identical by construction; the fp64 sums are accumulated in
    a different order, so dz / dgamma / dbeta agree to rounding.  Vector (C % 8 == 0) and scalar channel counts, ties
    (whole windows clamped to zero by the ReLU), with and without a skip gradient."""
    from unet_convlstm_b200 import ops
    T, B, H, W, C = shape
    dt = ops.act_dtype()
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.randn(shape, device="cuda", generator=g).to(dt)
    z[:, :, : H // 2, : W // 2] -= 3.0      # a quadrant where most windows are all-zero after the ReLU (ties)
    gamma = (0.5 + torch.rand(C, device="cuda", generator=g))
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    dy = torch.randn(shape, device="cuda", generator=g).to(dt)
    dp = torch.randn((T, B, H // 2, W // 2, C), device="cuda", generator=g).to(dt)

    def stats():
        rm = torch.zeros(C, device="cuda")
        rv = torch.ones(C, device="cuda")
        return rm, rv

    rm, rv = stats()
    (y1, p1), st1 = ops.bn_relu_fwd(z, gamma, beta, rm, rv, training, 1e-5, 0.1, pool=True)
    rm2, rv2 = stats()
    y0, st0 = ops.bn_relu_fwd(z, gamma, beta, rm2, rv2, training, 1e-5, 0.1)
    p0 = ops.maxpool2_fwd(y0)
    assert torch.equal(y1, y0) and torch.equal(p1, p0)
    assert torch.equal(rm, rm2) and torch.equal(rv, rv2)

    for skip in (True, False):
        g0 = ops.maxpool2_bwd(y0, dp, accumulate_into=dy.clone()) if skip else ops.maxpool2_bwd(y0, dp)
        dz0, dg0, db0, dc0 = ops.bn_relu_bwd(z, g0, st0, training, True)
        dz1, dg1, db1, dc1 = ops.bn_relu_bwd(z, dy if skip else None, st1, training, True, dpool=dp)
        tol = 1e-5 if mode == "fp32" else 6e-3     # bf16: an occasional last-bit difference of the stored dz
        assert rel(_np(dz1), _np(dz0)) < tol, (skip, rel(_np(dz1), _np(dz0)))
        assert rel2(_np(dz1), _np(dz0)) < (1e-6 if mode == "fp32" else 1e-3)
        for a, b in ((dg1, dg0), (db1, db0), (dc1, dc0)):
            assert np.abs(_np(a) - _np(b)).max() <= 1e-5 * max(np.abs(_np(b)).max(), 1.0)


@pytest.mark.parametrize("size", [(16, 16), (48, 40)])
def test_model_fused_pool_matches_pool_fork(mode, size):
    """TemporalUNetDualView with the fused BatchNorm/ReLU/max-pool passes against the same model on the separate kernels
    (ops.FUSE_BN_POOL off -> PoolFork): outputs bit-identical, every parameter gradient equal to summation-order noise
    (48x40 reaches the odd size 6x5 at the fourth level: that stage falls back to PoolFork by itself)."""
    from train.unet import TemporalUNetDualView
    from unet_convlstm_b200 import ops
    H, W = size
    torch.manual_seed(3)
    m = TemporalUNetDualView(base_ch=8, use_skip_lstm=True).cuda()
    x = torch.randn(2, 3, 2, H, W, device="cuda")
    w = torch.randn(3, 2, 1, H, W, device="cuda")
    res = []
    old = ops.FUSE_BN_POOL
    try:
        for fuse in (True, False):
            ops.FUSE_BN_POOL = fuse
            sd = {k: v.clone() for k, v in m.state_dict().items()}
            m.zero_grad(set_to_none=True)
            n0 = dict(__import__("unet_convlstm_b200")._lib.CALLS)
            out, _ = m(x)
            y = torch.stack(out, 0)
            (y * w).sum().backward()
            calls = {k: v - n0.get(k, 0) for k, v in __import__("unet_convlstm_b200")._lib.CALLS.items()}
            res.append((y.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()},
                        {k: v.clone() for k, v in m.state_dict().items()}, calls))
            m.load_state_dict(sd)
    finally:
        ops.FUSE_BN_POOL = old
    (y1, g1, s1, c1), (y0, g0, s0, c0) = res
    n_fused = 4 if size == (16, 16) else 3
    assert c1.get("b200_bn_relu_apply_pool", 0) == n_fused and c1.get("b200_bn_relu_pool_bwd_apply", 0) == n_fused
    assert c1.get("b200_maxpool2_bwd", 0) == 4 - n_fused
    assert c0.get("b200_bn_relu_apply_pool", 0) == 0 and c0.get("b200_maxpool2_bwd", 0) == 4
    assert torch.equal(y1, y0)
    for k in s0:
        assert torch.equal(s1[k], s0[k]), k
    for k in g0:
        a, b = _np(g1[k]), _np(g0[k])
        scale = max(np.abs(b).max(), 1e-6)
        assert np.abs(a - b).max() <= (1e-4 if mode == "fp32" else 3e-2) * scale, (k, np.abs(a - b).max() / scale)


# ------------------------------------------------------------------------------------------------
# BatchNorm + ReLU + 1x1 OutConv in one pass (the last DoubleConv feeding outc)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(3, 2, 8, 12, 64), (2, 3, 5, 7, 16), (2, 1, 4, 4, 8), (1, 2, 3, 3, 32)])
@pytest.mark.parametrize("training", [True, False])
def test_bn_relu_outconv_fused_matches_separate_kernels(mode, shape, training):
    """b200_bn_relu_outconv_fwd / _bwd_reduce / _bwd_apply against b200_bn_relu_apply + b200_outconv_fwd and
    b200_outconv_bwd + b200_bn_relu_bwd_reduce / _apply.  fp32 mode: only summation orders differ (1e-5).  bf16 mode: the
    separate kernels round the activation and the data gradient of the 1x1 convolution to bf16 on their way through HBM,
    the fused kernels never store either and do not round them: the two agree to that rounding."""
    from unet_convlstm_b200 import ops
    T, B, H, W, C = shape
    dt = ops.act_dtype()
    if not ops.bn_outconv_ok(C, dt, 1):
        pytest.skip("channel count not covered by the fused kernels in this mode")
    g = torch.Generator(device="cuda").manual_seed(11)
    z = torch.randn(shape, device="cuda", generator=g).to(dt)
    gamma = (0.5 + torch.rand(C, device="cuda", generator=g))
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    w = torch.randn(1, C, device="cuda", generator=g) / C ** 0.5
    b = torch.randn(1, device="cuda", generator=g)
    dout = torch.randn((T, B, H, W, 1), device="cuda", generator=g)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    out1, st1 = ops.bn_relu_fwd(z, gamma, beta, rm, rv, training, 1e-5, 0.1, outconv=(w.reshape(-1), b))
    rm2, rv2 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    y0, st0 = ops.bn_relu_fwd(z, gamma, beta, rm2, rv2, training, 1e-5, 0.1)
    out0 = ops.outconv_fwd(y0, w, b)
    assert out1.shape == out0.shape and out1.dtype == torch.float32
    assert rel(_np(out1), _np(out0)) < (1e-5 if mode == "fp32" else 5e-3)
    if mode == "bf16":   # ... and the fused output is the one closer to exact arithmetic on the same operands
        exact = (torch.relu(z.double() * st0[2].double().view(-1, 1, 1, 1, C) + st0[3].double().view(-1, 1, 1, 1, C))
                 * w.double().view(1, 1, 1, 1, C)).sum(-1, keepdim=True) + b.double()
        assert rel2(_np(out1), _np(exact)) <= rel2(_np(out0), _np(exact)) + 1e-6
    assert torch.equal(rm, rm2) and torch.equal(rv, rv2)

    dy0, dw0, db0 = ops.outconv_bwd(y0, w, dout)
    dz0, dg0, dbeta0, dc0 = ops.bn_relu_bwd(z, dy0, st0, training, True)
    dz1, dg1, dbeta1, dc1, dw1 = ops.bn_relu_bwd(z, dout, st1, training, True, outconv_w=w.reshape(-1))
    db1 = ops.colsum(dout.numel(), dout, 1)
    tol = 1e-5 if mode == "fp32" else 1.2e-2
    assert rel(_np(dz1), _np(dz0)) < tol, rel(_np(dz1), _np(dz0))
    assert rel2(_np(dz1), _np(dz0)) < (1e-6 if mode == "fp32" else 4e-3)
    for a, ref in ((dg1, dg0), (dbeta1, dbeta0), (dc1, dc0), (dw1, dw0.reshape(-1)), (db1, db0)):
        # (bf16: the separate path's gradient is rounded to bf16 element by element, 4e-3 each; over the 18 .. 1152 pixels of
        # these problems that does not average out below ~1e-2)
        assert np.abs(_np(a) - _np(ref)).max() <= (1e-5 if mode == "fp32" else 1.5e-2) * max(np.abs(_np(ref)).max(), 1.0)


def test_model_fused_outconv_matches_separate(mode):
    """TemporalUNetDualView with the last DoubleConv + OutConv fused against the separate kernels (ops.FUSE_BN_OUTCONV
    off): output, every parameter gradient (outc.conv.weight / bias included) and the input gradient."""
    from train.unet import TemporalUNetDualView
    from unet_convlstm_b200 import ops
    import unet_convlstm_b200 as pkg
    torch.manual_seed(4)
    m = TemporalUNetDualView(base_ch=16 if mode == "bf16" else 8, use_skip_lstm=False).cuda()
    x = torch.randn(2, 2, 2, 16, 16, device="cuda")
    w = torch.randn(2, 2, 1, 16, 16, device="cuda")
    res = []
    old = ops.FUSE_BN_OUTCONV
    try:
        for fuse in (True, False):
            ops.FUSE_BN_OUTCONV = fuse
            sd = {k: v.clone() for k, v in m.state_dict().items()}
            m.zero_grad(set_to_none=True)
            n0 = dict(pkg._lib.CALLS)
            xi = x.clone().requires_grad_(True)
            out, _ = m(xi)
            y = torch.stack(out, 0)
            (y * w).sum().backward()
            calls = {k: v - n0.get(k, 0) for k, v in pkg._lib.CALLS.items()}
            res.append((y.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()}, xi.grad.clone(), calls))
            m.load_state_dict(sd)
    finally:
        ops.FUSE_BN_OUTCONV = old
    (y1, g1, dx1, c1), (y0, g0, dx0, c0) = res
    assert c1.get("b200_bn_relu_outconv_fwd", 0) == 1 and c1.get("b200_outconv_fwd", 0) == 0
    assert c0.get("b200_bn_relu_outconv_fwd", 0) == 0 and c0.get("b200_outconv_fwd", 0) == 1
    assert rel(_np(y1), _np(y0)) < (1e-5 if mode == "fp32" else 5e-3)
    g1["x"], g0["x"] = dx1, dx0
    for k in g0:
        a, b = _np(g1[k]), _np(g0[k])
        scale = max(np.abs(b).max(), 1e-6)
        assert np.abs(a - b).max() <= (1e-4 if mode == "fp32" else 3e-2) * scale, (k, np.abs(a - b).max() / scale)


def test_image_too_small_for_four_down_stages_raises():
    """Like the reference (nn.MaxPool2d raises 'Output size is too small'): a 20x12 image is 2x1 at the fourth Down stage."""
    from train.unet import TemporalUNetDualView
    m = TemporalUNetDualView(base_ch=8).cuda()
    with pytest.raises(RuntimeError, match="too small"):
        m(torch.randn(1, 2, 2, 20, 12, device="cuda"))
