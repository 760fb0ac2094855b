"""The numpy oracle (oracle/unet_oracle.py) against fixtures produced by the reference itself
(tests/golden/make_golden.py): forward, hand-derived backward, BatchNorm buffers, state carry."""
import glob
import os

import numpy as np
import pytest

from oracle import unet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 2e-6  # fixtures store gradients as fp32; everything is computed in fp64


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    p = {k[2:]: z[k].astype(np.float64) if z[k].dtype != np.int64 else z[k] for k in z.files if k.startswith("p.")}
    return z, p


@pytest.mark.parametrize("name", ["convlstm_c8_l1_zero.npz", "convlstm_c6_12_l2_state.npz", "convlstm_c16_l1_state.npz"])
def test_convlstm_matches_reference(golden_dir, name):
    z, p = load(golden_dir, name)
    cin, ch, L, B, T, H, W, with_state = z["meta"]
    layers = [(p[f"layers.{l}.conv.weight"], p[f"layers.{l}.conv.bias"]) for l in range(L)]
    state = [(z[f"h0{l}"], z[f"c0{l}"]) for l in range(L)] if with_state else None
    out, new_state, caches = O.convlstm_fwd(list(z["x"]), layers, state)
    assert rel(np.stack(out), z["out"]) < 1e-10
    for l in range(L):
        assert rel(new_state[l][0], z[f"hT{l}"]) < 1e-10
        assert rel(new_state[l][1], z[f"cT{l}"]) < 1e-10
    dstate = [None] * (L - 1) + [(z["dh_last"], z["dc_last"])]
    dx, wg, d0 = O.convlstm_bwd(caches, list(z["dout"]), dstate)
    assert rel(np.stack(dx), z["dx"]) < 1e-10
    for l in range(L):
        assert rel(wg[l][0], z[f"g.layers.{l}.conv.weight"]) < TOL
        assert rel(wg[l][1], z[f"g.layers.{l}.conv.bias"]) < TOL
        if with_state:
            assert rel(d0[l][0], z[f"dh0{l}"]) < 1e-10
            assert rel(d0[l][1], z[f"dc0{l}"]) < 1e-10


@pytest.mark.parametrize("name,kind", [("double_3_8.npz", "double"), ("down_8_16.npz", "down"),
                                       ("up_16_8.npz", "up"), ("up_16_8_pad.npz", "up")])
def test_blocks_match_reference(golden_dir, name, kind):
    z, p = load(golden_dir, name)
    tp = O.Tape(p, training=True)
    if kind == "double":
        y, c = O.double_conv_fwd(tp, "net", z["x0"])
        dxs = (O.double_conv_bwd(tp, "net", c, z["dy"]),)
    elif kind == "down":
        y, c = O.down_fwd(tp, "", z["x0"]) if False else O.down_fwd(_Prefixed(tp), "X", z["x0"])
        dxs = (O.down_bwd(_Prefixed(tp), "X", c, z["dy"]),)
    else:
        y, c = O.up_fwd(_Prefixed(tp), "X", z["x0"], z["x1"])
        dxs = O.up_bwd(_Prefixed(tp), "X", c, z["dy"])
    assert rel(y, z["y_train"]) < 1e-10
    for i, dx in enumerate(dxs):
        assert rel(dx, z[f"dx{i}"]) < 1e-9
    for k in z.files:
        if k.startswith("g."):
            g = tp.grads[k[2:]]
            ref = z[k]
            # conv biases feeding a BatchNorm have an exactly-zero true gradient: compare absolutely
            if np.abs(ref).max() < 1e-9:
                assert np.abs(g).max() < 1e-9
            else:
                assert rel(g, ref) < TOL, k
        if k.startswith("after."):
            assert rel(np.asarray(tp.buf(k[6:]), dtype=np.float64), z[k].astype(np.float64)) < 1e-12, k
    tpe = O.Tape({**p, **{k: v for k, v in tp.new_buffers.items()}}, training=False)
    if kind == "double":
        ye, _ = O.double_conv_fwd(tpe, "net", z["x0"])
    elif kind == "down":
        ye, _ = O.down_fwd(_Prefixed(tpe), "X", z["x0"])
    else:
        ye, _ = O.up_fwd(_Prefixed(tpe), "X", z["x0"], z["x1"])
    assert rel(ye, z["y_eval"]) < 1e-10


class _Prefixed:
    """Lets the block-level oracle functions (which expect `<pre>.` keys) run on a bare block's
    state_dict: strips the dummy prefix 'X.'."""

    def __init__(self, tp):
        self._tp = tp
        self.training = tp.training
        self.p = _StripDict(tp.p)
        self.new_buffers = _StripDict(tp.new_buffers)
        self.grads = tp.grads

    def buf(self, name):
        return self._tp.buf(name[2:])

    def add(self, name, g):
        self._tp.add(name[2:], g)


class _StripDict:
    def __init__(self, d):
        self.d = d

    def __getitem__(self, k):
        return self.d[k[2:]]

    def __setitem__(self, k, v):
        self.d[k[2:]] = v

    def __contains__(self, k):
        return k[2:] in self.d

    def get(self, k, default=None):
        return self.d.get(k[2:], default)


@pytest.mark.parametrize("name", ["model_b4_skip.npz", "model_b2_noskip_l2.npz", "model_b4_skip_64.npz"])
def test_model_matches_reference(golden_dir, name):
    z, p = load(golden_dir, name)
    base_ch, skip, L, B, T, H, W = z["meta"]
    y, st, tp, caches = O.temporal_unet_fwd(p, z["x"], training=True)
    assert rel(y, z["y_train"]) < 1e-9
    for l in range(L):
        assert rel(st[l][0], z[f"hT{l}"]) < 1e-9
        assert rel(st[l][1], z[f"cT{l}"]) < 1e-9
    dx = O.temporal_unet_bwd(tp, caches, z["dy"])
    assert rel(dx, z["dx"]) < 1e-7
    for k in z.files:
        if k.startswith("g."):
            ref = z[k].astype(np.float64)
            g = tp.grads[k[2:]]
            if np.abs(ref).max() < 1e-7 * max(1.0, np.abs(z["dy"]).max()):
                assert np.abs(g).max() < 1e-6, k
            else:
                assert rel(g, ref) < 5e-6, k
        if k.startswith("after."):
            assert rel(np.asarray(tp.buf(k[6:]), dtype=np.float64), z[k].astype(np.float64)) < 1e-10, k
    # eval mode with the updated running statistics, and the state round trip (unet.py:185)
    p_eval = {**p, **tp.new_buffers}
    ye, _, _, _ = O.temporal_unet_fwd(p_eval, z["x"], training=False)
    assert rel(ye, z["y_eval"]) < 1e-9
    k = T // 2
    y1, s1, _, _ = O.temporal_unet_fwd(p_eval, z["x"][:, :k], training=False)
    y2, _, _, _ = O.temporal_unet_fwd(p_eval, z["x"][:, k:], state=s1, training=False)
    assert rel(np.concatenate([y1, y2], axis=1), z["y_eval_split"]) < 1e-9
    assert int(tp.buf("inc.net.1.num_batches_tracked")) == T


def test_fixture_set_is_complete(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "*.npz"))) >= 10


# --------------------------------------------------------------------------------------------
# oracle/torch_port.py (the CPU baseline of bench.py) against the same fixtures
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["model_b4_skip.npz", "model_b2_noskip_l2.npz", "model_b4_skip_64.npz"])
def test_torch_port_model_matches_reference(golden_dir, name):
    import torch
    from oracle import torch_port as TP
    z = np.load(os.path.join(golden_dir, name))
    base_ch, skip, L, B, T, H, W = z["meta"]
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
    p = TP.params_from_state_dict(sd, torch.float64)
    x = torch.from_numpy(z["x"]).double().requires_grad_(True)
    out, st = TP.temporal_unet(p, x, None, training=True)
    y = torch.stack(out, dim=1)
    assert rel(y.detach().numpy(), z["y_train"]) < 1e-9
    (y * torch.from_numpy(z["dy"]).double()).sum().backward()
    assert rel(x.grad.numpy(), z["dx"]) < 1e-7
    for k in z.files:
        if k.startswith("g."):
            ref = z[k].astype(np.float64)
            g = p[k[2:]].grad.numpy()
            if np.abs(ref).max() < 1e-7 * max(1.0, np.abs(z["dy"]).max()):
                assert np.abs(g).max() < 1e-6, k
            else:
                assert rel(g, ref) < 5e-6, k
        if k.startswith("after."):
            assert rel(p[k[6:]].double().numpy(), z[k].astype(np.float64)) < 1e-10, k
    for l in range(L):
        assert rel(st[l][0].detach().numpy(), z[f"hT{l}"]) < 1e-9
    with torch.no_grad():
        oe, _ = TP.temporal_unet(p, x.detach(), None, training=False)
    assert rel(torch.stack(oe, dim=1).numpy(), z["y_eval"]) < 1e-9


@pytest.mark.parametrize("name", ["convlstm_c6_12_l2_state.npz", "convlstm_c16_l1_state.npz"])
def test_torch_port_convlstm_matches_reference(golden_dir, name):
    import torch
    from oracle import torch_port as TP
    z = np.load(os.path.join(golden_dir, name))
    cin, ch, L, B, T, H, W, with_state = z["meta"]
    p = {"m." + k[2:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("p.")}
    state = [(torch.from_numpy(z[f"h0{l}"]).double(), torch.from_numpy(z[f"c0{l}"]).double()) for l in range(L)]
    out, ns = TP.convlstm(p, "m", [torch.from_numpy(a).double() for a in z["x"]], state)
    assert rel(torch.stack(out).numpy(), z["out"]) < 1e-10
    for l in range(L):
        assert rel(ns[l][1].numpy(), z[f"cT{l}"]) < 1e-10


def test_loss_oracle_matches_reference_compute_loss(golden_dir):
    """oracle/loss_oracle.py against the fixture produced by the reference's own main.compute_loss."""
    from oracle import loss_oracle as LO
    z = np.load(os.path.join(golden_dir, "loss_main_compute_loss.npz"))
    for name in ("a", "b", "c", "z"):
        for tag in ("mask", "nomask", "ignored"):
            if f"{name}.{tag}.loss" not in z.files:
                continue
            mask = None if tag == "nomask" else z[f"{name}.mask"]
            loss, grad = LO.compute_loss(z[f"{name}.yp"], z[f"{name}.y"], mask, use_mask=(tag != "ignored"))
            assert abs(loss - float(z[f"{name}.{tag}.loss"])) <= 1e-12 * max(1.0, abs(loss)), (name, tag)
            np.testing.assert_allclose(grad, z[f"{name}.{tag}.grad"], rtol=1e-10, atol=1e-14, err_msg=f"{name}.{tag}")


def test_metrics_oracle_matches_reference_evaluate(golden_dir):
    """oracle/metrics_oracle.py (+ the loss oracle) against the fixture produced by the reference's own
    main.evaluate with the reference's NPZSequenceDataset.denormalize (tests/golden/make_golden_metrics.py)."""
    from oracle import loss_oracle as LO
    from oracle import metrics_oracle as MO
    z = np.load(os.path.join(golden_dir, "metrics_main_evaluate.npz"))
    for tr in ("asinh", "signed_log", "none"):
        tmin, tmax, scale = z[f"{tr}.params"]
        batches = [(z[f"{tr}.b{i}.pred"], z[f"{tr}.b{i}.y"], z[f"{tr}.b{i}.mask"]) for i in range(3)]
        for use in (True, False):
            ref = z[f"{tr}.use{int(use)}.result"]
            mae, rmse, me = MO.epoch_metrics(batches, tmin, tmax, scale, tr, use_mask=use)
            tot = sum(LO.compute_loss(p, y, m, use_mask=use)[0] * p.shape[0] for p, y, m in batches)
            n = sum(p.shape[0] for p, _, _ in batches)
            # the reference accumulates float32 losses (.item()) and float32/64 NumPy lists
            np.testing.assert_allclose([tot / n, mae, rmse, me], ref, rtol=2e-6, atol=2e-7, err_msg=f"{tr} use={use}")
    p, y, m = z["asinh.b1.pred"], z["asinh.b1.y"], np.zeros_like(z["asinh.b1.mask"])
    tmin, tmax, scale = z["asinh.params"]
    assert MO.epoch_metrics([(p, y, m)], tmin, tmax, scale, "asinh") == (0.0, 0.0, 0.0)
    assert tuple(z["empty.result"][1:]) == (0.0, 0.0, 0.0)


def test_optimizer_oracle_matches_torch_clip_and_adamw():
    """oracle/optim_oracle.py against the calls the reference makes (main.py:106, :275), run on the CPU in fp64."""
    import torch
    from oracle import optim_oracle as OO
    rng = np.random.default_rng(9)
    shapes = [(5, 3), (17,), (2, 3, 3, 3)]
    ps = [rng.standard_normal(s) for s in shapes]
    tp = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in ps]
    kw = dict(lr=2e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.03)
    opt = torch.optim.AdamW(tp, **kw)
    m = [np.zeros(s) for s in shapes]
    v = [np.zeros(s) for s in shapes]
    for step in range(1, 5):
        gs = [rng.standard_normal(s) * (5.0 if step == 1 else 0.05) for s in shapes]   # a clipping and non-clipping steps
        for t, g in zip(tp, gs):
            t.grad = torch.tensor(g, dtype=torch.float64)
        total_ref = float(torch.nn.utils.clip_grad_norm_(tp, 1.0))
        total, gc = OO.clip_grad_norm(gs, 1.0)
        assert abs(total - total_ref) <= 1e-12 * total_ref
        for t, g in zip(tp, gc):
            np.testing.assert_allclose(t.grad.numpy(), g, rtol=1e-12, atol=0)
        opt.step()
        ps, m, v = OO.adamw_step(ps, gc, m, v, step, **kw)
        for t, p in zip(tp, ps):
            np.testing.assert_allclose(t.detach().numpy(), p, rtol=1e-11, atol=1e-14)


def test_weight_cache_invalidation_rules():
    """functional.WeightCache on CPU tensors (pure host logic): version bumps, the optimizer-step generation counter
    (torch's fused optimizers do not bump versions), and the single-use bookkeeping that gates the background
    weight-gradient stream."""
    import torch
    from unet_convlstm_b200.functional import WeightCache
    w = torch.nn.Parameter(torch.zeros(4))
    c, built = WeightCache(), []

    def get():
        return c.get("k", (w,), lambda: built.append(1) or len(built))

    c.begin_forward(True)
    assert get() == 1 and get() == 1                      # cached within a step
    assert c.note_backward() is True                      # one use of the weights in this graph
    c.begin_forward(False)                                # backward -> no-grad forward -> optimizer.step() -> forward
    assert get() == 1                                     # (nothing changed yet: still cached)
    opt = torch.optim.AdamW([w], lr=0.1, fused=False)
    w.grad = torch.ones(4)
    opt.step()
    c.begin_forward(True)
    assert get() == 2                                     # any optimizer step invalidates, version bump or not
    assert c.note_backward() is True
    c.begin_forward(False)
    assert get() == 2
    c.begin_forward(False)
    assert get() == 2                                     # inference after training keeps its packed weights
    with torch.no_grad():
        w.add_(1.0)                                       # an in-place torch op bumps the version
    assert get() == 3
    from unet_convlstm_b200 import functional as Fn
    g0 = Fn._OPT_GENERATION[0]
    torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1).step()   # somebody else's optimizer: conservative
    assert Fn._OPT_GENERATION[0] == g0 + 1
    assert get() == 4
    c.begin_forward(True), c.begin_forward(True)          # the module called twice in one graph
    assert c.note_backward() is False and c.note_backward() is False
    c.begin_forward(True)
    assert c.note_backward() is True                      # and alone again


def test_pack_refresh_plan_and_commit():
    """functional.pack_refresh_plan / pack_refresh_commit (host logic behind optim.AdamW's b200_adamw_pack route): an
    entry is offered to the optimizer only while it is valid and only if every recipe derives from the parameters being
    updated; after the commit it survives the optimizer's own generation bump, and nothing else."""
    import torch
    from unet_convlstm_b200 import functional as Fn, ops
    w, b, other = (torch.nn.Parameter(torch.zeros(4)) for _ in range(3))
    c, built, redone = Fn.WeightCache(), [], []

    def builder():
        built.append(1)
        ops._PACK_RECORDER.append(("pack", w, (2, 2, 1, "out", 0, 0, 0, 4, 2, 0, 0)))
        ops._PACK_RECORDER.append(("redo", (b,), lambda: redone.append(1)))
        return len(built)

    get = lambda: c.get("lstm", (w, b), builder)
    assert get() == 1 and ops._PACK_RECORDER is None
    upd = {p.data_ptr(): p for p in (w, b, other)}
    packs, redo, entries = Fn.pack_refresh_plan(upd)
    assert list(packs) == [w.data_ptr()] and packs[w.data_ptr()][0][3] == "out" and len(redo) == 1 and len(entries) == 1
    # an optimizer that does not update the bias cannot keep this entry fresh
    assert Fn.pack_refresh_plan({w.data_ptr(): w}) == ({}, [], [])
    # the "update": versions move, the closures run, the entry is re-stamped, the post-step hook bumps the generation
    with torch.no_grad():
        w.add_(1.0), b.add_(1.0)
    for fn in redo:
        fn()
    Fn.pack_refresh_commit(entries)
    Fn._bump_generation()
    assert get() == 1 and redone == [1]                    # still cached: no rebuild after the optimizer step
    Fn._bump_generation()                                  # somebody else's optimizer stepped
    assert get() == 2
    with torch.no_grad():
        w.add_(1.0)                                        # stale entry (version moved): not offered
    assert Fn.pack_refresh_plan(upd) == ({}, [], [])
    assert get() == 3
    # entries without recipes (e.g. the folded-BatchNorm coefficients) are never offered
    c2 = Fn.WeightCache()
    c2.get("bn_eval", (other,), lambda: 0)
    assert all(e[1] != 0 for e in Fn.pack_refresh_plan(upd)[2])


def test_background_stream_guards_on_leaf_state():
    """ops._leaf_takes_gradient_as_is: the background path is only safe when autograd merely stores the gradient."""
    import torch
    from unet_convlstm_b200 import ops
    p = torch.nn.Parameter(torch.zeros(3))
    assert ops._leaf_takes_gradient_as_is(p)
    p.grad = torch.zeros(3)
    assert not ops._leaf_takes_gradient_as_is(p)          # accumulation into an existing .grad
    p.grad = None
    h = p.register_hook(lambda g: g)
    assert not ops._leaf_takes_gradient_as_is(p)          # a tensor hook reads the gradient during backward
    h.remove()
    assert ops._leaf_takes_gradient_as_is(p)
    h = p.register_post_accumulate_grad_hook(lambda q: None)
    assert not ops._leaf_takes_gradient_as_is(p)
    p._b200_bg_aware = True                               # dist.GradReducer orders itself behind the stream
    assert ops._leaf_takes_gradient_as_is(p)
    h.remove()
    assert not ops._leaf_takes_gradient_as_is((p * 2))    # not a leaf
    assert not ops.background(p).active                   # outside a backward pass


def test_epoch_loop_batch_staging_order(monkeypatch):
    """loop._device_batches (host logic only): every batch is yielded exactly once and in order, the copy of
    batch i+1 is started BEFORE batch i is handed to the step, device batches pass through untouched, an empty
    loader yields nothing."""
    from unet_convlstm_b200 import loop

    class FakeTensor:
        def __init__(self, tag, cuda=False):
            self.tag, self.is_cuda = tag, cuda

        def is_pinned(self):
            return True

    events = []

    class FakePrefetcher:
        def __init__(self, device):
            self.pending = None

        def start(self, *ts):
            events.append(("start", ts[0].tag))
            self.pending = ts

        def get(self):
            ts, self.pending = self.pending, None
            events.append(("get", ts[0].tag))
            return [FakeTensor(t.tag, cuda=True) for t in ts]

    monkeypatch.setattr(loop, "DevicePrefetcher", FakePrefetcher)
    host = [(FakeTensor(i), FakeTensor(i), FakeTensor(i)) for i in range(4)]
    seen = []
    for x, y, m in loop._device_batches(host, "cuda"):
        events.append(("step", x.tag))
        seen.append(x.tag)
        assert x.is_cuda
    assert seen == [0, 1, 2, 3]
    assert events == [("start", 0), ("get", 0), ("start", 1), ("step", 0), ("get", 1), ("start", 2), ("step", 1),
                      ("get", 2), ("start", 3), ("step", 2), ("get", 3), ("step", 3)]
    events.clear()
    dev = [(FakeTensor(i, True), FakeTensor(i, True), FakeTensor(i, True)) for i in range(2)]
    assert [b[0].tag for b in loop._device_batches(dev, "cuda")] == [0, 1] and events == []
    assert list(loop._device_batches([], "cuda")) == []


def test_run_reference_launcher_runs_overfit_check_unchanged_cpu(tmp_path):
    """tools/run_reference.py on the CPU with the reference's OWN model (--impl reference): the launcher's shims (smp
    stub, constant overrides, bounded range, synthetic NPZ) let train/overfit_check.py run as shipped."""
    import json
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import run_reference as RR
        try:
            RR.reference_root()
        except RuntimeError:
            pytest.skip("no copy of the reference available")
    finally:
        sys.path.pop(0)
    out = tmp_path / "curve.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference.py"), "overfit", "--impl", "reference",
                        "--iters", "1", "--seq-len", "2", "--size", "16", "--out", str(out)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.loads(out.read_text())
    assert list(res["curve"]) == ["0"] and 0 < res["curve"]["0"] < 10
    assert "Starting Overfit Test (custom)" in r.stdout


def test_maxpool_tie_breaking_matches_aten():
    """The max-pool backward of this repository (maxpool2_bwd_kernel and the fused BatchNorm / ReLU / pool backward) and the
    numpy oracle route the gradient of a window to its FIRST maximum in scan order (0,0),(0,1),(1,0),(1,1).  After a ReLU
    whole windows are exactly zero, so ties are the common case, not a corner: pin the rule against ATen
    (nn.MaxPool2d(2), reference unet.py:81) on windows with every tie pattern, and the oracle against it."""
    import itertools
    import numpy as np
    import torch
    from oracle import unet_oracle as O
    pats = list(itertools.product([0.0, 1.0], repeat=4))            # 16 tie patterns of a 2x2 window
    x = torch.tensor(pats, dtype=torch.float64).reshape(1, 16, 2, 2).clone().requires_grad_(True)
    torch.nn.functional.max_pool2d(x, 2).sum().backward()
    got = x.grad.reshape(16, 4)
    for p, g in zip(pats, got):
        first = p.index(max(p))
        assert g.tolist() == [1.0 if k == first else 0.0 for k in range(4)], (p, g)
    y, cache = O.maxpool2_fwd(x.detach().numpy())
    dx = O.maxpool2_bwd(cache, np.ones_like(y))
    assert np.array_equal(dx.reshape(16, 4), got.numpy())
