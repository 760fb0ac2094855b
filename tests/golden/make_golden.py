"""Generates the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Usage (only in the build container, where /root/reference exists):
    python tests/golden/make_golden.py

Imports the unmodified reference classes from /root/reference/train/unet.py (CPU), runs them in
float64 on seeded inputs and stores inputs, state_dict, outputs and autograd gradients as .npz.
The reference ships no tests / golden vectors of its own, so these files
are what pins the oracle (oracle/unet_oracle.py) and, through it, the CUDA kernels.

Each fixture also records the reference's OWN numerical noise floor on the same inputs:
    e32.<key> = [max-rel, l2-rel] error of the reference run in fp32            vs its fp64 run
    e16.<key> = [max-rel, l2-rel] error of the reference run under bf16 autocast vs its fp64 run
Train-mode BatchNorm backward is ill-conditioned at small batch x spatial sizes (SURVEY.md section 7,
hard part 4): the parity tests accept max(stated tolerance, small multiple of this floor).
"""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from train.unet import ConvLSTM, DoubleConv, Down, TemporalUNetDualView, Up  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_default_dtype(torch.float64)


def round_params_to_fp32(module):
    """Parameters/buffers are made exactly fp32-representable so fixtures can store them as fp32."""
    for t in list(module.parameters()) + list(module.buffers()):
        if t.dtype == torch.float64:
            t.data = t.data.float().double()


def sd_np(module, prefix="p.", only_buffers=False):
    names = {k for k, _ in module.named_buffers()} if only_buffers else None
    out = {}
    for k, v in module.state_dict().items():
        if names is not None and k not in names:
            continue
        a = v.detach().cpu().numpy().copy()
        out[prefix + k] = a.astype(np.float32) if (a.dtype == np.float64 and prefix == "p.") else a
    return out


def randomize_bn(module, gen):
    """Non-trivial BN affine/running stats so that fixtures exercise them."""
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=gen)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=gen)
            m.running_mean.data = 0.1 * torch.randn(m.running_mean.shape, generator=gen)
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=gen)


def _errs(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.array([np.abs(a - b).max() / max(np.abs(b).max(), 1e-300),
                     np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)])


def with_noise_floor(d, module64, run):
    """run(module, dtype, autocast) -> dict key -> float64 ndarray of everything compared by the tests.
    Stores the fp64 results in d and the fp32 / bf16-autocast errors as e32.* / e16.*."""
    pristine = copy.deepcopy(module64)
    ref = run(module64, torch.float64, False)
    d.update(ref)
    r32 = run(copy.deepcopy(pristine).float(), torch.float32, False)
    r16 = run(copy.deepcopy(pristine).float(), torch.float32, True)
    for k, v in ref.items():
        if k in r32:
            d["e32." + k] = _errs(r32[k], v)
            d["e16." + k] = _errs(r16[k], v)
    return d


def _grads(module):
    return {"g." + k: p.grad.detach().double().numpy().copy() for k, p in module.named_parameters()}


def _f(t):
    return t.detach().double().numpy().copy()


def convlstm_fixture(name, cin, ch, layers, B, T, H, W, with_state, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = ConvLSTM(cin, ch, num_layers=layers)
    round_params_to_fp32(m)
    x = torch.randn(T, B, cin, H, W, generator=gen)
    st0 = [(torch.randn(B, ch, H, W, generator=gen) * 0.5, torch.randn(B, ch, H, W, generator=gen) * 0.5)
           for _ in range(layers)] if with_state else None
    wts = torch.randn(T, B, ch, H, W, generator=gen)
    wh = torch.randn(B, ch, H, W, generator=gen)
    wc = torch.randn(B, ch, H, W, generator=gen)
    d = sd_np(m)

    def run(mod, dtype, autocast):
        xs = [x[t].detach().clone().to(dtype).requires_grad_(True) for t in range(T)]
        state = None
        if with_state:
            state = [(h.detach().clone().to(dtype).requires_grad_(True), c.detach().clone().to(dtype).requires_grad_(True))
                     for h, c in st0]
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            out, new_state = mod(xs, state)
        loss = sum((o.to(dtype) * wts[t].to(dtype)).sum() for t, o in enumerate(out))
        loss = loss + (new_state[-1][0].to(dtype) * wh.to(dtype)).sum() + (new_state[-1][1].to(dtype) * wc.to(dtype)).sum()
        loss.backward()
        r = {"out": np.stack([_f(o) for o in out]), "dx": np.stack([_f(v.grad) for v in xs])}
        for l in range(layers):
            r[f"hT{l}"], r[f"cT{l}"] = _f(new_state[l][0]), _f(new_state[l][1])
            if with_state:
                r[f"dh0{l}"], r[f"dc0{l}"] = _f(state[l][0].grad), _f(state[l][1].grad)
        r.update(_grads(mod))
        return r

    with_noise_floor(d, m, run)
    d["x"], d["dout"], d["dh_last"], d["dc_last"] = x.numpy(), wts.numpy(), wh.numpy(), wc.numpy()
    if with_state:
        for l in range(layers):
            d[f"h0{l}"], d[f"c0{l}"] = st0[l][0].numpy(), st0[l][1].numpy()
    d["meta"] = np.array([cin, ch, layers, B, T, H, W, int(with_state)])
    _save(name, d)


def block_fixture(name, kind, cin, cout, B, H, W, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = {"double": DoubleConv, "down": Down, "up": Up}[kind](cin, cout)
    randomize_bn(m, gen)
    round_params_to_fp32(m)
    d = sd_np(m)
    if kind == "up":
        ins = (torch.randn(B, cin, H // 2, W // 2, generator=gen), torch.randn(B, cin // 2, H, W, generator=gen))
    else:
        ins = (torch.randn(B, cin, H, W, generator=gen),)
    oshape = (B, cout, H // 2, W // 2) if kind == "down" else (B, cout, H, W)
    w_ = torch.randn(oshape, generator=gen)

    def run(mod, dtype, autocast):
        args = [a.detach().clone().to(dtype).requires_grad_(True) for a in ins]
        mod.train()
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            y = mod(*args)
        (y.to(dtype) * w_.to(dtype)).sum().backward()
        r = {"y_train": _f(y)}
        for i, a in enumerate(args):
            r[f"dx{i}"] = _f(a.grad)
        r.update(_grads(mod))
        for k, v in sd_np(mod, "after.", only_buffers=True).items():
            r[k] = np.asarray(v, dtype=np.float64) if v.dtype != np.int64 else v
        mod.eval()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            r["y_eval"] = _f(mod(*[a.detach() for a in args]))
        return r

    with_noise_floor(d, m, run)
    d["dy"] = w_.numpy()
    for i, a in enumerate(ins):
        d[f"x{i}"] = a.numpy()
    _save(name, d)


def model_fixture(name, base_ch, skip, layers, B, T, H, W, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = TemporalUNetDualView(base_ch=base_ch, lstm_layers=layers, use_skip_lstm=skip)
    randomize_bn(m, gen)
    round_params_to_fp32(m)
    d = sd_np(m)
    x0 = torch.rand(B, T, 2, H, W, generator=gen)
    w_ = torch.randn(B, T, 1, H, W, generator=gen)

    def run(mod, dtype, autocast):
        x = x0.detach().clone().to(dtype).requires_grad_(True)
        mod.train()
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            out, st = mod(x)
        y = torch.stack(out, dim=1)
        (y.to(dtype) * w_.to(dtype)).sum().backward()
        r = {"y_train": _f(y), "dx": _f(x.grad)}
        for l in range(layers):
            r[f"hT{l}"], r[f"cT{l}"] = _f(st[l][0]), _f(st[l][1])
        r.update(_grads(mod))
        for k, v in sd_np(mod, "after.", only_buffers=True).items():
            r[k] = np.asarray(v, dtype=np.float64) if v.dtype != np.int64 else v
        # eval mode (running stats as updated by the train-mode forward above), plus the state
        # round trip of SURVEY 8c: model(x[:, :k]) then model(x[:, k:], state)
        mod.eval()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            xe = x.detach()
            oe, _ = mod(xe)
            k = T // 2
            o1, s1 = mod(xe[:, :k])
            o2, _ = mod(xe[:, k:], s1)
        r["y_eval"] = _f(torch.stack(oe, dim=1))
        r["y_eval_split"] = _f(torch.stack(o1 + o2, dim=1))
        return r

    with_noise_floor(d, m, run)
    d["x"], d["dy"] = x0.numpy(), w_.numpy()
    d["meta"] = np.array([base_ch, int(skip), layers, B, T, H, W])
    _save(name, d)


def _save(name, d):
    # parameter gradients are stored as fp32 (relative rounding 6e-8, far below every tolerance
    # used); outputs, input gradients and the noise-floor vectors stay float64
    out = {}
    for k, v in d.items():
        v = np.asarray(v)
        if v.dtype == np.float64 and k.startswith("g."):
            v = v.astype(np.float32)
        out[k] = v
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name)


if __name__ == "__main__":
    convlstm_fixture("convlstm_c8_l1_zero.npz", 8, 8, 1, 2, 4, 6, 6, False, 11)
    convlstm_fixture("convlstm_c6_12_l2_state.npz", 6, 12, 2, 2, 3, 5, 7, True, 12)
    convlstm_fixture("convlstm_c16_l1_state.npz", 16, 16, 1, 2, 4, 8, 8, True, 13)
    block_fixture("double_3_8.npz", "double", 3, 8, 3, 8, 8, 21)
    block_fixture("down_8_16.npz", "down", 8, 16, 2, 9, 10, 22)
    block_fixture("up_16_8.npz", "up", 16, 8, 2, 8, 8, 23)
    block_fixture("up_16_8_pad.npz", "up", 16, 8, 2, 9, 11, 24)
    model_fixture("model_b4_skip.npz", 4, True, 1, 2, 3, 16, 16, 31)
    model_fixture("model_b2_noskip_l2.npz", 2, False, 2, 2, 4, 32, 16, 32)
    # well-conditioned BatchNorm statistics (4 x 4 x 4 = 64 samples per channel at the bottleneck)
    model_fixture("model_b4_skip_64.npz", 4, True, 1, 4, 3, 64, 64, 33)
