"""Generates the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Usage (only in the build container, where /root/reference exists):
    python tests/golden/make_golden.py

Imports the unmodified reference classes from /root/reference/train/unet.py (CPU), runs them in
float64 on seeded inputs and stores inputs, state_dict, outputs and autograd gradients as .npz.
The reference ships no tests / golden vectors of its own, so these files
are what pins the oracle (oracle/unet_oracle.py) and, through it, the CUDA kernels.
"""
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


def grads_np(module, prefix="g."):
    # gradients are stored as fp32 (relative rounding 6e-8, far below every tolerance used)
    return {prefix + k: p.grad.detach().numpy().astype(np.float32) for k, p in module.named_parameters()}


def randomize_bn(module, gen):
    """Non-trivial BN affine/running stats so that fixtures exercise them."""
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=gen)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=gen)
            m.running_mean.data = 0.1 * torch.randn(m.running_mean.shape, generator=gen)
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=gen)


def convlstm_fixture(name, cin, ch, layers, B, T, H, W, with_state, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = ConvLSTM(cin, ch, num_layers=layers)
    round_params_to_fp32(m)
    xs = [torch.randn(B, cin, H, W, generator=gen, requires_grad=True) for _ in range(T)]
    state = None
    if with_state:
        state = [(torch.randn(B, ch, H, W, generator=gen, requires_grad=True) * 0.5,
                  torch.randn(B, ch, H, W, generator=gen, requires_grad=True) * 0.5) for _ in range(layers)]
        for h, c in state:
            h.retain_grad(), c.retain_grad()
    sd = sd_np(m)
    out, new_state = m(xs, state)
    wts = [torch.randn(B, ch, H, W, generator=gen) for _ in range(T)]
    wh = torch.randn(B, ch, H, W, generator=gen)
    wc = torch.randn(B, ch, H, W, generator=gen)
    loss = sum((o * w_).sum() for o, w_ in zip(out, wts)) + (new_state[-1][0] * wh).sum() + (new_state[-1][1] * wc).sum()
    loss.backward()
    d = dict(sd)
    d.update(grads_np(m))
    d["x"] = np.stack([x.detach().numpy() for x in xs])
    d["dx"] = np.stack([x.grad.numpy() for x in xs])
    d["out"] = np.stack([o.detach().numpy() for o in out])
    d["dout"] = np.stack([w_.numpy() for w_ in wts])
    d["dh_last"], d["dc_last"] = wh.numpy(), wc.numpy()
    for l in range(layers):
        d[f"hT{l}"], d[f"cT{l}"] = new_state[l][0].detach().numpy(), new_state[l][1].detach().numpy()
        if with_state:
            d[f"h0{l}"], d[f"c0{l}"] = state[l][0].detach().numpy(), state[l][1].detach().numpy()
            d[f"dh0{l}"], d[f"dc0{l}"] = state[l][0].grad.numpy(), state[l][1].grad.numpy()
    d["meta"] = np.array([cin, ch, layers, B, T, H, W, int(with_state)])
    np.savez_compressed(os.path.join(HERE, name), **d)
    print("wrote", name)


def block_fixture(name, kind, cin, cout, B, H, W, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = {"double": DoubleConv, "down": Down, "up": Up}[kind](cin, cout)
    randomize_bn(m, gen)
    round_params_to_fp32(m)
    d = sd_np(m)
    if kind == "up":
        x1 = torch.randn(B, cin, H // 2, W // 2, generator=gen, requires_grad=True)
        x2 = torch.randn(B, cin // 2, H, W, generator=gen, requires_grad=True)
        args = (x1, x2)
    else:
        args = (torch.randn(B, cin, H, W, generator=gen, requires_grad=True),)
    m.train()
    y = m(*args)
    w_ = torch.randn(y.shape, generator=gen)
    (y * w_).sum().backward()
    d.update(grads_np(m))
    d.update(sd_np(m, "after.", only_buffers=True))
    d["y_train"], d["dy"] = y.detach().numpy(), w_.numpy()
    for i, a in enumerate(args):
        d[f"x{i}"], d[f"dx{i}"] = a.detach().numpy(), a.grad.numpy()
    m.eval()
    d["y_eval"] = m(*args).detach().numpy()
    np.savez_compressed(os.path.join(HERE, name), **d)
    print("wrote", name)


def model_fixture(name, base_ch, skip, layers, B, T, H, W, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = TemporalUNetDualView(base_ch=base_ch, lstm_layers=layers, use_skip_lstm=skip)
    randomize_bn(m, gen)
    round_params_to_fp32(m)
    d = sd_np(m)
    x = torch.rand(B, T, 2, H, W, generator=gen, requires_grad=True)
    m.train()
    out, st = m(x)
    y = torch.stack(out, dim=1)
    w_ = torch.randn(y.shape, generator=gen)
    (y * w_).sum().backward()
    d.update(grads_np(m))
    d.update(sd_np(m, "after.", only_buffers=True))
    d["x"], d["dx"], d["y_train"], d["dy"] = x.detach().numpy(), x.grad.numpy(), y.detach().numpy(), w_.numpy()
    for l in range(layers):
        d[f"hT{l}"], d[f"cT{l}"] = st[l][0].detach().numpy(), st[l][1].detach().numpy()
    # eval mode (running stats as updated by the train-mode forward above), plus the state
    # round trip of SURVEY 8c: model(x[:, :k]) then model(x[:, k:], state)
    m.eval()
    with torch.no_grad():
        oe, _ = m(x)
        k = T // 2
        o1, s1 = m(x[:, :k])
        o2, s2 = m(x[:, k:], s1)
    d["y_eval"] = torch.stack(oe, dim=1).numpy()
    d["y_eval_split"] = torch.stack(o1 + o2, dim=1).numpy()
    d["meta"] = np.array([base_ch, int(skip), layers, B, T, H, W])
    np.savez_compressed(os.path.join(HERE, name), **d)
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
