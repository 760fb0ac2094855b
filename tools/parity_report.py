"""Prints per-tensor parity errors of the CUDA path against the golden fixtures (run on the GPU box).

    python tools/parity_report.py [fp32|bf16] ...

For every fixture: max-relative and L2-relative error of outputs, input gradients, parameter gradients
and BatchNorm buffers.  A development aid; the pass/fail criteria live in tests/test_gpu_parity.py.
"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unet_convlstm_b200 as pkg  # noqa: E402
from train.unet import ConvLSTM, DoubleConv, Down, TemporalUNetDualView, Up  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def errs(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return (np.abs(a - b).max() / max(np.abs(b).max(), 1e-30),
            np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30), np.abs(b).max())


FLOORS = None  # (npz, prefix) of the fixture being reported


def show(name, a, b):
    m, l2, mag = errs(a, b)
    fl = ""
    if FLOORS is not None and (FLOORS[1] + name) in FLOORS[0].files:
        f = FLOORS[0][FLOORS[1] + name]
        fl = f"  floor max {f[0]:8.2e} l2 {f[1]:8.2e}  ratio {l2 / max(f[1], 1e-30):6.1f}"
    print(f"    {name:44s} max-rel {m:9.2e}  l2-rel {l2:9.2e}  |ref|max {mag:9.2e}{fl}")


def _np(t):
    return t.detach().float().cpu().numpy()


def _cuda(a, grad=False):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).cuda().requires_grad_(grad)


def load(module, z):
    module.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}, strict=True)
    return module.cuda()


def grads(module, z):
    p = dict(module.named_parameters())
    for k in z.files:
        if k.startswith("g."):
            show(k, _np(p[k[2:]].grad), z[k])
    b = dict(module.named_buffers())
    for k in z.files:
        if k.startswith("after.") and "num_batches" not in k:
            show(k, _np(b[k[6:]].float()), z[k])


def run(mode):
    pkg.set_precision(mode)
    print(f"===== mode {mode}")
    for path in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        name = os.path.basename(path)
        z = np.load(path)
        global FLOORS
        FLOORS = (z, "e32." if mode == "fp32" else "e16.")
        print(f"  -- {name}")
        if name.startswith("convlstm"):
            cin, ch, L, B, T, H, W, ws = [int(v) for v in z["meta"]]
            m = load(ConvLSTM(cin, ch, num_layers=L), z)
            xs = [_cuda(z["x"][t], True) for t in range(T)]
            st = [(_cuda(z[f"h0{l}"], True), _cuda(z[f"c0{l}"], True)) for l in range(L)] if ws else None
            out, ns = m(xs, st)
            show("out", _np(torch.stack(out)), z["out"])
            loss = sum((o * _cuda(z["dout"][t])).sum() for t, o in enumerate(out))
            loss = loss + (ns[-1][0] * _cuda(z["dh_last"])).sum() + (ns[-1][1] * _cuda(z["dc_last"])).sum()
            loss.backward()
            show("dx", np.stack([_np(x.grad) for x in xs]), z["dx"])
            grads(m, z)
        elif name.startswith("model"):
            base_ch, skip, L, B, T, H, W = [int(v) for v in z["meta"]]
            m = load(TemporalUNetDualView(base_ch=base_ch, lstm_layers=L, use_skip_lstm=bool(skip)), z)
            x = _cuda(z["x"], True)
            m.train()
            out, st = m(x)
            y = torch.stack(out, dim=1)
            show("y_train", _np(y), z["y_train"])
            (y * _cuda(z["dy"])).sum().backward()
            show("dx", _np(x.grad), z["dx"])
            grads(m, z)
            m.eval()
            with torch.no_grad():
                oe, _ = m(x.detach())
            show("y_eval", _np(torch.stack(oe, dim=1)), z["y_eval"])
        else:
            kind = name.split("_")[0]
            cin, cout = int(name.split("_")[1]), int(name.split("_")[2].split(".")[0])
            m = load({"double": DoubleConv, "down": Down, "up": Up}[kind](cin, cout), z)
            args = [_cuda(z[f"x{i}"], True) for i in range(2 if kind == "up" else 1)]
            m.train()
            y = m(*args)
            show("y_train", _np(y), z["y_train"])
            (y * _cuda(z["dy"])).sum().backward()
            for i, a in enumerate(args):
                show(f"dx{i}", _np(a.grad), z[f"dx{i}"])
            grads(m, z)
            m.eval()
            with torch.no_grad():
                show("y_eval", _np(m(*[a.detach() for a in args])), z["y_eval"])


if __name__ == "__main__":
    for mode in (sys.argv[1:] or ["fp32", "bf16"]):
        run(mode)
