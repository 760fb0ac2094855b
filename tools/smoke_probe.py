"""Scratch: bf16 whole-model error against the fp64 port for several conditioning choices (to pick smoke()'s problem)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_convlstm_b200 as pkg
from oracle import torch_port as TP
from train.unet import TemporalUNetDualView

def rel2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))

pkg.set_precision("bf16")
for name, B, T, S, training, var, bc in [("train B8 T3 64x64 bc16", 8, 3, 64, True, None, 16), ("train B4 T2 64x64 bc16", 4, 2, 64, True, None, 16),
                                         ("eval fresh B4 T3 bc16", 4, 3, 64, False, None, 16), ("eval U(.5,1.5) bc16", 4, 3, 64, False, (0.5, 1.0), 16),
                                         ("eval U(.05,.15) B2 T2 bc64", 2, 2, 64, False, (0.05, 0.1), 64), ("eval U(.5,1.5) B2 T2 bc64", 2, 2, 64, False, (0.5, 1.0), 64),
                                         ("train B2 T2 bc64", 2, 2, 64, True, None, 64)]:

    import bench
    xb, yb, mb = bench.make_batch(B, T, S, 7)
    x = xb.numpy()
    # cotangent of the masked MSE the reference's overfit check trains on, at y = 0: a structured, non-cancelling
    # gradient signal instead of white noise
    dy = (-2.0 * yb * mb / mb.sum()).numpy() * 1000.0
    torch.manual_seed(1)
    m = TemporalUNetDualView(base_ch=bc, use_skip_lstm=True)
    g = torch.Generator().manual_seed(2)
    if var is not None:
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(0.05 * torch.randn(mod.num_features, generator=g))
                mod.running_var.copy_(var[0] + var[1] * torch.rand(mod.num_features, generator=g))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    res = {}
    for tag, dt, ac in (("f64", torch.float64, False), ("ref-bf16-autocast", torch.float32, True)):
        p = TP.params_from_state_dict(sd, dt)
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=ac):
            out, _ = TP.temporal_unet(p, torch.from_numpy(x).to(dt), None, training=training, track=False)
        y = torch.stack(out, dim=1).to(dt)
        (y * torch.from_numpy(dy).to(dt)).sum().backward()
        res[tag] = (y.detach().double().numpy(), {k: v.grad.double().numpy() for k, v in p.items() if v.requires_grad})
    m = m.cuda(); m.train(training)
    out, _ = m(torch.from_numpy(x).cuda())
    y = torch.stack(out, dim=1)
    (y * torch.from_numpy(dy).cuda()).sum().backward()
    torch.cuda.synchronize()
    yr, gr = res["f64"]
    ya, ga = res["ref-bf16-autocast"]
    gg = {k: v.grad.double().cpu().numpy() for k, v in m.named_parameters()}
    keys = [k for k in gr if np.abs(gr[k]).max() > 1e-9 * max(1.0, np.abs(dy).max())]
    e_b200 = {k: rel2(gg[k], gr[k]) for k in keys}
    e_ref = {k: rel2(ga[k], gr[k]) for k in keys}
    worst = max(e_b200, key=e_b200.get)
    print(f"{name:22s} y: b200 {rel2(y.detach().double().cpu().numpy(), yr):.2e} ref-autocast {rel2(ya, yr):.2e} | grads b200 median "
          f"{np.median(list(e_b200.values())):.2e} max {e_b200[worst]:.2e} ({worst}) | ref-autocast median {np.median(list(e_ref.values())):.2e} "
          f"max {max(e_ref.values()):.2e} | |y| {np.abs(yr).mean():.2e}", flush=True)
