"""Scratch: eval-mode structured-gradient errors of the base_ch=16 model per precision mode, per tensor."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import unet_convlstm_b200 as pkg
from oracle import torch_port as TP
from train.unet import TemporalUNetDualView
from unet_convlstm_b200 import ops

def rel2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))

B, T, S = 2, 2, 64
torch.manual_seed(21)
m0 = TemporalUNetDualView(base_ch=16, use_skip_lstm=True)
g = torch.Generator().manual_seed(5)
for mod in m0.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
        mod.running_mean.copy_(0.05 * torch.randn(mod.num_features, generator=g))
        mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
sd = {k: v.clone() for k, v in m0.state_dict().items()}
x, yt, mk = bench.make_batch(B, T, S, 7)
dy = -2000.0 * yt * mk / mk.sum()
p = TP.params_from_state_dict(sd, torch.float64)
out_r, _ = TP.temporal_unet(p, x.double(), None, training=False, track=False)
y_r = torch.stack(out_r, dim=1)
(y_r * dy.double()).sum().backward()
for mode, env in (("fp32", {}), ("tf32", {}), ("tf32", {"WGRAD_TF32": False}), ("bf16", {})):
    pkg.set_precision(mode)
    for k, v in env.items():
        setattr(ops, k, v)
    m = TemporalUNetDualView(base_ch=16, use_skip_lstm=True)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    out, _ = m(x.cuda())
    y = torch.stack(out, dim=1)
    (y * dy.cuda()).sum().backward()
    torch.cuda.synchronize()
    errs = {k: rel2(prm.grad.double().cpu().numpy(), p[k].grad.numpy()) for k, prm in m.named_parameters() if np.abs(p[k].grad.numpy()).max() > 1e-9}
    print(f"== {mode} {env}: y {rel2(y.detach().double().cpu().numpy(), y_r.detach().numpy()):.2e} median {np.median(list(errs.values())):.2e}")
    if mode == "tf32" and not env:
        for k, v in errs.items():
            print(f"   {v:.2e} {k}")
    ops.WGRAD_TF32 = True

# the reference's own GPU arithmetic on the same problem: its train/unet.py through stock PyTorch (cuDNN), TF32 on / off
Model, _, src = bench.load_reference()
for tf in (True, False):
    torch.backends.cudnn.allow_tf32 = tf
    torch.backends.cuda.matmul.allow_tf32 = tf
    rm = Model(base_ch=16, use_skip_lstm=True)
    rm.load_state_dict(sd)
    rm = rm.cuda().eval()
    out, _ = rm(x.cuda())
    y = torch.stack(out, dim=1)
    (y * dy.cuda()).sum().backward()
    torch.cuda.synchronize()
    errs = {k: rel2(prm.grad.double().cpu().numpy(), p[k].grad.numpy()) for k, prm in rm.named_parameters() if np.abs(p[k].grad.numpy()).max() > 1e-9}
    print(f"== reference on GPU (cuDNN, allow_tf32={tf}) from {src}: y {rel2(y.detach().double().cpu().numpy(), y_r.detach().numpy()):.2e} median {np.median(list(errs.values())):.2e} "
          f"bottleneck.net.1.net.0.weight {errs['bottleneck.net.1.net.0.weight']:.2e} temporal {errs['temporal.layers.0.conv.weight']:.2e} inc {errs['inc.net.0.weight']:.2e}")
