"""Golden fixture for the training loss, produced by RUNNING THE REFERENCE'S OWN
`compute_loss` (main.py:28-72) in float64 on seeded inputs.  Only in the build container:

    python tests/golden/make_golden_loss.py

main.py imports train.resnet18, which needs segmentation_models_pytorch (absent, no network): a stub module
is inserted for the import only -- compute_loss itself is plain torch.
"""
import os
import sys
import types

import numpy as np
import torch

sys.modules.setdefault("segmentation_models_pytorch", types.ModuleType("segmentation_models_pytorch"))
sys.path.insert(0, "/root/reference")
import main as ref_main  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(yp, y, mask, use_mask):
    yp = torch.from_numpy(yp).double().requires_grad_(True)
    loss = ref_main.compute_loss(yp, torch.from_numpy(y).double(),
                                 None if mask is None else torch.from_numpy(mask).double(), use_mask)
    loss.backward()
    return float(loss), yp.grad.numpy()


def main():
    rng = np.random.default_rng(21)
    out = {}
    for name, (B, T, H, W) in {"a": (2, 3, 8, 8), "b": (1, 2, 5, 12), "c": (3, 1, 16, 16)}.items():
        yp = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
        y = np.clip(rng.standard_normal((B, T, 1, H, W)), -1, 1).astype(np.float32)
        mask = (rng.random((B, T, 1, H, W)) < 0.6).astype(np.float32)
        out[f"{name}.yp"], out[f"{name}.y"], out[f"{name}.mask"] = yp, y, mask
        for tag, m, use in (("mask", mask, True), ("nomask", None, True), ("ignored", mask, False)):
            loss, g = run(yp, y, m, use)
            out[f"{name}.{tag}.loss"] = np.float64(loss)
            out[f"{name}.{tag}.grad"] = g
    # an all-zero mask: the epsilon in the denominators decides
    yp = rng.standard_normal((1, 2, 1, 6, 6)).astype(np.float32)
    y = rng.standard_normal((1, 2, 1, 6, 6)).astype(np.float32)
    mask = np.zeros((1, 2, 1, 6, 6), dtype=np.float32)
    out["z.yp"], out["z.y"], out["z.mask"] = yp, y, mask
    loss, g = run(yp, y, mask, True)
    out["z.mask.loss"], out["z.mask.grad"] = np.float64(loss), g
    np.savez_compressed(os.path.join(HERE, "loss_main_compute_loss.npz"), **out)
    print("wrote loss_main_compute_loss.npz", {k: float(v) for k, v in out.items() if k.endswith(".loss")})


if __name__ == "__main__":
    main()
