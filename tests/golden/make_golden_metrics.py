"""Golden fixture for the epoch metrics, produced by RUNNING THE REFERENCE'S OWN `evaluate`
(main.py:150-204) with the reference's NPZSequenceDataset (unet.py:208-327) as `dataset_obj`, on a stand-in
model that returns preset predictions.  Only in the build container:

    python tests/golden/make_golden_metrics.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

sys.modules.setdefault("segmentation_models_pytorch", types.ModuleType("segmentation_models_pytorch"))
sys.path.insert(0, "/root/reference")
import main as ref_main  # noqa: E402
from train.unet import NPZSequenceDataset  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class Preset(torch.nn.Module):
    """Returns the next preset prediction as a list of T frames, like TemporalUNetDualView.forward."""

    def __init__(self, preds):
        super().__init__()
        self.preds, self.i = preds, 0

    def forward(self, x):
        p = self.preds[self.i]
        self.i += 1
        return [p[:, t] for t in range(p.shape[1])], None


def main():
    rng = np.random.default_rng(33)
    out = {}
    N, T, H, W = 6, 3, 12, 10
    X = (rng.random((N, T, 2, H, W)) * 3.0).astype(np.float32)      # mask = X[:, :, 0:1] > 1.1: about 63 % valid
    Y = np.clip(rng.standard_normal((N, T, 1, H, W)) * 2.5, -7.0, 8.0).astype(np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "d.npz")
        np.savez(path, X=X, Y=Y)
        for tr in ("asinh", "signed_log", "none"):
            ds = NPZSequenceDataset(path, y_transform=tr)
            items = [ds[i] for i in range(N)]
            batches = []
            for i0, i1 in ((0, 2), (2, 3), (3, 6)):                 # ragged batch sizes
                x = torch.stack([items[i][0] for i in range(i0, i1)])
                y = torch.stack([items[i][1] for i in range(i0, i1)])
                m = torch.stack([items[i][2] for i in range(i0, i1)])
                batches.append((x, y, m))
            preds = [y + torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32)) * 0.3
                     for _, y, _ in batches]
            out[f"{tr}.params"] = np.array([ds.trans_min, ds.trans_max, ds.y_scale], dtype=np.float64)
            for bi, ((x, y, m), p) in enumerate(zip(batches, preds)):
                out[f"{tr}.b{bi}.x"], out[f"{tr}.b{bi}.y"] = x.numpy(), y.numpy()
                out[f"{tr}.b{bi}.mask"], out[f"{tr}.b{bi}.pred"] = m.numpy(), p.numpy()
            for use in (True, False):
                res = ref_main.evaluate(Preset(preds), batches, torch.device("cpu"), ds, use_mask=use)
                out[f"{tr}.use{int(use)}.result"] = np.array([float(v) for v in res], dtype=np.float64)
        # no valid pixel at all: the zeros branch of main.py:198-199
        ds = NPZSequenceDataset(path, y_transform="asinh")
        x, y, m = batches[1]
        res = ref_main.evaluate(Preset([preds[1]]), [(x, y, torch.zeros_like(m))], torch.device("cpu"), ds, use_mask=True)
        out["empty.result"] = np.array([float(v) for v in res], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "metrics_main_evaluate.npz"), **out)
    print({k: v for k, v in out.items() if k.endswith("result")})


if __name__ == "__main__":
    main()
