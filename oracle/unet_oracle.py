"""CPU oracle for the ConvLSTM / UNet-block hot path  --  TEST INFRASTRUCTURE ONLY.

A plain numpy (float64 by default) restatement of the arithmetic of the reference
`train/unet.py` (dordanino12/unet-convlstm), forward AND hand-derived backward, used as the
checker for the CUDA kernels.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py` may import this module; the product path
(`train/unet.py`, `unet_convlstm_b200/`) never does and fails loudly without its CUDA library.

Parity is PINNED: `tests/golden/make_golden.py` imports the reference classes from
/root/reference (CPU, fp32 and fp64), runs them on seeded inputs and stores inputs, state_dicts,
outputs and autograd gradients as fixtures; `tests/test_oracle.py` checks every function here
against those fixtures.  The arithmetic itself lives in PyTorch (pinned `torch==2.4.1` in the
reference's requirements.txt:59; operator semantics of Conv2d / BatchNorm2d / ConvTranspose2d /
MaxPool2d / sigmoid / tanh are version stable), call sites cited per function below.

Layouts follow the reference: activations NCHW, conv weights OIHW, ConvTranspose weights IOHW.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------
# elementary ops (each: forward returning (out, cache); backward(cache, dout) -> grads)
# --------------------------------------------------------------------------------------------


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def conv2d_fwd(x, w, b=None):
    """nn.Conv2d(cin, cout, k, padding=k//2), stride 1 (unet.py:19, :70-71, :104, :117)."""
    k = w.shape[2]
    p = k // 2
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
    win = np.lib.stride_tricks.sliding_window_view(xp, (k, k), axis=(2, 3))  # B,C,H,W,k,k
    out = np.einsum("bchwij,ocij->bohw", win, w, optimize=True)
    if b is not None:
        out = out + b[None, :, None, None]
    return out, (x, w, b is not None)


def conv2d_bwd(cache, dout):
    x, w, has_b = cache
    k = w.shape[2]
    p = k // 2
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
    win = np.lib.stride_tricks.sliding_window_view(xp, (k, k), axis=(2, 3))
    dw = np.einsum("bchwij,bohw->ocij", win, dout, optimize=True)
    db = dout.sum(axis=(0, 2, 3)) if has_b else None
    # dgrad: correlate dout with the spatially flipped, channel-transposed filter
    wf = np.flip(w, axis=(2, 3)).transpose(1, 0, 2, 3)
    dp = np.pad(dout, ((0, 0), (0, 0), (p, p), (p, p)))
    dwin = np.lib.stride_tricks.sliding_window_view(dp, (k, k), axis=(2, 3))
    dx = np.einsum("bohwij,coij->bchw", dwin, wf, optimize=True)
    return dx, dw, db


def batchnorm_fwd(x, gamma, beta, running_mean, running_var, training, eps=1e-5, momentum=0.1):
    """nn.BatchNorm2d defaults (unet.py:70-71): batch statistics over (B,H,W) of THIS call --
    the reference calls it once per timestep (unet.py:179-182, :196-202) -- biased variance for
    normalisation, unbiased for the running estimate.  Returns new running stats as well."""
    if training:
        n = x.shape[0] * x.shape[2] * x.shape[3]
        mean = x.mean(axis=(0, 2, 3))
        var = x.var(axis=(0, 2, 3))
        new_rm = (1 - momentum) * running_mean + momentum * mean
        new_rv = (1 - momentum) * running_var + momentum * var * (n / max(n - 1, 1))
    else:
        mean, var = running_mean, running_var
        new_rm, new_rv = running_mean, running_var
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean[None, :, None, None]) * rstd[None, :, None, None]
    y = gamma[None, :, None, None] * xhat + beta[None, :, None, None]
    return y, (xhat, gamma, rstd, training), new_rm, new_rv


def batchnorm_bwd(cache, dy):
    xhat, gamma, rstd, training = cache
    dgamma = (dy * xhat).sum(axis=(0, 2, 3))
    dbeta = dy.sum(axis=(0, 2, 3))
    g = (gamma * rstd)[None, :, None, None]
    if training:
        n = dy.shape[0] * dy.shape[2] * dy.shape[3]
        dx = g * (dy - dbeta[None, :, None, None] / n - xhat * dgamma[None, :, None, None] / n)
    else:
        dx = g * dy
    return dx, dgamma, dbeta


def relu_fwd(x):
    y = np.maximum(x, 0)
    return y, (y > 0)


def relu_bwd(mask, dy):
    return dy * mask


def maxpool2_fwd(x):
    """nn.MaxPool2d(2) (unet.py:81): floor division of H,W; the FIRST maximum in window scan
    order (0,0),(0,1),(1,0),(1,1) receives the gradient (ATen max_pool2d_with_indices)."""
    B, C, H, W = x.shape
    Ho, Wo = H // 2, W // 2
    xc = x[:, :, : Ho * 2, : Wo * 2].reshape(B, C, Ho, 2, Wo, 2).transpose(0, 1, 2, 4, 3, 5).reshape(B, C, Ho, Wo, 4)
    idx = xc.argmax(axis=-1)  # first occurrence
    y = np.take_along_axis(xc, idx[..., None], axis=-1)[..., 0]
    return y, (idx, x.shape)


def maxpool2_bwd(cache, dy):
    idx, shape = cache
    B, C, H, W = shape
    Ho, Wo = H // 2, W // 2
    d = np.zeros((B, C, Ho, Wo, 4), dtype=dy.dtype)
    np.put_along_axis(d, idx[..., None], dy[..., None], axis=-1)
    dx = np.zeros(shape, dtype=dy.dtype)
    dx[:, :, : Ho * 2, : Wo * 2] = d.reshape(B, C, Ho, Wo, 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(B, C, Ho * 2, Wo * 2)
    return dx


def convtranspose2x2_fwd(x, w, b):
    """nn.ConvTranspose2d(cin, cin//2, 2, stride=2) (unet.py:90); w is [cin, cout, 2, 2]."""
    B, C, H, W = x.shape
    out = np.einsum("bchw,coij->bohiwj", x, w, optimize=True).reshape(B, w.shape[1], 2 * H, 2 * W)
    return out + b[None, :, None, None], (x, w)


def convtranspose2x2_bwd(cache, dout):
    x, w = cache
    B, C, H, W = x.shape
    d = dout.reshape(B, w.shape[1], H, 2, W, 2)
    dx = np.einsum("bohiwj,coij->bchw", d, w, optimize=True)
    dw = np.einsum("bchw,bohiwj->coij", x, d, optimize=True)
    db = dout.sum(axis=(0, 2, 3))
    return dx, dw, db


# --------------------------------------------------------------------------------------------
# ConvLSTM  (unet.py:14-60)
# --------------------------------------------------------------------------------------------


def convlstm_cell_fwd(x, h, c, w, b):
    """ConvLSTMCell.forward (unet.py:21-36): gates = conv(cat[x,h]); i,f,g,o = chunk(4);
    sigma,sigma,tanh,sigma; c' = f*c + i*g; h' = o*tanh(c')."""
    Ch = w.shape[0] // 4
    xin = np.concatenate([x, h], axis=1)
    z, ccache = conv2d_fwd(xin, w, b)
    i, f, g, o = sigmoid(z[:, :Ch]), sigmoid(z[:, Ch:2 * Ch]), np.tanh(z[:, 2 * Ch:3 * Ch]), sigmoid(z[:, 3 * Ch:])
    c_next = f * c + i * g
    tc = np.tanh(c_next)
    h_next = o * tc
    return h_next, c_next, (ccache, i, f, g, o, c, tc, x.shape[1])


def convlstm_cell_bwd(cache, dh, dc):
    """Hand-derived autograd of unet.py:29-35 followed by conv backward and the cat split."""
    ccache, i, f, g, o, c_prev, tc, cin = cache
    do = dh * tc
    dc = dc + dh * o * (1 - tc * tc)
    di, dg, df, dc_prev = dc * g, dc * i, dc * c_prev, dc * f
    dz = np.concatenate([di * i * (1 - i), df * f * (1 - f), dg * (1 - g * g), do * o * (1 - o)], axis=1)
    dxin, dw, db = conv2d_bwd(ccache, dz)
    return dxin[:, :cin], dxin[:, cin:], dc_prev, dw, db


def convlstm_fwd(x_seq, layers, state=None):
    """ConvLSTM.forward (unet.py:46-60), layer-major.  layers: list of (w, b); x_seq: list of
    [B,C,H,W]; state: list of (h,c) or None per layer.  Returns (out_seq, new_states, caches)."""
    T = len(x_seq)
    out = x_seq
    new_states, caches = [], []
    for li, (w, b) in enumerate(layers):
        Ch = w.shape[0] // 4
        B, _, H, W = out[0].shape
        if state is None or state[li] is None:
            h = np.zeros((B, Ch, H, W), dtype=out[0].dtype)
            c = np.zeros((B, Ch, H, W), dtype=out[0].dtype)
        else:
            h, c = state[li]
        seq, lc = [], []
        for t in range(T):
            h, c, cache = convlstm_cell_fwd(out[t], h, c, w, b)
            seq.append(h)
            lc.append(cache)
        out = seq
        new_states.append((h, c))
        caches.append(lc)
    return out, new_states, caches


def convlstm_bwd(caches, dout_seq, dstate=None):
    """BPTT.  dout_seq: list of dL/dh_t of the LAST layer; dstate: optional list of (dh_T, dc_T).
    Returns (dx_seq, [(dw, db)] per layer, [(dh0, dc0)] per layer)."""
    L = len(caches)
    T = len(dout_seq)
    dseq = list(dout_seq)
    wgrads = [None] * L
    d0 = [None] * L
    for li in reversed(range(L)):
        lc = caches[li]
        dh_next = 0.0 if dstate is None or dstate[li] is None else dstate[li][0]
        dc_next = 0.0 if dstate is None or dstate[li] is None else dstate[li][1]
        dw_acc, db_acc = 0.0, 0.0
        dx_seq = [None] * T
        for t in reversed(range(T)):
            dx, dh_prev, dc_prev, dw, db = convlstm_cell_bwd(lc[t], dseq[t] + dh_next, dc_next)
            dx_seq[t] = dx
            dh_next, dc_next = dh_prev, dc_prev
            dw_acc = dw_acc + dw
            db_acc = db_acc + db
        wgrads[li] = (dw_acc, db_acc)
        d0[li] = (dh_next, dc_next)
        dseq = dx_seq
    return dseq, wgrads, d0


# --------------------------------------------------------------------------------------------
# UNet blocks (unet.py:66-107) on parameter dicts keyed like the reference state_dict
# --------------------------------------------------------------------------------------------


class Tape:
    """Collects parameter gradients and updated BatchNorm buffers, keyed by state_dict name."""

    def __init__(self, params, training):
        self.p = params
        self.training = training
        self.grads = {}
        self.new_buffers = {}

    def buf(self, name):
        return self.new_buffers.get(name, self.p[name])

    def add(self, name, g):
        self.grads[name] = self.grads.get(name, 0.0) + g


def double_conv_fwd(tp: Tape, pre, x):
    """DoubleConv (unet.py:66-75): Sequential children 0,1,3,4 carry parameters."""
    caches = []
    for ci, bi in (("0", "1"), ("3", "4")):
        z, cc = conv2d_fwd(x, tp.p[f"{pre}.{ci}.weight"], tp.p[f"{pre}.{ci}.bias"])
        y, bc, rm, rv = batchnorm_fwd(z, tp.p[f"{pre}.{bi}.weight"], tp.p[f"{pre}.{bi}.bias"],
                                      tp.buf(f"{pre}.{bi}.running_mean"), tp.buf(f"{pre}.{bi}.running_var"),
                                      tp.training)
        if tp.training:
            tp.new_buffers[f"{pre}.{bi}.running_mean"] = rm
            tp.new_buffers[f"{pre}.{bi}.running_var"] = rv
            nb = f"{pre}.{bi}.num_batches_tracked"
            tp.new_buffers[nb] = tp.buf(nb) + 1
        x, rc = relu_fwd(y)
        caches.append((cc, bc, rc))
    return x, caches


def double_conv_bwd(tp: Tape, pre, caches, dy):
    for (ci, bi), (cc, bc, rc) in zip((("3", "4"), ("0", "1")), reversed(caches)):
        dy = relu_bwd(rc, dy)
        dy, dg, db = batchnorm_bwd(bc, dy)
        tp.add(f"{pre}.{bi}.weight", dg)
        tp.add(f"{pre}.{bi}.bias", db)
        dy, dw, dbias = conv2d_bwd(cc, dy)
        tp.add(f"{pre}.{ci}.weight", dw)
        tp.add(f"{pre}.{ci}.bias", dbias)
    return dy


def down_fwd(tp, pre, x):
    """Down (unet.py:78-84): MaxPool2d(2) then DoubleConv; keys `<pre>.net.1.net.*`."""
    y, pc = maxpool2_fwd(x)
    y, dc = double_conv_fwd(tp, f"{pre}.net.1.net", y)
    return y, (pc, dc)


def down_bwd(tp, pre, cache, dy):
    pc, dc = cache
    return maxpool2_bwd(pc, double_conv_bwd(tp, f"{pre}.net.1.net", dc, dy))


def up_fwd(tp, pre, x1, x2):
    """Up (unet.py:87-98): ConvTranspose 2x2 s2, F.pad to the skip size, cat([skip, up]), DoubleConv."""
    u, uc = convtranspose2x2_fwd(x1, tp.p[f"{pre}.up.weight"], tp.p[f"{pre}.up.bias"])
    dY, dX = x2.shape[2] - u.shape[2], x2.shape[3] - u.shape[3]
    pads = (dY // 2, dY - dY // 2, dX // 2, dX - dX // 2)
    u = np.pad(u, ((0, 0), (0, 0), (pads[0], pads[1]), (pads[2], pads[3])))
    y, dc = double_conv_fwd(tp, f"{pre}.conv.net", np.concatenate([x2, u], axis=1))
    return y, (uc, pads, x2.shape[1], dc)


def up_bwd(tp, pre, cache, dy):
    uc, pads, c2, dc = cache
    d = double_conv_bwd(tp, f"{pre}.conv.net", dc, dy)
    dx2, du = d[:, :c2], d[:, c2:]
    H, W = du.shape[2], du.shape[3]
    du = du[:, :, pads[0]:H - pads[1], pads[2]:W - pads[3]]
    dx1, dw, db = convtranspose2x2_bwd(uc, du)
    tp.add(f"{pre}.up.weight", dw)
    tp.add(f"{pre}.up.bias", db)
    return dx1, dx2


# --------------------------------------------------------------------------------------------
# TemporalUNetDualView (unet.py:131-204)
# --------------------------------------------------------------------------------------------


def _lstm_layers(p, name):
    layers, l = [], 0
    while f"{name}.layers.{l}.conv.weight" in p:
        layers.append((p[f"{name}.layers.{l}.conv.weight"], p[f"{name}.layers.{l}.conv.bias"]))
        l += 1
    return layers


def temporal_unet_fwd(params, x_seq, state=None, training=True):
    """TemporalUNetDualView.forward (unet.py:174-204), use_attention=False.  params: dict
    name -> ndarray with the reference state_dict keys.  x_seq: [B,T,C,H,W].
    Returns (out [B,T,out,H,W], new_state, tape, caches)."""
    tp = Tape(params, training)
    B, T = x_seq.shape[:2]
    use_skip = "lstm_skip3.layers.0.conv.weight" in params
    enc = []
    for t in range(T):
        x0, c0 = double_conv_fwd(tp, "inc.net", x_seq[:, t])
        x1, c1 = down_fwd(tp, "down1", x0)
        x2, c2 = down_fwd(tp, "down2", x1)
        x3, c3 = down_fwd(tp, "down3", x2)
        xb, cb = down_fwd(tp, "bottleneck", x3)
        enc.append(((x0, x1, x2, x3, xb), (c0, c1, c2, c3, cb)))
    bott, new_state, tcache = convlstm_fwd([e[0][4] for e in enc], _lstm_layers(params, "temporal"), state)
    if use_skip:
        s3, _, s3cache = convlstm_fwd([e[0][3] for e in enc], _lstm_layers(params, "lstm_skip3"))
        s2, _, s2cache = convlstm_fwd([e[0][2] for e in enc], _lstm_layers(params, "lstm_skip2"))
    else:
        s3, s2 = [e[0][3] for e in enc], [e[0][2] for e in enc]
        s3cache = s2cache = None
    outs, dec = [], []
    for t in range(T):
        d3, u3 = up_fwd(tp, "up3", bott[t], s3[t])
        d2, u2 = up_fwd(tp, "up2", d3, s2[t])
        d1, u1 = up_fwd(tp, "up1", d2, enc[t][0][1])
        d0, u0 = up_fwd(tp, "up0", d1, enc[t][0][0])
        y, oc = conv2d_fwd(d0, params["outc.conv.weight"], params["outc.conv.bias"])
        outs.append(y)
        dec.append((u3, u2, u1, u0, oc))
    caches = (enc, tcache, s3cache, s2cache, dec)
    return np.stack(outs, axis=1), new_state, tp, caches


def temporal_unet_bwd(tp, caches, dout, dstate=None):
    """Backward of temporal_unet_fwd for dL/dout [B,T,out,H,W]; returns dL/dx_seq and fills
    tp.grads with parameter gradients keyed like the reference state_dict."""
    enc, tcache, s3cache, s2cache, dec = caches
    T = len(enc)
    dbott, ds3, ds2, dx1s, dx0s = [None] * T, [None] * T, [None] * T, [None] * T, [None] * T
    for t in range(T):
        u3, u2, u1, u0, oc = dec[t]
        d, dw, db = conv2d_bwd(oc, dout[:, t])
        tp.add("outc.conv.weight", dw)
        tp.add("outc.conv.bias", db)
        d, dx0s[t] = up_bwd(tp, "up0", u0, d)
        d, dx1s[t] = up_bwd(tp, "up1", u1, d)
        d, ds2[t] = up_bwd(tp, "up2", u2, d)
        dbott[t], ds3[t] = up_bwd(tp, "up3", u3, d)

    def lstm_back(name, cache, dseq, dst=None):
        dx, wg, _ = convlstm_bwd(cache, dseq, dst)
        for l, (dw, db) in enumerate(wg):
            tp.add(f"{name}.layers.{l}.conv.weight", dw)
            tp.add(f"{name}.layers.{l}.conv.bias", db)
        return dx

    dxb = lstm_back("temporal", tcache, dbott, dstate)
    if s3cache is not None:
        ds3 = lstm_back("lstm_skip3", s3cache, ds3)
        ds2 = lstm_back("lstm_skip2", s2cache, ds2)
    dxs = []
    for t in range(T):
        c0, c1, c2, c3, cb = enc[t][1]
        d = down_bwd(tp, "bottleneck", cb, dxb[t]) + ds3[t]
        d = down_bwd(tp, "down3", c3, d) + ds2[t]
        d = down_bwd(tp, "down2", c2, d) + dx1s[t]
        d = down_bwd(tp, "down1", c1, d) + dx0s[t]
        dxs.append(double_conv_bwd(tp, "inc.net", c0, d))
    return np.stack(dxs, axis=1)
