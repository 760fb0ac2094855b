"""Host-side dry run of the whole training step on CPU tensors (no GPU, no kernels): every C-ABI call the Python host
code would make -- model forward, loss, backward, clip + AdamW, over two steps, in each precision mode and with the
fusion switches on and off -- is checked against the prototype in include/b200_convlstm.h (arity, pointer vs scalar,
int vs float) and then skipped.  What this catches without a GPU: a wrapper out of step with the header, a call
sequence that changes when a switch is flipped, the packed-weight bookkeeping of optim.AdamW (no b200_pack_weight after
the first step), autograd Functions returning the wrong number of gradients.  It does NOT compute anything: the
numerics are the `-m gpu` suite's job."""
import ctypes

import pytest
import torch


@pytest.fixture
def dry(monkeypatch):
    """Replaces the device-facing helpers by checkers; returns the dict of calls seen."""
    import train.unet as U
    from unet_convlstm_b200 import _lib, loss, ops, optim
    protos = _lib.parse_header()
    seen = {}

    def fake_call(name, *args, tag="", work=None):
        assert name in protos, f"{name} is not declared in include/b200_convlstm.h"
        _, argtypes = protos[name]
        assert len(args) == len(argtypes), (name, len(args), len(argtypes))
        for i, (a, t) in enumerate(zip(args, argtypes)):
            if t is ctypes.c_void_p:
                assert a is None or hasattr(a, "data_ptr") or isinstance(a, (int, ctypes.Array)), (name, i, type(a))
            elif t in (ctypes.c_int, ctypes.c_longlong):
                assert isinstance(a, int), (name, i, type(a), a)
            else:
                assert isinstance(a, (float, int)) and not isinstance(a, bool), (name, i, type(a))
        seen[name] = seen.get(name, 0) + 1

    monkeypatch.setattr(_lib, "call", fake_call)
    monkeypatch.setattr(_lib, "supported", lambda name, *a: False)   # the CUDA-core route: same host plumbing
    monkeypatch.setattr(ops, "_chk", lambda t, name="tensor": t)
    monkeypatch.setattr(ops, "_st", lambda: 0)
    monkeypatch.setattr(optim, "_st", lambda: 0)
    monkeypatch.setattr(optim, "_check", lambda ts, what: None)
    monkeypatch.setattr(ops, "copy_", lambda dst, src, accumulate=False: (dst.add_(src) if accumulate else dst.copy_(src)))
    monkeypatch.setattr(ops, "_pack_ok", lambda w, taps: w.dtype == torch.float32 and taps <= 9)
    monkeypatch.setattr(U, "_require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(ops, "WGRAD_STREAM", False)   # (another test of the session may have opted in; streams need a device)
    if hasattr(loss, "_st"):
        monkeypatch.setattr(loss, "_st", lambda: 0)
    old = ops.get_precision()
    yield seen
    ops.set_precision(old)


def _train_two_steps(base_ch, size=(16, 16), clip=1.0):
    from train.unet import TemporalUNetDualView
    from unet_convlstm_b200 import optim
    torch.manual_seed(0)
    m = TemporalUNetDualView(base_ch=base_ch, use_skip_lstm=True)
    x = torch.randn(2, 3, 2, *size, requires_grad=True)
    opt = optim.AdamW(m.parameters(), lr=1e-3)
    for _ in range(2):
        out, state = m(x)
        assert isinstance(out, list) and len(out) == 3 and out[0].shape == (2, 1, *size)
        assert len(state) == 1 and state[0][0].shape == (2, 16 * base_ch, size[0] // 16, size[1] // 16)
        torch.stack(out, 1).sum().backward()
        missing = [k for k, p in m.named_parameters() if p.grad is None]
        assert not missing, missing
        assert x.grad is not None
        opt.step(clip_max_norm=clip)
        opt.zero_grad(set_to_none=True)
    return m


@pytest.mark.parametrize("precision,base_ch", [("fp32", 8), ("bf16", 16), ("tf32", 8)])
def test_every_call_of_a_training_step_matches_the_header(dry, precision, base_ch):
    from unet_convlstm_b200 import ops
    ops.set_precision(precision)
    _train_two_steps(base_ch)
    # the fused paths are the ones taken ...
    assert dry["b200_bn_relu_apply_pool"] == 2 * 4 and dry["b200_bn_relu_pool_bwd_reduce"] == 2 * 4
    assert dry["b200_bn_relu_pool_bwd_apply"] == 2 * 4
    assert dry["b200_bn_relu_outconv_fwd"] == 2 and dry["b200_bn_relu_outconv_bwd_reduce"] == 2
    assert dry["b200_bn_relu_outconv_bwd_apply"] == 2
    assert "b200_maxpool2_fwd" not in dry and "b200_maxpool2_bwd" not in dry and "b200_outconv_fwd" not in dry
    # ... 18 - 4 - 1 plain normalise passes per step remain
    assert dry["b200_bn_relu_apply"] == 2 * 13 and dry["b200_bn_relu_bwd_apply"] == 2 * 13
    # 21 conv + 4 ConvT weights: packed twice each (forward + data-gradient layout) in the FIRST step only; the optimizer
    # emits the packed copies afterwards (25 weights = two launches of 16 + 9 per step)
    assert dry["b200_pack_weight"] == 50
    assert dry["b200_adamw_pack_multi"] == 2 * 2
    assert dry["b200_grad_sqnorm_multi"] >= 2 and dry["b200_adamw_multi"] >= 2


def test_switches_restore_the_separate_kernels(dry, monkeypatch):
    from unet_convlstm_b200 import ops, optim
    ops.set_precision("bf16")
    monkeypatch.setattr(ops, "FUSE_BN_POOL", False)
    monkeypatch.setattr(ops, "FUSE_BN_OUTCONV", False)
    monkeypatch.setattr(optim, "ADAMW_PACK", False)
    _train_two_steps(16)
    assert "b200_bn_relu_apply_pool" not in dry and "b200_bn_relu_outconv_fwd" not in dry and "b200_adamw_pack_multi" not in dry
    assert dry["b200_maxpool2_fwd"] == 2 * 4 and dry["b200_maxpool2_bwd"] == 2 * 4
    assert dry["b200_outconv_fwd"] == 2 and dry["b200_outconv_bwd"] == 2
    assert dry["b200_bn_relu_apply"] == 2 * 18
    assert dry["b200_pack_weight"] == 2 * 50          # re-packed at first use after every optimizer step


def test_odd_level_sizes_fall_back_per_stage(dry):
    """48x40: the fourth encoder output is 6x5 -- that stage pools through the separate kernels, the others stay fused."""
    from unet_convlstm_b200 import ops
    ops.set_precision("bf16")
    _train_two_steps(16, size=(48, 40), clip=None)
    assert dry["b200_bn_relu_apply_pool"] == 2 * 3 and dry["b200_maxpool2_fwd"] == 2 and dry["b200_maxpool2_bwd"] == 2


def test_too_small_image_raises_before_any_kernel_runs_on_an_empty_tensor(dry):
    from train.unet import TemporalUNetDualView
    from unet_convlstm_b200 import ops
    ops.set_precision("fp32")
    with pytest.raises(RuntimeError, match="too small"):
        TemporalUNetDualView(base_ch=4)(torch.randn(1, 2, 2, 20, 12))
