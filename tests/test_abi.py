"""The C-ABI library loads and exports every symbol include/b200_convlstm.h declares (no compute calls:
this runs without a GPU), and the product path refuses CPU tensors instead of falling back."""
import ctypes
import os

import pytest
import torch


def test_header_symbols_exported():
    from unet_convlstm_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 24
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), name
    assert set(_lib.KERNELS_PER_CALL) <= set(protos)


def test_error_reporting_without_gpu():
    from unet_convlstm_b200 import _lib
    lib = _lib.lib()
    assert lib.b200_conv_tc_supported(4, 4, 4, 64, 64, 256, 1) == 1
    assert lib.b200_conv_tc_supported(4, 5, 6, 64, 64, 256, 1) == 0   # W not a power of two
    assert lib.b200_conv_tc_supported(4, 4, 4, 6, 12, 48, 1) == 0     # channels not multiples of 16
    with pytest.raises(RuntimeError, match="bad arguments"):
        _lib.call("b200_bn_stats", None, 1, 1, 1, 1, None, None, None)


def test_no_cpu_fallback():
    from train.unet import ConvLSTM, DoubleConv, TemporalUNetDualView
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TemporalUNetDualView(base_ch=2)(torch.zeros(1, 2, 2, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DoubleConv(2, 4)(torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ConvLSTM(2, 4)([torch.zeros(1, 2, 8, 8)])


def test_product_path_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel in ["train/unet.py"] + [os.path.join("unet_convlstm_b200", f)
                                    for f in os.listdir(os.path.join(root, "unet_convlstm_b200")) if f.endswith(".py")]:
        src = open(os.path.join(root, rel)).read()
        assert "oracle" not in src.replace("CPU oracle", ""), rel


def test_torch_custom_op_library_matches_header():
    """libb200convlstm_torch.so (csrc/torch_ops.cpp, GENERATED from include/b200_convlstm.h by tools/gen_torch_ops.py)
    registers one TORCH_LIBRARY(b200convlstm, ...) op per stream-taking entry point; a written (non-const) pointer of the
    C prototype is a mutable tensor of the op schema, a const pointer an immutable one, and the committed source is
    what the generator produces from the committed header."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_torch_ops.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sys.path.insert(0, os.path.join(root, "tools"))
    try:
        import gen_torch_ops as G
    finally:
        sys.path.pop(0)
    from unet_convlstm_b200 import _lib
    ops = _lib.torch_ops()
    protos = G.parse()
    assert len(protos) >= 30 and set(ops) == {name for name, _ in protos}
    for name, params in protos:
        schema = ops[name][0].default._schema
        assert len(schema.arguments) == len(params), name
        for arg, (ctype, pname) in zip(schema.arguments, params):
            assert arg.name == pname, (name, pname)
            kind = G.classify(ctype)
            written = arg.alias_info is not None and arg.alias_info.is_write
            assert written == (kind in ("t_out", "tlist_out")), (name, pname, ctype)
    # a CPU tensor is refused by the op itself
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.ops.b200convlstm.bn_stats(torch.zeros(4), 1, 1, 4, 1, None, None)
