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
