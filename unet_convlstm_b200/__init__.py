"""B200-native (sm_100a) kernels of the UNet-ConvLSTM hot path and their Python host layer.

    csrc/          hand-written CUDA: tcgen05/TMA implicit-GEMM conv + fused ConvLSTM epilogue, wgrad,
                   CUDA-core check-mode convs, HBM-bound pointwise/normalisation kernels, the C ABI
    _lib.py        ctypes binding generated from include/b200_convlstm.h (no fallback path)
    ops.py         tensor-level wrappers (allocation, dispatch tensor-core vs CUDA-core)
    functional.py  autograd Functions (hand-derived backward of the reference modules)
    dist.py        data-parallel gradient all-reduce (NCCL, bucketed, overlapped with backward)
    loss.py        the reference's compute_loss as two fused kernels
    optim.py       clip_grad_norm_ + AdamW as multi-tensor kernels
    loop.py        train_one_epoch / evaluate with on-device metric accumulators
    data.py        pinned, double-buffered host->device prefetch
    graph.py       optional CUDA-graph replay of the training step

The reference-facing module surface is train/unet.py at the repository root.
"""
from .ops import act_dtype, get_precision, set_precision  # noqa: F401
