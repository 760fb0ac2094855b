import sys, os, torch
sys.path.insert(0, os.getcwd())
from unet_convlstm_b200 import ops
w = torch.randn(4096, 2048, 3, 3, device="cuda")
dw = torch.randn(9, 4096, 2048, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("lstm pack   %.3f ms" % t(lambda: ops.pack_lstm_weight(w, None, torch.bfloat16)))
print("dgrad pack  %.3f ms" % t(lambda: ops.pack_conv_weight_dgrad(w, torch.bfloat16)))
print("fwd pack    %.3f ms" % t(lambda: ops.pack_conv_weight(w, torch.bfloat16)))
print("unpack      %.3f ms" % t(lambda: ops.unpack_conv_wgrad(dw, 2048)))
