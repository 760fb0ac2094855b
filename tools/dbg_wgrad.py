import sys, torch
sys.path.insert(0, ".")
from unet_convlstm_b200 import ops, _lib
dev = torch.device("cuda"); bf = torch.bfloat16
T, B, HW, Nz, C = [int(v) for v in sys.argv[1:6]]
dz = torch.randn(T, B, HW, HW, Nz, device=dev).to(bf)
src = torch.randn(T, B, HW, HW, C, device=dev).to(bf)
dw = torch.zeros(9, Nz, C, device=dev)
try:
    ops.conv_wgrad(dz, src, 3, dw, 0)
    torch.cuda.synchronize()
    print("ok")
except Exception as e:
    print("FAIL", str(e)[:200])
print("flag", _lib.lib().b200_device_error())
