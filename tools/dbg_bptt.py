import sys, torch
sys.path.insert(0, ".")
from unet_convlstm_b200 import ops, functional as Fn
T, B, H, W, Cin, Ch = 5, 8, 4, 4, 256, 256
bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(12)
w = torch.randn(4 * Ch, Cin + Ch, 3, 3, device="cuda", generator=g) / (9 * (Cin + Ch)) ** 0.5
gates = torch.rand(T, B, H, W, 4 * Ch, device="cuda", generator=g).to(bf)
c_all = torch.randn(T + 1, B, H, W, Ch, device="cuda", generator=g)
dh_seq = torch.randn(T, B, H, W, Ch, device="cuda", generator=g).to(bf)
wd = ops.pack_conv_weight_dgrad(w, bf)
for trial in range(3):
    junk = torch.full((64 << 20,), float("nan"), device="cuda")  # poison freed memory
    del junk
    dz_all = torch.full((T, B, H, W, 4 * Ch), float("nan"), device="cuda", dtype=bf)
    dx_seq = torch.full((T, B, H, W, Cin), float("nan"), device="cuda", dtype=bf)
    dc_buf = torch.full((2, B, H, W, Ch), float("nan"), device="cuda")
    ops.lstm_gates_bwd(gates[T - 1], c_all[T - 1], c_all[T], dh_seq[T - 1], None, None, dz_all[T - 1], dc_buf[(T - 1) & 1])
    ops.lstm_seq_bwd_fused(dz_all, wd, gates, c_all, dh_seq, dc_buf, dx_seq, None, Cin, False, 3)
    torch.cuda.synchronize()
    print("nan per t: dx", [int(torch.isnan(dx_seq[t].float()).sum()) for t in range(T)],
          "dz", [int(torch.isnan(dz_all[t].float()).sum()) for t in range(T)], "dc", [int(torch.isnan(dc_buf[i]).sum()) for i in range(2)])
