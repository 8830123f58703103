"""Device time of single fprop launches across channel widths (pruned vs full): is the pruned network slow because of tile
quantisation or because of unaligned pitches?  Usage: python scripts/time_conv_shapes.py"""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diff_pruning_b200  # noqa: F401,E402
from diff_pruning_b200 import _lib as L  # noqa: E402

lib = L.load()
S = lambda: torch.cuda.current_stream().cuda_stream
R = int(os.environ.get('CONV_R', '3'))
SHAPES = [(128, 128, 32, 128)] if os.environ.get('ONE_SHAPE') else None
for Cin, K, H, N in SHAPES or [(128, 128, 32, 128), (96, 96, 32, 128), (92, 92, 32, 128), (90, 90, 32, 128), (90, 128, 32, 128), (128, 90, 32, 128),
                     (256, 256, 16, 128), (192, 192, 16, 128), (180, 180, 16, 128), (179, 179, 16, 128), (358, 179, 16, 128)]:
    q = int(os.environ.get('DPB200_PITCH', '4'))
    ldx, ldy = (Cin + q - 1) // q * q, (K + q - 1) // q * q
    x = torch.randn(N, H, H, ldx, device="cuda")
    w = torch.randn(K, Cin, R, R, device="cuda") / (Cin * 9) ** 0.5
    y = torch.empty(N, H, H, ldy, device="cuda")
    C4, K4 = lib.dp_tc_weight_row(Cin), lib.dp_tc_weight_row(K)
    packs = [torch.empty(n, device="cuda", dtype=torch.float16) for n in (R * R * K * C4, R * R * K * C4, R * R * Cin * K4, R * R * Cin * K4)]
    slots = torch.zeros(2, dtype=torch.int32, device="cuda")     # [weight amax, activation amax]
    assert lib.dp_pack_conv_weight_tc(w.data_ptr(), K, Cin, R, R, *[p.data_ptr() for p in packs], slots.data_ptr(), S()) == 0
    assert lib.dp_amax(x.data_ptr(), ldx, N * H * H, Cin, slots.data_ptr() + 4, S()) == 0
    a = L.ConvArgs()
    a.N, a.H, a.W, a.C, a.P, a.Q, a.K = N, H, H, Cin, H, H, K
    a.R = a.S = R
    a.stride, a.pad_t, a.pad_l, a.splits = 1, R // 2, R // 2, 1
    a.x, a.ldx, a.y, a.ldy = x.data_ptr(), ldx, y.data_ptr(), ldy
    a.w, a.w_tc_hi, a.w_tc_lo = w.data_ptr(), packs[0].data_ptr(), packs[1].data_ptr()
    a.amax_w, a.amax_x = slots.data_ptr(), slots.data_ptr() + 4
    for _ in range(3):
        assert lib.dp_conv2d_fprop(C.byref(a), S()) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        assert lib.dp_conv2d_fprop(C.byref(a), S()) == 0
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    gf = 2.0 * N * H * H * K * Cin * R * R / 1e9
    stages = ((Cin + 63) // 64) * R * R
    print(f"Cin {Cin:4d} K {K:4d} @{H}x{H}: {us:7.1f} us  {gf / us * 1e3:6.1f} TF algorithmic  ({stages} stages/tile, {(K + 127) // 128} N tile(s), ld {ldx}/{ldy}) -> {us / stages / ((K + 127) // 128):.2f} us per stage-column")
