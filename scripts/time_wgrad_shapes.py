"""Device time of single tensor-core wgrad launches across channel widths / pixel strides (pruned vs full).
Usage: python scripts/time_wgrad_shapes.py"""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diff_pruning_b200  # noqa: F401,E402
from diff_pruning_b200 import _lib as L  # noqa: E402
from diff_pruning_b200.engine import _wgrad_splits  # noqa: E402

lib = L.load()
S = lambda: torch.cuda.current_stream().cuda_stream
R = 3
for Cin, K, H, N, ldx, ldy in [(128, 128, 32, 128, 128, 128), (96, 96, 32, 128, 96, 96), (96, 96, 32, 128, 128, 128), (96, 128, 32, 128, 96, 128),
                               (128, 96, 32, 128, 128, 96), (256, 256, 16, 128, 256, 256), (192, 192, 16, 128, 192, 192), (192, 192, 16, 128, 256, 256)]:
    x = torch.randn(N, H, H, ldx, device="cuda")
    dy = torch.randn(N, H, H, ldy, device="cuda")
    rows = N * H * H
    tiles = ((K + 127) // 128) * ((Cin + 127) // 128) * R * R
    splits = _wgrad_splits(tiles, rows // 64)
    ws = torch.empty(splits * K * R * R * Cin, device="cuda")
    slots = torch.zeros(2, dtype=torch.int32, device="cuda")
    assert lib.dp_amax(x.data_ptr(), ldx, rows, Cin, slots.data_ptr(), S()) == 0
    assert lib.dp_amax(dy.data_ptr(), ldy, rows, K, slots.data_ptr() + 4, S()) == 0
    a = L.ConvArgs()
    a.N, a.H, a.W, a.C, a.P, a.Q, a.K = N, H, H, Cin, H, H, K
    a.R = a.S = R
    a.stride, a.pad_t, a.pad_l, a.splits = 1, 1, 1, splits
    a.x, a.ldx, a.y, a.ldy, a.workspace = x.data_ptr(), ldx, dy.data_ptr(), ldy, ws.data_ptr()
    a.amax_x, a.amax_y = slots.data_ptr(), slots.data_ptr() + 4
    for _ in range(3):
        assert lib.dp_conv2d_wgrad(C.byref(a), S()) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        assert lib.dp_conv2d_wgrad(C.byref(a), S()) == 0
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    gf = 2.0 * rows * K * Cin * 9 / 1e9
    print(f"Cin {Cin:4d} K {K:4d} @{H}x{H} ld {ldx}/{ldy} splits {splits}: {us:7.1f} us  {gf / us * 1e3:6.1f} TF algorithmic")
