"""Pipeline timeline of conv_tc_ps_kernel (dp_conv_tc_set_trace): per-stage clock64() stamps of CTA 0, C1's dominant layer
shape (128 -> 128 3x3 @ 32x32, batch 128).  Prints per-actor intervals in SM clocks.  Usage: python scripts/trace_conv.py [Cin K H N]"""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import diff_pruning_b200  # noqa: F401,E402  (package shim)
from diff_pruning_b200 import _lib as L

lib = L.load()
Cin, K, H, N = (int(v) for v in (sys.argv[1:5] + ["128", "128", "32", "128"][len(sys.argv) - 1:]))
R = 3
S = lambda: torch.cuda.current_stream().cuda_stream
x = torch.randn(N, H, H, Cin, device="cuda")
w = torch.randn(K, Cin, R, R, device="cuda") / (Cin * 9) ** 0.5
y = torch.empty(N, H, H, K, device="cuda")
packs = [torch.empty(R * R * K * Cin, device="cuda") for _ in range(4)]
assert lib.dp_pack_conv_weight_tc(w.data_ptr(), K, Cin, R, R, *[p.data_ptr() for p in packs], S()) == 0
a = L.ConvArgs()
a.N, a.H, a.W, a.C, a.P, a.Q, a.K = N, H, H, Cin, H, H, K
a.R = a.S = R
a.stride, a.pad_t, a.pad_l, a.splits = 1, 1, 1, 1
a.x, a.ldx, a.y, a.ldy = x.data_ptr(), Cin, y.data_ptr(), K
a.w, a.w_tc_hi, a.w_tc_lo = w.data_ptr(), packs[0].data_ptr(), packs[1].data_ptr()
for _ in range(3):
    assert lib.dp_conv2d_fprop(C.byref(a), S()) == 0
torch.cuda.synchronize()
tr = torch.zeros(16640, dtype=torch.int64, device="cuda")
lib.dp_conv_tc_set_trace(tr.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
assert lib.dp_conv2d_fprop(C.byref(a), S()) == 0
e1.record()
torch.cuda.synchronize()
lib.dp_conv_tc_set_trace(None)
raw = tr.cpu().numpy()
import os
os.makedirs('gpurun_out', exist_ok=True)
np.save(os.environ.get('TRACE_OUT', 'gpurun_out/trace_raw.npy'), raw)
t = raw[:16384].reshape(1024, 16)
n = int((t[:, 0] > 0).sum())
t = t[:n].astype(np.int64)
t0 = t[0, 0]
print(f"kernel {e0.elapsed_time(e1) * 1e3:.1f} us; CTA0 ran {n} stages; span {t[n - 1, 5] - t0} clk => {(t[n - 1, 5] - t0) / n:.0f} clk/stage (MMA-bound ideal: 24 clk per K-float = 768 @BK32, 384 @BK16)")
names = ["P_wake", "P_issued", "S_wake", "S_arrived", "M_wake", "M_issued", "S_fenced"]
print("first 16 stages, clocks relative to start:")
print("  g " + " ".join(f"{k:>9}" for k in names))
for g in range(min(16, n)):
    print(f"{g:3d} " + " ".join(f"{t[g, k] - t0:9d}" for k in range(7)))
lo = min(40, n // 3)
hi = n - 4
sl = slice(lo, hi)
def st(name, v):
    print(f"{name:46s} mean {v.mean():8.0f}  p10 {np.percentile(v, 10):7.0f}  p50 {np.percentile(v, 50):7.0f}  p90 {np.percentile(v, 90):7.0f}")
print(f"steady state, stages {lo}..{hi}:")
st("stage period (M_issued[g+1]-M_issued[g])", np.diff(t[sl, 5]))
st("P issue cost (P_issued-P_wake)", t[sl, 1] - t[sl, 0])
st("TMA latency (S_wake-P_issued)", t[sl, 2] - t[sl, 1])
st("split work (S_fenced-S_wake)", t[sl, 6] - t[sl, 2])
st("arrive (S_arrived-S_fenced)", t[sl, 3] - t[sl, 6])
st("conv->MMA hop (M_wake-S_arrived)", t[sl, 4] - t[sl, 3])
st("MMA issue (M_issued-M_wake)", t[sl, 5] - t[sl, 4])
D = int(__import__("os").environ.get("PS_DEPTH", "7"))
st("MMA issued -> slot refilled (P_wake[g+D]-M_issued[g])", t[lo + D:hi, 0] - t[lo:hi - D, 5])
st("splitter idle before wake (S_wake[g+1]-S_arrived[g])", t[lo + 1:hi, 2] - t[lo:hi - 1, 3])
st("MMA lane idle (M_wake[g+1]-M_issued[g])", t[lo + 1:hi, 4] - t[lo:hi - 1, 5])
st("producer idle (P_wake[g+1]-P_issued[g])", t[lo + 1:hi, 0] - t[lo:hi - 1, 1])
