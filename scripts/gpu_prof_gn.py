"""ncu helper: GroupNorm(+SiLU) forward + backward and the attention operand splits / transposes on the shapes that dominate C1 (batch 128) and
C3 (batch 4), one call each inside cudaProfilerStart/Stop.
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_gn python scripts/gpu_prof_gn.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_pruning_b200 import _lib as L   # noqa: E402

lib = L.load()
S = lambda: torch.cuda.current_stream().cuda_stream
SHAPES = [(128, 32 * 32, 128, 32), (128, 16 * 16, 256, 32), (128, 32 * 32, 256, 32), (4, 256 * 256, 128, 32), (4, 128 * 128, 256, 32)]
if len(sys.argv) > 1:
    SHAPES = SHAPES[:int(sys.argv[1])]
runs = []
for N, HW, Cc, G in SHAPES:
    x = torch.randn(N, HW, Cc, device="cuda")
    y, dy, dx, add2 = (torch.randn(N, HW, Cc, device="cuda") for _ in range(4))
    gm, bt = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    dg, db = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    stats = torch.empty(2 * N * G, device="cuda")
    ws = torch.empty(lib.dp_groupnorm_workspace_bytes(N, HW, Cc, G) // 4 + 64, device="cuda")
    fin = torch.empty(2 * N * Cc, device="cuda")
    slots = torch.zeros(2, dtype=torch.int32, device="cuda")
    a = L.GnArgs()
    a.N, a.HW, a.C, a.G, a.eps, a.silu = N, HW, Cc, G, 1e-6, 1
    a.x, a.ldx, a.y, a.ldy = x.data_ptr(), Cc, y.data_ptr(), Cc
    a.gamma, a.beta, a.mean, a.rstd = gm.data_ptr(), bt.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * N * G
    a.workspace, a.amax_y, a.amax_dx = ws.data_ptr(), slots.data_ptr(), slots.data_ptr() + 4
    a.dy, a.lddy, a.dx, a.lddx = dy.data_ptr(), Cc, dx.data_ptr(), Cc
    a.dx_add2, a.ldadd2 = add2.data_ptr(), Cc
    a.dx_add, a.ldadd = dx.data_ptr(), Cc          # accumulate into the existing gradient (aliases dx), as most layers of a pass do
    a.dgamma, a.dbeta, a.fin = dg.data_ptr(), db.data_ptr(), fin.data_ptr()
    runs.append((a, (x, y, dy, dx, add2, gm, bt, dg, db, stats, ws, fin, slots)))
# attention operand split / transpose (C1: 128 images x 256 tokens x 256 channels)
q = torch.randn(128, 256, 256, device="cuda")
hi = torch.empty(128 * 256 * 256, dtype=torch.float16, device="cuda")
lo = torch.empty_like(hi)
P = torch.rand(128, 256, 256, device="cuda")
Pt = torch.empty_like(P)
one = torch.tensor([4.0], device="cuda").view(torch.int32)


def once():
    for a, _ in runs:
        assert lib.dp_groupnorm_fwd(C.byref(a), S()) == 0
        assert lib.dp_groupnorm_bwd(C.byref(a), S()) == 0
        assert lib.dp_groupnorm_bwd_param(C.byref(a), S()) == 0
    assert lib.dp_split_h3(q.data_ptr(), 256, 256 * 256, 128, 256, 256, 0, one.data_ptr(), hi.data_ptr(), lo.data_ptr(), S()) == 0
    assert lib.dp_split_h3(q.data_ptr(), 256, 256 * 256, 128, 256, 256, 1, one.data_ptr(), hi.data_ptr(), lo.data_ptr(), S()) == 0
    assert lib.dp_transpose_batched(P.data_ptr(), Pt.data_ptr(), 128, 256, 256, S()) == 0


once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
# plain timing (no profiler attached: CUDA events, L2 flushed by the 134 MB+ tensors of the neighbouring shapes)
if os.environ.get("GN_TIME", "1") == "1":
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for (a, _), shp in zip(runs, SHAPES):
        tf = tb = 0.0
        for _ in range(5):
            ev[0].record(); lib.dp_groupnorm_fwd(C.byref(a), S()); ev[1].record(); lib.dp_groupnorm_bwd(C.byref(a), S()); ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
        N, HW, Cc, G = shp
        mb = N * HW * Cc * 4 / 1e6
        print(f"gn N={N} HW={HW} C={Cc}: tensor {mb:.1f} MB  fwd {tf / 5 * 1e3:.1f} us ({3 * mb / (tf / 5) / 1e3:.2f} TB/s of 3 tensor passes)  "
              f"bwd {tb / 5 * 1e3:.1f} us ({6 * mb / (tb / 5) / 1e3:.2f} TB/s of 6 tensor passes)")
