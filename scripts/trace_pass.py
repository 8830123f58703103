"""Per-layer pipeline timeline INSIDE a Taylor pass (cold inputs, real layer mix): runs one eager C1 batch-128 pass and, for every
forward conv launch, arms dp_conv_tc_set_trace, runs the launch, and reports its device time, the stages CTA 0 ran and its mean /
median stage period.  Needs a traced kernel: DPB200_TC_PERSISTENT=3 (conv_tc_ps2_kernel) or =4 (conv_tc_pt_kernel).
Usage: DPB200_TC_PERSISTENT=4 python scripts/trace_pass.py [--bwd]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import diff_pruning_b200 as dp  # noqa: E402
from diff_pruning_b200 import _lib as L  # noqa: E402
from diff_pruning_b200.scoring import TaylorScorer  # noqa: E402

lib = L.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = dp.UNet2DModel(**dp.CIFAR10_DDPM_CONFIG).eval().to(dev)
clean, noise = bench.synth_batch(128, seed_off=0)
model.zero_grad()
sc = TaylorScorer(model, clean.to(dev), noise.to(dev), use_graph=False)
for k in range(2):
    sc.step(k)
torch.cuda.synchronize()
p = sc.plan
s_int = torch.cuda.current_stream().cuda_stream
tr = torch.zeros(16640, dtype=torch.int64, device=dev)
rows = []


def run(steps, want):
    for f in steps:
        what = getattr(f, "what", "")
        if what != want:
            f(s_int)
            continue
        tr.zero_()
        lib.dp_conv_tc_set_trace(tr.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(s_int); e1.record()
        torch.cuda.synchronize()
        lib.dp_conv_tc_set_trace(None)
        t = tr.cpu().numpy()[:16384].reshape(1024, 16)
        n = int((t[:, 5] > 0).sum())
        if n < 8:
            rows.append((what, e0.elapsed_time(e1) * 1e3, n, float("nan"), float("nan"), float("nan")))
            continue
        per = np.diff(t[:n, 5])
        tma = (t[:n, 2] - t[:n, 1])[2:]
        rows.append((what, e0.elapsed_time(e1) * 1e3, n, per.mean(), np.median(per), np.median(tma)))


p.t_dev.fill_(3)
L.check(lib.dp_add_noise(sc.clean.data_ptr(), sc.noise.data_ptr(), p.t_dev.data_ptr(), sc.acp.data_ptr(), p.x_in.ptr, sc.B, sc.C, sc.H,
                         sc.W, 1, p.x_in.ld, s_int))
run(p.fwd, "conv fprop")
if "--bwd" in sys.argv:
    gy = p.gradof(p.y_out)
    L.check(lib.dp_mse_loss_grad(p.y_out.ptr, sc.noise_nhwc.data_ptr(), gy.ptr, sc.n, sc.loss_scale, sc.grad_scale,
                                 sc.partial.data_ptr(), sc.loss.data_ptr(), s_int))
    p.gradof(p.silu_temb).t.zero_()
    run(p.bwd_steps, "conv dgrad")
print(f"{'launch':12s} {'us':>9s} {'stages(CTA0)':>13s} {'mean period':>12s} {'median':>8s} {'A-load->split woke (median clk)':>32s}")
for what, us, n, m, med, tma in rows:
    print(f"{what:12s} {us:9.1f} {n:13d} {m:12.0f} {med:8.0f} {tma:32.0f}")
tot = sum(r[1] for r in rows)
print(f"total {tot / 1e3:.2f} ms over {len(rows)} traced launches (tensor-bound ideal: 768 clk per 32-float stage)")
