"""GPU diagnostic: per-parameter gradient error of the engine vs torch autograd over the trace-mode module path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import diff_pruning_b200 as dp
from diff_pruning_b200.scoring import TaylorScorer

cfgname = sys.argv[1] if len(sys.argv) > 1 else "tiny"
cfg = {"tiny": dp.TINY_TEST_CONFIG, "cifar": dp.CIFAR10_DDPM_CONFIG,
       "lsun_small": dict(dp.LSUN256_DDPM_CONFIG, block_out_channels=(32, 32, 64, 64, 128, 128), sample_size=64)}[cfgname]
B, hw = {"tiny": (2, 16), "cifar": (2, 32), "lsun_small": (2, 64)}[cfgname]
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
m = dp.UNet2DModel(**cfg).eval().cuda()
g = torch.Generator().manual_seed(1)
clean, noise = torch.randn(B, 3, hw, hw, generator=g).cuda(), torch.randn(B, 3, hw, hw, generator=g).cuda()
t = torch.full((B,), 7, device="cuda", dtype=torch.long)
sched = dp.DDPMScheduler()
noisy = sched.add_noise(clean, noise, t)
m.zero_grad()
with dp.trace_mode():
    out_ref = m(noisy, t).sample
    loss_ref = F.mse_loss(out_ref, noise)
    loss_ref.backward()
ref = {k: p.grad.clone() for k, p in m.named_parameters()}
m.zero_grad(set_to_none=True)
sc = TaylorScorer(m, clean, noise, use_graph=False)
loss = sc.step(7)
print("loss", loss.item(), loss_ref.item())
rows = []
for k, p in m.named_parameters():
    e = float((p.grad - ref[k]).norm() / ref[k].norm().clamp_min(1e-30))
    rows.append((e, k, float(p.grad.norm()), float(ref[k].norm())))
for e, k, a, b in rows:
    flag = "  <<<<" if e > 1e-3 else ""
    print(f"{e:10.3e} {a:10.3e} {b:10.3e} {k}{flag}")
