"""ncu helper: one eager finetune step on the ratio-0.3 pruned network inside cudaProfilerStart/Stop.
usage: python scripts/gpu_prof_finetune.py [c1|c3] [fp32|bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diff_pruning_b200.scoring import FinetuneStepper
cfg = sys.argv[1] if len(sys.argv) > 1 else "c1"
compute = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda", 0)
c = bench.CONFIGS[cfg]
m = bench.pruned_model(cfg, dev, 0.3)
for mod in m.modules():
    if isinstance(mod, torch.nn.Dropout):
        mod.p = 0.1
m.train()
st = FinetuneStepper(m, use_graph=False, compute=compute)
g = torch.Generator().manual_seed(7)
B, hw = c["batch"], c["hw"]
clean, noise = torch.randn(B, 3, hw, hw, generator=g).to(dev), torch.randn(B, 3, hw, hw, generator=g).to(dev)
t = torch.randint(0, 1000, (B,), generator=g).to(dev)
for _ in range(2):
    st.step(clean, noise, t)
torch.cuda.synchronize()
torch.cuda.profiler.start()
st.step(clean, noise, t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
