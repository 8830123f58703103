"""ncu helper: one eager finetune step on the pruned C1 network inside cudaProfilerStart/Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diff_pruning_b200.scoring import FinetuneStepper
dev = torch.device("cuda", 0)
m = bench.pruned_c1_model(dev)
for mod in m.modules():
    if isinstance(mod, torch.nn.Dropout):
        mod.p = 0.1
m.train()
st = FinetuneStepper(m, use_graph=False)
g = torch.Generator().manual_seed(7)
clean, noise = torch.randn(128, 3, 32, 32, generator=g).to(dev), torch.randn(128, 3, 32, 32, generator=g).to(dev)
t = torch.randint(0, 1000, (128,), generator=g).to(dev)
for _ in range(2):
    st.step(clean, noise, t)
torch.cuda.synchronize()
torch.cuda.profiler.start()
st.step(clean, noise, t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
