"""2-GPU NCCL check of the sharded paths (skipped with < 2 GPUs): timestep-sharded Taylor scoring == sequential
accumulation; DDP-style finetune step (grad all-reduce mean) == single-GPU step on the concatenated batch."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import diff_pruning_b200 as dp
    from diff_pruning_b200.scoring import FinetuneStepper, TaylorScorer
    cfg = dp.TINY_TEST_CONFIG
    g = torch.Generator().manual_seed(1)
    clean, noise = torch.randn(4, 3, 16, 16, generator=g), torch.randn(4, 3, 16, 16, generator=g)
    ts = list(range(0, 80, 10))
    # ---- scoring: sharded run vs sequential
    torch.manual_seed(0)
    m = dp.UNet2DModel(**cfg).eval().cuda()
    sc = TaylorScorer(m, clean.cuda(), noise.cuda(), use_graph=True)
    losses = sc.run(ts, shard=True)
    arena = sc.plan.grad_arena.clone()
    torch.manual_seed(0)
    m2 = dp.UNet2DModel(**cfg).eval().cuda()
    sc2 = TaylorScorer(m2, clean.cuda(), noise.cuda(), use_graph=False)
    losses2 = sc2.run(ts, shard=False)
    e1 = float((arena - sc2.plan.grad_arena).norm() / sc2.plan.grad_arena.norm())
    e2 = float((losses - losses2).abs().max())
    # ---- finetune: each rank takes half the batch; compare with the full batch on one GPU
    torch.manual_seed(0)
    ma = dp.UNet2DModel(**cfg).cuda().train()
    sa = FinetuneStepper(ma, use_graph=False)
    t = torch.tensor([5, 400, 900, 77])
    lo, hi = rank * 2, rank * 2 + 2
    sa.step(clean[lo:hi].cuda(), noise[lo:hi].cuda(), t[lo:hi].cuda())
    dist.destroy_process_group()   # the single-GPU comparison below must not all-reduce
    torch.manual_seed(0)
    mb = dp.UNet2DModel(**cfg).cuda().train()
    sb = FinetuneStepper(mb, use_graph=False)
    sb.world = 1
    sb.step(clean.cuda(), noise.cuda(), t.cuda())
    # loss = sum_chw mean_b: mean over the full batch == mean of the two half-batch means
    e3 = float((sa.param_arena - sb.param_arena).norm() / sb.param_arena.norm())
    if rank == 0:
        q.put((e1, e2, e3))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharding_matches_single():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    e1, e2, e3 = q.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert e1 < 2e-6 and e2 < 1e-6, (e1, e2)     # same passes, different summation order across ranks
    assert e3 < 2e-6, e3
