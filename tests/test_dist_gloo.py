"""world_size-2 gloo test of the timestep-sharding rule: rank r takes t_k with k mod W == r,
one all-reduce(SUM) of the gradient arena reproduces the sequential accumulation.  CPU, uses the oracle as the
per-pass engine (test infrastructure)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import diff_pruning_b200 as dp
    from oracle import unet_oracle as orc
    cfg = dp.TINY_TEST_CONFIG
    torch.manual_seed(0)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in dp.UNet2DModel(**cfg).state_dict().items()}
    ac = orc.alphas_cumprod()
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    clean, noise = torch.randn(2, 3, 16, 16, generator=g1), torch.randn(2, 3, 16, 16, generator=g2)
    ts = list(range(0, 60, 10))
    for k, t in enumerate(ts):
        if k % world == rank:
            orc.taylor_pass(sd, cfg, ac, clean, noise, (t * torch.ones(2)).long())
    arena = torch.cat([v.grad.flatten() for v in sd.values()])
    dist.all_reduce(arena, op=dist.ReduceOp.SUM)
    if rank == 0:
        ref = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
        for t in ts:
            orc.taylor_pass(ref, cfg, ac, clean, noise, (t * torch.ones(2)).long())
        ref_arena = torch.cat([v.grad.flatten() for v in ref.values()])
        q.put(float((arena - ref_arena).norm() / ref_arena.norm()))
    dist.barrier()
    dist.destroy_process_group()


def test_timestep_sharding_allreduce_equals_sequential():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=300)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err < 1e-6
