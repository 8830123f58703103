"""End-to-end parity of the planned CUDA engine (through the C-ABI) against golden fixtures produced by the
unmodified reference, and against the oracle on fresh seeds."""
import copy

import pytest
import torch
import torch.nn.functional as F

from conftest import expand, load_golden, max_rel, rel_err, worst_grad_err
import diff_pruning_b200 as dp
from diff_pruning_b200.scoring import FinetuneStepper, TaylorScorer, group_importance, select_pruning_idxs
from diff_pruning_b200 import pruning

pytestmark = pytest.mark.gpu


def inputs(b, hw):
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    return torch.randn(b, 3, hw, hw, generator=g1), torch.randn(b, 3, hw, hw, generator=g2)


def build(cfg, seed=0):
    torch.manual_seed(seed)
    return dp.UNet2DModel(**cfg).eval().cuda()


@pytest.mark.parametrize("use_graph", [False, True])
def test_tiny_two_accumulated_passes(use_graph):
    """grads accumulate over passes (no zero_grad) — ddpm_prune.py:90,97-102."""
    G = load_golden("tiny_unet.pt")
    m = build(G["cfg"])
    clean, noise = inputs(2, 16)
    sc = TaylorScorer(m, clean.cuda(), noise.cuda(), use_graph=use_graph)
    losses = [sc.step(7).item(), sc.step(400).item()]
    assert losses == pytest.approx(G["losses"], rel=2e-6)
    worst = worst_grad_err(((k, p.grad) for k, p in m.named_parameters()), G["grads"])
    assert worst < 5e-5, worst


def test_tiny_autograd_boundary_matches_scripts_loop():
    """The unchanged script loop: scheduler.add_noise -> model(...).sample -> F.mse_loss -> loss.backward()."""
    G = load_golden("tiny_unet.pt")
    m = build(G["cfg"])
    sched = dp.DDPMScheduler(num_train_timesteps=1000)
    clean, noise = (t.cuda() for t in inputs(2, 16))
    m.zero_grad()
    losses = []
    for tt in (7, 400):
        t = (tt * torch.ones(2, device="cuda")).long()
        out = m(sched.add_noise(clean, noise, t), t).sample
        loss = F.mse_loss(out, noise)
        loss.backward()
        losses.append(loss.item())
    assert losses == pytest.approx(G["losses"], rel=2e-6)
    assert max_rel(out, G["out_last"]) < 1e-5
    worst = worst_grad_err(((k, p.grad) for k, p in m.named_parameters()), G["grads"])
    assert worst < 5e-5, worst
    with torch.no_grad():   # per-sample timesteps, inference path
        out2 = m(sched.add_noise(clean, noise, G["t2"].cuda()), G["t2"].cuda()).sample
    assert max_rel(out2, G["out_t2"]) < 1e-5
    # python-number timestep like the pipelines pass (pipeline_ddim.py:105)
    with torch.no_grad():
        o3 = m(clean, 5).sample
        o4 = m(clean, torch.tensor(5, device="cuda")).sample
    assert torch.equal(o3, o4)
    # model survives pickling / deepcopy with live plans (ddpm_prune.py:135, op_counter.py:18)
    m2 = copy.deepcopy(m)
    assert "_dpb200_plans" not in m2.__dict__


def test_cifar_eps_and_grads():
    """C1 (CIFAR UNet) seed 0: eps_hat within 1e-4 relative of the reference CPU fp32 (north_star tolerance)."""
    G = load_golden("cifar_fwd.pt")
    m = build(dp.CIFAR10_DDPM_CONFIG)
    sched = dp.DDPMScheduler()
    clean16, noise16 = inputs(16, 32)
    clean, noise = clean16[:2].cuda(), noise16[:2].cuda()
    with torch.no_grad():
        for tt, ref in G["eps_b2"].items():
            t = (tt * torch.ones(2, device="cuda")).long()
            assert max_rel(m(sched.add_noise(clean, noise, t), t).sample, ref) < 1e-4, tt
    m.zero_grad()
    sc = TaylorScorer(m, clean, noise, use_graph=False)
    assert sc.step(500).item() == pytest.approx(G["loss_b2_t500"], rel=5e-6)
    bad = []
    for k, p in m.named_parameters():
        s = G["grad_samples_b2_t500"][k]
        g = p.grad.flatten()[:64].cpu()
        if rel_err(g, s) > 2e-3 and float((g - s).abs().max()) > 1e-7:
            bad.append((k, rel_err(g, s)))
        f = G["grad_fp_b2_t500"][k]
        assert float((p.grad.double() ** 2).sum()) == pytest.approx(f[2], rel=2e-3), k
    assert not bad, bad[:5]


def test_cfg1_scores_and_masks_bit_exact():
    """BASELINE config 1 on the GPU: 100 timesteps B=16, ratio 0.3.  Losses, per-group importance vectors and the
    pruned channel-index sets (all three importance variants) against the unmodified reference; then the pruned
    network's eps_hat."""
    G = load_golden("cifar_cfg1.pt")
    m = build(dp.CIFAR10_DDPM_CONFIG)
    clean, noise = inputs(G["B"], 32)
    m.zero_grad()
    sc = TaylorScorer(m, clean.cuda(), noise.cuda(), use_graph=True)
    losses = sc.run(range(G["n_steps"])).cpu()
    assert rel_err(losses, torch.tensor(G["losses"])) < 5e-6
    for k, p in m.named_parameters():
        assert float((p.grad.double() ** 2).sum()) == pytest.approx(G["grad_fp"][k][2], rel=5e-3), k
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    sched = dp.DDPMScheduler()
    for variant, V in G["variants"].items():
        mv = copy.deepcopy(m)
        for k, p in mv.named_parameters():
            p.grad = grads[k].clone()
        # the whole product pipeline: structural groups in the reference's order -> device-side scores -> selection -> slicing
        rec = pruning.taylor_prune(mv, G["ratio"], variant, ignored_layers=[mv.conv_out])
        assert [r["root"] for r in rec] == [g["root"] for g in V["groups"]]
        worst_imp = 0.0
        for r, g in zip(rec, V["groups"]):
            worst_imp = max(worst_imp, rel_err(r["imp"], g["imp"]))
            assert sorted(r["idxs"]) == sorted(g["idxs"]), (variant, g["root"], worst_imp)   # bit-exact mask
        assert worst_imp < 2e-3, (variant, worst_imp)
        assert {k: list(v.shape) for k, v in mv.state_dict().items()} == V["pruned_shapes"]
        assert sum(p.numel() for p in mv.parameters()) == V["pruned"][1] == 19851157
        with torch.no_grad():
            t = (10 * torch.ones(2, device="cuda")).long()
            out = mv(sched.add_noise(clean[:2].cuda(), noise[:2].cuda(), t), t).sample
        assert max_rel(out, V["pruned_eps_b2_t10"]) < 1e-4, variant
    # the same through the reference's own call sequence (ddpm_prune.py:60,79-87,108-116) on the compat names
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diff-pruning_b200", "compat"))
    import torch_pruning as tp
    from diffusers.models.resnet import Downsample2D, Upsample2D
    mv = copy.deepcopy(m)
    for k, p in mv.named_parameters():
        p.grad = grads[k].clone()
    ex = {"sample": torch.randn(1, 3, 32, 32).cuda(), "timestep": torch.ones((1,)).long().cuda()}
    pr = tp.pruner.MagnitudePruner(mv, ex, importance=tp.importance.TaylorImportance(multivariable=True), iterative_steps=1,
                                   channel_groups={}, ch_sparsity=G["ratio"], ignored_layers=[mv.conv_out])
    got = []
    for g in pr.step(interactive=True):
        got.append(sorted(g.idxs))
        g.prune()
    for mod in mv.modules():
        if isinstance(mod, (Upsample2D, Downsample2D)):
            mod.channels = mod.conv.in_channels
    assert got == [sorted(g["idxs"]) for g in G["variants"]["taylor"]["groups"]]
    assert tp.utils.count_ops_and_params(mv, ex) == (G["variants"]["taylor"]["pruned"][0], 19851157.0)


def test_finetune_two_steps():
    """ddpm_train.py:437-469 (dropout 0): loss, clipped Adam update and EMA after 2 steps vs the reference."""
    G = load_golden("finetune_tiny.pt")
    for use_graph in (False, True):
        torch.manual_seed(0)
        m = dp.UNet2DModel(**G["cfg"]).cuda().train()
        st = FinetuneStepper(m, lr=2e-4, ema_decay=0.9999, max_grad_norm=1.0, use_graph=use_graph)
        for s in G["steps"]:
            loss = st.step(s["clean"].cuda(), s["noise"].cuda(), s["t"].cuda())
            assert loss.item() == pytest.approx(s["loss"], rel=2e-5)
            assert float(st.sumsq.sqrt()) == pytest.approx(s["grad_norm"], rel=2e-4)
        ema = st.ema_state()
        # Adam normalises the step (m/sqrt(v)): a zero-initialised bias moves by ~lr per step, so its relative
        # error after 2 steps equals the gradient's relative rounding error (~1e-5), not 1e-7.
        for k, p in m.named_parameters():
            assert rel_err(p, G["params"][k]) < 5e-5, k
            assert rel_err(ema[k], G["ema"][k]) < 5e-5, k


def test_lsun_family_block_one_pass_vs_oracle():
    """A narrow member of the LSUN-256 family (6 levels, attention at level 4, 64x64 input) vs the oracle."""
    from oracle import unet_oracle as orc
    cfg = dict(dp.LSUN256_DDPM_CONFIG, block_out_channels=(32, 32, 64, 64, 128, 128), sample_size=64)
    torch.manual_seed(3)
    m = dp.UNet2DModel(**cfg).eval()
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    clean, noise = torch.randn(2, 3, 64, 64, generator=g), torch.randn(2, 3, 64, 64, generator=g)
    t = torch.tensor([123, 877])
    ref = orc.taylor_pass(sd, cfg, orc.alphas_cumprod(), clean, noise, t)
    m = m.cuda()
    m.zero_grad()
    sc = TaylorScorer(m, clean.cuda(), noise.cuda(), use_graph=False)
    assert sc.step(t).item() == pytest.approx(ref.item(), rel=5e-6)
    worst = worst_grad_err(((k, p.grad) for k, p in m.named_parameters()), {k: v.grad for k, v in sd.items()})
    assert worst < 1e-4, worst


def test_fused_signed_scores_fall_out_of_backward():
    """North-star item: the wgrad reduce accumulates sum_t sum_k W*dW_t per channel while it adds dW_t into .grad; the
    `multivariable=True` Taylor score |.| computed from that equals the one computed from the accumulated gradient
    (scores are linear in dW: sum_t and sum_k commute)."""
    from diff_pruning_b200.scoring import taylor_layer_scores
    m = build(dp.TINY_TEST_CONFIG)
    clean, noise = inputs(2, 16)
    sc = TaylorScorer(m, clean.cuda(), noise.cuda(), use_graph=True, fused_scores=True)
    for t in (3, 250, 600, 990):
        sc.step(t)
    fused = sc.signed_scores()
    params = dict(m.named_parameters())
    assert len(fused) == sum(1 for p in params.values() if p.dim() >= 2)
    worst = 0.0
    for name, (so, si) in fused.items():
        ref = taylor_layer_scores(params[name], params[name].grad)
        scale = float(ref["out_abs"].max())          # sum |w dw|: the magnitude of the summands (signed sums can cancel)
        worst = max(worst, float((so - ref["out_signed"]).abs().max()) / scale, float((si - ref["in_signed"]).abs().max()) / scale)
    assert worst < 2e-5, worst


def test_lsun256_architecture_one_pass_vs_oracle():
    """BASELINE config 3 architecture (google/ddpm-ema-bedroom-256: 113.7 M params, 3x256x256), batch 1, one Taylor pass on
    the GPU vs the CPU oracle: loss and every parameter gradient."""
    from oracle import unet_oracle as orc
    cfg = dp.LSUN256_DDPM_CONFIG
    torch.manual_seed(0)
    m = dp.UNet2DModel(**cfg).eval()
    assert round(sum(p.numel() for p in m.parameters()) / 1e6, 3) == 113.673
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    clean, noise = torch.randn(1, 3, 256, 256, generator=g), torch.randn(1, 3, 256, 256, generator=g)
    t = torch.tensor([321])
    torch.set_num_threads(min(32, torch.get_num_threads()))
    ref = orc.taylor_pass(sd, cfg, orc.alphas_cumprod(), clean, noise, t)
    m = m.cuda()
    m.zero_grad()
    sc = TaylorScorer(m, clean.cuda(), noise.cuda(), use_graph=False)
    assert sc.step(t).item() == pytest.approx(ref.item(), rel=1e-5)
    worst = worst_grad_err(((k, p.grad) for k, p in m.named_parameters()), {k: v.grad for k, v in sd.items()})
    assert worst < 2e-4, worst


def test_ddim_sampling_matches_reference_pipeline():
    """DDIMPipeline + the reference's modified DDIMScheduler (skip_type, prev_timestep rule) on the device: images of a 10-step
    eta=0 uniform run and a 7-step eta=0.5 quad run (CPU generator for the initial latent and the variance noise, like the
    fixture) within 2e-4 of the reference's CPU pipeline (errors of the 1e-5-grade UNet compound over the chain)."""
    G = load_golden("ddim_tiny.pt")
    from diff_pruning_b200.sampling import DDIMPipeline
    m = build(G["cfg"])
    for name in ("uniform_eta0", "quad_eta05"):
        R = G[name]
        pipe = DDIMPipeline(unet=m, scheduler=dp.DDPMScheduler(num_train_timesteps=1000))
        pipe.scheduler.skip_type = R["skip_type"]
        g = torch.Generator().manual_seed(0)
        out = pipe(batch_size=2, generator=g, eta=R["eta"], num_inference_steps=R["steps"], output_type="numpy").images
        assert torch.equal(pipe.scheduler.timesteps, R["timesteps"])
        assert out.shape == tuple(R["images"].shape) and out.min() >= 0.0 and out.max() <= 1.0
        assert float((torch.from_numpy(out) - R["images"]).abs().max()) < 2e-4, name
