"""The import-compatible `diffusers` / `torch_pruning` surface (compat/), CPU structure checks."""
import os
import sys

import torch

from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, "diff-pruning_b200", "compat"))


def test_compat_names_and_counts():
    import diffusers
    import torch_pruning as tp
    from diffusers.models.resnet import Downsample2D, Upsample2D   # ddpm_prune.py:112
    import diff_pruning_b200 as dp
    assert diffusers.UNet2DModel is dp.UNet2DModel and Downsample2D is not None and Upsample2D is not None
    torch.manual_seed(0)
    m = diffusers.UNet2DModel(**dp.CIFAR10_DDPM_CONFIG).eval()
    ex = {"sample": torch.randn(1, 3, 32, 32), "timestep": torch.ones((1,)).long()}
    macs, params = tp.utils.count_ops_and_params(m, ex)
    G = load_golden("cifar_cfg1.pt")["variants"]["vendored"]
    assert params == G["base"][1] == 35746307
    assert macs == G["base"][0]                                   # same counter convention as the reference (6.064 G)
    # magnitude pruning through the unchanged call sequence of ddpm_prune.py:79-116 reaches the published architecture size
    pr = tp.pruner.MagnitudePruner(m, ex, importance=tp.importance.MagnitudeImportance(), iterative_steps=1, channel_groups={},
                                   ch_sparsity=0.3, ignored_layers=[m.conv_out])
    for g in pr.step(interactive=True):
        g.prune()
    for mod in m.modules():
        if isinstance(mod, (Upsample2D, Downsample2D)):
            mod.channels = mod.conv.in_channels
    macs2, params2 = tp.utils.count_ops_and_params(m, ex)
    assert params2 == G["pruned"][1] == 19851157 and macs2 == G["pruned"][0]    # 3.392 G (assets/exp.png)


def _run(code_or_args, env_extra=None, timeout=600):
    import subprocess
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "diff-pruning_b200", "compat"), ROOT, env.get("PYTHONPATH", "")])
    env.update(env_extra or {})
    return subprocess.run([sys.executable] + code_or_args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_script_import_lines_resolve():
    """Every import statement of ddpm_prune.py:1-13, ddpm_train.py:9-23 and ddpm_sample.py:1-5 that names diffusers / accelerate /
    torch_pruning resolves against compat/ (the VERDICT round-1 finding: DiffusionPipeline, accelerate, diffusers.optimization /
    training_utils / utils raised ImportError)."""
    code = ("from diffusers import DiffusionPipeline, DDPMPipeline, DDIMPipeline, DDIMScheduler, DDPMScheduler, UNet2DModel\n"
            "from diffusers.models import UNet2DModel\nimport torch_pruning as tp\nimport accelerate\nimport diffusers\n"
            "from accelerate import Accelerator\nfrom accelerate.logging import get_logger\nfrom accelerate.utils import ProjectConfiguration\n"
            "from diffusers.optimization import get_scheduler\nfrom diffusers.training_utils import EMAModel\n"
            "from diffusers.utils import is_accelerate_version, is_tensorboard_available, is_wandb_available\n"
            "from diffusers.models.resnet import Upsample2D, Downsample2D\n"
            "diffusers.utils.logging.set_verbosity_info(); diffusers.utils.logging.set_verbosity_error()\n"
            "a = Accelerator(gradient_accumulation_steps=1, mixed_precision='no', log_with='tensorboard', project_dir='/tmp/x', "
            "project_config=ProjectConfiguration())\n"
            "assert a.num_processes == 1 and a.is_main_process and a.is_local_main_process and a.sync_gradients\n"
            "get_logger('t', log_level='INFO').info(a.state, main_process_only=False)\n"
            "assert is_accelerate_version('>=', '0.17.0.dev0')\nprint('OK')")
    r = _run(["-c", code])
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_reference_written_pruned_pickle_loads_with_bare_torch_load():
    """`unet = torch.load(args.pruned_model_ckpt, map_location='cpu')` (ddpm_train.py:292, ddpm_sample.py:27) on a whole-module pickle
    WRITTEN BY THE REFERENCE (tests/golden/ref_pruned_small.pth, tools/gen_golden.py gen_ref_pickle): classes resolve to this package's
    modules, pruned widths and the stale attention scale survive, eps_hat equals the reference's (trace mode = torch CPU ops)."""
    code = ("import torch, diffusers\nimport diff_pruning_b200 as dp\nfrom diff_pruning_b200.models import Attention\n"
            "m = torch.load('tests/golden/ref_pruned_small.pth', map_location='cpu').eval()\n"
            "ref = torch.load('tests/golden/ref_pruned_small_out.pt', weights_only=False)\n"
            "assert type(m) is dp.UNet2DModel\n"
            "assert {k: list(v.shape) for k, v in m.state_dict().items()} == ref['shapes']\n"
            "assert {n: a.scale for n, a in m.named_modules() if isinstance(a, Attention)} == ref['scales']\n"
            "g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)\n"
            "clean, noise = torch.randn(2,3,16,16,generator=g1), torch.randn(2,3,16,16,generator=g2)\n"
            "s = diffusers.DDPMScheduler(num_train_timesteps=1000)\n"
            "with dp.trace_mode(), torch.no_grad():\n    out = m(s.add_noise(clean, noise, ref['t']), ref['t']).sample\n"
            "err = float((out - ref['eps']).abs().max() / ref['eps'].abs().max())\nassert err < 1e-5, err\n"
            "import io\nb = io.BytesIO(); torch.save(m, b); b.seek(0)\nassert type(torch.load(b)) is dp.UNet2DModel\nprint('OK')")
    r = _run(["-c", code])
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_deprecated_attention_keys_are_converted(tmp_path):
    """Hub DDPM checkpoints store attention projections as query/key/value/proj_attn (modeling_utils.py:809-851)."""
    import diff_pruning_b200 as dp
    from diff_pruning_b200 import checkpoint
    torch.manual_seed(0)
    m = dp.UNet2DModel(**dp.TINY_TEST_CONFIG)
    m.save_pretrained(str(tmp_path))
    sd = torch.load(os.path.join(tmp_path, checkpoint.WEIGHTS_NAME), weights_only=True)
    old = {}
    ren = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}
    for k, v in sd.items():
        for new, dep in ren.items():
            if f".{new}." in k and ".attentions." in k:
                k = k.replace(f".{new}.", f".{dep}.")
        old[k] = v
    assert any(".query." in k for k in old) and not any(".to_q." in k for k in old)
    torch.save(old, os.path.join(tmp_path, checkpoint.WEIGHTS_NAME))
    m2 = dp.UNet2DModel.from_pretrained(str(tmp_path))
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_ema_model_and_get_scheduler_match_reference_semantics():
    from diffusers.optimization import get_scheduler
    from diffusers.training_utils import EMAModel
    import json
    p = [torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.zeros(2), requires_grad=False)]
    ema = EMAModel(p, decay=0.9, use_ema_warmup=False, inv_gamma=1.0, power=0.75)
    with torch.no_grad():
        p[0].add_(1.0); p[1].add_(5.0)
    ema.step(p)
    assert ema.cur_decay_value == 0.9 and ema.optimization_step == 1            # constant decay from step 1 (training_utils.py:201)
    assert torch.allclose(ema.shadow_params[0], torch.full((3,), 0.1 * 2 + 0.9 * 1.0)) and torch.equal(ema.shadow_params[1], p[1].data)
    ema.store(p); ema.copy_to(p)
    assert torch.allclose(p[0].data, torch.full((3,), 1.1))
    ema.restore(p)
    assert torch.allclose(p[0].data, torch.full((3,), 2.0))
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "lr_schedules.json")))
    for key, vals in G.items():
        name, warm, total = key.split("|")
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
        sch = get_scheduler(name, opt, num_warmup_steps=int(warm), num_training_steps=int(total))
        for s in range(max(int(k) for k in vals) + 1):
            if str(s) in vals:
                assert abs(sch.get_last_lr()[0] - vals[str(s)]) < 1e-12, (key, s)
            opt.step(); sch.step()


def test_ddpm_prune_script_runs_unmodified_magnitude(tmp_path):
    """/root/reference/ddpm_prune.py executed AS IS (runpy) with compat/ first on sys.path, `--pruner magnitude --device cpu` on a
    saved random-init pipeline; the module tree runs in trace mode (CPU build container).  The script's last block hard-codes
    `pipeline.to("cuda")` (ddpm_prune.py:144), so without a GPU it stops there — after everything this test checks was written."""
    import pytest
    script = "/root/reference/ddpm_prune.py"
    if not os.path.isfile(script):
        pytest.skip("reference checkout not present (GPU box)")
    import diff_pruning_b200 as dp
    torch.manual_seed(0)
    m = dp.UNet2DModel(**dp.TINY_TEST_CONFIG)
    src, dst = str(tmp_path / "tiny_cifar_pipeline"), str(tmp_path / "pruned")
    dp.DDPMPipeline(unet=m, scheduler=dp.DDPMScheduler(num_train_timesteps=1000)).save_pretrained(src)
    r = _run([os.path.join(ROOT, "tests", "run_script_traced.py"), script, "--model_path", src, "--save_path", dst, "--pruner", "magnitude",
              "--pruning_ratio", "0.3", "--device", "cpu", "--batch_size", "2"])
    assert "#Params" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])
    if r.returncode != 0:
        assert "cuda" in r.stderr.lower(), r.stderr[-2000:]
    for rel in ("model_index.json", "unet/config.json", "scheduler/scheduler_config.json", "pruned/unet_pruned.pth"):
        assert os.path.isfile(os.path.join(dst, rel)), rel
    sys.path.insert(0, os.path.join(ROOT, "diff-pruning_b200", "compat"))
    import diffusers  # noqa: F401  (registers the allow-list for whole-module pickles)
    pm = torch.load(os.path.join(dst, "pruned", "unet_pruned.pth"), map_location="cpu")
    # the same call sequence in-process gives the same architecture
    import torch_pruning as tp
    ex = {"sample": torch.randn(1, 3, 32, 32), "timestep": torch.ones((1,)).long()}
    torch.manual_seed(0)
    m2 = dp.UNet2DModel(**dp.TINY_TEST_CONFIG).eval()
    pr = tp.pruner.MagnitudePruner(m2, ex, importance=tp.importance.MagnitudeImportance(), iterative_steps=1, channel_groups={},
                                   ch_sparsity=0.3, ignored_layers=[m2.conv_out])
    for g in pr.step(interactive=True):
        g.prune()
    assert {k: tuple(v.shape) for k, v in pm.state_dict().items()} == {k: tuple(v.shape) for k, v in m2.state_dict().items()}
    assert sum(p.numel() for p in pm.parameters()) < sum(p.numel() for p in m.parameters())
