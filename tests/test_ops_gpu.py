"""Per-kernel parity of the C-ABI (libdpb200.so) against plain torch fp32 CPU math (the ops the reference
dispatches).  Tolerances are fp32-accumulation-order tolerances, stated per test."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from diff_pruning_b200 import _lib as L
    return L.load()


def L_():
    from diff_pruning_b200 import _lib as L
    return L


def S():
    return torch.cuda.current_stream().cuda_stream


def nhwc(x):  # NCHW cpu -> NHWC cuda contiguous
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous().cpu()


def pack(lib, w):
    K, Cin = w.shape[0], w.shape[1]
    R = w.shape[2] if w.dim() == 4 else 1
    Sx = w.shape[3] if w.dim() == 4 else 1
    wd = w.contiguous().cuda()
    wck, wkc = torch.empty(w.numel(), device="cuda"), torch.empty(w.numel(), device="cuda")
    assert lib.dp_pack_conv_weight(wd.data_ptr(), K, Cin, R, Sx, wck.data_ptr(), wkc.data_ptr(), S()) == 0
    return wd, wck, wkc


CONV_CASES = [
    # N, C, H, W, K, R, stride, pad, ld_extra
    (2, 32, 8, 8, 64, 3, 1, 1, 0),
    (3, 3, 16, 16, 32, 3, 1, 1, 0),       # conv_in-like (C=3)
    (2, 48, 8, 8, 3, 3, 1, 1, 0),         # conv_out-like (K=3)
    (2, 32, 16, 16, 32, 3, 2, 0, 0),      # Downsample2D: stride 2, F.pad(0,1,0,1) folded
    (2, 64, 8, 8, 32, 1, 1, 0, 0),        # 1x1 shortcut
    (2, 179, 4, 4, 192, 1, 1, 0, 0),      # pruned attention inner dim (odd)
    (2, 96, 8, 8, 96, 3, 1, 1, 40),       # pruned width, input view inside a wider (concat) buffer
    (1, 130, 5, 7, 70, 3, 1, 1, 0),       # ragged non-power-of-two everything
    (5, 358, 1, 1, 96, 1, 1, 0, 0),       # time_emb_proj on pruned temb (linear)
]


@pytest.mark.parametrize("N,Cin,H,W,K,R,stride,pad,ldx", CONV_CASES)
def test_conv_fprop_dgrad_wgrad(lib, N, Cin, H, W, K, R, stride, pad, ldx):
    L = L_()
    g = torch.Generator().manual_seed(N * 1000 + Cin + K)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(K, Cin, R, R, generator=g) / math.sqrt(Cin * R * R)
    b = torch.randn(K, generator=g)
    if stride == 2 and pad == 0:
        ref_in = F.pad(x, (0, 1, 0, 1))
    else:
        ref_in = x
    xr = ref_in.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    y_ref = F.conv2d(xr, wr, b, stride=stride, padding=pad)
    P, Q = y_ref.shape[2], y_ref.shape[3]
    rowadd = torch.randn(N, K, generator=g)
    res = torch.randn(N, K, P, Q, generator=g)
    y_full = y_ref + rowadd[:, :, None, None] + res
    gy = torch.randn(N, K, P, Q, generator=g)
    y_ref.backward(gy)
    gx_ref = xr.grad[:, :, :H, :W]

    wd, wck, wkc = pack(lib, w)
    # x lives in channels [ldx_off, ldx_off+C) of a wider buffer
    xb = torch.randn(N, H, W, Cin + ldx, generator=g).cuda()
    xb[..., ldx:] = nhwc(x)
    y = torch.zeros(N, P, Q, K, device="cuda")
    a = L.ConvArgs()
    a.N, a.H, a.W, a.C, a.P, a.Q, a.K = N, H, W, Cin, P, Q, K
    a.R = a.S = R
    a.stride, a.pad_t, a.pad_l, a.splits = stride, pad, pad, 1
    a.x, a.ldx, a.y, a.ldy = xb.data_ptr() + 4 * ldx, Cin + ldx, y.data_ptr(), K
    a.w = wck.data_ptr()
    bd, rd, resd = b.cuda(), rowadd.cuda().contiguous(), nhwc(res)
    a.bias, a.rowadd, a.ld_rowadd, a.residual, a.ld_res = bd.data_ptr(), rd.data_ptr(), K, resd.data_ptr(), K
    assert lib.dp_conv2d_fprop(C.byref(a), S()) == 0
    assert rel_err(nchw(y), y_full) < 5e-6
    # accumulate flag: y += conv (no epilogue)
    a.flags, a.bias, a.rowadd, a.residual = 1, None, None, None
    assert lib.dp_conv2d_fprop(C.byref(a), S()) == 0
    assert rel_err(nchw(y), y_full + (y_ref - b[None, :, None, None])) < 5e-6
    # dgrad into a strided view, then accumulate
    gyd = nhwc(gy)
    gxb = torch.zeros(N, H, W, Cin + ldx, device="cuda")
    d = L.ConvArgs()
    C.memmove(C.byref(d), C.byref(a), C.sizeof(a))
    d.flags = 0
    d.x, d.ldx, d.y, d.ldy, d.w = gxb.data_ptr() + 4 * ldx, Cin + ldx, gyd.data_ptr(), K, wkc.data_ptr()
    assert lib.dp_conv2d_dgrad(C.byref(d), S()) == 0
    assert rel_err(nchw(gxb[..., ldx:]), gx_ref) < 5e-6
    assert float(gxb[..., :ldx].abs().sum()) == 0.0
    d.flags = 1
    assert lib.dp_conv2d_dgrad(C.byref(d), S()) == 0
    assert rel_err(nchw(gxb[..., ldx:]), 2 * gx_ref) < 5e-6
    # wgrad with several split counts (deterministic split-K) + reduce accumulating into dW
    for splits in (1, 3):
        ws = torch.empty(splits * K * R * R * Cin, device="cuda")
        wgr = L.ConvArgs()
        C.memmove(C.byref(wgr), C.byref(a), C.sizeof(a))
        wgr.flags, wgr.splits, wgr.y, wgr.ldy, wgr.workspace = 0, splits, gyd.data_ptr(), K, ws.data_ptr()
        assert lib.dp_conv2d_wgrad(C.byref(wgr), S()) == 0
        dw = torch.ones(K, Cin, R, R, device="cuda")
        so, si = torch.zeros(K, device="cuda"), torch.zeros(Cin, device="cuda")
        r = L.WgradReduceArgs()
        r.K, r.C, r.R, r.S, r.splits = K, Cin, R, R, splits
        r.workspace, r.dw, r.w, r.score_out, r.score_in = ws.data_ptr(), dw.data_ptr(), wd.data_ptr(), so.data_ptr(), si.data_ptr()
        assert lib.dp_conv2d_wgrad_reduce(C.byref(r), S()) == 0
        assert rel_err(dw.cpu() - 1, wr.grad) < 1e-5
        pr = (w * wr.grad)
        assert rel_err(so.cpu(), pr.sum((1, 2, 3))) < 1e-4 + 1e-4 and rel_err(si.cpu(), pr.sum((0, 2, 3))) < 2e-4


def test_conv_rejects_bad_arguments(lib):
    L = L_()
    a = L.ConvArgs()
    assert lib.dp_conv2d_fprop(C.byref(a), S()) == -5     # NULL pointers
    t = torch.zeros(16, device="cuda")
    a.x = a.y = a.w = t.data_ptr()
    assert lib.dp_conv2d_fprop(C.byref(a), S()) == -1     # zero extents
    a.N = a.H = a.W = a.C = a.P = a.Q = a.K = a.R = a.S = 1
    a.stride = 3
    assert lib.dp_conv2d_fprop(C.byref(a), S()) == -1     # unsupported stride
    assert b"extents" in lib.dp_strerror(-1)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_batched_layouts(lib, ta, tb):
    L = L_()
    g = torch.Generator().manual_seed(7 + ta * 2 + tb)
    Bn, M, N, K = 3, 70, 50, 45
    A = torch.randn(Bn, K, M, generator=g) if ta else torch.randn(Bn, M, K, generator=g)
    Bm = torch.randn(Bn, N, K, generator=g) if tb else torch.randn(Bn, K, N, generator=g)
    ref = 0.5 * torch.bmm(A.transpose(1, 2) if ta else A, Bm.transpose(1, 2) if tb else Bm)
    Ad, Bd = A.cuda(), Bm.cuda()
    Cd = torch.ones(Bn, M, N, device="cuda")
    a = L.GemmArgs()
    a.M, a.N, a.Kd, a.batch = M, N, K, Bn
    a.A, a.a_rs, a.a_cs, a.a_bs = Ad.data_ptr(), (1 if ta else K), (M if ta else 1), M * K
    a.B, a.b_rs, a.b_cs, a.b_bs = Bd.data_ptr(), (1 if tb else N), (K if tb else 1), N * K
    a.C, a.ldc, a.c_bs, a.alpha, a.accumulate = Cd.data_ptr(), N, M * N, 0.5, 0
    assert lib.dp_gemm_batched(C.byref(a), S()) == 0
    assert rel_err(Cd.cpu(), ref) < 2e-6
    a.accumulate = 1
    assert lib.dp_gemm_batched(C.byref(a), S()) == 0
    assert rel_err(Cd.cpu(), 2 * ref) < 2e-6


@pytest.mark.parametrize("rows,cols", [(37, 16), (64, 256), (5, 1000)])
def test_softmax_fwd_bwd(lib, rows, cols):
    g = torch.Generator().manual_seed(rows)
    s = (3 * torch.randn(rows, cols, generator=g)).requires_grad_(True)
    p = s.softmax(-1)
    gp = torch.randn(rows, cols, generator=g)
    p.backward(gp)
    sd = s.detach().cuda()
    assert lib.dp_softmax_fwd(sd.data_ptr(), sd.data_ptr(), rows, cols, S()) == 0
    assert rel_err(sd.cpu(), p) < 1e-6
    gd = gp.cuda()
    assert lib.dp_softmax_bwd(sd.data_ptr(), gd.data_ptr(), gd.data_ptr(), rows, cols, None, S()) == 0
    assert rel_err(gd.cpu(), s.grad) < 5e-6


def test_copy_rows(lib):
    """dp_copy_rows: strided [rows][cols] device copy (the pack list gathers to_q / to_k / to_v into the fused projection's operand)."""
    g = torch.Generator().manual_seed(5)
    src = torch.randn(37, 200, generator=g).cuda()
    dst = torch.full((40, 190), -7.0, device="cuda")
    assert lib.dp_copy_rows(src.data_ptr() + 4 * 3, 200, dst.data_ptr() + 4 * (2 * 190 + 5), 190, 37, 179, S()) == 0
    assert torch.equal(dst[2:39, 5:184], src[:, 3:182])
    assert bool((dst[:2] == -7).all()) and bool((dst[39:] == -7).all()) and bool((dst[2:39, :5] == -7).all()) and bool((dst[2:39, 184:] == -7).all())
    assert lib.dp_copy_rows(src.data_ptr(), 100, dst.data_ptr(), 190, 2, 179, S()) == -1     # lda < cols: DP_ERR_SHAPE


GN_CASES = [(2, 8, 8, 32, 8, 1, 0), (3, 4, 4, 96, 32, 1, 16), (2, 16, 16, 128, 32, 0, 0), (2, 4, 4, 512, 32, 1, 0),
            (2, 2, 2, 768, 32, 1, 0), (1, 32, 32, 192, 32, 1, 64), (4, 3, 5, 24, 3, 0, 0),
            (2, 8, 8, 30, 3, 1, 0), (2, 8, 8, 64, 8, 1, 2), (2, 4, 4, 358, 2, 1, 0),   # scalar path: C%4!=0 / misaligned view
            (1, 64, 64, 128, 32, 1, 0), (3, 32, 32, 96, 32, 1, 0), (2, 16, 16, 384, 32, 1, 8),   # chunked float4 path: 32 / 7 / 7 chunks per image
            (1, 128, 64, 128, 32, 1, 0), (2, 64, 64, 90, 30, 1, 0),        # 64 / 46 chunks (float4 / scalar): warp-per-group finalize, chunk-reduction launch
            (2, 32, 32, 90, 30, 1, 0), (2, 16, 16, 179, 1, 1, 0)]          # scalar path at pruned widths: 12 chunks (forward folded only) / 6 (both folded)


@pytest.mark.parametrize("N,H,W,Cc,G,silu,ldx", GN_CASES)
def test_groupnorm_fwd_bwd(lib, N, H, W, Cc, G, silu, ldx):
    L = L_()
    g = torch.Generator().manual_seed(Cc + H)
    x = (torch.randn(N, Cc, H, W, generator=g) * 1.5 + 0.3).requires_grad_(True)
    gamma = (torch.randn(Cc, generator=g) * 0.5 + 1).requires_grad_(True)
    beta = torch.randn(Cc, generator=g).requires_grad_(True)
    y = F.group_norm(x, G, gamma, beta, 1e-6)
    if silu:
        y = F.silu(y)
    gy = torch.randn(N, Cc, H, W, generator=g)
    y.backward(gy)
    xb = torch.randn(N, H, W, Cc + ldx, generator=g).cuda()
    xb[..., ldx:] = nhwc(x.detach())
    yd = torch.empty(N, H, W, Cc, device="cuda")
    stats = torch.empty(2 * N * G, device="cuda")
    ws = torch.empty(lib.dp_groupnorm_workspace_bytes(N, H * W, Cc, G) // 4 + 64, device="cuda")
    gm, bt = gamma.detach().cuda(), beta.detach().cuda()
    a = L.GnArgs()
    a.N, a.HW, a.C, a.G, a.eps, a.silu = N, H * W, Cc, G, 1e-6, silu
    a.x, a.ldx, a.y, a.ldy = xb.data_ptr() + 4 * ldx, Cc + ldx, yd.data_ptr(), Cc
    a.gamma, a.beta, a.mean, a.rstd = gm.data_ptr(), bt.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * N * G
    a.workspace = ws.data_ptr()
    assert lib.dp_groupnorm_fwd(C.byref(a), S()) == 0
    assert rel_err(nchw(yd), y) < 3e-6
    gyd = nhwc(gy)
    add = torch.randn(N, H, W, Cc, generator=g).cuda()
    add2 = torch.randn(N, H, W, Cc, generator=g).cuda()
    dx = add.clone()
    dg, db = torch.ones(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    a.dy, a.lddy, a.dx, a.lddx = gyd.data_ptr(), Cc, dx.data_ptr(), Cc
    a.dx_add, a.ldadd, a.dx_add2, a.ldadd2 = dx.data_ptr(), Cc, add2.data_ptr(), Cc
    a.dgamma, a.dbeta = dg.data_ptr(), db.data_ptr()
    assert lib.dp_groupnorm_bwd(C.byref(a), S()) == 0
    assert rel_err(nchw(dx - add - add2), x.grad) < 2e-5
    assert rel_err(dg.cpu() - 1, gamma.grad) < 2e-5 and rel_err(db.cpu() - 1, beta.grad) < 2e-5
    # the same call with a caller-owned `fin`: dx identical, dgamma / dbeta untouched until dp_groupnorm_bwd_param (the engine runs that
    # on its side stream), then bit-identical to the one-call form
    dx2 = add.clone()
    dg2, db2 = torch.ones(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    fin = torch.full((2 * N * Cc,), float("nan"), device="cuda")
    a.dx, a.dx_add, a.dgamma, a.dbeta, a.fin = dx2.data_ptr(), dx2.data_ptr(), dg2.data_ptr(), db2.data_ptr(), fin.data_ptr()
    assert lib.dp_groupnorm_bwd(C.byref(a), S()) == 0
    assert torch.equal(dx2, dx) and bool((dg2 == 1).all()) and bool((db2 == 1).all())
    ws.fill_(float("nan"))          # the shared workspace may be reused before the parameter gradients are taken
    assert lib.dp_groupnorm_bwd_param(C.byref(a), S()) == 0
    assert torch.equal(dg2, dg) and torch.equal(db2, db)


@pytest.mark.parametrize("rows,Cc", [(203, 320), (64, 640), (37, 1280), (130, 96), (9, 1002)])
def test_layernorm_rows_fwd_bwd(lib, rows, Cc):
    """nn.LayerNorm over the channels of every token (ldm attention.py:204-206) = dp_groupnorm with one group over one-pixel images.
    Up to 1280 channels (a multiple of 4) a warp-per-row kernel takes it (1002: the chunked GroupNorm kernels); vs torch fp64, incl. both gradient addends
    and the accumulating dgamma / dbeta."""
    L = L_()
    g = torch.Generator().manual_seed(rows + Cc)
    x = (torch.randn(rows, Cc, generator=g, dtype=torch.float64) * 1.7 + 0.4).requires_grad_(True)
    gamma = (torch.randn(Cc, generator=g, dtype=torch.float64) * 0.5 + 1).requires_grad_(True)
    beta = torch.randn(Cc, generator=g, dtype=torch.float64).requires_grad_(True)
    y = F.layer_norm(x, (Cc,), gamma, beta, 1e-5)
    gy = torch.randn(rows, Cc, generator=g, dtype=torch.float64)
    y.backward(gy)
    ld = Cc + 4
    xb = torch.randn(rows, ld, generator=g).cuda()
    xb[:, 4:] = x.detach().float().cuda()
    yd = torch.empty(rows, Cc, device="cuda")
    stats = torch.empty(2 * rows, device="cuda")
    ws = torch.empty(lib.dp_groupnorm_workspace_bytes(rows, 1, Cc, 1) // 4 + 64, device="cuda")
    gm, bt = gamma.detach().float().cuda(), beta.detach().float().cuda()
    slots = torch.zeros(2, dtype=torch.int32, device="cuda")
    a = L.GnArgs()
    a.N, a.HW, a.C, a.G, a.eps, a.silu = rows, 1, Cc, 1, 1e-5, 0
    a.x, a.ldx, a.y, a.ldy = xb.data_ptr() + 16, ld, yd.data_ptr(), Cc
    a.gamma, a.beta, a.mean, a.rstd, a.workspace = gm.data_ptr(), bt.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * rows, ws.data_ptr()
    a.amax_y = slots.data_ptr()
    n0 = lib.dp_launch_count()
    assert lib.dp_groupnorm_fwd(C.byref(a), S()) == 0
    assert lib.dp_launch_count() - n0 == (1 if Cc % 4 == 0 else 2)     # row kernel | chunked stats + apply (finalize folded in)
    assert rel_err(yd.cpu().double(), y.detach()) < 2e-6
    assert slots.view(torch.float32)[0].item() == float(yd.abs().max())
    gyd, add2 = gy.float().cuda(), torch.randn(rows, Cc, generator=g).cuda()
    add = torch.randn(rows, Cc, generator=g).cuda()
    dx = add.clone()
    dg, db = torch.ones(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    a.amax_y, a.amax_dx = None, slots.data_ptr() + 4
    a.dy, a.lddy, a.dx, a.lddx = gyd.data_ptr(), Cc, dx.data_ptr(), Cc
    a.dx_add, a.ldadd, a.dx_add2, a.ldadd2 = dx.data_ptr(), Cc, add2.data_ptr(), Cc
    a.dgamma, a.dbeta = dg.data_ptr(), db.data_ptr()
    assert lib.dp_groupnorm_bwd(C.byref(a), S()) == 0
    assert rel_err((dx - add - add2).cpu().double(), x.grad) < 2e-5
    assert rel_err(dg.cpu().double() - 1, gamma.grad) < 2e-5 and rel_err(db.cpu().double() - 1, beta.grad) < 2e-5
    assert slots.view(torch.float32)[1].item() == float(dx.abs().max())


def test_groupnorm_dropout_mask_consistent(lib):
    """Dropout folded behind SiLU: backward regenerates the forward mask; keep-rate ~ 1-p; mean preserved."""
    L = L_()
    N, H, W, Cc, G, p = 2, 16, 16, 64, 8, 0.25
    x = torch.randn(N, H, W, Cc, device="cuda")
    gm, bt = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    stats = torch.empty(2 * N * G, device="cuda")
    ws = torch.empty(lib.dp_groupnorm_workspace_bytes(N, H * W, Cc, G) // 4 + 64, device="cuda")
    a = L.GnArgs()
    a.N, a.HW, a.C, a.G, a.eps, a.silu = N, H * W, Cc, G, 1e-6, 1
    a.x, a.ldx, a.y, a.ldy = x.data_ptr(), Cc, y0.data_ptr(), Cc
    a.gamma, a.beta, a.mean, a.rstd, a.workspace = gm.data_ptr(), bt.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * N * G, ws.data_ptr()
    assert lib.dp_groupnorm_fwd(C.byref(a), S()) == 0
    seed = torch.tensor([12345], device="cuda", dtype=torch.int64)
    a.dropout_p, a.dropout_seed, a.dropout_seed_dev, a.y = p, 77, seed.data_ptr(), y1.data_ptr()
    assert lib.dp_groupnorm_fwd(C.byref(a), S()) == 0
    kept = (y1 != 0) | (y0 == 0)
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.02
    assert torch.allclose(y1[kept], y0[kept] / (1 - p), rtol=1e-6)
    gyd = torch.ones_like(x)
    dx1 = torch.empty_like(x)
    dg, db = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    a.dy, a.lddy, a.dx, a.lddx, a.dgamma, a.dbeta = gyd.data_ptr(), Cc, dx1.data_ptr(), Cc, dg.data_ptr(), db.data_ptr()
    assert lib.dp_groupnorm_bwd(C.byref(a), S()) == 0
    # reference: same mask applied to the upstream gradient, no dropout in the kernel
    mask = kept.float() / (1 - p)
    a.dropout_p = 0.0
    gm2 = (gyd * mask).contiguous()
    dx2 = torch.empty_like(x)
    a.dy, a.dx = gm2.data_ptr(), dx2.data_ptr()
    assert lib.dp_groupnorm_bwd(C.byref(a), S()) == 0
    assert rel_err(dx1, dx2) < 1e-6


def test_pointwise_ops(lib):
    g = torch.Generator().manual_seed(11)
    # silu
    x = torch.randn(1000, generator=g).requires_grad_(True)
    y = F.silu(x)
    gy = torch.randn(1000, generator=g)
    y.backward(gy)
    xd, yd, gd = x.detach().cuda(), torch.empty(1000, device="cuda"), torch.ones(1000, device="cuda")
    assert lib.dp_silu_fwd(xd.data_ptr(), yd.data_ptr(), 1000, S()) == 0
    assert lib.dp_silu_bwd(xd.data_ptr(), gy.cuda().data_ptr(), gd.data_ptr(), 1000, 1, S()) == 0
    assert rel_err(yd.cpu(), y) < 1e-6 and rel_err(gd.cpu() - 1, x.grad) < 2e-6
    # timestep embedding (embeddings.py:22-62), t up to 999
    from diff_pruning_b200.models import Timesteps, sinusoidal_frequencies
    t = torch.tensor([0, 1, 17, 500, 999])
    for flip in (False, True):
        ref = Timesteps(128, flip, 1)(t)
        fr = sinusoidal_frequencies(128, 1).cuda()
        out = torch.empty(5, 128, device="cuda")
        assert lib.dp_timestep_embedding(t.cuda().data_ptr(), fr.data_ptr(), out.data_ptr(), 5, 64, int(flip), S()) == 0
        assert float((out.cpu() - ref).abs().max()) < 2e-6
    # add_noise (scheduling_ddpm.py:408-429), both output layouts
    from diff_pruning_b200.models import ddpm_alphas_cumprod
    from oracle import unet_oracle as orc
    x0, nz = torch.randn(5, 3, 8, 8, generator=g), torch.randn(5, 3, 8, 8, generator=g)
    ac = ddpm_alphas_cumprod()
    ref = orc.add_noise(ac, x0, nz, t)
    for nh in (0, 1):
        out = torch.empty(5 * 3 * 64, device="cuda")
        assert lib.dp_add_noise(x0.cuda().data_ptr(), nz.cuda().data_ptr(), t.cuda().data_ptr(), ac.cuda().data_ptr(),
                                out.data_ptr(), 5, 3, 8, 8, nh, 0, S()) == 0
        got = out.view(5, 8, 8, 3).permute(0, 3, 1, 2).cpu() if nh else out.view(5, 3, 8, 8).cpu()
        assert float((got - ref).abs().max()) < 1e-6
    # layout round trip into a wider buffer
    v = torch.randn(2, 5, 4, 6, generator=g)
    buf = torch.zeros(2, 4, 6, 9, device="cuda")
    assert lib.dp_nchw_to_nhwc(v.cuda().data_ptr(), buf.data_ptr() + 16, 9, 2, 5, 4, 6, S()) == 0
    assert torch.equal(buf[..., 4:].permute(0, 3, 1, 2).cpu(), v)
    back = torch.ones(2, 5, 4, 6, device="cuda")
    assert lib.dp_nhwc_to_nchw(buf.data_ptr() + 16, 9, back.data_ptr(), 2, 5, 4, 6, 1, S()) == 0
    assert torch.equal(back.cpu(), v + 1)
    # mse loss + grad (mean reduction)
    n = 2 * 3 * 33 * 33
    pr, tg = torch.randn(n, generator=g), torch.randn(n, generator=g)
    prr = pr.clone().requires_grad_(True)
    l = F.mse_loss(prr, tg)
    l.backward()
    gr, part, lo = torch.empty(n, device="cuda"), torch.empty(lib.dp_mse_partials(n), device="cuda"), torch.zeros(1, device="cuda")
    assert lib.dp_mse_loss_grad(pr.cuda().data_ptr(), tg.cuda().data_ptr(), gr.data_ptr(), n, 1.0 / n, 2.0 / n,
                                part.data_ptr(), lo.data_ptr(), S()) == 0
    assert abs(lo.item() - l.item()) < 2e-6 * l.item() and rel_err(gr.cpu(), prr.grad) < 1e-6
    # upsample x2 and its backward
    u = torch.randn(2, 6, 3, 5, generator=g).requires_grad_(True)
    up = F.interpolate(u, scale_factor=2.0, mode="nearest")
    gu = torch.randn_like(up)
    up.backward(gu)
    ud = nhwc(u.detach())
    upd = torch.empty(2, 6, 10, 6, device="cuda")
    assert lib.dp_upsample2x_fwd(ud.data_ptr(), 6, upd.data_ptr(), 6, 2, 3, 5, 6, S()) == 0
    assert torch.equal(nchw(upd), up.detach())
    dxd = torch.ones(2, 3, 5, 6, device="cuda")
    assert lib.dp_upsample2x_bwd(nhwc(gu).data_ptr(), 6, dxd.data_ptr(), 6, 2, 3, 5, 6, 1, S()) == 0
    assert rel_err(nchw(dxd) - 1, u.grad) < 1e-6
    # segmented column sums
    m = torch.randn(96, 70, generator=g)
    md = m.cuda()
    o = torch.ones(4, 70, device="cuda")
    assert lib.dp_colsum(md.data_ptr(), 70, 96, 70, 24, o.data_ptr(), 70, 1, S()) == 0
    assert rel_err(o.cpu() - 1, m.view(4, 24, 70).sum(1)) < 1e-6
    # float4 path: 16-byte aligned view of 180 columns inside a 192-float pitch (ragged last column block), 300-row segments
    mw = torch.randn(900, 192, generator=g)
    mwd = mw.cuda()
    o2 = torch.zeros(3, 180, device="cuda")
    assert lib.dp_colsum(mwd.data_ptr() + 16, 192, 900, 180, 300, o2.data_ptr(), 180, 0, S()) == 0
    assert rel_err(o2.cpu(), mw[:, 4:184].reshape(3, 300, 180).sum(1)) < 1e-6
    # add views / scale
    a_, b_ = torch.randn(10, 7, generator=g), torch.randn(10, 7, generator=g)
    yv = torch.zeros(10, 9, device="cuda")
    assert lib.dp_add_views(a_.cuda().data_ptr(), 7, b_.cuda().data_ptr(), 7, yv.data_ptr(), 9, 10, 7, S()) == 0
    assert torch.allclose(yv[:, :7].cpu(), a_ + b_)
    assert lib.dp_scale(yv.data_ptr(), 90, 0.5, S()) == 0
    assert torch.allclose(yv[:, :7].cpu(), 0.5 * (a_ + b_))


def test_taylor_reduce_matches_reference_formulas(lib):
    """importance.py:385-418 per-layer reductions (all variants), conv / linear / GroupNorm gamma."""
    from diff_pruning_b200.scoring import taylor_layer_scores
    g = torch.Generator().manual_seed(5)
    for shape in [(40, 24, 3, 3), (33, 50), (96,)]:
        w, dw = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
        sc = taylor_layer_scores(w.cuda(), dw.cuda())
        if len(shape) == 1:
            assert rel_err(sc["out_abs"].cpu(), (w * dw).abs()) < 1e-6
            continue
        po = (w * dw).flatten(1)
        pi = (w * dw).transpose(0, 1).flatten(1)
        for key, ref in [("out_signed", po.sum(1)), ("out_abs", po.abs().sum(1)), ("out_sq", po.pow(2).sum(1)),
                         ("in_signed", pi.sum(1)), ("in_abs", pi.abs().sum(1)), ("in_sq", pi.pow(2).sum(1))]:
            assert rel_err(sc[key].cpu(), ref) < 5e-6, key


def test_adam_clip_ema_matches_torch(lib):
    """ddpm_train.py:462-469: clip_grad_norm_(1.0) -> Adam(lr 2e-4) -> EMA(0.9999), 3 steps."""
    L = L_()
    g = torch.Generator().manual_seed(3)
    n = 10007
    p0 = torch.randn(n, generator=g)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=2e-4, betas=(0.9, 0.999), eps=1e-8)
    ema_ref = p0.clone()
    pd, md, vd, ed = p0.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), p0.cuda()
    ss, part = torch.zeros(1, device="cuda"), torch.empty(lib.dp_sumsq_partials(n), device="cuda")
    scal = torch.zeros(2, device="cuda")
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * (10.0 if step < 3 else 1e-3)   # clipped, clipped, unclipped
        pt.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([pt], 1.0)
        opt.step()
        ema_ref = (1 - 0.9999) * pt.detach() + 0.9999 * ema_ref
        gd = gr.cuda()
        assert lib.dp_sumsq(gd.data_ptr(), n, part.data_ptr(), ss.data_ptr(), S()) == 0
        assert abs(ss.item() - float((gr.double() ** 2).sum())) < 1e-5 * ss.item()
        a = L.AdamArgs()
        a.n, a.p, a.g, a.m, a.v, a.ema, a.sumsq = n, pd.data_ptr(), gd.data_ptr(), md.data_ptr(), vd.data_ptr(), ed.data_ptr(), ss.data_ptr()
        a.max_norm, a.lr, a.beta1, a.beta2, a.eps, a.ema_decay, a.step, a.grad_scale = 1.0, 2e-4, 0.9, 0.999, 1e-8, 0.9999, step, 1.0
        if step == 2:   # device-scalar bias corrections (graph-replay form)
            scal.copy_(torch.tensor([2e-4 / (1 - 0.9 ** step), math.sqrt(1 - 0.999 ** step)]))
            a.step_scalars, a.step = scal.data_ptr(), 1
        assert lib.dp_adam_clip_ema(C.byref(a), S()) == 0
        assert rel_err(pd.cpu(), pt.detach()) < 1e-6 and rel_err(ed.cpu(), ema_ref) < 1e-6


def test_exp_importance_variants_on_device():
    """ddpm_exp criteria (FullTaylor order 1/2, AbsTaylor, Fisher — importance.py:438-781) through dp_taylor_reduce against the unmodified
    vendored classes (tests/golden/exp_importance_tiny.pt): gradients from two accumulated engine passes."""
    from conftest import expand, load_golden
    import diff_pruning_b200 as dp
    from diff_pruning_b200.scoring import EXP_VARIANTS, TaylorScorer, group_importance
    G = load_golden("exp_importance_tiny.pt")
    torch.manual_seed(0)
    m = dp.UNet2DModel(**G["cfg"]).eval().cuda()
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    clean, noise = torch.randn(2, 3, 16, 16, generator=g1), torch.randn(2, 3, 16, 16, generator=g2)
    m.zero_grad()
    sc = TaylorScorer(m, clean.cuda(), noise.cuda(), use_graph=False)
    sc.step(7); sc.step(400)
    w = {k: p for k, p in m.named_parameters()}
    dw = {k: p.grad for k, p in m.named_parameters()}
    for g in G["groups"]:
        items = [(n, k, expand(i)) for n, k, i in g["items"]]
        for variant in EXP_VARIANTS:
            got = group_importance(items, w, dw, variant).cpu()
            ref = g["imp"][variant]
            assert float((got - ref).abs().max()) <= 5e-4 * float(ref.abs().max()) + 1e-12, (g["root"], variant)
