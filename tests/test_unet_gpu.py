"""GPU parity of the U-Net epilogue kernels (K9/K10) and of the U-Net module built on them."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import unet as ounet

T = torch.from_numpy
SMALL = dict(image_size=64, num_channels=32, num_res_blocks=1, channel_mult="", learn_sigma=True,
             attention_resolutions=16, num_heads=4, num_head_channels=16, use_scale_shift_norm=True,
             resblock_updown=True)


def _gn_ref(x, gamma, beta, scale, shift, mode):
    """fp32 torch reference of the three epilogue variants (nn.py:17-19, unet_ffhq.py:319-323)."""
    h = F.group_norm(x.float(), 32, gamma, beta, 1e-5)
    if mode == 2:
        h = h * (1 + scale[:, :, None, None]) + shift[:, :, None, None]
    return h if mode == 0 else F.silu(h)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cl", [False, True])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", [(2, 128, 16, 16), (3, 384, 8, 8), (1, 256, 64, 64), (2, 96, 5, 7)])
def test_group_norm_epilogue(dtype, cl, mode, shape):
    from nhmc_b200 import kernels
    B, C, H, W = shape
    if cl and C % (4 if dtype == torch.float32 else 8):
        pytest.skip("channels-last path needs C % vec == 0")
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(shape, generator=g) * 1.7 + 0.4).cuda().to(dtype)
    dy = torch.randn(shape, generator=g).cuda().to(dtype)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).cuda(); beta = (0.1 * torch.randn(C, generator=g)).cuda()
    scale = (0.3 * torch.randn(B, C, generator=g)).cuda(); shift = (0.3 * torch.randn(B, C, generator=g)).cuda()
    if cl:
        x = x.contiguous(memory_format=torch.channels_last); dy = dy.contiguous(memory_format=torch.channels_last)
    xr = x.float().detach().clone().requires_grad_(True)
    yr = _gn_ref(xr, gamma, beta, scale, shift, mode)
    (dxr,) = torch.autograd.grad(yr, xr, dy.float())
    y = torch.empty_like(x); dx = torch.empty_like(x)
    stats = torch.empty(B, 32, 2, device="cuda")
    sc, sh = (scale, shift) if mode == 2 else (None, None)
    kernels.gn_fwd(mode, x, y, gamma, beta, sc, sh, stats, 32, 1e-5, cl)
    kernels.gn_bwd(mode, x, dy, dx, gamma, beta, sc, sh, stats, 32, cl)
    torch.cuda.synchronize()
    xs = x.float().reshape(B, 32, -1) if not cl else x.float().permute(0, 2, 3, 1).reshape(B, H * W, 32, C // 32).permute(0, 2, 1, 3).reshape(B, 32, -1)
    np.testing.assert_allclose(stats[:, :, 0].cpu().numpy(), xs.mean(-1).cpu().numpy(), atol=2e-5)
    np.testing.assert_allclose(stats[:, :, 1].cpu().numpy(), (xs.var(-1, unbiased=False) + 1e-5).rsqrt().cpu().numpy(), rtol=2e-5)
    if dtype == torch.float32:
        np.testing.assert_allclose(y.cpu().numpy(), yr.detach().cpu().numpy(), atol=2e-5, rtol=2e-5)
        np.testing.assert_allclose(dx.cpu().numpy(), dxr.cpu().numpy(), atol=5e-5 * float(dxr.abs().max()), rtol=1e-4)
    else:   # bf16 output rounding: 2^-8 relative
        np.testing.assert_allclose(y.float().cpu().numpy(), yr.detach().cpu().numpy(), atol=2e-2, rtol=1e-2)
        np.testing.assert_allclose(dx.float().cpu().numpy(), dxr.cpu().numpy(), atol=2e-2 * float(dxr.abs().max()), rtol=2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,cols", [(64, 256), (37, 64), (8, 1024), (5, 4096), (3, 100), (9, 512), (4, 6000), (2, 8192), (3, 9000), (2, 1022)])
def test_softmax_epilogue(dtype, rows, cols):
    from nhmc_b200 import kernels
    g = torch.Generator().manual_seed(2)
    logits = (torch.randn(rows, cols, generator=g) * 3).cuda().to(dtype)
    dp = torch.randn(rows, cols, generator=g).cuda().to(dtype)
    scale = 1 / math.sqrt(64)
    lr = logits.float().detach().clone().requires_grad_(True)
    pr = torch.softmax(lr * scale, dim=-1)
    (dlr,) = torch.autograd.grad(pr, lr, dp.float())
    p = torch.empty_like(logits); dl = torch.empty_like(logits)
    kernels.softmax_fwd(logits, p, scale)
    kernels.softmax_bwd(p, dp, dl, scale)
    if dtype == torch.float32:
        np.testing.assert_allclose(p.cpu().numpy(), pr.detach().cpu().numpy(), atol=1e-6, rtol=2e-5)
        np.testing.assert_allclose(dl.cpu().numpy(), dlr.cpu().numpy(), atol=1e-6, rtol=1e-4)
    else:
        np.testing.assert_allclose(p.float().cpu().numpy(), pr.detach().cpu().numpy(), atol=4e-3, rtol=1e-2)
    assert abs(float(p.float().sum(-1).mean()) - 1.0) < 5e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,T,heads,ch", [(2, 1024, 3, 32), (1, 256, 5, 32), (3, 64, 2, 32), (2, 200, 2, 32), (1, 384, 2, 64),
                                          (2, 1024, 1, 64), (1, 72, 3, 64)])
@pytest.mark.parametrize("ramp", [False, True])
def test_fused_attention_forward(dtype, B, T, heads, ch, ramp):
    """nhmc_attn_fwd (tcgen05: S / O in tensor memory, P through shared memory) against fp32 torch attention on the
    same rounded operands (unet_ffhq.py:408-425).  T = 200 / 72 exercise the ragged last key tile and query tile; `ramp`
    makes the keys grow along the sequence so that the row maxima keep rising by far more than the lazy-rescale threshold
    (the accumulator in tensor memory is rescaled many times per row)."""
    from nhmc_b200 import kernels
    g = torch.Generator().manual_seed(7)
    C = heads * ch
    qk = torch.randn(B, T, heads, 2, ch, generator=g) * 1.5
    if ramp:
        qk[:, :, :, 1] *= torch.linspace(0.1, 8.0, T).view(1, T, 1, 1)
        qk[:, :, :, 0] = qk[:, :, :, 0].abs()                    # q . k grows with the key index for about half the channels' signs
    qk = qk.cuda().to(dtype)
    v = torch.randn(B, C, T, generator=g).cuda().to(dtype)
    # the output sits between two guard bands: a store outside [B, T, C] (ragged query tiles) would change them
    guard = 4096
    buf = torch.full((guard + B * T * C + guard,), 777.0, device="cuda", dtype=dtype)
    out = buf[guard:guard + B * T * C].view(B, T, C)
    out.fill_(float("nan"))
    scale = 1 / math.sqrt(ch)
    kernels.attn_fwd(qk.view(B, T, 2 * C), v, out, heads, scale)
    assert bool((buf[:guard] == 777.0).all()) and bool((buf[guard + B * T * C:] == 777.0).all())
    q = qk[:, :, :, 0].float().permute(0, 2, 1, 3)                       # [B, h, T, ch]
    k = qk[:, :, :, 1].float().permute(0, 2, 1, 3)
    vv = v.float().view(B, heads, ch, T)
    probs = torch.softmax(torch.einsum("bhtc,bhsc->bhts", q, k) * scale, dim=-1)
    ref = torch.einsum("bhts,bhcs->bthc", probs, vv).reshape(B, T, C)
    assert torch.isfinite(out.float()).all()
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3                      # P and the output are rounded to the 8 / 11-bit type
    np.testing.assert_allclose(out.float().cpu().numpy(), ref.cpu().numpy(), atol=tol, rtol=tol)


def test_attention_block_fused_forward_matches_autograd_path():
    """AttentionBlock under no_grad takes the fused kernel; with a gradient wanted it takes GEMM + softmax + GEMM.  Same
    weights, same input: the two agree to bf16 rounding."""
    from nhmc_b200 import unet, kernels
    torch.manual_seed(3)
    blk = unet.AttentionBlock(128, num_head_channels=32).cuda().to(torch.bfloat16)
    for p_ in blk.parameters():
        p_.data.normal_(0, 0.2)
    x = torch.randn(2, 128, 16, 16, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with kernels.gn_scope(kernels.GnWorkspace(x.device, 2)):
        calls = {}
        orig = kernels.attn_fwd
        kernels.attn_fwd = lambda *a, **k: (calls.setdefault("n", 0), calls.__setitem__("n", calls["n"] + 1), orig(*a, **k))[-1]
        try:
            with torch.no_grad():
                y_fused = blk(x)
            assert calls["n"] == 1
            y_ref = blk(x.clone().requires_grad_(True))
            assert calls["n"] == 1
        finally:
            kernels.attn_fwd = orig
    d = (y_fused.float() - y_ref.float()).abs().max().item()
    assert d <= 3e-2 * max(1.0, y_ref.float().abs().max().item()), d


def _small_model(dtype):
    from nhmc_b200 import unet
    m = unet.create_model(**SMALL, dtype=dtype)
    cfg = ounet.resolve_config(**SMALL)
    sd = ounet.seeded_state_dict(cfg, seed=1234)
    assert list(m.state_dict()) == list(sd)
    m.load_state_dict(sd)
    return m.prepare("cuda"), sd, cfg


def test_unet_fp32_matches_reference_output(golden):
    """Our U-Net on the fused kernels vs the output of the reference's own UNetModel (golden) and the oracle."""
    g = golden("unet_small.npz")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd, cfg = _small_model(torch.float32)
    x = T(g["x"]).cuda().requires_grad_(True)
    out = m(x, T(g["t"]).cuda())
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], atol=5e-5, rtol=1e-4)
    (gx,) = torch.autograd.grad(out, x, T(g["w"]).cuda())
    np.testing.assert_allclose(gx.cpu().numpy(), g["grad_x"], atol=5e-5 * max(1.0, float(np.abs(g["grad_x"]).max())), rtol=1e-3)


def test_unet_bf16_close_to_fp32(golden):
    g = golden("unet_small.npz")
    m, sd, cfg = _small_model(torch.bfloat16)
    x = T(g["x"]).cuda().requires_grad_(True)
    out = m(x, T(g["t"]).cuda())
    assert out.dtype == torch.bfloat16
    ref = g["out"]
    o = out.detach().float().cpu().numpy()
    err = np.abs(o - ref).max() / np.abs(ref).max()
    rel2 = np.linalg.norm(o - ref) / np.linalg.norm(ref)
    assert err < 0.2 and rel2 < 0.08, (err, rel2)      # bf16 (8 mantissa bits) through ~60 layers
    (gx,) = torch.autograd.grad(out, x, T(g["w"]).cuda().to(torch.bfloat16))
    gerr = np.linalg.norm(gx.float().cpu().numpy() - g["grad_x"]) / np.linalg.norm(g["grad_x"])
    assert gx.dtype == torch.float32 and gerr < 0.2, gerr


def test_hmc_with_own_unet_small(golden):
    """End-to-end: one gradient evaluation with our U-Net (fp32) vs the oracle running its functional U-Net on CPU."""
    from nhmc_b200 import _lib, operators as P, sampler, schedule
    from oracle import hmc as ohmc, operators as oops, schedule as osch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd, cfg = _small_model(torch.float32)
    D, seq, seq_next = 64, [250, 500, 750], [-1, 250, 500]
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(1, 3, D, D, generator=gen)
    oop = oops.SuperResolution(3, D, 4)
    y = oop.H(torch.rand(1, 3, D, D, generator=gen) * 2 - 1) + 0.1 * torch.randn(1, 3 * 16 * 16, generator=gen)
    betas = T(osch.beta_schedule("linear", 1e-4, 0.02, 1000)).float()
    steps = schedule.ddim_steps(schedule.alpha_table(betas), seq, seq_next)
    op = P.SuperResolution(3, D, 4, "cuda")
    ev = sampler.GradEvaluator(m, steps, op, y.cuda(), x.cuda())
    loss, ga, gb, xc = ev(x.cuda())
    omodel = lambda xx, tt: ounet.unet_forward(sd, cfg, xx, tt)
    tab = osch.alpha_bar_table(osch.beta_schedule("linear", 1e-4, 0.02, 1000))
    xh, oloss, og = ohmc.grad_eval(x, ohmc.module_eps_net(omodel), tab, seq, seq_next, oop, y)
    np.testing.assert_allclose(xc.cpu().numpy(), xh.numpy(), atol=1e-3)                  # north-star tolerance
    np.testing.assert_allclose(loss.cpu().numpy(), oloss.numpy(), rtol=1e-3)
    # The x0 clip of each DDIM step is a hard threshold behind a 1/sqrt(abar_750) = 17x gain: a last-bit change in the
    # GroupNorm reduction order can flip ONE borderline pixel's clip decision, which changes the gradient inside that
    # pixel's receptive field (~18x18 values) by a few percent while everything else stays within tolerance.
    diff = ((ga + gb).cpu() - og).abs()
    gmax = float(og.abs().max())
    assert float((diff > 2e-3 * gmax).float().mean()) < 0.05, float((diff > 2e-3 * gmax).float().mean())
    assert float(diff.max()) < 0.05 * gmax, (float(diff.max()), gmax)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_resample_and_split_kernels(dtype):
    from nhmc_b200 import unet
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 64, 12, 20, generator=g).cuda().to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = unet._AvgPool2.apply(x)
    yr = F.avg_pool2d(x.detach().float(), 2)
    np.testing.assert_allclose(y.detach().float().cpu().numpy(), yr.cpu().numpy(), atol=1e-6 if dtype == torch.float32 else 2e-2)
    w = torch.randn(y.shape, generator=g).cuda().to(dtype)
    (gx,) = torch.autograd.grad(y, x, w)
    gr = F.interpolate(w.float(), scale_factor=2, mode="nearest") / 4
    np.testing.assert_allclose(gx.float().cpu().numpy(), gr.cpu().numpy(), atol=1e-6 if dtype == torch.float32 else 1e-2)
    u = unet._Upsample2.apply(x)
    assert torch.equal(u.detach(), F.interpolate(x.detach(), scale_factor=2, mode="nearest"))
    w2 = torch.randn(u.shape, generator=g).cuda().to(dtype)
    (gu,) = torch.autograd.grad(u, x, w2)
    gur = F.avg_pool2d(w2.float(), 2) * 4
    np.testing.assert_allclose(gu.float().cpu().numpy(), gur.cpu().numpy(), atol=1e-5 if dtype == torch.float32 else 5e-2)
    a = torch.randn(2, 32, 6, 10, generator=g).cuda().to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    b = torch.randn(2, 96, 6, 10, generator=g).cuda().to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    c = unet._Cat2.apply(a, b)
    assert torch.equal(c.detach(), torch.cat([a.detach(), b.detach()], 1))
    w3 = torch.randn(c.shape, generator=g).cuda().to(dtype)
    ga, gb = torch.autograd.grad(c, (a, b), w3)
    assert torch.equal(ga, w3[:, :32]) and torch.equal(gb, w3[:, 32:])
    assert ga.is_contiguous(memory_format=torch.channels_last) and gb.is_contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("resident", [5, 4, 3, 2, 1, 0])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("shape", [(8, 128, 128, 128), (3, 384, 64, 64), (2, 256, 256, 256), (5, 1024, 16, 16), (2, 128, 33, 7),
                                   (8, 512, 8, 8), (2, 224, 32, 32), (8, 768, 16, 16), (8, 256, 64, 64)])
def test_group_norm_channels_last_paths(resident, dtype, mode, shape):
    """Channels-last GroupNorm on every execution path (5 = + cluster-per-unit kernel for the 64x64 maps; 4 = one-CTA-per-unit kernel for the small maps first;
    3..1 = shared-memory resident kernels; 0 = statistics + apply kernels): samples spanning many CTAs, a sample filling
    the whole chip, ragged pixel counts, 24-channel groups, the folded conv bias and the fused skip-gradient add;
    repeated calls check that the self-resetting counters reset."""
    from nhmc_b200 import kernels, _lib
    B, C, H, W = shape
    if dtype == torch.float32 and B * C * H * W > 2 ** 25:
        pytest.skip("fp32 copy of the largest shape is covered in bf16")
    g = torch.Generator().manual_seed(3)
    cl = torch.channels_last
    x = (torch.randn(shape, generator=g) * 1.7 + 0.4).cuda().to(dtype).contiguous(memory_format=cl)
    dy = torch.randn(shape, generator=g).cuda().to(dtype).contiguous(memory_format=cl)
    add = torch.randn(shape, generator=g).cuda().to(dtype).contiguous(memory_format=cl)
    bias = (0.5 * torch.randn(C, generator=g)).cuda()
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).cuda(); beta = (0.1 * torch.randn(C, generator=g)).cuda()
    scale = (0.3 * torch.randn(B, C, generator=g)).cuda(); shift = (0.3 * torch.randn(B, C, generator=g)).cuda()
    sc, sh = (scale, shift) if mode == 2 else (None, None)
    xr = x.float().detach().clone().requires_grad_(True)
    yr = _gn_ref(xr + bias[None, :, None, None], gamma, beta, scale, shift, mode)
    (dxr,) = torch.autograd.grad(yr, xr, dy.float())
    dxr = dxr + add.float()
    # the per-call flag word replaces the former process-wide setters (include/nhmc.h NHMC_GN_F_*)
    res = min(resident, 3)
    flags = ((_lib.GN_F_NO_RESIDENT if res == 0 else 0) | (_lib.GN_F_NO_CLUSTER if res < 2 else 0) | (_lib.GN_F_SLICE_ALL if res >= 3 else 0)
             | (_lib.GN_F_NO_SOLO if resident < 4 else 0) | (_lib.GN_F_NO_SOLO_CLUSTER if resident < 5 else 0))
    prev = kernels.set_flags(flags)
    try:
        for rep in range(3):
            y = torch.empty_like(x); dx = torch.empty_like(x); stats = torch.empty(B, 32, 2, device="cuda")
            kernels.gn_fwd(mode, x, y, gamma, beta, sc, sh, stats, 32, 1e-5, True, bias)
            kernels.gn_bwd(mode, x, dy, dx, gamma, beta, sc, sh, stats, 32, True, bias, dx_add=add)
            kernels.gn_check()
            if dtype == torch.float32:
                tol = dict(atol=5e-5, rtol=1e-4)
            else:
                tol = dict(atol=3e-2, rtol=2e-2)
            yd = (y.float() - yr.detach()).abs()
            assert float((yd - tol["atol"] - tol["rtol"] * yr.detach().abs()).max()) <= 0, (rep, float(yd.max()))
            dd = (dx.float() - dxr).abs()
            assert float((dd - tol["atol"] * float(dxr.abs().max()) - tol["rtol"] * dxr.abs()).max()) <= 0, (rep, float(dd.max()))
    finally:
        kernels.set_flags(prev)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cl", [False, True])
def test_group_norm_with_folded_conv_bias(dtype, cl):
    """GN(x + bias[c]) computed from the raw conv output x: forward, statistics and input gradient."""
    from nhmc_b200 import kernels
    B, C, H, W = 2, 128, 24, 24
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(B, C, H, W, generator=g) * 1.3).cuda().to(dtype)
    dy = torch.randn(B, C, H, W, generator=g).cuda().to(dtype)
    bias = (0.8 * torch.randn(C, generator=g)).cuda()
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).cuda(); beta = (0.1 * torch.randn(C, generator=g)).cuda()
    scale = (0.3 * torch.randn(B, C, generator=g)).cuda(); shift = (0.3 * torch.randn(B, C, generator=g)).cuda()
    if cl:
        x = x.contiguous(memory_format=torch.channels_last); dy = dy.contiguous(memory_format=torch.channels_last)
    xr = x.float().detach().clone().requires_grad_(True)
    yr = _gn_ref(xr + bias[None, :, None, None], gamma, beta, scale, shift, 2)
    (dxr,) = torch.autograd.grad(yr, xr, dy.float())
    y = torch.empty_like(x); dx = torch.empty_like(x); stats = torch.empty(B, 32, 2, device="cuda")
    kernels.gn_fwd(2, x, y, gamma, beta, scale, shift, stats, 32, 1e-5, cl, bias)
    kernels.gn_bwd(2, x, dy, dx, gamma, beta, scale, shift, stats, 32, cl, bias)
    tol = dict(atol=5e-5, rtol=1e-4) if dtype == torch.float32 else dict(atol=3e-2, rtol=2e-2)
    np.testing.assert_allclose(y.float().cpu().numpy(), yr.detach().cpu().numpy(), **tol)
    np.testing.assert_allclose(dx.float().cpu().numpy(), dxr.cpu().numpy(), atol=tol["atol"] * float(dxr.abs().max()), rtol=tol["rtol"])


def test_bias_add_kernel():
    from nhmc_b200 import unet
    g = torch.Generator().manual_seed(9)
    a = torch.randn(2, 64, 10, 12, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    b = torch.randn(2, 64, 10, 12, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bias = torch.randn(64, generator=g).cuda()
    out = unet._BiasAdd.apply(a, b, bias)
    ref = (a.float() + b.float() + bias[None, :, None, None]).to(torch.bfloat16)
    assert torch.equal(out.detach(), ref)
    w = torch.randn_like(out)
    ga, gb = torch.autograd.grad(out, (a, b), w)
    assert torch.equal(ga, w) and torch.equal(gb, w)
    out1 = unet._BiasAdd.apply(a, None, bias)
    assert torch.equal(out1.detach(), (a.float() + bias[None, :, None, None]).to(torch.bfloat16))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 128, 128, 64, 64), (3, 256, 128, 32, 40), (8, 512, 256, 16, 16), (1, 64, 96, 9, 5)])
def test_group_norm_over_virtual_concat(dtype, shape):
    """GroupNorm+SiLU over cat([x0, x1], 1) without the concatenation: forward, both input gradients, fused adds."""
    from nhmc_b200 import unet, _lib
    B, c0, c1, H, W = shape
    g = torch.Generator().manual_seed(11)
    cl = torch.channels_last
    mk = lambda c: (torch.randn(B, c, H, W, generator=g) * 1.5 + 0.3).cuda().to(dtype).contiguous(memory_format=cl)
    x0, x1 = mk(c0).requires_grad_(True), mk(c1).requires_grad_(True)
    C = c0 + c1
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).cuda(); beta = (0.1 * torch.randn(C, generator=g)).cuda()
    y, a0, a1 = unet._CatGroupNormAct.apply(x0, x1, gamma, beta, _lib.GN_SILU)
    w = torch.randn(B, C, H, W, generator=g).cuda().to(dtype); w0 = torch.randn(B, c0, H, W, generator=g).cuda().to(dtype)
    w1 = torch.randn(B, c1, H, W, generator=g).cuda().to(dtype)
    g0, g1 = torch.autograd.grad([y, a0, a1], (x0, x1), [w, w0, w1])
    xr0, xr1 = x0.detach().float().requires_grad_(True), x1.detach().float().requires_grad_(True)
    yr = _gn_ref(torch.cat([xr0, xr1], 1), gamma, beta, None, None, 1)
    r0, r1 = torch.autograd.grad(yr, (xr0, xr1), w.float())
    r0, r1 = r0 + w0.float(), r1 + w1.float()
    tol = dict(atol=5e-5, rtol=1e-4) if dtype == torch.float32 else dict(atol=3e-2, rtol=2e-2)
    np.testing.assert_allclose(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy(), **tol)
    for got, ref in ((g0, r0), (g1, r1)):
        np.testing.assert_allclose(got.float().cpu().numpy(), ref.cpu().numpy(), atol=tol["atol"] * float(ref.abs().max()), rtol=tol["rtol"])
    assert torch.equal(a0.detach(), x0.detach()) and torch.equal(a1.detach(), x1.detach())


def test_bias_add_three_inputs():
    from nhmc_b200 import unet
    g = torch.Generator().manual_seed(10)
    mk = lambda: torch.randn(2, 64, 6, 10, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    a, b, c = mk(), mk(), mk()
    bias = torch.randn(64, generator=g).cuda()
    out = unet._BiasAdd.apply(a, b, bias, c)
    ref = (a.float() + (b.float() + c.float()) + bias[None, :, None, None]).to(torch.bfloat16)
    assert torch.equal(out.detach(), ref)
    w = torch.randn_like(out)
    ga, gb, gc = torch.autograd.grad(out, (a, b, c), w)
    assert torch.equal(ga, w) and torch.equal(gb, w) and torch.equal(gc, w)


def test_unet_virtual_concat_matches_materialised_concat(golden):
    """The same network with the skip concatenations kept virtual (default at >= 64x64) and fully materialised."""
    g = golden("unet_small.npz")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd, cfg = _small_model(torch.float32)
    outs = []
    for min_pix in (1, 10 ** 9):
        m.virtual_cat_min_pixels = min_pix
        x = T(g["x"]).cuda().requires_grad_(True)
        out = m(x, T(g["t"]).cuda())
        (gx,) = torch.autograd.grad(out, x, T(g["w"]).cuda())
        outs.append((out.detach(), gx))
        np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], atol=5e-5, rtol=1e-4)
        np.testing.assert_allclose(gx.cpu().numpy(), g["grad_x"], atol=5e-5 * max(1.0, float(np.abs(g["grad_x"]).max())), rtol=1e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("shape", [(8, 128, 128, 128), (2, 128, 33, 7), (3, 256, 16, 16)])
def test_group_norm_backward_two_addends(dtype, mode, shape):
    """dx = gn_bwd(dy) + dx_add + dx_add2 in one pass (ResBlock skip-branch gradient + long-skip gradient)."""
    from nhmc_b200 import kernels
    B, C, H, W = shape
    g = torch.Generator().manual_seed(5)
    cl = torch.channels_last
    x = (torch.randn(shape, generator=g) * 1.3 - 0.2).cuda().to(dtype).contiguous(memory_format=cl)
    dy, a1, a2 = (torch.randn(shape, generator=g).cuda().to(dtype).contiguous(memory_format=cl) for _ in range(3))
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).cuda(); beta = (0.1 * torch.randn(C, generator=g)).cuda()
    scale = (0.3 * torch.randn(B, C, generator=g)).cuda(); shift = (0.3 * torch.randn(B, C, generator=g)).cuda()
    sc, sh = (scale, shift) if mode == 2 else (None, None)
    xr = x.float().detach().clone().requires_grad_(True)
    yr = _gn_ref(xr, gamma, beta, scale, shift, mode)
    (dxr,) = torch.autograd.grad(yr, xr, dy.float())
    dxr = dxr + a1.float() + a2.float()
    for rep in range(2):
        y = torch.empty_like(x); dx = torch.empty_like(x); stats = torch.empty(B, 32, 2, device="cuda")
        kernels.gn_fwd(mode, x, y, gamma, beta, sc, sh, stats, 32, 1e-5, True)
        kernels.gn_bwd(mode, x, dy, dx, gamma, beta, sc, sh, stats, 32, True, dx_add=a1, dx_add2=a2)
        kernels.gn_check()
        tol = dict(atol=5e-5, rtol=1e-4) if dtype == torch.float32 else dict(atol=3e-2, rtol=2e-2)
        dd = (dx.float() - dxr).abs()
        assert float((dd - tol["atol"] * float(dxr.abs().max()) - tol["rtol"] * dxr.abs()).max()) <= 0, (rep, float(dd.max()))


def test_unet_long_skip_gradient_fused_into_group_norm_backward(golden):
    """fork_min_bytes = 0: every input block hands the skip list a pass-through output of the next block's first
    GroupNorm, whose backward kernel then adds the long-skip gradient (dx_add2); None: autograd's own accumulation."""
    g = golden("unet_small.npz")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd, cfg = _small_model(torch.float32)
    for fork in (0, None):
        m.fork_min_bytes = fork
        x = T(g["x"]).cuda().requires_grad_(True)
        out = m(x, T(g["t"]).cuda())
        (gx,) = torch.autograd.grad(out, x, T(g["w"]).cuda())
        np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], atol=5e-5, rtol=1e-4)
        np.testing.assert_allclose(gx.cpu().numpy(), g["grad_x"], atol=5e-5 * max(1.0, float(np.abs(g["grad_x"]).max())), rtol=1e-3)


def test_programmatic_dependent_launch_does_not_change_results():
    """The epilogue kernels launched with and without the programmatic-stream-serialization attribute (NHMC_F_NO_PDL):
    bit-identical GroupNorm / residual-add / pooling results, eagerly and replayed from a CUDA graph."""
    from nhmc_b200 import kernels, _lib, unet
    g = torch.Generator().manual_seed(17)
    cl = torch.channels_last
    shape = (4, 256, 32, 32)
    x = torch.randn(shape, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=cl)
    dy = torch.randn(shape, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=cl)
    big = torch.randn(8, 128, 128, 128, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=cl)
    gamma = (1 + 0.2 * torch.randn(256, generator=g)).cuda(); beta = (0.1 * torch.randn(256, generator=g)).cuda()
    bias = torch.randn(256, generator=g).cuda()

    def chain():
        s = torch.empty_like(x); y = torch.empty_like(x); dx = torch.empty_like(x); st = torch.empty(4, 32, 2, device="cuda")
        kernels.bias_add(x, dy, bias, s)                                       # residual add -> GroupNorm -> pooling
        kernels.gn_fwd(1, s, y, gamma, beta, None, None, st, 32, 1e-5, True)
        p = torch.empty((4, 256, 16, 16), dtype=x.dtype, device="cuda", memory_format=cl)
        kernels.pool2_gather(y, p, 0.25)
        kernels.gn_bwd(1, s, dy, dx, gamma, beta, None, None, st, 32, True, dx_add=y)
        yb = torch.empty_like(big); stb = torch.empty(8, 32, 2, device="cuda")   # statistics + apply pair
        kernels.gn_fwd(1, big, yb, gamma[:128].contiguous(), beta[:128].contiguous(), None, None, stb, 32, 1e-5, True)
        return s, y, p, dx, yb

    outs = {}
    scope = kernels.gn_scope(kernels.GnWorkspace(x.device, 8))    # one caller-owned GroupNorm workspace for eager calls and captures alike
    scope.__enter__()
    for pdl in (0, 1):
        prev = kernels.set_flags(0 if pdl else _lib.F_NO_PDL)
        try:
            outs[(pdl, "eager")] = [t.clone() for t in chain()]
            chain(); torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                res = chain()
            gr.replay(); gr.replay(); torch.cuda.synchronize()
            outs[(pdl, "graph")] = [t.clone() for t in res]
        finally:
            kernels.set_flags(prev)
    kernels.gn_check()
    scope.__exit__(None, None, None)
    ref = outs[(0, "eager")]
    for key, val in outs.items():
        for a, b in zip(ref, val):
            assert torch.equal(a, b), key


def test_two_samplers_on_two_host_threads_with_different_lane_counts(golden):
    """the library keeps no scratch and no switches of its own, so two HMCSamplers (1 lane and 2
    lanes) stepping concurrently from two host threads on their own streams -- each with its own GroupNorm workspaces,
    passed per call -- give what each gives when run alone."""
    import threading
    from nhmc_b200 import _lib, kernels, operators as P, sampler, schedule
    from oracle import operators as oops, schedule as osch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd, cfg = _small_model(torch.float32)
    D = 64
    betas = T(osch.beta_schedule("linear", 1e-4, 0.02, 1000)).float()
    steps = schedule.ddim_steps(schedule.alpha_table(betas), [250, 500, 750], [-1, 250, 500])
    gen = torch.Generator().manual_seed(11)
    oop = oops.SuperResolution(3, D, 4)

    def make(B, lanes):
        x = torch.randn(B, 3, D, D, generator=gen)
        y = oop.H(torch.rand(B, 3, D, D, generator=gen) * 2 - 1) + 0.1 * torch.randn(B, 3 * 16 * 16, generator=gen)
        p = torch.randn(B, 3, D, D, generator=gen)
        return dict(B=B, lanes=lanes, x=x, y=y, p=p)

    jobs = [make(4, 1), make(4, 2)]

    def run(job, out, key):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                op = P.SuperResolution(3, D, 4, "cuda")
                ev = sampler.GradEvaluator(m, steps, op, job["y"].cuda(), job["x"].cuda())
                smp = sampler.HMCSampler(job["x"].cuda(), ev, 1.0, 0.05, 0.1, sched=_lib.SCHED_PIXEL, lanes=job["lanes"])
                assert smp.n_lanes == job["lanes"] and len(smp.gn_workspaces()) == 1 + (job["lanes"] if job["lanes"] > 1 else 0)
                smp.iteration(job["p"].cuda(), torch.full((job["B"],), 0.5))
                smp.poll_gn()
                stream.synchronize()
                smp.check_gn()
                out[key] = (smp.x_prop.cpu().clone(), smp.chains.read()["H"].copy())
        except Exception as e:      # noqa: BLE001
            out[key] = e

    det = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        alone, together = {}, {}
        for i, job in enumerate(jobs):
            run(job, alone, i)
        threads = [threading.Thread(target=run, args=(job, together, i)) for i, job in enumerate(jobs)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=600)
    finally:
        torch.backends.cudnn.deterministic = det
    for i in range(2):
        for res in (alone[i], together[i]):
            if isinstance(res, Exception):
                raise res
        d = (alone[i][0] - together[i][0]).abs()
        # same arithmetic; what may differ is a borderline clip decision flipped by library-kernel nondeterminism next to
        # a concurrent stream (see test_lanes_groups_on_streams_equal_groups_run_alone)
        assert float((d > 1e-4).float().mean()) <= 0.02 and float(d.max()) <= 0.25, (i, float(d.max()))
        np.testing.assert_allclose(together[i][1], alone[i][1], rtol=2e-3)
