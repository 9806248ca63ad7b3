"""Latent model family: our latent U-Net and VQ-f4 first stage on the fused kernels against outputs
recorded from the reference's own ldm modules (tests/golden/make_golden.py gold_latent_models)."""
import types

import numpy as np
import pytest
import torch

from oracle.latent_fill import SMALL_FIRST_STAGE, SMALL_UNET, seeded_fill

T = torch.from_numpy


def _small_latent(dtype):
    from nhmc_b200 import latent
    m = latent.LatentDiffusion(unet_config=SMALL_UNET, first_stage_config=SMALL_FIRST_STAGE, dtype=dtype)
    un, fs = m.model.diffusion_model, m.first_stage_model
    un.load_state_dict(seeded_fill({k: v.shape for k, v in un.state_dict().items()}, seed=4321))
    fs.load_state_dict(seeded_fill({k: v.shape for k, v in fs.state_dict().items()}, seed=8765))
    return m


def test_state_dict_keys_match_the_reference_modules(golden):
    """Same parameter names and order as ldm's UNetModel / VQModelInterface(decoder, quantize, post_quant_conv): reference
    checkpoints load by key, and the seeded fill gives both sides identical weights."""
    from nhmc_b200 import latent
    g = golden("latent_models.npz")
    m = latent.LatentDiffusion(unet_config=SMALL_UNET, first_stage_config=SMALL_FIRST_STAGE, dtype=torch.float32)
    assert list(m.model.diffusion_model.state_dict()) == [str(k) for k in g["unet_keys"]]
    assert list(m.first_stage_model.state_dict()) == [str(k) for k in g["dec_keys"]]
    keys = list(m.state_dict())
    assert "model.diffusion_model.input_blocks.2.0.op.weight" in keys and "first_stage_model.decoder.up.1.upsample.conv.weight" in keys
    assert "first_stage_model.quantize.embedding.weight" in keys and "alphas_cumprod" in keys
    ac, acp = latent.ldm_alphas_cumprod()
    np.testing.assert_allclose([float(ac[250]), float(ac[500]), float(ac[750])], [0.5084, 0.1149, 0.0083], atol=2e-4)
    full = latent.create_latent_model(dtype=torch.float32)              # ffhq_latent: 274 M U-Net + VQ-f4 decode half
    n_unet = sum(p.numel() for p in full.model.diffusion_model.parameters())
    assert abs(n_unet - 274.06e6) < 0.05e6, n_unet


@pytest.mark.gpu
def test_latent_unet_matches_reference_output(golden):
    g = golden("latent_models.npz")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = _small_latent(torch.float32).prepare("cuda")
    out = m.apply_model(T(g["unet_x"]).cuda(), T(g["unet_t"]).cuda(), None)
    assert not out.requires_grad                                             # ddpm.py:892-893: no-grad
    np.testing.assert_allclose(out.float().cpu().numpy(), g["unet_out"], atol=1e-4, rtol=1e-4)
    mb = _small_latent(torch.bfloat16).prepare("cuda")
    ob = mb.apply_model(T(g["unet_x"]).cuda(), T(g["unet_t"]).cuda(), None).float().cpu().numpy()
    assert np.linalg.norm(ob - g["unet_out"]) / np.linalg.norm(g["unet_out"]) < 0.05


@pytest.mark.gpu
def test_first_stage_decode_matches_reference(golden):
    """quantize (nearest code, straight-through) -> post_quant_conv -> decoder: value, code indices and d/dz."""
    from nhmc_b200 import kernels
    g = golden("latent_models.npz")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = _small_latent(torch.float32).prepare("cuda")
    z = T(g["dec_z"]).cuda()
    E = m.first_stage_model.quantize.embedding.weight.detach()
    zq = torch.empty_like(z); idx = torch.empty(2, 64, dtype=torch.int32, device="cuda")
    kernels.vq_nearest(z.contiguous(), E.contiguous(), zq, idx)
    assert float(g["dec_margin"]) > 1e-5                                     # no near-tie in the recorded lookup
    assert np.array_equal(idx.cpu().numpy(), g["dec_idx"].astype(np.int32))
    assert np.array_equal(zq.cpu().numpy(), g["dec_zq"])
    zv = z.clone().requires_grad_(True)
    out = m.differentiable_decode_first_stage(zv)
    np.testing.assert_allclose(out.detach().float().cpu().numpy(), g["dec_out"], atol=2e-4, rtol=1e-4)
    (gz,) = torch.autograd.grad(out, zv, T(g["dec_w"]).cuda())
    np.testing.assert_allclose(gz.cpu().numpy(), g["dec_grad_z"], atol=2e-4 * float(np.abs(g["dec_grad_z"]).max()), rtol=1e-3)
    with torch.no_grad():
        o2 = m.decode_first_stage(z)
    assert torch.equal(o2, out.detach())
    # force_not_quantize skips the lookup (ddpm.py:766-823)
    o3 = m.differentiable_decode_first_stage(z, force_not_quantize=True)
    assert not torch.equal(o3, out.detach())


@pytest.mark.gpu
def test_vq_lookup_ragged_and_large_codebook():
    from nhmc_b200 import kernels
    g = torch.Generator().manual_seed(3)
    for (B, D, H, W, K) in [(1, 3, 5, 7, 1500), (2, 4, 16, 16, 8192), (3, 3, 64, 64, 8192)]:
        z = torch.randn(B, D, H, W, generator=g).cuda()
        E = (torch.randn(K, D, generator=g) * 0.7).cuda()
        zq = torch.empty_like(z); idx = torch.empty(B, H * W, dtype=torch.int32, device="cuda")
        kernels.vq_nearest(z, E, zq, idx)
        zf = z.permute(0, 2, 3, 1).reshape(-1, D).double()
        d = (zf ** 2).sum(1, keepdim=True) + (E.double() ** 2).sum(1) - 2 * zf @ E.double().t()
        ref = d.argmin(1).view(B, H * W)
        same = (idx.long() == ref)
        # fp32 distances may flip exact near-ties: the chosen code must then be as close as the fp64 winner
        if not bool(same.all()):
            dd = d.view(B, H * W, K)
            gap = dd.gather(2, idx.long()[..., None]) - dd.gather(2, ref[..., None])
            assert float(gap.abs().max()) < 1e-5
        assert float(same.float().mean()) > 0.999
        ev = E[idx.long().view(-1)]; zf32 = z.permute(0, 2, 3, 1).reshape(-1, D)
        assert torch.equal(zq.permute(0, 2, 3, 1).reshape(-1, D), zf32 + (ev - zf32))        # z + (e - z), as the reference forms it


@pytest.mark.gpu
def test_hmc_latent_runs_on_our_latent_model():
    """hmc_latent drop-in call with our LatentDiffusion object (small config): finite samples, schedule advances."""
    import nhmc_b200
    from nhmc_b200 import operators as P
    m = _small_latent(torch.float32).prepare("cuda")
    g = torch.Generator().manual_seed(5)
    D = 32                                             # decoded image size for a 16x16 latent at f = 2 (small config)
    op = P.SuperResolution(3, D, 4, "cuda")
    x_true = torch.rand(1, 3, D, D, generator=g) * 2 - 1
    y = op.H(x_true.cuda()) + 0.1 * torch.randn(1, 3 * 8 * 8, generator=g).cuda()
    z0 = torch.randn(1, 3, 16, 16, generator=g).cuda()
    opt = types.SimpleNamespace(tau=0.2, epsilon=0.05, m=1.0, sigma_0=0.1, sigma_y=0.5)
    smp = nhmc_b200.hmc_latent(z0, 1, [250, 500, 750], [-1, 250, 500], types.SimpleNamespace(model=m, cls_fn=None), opt, y, op, None,
                               max_iters=3, return_sampler=True, image_size=D, cuda_graphs=False)   # eager: the evaluation counters below count launches
    st = smp.chains.read()
    assert int(st["iters"][0]) == 3 and np.isfinite(st["H"]).all() and bool(torch.isfinite(smp.x).all())
    assert smp.ev.n_net_fwd == 3 * (4 + 1) * 3 and smp.ev.n_net_vjp == 0        # L = tau/eps = 4; eps carries no gradient
