"""Tiny stand-in eps-network for fast parity runs (oracle; test infrastructure).

Same call signature as the reference's score net, `model(x:[B,3,H,W], t:[B] float) -> [B,6,H,W]`
(/root/reference/guided_diffusion/unet_ffhq.py:705), a few thousand parameters, smooth and
t-dependent so the DDIM clips and the VJP are exercised.  Weights come from a seeded
generator so that the golden generator, the CPU tests and the GPU tests build the same net.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class TinyEpsNet(nn.Module):
    def __init__(self, width=8, seed=4321, out_channels=6):
        super().__init__()
        self.c1 = nn.Conv2d(3, width, 3, padding=1)
        self.c2 = nn.Conv2d(width, width, 3, padding=1)
        self.c3 = nn.Conv2d(width, out_channels, 3, padding=1)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.25 if p.ndim > 1 else 0.05))
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, x, t):
        s = (t.float() / 1000.0).view(-1, 1, 1, 1)
        h = torch.tanh(self.c1(x)) * (1.0 + s)
        h = F.silu(self.c2(h)) + 0.5 * s
        return self.c3(h) + 0.3 * x.repeat(1, self.c3.out_channels // 3, 1, 1)


class TinyLDM(nn.Module):
    """Stand-in for the LatentDiffusion object used by hmc_latent (main_sampling_latent.py:623-789).

    Exposes exactly what the sampler touches: `apply_model(z, t, cond)` under no_grad (ddpm.py:892-893),
    `differentiable_decode_first_stage(z)` / `decode_first_stage(z)` (a small f=4 decoder; the VQ lookup of the
    real VQ-f4 autoencoder is the un-pinned third-party boundary and is NOT part of this stand-in), and the
    schedule buffers `alphas_cumprod`, `alphas_cumprod_prev` (LDM sqrt-linear schedule)."""

    def __init__(self, seed=9876):
        super().__init__()
        from .schedule import ldm_alphas_cumprod
        self.eps = TinyEpsNet(width=8, seed=seed, out_channels=3)
        self.d1 = nn.Conv2d(3, 12, 3, padding=1)
        self.d2 = nn.Conv2d(12, 3 * 16, 3, padding=1)
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for p in list(self.d1.parameters()) + list(self.d2.parameters()):
                p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.ndim > 1 else 0.05))
        for p in self.parameters():
            p.requires_grad_(False)
        ac, acp = ldm_alphas_cumprod()
        self.register_buffer("alphas_cumprod", ac)
        self.register_buffer("alphas_cumprod_prev", acp)

    @torch.no_grad()
    def apply_model(self, z, t, cond=None):
        return self.eps(z, t)

    def differentiable_decode_first_stage(self, z):
        h = torch.tanh(self.d1(z))
        return F.pixel_shuffle(self.d2(h), 4) * 0.7

    @torch.no_grad()
    def decode_first_stage(self, z):
        return self.differentiable_decode_first_stage(z)
