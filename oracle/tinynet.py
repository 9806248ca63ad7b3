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
