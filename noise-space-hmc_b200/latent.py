"""Latent-diffusion model objects for the `hmc_latent` path on the B200 kernels.

What `hmc_latent` touches of the reference's `LatentDiffusion` object (main_sampling_latent.py:623-789;
algos/unconditional_latent.py:4-28) is four members: `apply_model(z, t, None)` (no-grad, ldm/models/diffusion/
ddpm.py:892-990), `differentiable_decode_first_stage(z)` (:766-823), `decode_first_stage(z)` (:705-763) and the
schedule buffers `alphas_cumprod`, `alphas_cumprod_prev` (:117-170).  This module provides them with the same
state_dict keys as the reference's `ffhq_latent` model (configs/config_ffhq_latent.yml), so `model.ckpt` loads with
`load_state_dict(strict=False)` (encoder / EMA / loss weights are not needed for sampling):

  * `model.diffusion_model.*`  -- the latent U-Net (ldm/modules/diffusionmodules/openaimodel.py:413-742): our
    `unet.UNetModel` in its conv-resampling, additive-embedding variant (resblock_updown=False,
    use_scale_shift_norm=False, num_head_channels=32);
  * `first_stage_model.{quantize, post_quant_conv, decoder}.*` -- VQ-f4 (ldm/models/autoencoder.py:274-282,
    ldm/modules/diffusionmodules/model.py:440-568): nearest-code lookup (nhmc_vq_nearest, straight-through
    gradient), 1x1 conv, and the decoder whose GroupNorm(eps 1e-6)+swish and T=4096 attention softmax run in the
    fused sm_100a epilogue kernels; convolutions stay PyTorch (cuDNN) in bf16 channels-last.

CUDA-only.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, kernels
from .unet import (UNetModel, _derived, _AttentionCore, _BiasAdd, _Norm, _ScaledSoftmax, _Upsample2, _cl, _conv_bias, _conv_nobias, group_norm_act,
                   group_norm_act_skip)

VQ_GN_EPS = 1e-6          # model.py:37-38 Normalize()


def ldm_alphas_cumprod(linear_start=0.0015, linear_end=0.0195, n_timestep=1000):
    """LDM 'linear' schedule (ldm/modules/diffusionmodules/util.py:21-25: linspace(sqrt(start), sqrt(end))**2 in
    float64) -> (alphas_cumprod, alphas_cumprod_prev) float32 buffers as register_schedule builds them (ddpm.py:117-170)."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas, axis=0)
    acp = np.append(1.0, ac[:-1])
    return torch.tensor(ac, dtype=torch.float32), torch.tensor(acp, dtype=torch.float32)


class _VQLookup(torch.autograd.Function):
    """z_q = z + (e[argmin] - z).detach(): value of the nearest code, identity gradient (taming VectorQuantizer2)."""

    @staticmethod
    def forward(ctx, z, codebook):
        z32 = z.detach().float().contiguous()
        zq = torch.empty_like(z32)
        kernels.vq_nearest(z32, codebook, zq)
        return zq.to(z.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


class VectorQuantizer(nn.Module):
    """Inference half of taming.modules.vqvae.quantize.VectorQuantizer2 (pinned taming-transformers==0.0.1, not
    vendored upstream): argmin_k |z - e_k|^2 over the `embedding` table, straight-through gradient."""

    def __init__(self, n_embed, embed_dim):
        super().__init__()
        self.embedding = nn.Embedding(n_embed, embed_dim)

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self.embedding.weight.data = self.embedding.weight.data.float()      # distances are taken in fp32
        return self

    def forward(self, z):
        return _VQLookup.apply(z, self.embedding.weight.detach().contiguous())


class ResnetBlock(nn.Module):
    """model.py:82-141 without time embedding (temb_channels = 0 in the first stage)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = _Norm(in_channels, VQ_GN_EPS)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = _Norm(out_channels, VQ_GN_EPS)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1)

    def _biases(self):
        b1, b2 = self.conv1.bias, self.conv2.bias
        has_nin = self.in_channels != self.out_channels

        def build():
            b_out = b2.detach().float()
            if has_nin:
                b_out = b_out + self.nin_shortcut.bias.detach().float()
            return b1.detach().float().contiguous(), b_out.contiguous()
        return _derived(self, "_bias_cache", (b1, b2) + ((self.nin_shortcut.bias,) if has_nin else ()), build)

    def forward(self, x):
        b1, b_out = self._biases()
        h, x = group_norm_act_skip(x, self.norm1, _lib.GN_SILU)
        h = _conv_nobias(self.conv1, h)
        h = group_norm_act(h, self.norm2, _lib.GN_SILU, in_bias=b1)          # conv1's bias folded into the norm
        h = _conv_nobias(self.conv2, h)
        skip = x if self.in_channels == self.out_channels else _conv_nobias(self.nin_shortcut, x)
        return _BiasAdd.apply(skip, h, b_out)


class AttnBlock(nn.Module):
    """Single-head spatial self-attention (model.py:150-202): softmax(q^T k / sqrt(C)) over T = H*W tokens."""

    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.norm = _Norm(channels, VQ_GN_EPS)
        self.q = nn.Conv2d(channels, channels, 1)
        self.k = nn.Conv2d(channels, channels, 1)
        self.v = nn.Conv2d(channels, channels, 1)
        self.proj_out = nn.Conv2d(channels, channels, 1)

    def _qkv(self):
        ws = tuple(m.weight for m in (self.q, self.k, self.v)) + tuple(m.bias for m in (self.q, self.k, self.v))
        return _derived(self, "_qkv_cache", ws, lambda: (torch.cat([m.weight.detach()[:, :, 0, 0] for m in (self.q, self.k, self.v)], 0).contiguous(),
                                                           torch.cat([m.bias.detach() for m in (self.q, self.k, self.v)], 0).contiguous()))

    def forward(self, x):
        B, C, H, W = x.shape
        T = H * W
        hn, x = group_norm_act_skip(x, self.norm, _lib.GN_PLAIN)
        tok = hn.permute(0, 2, 3, 1).reshape(B, T, C)                                 # channels-last: no copy
        Wqkv, bqkv = self._qkv()
        # the U-Net's attention core with one head (channel order {q, k, v} * C + c, scale C^-1/2): hand-written input gradient
        # whose GEMMs write dq, dk, dv into one buffer -- autograd's split backward concatenated them with a 224 us copy of
        # the [8, 4096, 1536] tensor -- and the residual add + proj_out bias in one pass of bias_add_kernel
        h = _AttentionCore.apply(tok, Wqkv, bqkv, self.proj_out.weight[:, :, 0, 0], None, 1)
        b32 = _derived(self, "_proj_bias_f32", (self.proj_out.bias,), lambda: self.proj_out.bias.detach().float().contiguous())
        return _BiasAdd.apply(x, h.view(B, H, W, C).permute(0, 3, 1, 2), b32)


class _UpsampleConv(nn.Module):
    """model.py:42-56: nearest x2 + 3x3 convolution; key `conv`."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return _conv_bias(self.conv, _Upsample2.apply(x))


class Decoder(nn.Module):
    """VQ-f4 decoder (model.py:440-568): conv_in, mid (Res, Attn, Res), up levels (num_res_blocks + 1 ResnetBlocks
    each, Upsample between levels), norm_out + swish + conv_out.  Module names follow the reference's state_dict."""

    def __init__(self, ch, out_ch, ch_mult, num_res_blocks, attn_resolutions, resolution, z_channels, **ignored):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = _UpsampleConv(block_in)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = _Norm(block_in, VQ_GN_EPS)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, padding=1)

    def forward(self, z):
        h = _conv_bias(self.conv_in, _cl(z))
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block](h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        h = group_norm_act(h, self.norm_out, _lib.GN_SILU)
        return self.conv_out(h)


class VQModelInterface(nn.Module):
    """Decode half of ldm.models.autoencoder.VQModelInterface (:274-282): quantize -> post_quant_conv -> decoder."""

    def __init__(self, embed_dim, n_embed, ddconfig, **ignored):
        super().__init__()
        self.quantize = VectorQuantizer(n_embed, embed_dim)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.decoder = Decoder(**ddconfig)

    def decode(self, h, force_not_quantize=False):
        quant = h if force_not_quantize else self.quantize(h)
        dt = self.post_quant_conv.weight.dtype
        return self.decoder(self.post_quant_conv(_cl(quant.to(dt))))


class _DiffusionWrapper(nn.Module):
    """ddpm.py DiffusionWrapper with conditioning_key None: holds `diffusion_model`."""

    def __init__(self, unet):
        super().__init__()
        self.diffusion_model = unet


FFHQ_LATENT = dict(
    linear_start=0.0015, linear_end=0.0195, timesteps=1000, scale_factor=1.0,
    unet_config=dict(image_size=64, in_channels=3, out_channels=3, model_channels=224, attention_resolutions=(8, 4, 2),
                     num_res_blocks=2, channel_mult=(1, 2, 3, 4), num_head_channels=32),
    first_stage_config=dict(embed_dim=3, n_embed=8192,
                            ddconfig=dict(double_z=False, z_channels=3, resolution=256, in_channels=3, out_ch=3, ch=128,
                                          ch_mult=(1, 2, 4), num_res_blocks=2, attn_resolutions=(), dropout=0.0)))
"""configs/config_ffhq_latent.yml:31-79 (the reference's `ffhq_latent`; identical to LDM's celebahq-ldm-vq-4)."""


class LatentDiffusion(nn.Module):
    """The members of ldm.models.diffusion.ddpm.LatentDiffusion that hmc_latent uses."""

    def __init__(self, unet_config, first_stage_config, linear_start=0.0015, linear_end=0.0195, timesteps=1000,
                 scale_factor=1.0, dtype=torch.bfloat16, **ignored):
        super().__init__()
        uc = dict(unet_config)
        unet = UNetModel(image_size=uc["image_size"], in_channels=uc["in_channels"], model_channels=uc["model_channels"],
                         out_channels=uc["out_channels"], num_res_blocks=uc["num_res_blocks"],
                         attention_resolutions=tuple(uc["attention_resolutions"]), channel_mult=tuple(uc["channel_mult"]),
                         num_heads=uc.get("num_heads", 1), num_head_channels=uc.get("num_head_channels", -1),
                         use_scale_shift_norm=uc.get("use_scale_shift_norm", False),
                         resblock_updown=uc.get("resblock_updown", False), dtype=dtype)
        self.model = _DiffusionWrapper(unet)
        fs = dict(first_stage_config)
        fs["ddconfig"] = dict(fs["ddconfig"])
        self.first_stage_model = VQModelInterface(**fs)
        self.compute_dtype = dtype
        self.scale_factor = float(scale_factor)
        ac, acp = ldm_alphas_cumprod(linear_start, linear_end, timesteps)
        self.register_buffer("alphas_cumprod", ac)
        self.register_buffer("alphas_cumprod_prev", acp)
        for p in self.parameters():
            p.requires_grad_(False)

    def prepare(self, device="cuda"):
        """Compute dtype + channels-last conv weights on the device; norm affines, the codebook and the schedule
        buffers stay fp32."""
        ac, acp = self.alphas_cumprod.float(), self.alphas_cumprod_prev.float()
        self.to(device=device, dtype=self.compute_dtype)
        self.to(memory_format=torch.channels_last)
        self.alphas_cumprod, self.alphas_cumprod_prev = ac.to(device), acp.to(device)
        return self.eval()

    @torch.no_grad()
    def apply_model(self, x_noisy, t, cond=None, return_ids=False):
        """ddpm.py:892-990 with conditioning_key None (__is_unconditional__): eps = diffusion_model(x, t)."""
        if cond is not None:
            raise _lib.NhmcError("only the unconditional latent model is on the accelerated path")
        return self.model.diffusion_model(x_noisy, t)

    def differentiable_decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        """ddpm.py:766-823: z / scale_factor -> first_stage_model.decode."""
        z = 1.0 / self.scale_factor * z
        return self.first_stage_model.decode(z, force_not_quantize=predict_cids or force_not_quantize)

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        return self.differentiable_decode_first_stage(z, predict_cids, force_not_quantize)


def create_latent_model(config=None, dtype=torch.bfloat16, ckpt_path=""):
    """`ffhq_latent` by default.  A missing checkpoint leaves the random initialisation (benchmarks on synthetic data)."""
    cfg = dict(FFHQ_LATENT if config is None else config)
    model = LatentDiffusion(dtype=dtype, **cfg)
    if ckpt_path:
        try:
            sd = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            model.load_state_dict(sd.get("state_dict", sd), strict=False)
        except Exception as e:
            print(f"Got exception: {e} / Randomly initialize")
    return model
