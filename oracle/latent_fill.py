"""Seeded parameter fill shared by tests/golden/make_golden.py (reference modules) and the GPU tests (our modules) of
the latent model family (oracle; test infrastructure).  Both sides hold the same state_dict keys in the same order,
so filling key by key from one generator gives both identical weights without committing them."""
import math

import torch

SMALL_UNET = dict(image_size=16, in_channels=3, out_channels=3, model_channels=32, attention_resolutions=(2,),
                  num_res_blocks=1, channel_mult=(1, 2), num_head_channels=16)
SMALL_FIRST_STAGE = dict(embed_dim=3, n_embed=64,
                         ddconfig=dict(double_z=False, z_channels=3, resolution=32, in_channels=3, out_ch=3, ch=32,
                                       ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(), dropout=0.0))


def seeded_fill(shapes, seed):
    """shapes: ordered {key: shape}.  Norm scales ~ 1, biases small, weights ~ 1/sqrt(fan_in), codebook ~ N(0, 0.6)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        shp = tuple(shp)
        r = torch.randn(shp, generator=g)
        if "embedding" in k:
            out[k] = 0.6 * r
        elif len(shp) == 1 and (".norm" in k or "in_layers.0" in k or "out_layers.0" in k or k.startswith("out.0") or "norm_out" in k) and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * r
        elif len(shp) == 1:
            out[k] = 0.1 * r
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            out[k] = r / math.sqrt(fan_in)
    return out
