"""Functional fp32 restatement of the ffhq_10m ADM U-Net forward (oracle; test infrastructure).

Restates UNetModel.forward and its blocks
(/root/reference/guided_diffusion/unet_ffhq.py:214-327 ResBlock, :330-376 AttentionBlock,
:399-425 QKVAttentionLegacy, :705-734 forward; /root/reference/guided_diffusion/nn.py:17-19
GroupNorm32, :103-121 timestep_embedding) as a pure function of a state_dict that uses the
reference's parameter names, so a reference checkpoint (or a reference-initialised model's
`state_dict()`) can be evaluated without importing the reference.  Plain torch ops only; it
runs on CPU and is differentiable w.r.t. x through autograd.

Pinned by tests/golden/unet_small.npz (outputs of the reference's own UNetModel).
"""
import math

import torch
import torch.nn.functional as F


DEFAULT_MULT = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}


def resolve_config(image_size, num_channels, num_res_blocks, channel_mult="", learn_sigma=False,
                   attention_resolutions="16", num_heads=1, num_head_channels=-1,
                   use_scale_shift_norm=False, resblock_updown=False, **_ignored):
    """Normalise create_model's kwargs (unet_ffhq.py:25-91) into a small dict."""
    if channel_mult == "":
        mult = DEFAULT_MULT[image_size]
    elif isinstance(channel_mult, str):
        mult = tuple(int(v) for v in channel_mult.split(","))
    else:
        mult = tuple(channel_mult)
    if isinstance(attention_resolutions, int):
        att = (image_size // attention_resolutions,)
    else:
        att = tuple(image_size // int(r) for r in str(attention_resolutions).split(","))
    return dict(image_size=image_size, ch=num_channels, nres=num_res_blocks, mult=mult, attn_ds=att,
                out_ch=6 if learn_sigma else 3, heads=num_heads, head_ch=num_head_channels,
                film=use_scale_shift_norm, updown=resblock_updown)


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(sd, key, x):
    return F.group_norm(x.float(), 32, sd[key + ".weight"], sd[key + ".bias"], 1e-5).to(x.dtype)


def _conv(sd, key, x, stride=1, padding=1):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=padding)


def _resblock(sd, pre, x, emb, film, up=False, down=False):
    h = F.silu(_gn(sd, pre + ".in_layers.0", x))
    if up:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif down:
        h = F.avg_pool2d(h, 2)
        x = F.avg_pool2d(x, 2)
    h = _conv(sd, pre + ".in_layers.2", h)
    e = F.linear(F.silu(emb), sd[pre + ".emb_layers.1.weight"], sd[pre + ".emb_layers.1.bias"])[:, :, None, None]
    if film:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = _gn(sd, pre + ".out_layers.0", h) * (1 + scale) + shift
        h = F.silu(h)
    else:
        h = F.silu(_gn(sd, pre + ".out_layers.0", h + e))
    h = _conv(sd, pre + ".out_layers.3", h)
    skip_w = pre + ".skip_connection.weight"
    if skip_w in sd:
        w = sd[skip_w]
        x = F.conv2d(x, w, sd[pre + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def _attention(sd, pre, x, heads, head_ch):
    B, C, Hh, Ww = x.shape
    n_heads = heads if head_ch == -1 else C // head_ch
    xf = x.reshape(B, C, -1)
    qkv = F.conv1d(_gn(sd, pre + ".norm", xf), sd[pre + ".qkv.weight"], sd[pre + ".qkv.bias"])
    T = qkv.shape[-1]
    ch = qkv.shape[1] // (3 * n_heads)
    q, k, v = qkv.reshape(B * n_heads, 3 * ch, T).split(ch, dim=1)
    sc = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * sc, k * sc)
    w = torch.softmax(w.float(), dim=-1).to(w.dtype)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(B, -1, T)
    h = F.conv1d(a, sd[pre + ".proj_out.weight"], sd[pre + ".proj_out.bias"])
    return (xf + h).reshape(B, C, Hh, Ww)


def build_plan(cfg):
    """List the blocks in module order with the state_dict prefixes the reference uses."""
    ch0 = int(cfg["mult"][0] * cfg["ch"])
    inp, ds, ch = [("conv", "input_blocks.0.0")], 1, ch0
    idx = 1
    for level, m in enumerate(cfg["mult"]):
        for _ in range(cfg["nres"]):
            blk = [("res", f"input_blocks.{idx}.0", {})]
            ch = int(m * cfg["ch"])
            if ds in cfg["attn_ds"]:
                blk.append(("attn", f"input_blocks.{idx}.1", {}))
            inp.append(blk); idx += 1
        if level != len(cfg["mult"]) - 1:
            if cfg["updown"]:
                inp.append([("res", f"input_blocks.{idx}.0", {"down": True})])
            else:
                inp.append([("downconv", f"input_blocks.{idx}.0.op", {})])
            idx += 1
            ds *= 2
    mid = [("res", "middle_block.0", {}), ("attn", "middle_block.1", {}), ("res", "middle_block.2", {})]
    out, idx = [], 0
    for level, m in list(enumerate(cfg["mult"]))[::-1]:
        for i in range(cfg["nres"] + 1):
            blk = [("res", f"output_blocks.{idx}.0", {})]
            k = 1
            if ds in cfg["attn_ds"]:
                blk.append(("attn", f"output_blocks.{idx}.{k}", {})); k += 1
            if level and i == cfg["nres"]:
                if cfg["updown"]:
                    blk.append(("res", f"output_blocks.{idx}.{k}", {"up": True}))
                else:
                    blk.append(("upconv", f"output_blocks.{idx}.{k}.conv", {}))
                ds //= 2
            out.append(blk); idx += 1
    return inp, mid, out


def _run(sd, cfg, blk, h, emb):
    for kind, pre, kw in blk:
        if kind == "res":
            h = _resblock(sd, pre, h, emb, cfg["film"], **kw)
        elif kind == "attn":
            h = _attention(sd, pre, h, cfg["heads"], cfg["head_ch"])
        elif kind == "downconv":
            h = _conv(sd, pre, h, stride=2)
        elif kind == "upconv":
            h = _conv(sd, pre, F.interpolate(h, scale_factor=2, mode="nearest"))
    return h


def unet_forward(sd, cfg, x, t):
    """x: [B,3,H,W] fp32, t: [B] float -> [B,out_ch,H,W]; unet_ffhq.py:705-734."""
    inp, mid, out = build_plan(cfg)
    emb = timestep_embedding(t, cfg["ch"])
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = _conv(sd, inp[0][1], x)
    hs.append(h)
    for blk in inp[1:]:
        h = _run(sd, cfg, blk, h, emb)
        hs.append(h)
    h = _run(sd, cfg, mid, h, emb)
    for blk in out:
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run(sd, cfg, blk, h, emb)
    h = F.silu(_gn(sd, "out.0", h))
    return _conv(sd, "out.2", h)


def param_shapes(cfg):
    """{name: shape} for every parameter, in the reference's registration order."""
    shapes = {}
    C = cfg["ch"]
    E = 4 * C
    shapes["time_embed.0.weight"] = (E, C); shapes["time_embed.0.bias"] = (E,)
    shapes["time_embed.2.weight"] = (E, E); shapes["time_embed.2.bias"] = (E,)

    def res(pre, cin, cout):
        shapes[pre + ".in_layers.0.weight"] = (cin,); shapes[pre + ".in_layers.0.bias"] = (cin,)
        shapes[pre + ".in_layers.2.weight"] = (cout, cin, 3, 3); shapes[pre + ".in_layers.2.bias"] = (cout,)
        ne = 2 * cout if cfg["film"] else cout
        shapes[pre + ".emb_layers.1.weight"] = (ne, E); shapes[pre + ".emb_layers.1.bias"] = (ne,)
        shapes[pre + ".out_layers.0.weight"] = (cout,); shapes[pre + ".out_layers.0.bias"] = (cout,)
        shapes[pre + ".out_layers.3.weight"] = (cout, cout, 3, 3); shapes[pre + ".out_layers.3.bias"] = (cout,)
        if cin != cout:
            shapes[pre + ".skip_connection.weight"] = (cout, cin, 1, 1); shapes[pre + ".skip_connection.bias"] = (cout,)

    def attn(pre, c):
        shapes[pre + ".norm.weight"] = (c,); shapes[pre + ".norm.bias"] = (c,)
        shapes[pre + ".qkv.weight"] = (3 * c, c, 1); shapes[pre + ".qkv.bias"] = (3 * c,)
        shapes[pre + ".proj_out.weight"] = (c, c, 1); shapes[pre + ".proj_out.bias"] = (c,)

    ch = int(cfg["mult"][0] * C)
    shapes["input_blocks.0.0.weight"] = (ch, 3, 3, 3); shapes["input_blocks.0.0.bias"] = (ch,)
    chans, ds, idx = [ch], 1, 1
    for level, m in enumerate(cfg["mult"]):
        for _ in range(cfg["nres"]):
            co = int(m * C)
            res(f"input_blocks.{idx}.0", ch, co); ch = co
            if ds in cfg["attn_ds"]:
                attn(f"input_blocks.{idx}.1", ch)
            chans.append(ch); idx += 1
        if level != len(cfg["mult"]) - 1:
            if cfg["updown"]:
                res(f"input_blocks.{idx}.0", ch, ch)
            else:
                shapes[f"input_blocks.{idx}.0.op.weight"] = (ch, ch, 3, 3); shapes[f"input_blocks.{idx}.0.op.bias"] = (ch,)
            chans.append(ch); idx += 1; ds *= 2
    res("middle_block.0", ch, ch); attn("middle_block.1", ch); res("middle_block.2", ch, ch)
    idx = 0
    for level, m in list(enumerate(cfg["mult"]))[::-1]:
        for i in range(cfg["nres"] + 1):
            ich = chans.pop(); co = int(C * m)
            res(f"output_blocks.{idx}.0", ch + ich, co); ch = co
            k = 1
            if ds in cfg["attn_ds"]:
                attn(f"output_blocks.{idx}.{k}", ch); k += 1
            if level and i == cfg["nres"]:
                if cfg["updown"]:
                    res(f"output_blocks.{idx}.{k}", ch, ch)
                else:
                    shapes[f"output_blocks.{idx}.{k}.conv.weight"] = (ch, ch, 3, 3); shapes[f"output_blocks.{idx}.{k}.conv.bias"] = (ch,)
                ds //= 2
            idx += 1
    shapes["out.0.weight"] = (ch,); shapes["out.0.bias"] = (ch,)
    shapes["out.2.weight"] = (cfg["out_ch"], int(cfg["mult"][0] * C), 3, 3); shapes["out.2.bias"] = (cfg["out_ch"],)
    return shapes


def seeded_state_dict(cfg, seed=1234, std=0.02):
    """Deterministic synthetic weights, independent of any nn.Module init order.

    Keys are visited in sorted order; norm weights ~ 1 + std*N(0,1), everything else std*N(0,1)
    (so the reference's zero-initialised convs are non-zero, SURVEY.md headline 5).
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    shapes = param_shapes(cfg)
    for k in sorted(shapes):
        w = torch.randn(shapes[k], generator=g) * std
        is_norm = (k.endswith(".weight") and len(shapes[k]) == 1)
        sd[k] = (1.0 + w) if is_norm else w
    return {k: sd[k] for k in shapes}      # values drawn in sorted-key order, dict in registration order
