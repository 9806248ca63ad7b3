#!/usr/bin/env python
"""L2 retention probe on the GPU at hand: (1) a plain reduction read twice in a row as a function of the buffer size;
(2) CUPTI durations of the GroupNorm statistics and apply kernels (two-kernel path forced) for 1..8 samples of
[*, 128, 256, 256]: does the apply kernel's re-read of x hit the L2?"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nhmc_b200  # noqa: E402
from nhmc_b200 import _lib, kernels  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def ev_time(fn):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3


for mb in (8, 16, 32, 48, 64, 96, 128, 192, 256):
    x = torch.empty(mb << 18, dtype=torch.float32, device=dev).normal_()
    x.sum(); torch.cuda.synchronize()
    res = []
    for _ in range(5):
        flush.zero_(); flush.view(torch.int32).sum(); torch.cuda.synchronize()
        t_cold = ev_time(lambda: x.sum())
        t_warm = ev_time(lambda: x.sum())
        res.append((t_cold, t_warm))
    res.sort()
    c, w = res[len(res) // 2]
    print(json.dumps({"probe": "sum_twice", "MB": mb, "cold_us": round(c, 1), "warm_us": round(w, 1),
                      "cold_GBps": round(mb * 1.048576 / c * 1e3, 0), "warm_GBps": round(mb * 1.048576 / w * 1e3, 0)}), flush=True)

kernels.set_flags(_lib.GN_F_NO_RESIDENT | _lib.GN_F_NO_SOLO)
for B in (1, 2, 3, 4, 6, 8):
    C, H = 128, 256
    x = torch.randn(B, C, H, H, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn_like(x); y = torch.empty_like(x); dx = torch.empty_like(x)
    gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    stats = torch.empty(B, 32, 2, device=dev)
    for _ in range(2):
        kernels.gn_fwd(_lib.GN_SILU, x, y, gamma, beta, None, None, stats, 32, 1e-5, True)
        kernels.gn_bwd(_lib.GN_SILU, x, dy, dx, gamma, beta, None, None, stats, 32, True)
    torch.cuda.synchronize()
    for name, fn in (("fwd", lambda: kernels.gn_fwd(_lib.GN_SILU, x, y, gamma, beta, None, None, stats, 32, 1e-5, True)),
                     ("bwd", lambda: kernels.gn_bwd(_lib.GN_SILU, x, dy, dx, gamma, beta, None, None, stats, 32, True))):
        flush.zero_(); flush.view(torch.int32).sum(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            fn(); torch.cuda.synchronize()
        d = {("stats" if "stats" in e.name else "apply"): e.device_time for e in p.events() if "gn_nhwc" in (e.name or "")}
        mb = x.numel() * 2 / 1e6
        print(json.dumps({"probe": "gn_" + name, "B": B, "tensor_MB": round(mb, 1), "stats_us": d.get("stats"), "apply_us": d.get("apply")}), flush=True)
