#!/usr/bin/env python
"""Per-kernel roofline microbenchmark: each custom kernel alone, on tensors larger than L2, CUDA-event timed.

    python tools/kernel_microbench.py [--only gn] [--reps 20]

Prints one JSON line per (kernel, shape): algorithmic bytes, microseconds, GB/s, fraction of the measured HBM peak.
Used with `ncu --set full` for the per-kernel captures committed under profiles/.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nhmc_b200  # noqa: E402
from nhmc_b200 import _lib, kernels, operators as P  # noqa: E402
from nhmc_b200.chain import ChainState, new_host_state  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.isfile(p) else 6650.0


GRAPH_CALLS = 1


def timeit(fn, reps, flush=None):
    """Median CUDA-event time of fn() with a cold, CLEAN L2: the flush buffer (> L2) is written, then read back, so that the
    timed kernel neither hits in L2 nor pays for the write-back of the flush's dirty lines.  Flush, start event, the call
    and the stop event are ONE captured CUDA graph (event-record nodes), replayed `reps` times: the host's launch latency
    (ctypes + python, several microseconds -- comparable to these kernels) never sits between the two events."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    calls = max(GRAPH_CALLS, 1)
    s = torch.cuda.Event(enable_timing=True, external=True)
    e = torch.cuda.Event(enable_timing=True, external=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        if flush is not None:
            flush.zero_()
            flush.view(torch.int32).sum()
        s.record()
        for _ in range(calls):
            fn()
        e.record()
    g.replay(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        g.replay()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    ms.sort()
    return ms[len(ms) // 2] / calls


def report(name, shape, nbytes, ms, pk):
    gbs = nbytes / 1e9 / (ms / 1e3)
    print(json.dumps({"kernel": name, "shape": shape, "alg_MB": round(nbytes / 1e6, 2), "us": round(ms * 1e3, 2),
                      "GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / pk, 3)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--shape", default="", help="GroupNorm: only this B,C,H (e.g. 8,128,256)")
    ap.add_argument("--resident", type=int, default=1, help="GroupNorm: 1 = shared-memory resident single kernel, 0 = statistics + apply kernels")
    ap.add_argument("--solo", type=int, default=2, help="GroupNorm: 1 = one-CTA-per-unit kernel for the small maps, 2 = + cluster-per-unit kernel for the 64x64 maps, 0 = off")
    ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch of the U-Net epilogue kernels (1 = default)")
    ap.add_argument("--warm", action="store_true", help="no L2 flush between repetitions (inputs L2-resident, as inside the U-Net for the small maps)")
    ap.add_argument("--small", action="store_true", help="GroupNorm: the small-map shapes of the ffhq_10m U-Net at 8 chains")
    ap.add_argument("--graph-calls", type=int, default=1, help="time this many back-to-back calls replayed from one CUDA graph (per-call time reported)")
    a = ap.parse_args()
    global GRAPH_CALLS
    GRAPH_CALLS = a.graph_calls
    pk = peak()
    dev = torch.device("cuda")
    flush = None if a.warm else torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    want = lambda k: (not a.only) or a.only in k
    kernels.set_flags((0 if a.resident else _lib.GN_F_NO_RESIDENT) | (_lib.GN_F_NO_SOLO if a.solo == 0 else (_lib.GN_F_NO_SOLO_CLUSTER if a.solo == 1 else 0))
                      | (0 if a.pdl else _lib.F_NO_PDL))

    gn_scope = kernels.gn_scope(kernels.GnWorkspace(dev, 64))      # caller-owned GroupNorm scratch, shared by eager calls and graph captures
    gn_scope.__enter__()
    if want("gn"):
        shapes = [(8, 128, 256), (8, 256, 256), (8, 256, 128), (8, 384, 128), (8, 512, 64), (8, 768, 64), (8, 1024, 32), (8, 1024, 16), (8, 512, 8), (1, 128, 256)]
        if a.small:
            shapes = [(8, 512, 8), (8, 1024, 8), (8, 256, 16), (8, 512, 16), (8, 768, 16), (8, 1024, 16), (8, 256, 32), (8, 512, 32),
                      (8, 768, 32), (8, 128, 64), (8, 256, 64), (8, 512, 64), (8, 128, 128), (1, 512, 8), (1, 256, 32)]
        if a.shape:
            shapes = [tuple(int(v) for v in a.shape.split(","))]
        for (B, C, H) in shapes:
            for dt in (torch.bfloat16,):
                x = torch.randn(B, C, H, H, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
                dy = torch.randn_like(x); y = torch.empty_like(x); dx = torch.empty_like(x)
                gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
                sc = torch.zeros(B, C, device=dev); sh = torch.zeros(B, C, device=dev)
                stats = torch.empty(B, 32, 2, device=dev)
                nb = x.numel() * x.element_size()
                for mode, nm in ((_lib.GN_SILU, "silu"),) if a.small else ((_lib.GN_SILU, "silu"), (_lib.GN_FILM_SILU, "film")):
                    s_, h_ = (sc, sh) if mode == _lib.GN_FILM_SILU else (None, None)
                    ms = timeit(lambda: kernels.gn_fwd(mode, x, y, gamma, beta, s_, h_, stats, 32, 1e-5, True), a.reps, flush)
                    report("gn_fwd_" + nm, [B, C, H, H], 2 * nb, ms, pk)
                    ms = timeit(lambda: kernels.gn_bwd(mode, x, dy, dx, gamma, beta, s_, h_, stats, 32, True), a.reps, flush)
                    report("gn_bwd_" + nm, [B, C, H, H], 3 * nb, ms, pk)
                kernels.gn_check()
    if want("leapfrog") or want("ddim") or want("op"):
        B = a.chains
        N = 3 * 256 * 256
        x = torch.randn(B, 3, 256, 256, device=dev); p = torch.randn_like(x); g = torch.randn_like(x); g2 = torch.randn_like(x)
        xo = torch.empty_like(x)
        ch = ChainState(new_host_state(B, 1.0, 0.05, 0.1), dev)
        kernels.begin_iteration(ch, _lib.SCHED_PIXEL)
        ws = kernels.leapfrog_workspace(B, dev)
        loss = torch.zeros(B, dtype=torch.float64, device=dev)
        if want("leapfrog"):
            for mode, nm, nbuf in ((_lib.LF_FIRST, "first", 6), (_lib.LF_MIDDLE, "middle", 6), (_lib.LF_LAST, "last", 5)):
                ms = timeit(lambda: kernels.leapfrog(mode, x, xo, p, g, g2, loss, ch, 1.0, ws), a.reps, flush)
                report("leapfrog_" + nm, [B, 3, 256, 256], nbuf * B * N * 4, ms, pk)
        if want("ddim"):
            eps = torch.randn(B, 6, 256, 256, device=dev, dtype=torch.bfloat16)
            mask = torch.empty(B, kernels.ddim_mask_words(N), dtype=torch.int32, device=dev)
            coef = (0.96, 0.28, 0.69, 0.72)
            ms = timeit(lambda: kernels.ddim_fwd(x, eps, xo, mask, coef), a.reps, flush)
            report("ddim_fwd", [B, 3, 256, 256], B * N * (4 + 2 + 4) + B * N // 8, ms, pk)
            ge = torch.zeros_like(eps); gx = torch.empty_like(x)
            ms = timeit(lambda: kernels.ddim_bwd(g, g2, mask, gx, ge, coef), a.reps, flush)
            report("ddim_bwd", [B, 3, 256, 256], B * N * (4 + 4 + 4 + 2) + B * N // 8, ms, pk)
        if want("op"):
            torch.manual_seed(5678)
            ops = {"inpaint": P.Inpainting.random(3, 256, dev), "sr4": P.SuperResolution(3, 256, 4, dev),
                   "sr16": P.SuperResolution(3, 256, 16, dev), "hdr": P.HDR(3, 256, dev),
                   "deblur": P.Deblurring2D(P.gaussian_taps(1.0), P.gaussian_taps(20.0), 3, 256, dev),
                   "phase": P.PhaseRetrievalOperator(2.0, dev, 3, 256)}
            for nm, op in ops.items():
                Bo = B if nm not in ("deblur", "phase") else min(B, 16)
                xs = x[:Bo].contiguous()
                y = op.H(xs)
                gg = torch.empty_like(xs); xc = torch.empty_like(xs); ls = torch.zeros(Bo, dtype=torch.float64, device=dev)
                ms = timeit(lambda: op.loss_grad(xs, y, g=gg, xc=xc, loss=ls), a.reps, flush)
                report("op_loss_grad_" + nm, [Bo, 3, 256, 256], (3 * xs.numel() + y.numel()) * 4, ms, pk)
                if nm == "deblur":
                    # tensor-bound: 4 products of 2 * 256^3 flop per (chain, channel), each issued as 3 tf32 MMAs (3xTF32)
                    fl = 4 * Bo * 3 * 2 * 256 ** 3
                    pkf = os.path.join(ROOT, "MEASURED_PEAKS.json")
                    tf32_peak = (float(json.load(open(pkf))["bf16_tflops"]) if os.path.isfile(pkf) else 1590.0) / 2     # tf32 = half the bf16 rate
                    print(json.dumps({"kernel": "op_loss_grad_deblur", "bound": "tensor", "shape": [Bo, 3, 256, 256], "us": round(ms * 1e3, 2),
                                      "fp32_grade_TFLOPs": round(fl / 1e12 / (ms / 1e3), 1), "tf32_issued_TFLOPs": round(3 * fl / 1e12 / (ms / 1e3), 1),
                                      "tf32_peak_TFLOPs": tf32_peak, "frac_of_tf32_peak": round(3 * fl / 1e12 / (ms / 1e3) / tf32_peak, 3),
                                      "note": "tf32 peak taken as half the measured cuBLAS bf16 burst peak (MEASURED_PEAKS.json)"}), flush=True)
    if want("pool") or want("split"):
        B, C, H = 8, 256, 128
        big = torch.randn(B, C, 2 * H, 2 * H, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        small = torch.empty(B, C, H, H, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        nb = (big.numel() + small.numel()) * 2
        report("pool2_gather", [B, C, 2 * H, 2 * H], nb, timeit(lambda: kernels.pool2_gather(big, small, 0.25), a.reps, flush), pk)
        report("pool2_scatter", [B, C, H, H], nb, timeit(lambda: kernels.pool2_scatter(small, big, 1.0), a.reps, flush), pk)
        o1 = torch.empty(B, 128, 2 * H, 2 * H, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        o2 = torch.empty_like(o1)
        report("split2", [B, C, 2 * H, 2 * H], 2 * big.numel() * 2, timeit(lambda: kernels.split2(big, o1, o2), a.reps, flush), pk)

    if want("attn"):
        # the latent U-Net's three attention shapes at 8 chains (ffhq_latent: 224 * (2, 3, 4) channels, 32 per head): the fused
        # forward against Q K^T (cuBLAS) -> softmax kernel -> P V (cuBLAS), same operands.  Bytes = q, k, v in and out once.
        import math
        for (B, T, nh, ch) in ((8, 1024, 14, 32), (8, 256, 21, 32), (8, 64, 28, 32)):
            C = nh * ch
            qk = torch.randn(B, T, 2 * C, device=dev, dtype=torch.bfloat16)
            v = torch.randn(B, C, T, device=dev, dtype=torch.bfloat16)
            out = torch.empty(B, T, C, device=dev, dtype=torch.bfloat16)
            scale = 1 / math.sqrt(ch)
            nb = (qk.numel() + v.numel() + out.numel()) * 2
            ms = timeit(lambda: kernels.attn_fwd(qk, v, out, nh, scale), a.reps, flush)
            report("attn_fwd_fused", [B, T, nh, ch], nb, ms, pk)
            exps = B * nh * T * T
            print(json.dumps({"kernel": "attn_fwd_fused", "bound": "sfu (ex2: 16 / clk / SM)", "shape": [B, T, nh, ch], "us": round(ms * 1e3, 2),
                              "Gexp_per_s": round(exps / 1e9 / (ms / 1e3), 1), "peak_Gexp_per_s_at_1.9GHz": round(148 * 16 * 1.9, 1),
                              "frac": round(exps / 1e9 / (ms / 1e3) / (148 * 16 * 1.9), 3)}), flush=True)
            qkv = torch.randn(B, 3 * C, T, device=dev, dtype=torch.bfloat16)
            q_, k_, v_ = qkv.view(B * nh, 3 * ch, T).split(ch, dim=1)
            probs = torch.empty(B * nh, T, T, device=dev, dtype=torch.bfloat16)

            def unfused():
                logits = torch.bmm(q_.transpose(1, 2), k_)
                kernels.softmax_fwd(logits, probs, scale)
                return torch.bmm(v_, probs.transpose(1, 2))
            ms = timeit(unfused, a.reps, flush)
            report("attn_fwd_unfused(bmm+softmax+bmm)", [B, T, nh, ch], nb, ms, pk)


if __name__ == "__main__":
    main()
