#!/usr/bin/env python
"""bench.py -- HMC leapfrog steps/s (chains x steps) at FFHQ-256, the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--deg sr4] [--chains 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE leapfrog step for every chain resident on the GPU (main_sampling.py:706-713): position update,
3-step DDIM map (3 U-Net forwards + fused epilogues), measurement loss + adjoint, full VJP (3 U-Net input-VJPs +
fused epilogue backwards) and the momentum update.  Workload = BASELINE config[1]: FFHQ-256 sr4, 8 chains per
GPU, random-init ffhq_10m-architecture U-Net (bf16, channels-last), synthetic measurements.  Chains are
independent: at N GPUs every rank runs its own 8 chains (weak scaling), no collective in the timed loop.

Rank 0 prints ONE JSON line (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "hmc_leapfrog_steps_per_sec"
UNIT = "chain-steps/s"
FFHQ = dict(image_size=256, num_channels=128, num_res_blocks=1, channel_mult="", learn_sigma=True, class_cond=False,
            use_checkpoint=False, attention_resolutions=16, num_heads=4, num_head_channels=64, num_heads_upsample=-1,
            use_scale_shift_norm=True, dropout=0.0, resblock_updown=True, use_fp16=False, use_new_attention_order=False)
SEQ, SEQ_NEXT = [250, 500, 750], [-1, 250, 500]
UNET_FWD_GFLOP = 387.9          # per 1x3x256x256 forward (SURVEY.md headline 4, FlopCounterMode on the reference)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--deg", default="sr4", help="sr4 | sr16 | inpaint_random | inpaint_box | hdr | deblur_aniso | deblur_gauss | phase_retrieval | color")
    ap.add_argument("--chains", type=int, default=8, help="chains per GPU")
    ap.add_argument("--cudnn-benchmark", type=int, default=1, help="torch.backends.cudnn.benchmark for the library convolutions (1 = autotune at warm-up)")
    ap.add_argument("--cudnn-limit", type=int, default=-1, help="torch.backends.cudnn.benchmark_limit (-1 = torch's default of 10, 0 = all algorithms)")
    ap.add_argument("--workload", default="pixel", choices=["pixel", "latent"], help="pixel: --algo hmc (BASELINE config[1], the metric's configuration); latent: --algo hmc_latent on the ffhq_latent architecture (BASELINE config[4])")
    ap.add_argument("--lanes", type=int, default=1, help="groups of chains evaluated on their own streams (parallel branches of the captured step)")
    ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch of the U-Net epilogue kernels (1 = library default)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32-line", action="store_true", help="skip the fp32 (TF32 off) timing + bf16-vs-fp32 comparison that accompanies the bf16 line at N=1")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-in-situ", action="store_true", help="skip the CUPTI breakdown of one step after the timed region")
    ap.add_argument("--profile-kernels", action="store_true", default=True)
    ap.add_argument("--no-graphs", action="store_true", help="run the step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--torch-profile", default="", help="write a torch.profiler kernel table of one step to this file")
    ap.add_argument("--cuda-profiler-range", action="store_true", help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, bf16=1400.0, src="fallback")


# ----------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# workload construction (synthetic FFHQ-256 measurements)
# ----------------------------------------------------------------------------------------------------
def make_operator(deg, device):
    from nhmc_b200 import operators as P
    if deg.startswith("sr"):
        return P.SuperResolution(3, 256, int(deg[2:]), device)
    if "inp" in deg and "box" in deg:
        return P.Inpainting.box(3, 256, device, left=64, up=64)
    if "inp" in deg:
        torch.manual_seed(5678)
        return P.Inpainting.random(3, 256, device)
    if "deblur_gauss" in deg:
        return P.Deblurring(P.gaussian_taps(10.0, half=2), 3, 256, device)
    if deg == "color":
        return P.Colorization(256, device)
    if "hdr" in deg:
        return P.HDR(3, 256, device)
    if deg == "deblur_aniso":
        return P.Deblurring2D(P.gaussian_taps(1.0), P.gaussian_taps(20.0), 3, 256, device)
    if "phase" in deg:
        return P.PhaseRetrievalOperator(2.0, device, 3, 256)
    raise SystemExit("unknown --deg " + deg)


def run_ours(args, rank, world, local_rank):
    import nhmc_b200
    from nhmc_b200 import _lib, kernels, sampler, schedule, unet
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    if args.cudnn_limit >= 0:
        torch.backends.cudnn.benchmark_limit = args.cudnn_limit           # 0 = try every algorithm cuDNN offers
    kernels.set_flags(0 if args.pdl else _lib.F_NO_PDL)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B, K, W = args.chains, args.steps, max(args.warmup, 3)

    torch.manual_seed(1234)
    latent_wl = args.workload == "latent"
    op = make_operator(args.deg, dev)
    gen = torch.Generator().manual_seed(5678 + rank)
    x_orig = torch.rand(B, 3, 256, 256, generator=gen) * 2 - 1
    sigma_0 = 0.1                                                                 # --sigma_0 0.05, doubled
    y = op.H(x_orig.to(dev))
    y = y + sigma_0 * torch.randn(y.shape, generator=gen).to(dev)
    if not latent_wl:
        model = unet.create_model(**FFHQ, dtype=dtype)
        unet.randomize_zero_tensors(model, seed=1234)
        model.prepare(dev)
        state_shape = (B, 3, 256, 256)
        x0 = torch.randn(state_shape, generator=gen)
        betas = torch.from_numpy(schedule.get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
        steps = schedule.ddim_steps(schedule.alpha_table(betas), SEQ, SEQ_NEXT)
        ev = sampler.GradEvaluator(model, steps, op, y, x0.to(dev))
        smp = sampler.HMCSampler(x0.to(dev), ev, 1.0, 0.05, sigma_0, sched=_lib.SCHED_PIXEL, lanes=args.lanes)
        sched_kind = _lib.SCHED_PIXEL
    else:
        from nhmc_b200 import latent
        ldm = latent.create_latent_model(dtype=dtype)                             # ffhq_latent (config_ffhq_latent.yml)
        unet.randomize_zero_tensors(ldm, seed=1234)
        ldm.prepare(dev)
        op.set_clamp(False)                                                       # main_sampling_latent.py:651
        state_shape = (B, 3, 64, 64)
        x0 = torch.randn(state_shape, generator=gen)
        steps = schedule.ddim_steps(schedule.latent_alpha_table(ldm.alphas_cumprod, ldm.alphas_cumprod_prev), SEQ, SEQ_NEXT)
        ev = sampler.GradEvaluator(lambda z, t: ldm.apply_model(z, t, None), steps, op, y, x0.to(dev), net_grad=False,
                                   decode=ldm.differentiable_decode_first_stage)
        smp = sampler.HMCSampler(x0.to(dev), ev, 1.0, 0.05, sigma_0, sched=_lib.SCHED_LATENT, sigma_y=0.5, n_slots=10,
                                 lanes=args.lanes)
        sched_kind = _lib.SCHED_LATENT
    kernels.begin_iteration(smp.chains, sched_kind)                                # sigma_y for epoch 0
    smp.p.copy_(torch.randn(state_shape, generator=gen).to(dev))
    smp.step(_lib.LF_FIRST)
    use_graphs = not args.no_graphs
    if use_graphs:
        smp.capture_graphs()

    def step():
        smp.step(_lib.LF_MIDDLE)

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events, per-kernel events for our kernels ----
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.3)
    launches0 = kernels.LAUNCHES
    if not use_graphs:
        kernels.profile_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.cuda_profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    if args.cuda_profiler_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = kernels.LAUNCHES - launches0
    clk = clocks.stop()
    if use_graphs:
        # per-kernel CUDA events IN SITU: the same step captured once more with an event-record node on either side of
        # every kernel of ours, replayed K times right after the timed region (same buffers, same launch sequence,
        # warm L2); the timed region itself carries no events
        pg, precs = smp.capture_profiled_step(_lib.LF_MIDDLE)
        pg.replay(); torch.cuda.synchronize()
        prof_sz = {}
        for _ in range(K):
            pg.replay()
            kernels.profile_read(precs, prof_sz)
        del pg
    else:
        prof_sz = kernels.profile_stop(by_size=True)
    prof = {}
    for (name, _b), (n, tms, nbytes) in prof_sz.items():
        a, b_, c = prof.get(name, (0, 0.0, 0))
        prof[name] = (a + n, b_ + tms, c + nbytes)
    finite = bool(torch.isfinite(smp.x_prop).all() and torch.isfinite(smp.p).all())

    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())

    in_situ = None
    if rank == 0 and not args.no_in_situ:
        # one graph replay under CUPTI (torch.profiler): device-side duration of every kernel IN SITU (warm L2,
        # back-to-back launches) -- explains the step; the timed region above ran without any profiler
        from torch.profiler import profile, ProfilerActivity
        step_p = step
        if use_graphs and args.pdl:
            # with programmatic dependent launch a kernel's CTAs are scheduled while its predecessor drains, so CUPTI's
            # per-kernel durations would include that wait: the breakdown is taken on a capture with plain launches
            kernels.set_flags(_lib.F_NO_PDL)
            g_plain, _ = smp.capture_profiled_step(_lib.LF_MIDDLE, events=False)
            kernels.set_flags(0)
            g_plain.replay(); torch.cuda.synchronize()
            step_p = g_plain.replay
        with profile(activities=[ProfilerActivity.CUDA]) as prof_t:
            step_p()
            torch.cuda.synchronize()
        fam = {}
        for ev in prof_t.events():
            if ev.device_type.name != "CUDA" or not ev.name or ev.name.startswith(("Memcpy", "Memset")):
                continue
            nm = ev.name
            if "nhmc::" in nm:
                key = nm.split("(")[0].replace("void ", "")
            elif any(t in nm.lower() for t in ("cutlass", "nvjet", "cudnn", "gemm", "nhwcaddpadding")):
                key = "library conv/gemm (cuDNN, cuBLAS)"
            else:
                key = "aten elementwise/copy"
            n, us = fam.get(key, (0, 0.0))
            fam[key] = (n + 1, us + ev.device_time)
        tot = sum(us for _, us in fam.values())
        if tot > 0:
            ours_n = sum(n for k, (n, _) in fam.items() if k.startswith("nhmc::"))
            in_situ = {"sum_ms": tot / 1e3, "ours_launches_per_step": ours_n, "all_launches_per_step": sum(n for n, _ in fam.values()), "families": {k: {"launches": n, "ms": us / 1e3} for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:24]},
                       "ours_ms": sum(us for k, (_, us) in fam.items() if k.startswith("nhmc::")) / 1e3,
                       "library_ms": fam.get("library conv/gemm (cuDNN, cuBLAS)", (0, 0.0))[1] / 1e3,
                       "note": "CUPTI kernel durations of ONE graph-replayed step after the timed region (captured with plain launches: "
                               "under programmatic dependent launch a kernel's duration would include its wait on the predecessor)"}
        if args.torch_profile:
            with open(args.torch_profile, "w") as f:
                f.write(prof_t.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
            prof_t.export_chrome_trace(args.torch_profile + ".trace.json")     # per-launch durations + grid sizes

    # ---- end-to-end: the same step through the public API with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(state_shape).pin_memory(); ph = torch.empty(state_shape).pin_memory()
        xh.copy_(smp.x_prop); ph.copy_(smp.p)
        for _ in range(2):
            smp.leapfrog_step_host(xh, ph)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        Ke = max(2, min(K, 10))
        t0.record()
        for _ in range(Ke):
            res = smp.leapfrog_step_host(xh, ph)
        t1.record()
        barrier()
        ms_e = t0.elapsed_time(t1)
        if dist is not None:
            t = torch.tensor([ms_e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        e2e = {"value": world * B * Ke / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(2 * xh.numel() * 4),
               "d2h_bytes_per_step": int(res["bytes"]), "steps": Ke, "ms_per_step": ms_e / Ke}

    # ---- the same workload with an fp32 score network (TF32 off, the reference's precision: use_fp16 False) ----
    fp32_line = None
    if (not latent_wl and args.dtype == "bf16" and world == 1 and not args.no_fp32_line):
        fp32_line = fp32_companion(args, dev, op, y, x_orig, x0, steps, smp, sigma_0, B, max(3, min(K, 5)))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = world * B * K / (ms / 1e3)
    # ---- roofline: our dominant kernel (family, tensor shape) inside the step -- HBM-bound GroupNorm epilogue ----
    kern = {}
    for name, (n, tms, nbytes) in sorted(prof.items()):
        kern[name] = {"calls_per_step": n / K, "ms_per_step": tms / K, "GBps": (nbytes / 1e9) / (tms / 1e3) if tms > 0 else None}
    ours_ms = sum(v["ms_per_step"] for v in kern.values())
    roofline = None
    if prof_sz:
        (dname, dbytes), (dn, dms, dtot) = max(prof_sz.items(), key=lambda kv: kv[1][1])
        ach = (dtot / 1e9) / (dms / 1e3)
        traffic = None
        for tn in ("r2_traffic.json", "r1_traffic.json"):               # dram bytes per call from the committed ncu --set full capture
            tp = os.path.join(ROOT, "profiles", tn)
            if os.path.isfile(tp):
                traffic = json.load(open(tp)).get("%s:%d" % (dname, dbytes))
                if traffic is not None:
                    break
        roofline = {"kernel": dname, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": ach / pk["hbm"], "traffic": traffic, "peak_source": pk["src"],
                    "algorithmic_bytes_per_call": int(dbytes), "calls_per_step": dn / K, "us_per_call": dms / dn * 1e3,
                    "family_all_shapes": {"GBps": kern[dname]["GBps"], "frac": kern[dname]["GBps"] / pk["hbm"], "calls_per_step": kern[dname]["calls_per_step"]},
                    "note": "dominant (kernel family, tensor shape) of ours by time; algorithmic bytes (forward: x read + y written; backward: "
                            "x, dy (+ skip gradient) read + dx written) / CUDA-event time per call on the launching stream, averaged over K steps"
                            + (" (event-record nodes in a second capture of the same step, replayed right after the timed region)" if use_graphs else " of the timed region")
                            + "; traffic = dram bytes per call from the ncu --set full capture in profiles/ (null if that shape was not captured)"}
    # every (kernel family, algorithmic bytes per call) = tensor shape of ours, in situ: where the epilogue time goes
    by_shape = [{"kernel": nm, "algorithmic_bytes_per_call": int(nb_), "calls_per_step": n / K, "us_per_call": tms / n * 1e3,
                 "ms_per_step": tms / K, "GBps": (tot / 1e9) / (tms / 1e3) if tms > 0 else None}
                for (nm, nb_), (n, tms, tot) in sorted(prof_sz.items(), key=lambda kv: -kv[1][1])[:40]]
    step_ms = ms / K
    if latent_wl:   # 3 latent U-Net forwards (no VJP: eps carries no gradient) + decoder forward and VJP
        unet_tflop = B * (3 * 202.4 + 2 * 670.8) / 1e3
    else:
        unet_tflop = B * 3 * 2 * UNET_FWD_GFLOP / 1e3                   # fwd + input-VJP ~ 2x fwd FLOPs, 3 timesteps
    # shares of the step: from the in-situ CUPTI durations when available (the eager event times of small kernels
    # include launch gaps), else from the per-call events
    sampler_names = ("leapfrog", "ddim_fwd", "ddim_bwd", "begin_iteration", "accept", "commit")
    sampler_ms = sum(v["ms_per_step"] for k, v in kern.items() if k in sampler_names or k.startswith("op_"))
    if in_situ:
        lib_ms, own_ms, src = in_situ["library_ms"], in_situ["ours_ms"], "in situ (CUPTI, one graph-replayed step)"
        total_ms = in_situ["sum_ms"]
    else:
        own_ms, lib_ms, total_ms, src = ours_ms, max(step_ms - ours_ms, 1e-9), step_ms, "CUDA events per call (eager)"
    shares = {"own_kernels": own_ms / total_ms, "library_conv_gemm": lib_ms / total_ms,
              "sampler_kernels_proper": sampler_ms / step_ms, "source": src,
              "note": "own_kernels = every nhmc:: kernel (GroupNorm/softmax/resampling epilogues inside the U-Net included); "
                      "sampler_kernels_proper = leapfrog + DDIM epilogues + operator loss/adjoint (the non-U-Net time of a leapfrog step)"}
    unet_rep = {"algorithmic_tflop_per_step": unet_tflop, "library_ms_per_step": lib_ms,
                "achieved_tflops": unet_tflop / (lib_ms / 1e3), "peak_tflops": pk["bf16"],
                "tensor_pipe_frac": unet_tflop / (lib_ms / 1e3) / pk["bf16"],
                "note": "cuDNN/cuBLAS bf16 convolutions and GEMMs (PyTorch call); time = " + src + "; peak = sustained bf16 of MEASURED_PEAKS.json"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 sampler state / %s score network" % args.dtype, "data": "synthetic",
        "config": {"workload": ("FFHQ-256 %s hmc_latent, %d chains per GPU, timesteps=3 tau=1.0 epsilon=0.05 sigma_0=0.05 sigma_y=0.5, "
                                "random-init ffhq_latent (VQ-f4 + latent U-Net) architecture (BASELINE config[4])" if latent_wl else
                                "FFHQ-256 %s HMC, %d chains per GPU, timesteps=3 tau=1.0 epsilon=0.05 sigma_0=0.05, "
                                "random-init ffhq_10m U-Net (BASELINE config[1])") % (args.deg, B),
                   "chains_per_gpu": B, "deg": args.deg, "l2": "activations per step >> 126 MB L2 (8 chains x 3 U-Net passes); no flush needed",
                   "parallelism": "chains sharded, dp%d, no collective in the loop" % world, "cuda_graphs": use_graphs,
                   "lanes": smp.n_lanes, "pdl": bool(args.pdl)},
        "e2e": e2e,
        # kernels of libnhmc.so launched inside the timed region: exact CUPTI count of one replayed step x K x ranks when the
        # in-situ pass ran, else the python-side count of the C calls (which counts a two-kernel GroupNorm call once)
        "gpu_launches": int(in_situ["ours_launches_per_step"] * K * world) if in_situ else int(launches), "clocks": clk,
        "roofline": roofline, "kernels": kern, "kernels_by_shape": by_shape, "in_situ": in_situ,
        "shares": shares,
        "unet": unet_rep,
        "finite": finite,
        "fp32": fp32_line,
    }
    if not args.no_cpu_baseline and world >= 1:
        out["cpu_baseline"] = cpu_baseline(args.deg, steps_timed=1, warm=1)
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def fp32_companion(args, dev, op, y, x_orig, x0, steps, smp_bf16, sigma_0, B, K):
    """The bf16 line's companion: the same step with an fp32 score network, TF32 off -- the reference's own
    precision (config_ffhq.yml: use_fp16 False) -- timed the same way, plus how far one gradient evaluation of the bf16
    network is from the fp32 one on the same state (the full-size parity tests assert these bounds: tests/test_fullsize_gpu.py)."""
    from nhmc_b200 import _lib, kernels, sampler, unet
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(1234)                                           # same initialisation as the bf16 model, kept in fp32
        m32 = unet.create_model(**FFHQ, dtype=torch.float32)
        unet.randomize_zero_tensors(m32, seed=1234)
        m32.prepare(dev)
        ev32 = sampler.GradEvaluator(m32, steps, op, y, x0.to(dev))
        s32 = sampler.HMCSampler(x0.to(dev), ev32, 1.0, 0.05, sigma_0, sched=_lib.SCHED_PIXEL)
        kernels.begin_iteration(s32.chains, _lib.SCHED_PIXEL)
        s32.p.copy_(smp_bf16.p)
        # one gradient evaluation of both networks at the same state
        xs = smp_bf16.x_prop.clone()
        l16, ga, gb, xc16 = smp_bf16.ev(xs)
        l16, xc16, g16 = l16.clone(), xc16.clone(), (ga + gb).clone()
        l32, ga, gb, xc32 = ev32(xs)
        g32 = ga + gb
        xo = x_orig.to(dev)
        psnr = lambda a: 10 * torch.log10(1.0 / (((a + 1) / 2).clamp(0, 1) - ((xo + 1) / 2).clamp(0, 1)).pow(2).mean(dim=(1, 2, 3)))
        cmp = {"psnr_diff_db_max": float((psnr(xc16) - psnr(xc32)).abs().max()), "x_hat_rms_diff": float((xc16 - xc32).pow(2).mean().sqrt()),
               "loss_rel_diff_max": float(((l16 - l32).abs() / l32.abs()).max()),
               "grad_cosine": float((g16 * g32).sum() / (g16.norm() * g32.norm())),
               "note": "one gradient evaluation (3 U-Net forwards + VJPs) of the bf16 and the fp32 network at the same state; north star: PSNR "
                       "difference < 0.05 dB.  With random weights the map is chaotic (17x gain before a hard clip), so whole "
                       "trajectories agree statistically only (tests/test_fullsize_gpu.py)"}
        s32.x_prop.copy_(xs)
        s32.step(_lib.LF_FIRST)
        s32.capture_graphs()
        for _ in range(3):
            s32.step(_lib.LF_MIDDLE)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            s32.step(_lib.LF_MIDDLE)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        return {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": K, "dtype": "f32 sampler state / fp32 score network (TF32 off)",
                "bf16_vs_fp32": cmp}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference path, timed on the host cores
# ----------------------------------------------------------------------------------------------------
def _cpu_step_fn(deg):
    """One leapfrog step of ONE chain at FFHQ-256 with the full ffhq_10m-architecture U-Net on CPU (fp32)."""
    from oracle import hmc as ohmc, operators as oops, schedule as osch, unet as ounet
    cfg = ounet.resolve_config(**FFHQ)
    sd = ounet.seeded_state_dict(cfg, seed=1234)
    model = lambda xx, tt: ounet.unet_forward(sd, cfg, xx, tt)
    if deg.startswith("sr"):
        op = oops.SuperResolution(3, 256, int(deg[2:]))
    elif "inp" in deg and "box" in deg:
        op = oops.Inpainting.from_missing_indices(3, 256, oops.box_missing_indices(3, 256, 64, 64).numpy())
    elif "inp" in deg:
        torch.manual_seed(5678)
        op = oops.Inpainting(3, 256, oops.random_inpaint_missing_pixels(256))
    elif "deblur_gauss" in deg:
        op = oops.Deblurring.from_kernel(oops.gaussian_taps(10.0, half=2), 3, 256)
    elif deg == "color":
        op = oops.Colorization(256)
    elif "hdr" in deg:
        op = oops.HDR()
    elif deg == "deblur_aniso":
        op = oops.Deblurring2D.from_kernels(oops.gaussian_taps(1.0), oops.gaussian_taps(20.0), 3, 256)
    else:
        op = oops.PhaseRetrieval(2.0)
    gen = torch.Generator().manual_seed(5678)
    y = op.H(torch.rand(1, 3, 256, 256, generator=gen) * 2 - 1)
    y = y + 0.1 * torch.randn(y.shape, generator=gen)
    x = torch.randn(1, 3, 256, 256, generator=gen); p = torch.randn(1, 3, 256, 256, generator=gen)
    tab = osch.alpha_bar_table(osch.beta_schedule("linear", 1e-4, 0.02, 1000))
    eps_net = ohmc.module_eps_net(model)
    state = {"x": x, "p": p}
    sigma_y, eps = 0.1 + 1.6, 0.05

    def step():
        state["x"] = ohmc.drift(state["x"], state["p"], eps, 1.0)
        _, _, g = ohmc.grad_eval(state["x"], eps_net, tab, SEQ, SEQ_NEXT, op, y)
        state["p"] = ohmc.kick(state["p"], state["x"], g, sigma_y, eps)
    return step


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU arm is to use every host core it may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if n > torch.get_num_threads():
        torch.set_num_threads(n)


def cpu_baseline(deg, steps_timed=1, warm=1):
    _use_all_host_threads()
    step = _cpu_step_fn(deg)
    cores = torch.get_num_threads()
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps_timed):
        step()
    dt = time.perf_counter() - t0
    return {"value": steps_timed / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "1 chain x %d leapfrog step(s) of the same workload (FFHQ-256 %s, full ffhq_10m U-Net, fp32) after %d warm-up; "
                      "oracle restatement of main_sampling.py:706-713 on torch CPU" % (steps_timed, deg, warm),
            "ms_per_step": dt / steps_timed * 1e3}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is Python and
    cannot travel to the GPU box), all host threads, bounded sample per step."""
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    _use_all_host_threads()
    step = _cpu_step_fn(args.deg)
    cores = torch.get_num_threads()
    budget_s = 240.0
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    W_run = max(0, min(W, 1) - 1)            # the probe above is the warm-up
    for _ in range(W_run):
        step()
    K_run = max(1, min(K, int(max(1.0, (budget_s - first * (1 + W_run)) / max(first, 1e-3)))))
    t0 = time.perf_counter()
    for _ in range(K_run):
        step()
    dt = time.perf_counter() - t0
    v = K_run / dt
    out = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": world, "steps": K_run, "warmup": 1 + W_run,
           "ms_per_step": dt / K_run * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "FFHQ-256 %s HMC, timesteps=3 tau=1.0 epsilon=0.05 sigma_0=0.05, random-init ffhq_10m U-Net "
                                  "(BASELINE config[1]); bounded sample: 1 chain x 1 leapfrog step per step" % args.deg,
                      "deg": args.deg},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "1 chain x 1 leapfrog step per step, %d timed steps (requested %d, capped by a %.0f s budget)" % (K_run, K, budget_s)},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (sm_100a); there is no CPU fallback")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
