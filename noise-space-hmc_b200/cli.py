"""Command-line driver with the reference's flag surface (main_sampling.py:923-1015 `get_parser`,
:261-351 `prepare_measurement`, :354-570 `sample_image`) for the `--algo hmc` path.

    python main_sampling.py --ni --dataset ffhq --doc ffhq --algo hmc --timesteps 3 --deg inpaint_random \\
        --sigma_0 0.05 -i exp/samples/ffhq/demo --tau 1.0 --epsilon 0.05 [--chains 8] [--synthetic]

Unknown flags are ignored like in the reference (`parse_known_args`, main_sampling.py:1031).  Additions:
`--chains` (independent chains per GPU), `--synthetic` (seeded synthetic ground truth instead of the FFHQ folder),
`--dtype`, `--cuda_graphs`, `--reuse_gradient`, `--lanes`, `--save_progress`.  Under torchrun, chains are sharded across the GPUs (dist.py).
"""
import argparse
import math
import os
import random
import types

import numpy as np
import torch
import yaml

from . import dist as ndist, operators as P, sampler, schedule, unet


def get_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--seed", type=int, default=5678, help="Random seed")
    p.add_argument("--exp", type=str, default="exp", help="Path for saving running related data.")
    p.add_argument("--dataset", type=str, nargs="?", default="celeba", help="dataset")
    p.add_argument("--default_lr", action="store_true")
    p.add_argument("--lr", type=float, nargs="?", default=1.0, help="Step-Size")
    p.add_argument("--N", type=int, default=1, help="N repeats")
    p.add_argument("--deg", type=str, required=True, help="Degradation")
    p.add_argument("--sigma_0", type=float, required=True, help="Sigma_0")
    p.add_argument("--tau", type=float, default=1.0, help="Tau for HMC")
    p.add_argument("--epsilon", type=float, default=0.05, help="Epsilon for HMC")
    p.add_argument("--sigma_y", type=float, default=0.5, help="sigma_y for HMC (measurement noise)")
    p.add_argument("--m", type=float, default=1.0, help="Mass Matrix Variance")
    p.add_argument("--annealed_temp", action="store_true", default=False)
    p.add_argument("--noise", type=str, default="ddpm", help="Type of Noise")
    p.add_argument("--num_timesteps", type=int, nargs="?", default=1000, help="Maximum timestep for beta schedule")
    p.add_argument("--timesteps", type=int, nargs="?", default=10, help="Number of timesteps for actual sampling")
    p.add_argument("--subset_start", type=int, default=-1)
    p.add_argument("--subset_end", type=int, default=-1)
    p.add_argument("--algo", type=str, nargs="?", default="resample", help="Algorithm to use for sampling")
    p.add_argument("--refine", action="store_true")
    p.add_argument("-i", "--image_folder", type=str, default="exp/samples/ffhq/00000", help="The folder name of samples")
    # additions of this implementation
    p.add_argument("--chains", type=int, default=1, help="independent HMC chains (batched on the GPU, sharded across GPUs)")
    p.add_argument("--synthetic", action="store_true", help="seeded synthetic ground truth instead of an image folder")
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"], help="score-network compute dtype")
    p.add_argument("--cuda_graphs", action="store_true", help="replay the leapfrog step as CUDA graphs")
    p.add_argument("--reuse_gradient", action="store_true", help="keep (loss, gradient) of the current state instead of re-evaluating it at the top of each iteration (20 instead of 21 evaluations; main_sampling.py:693-695)")
    p.add_argument("--lanes", type=int, default=1, help="pixel path: evaluate the chains in this many groups on parallel streams (default 1: measured fastest at 8 chains)")
    p.add_argument("--save_progress", action="store_true", help="PNG of every accepted epoch like the reference (main_sampling.py:735)")
    p.add_argument("--ldm_ckpt", type=str, default="", help="hmc_latent: LatentDiffusion checkpoint (model.ckpt); random init if absent")
    p.add_argument("--max_iters", type=int, default=None, help="stop after this many HMC iterations (smoke runs)")
    return p


def dict2namespace(config):
    ns = argparse.Namespace()
    for k, v in config.items():
        setattr(ns, k, dict2namespace(v) if isinstance(v, dict) else v)
    return ns


def prepare_measurement(opt, config, device):
    """Operator for --deg (main_sampling.py:261-351); doubles opt.sigma_0 like the reference (:348)."""
    deg, C, D = opt.deg, config.data.channels, config.data.image_size
    if deg.startswith("sr_bicubic"):
        factor = int(deg[10:])
        H = P.SRConv(P.bicubic_taps(factor), C, D, device, stride=factor)            # main_sampling.py:265-284
    elif "sr" in deg:
        H = P.SuperResolution(C, D, int(deg[2:]), device)
    elif "inp" in deg and "box" in deg:
        H = P.Inpainting.box(C, D, device)                         # two random.randint(16, 112) draws (:292-293)
    elif "inp" in deg:
        H = P.Inpainting.random(C, D, device)                      # torch.randperm on the global CPU generator (:302)
    elif "deblur_gauss" in deg:
        H = P.Deblurring(P.gaussian_taps(10.0, half=2), C, D, device)        # 5 taps, sigma 10 (:307-312)
    elif "phase" in deg:
        H = P.PhaseRetrievalOperator(2.0, device, C, D)
    elif "hdr" in deg:
        H = P.HDR(C, D, device)
    elif deg == "deblur_aniso":
        H = P.Deblurring2D(P.gaussian_taps(1.0), P.gaussian_taps(20.0), C, D, device)
    elif deg == "color":
        H = P.Colorization(D, device)
    elif deg.startswith("cs"):
        H = P.WalshHadamardCS.random(C, D, int(deg[2:]), device)                     # main_sampling.py:322-325
    elif deg in ("deno", "denoise", "denoising"):                  # Hfuncs.Denoising; the reference's driver has no --deg for it
        H = P.Denoising(C, D, device)
    else:
        raise SystemExit("ERROR: degradation type %r is not on the accelerated hmc path "
                         "(supported: inpaint_random, inpaint_box, sr<r>, sr_bicubic<r>, cs<r>, deblur_aniso, deblur_gauss, hdr, phase_retrieval, color, deno)" % deg)
    opt.sigma_0 = 2 * opt.sigma_0
    return H, opt.sigma_0, deg


def init_model(opt, config, model_config, device):
    if config.model_type != "ffhq":
        raise NotImplementedError("only model_type 'ffhq' is on the accelerated path")
    dtype = torch.bfloat16 if getattr(opt, "dtype", "bf16") == "bf16" else torch.float32
    model = unet.create_model(**model_config, dtype=dtype)
    unet.randomize_zero_tensors(model)          # no-op for a loaded checkpoint; keeps random-init runs non-degenerate
    return model.prepare(device)


def _load_images(opt, config, n, generator):
    D = config.data.image_size
    folder = os.path.join(opt.exp, "datasets", "ffhq", "samples_1k")
    if opt.synthetic or not os.path.isdir(folder):
        return torch.rand(n, config.data.channels, D, D, generator=generator)
    from PIL import Image
    files = sorted(f for f in os.listdir(folder) if f.lower().endswith((".png", ".jpg")))
    idx = torch.randperm(len(files), generator=generator)[:n].tolist()
    ims = [torch.from_numpy(np.asarray(Image.open(os.path.join(folder, files[i])).convert("RGB").resize((D, D)))).permute(2, 0, 1).float() / 255 for i in idx]
    return torch.stack(ims)


def _sample_latent(opt, config, model_config, device, rank, world, a0, a1, H_funcs, sigma_0, x_orig, y_0, cgen, mom, uni, on_it):
    """--algo hmc_latent (main_sampling_latent.py:354-570): chains live in the 3 x 64 x 64 latent space of the
    `ffhq_latent` model; the collected latents are decoded with decode_first_stage (:473)."""
    from . import latent
    n = a1 - a0
    dtype = torch.bfloat16 if getattr(opt, "dtype", "bf16") == "bf16" else torch.float32
    ldm = latent.create_latent_model(model_config or None, dtype=dtype, ckpt_path=getattr(opt, "ldm_ckpt", ""))
    from . import unet as _unet
    _unet.randomize_zero_tensors(ldm)          # random-init runs: zero-initialised output convs would make eps == 0
    ldm.prepare(device)
    hz = config.data.image_size // 4
    z = torch.stack([torch.randn(3, hz, hz, device=device, generator=g_) for g_ in cgen]) if n else torch.zeros(0, 3, hz, hz, device=device)
    seq, seq_next = schedule.timestep_sequence(opt.num_timesteps, opt.timesteps)
    algo = types.SimpleNamespace(model=ldm, H_funcs=H_funcs, sigma_0=sigma_0, cls_fn=None)
    out = None
    if n:
        lat = sampler.hmc_latent(z, n, seq, seq_next, algo, opt, y_0, H_funcs, x_orig, momentum=mom, uniform=uni,
                                 max_iters=opt.max_iters, on_iteration=on_it, image_size=config.data.image_size,
                                 reuse_gradient=opt.reuse_gradient)
        lat = [lat] if n == 1 else lat
        out = [ldm.decode_first_stage(l_).float() if l_.numel() else torch.zeros(0, 3, config.data.image_size, config.data.image_size, device=device)
               for l_ in lat]
        S = min(o.shape[0] for o in out)
        out = torch.stack([o[:S] for o in out]) if S else None
    # every rank enters the gather, also one that owns no chain or collected nothing (ranks agree on S inside)
    allc = ndist.gather_samples(out, opt.chains, (3, config.data.image_size, config.data.image_size), device) if world > 1 else out
    if rank == 0 and allc is not None and allc.numel():
        img = ((allc + 1) / 2).clamp(0, 1)
        orig = ((x_orig[0] + 1) / 2).clamp(0, 1)
        psnr = 10 * torch.log10(1 / ((img - orig) ** 2).mean(dim=(2, 3, 4)))
        print("Total Average PSNR: %.3f" % float(psnr.mean()))
        print("Number of samples: %d chains x %d" % (img.shape[0], img.shape[1]))
        torch.save(allc.cpu(), os.path.join(opt.image_folder, "hmc_latent_samples.pt"))
    return allc


def sample_image(opt, config, model_config, device):
    rank, world, local_rank = ndist.init_from_env()
    H_funcs, sigma_0, deg = prepare_measurement(opt, config, device)
    latent_mode = opt.algo == "hmc_latent"
    model = None if latent_mode else init_model(opt, config, model_config, device)
    betas = schedule.get_beta_schedule(beta_schedule=config.diffusion.beta_schedule, beta_start=config.diffusion.beta_start,
                                       beta_end=config.diffusion.beta_end,
                                       num_diffusion_timesteps=config.diffusion.num_diffusion_timesteps)
    b = torch.from_numpy(betas).float().to(device)
    algo = types.SimpleNamespace(model=model, H_funcs=H_funcs, sigma_0=sigma_0, cls_fn=None)
    a0, a1 = ndist.chain_shard(opt.chains, rank, world)
    n = a1 - a0
    gen = torch.Generator().manual_seed(opt.seed)
    x_orig = _load_images(opt, config, 1, gen).to(device) * 2 - 1          # data_transform: 2x - 1 (datasets/__init__.py:197)
    x_orig = x_orig.expand(max(n, 1), -1, -1, -1).contiguous()
    y_0 = H_funcs.H(x_orig)
    y_0 = y_0 + sigma_0 * torch.randn_like(y_0)
    cgen = [torch.Generator(device=device).manual_seed(ndist.chain_seed(opt.seed, c)) for c in range(a0, a1)]
    x = torch.stack([torch.randn(config.data.channels, config.data.image_size, config.data.image_size, device=device, generator=g_) for g_ in cgen]) \
        if n else torch.zeros(0, config.data.channels, config.data.image_size, config.data.image_size, device=device)
    seq, seq_next = schedule.timestep_sequence(opt.num_timesteps, opt.timesteps)
    os.makedirs(opt.image_folder, exist_ok=True)
    mom = lambda x_: torch.stack([torch.randn(x_.shape[1:], device=x_.device, generator=g_) for g_ in cgen]) * math.sqrt(opt.m)
    ugen = [torch.Generator().manual_seed(ndist.chain_seed(opt.seed, c) + 10 ** 6) for c in range(a0, a1)]
    uni = lambda B: torch.cat([torch.rand(1, generator=g_) for g_ in ugen])

    def on_it(it, st, xc):
        if rank == 0:
            print("iter %d  epochs %s  accepted %s  eps %s" % (it, st["epoch"].tolist(), st["accepted"].tolist(),
                                                               np.round(st["epsilon"], 4).tolist()), flush=True)
        if opt.save_progress and st["accept"].any():
            import torchvision.utils as tvu
            for j in np.nonzero(st["accept"])[0]:
                tvu.save_image(((xc[j] + 1) / 2).clamp(0, 1), os.path.join(opt.image_folder, "hmc_c%d_%d.png" % (a0 + j, st["epoch"][j])))

    if latent_mode:
        return _sample_latent(opt, config, model_config, device, rank, world, a0, a1, H_funcs, sigma_0, x_orig, y_0, cgen, mom, uni, on_it)
    out = None
    if n:
        out = sampler.hmc(x, n, b, seq, seq_next, algo, opt, y_0, H_funcs, x_orig, momentum=mom, uniform=uni,
                          max_iters=opt.max_iters, on_iteration=on_it, cuda_graphs=opt.cuda_graphs,
                          reuse_gradient=opt.reuse_gradient, lanes=opt.lanes)
        out = out[None] if n == 1 else out                                  # [n, S, C, H, W]
    allc = ndist.gather_samples(out, opt.chains, (config.data.channels, config.data.image_size, config.data.image_size), device) if world > 1 else out
    if rank == 0 and allc is not None and allc.numel():
        img = ((allc + 1) / 2).clamp(0, 1)                                  # inverse_data_transform (datasets/__init__.py:214)
        orig = ((x_orig[0] + 1) / 2).clamp(0, 1)
        mse = ((img - orig) ** 2).mean(dim=(2, 3, 4))
        psnr = 10 * torch.log10(1 / mse)
        print("Total Average PSNR: %.3f (%.4f)" % (float(psnr.mean()), float(psnr.std(dim=1).mean()) if psnr.shape[1] > 1 else 0.0))
        print("Number of samples: %d chains x %d" % (img.shape[0], img.shape[1]))
        torch.save(allc.cpu(), os.path.join(opt.image_folder, "hmc_samples.pt"))
    return allc


def main(argv=None, latent=False):
    parser = get_parser()
    if latent:                                   # main_sampling_latent.py:828-830: epsilon defaults to 0.1 there
        parser.set_defaults(epsilon=0.1, dataset="ffhq_latent")
    opt, _unknown = parser.parse_known_args(argv)
    if opt.algo not in ("hmc", "hmc_latent") or (opt.algo == "hmc_latent") != bool(latent):
        raise SystemExit("main_sampling.py implements --algo hmc, main_sampling_latent.py --algo hmc_latent (got %r)" % opt.algo)
    if not torch.cuda.is_available():
        raise SystemExit("a CUDA device (sm_100a) is required: there is no CPU fallback")
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    opt.config = os.path.join(here, "configs", "config_{}.yml".format(opt.dataset))
    with open(opt.config) as f:
        cfg = yaml.safe_load(f)
    model_config = cfg.get("model_params") if latent else cfg["model"]
    config = dict2namespace(cfg)
    torch.manual_seed(opt.seed); np.random.seed(opt.seed); random.seed(opt.seed)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    return sample_image(opt, config, model_config, torch.device("cuda", local_rank))


if __name__ == "__main__":
    main()
