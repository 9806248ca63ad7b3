"""Command-line driver with the reference's flag surface (main_sampling.py:923-1015 `get_parser`,
:261-351 `prepare_measurement`, :354-570 `sample_image`) for the `--algo hmc` path.

    python main_sampling.py --ni --dataset ffhq --doc ffhq --algo hmc --timesteps 3 --deg inpaint_random \\
        --sigma_0 0.05 -i exp/samples/ffhq/demo --tau 1.0 --epsilon 0.05 [--chains 8] [--synthetic]

Unknown flags are ignored like in the reference (`parse_known_args`, main_sampling.py:1031).  Additions:
`--chains` (independent chains per GPU), `--synthetic` (seeded synthetic ground truth instead of the FFHQ folder),
`--dtype`, `--no_cuda_graphs`, `--reuse_gradient`, `--lanes`, `--save_progress`, `--n_images`, `--batch`, `--adaptive_mass`.
Under torchrun the (image, chain) work items are sharded across the GPUs (dist.py); images leave through a side stream +
worker thread (metrics.py), never through the sampler's stream.
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
    p.add_argument("--cuda_graphs", action="store_true", help="(default behaviour now: the leapfrog step is replayed as CUDA graphs; kept for old command lines)")
    p.add_argument("--no_cuda_graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    p.add_argument("--n_images", type=int, default=1, help="images to reconstruct (folder order / --subset_start..--subset_end like the reference, or synthetic)")
    p.add_argument("--batch", type=int, default=0, help="(image, chain) work items per batched sampler call on one GPU (default: --chains)")
    p.add_argument("--adaptive_mass", action="store_true", help="--algo hmc with the adaptive diagonal mass of hmc_test_conditioning (main_sampling.py:776-894)")
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


def _image_list(opt, config):
    """The images of a run: the sorted files of exp/datasets/ffhq/samples_1k (the reference's FFHQ loader,
    datasets/__init__.py) restricted to [--subset_start, --subset_end) like main_sampling.py:371-376, the first
    --n_images of them; or --n_images seeded synthetic images.  Returns a list of (label, loader(generator) -> [C,D,D] in [0,1])."""
    D, C = config.data.image_size, config.data.channels
    folder = os.path.join(opt.exp, "datasets", "ffhq", "samples_1k")
    n = max(1, int(opt.n_images))
    if opt.synthetic or not os.path.isdir(folder):
        return [("synthetic%d" % i, (lambda g, i=i: torch.rand(C, D, D, generator=torch.Generator().manual_seed(opt.seed + 104729 * i))))
                for i in range(n)]
    files = sorted(f for f in os.listdir(folder) if f.lower().endswith((".png", ".jpg")))
    if opt.subset_start >= 0 and opt.subset_end > 0:
        assert opt.subset_end > opt.subset_start
        files = files[opt.subset_start: opt.subset_end]
    else:                                    # the reference shuffles with a seeded generator (main_sampling.py:396-403)
        order = torch.randperm(len(files), generator=torch.Generator().manual_seed(opt.seed)).tolist()
        files = [files[i] for i in order]
    files = files[:n]

    def load(path):
        from PIL import Image
        return torch.from_numpy(np.asarray(Image.open(path).convert("RGB").resize((D, D))).copy()).permute(2, 0, 1).float() / 255
    return [(os.path.splitext(f)[0], (lambda g, f=f: load(os.path.join(folder, f)))) for f in files]


def sample_image(opt, config, model_config, device):
    """main_sampling.py:354-570 for --algo hmc / hmc_latent: every (image, chain) pair is one independent work item;
    the items are cut into contiguous shards over the ranks (one process per GPU, no collective in the loop) and run
    `--batch` at a time as one batched sampler call.  All chains of an image share its measurement y_0.  Images leave
    through an AsyncImageWriter (side stream + worker thread), metrics are computed on the device at the end."""
    from . import metrics
    rank, world, local_rank = ndist.init_from_env()
    H_funcs, sigma_0, deg = prepare_measurement(opt, config, device)
    latent_mode = opt.algo == "hmc_latent"
    C, D = config.data.channels, config.data.image_size
    seq, seq_next = schedule.timestep_sequence(opt.num_timesteps, opt.timesteps)
    os.makedirs(opt.image_folder, exist_ok=True)
    writer = metrics.AsyncImageWriter(device)
    if latent_mode:
        from . import latent
        dtype = torch.bfloat16 if getattr(opt, "dtype", "bf16") == "bf16" else torch.float32
        ldm = latent.create_latent_model(model_config or None, dtype=dtype, ckpt_path=getattr(opt, "ldm_ckpt", ""))
        unet.randomize_zero_tensors(ldm)          # random-init runs: zero-initialised output convs would make eps == 0
        ldm.prepare(device)
        algo = types.SimpleNamespace(model=ldm, H_funcs=H_funcs, sigma_0=sigma_0, cls_fn=None)
        b = None
    else:
        model = init_model(opt, config, model_config, device)
        betas = schedule.get_beta_schedule(beta_schedule=config.diffusion.beta_schedule, beta_start=config.diffusion.beta_start,
                                           beta_end=config.diffusion.beta_end,
                                           num_diffusion_timesteps=config.diffusion.num_diffusion_timesteps)
        b = torch.from_numpy(betas).float().to(device)
        algo = types.SimpleNamespace(model=model, H_funcs=H_funcs, sigma_0=sigma_0, cls_fn=None)

    images = _image_list(opt, config)
    items = [(i, c) for i in range(len(images)) for c in range(opt.chains)]            # (image, chain), image-major
    a0, a1 = ndist.chain_shard(len(items), rank, world)
    mine = items[a0:a1]
    # ground truth and measurement per IMAGE (the same on every rank that holds one of its chains)
    x_orig, y_img = {}, {}
    for i in sorted({i for i, _ in mine} | ({0} if rank == 0 else set())):
        gi = torch.Generator().manual_seed(opt.seed + 7919 * i)
        xo = (images[i][1](gi).to(device) * 2 - 1)[None]                 # data_transform: 2x - 1 (datasets/__init__.py:197)
        yo = H_funcs.H(xo)
        yo = yo + sigma_0 * torch.randn(yo.shape, generator=gi).to(device)            # main_sampling.py:447-448
        x_orig[i], y_img[i] = xo, yo
    if rank == 0:                                                        # y0_ / orig_ images like main_sampling.py:452-458
        for i in x_orig:
            if any(ii == i for ii, _ in mine) or i == 0:
                try:
                    ypi = H_funcs.H_pinv(y_img[i]).reshape(1, C, D, D)
                    writer.save(metrics.inverse_data_transform(ypi[0]), os.path.join(opt.image_folder, "y0_%s.png" % images[i][0]))
                except Exception:      # noqa: BLE001  (operators without a pseudo-inverse image)
                    pass
                writer.save(metrics.inverse_data_transform(x_orig[i][0]), os.path.join(opt.image_folder, "orig_%s.png" % images[i][0]))

    hz = D // 4
    fn = sampler.hmc_test_conditioning if getattr(opt, "adaptive_mass", False) else sampler.hmc
    results = []                                                         # per item: [S, C, D, D] on the device
    bs = max(1, int(opt.batch or opt.chains))
    for k0 in range(0, len(mine), bs):
        grp = mine[k0:k0 + bs]
        n = len(grp)
        gid = [a0 + k0 + j for j in range(n)]                            # global item index: seeds do not depend on the sharding
        cgen = [torch.Generator(device=device).manual_seed(ndist.chain_seed(opt.seed, g_)) for g_ in gid]
        ugen = [torch.Generator().manual_seed(ndist.chain_seed(opt.seed, g_) + 10 ** 6) for g_ in gid]
        shape = (3, hz, hz) if latent_mode else (C, D, D)
        x = torch.stack([torch.randn(shape, device=device, generator=g_) for g_ in cgen])
        y_0 = torch.cat([y_img[i] for i, _ in grp], 0)
        xo = torch.cat([x_orig[i] for i, _ in grp], 0)
        mom = lambda x_, cgen=cgen: torch.stack([torch.randn(x_.shape[1:], device=x_.device, generator=g_) for g_ in cgen]) * math.sqrt(opt.m)
        uni = lambda B, ugen=ugen: torch.cat([torch.rand(1, generator=g_) for g_ in ugen])

        def on_it(it, st, xc, grp=grp, xo=xo):
            if rank == 0:
                print("iter %d  epochs %s  accepted %s  eps %s" % (it, st["epoch"].tolist(), st["accepted"].tolist(),
                                                                   np.round(st["epsilon"], 4).tolist()), flush=True)
            if opt.save_progress and st["accept"].any():                 # main_sampling.py:735-742, off the critical path
                img = xc if not latent_mode else None
                if img is not None:
                    acc = np.nonzero(st["accept"])[0]
                    ps = metrics.psnr_per_sample(img, xo)                # device tensor; read only with the next state read
                    for j in acc:
                        i, c = grp[j]
                        writer.save(metrics.inverse_data_transform(img[j]),
                                    os.path.join(opt.image_folder, "hmc_%s_c%d_%d.png" % (images[i][0], c, st["epoch"][j])))
                    if rank == 0:
                        print("  PSNR of the accepted proposals:", np.round(ps.cpu().numpy()[acc], 3).tolist(), flush=True)

        if latent_mode:
            lat = sampler.hmc_latent(x, n, seq, seq_next, algo, opt, y_0, H_funcs, xo, momentum=mom, uniform=uni,
                                     max_iters=opt.max_iters, on_iteration=on_it, image_size=D, reuse_gradient=opt.reuse_gradient,
                                     cuda_graphs=False if opt.no_cuda_graphs else None)
            lat = [lat] if n == 1 else lat
            outs = [ldm.decode_first_stage(l_).float() if l_.numel() else torch.zeros(0, C, D, D, device=device) for l_ in lat]
        else:
            kw = dict(momentum=mom, uniform=uni, max_iters=opt.max_iters, on_iteration=on_it,
                      cuda_graphs=False if opt.no_cuda_graphs else None)
            if fn is sampler.hmc:
                kw.update(reuse_gradient=opt.reuse_gradient, lanes=opt.lanes)
            smp = fn(x, n, b, seq, seq_next, algo, opt, y_0, H_funcs, xo, return_sampler=True, **kw)
            outs = smp.collected()
        results += list(outs)

    # ---- final gather (the only collective), metrics, images ----
    S = min((o.shape[0] for o in results), default=0)
    local = torch.stack([o[:S] for o in results]) if results and S > 0 else None
    allc = ndist.gather_samples(local, len(items), (C, D, D), device) if world > 1 else local
    report = None
    if rank == 0 and allc is not None and allc.numel():
        report = {"images": [], "chains": opt.chains, "samples_per_chain": int(allc.shape[1])}
        tot, tot_std = [], []
        for i, (label, _) in enumerate(images):
            if i not in x_orig:
                gi = torch.Generator().manual_seed(opt.seed + 7919 * i)
                x_orig[i] = (images[i][1](gi).to(device) * 2 - 1)[None]
            smp_i = allc[i * opt.chains:(i + 1) * opt.chains].reshape(-1, C, D, D)     # all samples of all chains of image i
            sm = metrics.summarize(smp_i, x_orig[i][0])
            tot.append(sm["psnr_mean"]); tot_std.append(sm["psnr_std"])
            writer.save(metrics.inverse_data_transform(smp_i[-1]), os.path.join(opt.image_folder, "%s.png" % label))   # :512-515
            if smp_i.shape[0] > 1:                                                                                      # :494-508
                norm_map, raw = metrics.std_dev_map(smp_i)
                writer.save(metrics.hot_colormap(norm_map), os.path.join(opt.image_folder, "std_dev_map_%s.png" % label))
                np.save(os.path.join(opt.image_folder, "std_dev_map_%s.npy" % label), raw.cpu().numpy())
            report["images"].append({"image": label, "psnr_mean": sm["psnr_mean"], "psnr_std": sm["psnr_std"], "n_samples": int(smp_i.shape[0])})
        print("Total Average PSNR: %.3f (%.4f)" % (float(np.mean(tot)), float(np.mean(tot_std))))
        print("Total Average SSIM: n/a   Total Average LPIPS: n/a   (skimage / lpips are not vendored with the reference)")
        print("Number of samples: %d" % len(images))
        torch.save(allc.cpu(), os.path.join(opt.image_folder, "hmc_latent_samples.pt" if latent_mode else "hmc_samples.pt"))
        import json
        with open(os.path.join(opt.image_folder, "metrics.json"), "w") as f:
            json.dump(report, f, indent=1)
    n_written = writer.close()
    if rank == 0:
        print("images written: %d" % n_written)
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
