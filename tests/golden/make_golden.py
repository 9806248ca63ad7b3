#!/usr/bin/env python
"""Record golden vectors by running the REFERENCE's own code (build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Imports /root/reference (read-only) through oracle/ref_harness.py and records inputs and
outputs of the functions on the `--algo hmc` path.  The reference ships no tests or golden
vectors, so these files are the pin for oracle/ (tests/test_oracle_golden.py) and, through
the oracle, for the CUDA path.  Nothing here runs on the GPU box.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from oracle.tinynet import TinyEpsNet  # noqa: E402
from oracle import unet as ounet  # noqa: E402

ms = ref_harness.import_reference()
from obs_functions import Hfuncs  # noqa: E402
from algos.unconditional import Unconditional  # noqa: E402
from guided_diffusion import unet_ffhq  # noqa: E402

torch.set_num_threads(8)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print("wrote", name, {k: tuple(v.shape) for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")


# ---------------------------------------------------------------------------- schedule
def gold_schedule():
    betas = ms.get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
    b = torch.from_numpy(betas).float()
    t_all = torch.arange(-1, 1000)
    tab = ms.compute_alpha(b, t_all).reshape(-1)
    skip = 1000 // 4
    seq = list(range(skip, 1000, skip))
    seq_next = [-1] + seq[:-1]
    save("schedule.npz", betas=betas, alpha_table=tab, seq=np.array(seq), seq_next=np.array(seq_next))


# ---------------------------------------------------------------------------- inpaint mask
def gold_inpaint_mask():
    torch.manual_seed(5678)
    perm = torch.randperm(256 ** 2)
    miss = perm[: int(256 ** 2 * 0.92)]
    keep = np.ones(256 ** 2, dtype=bool)
    keep[miss.numpy()] = False
    kept = np.nonzero(keep)[0].astype(np.int64)
    sha = hashlib.sha1(kept.tobytes()).hexdigest()
    # small case through the reference's own (slow, python-loop) constructor
    torch.manual_seed(99)
    D = 32
    missing_r = 3 * torch.randperm(D ** 2)[: int(D ** 2 * 0.92)].long()
    missing = torch.cat([missing_r, missing_r + 1, missing_r + 2], dim=0)
    H = Hfuncs.Inpainting(3, D, missing, "cpu")
    save("inpaint_mask.npz", perm_head=perm[:8], kept_head=kept[:10], kept_count=np.array(len(kept)),
         sha1=np.frombuffer(bytes.fromhex(sha), dtype=np.uint8), small_missing=missing,
         small_kept_indices=H.kept_indices)


# ---------------------------------------------------------------------------- operators
def _op_record(prefix, H, x, rec, y_shape_image=False):
    """Record H(x0), y = H(x0)+noise, grad of sum((y-H(clip(x)))^2) at x (autograd), H_pinv(y)."""
    g = torch.Generator().manual_seed(7)
    x_true = torch.rand(x.shape, generator=g) * 2 - 1
    y = H.H(x_true).detach()
    y = y + 0.1 * torch.randn(y.shape, generator=g)
    xv = x.clone().requires_grad_(True)
    loss = torch.sum((y - H.H(xv.clip(-1, 1))) ** 2)
    (grad,) = torch.autograd.grad(loss, xv)
    rec[prefix + "_x"] = x
    rec[prefix + "_y"] = y
    rec[prefix + "_Hx"] = H.H(x.clip(-1, 1)).detach()
    rec[prefix + "_loss"] = loss.detach()
    rec[prefix + "_grad"] = grad
    try:
        rec[prefix + "_pinv"] = H.H_pinv(y.clone()).detach()
    except Exception as e:  # pragma: no cover
        print("pinv failed for", prefix, e)


def gold_operators():
    rec = {}
    D, B = 32, 2
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, D, D, generator=g) * 0.8      # some values beyond +-1 to exercise the clip

    torch.manual_seed(99)
    missing_r = 3 * torch.randperm(D ** 2)[: int(D ** 2 * 0.92)].long()
    missing = torch.cat([missing_r, missing_r + 1, missing_r + 2], dim=0)
    rec["inp_missing"] = missing
    _op_record("inp", Hfuncs.Inpainting(3, D, missing, "cpu"), x, rec)

    _op_record("sr4", Hfuncs.SuperResolution(3, D, 4, "cpu"), x, rec)
    _op_record("sr16", Hfuncs.SuperResolution(3, D, 16, "cpu"), x, rec)
    _op_record("sr2", Hfuncs.SuperResolution(3, D, 2, "cpu"), x, rec)
    _op_record("hdr", Hfuncs.HDR(), x, rec)

    # anisotropic deblur: kernels exactly as main_sampling.py:327-335
    sigma = 20
    pdf = lambda v: torch.exp(torch.Tensor([-0.5 * (v / sigma) ** 2]))
    kernel2 = torch.Tensor([pdf(i) for i in range(-4, 5)])
    sigma = 1
    pdf = lambda v: torch.exp(torch.Tensor([-0.5 * (v / sigma) ** 2]))
    kernel1 = torch.Tensor([pdf(i) for i in range(-4, 5)])
    Hd = Hfuncs.Deblurring2D(kernel1 / kernel1.sum(), kernel2 / kernel2.sum(), 3, D, "cpu")
    for k in ("U_small1", "V_small1", "U_small2", "V_small2", "_singulars", "_perm"):
        rec["deblur_tab_" + k.strip("_")] = getattr(Hd, k)
    rec["deblur_kernel1"] = kernel1 / kernel1.sum()
    rec["deblur_kernel2"] = kernel2 / kernel2.sum()
    _op_record("deblur", Hd, x, rec)

    Hp = Hfuncs.PhaseRetrievalOperator(oversample=2.0, device="cpu")
    _op_record("phase", Hp, x, rec)
    save("operators.npz", **rec)

    # 256-size deblur tables statistics: counts only, tables are too big to commit
    Hd256 = Hfuncs.Deblurring2D(kernel1 / kernel1.sum(), kernel2 / kernel2.sum(), 3, 256, "cpu")
    nz1 = int((Hd256.singulars_small1 != 0).sum()); nz2 = int((Hd256.singulars_small2 != 0).sum())
    save("deblur256_stats.npz", nz1=np.array(nz1), nz2=np.array(nz2), nnz=np.array(int((Hd256._singulars != 0).sum())),
         s_head=Hd256._singulars[:16])


# ---------------------------------------------------------------------------- operators, next rows
def gold_operators_ext():
    """deblur_gauss, color, deno and box inpainting, recorded from the reference's own Hfuncs classes."""
    rec = {}
    D, B = 32, 2
    g = torch.Generator().manual_seed(12)
    x = torch.randn(B, 3, D, D, generator=g) * 0.8

    # Gaussian deblurring: kernel exactly as main_sampling.py:307-312
    sigma = 10
    pdf = lambda v: torch.exp(torch.Tensor([-0.5 * (v / sigma) ** 2]))
    kernel = torch.Tensor([pdf(-2), pdf(-1), pdf(0), pdf(1), pdf(2)])
    Hg = Hfuncs.Deblurring(kernel / kernel.sum(), 3, D, "cpu")
    for k in ("U_small", "V_small", "_singulars", "_perm"):
        rec["dgauss_tab_" + k.strip("_")] = getattr(Hg, k)
    rec["dgauss_kernel"] = kernel / kernel.sum()
    _op_record("dgauss", Hg, x, rec)

    Hc = Hfuncs.Colorization(D, "cpu")
    for k in ("U_small", "singulars_small", "V_small"):
        rec["color_tab_" + k] = getattr(Hc, k)
    _op_record("color", Hc, x, rec)

    _op_record("deno", Hfuncs.Denoising(3, D, "cpu"), x, rec)

    # bicubic SRConv: kernel exactly as main_sampling.py:266-284 (factor 4)
    factor = 4
    def bicubic_kernel(x_, a=-0.5):
        if abs(x_) <= 1:
            return (a + 2) * abs(x_) ** 3 - (a + 3) * abs(x_) ** 2 + 1
        elif 1 < abs(x_) and abs(x_) < 2:
            return a * abs(x_) ** 3 - 5 * a * abs(x_) ** 2 + 8 * a * abs(x_) - 4 * a
        else:
            return 0
    kk = np.zeros((factor * 4))
    for i in range(factor * 4):
        x_ = (1 / factor) * (i - np.floor(factor * 4 / 2) + 0.5)
        kk[i] = bicubic_kernel(x_)
    kk = kk / np.sum(kk)
    kern = torch.from_numpy(kk).float()
    Hb = Hfuncs.SRConv(kern / kern.sum(), 3, D, "cpu", stride=factor)
    for k in ("U_small", "singulars_small", "V_small"):
        rec["srbic_tab_" + k] = getattr(Hb, k)
    rec["srbic_kernel"] = kern / kern.sum()
    _op_record("srbic", Hb, x, rec)

    # Walsh-Hadamard compressive sensing (main_sampling.py:322-325), ratio 4
    pg = torch.Generator().manual_seed(31)
    perm = torch.randperm(D ** 2, generator=pg)
    rec["cs_perm"] = perm
    _op_record("cs", Hfuncs.WalshHadamardCS(3, D, 4, perm, "cpu"), x, rec)

    # box inpainting: the index construction of main_sampling.py:290-299 at D = 32 with a 16 x 16 box
    left, up, size = 5, 9, 16
    missing = torch.zeros([D, D, 3])
    missing[left:left + size, up:up + size, :] = 1.0
    missing = torch.nonzero(missing.view(-1).long()).squeeze()
    rec["box_missing"] = missing
    rec["box_geom"] = np.array([left, up, size])
    _op_record("box", Hfuncs.Inpainting(3, D, missing, "cpu"), x, rec)
    save("operators_ext.npz", **rec)


# ---------------------------------------------------------------------------- DDIM epilogue
def gold_ddim():
    net = TinyEpsNet()
    algo = Unconditional(net, None, 0.1)
    betas = ms.get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
    b = torch.from_numpy(betas).float()
    g = torch.Generator().manual_seed(21)
    B, D = 2, 32
    x = torch.randn(B, 3, D, D, generator=g)
    opt = types.SimpleNamespace(algo="hmc", noise="ddpm")
    seq, seq_next = [250, 500, 750], [-1, 250, 500]
    xv = x.clone().requires_grad_(True)
    xt = ms.iterative_sampling(xv, B, b, seq, seq_next, algo, opt, None, tqdm_disable=True)
    w = torch.randn(xt.shape, generator=g)
    (gx,) = torch.autograd.grad((xt * w).sum(), xv)
    # single step at t=750 as well
    t = torch.ones(B) * 750
    at = ms.compute_alpha(b, t.long()); at_next = ms.compute_alpha(b, (torch.ones(B) * 500).long())
    xv2 = x.clone().requires_grad_(True)
    x0_t, add_up = algo.cal_x0(xv2, t, at, at_next, None, "ddpm")
    xt1 = algo.map_back(x0_t, None, add_up, at_next, at)
    save("ddim.npz", x=x, w=w, x_hat=xt.detach(), grad_x=gx, step_eps=algo.et.detach(), step_x0=x0_t.detach(),
         step_xt=xt1.detach())


# ---------------------------------------------------------------------------- U-Net
def gold_unet():
    kw = dict(image_size=64, num_channels=32, num_res_blocks=1, channel_mult="", learn_sigma=True,
              class_cond=False, use_checkpoint=False, attention_resolutions=16, num_heads=4,
              num_head_channels=16, num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0,
              resblock_updown=True, use_fp16=False, use_new_attention_order=False, model_path="")
    model = unet_ffhq.create_model(**kw).eval()
    cfg = ounet.resolve_config(**kw)
    sd = ounet.seeded_state_dict(cfg, seed=1234)
    ref_sd = model.state_dict()
    assert set(sd) == set(ref_sd), (set(sd) ^ set(ref_sd))
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref_sd[k].shape), k
    assert list(ounet.param_shapes(cfg)) == list(ref_sd), "registration order differs"
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 64, 64, generator=g)
    t = torch.tensor([750.0, 250.0])
    xv = x.clone().requires_grad_(True)
    out = model(xv, t)
    w = torch.randn(out.shape, generator=g)
    (gx,) = torch.autograd.grad((out * w).sum(), xv)
    save("unet_small.npz", x=x, t=t, w=w, out=out.detach(), grad_x=gx)
    # the real ffhq_10m architecture: parameter count / names only
    import yaml
    with open(os.path.join(ref_harness.REF, "configs", "config_ffhq.yml")) as f:
        full = yaml.safe_load(f)["model"]
    cfg_full = ounet.resolve_config(**full)
    shapes = ounet.param_shapes(cfg_full)
    big = unet_ffhq.create_model(**full)
    bsd = big.state_dict()
    assert list(shapes) == list(bsd)
    for k in shapes:
        assert tuple(shapes[k]) == tuple(bsd[k].shape), k
    nparam = sum(int(np.prod(s)) for s in shapes.values())
    names_sha = hashlib.sha1("\n".join(f"{k}:{tuple(v)}" for k, v in shapes.items()).encode()).hexdigest()
    save("unet_ffhq10m_meta.npz", nparam=np.array(nparam), names_sha1=np.frombuffer(bytes.fromhex(names_sha), dtype=np.uint8))
    print("ffhq_10m params", nparam)


# ---------------------------------------------------------------------------- full hmc() trace
def gold_hmc_trace():
    D = 32
    net = TinyEpsNet()
    torch.manual_seed(99)
    missing_r = 3 * torch.randperm(D ** 2)[: int(D ** 2 * 0.92)].long()
    missing = torch.cat([missing_r, missing_r + 1, missing_r + 2], dim=0)
    H = Hfuncs.Inpainting(3, D, missing, "cpu")
    sigma_0 = 0.1                                             # already doubled
    algo = Unconditional(net, H, sigma_0)
    betas = ms.get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
    b = torch.from_numpy(betas).float()
    g = torch.Generator().manual_seed(31)
    x_orig = torch.rand(1, 3, D, D, generator=g) * 2 - 1
    y_0 = H.H(x_orig) + sigma_0 * torch.randn(1, H.kept_indices.shape[0], generator=g)
    x = torch.randn(1, 3, D, D, generator=g)
    opt = types.SimpleNamespace(tau=1.0, epsilon=0.05, m=1.0, sigma_0=sigma_0, algo="hmc", noise="ddpm",
                                image_folder="/tmp/nhmc_golden_unused")
    seq, seq_next = [250, 500, 750], [-1, 250, 500]

    rec = dict(x_start=[], prob=[], u=[], saved=[])
    real_randn_like, real_rand, real_save = torch.randn_like, torch.rand, ms.tvu.save_image

    def randn_like(t, *a, **kw):
        rec["x_start"].append(t.detach().clone())
        return real_randn_like(t, *a, **kw)

    def rand(*a, **kw):
        v = real_rand(*a, **kw)
        rec["u"].append(float(v))
        return v

    def rec_min(one, ratio):
        rec["prob"].append(float(ratio))
        return min(one, ratio)

    def save_image(img, path, *a, **kw):
        rec["saved"].append(img.detach().clone())

    torch.randn_like, torch.rand, ms.tvu.save_image = randn_like, rand, save_image
    ms.min = rec_min
    ms.print = lambda *a, **k: None
    try:
        torch.manual_seed(777)
        out = ms.hmc(x, 1, b, seq, seq_next, algo, opt, y_0, H, x_orig)
    finally:
        torch.randn_like, torch.rand, ms.tvu.save_image = real_randn_like, real_rand, real_save
        del ms.min, ms.print
    n_it = len(rec["u"])
    accept = np.array([rec["u"][i] < min(1.0, rec["prob"][i]) for i in range(n_it)])
    print("hmc trace: iterations", n_it, "accepted", int(accept.sum()), "samples", tuple(out.shape))
    K = 32
    save("hmc_trace.npz", missing=missing, x_orig=x_orig, y_0=y_0, x_init=x, sigma_0=np.array(sigma_0),
         u=np.array(rec["u"]), prob=np.array(rec["prob"]), accept=accept,
         x_start=torch.cat(rec["x_start"][:K]), saved=torch.stack(rec["saved"][:K]),
         samples=out.detach(), seed=np.array(777))


# ---------------------------------------------------------------------------- hmc_test_conditioning() trace
def gold_hmc_cond_trace():
    """The adaptive-mass variant (main_sampling.py:776-894) run for real with the tiny eps-net: per-iteration start state,
    uniform draws, acceptance ratios, the variance handed to torch.sort at every mass update (and the iteration it
    happened in), the images saved on accept and the returned samples."""
    D = 32
    net = TinyEpsNet()
    torch.manual_seed(99)
    missing_r = 3 * torch.randperm(D ** 2)[: int(D ** 2 * 0.92)].long()
    missing = torch.cat([missing_r, missing_r + 1, missing_r + 2], dim=0)
    H = Hfuncs.Inpainting(3, D, missing, "cpu")
    sigma_0 = 0.1
    algo = Unconditional(net, H, sigma_0)
    betas = ms.get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
    b = torch.from_numpy(betas).float()
    g = torch.Generator().manual_seed(31)
    x_orig = torch.rand(1, 3, D, D, generator=g) * 2 - 1
    y_0 = H.H(x_orig) + sigma_0 * torch.randn(1, H.kept_indices.shape[0], generator=g)
    x = torch.randn(1, 3, D, D, generator=g)
    opt = types.SimpleNamespace(tau=1.0, epsilon=0.05, m=1.0, sigma_0=sigma_0, algo="hmc", noise="ddpm",
                                image_folder="/tmp/nhmc_golden_unused")
    seq, seq_next = [250, 500, 750], [-1, 250, 500]
    rec = dict(x_start=[], prob=[], u=[], saved=[], var=[], var_it=[])
    real_randn_like, real_rand, real_save, real_sort = torch.randn_like, torch.rand, ms.tvu.save_image, torch.sort

    def randn_like(t, *a, **kw):
        rec["x_start"].append(t.detach().clone())
        return real_randn_like(t, *a, **kw)

    def rand(*a, **kw):
        v = real_rand(*a, **kw)
        rec["u"].append(float(v))
        return v

    def rec_min(one, ratio):
        rec["prob"].append(float(ratio))
        return min(one, ratio)

    def save_image(img, path, *a, **kw):
        rec["saved"].append(img.detach().clone())

    def sort(t, *a, **kw):
        rec["var"].append(t.detach().clone())
        rec["var_it"].append(len(rec["u"]) - 1)          # the iteration whose accept triggered this update
        return real_sort(t, *a, **kw)

    torch.randn_like, torch.rand, ms.tvu.save_image, torch.sort = randn_like, rand, save_image, sort
    ms.min = rec_min
    ms.print = lambda *a, **k: None
    try:
        torch.manual_seed(778)
        out = ms.hmc_test_conditioning(x, 1, b, seq, seq_next, algo, opt, y_0, H, x_orig)
    finally:
        torch.randn_like, torch.rand, ms.tvu.save_image, torch.sort = real_randn_like, real_rand, real_save, real_sort
        del ms.min, ms.print
    n_it = len(rec["u"])
    accept = np.array([rec["u"][i] < min(1.0, rec["prob"][i]) for i in range(n_it)])
    print("hmc_test_conditioning trace: iterations", n_it, "accepted", int(accept.sum()), "mass updates", len(rec["var"]),
          "samples", tuple(out.shape))
    # keep the iterations up to a few mass updates past the point where the running variance switches on (epoch 19)
    acc_cum = np.cumsum(accept)
    K = int(min(n_it, np.searchsorted(acc_cum, 26) + 1))
    print("kept iterations", K, "first mass update at iteration", rec["var_it"][0] if rec["var_it"] else None)
    keep_var = [i for i, it in enumerate(rec["var_it"]) if it < K]
    save("hmc_cond_trace.npz", missing=missing, x_orig=x_orig, y_0=y_0, x_init=x, sigma_0=np.array(sigma_0),
         u=np.array(rec["u"]), prob=np.array(rec["prob"]), accept=accept,
         x_start=torch.cat(rec["x_start"][:K]), saved=torch.stack(rec["saved"][:K]),
         var=torch.stack([rec["var"][i] for i in keep_var]), var_it=np.array([rec["var_it"][i] for i in keep_var]),
         n_samples=np.array(out.shape[0]), last_sample=out[-1].detach(), seed=np.array(778))


# ---------------------------------------------------------------------------- full hmc_latent() trace
def gold_hmc_latent_trace():
    """The REAL hmc_latent (main_sampling_latent.py:623-762) driven with a tiny stand-in LDM."""
    import types as _t
    class _Any(_t.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return lambda *a, **kw: 0.0
    for n in ("omegaconf", "pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.utilities.distributed",
              "taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize", "ldm_loader"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = _Any(n)
    import main_sampling_latent as msl
    from algos.unconditional_latent import Unconditional_Latent
    from oracle.tinynet import TinyLDM
    msl.config, msl.device = ms.config, torch.device("cpu")
    ldm = TinyLDM()
    D, h = 64, 16
    H = Hfuncs.SuperResolution(3, D, 4, "cpu")
    sigma_0 = 0.1
    algo = Unconditional_Latent(ldm, H, sigma_0)
    g = torch.Generator().manual_seed(41)
    x_orig = torch.rand(1, 3, D, D, generator=g) * 2 - 1
    y_0 = H.H(x_orig) + sigma_0 * torch.randn(1, 3 * 16 * 16, generator=g)
    z = torch.randn(1, 3, h, h, generator=g)
    opt = types.SimpleNamespace(tau=1.0, epsilon=0.05, m=1.0, sigma_0=sigma_0, sigma_y=0.5, algo="hmc_latent", noise="ddpm",
                                image_folder="/tmp/nhmc_golden_unused")
    seq, seq_next = [250, 500, 750], [-1, 250, 500]
    rec = dict(x_start=[], prob=[], u=[])
    real_randn_like, real_rand = torch.randn_like, torch.rand

    def randn_like(t, *a, **kw):
        rec["x_start"].append(t.detach().clone())
        return real_randn_like(t, *a, **kw)

    def rand(*a, **kw):
        v = real_rand(*a, **kw)
        rec["u"].append(float(v))
        return v

    def rec_min(one, ratio):
        rec["prob"].append(float(ratio))
        return min(one, ratio)

    torch.randn_like, torch.rand = randn_like, rand
    msl.min = rec_min
    msl.print = lambda *a, **k: None
    try:
        torch.manual_seed(778)
        out = msl.hmc_latent(z, 1, seq, seq_next, algo, opt, y_0, H, x_orig)
    finally:
        torch.randn_like, torch.rand = real_randn_like, real_rand
        del msl.min, msl.print
    n_it = len(rec["u"])
    accept = np.array([rec["u"][i] < min(1.0, rec["prob"][i]) for i in range(n_it)])
    print("hmc_latent trace: iterations", n_it, "accepted", int(accept.sum()), "samples", tuple(out.shape))
    save("hmc_latent_trace.npz", x_orig=x_orig, y_0=y_0, z_init=z, sigma_0=np.array(sigma_0), sigma_y=np.array(0.5),
         u=np.array(rec["u"]), prob=np.array(rec["prob"]), accept=accept, x_start=torch.cat(rec["x_start"]),
         samples=out.detach(), seed=np.array(778))


# ---------------------------------------------------------------------------- latent model family
def gold_latent_models():
    """The reference's own latent U-Net (ldm openaimodel.UNetModel) and VQ decoder (ldm model.Decoder) at a small
    configuration with seeded weights: forward outputs, and for the first stage the decode output and its gradient
    with respect to z through post_quant_conv + decoder (straight-through over the nearest-code lookup).
    taming's VectorQuantizer2 is not vendored upstream: its lookup is restated inline (argmin of
    sum z^2 + sum e^2 - 2 z.e, first minimum) -- that boundary is un-pinned."""
    import types as _t
    class _Any(_t.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return lambda *a, **kw: 0.0
    for n in ("omegaconf", "pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.utilities.distributed",
              "taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = _Any(n)
    from ldm.modules.diffusionmodules.openaimodel import UNetModel as RefUNet
    from ldm.modules.diffusionmodules.model import Decoder as RefDecoder
    from oracle.latent_fill import SMALL_UNET, SMALL_FIRST_STAGE, seeded_fill
    rec = {}
    g = torch.Generator().manual_seed(51)
    # --- latent U-Net
    uc = dict(SMALL_UNET); uc["attention_resolutions"] = list(uc["attention_resolutions"]); uc["channel_mult"] = list(uc["channel_mult"])
    net = RefUNet(**uc).eval()
    sd = seeded_fill({k: v.shape for k, v in net.state_dict().items()}, seed=4321)
    net.load_state_dict(sd)
    x = torch.randn(2, 3, 16, 16, generator=g)
    t = torch.tensor([750.0, 250.0])
    with torch.no_grad():
        rec["unet_out"] = net(x, t)
    rec["unet_x"], rec["unet_t"] = x, t
    rec["unet_keys"] = np.array(list(sd.keys()))
    # --- first stage decode
    fs = SMALL_FIRST_STAGE
    dd = dict(fs["ddconfig"]); dd["ch_mult"] = tuple(dd["ch_mult"]); dd["attn_resolutions"] = list(dd["attn_resolutions"])
    dec = RefDecoder(**dd).eval()
    pq = torch.nn.Conv2d(fs["embed_dim"], dd["z_channels"], 1)
    shapes = {"quantize.embedding.weight": (fs["n_embed"], fs["embed_dim"]), "post_quant_conv.weight": pq.weight.shape,
              "post_quant_conv.bias": pq.bias.shape}
    shapes.update({"decoder." + k: v.shape for k, v in dec.state_dict().items()})
    sd2 = seeded_fill(shapes, seed=8765)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd2.items() if k.startswith("decoder.")})
    with torch.no_grad():
        pq.weight.copy_(sd2["post_quant_conv.weight"]); pq.bias.copy_(sd2["post_quant_conv.bias"])
    E = sd2["quantize.embedding.weight"]
    z = (torch.randn(2, 3, 8, 8, generator=g) * 0.8).requires_grad_(True)
    zf = z.permute(0, 2, 3, 1).reshape(-1, 3)
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) - 2 * torch.einsum("bd,dn->bn", zf, E.t())
    idx = torch.argmin(d, dim=1)
    zq = E[idx].view(2, 8, 8, 3)
    zq = z.permute(0, 2, 3, 1) + (zq - z.permute(0, 2, 3, 1)).detach()
    zq = zq.permute(0, 3, 1, 2).contiguous()
    out = dec(pq(zq))
    w = torch.randn(out.shape, generator=g)
    (gz,) = torch.autograd.grad(out, z, w)
    rec.update(dec_z=z.detach(), dec_idx=idx.view(2, 64), dec_zq=zq.detach(), dec_out=out.detach(), dec_w=w, dec_grad_z=gz,
               dec_keys=np.array(list(sd2.keys())))
    # second-nearest margin: how close the lookup is to a tie (for the un-pinned boundary's tolerance)
    d2 = d.detach().clone(); d2[torch.arange(d2.shape[0]), idx] = float("inf")
    rec["dec_margin"] = (d2.min(1).values - d.detach().min(1).values).min()
    save("latent_models.npz", **rec)


if __name__ == "__main__":
    which = sys.argv[1:] or ["schedule", "inpaint_mask", "operators", "operators_ext", "ddim", "unet", "hmc_trace", "hmc_cond_trace", "hmc_latent_trace", "latent_models"]
    for w in which:
        globals()["gold_" + w]()
