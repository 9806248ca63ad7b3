"""Noise-space HMC, one chain at a time (oracle; test infrastructure).

Restates hmc() (/root/reference/main_sampling.py:660-774) and hmc_latent()
(/root/reference/main_sampling_latent.py:623-762) with autograd used ONLY for the
eps-network / decoder VJP; every other derivative is the closed form of oracle/ddim.py and
oracle/operators.py (validated executable-exact against the real hmc()).

All tensors are CPU torch; `dtype` selects fp32 (the reference's arithmetic) or fp64.
"""
import math

import numpy as np
import torch

from . import ddim
from .schedule import PixelSchedule, LatentSchedule


# ----------------------------------------------------------------------------------------
# eps-network adapters
# ----------------------------------------------------------------------------------------
def module_eps_net(model, n_chains=None, out_channels=3):
    """Wrap `model(x, t)` ([B,6,H,W] or [B,3,H,W]) as eps_net(xt, t_value) -> (eps, vjp).

    main_sampling.py:903 builds t as `(ones(n) * i)` (a float tensor);
    algos/unconditional.py:17-18 keeps the first three output channels.
    """
    def eps_net(xt, t_value):
        xin = xt.detach().requires_grad_(True)
        with torch.enable_grad():
            t = torch.ones(xin.shape[0]) * t_value
            out = model(xin.float() if xin.dtype != torch.float32 else xin, t)
            eps = out[:, :out_channels]
        def vjp(g_eps):
            (g,) = torch.autograd.grad(eps, xin, g_eps.to(eps.dtype))
            return g.to(xt.dtype)
        return eps.detach().to(xt.dtype), vjp
    return eps_net


def latent_eps_net(ldm_model):
    """eps = model.apply_model(z, t, None) under no_grad (ddpm.py:892-893): zero VJP."""
    def eps_net(zt, t_value):
        with torch.no_grad():
            t = torch.ones(zt.shape[0]) * t_value
            eps = ldm_model.apply_model(zt, t, None)[:, :3]
        return eps.to(zt.dtype), (lambda g: torch.zeros_like(g))
    return eps_net


# ----------------------------------------------------------------------------------------
# gradient evaluation: loss and d loss / d x   (main_sampling.py:693-695, 709-711)
# ----------------------------------------------------------------------------------------
def grad_eval(x, eps_net, alpha_tab, seq, seq_next, op, y, decode=None):
    """Returns (x_hat, loss[B] float64, g = dloss/dx).

    decode: optional (dec, dec_vjp) for the latent path:
      loss = sum (y - H(decode(clip(z_hat))))^2   (main_sampling_latent.py:651,670).
    """
    x_hat, tape = ddim.ddim_map(x, eps_net, alpha_tab, seq, seq_next)
    if decode is None:
        _, loss = op.residual(x_hat, y)
        g_hat = op.grad(x_hat, y)
    else:
        zc = x_hat.clamp(-1, 1).detach().requires_grad_(True)
        with torch.enable_grad():
            img = decode(zc)
            r = y - op.H(img)
            total = (r ** 2).sum()
        (g_zc,) = torch.autograd.grad(total, zc)
        loss = (r.detach().double() ** 2).reshape(r.shape[0], -1).sum(1)
        g_hat = g_zc * ((x_hat >= -1) & (x_hat <= 1)).to(g_zc.dtype)
    g = ddim.ddim_map_vjp(g_hat, tape)
    return x_hat.clamp(-1, 1), loss, g


# ----------------------------------------------------------------------------------------
# leapfrog pieces  (main_sampling.py:702, 706, 713, 715) -- fp32 op-by-op, no FMA contraction
# ----------------------------------------------------------------------------------------
def _f(v, dtype):
    # python float scalar -> tensor scalar of the working dtype (what `py_float * tensor` does)
    return torch.tensor(v, dtype=dtype)


def force(x, g, sigma_y):
    """x + 1/(2 sigma_y^2) * g  (the bracket of main_sampling.py:702,713,715)."""
    return x + _f(1 / (2 * sigma_y ** 2), x.dtype) * g


def kick(p, x, g, sigma_y, step):
    """p - step * (x + g/(2 sigma^2)); step = eps/2, eps, or -eps/2."""
    return p - _f(step, x.dtype) * force(x, g, sigma_y)


def drift(x, p, epsilon, m):
    """x + eps * m^-1 * p  (main_sampling.py:706; the scalar product is formed in python double)."""
    return x + _f(epsilon * m ** (-1), x.dtype) * p


def hamiltonian(x, p, loss, sigma_y, m):
    """fp32 total in the reference's order (main_sampling.py:697,717): prior + data + kinetic."""
    dt = x.dtype
    prior = _f(0.5, dt) * torch.sum(x ** 2, dim=(1, 2, 3))
    data = _f(1 / (2 * sigma_y ** 2), dt) * loss.to(dt)
    kin = _f(0.5, dt) * torch.sum(p * p, dim=(1, 2, 3)) * _f(m ** (-1), dt)
    return prior + data + kin


def accept_prob(H0, H1):
    """min(1, exp(-(H1 - H0))) (main_sampling.py:718-719)."""
    return torch.minimum(torch.ones_like(H0), torch.exp(-(H1 - H0)))


def hmc_iteration(x, p, u, sigma_y, epsilon, L, m, eval_fn):
    """One HMC iteration from state x with momentum draw p and uniform draw u.

    eval_fn(x) -> (x_hat, loss[B], g).  Returns dict with x_prop, p_prop, x_hat (at the proposal),
    H0, H1, accept (bool per chain), loss0, loss1.
    """
    _, loss0, g = eval_fn(x)
    H0 = hamiltonian(x, p, loss0, sigma_y, m)
    xp = x.clone()
    p = kick(p, xp, g, sigma_y, epsilon / 2)
    x_hat = None
    for _ in range(L):
        xp = drift(xp, p, epsilon, m)
        x_hat, loss, g = eval_fn(xp)
        p = kick(p, xp, g, sigma_y, epsilon)
    p = kick(p, xp, g, sigma_y, -(epsilon / 2))
    H1 = hamiltonian(xp, p, loss, sigma_y, m)
    prob = accept_prob(H0, H1)
    accept = None if u is None else (torch.as_tensor(u, dtype=prob.dtype) < prob)
    return dict(x_prop=xp, p_prop=p, x_hat=x_hat, H0=H0, H1=H1, prob=prob, accept=accept,
                loss0=loss0, loss1=loss)


# ----------------------------------------------------------------------------------------
# full chains
# ----------------------------------------------------------------------------------------
def run_pixel_chain(x, eps_net, alpha_tab, seq, seq_next, op, y, tau, epsilon, sigma_0, m=1.0,
                    draw_p=None, draw_u=None, max_iters=None, trace=None):
    """hmc() for one chain (x: [1,C,H,W]). sigma_0 is the DOUBLED value (main_sampling.py:348).

    draw_p(x) / draw_u() default to torch.randn_like / torch.rand(1) from the global CPU
    generator, in the reference's call order (p first, u after the trajectory).
    Returns [S,C,H,W] collected samples (S = 20 when run to completion).
    """
    draw_p = draw_p or (lambda x_: torch.randn_like(x_))
    draw_u = draw_u or (lambda: torch.rand(1))
    sch = PixelSchedule(tau, epsilon, sigma_0)
    eval_fn = lambda x_: grad_eval(x_, eps_net, alpha_tab, seq, seq_next, op, y)
    samples, it = [], 0
    while not sch.done():
        if max_iters is not None and it >= max_iters:
            break
        sigma_y, eps_now = sch.begin_iteration()
        p = draw_p(x) * math.sqrt(m)
        out = hmc_iteration(x, p, draw_u(), sigma_y, eps_now, sch.L, m, eval_fn)
        acc = bool(out["accept"].item())
        if trace is not None:
            trace.append(dict(it=it, epoch=sch.epoch, sigma_y=sigma_y, epsilon=eps_now,
                              H0=float(out["H0"]), H1=float(out["H1"]), prob=float(out["prob"]),
                              accept=acc, loss0=float(out["loss0"]), loss1=float(out["loss1"])))
        collect = sch.end_iteration(acc)
        if acc:
            x = out["x_prop"].detach().clone()
            if collect:
                samples.append(out["x_hat"][0].detach().clone())
        it += 1
    return (torch.stack(samples) if samples else torch.empty(0)), x


def run_latent_chain(z, eps_net, alpha_tab, seq, seq_next, op, y, decode, tau, epsilon, sigma_0,
                     sigma_y, m=1.0, draw_p=None, draw_u=None, max_iters=None, trace=None):
    """hmc_latent() for one chain (z: [1,3,h,w]); returns the last <=10 collected latents."""
    draw_p = draw_p or (lambda x_: torch.randn_like(x_))
    draw_u = draw_u or (lambda: torch.rand(1))
    sch = LatentSchedule(tau, epsilon, sigma_0, sigma_y)
    eval_fn = lambda z_: grad_eval(z_, eps_net, alpha_tab, seq, seq_next, op, y, decode=decode)
    samples, it, z_accept = [], 0, None
    while not sch.done():
        if max_iters is not None and it >= max_iters:
            break
        sig, eps_now = sch.begin_iteration()
        p = draw_p(z) * math.sqrt(m)
        out = hmc_iteration(z, p, draw_u(), sig, eps_now, sch.L, m, eval_fn)
        acc = bool(out["accept"].item())
        if trace is not None:
            trace.append(dict(it=it, epoch=sch.epoch, sigma_y=sig, epsilon=eps_now, H0=float(out["H0"]),
                              H1=float(out["H1"]), prob=float(out["prob"]), accept=acc))
        collect = sch.end_iteration(acc)
        if acc:
            if collect and z_accept is not None:
                samples.append(z_accept[0])
            z_accept = out["x_hat"].detach().clone()
            z = out["x_prop"].detach().clone()
        it += 1
    samples = samples[-10:]
    return (torch.stack(samples) if samples else torch.empty(0)), z


# ----------------------------------------------------------------------------------------
# adaptive diagonal mass: hmc_test_conditioning (main_sampling.py:776-894)
# ----------------------------------------------------------------------------------------
def rank_mass(m2, L):
    """Diagonal mass of main_sampling.py:858-868 from the trajectory's Welford M2 (one chain, flat [N] fp32):
    variance = M2 / (L - 1); rank transform to scores in [-1, 1]; M = exp(scores).  The sort of tied values (the
    all-zero variance of the first updates) follows torch.sort's own order -- implementation-defined, as in the reference."""
    variance = (m2 / (L - 1)).reshape(-1)
    _, sorted_idx = torch.sort(variance)
    ranks = torch.zeros_like(sorted_idx, dtype=torch.float)
    ranks[sorted_idx] = torch.arange(len(variance), dtype=torch.float)
    scores = 2.0 * (ranks / (variance.numel() - 1)) - 1.0
    return torch.exp(1 * scores)


def hmc_cond_iteration(x, p, u, sigma_y, epsilon, L, inv_m, welford, eval_fn):
    """One iteration of hmc_test_conditioning from state x with momentum p (already scaled by sqrt(M)) and uniform u.
    inv_m: [1,C,H,W] fp32 inverse diagonal mass.  welford: accumulate the running mean / M2 of the positions
    (main_sampling.py:845-849).  fp32 op by op in the reference's order."""
    dt = x.dtype
    _, loss0, g = eval_fn(x)
    H0 = (_f(0.5, dt) * torch.sum(x ** 2, dim=(1, 2, 3)) + _f(1 / (2 * sigma_y ** 2), dt) * loss0.to(dt)
          + _f(0.5, dt) * torch.sum(inv_m * p ** 2))
    xp = x.clone()
    p = kick(p, xp, g, sigma_y, epsilon / 2)
    mean, m2 = torch.zeros_like(x), torch.zeros_like(x)
    x_hat = None
    for l in range(L):
        xp = xp + _f(epsilon, dt) * p * inv_m                       # main_sampling.py:835
        x_hat, loss, g = eval_fn(xp)
        p = kick(p, xp, g, sigma_y, epsilon)
        if welford:
            delta = xp - mean
            mean = mean + delta / (l + 1)
            delta2 = xp - mean
            m2 = m2 + delta * delta2
    p = kick(p, xp, g, sigma_y, -(epsilon / 2))
    H1 = (_f(0.5, dt) * torch.sum(xp ** 2, dim=(1, 2, 3)) + _f(1 / (2 * sigma_y ** 2), dt) * loss.to(dt)
          + _f(0.5, dt) * torch.sum(inv_m * p ** 2))
    prob = accept_prob(H0, H1)
    accept = None if u is None else (torch.as_tensor(u, dtype=prob.dtype) < prob)
    return dict(x_prop=xp, p_prop=p, x_hat=x_hat, H0=H0, H1=H1, prob=prob, accept=accept, loss0=loss0, loss1=loss, m2=m2, mean=mean)


def run_cond_chain(x, eps_net, alpha_tab, seq, seq_next, op, y, tau, epsilon, sigma_0, draw_p=None, draw_u=None,
                   max_iters=None, trace=None):
    """hmc_test_conditioning() for one chain (x: [1,C,H,W]); returns ([S,C,H,W] collected samples, final x, final M)."""
    from .schedule import CondSchedule
    draw_p = draw_p or (lambda x_: torch.randn_like(x_))
    draw_u = draw_u or (lambda: torch.rand(1))
    sch = CondSchedule(tau, epsilon, sigma_0)
    eval_fn = lambda x_: grad_eval(x_, eps_net, alpha_tab, seq, seq_next, op, y)
    M = torch.ones_like(x)
    samples, it = [], 0
    while not sch.done():
        if max_iters is not None and it >= max_iters:
            break
        sigma_y, eps_now = sch.begin_iteration()
        p = draw_p(x) * torch.sqrt(M)
        out = hmc_cond_iteration(x, p, draw_u(), sigma_y, eps_now, sch.L, 1.0 / M, sch.welford_on(), eval_fn)
        acc = bool(out["accept"].item())
        if trace is not None:
            trace.append(dict(it=it, epoch=sch.epoch, sigma_y=sigma_y, epsilon=eps_now, H0=float(out["H0"]), H1=float(out["H1"]),
                              prob=float(out["prob"]), accept=acc))
        if acc and sch.mass_update_on():
            M = rank_mass(out["m2"], sch.L).view_as(x)
        collect = sch.end_iteration(acc)
        if acc:
            x = out["x_prop"].detach().clone()
            if collect:
                samples.append(out["x_hat"][0].detach().clone())
        it += 1
    return (torch.stack(samples) if samples else torch.empty(0)), x, M
