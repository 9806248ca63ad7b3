"""Host-side schedule tables for the sampler.

The reference recomputes these on the device inside the hot loop: `compute_alpha` (cat + cumprod over
1001 entries + index_select, /root/reference/main_sampling.py:70-73) twice per DDIM step plus two tiny
H2D copies for `t` (main_sampling.py:903-904).  Here they are built once on the host and the four fp32
coefficients of each step are passed to the kernels by value.
"""
import math

import numpy as np
import torch


def get_beta_schedule(beta_schedule, *, beta_start, beta_end, num_diffusion_timesteps):
    """Same surface and values as main_sampling.py:38-68 (float64 numpy)."""
    n = num_diffusion_timesteps
    if beta_schedule == "quad":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float64) ** 2
    elif beta_schedule == "linear":
        betas = np.linspace(beta_start, beta_end, n, dtype=np.float64)
    elif beta_schedule == "const":
        betas = beta_end * np.ones(n, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(n, 1, n, dtype=np.float64)
    elif beta_schedule == "sigmoid":
        z = np.linspace(-6, 6, n)
        betas = (1 / (np.exp(-z) + 1)) * (beta_end - beta_start) + beta_start
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (n,)
    return betas


def alpha_table(b):
    """fp32 table T, T[t+1] = abar_t, T[0] = 1 -- the tensor `compute_alpha` indexes (main_sampling.py:71-72).

    `b` is the fp32 beta tensor the reference passes around (any device); the cumprod runs in fp32 on the
    host, exactly the op sequence of the reference.
    """
    b = torch.as_tensor(b).detach().to("cpu", torch.float32)
    b = torch.cat([torch.zeros(1), b], dim=0)
    return (1 - b).cumprod(dim=0)


def compute_alpha(beta, t):
    """Drop-in for main_sampling.compute_alpha (kept for callers that still want the tensor form)."""
    beta = torch.cat([torch.zeros(1).to(beta.device), beta], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)


def timestep_sequence(num_timesteps, timesteps):
    """seq / seq_next of main_sampling.py:469-471."""
    skip = num_timesteps // (timesteps + 1)
    seq = list(range(skip, num_timesteps, skip))
    return seq, [-1] + list(seq[:-1])


def latent_alpha_table(alphas_cumprod, alphas_cumprod_prev):
    """main_sampling_latent.py:771-773: cat([alphas_cumprod_prev[0:1], alphas_cumprod])."""
    ac = torch.as_tensor(alphas_cumprod).detach().to("cpu", torch.float32)
    acp = torch.as_tensor(alphas_cumprod_prev).detach().to("cpu", torch.float32)
    return torch.cat([acp[0:1], ac], dim=0)


def ddim_steps(table, seq, seq_next):
    """[(t, c1, sa, c2, ss)] in execution order (t descending), coefficients as python floats holding
    fp32 values computed with the reference's fp32 tensor ops ((1-at).sqrt(), at.sqrt(), ...)."""
    steps = []
    for i, j in zip(reversed(list(seq)), reversed(list(seq_next))):
        a, s = table[i + 1], table[j + 1]
        steps.append((int(i), float((1 - a).sqrt()), float(a.sqrt()), float((1 - s).sqrt()), float(s.sqrt())))
    return steps


def leapfrog_steps(tau, epsilon):
    """L of main_sampling.py:664 (fixed for the whole run)."""
    return max(1, math.floor(tau / epsilon))
