"""Noise schedule, timestep sequence and HMC annealing schedule (oracle; test infrastructure).

Restates:
  * get_beta_schedule          /root/reference/main_sampling.py:38-68
  * compute_alpha              /root/reference/main_sampling.py:70-73  (dup functions/denoising.py:6-9)
  * seq / seq_next             /root/reference/main_sampling.py:469-471
  * pixel sigma_y/eps schedule /root/reference/main_sampling.py:683-689, 743-749
  * latent schedule            /root/reference/main_sampling_latent.py:646,693-707,726-732
  * LDM alphas_cumprod         /root/reference/ldm/modules/diffusionmodules/util.py:21-25,
                               /root/reference/ldm/models/diffusion/ddpm.py:117-170
"""
import math

import numpy as np
import torch


def beta_schedule(kind, beta_start, beta_end, num_steps):
    """float64 numpy betas; main_sampling.py:38-68."""
    if kind == "linear":
        b = np.linspace(beta_start, beta_end, num_steps, dtype=np.float64)
    elif kind == "quad":
        b = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_steps, dtype=np.float64) ** 2
    elif kind == "const":
        b = beta_end * np.ones(num_steps, dtype=np.float64)
    elif kind == "jsd":
        b = 1.0 / np.linspace(num_steps, 1, num_steps, dtype=np.float64)
    elif kind == "sigmoid":
        z = np.linspace(-6, 6, num_steps)
        b = 1.0 / (np.exp(-z) + 1.0) * (beta_end - beta_start) + beta_start
    else:
        raise NotImplementedError(kind)
    assert b.shape == (num_steps,)
    return b


def alpha_bar_table(betas):
    """fp32 table T with T[t+1] = abar_t and T[0] = abar_{-1} = 1.

    main_sampling.py:70-73: `cat([0], beta)`, `(1-beta).cumprod(0)` in fp32 (b was
    `.float()`-ed at main_sampling.py:475), `index_select(t+1)`.
    """
    b = torch.as_tensor(np.asarray(betas)).float()
    b = torch.cat([torch.zeros(1), b], dim=0)
    return (1 - b).cumprod(dim=0)


def compute_alpha(betas, t):
    """abar for integer timesteps t (tensor/list), shaped [n,1,1,1]; main_sampling.py:70-73."""
    tab = alpha_bar_table(betas)
    t = torch.as_tensor(t).long()
    return tab.index_select(0, t + 1).view(-1, 1, 1, 1)


def timestep_sequence(num_timesteps, timesteps):
    """(seq, seq_next) exactly as main_sampling.py:469-471."""
    skip = num_timesteps // (timesteps + 1)
    seq = list(range(skip, num_timesteps, skip))
    seq_next = [-1] + list(seq[:-1])
    return seq, seq_next


def ldm_alphas_cumprod(linear_start=0.0015, linear_end=0.0195, n_timestep=1000):
    """LDM 'linear' (sqrt-linear) schedule, float64 -> float32 buffers.

    ldm/modules/diffusionmodules/util.py:21-25 (`linspace(sqrt(start), sqrt(end))**2` in
    float64) and ddpm.py:117-170 (`alphas_cumprod = cumprod(1-betas)`, registered as fp32;
    `alphas_cumprod_prev = [1, alphas_cumprod[:-1]]`).
    Returns (alphas_cumprod, alphas_cumprod_prev) as fp32 torch tensors.
    """
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas, axis=0)
    acp = np.append(1.0, ac[:-1])
    return torch.tensor(ac, dtype=torch.float32), torch.tensor(acp, dtype=torch.float32)


def latent_alpha_table(alphas_cumprod, alphas_cumprod_prev):
    """T[t+1] = alphas_cumprod[t], T[0] = alphas_cumprod_prev[0]; main_sampling_latent.py:771-779."""
    return torch.cat([alphas_cumprod_prev[0:1], alphas_cumprod], dim=0)


class PixelSchedule:
    """Host-side state machine of main_sampling.py:660-749 for ONE chain.

    Only accepted iterations advance `epoch` (the `while`), sigma_y is recomputed from
    `epoch` at the top of every iteration, tau/epsilon are clamped once at epoch == 60
    and decay by 0.95 on every reject once `rejected >= 2` (`rejected` is reset only by
    an accept).  L is fixed from the initial tau/epsilon (main_sampling.py:664).
    """
    EPOCHS = 60
    SAMPLING = 20
    ANNEAL = 1.6
    DECAY = 0.95

    def __init__(self, tau, epsilon, sigma_0):
        self.tau = float(tau)
        self.epsilon = float(epsilon)
        self.sigma_0 = float(sigma_0)          # already doubled (main_sampling.py:348)
        self.L = max(1, math.floor(self.tau / self.epsilon))
        self.epoch = 0
        self.accepted = 0
        self.rejected = 0
        self.sigma_y = None

    def done(self):
        return self.epoch >= self.EPOCHS + 2 * self.SAMPLING

    def begin_iteration(self):
        """main_sampling.py:683-689. Returns (sigma_y, epsilon) for this iteration."""
        if self.epoch < self.EPOCHS:
            self.sigma_y = self.sigma_0 + self.ANNEAL * (1 - self.epoch / self.EPOCHS) ** 2
        elif self.epoch == self.EPOCHS:
            self.sigma_y = self.sigma_0
            if self.tau > 0.1:
                self.tau = 0.1
                self.epsilon = 0.01
        return self.sigma_y, self.epsilon

    def end_iteration(self, accept):
        """main_sampling.py:721-749. Returns True iff the accepted xt is a collected sample."""
        if accept:
            self.accepted += 1
            self.rejected = 0
            collect = self.epoch >= self.EPOCHS + self.SAMPLING
            self.epoch += 1
            return collect
        self.rejected += 1
        if self.rejected >= 2:
            self.tau *= self.DECAY
            self.epsilon *= self.DECAY
        return False


class LatentSchedule:
    """Host-side state machine of main_sampling_latent.py:623-733 for ONE chain.

    `for epoch in range(70)`: rejected iterations consume an epoch too.  sigma_y starts at
    opt.sigma_y and is re-set on accept (geometric interpolation for epoch < 50, else
    sigma_0 with tau=0.1, eps=0.01).  Reject: after 2 rejects tau,eps *= 0.9 and the counter
    resets.  On accept with epoch >= 50 the PREVIOUS accepted sample is collected
    (main_sampling_latent.py:705-713: append happens before x_accept is refreshed).
    """
    EPOCHS = 50
    SAMPLING = 10
    DECAY = 0.9

    def __init__(self, tau, epsilon, sigma_0, sigma_y):
        self.tau = float(tau)
        self.epsilon = float(epsilon)
        self.sigma_0 = float(sigma_0)
        self.sigma_y0 = float(sigma_y)
        self.sigma_y = float(sigma_y)
        self.L = max(1, math.floor(self.tau / self.epsilon))
        self.epoch = 0
        self.accepted = 0
        self.rejected = 0

    def done(self):
        return self.epoch >= self.EPOCHS + 2 * self.SAMPLING

    def begin_iteration(self):
        return self.sigma_y, self.epsilon

    def end_iteration(self, accept):
        """Returns True iff the previously accepted sample is collected on this accept."""
        collect = False
        if accept:
            self.accepted += 1
            self.rejected = 0
            if self.epoch < self.EPOCHS:
                self.sigma_y = self.sigma_y0 * (self.sigma_0 / self.sigma_y0) ** (self.epoch / self.EPOCHS)
            else:
                self.sigma_y = self.sigma_0
                self.tau = 0.1
                self.epsilon = 0.01
                collect = True
        else:
            self.rejected += 1
            if self.rejected >= 2:
                self.tau *= self.DECAY
                self.epsilon *= self.DECAY
                self.rejected = 0
        self.epoch += 1
        return collect


class CondSchedule:
    """Host-side state machine of hmc_test_conditioning (main_sampling.py:776-894) for ONE chain: the adaptive-mass
    variant.  burn = 5, epochs = 40, sampling = 10; the `while` runs until epoch reaches burn + epochs + 4 * sampling
    = 85 accepted iterations; sigma_y = sigma_0 + 0.9 for epoch < burn, sigma_0 + 0.9 (1 - (epoch - burn)/epochs)^3 for
    epoch < epochs, sigma_0 from epoch == epochs on (tau, epsilon clamped to 0.1 / 0.01 there); accepted samples are
    collected from epoch >= epochs + sampling; rejects decay tau / epsilon by 0.95 once rejected >= 2."""
    BURN = 5
    EPOCHS = 40
    SAMPLING = 10
    DECAY = 0.95

    def __init__(self, tau, epsilon, sigma_0):
        self.tau = float(tau)
        self.epsilon = float(epsilon)
        self.sigma_0 = float(sigma_0)
        self.L = max(1, math.floor(self.tau / self.epsilon))
        self.epoch = 0
        self.accepted = 0
        self.rejected = 0
        self.sigma_y = None

    def done(self):
        return self.epoch >= self.BURN + self.EPOCHS + 4 * self.SAMPLING

    def begin_iteration(self):
        """main_sampling.py:809-817."""
        if self.epoch < self.BURN:
            self.sigma_y = self.sigma_0 + 0.9
        elif self.epoch < self.EPOCHS:
            self.sigma_y = self.sigma_0 + 0.9 * (1 - (self.epoch - self.BURN) / self.EPOCHS) ** 3
        elif self.epoch == self.EPOCHS:
            self.sigma_y = self.sigma_0
            if self.tau > 0.1:
                self.tau = 0.1
                self.epsilon = 0.01
        return self.sigma_y, self.epsilon

    def welford_on(self):
        """main_sampling.py:845: the trajectory's running variance is accumulated only from here on."""
        return (self.epoch - self.BURN) > self.EPOCHS // 3

    def mass_update_on(self):
        """main_sampling.py:857: an ACCEPTED iteration re-estimates the diagonal mass from here on (note: a few epochs
        earlier than welford_on, so those updates see an all-zero variance: ranks = the sort's tie order)."""
        return self.epoch > self.EPOCHS // 3

    def end_iteration(self, accept):
        """main_sampling.py:856-892.  Returns True iff the accepted xt is a collected sample."""
        if accept:
            self.accepted += 1
            self.rejected = 0
            collect = self.epoch >= self.EPOCHS + self.SAMPLING
            self.epoch += 1
            return collect
        self.rejected += 1
        if self.rejected >= 2:
            self.tau *= self.DECAY
            self.epsilon *= self.DECAY
        return False
