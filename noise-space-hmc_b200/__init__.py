"""nhmc_b200 -- B200-native noise-space HMC sampler (drop-in for the `--algo hmc` / `--algo hmc_latent`
path of Sunsett5/Noise-space-HMC).  Host code is Python/PyTorch; all arithmetic around the score network
runs in hand-written sm_100a kernels behind the C ABI of include/nhmc.h (libnhmc.so, loaded via ctypes).
There is no CPU fallback: calls raise if the library is missing or the tensors are not on a CUDA device.
"""
from . import _lib, schedule, chain                       # noqa: F401
from ._lib import NhmcError, build, lib                   # noqa: F401
from . import kernels, operators, sampler                 # noqa: F401
from .sampler import hmc, hmc_latent, hmc_test_conditioning, HMCSampler, GradEvaluator   # noqa: F401
from .schedule import get_beta_schedule, compute_alpha    # noqa: F401

__all__ = ["hmc", "hmc_latent", "hmc_test_conditioning", "HMCSampler", "GradEvaluator", "operators", "kernels", "schedule", "build", "lib",
           "NhmcError", "get_beta_schedule", "compute_alpha"]
