"""Per-chain sampler state shared with the kernels (struct nhmc_chain of include/nhmc.h)."""
import numpy as np
import torch

from . import _lib

CHAIN_DTYPE = np.dtype([
    ("epsilon", "<f8"), ("tau", "<f8"), ("sigma_y", "<f8"), ("sigma_0", "<f8"), ("sigma_y_init", "<f8"),
    ("sum_x2", "<f8", (2,)), ("sum_p2", "<f8", (2,)), ("loss", "<f8", (2,)),
    ("H", "<f4", (2,)), ("prob", "<f4"), ("u", "<f4"),
    ("epoch", "<i4"), ("accepted", "<i4"), ("rejected", "<i4"), ("n_collected", "<i4"), ("iters", "<i4"),
    ("accept", "<i4"), ("collect", "<i4"), ("slot", "<i4"), ("active", "<i4"), ("reserved", "<i4"),
], align=False)


def new_host_state(n_chains, tau, epsilon, sigma_0, sigma_y=None):
    """Initial python-side locals of hmc()/hmc_latent() for every chain (main_sampling.py:661-679)."""
    st = np.zeros(n_chains, dtype=CHAIN_DTYPE)
    st["epsilon"] = epsilon
    st["tau"] = tau
    st["sigma_0"] = sigma_0
    st["sigma_y_init"] = sigma_y if sigma_y is not None else sigma_0
    st["sigma_y"] = sigma_y if sigma_y is not None else sigma_0
    st["slot"] = -1
    st["active"] = 1
    return st


class ChainState:
    """Device array of nhmc_chain + host mirror helpers."""

    def __init__(self, host_state, device):
        assert host_state.dtype == CHAIN_DTYPE
        if device.type == "cuda":
            assert _lib.lib().nhmc_sizeof_chain() == CHAIN_DTYPE.itemsize, "struct layout mismatch"
        self.n = len(host_state)
        raw = torch.from_numpy(host_state.view(np.uint8).reshape(self.n, CHAIN_DTYPE.itemsize).copy())
        self.dev = raw.to(device)

    def ptr(self):
        return self.dev.data_ptr()

    def view(self, start, n):
        """Chains [start, start + n) of this array as an object the kernel wrappers accept (ptr(), n)."""
        return _ChainView(self, start, n)

    def field_i32(self, name):
        """Device view [n] of a 4-byte integer field (no copy, no sync): e.g. "accept" or "epoch" for device-side masks."""
        off = CHAIN_DTYPE.fields[name][1]
        assert CHAIN_DTYPE.fields[name][0] == np.dtype("<i4") and off % 4 == 0
        return self.dev.view(torch.int32)[:, off // 4]

    def read(self):
        """Host copy (synchronises)."""
        return self.dev.cpu().numpy().reshape(-1).view(CHAIN_DTYPE).copy()

    def write(self, host_state):
        raw = torch.from_numpy(host_state.view(np.uint8).reshape(self.n, CHAIN_DTYPE.itemsize).copy())
        self.dev.copy_(raw)


class _ChainView:
    def __init__(self, parent, start, n):
        assert 0 <= start and start + n <= parent.n
        self.parent, self.start, self.n = parent, start, n

    def ptr(self):
        return self.parent.dev.data_ptr() + self.start * CHAIN_DTYPE.itemsize
