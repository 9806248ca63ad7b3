"""Import the REAL reference (read-only tree) under stubs (oracle; test infrastructure).

Only usable in the build container where /root/reference exists; the GPU box never has it.
Used by tests/golden/make_golden.py to record golden vectors and by the optional
`tests/test_oracle_vs_reference.py` (skipped when the tree is absent).

Stubs: `skimage.metrics`, `matplotlib.pyplot`, `lpips` are not installed here and are only
used for metrics/plots after the sampler (main_sampling.py:7-8,11).
"""
import os
import sys
import types

REF = os.environ.get("NHMC_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "main_sampling.py"))


class _Any(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return lambda *a, **kw: 0.0


def _install_stubs():
    for n in ("skimage", "skimage.metrics", "matplotlib", "matplotlib.pyplot", "lpips"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = _Any(n)


def import_reference():
    """Returns the reference's `main_sampling` module with config/device injected for CPU."""
    import torch
    import yaml
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _install_stubs()
    import main_sampling as ms
    with open(os.path.join(REF, "configs", "config_ffhq.yml")) as f:
        cfg = yaml.safe_load(f)
    ms.model_config_dict = cfg["model"]
    ms.config = ms.dict2namespace(cfg)
    ms.device = torch.device("cpu")
    return ms
