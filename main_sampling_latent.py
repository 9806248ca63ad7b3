#!/usr/bin/env python
"""Entry point with the reference's latent command line (README: `python main_sampling_latent.py --ni --dataset
ffhq_latent --doc ffhq --algo hmc_latent --timesteps 3 --deg inpaint_random --sigma_0 0.05 -i <dir> --tau 1.0
--epsilon 0.05`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nhmc_b200  # noqa: E402
from nhmc_b200 import cli  # noqa: E402

if __name__ == "__main__":
    cli.main(latent=True)
