"""oracle/ -- CPU restatement of the reference's `--algo hmc` / `--algo hmc_latent` path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import anything from this package, and only as the checker
(or as the timed CPU baseline) -- never on the product path.  The product package
(`noise-space-hmc_b200/`, import name `nhmc_b200`) must not import `oracle`.

Every function cites the reference file:line it restates (paths relative to the
upstream tree Sunsett5/Noise-space-HMC, mounted read-only at /root/reference in the
build container; it does NOT exist on the GPU box).

Parity pinning: the reference ships no tests and no golden vectors,
so the oracle is pinned against *outputs of the reference itself run in the build
container*: `tests/golden/make_golden.py` imports the reference's own modules
(`obs_functions.Hfuncs`, `algos.unconditional`, `guided_diffusion.unet_ffhq`,
`main_sampling.hmc` under import stubs), records input/output vectors into
`tests/golden/*.npz`, and `tests/test_oracle_golden.py` checks every oracle function
against those vectors.  The one un-pinned boundary is the latent path's
`taming.modules.vqvae.quantize.VectorQuantizer2` (third-party, pinned
`taming-transformers==0.0.1`, absent from /root/reference): "parity unpinned" there.

Arithmetic is floating point (fp32 in the reference), so the restatement is written
with torch CPU tensors (fp32 by default, fp64 on request) and numpy for index work.
"""
