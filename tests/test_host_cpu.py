"""CPU-side checks: host logic of the product package, the C-ABI library (loads, exports, layouts), and
that the product path fails loudly instead of falling back when there is no GPU."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

import nhmc_b200
from nhmc_b200 import _lib, chain, schedule, operators as P
from oracle import schedule as osch, unet as ounet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nhmc.h")).read()
    declared = set(re.findall(r"\b(nhmc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    l = nhmc_b200.lib()
    for name in declared:
        assert hasattr(l, name), name
    assert declared == set(_lib.EXPORTS)
    assert l.nhmc_version() >= 100


def test_struct_layouts_match():
    l = nhmc_b200.lib()
    assert l.nhmc_sizeof_chain() == chain.CHAIN_DTYPE.itemsize == 144
    assert ctypes.sizeof(_lib.OpDesc) == l.nhmc_sizeof_op() == 152
    st = chain.new_host_state(3, 1.0, 0.05, 0.1)
    assert st["active"].tolist() == [1, 1, 1] and st["slot"].tolist() == [-1, -1, -1]


def test_argument_errors_are_reported_not_crashed():
    l = nhmc_b200.lib()
    assert l.nhmc_leapfrog(7, None, None, None, None, None, None, None, 1.0, 1, 4, None, None) == 1
    assert b"mode" in l.nhmc_last_error()
    assert l.nhmc_ddim_fwd(None, None, 0, 0, None, None, 0.1, 0.5, 0.1, 0.5, 1, 4, None) == 1
    assert l.nhmc_op_meas_size(None) == -1
    d = _lib.OpDesc(); d.kind, d.channels, d.img_dim, d.ratio = _lib.OP_SR, 3, 256, 4
    assert l.nhmc_op_meas_size(ctypes.byref(d)) == 3 * 64 * 64
    d.kind, d.pad = _lib.OP_PHASE, 64
    assert l.nhmc_op_meas_size(ctypes.byref(d)) == 3 * 384 * 384
    assert l.nhmc_ddim_mask_bytes(196608) == 196608 // 8
    # fused attention: argument and support checks happen before anything touches a device
    assert l.nhmc_attn_fwd(1, None, None, None, 1, 64, 1, 32, 1.0, 0, None) == 1
    buf = ctypes.create_string_buffer(64)
    a = ctypes.addressof(buf) + (-ctypes.addressof(buf)) % 16
    assert l.nhmc_attn_fwd(1, a, a, a, 1, 60, 1, 32, 1.0, 0, None) == 1 and b"multiple of 8" in l.nhmc_last_error()
    assert l.nhmc_attn_fwd(0, a, a, a, 1, 64, 1, 32, 1.0, 0, None) == 3 and b"head_dim" in l.nhmc_last_error()      # fp32: unsupported
    assert l.nhmc_attn_fwd(1, a, a, a, 1, 64, 1, 48, 1.0, 0, None) == 3


def test_schedule_tables_match_golden(golden):
    g = golden("schedule.npz")
    betas = schedule.get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
    assert np.array_equal(betas, g["betas"])
    tab = schedule.alpha_table(torch.from_numpy(betas).float())
    assert np.array_equal(tab.numpy(), g["alpha_table"])
    seq, seq_next = schedule.timestep_sequence(1000, 3)
    assert seq == [250, 500, 750] and seq_next == [-1, 250, 500]
    steps = schedule.ddim_steps(tab, seq, seq_next)
    assert [s[0] for s in steps] == [750, 500, 250]
    assert steps[-1][3] == 0.0 and steps[-1][4] == 1.0          # abar_{-1} = 1: last step returns clip(x0)
    a = schedule.compute_alpha(torch.from_numpy(betas).float(), torch.tensor([750, -1]))
    assert a.shape == (2, 1, 1, 1) and float(a[1]) == 1.0
    assert schedule.leapfrog_steps(1.0, 0.05) == 20 and schedule.leapfrog_steps(0.01, 0.05) == 1
    for kind in ("quad", "const", "jsd", "sigmoid"):
        np.testing.assert_allclose(schedule.get_beta_schedule(kind, beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=50),
                                   osch.beta_schedule(kind, 1e-4, 0.02, 50))


def test_operator_tables_built_like_the_reference(golden):
    g = golden("operators.npz")
    op = P.Inpainting(3, 32, torch.from_numpy(g["inp_missing"]), "cpu")
    gi = golden("inpaint_mask.npz")
    assert np.array_equal(op.kept_indices.numpy(), gi["small_kept_indices"])
    p2k = op.pix_to_kept.numpy()
    assert (p2k >= 0).sum() == op.desc.n_kept and np.array_equal(np.nonzero(p2k >= 0)[0], op.kept_pix.numpy())
    torch.manual_seed(5678)
    big = P.Inpainting.random(3, 256, "cpu")
    assert big.desc.n_kept == 5243 and big.kept_pix[:10].tolist() == gi["kept_head"].tolist()
    d = P.Deblurring2D(torch.from_numpy(g["deblur_kernel1"]), torch.from_numpy(g["deblur_kernel2"]), 3, 32, "cpu")
    assert np.array_equal(d._singulars.numpy(), g["deblur_tab_singulars"]) and np.array_equal(d._perm.numpy(), g["deblur_tab_perm"])
    assert np.array_equal(d.U_small1.numpy(), g["deblur_tab_U_small1"])
    assert np.array_equal(P.gaussian_taps(1.0).numpy(), g["deblur_kernel1"])
    # from_reference duck-typing: an object carrying the reference's attribute names
    fake = types.SimpleNamespace(U_small1=d.U_small1, V_small1=d.V_small1, U_small2=d.U_small2, V_small2=d.V_small2,
                                 _singulars=d._singulars, _perm=d._perm, channels=3, img_dim=32)
    d2 = P.from_reference(fake, 3, 32, "cpu")
    assert torch.equal(d2.gain, d.gain)
    sr = P.from_reference(types.SimpleNamespace(ratio=4, y_dim=8, V_small=None, channels=3, img_dim=32), 3, 32, "cpu")
    assert sr.meas_size() == 3 * 8 * 8


def test_product_fails_loudly_without_cuda():
    op = P.SuperResolution(3, 32, 4, "cpu")
    with pytest.raises(nhmc_b200.NhmcError):
        op.H(torch.zeros(1, 3, 32, 32))
    with pytest.raises(nhmc_b200.NhmcError):
        op.loss_grad(torch.zeros(1, 3, 32, 32), torch.zeros(1, 192))
    opt = types.SimpleNamespace(tau=1.0, epsilon=0.05, m=1.0, sigma_0=0.1)
    with pytest.raises(nhmc_b200.NhmcError):
        nhmc_b200.hmc(torch.zeros(1, 3, 32, 32), 1, torch.zeros(1000), [250], [-1], types.SimpleNamespace(model=None, cls_fn=None),
                      opt, torch.zeros(1, 192), op, None)
    from nhmc_b200 import unet
    m = unet.create_model(image_size=64, num_channels=32, num_res_blocks=1, learn_sigma=True, attention_resolutions=16,
                          num_head_channels=16, use_scale_shift_norm=True, resblock_updown=True, dtype=torch.float32)
    with pytest.raises(nhmc_b200.NhmcError):
        m(torch.zeros(1, 3, 64, 64), torch.zeros(1))
    from nhmc_b200 import kernels
    assert kernels.attn_fwd_supported(torch.bfloat16, 1024, 32) and not kernels.attn_fwd_supported(torch.float32, 1024, 32)
    assert not kernels.attn_fwd_supported(torch.bfloat16, 1023, 32) and not kernels.attn_fwd_supported(torch.bfloat16, 256, 16)
    with pytest.raises(nhmc_b200.NhmcError):
        kernels.attn_fwd(torch.zeros(1, 64, 64, dtype=torch.bfloat16), torch.zeros(1, 32, 64, dtype=torch.bfloat16),
                         torch.zeros(1, 64, 32, dtype=torch.bfloat16), 1, 0.1)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "noise-space-hmc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_unet_state_dict_keys_equal_reference_names(golden):
    from nhmc_b200 import unet
    full = dict(image_size=256, num_channels=128, num_res_blocks=1, channel_mult="", learn_sigma=True,
                attention_resolutions=16, num_heads=4, num_head_channels=64, use_scale_shift_norm=True,
                resblock_updown=True)
    with torch.device("meta"):
        m = unet.create_model(**full)
    shapes = ounet.param_shapes(ounet.resolve_config(**full))
    sd = m.state_dict()
    assert list(sd) == list(shapes)
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    assert sum(int(np.prod(s)) for s in shapes.values()) == int(golden("unet_ffhq10m_meta.npz")["nparam"])


def test_cli_flag_surface_matches_reference():
    """Same flags and defaults as main_sampling.get_parser (main_sampling.py:923-1015); unknown flags ignored."""
    from nhmc_b200 import cli
    p = cli.get_parser()
    opt, unknown = p.parse_known_args("--ni --dataset ffhq --doc ffhq --algo hmc --timesteps 3 --deg inpaint_random "
                                      "--sigma_0 0.05 -i exp/x --tau 1.0 --epsilon 0.05".split())
    assert unknown == ["--ni", "--doc", "ffhq"]
    assert (opt.seed, opt.tau, opt.epsilon, opt.m, opt.sigma_y, opt.num_timesteps, opt.timesteps) == (5678, 1.0, 0.05, 1.0, 0.5, 1000, 3)
    assert opt.image_folder == "exp/x" and opt.noise == "ddpm" and opt.chains == 1
    d = p.parse_known_args(["--deg", "sr4", "--sigma_0", "0.05"])[0]
    assert d.dataset == "celeba" and d.algo == "resample" and d.timesteps == 10 and d.subset_start == -1
    import yaml
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "config_ffhq.yml")))
    ns = cli.dict2namespace(cfg)
    assert ns.model_type == "ffhq" and ns.data.image_size == 256 and ns.diffusion.beta_end == 0.02
    assert cfg["model"]["num_channels"] == 128 and cfg["model"]["attention_resolutions"] == 16 and cfg["sampling"]["batch_size"] == 1
    opt2 = types.SimpleNamespace(deg="sr4", sigma_0=0.05)
    H, s0, deg = cli.prepare_measurement(opt2, ns, "cpu")
    assert s0 == 0.1 and opt2.sigma_0 == 0.1 and H.meas_size() == 3 * 64 * 64       # sigma_0 doubled (main_sampling.py:348)
    Hc, _, _ = cli.prepare_measurement(types.SimpleNamespace(deg="color", sigma_0=0.05), ns, "cpu")
    assert Hc.meas_size() == 256 * 256
    with pytest.raises(SystemExit):      # needs the un-vendored bkse weights upstream (Hfuncs.py:379,386): out of scope
        cli.prepare_measurement(types.SimpleNamespace(deg="deblur_nonlinear", sigma_0=0.05), ns, "cpu")


def test_library_sources_keep_the_abi_contract():
    """include/nhmc.h promises: no device allocation, no process-wide switches, no environment reads, no synchronisation
    inside the library.  Pinned on the sources, so a regression fails here, on the CPU."""
    import glob
    csrc = os.path.join(ROOT, "noise-space-hmc_b200", "csrc")
    banned = [r"\bgetenv\s*\(", r"\bcudaMalloc\w*\s*\(", r"\bcudaFree\w*\s*\(", r"\bcudaMemset\s*\(", r"\bcudaMemcpy\s*\(",
              r"\bcudaDeviceSynchronize\s*\(", r"\bcudaStreamSynchronize\s*\(", r"^static\s+(?!const|constexpr|inline|thread_local|int\s+\w+\(|bool\s+\w+\(|size_t\s+\w+\()[\w:<> ]+\bg_\w+"]   # (the thread-local error text is the one exception)
    for path in sorted(glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cuh"))):
        src = open(path).read()
        for pat in banned:
            m = re.search(pat, src, flags=re.M)
            assert m is None, "%s: %r" % (os.path.basename(path), m.group(0))
    hdr = open(os.path.join(ROOT, "include", "nhmc.h")).read()
    assert not re.search(r"nhmc_\w*set_\w+\s*\(", hdr), "no setters in the ABI: execution-path choices are per-call flags"
