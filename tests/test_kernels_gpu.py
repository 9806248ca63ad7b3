"""GPU parity: every kernel called through the C ABI (libnhmc.so) against the CPU oracle and the golden vectors.

Tolerances: the elementwise kernels follow the reference's fp32 operation order with no FMA contraction and are
checked BIT-EXACT; reductions are fp64 and are checked to 1e-12 relative; operators that sum in a different order
than the reference (block averages, GEMMs) are checked to 2e-5 relative (fp32 round-off), the stated tolerance of
the north star being max-abs 1e-3.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ddim as oddim, hmc as ohmc, operators as oops, schedule as osch

T = torch.from_numpy


@pytest.fixture(scope="module")
def nb():
    import nhmc_b200
    nhmc_b200.lib()
    return nhmc_b200


def dev(t):
    return t.cuda().contiguous()


# ------------------------------------------------------------------------------------------ leapfrog
@pytest.mark.parametrize("B,shape", [(1, (3, 32, 32)), (3, (3, 64, 64)), (2, (3, 256, 256)), (2, (3, 5, 7))])
def test_leapfrog_modes_bit_exact(nb, B, shape):
    from nhmc_b200 import _lib, kernels
    from nhmc_b200.chain import ChainState, new_host_state
    g = torch.Generator().manual_seed(3)
    x = torch.randn((B,) + shape, generator=g); p = torch.randn((B,) + shape, generator=g)
    ga = torch.randn((B,) + shape, generator=g) * 30; gb = torch.randn((B,) + shape, generator=g) * 30
    eps = [0.05, 0.0475, 0.01][:B] + [0.03] * max(0, B - 3)
    sig = [1.7, 0.1, 0.9][:B] + [0.5] * max(0, B - 3)
    hs = new_host_state(B, 1.0, 0.05, 0.1)
    hs["epsilon"] = eps[:B]; hs["sigma_y"] = sig[:B]
    m = 1.0
    loss = torch.tensor([123.5 + b for b in range(B)], dtype=torch.float64)
    for use_gb in (False, True):
        gtot = ga + gb if use_gb else ga
        # oracle, chain by chain (python-double scalars like the reference)
        exp = {}
        for mode in ("first", "middle", "last"):
            xs, ps = [], []
            for b in range(B):
                xb, pb, gbb = x[b:b + 1], p[b:b + 1], gtot[b:b + 1]
                if mode == "first":
                    pn = ohmc.kick(pb, xb, gbb, sig[b], eps[b] / 2); xn = ohmc.drift(xb, pn, eps[b], m)
                elif mode == "middle":
                    pn = ohmc.kick(pb, xb, gbb, sig[b], eps[b]); xn = ohmc.drift(xb, pn, eps[b], m)
                else:
                    pn = ohmc.kick(pb, xb, gbb, sig[b], eps[b]); pn = ohmc.kick(pn, xb, gbb, sig[b], -(eps[b] / 2)); xn = xb
                xs.append(xn); ps.append(pn)
            exp[mode] = (torch.cat(xs), torch.cat(ps))
        for mode, code in (("first", _lib.LF_FIRST), ("middle", _lib.LF_MIDDLE), ("last", _lib.LF_LAST)):
            ch = ChainState(hs.copy(), torch.device("cuda"))
            ws = kernels.leapfrog_workspace(B, "cuda")
            xd, pd = dev(x), dev(p)
            xo = torch.empty_like(xd) if mode == "first" else xd
            kernels.leapfrog(code, xd, None if mode == "last" else xo, pd, dev(ga), dev(gb) if use_gb else None,
                             dev(loss), ch, m, ws)
            torch.cuda.synchronize()
            assert torch.equal(pd.cpu(), exp[mode][1]), (mode, use_gb)
            if mode != "last":
                assert torch.equal(xo.cpu(), exp[mode][0]), (mode, use_gb)
            st = ch.read()
            if mode != "middle":
                w = 0 if mode == "first" else 1
                pe = p if mode == "first" else exp[mode][1]
                sx = (x.double() ** 2).reshape(B, -1).sum(1).numpy(); sp = (pe.double() ** 2).reshape(B, -1).sum(1).numpy()
                np.testing.assert_allclose(st["sum_x2"][:, w], sx, rtol=1e-12)
                np.testing.assert_allclose(st["sum_p2"][:, w], sp, rtol=1e-12)
                np.testing.assert_array_equal(st["loss"][:, w], loss.numpy())
            # second launch on the same workspace: tickets must have reset themselves
            if mode == "first":
                kernels.leapfrog(code, xd, xo, pd, dev(ga), None, dev(loss), ch, m, ws)
                torch.cuda.synchronize()
                np.testing.assert_allclose(ch.read()["sum_x2"][:, 0], sx, rtol=1e-12)


# ------------------------------------------------------------------------------------------ DDIM epilogue
@pytest.mark.parametrize("B,D", [(2, 32), (1, 256), (3, 20)])
def test_ddim_fwd_bwd(nb, B, D):
    from nhmc_b200 import kernels, schedule
    tab = schedule.alpha_table(T(osch.beta_schedule("linear", 1e-4, 0.02, 1000)).float())
    steps = schedule.ddim_steps(tab, [250, 500, 750], [-1, 250, 500])
    g = torch.Generator().manual_seed(5)
    N = 3 * D * D
    for (t, c1, sa, c2, ss) in steps:
        xt = torch.randn(B, 3, D, D, generator=g) * (1.0 if t == 750 else 0.6)
        out6 = torch.randn(B, 6, D, D, generator=g)
        gn_ = torch.randn(B, 3, D, D, generator=g); gn2 = torch.randn(B, 3, D, D, generator=g)
        a, s = tab[t + 1], tab[{750: 500, 500: 250, 250: -1}[t] + 1]
        ref, mask = oddim.epilogue_fwd(xt, out6[:, :3], a, s)
        xo = torch.empty(B, 3, D, D, device="cuda")
        mk = torch.zeros(B, kernels.ddim_mask_words(N), dtype=torch.int32, device="cuda")
        kernels.ddim_fwd(dev(xt), dev(out6), xo, mk, (c1, sa, c2, ss))
        assert torch.equal(xo.cpu(), ref), t                                   # bit-exact fp32
        # unpack the mask: word j of group gidx, bit l <-> element 128*gidx + 4*l + j
        mw = mk.cpu().numpy().view(np.uint32).reshape(B, -1, 4)
        bits = ((mw[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)   # [B, G, j, l]
        flat = bits.transpose(0, 1, 3, 2).reshape(B, -1)[:, :N]
        assert np.array_equal(flat, mask.reshape(B, -1).numpy()), t
        # backward, with and without the second gradient input, fp32 and bf16 cotangent buffers
        gx_ref, ge_ref = oddim.epilogue_bwd(gn_ + gn2, mask, a, s)
        gx = torch.empty(B, 3, D, D, device="cuda"); ge = torch.zeros(B, 6, D, D, device="cuda")
        kernels.ddim_bwd(dev(gn_), dev(gn2), mk, gx, ge, (c1, sa, c2, ss))
        assert torch.equal(gx.cpu(), gx_ref) and torch.equal(ge[:, :3].cpu(), ge_ref)
        assert float(ge[:, 3:].abs().max()) == 0.0
        gx_ref1, ge_ref1 = oddim.epilogue_bwd(gn_, mask, a, s)
        ge16 = torch.zeros(B, 6, D, D, device="cuda", dtype=torch.bfloat16)
        kernels.ddim_bwd(dev(gn_), None, mk, gx, ge16, (c1, sa, c2, ss))
        assert torch.equal(gx.cpu(), gx_ref1)
        assert torch.equal(ge16[:, :3].cpu(), ge_ref1.to(torch.bfloat16))
        kernels.ddim_bwd(dev(gn_), None, mk, gx, None, (c1, sa, c2, ss))
        assert torch.equal(gx.cpu(), gx_ref1)
        # bf16 network output
        xo16 = torch.empty(B, 3, D, D, device="cuda")
        kernels.ddim_fwd(dev(xt), dev(out6).to(torch.bfloat16), xo16, None, (c1, sa, c2, ss))
        ref16, _ = oddim.epilogue_fwd(xt, out6[:, :3].to(torch.bfloat16).float(), a, s)
        assert torch.equal(xo16.cpu(), ref16)


def test_ddim_rejects_bad_arguments(nb):
    from nhmc_b200 import kernels
    x = torch.zeros(1, 3, 5, 5, device="cuda")           # N = 75 is not a multiple of 4
    with pytest.raises(nb.NhmcError):
        kernels.ddim_fwd(x, x, torch.empty_like(x), None, (0.1, 0.9, 0.2, 0.8))
    with pytest.raises(nb.NhmcError):
        kernels.ddim_fwd(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4), None, (0.1, 0.9, 0.2, 0.8))


# ------------------------------------------------------------------------------------------ operators
def _ops_from_golden(g, D=32):
    from nhmc_b200 import operators as P
    pairs = {
        "inp": (P.Inpainting(3, D, T(g["inp_missing"]), "cuda"), oops.Inpainting.from_missing_indices(3, D, g["inp_missing"])),
        "sr4": (P.SuperResolution(3, D, 4, "cuda"), oops.SuperResolution(3, D, 4)),
        "sr16": (P.SuperResolution(3, D, 16, "cuda"), oops.SuperResolution(3, D, 16)),
        "sr2": (P.SuperResolution(3, D, 2, "cuda"), oops.SuperResolution(3, D, 2)),
        "hdr": (P.HDR(3, D, "cuda"), oops.HDR()),
    }
    tabs = [T(g["deblur_tab_" + k]) for k in ("U_small1", "V_small1", "U_small2", "V_small2", "singulars", "perm")]
    pairs["deblur"] = (P.Deblurring2D(None, None, 3, D, "cuda", tables=tabs), oops.Deblurring2D(3, D, *tabs))
    pairs["phase"] = (P.PhaseRetrievalOperator(2.0, "cuda", 3, D), oops.PhaseRetrieval(2.0))      # L = 160 = 5 * 32
    return pairs


@pytest.mark.parametrize("name", ["inp", "sr4", "sr16", "sr2", "hdr", "deblur", "phase"])
def test_operator_vs_golden_and_oracle(nb, golden, name):
    g = golden("operators.npz")
    op, oop = _ops_from_golden(g)[name]
    x, y = T(g[name + "_x"]), T(g[name + "_y"])
    B = x.shape[0]
    exact = name in ("inp", "hdr")
    tol = dict(atol=0, rtol=0) if exact else dict(atol=3e-5, rtol=2e-5)
    Hx = op.H(dev(x.clamp(-1, 1))).cpu()
    np.testing.assert_allclose(Hx.numpy(), g[name + "_Hx"].reshape(Hx.shape), **tol)
    loss, grad, xc = op.loss_grad(dev(x), dev(y))
    torch.cuda.synchronize()
    assert torch.equal(xc.cpu(), x.clamp(-1, 1))
    np.testing.assert_allclose(loss.sum().item(), float(g[name + "_loss"]), rtol=2e-5)
    if name == "phase":
        pinv = op.H_pinv(dev(y)).cpu()
        np.testing.assert_allclose(pinv.numpy(), g["phase_pinv"], atol=2e-5, rtol=1e-4)
        _, oloss = oop.residual(x, y)
        np.testing.assert_allclose(loss.cpu().numpy(), oloss.numpy(), rtol=2e-5)
        scale = float(np.abs(g["phase_grad"]).max())
        np.testing.assert_allclose(grad.cpu().numpy(), g["phase_grad"], atol=5e-5 * scale, rtol=1e-4)
        return
    _, oloss = oop.residual(x, y)
    np.testing.assert_allclose(loss.cpu().numpy(), oloss.numpy(), rtol=1e-12 if exact else 2e-5)
    scale = max(1.0, float(np.abs(g[name + "_grad"]).max()))
    np.testing.assert_allclose(grad.cpu().numpy(), g[name + "_grad"], atol=0 if exact else 3e-5 * scale, rtol=0 if exact else 2e-5)
    pinv = op.H_pinv(dev(y)).cpu()
    np.testing.assert_allclose(pinv.reshape(g[name + "_pinv"].shape).numpy(), g[name + "_pinv"],
                               atol=5e-4 if name == "deblur" else (0 if name in ("inp", "hdr") else 1e-6), rtol=1e-4 if name == "deblur" else 0)
    if op.is_linear():
        r = torch.randn(op.meas_shape(B), generator=torch.Generator().manual_seed(1))
        Ht = op.Ht(dev(r)).cpu()
        np.testing.assert_allclose(Ht.numpy(), oop.Ht(r).reshape(Ht.shape).numpy(), atol=3e-5, rtol=2e-5)
        # adjoint identity <Hx, r> = <x, H^T r>
        lhs = float((op.H(dev(x)).cpu().double() * r.double()).sum()); rhs = float((x.reshape(B, -1).double() * Ht.double()).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def _ops_ext_from_golden(g, D=32):
    from nhmc_b200 import operators as P
    tabs = [T(g["dgauss_tab_" + k]) for k in ("U_small", "V_small", "singulars", "perm")]
    fac = [T(g["color_tab_" + k]) for k in ("U_small", "singulars_small", "V_small")]
    return {
        "dgauss": (P.Deblurring(None, 3, D, "cuda", tables=tabs), oops.Deblurring(3, D, *tabs)),
        "color": (P.Colorization(D, "cuda", factors=fac), oops.Colorization(D, *fac)),
        "deno": (P.Denoising(3, D, "cuda"), oops.Denoising()),
        "box": (P.Inpainting(3, D, T(g["box_missing"]), "cuda"), oops.Inpainting.from_missing_indices(3, D, g["box_missing"])),
        "cs": (P.WalshHadamardCS(3, D, 4, T(g["cs_perm"]), "cuda"), oops.WalshHadamardCS(3, D, 4, g["cs_perm"])),
        "srbic": (P.SRConv(None, 3, D, "cuda", stride=4, factors=[T(g["srbic_tab_" + k]) for k in ("U_small", "singulars_small", "V_small")]),
                  oops.SRConv(3, D, 4, *[T(g["srbic_tab_" + k]) for k in ("U_small", "singulars_small", "V_small")])),
    }


@pytest.mark.parametrize("name", ["dgauss", "color", "deno", "box", "srbic", "cs"])
def test_next_row_operator_vs_golden_and_oracle(nb, golden, name):
    """SURVEY.md 8f N4 operators through the C ABI against the reference's recorded outputs and the oracle."""
    g = golden("operators_ext.npz")
    op, oop = _ops_ext_from_golden(g)[name]
    x, y = T(g[name + "_x"]), T(g[name + "_y"])
    B = x.shape[0]
    exact = name in ("deno", "box", "cs")                 # cs: same butterfly order as the reference -> bit-identical fp32
    tol = dict(atol=0, rtol=0) if exact else dict(atol=3e-5, rtol=2e-5)
    Hx = op.H(dev(x.clamp(-1, 1))).cpu()
    np.testing.assert_allclose(Hx.numpy(), g[name + "_Hx"].reshape(Hx.shape), **tol)
    loss, grad, xc = op.loss_grad(dev(x), dev(y))
    torch.cuda.synchronize()
    assert torch.equal(xc.cpu(), x.clamp(-1, 1))
    np.testing.assert_allclose(loss.sum().item(), float(g[name + "_loss"]), rtol=2e-5)
    _, oloss = oop.residual(x, y)
    np.testing.assert_allclose(loss.cpu().numpy(), oloss.numpy(), rtol=1e-12 if exact else 2e-5)
    scale = max(1.0, float(np.abs(g[name + "_grad"]).max()))
    # cs: H, loss and H^+ follow the reference's butterfly order exactly; its autograd gradient runs the stages in reverse
    # order (a different fp32 rounding sequence of the same transform), so the gradient agrees to rounding only
    gexact = exact and name != "cs"
    np.testing.assert_allclose(grad.cpu().numpy(), g[name + "_grad"], atol=0 if gexact else 3e-5 * scale, rtol=0 if gexact else 2e-5)
    pinv = op.H_pinv(dev(y)).cpu()
    np.testing.assert_allclose(pinv.reshape(g[name + "_pinv"].shape).numpy(), g[name + "_pinv"],
                               atol=5e-4 if name == "dgauss" else (0 if exact else 2e-6), rtol=1e-4 if name == "dgauss" else 0)
    r = torch.randn(op.meas_shape(B), generator=torch.Generator().manual_seed(1))
    Ht = op.Ht(dev(r)).cpu()
    np.testing.assert_allclose(Ht.numpy(), oop.Ht(r).reshape(Ht.shape).numpy(), atol=3e-5, rtol=2e-5)
    lhs = float((op.H(dev(x)).cpu().double() * r.double()).sum()); rhs = float((x.reshape(B, -1).double() * Ht.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_next_row_operators_full_size_and_from_reference(nb):
    """FFHQ-256 shapes of the N4 operators (loss / gradient vs the oracle) and construction from duck-typed
    reference objects (`from_reference` reads the attributes the reference's classes expose)."""
    import types
    from nhmc_b200 import operators as P
    D, B = 256, 2
    g = torch.Generator().manual_seed(19)
    x = torch.randn(B, 3, D, D, generator=g) * 0.8
    x_true = torch.rand(B, 3, D, D, generator=g) * 2 - 1
    k5 = P.gaussian_taps(10.0, half=2)
    dg = P.Deblurring(k5, 3, D, "cuda")
    ref_like = types.SimpleNamespace(U_small=dg.U_small.cpu(), V_small=dg.V_small.cpu(), _singulars=dg._singulars.cpu(), _perm=dg._perm.cpu(),
                                     singulars_small=torch.zeros(D), channels=3, img_dim=D)
    dg2 = P.from_reference(ref_like, 3, D, "cuda")
    assert isinstance(dg2, P.Deblurring) and torch.equal(dg2.gain, dg.gain)
    col = P.Colorization(D, "cuda")
    col2 = P.from_reference(types.SimpleNamespace(U_small=col.U_small.cpu(), singulars_small=col.singulars_small.cpu(), V_small=col.V_small.cpu(),
                                                  Vt_small=col.V_small.cpu().t(), img_dim=D, channels=3), 3, D, "cuda")
    assert isinstance(col2, P.Colorization) and list(col2.desc.coef) == list(col.desc.coef)
    import random
    random.seed(5678)
    left, up = random.randint(16, 112), random.randint(16, 112)
    random.seed(5678)
    box = P.Inpainting.box(3, D, "cuda")                                   # draws left, up like main_sampling.py:292-293
    assert box.desc.n_kept == D * D - 128 * 128
    cases = {
        "dgauss": (dg, oops.Deblurring(3, D, dg.U_small.cpu(), dg.V_small.cpu(), dg._singulars.cpu(), dg._perm.cpu())),
        "color": (col, oops.Colorization(D, col.U_small.cpu(), col.singulars_small.cpu(), col.V_small.cpu())),
        "deno": (P.Denoising(3, D, "cuda"), oops.Denoising()),
        "box": (box, oops.Inpainting.from_missing_indices(3, D, oops.box_missing_indices(3, D, left, up).numpy())),
    }
    wh = P.WalshHadamardCS.random(3, D, 4, "cuda")                                # --deg cs4 (main_sampling.py:322-325)
    assert wh.meas_shape(B) == (B, 3 * D * D // 4)
    cases["cs"] = (wh, oops.WalshHadamardCS(3, D, 4, wh.perm.cpu()))
    srb = P.SRConv(P.bicubic_taps(4), 3, D, "cuda", stride=4)                     # --deg sr_bicubic4 (main_sampling.py:265-284)
    assert srb.meas_shape(B) == (B, 3 * 64 * 64)
    cases["srbic"] = (srb, oops.SRConv(3, D, 4, srb.U_small.cpu(), srb.singulars_small.cpu(), srb.V_small.cpu()))
    srb2 = P.from_reference(types.SimpleNamespace(U_small=srb.U_small.cpu(), singulars_small=srb.singulars_small.cpu(), V_small=srb.V_small.cpu(),
                                                  small_dim=64, ratio=4, channels=3, img_dim=D), 3, D, "cuda")
    assert isinstance(srb2, P.SRConv) and torch.equal(srb2.L, srb.L)
    for name, (op, oop) in cases.items():
        y = oop.H(x_true); y = y + 0.1 * torch.randn(y.shape, generator=g)
        Hx = op.H(dev(x_true)).cpu()
        np.testing.assert_allclose(Hx.numpy(), oop.H(x_true).reshape(Hx.shape).numpy(), atol=5e-5, rtol=2e-5, err_msg=name)
        loss, grad, xc = op.loss_grad(dev(x), dev(y))
        _, oloss = oop.residual(x, y)
        np.testing.assert_allclose(loss.cpu().numpy(), oloss.numpy(), rtol=2e-5, err_msg=name)
        ograd = oop.grad(x, y)
        np.testing.assert_allclose(grad.cpu().numpy(), ograd.numpy(), atol=5e-5 * max(1.0, float(ograd.abs().max())), rtol=2e-5, err_msg=name)
        assert float(grad.cpu()[x.abs() > 1].abs().max()) == 0.0


@pytest.mark.parametrize("kind", ["inp", "sr4", "sr16", "hdr", "deblur", "phase"])
def test_operator_full_size(nb, kind):
    """FFHQ-256 shapes (BASELINE configs 1-3), B = 2 chains: loss/gradient against the oracle + properties."""
    from nhmc_b200 import operators as P
    D, B = 256, 2
    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, 3, D, D, generator=g) * 0.8
    if kind == "inp":
        torch.manual_seed(5678)
        op = P.Inpainting.random(3, D, "cuda")
        assert op.desc.n_kept == 5243 and op.kept_pix[:3].tolist() == [20, 51, 59]           # SURVEY.md 8c pin
        oop = oops.Inpainting(3, D, np.setdiff1d(np.arange(D * D), op.kept_pix.cpu().numpy()))
    elif kind.startswith("sr"):
        r = int(kind[2:]); op = P.SuperResolution(3, D, r, "cuda"); oop = oops.SuperResolution(3, D, r)
    elif kind == "hdr":
        op = P.HDR(3, D, "cuda"); oop = oops.HDR()
    elif kind == "phase":
        op = P.PhaseRetrievalOperator(2.0, "cuda", 3, D); oop = oops.PhaseRetrieval(2.0)      # L = 384 = 3 * 128
        assert op.meas_shape(B) == (B, 3, 384, 384)
    else:
        op = P.Deblurring2D(P.gaussian_taps(1.0), P.gaussian_taps(20.0), 3, D, "cuda")
        assert int((op.gain != 0).sum()) == 3 * 49486
        oop = oops.Deblurring2D(3, D, op.U_small1.cpu(), op.V_small1.cpu(), op.U_small2.cpu(), op.V_small2.cpu(),
                                op._singulars.cpu(), op._perm.cpu())
    x_true = torch.rand(B, 3, D, D, generator=g) * 2 - 1
    y = oop.H(x_true); y = y + 0.1 * torch.randn(y.shape, generator=g)
    Hx = op.H(dev(x_true)).cpu()
    np.testing.assert_allclose(Hx.numpy(), oop.H(x_true).reshape(Hx.shape).numpy(), atol=5e-5, rtol=2e-5)
    loss, grad, xc = op.loss_grad(dev(x), dev(y))
    _, oloss = oop.residual(x, y)
    np.testing.assert_allclose(loss.cpu().numpy(), oloss.numpy(), rtol=2e-5)
    ograd = oop.grad(x, y)
    np.testing.assert_allclose(grad.cpu().numpy(), ograd.numpy(), atol=5e-5 * max(1.0, float(ograd.abs().max())), rtol=2e-5)
    # gradient is zero wherever the final clip saturates
    sat = (x.abs() > 1)
    assert float(grad.cpu()[sat].abs().max()) == 0.0
    # no_clamp mode (latent path): clip and mask disabled
    op.set_clamp(False)
    loss2, grad2, xc2 = op.loss_grad(dev(x), dev(y))
    assert torch.equal(xc2.cpu(), x)
    r2 = y - oop.H(x)
    np.testing.assert_allclose(loss2.cpu().numpy(), (r2.double() ** 2).reshape(B, -1).sum(1).numpy(), rtol=2e-5)
    op.set_clamp(True)


def test_inpaint_empty_and_full_masks(nb):
    from nhmc_b200 import operators as P
    D = 16
    x = torch.randn(1, 3, D, D)
    none_missing = P.Inpainting(3, D, torch.zeros(0, dtype=torch.long), "cuda")
    assert none_missing.desc.n_kept == D * D
    y = none_missing.H(dev(x)).cpu()
    assert torch.equal(y.reshape(1, D * D, 3).permute(0, 2, 1).reshape(1, 3, D, D), x)
    all_missing = P.Inpainting(3, D, torch.arange(3 * D * D), "cuda")
    assert all_missing.desc.n_kept == 0 and all_missing.H(dev(x)).shape == (1, 0)
    loss, grad, _ = all_missing.loss_grad(dev(x), torch.zeros(1, 0, device="cuda"))
    assert float(loss) == 0.0 and float(grad.abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ accept / commit
@pytest.mark.parametrize("sched", ["pixel", "latent"])
def test_accept_schedule_matches_oracle_state_machine(nb, sched):
    from nhmc_b200 import _lib, kernels
    from nhmc_b200.chain import ChainState, new_host_state
    B, N = 4, 64
    rng = np.random.RandomState(7)
    code = _lib.SCHED_PIXEL if sched == "pixel" else _lib.SCHED_LATENT
    slots = 20 if sched == "pixel" else 10
    hs = new_host_state(B, 1.0, 0.05, 0.1, 0.5 if sched == "latent" else None)
    ch = ChainState(hs, torch.device("cuda"))
    if sched == "pixel":
        ref = [osch.PixelSchedule(1.0, 0.05, 0.1) for _ in range(B)]
    else:
        ref = [osch.LatentSchedule(1.0, 0.05, 0.1, 0.5) for _ in range(B)]
    x = torch.zeros(B, N, device="cuda"); xk = torch.zeros(B, N, device="cuda")
    samples = torch.zeros(B, slots, N, device="cuda")
    expect_samples = [[] for _ in range(B)]; keep = [None] * B
    p_acc = [0.9, 0.6, 0.45, 0.3]
    for it in range(400):
        if all(r.done() for r in ref):
            break
        kernels.begin_iteration(ch, code)
        st = ch.read()
        for b in range(B):
            if not ref[b].done():
                sig, eps = ref[b].begin_iteration()
                assert abs(st["sigma_y"][b] - sig) <= 1e-14 * sig and st["epsilon"][b] == eps, (it, b)
        # force the decision through the energies: H1 - H0 = -inf-ish (accept) or +large (reject)
        acc = rng.rand(B) < p_acc
        st["sum_x2"][:, 0] = 0; st["sum_p2"][:, 0] = 0; st["loss"][:, 0] = 0
        st["sum_x2"][:, 1] = np.where(acc, 0.0, 200.0); st["sum_p2"][:, 1] = 0; st["loss"][:, 1] = 0
        ch.write(st)
        u = torch.full((B,), 0.5)
        kernels.accept(ch, u.cuda(), 1.0, code, slots)
        xp = torch.full((B, N), float(it + 1), device="cuda"); xh = torch.full((B, N), -float(it + 1), device="cuda")
        kernels.commit(ch, x, xp, xh, xk if sched == "latent" else None, samples, code, slots)
        st = ch.read()
        for b in range(B):
            if ref[b].done():
                assert st["accept"][b] == 0 and st["active"][b] == 0
                continue
            col = ref[b].end_iteration(bool(acc[b]))
            assert bool(st["accept"][b]) == bool(acc[b])
            assert st["epoch"][b] == ref[b].epoch and st["rejected"][b] == ref[b].rejected
            assert st["tau"][b] == ref[b].tau and st["epsilon"][b] == ref[b].epsilon
            if sched == "latent":
                assert abs(st["sigma_y"][b] - ref[b].sigma_y) <= 1e-14 * ref[b].sigma_y   # device pow vs libm pow
                if acc[b]:
                    if col and keep[b] is not None:
                        expect_samples[b].append(keep[b])
                    keep[b] = -float(it + 1)
            elif col:
                expect_samples[b].append(-float(it + 1))
            assert bool(st["active"][b]) == (not ref[b].done())
            if acc[b]:
                assert float(x[b, 0]) == float(it + 1)
    assert all(r.done() for r in ref)
    st = ch.read()
    for b in range(B):
        n = int(st["n_collected"][b])
        assert n == len(expect_samples[b])
        got = samples[b].cpu()[:, 0].tolist()
        if n <= slots:
            assert got[:n] == expect_samples[b]
        else:
            k = n % slots
            assert got[k:] + got[:k] == expect_samples[b][-slots:]
    if sched == "pixel":
        assert all(int(st["n_collected"][b]) == 20 and int(st["accepted"][b]) == 100 for b in range(B))


def test_accept_probability_and_energy_order(nb):
    """H in fp32, reference order; prob = min(1, exp(-dH)); u < prob decides (main_sampling.py:697,717-720)."""
    from nhmc_b200 import _lib, kernels
    from nhmc_b200.chain import ChainState, new_host_state
    B = 3
    hs = new_host_state(B, 1.0, 0.05, 0.1)
    hs["sigma_y"] = [1.7, 0.3, 0.1]
    hs["sum_x2"] = [[196000.3, 196010.9], [3000.0, 2999.0], [100.0, 100.5]]
    hs["sum_p2"] = [[197001.1, 196995.2], [3100.0, 3102.5], [90.0, 89.0]]
    hs["loss"] = [[5000.0, 4990.0], [12.0, 12.4], [3.0, 2.9]]
    ch = ChainState(hs, torch.device("cuda"))
    u = torch.tensor([0.3, 0.3, 0.3])
    kernels.accept(ch, u.cuda(), 1.0, _lib.SCHED_PIXEL, 20)
    st = ch.read()
    for b in range(B):
        H = []
        for w in range(2):
            x2 = torch.tensor(hs["sum_x2"][b][w], dtype=torch.float64).float(); p2 = torch.tensor(hs["sum_p2"][b][w], dtype=torch.float64).float()
            ls = torch.tensor(hs["loss"][b][w], dtype=torch.float64).float()
            Hw = torch.tensor(0.5) * x2 + torch.tensor(1 / (2 * float(hs["sigma_y"][b]) ** 2), dtype=torch.float32) * ls + torch.tensor(0.5) * p2 * torch.tensor(1.0)
            H.append(Hw)
        assert st["H"][b][0] == H[0].item() and st["H"][b][1] == H[1].item()
        prob = min(1.0, float(torch.exp(-(H[1] - H[0]))))
        assert abs(st["prob"][b] - prob) <= 2e-6 * max(prob, 1e-30)
        assert bool(st["accept"][b]) == (0.3 < st["prob"][b])
