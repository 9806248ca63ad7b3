"""Measurement operators behind the reference's `H_funcs` interface (obs_functions/Hfuncs.py).

Each class keeps the reference's duck-typed surface -- `H(vec)`, `Ht(vec)`, `H_pinv(vec)`, `is_linear()` --
and adds `loss_grad(x_hat, y)`, the fused residual / loss / adjoint used by the sampler.  Tables (kept
pixel indices, SVD factors, spectral gain) are built once on the host with the same torch calls as the
reference, or extracted from a reference `H_funcs` object by `from_reference`, and live on the device.
All arithmetic runs in the sm_100a kernels (no CPU path).
"""
import ctypes

import numpy as np
import torch

from . import _lib, kernels
from ._lib import OpDesc, check, lib


class Operator:
    kind = None
    linear = True
    loss_grad_launches = 1     # kernels launched by one nhmc_op_loss_grad call

    def __init__(self, channels, img_dim, device):
        self.channels, self.img_dim = int(channels), int(img_dim)
        self.device = torch.device(device)
        self.desc = OpDesc()
        self.desc.kind = self.kind
        self.desc.channels = self.channels
        self.desc.img_dim = self.img_dim
        self._ws = {}

    def set_clamp(self, on):
        """loss_grad clamps x_hat to [-1,1] first (pixel path, main_sampling.py:709); the latent path
        feeds the un-clipped decoder output (main_sampling_latent.py:651)."""
        self.desc.no_clamp = 0 if on else 1

    # ---- reference surface -----------------------------------------------------------------
    def is_linear(self):
        return self.linear

    def meas_size(self):
        return int(lib().nhmc_op_meas_size(ctypes.byref(self.desc)))

    def meas_shape(self, B):
        return (B, self.meas_size())

    def unclamped(self):
        """A view of this operator whose loss_grad skips clamp(x_hat, -1, 1) and its mask (the latent path feeds the
        un-clipped decoder output, main_sampling_latent.py:651): own descriptor and workspaces, shared tables.  The
        original object is not modified."""
        import copy
        o = copy.copy(self)
        o.desc = OpDesc.from_buffer_copy(self.desc)
        o.desc.no_clamp = 1
        o._ws = {}
        return o

    def _prepare_tables(self):
        """Operator-owned derived tables (the tf32 hi/lo splits of the SVD factors the tensor-core GEMMs read), built once
        by the library into a buffer this object owns; without them every call rebuilds them in its workspace."""
        if self.device.type != "cuda":
            return                                   # host-side construction (argument checks, table extraction): nothing to build
        nb = int(lib().nhmc_op_tables_bytes(ctypes.byref(self.desc)))
        if nb:
            self._tc_tables = torch.empty((nb + 3) // 4, dtype=torch.float32, device=self.device)
            check(lib().nhmc_op_prepare_tables(ctypes.byref(self.desc), self._tc_tables.data_ptr(), kernels._stream()),
                  "nhmc_op_prepare_tables")
            self.desc.tc_tables = self._tc_tables.data_ptr()

    def _workspace(self, B):
        # calls that may run concurrently (lanes / host threads on their own streams) never share reduction scratch
        key = (B, torch.cuda.current_stream(self.device).cuda_stream)
        if key not in self._ws:
            self._ws[key] = kernels.zeros_workspace(lib().nhmc_op_workspace_bytes(ctypes.byref(self.desc), B), self.device)
        return self._ws[key]

    def _img(self, vec):
        B = vec.shape[0]
        x = vec.reshape(B, self.channels, self.img_dim, self.img_dim)
        return kernels._req(x.contiguous().float(), torch.float32, "image")

    def H(self, vec, **kwargs):
        x = self._img(vec)
        B = x.shape[0]
        y = torch.empty(self.meas_shape(B), dtype=torch.float32, device=x.device)
        check(lib().nhmc_op_forward(ctypes.byref(self.desc), x.data_ptr(), y.data_ptr(), B,
                                    self._workspace(B).data_ptr(), kernels._stream()), "nhmc_op_forward")
        return y

    forward = H

    def Ht(self, vec):
        B = vec.shape[0]
        r = kernels._req(vec.reshape(self.meas_shape(B)).contiguous().float(), torch.float32, "measurement")
        out = torch.empty(B, self.channels * self.img_dim ** 2, dtype=torch.float32, device=r.device)
        check(lib().nhmc_op_adjoint(ctypes.byref(self.desc), r.data_ptr(), out.data_ptr(), B,
                                    self._workspace(B).data_ptr(), kernels._stream()), "nhmc_op_adjoint")
        return out

    def H_pinv(self, vec):
        B = vec.shape[0]
        y = kernels._req(vec.reshape(self.meas_shape(B)).contiguous().float(), torch.float32, "measurement")
        out = torch.empty(B, self.channels * self.img_dim ** 2, dtype=torch.float32, device=y.device)
        check(lib().nhmc_op_pinv(ctypes.byref(self.desc), y.data_ptr(), out.data_ptr(), B,
                                 self._workspace(B).data_ptr(), kernels._stream()), "nhmc_op_pinv")
        return out

    # ---- sampler surface -------------------------------------------------------------------
    def loss_grad(self, x_hat, y, g=None, xc=None, loss=None):
        """(loss[B] fp64, g = dloss/dx_hat, clamp(x_hat)) of main_sampling.py:709-711, one fused call."""
        x_hat = kernels._req(x_hat, torch.float32, "x_hat")
        B = x_hat.shape[0]
        y = kernels._req(y.reshape(self.meas_shape(B)), torch.float32, "y")
        g = torch.empty_like(x_hat) if g is None else g
        xc = torch.empty_like(x_hat) if xc is None else xc
        loss = torch.empty(B, dtype=torch.float64, device=x_hat.device) if loss is None else loss
        nb = (3 * x_hat.numel() + y.numel()) * 4          # read x_hat + y, write g + clamped image
        yp = self._prepared(y, B)
        with kernels._timed("op_loss_grad_" + type(self).__name__, self.loss_grad_launches, nb):
            if yp is not None:
                check(lib().nhmc_op_loss_grad_prepared(ctypes.byref(self.desc), x_hat.data_ptr(), yp.data_ptr(), g.data_ptr(),
                                                       xc.data_ptr(), loss.data_ptr(), B, self._workspace(B).data_ptr(),
                                                       kernels._stream()), "nhmc_op_loss_grad_prepared")
            else:
                check(lib().nhmc_op_loss_grad(ctypes.byref(self.desc), x_hat.data_ptr(), y.data_ptr(), g.data_ptr(),
                                              xc.data_ptr(), loss.data_ptr(), B, self._workspace(B).data_ptr(),
                                              kernels._stream()), "nhmc_op_loss_grad")
        return loss, g, xc

    def _prepared(self, y, B):
        """Operators with a prepared measurement form (the spectral deblurring operators: nhmc_op_prepare_measurement) keep
        it per measurement tensor (storage, version): a sampler evaluates thousands of times against one y."""
        if not getattr(self, "has_prepared_measurement", False) or getattr(self, "_tc_tables", None) is None:
            return None
        key = (y.data_ptr(), y._version, tuple(y.shape))
        c = self.__dict__.get("_prep")
        if c is None or c[0] != key:
            if torch.cuda.is_current_stream_capturing():
                return None                                  # a new measurement inside a capture: the self-contained entry point
            yp = torch.empty_like(y)
            check(lib().nhmc_op_prepare_measurement(ctypes.byref(self.desc), y.data_ptr(), yp.data_ptr(), B,
                                                    self._workspace(B).data_ptr(), kernels._stream()), "nhmc_op_prepare_measurement")
            c = self.__dict__["_prep"] = (key, yp)
        return c[1]


class Inpainting(Operator):
    """Row-selection operator: y[b, 3k+c] = x[b, c, kept_pix[k]] (Hfuncs.py:119-154)."""
    kind = _lib.OP_INPAINT

    def __init__(self, channels, img_dim, missing_indices, device):
        """`missing_indices`: the reference's channel-interleaved index vector (3*pix + c), main_sampling.py:302-305.

        The reference spends 44 s in a python `i not in tensor` loop here (Hfuncs.py:125); a boolean mask gives
        the same ascending kept list."""
        super().__init__(channels, img_dim, device)
        P = img_dim ** 2
        mi = torch.as_tensor(missing_indices).detach().cpu().numpy().astype(np.int64)
        miss = np.zeros(P, dtype=bool)
        miss[np.unique(mi // channels)] = True
        kept = np.nonzero(~miss)[0].astype(np.int32)
        p2k = np.full(P, -1, dtype=np.int32)
        p2k[kept] = np.arange(len(kept), dtype=np.int32)
        self.missing_indices = torch.as_tensor(missing_indices).to(self.device)
        self.kept_pix = torch.from_numpy(kept).to(self.device)
        self.pix_to_kept = torch.from_numpy(p2k).to(self.device)
        # the reference's attribute, for callers that read it (channel-interleaved, ascending)
        self.kept_indices = (channels * self.kept_pix.long()[:, None] + torch.arange(channels, device=self.device)[None]).reshape(-1)
        self.desc.n_kept = len(kept)
        self.desc.kept_pix = self.kept_pix.data_ptr()
        self.desc.pix_to_kept = self.pix_to_kept.data_ptr()

    @classmethod
    def box(cls, channels, img_dim, device, left=None, up=None, size=128):
        """Box inpainting (main_sampling.py:290-299); left/up default to the reference's random.randint(16, 112) draws."""
        import random
        left = random.randint(16, 112) if left is None else left
        up = random.randint(16, 112) if up is None else up
        return cls(channels, img_dim, box_missing_indices(channels, img_dim, left, up, size), device)

    @classmethod
    def random(cls, channels, img_dim, device, frac=0.92):
        """main_sampling.py:302-305: randperm on the global CPU generator, first 92 % of the pixels dropped."""
        missing_r = channels * torch.randperm(img_dim ** 2)[: int(img_dim ** 2 * frac)].long()
        missing = torch.cat([missing_r + c for c in range(channels)], dim=0)
        return cls(channels, img_dim, missing, device)


def box_missing_indices(channels, img_dim, left, up, size=128):
    """main_sampling.py:290-299 (`inp` + `box` in --deg): a size x size box whose first image row is `left` and first
    column is `up` (the reference's names), as channel-interleaved indices (3*pix + c), ascending.  The reference
    draws left, up = random.randint(16, 112) twice from python's `random` (seeded with --seed)."""
    m = torch.zeros(img_dim, img_dim, channels)
    m[left:left + size, up:up + size, :] = 1.0
    return torch.nonzero(m.view(-1)).squeeze(1)


class SuperResolution(Operator):
    """ratio x ratio block average in CHW order (Hfuncs.py:180-234)."""
    kind = _lib.OP_SR

    def __init__(self, channels, img_dim, ratio, device):
        assert img_dim % ratio == 0
        super().__init__(channels, img_dim, device)
        self.ratio = int(ratio)
        self.y_dim = img_dim // ratio
        self.desc.ratio = self.ratio


class HDR(Operator):
    """y = clip(x/0.5, -1, 1) (Hfuncs.py:406-445); image-shaped measurement."""
    kind = _lib.OP_HDR
    linear = False

    def __init__(self, channels=3, img_dim=256, device="cuda"):
        super().__init__(channels, img_dim, device)

    def meas_shape(self, B):
        return (B, self.channels, self.img_dim, self.img_dim)

    def H_pinv(self, vec):
        return vec


def _band_matrix(kernel, img_dim):
    """Hfuncs.py:460-471: the loop `range(i - k//2, i + k//2)` drops the last tap."""
    k = torch.as_tensor(kernel).detach().cpu().float()
    h = k.shape[0] // 2
    M = torch.zeros(img_dim, img_dim)
    for d in range(-h, h):
        v = k[d + h]
        idx = torch.arange(max(0, -d), min(img_dim, img_dim - d))
        M[idx, idx + d] = v
    return M


def gaussian_taps(sigma, half=4):
    """Normalised Gaussian taps as main_sampling.py:327-335 builds them (fp32 exp, then / sum)."""
    pdf = lambda v: torch.exp(torch.Tensor([-0.5 * (v / sigma) ** 2]))
    k = torch.Tensor([pdf(i) for i in range(-half, half + 1)])
    return k / k.sum()


class Deblurring2D(Operator):
    """Y_c = U1 (G_c .* (V1^T X_c V2)) U2^T (Hfuncs.py:448-523, tiled-singulars quirk included)."""
    kind = _lib.OP_DEBLUR2D
    loss_grad_launches = 6     # split of clamp(x_hat) + 4 tensor-core products (prepared measurement) + the clamped copy
    has_prepared_measurement = True

    def __init__(self, kernel1, kernel2, channels, img_dim, device, tables=None, zero=3e-2):
        super().__init__(channels, img_dim, device)
        if tables is None:
            U1, s1, V1 = torch.svd(_band_matrix(kernel1, img_dim), some=False)
            U2, s2, V2 = torch.svd(_band_matrix(kernel2, img_dim), some=False)
            s1 = s1.clone(); s2 = s2.clone()
            s1[s1 < zero] = 0
            s2[s2 < zero] = 0
            s = torch.matmul(s1.reshape(img_dim, 1), s2.reshape(1, img_dim)).reshape(img_dim ** 2)
            s_sorted, perm = s.sort(descending=True)
        else:
            U1, V1, U2, V2, s_sorted, perm = [torch.as_tensor(t).detach().cpu() for t in tables]
        n = img_dim ** 2
        rank = torch.empty(n, dtype=torch.long)
        rank[perm.long()] = torch.arange(n)
        # Hfuncs.py:519-520 tiles the sorted singulars while Vt (:493-499) interleaves channels
        gain = torch.stack([s_sorted[(channels * rank + c) % n] for c in range(channels)], 0).float()
        gpinv = torch.where(gain != 0, 1.0 / torch.where(gain != 0, gain, torch.ones_like(gain)), torch.zeros_like(gain))
        dev = self.device
        self.U_small1, self.V_small1 = U1.float().contiguous().to(dev), V1.float().contiguous().to(dev)
        self.U_small2, self.V_small2 = U2.float().contiguous().to(dev), V2.float().contiguous().to(dev)
        self._singulars, self._perm = s_sorted.to(dev), perm.to(dev)
        self.gain, self.gain_pinv = gain.contiguous().to(dev), gpinv.contiguous().to(dev)
        d = self.desc
        d.U1, d.V1, d.U2, d.V2 = (t.data_ptr() for t in (self.U_small1, self.V_small1, self.U_small2, self.V_small2))
        d.gain, d.gain_pinv = self.gain.data_ptr(), self.gain_pinv.data_ptr()
        self._prepare_tables()


class Deblurring(Deblurring2D):
    """Separable (Gaussian / uniform) deblurring, Hfuncs.py:236-316: the same spectral form as Deblurring2D with one
    1-D kernel on both axes -- Y_c = U (G_c .* (V^T X_c V)) U^T -- including the same tiled-singulars quirk
    (`singulars()` = `_singulars.repeat(1, 3)`, :293-294, against the channel-interleaved `Vt`, :266-272)."""

    def __init__(self, kernel, channels, img_dim, device, tables=None, zero=3e-2):
        if tables is not None:
            U, V, s_sorted, perm = tables
            tables = (U, V, U, V, s_sorted, perm)
        super().__init__(kernel, kernel, channels, img_dim, device, tables=tables, zero=zero)
        self.U_small, self.V_small = self.U_small1, self.V_small1


def bicubic_taps(factor, a=-0.5):
    """The 4*factor bicubic taps of main_sampling.py:266-281 (float64 numpy, normalised, then float32 and normalised again)."""
    def bk(x):
        x = abs(x)
        if x <= 1:
            return (a + 2) * x ** 3 - (a + 3) * x ** 2 + 1
        if 1 < x < 2:
            return a * x ** 3 - 5 * a * x ** 2 + 8 * a * x - 4 * a
        return 0
    k = np.zeros((factor * 4))
    for i in range(factor * 4):
        k[i] = bk((1 / factor) * (i - np.floor(factor * 4 / 2) + 0.5))
    k = k / np.sum(k)
    kernel = torch.from_numpy(k).float()
    return kernel / kernel.sum()


def _srconv_matrix(kernel, img_dim, stride):
    """1-D strided convolution matrix with reflective padding (Hfuncs.py:541-551)."""
    k = torch.as_tensor(kernel).detach().cpu().float()
    small = img_dim // stride
    Hs = torch.zeros(small, img_dim)
    half = k.shape[0] // 2
    for i in range(stride // 2, img_dim + stride // 2, stride):
        for j in range(i - half, i + half):
            je = j
            if je < 0:
                je = -je - 1
            if je >= img_dim:
                je = (img_dim - 1) - (je - img_dim)
            Hs[i // stride, je] += k[j - i + half]
    return Hs


class SRConv(Operator):
    """Convolution-based (bicubic) super-resolution, Hfuncs.py:527-607: Y_c = L X_c L^T with L = U_small diag(s)
    V_small[:, :d]^T (d = D / stride; the permutation of :558-559 selects the leading d x d block of V^T X V),
    H^T R = L^T R L, H^+ Y = P Y P^T with P = V_small[:, :d] diag(1/s) U_small^T.  L and P are formed once from the SVD
    factors (the reference's own torch.svd call, or taken from a reference object)."""
    kind = _lib.OP_SEPMAT
    loss_grad_launches = 5

    def __init__(self, kernel, channels, img_dim, device, stride=1, factors=None, zero=3e-2):
        super().__init__(channels, img_dim, device)
        self.ratio, self.small_dim = int(stride), img_dim // int(stride)
        if factors is None:
            U, S, V = torch.svd(_srconv_matrix(kernel, img_dim, self.ratio), some=False)
            S = S.clone()
            S[S < zero] = 0
        else:
            U, S, V = [torch.as_tensor(t).detach().cpu().float() for t in factors]
        d = self.small_dim
        Sinv = torch.where(S != 0, 1.0 / torch.where(S != 0, S, torch.ones_like(S)), torch.zeros_like(S))
        Vd = V[:, :d]
        self.U_small, self.singulars_small, self.V_small = U.to(self.device), S.to(self.device), V.to(self.device)
        self.L = ((U * S[None, :]) @ Vd.t()).contiguous().to(self.device)            # [d, D]
        self.P = ((Vd * Sinv[None, :]) @ U.t()).contiguous().to(self.device)         # [D, d]
        self.desc.ratio = self.ratio
        self.desc.U1, self.desc.V1 = self.L.data_ptr(), self.P.data_ptr()
        self._prepare_tables()


class WalshHadamardCS(Operator):
    """Walsh-Hadamard compressive sensing, Hfuncs.py:611-651 (--deg cs<ratio>): y[b, k*C + c] = FWHT(x_c)[perm[k]] / D for
    the first D*D/ratio entries of a pixel permutation; all singular values are 1, so H^+ = H^T = FWHT(scatter(.))."""
    kind = _lib.OP_CS
    loss_grad_launches = 7

    def __init__(self, channels, img_dim, ratio, perm, device):
        super().__init__(channels, img_dim, device)
        self.ratio = int(ratio)
        self.perm = torch.as_tensor(perm).to(self.device).long()
        assert self.perm.numel() == img_dim ** 2
        self._perm32 = self.perm.to(torch.int32).contiguous()
        self.desc.ratio = self.ratio
        self.desc.n_kept = img_dim ** 2 // self.ratio
        self.desc.kept_pix = self._perm32.data_ptr()

    @classmethod
    def random(cls, channels, img_dim, ratio, device):
        """main_sampling.py:322-325: torch.randperm(D*D) drawn on the DEVICE generator."""
        return cls(channels, img_dim, ratio, torch.randperm(img_dim ** 2, device=device), device)


class Denoising(Operator):
    """H = identity on the CHW-flattened image (Hfuncs.py:157-177)."""
    kind = _lib.OP_DENOISE


class Colorization(Operator):
    """y[b, pix] = grey value of the pixel (Hfuncs.py:655-695): H = U S V^T of the 1x3 matrix [0.3333, 0.3334, 0.3333]
    applied per pixel.  The three weight triples handed to the kernels are products of the SVD factors, which are
    built with the reference's own torch.svd call (or taken from a reference object)."""
    kind = _lib.OP_COLOR

    def __init__(self, img_dim, device, factors=None):
        super().__init__(3, img_dim, device)
        if factors is None:
            Hs = torch.Tensor([[0.3333, 0.3334, 0.3333]])
            U, S, V = torch.svd(Hs, some=False)
        else:
            U, S, V = [torch.as_tensor(t).detach().cpu().float() for t in factors]
        self.U_small, self.singulars_small, self.V_small = U.to(self.device), S.to(self.device), V.to(self.device)
        u, sv = float(U[0, 0]), float(S[0])
        v0 = [float(V[c, 0]) for c in range(3)]
        f32 = lambda t: float(np.float32(t))
        for c in range(3):
            self.desc.coef[c] = f32(np.float32(u) * (np.float32(sv) * np.float32(v0[c])))            # H:   U (s (V^T x)_0)
            self.desc.coef[3 + c] = f32(np.float32(v0[c]) * (np.float32(sv) * np.float32(u)))        # H^T: V (s (U^T r))
            self.desc.coef[6 + c] = f32(np.float32(v0[c]) * (np.float32(u) / np.float32(sv)))        # H^+: V ((U^T y) / s)

    def meas_shape(self, B):
        return (B, self.img_dim ** 2)


class PhaseRetrievalOperator(Operator):
    """y = |fft2c(pad(x, 64))| (Hfuncs.py:318-366, fastmri_utils.py:67-89); image-shaped measurement."""
    kind = _lib.OP_PHASE
    linear = False

    def __init__(self, oversample, device, channels=3, img_dim=256):
        super().__init__(channels, img_dim, device)
        self.pad = int((oversample / 8.0) * 256)
        self.desc.pad = self.pad
        self._prepare_tables()           # FFT twiddle table, built once

    def meas_shape(self, B):
        L = self.img_dim + 2 * self.pad
        return (B, self.channels, L, L)

    def H_pinv(self, vec):
        out = super().H_pinv(vec)
        return out.reshape(vec.shape[0], self.channels, self.img_dim, self.img_dim)


def from_reference(H_funcs, channels, img_dim, device):
    """Build the device operator from one of the reference's own `H_funcs` objects (duck-typed by
    attribute, so the reference module does not have to be importable)."""
    if isinstance(H_funcs, Operator):
        return H_funcs
    name = type(H_funcs).__name__
    if hasattr(H_funcs, "kept_indices") and hasattr(H_funcs, "missing_indices"):
        return Inpainting(getattr(H_funcs, "channels", channels), getattr(H_funcs, "img_dim", img_dim),
                          H_funcs.missing_indices, device)
    if hasattr(H_funcs, "U_small1") and hasattr(H_funcs, "_perm"):
        return Deblurring2D(None, None, H_funcs.channels, H_funcs.img_dim, device,
                            tables=(H_funcs.U_small1, H_funcs.V_small1, H_funcs.U_small2, H_funcs.V_small2,
                                    H_funcs._singulars, H_funcs._perm))
    if name == "Deblurring" or (hasattr(H_funcs, "U_small") and hasattr(H_funcs, "_perm") and hasattr(H_funcs, "singulars_small")
                                and not hasattr(H_funcs, "small_dim") and tuple(H_funcs.U_small.shape) == (img_dim, img_dim)):
        return Deblurring(None, H_funcs.channels, H_funcs.img_dim, device,
                          tables=(H_funcs.U_small, H_funcs.V_small, H_funcs._singulars, H_funcs._perm))
    if name == "SRConv" or (hasattr(H_funcs, "small_dim") and hasattr(H_funcs, "U_small") and hasattr(H_funcs, "singulars_small")):
        return SRConv(None, H_funcs.channels, H_funcs.img_dim, device, stride=H_funcs.ratio,
                      factors=(H_funcs.U_small, H_funcs.singulars_small, H_funcs.V_small))
    if name == "WalshHadamardCS" or (hasattr(H_funcs, "perm") and hasattr(H_funcs, "fwht")):
        return WalshHadamardCS(H_funcs.channels, H_funcs.img_dim, H_funcs.ratio, H_funcs.perm, device)
    if name == "Colorization" or hasattr(H_funcs, "Vt_small"):
        return Colorization(H_funcs.img_dim, device, factors=(H_funcs.U_small, H_funcs.singulars_small, H_funcs.V_small))
    if name == "Denoising":
        return Denoising(channels, img_dim, device)
    if hasattr(H_funcs, "ratio") and hasattr(H_funcs, "y_dim") and hasattr(H_funcs, "V_small"):
        return SuperResolution(H_funcs.channels, H_funcs.img_dim, H_funcs.ratio, device)
    if name == "HDR":
        return HDR(channels, img_dim, device)
    if name == "PhaseRetrievalOperator" or (hasattr(H_funcs, "pad") and hasattr(H_funcs, "undo_padding")):
        op = PhaseRetrievalOperator(2.0, device, channels, img_dim)
        op.pad = int(H_funcs.pad); op.desc.pad = op.pad
        return op
    raise _lib.NhmcError("operator %s is not on the accelerated path (inpaint, sr, hdr, deblur_aniso, deblur_gauss, sr_bicubic, cs, phase, color, deno)" % name)
