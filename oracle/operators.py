"""Measurement operators A, A^T, A^+ in closed form (oracle; test infrastructure).

Each class restates one operator of /root/reference/obs_functions/Hfuncs.py without the SVD
scaffolding (H = U * S * V^T there); the closed forms were checked against the reference's
own `H`, `H_pinv` and autograd gradients by tests/golden/make_golden.py.

Common interface (x is [B,C,D,D], any float dtype; CPU torch tensors):
    H(x)        -> measurement in the reference's own layout (flat [B,M] or image-shaped)
    residual(x_hat, y) -> (r, loss[B])  with r = y - H(clamp(x_hat,-1,1)), loss = sum r^2 per chain
    grad(x_hat, y)     -> d loss / d x_hat  (includes the final-clip mask, main_sampling.py:709)
    H_pinv(y)   -> [B, C*D*D] (flat, as the reference returns it) or image-shaped for hdr/phase
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _inrange(x, lo=-1.0, hi=1.0):
    # torch.clamp's backward passes the gradient where lo <= x <= hi (inclusive).
    return ((x >= lo) & (x <= hi)).to(x.dtype)


class _Base:
    linear = True

    def is_linear(self):
        return self.linear

    def residual(self, x_hat, y):
        xc = x_hat.clamp(-1, 1)
        r = y - self.H(xc)
        loss = (r.double() ** 2).reshape(r.shape[0], -1).sum(1)
        return r, loss

    def grad(self, x_hat, y):
        """d/dx_hat sum((y - H(clamp(x_hat)))^2), per chain (block-diagonal)."""
        r, _ = self.residual(x_hat, y)
        return -2.0 * self.Ht(r).reshape(x_hat.shape) * _inrange(x_hat)


# --------------------------------------------------------------------------------------
# Random inpainting: Hfuncs.py:119-154 + mask construction main_sampling.py:301-305
# --------------------------------------------------------------------------------------
def random_inpaint_missing_pixels(img_dim, frac=0.92, generator=None):
    """Pixel indices (row-major, 0..D*D-1) that are DROPPED; main_sampling.py:302.

    The reference draws `torch.randperm(D*D)` from the global CPU generator and keeps the
    first int(D*D*0.92) entries as missing pixels (times 3, +0/+1/+2 for the channels of
    the HWC-interleaved vector).
    """
    perm = torch.randperm(img_dim ** 2, generator=generator)
    return perm[: int(img_dim ** 2 * frac)].clone()


class Inpainting(_Base):
    """y[b, 3k+c] = x[b, c, kept_pix[k]], kept pixels ascending (Hfuncs.py:119-154).

    In the reference `Vt` flattens the image pixel-major/channel-minor (permute(0,2,1)),
    gathers `kept_indices` (ascending, Hfuncs.py:125) to the front and all singular values
    are 1, `U` is the identity -> H is a pure gather.
    """

    def __init__(self, channels, img_dim, missing_pixels):
        self.channels, self.img_dim = channels, img_dim
        miss = np.zeros(img_dim ** 2, dtype=bool)
        miss[np.asarray(missing_pixels, dtype=np.int64)] = True
        self.kept_pix = torch.from_numpy(np.nonzero(~miss)[0].astype(np.int64))

    @classmethod
    def from_missing_indices(cls, channels, img_dim, missing_indices):
        """From the reference's channel-interleaved index vector (3*pix + c)."""
        mi = np.asarray(missing_indices, dtype=np.int64)
        return cls(channels, img_dim, np.unique(mi // channels))

    def H(self, x):
        B = x.shape[0]
        xf = x.reshape(B, self.channels, -1)
        return xf[:, :, self.kept_pix].permute(0, 2, 1).reshape(B, -1)

    def Ht(self, r):
        B = r.shape[0]
        out = torch.zeros(B, self.channels, self.img_dim ** 2, dtype=r.dtype)
        out[:, :, self.kept_pix] = r.reshape(B, -1, self.channels).permute(0, 2, 1)
        return out.reshape(B, -1)

    def H_pinv(self, y):
        return self.Ht(y)


# --------------------------------------------------------------------------------------
# Block-average super-resolution: Hfuncs.py:180-234
# --------------------------------------------------------------------------------------
class SuperResolution(_Base):
    """y[b, c*yd*yd + i*yd + j] = mean of the ratio x ratio block (Hfuncs.py:180-234).

    H_small = [1/r^2]*r^2 has the single singular value 1/r with V[:,0] = +-1/r, U = +-1 (the
    two signs cancel), so H = r x r average pooling in CHW order, H^T = nearest-upsample / r^2
    and H^+ = nearest-upsample.
    """

    def __init__(self, channels, img_dim, ratio):
        assert img_dim % ratio == 0
        self.channels, self.img_dim, self.ratio = channels, img_dim, ratio
        self.y_dim = img_dim // ratio

    def H(self, x):
        B = x.shape[0]
        return F.avg_pool2d(x.reshape(B, self.channels, self.img_dim, self.img_dim), self.ratio).reshape(B, -1)

    def _up(self, y):
        B = y.shape[0]
        y = y.reshape(B, self.channels, self.y_dim, self.y_dim)
        return y.repeat_interleave(self.ratio, 2).repeat_interleave(self.ratio, 3)

    def Ht(self, r):
        return (self._up(r) / self.ratio ** 2).reshape(r.shape[0], -1)

    def H_pinv(self, y):
        return self._up(y).reshape(y.shape[0], -1)


# --------------------------------------------------------------------------------------
# HDR: Hfuncs.py:406-445
# --------------------------------------------------------------------------------------
class HDR(_Base):
    """y = clip(x / 0.5, -1, 1) (Hfuncs.py:415-419); d/dx = 2 * 1[-1 <= 2x <= 1]."""
    linear = False

    def H(self, x):
        return torch.clip(x / 0.5, -1, 1)

    def grad(self, x_hat, y):
        r, _ = self.residual(x_hat, y)
        xc = x_hat.clamp(-1, 1)
        return -2.0 * r * 2.0 * _inrange(xc / 0.5) * _inrange(x_hat)

    def H_pinv(self, y):
        return y


# --------------------------------------------------------------------------------------
# Anisotropic deblurring: Hfuncs.py:448-523 + kernels main_sampling.py:327-335
# --------------------------------------------------------------------------------------
def gaussian_taps(sigma, half=4):
    """Normalised 2*half+1 Gaussian taps exactly as main_sampling.py:329-334 builds them."""
    pdf = lambda v: torch.exp(torch.Tensor([-0.5 * (v / sigma) ** 2]))
    k = torch.Tensor([pdf(i) for i in range(-half, half + 1)])
    return k / k.sum()


def band_matrix(kernel, img_dim):
    """1-D blur matrix of Hfuncs.py:460-471; the loop `range(i-k//2, i+k//2)` drops the last tap."""
    h = kernel.shape[0] // 2
    M = torch.zeros(img_dim, img_dim)
    for i in range(img_dim):
        for j in range(i - h, i + h):
            if 0 <= j < img_dim:
                M[i, j] = kernel[j - i + h]
    return M


def deblur2d_tables(kernel1, kernel2, img_dim, zero=3e-2):
    """SVD tables of Hfuncs.py:473-481: (U1, V1, U2, V2, singulars_sorted, perm).

    `torch.svd` signs and the unstable descending sort over tied zeros are implementation
    defined: always build ONE set of tables and hand the same tensors to every
    implementation under comparison.
    """
    U1, s1, V1 = torch.svd(band_matrix(kernel1, img_dim), some=False)
    U2, s2, V2 = torch.svd(band_matrix(kernel2, img_dim), some=False)
    s1 = s1.clone(); s2 = s2.clone()
    s1[s1 < zero] = 0
    s2[s2 < zero] = 0
    s = torch.matmul(s1.reshape(img_dim, 1), s2.reshape(1, img_dim)).reshape(img_dim ** 2)
    s_sorted, perm = s.sort(descending=True)
    return U1, V1, U2, V2, s_sorted, perm


def deblur2d_spectral_gain(s_sorted, perm, channels=3):
    """D[c, j] = s_sorted[(channels*rank(j) + c) mod D*D]  with rank(perm[k]) = k.

    Hfuncs.py:519-520: `singulars()` TILES the sorted vector (`repeat(1,3)`) while `Vt`
    (Hfuncs.py:493-499) interleaves channels (index channels*k + c), so spectral coefficient
    `perm[k]` of channel c is scaled by s_sorted[(channels*k + c) mod n]  (SURVEY.md headline 6).
    """
    n = s_sorted.shape[0]
    rank = torch.empty(n, dtype=torch.long)
    rank[perm] = torch.arange(n)
    D = torch.stack([s_sorted[(channels * rank + c) % n] for c in range(channels)], 0)
    return D


class Deblurring2D(_Base):
    """Y_c = U1 (D_c . (V1^T X_c V2)) U2^T, flattened CHW (Hfuncs.py:448-523)."""

    def __init__(self, channels, img_dim, U1, V1, U2, V2, s_sorted, perm):
        self.channels, self.img_dim = channels, img_dim
        self.U1, self.V1, self.U2, self.V2 = U1, V1, U2, V2
        self.s_sorted, self.perm = s_sorted, perm
        self.D = deblur2d_spectral_gain(s_sorted, perm, channels).reshape(channels, img_dim, img_dim)

    @classmethod
    def from_kernels(cls, kernel1, kernel2, channels, img_dim):
        return cls(channels, img_dim, *deblur2d_tables(kernel1, kernel2, img_dim))

    def _cast(self, dt):
        return [t.to(dt) for t in (self.U1, self.V1, self.U2, self.V2, self.D)]

    def H(self, x):
        B = x.shape[0]
        U1, V1, U2, V2, D = self._cast(x.dtype)
        X = x.reshape(B, self.channels, self.img_dim, self.img_dim)
        T = V1.t() @ X @ V2
        return (U1 @ (D * T) @ U2.t()).reshape(B, -1)

    def Ht(self, r):
        B = r.shape[0]
        U1, V1, U2, V2, D = self._cast(r.dtype)
        R = r.reshape(B, self.channels, self.img_dim, self.img_dim)
        T = U1.t() @ R @ U2
        return (V1 @ (D * T) @ V2.t()).reshape(B, -1)

    def H_pinv(self, y):
        B = y.shape[0]
        U1, V1, U2, V2, D = self._cast(y.dtype)
        Dinv = torch.where(D != 0, 1.0 / torch.where(D != 0, D, torch.ones_like(D)), torch.zeros_like(D))
        R = y.reshape(B, self.channels, self.img_dim, self.img_dim)
        T = U1.t() @ R @ U2
        return (V1 @ (Dinv * T) @ V2.t()).reshape(B, -1)


def deblur1d_tables(kernel, img_dim, zero=3e-2):
    """SVD tables of the separable `Deblurring` (Hfuncs.py:245-262): (U, V, singulars_sorted, perm)."""
    U, s, V = torch.svd(band_matrix(kernel, img_dim), some=False)
    s = s.clone()
    s[s < zero] = 0
    ss = torch.matmul(s.reshape(img_dim, 1), s.reshape(1, img_dim)).reshape(img_dim ** 2)
    s_sorted, perm = ss.sort(descending=True)
    return U, V, s_sorted, perm


class Deblurring(Deblurring2D):
    """Separable deblurring (`deblur_gauss`, Hfuncs.py:236-316): Deblurring2D with the same factors on both axes;
    `singulars()` (:293-294) tiles exactly as Deblurring2D's does, so the same D_c gain applies."""

    def __init__(self, channels, img_dim, U, V, s_sorted, perm):
        super().__init__(channels, img_dim, U, V, U, V, s_sorted, perm)

    @classmethod
    def from_kernel(cls, kernel, channels, img_dim):
        return cls(channels, img_dim, *deblur1d_tables(kernel, img_dim))


# --------------------------------------------------------------------------------------
# Denoising (identity): Hfuncs.py:157-177
# --------------------------------------------------------------------------------------
class Denoising(_Base):
    def H(self, x):
        return x.reshape(x.shape[0], -1)

    def Ht(self, r):
        return r.reshape(r.shape[0], -1)

    def H_pinv(self, y):
        return y.reshape(y.shape[0], -1)


# --------------------------------------------------------------------------------------
# Colorization: Hfuncs.py:655-695
# --------------------------------------------------------------------------------------
class Colorization(_Base):
    """y[b, pix] = U00 * s0 * sum_c V[c,0] x[b,c,pix]: `Vt` (:672-679) rotates every pixel's RGB needle by V^T and
    stores component-major, `H` keeps the first D*D entries (component 0) times the single singular value, `U` is the
    1x1 factor.  H^T r = V[:,0] s0 U00 r, H^+ y = V[:,0] U00 y / s0."""

    def __init__(self, img_dim, U=None, S=None, V=None):
        self.img_dim = img_dim
        if U is None:
            U, S, V = torch.svd(torch.Tensor([[0.3333, 0.3334, 0.3333]]), some=False)
        self.u, self.s, self.v0 = U[0, 0], S[0], V[:, 0]

    def H(self, x):
        B = x.shape[0]
        xf = x.reshape(B, 3, -1)
        v0 = self.v0.to(x.dtype)
        return self.u.to(x.dtype) * (self.s.to(x.dtype) * (v0[None, :, None] * xf).sum(1))

    def Ht(self, r):
        B = r.shape[0]
        k = (self.v0 * self.s * self.u).to(r.dtype)
        return (k[None, :, None] * r.reshape(B, 1, -1)).reshape(B, -1)

    def H_pinv(self, y):
        B = y.shape[0]
        k = (self.v0 * self.u / self.s).to(y.dtype)
        return (k[None, :, None] * y.reshape(B, 1, -1)).reshape(B, -1)


# --------------------------------------------------------------------------------------
# Convolution-based (bicubic) super-resolution: Hfuncs.py:527-607, kernel main_sampling.py:266-284
# --------------------------------------------------------------------------------------
def bicubic_taps(factor, a=-0.5):
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


def srconv_matrix(kernel, img_dim, stride):
    """Hfuncs.py:541-551: strided 1-D convolution matrix [D/stride, D] with reflective padding."""
    small = img_dim // stride
    Hs = torch.zeros(small, img_dim)
    half = kernel.shape[0] // 2
    for i in range(stride // 2, img_dim + stride // 2, stride):
        for j in range(i - half, i + half):
            je = j
            if je < 0:
                je = -je - 1
            if je >= img_dim:
                je = (img_dim - 1) - (je - img_dim)
            Hs[i // stride, je] += kernel[j - i + half]
    return Hs


class SRConv(_Base):
    """Y_c = L X_c L^T, L = U diag(s) V[:, :d]^T: `Vt` (:571-578) forms V^T X V and moves its leading d x d block (the
    `_perm` of :558-559) to the front, `singulars()` (:598-599) scales coefficient (i, j) by s_i s_j for every channel,
    `U` (:580-588) maps back with U . U^T."""

    def __init__(self, channels, img_dim, stride, U, S, V):
        self.channels, self.img_dim, self.d = channels, img_dim, img_dim // stride
        Vd = V[:, :self.d]
        Sinv = torch.where(S != 0, 1.0 / torch.where(S != 0, S, torch.ones_like(S)), torch.zeros_like(S))
        self.L = (U * S[None, :]) @ Vd.t()
        self.P = (Vd * Sinv[None, :]) @ U.t()

    @classmethod
    def from_kernel(cls, kernel, channels, img_dim, stride, zero=3e-2):
        U, S, V = torch.svd(srconv_matrix(kernel, img_dim, stride), some=False)
        S = S.clone()
        S[S < zero] = 0
        return cls(channels, img_dim, stride, U, S, V)

    def H(self, x):
        B = x.shape[0]
        L = self.L.to(x.dtype)
        return (L @ x.reshape(B, self.channels, self.img_dim, self.img_dim) @ L.t()).reshape(B, -1)

    def Ht(self, r):
        B = r.shape[0]
        L = self.L.to(r.dtype)
        return (L.t() @ r.reshape(B, self.channels, self.d, self.d) @ L).reshape(B, -1)

    def H_pinv(self, y):
        B = y.shape[0]
        P = self.P.to(y.dtype)
        return (P @ y.reshape(B, self.channels, self.d, self.d) @ P.t()).reshape(B, -1)


# --------------------------------------------------------------------------------------
# Walsh-Hadamard compressive sensing: Hfuncs.py:611-651
# --------------------------------------------------------------------------------------
def fwht(x, img_dim):
    """Natural-order fast Walsh-Hadamard transform over the flattened D*D pixels of every channel, divided by D
    (Hfuncs.py:613-622); the same butterfly order as the reference, so fp32 results are bit-identical."""
    B, C = x.shape[0], x.shape[1]
    a = x.reshape(B, C, img_dim ** 2)
    h = 1
    while h < img_dim ** 2:
        a = a.reshape(B, C, -1, h * 2)
        lo, hi = a[:, :, :, :h], a[:, :, :, h:2 * h]
        a = torch.cat([lo + hi, lo - hi], dim=3)
        h *= 2
    return a.reshape(B, C, img_dim ** 2) / img_dim


class WalshHadamardCS(_Base):
    """y[b, k*C + c] = fwht(x)[b, c, perm[k]], k < D*D/ratio (`Vt` :631-632 + the truncation in `H`, Hfuncs.py:65-71)."""

    def __init__(self, channels, img_dim, ratio, perm):
        self.channels, self.img_dim, self.ratio = channels, img_dim, ratio
        self.perm = torch.as_tensor(perm).long()
        self.n = img_dim ** 2 // ratio

    def H(self, x):
        B = x.shape[0]
        t = fwht(x.reshape(B, self.channels, self.img_dim, self.img_dim), self.img_dim)
        return t[:, :, self.perm[: self.n]].permute(0, 2, 1).reshape(B, -1)

    def Ht(self, r):
        B = r.shape[0]
        temp = torch.zeros(B, self.channels, self.img_dim ** 2, dtype=r.dtype)
        temp[:, :, self.perm[: self.n]] = r.reshape(B, self.n, self.channels).permute(0, 2, 1)
        return fwht(temp, self.img_dim).reshape(B, -1)

    def H_pinv(self, y):
        return self.Ht(y)


def box_missing_indices(channels, img_dim, left, up, size=128):
    """main_sampling.py:290-299: channel-interleaved indices (3*pix + c) of a size x size box at rows left.., cols up.."""
    m = torch.zeros(img_dim, img_dim, channels)
    m[left:left + size, up:up + size, :] = 1.0
    return torch.nonzero(m.view(-1)).squeeze(1)


# --------------------------------------------------------------------------------------
# Phase retrieval: Hfuncs.py:318-366, fft2_m Hfuncs.py:8-12, fft2c_new fastmri_utils.py:67-89
# --------------------------------------------------------------------------------------
class PhaseRetrieval(_Base):
    """y = | fftshift( fft2( ifftshift( pad(x, 64) ), norm='ortho' ) ) |  per channel.

    pad = int(oversample/8 * 256) (Hfuncs.py:321) regardless of the image size.
    """
    linear = False

    def __init__(self, oversample=2.0):
        self.pad = int((oversample / 8.0) * 256)

    def _F(self, x):
        p = self.pad
        xp = F.pad(x, (p, p, p, p))
        z = torch.fft.ifftshift(xp, dim=(-2, -1))
        z = torch.fft.fft2(z, norm="ortho")
        return torch.fft.fftshift(z, dim=(-2, -1))

    def H(self, x):
        return self._F(x).abs()

    def grad(self, x_hat, y):
        """2 Re crop F^H((|Fx|-y) Fx/|Fx|) times the final-clip mask.

        d|z|/dz at z == 0 is taken as 0, which is what torch's abs backward (sgn) does.
        """
        p = self.pad
        xc = x_hat.clamp(-1, 1)
        z = self._F(xc)
        mag = z.abs()
        sgn = torch.where(mag > 0, z / torch.where(mag > 0, mag, torch.ones_like(mag)), torch.zeros_like(z))
        w = (mag - y) * sgn
        w = torch.fft.ifftshift(w, dim=(-2, -1))
        w = torch.fft.ifft2(w, norm="ortho")
        w = torch.fft.fftshift(w, dim=(-2, -1))
        g = 2.0 * w.real[..., p:-p, p:-p]
        return g * _inrange(x_hat)

    def H_pinv(self, y):
        p = self.pad
        z = torch.fft.ifftshift(y.to(torch.complex64 if y.dtype == torch.float32 else torch.complex128), dim=(-2, -1))
        z = torch.fft.ifft2(z, norm="ortho")
        z = torch.fft.fftshift(z, dim=(-2, -1)).abs()
        return z[..., p:-p, p:-p]
