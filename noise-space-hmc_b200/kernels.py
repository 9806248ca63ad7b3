"""Torch-tensor wrappers around the C ABI (raw device pointers + torch's current stream)."""
import threading

import torch

from . import _lib
from ._lib import check, lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}

# ---- launch accounting / optional per-kernel timing (bench.py) -------------------------------------
LAUNCHES = 0          # kernels of libnhmc.so launched so far (python-side count of the C calls' launches)
_PROFILE = None       # when a dict: name -> list of (start_event, end_event, algorithmic_bytes)


def profile_start():
    global _PROFILE
    _PROFILE = {}


def profile_take():
    """Stops recording and returns the raw {name: [(start_event, end_event, algorithmic_bytes), ...]} records WITHOUT
    reading them (for events captured into a CUDA graph: read them with profile_read after each replay)."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    return prof or {}


def profile_read(records, into, by_size=True):
    """Accumulates the current event times of `records` into `into` ({key: (n, ms, bytes)}); synchronises."""
    torch.cuda.synchronize()
    for name, recs in records.items():
        for s, e, b in recs:
            key = (name, b) if by_size else name
            n, ms, tot = into.get(key, (0, 0.0, 0))
            into[key] = (n + 1, ms + s.elapsed_time(e), tot + b)
    return into


def profile_stop(by_size=False):
    """Returns {name: (n_calls, total_ms, total_algorithmic_bytes)} (synchronises); with by_size the key is
    (name, algorithmic_bytes_per_call), i.e. one entry per tensor shape."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for name, recs in (prof or {}).items():
        if not by_size:
            ms = sum(s.elapsed_time(e) for s, e, _ in recs)
            out[name] = (len(recs), ms, sum(b for _, _, b in recs))
            continue
        for s, e, b in recs:
            n, ms, tot = out.get((name, b), (0, 0.0, 0))
            out[(name, b)] = (n + 1, ms + s.elapsed_time(e), tot + b)
    return out


class _timed:
    """Counts launches; with profiling on, brackets the call with CUDA events on the launching stream."""

    def __init__(self, name, n_launches, nbytes):
        self.name, self.n, self.nbytes = name, n_launches, nbytes

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += self.n
        if _PROFILE is not None:
            # inside a stream capture the events become event-record NODES of the graph (external=True): every
            # replay re-records them, so per-kernel CUDA-event times can be read in situ after each replay
            ext = torch.cuda.is_current_stream_capturing()
            self.s = torch.cuda.Event(enable_timing=True, external=ext)
            self.e = torch.cuda.Event(enable_timing=True, external=ext)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.e.record()
            _PROFILE.setdefault(self.name, []).append((self.s, self.e, self.nbytes))
        return False


# ---- GroupNorm workspaces and the per-call flag word ---------------------------------------------------------------
# The library keeps no scratch and no switches of its own (include/nhmc.h): the accumulators / tickets / barrier counters
# of the GroupNorm kernels live in a caller-owned workspace, and the execution-path choices travel with every call.
# Here the "caller" is this module: a thread-local scope (HMCSampler opens one per lane around its launches, so two
# samplers on two host threads never share accumulators) and, outside any scope, one workspace per (device, stream).
GN_MAX_BATCH = 64     # samples one workspace is sized for (nhmc_gn_workspace_bytes); larger batches get a larger one
FLAGS = 0             # process-wide default flag word (tests / A-B tools set it through set_flags); scopes OR theirs in
_tls = threading.local()
_default_ws = {}      # (device index, stream id) -> GnWorkspace


class GnWorkspace:
    """Device scratch of the GroupNorm kernels + a pinned host word for its timeout flag."""

    def __init__(self, device, max_batch=GN_MAX_BATCH):
        self.max_batch = int(max_batch)
        self.nbytes = int(lib().nhmc_gn_workspace_bytes(self.max_batch))
        self.buf = torch.zeros((self.nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._pending = False

    def poll(self):
        """Queue a copy of the timeout flag to pinned host memory on the current stream (no synchronisation)."""
        check(lib().nhmc_gn_check(self.buf.data_ptr(), self.nbytes, self.flag.data_ptr(), _stream()), "nhmc_gn_check")
        self._pending = True

    def raise_if_timed_out(self):
        """After a synchronisation that covers the last poll(): raise (and reset the workspace) if a barrier timed out."""
        if self._pending and int(self.flag[0]) != 0:
            self.buf.zero_()
            self.flag.zero_()
            raise _lib.NhmcError("resident GroupNorm: a per-sample barrier timed out; the results of the affected calls are "
                                 "invalid (workspace reset)")
        self._pending = False


def set_flags(flags):
    """Process-wide default flag word (NHMC_F_NO_PDL, NHMC_GN_F_*); returns the previous one.  For tests and A/B tools."""
    global FLAGS
    prev, FLAGS = FLAGS, int(flags)
    return prev


class gn_scope:
    """with gn_scope(ws, n_lanes, flags): the GroupNorm / pooling / residual-add calls of this thread use `ws` and carry
    `flags` (lane count included) until the scope closes."""

    def __init__(self, ws, n_lanes=1, flags=0):
        self.state = (ws, int(flags) | ((max(1, min(int(n_lanes), 8)) - 1) << _lib.GN_F_LANES_SHIFT))

    def __enter__(self):
        self.prev = getattr(_tls, "scope", None)
        _tls.scope = self.state
        return self

    def __exit__(self, *exc):
        _tls.scope = self.prev
        return False


def current_scope():
    """The calling thread's scope state (or None): autograd Functions save it in forward and re-enter it in backward,
    because autograd runs backward nodes on its own device thread where this thread's scope is not visible."""
    return getattr(_tls, "scope", None)


class _reenter:
    def __init__(self, state):
        self.state = state

    def __enter__(self):
        self.prev = getattr(_tls, "scope", None)
        if self.state is not None:
            _tls.scope = self.state
        return self

    def __exit__(self, *exc):
        _tls.scope = self.prev
        return False


def scope_of(state):
    """Context manager re-entering a scope state saved with current_scope() (no-op for None)."""
    return _reenter(state)


def _flags():
    sc = getattr(_tls, "scope", None)
    return FLAGS | (sc[1] if sc is not None else 0)


def _gn_ws(device, batch):
    sc = getattr(_tls, "scope", None)
    if sc is not None and sc[0] is not None:
        ws = sc[0]
        if batch > ws.max_batch:
            raise _lib.NhmcError("GroupNorm workspace of this scope is sized for %d samples, got %d" % (ws.max_batch, batch))
        return ws
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    ws = _default_ws.get(key)
    if ws is None or ws.max_batch < batch:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.NhmcError("no GroupNorm workspace for this stream yet: warm up once before capturing, or open a kernels.gn_scope")
        ws = _default_ws[key] = GnWorkspace(device, max(GN_MAX_BATCH, batch))
    return ws


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _req(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise _lib.NhmcError("%s must live on a CUDA device: the sm_100a kernels have no CPU fallback" % name)
    if not t.is_contiguous():
        raise _lib.NhmcError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise _lib.NhmcError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def zeros_workspace(nbytes, device):
    """Device scratch, zero-filled once (reduction tickets reset themselves afterwards)."""
    return torch.zeros((int(nbytes) + 7) // 8, dtype=torch.int64, device=device)


def begin_iteration(chains, sched):
    with _timed("begin_iteration", 1, chains.n * 144 * 2):
        check(lib().nhmc_begin_iteration(chains.ptr(), chains.n, sched, _stream()), "nhmc_begin_iteration")


def leapfrog_workspace(n_chains, device):
    return zeros_workspace(lib().nhmc_leapfrog_workspace_bytes(n_chains), device)


def leapfrog(mode, x_in, x_out, p, g_a, g_b, loss, chains, mass, ws):
    B = x_in.shape[0]
    N = x_in[0].numel()
    _req(x_in, torch.float32, "x_in"); _req(p, torch.float32, "p"); _req(g_a, torch.float32, "g_a")
    if x_out is not None: _req(x_out, torch.float32, "x_out")
    if g_b is not None: _req(g_b, torch.float32, "g_b")
    if loss is not None: _req(loss, torch.float64, "loss")
    nb = B * N * 4 * ((3 if mode == _lib.LF_LAST else 5) + (1 if g_b is not None else 0))
    with _timed("leapfrog", 1, nb):
        check(lib().nhmc_leapfrog(mode, _p(x_in), _p(x_out), _p(p), _p(g_a), _p(g_b), _p(loss), chains.ptr(),
                                  float(mass), B, N, _p(ws), _stream()), "nhmc_leapfrog")


def leapfrog_diag(mode, x_in, x_out, p, g_a, g_b, loss, chains, inv_mass, wf_mean, wf_m2, ws):
    """Diagonal-mass leapfrog step (hmc_test_conditioning): per-element inverse mass + the trajectory's Welford mean / M2."""
    B = x_in.shape[0]
    N = x_in[0].numel()
    for t, nm in ((x_in, "x_in"), (p, "p"), (g_a, "g_a"), (inv_mass, "inv_mass")):
        _req(t, torch.float32, nm)
    if g_b is not None: _req(g_b, torch.float32, "g_b")
    nb = B * N * 4 * ((4 if mode == _lib.LF_LAST else 6) + (1 if g_b is not None else 0) + (4 if (wf_mean is not None and mode != _lib.LF_FIRST) else 0))
    with _timed("leapfrog_diag", 1 if mode == _lib.LF_FIRST else 2, nb):
        check(lib().nhmc_leapfrog_diag(mode, _p(x_in), _p(x_out), _p(p), _p(g_a), _p(g_b), _p(loss), chains.ptr(), _p(inv_mass),
                                       _p(wf_mean), _p(wf_m2), B, N, _p(ws), _stream()), "nhmc_leapfrog_diag")


def ddim_mask_words(n_per_chain):
    return lib().nhmc_ddim_mask_bytes(n_per_chain) // 4


def ddim_fwd(xt, eps, xt_next, mask, coef):
    """eps: network output [B, Ce, H, W] (first C channels used), any of f32/bf16/f16."""
    B = xt.shape[0]
    N = xt[0].numel()
    _req(xt, torch.float32, "xt"); _req(xt_next, torch.float32, "xt_next"); _req(eps, None, "eps")
    c1, sa, c2, ss = coef
    with _timed("ddim_fwd", 1, B * N * (8 + eps.element_size()) + B * N // 8):
        check(lib().nhmc_ddim_fwd(_p(xt), _p(eps), _DT[eps.dtype], eps[0].numel(), _p(xt_next), _p(mask),
                                  c1, sa, c2, ss, B, N, _stream()), "nhmc_ddim_fwd")


def ddim_bwd(g_a, g_b, mask, g_xt, g_eps, coef):
    B = g_a.shape[0]
    N = g_a[0].numel()
    _req(g_a, torch.float32, "g_a"); _req(g_xt, torch.float32, "g_xt"); _req(mask, torch.int32, "mask")
    if g_b is not None: _req(g_b, torch.float32, "g_b")
    dt, stride = _lib.F32, N
    if g_eps is not None:
        _req(g_eps, None, "g_eps")
        dt, stride = _DT[g_eps.dtype], g_eps[0].numel()
    c1, sa, c2, ss = coef
    nb = B * N * (8 + (4 if g_b is not None else 0) + (g_eps.element_size() if g_eps is not None else 0)) + B * N // 8
    with _timed("ddim_bwd", 1, nb):
        check(lib().nhmc_ddim_bwd(_p(g_a), _p(g_b), _p(mask), _p(g_xt), _p(g_eps), dt, stride, c1, sa, c2, ss, B, N,
                                  _stream()), "nhmc_ddim_bwd")


def clip_fwd(z, zc, zc_copy=None):
    """zc = clamp(z, -1, 1) (and a second copy), fp32."""
    _req(z, torch.float32, "z"); _req(zc, torch.float32, "zc")
    with _timed("clip", 1, z.numel() * 4 * (2 + (1 if zc_copy is not None else 0))):
        check(lib().nhmc_clip_fwd(_p(z), _p(zc), _p(zc_copy), z.numel(), _stream()), "nhmc_clip_fwd")


def clip_bwd(g, z, out):
    """out = g * 1[-1 <= z <= 1]; g in any network dtype, out fp32."""
    _req(g, None, "g"); _req(z, torch.float32, "z"); _req(out, torch.float32, "out")
    with _timed("clip", 1, z.numel() * (8 + g.element_size())):
        check(lib().nhmc_clip_bwd(_p(g), _DT[g.dtype], _p(z), _p(out), z.numel(), _stream()), "nhmc_clip_bwd")


def accept(chains, u, mass, sched, n_slots):
    _req(u, torch.float32, "u")
    with _timed("accept", 1, chains.n * 144 * 2):
        check(lib().nhmc_accept(chains.ptr(), _p(u), float(mass), sched, n_slots, chains.n, _stream()), "nhmc_accept")


def commit(chains, x, x_prop, x_hat, x_keep, samples, sched, n_slots):
    N = x[0].numel()
    for t, n in ((x, "x"), (x_prop, "x_prop"), (x_hat, "x_hat")):
        _req(t, torch.float32, n)
    with _timed("commit", 1, chains.n * N * 4 * 4):
        check(lib().nhmc_commit(chains.ptr(), _p(x), _p(x_prop), _p(x_hat), _p(x_keep), _p(samples), sched, n_slots,
                                chains.n, N, _stream()), "nhmc_commit")


def _film_stride(scale, shift, channels):
    """scale / shift: fp32 [B, C] rows, unit stride along C, possibly views into one [B, 2C] projection."""
    if scale is None:
        return 0
    for t, n in ((scale, "scale"), (shift, "shift")):
        if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != channels or t.stride(1) != 1:
            raise _lib.NhmcError("%s must be fp32 [B, C] with unit stride along C" % n)
    if scale.stride(0) != shift.stride(0):
        raise _lib.NhmcError("scale and shift must share their row stride")
    return scale.stride(0)


def commit_grad(chains, g_state, g_a, g_b, loss_state, loss_prop):
    N = g_state[0].numel()
    _req(g_state, torch.float32, "g_state"); _req(g_a, torch.float32, "g_a")
    _req(loss_state, torch.float64, "loss_state"); _req(loss_prop, torch.float64, "loss_prop")
    if g_b is not None: _req(g_b, torch.float32, "g_b")
    with _timed("commit", 1, chains.n * N * 4 * (2 + (1 if g_b is not None else 0))):
        check(lib().nhmc_commit_grad(chains.ptr(), _p(g_state), _p(g_a), _p(g_b), _p(loss_state), _p(loss_prop), chains.n, N,
                                     _stream()), "nhmc_commit_grad")


def _in_bias_stride(in_bias, channels):
    """in_bias: fp32 [C] (shared) or [B, C] (per sample, e.g. conv bias + embedding projection)."""
    if in_bias is None or in_bias.dim() == 1:
        return 0
    if in_bias.dtype != torch.float32 or in_bias.dim() != 2 or in_bias.shape[1] != channels or in_bias.stride(1) != 1:
        raise _lib.NhmcError("in_bias must be fp32 [C] or [B, C] with unit stride along C")
    return in_bias.stride(0)


def gn_fwd(mode, x, y, gamma, beta, scale, shift, stats, groups, eps, channels_last, in_bias=None):
    """x,y: [B,C,H,W] tensors, either contiguous (NCHW) or channels_last memory format."""
    B, Cc = x.shape[0], x.shape[1]
    hw = x[0, 0].numel()
    ws = _gn_ws(x.device, B)
    with _timed("gn_fwd", 1, 2 * x.numel() * x.element_size()):
        check(lib().nhmc_gn_fwd(mode, _lib.LAYOUT_NHWC if channels_last else _lib.LAYOUT_NCHW, _DT[x.dtype], _p(x), _p(y),
                                _p(gamma), _p(beta), _p(scale), _p(shift), _film_stride(scale, shift, Cc), _p(in_bias),
                                _in_bias_stride(in_bias, Cc), _p(stats), B, Cc, hw, groups, float(eps), ws.buf.data_ptr(), ws.nbytes,
                                _flags(), _stream()), "nhmc_gn_fwd")


def gn_bwd(mode, x, dy, dx, gamma, beta, scale, shift, stats, groups, channels_last, in_bias=None, dx_add=None, dx_add2=None):
    """dx = d/dx act(GN(x)) . dy (+ dx_add, e.g. the gradient of the ResBlock's skip branch, + dx_add2, the gradient
    arriving through the U-Net's long skip connection, in the same pass)."""
    B, Cc = x.shape[0], x.shape[1]
    hw = x[0, 0].numel()
    n_add = (1 if dx_add is not None else 0) + (1 if dx_add2 is not None else 0)
    ws = _gn_ws(x.device, B)
    with _timed("gn_bwd", 1, (3 + n_add) * x.numel() * x.element_size()):
        check(lib().nhmc_gn_bwd(mode, _lib.LAYOUT_NHWC if channels_last else _lib.LAYOUT_NCHW, _DT[x.dtype], _p(x), _p(dy),
                                _p(dx), _p(gamma), _p(beta), _p(scale), _p(shift), _film_stride(scale, shift, Cc),
                                _p(in_bias), _in_bias_stride(in_bias, Cc), _p(stats), 0, 0, _p(dx_add), _p(dx_add2), B, Cc, hw,
                                groups, ws.buf.data_ptr(), ws.nbytes, _flags(), _stream()), "nhmc_gn_bwd")


def gn_check(ws=None):
    """Raises if a per-sample barrier of the resident GroupNorm path timed out in `ws` (default: every workspace this
    module created for the plain per-stream use, plus the current scope's).  Synchronises; the hot loop uses
    GnWorkspace.poll() + raise_if_timed_out() around the synchronisation it has anyway."""
    sc = getattr(_tls, "scope", None)
    todo = [ws] if ws is not None else list(_default_ws.values()) + ([sc[0]] if sc is not None and sc[0] is not None else [])
    for w in todo:
        w.poll()
    torch.cuda.synchronize()
    for w in todo:
        w.raise_if_timed_out()


def softmax_fwd(logits, probs, scale):
    cols = logits.shape[-1]
    rows = logits.numel() // cols
    with _timed("softmax_fwd", 1, 2 * logits.numel() * logits.element_size()):
        check(lib().nhmc_softmax_fwd(_DT[logits.dtype], _p(logits), _p(probs), float(scale), rows, cols, _stream()),
              "nhmc_softmax_fwd")


def attn_fwd_supported(dtype, tokens, head_dim):
    """Shapes the fused attention forward (nhmc_attn_fwd) takes; everything else keeps GEMM + softmax + GEMM."""
    return dtype in (torch.bfloat16, torch.float16) and head_dim in (32, 64) and tokens % 8 == 0


def attn_fwd(qk, v, out, heads, scale):
    """out[b, t, h*ch + c] = sum_s softmax_s(scale * q[b,t,h,:] . k[b,s,h,:]) v[b, h*ch + c, s]
    qk [B, T, heads, 2, ch] (= [B, T, 2C]) token-major, v [B, C, T] channel-major, out [B, T, C] token-major."""
    for t_, n_ in ((qk, "qk"), (v, "v"), (out, "out")):
        _req(t_, name=n_)
    B, T, C2 = qk.shape
    ch = C2 // (2 * heads)
    assert v.shape == (B, C2 // 2, T) and out.shape == (B, T, C2 // 2) and qk.dtype == v.dtype == out.dtype
    nb = (qk.numel() + v.numel() + out.numel()) * qk.element_size()
    with _timed("attn_fwd", 1, nb):
        check(lib().nhmc_attn_fwd(_DT[qk.dtype], _p(qk), _p(v), _p(out), B, T, heads, ch, float(scale), _flags(), _stream()),
              "nhmc_attn_fwd")


def softmax_bwd(probs, dprobs, dlogits, scale):
    cols = probs.shape[-1]
    rows = probs.numel() // cols
    with _timed("softmax_bwd", 1, 3 * probs.numel() * probs.element_size()):
        check(lib().nhmc_softmax_bwd(_DT[probs.dtype], _p(probs), _p(dprobs), _p(dlogits), float(scale), rows, cols,
                                     _stream()), "nhmc_softmax_bwd")


def pool2_gather(in_big, out_small, scale):
    """channels-last 4-D tensors; out_small[b,:,i,j] = scale * sum of the 2x2 block of in_big."""
    B, Cc, h, w = out_small.shape
    with _timed("pool2_gather", 1, (in_big.numel() + out_small.numel()) * in_big.element_size()):
        check(lib().nhmc_pool2_gather(_DT[in_big.dtype], _p(in_big), _p(out_small), B, h, w, Cc, float(scale), _flags(), _stream()),
              "nhmc_pool2_gather")


def pool2_scatter(in_small, out_big, scale):
    B, Cc, h, w = in_small.shape
    with _timed("pool2_scatter", 1, (in_small.numel() + out_big.numel()) * in_small.element_size()):
        check(lib().nhmc_pool2_scatter(_DT[in_small.dtype], _p(in_small), _p(out_big), B, h, w, Cc, float(scale), _flags(), _stream()),
              "nhmc_pool2_scatter")


def split2(g, o1, o2):
    """g: channels-last [B, C1+C2, H, W] -> o1 [B,C1,H,W], o2 [B,C2,H,W] (channels-last)."""
    rows = g.shape[0] * g.shape[2] * g.shape[3]
    with _timed("split2", 1, 2 * g.numel() * g.element_size()):
        check(lib().nhmc_split2(_DT[g.dtype], _p(g), _p(o1), _p(o2), rows, o1.shape[1], o2.shape[1], _stream()), "nhmc_split2")


def bias_add(a, b, bias, out, c=None):
    """channels-last 4-D tensors: out = a (+ b (+ c)) (+ bias[ch]); out may alias a."""
    B, Cc, H, W = a.shape
    nb = (a.numel() * (2 + (1 if b is not None else 0) + (1 if c is not None else 0))) * a.element_size()
    with _timed("bias_add", 1, nb):
        check(lib().nhmc_bias_add(_DT[a.dtype], _p(a), _p(b), _p(c), _p(bias), _p(out), B * H * W, Cc, _flags(), _stream()), "nhmc_bias_add")


def gn_cat_fwd(mode, x0, x1, y, gamma, beta, scale, shift, stats, groups, eps):
    """GroupNorm (+ follow-up) over the virtual concatenation cat([x0, x1], 1); all tensors channels-last."""
    B, c0, c1 = x0.shape[0], x0.shape[1], x1.shape[1]
    hw = x0[0, 0].numel()
    ws = _gn_ws(x0.device, B)
    with _timed("gn_fwd", 2, 2 * y.numel() * y.element_size()):
        check(lib().nhmc_gn_cat_fwd(mode, _DT[x0.dtype], _p(x0), _p(x1), c0, c1, _p(y), _p(gamma), _p(beta), _p(scale),
                                    _p(shift), _film_stride(scale, shift, c0 + c1), _p(stats), B, hw, groups, float(eps),
                                    ws.buf.data_ptr(), ws.nbytes, _flags(), _stream()), "nhmc_gn_cat_fwd")


def gn_cat_bwd(mode, x0, x1, dy, dx0, dx1, gamma, beta, scale, shift, stats, groups, add0=None, add1=None):
    B, c0, c1 = x0.shape[0], x0.shape[1], x1.shape[1]
    hw = x0[0, 0].numel()
    n_add = (add0.numel() if add0 is not None else 0) + (add1.numel() if add1 is not None else 0)
    ws = _gn_ws(x0.device, B)
    with _timed("gn_bwd", 2, (3 * dy.numel() + n_add) * dy.element_size()):
        check(lib().nhmc_gn_cat_bwd(mode, _DT[x0.dtype], _p(x0), _p(x1), c0, c1, _p(dy), _p(dx0), _p(dx1), _p(gamma), _p(beta),
                                    _p(scale), _p(shift), _film_stride(scale, shift, c0 + c1), _p(stats), _p(add0), _p(add1),
                                    B, hw, groups, ws.buf.data_ptr(), ws.nbytes, _flags(), _stream()), "nhmc_gn_cat_bwd")


def vq_nearest(z, codebook, zq, idx=None):
    """z, zq: fp32 [B, D, H, W] contiguous; codebook fp32 [K, D]; idx int32 [B, H*W] or None."""
    _req(z, torch.float32, "z"); _req(zq, torch.float32, "zq"); _req(codebook, torch.float32, "codebook")
    B, D = z.shape[0], z.shape[1]
    P = z[0, 0].numel()
    with _timed("vq_nearest", 1, 2 * z.numel() * 4):
        check(lib().nhmc_vq_nearest(_p(z), _p(codebook), _p(zq), _p(idx), B, D, P, codebook.shape[0], _stream()), "nhmc_vq_nearest")
