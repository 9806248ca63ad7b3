"""Noise-space HMC sampler: drop-in `hmc()` / `hmc_latent()` on the sm_100a kernels.

Mirrors /root/reference/main_sampling.py:660-774 (`hmc`) and
/root/reference/main_sampling_latent.py:623-762 (`hmc_latent`): same signatures, same `opt` fields
(tau, epsilon, m, sigma_0, sigma_y), same algo / H_funcs / model objects.  What changes is where the
work runs:

  * the score network (any `model(x, t)`; ours or the reference's) stays a PyTorch call, its input-VJP
    a `torch.autograd.grad` call;
  * everything around it is a handful of fused kernels (include/nhmc.h): DDIM epilogue fwd/bwd,
    operator loss+gradient, leapfrog with in-register energies, on-device Metropolis + schedule;
  * chains are a batch dimension: every chain carries its own epoch / sigma_y / epsilon / reject
    counter in a device array, so B independent replicas of the (batch-1-only) reference algorithm
    advance in lock step with no host round trip inside an iteration.
"""
import copy
import math
import warnings

import numpy as np
import torch

from . import _lib, kernels, operators, schedule
from .chain import ChainState, new_host_state


class GradEvaluator:
    """loss(x) = sum (y - A(clip(G(x))))^2 and its gradient, G = the few-step DDIM map.

    Replaces `iterative_sampling` + `cal_x0`/`map_back` + `H_funcs.H` + `autograd.grad`
    (main_sampling.py:693-695,709-711,898-915; algos/unconditional.py:9-28).  Autograd is used only
    across the score network; the epilogues and the operator are differentiated by the kernels.
    """

    def __init__(self, model, steps, op, y, example_x, net_grad=True, decode=None):
        self.model, self.steps, self.op, self.y = model, steps, op, y
        self.net_grad = net_grad
        self.decode = decode
        B = example_x.shape[0]
        dev = example_x.device
        self.B, self.N = B, example_x[0].numel()
        self.t = [torch.ones(B, device=dev) * float(s[0]) for s in steps]       # main_sampling.py:903
        self.xt = [torch.empty_like(example_x) for _ in steps]                  # outputs of each epilogue
        words = kernels.ddim_mask_words(self.N)
        self.mask = [torch.empty(B, words, dtype=torch.int32, device=dev) for _ in steps]
        self.gxt = [torch.empty_like(example_x) for _ in range(2)]
        self.g_hat = torch.empty_like(example_x)
        self.xc = torch.empty_like(example_x)
        self.loss = torch.zeros(B, dtype=torch.float64, device=dev)
        self.g_eps = None     # cotangent buffer for the network output, allocated on first use
        self.n_net_fwd = 0
        self.n_net_vjp = 0

    def split(self, bounds):
        """Evaluators for the chain groups [bounds[i], bounds[i+1]): same model / operator / schedule, every buffer a
        VIEW of this evaluator's (chains are contiguous along dim 0), so `loss`, `xc`, `g_hat` stay whole-batch tensors
        for the accept / commit kernels.  Chains are independent, so groups can run on separate streams."""
        subs = []
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            e = object.__new__(GradEvaluator)
            e.model, e.steps, e.op, e.net_grad, e.decode = self.model, self.steps, self.op, self.net_grad, self.decode
            e.y = self.y[lo:hi]
            e.B, e.N = hi - lo, self.N
            e.t = [t[lo:hi] for t in self.t]
            e.xt = [t[lo:hi] for t in self.xt]
            e.mask = [t[lo:hi] for t in self.mask]
            e.gxt = [t[lo:hi] for t in self.gxt]
            e.g_hat, e.xc, e.loss = self.g_hat[lo:hi], self.xc[lo:hi], self.loss[lo:hi]
            e.g_eps = None
            e.n_net_fwd = e.n_net_vjp = 0
            subs.append(e)
        return subs

    def _net(self, x_in, k):
        if self.net_grad:
            xk = x_in.detach().requires_grad_(True)
            with torch.enable_grad():
                out = self.model(xk, self.t[k])
        else:
            xk = None
            with torch.no_grad():
                out = self.model(x_in, self.t[k])
        self.n_net_fwd += 1
        return xk, out

    def __call__(self, x):
        """Returns (loss[B] fp64, g_a, g_b, xc): dloss/dx = g_a + g_b (g_b may be None)."""
        tape = []
        x_in = x
        for k, (_, c1, sa, c2, ss) in enumerate(self.steps):
            xk, out = self._net(x_in, k)
            eps = out.detach()
            if not eps.is_contiguous():
                eps = eps.contiguous()
            kernels.ddim_fwd(x_in, eps, self.xt[k], self.mask[k], (c1, sa, c2, ss))
            tape.append((xk, out))
            x_in = self.xt[k]

        if self.decode is None:
            self.op.loss_grad(x_in, self.y, g=self.g_hat, xc=self.xc, loss=self.loss)
            g_a, g_b = self.g_hat, None
        else:
            g_a, g_b = self._decode_loss_grad(x_in), None

        for k in reversed(range(len(self.steps))):
            _, c1, sa, c2, ss = self.steps[k]
            xk, out = tape[k]
            g_xt = self.gxt[k % 2]
            if self.net_grad:
                if self.g_eps is None or self.g_eps.shape != out.shape or self.g_eps.dtype != out.dtype:
                    self.g_eps = torch.zeros(out.shape, dtype=out.dtype, device=out.device)  # channels >= C stay 0
                kernels.ddim_bwd(g_a, g_b, self.mask[k], g_xt, self.g_eps, (c1, sa, c2, ss))
                (vjp,) = torch.autograd.grad(out, xk, self.g_eps)
                self.n_net_vjp += 1
                g_a, g_b = g_xt, vjp.contiguous()
            else:
                kernels.ddim_bwd(g_a, g_b, self.mask[k], g_xt, None, (c1, sa, c2, ss))
                g_a, g_b = g_xt, None
        return self.loss, g_a, g_b, self.xc

    def _decode_loss_grad(self, z_hat):
        """Latent path: loss = sum (y - H(decode(clip(z_hat))))^2 (main_sampling_latent.py:651,670).
        The decoder and its VJP stay PyTorch; the latent clip mask is applied here."""
        zc = torch.empty_like(z_hat)
        kernels.clip_fwd(z_hat, zc, self.xc)                           # clamp + the kept sample, one launch
        zv = zc.requires_grad_(True)
        with torch.enable_grad():
            img = self.decode(zv)
        img_c = img.detach().float().contiguous()
        _, g_img, _ = self.op.loss_grad(img_c, self.y, loss=self.loss)
        (g_z,) = torch.autograd.grad(img, zv, g_img.to(img.dtype))
        kernels.clip_bwd(g_z.contiguous(), z_hat, self.g_hat)          # gradient mask of the clip (+ cast), one launch
        return self.g_hat


class HMCSampler:
    """State + one-iteration stepper shared by hmc() and hmc_latent()."""

    def __init__(self, x, evaluator, tau, epsilon, sigma_0, mass=1.0, sched=_lib.SCHED_PIXEL, sigma_y=None,
                 n_slots=20, reuse_gradient=False, lanes=1, diag_mass=False):
        """lanes > 1: the chains are cut into that many contiguous groups, each evaluated (score network, epilogues,
        operator, VJP, leapfrog kernel) on its own stream; in a captured step the groups are parallel branches of the
        graph, so one group's bandwidth-bound / latency-bound kernels overlap the other's convolutions.  Per-chain
        results do not depend on the grouping (no quantity couples two chains).

        reuse_gradient (opt-in): keep (loss, gradient) of the current state per chain instead
        of re-evaluating them at the top of every iteration (main_sampling.py:693-695): 20 instead of 21 gradient
        evaluations per iteration, same values (the reference recomputes exactly what it already had)."""
        self.ev = evaluator
        self.sched, self.mass, self.n_slots = sched, float(mass), n_slots
        self.x = x.detach().clone().float().contiguous()
        self.B = self.x.shape[0]
        dev = self.x.device
        self.L = schedule.leapfrog_steps(tau, epsilon)
        self.chains = ChainState(new_host_state(self.B, tau, epsilon, sigma_0, sigma_y), dev)
        self.x_prop = torch.empty_like(self.x)
        self.p = torch.empty_like(self.x)
        self.x_keep = torch.zeros_like(self.x) if sched == _lib.SCHED_LATENT else None
        self.samples = torch.zeros((self.B, n_slots) + tuple(self.x.shape[1:]), dtype=torch.float32, device=dev)
        self.ws = kernels.leapfrog_workspace(self.B, dev)
        self.u_dev = torch.empty(self.B, dtype=torch.float32, device=dev)
        self.leapfrog_steps_done = 0
        self._graphs = None          # {mode: (CUDAGraph, launches)} once capture_graphs() has run
        self.reuse_gradient = bool(reuse_gradient)
        self.g_state = torch.zeros_like(self.x) if self.reuse_gradient else None
        self.loss_state = torch.zeros(self.B, dtype=torch.float64, device=dev) if self.reuse_gradient else None
        self._state_valid = False    # (loss_state, g_state) hold the values at self.x
        self._last = {}              # mode -> [(lo, hi, g_a, g_b)] of that step variant's evaluation (static buffers under graphs)
        # diagonal mass (hmc_test_conditioning, main_sampling.py:776-894): per-element mass re-estimated from the
        # trajectory's running variance on accepted iterations; M = 1 until the first update
        self.diag = bool(diag_mass)
        if self.diag:
            assert not reuse_gradient and int(lanes) <= 1, "the diagonal-mass variant runs single-lane without gradient reuse"
            self.inv_mass = torch.ones_like(self.x)
            self.std_mass = torch.ones_like(self.x)
            self.wf_mean = torch.zeros_like(self.x)
            self.wf_m2 = torch.zeros_like(self.x)
        self.n_lanes = max(1, min(int(lanes), self.B, 8))
        self.lanes = None
        # GroupNorm scratch of the score network's fused epilogues: owned by THIS sampler (one per lane), so samplers on
        # other host threads / streams never share accumulators with it (include/nhmc.h: the library keeps no state)
        self.gn_ws = kernels.GnWorkspace(dev, self.B)
        if self.n_lanes > 1:
            bounds = [(i * self.B) // self.n_lanes for i in range(self.n_lanes + 1)]
            self.lanes = [dict(idx=i, lo=lo, hi=hi, ev=e, chains=self.chains.view(lo, hi - lo),
                               ws=kernels.leapfrog_workspace(hi - lo, dev), stream=torch.cuda.Stream(device=dev),
                               gn_ws=kernels.GnWorkspace(dev, hi - lo))
                          for i, (lo, hi, e) in enumerate(zip(bounds[:-1], bounds[1:], evaluator.split(bounds)))]

    # ---- one gradient evaluation + leapfrog kernel, eager or as a replayed CUDA graph -------------------
    def _step_eager(self, mode):
        src = self.x if mode == _lib.LF_FIRST else self.x_prop
        if self.lanes is None:
            with kernels.gn_scope(self.gn_ws, 1):
                loss, g_a, g_b, _ = self.ev(src)
            self._last[mode] = [(0, self.B, g_a, g_b)]
            if self.diag:
                kernels.leapfrog_diag(mode, src, self.x_prop, self.p, g_a, g_b, loss, self.chains, self.inv_mass, self.wf_mean,
                                      self.wf_m2, self.ws)
            else:
                kernels.leapfrog(mode, src, self.x_prop, self.p, g_a, g_b, loss, self.chains, self.mass, self.ws)
            return
        cur = torch.cuda.current_stream()
        last = []
        for ln in self.lanes:                           # fork: every group on its own stream, with its own GroupNorm scratch
            lo, hi = ln["lo"], ln["hi"]
            ln["stream"].wait_stream(cur)
            with torch.cuda.stream(ln["stream"]), kernels.gn_scope(ln["gn_ws"], self.n_lanes):
                loss, g_a, g_b, _ = ln["ev"](src[lo:hi])
                kernels.leapfrog(mode, src[lo:hi], self.x_prop[lo:hi], self.p[lo:hi], g_a, g_b, loss, ln["chains"],
                                 self.mass, ln["ws"])
            last.append((lo, hi, g_a, g_b))
        for ln in self.lanes:                           # join
            cur.wait_stream(ln["stream"])
        self._last[mode] = last
        self.ev.n_net_fwd = sum(ln["ev"].n_net_fwd for ln in self.lanes) // self.n_lanes
        self.ev.n_net_vjp = sum(ln["ev"].n_net_vjp for ln in self.lanes) // self.n_lanes

    def _first_from_state(self):
        """FIRST step without a gradient evaluation: energies, half kick and first drift from the kept (loss, g)."""
        kernels.leapfrog(_lib.LF_FIRST, self.x, self.x_prop, self.p, self.g_state, None, self.loss_state, self.chains,
                         self.mass, self.ws)

    def step(self, mode):
        """FIRST: evaluate at x, half kick, first drift (x -> x_prop). MIDDLE/LAST: evaluate at x_prop, kick (+ drift)."""
        if self._graphs is None:
            self._step_eager(mode)
        else:
            g, n = self._graphs[mode]
            g.replay()
            kernels.LAUNCHES += n

    def capture_graphs(self, warmup=2, autotune=False):
        """Capture the three step variants (score-network forwards, fused epilogues, operator, VJPs, leapfrog
        kernel: several thousand launches each) as CUDA graphs sharing one memory pool.  Per-chain scalars live
        in device memory (nhmc_chain), so the graphs stay valid for every iteration of every chain.  The sampler
        state touched while warming up / capturing is restored afterwards.

        autotune: `torch.backends.cudnn.benchmark` is switched on for the warm-up and the capture only (and restored):
        the library convolutions the graphs replay are the ones cuDNN measured fastest for these shapes, not its
        heuristic picks (17.8 -> 16.3 ms of convolutions per 8-chain step, 228 -> 238 chain-steps/s).  The drop-in
        functions ask for it when the score network computes in bf16 / fp16 only: for an fp32 network the measured-fastest
        algorithm may be a transform-domain one whose rounding differs from run to run of the tuner, and the fp32 path is
        the parity path."""
        prev_benchmark = torch.backends.cudnn.benchmark
        if autotune:
            torch.backends.cudnn.benchmark = True
        try:
            return self._capture_graphs(warmup)
        finally:
            torch.backends.cudnn.benchmark = prev_benchmark

    def _capture_graphs(self, warmup):
        keep = [t.clone() for t in self._state_tensors()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # cuDNN autotuning, scratch allocation: outside capture
                for mode in (_lib.LF_FIRST, _lib.LF_MIDDLE, _lib.LF_LAST):
                    self._step_eager(mode)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs, pool = {}, None
        for mode in (_lib.LF_FIRST, _lib.LF_MIDDLE, _lib.LF_LAST):
            g = torch.cuda.CUDAGraph()
            n0 = kernels.LAUNCHES
            with torch.cuda.graph(g, pool=pool):
                self._step_eager(mode)
            graphs[mode] = (g, kernels.LAUNCHES - n0)
            pool = g.pool()
        torch.cuda.synchronize()
        for dst, src in zip(self._state_tensors(), keep):
            dst.copy_(src)
        self._graphs = graphs
        return self

    def _state_tensors(self):
        """Everything a step variant mutates (restored after warm-up / capture)."""
        return (self.x, self.x_prop, self.p, self.chains.dev) + ((self.wf_mean, self.wf_m2) if self.diag else ())

    def _mass_update(self):
        """Accepted iterations from epoch > epochs // 3 on re-estimate the diagonal mass from the trajectory's variance
        (main_sampling.py:857-868): variance = M2 / (L - 1), rank transform to scores in [-1, 1], M = exp(scores).  Per
        chain, on device, masked by the accept decision; the sort (ties: the all-zero variance of the first updates) is
        torch.sort, the reference's own call -- its tie order is implementation-defined and taken as is."""
        B, N = self.B, self.x[0].numel()
        epoch_after = self.chains.field_i32("epoch")
        upd = (self.chains.field_i32("accept") != 0) & ((epoch_after - 1) > 40 // 3)           # epoch before the increment
        variance = (self.wf_m2 / (self.L - 1)).reshape(B, N)
        _, sorted_idx = torch.sort(variance, dim=1)
        ranks = torch.empty(B, N, dtype=torch.float32, device=self.x.device)
        ranks.scatter_(1, sorted_idx, torch.arange(N, dtype=torch.float32, device=self.x.device).expand(B, N).contiguous())
        scores = 2.0 * (ranks / (N - 1)) - 1.0
        M = torch.exp(1 * scores).view_as(self.x)
        m = upd.view(B, *([1] * (self.x.dim() - 1)))
        self.std_mass.copy_(torch.where(m, torch.sqrt(M), self.std_mass))
        self.inv_mass.copy_(torch.where(m, 1.0 / M, self.inv_mass))

    def capture_profiled_step(self, mode=_lib.LF_MIDDLE, events=True):
        """One step variant captured as a CUDA graph WITH a pair of timing events around every kernel of ours
        (event-record nodes).  Returns (graph, records): replay, then kernels.profile_read(records, acc) gives the
        per-kernel CUDA-event times of that replay -- measured in situ (warm L2, back-to-back launches) on the
        launching stream.  Sampler state is restored after the capture.  events=False: the same capture without the
        event nodes (e.g. with other library settings than the sampler's own graphs, for A/B or profiler passes)."""
        keep = [t.clone() for t in self._state_tensors()]
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        pool = next(iter(self._graphs.values()))[0].pool() if self._graphs else None
        if events:
            kernels.profile_start()
        with torch.cuda.graph(g, pool=pool):
            self._step_eager(mode)
        records = kernels.profile_take() if events else {}
        torch.cuda.synchronize()
        for dst, src in zip(self._state_tensors(), keep):
            dst.copy_(src)
        return g, records

    def iteration(self, p, u):
        """One HMC iteration for every chain. p: momentum draw [B,C,H,W]; u: uniform draws [B]."""
        ev, ch, m = self.ev, self.chains, self.mass
        kernels.begin_iteration(ch, self.sched)
        self.p.copy_(p)
        if self.diag:
            self.p.mul_(self.std_mass)                                           # p ~ N(0, M) (main_sampling.py:819)
            self.wf_mean.zero_(); self.wf_m2.zero_()                             # :806-807
        if self.reuse_gradient and self._state_valid:
            self._first_from_state()                                             # no evaluation #0
        else:
            self.step(_lib.LF_FIRST)                                             # main_sampling.py:693-702,706
            if self.reuse_gradient:   # first iteration: every chain's state values come from this evaluation
                for lo, hi, g_a, g_b in self._last[_lib.LF_FIRST]:
                    torch.add(g_a, g_b, out=self.g_state[lo:hi]) if g_b is not None else self.g_state[lo:hi].copy_(g_a)
                self.loss_state.copy_(ev.loss)
                self._state_valid = True
        for l in range(self.L):                                                  # main_sampling.py:704-715
            self.step(_lib.LF_LAST if l == self.L - 1 else _lib.LF_MIDDLE)
        xc = ev.xc
        self.leapfrog_steps_done += self.L
        self.u_dev.copy_(torch.as_tensor(u, dtype=torch.float32).reshape(self.B), non_blocking=True)
        kernels.accept(ch, self.u_dev, 1.0 if self.diag else m, self.sched, self.n_slots)   # main_sampling.py:717-749
        kernels.commit(ch, self.x, self.x_prop, xc, self.x_keep, self.samples, self.sched, self.n_slots)
        if self.diag:
            self._mass_update()
        if self.reuse_gradient:
            for lo, hi, g_a, g_b in self._last[_lib.LF_LAST]:                    # evaluation at the proposal
                kernels.commit_grad(ch.view(lo, hi - lo), self.g_state[lo:hi], g_a, g_b, self.loss_state[lo:hi], ev.loss[lo:hi])
        return xc

    def leapfrog_step_host(self, x_host, p_host):
        """One leapfrog step (main_sampling.py:706-713) on HOST buffers: position and momentum come from
        (pinned) host memory, the step runs on the device, the per-chain loss and the updated (x, p) go back.
        Used for end-to-end timing; returns {"loss": [B] float64 (host), "bytes": D2H bytes}."""
        self.x_prop.copy_(x_host, non_blocking=True)
        self.p.copy_(p_host, non_blocking=True)
        self.step(_lib.LF_MIDDLE)
        x_host.copy_(self.x_prop, non_blocking=True)
        p_host.copy_(self.p, non_blocking=True)
        loss_h = self.ev.loss.cpu()               # synchronises: the step's result is on the host
        return {"loss": loss_h, "bytes": 2 * x_host.numel() * 4 + loss_h.numel() * 8}

    def gn_workspaces(self):
        return [self.gn_ws] + ([ln["gn_ws"] for ln in self.lanes] if self.lanes else [])

    def poll_gn(self):
        """Queue copies of the GroupNorm workspaces' timeout flags (no sync); check_gn() after the next synchronisation."""
        for w in self.gn_workspaces():
            w.poll()

    def check_gn(self):
        for w in self.gn_workspaces():
            w.raise_if_timed_out()

    def collected(self):
        """[B, S, ...] samples in collection order (S = min(n_collected, n_slots), same for all finished chains)."""
        st = self.chains.read()
        out = []
        for b in range(self.B):
            n = int(st["n_collected"][b])
            if n <= self.n_slots:
                out.append(self.samples[b, :n])
            else:   # ring buffer wrapped: oldest kept sample first
                k = n % self.n_slots
                out.append(torch.cat([self.samples[b, k:], self.samples[b, :k]], 0))
        return out


def _default_draws(m):
    def momentum(x):
        return torch.randn_like(x) * math.sqrt(m)             # main_sampling.py:692 (device generator)

    def uniform(B):
        return torch.rand(B)                                   # main_sampling.py:720 (global CPU generator)
    return momentum, uniform


def _run(sampler, momentum, uniform, max_iters, on_iteration):
    it = 0
    while True:
        if max_iters is not None and it >= max_iters:
            break
        p = momentum(sampler.x)
        xc = sampler.iteration(p, uniform(sampler.B))
        sampler.poll_gn()
        st = sampler.chains.read()                             # the one host sync per iteration
        sampler.check_gn()                                     # a barrier timeout of the resident GroupNorm kernels raises here
        if on_iteration is not None:
            on_iteration(it, st, xc)
        it += 1
        if not st["active"].any():
            break
    return it


def _model_of(algo):
    if getattr(algo, "cls_fn", None) is not None:
        raise _lib.NhmcError("classifier guidance (cls_fn) is not on the accelerated hmc path")
    return algo.model


def _unconditional_step(xt, eps, at, at_next):
    """Unconditional.cal_x0 + map_back in closed form (algos/unconditional.py:17-28)."""
    x0 = ((xt - eps * (1 - at).sqrt()) / at.sqrt()).clip(-1, 1)
    return at_next.sqrt() * x0 + (1 - at_next).sqrt() * eps


def plugin_is_unconditional(algo, channels, device, noise="ddpm"):
    """True when `algo` has no cal_x0 / map_back of its own (a bare object with `.model`) or when its cal_x0 + map_back
    reproduce Unconditional's (algos/unconditional.py:9-28) on a probe: random xt through the plugin with a stub score
    network against the closed form the fused epilogue kernels implement.  A Base_Algo subclass that does anything else
    (base_algo.py:3-16) is NOT silently replaced by the fused epilogue: hmc() then differentiates the plugin itself."""
    if not (callable(getattr(algo, "cal_x0", None)) and callable(getattr(algo, "map_back", None))):
        return True
    gen = torch.Generator().manual_seed(7)
    n, d = 2, 8
    xt = (3.0 * torch.randn(n, channels, d, d, generator=gen)).to(device)
    eps = torch.randn(n, 2 * channels, d, d, generator=gen).to(device)
    y_0 = torch.zeros(n, channels * d * d, device=device)
    at = torch.full((n, 1, 1, 1), 0.0778, device=device)
    at_next = torch.full((n, 1, 1, 1), 0.5214, device=device)
    t = torch.ones(n, device=device) * 500
    real = algo.model
    try:
        try:
            algo.model = lambda x_, t_, *a, **k: eps
        except Exception:
            return False                                   # cannot probe: treat as a custom plugin
        with torch.no_grad():
            x0_t, add_up = algo.cal_x0(xt, t, at, at_next, y_0, noise)
            out = algo.map_back(x0_t, y_0, add_up, at_next, at)
    except Exception:
        return False
    finally:
        try:
            algo.model = real
        except Exception:
            pass
    want = _unconditional_step(xt, eps[:, :channels], at, at_next)
    return out.shape == want.shape and bool((out.float() - want).abs().max() <= 1e-5 * float(want.abs().max()))


class PluginGradEvaluator:
    """loss and gradient through ANY Base_Algo plugin: cal_x0 / map_back are called exactly as iterative_sampling does
    (main_sampling.py:898-915) and differentiated by autograd; only the measurement loss + adjoint run in the fused
    operator kernel.  The slow but faithful route for plugins that are not `Unconditional`."""

    def __init__(self, algo, b, seq, seq_next, op, y, y_0, example_x, noise):
        self.algo, self.b, self.seq, self.seq_next, self.op, self.y, self.y_0, self.noise = algo, b, list(seq), list(seq_next), op, y, y_0, noise
        self.B, self.N = example_x.shape[0], example_x[0].numel()
        self.g_hat = torch.empty_like(example_x)
        self.xc = torch.empty_like(example_x)
        self.loss = torch.zeros(self.B, dtype=torch.float64, device=example_x.device)
        self.n_net_fwd = self.n_net_vjp = 0
        # (t, at, at_next) of every step, formed once exactly as iterative_sampling forms them per call
        # (main_sampling.py:903-906): constants of the run, and host->device copies could not be captured into a graph
        self.sched = []
        dev = example_x.device
        for i, j in zip(reversed(self.seq), reversed(self.seq_next)):
            t = (torch.ones(self.B) * i).to(dev)
            next_t = (torch.ones(self.B) * j).to(dev)
            self.sched.append((t, schedule.compute_alpha(b.to(dev), t.long()), schedule.compute_alpha(b.to(dev), next_t.long())))

    def split(self, bounds):
        raise _lib.NhmcError("lanes are not available with a custom algo plugin")

    def __call__(self, x):
        xk = x.detach().requires_grad_(True)
        with torch.enable_grad():
            xt = xk
            for t, at, at_next in self.sched:
                x0_t, add_up = self.algo.cal_x0(xt, t, at, at_next, self.y_0, self.noise)
                xt = self.algo.map_back(x0_t, self.y_0, add_up, at_next, at)
                self.n_net_fwd += 1
        self.op.loss_grad(xt.detach().float().contiguous(), self.y, g=self.g_hat, xc=self.xc, loss=self.loss)
        (g,) = torch.autograd.grad(xt, xk, self.g_hat.to(xt.dtype))
        self.n_net_vjp += len(self.seq)
        return self.loss, g.float().contiguous(), None, self.xc


def _capture(smp, cuda_graphs, model):
    """cuda_graphs True: capture (errors propagate); False: eager; None (default): capture, and fall back to eager
    launches with a warning if the score network cannot be captured (host syncs / data-dependent control flow inside a
    third-party model) -- our own networks always capture."""
    if cuda_graphs is False:
        return
    params = getattr(model, "parameters", None)
    autotune = callable(params) and any(p.dtype in (torch.bfloat16, torch.float16) for p in params())
    if cuda_graphs is True:
        smp.capture_graphs(autotune=autotune)
        return
    try:
        smp.capture_graphs(autotune=autotune)
    except Exception as e:                                     # noqa: BLE001  (third-party model code)
        from . import unet
        if isinstance(model, unet.UNetModel):
            raise
        smp._graphs = None
        torch.cuda.synchronize()
        warnings.warn("nhmc_b200: the score network could not be captured into CUDA graphs (%s: %s); running eager "
                      "launches (slower)" % (type(e).__name__, e))


def hmc(x, n, b, seq, seq_next, algo, opt, y_0, H_funcs, x_orig, *, momentum=None, uniform=None,
        max_iters=None, on_iteration=None, return_sampler=False, cuda_graphs=None, reuse_gradient=False, lanes=1):
    """Drop-in for main_sampling.hmc (main_sampling.py:660-774).

    x: [B,3,H,W] initial noise; b: fp32 betas [1000]; seq/seq_next: timestep lists; algo: object with
    `.model` (`model(x,t) -> [B,6,H,W]` or 3 channels); opt: `.tau .epsilon .m .sigma_0` (sigma_0 already
    doubled by prepare_measurement, main_sampling.py:348); y_0: measurement; H_funcs: a reference
    operator object or one of nhmc_b200.operators.  Returns the 20 collected samples, `[20,3,H,W]` for
    B == 1 (exactly the reference's return) and `[B,20,3,H,W]` for B > 1 (B independent chains -- the
    reference itself cannot run B > 1, main_sampling.py:719-720; if the chains collected different numbers of
    samples -- max_iters cut them short -- every chain is truncated to the smallest count).

    The three step variants are captured as CUDA graphs by default (cuda_graphs=None: with an eager fallback for
    score networks that cannot be captured; True: capture or raise; False / opt.cuda_graphs == False: eager).
    A plugin whose cal_x0 / map_back are not Unconditional's (algos/unconditional.py) is differentiated by autograd
    instead of being replaced by the fused DDIM epilogue (PluginGradEvaluator).
    """
    if not x.is_cuda:
        raise _lib.NhmcError("hmc() runs on a CUDA device (sm_100a); there is no CPU fallback")
    B, C, D = x.shape[0], x.shape[1], x.shape[2]
    assert n == B, "n must equal the batch size (main_sampling.py:472)"
    model = _model_of(algo)
    op = operators.from_reference(H_funcs, C, D, x.device)
    y = y_0.detach().float().contiguous().reshape(op.meas_shape(B))
    m = float(getattr(opt, "m", 1.0))
    noise = getattr(opt, "noise", "ddpm")
    if plugin_is_unconditional(algo, C, x.device, noise):
        steps = schedule.ddim_steps(schedule.alpha_table(b), seq, seq_next)
        ev = GradEvaluator(model, steps, op, y, x.detach().float().contiguous(), net_grad=True)
    else:
        ev = PluginGradEvaluator(algo, b, seq, seq_next, op, y, y_0, x.detach().float().contiguous(), noise)
    smp = HMCSampler(x, ev, opt.tau, opt.epsilon, opt.sigma_0, mass=m, sched=_lib.SCHED_PIXEL, n_slots=20,
                     reuse_gradient=reuse_gradient or getattr(opt, "reuse_gradient", False),
                     lanes=max(int(lanes), int(getattr(opt, "lanes", 1) or 1)))
    if cuda_graphs is None and getattr(opt, "cuda_graphs", None) is not None:
        cuda_graphs = bool(opt.cuda_graphs)
    _capture(smp, cuda_graphs, model)
    dm, du = _default_draws(m)
    _run(smp, momentum or dm, uniform or du, max_iters, on_iteration)
    if return_sampler:
        return smp
    out = smp.collected()
    if B == 1:
        return out[0]
    s_min = min(o.shape[0] for o in out)
    return torch.stack([o[:s_min] for o in out])


def hmc_test_conditioning(x, n, b, seq, seq_next, algo, opt, y_0, H_funcs, x_orig, *, momentum=None, uniform=None,
                          max_iters=None, on_iteration=None, return_sampler=False, cuda_graphs=None):
    """Drop-in for main_sampling.hmc_test_conditioning (main_sampling.py:776-894): HMC with an adaptive DIAGONAL mass.
    Same signature and schedule (burn 5, epochs 40, 85 accepted iterations, sigma_y = sigma_0 + 0.9 (1 - (e - 5)/40)^3);
    the momentum is drawn as randn * sqrt(M), the kinetic energy is sum(p^2 / M), the drift x + eps p / M, and every
    accepted iteration from epoch 14 on sets M = exp(rank-transformed variance of the trajectory's positions), the
    variance coming from the Welford pass fused into the leapfrog kernel (nhmc_leapfrog_diag).  Returns the 35 samples
    collected from epoch 50 on ([35,3,H,W] for B == 1, [B,35,3,H,W] for B > 1 independent chains)."""
    if not x.is_cuda:
        raise _lib.NhmcError("hmc_test_conditioning() runs on a CUDA device (sm_100a); there is no CPU fallback")
    B, C, D = x.shape[0], x.shape[1], x.shape[2]
    assert n == B, "n must equal the batch size"
    model = _model_of(algo)
    op = operators.from_reference(H_funcs, C, D, x.device)
    y = y_0.detach().float().contiguous().reshape(op.meas_shape(B))
    noise = getattr(opt, "noise", "ddpm")
    if plugin_is_unconditional(algo, C, x.device, noise):
        steps = schedule.ddim_steps(schedule.alpha_table(b), seq, seq_next)
        ev = GradEvaluator(model, steps, op, y, x.detach().float().contiguous(), net_grad=True)
    else:
        ev = PluginGradEvaluator(algo, b, seq, seq_next, op, y, y_0, x.detach().float().contiguous(), noise)
    smp = HMCSampler(x, ev, opt.tau, opt.epsilon, opt.sigma_0, mass=1.0, sched=_lib.SCHED_COND, n_slots=35, diag_mass=True)
    if cuda_graphs is None and getattr(opt, "cuda_graphs", None) is not None:
        cuda_graphs = bool(opt.cuda_graphs)
    _capture(smp, cuda_graphs, model)
    dm, du = _default_draws(1.0)
    _run(smp, momentum or dm, uniform or du, max_iters, on_iteration)
    if return_sampler:
        return smp
    out = smp.collected()
    if B == 1:
        return out[0]
    s_min = min(o.shape[0] for o in out)
    return torch.stack([o[:s_min] for o in out])


def hmc_latent(x, n, seq, seq_next, algo, opt, y_0, H_funcs, x_orig, *, momentum=None, uniform=None,
               max_iters=None, on_iteration=None, return_sampler=False, image_size=256, reuse_gradient=False, cuda_graphs=None):
    """Drop-in for main_sampling_latent.hmc_latent (main_sampling_latent.py:623-762).

    algo.model is an LDM object exposing `.apply_model(z,t,None)` (no-grad, ddpm.py:892-893),
    `.differentiable_decode_first_stage(z)`, `.alphas_cumprod`, `.alphas_cumprod_prev`.  Returns the last
    <= 10 collected latents ([S,3,h,w] for B == 1, list per chain otherwise).
    """
    if not x.is_cuda:
        raise _lib.NhmcError("hmc_latent() runs on a CUDA device (sm_100a); there is no CPU fallback")
    B = x.shape[0]
    assert n == B
    ldm = _model_of(algo)
    # the decoded image is NOT clipped in the latent loss (main_sampling_latent.py:651): an unclamped VIEW of the operator
    # (own descriptor, shared tables) -- the caller's object keeps its clamp for any pixel-space use
    op = operators.from_reference(H_funcs, 3, image_size, x.device).unclamped()
    table = schedule.latent_alpha_table(ldm.alphas_cumprod, ldm.alphas_cumprod_prev)
    steps = schedule.ddim_steps(table, seq, seq_next)
    y = y_0.detach().float().contiguous().reshape(op.meas_shape(B))
    m = float(getattr(opt, "m", 1.0))
    net = lambda z, t: ldm.apply_model(z, t, None)
    ev = GradEvaluator(net, steps, op, y, x.detach().float().contiguous(), net_grad=False,
                       decode=ldm.differentiable_decode_first_stage)
    smp = HMCSampler(x, ev, opt.tau, opt.epsilon, opt.sigma_0, mass=m, sched=_lib.SCHED_LATENT,
                     sigma_y=opt.sigma_y, n_slots=10,
                     reuse_gradient=reuse_gradient or getattr(opt, "reuse_gradient", False))
    if cuda_graphs is None and getattr(opt, "cuda_graphs", None) is not None:
        cuda_graphs = bool(opt.cuda_graphs)
    _capture(smp, cuda_graphs, ldm)
    dm, du = _default_draws(m)
    _run(smp, momentum or dm, uniform or du, max_iters, on_iteration)
    if return_sampler:
        return smp
    out = smp.collected()
    return out[0] if B == 1 else out
