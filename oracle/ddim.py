"""Few-step DDIM epilogue, forward and hand-derived backward (oracle; test infrastructure).

Restates Unconditional.cal_x0 / map_back (/root/reference/algos/unconditional.py:9-28;
same x0-prediction formula as /root/reference/functions/denoising.py:62) and the loop of
iterative_sampling (/root/reference/main_sampling.py:898-915).

With a = abar_t, s = abar_next (fp32 scalars taken from the alpha table):
    x0raw = (xt - sqrt(1-a) * eps) / sqrt(a)
    xt'   = sqrt(s) * clamp(x0raw, -1, 1) + sqrt(1-s) * eps
Backward (clamp passes gradient where -1 <= x0raw <= 1, inclusive):
    g_x0  = sqrt(s) * g' * mask
    g_xt  = g_x0 / sqrt(a)
    g_eps = sqrt(1-s) * g' - sqrt(1-a) * g_xt
"""
import torch


def coeffs(a, s):
    """fp32 coefficients in the reference's own evaluation order.

    a, s are 0-dim fp32 tensors (alpha table entries); `(1 - at).sqrt()`, `at.sqrt()`,
    `(1 - at_next).sqrt()`, `at_next.sqrt()` are all fp32 tensor ops in the reference.
    """
    a = torch.as_tensor(a, dtype=torch.float32)
    s = torch.as_tensor(s, dtype=torch.float32)
    return (1 - a).sqrt(), a.sqrt(), (1 - s).sqrt(), s.sqrt()


def epilogue_fwd(xt, eps, a, s):
    """Returns (xt_next, mask). xt, eps: [B,C,H,W] of one dtype (fp32 or fp64)."""
    c1, sa, c2, ss = [c.to(xt.dtype) for c in coeffs(a, s)]
    x0raw = (xt - eps * c1) / sa
    mask = (x0raw >= -1) & (x0raw <= 1)
    x0 = x0raw.clip(-1, 1)
    add_up = c2 * eps
    return ss * x0 + add_up, mask


def epilogue_bwd(g_next, mask, a, s):
    """Returns (g_xt_direct, g_eps) for upstream gradient g_next = dL/dxt'."""
    c1, sa, c2, ss = [c.to(g_next.dtype) for c in coeffs(a, s)]
    g_x0 = ss * g_next * mask.to(g_next.dtype)
    g_xt = g_x0 / sa
    g_eps = c2 * g_next - g_xt * c1
    return g_xt, g_eps


def ddim_map(x, eps_net, alpha_tab, seq, seq_next):
    """G: x_T -> x_hat_0, keeping what the backward pass needs.

    eps_net(xt, t_value) -> (eps [B,C,H,W], vjp_fn(g_eps) -> g_xt).  alpha_tab[t+1] = abar_t.
    Returns (x_hat, tape) where tape is a list of (mask, a, s, vjp_fn) per step.
    """
    xt = x
    tape = []
    for i, j in zip(reversed(seq), reversed(seq_next)):
        a, s = alpha_tab[i + 1], alpha_tab[j + 1]
        eps, vjp = eps_net(xt, i)
        xt, mask = epilogue_fwd(xt, eps, a, s)
        tape.append((mask, a, s, vjp))
    return xt, tape


def ddim_map_vjp(g_xhat, tape):
    """Pull g_xhat = dL/dx_hat back through G to dL/dx."""
    g = g_xhat
    for mask, a, s, vjp in reversed(tape):
        g_xt, g_eps = epilogue_bwd(g, mask, a, s)
        g = g_xt + vjp(g_eps)
    return g
