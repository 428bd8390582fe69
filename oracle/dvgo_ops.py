"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's native ops.

Nothing under oracle/ is shipped or measured as the product; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.

Restates, in torch-CPU fp32 with the reference's float/double mixing kept where it
changes bits:
  * sample_pts_on_rays          lib/cuda/render_utils_kernel.cu:12-73,138-236
  * raw2alpha / backward        lib/cuda/render_utils_kernel.cu:358-428
  * alpha2weight / backward     lib/cuda/render_utils_kernel.cu:431-561
  * adam_upd / masked / perlr   lib/cuda/adam_upd_kernel.cu:9-132
Arithmetic contract (shared with the CUDA kernels in articulated_point_nerf_b200/csrc):
every a*b+c is a separate IEEE multiply and add (no FMA contraction), so torch-CPU
elementwise ops and the kernels' __fmul_rn/__fadd_rn produce identical bits for the
sampling chain.  The reference's nvcc build may contract some of these to FMA; that
is a <=1 ulp difference that no reference test pins.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

F32 = torch.float32


def _f(x) -> torch.Tensor:
    return torch.tensor(float(np.float32(x)), dtype=F32)


def _sqrt_rn(x: torch.Tensor) -> torch.Tensor:
    """Correctly rounded fp32 square root (what CUDA's sqrtf / __fsqrt_rn return).  torch's vectorised CPU
    sqrt is off by one ulp in near-halfway cases (observed: sqrt(1.0474659f) -> 0x1.060154p+0 instead of
    0x1.060156p+0), numpy's is IEEE."""
    return torch.from_numpy(np.sqrt(x.detach().contiguous().numpy()))


# ----------------------------------------------------------------------------------
# ray sampling
# ----------------------------------------------------------------------------------
@torch.no_grad()   # native extension ops build no autograd graph
def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    """render_utils_kernel.cu:12-35."""
    near, far = _f(near), _f(far)
    v = torch.where(rays_d == 0, torch.full_like(rays_d, float(np.float32(1e-6))), rays_d)
    a = (xyz_max[None] - rays_o) / v
    b = (xyz_min[None] - rays_o) / v
    lo = torch.minimum(a, b)
    hi = torch.maximum(a, b)
    t_min = torch.maximum(torch.minimum(torch.maximum(torch.maximum(lo[:, 0], lo[:, 1]), lo[:, 2]), far), near)
    t_max = torch.maximum(torch.minimum(torch.minimum(torch.minimum(hi[:, 0], hi[:, 1]), hi[:, 2]), far), near)
    return t_min, t_max


@torch.no_grad()   # native extension ops build no autograd graph
def infer_n_samples(t_min, t_max, stepdist):
    """render_utils_kernel.cu:38-49 (float divide, ceil, at least one sample)."""
    n = torch.ceil((t_max - t_min) / _f(stepdist))
    return torch.clamp_min(n, 1.0).to(torch.int64)


@torch.no_grad()   # native extension ops build no autograd graph
def infer_ray_start_dir(rays_o, rays_d, t_min):
    """render_utils_kernel.cu:52-73."""
    d0, d1, d2 = rays_d[:, 0], rays_d[:, 1], rays_d[:, 2]
    rnorm = _sqrt_rn((d0 * d0 + d1 * d1) + d2 * d2)
    start = rays_o + rays_d * t_min[:, None]
    direc = rays_d / rnorm[:, None]
    return start, direc


@torch.no_grad()   # native extension ops build no autograd graph
def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """render_utils_kernel.cu:190-236. Returns the reference's 7-tuple."""
    rays_o = rays_o.to(F32)
    rays_d = rays_d.to(F32)
    xyz_min = xyz_min.to(F32)
    xyz_max = xyz_max.to(F32)
    t_min, t_max = infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far)
    n_steps = infer_n_samples(t_min, t_max, stepdist)
    cum = n_steps.cumsum(0)
    n_rays = len(rays_o)
    ray_id = torch.repeat_interleave(torch.arange(n_rays), n_steps)
    first = cum - n_steps
    step_id = torch.arange(int(cum[-1]) if n_rays else 0) - first[ray_id]
    start, direc = infer_ray_start_dir(rays_o, rays_d, t_min)
    dist = _f(stepdist) * step_id.to(F32)
    pts = start[ray_id] + direc[ray_id] * dist[:, None]
    mask_out = ((xyz_min[None] > pts) | (xyz_max[None] < pts)).any(-1)
    return [pts, mask_out, ray_id, step_id, n_steps, t_min, t_max]


# ----------------------------------------------------------------------------------
# raw2alpha
# ----------------------------------------------------------------------------------
@torch.no_grad()   # native extension ops build no autograd graph
def raw2alpha(density, shift, interval):
    """render_utils_kernel.cu:358-370."""
    e = torch.exp(density + _f(shift))
    alpha = 1 - torch.pow(1 + e, -_f(interval))
    return [e, alpha]


@torch.no_grad()   # native extension ops build no autograd graph
def raw2alpha_backward(exp_d, grad_back, interval):
    """render_utils_kernel.cu:396-406: min(e,1e10) and the product are evaluated in double."""
    itv = float(np.float32(interval))
    p = torch.pow(1 + exp_d, _f(np.float32(-itv) - np.float32(1)))
    g = torch.clamp_max(exp_d.double(), 1e10) * p.double() * itv * grad_back.double()
    return g.to(F32)


# ----------------------------------------------------------------------------------
# alpha2weight
# ----------------------------------------------------------------------------------
def _segments(ray_id, n_rays):
    """render_utils_kernel.cu:461-471 + :489."""
    n = len(ray_id)
    i_start = torch.zeros(n_rays, dtype=torch.int64)
    i_end = torch.zeros(n_rays, dtype=torch.int64)
    if n == 0:
        return i_start, i_end
    chg = torch.nonzero(ray_id[1:] != ray_id[:-1]).flatten() + 1
    i_start[ray_id[chg]] = chg
    i_end[ray_id[chg - 1]] = chg
    i_end[ray_id[n - 1]] = n
    return i_start, i_end


@torch.no_grad()   # native extension ops build no autograd graph
def alpha2weight(alpha, ray_id, n_rays):
    """render_utils_kernel.cu:431-459,473-505.

    Step-synchronous vectorised loop: every ray advances one sample per iteration, so each
    ray's product is taken in the reference's order; T_cum is float, updated through a
    double product (`T_cum *= (1. - alpha[i])`), stop when T_cum < 1e-3.
    """
    alpha = alpha.to(F32)
    n = len(alpha)
    weight = torch.zeros_like(alpha)
    T = torch.ones_like(alpha)
    last = torch.ones(n_rays, dtype=F32)
    i_start, i_end = _segments(ray_id, n_rays)
    if n == 0:
        return [weight, T, last, i_start, i_end]
    pos = i_start.clone()
    act = torch.nonzero(pos < i_end).flatten()
    t_cum = torch.ones(n_rays, dtype=F32)
    while len(act):
        idx = pos[act]
        a = alpha[idx]
        tc = t_cum[act]
        T[idx] = tc
        weight[idx] = tc * a
        tn = (tc.double() * (1.0 - a.double())).to(F32)
        t_cum[act] = tn
        pos[act] = idx + 1
        keep = (tn.double() >= 1e-3) & (pos[act] < i_end[act])
        act = act[keep]
    return [weight, T, t_cum, i_start, pos]


@torch.no_grad()   # native extension ops build no autograd graph
def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    """render_utils_kernel.cu:508-531."""
    grad = torch.zeros_like(alpha)
    if n_rays == 0 or len(alpha) == 0:
        return grad
    back = (grad_last * alphainv_last).to(F32).clone()
    pos = i_end.clone() - 1
    act = torch.nonzero(pos >= i_start).flatten()
    # rays whose i_start==i_end==0 have pos=-1 < 0 = i_start: inactive, as in the kernel
    while len(act):
        idx = pos[act]
        gw = grad_weights[idx]
        b = back[act]
        denom = (1 - alpha[idx]).double() + 1e-10
        grad[idx] = ((gw * T[idx]).double() - b.double() / denom).to(F32)
        back[act] = b + gw * weight[idx]
        pos[act] = idx - 1
        act = act[pos[act] >= i_start[act]]
    return grad


# ----------------------------------------------------------------------------------
# Adam (lib/cuda/adam_upd_kernel.cu; host scalar math in float as in :72)
# ----------------------------------------------------------------------------------
def _step_size(step, beta1, beta2, lr):
    f = np.float32
    b1, b2, lr = f(beta1), f(beta2), f(lr)
    return f(lr * f(np.sqrt(f(1) - f(np.power(b2, f(step))))) / f(f(1) - f(np.power(b1, f(step)))))


def _adam_core(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps, mask=None, perlr=None):
    f = np.float32
    ss = float(_step_size(step, beta1, beta2, lr))
    b1, b2, eps = float(f(beta1)), float(f(beta2)), float(f(eps))
    omb1, omb2 = float(f(1) - f(beta1)), float(f(1) - f(beta2))
    m = _f(b1) * exp_avg + _f(omb1) * grad
    v = _f(b2) * exp_avg_sq + (_f(omb2) * grad) * grad
    num = _f(ss) * m if perlr is None else (_f(ss) * perlr) * m
    p = param - num / (_sqrt_rn(v) + _f(eps))
    if mask is None:
        exp_avg.copy_(m), exp_avg_sq.copy_(v), param.copy_(p)
    else:
        exp_avg.copy_(torch.where(mask, m, exp_avg))
        exp_avg_sq.copy_(torch.where(mask, v, exp_avg_sq))
        param.copy_(torch.where(mask, p, param))


@torch.no_grad()
def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam_core(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps)


@torch.no_grad()
def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam_core(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps, mask=(grad != 0))


@torch.no_grad()
def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    _adam_core(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps, perlr=perlr)


render_utils_namespace = SimpleNamespace(
    infer_t_minmax=lambda *a: list(infer_t_minmax(*a)),
    infer_n_samples=infer_n_samples,
    infer_ray_start_dir=lambda *a: list(infer_ray_start_dir(*a)),
    sample_pts_on_rays=sample_pts_on_rays,
    raw2alpha=raw2alpha,
    raw2alpha_backward=raw2alpha_backward,
    alpha2weight=alpha2weight,
    alpha2weight_backward=alpha2weight_backward,
)
adam_namespace = SimpleNamespace(adam_upd=adam_upd, masked_adam_upd=masked_adam_upd,
                                 adam_upd_with_perlr=adam_upd_with_perlr)
