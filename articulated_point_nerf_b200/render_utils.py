"""Reference-compatible native-op surface: the functions the reference binds with pybind in
lib/cuda/render_utils.cpp:144-155 and lib/cuda/adam_upd.cpp:79-86, here served by libapn_sm100.so,
plus the two autograd wrappers of lib/tineuvox.py:627-670.

    render_utils_cuda.sample_pts_on_rays / infer_* / raw2alpha(+backward) / alpha2weight(+backward)
    adam_upd_cuda.adam_upd / masked_adam_upd / adam_upd_with_perlr

Same argument order, same return tuples, int64 ids, CUDA + contiguous inputs required
(RuntimeError otherwise, like CHECK_INPUT in lib/cuda/render_utils.cpp:40-42).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import _lib, ops
from ._lib import check, ptr, stream


def _chk(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("input must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError("input must be contiguous")
        if t.is_floating_point() and t.dtype != torch.float32:
            raise RuntimeError("the B200 path computes in fp32")


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    _chk(rays_o, rays_d, xyz_min, xyz_max)
    R = rays_o.shape[0]
    t_min = torch.empty(R, device=rays_o.device)
    t_max = torch.empty(R, device=rays_o.device)
    check(_lib.load().apn_infer_t_minmax(ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max), float(near), float(far), R,
                                         ptr(t_min), ptr(t_max), stream()), "apn_infer_t_minmax")
    return [t_min, t_max]


def infer_n_samples(t_min, t_max, stepdist):
    _chk(t_min, t_max)
    R = t_min.shape[0]
    n = torch.empty(R, device=t_min.device, dtype=torch.int64)
    check(_lib.load().apn_infer_n_samples(ptr(t_min), ptr(t_max), float(stepdist), R, ptr(n), stream()),
          "apn_infer_n_samples")
    return n


def infer_ray_start_dir(rays_o, rays_d, t_min):
    _chk(rays_o, rays_d, t_min)
    R = rays_o.shape[0]
    start = torch.empty(R, 3, device=rays_o.device)
    direc = torch.empty(R, 3, device=rays_o.device)
    check(_lib.load().apn_infer_ray_start_dir(ptr(rays_o), ptr(rays_d), ptr(t_min), R, ptr(start), ptr(direc), stream()),
          "apn_infer_ray_start_dir")
    return [start, direc]


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """lib/cuda/render_utils_kernel.cu:190-236 -> [pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max]."""
    _chk(rays_o, rays_d, xyz_min, xyz_max)
    assert rays_o.dim() == 2 and rays_o.shape[1] == 3
    R = rays_o.shape[0]
    dev = rays_o.device
    t_min, t_max = infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far)
    n_steps = infer_n_samples(t_min, t_max, stepdist)
    cum = n_steps.cumsum(0)
    total = int(cum[-1].item()) if R else 0     # the reference synchronises here too (N_steps.sum().item())
    start, direc = infer_ray_start_dir(rays_o, rays_d, t_min)
    pts = torch.empty(total, 3, device=dev)
    mask = torch.empty(total, device=dev, dtype=torch.bool)
    ray_id = torch.empty(total, device=dev, dtype=torch.int64)
    step_id = torch.empty(total, device=dev, dtype=torch.int64)
    check(_lib.load().apn_sample_pts_on_rays_fill(ptr(start), ptr(direc), ptr(xyz_min), ptr(xyz_max), ptr(cum), R, total,
                                                  float(stepdist), ptr(pts), ptr(mask), ptr(ray_id), ptr(step_id), stream()),
          "apn_sample_pts_on_rays_fill")
    return [pts, mask, ray_id, step_id, n_steps, t_min, t_max]


def raw2alpha(density, shift, interval):
    _chk(density)
    n = density.numel()
    e = torch.empty_like(density)
    a = torch.empty_like(density)
    check(_lib.load().apn_raw2alpha(ptr(density), float(shift), float(interval), n, ptr(e), ptr(a), stream()), "apn_raw2alpha")
    return [e, a]


def raw2alpha_backward(exp_d, grad_back, interval):
    _chk(exp_d, grad_back)
    g = torch.empty_like(exp_d)
    check(_lib.load().apn_raw2alpha_backward(ptr(exp_d), ptr(grad_back), float(interval), exp_d.numel(), ptr(g), stream()),
          "apn_raw2alpha_backward")
    return g


def alpha2weight(alpha, ray_id, n_rays):
    _chk(alpha, ray_id)
    assert alpha.dim() == 1 and ray_id.dim() == 1 and alpha.shape[0] == ray_id.shape[0]
    if ray_id.dtype != torch.int64:
        raise RuntimeError("ray_id must be int64")
    n, dev = alpha.shape[0], alpha.device
    weight = torch.zeros_like(alpha)
    T = torch.ones_like(alpha)
    last = torch.ones(n_rays, device=dev)
    i_start = torch.zeros(n_rays, device=dev, dtype=torch.int64)
    i_end = torch.zeros(n_rays, device=dev, dtype=torch.int64)
    check(_lib.load().apn_alpha2weight(ptr(alpha), ptr(ray_id), n, int(n_rays), ptr(weight), ptr(T), ptr(last), ptr(i_start),
                                       ptr(i_end), stream()), "apn_alpha2weight")
    return [weight, T, last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    _chk(alpha, weight, T, alphainv_last, i_start, i_end, grad_weights, grad_last)
    grad = torch.zeros_like(alpha)
    check(_lib.load().apn_alpha2weight_backward(ptr(alpha), ptr(weight), ptr(T), ptr(alphainv_last), ptr(i_start), ptr(i_end),
                                                int(n_rays), ptr(grad_weights), ptr(grad_last), ptr(grad), stream()),
          "apn_alpha2weight_backward")
    return grad


def _adam(mode):
    def fn(param, grad, exp_avg, exp_avg_sq, *rest):
        if mode == 2:
            perlr, step, beta1, beta2, lr, eps = rest
        else:
            perlr = None
            step, beta1, beta2, lr, eps = rest
        _chk(param, grad, exp_avg, exp_avg_sq)
        ss = ops.adam_step_size(step, beta1, beta2, lr)
        ops.adam_multi([(param, grad, exp_avg, exp_avg_sq, perlr, ss, mode)], beta1, beta2, eps)
        torch.autograd.graph.increment_version([param, exp_avg, exp_avg_sq])     # in-place update through raw pointers
    return fn


render_utils_cuda = SimpleNamespace(
    infer_t_minmax=infer_t_minmax, infer_n_samples=infer_n_samples, infer_ray_start_dir=infer_ray_start_dir,
    sample_pts_on_rays=sample_pts_on_rays, raw2alpha=raw2alpha, raw2alpha_backward=raw2alpha_backward,
    alpha2weight=alpha2weight, alpha2weight_backward=alpha2weight_backward)

adam_upd_cuda = SimpleNamespace(adam_upd=_adam(0), masked_adam_upd=_adam(1), adam_upd_with_perlr=_adam(2))


class Raw2Alpha(torch.autograd.Function):
    """lib/tineuvox.py:646-670."""

    @staticmethod
    def forward(ctx, density, shift, interval):
        exp_d, alpha = raw2alpha(density.contiguous(), shift, interval)
        if density.requires_grad:
            ctx.save_for_backward(exp_d)
            ctx.interval = interval
        return alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_back):
        exp_d = ctx.saved_tensors[0]
        return raw2alpha_backward(exp_d, grad_back.contiguous(), ctx.interval), None, None


class Alphas2Weights(torch.autograd.Function):
    """lib/tineuvox.py:627-643."""

    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        weights, T, alphainv_last, i_start, i_end = alpha2weight(alpha.contiguous(), ray_id.contiguous(), N)
        if alpha.requires_grad:
            ctx.save_for_backward(alpha, weights, T, alphainv_last, i_start, i_end)
            ctx.n_rays = N
        return weights, alphainv_last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, alphainv_last, i_start, i_end = ctx.saved_tensors
        grad = alpha2weight_backward(alpha, weights, T, alphainv_last, i_start, i_end, ctx.n_rays,
                                     grad_weights.contiguous(), grad_last.contiguous())
        return grad, None, None
