"""GPU parity against the REFERENCE'S OWN compiled CUDA (oracle/_ref/, built by oracle/build_ref.py from the unmodified
/root/reference/lib/cuda sources for sm_100a): the secondary oracle of SURVEY.md §8(c)(iii).

Pins, on the B200 itself:
  * oracle/dvgo_ops.py (the CPU restatement every other test leans on) against the real kernels it restates;
  * the reference-compatible ops of the product (render_utils_cuda.*, adam_upd_cuda.*: lib/cuda/render_utils.cpp:144-155,
    lib/cuda/adam_upd.cpp:79-86) against the real kernels, bit for bit where the contract is bit-exact;
  * the fused product kernels (apn_composite_fwd/bwd, MaskedAdam -> apn_adam_multi) against chains of the real kernels.
Skipped (with the reason) when the reference extensions were not prebuilt into oracle/_ref/.
"""
import json
import os

import pytest
import torch

from conftest import ROOT, RTOL, rel_err

pytestmark = pytest.mark.gpu

# The reference's kernels are compiled by nvcc with its default FMA contraction, the CPU restatement (torch CPU ops) and
# the product's contract arithmetic are not contracted: where an expression has the shape a*b + c the real kernels may
# differ from both in the last bit.  Every comparison below is recorded (count of differing elements, max |diff|) in
# gpurun_out/ref_kernel_report.json; the assertions hold what was measured on the B200 (see the report under profiles/).
REPORT = {}


def _cmp(tag, a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape and a.dtype == b.dtype, (tag, a.shape, b.shape, a.dtype, b.dtype)
    neq = int((a != b).sum())
    mx = float((a.double() - b.double()).abs().max()) if a.numel() else 0.0
    REPORT[tag] = {"n": a.numel(), "n_differ": neq, "max_abs_diff": mx, "scale": float(b.double().abs().max()) if b.numel() else 0.0}
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "ref_kernel_report.json"), "w") as fh:
            json.dump(REPORT, fh, indent=1)
    except OSError:
        pass
    return neq


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref/*.so not built (python -m oracle.build_ref needs /root/reference)")
    return build_ref.load("render_utils_cuda"), build_ref.load("adam_upd_cuda")


def _rays(golden_tiny, n=None):
    g = golden_tiny
    ro, rd = g["rays_o"].cuda(), g["rays_d"].cuda()
    if n:
        ro, rd = ro[:n].contiguous(), rd[:n].contiguous()
    lo = (g["canonical_pcd"].min(0)[0] - 0.01).cuda()
    hi = (g["canonical_pcd"].max(0)[0] + 0.01).cuda()
    return ro, rd, lo, hi


def test_sampler_vs_reference_kernels(ref, golden_tiny):
    """sample_pts_on_rays: positions, masks, ids, step counts, t_min/t_max — product op, CPU restatement and the
    reference's kernels agree bit for bit."""
    from articulated_point_nerf_b200 import render_utils_cuda as ru
    from articulated_point_nerf_b200.scene import make_scene
    from oracle import dvgo_ops
    ru_ref, _ = ref
    scene = make_scene(golden_tiny["config"])
    ro, rd, lo, hi = _rays(golden_tiny)
    stepdist = scene.cfg.stepsize * scene.voxel_size
    a = ru_ref.sample_pts_on_rays(ro, rd, lo, hi, scene.cfg.near, scene.cfg.far, stepdist)
    b = ru.sample_pts_on_rays(ro, rd, lo, hi, scene.cfg.near, scene.cfg.far, stepdist)
    c = dvgo_ops.sample_pts_on_rays(ro.cpu(), rd.cpu(), lo.cpu(), hi.cpu(), scene.cfg.near, scene.cfg.far, stepdist)
    names = ["pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"]
    assert len(a) == len(b) == len(c) == 7
    for n, x, y, z in zip(names, a, b, c):
        assert x.shape == y.shape == z.shape, n
        d_k, d_c = _cmp(f"sampler.{n}: product vs reference kernel", y, x), _cmp(f"sampler.{n}: CPU restatement vs reference kernel", z, x.cpu())
        if x.dtype in (torch.int64, torch.bool) and n != "mask_outbbox":
            assert d_k == 0 and d_c == 0, n                       # structure: identical
        elif n == "mask_outbbox":
            # a sample whose coordinate lies within an ulp of a bbox face may flip (a*b+c contraction in the real kernel)
            # measured on the B200: 34 of 4876 (0.7 %), the same 34 for the product and for the CPU restatement
            assert d_k <= x.numel() // 50 and d_c <= x.numel() // 50, (n, d_k, d_c)
            assert torch.equal(y.cpu(), z), "product and CPU restatement agree with each other bit for bit"
        else:
            assert rel_err(y, x) < 5e-7 and rel_err(z, x.cpu()) < 5e-7, n
    # the three infer_* helpers of the pybind surface
    t0, t1 = ru_ref.infer_t_minmax(ro, rd, lo, hi, scene.cfg.near, scene.cfg.far)
    p0, p1 = ru.infer_t_minmax(ro, rd, lo, hi, scene.cfg.near, scene.cfg.far)
    _cmp("infer_t_minmax.t_min", p0, t0), _cmp("infer_t_minmax.t_max", p1, t1)
    assert rel_err(p0, t0) < 5e-7 and rel_err(p1, t1) < 5e-7
    assert _cmp("infer_n_samples", ru.infer_n_samples(t0, t1, stepdist), ru_ref.infer_n_samples(t0, t1, stepdist)) == 0
    s0, d0 = ru_ref.infer_ray_start_dir(ro, rd, t0)
    s1, d1 = ru.infer_ray_start_dir(ro, rd, t0)
    _cmp("infer_ray_start_dir.start", s1, s0), _cmp("infer_ray_start_dir.dir", d1, d0)
    assert rel_err(s1, s0) < 5e-7 and rel_err(d1, d0) < 5e-7


def test_raw2alpha_vs_reference_kernels(ref):
    from articulated_point_nerf_b200 import render_utils_cuda as ru
    from oracle import dvgo_ops
    ru_ref, _ = ref
    gen = torch.Generator().manual_seed(0)
    density = (torch.randn(100003, generator=gen) * 6).cuda()
    shift, interval = -6.9, 0.5
    e0, a0 = ru_ref.raw2alpha(density, shift, interval)
    e1, a1 = ru.raw2alpha(density, shift, interval)
    e2, a2 = dvgo_ops.raw2alpha(density.cpu(), shift, interval)
    # exp / pow come from different math libraries (CUDA libdevice vs glibc): a few ulp, far inside the 1e-4 bar
    _cmp("raw2alpha.alpha: product vs reference kernel", a1, a0), _cmp("raw2alpha.exp: product vs reference kernel", e1, e0)
    _cmp("raw2alpha.alpha: CPU restatement vs reference kernel", a2, a0.cpu())
    assert rel_err(a1, a0) < 1e-6 and rel_err(e1, e0) < 1e-6
    assert rel_err(a2, a0) < 1e-6 and rel_err(e2, e0) < 1e-6
    gb = torch.randn(100003, generator=gen).cuda()
    g0 = ru_ref.raw2alpha_backward(e0, gb, interval)
    g1 = ru.raw2alpha_backward(e0, gb, interval)
    g2 = dvgo_ops.raw2alpha_backward(e0.cpu(), gb.cpu(), interval)
    _cmp("raw2alpha_backward: product vs reference kernel", g1, g0), _cmp("raw2alpha_backward: CPU restatement vs reference kernel", g2, g0.cpu())
    assert rel_err(g1, g0) < 1e-6 and rel_err(g2, g0) < 1e-6


def _ragged(seed=1, n_rays=5000, max_len=40, empty_frac=0.3, opaque_frac=0.1):
    gen = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, max_len, (n_rays,), generator=gen)
    lens[torch.rand(n_rays, generator=gen) < empty_frac] = 0
    ray_id = torch.repeat_interleave(torch.arange(n_rays), lens)
    alpha = torch.rand(len(ray_id), generator=gen) * 0.5
    big = torch.rand(len(ray_id), generator=gen) < opaque_frac        # triggers the T < 1e-3 early stop
    alpha[big] = 0.97 + 0.03 * torch.rand(int(big.sum()), generator=gen)
    return alpha, ray_id, lens, gen


def test_alpha2weight_vs_reference_kernels(ref):
    """alpha2weight (+backward): the compat op, the CPU restatement and the reference's kernel give identical bits
    (per-ray serial product in float, lib/cuda/render_utils_kernel.cu:445-457,506-530)."""
    from articulated_point_nerf_b200 import render_utils_cuda as ru
    from oracle import dvgo_ops
    ru_ref, _ = ref
    alpha, ray_id, lens, gen = _ragged()
    n_rays = len(lens)
    a_d, r_d = alpha.cuda(), ray_id.cuda()
    out_ref = ru_ref.alpha2weight(a_d, r_d, n_rays)
    out_k = ru.alpha2weight(a_d, r_d, n_rays)
    out_c = dvgo_ops.alpha2weight(alpha, ray_id, n_rays)
    names = ["weight", "T", "alphainv_last", "i_start", "i_end"]
    for n, x, y, z in zip(names, out_ref, out_k, out_c):
        # forward: products only (T *= 1 - alpha; w = T * alpha) — nothing to contract, identical bits
        assert _cmp(f"alpha2weight.{n}: product vs reference kernel", y, x) == 0, n
        assert _cmp(f"alpha2weight.{n}: CPU restatement vs reference kernel", z, x.cpu()) == 0, n
    gw = torch.randn(len(alpha), generator=gen).cuda()
    gl = torch.randn(n_rays, generator=gen).cuda()
    w, T, last, i0, i1 = out_ref
    g_ref = ru_ref.alpha2weight_backward(a_d, w, T, last, i0, i1, n_rays, gw, gl)
    g_k = ru.alpha2weight_backward(a_d, w, T, last, i0, i1, n_rays, gw, gl)
    g_c = dvgo_ops.alpha2weight_backward(alpha, w.cpu(), T.cpu(), last.cpu(), i0.cpu(), i1.cpu(), n_rays, gw.cpu(), gl.cpu())
    # backward: grad = gw*T - cum/(1-alpha+1e-10), cum += gw*w: shapes a*b+c the real kernel may contract
    _cmp("alpha2weight_backward: product vs reference kernel", g_k, g_ref)
    _cmp("alpha2weight_backward: CPU restatement vs reference kernel", g_c, g_ref.cpu())
    assert rel_err(g_k, g_ref) < 1e-6 and rel_err(g_c, g_ref.cpu()) < 1e-6


def test_fused_compositing_vs_chain_of_reference_kernels(ref):
    """apn_composite_fwd/bwd (pre-mask + Alphas2Weights + post-mask + segment sums in one kernel each) against the
    reference's chain: boolean masks (torch) -> the REAL alpha2weight kernel -> index_add (lib/temporalpoints.py:611-677)."""
    from articulated_point_nerf_b200 import ops
    ru_ref, _ = ref
    alpha, ray_id, lens, gen = _ragged(seed=2)
    alpha[torch.rand(len(alpha), generator=gen) < 0.15] *= 1e-4          # exercises the alpha <= thres mask
    n_rays = len(lens)
    rgb = torch.rand(len(alpha), 3, generator=gen)
    step_id = torch.cat([torch.arange(int(l)) for l in lens]) if len(alpha) else torch.zeros(0, dtype=torch.long)
    thres, bg = 1e-4, 1.0
    d = "cuda"

    class A2W(torch.autograd.Function):           # lib/tineuvox.py:627-643 on the reference's kernels
        @staticmethod
        def forward(ctx, a, rid, n):
            w, T, last, i0, i1 = ru_ref.alpha2weight(a, rid, n)
            ctx.save_for_backward(a, w, T, last, i0, i1)
            ctx.n = n
            return w, last

        @staticmethod
        def backward(ctx, gw, gl):
            a, w, T, last, i0, i1 = ctx.saved_tensors
            return ru_ref.alpha2weight_backward(a, w, T, last, i0, i1, ctx.n, gw.contiguous(), gl.contiguous()), None, None

    a_r = alpha.to(d).requires_grad_(True)
    c_r = rgb.to(d).requires_grad_(True)
    rid, sid = ray_id.to(d), step_id.to(d)
    m1 = torch.where(a_r > thres)[0]
    a1, c1, rid1, sid1 = a_r[m1], c_r[m1], rid[m1], sid[m1]
    w, last = A2W.apply(a1.contiguous(), rid1.contiguous(), n_rays)
    m2 = torch.where(w > thres)[0]
    w2, c2, rid2, sid2 = w[m2], c1[m2], rid1[m2], sid1[m2]
    rgb_ref = torch.zeros(n_rays, 3, device=d).index_add_(0, rid2, w2[:, None] * c2) + last[:, None] * bg
    depth_ref = torch.zeros(n_rays, device=d).index_add_(0, rid2, w2 * sid2)

    a_k = alpha.to(d).requires_grad_(True)
    c_k = rgb.to(d).requires_grad_(True)
    ray_start = torch.zeros(n_rays + 1, dtype=torch.int32)
    ray_start[1:] = torch.cumsum(lens, 0).int()
    rgb_k, last_k, depth_k, _ = ops.composite(a_k, c_k, step_id.int().to(d), ray_start.to(d), n_rays, thres, bg, want_depth=True)
    assert _cmp("composite.alphainv_last: fused product kernel vs reference chain", last_k, last.detach()) == 0   # bit-exact
    assert rel_err(rgb_k, rgb_ref) < 1e-6 and rel_err(depth_k, depth_ref) < 1e-6      # sums: order-free
    w1 = torch.randn(n_rays, 3, generator=gen).to(d)
    w2_ = torch.randn(n_rays, generator=gen).to(d)
    w3 = torch.randn(n_rays, generator=gen).to(d)
    ((rgb_ref * w1).sum() + (last * w2_).sum() + (depth_ref * w3).sum()).backward()
    ((rgb_k * w1).sum() + (last_k * w2_).sum() + (depth_k * w3).sum()).backward()
    assert rel_err(a_k.grad, a_r.grad) < RTOL
    assert rel_err(c_k.grad, c_r.grad) < RTOL


def test_adam_vs_reference_kernels(ref):
    """adam_upd / masked_adam_upd / adam_upd_with_perlr: compat ops and the multi-tensor MaskedAdam against the
    reference's kernels (lib/cuda/adam_upd_kernel.cu:9-132), identical bits after three steps."""
    from articulated_point_nerf_b200 import MaskedAdam, adam_upd_cuda
    from oracle import dvgo_ops
    _, ad_ref = ref
    gen = torch.Generator().manual_seed(3)
    n = 70001
    p0 = torch.randn(n, generator=gen)
    g0 = torch.randn(n, generator=gen) * 0.1
    g0[torch.rand(n, generator=gen) < 0.4] = 0
    pl = torch.rand(n, generator=gen)
    for kind in ("adam_upd", "masked_adam_upd", "adam_upd_with_perlr"):
        pr, mr, vr = p0.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
        pk, mk, vk = p0.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
        pc, mc, vc = p0.clone(), torch.zeros(n), torch.zeros(n)
        for step in (1, 2, 3):
            g = (g0 * step).cuda()
            extra = (pl.cuda(),) if kind == "adam_upd_with_perlr" else ()
            getattr(ad_ref, kind)(pr, g, mr, vr, *extra, step, 0.9, 0.99, 1e-3, 1e-8)
            getattr(adam_upd_cuda, kind)(pk, g, mk, vk, *extra, step, 0.9, 0.99, 1e-3, 1e-8)
            getattr(dvgo_ops, kind)(pc, g0 * step, mc, vc, *((pl,) if extra else ()), step, 0.9, 0.99, 1e-3, 1e-8)
        for nm, k_, c_, r_ in (("param", pk, pc, pr), ("exp_avg", mk, mc, mr), ("exp_avg_sq", vk, vc, vr)):
            _cmp(f"{kind}.{nm}: product vs reference kernel", k_, r_)
            _cmp(f"{kind}.{nm}: CPU restatement vs reference kernel", c_, r_.cpu())
            # m = b1*m + (1-b1)*g and v = b2*v + (1-b2)*g*g are a*b+c shapes: last-bit differences allowed, nothing more
            assert rel_err(k_, r_) < 5e-7 and rel_err(c_, r_.cpu()) < 5e-7, (kind, nm)
    # the optimiser (one multi-tensor launch for both groups) against per-tensor launches of the reference kernels
    pa = torch.nn.Parameter(p0.clone().cuda())
    pb = torch.nn.Parameter((p0 * 0.5).cuda())
    opt = MaskedAdam([{"params": [pa], "lr": 1e-3, "skip_zero_grad": False}, {"params": [pb], "lr": 5e-4, "skip_zero_grad": True}])
    ra, rb = p0.clone().cuda(), (p0 * 0.5).cuda()
    st = [torch.zeros(n).cuda() for _ in range(4)]
    for step in (1, 2, 3):
        g = (g0 * step).cuda()
        pa.grad, pb.grad = g.clone(), g.clone()
        opt.step()
        ad_ref.adam_upd(ra, g, st[0], st[1], step, 0.9, 0.99, 1e-3, 1e-8)
        ad_ref.masked_adam_upd(rb, g, st[2], st[3], step, 0.9, 0.99, 5e-4, 1e-8)
    _cmp("MaskedAdam(group 0, plain): product vs reference kernel", pa.detach(), ra)
    _cmp("MaskedAdam(group 1, skip_zero_grad): product vs reference kernel", pb.detach(), rb)
    assert rel_err(pa.detach(), ra) < 5e-7 and rel_err(pb.detach(), rb) < 5e-7
