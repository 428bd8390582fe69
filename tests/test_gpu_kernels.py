"""GPU parity of each kernel group against the CPU oracle (oracle/), through the C ABI (ctypes).

Bars (BASELINE.json north_star): neighbour / sample indices and every integer output bit-exact;
fp32 outputs and gradients within RTOL = 1e-4 of the oracle, relative to the tensor's scale.
"""
import numpy as np
import pytest
import torch

from conftest import RTOL, model_from_golden, oracle_for_scene, oracle_from_golden, rel_err

pytestmark = pytest.mark.gpu


def _ops():
    from articulated_point_nerf_b200 import ops
    return ops


# ----------------------------------------------------------------------------------------
# K1 linear blend skinning
# ----------------------------------------------------------------------------------------
def _lbs_case(N, J, seed, merge=False):
    g = torch.Generator().manual_seed(seed)
    raw = torch.rand(N, J, generator=g) * 2
    theta = torch.tensor([0.1])
    T = torch.eye(4).repeat(J, 1, 1)
    A = torch.linalg.qr(torch.randn(J, 3, 3, generator=g))[0]
    T[:, :3, :3] = A * torch.sign(torch.linalg.det(A))[:, None, None]
    T[:, :3, 3] = torch.randn(J, 3, generator=g) * 0.1
    xyz = torch.randn(N, 3, generator=g) * 0.3
    gt = torch.randn(3, generator=g) * 0.05
    rules = torch.arange(J)
    if merge:
        rules[J // 2] = 1
        rules[J - 1] = 0
    return raw, theta, T, xyz, gt, rules


def _lbs_oracle(raw, theta, T, xyz, gt, rules, eps=1e-6):
    from oracle.path_oracle import get_weights
    w = get_weights(raw, theta, eps, rules)
    G = (T * w[:, :, None, None]).sum(1)
    xh = torch.cat([xyz, torch.ones(len(xyz), 1)], -1)
    out = torch.bmm(G, xh.unsqueeze(-1)).squeeze(-1)[:, :3] + gt
    ginv = torch.inverse(G)[:, :3, :3]
    return out, ginv, w, G


# J = 16 / 17 / 80 / 81: bone-tile boundaries of the tensor-core backward (J <= 80) and its CUDA-core fallback (81, merge rules)
@pytest.mark.parametrize("N,J,merge", [(1, 2, False), (127, 21, False), (4099, 29, True), (20000, 65, False), (64, 16, False),
                                       (40, 17, False), (33, 80, False), (1000, 81, False)])
def test_lbs_forward_backward(N, J, merge):
    ops = _ops()
    raw, theta, T, xyz, gt, rules = _lbs_case(N, J, N + J, merge)
    leaves = [t.clone().requires_grad_(True) for t in (raw, theta, T, gt)]
    o_xyz, o_ginv, o_w, _ = _lbs_oracle(leaves[0], leaves[1], leaves[2], xyz, leaves[3], rules)
    gen = torch.Generator().manual_seed(1)
    c1, c2, c3 = torch.randn(N, 3, generator=gen), torch.randn(N, 3, 3, generator=gen), torch.randn(N, J, generator=gen)
    (o_xyz * c1).sum().add((o_ginv * c2).sum()).add((o_w * c3).sum()).backward()

    d = "cuda"
    cl = [t.detach().clone().to(d).requires_grad_(True) for t in (raw, theta, T, gt)]
    k_xyz, k_ginv, k_w, bbox = ops.lbs(cl[0], cl[1], cl[2], cl[3], xyz.to(d), rules=rules.to(d) if merge else None)
    assert rel_err(k_xyz, o_xyz) < 1e-5
    assert rel_err(k_ginv.view(N, 3, 3), o_ginv) < 1e-5
    assert rel_err(k_w, o_w) < 1e-5
    assert torch.equal(bbox.cpu(), torch.cat([k_xyz.min(0)[0], k_xyz.max(0)[0]]).cpu())
    (k_xyz * c1.to(d)).sum().add((k_ginv.view(N, 3, 3) * c2.to(d)).sum()).add((k_w * c3.to(d)).sum()).backward()
    for a, b, name in zip(cl, leaves, ["raw_w", "theta", "bone_T", "global_t"]):
        ref = b.grad if name != "bone_T" else b.grad * torch.tensor([1., 1., 1., 0.])[None, :, None]
        got = a.grad if name != "bone_T" else a.grad
        assert rel_err(got, ref) < RTOL, name


def test_lbs_preblended_weights_and_frames():
    """theta_weight=None: the PointWarper.forward contract (final weights in, frames out)."""
    ops = _ops()
    raw, theta, T, xyz, gt, rules = _lbs_case(513, 24, 5)
    o_xyz, o_ginv, o_w, o_G = _lbs_oracle(raw, theta, T, xyz, gt, rules)
    d = "cuda"
    k_xyz, k_ginv, k_w, bbox, k_G = ops.lbs(o_w.to(d), None, T.to(d), gt.to(d), xyz.to(d), want_frames=True)
    assert rel_err(k_xyz, o_xyz) < 1e-5
    o_G = o_G.clone()
    assert rel_err(k_G, o_G) < 1e-5
    assert torch.equal(k_w.cpu(), o_w)


# ----------------------------------------------------------------------------------------
# K2 grid + sampling + exact 8-NN
# ----------------------------------------------------------------------------------------
def _cloud_and_rays(config="tiny", view=0):
    from articulated_point_nerf_b200.scene import make_scene
    scene = make_scene(config)
    ro, rd, vd = [x.reshape(-1, 3).contiguous() for x in scene.rays(view)]
    return scene, ro, rd, vd


@pytest.mark.parametrize("search", ["warp", "thread0", "thread1", "sorted"])
@pytest.mark.parametrize("config", ["tiny", "small"])
def test_sample_and_knn_bit_exact(config, search, monkeypatch):
    """pts / ray_id / step_id / neighbour indices identical to the oracle's brute force (ties -> lower index)."""
    monkeypatch.setenv("APN_KNN_FORCE", search)          # both k-NN searches (csrc/grid_knn.cu) against the same oracle
    ops = _ops()
    from oracle.path_oracle import OraclePath
    scene, ro, rd, vd = _cloud_and_rays(config)
    cfg = scene.cfg
    g = torch.Generator().manual_seed(3)
    xyz = scene.canonical_pcd + torch.randn(scene.canonical_pcd.shape, generator=g) * 0.002
    orc = OraclePath.__new__(OraclePath)
    orc.K, orc.voxel_size = 8, scene.voxel_size
    ref = orc.sample_and_knn(xyz, ro, rd, cfg.near, cfg.far, cfg.stepsize, 0.01)
    d = "cuda"
    xyz_d = xyz.to(d)
    bbox = torch.cat([xyz_d.min(0)[0], xyz_d.max(0)[0]])
    grid = ops.Grid(xyz_d, bbox, 0.01, 0.01, 1.5 * scene.lattice_h)
    desc = grid.describe()
    assert desc["overflow"] == 0 and desc["n_points"] == len(xyz)
    smp, dbg = ops.sample_and_knn(grid, ro.to(d), rd.to(d), cfg.near, cfg.far, cfg.stepsize * scene.voxel_size, return_d2=True)
    assert smp.M == len(ref["pts"])
    assert torch.equal(smp.pts.cpu(), ref["pts"])
    assert torch.equal(smp.ray_id.cpu().long(), ref["ray_id"])
    assert torch.equal(smp.step_id.cpu().long(), ref["step_id"])
    assert torch.equal(smp.nn_idx.cpu().long(), ref["s_i"])
    # d2 of kept candidates are the oracle's bits
    keep = dbg["keep"].bool()
    assert torch.equal(dbg["d2"][keep].cpu(), ref["d2_all"][ref["keep"]])
    # ray_start is the CSR of ray_id
    rs = smp.ray_start.cpu().long()
    counts = torch.bincount(ref["ray_id"], minlength=len(ro))
    assert torch.equal(rs[1:] - rs[:-1], counts)


def test_grid_overflow_is_surfaced_and_recovered():
    """A cell table too small for the padded bbox must never be silent (background frames, index -1 in gathers):
    sample_and_knn and Grid.knn notice the header's overflow flag, grow the table in place and return the same
    results as a grid that was large enough from the start; at the capacity limit they raise."""
    ops = _ops()
    scene, ro, rd, vd = _cloud_and_rays("tiny")
    cfg = scene.cfg
    d = "cuda"
    xyz_d = scene.canonical_pcd.to(d)
    bbox = torch.cat([xyz_d.min(0)[0], xyz_d.max(0)[0]])
    good = ops.Grid(xyz_d, bbox, 0.01, 0.01, 1.5 * scene.lattice_h)
    small = ops.Grid(xyz_d, bbox, 0.01, 0.01, 1.5 * scene.lattice_h, cell_capacity=1024)
    assert small.overflowed() and not good.overflowed()
    a = ops.sample_and_knn(good, ro.to(d), rd.to(d), cfg.near, cfg.far, cfg.stepsize * scene.voxel_size)
    b = ops.sample_and_knn(small, ro.to(d), rd.to(d), cfg.near, cfg.far, cfg.stepsize * scene.voxel_size)
    assert small.cell_capacity > 1024 and not small.overflowed()
    assert a.M == b.M > 0 and torch.equal(a.nn_idx, b.nn_idx) and torch.equal(a.pts, b.pts)
    small2 = ops.Grid(xyz_d, bbox, 0.01, 0.01, 1.5 * scene.lattice_h, cell_capacity=1024)
    i1, d1 = good.knn(xyz_d[:500], 8)
    i2, d2 = small2.knn(xyz_d[:500], 8)
    assert int(i2.min()) >= 0 and torch.equal(i1, i2) and torch.equal(d1, d2)
    old_max = ops.Grid.MAX_CAPACITY
    try:
        ops.Grid.MAX_CAPACITY = 1024
        small3 = ops.Grid(xyz_d, bbox, 0.01, 0.01, 1.5 * scene.lattice_h, cell_capacity=1024)
        with pytest.raises(ops.GridOverflow):
            ops.sample_and_knn(small3, ro.to(d), rd.to(d), cfg.near, cfg.far, cfg.stepsize * scene.voxel_size)
    finally:
        ops.Grid.MAX_CAPACITY = old_max


def test_knn_points_matches_bruteforce_with_lattice_ties():
    """Self k-NN of the canonical lattice cloud (exact ties everywhere): lowest index wins."""
    ops = _ops()
    from oracle.path_oracle import knn_bruteforce
    scene, *_ = _cloud_and_rays("small")
    pcd = scene.canonical_pcd
    ref_d, ref_i = knn_bruteforce(pcd, pcd, 8)
    d = "cuda"
    p = pcd.to(d)
    grid = ops.Grid(p, torch.cat([p.min(0)[0], p.max(0)[0]]), 0.01, 0.01, scene.lattice_h)
    idx, d2 = grid.knn(p, 8)
    assert torch.equal(idx.cpu().long(), ref_i)
    assert torch.equal(d2.cpu(), ref_d)
    # far-away queries still get the exact answer (falls back to wider levels / full scan)
    q = torch.tensor([[3.0, 3.0, 3.0], [-2.0, 0.1, 0.4], [0.0, 0.0, 0.9]])
    rd_, ri_ = knn_bruteforce(q, pcd, 8)
    i2, d22 = grid.knn(q.to(d), 8)
    assert torch.equal(i2.cpu().long(), ri_)
    assert torch.equal(d22.cpu(), rd_)


def test_rays_missing_the_cloud_give_no_samples():
    ops = _ops()
    scene, ro, rd, vd = _cloud_and_rays("tiny")
    d = "cuda"
    p = scene.canonical_pcd.to(d)
    grid = ops.Grid(p, torch.cat([p.min(0)[0], p.max(0)[0]]), 0.01, 0.01, 1.5 * scene.lattice_h)
    smp = ops.sample_and_knn(grid, ro[:7].to(d), (-rd[:7]).contiguous().to(d), scene.cfg.near, scene.cfg.far,
                             scene.cfg.stepsize * scene.voxel_size)
    assert smp.M == 0 and smp.ray_start.cpu().tolist() == [0] * 8


# ----------------------------------------------------------------------------------------
# K3 aggregation (exact path)
# ----------------------------------------------------------------------------------------
def _agg_setup(golden_tiny, need_grad):
    ops = _ops()
    g = golden_tiny
    orc, cfg = oracle_from_golden(g)
    if need_grad:
        for v in orc.s.values():
            if v.is_floating_point():
                v.requires_grad_(True)
    with torch.set_grad_enabled(need_grad):
        wp = orc.warp(g["render"]["t"])
        Ginv = torch.inverse(wp["G"])
        smp = orc.sample_and_knn(wp["xyz"], g["rays_o"], g["rays_d"], cfg.near, cfg.far, cfg.stepsize, 0.01)
        xyz = wp["xyz"].detach().requires_grad_(need_grad)
        gi = Ginv.detach().requires_grad_(need_grad)
        o = orc.aggregate(xyz, gi, smp, g["viewdirs"], cfg.stepsize)
    model, scene = model_from_golden(g)
    d = "cuda"
    c = ops.AggConst(pts=smp["pts"].to(d), nn_idx=smp["s_i"].to(d).int().contiguous(), ray_id=smp["ray_id"].to(d).int(),
                     viewdirs=g["viewdirs"].to(d), canonical_alpha=model.canonical_alpha.detach(),
                     canonical_rgbs=model.canonical_rgbs.detach(), direct_eps=model.direct_eps.detach(),
                     mean_min_distance=float(g["mean_min_distance"]), eps=1e-6, act_shift=g["act_shift"],
                     interval=cfg.stepsize * g["voxel_size_ratio"], direct=True)
    return ops, orc, o, xyz, gi, smp, model, c


def test_aggregate_forward(golden_tiny):
    ops, orc, o, xyz, gi, smp, model, c = _agg_setup(golden_tiny, False)
    rgb, alpha, rgb_d, alpha_d, _ = o
    with torch.no_grad():
        k_alpha, k_rgb, k_ad, k_rd, k_idw = ops.aggregate(c, xyz.cuda(), gi[:, :3, :3].reshape(-1, 9).contiguous().cuda(),
                                                          model.canonical_feat, None, model._mlp_weights())
    assert rel_err(k_idw, orc.trace["idw"]) < 1e-5
    assert rel_err(k_alpha, alpha) < RTOL
    assert rel_err(k_rgb, rgb) < RTOL
    assert rel_err(k_ad, alpha_d) < RTOL
    assert rel_err(k_rd, rgb_d) < RTOL


def test_aggregate_backward(golden_tiny):
    ops, orc, o, xyz, gi, smp, model, c = _agg_setup(golden_tiny, True)
    rgb, alpha, _, _, _ = o
    gen = torch.Generator().manual_seed(2)
    ca, cr = torch.randn(alpha.shape, generator=gen), torch.randn(rgb.shape, generator=gen)
    ((alpha * ca).sum() + (rgb * cr).sum()).backward()
    kx = xyz.detach().cuda().requires_grad_(True)
    kg = gi.detach()[:, :3, :3].reshape(-1, 9).contiguous().cuda().requires_grad_(True)
    model.zero_grad()
    k_alpha, k_rgb, *_ = ops.aggregate(c, kx, kg, model.canonical_feat, None, model._mlp_weights())
    ((k_alpha * ca.cuda()).sum() + (k_rgb * cr.cuda()).sum()).backward()
    assert rel_err(kx.grad, xyz.grad) < RTOL
    assert rel_err(kg.grad.view(-1, 3, 3), gi.grad[:, :3, :3]) < RTOL
    named = dict(model.named_parameters())
    for k in ["canonical_feat", "feat_net.0.weight", "feat_net.0.bias", "feat_net.2.0.weight", "feat_net.3.0.bias",
              "feat_net.4.weight", "feat_net.4.bias", "densitynet.weight", "densitynet.bias",
              "rgbnet.feature_linears.weight", "rgbnet.feature_linears.bias", "rgbnet.views_linears.0.weight",
              "rgbnet.views_linears.0.bias", "rgbnet.views_linears.2.weight", "rgbnet.views_linears.2.bias"]:
        assert named[k].grad is not None, k
        assert rel_err(named[k].grad, orc.s[k].grad) < RTOL, k


# ----------------------------------------------------------------------------------------
# K4 compositing
# ----------------------------------------------------------------------------------------
def _ragged_rays(R, seed, max_len=40):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0, max_len, (R,), generator=g)
    lens[::7] = 0                       # empty rays
    ray_id = torch.repeat_interleave(torch.arange(R), lens)
    M = int(lens.sum())
    alpha = torch.rand(M, generator=g) ** 3
    alpha[torch.rand(M, generator=g) < 0.15] = 5e-5          # below the 1e-4 pre-mask
    alpha[torch.rand(M, generator=g) < 0.05] = 0.97          # triggers early stop quickly
    rgb = torch.rand(M, 3, generator=g)
    first = torch.cumsum(lens, 0) - lens
    step_id = torch.arange(M) - first[ray_id] + 3
    ray_start = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(lens, 0)])
    return alpha, rgb, ray_id, step_id, ray_start


@pytest.mark.parametrize("R,seed", [(1, 0), (257, 1), (5000, 2)])
def test_composite_forward_backward(R, seed):
    ops = _ops()
    from oracle.path_oracle import OraclePath
    alpha, rgb, ray_id, step_id, ray_start = _ragged_rays(R, seed)
    orc = OraclePath.__new__(OraclePath)
    orc.thres = 1e-4
    a = alpha.clone().requires_grad_(True)
    c = rgb.clone().requires_grad_(True)
    extra = torch.rand(len(alpha), 3, generator=torch.Generator().manual_seed(9))
    rgb_m, last, depth, ex, _, _ = orc.composite(a, c, ray_id, step_id.float(), R, 1.0, extra)
    gen = torch.Generator().manual_seed(4)
    w1, w2, w3 = torch.randn(R, 3, generator=gen), torch.randn(R, generator=gen), torch.randn(R, generator=gen)
    ((rgb_m * w1).sum() + (last * w2).sum() + (depth * w3).sum()).backward()
    d = "cuda"
    ka = alpha.to(d).requires_grad_(True)
    kc = rgb.to(d).requires_grad_(True)
    k_rgb, k_last, k_depth, k_ex = ops.composite(ka, kc, step_id.int().to(d), ray_start.int().to(d), R, 1e-4, 1.0,
                                                 extra=extra.to(d), want_depth=True)
    assert torch.equal(k_last.cpu(), last.detach())          # same fp32 product chain, same early stop
    assert rel_err(k_rgb, rgb_m) < 1e-6
    assert rel_err(k_depth, depth) < 1e-6
    assert rel_err(k_ex, ex + last.detach()[:, None] * 1.0) < 1e-6
    ((k_rgb * w1.to(d)).sum() + (k_last * w2.to(d)).sum() + (k_depth * w3.to(d)).sum()).backward()
    assert rel_err(ka.grad, a.grad) < RTOL
    assert rel_err(kc.grad, c.grad) < RTOL


# ----------------------------------------------------------------------------------------
# K4b Adam + reference-compatible single ops
# ----------------------------------------------------------------------------------------
def test_adam_bit_exact_all_modes():
    from articulated_point_nerf_b200 import MaskedAdam
    from oracle import dvgo_ops
    gen = torch.Generator().manual_seed(0)
    shapes = [(1,), (3, 5), (1031,), (128, 191), (4096 * 3 + 7,)]
    ps = [torch.randn(s, generator=gen) for s in shapes]
    gs = [torch.randn(s, generator=gen) * 0.1 for s in shapes]
    for g_ in gs:
        g_[torch.rand(g_.shape, generator=gen) < 0.4] = 0
    params = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    opt = MaskedAdam([{"params": params[:3], "lr": 1e-3, "skip_zero_grad": False},
                      {"params": params[3:], "lr": 5e-4, "skip_zero_grad": True}])
    ref_p = [p.clone() for p in ps]
    ref_m = [torch.zeros_like(p) for p in ps]
    ref_v = [torch.zeros_like(p) for p in ps]
    for step in range(1, 4):
        for p, g_ in zip(params, gs):
            p.grad = (g_ * step).cuda()
        opt.step()
        for i in range(len(ps)):
            fn = dvgo_ops.adam_upd if i < 3 else dvgo_ops.masked_adam_upd
            fn(ref_p[i], gs[i] * step, ref_m[i], ref_v[i], step, 0.9, 0.99, 1e-3 if i < 3 else 5e-4, 1e-8)
    for i, p in enumerate(params):
        assert torch.equal(p.detach().cpu(), ref_p[i]), i
        assert torch.equal(opt.state[p]["exp_avg"].cpu(), ref_m[i]), i
        assert torch.equal(opt.state[p]["exp_avg_sq"].cpu(), ref_v[i]), i


def test_adam_matches_reference_optimizer_golden(golden_tiny):
    from articulated_point_nerf_b200 import MaskedAdam
    a = golden_tiny["adam"]
    p0 = torch.nn.Parameter(a["before"][0].clone().cuda())
    p1 = torch.nn.Parameter(a["before"][1].clone().cuda())
    opt = MaskedAdam([{"params": [p0], "lr": a["lrs"][0], "skip_zero_grad": False},
                      {"params": [p1], "lr": a["lrs"][1], "skip_zero_grad": True}])
    p0.grad, p1.grad = a["grads"][0].cuda(), a["grads"][1].cuda()
    for _ in range(a["steps"]):
        opt.step()
    assert torch.equal(p0.detach().cpu(), a["after"][0])
    assert torch.equal(p1.detach().cpu(), a["after"][1])


def test_adam_follows_load_state_dict():
    """The cached descriptor table must see moment tensors replaced by optimizer.load_state_dict() (resume inside a
    live process): step, load a saved state, step -> same as an optimizer that never cached anything."""
    from articulated_point_nerf_b200 import MaskedAdam
    from oracle import dvgo_ops
    gen = torch.Generator().manual_seed(11)
    p0, g0 = torch.randn(515, generator=gen), torch.randn(515, generator=gen)
    p = torch.nn.Parameter(p0.clone().cuda())
    opt = MaskedAdam([{"params": [p], "lr": 1e-3, "skip_zero_grad": False}])
    p.grad = g0.cuda()
    opt.step()
    saved = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()}
    sd = opt.state_dict()
    sd = {"state": {0: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in sd["state"][0].items()}},
          "param_groups": sd["param_groups"]}
    p_after1 = p.detach().clone()
    opt.step()                                    # moves the live moments away from the saved ones
    with torch.no_grad():
        p.copy_(p_after1)
    opt.load_state_dict(sd)                       # replaces state[p]['exp_avg'/'exp_avg_sq'] by new tensors
    opt.step()
    rp, rm, rv = p_after1.cpu().clone(), saved["exp_avg"].cpu().clone(), saved["exp_avg_sq"].cpu().clone()
    dvgo_ops.adam_upd(rp, g0, rm, rv, 2, 0.9, 0.99, 1e-3, 1e-8)
    assert torch.equal(p.detach().cpu(), rp)
    assert torch.equal(opt.state[p]["exp_avg"].cpu(), rm) and torch.equal(opt.state[p]["exp_avg_sq"].cpu(), rv)


def test_adam_perlr_mode():
    from articulated_point_nerf_b200 import adam_upd_cuda
    from oracle import dvgo_ops
    gen = torch.Generator().manual_seed(5)
    p, g_, pl = torch.randn(777, generator=gen), torch.randn(777, generator=gen), torch.rand(777, generator=gen)
    m, v = torch.zeros(777), torch.zeros(777)
    kp, km, kv = p.clone().cuda(), m.clone().cuda(), v.clone().cuda()
    for step in (1, 2):
        dvgo_ops.adam_upd_with_perlr(p, g_, m, v, pl, step, 0.9, 0.99, 1e-2, 1e-8)
        adam_upd_cuda.adam_upd_with_perlr(kp, g_.cuda(), km, kv, pl.cuda(), step, 0.9, 0.99, 1e-2, 1e-8)
    assert torch.equal(kp.cpu(), p) and torch.equal(km.cpu(), m) and torch.equal(kv.cpu(), v)


def test_reference_compatible_render_utils(golden_tiny):
    """render_utils_cuda.{sample_pts_on_rays, raw2alpha(+bwd), alpha2weight(+bwd)} vs the restated reference kernels."""
    from articulated_point_nerf_b200 import render_utils_cuda as ru, Raw2Alpha, Alphas2Weights
    from oracle import dvgo_ops
    g = golden_tiny
    d = "cuda"
    lo = g["canonical_pcd"].min(0)[0] - 0.01
    hi = g["canonical_pcd"].max(0)[0] + 0.01
    stepdist = 0.5 * g["voxel_size"]
    ref = dvgo_ops.sample_pts_on_rays(g["rays_o"], g["rays_d"], lo, hi, 2.0, 6.0, stepdist)
    got = ru.sample_pts_on_rays(g["rays_o"].to(d), g["rays_d"].to(d), lo.to(d), hi.to(d), 2.0, 6.0, stepdist)
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)
    dens = torch.randn(5000, generator=torch.Generator().manual_seed(0)) * 4
    e_r, a_r = dvgo_ops.raw2alpha(dens, -6.9, 0.5)
    e_k, a_k = ru.raw2alpha(dens.to(d), -6.9, 0.5)
    assert rel_err(e_k, e_r) < 1e-6 and rel_err(a_k, a_r) < 1e-6
    gb = torch.randn(5000, generator=torch.Generator().manual_seed(1))
    assert rel_err(ru.raw2alpha_backward(e_k, gb.to(d), 0.5), dvgo_ops.raw2alpha_backward(e_r, gb, 0.5)) < 1e-5
    alpha, rgb, ray_id, step_id, ray_start = _ragged_rays(300, 7)
    w_r, T_r, l_r, s_r, en_r = dvgo_ops.alpha2weight(alpha, ray_id, 300)
    w_k, T_k, l_k, s_k, en_k = ru.alpha2weight(alpha.to(d), ray_id.to(d), 300)
    assert torch.equal(w_k.cpu(), w_r) and torch.equal(T_k.cpu(), T_r) and torch.equal(l_k.cpu(), l_r)
    assert torch.equal(s_k.cpu(), s_r) and torch.equal(en_k.cpu(), en_r)
    gw, gl = torch.randn(len(alpha)), torch.randn(300)
    g_r = dvgo_ops.alpha2weight_backward(alpha, w_r, T_r, l_r, s_r, en_r, 300, gw, gl)
    g_k = ru.alpha2weight_backward(alpha.to(d), w_k, T_k, l_k, s_k, en_k, 300, gw.to(d), gl.to(d))
    assert rel_err(g_k, g_r) < 1e-5
    # autograd wrappers (lib/tineuvox.py:627-670)
    x = (torch.randn(len(alpha), generator=torch.Generator().manual_seed(2)) * 4).to(d).requires_grad_(True)
    al = Raw2Alpha.apply(x, -6.9, 0.5)
    w, last = Alphas2Weights.apply(al, ray_id.to(d), 300)
    (w.sum() + last.sum()).backward()
    assert torch.isfinite(x.grad).all()
    # CPU tensors are rejected, like CHECK_INPUT
    with pytest.raises(RuntimeError):
        ru.raw2alpha(dens, -6.9, 0.5)


# ----------------------------------------------------------------------------------------
# K3 on the tcgen05 tensor cores
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [(1, RTOL), (0, 3e-2)])
def test_aggregate_tensor_core_path(golden_tiny, precision, tol):
    """Split-fp16 tcgen05 decoder == oracle within the 1e-4 parity bar; the single-pass fp16 mode within its own bound."""
    ops, orc, o, xyz, gi, smp, model, c = _agg_setup(golden_tiny, False)
    rgb, alpha, rgb_d, alpha_d, _ = o
    packed = ops.PackedDecoder()
    k_alpha, k_rgb, k_ad, k_rd, k_idw = ops.aggregate_tc(c, xyz.cuda(), gi[:, :3, :3].reshape(-1, 9).contiguous().cuda(),
                                                         model.canonical_feat, None, model._mlp_weights(), packed, precision)
    assert rel_err(k_idw, orc.trace["idw"]) < 1e-5
    assert rel_err(k_ad, alpha_d) < RTOL and rel_err(k_rd, rgb_d) < RTOL
    assert rel_err(k_alpha, alpha) < tol, rel_err(k_alpha, alpha)
    assert rel_err(k_rgb, rgb) < tol, rel_err(k_rgb, rgb)


@pytest.mark.parametrize("M,pose", [(1, False), (15, False), (16, True), (17, False), (4099, True)])
def test_aggregate_tensor_core_matches_fp32_kernel_on_ragged_sizes(M, pose):
    """Tile tails (M not a multiple of 16), more tiles than SMs, and the pose-embedding fold (d_in = 255)."""
    ops = _ops()
    g = torch.Generator().manual_seed(M)
    N, d = 3000, "cuda"
    xyz = torch.rand(N, 3, generator=g)
    A = torch.eye(3) + 0.2 * torch.randn(N, 3, 3, generator=g)
    feat = torch.relu(torch.randn(N, 128, generator=g)) * 0.5
    nn_idx = torch.randint(0, N, (M, 8), generator=g).int()
    pts = xyz[nn_idx[:, 0].long()] + 0.02 * torch.randn(M, 3, generator=g)
    ray_id = torch.sort(torch.randint(0, 50, (M,), generator=g))[0].int()
    vd = torch.nn.functional.normalize(torch.randn(50, 3, generator=g), dim=-1)
    d_in = 255 if pose else 191
    lin = [torch.nn.Linear(d_in, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 128),
           torch.nn.Linear(128, 1), torch.nn.Linear(128, 128), torch.nn.Linear(155, 64), torch.nn.Linear(64, 3)]
    ws = []
    for l in lin:
        ws += [l.weight.detach().to(d), l.bias.detach().to(d)]
    pe = (torch.randn(64, generator=g) * 0.3).to(d) if pose else None
    c = ops.AggConst(pts=pts.to(d), nn_idx=nn_idx.to(d), ray_id=ray_id.to(d), viewdirs=vd.to(d),
                     canonical_alpha=torch.rand(N, generator=g).to(d), canonical_rgbs=torch.rand(N, 3, generator=g).to(d),
                     direct_eps=torch.full((N,), 0.05).to(d), mean_min_distance=0.02, eps=1e-6, act_shift=0.0, interval=0.5)
    # act_shift = 0 keeps alpha = 1 - (1 + e^d)^-0.5 away from the cancellation regime where one ulp of 1.0 is 1e-4 of alpha
    args = (c, xyz.to(d), A.reshape(N, 9).contiguous().to(d), feat.to(d), pe)
    with torch.no_grad():
        ref = ops.aggregate(*args, ws)
    got = ops.aggregate_tc(*args, ws, ops.PackedDecoder(), precision=1)
    for a, b, name in zip(got, ref, ["alpha", "rgb", "alpha_direct", "rgb_direct", "idw"]):
        assert rel_err(a, b) < RTOL, (name, rel_err(a, b))
    fast = ops.aggregate_tc(*args, ws, ops.PackedDecoder(), precision=0)
    assert rel_err(fast[0], ref[0]) < 3e-2 and rel_err(fast[1], ref[1]) < 3e-2


def _kink_free_sample_mask(orc, xyz, gi, smp, tol=1e-5):
    """Samples none of whose 8 x 384 feat_net pre-activations (layers 0-2) lies within `tol` of the LeakyReLU kink.
    A kernel that is not bit-identical to the reference cannot agree on the derivative AT the kink (a 1e-7 rounding
    difference flips 1 <-> 0.01); everywhere else it must meet the 1e-4 bar."""
    from oracle.path_oracle import poc_fre
    with torch.no_grad():
        s_i, pts = smp["s_i"], smp["pts"]
        rel_p = pts[:, None, :] - xyz.detach()[s_i, :]
        frames = gi.detach()[s_i]
        rel_c = torch.bmm(frames[..., :3, :3].reshape(-1, 3, 3), rel_p.reshape(-1, 3).unsqueeze(-1)).squeeze(-1)
        x = torch.cat([poc_fre(rel_c, orc.pos_poc), orc.s["canonical_feat"][s_i, :].reshape(-1, 128)], -1)
        near = torch.zeros(len(x), dtype=torch.bool)
        for name in ["feat_net.0", "feat_net.2.0", "feat_net.3.0", "feat_net.4"]:
            pre = torch.nn.functional.linear(x, orc.s[name + ".weight"], orc.s[name + ".bias"])
            near |= (pre.abs() < tol).any(dim=1)
            x = torch.nn.functional.leaky_relu(pre, 0.01)
    return ~near.view(-1, 8).any(dim=1)


def test_aggregate_tensor_core_backward(golden_tiny):
    """Tensor-core training path (split-fp16 forward with tape + dgrad/wgrad kernels) against the oracle's autograd:
    1e-4 on every gradient for upstream gradients supported on kink-free samples."""
    ops, orc, o, xyz, gi, smp, model, c = _agg_setup(golden_tiny, True)
    rgb, alpha, _, _, _ = o
    ok = _kink_free_sample_mask(orc, xyz, gi, smp)
    assert ok.float().mean() > 0.5
    gen = torch.Generator().manual_seed(2)
    ca = torch.randn(alpha.shape, generator=gen) * ok
    cr = torch.randn(rgb.shape, generator=gen) * ok[:, None]
    ((alpha * ca).sum() + (rgb * cr).sum()).backward()
    kx = xyz.detach().cuda().requires_grad_(True)
    kg = gi.detach()[:, :3, :3].reshape(-1, 9).contiguous().cuda().requires_grad_(True)
    model.zero_grad()
    k_alpha, k_rgb, *_ = ops.aggregate_tc_train(c, kx, kg, model.canonical_feat, None, model._mlp_weights(), ops.PackedDecoder())
    assert rel_err(k_alpha, alpha) < RTOL and rel_err(k_rgb, rgb) < RTOL
    ((k_alpha * ca.cuda()).sum() + (k_rgb * cr.cuda()).sum()).backward()
    errs = {"d_xyz": rel_err(kx.grad, xyz.grad), "d_ginv": rel_err(kg.grad.view(-1, 3, 3), gi.grad[:, :3, :3])}
    named = dict(model.named_parameters())
    for k in ["canonical_feat", "feat_net.0.weight", "feat_net.0.bias", "feat_net.2.0.weight", "feat_net.2.0.bias",
              "feat_net.3.0.weight", "feat_net.3.0.bias", "feat_net.4.weight", "feat_net.4.bias", "densitynet.weight",
              "densitynet.bias", "rgbnet.feature_linears.weight", "rgbnet.feature_linears.bias",
              "rgbnet.views_linears.0.weight", "rgbnet.views_linears.0.bias", "rgbnet.views_linears.2.weight",
              "rgbnet.views_linears.2.bias"]:
        assert named[k].grad is not None, k
        errs[k] = rel_err(named[k].grad, orc.s[k].grad)
    bad = {k: v for k, v in errs.items() if not v < RTOL}
    assert not bad, (bad, errs)


@pytest.mark.parametrize("M", [1, 17, 2500])
def test_aggregate_tensor_core_backward_matches_fp32_kernels_on_ragged_sizes(M):
    """Tile tails and multi-tile accumulation (CTAs that walk several tiles; slab reduction of the wgrad) of the
    tensor-core backward against the fp32 CUDA-core backward on random weights: every entry within 5e-4 of the tensor's
    scale (split-fp16 carries 22 mantissa bits; the PE backward multiplies its rounding by up to 2^9).
    The weights are seeded so that no LeakyReLU kink flip separates the two forwards (measured while
    the kernel was brought up: a flip in a dominant row moves that step's layer-0 gradients by ~1e-2, in either implementation)."""
    ops = _ops()
    torch.manual_seed(1)
    g = torch.Generator().manual_seed(100 + M)
    N, d = 2000, "cuda"
    xyz = torch.rand(N, 3, generator=g)
    A = torch.eye(3) + 0.2 * torch.randn(N, 3, 3, generator=g)
    feat = torch.relu(torch.randn(N, 128, generator=g)) * 0.5
    nn_idx = torch.randint(0, N, (M, 8), generator=g).int()
    pts = xyz[nn_idx[:, 0].long()] + 0.02 * torch.randn(M, 3, generator=g)
    ray_id = torch.sort(torch.randint(0, 50, (M,), generator=g))[0].int()
    vd = torch.nn.functional.normalize(torch.randn(50, 3, generator=g), dim=-1)
    lin = [torch.nn.Linear(191, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 128),
           torch.nn.Linear(128, 1), torch.nn.Linear(128, 128), torch.nn.Linear(155, 64), torch.nn.Linear(64, 3)]
    c = ops.AggConst(pts=pts.to(d), nn_idx=nn_idx.to(d), ray_id=ray_id.to(d), viewdirs=vd.to(d),
                     canonical_alpha=torch.rand(N, generator=g).to(d), canonical_rgbs=torch.rand(N, 3, generator=g).to(d),
                     direct_eps=torch.full((N,), 0.05).to(d), mean_min_distance=0.02, eps=1e-6, act_shift=0.0, interval=0.5)
    ca, cr = torch.randn(M, generator=g).to(d), torch.randn(M, 3, generator=g).to(d)

    def run(tc):
        leaves = [xyz.to(d).requires_grad_(True), A.reshape(N, 9).contiguous().to(d).requires_grad_(True),
                  feat.to(d).requires_grad_(True)]
        ws = []
        for l in lin:
            ws += [l.weight.detach().to(d).requires_grad_(True), l.bias.detach().to(d).requires_grad_(True)]
        if tc:
            out = ops.aggregate_tc_train(c, *leaves, None, ws, ops.PackedDecoder())
        else:
            out = ops.aggregate(c, *leaves, None, ws)
        ((out[0] * ca).sum() + (out[1] * cr).sum()).backward()
        return [t.grad for t in leaves + ws]

    names = ["xyz", "ginv", "feat"] + [f"w{i}" for i in range(16)]
    report = {}
    for name, a, b in zip(names, run(True), run(False)):
        err = ((a - b).abs() / (b.abs().max() + 1e-30))[b != 0]
        if err.numel():
            report[name] = (float(err.max()), float((err > 5e-4).float().mean()))
    bad = {k: v for k, v in report.items() if v[0] > 2e-3 or v[1] > 1e-3}
    assert not bad, (bad, report)


# ----------------------------------------------------------------------------------------
# K0 pose chain (one launch) against the PyTorch ops that restate lib/pointwarper.py:217-236
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("masks", [False, True])
def test_fused_pose_chain_matches_pytorch_chain(golden_tiny, masks):
    from articulated_point_nerf_b200 import poc_fre
    g = golden_tiny
    model, scene = model_from_golden(g)
    fw = model.forward_warp
    J = len(model.joints)
    if masks:
        m = torch.zeros(J, dtype=torch.bool)
        m[[3, 11]] = True
        fw.set_rotation_mask(~m.cuda())
        sib = torch.arange(J)
        sib[6], sib[14] = 5, 13
        fw.set_sibling_mask(sib.cuda())
    t_embed = poc_fre(g["train"]["t"].cuda(), model.time_poc)
    gen = torch.Generator().manual_seed(0)
    c1, c2, c3 = torch.randn(J, 4, 4, generator=gen).cuda(), torch.randn(3, generator=gen).cuda(), torch.randn(J, generator=gen).cuda()
    c1[:, 3] = 0          # the last row of a bone transform is constant

    def run(fused):
        fw.fused_pose = fused
        model.zero_grad(set_to_none=True)
        bone_Ts, global_t = fw.pose(model.joints, t=t_embed)
        ((bone_Ts * c1).sum() + (global_t * c2).sum() + (fw.prev_thetas * c3).sum()).backward()
        grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        return bone_Ts.detach(), global_t.detach(), fw.prev_thetas.detach(), grads

    ref, got = run(False), run(True)
    assert fw._fused_tables(model.joints.device) is not None
    for a, b in zip(got[:3], ref[:3]):
        assert rel_err(a, b) < 1e-5
    assert set(got[3]) == set(ref[3]) and "joints" in got[3]
    for k in ref[3]:
        assert rel_err(got[3][k], ref[3][k]) < RTOL, k


@pytest.mark.parametrize("tail", [1, 2, 3])
def test_composite_backward_staged_spans_stay_inside_their_arrays(tail):
    """The backward moves each warp's 32-ray span through shared memory with 16-byte bulk copies and 16-byte stores:
    long rays (several passes per span), a sample count that is not a multiple of four, spans that start and end off
    the 16-byte grid, and outputs embedded in canary-filled buffers (nothing outside [0, M) may be written)."""
    from articulated_point_nerf_b200 import _lib
    lib, P, S = _lib.load(), _lib.ptr, _lib.stream
    g = torch.Generator().manual_seed(tail)
    lens = torch.randint(0, 30, (300,), generator=g)
    lens[5], lens[130], lens[131] = 2100, 700, 0                  # > 4 passes, > 1 pass
    lens[-1] += (tail - int(lens.sum()) % 4) % 4                   # M % 4 == tail
    R, M = len(lens), int(lens.sum())
    assert M % 4 == tail
    d = "cuda"
    ray_start = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(lens, 0)]).int().to(d)
    alpha = (torch.rand(M, generator=g) ** 3).to(d)
    rgb = torch.rand(M, 3, generator=g).to(d)
    step = torch.randint(0, 200, (M,), generator=g).int().to(d)
    rgb_m, last, depth = torch.empty(R, 3, device=d), torch.empty(R, device=d), torch.empty(R, device=d)
    T, used = torch.empty(M, device=d), torch.empty(R, dtype=torch.int32, device=d)
    _lib.check(lib.apn_composite_fwd(P(alpha), P(rgb), P(step), None, 0, P(ray_start), R, 1e-4, 1.0, P(rgb_m), P(last), P(depth),
                                     None, P(T), P(used), S()), "fwd")
    d_rgb_m, d_last, d_depth = [torch.randn(*s, generator=g).to(d) for s in ((R, 3), (R,), (R,))]
    pad = 64
    buf_a = torch.full((M + 2 * pad,), 777.0, device=d)
    buf_c = torch.full((3 * M + 2 * pad,), 777.0, device=d)
    _lib.check(lib.apn_composite_bwd(P(alpha), P(rgb), P(step), P(ray_start), R, 1e-4, 1.0, P(T), P(used), P(last), P(d_rgb_m),
                                     P(d_last), P(d_depth), P(buf_a[pad:]), P(buf_c[pad:]), S()), "bwd")
    torch.cuda.synchronize()
    for buf, n in ((buf_a, M), (buf_c, 3 * M)):
        assert bool((buf[:pad] == 777.0).all()) and bool((buf[pad + n:] == 777.0).all())
        assert bool((buf[pad:pad + n] != 777.0).all())            # every element written
    # against autograd through the oracle's compositing on the CPU
    from oracle.path_oracle import OraclePath
    orc = OraclePath.__new__(OraclePath)
    orc.thres = 1e-4
    a, c = alpha.cpu().requires_grad_(True), rgb.cpu().requires_grad_(True)
    ray_id = torch.repeat_interleave(torch.arange(R), lens)
    o_rgb, o_last, o_depth, _, _, _ = orc.composite(a, c, ray_id, step.cpu().float(), R, 1.0)
    ((o_rgb * d_rgb_m.cpu()).sum() + (o_last * d_last.cpu()).sum() + (o_depth * d_depth.cpu()).sum()).backward()
    assert torch.equal(last.cpu(), o_last.detach())
    assert rel_err(buf_a[pad:pad + M], a.grad) < RTOL
    assert rel_err(buf_c[pad:pad + 3 * M].view(M, 3), c.grad) < RTOL


def test_time_embed_and_render_loss_match_torch():
    """One-launch helpers of the fused training step against the torch expressions they replace
    (lib/tineuvox.py:872-878 on the scalar time; run.py:617-621)."""
    ops = _ops()
    from articulated_point_nerf_b200.heads import poc_fre
    freqs = torch.tensor([2.0 ** i for i in range(8)], device="cuda")
    for tv in (0.0, 0.37, 1.0):
        t = torch.tensor([tv], device="cuda")
        assert torch.equal(ops.time_embed(t, freqs), poc_fre(t, freqs))
    g = torch.Generator().manual_seed(0)
    for n in (1, 777, 8192):
        pred = torch.rand(n, 3, generator=g).cuda().requires_grad_(True)
        target = torch.rand(n, 3, generator=g).cuda()
        ref = 200.0 * torch.nn.functional.mse_loss(pred, target)
        ref.backward()
        loss, grad = ops.mse_loss_grad(pred.detach(), target, 200.0)
        assert abs(loss.item() - ref.item()) < 1e-5 * ref.item()
        assert rel_err(grad, pred.grad) < 1e-6


@pytest.mark.parametrize("N,K,J", [(700, 8, 21), (1500, 8, 65), (257, 5, 3)])
def test_regulariser_kernels_match_torch_autograd(N, K, J):
    """apn_point_regularisers / apn_pose_regularisers: losses and gradients against the torch expressions of
    lib/temporalpoints.py:714-733,797-800 differentiated by autograd (J > 32 exercises the column loop)."""
    from articulated_point_nerf_b200 import ops
    gen = torch.Generator().manual_seed(N + J)
    xyz0 = torch.rand(N, 3, generator=gen).cuda()
    nn_i = torch.randint(0, N, (N, K), generator=gen).cuda()
    nn_i[:, 0] = torch.arange(N).cuda()                                   # a point is its own first neighbour
    eps = 1e-6
    nn_dist = torch.sqrt(((xyz0[:, None] - xyz0[nn_i]) ** 2).sum(-1) + eps)
    xyz = (xyz0 + 0.02 * torch.randn(N, 3, generator=gen).cuda()).requires_grad_(True)
    w = torch.softmax(3 * torch.randn(N, J, generator=gen).cuda(), dim=-1).requires_grad_(True)
    wa, wt, ws = 5e-3, 10.0, 0.2
    arap = wa * (nn_dist - torch.sqrt((xyz[:, None] - xyz[nn_i]).pow(2).sum(-1) + eps)).abs().sum()
    tv = wt * (w[:, None, :] - w[nn_i, :]).abs().mean()
    sp = ws * -(w * torch.log(w + eps) + (1 - w) * torch.log(1 - w + eps)).mean()
    (arap + tv + sp).backward()
    d_xyz = torch.full((N, 3), 0.5, device="cuda")                        # accumulated into
    losses, d_w = ops.point_regularisers(xyz.detach(), w.detach(), nn_i.int().contiguous(), nn_dist, eps, wa, wt, ws, d_xyz)
    for got, ref in zip(losses.tolist(), (arap, tv, sp)):
        assert abs(got - float(ref)) < 1e-4 * abs(float(ref))
    assert rel_err(d_xyz - 0.5, xyz.grad) < 1e-4
    assert rel_err(d_w, w.grad) < 1e-4
    # single terms: a zero weight switches a term off
    d2 = torch.zeros(N, 3, device="cuda")
    l2, dw2 = ops.point_regularisers(xyz.detach(), w.detach(), nn_i.int().contiguous(), nn_dist, eps, wa, 0.0, 0.0, d2)
    assert dw2 is None and l2[1] == 0 and l2[2] == 0 and rel_err(d2, xyz.grad) < 1e-4
    # pose side
    S = 300
    skel = torch.rand(S, 3, generator=gen).cuda()
    joints = torch.rand(J, 3, generator=gen).cuda().requires_grad_(True)
    thetas = torch.randn(J, generator=gen).cuda().requires_grad_(True)
    thetas.data[0] = 0.0
    gt = torch.randn(3, generator=gen).cuda().requires_grad_(True)
    treg = 0.1 * (gt.abs().sum() + thetas.abs().sum()) / J
    d = ((joints[:, None] - skel[None]) ** 2).sum(-1)
    jc = 1.5 * d.min(dim=1)[0].sum()
    (treg + jc).backward()
    lp, d_th, d_gt, d_j = ops.pose_regularisers(thetas.detach(), gt.detach(), joints.detach(), skel, 0.1, 1.5)
    assert abs(float(lp[0]) - float(treg)) < 1e-5 * float(treg) and abs(float(lp[1]) - float(jc)) < 1e-5 * float(jc)
    assert rel_err(d_th, thetas.grad) < 1e-6 and rel_err(d_gt, gt.grad) < 1e-6 and rel_err(d_j, joints.grad) < 1e-5
    lp, d_th, d_gt, d_j = ops.pose_regularisers(thetas.detach(), gt.detach(), joints.detach(), None, 0.1, 0.0)
    assert d_j is None and float(lp[1]) == 0.0 and rel_err(d_th, thetas.grad) < 1e-6
