"""Alpha / early-termination statistics of a synthetic scene ('record mean alpha and % rays early-stopped')."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from articulated_point_nerf_b200 import ops
from articulated_point_nerf_b200.scene import build_model, make_scene

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
scene = make_scene(name)
model = build_model(scene, seed=0).cuda()
rk = scene.render_kwargs()
for view in (0, 1, 2):
    ro, rd, vd = [x.reshape(-1, 3).contiguous().cuda() for x in scene.rays(view)]
    t = torch.tensor([view / 7.0], device="cuda")
    with torch.no_grad():
        warped = model.warp(t)
        grid = model.build_grid(warped)
        smp = ops.sample_and_knn(grid, ro, rd, scene.cfg.near, scene.cfg.far, scene.cfg.stepsize * scene.voxel_size)
        c = ops.AggConst(pts=smp.pts, nn_idx=smp.nn_idx, ray_id=smp.ray_id, viewdirs=vd, canonical_alpha=model.canonical_alpha.detach(),
                         canonical_rgbs=model.canonical_rgbs.detach(), direct_eps=model.direct_eps.detach(),
                         mean_min_distance=model._mmd_float, eps=float(model.eps), act_shift=float(model.tineuvox.act_shift),
                         interval=float(rk['stepsize']) * float(model.tineuvox.voxel_size_ratio), direct=True)
        alpha, rgb, *_ = ops.aggregate_tc(c, warped['xyz'], warped['ginv'], model.canonical_feat, None, model._mlp_weights(),
                                          model._packed_decoder, precision=1)
        out = model(t, render_depth=True, render_kwargs=dict(rk, rays_o=ro, rays_d=rd, viewdirs=vd), warped=warped, grid=grid)
    hit = (smp.ray_start[1:] - smp.ray_start[:-1]) > 0
    last = out["alphainv_last"]
    q = torch.quantile(alpha, torch.tensor([0.1, 0.5, 0.9], device="cuda"))
    print(f"{name} view {view}: M={smp.M} rays hit={int(hit.sum())}/{len(ro)} alpha mean={float(alpha.mean()):.3f} "
          f"q10/50/90={[round(float(x), 3) for x in q]} rays early-stopped (T<1e-3)={float((last[hit] < 1e-3).float().mean()):.3f} "
          f"mean T_last(hit)={float(last[hit].mean()):.3f} rgb range=({float(rgb.min()):.2f},{float(rgb.max()):.2f})")
