import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import model_from_golden, rel_err
g = torch.load(os.path.join(ROOT, "tests/golden/ref_tiny.pt"), weights_only=False)
model, scene = model_from_golden(g)
rk = scene.render_kwargs(); rk.update(rays_o=g["rays_o"].cuda(), rays_d=g["rays_d"].cuda(), viewdirs=g["viewdirs"].cuda())
outs = {}
for fused in (False, True, True):
    model.forward_warp.fused_pose = fused
    with torch.no_grad():
        w = model.warp(g["render"]["t"].cuda())
        out = model(g["render"]["t"].cuda(), render_depth=True, render_kwargs=rk)
    print("fused", fused, "M", model.last_counts, "xyz vs golden", rel_err(out["t_hat_pcd"], g["render"]["out"]["t_hat_pcd"]),
          "rgb vs golden", rel_err(out["rgb_marched"], g["render"]["out"]["rgb_marched"]),
          "bone_T max", w["bone_Ts"].abs().max().item(), "global_t", w["global_t"].tolist() if w["global_t"] is not None else None)
    outs[fused] = (w["bone_Ts"].clone(), w["xyz"].clone(), out["rgb_marched"].clone())
print("bone_T fused vs torch", (outs[True][0] - outs[False][0]).abs().max().item(), "xyz", (outs[True][1] - outs[False][1]).abs().max().item())
# ---- is the sampler / k-NN result exact for the fused-pose cloud?
from oracle.path_oracle import OraclePath
from articulated_point_nerf_b200 import ops
orc = OraclePath.__new__(OraclePath); orc.K, orc.voxel_size = 8, scene.voxel_size
for fused in (False, True):
    xyz = outs[fused][1]
    ref = orc.sample_and_knn(xyz.cpu(), g["rays_o"], g["rays_d"], scene.cfg.near, scene.cfg.far, scene.cfg.stepsize, 0.01)
    bbox = torch.cat([xyz.min(0)[0], xyz.max(0)[0]])
    grid = ops.Grid(xyz, bbox, 0.01, 0.01, 0.02)
    smp = ops.sample_and_knn(grid, g["rays_o"].cuda(), g["rays_d"].cuda(), scene.cfg.near, scene.cfg.far, scene.cfg.stepsize * scene.voxel_size)
    print("fused", fused, "oracle M", len(ref["pts"]), "kernel M", smp.M, "S", ref["S"], "bbox", bbox.tolist())
