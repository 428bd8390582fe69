import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from dataclasses import replace
from conftest import load_golden, model_from_golden, rel_err
from articulated_point_nerf_b200.train import regulariser_losses, Regularisers
g, gf = load_golden("tiny"), load_golden("tiny_fullstep")
for graph_pose in (True, False):
    for order in ("regs_first", "arap_first", "arap_only"):
        model, scene = model_from_golden(g, fused_pose=False)
        model.graph_pose = graph_pose
        rk = dict(scene.render_kwargs(), rays_o=g["rays_o"].cuda(), rays_d=g["rays_d"].cuda(), viewdirs=g["viewdirs"].cuda())
        res = model(gf["t"].cuda(), False, rk, render_pcd_direct=False)
        reg = Regularisers()
        ps = [model.joints, model.weights, model.theta_weight]
        def arap():
            gr = torch.autograd.grad(5e-3 * model.get_arap_loss(res["t_hat_pcd"]), ps, retain_graph=True, allow_unused=True)
            return " ".join(f"{k}={rel_err(a, gf['grads_arap'][k]):.3e}" for k, a in zip(("joints", "weights", "theta_weight"), gr))
        def regs():
            gr = torch.autograd.grad(regulariser_losses(model, res["t_hat_pcd"], replace(reg, arap=0.0)), ps, retain_graph=True, allow_unused=True)
            return "regs ok"
        if order == "regs_first":
            regs(); out = arap()
        elif order == "arap_first":
            out = arap(); regs()
        else:
            out = arap()
        print(f"graph_pose={graph_pose} {order}: {out}")
