"""TEST INFRASTRUCTURE ONLY — generate tests/golden/ref_tiny.pt from the reference itself.

Run in the authoring container:  python -m oracle.make_golden
It executes the UNMODIFIED reference Python (lib/temporalpoints.py, lib/pointwarper.py,
lib/tineuvox.py, lib/masked_adam.py imported from /root/reference) on CPU under the shims
of oracle/ref_harness.py on a seeded synthetic scene, and stores inputs, parameters,
outputs, stage tensors and gradients.  tests/ compare both oracle/path_oracle.py and the
CUDA path against this file.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from articulated_point_nerf_b200.scene import make_scene  # noqa: E402
from oracle import ref_harness  # noqa: E402


def run(config="tiny", out_path=None):
    scene = make_scene(config)
    model, tv = ref_harness.build_reference_model(scene)
    tineuvox, temporalpoints, _, masked_adam = ref_harness.import_reference()
    cfg = scene.cfg
    rk = scene.render_kwargs()
    rays_o, rays_d, viewdirs = scene.rays(0)
    rays_o, rays_d, viewdirs = [x.reshape(-1, 3).contiguous() for x in (rays_o, rays_d, viewdirs)]
    rk.update(rays_o=rays_o, rays_d=rays_d, viewdirs=viewdirs)
    g = {"config": config, "n_points": len(scene.canonical_pcd)}
    g["state_dict"] = {k: v.detach().clone() for k, v in model.state_dict().items() if not k.startswith("tineuvox.")}
    g["canonical_pcd"] = scene.canonical_pcd.clone()
    g["bones"] = scene.bones
    g["mean_min_distance"] = model.mean_min_distance.detach().clone()
    g["nn_i"] = model.nn_i.clone()
    g["act_shift"] = float(tv.act_shift)
    g["voxel_size_ratio"] = float(tv.voxel_size_ratio)
    g["voxel_size"] = scene.voxel_size
    g["rays_o"], g["rays_d"], g["viewdirs"] = rays_o, rays_d, viewdirs

    # capture aggregate_pts outputs
    cap = {}
    orig = model.aggregate_pts

    def wrapped(*a, **k):
        r = orig(*a, **k)
        cap["agg"] = r
        return r

    model.aggregate_pts = wrapped

    # ---- render call, as run.py:149-151 ---------------------------------------------
    t = torch.tensor([0.37])
    with torch.no_grad():
        out = model(t, render_depth=True, render_kwargs=rk, render_weights=True, poses=scene.poses[0][None],
                    Ks=scene.Ks[0][None], cam_per_ray=torch.zeros(len(rays_o))[:, None], get_skeleton=True)
    rgbs, alpha, rgbs_direct, alpha_direct, lbs_w, ray_pts, ray_id, step_id, _ = cap["agg"]
    g["render"] = {
        "t": t,
        "out": {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in out.items() if k != "grid"},
        "agg": dict(rgbs=rgbs, alpha=alpha, rgbs_direct=rgbs_direct, alpha_direct=alpha_direct, lbs_w=lbs_w,
                    ray_pts=ray_pts, ray_id=ray_id, step_id=step_id),
        "last_weights": model._last_weights.detach().clone(),
        "prev_thetas": model.forward_warp.prev_thetas.detach().clone(),
        "prev_global_t": model.forward_warp.prev_global_t.detach().clone(),
    }

    # ---- repose call, as run.py:287 -----------------------------------------------------
    gen = torch.Generator().manual_seed(1)
    J = len(scene.joints)
    rot_params = torch.randn(J, 4, generator=gen) * 0.2
    rot_params[0] = 0
    with torch.no_grad():
        out = model(None, render_depth=True, render_kwargs=rk, render_weights=True, rot_params=rot_params,
                    calc_min_max=True, get_skeleton=True, poses=scene.poses[0][None], Ks=scene.Ks[0][None])
    g["repose"] = {"rot_params": rot_params,
                   "out": {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in out.items() if k != "grid"}}

    # ---- train call + backward, as run.py:615-631,713 ---------------------------------
    target = torch.rand(len(rays_o), 3, generator=gen)
    model.zero_grad(set_to_none=True)
    res = model(t, False, rk, render_pcd_direct=False, poses=scene.poses, Ks=scene.Ks,
                cam_per_ray=torch.zeros(len(rays_o), 1, dtype=torch.long))
    loss = torch.nn.functional.mse_loss(res["rgb_marched"], target) * 200.0
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()
             if p.grad is not None and not k.startswith("tineuvox.")}
    g["train"] = {"t": t, "target": target, "loss": loss.detach().clone(),
                  "rgb_marched": res["rgb_marched"].detach().clone(), "grads": grads}

    # ---- regulariser losses on the same forward (run.py:633-657) ------------------------
    with torch.no_grad():
        g["losses"] = {
            "arap": model.get_arap_loss(res["t_hat_pcd"]).clone(),
            "weight_tv": model.get_neighbour_weight_tv_loss().clone(),
            "sparsity": model.get_weight_sparsity_loss().clone(),
            "transformation_reg": model.get_transformation_regularisation_loss().clone(),
            "joint_chamfer": model.get_joint_chamfer_loss().clone(),
        }

    # ---- one MaskedAdam step over three tensors (lib/masked_adam.py:39-72) -----------------
    p_plain = torch.nn.Parameter(model.feat_net[0].weight.detach().clone())
    p_mask = torch.nn.Parameter(model.weights.detach().clone())
    p_plain.grad = grads["feat_net.0.weight"].clone()
    gm = grads["weights"].clone()
    gm[gm.abs() < gm.abs().median()] = 0  # exercise the skip-zero-grad branch
    p_mask.grad = gm
    opt = masked_adam.MaskedAdam([
        {"params": [p_plain], "lr": 1e-3, "skip_zero_grad": False},
        {"params": [p_mask], "lr": 1e-4, "skip_zero_grad": True}])
    before = (p_plain.detach().clone(), p_mask.detach().clone())
    for _ in range(3):
        opt.step()
    g["adam"] = {"before": before, "grads": (p_plain.grad.clone(), gm.clone()),
                 "after": (p_plain.detach().clone(), p_mask.detach().clone()),
                 "exp_avg": (opt.state[p_plain]["exp_avg"].clone(), opt.state[p_mask]["exp_avg"].clone()),
                 "exp_avg_sq": (opt.state[p_plain]["exp_avg_sq"].clone(), opt.state[p_mask]["exp_avg_sq"].clone()),
                 "steps": 3, "lrs": (1e-3, 1e-4)}

    out_path = out_path or os.path.join(ROOT, "tests", "golden", f"ref_{config}.pt")
    torch.save(g, out_path)
    n_kept = len(ray_id)
    print(f"wrote {out_path}: N={len(scene.canonical_pcd)} R={len(rays_o)} kept={n_kept} "
          f"loss={float(loss):.6f} size={os.path.getsize(out_path) / 1e6:.2f} MB")
    return g


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "tiny")
