"""TEST INFRASTRUCTURE ONLY — tests/golden/ref_tiny_fullstep.pt: the reference's COMPLETE stage-2 iteration loss
(run.py:615-694: render loss + ARAP + weight TV + sparsity + transformation regulariser + joint chamfer + 2-D chamfer),
evaluated by the UNMODIFIED reference (lib/temporalpoints.py, lib/utils.py from /root/reference, CPU, under the shims of
oracle/ref_harness.py) on the `tiny` scene of ref_tiny.pt, with the gradients of every parameter.

    python -m oracle.make_golden_fullstep

The model state is the one of ref_tiny.pt (same seeded build), so only inputs of the extra terms, the loss terms and the
gradients are stored.  The 2-D chamfer term uses every projected point (N=None) instead of run.py's random 3000-point
subset so that the result does not depend on an RNG stream; the mask pixels are a seeded synthetic set.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from articulated_point_nerf_b200.scene import make_scene  # noqa: E402
from oracle import ref_harness  # noqa: E402

WEIGHTS = dict(render=2e2, arap=5e-3, tv=1e1, sparsity=2e-1, transformation_reg=1e-1, joint_chamfer=1.0, chamfer2D=5e-3)


def run(config="tiny", out_path=None):
    scene = make_scene(config)
    model, tv = ref_harness.build_reference_model(scene)
    ref_harness.import_reference()
    from lib import utils as ref_utils                      # the reference's own projection (lib/utils.py:435-450)
    rk = scene.render_kwargs()
    rays_o, rays_d, viewdirs = [x.reshape(-1, 3).contiguous() for x in scene.rays(0)]
    rk.update(rays_o=rays_o, rays_d=rays_d, viewdirs=viewdirs)
    gen = torch.Generator().manual_seed(1)
    J = len(scene.joints)
    _ = torch.randn(J, 4, generator=gen)                    # same stream position as make_golden.run -> same target
    target = torch.rand(len(rays_o), 3, generator=gen)
    t = torch.tensor([0.37])
    # synthetic mask pixels for the 2-D chamfer term: B views x M (row, col) coordinates
    B, M = min(3, len(scene.poses)), 257
    H, W = scene.cfg.H, scene.cfg.W
    mask_pcd = torch.stack([torch.randint(0, H, (B, M), generator=gen), torch.randint(0, W, (B, M), generator=gen)], dim=-1).float()
    poses_c, Ks_c = scene.poses[:B].float(), scene.Ks[:B].float()

    model.zero_grad(set_to_none=True)
    res = model(t, False, rk, render_pcd_direct=False, poses=scene.poses, Ks=scene.Ks,
                cam_per_ray=torch.zeros(len(rays_o), 1, dtype=torch.long))
    t_hat_pcd = res["t_hat_pcd"]
    terms = {"render": WEIGHTS["render"] * torch.nn.functional.mse_loss(res["rgb_marched"], target),      # run.py:617-631
             "arap": WEIGHTS["arap"] * model.get_arap_loss(t_hat_pcd),                                      # run.py:633-636
             "tv": WEIGHTS["tv"] * model.get_neighbour_weight_tv_loss(),                                   # run.py:638-641
             "sparsity": WEIGHTS["sparsity"] * model.get_weight_sparsity_loss(),                           # run.py:643-648
             "transformation_reg": WEIGHTS["transformation_reg"] * model.get_transformation_regularisation_loss(),   # :650-653
             "joint_chamfer": WEIGHTS["joint_chamfer"] * model.get_joint_chamfer_loss()}                   # run.py:655-658
    proj = ref_utils.project_point_to_image_plane(t_hat_pcd, poses_c, Ks_c)                                # run.py:675-679
    if not rk["inverse_y"]:
        proj[:, :, 0] = (H - 1) - proj[:, :, 0]
    proj = proj.flip(-1)
    terms["chamfer2D"] = WEIGHTS["chamfer2D"] * model.get_batch_chamfer_loss(proj, mask_pcd, N=None, M=None)   # run.py:688-690
    loss = sum(terms.values())
    # the regulariser part alone (everything but the render loss): unlike the render gradient it does not pass through the
    # 2^9 positional encoding, so it is well conditioned and can be compared at 1e-4 for EVERY parameter
    named = [(k, p) for k, p in model.named_parameters() if p.requires_grad and not k.startswith("tineuvox.")]
    # ARAP is kept apart: |D0 - d| has its kink exactly where a neighbourhood moves rigidly (d == D0 up to rounding, the
    # common case), so its gradient there is sign(rounding noise) and no two implementations of the warp agree on it
    reg_loss = sum(v for k, v in terms.items() if k not in ("render", "arap"))
    gr = torch.autograd.grad(reg_loss, [p for _, p in named], retain_graph=True, allow_unused=True)
    grads_reg = {k: g_.detach().clone() for (k, _), g_ in zip(named, gr) if g_ is not None}
    gr = torch.autograd.grad(terms["arap"], [p for _, p in named], retain_graph=True, allow_unused=True)
    grads_arap = {k: g_.detach().clone() for (k, _), g_ in zip(named, gr) if g_ is not None}
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()
             if p.grad is not None and not k.startswith("tineuvox.")}
    g = {"config": config, "weights": WEIGHTS, "t": t, "target": target, "mask_pcd": mask_pcd, "n_views": B,
         "terms": {k: v.detach().clone() for k, v in terms.items()}, "loss": loss.detach().clone(), "grads": grads, "grads_reg": grads_reg, "grads_arap": grads_arap}
    out_path = out_path or os.path.join(ROOT, "tests", "golden", f"ref_{config}_fullstep.pt")
    torch.save(g, out_path)
    print(f"wrote {out_path}: loss={float(loss):.6f} terms={ {k: round(float(v), 6) for k, v in terms.items()} } "
          f"size={os.path.getsize(out_path) / 1e6:.2f} MB")
    return g


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "tiny")
