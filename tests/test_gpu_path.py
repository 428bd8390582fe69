"""GPU parity of the whole drop-in path: TemporalPoints.forward (render / repose / train + backward)
against the committed golden tensors that the reference's own Python produced (tests/golden/ref_tiny.pt,
oracle/make_golden.py) and against the CPU oracle on a second seeded scene."""
import pytest
import torch
import torch.nn.functional as F

from conftest import RTOL, model_from_golden, oracle_for_scene, rel_err

pytestmark = pytest.mark.gpu


def _rk(scene, g=None, rays=None, device="cuda"):
    rk = scene.render_kwargs()
    if g is not None:
        rays = (g["rays_o"], g["rays_d"], g["viewdirs"])
    rk.update(rays_o=rays[0].to(device), rays_d=rays[1].to(device), viewdirs=rays[2].to(device))
    return rk


def test_render_matches_reference_golden(golden_tiny):
    g = golden_tiny
    model, scene = model_from_golden(g)
    rk = _rk(scene, g)
    with torch.no_grad():
        out = model(g["render"]["t"].cuda(), render_depth=True, render_kwargs=rk, render_weights=True,
                    poses=scene.poses[0][None].cuda(), Ks=scene.Ks[0][None].cuda(), get_skeleton=True)
    ref = g["render"]["out"]
    for k in ["t_hat_pcd", "rgb_marched", "alphainv_last", "depth", "rgb_marched_direct", "alphainv_last_direct", "weights",
              "joints"]:
        assert rel_err(out[k], ref[k]) < RTOL, k
    assert out["bones"] == ref["bones"]
    assert model.last_counts["M"] == len(g["render"]["agg"]["ray_id"])
    assert rel_err(model._last_weights, g["render"]["last_weights"]) < 1e-5
    assert rel_err(model.forward_warp.prev_thetas, g["render"]["prev_thetas"]) < 1e-5
    # neighbourhood tables built at first use equal the reference's KeOps argKmin
    assert torch.equal(model.nn_i.cpu(), g["nn_i"])
    assert rel_err(model.mean_min_distance, g["mean_min_distance"]) < 1e-6


def test_repose_matches_reference_golden(golden_tiny):
    g = golden_tiny
    model, scene = model_from_golden(g)
    rk = _rk(scene, g)
    with torch.no_grad():
        out = model(None, render_depth=True, render_kwargs=rk, render_weights=True, rot_params=g["repose"]["rot_params"].cuda(),
                    calc_min_max=True, get_skeleton=True, poses=scene.poses[0][None].cuda(), Ks=scene.Ks[0][None].cuda())
    ref = g["repose"]["out"]
    for k in ["t_hat_pcd", "rgb_marched", "alphainv_last", "depth", "rgb_marched_direct", "weights"]:
        assert rel_err(out[k], ref[k]) < RTOL, k


def test_train_step_gradients_match_reference_golden(golden_tiny):
    g = golden_tiny
    model, scene = model_from_golden(g)
    rk = _rk(scene, g)
    model.zero_grad(set_to_none=True)
    res = model(g["train"]["t"].cuda(), False, rk, render_pcd_direct=False, poses=scene.poses.cuda(), Ks=scene.Ks.cuda())
    loss = F.mse_loss(res["rgb_marched"], g["train"]["target"].cuda()) * 200.0
    loss.backward()
    assert abs(loss.item() - g["train"]["loss"].item()) < RTOL * g["train"]["loss"].item()
    assert rel_err(res["rgb_marched"], g["train"]["rgb_marched"]) < RTOL
    named = dict(model.named_parameters())
    for k, ref in g["train"]["grads"].items():
        assert named[k].grad is not None, k
        assert rel_err(named[k].grad, ref) < RTOL, k


def test_regulariser_losses_match_reference_golden(golden_tiny):
    g = golden_tiny
    model, scene = model_from_golden(g)
    rk = _rk(scene, g)
    with torch.no_grad():
        res = model(g["train"]["t"].cuda(), False, rk)
        got = {"arap": model.get_arap_loss(res["t_hat_pcd"]), "weight_tv": model.get_neighbour_weight_tv_loss(),
               "sparsity": model.get_weight_sparsity_loss(), "transformation_reg": model.get_transformation_regularisation_loss(),
               "joint_chamfer": model.get_joint_chamfer_loss()}
    for k, v in got.items():
        ref = g["losses"][k]
        assert abs(float(v) - float(ref)) <= 1e-4 * max(abs(float(ref)), 1e-3), (k, float(v), float(ref))


def test_empty_batch_returns_background(golden_tiny):
    g = golden_tiny
    model, scene = model_from_golden(g)
    rk = _rk(scene, rays=(g["rays_o"][:7], (-g["rays_d"][:7]).contiguous(), (-g["viewdirs"][:7]).contiguous()))
    with torch.no_grad():
        out = model(g["render"]["t"].cuda(), render_depth=True, render_kwargs=rk)
    assert out["alphainv_last"] is None
    assert torch.equal(out["rgb_marched"].cpu(), torch.ones(7, 3) * scene.cfg.bg)
    assert torch.equal(out["depth"].cpu(), torch.zeros(7))


def test_second_scene_against_oracle_with_merged_weights():
    """A seeded scene the golden file does not cover: other view, merged skinning columns, frozen rotations."""
    from articulated_point_nerf_b200.scene import make_scene, build_model
    scene = make_scene("small")
    model = build_model(scene, seed=3)
    J = len(scene.joints)
    rules = torch.arange(J)
    rules[7], rules[8] = 6, 6
    model.flat_merging_rules = rules
    mask = torch.zeros(J, dtype=torch.bool)
    mask[[7, 8]] = True
    model.forward_warp.rot_mask = mask
    orc = oracle_for_scene(scene, model)
    ro, rd, vd = [x.reshape(-1, 3).contiguous() for x in scene.rays(2)]
    t = torch.tensor([0.81])
    with torch.no_grad():
        ref = orc.forward(t, rays_o=ro, rays_d=rd, viewdirs=vd, near=scene.cfg.near, far=scene.cfg.far,
                          stepsize=scene.cfg.stepsize, bg=scene.cfg.bg)
    model = model.cuda()
    rk = _rk(scene, rays=(ro, rd, vd))
    with torch.no_grad():
        out = model(t.cuda(), render_depth=True, render_kwargs=rk)
    assert model.last_counts["M"] == len(orc.trace["pts"])
    for k in ["t_hat_pcd", "rgb_marched", "alphainv_last", "depth", "rgb_marched_direct", "alphainv_last_direct"]:
        assert rel_err(out[k], ref[k]) < RTOL, k


def test_pointwarper_forward_contract(golden_tiny):
    """PointWarper.forward(weights, joints, t) -> [xyz, joints_rel, G] like lib/pointwarper.py:213-278."""
    g = golden_tiny
    model, scene = model_from_golden(g)
    from articulated_point_nerf_b200 import poc_fre
    with torch.no_grad():
        w = model.get_weights()
        t_embed = poc_fre(g["render"]["t"].cuda(), model.time_poc)
        xyz, joints_rel, G, joints_w, bones = model.forward_warp(w, model.joints, t_embed, get_frames=True, get_skeleton=True)
    assert rel_err(xyz, g["render"]["out"]["t_hat_pcd"]) < 1e-5
    assert G.shape == (len(xyz), 4, 4) and xyz.is_contiguous()
    assert torch.equal(G[:, 3].cpu(), torch.tensor([0., 0., 0., 1.]).expand(len(xyz), 4))
