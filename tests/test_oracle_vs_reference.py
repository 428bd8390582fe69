"""CPU: pins oracle/path_oracle.py + oracle/dvgo_ops.py against tests/golden/ref_tiny.pt and ref_tiny_pose.pt, which
oracle/make_golden.py produced by running the reference's own Python under shims."""
import torch
import torch.nn.functional as F

from oracle import dvgo_ops

RTOL = 1e-4  # BASELINE.json north_star: rgb/alpha and gradients within 1e-4 relative in fp32


def _rel(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def _call(orc, cfg, g, t=None, rot_params=None, **kw):
    return orc.forward(t, rot_params, rays_o=g["rays_o"], rays_d=g["rays_d"], viewdirs=g["viewdirs"], near=cfg.near,
                       far=cfg.far, stepsize=cfg.stepsize, bg=cfg.bg, **kw)


def test_render_outputs_match_reference(golden_any, oracle_any):
    orc, cfg = oracle_any
    g = golden_any
    with torch.no_grad():
        out = _call(orc, cfg, g, t=g["render"]["t"], render_weights=True)
    ref = g["render"]["out"]
    for k in ["t_hat_pcd", "rgb_marched", "alphainv_last", "depth", "rgb_marched_direct", "alphainv_last_direct",
              "weights"]:
        assert _rel(out[k], ref[k]) < RTOL, k
    agg = g["render"]["agg"]
    assert torch.equal(orc.trace["ray_id"], agg["ray_id"])
    assert torch.equal(orc.trace["step_id"], agg["step_id"])
    assert torch.equal(orc.trace["pts"], agg["ray_pts"])
    assert _rel(orc.trace["alpha"], agg["alpha"]) < RTOL
    assert _rel(orc.trace["rgb"], agg["rgbs"]) < RTOL
    assert _rel(orc.trace["alpha_direct"], agg["alpha_direct"]) < RTOL
    assert _rel(orc.trace["weights"], g["render"]["last_weights"]) < RTOL


def test_repose_outputs_match_reference(golden_any, oracle_any):
    orc, cfg = oracle_any
    g = golden_any
    ref = g["repose"]["out"]
    with torch.no_grad():
        out = _call(orc, cfg, g, rot_params=g["repose"]["rot_params"], render_weights=True)
        # the warp itself: last-bit agreement (the oracle sums the blend in another order than the reference's bmm)
        assert _rel(out["t_hat_pcd"], ref["t_hat_pcd"]) < 2e-6
        # everything behind the warp on the reference's own cloud: the sampler is discontinuous in the last bit of
        # the cloud bbox, so a 1-ulp difference of min/max(xyz') may move a handful of samples across the bbox faces
        out = _call(orc, cfg, g, rot_params=g["repose"]["rot_params"], render_weights=True, cloud=ref["t_hat_pcd"])
    for k in ["rgb_marched", "alphainv_last", "depth", "rgb_marched_direct", "weights"]:
        assert _rel(out[k], ref[k]) < RTOL, k


def test_train_gradients_match_reference(golden_any):
    from conftest import oracle_from_golden
    g = golden_any
    orc, cfg = oracle_from_golden(g)
    for v in orc.s.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    out = _call(orc, cfg, g, t=g["train"]["t"])
    loss = F.mse_loss(out["rgb_marched"], g["train"]["target"]) * 200.0
    loss.backward()
    assert abs(loss.item() - g["train"]["loss"].item()) < 1e-4 * g["train"]["loss"].item()
    for k, ref in g["train"]["grads"].items():
        assert orc.s[k].grad is not None, k
        assert _rel(orc.s[k].grad, ref) < RTOL, k


def test_adam_restatement_matches_reference_optimizer(golden_tiny):
    a = golden_tiny["adam"]
    for i, masked in enumerate([False, True]):
        p = a["before"][i].clone()
        gr = a["grads"][i]
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for step in range(1, a["steps"] + 1):
            fn = dvgo_ops.masked_adam_upd if masked else dvgo_ops.adam_upd
            fn(p, gr, m, v, step, 0.9, 0.99, a["lrs"][i], 1e-8)
        assert torch.equal(p, a["after"][i])
        assert torch.equal(m, a["exp_avg"][i])
        assert torch.equal(v, a["exp_avg_sq"][i])


def test_empty_ray_batch_falls_back_to_background(oracle_any, golden_any):
    orc, cfg = oracle_any
    g = golden_any
    ro = g["rays_o"][:7]
    rd = -g["rays_d"][:7]            # looking away from the cloud: no sample in the bbox survives
    with torch.no_grad():
        out = orc.forward(g["render"]["t"], rays_o=ro, rays_d=rd, viewdirs=-g["viewdirs"][:7], near=cfg.near,
                          far=cfg.far, stepsize=cfg.stepsize, bg=cfg.bg)
    assert out["alphainv_last"] is None
    assert torch.equal(out["rgb_marched"], torch.ones(7, 3) * cfg.bg)


def test_regulariser_losses_match_reference(golden_any, oracle_any):
    from articulated_point_nerf_b200.scene import make_scene
    orc, cfg = oracle_any
    g = golden_any
    scene = make_scene(g["config"])
    with torch.no_grad():
        out = _call(orc, cfg, g, t=g["train"]["t"])
        got = {"arap": orc.arap_loss(out["t_hat_pcd"]), "weight_tv": orc.weight_tv_loss(orc.trace["weights"]),
               "sparsity": orc.sparsity_loss(orc.trace["weights"]),
               "transformation_reg": orc.transformation_reg_loss(out["global_t"], out["thetas"]),
               "joint_chamfer": orc.joint_chamfer_loss(scene.skeleton_pcd)}
    assert torch.equal(orc.nn_i, g["nn_i"])
    for k, v in got.items():
        ref = float(g["losses"][k])
        assert abs(float(v) - ref) <= 1e-4 * max(abs(ref), 1e-3), (k, float(v), ref)


def test_batch_chamfer_loss_matches_reference():
    """lib/temporalpoints.py:765-795 (2-D / 3-D, integer pixel grids with exact ties) against the reference's own run."""
    import os
    from conftest import GOLDEN_DIR
    from oracle.path_oracle import OraclePath
    sk = torch.load(os.path.join(GOLDEN_DIR, "ref_skeleton.pt"), weights_only=False)
    for c in sk["batch_chamfer"]:
        p1 = c["pcd1"].clone().requires_grad_(True)
        loss = OraclePath.batch_chamfer_loss(p1, c["pcd2"])
        loss.backward()
        assert _rel(loss.detach(), c["loss"]) < 1e-6
        assert _rel(p1.grad, c["grad1"]) < 1e-6


def test_frozen_view_dir_matches_reference(golden_frozen_view):
    """`frozen_view_dir` (run.py:480-481, lib/temporalpoints.py:157-159,507-508): the reference's own run with one global
    view direction; render outputs, loss and the gradients behind the RGB head."""
    from conftest import oracle_from_golden
    g, v = golden_frozen_view
    orc, cfg = oracle_from_golden(g, frozen_view_dir=True)
    with torch.no_grad():
        out = _call(orc, cfg, g, t=v["t"])
    for k in ["rgb_marched", "alphainv_last", "depth", "rgb_marched_direct", "t_hat_pcd"]:
        assert _rel(out[k], v["render"][k]) < RTOL, k
    assert _rel(out["rgb_marched"], g["render"]["out"]["rgb_marched"]) > 1e-3      # the variant does change the image
    for k, p in orc.s.items():
        if p.is_floating_point() and k != "viewdirs_emb":
            p.requires_grad_(True)
    out = _call(orc, cfg, g, t=v["t"])
    loss = F.mse_loss(out["rgb_marched"], v["target"]) * 200.0
    loss.backward()
    assert abs(loss.item() - v["train"]["loss"].item()) < 1e-4 * v["train"]["loss"].item()
    for k, ref in v["train"]["grads"].items():
        assert _rel(orc.s[k].grad, ref) < RTOL, k


def test_no_view_dir_head_is_the_view_head_with_zero_view_columns(golden_tiny):
    """`no_view_dir=True` has no reference behaviour on this path (lib/temporalpoints.py:504-514 raises UnboundLocalError,
    oracle/make_golden_viewdir.py); the evident intent — rgbnet(h) without view columns — equals the view head with its
    view columns zeroed, which is how the kernels serve it."""
    from conftest import oracle_from_golden
    g = dict(golden_tiny)
    w = g["state_dict"]["rgbnet.views_linears.0.weight"]
    g["state_dict"] = dict(g["state_dict"], **{"rgbnet.views_linears.0.weight": w[:, :128].clone()})
    o_nv, cfg = oracle_from_golden(g, no_view_dir=True)
    wz = w.clone()
    wz[:, 128:] = 0
    g["state_dict"] = dict(g["state_dict"], **{"rgbnet.views_linears.0.weight": wz})
    o_z, _ = oracle_from_golden(g)
    with torch.no_grad():
        a = _call(o_nv, cfg, g, t=g["render"]["t"])
        b = _call(o_z, cfg, g, t=g["render"]["t"])
    assert _rel(a["rgb_marched"], b["rgb_marched"]) < 1e-6
    assert _rel(a["rgb_marched"], g["render"]["out"]["rgb_marched"]) > 1e-3
