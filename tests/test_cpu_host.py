"""CPU: the C-ABI library loads and exports everything include/apn.h declares; host-side logic of the
drop-in modules (tree tables, pose maths, state-dict layout, optimiser bookkeeping) against the oracle /
the reference golden file.  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, model_from_golden, rel_err


def test_library_builds_loads_and_exports_every_declared_symbol():
    from articulated_point_nerf_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "apn.h")).read()
    declared = set(re.findall(r"\b(apn_[a-z0-9_]+)\s*\(", header))
    declared -= {"apn_stream_t"}
    lib = ctypes.CDLL(path)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/apn.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().apn_version() == 100
    assert _lib.load().apn_launch_count() == 0


def test_ctypes_struct_layout_matches_header_sizes():
    """Guards the hand-written ctypes mirrors against drift: sizes computed from the C declarations."""
    from articulated_point_nerf_b200 import _lib
    P = ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.MlpWeights) == 16 * P
    assert ctypes.sizeof(_lib.AggInputs) == 4 * 4 + 11 * P + 4 * 4 + P     # 3 ints (+pad), 11 pointers, 4 floats, m_dev
    assert ctypes.sizeof(_lib.AggOutputs) == 14 * P
    assert ctypes.sizeof(_lib.AggGrads) == 22 * P
    assert ctypes.sizeof(_lib.AdamTensor) == 5 * P + 8 + 4 + 4


def test_cpu_tensors_are_rejected_loudly():
    from articulated_point_nerf_b200 import ops, _lib
    with pytest.raises(_lib.ApnError):
        ops.lbs(torch.rand(4, 3), torch.tensor([0.1]), torch.eye(4).repeat(3, 1, 1), None, torch.rand(4, 3))


def test_state_dict_keys_and_kwargs_follow_the_reference(golden_tiny):
    model, scene = model_from_golden(golden_tiny, device="cpu")
    ours = {k for k in model.state_dict() if not k.startswith("tineuvox.")}
    assert ours == set(golden_tiny["state_dict"])
    kw = model.get_kwargs()
    for k in ["canonical_pcd", "skeleton_pcd", "canonical_alpha", "canonical_feat", "canonical_rgbs", "joints", "bones",
              "neighbours", "timebase_pe", "eps", "stepsize", "weights", "xyz_min", "xyz_max", "tineuvox", "voxel_size",
              "fast_color_thres", "embedding", "frozen_view_dir", "over_parameterized_rot", "feat_depth",
              "pose_embedding_dim"]:
        assert k in kw, k
    # rebuilding from get_kwargs (utils.load_model: model_class(**ckpt['model_kwargs'])) keeps the parameters
    from articulated_point_nerf_b200 import TemporalPoints
    clone = TemporalPoints(**kw)
    clone.load_state_dict(model.state_dict(), strict=False)
    assert torch.equal(clone.weights, model.weights)
    with pytest.raises(AssertionError):
        model(torch.tensor([0.1]), render_kwargs={}, rot_params=torch.zeros(21, 4))     # t XOR rot_params


def test_initial_skinning_weights_match_reference(golden_tiny):
    """_weights_from_bones (lib/temporalpoints.py:235-254) == the reference's initial `weights` parameter."""
    from articulated_point_nerf_b200.scene import make_scene, build_model
    scene = make_scene("tiny")
    model = build_model(scene)
    assert rel_err(model.weights, golden_tiny["state_dict"]["weights"]) < 1e-6


def test_pose_chain_matches_oracle(golden_tiny, oracle_tiny):
    """TransformNet -> Rodrigues -> kinematic chain on CPU == oracle.bone_transforms (lib/pointwarper.py:118-193)."""
    from articulated_point_nerf_b200 import poc_fre
    model, scene = model_from_golden(golden_tiny, device="cpu")
    orc, cfg = oracle_tiny
    t = golden_tiny["render"]["t"]
    with torch.no_grad():
        bone_Ts, global_t = model.forward_warp.pose(model.joints, t=poc_fre(t, model.time_poc))
        ref = orc.warp(t)
    assert rel_err(bone_Ts, ref["bone_Ts"]) < 1e-6
    assert rel_err(global_t, ref["global_t"]) < 1e-6
    assert rel_err(model.forward_warp.prev_thetas, golden_tiny["render"]["prev_thetas"]) < 1e-6
    assert rel_err(model.get_weights(), ref["weights"]) < 1e-6
    # rot_params path: no global translation, frozen + sibling-shared rotations
    rp = golden_tiny["repose"]["rot_params"]
    model.forward_warp.set_rotation_mask(~torch.tensor([i in (3, 4) for i in range(len(rp))]))
    sib = torch.arange(len(rp))
    sib[6] = 5
    model.forward_warp.set_sibling_mask(sib)
    orc2_state = dict(orc.s)
    orc2_state["forward_warp.rot_mask"] = model.forward_warp.rot_mask
    orc2_state["forward_warp.sibling_mask"] = sib
    saved = orc.s
    orc.s = orc2_state
    try:
        with torch.no_grad():
            b2, g2 = model.forward_warp.pose(model.joints, rot_params=rp)
            ref2 = orc.warp(None, rp)
    finally:
        orc.s = saved
    assert g2 is None and rel_err(b2, ref2["bone_Ts"]) < 1e-6


def test_masked_adam_argument_validation_and_grouping():
    from articulated_point_nerf_b200 import MaskedAdam
    p = torch.nn.Parameter(torch.zeros(3))
    with pytest.raises(ValueError):
        MaskedAdam([p], lr=-1.0)
    with pytest.raises(ValueError):
        MaskedAdam([p], betas=(1.0, 0.99))
    opt = MaskedAdam([{"params": [p], "lr": 1e-3, "skip_zero_grad": True}])
    assert opt.param_groups[0]["betas"] == (0.9, 0.99) and opt.param_groups[0]["eps"] == 1e-8
    opt.step()                                   # no grads: nothing to launch, no device needed
    assert len(opt.state) == 0


def test_synthetic_scene_is_deterministic():
    from articulated_point_nerf_b200.scene import make_scene
    a, b = make_scene("tiny"), make_scene("tiny")
    assert torch.equal(a.canonical_pcd, b.canonical_pcd) and torch.equal(a.canonical_feat, b.canonical_feat)
    assert a.bones == b.bones and all(p < c for p, c in a.bones)
    assert [c for _, c in a.bones] == list(range(1, len(a.joints)))      # bone i = [parent, i+1]


def test_reference_gradients_are_ill_conditioned_in_the_warped_cloud(golden_tiny):
    """Evidence for the tolerance split in tests/test_gpu_path.py: in the reference's own arithmetic (the CPU oracle)
    a +-4-ulp change of the warped cloud moves the gradients that pass through PE(2^9 * rel_c) and the LeakyReLU kinks
    of feat_net layer 0 by more than the 1e-4 parity bar, while the heads stay far below it."""
    import torch.nn.functional as F
    from conftest import oracle_from_golden
    g = golden_tiny
    keys = ["canonical_feat", "feat_net.0.bias", "rgbnet.views_linears.0.weight"]

    def grads(ulp):
        orc, cfg = oracle_from_golden(g)
        for k in keys:
            orc.s[k].requires_grad_(True)
        with torch.no_grad():
            wp = orc.warp(g["train"]["t"])
            Ginv = torch.inverse(wp["G"])
        xyz = wp["xyz"].clone()
        if ulp:
            d = torch.randint(-ulp, ulp + 1, xyz.shape, generator=torch.Generator().manual_seed(0)).int()
            xyz = (xyz.view(torch.int32) + d).view(torch.float32)
        xyz.requires_grad_(True)
        smp = orc.sample_and_knn(xyz, g["rays_o"], g["rays_d"], cfg.near, cfg.far, cfg.stepsize, 0.01)
        rgb, alpha, *_ = orc.aggregate(xyz, Ginv, smp, g["viewdirs"], cfg.stepsize)
        rgb_m, *_ = orc.composite(alpha, rgb, smp["ray_id"], smp["step_id"], len(g["rays_o"]), cfg.bg)
        (F.mse_loss(rgb_m, g["train"]["target"]) * 200.0).backward()
        return {"d_xyz": xyz.grad, **{k: orc.s[k].grad for k in keys}}

    a, b = grads(0), grads(4)
    assert rel_err(b["d_xyz"], a["d_xyz"]) > 1e-4
    assert rel_err(b["canonical_feat"], a["canonical_feat"]) > 1e-4
    assert rel_err(b["feat_net.0.bias"], a["feat_net.0.bias"]) > 1e-4
    assert rel_err(b["rgbnet.views_linears.0.weight"], a["rgbnet.views_linears.0.weight"]) < 1e-5


def test_checkpoint_round_trip_keeps_the_reference_layout(tmp_path):
    """temporalpoints_last.tar: {'global_step', 'model_kwargs', 'model_state_dict', 'optimizer_state_dict'} (run.py:1234-1235);
    load = model_class(**model_kwargs) + load_state_dict (lib/utils.py:519-523)."""
    import torch
    from articulated_point_nerf_b200 import load_checkpoint, save_checkpoint
    from articulated_point_nerf_b200.scene import build_model, make_scene
    scene = make_scene("tiny")
    model = build_model(scene, seed=3)
    with torch.no_grad():
        model.theta_weight.fill_(0.123)
        model.flat_merging_rules[2] = 1
    path = str(tmp_path / "temporalpoints_last.tar")
    save_checkpoint(path, model, optimizer=None, global_step=42)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) == {"global_step", "model_kwargs", "model_state_dict", "optimizer_state_dict"} and raw["global_step"] == 42
    for k in ("canonical_pcd", "skeleton_pcd", "bones", "joints", "weights", "xyz_min", "xyz_max", "tineuvox", "stepsize",
              "voxel_size", "fast_color_thres", "feat_depth", "pose_embedding_dim"):
        assert k in raw["model_kwargs"], k
    loaded, ckpt = load_checkpoint(path, device="cpu")
    a, b = model.state_dict(), loaded.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(loaded.canonical_pcd, model.canonical_pcd) and loaded.bones == model.bones


# ---------------------------------------------------------------------------------------------------------------------
# skeleton simplification: lib/treeprune.py merge_joints and lib/temporalpoints.py simplify_skeleton
# against the reference's own outputs (tests/golden/ref_skeleton.pt, oracle/make_golden_skeleton.py)
@pytest.fixture(scope="module")
def golden_skeleton():
    return torch.load(os.path.join(ROOT, "tests", "golden", "ref_skeleton.pt"), weights_only=False)


def test_merge_joints_matches_reference_on_its_fixture_and_random_trees(golden_skeleton):
    import numpy as np
    from articulated_point_nerf_b200.treeprune import merge_joints
    names = ["new_joints", "new_bones", "merging_rules", "joints_to_keep", "rotations_to_keep", "rotation_switch_mask",
             "sibling_transfer_rules"]
    n_ok = n_sibling = 0
    for c in golden_skeleton["merge_joints"]:
        if not c["ok"]:
            with pytest.raises(Exception):         # degenerate trees the reference itself cannot handle
                merge_joints(c["joints"], c["bones"], c["prune"].copy(), c["sim"], convert_merging_rules=c["convert"])
            continue
        out = merge_joints(c["joints"], c["bones"], c["prune"].copy(), c["sim"], convert_merging_rules=c["convert"])
        for name, ours, ref in zip(names, out, c["out"]):
            assert np.array_equal(np.asarray(ours), ref), (c["name"], name)
            assert np.asarray(ours).dtype == ref.dtype, (c["name"], name)
        n_ok += 1
        n_sibling += int((c["out"][6] != np.arange(len(c["out"][6]))).any())
    assert n_ok > 100 and n_sibling > 20           # the sibling-merge branch is exercised


def test_simplify_skeleton_matches_reference(golden_tiny, golden_skeleton):
    import numpy as np
    for s in golden_skeleton["simplify"]:
        model, scene = model_from_golden(golden_tiny, device="cpu")
        joints, bones, new_joints, new_bones, prune_bones, merging_rules, rot_keep, res = model.simplify_skeleton(
            s["times"], deg_threshold=s["deg_threshold"], five_percent_heuristic=s["five_percent"])
        assert torch.equal(prune_bones.cpu(), s["prune_bones"])
        assert np.array_equal(merging_rules, s["merging_rules"])
        assert np.array_equal(new_joints, s["new_joints"]) and np.array_equal(new_bones, s["new_bones"])
        assert torch.equal(rot_keep, s["rotations_to_keep"])
        assert torch.equal(model.flat_merging_rules.long(), s["flat_merging_rules"])
        assert torch.equal(model.sibling_merging_rules.long(), s["sibling_merging_rules"])
        assert torch.equal(model.forward_warp.rot_mask, s["rot_mask"])
        assert torch.equal(model.forward_warp.sibling_mask, s["sibling_mask"])
        # merged skinning weights (the reference's (J,J,J) merging_mat bmm, lib/temporalpoints.py:405-414)
        assert rel_err(model.get_weights(), s["last_weights"]) < 1e-5
        # a checkpoint written after the simplification loads back (load_model path, lib/utils.py:519-523)
        fresh, _ = model_from_golden(golden_tiny, device="cpu")
        fresh.load_state_dict(model.state_dict(), strict=False)
        assert torch.equal(fresh.flat_merging_rules.long(), s["flat_merging_rules"])
        assert torch.equal(fresh.forward_warp.rot_mask, s["rot_mask"])


def test_rotation_angle_of_relative_rotations():
    from articulated_point_nerf_b200 import TemporalPoints
    from articulated_point_nerf_b200.pointwarper import rodrigues
    torch.manual_seed(0)
    axis = torch.randn(64, 3)
    ang = torch.linspace(0.0, 3.1, 64)
    R, _ = rodrigues(torch.cat([axis, ang[:, None]], -1))
    # the 1e-5 inside the axis normalisation (lib/pointwarper.py:127) makes R slightly non-orthogonal: loose bound
    assert (TemporalPoints._rotation_angle(R) - ang).abs().max() < 1e-3


def test_pcds_formats_round_trip_into_the_stage2_model(tmp_path):
    """pcds/canonical.tar + pcds/skeleton.tar in the reference's layout (run.py:1090-1103, 1214-1230) and the stage-2
    model construction from them (run.py:457-503)."""
    import numpy as np
    from articulated_point_nerf_b200 import model_from_pcds, save_pcds
    from articulated_point_nerf_b200.heads import TiNeuVoxHeads
    from articulated_point_nerf_b200.render import CANONICAL_KEYS, SKELETON_KEYS
    from articulated_point_nerf_b200.scene import make_scene
    scene = make_scene("tiny")
    folder = os.path.join(tmp_path, "run", "pcds")
    save_pcds(folder, pcd=scene.canonical_pcd, rgbs=scene.canonical_rgbs, feat=scene.canonical_feat,
              alphas=scene.canonical_alpha, skeleton_pcd=scene.skeleton_pcd, joints=scene.joints, bones=scene.bones,
              xyz_min=scene.xyz_min, xyz_max=scene.xyz_max, voxel_size=scene.voxel_size, t=0.0)
    can = torch.load(os.path.join(folder, "canonical.tar"), weights_only=False)
    skel = torch.load(os.path.join(folder, "skeleton.tar"), weights_only=False)
    assert tuple(can.keys()) == CANONICAL_KEYS and tuple(skel.keys()) == SKELETON_KEYS     # the reference's dict layouts
    assert isinstance(skel["joints"], np.ndarray) and np.array_equal(skel["root"], skel["joints"][0])
    heads = TiNeuVoxHeads(scene.xyz_min.numpy(), scene.xyz_max.numpy(), num_voxels=scene.cfg.num_voxels,
                          num_voxels_base=scene.cfg.num_voxels, alpha_init=1e-3, net_width=128, no_view_dir=False)
    model = model_from_pcds(os.path.join(tmp_path, "run"), heads, world_bound_scale=1.05, stepsize=scene.cfg.stepsize,
                            fast_color_thres=scene.cfg.fast_color_thres)
    assert torch.equal(model.canonical_pcd, scene.canonical_pcd)
    assert torch.equal(model.canonical_feat.detach(), scene.canonical_feat)
    assert torch.equal(model.joints.detach(), scene.joints.float())
    assert model.bones == [list(map(int, b)) for b in scene.bones]
    assert torch.allclose(model.xyz_max, scene.xyz_max.float() * 1.05) and model.voxel_size == scene.voxel_size
    assert model.weights.shape == (len(scene.canonical_pcd), len(scene.joints))          # initial skinning weights from the bones


def test_ctypes_arities_match_the_header_prototypes():
    """Every prototype of include/apn.h has as many parameters as its ctypes signature (catches a binding that drifts
    from the header when an entry point gains an argument)."""
    from articulated_point_nerf_b200 import _lib
    header = open(os.path.join(ROOT, "include", "apn.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    header = re.sub(r"//[^\n]*", " ", header)
    protos = re.findall(r"\b(apn_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", header, flags=re.S)
    seen = {}
    for name, params in protos:
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        seen[name] = n
    assert set(seen) == set(_lib.SIGNATURES)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == seen[name], (name, len(argtypes), seen[name])
    # ... and of the same kind, parameter by parameter: pointer / int / float / 64-bit size
    def kind_c(decl):
        decl = decl.strip()
        if "*" in decl or decl.startswith("apn_stream_t"):
            return "ptr"
        base = decl.rsplit(" ", 1)[0].replace("const", "").strip()
        return {"int": "int", "int32_t": "int", "float": "float", "size_t": "size", "long long": "ll",
                "unsigned long long": "ll"}[base]
    kinds_py = {_lib.I: "int", _lib.F: "float", _lib.P: "ptr", _lib.LL: "ll", _lib.SZ: "size"}
    for name, params in protos:
        params = " ".join(params.split())
        decls = [] if params in ("", "void") else params.split(",")
        for i, (d, a) in enumerate(zip(decls, _lib.SIGNATURES[name][1])):
            assert kind_c(d) == kinds_py.get(a, "ptr"), (name, i, d.strip(), a)


def test_reference_written_checkpoint_loads_without_the_reference():
    """`temporalpoints_last.tar` as run.py:813-819 writes it (oracle/make_golden_checkpoint.py ran the reference): model_kwargs
    pickles a lib.tineuvox.TiNeuVox; here `lib` is not importable, the loader maps it onto heads.TiNeuVoxHeads.  Every
    state-dict key of the reference model is present and loaded; save -> load round-trips; pcds/*.tar build a model."""
    import sys
    import tempfile
    from articulated_point_nerf_b200 import heads
    from articulated_point_nerf_b200.render import load_checkpoint, model_from_pcds, save_checkpoint
    gd = os.path.join(ROOT, "tests", "golden")
    model, ck = load_checkpoint(os.path.join(gd, "ref_mini_last.tar"), device="cpu")
    ref = torch.load(os.path.join(gd, "ref_mini_render.pt"), weights_only=False)
    assert ck["global_step"] == 1234 and not ck["missing_keys"] and not ck["unexpected_keys"]
    assert isinstance(model.tineuvox, (heads.TiNeuVoxHeads,)) or type(model.tineuvox).__name__ == "TiNeuVox"
    assert sorted(model.state_dict().keys()) == ref["state_keys"]
    N = len(model.canonical_pcd)
    assert 250 <= N <= 350 and model.weights.shape == (N, len(model.joints))
    assert abs(float(model.tineuvox.act_shift) - float(torch.log(torch.tensor(1 / (1 - 1e-3) - 1)))) < 1e-6
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "temporalpoints_last.tar")
        save_checkpoint(p, model, global_step=7)
        m2, ck2 = load_checkpoint(p, device="cpu")
        assert ck2["global_step"] == 7
        for k, v in model.state_dict().items():
            assert torch.equal(v, m2.state_dict()[k]), k
    m3 = model_from_pcds(os.path.join(gd, "ref_mini_pcds"), model.tineuvox, stepsize=0.5, fast_color_thres=1e-4)
    assert torch.equal(m3.canonical_pcd, model.canonical_pcd) and torch.equal(m3.canonical_feat.detach(), torch.load(
        os.path.join(gd, "ref_mini_pcds", "pcds", "canonical.tar"), weights_only=False)["feat"])


class _CapsuleField:
    """Analytic stand-in for the stage-1 voxel model behind export_point_cloud: density of a capsule around one segment, with
    the two methods the reference calls on its TiNeuVox (lib/tineuvox.py:238-250, 253-372)."""
    voxel_size = 0.05

    def __init__(self):
        self.xyz_min, self.xyz_max = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
        self.world_size = torch.tensor([24, 24, 24])
        self.calls = 0

    def get_grid_xyz(self, f):
        ax = [torch.linspace(float(self.xyz_min[i]), float(self.xyz_max[i]), int(int(self.world_size[i]) * f)) for i in range(3)]
        return torch.stack(torch.meshgrid(*ax, indexing="ij"), -1)

    def _alpha(self, p):
        a, b = torch.tensor([-0.5, 0., 0.]), torch.tensor([0.5, 0.1, 0.])
        tt = ((p - a) @ (b - a) / (b - a).dot(b - a)).clamp(0, 1)
        d = (p - (a + tt[:, None] * (b - a))).norm(dim=-1)
        return torch.sigmoid((0.25 - d) * 40)

    def get_grid_as_point_cloud(self, stepsize, time_sel, viewdir, threshold, sampling_freq, N_batch, alpha_xyz_only, grid_xyz=None):
        self.calls += 1
        g = self.get_grid_xyz(sampling_freq) if grid_xyz is None else grid_xyz
        shape = g.shape[:-1]
        alpha = self._alpha(g.reshape(-1, 3)).reshape(shape)
        if alpha_xyz_only:
            return None, None, None, None, None, None, g, alpha
        pts = g.reshape(-1, 3)
        return pts, alpha.reshape(-1), torch.rand(len(pts), 3), torch.rand(len(pts), 128), torch.rand(len(pts), 12), None, g, alpha


def test_export_point_cloud_on_an_analytic_field(tmp_path):
    """export.export_point_cloud (run.py:1081-1240): the frequency search ends near the requested point count, the volume clean-up
    keeps one component without small holes, canonical.tar / skeleton.tar have the reference's keys and feed model_from_pcds."""
    import numpy as np
    from articulated_point_nerf_b200 import heads
    from articulated_point_nerf_b200.export import export_point_cloud, preprocess_volume
    from articulated_point_nerf_b200.render import CANONICAL_KEYS, SKELETON_KEYS, model_from_pcds
    vol = np.zeros((20, 20, 20))
    vol[2:12, 2:12, 2:12] = 1.0
    vol[5:7, 5:7, 5:7] = 0.0                       # a small hole: filled
    vol[15:18, 15:18, 15:18] = 1.0                 # a second, smaller component: dropped
    m = preprocess_volume(vol, 0.5)
    assert m[5, 5, 5] and not m[16, 16, 16] and m.sum() == 1000
    field = _CapsuleField()

    def skeleton(binary_volume, grid_xyz, bone_length):
        pts = grid_xyz[binary_volume]
        j = np.stack([pts[pts[:, 0].argmin()], pts.mean(0), pts[pts[:, 0].argmax()]]).astype(np.float32)
        return {'skeleton_pcd': j, 'joints': j, 'root': j[0], 'bones': [[0, 1], [1, 2]], 'pcd': None, 'weights': None,
                'binary_volume': binary_volume}

    target = 3000
    can = export_point_cloud(field, str(tmp_path), viewdir=[0., 0., -1.], stepsize=0.5, threshold=0.2, canonical_pcd_num=target,
                             create_skeleton=skeleton)
    assert set(CANONICAL_KEYS) <= set(can) and abs(len(can["pcd"]) - target) < 0.15 * target      # the lattice size is int(world_size * freq): counts move in steps
    assert can['feat'].shape == (len(can['pcd']), 128) and can['alphas'].min() > 0.2
    sk = torch.load(tmp_path / 'pcds' / 'skeleton.tar', weights_only=False)
    assert set(SKELETON_KEYS) <= set(sk)
    calls = field.calls
    export_point_cloud(field, str(tmp_path), viewdir=[0., 0., -1.], stepsize=0.5, canonical_pcd_num=target, create_skeleton=skeleton)
    assert field.calls == calls                    # existing exports are left alone (run.py:1087-1089)
    tv = heads.TiNeuVoxHeads(can['xyz_min'].numpy(), can['xyz_max'].numpy(), num_voxels=16 ** 3, num_voxels_base=16 ** 3)
    model = model_from_pcds(str(tmp_path), tv, stepsize=0.5, fast_color_thres=1e-4)
    assert len(model.canonical_pcd) == len(can['pcd']) and len(model.joints) == 3


def test_nvtx_ranges_are_callable_without_a_tool():
    """apn_range_push / apn_range_pop (nvtx3, header-only): without an attached profiler the calls are no-ops that must not
    fail; `_lib.stage` emits them when APN_NVTX / _lib.nvtx(True) is on, named after the reference's profiler ranges."""
    from articulated_point_nerf_b200 import _lib
    lib = _lib.load()
    assert isinstance(lib.apn_range_push(b"sample_ray+knn+knn-post"), int)
    assert isinstance(lib.apn_range_pop(), int)
    ref_names = {"transform_net", "calc_rec_abs_T", "weighted_G_tw", "sample_ray", "knn", "knn-post", "feat_net", "densitynet", "rgbnet",
                 "poc_fre", "forward_warp", "pre-mask", "Alphas2Weights", "post-mask", "segment_coo"}   # lib/temporalpoints.py:421-653, pointwarper.py:217-241
    used = set()
    for v in _lib.NVTX_NAMES.values():
        used |= set(v.decode().split("+"))
    assert ref_names <= used | {"grid_build"}, ref_names - used
    _lib.nvtx(True)
    try:
        with _lib.stage("forward_warp"):
            pass
    finally:
        _lib.nvtx(False)
