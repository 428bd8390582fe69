"""Seeded synthetic scenes shaped like the reference's datasets.

Nothing here is on the hot path; it only manufactures inputs (skeleton, canonical
point cloud, features, cameras, rays) for tests, `bench.py` and `smoke()`.  Pure
CPU torch/numpy so the GPU box (no datasets, no /root/reference) can regenerate
the same tensors from the seed.

Conventions restated from the reference:
  * bone i == [parent, i+1] with parent < i+1   (lib/pointwarper.py:105-111)
  * cameras: pose_spherical(theta, phi, r)      (lib/load_dnerf.py:62-67)
  * rays: get_rays / get_rays_of_a_view         (lib/tineuvox.py:675-738)
  * scene bbox = frustum bbox * world_bound_scale (run.py:403-415, 824-827)
  * voxel_size = (prod(extent)/num_voxels)^(1/3) (lib/tineuvox.py:172)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

DNERF_CAMERA_ANGLE_X = 0.6911112070083618  # D-NeRF dataset constant (not in the reference repo)


# --------------------------------------------------------------------------------------
# skeletons
# --------------------------------------------------------------------------------------
def humanoid_skeleton(seed: int = 0):
    """21 joints / 20 bones, jumping-jack-like T-pose, z up.  Bone i = [parent, i+1]."""
    j = [
        (0.00, 0.00, 0.00),    # 0 pelvis (root)
        (0.00, 0.00, 0.12),    # 1 spine1
        (0.00, 0.00, 0.25),    # 2 spine2
        (0.00, 0.00, 0.37),    # 3 neck
        (0.00, 0.00, 0.50),    # 4 head
        (0.12, 0.00, 0.33),    # 5 l shoulder
        (0.30, 0.00, 0.33),    # 6 l elbow
        (0.46, 0.00, 0.33),    # 7 l wrist
        (0.54, 0.00, 0.33),    # 8 l hand
        (-0.12, 0.00, 0.33),   # 9 r shoulder
        (-0.30, 0.00, 0.33),   # 10
        (-0.46, 0.00, 0.33),   # 11
        (-0.54, 0.00, 0.33),   # 12
        (0.08, 0.00, -0.05),   # 13 l hip
        (0.09, 0.00, -0.32),   # 14 l knee
        (0.09, 0.00, -0.58),   # 15 l ankle
        (0.09, 0.08, -0.62),   # 16 l foot
        (-0.08, 0.00, -0.05),  # 17 r hip
        (-0.09, 0.00, -0.32),  # 18
        (-0.09, 0.00, -0.58),  # 19
        (-0.09, 0.08, -0.62),  # 20
    ]
    parents = [0, 1, 2, 3, 2, 5, 6, 7, 2, 9, 10, 11, 0, 13, 14, 15, 0, 17, 18, 19]
    rng = np.random.RandomState(seed)
    joints = np.asarray(j, dtype=np.float32) + rng.normal(0, 0.004, (len(j), 3)).astype(np.float32)
    bones = [[int(p), i + 1] for i, p in enumerate(parents)]
    return torch.from_numpy(joints), bones


def smpl_like_skeleton(seed: int = 0):
    """24 joints in the SMPL parent layout (zju_skeletons.py:5-9 describes the same tree)."""
    parents = [0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
    off = {
        1: (0.07, 0, -0.09), 2: (-0.07, 0, -0.09), 3: (0, 0, 0.11), 4: (0.03, 0, -0.38), 5: (-0.03, 0, -0.38),
        6: (0, 0, 0.13), 7: (0, 0, -0.40), 8: (0, 0, -0.40), 9: (0, 0, 0.06), 10: (0, 0.10, -0.05), 11: (0, 0.10, -0.05),
        12: (0, 0, 0.21), 13: (0.08, 0, 0.12), 14: (-0.08, 0, 0.12), 15: (0, 0, 0.09), 16: (0.10, 0, 0.03),
        17: (-0.10, 0, 0.03), 18: (0.26, 0, 0), 19: (-0.26, 0, 0), 20: (0.25, 0, 0), 21: (-0.25, 0, 0),
        22: (0.08, 0, 0), 23: (-0.08, 0, 0),
    }
    rng = np.random.RandomState(seed)
    joints = np.zeros((24, 3), dtype=np.float32)
    for c in range(1, 24):
        joints[c] = joints[parents[c - 1]] + np.asarray(off[c], dtype=np.float32)
    joints += rng.normal(0, 0.003, joints.shape).astype(np.float32)
    bones = [[int(parents[c - 1]), c] for c in range(1, 24)]
    return torch.from_numpy(joints), bones


def random_tree_skeleton(n_joints: int, seed: int = 0, seg=(0.08, 0.15)):
    """Seeded random kinematic tree; bone i = [parent < i+1, i+1]."""
    rng = np.random.RandomState(seed)
    joints = np.zeros((n_joints, 3), dtype=np.float32)
    bones = []
    for c in range(1, n_joints):
        lo = max(0, c - 4)
        p = int(rng.randint(lo, c)) if rng.rand() < 0.8 else int(rng.randint(0, c))
        d = rng.normal(size=3)
        d /= np.linalg.norm(d) + 1e-9
        joints[c] = joints[p] + d * rng.uniform(*seg)
        bones.append([p, c])
    joints -= joints.mean(0, keepdims=True)
    return torch.from_numpy(joints.astype(np.float32)), bones


# --------------------------------------------------------------------------------------
# canonical cloud: lattice nodes inside capsules around bones (mimics run.py:1152-1201,
# where canonical points are voxel-grid nodes)
# --------------------------------------------------------------------------------------
def _seg_dist(p: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    s = b - a
    w = p - a
    l2 = float((s * s).sum())
    t = np.clip((w @ s) / max(l2, 1e-12), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[:, None] * s), axis=-1)


def _lattice_in_capsules(joints: np.ndarray, bones, h: float, radius: float) -> np.ndarray:
    keys = []
    for (pa, ch) in bones:
        a, b = joints[pa].astype(np.float64), joints[ch].astype(np.float64)
        lo = np.floor((np.minimum(a, b) - radius) / h).astype(np.int64)
        hi = np.ceil((np.maximum(a, b) + radius) / h).astype(np.int64)
        ax = [np.arange(lo[d], hi[d] + 1) for d in range(3)]
        g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
        d = _seg_dist(g * h, a, b)
        g = g[d < radius]
        keys.append(((g[:, 0] + (1 << 20)) << 42) | ((g[:, 1] + (1 << 20)) << 21) | (g[:, 2] + (1 << 20)))
    k = np.unique(np.concatenate(keys))
    g = np.stack([(k >> 42) - (1 << 20), ((k >> 21) & ((1 << 21) - 1)) - (1 << 20), (k & ((1 << 21) - 1)) - (1 << 20)], -1)
    return (g * h).astype(np.float32)


def capsule_lattice_cloud(joints: torch.Tensor, bones, n_target: int, radius: float = 0.045):
    jn = joints.numpy().astype(np.float64)
    vol = sum(math.pi * radius ** 2 * np.linalg.norm(jn[a] - jn[b]) + 4 / 3 * math.pi * radius ** 3 for a, b in bones)
    h = (vol / n_target) ** (1 / 3)
    pts = _lattice_in_capsules(jn, bones, h, radius)
    for _ in range(6):
        if abs(len(pts) - n_target) <= max(2, 0.002 * n_target):
            break
        h *= (len(pts) / n_target) ** (1 / 3)
        pts = _lattice_in_capsules(jn, bones, h, radius)
    return torch.from_numpy(pts), float(h)


# --------------------------------------------------------------------------------------
# cameras and rays
# --------------------------------------------------------------------------------------
def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> torch.Tensor:
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    trans = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=np.float64)
    rphi = np.array([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0], [0, math.sin(ph), math.cos(ph), 0], [0, 0, 0, 1]])
    rth = np.array([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1]])
    swap = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
    return torch.from_numpy((swap @ rth @ rphi @ trans).astype(np.float32))


def intrinsics(H: int, W: int, focal: float) -> torch.Tensor:
    return torch.tensor([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=torch.float32)


def get_rays(H, W, K, c2w, inverse_y=False, flip_x=False, flip_y=False):
    """Pixel-centre rays; restates lib/tineuvox.py:675-703 (mode='center')."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t().float() + 0.5
    j = j.t().float() + 0.5
    if flip_x:
        i = i.flip((1,))
    if flip_y:
        j = j.flip((0,))
    if inverse_y:
        dirs = torch.stack([(i - K[0][2]) / K[0][0], (j - K[1][2]) / K[1][1], torch.ones_like(i)], -1)
    else:
        dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    return rays_o, rays_d


def get_rays_of_a_view(H, W, K, c2w, inverse_y=False, flip_x=False, flip_y=False):
    """lib/tineuvox.py:733-738 with ndc=False."""
    rays_o, rays_d = get_rays(H, W, K, c2w, inverse_y, flip_x, flip_y)
    viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
    return rays_o, rays_d, viewdirs


def frustum_bbox(HW, Ks, poses, near, far, inverse_y=False, scale=1.05):
    """run.py:403-415 + world_bound_scale (run.py:824-827)."""
    lo = torch.full((3,), float("inf"))
    hi = -lo
    for (H, W), K, c2w in zip(HW, Ks, poses):
        # the extremes of o + viewdir*{near,far} are reached on the image border or centre
        Hs, Ws = min(H, 65), min(W, 65)
        Kc = K.clone()
        Kc[0] *= Ws / W
        Kc[1] *= Hs / H
        ro, rd, vd = get_rays_of_a_view(Hs, Ws, Kc, c2w, inverse_y=inverse_y)
        pts = torch.stack([ro + vd * near, ro + vd * far])
        lo = torch.minimum(lo, pts.amin((0, 1, 2)))
        hi = torch.maximum(hi, pts.amax((0, 1, 2)))
    shift = (hi - lo) * (scale - 1) / 2
    return lo - shift, hi + shift


# --------------------------------------------------------------------------------------
# skinning-weight initialisation (lib/temporalpoints.py:206-254)
# --------------------------------------------------------------------------------------
def weights_from_bones(joints: torch.Tensor, bones, pcd: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    a = torch.stack([joints[b[0]] for b in bones])  # (B,3)
    b = torch.stack([joints[b[1]] for b in bones])
    s = b - a
    w = pcd[None] - a[:, None]                       # (B,N,3)
    ps = (w * s[:, None]).sum(-1)
    l2 = (s * s).sum(-1)[:, None]
    t = (ps / l2.clamp_min(1e-20)).clamp(0, 1)
    d = (pcd[None] - (a[:, None] + t[..., None] * s[:, None])).norm(dim=-1)   # (B,N)
    wts = (1.0 / (0.5 * math.e ** d + eps)).T.contiguous()
    return torch.cat([torch.zeros(len(wts), 1), wts], -1)


# --------------------------------------------------------------------------------------
# scene description
# --------------------------------------------------------------------------------------
@dataclass
class SceneConfig:
    name: str = "c1"
    n_points: int = 30000
    skeleton: str = "humanoid"          # humanoid | smpl | random
    n_joints: int = 21
    body_scale: float = 1.0
    H: int = 400
    W: int = 400
    n_views: int = 8
    cam_radius: float = 4.0
    cam_phi: float = -20.0
    near: float = 2.0
    far: float = 6.0
    bg: float = 1.0
    inverse_y: bool = False
    pose_embedding_dim: int = 0
    stepsize: float = 0.5
    fast_color_thres: float = 1e-4
    num_voxels: int = 160 ** 3
    feat_dim: int = 128
    seed: int = 0


CONFIGS = {
    # BASELINE.json configs[0..4]
    "tiny": SceneConfig(name="tiny", n_points=1500, H=40, W=40, n_views=4),
    "small": SceneConfig(name="small", n_points=6000, H=64, W=64, n_views=4),
    # c4 in miniature (configs/zju/default.py:72,112; lib/load_data.py:45-46): pose embedding, black background, OpenCV rays
    "tiny_pose": SceneConfig(name="tiny_pose", n_points=1500, skeleton="random", n_joints=10, H=40, W=40, n_views=4,
                             cam_radius=3.0, near=1.0, far=4.0, bg=0.0, inverse_y=True, pose_embedding_dim=64),
    "c1": SceneConfig(name="c1", n_points=30000, H=400, W=400, n_views=8),
    "c2": SceneConfig(name="c2", n_points=30000, H=400, W=400, n_views=8),
    "c3": SceneConfig(name="c3", n_points=30000, skeleton="random", n_joints=24, body_scale=2.5, H=1024, W=1024,
                      n_views=8, near=1.0, far=6.0),
    "c4": SceneConfig(name="c4", n_points=100000, skeleton="smpl", n_joints=24, body_scale=1.0, H=512, W=512,
                      n_views=21, cam_radius=3.0, near=1.0, far=4.0, bg=0.0, inverse_y=True, pose_embedding_dim=64),
    "c5": SceneConfig(name="c5", n_points=1000000, skeleton="random", n_joints=65, body_scale=2.5, H=2048, W=2048,
                      n_views=8, near=1.0, far=6.0),
}


@dataclass
class Scene:
    cfg: SceneConfig
    joints: torch.Tensor
    bones: List[List[int]]
    canonical_pcd: torch.Tensor
    lattice_h: float
    canonical_feat: torch.Tensor
    canonical_rgbs: torch.Tensor
    canonical_alpha: torch.Tensor
    skeleton_pcd: torch.Tensor
    xyz_min: torch.Tensor
    xyz_max: torch.Tensor
    voxel_size: float
    HW: List
    Ks: torch.Tensor
    poses: torch.Tensor
    extra: dict = field(default_factory=dict)

    def rays(self, view: int = 0):
        H, W = self.HW[view]
        return get_rays_of_a_view(H, W, self.Ks[view], self.poses[view], inverse_y=self.cfg.inverse_y)

    def render_kwargs(self):
        c = self.cfg
        return dict(near=c.near, far=c.far, bg=c.bg, stepsize=c.stepsize, inverse_y=c.inverse_y,
                    flip_x=False, flip_y=False)


def make_scene(cfg: SceneConfig | str) -> Scene:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    g = torch.Generator().manual_seed(cfg.seed)
    if cfg.skeleton == "humanoid":
        joints, bones = humanoid_skeleton(cfg.seed)
    elif cfg.skeleton == "smpl":
        joints, bones = smpl_like_skeleton(cfg.seed)
    else:
        joints, bones = random_tree_skeleton(cfg.n_joints, cfg.seed)
    joints = joints * cfg.body_scale
    pcd, h = capsule_lattice_cloud(joints, bones, cfg.n_points, radius=0.045 * cfg.body_scale)
    N = len(pcd)
    feat = torch.relu(torch.randn(N, cfg.feat_dim, generator=g)) * 0.5
    rgbs = torch.rand(N, 3, generator=g)
    alpha = torch.rand(N, generator=g)
    # skeleton point cloud: points along the bones (stands in for the thinned volume, skeletonizer.py)
    sk = []
    for a, b in bones:
        tt = torch.linspace(0, 1, 8)[:, None]
        sk.append(joints[a][None] * (1 - tt) + joints[b][None] * tt)
    skeleton_pcd = torch.cat(sk)

    focal_scale = 0.5 / math.tan(0.5 * DNERF_CAMERA_ANGLE_X)
    HW, Ks, poses = [], [], []
    for v in range(cfg.n_views):
        th = 180.0 - 360.0 * v / cfg.n_views
        c2w = pose_spherical(th, cfg.cam_phi, cfg.cam_radius)
        if cfg.inverse_y:
            c2w = c2w @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))  # OpenCV-style camera
        HW.append((cfg.H, cfg.W))
        Ks.append(intrinsics(cfg.H, cfg.W, focal_scale * cfg.W))
        poses.append(c2w)
    Ks = torch.stack(Ks)
    poses = torch.stack(poses)
    lo, hi = frustum_bbox(HW, Ks, poses, cfg.near, cfg.far, inverse_y=cfg.inverse_y)
    voxel_size = float(((hi - lo).prod() / cfg.num_voxels) ** (1 / 3))
    return Scene(cfg, joints, bones, pcd, h, feat, rgbs, alpha, skeleton_pcd, lo, hi, voxel_size, HW, Ks, poses)


# --------------------------------------------------------------------------------------
# model construction on top of a scene (tests, bench.py, smoke())
# --------------------------------------------------------------------------------------
def build_model(scene: Scene, seed: int = 0, density_bias: float = 7.0, theta_std: float = 0.2,
                density_gain: float = 300.0, rgb_gain: float = 8.0, device=None, density_std: float = 2.5,
                no_view_dir: bool = False, frozen_view_dir=None):
    """TemporalPoints on random-init weights of the reference's architecture, with the output heads rescaled so
    that kept-sample alpha spreads over (0,1) and early ray termination triggers."""
    from .heads import TiNeuVoxHeads, poc_fre
    from .temporalpoints import TemporalPoints
    torch.manual_seed(seed)
    cfg = scene.cfg
    heads = TiNeuVoxHeads(scene.xyz_min.numpy(), scene.xyz_max.numpy(), num_voxels=cfg.num_voxels,
                          num_voxels_base=cfg.num_voxels, alpha_init=1e-3, net_width=128, no_view_dir=no_view_dir)
    model = TemporalPoints(
        canonical_pcd=scene.canonical_pcd.clone(), canonical_alpha=scene.canonical_alpha.clone(),
        canonical_feat=scene.canonical_feat.clone(), canonical_rgbs=scene.canonical_rgbs.clone(),
        skeleton_pcd=scene.skeleton_pcd.clone(), joints=scene.joints.clone(), bones=scene.bones,
        xyz_min=scene.xyz_min.numpy(), xyz_max=scene.xyz_max.numpy(), tineuvox=heads, stepsize=cfg.stepsize,
        voxel_size=scene.voxel_size, fast_color_thres=cfg.fast_color_thres, pose_embedding_dim=cfg.pose_embedding_dim,
        frozen_view_dir=frozen_view_dir)
    with torch.no_grad():
        # density head calibrated on a CPU sample of decoder features (random points, offsets of the size the k-NN
        # produces): pre-activation density + act_shift ~ N(0, density_std) => median kept-sample alpha ~ 0.3, a spread over
        # (0.02, 0.9) and early ray termination on the thicker parts; `density_gain` is the fallback scale
        gcal = torch.Generator().manual_seed(seed + 1234)
        sel = torch.randint(0, len(scene.canonical_pcd), (512,), generator=gcal)
        rel = torch.randn(512, 3, generator=gcal) * (0.7 * float(scene.lattice_h))
        x = torch.cat([poc_fre(rel, model.pos_poc), model.canonical_feat[sel]], dim=-1)
        if cfg.pose_embedding_dim > 0:      # rest pose: joint offsets 0 (lib/temporalpoints.py:571-574)
            e = model.pose_embedding_net(poc_fre(torch.zeros_like(model.joints), model.pos_poc).view(1, -1))
            x = torch.cat([x, e.expand(len(x), -1)], dim=-1)
        if x.shape[-1] == model.feat_net[0].in_features:
            z = model.densitynet(model.feat_net(x)).reshape(-1)
            gain = density_std / max(float(z.std()), 1e-6)
            model.densitynet.weight.mul_(gain)
            model.densitynet.bias.fill_(-float(heads.act_shift) - gain * float(z.mean() - model.densitynet.bias.reshape(-1)[0]))
        else:                      # pose-embedding configs: the decoder input has extra columns
            model.densitynet.bias.fill_(density_bias)
            model.densitynet.weight.mul_(density_gain)
        model.rgbnet.views_linears[2].weight.mul_(rgb_gain)
        t_embed = poc_fre(torch.tensor([0.37]), model.time_poc)
        out = model.forward_warp.transform_net(t_embed.unsqueeze(0))
        model.forward_warp.transform_net.net[-1].weight.mul_(theta_std / float(out.std()))
    if device is not None:
        model = model.to(device)
    return model
