"""TemporalPoints with the reference's call surface (lib/temporalpoints.py), on the B200 kernels.

`forward(t | rot_params, render_kwargs{rays_o, rays_d, viewdirs, near, far, stepsize, bg, ...})`
returns the reference's dict (lib/temporalpoints.py:598-609,679-712): t_hat_pcd, rgb_marched,
alphainv_last, alphainv_last_direct, rgb_marched_direct, depth, weights, joints, bones, grid.

Pipeline (reference stage -> kernel):
  forward_warp  get_weights + PointWarper blend + torch.inverse + bbox   -> csrc/lbs.cu (one launch)
  sample_ray / knn / knn-post  DVGO sampler + KeOps brute force + masks  -> csrc/grid_knn.cu
  feat_net / densitynet / rgbnet                                         -> csrc/aggregate*.cu
  pre-mask / Alphas2Weights / post-mask / segment_coo                    -> csrc/composite.cu
State-dict keys, constructor kwargs and side-effect attributes (`_last_weights`, `nn_i`,
`forward_warp.prev_*`) follow the reference so run.py's callers and checkpoints keep working.
"""
from __future__ import annotations

import colorsys
import math
from typing import Optional

import numpy as np
import torch

from . import ops
from .heads import poc_fre
from .pointwarper import PointWarper


class NoPointsException(Exception):
    pass


class _PoseChain(torch.nn.Module):
    """t_embed, joints -> bone transforms: TransformNet + Rodrigues + kinematic chain (lib/pointwarper.py:217-236)
    as ONE unit, so that the ~250 tiny launches of its forward + backward can be replayed as two CUDA graphs
    (torch.cuda.make_graphed_callables) inside the training step.  Never registered as a sub-module of
    TemporalPoints: the parameters stay under `forward_warp.*`."""

    def __init__(self, warper: PointWarper):
        super().__init__()
        self.warper = warper

    def forward(self, t_embed, joints):
        bone_Ts, global_t = self.warper.pose(joints, t=t_embed)
        jh = torch.cat([joints, torch.ones((len(joints), 1), device=joints.device, dtype=joints.dtype)], dim=-1)
        joints_rel = torch.bmm(bone_Ts, jh.unsqueeze(-1)).squeeze(-1)[:, :3]
        return bone_Ts, global_t, self.warper.prev_thetas, joints_rel


def hls_palette(n: int):
    """seaborn.color_palette('hls', n) (lib/temporalpoints.py:692) without the seaborn dependency."""
    hues = np.linspace(0, 1, int(n) + 1)[:-1] + 0.01
    hues %= 1
    return [colorsys.hls_to_rgb(float(h), 0.6, 0.65) for h in hues]


def project_point_to_image_plane(points, pose, intrinsic):
    """lib/utils.py:435-450."""
    points = points.unsqueeze(0).expand(len(pose), -1, -1)
    pose = pose.inverse()
    points = torch.bmm(pose[:, :3, :3], points.transpose(1, 2)).transpose(1, 2) + pose[:, :3, 3:].transpose(1, 2)
    points = torch.bmm(intrinsic, points.transpose(1, 2)).transpose(1, 2)
    return points[:, :, :2] / points[:, :, 2:]


class TemporalPoints(torch.nn.Module):
    def __init__(self, canonical_pcd, canonical_alpha, canonical_feat, canonical_rgbs, skeleton_pcd, joints, bones,
                 xyz_min, xyz_max, tineuvox, neighbours=8, timebase_pe=8, eps=1e-6, stepsize=None, voxel_size=None,
                 fast_color_thres=0, embedding='full', frozen_view_dir=None, over_parameterized_rot=True,
                 re_init_feat=False, re_init_mlps=False, feat_depth=4, pose_embedding_dim=0, **kwargs):
        super().__init__()
        assert neighbours == ops.K_NEIGHBOURS, "the B200 kernels are specialised for 8 neighbours (lib/temporalpoints.py:42)"
        assert feat_depth == 4, "the fused decoder is specialised for feat_depth=4 (lib/temporalpoints.py:53)"
        assert canonical_feat.shape[-1] == ops.FEAT_DIM, "the fused decoder is specialised for 128 channels"
        canonical_pcd = torch.as_tensor(canonical_pcd).float()
        joints = torch.as_tensor(joints).float()
        self.canonical_pcd = canonical_pcd
        self.skeleton_pcd = skeleton_pcd
        self.bones = bones
        self.bone_arap_mask = torch.tensor(bones).reshape(-1)
        self.register_buffer('xyz_min', torch.as_tensor(np.asarray(xyz_min)).float())
        self.register_buffer('xyz_max', torch.as_tensor(np.asarray(xyz_max)).float())
        self.eps = torch.tensor(float(eps))
        self.feat_depth = feat_depth
        self.timebase_pe = timebase_pe
        self.t_dim = 1 + self.timebase_pe * 2
        self.stepsize = stepsize
        self.voxel_size = voxel_size
        self.fast_color_thres = fast_color_thres
        self.embedding = embedding
        self.over_parameterized_rot = over_parameterized_rot
        self.joints_to_keep = None
        self.forward_warp_t_dim = self.t_dim

        w0 = kwargs.get('weights', None)
        if w0 is None:
            w0 = self._weights_from_bones(joints, bones, canonical_pcd, add_zero_weight=True)
        self.weights = torch.nn.Parameter(torch.as_tensor(w0).detach().clone().float(), requires_grad=True)
        self.forward_warp = PointWarper(canonical_pcd=canonical_pcd, t_dim=self.forward_warp_t_dim, joints=joints,
                                        bones=bones, over_parameterized_rot=over_parameterized_rot)
        self.original_joints = torch.nn.Parameter(joints.clone(), requires_grad=False)
        self.joints = torch.nn.Parameter(joints.clone(), requires_grad=True)
        self.canonical_feat = torch.nn.Parameter(torch.as_tensor(canonical_feat).detach().clone().float(), requires_grad=True)
        if re_init_feat:
            self.canonical_feat.data = torch.randn_like(self.canonical_feat)
        self.theta_weight = torch.nn.Parameter(torch.tensor([0.1]), requires_grad=True)
        self.merging_dict = None
        gammas = torch.ones(len(canonical_pcd))
        self.gammas = torch.nn.Parameter(gammas + torch.randn_like(gammas) * 1e-2, requires_grad=True)
        self.pruned_joints = torch.zeros(len(joints), dtype=bool)
        self.register_buffer('flat_merging_rules', torch.arange(0, len(joints)))
        self.register_buffer('sibling_merging_rules', torch.zeros(len(joints), dtype=bool))
        self.canonical_rgbs = torch.nn.Parameter(torch.as_tensor(canonical_rgbs).detach().clone().float(), requires_grad=True)
        self.canonical_alpha = torch.nn.Parameter(torch.as_tensor(canonical_alpha).detach().clone().float(), requires_grad=True)
        self.direct_eps = torch.nn.Parameter(torch.tensor([0.05] * len(canonical_pcd)), requires_grad=True)
        self.register_buffer('time_poc', torch.FloatTensor([(2 ** i) for i in range(timebase_pe)]))
        self.neighbours = neighbours
        # canonical neighbourhood (lib/temporalpoints.py:104-111): built lazily on the device
        self.nn_i = None
        self.nn_distance = None
        self.mean_min_distance = None

        feat_w = self.canonical_feat.shape[-1]
        feat_input_dim = feat_w + 3 + 3 * tineuvox.posbase_pe * 2 + pose_embedding_dim
        self.feat_net = torch.nn.Sequential(
            torch.nn.Linear(feat_input_dim, feat_w), torch.nn.LeakyReLU(inplace=True),
            *[torch.nn.Sequential(torch.nn.Linear(feat_w, feat_w), torch.nn.LeakyReLU(inplace=True))
              for _ in range(feat_depth - 2)],
            torch.nn.Linear(feat_w, feat_w), torch.nn.LeakyReLU(inplace=True))
        self.rgbnet = tineuvox.rgbnet
        self.densitynet = tineuvox.densitynet
        self.timenet = tineuvox.timenet
        if re_init_mlps:
            for net in (self.rgbnet, self.densitynet, self.timenet):
                for m in net.modules():
                    if hasattr(m, 'reset_parameters'):
                        m.reset_parameters()
        self.view_poc = tineuvox.view_poc
        self.pos_poc = tineuvox.pos_poc
        # no_view_dir (lib/tineuvox.py:112-113, lib/temporalpoints.py:504-505): the RGB head has no view columns.  The
        # kernels always form [f | PE(view)]: the head's first view layer is handed over with 27 zero columns appended
        # (_mlp_weights), which contributes exact zeros — the product of finite PE values with 0.
        self.no_view_dir = tineuvox.no_view_dir
        self.tineuvox = tineuvox
        self.register_buffer('xyz_max_canonical', canonical_pcd.max(dim=0)[0])
        self.register_buffer('xyz_min_canonical', canonical_pcd.min(dim=0)[0])
        # frozen_view_dir (run.py:480-481 `use_global_view_dir`; lib/temporalpoints.py:155-159,507-508): one view direction
        # for every ray.  `viewdirs_emb` is the reference's (frozen) parameter — state-dict surface; the kernels embed the
        # direction themselves, so every ray simply carries `frozen_view_dir` (_view_dirs).
        self.frozen_view_dir = None if frozen_view_dir is None else torch.as_tensor(frozen_view_dir).float().reshape(3)
        if self.frozen_view_dir is not None:
            self.viewdirs_emb = torch.nn.Parameter(poc_fre(self.frozen_view_dir, self.view_poc)[None], requires_grad=False)
        object.__setattr__(self, '_frozen_dirs', None)
        self.pose_embedding_dim = pose_embedding_dim
        if pose_embedding_dim > 0:
            d = len(joints) * (3 * len(self.pos_poc) * 2 + 3)
            self.pose_embedding_net = torch.nn.Sequential(
                torch.nn.Linear(d, d // 2), torch.nn.LeakyReLU(inplace=True),
                *[torch.nn.Sequential(torch.nn.Linear(d // 2, d // 2), torch.nn.LeakyReLU(inplace=True))
                  for _ in range(feat_depth - 2)],
                torch.nn.Linear(d // 2, pose_embedding_dim), torch.nn.LeakyReLU(inplace=True))
        self.beta = torch.nn.Parameter(torch.tensor([0.5]), requires_grad=True)
        self.beta_min = torch.nn.Parameter(torch.tensor([0.0001]), requires_grad=False)
        self._last_weights_value = None
        self.last_counts = {}
        # decoder used when no gradient is needed: "tc" = tcgen05 split-fp16 (fp32-class), "tc_fast" = tcgen05 fp16
        # operands, "fp32" = CUDA-core exact path (always used when autograd is recording)
        self.decoder = "tc"
        # decoder used while autograd records: "tc" = tensor-core forward + backward (split-fp16; d_in = 191 or 255),
        # "fp32" = CUDA-core exact path
        self.decoder_train = "fp32"
        self._packed_decoder = ops.PackedDecoder()
        # replay the PyTorch pose chain (fwd + bwd) as CUDA graphs while training; only relevant when the fused pose
        # kernel (forward_warp.fused_pose) does not cover the tree / MLP shape.  Opt-in: torch.cuda.make_graphed_callables
        # returns its gradients in static buffers, and a SECOND backward through one forward (retain_graph=True; e.g. separate
        # autograd.grad calls per loss term) silently yields wrong joint gradients (found with the full stage-2 loss test)
        self.graph_pose = False
        object.__setattr__(self, '_pose_graph', None)
        object.__setattr__(self, '_pose_graph_key', None)

    # ------------------------------------------------------------------------------------------
    def get_kwargs(self):
        """lib/temporalpoints.py:176-200."""
        return {
            'canonical_pcd': self.canonical_pcd, 'skeleton_pcd': self.skeleton_pcd, 'canonical_alpha': self.canonical_alpha,
            'canonical_feat': self.canonical_feat, 'canonical_rgbs': self.canonical_rgbs, 'joints': self.joints,
            'bones': self.bones, 'neighbours': self.neighbours, 'timebase_pe': self.timebase_pe, 'eps': self.eps,
            'stepsize': self.stepsize, 'weights': self.weights, 'xyz_min': self.xyz_min.cpu().numpy(),
            'xyz_max': self.xyz_max.cpu().numpy(), 'tineuvox': self.tineuvox, 'voxel_size': self.voxel_size,
            'fast_color_thres': self.fast_color_thres, 'embedding': self.embedding, 'frozen_view_dir': self.frozen_view_dir,
            'over_parameterized_rot': self.over_parameterized_rot, 'feat_depth': self.feat_depth,
            'pose_embedding_dim': self.pose_embedding_dim,
        }

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        # plain-tensor attributes the reference leaves to the default CUDA tensor type (run.py:1248-1249)
        self.canonical_pcd = fn(self.canonical_pcd)
        self.forward_warp.canonical_pcd = self.canonical_pcd
        # aliases of the heads' frequency buffers (lib/temporalpoints.py:150-151): follow the moved buffers
        self.view_poc, self.pos_poc = self.tineuvox.view_poc, self.tineuvox.pos_poc
        if torch.is_tensor(self.skeleton_pcd):
            self.skeleton_pcd = fn(self.skeleton_pcd)
        for name in ('nn_i', 'nn_distance', 'mean_min_distance'):
            v = getattr(self, name)
            if v is not None:
                setattr(self, name, fn(v))
        return self

    @property
    def device(self):
        return self.weights.device

    def reinitialise_weights(self):
        self.weights.data = self._weights_from_bones(self.joints.detach(), self.bones, self.canonical_pcd,
                                                     add_zero_weight=True).to(self.weights.device)
        self.theta_weight.data = torch.tensor([0.1], device=self.theta_weight.device)

    @staticmethod
    def dist_batch(p, a, b):
        """Point-to-segment distances (B segments x N points), lib/temporalpoints.py:206-233."""
        s = b - a
        w = p[None] - a[:, None]
        ps = (w * s[:, None]).sum(-1)
        l2 = (s * s).sum(-1)[:, None]
        t = (ps / l2.clamp_min(1e-30)).clamp(0, 1)
        t = torch.where(ps <= 0, torch.zeros_like(t), t)
        return (p[None] - (a[:, None] + t[..., None] * s[:, None])).norm(dim=-1)

    def _weights_from_bones(self, joints, bones, pcd, add_noise=False, noise_var=0, val=1, soft_weights=True,
                            add_zero_weight=False):
        """lib/temporalpoints.py:235-254."""
        d = self.dist_batch(pcd, torch.stack([joints[b[0]] for b in bones]), torch.stack([joints[b[1]] for b in bones]))
        if soft_weights:
            weights = (1 / (0.5 * torch.e ** d + float(self.eps))).T.contiguous()
        else:
            am = torch.argmin(d, dim=0)
            weights = torch.zeros((len(am), len(bones)), device=pcd.device)
            weights[torch.arange(len(am)), am] = val
        if add_zero_weight:
            weights = torch.cat([torch.zeros((len(weights), 1), device=weights.device), weights], dim=-1)
        if add_noise:
            weights = weights + torch.randn_like(weights) * noise_var
        return weights

    def _ensure_neighbourhood(self):
        """nn_i / nn_distance / mean_min_distance (lib/temporalpoints.py:104-111), via the grid k-NN instead
        of the N x N KeOps reduction."""
        if self.nn_i is not None:
            return
        pcd = self.canonical_pcd
        if not pcd.is_cuda:
            raise RuntimeError("TemporalPoints needs its tensors on a CUDA device (move the module with .cuda())")
        lo, hi = pcd.min(0)[0], pcd.max(0)[0]
        bbox = torch.cat([lo, hi])
        vol = float((hi - lo).clamp_min(1e-6).prod())
        spacing = (vol / max(len(pcd), 1)) ** (1 / 3)
        # the point spacing inside the occupied volume is finer than the bbox average; refine once
        grid = ops.Grid(pcd, bbox, query_radius=0.01, bbox_pad=0.01, cell_hint=max(spacing * 0.5, 1e-4))
        idx, d2 = grid.knn(pcd, self.neighbours)
        assert int(idx.min()) >= 0, "k-NN returned an invalid neighbour index"
        self.nn_i = idx.long()
        self.nn_distance = torch.sqrt(((pcd[:, None, :] - pcd[self.nn_i, :]) ** 2).sum(-1) + self.eps.to(pcd.device))
        self.mean_min_distance = self.nn_distance[:, 1].mean()
        self._mmd_float = float(self.mean_min_distance)
        bam = self.bone_arap_mask.to(pcd.device)
        self.og_joint_distance = (self.original_joints[bam][0::2, :] - self.original_joints[bam][1::2, :])

    def get_weights(self):
        """lib/temporalpoints.py:401-414: softmax(raw / max(eps, theta)) then column merge."""
        theta = torch.max(self.eps.to(self.weights.device), self.theta_weight)
        w = torch.softmax(self.weights / theta, dim=-1)
        rules = self.flat_merging_rules.to(w.device)
        return torch.zeros_like(w).index_add_(1, rules, w)

    def repose(self, rot_params):
        return self.forward_warp(self.get_weights(), self.joints, rot_params=rot_params)

    def flatten_merging_rules(self, merging_rules):
        endpoints = []
        for i in range(len(merging_rules)):
            j = i
            while True:
                j = merging_rules[j]
                if j == merging_rules[j]:
                    endpoints.append(j)
                    break
        return endpoints

    @staticmethod
    def _rotation_angle(rel):
        """Rotation angle in [0, pi] of (..., 3, 3) rotation matrices — the norm of roma.rotmat_to_rotvec
        (lib/temporalpoints.py:358-361) — as atan2(|axial vector|, (trace - 1) / 2)."""
        vx = rel[..., 2, 1] - rel[..., 1, 2]
        vy = rel[..., 0, 2] - rel[..., 2, 0]
        vz = rel[..., 1, 0] - rel[..., 0, 1]
        s = 0.5 * torch.sqrt(vx * vx + vy * vy + vz * vz)
        c = 0.5 * (rel[..., 0, 0] + rel[..., 1, 1] + rel[..., 2, 2] - 1.0)
        return torch.atan2(s, c)

    def _are_rotations_similar(self, rot1, rot2, deg_threshold=20, five_percent_heuristic=False):
        """lib/temporalpoints.py:356-369 for (T, ..., 3, 3) stacks; reduces over the leading (time) axis."""
        angle = self._rotation_angle(rot1 @ rot2.transpose(-1, -2))
        if not five_percent_heuristic:
            return torch.rad2deg(torch.sqrt((angle ** 2).mean(dim=0))) <= deg_threshold
        th = int(len(rot1) * 0.05)
        return (torch.rad2deg(angle) >= deg_threshold).sum(dim=0) <= th

    @torch.no_grad()
    def simplify_skeleton(self, times, deg_threshold=10, mass_threshold=0.0, update_skeleton=False,
                          five_percent_heuristic=False, visualise_canonical=False):
        """lib/temporalpoints.py:256-343 (run.py:1302-1308, --degree_threshold): evaluate the pose network at every
        training time, freeze joints that never rotate by more than the threshold, merge siblings that rotate alike,
        and install the result as `forward_warp.rot_mask`, `forward_warp.sibling_mask` and `flat_merging_rules`
        (the skinning-weight column merge done inside the LBS kernel).  Same return tuple as the reference.
        Reference quirks kept: the 'average' heuristic thresholds the mean SQUARED angle converted to degrees
        (no square root, :288); the last returned item is whatever the reference's `res` held last."""
        from .treeprune import merge_joints
        J = len(self.joints)
        times = torch.as_tensor(times, dtype=torch.float32, device=self.time_poc.device).reshape(-1, 1)
        assert len(times) > 1, "simplify_skeleton needs more than one time step (TransformNet drops the batch axis for one)"
        params = self.forward_warp.transform_net(poc_fre(times, self.time_poc))          # (T, J+1, 4)
        T = len(times)
        if self.over_parameterized_rot:
            rot_angles = params[:, :J, -1]
            R, _ = self.forward_warp.Rodrigues(params[:, :J, :].reshape(T * J, 4))
        else:
            rot_angles = (params[:, :J, :3] ** 2).sum(-1).sqrt() % (2 * np.pi)
            R, _ = self.forward_warp.Rodrigues(params[:, :J, :3].reshape(T * J, 3))
        R = R.reshape(T, J, 3, 3)
        # all joint pairs at once (the reference loops over the lower triangle and mirrors it)
        pair = self._are_rotations_similar(R[:, :, None], R[:, None, :], deg_threshold=deg_threshold,
                                           five_percent_heuristic=five_percent_heuristic)    # (J, J)
        low = torch.tril(pair, diagonal=-1)
        rotation_similarity_mat = (low | low.T | torch.eye(J, dtype=torch.bool, device=pair.device)).cpu()
        if five_percent_heuristic:
            th = int(T * 0.05)
            res = (torch.rad2deg(rot_angles).abs() >= deg_threshold).sum(dim=0)
            zero_motion = res <= th
        else:
            res = pair[J - 1, J - 2] if J > 1 else None
            zero_motion = torch.rad2deg((rot_angles ** 2).mean(dim=0)) <= deg_threshold
        prune_bones = zero_motion
        prune_bones[0] = False                                   # the (imaginary) root bone is never pruned

        joints = self.joints.detach().cpu().numpy()
        bones = self.bones
        new_joints, new_bones, merging_rules, joints_to_keep, rotations_to_keep, _, sibling_transfer_rules = \
            merge_joints(joints, bones, prune_bones.cpu().numpy(), rotation_similarity_mat.numpy(),
                         convert_merging_rules=False)
        rotations_to_keep = torch.tensor(rotations_to_keep)
        self.merging_rules = merging_rules
        self.joints_to_keep = joints_to_keep
        self.new_bones = new_bones

        dev = self.forward_warp.rot_mask.device
        self.forward_warp.set_rotation_mask(~prune_bones.to(dev))
        self.forward_warp.set_sibling_mask(torch.tensor(sibling_transfer_rules).to(dev))
        flat = [int(v) for v in self.flatten_merging_rules(merging_rules)]
        self.flat_merging_rules = torch.tensor(flat, dtype=torch.long, device=self.flat_merging_rules.device)
        self.sibling_merging_rules = torch.tensor(sibling_transfer_rules).to(self.sibling_merging_rules.device)
        object.__setattr__(self, '_pose_graph', None)            # masks changed: captured pose graphs are stale
        object.__setattr__(self, '_pose_graph_key', None)

        print(f"Frozen joints/weights: {int(prune_bones.sum())} of {len(prune_bones)} ")
        print(f"Joints kept: {[i for i, v in enumerate(~prune_bones) if v]}")
        print(f"Actually pruned joints: {len(joints) - len(new_joints)} of {len(joints)}")
        return joints, bones, new_joints, new_bones, prune_bones, merging_rules, rotations_to_keep, res

    # ------------------------------------------------------------------------------------------
    def _mlp_weights(self):
        fn, rn = self.feat_net, self.rgbnet
        lin = [fn[0], fn[2][0], fn[3][0], fn[4]]
        ws = []
        for l in lin:
            ws += [l.weight, l.bias]
        v0_w = rn.views_linears[0].weight
        if self.no_view_dir:       # (64, 128) -> (64, 155): zero view columns; differentiable, so autograd slices the gradient back
            v0_w = torch.cat([v0_w, v0_w.new_zeros(v0_w.shape[0], ops.FEAT_DIM + ops.PE_VIEW - v0_w.shape[1])], dim=1)
        ws += [self.densitynet.weight, self.densitynet.bias, rn.feature_linears.weight, rn.feature_linears.bias,
               v0_w, rn.views_linears[0].bias, rn.views_linears[2].weight, rn.views_linears[2].bias]
        return ws

    def _view_dirs(self, viewdirs, n_rays: int):
        """The (R,3) view directions the heads see: the caller's, or `frozen_view_dir` for every ray
        (lib/temporalpoints.py:507-512)."""
        if self.frozen_view_dir is None:
            return ops._f32(viewdirs)
        dev = self.joints.device
        fd = self._frozen_dirs
        if fd is None or fd.shape[0] < n_rays or fd.device != dev:
            fd = self.frozen_view_dir.to(dev).reshape(1, 3).expand(max(int(n_rays), 1), 3).contiguous()
            object.__setattr__(self, '_frozen_dirs', fd)
        return fd[:n_rays]

    def _merge_rules_i32(self):
        """int32 copy of flat_merging_rules for the kernel, None while the rules are the identity
        (re-derived only when the buffer is replaced, e.g. by simplify_skeleton or load_state_dict)."""
        r = self.flat_merging_rules
        key = (r.data_ptr(), r._version, r.device)
        if getattr(self, '_rules_key', None) != key:
            ident = bool((r == torch.arange(len(r), device=r.device)).all())
            self._rules_cache = None if ident else r.to(torch.int32).contiguous()
            self._rules_key = key
        return self._rules_cache

    @property
    def _last_weights(self):
        """The merged skinning weights of the last forward (lib/temporalpoints.py:555-556).  A no-grad render does not
        store them (the LBS kernel skips that (N,J) write unless something reads it): they are then evaluated on demand."""
        if self._last_weights_value is None and self.weights.is_cuda:
            with torch.no_grad():
                self._last_weights_value = self.get_weights()
        return self._last_weights_value

    @_last_weights.setter
    def _last_weights(self, w):
        self._last_weights_value = w

    def warp(self, t=None, rot_params=None, want_weights=None):
        """forward_warp stage: -> dict(xyz, ginv, weights, bbox, bone_Ts, global_t, joints_rel).
        want_weights (default: whenever autograd records, i.e. training, where the regularisers read `_last_weights`):
        also return the merged skinning weights."""
        if want_weights is None:
            want_weights = torch.is_grad_enabled()
        self._ensure_neighbourhood()
        t_embed = poc_fre(t, self.time_poc) if rot_params is None else None
        joints_rel = None
        fused = (rot_params is None and self.joints.is_cuda and self.forward_warp.fused_pose
                 and self.forward_warp._fused_tables(self.joints.device) is not None)
        if (not fused and self.graph_pose and rot_params is None and torch.is_grad_enabled() and self.joints.is_cuda
                and self.joints.requires_grad):
            bone_Ts, global_t, joints_rel = self._pose_graphed(t_embed)
        else:
            bone_Ts, global_t = self.forward_warp.pose(self.joints, t=t_embed, rot_params=rot_params)
        rules = self._merge_rules_i32()
        xyz, ginv, w, bbox = ops.lbs(self.weights, self.theta_weight, bone_Ts, global_t, self.canonical_pcd, rules=rules,
                                     eps=float(self.eps), want_weights=want_weights)
        self._last_weights = w                      # None: evaluated lazily by the property if someone asks
        if joints_rel is None:
            jh = torch.cat([self.joints, torch.ones((len(self.joints), 1), device=self.joints.device)], dim=-1)
            joints_rel = torch.bmm(bone_Ts, jh.unsqueeze(-1)).squeeze(-1)[:, :3]
        return dict(xyz=xyz, ginv=ginv, weights=w, bbox=bbox, bone_Ts=bone_Ts, global_t=global_t, joints_rel=joints_rel)

    def _pose_graphed(self, t_embed):
        """Pose chain through CUDA graphs (captured on first use, re-captured when the masks change)."""
        fw = self.forward_warp
        key = (id(fw.rot_mask), fw.rot_mask._version, id(fw.sibling_mask), fw.sibling_mask._version, self.joints.data_ptr())
        if self._pose_graph is None or self._pose_graph_key != key:
            chain = _PoseChain(fw)
            sample = (t_embed.detach().clone(), self.joints)
            graphed = torch.cuda.make_graphed_callables(chain, sample)
            object.__setattr__(self, '_pose_graph', graphed)
            object.__setattr__(self, '_pose_graph_key', key)
        bone_Ts, global_t, thetas, joints_rel = self._pose_graph(t_embed.detach(), self.joints)
        fw.prev_thetas, fw.prev_global_t = thetas, global_t
        return bone_Ts, global_t, joints_rel

    def build_grid(self, warped, query_radius=0.01):
        return ops.Grid(warped['xyz'], warped['bbox'], query_radius=query_radius, bbox_pad=query_radius,
                        cell_hint=1.5 * self._mmd_float)

    def forward(self, t, render_depth=False, render_kwargs=None, query_radius=0.01, render_weights=False, rot_params=None,
                render_pcd_direct=False, poses=None, Ks=None, cam_per_ray=None, calc_min_max=True, get_skeleton=False,
                warped=None, grid=None):
        """lib/temporalpoints.py:540-712.  `warped` / `grid` (optional, not in the reference) let a caller
        reuse the per-pose warp across ray chunks instead of re-warping for every chunk."""
        assert (t is None) ^ (rot_params is None)
        assert render_kwargs is not None
        assert calc_min_max, "the reference's callers always sample inside the warped-cloud bbox"
        if warped is None:
            warped = self.warp(t, rot_params, want_weights=True if render_weights else None)
        t_hat_pcd = warped['xyz']
        joints, bones = None, None
        pose_embedding = None
        if self.pose_embedding_dim > 0:
            delta_joint = (self.joints - warped['joints_rel']).clone().detach()
            pose_embedding = self.pose_embedding_net(poc_fre(delta_joint, self.pos_poc).view(1, -1))
        if get_skeleton:
            gt = warped['global_t'] if warped['global_t'] is not None else torch.zeros(3, device=t_hat_pcd.device)
            joints = project_point_to_image_plane(warped['joints_rel'] + gt, poses, Ks.to(torch.float32))
            bones = self.bones
            if self.joints_to_keep is not None:
                joints = joints[:, self.joints_to_keep]
                bones = self.new_bones

        rays_o, rays_d, viewdirs = render_kwargs['rays_o'], render_kwargs['rays_d'], render_kwargs['viewdirs']
        R = len(rays_o)
        bg = float(render_kwargs['bg'])
        dev = t_hat_pcd.device
        if grid is None:
            grid = self.build_grid(warped, query_radius)
        stepdist = float(render_kwargs['stepsize']) * float(self.voxel_size)
        smp = ops.sample_and_knn(grid, rays_o, rays_d, float(render_kwargs['near']), float(render_kwargs['far']), stepdist)
        self.last_counts = dict(R=R, candidates=smp.n_candidates, M=smp.M, N=len(t_hat_pcd))
        if smp.M == 0:     # NoPointsException path (lib/temporalpoints.py:598-609)
            return {
                'rgb_marched': torch.ones(R, 3, device=dev) * bg, 'rgb_marched_direct': torch.ones(R, 3, device=dev) * bg,
                'depth': torch.zeros(R, device=dev), 'weights': torch.ones(R, 3, device=dev) * bg, 't_hat_pcd': t_hat_pcd,
                'alphainv_last': None, 'grid': None, 'joints': joints, 'bones': bones,
            }
        interval = float(render_kwargs['stepsize']) * float(self.tineuvox.voxel_size_ratio)
        c = ops.AggConst(pts=smp.pts, nn_idx=smp.nn_idx, ray_id=smp.ray_id, viewdirs=self._view_dirs(viewdirs, R),
                         canonical_alpha=self.canonical_alpha.detach(), canonical_rgbs=self.canonical_rgbs.detach(),
                         direct_eps=self.direct_eps.detach(), mean_min_distance=self._mmd_float, eps=float(self.eps),
                         act_shift=float(self.tineuvox.act_shift), interval=interval, direct=True)
        if self.decoder in ("tc", "tc_fast") and not torch.is_grad_enabled():
            alpha, rgb, alpha_d, rgb_d, idw = ops.aggregate_tc(c, t_hat_pcd, warped['ginv'], self.canonical_feat, pose_embedding,
                                                               self._mlp_weights(), self._packed_decoder,
                                                               precision=1 if self.decoder == "tc" else 0)
        elif self.decoder_train == "tc" and torch.is_grad_enabled():
            alpha, rgb, alpha_d, rgb_d, idw = ops.aggregate_tc_train(c, t_hat_pcd, warped['ginv'], self.canonical_feat,
                                                                     pose_embedding, self._mlp_weights(), self._packed_decoder)
        else:
            alpha, rgb, alpha_d, rgb_d, idw = ops.aggregate(c, t_hat_pcd, warped['ginv'], self.canonical_feat, pose_embedding,
                                                            self._mlp_weights())
        extra = None
        if render_weights:
            # lib/temporalpoints.py:517-519,690-701: per-sample LBS weights -> one colour per active bone
            lw = warped['weights'] if warped.get('weights') is not None else self.get_weights()
            lw = lw.detach()
            mask = lw.sum(dim=0) > 0
            cols = torch.tensor(hls_palette(int(mask.sum())), dtype=torch.float32)
            gen = torch.Generator().manual_seed(0)
            cols = cols[torch.randperm(cols.shape[0], generator=gen)].to(dev)
            point_cols = lw[:, mask] @ cols                                   # (N,3) colour of every point
            extra = (point_cols[smp.nn_idx.long()] * idw.unsqueeze(-1)).sum(1)
        rgb_m, last, depth, w_img = ops.composite(alpha, rgb, smp.step_id, smp.ray_start, R, self.fast_color_thres, bg,
                                                  extra=extra, want_depth=render_depth)
        rgb_md, last_d, _, _ = ops.composite(alpha_d, rgb_d, smp.step_id, smp.ray_start, R, self.fast_color_thres, bg,
                                             want_depth=False)
        ret = {'t_hat_pcd': t_hat_pcd, 'rgb_marched': rgb_m, 'alphainv_last': last, 'alphainv_last_direct': last_d,
               'grid': None, 'rgb_marched_direct': rgb_md, 'joints': joints, 'bones': bones}
        if render_depth:
            ret['depth'] = depth
        if render_weights:
            ret['weights'] = w_img
        return ret

    # -- regulariser losses (lib/temporalpoints.py:714-800) ----------------------------------------
    def get_neighbour_weight_tv_loss(self):
        diff = self._last_weights[:, None, :] - self._last_weights[self.nn_i, :]
        return torch.abs(diff).mean()

    def get_weight_sparsity_loss(self):
        w, e = self._last_weights, self.eps.to(self._last_weights.device)
        return -(w * torch.log(w + e) + (1 - w) * torch.log(1 - w + e)).mean()

    def get_arap_loss(self, warped_pcd, c=0.03):
        d = torch.sqrt((warped_pcd[:, None, :] - warped_pcd[self.nn_i, :]).pow(2).sum(-1) + self.eps.to(warped_pcd.device))
        return (self.nn_distance - d).abs().sum()

    def get_joint_arap_loss(self):
        bam = self.bone_arap_mask.to(self.joints.device)
        jd = (self.joints[bam][0::2, :] - self.joints[bam][1::2, :])
        return ((self.og_joint_distance - jd) ** 2).sum()

    def get_joint_chamfer_loss(self):
        _, c2 = self.get_chamfer_loss(self.skeleton_pcd, self.joints, c=None, get_raw=True)
        return c2.sum()

    @staticmethod
    def _rho(x, c):
        return (2 * (x / c) ** 2) / ((x / c) ** 2 + 4)

    @staticmethod
    def _nn1(query, target):
        """argKmin(K=1) through the grid (replaces the KeOps reduction of lib/temporalpoints.py:747-751)."""
        tgt = target.detach().contiguous()
        lo, hi = tgt.min(0)[0], tgt.max(0)[0]
        vol = float((hi - lo).clamp_min(1e-4).prod())
        cell = max((vol / max(len(tgt), 1)) ** (1 / 3), 1e-4)
        g = ops.Grid(tgt, torch.cat([lo, hi]), query_radius=0.01, bbox_pad=0.01, cell_hint=cell)
        idx, _ = g.knn(query.detach().contiguous(), 1)      # checks the grid's overflow flag, grows the table if needed
        return idx.long()

    def get_chamfer_loss(self, pcd1, pcd2, N=None, M=None, c=0.03, get_raw=False):
        if N is not None:
            pcd1 = pcd1[torch.randint(0, pcd1.shape[0], (N,), device=pcd1.device)]
        if M is not None:
            pcd2 = pcd2[torch.randint(0, pcd2.shape[0], (M,), device=pcd2.device)]
        nn_i1 = self._nn1(pcd1, pcd2)
        nn_i2 = self._nn1(pcd2, pcd1)
        d1 = ((pcd1[:, None, :] - pcd2[nn_i1, :]) ** 2).sum(-1)
        d2 = ((pcd2[:, None, :] - pcd1[nn_i2, :]) ** 2).sum(-1)
        if get_raw:
            return d1, d2
        if c is None:
            return d1.mean() + d2.mean()
        return self._rho(d1, c).mean() + self._rho(d2, c).mean()

    def get_batch_chamfer_loss(self, pcd1, pcd2, N=None, M=None):
        """lib/temporalpoints.py:765-795 (2-D chamfer of the projected cloud against mask pixels, run.py:659-690):
        (B, N, D) vs (B, M, D); gradients flow through the gathered coordinates, the indices come from
        `apn_nn1_batched`."""
        assert len(pcd1) == len(pcd2)
        if N is not None:
            pcd1 = pcd1[:, torch.randint(0, pcd1.shape[1], (N,), device=pcd1.device)]
        if M is not None:
            pcd2 = pcd2[:, torch.randint(0, pcd2.shape[1], (M,), device=pcd1.device)]
        D = pcd1.shape[-1]
        i12 = ops.nn1_batched(pcd1, pcd2)[..., None].expand(-1, -1, D)
        i21 = ops.nn1_batched(pcd2, pcd1)[..., None].expand(-1, -1, D)
        d1 = (pcd1 - torch.gather(pcd2, 1, i12)).pow(2)
        d2 = (pcd2 - torch.gather(pcd1, 1, i21)).pow(2)
        return d1.sum(-1).mean() + d2.sum(-1).mean()

    def get_transformation_regularisation_loss(self, d=0.0873):
        t = self.forward_warp.prev_global_t.abs()
        thetas = self.forward_warp.prev_thetas.abs()
        return (torch.abs(t).sum() + thetas.sum()) / len(thetas + 1)
