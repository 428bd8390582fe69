"""PointWarper / TransformNet with the reference's call surface (lib/pointwarper.py), B200 path.

Split of work:
  * pose -> per-bone rigid transforms: TransformNet MLP (lib/pointwarper.py:5-37), 4-parameter
    Rodrigues (:118-143) and the kinematic chain product (:145-193) act on J <= 128 joints; they stay
    PyTorch autograd on the device (a few KB of data, exact fp32).
  * the O(N*J) part — blend of the J transforms by the skinning weights, point transform, global
    translation (:241-266) — is the fused sm_100a kernel csrc/lbs.cu, which also emits the inverse
    frames and the cloud bbox that TemporalPoints needs next.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class TransformNet(torch.nn.Module):
    """lib/pointwarper.py:5-37: MLP t_dim -> 256 x (num_layers-1) -> components*params, last layer bias-free."""

    def __init__(self, input_dim, num_components, num_params_per_component, num_layers=3, hidden_dim=256):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.num_components = num_components
        self.num_params_per_component = num_params_per_component
        self.out_dim = num_components * num_params_per_component
        self.register_buffer('rotation_switch_mask', torch.arange(0, num_components).long())
        layers = []
        for i in range(num_layers - 1):
            layers.append(torch.nn.Linear(input_dim if i == 0 else self.hidden_dim, self.hidden_dim))
            layers.append(torch.nn.ReLU())
        layers.append(torch.nn.Linear(self.hidden_dim, self.out_dim, bias=False))
        self.net = torch.nn.Sequential(*layers)

    def forward(self, x):
        b, _ = x.shape
        out = self.net(x)
        if b > 1:
            return out.reshape(b, self.num_components, self.num_params_per_component)
        return out.reshape(self.num_components, self.num_params_per_component)


def rodrigues(rvec: torch.Tensor):
    """lib/pointwarper.py:118-143.  (J,3): angle = |v| ; (J,4): (unnormalised axis, angle)."""
    if rvec.shape[-1] == 3:
        theta = torch.sqrt(1e-5 + torch.sum(rvec ** 2, dim=1))
        rvec = rvec / theta[:, None]
    elif rvec.shape[-1] == 4:
        theta = rvec[:, -1]
        rvec = rvec[:, :3]
        rvec = rvec / torch.sqrt(1e-5 + torch.sum(rvec ** 2, dim=1))[:, None]
    else:
        raise ValueError()
    c, s = torch.cos(theta), torch.sin(theta)
    x, y, z = rvec[:, 0], rvec[:, 1], rvec[:, 2]
    R = torch.stack((
        x ** 2 + (1. - x ** 2) * c, x * y * (1. - c) - z * s, x * z * (1. - c) + y * s,
        x * y * (1. - c) + z * s, y ** 2 + (1. - y ** 2) * c, y * z * (1. - c) - x * s,
        x * z * (1. - c) - y * s, y * z * (1. - c) + x * s, z ** 2 + (1. - z ** 2) * c), dim=1).view(-1, 3, 3)
    return R, theta


class PointWarper(torch.nn.Module):
    def __init__(self, t_dim, canonical_pcd, joints, bones, num_layers=5, over_parameterized_rot=True):
        super().__init__()
        self.t_dim = t_dim
        self.params_per_compoent = 4
        self.canonical_pcd = canonical_pcd
        self.num_layers = num_layers
        self.over_parameterized_rot = over_parameterized_rot
        self.init_tree(joints, bones)
        self.transform_net = TransformNet(t_dim, len(joints) + 1, self.params_per_compoent, num_layers=self.num_layers)
        self.register_buffer('rot_mask', torch.zeros(len(joints), dtype=torch.bool))
        self.register_buffer('sibling_mask', torch.arange(0, len(joints)).long())
        self.prev_params = self.prev_thetas = self.prev_global_t = None
        self.fused_pose = True       # one-launch pose kernel (csrc/pose.cu) when the tree / MLP shape allows it

    # -- kinematic tree tables (lib/pointwarper.py:95-116, old=False branch) --------------------
    def init_tree(self, joints, bones, old=False):
        self.bones = bones
        self.parent_joint = {int(b[1]): int(b[0]) for b in bones}
        self.child_joints = {k: [] for k in range(len(joints))}
        for k, p in self.parent_joint.items():
            self.child_joints[p].append(k)
        chains = [[0]]
        for i in range(len(bones)):
            j, inds = i + 1, []
            while j >= 0:
                inds.append(j)
                j = self.parent_joint.get(j, -1)
            chains.append(inds[::-1])
        depth = max(len(c) for c in chains)
        table = np.full((len(chains), depth), -1, dtype=np.int64)
        for i, c in enumerate(chains):
            table[i, :len(c)] = c
        self.parent_indices = torch.from_numpy(table)
        self.parent_joint_ex = torch.tensor([self.parent_joint.get(i, 0) for i in range(len(chains))], dtype=torch.long)

    def _tables(self, device):
        if self.parent_indices.device != device:
            self.parent_indices = self.parent_indices.to(device)
            self.parent_joint_ex = self.parent_joint_ex.to(device)
        return self.parent_indices, self.parent_joint_ex

    @classmethod
    def matrix_chain_product(cls, chain: torch.Tensor) -> torch.Tensor:
        """Product of chain[:, 0] @ chain[:, 1] @ ... by recursive halving (lib/pointwarper.py:145-153)."""
        n = chain.shape[1]
        if n == 1:
            return chain
        return cls.matrix_chain_product(chain[:, :n // 2]) @ cls.matrix_chain_product(chain[:, n // 2:])

    def calc_rec_abs_T_fast(self, R_t: torch.Tensor, joints: torch.Tensor) -> torch.Tensor:
        """lib/pointwarper.py:156-193: node i rotates by R_i about its PARENT joint's position
        (the root about itself); absolute transform = product along the root -> i chain."""
        pi, pj = self._tables(joints.device)
        J = R_t.shape[0]
        pivot = joints[pj]
        top = torch.cat((R_t, pivot[..., None] - R_t @ pivot[..., None]), -1)                      # (J,3,4)
        eye = torch.eye(4, device=joints.device, dtype=joints.dtype)     # built on the device: CUDA-graph capturable
        M = torch.cat((top, eye[3:4].expand(J, 1, 4)), -2)
        M = torch.cat((eye[None], M), 0)                                                            # slot 0 = identity pad
        return self.matrix_chain_product(M[pi + 1])[:, 0]

    def get_thetas(self, ts_embed):
        params = self.transform_net(ts_embed)
        rot_params = params[:, :-1, :3]
        shape = rot_params.shape[:2]
        _, thetas = rodrigues(rot_params.reshape(shape[0] * shape[1], 3))
        return thetas.reshape(shape)

    def set_rotation_mask(self, rotations_to_keep):
        mask = ~rotations_to_keep
        if self.rot_mask is not None:
            mask = torch.logical_or(mask, self.rot_mask)
        self.rot_mask = mask

    def set_sibling_mask(self, sibling_mask):
        self.sibling_mask = sibling_mask.long()

    Rodrigues = staticmethod(rodrigues)

    # -- fused pose chain (csrc/pose.cu) -----------------------------------------------------
    def _fused_tables(self, device):
        """Device tables for the one-launch pose kernel, or None when the tree / MLP shape is outside what it covers
        (then the PyTorch ops below are used: same maths, more launches)."""
        key = (device, id(self.rot_mask), self.rot_mask._version, id(self.sibling_mask), self.sibling_mask._version)
        if getattr(self, '_fused_key', None) != key:
            self._fused_key, self._fused = key, None
            J = len(self.parent_joint_ex)
            parent_node = [self.parent_joint.get(i, -1) for i in range(J)]
            lin = [m for m in self.transform_net.net if isinstance(m, torch.nn.Linear)]
            ok = (J <= 128 and all(p < i for i, p in enumerate(parent_node)) and len(lin) == 5
                  and lin[0].in_features <= 64 and all(l.out_features == 256 for l in lin[:4]) and lin[4].bias is None
                  and lin[4].out_features == (J + 1) * 4)
            if ok:
                i32 = dict(dtype=torch.int32, device=device)
                rm = self.rot_mask
                self._fused = ops.PoseTables(
                    parent_node=torch.tensor(parent_node, **i32), pivot=self.parent_joint_ex.to(**i32),
                    sibling=self.sibling_mask.to(**i32),
                    rot_mask=rm.to(device=device, dtype=torch.uint8) if rm is not None else None)
        return self._fused

    def _mlp_params(self):
        lin = [m for m in self.transform_net.net if isinstance(m, torch.nn.Linear)]
        out = []
        for l in lin[:4]:
            out += [l.weight, l.bias]
        return out + [lin[4].weight]

    # -- pose -> bone transforms -------------------------------------------------------------
    def pose(self, joints, t=None, rot_params=None, global_t=None):
        """-> bone_Ts (J,4,4), global_t (3) or None.  Sets prev_params / prev_thetas / prev_global_t like
        lib/pointwarper.py:217-228."""
        assert (t is None) ^ (rot_params is None)
        if rot_params is None and joints.is_cuda and self.fused_pose:
            tb = self._fused_tables(joints.device)
            if tb is not None:
                bone_Ts, global_t, thetas = ops.pose_chain(tb, t, joints, self._mlp_params())
                self.prev_params = None
                self.prev_thetas, self.prev_global_t = thetas, global_t
                return bone_Ts, global_t
        if rot_params is None:
            params = self.transform_net(t.unsqueeze(0))
            self.prev_params = params
            global_t = params[-1, :3]
            R_t, self.prev_thetas = rodrigues(params[:len(joints), :])
            self.prev_global_t = global_t
        else:
            R_t, self.prev_thetas = rodrigues(rot_params)
        R_t = R_t[self.sibling_mask]
        if self.rot_mask is not None:      # no host sync: select instead of masked assignment
            eye = torch.eye(3, device=R_t.device, dtype=R_t.dtype)
            R_t = torch.where(self.rot_mask[:, None, None], eye[None], R_t)
        return self.calc_rec_abs_T_fast(R_t, joints), global_t

    def forward(self, weights, joints, t=None, rot_params=None, global_t=None, get_frames=False, avg_procrustes=False,
                get_skeleton=False):
        """Same contract as lib/pointwarper.py:213-278; `weights` are final (soft-maxed, merged) (N,J)."""
        if avg_procrustes:
            raise NotImplementedError("avg_procrustes is never enabled by the reference's callers "
                                      "(lib/temporalpoints.py:558-563)")
        bone_Ts, global_t = self.pose(joints, t=t, rot_params=rot_params, global_t=global_t)
        xyz = self.canonical_pcd
        if xyz.device != joints.device:
            xyz = self.canonical_pcd = xyz.to(joints.device)
        res = ops.lbs(weights, None, bone_Ts, global_t, xyz, want_frames=get_frames, want_weights=False)
        jointsh = torch.cat([joints, torch.ones((len(joints), 1), device=joints.device, dtype=joints.dtype)], dim=-1)
        joints_warped_rel = torch.bmm(bone_Ts, jointsh.unsqueeze(-1)).squeeze(-1)[:, :3]
        out = [res[0], joints_warped_rel]
        if get_frames:
            out.append(res[4])
        if get_skeleton:
            gt = global_t if global_t is not None else torch.zeros(3, device=joints.device)
            out.append(joints_warped_rel + gt)
            out.append(self.bones)
        return out
