"""TEST INFRASTRUCTURE ONLY — torch-CPU fp32 restatement of the PCD render path.

Never imported by the product package; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs use it (as the checker / the CPU arm).

Follows, function by function:
  poc_fre                      lib/tineuvox.py:872-878
  TransformNet                 lib/pointwarper.py:5-37
  Rodrigues (4-param)          lib/pointwarper.py:118-143
  init_tree / chain product    lib/pointwarper.py:70-116,145-193
  PointWarper.forward          lib/pointwarper.py:213-278
  get_weights                  lib/temporalpoints.py:401-414
  sample_ray                   lib/temporalpoints.py:373-399
  aggregate_pts                lib/temporalpoints.py:416-521
  TemporalPoints.forward       lib/temporalpoints.py:540-712
  RGBNet.forward               lib/tineuvox.py:65-88
  Raw2Alpha / Alphas2Weights   lib/tineuvox.py:627-670
  KeOps Kmin_argKmin           third party (pykeops, unpinned in requirements.txt:15): exact K
                               smallest squared distances, ascending; restated as blocked brute
                               force with the contract  d2 = (dx*dx + dy*dy) + dz*dz  in fp32, ties
                               broken by the lower point index.
  torch_scatter.segment_coo    third party (unpinned, requirements.txt:2): index_add_.

Pinning: the reference ships no tests or golden vectors for this path, so
this oracle is pinned against tensors produced by the reference's own Python executed in
the authoring container under third-party shims (oracle/ref_harness.py ->
tests/golden/ref_tiny.pt; tests/test_oracle_vs_reference.py).
"""
from __future__ import annotations

import colorsys
import math
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as Fnn

from . import dvgo_ops

F32 = torch.float32


# ----------------------------------------------------------------------------------
# small pieces
# ----------------------------------------------------------------------------------
def poc_fre(x: torch.Tensor, poc: torch.Tensor) -> torch.Tensor:
    emb = (x.unsqueeze(-1) * poc).flatten(-2)
    return torch.cat([x, emb.sin(), emb.cos()], -1)


def rodrigues4(p: torch.Tensor):
    theta = p[:, -1]
    r = p[:, :3]
    r = r / torch.sqrt(1e-5 + torch.sum(r ** 2, dim=1))[:, None]
    c, s = torch.cos(theta), torch.sin(theta)
    x, y, z = r[:, 0], r[:, 1], r[:, 2]
    R = torch.stack((
        x ** 2 + (1. - x ** 2) * c, x * y * (1. - c) - z * s, x * z * (1. - c) + y * s,
        x * y * (1. - c) + z * s, y ** 2 + (1. - y ** 2) * c, y * z * (1. - c) - x * s,
        x * z * (1. - c) - y * s, y * z * (1. - c) + x * s, z ** 2 + (1. - z ** 2) * c), dim=1).view(-1, 3, 3)
    return R, theta


def build_tree(bones, n_joints: int):
    """lib/pointwarper.py:95-116 (old=False branch)."""
    parent = {int(b[1]): int(b[0]) for b in bones}
    chains = [[0]]
    for i in range(len(bones)):
        j, inds = i + 1, []
        while j >= 0:
            inds.append(j)
            j = parent.get(j, -1)
        chains.append(inds[::-1])
    depth = max(len(c) for c in chains)
    pi = torch.full((len(chains), depth), -1, dtype=torch.long)
    for i, c in enumerate(chains):
        pi[i, :len(c)] = torch.tensor(c)
    pj = torch.tensor([parent.get(i, 0) for i in range(len(chains))], dtype=torch.long)
    return pi, pj


def _chain_product(m: torch.Tensor) -> torch.Tensor:
    n = m.shape[1]
    if n == 1:
        return m
    return _chain_product(m[:, :n // 2]) @ _chain_product(m[:, n // 2:])


def bone_transforms(R_t: torch.Tensor, joints: torch.Tensor, parent_indices, parent_joint_ex) -> torch.Tensor:
    """lib/pointwarper.py:156-193: M_i = [R_i | p - R_i p], p = joints[parent(i)]; product root -> i."""
    J = R_t.shape[0]
    piv = joints[parent_joint_ex]
    hom = torch.tensor([0., 0., 0., 1.])
    top = torch.cat((R_t, piv[..., None] + R_t @ -piv[..., None]), -1)
    M = torch.cat((top, hom[None, None].repeat(J, 1, 1)), -2)
    M = torch.cat((torch.eye(4)[None], M), 0)
    return _chain_product(M[parent_indices + 1])[:, 0]


def get_weights(raw: torch.Tensor, theta_weight: torch.Tensor, eps: float, rules: torch.Tensor) -> torch.Tensor:
    """lib/temporalpoints.py:401-414; the (J,J,J) bmm is the index-add  out[:, rules[j]] += w[:, j]."""
    th = torch.max(torch.tensor(eps), theta_weight)
    w = torch.softmax(raw / th, dim=-1)
    J = w.shape[1]
    onehot = torch.zeros(J, J, dtype=w.dtype)
    onehot[torch.arange(J), rules.long()] = 1.0
    return w @ onehot


def knn_bruteforce(q: torch.Tensor, pts: torch.Tensor, K: int, block_bytes: int = 1 << 30):
    """Exact K smallest of d2 = (dx*dx + dy*dy) + dz*dz, ascending by (d2, index)."""
    S, N = len(q), len(pts)
    out_d = torch.empty(S, K, dtype=F32)
    out_i = torch.empty(S, K, dtype=torch.int64)
    px, py, pz = pts[:, 0][None], pts[:, 1][None], pts[:, 2][None]
    idx = torch.arange(N, dtype=torch.int64)[None]
    blk = max(1, block_bytes // (N * 24))
    for s in range(0, S, blk):
        qq = q[s:s + blk]
        dx = qq[:, 0:1] - px
        dy = qq[:, 1:2] - py
        dz = qq[:, 2:3] - pz
        d2 = (dx * dx + dy * dy) + dz * dz
        key = (d2.view(torch.int32).to(torch.int64) << 32) | idx
        kk = torch.topk(key, K, dim=1, largest=False, sorted=True).values
        out_i[s:s + blk] = kk & 0xFFFFFFFF
        out_d[s:s + blk] = (kk >> 32).to(torch.int32).view(F32)
    return out_d, out_i


class Raw2Alpha(torch.autograd.Function):
    @staticmethod
    def forward(ctx, density, shift, interval):
        e, alpha = dvgo_ops.raw2alpha(density, shift, interval)
        ctx.save_for_backward(e)
        ctx.interval = interval
        return alpha

    @staticmethod
    def backward(ctx, g):
        return dvgo_ops.raw2alpha_backward(ctx.saved_tensors[0], g.contiguous(), ctx.interval), None, None


class Alphas2Weights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        w, T, last, i0, i1 = dvgo_ops.alpha2weight(alpha, ray_id, N)
        ctx.save_for_backward(alpha, w, T, last, i0, i1)
        ctx.n = N
        return w, last

    @staticmethod
    def backward(ctx, gw, gl):
        alpha, w, T, last, i0, i1 = ctx.saved_tensors
        return dvgo_ops.alpha2weight_backward(alpha, w, T, last, i0, i1, ctx.n, gw.contiguous(), gl.contiguous()), None, None


def hls_palette(n: int):
    """seaborn.color_palette('hls', n) restated: hues linspace(0,1,n+1)[:-1]+0.01, l=.6, s=.65."""
    hues = np.linspace(0, 1, int(n) + 1)[:-1] + 0.01
    hues %= 1
    return [colorsys.hls_to_rgb(float(h), 0.6, 0.65) for h in hues]


def leaky(x):
    return Fnn.leaky_relu(x, 0.01)


# ----------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------
class OraclePath:
    """Functional restatement; `state` uses the reference's state_dict key names."""

    def __init__(self, state: Dict[str, torch.Tensor], canonical_pcd, bones, *, stepsize, voxel_size,
                 fast_color_thres, act_shift, voxel_size_ratio, mean_min_distance=None, eps=1e-6,
                 neighbours=8, pose_embedding_dim=0, feat_depth=4, posbase_pe=10, viewbase_pe=4, timebase_pe=8,
                 no_view_dir=False, frozen_view_dir=False):
        """no_view_dir: the RGB head takes no view columns (lib/tineuvox.py:112-113, lib/temporalpoints.py:504-505);
        frozen_view_dir: every ray uses state['viewdirs_emb'] (lib/temporalpoints.py:157-159,507-508)."""
        self.no_view_dir, self.frozen_view_dir = no_view_dir, frozen_view_dir
        self.s = state
        self.pcd = canonical_pcd.to(F32)
        self.bones = bones
        self.stepsize, self.voxel_size = stepsize, float(voxel_size)
        self.thres = fast_color_thres
        self.act_shift, self.vsr = float(act_shift), float(voxel_size_ratio)
        self.eps, self.K = eps, neighbours
        self.pose_dim, self.feat_depth = pose_embedding_dim, feat_depth
        self.pos_poc = torch.tensor([2.0 ** i for i in range(posbase_pe)])
        self.view_poc = torch.tensor([2.0 ** i for i in range(viewbase_pe)])
        self.time_poc = torch.tensor([2.0 ** i for i in range(timebase_pe)])
        J = state['joints'].shape[0]
        self.parent_indices, self.parent_joint_ex = build_tree(bones, J)
        self.nn_i = self.nn_distance = None
        if mean_min_distance is None:
            self.neighbourhood()
            mean_min_distance = self.nn_distance[:, 1].mean()
        self.mean_min_distance = torch.as_tensor(mean_min_distance, dtype=F32)
        self.trace: Dict[str, torch.Tensor] = {}

    def neighbourhood(self):
        """lib/temporalpoints.py:104-111: static 8-NN of the canonical cloud (self included)."""
        if self.nn_i is None:
            _, self.nn_i = knn_bruteforce(self.pcd, self.pcd, self.K)
            self.nn_distance = torch.sqrt(((self.pcd[:, None] - self.pcd[self.nn_i]) ** 2).sum(-1) + self.eps)
        return self.nn_i, self.nn_distance

    # -- regulariser losses (lib/temporalpoints.py:714-733,797-800) -----------------------------
    def arap_loss(self, warped_pcd):
        nn_i, nn_d = self.neighbourhood()
        d = torch.sqrt((warped_pcd[:, None, :] - warped_pcd[nn_i, :]).pow(2).sum(-1) + self.eps)
        return (nn_d - d).abs().sum()

    def weight_tv_loss(self, weights):
        nn_i, _ = self.neighbourhood()
        return torch.abs(weights[:, None, :] - weights[nn_i, :]).mean()

    def sparsity_loss(self, weights):
        return -(weights * torch.log(weights + self.eps) + (1 - weights) * torch.log(1 - weights + self.eps)).mean()

    @staticmethod
    def transformation_reg_loss(global_t, thetas):
        return (global_t.abs().sum() + thetas.abs().sum()) / len(thetas)

    def joint_chamfer_loss(self, skeleton_pcd):
        """lib/temporalpoints.py:731-733,738-763 (get_raw=True, second direction): every joint to its nearest
        skeleton point."""
        joints = self.s['joints']
        _, i2 = knn_bruteforce(joints.detach(), skeleton_pcd, 1)
        return ((joints[:, None, :] - skeleton_pcd[i2, :]) ** 2).sum(-1).sum()

    @staticmethod
    def batch_chamfer_loss(pcd1, pcd2):
        """lib/temporalpoints.py:765-795 without the random sub-sampling: (B, N, D) vs (B, M, D), D = 2 or 3; nearest
        neighbour per batch item by d2 = (dx*dx + dy*dy) [+ dz*dz], ties -> lowest index; gradients flow through the
        gathered coordinates only."""
        def nn1(a, b):                                            # (B, Na, D), (B, Nb, D) -> (B, Na) int64
            diff = a.detach()[:, :, None, :] - b.detach()[:, None, :, :]
            d2 = diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]
            if a.shape[-1] == 3:
                d2 = d2 + diff[..., 2] * diff[..., 2]
            key = (d2.contiguous().view(torch.int32).to(torch.int64) << 32) | torch.arange(b.shape[1], dtype=torch.int64)
            return key.min(-1).values & 0xFFFFFFFF
        D = pcd1.shape[-1]
        i12 = nn1(pcd1, pcd2)[..., None].expand(-1, -1, D)
        i21 = nn1(pcd2, pcd1)[..., None].expand(-1, -1, D)
        d1 = (pcd1 - torch.gather(pcd2, 1, i12)).pow(2)
        d2 = (pcd2 - torch.gather(pcd1, 1, i21)).pow(2)
        return d1.sum(-1).mean() + d2.sum(-1).mean()

    # -- sub-networks ---------------------------------------------------------------
    def transform_net(self, x):
        s = self.s
        for i in (0, 2, 4, 6):
            x = torch.relu(Fnn.linear(x, s[f'forward_warp.transform_net.net.{i}.weight'],
                                      s[f'forward_warp.transform_net.net.{i}.bias']))
        return Fnn.linear(x, s['forward_warp.transform_net.net.8.weight'])

    def feat_net(self, x):
        s = self.s
        names = ['feat_net.0'] + [f'feat_net.{i}.0' for i in range(2, self.feat_depth)] + [f'feat_net.{self.feat_depth}']
        for n in names:
            x = leaky(Fnn.linear(x, s[n + '.weight'], s[n + '.bias']))
        return x

    def pose_embedding_net(self, x):
        s = self.s
        names = ['pose_embedding_net.0'] + [f'pose_embedding_net.{i}.0' for i in range(2, self.feat_depth)] + \
                [f'pose_embedding_net.{self.feat_depth}']
        for n in names:
            x = leaky(Fnn.linear(x, s[n + '.weight'], s[n + '.bias']))
        return x

    def rgbnet(self, h, views):
        s = self.s
        f = Fnn.linear(h, s['rgbnet.feature_linears.weight'], s['rgbnet.feature_linears.bias'])
        x = f if views is None else torch.cat([f, views], -1)           # lib/tineuvox.py:80-88
        x = torch.relu(Fnn.linear(x, s['rgbnet.views_linears.0.weight'], s['rgbnet.views_linears.0.bias']))
        return Fnn.linear(x, s['rgbnet.views_linears.2.weight'], s['rgbnet.views_linears.2.bias'])

    # -- warp -----------------------------------------------------------------------
    def warp(self, t=None, rot_params=None):
        s = self.s
        J = s['joints'].shape[0]
        rules = s.get('flat_merging_rules', torch.arange(J))
        weights = get_weights(s['weights'], s['theta_weight'], self.eps, rules)
        if rot_params is None:
            t_embed = poc_fre(t, self.time_poc)
            params = self.transform_net(t_embed.unsqueeze(0)).reshape(J + 1, 4)
            global_t = params[-1, :3]
            R_t, thetas = rodrigues4(params[:J])
        else:
            R_t, thetas = rodrigues4(rot_params)
            global_t = torch.zeros(3)
        sib = s.get('forward_warp.sibling_mask', torch.arange(J))
        R_t = R_t[sib.long()]
        rot_mask = s.get('forward_warp.rot_mask', torch.zeros(J, dtype=torch.bool))
        if rot_mask.any():
            R_t = torch.where(rot_mask[:, None, None], torch.eye(3)[None], R_t)
        bone_Ts = bone_transforms(R_t, s['joints'], self.parent_indices, self.parent_joint_ex)
        G = (bone_Ts * weights[:, :, None, None]).sum(dim=1)
        xyzh = torch.cat([self.pcd, torch.ones(len(self.pcd), 1)], -1)
        xyz = torch.bmm(G, xyzh.unsqueeze(-1)).squeeze(-1)[:, :3] + global_t
        jh = torch.cat([s['joints'], torch.ones(J, 1)], -1)
        joints_rel = torch.bmm(bone_Ts, jh.unsqueeze(-1)).squeeze(-1)[:, :3]
        return dict(weights=weights, bone_Ts=bone_Ts, G=G, xyz=xyz.contiguous(), joints_rel=joints_rel,
                    global_t=global_t, thetas=thetas, joints_warped=joints_rel + global_t)

    # -- sampling + knn ---------------------------------------------------------------
    def sample_and_knn(self, xyz, rays_o, rays_d, near, far, stepsize, query_radius=0.01):
        xyz_d = xyz.detach()
        lo = xyz_d.min(0)[0] - query_radius
        hi = xyz_d.max(0)[0] + query_radius
        stepdist = stepsize * self.voxel_size
        pts, mask_out, ray_id, step_id, n_steps, _, _ = dvgo_ops.sample_pts_on_rays(
            rays_o.contiguous(), rays_d.contiguous(), lo, hi, near, far, stepdist)
        inb = ~mask_out
        pts, ray_id, step_id = pts[inb], ray_id[inb], step_id[inb]
        T, S = len(mask_out), len(pts)
        if S == 0:
            return None
        d2, s_i = knn_bruteforce(pts, xyz_d, self.K)
        keep = torch.where(d2[:, -1] <= query_radius)[0]
        out = dict(bbox_min=lo, bbox_max=hi, T=T, S=S, pts=pts[keep], ray_id=ray_id[keep], step_id=step_id[keep],
                   s_i=s_i[keep], d2_all=d2, s_i_all=s_i, keep=keep, pts_all=pts)
        return out if len(keep) else None

    # -- aggregate (lib/temporalpoints.py:452-521) -------------------------------------
    def aggregate(self, xyz, Ginv, smp, viewdirs, stepsize, pose_embedding=None, merged_weights=None):
        s = self.s
        K = self.K
        s_i, pts, ray_id = smp['s_i'], smp['pts'], smp['ray_id']
        rel_p = pts[:, None, :] - xyz[s_i, :]
        to_nn = (rel_p ** 2).sum(-1)
        feats = s['canonical_feat'][s_i, :]
        frames = Ginv[s_i]
        # direct branch (always on: lib/temporalpoints.py:592)
        sig = self.mean_min_distance * torch.max(s['direct_eps'], torch.tensor(0.))
        w_direct = torch.exp(-(to_nn ** 2) / (2 * (sig[s_i]) ** 2 + 1e-12))
        w_dd = (torch.tensor(1. / K) * w_direct).unsqueeze(-1)
        w_direct = (w_direct / (w_direct.sum(dim=-1) + 1e-12)[:, None]).unsqueeze(-1)
        a_k = s['canonical_alpha'].clip(0, 1)[s_i].unsqueeze(-1)
        c_k = s['canonical_rgbs'].clip(0, 1)[s_i, :]
        rgbs_direct = (w_direct * c_k).sum(dim=1)
        alpha_direct = (w_dd * a_k).sum(dim=1).squeeze(-1)
        # point-nerf branch
        w = 1 / (to_nn + self.eps)
        w = (w / w.sum(dim=-1)[:, None]).unsqueeze(-1)
        rel_c = torch.bmm(frames[..., :3, :3].reshape(-1, 3, 3), rel_p.reshape(-1, 3).unsqueeze(-1)).squeeze(-1)
        emb = poc_fre(rel_c, self.pos_poc)
        x = [emb, feats.reshape(-1, feats.shape[-1])]
        if pose_embedding is not None:
            x.append(pose_embedding.expand(len(emb), -1))
        x = torch.cat(x, -1)
        out = self.feat_net(x).reshape(len(s_i), K, -1)
        h = (out * w).sum(dim=1)
        density = Fnn.linear(h, s['densitynet.weight'], s['densitynet.bias']).squeeze(-1)
        interval = stepsize * self.vsr
        alpha = Raw2Alpha.apply(density.flatten(), self.act_shift, interval)
        if self.no_view_dir:                                             # lib/temporalpoints.py:504-512
            vemb = None
        elif self.frozen_view_dir:
            vemb = s['viewdirs_emb'].expand(len(ray_id), -1)
        else:
            vemb = poc_fre(viewdirs, self.view_poc)[ray_id]
        rgb = torch.sigmoid(self.rgbnet(h, vemb))
        lbs_w = None
        if merged_weights is not None:
            lbs_w = (merged_weights[s_i, :] * w).sum(dim=1)
        self.trace.update(rel_p=rel_p, to_nn=to_nn, rel_c=rel_c, h=h, density=density, alpha=alpha, rgb=rgb,
                          alpha_direct=alpha_direct, rgbs_direct=rgbs_direct, idw=w.squeeze(-1))
        return rgb, alpha, rgbs_direct, alpha_direct, lbs_w

    # -- compositing (lib/temporalpoints.py:611-677) ------------------------------------
    def composite(self, alpha, rgb, ray_id, step_id, n_rays, bg, extra=None):
        thres = self.thres
        if thres > 0:
            m = torch.where(alpha > thres)[0]
            alpha, rgb, ray_id, step_id = alpha[m], rgb[m], ray_id[m], step_id[m]
            if extra is not None:
                extra = extra[m]
        weights, last = Alphas2Weights.apply(alpha, ray_id, n_rays)
        if thres > 0:
            m = torch.where(weights > thres)[0]
            weights, alpha, rgb, ray_id, step_id = weights[m], alpha[m], rgb[m], ray_id[m], step_id[m]
            if extra is not None:
                extra = extra[m]
        rgb_marched = torch.zeros(n_rays, 3).index_add_(0, ray_id, weights.unsqueeze(-1) * rgb)
        rgb_marched = rgb_marched + last.unsqueeze(-1) * bg
        depth = torch.zeros(n_rays).index_add_(0, ray_id, weights * step_id)
        ex = None
        if extra is not None:
            ex = torch.zeros(n_rays, extra.shape[-1]).index_add_(0, ray_id, weights.unsqueeze(-1) * extra)
        return rgb_marched, last, depth, ex, weights, ray_id

    # -- full forward -------------------------------------------------------------------
    def forward(self, t=None, rot_params=None, *, rays_o, rays_d, viewdirs, near, far, stepsize, bg,
                query_radius=0.01, render_weights=False, cloud=None, ginv3=None):
        """`cloud` (test hook, not in the reference): evaluate everything downstream of the warp on THIS warped cloud
        (values replaced, autograd graph of the warp kept).  The reference's sampler is discontinuous in the last bit of
        the cloud's bbox (lib/cuda/render_utils_kernel.cu:23-33), so two correct warps that differ by one ulp keep
        slightly different sample sets; comparisons that are meant to pin the stages AFTER the warp pass the other
        side's cloud (and inverse frames, `ginv3`) in."""
        assert (t is None) ^ (rot_params is None)
        s = self.s
        wp = self.warp(t, rot_params)
        if cloud is not None:
            wp['xyz'] = wp['xyz'] + (cloud.to(F32) - wp['xyz']).detach()
        Ginv = torch.inverse(wp['G'])
        if ginv3 is not None:                     # same hook for the inverse frames (N,3,3)
            pad = torch.zeros_like(Ginv)
            pad[:, :3, :3] = ginv3.to(F32) - Ginv[:, :3, :3].detach()
            Ginv = Ginv + pad
        pose_embedding = None
        if self.pose_dim > 0:
            delta = (s['joints'] - wp['joints_rel']).clone().detach()
            pose_embedding = self.pose_embedding_net(poc_fre(delta, self.pos_poc).view(1, -1))
        n_rays = len(rays_o)
        smp = self.sample_and_knn(wp['xyz'], rays_o, rays_d, near, far, stepsize, query_radius)
        self.trace = dict(xyz=wp['xyz'], G=wp['G'], Ginv=Ginv, weights=wp['weights'], bone_Ts=wp['bone_Ts'],
                          pose_embedding=pose_embedding)
        ret = dict(t_hat_pcd=wp['xyz'], joints_rel=wp['joints_rel'], global_t=wp['global_t'], thetas=wp['thetas'])
        if smp is None:
            ret.update(rgb_marched=torch.ones(n_rays, 3) * bg, rgb_marched_direct=torch.ones(n_rays, 3) * bg,
                       depth=torch.zeros(n_rays), alphainv_last=None)
            return ret
        self.trace.update({k: smp[k] for k in ('bbox_min', 'bbox_max', 'pts', 'ray_id', 'step_id', 's_i', 'T', 'S')})
        rgb, alpha, rgbs_d, alpha_d, lbs_w = self.aggregate(
            wp['xyz'], Ginv, smp, viewdirs, stepsize, pose_embedding, wp['weights'] if render_weights else None)
        extra = None
        if render_weights:
            mask = wp['weights'].sum(dim=0) > 0
            cols = torch.tensor(hls_palette(int(mask.sum())), dtype=F32)
            gen = torch.Generator().manual_seed(0)
            cols = cols[torch.randperm(cols.shape[0], generator=gen)]
            extra = lbs_w[:, torch.where(mask)[0]] @ cols
        rgb_m, last, depth, wimg, wts, rid = self.composite(alpha, rgb, smp['ray_id'], smp['step_id'], n_rays, bg, extra)
        rgb_md, last_d, _, _, _, _ = self.composite(alpha_d, rgbs_d, smp['ray_id'], smp['step_id'], n_rays, bg)
        self.trace.update(weights_kept=wts, ray_id_kept=rid)
        ret.update(rgb_marched=rgb_m, alphainv_last=last, depth=depth, rgb_marched_direct=rgb_md,
                   alphainv_last_direct=last_d)
        if render_weights:
            ret['weights'] = wimg + last.unsqueeze(-1) * bg
        return ret
