"""Torch-facing operators over libapn_sm100.so: tensors in, tensors out, autograd where the
reference differentiates.  Every function launches hand-written sm_100a kernels through the C ABI
(include/apn.h); nothing here computes on the CPU or falls back to PyTorch ops.

Stage names follow the reference's profiler ranges (lib/temporalpoints.py:421-653):
forward_warp -> sample_ray / knn / knn-post -> feat_net / densitynet / rgbnet -> pre-mask /
Alphas2Weights / post-mask / segment_coo.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import AggGrads, AggInputs, AggOutputs, AdamTensor, MlpWeights, check, ptr, stage, stream

K_NEIGHBOURS = 8
FEAT_DIM = 128
PE_POS = 63
PE_VIEW = 27
FV_LD = 160
V0_DIM = 64
KNN_SORTED_MIN = 16384          # candidates from which the cell-sorted k-NN search is used (below: warp / thread searches)


# set by train.GradBucket: backward kernels that accumulate may write straight into an existing leaf .grad
DIRECT_GRAD_ACCUM = False


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_ITEMSIZE = {torch.float32: 4, torch.int32: 4, torch.uint8: 1, torch.int64: 8, torch.float16: 2}


def _zeros_like_many(tensors):
    """Zero gradients for a list of tensors as views of ONE flat buffer (one memset instead of one per tensor)."""
    sizes = [(t.numel() + 3) // 4 * 4 for t in tensors]          # keep every view 16-byte aligned
    flat = torch.zeros(sum(sizes), device=tensors[0].device, dtype=torch.float32)
    out, o = [], 0
    for t, n in zip(tensors, sizes):
        out.append(flat[o:o + t.numel()].view(t.shape))
        o += n
    return out


def _empty(shape, device, dtype=torch.float32):
    """torch.empty whose storage size is rounded up to quarter-octave buckets once it exceeds 256 KiB: the sizes
    of the per-step buffers follow the (data-dependent) sample counts, and a handful of bucket sizes keeps them
    inside the caching allocator instead of reaching cudaMalloc whenever a batch is a little larger than any
    before it."""
    if isinstance(shape, int):
        shape = (shape,)
    n = 1
    for d in shape:
        n *= int(d)
    itemsize = _ITEMSIZE.get(dtype) or torch.empty((), dtype=dtype).element_size()
    if n * itemsize <= (1 << 18):
        return torch.empty(shape, device=device, dtype=dtype)
    q = 1 << max(n.bit_length() - 3, 0)
    cap = (n + q - 1) // q * q
    return torch.empty((cap,), device=device, dtype=dtype)[:n].view(shape)


# --------------------------------------------------------------------------------------
# K0: pose chain (TransformNet + Rodrigues + kinematic chain)
# --------------------------------------------------------------------------------------
@dataclass
class PoseTables:
    """Static int tables of the kinematic tree on the device (lib/pointwarper.py:95-116, 232-234)."""
    parent_node: torch.Tensor   # (J) int32, -1 for the root
    pivot: torch.Tensor         # (J) int32
    sibling: torch.Tensor       # (J) int32
    rot_mask: Optional[torch.Tensor]   # (J) uint8 or None


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t)
    return arr


class _Pose(torch.autograd.Function):
    """t_embed, joints, TransformNet weights -> bone_Ts (J,4,4), global_t (3), thetas (J) in one launch
    (csrc/pose.cu); backward in one launch."""

    @staticmethod
    def forward(ctx, tb: PoseTables, t_embed, joints, *wb):
        lib = _lib.load()
        t_embed, joints = _f32(t_embed).reshape(-1), _f32(joints)
        wb = [_f32(x) for x in wb]                    # w0, b0, w1, b1, w2, b2, w3, b3, w4
        ws, bs = [wb[0], wb[2], wb[4], wb[6], wb[8]], [wb[1], wb[3], wb[5], wb[7]]
        J, dev = joints.shape[0], joints.device
        bone_T, global_t, thetas = _empty((J, 4, 4), dev), _empty((3,), dev), _empty((J,), dev)
        saved = _empty((lib.apn_pose_saved_bytes(J) // 4,), dev)
        with stage("transform_net"):
            check(lib.apn_pose_fwd(ptr(t_embed), t_embed.numel(), _ptr_array(ws), _ptr_array(bs), ptr(joints), ptr(tb.parent_node),
                                   ptr(tb.pivot), ptr(tb.sibling), ptr(tb.rot_mask), J, ptr(bone_T), ptr(global_t), ptr(thetas),
                                   ptr(saved), stream()), "apn_pose_fwd")
        ctx.tb = tb
        ctx.save_for_backward(t_embed, joints, saved, *wb)
        return bone_T, global_t, thetas

    @staticmethod
    def backward(ctx, d_bone_T, d_global_t, d_thetas):
        lib = _lib.load()
        tb = ctx.tb
        t_embed, joints, saved, *wb = ctx.saved_tensors
        ws, bs = [wb[0], wb[2], wb[4], wb[6], wb[8]], [wb[1], wb[3], wb[5], wb[7]]
        J, dev = joints.shape[0], joints.device
        d_bone_T = torch.zeros(J, 4, 4, device=dev) if d_bone_T is None else _f32(d_bone_T)
        d_global_t = None if d_global_t is None else _f32(d_global_t)
        d_thetas = None if d_thetas is None else _f32(d_thetas)
        # the kernel OVERWRITES its outputs; a caller that owns zeroed gradient storage (train.FusedTrainStep) passes it in
        out = getattr(ctx, "grad_out", None)
        d_wb = out["wb"] if out else [torch.empty_like(x) for x in wb]
        d_ws, d_bs = [d_wb[0], d_wb[2], d_wb[4], d_wb[6], d_wb[8]], [d_wb[1], d_wb[3], d_wb[5], d_wb[7]]
        d_joints = out["joints"] if out else torch.empty_like(joints)
        with stage("transform_net_bwd"):
            check(lib.apn_pose_bwd(ptr(t_embed), t_embed.numel(), _ptr_array(ws), _ptr_array(bs), ptr(joints), ptr(tb.parent_node),
                                   ptr(tb.pivot), ptr(tb.sibling), ptr(tb.rot_mask), J, ptr(saved), ptr(d_bone_T), ptr(d_global_t),
                                   ptr(d_thetas), _ptr_array(d_ws), _ptr_array(d_bs), ptr(d_joints), stream()), "apn_pose_bwd")
        return (None, None, d_joints, *d_wb)


def pose_chain(tb: PoseTables, t_embed, joints, weights_and_biases):
    """-> bone_Ts (J,4,4), global_t (3), thetas (J).  weights_and_biases = [w0,b0,w1,b1,w2,b2,w3,b3,w4]."""
    return _Pose.apply(tb, t_embed, joints, *weights_and_biases)


# --------------------------------------------------------------------------------------
# K1: linear blend skinning
# --------------------------------------------------------------------------------------
class _LBS(torch.autograd.Function):
    """get_weights + blend + transform + frame inverse + bbox (lib/temporalpoints.py:401-414,424,569;
    lib/pointwarper.py:241-266).  theta_weight=None: `raw_w` already holds the final weights."""

    @staticmethod
    def forward(ctx, raw_w, theta_weight, bone_T, global_t, xyz, rules, eps, want_frames, want_weights=True):
        lib = _lib.load()
        raw_w, bone_T, xyz = _f32(raw_w), _f32(bone_T), _f32(xyz)
        theta_weight = None if theta_weight is None else _f32(theta_weight)
        global_t = None if global_t is None else _f32(global_t)
        N, J = raw_w.shape
        dev = raw_w.device
        xyz_out = _empty((N, 3), dev)
        ginv = _empty((N, 9), dev)
        # merged skinning weights: a second 4J bytes per point of store traffic, skipped when nobody reads them
        w_out = _empty((N, J), dev) if want_weights else None
        g_out = _empty((N, 4, 4), dev) if want_frames else None
        bbox = _empty((6,), dev)
        with stage("forward_warp"):
            check(lib.apn_lbs_fwd(ptr(raw_w), ptr(theta_weight), float(eps), ptr(rules), ptr(bone_T), ptr(xyz), ptr(global_t),
                                  N, J, ptr(xyz_out), ptr(ginv), ptr(w_out), ptr(g_out), ptr(bbox), stream()), "apn_lbs_fwd")
        ctx.save_for_backward(raw_w, theta_weight, bone_T, xyz, ginv, rules)
        ctx.eps = float(eps)
        ctx.has_gt = global_t is not None
        ctx.want_frames = want_frames
        ctx.mark_non_differentiable(bbox)
        return (xyz_out, ginv, (w_out if want_weights else torch.zeros(0, device=dev)), bbox,
                (g_out if want_frames else torch.zeros(0, device=dev)))

    @staticmethod
    def backward(ctx, d_xyz, d_ginv, d_w, _d_bbox, d_g):
        lib = _lib.load()
        raw_w, theta_weight, bone_T, xyz, ginv, rules = ctx.saved_tensors
        N, J = raw_w.shape
        dev = raw_w.device
        d_xyz = None if d_xyz is None else _f32(d_xyz)
        d_ginv = None if d_ginv is None else _f32(d_ginv)
        d_w = None if d_w is None else _f32(d_w)
        d_g = _f32(d_g) if (ctx.want_frames and d_g is not None) else None
        out = getattr(ctx, "grad_out", None)          # see _Pose.backward
        d_raw = out["raw"] if out else _empty((N, J), dev)
        d_theta = (out["theta"] if out else _empty((1,), dev)) if theta_weight is not None else None
        d_bone = _empty((J, 4, 4), dev)
        d_gt = _empty((3,), dev)
        ws_bytes = lib.apn_lbs_bwd_workspace_bytes(N, J)
        ws = _empty((ws_bytes,), dev, torch.uint8)
        with stage("forward_warp_bwd"):
            check(lib.apn_lbs_bwd(ptr(raw_w), ptr(theta_weight), ctx.eps, ptr(rules), ptr(bone_T), ptr(xyz), N, J, ptr(ginv),
                                  ptr(d_xyz), ptr(d_ginv), ptr(d_w), ptr(d_g), ptr(d_raw), ptr(d_theta), ptr(d_bone), ptr(d_gt),
                                  ptr(ws), ws_bytes, stream()), "apn_lbs_bwd")
        return (d_raw, None if d_theta is None else d_theta.reshape(theta_weight.shape), d_bone,
                (d_gt if ctx.has_gt else None), None, None, None, None, None)


def lbs(raw_w, theta_weight, bone_T, global_t, xyz, rules: Optional[torch.Tensor] = None, eps: float = 1e-6,
        want_frames: bool = False, want_weights: bool = True):
    """-> warped xyz (N,3), inverse frames (N,9), merged skinning weights (N,J) (None unless want_weights), bbox (6)
    [min, max], blended frames (N,4,4) if want_frames."""
    if rules is not None:
        rules = rules.to(torch.int32).contiguous()
    xyz_out, ginv, w, bbox, g = _LBS.apply(raw_w, theta_weight, bone_T, global_t, xyz, rules, eps, want_frames, want_weights)
    w = w if want_weights else None
    return (xyz_out, ginv, w, bbox, g) if want_frames else (xyz_out, ginv, w, bbox)


# --------------------------------------------------------------------------------------
# K2: grid + ray samples + exact 8-NN
# --------------------------------------------------------------------------------------
class GridOverflow(_lib.ApnError):
    """The padded bbox of the cloud holds more grid leaves than the grid's cell table."""


class Grid:
    """Opaque multi-level uniform grid over a warped cloud (csrc/grid_knn.cu)."""

    def __init__(self, xyz: torch.Tensor, bbox: torch.Tensor, query_radius: float, bbox_pad: float, cell_hint: float,
                 cell_capacity: Optional[int] = None):
        lib = _lib.load()
        self.xyz = _f32(xyz.detach())
        self.N = self.xyz.shape[0]
        if cell_capacity is None:
            cell_capacity = self.default_capacity(self.N)
        self.cell_capacity = cell_capacity
        self._args = (query_radius, bbox_pad, cell_hint)
        self._bbox = bbox
        self.bytes = lib.apn_grid_workspace_bytes(self.N, cell_capacity)
        self.blob = _empty((self.bytes,), self.xyz.device, torch.uint8)
        with stage("grid_build"):
            check(lib.apn_grid_build(ptr(self.xyz), ptr(_f32(bbox)), self.N, float(query_radius), float(bbox_pad),
                                     float(cell_hint), cell_capacity, ptr(self.blob), self.bytes, stream()), "apn_grid_build")

    MAX_CAPACITY = 1 << 27

    @staticmethod
    def default_capacity(n_points: int) -> int:
        # 2^20 leaves = 16384 top cells of edge 1.01*sqrt(r) = 17 units^3 of padded bbox at the reference's radius
        return int(min(max(8 * n_points, 1 << 20), 1 << 25))

    def grow(self) -> None:
        """Rebuilds this grid IN PLACE with 8x the cell capacity (after an overflow: the padded bbox holds more leaves
        than the table); callers that cache the grid per pose keep a valid object."""
        if self.cell_capacity >= self.MAX_CAPACITY:
            raise GridOverflow(f"the warped cloud's bbox needs more than {self.MAX_CAPACITY} grid leaves")
        g = Grid(self.xyz, self._bbox, *self._args, cell_capacity=min(self.cell_capacity * 8, self.MAX_CAPACITY))
        self.cell_capacity, self.bytes, self.blob = g.cell_capacity, g.bytes, g.blob

    def overflowed(self) -> bool:
        """Synchronising check of the header's overflow flag (set on the device by apn_grid_build)."""
        return bool(self.describe()["overflow"])

    def describe(self) -> dict:
        """Synchronising debug helper: header fields of the grid."""
        lib = _lib.load()
        host = (C.c_float * 64)()
        check(lib.apn_grid_describe(ptr(self.blob), C.cast(host, C.c_void_p), stream()), "apn_grid_describe")
        raw = bytes(host)
        import struct
        f = struct.unpack_from("<6f", raw, 0)
        i = struct.unpack_from("<6i", raw, 24)
        r2 = struct.unpack_from("<f", raw, 48)[0]
        bm = struct.unpack_from("<6f", raw, 52)
        n_points, overflow = struct.unpack_from("<2i", raw, 76)
        return dict(origin=f[0:3], cell=f[3], inv_cell=f[4], top_cell=f[5], L=i[0], top_dim=i[1:4], n_top=i[4],
                    n_cells=i[5], r2=r2, bmin=bm[0:3], bmax=bm[3:6], n_points=n_points, overflow=overflow)

    def knn(self, query: torch.Tensor, k: int = K_NEIGHBOURS, check_overflow: bool = True):
        """Exact k-NN (k <= 8) of arbitrary points: indices (n,k) int32 ascending by (d2, index), d2 (n,k)."""
        lib = _lib.load()
        q = _f32(query.detach())
        n = q.shape[0]
        idx = _empty((n, k), q.device, torch.int32)
        d2 = _empty((n, k), q.device)
        check(lib.apn_knn_points(ptr(q), n, ptr(self.blob), k, ptr(idx), ptr(d2), stream()), "apn_knn_points")
        if check_overflow and self.overflowed():     # the kernel wrote index -1 everywhere: never hand that to a gather
            self.grow()
            return self.knn(query, k, check_overflow)
        return idx, d2


def nn1_batched(query: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """(B, Nq, D), (B, Nt, D), D in {2, 3} -> int64 (B, Nq): nearest target of every query inside its batch item
    (KeOps argKmin K=1 of lib/temporalpoints.py:783-787); ties -> lowest index."""
    lib = _lib.load()
    q, t = query.detach().float().contiguous(), target.detach().float().contiguous()
    assert q.dim() == 3 and t.dim() == 3 and q.shape[0] == t.shape[0] and q.shape[2] == t.shape[2]
    B, nq, d = q.shape
    idx = _empty((B, nq), q.device, torch.int32)
    check(lib.apn_nn1_batched(ptr(q), ptr(t), B, nq, t.shape[1], d, ptr(idx), stream()), "apn_nn1_batched")
    return idx.long()


def rays_of_a_view(H: int, W: int, K, c2w, device, inverse_y=False, flip_x=False, flip_y=False, pixel_ids=None,
                   first_pixel: int = 0, n: Optional[int] = None):
    """get_rays_of_a_view (lib/tineuvox.py:675-738) on the device, from the camera alone: -> rays_o, rays_d, viewdirs (n,3).
    `pixel_ids` (int32 device tensor) selects pixels (row-major indices); otherwise n pixels from first_pixel (default all)."""
    lib = _lib.load()
    Kh = torch.as_tensor(K, dtype=torch.float32).detach().cpu().contiguous().reshape(-1)
    ch = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu().contiguous()
    assert Kh.numel() == 9 and ch.dim() == 2 and ch.shape[1] == 4 and ch.shape[0] in (3, 4)
    if pixel_ids is not None:
        pixel_ids = pixel_ids.to(device=device, dtype=torch.int32).contiguous()
        n = pixel_ids.numel()
    elif n is None:
        n = H * W - first_pixel
    ro, rd, vd = (_empty((n, 3), device) for _ in range(3))
    check(lib.apn_rays_of_a_view(Kh.data_ptr(), ch.data_ptr(), ch.shape[0], int(H), int(W), int(bool(inverse_y)), int(bool(flip_x)),
                                 int(bool(flip_y)), ptr(pixel_ids), int(first_pixel), int(n), ptr(ro), ptr(rd), ptr(vd), stream()),
          "apn_rays_of_a_view")
    return ro, rd, vd


def time_embed(t: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """poc_fre of the scalar time (lib/tineuvox.py:872-878) in one launch: (1) -> (1 + 2 F)."""
    lib = _lib.load()
    t, freqs = _f32(t).reshape(-1), _f32(freqs)
    assert t.numel() == 1
    out = torch.empty(1 + 2 * freqs.numel(), device=t.device, dtype=torch.float32)
    check(lib.apn_time_embed(ptr(t), ptr(freqs), freqs.numel(), ptr(out), stream()), "apn_time_embed")
    return out


def mse_loss_grad(pred: torch.Tensor, target: torch.Tensor, weight: float):
    """-> (loss (0-dim) = weight * mse(pred, target), d loss / d pred); run.py:617-621 in one launch."""
    lib = _lib.load()
    pred, target = _f32(pred), _f32(target)
    assert pred.shape == target.shape
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred)
    check(lib.apn_mse_loss_grad(ptr(pred), ptr(target), pred.numel(), float(weight), ptr(loss), ptr(grad), stream()),
          "apn_mse_loss_grad")
    return loss[0], grad


def point_regularisers(xyz, w, nn_i32, nn_dist, eps: float, weight_arap: float, weight_tv: float, weight_sparsity: float,
                       d_xyz: torch.Tensor, d_w: Optional[torch.Tensor] = None, losses: Optional[torch.Tensor] = None):
    """ARAP + neighbour weight TV + weight sparsity (lib/temporalpoints.py:714-725, weighted as run.py:633-648 does) with
    their gradients in one launch: d_xyz (N,3) is accumulated into, d_w (N,J) overwritten.  -> (losses (3), d_w)."""
    lib = _lib.load()
    N, K = nn_i32.shape
    assert nn_i32.dtype == torch.int32 and nn_i32.is_contiguous() and d_xyz.is_contiguous()
    need_w = weight_tv != 0.0 or weight_sparsity != 0.0
    J = w.shape[1] if w is not None else 1
    if need_w and d_w is None:
        d_w = _empty((N, J), xyz.device)
    if losses is None:
        losses = _empty((3,), xyz.device)
    check(lib.apn_point_regularisers(ptr(_f32(xyz)), ptr(_f32(w)) if need_w else None, ptr(nn_i32), ptr(_f32(nn_dist)), N, K, J,
                                     float(eps), float(weight_arap), float(weight_tv), float(weight_sparsity), ptr(d_xyz),
                                     ptr(d_w) if need_w else None, ptr(losses), stream()), "apn_point_regularisers")
    return losses, (d_w if need_w else None)


def pose_regularisers(thetas, global_t, joints, skeleton, weight_transformation_reg: float, weight_joint_chamfer: float,
                      losses: Optional[torch.Tensor] = None):
    """Transformation regulariser + joint chamfer loss (lib/temporalpoints.py:797-800, 731-733) with gradients in one launch.
    -> (losses (2), d_thetas (J) or None, d_global_t (3) or None, d_joints (J,3) or None)."""
    lib = _lib.load()
    J = joints.shape[0]
    dev = joints.device
    d_thetas = _empty((J,), dev) if weight_transformation_reg != 0.0 else None
    d_global_t = _empty((3,), dev) if weight_transformation_reg != 0.0 else None
    d_joints = _empty((J, 3), dev) if weight_joint_chamfer != 0.0 else None
    if losses is None:
        losses = _empty((2,), dev)
    skeleton = None if skeleton is None else _f32(skeleton)
    check(lib.apn_pose_regularisers(ptr(_f32(thetas)), ptr(_f32(global_t)), ptr(_f32(joints)), ptr(skeleton), J,
                                    0 if skeleton is None else skeleton.shape[0], float(weight_transformation_reg),
                                    float(weight_joint_chamfer), ptr(d_thetas), ptr(d_global_t), ptr(d_joints), ptr(losses), stream()),
          "apn_pose_regularisers")
    return losses, d_thetas, d_global_t, d_joints


def exclusive_scan(x: torch.Tensor) -> torch.Tensor:
    """int32 (n) -> int32 (n+1), out[n] = total."""
    lib = _lib.load()
    n = x.numel()
    out = _empty((n + 1,), x.device, torch.int32)
    ws_bytes = lib.apn_scan_workspace_bytes(n)
    ws = _empty((max(ws_bytes, 1),), x.device, torch.uint8)
    check(lib.apn_exclusive_scan_i32(ptr(x), ptr(out), n, ptr(ws), ws_bytes, stream()), "apn_exclusive_scan_i32")
    return out


@dataclass
class Samples:
    """Kept ray samples, ray-major and near-to-far (what lib/temporalpoints.py:427-447 leaves)."""
    pts: torch.Tensor        # (M,3)
    ray_id: torch.Tensor     # (M) int32
    step_id: torch.Tensor    # (M) int32
    nn_idx: torch.Tensor     # (M,8) int32
    ray_start: torch.Tensor  # (R+1) int32
    n_rays: int
    n_candidates: int
    # static (sync-free) mode: the arrays above are CAPACITY-sized and the true sample count lives on the device
    m_dev: Optional[torch.Tensor] = None     # (1) int32 view of counts[1]
    counts: Optional[torch.Tensor] = None    # (5) int32: candidates used, samples kept, flags, candidates found, samples found

    @property
    def M(self) -> int:
        return self.pts.shape[0]


def sample_and_knn(grid: Grid, rays_o: torch.Tensor, rays_d: torch.Tensor, near: float, far: float, stepdist: float,
                   return_d2: bool = False):
    """sample_ray + Kmin_argKmin + radius rule, fused (lib/temporalpoints.py:421-447)."""
    with stage("sample_ray+knn"):
        while True:
            try:
                return _sample_and_knn(grid, rays_o, rays_d, near, far, stepdist, return_d2)
            except GridOverflow:               # widely spread cloud: larger cell table (raises at the limit)
                grid.grow()


def _sample_and_knn(grid, rays_o, rays_d, near, far, stepdist, return_d2):
    lib = _lib.load()
    rays_o, rays_d = _f32(rays_o), _f32(rays_d)
    R = rays_o.shape[0]
    dev = rays_o.device
    st = stream()
    count = _empty((R,), dev, torch.int32)
    check(lib.apn_ray_candidates(ptr(rays_o), ptr(rays_d), R, near, far, stepdist, ptr(grid.blob), 0, ptr(count), None,
                                 None, None, st), "apn_ray_candidates(count)")
    base = exclusive_scan(count)
    n_cand = int(base[R].item())
    if n_cand < 0:                     # grid overflow, reported through the count (csrc/grid_knn.cu ray_candidates_kernel)
        raise GridOverflow("grid cell table too small for the warped cloud's bbox")
    cand_ray = _empty((n_cand,), dev, torch.int32)
    cand_step = _empty((n_cand,), dev, torch.int32)
    nn_c = _empty((n_cand, K_NEIGHBOURS), dev, torch.int32)
    keep = _empty((n_cand,), dev, torch.int32)
    d2_c = _empty((n_cand, K_NEIGHBOURS), dev) if return_d2 else None
    if n_cand > 0:
        check(lib.apn_ray_candidates(ptr(rays_o), ptr(rays_d), R, near, far, stepdist, ptr(grid.blob), 1, None, ptr(base),
                                     ptr(cand_ray), ptr(cand_step), st), "apn_ray_candidates(fill)")
        # search: the cell-sorted, shared-memory staged kernel for batches large enough to amortise its sort; the warp /
        # thread searches below that (APN_KNN_FORCE=sorted|warp|thread|thread0|thread1 pins one: the tests run them all)
        forced = os.environ.get("APN_KNN_FORCE", "")
        if forced.startswith("s") or (not forced and n_cand >= KNN_SORTED_MIN):
            wsb = lib.apn_knn_sorted_workspace_bytes(n_cand)
            ws = _empty((wsb,), dev, torch.uint8)
            check(lib.apn_knn_sorted(ptr(rays_o), ptr(rays_d), near, far, stepdist, ptr(grid.blob), ptr(cand_ray), ptr(cand_step),
                                     n_cand, ptr(nn_c), ptr(d2_c), ptr(keep), ptr(ws), wsb, st), "apn_knn_sorted")
        else:
            check(lib.apn_knn(ptr(rays_o), ptr(rays_d), near, far, stepdist, ptr(grid.blob), ptr(cand_ray), ptr(cand_step),
                              n_cand, ptr(nn_c), ptr(d2_c), ptr(keep), st), "apn_knn")
    kept_pos = exclusive_scan(keep)
    M = int(kept_pos[n_cand].item())
    pts = _empty((M, 3), dev)
    ray_id = _empty((M,), dev, torch.int32)
    step_id = _empty((M,), dev, torch.int32)
    nn_idx = _empty((M, K_NEIGHBOURS), dev, torch.int32)
    ray_start = _empty((R + 1,), dev, torch.int32)
    check(lib.apn_compact_samples(ptr(rays_o), ptr(rays_d), near, far, stepdist, ptr(grid.blob), ptr(cand_ray),
                                  ptr(cand_step), ptr(base), ptr(keep), ptr(kept_pos), ptr(nn_c), n_cand, R, ptr(pts),
                                  ptr(ray_id), ptr(step_id), ptr(nn_idx), ptr(ray_start), st), "apn_compact_samples")
    smp = Samples(pts, ray_id, step_id, nn_idx, ray_start, R, n_cand)
    if return_d2:
        return smp, dict(cand_ray=cand_ray, cand_step=cand_step, keep=keep, d2=d2_c, nn=nn_c)
    return smp


class StaticSampler:
    """Fixed-capacity workspace + output arrays of the sync-free sampling stage (apn_sample_knn_static): nothing is sized
    from data, no count is read back, so the stage (and the training step around it) can be enqueued ahead of the GPU and
    captured in a CUDA graph.  Flags in counts[2] (1 grid overflow | 2 candidate list truncated | 4 samples truncated)."""

    def __init__(self, n_rays: int, cand_cap: int, m_cap: int, device):
        lib = _lib.load()
        self.R, self.cand_cap, self.m_cap = int(n_rays), int(cand_cap), int(m_cap)
        self.ws_bytes = lib.apn_sample_knn_static_workspace_bytes(self.R, self.cand_cap)
        self.ws = torch.empty(self.ws_bytes, device=device, dtype=torch.uint8)
        self.pts = torch.empty(self.m_cap, 3, device=device)
        self.ray_id = torch.empty(self.m_cap, device=device, dtype=torch.int32)
        self.step_id = torch.empty(self.m_cap, device=device, dtype=torch.int32)
        self.nn_idx = torch.empty(self.m_cap, K_NEIGHBOURS, device=device, dtype=torch.int32)
        self.ray_start = torch.empty(self.R + 1, device=device, dtype=torch.int32)
        self.counts = torch.zeros(8, device=device, dtype=torch.int32)

    def run(self, grid: "Grid", rays_o, rays_d, near: float, far: float, stepdist: float) -> "Samples":
        rays_o, rays_d = _f32(rays_o), _f32(rays_d)
        assert rays_o.shape[0] == self.R
        with stage("sample_ray+knn"):
            check(_lib.load().apn_sample_knn_static(ptr(rays_o), ptr(rays_d), self.R, near, far, stepdist, ptr(grid.blob),
                                                    self.cand_cap, self.m_cap, ptr(self.ws), self.ws_bytes, ptr(self.pts),
                                                    ptr(self.ray_id), ptr(self.step_id), ptr(self.nn_idx), ptr(self.ray_start),
                                                    ptr(self.counts), stream()), "apn_sample_knn_static")
        return Samples(self.pts, self.ray_id, self.step_id, self.nn_idx, self.ray_start, self.R, self.cand_cap,
                       m_dev=self.counts[1:2], counts=self.counts)


# --------------------------------------------------------------------------------------
# K3: aggregation (exact fp32 path, differentiable)
# --------------------------------------------------------------------------------------
MLP_KEYS = ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3", "density_w", "density_b", "rgb_feat_w", "rgb_feat_b",
            "rgb_v0_w", "rgb_v0_b", "rgb_v2_w", "rgb_v2_b")


def _mlp_struct(ws: Sequence[torch.Tensor]) -> MlpWeights:
    s = MlpWeights()
    for l in range(4):
        s.w[l] = ptr(ws[2 * l])
        s.b[l] = ptr(ws[2 * l + 1])
    (s.density_w, s.density_b, s.rgb_feat_w, s.rgb_feat_b, s.rgb_v0_w, s.rgb_v0_b, s.rgb_v2_w,
     s.rgb_v2_b) = [ptr(t) for t in ws[8:16]]
    return s


@dataclass
class AggConst:
    """Non-tensor / non-differentiable inputs of the aggregation."""
    pts: torch.Tensor
    nn_idx: torch.Tensor
    ray_id: torch.Tensor
    viewdirs: torch.Tensor
    canonical_alpha: torch.Tensor
    canonical_rgbs: torch.Tensor
    direct_eps: torch.Tensor
    mean_min_distance: float
    eps: float
    act_shift: float
    interval: float
    direct: bool = True
    # static mode: pts / nn_idx / ray_id are capacity-sized, the true sample count is this (1) int32 device tensor
    m_dev: Optional[torch.Tensor] = None


def _agg_inputs(c: AggConst, xyz, ginv, feat, pose_emb, M, d_in) -> AggInputs:
    a = AggInputs()
    a.M, a.N, a.d_in = M, xyz.shape[0], d_in
    a.pts, a.nn_idx, a.ray_id = ptr(c.pts), ptr(c.nn_idx), ptr(c.ray_id)
    a.xyz, a.ginv, a.feat, a.pose_emb = ptr(xyz), ptr(ginv), ptr(feat), ptr(pose_emb)
    a.viewdirs = ptr(c.viewdirs)
    a.canonical_alpha, a.canonical_rgbs, a.direct_eps = ptr(c.canonical_alpha), ptr(c.canonical_rgbs), ptr(c.direct_eps)
    a.mean_min_distance, a.eps, a.act_shift, a.interval = c.mean_min_distance, c.eps, c.act_shift, c.interval
    a.m_dev = ptr(c.m_dev)
    return a


class _Aggregate(torch.autograd.Function):
    """lib/temporalpoints.py:446-515 after the k-NN.  Differentiable w.r.t. the warped cloud, the inverse
    frames, canonical_feat, the pose embedding and all MLP weights.  The direct branch
    (lib/temporalpoints.py:459-470) is returned without a graph: no loss in run.py reads it."""

    @staticmethod
    def forward(ctx, c: AggConst, xyz, ginv, feat, pose_emb, *ws):
        lib = _lib.load()
        xyz, ginv, feat = _f32(xyz), _f32(ginv), _f32(feat)
        pose_emb = None if pose_emb is None else _f32(pose_emb).reshape(-1)
        ws = [_f32(w) for w in ws]
        M = c.pts.shape[0]
        dev = xyz.device
        d_in = PE_POS + FEAT_DIM + (0 if pose_emb is None else pose_emb.numel())
        ld0 = (d_in + 3) // 4 * 4
        rows = M * K_NEIGHBOURS
        alpha, rgb = _empty((M,), dev), _empty((M, 3), dev)
        alpha_d = _empty((M,), dev) if c.direct else None
        rgb_d = _empty((M, 3), dev) if c.direct else None
        idw = _empty((M, K_NEIGHBOURS), dev)
        need_grad = any(ctx.needs_input_grad)
        out = AggOutputs()
        out.alpha, out.rgb, out.alpha_direct, out.rgb_direct, out.idw = ptr(alpha), ptr(rgb), ptr(alpha_d), ptr(rgb_d), ptr(idw)
        a = _agg_inputs(c, xyz, ginv, feat, pose_emb, M, d_in)
        w = _mlp_struct(ws)
        saved = None
        scratch, scratch_bytes = None, 0
        if M > 0:
            if need_grad:
                saved = dict(x0=_empty((rows, ld0), dev), act=[_empty((rows, FEAT_DIM), dev) for _ in range(4)],
                             h=_empty((M, FEAT_DIM), dev), exp_d=_empty((M,), dev), fv=_empty((M, FV_LD), dev),
                             v0=_empty((M, V0_DIM), dev))
                out.x0 = ptr(saved["x0"])
                for l in range(4):
                    out.act[l] = ptr(saved["act"][l])
                out.h, out.exp_d, out.fv, out.v0 = ptr(saved["h"]), ptr(saved["exp_d"]), ptr(saved["fv"]), ptr(saved["v0"])
            else:
                scratch_bytes = lib.apn_aggregate_scratch_bytes(M, d_in)
                scratch = _empty((scratch_bytes,), dev, torch.uint8)
            with stage("feat_net"):
                check(lib.apn_aggregate_fwd(C.byref(a), C.byref(w), C.byref(out), ptr(scratch), scratch_bytes, stream()),
                      "apn_aggregate_fwd")
        # every tensor goes through save_for_backward: keeping outputs (alpha, rgb, idw) as plain attributes of ctx
        # would create a node -> tensor -> node reference cycle that only the Python GC frees, i.e. a slow leak of
        # device memory between collections
        ctx.c, ctx.d_in, ctx.has_pose, ctx.has_saved = c, d_in, pose_emb is not None, saved is not None
        extra = []
        if saved is not None:
            extra = [saved["x0"], *saved["act"], saved["h"], saved["exp_d"], saved["fv"], saved["v0"]]
        ctx.save_for_backward(xyz, ginv, feat, pose_emb if pose_emb is not None else xyz.new_empty(0), *ws, alpha, rgb, idw, *extra)
        ctx.mark_non_differentiable(idw)
        if c.direct:
            ctx.mark_non_differentiable(alpha_d, rgb_d)
        return alpha, rgb, alpha_d, rgb_d, idw

    @staticmethod
    def backward(ctx, d_alpha, d_rgb, *_):
        lib = _lib.load()
        c = ctx.c
        sv_all = ctx.saved_tensors
        xyz, ginv, feat, pose_emb = sv_all[0:4]
        if not ctx.has_pose:
            pose_emb = None
        ws = list(sv_all[4:20])
        alpha, rgb, idw = sv_all[20:23]
        M = c.pts.shape[0]
        dev = xyz.device
        need = ctx.needs_input_grad      # (c, xyz, ginv, feat, pose_emb, *ws)
        wanted = [t for t, n in ((xyz, need[1]), (ginv, need[2]), (feat, need[3])) if n]
        if pose_emb is not None and need[4]:
            wanted.append(pose_emb)
        zs = _zeros_like_many(wanted + list(ws))
        zi = iter(zs)
        d_xyz = next(zi) if need[1] else None
        d_ginv = next(zi) if need[2] else None
        d_feat = next(zi) if need[3] else None
        d_pose = next(zi) if (pose_emb is not None and need[4]) else None
        d_ws = list(zi)
        if M > 0:
            assert ctx.has_saved
            d_alpha = torch.zeros_like(alpha) if d_alpha is None else _f32(d_alpha)
            d_rgb = torch.zeros_like(rgb) if d_rgb is None else _f32(d_rgb)
            x0, a0, a1, a2, a3, h, exp_d, fv, v0 = sv_all[23:32]
            out = AggOutputs()
            out.alpha, out.rgb, out.idw = ptr(alpha), ptr(rgb), ptr(idw)
            out.x0 = ptr(x0)
            for l, a_ in enumerate((a0, a1, a2, a3)):
                out.act[l] = ptr(a_)
            out.h, out.exp_d, out.fv, out.v0 = ptr(h), ptr(exp_d), ptr(fv), ptr(v0)
            g = AggGrads()
            g.d_alpha, g.d_rgb = ptr(d_alpha), ptr(d_rgb)
            g.d_xyz, g.d_ginv, g.d_feat, g.d_pose_emb = ptr(d_xyz), ptr(d_ginv), ptr(d_feat), ptr(d_pose)
            for l in range(4):
                g.d_w[l] = ptr(d_ws[2 * l])
                g.d_b[l] = ptr(d_ws[2 * l + 1])
            (g.d_density_w, g.d_density_b, g.d_rgb_feat_w, g.d_rgb_feat_b, g.d_rgb_v0_w, g.d_rgb_v0_b, g.d_rgb_v2_w,
             g.d_rgb_v2_b) = [ptr(t) for t in d_ws[8:16]]
            a = _agg_inputs(c, xyz, ginv, feat, pose_emb, M, ctx.d_in)
            w = _mlp_struct(ws)
            sb = lib.apn_aggregate_bwd_scratch_bytes(M, ctx.d_in)
            scratch = _empty((sb,), dev, torch.uint8)
            with stage("feat_net_bwd"):
                check(lib.apn_aggregate_bwd(C.byref(a), C.byref(w), C.byref(out), C.byref(g), ptr(scratch), sb, stream()),
                      "apn_aggregate_bwd")
        if d_pose is not None and pose_emb is not None:
            d_pose = d_pose.reshape(pose_emb.shape)
        return (None, d_xyz, d_ginv, d_feat, d_pose, *[dw if need[5 + i] else None for i, dw in enumerate(d_ws)])


def aggregate(c: AggConst, xyz, ginv, feat, pose_emb, weights: Sequence[torch.Tensor]):
    """-> alpha (M), rgb (M,3), alpha_direct (M)|None, rgb_direct (M,3)|None, idw (M,8)."""
    assert len(weights) == 16, "expected feat_net (4x w,b), densitynet, rgbnet (3x w,b)"
    if pose_emb is not None:
        pose_emb = pose_emb.reshape(-1)          # (1, 64) -> (64): the gradient comes back in this shape through autograd
    return _Aggregate.apply(c, xyz, ginv, feat, pose_emb, *weights)


class PackedDecoder:
    """Derived decoder state of the tensor-core kernel (csrc/aggregate_tc.cu): feat_net weights in the kernel's
    shared-memory image and the per-point table canonical_feat @ W0_feat^T.  Rebuilt whenever a source tensor
    changes (tracked through the tensors' version counters)."""

    def __init__(self):
        self.buf = self.table = self.buf_bwd = None
        self.key = self.table_key = self.key_bwd = None

    # set while a CUDA graph is being captured: the packing launches must be part of the graph whatever the caches say
    # (the optimiser changes the weights between replays)
    force = False

    def get_bwd(self, ws: Sequence[torch.Tensor], d_in: int) -> torch.Tensor:
        """Transposed weight tiles of the dgrad kernel."""
        lib = _lib.load()
        key = tuple((w.data_ptr(), w._version) for w in ws[:8:2]) + (d_in,)
        if self.key_bwd != key or self.force:
            if self.buf_bwd is None:
                self.buf_bwd = _empty((lib.apn_aggregate_tc_bwd_weights_bytes(),), ws[0].device, torch.uint8)
            w = _mlp_struct(ws)
            check(lib.apn_aggregate_tc_pack_weights_bwd(C.byref(w), d_in, ptr(self.buf_bwd), stream()),
                  "apn_aggregate_tc_pack_weights_bwd")
            self.key_bwd = key
        return self.buf_bwd

    def get(self, ws: Sequence[torch.Tensor], d_in: int, feat: torch.Tensor):
        lib = _lib.load()
        key = tuple((w.data_ptr(), w._version) for w in ws[:8:2]) + (d_in,)
        if self.key != key or self.force:
            if self.buf is None:
                self.buf = _empty((lib.apn_aggregate_tc_weights_bytes(d_in),), ws[0].device, torch.uint8)
            w = _mlp_struct(ws)
            check(lib.apn_aggregate_tc_pack_weights(C.byref(w), d_in, ptr(self.buf), stream()), "apn_aggregate_tc_pack_weights")
            self.key = key
        tkey = (feat.data_ptr(), feat._version, feat.shape[0], ws[0].data_ptr(), ws[0]._version, d_in)
        if self.table_key != tkey or self.force:
            if self.table is None or self.table.shape[0] != feat.shape[0]:
                self.table = _empty((feat.shape[0], FEAT_DIM), feat.device)
            check(lib.apn_aggregate_tc_point_table(ptr(feat), ptr(ws[0]), d_in, feat.shape[0], ptr(self.table), stream()),
                  "apn_aggregate_tc_point_table")
            self.table_key = tkey
        return self.buf, self.table


def aggregate_tc(c: AggConst, xyz, ginv, feat, pose_emb, weights: Sequence[torch.Tensor], packed: PackedDecoder,
                 precision: int = 1):
    """Inference-only aggregation on the tcgen05 tensor cores (no autograd graph).
    precision 0: fp16 operands; 1: split-fp16 (fp32-class).  -> alpha, rgb, alpha_direct, rgb_direct, idw."""
    assert len(weights) == 16
    lib = _lib.load()
    with torch.no_grad():
        xyz, ginv, feat = _f32(xyz.detach()), _f32(ginv.detach()), _f32(feat.detach())
        pose_emb = None if pose_emb is None else _f32(pose_emb.detach()).reshape(-1)
        ws = [_f32(w.detach()) for w in weights]
        M = c.pts.shape[0]
        dev = xyz.device
        d_in = PE_POS + FEAT_DIM + (0 if pose_emb is None else pose_emb.numel())
        alpha, rgb = _empty((M,), dev), _empty((M, 3), dev)
        alpha_d = _empty((M,), dev) if c.direct else None
        rgb_d = _empty((M, 3), dev) if c.direct else None
        idw = _empty((M, K_NEIGHBOURS), dev)
        if M > 0:
            out = AggOutputs()
            out.alpha, out.rgb, out.alpha_direct, out.rgb_direct, out.idw = ptr(alpha), ptr(rgb), ptr(alpha_d), ptr(rgb_d), ptr(idw)
            a = _agg_inputs(c, xyz, ginv, feat, pose_emb, M, d_in)
            w = _mlp_struct(ws)
            pk, table = packed.get(ws, d_in, feat)
            sb = lib.apn_aggregate_tc_scratch_bytes(M)
            scratch = _empty((sb,), dev, torch.uint8)
            with stage("feat_net"):
                check(lib.apn_aggregate_fwd_tc(C.byref(a), C.byref(w), ptr(pk), ptr(table), C.byref(out), int(precision),
                                               None, 0, ptr(scratch), sb, stream()), "apn_aggregate_fwd_tc")
    return alpha, rgb, alpha_d, rgb_d, idw


def _aligned_bytes(n: int, device, align: int = 1024) -> torch.Tensor:
    t = _empty((n + align,), device, torch.uint8)
    off = (-t.data_ptr()) % align
    return t[off:off + n]


class _AggregateTC(torch.autograd.Function):
    """Training aggregation on the tcgen05 tensor cores: split-fp16 forward that records a tape, tensor-core dgrad
    + wgrad backward (csrc/aggregate_tc.cu, csrc/aggregate_tc_bwd.cu).  Same contract as _Aggregate, including the
    pose embedding of the ZJU configs (d_in = 255, lib/temporalpoints.py:483-490)."""

    @staticmethod
    def forward(ctx, c: AggConst, packed: PackedDecoder, xyz, ginv, feat, pose_emb, *ws):
        lib = _lib.load()
        xyz, ginv, feat = _f32(xyz), _f32(ginv), _f32(feat)
        pose_emb = None if pose_emb is None else _f32(pose_emb).reshape(-1)
        ws = [_f32(w) for w in ws]
        M = c.pts.shape[0]
        dev = xyz.device
        d_in = PE_POS + FEAT_DIM + (0 if pose_emb is None else pose_emb.numel())
        alpha, rgb = _empty((M,), dev), _empty((M, 3), dev)
        alpha_d = _empty((M,), dev) if c.direct else None
        rgb_d = _empty((M, 3), dev) if c.direct else None
        idw = _empty((M, K_NEIGHBOURS), dev)
        h, exp_d = _empty((M, FEAT_DIM), dev), _empty((M,), dev)
        fv, v0 = _empty((M, FV_LD), dev), _empty((M, V0_DIM), dev)
        tape_bytes = lib.apn_aggregate_tc_tape_bytes(M)
        tape = _aligned_bytes(max(tape_bytes, 1), dev)
        if M > 0:
            out = AggOutputs()
            out.alpha, out.rgb, out.alpha_direct, out.rgb_direct, out.idw = ptr(alpha), ptr(rgb), ptr(alpha_d), ptr(rgb_d), ptr(idw)
            out.h, out.exp_d, out.fv, out.v0 = ptr(h), ptr(exp_d), ptr(fv), ptr(v0)
            a = _agg_inputs(c, xyz, ginv, feat, pose_emb, M, d_in)
            w = _mlp_struct(ws)
            pk, table = packed.get(ws, d_in, feat)
            with stage("feat_net"):
                check(lib.apn_aggregate_fwd_tc(C.byref(a), C.byref(w), ptr(pk), ptr(table), C.byref(out), 1, ptr(tape),
                                               tape_bytes, None, 0, stream()), "apn_aggregate_fwd_tc(train)")
        ctx.c, ctx.packed, ctx.d_in, ctx.has_pose = c, packed, d_in, pose_emb is not None
        ctx.save_for_backward(xyz, ginv, feat, pose_emb if pose_emb is not None else xyz.new_empty(0), *ws, alpha, rgb, idw, h,
                              exp_d, fv, v0, tape)
        ctx.mark_non_differentiable(idw)
        if c.direct:
            ctx.mark_non_differentiable(alpha_d, rgb_d)
        return alpha, rgb, alpha_d, rgb_d, idw

    @staticmethod
    def backward(ctx, d_alpha, d_rgb, *_):
        state = _AggregateTC.backward_prepare(ctx, d_alpha, d_rgb)
        _AggregateTC.backward_launch(state, 0)
        return state["result"]

    @staticmethod
    def backward_prepare(ctx, d_alpha, d_rgb):
        """Allocates / binds every gradient buffer and builds the argument structs; `backward_launch(state, phase)` runs the
        kernels: phase 0 = all, or 1 (through canonical_feat.grad) then 2 (the rest) for a data-parallel step that starts
        exchanging the point-feature gradient while the weight gradients are still being computed (apn_aggregate_bwd_tc_phase);
        or 3 (through tc_dgrad: d_xyz / d_ginv final) then 4 (every parameter gradient) — the two branches of the one-GPU graph."""
        lib = _lib.load()
        c = ctx.c
        sv = ctx.saved_tensors
        xyz, ginv, feat, pose_emb = sv[0:4]
        if not ctx.has_pose:
            pose_emb = None
        ws = list(sv[4:20])
        alpha, rgb, idw, h, exp_d, fv, v0, tape = sv[20:28]
        M = c.pts.shape[0]
        dev = xyz.device
        need = ctx.needs_input_grad      # (c, packed, xyz, ginv, feat, pose_emb, *ws)
        # The backward kernels ACCUMULATE into caller-provided buffers.  With DIRECT_GRAD_ACCUM (set by
        # train.GradBucket.direct_accum) a leaf parameter whose .grad already exists (a slice of the flat gradient bucket)
        # receives its gradient in place and autograd is handed None: no temporary, no memset, no AccumulateGrad add kernel.
        direct = [DIRECT_GRAD_ACCUM and n and t.is_leaf and t.grad is not None and t.grad.is_contiguous()
                  and t.grad.dtype == torch.float32 for t, n in zip([feat] + list(ws), [need[4]] + list(need[6:]))]
        want_pose = pose_emb is not None and need[5]
        wanted = [t for t, n in ((xyz, need[2]), (ginv, need[3])) if n]
        if want_pose:
            wanted.append(pose_emb)
        wanted += [t for t, dr in zip([feat] + list(ws), direct) if not dr]
        zs = _zeros_like_many(wanted) if wanted else []
        zi = iter(zs)
        d_xyz = next(zi) if need[2] else None
        d_ginv = next(zi) if need[3] else None
        d_pose = next(zi) if want_pose else None
        d_all = [t.grad if dr else next(zi) for t, dr in zip([feat] + list(ws), direct)]
        d_feat = d_all[0] if need[4] else None
        d_ws = d_all[1:]
        state = dict(M=M, keep=[])
        if M > 0:
            d_alpha = torch.zeros_like(alpha) if d_alpha is None else _f32(d_alpha)
            d_rgb = torch.zeros_like(rgb) if d_rgb is None else _f32(d_rgb)
            out = AggOutputs()
            out.alpha, out.rgb, out.idw = ptr(alpha), ptr(rgb), ptr(idw)
            out.h, out.exp_d, out.fv, out.v0 = ptr(h), ptr(exp_d), ptr(fv), ptr(v0)
            g = AggGrads()
            g.d_alpha, g.d_rgb = ptr(d_alpha), ptr(d_rgb)
            g.d_xyz, g.d_ginv, g.d_feat, g.d_pose_emb = ptr(d_xyz), ptr(d_ginv), ptr(d_feat), ptr(d_pose)
            for l in range(4):
                g.d_w[l] = ptr(d_ws[2 * l])
                g.d_b[l] = ptr(d_ws[2 * l + 1])
            (g.d_density_w, g.d_density_b, g.d_rgb_feat_w, g.d_rgb_feat_b, g.d_rgb_v0_w, g.d_rgb_v0_b, g.d_rgb_v2_w,
             g.d_rgb_v2_b) = [ptr(t) for t in d_ws[8:16]]
            a = _agg_inputs(c, xyz, ginv, feat, pose_emb, M, ctx.d_in)
            w = _mlp_struct(ws)
            pk = ctx.packed.get_bwd(ws, ctx.d_in)
            sb = lib.apn_aggregate_tc_bwd_scratch_bytes(M, xyz.shape[0])
            scratch = _aligned_bytes(sb, dev)
            state.update(a=a, w=w, pk=pk, out=out, g=g, tape=tape, scratch=scratch, sb=sb)
            state["keep"] = [d_alpha, d_rgb, scratch, pk, tape, xyz, ginv, feat, pose_emb, ws, alpha, rgb, idw, h, exp_d, fv, v0,
                             d_xyz, d_ginv, d_pose, d_all]
        state["result"] = (None, None, d_xyz, d_ginv, (None if direct[0] else d_feat), d_pose,
                           *[dw if (need[6 + i] and not direct[1 + i]) else None for i, dw in enumerate(d_ws)])
        return state

    @staticmethod
    def backward_launch(state, phase: int):
        if state["M"] <= 0:
            return
        lib = _lib.load()
        with stage("feat_net_bwd"):
            if phase == 0:
                check(lib.apn_aggregate_bwd_tc(C.byref(state["a"]), C.byref(state["w"]), ptr(state["pk"]), C.byref(state["out"]),
                                               ptr(state["tape"]), C.byref(state["g"]), ptr(state["scratch"]), state["sb"],
                                               stream()), "apn_aggregate_bwd_tc")
            else:
                check(lib.apn_aggregate_bwd_tc_phase(C.byref(state["a"]), C.byref(state["w"]), ptr(state["pk"]),
                                                     C.byref(state["out"]), ptr(state["tape"]), C.byref(state["g"]),
                                                     ptr(state["scratch"]), state["sb"], int(phase), stream()),
                      "apn_aggregate_bwd_tc_phase")


def aggregate_tc_train(c: AggConst, xyz, ginv, feat, pose_emb, weights: Sequence[torch.Tensor], packed: PackedDecoder):
    """Differentiable aggregation on the tensor cores (d_in = 191, or 255 with a pose embedding).
    -> alpha, rgb, alpha_direct, rgb_direct, idw."""
    assert len(weights) == 16
    if pose_emb is not None:
        pose_emb = pose_emb.reshape(-1)          # (1, 64) -> (64): the gradient comes back in this shape through autograd
    return _AggregateTC.apply(c, packed, xyz, ginv, feat, pose_emb, *weights)


# --------------------------------------------------------------------------------------
# K4: compositing
# --------------------------------------------------------------------------------------
# measurement hook (bench.py): a list here makes every compositing forward also return, per ray, how many samples the walk
# visited before the early stop (T < 1e-3) — the samples the kernel actually read; appended as (R,) int32 device tensors
RECORD_VISITED = None


class _Composite(torch.autograd.Function):
    """pre-mask + Alphas2Weights + post-mask + segment_coo (lib/temporalpoints.py:611-677)."""

    @staticmethod
    def forward(ctx, alpha, rgb, step_id, extra, ray_start, n_rays, thres, bg, want_depth):
        lib = _lib.load()
        alpha, rgb = _f32(alpha), _f32(rgb)
        dev = alpha.device
        M, R = alpha.shape[0], int(n_rays)
        rgb_m = _empty((R, 3), dev)
        last = _empty((R,), dev)
        depth = _empty((R,), dev) if want_depth else None
        n_extra = 0 if extra is None else extra.shape[1]
        extra = None if extra is None else _f32(extra)
        extra_m = _empty((R, n_extra), dev) if n_extra else None
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        T_save = _empty((M,), dev) if need_grad else None
        n_used = _empty((R,), dev, torch.int32) if (need_grad or RECORD_VISITED is not None) else None
        if RECORD_VISITED is not None:
            RECORD_VISITED.append(n_used)
        with stage("Alphas2Weights"):
            check(lib.apn_composite_fwd(ptr(alpha), ptr(rgb), ptr(step_id) if want_depth else None, ptr(extra), n_extra,
                                        ptr(ray_start), R, float(thres), float(bg), ptr(rgb_m), ptr(last), ptr(depth),
                                        ptr(extra_m), ptr(T_save), ptr(n_used), stream()), "apn_composite_fwd")
        ctx.save_for_backward(alpha, rgb, step_id, ray_start, T_save, n_used, last)
        ctx.meta = (R, float(thres), float(bg), want_depth)
        outs = [rgb_m, last, depth if want_depth else torch.zeros(0, device=dev),
                extra_m if n_extra else torch.zeros(0, device=dev)]
        ctx.mark_non_differentiable(outs[3])
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_rgb_m, d_last, d_depth, _d_extra):
        lib = _lib.load()
        alpha, rgb, step_id, ray_start, T_save, n_used, last = ctx.saved_tensors
        R, thres, bg, want_depth = ctx.meta
        d_alpha = torch.empty_like(alpha)
        d_rgb = torch.empty_like(rgb)
        d_rgb_m = None if d_rgb_m is None else _f32(d_rgb_m)
        d_last = None if d_last is None else _f32(d_last)
        d_depth = _f32(d_depth) if (want_depth and d_depth is not None) else None
        with stage("Alphas2Weights_bwd"):
            check(lib.apn_composite_bwd(ptr(alpha), ptr(rgb), ptr(step_id) if want_depth else None, ptr(ray_start), R, thres, bg,
                                        ptr(T_save), ptr(n_used), ptr(last), ptr(d_rgb_m), ptr(d_last), ptr(d_depth),
                                        ptr(d_alpha), ptr(d_rgb), stream()), "apn_composite_bwd")
        return d_alpha, d_rgb, None, None, None, None, None, None, None


def composite(alpha, rgb, step_id, ray_start, n_rays: int, thres: float, bg: float, extra=None, want_depth=True):
    """-> rgb_marched (R,3), alphainv_last (R), depth (R)|None, extra_marched (R,n_extra)|None."""
    rgb_m, last, depth, extra_m = _Composite.apply(alpha, rgb, step_id, extra, ray_start, n_rays, thres, bg, want_depth)
    return rgb_m, last, (depth if want_depth else None), (extra_m if extra is not None else None)


# --------------------------------------------------------------------------------------
# K4b: Adam
# --------------------------------------------------------------------------------------
def adam_step_size(step: int, beta1: float, beta2: float, lr: float) -> float:
    return float(_lib.load().apn_adam_step_size(int(step), beta1, beta2, lr))


class AdamPlan:
    """A cached descriptor table for apn_adam_multi: the device pointers stay fixed across steps (parameters, gradient
    bucket views, moments), only the step sizes change."""

    def __init__(self, entries):
        self.n = len(entries)
        self.arr = (AdamTensor * self.n)()
        self.keep = entries                      # keeps the tensors alive
        for i, (p, g, m, v, pl, ss, mode) in enumerate(entries):
            for t in (p, g, m, v):
                if t.dtype != torch.float32:
                    raise _lib.ApnError("Adam tensors must be fp32")
            self.arr[i].param, self.arr[i].grad, self.arr[i].exp_avg, self.arr[i].exp_avg_sq = ptr(p), ptr(g), ptr(m), ptr(v)
            self.arr[i].perlr = ptr(pl)
            self.arr[i].numel = p.numel()
            self.arr[i].step_size = ss
            self.arr[i].mode = mode
        self.cptr = C.cast(self.arr, C.c_void_p)

    def launch(self, step_sizes, beta1, beta2, eps):
        for i, ss in enumerate(step_sizes):
            self.arr[i].step_size = ss
        with stage("adam"):
            check(_lib.load().apn_adam_multi(self.cptr, self.n, beta1, beta2, eps, stream()), "apn_adam_multi")

    def launch_dev(self, step_sizes_dev: torch.Tensor, skip_dev: Optional[torch.Tensor], beta1, beta2, eps):
        """Step sizes from device memory (n floats) and an optional device skip word: the launch a CUDA graph can replay."""
        assert step_sizes_dev.numel() >= self.n and step_sizes_dev.dtype == torch.float32
        with stage("adam"):
            check(_lib.load().apn_adam_multi_dev(self.cptr, self.n, beta1, beta2, eps, ptr(step_sizes_dev), ptr(skip_dev), stream()),
                  "apn_adam_multi_dev")


def adam_multi(entries, beta1: float, beta2: float, eps: float) -> None:
    """entries: list of (param, grad, exp_avg, exp_avg_sq, perlr|None, step_size, mode)."""
    if not entries:
        return
    lib = _lib.load()
    arr = (AdamTensor * len(entries))()
    for i, (p, g, m, v, pl, ss, mode) in enumerate(entries):
        for t in (p, g, m, v):
            if t.dtype != torch.float32:
                raise _lib.ApnError("Adam tensors must be fp32")
        arr[i].param, arr[i].grad, arr[i].exp_avg, arr[i].exp_avg_sq = ptr(p), ptr(g), ptr(m), ptr(v)
        arr[i].perlr = ptr(pl)
        arr[i].numel = p.numel()
        arr[i].step_size = ss
        arr[i].mode = mode
    with stage("adam"):
        check(lib.apn_adam_multi(C.cast(arr, C.c_void_p), len(entries), beta1, beta2, eps, stream()), "apn_adam_multi")
