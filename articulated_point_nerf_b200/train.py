"""Stage-2 (PCD) training step with the reference's structure (run.py:574-721) on the B200 kernels, plus the
one parallel strategy this path has: rays sharded across the GPUs of one box, one NCCL all-reduce (AVG) of a
flat fp32 gradient bucket per step.  torch.distributed is plumbing only.

  create_optimizer         lib/utils.py:480-513  (groups from the `lrate_*` keys)
  STAGE2_LRATES            configs/nerf/default.py:80-92
  train_step               run.py:575,615-631,713-721 (zero_grad, forward, 200*MSE, backward, step, lr decay)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Optional

import torch
import torch.nn.functional as F

from .masked_adam import MaskedAdam

# configs/nerf/default.py:80-92 (keys without a matching attribute on the model are skipped, lib/utils.py:490)
STAGE2_LRATES = dict(rgbnet=1e-4, densitynet=1e-4, featurenet=1e-4, canonical_feat=1e-4, gammas=1e-3, weights=1e-4,
                     theta_weight=1e-4, forward_warp=1e-4, joints=1e-5, theta=1e-5, feat_net=1e-3)
WEIGHT_RENDER = 2e2          # configs/nerf/default.py:95
LRATE_DECAY = 160            # N_iters // 1000 (configs/nerf/default.py:76)


@dataclass(frozen=True)
class Regularisers:
    """Weights of the stage-2 regulariser losses that run.py:633-657 adds to the render loss in every iteration
    (defaults: configs/nerf/default.py:96-103 = configs/zju/default.py:95-102).  A zero weight switches a term off.
    `sparsity` only applies from `weight_start_iter` on (run.py:645): pass 0 before that iteration.
    The 2-D chamfer term (weight_chamfer2D, run.py:659-690) needs data (mask pixels, cameras): it enters through the
    `extra_loss` hook of FusedTrainStep (see Chamfer2D)."""
    arap: float = 5e-3
    tv: float = 1e1
    sparsity: float = 2e-1
    transformation_reg: float = 1e-1
    joint_chamfer: float = 1.0

    def any_point(self) -> bool:
        return self.arap != 0 or self.tv != 0 or self.sparsity != 0

    def any_pose(self) -> bool:
        return self.transformation_reg != 0 or self.joint_chamfer != 0


class Chamfer2D:
    """The 2-D chamfer term of run.py:659-690 as an `extra_loss` for FusedTrainStep / GraphedTrainStep: projects the warped
    cloud into B views and compares it with M mask pixels per view (get_batch_chamfer_loss, lib/temporalpoints.py:765-795;
    nearest neighbours by apn_nn1_batched).  `poses` (B,4,4), `Ks` (B,3,3) and `mask_pcd` (B,M,2) [(row, col) pixel
    coordinates] are STATIC device buffers: the caller refreshes their contents per step (`update`), so the term can live
    inside a captured graph.  `image_height` mirrors the y flip of run.py:677-678 (None = render_kwargs['inverse_y'])."""

    def __init__(self, model, poses, Ks, mask_pcd, weight: float = 5e-3, n_points: Optional[int] = 3000, image_height=None):
        self.model, self.weight, self.image_height = model, float(weight), image_height
        self.n_points = None if n_points is None else int(n_points)
        dev = mask_pcd.device
        # world->camera matrices: the reference inverts the poses inside project_point_to_image_plane (lib/utils.py:444) in
        # every iteration; torch.inverse reads an error flag back to the host, so the inversion happens here, on the host
        # copy of the (B,4,4) poses, and only the product with the points is part of the (capturable) step
        self.w2c = torch.linalg.inv(poses.detach().double().cpu()).float().to(dev)
        self.Ks, self.mask_pcd = Ks.detach().clone().float().to(dev), mask_pcd.detach().clone().float()

    def update(self, poses, Ks, mask_pcd):
        self.w2c.copy_(torch.linalg.inv(poses.detach().double().cpu()).float(), non_blocking=True)
        self.Ks.copy_(Ks, non_blocking=True)
        self.mask_pcd.copy_(mask_pcd, non_blocking=True)

    def __call__(self, xyz):
        pts = torch.matmul(xyz, self.w2c[:, :3, :3].transpose(1, 2)) + self.w2c[:, None, :3, 3]       # (B, N, 3) camera frame
        pts = torch.matmul(pts, self.Ks.transpose(1, 2))
        proj = pts[:, :, :2] / pts[:, :, 2:]                                                         # lib/utils.py:446-448
        if self.image_height is not None:
            proj = torch.cat([(self.image_height - 1) - proj[:, :, :1], proj[:, :, 1:]], dim=-1)
        proj = proj.flip(-1)
        return self.weight * self.model.get_batch_chamfer_loss(proj, self.mask_pcd, N=self.n_points, M=None)


def create_optimizer(model, lrates: Optional[Dict[str, float]] = None, global_step: int = 0,
                     lrate_decay: int = LRATE_DECAY, skip_zero_grad_fields: Iterable[str] = ()) -> MaskedAdam:
    """lib/utils.py:480-513."""
    lrates = dict(STAGE2_LRATES if lrates is None else lrates)
    decay_factor = 0.1 ** (global_step / (lrate_decay * 1000))
    groups = []
    for k, lr0 in lrates.items():
        if not hasattr(model, k):
            continue
        param = getattr(model, k)
        if param is None:
            continue
        lr = lr0 * decay_factor
        if lr > 0:
            if isinstance(param, torch.nn.Module):
                param = list(param.parameters())
            else:
                param = [param]
            groups.append({'params': param, 'name': k, 'lr': lr, 'skip_zero_grad': k in skip_zero_grad_fields})
        else:
            param.requires_grad = False
    return MaskedAdam(groups)


class GradBucket:
    """All gradients of the optimiser's parameters as views of ONE flat fp32 buffer: zeroing is one memset and the
    data-parallel exchange is an all-reduce over it.

    Layout: [early | late | status].  `early` holds the parameters named by `early` (GraphedTrainStep passes the
    decoder's: canonical_feat, feat_net, rgbnet, densitynet — ~80 % of the bytes — whose gradients are complete as soon as
    the decoder backward has run), `late` the rest (skinning weights, joints, pose network: complete after the LBS and
    pose backward).  The two parts can be reduced separately, the first one overlapping the LBS / pose backward.
    `status` (64 floats) is a side channel that travels with the `late` all-reduce (a device-side flag every rank has to
    agree on, e.g. "a rank's sample workspace overflowed: skip this update")."""

    STATUS = 64

    def __init__(self, optimizer: torch.optim.Optimizer, early=(), first=()):
        """`first` (a subset of `early`) goes to the very front: [first | rest of early | late | status]; `first_split` /
        `split` are the two boundaries.  GraphedTrainStep puts canonical_feat there: its gradient is final before the
        decoder's weight gradients are (apn_aggregate_bwd_tc_phase) and is ~80 % of the bytes."""
        params = [p for g in optimizer.param_groups for p in g['params'] if p.requires_grad]
        assert params, "no trainable parameters"
        early_ids = {id(p) for p in early}
        first_ids = {id(p) for p in first} & early_ids
        params = ([p for p in params if id(p) in first_ids] + [p for p in params if id(p) in early_ids and id(p) not in first_ids]
                  + [p for p in params if id(p) not in early_ids])
        dev = params[0].device
        offs, total, split, first_split = [], 0, 0, 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64          # 256-byte aligned slices
            if id(p) in early_ids:
                split = total
            if id(p) in first_ids:
                first_split = total
        self.split, self.total, self.first_split = split, total, first_split
        self.flat = torch.zeros(total + self.STATUS, device=dev, dtype=torch.float32)
        self.status = self.flat[total:total + self.STATUS]
        self.params = params
        self.offsets = offs
        self.attach()
        self.numel = sum(p.numel() for p in params)

    def attach(self):
        """(Re-)installs the bucket slices as the parameters' .grad."""
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view_as(p)

    def attached(self) -> bool:
        """False once something detached a gradient from the bucket, e.g. optimizer.zero_grad(set_to_none=True) (the
        reference's default, run.py:575): an all-reduce of `flat` would then exchange zeros."""
        base = self.flat.data_ptr()
        return all(p.grad is not None and p.grad.data_ptr() == base + 4 * o for p, o in zip(self.params, self.offsets))

    def direct_accum(self):
        """Context manager: inside it the decoder's backward kernels accumulate straight into the bucket slices (no
        temporaries, no AccumulateGrad adds).  Scoped to the training step: outside it autograd sees ordinary
        gradients, so torch.autograd.grad / gradcheck / hooks on the same model keep working."""
        return _DirectAccum()

    def zero(self):
        if not self.attached():
            self.attach()
        self.flat.zero_()

    def all_reduce_avg(self, group=None, part=None):
        """part None: the whole bucket (+ status); 0: the early slice; 1: the late slice + status; "first": the front of the
        early slice (canonical_feat); "early_rest": the remainder of the early slice."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if not self.attached():
                raise RuntimeError("GradBucket: a parameter's .grad no longer aliases the flat bucket (zero_grad(set_to_none="
                                   "True)?); call bucket.zero() instead of optimizer.zero_grad()")
            if part == "first":
                buf = self.flat[:self.first_split]
            elif part == "early_rest":
                buf = self.flat[self.first_split:self.split]
            else:
                buf = self.flat if part is None else (self.flat[:self.split] if part == 0 else self.flat[self.split:])
            if buf.numel() == 0:
                return
            avg = dist.get_backend(group) == "nccl"        # gloo has no AVG
            dist.all_reduce(buf, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=group)
            if not avg:
                buf.div_(dist.get_world_size(group))


class _DirectAccum:
    def __enter__(self):
        from . import ops
        self.prev = ops.DIRECT_GRAD_ACCUM
        ops.DIRECT_GRAD_ACCUM = True

    def __exit__(self, *exc):
        from . import ops
        ops.DIRECT_GRAD_ACCUM = self.prev
        return False


def decoder_parameters(model):
    """The parameters whose gradients the decoder backward completes (before the LBS / pose backward run)."""
    ps = [model.canonical_feat]
    for mod in (model.feat_net, model.rgbnet, model.densitynet):
        ps += list(mod.parameters())
    return ps


def warp_parameters(model):
    """The parameters whose gradients only the LBS / pose backward completes — and the only ones the sampling stage of the
    NEXT step reads (pose network, joints, skinning weights): the `late` slice of a pipelined bucket."""
    return [model.weights, model.theta_weight, model.joints, *model.forward_warp.parameters()]


OVERLAP_MIN_FLOATS = 8 << 20       # 32 MiB of gradients


def make_bucket(model, optimizer, overlap="pipeline") -> "GradBucket":
    """The flat gradient bucket of a data-parallel run.
    "pipeline" (default): [canonical_feat | every other parameter the warp does not own | warp parameters | status] —
        GraphedTrainStep then reduces the first two parts (82 % of the bytes at c2 / c4 sizes) on the communication stream BESIDE THE NEXT
        STEP's pose -> LBS -> grid -> k-NN chain, which only reads the warp parameters; only the small warp slice is exchanged
        inside the step.
    True: [canonical_feat | decoder MLPs | rest | status], reduced in three parts beside the backward of the same step
        (measured on 8 x B200, gpurun_out/r2l_*, r2m_*: wins at c4 = 61 MB, 2.30 vs 2.33 ms, loses at c2 = 17.6 MB, 1.03 vs
        1.01 ms: the persistent decoder kernels occupy every SM, and each extra graph boundary costs ~10 us).
    "auto": True from 32 MiB of gradients, else False.   False: one all-reduce of the whole bucket after the backward."""
    if overlap == "pipeline":
        warp = {id(p) for p in warp_parameters(model)}
        early = [p for g in optimizer.param_groups for p in g['params'] if p.requires_grad and id(p) not in warp]
        return GradBucket(optimizer, early=early, first=[model.canonical_feat])
    if overlap == "auto":
        n = sum(p.numel() for g in optimizer.param_groups for p in g['params'] if p.requires_grad)
        overlap = n >= OVERLAP_MIN_FLOATS
    if not overlap:
        return GradBucket(optimizer)
    return GradBucket(optimizer, early=decoder_parameters(model), first=[model.canonical_feat])


def shard_rays(n_rays: int, rank: int, world: int):
    """Contiguous, near-equal ray ranges per rank: [start, stop)."""
    base, rem = divmod(n_rays, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class _Ctx:
    """Stand-in for the autograd context: runs the autograd.Function bodies of ops.py outside the autograd engine."""
    __slots__ = ("needs_input_grad", "saved_tensors", "__dict__")

    def __init__(self, needs):
        self.needs_input_grad = tuple(needs)
        self.saved_tensors = ()

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors

    def mark_non_differentiable(self, *a):
        pass


class FusedTrainStep:
    """The render-loss training step as one explicit forward + backward chain over the kernels (no autograd engine, no
    AccumulateGrad, every gradient written by its kernel straight into the flat bucket).  Same arithmetic as
    `loss.backward()` through TemporalPoints.forward: it calls the very same forward/backward bodies (ops._Pose, _LBS,
    _AggregateTC, _Composite) in the order autograd would.  Covers the configuration the stage-2 loop runs in
    (time input, fused pose kernel, tensor-core decoder, with or without the pose embedding); `eligible()` says whether
    it applies.

    `sampler` (ops.StaticSampler): the sync-free variant — fixed-capacity sample arrays, every count stays on the device,
    no host read-back anywhere in the step (GraphedTrainStep captures exactly this body in CUDA graphs)."""

    def __init__(self, model, optimizer: MaskedAdam, bucket: GradBucket, regularisers: Optional[Regularisers] = None,
                 extra_loss: Optional[Callable] = None):
        """`regularisers`: the stage-2 regulariser terms (run.py:633-657), evaluated with their gradients by two kernels
        (apn_point_regularisers, apn_pose_regularisers) that add into the LBS / pose backward's incoming gradients.
        `extra_loss(xyz) -> scalar`: any further differentiable (torch) term on the warped cloud, e.g. Chamfer2D; its
        gradient w.r.t. xyz joins d_xyz before the LBS backward."""
        self.model, self.opt, self.bucket = model, optimizer, bucket
        self.reg, self.extra_loss = regularisers, extra_loss
        model._ensure_neighbourhood()
        dev = model.joints.device
        # [arap, tv, sparsity, transformation_reg, joint_chamfer, extra, render, total]
        self.loss_terms = torch.zeros(8, device=dev)
        self._nn_i32 = model.nn_i.to(torch.int32).contiguous() if regularisers is not None and regularisers.any_point() else None
        self._d_w = None

    @staticmethod
    def eligible(model) -> bool:
        fw = model.forward_warp
        # no_view_dir models hand the kernels a zero-padded COPY of the first view layer (TemporalPoints._mlp_weights): its
        # gradient has to flow back through that padding, which only the autograd path does
        return (model.decoder_train == "tc" and model.joints.is_cuda and fw.fused_pose and not model.no_view_dir
                and fw._fused_tables(model.joints.device) is not None)

    @torch.no_grad()
    def run(self, t, render_kwargs, target, sampler=None):
        """-> loss (device scalar), or None when the batch keeps no sample (dynamic mode only; the static mode never
        looks at the count: an empty batch simply produces zero gradients and the constant background loss)."""
        st = self.forward_sampling(t, render_kwargs, sampler)
        if st is None:
            return None
        return self.decode_and_backward(st, render_kwargs, target)

    def has_extra_terms(self) -> bool:
        return self.reg is not None or self.extra_loss is not None

    # ---- stage A: pose -> LBS -> grid -> ray samples + exact 8-NN
    @torch.no_grad()
    def refresh_decoder_state(self):
        """Re-derives the decoder's packed state on the CURRENT stream: forward and backward weight tiles + the per-point
        layer-0 table (ops.PackedDecoder).  None of it depends on the pose, so GraphedTrainStep runs it on a side stream
        beside the pose -> LBS -> grid -> k-NN chain; the decoder's own `get` calls then find the caches fresh."""
        from . import ops
        m = self.model
        ws = [ops._f32(w) for w in m._mlp_weights()]
        d_in = ops.PE_POS + ops.FEAT_DIM + int(m.pose_embedding_dim)
        m._packed_decoder.get(ws, d_in, ops._f32(m.canonical_feat))
        m._packed_decoder.get_bwd(ws, d_in)

    def forward_sampling(self, t, render_kwargs, sampler=None, zero_bucket: bool = True):
        from . import ops
        m, fw = self.model, self.model.forward_warp
        dev = m.joints.device
        if zero_bucket:
            self.bucket.zero()
        t_embed = ops.time_embed(t, m.time_poc)
        wb = fw._mlp_params()
        cp = _Ctx((False, False, m.joints.requires_grad, *[w.requires_grad for w in wb]))
        bone_Ts, global_t, thetas = ops._Pose.forward(cp, fw._fused_tables(dev), t_embed, m.joints, *wb)
        fw.prev_params, fw.prev_thetas, fw.prev_global_t = None, thetas, global_t
        cl = _Ctx((m.weights.requires_grad, m.theta_weight.requires_grad, True, True, False, False, False, False))
        xyz, ginv, w, bbox, _ = ops._LBS.forward(cl, m.weights, m.theta_weight, bone_Ts, global_t, m.canonical_pcd,
                                                 m._merge_rules_i32(), float(m.eps), False)
        m._last_weights = w
        warped = dict(xyz=xyz, ginv=ginv, weights=w, bbox=bbox, bone_Ts=bone_Ts, global_t=global_t, joints_rel=None)
        grid = m.build_grid(warped, 0.01)
        rays_o, rays_d = render_kwargs['rays_o'], render_kwargs['rays_d']
        R = len(rays_o)
        stepdist = float(render_kwargs['stepsize']) * float(m.voxel_size)
        near, far = float(render_kwargs['near']), float(render_kwargs['far'])
        if sampler is None:
            smp = ops.sample_and_knn(grid, rays_o, rays_d, near, far, stepdist)
            m.last_counts = dict(R=R, candidates=smp.n_candidates, M=smp.M, N=len(xyz))
            if smp.M == 0 and not self.has_extra_terms():
                return None
        else:
            smp = sampler.run(grid, rays_o, rays_d, near, far, stepdist)
        return dict(cp=cp, cl=cl, wb=wb, xyz=xyz, ginv=ginv, bone_Ts=bone_Ts, smp=smp, R=R)

    # ---- stage B: decoder -> compositing -> loss -> backward of everything (gradients land in the bucket)
    @torch.no_grad()
    def decode_and_backward(self, st, render_kwargs, target, warp_backward: bool = True, stop_after_feat: bool = False,
                            stop_after_dgrad: bool = False):
        """warp_backward=False stops after the decoder backward (every decoder gradient is then final in the bucket: a
        data-parallel caller can start reducing them) and leaves the rest to `warp_backward(st)`.
        stop_after_feat=True stops even earlier, as soon as canonical_feat.grad is final (phase 1 of the decoder backward,
        apn_aggregate_bwd_tc_phase); `decoder_backward_rest(st)` continues.  -> loss (the render loss in that case).
        stop_after_dgrad=True stops as soon as d_xyz / d_ginv are final (phase 3): the two halves that remain are independent
        — `decoder_backward_params(st)` (phase 4: every parameter gradient of the decoder) and
        `regularise_and_warp_backward(st)` (regularisers, LBS + pose backward; -> total loss) — and may run on two streams."""
        from . import ops
        from .heads import poc_fre
        m = self.model
        dev = m.joints.device
        cp, cl, wb, xyz, ginv, bone_Ts, smp, R = (st[k] for k in ("cp", "cl", "wb", "xyz", "ginv", "bone_Ts", "smp", "R"))
        c = ops.AggConst(pts=smp.pts, nn_idx=smp.nn_idx, ray_id=smp.ray_id, viewdirs=m._view_dirs(render_kwargs['viewdirs'], R),
                         canonical_alpha=m.canonical_alpha, canonical_rgbs=m.canonical_rgbs, direct_eps=m.direct_eps,
                         mean_min_distance=m._mmd_float, eps=float(m.eps), act_shift=float(m.tineuvox.act_shift),
                         interval=float(render_kwargs['stepsize']) * float(m.tineuvox.voxel_size_ratio), direct=False,
                         m_dev=smp.m_dev)
        ws = m._mlp_weights()
        # pose embedding (lib/temporalpoints.py:571-576; ZJU configs): a (1, 64) vector from the DETACHED joint offsets, so
        # gradients reach pose_embedding_net only — and only if the optimiser owns it (the reference's stage-2 configs
        # have no lrate_pose_embedding_net: the net stays at its initialisation, configs/zju/default.py:79-91)
        pose_emb, pose_graph = None, False
        if m.pose_embedding_dim > 0:
            jh = torch.cat([m.joints, torch.ones((len(m.joints), 1), device=dev)], dim=-1)
            delta = m.joints - torch.bmm(bone_Ts, jh.unsqueeze(-1)).squeeze(-1)[:, :3]
            pose_graph = any(p.grad is not None for p in m.pose_embedding_net.parameters())
            with torch.set_grad_enabled(pose_graph):
                pose_emb = m.pose_embedding_net(poc_fre(delta.detach(), m.pos_poc).view(1, -1)).reshape(-1)
        ca = _Ctx((False, False, True, True, m.canonical_feat.requires_grad, pose_graph, *[x.requires_grad for x in ws]))
        alpha, rgb, _, _, _ = ops._AggregateTC.forward(ca, c, m._packed_decoder, xyz, ginv, m.canonical_feat,
                                                       None if pose_emb is None else pose_emb.detach(), *ws)
        cc = _Ctx((True, True, False, False, False, False, False, False, False))
        rgb_m, last, _, _ = ops._Composite.forward(cc, alpha, rgb, smp.step_id, None, smp.ray_start, R, m.fast_color_thres,
                                                   float(render_kwargs['bg']), False)
        # ---- loss (run.py:617-621: img2mse weighted by weight_render) and its gradient
        loss, d_rgb_m = ops.mse_loss_grad(rgb_m, target, WEIGHT_RENDER)
        # ---- backward, in autograd's order
        d_alpha, d_rgb = ops._Composite.backward(cc, d_rgb_m, None, None, None)[:2]
        state = ops._AggregateTC.backward_prepare(ca, d_alpha, d_rgb)
        st.update(agg_state=state, ws=ws, pose_emb=pose_emb, pose_graph=pose_graph, render_loss=loss, ga=state["result"])
        if stop_after_feat or stop_after_dgrad:
            ops._AggregateTC.backward_launch(state, 1 if stop_after_feat else 3)
            return loss
        ops._AggregateTC.backward_launch(state, 0)
        return self._after_decoder(st, warp_backward)

    @torch.no_grad()
    def decoder_backward_rest(self, st, warp_backward: bool = True):
        """Second part of a decoder backward that was stopped after canonical_feat.grad (stop_after_feat=True).  -> loss."""
        from . import ops
        ops._AggregateTC.backward_launch(st["agg_state"], 2)
        return self._after_decoder(st, warp_backward)

    @torch.no_grad()
    def decoder_backward_params(self, st):
        """After stop_after_dgrad=True: every parameter gradient of the decoder (point features, feat_net, pose embedding)."""
        from . import ops
        ops._AggregateTC.backward_launch(st["agg_state"], 4)
        self._decoder_param_grads(st)

    @torch.no_grad()
    def decoder_backward_feat(self, st):
        """After stop_after_dgrad=True: canonical_feat.grad only (phase 5); independent of decoder_backward_weights."""
        from . import ops
        ops._AggregateTC.backward_launch(st["agg_state"], 5)
        self._accumulate([self.model.canonical_feat], st["ga"][4:5])

    @torch.no_grad()
    def decoder_backward_weights(self, st):
        """After stop_after_dgrad=True: the decoder's weight / pose-embedding gradients (phase 6)."""
        from . import ops
        ops._AggregateTC.backward_launch(st["agg_state"], 6)
        self._accumulate(list(st["ws"]), st["ga"][6:])
        if st["pose_graph"]:
            with torch.enable_grad():
                st["pose_emb"].backward(st["ga"][5])

    @torch.no_grad()
    def regularise_and_warp_backward(self, st):
        """After stop_after_dgrad=True: regularisers (+ extra loss) on d_xyz, then the LBS and pose backward.  -> total loss."""
        loss = st["render_loss"]
        if self.has_extra_terms():
            loss = self._regularise(st, st["ga"][2], loss)
        self.warp_backward(st)
        return loss

    def _decoder_param_grads(self, st):
        m = self.model
        ga, ws = st["ga"], st["ws"]
        self._accumulate([m.canonical_feat], ga[4:5])                      # None where the kernels wrote into .grad directly
        self._accumulate(list(ws), ga[6:])
        if st["pose_graph"]:
            with torch.enable_grad():
                st["pose_emb"].backward(ga[5])                              # accumulates into the bucket slices

    def _after_decoder(self, st, warp_backward):
        loss = st["render_loss"]
        self._decoder_param_grads(st)
        if self.has_extra_terms():
            loss = self._regularise(st, st["ga"][2], loss)
        if warp_backward:
            self.warp_backward(st)
        return loss

    def _regularise(self, st, d_xyz, render_loss):
        """Regulariser losses + gradients (run.py:633-694): d_xyz (N,3) receives the ARAP / extra-loss gradients, st gets d_w
        (merged skinning weights), d_thetas / d_global_t / d_joints for the LBS and pose backward.  -> total loss."""
        from . import ops
        m, fw, reg, lt = self.model, self.model.forward_warp, self.reg, self.loss_terms
        xyz = st["xyz"]
        lt.zero_()
        with _lib_stage("regularisers"):
            if reg is not None and reg.any_point():
                N, J = m._last_weights.shape
                if self._d_w is None or self._d_w.shape != (N, J):
                    self._d_w = torch.empty(N, J, device=xyz.device)
                _, st["d_w"] = ops.point_regularisers(xyz, m._last_weights, self._nn_i32, m.nn_distance, float(m.eps), reg.arap,
                                                      reg.tv, reg.sparsity, d_xyz, self._d_w, lt[0:3])
            if reg is not None and reg.any_pose():
                _, st["d_thetas"], st["d_gt_reg"], st["d_joints_reg"] = ops.pose_regularisers(
                    fw.prev_thetas, fw.prev_global_t, m.joints, m.skeleton_pcd if reg.joint_chamfer != 0 else None,
                    reg.transformation_reg, reg.joint_chamfer, lt[3:5])
        if self.extra_loss is not None:
            with _lib_stage("extra_loss"):
                with torch.enable_grad():
                    leaf = xyz.detach().requires_grad_(True)
                    extra = self.extra_loss(leaf)
                    (g,) = torch.autograd.grad(extra, leaf)
                d_xyz.add_(g)
                lt[5:6].copy_(extra.detach().reshape(1))
        lt[6:7].copy_(render_loss.reshape(1))
        torch.sum(lt[0:7], dim=0, keepdim=True, out=lt[7:8])
        return lt[7]

    # ---- stage C: LBS backward + pose backward (skinning weights, joints, pose network)
    @torch.no_grad()
    def warp_backward(self, st):
        from . import ops
        m = self.model
        cp, cl, wb, ga = st["cp"], st["cl"], st["wb"], st["ga"]
        if m.weights.grad is not None and m.theta_weight.grad is not None:
            cl.grad_out = dict(raw=m.weights.grad, theta=m.theta_weight.grad.reshape(1))
        gl = ops._LBS.backward(cl, ga[2], ga[3], st.get("d_w"), None, None)
        if not hasattr(cl, "grad_out"):
            self._accumulate([m.weights, m.theta_weight], gl[0:2])
        if all(p.grad is not None for p in [m.joints] + list(wb)):
            cp.grad_out = dict(wb=[p.grad for p in wb], joints=m.joints.grad)
        d_gt = gl[3]
        if st.get("d_gt_reg") is not None:          # transformation regulariser on the global translation
            d_gt = st["d_gt_reg"] if d_gt is None else d_gt.add_(st["d_gt_reg"])
        gp = ops._Pose.backward(cp, gl[2], d_gt, st.get("d_thetas"))
        if not hasattr(cp, "grad_out"):
            self._accumulate([m.joints] + list(wb), gp[2:])
        if st.get("d_joints_reg") is not None:      # joint chamfer: the pose backward has overwritten d_joints, add on top
            self._accumulate([m.joints], [st["d_joints_reg"]])

    @staticmethod
    def _accumulate(params, grads):
        for p, g in zip(params, grads):
            if g is not None and p.requires_grad and p.grad is not None:
                p.grad.add_(g.reshape(p.grad.shape))


def _lib_stage(name):
    from . import _lib
    return _lib.stage(name)


class WorkspaceOverflow(RuntimeError):
    """A graph-captured step found more candidates / kept samples than its fixed workspace holds.  The step was skipped on
    the device (no optimiser update on any rank), the workspace has been enlarged; feed the batch again."""


class GraphedTrainStep:
    """The stage-2 iteration with NO host synchronisation, replayed from CUDA graphs.

    The reference reads its sample count back to the host in every forward (lib/cuda/render_utils_kernel.cu:205-206) and
    so did the dynamic path (two `.item()` per step): the host cannot run ahead of the GPU, and every step pays ~80 launch
    latencies.  Here the sample arrays have fixed capacities (ops.StaticSampler), every kernel reads its counts from
    device memory (apn_agg_inputs.m_dev), Adam takes its step sizes from device memory (apn_adam_multi_dev), and the whole
    chain is captured once:
        graph A   zero bucket, pose, LBS, grid, samples + 8-NN, decoder, compositing, loss, full backward
        (eager)   the one NCCL all-reduce of the flat bucket — only with more than one rank
        graph B   Adam
    Per step the host copies the inputs into the static buffers, writes ~30 step sizes, and launches two graphs.

    More than one rank, bucket from make_bucket(..., "pipeline"): the exchange is software-pipelined ACROSS steps —
        main stream    graph P1 [zero warp slice, pose, LBS, grid, samples + 8-NN]   (reads only the warp parameters)
                       wait for the communication stream
                       graph P2 [decoder, compositing, loss, backward with its three branches]
                       all-reduce of the warp slice + status (small), Adam of the warp parameters
        comm stream    all-reduce of the decoder slice (82 % of the bytes at c2 / c4 sizes), Adam of those parameters, graph PP [zero the
                       decoder slice, weight tiles, per-point table] — all of it beside the NEXT step's graph P1.
    Same arithmetic as the unpipelined step: every parameter is updated before its next reader runs.

    Branches inside the graph (captured fork / join on a side stream; they also run, the same way, in the eager mode):
      * the decoder's derived state (weight tiles, per-point table) and the bucket memset do not depend on the pose: they
        run beside the pose -> LBS -> grid -> k-NN chain, whose kernels leave most SMs idle, and join before the decoder;
      * (unless the gradient exchange is split) the backward forks three ways as soon as tc_dgrad has produced d_xyz / d_ginv
        (apn_aggregate_bwd_tc_phase 3 | 5 | 6): [d_feat GEMM, Adam of canonical_feat (~80 % of the optimiser's bytes)] beside
        [point-table wgrad GEMM || tc_wgrad, Adam of the decoder's MLPs] beside [regularisers, LBS backward, pose backward
        (one 8-CTA cluster), Adam of the skinning weights / joints / pose network]; the Adam parts only with one rank (with
        more, Adam follows the all-reduce).

    Overflow: if a batch yields more samples than the workspace holds, the kernels truncate, raise a flag that travels with
    the bucket through the all-reduce (so every rank sees it) and Adam skips the update on the device.  The host notices
    on a later call (it polls, it never waits), enlarges the workspace, re-captures and raises WorkspaceOverflow; `flush()`
    waits for the steps still in flight."""

    RING = 8

    def __init__(self, model, optimizer: MaskedAdam, bucket: GradBucket, n_rays: int, render_kwargs, *, cand_cap=None,
                 m_cap=None, calibrate=None, use_graph: bool = True, packed_inputs: bool = False,
                 regularisers: Optional[Regularisers] = None, extra_loss: Optional[Callable] = None):
        from . import ops
        assert FusedTrainStep.eligible(model), "GraphedTrainStep needs the fused pose kernel and the tensor-core decoder"
        self.model, self.opt, self.bucket = model, optimizer, bucket
        self.fused = FusedTrainStep(model, optimizer, bucket, regularisers, extra_loss)
        dev = model.joints.device
        self.dev, self.R, self.use_graph = dev, int(n_rays), use_graph
        self.rk = {k: render_kwargs[k] for k in ("near", "far", "bg", "stepsize", "inverse_y", "flip_x", "flip_y") if k in render_kwargs}
        # static inputs
        self.t = torch.zeros(1, device=dev)
        self.rays_o, self.rays_d, self.viewdirs, self.target = (torch.zeros(self.R, 3, device=dev) for _ in range(4))
        self.rk.update(rays_o=self.rays_o, rays_d=self.rays_d, viewdirs=self.viewdirs)
        self.loss = torch.zeros(1, device=dev)
        # optional packed input: one (R, 12) buffer [rays_o | rays_d | viewdirs | target] that arrives with ONE host->device
        # copy; the split into the four static tensors is then part of the captured graph
        self.packed = torch.zeros(self.R, 12, device=dev) if packed_inputs else None
        self.history = []                 # (candidates, kept samples) of every finished step, filled by _poll
        self.launches_per_step = 0
        # capacities: candidates = every step of every ray between near and far (cannot overflow), unless that is absurdly
        # large; kept samples = 4x what a calibration batch keeps
        stepdist = float(self.rk['stepsize']) * float(model.voxel_size)
        worst = self.R * (int((float(self.rk['far']) - float(self.rk['near'])) / stepdist) + 3)
        if calibrate is not None and (cand_cap is None or m_cap is None):
            n_cand, n_kept = self._calibrate(calibrate)
            cand_cap = cand_cap or min(worst, max(4 * n_cand, 1 << 16))
            m_cap = m_cap or max(2 * n_kept, 1 << 14)       # grids and split-K factors of the backward are sized by this
        self.cand_cap = int(cand_cap or min(worst, 1 << 24))
        self.m_cap = int(m_cap or max(self.cand_cap // 8, 1 << 14))
        self.m_cap = (self.m_cap + 127) // 128 * 128
        self.sampler = None
        self.graphs = None
        self._plan_key = None
        self._pending = []          # (event, pinned status, step index)
        self._ring = 0
        self._steps = 0
        self.world = 1
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size()
        # the status words live behind the gradients in the bucket and travel with its (late) all-reduce
        self.status = bucket.status
        self.skip_word = self.status[:1]          # non-zero (on any rank, after the all-reduce) => Adam skips
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self._side = torch.cuda.Stream(device=dev)            # branches of the step (decoder state; parameter gradients + Adam)
        self._side2 = torch.cuda.Stream(device=dev)
        self.branches = True                                   # False: everything on one stream (A/B measurements, tests)
        self._adam_feat, self._adam_early, self._adam_late, self._ss_perm = [], [], [], []
        # pipelined exchange: needs the bucket split exactly at the warp parameters (make_bucket(..., "pipeline"))
        self.pipelined = False
        if self.world > 1 and bucket.split > 0:
            warp = {p.data_ptr() for p in warp_parameters(model)}
            self.pipelined = all((p.data_ptr() in warp) == (o >= bucket.split) for p, o in zip(bucket.params, bucket.offsets))
        self.skip_copy = torch.zeros(1, device=dev)          # the reduced skip word, for the Adam part on the comm stream
        self._pg, self._need_pp = None, True
        self._pinned = [torch.zeros(9, dtype=torch.float32).pin_memory() for _ in range(self.RING)]   # status[0:8] | loss
        self._pinned_ss = None
        self.step_sizes = None

    # ------------------------------------------------------------------------------------------
    def _calibrate(self, batch):
        """One dynamic (synchronising) sampling pass on a representative batch: -> (candidates, kept samples)."""
        from . import ops
        t, ro, rd = batch[0], batch[1], batch[2]
        m = self.model
        with torch.no_grad():
            warped = m.warp(t)
            grid = m.build_grid(warped, 0.01)
            smp = ops.sample_and_knn(grid, ro, rd, float(self.rk['near']), float(self.rk['far']),
                                     float(self.rk['stepsize']) * float(m.voxel_size))
        return smp.n_candidates, smp.M

    def reserve(self, batch):
        """Grows the fixed workspace so that `batch` = (t, rays_o, rays_d, ...) fits (one synchronising sampling pass, like the
        constructor's `calibrate`): call it for representative batches BEFORE training — e.g. the densest view — so that no
        step overflows and gets skipped.  Capacities only grow; graphs are re-captured on the next step if they did."""
        n_cand, n_kept = self._calibrate(batch)
        stepdist = float(self.rk['stepsize']) * float(self.model.voxel_size)
        worst = self.R * (int((float(self.rk['far']) - float(self.rk['near'])) / stepdist) + 3)
        cand_cap = max(self.cand_cap, min(worst, max(4 * n_cand, 1 << 16)))
        m_cap = max(self.m_cap, (max(2 * n_kept, 1 << 14) + 127) // 128 * 128)
        if cand_cap != self.cand_cap or m_cap != self.m_cap:
            self.flush_no_raise()
            self.cand_cap, self.m_cap = cand_cap, m_cap
            self.graphs, self.sampler = None, None

    def _body_a(self, split: bool = False, adam_skip=False):
        """forward + decoder backward [+ LBS / pose backward unless `split`: then only up to canonical_feat.grad].
        `adam_skip` (one rank only): None / a device word = also run the optimiser, its early part beside the LBS / pose
        backward; False = no optimiser launch here."""
        cur = torch.cuda.current_stream(self.dev)
        if self.packed is not None:
            for i, dst in enumerate((self.rays_o, self.rays_d, self.viewdirs, self.target)):
                dst.copy_(self.packed[:, 3 * i:3 * i + 3])
        packed = self.model._packed_decoder
        force = packed.force
        if self.branches:
            # branch 1: bucket memset + decoder state, beside the pose -> LBS -> grid -> k-NN chain
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.bucket.zero()
                self.fused.refresh_decoder_state()
            packed.force = False               # the decoder's own look-ups hit the caches refreshed above
        try:
            st = self.fused.forward_sampling(self.t, self.rk, self.sampler, zero_bucket=not self.branches)
            if self.branches:
                cur.wait_stream(self._side)
            self._body_status()               # flags are final after sampling; the bucket (and its status words) is zeroed
            inline_adam = adam_skip is not False and not split
            fork = self.branches and not split       # (the split mode's graphs end where the all-reduces start)
            loss = self.fused.decode_and_backward(st, self.rk, self.target, warp_backward=not split and not inline_adam,
                                                  stop_after_feat=split, stop_after_dgrad=fork)
        finally:
            packed.force = force
        self._st = st
        if split:
            return
        if fork:
            loss = self._backward_branches(st, cur, inline_adam, adam_skip)
        elif inline_adam:
            self.fused.warp_backward(st)
            self._body_b(adam_skip)
        self.loss.copy_(loss.reshape(1))

    def _backward_branches(self, st, cur, inline_adam, adam_skip):
        """After decode_and_backward(stop_after_dgrad=True), i.e. as soon as tc_dgrad has produced d_xyz / d_ginv: the decoder's
        parameter gradients (+ the Adam update of those parameters) beside [regularisers, LBS + pose backward (+ Adam of the
        rest)].  -> total loss."""
        self._side.wait_stream(cur)
        self._side2.wait_stream(cur)
        with torch.cuda.stream(self._side):          # point features: d_feat GEMM, then their Adam update (~80 % of its bytes)
            self.fused.decoder_backward_feat(st)
            if inline_adam:
                self._launch_adam(self._adam_feat, adam_skip)
        with torch.cuda.stream(self._side2):         # decoder weights: point-table wgrad GEMM || tc_wgrad, then their Adam
            self.fused.decoder_backward_weights(st)
            if inline_adam:
                self._launch_adam(self._adam_early, adam_skip)
        loss = self.fused.regularise_and_warp_backward(st)
        if inline_adam:                              # (more than one rank: Adam follows the all-reduce)
            self._launch_adam(self._adam_late, adam_skip)
        cur.wait_stream(self._side)
        cur.wait_stream(self._side2)
        return loss

    # ---- pipelined exchange (more than one rank): P1 | PP | P2, see the class docstring
    def _body_p1(self):
        if self.packed is not None:
            for i, dst in enumerate((self.rays_o, self.rays_d, self.viewdirs, self.target)):
                dst.copy_(self.packed[:, 3 * i:3 * i + 3])
        if not self.bucket.attached():
            self.bucket.attach()
        self.bucket.flat[self.bucket.split:].zero_()                   # warp slice + status words
        self._st = self.fused.forward_sampling(self.t, self.rk, self.sampler, zero_bucket=False)
        self._body_status()

    def _body_pp(self):
        self.bucket.flat[:self.bucket.split].zero_()                   # decoder slice
        self.fused.refresh_decoder_state()

    def _body_p2(self):
        cur = torch.cuda.current_stream(self.dev)
        packed = self.model._packed_decoder
        force, packed.force = packed.force, False                       # PP has refreshed the decoder state
        try:
            if self.branches:
                self.fused.decode_and_backward(self._st, self.rk, self.target, stop_after_dgrad=True)
                loss = self._backward_branches(self._st, cur, False, None)
            else:
                loss = self.fused.decode_and_backward(self._st, self.rk, self.target)
        finally:
            packed.force = force
        self.loss.copy_(loss.reshape(1))

    def _body_a2(self):
        """rest of the decoder backward (weight gradients) + regularisers"""
        loss = self.fused.decoder_backward_rest(self._st, warp_backward=False)
        self.loss.copy_(loss.reshape(1))

    def _body_a3(self):
        self.fused.warp_backward(self._st)

    def _body_status(self):
        self.status[:1].copy_(self.sampler.counts[2:3])            # int flags -> float status word (0.0 = clean)
        self.status[1:6].copy_(self.sampler.counts[0:5])           # counts, for the host's bookkeeping

    def _launch_adam(self, parts, skip=None):
        for cls, ap, off in parts:
            ap.launch_dev(self.step_sizes[off:off + ap.n], self.skip_word if skip is None else skip, *cls)

    def _body_b(self, skip=None):
        self._launch_adam(self._adam_feat, skip)
        self._launch_adam(self._adam_early, skip)
        self._launch_adam(self._adam_late, skip)

    def _plan_adam(self, launches):
        """Splits every (betas, eps) class of the optimiser's launch plan by the branch of the backward that completes a
        parameter's gradient: `feat` (canonical_feat), `late` (what the warp backward writes: skinning weights, theta_weight,
        joints, pose network), `early` (the rest: the decoder's MLPs, ...), as separate descriptor tables; `_ss_perm[k]` =
        index into the plan-ordered step sizes of the k-th slot of the device step-size vector.  With more than one rank
        Adam follows the all-reduce in one piece (everything is `late`) unless the exchange is pipelined: then `late` follows
        the all-reduce of the warp slice on the main stream, `feat` + `early` that of the decoder slice on the other."""
        from . import ops
        m = self.model
        late_ptrs = {p.data_ptr() for p in warp_parameters(m)}
        feat_ptr = m.canonical_feat.data_ptr()
        self._adam_feat, self._adam_early, self._adam_late, self._ss_perm = [], [], [], []
        base = 0
        for cls, ap, sizes in launches:
            one_piece = (self.world > 1 and not self.pipelined) or (self.world == 1 and not self.branches)
            label = [2 if (one_piece or e[0].data_ptr() in late_ptrs) else (0 if e[0].data_ptr() == feat_ptr else 1)
                     for e in ap.keep]
            if len(set(label)) == 1:
                (self._adam_feat, self._adam_early, self._adam_late)[label[0]].append((cls, ap, len(self._ss_perm)))
                self._ss_perm += [base + i for i in range(ap.n)]
            else:
                for part, dst in enumerate((self._adam_feat, self._adam_early, self._adam_late)):
                    idx = [i for i, l in enumerate(label) if l == part]
                    if idx:
                        dst.append((cls, ops.AdamPlan([ap.keep[i] for i in idx]), len(self._ss_perm)))
                        self._ss_perm += [base + i for i in idx]
            base += len(sizes)

    def _capture(self, launches):
        from . import ops
        dev = self.dev
        self.sampler = ops.StaticSampler(self.R, self.cand_cap, self.m_cap, dev)
        self._plan_adam(launches)
        n = len(self._ss_perm)
        self.step_sizes = torch.zeros(max(n, 1), device=dev)
        self._pinned_ss = [torch.zeros(max(n, 1), dtype=torch.float32).pin_memory() for _ in range(self.RING)]
        self._ss_events = [None] * self.RING
        packed = self.model._packed_decoder
        self.graphs, self._pg, self._need_pp = None, None, True
        if not self.use_graph:
            return
        if self.pipelined:
            return self._capture_pipelined()
        one = self.world == 1
        # warm-up on a side stream (allocator + lazy module state), then capture
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s), self.bucket.direct_accum():
            packed.force = True
            warm_skip = torch.ones(1, device=dev)     # Adam skips during warm-up: parameters and moments stay untouched
            from . import _lib
            for _ in range(2):
                n0 = _lib.launch_count()
                if one:
                    self._body_a(adam_skip=warm_skip)
                else:
                    self._body_a()
                    self._body_b(warm_skip)
                self.launches_per_step = _lib.launch_count() - n0     # our kernels in one replayed step
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        ga, ga2, ga3, gb = torch.cuda.CUDAGraph(), None, None, None
        with self.bucket.direct_accum():
            if one:                                  # no collective between the backward and Adam: ONE graph per step
                with torch.cuda.graph(ga):
                    self._body_a(adam_skip=None)
            else:
                # four graphs around the three all-reduces: [forward + decoder backward up to canonical_feat.grad] | that
                # gradient (~80 % of the bytes) reduces on the communication stream while [decoder weight gradients +
                # regularisers] run | the decoder's weight gradients reduce while [LBS + pose backward] runs | late slice +
                # status reduce | [Adam]
                split = self.bucket.split > 0
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga):
                    self._body_a(split=split)
                if split:
                    ga2, ga3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(ga2, pool=ga.pool()):
                        self._body_a2()
                    with torch.cuda.graph(ga3, pool=ga.pool()):
                        self._body_a3()
                with torch.cuda.graph(gb, pool=ga.pool()):
                    self._body_b()
        packed.force = False
        self.graphs = (ga, ga2, ga3, gb)

    def _capture_pipelined(self):
        from . import _lib
        dev, packed = self.dev, self.model._packed_decoder
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s), self.bucket.direct_accum():
            packed.force = True
            warm_skip = torch.ones(1, device=dev)     # Adam skips during warm-up: parameters and moments stay untouched
            for _ in range(2):
                n0 = _lib.launch_count()
                self._body_p1()
                self._body_pp()
                self._body_p2()
                self._body_b(warm_skip)
                self.launches_per_step = _lib.launch_count() - n0
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        g1, gp, g2, gbl, gbe = (torch.cuda.CUDAGraph() for _ in range(5))
        with self.bucket.direct_accum():
            with torch.cuda.graph(g1):
                self._body_p1()
            with torch.cuda.graph(gp):                # (allocates nothing: its own pool is fine beside P1 on another stream)
                self._body_pp()
            with torch.cuda.graph(g2, pool=g1.pool()):
                self._body_p2()
            with torch.cuda.graph(gbl):
                self._launch_adam(self._adam_late)
            with torch.cuda.graph(gbe):
                self._launch_adam(self._adam_feat, self.skip_copy)
                self._launch_adam(self._adam_early, self.skip_copy)
        packed.force = False
        self._pg = (g1, gp, g2, gbl, gbe)

    def _step_pipelined(self, ss, params):
        cur, comm, pg = torch.cuda.current_stream(self.dev), self.comm_stream, self._pg
        pg[0].replay() if pg else self._body_p1()
        if self._need_pp:                             # first step after a (re-)capture: nothing has prepared the decoder state yet
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                pg[1].replay() if pg else self._body_pp()
            self._need_pp = False
        cur.wait_stream(comm)                         # Adam of the decoder slice (previous step) + PP have run
        self.step_sizes.copy_(ss, non_blocking=True)  # (only now: that Adam read the previous step's sizes)
        pg[2].replay() if pg else self._body_p2()
        with _lib_stage("allreduce_late"):
            self.bucket.all_reduce_avg(part=1)        # warp slice + status: issued FIRST, the only exchange inside the step
        self.skip_copy.copy_(self.status[:1])
        comm.wait_stream(cur)
        if params:                                    # (host-side bookkeeping; before PP keys its caches on the versions)
            torch.autograd.graph.increment_version(params)
        with torch.cuda.stream(comm):
            with _lib_stage("allreduce_early"):
                self.bucket.all_reduce_avg(part=0)    # decoder slice: beside the next step's P1
            if pg:
                pg[4].replay()
                pg[1].replay()
            else:
                self._launch_adam(self._adam_feat, self.skip_copy)
                self._launch_adam(self._adam_early, self.skip_copy)
                self._body_pp()
        pg[3].replay() if pg else self._launch_adam(self._adam_late)

    # ------------------------------------------------------------------------------------------
    POLL_LAG = 2

    def _poll(self, wait: bool = False):
        """Looks at the status of finished steps (never waits unless `wait`).  With more than one rank the ranks must notice
        an overflowed step in the SAME call (one that raises issues no collective while the others would): a step is looked
        at exactly POLL_LAG calls later — its status words are all-reduced, so every rank reads the same flag, and a step
        that old has long finished, so the wait is free."""
        overflow = None
        keep = []
        for ev, pin, idx in self._pending:
            due = wait or (self.world > 1 and idx <= self._steps - self.POLL_LAG)
            if due:
                ev.synchronize()
            if due or (self.world == 1 and ev.query()):
                flags = pin[0].item()
                self.last_counts = dict(R=self.R, candidates=int(pin[1].item()), M=int(pin[2].item()),     # mean over ranks
                                        N=len(self.model.canonical_pcd))
                self.model.last_counts = self.last_counts
                self.history.append((self.last_counts["candidates"], self.last_counts["M"]))
                if flags != 0.0 and overflow is None:
                    overflow = (idx, int(pin[4].item()) + 1, int(pin[5].item()) + 1)       # (rank-averaged) counts found
            else:
                keep.append((ev, pin, idx))
        self._pending = keep
        if overflow is not None:
            idx, n_cand, n_kept = overflow
            self.flush_no_raise()
            # the averaged counts under-estimate the worst rank: at least double what there was
            self.cand_cap = max(2 * self.cand_cap if n_cand > self.cand_cap // 2 else self.cand_cap, 2 * n_cand)
            self.m_cap = (max(2 * self.m_cap if n_kept > self.m_cap // 2 else self.m_cap, 2 * n_kept) + 127) // 128 * 128
            self.graphs, self.sampler = None, None
            self.opt.undo_step_count()
            raise WorkspaceOverflow(f"step {idx}: {n_cand} candidates / {n_kept} kept samples exceeded the workspace; the step was "
                                    f"skipped on the device, capacities are now {self.cand_cap} / {self.m_cap}")

    def loss_reader(self):
        """-> a callable that returns THIS step's loss as a Python float: it waits for the step's own completion event and
        reads the value from pinned host memory (the copy was enqueued with the step).  Calling it one step late (after the
        next step has been enqueued) reads every step's loss without ever draining the GPU."""
        done, pin = self._last_done

        def read():
            done.synchronize()
            return float(pin[8])
        return read

    def flush_no_raise(self):
        torch.cuda.synchronize(self.dev)
        self._pending = []

    def sync_parameters(self):
        """Pipelined exchange only: the decoder's parameters of the last step are updated on the communication stream, beside
        whatever the main stream does next.  Anything OTHER than the next step() that reads them on the current stream (a
        validation render, a checkpoint) calls this first: the current stream then waits for that update (no host wait)."""
        if self.pipelined:
            torch.cuda.current_stream(self.dev).wait_stream(self.comm_stream)

    def flush(self):
        """Waits for the steps in flight (on every stream) and raises WorkspaceOverflow if one of them was skipped."""
        self.sync_parameters()
        if self.pipelined:
            self.comm_stream.synchronize()
        self._poll(wait=True)

    @torch.no_grad()
    def step_packed(self, t, packed, decay_factor: float = 1.0):
        """One iteration from a packed (R, 12) [rays_o | rays_d | viewdirs | target] buffer (host-pinned or device)."""
        assert self.packed is not None, "construct with packed_inputs=True"
        self.packed.copy_(packed, non_blocking=True)
        return self.step(t, self.rays_o, self.rays_d, self.viewdirs, self.target, decay_factor)

    @torch.no_grad()
    def step(self, t, rays_o, rays_d, viewdirs, target, decay_factor: float = 1.0):
        """One iteration.  -> loss (a (1,) device tensor that the NEXT call overwrites)."""
        self._poll()
        for dst, src in ((self.t, t), (self.rays_o, rays_o), (self.rays_d, rays_d), (self.viewdirs, viewdirs), (self.target, target)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src.reshape(dst.shape), non_blocking=True)
        launches, params = self.opt.prepare_step()
        if self.sampler is None or self._plan_key != self.opt._plan_key:
            self._capture(launches)
            self._plan_key = self.opt._plan_key
        slot = self._ring
        self._ring = (self._ring + 1) % self.RING
        if self._ss_events[slot] is not None:
            self._ss_events[slot].synchronize()                 # the copy that last used this pinned slot has run
        ss = self._pinned_ss[slot]
        flat_sizes = [v for _, _, sizes in launches for v in sizes]
        for k, src in enumerate(self._ss_perm):
            ss[k] = flat_sizes[src]
        if not self.pipelined:
            self.step_sizes.copy_(ss, non_blocking=True)
        with self.bucket.direct_accum():
            split = self.world > 1 and self.bucket.split > 0
            if self.pipelined:
                self._step_pipelined(ss, params)
                params = None
            elif self.graphs is not None:
                self.graphs[0].replay()
            else:
                self._body_a(split=split, adam_skip=None if self.world == 1 else False)
            if self.world > 1 and not self.pipelined:
                cur = torch.cuda.current_stream(self.dev)
                if split:
                    # canonical_feat.grad (~80 % of the bucket) is final: reduce it on the communication stream while the
                    # decoder's weight gradients are computed here; then those, beside the LBS / pose backward
                    self.comm_stream.wait_stream(cur)
                    with torch.cuda.stream(self.comm_stream), _lib_stage("allreduce_feat"):
                        self.bucket.all_reduce_avg(part="first")
                    if self.graphs is not None:
                        self.graphs[1].replay()
                    else:
                        self._body_a2()
                    self.comm_stream.wait_stream(cur)
                    with torch.cuda.stream(self.comm_stream), _lib_stage("allreduce_early"):
                        self.bucket.all_reduce_avg(part="early_rest")
                    if self.graphs is not None:
                        self.graphs[2].replay()
                    else:
                        self._body_a3()
                    with _lib_stage("allreduce_late"):
                        self.bucket.all_reduce_avg(part=1)
                    cur.wait_stream(self.comm_stream)
                else:
                    with _lib_stage("allreduce"):
                        self.bucket.all_reduce_avg()
            if self.pipelined:
                pass
            elif self.graphs is None:
                if self.world > 1:
                    self._body_b()
            elif self.graphs[3] is not None:
                self.graphs[3].replay()
        ev = torch.cuda.Event()
        ev.record()
        self._ss_events[slot] = ev                     # the copy out of this pinned step-size slot has run
        pin = self._pinned[slot]
        pin[:8].copy_(self.status[:8], non_blocking=True)
        pin[8:9].copy_(self.loss, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        self._pending.append((done, pin, self._steps))
        self._last_done = (done, pin)
        self._steps += 1
        if params:
            torch.autograd.graph.increment_version(params)
        if decay_factor != 1.0:
            for g in self.opt.param_groups:                  # run.py:718-721
                g['lr'] = g['lr'] * decay_factor
        return self.loss


def _finish_step(optimizer, bucket, decay_factor):
    bucket.all_reduce_avg()
    optimizer.step()
    if decay_factor != 1.0:
        for g in optimizer.param_groups:                  # run.py:718-721
            g['lr'] = g['lr'] * decay_factor


def regulariser_losses(model, t_hat_pcd, reg: Regularisers):
    """The regulariser terms of run.py:633-657 through the model's own loss getters (plain torch expressions + autograd):
    what the kernels of FusedTrainStep._regularise replace, kept as the autograd fallback and as their test reference.
    Call after a forward (reads model._last_weights and forward_warp.prev_thetas / prev_global_t)."""
    loss = 0
    if reg.arap != 0:
        loss = loss + reg.arap * model.get_arap_loss(t_hat_pcd)
    if reg.tv != 0:
        loss = loss + reg.tv * model.get_neighbour_weight_tv_loss()
    if reg.sparsity != 0:
        loss = loss + reg.sparsity * model.get_weight_sparsity_loss()
    if reg.transformation_reg != 0:
        loss = loss + reg.transformation_reg * model.get_transformation_regularisation_loss()
    if reg.joint_chamfer != 0:
        loss = loss + reg.joint_chamfer * model.get_joint_chamfer_loss()
    return loss


def train_step(model, optimizer: MaskedAdam, bucket: GradBucket, t, render_kwargs, target, *, decay_factor: float = 1.0,
               fused: bool = True, regularisers: Optional[Regularisers] = None, extra_loss: Optional[Callable] = None):
    """One stage-2 iteration on one rank's ray batch.  Returns the (device) loss tensor.

    A batch that keeps no sample (every ray of this rank's shard misses the cloud: the reference's NoPointsException
    path, lib/temporalpoints.py:598-609) has a constant render loss: its gradient is zero, so the bucket stays zeroed —
    but the rank still takes part in the all-reduce and the optimiser step, otherwise the other ranks of a sharded
    batch would wait in NCCL for ever."""
    with bucket.direct_accum():
        if fused and FusedTrainStep.eligible(model):
            fs = getattr(bucket, "_fused_step", None)
            if fs is None or fs.model is not model or fs.reg != regularisers or fs.extra_loss is not extra_loss:
                fs = bucket._fused_step = FusedTrainStep(model, optimizer, bucket, regularisers, extra_loss)
            loss = fs.run(t, render_kwargs, target)
            if loss is None:
                bg = float(render_kwargs['bg'])
                loss = WEIGHT_RENDER * F.mse_loss(torch.full_like(target, bg), target)
            _finish_step(optimizer, bucket, decay_factor)
            return loss
        bucket.zero()
        res = model(t, False, render_kwargs, render_pcd_direct=False)
        loss = WEIGHT_RENDER * F.mse_loss(res['rgb_marched'], target)
        if regularisers is not None:
            loss = loss + regulariser_losses(model, res['t_hat_pcd'], regularisers)
        if extra_loss is not None:
            loss = loss + extra_loss(res['t_hat_pcd'])
        if loss.requires_grad:
            loss.backward()
        _finish_step(optimizer, bucket, decay_factor)
        return loss
