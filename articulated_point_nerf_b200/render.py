"""Whole-frame render entry points: the callers of the hot path.

`render_viewpoints` / `render_repose` keep the call signatures and return values of `run.py:82-260` / `run.py:262-340`
(the reference renders a frame as 8192-ray chunks and re-warps the cloud for every chunk, `run.py:136-166,283-298`).
Here a frame is ONE pass of the hot path:

  * the pose chain, LBS and the grid run once per distinct pose (`PoseCache`: consecutive views of the same time step
    or the same `rot_params` — WIM / ZJU multi-view sets — reuse the warped cloud and its grid);
  * all rays of the frame go through sampling / k-NN / decoder / compositing in one call (`chunk_rays` bounds the
    number of rays per call for very large frames; results are identical either way because rays are independent);
  * frames are assembled on the device and leave it with one D2H copy per frame.

Image metrics other than PSNR, PNG writing and the skeleton overlay (`cv2`) belong to the reference's CLI, not to the
hot path: PSNR is computed here, the others raise if requested.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .temporalpoints import TemporalPoints


def tile_pixels(H: int, W: int, rank: int, world: int, tile: int = 16) -> torch.Tensor:
    """Row-major pixel indices of the `tile` x `tile` image tiles that rank `rank` of `world` owns (tiles dealt round-robin
    in raster order: only ~10-40 % of the rays of a frame hit the cloud, contiguous bands would be badly balanced;
    SURVEY.md §8(e)).  CPU int32 tensor, sorted."""
    ty, tx = (H + tile - 1) // tile, (W + tile - 1) // tile
    owner = (torch.arange(ty * tx) % world).reshape(ty, tx)
    mask = owner.repeat_interleave(tile, 0).repeat_interleave(tile, 1)[:H, :W] == rank
    return torch.nonzero(mask.reshape(-1)).reshape(-1).to(torch.int32)


class PoseCache:
    """warp() + build_grid() once per distinct pose."""

    def __init__(self, model: TemporalPoints, query_radius: float = 0.01):
        self.model, self.query_radius = model, query_radius
        self.key, self.warped, self.grid = None, None, None
        self.hits = self.misses = 0

    @staticmethod
    def _key(t, rot_params):
        x = t if rot_params is None else rot_params
        x = x.detach().reshape(-1).float().cpu()
        return ("t" if rot_params is None else "rot", tuple(x.tolist()))

    def get(self, t=None, rot_params=None):
        k = self._key(t, rot_params)
        if k != self.key:
            self.warped = self.model.warp(t, rot_params)
            self.grid = self.model.build_grid(self.warped, self.query_radius)
            self.key = k
            self.misses += 1
        else:
            self.hits += 1
        return self.warped, self.grid


def _scale_cameras(HW, Ks, render_factor):
    if render_factor == 0:
        return HW, Ks
    HW = np.copy(HW) // render_factor
    Ks = torch.as_tensor(Ks).clone()
    Ks[:, :2, :3] = Ks[:, :2, :3] // render_factor
    return HW, Ks


def _render_frame(model, cache, H, W, K, c2w, render_kwargs, *, t=None, rot_params=None, render_pcd_direct=False,
                  fixed_viewdirs=None, chunk_rays=None, inverse_y=False, flip_x=False, flip_y=False, Ks_i=None, pixel_ids=None):
    """One frame (or, with `pixel_ids`, this rank's pixels of it) -> dict of (H, W, C) device tensors (unowned pixels 0).
    The rays are produced ON THE DEVICE from (K, c2w) (ops.rays_of_a_view: lib/tineuvox.py:675-738 in one launch): a frame
    costs ~100 bytes of host->device traffic instead of 36 bytes per pixel."""
    dev = model.device
    rays_o, rays_d, viewdirs = ops.rays_of_a_view(H, W, K.to(torch.float32), c2w, dev, inverse_y=inverse_y, flip_x=flip_x,
                                                  flip_y=flip_y, pixel_ids=pixel_ids)
    if fixed_viewdirs is not None:
        viewdirs = torch.as_tensor(fixed_viewdirs).reshape(-1, 3).to(dev).contiguous()
        if pixel_ids is not None:
            viewdirs = viewdirs[pixel_ids.to(dev).long()].contiguous()
    warped, grid = cache.get(t=t, rot_params=rot_params)
    R = rays_o.shape[0]
    step = R if not chunk_rays else int(chunk_rays)
    outs = []
    for s in range(0, R, step):
        rk = dict(render_kwargs, rays_o=rays_o[s:s + step], rays_d=rays_d[s:s + step], viewdirs=viewdirs[s:s + step])
        out = model(t, render_depth=True, render_kwargs=rk, render_weights=True, rot_params=rot_params,
                    render_pcd_direct=render_pcd_direct, poses=c2w[None].to(dev), Ks=(K if Ks_i is None else Ks_i)[None].to(dev),
                    get_skeleton=True, warped=warped, grid=grid)
        if render_pcd_direct:
            out['rgb_marched'] = out['rgb_marched_direct']
        outs.append(out)
    cat = {k: (outs[0][k] if len(outs) == 1 else torch.cat([o[k] for o in outs])).reshape(R, -1)
           for k in ('rgb_marched', 'depth', 'weights')}
    if pixel_ids is None:
        cat = {k: v.reshape(H, W, -1) for k, v in cat.items()}
    else:                                  # scatter this rank's pixels into a full (zero) frame
        ids = pixel_ids.to(dev).long()
        cat = {k: torch.zeros(H * W, v.shape[-1], device=dev).index_copy_(0, ids, v).reshape(H, W, -1) for k, v in cat.items()}
    return cat, outs[0]['joints'], outs[0]['bones']


def _finish(rgbs, depths, weights, joints, gt_imgs, render_factor, eval_psnr, other_metrics, savedir):
    if other_metrics:
        raise NotImplementedError("SSIM / LPIPS belong to the reference's evaluation CLI (lib/utils.py), not to the render path")
    if savedir is not None:
        raise NotImplementedError("image writing belongs to the reference's CLI (imageio / cv2 are not part of this path)")
    psnrs = []
    if gt_imgs is not None and render_factor == 0 and eval_psnr:
        for i, rgb in enumerate(rgbs):
            psnrs.append(-10. * np.log10(np.mean(np.square(rgb - np.asarray(gt_imgs[i])))))
        print('Testing psnr', np.mean(psnrs), '(avg)')
    return np.array(rgbs), np.array(depths), np.array(weights), psnrs


@torch.no_grad()
def render_viewpoints(model, render_poses, HW, Ks, ndc, render_kwargs, gt_imgs=None, savedir=None, test_times=None,
                      render_factor=0, eval_psnr=False, eval_ssim=False, eval_lpips_alex=False, eval_lpips_vgg=False,
                      inverse_y=False, flip_x=False, flip_y=False, batch_size=None, verbose=True, render_pcd_direct=False,
                      render_flow=False, fixed_viewdirs=None, return_joints=False, rank: int = 0, world: int = 1,
                      shard: str = "views", gather: bool = True):
    """run.py:82-260.  -> rgbs (V,H,W,3), depths (V,H,W,1), weights (V,H,W,3), flows (empty) as numpy arrays.
    `batch_size=None` renders each frame in one pass; an integer reproduces the reference's ray chunking (same result).

    Multi-GPU (not in the reference): with `world` > 1 this rank renders its share — whole frames
    (`shard="views"`: frame i belongs to rank i % world, e.g. one 1024^2 view per GPU) or its 16x16 tiles of every frame
    (`shard="tiles"`, round-robin: one large frame split over the GPUs).  Rays are independent given the warped cloud,
    every rank warps the full cloud itself, so the render needs NO collective; `gather=True` (default) finally sums the
    disjoint shares with one all-reduce so that every rank returns complete frames (skipped when torch.distributed is
    not initialised: the caller then holds this rank's share, unowned pixels / frames zero)."""
    assert len(render_poses) == len(HW) and len(HW) == len(Ks)
    assert isinstance(model, TemporalPoints), "this entry point drives the point-cloud model"
    assert not ndc, "the PCD path never uses NDC rays (configs keep ndc=False)"
    assert not render_flow, "scene flow is not part of the PCD hot path"
    HW, Ks = _scale_cameras(HW, Ks, render_factor)
    assert shard in ("views", "tiles") and 0 <= rank < max(world, 1)
    cache = PoseCache(model)
    rgbs, depths, weights, joints = [], [], [], {}
    bones = None
    tiles = {}
    frames_dev = []
    for i, c2w in enumerate(render_poses):
        H, W = int(HW[i][0]), int(HW[i][1])
        if world > 1 and shard == "views" and i % world != rank:
            frames_dev.append(torch.zeros(H, W, 7, device=model.device))
            continue
        pix = None
        if world > 1 and shard == "tiles":
            if (H, W) not in tiles:
                tiles[(H, W)] = tile_pixels(H, W, rank, world).to(model.device)
            pix = tiles[(H, W)]
        t = torch.as_tensor(test_times[i], dtype=torch.float32, device=model.device).reshape(1)
        frame, jt, bn = _render_frame(model, cache, H, W, torch.as_tensor(Ks[i]), torch.as_tensor(c2w), render_kwargs, t=t,
                                      render_pcd_direct=render_pcd_direct, fixed_viewdirs=fixed_viewdirs, chunk_rays=batch_size,
                                      inverse_y=inverse_y, flip_x=flip_x, flip_y=flip_y, pixel_ids=pix)
        if jt is not None and i not in joints:
            jt = jt.clone()
            if not render_kwargs['inverse_y']:
                jt[:, :, 0] = (int(HW[0][0]) - 1) - jt[:, :, 0]
            joints[i] = jt[0].cpu().numpy()
            bones = bn
        frames_dev.append(torch.cat([frame['rgb_marched'], frame['depth'], frame['weights']], dim=-1))
    if world > 1 and gather:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            for f in frames_dev:               # disjoint shares: the sum IS the assembled frame
                dist.all_reduce(f, op=dist.ReduceOp.SUM)
    for f in frames_dev:
        host = f.cpu().numpy()                                                                  # one D2H per frame
        rgbs.append(host[..., 0:3])
        depths.append(host[..., 3:4])
        weights.append(host[..., 4:7])
    rgbs, depths, weights, _ = _finish(rgbs, depths, weights, joints, gt_imgs, render_factor, eval_psnr,
                                       eval_ssim or eval_lpips_alex or eval_lpips_vgg, savedir)
    if verbose:
        print(f'render_viewpoints: {len(rgbs)} frames, {cache.misses} warps ({cache.hits} reused)')
    flows = np.array([])
    if return_joints:
        return rgbs, depths, weights, flows, [joints[i] for i in sorted(joints)], bones
    return rgbs, depths, weights, flows


@torch.no_grad()
def render_repose(rot_params, render_poses, HW, Ks, ndc, model, render_kwargs, gt_imgs=None, savedir=None, render_factor=0,
                  eval_psnr=False, eval_ssim=False, eval_lpips_alex=False, eval_lpips_vgg=False, inverse_y=False,
                  flip_x=False, flip_y=False, batch_size=None):
    """run.py:262-340: one frame per (rot_params[i], render_poses[i]).  -> rgbs, depths, weights (numpy)."""
    assert len(render_poses) == len(HW) and len(HW) == len(Ks)
    assert isinstance(model, TemporalPoints)
    assert not ndc
    HW, Ks = _scale_cameras(HW, Ks, render_factor)
    cache = PoseCache(model)
    rgbs, depths, weights = [], [], []
    for i, c2w in enumerate(render_poses):
        H, W = int(HW[i][0]), int(HW[i][1])
        rp = torch.as_tensor(rot_params[i], dtype=torch.float32, device=model.device)
        frame, _, _ = _render_frame(model, cache, H, W, torch.as_tensor(Ks[i]), torch.as_tensor(c2w), render_kwargs, rot_params=rp,
                                    chunk_rays=batch_size, inverse_y=inverse_y, flip_x=flip_x, flip_y=flip_y)
        host = torch.cat([frame['rgb_marched'], frame['depth'], frame['weights']], dim=-1).cpu().numpy()
        rgbs.append(host[..., 0:3])
        depths.append(host[..., 3:4])
        weights.append(host[..., 4:7])
    rgbs, depths, weights, _ = _finish(rgbs, depths, weights, {}, gt_imgs, render_factor, eval_psnr,
                                       eval_ssim or eval_lpips_alex or eval_lpips_vgg, savedir)
    return rgbs, depths, weights


# ----------------------------------------------------------------------------------------------
# On-disk formats: the reference's checkpoints are torch.save dicts
#   temporalpoints_last.tar : {'global_step', 'model_kwargs', 'model_state_dict', 'optimizer_state_dict'}  (run.py:1234-1235,
#                             lib/utils.py:519-523 load_model: model_class(**ckpt['model_kwargs']); load_state_dict)
#   pcds/canonical.tar, pcds/skeleton.tar : point-cloud / skeleton dicts written by export_point_cloud (run.py:1091-1103)
# ----------------------------------------------------------------------------------------------
def save_checkpoint(path: str, model: TemporalPoints, optimizer=None, global_step: int = 0) -> None:
    """Writes the reference's `temporalpoints_last.tar` layout (run.py:1234-1235)."""
    kw = model.get_kwargs()
    torch.save({'global_step': int(global_step), 'model_kwargs': kw, 'model_state_dict': model.state_dict(),
                'optimizer_state_dict': None if optimizer is None else optimizer.state_dict()}, path)


class _ReferenceObject(torch.nn.Module):
    """Stand-in for a reference class (module `lib.*`) met while unpickling a reference-written checkpoint when the
    reference's code is not importable: pickle restores the instance state (an nn.Module's parameters, buffers, sub-modules
    and plain attributes) without ever calling the class, so the tensors and attribute names survive."""

    def forward(self, *a, **k):
        raise RuntimeError("a reference object loaded without the reference's code cannot be called")


def _reference_pickle_module():
    """A pickle module whose Unpickler maps the reference's classes onto this package's: lib.tineuvox.TiNeuVox ->
    heads.TiNeuVoxHeads (same attribute names for everything TemporalPoints reads: lib/temporalpoints.py:118-147,498-500),
    lib.tineuvox.RGBNet -> heads.RGBNet, anything else under `lib.` -> _ReferenceObject.  Only used when the class cannot be
    imported (a reader that has the reference on its path gets the real classes, as the reference's load_model does)."""
    import importlib
    import pickle
    import types
    from . import heads

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module == "lib" or module.startswith("lib."):
                try:
                    return getattr(importlib.import_module(module), name)
                except Exception:
                    if module == "lib.tineuvox" and name == "TiNeuVox":
                        return heads.TiNeuVoxHeads
                    if module == "lib.tineuvox" and name == "RGBNet":
                        return heads.RGBNet
                    return _ReferenceObject            # one module-level class: the loaded model stays picklable
            return super().find_class(module, name)

    return types.SimpleNamespace(Unpickler=Unpickler, load=lambda f, **kw: Unpickler(f, **kw).load(), __name__="pickle",
                                 loads=pickle.loads, dumps=pickle.dumps, dump=pickle.dump, HIGHEST_PROTOCOL=pickle.HIGHEST_PROTOCOL)


def load_checkpoint(path: str, tineuvox=None, device='cuda', strict: bool = False):
    """lib/utils.py:519-523 `load_model` for the point-cloud model: rebuilds `TemporalPoints(**model_kwargs)` and loads
    the state dict (`strict=False`, as the reference does).  Reads checkpoints written by this package AND
    `temporalpoints_last.tar` files written by the reference (run.py:813-819): their `model_kwargs['tineuvox']` is a pickled
    lib.tineuvox.TiNeuVox, which is mapped onto heads.TiNeuVoxHeads when the reference's code is not importable.
    `tineuvox` (optional) overrides the heads object.  -> (model, checkpoint dict)."""
    ckpt = torch.load(path, map_location='cpu', weights_only=False, pickle_module=_reference_pickle_module())
    kw = dict(ckpt['model_kwargs'])
    if tineuvox is not None:            # default: the heads object pickled inside model_kwargs, as the reference does
        kw['tineuvox'] = tineuvox
    model = TemporalPoints(**kw)
    missing, unexpected = model.load_state_dict(ckpt['model_state_dict'], strict=strict)
    ckpt['missing_keys'], ckpt['unexpected_keys'] = list(missing), list(unexpected)
    model = model.to(device)
    return model, ckpt


# pcds/canonical.tar and pcds/skeleton.tar: what stage 1 hands to stage 2 (written by export_point_cloud, run.py:1090-1103,
# 1229-1235; read by train_pcd, run.py:459-477).  export_point_cloud itself samples the stage-1 voxel model and runs the
# skeletonizer (out of scope); the FORMATS are what lets existing exports flow into the B200 path.
CANONICAL_KEYS = ('pcd', 'rgbs', 'feat', 'raw_feat', 'alphas', 't', 'xyz_min', 'xyz_max', 'voxel_size')
SKELETON_KEYS = ('skeleton_pcd', 'joints', 'root', 'bones', 'pcd', 'weights', 'binary_volume')


def save_pcds(folder: str, *, pcd, rgbs, feat, alphas, skeleton_pcd, joints, bones, xyz_min, xyz_max, voxel_size,
              raw_feat=None, t: float = 0.0) -> None:
    """Writes `canonical.tar` and `skeleton.tar` under `folder` in the reference's layout (run.py:1090-1103, 1214-1230)."""
    import os
    os.makedirs(folder, exist_ok=True)
    torch.save({'pcd': pcd, 'rgbs': rgbs, 'feat': feat, 'raw_feat': raw_feat, 'alphas': alphas, 't': float(t),
                'xyz_min': xyz_min, 'xyz_max': xyz_max, 'voxel_size': voxel_size}, os.path.join(folder, 'canonical.tar'))
    joints_np = np.asarray(torch.as_tensor(joints).cpu())
    torch.save({'skeleton_pcd': np.asarray(torch.as_tensor(skeleton_pcd).cpu()), 'joints': joints_np, 'root': joints_np[0],
                'bones': [list(map(int, b)) for b in bones], 'pcd': None, 'weights': None, 'binary_volume': None},
               os.path.join(folder, 'skeleton.tar'))


def model_from_pcds(read_path: str, tineuvox, *, world_bound_scale: float = 1.05, **model_kwargs) -> TemporalPoints:
    """run.py:457-503 (train_pcd): builds the stage-2 model from `<read_path>/pcds/canonical.tar` + `skeleton.tar`.
    `tineuvox` supplies the frozen stage-1 heads; `model_kwargs` are the reference's `cfg.model_and_render` entries the
    constructor understands (stepsize, fast_color_thres, timebase_pe, pose_embedding_dim, ...)."""
    import os
    can = torch.load(os.path.join(read_path, 'pcds', 'canonical.tar'), map_location='cpu', weights_only=False)
    skel = torch.load(os.path.join(read_path, 'pcds', 'skeleton.tar'), map_location='cpu', weights_only=False)
    missing = [k for k in ('pcd', 'feat', 'alphas', 'rgbs', 'xyz_min', 'xyz_max', 'voxel_size') if k not in can]
    if missing:
        raise KeyError(f"canonical.tar lacks {missing}")
    model_kwargs.pop('world_bound_scale', None)
    return TemporalPoints(
        canonical_pcd=can['pcd'], canonical_feat=can['feat'], canonical_alpha=can['alphas'], canonical_rgbs=can['rgbs'],
        skeleton_pcd=torch.as_tensor(skel['skeleton_pcd']), joints=torch.as_tensor(skel['joints']), bones=skel['bones'],
        xyz_min=torch.as_tensor(can['xyz_min']) * world_bound_scale, xyz_max=torch.as_tensor(can['xyz_max']) * world_bound_scale,
        voxel_size=can['voxel_size'], tineuvox=tineuvox, **model_kwargs)
