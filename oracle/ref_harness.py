"""TEST INFRASTRUCTURE ONLY — runs the reference's OWN Python (from /root/reference) on CPU.

Only usable in the authoring container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to produce tests/golden/*.pt, which pin oracle/path_oracle.py.

Third-party modules the reference imports but that are absent here are shimmed exactly as
specified in SURVEY.md Appendix A:
  tkinter.W, seaborn.color_palette, matplotlib(.pyplot)           -> inert stand-ins
  roma.rotmat_to_rotvec                                          -> axis * angle in float64 (only its norm is used)
  torch_scatter.segment_coo                                      -> out.index_add_
  pykeops.torch.LazyTensor                                       -> dense blocked brute force,
        d2 = (dx*dx + dy*dy) + dz*dz, K smallest ascending by (d2, index)
  torch.utils.cpp_extension.load                                 -> oracle.dvgo_ops restatements of
        lib/cuda/render_utils_kernel.cu and lib/cuda/adam_upd_kernel.cu (the reference .cu does
        not compile against torch 2.11: 10 x `.type()` -> ScalarType errors)
No reference source is copied; the reference modules are imported from where they lie.
"""
from __future__ import annotations

import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

from . import dvgo_ops
from .path_oracle import hls_palette

REFERENCE_ROOT = "/root/reference"


class _Lazy:
    """Minimal LazyTensor algebra: (a - b) ** 2 -> .sum(-1) -> argKmin / Kmin_argKmin."""

    def __init__(self, t, kind="leaf", a=None, b=None):
        self.t, self.kind, self.a, self.b = t, kind, a, b

    def __sub__(self, o):
        return _Lazy(None, "diff", self, o)

    def __pow__(self, p):
        assert p == 2 and self.kind == "diff"
        return _Lazy(None, "sq", self.a, self.b)

    def sum(self, dim):
        assert dim == -1 and self.kind == "sq"
        return _Lazy(None, "dist", self.a, self.b)

    def _dense_keys(self, rows):
        a, b = self.a.t, self.b.t
        a = a if a.shape[-3] == 1 else a[..., rows, :, :]
        dx = a[..., 0] - b[..., 0]
        dy = a[..., 1] - b[..., 1]
        dz = a[..., 2] - b[..., 2]
        return (dx * dx + dy * dy) + dz * dz

    def _reduce(self, dim, K, with_values):
        assert self.kind == "dist"
        a, b = self.a.t, self.b.t
        nd = a.dim()
        i_axis, j_axis = nd - 3, nd - 2
        n_i, n_j = a.shape[i_axis], b.shape[j_axis]
        reduce_j = (dim == j_axis)
        assert reduce_j or dim == i_axis
        n_out = n_i if reduce_j else n_j
        n_red = n_j if reduce_j else n_i
        batch = a.shape[:i_axis]
        vals = torch.empty(*batch, n_out, K, dtype=torch.float32)
        inds = torch.empty(*batch, n_out, K, dtype=torch.int64)
        blk = max(1, (1 << 28) // max(1, n_red * int(np.prod(batch) if len(batch) else 1)))
        for s in range(0, n_out, blk):
            sl = slice(s, min(n_out, s + blk))
            if reduce_j:
                aa, bb = a[..., sl, :, :], b
            else:
                aa, bb = a, b[..., sl, :]
            dx = aa[..., 0] - bb[..., 0]
            dy = aa[..., 1] - bb[..., 1]
            d2 = dx * dx + dy * dy                      # (..., i, j)
            if aa.shape[-1] == 3:
                dz = aa[..., 2] - bb[..., 2]
                d2 = d2 + dz * dz
            if not reduce_j:
                d2 = d2.transpose(-1, -2)               # (..., j_block, i)
            idx = torch.arange(n_red, dtype=torch.int64)
            key = (d2.contiguous().view(torch.int32).to(torch.int64) << 32) | idx
            kk = torch.topk(key, K, dim=-1, largest=False, sorted=True).values
            inds[..., sl, :] = kk & 0xFFFFFFFF
            vals[..., sl, :] = (kk >> 32).to(torch.int32).view(torch.float32)
        return (vals, inds) if with_values else inds

    def argKmin(self, K, dim):
        return self._reduce(dim, K, False)

    def Kmin_argKmin(self, K, dim):
        return self._reduce(dim, K, True)


def _lazy_tensor(t):
    return _Lazy(t)


def install_shims():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("tkinter", W="w")
    def rotmat_to_rotvec(R):
        """roma.rotmat_to_rotvec restated from its documented semantics (rotation vector = unit axis * angle,
        angle in [0, pi]); evaluated in float64.  Only its norm is consumed (lib/temporalpoints.py:358-361)."""
        Rd = R.double()
        cos = ((Rd[..., 0, 0] + Rd[..., 1, 1] + Rd[..., 2, 2]) - 1.0) / 2.0
        angle = torch.acos(cos.clamp(-1.0, 1.0))
        axis = torch.stack([Rd[..., 2, 1] - Rd[..., 1, 2], Rd[..., 0, 2] - Rd[..., 2, 0], Rd[..., 1, 0] - Rd[..., 0, 1]], -1)
        axis = axis / axis.norm(dim=-1, keepdim=True).clamp_min(1e-300)
        return (axis * angle[..., None]).to(R.dtype)

    mod("roma", rotmat_to_rotvec=rotmat_to_rotvec)
    mod("seaborn", color_palette=lambda name, n: hls_palette(int(n)))
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot")

    def segment_coo(src, index, out=None, reduce="sum"):
        assert reduce == "sum"
        return out.index_add_(0, index, src)

    mod("torch_scatter", segment_coo=segment_coo)
    pk = mod("pykeops")
    pk.torch = mod("pykeops.torch", LazyTensor=_lazy_tensor)

    import torch.utils.cpp_extension as cpp

    def fake_load(name, sources=None, **kw):
        if name == "render_utils_cuda":
            return dvgo_ops.render_utils_namespace
        if name == "adam_upd_cuda":
            return dvgo_ops.adam_namespace
        return SimpleNamespace()

    cpp.load = fake_load


_REF = None


def import_reference():
    """Returns (tineuvox, temporalpoints, pointwarper, masked_adam) modules of the reference."""
    global _REF
    if _REF is None:
        install_shims()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        from lib import tineuvox, temporalpoints, pointwarper, masked_adam  # type: ignore
        _REF = (tineuvox, temporalpoints, pointwarper, masked_adam)
    return _REF


def build_reference_model(scene, seed=0, density_bias=7.0, theta_std=0.2, density_gain=300.0, rgb_gain=8.0,
                          no_view_dir=False, frozen_view_dir=None):
    """Instantiate the reference TiNeuVox (tiny grid; only its heads are used) and TemporalPoints."""
    tineuvox, temporalpoints, _, _ = import_reference()
    torch.manual_seed(seed)
    np.random.seed(seed)
    cfg = scene.cfg
    tv = tineuvox.TiNeuVox(scene.xyz_min.numpy(), scene.xyz_max.numpy(), num_voxels=16 ** 3, num_voxels_base=16 ** 3,
                           alpha_init=1e-3, fast_color_thres=cfg.fast_color_thres, voxel_dim=4, defor_depth=3,
                           net_width=128, no_view_dir=no_view_dir)
    model = temporalpoints.TemporalPoints(
        canonical_pcd=scene.canonical_pcd.clone(), canonical_alpha=scene.canonical_alpha.clone(),
        canonical_feat=scene.canonical_feat.clone(), canonical_rgbs=scene.canonical_rgbs.clone(),
        skeleton_pcd=scene.skeleton_pcd.clone(), joints=scene.joints.clone(), bones=scene.bones,
        xyz_min=scene.xyz_min.numpy(), xyz_max=scene.xyz_max.numpy(), tineuvox=tv,
        stepsize=cfg.stepsize, voxel_size=scene.voxel_size, fast_color_thres=cfg.fast_color_thres,
        pose_embedding_dim=cfg.pose_embedding_dim, frozen_view_dir=frozen_view_dir)
    with torch.no_grad():
        model.densitynet.bias.fill_(density_bias)
        model.densitynet.weight.mul_(density_gain)      # spread alpha over (0, 1)
        model.rgbnet.views_linears[2].weight.mul_(rgb_gain)
        # scale TransformNet's last layer so rotation angles ~ N(0, theta_std)  
        t_embed = tineuvox.poc_fre(torch.tensor([0.37]), model.time_poc)
        out = model.forward_warp.transform_net(t_embed.unsqueeze(0))
        model.forward_warp.transform_net.net[-1].weight.mul_(theta_std / float(out.std()))
    return model, tv
