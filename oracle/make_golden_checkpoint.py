"""TEST INFRASTRUCTURE ONLY — tests/golden/ref_mini_last.tar (+ ref_mini_pcds/): files WRITTEN BY THE REFERENCE.

    python -m oracle.make_golden_checkpoint

The unmodified reference (lib/temporalpoints.py under the shims of oracle/ref_harness.py) builds a 300-point model and saves
  * `temporalpoints_last.tar` exactly as run.py:813-819 does: {'global_step', 'model_kwargs': model.get_kwargs(),
    'model_state_dict'} — model_kwargs pickles the reference's own TiNeuVox object (class lib.tineuvox.TiNeuVox), which is
    what a loader without /root/reference on its path has to cope with;
  * `pcds/canonical.tar` / `pcds/skeleton.tar` in the layout of run.py:1091-1103, 1229-1235 (written here with the
    reference's key set; export_point_cloud itself needs the stage-1 voxel model and open3d);
plus the rendering of one small ray batch by the reference model, so that the loaded model can be checked end to end.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from articulated_point_nerf_b200.scene import SceneConfig, make_scene  # noqa: E402
from oracle import ref_harness  # noqa: E402


def run():
    cfg = SceneConfig(name="mini", n_points=300, H=24, W=24, n_views=2)
    scene = make_scene(cfg)
    model, tv = ref_harness.build_reference_model(scene)
    out_dir = os.path.join(ROOT, "tests", "golden")
    ckpt = os.path.join(out_dir, "ref_mini_last.tar")
    torch.save({'global_step': 1234, 'model_kwargs': model.get_kwargs(), 'model_state_dict': model.state_dict()}, ckpt)   # run.py:813-819
    pcds = os.path.join(out_dir, "ref_mini_pcds", "pcds")
    os.makedirs(pcds, exist_ok=True)
    torch.save({'pcd': scene.canonical_pcd, 'rgbs': scene.canonical_rgbs, 'feat': scene.canonical_feat, 'raw_feat': None,
                'alphas': scene.canonical_alpha, 't': 0.0, 'xyz_min': scene.canonical_pcd.min(0)[0],
                'xyz_max': scene.canonical_pcd.max(0)[0], 'voxel_size': scene.voxel_size}, os.path.join(pcds, 'canonical.tar'))  # run.py:1091-1103
    torch.save({'skeleton_pcd': scene.skeleton_pcd.numpy(), 'joints': scene.joints.numpy(), 'root': scene.joints[0].numpy(),
                'bones': scene.bones, 'pcd': None, 'weights': None, 'binary_volume': None}, os.path.join(pcds, 'skeleton.tar'))  # :1229-1235
    rk = scene.render_kwargs()
    ro, rd, vd = [x.reshape(-1, 3).contiguous() for x in scene.rays(0)]
    rk.update(rays_o=ro, rays_d=rd, viewdirs=vd)
    t = torch.tensor([0.6])
    with torch.no_grad():
        out = model(t, render_depth=True, render_kwargs=rk, poses=scene.poses[0][None], Ks=scene.Ks[0][None])
    torch.save({"t": t, "rays_o": ro, "rays_d": rd, "viewdirs": vd, "render_kwargs": {k: v for k, v in rk.items() if not torch.is_tensor(v)},
                "rgb_marched": out["rgb_marched"], "depth": out["depth"], "t_hat_pcd": out["t_hat_pcd"],
                "state_keys": sorted(model.state_dict().keys())}, os.path.join(out_dir, "ref_mini_render.pt"))
    print(f"wrote {ckpt} ({os.path.getsize(ckpt) / 1e6:.2f} MB), pcds/, ref_mini_render.pt; rays hit: "
          f"{int((out['alphainv_last'] < 0.999).sum()) if out.get('alphainv_last') is not None else 0} of {len(ro)}")


if __name__ == "__main__":
    run()
