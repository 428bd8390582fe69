"""TEST INFRASTRUCTURE ONLY — generate tests/golden/ref_skeleton.pt from the reference itself.

Run in the authoring container:  python -m oracle.make_golden_skeleton
Pins the skeleton-simplification row against the UNMODIFIED reference:
  * lib/treeprune.py `merge_joints` on the reference's own 29-joint fixture (the arrays of its `__main__` block,
    lib/treeprune.py:301-478, evaluated from the file where it lies) and on seeded random trees / prune masks /
    similarity matrices, with and without `convert_merging_rules`;
  * lib/temporalpoints.py `simplify_skeleton` (:256-343) on the tiny golden scene under the shims of
    oracle/ref_harness.py, both heuristics, followed by the reference's render of the simplified model;
  * lib/temporalpoints.py `get_batch_chamfer_loss` (:765-795) with its gradient, on seeded 2-D / 3-D point sets.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from articulated_point_nerf_b200.scene import make_scene  # noqa: E402
from oracle import ref_harness  # noqa: E402


def reference_fixture():
    """joints / prune / bones / rotation_similarity_matrix arrays of lib/treeprune.py's __main__ block."""
    src = open(os.path.join(ref_harness.REFERENCE_ROOT, "lib", "treeprune.py")).read()
    body = src.split("if __name__ == '__main__':", 1)[1]
    body = body.split("new_joints, new_bones, merging_rules", 1)[0]
    lines = [ln[4:] if ln.startswith("    ") else ln for ln in body.splitlines()]
    lines = [ln for ln in lines if not ln.strip().startswith("import matplotlib")]
    ns = {"np": np}
    exec("\n".join(lines), ns)
    return ns["joints"], ns["bones"], ns["prune"], ns["rotation_similarity_matrix"]


def random_tree(rng, J):
    parents = [int(rng.integers(0, i + 1)) for i in range(J - 1)]
    if rng.random() < 0.5:                                   # chain-heavy trees as well as bushy ones
        parents = [i if rng.random() < 0.7 else p for i, p in enumerate(parents)]
    bones = [[p, i + 1] for i, p in enumerate(parents)]
    joints = rng.normal(size=(J, 3)).astype(np.float32)
    return joints, bones


def run(out_path=None):
    ref_harness.import_reference()
    from lib import treeprune  # type: ignore  (the reference's module)

    def call(joints, bones, prune, sim, convert):
        try:
            out = treeprune.merge_joints(joints, bones, prune.copy(), sim, convert_merging_rules=convert)
            return {"ok": True, "out": [np.asarray(o) for o in out]}
        except Exception as e:  # the reference fails on some degenerate trees (e.g. everything pruned)
            return {"ok": False, "error": type(e).__name__}

    cases = []
    j, b, p, s = reference_fixture()
    for convert in (False, True):
        cases.append(dict(name="reference fixture", joints=j, bones=np.asarray(b).tolist(), prune=p, sim=s,
                          convert=convert, **call(j, b, p, s, convert)))
    rng = np.random.default_rng(0)
    for k in range(120):
        J = int(rng.integers(3, 41))
        joints, bones = random_tree(rng, J)
        prune = rng.random(J) < rng.choice([0.2, 0.5, 0.8])
        prune[0] = False
        sim = rng.random((J, J)) < rng.choice([0.1, 0.4, 0.9])
        sim = sim | sim.T | np.eye(J, dtype=bool)
        convert = bool(k % 2)
        cases.append(dict(name=f"random {k}", joints=joints, bones=bones, prune=prune, sim=sim, convert=convert,
                          **call(joints, bones, prune, sim, convert)))
    n_ok = sum(c["ok"] for c in cases)

    # ---- simplify_skeleton on the tiny scene -----------------------------------------------------------------------
    scene = make_scene("tiny")
    rk = scene.render_kwargs()
    rays_o, rays_d, viewdirs = [x.reshape(-1, 3).contiguous() for x in scene.rays(0)]
    rk.update(rays_o=rays_o, rays_d=rays_d, viewdirs=viewdirs)
    times = torch.linspace(0, 1, 40).unsqueeze(-1)
    simplify = []
    for five, thr in ((True, 8.0), (False, 2.0), (True, 14.0)):
        model, _ = ref_harness.build_reference_model(scene)
        with torch.no_grad():
            joints, bones, new_joints, new_bones, prune_bones, merging_rules, rot_keep, res = model.simplify_skeleton(
                times, deg_threshold=thr, five_percent_heuristic=five)
            t = torch.tensor([0.37])
            out = model(t, render_depth=True, render_kwargs=rk, render_weights=True, poses=scene.poses[0][None],
                        Ks=scene.Ks[0][None], cam_per_ray=torch.zeros(len(rays_o))[:, None], get_skeleton=True)
        simplify.append(dict(
            five_percent=five, deg_threshold=thr, times=times.clone(), t=t,
            new_joints=np.asarray(new_joints), new_bones=np.asarray(new_bones), prune_bones=prune_bones.clone(),
            merging_rules=np.asarray(merging_rules), rotations_to_keep=rot_keep.clone(),
            flat_merging_rules=model.flat_merging_rules.clone().long(),
            sibling_merging_rules=model.sibling_merging_rules.clone().long(),
            rot_mask=model.forward_warp.rot_mask.clone(), sibling_mask=model.forward_warp.sibling_mask.clone(),
            last_weights=model._last_weights.detach().clone(),
            out={k: v.detach().clone() for k, v in out.items() if torch.is_tensor(v)}))
        print(f"simplify five={five} thr={thr}: frozen {int(prune_bones.sum())}/{len(prune_bones)}, "
              f"merged columns {int((model.flat_merging_rules != torch.arange(len(prune_bones))).sum())}, "
              f"sibling transfers {int((model.forward_warp.sibling_mask != torch.arange(len(prune_bones))).sum())}")

    # ---- batch chamfer loss (lib/temporalpoints.py:765-795; run.py:659-690: projected cloud vs integer mask pixels) ----
    gen = torch.Generator().manual_seed(3)
    chamfer = []
    for B, N, M, D in ((3, 500, 400, 2), (2, 300, 1300, 3), (1, 7, 5, 2)):
        p1 = (torch.rand(B, N, D, generator=gen) * 40).requires_grad_(True)
        p2 = torch.randint(0, 40, (B, M, D), generator=gen).float()          # pixel grid: many exact ties
        loss = model.get_batch_chamfer_loss(p1, p2)
        loss.backward()
        chamfer.append(dict(pcd1=p1.detach().clone(), pcd2=p2.clone(), loss=loss.detach().clone(), grad1=p1.grad.clone()))

    out_path = out_path or os.path.join(ROOT, "tests", "golden", "ref_skeleton.pt")
    torch.save({"merge_joints": cases, "simplify": simplify, "batch_chamfer": chamfer}, out_path)
    print(f"wrote {out_path}: {len(cases)} merge_joints cases ({n_ok} the reference completes), "
          f"{len(simplify)} simplify_skeleton runs, {os.path.getsize(out_path) / 1e6:.2f} MB")


if __name__ == "__main__":
    run()
