"""TEST INFRASTRUCTURE ONLY — generate tests/golden/ref_tiny_viewdir.pt from the reference itself.

Run in the authoring container:  python -m oracle.make_golden_viewdir
The view-direction variant of the RGB head that no shipped config switches on (lib/temporalpoints.py:504-512):
  frozen   `frozen_view_dir` = one direction for every ray (run.py:480-481 `use_global_view_dir`: the median training
           direction), embedded once into the frozen parameter `viewdirs_emb` (lib/temporalpoints.py:157-159)
Runs the UNMODIFIED reference Python on the `tiny` scene under the shims of oracle/ref_harness.py; stores the state dict,
one render (run.py:149-151) and one training forward + backward (run.py:615-631): loss, rgb_marched and the gradients
downstream of the positional encoding (the view handling only touches the RGB head).

The other variant, `tineuvox.no_view_dir=True` (rgbnet without view columns, lib/tineuvox.py:112-113,80-88), has NO
reference behaviour on this path: lib/temporalpoints.py:504-514 computes `rgbnet(h_feature)` and then falls through to
`rgbnet(h_feature, viewdirs_emb_reshape)` with the name unbound — the reference raises UnboundLocalError (checked here:
`run(variants=("noview",))`).  The product and the oracle implement the evident intent (the first call), tested against
each other only.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from articulated_point_nerf_b200.scene import make_scene  # noqa: E402
from oracle import ref_harness  # noqa: E402

GRAD_KEYS = ("rgbnet.", "densitynet.", "feat_net.2.", "feat_net.3.", "feat_net.4.")


def run(out_path=None, variants=("frozen",)):
    scene = make_scene("tiny")
    rk = scene.render_kwargs()
    rays_o, rays_d, viewdirs = [x.reshape(-1, 3).contiguous() for x in scene.rays(0)]
    rk.update(rays_o=rays_o, rays_d=rays_d, viewdirs=viewdirs)
    t = torch.tensor([0.37])
    gen = torch.Generator().manual_seed(1)
    target = torch.rand(len(rays_o), 3, generator=gen)
    # same scene, seed, rays and time as tests/golden/ref_tiny.pt: every parameter but `viewdirs_emb` equals that file's
    # state dict (asserted below), so this fixture only stores what differs
    base = torch.load(os.path.join(ROOT, "tests", "golden", "ref_tiny.pt"), weights_only=False)
    assert torch.equal(base["rays_o"], rays_o) and torch.equal(base["viewdirs"], viewdirs) and torch.equal(base["render"]["t"], t)
    g = {"config": "tiny", "base": "ref_tiny.pt", "t": t, "target": target}
    for name in variants:
        frozen = viewdirs.median(dim=0)[0] if name == "frozen" else None          # run.py:481
        model, tv = ref_harness.build_reference_model(scene, no_view_dir=(name == "noview"), frozen_view_dir=frozen)
        sd = {k: p.detach().clone() for k, p in model.state_dict().items() if not k.startswith("tineuvox.")}
        assert all(torch.equal(sd[k], p) for k, p in base["state_dict"].items())
        v = {"frozen_view_dir": frozen, "state_dict_extra": {k: p for k, p in sd.items() if k not in base["state_dict"]}}
        with torch.no_grad():
            out = model(t, render_depth=True, render_kwargs=rk)
        v["render"] = {k: out[k].detach().clone() for k in ("rgb_marched", "alphainv_last", "depth", "rgb_marched_direct",
                                                            "t_hat_pcd")}
        model.zero_grad(set_to_none=True)
        res = model(t, False, rk, render_pcd_direct=False)
        loss = torch.nn.functional.mse_loss(res["rgb_marched"], target) * 200.0
        loss.backward()
        v["train"] = {"loss": loss.detach().clone(), "rgb_marched": res["rgb_marched"].detach().clone(),
                      "grads": {k: p.grad.detach().clone() for k, p in model.named_parameters()
                                if p.grad is not None and k.startswith(GRAD_KEYS)}}
        g[name] = v
        print(f"{name}: loss={float(loss):.6f} rgbnet.views_linears.0.weight {tuple(model.rgbnet.views_linears[0].weight.shape)}")
    out_path = out_path or os.path.join(ROOT, "tests", "golden", "ref_tiny_viewdir.pt")
    torch.save(g, out_path)
    print(f"wrote {out_path}: {os.path.getsize(out_path) / 1e6:.2f} MB")
    return g


if __name__ == "__main__":
    run()
