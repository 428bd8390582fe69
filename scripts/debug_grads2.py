"""Dissect the end-to-end gradient mismatch: is it conditioning (PE x512 + LeakyReLU kinks) or a kernel bug?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from conftest import model_from_golden, oracle_from_golden, rel_err
g = torch.load(os.path.join(ROOT, "tests/golden/ref_tiny.pt"), weights_only=False)
model, scene = model_from_golden(g)
rk = scene.render_kwargs(); rk.update(rays_o=g["rays_o"].cuda(), rays_d=g["rays_d"].cuda(), viewdirs=g["viewdirs"].cuda())
model.zero_grad(set_to_none=True)
warped = model.warp(g["train"]["t"].cuda())
warped["xyz"].retain_grad(); warped["ginv"].retain_grad()
res = model(g["train"]["t"].cuda(), False, rk, warped=warped)
loss = F.mse_loss(res["rgb_marched"], g["train"]["target"].cuda()) * 200.0
loss.backward()
named = dict(model.named_parameters())
xyz_k = warped["xyz"].detach().cpu(); ginv_k = warped["ginv"].detach().cpu().view(-1, 3, 3)
orc, cfg = oracle_from_golden(g)
with torch.no_grad():
    wp = orc.warp(g["train"]["t"]); Ginv_o = torch.inverse(wp["G"])
def ulps(a, b):
    return (a.view(torch.int32).long() - b.view(torch.int32).long()).abs()
print("xyz kernel vs oracle: max ulp", ulps(xyz_k, wp["xyz"]).max().item(), "rel", rel_err(xyz_k, wp["xyz"]))
print("ginv kernel vs oracle rel", rel_err(ginv_k, Ginv_o[:, :3, :3]))
# oracle evaluated at the kernel's warped cloud
keys = [k for k in g["train"]["grads"] if not k.startswith("forward_warp") and k not in ("weights", "joints", "theta_weight")]
for k in keys: orc.s[k].requires_grad_(True)
x = xyz_k.clone().requires_grad_(True)
Gi = torch.eye(4).repeat(len(x), 1, 1); Gi[:, :3, :3] = ginv_k; Gi.requires_grad_(True)
smp = orc.sample_and_knn(x, g["rays_o"], g["rays_d"], cfg.near, cfg.far, cfg.stepsize, 0.01)
rgb, alpha, *_ = orc.aggregate(x, Gi, smp, g["viewdirs"], cfg.stepsize)
rgb_m, last, depth, _, _, _ = orc.composite(alpha, rgb, smp["ray_id"], smp["step_id"], len(g["rays_o"]), cfg.bg)
l2 = F.mse_loss(rgb_m, g["train"]["target"]) * 200.0
l2.backward()
print("loss kernel", loss.item(), "oracle@kernel-cloud", l2.item(), "golden", g["train"]["loss"].item())
print("d_xyz kernel vs oracle@kernel-cloud", rel_err(warped["xyz"].grad, x.grad))
print("d_ginv kernel vs oracle@kernel-cloud", rel_err(warped["ginv"].grad.view(-1, 3, 3), Gi.grad[:, :3, :3]))
for k in keys:
    print(f"{k:40s} kernel vs oracle@kernel-cloud {rel_err(named[k].grad, orc.s[k].grad):.2e}   kernel vs golden {rel_err(named[k].grad, g['train']['grads'][k]):.2e}"
          f"   oracle@kernel-cloud vs golden {rel_err(orc.s[k].grad, g['train']['grads'][k]):.2e}")
