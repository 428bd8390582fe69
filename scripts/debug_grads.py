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
print("loss", loss.item(), g["train"]["loss"].item())
for k, ref in g["train"]["grads"].items():
    print(f"{k:45s} rel_err {rel_err(named[k].grad, ref):.3e}  max|ref| {ref.abs().max().item():.3e}")
# oracle side with intermediate grads
orc, cfg = oracle_from_golden(g)
for v in orc.s.values():
    if v.is_floating_point(): v.requires_grad_(True)
wp = orc.warp(g["train"]["t"])
Ginv = torch.inverse(wp["G"]); Ginv.retain_grad(); wp["xyz"].retain_grad()
smp = orc.sample_and_knn(wp["xyz"], g["rays_o"], g["rays_d"], cfg.near, cfg.far, cfg.stepsize, 0.01)
rgb, alpha, rd_, ad_, _ = orc.aggregate(wp["xyz"], Ginv, smp, g["viewdirs"], cfg.stepsize)
rgb_m, last, depth, _, _, _ = orc.composite(alpha, rgb, smp["ray_id"], smp["step_id"], len(g["rays_o"]), cfg.bg)
l2 = F.mse_loss(rgb_m, g["train"]["target"]) * 200.0
l2.backward()
print("oracle loss", l2.item())
dx_k, dg_k = warped["xyz"].grad.cpu(), warped["ginv"].grad.cpu().view(-1, 3, 3)
print("d_xyz rel", rel_err(dx_k, wp["xyz"].grad), "d_ginv rel", rel_err(dg_k, Ginv.grad[:, :3, :3]))
print("oracle weights grad vs golden", rel_err(orc.s["weights"].grad, g["train"]["grads"]["weights"]))
# feed kernel upstream grads through oracle warp autograd
orc2, _ = oracle_from_golden(g)
for v in orc2.s.values():
    if v.is_floating_point(): v.requires_grad_(True)
wp2 = orc2.warp(g["train"]["t"])
Ginv2 = torch.inverse(wp2["G"])
((wp2["xyz"] * dx_k).sum() + (Ginv2[:, :3, :3] * dg_k).sum()).backward()
print("mixed (kernel upstream -> oracle warp bwd) vs kernel weights grad", rel_err(named["weights"].grad, orc2.s["weights"].grad))
print("mixed vs golden", rel_err(orc2.s["weights"].grad, g["train"]["grads"]["weights"]))
# fp64 truth of the warp backward given the oracle's upstream
orc3, _ = oracle_from_golden(g)
s64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in orc3.s.items()}
from oracle.path_oracle import get_weights, poc_fre, rodrigues4, bone_transforms
import oracle.path_oracle as po
w = torch.softmax(s64["weights"] / torch.max(torch.tensor(1e-6, dtype=torch.float64), s64["theta_weight"]), -1)
bT = wp["bone_Ts"].detach().double()
G = (bT * w[:, :, None, None]).sum(1)
xh = torch.cat([orc3.pcd.double(), torch.ones(len(orc3.pcd), 1, dtype=torch.float64)], -1)
xyz64 = torch.bmm(G, xh.unsqueeze(-1)).squeeze(-1)[:, :3]
Gi64 = torch.inverse(G)
((xyz64 * wp["xyz"].grad.double()).sum() + (Gi64 * Ginv.grad.double()).sum()).backward()
print("fp64 warp-bwd (oracle upstream) vs oracle fp32 weights grad", rel_err(orc.s["weights"].grad, s64["weights"].grad))
print("fp64 warp-bwd (oracle upstream) vs kernel weights grad", rel_err(named["weights"].grad, s64["weights"].grad))
