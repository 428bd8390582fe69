"""fp64 ground truth for the ragged TC-backward test: which path is off on feat_net.0 grads?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from articulated_point_nerf_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
torch.manual_seed(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
g = torch.Generator().manual_seed(100 + M)
N, d = 2000, "cuda"
xyz = torch.rand(N, 3, generator=g)
A = torch.eye(3) + 0.2 * torch.randn(N, 3, 3, generator=g)
feat = torch.relu(torch.randn(N, 128, generator=g)) * 0.5
nn_idx = torch.randint(0, N, (M, 8), generator=g).int()
pts = xyz[nn_idx[:, 0].long()] + 0.02 * torch.randn(M, 3, generator=g)
ray_id = torch.sort(torch.randint(0, 50, (M,), generator=g))[0].int()
vd = torch.nn.functional.normalize(torch.randn(50, 3, generator=g), dim=-1)
lin = [torch.nn.Linear(191, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 128),
       torch.nn.Linear(128, 1), torch.nn.Linear(128, 128), torch.nn.Linear(155, 64), torch.nn.Linear(64, 3)]
c = ops.AggConst(pts=pts.to(d), nn_idx=nn_idx.to(d), ray_id=ray_id.to(d), viewdirs=vd.to(d),
                 canonical_alpha=torch.rand(N, generator=g).to(d), canonical_rgbs=torch.rand(N, 3, generator=g).to(d),
                 direct_eps=torch.full((N,), 0.05).to(d), mean_min_distance=0.02, eps=1e-6, act_shift=0.0, interval=0.5)
ca, cr = torch.randn(M, generator=g).to(d), torch.randn(M, 3, generator=g).to(d)
def run(tc):
    leaves = [xyz.to(d).requires_grad_(True), A.reshape(N, 9).contiguous().to(d).requires_grad_(True), feat.to(d).requires_grad_(True)]
    ws = []
    for l in lin: ws += [l.weight.detach().to(d).requires_grad_(True), l.bias.detach().to(d).requires_grad_(True)]
    out = ops.aggregate_tc_train(c, *leaves, ws, ops.PackedDecoder()) if tc else ops.aggregate(c, *leaves, None, ws)
    ((out[0] * ca).sum() + (out[1] * cr).sum()).backward()
    return [t.grad for t in leaves + ws], out
# fp64 truth
def truth():
    D = torch.float64
    x = xyz.to(d, D).requires_grad_(True); Ai = A.to(d, D).requires_grad_(True); f = feat.to(d, D).requires_grad_(True)
    ws = []
    for l in lin: ws += [l.weight.detach().to(d, D).requires_grad_(True), l.bias.detach().to(d, D).requires_grad_(True)]
    si = nn_idx.long().to(d)
    rel_p = pts.to(d, D)[:, None, :] - x[si]
    to_nn = (rel_p ** 2).sum(-1)
    w = 1 / (to_nn + 1e-6); w = (w / w.sum(-1, keepdim=True)).unsqueeze(-1)
    rel_c = torch.einsum("mkab,mkb->mka", Ai[si], rel_p).reshape(-1, 3)
    poc = torch.tensor([2.0 ** i for i in range(10)], device=d, dtype=D)
    emb = (rel_c.unsqueeze(-1) * poc).flatten(-2)
    pe = torch.cat([rel_c, emb.sin(), emb.cos()], -1)
    h = torch.cat([pe, f[si].reshape(-1, 128)], -1)
    pres = []
    for i in range(4):
        pre = torch.nn.functional.linear(h, ws[2 * i], ws[2 * i + 1]); pre.retain_grad(); pres.append(pre)
        h = torch.nn.functional.leaky_relu(pre, 0.01)
    hh = (h.reshape(M, 8, 128) * w).sum(1); hh.retain_grad()
    truth.pres, truth.hh, truth.acts_max = pres, hh, [p_.abs().max().item() for p_ in pres]
    dens = torch.nn.functional.linear(hh, ws[8], ws[9]).squeeze(-1)
    alpha = 1 - (1 + torch.exp(dens + 0.0)) ** (-0.5)
    vpoc = torch.tensor([2.0 ** i for i in range(4)], device=d, dtype=D)
    v = vd.to(d, D); ve = (v.unsqueeze(-1) * vpoc).flatten(-2); vemb = torch.cat([v, ve.sin(), ve.cos()], -1)[ray_id.long().to(d)]
    ff = torch.nn.functional.linear(hh, ws[10], ws[11])
    r = torch.sigmoid(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(torch.cat([ff, vemb], -1), ws[12], ws[13])), ws[14], ws[15]))
    ((alpha * ca.double()).sum() + (r * cr.double()).sum()).backward()
    return [x.grad, Ai.grad.reshape(N, 9), f.grad] + [t.grad for t in ws]
tc, otc = run(True); f32, of32 = run(False); tr = truth()
print("forward tc vs fp32: alpha", (otc[0] - of32[0]).abs().max().item(), "rgb", (otc[1] - of32[1]).abs().max().item())
q = torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], device=d, dtype=torch.float64)
print("|d_h| quantiles", torch.quantile(truth.hh.grad.abs().flatten()[::7], q).tolist())
for i, p_ in enumerate(truth.pres):
    print(f"|dpre{i}| quantiles", torch.quantile(p_.grad.abs().flatten()[::37], q).tolist(), "max |pre|", truth.acts_max[i])
names = ["xyz", "ginv", "feat"] + [f"w{i}" for i in range(16)]
for n_, a, b, t in list(zip(names, tc, f32, tr))[:5]:
    sc = t.abs().max().item() + 1e-30
    print(f"{n_:5s} tc-truth {((a.double() - t).abs().max() / sc).item():.2e}   fp32-truth {((b.double() - t).abs().max() / sc).item():.2e}   frac>5e-4 tc {(((a.double()-t).abs()/sc) > 5e-4).double().mean().item():.3f}")
