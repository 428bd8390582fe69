import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import rel_err
import test_gpu_kernels as T
g = torch.load(os.path.join(ROOT, "tests/golden/ref_tiny.pt"), weights_only=False)
ops, orc, o, xyz, gi, smp, model, c = T._agg_setup(g, True)
rgb, alpha, _, _, _ = o
gen = torch.Generator().manual_seed(2)
ca, cr = torch.randn(alpha.shape, generator=gen), torch.randn(rgb.shape, generator=gen)
((alpha * ca).sum() + (rgb * cr).sum()).backward()
keys = ["canonical_feat", "feat_net.0.weight", "feat_net.0.bias", "feat_net.2.0.weight", "feat_net.2.0.bias", "feat_net.3.0.weight",
        "feat_net.3.0.bias", "feat_net.4.weight", "feat_net.4.bias", "densitynet.weight", "densitynet.bias",
        "rgbnet.feature_linears.weight", "rgbnet.views_linears.2.bias"]
def run(fn):
    kx = xyz.detach().cuda().requires_grad_(True)
    kg = gi.detach()[:, :3, :3].reshape(-1, 9).contiguous().cuda().requires_grad_(True)
    model.zero_grad()
    k_alpha, k_rgb, *_ = fn(kx, kg)
    ((k_alpha * ca.cuda()).sum() + (k_rgb * cr.cuda()).sum()).backward()
    named = dict(model.named_parameters())
    out = {"d_xyz": kx.grad.clone(), "d_ginv": kg.grad.clone()}
    for k in keys: out[k] = named[k].grad.clone()
    return out, k_alpha, k_rgb
tc, a1, r1 = run(lambda kx, kg: ops.aggregate_tc_train(c, kx, kg, model.canonical_feat, model._mlp_weights(), ops.PackedDecoder()))
f32, a2, r2 = run(lambda kx, kg: ops.aggregate(c, kx, kg, model.canonical_feat, None, model._mlp_weights()))
print("fwd alpha tc vs oracle", rel_err(a1, alpha), "rgb", rel_err(r1, rgb))
ref = {"d_xyz": xyz.grad, "d_ginv": gi.grad[:, :3, :3].reshape(-1, 9)}
for k in keys: ref[k] = orc.s[k].grad
for k in tc:
    print(f"{k:34s} tc vs oracle {rel_err(tc[k].reshape(ref[k].shape), ref[k]):.2e}   fp32 vs oracle {rel_err(f32[k].reshape(ref[k].shape), ref[k]):.2e}   tc vs fp32 {rel_err(tc[k], f32[k]):.2e}")
# where is the d_xyz error concentrated
d = (tc["d_xyz"] - f32["d_xyz"]).abs().cpu()
print("d_xyz abs err: max", d.max().item(), "n > 1e-4*max", int((d > 1e-4 * f32["d_xyz"].abs().max().cpu()).sum()), "of", d.numel())
# ---- kink-flip hypothesis: compare LeakyReLU masks of the TC tape with the oracle's pre-activation signs
import ctypes as C
from articulated_point_nerf_b200 import _lib
from articulated_point_nerf_b200._lib import ptr, stream, check, AggOutputs
lib = _lib.load()
M = c.pts.shape[0]; dev = "cuda"
ws = [w.detach().float().contiguous() for w in model._mlp_weights()]
pk, table = ops.PackedDecoder().get(ws, 191, model.canonical_feat.detach())
tape_bytes = lib.apn_aggregate_tc_tape_bytes(M)
tape = ops._aligned_bytes(tape_bytes, dev)
outs = [torch.empty(s, device=dev) for s in [(M,), (M, 3), (M,), (M, 3), (M, 8), (M, 128), (M,), (M, 160), (M, 64)]]
out = AggOutputs(); out.alpha, out.rgb, out.alpha_direct, out.rgb_direct, out.idw, out.h, out.exp_d, out.fv, out.v0 = [ptr(t) for t in outs]
a = ops._agg_inputs(c, xyz.detach().cuda(), gi.detach()[:, :3, :3].reshape(-1, 9).contiguous().cuda(), model.canonical_feat.detach(), None, M, 191)
w = ops._mlp_struct(ws)
check(lib.apn_aggregate_fwd_tc(C.byref(a), C.byref(w), ptr(pk), ptr(table), C.byref(out), 1, ptr(tape), tape_bytes, None, 0, stream()))
torch.cuda.synchronize()
tp = tape.cpu().numpy()
import numpy as np
TILE = 18 * 16384 + 3 * 2048
n_tiles = (M + 15) // 16
# oracle pre-activations per layer
with torch.no_grad():
    s_i, pts = smp["s_i"], smp["pts"]
    rel_p = pts[:, None, :] - xyz.detach()[s_i, :]
    frames = gi.detach()[s_i]
    rel_c = torch.bmm(frames[..., :3, :3].reshape(-1, 3, 3), rel_p.reshape(-1, 3).unsqueeze(-1)).squeeze(-1)
    from oracle.path_oracle import poc_fre
    x = torch.cat([poc_fre(rel_c, orc.pos_poc), orc.s["canonical_feat"][s_i, :].reshape(-1, 128)], -1)
    pres = []
    for nme in ["feat_net.0", "feat_net.2.0", "feat_net.3.0"]:
        pre = torch.nn.functional.linear(x, orc.s[nme + ".weight"], orc.s[nme + ".bias"]); pres.append(pre)
        x = torch.nn.functional.leaky_relu(pre, 0.01)
for l in range(3):
    bits = np.zeros((n_tiles * 128, 128), dtype=bool)
    for t in range(n_tiles):
        mk = tp[t * TILE + 18 * 16384 + l * 2048: t * TILE + 18 * 16384 + (l + 1) * 2048].view(np.uint16).reshape(128, 8)
        for grp in range(8):
            ph, cq = grp // 4, grp % 4
            for i in range(16):
                bits[t * 128:(t + 1) * 128, ph * 64 + cq * 16 + i] = (mk[:, grp] >> i) & 1
    ref_bits = (pres[l] > 0).numpy()
    nrows = ref_bits.shape[0]
    diff = bits[:nrows] != ref_bits
    idxs = np.argwhere(diff)
    print(f"layer {l}: mask flips {diff.sum()} of {diff.size}; |pre| at flips:", [float(pres[l][r, cc].abs()) for r, cc in idxs[:8]])
