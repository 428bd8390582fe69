import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from articulated_point_nerf_b200 import ops, render_utils_cuda as ru
from articulated_point_nerf_b200.scene import make_scene
from oracle import dvgo_ops
scene = make_scene("tiny")
ro, rd, vd = [x.reshape(-1, 3).contiguous() for x in scene.rays(0)]
g = torch.Generator().manual_seed(3)
xyz = scene.canonical_pcd + torch.randn(scene.canonical_pcd.shape, generator=g) * 0.002
lo = xyz.min(0)[0] - 0.01
hi = xyz.max(0)[0] + 0.01
sd = scene.cfg.stepsize * scene.voxel_size
ref = dvgo_ops.sample_pts_on_rays(ro, rd, lo, hi, 2.0, 6.0, sd)
got = ru.sample_pts_on_rays(ro.cuda(), rd.cuda(), lo.cuda(), hi.cuda(), 2.0, 6.0, sd)
names = ["pts", "mask", "ray_id", "step_id", "n_steps", "t_min", "t_max"]
for n, a, b in zip(names, got, ref):
    a = a.cpu()
    if a.shape != b.shape:
        print(n, "shape", a.shape, b.shape); continue
    neq = (a != b)
    print(n, "mismatch", int(neq.sum()), "of", a.numel(), "maxabs", float((a.double() - b.double()).abs().max()) if a.numel() else 0)
t_min = ref[5]
s_r, d_r = dvgo_ops.infer_ray_start_dir(ro, rd, t_min)
s_k, d_k = ru.infer_ray_start_dir(ro.cuda(), rd.cuda(), t_min.cuda())
print("start mismatch", int((s_k.cpu() != s_r).sum()), "dir mismatch", int((d_k.cpu() != d_r).sum()))
# which op: norm?
d0, d1, d2 = rd[:, 0], rd[:, 1], rd[:, 2]
rn = torch.sqrt((d0 * d0 + d1 * d1) + d2 * d2)
rn_g = torch.sqrt((d0.cuda() * d0.cuda() + d1.cuda() * d1.cuda()) + d2.cuda() * d2.cuda())
print("torch-gpu vs torch-cpu norm mismatch", int((rn_g.cpu() != rn).sum()))
dk = d_k.cpu()
i = (dk != d_r).nonzero()
print(i[:5])
for r, c in i[:5].tolist():
    print(r, c, rd[r].tolist(), float(rn[r]), dk[r, c].item().hex() if hasattr(dk[r,c].item(),'hex') else dk[r,c].item(), d_r[r, c].item(), (rd[r, c] / rn[r]).item())
