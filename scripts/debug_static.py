"""debug: where do the dynamic and the static training paths diverge? (gradient checksums per step)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import load_golden, model_from_golden
from articulated_point_nerf_b200.train import GradBucket, GraphedTrainStep, create_optimizer, train_step
g = load_golden("tiny")
gen = torch.Generator().manual_seed(5)
R = len(g["rays_o"])
batches = []
for i in range(4):
    sel = torch.randperm(R, generator=gen)
    batches.append((torch.tensor([0.2 + 0.2 * i]).cuda(), g["rays_o"][sel].cuda(), g["rays_d"][sel].cuda(),
                    g["viewdirs"][sel].cuda(), torch.rand(R, 3, generator=gen).cuda()))
res = {}
for mode in ("dynamic", "dynamic2", "static", "graph"):
    model, scene = model_from_golden(g, fused_pose=True)
    model.decoder_train = "tc"
    opt = create_optimizer(model)
    bucket = GradBucket(opt)
    rk0 = scene.render_kwargs()
    rows = []
    gs = None
    if mode in ("static", "graph"):
        gs = GraphedTrainStep(model, opt, bucket, R, rk0, calibrate=batches[0], use_graph=mode == "graph")
    for t, ro, rd, vd, tgt in batches:
        if gs is None:
            loss = train_step(model, opt, bucket, t, dict(rk0, rays_o=ro, rays_d=rd, viewdirs=vd), tgt)
        else:
            loss = gs.step(t, ro, rd, vd, tgt)
        torch.cuda.synchronize()
        named = dict(model.named_parameters())
        rows.append((float(loss), {k: (float(p.grad.double().sum()), float(p.grad.double().abs().sum())) for k, p in named.items() if p.grad is not None},
                     {k: float(p.detach().double().sum()) for k, p in named.items()}, dict(model.last_counts) if gs is None else None))
    if gs is not None:
        gs.flush(); print(mode, "history", gs.history, "caps", gs.cand_cap, gs.m_cap)
    res[mode] = rows
for i in range(4):
    print("step", i, {m: res[m][i][0] for m in res}, res["dynamic"][i][3])
    for k in res["dynamic"][i][1]:
        a = res["dynamic"][i][1][k]
        line = []
        for m in ("dynamic2", "static", "graph"):
            b = res[m][i][1].get(k)
            line.append("%.2e" % (abs(a[0] - b[0]) / (a[1] + 1e-30)) if b else "none")
        pa = res["dynamic"][i][2][k]
        pl = ["%.2e" % abs(pa - res[m][i][2][k]) for m in ("dynamic2", "static", "graph")]
        if any(float(x) > 1e-4 for x in line if x != "none") or i == 0:
            print(f"   {k:45s} grad-sum diff/abs-sum {line}   param-sum diff {pl}")
