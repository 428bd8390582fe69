import sys, os, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_batches, pack_host, unpack_dev
from articulated_point_nerf_b200.scene import make_scene, build_model
from articulated_point_nerf_b200.train import GradBucket, create_optimizer, train_step
scene = make_scene("c2")
model = build_model(scene, seed=0).cuda(); model.decoder_train = os.environ.get("DT", "tc")
host = [pack_host(b, True) for b in make_batches(scene, "train", 20, 0)]
dev_in = [(t.cuda(), b.cuda()) for t, b in host]
opt = create_optimizer(model); bucket = GradBucket(opt)
rk = scene.render_kwargs()
def step(i):
    t, ro, rd, vd, tgt = unpack_dev(*dev_in[i % 20])
    return train_step(model, opt, bucket, t, dict(rk, rays_o=ro, rays_d=rd, viewdirs=vd), tgt)
for i in range(20): step(i)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(100): step(i)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) * 10)
pr = cProfile.Profile(); pr.enable()
for i in range(100): step(i)
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
