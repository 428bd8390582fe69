"""torch.profiler summary + host-side wall breakdown of one c2 training step (debug aid, not a bench)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_batches, pack_host, unpack_dev
from articulated_point_nerf_b200.scene import make_scene, build_model
from articulated_point_nerf_b200.train import GradBucket, create_optimizer, train_step
from articulated_point_nerf_b200 import ops, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
scene = make_scene(wl)
model = build_model(scene, seed=0).cuda()
model.decoder_train = os.environ.get("DT", "tc")
host = [pack_host(b, True) for b in make_batches(scene, "train", 12, 0)]
dev_in = [(t.cuda(), b.cuda()) for t, b in host]
opt = create_optimizer(model); bucket = GradBucket(opt)
rk = scene.render_kwargs()
def step(i):
    t, ro, rd, vd, tgt = unpack_dev(*dev_in[i])
    return train_step(model, opt, bucket, t, dict(rk, rays_o=ro, rays_d=rd, viewdirs=vd), tgt)
for i in range(5): step(i)
torch.cuda.synchronize()
# wall-clock per step without profiler
t0 = time.perf_counter()
for i in range(5, 10): step(i)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
# host-side breakdown of sample_and_knn with syncs
import types
orig = ops._sample_and_knn
def timed(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(*a, **k)
    torch.cuda.synchronize(); print("  sample_and_knn wall ms", (time.perf_counter() - t0) * 1e3, "cands", r.n_candidates, "M", r.M)
    return r
ops._sample_and_knn = timed
step(10)
ops._sample_and_knn = orig
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=70))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=25, max_name_column_width=60))
