"""Which allocations still reach cudaMalloc in steady-state training steps? (debug aid)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_batches, pack_host, unpack_dev
from articulated_point_nerf_b200.scene import make_scene, build_model
from articulated_point_nerf_b200.train import GradBucket, create_optimizer, train_step
scene = make_scene("c2")
model = build_model(scene, seed=0).cuda()
host = [pack_host(b, True) for b in make_batches(scene, "train", 40, 0)]
dev_in = [(t.cuda(), b.cuda()) for t, b in host]
opt = create_optimizer(model); bucket = GradBucket(opt)
rk = scene.render_kwargs()
def step(i):
    t, ro, rd, vd, tgt = unpack_dev(*dev_in[i])
    return train_step(model, opt, bucket, t, dict(rk, rays_o=ro, rays_d=rd, viewdirs=vd), tgt)
for i in range(10): step(i)
torch.cuda.synchronize()
torch.cuda.memory._record_memory_history(max_entries=200000, stacks="python")
n0 = torch.cuda.memory_stats()["num_device_alloc"]
for i in range(10, 40):
    step(i)
torch.cuda.synchronize()
print("device allocs in 30 steady steps:", torch.cuda.memory_stats()["num_device_alloc"] - n0)
snap = torch.cuda.memory._snapshot()
for tr in snap["device_traces"]:
    for ev in tr:
        if ev["action"] in ("segment_alloc", "segment_free"):
            fr = [f"{os.path.basename(f['filename'])}:{f['line']}:{f['name']}" for f in ev.get("frames", []) if "articulated" in f["filename"] or "bench" in f["filename"] or "train" in f["filename"]][:3]
            print(ev["action"], ev["size"] / 1e6, "MB", fr)
