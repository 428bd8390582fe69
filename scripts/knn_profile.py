"""Where does the k-NN time go on a c1 frame?  (debug aid)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from articulated_point_nerf_b200.scene import make_scene, build_model
from articulated_point_nerf_b200 import ops, _lib
from articulated_point_nerf_b200._lib import ptr, stream, check
wl = sys.argv[1] if len(sys.argv) > 1 else "c1"
scene = make_scene(wl)
model = build_model(scene, seed=0).cuda()
ro, rd, vd = [x.reshape(-1, 3).contiguous().cuda() for x in scene.rays(0)]
with torch.no_grad():
    warped = model.warp(torch.tensor([0.3]).cuda())
    grid = model.build_grid(warped)
    print(grid.describe())
    stepdist = scene.cfg.stepsize * scene.voxel_size
    smp, dbg = ops.sample_and_knn(grid, ro, rd, scene.cfg.near, scene.cfg.far, stepdist, return_d2=True)
keep = dbg["keep"].bool()
print("candidates", len(keep), "kept", int(keep.sum()))
if len(sys.argv) > 2 and sys.argv[2] == "time":       # end-to-end sample_and_knn with the search the environment selects
    if os.environ.get("APN_KNN_LEGACY"):
        ops.KNN_SORTED_MIN = 1 << 60
    for _ in range(3):
        ops.sample_and_knn(grid, ro, rd, scene.cfg.near, scene.cfg.far, stepdist)
    torch.cuda.synchronize()
    ts = []
    for _ in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s2 = ops.sample_and_knn(grid, ro, rd, scene.cfg.near, scene.cfg.far, stepdist)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"sample_and_knn {wl} FORCE={os.environ.get('APN_KNN_FORCE')} LEGACY={os.environ.get('APN_KNN_LEGACY')}: median {ts[len(ts) // 2]:.3f} ms, "
          f"min {ts[0]:.3f} ms, max {ts[-1]:.3f} ms per frame, M={s2.M}")
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "stats":       # needs a build with APN_EXTRA_NVCC_FLAGS=-DKS_STATS
    import ctypes
    raw = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_ulonglong * 8)()
    torch.cuda.synchronize(); raw.apn_knn_sorted_stats(buf)
    ops.sample_and_knn(grid, ro, rd, scene.cfg.near, scene.cfg.far, stepdist)
    torch.cuda.synchronize(); raw.apn_knn_sorted_stats(buf)
    v = list(buf)
    q = max(v[7], 1)
    print(f"stats {wl}: queries {v[7]}, groups {v[0]} ({v[7] / max(v[0], 1):.1f} q/group), rounds/group {v[1] / max(v[0], 1):.2f}, per QUERY: "
          f"leaf slots {v[2] / q:.1f}, leaves read {v[3] / q:.1f}, points loaded {v[4] / q:.1f}, staged {v[5] / q:.1f}, distance evals {v[6] / q:.1f}")
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "short":      # under ncu: the one sample_and_knn call above is all that is profiled
    torch.cuda.synchronize()
    sys.exit(0)
d8 = dbg["d2"][keep][:, 7].sqrt()
qs = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9, 0.99]).cuda()
print("d8 quantiles of kept", torch.quantile(d8[:: max(1, len(d8) // 200000)], qs).tolist(), "cell", grid.describe()["cell"])
lib = _lib.load()
def time_knn(cr, cs, name):
    n = len(cr)
    nn = torch.empty(n, 8, dtype=torch.int32, device="cuda"); kp = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        check(lib.apn_knn(ptr(ro), ptr(rd), scene.cfg.near, scene.cfg.far, stepdist, ptr(grid.blob), ptr(cr), ptr(cs), n, ptr(nn), None, ptr(kp), stream()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        check(lib.apn_knn(ptr(ro), ptr(rd), scene.cfg.near, scene.cfg.far, stepdist, ptr(grid.blob), ptr(cr), ptr(cs), n, ptr(nn), None, ptr(kp), stream()))
    b.record(); torch.cuda.synchronize()
    print(f"{name:28s} n={n:8d}  {a.elapsed_time(b) / 5 * 1e3:9.1f} us  {a.elapsed_time(b) / 5 * 1e6 / max(n,1):7.1f} ns/query")
cr, cs = dbg["cand_ray"], dbg["cand_step"]
time_knn(cr, cs, "all candidates")
time_knn(cr[keep].contiguous(), cs[keep].contiguous(), "kept only")
time_knn(cr[~keep].contiguous(), cs[~keep].contiguous(), "rejected only")
d8all = torch.full((len(keep),), 1.0, device="cuda"); d8all[keep] = d8
for lo, hi in [(0, 0.0126), (0.0126, 0.025), (0.025, 0.05), (0.05, 0.1001)]:
    m = keep & (d8all >= lo) & (d8all < hi)
    time_knn(cr[m].contiguous(), cs[m].contiguous(), f"kept, d8 in [{lo},{hi})")
