"""Per-kernel HBM roofline of the bandwidth-bound kernels (LBS fwd/bwd, grid build, compositing fwd/bwd, Adam) at the
scale-sweep sizes of BASELINE.json configs[4] (N = 1M points, J = 65 bones; M = 16M kept samples on 4.2M rays).

    python scripts/kernel_roofline.py [--n 1000000] [--j 65] > profiles/r01_kernel_roofline.json

Each kernel is timed alone with CUDA events on the launching stream (10 launches after 3 warm-ups, L2 flushed by a
256 MiB write before every launch); `achieved` = ALGORITHMIC bytes / mean time, `peak` =
MEASURED_PEAKS.json hbm_gbs (burst figure: the kernel is timed alone).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from articulated_point_nerf_b200 import _lib, ops  # noqa: E402


def timeit(fn, flush, n=10, warm=3):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sum(ms) / len(ms), min(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--j", type=int, default=65)
    ap.add_argument("--rays", type=int, default=2048 * 2048)
    ap.add_argument("--samples-per-ray", type=int, default=4)
    args = ap.parse_args()
    dev = torch.device("cuda")
    lib = _lib.load()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device="cuda").manual_seed(0)
    N, J = args.n, args.j
    out = {"peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs", "N": N, "J": J, "kernels": {}}

    def rec(name, bytes_alg, fn, note=""):
        mean, best = timeit(fn, flush)
        out["kernels"][name] = {"algorithmic_bytes": bytes_alg, "ms_mean": mean, "ms_best": best,
                                "achieved_gbs": bytes_alg / (mean * 1e-3) / 1e9, "frac": bytes_alg / (mean * 1e-3) / 1e9 / peak,
                                "note": note}

    # ---------------------------------------------------------------- LBS
    raw_w = torch.randn(N, J, device=dev, generator=g)
    theta = torch.tensor([0.1], device=dev)
    xyz = torch.rand(N, 3, device=dev, generator=g) * 2 - 1
    ang = torch.randn(J, 3, device=dev, generator=g) * 0.2
    bone_T = torch.eye(4, device=dev).repeat(J, 1, 1)
    bone_T[:, :3, 3] = ang * 0.1
    bone_T[:, 0, 1], bone_T[:, 1, 0] = -ang[:, 2], ang[:, 2]
    gt = torch.zeros(3, device=dev)
    xyz_out = torch.empty(N, 3, device=dev)
    ginv = torch.empty(N, 9, device=dev)
    w_out = torch.empty(N, J, device=dev)
    bbox = torch.empty(6, device=dev)
    P, S = _lib.ptr, _lib.stream

    def lbs_fwd():
        _lib.check(lib.apn_lbs_fwd(P(raw_w), P(theta), 1e-6, None, P(bone_T), P(xyz), P(gt), N, J, P(xyz_out), P(ginv), P(w_out),
                                   None, P(bbox), S()), "lbs_fwd")
    rec("lbs_fwd", N * (4 * J + 12 + 12 + 36 + 4 * J) + 64 * J, lbs_fwd, "raw weights in, xyz in/out, 3x3 inverse out, merged weights out")

    def lbs_fwd_render():          # w_out = NULL: what a no-grad render launches (SURVEY.md §8(d)'s byte count)
        _lib.check(lib.apn_lbs_fwd(P(raw_w), P(theta), 1e-6, None, P(bone_T), P(xyz), P(gt), N, J, P(xyz_out), P(ginv), None,
                                   None, P(bbox), S()), "lbs_fwd")
    rec("lbs_fwd_render", N * (4 * J + 12 + 12 + 36) + 64 * J, lbs_fwd_render,
        "SURVEY 8(d): raw weights in, xyz in/out, 3x3 inverse out (merged weights not stored)")
    d_xyz = torch.randn(N, 3, device=dev, generator=g)
    d_ginv = torch.randn(N, 9, device=dev, generator=g)
    d_raw = torch.empty(N, J, device=dev)
    d_theta = torch.empty(1, device=dev)
    d_bone = torch.empty(J, 4, 4, device=dev)
    d_gt = torch.empty(3, device=dev)
    ws_bytes = lib.apn_lbs_bwd_workspace_bytes(N, J)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def lbs_bwd():
        _lib.check(lib.apn_lbs_bwd(P(raw_w), P(theta), 1e-6, None, P(bone_T), P(xyz), N, J, P(ginv), P(d_xyz), P(d_ginv), None, None,
                                   P(d_raw), P(d_theta), P(d_bone), P(d_gt), P(ws), ws_bytes, S()), "lbs_bwd")
    rec("lbs_bwd", N * (8 * J + 12 + 12 + 36 + 36), lbs_bwd, "re-read raw w, xyz, d x', dG^-1, G^-1; write dw")

    # ---------------------------------------------------------------- grid build
    bb = torch.cat([xyz_out.min(0).values, xyz_out.max(0).values])
    cap = int(min(max(8 * N, 1 << 18), 1 << 25))
    gbytes = lib.apn_grid_workspace_bytes(N, cap)
    blob = torch.empty(gbytes, dtype=torch.uint8, device=dev)

    def grid_build():
        _lib.check(lib.apn_grid_build(P(xyz_out), P(bb), N, 0.01, 0.01, 0.03, cap, P(blob), gbytes, S()), "grid_build")
    rec("grid_build", N * 32, grid_build, "5 launches: count, scan, scatter (xyz read twice, sorted float4 written)")

    # ---------------------------------------------------------------- compositing
    R = args.rays
    spr = args.samples_per_ray
    # image-space coherent sample counts, as a rendered object produces them: a disc covering ~30 % of the frame whose
    # per-ray count follows the chord length through a sphere (max 8 * spr samples), +-1 sample of noise
    side = int(round(R ** 0.5))
    yy, xx = torch.meshgrid(torch.arange(side, device=dev), torch.arange(side, device=dev), indexing="ij")
    rad = (0.3 / 3.14159) ** 0.5 * side
    d2 = ((xx - side / 2) ** 2 + (yy - side / 2) ** 2).float() / (rad * rad)
    chord = torch.sqrt(torch.clamp(1.0 - d2, min=0.0))
    cnt = torch.round(chord * 8 * spr + (chord > 0) * (torch.rand(side, side, device=dev, generator=g) * 2 - 1)).clamp(min=0)
    cnt = cnt.reshape(-1)[:R].to(torch.int32)
    if cnt.numel() < R:
        cnt = torch.cat([cnt, torch.zeros(R - cnt.numel(), dtype=torch.int32, device=dev)])
    ray_start = torch.zeros(R + 1, dtype=torch.int32, device=dev)
    ray_start[1:] = torch.cumsum(cnt, 0)
    M = int(ray_start[-1].item())
    alpha = torch.rand(M, device=dev, generator=g) * 0.5
    rgb = torch.rand(M, 3, device=dev, generator=g)
    step_id = torch.randint(0, 300, (M,), device=dev, generator=g, dtype=torch.int32)
    rgb_m = torch.empty(R, 3, device=dev)
    last = torch.empty(R, device=dev)
    depth = torch.empty(R, device=dev)
    T_save = torch.empty(M, device=dev)
    n_used = torch.empty(R, dtype=torch.int32, device=dev)

    def comp_fwd():
        _lib.check(lib.apn_composite_fwd(P(alpha), P(rgb), P(step_id), None, 0, P(ray_start), R, 1e-4, 1.0, P(rgb_m), P(last), P(depth),
                                         None, P(T_save), P(n_used), S()), "composite_fwd")
    rec("composite_fwd", M * (4 + 12 + 4 + 4) + R * (4 + 12 + 4 + 4 + 4), comp_fwd, f"M={M} samples on R={R} rays")
    d_rgb_m = torch.randn(R, 3, device=dev, generator=g)
    d_last = torch.randn(R, device=dev, generator=g)
    d_alpha = torch.empty(M, device=dev)
    d_rgb = torch.empty(M, 3, device=dev)

    def comp_bwd():
        _lib.check(lib.apn_composite_bwd(P(alpha), P(rgb), None, P(ray_start), R, 1e-4, 1.0, P(T_save), P(n_used), P(last), P(d_rgb_m),
                                         P(d_last), None, P(d_alpha), P(d_rgb), S()), "composite_bwd")
    rec("composite_bwd", M * (4 + 12 + 4 + 4 + 12) + R * (4 + 4 + 4 + 12 + 4), comp_bwd, "alpha, rgb, T in; d_alpha, d_rgb out")

    # ---------------------------------------------------------------- Adam (canonical_feat-sized tensor)
    n_par = N * 128
    p_ = torch.randn(n_par, device=dev, generator=g)
    g_ = torch.randn(n_par, device=dev, generator=g)
    m_ = torch.zeros(n_par, device=dev)
    v_ = torch.zeros(n_par, device=dev)
    plan = ops.AdamPlan([(p_, g_, m_, v_, None, 1e-3, 0)])

    def adam():
        plan.launch([1e-3], 0.9, 0.99, 1e-8)
    rec("adam_multi", n_par * 28, adam, "16 B read + 12 B written per parameter")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
