"""Scratch tool: builds the library with -DTC_TRACE into build/trace/, renders one c1 frame and prints the
clock64 timeline CTA 0 recorded inside agg_tc_fwd_kernel (see TRACE() in csrc/aggregate_tc.cu).

    python scripts/tc_trace.py build        # here (no GPU needed)
    python scripts/tc_trace.py run [tc|tc_fast]   # on the GPU box
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "build", "trace")
LIB = os.path.join(OUT, "libapn_trace.so")


def build():
    from articulated_point_nerf_b200 import build as b
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for src in b.sources():
        obj = os.path.join(OUT, src[:-3] + ".o")
        cmd = [b._nvcc(), *b.NVCC_FLAGS, "-DTC_TRACE", *sys.argv[2:], "-c", os.path.join(b.CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
        objs.append(obj)
    subprocess.run([b._nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    print(LIB)


def run(decoder):
    import torch
    from articulated_point_nerf_b200 import _lib
    _lib.LIB_PATH = LIB
    from articulated_point_nerf_b200.scene import build_model, make_scene
    scene = make_scene("c1")
    model = build_model(scene, seed=0).cuda()
    model.decoder = decoder
    ro, rd, vd = [x.reshape(-1, 3).contiguous().cuda() for x in scene.rays(0)]
    rk = scene.render_kwargs()
    rk.update(rays_o=ro, rays_d=rd, viewdirs=vd)
    t = torch.tensor([0.3], device="cuda")
    lib = _lib.load()
    lib.apn_tc_trace_buffer.restype = C.c_void_p
    lib.apn_tc_trace_buffer.argtypes = [C.c_int]
    with torch.no_grad():
        for _ in range(2):
            model(t, render_depth=True, render_kwargs=rk)
        torch.cuda.synchronize()
        n_ptr = lib.apn_tc_trace_buffer(1)
        C.cdll.LoadLibrary("libcudart.so.12").cudaMemset(C.c_void_p(n_ptr), 0, 16)
        model(t, render_depth=True, render_kwargs=rk)
        torch.cuda.synchronize()
    rt = C.cdll.LoadLibrary("libcudart.so.12")
    n = (C.c_int * 4)()
    rt.cudaMemcpy(n, C.c_void_p(n_ptr), 16, 2)
    buf = (C.c_longlong * (4 * 4096))()
    rt.cudaMemcpy(buf, C.c_void_p(lib.apn_tc_trace_buffer(0)), 4 * 4096 * 8, 2)
    ev = []
    for role in range(4):
        for k in range(n[role]):
            ev.append((buf[role * 4096 + 2 * k + 1], role, buf[role * 4096 + 2 * k]))
    ev.sort()
    t0 = ev[0][0]
    names = {0: "grpA", 1: "grpB", 2: "mmaA", 3: "mmaB"}
    for tt, role, tag in ev[:700]:
        print(f"{tt - t0:9d} {names[role]} {tag}")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2] if len(sys.argv) > 2 else "tc")
