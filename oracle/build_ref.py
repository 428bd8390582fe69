"""TEST INFRASTRUCTURE ONLY — compiles the reference's own CUDA extensions into oracle/_ref/ (git-ignored).

    python -m oracle.build_ref [--force]

`render_utils_cuda` (lib/cuda/render_utils.cpp + render_utils_kernel.cu, pybind surface lib/cuda/render_utils.cpp:144-155)
and `adam_upd_cuda` (lib/cuda/adam_upd.cpp + adam_upd_kernel.cu, lib/cuda/adam_upd.cpp:79-86) are built for sm_100a from the
sources WHERE THEY LIE under /root/reference — nothing is copied or patched; the one incompatibility with torch 2.x (the
deprecated `tensor.type()` in AT_DISPATCH) is bridged by the force-included oracle/ref_shim.h.  Only the authoring
container has /root/reference; the built .so files travel to the GPU box inside oracle/_ref/ and are loaded there by
`load()`.  tests/test_gpu_reference_kernels.py uses them as the secondary oracle for the DVGO ops.
"""
from __future__ import annotations

import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CUDA = "/root/reference/lib/cuda"
OUT = os.path.join(ROOT, "oracle", "_ref")
SHIM = os.path.join(ROOT, "oracle", "ref_shim.h")
MODULES = {
    "render_utils_cuda": ["render_utils.cpp", "render_utils_kernel.cu"],
    "adam_upd_cuda": ["adam_upd.cpp", "adam_upd_kernel.cu"],
}


def so_path(name: str) -> str:
    return os.path.join(OUT, name, name + ".so")


def available() -> bool:
    return all(os.path.exists(so_path(n)) for n in MODULES)


def build(force: bool = False, verbose: bool = False):
    """Builds both modules when /root/reference is present; returns the list of .so paths (or None without the reference)."""
    if not os.path.isdir(REF_CUDA):
        return [so_path(n) for n in MODULES] if available() else None
    import torch  # noqa: F401
    from torch.utils import cpp_extension
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(min(8, os.cpu_count() or 1)))
    outs = []
    for name, files in MODULES.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        if force or not os.path.exists(so_path(name)):
            cpp_extension.load(
                name=name, sources=[os.path.join(REF_CUDA, f) for f in files], build_directory=bdir, verbose=verbose,
                extra_cflags=["-O2", "-include", SHIM],
                extra_cuda_cflags=["-O2", "-include", SHIM, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"],
                is_python_module=True)
        outs.append(so_path(name))
    return outs


def load(name: str):
    """Imports a prebuilt reference extension from oracle/_ref (no compilation; needs a CUDA device to be useful)."""
    import torch  # noqa: F401  (the pybind module links against libtorch)
    path = so_path(name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `python -m oracle.build_ref` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    r = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("reference CUDA extensions:", r if r else "unavailable (no /root/reference and nothing prebuilt)")
