"""ctypes binding of libapn_sm100.so (include/apn.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module without
the built library raises; calling into it without a CUDA device raises from the CUDA runtime.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libapn_sm100.so")

c_float_p = C.POINTER(C.c_float)
c_void_p = C.c_void_p


class ApnError(RuntimeError):
    pass


class MlpWeights(C.Structure):
    _fields_ = [("w", c_void_p * 4), ("b", c_void_p * 4), ("density_w", c_void_p), ("density_b", c_void_p),
                ("rgb_feat_w", c_void_p), ("rgb_feat_b", c_void_p), ("rgb_v0_w", c_void_p), ("rgb_v0_b", c_void_p),
                ("rgb_v2_w", c_void_p), ("rgb_v2_b", c_void_p)]


class AggInputs(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("d_in", C.c_int), ("pts", c_void_p), ("nn_idx", c_void_p),
                ("ray_id", c_void_p), ("xyz", c_void_p), ("ginv", c_void_p), ("feat", c_void_p), ("pose_emb", c_void_p),
                ("viewdirs", c_void_p), ("canonical_alpha", c_void_p), ("canonical_rgbs", c_void_p),
                ("direct_eps", c_void_p), ("mean_min_distance", C.c_float), ("eps", C.c_float),
                ("act_shift", C.c_float), ("interval", C.c_float), ("m_dev", c_void_p)]


class AggOutputs(C.Structure):
    _fields_ = [("alpha", c_void_p), ("rgb", c_void_p), ("alpha_direct", c_void_p), ("rgb_direct", c_void_p),
                ("idw", c_void_p), ("x0", c_void_p), ("act", c_void_p * 4), ("h", c_void_p), ("exp_d", c_void_p),
                ("fv", c_void_p), ("v0", c_void_p)]


class AggGrads(C.Structure):
    _fields_ = [("d_alpha", c_void_p), ("d_rgb", c_void_p), ("d_xyz", c_void_p), ("d_ginv", c_void_p),
                ("d_feat", c_void_p), ("d_pose_emb", c_void_p), ("d_w", c_void_p * 4), ("d_b", c_void_p * 4),
                ("d_density_w", c_void_p), ("d_density_b", c_void_p), ("d_rgb_feat_w", c_void_p),
                ("d_rgb_feat_b", c_void_p), ("d_rgb_v0_w", c_void_p), ("d_rgb_v0_b", c_void_p),
                ("d_rgb_v2_w", c_void_p), ("d_rgb_v2_b", c_void_p)]


class AdamTensor(C.Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("perlr", c_void_p), ("numel", C.c_longlong), ("step_size", C.c_float), ("mode", C.c_int)]


I, F, P, LL, SZ = C.c_int, C.c_float, c_void_p, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); mirrors include/apn.h one to one
SIGNATURES = {
    "apn_version": (I, []),
    "apn_last_error": (C.c_char_p, []),
    "apn_launch_count": (C.c_ulonglong, []),
    "apn_range_push": (I, [C.c_char_p]),
    "apn_range_pop": (I, []),
    "apn_pose_saved_bytes": (SZ, [I]),
    "apn_pose_fwd": (I, [P, I, P, P, P, P, P, P, P, I, P, P, P, P, P]),
    "apn_pose_bwd": (I, [P, I, P, P, P, P, P, P, P, I, P, P, P, P, P, P, P, P]),
    "apn_lbs_fwd": (I, [P, P, F, P, P, P, P, I, I, P, P, P, P, P, P]),
    "apn_lbs_bwd_workspace_bytes": (SZ, [I, I]),
    "apn_lbs_bwd": (I, [P, P, F, P, P, P, I, I, P, P, P, P, P, P, P, P, P, P, SZ, P]),
    "apn_grid_workspace_bytes": (SZ, [I, I]),
    "apn_grid_build": (I, [P, P, I, F, F, F, I, P, SZ, P]),
    "apn_grid_describe": (I, [P, P, P]),
    "apn_scan_workspace_bytes": (SZ, [I]),
    "apn_exclusive_scan_i32": (I, [P, P, I, P, SZ, P]),
    "apn_ray_candidates": (I, [P, P, I, F, F, F, P, I, P, P, P, P, P]),
    "apn_knn": (I, [P, P, F, F, F, P, P, P, I, P, P, P, P]),
    "apn_knn_sorted_workspace_bytes": (SZ, [I]),
    "apn_knn_sorted": (I, [P, P, F, F, F, P, P, P, I, P, P, P, P, SZ, P]),
    "apn_compact_samples": (I, [P, P, F, F, F, P, P, P, P, P, P, P, I, I, P, P, P, P, P, P]),
    "apn_knn_points": (I, [P, I, P, I, P, P, P]),
    "apn_nn1_batched": (I, [P, P, I, I, I, I, P, P]),
    "apn_rays_of_a_view": (I, [P, P, I, I, I, I, I, I, P, LL, I, P, P, P, P]),
    "apn_point_regularisers": (I, [P, P, P, P, I, I, I, F, F, F, F, P, P, P, P]),
    "apn_pose_regularisers": (I, [P, P, P, P, I, I, F, F, P, P, P, P, P]),
    "apn_time_embed": (I, [P, P, I, P, P]),
    "apn_mse_loss_grad": (I, [P, P, I, F, P, P, P]),
    "apn_aggregate_scratch_bytes": (SZ, [I, I]),
    "apn_aggregate_fwd": (I, [P, P, P, P, SZ, P]),
    "apn_aggregate_bwd_scratch_bytes": (SZ, [I, I]),
    "apn_aggregate_bwd": (I, [P, P, P, P, P, SZ, P]),
    "apn_aggregate_tc_weights_bytes": (SZ, [I]),
    "apn_aggregate_tc_pack_weights": (I, [P, I, P, P]),
    "apn_aggregate_tc_scratch_bytes": (SZ, [I]),
    "apn_aggregate_tc_point_table": (I, [P, P, I, I, P, P]),
    "apn_aggregate_tc_tape_bytes": (SZ, [I]),
    "apn_aggregate_fwd_tc": (I, [P, P, P, P, P, I, P, SZ, P, SZ, P]),
    "apn_aggregate_tc_bwd_weights_bytes": (SZ, []),
    "apn_aggregate_tc_pack_weights_bwd": (I, [P, I, P, P]),
    "apn_aggregate_tc_bwd_scratch_bytes": (SZ, [I, I]),
    "apn_aggregate_bwd_tc": (I, [P, P, P, P, P, P, P, SZ, P]),
    "apn_aggregate_bwd_tc_phase": (I, [P, P, P, P, P, P, P, SZ, I, P]),
    "apn_composite_fwd": (I, [P, P, P, P, I, P, I, F, F, P, P, P, P, P, P, P]),
    "apn_composite_bwd": (I, [P, P, P, P, I, F, F, P, P, P, P, P, P, P, P, P]),
    "apn_adam_step_size": (F, [I, F, F, F]),
    "apn_adam_multi": (I, [P, I, F, F, F, P]),
    "apn_adam_multi_dev": (I, [P, I, F, F, F, P, P, P]),
    "apn_sample_knn_static_workspace_bytes": (SZ, [I, I]),
    "apn_sample_knn_static": (I, [P, P, I, F, F, F, P, I, I, P, SZ, P, P, P, P, P, P, P]),
    "apn_infer_t_minmax": (I, [P, P, P, P, F, F, I, P, P, P]),
    "apn_infer_n_samples": (I, [P, P, F, I, P, P]),
    "apn_infer_ray_start_dir": (I, [P, P, P, I, P, P, P]),
    "apn_sample_pts_on_rays_fill": (I, [P, P, P, P, P, I, LL, F, P, P, P, P, P]),
    "apn_raw2alpha": (I, [P, F, F, LL, P, P, P]),
    "apn_raw2alpha_backward": (I, [P, P, F, LL, P, P]),
    "apn_alpha2weight": (I, [P, P, LL, I, P, P, P, P, P, P]),
    "apn_alpha2weight_backward": (I, [P, P, P, P, P, P, I, P, P, P, P]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the library once; raises ApnError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ApnError(f"{LIB_PATH} is missing: run `python -m articulated_point_nerf_b200.build` "
                           "(there is no fallback path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().apn_last_error().decode(errors="replace")
        raise ApnError(f"{what or 'libapn_sm100'} failed ({rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ApnError("expected a CUDA tensor (the B200 path has no CPU fallback)")
    if not t.is_contiguous():
        raise ApnError("expected a contiguous tensor")
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the stream every kernel is enqueued on)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().apn_launch_count())


class _StageTimer:
    """Optional CUDA-event brackets around the C-ABI call groups (bench.py's per-stage breakdown and the
    roofline of the dominant kernel group).  Disabled by default: zero overhead on the product path."""

    def __init__(self):
        self.enabled = False
        self.records = []          # (name, start_event, end_event)

    def reset(self, enabled: bool):
        self.enabled = enabled
        self.records = []

    def totals(self):
        """name -> (calls, total ms); synchronises."""
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.records:
            n, ms = out.get(name, (0, 0.0))
            out[name] = (n + 1, ms + a.elapsed_time(b))
        return out


STAGES = _StageTimer()

# NVTX ranges around the same call groups (apn_range_push / apn_range_pop): on with APN_NVTX=1 or nvtx(True).  Each name is the
# reference's torch.profiler.record_function name where the group is one reference range, and the reference names joined
# with '+' where one launch group covers several (lib/temporalpoints.py:421-653, lib/pointwarper.py:217-241):
NVTX_NAMES = {
    "transform_net": b"poc_fre+transform_net+calc_rec_abs_T",      # pose chain (one fused cluster kernel)
    "forward_warp": b"forward_warp+weighted_G_tw",                  # skinning-weight softmax/merge + LBS
    "grid_build": b"grid_build",                                    # no reference range (KeOps has no build step)
    "sample_ray+knn": b"sample_ray+knn+knn-post",
    "feat_net": b"feat_net+densitynet+rgbnet",
    "Alphas2Weights": b"pre-mask+Alphas2Weights+post-mask+segment_coo",
}
_nvtx_on = bool(os.environ.get("APN_NVTX"))


def nvtx(enabled: bool) -> None:
    global _nvtx_on
    _nvtx_on = bool(enabled)



class stage:
    """with stage("aggregate_fwd"): ...   (names follow the reference's profiler ranges where one exists)"""
    __slots__ = ("name", "a")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _nvtx_on:
            load().apn_range_push(NVTX_NAMES.get(self.name) or self.name.encode())
        if STAGES.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _nvtx_on:
            load().apn_range_pop()
        if STAGES.enabled:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            STAGES.records.append((self.name, self.a, b))
        return False
