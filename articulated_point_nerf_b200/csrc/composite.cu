// K4 — ray compositing, fused: alpha>thres mask -> transmittance chain with early stop at
// T<1e-3 -> weight>thres mask -> sums of rgb / depth / extra channels -> + alphainv_last*bg.
// Replaces lib/temporalpoints.py:611-677 (two threshold compactions, Alphas2Weights,
// three torch_scatter.segment_coo sums) and lib/cuda/render_utils_kernel.cu:431-531.
//
// One thread owns one ray and walks its contiguous, near-to-far sample segment.  The chain
//   T[i] = Tc;  w[i] = Tc*alpha[i];  Tc = float(double(Tc) * (1.0 - double(alpha[i])))
// is evaluated in exactly the reference's order and mixed precision
// (render_utils_kernel.cu:445-457: `float T_cum; T_cum *= (1. - alpha[i])`), because the
// early-stop decision `T_cum < 1e-3` selects which samples contribute; a re-associated warp
// scan would change those bits.  Adjacent threads own adjacent segments of the same arrays,
// so the sequential walks of a warp stay inside a few shared cache lines.
// Sums are accumulated near-to-far (deterministic, unlike segment_coo's atomics).
//
// Algorithmic bytes: fwd M*(4 alpha + 12 rgb + 4 step) + M*4 (T saved) + R*(4 + 20);
//                    bwd M*(4+12+4+4) + M*16 + R*(4+4+16).
#include "common.cuh"

#define COMP_MAX_EXTRA 4

__global__ void __launch_bounds__(128)
composite_fwd_kernel(const float* __restrict__ alpha, const float* __restrict__ rgb, const int* __restrict__ step_id,
                     const float* __restrict__ extra, int n_extra, const int* __restrict__ ray_start, int R, float thres,
                     float bg, float* __restrict__ rgb_marched, float* __restrict__ alphainv_last, float* __restrict__ depth,
                     float* __restrict__ extra_marched, float* __restrict__ T_save, int* __restrict__ n_used) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int s = ray_start[r], e = ray_start[r + 1];
  float Tc = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f;
  float ex[COMP_MAX_EXTRA] = {0.f, 0.f, 0.f, 0.f};
  int i = s;
  for (; i < e; ++i) {
    const float a = alpha[i];
    if (!(a > thres)) {               // dropped by the pre-mask: not part of the chain
      if (T_save) T_save[i] = 1.f;
      continue;
    }
    if (T_save) T_save[i] = Tc;
    const float w = __fmul_rn(Tc, a);
    if (w > thres) {
      cr = __fadd_rn(cr, __fmul_rn(w, rgb[3 * (size_t)i]));
      cg = __fadd_rn(cg, __fmul_rn(w, rgb[3 * (size_t)i + 1]));
      cb = __fadd_rn(cb, __fmul_rn(w, rgb[3 * (size_t)i + 2]));
      if (step_id) dep = __fadd_rn(dep, __fmul_rn(w, (float)step_id[i]));
      for (int c = 0; c < n_extra; ++c) ex[c] = __fadd_rn(ex[c], __fmul_rn(w, extra[(size_t)i * n_extra + c]));
    }
    Tc = (float)((double)Tc * (1.0 - (double)a));
    if ((double)Tc < 1e-3) {
      ++i;
      break;
    }
  }
  if (n_used) n_used[r] = i - s;      // samples visited before the early stop
  if (T_save)
    for (int k = i; k < e; ++k) T_save[k] = 1.f;
  rgb_marched[3 * (size_t)r] = __fadd_rn(cr, __fmul_rn(Tc, bg));
  rgb_marched[3 * (size_t)r + 1] = __fadd_rn(cg, __fmul_rn(Tc, bg));
  rgb_marched[3 * (size_t)r + 2] = __fadd_rn(cb, __fmul_rn(Tc, bg));
  alphainv_last[r] = Tc;
  if (depth) depth[r] = dep;
  if (extra_marched)
    for (int c = 0; c < n_extra; ++c) extra_marched[(size_t)r * n_extra + c] = __fadd_rn(ex[c], __fmul_rn(Tc, bg));
}

// Backward of the fused op.  With gw[i] = d(rgb_marched)·rgb[i] + d(depth)*step[i] for samples
// that passed the weight mask (0 otherwise) and gl = d(alphainv_last) + bg*sum_c d(rgb_marched)_c:
//   back = gl*alphainv_last;  for i reversed over the chain:
//     d_alpha[i] = gw[i]*T[i] - back/(1-alpha[i]+1e-10);  back += gw[i]*w[i]
// (render_utils_kernel.cu:520-531, same float/double mixing).  d_rgb[i] = w[i]*d(rgb_marched).
__global__ void __launch_bounds__(128)
composite_bwd_kernel(const float* __restrict__ alpha, const float* __restrict__ rgb, const int* __restrict__ step_id,
                     const int* __restrict__ ray_start, int R, float thres, float bg, const float* __restrict__ T_save,
                     const int* __restrict__ n_used, const float* __restrict__ alphainv_last,
                     const float* __restrict__ d_rgb_marched, const float* __restrict__ d_alphainv_last,
                     const float* __restrict__ d_depth, float* __restrict__ d_alpha, float* __restrict__ d_rgb) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int s = ray_start[r], e = ray_start[r + 1];
  const int stop = s + n_used[r];
  const float gr = d_rgb_marched ? d_rgb_marched[3 * (size_t)r] : 0.f;
  const float gg = d_rgb_marched ? d_rgb_marched[3 * (size_t)r + 1] : 0.f;
  const float gb = d_rgb_marched ? d_rgb_marched[3 * (size_t)r + 2] : 0.f;
  const float gd = d_depth ? d_depth[r] : 0.f;
  float gl = d_alphainv_last ? d_alphainv_last[r] : 0.f;
  gl += bg * (gr + gg + gb);
  for (int i = stop; i < e; ++i) {    // never visited: no gradient
    d_alpha[i] = 0.f;
    d_rgb[3 * (size_t)i] = 0.f; d_rgb[3 * (size_t)i + 1] = 0.f; d_rgb[3 * (size_t)i + 2] = 0.f;
  }
  float back = __fmul_rn(gl, alphainv_last[r]);
  for (int i = stop - 1; i >= s; --i) {
    const float a = alpha[i];
    float da = 0.f, dr = 0.f, dg = 0.f, db = 0.f;
    if (a > thres) {
      const float T = T_save[i];
      const float w = __fmul_rn(T, a);
      float gw = 0.f;
      if (w > thres) {
        gw = gr * rgb[3 * (size_t)i] + gg * rgb[3 * (size_t)i + 1] + gb * rgb[3 * (size_t)i + 2];
        if (step_id) gw += gd * (float)step_id[i];
        dr = w * gr; dg = w * gg; db = w * gb;
      }
      da = (float)((double)__fmul_rn(gw, T) - (double)back / ((double)(1.f - a) + 1e-10));
      back = __fadd_rn(back, __fmul_rn(gw, w));
    }
    d_alpha[i] = da;
    d_rgb[3 * (size_t)i] = dr; d_rgb[3 * (size_t)i + 1] = dg; d_rgb[3 * (size_t)i + 2] = db;
  }
}

extern "C" int apn_composite_fwd(const float* alpha, const float* rgb, const int32_t* step_id, const float* extra,
                                 int n_extra, const int32_t* ray_start, int R, float thres, float bg, float* rgb_marched,
                                 float* alphainv_last, float* depth, float* extra_marched, float* T_save, int32_t* n_used,
                                 apn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (R <= 0) return 0;
  APN_CHECK_ARG(ray_start && rgb_marched && alphainv_last, "null pointer");
  APN_CHECK_ARG(n_extra >= 0 && n_extra <= COMP_MAX_EXTRA, "0 <= n_extra <= 4");
  APN_CHECK_ARG(n_extra == 0 || extra_marched, "extra channels need an output buffer");   // `extra` is NULL when M == 0
  composite_fwd_kernel<<<apn_div_up(R, 128), 128, 0, stream>>>(alpha, rgb, step_id, extra, n_extra, ray_start, R, thres, bg,
                                                              rgb_marched, alphainv_last, depth, extra_marched, T_save, n_used);
  APN_LAUNCH_CHECK();
  return 0;
}

extern "C" int apn_composite_bwd(const float* alpha, const float* rgb, const int32_t* step_id, const int32_t* ray_start,
                                 int R, float thres, float bg, const float* T_save, const int32_t* n_used,
                                 const float* alphainv_last, const float* d_rgb_marched, const float* d_alphainv_last,
                                 const float* d_depth, float* d_alpha, float* d_rgb, apn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (R <= 0) return 0;
  APN_CHECK_ARG(ray_start && n_used && alphainv_last, "null pointer");   // per-sample arrays are NULL when M == 0
  composite_bwd_kernel<<<apn_div_up(R, 128), 128, 0, stream>>>(alpha, rgb, step_id, ray_start, R, thres, bg, T_save, n_used,
                                                              alphainv_last, d_rgb_marched, d_alphainv_last, d_depth, d_alpha,
                                                              d_rgb);
  APN_LAUNCH_CHECK();
  return 0;
}
