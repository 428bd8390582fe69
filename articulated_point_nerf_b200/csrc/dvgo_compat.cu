// Reference-compatible single ops: the pybind surface of lib/cuda/render_utils.cpp:144-155
// (infer_t_minmax, infer_n_samples, infer_ray_start_dir, sample_pts_on_rays, raw2alpha(+bwd),
// alpha2weight(+bwd)) for callers that use the ops one by one (lib/tineuvox.py:627-670,
// lib/temporalpoints.py:392).  The fused path (grid_knn.cu, aggregate.cu, composite.cu) does not
// go through these.  int64 ids and the reference's float/double mixing are kept.
#include "common.cuh"

__global__ void t_minmax_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                const float* __restrict__ xyz_min, const float* __restrict__ xyz_max, float near, float far,
                                int R, float* __restrict__ t_min, float* __restrict__ t_max) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float bmin[3] = {xyz_min[0], xyz_min[1], xyz_min[2]}, bmax[3] = {xyz_max[0], xyz_max[1], xyz_max[2]};
  const RaySetup s = ray_setup(rays_o, rays_d, r, bmin, bmax, near, far, 1.f);
  t_min[r] = s.t_min;
  t_max[r] = s.t_max;
}

__global__ void n_samples_kernel(const float* __restrict__ t_min, const float* __restrict__ t_max, float stepdist, int R,
                                 long long* __restrict__ n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  n[r] = (long long)fmaxf(ceilf(__fdiv_rn(__fsub_rn(t_max[r], t_min[r]), stepdist)), 1.f);
}

__global__ void start_dir_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                 const float* __restrict__ t_min, int R, float* __restrict__ start, float* __restrict__ dir) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float rx = rays_d[3 * r], ry = rays_d[3 * r + 1], rz = rays_d[3 * r + 2];
  const float rnorm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
  const float t = t_min[r];
  start[3 * r] = __fadd_rn(rays_o[3 * r], __fmul_rn(rx, t));
  start[3 * r + 1] = __fadd_rn(rays_o[3 * r + 1], __fmul_rn(ry, t));
  start[3 * r + 2] = __fadd_rn(rays_o[3 * r + 2], __fmul_rn(rz, t));
  dir[3 * r] = __fdiv_rn(rx, rnorm);
  dir[3 * r + 1] = __fdiv_rn(ry, rnorm);
  dir[3 * r + 2] = __fdiv_rn(rz, rnorm);
}

// one warp per ray: lanes stride over the ray's steps, so writes of pts/ids are coalesced
__global__ void sample_fill_kernel(const float* __restrict__ start, const float* __restrict__ dir,
                                   const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                                   const long long* __restrict__ cum, int R, float stepdist, float* __restrict__ pts,
                                   unsigned char* __restrict__ mask_out, long long* __restrict__ ray_id,
                                   long long* __restrict__ step_id) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < R; r += gridDim.x * warps_per_block) {
    const long long first = r ? cum[r - 1] : 0, n = cum[r] - first;
    const float sx = start[3 * r], sy = start[3 * r + 1], sz = start[3 * r + 2];
    const float dx = dir[3 * r], dy = dir[3 * r + 1], dz = dir[3 * r + 2];
    for (long long k = lane; k < n; k += 32) {
      const float dist = __fmul_rn(stepdist, (float)k);
      const float px = __fadd_rn(sx, __fmul_rn(dx, dist)), py = __fadd_rn(sy, __fmul_rn(dy, dist)),
                  pz = __fadd_rn(sz, __fmul_rn(dz, dist));
      const long long o = first + k;
      pts[3 * o] = px; pts[3 * o + 1] = py; pts[3 * o + 2] = pz;
      mask_out[o] = (xyz_min[0] > px) | (xyz_min[1] > py) | (xyz_min[2] > pz) | (xyz_max[0] < px) | (xyz_max[1] < py) |
                    (xyz_max[2] < pz);
      ray_id[o] = r;
      step_id[o] = k;
    }
  }
}

__global__ void raw2alpha_kernel(const float* __restrict__ density, float shift, float interval, long long n,
                                 float* __restrict__ exp_d, float* __restrict__ alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = expf(density[i] + shift);
  exp_d[i] = e;
  alpha[i] = 1.f - powf(1.f + e, -interval);
}

__global__ void raw2alpha_bwd_kernel(const float* __restrict__ exp_d, const float* __restrict__ grad_back, float interval,
                                     long long n, float* __restrict__ grad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = exp_d[i];
  grad[i] = (float)(fmin((double)e, 1e10) * (double)powf(1.f + e, -interval - 1.f) * (double)interval * (double)grad_back[i]);
}

__global__ void seg_bounds_kernel(const long long* __restrict__ ray_id, long long n, long long* __restrict__ i_start,
                                  long long* __restrict__ i_end) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i > 0 && ray_id[i] != ray_id[i - 1]) {
    i_start[ray_id[i]] = i;
    i_end[ray_id[i - 1]] = i;
  }
  if (i == n - 1) i_end[ray_id[i]] = n;
}

__global__ void alpha2weight_kernel(const float* __restrict__ alpha, int n_rays, float* __restrict__ weight,
                                    float* __restrict__ T, float* __restrict__ last, long long* __restrict__ i_start,
                                    long long* __restrict__ i_end) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const long long s = i_start[r], e = i_end[r];
  float Tc = 1.f;
  long long i = s;
  for (; i < e; ++i) {
    T[i] = Tc;
    weight[i] = __fmul_rn(Tc, alpha[i]);
    Tc = (float)((double)Tc * (1.0 - (double)alpha[i]));
    if ((double)Tc < 1e-3) {
      ++i;
      break;
    }
  }
  i_end[r] = i;
  last[r] = Tc;
}

__global__ void alpha2weight_bwd_kernel(const float* __restrict__ alpha, const float* __restrict__ weight,
                                        const float* __restrict__ T, const float* __restrict__ last,
                                        const long long* __restrict__ i_start, const long long* __restrict__ i_end,
                                        int n_rays, const float* __restrict__ gw, const float* __restrict__ gl,
                                        float* __restrict__ grad) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const long long s = i_start[r], e = i_end[r];
  float back = __fmul_rn(gl[r], last[r]);
  for (long long i = e - 1; i >= s; --i) {
    grad[i] = (float)((double)__fmul_rn(gw[i], T[i]) - (double)back / ((double)(1.f - alpha[i]) + 1e-10));
    back = __fadd_rn(back, __fmul_rn(gw[i], weight[i]));
  }
}

#define STREAM cudaStream_t stream = (cudaStream_t)stream_

extern "C" int apn_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                                  float near, float far, int R, float* t_min, float* t_max, apn_stream_t stream_) {
  STREAM;
  if (R <= 0) return 0;
  APN_CHECK_ARG(rays_o && rays_d && xyz_min && xyz_max && t_min && t_max, "null pointer");
  t_minmax_kernel<<<apn_div_up(R, 256), 256, 0, stream>>>(rays_o, rays_d, xyz_min, xyz_max, near, far, R, t_min, t_max);
  APN_LAUNCH_CHECK();
  return 0;
}
extern "C" int apn_infer_n_samples(const float* t_min, const float* t_max, float stepdist, int R, int64_t* n_samples,
                                   apn_stream_t stream_) {
  STREAM;
  if (R <= 0) return 0;
  APN_CHECK_ARG(t_min && t_max && n_samples, "null pointer");
  n_samples_kernel<<<apn_div_up(R, 256), 256, 0, stream>>>(t_min, t_max, stepdist, R, (long long*)n_samples);
  APN_LAUNCH_CHECK();
  return 0;
}
extern "C" int apn_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min, int R,
                                       float* rays_start, float* rays_dir, apn_stream_t stream_) {
  STREAM;
  if (R <= 0) return 0;
  APN_CHECK_ARG(rays_o && rays_d && t_min && rays_start && rays_dir, "null pointer");
  start_dir_kernel<<<apn_div_up(R, 256), 256, 0, stream>>>(rays_o, rays_d, t_min, R, rays_start, rays_dir);
  APN_LAUNCH_CHECK();
  return 0;
}
extern "C" int apn_sample_pts_on_rays_fill(const float* rays_start, const float* rays_dir, const float* xyz_min,
                                           const float* xyz_max, const int64_t* n_cumsum, int R, long long total,
                                           float stepdist, float* pts, uint8_t* mask_outbbox, int64_t* ray_id,
                                           int64_t* step_id, apn_stream_t stream_) {
  STREAM;
  if (R <= 0 || total <= 0) return 0;
  APN_CHECK_ARG(rays_start && rays_dir && xyz_min && xyz_max && n_cumsum && pts && mask_outbbox && ray_id && step_id,
                "null pointer");
  const int blocks = min(apn_div_up(R, 8), APN_SM_COUNT * 16);
  sample_fill_kernel<<<blocks, 256, 0, stream>>>(rays_start, rays_dir, xyz_min, xyz_max, (const long long*)n_cumsum, R,
                                                stepdist, pts, mask_outbbox, (long long*)ray_id, (long long*)step_id);
  APN_LAUNCH_CHECK();
  return 0;
}
extern "C" int apn_raw2alpha(const float* density, float shift, float interval, long long n, float* exp_d, float* alpha,
                             apn_stream_t stream_) {
  STREAM;
  if (n <= 0) return 0;
  APN_CHECK_ARG(density && exp_d && alpha, "null pointer");
  raw2alpha_kernel<<<apn_div_up(n, 256), 256, 0, stream>>>(density, shift, interval, n, exp_d, alpha);
  APN_LAUNCH_CHECK();
  return 0;
}
extern "C" int apn_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval, long long n, float* grad,
                                      apn_stream_t stream_) {
  STREAM;
  if (n <= 0) return 0;
  APN_CHECK_ARG(exp_d && grad_back && grad, "null pointer");
  raw2alpha_bwd_kernel<<<apn_div_up(n, 256), 256, 0, stream>>>(exp_d, grad_back, interval, n, grad);
  APN_LAUNCH_CHECK();
  return 0;
}
extern "C" int apn_alpha2weight(const float* alpha, const int64_t* ray_id, long long n_pts, int n_rays, float* weight,
                                float* T, float* alphainv_last, int64_t* i_start, int64_t* i_end, apn_stream_t stream_) {
  STREAM;
  if (n_pts <= 0 || n_rays <= 0) return 0;
  APN_CHECK_ARG(alpha && ray_id && weight && T && alphainv_last && i_start && i_end, "null pointer");
  seg_bounds_kernel<<<apn_div_up(n_pts, 256), 256, 0, stream>>>((const long long*)ray_id, n_pts, (long long*)i_start,
                                                               (long long*)i_end);
  APN_LAUNCH_CHECK();
  alpha2weight_kernel<<<apn_div_up(n_rays, 128), 128, 0, stream>>>(alpha, n_rays, weight, T, alphainv_last,
                                                                  (long long*)i_start, (long long*)i_end);
  APN_LAUNCH_CHECK();
  return 0;
}
extern "C" int apn_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                                         const float* alphainv_last, const int64_t* i_start, const int64_t* i_end,
                                         int n_rays, const float* grad_weights, const float* grad_last, float* grad,
                                         apn_stream_t stream_) {
  STREAM;
  if (n_rays <= 0) return 0;
  APN_CHECK_ARG(alpha && weight && T && alphainv_last && i_start && i_end && grad_weights && grad_last && grad,
                "null pointer");
  alpha2weight_bwd_kernel<<<apn_div_up(n_rays, 128), 128, 0, stream>>>(alpha, weight, T, alphainv_last,
                                                                      (const long long*)i_start, (const long long*)i_end,
                                                                      n_rays, grad_weights, grad_last, grad);
  APN_LAUNCH_CHECK();
  return 0;
}
