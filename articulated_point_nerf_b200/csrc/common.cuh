// Shared host/device helpers for libapn_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/apn.h"

#define APN_SM_COUNT 148  // B200: 2 dies x 74 SMs

void apn_set_error(const char* fmt, ...);
void apn_count_launch(int n = 1);
// aggregate.cu: [K-reduce of act3 unless act3 == NULL] + densitynet/Raw2Alpha + RGBNet on the reduced feature h
int agg_heads_launch(cudaStream_t st, const apn_agg_inputs* in, const apn_mlp_weights* w, const float* act3, const float* idw,
                     float* h, float* exp_d, float* alpha, float* fv, float* v0, float* rgb);
// heads_tc.cu: densitynet / Raw2Alpha + RGBNet on the tensor cores (inference); `packed` >= heads_tc_weights_bytes()
size_t heads_tc_weights_bytes();
// exp_d / fv / v0: the tape of the fp32 heads backward (training), or NULL
int agg_heads_tc_launch(cudaStream_t st, const apn_agg_inputs* in, const apn_mlp_weights* w, const float* h, void* packed,
                        float* alpha, float* rgb, float* exp_d, float* fv, float* v0);
// runtime.cu: library-owned side streams + events for fork/join inside one entry point
struct ApnSide {
  cudaStream_t s[2];
  cudaEvent_t fork[3], join[2];
};
int apn_side_streams(ApnSide** out);
// aggregate.cu: RGBNet backward: weight gradients + d_h (M,128) of the rgb branch.  With `side`, the weight-gradient
// GEMMs run on the side streams (forked from `st`) and are NOT joined: the caller joins side->join[*] before it returns.
int agg_rgbnet_bwd_launch(cudaStream_t st, const apn_agg_inputs* in, const apn_mlp_weights* w, const apn_agg_outputs* sv,
                          const apn_agg_grads* g, float* d_v0, float* d_fv, float* d_h, ApnSide* side = nullptr, bool tensor_cores = false);

#define APN_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      apn_set_error("%s: %s (%s:%d)", __func__, msg, __FILE__, __LINE__); \
      return -1;                                                   \
    }                                                              \
  } while (0)

#define APN_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      apn_set_error("%s: CUDA error %s (%s:%d)", __func__, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return -2;                                                                        \
    }                                                                                   \
  } while (0)

#define APN_LAUNCH_CHECK()                    \
  do {                                        \
    apn_count_launch();                       \
    APN_CUDA(cudaGetLastError());             \
  } while (0)

static inline int apn_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t apn_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ----------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------
// run-time element count of a capacity-sized array: min(*dev, cap), or cap when no device counter is given
__device__ __forceinline__ int apn_rt_count(const int32_t* dev, int cap) {
  if (!dev) return cap;
  const int n = *reinterpret_cast<const volatile int32_t*>(dev);
  return n < 0 ? 0 : (n < cap ? n : cap);
}
__device__ __forceinline__ float atomic_min_float(float* addr, float v) {
  return (v >= 0.f) ? __int_as_float(atomicMin((int*)addr, __float_as_int(v)))
                    : __uint_as_float(atomicMax((unsigned int*)addr, __float_as_uint(v)));
}
__device__ __forceinline__ float atomic_max_float(float* addr, float v) {
  return (v >= 0.f) ? __int_as_float(atomicMax((int*)addr, __float_as_int(v)))
                    : __uint_as_float(atomicMin((unsigned int*)addr, __float_as_uint(v)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------------
// Ray sampling arithmetic shared by every kernel that needs a sample position.
// Restates lib/cuda/render_utils_kernel.cu:12-73,161-188 with the repo-wide contract that
// every multiply/add is a separate IEEE operation (no FMA contraction), so the torch-CPU
// oracle reproduces the same bits.
// ----------------------------------------------------------------------------------------
struct RaySetup {
  float sx, sy, sz;   // start = o + d * t_min
  float dx, dy, dz;   // normalised direction
  float t_min, t_max;
  int n_steps;
};

__device__ __forceinline__ RaySetup ray_setup(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int r,
                                              const float bmin[3], const float bmax[3], float near, float far,
                                              float stepdist) {
  RaySetup s;
  const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
  const float rx = rays_d[3 * r], ry = rays_d[3 * r + 1], rz = rays_d[3 * r + 2];
  const float vx = (rx == 0.f) ? 1e-6f : rx, vy = (ry == 0.f) ? 1e-6f : ry, vz = (rz == 0.f) ? 1e-6f : rz;
  const float ax = __fdiv_rn(__fsub_rn(bmax[0], ox), vx), ay = __fdiv_rn(__fsub_rn(bmax[1], oy), vy),
              az = __fdiv_rn(__fsub_rn(bmax[2], oz), vz);
  const float bx = __fdiv_rn(__fsub_rn(bmin[0], ox), vx), by = __fdiv_rn(__fsub_rn(bmin[1], oy), vy),
              bz = __fdiv_rn(__fsub_rn(bmin[2], oz), vz);
  s.t_min = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), near);
  s.t_max = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), near);
  const float n = ceilf(__fdiv_rn(__fsub_rn(s.t_max, s.t_min), stepdist));
  s.n_steps = (int)fmaxf(n, 1.f);
  const float rnorm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
  s.sx = __fadd_rn(ox, __fmul_rn(rx, s.t_min));
  s.sy = __fadd_rn(oy, __fmul_rn(ry, s.t_min));
  s.sz = __fadd_rn(oz, __fmul_rn(rz, s.t_min));
  s.dx = __fdiv_rn(rx, rnorm);
  s.dy = __fdiv_rn(ry, rnorm);
  s.dz = __fdiv_rn(rz, rnorm);
  return s;
}

__device__ __forceinline__ void ray_point(const RaySetup& s, int step, float stepdist, float& px, float& py, float& pz) {
  const float dist = __fmul_rn(stepdist, (float)step);
  px = __fadd_rn(s.sx, __fmul_rn(s.dx, dist));
  py = __fadd_rn(s.sy, __fmul_rn(s.dy, dist));
  pz = __fadd_rn(s.sz, __fmul_rn(s.dz, dist));
}

// squared distance under the neighbour contract: (dx*dx + dy*dy) + dz*dz, no FMA
__device__ __forceinline__ float dist2_contract(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
