// K3 (exact path) — neighbour aggregation in fp32 on CUDA cores: gather + canonical-frame rotation
// + positional encoding -> feat_net (4 Linear+LeakyReLU) -> inverse-distance reduce over the 8
// neighbours -> densitynet/Raw2Alpha and RGBNet heads; and its backward.
// Replaces lib/temporalpoints.py:446-515, lib/tineuvox.py:65-88,158,396-400,646-670,872-878 and
// lib/cuda/render_utils_kernel.cu:358-428.
//
// This is the parity / training path (every product in fp32, activations saved for backward).
// The tcgen05 inference path lives in aggregate_tc.cu and shares the prologue maths below.
//
// Row layout: MLP row 8*m + k is neighbour k of kept sample m.  x0 row = [PE(rel_c) 63 | feat 128 |
// pose 0/64 | zero pad] with leading dimension ld0 = round_up(d_in, 4), i.e. the reference's own
// column order so that torch-layout weights are consumed unchanged.
#include "tgemm.cuh"

#define AGG_C APN_C
#define AGG_K APN_K
#define AGG_FV_LD 160   // [rgb feature 128 | view PE 27 | pad 5]
#define AGG_V0 64

static inline int agg_ld0(int d_in) { return (d_in + 3) & ~3; }

// ---------------------------------------------------------------------------------------
// prologue: one warp per kept sample
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
agg_prep_kernel(const apn_agg_inputs in, int ld0, float* __restrict__ x0, float* __restrict__ idw,
                float* __restrict__ alpha_direct, float* __restrict__ rgb_direct) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < in.M; m += warps) {
    const float px = __ldg(in.pts + 3 * (size_t)m), py = __ldg(in.pts + 3 * (size_t)m + 1), pz = __ldg(in.pts + 3 * (size_t)m + 2);
    // lanes 0..7 own neighbour `lane` for the per-sample weights
    const int kk = lane & 7;
    const int my_idx = __ldg(in.nn_idx + (size_t)m * AGG_K + kk);
    const float rx = px - __ldg(in.xyz + 3 * (size_t)my_idx), ry = py - __ldg(in.xyz + 3 * (size_t)my_idx + 1),
                rz = pz - __ldg(in.xyz + 3 * (size_t)my_idx + 2);
    const float d2 = (rx * rx + ry * ry) + rz * rz;
    // inverse-distance weights (lib/temporalpoints.py:473-475)
    const float u = 1.0f / (d2 + in.eps);
    float su = u;
    su += __shfl_xor_sync(0xffffffffu, su, 1);
    su += __shfl_xor_sync(0xffffffffu, su, 2);
    su += __shfl_xor_sync(0xffffffffu, su, 4);
    const float w = u / su;
    if (lane < AGG_K) idw[(size_t)m * AGG_K + lane] = w;
    // direct branch (lib/temporalpoints.py:459-470)
    if (alpha_direct) {
      const float sig = in.mean_min_distance * fmaxf(__ldg(in.direct_eps + my_idx), 0.f);
      const float wd = expf(-(d2 * d2) / (2.f * sig * sig + 1e-12f));
      float sw = wd;
      sw += __shfl_xor_sync(0xffffffffu, sw, 1);
      sw += __shfl_xor_sync(0xffffffffu, sw, 2);
      sw += __shfl_xor_sync(0xffffffffu, sw, 4);
      const float wn = wd / (sw + 1e-12f);
      const float ca = fminf(fmaxf(__ldg(in.canonical_alpha + my_idx), 0.f), 1.f);
      float a = (1.0f / AGG_K) * wd * ca;
      float cr = wn * fminf(fmaxf(__ldg(in.canonical_rgbs + 3 * (size_t)my_idx), 0.f), 1.f);
      float cg = wn * fminf(fmaxf(__ldg(in.canonical_rgbs + 3 * (size_t)my_idx + 1), 0.f), 1.f);
      float cb = wn * fminf(fmaxf(__ldg(in.canonical_rgbs + 3 * (size_t)my_idx + 2), 0.f), 1.f);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        cr += __shfl_xor_sync(0xffffffffu, cr, o);
        cg += __shfl_xor_sync(0xffffffffu, cg, o);
        cb += __shfl_xor_sync(0xffffffffu, cb, o);
      }
      if (lane == 0) {
        alpha_direct[m] = a;
        rgb_direct[3 * (size_t)m] = cr; rgb_direct[3 * (size_t)m + 1] = cg; rgb_direct[3 * (size_t)m + 2] = cb;
      }
    }
    // MLP input rows
    for (int k = 0; k < AGG_K; ++k) {
      const int idx = __shfl_sync(0xffffffffu, my_idx, k);
      const float qx = __shfl_sync(0xffffffffu, rx, k), qy = __shfl_sync(0xffffffffu, ry, k), qz = __shfl_sync(0xffffffffu, rz, k);
      float* row = x0 + ((size_t)m * AGG_K + k) * ld0;
      // rel_c = Ginv[idx] * rel_p; lane d (0..2) computes component d, others a copy of component lane%3
      const int d = lane % 3;
      const float* Gi = in.ginv + 9 * (size_t)idx + 3 * d;
      const float rc = __ldg(Gi) * qx + __ldg(Gi + 1) * qy + __ldg(Gi + 2) * qz;
      if (lane < 3) row[lane] = rc;
      if (lane < 30) {
        // poc_fre: column 3 + d*10 + i = sin(rel_c[d] * 2^i), column 33 + d*10 + i = cos(...)
        const int dd = lane / 10, i = lane - dd * 10;
        const float v = __shfl_sync(0x3fffffffu, rc, dd) * (float)(1 << i);
        float s, c;
        sincosf(v, &s, &c);
        row[3 + lane] = s;
        row[33 + lane] = c;
      }
      const float4 f = __ldg(reinterpret_cast<const float4*>(in.feat + (size_t)idx * AGG_C) + lane);
      float* fr = row + APN_PE_POS + 4 * lane;
      fr[0] = f.x; fr[1] = f.y; fr[2] = f.z; fr[3] = f.w;
      for (int c = APN_PE_POS + AGG_C + lane; c < ld0; c += 32)
        row[c] = (c < in.d_in) ? __ldg(in.pose_emb + (c - APN_PE_POS - AGG_C)) : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------
// K-reduce + densitynet + Raw2Alpha + view PE: one warp per sample
// ---------------------------------------------------------------------------------------
// FROM_H: `h` already holds the reduced feature (written by the tensor-core kernel); only the heads' inputs are built
template <bool FROM_H>
__global__ void __launch_bounds__(256)
agg_reduce_kernel(const apn_agg_inputs in, const float* __restrict__ act3, const float* __restrict__ idw,
                  const float* __restrict__ density_w, const float* __restrict__ density_b, float* __restrict__ h,
                  float* __restrict__ exp_d, float* __restrict__ alpha, float* __restrict__ fv) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const float4 wd = __ldg(reinterpret_cast<const float4*>(density_w) + lane);
  const float bd = __ldg(density_b);
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < in.M; m += warps) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FROM_H) {
      acc = reinterpret_cast<const float4*>(h + (size_t)m * AGG_C)[lane];
    } else {
#pragma unroll
      for (int k = 0; k < AGG_K; ++k) {
        const float w = __ldg(idw + (size_t)m * AGG_K + k);
        const float4 a = __ldg(reinterpret_cast<const float4*>(act3 + ((size_t)m * AGG_K + k) * AGG_C) + lane);
        acc.x += a.x * w; acc.y += a.y * w; acc.z += a.z * w; acc.w += a.w * w;
      }
      reinterpret_cast<float4*>(h + (size_t)m * AGG_C)[lane] = acc;
    }
    const float dens = warp_sum(acc.x * wd.x + acc.y * wd.y + acc.z * wd.z + acc.w * wd.w) + bd;
    if (lane == 0) {
      // lib/cuda/render_utils_kernel.cu:358-370
      const float e = expf(dens + in.act_shift);
      exp_d[m] = e;
      alpha[m] = 1.f - powf(1.f + e, -in.interval);
    }
    // view-direction encoding of the sample's ray (lib/tineuvox.py:872-878 with 4 frequencies)
    const int r = __ldg(in.ray_id + m);
    float* f = fv + (size_t)m * AGG_FV_LD + AGG_C;
    if (lane < 3) f[lane] = __ldg(in.viewdirs + 3 * (size_t)r + lane);
    if (lane < 12) {
      const int dd = lane >> 2, i = lane & 3;
      const float v = __ldg(in.viewdirs + 3 * (size_t)r + dd) * (float)(1 << i);
      float s, c;
      sincosf(v, &s, &c);
      f[3 + lane] = s;
      f[15 + lane] = c;
    }
    if (lane >= 27) f[lane] = 0.f;   // pad columns 155..159
  }
}

// rgb = sigmoid(v0 W2^T + b2): one thread per sample
__global__ void agg_rgb_out_kernel(const float* __restrict__ v0, const float* __restrict__ W2, const float* __restrict__ b2, int M,
                                   float* __restrict__ rgb) {
  __shared__ float sW[3 * AGG_V0 + 3];
  for (int i = threadIdx.x; i < 3 * AGG_V0 + 3; i += blockDim.x) sW[i] = (i < 3 * AGG_V0) ? W2[i] : b2[i - 3 * AGG_V0];
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float a0 = sW[3 * AGG_V0], a1 = sW[3 * AGG_V0 + 1], a2 = sW[3 * AGG_V0 + 2];
  const float4* v = reinterpret_cast<const float4*>(v0 + (size_t)m * AGG_V0);
#pragma unroll
  for (int j = 0; j < AGG_V0 / 4; ++j) {
    const float4 x = __ldg(v + j);
    a0 += x.x * sW[4 * j] + x.y * sW[4 * j + 1] + x.z * sW[4 * j + 2] + x.w * sW[4 * j + 3];
    a1 += x.x * sW[AGG_V0 + 4 * j] + x.y * sW[AGG_V0 + 4 * j + 1] + x.z * sW[AGG_V0 + 4 * j + 2] + x.w * sW[AGG_V0 + 4 * j + 3];
    a2 += x.x * sW[2 * AGG_V0 + 4 * j] + x.y * sW[2 * AGG_V0 + 4 * j + 1] + x.z * sW[2 * AGG_V0 + 4 * j + 2] + x.w * sW[2 * AGG_V0 + 4 * j + 3];
  }
  rgb[3 * (size_t)m] = 1.f / (1.f + expf(-a0));
  rgb[3 * (size_t)m + 1] = 1.f / (1.f + expf(-a1));
  rgb[3 * (size_t)m + 2] = 1.f / (1.f + expf(-a2));
}

// K-reduce (unless act3 == NULL: h given) + densitynet/Raw2Alpha + RGBNet; shared by the fp32 and tensor-core paths
int agg_heads_launch(cudaStream_t st, const apn_agg_inputs* in, const apn_mlp_weights* w, const float* act3, const float* idw,
                     float* h, float* exp_d, float* alpha, float* fv, float* v0, float* rgb) {
  const int M = in->M;
  const int wblocks = min(apn_div_up(M, 8), APN_SM_COUNT * 16);
  if (act3)
    agg_reduce_kernel<false><<<wblocks, 256, 0, st>>>(*in, act3, idw, w->density_w, w->density_b, h, exp_d, alpha, fv);
  else
    agg_reduce_kernel<true><<<wblocks, 256, 0, st>>>(*in, nullptr, idw, w->density_w, w->density_b, h, exp_d, alpha, fv);
  APN_LAUNCH_CHECK();
  // RGBNet (lib/tineuvox.py:77-88): feature_linears has no activation
  APN_CHECK_ARG(gemm_forward(st, h, AGG_C, w->rgb_feat_w, AGG_C, w->rgb_feat_b, fv, AGG_FV_LD, M, AGG_C, AGG_C, 1.f) == 0, "gemm rgb feat");
  APN_CHECK_ARG(gemm_forward(st, fv, AGG_FV_LD, w->rgb_v0_w, AGG_C + APN_PE_VIEW, w->rgb_v0_b, v0, AGG_V0, M, AGG_V0,
                             AGG_C + APN_PE_VIEW, 0.f) == 0, "gemm rgb v0");
  agg_rgb_out_kernel<<<apn_div_up(M, 128), 128, 0, st>>>(v0, w->rgb_v2_w, w->rgb_v2_b, M, rgb);
  APN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// scratch layouts
// ---------------------------------------------------------------------------------------
struct AggFwdBufs {
  float *x0, *act[4], *h, *exp_d, *fv, *v0;
  size_t total;
};
static AggFwdBufs agg_fwd_layout(char* base, int M, int d_in) {
  AggFwdBufs b;
  const size_t rows = (size_t)M * AGG_K;
  size_t o = 0;
  auto take = [&](size_t n_float) {
    float* p = (float*)(base + o);
    o = apn_align(o + n_float * sizeof(float));
    return p;
  };
  b.x0 = take(rows * agg_ld0(d_in));
  for (int l = 0; l < 4; ++l) b.act[l] = take(rows * AGG_C);
  b.h = take((size_t)M * AGG_C);
  b.exp_d = take(M);
  b.fv = take((size_t)M * AGG_FV_LD);
  b.v0 = take((size_t)M * AGG_V0);
  b.total = o;
  return b;
}

extern "C" size_t apn_aggregate_scratch_bytes(int M, int d_in) {
  if (M <= 0) return 0;
  return agg_fwd_layout(nullptr, M, d_in).total;
}

static int agg_check_inputs(const apn_agg_inputs* in, const apn_mlp_weights* w) {
  APN_CHECK_ARG(in && w, "null struct");
  APN_CHECK_ARG(in->d_in == APN_PE_POS + AGG_C || (in->d_in > APN_PE_POS + AGG_C && in->d_in <= 256 && in->pose_emb),
                "d_in must be 191, or 192..256 with a pose embedding");
  APN_CHECK_ARG(in->m_dev == nullptr, "the fp32 path takes an exact host-side sample count (m_dev is for the tensor-core entry points)");
  APN_CHECK_ARG(in->pts && in->nn_idx && in->ray_id && in->xyz && in->ginv && in->feat && in->viewdirs, "null input pointer");
  for (int l = 0; l < 4; ++l) APN_CHECK_ARG(w->w[l] && w->b[l], "null feat_net weight");
  APN_CHECK_ARG(w->density_w && w->density_b && w->rgb_feat_w && w->rgb_feat_b && w->rgb_v0_w && w->rgb_v0_b && w->rgb_v2_w &&
                    w->rgb_v2_b, "null head weight");
  return 0;
}

extern "C" int apn_aggregate_fwd(const apn_agg_inputs* in, const apn_mlp_weights* w, const apn_agg_outputs* out, void* scratch,
                                 size_t scratch_bytes, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (agg_check_inputs(in, w)) return -1;
  APN_CHECK_ARG(out && out->alpha && out->rgb && out->idw, "alpha, rgb and idw outputs are required");
  APN_CHECK_ARG((out->alpha_direct == nullptr) == (out->rgb_direct == nullptr), "direct outputs come as a pair");
  APN_CHECK_ARG(!out->alpha_direct || (in->canonical_alpha && in->canonical_rgbs && in->direct_eps), "direct branch inputs missing");
  const int M = in->M;
  if (M <= 0) return 0;
  AggFwdBufs b;
  if (out->x0) {
    APN_CHECK_ARG(out->act[0] && out->act[1] && out->act[2] && out->act[3] && out->h && out->exp_d && out->fv && out->v0,
                  "saved-activation buffers must all be given");
    b.x0 = out->x0;
    for (int l = 0; l < 4; ++l) b.act[l] = out->act[l];
    b.h = out->h; b.exp_d = out->exp_d; b.fv = out->fv; b.v0 = out->v0;
  } else {
    APN_CHECK_ARG(scratch && scratch_bytes >= apn_aggregate_scratch_bytes(M, in->d_in), "scratch too small");
    b = agg_fwd_layout((char*)scratch, M, in->d_in);
  }
  const int ld0 = agg_ld0(in->d_in);
  const int rows = M * AGG_K;
  const int wblocks = min(apn_div_up(M, 8), APN_SM_COUNT * 16);
  agg_prep_kernel<<<wblocks, 256, 0, st>>>(*in, ld0, b.x0, out->idw, out->alpha_direct, out->rgb_direct);
  APN_LAUNCH_CHECK();
  // feat_net (lib/temporalpoints.py:123-130): LeakyReLU(0.01) after every layer
  APN_CHECK_ARG(gemm_forward(st, b.x0, ld0, w->w[0], in->d_in, w->b[0], b.act[0], AGG_C, rows, AGG_C, in->d_in, 0.01f) == 0, "gemm L0");
  for (int l = 1; l < 4; ++l)
    APN_CHECK_ARG(gemm_forward(st, b.act[l - 1], AGG_C, w->w[l], AGG_C, w->b[l], b.act[l], AGG_C, rows, AGG_C, AGG_C, 0.01f) == 0, "gemm L");
  return agg_heads_launch(st, in, w, b.act[3], out->idw, b.h, b.exp_d, out->alpha, b.fv, b.v0, out->rgb);
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
// d_pre = d_rgb * rgb (1-rgb); dW2 += d_pre^T v0; db2 += sum d_pre; d_v0 = relu'(v0) * (d_pre W2)
__global__ void __launch_bounds__(256)
agg_rgb_out_bwd_kernel(const float* __restrict__ d_rgb, const float* __restrict__ rgb, const float* __restrict__ v0,
                       const float* __restrict__ W2, int M_cap, const int32_t* __restrict__ m_dev, float* __restrict__ d_v0,
                       float* __restrict__ dW2, float* __restrict__ db2) {
  const int M = apn_rt_count(m_dev, M_cap);
  __shared__ float sW[3 * AGG_V0];
  __shared__ float sAcc[3 * AGG_V0 + 3];
  for (int i = threadIdx.x; i < 3 * AGG_V0; i += blockDim.x) sW[i] = W2[i];
  for (int i = threadIdx.x; i < 3 * AGG_V0 + 3; i += blockDim.x) sAcc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  float aw[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}}, ab[3] = {0.f, 0.f, 0.f};
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    float dp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float y = __ldg(rgb + 3 * (size_t)m + c);
      dp[c] = __ldg(d_rgb + 3 * (size_t)m + c) * y * (1.f - y);
      ab[c] += dp[c];
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int j = lane + 32 * hh;
      const float v = __ldg(v0 + (size_t)m * AGG_V0 + j);
      float g = dp[0] * sW[j] + dp[1] * sW[AGG_V0 + j] + dp[2] * sW[2 * AGG_V0 + j];
      d_v0[(size_t)m * AGG_V0 + j] = (v > 0.f) ? g : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) aw[c][hh] += dp[c] * v;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    atomicAdd(&sAcc[c * AGG_V0 + lane], aw[c][0]);
    atomicAdd(&sAcc[c * AGG_V0 + lane + 32], aw[c][1]);
    if (lane == 0) atomicAdd(&sAcc[3 * AGG_V0 + c], ab[c]);   // every lane holds the same ab
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * AGG_V0; i += blockDim.x) atomicAdd(dW2 + i, sAcc[i]);
  if (threadIdx.x < 3) atomicAdd(db2 + threadIdx.x, sAcc[3 * AGG_V0 + threadIdx.x]);
}

// column sums: out[n] += sum_r A[r*lda + n]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ A, int lda, int rows_cap, const int32_t* __restrict__ rows_dev, int N, int rows_per_block,
              float* __restrict__ out) {
  __shared__ float sRed[8][32];
  const int rows = apn_rt_count(rows_dev, rows_cap);
  if (blockIdx.x * rows_per_block >= rows) return;
  const int n = blockIdx.y * 32 + (threadIdx.x & 31);
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + (threadIdx.x >> 5); r < r1; r += 8) s += __ldg(A + (size_t)r * lda + n);
  sRed[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sRed[i][threadIdx.x];
    atomicAdd(out + n, t);
  }
}
static int colsum(cudaStream_t st, const float* A, int lda, int rows, int N, float* out, const int32_t* rows_dev = nullptr) {
  if (rows <= 0) return 0;
  const int rpb = 1024;
  dim3 grid(apn_div_up(rows, rpb), apn_div_up(N, 32));
  colsum_kernel<<<grid, 256, 0, st>>>(A, lda, rows, rows_dev, N, rpb, out);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// RGBNet backward (lib/tineuvox.py:77-88): d_rgb -> weight gradients of the three Linear layers and d_h (M,128) of
// the rgb branch; shared by the fp32 and tensor-core backward paths
int agg_rgbnet_bwd_launch(cudaStream_t st, const apn_agg_inputs* in, const apn_mlp_weights* w, const apn_agg_outputs* sv,
                          const apn_agg_grads* g, float* d_v0, float* d_fv, float* d_h, ApnSide* side, bool tensor_cores) {
  const int M = in->M;
  const int wblocks = min(apn_div_up(M, 8), APN_SM_COUNT * 8);
  const int KV = AGG_C + APN_PE_VIEW;   // 155
  // The chain d_rgb -> d_v0 -> d_fv -> d_h is serial; the two weight gradients (+ bias column sums) only READ d_v0 /
  // d_fv.  On small batches every one of these GEMMs fills less than half the GPU, so with side streams they run beside
  // the chain instead of in it.
  cudaStream_t s0 = side ? side->s[0] : st, s1 = side ? side->s[1] : st;
  const int32_t* md = in->m_dev;         // device-side sample count (or NULL): every row loop / GEMM extent below follows it
  agg_rgb_out_bwd_kernel<<<wblocks, 256, 0, st>>>(g->d_rgb, sv->rgb, sv->v0, w->rgb_v2_w, M, md, d_v0, g->d_rgb_v2_w, g->d_rgb_v2_b);
  APN_LAUNCH_CHECK();
  if (side) {
    APN_CUDA(cudaEventRecord(side->fork[0], st));
    APN_CUDA(cudaStreamWaitEvent(s0, side->fork[0], 0));
  }
  // tensor_cores (the tensor-core training path): the same four products as 3xTF32 GEMMs (tgemm.cuh); the exact fp32 path
  // keeps the CUDA-core SGEMMs
  auto wgrad = tensor_cores ? tgemm_wgrad : gemm_wgrad;
  auto dgrad = tensor_cores ? tgemm_dgrad : gemm_dgrad;
  APN_CHECK_ARG(wgrad(s0, d_v0, AGG_V0, sv->fv, AGG_FV_LD, g->d_rgb_v0_w, KV, M, AGG_V0, KV, md) == 0, "wgrad v0");
  APN_CHECK_ARG(colsum(s0, d_v0, AGG_V0, M, AGG_V0, g->d_rgb_v0_b, md) == 0, "colsum v0");
  APN_CHECK_ARG(dgrad(st, d_v0, AGG_V0, w->rgb_v0_w, KV, d_fv, AGG_FV_LD, M, KV, AGG_V0, nullptr, 0, 1.f, md) == 0, "dgrad v0");
  if (side) {
    APN_CUDA(cudaEventRecord(side->fork[1], st));
    APN_CUDA(cudaStreamWaitEvent(s1, side->fork[1], 0));
  }
  APN_CHECK_ARG(wgrad(s1, d_fv, AGG_FV_LD, sv->h, AGG_C, g->d_rgb_feat_w, AGG_C, M, AGG_C, AGG_C, md) == 0, "wgrad rgb feat");
  APN_CHECK_ARG(colsum(s1, d_fv, AGG_FV_LD, M, AGG_C, g->d_rgb_feat_b, md) == 0, "colsum rgb feat");
  APN_CHECK_ARG(dgrad(st, d_fv, AGG_FV_LD, w->rgb_feat_w, AGG_C, d_h, AGG_C, M, AGG_C, AGG_C, nullptr, 0, 1.f, md) == 0, "dgrad rgb feat");
  return 0;
}

// one warp per sample: raw2alpha/densitynet backward, d_h -> d_act3 (masked), idw backward -> d_d2
__global__ void __launch_bounds__(256)
agg_reduce_bwd_kernel(const apn_agg_inputs in, const float* __restrict__ act3, const float* __restrict__ idw,
                      const float* __restrict__ h, const float* __restrict__ exp_d, const float* __restrict__ density_w,
                      const float* __restrict__ d_alpha, const float* __restrict__ d_h_in, const float* __restrict__ xyz,
                      float* __restrict__ d_act3, float* __restrict__ d_d2, float* __restrict__ d_density_w,
                      float* __restrict__ d_density_b) {
  __shared__ float sAcc[AGG_C + 1];
  for (int i = threadIdx.x; i < AGG_C + 1; i += blockDim.x) sAcc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const float4 wd = __ldg(reinterpret_cast<const float4*>(density_w) + lane);
  float4 a_dw = make_float4(0.f, 0.f, 0.f, 0.f);
  float a_db = 0.f;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < in.M; m += warps) {
    // lib/cuda/render_utils_kernel.cu:396-406
    const float e = __ldg(exp_d + m);
    const float dd = (float)(fmin((double)e, 1e10) * (double)powf(1.f + e, -in.interval - 1.f) * (double)in.interval *
                             (double)__ldg(d_alpha + m));
    const float4 hv = __ldg(reinterpret_cast<const float4*>(h + (size_t)m * AGG_C) + lane);
    a_dw.x += dd * hv.x; a_dw.y += dd * hv.y; a_dw.z += dd * hv.z; a_dw.w += dd * hv.w;
    a_db += dd;
    float4 dh = __ldg(reinterpret_cast<const float4*>(d_h_in + (size_t)m * AGG_C) + lane);
    dh.x += dd * wd.x; dh.y += dd * wd.y; dh.z += dd * wd.z; dh.w += dd * wd.w;
    float dw_k[AGG_K];
#pragma unroll
    for (int k = 0; k < AGG_K; ++k) {
      const size_t row = (size_t)m * AGG_K + k;
      const float w = __ldg(idw + row);
      const float4 a = __ldg(reinterpret_cast<const float4*>(act3 + row * AGG_C) + lane);
      dw_k[k] = warp_sum(dh.x * a.x + dh.y * a.y + dh.z * a.z + dh.w * a.w);
      float4 g;
      g.x = w * dh.x * (a.x > 0.f ? 1.f : 0.01f);
      g.y = w * dh.y * (a.y > 0.f ? 1.f : 0.01f);
      g.z = w * dh.z * (a.z > 0.f ? 1.f : 0.01f);
      g.w = w * dh.w * (a.w > 0.f ? 1.f : 0.01f);
      reinterpret_cast<float4*>(d_act3 + row * AGG_C)[lane] = g;
    }
    // idw_k = u_k / S, u_k = 1/(d2_k + eps)
    float dot = 0.f, wk = 0.f, dwk = 0.f;
#pragma unroll
    for (int k = 0; k < AGG_K; ++k) {
      const float w = __ldg(idw + (size_t)m * AGG_K + k);
      dot += dw_k[k] * w;
      if (lane == k) { wk = w; dwk = dw_k[k]; }
    }
    if (lane < AGG_K) {
      const int idx = __ldg(in.nn_idx + (size_t)m * AGG_K + lane);
      const float rx = __ldg(in.pts + 3 * (size_t)m) - __ldg(xyz + 3 * (size_t)idx);
      const float ry = __ldg(in.pts + 3 * (size_t)m + 1) - __ldg(xyz + 3 * (size_t)idx + 1);
      const float rz = __ldg(in.pts + 3 * (size_t)m + 2) - __ldg(xyz + 3 * (size_t)idx + 2);
      const float u = 1.0f / ((rx * rx + ry * ry) + rz * rz + in.eps);
      // S = u / w ; d_u = (d_w - dot) / S ; d_d2 = -d_u * u^2
      const float S = u / wk;
      d_d2[(size_t)m * AGG_K + lane] = -((dwk - dot) / S) * u * u;
    }
  }
  atomicAdd(&sAcc[4 * lane], a_dw.x);
  atomicAdd(&sAcc[4 * lane + 1], a_dw.y);
  atomicAdd(&sAcc[4 * lane + 2], a_dw.z);
  atomicAdd(&sAcc[4 * lane + 3], a_dw.w);
  if (lane == 0) atomicAdd(&sAcc[AGG_C], a_db);
  __syncthreads();
  for (int i = threadIdx.x; i < AGG_C; i += blockDim.x) atomicAdd(d_density_w + i, sAcc[i]);
  if (threadIdx.x == 0) atomicAdd(d_density_b, sAcc[AGG_C]);
}

// one warp per sample: d_x0 -> d_feat (scatter), PE backward -> d_rel_c -> d_ginv, d_xyz
__global__ void __launch_bounds__(256)
agg_scatter_bwd_kernel(const apn_agg_inputs in, int ld0, const float* __restrict__ x0, const float* __restrict__ d_x0,
                       const float* __restrict__ d_d2, float* __restrict__ d_xyz, float* __restrict__ d_ginv,
                       float* __restrict__ d_feat) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < in.M; m += warps) {
    const float px = __ldg(in.pts + 3 * (size_t)m), py = __ldg(in.pts + 3 * (size_t)m + 1), pz = __ldg(in.pts + 3 * (size_t)m + 2);
    for (int k = 0; k < AGG_K; ++k) {
      const size_t rowi = (size_t)m * AGG_K + k;
      const int idx = __ldg(in.nn_idx + rowi);
      const float* xr = x0 + rowi * ld0;
      const float* dr = d_x0 + rowi * ld0;
      if (d_feat) {
        const float* s = dr + APN_PE_POS + 4 * lane;
        float* t = d_feat + (size_t)idx * AGG_C + 4 * lane;
        atomicAdd(t, __ldg(s));
        atomicAdd(t + 1, __ldg(s + 1));
        atomicAdd(t + 2, __ldg(s + 2));
        atomicAdd(t + 3, __ldg(s + 3));
      }
      // PE backward: d rel_c[d] = dx[d] + sum_i 2^i (cos_i dsin_i - sin_i dcos_i)
      float c0 = 0.f;
      if (lane < 30) {
        const int i = lane % 10;
        c0 = (float)(1 << i) * (__ldg(xr + 33 + lane) * __ldg(dr + 3 + lane) - __ldg(xr + 3 + lane) * __ldg(dr + 33 + lane));
      }
      const int dd = lane / 10;
      float dc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) dc[d] = warp_sum((dd == d && lane < 30) ? c0 : 0.f) + __ldg(dr + d);
      const float rx = px - __ldg(in.xyz + 3 * (size_t)idx), ry = py - __ldg(in.xyz + 3 * (size_t)idx + 1),
                  rz = pz - __ldg(in.xyz + 3 * (size_t)idx + 2);
      const float rp[3] = {rx, ry, rz};
      if (lane < 9) {
        // d_ginv[r][c] += d_rel_c[r] * rel_p[c]
        const int r = lane / 3, c = lane - 3 * r;
        if (d_ginv) atomicAdd(d_ginv + 9 * (size_t)idx + lane, dc[r] * rp[c]);
      } else if (lane < 12) {
        // d_rel_p[c] = sum_r Ginv[r][c] d_rel_c[r] + 2 rel_p[c] d_d2 ; d_xyz[idx] -= d_rel_p
        const int c = lane - 9;
        const float* G = in.ginv + 9 * (size_t)idx;
        const float g = __ldg(G + c) * dc[0] + __ldg(G + 3 + c) * dc[1] + __ldg(G + 6 + c) * dc[2] +
                        2.f * rp[c] * __ldg(d_d2 + rowi);
        if (d_xyz) atomicAdd(d_xyz + 3 * (size_t)idx + c, -g);
      }
    }
  }
}

struct AggBwdBufs {
  float *d_act_a, *d_act_b, *d_x0, *d_h, *d_fv, *d_v0, *d_d2;
  size_t total;
};
static AggBwdBufs agg_bwd_layout(char* base, int M, int d_in) {
  AggBwdBufs b;
  const size_t rows = (size_t)M * AGG_K;
  size_t o = 0;
  auto take = [&](size_t n_float) {
    float* p = (float*)(base + o);
    o = apn_align(o + n_float * sizeof(float));
    return p;
  };
  b.d_act_a = take(rows * AGG_C);
  b.d_act_b = take(rows * AGG_C);
  b.d_x0 = take(rows * agg_ld0(d_in));
  b.d_h = take((size_t)M * AGG_C);
  b.d_fv = take((size_t)M * AGG_FV_LD);
  b.d_v0 = take((size_t)M * AGG_V0);
  b.d_d2 = take((size_t)M * AGG_K);
  b.total = o;
  return b;
}
extern "C" size_t apn_aggregate_bwd_scratch_bytes(int M, int d_in) {
  if (M <= 0) return 0;
  return agg_bwd_layout(nullptr, M, d_in).total;
}

extern "C" int apn_aggregate_bwd(const apn_agg_inputs* in, const apn_mlp_weights* w, const apn_agg_outputs* sv,
                                 const apn_agg_grads* g, void* scratch, size_t scratch_bytes, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (agg_check_inputs(in, w)) return -1;
  APN_CHECK_ARG(sv && g, "null struct");
  const int M = in->M;
  if (M <= 0) return 0;
  APN_CHECK_ARG(sv->x0 && sv->act[0] && sv->act[1] && sv->act[2] && sv->act[3] && sv->h && sv->exp_d && sv->fv && sv->v0 &&
                    sv->idw && sv->rgb, "saved forward tensors missing");
  APN_CHECK_ARG(g->d_alpha && g->d_rgb, "incoming gradients missing");
  for (int l = 0; l < 4; ++l) APN_CHECK_ARG(g->d_w[l] && g->d_b[l], "null feat_net gradient buffer");
  APN_CHECK_ARG(g->d_density_w && g->d_density_b && g->d_rgb_feat_w && g->d_rgb_feat_b && g->d_rgb_v0_w && g->d_rgb_v0_b &&
                    g->d_rgb_v2_w && g->d_rgb_v2_b, "null head gradient buffer");
  APN_CHECK_ARG(scratch && scratch_bytes >= apn_aggregate_bwd_scratch_bytes(M, in->d_in), "scratch too small");
  const AggBwdBufs b = agg_bwd_layout((char*)scratch, M, in->d_in);
  const int ld0 = agg_ld0(in->d_in);
  const int rows = M * AGG_K;
  const int wblocks = min(apn_div_up(M, 8), APN_SM_COUNT * 8);
  const int KV = AGG_C + APN_PE_VIEW;   // 155
  // RGBNet
  if (agg_rgbnet_bwd_launch(st, in, w, sv, g, b.d_v0, b.d_fv, b.d_h)) return -1;
  // densitynet + K-reduce
  agg_reduce_bwd_kernel<<<wblocks, 256, 0, st>>>(*in, sv->act[3], sv->idw, sv->h, sv->exp_d, w->density_w, g->d_alpha, b.d_h, in->xyz,
                                                 b.d_act_a, b.d_d2, g->d_density_w, g->d_density_b);
  APN_LAUNCH_CHECK();
  // feat_net layers 3..1
  float* dy = b.d_act_a;
  float* dx = b.d_act_b;
  for (int l = 3; l >= 1; --l) {
    APN_CHECK_ARG(gemm_wgrad(st, dy, AGG_C, sv->act[l - 1], AGG_C, g->d_w[l], AGG_C, rows, AGG_C, AGG_C) == 0, "wgrad L");
    APN_CHECK_ARG(colsum(st, dy, AGG_C, rows, AGG_C, g->d_b[l]) == 0, "colsum L");
    APN_CHECK_ARG(gemm_dgrad(st, dy, AGG_C, w->w[l], AGG_C, dx, AGG_C, rows, AGG_C, AGG_C, sv->act[l - 1], AGG_C, 0.01f) == 0, "dgrad L");
    float* t = dy; dy = dx; dx = t;
  }
  // layer 0
  APN_CHECK_ARG(gemm_wgrad(st, dy, AGG_C, sv->x0, ld0, g->d_w[0], in->d_in, rows, AGG_C, in->d_in) == 0, "wgrad L0");
  APN_CHECK_ARG(colsum(st, dy, AGG_C, rows, AGG_C, g->d_b[0]) == 0, "colsum L0");
  if (g->d_xyz || g->d_ginv || g->d_feat || g->d_pose_emb) {
    // PE and feature columns always; the pose-embedding columns only when their gradient is wanted
    const int n_cols = g->d_pose_emb ? in->d_in : APN_PE_POS + AGG_C;
    APN_CHECK_ARG(gemm_dgrad(st, dy, AGG_C, w->w[0], in->d_in, b.d_x0, ld0, rows, n_cols, AGG_C, nullptr, 0, 1.f) == 0, "dgrad L0");
    if (g->d_pose_emb && in->d_in > APN_PE_POS + AGG_C)
      APN_CHECK_ARG(colsum(st, b.d_x0 + APN_PE_POS + AGG_C, ld0, rows, in->d_in - APN_PE_POS - AGG_C, g->d_pose_emb) == 0, "colsum pose");
    agg_scatter_bwd_kernel<<<wblocks, 256, 0, st>>>(*in, ld0, sv->x0, b.d_x0, b.d_d2, g->d_xyz, g->d_ginv, g->d_feat);
    APN_LAUNCH_CHECK();
  }
  return 0;
}
