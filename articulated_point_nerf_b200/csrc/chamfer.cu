// Small helpers of the training step that would otherwise cost several framework launches each, and the batched NN-1.
//
// Batched exact nearest neighbour (K = 1) between two small point sets per batch item, 2-D or 3-D.
// Replaces the KeOps reductions of the batch chamfer loss, lib/temporalpoints.py:783-787
// (D_ij.argKmin(dim=2, K=1) and D_ij.argKmin(dim=1, K=1) over (B, N, M) squared distances; run.py:659-690 calls it
// with B <= 5 views, N = M = 3000 projected 2-D points).  Brute force on purpose: 5 x 3000 x 3000 pair distances is
// 45 M multiply-adds; the target set is staged through shared memory once per block of queries.
// Contract (same as the k-NN): d2 = (dx*dx + dy*dy) [+ dz*dz] in fp32 without FMA contraction, ties -> lowest index.
#include "common.cuh"

#define NN1_THREADS 256
#define NN1_TILE 1024

template <int DIM>
__global__ void __launch_bounds__(NN1_THREADS)
nn1_batched_kernel(const float* __restrict__ query, const float* __restrict__ target, int n_query, int n_target,
                   int* __restrict__ nn_idx) {
  __shared__ float tile[NN1_TILE * DIM];
  const int b = blockIdx.y;
  const float* q = query + (size_t)b * n_query * DIM;
  const float* t = target + (size_t)b * n_target * DIM;
  const int i = blockIdx.x * NN1_THREADS + threadIdx.x;
  float qc[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) qc[c] = i < n_query ? q[(size_t)i * DIM + c] : 0.f;
  float best = INFINITY;
  int best_j = 0;
  for (int j0 = 0; j0 < n_target; j0 += NN1_TILE) {
    const int n = min(NN1_TILE, n_target - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < n * DIM; e += NN1_THREADS) tile[e] = t[(size_t)j0 * DIM + e];
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      const float dx = __fsub_rn(qc[0], tile[j * DIM]), dy = __fsub_rn(qc[1], tile[j * DIM + 1]);
      float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      if (DIM == 3) {
        const float dz = __fsub_rn(qc[DIM - 1], tile[j * DIM + DIM - 1]);
        d = __fadd_rn(d, __fmul_rn(dz, dz));
      }
      if (d < best) {          // strict: the first (lowest-index) minimum wins; NaN distances never win
        best = d;
        best_j = j0 + j;
      }
    }
  }
  if (i < n_query) nn_idx[(size_t)b * n_query + i] = best_j;
}

extern "C" int apn_nn1_batched(const float* query, const float* target, int n_batch, int n_query, int n_target, int dim,
                               int32_t* nn_idx, apn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  APN_CHECK_ARG(dim == 2 || dim == 3, "dim must be 2 or 3");
  APN_CHECK_ARG(n_batch >= 0 && n_query >= 0 && n_target > 0, "need n_target > 0");
  if (n_batch == 0 || n_query == 0) return 0;
  APN_CHECK_ARG(n_batch <= 65535, "at most 65535 batch items");
  APN_CHECK_ARG(query && target && nn_idx, "null pointer");
  const dim3 grid(apn_div_up(n_query, NN1_THREADS), n_batch);
  if (dim == 2) nn1_batched_kernel<2><<<grid, NN1_THREADS, 0, stream>>>(query, target, n_query, n_target, nn_idx);
  else nn1_batched_kernel<3><<<grid, NN1_THREADS, 0, stream>>>(query, target, n_query, n_target, nn_idx);
  APN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Time embedding of the pose network input (lib/tineuvox.py:872-878 poc_fre on the scalar time, lib/temporalpoints.py:
// 546-550): out = [t, sin(t f_0..f_{F-1}), cos(t f_0..f_{F-1})]  (1 + 2F floats); one launch instead of mul / sin / cos / cat.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void time_embed_kernel(const float* __restrict__ t, const float* __restrict__ freqs, int n_freq, float* __restrict__ out) {
  const int i = threadIdx.x;
  const float tv = t[0];
  if (i == 0) out[0] = tv;
  if (i < n_freq) {
    const float x = __fmul_rn(tv, freqs[i]);
    out[1 + i] = sinf(x);
    out[1 + n_freq + i] = cosf(x);
  }
}

extern "C" int apn_time_embed(const float* t, const float* freqs, int n_freq, float* out, apn_stream_t stream_) {
  APN_CHECK_ARG(t && freqs && out, "null pointer");
  APN_CHECK_ARG(n_freq >= 0 && n_freq <= 1024, "0 <= n_freq <= 1024");
  time_embed_kernel<<<1, n_freq < 32 ? 32 : ((n_freq + 31) & ~31), 0, (cudaStream_t)stream_>>>(t, freqs, n_freq, out);
  APN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Render loss and its gradient (run.py:617-621: weight * mse(rgb_marched, target)):
//   loss = weight * mean((pred - target)^2),  grad = (pred - target) * 2 weight / n
// One block, fixed summation order (deterministic); n is a ray batch (8192 x 3), not a frame.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
mse_loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target, int n, float weight,
                     float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float sred[32];
  const float gscale = 2.0f * weight / (float)n;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = pred[i] - target[i];
    acc = fmaf(d, d, acc);
    grad[i] = d * gscale;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sred[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] = weight * (v / (float)n);
  }
}

extern "C" int apn_mse_loss_grad(const float* pred, const float* target, int n, float weight, float* loss, float* grad,
                                 apn_stream_t stream_) {
  APN_CHECK_ARG(n > 0 && pred && target && loss && grad, "need n > 0 and non-null pointers");
  mse_loss_grad_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(pred, target, n, weight, loss, grad);
  APN_LAUNCH_CHECK();
  return 0;
}
