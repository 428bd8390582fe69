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
