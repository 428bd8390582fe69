// fp32 CUDA-core GEMM used by the exact (parity / training) aggregation path.
//   C[M x N] (+)= sum_k A(m,k) * B(k,n)
// Both operands may be stored with either index contiguous, which covers the three shapes the
// decoder MLP needs without materialising a transpose:
//   forward  Y  = act(X W^T + b)     A = X  (k contiguous)   B = W  (k contiguous)
//   dgrad    dX = (dY W) * act'      A = dY (k contiguous)   B = W  (n contiguous)
//   wgrad    dW += dY^T X            A = dY (m contiguous)   B = X  (n contiguous), split over k
// 128x128 tile, 16-deep k slices, 256 threads, 8x8 outputs per thread held as two 4-wide strips
// per dimension (conflict-free float4 shared-memory reads), register-staged prefetch of the next
// k slice.  Accumulation order is fixed per launch configuration (deterministic) except for the
// split-k wgrad, which ends in fp32 atomics.
#pragma once
#include "common.cuh"

enum { GEMM_EPI_BIAS_ACT = 0, GEMM_EPI_MASK = 1, GEMM_EPI_ATOMIC = 2, GEMM_EPI_ACCUM = 3 };   // ACCUM: C += A B (one writer per element)

struct GemmArgs {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int M, N, K;
  const float* bias;   // BIAS_ACT: per-n bias or NULL
  float slope;         // BIAS_ACT: y<0 ? slope*y : y (1 = identity); MASK: derivative for mask<=0
  const float* mask;   // MASK: activation values whose sign selects the derivative (M x N, ld ldm) or NULL
  int ldm;
  int k_chunk;         // ATOMIC: k range per blockIdx.z
  // optional device-side row count (see apn_agg_inputs::m_dev): the true extent of the SAMPLE dimension — M for the
  // forward / dgrad shapes (rows_is_k = 0), K for the wgrad shape (rows_is_k = 1); the host value is then a capacity
  const int32_t* rows_dev;
  int rows_is_k;
};

#define GEMM_BM 128
#define GEMM_BN 128
#define GEMM_BK 16
#define GEMM_LDS (GEMM_BM + 4)

// loads one (128 x 16) operand slice into registers: r[8]
//   KCONTIG: element (row, k) at p[row*ld + k]  -> thread covers rows (tid&63)+64*j, k = (tid>>6)*4 .. +3
//   else   : element (row, k) at p[k*ld + row]  -> thread covers k = (tid>>5)+8*j, rows (tid&31)*4 .. +3
template <bool KCONTIG>
__device__ __forceinline__ void gemm_load_slice(const float* __restrict__ p, int ld, int row0, int n_rows, int k0, int k_end,
                                                bool vec_ok, float (&r)[8]) {
  const int tid = threadIdx.x;
  if (KCONTIG) {
    const int kq = k0 + ((tid >> 6) << 2);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = row0 + (tid & 63) + 64 * j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n_rows) {
        const float* q = p + (size_t)row * ld + kq;
        if (vec_ok && kq + 3 < k_end) {
          v = __ldg(reinterpret_cast<const float4*>(q));
        } else {
          if (kq < k_end) v.x = __ldg(q);
          if (kq + 1 < k_end) v.y = __ldg(q + 1);
          if (kq + 2 < k_end) v.z = __ldg(q + 2);
          if (kq + 3 < k_end) v.w = __ldg(q + 3);
        }
      }
      r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
    }
  } else {
    const int rq = row0 + ((tid & 31) << 2);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = k0 + (tid >> 5) + 8 * j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < k_end) {
        const float* q = p + (size_t)k * ld + rq;
        if (vec_ok && rq + 3 < n_rows) {
          v = __ldg(reinterpret_cast<const float4*>(q));
        } else {
          if (rq < n_rows) v.x = __ldg(q);
          if (rq + 1 < n_rows) v.y = __ldg(q + 1);
          if (rq + 2 < n_rows) v.z = __ldg(q + 2);
          if (rq + 3 < n_rows) v.w = __ldg(q + 3);
        }
      }
      r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
    }
  }
}

template <bool KCONTIG>
__device__ __forceinline__ void gemm_store_slice(float (*s)[GEMM_LDS], const float (&r)[8]) {
  const int tid = threadIdx.x;
  if (KCONTIG) {
    const int kq = (tid >> 6) << 2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = (tid & 63) + 64 * j;
#pragma unroll
      for (int i = 0; i < 4; ++i) s[kq + i][row] = r[4 * j + i];
    }
  } else {
    const int rq = (tid & 31) << 2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = (tid >> 5) + 8 * j;
      *reinterpret_cast<float4*>(&s[k][rq]) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
  }
}

template <bool A_KCONTIG, bool B_KCONTIG, int EPI>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[GEMM_BK][GEMM_LDS];
  __shared__ __align__(16) float Bs[GEMM_BK][GEMM_LDS];
  const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * GEMM_BN;
  int gM = g.M, gK = g.K;
  if (g.rows_dev) {
    if (g.rows_is_k) gK = apn_rt_count(g.rows_dev, g.K);
    else gM = apn_rt_count(g.rows_dev, g.M);
    if (m0 >= gM) return;                      // uniform over the block
  }
  int k_begin = 0, k_end = gK;
  if (EPI == GEMM_EPI_ATOMIC) {
    k_begin = blockIdx.z * g.k_chunk;
    k_end = min(gK, k_begin + g.k_chunk);
    if (k_begin >= k_end) return;
  }
  const bool a_vec = ((g.lda & 3) == 0) && ((((uintptr_t)g.A) & 15) == 0) && (A_KCONTIG ? true : true);
  const bool b_vec = ((g.ldb & 3) == 0) && ((((uintptr_t)g.B) & 15) == 0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  gemm_load_slice<A_KCONTIG>(g.A, g.lda, m0, gM, k_begin, k_end, a_vec, ra);
  gemm_load_slice<B_KCONTIG>(g.B, g.ldb, n0, g.N, k_begin, k_end, b_vec, rb);
  for (int k0 = k_begin; k0 < k_end; k0 += GEMM_BK) {
    __syncthreads();
    gemm_store_slice<A_KCONTIG>(As, ra);
    gemm_store_slice<B_KCONTIG>(Bs, rb);
    __syncthreads();
    if (k0 + GEMM_BK < k_end) {
      gemm_load_slice<A_KCONTIG>(g.A, g.lda, m0, gM, k0 + GEMM_BK, k_end, a_vec, ra);
      gemm_load_slice<B_KCONTIG>(g.B, g.ldb, n0, g.N, k0 + GEMM_BK, k_end, b_vec, rb);
    }
#pragma unroll
    for (int k = 0; k < GEMM_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= gM) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n >= g.N) continue;
      float v[4] = {acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]};
      float* c = g.C + (size_t)m * g.ldc + n;
      if (EPI == GEMM_EPI_ATOMIC) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < g.N) atomicAdd(c + j, v[j]);
      } else {
        if (EPI == GEMM_EPI_BIAS_ACT) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j < g.N) {
              float y = v[j] + (g.bias ? __ldg(g.bias + n + j) : 0.f);
              v[j] = (y < 0.f) ? y * g.slope : y;
            }
          }
        } else if (EPI == GEMM_EPI_ACCUM) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < g.N) v[j] += c[j];
        } else if (g.mask) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < g.N) v[j] = (__ldg(g.mask + (size_t)m * g.ldm + n + j) > 0.f) ? v[j] : v[j] * g.slope;
        }
        if (n + 3 < g.N && ((g.ldc & 3) == 0) && ((((uintptr_t)g.C) & 15) == 0)) {
          *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < g.N) c[j] = v[j];
        }
      }
    }
  }
}

// Y = act(X W^T + b):  X (M x K, ld lda), W (N x K torch layout, ld ldw), Y (M x N, ld ldc)
static inline int gemm_forward(cudaStream_t st, const float* X, int lda, const float* W, int ldw, const float* bias, float* Y,
                               int ldc, int M, int N, int K, float slope, const int32_t* rows_dev = nullptr) {
  if (M <= 0) return 0;
  GemmArgs g = {X, lda, W, ldw, Y, ldc, M, N, K, bias, slope, nullptr, 0, 0, rows_dev, 0};
  dim3 grid(apn_div_up(M, GEMM_BM), apn_div_up(N, GEMM_BN), 1);
  sgemm_kernel<true, true, GEMM_EPI_BIAS_ACT><<<grid, 256, 0, st>>>(g);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
// dX = (dY W) * act'(mask):  dY (M x K), W (K x N torch layout: out=K, in=N), dX (M x N)
static inline int gemm_dgrad(cudaStream_t st, const float* dY, int lda, const float* W, int ldw, float* dX, int ldc, int M, int N,
                             int K, const float* mask, int ldm, float slope, const int32_t* rows_dev = nullptr) {
  if (M <= 0) return 0;
  GemmArgs g = {dY, lda, W, ldw, dX, ldc, M, N, K, nullptr, slope, mask, ldm, 0, rows_dev, 0};
  dim3 grid(apn_div_up(M, GEMM_BM), apn_div_up(N, GEMM_BN), 1);
  sgemm_kernel<true, false, GEMM_EPI_MASK><<<grid, 256, 0, st>>>(g);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
// dX += dY W (no activation mask): the accumulating form, for gradients that land in a caller-owned buffer
static inline int gemm_dgrad_accum(cudaStream_t st, const float* dY, int lda, const float* W, int ldw, float* dX, int ldc, int M,
                                   int N, int K) {
  if (M <= 0) return 0;
  GemmArgs g = {dY, lda, W, ldw, dX, ldc, M, N, K, nullptr, 1.f, nullptr, 0, 0, nullptr, 0};
  dim3 grid(apn_div_up(M, GEMM_BM), apn_div_up(N, GEMM_BN), 1);
  sgemm_kernel<true, false, GEMM_EPI_ACCUM><<<grid, 256, 0, st>>>(g);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
// dW (n_out x n_in, ld ldw) += dY^T X : dY (rows x n_out, ld ldy), X (rows x n_in, ld ldx)
static inline int gemm_wgrad(cudaStream_t st, const float* dY, int ldy, const float* X, int ldx, float* dW, int ldw, int rows,
                             int n_out, int n_in, const int32_t* rows_dev = nullptr) {
  if (rows <= 0) return 0;
  const int tiles = apn_div_up(n_out, GEMM_BM) * apn_div_up(n_in, GEMM_BN);
  int splits = (2 * APN_SM_COUNT + tiles - 1) / tiles;
  int k_chunk = apn_div_up(rows, splits);
  k_chunk = ((k_chunk + GEMM_BK - 1) / GEMM_BK) * GEMM_BK;
  if (k_chunk < 64) k_chunk = 64;        // enough splits to fill the machine on the 8192-ray training batches (K ~ 1e4)
  splits = apn_div_up(rows, k_chunk);
  GemmArgs g = {dY, ldy, X, ldx, dW, ldw, n_out, n_in, rows, nullptr, 1.f, nullptr, 0, k_chunk, rows_dev, 1};
  dim3 grid(apn_div_up(n_out, GEMM_BM), apn_div_up(n_in, GEMM_BN), splits);
  sgemm_kernel<false, false, GEMM_EPI_ATOMIC><<<grid, 256, 0, st>>>(g);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
