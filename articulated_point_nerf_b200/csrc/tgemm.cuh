// fp32-class GEMM on the tensor cores (3xTF32, mma.sync m16n8k8) for the small dense products AROUND the tcgen05 decoder
// in the training step: the per-point table P = feat W0_feat^T (forward) and d_feat = dP W0_feat, dW0_feat = dP^T feat
// (backward), and the RGBNet head backward (d_fv, d_h, dW_v0, dW_feat).  These were CUDA-core SGEMMs (sgemm.cuh): on the
// 8192-ray batches every one of them is a handful of 128x128 tiles that runs at ~9 TFLOP/s and sits on the critical path
// of the step.  Same interface and epilogues as sgemm.cuh (GemmArgs):
//   C[M x N] (+)= sum_k A(m,k) B(k,n),   either operand with either index contiguous.
// Every fp32 operand is split x = hi + lo into two TF32 values (10-bit mantissas, cvt.rna) when its fragment is read from
// shared memory, and a_lo b_hi + a_hi b_lo + a_hi b_hi is accumulated in fp32: ~2^-21 relative error per product, the same
// class as an FFMA chain (tests/test_gpu_kernels.py compares with the fp32 oracle at the unchanged tolerances).
// 64 x 128 tile (twice as many CTAs as sgemm's 128 x 128 on the short sample dimension), 16-deep k slices, 8 warps as
// 2 (m) x 4 (n), each 32 x 32 = 2 x 4 MMA tiles; shared-memory rows padded to 72 / 136 floats so that the fragment reads
// (bank = 8 t + g) are conflict-free; register-staged prefetch of the next k slice.
#pragma once
#include "sgemm.cuh"

#define TG_BM 64
#define TG_BN 128
#define TG_BK 16
#define TG_LDA (TG_BM + 8)
#define TG_LDB (TG_BN + 8)

__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// one (ROWS x 16) operand slice -> registers (ROWS / 16 floats per thread) -> shared memory s[k][row]
//   KCONTIG: element (row, k) at p[row * ld + k]   thread: row = (tid & 63) + 64 j, k = (tid >> 6) * 4 .. + 3
//   else   : element (row, k) at p[k * ld + row]   thread: k = tid / (ROWS / 4) + (1024 / ROWS) j, rows (tid % (ROWS / 4)) * 4 .. + 3
template <bool KCONTIG, int ROWS>
__device__ __forceinline__ void tg_load(const float* __restrict__ p, int ld, int row0, int n_rows, int k0, int k_end, bool vec_ok,
                                        float (&r)[ROWS / 16]) {
  const int tid = threadIdx.x;
  constexpr int NV = ROWS / 64;             // float4 per thread
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KCONTIG) {
      const int row = row0 + (tid & 63) + 64 * j, kq = k0 + ((tid >> 6) << 2);
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
    } else {
      constexpr int TPR = ROWS / 4;         // threads per k row
      const int k = k0 + tid / TPR + (256 / TPR) * j, rq = row0 + (tid % TPR) * 4;
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
    }
    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
  }
}

template <bool KCONTIG, int ROWS, int LDS>
__device__ __forceinline__ void tg_store(float (*s)[LDS], const float (&r)[ROWS / 16]) {
  const int tid = threadIdx.x;
  constexpr int NV = ROWS / 64;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (KCONTIG) {
      const int row = (tid & 63) + 64 * j, kq = (tid >> 6) << 2;
#pragma unroll
      for (int i = 0; i < 4; ++i) s[kq + i][row] = r[4 * j + i];
    } else {
      constexpr int TPR = ROWS / 4;
      const int k = tid / TPR + (256 / TPR) * j, rq = (tid % TPR) * 4;
      *reinterpret_cast<float4*>(&s[k][rq]) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
  }
}

template <bool A_KCONTIG, bool B_KCONTIG, int EPI>
__global__ void __launch_bounds__(256) tgemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[TG_BK][TG_LDA];
  __shared__ __align__(16) float Bs[TG_BK][TG_LDB];
  const int m0 = blockIdx.x * TG_BM, n0 = blockIdx.y * TG_BN;
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
  const bool a_vec = ((g.lda & 3) == 0) && ((((uintptr_t)g.A) & 15) == 0);
  const bool b_vec = ((g.ldb & 3) == 0) && ((((uintptr_t)g.B) & 15) == 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  float ra[4], rb[8];
  tg_load<A_KCONTIG, TG_BM>(g.A, g.lda, m0, gM, k_begin, k_end, a_vec, ra);
  tg_load<B_KCONTIG, TG_BN>(g.B, g.ldb, n0, g.N, k_begin, k_end, b_vec, rb);
  for (int k0 = k_begin; k0 < k_end; k0 += TG_BK) {
    __syncthreads();
    tg_store<A_KCONTIG, TG_BM, TG_LDA>(As, ra);
    tg_store<B_KCONTIG, TG_BN, TG_LDB>(Bs, rb);
    __syncthreads();
    if (k0 + TG_BK < k_end) {
      tg_load<A_KCONTIG, TG_BM>(g.A, g.lda, m0, gM, k0 + TG_BK, k_end, a_vec, ra);
      tg_load<B_KCONTIG, TG_BN>(g.B, g.ldb, n0, g.N, k0 + TG_BK, k_end, b_vec, rb);
    }
#pragma unroll
    for (int kk = 0; kk < TG_BK; kk += 8) {
      uint32_t a_hi[2][4], a_lo[2][4], b_hi[4][2], b_lo[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = wm + mi * 16 + gq;
        tf32_split(As[kk + tq][r], a_hi[mi][0], a_lo[mi][0]);
        tf32_split(As[kk + tq][r + 8], a_hi[mi][1], a_lo[mi][1]);
        tf32_split(As[kk + tq + 4][r], a_hi[mi][2], a_lo[mi][2]);
        tf32_split(As[kk + tq + 4][r + 8], a_hi[mi][3], a_lo[mi][3]);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int c = wn + ni * 8 + gq;
        tf32_split(Bs[kk + tq][c], b_hi[ni][0], b_lo[ni][0]);
        tf32_split(Bs[kk + tq + 4][c], b_hi[ni][1], b_lo[ni][1]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          mma_tf32_16x8x8(acc[mi][ni], a_lo[mi], b_hi[ni]);      // small terms first
          mma_tf32_16x8x8(acc[mi][ni], a_hi[mi], b_lo[ni]);
          mma_tf32_16x8x8(acc[mi][ni], a_hi[mi], b_hi[ni]);
        }
    }
  }
  // epilogue: c0 c1 = C[g][2t, 2t+1], c2 c3 = C[g+8][2t, 2t+1]
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm + mi * 16 + gq + 8 * half;
      if (m >= gM) continue;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int n = n0 + wn + ni * 8 + 2 * tq;
        if (n >= g.N) continue;
        float v[2] = {acc[mi][ni][2 * half], acc[mi][ni][2 * half + 1]};
        float* c = g.C + (size_t)m * g.ldc + n;
        const bool two = n + 1 < g.N;
        if (EPI == GEMM_EPI_ATOMIC) {
          atomicAdd(c, v[0]);
          if (two) atomicAdd(c + 1, v[1]);
          continue;
        }
        if (EPI == GEMM_EPI_BIAS_ACT) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (n + j < g.N) {
              const float y = v[j] + (g.bias ? __ldg(g.bias + n + j) : 0.f);
              v[j] = (y < 0.f) ? y * g.slope : y;
            }
        } else if (EPI == GEMM_EPI_ACCUM) {
          v[0] += c[0];
          if (two) v[1] += c[1];
        } else if (g.mask) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (n + j < g.N) v[j] = (__ldg(g.mask + (size_t)m * g.ldm + n + j) > 0.f) ? v[j] : v[j] * g.slope;
        }
        if (two && ((g.ldc & 1) == 0) && ((((uintptr_t)g.C) & 7) == 0)) {
          *reinterpret_cast<float2*>(c) = make_float2(v[0], v[1]);
        } else {
          c[0] = v[0];
          if (two) c[1] = v[1];
        }
      }
    }
}

// the four call shapes of sgemm.cuh on the tensor cores (same argument meaning)
static inline int tgemm_forward(cudaStream_t st, const float* X, int lda, const float* W, int ldw, const float* bias, float* Y,
                                int ldc, int M, int N, int K, float slope, const int32_t* rows_dev = nullptr) {
  if (M <= 0) return 0;
  GemmArgs g = {X, lda, W, ldw, Y, ldc, M, N, K, bias, slope, nullptr, 0, 0, rows_dev, 0};
  dim3 grid(apn_div_up(M, TG_BM), apn_div_up(N, TG_BN), 1);
  tgemm_kernel<true, true, GEMM_EPI_BIAS_ACT><<<grid, 256, 0, st>>>(g);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
static inline int tgemm_dgrad(cudaStream_t st, const float* dY, int lda, const float* W, int ldw, float* dX, int ldc, int M, int N,
                              int K, const float* mask, int ldm, float slope, const int32_t* rows_dev = nullptr) {
  if (M <= 0) return 0;
  GemmArgs g = {dY, lda, W, ldw, dX, ldc, M, N, K, nullptr, slope, mask, ldm, 0, rows_dev, 0};
  dim3 grid(apn_div_up(M, TG_BM), apn_div_up(N, TG_BN), 1);
  tgemm_kernel<true, false, GEMM_EPI_MASK><<<grid, 256, 0, st>>>(g);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
static inline int tgemm_dgrad_accum(cudaStream_t st, const float* dY, int lda, const float* W, int ldw, float* dX, int ldc, int M,
                                    int N, int K) {
  if (M <= 0) return 0;
  GemmArgs g = {dY, lda, W, ldw, dX, ldc, M, N, K, nullptr, 1.f, nullptr, 0, 0, nullptr, 0};
  dim3 grid(apn_div_up(M, TG_BM), apn_div_up(N, TG_BN), 1);
  tgemm_kernel<true, false, GEMM_EPI_ACCUM><<<grid, 256, 0, st>>>(g);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
static inline int tgemm_wgrad(cudaStream_t st, const float* dY, int ldy, const float* X, int ldx, float* dW, int ldw, int rows,
                              int n_out, int n_in, const int32_t* rows_dev = nullptr) {
  if (rows <= 0) return 0;
  const int tiles = apn_div_up(n_out, TG_BM) * apn_div_up(n_in, TG_BN);
  int splits = (2 * APN_SM_COUNT + tiles - 1) / tiles;
  int k_chunk = apn_div_up(rows, splits);
  k_chunk = ((k_chunk + TG_BK - 1) / TG_BK) * TG_BK;
  if (k_chunk < 64) k_chunk = 64;
  splits = apn_div_up(rows, k_chunk);
  GemmArgs g = {dY, ldy, X, ldx, dW, ldw, n_out, n_in, rows, nullptr, 1.f, nullptr, 0, k_chunk, rows_dev, 1};
  dim3 grid(apn_div_up(n_out, TG_BM), apn_div_up(n_in, TG_BN), splits);
  tgemm_kernel<false, false, GEMM_EPI_ATOMIC><<<grid, 256, 0, st>>>(g);
  apn_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
