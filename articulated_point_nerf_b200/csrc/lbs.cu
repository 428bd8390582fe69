// K1 — fused skinning: softmax(raw/theta) -> merge rules -> blend of bone transforms -> point
// transform (+ global translation) -> closed-form inverse of the blended 3x3 frame -> bbox.
// Replaces lib/temporalpoints.py:401-414,424,569 and lib/pointwarper.py:241-266.
//
// Bandwidth-bound: algorithmic bytes fwd = N*(4J + 12 + 12 + 36) (+4J when merged weights are
// written); bwd = N*(8J + 12 + 12 + 36 + 36).  Warp-centric: a warp owns 32 points at a time, reads and writes
// the (N x J) weight rows coalesced with its lanes over the bones, and switches to lane = point (rows in a
// warp-private shared tile with an odd stride, bone matrices as warp broadcasts) for the per-point algebra.
// No block-wide barriers inside the loop; persistent grid (multiple of the SM count).
#include "common.cuh"

#define LBS_MAX_J 128

__global__ void lbs_init_bbox_kernel(float* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = __int_as_float(0x7f800000);       // +inf
  else if (threadIdx.x < 6) bbox[threadIdx.x] = __int_as_float(0xff800000);  // -inf
}

__device__ __forceinline__ void inverse3x3(const float a[9], float inv[9]) {
  const float c00 = a[4] * a[8] - a[5] * a[7];
  const float c01 = a[5] * a[6] - a[3] * a[8];
  const float c02 = a[3] * a[7] - a[4] * a[6];
  const float det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const float r = 1.0f / det;
  inv[0] = c00 * r;
  inv[1] = (a[2] * a[7] - a[1] * a[8]) * r;
  inv[2] = (a[1] * a[5] - a[2] * a[4]) * r;
  inv[3] = c01 * r;
  inv[4] = (a[0] * a[8] - a[2] * a[6]) * r;
  inv[5] = (a[2] * a[3] - a[0] * a[5]) * r;
  inv[6] = c02 * r;
  inv[7] = (a[1] * a[6] - a[0] * a[7]) * r;
  inv[8] = (a[0] * a[4] - a[1] * a[3]) * r;
}

// softmax over a shared-memory row (in place) + in-place merge; returns nothing.
__device__ __forceinline__ void softmax_row(float* row, int J, float theta) {
  float mx = -INFINITY;
  for (int j = 0; j < J; ++j) {
    const float z = row[j] / theta;
    row[j] = z;
    mx = fmaxf(mx, z);
  }
  float sum = 0.f;
  for (int j = 0; j < J; ++j) {
    const float e = expf(row[j] - mx);
    row[j] = e;
    sum += e;
  }
  const float inv = 1.0f / sum;
  for (int j = 0; j < J; ++j) row[j] *= inv;
}

// softmax(raw / theta) of one shared-memory row by ONE thread, in place (three passes over the row).
// x / theta is evaluated as q = x * (1/theta) followed by one residual correction, q + (x - q theta) * (1/theta):
// the correctly rounded quotient for all but a vanishing fraction of operands, at 3 instructions instead of the
// ~10 of the IEEE division sequence.
__device__ __forceinline__ void softmax_row(float* row, int J, float theta, float inv_theta) {
  float mx = -INFINITY;
  for (int j = 0; j < J; ++j) {
    const float x = row[j];
    const float q = x * inv_theta;
    const float z = fmaf(fmaf(-q, theta, x), inv_theta, q);
    row[j] = z;
    mx = fmaxf(mx, z);
  }
  float sum = 0.f;
  for (int j = 0; j < J; ++j) {
    const float e = expf(row[j] - mx);
    row[j] = e;
    sum += e;
  }
  const float inv = 1.0f / sum;
  for (int j = 0; j < J; ++j) row[j] *= inv;
}

// Forward: warp-centric, no block-wide barriers.  A warp owns 32 consecutive points per step:
//   phase 1  lanes run over the bones of one point at a time: coalesced row reads straight from global memory (four rows
//            in flight, one batch ahead of their use) into the warp's shared-memory tile (odd row stride);
//   phase 2  lane = point: softmax in three passes over its row (max; exp + sum; normalise), the last one fused with
//            the blend of the bone 3x4s (shared-memory broadcasts, packed fp32x2 FMAs) unless merge rules apply;
//            transform, closed-form inverse, bbox;
//   phase 3  the (merged) weights go back to global memory row by row, coalesced.
// The load/store pipe, not HBM, bounds this kernel (ncu: lsu 65 %, issue 65 % active, DRAM 33 % before the passes were
// fused), so the passes keep shared-memory traffic to one load (+ one store where the value is needed later) per element.
#define LBS_WARPS 4
#define LBS_MAX_K (LBS_MAX_J / 32)

template <int K>          // K = ceil(J / 32) registers per lane and row
__global__ void __launch_bounds__(32 * LBS_WARPS)
lbs_fwd_kernel(const float* __restrict__ raw_w, const float* __restrict__ theta_weight, float eps,
               const int* __restrict__ rules, const float* __restrict__ bone_T, const float* __restrict__ xyz,
               const float* __restrict__ global_t, int N, int J, float* __restrict__ xyz_out,
               float* __restrict__ ginv_out, float* __restrict__ w_out, float* __restrict__ g_out,
               float* __restrict__ bbox) {
  extern __shared__ float smem[];
  const int JP = J | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sT = smem;                                  // J*12
  int* sR = (int*)(sT + J * 12);                     // J
  float* sW = (float*)(sR + J) + warp * 32 * JP;     // this warp's 32 x JP tile
  for (int i = threadIdx.x; i < J * 12; i += blockDim.x) {
    const int j = i / 12, c = i - j * 12;
    sT[i] = bone_T[j * 16 + c];        // rows 0..2 of the 4x4
  }
  for (int j = threadIdx.x; j < J; j += blockDim.x) sR[j] = rules ? rules[j] : j;
  __syncthreads();
  const float theta = theta_weight ? fmaxf(eps, theta_weight[0]) : 1.f;
  const float inv_theta = 1.0f / theta;
  const float gx = global_t ? global_t[0] : 0.f, gy = global_t ? global_t[1] : 0.f, gz = global_t ? global_t[2] : 0.f;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mxv[3] = {-INFINITY, -INFINITY, -INFINITY};
  const int n_chunks = (N + 31) / 32;
  const int n_warps = gridDim.x * LBS_WARPS;
  // rows are fetched four at a time, one batch ahead of their use (across chunk boundaries too)
  auto load_rows = [&](int chunk, int p0, float (&v)[4][K]) {
    const int base = chunk * 32;
    const int last = min(32, N - base) - 1;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t row = (size_t)(base + min(p0 + u, last)) * J;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        v[u][k] = j < J ? __ldg(raw_w + row + j) : -INFINITY;
      }
    }
  };
  float vn[4][K];
  int chunk = blockIdx.x * LBS_WARPS + warp;
  if (chunk < n_chunks) load_rows(chunk, 0, vn);
  for (; chunk < n_chunks; chunk += n_warps) {
    const int base = chunk * 32;
    const int n_valid = min(32, N - base);
    // ---------------------------------------------------------------- phase 1: rows -> warp tile, lanes over bones
    for (int p0 = 0; p0 < n_valid; p0 += 4) {
      float v[4][K];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) v[u][k] = vn[u][k];
      if (p0 + 4 < n_valid) load_rows(chunk, p0 + 4, vn);
      else if (chunk + n_warps < n_chunks) load_rows(chunk + n_warps, 0, vn);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (p0 + u >= n_valid) break;
        float* srow = sW + (p0 + u) * JP;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = lane + 32 * k;
          if (j < J) srow[j] = v[u][k];
        }
      }
    }
    __syncwarp();
    // ---------------------------------------------------------------- phase 2: lane = point
    const bool active = lane < n_valid;   // every lane runs the algebra (keeps the bone loop warp-uniform); only stores are guarded
    {
      float* row = sW + lane * JP;
      float2 G2[6];                       // G[0..11] as six fp32x2 pairs: one FFMA2 updates two entries, same roundings
#pragma unroll
      for (int c = 0; c < 6; ++c) G2[c] = make_float2(0.f, 0.f);
#define LBS_BLEND(w, j)                                                                                               \
  do {                                                                                                                \
    const float4* T = reinterpret_cast<const float4*>(sT + (j) * 12);                                                 \
    const float4 t0 = T[0], t1 = T[1], t2 = T[2];                                                                     \
    const float2 ww = make_float2(w, w);                                                                              \
    G2[0] = __ffma2_rn(ww, make_float2(t0.x, t0.y), G2[0]); G2[1] = __ffma2_rn(ww, make_float2(t0.z, t0.w), G2[1]);   \
    G2[2] = __ffma2_rn(ww, make_float2(t1.x, t1.y), G2[2]); G2[3] = __ffma2_rn(ww, make_float2(t1.z, t1.w), G2[3]);   \
    G2[4] = __ffma2_rn(ww, make_float2(t2.x, t2.y), G2[4]); G2[5] = __ffma2_rn(ww, make_float2(t2.z, t2.w), G2[5]);   \
  } while (0)
      if (theta_weight) {
        // x / theta as q = x * (1/theta) plus one residual correction (see softmax_row)
        float mx = -INFINITY;
#pragma unroll 4
        for (int j = 0; j < J; ++j) {
          const float x = row[j];
          const float q = x * inv_theta;
          mx = fmaxf(mx, fmaf(fmaf(-q, theta, x), inv_theta, q));
        }
        float sum = 0.f;
#pragma unroll 4
        for (int j = 0; j < J; ++j) {
          const float x = row[j];
          const float q = x * inv_theta;
          const float e = expf(fmaf(fmaf(-q, theta, x), inv_theta, q) - mx);
          row[j] = e;
          sum += e;
        }
        const float inv = 1.0f / sum;
        if (!rules) {
          if (w_out) {
#pragma unroll 4
            for (int j = 0; j < J; ++j) {
              const float w = row[j] * inv;
              row[j] = w;
              LBS_BLEND(w, j);
            }
          } else {
#pragma unroll 4
            for (int j = 0; j < J; ++j) {
              const float w = row[j] * inv;
              LBS_BLEND(w, j);
            }
          }
        } else {
#pragma unroll 4
          for (int j = 0; j < J; ++j) row[j] *= inv;
        }
      }
      if (rules) {
        for (int j = 0; j < J; ++j) {
          const int t = sR[j];
          if (t != j) {
            row[t] += row[j];
            row[j] = 0.f;
          }
        }
      }
      if (rules || !theta_weight) {
#pragma unroll 4
        for (int j = 0; j < J; ++j) {
          const float w = row[j];
          LBS_BLEND(w, j);
        }
      }
#undef LBS_BLEND
      const float G[12] = {G2[0].x, G2[0].y, G2[1].x, G2[1].y, G2[2].x, G2[2].y, G2[3].x, G2[3].y, G2[4].x, G2[4].y, G2[5].x, G2[5].y};
      if (active) {
      const size_t n = (size_t)base + lane;
      const float x = xyz[3 * n], y = xyz[3 * n + 1], z = xyz[3 * n + 2];
      const float ox = G[0] * x + G[1] * y + G[2] * z + G[3] + gx;
      const float oy = G[4] * x + G[5] * y + G[6] * z + G[7] + gy;
      const float oz = G[8] * x + G[9] * y + G[10] * z + G[11] + gz;
      xyz_out[3 * n] = ox;
      xyz_out[3 * n + 1] = oy;
      xyz_out[3 * n + 2] = oz;
      mn[0] = fminf(mn[0], ox); mn[1] = fminf(mn[1], oy); mn[2] = fminf(mn[2], oz);
      mxv[0] = fmaxf(mxv[0], ox); mxv[1] = fmaxf(mxv[1], oy); mxv[2] = fmaxf(mxv[2], oz);
      const float A[9] = {G[0], G[1], G[2], G[4], G[5], G[6], G[8], G[9], G[10]};
      float inv[9];
      inverse3x3(A, inv);
#pragma unroll
      for (int c = 0; c < 9; ++c) ginv_out[9 * n + c] = inv[c];
      if (g_out) {
        float4* go = reinterpret_cast<float4*>(g_out + 16 * n);
        go[0] = make_float4(G[0], G[1], G[2], G[3]);
        go[1] = make_float4(G[4], G[5], G[6], G[7]);
        go[2] = make_float4(G[8], G[9], G[10], G[11]);
        go[3] = make_float4(0.f, 0.f, 0.f, 1.f);
      }
      }
    }
    __syncwarp();
    if (w_out) {                                       // (merged) weights: the tile, row by row, coalesced
      for (int p = 0; p < n_valid; ++p) {
        const float* srow = sW + p * JP;
        float* grow = w_out + (size_t)(base + p) * J;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = lane + 32 * k;
          if (j < J) grow[j] = srow[j];
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = warp_min(mn[c]), b = warp_max(mxv[c]);
    if (lane == 0) {
      if (a < INFINITY) atomic_min_float(bbox + c, a);
      if (b > -INFINITY) atomic_max_float(bbox + 3 + c, b);
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward
// partial layout per block: [J*12 dT | 1 dtheta | 3 dglobal_t]
// Warp-centric like the forward; per 32-point chunk:
//   phase 1  lanes over bones: raw row -> softmax weights w (warp tile sW), incoming d_w row (warp tile sDM);
//   phase 2  lane = point: dA = -B^T dB B^T, dG (registers + warp tile sDG), merged weights (sM, only with merge
//            rules), dm_j = d_w_j + <dG, T_j> (sDM, in place);
//   phase 3  lanes over bones: softmax backward -> d_raw row (coalesced store) and the theta gradient;
//            dT_j += sum_p m[p][j] dG[p][:] in registers (12 K accumulators per lane, kept across chunks).
// ---------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(32 * LBS_WARPS)
lbs_bwd_kernel(const float* __restrict__ raw_w, const float* __restrict__ theta_weight, float eps,
               const int* __restrict__ rules, const float* __restrict__ bone_T, const float* __restrict__ xyz, int N,
               int J, const float* __restrict__ ginv, const float* __restrict__ d_xyz,
               const float* __restrict__ d_ginv, const float* __restrict__ d_w, const float* __restrict__ d_g,
               float* __restrict__ d_raw, float* __restrict__ partial) {
  extern __shared__ float smem[];
  const int JP = J | 1;
  const int n_out = J * 12;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // every float4-read array starts on a 16-byte boundary: sT, sAcc, then per warp [sDG | sW | sM | sDM], sR last
  float* sT = smem;                                        // J*12
  float* sAcc = sT + n_out;                                // J*12 + 4 block accumulators
  const int n_tiles_w = rules ? 3 : 2;                     // the merged-weight tile exists only with merge rules
  const int wstride = n_tiles_w * 32 * JP + 32 * 12;
  float* wbase = sAcc + n_out + 4 + warp * wstride;
  float* sDG = wbase;                                      // dG per point                 32*12
  float* sW = sDG + 32 * 12;                               // softmax weights              32*JP
  float* sDM = sW + 32 * JP;                               // d_w in, dm out               32*JP
  float* sM = sDM + 32 * JP;                               // merged weights (rules only)  32*JP
  int* sR = (int*)(sAcc + n_out + 4 + LBS_WARPS * wstride);                   // J
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    const int j = i / 12, c = i - j * 12;
    sT[i] = bone_T[j * 16 + c];
  }
  for (int i = threadIdx.x; i < n_out + 4; i += blockDim.x) sAcc[i] = 0.f;
  for (int j = threadIdx.x; j < J; j += blockDim.x) sR[j] = rules ? rules[j] : j;
  __syncthreads();
  const float theta = theta_weight ? fmaxf(eps, theta_weight[0]) : 1.f;
  const float inv_theta = 1.0f / theta;
  double acc_theta = 0.0;                                  // heavy cancellation over N*J terms: accumulated in fp64
  float acc_g[3] = {0.f, 0.f, 0.f};
  float accT[K][12];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int c = 0; c < 12; ++c) accT[k][c] = 0.f;
  const float* sMm = rules ? sM : sW;                      // weights that multiplied T_j in the forward
  const int n_chunks = (N + 31) / 32;
  const int n_warps = gridDim.x * LBS_WARPS;
  auto load_rows = [&](int chunk, int p0, float (&v)[4][K], float (&dwv)[4][K]) {
    const int base = chunk * 32;
    const int last = min(32, N - base) - 1;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t row = (size_t)(base + min(p0 + u, last)) * J;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        v[u][k] = j < J ? __ldg(raw_w + row + j) : -INFINITY;
        dwv[u][k] = (d_w && j < J) ? __ldg(d_w + row + j) : 0.f;
      }
    }
  };
  float vn[4][K], dwn[4][K];
  int chunk = blockIdx.x * LBS_WARPS + warp;
  if (chunk < n_chunks) load_rows(chunk, 0, vn, dwn);
  for (; chunk < n_chunks; chunk += n_warps) {
    const int base = chunk * 32;
    const int n_valid = min(32, N - base);
    // ---------------------------------------------------------------- phase 1: raw and d_w rows -> warp tiles
    for (int p0 = 0; p0 < n_valid; p0 += 4) {
      float v[4][K], dwv[4][K];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) {
          v[u][k] = vn[u][k];
          dwv[u][k] = dwn[u][k];
        }
      if (p0 + 4 < n_valid) load_rows(chunk, p0 + 4, vn, dwn);
      else if (chunk + n_warps < n_chunks) load_rows(chunk + n_warps, 0, vn, dwn);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (p0 + u >= n_valid) break;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = lane + 32 * k;
          if (j < J) {
            sW[(p0 + u) * JP + j] = v[u][k];
            sDM[(p0 + u) * JP + j] = dwv[u][k];
          }
        }
      }
    }
    __syncwarp();
    // ---------------------------------------------------------------- phase 2: lane = point
    float* dg = sDG + lane * 12;
    if (lane < n_valid) {
      if (theta_weight) softmax_row(sW + lane * JP, J, theta, inv_theta);
      if (rules) {
        float* mrow = sM + lane * JP;
        const float* row = sW + lane * JP;
        for (int j = 0; j < J; ++j) mrow[j] = row[j];
        for (int j = 0; j < J; ++j) {
          const int t = sR[j];
          if (t != j) {
            mrow[t] += mrow[j];
            mrow[j] = 0.f;
          }
        }
      }
      const size_t n = (size_t)base + lane;
      float B[9], dB[9], dA[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        B[c] = ginv[9 * n + c];
        dB[c] = d_ginv ? d_ginv[9 * n + c] : 0.f;
      }
      // dA = -B^T dB B^T
      float t1[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) t1[r * 3 + c] = B[0 * 3 + r] * dB[0 * 3 + c] + B[1 * 3 + r] * dB[1 * 3 + c] + B[2 * 3 + r] * dB[2 * 3 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) dA[r * 3 + c] = -(t1[r * 3 + 0] * B[c * 3 + 0] + t1[r * 3 + 1] * B[c * 3 + 1] + t1[r * 3 + 2] * B[c * 3 + 2]);
      const float x = xyz[3 * n], y = xyz[3 * n + 1], z = xyz[3 * n + 2];
      const float gx = d_xyz ? d_xyz[3 * n] : 0.f, gy = d_xyz ? d_xyz[3 * n + 1] : 0.f, gz = d_xyz ? d_xyz[3 * n + 2] : 0.f;
      acc_g[0] += gx; acc_g[1] += gy; acc_g[2] += gz;
      // dG rows: [dA row | db]
      float g12[12] = {dA[0] + gx * x, dA[1] + gx * y, dA[2] + gx * z, gx, dA[3] + gy * x, dA[4] + gy * y, dA[5] + gy * z, gy,
                       dA[6] + gz * x, dA[7] + gz * y, dA[8] + gz * z, gz};
      if (d_g) {
#pragma unroll
        for (int c = 0; c < 12; ++c) g12[c] += d_g[16 * n + c];
      }
#pragma unroll
      for (int c = 0; c < 12; ++c) dg[c] = g12[c];
      // dm_j = d_w_j + <dG, T_j>
      float* dmrow = sDM + lane * JP;
      for (int j = 0; j < J; ++j) {
        const float4* T = reinterpret_cast<const float4*>(sT + j * 12);
        const float4 t0 = T[0], t1v = T[1], t2 = T[2];
        float sacc = dmrow[j];
        sacc = fmaf(g12[0], t0.x, sacc); sacc = fmaf(g12[1], t0.y, sacc); sacc = fmaf(g12[2], t0.z, sacc); sacc = fmaf(g12[3], t0.w, sacc);
        sacc = fmaf(g12[4], t1v.x, sacc); sacc = fmaf(g12[5], t1v.y, sacc); sacc = fmaf(g12[6], t1v.z, sacc); sacc = fmaf(g12[7], t1v.w, sacc);
        sacc = fmaf(g12[8], t2.x, sacc); sacc = fmaf(g12[9], t2.y, sacc); sacc = fmaf(g12[10], t2.z, sacc); sacc = fmaf(g12[11], t2.w, sacc);
        dmrow[j] = sacc;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 12; ++c) dg[c] = 0.f;
    }
    __syncwarp();
    // ---------------------------------------------------------------- phase 3: lanes over bones
    for (int p = 0; p < n_valid; ++p) {
      const float* wrow = sW + p * JP;
      const float* dmrow = sDM + p * JP;
      const size_t row = (size_t)(base + p) * J;
      float wv[K], dmr[K];
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        wv[k] = j < J ? wrow[j] : 0.f;
        dmr[k] = j < J ? dmrow[sR[j]] : 0.f;
        dot = fmaf(wv[k], dmr[k], dot);
      }
      if (theta_weight) {
        dot = warp_sum(dot);
        float th = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = lane + 32 * k;
          if (j < J) {
            const float dz = wv[k] * (dmr[k] - dot);
            th = fmaf(-dz, __ldg(raw_w + row + j), th);
            d_raw[row + j] = dz * inv_theta;
          }
        }
        acc_theta += (double)th;
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = lane + 32 * k;
          if (j < J) d_raw[row + j] = dmr[k];
        }
      }
      // dT_j += m[p][j] * dG[p][:]
      const float4* gp = reinterpret_cast<const float4*>(sDG + p * 12);
      const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2];
      const float* mrow = sMm + p * JP;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        const float m = j < J ? mrow[j] : 0.f;
        accT[k][0] = fmaf(m, g0.x, accT[k][0]); accT[k][1] = fmaf(m, g0.y, accT[k][1]); accT[k][2] = fmaf(m, g0.z, accT[k][2]);
        accT[k][3] = fmaf(m, g0.w, accT[k][3]); accT[k][4] = fmaf(m, g1.x, accT[k][4]); accT[k][5] = fmaf(m, g1.y, accT[k][5]);
        accT[k][6] = fmaf(m, g1.z, accT[k][6]); accT[k][7] = fmaf(m, g1.w, accT[k][7]); accT[k][8] = fmaf(m, g2.x, accT[k][8]);
        accT[k][9] = fmaf(m, g2.y, accT[k][9]); accT[k][10] = fmaf(m, g2.z, accT[k][10]); accT[k][11] = fmaf(m, g2.w, accT[k][11]);
      }
    }
    __syncwarp();
  }
  // block reduction (4 warps, shared-memory atomics: the order varies, the values are 4 partial sums per output)
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int j = lane + 32 * k;
    if (j < J) {
#pragma unroll
      for (int c = 0; c < 12; ++c) atomicAdd(&sAcc[j * 12 + c], accT[k][c]);
    }
  }
  {
    double t = acc_theta;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const float th = (float)(t / ((double)theta * (double)theta));
    const float a0 = warp_sum(acc_g[0]), a1 = warp_sum(acc_g[1]), a2 = warp_sum(acc_g[2]);
    if (lane == 0) {
      atomicAdd(&sAcc[n_out], th);
      atomicAdd(&sAcc[n_out + 1], a0);
      atomicAdd(&sAcc[n_out + 2], a1);
      atomicAdd(&sAcc[n_out + 3], a2);
    }
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * (n_out + 4);
  for (int o = threadIdx.x; o < n_out + 4; o += blockDim.x) out[o] = sAcc[o];
}

// ---------------------------------------------------------------------------------------
// backward on the tensor cores (no merge rules, J <= 80): the two contractions of the backward are small GEMMs,
//   dm^T (J x 32) = T (J x 12) * dG^T (12 x 32 points)        "how much does the loss want bone j's weight at point p"
//   dT   (J x 12) += w^T (J x 32 points) * dG (32 x 12)       accumulated over all points
// and run as mma.sync m16n8k8 TF32 with the 3xTF32 split (x = hi + lo; lo*hi + hi*lo + hi*hi, fp32 accumulate:
// fp32-class results).  A warp owns 32 points per step and holds the (bones x points) tile of raw / softmax weights
// directly in the MMA accumulator layout (bone rows g, g+8 of each 16-row tile; points 2q, 2q+1 of each 8-point tile), so
//   * the raw rows are read from and d_raw written to global memory in that layout (8 consecutive bones per quad: whole
//     32-byte sectors) — no shared-memory tile, no transposition;
//   * the softmax over bones is register arithmetic + three xor-shuffles over the lanes that share a point;
//   * the same registers are the A fragments of the second GEMM (its contraction index — the point — may be permuted
//     freely as long as dG's fragment uses the same permutation: k-slot q <-> point 2q, slot q+4 <-> point 2q+1).
// Only dG (32 x 12 per warp) goes through shared memory, to change from lane = point (where it is computed from d_xyz,
// d_ginv and the inverse frame) to fragment layout.  ~70 warp instructions per point instead of ~700.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_hi(x);
  lo = tf32_hi(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// max / sum over the 8 lanes that share q = lane & 3 (all bones of one point)
__device__ __forceinline__ float quad_col_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
}
__device__ __forceinline__ float quad_col_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v + __shfl_xor_sync(0xffffffffu, v, 16);
}

#define LBS_MMA_WARPS 4
#define LBS_DG_LD 13                     // row stride of the per-warp dG tile (odd: conflict-free lane = point stores)

template <int MT>                        // MT = ceil(J / 16) bone tiles, J <= 16 * MT
__global__ void __launch_bounds__(32 * LBS_MMA_WARPS)
lbs_bwd_mma_kernel(const float* __restrict__ raw_w, const float* __restrict__ theta_weight, float eps,
                   const float* __restrict__ bone_T, const float* __restrict__ xyz, int N, int J,
                   const float* __restrict__ ginv, const float* __restrict__ d_xyz, const float* __restrict__ d_ginv,
                   const float* __restrict__ d_w, float* __restrict__ d_raw, float* __restrict__ partial) {
  extern __shared__ float smem[];
  const int n_out = J * 12;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform: no divergence scaffolding around the MMAs
  const int g = lane >> 2, q = lane & 3;
  // [T A-fragments hi | lo : MT*2*4*32 each] [block accumulators n_out + 4] [per warp dG 32 x LBS_DG_LD]
  uint32_t* sTA_hi = reinterpret_cast<uint32_t*>(smem);
  uint32_t* sTA_lo = sTA_hi + MT * 2 * 4 * 32;
  float* sAcc = reinterpret_cast<float*>(sTA_lo + MT * 2 * 4 * 32);
  float* sDG = sAcc + n_out + 4 + warp * 32 * LBS_DG_LD;
  for (int i = threadIdx.x; i < MT * 2 * 4 * 32; i += blockDim.x) {     // [mt][kt][lane][reg]: one 16-byte load per fragment
    const int reg = i & 3, l = (i >> 2) & 31, kt = (i >> 7) & 1, mt = i >> 8;
    const int bone = 16 * mt + (l >> 2) + 8 * (reg & 1), c = 8 * kt + (l & 3) + 4 * (reg >> 1);
    const float v = (bone < J && c < 12) ? bone_T[bone * 16 + c] : 0.f;
    uint32_t hi, lo;
    tf32_split(v, hi, lo);
    sTA_hi[i] = hi;
    sTA_lo[i] = lo;
  }
  for (int i = threadIdx.x; i < n_out + 4; i += blockDim.x) sAcc[i] = 0.f;
  __syncthreads();
  const float theta = fmaxf(eps, theta_weight[0]);
  const float inv_theta = 1.0f / theta;
  double acc_theta = 0.0;
  float acc_g[3] = {0.f, 0.f, 0.f};
  float accT[MT][2][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
      for (int r = 0; r < 4; ++r) accT[mt][ct][r] = 0.f;
  const int n_chunks = (N + 31) / 32;
  const int n_warps = gridDim.x * LBS_MMA_WARPS;
  // raw values of one 8-point tile in accumulator layout: element r of bone tile mt = bone 16 mt + g + 8 (r >> 1),
  // point base + 8 nt + 2 q + (r & 1); anything beyond J or N reads as 0 (bones beyond J are masked to -inf in the softmax,
  // points beyond N have dG = 0: every product they enter is 0)
  // only the last bone tile can reach beyond J; a point is valid iff its index is below N
#define LBS_BONE_OK(mt, r) ((mt) < MT - 1 || 16 * (mt) + g + 8 * ((r) >> 1) < J)
  auto load_raw = [&](int base, int nt, float (&v)[MT][4]) {
    const int p0 = base + 8 * nt + 2 * q;
    const bool ok0 = p0 < N, ok1 = p0 + 1 < N;
    const float* rp = raw_w + (size_t)p0 * J + g;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const bool bone_ok = LBS_BONE_OK(mt, r);
        v[mt][r] = (bone_ok && ((r & 1) ? ok1 : ok0)) ? __ldg(rp + (r & 1) * J + 16 * mt + 8 * (r >> 1)) : 0.f;
      }
  };
  for (int chunk = blockIdx.x * LBS_MMA_WARPS + warp; chunk < n_chunks; chunk += n_warps) {
    const int base = chunk * 32;
    float nxt[MT][4];
    load_raw(base, 0, nxt);               // issued first: longest latency
    // ------------------------------------------------ lane = point: dG from d_xyz, d_ginv and the inverse frame
    {
      const int n = base + lane;
      float g12[12];
#pragma unroll
      for (int c = 0; c < 12; ++c) g12[c] = 0.f;
      if (n < N) {
        float B[9], dB[9], dA[9], t1[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          B[c] = ginv[9 * (size_t)n + c];
          dB[c] = d_ginv ? d_ginv[9 * (size_t)n + c] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) t1[r * 3 + c] = B[0 * 3 + r] * dB[0 * 3 + c] + B[1 * 3 + r] * dB[1 * 3 + c] + B[2 * 3 + r] * dB[2 * 3 + c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) dA[r * 3 + c] = -(t1[r * 3 + 0] * B[c * 3 + 0] + t1[r * 3 + 1] * B[c * 3 + 1] + t1[r * 3 + 2] * B[c * 3 + 2]);
        const float x = xyz[3 * (size_t)n], y = xyz[3 * (size_t)n + 1], z = xyz[3 * (size_t)n + 2];
        const float gx = d_xyz ? d_xyz[3 * (size_t)n] : 0.f, gy = d_xyz ? d_xyz[3 * (size_t)n + 1] : 0.f,
                    gz = d_xyz ? d_xyz[3 * (size_t)n + 2] : 0.f;
        acc_g[0] += gx; acc_g[1] += gy; acc_g[2] += gz;
        g12[0] = dA[0] + gx * x; g12[1] = dA[1] + gx * y; g12[2] = dA[2] + gx * z; g12[3] = gx;
        g12[4] = dA[3] + gy * x; g12[5] = dA[4] + gy * y; g12[6] = dA[5] + gy * z; g12[7] = gy;
        g12[8] = dA[6] + gz * x; g12[9] = dA[7] + gz * y; g12[10] = dA[8] + gz * z; g12[11] = gz;
      }
      __syncwarp();                       // the previous step's fragment reads are done
#pragma unroll
      for (int c = 0; c < 12; ++c) sDG[lane * LBS_DG_LD + c] = g12[c];
      __syncwarp();
    }
    float th4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int nt = 0; nt < 4; ++nt) {      // 8 points at a time: everything below lives in registers
      float rawv[MT][4], w[MT][4];
      const int p0 = base + 8 * nt + 2 * q;
      const bool ok0 = p0 < N, ok1 = p0 + 1 < N;
      const size_t at0 = (size_t)p0 * J + g;           // element (mt, r) of this lane: at0 + (r & 1) J + 16 mt + 8 (r >> 1)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int r = 0; r < 4; ++r) rawv[mt][r] = nxt[mt][r];
      if (nt < 3) load_raw(base, nt + 1, nxt);
      // ---- softmax over bones for this lane's two points (h = 0, 1), shuffles over the lanes that share q
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float mxa = -INFINITY, mxb = -INFINITY;            // two partial chains each (rows g and g + 8)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const float xa = LBS_BONE_OK(mt, h) ? rawv[mt][h] * inv_theta : -INFINITY;
          const float xb = LBS_BONE_OK(mt, h + 2) ? rawv[mt][h + 2] * inv_theta : -INFINITY;
          w[mt][h] = xa;
          w[mt][h + 2] = xb;
          mxa = fmaxf(mxa, xa);
          mxb = fmaxf(mxb, xb);
        }
        const float mx = quad_col_max(fmaxf(mxa, mxb));
        float suma = 0.f, sumb = 0.f;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const float ea = __expf(w[mt][h] - mx), eb = __expf(w[mt][h + 2] - mx);   // ex2.approx: ~2 ulp
          w[mt][h] = ea;
          w[mt][h + 2] = eb;
          suma += ea;
          sumb += eb;
        }
        const float inv = 1.0f / quad_col_sum(suma + sumb);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) w[mt][h + 2 * rr] *= inv;
      }
      // ---- dm^T = T dG^T for these 8 points:  B fragment b0 = dG[8 nt + g][8 kt + q], b1: column + 4
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int kt = 0; kt < 2; ++kt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = 8 * kt + q + 4 * h;
          const float v = c < 12 ? sDG[(8 * nt + g) * LBS_DG_LD + c] : 0.f;
          tf32_split(v, bh[kt][h], bl[kt][h]);
        }
      float acc[MT][4];
      float dot0 = 0.f, dot1 = 0.f, dot2 = 0.f, dot3 = 0.f;    // per accumulator element: four independent chains
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[mt][r] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          const uint4 ah = *reinterpret_cast<const uint4*>(sTA_hi + ((mt * 2 + kt) * 32 + lane) * 4);
          const uint4 al = *reinterpret_cast<const uint4*>(sTA_lo + ((mt * 2 + kt) * 32 + lane) * 4);
          mma_tf32(acc[mt], al.x, al.y, al.z, al.w, bh[kt][0], bh[kt][1]);
          mma_tf32(acc[mt], ah.x, ah.y, ah.z, ah.w, bl[kt][0], bl[kt][1]);
          mma_tf32(acc[mt], ah.x, ah.y, ah.z, ah.w, bh[kt][0], bh[kt][1]);
        }
        if (d_w) {                        // gradient arriving through the weights output (weight TV / sparsity losses)
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (LBS_BONE_OK(mt, r) && ((r & 1) ? ok1 : ok0)) acc[mt][r] += __ldg(d_w + at0 + (r & 1) * J + 16 * mt + 8 * (r >> 1));
        }
        dot0 = fmaf(w[mt][0], acc[mt][0], dot0); dot2 = fmaf(w[mt][2], acc[mt][2], dot2);
        dot1 = fmaf(w[mt][1], acc[mt][1], dot1); dot3 = fmaf(w[mt][3], acc[mt][3], dot3);
      }
      dot0 = quad_col_sum(dot0 + dot2);
      dot1 = quad_col_sum(dot1 + dot3);
      // ---- dz = w (dm - dot) -> d_raw, theta gradient
      float* dp = d_raw + at0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {                                 // branch-free: w = 0 beyond J, dm = dot = 0 beyond N, so dz = 0 there
          const float dz = w[mt][r] * (acc[mt][r] - ((r & 1) ? dot1 : dot0));
          th4[r] = fmaf(-dz, rawv[mt][r], th4[r]);
          if (LBS_BONE_OK(mt, r) && ((r & 1) ? ok1 : ok0)) dp[(r & 1) * J + 16 * mt + 8 * (r >> 1)] = dz * inv_theta;
        }
      // ---- dT += w^T dG:  A = (w r0, r2, r1, r3), B: b0 = dG[8 nt + 2 q][8 ct + g], b1: next point
      uint32_t gh[2][2], gl[2][2];
#pragma unroll
      for (int ct = 0; ct < 2; ++ct)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = 8 * ct + g;
          const float v = c < 12 ? sDG[(8 * nt + 2 * q + h) * LBS_DG_LD + c] : 0.f;
          tf32_split(v, gh[ct][h], gl[ct][h]);
        }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t ah[4], al[4];
        tf32_split(w[mt][0], ah[0], al[0]);
        tf32_split(w[mt][2], ah[1], al[1]);
        tf32_split(w[mt][1], ah[2], al[2]);
        tf32_split(w[mt][3], ah[3], al[3]);
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          mma_tf32(accT[mt][ct], al[0], al[1], al[2], al[3], gh[ct][0], gh[ct][1]);
          mma_tf32(accT[mt][ct], ah[0], ah[1], ah[2], ah[3], gl[ct][0], gl[ct][1]);
          mma_tf32(accT[mt][ct], ah[0], ah[1], ah[2], ah[3], gh[ct][0], gh[ct][1]);
        }
      }
    }
    acc_theta += (double)((th4[0] + th4[1]) + (th4[2] + th4[3]));
  }
  // block reduction (shared-memory atomics), then one partial slab per block as in lbs_bwd_kernel
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int bone = 16 * mt + g + 8 * (r >> 1), c = 8 * ct + 2 * q + (r & 1);
        if (bone < J && c < 12) atomicAdd(&sAcc[bone * 12 + c], accT[mt][ct][r]);
      }
  {
    double t = acc_theta;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const float thf = (float)(t / ((double)theta * (double)theta));
    const float a0 = warp_sum(acc_g[0]), a1 = warp_sum(acc_g[1]), a2 = warp_sum(acc_g[2]);
    if (lane == 0) {
      atomicAdd(&sAcc[n_out], thf);
      atomicAdd(&sAcc[n_out + 1], a0);
      atomicAdd(&sAcc[n_out + 2], a1);
      atomicAdd(&sAcc[n_out + 3], a2);
    }
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * (n_out + 4);
  for (int o = threadIdx.x; o < n_out + 4; o += blockDim.x) out[o] = sAcc[o];
}

#undef LBS_BONE_OK
static size_t lbs_bwd_mma_smem(int J, int MT) {
  return sizeof(float) * ((size_t)2 * MT * 2 * 4 * 32 + J * 12 + 4 + (size_t)LBS_MMA_WARPS * 32 * LBS_DG_LD);
}

// fixed-order reduction of the per-block partials (deterministic): one warp per output, lanes stride over the blocks
// and a fixed shuffle tree combines them
__global__ void __launch_bounds__(128)
lbs_bwd_reduce_kernel(const float* __restrict__ partial, int n_blocks, int J, const float* theta_weight, float eps,
                      float* __restrict__ d_theta, float* __restrict__ d_bone_T, float* __restrict__ d_global_t) {
  const int n_out = J * 12;
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (o >= n_out + 4) return;
  float s = 0.f;
  for (int b = lane; b < n_blocks; b += 32) s += partial[(size_t)b * (n_out + 4) + o];
  s = warp_sum(s);
  if (lane != 0) return;
  if (o < n_out) {
    const int j = o / 12, c = o - j * 12;
    d_bone_T[j * 16 + c] = s;
    if (c < 4) d_bone_T[j * 16 + 12 + c] = 0.f;
  } else if (o == n_out) {
    if (d_theta) d_theta[0] = (theta_weight && theta_weight[0] > eps) ? s : 0.f;   // torch.max(eps, theta): gradient to the larger
  } else if (d_global_t) {
    d_global_t[o - n_out - 1] = s;
  }
}

extern "C" int apn_lbs_fwd(const float* raw_w, const float* theta_weight, float eps, const int32_t* merge_rules,
                           const float* bone_T, const float* xyz, const float* global_t, int N, int J, float* xyz_out,
                           float* ginv_out, float* w_out, float* g_out, float* bbox, apn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  APN_CHECK_ARG(N > 0 && J > 0 && J <= LBS_MAX_J, "need N > 0 and 0 < J <= 128");
  APN_CHECK_ARG(raw_w && bone_T && xyz && xyz_out && ginv_out && bbox, "null pointer");
  const int JP = J | 1;
  const size_t smem = sizeof(float) * (J * 12 + J + (size_t)LBS_WARPS * 32 * JP);
  lbs_init_bbox_kernel<<<1, 32, 0, stream>>>(bbox);
  APN_LAUNCH_CHECK();
  const int blocks_needed = apn_div_up(apn_div_up(N, 32), LBS_WARPS);
  const int per_sm = (int)((200 * 1024) / (smem + 1024));
  const int cap = APN_SM_COUNT * (per_sm < 1 ? 1 : per_sm > 12 ? 12 : per_sm);
  const int grid = blocks_needed < cap ? blocks_needed : cap;
#define LBS_FWD_LAUNCH(KK)                                                                                              \
  do {                                                                                                                   \
    APN_CUDA(cudaFuncSetAttribute(lbs_fwd_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    lbs_fwd_kernel<KK><<<grid, 32 * LBS_WARPS, smem, stream>>>(raw_w, theta_weight, eps, merge_rules, bone_T, xyz,       \
                                                              global_t, N, J, xyz_out, ginv_out, w_out, g_out, bbox);   \
  } while (0)
  const int K = (J + 31) / 32;
  if (K == 1) LBS_FWD_LAUNCH(1);
  else if (K == 2) LBS_FWD_LAUNCH(2);
  else if (K == 3) LBS_FWD_LAUNCH(3);
  else LBS_FWD_LAUNCH(4);
#undef LBS_FWD_LAUNCH
  APN_LAUNCH_CHECK();
  return 0;
}

static size_t lbs_bwd_smem(int J, bool rules) {
  const int JP = J | 1;
  return sizeof(float) * (J * 12 + J + J * 12 + 4 + (size_t)LBS_WARPS * ((rules ? 3 : 2) * 32 * JP + 32 * 12));
}
static int lbs_bwd_grid(int N, int J) {
  const int blocks_needed = apn_div_up(apn_div_up(N, 32), LBS_WARPS);
  const int per_sm = (int)((200 * 1024) / (lbs_bwd_smem(J, false) + 1024));   // workspace sizing: the larger grid
  const int cap = APN_SM_COUNT * (per_sm < 1 ? 1 : per_sm > 8 ? 8 : per_sm);
  return blocks_needed < cap ? (blocks_needed > 0 ? blocks_needed : 1) : cap;
}

// tensor-core backward: persistent grid of up to 4 blocks per SM (the occupancy API gives the resident count per MT)
static int lbs_bwd_mma_grid_cap(int N) {
  const int blocks_needed = apn_div_up(apn_div_up(N, 32), LBS_MMA_WARPS);
  const int cap = APN_SM_COUNT * 4;
  return blocks_needed < cap ? (blocks_needed > 0 ? blocks_needed : 1) : cap;
}

extern "C" size_t apn_lbs_bwd_workspace_bytes(int N, int J) {
  const int g = lbs_bwd_grid(N, J), gm = lbs_bwd_mma_grid_cap(N);
  return sizeof(float) * (size_t)(g > gm ? g : gm) * (J * 12 + 4);
}

extern "C" int apn_lbs_bwd(const float* raw_w, const float* theta_weight, float eps, const int32_t* merge_rules,
                           const float* bone_T, const float* xyz, int N, int J, const float* ginv, const float* d_xyz,
                           const float* d_ginv, const float* d_w, const float* d_g, float* d_raw, float* d_theta,
                           float* d_bone_T,
                           float* d_global_t, void* workspace, size_t workspace_bytes, apn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  APN_CHECK_ARG(N > 0 && J > 0 && J <= LBS_MAX_J, "need N > 0 and 0 < J <= 128");
  APN_CHECK_ARG(raw_w && bone_T && xyz && ginv && d_raw && d_bone_T && workspace, "null pointer");
  APN_CHECK_ARG(!theta_weight || d_theta, "d_theta is required when theta_weight is given");
  APN_CHECK_ARG(workspace_bytes >= apn_lbs_bwd_workspace_bytes(N, J), "workspace too small");
  const int grid = lbs_bwd_grid(N, J);
  if (theta_weight && !merge_rules && !d_g && J <= 80) {          // tensor-core path
    const int MT = (J + 15) / 16;
    const size_t sm = lbs_bwd_mma_smem(J, MT);
    int g2 = lbs_bwd_mma_grid_cap(N);                               // never more slabs than the workspace holds
#define LBS_BWD_MMA_LAUNCH(MM)                                                                                          \
  do {                                                                                                                   \
    APN_CUDA(cudaFuncSetAttribute(lbs_bwd_mma_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));        \
    int per_sm = 0;                                                                                                      \
    APN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbs_bwd_mma_kernel<MM>, 32 * LBS_MMA_WARPS, sm));    \
    per_sm = per_sm < 1 ? 1 : per_sm > 4 ? 4 : per_sm;                                                                   \
    if (g2 > APN_SM_COUNT * per_sm) g2 = APN_SM_COUNT * per_sm;                                                          \
    lbs_bwd_mma_kernel<MM><<<g2, 32 * LBS_MMA_WARPS, sm, stream>>>(raw_w, theta_weight, eps, bone_T, xyz, N, J, ginv,    \
                                                                  d_xyz, d_ginv, d_w, d_raw, (float*)workspace);        \
  } while (0)
    if (MT == 1) LBS_BWD_MMA_LAUNCH(1);
    else if (MT == 2) LBS_BWD_MMA_LAUNCH(2);
    else if (MT == 3) LBS_BWD_MMA_LAUNCH(3);
    else if (MT == 4) LBS_BWD_MMA_LAUNCH(4);
    else LBS_BWD_MMA_LAUNCH(5);
#undef LBS_BWD_MMA_LAUNCH
    APN_LAUNCH_CHECK();
    const int n = J * 12 + 4;
    lbs_bwd_reduce_kernel<<<(n + 3) / 4, 128, 0, stream>>>((const float*)workspace, g2, J, theta_weight, eps, d_theta,
                                                            d_bone_T, d_global_t);
    APN_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = lbs_bwd_smem(J, merge_rules != nullptr);
  APN_CHECK_ARG(smem <= 227 * 1024, "J too large for the backward tile");
#define LBS_BWD_LAUNCH(KK)                                                                                               \
  do {                                                                                                                    \
    APN_CUDA(cudaFuncSetAttribute(lbs_bwd_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    lbs_bwd_kernel<KK><<<grid, 32 * LBS_WARPS, smem, stream>>>(raw_w, theta_weight, eps, merge_rules, bone_T, xyz, N, J,  \
                                                              ginv, d_xyz, d_ginv, d_w, d_g, d_raw, (float*)workspace);  \
  } while (0)
  const int K = (J + 31) / 32;
  if (K == 1) LBS_BWD_LAUNCH(1);
  else if (K == 2) LBS_BWD_LAUNCH(2);
  else if (K == 3) LBS_BWD_LAUNCH(3);
  else LBS_BWD_LAUNCH(4);
#undef LBS_BWD_LAUNCH
  APN_LAUNCH_CHECK();
  const int n = J * 12 + 4;
  lbs_bwd_reduce_kernel<<<(n + 3) / 4, 128, 0, stream>>>((const float*)workspace, grid, J, theta_weight, eps, d_theta,
                                                            d_bone_T, d_global_t);
  APN_LAUNCH_CHECK();
  return 0;
}
