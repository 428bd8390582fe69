// K1 — fused skinning: softmax(raw/theta) -> merge rules -> blend of bone transforms -> point
// transform (+ global translation) -> closed-form inverse of the blended 3x3 frame -> bbox.
// Replaces lib/temporalpoints.py:401-414,424,569 and lib/pointwarper.py:241-266.
//
// Bandwidth-bound: algorithmic bytes fwd = N*(4J + 12 + 12 + 36) (+4J when merged weights are
// written); bwd = N*(8J + 12 + 12 + 36 + 36).  One tile = 128 points; the (128 x J) weight tile
// is staged through shared memory with float4-coalesced global access and an odd row stride so
// that the per-point passes are bank-conflict free; bone matrices live in shared memory and
// are read as warp broadcasts.  Persistent grid (multiple of the SM count).
#include "common.cuh"

#define LBS_TILE 128
#define LBS_MAX_J 128

__global__ void lbs_init_bbox_kernel(float* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = __int_as_float(0x7f800000);       // +inf
  else if (threadIdx.x < 6) bbox[threadIdx.x] = __int_as_float(0xff800000);  // -inf
}

// cooperative (LBS_TILE x J) global <-> shared copy; global rows are contiguous (stride J),
// shared rows have stride JP.
__device__ __forceinline__ void tile_load(float* __restrict__ s, const float* __restrict__ g, int n_valid, int J, int JP) {
  const int total = n_valid * J;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = i / J, j = i - p * J;
    s[p * JP + j] = __ldg(g + i);
  }
}
__device__ __forceinline__ void tile_store(float* __restrict__ g, const float* __restrict__ s, int n_valid, int J, int JP) {
  const int total = n_valid * J;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = i / J, j = i - p * J;
    g[i] = s[p * JP + j];
  }
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

__global__ void __launch_bounds__(LBS_TILE)
lbs_fwd_kernel(const float* __restrict__ raw_w, const float* __restrict__ theta_weight, float eps,
               const int* __restrict__ rules, const float* __restrict__ bone_T, const float* __restrict__ xyz,
               const float* __restrict__ global_t, int N, int J, float* __restrict__ xyz_out,
               float* __restrict__ ginv_out, float* __restrict__ w_out, float* __restrict__ g_out,
               float* __restrict__ bbox) {
  extern __shared__ float smem[];
  const int JP = J | 1;
  float* sT = smem;                    // J*12
  int* sR = (int*)(sT + J * 12);       // J
  float* sW = (float*)(sR + J);        // LBS_TILE*JP
  for (int i = threadIdx.x; i < J * 12; i += blockDim.x) {
    const int j = i / 12, c = i - j * 12;
    sT[i] = bone_T[j * 16 + c];        // rows 0..2 of the 4x4
  }
  for (int j = threadIdx.x; j < J; j += blockDim.x) sR[j] = rules ? rules[j] : j;
  const float theta = theta_weight ? fmaxf(eps, theta_weight[0]) : 1.f;
  const float gx = global_t ? global_t[0] : 0.f, gy = global_t ? global_t[1] : 0.f, gz = global_t ? global_t[2] : 0.f;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mxv[3] = {-INFINITY, -INFINITY, -INFINITY};
  const int n_tiles = (N + LBS_TILE - 1) / LBS_TILE;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int base = tile * LBS_TILE;
    const int n_valid = min(LBS_TILE, N - base);
    __syncthreads();
    tile_load(sW, raw_w + (size_t)base * J, n_valid, J, JP);
    __syncthreads();
    const int p = threadIdx.x;
    if (p < n_valid) {
      float* row = sW + p * JP;
      if (theta_weight) softmax_row(row, J, theta);   // NULL: the caller passes final weights
      if (rules) {
        for (int j = 0; j < J; ++j) {
          const int t = sR[j];
          if (t != j) {
            row[t] += row[j];
            row[j] = 0.f;
          }
        }
      }
      float G[12];
#pragma unroll
      for (int c = 0; c < 12; ++c) G[c] = 0.f;
      for (int j = 0; j < J; ++j) {
        const float w = row[j];
        const float* T = sT + j * 12;
#pragma unroll
        for (int c = 0; c < 12; ++c) G[c] = fmaf(w, T[c], G[c]);
      }
      const size_t n = (size_t)base + p;
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
    if (w_out) {
      __syncthreads();
      tile_store(w_out + (size_t)base * J, sW, n_valid, J, JP);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = warp_min(mn[c]), b = warp_max(mxv[c]);
    if ((threadIdx.x & 31) == 0) {
      if (a < INFINITY) atomic_min_float(bbox + c, a);
      if (b > -INFINITY) atomic_max_float(bbox + 3 + c, b);
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward
// partial layout per block: [J*12 dT | 1 dtheta | 3 dglobal_t]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LBS_TILE)
lbs_bwd_kernel(const float* __restrict__ raw_w, const float* __restrict__ theta_weight, float eps,
               const int* __restrict__ rules, const float* __restrict__ bone_T, const float* __restrict__ xyz, int N,
               int J, const float* __restrict__ ginv, const float* __restrict__ d_xyz,
               const float* __restrict__ d_ginv, const float* __restrict__ d_w, const float* __restrict__ d_g,
               float* __restrict__ d_raw, float* __restrict__ partial) {
  extern __shared__ float smem[];
  const int JP = J | 1;
  const int n_out = J * 12;
  float* sT = smem;                          // J*12
  int* sR = (int*)(sT + n_out);              // J
  float* sW = (float*)(sR + J);              // softmax weights, later d_raw          TILE*JP
  float* sM = sW + LBS_TILE * JP;            // merged weights                         TILE*JP
  float* sDW = sM + LBS_TILE * JP;           // incoming d_w tile                      TILE*JP
  float* sDG = sDW + LBS_TILE * JP;          // dG per point                           TILE*13
  float* sAcc = sDG + LBS_TILE * 13;         // dT accumulators                        J*12
  __shared__ float sRed[4][4];
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    const int j = i / 12, c = i - j * 12;
    sT[i] = bone_T[j * 16 + c];
    sAcc[i] = 0.f;
  }
  for (int j = threadIdx.x; j < J; j += blockDim.x) sR[j] = rules ? rules[j] : j;
  const float theta = theta_weight ? fmaxf(eps, theta_weight[0]) : 1.f;
  float acc_theta = 0.f, acc_g[3] = {0.f, 0.f, 0.f};
  const int n_tiles = (N + LBS_TILE - 1) / LBS_TILE;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int base = tile * LBS_TILE;
    const int n_valid = min(LBS_TILE, N - base);
    __syncthreads();
    tile_load(sW, raw_w + (size_t)base * J, n_valid, J, JP);
    if (d_w) tile_load(sDW, d_w + (size_t)base * J, n_valid, J, JP);
    __syncthreads();
    const int p = threadIdx.x;
    float* dg = sDG + p * 13;
    if (p < n_valid) {
      float* row = sW + p * JP;
      float* mrow = sM + p * JP;
      if (theta_weight) softmax_row(row, J, theta);
      for (int j = 0; j < J; ++j) mrow[j] = row[j];
      if (rules) {
        for (int j = 0; j < J; ++j) {
          const int t = sR[j];
          if (t != j) {
            mrow[t] += mrow[j];
            mrow[j] = 0.f;
          }
        }
      }
      const size_t n = (size_t)base + p;
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
      dg[0] = dA[0] + gx * x; dg[1] = dA[1] + gx * y; dg[2] = dA[2] + gx * z; dg[3] = gx;
      dg[4] = dA[3] + gy * x; dg[5] = dA[4] + gy * y; dg[6] = dA[5] + gy * z; dg[7] = gy;
      dg[8] = dA[6] + gz * x; dg[9] = dA[7] + gz * y; dg[10] = dA[8] + gz * z; dg[11] = gz;
      if (d_g) {
#pragma unroll
        for (int c = 0; c < 12; ++c) dg[c] += d_g[16 * n + c];
      }
      // dm_j, then dw_j = dm_{rules[j]}, softmax backward
      float* dwrow = sDW + p * JP;
      for (int j = 0; j < J; ++j) {
        const float* T = sT + j * 12;
        float s = d_w ? dwrow[j] : 0.f;
#pragma unroll
        for (int c = 0; c < 12; ++c) s = fmaf(dg[c], T[c], s);
        dwrow[j] = s;  // dm_j
      }
      if (theta_weight) {
        float dot = 0.f;
        for (int j = 0; j < J; ++j) dot = fmaf(row[j], dwrow[sR[j]], dot);
        float th = 0.f;
        for (int j = 0; j < J; ++j) {
          const float dz = row[j] * (dwrow[sR[j]] - dot);
          const float rawv = __ldg(raw_w + n * J + j);
          th = fmaf(-dz, rawv, th);
          row[j] = dz / theta;  // d_raw
        }
        acc_theta += th / (theta * theta);
      } else {
        for (int j = 0; j < J; ++j) row[j] = dwrow[sR[j]];
      }
    } else {
#pragma unroll
      for (int c = 0; c < 12; ++c) dg[c] = 0.f;
      float* mrow = sM + p * JP;
      for (int j = 0; j < J; ++j) mrow[j] = 0.f;
    }
    __syncthreads();
    tile_store(d_raw + (size_t)base * J, sW, n_valid, J, JP);
    // dT_j += sum_p m[p][j] * dG[p][:]
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
      const int j = o / 12, c = o - j * 12;
      float a = 0.f;
#pragma unroll 8
      for (int q = 0; q < LBS_TILE; ++q) a = fmaf(sM[q * JP + j], sDG[q * 13 + c], a);
      sAcc[o] += a;
    }
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * (n_out + 4);
  for (int o = threadIdx.x; o < n_out; o += blockDim.x) out[o] = sAcc[o];
  float v[4] = {acc_theta, acc_g[0], acc_g[1], acc_g[2]};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float s = warp_sum(v[c]);
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5][c] = s;
  }
  __syncthreads();
  if (threadIdx.x < 4) out[n_out + threadIdx.x] = sRed[0][threadIdx.x] + sRed[1][threadIdx.x] + sRed[2][threadIdx.x] + sRed[3][threadIdx.x];
}

// fixed-order reduction of the per-block partials (deterministic)
__global__ void lbs_bwd_reduce_kernel(const float* __restrict__ partial, int n_blocks, int J, const float* theta_weight,
                                      float eps, float* __restrict__ d_theta, float* __restrict__ d_bone_T,
                                      float* __restrict__ d_global_t) {
  const int n_out = J * 12;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out + 4) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += partial[(size_t)b * (n_out + 4) + o];
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

static int lbs_grid(int N) {
  const int tiles = (N + LBS_TILE - 1) / LBS_TILE;
  const int cap = APN_SM_COUNT * 8;
  return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

extern "C" int apn_lbs_fwd(const float* raw_w, const float* theta_weight, float eps, const int32_t* merge_rules,
                           const float* bone_T, const float* xyz, const float* global_t, int N, int J, float* xyz_out,
                           float* ginv_out, float* w_out, float* g_out, float* bbox, apn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  APN_CHECK_ARG(N > 0 && J > 0 && J <= LBS_MAX_J, "need N > 0 and 0 < J <= 128");
  APN_CHECK_ARG(raw_w && bone_T && xyz && xyz_out && ginv_out && bbox, "null pointer");
  const int JP = J | 1;
  const size_t smem = sizeof(float) * (J * 12 + J + (size_t)LBS_TILE * JP);
  APN_CUDA(cudaFuncSetAttribute(lbs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lbs_init_bbox_kernel<<<1, 32, 0, stream>>>(bbox);
  APN_LAUNCH_CHECK();
  lbs_fwd_kernel<<<lbs_grid(N), LBS_TILE, smem, stream>>>(raw_w, theta_weight, eps, merge_rules, bone_T, xyz, global_t, N,
                                                         J, xyz_out, ginv_out, w_out, g_out, bbox);
  APN_LAUNCH_CHECK();
  return 0;
}

static int lbs_bwd_grid(int N) {
  const int tiles = (N + LBS_TILE - 1) / LBS_TILE;
  const int cap = APN_SM_COUNT * 4;
  return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

extern "C" size_t apn_lbs_bwd_workspace_bytes(int N, int J) {
  return sizeof(float) * (size_t)lbs_bwd_grid(N) * (J * 12 + 4);
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
  const int JP = J | 1;
  const size_t smem = sizeof(float) * (J * 12 + J + 3 * (size_t)LBS_TILE * JP + LBS_TILE * 13 + J * 12);
  APN_CHECK_ARG(smem <= 227 * 1024, "J too large for the backward tile");
  APN_CUDA(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = lbs_bwd_grid(N);
  lbs_bwd_kernel<<<grid, LBS_TILE, smem, stream>>>(raw_w, theta_weight, eps, merge_rules, bone_T, xyz, N, J, ginv, d_xyz,
                                                  d_ginv, d_w, d_g, d_raw, (float*)workspace);
  APN_LAUNCH_CHECK();
  const int n = J * 12 + 4;
  lbs_bwd_reduce_kernel<<<(n + 127) / 128, 128, 0, stream>>>((const float*)workspace, grid, J, theta_weight, eps, d_theta,
                                                            d_bone_T, d_global_t);
  APN_LAUNCH_CHECK();
  return 0;
}
