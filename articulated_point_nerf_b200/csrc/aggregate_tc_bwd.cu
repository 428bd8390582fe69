// K3 backward on the tensor cores (training path; split-fp16 operands, fp32 accumulation in tensor memory).
// Counterpart of aggregate_tc.cu; replaces autograd through lib/temporalpoints.py:446-494.
//
//   tc_dgrad_kernel   per 128-row tile: dY3 = idw * d_h * LeakyReLU'(act3)  ->  three dgrad MMAs
//                     dY_{l-1} = (dY_l W_l) * LeakyReLU'_{l-1}  ->  dPE = dY0 W0_pe (N = 64)  ->  PE backward,
//                     IDW backward, scatter into d_xyz / d_ginv and into the per-point table gradient dP;
//                     every dY_l tile is also written out as a shared-memory image for the wgrad kernel.
//   tc_wgrad_kernel   dW_l += dY_l^T X_l over all tiles (X_0 = PE tile, X_l = act_{l-1} from the tape): both
//                     operands are the SAME [row][column] tiles, read MN-major (the reduction runs over rows);
//                     db_l comes from one extra N=8 MMA against a tile of ones.  Accumulators for all four
//                     layers stay in tensor memory (480 of 512 columns) until the CTA has consumed its tiles.
// The feature columns of layer 0 go through the per-point table: dP (N x 128) is scattered here and turned
// into d_feat = dP W0_feat and dW0_feat = dP^T feat by two exact fp32 GEMMs over the N points.
#include "tgemm.cuh"
#include "aggregate_tc.cuh"

#define TCB_NCHUNKS 8          // W3^T (2), W2^T (2), W1^T (2), W0_pe^T (2): [n = in feature][k = out feature] tiles
#define TCB_SLOTS 3

// ---------------------------------------------------------------------------------------
// packed transposed weights
// ---------------------------------------------------------------------------------------
__global__ void tc_pack_bwd_kernel(const apn_mlp_weights w, int d_in, uint8_t* __restrict__ packed) {
  const int c = blockIdx.x;
  const int layer = 3 - c / 2, kc = c & 1;
  const float* W = w.w[layer];
  const int ld = layer == 0 ? d_in : APN_C;
  for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
    const int n = e >> 6, k = e & 63;                 // n: input feature (row of the B tile), k: output feature
    if (layer == 0 && n >= 64) continue;              // the layer-0 tile has 64 rows (the PE tile's column order)
    const int col = layer == 0 ? tc_pe_ref_col(n) : n;
    const float v = col >= 0 ? W[(size_t)(kc * 64 + k) * ld + col] : 0.f;
    __half hi, lo;
    split_half(v, hi, lo);
    uint8_t* base = packed + (size_t)c * TC_CHUNK_GBYTES;
    const uint32_t o = sw128_offset(n, k);
    *reinterpret_cast<__half*>(base + o) = hi;
    *reinterpret_cast<__half*>(base + TC_TILE_BYTES + o) = lo;
  }
}

extern "C" size_t apn_aggregate_tc_bwd_weights_bytes(void) { return (size_t)TCB_NCHUNKS * TC_CHUNK_GBYTES; }

extern "C" int apn_aggregate_tc_pack_weights_bwd(const apn_mlp_weights* w, int d_in, void* packed, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  APN_CHECK_ARG(w && packed, "null pointer");
  APN_CHECK_ARG(d_in >= APN_PE_POS + APN_C && d_in <= 256, "d_in must be 191..256");
  APN_CHECK_ARG((((uintptr_t)packed) & 15) == 0, "packed weights must be 16-byte aligned");
  APN_CUDA(cudaMemsetAsync(packed, 0, apn_aggregate_tc_bwd_weights_bytes(), st));
  tc_pack_bwd_kernel<<<TCB_NCHUNKS, 256, 0, st>>>(*w, d_in, (uint8_t*)packed);
  APN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// dgrad chain
// ---------------------------------------------------------------------------------------
struct TcBwdParams {
  apn_agg_inputs in;
  const float* d_h;           // (M,128) gradient on the reduced feature (heads backward)
  const float* idw;           // (M,8)
  const uint8_t* tape;
  const uint8_t* packed;      // transposed weights
  uint8_t* dy;                // n_tiles * TC_DY_TILE_BYTES
  float* d_ptable;            // (N,128), zeroed by the host
  float* d_xyz;               // (N,3) accumulated, or NULL
  float* d_ginv;              // (N,9) accumulated, or NULL
  const float* hmax;          // max |d_h| (device scalar): sets the power-of-two gradient scale
  int n_tiles;
};

// Gradients are far below fp16's normal range (6e-5), activations are not: every dY is carried as S * dY with one
// power of two S per launch (max |d_h| -> 2^10, leaving 2^6 of head-room for growth down the chain) and unscaled
// where it leaves the tensor cores (dP, dPE, dW, db).
__device__ __forceinline__ float tc_grad_scale(const float* hmax) {
  const float m = __ldg(hmax);
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  const float e = fminf(fmaxf(floorf(log2f(1024.f / m)), -100.f), 100.f);
  return exp2f(e);
}

struct TcBwdSmem {
  static constexpr int OFF_A = 0;                                   // dY operand: (kc, split) tiles
  static constexpr int OFF_W = OFF_A + 4 * TC_TILE_BYTES;
  static constexpr int OFF_DOT = OFF_W + TCB_SLOTS * 2 * TC_TILE_BYTES;   // 128 floats
  static constexpr int OFF_DRC = OFF_DOT + 128 * 4;                  // 128 x 3 floats
  static constexpr int OFF_IDX = OFF_DRC + 128 * 3 * 4;              // 128 ints
  static constexpr int OFF_BAR = OFF_IDX + 128 * 4;
  static constexpr int N_BAR = 2 * TCB_SLOTS + 4;
  static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1) tc_dgrad_kernel(const TcBwdParams p) {
  using S = TcBwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem + S::OFF_A;
  uint8_t* sW = smem + S::OFF_W;
  float* sDot = (float*)(smem + S::OFF_DOT);
  float* sDrc = (float*)(smem + S::OFF_DRC);
  int* sIdx = (int*)(smem + S::OFF_IDX);
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_free = bars + TCB_SLOTS;
  uint64_t* a_ready = bars + 2 * TCB_SLOTS;        // [2]
  uint64_t* acc_ready = bars + 2 * TCB_SLOTS + 2;  // [2]
  uint32_t* sTmem = (uint32_t*)(smem + S::OFF_TMEM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const apn_agg_inputs& in = p.in;
  const int Mrt = apn_rt_count(in.m_dev, in.M);                     // sample count: host-exact or read from the device counter
  const int n_tiles = (Mrt + TC_SAMPLES - 1) / TC_SAMPLES;

  if (tid == 0) {
    for (int i = 0; i < TCB_SLOTS; ++i) {
      mbar_init(w_full + i, 1);
      mbar_init(w_free + i, 1);
    }
    mbar_init(a_ready, TC_COMPUTE_THREADS);
    mbar_init(a_ready + 1, TC_COMPUTE_THREADS);
    mbar_init(acc_ready, 1);
    mbar_init(acc_ready + 1, 1);
    fence_barrier_init();
  }
  if (warp == TC_COMPUTE_WARPS + 1) tmem_alloc<TC_TMEM_COLS>(sTmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sTmem;

  if (warp == TC_COMPUTE_WARPS) {
    // ================================================================= weight producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int c = 0; c < TCB_NCHUNKS; ++c, ++it) {
          const uint32_t slot = it % TCB_SLOTS;
          mbar_wait(w_free + slot, ((it / TCB_SLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(w_full + slot, 2 * TC_TILE_BYTES);
          bulk_g2s(sW + (size_t)slot * 2 * TC_TILE_BYTES, p.packed + (size_t)c * TC_CHUNK_GBYTES, 2 * TC_TILE_BYTES, w_full + slot);
        }
    }
  } else if (warp == TC_COMPUTE_WARPS + 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t a_base = smem_u32(sA), w_base = smem_u32(sW);
      uint32_t it = 0, ph_a0 = 0, ph_a1 = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int g = 0; g < 4; ++g) {                   // g = 0..2: dX_{3-g}; g = 3: dPE (N = 64)
          const uint32_t idesc = umma_idesc_f16(128, g == 3 ? 64 : 128);
          const uint32_t acc = tmem_base + (uint32_t)((g & 1) * 128);
          for (int kc = 0; kc < 2; ++kc, ++it) {
            if (kc == 0) {
              mbar_wait(a_ready, ph_a0);
              ph_a0 ^= 1;
            } else {
              mbar_wait(a_ready + 1, ph_a1);
              ph_a1 ^= 1;
            }
            tc_fence_after();
            // the dY operand tile is also the wgrad kernel's input: stream it out from shared memory
            bulk_s2g(p.dy + (size_t)tile * TC_DY_TILE_BYTES + TC_TAPE_ACT(3 - g, kc, 0), sA + (size_t)(kc * 2) * TC_TILE_BYTES,
                     2 * TC_TILE_BYTES);
            bulk_commit();
            if (kc == 1) bulk_wait_read();      // the epilogue released by this group's commit overwrites both tiles
            const uint32_t slot = it % TCB_SLOTS;
            mbar_wait(w_full + slot, (it / TCB_SLOTS) & 1);
            tc_fence_after();
            const uint32_t a_hi = a_base + (uint32_t)(kc * 2) * TC_TILE_BYTES;
            const uint32_t b_hi = w_base + slot * (uint32_t)(2 * TC_TILE_BYTES);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = umma_desc_k_sw128(a_hi + ks * 32), db = umma_desc_k_sw128(b_hi + ks * 32);
              const uint64_t da_lo = umma_desc_k_sw128(a_hi + TC_TILE_BYTES + ks * 32);
              const uint64_t db_lo = umma_desc_k_sw128(b_hi + TC_TILE_BYTES + ks * 32);
              umma_f16(acc, da, db, idesc, (kc | ks) ? 1u : 0u);
              umma_f16(acc, da, db_lo, idesc, 1u);
              umma_f16(acc, da_lo, db, idesc, 1u);
            }
            umma_commit(w_free + slot);
          }
          umma_commit(acc_ready + (g & 1));
        }
      }
      bulk_wait_all();
    }
  } else {
    // ================================================================= compute warps
    uint32_t ph_acc0 = 0, ph_acc1 = 0;
    const int q = warp & 3, cq = warp >> 2;
    const int erow = q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const float gscale = tc_grad_scale(p.hmax), inv_gscale = 1.f / gscale;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = tile * TC_SAMPLES;
      const uint8_t* tp = p.tape + (size_t)tile * TC_TAPE_TILE_BYTES;
      const int ms = m0 + (erow >> 3);
      const bool valid = ms < Mrt;
      const int mc = min(ms, Mrt - 1);
      if (tid < 128) {
        sDot[tid] = 0.f;
        sDrc[3 * tid] = 0.f; sDrc[3 * tid + 1] = 0.f; sDrc[3 * tid + 2] = 0.f;
        sIdx[tid] = __ldg(in.nn_idx + (size_t)min(m0 + (tid >> 3), Mrt - 1) * APN_K + (tid & 7));
      }
      compute_sync();
      // ---------------------------------------------------------------- dY3 = idw * d_h * LeakyReLU'(act3); dw_k = <d_h, act3_k>
      {
        const float w = valid ? __ldg(p.idw + (size_t)mc * APN_K + (erow & 7)) : 0.f;
        const float ws = w * gscale;
        float dot = 0.f;
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          const float4* dh4 = reinterpret_cast<const float4*>(p.d_h + (size_t)mc * APN_C + ph * 64 + cq * 16);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t o = (uint32_t)(erow * 128 + ((((cq * 2 + u) ^ (erow & 7)) & 7) << 4));
            const uint4 ah = __ldg(reinterpret_cast<const uint4*>(tp + TC_TAPE_ACT(3, ph, 0) + o));
            const uint4 al = __ldg(reinterpret_cast<const uint4*>(tp + TC_TAPE_ACT(3, ph, 1) + o));
            const float4 d0 = __ldg(dh4 + 2 * u), d1 = __ldg(dh4 + 2 * u + 1);
            const float dh[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            const uint32_t ahw[4] = {ah.x, ah.y, ah.z, ah.w}, alw[4] = {al.x, al.y, al.z, al.w};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&ahw[e]));
              const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&alw[e]));
              const float a0 = fh.x + fl.x, a1 = fh.y + fl.y;
              dot = fmaf(dh[2 * e], a0, dot);
              dot = fmaf(dh[2 * e + 1], a1, dot);
              const float g0 = ws * dh[2 * e] * (a0 > 0.f ? 1.f : 0.01f), g1 = ws * dh[2 * e + 1] * (a1 > 0.f ? 1.f : 0.01f);
              split_half2(g0, g1, hi[e], lo[e]);
            }
            uint8_t* t = sA + (size_t)(ph * 2) * TC_TILE_BYTES + o;
            *reinterpret_cast<uint4*>(t) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(t + TC_TILE_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async_smem();
          mbar_arrive(a_ready + ph);
        }
        atomicAdd(&sDot[erow], valid ? dot : 0.f);
      }
      // ---------------------------------------------------------------- dY_{l} = (dY_{l+1} W_{l+1}) * LeakyReLU'_l,  l = 2, 1, 0
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int l = 2 - g;                               // layer whose pre-activation gradient is produced
        if (g & 1) {
          mbar_wait(acc_ready + 1, ph_acc1);
          ph_acc1 ^= 1;
        } else {
          mbar_wait(acc_ready, ph_acc0);
          ph_acc0 ^= 1;
        }
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)((g & 1) * 128) + tlane;
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          uint32_t v[16];
          tmem_ld16(tacc + ph * 64 + cq * 16, v);
          tmem_ld_wait();
          const uint32_t mb = __ldg(reinterpret_cast<const uint16_t*>(tp + TC_TAPE_MASK(l) + ((size_t)erow * 8 + ph * 4 + cq) * 2));
          float y[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(v[i]) * (((mb >> i) & 1u) ? 1.f : 0.01f);
          if (l == 0 && valid) {
            // feature columns of layer 0: gradient of the per-point table row
            float* dp = p.d_ptable + (size_t)sIdx[erow] * APN_C + ph * 64 + cq * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              red_add_v4(dp + i, y[i] * inv_gscale, y[i + 1] * inv_gscale, y[i + 2] * inv_gscale, y[i + 3] * inv_gscale);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_half2(y[u * 8 + 2 * e], y[u * 8 + 2 * e + 1], hi[e], lo[e]);
            const uint32_t o = (uint32_t)(erow * 128 + ((((cq * 2 + u) ^ (erow & 7)) & 7) << 4));
            uint8_t* t = sA + (size_t)(ph * 2) * TC_TILE_BYTES + o;
            *reinterpret_cast<uint4*>(t) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(t + TC_TILE_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(a_ready + ph);
        }
      }
      // ---------------------------------------------------------------- dPE = dY0 W0_pe  ->  d rel_c (PE backward)
      mbar_wait(acc_ready + 1, ph_acc1);
      ph_acc1 ^= 1;
      tc_fence_after();
      {
        uint32_t v[16];
        tmem_ld16(tmem_base + 128u + tlane + cq * 16, v);
        tmem_ld_wait();
        tc_fence_before();
        // recompute rel_c of this row (lib/temporalpoints.py:478-480)
        const int idx = sIdx[erow];
        const float px = __ldg(in.pts + 3 * (size_t)mc), py = __ldg(in.pts + 3 * (size_t)mc + 1), pz = __ldg(in.pts + 3 * (size_t)mc + 2);
        const float rx = px - __ldg(in.xyz + 3 * (size_t)idx), ry = py - __ldg(in.xyz + 3 * (size_t)idx + 1),
                    rz = pz - __ldg(in.xyz + 3 * (size_t)idx + 2);
        const float* G = in.ginv + 9 * (size_t)idx;
        const float rc0 = __ldg(G) * rx + __ldg(G + 1) * ry + __ldg(G + 2) * rz;
        const float rc1 = __ldg(G + 3) * rx + __ldg(G + 4) * ry + __ldg(G + 5) * rz;
        const float rc2 = __ldg(G + 6) * rx + __ldg(G + 7) * ry + __ldg(G + 8) * rz;
        // d rel_c[d] = g_x[d] + sum_i 2^i (cos_i g_sin_i - sin_i g_cos_i), in the PE tile's column layout (aggregate_tc.cuh)
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        if (cq < 3) {                                       // columns 16 d + [sin i = 0..7 | cos i = 0..7]
          float sn[8], cs[8];
          tc_pe_octaves<0, 8>(cq == 0 ? rc0 : cq == 1 ? rc1 : rc2, sn, cs);
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            acc += (float)(1 << i) * (cs[i] * __uint_as_float(v[i]) - sn[i] * __uint_as_float(v[8 + i]));
          acc *= inv_gscale;
          d0 = cq == 0 ? acc : 0.f;
          d1 = cq == 1 ? acc : 0.f;
          d2 = cq == 2 ? acc : 0.f;
        } else {                                            // 4 d + [sin 8, sin 9 | cos 8, cos 9], then rel_c
          float dd[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float sn[2], cs[2];
            tc_pe_octaves<8, 2>(d == 0 ? rc0 : d == 1 ? rc1 : rc2, sn, cs);
            dd[d] = 256.f * (cs[0] * __uint_as_float(v[4 * d]) - sn[0] * __uint_as_float(v[4 * d + 2])) +
                    512.f * (cs[1] * __uint_as_float(v[4 * d + 1]) - sn[1] * __uint_as_float(v[4 * d + 3])) +
                    __uint_as_float(v[12 + d]);
          }
          d0 = dd[0] * inv_gscale;
          d1 = dd[1] * inv_gscale;
          d2 = dd[2] * inv_gscale;
        }
        atomicAdd(&sDrc[3 * erow], d0);
        atomicAdd(&sDrc[3 * erow + 1], d1);
        atomicAdd(&sDrc[3 * erow + 2], d2);
      }
      compute_sync();
      // ---------------------------------------------------------------- IDW backward + scatter (one thread per row)
      if (tid < 128) {
        const int r = tid, m = m0 + (r >> 3);
        const int mm = min(m, Mrt - 1);
        const int idx = sIdx[r];
        const float px = __ldg(in.pts + 3 * (size_t)mm), py = __ldg(in.pts + 3 * (size_t)mm + 1), pz = __ldg(in.pts + 3 * (size_t)mm + 2);
        const float rp[3] = {px - __ldg(in.xyz + 3 * (size_t)idx), py - __ldg(in.xyz + 3 * (size_t)idx + 1),
                             pz - __ldg(in.xyz + 3 * (size_t)idx + 2)};
        const float u = 1.0f / ((rp[0] * rp[0] + rp[1] * rp[1]) + rp[2] * rp[2] + in.eps);
        float su = u;
        su += __shfl_xor_sync(0xffffffffu, su, 1);
        su += __shfl_xor_sync(0xffffffffu, su, 2);
        su += __shfl_xor_sync(0xffffffffu, su, 4);
        const float w = u / su, dwk = sDot[r];
        float dot = dwk * w;
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        // w_k = u_k / S:  d_u = (d_w - <d_w, w>) / S ;  d_d2 = -d_u u^2
        const float d_d2 = -((dwk - dot) / su) * u * u;
        if (m < Mrt) {
          const float dc[3] = {sDrc[3 * r], sDrc[3 * r + 1], sDrc[3 * r + 2]};
          const float* G = in.ginv + 9 * (size_t)idx;
          if (p.d_ginv) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int c = 0; c < 3; ++c) atomicAdd(p.d_ginv + 9 * (size_t)idx + 3 * a + c, dc[a] * rp[c]);
          }
          if (p.d_xyz) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float g = __ldg(G + c) * dc[0] + __ldg(G + 3 + c) * dc[1] + __ldg(G + 6 + c) * dc[2] + 2.f * rp[c] * d_d2;
              atomicAdd(p.d_xyz + 3 * (size_t)idx + c, -g);
            }
          }
        }
      }
      compute_sync();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_COMPUTE_WARPS + 1) tmem_dealloc<TC_TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------
struct TcWgradParams {
  const uint8_t* tape;
  const uint8_t* dy;
  float* partial;             // gridDim.x slabs of TCW_SLAB floats: per-CTA sums, reduced by tc_wgrad_reduce_kernel
  int M;                      // sample count (capacity when m_dev is given)
  const int32_t* m_dev;
};

// slab layout: dW1 | dW2 | dW3 (128 x 128 each) | dW0_pe (128 x 64) | db0..db3 (128 each)
#define TCW_SLAB (3 * 128 * 128 + 128 * 64 + 4 * 128)
__host__ __device__ constexpr int tcw_slab_off(int layer) { return layer == 0 ? 3 * 128 * 128 : (layer - 1) * 128 * 128; }
#define TCW_SLAB_BIAS (3 * 128 * 128 + 128 * 64)

#define TCW_STAGES 3
#define TCW_HALF 8192            // bytes of a 64-row half of one [128 x 64] fp16 tile
#define TCW_STAGE_BYTES (8 * TCW_HALF)
#define TCW_THREADS 192          // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
#define TCW_TMEM_COLS 512
// accumulator columns: dW1 0, dW2 128, dW3 256, dW0_pe 384 (64), db_l 448 + 16 l (N = 16 is the smallest M=128 shape)
__host__ __device__ constexpr uint32_t tcw_acc_col(int layer) { return layer == 0 ? 384u : (uint32_t)(layer - 1) * 128u; }

struct TcWSmem {
  static constexpr int OFF_STAGE = 0;
  static constexpr int OFF_ONES = TCW_STAGES * TCW_STAGE_BYTES;   // 2 KiB of fp16 1.0
  static constexpr int OFF_BAR = OFF_ONES + 2048;
  static constexpr int OFF_TMEM = OFF_BAR + (2 * TCW_STAGES + 1) * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
};

__global__ void __launch_bounds__(TCW_THREADS, 1) tc_wgrad_kernel(const TcWgradParams p) {
  using S = TcWSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sStage = smem + S::OFF_STAGE;      // stage s: [dY: kc0 hi | kc0 lo | kc1 hi | kc1 lo][X: same], 8 KiB each
  uint8_t* sOnes = smem + S::OFF_ONES;
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + TCW_STAGES;
  uint64_t* done = bars + 2 * TCW_STAGES;
  uint32_t* sTmem = (uint32_t*)(smem + S::OFF_TMEM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (apn_rt_count(p.m_dev, p.M) + TC_SAMPLES - 1) / TC_SAMPLES;
  if ((int)blockIdx.x >= n_tiles) return;      // no tile for this CTA (device-side count below the grid's capacity): its slab is not summed
  if (tid == 0) {
    for (int i = 0; i < TCW_STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 1024; i += TCW_THREADS) reinterpret_cast<__half*>(sOnes)[i] = __float2half_rn(1.f);
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<TCW_TMEM_COLS>(sTmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sTmem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint8_t* tp = p.tape + (size_t)tile * TC_TAPE_TILE_BYTES;
        const uint8_t* dyt = p.dy + (size_t)tile * TC_DY_TILE_BYTES;
        for (int layer = 0; layer < 4; ++layer)
          for (int half = 0; half < 2; ++half, ++it) {
            const uint32_t s = it % TCW_STAGES;
            mbar_wait(empty + s, ((it / TCW_STAGES) & 1) ^ 1);
            uint8_t* st = sStage + (size_t)s * TCW_STAGE_BYTES;
            const int nx = layer == 0 ? 2 : 4;                       // X pieces: PE has one K chunk
            mbar_arrive_expect_tx(full + s, (uint32_t)((4 + nx) * TCW_HALF));
#pragma unroll
            for (int pc = 0; pc < 4; ++pc)                           // pc = kc * 2 + split
              bulk_g2s(st + pc * TCW_HALF, dyt + TC_TAPE_ACT(layer, pc >> 1, pc & 1) + half * TCW_HALF, TCW_HALF, full + s);
            for (int pc = 0; pc < nx; ++pc) {
              const uint8_t* src = layer == 0 ? tp + TC_TAPE_PE(pc) : tp + TC_TAPE_ACT(layer - 1, pc >> 1, pc & 1);
              bulk_g2s(st + (4 + pc) * TCW_HALF, src + half * TCW_HALF, TCW_HALF, full + s);
            }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t ones = smem_u32(sOnes);
      bool first_tile = true;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first_tile = false) {
        for (int layer = 0; layer < 4; ++layer) {
          // A = dY^T (M = 128 out features), B = X^T (N = in features); both MN-major over the [row][col] tiles
          const uint32_t idesc = umma_idesc_f16_major(128, layer == 0 ? 64 : 128, true, true);
          const uint32_t idesc_b = umma_idesc_f16_major(128, 16, true, true);
          const uint32_t acc = tmem_base + tcw_acc_col(layer), acc_b = tmem_base + 448u + 16u * layer;
          for (int half = 0; half < 2; ++half, ++it) {
            const uint32_t s = it % TCW_STAGES;
            mbar_wait(full + s, (it / TCW_STAGES) & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sStage + (size_t)s * TCW_STAGE_BYTES), b0 = a0 + 4 * TCW_HALF;
            // the 64-column atoms of one operand are 2 pieces apart (kc0 hi, kc0 lo, kc1 hi, kc1 lo)
            const uint32_t lbo_a = 2 * TCW_HALF, lbo_b = 2 * TCW_HALF;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc_flag = (first_tile && half == 0 && ks == 0) ? 0u : 1u;
              const uint64_t da = umma_desc_mn_sw128(a0 + ks * 2048, lbo_a), da_lo = umma_desc_mn_sw128(a0 + TCW_HALF + ks * 2048, lbo_a);
              const uint64_t db = umma_desc_mn_sw128(b0 + ks * 2048, lbo_b), db_lo = umma_desc_mn_sw128(b0 + TCW_HALF + ks * 2048, lbo_b);
              umma_f16(acc, da, db, idesc, acc_flag);
              umma_f16(acc, da, db_lo, idesc, 1u);
              umma_f16(acc, da_lo, db, idesc, 1u);
              const uint64_t d1 = umma_desc_mn_sw128(ones, 1024);
              umma_f16(acc_b, da, d1, idesc_b, acc_flag);
              umma_f16(acc_b, da_lo, d1, idesc_b, 1u);
            }
            umma_commit(empty + s);
          }
        }
      }
      umma_commit(done);
    }
  } else {
    // ================================================================= epilogue: accumulators -> this CTA's slab
    // (148 CTAs adding into the same 57k addresses would serialise in L2; a second kernel sums the slabs)
    mbar_wait(done, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int out = q * 32 + lane;                       // output feature = accumulator row
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    float* slab = p.partial + (size_t)blockIdx.x * TCW_SLAB;
    for (int layer = 0; layer < 4; ++layer) {
      const int n_in = layer == 0 ? 64 : 128;
      float* dst = slab + tcw_slab_off(layer) + (size_t)out * n_in;
      for (int c0 = 0; c0 < n_in; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + tcw_acc_col(layer) + tlane + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                 __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
      }
      uint32_t vb[16];
      tmem_ld16(tmem_base + 448u + tlane + 16u * layer, vb);
      tmem_ld_wait();
      slab[TCW_SLAB_BIAS + layer * 128 + out] = __uint_as_float(vb[0]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCW_TMEM_COLS>(tmem_base);
}

// sums the per-CTA slabs, removes the gradient scale and accumulates into the (caller-zeroed) gradient buffers
__global__ void __launch_bounds__(256)
tc_wgrad_reduce_kernel(const float* __restrict__ partial, int grid_slabs, int M_cap, const int32_t* __restrict__ m_dev,
                       const float* __restrict__ hmax, int d_in, float* dw0, float* dw1, float* dw2, float* dw3, float* db0,
                       float* db1, float* db2, float* db3) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= TCW_SLAB) return;
  const int n_slabs = min(grid_slabs, (apn_rt_count(m_dev, M_cap) + TC_SAMPLES - 1) / TC_SAMPLES);   // CTAs that had a tile
  float s = 0.f;
  for (int b = 0; b < n_slabs; ++b) s += __ldg(partial + (size_t)b * TCW_SLAB + e);
  s *= 1.f / tc_grad_scale(hmax);
  if (e < 3 * 128 * 128) {
    float* dw = e < 128 * 128 ? dw1 : e < 2 * 128 * 128 ? dw2 : dw3;
    dw[e & (128 * 128 - 1)] += s;
  } else if (e < TCW_SLAB_BIAS) {
    const int r = (e - 3 * 128 * 128) >> 6, c = (e - 3 * 128 * 128) & 63;
    const int col = tc_pe_ref_col(c);
    if (col >= 0) dw0[(size_t)r * d_in + col] += s;
  } else {
    const int l = (e - TCW_SLAB_BIAS) >> 7, n = (e - TCW_SLAB_BIAS) & 127;
    float* db = l == 0 ? db0 : l == 1 ? db1 : l == 2 ? db2 : db3;
    db[n] += s;
  }
}

// Pose embedding (lib/temporalpoints.py:483-490,571-576: the same 64-vector e is appended to EVERY decoder input row).
// The forward folds  W0[:, 191:] . e  into the layer-0 bias, so its backward needs no per-row work at all:
//   d_e       = W0[:, 191:]^T . db0          db0 = sum over rows of dY0 = this launch's layer-0 bias gradient
//   dW0[:, 191:] += db0 (x) e
// One block; thread n owns output feature n.  db0 is re-summed from the per-CTA slabs (g->d_b[0] may already hold
// gradient from an earlier accumulation).
__global__ void __launch_bounds__(128)
tc_pose_bwd_kernel(const float* __restrict__ partial, int grid_slabs, int M_cap, const int32_t* __restrict__ m_dev,
                   const float* __restrict__ hmax, int d_in, const float* __restrict__ w0, const float* __restrict__ pose_emb,
                   float* __restrict__ dw0, float* __restrict__ d_pose_emb) {
  __shared__ float sDb[128];
  const int n_slabs = min(grid_slabs, (apn_rt_count(m_dev, M_cap) + TC_SAMPLES - 1) / TC_SAMPLES);
  const int n = threadIdx.x;
  const int n_pose = d_in - (APN_PE_POS + APN_C);
  float s = 0.f;
  for (int b = 0; b < n_slabs; ++b) s += __ldg(partial + (size_t)b * TCW_SLAB + TCW_SLAB_BIAS + n);
  s *= 1.f / tc_grad_scale(hmax);
  sDb[n] = s;
  float* dwr = dw0 + (size_t)n * d_in + APN_PE_POS + APN_C;
  for (int j = 0; j < n_pose; ++j) dwr[j] += s * __ldg(pose_emb + j);
  __syncthreads();
  if (d_pose_emb)
    for (int j = n; j < n_pose; j += 128) {
      float a = 0.f;
      for (int r = 0; r < 128; ++r) a = fmaf(__ldg(w0 + (size_t)r * d_in + APN_PE_POS + APN_C + j), sDb[r], a);
      d_pose_emb[j] += a;
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// d density / Raw2Alpha backward on the reduced feature: d_h += dd * w_density (lib/cuda/render_utils_kernel.cu:396-406)
__global__ void __launch_bounds__(256)
tc_density_bwd_kernel(int M_cap, const int32_t* __restrict__ m_dev, float interval, const float* __restrict__ h,
                      const float* __restrict__ exp_d, const float* __restrict__ density_w, const float* __restrict__ d_alpha,
                      float* __restrict__ d_h, float* __restrict__ d_density_w, float* __restrict__ d_density_b,
                      float* __restrict__ hmax) {
  __shared__ float sAcc[APN_C + 1];
  const int M = apn_rt_count(m_dev, M_cap);
  for (int i = threadIdx.x; i < APN_C + 1; i += blockDim.x) sAcc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const float4 wd = __ldg(reinterpret_cast<const float4*>(density_w) + lane);
  float4 a_dw = make_float4(0.f, 0.f, 0.f, 0.f);
  float a_db = 0.f, gmax = 0.f;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    const float e = __ldg(exp_d + m);
    const float dd = (float)(fmin((double)e, 1e10) * (double)powf(1.f + e, -interval - 1.f) * (double)interval * (double)__ldg(d_alpha + m));
    const float4 hv = __ldg(reinterpret_cast<const float4*>(h + (size_t)m * APN_C) + lane);
    a_dw.x += dd * hv.x; a_dw.y += dd * hv.y; a_dw.z += dd * hv.z; a_dw.w += dd * hv.w;
    a_db += dd;
    float4* dh = reinterpret_cast<float4*>(d_h + (size_t)m * APN_C) + lane;
    float4 g = *dh;
    g.x += dd * wd.x; g.y += dd * wd.y; g.z += dd * wd.z; g.w += dd * wd.w;
    *dh = g;
    gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(g.x), fabsf(g.y)), fmaxf(fabsf(g.z), fabsf(g.w))));
  }
  gmax = warp_max(gmax);
  if (lane == 0 && isfinite(gmax)) atomicMax(reinterpret_cast<unsigned int*>(hmax), __float_as_uint(gmax));
  atomicAdd(&sAcc[4 * lane], a_dw.x);
  atomicAdd(&sAcc[4 * lane + 1], a_dw.y);
  atomicAdd(&sAcc[4 * lane + 2], a_dw.z);
  atomicAdd(&sAcc[4 * lane + 3], a_dw.w);
  if (lane == 0) atomicAdd(&sAcc[APN_C], a_db);
  __syncthreads();
  for (int i = threadIdx.x; i < APN_C; i += blockDim.x) atomicAdd(d_density_w + i, sAcc[i]);
  if (threadIdx.x == 0) atomicAdd(d_density_b, sAcc[APN_C]);
}

struct TcBwdScratch {
  uint8_t* dy;
  float *d_v0, *d_fv, *d_h, *d_ptable, *hmax, *partial;
  size_t total;
};
static TcBwdScratch tc_bwd_layout(char* base, int M, int N) {
  TcBwdScratch b;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    char* p = base + o;
    o = apn_align(o + bytes, 1024);
    return p;
  };
  b.dy = (uint8_t*)take((size_t)apn_div_up(M, TC_SAMPLES) * TC_DY_TILE_BYTES);
  b.d_v0 = (float*)take((size_t)M * 64 * 4);
  b.d_fv = (float*)take((size_t)M * 160 * 4);
  b.d_h = (float*)take((size_t)M * APN_C * 4);
  b.d_ptable = (float*)take((size_t)N * APN_C * 4 + 1024);   // + the max |d_h| scalar right behind the table
  b.hmax = b.d_ptable + (size_t)N * APN_C;
  b.partial = (float*)take((size_t)APN_SM_COUNT * TCW_SLAB * 4);
  b.total = o;
  return b;
}
extern "C" size_t apn_aggregate_tc_bwd_scratch_bytes(int M, int N) { return (M > 0 && N > 0) ? tc_bwd_layout(nullptr, M, N).total : 0; }

// phase 0: the whole backward.  phase 1: everything up to the gradient of the point features (heads, density, dgrad chain,
// d_feat = dP W0_feat) — canonical_feat.grad, the bulk of a data-parallel gradient exchange, is final when it returns.
// phase 2: the rest (weight gradients of feat_net, the point-table weight gradient, the pose-embedding gradient) from the
// SAME scratch buffer.  1 followed by 2 == 0.
static int aggregate_bwd_tc_impl(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_bwd,
                                 const apn_agg_outputs* sv, const void* tape, const apn_agg_grads* g, void* scratch,
                                 size_t scratch_bytes, int phase, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  // phase 0: everything.  1 | 2: through canonical_feat.grad | the rest (data-parallel step: the point-feature gradient is
  // exchanged while the weight gradients are computed).  3 | 4: through tc_dgrad (d_xyz, d_ginv final: the LBS / pose
  // backward can start) | every parameter gradient (point table, feat_net) — the branches of the one-GPU graph; 4 = 5 (d_feat
  // only, on the caller's stream) + 6 (the weight gradients), which are independent of each other.
  APN_CHECK_ARG(phase >= 0 && phase <= 6, "phase must be 0..6");
  const bool do1 = phase == 0 || phase == 1 || phase == 3, do2 = phase == 0 || phase == 2 || phase == 4 || phase == 6;
  const bool do_feat = phase == 0 || phase == 1 || phase == 4 || phase == 5;
  APN_CHECK_ARG(in && w && packed_bwd && sv && tape && g, "null pointer");
  APN_CHECK_ARG(in->d_in == APN_PE_POS + APN_C || (in->d_in > APN_PE_POS + APN_C && in->d_in <= 256 && in->pose_emb),
                "d_in must be 191, or 192..256 with a pose embedding");
  const int M = in->M, N = in->N;
  if (M <= 0) return 0;
  APN_CHECK_ARG(sv->rgb && sv->idw && sv->h && sv->exp_d && sv->fv && sv->v0, "saved forward tensors missing");
  APN_CHECK_ARG(g->d_alpha && g->d_rgb, "incoming gradients missing");
  for (int l = 0; l < 4; ++l) APN_CHECK_ARG(g->d_w[l] && g->d_b[l], "null feat_net gradient buffer");
  APN_CHECK_ARG(g->d_density_w && g->d_density_b && g->d_rgb_feat_w && g->d_rgb_feat_b && g->d_rgb_v0_w && g->d_rgb_v0_b &&
                    g->d_rgb_v2_w && g->d_rgb_v2_b, "null head gradient buffer");
  APN_CHECK_ARG(scratch && scratch_bytes >= apn_aggregate_tc_bwd_scratch_bytes(M, N) && (((uintptr_t)scratch) & 1023) == 0,
                "scratch too small or not 1 KiB aligned");
  const TcBwdScratch b = tc_bwd_layout((char*)scratch, M, N);
  // heads: RGBNet backward -> d_h, then densitynet / Raw2Alpha
  // small batches: the weight-gradient GEMMs of the heads and of the point table run on the library's side streams,
  // beside the serial chain (each fills less than half the GPU); large batches fill it and stay on one stream
  ApnSide* side = nullptr;
  if (M <= 32768 && apn_side_streams(&side)) return -2;
  const int n_tiles = apn_div_up(M, TC_SAMPLES);
  const int grid = n_tiles < APN_SM_COUNT ? n_tiles : APN_SM_COUNT;
  if (do1) {
    if (agg_rgbnet_bwd_launch(st, in, w, sv, g, b.d_v0, b.d_fv, b.d_h, side, true)) return -1;
    const int wblocks = min(apn_div_up(M, 8), APN_SM_COUNT * 8);
    APN_CUDA(cudaMemsetAsync(b.d_ptable, 0, (size_t)N * APN_C * sizeof(float) + 1024, st));   // table + hmax
    tc_density_bwd_kernel<<<wblocks, 256, 0, st>>>(M, in->m_dev, in->interval, sv->h, sv->exp_d, w->density_w, g->d_alpha, b.d_h,
                                                   g->d_density_w, g->d_density_b, b.hmax);
    APN_LAUNCH_CHECK();
    TcBwdParams p;
    p.in = *in;
    p.d_h = b.d_h;
    p.idw = sv->idw;
    p.tape = (const uint8_t*)tape;
    p.packed = (const uint8_t*)packed_bwd;
    p.dy = b.dy;
    p.d_ptable = b.d_ptable;
    p.d_xyz = g->d_xyz;
    p.d_ginv = g->d_ginv;
    p.hmax = b.hmax;
    p.n_tiles = n_tiles;
    static_assert(TcBwdSmem::TOTAL <= 227 * 1024, "shared memory budget");
    APN_CUDA(cudaFuncSetAttribute(tc_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcBwdSmem::TOTAL));
    tc_dgrad_kernel<<<grid, TC_THREADS, TcBwdSmem::TOTAL, st>>>(p);
    APN_LAUNCH_CHECK();
  }
  cudaStream_t sw = st;
  if (side && do2) {                            // the point-table GEMMs only need tc_dgrad's d_ptable
    APN_CUDA(cudaEventRecord(side->fork[2], st));
    APN_CUDA(cudaStreamWaitEvent(side->s[0], side->fork[2], 0));
    sw = side->s[0];
  }
  // feature columns of layer 0 through the per-point table: d_feat = dP W0_feat, dW0_feat += dP^T feat
  if (do_feat && g->d_feat)                     // phase 1 keeps it on the caller's stream: it is what the caller waits for
    APN_CHECK_ARG(tgemm_dgrad_accum(sw, b.d_ptable, APN_C, w->w[0] + APN_PE_POS, in->d_in, g->d_feat, APN_C, N, APN_C, APN_C) == 0,
                  "dgrad point table");      // accumulates, like every other gradient of this entry point
  if (do2) {
    APN_CHECK_ARG(tgemm_wgrad(sw, b.d_ptable, APN_C, in->feat, APN_C, g->d_w[0] + APN_PE_POS, in->d_in, N, APN_C, APN_C) == 0,
                  "wgrad point table");
    TcWgradParams p;
    p.tape = (const uint8_t*)tape;
    p.dy = b.dy;
    p.partial = b.partial;
    p.M = M;
    p.m_dev = in->m_dev;
    static_assert(TcWSmem::TOTAL <= 227 * 1024, "shared memory budget");
    APN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcWSmem::TOTAL));
    tc_wgrad_kernel<<<grid, TCW_THREADS, TcWSmem::TOTAL, st>>>(p);
    APN_LAUNCH_CHECK();
    tc_wgrad_reduce_kernel<<<apn_div_up(TCW_SLAB, 256), 256, 0, st>>>(b.partial, grid, M, in->m_dev, b.hmax, in->d_in, g->d_w[0],
                                                                       g->d_w[1], g->d_w[2], g->d_w[3], g->d_b[0], g->d_b[1],
                                                                       g->d_b[2], g->d_b[3]);
    APN_LAUNCH_CHECK();
    if (in->d_in > APN_PE_POS + APN_C) {
      tc_pose_bwd_kernel<<<1, 128, 0, st>>>(b.partial, grid, M, in->m_dev, b.hmax, in->d_in, w->w[0], in->pose_emb, g->d_w[0],
                                            g->d_pose_emb);
      APN_LAUNCH_CHECK();
    }
  }
  if (side) {                                   // join: everything this call launched is ordered before what follows on st
    // only the side streams THIS call forked (under stream capture a wait on an event of a stream that is not part of the
    // capture is an error): phase 1 / 0 fork both in the heads backward, phase 2 only the point-table stream
    for (int i = 0; i < (do1 ? 2 : (do2 ? 1 : 0)); ++i) {
      APN_CUDA(cudaEventRecord(side->join[i], side->s[i]));
      APN_CUDA(cudaStreamWaitEvent(st, side->join[i], 0));
    }
  }
  return 0;
}

extern "C" int apn_aggregate_bwd_tc(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_bwd,
                                    const apn_agg_outputs* sv, const void* tape, const apn_agg_grads* g, void* scratch,
                                    size_t scratch_bytes, apn_stream_t stream_) {
  return aggregate_bwd_tc_impl(in, w, packed_bwd, sv, tape, g, scratch, scratch_bytes, 0, stream_);
}

extern "C" int apn_aggregate_bwd_tc_phase(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_bwd,
                                          const apn_agg_outputs* sv, const void* tape, const apn_agg_grads* g, void* scratch,
                                          size_t scratch_bytes, int phase, apn_stream_t stream_) {
  return aggregate_bwd_tc_impl(in, w, packed_bwd, sv, tape, g, scratch, scratch_bytes, phase, stream_);
}
