// K3 (tensor-core path) — the decoder MLP on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators
// in tensor memory), fused with everything that feeds it and the neighbour reduction that follows it:
//   gather of the 8 neighbours -> canonical-frame offset -> positional encoding -> feat_net (4 x Linear +
//   LeakyReLU) -> inverse-distance reduce over the neighbours  (+ the direct branch and the idw weights).
// Replaces lib/temporalpoints.py:446-494 and lib/tineuvox.py:872-878; the heads (densitynet / Raw2Alpha /
// RGBNet, 4.5 % of the flops) run on the reduced feature through agg_heads_launch (aggregate.cu).
//
// Layer 0 is split algebraically: the feature columns of its input are a pure gather of per-point rows, so
//   W0 [PE | feat[idx]] = W0_pe PE + (feat W0_feat^T)[idx]
// and the second term is a per-point table P (N x 128 fp32, exact CUDA-core GEMM, rebuilt only when the
// features or W0 change: apn_aggregate_tc_point_table).  The kernel adds P[idx] in the layer-0 epilogue; the
// tensor cores only see K = 64 for layer 0 instead of 192.
//
// One persistent CTA per SM walks 128-row tiles (16 kept samples x 8 neighbours):
//   warps 0-15 "compute": build the PE operand tile (fp16, K-major, 128-byte swizzle); after every layer read
//              the fp32 accumulator from tensor memory, add bias (+P row), LeakyReLU, write the next layer's
//              operand tile — K-chunk 0 first, then K-chunk 1, each announced on its own mbarrier so the next
//              layer's MMAs start on chunk 0 while chunk 1 is still being written (accumulators are double
//              buffered in tensor memory); after layer 3 the weighted 8-row reduce produces h.
//              The next tile's prologue runs between the layer-2 epilogue and the final epilogue, i.e. under
//              the layer-3 MMAs, and the next tile's layer-0 MMAs run under the final epilogue.
//   warp 16    streams the packed weight chunks ([128 out x 64 in] fp16 tiles, pre-swizzled by
//              apn_aggregate_tc_pack_weights) with 1-D bulk async copies into a ring of shared-memory slots
//              (precision 0: all 7 chunks stay resident);
//   warp 17    one elected thread issues tcgen05.mma (M=128, N=128, K=16) and commits to mbarriers.
// Precision 0: fp16 operands, fp32 accumulate (1 MMA per K step).
// Precision 1: every operand is split x = hi + lo (two fp16 terms, 22 mantissa bits) and the three products
//              hi*hi + hi*lo + lo*hi are accumulated in fp32: fp32-class results (the parity mode).
// PE chunk columns: [rel_c(3) sin(30) cos(30) 0] (the reference's own order).  A pose embedding (d_in = 255) is
// constant over rows: W0[:,191:255] * pose is folded into the layer-0 bias by every CTA at start-up.
#include "sgemm.cuh"
#include "aggregate_tc.cuh"

struct TcParams {
  apn_agg_inputs in;
  const float* bias[4];
  const float* w0;            // fp32 W0 (128, d_in) for the pose-embedding fold
  const uint8_t* packed;
  const float* ptable;        // (N,128) feat W0_feat^T
  float* h;                   // (M,128)
  float* idw;                 // (M,8)
  float* alpha_direct;        // (M) or NULL
  float* rgb_direct;          // (M,3) or NULL
  uint8_t* tape;              // training: n_tiles * TC_TAPE_TILE_BYTES, NULL in inference
  int n_tiles;
};

// ---------------------------------------------------------------------------------------
// weight packing: the shared-memory image of every chunk
// ---------------------------------------------------------------------------------------
__global__ void tc_pack_kernel(const apn_mlp_weights w, int d_in, uint8_t* __restrict__ packed) {
  const int c = blockIdx.x;                                  // chunk
  const int layer = c == 0 ? 0 : 1 + (c - 1) / 2;
  const int kc = c == 0 ? 0 : (c - 1) % 2;
  const float* W = w.w[layer];
  const int ld = layer == 0 ? d_in : APN_C;
  for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
    const int n = e >> 6, k = e & 63;
    const int col = layer == 0 ? (k < APN_PE_POS ? k : -1) : kc * 64 + k;
    const float v = col >= 0 ? W[(size_t)n * ld + col] : 0.f;
    __half hi, lo;
    split_half(v, hi, lo);
    uint8_t* base = packed + (size_t)c * TC_CHUNK_GBYTES;
    const uint32_t o = sw128_offset(n, k);
    *reinterpret_cast<__half*>(base + o) = hi;
    *reinterpret_cast<__half*>(base + TC_TILE_BYTES + o) = lo;
  }
}

extern "C" size_t apn_aggregate_tc_weights_bytes(int d_in) {
  (void)d_in;
  return (size_t)TC_NCHUNKS * TC_CHUNK_GBYTES;
}

extern "C" int apn_aggregate_tc_pack_weights(const apn_mlp_weights* w, int d_in, void* packed, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  APN_CHECK_ARG(w && packed, "null pointer");
  APN_CHECK_ARG(d_in >= APN_PE_POS + APN_C && d_in <= 256, "d_in must be 191..256");
  APN_CHECK_ARG((((uintptr_t)packed) & 15) == 0, "packed weights must be 16-byte aligned");
  for (int l = 0; l < 4; ++l) APN_CHECK_ARG(w->w[l], "null feat_net weight");
  tc_pack_kernel<<<TC_NCHUNKS, 256, 0, st>>>(*w, d_in, (uint8_t*)packed);
  APN_LAUNCH_CHECK();
  return 0;
}

// P = feat W0[:, 63:191]^T  (N x 128), exact fp32
extern "C" int apn_aggregate_tc_point_table(const float* feat, const float* w0, int d_in, int N, float* ptable,
                                            apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  APN_CHECK_ARG(feat && w0 && ptable, "null pointer");
  APN_CHECK_ARG(d_in >= APN_PE_POS + APN_C && d_in <= 256 && N > 0, "bad sizes");
  APN_CHECK_ARG(gemm_forward(st, feat, APN_C, w0 + APN_PE_POS, d_in, nullptr, ptable, APN_C, N, APN_C, APN_C, 1.f) == 0,
                "point table gemm");
  return 0;
}

// ---------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------
template <int NSPLIT>
struct TcSmem {
  static constexpr int NSLOT = NSPLIT == 1 ? TC_NCHUNKS : 3;
  static constexpr int OFF_PE = 0;                                           // NSPLIT tiles
  static constexpr int OFF_ACT = OFF_PE + NSPLIT * TC_TILE_BYTES;            // 2 x NSPLIT tiles: (kc, split)
  static constexpr int OFF_W = OFF_ACT + 2 * NSPLIT * TC_TILE_BYTES;         // NSLOT x NSPLIT tiles
  static constexpr int OFF_BIAS = OFF_W + NSLOT * NSPLIT * TC_TILE_BYTES;    // 4 x 128 floats
  static constexpr int OFF_IDW = OFF_BIAS + 4 * 128 * 4;                     // 2 x 128 floats
  static constexpr int OFF_IDX = OFF_IDW + 2 * 128 * 4;                      // 2 x 128 ints
  static constexpr int OFF_BAR = OFF_IDX + 2 * 128 * 4;
  static constexpr int N_BAR = 2 * NSLOT + 5;
  static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;                         // + slack for the 1024-byte alignment
};

// Builds the per-tile sample state (idw, neighbour indices, direct branch) and the PE operand tile.
// 4 threads per row: part p owns the (dimension, frequency) pairs j = p, p+4, ...
template <int NSPLIT>
__device__ __forceinline__ void tc_prologue(const TcParams& p, int tile, int tid, uint8_t* sPE, float* sIdw, int* sIdx) {

  const apn_agg_inputs& in = p.in;
  const int r = tid & 127, part = tid >> 7;
  const int s = r >> 3;
  const int m0 = tile * TC_SAMPLES;
  const int m = min(m0 + s, in.M - 1);
  const bool valid = (m0 + s) < in.M;
  const int idx = __ldg(in.nn_idx + (size_t)m * APN_K + (r & 7));
  const float px = __ldg(in.pts + 3 * (size_t)m), py = __ldg(in.pts + 3 * (size_t)m + 1), pz = __ldg(in.pts + 3 * (size_t)m + 2);
  const float rx = px - __ldg(in.xyz + 3 * (size_t)idx), ry = py - __ldg(in.xyz + 3 * (size_t)idx + 1),
              rz = pz - __ldg(in.xyz + 3 * (size_t)idx + 2);
  if (part == 0) {
    const float d2 = (rx * rx + ry * ry) + rz * rz;
    // inverse-distance weights (lib/temporalpoints.py:473-475): the 8 rows of a sample are 8 consecutive lanes
    const float u = 1.0f / (d2 + in.eps);
    float su = u;
    su += __shfl_xor_sync(0xffffffffu, su, 1);
    su += __shfl_xor_sync(0xffffffffu, su, 2);
    su += __shfl_xor_sync(0xffffffffu, su, 4);
    const float w = u / su;
    sIdw[r] = w;
    sIdx[r] = idx;
    if (valid) p.idw[(size_t)m * APN_K + (r & 7)] = w;
    if (p.alpha_direct) {
      // direct branch (lib/temporalpoints.py:459-470)
      const float sig = in.mean_min_distance * fmaxf(__ldg(in.direct_eps + idx), 0.f);
      const float wd = expf(-(d2 * d2) / (2.f * sig * sig + 1e-12f));
      float sw = wd;
      sw += __shfl_xor_sync(0xffffffffu, sw, 1);
      sw += __shfl_xor_sync(0xffffffffu, sw, 2);
      sw += __shfl_xor_sync(0xffffffffu, sw, 4);
      const float wn = wd / (sw + 1e-12f);
      float a = (1.0f / APN_K) * wd * fminf(fmaxf(__ldg(in.canonical_alpha + idx), 0.f), 1.f);
      float cr = wn * fminf(fmaxf(__ldg(in.canonical_rgbs + 3 * (size_t)idx), 0.f), 1.f);
      float cg = wn * fminf(fmaxf(__ldg(in.canonical_rgbs + 3 * (size_t)idx + 1), 0.f), 1.f);
      float cb = wn * fminf(fmaxf(__ldg(in.canonical_rgbs + 3 * (size_t)idx + 2), 0.f), 1.f);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        cr += __shfl_xor_sync(0xffffffffu, cr, o);
        cg += __shfl_xor_sync(0xffffffffu, cg, o);
        cb += __shfl_xor_sync(0xffffffffu, cb, o);
      }
      if (valid && (r & 7) == 0) {
        p.alpha_direct[m] = a;
        p.rgb_direct[3 * (size_t)m] = cr; p.rgb_direct[3 * (size_t)m + 1] = cg; p.rgb_direct[3 * (size_t)m + 2] = cb;
      }
    }
  }
  // canonical-frame offset (lib/temporalpoints.py:478-480)
  const float* G = in.ginv + 9 * (size_t)idx;
  float rc[3];
  rc[0] = __ldg(G) * rx + __ldg(G + 1) * ry + __ldg(G + 2) * rz;
  rc[1] = __ldg(G + 3) * rx + __ldg(G + 4) * ry + __ldg(G + 5) * rz;
  rc[2] = __ldg(G + 6) * rx + __ldg(G + 7) * ry + __ldg(G + 8) * rz;
  auto put = [&](int col, float v) {
    __half hi, lo;
    split_half(v, hi, lo);
    const uint32_t o = sw128_offset(r, col);
    *reinterpret_cast<__half*>(sPE + o) = hi;
    if (NSPLIT == 2) *reinterpret_cast<__half*>(sPE + TC_TILE_BYTES + o) = lo;
  };
  if (part < 3) put(part, part == 0 ? rc[0] : part == 1 ? rc[1] : rc[2]);
  else put(63, 0.f);
  // poc_fre (lib/tineuvox.py:872-878): column 3 + d*10 + i = sin(rel_c[d] * 2^i), column 33 + d*10 + i = cos(...)
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = part + 4 * jj;
    if (j < 30) {
      const int d = j / 10, i = j - d * 10;
      float sn, cs;
      sincosf((d == 0 ? rc[0] : d == 1 ? rc[1] : rc[2]) * (float)(1 << i), &sn, &cs);
      put(3 + j, sn);
      put(33 + j, cs);
    }
  }
}

template <int NSPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) agg_tc_fwd_kernel(const TcParams p) {
  using S = TcSmem<NSPLIT>;
  constexpr int NSLOT = S::NSLOT;
  constexpr bool RESIDENT = (NSPLIT == 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sPE = smem + S::OFF_PE;               // [hi tile][lo tile]
  uint8_t* sAct = smem + S::OFF_ACT;             // tile (kc, split) at (kc * NSPLIT + split) * TC_TILE_BYTES
  uint8_t* sW = smem + S::OFF_W;                 // slot s: [hi tile][lo tile]
  float* sBias = (float*)(smem + S::OFF_BIAS);
  float* sIdw = (float*)(smem + S::OFF_IDW);     // [2][128]
  int* sIdx = (int*)(smem + S::OFF_IDX);         // [2][128]
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);
  uint64_t* w_full = bars;                        // [NSLOT] weights landed
  uint64_t* w_free = bars + NSLOT;                // [NSLOT] MMAs reading the slot have completed
  uint64_t* pe_ready = bars + 2 * NSLOT;          // PE tile written (all compute threads)
  uint64_t* a_ready = bars + 2 * NSLOT + 1;       // [2] activation K-chunk written (all compute threads)
  uint64_t* acc_ready = bars + 2 * NSLOT + 3;     // [2] accumulator buffer complete
  uint32_t* sTmem = (uint32_t*)(smem + S::OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const apn_agg_inputs& in = p.in;

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(w_full + i, 1);
      mbar_init(w_free + i, 1);
    }
    mbar_init(pe_ready, TC_COMPUTE_THREADS);
    mbar_init(a_ready, TC_COMPUTE_THREADS);
    mbar_init(a_ready + 1, TC_COMPUTE_THREADS);
    mbar_init(acc_ready, 1);
    mbar_init(acc_ready + 1, 1);
    fence_barrier_init();
  }
  if (warp == TC_COMPUTE_WARPS + 1) tmem_alloc<TC_TMEM_COLS>(sTmem);
  // biases (+ the pose-embedding fold into the layer-0 bias)
  for (int i = tid; i < 4 * 128; i += TC_THREADS) {
    const int l = i >> 7, n = i & 127;
    float b = __ldg(p.bias[l] + n);
    if (l == 0 && in.d_in > APN_PE_POS + APN_C) {
      const float* wr = p.w0 + (size_t)n * in.d_in + APN_PE_POS + APN_C;
      for (int j = 0; j < in.d_in - APN_PE_POS - APN_C; ++j) b = fmaf(__ldg(wr + j), __ldg(in.pose_emb + j), b);
    }
    sBias[i] = b;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sTmem;

  if (warp == TC_COMPUTE_WARPS) {
    // ================================================================= weight producer
    if (lane == 0) {
      if (RESIDENT) {
        for (int c = 0; c < TC_NCHUNKS; ++c) {
          mbar_arrive_expect_tx(w_full + c, TC_TILE_BYTES);
          bulk_g2s(sW + (size_t)c * TC_TILE_BYTES, p.packed + (size_t)c * TC_CHUNK_GBYTES, TC_TILE_BYTES, w_full + c);
        }
      } else {
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
          for (int c = 0; c < TC_NCHUNKS; ++c, ++it) {
            const uint32_t slot = it % NSLOT;
            mbar_wait(w_free + slot, ((it / NSLOT) & 1) ^ 1);
            mbar_arrive_expect_tx(w_full + slot, NSPLIT * TC_TILE_BYTES);
            bulk_g2s(sW + (size_t)slot * NSPLIT * TC_TILE_BYTES, p.packed + (size_t)c * TC_CHUNK_GBYTES, NSPLIT * TC_TILE_BYTES,
                     w_full + slot);
          }
        }
      }
    }
  } else if (warp == TC_COMPUTE_WARPS + 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 128);
      const uint32_t pe_base = smem_u32(sPE), act_base = smem_u32(sAct), w_base = smem_u32(sW);
      uint32_t it = 0, ph_pe = 0, ph_a0 = 0, ph_a1 = 0;
      // one K chunk of 64: 4 K steps, NSPLIT == 2 adds the two cross products
      auto issue_chunk = [&](uint32_t a_hi, uint32_t acc, bool first_chunk, int c) {
        const uint32_t slot = RESIDENT ? (uint32_t)c : it % NSLOT;
        mbar_wait(w_full + slot, RESIDENT ? 0u : ((it / NSLOT) & 1));
        tc_fence_after();
        const uint32_t b_hi = w_base + slot * (uint32_t)(NSPLIT * TC_TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da = umma_desc_k_sw128(a_hi + ks * 32), db = umma_desc_k_sw128(b_hi + ks * 32);
          umma_f16(acc, da, db, idesc, (first_chunk && ks == 0) ? 0u : 1u);
          if (NSPLIT == 2) {
            const uint64_t da_lo = umma_desc_k_sw128(a_hi + TC_TILE_BYTES + ks * 32);
            const uint64_t db_lo = umma_desc_k_sw128(b_hi + TC_TILE_BYTES + ks * 32);
            umma_f16(acc, da, db_lo, idesc, 1u);
            umma_f16(acc, da_lo, db, idesc, 1u);
          }
        }
        if (!RESIDENT) umma_commit(w_free + slot);
        ++it;
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        // layer 0: PE chunk -> accumulator 0
        mbar_wait(pe_ready, ph_pe);
        ph_pe ^= 1;
        tc_fence_after();
        // training: the operand tiles ARE the tape (same layout): stream them out with bulk stores from shared memory
        uint8_t* tape_tile = p.tape ? p.tape + (size_t)tile * TC_TAPE_TILE_BYTES : nullptr;
        if (NSPLIT == 2 && tape_tile) {
          bulk_s2g(tape_tile + TC_TAPE_PE(0), sPE, 2 * TC_TILE_BYTES);
          bulk_commit();
        }
        issue_chunk(pe_base, tmem_base, true, 0);
        umma_commit(acc_ready);
        // layers 1..3: activation chunks 0, 1 -> accumulator (layer & 1)
        for (int layer = 1; layer < 4; ++layer) {
          const uint32_t acc = tmem_base + (uint32_t)((layer & 1) * 128);
          mbar_wait(a_ready, ph_a0);
          ph_a0 ^= 1;
          tc_fence_after();
          if (NSPLIT == 2 && tape_tile) {
            bulk_s2g(tape_tile + TC_TAPE_ACT(layer - 1, 0, 0), sAct, 2 * TC_TILE_BYTES);
            bulk_commit();
          }
          issue_chunk(act_base, acc, true, 2 * layer - 1);
          mbar_wait(a_ready + 1, ph_a1);
          ph_a1 ^= 1;
          tc_fence_after();
          if (NSPLIT == 2 && tape_tile) {
            bulk_s2g(tape_tile + TC_TAPE_ACT(layer - 1, 1, 0), sAct + 2 * TC_TILE_BYTES, 2 * TC_TILE_BYTES);
            bulk_commit();
            bulk_wait_read();      // the epilogue released by the commit below overwrites these tiles (and, later, sPE)
          }
          issue_chunk(act_base + NSPLIT * TC_TILE_BYTES, acc, false, 2 * layer);
          umma_commit(acc_ready + (layer & 1));
        }
      }
      if (NSPLIT == 2 && p.tape) bulk_wait_all();
    }
  } else {
    // ================================================================= compute warps
    uint32_t ph_acc0 = 0, ph_acc1 = 0;
    const int q = warp & 3, cq = warp >> 2;
    const int erow = q * 32 + lane;                       // accumulator row (tensor-memory lane) owned in the epilogues
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    int n = 0;
    if ((int)blockIdx.x < p.n_tiles) {
      tc_prologue<NSPLIT>(p, blockIdx.x, tid, sPE, sIdw, sIdx);
      fence_proxy_async_smem();
      mbar_arrive(pe_ready);
      compute_sync();
    }
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++n) {
      const int buf = n & 1;
      const int m0 = tile * TC_SAMPLES;
      // ---------------------------------------------------------------- epilogues of layers 0..2
#pragma unroll
      for (int layer = 0; layer < 3; ++layer) {
        float4 pf[2][4];
        if (layer == 0) {
          // P[idx] rows for this thread's 2 x 16 columns, in flight while the layer-0 MMAs finish
          const float* prow = p.ptable + (size_t)sIdx[buf * 128 + erow] * APN_C + cq * 16;
#pragma unroll
          for (int ph = 0; ph < 2; ++ph)
#pragma unroll
            for (int u = 0; u < 4; ++u) pf[ph][u] = __ldg(reinterpret_cast<const float4*>(prow + ph * 64) + u);
        }
        if (layer & 1) {
          mbar_wait(acc_ready + 1, ph_acc1);
          ph_acc1 ^= 1;
        } else {
          mbar_wait(acc_ready, ph_acc0);
          ph_acc0 ^= 1;
        }
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)((layer & 1) * 128) + tlane;
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {                  // K chunk `ph` of the next layer's operand
          uint32_t v[16];
          tmem_ld16(tacc + ph * 64 + cq * 16, v);
          tmem_ld_wait();
          const float* bias = sBias + layer * 128 + ph * 64 + cq * 16;
          uint8_t* t_hi = sAct + (size_t)(ph * NSPLIT) * TC_TILE_BYTES + (size_t)erow * 128;
          uint8_t* tp = p.tape ? p.tape + (size_t)tile * TC_TAPE_TILE_BYTES : nullptr;
          uint32_t mbits = 0;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float y0 = __uint_as_float(v[u * 8 + 2 * e]) + bias[u * 8 + 2 * e];
              float y1 = __uint_as_float(v[u * 8 + 2 * e + 1]) + bias[u * 8 + 2 * e + 1];
              if (layer == 0) {
                const float4 f = pf[ph][u * 2 + (e >> 1)];
                y0 += (e & 1) ? f.z : f.x;
                y1 += (e & 1) ? f.w : f.y;
              }
              mbits |= (y0 > 0.f ? 1u : 0u) << (u * 8 + 2 * e);
              mbits |= (y1 > 0.f ? 1u : 0u) << (u * 8 + 2 * e + 1);
              split_half2(leaky(y0), leaky(y1), hi[e], lo[e]);
            }
            const uint32_t o = (uint32_t)((((cq * 2 + u) ^ (erow & 7)) & 7) << 4);
            *reinterpret_cast<uint4*>(t_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (NSPLIT == 2) *reinterpret_cast<uint4*>(t_hi + TC_TILE_BYTES + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          if (tp) *reinterpret_cast<uint16_t*>(tp + TC_TAPE_MASK(layer) + ((size_t)erow * 8 + ph * 4 + cq) * 2) = (uint16_t)mbits;
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(a_ready + ph);
        }
      }
      // ---------------------------------------------------------------- next tile's prologue (under the layer-3 MMAs)
      const int next = tile + gridDim.x;
      if (next < p.n_tiles) {
        tc_prologue<NSPLIT>(p, next, tid, sPE, sIdw + (buf ^ 1) * 128, sIdx + (buf ^ 1) * 128);
        fence_proxy_async_smem();
        mbar_arrive(pe_ready);
      }
      // ---------------------------------------------------------------- layer 3: out_k = LeakyReLU(acc + b3); h = sum_k idw_k out_k
      mbar_wait(acc_ready + 1, ph_acc1);
      ph_acc1 ^= 1;
      tc_fence_after();
      {
        float v[32];
        {
          uint32_t t0[16], t1[16];
          const uint32_t tacc = tmem_base + 128u + tlane;
          tmem_ld16(tacc + cq * 16, t0);
          tmem_ld16(tacc + 64 + cq * 16, t1);
          tmem_ld_wait();
          const float w = sIdw[buf * 128 + erow];
          const float* bias = sBias + 3 * 128 + cq * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = leaky(__uint_as_float(t0[i]) + bias[i]);
            v[16 + i] = leaky(__uint_as_float(t1[i]) + bias[64 + i]);
          }
          if (p.tape) {
            uint8_t* tp = p.tape + (size_t)tile * TC_TAPE_TILE_BYTES + (size_t)erow * 128;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph)
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_half2(v[ph * 16 + u * 8 + 2 * e], v[ph * 16 + u * 8 + 2 * e + 1], hi[e], lo[e]);
                uint8_t* g = tp + TC_TAPE_ACT(3, ph, 0) + (uint32_t)((((cq * 2 + u) ^ (erow & 7)) & 7) << 4);
                *reinterpret_cast<uint4*>(g) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(g + TC_TILE_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= w;
        }
        tc_fence_before();
        // reduce-scatter over the 8 lanes (rows) of a sample: 32 -> 16 -> 8 -> 4 columns per lane
        const bool b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
        float x16[16], x8[8], x4[4];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float send = b4 ? v[i] : v[16 + i];
          const float keep = b4 ? v[16 + i] : v[i];
          x16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float send = b2 ? x16[i] : x16[8 + i];
          const float keep = b2 ? x16[8 + i] : x16[i];
          x8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = b1 ? x8[i] : x8[4 + i];
          const float keep = b1 ? x8[4 + i] : x8[i];
          x4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
        const int m = m0 + (erow >> 3);
        if (m < in.M) {
          const int col = (b4 ? 64 : 0) + cq * 16 + (b2 ? 8 : 0) + (b1 ? 4 : 0);
          *reinterpret_cast<float4*>(p.h + (size_t)m * APN_C + col) = make_float4(x4[0], x4[1], x4[2], x4[3]);
        }
      }
      compute_sync();     // next tile's sIdx / sIdw visible to every compute thread; this tile's are free again
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_COMPUTE_WARPS + 1) tmem_dealloc<TC_TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct TcScratch {
  float *h, *exp_d, *fv, *v0;
  size_t total;
};
static TcScratch tc_scratch_layout(char* base, int M) {
  TcScratch b;
  size_t o = 0;
  auto take = [&](size_t n_float) {
    float* p = (float*)(base + o);
    o = apn_align(o + n_float * sizeof(float));
    return p;
  };
  b.h = take((size_t)M * APN_C);
  b.exp_d = take(M);
  b.fv = take((size_t)M * 160);
  b.v0 = take((size_t)M * 64);
  b.total = o;
  return b;
}
extern "C" size_t apn_aggregate_tc_scratch_bytes(int M) { return M > 0 ? tc_scratch_layout(nullptr, M).total : 0; }

template <int NSPLIT>
static int tc_launch(cudaStream_t st, const TcParams& p) {
  using S = TcSmem<NSPLIT>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory budget");
  APN_CUDA(cudaFuncSetAttribute(agg_tc_fwd_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const int grid = p.n_tiles < APN_SM_COUNT ? p.n_tiles : APN_SM_COUNT;
  agg_tc_fwd_kernel<NSPLIT><<<grid, TC_THREADS, S::TOTAL, st>>>(p);
  APN_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t apn_aggregate_tc_tape_bytes(int M) {
  return M > 0 ? (size_t)apn_div_up(M, TC_SAMPLES) * TC_TAPE_TILE_BYTES : 0;
}

extern "C" int apn_aggregate_fwd_tc(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_weights,
                                    const float* point_table, const apn_agg_outputs* out, int precision, void* tape,
                                    size_t tape_bytes, void* scratch, size_t scratch_bytes, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  APN_CHECK_ARG(in && w && out && packed_weights && point_table, "null pointer");
  APN_CHECK_ARG(precision == 0 || precision == 1, "precision: 0 = fp16 operands, 1 = split fp16 (fp32-class)");
  APN_CHECK_ARG(in->d_in == APN_PE_POS + APN_C || (in->d_in > APN_PE_POS + APN_C && in->d_in <= 256 && in->pose_emb),
                "d_in must be 191, or 192..256 with a pose embedding");
  APN_CHECK_ARG(in->pts && in->nn_idx && in->ray_id && in->xyz && in->ginv && in->viewdirs, "null input pointer");
  APN_CHECK_ARG(out->alpha && out->rgb && out->idw, "alpha, rgb and idw outputs are required");
  APN_CHECK_ARG((out->alpha_direct == nullptr) == (out->rgb_direct == nullptr), "direct outputs come as a pair");
  APN_CHECK_ARG(!out->alpha_direct || (in->canonical_alpha && in->canonical_rgbs && in->direct_eps), "direct branch inputs missing");
  APN_CHECK_ARG((((uintptr_t)point_table) & 15) == 0, "point table must be 16-byte aligned");
  const int M = in->M;
  if (M <= 0) return 0;
  TcScratch b;
  if (tape) {
    // training: the split-precision kernel records the tape; the heads' intermediates are kept by the caller
    APN_CHECK_ARG(precision == 1, "the training forward runs in split precision");
    APN_CHECK_ARG(tape_bytes >= apn_aggregate_tc_tape_bytes(M) && (((uintptr_t)tape) & 1023) == 0, "tape too small or not 1 KiB aligned");
    APN_CHECK_ARG(out->h && out->exp_d && out->fv && out->v0, "training needs the h / exp_d / fv / v0 buffers");
    b.h = out->h; b.exp_d = out->exp_d; b.fv = out->fv; b.v0 = out->v0;
  } else {
    APN_CHECK_ARG(scratch && scratch_bytes >= apn_aggregate_tc_scratch_bytes(M), "scratch too small");
    b = tc_scratch_layout((char*)scratch, M);
  }
  TcParams p;
  p.in = *in;
  for (int l = 0; l < 4; ++l) p.bias[l] = w->b[l];
  p.w0 = w->w[0];
  p.packed = (const uint8_t*)packed_weights;
  p.ptable = point_table;
  p.h = b.h;
  p.idw = out->idw;
  p.alpha_direct = out->alpha_direct;
  p.rgb_direct = out->rgb_direct;
  p.tape = (uint8_t*)tape;
  p.n_tiles = apn_div_up(M, TC_SAMPLES);
  const int rc = precision == 0 ? tc_launch<1>(st, p) : tc_launch<2>(st, p);
  if (rc) return rc;
  return agg_heads_launch(st, in, w, nullptr, out->idw, b.h, b.exp_d, out->alpha, b.fv, b.v0, out->rgb);
}
