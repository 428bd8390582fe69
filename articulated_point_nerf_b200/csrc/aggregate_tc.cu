// K3 (tensor-core path) — the decoder MLP on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators
// in tensor memory), fused with everything that feeds it and the neighbour reduction that follows it:
//   gather of the 8 neighbours -> canonical-frame offset -> positional encoding -> feat_net (4 x Linear +
//   LeakyReLU) -> inverse-distance reduce over the neighbours  (+ the direct branch and the idw weights).
// Replaces lib/temporalpoints.py:446-494 and lib/tineuvox.py:872-878; the heads (densitynet / Raw2Alpha /
// RGBNet, 4.5 % of the flops) run on the reduced feature through agg_heads_launch (aggregate.cu).
//
// One persistent CTA per SM walks 128-row tiles (16 kept samples x 8 neighbours):
//   warps 0-7  build the layer-0 operand tile in shared memory (fp16, K-major, 128-byte swizzle), and after every
//              layer read the fp32 accumulator from tensor memory, apply bias + LeakyReLU and write the next
//              layer's operand tile (layers 0-2) or do the weighted 8-row reduce and store h (layer 3);
//   warp 8     streams the packed weight chunks ([128 out x 64 in] fp16 tiles, pre-swizzled by
//              apn_aggregate_tc_pack_weights) with 1-D bulk async copies into a ring of shared-memory slots
//              (precision 0: all 9 chunks stay resident);
//   warp 9     one elected thread issues tcgen05.mma (M=128, N=128, K=16) and commits to mbarriers.
// Precision 0: fp16 operands, fp32 accumulate (1 MMA per K step).
// Precision 1: every operand is split x = hi + lo (two fp16 terms, 22 mantissa bits) and the three products
//              hi*hi + hi*lo + lo*hi are accumulated in fp32: fp32-class results (the parity mode).
// Column order of the layer-0 operand: [feat 0..63 | feat 64..127 | rel_c(3) sin(30) cos(30) 0]; the packed
// W0 uses the same permutation.  A pose embedding (d_in = 255) is constant over rows: W0[:,191:255] * pose is
// folded into the layer-0 bias by every CTA at start-up.
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

#define TC_ROWS 128
#define TC_SAMPLES 16
#define TC_NCHUNKS 9                 // K chunks of 64: layer 0 has 3, layers 1-3 have 2
#define TC_TILE_BYTES 16384          // [128 x 64] fp16
#define TC_CHUNK_GBYTES (2 * TC_TILE_BYTES)   // packed global: hi tile then lo tile
#define TC_COMPUTE_WARPS 8
#define TC_COMPUTE_THREADS (32 * TC_COMPUTE_WARPS)
#define TC_THREADS (TC_COMPUTE_THREADS + 64)
#define TC_TMEM_COLS 128

struct TcParams {
  apn_agg_inputs in;
  const float* bias[4];
  const float* w0;            // fp32 W0 (128, d_in) for the pose-embedding fold
  const uint8_t* packed;
  float* h;                   // (M,128)
  float* idw;                 // (M,8)
  float* alpha_direct;        // (M) or NULL
  float* rgb_direct;          // (M,3) or NULL
  int n_tiles;
};

// ---------------------------------------------------------------------------------------
// weight packing: the shared-memory image of every chunk
// ---------------------------------------------------------------------------------------
__global__ void tc_pack_kernel(const apn_mlp_weights w, int d_in, uint8_t* __restrict__ packed) {
  const int c = blockIdx.x;                                  // chunk
  const int layer = c < 3 ? 0 : 1 + (c - 3) / 2;
  const int kc = c < 3 ? c : (c - 3) % 2;
  const float* W = w.w[layer];
  const int ld = layer == 0 ? d_in : APN_C;
  for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
    const int n = e >> 6, k = e & 63;
    int col;
    if (layer == 0) col = (kc < 2) ? APN_PE_POS + kc * 64 + k : (k < APN_PE_POS ? k : -1);
    else col = kc * 64 + k;
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

// ---------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------
template <int NSPLIT>
struct TcSmem {
  static constexpr int NSLOT = NSPLIT == 1 ? TC_NCHUNKS : 3;
  static constexpr int A_BYTES = 3 * NSPLIT * TC_TILE_BYTES;
  static constexpr int W_BYTES = NSLOT * NSPLIT * TC_TILE_BYTES;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = A_BYTES;
  static constexpr int OFF_BIAS = OFF_W + W_BYTES;          // 4 x 128 floats
  static constexpr int OFF_IDW = OFF_BIAS + 4 * 128 * 4;     // 128 floats
  static constexpr int OFF_IDX = OFF_IDW + 128 * 4;          // 128 ints
  static constexpr int OFF_BAR = OFF_IDX + 128 * 4;          // barriers
  static constexpr int N_BAR = 2 * NSLOT + 2;
  static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;         // + slack for the 1024-byte alignment of the base
};

__device__ __forceinline__ float leaky(float y) { return y < 0.f ? y * 0.01f : y; }

// bar.sync among the compute warps only
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE_THREADS) : "memory"); }

template <int NSPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) agg_tc_fwd_kernel(const TcParams p) {
  using S = TcSmem<NSPLIT>;
  constexpr int NSLOT = S::NSLOT;
  constexpr bool RESIDENT = (NSPLIT == 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem + S::OFF_A;                 // tile (kc, split) at (kc * NSPLIT + split) * TC_TILE_BYTES
  uint8_t* sW = smem + S::OFF_W;                 // slot s: [hi tile][lo tile]
  float* sBias = (float*)(smem + S::OFF_BIAS);
  float* sIdw = (float*)(smem + S::OFF_IDW);
  int* sIdx = (int*)(smem + S::OFF_IDX);
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);
  uint64_t* w_full = bars;                        // [NSLOT] weights landed
  uint64_t* w_free = bars + NSLOT;                // [NSLOT] MMAs reading the slot have completed
  uint64_t* a_ready = bars + 2 * NSLOT;           // operand tile written (TC_COMPUTE_THREADS arrivals)
  uint64_t* acc_ready = bars + 2 * NSLOT + 1;     // accumulator of the layer complete
  uint32_t* sTmem = (uint32_t*)(smem + S::OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const apn_agg_inputs& in = p.in;

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(w_full + i, 1);
      mbar_init(w_free + i, 1);
    }
    mbar_init(a_ready, TC_COMPUTE_THREADS);
    mbar_init(acc_ready, 1);
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
      const uint32_t a_base = smem_u32(sA), w_base = smem_u32(sW);
      uint32_t it = 0, ph_a = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int c = 0;
        for (int layer = 0; layer < 4; ++layer) {
          mbar_wait(a_ready, ph_a);
          ph_a ^= 1;
          tc_fence_after();
          const int nk = layer == 0 ? 3 : 2;
          for (int kc = 0; kc < nk; ++kc, ++c, ++it) {
            const uint32_t slot = RESIDENT ? (uint32_t)c : it % NSLOT;
            mbar_wait(w_full + slot, RESIDENT ? 0u : ((it / NSLOT) & 1));
            tc_fence_after();
            const uint32_t a_hi = a_base + (uint32_t)(kc * NSPLIT) * TC_TILE_BYTES;
            const uint32_t b_hi = w_base + slot * (uint32_t)(NSPLIT * TC_TILE_BYTES);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = umma_desc_k_sw128(a_hi + ks * 32), db = umma_desc_k_sw128(b_hi + ks * 32);
              umma_f16(tmem_base, da, db, idesc, (kc | ks) ? 1u : 0u);
              if (NSPLIT == 2) {
                const uint64_t da_lo = umma_desc_k_sw128(a_hi + TC_TILE_BYTES + ks * 32);
                const uint64_t db_lo = umma_desc_k_sw128(b_hi + TC_TILE_BYTES + ks * 32);
                umma_f16(tmem_base, da, db_lo, idesc, 1u);
                umma_f16(tmem_base, da_lo, db, idesc, 1u);
              }
            }
            if (!RESIDENT) umma_commit(w_free + slot);
          }
          umma_commit(acc_ready);
        }
      }
    }
  } else {
    // ================================================================= compute warps
    uint32_t ph_acc = 0;
    const int q = warp & 3, chalf = warp >> 2;
    const int erow = q * 32 + lane;                       // accumulator row owned in the epilogues
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int m0 = tile * TC_SAMPLES;
      // ---------------------------------------------------------------- prologue 1: geometry + PE (2 threads per row)
      {
        const int r = tid & 127, half = tid >> 7;
        const int s = r >> 3;
        const int m = min(m0 + s, in.M - 1);
        const bool valid = (m0 + s) < in.M;
        const int idx = __ldg(in.nn_idx + (size_t)m * APN_K + (r & 7));
        const float px = __ldg(in.pts + 3 * (size_t)m), py = __ldg(in.pts + 3 * (size_t)m + 1), pz = __ldg(in.pts + 3 * (size_t)m + 2);
        const float rx = px - __ldg(in.xyz + 3 * (size_t)idx), ry = py - __ldg(in.xyz + 3 * (size_t)idx + 1),
                    rz = pz - __ldg(in.xyz + 3 * (size_t)idx + 2);
        const float d2 = (rx * rx + ry * ry) + rz * rz;
        if (half == 0) {
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
        // canonical-frame offset and its positional encoding -> chunk 2 of the operand tile
        const float* G = in.ginv + 9 * (size_t)idx;
        float rc[3];
        rc[0] = __ldg(G) * rx + __ldg(G + 1) * ry + __ldg(G + 2) * rz;
        rc[1] = __ldg(G + 3) * rx + __ldg(G + 4) * ry + __ldg(G + 5) * rz;
        rc[2] = __ldg(G + 6) * rx + __ldg(G + 7) * ry + __ldg(G + 8) * rz;
        uint8_t* t_hi = sA + (size_t)(2 * NSPLIT) * TC_TILE_BYTES;
        uint8_t* t_lo = t_hi + TC_TILE_BYTES;
        auto put = [&](int col, float v) {
          __half hi, lo;
          split_half(v, hi, lo);
          const uint32_t o = sw128_offset(r, col);
          *reinterpret_cast<__half*>(t_hi + o) = hi;
          if (NSPLIT == 2) *reinterpret_cast<__half*>(t_lo + o) = lo;
        };
        if (half == 0) {
          put(0, rc[0]); put(1, rc[1]); put(2, rc[2]);
          put(63, 0.f);
        }
        // poc_fre: column 3 + d*10 + i = sin(rel_c[d] * 2^i), column 33 + d*10 + i = cos(...); half h owns i = 5h..5h+4
#pragma unroll
        for (int d = 0; d < 3; ++d) {
#pragma unroll
          for (int ii = 0; ii < 5; ++ii) {
            const int i = half * 5 + ii;
            float sn, cs;
            sincosf(rc[d] * (float)(1 << i), &sn, &cs);
            put(3 + d * 10 + i, sn);
            put(33 + d * 10 + i, cs);
          }
        }
      }
      compute_sync();     // sIdx visible
      // ---------------------------------------------------------------- prologue 2: feature gather (one warp per row)
      {
#pragma unroll 4
        for (int j = 0; j < TC_ROWS / TC_COMPUTE_WARPS; ++j) {
          const int r = warp * (TC_ROWS / TC_COMPUTE_WARPS) + j;
          const float4 f = __ldg(reinterpret_cast<const float4*>(in.feat + (size_t)sIdx[r] * APN_C) + lane);
          __half h0, l0, h1, l1, h2, l2, h3, l3;
          split_half(f.x, h0, l0); split_half(f.y, h1, l1); split_half(f.z, h2, l2); split_half(f.w, h3, l3);
          const int kc = lane >> 4, unit = (lane & 15) >> 1;
          const uint32_t o = (uint32_t)(r * 128 + (((unit ^ (r & 7)) & 7) << 4) + ((lane & 1) << 3));
          uint8_t* t_hi = sA + (size_t)(kc * NSPLIT) * TC_TILE_BYTES;
          *reinterpret_cast<uint2*>(t_hi + o) = make_uint2(pack_half2(h0, h1), pack_half2(h2, h3));
          if (NSPLIT == 2) *reinterpret_cast<uint2*>(t_hi + TC_TILE_BYTES + o) = make_uint2(pack_half2(l0, l1), pack_half2(l2, l3));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(a_ready);
      // ---------------------------------------------------------------- layers
      for (int layer = 0; layer < 4; ++layer) {
        mbar_wait(acc_ready, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(chalf * 64);
        const float* bias = sBias + layer * 128 + chalf * 64;
        if (layer < 3) {
          uint8_t* t_hi = sA + (size_t)(chalf * NSPLIT) * TC_TILE_BYTES + (size_t)erow * 128;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld32(taddr + hh * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint32_t ph[4], pl[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int cidx = hh * 32 + u * 8 + 2 * e;
                const float y0 = leaky(__uint_as_float(v[u * 8 + 2 * e]) + bias[cidx]);
                const float y1 = leaky(__uint_as_float(v[u * 8 + 2 * e + 1]) + bias[cidx + 1]);
                __half a0, b0, a1, b1;
                split_half(y0, a0, b0);
                split_half(y1, a1, b1);
                ph[e] = pack_half2(a0, a1);
                pl[e] = pack_half2(b0, b1);
              }
              const uint32_t o = (uint32_t)((((hh * 4 + u) ^ (erow & 7)) & 7) << 4);
              *reinterpret_cast<uint4*>(t_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
              if (NSPLIT == 2) *reinterpret_cast<uint4*>(t_hi + TC_TILE_BYTES + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(a_ready);
        } else {
          // out_k = LeakyReLU(acc + b3); h = sum_k idw_k out_k: reduce-scatter over the 8 lanes of a sample
          float v[64];
          {
            uint32_t t0[32], t1[32];
            tmem_ld32(taddr, t0);
            tmem_ld32(taddr + 32, t1);
            tmem_ld_wait();
            const float w = sIdw[erow];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = w * leaky(__uint_as_float(t0[i]) + bias[i]);
              v[32 + i] = w * leaky(__uint_as_float(t1[i]) + bias[32 + i]);
            }
          }
          tc_fence_before();
          const bool b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
          float x32[32], x16[16], x8[8];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float send = b4 ? v[i] : v[32 + i];
            const float keep = b4 ? v[32 + i] : v[i];
            x32[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float send = b2 ? x32[i] : x32[16 + i];
            const float keep = b2 ? x32[16 + i] : x32[i];
            x16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float send = b1 ? x16[i] : x16[8 + i];
            const float keep = b1 ? x16[8 + i] : x16[i];
            x8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
          const int s = erow >> 3, m = m0 + s;
          if (m < in.M) {
            const int col = chalf * 64 + (b4 ? 32 : 0) + (b2 ? 16 : 0) + (b1 ? 8 : 0);
            float4* dst = reinterpret_cast<float4*>(p.h + (size_t)m * APN_C + col);
            dst[0] = make_float4(x8[0], x8[1], x8[2], x8[3]);
            dst[1] = make_float4(x8[4], x8[5], x8[6], x8[7]);
          }
        }
      }
      compute_sync();     // sIdw / sIdx are rewritten by the next tile's prologue
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

extern "C" int apn_aggregate_fwd_tc(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_weights,
                                    const apn_agg_outputs* out, int precision, void* scratch, size_t scratch_bytes,
                                    apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  APN_CHECK_ARG(in && w && out && packed_weights, "null pointer");
  APN_CHECK_ARG(precision == 0 || precision == 1, "precision: 0 = fp16 operands, 1 = split fp16 (fp32-class)");
  APN_CHECK_ARG(in->d_in == APN_PE_POS + APN_C || (in->d_in > APN_PE_POS + APN_C && in->d_in <= 256 && in->pose_emb),
                "d_in must be 191, or 192..256 with a pose embedding");
  APN_CHECK_ARG(in->pts && in->nn_idx && in->ray_id && in->xyz && in->ginv && in->feat && in->viewdirs, "null input pointer");
  APN_CHECK_ARG(out->alpha && out->rgb && out->idw, "alpha, rgb and idw outputs are required");
  APN_CHECK_ARG((out->alpha_direct == nullptr) == (out->rgb_direct == nullptr), "direct outputs come as a pair");
  APN_CHECK_ARG(!out->alpha_direct || (in->canonical_alpha && in->canonical_rgbs && in->direct_eps), "direct branch inputs missing");
  APN_CHECK_ARG((((uintptr_t)in->feat) & 15) == 0, "feat must be 16-byte aligned");
  const int M = in->M;
  if (M <= 0) return 0;
  APN_CHECK_ARG(scratch && scratch_bytes >= apn_aggregate_tc_scratch_bytes(M), "scratch too small");
  const TcScratch b = tc_scratch_layout((char*)scratch, M);
  TcParams p;
  p.in = *in;
  for (int l = 0; l < 4; ++l) p.bias[l] = w->b[l];
  p.w0 = w->w[0];
  p.packed = (const uint8_t*)packed_weights;
  p.h = b.h;
  p.idw = out->idw;
  p.alpha_direct = out->alpha_direct;
  p.rgb_direct = out->rgb_direct;
  p.n_tiles = apn_div_up(M, TC_SAMPLES);
  const int rc = precision == 0 ? tc_launch<1>(st, p) : tc_launch<2>(st, p);
  if (rc) return rc;
  return agg_heads_launch(st, in, w, nullptr, out->idw, b.h, b.exp_d, out->alpha, b.fv, b.v0, out->rgb);
}
