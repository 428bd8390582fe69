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
#include "tgemm.cuh"
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
    const int col = layer == 0 ? tc_pe_ref_col(k) : kc * 64 + k;
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
  APN_CHECK_ARG(tgemm_forward(st, feat, APN_C, w0 + APN_PE_POS, d_in, nullptr, ptable, APN_C, N, APN_C, APN_C, 1.f) == 0,
                "point table gemm");
  return 0;
}

// ---------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------
#ifdef TC_TRACE
// scratch builds only (-DTC_TRACE): CTA 0 records (tag, clock64) pairs per role into a device buffer
__device__ long long g_tc_trace[4][4096];
__device__ int g_tc_trace_n[4];
extern "C" void* apn_tc_trace_buffer(int which) {
  void* p = nullptr;
  if (which == 0) cudaGetSymbolAddress(&p, g_tc_trace); else cudaGetSymbolAddress(&p, g_tc_trace_n);
  return p;
}
#define TRACE_DECL int trace_k__ = 0
#define TRACE(role, tag)                                                        \
  do {                                                                          \
    if (blockIdx.x == 0 && trace_k__ < 2047) {                                  \
      g_tc_trace[role][2 * trace_k__] = (tag);                                  \
      g_tc_trace[role][2 * trace_k__ + 1] = clock64();                          \
      g_tc_trace_n[role] = ++trace_k__;                                         \
    }                                                                           \
  } while (0)
#else
#define TRACE_DECL do {} while (0)
#define TRACE(role, tag) do {} while (0)
#endif

template <int NSPLIT>
struct TcSmem {
  static constexpr int NSLOT = NSPLIT == 1 ? TC_NCHUNKS : 3;
  static constexpr int ACT_BYTES = 2 * NSPLIT * TC_TILE_BYTES;               // per group: (kc, split) tiles; PE aliases kc = 0
  static constexpr int OFF_ACT = 0;                                          // 2 groups
  static constexpr int OFF_W = OFF_ACT + 2 * ACT_BYTES;                      // NSLOT x NSPLIT tiles
  static constexpr int OFF_BIAS = OFF_W + NSLOT * NSPLIT * TC_TILE_BYTES;    // 4 x 128 floats
  static constexpr int OFF_BAR = OFF_BIAS + 4 * 128 * 4;
  static constexpr int N_BAR = 2 * NSLOT + 10;
  static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;                                // the dynamic window itself is 1024-byte aligned
};

// Per-row sample state and the PE operand values of one tile row, kept in registers.
// Two threads share a row: half 0 owns PE columns 0..31 (dimensions 0, 1), half 1 columns 32..63.
struct TcRow {
  int idx;          // neighbour index of this row
  float w;          // inverse-distance weight
};

// stage 1 of the next tile's prologue: the loads every later gather depends on
struct TcRowLoad {
  int idx;
  float px, py, pz;
};
__device__ __forceinline__ TcRowLoad tc_row_load(const apn_agg_inputs& in, int M, int tile, int erow) {
  TcRowLoad l;
  const int m = min(tile * TC_SAMPLES + (erow >> 3), M - 1);
  l.idx = __ldg(in.nn_idx + (size_t)m * APN_K + (erow & 7));
  l.px = __ldg(in.pts + 3 * (size_t)m);
  l.py = __ldg(in.pts + 3 * (size_t)m + 1);
  l.pz = __ldg(in.pts + 3 * (size_t)m + 2);
  return l;
}

// stage 2: geometry, inverse-distance weights, direct branch and the 32 PE columns of this thread as packed
// half2 (hi[16], lo[16]: four 16-byte units each)
template <int NSPLIT>
__device__ __forceinline__ TcRow tc_prologue(const TcParams& p, int M, int tile, int erow, int half, const TcRowLoad& l,
                                             uint32_t (&hi)[16], uint32_t (&lo)[16]) {
  const apn_agg_inputs& in = p.in;
  const int m0 = tile * TC_SAMPLES, s = erow >> 3;
  const int m = min(m0 + s, M - 1);
  const bool valid = (m0 + s) < M;
  const int idx = l.idx;
  const float rx = l.px - __ldg(in.xyz + 3 * (size_t)idx), ry = l.py - __ldg(in.xyz + 3 * (size_t)idx + 1),
              rz = l.pz - __ldg(in.xyz + 3 * (size_t)idx + 2);
  const float* G = in.ginv + 9 * (size_t)idx;
  float g[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = __ldg(G + i);
  const float d2 = (rx * rx + ry * ry) + rz * rz;
  // inverse-distance weights (lib/temporalpoints.py:473-475): the 8 rows of a sample are 8 consecutive lanes
  const float u = 1.0f / (d2 + in.eps);
  float su = u;
  su += __shfl_xor_sync(0xffffffffu, su, 1);
  su += __shfl_xor_sync(0xffffffffu, su, 2);
  su += __shfl_xor_sync(0xffffffffu, su, 4);
  TcRow row;
  row.idx = idx;
  row.w = u / su;
  if (half == 0) {
    if (valid) p.idw[(size_t)m * APN_K + (erow & 7)] = row.w;
  } else if (p.alpha_direct) {
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
    if (valid && (erow & 7) == 0) {
      p.alpha_direct[m] = a;
      p.rgb_direct[3 * (size_t)m] = cr; p.rgb_direct[3 * (size_t)m + 1] = cg; p.rgb_direct[3 * (size_t)m + 2] = cb;
    }
  }
  // canonical-frame offset (lib/temporalpoints.py:478-480)
  const float rc0 = g[0] * rx + g[1] * ry + g[2] * rz;
  const float rc1 = g[3] * rx + g[4] * ry + g[5] * rz;
  const float rc2 = g[6] * rx + g[7] * ry + g[8] * rz;
  // poc_fre (lib/tineuvox.py:872-878) in the tile's column layout (aggregate_tc.cuh)
  auto dim16 = [&](float x, uint32_t* h8, uint32_t* l8) {          // sin i = 0..7 | cos i = 0..7
    float sn[8], cs[8];
    tc_pe_octaves<0, 8>(x, sn, cs);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      split_half2(sn[2 * e], sn[2 * e + 1], h8[e], l8[e]);
      split_half2(cs[2 * e], cs[2 * e + 1], h8[4 + e], l8[4 + e]);
    }
  };
  if (half == 0) {
    dim16(rc0, hi, lo);
    dim16(rc1, hi + 8, lo + 8);
  } else {
    dim16(rc2, hi, lo);
#pragma unroll
    for (int d = 0; d < 3; ++d) {                                   // [sin 8, sin 9 | cos 8, cos 9]
      float sn[2], cs[2];
      tc_pe_octaves<8, 2>(d == 0 ? rc0 : d == 1 ? rc1 : rc2, sn, cs);
      split_half2(sn[0], sn[1], hi[8 + 2 * d], lo[8 + 2 * d]);
      split_half2(cs[0], cs[1], hi[9 + 2 * d], lo[9 + 2 * d]);
    }
    split_half2(rc0, rc1, hi[14], lo[14]);
    split_half2(rc2, 0.f, hi[15], lo[15]);
  }
  return row;
}

// the thread's four 16-byte units of the PE tile (K chunk 0 of the group's operand buffer)
template <int NSPLIT>
__device__ __forceinline__ void tc_store_pe(uint32_t act, int erow, int half, const uint32_t (&hi)[16], const uint32_t (&lo)[16]) {
  const uint32_t t = act + (uint32_t)erow * 128u;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t o = (uint32_t)((((half * 4 + u) ^ (erow & 7)) & 7) << 4);
    sts128(t + o, hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
    if (NSPLIT == 2) sts128(t + TC_TILE_BYTES + o, lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
  }
}

// Layer-0 table rows P[idx] -> accumulator 0 of the group (the layer-0 MMAs then accumulate on top of them).
// A row-per-thread gather costs one L1 tag look-up per lane and instruction (32 distinct lines); instead the warp
// reads its 32 rows line by line (WB bytes of a row per group of lanes), transposes through a private scratch
// (the K-chunk-1 operand tiles, idle between the layer-3 MMAs and the layer-0 epilogue) and stores row-per-lane
// registers to tensor memory.  Warp (q, half) covers rows 32 q .. 32 q + 31, columns 64 half .. 64 half + 63.
template <int NSPLIT>
__device__ __forceinline__ void tc_stage_ptable(const float* __restrict__ ptable, int idx, uint32_t scratch, uint32_t tacc,
                                                int half, int lane) {
  constexpr int WB = 64 * NSPLIT;             // bytes of a row per pass (scratch: 32 rows x WB per warp)
  constexpr int U = WB / 16;                  // 16-byte units per row and pass
  constexpr int RPI = 32 / U;                 // rows per load instruction
  constexpr int NPASS = 256 / WB;
  const int lu = lane % U, lr = lane / U;
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    const int col0 = half * 64 + pass * (WB / 4);
    float4 ld[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int r = j * RPI + lr;
      const int ridx = __shfl_sync(0xffffffffu, idx, r);
      ld[j] = __ldg(reinterpret_cast<const float4*>(ptable + (size_t)ridx * APN_C + col0) + lu);
    }
    __syncwarp();                             // the previous pass has been read out of the scratch
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int r = j * RPI + lr;
      const int sw = NSPLIT == 2 ? (r & 7) : ((r >> 1) & 3);
      sts128(scratch + (uint32_t)(r * WB + ((lu ^ sw) << 4)), __float_as_uint(ld[j].x), __float_as_uint(ld[j].y),
             __float_as_uint(ld[j].z), __float_as_uint(ld[j].w));
    }
    __syncwarp();
    uint32_t v[4 * U];
    const int sw = NSPLIT == 2 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint4 x = lds128(scratch + (uint32_t)(lane * WB + ((k ^ sw) << 4)));
      v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
    }
    if constexpr (NSPLIT == 2) tmem_st32(tacc + col0, v);
    else tmem_st16(tacc + col0, v);
  }
  tmem_st_wait();
}

// Two tiles are in flight per CTA.  Each of the two groups of 8 compute warps owns one tile (its operand buffer
// in shared memory and two accumulators in tensor memory); the MMA warp walks the 7 weight chunks in lock-step over
// both tiles (A then B on the same weight slot), so that one tile's epilogue runs under the other tile's MMAs and
// every weight chunk is fetched once per tile pair.
template <int NSPLIT, bool TAPE>
__global__ void __launch_bounds__(TC_FWD_THREADS, 1) agg_tc_fwd_kernel(const TcParams p) {
  using S = TcSmem<NSPLIT>;
  constexpr int NSLOT = S::NSLOT;
  constexpr bool RESIDENT = (NSPLIT == 1);
  // no static shared memory in this kernel: the dynamic window starts at the CTA's shared base, which is 1024-byte
  // aligned (checked below) — the 227 KB budget has no room for alignment slack
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sAct = smem + S::OFF_ACT;             // group g, tile (kc, split) at g * ACT_BYTES + (kc * NSPLIT + split) * TC_TILE_BYTES
  uint8_t* sW = smem + S::OFF_W;                 // slot s: [hi tile][lo tile]
  float* sBias = (float*)(smem + S::OFF_BIAS);
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);
  uint64_t* w_full = bars;                        // [NSLOT] weights landed
  uint64_t* w_free = bars + NSLOT;                // [NSLOT] MMAs reading the slot have completed
  uint64_t* pe_ready = bars + 2 * NSLOT;          // [2 groups] PE tile written
  uint64_t* a_ready = bars + 2 * NSLOT + 2;       // [2 groups][2] activation K-chunk written
  uint64_t* acc_ready = bars + 2 * NSLOT + 6;     // [2 groups][2] accumulator buffer complete
  uint32_t* sTmem = (uint32_t*)(smem + S::OFF_TMEM);
  uint32_t* sLock = sTmem + 2;                    // tensor-pipe turn of the two MMA issuers

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const apn_agg_inputs& in = p.in;
  TRACE_DECL;
  // sample count: exact (host) or read here from the device counter (p.n_tiles then only sized the grid)
  const int Mrt = apn_rt_count(in.m_dev, in.M);
  const int n_tiles = (Mrt + TC_SAMPLES - 1) / TC_SAMPLES;
  // contiguous tile range of this CTA; group g takes every second tile starting at t_begin + g
  const int t_begin = (int)(((long long)n_tiles * blockIdx.x) / gridDim.x);
  const int t_end = (int)(((long long)n_tiles * (blockIdx.x + 1)) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(w_full + i, 1);
      mbar_init(w_free + i, 2);            // both issuers
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(pe_ready + g, TC_GROUP_THREADS);
      mbar_init(a_ready + 2 * g, TC_GROUP_THREADS);
      mbar_init(a_ready + 2 * g + 1, TC_GROUP_THREADS);
      mbar_init(acc_ready + 2 * g, 1);
      mbar_init(acc_ready + 2 * g + 1, 1);
    }
    *sLock = 0u;
    fence_barrier_init();
  }
  if (warp == TC_COMPUTE_WARPS + 1) tmem_alloc<TC_FWD_TMEM_COLS>(sTmem);
  // biases (+ the pose-embedding fold into the layer-0 bias)
  for (int i = tid; i < 4 * 128; i += TC_FWD_THREADS) {
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
  const int n_rounds = (t_end - t_begin + 1) >> 1;

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
        for (int r = 0; r < n_rounds; ++r) {
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
  } else if (warp > TC_COMPUTE_WARPS) {
    // ================================================================= MMA issuers: one thread per tile group
    // (tcgen05.mma issue is throttled at execution rate, so a single issuer would leave the tensor pipe idle during
    //  its own barrier waits and commits; two independent issuers interleave on the pipe and share every weight slot)
    const int g = warp - (TC_COMPUTE_WARPS + 1);
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 128);
      const uint32_t act_base = smem_u32(sAct) + (uint32_t)(g * S::ACT_BYTES), w_base = smem_u32(sW);
      uint32_t it = 0;
      uint32_t ph_pe = 0, ph_a[2] = {0, 0};
      // one K chunk of 64: 4 K steps, NSPLIT == 2 adds the two cross products
      auto mma_chunk = [&](uint32_t a_hi, uint32_t b_hi, uint32_t acc, bool first_chunk) {
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
      };
      // layer 0 accumulates onto the table rows P[idx] the compute warps stored in accumulator 0
      for (int r = 0; r < n_rounds; ++r) {
        const int tile = t_begin + 2 * r + g;
        const bool solo = (t_begin + 2 * r + 1 >= t_end);       // last round of an odd range: group 0 only
        if (tile >= t_end) break;
        for (int c = 0; c < TC_NCHUNKS; ++c, ++it) {
          const int layer = c == 0 ? 0 : (c + 1) >> 1, kc = c == 0 ? 0 : (c - 1) & 1;
          const uint32_t slot = RESIDENT ? (uint32_t)c : it % NSLOT;
          mbar_wait(w_full + slot, RESIDENT ? 0u : ((it / NSLOT) & 1));
          const uint32_t b_hi = w_base + slot * (uint32_t)(NSPLIT * TC_TILE_BYTES);
          uint64_t* ready = c == 0 ? pe_ready + g : a_ready + 2 * g + kc;
          uint32_t& ph = c == 0 ? ph_pe : ph_a[kc];
          TRACE(2 + g, 1000 + c * 10);
          mbar_wait(ready, ph);
          ph ^= 1;
          tc_fence_after();
          TRACE(2 + g, 2000 + c * 10);
          const uint32_t a_off = (uint32_t)(kc * NSPLIT * TC_TILE_BYTES);
          if (TAPE) {
            // training: the operand tiles ARE the tape (same layout): stream them out with bulk stores from shared memory
            uint8_t* tape_tile = p.tape + (size_t)tile * TC_TAPE_TILE_BYTES;
            bulk_s2g(tape_tile + (c == 0 ? TC_TAPE_PE(0) : TC_TAPE_ACT(layer - 1, kc, 0)), sAct + (size_t)g * S::ACT_BYTES + a_off,
                     2 * TC_TILE_BYTES);
            bulk_commit();
          }
          const uint32_t acc = tmem_base + (uint32_t)(g * 256 + (layer & 1) * 128);
          // one chunk at a time on the tensor pipe: fair interleaving of the two issuers would keep both tiles in
          // phase (MMA, then both epilogues); whole-chunk turns stagger them so one tile's epilogue hides under the
          // other tile's MMAs
          while (atomicCAS(sLock, 0u, 1u) != 0u) {
          }
          mma_chunk(act_base + a_off, b_hi, acc, kc == 0 && c != 0);
          atomicExch(sLock, 0u);
          if (c == 0 || kc == 1) {
            if (TAPE) bulk_wait_read();          // the epilogue released by this commit overwrites the stored tiles
            umma_commit(acc_ready + 2 * g + (layer & 1));
          }
          if (!RESIDENT) {
            umma_commit(w_free + slot);
            if (solo) umma_commit(w_free + slot);
          }
          TRACE(2 + g, 3000 + c * 10);
        }
      }
      if (TAPE) bulk_wait_all();
    }
  } else {
    // ================================================================= compute warps: group g owns tiles t_begin + g + 2 n
    const int g = warp >> 3, wg = warp & 7;
    const int q = wg & 3, half = wg >> 2;
    const int erow = q * 32 + lane;                       // accumulator row (tensor-memory lane) and PE row of this thread
    const uint32_t tacc0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 256);
    const uint32_t act = smem_u32(sAct) + (uint32_t)(g * S::ACT_BYTES);   // shared-space address of the group's operand buffer
    const uint32_t bias_s = smem_u32(sBias);
    const uint32_t scratch = act + (uint32_t)(NSPLIT * TC_TILE_BYTES) + (uint32_t)(wg * 2048 * NSPLIT);   // in the K-chunk-1 tiles
    uint64_t* my_pe = pe_ready + g;
    uint64_t* my_a = a_ready + 2 * g;
    uint64_t* my_acc = acc_ready + 2 * g;
    uint32_t ph_acc0 = 0, ph_acc1 = 0;
    TcRow row{0, 0.f};
    int tile = t_begin + g;
    if (tile < t_end) {
      uint32_t hi[16], lo[16];
      row = tc_prologue<NSPLIT>(p, Mrt, tile, erow, half, tc_row_load(in, Mrt, tile, erow), hi, lo);
      tc_store_pe<NSPLIT>(act, erow, half, hi, lo);
      tc_stage_ptable<NSPLIT>(p.ptable, row.idx, scratch, tacc0, half, lane);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(my_pe);
    }
    for (; tile < t_end; tile += 2) {
      const int m0 = tile * TC_SAMPLES;
      const int next = tile + 2;
      TcRowLoad nl{0, 0.f, 0.f, 0.f};
      uint8_t* tp = TAPE ? p.tape + (size_t)tile * TC_TAPE_TILE_BYTES : nullptr;
      // ---------------------------------------------------------------- epilogues of layers 0..2
#pragma unroll
      for (int layer = 0; layer < 3; ++layer) {
        // stage 1 of the next tile's prologue (4 registers, in flight during the layer-1 and layer-2 epilogues)
        if (layer == 1 && next < t_end) nl = tc_row_load(in, Mrt, next, erow);
        if (layer & 1) {
          mbar_wait(my_acc + 1, ph_acc1);
          ph_acc1 ^= 1;
        } else {
          mbar_wait(my_acc, ph_acc0);
          ph_acc0 ^= 1;
        }
        tc_fence_after();
        if (wg == 0 && lane == 0) TRACE(g, 100 + layer);
        const uint32_t tacc = tacc0 + (uint32_t)((layer & 1) * 128);
        // 4 pieces of 16 columns: (ph, pc); the tensor-memory load of the next piece is in flight while one is processed
        uint32_t v[2][16];
        tmem_ld16(tacc + half * 32, v[0]);
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) {
          const int ph = pi >> 1, pc = pi & 1;            // K chunk `ph` of the next layer's operand, 16-column piece pc
          const int c0 = half * 32 + pc * 16;             // column inside the chunk
          tmem_ld_wait();
          if (pi < 3) tmem_ld16(tacc + ((pi + 1) >> 1) * 64 + half * 32 + ((pi + 1) & 1) * 16, v[(pi + 1) & 1]);
          const uint32_t(&vv)[16] = v[pi & 1];
          const uint32_t bias = bias_s + (uint32_t)((layer * 128 + ph * 64 + c0) * 4);
          const uint32_t t_hi = act + (uint32_t)((ph * NSPLIT) * TC_TILE_BYTES) + (uint32_t)erow * 128u;
          uint32_t mbits = 0;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t hi[4], lo[4];
            const float4 b0 = lds128f(bias + u * 32), b1 = lds128f(bias + u * 32 + 16);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 bb = e < 2 ? b0 : b1;
              float2 y = __fadd2_rn(make_float2(__uint_as_float(vv[u * 8 + 2 * e]), __uint_as_float(vv[u * 8 + 2 * e + 1])),
                                    (e & 1) ? make_float2(bb.z, bb.w) : make_float2(bb.x, bb.y));
              if (TAPE) {
                mbits |= (y.x > 0.f ? 1u : 0u) << (u * 8 + 2 * e);
                mbits |= (y.y > 0.f ? 1u : 0u) << (u * 8 + 2 * e + 1);
              }
              split_half2(leaky2(y), hi[e], lo[e]);
            }
            const uint32_t o = (uint32_t)(((((c0 >> 3) + u) ^ (erow & 7)) & 7) << 4);
            sts128(t_hi + o, hi[0], hi[1], hi[2], hi[3]);
            if (NSPLIT == 2) sts128(t_hi + TC_TILE_BYTES + o, lo[0], lo[1], lo[2], lo[3]);
          }
          if (TAPE)
            *reinterpret_cast<uint16_t*>(tp + TC_TAPE_MASK(layer) + ((size_t)erow * 8 + ph * 4 + (c0 >> 4)) * 2) = (uint16_t)mbits;
          if (pc == 1) {
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(my_a + ph);
            if (wg == 0 && lane == 0) TRACE(g, 200 + layer * 10 + ph);
          }
        }
      }
      // ---------------------------------------------------------------- next tile's prologue, in registers (under the layer-3 MMAs)
      uint32_t nhi[16], nlo[16];
      TcRow nrow{0, 0.f};
      if (next < t_end) nrow = tc_prologue<NSPLIT>(p, Mrt, next, erow, half, nl, nhi, nlo);
      if (wg == 0 && lane == 0) TRACE(g, 300);
      // ---------------------------------------------------------------- layer 3 complete: the operand buffer is free
      mbar_wait(my_acc + 1, ph_acc1);
      ph_acc1 ^= 1;
      tc_fence_after();
      if (wg == 0 && lane == 0) TRACE(g, 103);
      if (next < t_end) {
        tc_store_pe<NSPLIT>(act, erow, half, nhi, nlo);
        tc_stage_ptable<NSPLIT>(p.ptable, nrow.idx, scratch, tacc0, half, lane);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(my_pe);                               // the next tile's layer-0 MMAs run under the final epilogue
      }
      // ---------------------------------------------------------------- out_k = LeakyReLU(acc + b3); h = sum_k idw_k out_k
      const bool b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const int c0 = ph * 64 + half * 32;
        float v[32];
        {
          uint32_t t0[16], t1[16];
          tmem_ld16(tacc0 + 128u + c0, t0);
          tmem_ld16(tacc0 + 128u + c0 + 16, t1);
          tmem_ld_wait();
          const uint32_t bias = bias_s + (uint32_t)((3 * 128 + c0) * 4);
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 ba = lds128f(bias + i4 * 16), bb = lds128f(bias + 64 + i4 * 16);
            const float2 a0 = leaky2(__fadd2_rn(make_float2(__uint_as_float(t0[4 * i4]), __uint_as_float(t0[4 * i4 + 1])), make_float2(ba.x, ba.y)));
            const float2 a1 = leaky2(__fadd2_rn(make_float2(__uint_as_float(t0[4 * i4 + 2]), __uint_as_float(t0[4 * i4 + 3])), make_float2(ba.z, ba.w)));
            const float2 c0_ = leaky2(__fadd2_rn(make_float2(__uint_as_float(t1[4 * i4]), __uint_as_float(t1[4 * i4 + 1])), make_float2(bb.x, bb.y)));
            const float2 c1_ = leaky2(__fadd2_rn(make_float2(__uint_as_float(t1[4 * i4 + 2]), __uint_as_float(t1[4 * i4 + 3])), make_float2(bb.z, bb.w)));
            v[4 * i4] = a0.x; v[4 * i4 + 1] = a0.y; v[4 * i4 + 2] = a1.x; v[4 * i4 + 3] = a1.y;
            v[16 + 4 * i4] = c0_.x; v[16 + 4 * i4 + 1] = c0_.y; v[16 + 4 * i4 + 2] = c1_.x; v[16 + 4 * i4 + 3] = c1_.y;
          }
        }
        if (TAPE) {
          uint8_t* g3 = tp + TC_TAPE_ACT(3, ph, 0) + (size_t)erow * 128;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_half2(v[u * 8 + 2 * e], v[u * 8 + 2 * e + 1], hi[e], lo[e]);
            const uint32_t o = (uint32_t)((((half * 4 + u) ^ (erow & 7)) & 7) << 4);
            *reinterpret_cast<uint4*>(g3 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(g3 + TC_TILE_BYTES + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= row.w;
        // reduce-scatter over the 8 lanes (rows) of a sample: 32 -> 16 -> 8 -> 4 columns per lane
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
        if (m < Mrt) {
          const int col = c0 + (b4 ? 16 : 0) + (b2 ? 8 : 0) + (b1 ? 4 : 0);
          *reinterpret_cast<float4*>(p.h + (size_t)m * APN_C + col) = make_float4(x4[0], x4[1], x4[2], x4[3]);
        }
      }
      tc_fence_before();
      if (wg == 0 && lane == 0) TRACE(g, 400);
      row = nrow;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_COMPUTE_WARPS + 1) tmem_dealloc<TC_FWD_TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// inference scratch: the reduced feature h and the packed head weights (heads_tc.cu)
struct TcScratch {
  float* h;
  void* heads_packed;
  size_t total;
};
static TcScratch tc_scratch_layout(char* base, int M) {
  TcScratch b;
  size_t o = 0;
  b.h = (float*)(base + o);
  o = apn_align(o + (size_t)M * APN_C * sizeof(float));
  b.heads_packed = base + o;
  o = apn_align(o + heads_tc_weights_bytes());
  b.total = o;
  return b;
}
extern "C" size_t apn_aggregate_tc_scratch_bytes(int M) { return M > 0 ? tc_scratch_layout(nullptr, M).total : 0; }

template <int NSPLIT, bool TAPE>
static int tc_launch(cudaStream_t st, const TcParams& p) {
  using S = TcSmem<NSPLIT>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory budget");
  APN_CUDA(cudaFuncSetAttribute(agg_tc_fwd_kernel<NSPLIT, TAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const int grid = p.n_tiles < APN_SM_COUNT ? p.n_tiles : APN_SM_COUNT;
  agg_tc_fwd_kernel<NSPLIT, TAPE><<<grid, TC_FWD_THREADS, S::TOTAL, st>>>(p);
  APN_LAUNCH_CHECK();
  return 0;
}

// per-tile tape of the decoder + (at the end) the packed head weights of this step (heads_tc.cu)
static size_t tc_tape_tiles_bytes(int M) { return (size_t)apn_div_up(M, TC_SAMPLES) * TC_TAPE_TILE_BYTES; }
extern "C" size_t apn_aggregate_tc_tape_bytes(int M) { return M > 0 ? tc_tape_tiles_bytes(M) + apn_align(heads_tc_weights_bytes()) : 0; }

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
    b.h = out->h;
    b.heads_packed = (char*)tape + tc_tape_tiles_bytes(M);
  } else {
    APN_CHECK_ARG(scratch && scratch_bytes >= apn_aggregate_tc_scratch_bytes(M) && (((uintptr_t)scratch) & 15) == 0,
                  "scratch too small or not 16-byte aligned");
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
  const int rc = precision == 0 ? tc_launch<1, false>(st, p) : tape ? tc_launch<2, true>(st, p) : tc_launch<2, false>(st, p);
  if (rc) return rc;
  // tensor-core heads; in training they also leave exp_d / fv / v0, the tape of the (fp32) heads backward
  return agg_heads_tc_launch(st, in, w, b.h, b.heads_packed, out->alpha, out->rgb, tape ? out->exp_d : nullptr,
                             tape ? out->fv : nullptr, tape ? out->v0 : nullptr);
}
