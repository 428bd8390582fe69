// K3 heads on the tensor cores (inference): densitynet / Raw2Alpha and RGBNet on the reduced feature h (M,128).
// Replaces lib/tineuvox.py:65-88,158,396-400 + lib/temporalpoints.py:503-515 for the render path; the training
// path keeps the fp32 heads of aggregate.cu (their intermediates are its tape).
//
// One CTA (4 warps) per 128-sample tile, persistent over tiles; everything in split-fp16 (x = hi + lo, three MMAs per
// K step, fp32 accumulate in tensor memory: fp32-class results):
//   A   warp-per-row, coalesced: h row -> density dot (warp reduce) + hi/lo operand tiles (K-major, 128-byte swizzle)
//   G1  f = h Wf^T                     tcgen05.mma 128 x 128 x 16, 2 K-chunks          (feature_linears, no activation)
//       meanwhile: view-direction PE of every row -> operand chunk 2
//   B   thread = row: f + b_f -> hi/lo operand tiles (in place of h)
//   G2  v0 = [f | PE(view)] Wv0^T      tcgen05.mma 128 x 64 x 16, 2 K-chunks + 32 PE columns
//   C   thread = row: ReLU(v0 + b_v0) -> 64 -> 3 on the CUDA cores -> sigmoid -> rgb; alpha from the density
// Weights (Wf 64 KB, Wv0 48 KB as pre-swizzled hi/lo tiles) are fetched once per CTA with bulk async copies.
#include "aggregate_tc.cuh"

#define HT_THREADS 128
#define HT_WF_BYTES (2 * 2 * TC_TILE_BYTES)            // 2 K-chunks x (hi, lo) x [128 x 64]
#define HT_WV_TILE 8192                                 // [64 x 64] fp16
#define HT_WV_BYTES (3 * 2 * HT_WV_TILE)                // 3 K-chunks x (hi, lo)
#define HT_PACKED_BYTES (HT_WF_BYTES + HT_WV_BYTES)
#define HT_TMEM_COLS 256                                // f: columns 0..127, v0: 128..191

size_t heads_tc_weights_bytes() { return HT_PACKED_BYTES; }

// blocks 0,1: Wf K-chunks; blocks 2..4: Wv0 K-chunks (input columns [f(128) | view PE(27)], zero padded to 192)
__global__ void heads_tc_pack_kernel(const float* __restrict__ Wf, const float* __restrict__ Wv0, uint8_t* __restrict__ packed) {
  const int c = blockIdx.x;
  if (c < 2) {
    uint8_t* base = packed + (size_t)c * 2 * TC_TILE_BYTES;
    for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
      const int n = e >> 6, k = e & 63;
      __half hi, lo;
      split_half(Wf[(size_t)n * APN_C + c * 64 + k], hi, lo);
      const uint32_t o = sw128_offset(n, k);
      *reinterpret_cast<__half*>(base + o) = hi;
      *reinterpret_cast<__half*>(base + TC_TILE_BYTES + o) = lo;
    }
  } else {
    const int kc = c - 2;
    uint8_t* base = packed + HT_WF_BYTES + (size_t)kc * 2 * HT_WV_TILE;
    const int KV = APN_C + APN_PE_VIEW;
    for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
      const int n = e >> 6, k = e & 63;
      const int col = kc * 64 + k;
      __half hi, lo;
      split_half(col < KV ? Wv0[(size_t)n * KV + col] : 0.f, hi, lo);
      const uint32_t o = sw128_offset(n, k);
      *reinterpret_cast<__half*>(base + o) = hi;
      *reinterpret_cast<__half*>(base + HT_WV_TILE + o) = lo;
    }
  }
}

struct HtParams {
  const float* h;              // (M,128)
  const int* ray_id;           // (M)
  const float* viewdirs;       // (R,3)
  const uint8_t* packed;
  const float *bf, *bv0, *W2, *b2, *wd, *bd;
  float act_shift, interval;
  float* alpha;                // (M)
  float* rgb;                  // (M,3)
  // training tape of the fp32 heads backward (aggregate.cu: agg_rgbnet_bwd_launch, density backward), or NULL
  float* exp_d;                // (M)      exp(density + act_shift)
  float* fv;                   // (M,160)  [f (128) | view PE (27) | 0 (5)]
  float* v0;                   // (M,64)   ReLU(views_linears.0)
  int M, n_tiles;              // capacity when m_dev is given
  const int32_t* m_dev;        // device-side sample count or NULL
};

struct HtSmem {
  static constexpr int OFF_A = 0;                                   // 3 K-chunks x (hi, lo) x [128 x 64]
  static constexpr int OFF_WF = OFF_A + 3 * 2 * TC_TILE_BYTES;
  static constexpr int OFF_WV = OFF_WF + HT_WF_BYTES;
  static constexpr int OFF_F = OFF_WV + HT_WV_BYTES;                // floats: bf 128 | bv0 64 | W2 192 | b2 4 | wd 128
  static constexpr int N_F = 128 + 64 + 192 + 4 + 128;
  static constexpr int OFF_BAR = OFF_F + N_F * 4;                   // 3 mbarriers
  static constexpr int OFF_TMEM = OFF_BAR + 3 * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
};

__global__ void __launch_bounds__(HT_THREADS, 1) heads_tc_kernel(const HtParams p) {
  using S = HtSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* sF = (float*)(smem + S::OFF_F);
  float *sBf = sF, *sBv = sF + 128, *sW2 = sF + 192, *sB2 = sF + 384, *sWd = sF + 388;
  uint64_t* bars = (uint64_t*)(smem + S::OFF_BAR);      // [0] weights, [1] f complete, [2] v0 complete
  uint32_t* sTmem = (uint32_t*)(smem + S::OFF_TMEM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sA = smem_u32(smem + S::OFF_A), sWf = smem_u32(smem + S::OFF_WF), sWv = smem_u32(smem + S::OFF_WV);
  const int M = apn_rt_count(p.m_dev, p.M), n_tiles = (M + 127) >> 7;

  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    mbar_init(bars + 2, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bars, HT_PACKED_BYTES);
    bulk_g2s(smem + S::OFF_WF, p.packed, HT_WF_BYTES, bars);
    bulk_g2s(smem + S::OFF_WV, p.packed + HT_WF_BYTES, HT_WV_BYTES, bars);
  }
  if (warp == 0) tmem_alloc<HT_TMEM_COLS>(sTmem);
  for (int i = tid; i < S::N_F; i += HT_THREADS) {
    float v = 0.f;
    if (i < 128) v = __ldg(p.bf + i);
    else if (i < 192) v = __ldg(p.bv0 + i - 128);
    else if (i < 384) v = __ldg(p.W2 + i - 192);
    else if (i < 387) v = __ldg(p.b2 + i - 384);
    else if (i >= 388) v = __ldg(p.wd + i - 388);
    sF[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sTmem;
  const float bd = __ldg(p.bd);
  const float4 wd4 = *reinterpret_cast<const float4*>(sWd + 4 * lane);
  constexpr uint32_t idesc1 = umma_idesc_f16(128, 128), idesc2 = umma_idesc_f16(128, 64);
  uint32_t ph = 0;
  bool w_ready = false;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ph ^= 1) {
    const int m0 = tile * 128;
    // ---------------------------------------------------------------- A: rows of h, one warp per row
    float my_dens = 0.f;
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      const int r = warp * 32 + rr;
      const int m = min(m0 + r, M - 1);
      const float4 x = __ldg(reinterpret_cast<const float4*>(p.h + (size_t)m * APN_C) + lane);
      const float d = warp_sum(x.x * wd4.x + x.y * wd4.y + x.z * wd4.z + x.w * wd4.w);
      if (lane == rr) my_dens = d;
      uint32_t hi0, lo0, hi1, lo1;
      split_half2(make_float2(x.x, x.y), hi0, lo0);
      split_half2(make_float2(x.z, x.w), hi1, lo1);
      const int c = (4 * lane) & 63;
      const uint32_t a = sA + (uint32_t)((lane >> 4) * 2 * TC_TILE_BYTES) + sw128_offset(r, c);
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(hi0), "r"(hi1) : "memory");
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a + TC_TILE_BYTES), "r"(lo0), "r"(lo1) : "memory");
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (!w_ready) {
        mbar_wait(bars, 0);
        w_ready = true;
      }
      tc_fence_after();
#pragma unroll
      for (int kc = 0; kc < 2; ++kc)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a_hi = sA + kc * 2 * TC_TILE_BYTES + ks * 32, b_hi = sWf + kc * 2 * TC_TILE_BYTES + ks * 32;
          const uint64_t da = umma_desc_k_sw128(a_hi), db = umma_desc_k_sw128(b_hi);
          const uint64_t da_lo = umma_desc_k_sw128(a_hi + TC_TILE_BYTES), db_lo = umma_desc_k_sw128(b_hi + TC_TILE_BYTES);
          umma_f16(tmem_base, da, db, idesc1, (kc | ks) ? 1u : 0u);
          umma_f16(tmem_base, da, db_lo, idesc1, 1u);
          umma_f16(tmem_base, da_lo, db, idesc1, 1u);
        }
      umma_commit(bars + 1);
    }
    // ---------------------------------------------------------------- (under G1) alpha + view PE -> operand chunk 2
    const int mrow = m0 + tid;
    const int mc = min(mrow, M - 1);
    {
      // lib/cuda/render_utils_kernel.cu:358-370 on this lane's row of the warp
      const float e = expf(my_dens + bd + p.act_shift);
      if (mrow < M) {
        p.alpha[mrow] = 1.f - powf(1.f + e, -p.interval);
        if (p.exp_d) p.exp_d[mrow] = e;
      }
      // lib/tineuvox.py:872-878 with 4 frequencies: [v(3) | sin(v_d 2^i) d-major | cos(...)], zero padded to 32 columns
      const int ray = __ldg(p.ray_id + mc);
      float pe[32];
#pragma unroll
      for (int i = 27; i < 32; ++i) pe[i] = 0.f;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float v = __ldg(p.viewdirs + 3 * (size_t)ray + d);
        pe[d] = v;
#pragma unroll
        for (int i = 0; i < 4; ++i) sincosf(v * (float)(1 << i), &pe[3 + d * 4 + i], &pe[15 + d * 4 + i]);
      }
      if (p.fv && mrow < M) {
        float4* dst = reinterpret_cast<float4*>(p.fv + (size_t)mrow * 160 + APN_C);
#pragma unroll
        for (int u = 0; u < 8; ++u) dst[u] = make_float4(pe[4 * u], pe[4 * u + 1], pe[4 * u + 2], pe[4 * u + 3]);
      }
      const uint32_t t = sA + 2 * 2 * TC_TILE_BYTES + (uint32_t)tid * 128u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) split_half2(make_float2(pe[u * 8 + 2 * e2], pe[u * 8 + 2 * e2 + 1]), hi[e2], lo[e2]);
        const uint32_t o = (uint32_t)(((u ^ (tid & 7)) & 7) << 4);
        sts128(t + o, hi[0], hi[1], hi[2], hi[3]);
        sts128(t + TC_TILE_BYTES + o, lo[0], lo[1], lo[2], lo[3]);
      }
    }
    // ---------------------------------------------------------------- B: f = acc + b_f -> operand chunks 0, 1
    mbar_wait(bars + 1, ph);
    tc_fence_after();
    const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int pc = 0; pc < 8; ++pc) {
      uint32_t v[16];
      tmem_ld16(tacc + pc * 16, v);
      tmem_ld_wait();
      const uint32_t t = sA + (uint32_t)((pc >> 2) * 2 * TC_TILE_BYTES) + (uint32_t)tid * 128u;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t hi[4], lo[4];
        float fr[8];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const int col = pc * 16 + u * 8 + 2 * e2;
          fr[2 * e2] = __uint_as_float(v[u * 8 + 2 * e2]) + sBf[col];
          fr[2 * e2 + 1] = __uint_as_float(v[u * 8 + 2 * e2 + 1]) + sBf[col + 1];
          split_half2(make_float2(fr[2 * e2], fr[2 * e2 + 1]), hi[e2], lo[e2]);
        }
        if (p.fv && mrow < M) {
          float4* dst = reinterpret_cast<float4*>(p.fv + (size_t)mrow * 160 + pc * 16 + u * 8);
          dst[0] = make_float4(fr[0], fr[1], fr[2], fr[3]);
          dst[1] = make_float4(fr[4], fr[5], fr[6], fr[7]);
        }
        const uint32_t o = (uint32_t)(((((pc & 3) * 2 + u) ^ (tid & 7)) & 7) << 4);
        sts128(t + o, hi[0], hi[1], hi[2], hi[3]);
        sts128(t + TC_TILE_BYTES + o, lo[0], lo[1], lo[2], lo[3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t acc2 = tmem_base + 128u;
#pragma unroll
      for (int kc = 0; kc < 3; ++kc)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (kc == 2 && ks >= 2) break;                              // the view PE has 32 columns
          const uint32_t a_hi = sA + kc * 2 * TC_TILE_BYTES + ks * 32, b_hi = sWv + kc * 2 * HT_WV_TILE + ks * 32;
          const uint64_t da = umma_desc_k_sw128(a_hi), db = umma_desc_k_sw128(b_hi);
          const uint64_t da_lo = umma_desc_k_sw128(a_hi + TC_TILE_BYTES), db_lo = umma_desc_k_sw128(b_hi + HT_WV_TILE);
          umma_f16(acc2, da, db, idesc2, (kc | ks) ? 1u : 0u);
          umma_f16(acc2, da, db_lo, idesc2, 1u);
          umma_f16(acc2, da_lo, db, idesc2, 1u);
        }
      umma_commit(bars + 2);
    }
    // ---------------------------------------------------------------- C: ReLU -> 64 -> 3 -> sigmoid
    mbar_wait(bars + 2, ph);
    tc_fence_after();
    float a0 = sB2[0], a1 = sB2[1], a2 = sB2[2];
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
      uint32_t v[16];
      tmem_ld16(tacc + 128u + pc * 16, v);
      tmem_ld_wait();
#pragma unroll
      float xr[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = pc * 16 + i;
        const float x = fmaxf(__uint_as_float(v[i]) + sBv[col], 0.f);
        xr[i] = x;
        a0 = fmaf(x, sW2[col], a0);
        a1 = fmaf(x, sW2[64 + col], a1);
        a2 = fmaf(x, sW2[128 + col], a2);
      }
      if (p.v0 && mrow < M) {
        float4* dst = reinterpret_cast<float4*>(p.v0 + (size_t)mrow * 64 + pc * 16);
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[u] = make_float4(xr[4 * u], xr[4 * u + 1], xr[4 * u + 2], xr[4 * u + 3]);
      }
    }
    if (mrow < M) {
      p.rgb[3 * (size_t)mrow] = 1.f / (1.f + expf(-a0));
      p.rgb[3 * (size_t)mrow + 1] = 1.f / (1.f + expf(-a1));
      p.rgb[3 * (size_t)mrow + 2] = 1.f / (1.f + expf(-a2));
    }
    tc_fence_before();
    __syncthreads();          // tensor-memory reads and operand tiles are done before the next tile reuses them
  }
  if (!w_ready && tid == 0) mbar_wait(bars, 0);      // never leave with a bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<HT_TMEM_COLS>(tmem_base);
}

// packs the head weights into `packed` (HT_PACKED_BYTES, 16-byte aligned) and runs the heads on h
int agg_heads_tc_launch(cudaStream_t st, const apn_agg_inputs* in, const apn_mlp_weights* w, const float* h, void* packed,
                        float* alpha, float* rgb, float* exp_d, float* fv, float* v0) {
  const int M = in->M;
  if (M <= 0) return 0;
  heads_tc_pack_kernel<<<5, 256, 0, st>>>(w->rgb_feat_w, w->rgb_v0_w, (uint8_t*)packed);
  APN_LAUNCH_CHECK();
  HtParams p;
  p.h = h; p.ray_id = in->ray_id; p.viewdirs = in->viewdirs; p.packed = (const uint8_t*)packed;
  p.bf = w->rgb_feat_b; p.bv0 = w->rgb_v0_b; p.W2 = w->rgb_v2_w; p.b2 = w->rgb_v2_b; p.wd = w->density_w; p.bd = w->density_b;
  p.act_shift = in->act_shift; p.interval = in->interval;
  p.alpha = alpha; p.rgb = rgb; p.exp_d = exp_d; p.fv = fv; p.v0 = v0; p.M = M; p.n_tiles = apn_div_up(M, 128);
  p.m_dev = in->m_dev;
  static_assert(HtSmem::TOTAL <= 227 * 1024, "shared memory budget");
  APN_CUDA(cudaFuncSetAttribute(heads_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HtSmem::TOTAL));
  const int grid = p.n_tiles < APN_SM_COUNT ? p.n_tiles : APN_SM_COUNT;
  heads_tc_kernel<<<grid, HT_THREADS, HtSmem::TOTAL, st>>>(p);
  APN_LAUNCH_CHECK();
  return 0;
}
