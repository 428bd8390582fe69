// Shared definitions of the tensor-core decoder kernels (aggregate_tc.cu: forward, aggregate_tc_bwd.cu: backward).
#pragma once
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

#define TC_ROWS 128
#define TC_SAMPLES 16
#define TC_NCHUNKS 7                 // chunk 0: layer-0 PE columns; 1..6: layers 1-3, two K chunks of 64 each
#define TC_TILE_BYTES 16384          // [128 x 64] fp16
#define TC_CHUNK_GBYTES (2 * TC_TILE_BYTES)   // packed global: hi tile then lo tile
#define TC_COMPUTE_WARPS 16
#define TC_COMPUTE_THREADS (32 * TC_COMPUTE_WARPS)
#define TC_THREADS (TC_COMPUTE_THREADS + 64)          // backward kernels: + producer warp + MMA warp
#define TC_FWD_THREADS (TC_COMPUTE_THREADS + 96)      // forward: + producer warp + one MMA warp per tile group
#define TC_TMEM_COLS 256             // two 128-column fp32 accumulators (backward kernels)
#define TC_FWD_TMEM_COLS 512         // forward: two tiles in flight, two accumulators each
#define TC_GROUP_WARPS 8             // forward: compute warps per tile group
#define TC_GROUP_THREADS (32 * TC_GROUP_WARPS)


// ---------------------------------------------------------------------------------------
// The "tape": what the training forward leaves for the backward, per 128-row tile, as shared-memory images
// (pre-swizzled [128 x 64] fp16 tiles) so that the backward kernels fetch them with plain bulk copies.
//   ACT(layer 0..3, kc 0..1, split hi/lo)   post-activation of every feat_net layer
//   PE(split)                               layer-0 positional-encoding operand
//   MASK(layer 0..2)                        uint16 per (row, 16-column group): pre-activation > 0
// ---------------------------------------------------------------------------------------
#define TC_TAPE_ACT(layer, kc, split) ((size_t)((((layer) * 2 + (kc)) * 2 + (split)) * TC_TILE_BYTES))
#define TC_TAPE_PE(split) ((size_t)((16 + (split)) * TC_TILE_BYTES))
#define TC_TAPE_MASK(layer) ((size_t)(18 * TC_TILE_BYTES + (layer) * 128 * 8 * 2))
#define TC_TAPE_TILE_BYTES ((size_t)(18 * TC_TILE_BYTES + 3 * 128 * 8 * 2))   // 294 KiB
// gradient tiles written by the dgrad kernel for the wgrad kernel: DY(layer 0..3, kc, split)
#define TC_DY_TILE_BYTES ((size_t)(16 * TC_TILE_BYTES))

// ---------------------------------------------------------------------------------------
// Column layout of the layer-0 positional-encoding operand tile (64 columns).  The order is private to the
// tensor-core kernels (the packed W0 tiles use the same permutation), chosen so that every thread produces
// whole 16-byte units from few accurate sincosf calls:
//   [16d, 16d+8)       sin(rel_c[d] * 2^i), i = 0..7        d = 0..2
//   [16d+8, 16d+16)    cos(rel_c[d] * 2^i), i = 0..7
//   48 + 4d + {0,1}    sin, i = 8, 9;   48 + 4d + {2,3}  cos, i = 8, 9
//   60..62             rel_c;   63  zero padding
// tc_pe_ref_col maps a tile column to the reference column of poc_fre (lib/tineuvox.py:872-878):
// [x(3) | sin: 3 + 10 d + i | cos: 33 + 10 d + i], or -1 for the padding column.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int tc_pe_ref_col(int c) {
  if (c < 48) return 3 + 30 * ((c >> 3) & 1) + 10 * (c >> 4) + (c & 7);
  if (c < 60) return 3 + 30 * (((c - 48) >> 1) & 1) + 10 * ((c - 48) >> 2) + 8 + ((c - 48) & 1);
  return c < 63 ? c - 60 : -1;
}
// sin / cos of x * 2^(I0 + i), i < N: one accurate sincosf every 4 octaves, exact angle doubling in between
// (the absolute error grows by at most 8x over three doublings: <= 1e-6, below the fp16 hi/lo split of the operand)
template <int I0, int N>
__device__ __forceinline__ void tc_pe_octaves(float x, float (&sn)[N], float (&cs)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if ((i & 3) == 0) {
      sincosf(x * (float)(1 << (I0 + i)), &sn[i], &cs[i]);
    } else {
      sn[i] = 2.f * sn[i - 1] * cs[i - 1];
      cs[i] = fmaf(-2.f * sn[i - 1], sn[i - 1], 1.f);
    }
  }
}

__device__ __forceinline__ float leaky(float y) { return fmaxf(y, 0.01f * y); }

// bar.sync among the compute warps only (backward kernels)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE_THREADS) : "memory"); }
