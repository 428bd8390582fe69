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
#define TC_THREADS (TC_COMPUTE_THREADS + 64)
#define TC_TMEM_COLS 256             // two 128-column fp32 accumulators


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

__device__ __forceinline__ float leaky(float y) { return fmaxf(y, 0.01f * y); }

// bar.sync among the compute warps only
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE_THREADS) : "memory"); }
