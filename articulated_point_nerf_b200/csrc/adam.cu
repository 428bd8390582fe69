// K4b — multi-tensor Adam.  Replaces adam_upd_cuda.{adam_upd, masked_adam_upd, adam_upd_with_perlr}
// (lib/cuda/adam_upd_kernel.cu:9-132) as called once per parameter tensor by
// lib/masked_adam.py:39-72: here one launch updates up to 64 tensors (the ~30 parameter tensors
// of stage 2 are one launch instead of ~30).
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= step_size*m/(sqrt(v)+eps)
// Every multiply/add is a separate IEEE op (repo-wide contract, see common.cuh) so the result is
// bit-identical to the torch-CPU oracle.  HBM-bound: 28 B/element (16 read, 12 written); masked
// elements (g == 0) cost 4 B.  float4 vectorised main body, scalar tail and unaligned fallback.
#include <math.h>

#include "common.cuh"

#define ADAM_MAX_TENSORS 64
#define ADAM_CHUNK 4096   // elements per block-iteration

struct AdamBatch {
  apn_adam_tensor t[ADAM_MAX_TENSORS];
  int chunk_start[ADAM_MAX_TENSORS + 1];   // prefix of ceil(numel/ADAM_CHUNK)
  int n;
  const float* step_sizes;                 // device array (one per tensor of this batch) overriding t.step_size, or NULL
  const int32_t* skip;                     // device word: a non-zero value turns the launch into a no-op, or NULL
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float perlr, float ss, float b1, float b2,
                                          float omb1, float omb2, float eps, int mode) {
  if (mode == 1 && g == 0.f) return;
  m = __fadd_rn(__fmul_rn(b1, m), __fmul_rn(omb1, g));
  v = __fadd_rn(__fmul_rn(b2, v), __fmul_rn(__fmul_rn(omb2, g), g));
  const float s = (mode == 2) ? __fmul_rn(ss, perlr) : ss;
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(s, m), __fadd_rn(__fsqrt_rn(v), eps)));
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamBatch batch, float b1, float b2, float eps) {
  const float omb1 = __fsub_rn(1.f, b1), omb2 = __fsub_rn(1.f, b2);
  if (batch.skip && *reinterpret_cast<const volatile int32_t*>(batch.skip) != 0) return;   // e.g. a truncated sample workspace
  const int total_chunks = batch.chunk_start[batch.n];
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    int ti = 0;
    while (batch.chunk_start[ti + 1] <= c) ++ti;     // n <= 64: linear search in constant bank
    const apn_adam_tensor& t = batch.t[ti];
    const float step_size = batch.step_sizes ? __ldg(batch.step_sizes + ti) : t.step_size;
    const long long base = (long long)(c - batch.chunk_start[ti]) * ADAM_CHUNK;
    const long long n = min((long long)ADAM_CHUNK, t.numel - base);
    float* p = t.param + base;
    const float* g = t.grad + base;
    float* m = t.exp_avg + base;
    float* v = t.exp_avg_sq + base;
    const float* pl = t.perlr ? t.perlr + base : nullptr;
    const bool aligned = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)pl) & 15) == 0);
    const long long n4 = aligned ? (n >> 2) : 0;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 P = reinterpret_cast<float4*>(p)[i];
      const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
      if (t.mode == 1 && G.x == 0.f && G.y == 0.f && G.z == 0.f && G.w == 0.f) continue;
      float4 M = reinterpret_cast<float4*>(m)[i];
      float4 V = reinterpret_cast<float4*>(v)[i];
      float4 PL = make_float4(1.f, 1.f, 1.f, 1.f);
      if (pl) PL = __ldg(reinterpret_cast<const float4*>(pl) + i);
      adam_elem(P.x, G.x, M.x, V.x, PL.x, step_size, b1, b2, omb1, omb2, eps, t.mode);
      adam_elem(P.y, G.y, M.y, V.y, PL.y, step_size, b1, b2, omb1, omb2, eps, t.mode);
      adam_elem(P.z, G.z, M.z, V.z, PL.z, step_size, b1, b2, omb1, omb2, eps, t.mode);
      adam_elem(P.w, G.w, M.w, V.w, PL.w, step_size, b1, b2, omb1, omb2, eps, t.mode);
      reinterpret_cast<float4*>(p)[i] = P;
      reinterpret_cast<float4*>(m)[i] = M;
      reinterpret_cast<float4*>(v)[i] = V;
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float P = p[i], M = m[i], V = v[i];
      adam_elem(P, g[i], M, V, pl ? pl[i] : 1.f, step_size, b1, b2, omb1, omb2, eps, t.mode);
      p[i] = P; m[i] = M; v[i] = V;
    }
  }
}

// lib/cuda/adam_upd_kernel.cu:72 — evaluated in float, like the reference host code
extern "C" float apn_adam_step_size(int step, float beta1, float beta2, float lr) {
  const float fs = (float)step;
  return lr * sqrtf(1.f - powf(beta2, fs)) / (1.f - powf(beta1, fs));
}

static int adam_multi_launch(const apn_adam_tensor* tensors, int n_tensors, float beta1, float beta2, float eps,
                             const float* step_sizes_dev, const int32_t* skip_dev, cudaStream_t stream) {
  APN_CHECK_ARG(n_tensors >= 0 && (n_tensors == 0 || tensors), "bad tensor list");
  for (int s = 0; s < n_tensors; s += ADAM_MAX_TENSORS) {
    AdamBatch b;
    memset(&b, 0, sizeof(b));
    b.n = (n_tensors - s < ADAM_MAX_TENSORS) ? n_tensors - s : ADAM_MAX_TENSORS;
    b.step_sizes = step_sizes_dev ? step_sizes_dev + s : nullptr;
    b.skip = skip_dev;
    long long chunks = 0;
    for (int i = 0; i < b.n; ++i) {
      const apn_adam_tensor& t = tensors[s + i];
      APN_CHECK_ARG(t.param && t.grad && t.exp_avg && t.exp_avg_sq && t.numel >= 0, "null tensor in Adam list");
      APN_CHECK_ARG(t.mode >= 0 && t.mode <= 2 && (t.mode != 2 || t.perlr), "bad Adam mode");
      b.t[i] = t;
      b.chunk_start[i] = (int)chunks;
      chunks += (t.numel + ADAM_CHUNK - 1) / ADAM_CHUNK;
      APN_CHECK_ARG(chunks < (1ll << 31), "too many elements for one batch");
    }
    b.chunk_start[b.n] = (int)chunks;
    if (chunks == 0) continue;
    const int grid = (int)(chunks < (long long)APN_SM_COUNT * 8 ? chunks : (long long)APN_SM_COUNT * 8);
    adam_multi_kernel<<<grid, 256, 0, stream>>>(b, beta1, beta2, eps);
    APN_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int apn_adam_multi(const apn_adam_tensor* tensors, int n_tensors, float beta1, float beta2, float eps,
                              apn_stream_t stream_) {
  return adam_multi_launch(tensors, n_tensors, beta1, beta2, eps, nullptr, nullptr, (cudaStream_t)stream_);
}

// Same update with the per-tensor step sizes read from DEVICE memory (step_sizes_dev[n_tensors], refreshed by the host
// before a captured CUDA graph is replayed: kernel arguments are frozen at capture, the bias-corrected step size changes
// every iteration) and an optional device-side skip word (non-zero: no update — the step's sample workspace overflowed and
// the step is re-run; apn_sample_knn_static counts[2]).
extern "C" int apn_adam_multi_dev(const apn_adam_tensor* tensors, int n_tensors, float beta1, float beta2, float eps,
                                  const float* step_sizes_dev, const int32_t* skip_dev, apn_stream_t stream_) {
  APN_CHECK_ARG(step_sizes_dev, "step_sizes_dev is required (use apn_adam_multi for host-side step sizes)");
  return adam_multi_launch(tensors, n_tensors, beta1, beta2, eps, step_sizes_dev, skip_dev, (cudaStream_t)stream_);
}
