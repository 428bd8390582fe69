// Pose chain in one launch: TransformNet MLP -> 4-parameter Rodrigues -> per-joint pivot transform -> kinematic chain
// (lib/pointwarper.py:5-37,118-193,217-236), forward and backward.  J <= 128 joints and a 256-wide, 5-layer MLP on a
// single time sample are a few hundred tiny PyTorch launches per training step; here they are two kernels, each ONE
// thread-block cluster of 8 CTAs: the work is ~1-2 MB of weight traffic whose latency is the whole cost, so every layer's
// rows are split over the 8 SMs of the cluster (8x the loads in flight) and the 256-float activations / partial sums are
// exchanged through distributed shared memory with one cluster barrier per layer.
//
// Conventions (checked by the Python wrapper): node i's parent node index is smaller than i (bone i = [parent, i+1],
// lib/pointwarper.py:105-111), so  T_i = T_parent(i) * M_i  can be evaluated in index order and differentiated in
// reverse order;  M_i = [R_i | p - R_i p] with p = joints[pivot(i)] (the parent joint; the root pivots on itself).
#include <cooperative_groups.h>

#include "common.cuh"
namespace cg = cooperative_groups;

#define POSE_CLUSTER 8        // CTAs per launch (one cluster)
#define POSE_H 256            // hidden width of TransformNet (lib/pointwarper.py:6)
#define POSE_MAX_J 128
#define POSE_MAX_T 64
#define POSE_THREADS 256

struct PoseArgs {
  const float* t_embed;       // (t_dim)
  const float* w[5];          // (256,t_dim) (256,256)x3 ((J+1)*4,256)
  const float* b[4];
  const float* joints;        // (J,3)
  const int* parent_node;     // (J)  -1 for the root
  const int* pivot;           // (J)  joint whose position is the rotation pivot
  const int* sibling;         // (J)  rotation shared from node sibling[i] (lib/pointwarper.py:232)
  const uint8_t* rot_mask;    // (J)  1: rotation frozen to identity (:233-234), or NULL
  int J, t_dim;
};

// R = c I + (1-c) n n^T + s [n]x, written out as lib/pointwarper.py:128-141
__device__ __forceinline__ void rodrigues4(const float p[4], float R[9], float n[3], float& r_len, float& cs, float& sn) {
  r_len = sqrtf(1e-5f + p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  n[0] = p[0] / r_len; n[1] = p[1] / r_len; n[2] = p[2] / r_len;
  sincosf(p[3], &sn, &cs);
  const float x = n[0], y = n[1], z = n[2], C = 1.f - cs;
  R[0] = x * x + (1.f - x * x) * cs; R[1] = x * y * C - z * sn;       R[2] = x * z * C + y * sn;
  R[3] = x * y * C + z * sn;       R[4] = y * y + (1.f - y * y) * cs; R[5] = y * z * C - x * sn;
  R[6] = x * z * C - y * sn;       R[7] = y * z * C + x * sn;       R[8] = z * z + (1.f - z * z) * cs;
}

// y[n] = act(dot(W[n,:K], x) + b[n]) for the rows n of THIS CTA's slice of [0, n_out); every result is stored into the
// shared memory of all CTAs of the cluster (lane r of the writing warp stores to rank r), so after the next
// cluster barrier each CTA holds the complete activation vector.  One warp per row, 4 rows in flight per warp.
__device__ __forceinline__ void dense_layer_cluster(cg::cluster_group& cluster, const float* __restrict__ W,
                                                    const float* __restrict__ b, const float* x, float* y, int n_out, int K,
                                                    bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int per = (n_out + POSE_CLUSTER - 1) / POSE_CLUSTER;
  const int r0 = (int)cluster.block_rank() * per, r1 = min(n_out, r0 + per);
  constexpr int RB = 4;
  float* y_remote = lane < POSE_CLUSTER ? cluster.map_shared_rank(y, lane) : nullptr;
  for (int n0 = r0 + warp * RB; n0 < r1; n0 += nw * RB) {
    float s[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) s[r] = 0.f;
    for (int k = lane; k < K; k += 32) {
      float wv[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) wv[r] = (n0 + r < r1) ? __ldg(W + (size_t)(n0 + r) * K + k) : 0.f;
      const float xv = x[k];
#pragma unroll
      for (int r = 0; r < RB; ++r) s[r] = fmaf(wv[r], xv, s[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const float t = warp_sum(s[r]);
      if (n0 + r < r1) {
        float v = t + (b ? __ldg(b + n0 + r) : 0.f);
        v = relu ? fmaxf(v, 0.f) : v;
        if (lane < POSE_CLUSTER) y_remote[n0 + r] = v;
      }
    }
  }
}

// saved layout (floats): h1..h4 (4 x 256) | params ((J+1)*4) | R_raw (J*9) | M (J*12) | T (J*12)
__host__ __device__ inline int pose_saved_floats(int J) { return 4 * POSE_H + (J + 1) * 4 + J * 9 + J * 12 + J * 12; }

__global__ void __cluster_dims__(POSE_CLUSTER, 1, 1) __launch_bounds__(POSE_THREADS)
pose_fwd_kernel(const PoseArgs a, float* __restrict__ bone_T, float* __restrict__ global_t, float* __restrict__ thetas,
                float* __restrict__ saved) {
  __shared__ float sh[5][POSE_H];
  __shared__ float sP[(POSE_MAX_J + 1) * 4];
  __shared__ float sR[POSE_MAX_J * 9];
  __shared__ float sM[POSE_MAX_J * 12];
  __shared__ float sT[POSE_MAX_J * 12];
  cg::cluster_group cluster = cg::this_cluster();
  const int J = a.J, tid = threadIdx.x;
  for (int i = tid; i < a.t_dim; i += blockDim.x) sh[0][i] = a.t_embed[i];
  cluster.sync();                                   // every CTA's shared memory exists before the first remote store
  dense_layer_cluster(cluster, a.w[0], a.b[0], sh[0], sh[1], POSE_H, a.t_dim, true);
  cluster.sync();
  for (int l = 1; l < 4; ++l) {
    dense_layer_cluster(cluster, a.w[l], a.b[l], sh[l], sh[l + 1], POSE_H, POSE_H, true);
    cluster.sync();
  }
  dense_layer_cluster(cluster, a.w[4], nullptr, sh[4], sP, (J + 1) * 4, POSE_H, false);
  cluster.sync();                                   // last remote store done: from here on only rank 0 works
  if (cluster.block_rank() != 0) return;
  // raw rotations
  if (tid < J) {
    float R[9], n[3], rl, cs, sn;
    rodrigues4(sP + 4 * tid, R, n, rl, cs, sn);
#pragma unroll
    for (int c = 0; c < 9; ++c) sR[9 * tid + c] = R[c];
    thetas[tid] = sP[4 * tid + 3];
  }
  if (tid < 3) global_t[tid] = sP[4 * J + tid];
  __syncthreads();
  // local transforms M_i = [R_i | p - R_i p]
  if (tid < J) {
    float R[9];
    const bool frozen = a.rot_mask && a.rot_mask[tid];
    const int src = a.sibling[tid];
#pragma unroll
    for (int c = 0; c < 9; ++c) R[c] = frozen ? ((c % 4 == 0) ? 1.f : 0.f) : sR[9 * src + c];
    const int pj = a.pivot[tid];
    const float p[3] = {a.joints[3 * pj], a.joints[3 * pj + 1], a.joints[3 * pj + 2]};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      sM[12 * tid + 4 * r] = R[3 * r]; sM[12 * tid + 4 * r + 1] = R[3 * r + 1]; sM[12 * tid + 4 * r + 2] = R[3 * r + 2];
      sM[12 * tid + 4 * r + 3] = p[r] - (R[3 * r] * p[0] + R[3 * r + 1] * p[1] + R[3 * r + 2] * p[2]);
    }
  }
  __syncthreads();
  // chain: T_i = T_parent * M_i  (parents first; one thread, J <= 128)
  if (tid == 0) {
    for (int i = 0; i < J; ++i) {
      const int pn = a.parent_node[i];
      const float* M = sM + 12 * i;
      float* T = sT + 12 * i;
      if (pn < 0) {
#pragma unroll
        for (int c = 0; c < 12; ++c) T[c] = M[c];
      } else {
        const float* P = sT + 12 * pn;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            T[4 * r + c] = P[4 * r] * M[c] + P[4 * r + 1] * M[4 + c] + P[4 * r + 2] * M[8 + c] + (c == 3 ? P[4 * r + 3] : 0.f);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < J * 16; i += blockDim.x) {
    const int j = i >> 4, c = i & 15;
    bone_T[i] = c < 12 ? sT[12 * j + c] : (c == 15 ? 1.f : 0.f);
  }
  if (saved) {
    float* s = saved;
    for (int i = tid; i < 4 * POSE_H; i += blockDim.x) s[i] = sh[1 + i / POSE_H][i % POSE_H];
    s += 4 * POSE_H;
    for (int i = tid; i < (J + 1) * 4; i += blockDim.x) s[i] = sP[i];
    s += (J + 1) * 4;
    for (int i = tid; i < J * 9; i += blockDim.x) s[i] = sR[i];
    s += J * 9;
    for (int i = tid; i < J * 12; i += blockDim.x) s[i] = sM[i];
    s += J * 12;
    for (int i = tid; i < J * 12; i += blockDim.x) s[i] = sT[i];
  }
}

struct PoseGrads {
  const float* d_bone_T;      // (J,16)
  const float* d_global_t;    // (3) or NULL
  const float* d_thetas;      // (J) or NULL
  float* d_w[5];
  float* d_b[4];
  float* d_joints;            // (J,3)
};

// The tree part (chain, pivots, Rodrigues: a few thousand flops) is evaluated redundantly by every CTA of the cluster, so
// that all of them hold dP; the MLP part splits every layer's rows over the CTAs: each CTA writes its rows of dW_l and a
// partial W_l^T d over its rows, the partials are all-gathered through distributed shared memory and summed in rank order.
__global__ void __cluster_dims__(POSE_CLUSTER, 1, 1) __launch_bounds__(POSE_THREADS)
pose_bwd_kernel(const PoseArgs a, const float* __restrict__ saved, const PoseGrads g) {
  __shared__ float sdT[POSE_MAX_J * 12];
  __shared__ float sdR[POSE_MAX_J * 9];      // gradient on the RAW rotations (after the sibling gather)
  __shared__ float sdJ[POSE_MAX_J * 3];
  __shared__ float sdP[(POSE_MAX_J + 1) * 4];
  __shared__ float sd[2][POSE_H];
  __shared__ float sPart[2][POSE_CLUSTER][POSE_H];   // [layer parity][source rank][column]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int J = a.J, tid = threadIdx.x;
  cluster.sync();                                   // every CTA's shared memory exists before the first remote store
  const float* h = saved;                        // h[l-1] = activation of hidden layer l (1..4)
  const float* sP = saved + 4 * POSE_H;
  const float* sR = sP + (J + 1) * 4;
  const float* sM = sR + J * 9;
  const float* sT = sM + J * 12;
  for (int i = tid; i < J * 12; i += blockDim.x) sdT[i] = g.d_bone_T[(i / 12) * 16 + (i % 12)];
  for (int i = tid; i < J * 9; i += blockDim.x) sdR[i] = 0.f;
  for (int i = tid; i < J * 3; i += blockDim.x) sdJ[i] = 0.f;
  __syncthreads();
  // chain backward, children first:  dT_p += [dA R^T + da t^T | da],  dM = [A_p^T dA | A_p^T da]
  // (sdT[i] is overwritten by dM_i once node i is processed)
  if (tid == 0) {
    for (int i = J - 1; i >= 0; --i) {
      const int pn = a.parent_node[i];
      if (pn < 0) continue;                      // T_root = M_root
      const float* M = sM + 12 * i;
      const float* P = sT + 12 * pn;
      float* dT = sdT + 12 * i;
      float* dP = sdT + 12 * pn;
      float dM[12];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          // dA_p[r][c] += sum_k dA[r][k] R[c][k] + da[r] t[c]
          dP[4 * r + c] += dT[4 * r] * M[4 * c] + dT[4 * r + 1] * M[4 * c + 1] + dT[4 * r + 2] * M[4 * c + 2] + dT[4 * r + 3] * M[4 * c + 3];
        }
        dP[4 * r + 3] += dT[4 * r + 3];
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) dM[4 * r + c] = P[r] * dT[c] + P[4 + r] * dT[4 + c] + P[8 + r] * dT[8 + c];
#pragma unroll
      for (int c = 0; c < 12; ++c) dT[c] = dM[c];
    }
  }
  __syncthreads();
  // M = [R | p - R p]  ->  dR (to the raw rotation of the sibling source), d pivot joint
  if (tid < J) {
    const float* dM = sdT + 12 * tid;
    const float* M = sM + 12 * tid;
    const int pj = a.pivot[tid];
    const float p[3] = {a.joints[3 * pj], a.joints[3 * pj + 1], a.joints[3 * pj + 2]};
    const float dt[3] = {dM[3], dM[7], dM[11]};
    // dp = dt - R^T dt
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&sdJ[3 * pj + c], dt[c] - (M[c] * dt[0] + M[4 + c] * dt[1] + M[8 + c] * dt[2]));
    const bool frozen = a.rot_mask && a.rot_mask[tid];
    if (!frozen) {
      const int src = a.sibling[tid];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(&sdR[9 * src + 3 * r + c], dM[4 * r + c] - dt[r] * p[c]);
    }
  }
  __syncthreads();
  // Rodrigues backward
  if (tid < J) {
    float R[9], n[3], rl, cs, sn;
    const float* pp = sP + 4 * tid;
    const float p4[4] = {pp[0], pp[1], pp[2], pp[3]};
    rodrigues4(p4, R, n, rl, cs, sn);
    const float* G = sdR + 9 * tid;
    const float x = n[0], y = n[1], z = n[2], C = 1.f - cs;
    const float Gn[3] = {G[0] * x + G[1] * y + G[2] * z, G[3] * x + G[4] * y + G[5] * z, G[6] * x + G[7] * y + G[8] * z};
    const float GTn[3] = {G[0] * x + G[3] * y + G[6] * z, G[1] * x + G[4] * y + G[7] * z, G[2] * x + G[5] * y + G[8] * z};
    const float gc = (G[0] + G[4] + G[8]) - (x * Gn[0] + y * Gn[1] + z * Gn[2]);
    const float gs = -z * G[1] + y * G[2] + z * G[3] - x * G[5] - y * G[6] + x * G[7];
    float dn[3];
    dn[0] = C * (Gn[0] + GTn[0]) + sn * (G[7] - G[5]);
    dn[1] = C * (Gn[1] + GTn[1]) + sn * (G[2] - G[6]);
    dn[2] = C * (Gn[2] + GTn[2]) + sn * (G[3] - G[1]);
    const float ndn = x * dn[0] + y * dn[1] + z * dn[2];
    sdP[4 * tid] = (dn[0] - x * ndn) / rl;
    sdP[4 * tid + 1] = (dn[1] - y * ndn) / rl;
    sdP[4 * tid + 2] = (dn[2] - z * ndn) / rl;
    sdP[4 * tid + 3] = -sn * gc + cs * gs + (g.d_thetas ? g.d_thetas[tid] : 0.f);
  }
  if (tid < 4) sdP[4 * J + tid] = (tid < 3 && g.d_global_t) ? g.d_global_t[tid] : 0.f;
  if (rank == 0)
    for (int i = tid; i < J * 3; i += blockDim.x) g.d_joints[i] = sdJ[i];
  __syncthreads();
  // all-gather of this CTA's partial column sums + rank-ordered reduction: -> full vector in every CTA
  auto allreduce_cols = [&](int par, float partial) -> float {
#pragma unroll
    for (int r = 0; r < POSE_CLUSTER; ++r) cluster.map_shared_rank(&sPart[par][rank][0], r)[tid] = partial;
    cluster.sync();
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < POSE_CLUSTER; ++r) s += sPart[par][r][tid];
    return s;
  };
  // output layer (no bias): dW4[r,:] = dP[r] h4 ; dh4 = W4^T dP, rows split over the cluster
  const int n_out = (J + 1) * 4;
  {
    const int per = (n_out + POSE_CLUSTER - 1) / POSE_CLUSTER;
    const int r0 = rank * per, r1 = min(n_out, r0 + per);
    for (int i = r0 * POSE_H + tid; i < r1 * POSE_H; i += blockDim.x) g.d_w[4][i] = sdP[i / POSE_H] * h[3 * POSE_H + (i % POSE_H)];
    float s = 0.f;
#pragma unroll 8
    for (int r = r0; r < r1; ++r) s = fmaf(__ldg(a.w[4] + (size_t)r * POSE_H + tid), sdP[r], s);
    s = allreduce_cols(0, s);
    sd[0][tid] = h[3 * POSE_H + tid] > 0.f ? s : 0.f;
  }
  __syncthreads();
  int cur = 0;
  constexpr int ROWS = POSE_H / POSE_CLUSTER;      // 32 rows of every hidden layer per CTA
  for (int l = 3; l >= 1; --l) {                 // hidden layers 3..1: input h[l-1] (activation of layer l)
    const float* x = h + (size_t)(l - 1) * POSE_H;
    const int n0 = rank * ROWS;
    if (tid >= n0 && tid < n0 + ROWS) g.d_b[l][tid] = sd[cur][tid];
    for (int i = n0 * POSE_H + tid; i < (n0 + ROWS) * POSE_H; i += blockDim.x) g.d_w[l][i] = sd[cur][i / POSE_H] * x[i % POSE_H];
    float s = 0.f;
#pragma unroll 32
    for (int n = n0; n < n0 + ROWS; ++n) s = fmaf(__ldg(a.w[l] + (size_t)n * POSE_H + tid), sd[cur][n], s);
    s = allreduce_cols((4 - l) & 1, s);
    sd[cur ^ 1][tid] = x[tid] > 0.f ? s : 0.f;
    __syncthreads();
    cur ^= 1;
  }
  {
    const int n0 = rank * ROWS;
    if (tid >= n0 && tid < n0 + ROWS) g.d_b[0][tid] = sd[cur][tid];
    for (int i = n0 * a.t_dim + tid; i < (n0 + ROWS) * a.t_dim; i += blockDim.x) g.d_w[0][i] = sd[cur][i / a.t_dim] * a.t_embed[i % a.t_dim];
  }
  cluster.sync();                                   // no CTA leaves while its shared memory can still be written remotely
}

static int pose_check(const PoseArgs& a) {
  APN_CHECK_ARG(a.J > 0 && a.J <= POSE_MAX_J && a.t_dim > 0 && a.t_dim <= POSE_MAX_T, "need J <= 128 and t_dim <= 64");
  APN_CHECK_ARG(a.t_embed && a.joints && a.parent_node && a.pivot && a.sibling, "null pointer");
  for (int l = 0; l < 5; ++l) APN_CHECK_ARG(a.w[l], "null weight");
  for (int l = 0; l < 4; ++l) APN_CHECK_ARG(a.b[l], "null bias");
  return 0;
}

extern "C" size_t apn_pose_saved_bytes(int J) { return sizeof(float) * (size_t)pose_saved_floats(J); }

extern "C" int apn_pose_fwd(const float* t_embed, int t_dim, const float* const* w5, const float* const* b4, const float* joints,
                            const int32_t* parent_node, const int32_t* pivot, const int32_t* sibling, const uint8_t* rot_mask,
                            int J, float* bone_T, float* global_t, float* thetas, void* saved, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  PoseArgs a;
  a.t_embed = t_embed; a.joints = joints; a.parent_node = parent_node; a.pivot = pivot; a.sibling = sibling; a.rot_mask = rot_mask;
  a.J = J; a.t_dim = t_dim;
  APN_CHECK_ARG(w5 && b4, "null pointer");
  for (int l = 0; l < 5; ++l) a.w[l] = w5[l];
  for (int l = 0; l < 4; ++l) a.b[l] = b4[l];
  if (pose_check(a)) return -1;
  APN_CHECK_ARG(bone_T && global_t && thetas, "null output");
  pose_fwd_kernel<<<POSE_CLUSTER, POSE_THREADS, 0, st>>>(a, bone_T, global_t, thetas, (float*)saved);
  APN_LAUNCH_CHECK();
  return 0;
}

extern "C" int apn_pose_bwd(const float* t_embed, int t_dim, const float* const* w5, const float* const* b4, const float* joints,
                            const int32_t* parent_node, const int32_t* pivot, const int32_t* sibling, const uint8_t* rot_mask,
                            int J, const void* saved, const float* d_bone_T, const float* d_global_t, const float* d_thetas,
                            float* const* d_w5, float* const* d_b4, float* d_joints, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  PoseArgs a;
  a.t_embed = t_embed; a.joints = joints; a.parent_node = parent_node; a.pivot = pivot; a.sibling = sibling; a.rot_mask = rot_mask;
  a.J = J; a.t_dim = t_dim;
  APN_CHECK_ARG(w5 && b4 && d_w5 && d_b4, "null pointer");
  for (int l = 0; l < 5; ++l) a.w[l] = w5[l];
  for (int l = 0; l < 4; ++l) a.b[l] = b4[l];
  if (pose_check(a)) return -1;
  APN_CHECK_ARG(saved && d_bone_T && d_joints, "null pointer");
  PoseGrads g;
  g.d_bone_T = d_bone_T; g.d_global_t = d_global_t; g.d_thetas = d_thetas; g.d_joints = d_joints;
  for (int l = 0; l < 5; ++l) {
    APN_CHECK_ARG(d_w5[l], "null weight gradient");
    g.d_w[l] = d_w5[l];
  }
  for (int l = 0; l < 4; ++l) {
    APN_CHECK_ARG(d_b4[l], "null bias gradient");
    g.d_b[l] = d_b4[l];
  }
  pose_bwd_kernel<<<POSE_CLUSTER, POSE_THREADS, 0, st>>>(a, (const float*)saved, g);
  APN_LAUNCH_CHECK();
  return 0;
}
