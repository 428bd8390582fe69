// Stage-2 regulariser losses and their gradients as two launches (run.py:633-657 adds them to every training iteration).
//
//   point_reg_kernel   over the canonical points and their static neighbourhood nn_i (N, K):
//       ARAP       lib/temporalpoints.py:723-725   sum_ik | D0_ik - sqrt(|x_i - x_n|^2 + eps) |          -> d_xyz (N,3)
//       weight TV  lib/temporalpoints.py:714-716   mean_ikj | w_ij - w_nj |                               -> d_w   (N,J)
//       sparsity   lib/temporalpoints.py:718-721   -mean_ij [ w log(w+eps) + (1-w) log(1-w+eps) ]         -> d_w   (N,J)
//     The reference evaluates each with ~6 framework launches over (N,K,3) / (N,K,J) temporaries plus as many in
//     autograd's backward; here one warp owns one point, reads its K neighbour rows once and emits loss and gradient
//     together.  d_xyz feeds apn_lbs_bwd's d_xyz, d_w its d_w (gradient of the MERGED skinning weights).
//   pose_reg_kernel    transformation regulariser lib/temporalpoints.py:797-800 ((sum|t| + sum|theta|) / J) -> d_thetas,
//     d_global_t; joint chamfer lib/temporalpoints.py:731-733 (sum_j min_s |joint_j - skeleton_s|^2, nearest by the k-NN
//     contract: (dx*dx + dy*dy) + dz*dz without FMA, ties -> lowest index) -> d_joints.
// abs() differentiates to sign() with sign(0) = 0, as torch does.
#include "common.cuh"

#define REG_WARPS 8

__device__ __forceinline__ float sgnf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

__global__ void __launch_bounds__(32 * REG_WARPS)
point_reg_kernel(const float* __restrict__ xyz, const float* __restrict__ w, const int32_t* __restrict__ nn_i,
                 const float* __restrict__ nn_dist, int N, int K, int J, float eps, float c_arap, float c_tv, float c_sp,
                 float* __restrict__ d_xyz, float* __restrict__ d_w, float* __restrict__ losses) {
  __shared__ float sred[3][REG_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = gridDim.x * REG_WARPS;
  float l_arap = 0.f, l_tv = 0.f, l_sp = 0.f;
  for (int i = blockIdx.x * REG_WARPS + warp; i < N; i += n_warps) {
    // neighbour k of this point lives in lane k (K <= 32)
    const int nk = lane < K ? __ldg(nn_i + (size_t)i * K + lane) : -1;
    if (c_arap != 0.f) {
      const float xi = __ldg(xyz + 3 * (size_t)i), yi = __ldg(xyz + 3 * (size_t)i + 1), zi = __ldg(xyz + 3 * (size_t)i + 2);
      float gx = 0.f, gy = 0.f, gz = 0.f;
      if (nk >= 0) {
        const float dx = xi - __ldg(xyz + 3 * (size_t)nk), dy = yi - __ldg(xyz + 3 * (size_t)nk + 1),
                    dz = zi - __ldg(xyz + 3 * (size_t)nk + 2);
        const float d = sqrtf(dx * dx + dy * dy + dz * dz + eps);
        const float r = __ldg(nn_dist + (size_t)i * K + lane) - d;
        l_arap += fabsf(r);
        const float s = -c_arap * sgnf(r) / d;      // d|D0 - d| / dd = -sign(D0 - d);  dd / dx_i = (x_i - x_n) / d
        gx = s * dx; gy = s * dy; gz = s * dz;
        if (s != 0.f && nk != i) {
          atomicAdd(d_xyz + 3 * (size_t)nk, -gx);
          atomicAdd(d_xyz + 3 * (size_t)nk + 1, -gy);
          atomicAdd(d_xyz + 3 * (size_t)nk + 2, -gz);
        } else {
          gx = gy = gz = 0.f;                       // a point that is its own neighbour: the two contributions cancel
        }
      }
      gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
      if (lane == 0 && (gx != 0.f || gy != 0.f || gz != 0.f)) {
        atomicAdd(d_xyz + 3 * (size_t)i, gx);
        atomicAdd(d_xyz + 3 * (size_t)i + 1, gy);
        atomicAdd(d_xyz + 3 * (size_t)i + 2, gz);
      }
    }
    if (c_tv != 0.f || c_sp != 0.f) {
      for (int j0 = 0; j0 < J; j0 += 32) {
        const int j = j0 + lane;
        const bool on = j < J;
        const float wi = on ? __ldg(w + (size_t)i * J + j) : 0.f;
        float g = 0.f;
        if (c_sp != 0.f && on) {
          const float a = wi + eps, b = 1.f - wi + eps;
          const float la = logf(a), lb = logf(b);
          l_sp -= wi * la + (1.f - wi) * lb;
          g = -c_sp * (la + wi / a - lb - (1.f - wi) / b);
        }
        if (c_tv != 0.f) {
          for (int k = 0; k < K; ++k) {
            const int n = __shfl_sync(0xffffffffu, nk, k);
            if (!on || n == i) continue;
            const float df = wi - __ldg(w + (size_t)n * J + j);
            l_tv += fabsf(df);
            const float s = c_tv * sgnf(df);
            if (s != 0.f) {
              g += s;
              atomicAdd(d_w + (size_t)n * J + j, -s);
            }
          }
        }
        if (on && g != 0.f) atomicAdd(d_w + (size_t)i * J + j, g);
      }
    }
  }
  l_arap = warp_sum(l_arap); l_tv = warp_sum(l_tv); l_sp = warp_sum(l_sp);
  if (lane == 0) { sred[0][warp] = l_arap; sred[1][warp] = l_tv; sred[2][warp] = l_sp; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float v = 0.f;
    for (int q = 0; q < REG_WARPS; ++q) v += sred[threadIdx.x][q];
    if (v != 0.f) atomicAdd(losses + threadIdx.x, v);
  }
}

__global__ void point_reg_scale_kernel(float* __restrict__ losses, float s_arap, float s_tv, float s_sp) {
  if (threadIdx.x == 0) losses[0] *= s_arap;
  if (threadIdx.x == 1) losses[1] *= s_tv;
  if (threadIdx.x == 2) losses[2] *= s_sp;
}

extern "C" int apn_point_regularisers(const float* xyz, const float* w, const int32_t* nn_i, const float* nn_dist, int N, int K,
                                      int J, float eps, float weight_arap, float weight_tv, float weight_sparsity, float* d_xyz,
                                      float* d_w, float* losses3, apn_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  APN_CHECK_ARG(N > 0 && K > 0 && K <= 32 && J > 0, "need N > 0, 0 < K <= 32, J > 0");
  APN_CHECK_ARG(nn_i && losses3, "null pointer");
  APN_CHECK_ARG(weight_arap == 0.f || (xyz && nn_dist && d_xyz), "ARAP needs xyz, nn_dist and d_xyz");
  APN_CHECK_ARG((weight_tv == 0.f && weight_sparsity == 0.f) || (w && d_w), "the weight losses need w and d_w");
  APN_CUDA(cudaMemsetAsync(losses3, 0, 3 * sizeof(float), st));
  if (d_w) APN_CUDA(cudaMemsetAsync(d_w, 0, (size_t)N * J * sizeof(float), st));
  // gradient coefficients carry the mean's 1 / count; the loss sums are scaled once at the end
  const float c_tv = weight_tv / ((float)N * (float)K * (float)J), c_sp = weight_sparsity / ((float)N * (float)J);
  const int blocks = min(apn_div_up(N, REG_WARPS), APN_SM_COUNT * 8);
  point_reg_kernel<<<blocks, 32 * REG_WARPS, 0, st>>>(xyz, w, nn_i, nn_dist, N, K, J, eps, weight_arap, c_tv, c_sp, d_xyz, d_w, losses3);
  APN_LAUNCH_CHECK();
  point_reg_scale_kernel<<<1, 32, 0, st>>>(losses3, weight_arap, c_tv, c_sp);
  APN_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(128)
pose_reg_kernel(const float* __restrict__ thetas, const float* __restrict__ global_t, const float* __restrict__ joints,
                const float* __restrict__ skeleton, int J, int S, float c_treg, float c_jc, float* __restrict__ d_thetas,
                float* __restrict__ d_global_t, float* __restrict__ d_joints, float* __restrict__ losses) {
  __shared__ float sred[2][4];
  float l_t = 0.f, l_j = 0.f;
  const float ct = c_treg / (float)J;          // len(thetas + 1) == J (lib/temporalpoints.py:800)
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    if (d_thetas) {
      const float th = thetas[j];
      l_t += fabsf(th);
      d_thetas[j] = ct * sgnf(th);
    }
    if (d_joints) {
      const float qx = joints[3 * j], qy = joints[3 * j + 1], qz = joints[3 * j + 2];
      float best = INFINITY;
      int bs = 0;
      for (int s = 0; s < S; ++s) {
        const float d = dist2_contract(qx, qy, qz, __ldg(skeleton + 3 * s), __ldg(skeleton + 3 * s + 1), __ldg(skeleton + 3 * s + 2));
        if (d < best) { best = d; bs = s; }
      }
      const float dx = qx - skeleton[3 * bs], dy = qy - skeleton[3 * bs + 1], dz = qz - skeleton[3 * bs + 2];
      l_j += dx * dx + dy * dy + dz * dz;
      d_joints[3 * j] = 2.f * c_jc * dx;
      d_joints[3 * j + 1] = 2.f * c_jc * dy;
      d_joints[3 * j + 2] = 2.f * c_jc * dz;
    }
  }
  if (d_thetas && threadIdx.x < 3) {
    const float tv = global_t[threadIdx.x];
    l_t += fabsf(tv);
    if (d_global_t) d_global_t[threadIdx.x] = ct * sgnf(tv);
  }
  l_t = warp_sum(l_t); l_j = warp_sum(l_j);
  if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = l_t; sred[1][threadIdx.x >> 5] = l_j; }
  __syncthreads();
  if (threadIdx.x == 0) {
    losses[0] = ct * (sred[0][0] + sred[0][1] + sred[0][2] + sred[0][3]);
    losses[1] = c_jc * (sred[1][0] + sred[1][1] + sred[1][2] + sred[1][3]);
  }
}

extern "C" int apn_pose_regularisers(const float* thetas, const float* global_t, const float* joints, const float* skeleton, int J,
                                     int S, float weight_transformation_reg, float weight_joint_chamfer, float* d_thetas,
                                     float* d_global_t, float* d_joints, float* losses2, apn_stream_t stream_) {
  APN_CHECK_ARG(J > 0 && losses2, "need J > 0 and a loss buffer");
  APN_CHECK_ARG(!d_thetas || (thetas && global_t), "the transformation regulariser needs thetas and global_t");
  APN_CHECK_ARG(!d_joints || (joints && skeleton && S > 0), "the joint chamfer loss needs joints and a skeleton cloud");
  pose_reg_kernel<<<1, 128, 0, (cudaStream_t)stream_>>>(thetas, global_t, joints, skeleton, J, S, weight_transformation_reg,
                                                        weight_joint_chamfer, d_thetas, d_global_t, d_joints, losses2);
  APN_LAUNCH_CHECK();
  return 0;
}
