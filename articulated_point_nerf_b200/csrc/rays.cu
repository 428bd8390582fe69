// Camera rays on the device.  Replaces get_rays / get_rays_of_a_view (lib/tineuvox.py:675-738, mode='center', ndc=False):
// the reference builds (H, W, 3) origin / direction / unit-direction tensors with ~15 torch launches (on the host in the
// render loops of run.py:136-150, then copies them); a 1024^2 frame is 37.7 MB of rays for 100 bytes of camera.  Here a
// frame's rays are produced by one launch from (K, c2w) passed by value, for all pixels or for a list of pixel indices
// (a rank's image tiles / a training batch's random pixels).
// Arithmetic follows the reference expression by expression with separate IEEE operations (no FMA contraction):
//   dirs = [(i - cx)/fx, -(j - cy)/fy, -1]  (inverse_y: [(i - cx)/fx, (j - cy)/fy, 1]),  i = col + 0.5, j = row + 0.5
//   rays_d[k] = (dirs[0]*c2w[k][0] + dirs[1]*c2w[k][1]) + dirs[2]*c2w[k][2]      (torch.sum over the last axis)
//   viewdirs = rays_d / sqrt((rx*rx + ry*ry) + rz*rz)                            (rays_d.norm(dim=-1))
#include "common.cuh"

struct RayCam {
  float fx, fy, cx, cy;
  float c[3][4];      // c2w[:3, :4]
  int H, W, inverse_y, flip_x, flip_y;
};

__global__ void __launch_bounds__(256)
rays_kernel(const RayCam cam, const int32_t* __restrict__ pixel_ids, long long first, int n, float* __restrict__ rays_o,
            float* __restrict__ rays_d, float* __restrict__ viewdirs) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const long long p = pixel_ids ? (long long)__ldg(pixel_ids + q) : first + q;
  const int row = (int)(p / cam.W), col = (int)(p - (long long)row * cam.W);
  const float i = __fadd_rn((float)(cam.flip_x ? cam.W - 1 - col : col), 0.5f);
  const float j = __fadd_rn((float)(cam.flip_y ? cam.H - 1 - row : row), 0.5f);
  const float dx = __fdiv_rn(__fsub_rn(i, cam.cx), cam.fx);
  float dy = __fdiv_rn(__fsub_rn(j, cam.cy), cam.fy);
  float dz = 1.f;
  if (!cam.inverse_y) {
    dy = -dy;
    dz = -1.f;
  }
  float r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    r[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam.c[k][0]), __fmul_rn(dy, cam.c[k][1])), __fmul_rn(dz, cam.c[k][2]));
  const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], r[0]), __fmul_rn(r[1], r[1])), __fmul_rn(r[2], r[2])));
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    rays_o[3 * (size_t)q + k] = cam.c[k][3];
    rays_d[3 * (size_t)q + k] = r[k];
    if (viewdirs) viewdirs[3 * (size_t)q + k] = __fdiv_rn(r[k], nrm);
  }
}

extern "C" int apn_rays_of_a_view(const float* K_host9, const float* c2w_host, int c2w_rows, int H, int W, int inverse_y,
                                  int flip_x, int flip_y, const int32_t* pixel_ids, long long first_pixel, int n,
                                  float* rays_o, float* rays_d, float* viewdirs, apn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  APN_CHECK_ARG(K_host9 && c2w_host && rays_o && rays_d, "null pointer");
  APN_CHECK_ARG(H > 0 && W > 0 && n >= 0 && (c2w_rows == 3 || c2w_rows == 4), "bad sizes");
  APN_CHECK_ARG(pixel_ids || (first_pixel >= 0 && first_pixel + n <= (long long)H * W), "pixel range outside the image");
  if (n == 0) return 0;
  RayCam cam;
  cam.fx = K_host9[0]; cam.cx = K_host9[2]; cam.fy = K_host9[4]; cam.cy = K_host9[5];
  for (int k = 0; k < 3; ++k)
    for (int c = 0; c < 4; ++c) cam.c[k][c] = c2w_host[4 * k + c];
  cam.H = H; cam.W = W; cam.inverse_y = inverse_y; cam.flip_x = flip_x; cam.flip_y = flip_y;
  rays_kernel<<<apn_div_up(n, 256), 256, 0, stream>>>(cam, pixel_ids, first_pixel, n, rays_o, rays_d, viewdirs);
  APN_LAUNCH_CHECK();
  return 0;
}
