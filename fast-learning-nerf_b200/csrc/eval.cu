// eval.cu -- image metrics of the eval path (render.py:94-146: render_path reports PSNR and SSIM per rendered frame).
//   PSNR : -10 log10(mean((rgb - gt)^2))                                        (render.py:120, run_nerf_helpers.py:9-10)
//   SSIM : compute_ssim (run_nerf_helpers.py:158-228; modelled after tf.image.ssim): 11-tap Gaussian (sigma 1.5),
//          separable, ZERO padded (F.conv2d padding = 5), k1 = 0.01, k2 = 0.03, covariance clipped to sqrt(var0 var1),
//          mean of the per-pixel, per-channel map.
// One block = one 16x16 pixel tile of one channel; the two images' 26x26 halo regions are staged in shared memory, blurred
// along x into five moment planes (a, b, a^2, b^2, ab) and then along y; block sums leave through two double atomics.
#include "common.cuh"

namespace {

constexpr int kT = 16, kR = 5, kS = kT + 2 * kR;  // tile, filter radius, staged side (26)

__global__ void __launch_bounds__(kT * kT) ssim_psnr_kernel(int H, int W, const float *__restrict__ img0,
                                                            const float *__restrict__ img1, float c1, float c2,
                                                            double *__restrict__ out) {
  __shared__ float a[kS][kS], b[kS][kS];
  __shared__ float hb[5][kS][kT];
  __shared__ float taps[2 * kR + 1];
  __shared__ double red[2][kT * kT / 32];
  const int tx = threadIdx.x % kT, ty = threadIdx.x / kT;
  const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT, ch = blockIdx.z;
  if (threadIdx.x < 2 * kR + 1) {
    float s = 0.f;
    for (int i = 0; i < 2 * kR + 1; ++i) { const float d = (float)(i - kR) / 1.5f; s += expf(-0.5f * d * d); }
    const float d = (float)((int)threadIdx.x - kR) / 1.5f;
    taps[threadIdx.x] = expf(-0.5f * d * d) / s;
  }
  for (int i = threadIdx.x; i < kS * kS; i += kT * kT) {
    const int r = i / kS, c = i % kS;
    const int y = y0 + r - kR, x = x0 + c - kR;
    const bool in = y >= 0 && y < H && x >= 0 && x < W;
    const int64_t idx = ((int64_t)y * W + x) * 3 + ch;
    a[r][c] = in ? img0[idx] : 0.f;
    b[r][c] = in ? img1[idx] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kS * kT; i += kT * kT) {   // blur along x: rows 0..25, output columns 0..15
    const int r = i / kT, c = i % kT;
    float m0 = 0.f, m1 = 0.f, s00 = 0.f, s11 = 0.f, s01 = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kR + 1; ++k) {
      const float w = taps[k], va = a[r][c + k], vb = b[r][c + k];
      m0 = fmaf(w, va, m0); m1 = fmaf(w, vb, m1);
      s00 = fmaf(w, va * va, s00); s11 = fmaf(w, vb * vb, s11); s01 = fmaf(w, va * vb, s01);
    }
    hb[0][r][c] = m0; hb[1][r][c] = m1; hb[2][r][c] = s00; hb[3][r][c] = s11; hb[4][r][c] = s01;
  }
  __syncthreads();
  double ssim = 0.0, se = 0.0;
  const int y = y0 + ty, x = x0 + tx;
  if (y < H && x < W) {
    float m0 = 0.f, m1 = 0.f, s00 = 0.f, s11 = 0.f, s01 = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kR + 1; ++k) {
      const float w = taps[k];
      m0 = fmaf(w, hb[0][ty + k][tx], m0); m1 = fmaf(w, hb[1][ty + k][tx], m1);
      s00 = fmaf(w, hb[2][ty + k][tx], s00); s11 = fmaf(w, hb[3][ty + k][tx], s11); s01 = fmaf(w, hb[4][ty + k][tx], s01);
    }
    const float mu00 = m0 * m0, mu11 = m1 * m1, mu01 = m0 * m1;
    const float v0 = fmaxf(0.f, s00 - mu00), v1 = fmaxf(0.f, s11 - mu11);
    float cov = s01 - mu01;
    cov = copysignf(fminf(sqrtf(v0 * v1), fabsf(cov)), cov);
    if (cov == 0.f) cov = 0.f;   // torch.sign(0) * ... = 0
    ssim = (double)(((2.f * mu01 + c1) * (2.f * cov + c2)) / ((mu00 + mu11 + c1) * (v0 + v1 + c2)));
    const float d = a[ty + kR][tx + kR] - b[ty + kR][tx + kR];
    se = (double)(d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssim += __shfl_xor_sync(0xffffffffu, ssim, o);
    se += __shfl_xor_sync(0xffffffffu, se, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ssim; red[1][threadIdx.x >> 5] = se; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, e = 0.0;
    for (int i = 0; i < kT * kT / 32; ++i) { s += red[0][i]; e += red[1][i]; }
    atomicAdd(out, s);
    atomicAdd(out + 1, e);
  }
}

}  // namespace

extern "C" int flnerf_ssim_psnr(flnerf_ctx *ctx, int H, int W, const float *img0, const float *img1, double max_val,
                                double *sums, void *stream) {
  FL_REQUIRE(ctx && img0 && img1 && sums && H > 0 && W > 0, "flnerf_ssim_psnr: bad arguments");
  FL_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), (cudaStream_t)stream));
  const float c1 = (float)((0.01 * max_val) * (0.01 * max_val)), c2 = (float)((0.03 * max_val) * (0.03 * max_val));
  dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, 3);
  FL_LAUNCH(ssim_psnr_kernel, grid, kT * kT, 0, stream, H, W, img0, img1, c1, c2, sums);
  return 0;
}
