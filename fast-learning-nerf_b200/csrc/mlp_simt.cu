// mlp_simt.cu -- FLNERF_MODE_FP32: the NeRF MLP (model.py:38-63) forward and backward as fp32
// CUDA-core GEMMs with fused bias / ReLU / ReLU-mask epilogues.  This is the parity path: fp32
// inputs, fp32 FMA accumulation, same operand precision as the reference's cuBLAS SGEMM (TF32 off).
// The throughput path is mlp_tc.cu (tcgen05).
//
// One generic tiled kernel: C[M,N] (+)= A[M,K] * B[K,N] with arbitrary element strides for A and B
// (so X*W^T, dY*W and dY^T*X are the same kernel), 128x128x8 tiles, 8x8 register micro-tiles,
// optional split-K with atomicAdd for the weight gradients whose reduction dimension is the sample
// count.
#include "common.cuh"
#include "mlp_layout.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, NT = 256;

struct GemmArgs {
  const float *A; int64_t a_rs, a_cs;
  const float *B; int64_t b_rs, b_cs;
  float *C; int64_t ldc;
  int64_t M; int N; int64_t K;
  const float *bias;
  const float *mask; int64_t ldm;
  int relu, accumulate, atomic;
  int64_t k_split;
};

__global__ void __launch_bounds__(NT) gemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * g.k_split;
  const int64_t k_end = min(g.K, k_begin + g.k_split);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const bool a_mfast = (g.a_rs == 1 && g.a_cs != 1);
  const bool b_nfast = (g.b_cs == 1);

  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (a_mfast) { m = tid & 127; k = (tid >> 7) + 2 * i; } else { k = tid & 7; m = (tid >> 3) + 32 * i; }
      int64_t gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < g.M && gk < k_end) ? g.A[gm * g.a_rs + gk * g.a_cs] : 0.f;
      int n, kb;
      if (b_nfast) { n = tid & 127; kb = (tid >> 7) + 2 * i; } else { kb = tid & 7; n = (tid >> 3) + 32 * i; }
      int gn = n0 + n;
      int64_t gkb = k0 + kb;
      Bs[kb][n] = (gn < g.N && gkb < k_end) ? g.B[gkb * g.b_rs + (int64_t)gn * g.b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = acc[i][j];
      float *c = g.C + m * g.ldc + n;
      if (g.atomic) {
        atomicAdd(c, v);
      } else {
        if (g.accumulate) v += *c;
        if (g.bias) v += g.bias[n];
        if (g.relu) v = fmaxf(v, 0.f);
        if (g.mask) v = (g.mask[m * g.ldm + n] > 0.f) ? v : 0.f;
        *c = v;
      }
    }
  }
}

// out[c] += sum_r X[r*ld + c], c < N <= 256
__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ X, int64_t rows, int64_t ld, int N,
                                                     int64_t rows_per_block, float *__restrict__ out) {
  int c = threadIdx.x;
  if (c >= N) return;
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += X[r * ld + c];
  atomicAdd(out + c, s);
}

int launch_gemm(const GemmArgs &g0, cudaStream_t st) {
  GemmArgs g = g0;
  unsigned gz = 1;
  if (g.atomic) {
    g.k_split = 4096;
    gz = (unsigned)ceil_div64(g.K, g.k_split);
  } else {
    g.k_split = g.K;
  }
  dim3 grid((unsigned)ceil_div64(g.M, BM), (unsigned)((g.N + BN - 1) / BN), gz);
  FL_LAUNCH(gemm_kernel, grid, NT, 0, st, g);
  return 0;
}

int launch_colsum(const float *X, int64_t rows, int64_t ld, int N, float *out, cudaStream_t st) {
  const int64_t rpb = 2048;
  FL_LAUNCH(colsum_kernel, (unsigned)ceil_div64(rows, rpb), 256, 0, st, X, rows, ld, N, rpb, out);
  return 0;
}

// Y[n,N] = act(X[n,K] * W[N,K]^T (+ Y) + bias)
GemmArgs fwd_args(const float *X, int64_t ldx, const float *W, int64_t ldw, float *Y, int64_t ldy, int64_t n, int N,
                  int K, const float *bias, int relu, int accumulate) {
  GemmArgs g{};
  g.A = X; g.a_rs = ldx; g.a_cs = 1;
  g.B = W; g.b_rs = 1; g.b_cs = ldw;
  g.C = Y; g.ldc = ldy; g.M = n; g.N = N; g.K = K;
  g.bias = bias; g.relu = relu; g.accumulate = accumulate;
  return g;
}
// dX[n,K] = (dY[n,N] * W[N,K] (+ dX)) masked by (mask > 0)
GemmArgs dgrad_args(const float *dY, int64_t ldy, const float *W, int64_t ldw, float *dX, int64_t ldx, int64_t n,
                    int N_out, int K_in, const float *mask, int64_t ldm, int accumulate) {
  GemmArgs g{};
  g.A = dY; g.a_rs = ldy; g.a_cs = 1;
  g.B = W; g.b_rs = ldw; g.b_cs = 1;
  g.C = dX; g.ldc = ldx; g.M = n; g.N = K_in; g.K = N_out;
  g.mask = mask; g.ldm = ldm; g.accumulate = accumulate;
  return g;
}
// dW[N,K] += dY[n,N]^T * X[n,K]   (split-K atomics)
GemmArgs wgrad_args(const float *dY, int64_t ldy, const float *X, int64_t ldx, float *dW, int64_t ldw, int64_t n,
                    int N_out, int K_in) {
  GemmArgs g{};
  g.A = dY; g.a_rs = 1; g.a_cs = ldy;
  g.B = X; g.b_rs = ldx; g.b_cs = 1;
  g.C = dW; g.ldc = ldw; g.M = N_out; g.N = K_in; g.K = n;
  g.atomic = 1;
  return g;
}

#define RUN(x)             \
  do {                     \
    int rc_ = (x);         \
    if (rc_) return rc_;   \
  } while (0)

}  // namespace

// Parameter offsets (floats) for D=8, W=256, skip after layer 4, with in_pts position channels and in_views view
// channels, in the flat order of mlp_layout.h (pts_linears.0..7 {weight,bias}, views_linears.0, feature_linear,
// alpha_linear, rgb_linear).  in_pts = 63, in_views = 27 reproduces mlp_layout.h; in_pts = 84 is the nerf++ background
// network (nerf++-ours/nerf_network.py:70-118 has the same chain, see oracle/nerfpp_oracle.py).
struct SimtLayout {
  int in_pts, in_views;
  int w_pts[8], b_pts[8], k_pts[8];
  int w_views, b_views, w_feat, b_feat, w_alpha, b_alpha, w_rgb, b_rgb, total;
};

static SimtLayout make_layout(int in_pts, int in_views) {
  SimtLayout L{};
  L.in_pts = in_pts; L.in_views = in_views;
  int off = 0;
  for (int l = 0; l < 8; ++l) {
    L.k_pts[l] = (l == 0) ? in_pts : (l == 5 ? in_pts + 256 : 256);
    L.w_pts[l] = off; off += 256 * L.k_pts[l];
    L.b_pts[l] = off; off += 256;
  }
  L.w_views = off; off += 128 * (256 + in_views);
  L.b_views = off; off += 128;
  L.w_feat = off; off += 256 * 256;
  L.b_feat = off; off += 256;
  L.w_alpha = off; off += 256;
  L.b_alpha = off; off += 1;
  L.w_rgb = off; off += 3 * 128;
  L.b_rgb = off; off += 3;
  L.total = off;
  return L;
}

int64_t mlp_simt_param_count(int in_pts, int in_views) { return make_layout(in_pts, in_views).total; }

size_t mlp_simt_stash_bytes(int64_t n) { return (size_t)n * (8 * 256 + 256 + 128) * sizeof(float); }
size_t mlp_simt_bwd_workspace_bytes(int64_t n) { return (size_t)n * 2 * 256 * sizeof(float); }

// x: fp32 [n, in_pts + in_views] (the reference's embedded layout)
int mlp_simt_forward_g(int in_pts, int in_views, const float *P, int64_t n, const float *x, float *raw, float *stash,
                       cudaStream_t st) {
  const SimtLayout L = make_layout(in_pts, in_views);
  const int ldx = in_pts + in_views, kv = 256 + in_views;
  float *H[8];
  for (int l = 0; l < 8; ++l) H[l] = stash + (size_t)l * n * 256;
  float *F = stash + (size_t)8 * n * 256;
  float *H9 = F + (size_t)n * 256;
  RUN(launch_gemm(fwd_args(x, ldx, P + L.w_pts[0], in_pts, H[0], 256, n, 256, in_pts, P + L.b_pts[0], 1, 0), st));
  for (int l = 1; l < 8; ++l) {
    if (l == 5) {  // skip: input = [x_pts, h(256)] (model.py:47)
      RUN(launch_gemm(fwd_args(x, ldx, P + L.w_pts[5], L.k_pts[5], H[5], 256, n, 256, in_pts, nullptr, 0, 0), st));
      RUN(launch_gemm(fwd_args(H[4], 256, P + L.w_pts[5] + in_pts, L.k_pts[5], H[5], 256, n, 256, 256, P + L.b_pts[5], 1, 1), st));
    } else {
      RUN(launch_gemm(fwd_args(H[l - 1], 256, P + L.w_pts[l], 256, H[l], 256, n, 256, 256, P + L.b_pts[l], 1, 0), st));
    }
  }
  RUN(launch_gemm(fwd_args(H[7], 256, P + L.w_alpha, 256, raw + 3, 4, n, 1, 256, P + L.b_alpha, 0, 0), st));
  RUN(launch_gemm(fwd_args(H[7], 256, P + L.w_feat, 256, F, 256, n, 256, 256, P + L.b_feat, 0, 0), st));
  RUN(launch_gemm(fwd_args(F, 256, P + L.w_views, kv, H9, 128, n, 128, 256, nullptr, 0, 0), st));
  RUN(launch_gemm(fwd_args(x + in_pts, ldx, P + L.w_views + 256, kv, H9, 128, n, 128, in_views, P + L.b_views, 1, 1), st));
  RUN(launch_gemm(fwd_args(H9, 128, P + L.w_rgb, 128, raw, 4, n, 3, 128, P + L.b_rgb, 0, 0), st));
  return 0;
}

int mlp_simt_backward_g(int in_pts, int in_views, const float *P, int64_t n, const float *x, const float *stash,
                        const float *draw, float *G, float *ws, cudaStream_t st) {
  const SimtLayout L = make_layout(in_pts, in_views);
  const int ldx = in_pts + in_views, kv = 256 + in_views;
  const float *H[8];
  for (int l = 0; l < 8; ++l) H[l] = stash + (size_t)l * n * 256;
  const float *F = stash + (size_t)8 * n * 256;
  const float *H9 = F + (size_t)n * 256;
  float *bufA = ws, *bufB = ws + (size_t)n * 256;
  // rgb head
  RUN(launch_gemm(wgrad_args(draw, 4, H9, 128, G + L.w_rgb, 128, n, 3, 128), st));
  RUN(launch_colsum(draw, n, 4, 3, G + L.b_rgb, st));
  RUN(launch_colsum(draw + 3, n, 4, 1, G + L.b_alpha, st));
  float *G9 = bufA;  // [n,128]
  RUN(launch_gemm(dgrad_args(draw, 4, P + L.w_rgb, 128, G9, 128, n, 3, 128, H9, 128, 0), st));
  // views layer
  RUN(launch_gemm(wgrad_args(G9, 128, F, 256, G + L.w_views, kv, n, 128, 256), st));
  RUN(launch_gemm(wgrad_args(G9, 128, x + in_pts, ldx, G + L.w_views + 256, kv, n, 128, in_views), st));
  RUN(launch_colsum(G9, n, 128, 128, G + L.b_views, st));
  float *GF = bufB;  // d feature [n,256]
  RUN(launch_gemm(dgrad_args(G9, 128, P + L.w_views, kv, GF, 256, n, 128, 256, nullptr, 0, 0), st));
  // feature + alpha
  RUN(launch_gemm(wgrad_args(GF, 256, H[7], 256, G + L.w_feat, 256, n, 256, 256), st));
  RUN(launch_colsum(GF, n, 256, 256, G + L.b_feat, st));
  RUN(launch_gemm(wgrad_args(draw + 3, 4, H[7], 256, G + L.w_alpha, 256, n, 1, 256), st));
  float *cur = bufA;  // dH7 (pre-activation gradient of layer 7)
  RUN(launch_gemm(dgrad_args(GF, 256, P + L.w_feat, 256, cur, 256, n, 256, 256, nullptr, 0, 0), st));
  RUN(launch_gemm(dgrad_args(draw + 3, 4, P + L.w_alpha, 256, cur, 256, n, 1, 256, H[7], 256, 1), st));
  float *other = bufB;
  for (int l = 7; l >= 1; --l) {
    RUN(launch_colsum(cur, n, 256, 256, G + L.b_pts[l], st));
    if (l == 5) {
      RUN(launch_gemm(wgrad_args(cur, 256, x, ldx, G + L.w_pts[5], L.k_pts[5], n, 256, in_pts), st));
      RUN(launch_gemm(wgrad_args(cur, 256, H[4], 256, G + L.w_pts[5] + in_pts, L.k_pts[5], n, 256, 256), st));
      RUN(launch_gemm(dgrad_args(cur, 256, P + L.w_pts[5] + in_pts, L.k_pts[5], other, 256, n, 256, 256, H[4], 256, 0), st));
    } else {
      RUN(launch_gemm(wgrad_args(cur, 256, H[l - 1], 256, G + L.w_pts[l], 256, n, 256, 256), st));
      RUN(launch_gemm(dgrad_args(cur, 256, P + L.w_pts[l], 256, other, 256, n, 256, 256, H[l - 1], 256, 0), st));
    }
    float *t = cur; cur = other; other = t;
  }
  RUN(launch_colsum(cur, n, 256, 256, G + L.b_pts[0], st));
  RUN(launch_gemm(wgrad_args(cur, 256, x, ldx, G + L.w_pts[0], in_pts, n, 256, in_pts), st));
  return 0;
}

// the nerf-ours network: 63 position + 27 view channels (offsets identical to mlp_layout.h, checked once)
int mlp_simt_forward(const float *P, int64_t n, const float *x90, float *raw, float *stash, cudaStream_t st) {
  return mlp_simt_forward_g(63, 27, P, n, x90, raw, stash, st);
}

int mlp_simt_backward(const float *P, int64_t n, const float *x90, const float *stash, const float *draw, float *G,
                      float *ws, cudaStream_t st) {
  return mlp_simt_backward_g(63, 27, P, n, x90, stash, draw, G, ws, st);
}

int mlp_simt_layout_selfcheck() {
  using namespace mlp_layout;
  const SimtLayout L = make_layout(63, 27);
  bool ok = L.total == TOTAL && L.w_views == W_VIEWS && L.b_views == B_VIEWS && L.w_feat == W_FEAT && L.b_feat == B_FEAT &&
            L.w_alpha == W_ALPHA && L.b_alpha == B_ALPHA && L.w_rgb == W_RGB && L.b_rgb == B_RGB;
  for (int l = 0; l < 8; ++l) ok = ok && L.w_pts[l] == W_PTS[l] && L.b_pts[l] == B_PTS[l] && L.k_pts[l] == K_PTS[l];
  return ok ? 0 : 1;
}
