// composite.cu -- alpha compositing (forward/backward) and inverse-CDF resampling + sort-merge.
// Reference: nerf-ours/render.py:149-192 (raw2outputs), run_nerf_helpers.py:112-155 (sample_pdf),
// render.py:279-284,299 (merge, z_std).  Math restated in SURVEY.md appendix A.1-A.3.
//
// One warp per ray.  A ray's S samples are walked in chunks of 32 (sample = chunk*32 + lane) so that the
// raw[.,4] float4 loads are 512-byte coalesced; transmittance is a multiplicative warp scan with a carry
// between chunks; the backward pass is the mirrored suffix-sum scan.  HBM-bound: algorithmic bytes per ray
// are 20*S+12 read, 24 (+4*S weights) written in forward; 20*S+12+24 read, 16*S written in backward.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int kRaysPerBlock = 4;  // 128 threads

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive product scan across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
// inclusive sum scan across the warp
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// inclusive suffix-sum scan across the warp (lane i gets sum over lanes >= i)
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

struct SampleTerms {
  float alpha, om, dist, sig;
};

__device__ __forceinline__ SampleTerms sample_terms(float raw3, float noise, float z0, float z1, bool last, float nrm) {
  SampleTerms t;
  float d = last ? 1e10f : __fsub_rn(z1, z0);
  t.dist = __fmul_rn(d, nrm);
  t.sig = raw3 + noise;
  float s = fmaxf(t.sig, 0.0f);
  t.alpha = 1.0f - expf(-__fmul_rn(s, t.dist));       // render.py:162,181
  t.om = __fadd_rn(__fsub_rn(1.0f, t.alpha), 1e-10f);  // 1-alpha+1e-10 (render.py:183)
  return t;
}

__device__ __forceinline__ float ray_norm(const float *d) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
}

__global__ void __launch_bounds__(kRaysPerBlock * 32)
composite_fwd_kernel(int64_t B, int S, const float4 *__restrict__ raw, const float *__restrict__ z,
                     const float *__restrict__ rays_d, int64_t d_stride, const float *__restrict__ noise,
                     int white_bkgd, float *__restrict__ rgb, float *__restrict__ disp, float *__restrict__ acc,
                     float *__restrict__ depth, float *__restrict__ weights) {
  int lane = threadIdx.x & 31;
  int64_t ray = (int64_t)blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
  if (ray >= B) return;
  float nrm = ray_norm(rays_d + ray * d_stride);
  const float4 *rr = raw + ray * S;
  const float *zz = z + ray * S;
  float carry = 1.0f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    bool valid = i < S;
    float4 r = valid ? rr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float z0 = valid ? zz[i] : 0.f;
    float z1 = (i + 1 < S) ? zz[i + 1] : 0.f;
    float nz = (noise && valid) ? noise[ray * S + i] : 0.f;
    SampleTerms t = sample_terms(r.w, nz, z0, z1, i == S - 1, nrm);
    float om = valid ? t.om : 1.0f;
    float incl = warp_scan_mul(om, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    float T = carry * excl;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    float w = valid ? t.alpha * T : 0.f;
    if (weights && valid) weights[ray * S + i] = w;
    sr += w * sigmoidf_(r.x);
    sg += w * sigmoidf_(r.y);
    sb += w * sigmoidf_(r.z);
    sd += w * z0;
    sa += w;
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
  if (lane == 0) {
    float q = sd / sa;  // NaN when acc == 0; torch.max propagates it (render.py:187)
    float dsp = (q != q) ? q : 1.0f / fmaxf(1e-10f, q);
    if (white_bkgd) {
      float bg = 1.0f - sa;
      sr += bg; sg += bg; sb += bg;
    }
    rgb[ray * 3] = sr; rgb[ray * 3 + 1] = sg; rgb[ray * 3 + 2] = sb;
    if (disp) disp[ray] = dsp;
    if (acc) acc[ray] = sa;
    if (depth) depth[ray] = sd;
  }
}

// dynamic smem: per ray 2*S floats (alpha, T)
__global__ void __launch_bounds__(kRaysPerBlock * 32)
composite_bwd_kernel(int64_t B, int S, const float4 *__restrict__ raw, const float *__restrict__ z,
                     const float *__restrict__ rays_d, int64_t d_stride, const float *__restrict__ noise,
                     int white_bkgd, const float *__restrict__ g_rgb, const float *__restrict__ g_disp,
                     const float *__restrict__ g_acc, const float *__restrict__ g_depth, float4 *__restrict__ draw) {
  extern __shared__ float sm[];
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t ray = (int64_t)blockIdx.x * kRaysPerBlock + wib;
  if (ray >= B) return;
  float *s_alpha = sm + (size_t)wib * 2 * S;
  float *s_T = s_alpha + S;
  float nrm = ray_norm(rays_d + ray * d_stride);
  const float4 *rr = raw + ray * S;
  const float *zz = z + ray * S;
  // ---- pass 1: alpha_i, T_i and the totals A = sum w, D = sum w z
  float carry = 1.0f, sd = 0.f, sa = 0.f;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    bool valid = i < S;
    float r3 = valid ? rr[i].w : 0.f;
    float z0 = valid ? zz[i] : 0.f;
    float z1 = (i + 1 < S) ? zz[i + 1] : 0.f;
    float nz = (noise && valid) ? noise[ray * S + i] : 0.f;
    SampleTerms t = sample_terms(r3, nz, z0, z1, i == S - 1, nrm);
    float om = valid ? t.om : 1.0f;
    float incl = warp_scan_mul(om, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    float T = carry * excl;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    if (valid) {
      s_alpha[i] = t.alpha;
      s_T[i] = T;
      float w = t.alpha * T;
      sd += w * z0;
      sa += w;
    }
  }
  sd = warp_sum(sd); sa = warp_sum(sa);
  __syncwarp();
  float gc0 = g_rgb ? g_rgb[ray * 3] : 0.f, gc1 = g_rgb ? g_rgb[ray * 3 + 1] : 0.f, gc2 = g_rgb ? g_rgb[ray * 3 + 2] : 0.f;
  float gA = g_acc ? g_acc[ray] : 0.f, gD = g_depth ? g_depth[ray] : 0.f;
  if (g_disp) {  // disp = 1/max(1e-10, D/A)
    float q = sd / sa;
    if (q > 1e-10f) {
      float gq = -g_disp[ray] / (q * q);
      gD += gq / sa;
      gA += gq * (-sd / (sa * sa));
    }
  }
  if (white_bkgd) gA -= (gc0 + gc1 + gc2);
  // ---- pass 2 (reverse): suffix sums of q_k w_k
  float rcarry = 0.f;
  int nchunk = (S + 31) / 32;
  for (int c = nchunk - 1; c >= 0; --c) {
    int i = c * 32 + lane;
    bool valid = i < S;
    float4 r = valid ? rr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float z0 = valid ? zz[i] : 0.f;
    float z1 = (i + 1 < S) ? zz[i + 1] : 0.f;
    float nz = (noise && valid) ? noise[ray * S + i] : 0.f;
    float alpha = valid ? s_alpha[i] : 0.f, T = valid ? s_T[i] : 0.f;
    float c0 = sigmoidf_(r.x), c1 = sigmoidf_(r.y), c2 = sigmoidf_(r.z);
    float w = alpha * T;
    float q = gc0 * c0 + gc1 * c1 + gc2 * c2 + gA + gD * z0;
    float s = valid ? q * w : 0.f;
    float incl = warp_rscan_add(s, lane);
    float R = rcarry + (incl - s);  // exclusive suffix sum
    rcarry += __shfl_sync(0xffffffffu, incl, 0);
    if (valid) {
      float d = (i == S - 1) ? 1e10f : __fsub_rn(z1, z0);
      float dist = __fmul_rn(d, nrm);
      float om = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      float dalpha = q * T - R / om;
      float dsig = ((r.w + nz) > 0.f) ? dalpha * dist * (1.0f - alpha) : 0.f;
      draw[ray * S + i] = make_float4(w * gc0 * c0 * (1.f - c0), w * gc1 * c1 * (1.f - c1), w * gc2 * c2 * (1.f - c2), dsig);
    }
  }
}

// dynamic smem per warp: (Nc-1) cdf floats + P2 merge floats
// kBins = false: z = coarse depths [B,Nc], wts = compositing weights [B,Nc]; bins are the mid-points and the
//                 weights are wts[1:-1] (render.py:279-280); output merged + sorted with z.
// kBins = true : z = bins [B,Nc-1] and wts = weights [B,Nc-2] taken as they are (the generic sample_pdf API).
// kPP = 1: the nerf++ variant of the resampler (nerf++-ours/ddp_train_nerf.py:84-133): 1e-6 instead of 1e-5 floors, the
// inverse CDF counts u >= cdf[:M] (no clamp of the upper index), and the bin width carries a +1e-6
template <bool kBins, int kPP>
__global__ void __launch_bounds__(kRaysPerBlock * 32)
sample_pdf_merge_kernel(int64_t B, int Nc, int Nf, int P2, const float *__restrict__ z, const float *__restrict__ wts,
                        const float *__restrict__ u_in, int det, uint64_t seed, uint64_t offset,
                        float *__restrict__ z_merged, float *__restrict__ z_samples, float *__restrict__ z_std,
                        const flnerf_step_record *__restrict__ rec) {
  extern __shared__ float sm[];
  if (rec) offset += rec->rng_offset;
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t ray = (int64_t)blockIdx.x * kRaysPerBlock + wib;
  if (ray >= B) return;
  const int ncdf = Nc - 1, M = Nc - 2;
  float *cdf = sm + (size_t)wib * (ncdf + P2);
  float *buf = cdf + ncdf;
  const float *zz = z + ray * (kBins ? Nc - 1 : Nc);
  const float *ww = kBins ? wts + ray * (Nc - 2) - 1 : wts + ray * Nc;   // ww[k + 1] = k-th bin weight
  for (int i = lane; i < (kBins ? Nc - 1 : Nc); i += 32) buf[i] = zz[i];
  __syncwarp();
  // pdf = (w[1:-1] + 1e-5) / sum ; cdf = [0, cumsum(pdf)]      (helpers:114-117)
  float part = 0.f;
  const float eps = kPP ? 1e-6f : 1e-5f;
  for (int k = lane; k < M; k += 32) part += __fadd_rn(ww[k + 1], eps);
  float total = warp_sum(part);
  // torch.cumsum on the CPU accumulates fp32 inputs in double and rounds every prefix once; do the same so that the
  // searchsorted / den<1e-5 decisions below see the reference's cdf
  double carry = 0.0;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < M; base += 32) {
    int k = base + lane;
    double incl = (k < M) ? (double)__fdiv_rn(__fadd_rn(ww[k + 1], eps), total) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (k < M) cdf[k + 1] = (float)(carry + incl);
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  // inverse CDF (helpers:137-153)
  float step = 1.0f / (float)(Nf - 1);
  float s1 = 0.f;
  for (int j = lane; j < Nf; j += 32) {
    float u;
    if (det) {
      u = (j < Nf / 2) ? (float)j * step : 1.0f - (float)(Nf - 1 - j) * step;  // torch.linspace(0,1,Nf)
    } else if (u_in) {
      u = u_in[ray * Nf + j];
    } else {
      uint32_t r[4];
      philox4x32(seed, offset + (uint64_t)(ray * Nf + j), 0x5D0Full, r);
      u = u32_to_unit(r[0]);
    }
    int lo = 0, hi = kPP ? M : ncdf;  // searchsorted(right=True): count of cdf entries <= u (nerf++: among cdf[:M])
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    int below = max(0, lo - 1), above = min(ncdf - 1, lo);
    float cb = cdf[below], ca = cdf[above];
    float bb = kBins ? buf[below] : __fmul_rn(0.5f, __fadd_rn(buf[below + 1], buf[below]));
    float ba = kBins ? buf[above] : __fmul_rn(0.5f, __fadd_rn(buf[above + 1], buf[above]));
    float den = __fsub_rn(ca, cb);
    if (den < eps) den = 1.0f;
    float t = __fdiv_rn(__fsub_rn(u, cb), den);
    float smp = __fadd_rn(bb, __fmul_rn(t, kPP ? __fadd_rn(__fsub_rn(ba, bb), 1e-6f) : __fsub_rn(ba, bb)));
    if (!kBins) buf[Nc + j] = smp;
    if (z_samples) z_samples[ray * Nf + j] = smp;
    s1 += smp;
  }
  if (kBins) return;
  for (int i = Nc + Nf + lane; i < P2; i += 32) buf[i] = CUDART_INF_F;
  __syncwarp();
  if (z_std) {  // torch.std(unbiased=False): two-pass (render.py:299)
    float mean = warp_sum(s1) / (float)Nf;
    float s2 = 0.f;
    for (int j = lane; j < Nf; j += 32) {
      float dlt = buf[Nc + j] - mean;
      s2 += dlt * dlt;
    }
    s2 = warp_sum(s2);
    if (lane == 0) z_std[ray] = sqrtf(s2 / (float)Nf);
  }
  // bitonic sort of buf[0..P2) ascending (render.py:283); sorted values are unique as a multiset
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (P2 >> 1); t += 32) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int p = i | j;
        float a = buf[i], b = buf[p];
        bool up = (i & k) == 0;
        if ((a > b) == up) {
          buf[i] = b;
          buf[p] = a;
        }
      }
      __syncwarp();
    }
  }
  float *out = z_merged + ray * (Nc + Nf);
  for (int i = lane; i < Nc + Nf; i += 32) out[i] = buf[i];
}

// template instantiations behind plain names (a comma inside FL_LAUNCH's first argument would split it)
constexpr auto k_merge = sample_pdf_merge_kernel<false, 0>;
constexpr auto k_bins = sample_pdf_merge_kernel<true, 0>;
constexpr auto k_merge_pp = sample_pdf_merge_kernel<false, 1>;

}  // namespace

extern "C" {

int flnerf_composite_forward(flnerf_ctx *ctx, int64_t B, int S, const float *raw, const float *z, const float *rays_d,
                             int64_t rays_d_stride, const float *noise, int white_bkgd, float *rgb, float *disp,
                             float *acc, float *depth, float *weights, void *stream) {
  FL_REQUIRE(ctx && raw && z && rays_d && rgb && S > 0 && B >= 0, "flnerf_composite_forward: bad arguments");
  FL_REQUIRE(((uintptr_t)raw & 15) == 0, "flnerf_composite_forward: raw must be 16-byte aligned");
  if (B == 0) return 0;
  FL_LAUNCH(composite_fwd_kernel, (unsigned)ceil_div64(B, kRaysPerBlock), kRaysPerBlock * 32, 0, stream, B, S,
            (const float4 *)raw, z, rays_d, rays_d_stride, noise, white_bkgd, rgb, disp, acc, depth, weights);
  return 0;
}

int flnerf_composite_backward(flnerf_ctx *ctx, int64_t B, int S, const float *raw, const float *z, const float *rays_d,
                              int64_t rays_d_stride, const float *noise, int white_bkgd, const float *g_rgb,
                              const float *g_disp, const float *g_acc, const float *g_depth, float *draw,
                              void *stream) {
  FL_REQUIRE(ctx && raw && z && rays_d && draw && S > 0 && B >= 0, "flnerf_composite_backward: bad arguments");
  FL_REQUIRE((((uintptr_t)raw | (uintptr_t)draw) & 15) == 0, "flnerf_composite_backward: raw/draw must be 16-byte aligned");
  if (B == 0) return 0;
  size_t smem = (size_t)kRaysPerBlock * 2 * S * sizeof(float);
  FL_REQUIRE(smem <= 48 * 1024, "flnerf_composite_backward: S=%d too large (max 1536)", S);
  FL_LAUNCH(composite_bwd_kernel, (unsigned)ceil_div64(B, kRaysPerBlock), kRaysPerBlock * 32, smem, stream, B, S,
            (const float4 *)raw, z, rays_d, rays_d_stride, noise, white_bkgd, g_rgb, g_disp, g_acc, g_depth,
            (float4 *)draw);
  return 0;
}

int flnerf_sample_pdf_merge(flnerf_ctx *ctx, int64_t B, int Nc, int Nf, const float *z, const float *weights,
                            const float *u, int det, uint64_t seed, uint64_t offset, float *z_merged,
                            float *z_samples, float *z_std, void *stream) {
  FL_REQUIRE(ctx && z && weights && z_merged && Nc >= 3 && Nf >= 2 && B >= 0, "flnerf_sample_pdf_merge: bad arguments");
  if (B == 0) return 0;
  int P2 = 1;
  while (P2 < Nc + Nf) P2 <<= 1;
  size_t smem = (size_t)kRaysPerBlock * (Nc - 1 + P2) * sizeof(float);
  FL_REQUIRE(smem <= 48 * 1024, "flnerf_sample_pdf_merge: Nc+Nf=%d too large", Nc + Nf);
  FL_LAUNCH(k_merge, (unsigned)ceil_div64(B, kRaysPerBlock), kRaysPerBlock * 32, smem, stream, B,
            Nc, Nf, P2, z, weights, u, det, seed, offset, z_merged, z_samples, z_std, ctx->step_rec);
  return 0;
}

int flnerf_sample_pdf(flnerf_ctx *ctx, int64_t B, int n_bins, int Nf, const float *bins, const float *weights,
                      const float *u, int det, uint64_t seed, uint64_t offset, float *z_samples, void *stream) {
  FL_REQUIRE(ctx && bins && weights && z_samples && n_bins >= 2 && Nf >= 2 && B >= 0, "flnerf_sample_pdf: bad arguments");
  if (B == 0) return 0;
  int Nc = n_bins + 1;
  size_t smem = (size_t)kRaysPerBlock * (Nc - 1 + Nc) * sizeof(float);
  FL_REQUIRE(smem <= 48 * 1024, "flnerf_sample_pdf: too many bins (%d)", n_bins);
  FL_LAUNCH(k_bins, (unsigned)ceil_div64(B, kRaysPerBlock), kRaysPerBlock * 32, smem, stream, B,
            Nc, Nf, Nc, bins, weights, u, det, seed, offset, nullptr, z_samples, nullptr, (const flnerf_step_record *)nullptr);
  return 0;
}

/* nerf++ level-1 sample placement (nerf++-ours/ddp_train_nerf.py:369-382): resample Nf depths from the previous level's
 * weights[..., 1:-1] over the mid-point bins with the nerf++ sample_pdf, then sort-merge with the previous depths. */
int flnerf_pp_sample_pdf_merge(flnerf_ctx *ctx, int64_t B, int Nc, int Nf, const float *z, const float *weights,
                               const float *u, int det, uint64_t seed, uint64_t offset, float *z_merged,
                               float *z_samples, void *stream) {
  FL_REQUIRE(ctx && z && weights && z_merged && Nc >= 3 && Nf >= 2 && B >= 0, "flnerf_pp_sample_pdf_merge: bad arguments");
  if (B == 0) return 0;
  int P2 = 1;
  while (P2 < Nc + Nf) P2 <<= 1;
  size_t smem = (size_t)kRaysPerBlock * (Nc - 1 + P2) * sizeof(float);
  FL_REQUIRE(smem <= 48 * 1024, "flnerf_pp_sample_pdf_merge: Nc+Nf=%d too large", Nc + Nf);
  FL_LAUNCH(k_merge_pp, (unsigned)ceil_div64(B, kRaysPerBlock), kRaysPerBlock * 32, smem, stream, B,
            Nc, Nf, P2, z, weights, u, det, seed, offset, z_merged, z_samples, nullptr, ctx->step_rec);
  return 0;
}

}  // extern "C"
