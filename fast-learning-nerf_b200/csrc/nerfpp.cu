// nerfpp.cu -- first kernels of the NEXT row (nerf++-ours dual-MLP path): level-0 sample placement,
// inverted-sphere background points + their positional encoding, and the foreground/background compositing with its
// backward.  Reference: nerf++-ours/ddp_train_nerf.py:54-81,352-366 and ddp_model.py:16-45,74-143; oracle:
// oracle/nerfpp_oracle.py (pinned against the unmodified reference).  The foreground network's encode and MLP are the
// nerf-ours kernels (same GEMM chain, see nerfpp_oracle.mlp_params_to_nerf_layout); the 84-channel background MLP on
// the tensor cores is not built yet -- these kernels are parity-tested building blocks, not a complete path.
//
// HBM-bound, one warp per ray for the scans (as composite.cu), one thread per sample for the pointwise kernels.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int kRaysPerBlock = 4;
constexpr float kTiny = 1e-6f;   // utils.py:8 TINY_NUMBER
constexpr float kHuge = 1e10f;   // utils.py:7 HUGE_NUMBER

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// intersect_sphere (ddp_train_nerf.py:54-69): depth at which the ray leaves the unit sphere; NaN if the camera is outside
__device__ __forceinline__ float sphere_exit(const float *o, const float *d, float *d1_out, float *pmid_norm_out) {
  float d1 = -dot3(d, o) / dot3(d, d);
  float p[3] = {o[0] + d1 * d[0], o[1] + d1 * d[1], o[2] + d1 * d[2]};
  float pn2 = dot3(p, p);
  float cosv = 1.0f / sqrtf(dot3(d, d));
  if (d1_out) *d1_out = d1;
  if (pmid_norm_out) *pmid_norm_out = sqrtf(pn2);
  return d1 + sqrtf(1.0f - pn2) * cosv;
}

// level-0 depths (ddp_train_nerf.py:352-366): fg linear in [1e-4, sphere exit], bg inverse depth linspace(0,1), both
// jittered inside their mid-point intervals (perturb_samples, :72-81).  One thread per (ray, sample).
__global__ void pp_depths0_kernel(int64_t B, int N, const float *__restrict__ ro, const float *__restrict__ rd,
                                  const float *__restrict__ t_fg, const float *__restrict__ t_bg, int perturb, uint64_t seed,
                                  uint64_t offset, const flnerf_step_record *__restrict__ rec, float *__restrict__ fg_far,
                                  float *__restrict__ fg_z, float *__restrict__ bg_z) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  if (rec) offset += rec->rng_offset;
  int64_t ray = idx / N;
  int i = (int)(idx % N);
  float far = sphere_exit(ro + ray * 3, rd + ray * 3, nullptr, nullptr);
  if (i == 0) fg_far[ray] = far;
  const float near = 1e-4f;
  float step = __fdiv_rn(__fsub_rn(far, near), (float)(N - 1));
  auto fg_at = [&](int k) { return __fadd_rn(near, __fmul_rn((float)k, step)); };
  float lin_step = 1.0f / (float)(N - 1);
  auto bg_at = [&](int k) {  // torch.linspace(0, 1, N): symmetric evaluation
    return (k < N / 2) ? (float)k * lin_step : 1.0f - (float)(N - 1 - k) * lin_step;
  };
  float zf = fg_at(i), zb = bg_at(i);
  if (perturb) {
    float uf, ub;
    if (t_fg) {
      uf = t_fg[idx]; ub = t_bg[idx];
    } else {
      uint32_t r[4];
      philox4x32(seed, offset + (uint64_t)idx, 0x9E3Dull, r);
      uf = u32_to_unit(r[0]); ub = u32_to_unit(r[1]);
    }
    float lo = i > 0 ? __fmul_rn(0.5f, __fadd_rn(zf, fg_at(i - 1))) : zf;
    float hi = i < N - 1 ? __fmul_rn(0.5f, __fadd_rn(fg_at(i + 1), zf)) : zf;
    zf = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), uf));
    lo = i > 0 ? __fmul_rn(0.5f, __fadd_rn(zb, bg_at(i - 1))) : zb;
    hi = i < N - 1 ? __fmul_rn(0.5f, __fadd_rn(bg_at(i + 1), zb)) : zb;
    zb = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), ub));
  }
  fg_z[idx] = zf;
  bg_z[idx] = zb;
}

// depth2pts_outside (ddp_model.py:16-45) + Embedder(4 -> 84) + Embedder(viewdir 3 -> 27), written in the FLIPPED sample
// order NerfNet.forward feeds its background network (ddp_model.py:116-117): output sample j = input sample N-1-j.
// x111 [B, N, 111] fp32 (the reference's layout), bg_z_flip [B, N], pts4 [B, N, 4] (optional, un-flipped, for tests).
__global__ void pp_bg_encode_kernel(int64_t B, int N, const float *__restrict__ ro, const float *__restrict__ rd,
                                    const float *__restrict__ bg_z, float *__restrict__ x111, float *__restrict__ bg_z_flip,
                                    float *__restrict__ pts4) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  int64_t ray = idx / N;
  int i = (int)(idx % N);
  const float *o = ro + ray * 3, *d = rd + ray * 3;
  float depth = bg_z[idx];
  float d1, pmn;
  float exit_t = sphere_exit(o, d, &d1, &pmn);
  float ps[3] = {o[0] + exit_t * d[0], o[1] + exit_t * d[1], o[2] + exit_t * d[2]};
  float ax[3] = {o[1] * ps[2] - o[2] * ps[1], o[2] * ps[0] - o[0] * ps[2], o[0] * ps[1] - o[1] * ps[0]};
  float an = sqrtf(dot3(ax, ax));
  ax[0] /= an; ax[1] /= an; ax[2] /= an;
  float phi = asinf(pmn), theta = asinf(pmn * depth);
  float ang = phi - theta, ca = cosf(ang), sa = sinf(ang);
  float cr[3] = {ax[1] * ps[2] - ax[2] * ps[1], ax[2] * ps[0] - ax[0] * ps[2], ax[0] * ps[1] - ax[1] * ps[0]};
  float ad = dot3(ax, ps) * (1.0f - ca);
  float pn[3] = {ps[0] * ca + cr[0] * sa + ax[0] * ad, ps[1] * ca + cr[1] * sa + ax[1] * ad, ps[2] * ca + cr[2] * sa + ax[2] * ad};
  float nn = sqrtf(dot3(pn, pn));
  float p4[4] = {pn[0] / nn, pn[1] / nn, pn[2] / nn, depth};
  if (pts4) {
#pragma unroll
    for (int c = 0; c < 4; ++c) pts4[idx * 4 + c] = p4[c];
  }
  int64_t out = ray * N + (N - 1 - i);
  bg_z_flip[out] = depth;
  float *x = x111 + out * 111;
#pragma unroll
  for (int c = 0; c < 4; ++c) x[c] = p4[c];
  for (int k = 0; k < 10; ++k) {
    float f = (float)(1u << k);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      x[4 + 8 * k + c] = sinf(p4[c] * f);
      x[8 + 8 * k + c] = cosf(p4[c] * f);
    }
  }
  float dn = sqrtf(dot3(d, d));
  float v[3] = {d[0] / dn, d[1] / dn, d[2] / dn};
  float *xv = x + 84;
#pragma unroll
  for (int c = 0; c < 3; ++c) xv[c] = v[c];
  for (int k = 0; k < 4; ++k) {
    float f = (float)(1u << k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      xv[3 + 6 * k + c] = sinf(v[c] * f);
      xv[6 + 6 * k + c] = cosf(v[c] * f);
    }
  }
}

// One pass of the nerf++ alpha compositing over S samples of a ray: sigma = |raw.w|, alpha = 1 - exp(-sigma * dist),
// T = exclusive cumprod(1 - alpha + 1e-6).  kFg: dist = (z[i+1] - z[i]) * |d|, last = (z_max - z[S-1]) * |d|
// (ddp_model.py:96-98); else (background, z descending): dist = z[i] - z[i+1], last = 1e10 (:118-121).
// Returns the weighted sums and the final transmittance; optionally stores the weights.
template <bool kFg>
__device__ __forceinline__ void pp_pass(int S, const float4 *__restrict__ rr, const float *__restrict__ zz, float nrm,
                                        float z_max, int lane, float *__restrict__ w_out, float &sr, float &sg, float &sb,
                                        float &sd, float &t_final) {
  float carry = 1.0f;
  sr = sg = sb = sd = 0.f;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    bool valid = i < S;
    float4 r = valid ? rr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float z0 = valid ? zz[i] : 0.f;
    float z1 = (i + 1 < S) ? zz[i + 1] : 0.f;
    float dist;
    if (kFg) dist = __fmul_rn(nrm, i == S - 1 ? __fsub_rn(z_max, z0) : __fsub_rn(z1, z0));
    else dist = i == S - 1 ? kHuge : __fsub_rn(z0, z1);
    float alpha = 1.0f - expf(-__fmul_rn(fabsf(r.w), dist));
    float om = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), kTiny) : 1.0f;
    float incl = warp_scan_mul(om, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    float T = carry * excl;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    float w = valid ? alpha * T : 0.f;
    if (w_out && valid) w_out[i] = w;
    sr += w * sigmoidf_(r.x);
    sg += w * sigmoidf_(r.y);
    sb += w * sigmoidf_(r.z);
    sd += w * z0;
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd);
  t_final = carry;
}

// NerfNet.forward compositing (ddp_model.py:93-133) from the two networks' RAW outputs (rgb before the sigmoid, sigma
// before |.|): rgb = fg + bg_lambda * bg with bg_lambda = the foreground's final transmittance.  Warp per ray.
__global__ void __launch_bounds__(kRaysPerBlock * 32)
pp_composite_fwd_kernel(int64_t B, int Sf, int Sb, const float4 *__restrict__ raw_fg, const float *__restrict__ fg_z,
                        const float *__restrict__ fg_far, const float4 *__restrict__ raw_bg,
                        const float *__restrict__ bg_z, const float *__restrict__ rd, float *__restrict__ rgb,
                        float *__restrict__ fg_w, float *__restrict__ bg_w, float *__restrict__ aux /* [B][9] */) {
  int lane = threadIdx.x & 31;
  int64_t ray = (int64_t)blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
  if (ray >= B) return;
  float nrm = sqrtf(dot3(rd + ray * 3, rd + ray * 3));
  float fr, fg, fb, fd, lam, br, bgc, bb, bd, tb;
  pp_pass<true>(Sf, raw_fg + ray * Sf, fg_z + ray * Sf, nrm, fg_far[ray], lane, fg_w ? fg_w + ray * Sf : nullptr, fr, fg, fb,
                fd, lam);
  pp_pass<false>(Sb, raw_bg + ray * Sb, bg_z + ray * Sb, nrm, 0.f, lane, bg_w ? bg_w + ray * Sb : nullptr, br, bgc, bb, bd, tb);
  if (lane == 0) {
    rgb[ray * 3] = fr + lam * br; rgb[ray * 3 + 1] = fg + lam * bgc; rgb[ray * 3 + 2] = fb + lam * bb;
    if (aux) {  // fg_rgb(3), fg_depth, bg_rgb(3, scaled), bg_depth (scaled), bg_lambda
      float *a = aux + ray * 9;
      a[0] = fr; a[1] = fg; a[2] = fb; a[3] = fd; a[4] = lam * br; a[5] = lam * bgc; a[6] = lam * bb; a[7] = lam * bd; a[8] = lam;
    }
  }
}

// Backward of one pass: with q_i = dL/dw_i (per sample, without T), dL/dalpha_i = q_i T_i - (sum_{k>i} q_k w_k + tail) /
// (1 - alpha_i + 1e-6); tail = the gradient flowing into the pass's final transmittance times that transmittance
// (the background term for the foreground pass, 0 for the background pass).  smem: alpha[S], T[S] per warp.
template <bool kFg>
__device__ __forceinline__ void pp_pass_bwd(int S, const float4 *__restrict__ rr, const float *__restrict__ zz, float nrm,
                                            float z_max, int lane, float gx, float gy, float gz, float scale, float tail,
                                            float *s_alpha, float *s_T, float4 *__restrict__ draw) {
  float carry = 1.0f;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    bool valid = i < S;
    float r3 = valid ? rr[i].w : 0.f;
    float z0 = valid ? zz[i] : 0.f;
    float z1 = (i + 1 < S) ? zz[i + 1] : 0.f;
    float dist;
    if (kFg) dist = __fmul_rn(nrm, i == S - 1 ? __fsub_rn(z_max, z0) : __fsub_rn(z1, z0));
    else dist = i == S - 1 ? kHuge : __fsub_rn(z0, z1);
    float alpha = 1.0f - expf(-__fmul_rn(fabsf(r3), dist));
    float om = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), kTiny) : 1.0f;
    float incl = warp_scan_mul(om, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    if (valid) { s_alpha[i] = alpha; s_T[i] = carry * excl; }
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  float suffix = tail;  // sum over samples after the current chunk of q_k w_k (+ tail)
  for (int base = ((S - 1) / 32) * 32; base >= 0; base -= 32) {
    int i = base + lane;
    bool valid = i < S;
    float4 r = valid ? rr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float z0 = valid ? zz[i] : 0.f;
    float z1 = (i + 1 < S) ? zz[i + 1] : 0.f;
    float alpha = valid ? s_alpha[i] : 0.f, T = valid ? s_T[i] : 0.f;
    float w = alpha * T;
    float cx = sigmoidf_(r.x), cy = sigmoidf_(r.y), cz = sigmoidf_(r.z);
    float q = scale * (gx * cx + gy * cy + gz * cz);
    float qw = valid ? q * w : 0.f;
    float rs = warp_rscan_add(qw, lane);                // inclusive suffix inside the chunk
    float after = rs - qw + suffix;                      // strictly-after sum (+ later chunks + tail)
    suffix += __shfl_sync(0xffffffffu, rs, 0);
    if (valid) {
      float dist;
      if (kFg) dist = __fmul_rn(nrm, i == S - 1 ? __fsub_rn(z_max, z0) : __fsub_rn(z1, z0));
      else dist = i == S - 1 ? kHuge : __fsub_rn(z0, z1);
      float dalpha = q * T - after / __fadd_rn(__fsub_rn(1.0f, alpha), kTiny);
      float sig = fabsf(r.w);
      float dsig = dalpha * dist * expf(-__fmul_rn(sig, dist));      // d alpha / d sigma = dist * exp(-sigma dist)
      float sgn = r.w > 0.f ? 1.f : (r.w < 0.f ? -1.f : 0.f);        // torch.abs backward: sign(x), 0 at 0
      draw[i] = make_float4(scale * gx * w * cx * (1.f - cx), scale * gy * w * cy * (1.f - cy), scale * gz * w * cz * (1.f - cz),
                            dsig * sgn);
    }
  }
}

__global__ void __launch_bounds__(kRaysPerBlock * 32)
pp_composite_bwd_kernel(int64_t B, int Sf, int Sb, const float4 *__restrict__ raw_fg, const float *__restrict__ fg_z,
                        const float *__restrict__ fg_far, const float4 *__restrict__ raw_bg,
                        const float *__restrict__ bg_z, const float *__restrict__ rd, const float *__restrict__ g_rgb,
                        float4 *__restrict__ draw_fg, float4 *__restrict__ draw_bg) {
  extern __shared__ float sm[];
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t ray = (int64_t)blockIdx.x * kRaysPerBlock + wib;
  if (ray >= B) return;
  const int Smax = Sf > Sb ? Sf : Sb;
  float *s_alpha = sm + (size_t)wib * 2 * Smax, *s_T = s_alpha + Smax;
  float nrm = sqrtf(dot3(rd + ray * 3, rd + ray * 3));
  float gx = g_rgb[ray * 3], gy = g_rgb[ray * 3 + 1], gz = g_rgb[ray * 3 + 2];
  // forward quantities the backward needs: bg_lambda and the un-scaled background colour
  float fr, fg, fb, fd, lam, br, bgc, bb, bd, tb;
  pp_pass<true>(Sf, raw_fg + ray * Sf, fg_z + ray * Sf, nrm, fg_far[ray], lane, nullptr, fr, fg, fb, fd, lam);
  pp_pass<false>(Sb, raw_bg + ray * Sb, bg_z + ray * Sb, nrm, 0.f, lane, nullptr, br, bgc, bb, bd, tb);
  const float g_cb = gx * br + gy * bgc + gz * bb;   // dL/d(bg_lambda)
  pp_pass_bwd<true>(Sf, raw_fg + ray * Sf, fg_z + ray * Sf, nrm, fg_far[ray], lane, gx, gy, gz, 1.0f, g_cb * lam, s_alpha, s_T,
                    draw_fg + ray * Sf);
  __syncwarp();
  pp_pass_bwd<false>(Sb, raw_bg + ray * Sb, bg_z + ray * Sb, nrm, 0.f, lane, gx, gy, gz, lam, 0.f, s_alpha, s_T,
                     draw_bg + ray * Sb);
}

}  // namespace

extern "C" {

int flnerf_pp_depths0(flnerf_ctx *ctx, int64_t B, int N, const float *rays_o, const float *rays_d, const float *t_fg,
                      const float *t_bg, int perturb, uint64_t seed, uint64_t offset, float *fg_far, float *fg_z, float *bg_z,
                      void *stream) {
  FL_REQUIRE(ctx && rays_o && rays_d && fg_far && fg_z && bg_z && N >= 2 && B >= 0 && (!t_fg == !t_bg),
             "flnerf_pp_depths0: bad arguments");
  if (B == 0) return 0;
  FL_LAUNCH(pp_depths0_kernel, (unsigned)ceil_div64(B * N, 256), 256, 0, stream, B, N, rays_o, rays_d, t_fg, t_bg, perturb, seed,
            offset, ctx->step_rec, fg_far, fg_z, bg_z);
  return 0;
}

int flnerf_pp_bg_encode(flnerf_ctx *ctx, int64_t B, int N, const float *rays_o, const float *rays_d, const float *bg_z,
                        float *x111, float *bg_z_flip, float *pts4, void *stream) {
  FL_REQUIRE(ctx && rays_o && rays_d && bg_z && x111 && bg_z_flip && N >= 1 && B >= 0, "flnerf_pp_bg_encode: bad arguments");
  if (B == 0) return 0;
  FL_LAUNCH(pp_bg_encode_kernel, (unsigned)ceil_div64(B * N, 128), 128, 0, stream, B, N, rays_o, rays_d, bg_z, x111, bg_z_flip,
            pts4);
  return 0;
}

int flnerf_pp_composite_forward(flnerf_ctx *ctx, int64_t B, int Sf, int Sb, const float *raw_fg, const float *fg_z,
                                const float *fg_far, const float *raw_bg, const float *bg_z_flip, const float *rays_d,
                                float *rgb, float *fg_weights, float *bg_weights, float *aux9, void *stream) {
  FL_REQUIRE(ctx && raw_fg && fg_z && fg_far && raw_bg && bg_z_flip && rays_d && rgb && Sf >= 2 && Sb >= 2 && B >= 0,
             "flnerf_pp_composite_forward: bad arguments");
  FL_REQUIRE((((uintptr_t)raw_fg | (uintptr_t)raw_bg) & 15) == 0, "flnerf_pp_composite_forward: raw must be 16-byte aligned");
  if (B == 0) return 0;
  FL_LAUNCH(pp_composite_fwd_kernel, (unsigned)ceil_div64(B, kRaysPerBlock), kRaysPerBlock * 32, 0, stream, B, Sf, Sb,
            (const float4 *)raw_fg, fg_z, fg_far, (const float4 *)raw_bg, bg_z_flip, rays_d, rgb, fg_weights, bg_weights, aux9);
  return 0;
}

int flnerf_pp_composite_backward(flnerf_ctx *ctx, int64_t B, int Sf, int Sb, const float *raw_fg, const float *fg_z,
                                 const float *fg_far, const float *raw_bg, const float *bg_z_flip, const float *rays_d,
                                 const float *g_rgb, float *draw_fg, float *draw_bg, void *stream) {
  FL_REQUIRE(ctx && raw_fg && fg_z && fg_far && raw_bg && bg_z_flip && rays_d && g_rgb && draw_fg && draw_bg && Sf >= 2 &&
                 Sb >= 2 && B >= 0,
             "flnerf_pp_composite_backward: bad arguments");
  FL_REQUIRE((((uintptr_t)raw_fg | (uintptr_t)raw_bg | (uintptr_t)draw_fg | (uintptr_t)draw_bg) & 15) == 0,
             "flnerf_pp_composite_backward: raw/draw must be 16-byte aligned");
  if (B == 0) return 0;
  size_t smem = (size_t)kRaysPerBlock * 2 * (Sf > Sb ? Sf : Sb) * sizeof(float);
  FL_REQUIRE(smem <= 48 * 1024, "flnerf_pp_composite_backward: too many samples per ray");
  FL_LAUNCH(pp_composite_bwd_kernel, (unsigned)ceil_div64(B, kRaysPerBlock), kRaysPerBlock * 32, smem, stream, B, Sf, Sb,
            (const float4 *)raw_fg, fg_z, fg_far, (const float4 *)raw_bg, bg_z_flip, rays_d, g_rgb, (float4 *)draw_fg,
            (float4 *)draw_bg);
  return 0;
}

}  // extern "C"
