// rays.cu -- ray generation, NDC warp + packing, stratified depths, positional encoding.
// Reference: nerf-ours/run_nerf_helpers.py:15-108, render.py:59-80, 244-268, run_nerf.py:50-64.
//
// All of these are HBM-bound elementwise kernels: one thread per output element (or per 16-byte
// output chunk), unit-stride stores, inputs re-read through L1/L2.  Arithmetic uses the *_rn
// intrinsics wherever the reference performs separate tensor ops, so that no FMA contraction
// changes the rounding relative to torch (x*freq is exact; sinf/cosf are the accurate versions).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

struct Cam {
  float cx, cy, fx, fy;
};
struct Pose {
  float m[12];  // 3x4 row-major c2w
};

__device__ __forceinline__ void pixel_ray(const Cam &c, const float *P, int row, int col, float o[3], float d[3]) {
  // dirs = [(i-cx)/fx, -(j-cy)/fy, -1]; rays_d[a] = sum_b dirs[b]*R[a][b]   (helpers:72-75)
  float x = __fdiv_rn(__fsub_rn((float)col, c.cx), c.fx);
  float y = -__fdiv_rn(__fsub_rn((float)row, c.cy), c.fy);
  float z = -1.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float s = __fadd_rn(__fmul_rn(x, P[a * 4 + 0]), __fmul_rn(y, P[a * 4 + 1]));
    d[a] = __fadd_rn(s, __fmul_rn(z, P[a * 4 + 2]));
    o[a] = P[a * 4 + 3];
  }
}

__global__ void raygen_kernel(int H, int W, Cam cam, Pose pose, float *__restrict__ ro, float *__restrict__ rd) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (int64_t)H * W) return;
  int row = (int)(p / W), col = (int)(p % W);
  float o[3], d[3];
  pixel_ray(cam, pose.m, row, col, o, d);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    ro[p * 3 + a] = o[a];
    rd[p * 3 + a] = d[a];
  }
}

// viewdir normalisation (from the un-warped direction, render.py:59-66), optional NDC warp (run_nerf_helpers.py:91-108 with
// near = 1) and the [o, d, near, far, viewdir] row of render.py:74-80
__device__ __forceinline__ void pack_ray(float o0, float o1, float o2, float d0, float d1, float d2, float near_, float far_, int ndc,
                                         float sx, float sy, float q[11]) {
  float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
  float v0 = __fdiv_rn(d0, nrm), v1 = __fdiv_rn(d1, nrm), v2 = __fdiv_rn(d2, nrm);
  if (ndc) {
    const float nr = 1.0f;
    float t = -__fdiv_rn(__fadd_rn(nr, o2), d2);
    o0 = __fadd_rn(o0, __fmul_rn(t, d0));
    o1 = __fadd_rn(o1, __fmul_rn(t, d1));
    o2 = __fadd_rn(o2, __fmul_rn(t, d2));
    float a0 = __fdiv_rn(__fmul_rn(sx, o0), o2);
    float a1 = __fdiv_rn(__fmul_rn(sy, o1), o2);
    float a2 = __fadd_rn(1.0f, __fdiv_rn(2.0f * nr, o2));
    float b0 = __fmul_rn(sx, __fsub_rn(__fdiv_rn(d0, d2), __fdiv_rn(o0, o2)));
    float b1 = __fmul_rn(sy, __fsub_rn(__fdiv_rn(d1, d2), __fdiv_rn(o1, o2)));
    float b2 = __fdiv_rn(-2.0f * nr, o2);
    o0 = a0; o1 = a1; o2 = a2; d0 = b0; d1 = b1; d2 = b2;
  }
  q[0] = o0; q[1] = o1; q[2] = o2; q[3] = d0; q[4] = d1; q[5] = d2;
  q[6] = near_; q[7] = far_; q[8] = v0; q[9] = v1; q[10] = v2;
}

__global__ void pack_rays_kernel(int64_t B, const float *__restrict__ ro, const float *__restrict__ rd, float near_,
                                 float far_, int ndc, float sx, float sy, float *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float q[11];
  pack_ray(ro[i * 3], ro[i * 3 + 1], ro[i * 3 + 2], rd[i * 3], rd[i * 3 + 1], rd[i * 3 + 2], near_, far_, ndc, sx, sy, q);
#pragma unroll
  for (int a = 0; a < 11; ++a) out[i * 11 + a] = q[a];
}

__device__ __forceinline__ float base_depth(float near_, float far_, float t, int lindisp) {
  float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near_, omt), __fmul_rn(far_, t));                       // render.py:246
  float inv = __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, near_), omt), __fmul_rn(__fdiv_rn(1.0f, far_), t));  // :248
  return __fdiv_rn(1.0f, inv);
}

__global__ void coarse_depths_kernel(int64_t B, int Nc, const float *__restrict__ rays11, const float *__restrict__ tv,
                                     const float *__restrict__ t_rand, int perturb, int lindisp, uint64_t seed,
                                     uint64_t offset, const flnerf_step_record *__restrict__ rec, float *__restrict__ z) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Nc) return;
  if (rec) offset += rec->rng_offset;
  int64_t ray = idx / Nc;
  int j = (int)(idx % Nc);
  float near_ = rays11[ray * 11 + 6], far_ = rays11[ray * 11 + 7];
  float zj = base_depth(near_, far_, tv[j], lindisp);
  if (perturb) {  // render.py:252-266
    float lo = zj, hi = zj;
    if (j > 0) lo = __fmul_rn(0.5f, __fadd_rn(zj, base_depth(near_, far_, tv[j - 1], lindisp)));
    if (j < Nc - 1) hi = __fmul_rn(0.5f, __fadd_rn(base_depth(near_, far_, tv[j + 1], lindisp), zj));
    float u;
    if (t_rand) {
      u = t_rand[idx];
    } else {
      uint32_t r[4];
      philox4x32(seed, offset + (uint64_t)idx, 0x5A17ull, r);
      u = u32_to_unit(r[0]);
    }
    zj = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u));
  }
  z[idx] = zj;
}

// channel ch of PE_L(x): ch<3 -> x[ch]; else k=(ch-3)/6, sin for (ch-3)%6<3 else cos, coord=(ch-3)%3
__device__ __forceinline__ float pe_channel(const float x[3], int ch) {
  if (ch < 3) return x[ch];
  int q = ch - 3;
  int k = q / 6, r = q % 6;
  float a = x[r % 3] * (float)(1u << k);  // exact: power-of-two scale (helpers:32,38)
  return r < 3 ? sinf(a) : cosf(a);
}

__global__ void posenc_kernel(int64_t n, int L, const float *__restrict__ x, float *__restrict__ out) {
  int C = 3 + 6 * L;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * C) return;
  int64_t p = idx / C;
  int ch = (int)(idx % C);
  float v[3] = {x[p * 3], x[p * 3 + 1], x[p * 3 + 2]};
  out[idx] = pe_channel(v, ch);
}

__device__ __forceinline__ void sample_point(const float *__restrict__ r11, float z, float p[3]) {
  // pts = o + d*z with separate rounding (render.py:268)
#pragma unroll
  for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(r11[a], __fmul_rn(r11[3 + a], z));
}

__global__ void encode_f32_kernel(int64_t B, int S, const float *__restrict__ rays11, const float *__restrict__ z,
                                  float *__restrict__ x90) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = B * S;
  if (idx >= n * 90) return;
  int64_t row = idx / 90;
  int ch = (int)(idx % 90);
  const float *r11 = rays11 + (row / S) * 11;
  float v[3];
  if (ch < 63) {
    sample_point(r11, z[row], v);
    x90[idx] = pe_channel(v, ch);
  } else {
    v[0] = r11[8]; v[1] = r11[9]; v[2] = r11[10];
    x90[idx] = pe_channel(v, ch - 63);
  }
}

// one thread = one row (63 channels + zero pad) of a [128 x 64] bf16 SWIZZLE_128B tile.  The tile holds bf16 (8
// mantissa bits), so only the octaves k = 0 and k = 5 take an exact sincosf; the four octaves after each come from the
// double-angle recurrence (error <= 2^4 ulp(fp32) ~ 2e-6, 2000x below the bf16 rounding step).  The fp32 parity path
// (encode_f32_kernel / posenc_kernel) evaluates every channel exactly, and so does kX3 (FLNERF_MODE_BF16X3), which
// also writes the residual tile lo = bf16(v - hi) at tiles + lo_off.
// kFrame (the eval path, render.py:94-146 -> render(c2w=...)): the rays are not read but GENERATED -- row -> ray = row / S ->
// pixel pixel0 + ray of the frame -> get_rays + (ndc) + packing, depths = the un-jittered coarse depths of render.py:244-249 --
// so one launch goes from the camera pose to the MLP's input tiles; the first sample's thread also writes rays11[ray],
// dirpe[ray] (for the fine pass and the compositing kernels) and every thread its z.
struct FrameArgs {
  Cam cam;
  Pose pose;
  int W;
  int ndc, lindisp;
  float near_, far_, sx, sy;
  int64_t pixel0;
  const float *tv;      // linspace(0, 1, S)
  float *rays11, *z, *dirpe;
};

template <bool kX3, bool kFrame>
__global__ void __launch_bounds__(128) encode_tc_kernel(int64_t n, int64_t n_pad, int S, const float *__restrict__ rays11,
                                                        const float *__restrict__ z, uint8_t *__restrict__ tiles, size_t lo_off,
                                                        FrameArgs fa) {
  // block = one 128-row tile (n_pad is a multiple of 128: no partial blocks).  The tile is assembled in shared memory in the
  // very image the MLP kernels consume, then leaves as ONE 16 KB bulk store per tile set (cp.async.bulk: full 128-byte lines
  // instead of 32 scattered 16-byte pieces per store instruction).
  __shared__ __align__(1024) uint8_t s_tile[kX3 ? 2 : 1][16384];
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t pk[32], pl[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { pk[i] = 0u; pl[i] = 0u; }
  if (row < n) {
    float v[64];
    float p[3], sn[3], cs[3];
    if (kFrame) {
      const int64_t ray = row / S;
      const int j = (int)(row - ray * S);
      const int64_t pix = fa.pixel0 + ray;
      float o[3], d[3], q[11];
      pixel_ray(fa.cam, fa.pose.m, (int)(pix / fa.W), (int)(pix % fa.W), o, d);
      pack_ray(o[0], o[1], o[2], d[0], d[1], d[2], fa.near_, fa.far_, fa.ndc, fa.sx, fa.sy, q);
      const float zj = base_depth(fa.near_, fa.far_, fa.tv[j], fa.lindisp);
      fa.z[row] = zj;
      if (j == 0) {
#pragma unroll
        for (int a = 0; a < 11; ++a) fa.rays11[ray * 11 + a] = q[a];
        const float vd[3] = {q[8], q[9], q[10]};
        for (int ch = 0; ch < 32; ++ch) fa.dirpe[ray * 32 + ch] = ch < 27 ? pe_channel(vd, ch) : 0.0f;
      }
      sample_point(q, zj, p);
    } else {
      const int64_t ray = row / S;
      sample_point(rays11 + ray * 11, z[row], p);
      if (fa.dirpe && row == ray * S) {       // the ray's first sample also writes its view-direction PE (saves a launch)
        const float vd[3] = {rays11[ray * 11 + 8], rays11[ray * 11 + 9], rays11[ray * 11 + 10]};
        for (int ch = 0; ch < 32; ++ch) fa.dirpe[ray * 32 + ch] = ch < 27 ? pe_channel(vd, ch) : 0.0f;
      }
    }
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
    v[63] = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (kX3 || k == 0 || k == 5) {
          sincosf(p[a] * (float)(1u << k), &sn[a], &cs[a]);  // exact power-of-two scale (helpers:32,38)
        } else {
          const float s2 = 2.f * sn[a] * cs[a], c2 = fmaf(-2.f * sn[a], sn[a], 1.f);
          sn[a] = s2;
          cs[a] = c2;
        }
        v[3 + 6 * k + a] = sn[a];
        v[6 + 6 * k + a] = cs[a];
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      pk[i] = *reinterpret_cast<uint32_t *>(&h);
      if (kX3) {
        __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * i] - __low2float(h), v[2 * i + 1] - __high2float(h));
        pl[i] = *reinterpret_cast<uint32_t *>(&l);
      }
    }
  }
  const uint32_t r = (uint32_t)(row & 127);
  const uint32_t roff = (r >> 3) * 1024u + (r & 7u) * 128u;
#pragma unroll
  for (uint32_t q = 0; q < 8; ++q) {
    *reinterpret_cast<uint4 *>(s_tile[0] + roff + ((q ^ (r & 7u)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    if (kX3)
      *reinterpret_cast<uint4 *>(s_tile[kX3 ? 1 : 0] + roff + ((q ^ (r & 7u)) << 4)) =
          make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);
  }
  tc::fence_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint8_t *dst = tiles + (row >> 7) * 16384;
    tc::bulk_s2g(dst, tc::smem_u32(s_tile[0]), 16384u);
    if (kX3) tc::bulk_s2g(dst + lo_off, tc::smem_u32(s_tile[kX3 ? 1 : 0]), 16384u);
    tc::bulk_commit();
    tc::bulk_wait_all0();        // the source is this block's shared memory: it must outlive the copy
  }
}

// generic NeRF.forward(x[n,90]) entry for the tensor-core MLP: already-embedded fp32 rows -> bf16 tile image
__global__ void pack_x90_kernel(int64_t n, int64_t n_pad, const float *__restrict__ x90, uint8_t *__restrict__ tiles,
                                float *__restrict__ dirpe, size_t lo_off) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad * 8) return;
  int64_t row = idx >> 3;
  int q = (int)(idx & 7);
  uint32_t packed[4] = {0u, 0u, 0u, 0u}, packed_lo[4] = {0u, 0u, 0u, 0u};
  if (row < n) {
    const float *src = x90 + row * 90;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int c0 = q * 8 + 2 * e;
      float a = src[c0];
      float b = (c0 + 1 < 63) ? src[c0 + 1] : 0.0f;
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      packed[e] = *reinterpret_cast<uint32_t *>(&h);
      __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
      packed_lo[e] = *reinterpret_cast<uint32_t *>(&l);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int j = q * 4 + e;
      dirpe[row * 32 + j] = j < 27 ? src[63 + j] : 0.0f;
    }
  }
  int64_t tile = row >> 7;
  uint32_t r = (uint32_t)(row & 127);
  *reinterpret_cast<uint4 *>(tiles + tile * 16384 + sw128_offset(r, (uint32_t)q * 8)) =
      make_uint4(packed[0], packed[1], packed[2], packed[3]);
  if (lo_off)  // FLNERF_MODE_BF16X3: the residual tile set
    *reinterpret_cast<uint4 *>(tiles + lo_off + tile * 16384 + sw128_offset(r, (uint32_t)q * 8)) =
        make_uint4(packed_lo[0], packed_lo[1], packed_lo[2], packed_lo[3]);
}

// already-embedded rows x[n, in_pts + in_views] (in_views = 27) -> tensor-core tiles with ceil(in_pts / 64) slabs of [128 x 64]
// per 128-row tile (slab-major inside a tile; zero pad behind in_pts), hi set and -- lo_off != 0 -- lo set; the view part of
// the FIRST row of every S-row ray -> dirpe[ray, 32].  One thread = one 16-byte chunk.
__global__ void pack_xrows_kernel(int64_t n, int64_t n_pad, int in_pts, int n_slabs, int ld, int S, const float *__restrict__ x,
                                  uint8_t *__restrict__ tiles, float *__restrict__ dirpe, size_t lo_off) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = n_slabs * 8;
  if (idx >= n_pad * per_row) return;
  const int64_t row = idx / per_row;
  const int qq = (int)(idx % per_row), sl = qq >> 3, q = qq & 7;
  uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
  if (row < n) {
    const float *src = x + row * ld;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c0 = sl * 64 + q * 8 + 2 * e;
      const float a = c0 < in_pts ? src[c0] : 0.0f, b = c0 + 1 < in_pts ? src[c0 + 1] : 0.0f;
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      hi[e] = *reinterpret_cast<uint32_t *>(&h);
      __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
      lo[e] = *reinterpret_cast<uint32_t *>(&l);
    }
    if (sl == 0 && row % S == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = q * 4 + e;
        dirpe[(row / S) * 32 + j] = j < 27 ? src[in_pts + j] : 0.0f;
      }
    }
  }
  const int64_t tile = row >> 7;
  const uint32_t r = (uint32_t)(row & 127);
  const size_t off = ((size_t)tile * n_slabs + sl) * 16384 + sw128_offset(r, (uint32_t)q * 8);
  *reinterpret_cast<uint4 *>(tiles + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (lo_off) *reinterpret_cast<uint4 *>(tiles + lo_off + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// images: fp32 [n,H,W,3], or -- lut != nullptr -- uint8 [n,H,W,3] decoded through lut[256] (the loaders' float32(u / 255.)
// values, load_blender.py:37: a quarter of the bytes, the same floats)
__global__ void gather_batch_kernel(int64_t B, int64_t first, int64_t stride, const int32_t *__restrict__ ray_pix,
                                    const int32_t *__restrict__ ray_gid, int cap, int H, int W, Cam cam,
                                    const float *__restrict__ poses, const void *__restrict__ images_any,
                                    const float *__restrict__ lut,
                                    float *__restrict__ ro, float *__restrict__ rd, float *__restrict__ target,
                                    int32_t *__restrict__ leaf_gid, const flnerf_step_record *__restrict__ rec) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  int64_t j = first + (rec ? rec->first : 0) + k * stride;
  int pix = ray_pix[j], gid = ray_gid[j];
  int img = gid / cap;
  int row = pix / W, col = pix % W;
  float o[3], d[3];
  pixel_ray(cam, poses + (int64_t)img * 12, row, col, o, d);
  const int64_t pidx = (((int64_t)img * H + row) * W + col) * 3;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    ro[k * 3 + a] = o[a];
    rd[k * 3 + a] = d[a];
    target[k * 3 + a] = lut ? lut[reinterpret_cast<const uint8_t *>(images_any)[pidx + a]]
                            : reinterpret_cast<const float *>(images_any)[pidx + a];
  }
  if (leaf_gid) leaf_gid[k] = gid;
}

// gen_rays_v3's gather (tree.py:270-285): colour, direction and origin images of image `img` sampled with
// F.grid_sample(bilinear, zeros padding, align_corners=False) at the grid ((x / (H/2)) - 1, (y / (W/2)) - 1).  grid_sample reads
// a grid's FIRST component as the WIDTH coordinate, so the reference's (row, col) pairs land TRANSPOSED: ix = (gx + 1) W / 2 - 0.5
// with gx made from the ROW x, iy from the column y (a quirk of the reference, reproduced: rays and colours stay consistent).
// The direction / origin images are get_rays of the image's pose, regenerated per corner.
__global__ void gather_sub_kernel(int64_t B, const float *__restrict__ ray_xy, const int32_t *__restrict__ ray_gid, int cap, int H,
                                  int W, Cam cam, const float *__restrict__ poses, const void *__restrict__ images_any,
                                  const float *__restrict__ lut, float *__restrict__ ro, float *__restrict__ rd,
                                  float *__restrict__ target) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  const int img = ray_gid[k] / cap;
  const float gx = __fsub_rn(__fdiv_rn(ray_xy[k * 2], (float)H / 2.0f), 1.0f);
  const float gy = __fsub_rn(__fdiv_rn(ray_xy[k * 2 + 1], (float)W / 2.0f), 1.0f);
  const float ix = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W / 2.0f), 0.5f);
  const float iy = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H / 2.0f), 0.5f);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = __fsub_rn(ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
  const float wy1 = __fsub_rn(iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
  const float w4[4] = {__fmul_rn(wx0, wy0), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy1), __fmul_rn(wx1, wy1)};   // nw, ne, sw, se
  float c[3] = {0.f, 0.f, 0.f}, o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = x0 + (q & 1), row = y0 + (q >> 1);
    if (col < 0 || col >= W || row < 0 || row >= H) continue;          // zeros padding
    float po[3], pd[3];
    pixel_ray(cam, poses + (int64_t)img * 12, row, col, po, pd);
    const int64_t pidx = (((int64_t)img * H + row) * W + col) * 3;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = lut ? lut[reinterpret_cast<const uint8_t *>(images_any)[pidx + a]]
                          : reinterpret_cast<const float *>(images_any)[pidx + a];
      c[a] = __fadd_rn(c[a], __fmul_rn(v, w4[q]));
      o[a] = __fadd_rn(o[a], __fmul_rn(po[a], w4[q]));
      d[a] = __fadd_rn(d[a], __fmul_rn(pd[a], w4[q]));
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    ro[k * 3 + a] = o[a];
    rd[k * 3 + a] = d[a];
    target[k * 3 + a] = c[a];
  }
}

Cam make_cam(const double *K) {
  Cam c;
  c.fx = (float)K[0]; c.cx = (float)K[2]; c.fy = (float)K[4]; c.cy = (float)K[5];
  return c;
}

}  // namespace

extern "C" {

int flnerf_raygen(flnerf_ctx *ctx, int H, int W, const double *h_K, const float *h_c2w, float *rays_o, float *rays_d,
                  void *stream) {
  FL_REQUIRE(ctx && h_K && h_c2w && rays_o && rays_d && H > 0 && W > 0, "flnerf_raygen: bad arguments");
  Pose P;
  for (int i = 0; i < 12; ++i) P.m[i] = h_c2w[i];
  int64_t n = (int64_t)H * W;
  FL_LAUNCH(raygen_kernel, (unsigned)ceil_div64(n, 256), 256, 0, stream, H, W, make_cam(h_K), P, rays_o, rays_d);
  return 0;
}

int flnerf_pack_rays(flnerf_ctx *ctx, int64_t B, const float *rays_o, const float *rays_d, float near_, float far_,
                     int ndc, int H, int W, double focal, float *rays11, void *stream) {
  if (B == 0) return 0;
  FL_REQUIRE(ctx && rays_o && rays_d && rays11 && B > 0, "flnerf_pack_rays: bad arguments");
  float sx = (float)(-1.0 / (W / (2.0 * focal))), sy = (float)(-1.0 / (H / (2.0 * focal)));
  FL_LAUNCH(pack_rays_kernel, (unsigned)ceil_div64(B, 256), 256, 0, stream, B, rays_o, rays_d, near_, far_, ndc, sx, sy,
            rays11);
  return 0;
}

int flnerf_coarse_depths(flnerf_ctx *ctx, int64_t B, int Nc, const float *rays11, const float *t_vals,
                         const float *t_rand, int perturb, int lindisp, uint64_t seed, uint64_t offset, float *z,
                         void *stream) {
  FL_REQUIRE(ctx && rays11 && t_vals && z && Nc > 0 && B >= 0, "flnerf_coarse_depths: bad arguments");
  if (B == 0) return 0;
  FL_LAUNCH(coarse_depths_kernel, (unsigned)ceil_div64(B * Nc, 256), 256, 0, stream, B, Nc, rays11, t_vals, t_rand,
            perturb, lindisp, seed, offset, ctx->step_rec, z);
  return 0;
}

int flnerf_posenc(flnerf_ctx *ctx, int64_t n, int L, const float *x, float *out, void *stream) {
  FL_REQUIRE(ctx && x && out && L >= 0 && L <= 16 && n >= 0, "flnerf_posenc: bad arguments");
  if (n == 0) return 0;
  FL_LAUNCH(posenc_kernel, (unsigned)ceil_div64(n * (3 + 6 * L), 256), 256, 0, stream, n, L, x, out);
  return 0;
}

int flnerf_encode_f32(flnerf_ctx *ctx, int64_t B, int S, const float *rays11, const float *z, float *x90,
                      void *stream) {
  FL_REQUIRE(ctx && rays11 && z && x90 && S > 0 && B >= 0, "flnerf_encode_f32: bad arguments");
  if (B == 0) return 0;
  FL_LAUNCH(encode_f32_kernel, (unsigned)ceil_div64(B * S * 90, 256), 256, 0, stream, B, S, rays11, z, x90);
  return 0;
}

int64_t flnerf_padded_rows(int64_t n) { return ceil_div64(n, FLNERF_PAIR_ROWS) * FLNERF_PAIR_ROWS; }

int flnerf_encode_tc(flnerf_ctx *ctx, int64_t B, int S, const float *rays11, const float *z, void *pe_tiles,
                     float *dirpe, void *stream) {
  FL_REQUIRE(ctx && rays11 && z && pe_tiles && dirpe && S > 0 && B >= 0, "flnerf_encode_tc: bad arguments");
  if (B == 0) return 0;
  int64_t n = B * S, n_pad = flnerf_padded_rows(n);
  FrameArgs fa{};
  fa.dirpe = dirpe;
  FL_LAUNCH((encode_tc_kernel<false, false>), (unsigned)ceil_div64(n_pad, 128), 128, 0, stream, n, n_pad, S, rays11, z,
            (uint8_t *)pe_tiles, (size_t)0, fa);
  return 0;
}

int flnerf_encode_tc_x3(flnerf_ctx *ctx, int64_t B, int S, const float *rays11, const float *z, void *pe_tiles,
                        float *dirpe, void *stream) {
  FL_REQUIRE(ctx && rays11 && z && pe_tiles && dirpe && S > 0 && B >= 0, "flnerf_encode_tc_x3: bad arguments");
  if (B == 0) return 0;
  int64_t n = B * S, n_pad = flnerf_padded_rows(n);
  FrameArgs fa{};
  fa.dirpe = dirpe;
  FL_LAUNCH((encode_tc_kernel<true, false>), (unsigned)ceil_div64(n_pad, 128), 128, 0, stream, n, n_pad, S, rays11, z,
            (uint8_t *)pe_tiles, (size_t)(n_pad / 128) * 16384, fa);
  return 0;
}

int flnerf_encode_frame_tc(flnerf_ctx *ctx, int x3, int H, int W, const double *h_K, const float *h_c2w, float near_, float far_,
                           int ndc, int lindisp, int64_t pixel0, int64_t B, int S, const float *t_vals, float *rays11, float *z,
                           void *pe_tiles, float *dirpe, void *stream) {
  FL_REQUIRE(ctx && h_K && h_c2w && t_vals && rays11 && z && pe_tiles && dirpe && S > 0 && B >= 0 && H > 0 && W > 0 &&
                 pixel0 >= 0 && pixel0 + B <= (int64_t)H * W,
             "flnerf_encode_frame_tc: bad arguments");
  if (B == 0) return 0;
  FrameArgs fa;
  fa.cam = make_cam(h_K);
  for (int i = 0; i < 12; ++i) fa.pose.m[i] = h_c2w[i];
  fa.W = W; fa.ndc = ndc; fa.lindisp = lindisp; fa.near_ = near_; fa.far_ = far_;
  fa.sx = (float)(-1.0 / (W / (2.0 * h_K[0]))); fa.sy = (float)(-1.0 / (H / (2.0 * h_K[0])));   // focal = K[0][0] (render.py:71)
  fa.pixel0 = pixel0; fa.tv = t_vals; fa.rays11 = rays11; fa.z = z; fa.dirpe = dirpe;
  int64_t n = B * S, n_pad = flnerf_padded_rows(n);
  if (x3)
    FL_LAUNCH((encode_tc_kernel<true, true>), (unsigned)ceil_div64(n_pad, 128), 128, 0, stream, n, n_pad, S, nullptr, nullptr,
              (uint8_t *)pe_tiles, (size_t)(n_pad / 128) * 16384, fa);
  else
    FL_LAUNCH((encode_tc_kernel<false, true>), (unsigned)ceil_div64(n_pad, 128), 128, 0, stream, n, n_pad, S, nullptr, nullptr,
              (uint8_t *)pe_tiles, (size_t)0, fa);
  return 0;
}

int flnerf_pack_x90(flnerf_ctx *ctx, int64_t n, const float *x90, void *pe_tiles, float *dirpe, void *stream) {
  FL_REQUIRE(ctx && x90 && pe_tiles && dirpe && n >= 0, "flnerf_pack_x90: bad arguments");
  if (n == 0) return 0;
  int64_t n_pad = flnerf_padded_rows(n);
  FL_LAUNCH(pack_x90_kernel, (unsigned)ceil_div64(n_pad * 8, 256), 256, 0, stream, n, n_pad, x90, (uint8_t *)pe_tiles,
            dirpe, (size_t)0);
  return 0;
}

int flnerf_pack_x90_x3(flnerf_ctx *ctx, int64_t n, const float *x90, void *pe_tiles, float *dirpe, void *stream) {
  FL_REQUIRE(ctx && x90 && pe_tiles && dirpe && n >= 0, "flnerf_pack_x90_x3: bad arguments");
  if (n == 0) return 0;
  int64_t n_pad = flnerf_padded_rows(n);
  FL_LAUNCH(pack_x90_kernel, (unsigned)ceil_div64(n_pad * 8, 256), 256, 0, stream, n, n_pad, x90, (uint8_t *)pe_tiles,
            dirpe, (size_t)(n_pad / 128) * 16384);
  return 0;
}

int flnerf_pack_xrows(flnerf_ctx *ctx, int x3, int in_pts, int in_views, int64_t n, int S, const float *x, void *pe_tiles,
                      float *dirpe, void *stream) {
  FL_REQUIRE(ctx && x && pe_tiles && dirpe && n >= 0 && S > 0 && in_pts > 0 && in_pts <= 128 && in_views == 27,
             "flnerf_pack_xrows: bad arguments (in_views must be 27, in_pts <= 128)");
  if (n == 0) return 0;
  const int n_slabs = (in_pts + 63) / 64;
  int64_t n_pad = flnerf_padded_rows(n);
  FL_LAUNCH(pack_xrows_kernel, (unsigned)ceil_div64(n_pad * n_slabs * 8, 256), 256, 0, stream, n, n_pad, in_pts, n_slabs,
            in_pts + in_views, S, x, (uint8_t *)pe_tiles, dirpe, x3 ? (size_t)(n_pad / 128) * n_slabs * 16384 : (size_t)0);
  return 0;
}

int flnerf_gather_batch(flnerf_ctx *ctx, int64_t B, int64_t first, int64_t stride, const int32_t *ray_pix,
                        const int32_t *ray_gid, int cap, int H, int W, const double *h_K, const float *poses,
                        const float *images, float *rays_o, float *rays_d, float *target, int32_t *leaf_gid,
                        void *stream) {
  FL_REQUIRE(ctx && ray_pix && ray_gid && h_K && poses && images && rays_o && rays_d && target && cap > 0,
             "flnerf_gather_batch: bad arguments");
  if (B == 0) return 0;
  FL_LAUNCH(gather_batch_kernel, (unsigned)ceil_div64(B, 256), 256, 0, stream, B, first, stride, ray_pix, ray_gid, cap,
            H, W, make_cam(h_K), poses, (const void *)images, (const float *)nullptr, rays_o, rays_d, target, leaf_gid, ctx->step_rec);
  return 0;
}

int flnerf_gather_sub(flnerf_ctx *ctx, int64_t B, const float *ray_xy, const int32_t *ray_gid, int cap, int H, int W,
                      const double *h_K, const float *poses, const void *images, const float *lut256_or_null, float *rays_o,
                      float *rays_d, float *target, void *stream) {
  FL_REQUIRE(ctx && ray_xy && ray_gid && h_K && poses && images && rays_o && rays_d && target && cap > 0,
             "flnerf_gather_sub: bad arguments");
  if (B == 0) return 0;
  FL_LAUNCH(gather_sub_kernel, (unsigned)ceil_div64(B, 256), 256, 0, stream, B, ray_xy, ray_gid, cap, H, W, make_cam(h_K), poses,
            images, lut256_or_null, rays_o, rays_d, target);
  return 0;
}

int flnerf_gather_batch_u8(flnerf_ctx *ctx, int64_t B, int64_t first, int64_t stride, const int32_t *ray_pix,
                           const int32_t *ray_gid, int cap, int H, int W, const double *h_K, const float *poses,
                           const uint8_t *images, const float *lut256, float *rays_o, float *rays_d, float *target,
                           int32_t *leaf_gid, void *stream) {
  FL_REQUIRE(ctx && ray_pix && ray_gid && h_K && poses && images && lut256 && rays_o && rays_d && target && cap > 0,
             "flnerf_gather_batch_u8: bad arguments");
  if (B == 0) return 0;
  FL_LAUNCH(gather_batch_kernel, (unsigned)ceil_div64(B, 256), 256, 0, stream, B, first, stride, ray_pix, ray_gid, cap,
            H, W, make_cam(h_K), poses, (const void *)images, lut256, rays_o, rays_d, target, leaf_gid, ctx->step_rec);
  return 0;
}

}  // extern "C"
