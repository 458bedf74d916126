// train_ops.cu -- loss (+ per-leaf max table), Adam, and the GPU-resident quadtree ray selector.
// Reference: run_nerf_helpers.py:9 (img2mse), run_nerf.py:99,482-502 (loss, Adam),
// tree.py:17-94 (QuadTree), :569-626 (gen_rays_v3_1_subThread), :533-557,629-652 (adjust_tree).
#include "common.cuh"
#include <math.h>

namespace {

// ------------------------------------------------------------------------------------------------
// a9 + a13 accumulation.  One block (deterministic sums); B*3 elements is tiny next to the MLP.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
mse_leafmax_kernel(int64_t B, const float *__restrict__ rgb, const float *__restrict__ rgb0,
                   const float *__restrict__ target, float inv_count, const int32_t *__restrict__ leaf_gid,
                   float *__restrict__ loss_out, float *__restrict__ d_rgb, float *__restrict__ d_rgb0,
                   float *__restrict__ leaf_max) {
  __shared__ float red[2][32];
  float s_f = 0.f, s_c = 0.f;
  for (int64_t r = threadIdx.x; r < B; r += blockDim.x) {
    float mx = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float t = target[r * 3 + c];
      float df = rgb[r * 3 + c] - t;
      s_f += df * df;
      if (d_rgb) d_rgb[r * 3 + c] = 2.0f * df * inv_count;
      mx = fmaxf(mx, fabsf(t - rgb[r * 3 + c]));  // |gt - pred| (tree.py:538), max over channels (:642)
      if (rgb0) {
        float dc = rgb0[r * 3 + c] - t;
        s_c += dc * dc;
        if (d_rgb0) d_rgb0[r * 3 + c] = 2.0f * dc * inv_count;
      }
    }
    if (leaf_gid && leaf_max) {
      int g = leaf_gid[r];
      // non-negative floats order like their int patterns; the table is initialised to -1.0f
      if (g >= 0) atomicMax(reinterpret_cast<int *>(leaf_max) + g, __float_as_int(mx));
    }
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_f += __shfl_xor_sync(0xffffffffu, s_f, o);
    s_c += __shfl_xor_sync(0xffffffffu, s_c, o);
  }
  if (lane == 0) { red[0][w] = s_f; red[1][w] = s_c; }
  __syncthreads();
  if (w == 0) {
    s_f = lane < (int)(blockDim.x >> 5) ? red[0][lane] : 0.f;
    s_c = lane < (int)(blockDim.x >> 5) ? red[1][lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_f += __shfl_xor_sync(0xffffffffu, s_f, o);
      s_c += __shfl_xor_sync(0xffffffffu, s_c, o);
    }
    if (lane == 0) {
      loss_out[0] = s_f * inv_count;
      loss_out[1] = s_c * inv_count;
    }
  }
}

// The refinement statistic of the nerf++ / plenoxels copies of the quadtree (nerf++-ours/tree.py:613-622: a leaf splits when
// the MEAN of |gt - pred| over its rays and channels exceeds the threshold; nerf-ours uses the max, tree.py:642).  Sums are
// accumulated in double (the order of the atomics cannot matter), counts in int32.
__global__ void leaf_sum_kernel(int64_t B, const float *__restrict__ pred, const float *__restrict__ target,
                                const int32_t *__restrict__ leaf_gid, double *__restrict__ leaf_sum,
                                int32_t *__restrict__ leaf_cnt) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const int g = leaf_gid[r];
  if (g < 0) return;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) s += fabsf(target[r * 3 + c] - pred[r * 3 + c]);
  atomicAdd(leaf_sum + g, (double)s);
  atomicAdd(leaf_cnt + g, 1);
}
// leaf_stat[g] = mean, or -1 for a leaf without rays (the reference's NaN mean never splits either)
__global__ void leaf_mean_kernel(int64_t n, const double *__restrict__ leaf_sum, const int32_t *__restrict__ leaf_cnt,
                                 float *__restrict__ leaf_stat) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int c = leaf_cnt[g];
  leaf_stat[g] = c > 0 ? (float)(leaf_sum[g] / (3.0 * c)) : -1.0f;
}

// torch.optim.Adam single-tensor update order (exp_avg.lerp, exp_avg_sq.mul.addcmul, sqrt/bc2_sqrt + eps, addcdiv)
__global__ void adam_kernel(int64_t n, float *__restrict__ p, float *__restrict__ m, float *__restrict__ v,
                            const float *__restrict__ g, float omb1, float b2, float omb2, float eps, float step_size,
                            float bc2_sqrt, const flnerf_step_record *__restrict__ rec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (rec) { step_size = rec->adam_step_size; bc2_sqrt = rec->adam_bc2_sqrt; }
  float gi = g[i];
  float mi = m[i] + (gi - m[i]) * omb1;
  float vi = v[i] * b2 + omb2 * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float den = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - step_size * (mi / den);
}

// ------------------------------------------------------------------------------------------------
// quadtree
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void child_box(const double b[4], int q, double out[4]) {
  // tree.py:61-72: TL (x0,y0,mx,my), (mx,y0,x1,my), (x0,my,mx,y1), (mx,my,x1,y1)
  double mx = (b[0] + b[2]) / 2, my = (b[1] + b[3]) / 2;
  out[0] = (q & 1) ? mx : b[0];
  out[2] = (q & 1) ? b[2] : mx;
  out[1] = (q & 2) ? my : b[1];
  out[3] = (q & 2) ? b[3] : my;
}

__global__ void qt_init_kernel(int n_images, int cap, int H, int W, int levels, int leaves, double *__restrict__ boxes,
                               int32_t *__restrict__ count, double *__restrict__ min_area) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n_images * leaves) return;
  int img = (int)(idx / leaves), j = (int)(idx % leaves);
  double b[4] = {0.0, 0.0, (double)H, (double)W}, c[4];
  for (int l = levels - 1; l >= 0; --l) {  // most significant base-4 digit = first split (DFS order)
    child_box(b, (j >> (2 * l)) & 3, c);
    b[0] = c[0]; b[1] = c[1]; b[2] = c[2]; b[3] = c[3];
  }
  double *o = boxes + ((int64_t)img * cap + j) * 4;
  o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3];
  if (j == 0) {
    count[img] = leaves;
    min_area[img] = (double)H * (double)W / (double)leaves;  // tree.py:94
  }
}

// block-wide exclusive scan of one int per thread (blockDim.x = 256)
__device__ __forceinline__ int block_excl_scan(int v, int *total, int *s_warp) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    int x = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
    int xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    s_warp[lane] = xi - x;  // exclusive warp offsets
    if (lane == 31) s_warp[32] = xi;
  }
  __syncthreads();
  int res = s_warp[w] + incl - v;
  *total = s_warp[32];
  __syncthreads();
  return res;
}

// one block per image
__global__ void __launch_bounds__(256)
qt_refine_kernel(int cap, const double *__restrict__ boxes_in, const int32_t *__restrict__ count_in,
                 double *__restrict__ min_area, const float *__restrict__ leaf_max, float thres,
                 double *__restrict__ boxes_out, int32_t *__restrict__ count_out) {
  __shared__ int s_warp[33];
  __shared__ int s_any;
  int img = blockIdx.x;
  int n = count_in[img];
  double ma = min_area[img];
  const double *bi = boxes_in + (int64_t)img * cap * 4;
  double *bo = boxes_out + (int64_t)img * cap * 4;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  int base_out = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    int j = base + threadIdx.x;
    double b[4] = {0, 0, 0, 0};
    int split = 0;
    if (j < n) {
      b[0] = bi[j * 4]; b[1] = bi[j * 4 + 1]; b[2] = bi[j * 4 + 2]; b[3] = bi[j * 4 + 3];
      double area = (b[2] - b[0]) * (b[3] - b[1]);
      // tree.py:642-646: loss.max() > thres (fp32 compare) and area == minArea (float equality)
      split = (leaf_max[(int64_t)img * cap + j] > thres) && (area == ma);
    }
    int total;
    int pos = base_out + block_excl_scan(j < n ? (split ? 4 : 1) : 0, &total, s_warp);
    if (j < n) {
      if (split) {
        s_any = 1;
        for (int q = 0; q < 4; ++q) {
          if (pos + q < cap) child_box(b, q, bo + (int64_t)(pos + q) * 4);
        }
      } else if (pos < cap) {
        bo[pos * 4] = b[0]; bo[pos * 4 + 1] = b[1]; bo[pos * 4 + 2] = b[2]; bo[pos * 4 + 3] = b[3];
      }
    }
    base_out += total;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    count_out[img] = min(base_out, cap);
    if (s_any) min_area[img] = ma / 4;  // tree.py:649-650 (once per tree per adjust)
  }
}

// single block: per-leaf ray counts + exclusive scan over all (image, leaf) slots
__global__ void __launch_bounds__(256)
qt_count_kernel(int n_images, int cap, const double *__restrict__ boxes, const int32_t *__restrict__ count,
                const double *__restrict__ min_area, double rays_per_pixel, int64_t *__restrict__ ray_offset) {
  __shared__ int s_warp[33];
  int64_t running = 0;
  int64_t slots = (int64_t)n_images * cap;
  for (int64_t base = 0; base < slots; base += blockDim.x) {
    int64_t s = base + threadIdx.x;
    int c = 0;
    if (s < slots) {
      int img = (int)(s / cap), j = (int)(s % cap);
      if (j < count[img]) {
        const double *b = boxes + s * 4;
        double area = (b[2] - b[0]) * (b[3] - b[1]);
        c = (area > min_area[img] + 0.01) ? 10 : (int)(area * rays_per_pixel);  // tree.py:578-581
      }
    }
    int total;
    int ex = block_excl_scan(c, &total, s_warp);
    if (s < slots) ray_offset[s] = running + ex;
    running += total;
  }
  if (threadIdx.x == 0) ray_offset[slots] = running;
}

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// keyed bijection of [0,N): balanced Feistel network on 2*half bits + cycle walking
__host__ __device__ inline uint64_t feistel_perm(uint64_t j, uint64_t N, int half, uint64_t seed) {
  uint64_t mask = (1ull << half) - 1;
  uint64_t x = j;
  do {
    uint32_t L = (uint32_t)(x >> half), R = (uint32_t)(x & mask);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      uint32_t k = (uint32_t)(seed >> ((r & 1) * 32)) + 0x9E3779B9u * (uint32_t)(r + 1);
      uint32_t f = R * 0x85EBCA6Bu + k;
      f ^= f >> 16; f *= 0x7feb352dU; f ^= f >> 15; f *= 0x846ca68bU; f ^= f >> 16;
      uint32_t nl = R;
      R = (L ^ f) & (uint32_t)mask;
      L = nl;
    }
    x = ((uint64_t)L << half) | R;
  } while (x >= N);
  return x;
}

__global__ void qt_emit_kernel(int n_images, int cap, int W, const double *__restrict__ boxes,
                               const int64_t *__restrict__ ray_offset, int64_t N, int half, uint64_t seed,
                               int32_t *__restrict__ ray_pix, int32_t *__restrict__ ray_gid) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  // slot = last s with ray_offset[s] <= j  (empty slots have equal offsets, upper bound skips them)
  int64_t lo = 0, hi = (int64_t)n_images * cap;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (ray_offset[mid] <= j) lo = mid; else hi = mid;
  }
  const double *b = boxes + lo * 4;
  // tree.py:598-599: x ~ randint[ceil(x0), ceil(x1)), y ~ randint[ceil(y0), ceil(y1-0.01)), with replacement
  int x_lo = (int)ceil(b[0]), x_hi = (int)ceil(b[2]);
  int y_lo = (int)ceil(b[1]), y_hi = (int)ceil(b[3] - 0.01);
  uint32_t r[4];
  philox4x32(seed, (uint64_t)j, 0x9E17ull, r);
  int row = x_lo + (int)(r[0] % (uint32_t)max(1, x_hi - x_lo));
  int col = y_lo + (int)(r[1] % (uint32_t)max(1, y_hi - y_lo));
  int64_t pos = (int64_t)feistel_perm((uint64_t)j, (uint64_t)N, half, seed ^ 0xA5A5A5A55A5A5A5Aull);
  ray_pix[pos] = row * W + col;
  ray_gid[pos] = (int32_t)lo;
}

// gen_rays_v3 (tree.py:231-307): SUB-PIXEL positions on a 1/1000 grid -- x = randint[int(1000 x0), int(1000 (x1 - 0.01))) / 1000,
// y likewise (:265-268) -- written as float pairs at the shuffled position
__global__ void qt_emit_sub_kernel(int n_images, int cap, const double *__restrict__ boxes, const int64_t *__restrict__ ray_offset,
                                   int64_t N, int half, uint64_t seed, float *__restrict__ ray_xy, int32_t *__restrict__ ray_gid) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  int64_t lo = 0, hi = (int64_t)n_images * cap;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (ray_offset[mid] <= j) lo = mid; else hi = mid;
  }
  const double *b = boxes + lo * 4;
  const int x_lo = (int)(b[0] * 1000), x_hi = (int)((b[2] - 0.01) * 1000);
  const int y_lo = (int)(b[1] * 1000), y_hi = (int)((b[3] - 0.01) * 1000);
  uint32_t r[4];
  philox4x32(seed, (uint64_t)j, 0x9E18ull, r);
  const int kx = x_lo + (int)(r[0] % (uint32_t)max(1, x_hi - x_lo));
  const int ky = y_lo + (int)(r[1] % (uint32_t)max(1, y_hi - y_lo));
  int64_t pos = (int64_t)feistel_perm((uint64_t)j, (uint64_t)N, half, seed ^ 0xA5A5A5A55A5A5A5Aull);
  ray_xy[pos * 2] = __fdiv_rn((float)kx, 1000.0f);
  ray_xy[pos * 2 + 1] = __fdiv_rn((float)ky, 1000.0f);
  ray_gid[pos] = (int32_t)lo;
}

// ---------------------------------------------------------------------------------------------
// probability-guided pixel sampling (prob=True): image_process.py:26-96 + tree.py:583-595
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {  // cv2 BORDER_REFLECT_101 (the cv2.blur default)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i < 0 ? 0 : i;
}

// get_sharp_img (image_process.py:26-39): per channel sqrt|blur3x3(x^2) - blur3x3(x)^2|, then the BGR2GRAY weights
// on the channel-reversed image = 0.299 R + 0.587 G + 0.114 B.  cv2.blur sums a float image in double.
__global__ void sharp_map_kernel(int n_images, int H, int W, const float *__restrict__ images, float *__restrict__ sharp) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)n_images * H * W;
  if (idx >= total) return;
  int img = (int)(idx / ((int64_t)H * W));
  int rem = (int)(idx % ((int64_t)H * W));
  int r = rem / W, c = rem % W;
  const float *im = images + (int64_t)img * H * W * 3;
  double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
  for (int dr = -1; dr <= 1; ++dr)
    for (int dc = -1; dc <= 1; ++dc) {
      const float *px = im + ((int64_t)reflect101(r + dr, H) * W + reflect101(c + dc, W)) * 3;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float v = px[ch];
        s1[ch] += (double)v;
        s2[ch] += (double)__fmul_rn(v, v);  // img ** 2 is a float32 array
      }
    }
  float g[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float e2 = (float)(s2[ch] * (1.0 / 9.0)), e1 = (float)(s1[ch] * (1.0 / 9.0));
    g[ch] = sqrtf(fabsf(__fsub_rn(e2, __fmul_rn(e1, e1))));
  }
  sharp[idx] = g[2] * 0.114f + g[1] * 0.587f + g[0] * 0.299f;
}

// per-slot block heights int(x1) - int(x0) (tree.py:587: sharp[int(x0):int(x1), int(y0):int(y1)]) -> exclusive scan
__global__ void __launch_bounds__(256)
qt_prob_rows_kernel(int n_images, int cap, const double *__restrict__ boxes, const int32_t *__restrict__ count,
                    int64_t *__restrict__ row_offset) {
  __shared__ int s_warp[33];
  int64_t running = 0;
  int64_t slots = (int64_t)n_images * cap;
  for (int64_t base = 0; base < slots; base += blockDim.x) {
    int64_t s = base + threadIdx.x;
    int c = 0;
    if (s < slots && (int)(s % cap) < count[s / cap]) {
      const double *b = boxes + s * 4;
      c = max(0, (int)b[2] - (int)b[0]);
    }
    int total;
    int ex = block_excl_scan(c, &total, s_warp);
    if (s < slots) row_offset[s] = running + ex;
    running += total;
  }
  if (threadIdx.x == 0) row_offset[slots] = running;
}

// to_prob_v2 (image_process.py:59-74) of one leaf block: g = gray + 1e-6 (float64); the weights are
// clip(g, 0.01*mean(g), max(g)) / max(g) / sum -- the common scale cancels in sampling, so w = max(g, 0.01*mean(g)).
// One CTA per leaf: thr, then the inclusive prefix of the per-row weight sums (float64).
__global__ void __launch_bounds__(256)
qt_prob_prepare_kernel(int cap, int H, int W, const double *__restrict__ boxes, const int32_t *__restrict__ count,
                       const float *__restrict__ sharp, const int64_t *__restrict__ row_offset,
                       double *__restrict__ row_cdf, double *__restrict__ leaf_thr) {
  const int64_t s = blockIdx.x;
  const int img = (int)(s / cap), j = (int)(s % cap);
  if (j >= count[img]) return;
  const double *b = boxes + s * 4;
  const int X0 = (int)b[0], X1 = (int)b[2], Y0 = (int)b[1], Y1 = (int)b[3];
  const int h = X1 - X0, w = Y1 - Y0;
  if (h <= 0 || w <= 0) { if (threadIdx.x == 0) leaf_thr[s] = 0.0; return; }
  const float *g = sharp + (int64_t)img * H * W;
  __shared__ double s_red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) acc += (double)g[(int64_t)(X0 + i / w) * W + Y0 + i % w] + 1e-6;
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  const double thr = 0.01 * (s_red[0] / (double)(h * w));
  __syncthreads();
  double *cdf = row_cdf + row_offset[s];
  // row sums: one warp per row (lanes stride the columns), then a serial prefix by thread 0 (h <= a few hundred)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < h; r += 8) {
    double rs = 0.0;
    for (int c = lane; c < w; c += 32) rs += fmax((double)g[(int64_t)(X0 + r) * W + Y0 + c] + 1e-6, thr);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
    if (lane == 0) cdf[r] = rs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0.0;
    for (int r = 0; r < h; ++r) { run += cdf[r]; cdf[r] = run; }
    leaf_thr[s] = thr;
  }
}

// tree.py:583-599 with prob=True: of a leaf's ray_num rays the first int(ray_num*(1-rand_frac)) are drawn from the
// sharpness distribution of its block (np.random.choice(p): inverse CDF over the row-major flattened block, offset
// (int(x0), int(y0))), the rest uniformly as in the prob=False path.  u (optional, [N][2]) replaces Philox.
__global__ void qt_emit_prob_kernel(int n_images, int cap, int H, int W, const double *__restrict__ boxes,
                                    const int64_t *__restrict__ ray_offset, int64_t N, int half, uint64_t seed,
                                    double rand_frac, const float *__restrict__ sharp,
                                    const int64_t *__restrict__ row_offset, const double *__restrict__ row_cdf,
                                    const double *__restrict__ leaf_thr, const float *__restrict__ u, int shuffle,
                                    int32_t *__restrict__ ray_pix, int32_t *__restrict__ ray_gid) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  int64_t lo = 0, hi = (int64_t)n_images * cap;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (ray_offset[mid] <= j) lo = mid; else hi = mid;
  }
  const double *b = boxes + lo * 4;
  const int ray_num = (int)(ray_offset[lo + 1] - ray_offset[lo]);
  const int ray_num1 = (int)((double)ray_num * (1.0 - rand_frac));  // int(ray_num * (1 - randSamp_proc))
  const int k = (int)(j - ray_offset[lo]);
  double u0, u1;
  if (u) {
    u0 = (double)u[j * 2]; u1 = (double)u[j * 2 + 1];
  } else {
    uint32_t r[4];
    philox4x32(seed, (uint64_t)j, 0x9E17ull, r);
    u0 = ((double)r[0] * 4294967296.0 + (double)r[1]) * (1.0 / 18446744073709551616.0);  // 64-bit uniform, like numpy
    u1 = (double)u32_to_unit(r[2]);
  }
  int row, col;
  const int X0 = (int)b[0], X1 = (int)b[2], Y0 = (int)b[1], Y1 = (int)b[3];
  const int h = X1 - X0, w = Y1 - Y0;
  if (k < ray_num1 && h > 0 && w > 0) {
    const double *cdf = row_cdf + row_offset[lo];
    const double thr = leaf_thr[lo];
    const double target = u0 * cdf[h - 1];
    int a = 0, e = h - 1;   // first row whose inclusive prefix exceeds the target (searchsorted side='right')
    while (a < e) {
      int m = (a + e) >> 1;
      if (cdf[m] > target) e = m; else a = m + 1;
    }
    double run = a > 0 ? cdf[a - 1] : 0.0;
    const float *g = sharp + ((int64_t)(lo / cap) * H + X0 + a) * W + Y0;
    int c = 0;
    for (; c < w - 1; ++c) {
      run += fmax((double)g[c] + 1e-6, thr);
      if (run > target) break;
    }
    row = X0 + a;
    col = Y0 + c;
  } else {
    int x_lo = (int)ceil(b[0]), x_hi = (int)ceil(b[2]);
    int y_lo = (int)ceil(b[1]), y_hi = (int)ceil(b[3] - 0.01);
    if (u) {
      row = x_lo + min(max(1, x_hi - x_lo) - 1, (int)(u0 * (double)max(1, x_hi - x_lo)));
      col = y_lo + min(max(1, y_hi - y_lo) - 1, (int)(u1 * (double)max(1, y_hi - y_lo)));
    } else {
      uint32_t r[4];
      philox4x32(seed, (uint64_t)j, 0x9E18ull, r);
      row = x_lo + (int)(r[0] % (uint32_t)max(1, x_hi - x_lo));
      col = y_lo + (int)(r[1] % (uint32_t)max(1, y_hi - y_lo));
    }
  }
  int64_t pos = shuffle ? (int64_t)feistel_perm((uint64_t)j, (uint64_t)N, half, seed ^ 0xA5A5A5A55A5A5A5Aull) : j;
  ray_pix[pos] = row * W + col;
  ray_gid[pos] = (int32_t)lo;
}

}  // namespace

extern "C" {

int flnerf_mse_leafmax(flnerf_ctx *ctx, int64_t B, const float *rgb, const float *rgb0, const float *target,
                       int64_t denom, const int32_t *leaf_gid, float *loss_out, float *d_rgb, float *d_rgb0,
                       float *leaf_max, void *stream) {
  FL_REQUIRE(ctx && rgb && target && loss_out && denom > 0 && B >= 0, "flnerf_mse_leafmax: bad arguments");
  float inv = (float)(1.0 / (3.0 * (double)denom));
  FL_LAUNCH(mse_leafmax_kernel, 1, 1024, 0, stream, B, rgb, rgb0, target, inv, leaf_gid, loss_out, d_rgb, d_rgb0,
            leaf_max);
  return 0;
}

int flnerf_leaf_sum(flnerf_ctx *ctx, int64_t B, const float *pred, const float *target, const int32_t *leaf_gid,
                    double *leaf_sum, int32_t *leaf_cnt, void *stream) {
  FL_REQUIRE(ctx && pred && target && leaf_gid && leaf_sum && leaf_cnt && B >= 0, "flnerf_leaf_sum: bad arguments");
  if (B == 0) return 0;
  FL_LAUNCH(leaf_sum_kernel, (unsigned)ceil_div64(B, 256), 256, 0, stream, B, pred, target, leaf_gid, leaf_sum, leaf_cnt);
  return 0;
}

int flnerf_leaf_mean(flnerf_ctx *ctx, int64_t n_slots, const double *leaf_sum, const int32_t *leaf_cnt, float *leaf_stat,
                     void *stream) {
  FL_REQUIRE(ctx && leaf_sum && leaf_cnt && leaf_stat && n_slots >= 0, "flnerf_leaf_mean: bad arguments");
  if (n_slots == 0) return 0;
  FL_LAUNCH(leaf_mean_kernel, (unsigned)ceil_div64(n_slots, 256), 256, 0, stream, n_slots, leaf_sum, leaf_cnt, leaf_stat);
  return 0;
}

int flnerf_adam_step(flnerf_ctx *ctx, int64_t n, float *param, float *m, float *v, const float *grad, double lr,
                     double b1, double b2, double eps, int64_t t, void *stream) {
  FL_REQUIRE(ctx && param && m && v && grad && n >= 0 && t >= 1, "flnerf_adam_step: bad arguments");
  if (n == 0) return 0;
  double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
  FL_LAUNCH(adam_kernel, (unsigned)ceil_div64(n, 256), 256, 0, stream, n, param, m, v, grad, (float)(1.0 - b1),
            (float)b2, (float)(1.0 - b2), (float)eps, (float)(lr / bc1), (float)sqrt(bc2), ctx->step_rec);
  return 0;
}

__global__ void step_record_kernel(flnerf_step_record *rec, flnerf_step_record v) { *rec = v; }

int flnerf_set_step_record(flnerf_ctx *ctx, const flnerf_step_record *rec) {
  FL_REQUIRE(ctx, "flnerf_set_step_record: bad arguments");
  ctx->step_rec = rec;
  return 0;
}

int flnerf_step_record_write(flnerf_ctx *ctx, flnerf_step_record *rec, int64_t first, uint64_t rng_offset, double lr,
                             double b1, double b2, int64_t t, void *stream) {
  FL_REQUIRE(ctx && rec && t >= 1, "flnerf_step_record_write: bad arguments");
  flnerf_step_record v;
  v.first = first;
  v.rng_offset = rng_offset;
  v.adam_step_size = (float)(lr / (1.0 - pow(b1, (double)t)));     // the same host-side doubles as flnerf_adam_step
  v.adam_bc2_sqrt = (float)sqrt(1.0 - pow(b2, (double)t));
  FL_LAUNCH(step_record_kernel, 1, 1, 0, stream, rec, v);
  return 0;
}

int flnerf_qt_init(flnerf_ctx *ctx, int n_images, int cap, int H, int W, int max_depth, double *boxes, int32_t *count,
                   double *min_area, void *stream) {
  FL_REQUIRE(ctx && boxes && count && min_area && n_images > 0 && max_depth >= 1 && max_depth <= 12,
             "flnerf_qt_init: bad arguments");
  int levels = max_depth - 1;
  int leaves = 1 << (2 * levels);
  FL_REQUIRE(leaves <= cap, "flnerf_qt_init: capacity %d < %d leaves", cap, leaves);
  FL_LAUNCH(qt_init_kernel, (unsigned)ceil_div64((int64_t)n_images * leaves, 256), 256, 0, stream, n_images, cap, H, W,
            levels, leaves, boxes, count, min_area);
  return 0;
}

int flnerf_qt_refine(flnerf_ctx *ctx, int n_images, int cap, const double *boxes_in, const int32_t *count_in,
                     double *min_area, const float *leaf_max, float thres, double *boxes_out, int32_t *count_out,
                     void *stream) {
  FL_REQUIRE(ctx && boxes_in && count_in && min_area && leaf_max && boxes_out && count_out && boxes_in != boxes_out,
             "flnerf_qt_refine: bad arguments");
  FL_LAUNCH(qt_refine_kernel, n_images, 256, 0, stream, cap, boxes_in, count_in, min_area, leaf_max, thres, boxes_out,
            count_out);
  return 0;
}

int flnerf_qt_count(flnerf_ctx *ctx, int n_images, int cap, const double *boxes, const int32_t *count,
                    const double *min_area, double rays_per_pixel, int64_t *ray_offset, void *stream) {
  FL_REQUIRE(ctx && boxes && count && min_area && ray_offset, "flnerf_qt_count: bad arguments");
  FL_LAUNCH(qt_count_kernel, 1, 256, 0, stream, n_images, cap, boxes, count, min_area, rays_per_pixel, ray_offset);
  return 0;
}

int flnerf_qt_emit(flnerf_ctx *ctx, int n_images, int cap, int W, const double *boxes, const int32_t *count,
                   const int64_t *ray_offset, int64_t n_rays, uint64_t seed, int32_t *ray_pix, int32_t *ray_gid,
                   void *stream) {
  (void)count;
  FL_REQUIRE(ctx && boxes && ray_offset && ray_pix && ray_gid && n_rays >= 0, "flnerf_qt_emit: bad arguments");
  if (n_rays == 0) return 0;
  int bits = 2;
  while ((1ull << bits) < (uint64_t)n_rays) ++bits;
  int half = (bits + 1) / 2;
  FL_LAUNCH(qt_emit_kernel, (unsigned)ceil_div64(n_rays, 256), 256, 0, stream, n_images, cap, W, boxes, ray_offset,
            n_rays, half, seed, ray_pix, ray_gid);
  return 0;
}

int flnerf_qt_emit_sub(flnerf_ctx *ctx, int n_images, int cap, const double *boxes, const int64_t *ray_offset, int64_t n_rays,
                       uint64_t seed, float *ray_xy, int32_t *ray_gid, void *stream) {
  FL_REQUIRE(ctx && boxes && ray_offset && ray_xy && ray_gid && n_rays >= 0, "flnerf_qt_emit_sub: bad arguments");
  if (n_rays == 0) return 0;
  int bits = 2;
  while ((1ull << bits) < (uint64_t)n_rays) ++bits;
  FL_LAUNCH(qt_emit_sub_kernel, (unsigned)ceil_div64(n_rays, 256), 256, 0, stream, n_images, cap, boxes, ray_offset, n_rays,
            (bits + 1) / 2, seed, ray_xy, ray_gid);
  return 0;
}

int flnerf_sharp_map(flnerf_ctx *ctx, int n_images, int H, int W, const float *images, float *sharp, void *stream) {
  FL_REQUIRE(ctx && images && sharp && n_images > 0 && H > 1 && W > 1, "flnerf_sharp_map: bad arguments");
  FL_LAUNCH(sharp_map_kernel, (unsigned)ceil_div64((int64_t)n_images * H * W, 256), 256, 0, stream, n_images, H, W, images,
            sharp);
  return 0;
}

int flnerf_qt_prob_rows(flnerf_ctx *ctx, int n_images, int cap, const double *boxes, const int32_t *count,
                        int64_t *row_offset, void *stream) {
  FL_REQUIRE(ctx && boxes && count && row_offset, "flnerf_qt_prob_rows: bad arguments");
  FL_LAUNCH(qt_prob_rows_kernel, 1, 256, 0, stream, n_images, cap, boxes, count, row_offset);
  return 0;
}

int flnerf_qt_prob_prepare(flnerf_ctx *ctx, int n_images, int cap, int H, int W, const double *boxes,
                           const int32_t *count, const float *sharp, const int64_t *row_offset, double *row_cdf,
                           double *leaf_thr, void *stream) {
  FL_REQUIRE(ctx && boxes && count && sharp && row_offset && row_cdf && leaf_thr, "flnerf_qt_prob_prepare: bad arguments");
  FL_LAUNCH(qt_prob_prepare_kernel, (unsigned)((int64_t)n_images * cap), 256, 0, stream, cap, H, W, boxes, count, sharp,
            row_offset, row_cdf, leaf_thr);
  return 0;
}

int flnerf_qt_emit_prob(flnerf_ctx *ctx, int n_images, int cap, int H, int W, const double *boxes, const int32_t *count,
                        const int64_t *ray_offset, int64_t n_rays, uint64_t seed, double rand_frac, const float *sharp,
                        const int64_t *row_offset, const double *row_cdf, const double *leaf_thr, const float *u,
                        int shuffle, int32_t *ray_pix, int32_t *ray_gid, void *stream) {
  (void)count;
  FL_REQUIRE(ctx && boxes && ray_offset && ray_pix && ray_gid && sharp && row_offset && row_cdf && leaf_thr && n_rays >= 0 &&
                 rand_frac >= 0.0 && rand_frac <= 1.0,
             "flnerf_qt_emit_prob: bad arguments");
  if (n_rays == 0) return 0;
  int bits = 2;
  while ((1ull << bits) < (uint64_t)n_rays) ++bits;
  int half = (bits + 1) / 2;
  FL_LAUNCH(qt_emit_prob_kernel, (unsigned)ceil_div64(n_rays, 256), 256, 0, stream, n_images, cap, H, W, boxes, ray_offset,
            n_rays, half, seed, rand_frac, sharp, row_offset, row_cdf, leaf_thr, u, shuffle, ray_pix, ray_gid);
  return 0;
}

}  // extern "C"
