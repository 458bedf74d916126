// mlp_tc.cu -- FLNERF_MODE_BF16: the NeRF MLP (model.py:38-63) on the 5th-gen tensor cores.
//
//   * persistent, warp-specialised kernels on CTA PAIRS (clusters of 2, tcgen05.mma.cta_group::2, M = 256 = 128 rows
//     in each CTA): warp 0 = bulk-copy producer (UBLKCP into an mbarrier ring), warp 1 = single-thread MMA issuer in
//     the leader CTA / "stage full" relay in the follower, warps 2..17 = epilogue (tcgen05.ld -> bias/ReLU/bf16 -> smem);
//   * the activations of a 128-row tile never leave the SM between layers: the epilogue writes the next layer's A
//     operand straight into shared memory in the SWIZZLE_128B K-major image the UMMA descriptor reads; every CTA carries
//     TWO 128-row tiles (tile sets A and B of the pair) half a period apart: A's MMAs overlap B's epilogue;
//   * accumulators live in TMEM (2 tiles x 256 fp32 columns = all 512 columns);
//   * weights are pre-packed once per optimiser step (flnerf_mlp_pack_weights) into the exact shared-memory image; each
//     CTA of a pair fetches only ITS half of a chunk (N/2 rows, 16 KB, one contiguous bulk copy -- no tensor maps, no
//     driver API) and a layer's chunks stay in the 5 x 16 KB ring for both tile sets.  Shared-memory traffic per
//     128-row tile-layer: 128 KB operand reads + 32 KB chunk writes + 64 KB epilogue stores (+ 64 KB read by the stash
//     bulk stores) against 128 B/clk x 2048 clk = 256 KB at the MMA floor (a single-CTA M=128 design moves 384..448 KB);
//   * for training every epilogue warp bulk-stores its own 4 KB pieces of the activation tile (and a ReLU bitmask) to
//     HBM; the same image is read back as an MN-major operand by the weight-gradient kernel.
//
// Three kernels: mlp_fwd_tc (10 tensor layers + alpha/rgb heads on CUDA cores in the epilogue),
// mlp_dgrad_tc (9 tensor layers, ReLU masks from the bitmask stash), mlp_wgrad_tc (per-layer dY^T X with the
// 256x256 fp32 accumulator resident in TMEM over a row range, flushed once with red.global.add).
#include "common.cuh"
#include "mlp_layout.h"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace tc {

using namespace mlp_layout;

// CTA = 18 warps: warp 0 producer, warp 1 MMA issuer / relay, warps 2..17 epilogue (4 lane quarters x 4 column
// quarters, each warp serving both tiles)
constexpr int kEpiWarps = 16;
constexpr int kThreads = (2 + kEpiWarps) * 32;  // 576
constexpr int kEpiWarp0 = 2;
constexpr uint32_t ACT_BYTES = 65536, SLAB_BYTES = 16384, PE_BYTES = 16384, WSTAGE = 16384;
// biases of pts_linears.0..7 + feature_linear, then fp32 copies of the small heads: W_rgb[3][128], b_rgb[3],
// b_alpha[1], w_alpha[256]
constexpr uint32_t BIAS_FLOATS = 9 * 256, HEAD_FLOATS = 384 + 4 + 256;
constexpr uint32_t OFF_ACT = 0, OFF_W = 2 * ACT_BYTES;
// shared-memory layout after the two activation tiles: ring of kNst x 16 KB stages (half weight chunks AND the PE slabs
// travel through it), [bias table], heads, barrier block, [alpha partial sums].  The forward kernel trades one ring
// stage for its resident bias table (5 stages); dgrad needs neither biases nor alpha sums and runs 6 stages.
template <int kNst, bool kFwd>
struct Lay {
  static constexpr int NST = kNst;
  static constexpr uint32_t OFF_VEC = OFF_W + kNst * WSTAGE;
  static constexpr uint32_t OFF_HEAD = OFF_VEC + (kFwd ? BIAS_FLOATS * 4 : 0);
  static constexpr uint32_t OFF_BAR = OFF_HEAD + HEAD_FLOATS * 4;
  static constexpr uint32_t OFF_ALPHA = OFF_BAR + 256;  // [2 tiles][4 column quarters][128 rows] fp32 partial alpha sums
  static constexpr uint32_t SMEM = OFF_ALPHA + (kFwd ? 2 * 4 * 128 * 4 : 0);
  static_assert(OFF_BAR % 16 == 0 && SMEM <= 232448, "shared memory budget");
  static_assert(8u * (2 * kNst + 4) <= 128u, "barrier block: the TMEM slot sits at +128");
  // mbarrier addresses (bytes from the barrier block): no arrays, so nothing lands in local memory
  __device__ static __forceinline__ uint32_t w_full(uint32_t base, uint32_t i) { return base + 8u * i; }
  __device__ static __forceinline__ uint32_t w_empty(uint32_t base, uint32_t i) { return base + 8u * (kNst + i); }
  __device__ static __forceinline__ uint32_t acc_full(uint32_t base, uint32_t t) { return base + 8u * (2 * kNst + t); }
  __device__ static __forceinline__ uint32_t act_ready(uint32_t base, uint32_t t) { return base + 8u * (2 * kNst + 2 + t); }
  // the issuer addresses ring items by their running index (a layer's items are used twice, by tile set A then B)
  __device__ static __forceinline__ uint32_t item_stage(uint32_t q) { return q % kNst; }
  __device__ static __forceinline__ uint32_t item_phase(uint32_t q) { return (q / kNst) & 1u; }
  struct Ring {
    uint32_t stage = 0, phase = 0;
    __device__ __forceinline__ void next() {
      if (++stage == (uint32_t)kNst) { stage = 0; phase ^= 1; }
    }
  };
};
using LayF = Lay<5, true>;
using LayD = Lay<6, false>;

// packed weight image of one net
constexpr int FWD_CHUNKS = 38, DG_CHUNKS = 34;
constexpr size_t FWD_BYTES = (size_t)34 * 32768 + 4 * 16384;
constexpr size_t DG_BYTES = (size_t)34 * 32768;
constexpr size_t PACKED_BYTES = FWD_BYTES + DG_BYTES;

// Where a network's tensors live inside its flat fp32 parameter buffer (parameters() order) and how many 64-wide slabs its
// positional input takes.  kind 0: 63 position channels (1 slab) -- nerf-ours model.NeRF and the nerf++ foreground MLPNet;
// kind 1: 84 channels (2 slabs, PE of (x, y, z, 1/r)) -- the nerf++ background MLPNet (nerf++-ours/nerf_network.py:70-142).
// Everything behind the first layer and the skip concatenation has the same shape in both.
constexpr int kNetKinds = 2;
constexpr int MAX_FWD_CHUNKS = 40;     // kind 1: 36 full + 4 half chunks
struct NetDesc {
  int kind, in_pts, pe_slabs;
  int fwd_full;                        // full (32 KB) forward chunks = 32 + 2 * pe_slabs; 4 half chunks (views layer) follow
  int k5;                              // row length of pts_linears.5 = 256 + in_pts
  int w_pts[8], b_pts[8];
  int w_views, b_views, w_feat, b_feat, w_alpha, b_alpha, w_rgb, b_rgb, total;
  size_t fwd_bytes, packed_bytes;      // one part (hi or lo) of the packed image
};
static NetDesc make_desc(int kind) {
  NetDesc d{};
  d.kind = kind;
  d.in_pts = kind == 0 ? 63 : 84;
  d.pe_slabs = (d.in_pts + 63) / 64;
  d.fwd_full = 32 + 2 * d.pe_slabs;
  d.k5 = 256 + d.in_pts;
  int off = 0;
  for (int l = 0; l < 8; ++l) {
    const int K = l == 0 ? d.in_pts : (l == 5 ? d.k5 : 256);
    d.w_pts[l] = off; off += 256 * K;
    d.b_pts[l] = off; off += 256;
  }
  d.w_views = off; off += 128 * 283;
  d.b_views = off; off += 128;
  d.w_feat = off; off += 256 * 256;
  d.b_feat = off; off += 256;
  d.w_alpha = off; off += 256;
  d.b_alpha = off; off += 1;
  d.w_rgb = off; off += 3 * 128;
  d.b_rgb = off; off += 3;
  d.total = off;
  d.fwd_bytes = (size_t)d.fwd_full * 32768 + 4 * 16384;
  d.packed_bytes = d.fwd_bytes + DG_BYTES;
  return d;
}

// stash (bf16 mode): [viewbias B*128 fp32 (1 KB aligned)] [acts: tiles x 10 x 64 KB] [masks: tiles x 9 x 4 x 128 x 8 B]
constexpr size_t TILE_ACT_BYTES = 10 * 65536;
constexpr size_t TILE_MASK_BYTES = 9 * 128 * 32;

struct ChunkSrc {
  int base;     // float offset of element (row 0, col 0) in the flat parameter buffer
  int row_mul;  // float stride per chunk row
  int col_mul;  // float stride per chunk column
  int kvalid;   // columns >= kvalid are zero
  int nrows;    // 256 or 128
  int byte_off; // byte offset of the chunk in the packed image
};
__constant__ ChunkSrc c_chunks[kNetKinds][MAX_FWD_CHUNKS + DG_CHUNKS];
__constant__ NetDesc c_desc[kNetKinds];   // the kernels take the network KIND as a parameter and read its layout from here

// one thread = one 16-byte chunk (8 bf16) of the packed image; blockIdx.z = part: 0 = hi = bf16(w), 1 = lo =
// bf16(w - hi) (the split-precision mode multiplies by hi + lo; the bf16 mode reads part 0 only)
__global__ void pack_weights_kernel(const float *__restrict__ P, uint8_t *__restrict__ packed, int kind, size_t part_bytes) {
  int ci = blockIdx.y;
  const bool lo = blockIdx.z != 0;
  packed += (size_t)blockIdx.z * part_bytes;
  ChunkSrc s = c_chunks[kind][ci];
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= s.nrows * 8) return;
  int r = idx >> 3, q = idx & 7;
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int c0 = q * 8 + 2 * e;
    float a = (c0 < s.kvalid) ? P[s.base + r * s.row_mul + c0 * s.col_mul] : 0.f;
    float b = (c0 + 1 < s.kvalid) ? P[s.base + r * s.row_mul + (c0 + 1) * s.col_mul] : 0.f;
    w[e] = pack_bf16(a, b);
    if (lo) w[e] = pack_bf16(a - bf16_lo(w[e]), b - bf16_hi(w[e]));
  }
  *reinterpret_cast<uint4 *>(packed + s.byte_off + sw128_offset(r, q * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
}

// viewbias[ray][c] = b_v[c] + sum_j W_v[c][256+j] * PE4(viewdir)[j]  -- the 27 view-direction inputs of
// views_linears.0 are constant along a ray, so their contribution is a per-ray bias (fp32, exact)
__global__ void __launch_bounds__(256) viewbias_kernel(int64_t B, const float *__restrict__ P, const float *__restrict__ dirpe,
                                                       float *__restrict__ vb, int W_VIEWS, int B_VIEWS) {
  // block = 32 rays x 128 output channels.  The 128 x 27 view-direction weights (+ bias) and the block's 32 x 27 PE values are
  // staged in shared memory once; thread = (channel c, ray group): 16 rays each, weights in registers, PE values as warp-uniform
  // shared-memory broadcasts, coalesced 128-float stores.  (The first version looped 16 rays per thread over __ldg loads with
  // 256 blocks of 128 threads: latency-bound at 21 us for 4096 rays -- as long as the whole positional encoding.)
  __shared__ float s_w[128 * 28];
  __shared__ float s_pe[32 * 28];
  const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.x; i < 128 * 28; i += 256) {
    const int cc = i / 28, j = i % 28;
    s_w[i] = j < 27 ? P[W_VIEWS + cc * 283 + 256 + j] : P[B_VIEWS + cc];
  }
  for (int i = threadIdx.x; i < 32 * 28; i += 256) {
    const int rr = i / 28, j = i % 28;
    s_pe[i] = (j < 27 && r0 + rr < B) ? dirpe[(r0 + rr) * 32 + j] : 0.f;
  }
  __syncthreads();
  float w[28];
#pragma unroll
  for (int j = 0; j < 28; ++j) w[j] = s_w[c * 28 + j];
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const int rr = half * 16 + k;
    if (r0 + rr >= B) break;
    float acc = w[27];
#pragma unroll
    for (int j = 0; j < 27; ++j) acc = fmaf(w[j], s_pe[rr * 28 + j], acc);
    vb[(r0 + rr) * 128 + c] = acc;
  }
}

// pair MMA over one 64-wide K chunk: A = this tile set's K-major SW128 slab [128 rows x 64] (same offset in both
// CTAs), B = this CTA's half [N/2 rows x 64] of the weight chunk; 4 k-steps of 16
__device__ __forceinline__ void issue_chunk(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, uint32_t idesc, bool first) {
  uint64_t da = make_smem_desc(a_smem, 0, 1024), db = make_smem_desc(b_smem, 0, 1024);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma2_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
}

// barrier / TMEM / constant-vector set-up shared by the two kernels; returns the TMEM base
template <class L>
__device__ __forceinline__ uint32_t pair_setup(uint8_t *smem, uint32_t bar, uint32_t cr, int warp, uint32_t lane,
                                               const float *P, int kind) {
  const NetDesc &nd = c_desc[kind];
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::OFF_BAR + 128);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  // biases of the 9 tensor layers that have one (pts 0..7, feature) and the small heads: resident for the whole
  // launch, read as warp-uniform LDS (a global __ldg here costs an exposed L1/L2 latency per 4 columns)
  if (L::OFF_HEAD != L::OFF_VEC) {
    float *bias = reinterpret_cast<float *>(smem + L::OFF_VEC);
    for (int i = threadIdx.x; i < (int)BIAS_FLOATS; i += blockDim.x) {
      const int l = i >> 8, c = i & 255;
      bias[i] = P[(l < 8 ? nd.b_pts[l] : nd.b_feat) + c];
    }
  }
  float *head = reinterpret_cast<float *>(smem + L::OFF_HEAD);
  for (int i = threadIdx.x; i < (int)HEAD_FLOATS; i += blockDim.x)
    head[i] = i < 387 ? P[nd.w_rgb + i] : (i == 387 ? P[nd.b_alpha] : P[nd.w_alpha + (i - 388)]);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < L::NST; ++i) {
      mbar_init(L::w_full(bar, i), cr == 0 ? 2 : 1);  // leader: own producer + the follower's relay
      mbar_init(L::w_empty(bar, i), 1);
    }
    for (int t = 0; t < 2; ++t) { mbar_init(L::acc_full(bar, t), 1); mbar_init(L::act_ready(bar, t), 2 * kEpiWarps); }
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs exist before any remote arrive / multicast commit
  if (warp == 0) { tmem_alloc2(smem_u32(tmem_slot), 512); tmem_relinquish2(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both halves of the pair allocation are in place before the leader's first MMA
  tc_fence_after();
  return *tmem_slot;
}

// epilogue helper: store 32 consecutive bf16 columns [c0, c0+32) of row r (packed as 16 u32) into an act tile
__device__ __forceinline__ void store_cols32(uint8_t *act_tile, uint32_t r, uint32_t c0, const uint32_t pk[16]) {
  uint8_t *slab = act_tile + (c0 >> 6) * SLAB_BYTES + (r >> 3) * 1024u + (r & 7u) * 128u;
  uint32_t q0 = (c0 & 63u) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t pos = ((q0 + q) ^ (r & 7u)) << 4;
    *reinterpret_cast<uint4 *>(slab + pos) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}

// A warp owns rows [32*quarter, +32) x column half ch of its tile = a contiguous 4 KB piece of each of its slabs (in
// shared memory and in the HBM image alike), so it ships its own pieces: no CTA-level barrier around the stash stores.
__device__ __forceinline__ void warp_store_slabs(uint8_t *dst_tile_layer, const uint8_t *act_tile, uint32_t quarter,
                                                 uint32_t slab0, uint32_t nslabs) {
  for (uint32_t s = 0; s < nslabs; ++s) {
    const uint32_t off = (slab0 + s) * SLAB_BYTES + quarter * 4096u;
    bulk_s2g(dst_tile_layer + off, smem_u32(act_tile + off), 4096u);
  }
  bulk_commit();
}

// ---- epilogue building blocks ---------------------------------------------------------------------
// forward: 32 accumulator columns [c0, c0+32) of row r -> +bias -> (relu) -> bf16 -> act tile; kType 0 relu,
// 1 relu + alpha head, 2 linear.  Returns the ReLU mask word of the 32 outputs (bit 31-i set = output i inactive,
// i.e. pre-activation negative; 0 when not needed).
template <int kType, bool kMask>
__device__ __forceinline__ uint32_t fwd_block(const uint32_t v[32], const float *s_b, uint8_t *act_tile, uint32_t r,
                                              uint32_t c0, const float *s_wa, float &alpha) {
  uint32_t pk[16], m = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4 *>(s_b + c0 + 4 * q);
    // packed fp32x2 adds (FADD2): half the issue slots of 4 scalar adds
    const float2 x01 = __fadd2_rn(make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])), make_float2(b.x, b.y));
    const float2 x23 = __fadd2_rn(make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), make_float2(b.z, b.w));
    const uint32_t w0 = kType == 2 ? pack_bf16_fast(x01.x, x01.y) : pack_bf16_relu(x01.x, x01.y);
    const uint32_t w1 = kType == 2 ? pack_bf16_fast(x23.x, x23.y) : pack_bf16_relu(x23.x, x23.y);
    pk[2 * q] = w0;
    pk[2 * q + 1] = w1;
    if (kMask) {  // one funnel shift per output collects the SIGN bits: element i ends up at bit 31 - i, 1 = inactive
      m = __funnelshift_l(__float_as_uint(x01.x), m, 1);
      m = __funnelshift_l(__float_as_uint(x01.y), m, 1);
      m = __funnelshift_l(__float_as_uint(x23.x), m, 1);
      m = __funnelshift_l(__float_as_uint(x23.y), m, 1);
    }
    if (kType == 1) {  // alpha_linear on the bf16-rounded activations the next layers also see
      const float4 a = *reinterpret_cast<const float4 *>(s_wa + c0 + 4 * q);
      alpha = fmaf(bf16_lo(w0), a.x, alpha);
      alpha = fmaf(bf16_hi(w0), a.y, alpha);
      alpha = fmaf(bf16_lo(w1), a.z, alpha);
      alpha = fmaf(bf16_hi(w1), a.w, alpha);
    }
  }
  store_cols32(act_tile, r, c0, pk);
  return m;
}

// 64 columns [cq*64, cq*64+64) of one row: every epilogue warp owns one lane quarter x one column quarter of BOTH
// tiles of its CTA.  One 32-column TMEM load at a time: a tcgen05.ld costs ~40 cycles (tools/ub/tmem_bw.cu), while a
// second register block in flight costs spills at the 96-register cap (5 warps per scheduler) -- measured: -11 % dgrad.
template <int kType, bool kMask>
__device__ __forceinline__ uint2 fwd_epilogue_q(uint32_t tmem_rc, uint32_t cq, const float *s_bias, uint8_t *act_tile,
                                                uint32_t r, const float *s_wa, float &alpha) {
  uint32_t va[32];
  const uint32_t c0 = cq * 64;
  uint2 mk;
  tmem_ld32(tmem_rc, va);
  tmem_ld_wait(va);
  mk.x = fwd_block<kType, kMask>(va, s_bias, act_tile, r, c0, s_wa, alpha);
  tmem_ld32(tmem_rc + 32, va);
  tmem_ld_wait(va);
  mk.y = fwd_block<kType, kMask>(va, s_bias, act_tile, r, c0 + 32, s_wa, alpha);
  return mk;
}

// views_linears.0 (N=128: 64 columns per warp) + rgb_linear on CUDA cores -> raw[row].rgb (accumulated by both halves)
template <class Params>
__device__ __forceinline__ uint2 fwd_views_rgb(const Params &p, uint32_t tmem_row, uint32_t ch, uint8_t *act_tile, uint32_t r,
                                               int64_t row, bool live, const float *s_head) {
  const int64_t ray = (row < p.n ? row : p.n - 1) / p.S;
  const float4 *vb4 = reinterpret_cast<const float4 *>(p.viewbias + ray * 128);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f;
  uint32_t mk2[2];
  uint32_t va[32], vb[32];
  tmem_ld32(tmem_row + ch * 64, va);
  tmem_ld32(tmem_row + ch * 64 + 32, vb);
  tmem_ld_wait2(va, vb);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t cb = ch * 2 + i;
    const uint32_t *v = i == 0 ? va : vb;
    uint32_t pk[16], m = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = __ldg(vb4 + cb * 8 + q);
      const float x0 = __uint_as_float(v[4 * q]) + b.x, x1 = __uint_as_float(v[4 * q + 1]) + b.y;
      const float x2 = __uint_as_float(v[4 * q + 2]) + b.z, x3 = __uint_as_float(v[4 * q + 3]) + b.w;
      const uint32_t w0 = pack_bf16_relu(x0, x1);
      const uint32_t w1 = pack_bf16_relu(x2, x3);
      pk[2 * q] = w0;
      pk[2 * q + 1] = w1;
      m = __funnelshift_l(__float_as_uint(x0), m, 1);
      m = __funnelshift_l(__float_as_uint(x1), m, 1);
      m = __funnelshift_l(__float_as_uint(x2), m, 1);
      m = __funnelshift_l(__float_as_uint(x3), m, 1);
      const float h[4] = {bf16_lo(w0), bf16_hi(w0), bf16_lo(w1), bf16_hi(w1)};
      const int k = cb * 32 + 4 * q;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        c0 = fmaf(h[x], s_head[k + x], c0);
        c1 = fmaf(h[x], s_head[128 + k + x], c1);
        c2 = fmaf(h[x], s_head[256 + k + x], c2);
      }
    }
    mk2[i] = m;
    store_cols32(act_tile, r, cb * 32, pk);
  }
  if (live && row < p.n) {
    const float bsel = ch == 0 ? 1.f : 0.f;
    atomicAdd(p.raw + row * 4 + 0, c0 + bsel * s_head[384]);
    atomicAdd(p.raw + row * 4 + 1, c1 + bsel * s_head[385]);
    atomicAdd(p.raw + row * 4 + 2, c2 + bsel * s_head[386]);
  }
  return make_uint2(mk2[0], mk2[1]);
}

// backward: gradient columns [c0, c0+32) -> (+ d_sigma * w_alpha) -> relu mask -> bf16 -> act tile
template <bool kAlpha, bool kUseMask>
__device__ __forceinline__ void dgrad_block(const uint32_t v[32], uint32_t m, float dsig, const float *s_wa,
                                            uint8_t *act_tile, uint32_t r, uint32_t c0) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float g0 = __uint_as_float(v[i]), g1 = __uint_as_float(v[i + 1]);
    if (kAlpha) {
      g0 = fmaf(dsig, s_wa[c0 + i], g0);
      g1 = fmaf(dsig, s_wa[c0 + i + 1], g1);
    }
    if (kUseMask) {  // mask bit 31 - i set = output i of the forward layer was inactive
      g0 = (m & (0x80000000u >> i)) ? 0.f : g0;
      g1 = (m & (0x40000000u >> i)) ? 0.f : g1;
    }
    pk[i >> 1] = pack_bf16_fast(g0, g1);
  }
  store_cols32(act_tile, r, c0, pk);
}

// mk: this row's 64 ReLU mask bits of its column quarter (prefetched by the caller before the accumulator wait)
template <bool kAlpha, bool kUseMask>
__device__ __forceinline__ void dgrad_epilogue_q(uint32_t tmem_rc, uint32_t cq, const uint2 mk, float dsig, const float *s_wa,
                                                 uint8_t *act_tile, uint32_t r) {
  uint32_t va[32];
  const uint32_t c0 = cq * 64;
  tmem_ld32(tmem_rc, va);
  tmem_ld_wait(va);
  dgrad_block<kAlpha, kUseMask>(va, mk.x, dsig, s_wa, act_tile, r, c0);
  tmem_ld32(tmem_rc + 32, va);
  tmem_ld_wait(va);
  dgrad_block<kAlpha, kUseMask>(va, mk.y, dsig, s_wa, act_tile, r, c0 + 32);
}

// backward stage -1: G9 = (d_rgb * W_rgb) masked by relu(h9) -> act slabs 0,1 (64 columns per warp)
__device__ __forceinline__ void dgrad_g9(const float4 dr, const uint2 mk, uint32_t ch, const float *s_head,
                                         uint8_t *act_tile, uint32_t r) {
  const uint32_t mw[2] = {mk.x, mk.y};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint32_t cb = ch * 2 + j;
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const int k = cb * 32 + i;
      float g0 = dr.x * s_head[k] + dr.y * s_head[128 + k] + dr.z * s_head[256 + k];
      float g1 = dr.x * s_head[k + 1] + dr.y * s_head[128 + k + 1] + dr.z * s_head[256 + k + 1];
      g0 = (mw[j] & (0x80000000u >> i)) ? 0.f : g0;
      g1 = (mw[j] & (0x40000000u >> i)) ? 0.f : g1;
      pk[i >> 1] = pack_bf16_fast(g0, g1);
    }
    store_cols32(act_tile, r, cb * 32, pk);
  }
}

// ReLU bitmask stash: [tile][slot 0..8][column quarter][row][2 x u32] -- a warp's 32 rows are 256 contiguous bytes;
// word w covers columns 64 cq + 32 w + i at bit 31 - i, set = inactive (sign bit of the pre-activation).
// Slots 0..7 = H0..H7, slot 8 = h9 (views layer, 128 columns: column quarters 0,1 only).
__device__ __forceinline__ size_t mask_word_offset(int64_t tile, int slot, uint32_t cq, uint32_t r) {
  return ((((size_t)tile * 9 + (size_t)slot) * 4 + cq) * 128 + r) * 2;
}

// =================================================================================================
// forward
// =================================================================================================
struct FwdParams {
  const float *P;          // fp32 parameters (biases, alpha / rgb / view weights)
  const uint8_t *packed;   // bf16 chunks
  const uint8_t *pe_tiles; // [tiles][16 KB]
  const float *viewbias;   // [B][128]
  float *raw;              // [n][4], ZEROED by the host: the two column-half warps of a row accumulate into it
  uint8_t *stash_act;      // or null
  uint32_t *stash_mask;    // or null
  int64_t n;
  int S;
  int n_pairs;
  int dbg;   // profiling switches (bit 0: skip mask generation, bit 1: skip the activation bulk stores)
  long long *prof;  // per-CTA cycle accounting of the roles (PROF_SLOTS per CTA) when FLNERF_TC_PROF is set, else null
  int kind;         // network kind (index into c_desc)
  int stash_lo;     // split-precision kernels: 1 = the stash keeps the lo images too ([hi 64 KB][lo 64 KB] per slot), 0 = hi only
};
constexpr int PROF_SLOTS = 24;

// ring items of one iteration, in producer order (each = this CTA's half chunk or its own PE slab, <= 16 KB):
//   L0     : PE_A, W0, PE_B                            (W0 feeds both tile sets)
//   L1..L4 : 4 chunks each
//   L5     : W1..W4, Wpe, PE_A, W1..W4, Wpe, PE_B      (PE + 5 chunks do not fit the five stages twice over: layer 5
//            is fetched once per tile set and released chunk by chunk, the PE chunk last)
//   L6..L9 : 4 chunks each
constexpr int FWD_ITEMS = 3 + 16 + 12 + 16;
constexpr int DG_ITEMS = 34;

template <bool kProf>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mlp_fwd_tc(FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t bar = smem_u32(smem + LayF::OFF_BAR);
  const uint32_t cr = cluster_ctarank();
  const uint32_t tmem_base = pair_setup<LayF>(smem, bar, cr, warp, lane, p.P, p.kind);
  const uint32_t s_act = smem_u32(smem + OFF_ACT), s_w = smem_u32(smem + OFF_W);
  const int iters = (p.n_pairs + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool prof_on = kProf && p.prof != nullptr;  // the accounting is compiled out of the production instantiation

  if (warp == 0) {
    // ---------------------------------------------------------------- producer (both CTAs): local halves
    if (lane == 0) {
      LayF::Ring ring;
      long long pw = 0;
      const long long pt0 = prof_on ? clock64() : 0;
      auto push = [&](const uint8_t *src, uint32_t bytes) {
        const long long c0 = prof_on ? clock64() : 0;
        mbar_wait(LayF::w_empty(bar, ring.stage), ring.phase ^ 1);
        if (prof_on) pw += clock64() - c0;
        mbar_arrive_expect_tx(LayF::w_full(bar, ring.stage), bytes);
        bulk_g2s(s_w + ring.stage * WSTAGE, src, bytes, LayF::w_full(bar, ring.stage));
        ring.next();
      };
      auto push_w = [&](int cc) {  // this CTA's half (N/2 rows) of forward chunk cc
        const uint32_t half = cc < 34 ? 16384u : 8192u;
        const size_t off = cc < 34 ? (size_t)cc * 32768 : (size_t)34 * 32768 + (size_t)(cc - 34) * 16384;
        push(p.packed + off + (size_t)cr * half, half);
      };
      for (int it = 0; it < iters; ++it) {
        const int pair = min(p.n_pairs - 1, (int)blockIdx.x + it * (int)gridDim.x);
        const uint8_t *pe = p.pe_tiles + (size_t)pair * 2 * PE_BYTES;
        int ci = 0;
        for (int L = 0; L < 10; ++L) {
          if (L == 0) {
            push(pe, PE_BYTES); push_w(0); push(pe + PE_BYTES, PE_BYTES);
            ci = 1;
          } else if (L == 5) {
            for (int t = 0; t < 2; ++t) {
              for (int c = 1; c < 5; ++c) push_w(ci + c);
              push_w(ci);
              push(pe + (size_t)t * PE_BYTES, PE_BYTES);
            }
            ci += 5;
          } else {
            for (int c = 0; c < 4; ++c) push_w(ci + c);
            ci += 4;
          }
        }
      }
      if (prof_on) { p.prof[blockIdx.x * PROF_SLOTS + 4] = pw; p.prof[blockIdx.x * PROF_SLOTS + 5] = clock64() - pt0; }
    }
  } else if (warp == 1) {
    if (lane == 0 && cr != 0) {
      // -------------------------------------------------------------- relay (follower): local full -> leader's full
      LayF::Ring ring;
      const uint32_t remote_full = mapa_cluster(LayF::w_full(bar, 0), 0);
      const int total = iters * FWD_ITEMS;
      for (int q = 0; q < total; ++q) {
        mbar_wait(LayF::w_full(bar, ring.stage), ring.phase);
        mbar_arrive_cluster(remote_full + 8u * ring.stage);
        ring.next();
      }
    } else if (lane == 0) {
      // -------------------------------------------------------------- pair MMA issuer (leader)
      uint32_t n_act[2] = {0u, 0u};
      const uint32_t idesc256 = make_idesc(256, 256, 0, 0), idesc128 = make_idesc(256, 128, 0, 0);
      long long wa0 = 0, wa1 = 0, ww = 0, ww0 = 0, ww5 = 0;
      const long long mt0 = prof_on ? clock64() : 0;
      uint32_t q0 = 0;  // ring index of the current layer's first item
      auto wait_full = [&](uint32_t q) { mbar_wait(LayF::w_full(bar, LayF::item_stage(q)), LayF::item_phase(q)); };
      auto release = [&](uint32_t q) { umma2_commit_multicast(LayF::w_empty(bar, LayF::item_stage(q)), (uint16_t)3); };
      auto st = [&](uint32_t q) { return s_w + LayF::item_stage(q) * WSTAGE; };
      for (int it = 0; it < iters; ++it) {
        for (int L = 0; L < 10; ++L) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (!(it == 0 && L == 0)) {  // both CTAs' epilogue warps of tile set t: inputs written, accumulator drained
              const long long c0 = prof_on ? clock64() : 0;
              mbar_wait(LayF::act_ready(bar, t), n_act[t] & 1);
              if (prof_on) { if (t == 0) wa0 += clock64() - c0; else wa1 += clock64() - c0; }
              ++n_act[t];
            }
            tc_fence_after();
            const uint32_t d = tmem_base + t * 256, act = s_act + t * ACT_BYTES;
            const long long c1 = prof_on ? clock64() : 0;
            long long wsum = 0;
            if (L == 0) {
              const uint32_t pq = q0 + (t == 0 ? 0u : 2u), wq = q0 + 1;
              wait_full(pq);
              if (t == 0) wait_full(wq);
              if (prof_on) wsum = clock64() - c1;
              tc_fence_after();
              issue_chunk(d, st(pq), st(wq), idesc256, true);
              release(pq);
              if (t == 1) release(wq);
            } else if (L == 5) {
              // the five-stage ring cannot hold PE + 5 chunks for two tile sets: layer 5 is fetched once per tile set
              // and released chunk by chunk, in the same order for both (identical fp32 accumulation order keeps a row's
              // result independent of its tile).  The short-lived PE item comes last: every refill then has four
              // chunk-times of cover.
              const uint32_t b = q0 + (t == 0 ? 0u : 6u);
              for (uint32_t c = 1; c < 5; ++c) {
                const long long c2 = prof_on ? clock64() : 0;
                wait_full(b + c - 1);
                if (prof_on) wsum += clock64() - c2;
                tc_fence_after();
                issue_chunk(d, act + (c - 1) * SLAB_BYTES, st(b + c - 1), idesc256, c == 1);
                release(b + c - 1);
              }
              const long long c2 = prof_on ? clock64() : 0;
              wait_full(b + 4); wait_full(b + 5);
              if (prof_on) wsum += clock64() - c2;
              tc_fence_after();
              issue_chunk(d, st(b + 5), st(b + 4), idesc256, false);
              release(b + 4); release(b + 5);
            } else {
              for (uint32_t c = 0; c < 4; ++c) {
                if (t == 0) {
                  const long long c2 = prof_on ? clock64() : 0;
                  wait_full(q0 + c);
                  if (prof_on) wsum += clock64() - c2;
                  tc_fence_after();
                }
                issue_chunk(d, act + c * SLAB_BYTES, st(q0 + c), L == 9 ? idesc128 : idesc256, c == 0);
                if (t == 1) release(q0 + c);  // both tile sets are done with this chunk
              }
            }
            if (prof_on) { if (L == 0) ww0 += wsum; else if (L == 5) ww5 += wsum; else ww += wsum; }
            umma2_commit_multicast(LayF::acc_full(bar, t), (uint16_t)3);
          }
          q0 += (L == 0) ? 3u : (L == 5 ? 12u : 4u);
        }
      }
      if (prof_on) {
        long long *o = p.prof + blockIdx.x * PROF_SLOTS;
        o[0] = clock64() - mt0; o[1] = wa0; o[2] = wa1; o[3] = ww + ww0 + ww5; o[16] = ww0; o[17] = ww5;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 16 warps (both CTAs)
    // Every warp owns one lane quarter x one 64-column quarter of BOTH tiles and serves them alternately: a tile's
    // accumulator is drained by all 16 warps at once (half the latency of 8 warps per tile working side by side), so
    // tile A's next MMAs start while the same warps drain tile B.
    const int e = warp - kEpiWarp0;
    const uint32_t cq = (uint32_t)e >> 2;    // column quarter: columns [64 cq, 64 cq + 64) = slab cq
    const uint32_t quarter = warp & 3;       // TMEM lane quarter this warp may access
    const uint32_t r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const float *s_bias = reinterpret_cast<const float *>(smem + LayF::OFF_VEC);
    const float *s_head = reinterpret_cast<const float *>(smem + LayF::OFF_HEAD);
    const float *s_wa = s_head + 388;
    float *s_alpha = reinterpret_cast<float *>(smem + LayF::OFF_ALPHA);  // [tile][cq][row] partial alpha sums (layer 7)
    uint32_t n_layer = 0;  // both tiles' accumulator barriers flip once per layer, in order
    bool store_pending = false;
    const bool prof = prof_on && e == 0 && lane == 0;
    long long e_acc = 0, e_st = 0, e_body = 0, e_tail = 0;
    const long long et0 = prof ? clock64() : 0;
    for (int it = 0; it < iters; ++it) {
      const int pair_raw = (int)blockIdx.x + it * (int)gridDim.x;
      const bool live = pair_raw < p.n_pairs;          // surplus iteration: compute, but write nothing
      const int pair = live ? pair_raw : p.n_pairs - 1;
      uint8_t *stash_act = (live && !(p.dbg & 2)) ? p.stash_act : nullptr;
      uint32_t *stash_mask = (live && !(p.dbg & 1)) ? p.stash_mask : nullptr;
      for (int L = 0; L < 10; ++L, ++n_layer) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int64_t tile = (int64_t)pair * 2 + t;
          const int64_t row = tile * 128 + r;
          uint8_t *act_tile = smem + OFF_ACT + t * ACT_BYTES;
          const uint32_t tmem_rc = tmem_base + ((quarter * 32) << 16) + t * 256 + cq * 64;
          const bool has_cols = L < 9 || cq < 2;   // the views layer has 128 outputs: column quarters 0,1 only
          const long long c0 = prof ? clock64() : 0;
          mbar_wait(LayF::acc_full(bar, t), n_layer & 1);
          tc_fence_after();
          const long long c1 = prof ? clock64() : 0;
          if (p.stash_act) {  // my previous store of THIS tile's piece (two groups ago) must have finished reading it
            if (lane == 0 && store_pending) bulk_wait_read1();
            __syncwarp();
          }
          const long long c2 = prof ? clock64() : 0;
          float alpha = 0.f;
          uint2 mk = make_uint2(0u, 0u);
          const bool want_mask = stash_mask != nullptr;
          if (L < 7) {
            if (want_mask) mk = fwd_epilogue_q<0, true>(tmem_rc, cq, s_bias + L * 256, act_tile, r, s_wa, alpha);
            else fwd_epilogue_q<0, false>(tmem_rc, cq, s_bias + L * 256, act_tile, r, s_wa, alpha);
          } else if (L == 7) {
            if (want_mask) mk = fwd_epilogue_q<1, true>(tmem_rc, cq, s_bias + 7 * 256, act_tile, r, s_wa, alpha);
            else fwd_epilogue_q<1, false>(tmem_rc, cq, s_bias + 7 * 256, act_tile, r, s_wa, alpha);
            // alpha_linear: the four column quarters of a row are summed in a fixed order (bit-reproducible)
            s_alpha[(t * 4 + cq) * 128 + r] = alpha;
            named_bar_sync(1 + quarter, 128);
            if (cq == 0 && live && row < p.n) {
              const float *pa = s_alpha + t * 512 + r;
              p.raw[row * 4 + 3] = ((pa[0] + pa[128]) + (pa[256] + pa[384])) + s_head[387];
            }
          } else if (L == 8) {
            fwd_epilogue_q<2, false>(tmem_rc, cq, s_bias + 8 * 256, act_tile, r, s_wa, alpha);
          } else if (has_cols) {
            mk = fwd_views_rgb(p, tmem_rc - cq * 64, cq, act_tile, r, row, live, s_head);  // adds its own column offset
          }
          tc_fence_before();
          fence_async_smem();
          __syncwarp();
          const long long c3 = prof ? clock64() : 0;
          if (lane == 0) {
            // arrive FIRST: the issuer waits for it, while the stash store only has to leave before this warp rewrites
            // its piece one layer later (bulk_wait_read1 above); both only read the tile
            mbar_arrive_cluster(mapa_cluster(LayF::act_ready(bar, t), 0));  // the leader's barrier
            if (stash_act && has_cols) {
              warp_store_slabs(stash_act + (size_t)tile * TILE_ACT_BYTES + (size_t)L * 65536, act_tile, quarter, cq, 1);
              store_pending = true;
            }
          }
          // the mask words leave AFTER the arrive: its release would otherwise wait for these global stores
          if (want_mask && L != 8 && has_cols)
            *reinterpret_cast<uint2 *>(stash_mask + mask_word_offset(tile, L < 9 ? L : 8, cq, r)) = mk;
          if (prof) { e_acc += c1 - c0; e_st += c2 - c1; e_body += c3 - c2; e_tail += clock64() - c3; }
        }
      }
    }
    if (prof) {
      long long *o = p.prof + blockIdx.x * PROF_SLOTS + 6;
      o[0] = clock64() - et0; o[1] = e_acc; o[2] = e_st; o[3] = e_body; o[4] = e_tail;
      for (int k = 5; k < 10; ++k) o[k] = 0;
    }
    if (lane == 0 && store_pending) bulk_wait_all0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair may still be working for it
  if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

// =================================================================================================
// backward, data gradient chain:  G9 -> dF -> dH7 -> ... -> dH0   (pre-activation gradients, bf16)
// =================================================================================================
struct DgradParams {
  const float *P;
  const uint8_t *packed_dg;  // 34 chunks of 32 KB
  const float *draw;         // [n][4]
  const uint32_t *stash_mask;
  uint8_t *dy;               // [tiles][10][64 KB]: slot l<8 = dH_l, slot 8 = dF, slot 9 = G9 (32 KB)
  int64_t n;
  int n_pairs;
  long long *prof;
  int kind;
  int stash_lo;      // as FwdParams::stash_lo, for the gradient slots
};

template <bool kProf>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mlp_dgrad_tc(DgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t bar = smem_u32(smem + LayD::OFF_BAR);
  const uint32_t cr = cluster_ctarank();
  const uint32_t tmem_base = pair_setup<LayD>(smem, bar, cr, warp, lane, p.P, p.kind);
  const uint32_t s_act = smem_u32(smem + OFF_ACT), s_w = smem_u32(smem + OFF_W);
  const int iters = (p.n_pairs + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool prof_on = kProf && p.prof != nullptr;  // the accounting is compiled out of the production instantiation

  if (warp == 0) {
    if (lane == 0) {
      LayD::Ring ring;
      long long pw = 0;
      const long long pt0 = prof_on ? clock64() : 0;
      for (int it = 0; it < iters; ++it)
        for (int ci = 0; ci < DG_CHUNKS; ++ci) {
          const long long c0 = prof_on ? clock64() : 0;
          mbar_wait(LayD::w_empty(bar, ring.stage), ring.phase ^ 1);
          if (prof_on) pw += clock64() - c0;
          mbar_arrive_expect_tx(LayD::w_full(bar, ring.stage), 16384u);
          bulk_g2s(s_w + ring.stage * WSTAGE, p.packed_dg + (size_t)ci * 32768 + (size_t)cr * 16384, 16384u,
                   LayD::w_full(bar, ring.stage));
          ring.next();
        }
      if (prof_on) { p.prof[blockIdx.x * PROF_SLOTS + 4] = pw; p.prof[blockIdx.x * PROF_SLOTS + 5] = clock64() - pt0; }
    }
  } else if (warp == 1) {
    if (lane == 0 && cr != 0) {
      LayD::Ring ring;
      const uint32_t remote_full = mapa_cluster(LayD::w_full(bar, 0), 0);
      const int total = iters * DG_ITEMS;
      for (int q = 0; q < total; ++q) {
        mbar_wait(LayD::w_full(bar, ring.stage), ring.phase);
        mbar_arrive_cluster(remote_full + 8u * ring.stage);
        ring.next();
      }
    } else if (lane == 0) {
      uint32_t n_act[2] = {0u, 0u};
      const uint32_t idesc = make_idesc(256, 256, 0, 0);
      long long wa0 = 0, wa1 = 0, ww = 0;
      const long long mt0 = prof_on ? clock64() : 0;
      uint32_t q0 = 0;
      for (int it = 0; it < iters; ++it) {
        for (int D = 0; D < 9; ++D) {  // D=0: dF = G9 * Wv (K=128); D>=1: K=256
          const uint32_t nch = (D == 0) ? 2u : 4u;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const long long c0 = prof_on ? clock64() : 0;
            mbar_wait(LayD::act_ready(bar, t), n_act[t] & 1);
            if (prof_on) { if (t == 0) wa0 += clock64() - c0; else wa1 += clock64() - c0; }
            ++n_act[t];
            tc_fence_after();
            for (uint32_t c = 0; c < nch; ++c) {
              const uint32_t wq = q0 + c;
              if (t == 0) {
                const long long c1 = prof_on ? clock64() : 0;
                mbar_wait(LayD::w_full(bar, LayD::item_stage(wq)), LayD::item_phase(wq));
                if (prof_on) ww += clock64() - c1;
                tc_fence_after();
              }
              issue_chunk(tmem_base + t * 256, s_act + t * ACT_BYTES + c * SLAB_BYTES, s_w + LayD::item_stage(wq) * WSTAGE, idesc,
                          c == 0);
              if (t == 1) umma2_commit_multicast(LayD::w_empty(bar, LayD::item_stage(wq)), (uint16_t)3);
            }
            umma2_commit_multicast(LayD::acc_full(bar, t), (uint16_t)3);
          }
          q0 += nch;
        }
      }
      if (prof_on) {
        long long *o = p.prof + blockIdx.x * PROF_SLOTS;
        o[0] = clock64() - mt0; o[1] = wa0; o[2] = wa1; o[3] = ww; o[16] = 0; o[17] = 0;
      }
    }
  } else {
    // epilogue: 16 warps, each one lane quarter x one 64-column quarter of BOTH tiles (see mlp_fwd_tc)
    const int e = warp - kEpiWarp0;
    const uint32_t cq = (uint32_t)e >> 2;
    const uint32_t quarter = warp & 3;
    const uint32_t r = quarter * 32 + lane;
    const float *s_head = reinterpret_cast<const float *>(smem + LayD::OFF_HEAD);
    const float *s_wa = s_head + 388;
    uint32_t n_layer = 0;
    bool store_pending = false;
    const bool prof = prof_on && e == 0 && lane == 0;
    long long e_acc = 0, e_st = 0, e_body = 0, e_tail = 0;
    const long long et0 = prof ? clock64() : 0;
    for (int it = 0; it < iters; ++it) {
      const int pair_raw = (int)blockIdx.x + it * (int)gridDim.x;
      const bool live = pair_raw < p.n_pairs;
      const int pair = live ? pair_raw : p.n_pairs - 1;
      // stage -1: G9 from d_rgb; stages 0..8 = tensor layers.  ReLU mask of the activation the gradient flows into:
      // D=-1 -> h9 (slot 8), D=0 -> none (feature is linear), D=1 -> H7, D=2 -> H6, ..., D=8 -> H0
      for (int D = -1; D < 9; ++D) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int64_t tile = (int64_t)pair * 2 + t;
          uint8_t *act_tile = smem + OFF_ACT + t * ACT_BYTES;
          const uint32_t tmem_rc = tmem_base + ((quarter * 32) << 16) + t * 256 + cq * 64;
          const bool has_cols = D >= 0 || cq < 2;  // G9 has 128 columns: column quarters 0,1 only
          // the mask words (and, for the two stages that use it, the row's d_raw) come from HBM / L2: fetch them BEFORE
          // waiting for the accumulator; d_raw is not kept in registers across the stage loop
          float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
          if (D <= 1 && D != 0) {
            const int64_t row = tile * 128 + r;
            if (row < p.n) dr = __ldg(reinterpret_cast<const float4 *>(p.draw) + row);
          }
          uint2 mk = make_uint2(0u, 0u);
          if (D != 0 && has_cols)
            mk = __ldg(reinterpret_cast<const uint2 *>(p.stash_mask + mask_word_offset(tile, D < 0 ? 8 : 8 - D, cq, r)));
          const long long c0 = prof ? clock64() : 0;
          if (D >= 0) {
            mbar_wait(LayD::acc_full(bar, t), n_layer & 1);
            tc_fence_after();
          }
          const long long c1 = prof ? clock64() : 0;
          if (lane == 0 && store_pending) bulk_wait_read1();  // my store of THIS tile's piece is two groups back
          __syncwarp();
          const long long c2 = prof ? clock64() : 0;
          if (D < 0) {
            if (has_cols) dgrad_g9(dr, mk, cq, s_head, act_tile, r);
          } else if (D == 0) {
            dgrad_epilogue_q<false, false>(tmem_rc, cq, mk, 0.f, s_wa, act_tile, r);
          } else if (D == 1) {  // dH7 also receives d_sigma * w_alpha (alpha_linear reads H7)
            dgrad_epilogue_q<true, true>(tmem_rc, cq, mk, dr.w, s_wa, act_tile, r);
          } else {
            dgrad_epilogue_q<false, true>(tmem_rc, cq, mk, 0.f, s_wa, act_tile, r);
          }
          tc_fence_before();
          fence_async_smem();
          __syncwarp();
          const long long c3 = prof ? clock64() : 0;
          if (lane == 0) {
            if (D < 8) mbar_arrive_cluster(mapa_cluster(LayD::act_ready(bar, t), 0));  // the last stage feeds no further MMA
            if (live && has_cols) {
              const int slot = (D < 0) ? 9 : 8 - D;  // D=0 -> dF (8), D=1 -> dH7 (7) ... D=8 -> dH0 (0)
              warp_store_slabs(p.dy + (size_t)tile * TILE_ACT_BYTES + (size_t)slot * 65536, act_tile, quarter, cq, 1);
              store_pending = true;
            }
          }
          if (prof) { e_acc += c1 - c0; e_st += c2 - c1; e_body += c3 - c2; e_tail += clock64() - c3; }
        }
        if (D >= 0) ++n_layer;
      }
    }
    if (prof) {
      long long *o = p.prof + blockIdx.x * PROF_SLOTS + 6;
      o[0] = clock64() - et0; o[1] = e_acc; o[2] = e_st; o[3] = e_body; o[4] = e_tail;
      for (int k = 5; k < 10; ++k) o[k] = 0;
    }
    if (lane == 0 && store_pending) bulk_wait_all0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

#include "mlp_tc_x3.cuh"

// =================================================================================================
// backward, weight gradients:  dW[out][in] += sum_rows dY[row][out] * X[row][in]
// Both operands are the stashed [rows x features] SWIZZLE_128B images read as MN-major UMMA operands
// (LBO = distance between 64-feature slabs, SBO = 1024 = distance between 8-row groups).
// =================================================================================================
struct WUnit {
  int a_slot;   // dY slot (in dy buffer)
  int a_slabs;  // 4 (M=256) or 2 (M=128)
  int b_kind;   // 0 = stash act slot, 1 = PE tiles
  int b_slot;
  int b_slabs;  // 4 (N=256) or 1 (N=64)
  int w_off;    // float offset of dW(0,0)
  int ldw;
  int bias_off; // float offset of the bias gradient (column sums of dY), -1 = none
  int alpha;    // 1: also accumulate d w_alpha = sum_rows d_sigma[row] * X[row][:] and d b_alpha
                // 2: the views unit also carries the CUDA-core heads: d W_rgb, d b_rgb (needs the h9 half tile as a
                //    third operand) and the 27 view-direction columns of d W_views (per-ray sums of G9 x PE(dir))
  int splits;   // number of row ranges this unit is cut into
};
constexpr int kUnits = 11;
__constant__ WUnit c_units[kNetKinds][kUnits];

struct WgradParams {
  const uint8_t *dy;
  const uint8_t *stash_act;
  const uint8_t *pe_tiles;
  const float *draw;
  const float *dirpe;   // [B][32]
  float *G;
  long long *dbg;       // per-CTA {unit, cycles main loop, cycles total} when profiling the work split, else null
  int64_t n;
  int S;
  int n_tiles;  // 128-row tiles
  // operand addressing: a tile's slots are slot_stride apart, tiles tile_stride apart; a_part / b_part select the hi (0)
  // or lo (65536) image of the dY / activation slot, pe_part the hi (0) or lo PE tile set (split-precision mode; the
  // bf16 mode has slot_stride 65536 and all parts 0)
  size_t tile_stride, slot_stride, a_part, b_part, pe_part;
  // which CUDA-core reductions this pass carries (the split-precision mode runs three passes over the same rows):
  // bit 0: sums over dY (biases, view-direction columns), bit 1: sums over activations (w_alpha, W_rgb),
  // bit 2: sums over d_raw alone (b_alpha, b_rgb)
  int helper_flags;
  int kind;
};

constexpr int WG_STAGES = 3;
constexpr uint32_t WG_A_BYTES = 32768, WG_B_BYTES = 32768;  // 64 rows x 256 features each
constexpr uint32_t WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr uint32_t WG_OFF_BAR = WG_STAGES * WG_STAGE_BYTES;
constexpr uint32_t WG_OFF_DRAW = WG_OFF_BAR + 256;            // per stage: draw[64 rows][4] fp32 (units with heads)
constexpr uint32_t SMEM_WG = WG_OFF_DRAW + WG_STAGES * 1024;
static_assert(SMEM_WG <= 232448, "shared memory budget");

__device__ __forceinline__ float bf16_at(const uint8_t *p) {
  return __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t *>(p)) << 16);
}

constexpr int kWgThreads = 320;
__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_tc(WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t bar0 = smem_u32(smem + WG_OFF_BAR);
  // barriers: full[3], empty[3], acc_done
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + WG_OFF_BAR + 128);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  // helpers (8 warps) + MMA commit free a stage: 1 (commit) + 8 (warps)
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(bar0 + 8 * i, 1); mbar_init(bar0 + 8 * (WG_STAGES + i), 9); }
    mbar_init(bar0 + 8 * 2 * WG_STAGES, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // which (unit, row range) does this CTA own?
  int u = 0, part = blockIdx.x;
  while (u < kUnits - 1 && part >= c_units[p.kind][u].splits) { part -= c_units[p.kind][u].splits; ++u; }
  const WUnit un = c_units[p.kind][u];
  const int W_ALPHA = c_desc[p.kind].w_alpha, B_ALPHA = c_desc[p.kind].b_alpha, W_RGB = c_desc[p.kind].w_rgb,
            B_RGB = c_desc[p.kind].b_rgb, W_VIEWS = c_desc[p.kind].w_views;
  const int in_pts = c_desc[p.kind].in_pts, pe_slabs = c_desc[p.kind].pe_slabs;
  const int nhalf = p.n_tiles * 2;  // 64-row half tiles
  const int h_begin = (int)((int64_t)nhalf * part / un.splits), h_end = (int)((int64_t)nhalf * (part + 1) / un.splits);
  const uint32_t a_bytes = un.a_slabs * 8192u, b_bytes = un.b_slabs * 8192u;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int h = h_begin; h < h_end; ++h) {
        mbar_wait(bar0 + 8 * (WG_STAGES + stage), phase ^ 1);
        uint32_t full = bar0 + 8 * stage;
        const int64_t row0 = (int64_t)h * 64;
        const int64_t nv = p.n - row0;   // valid rows of this half tile (the tail of the last pair is padding)
        const uint32_t draw_bytes = (un.alpha && nv > 0) ? (uint32_t)(nv < 64 ? nv : 64) * 16u : 0u;
        mbar_arrive_expect_tx(full, a_bytes + b_bytes + (un.alpha == 2 ? 16384u : 0u) + draw_bytes);
        const size_t tile = (size_t)(h >> 1), half_off = (size_t)(h & 1) * 8192;
        const uint8_t *a_src = p.dy + tile * p.tile_stride + (size_t)un.a_slot * p.slot_stride + p.a_part + half_off;
        uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES), sb = sa + WG_A_BYTES;
        for (int s = 0; s < un.a_slabs; ++s) bulk_g2s(sa + s * 8192, a_src + (size_t)s * SLAB_BYTES, 8192u, full);
        const uint8_t *b_src = un.b_kind == 0 ? p.stash_act + tile * p.tile_stride + (size_t)un.b_slot * p.slot_stride + p.b_part + half_off
                                              : p.pe_tiles + p.pe_part + tile * (size_t)(pe_slabs * PE_BYTES) + half_off;
        for (int s = 0; s < un.b_slabs; ++s) bulk_g2s(sb + s * 8192, b_src + (size_t)s * SLAB_BYTES, 8192u, full);
        if (un.alpha == 2) {  // h9 (stash slot 9, 2 slabs) rides in the unused upper half of the A region
          const uint8_t *h9 = p.stash_act + tile * p.tile_stride + (size_t)9 * p.slot_stride + p.b_part + half_off;
          for (int s = 0; s < 2; ++s) bulk_g2s(sa + 16384 + s * 8192, h9 + (size_t)s * SLAB_BYTES, 8192u, full);
        }
        if (draw_bytes) bulk_g2s(smem_u32(smem + WG_OFF_DRAW + stage * 1024), p.draw + row0 * 4, draw_bytes, full);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t idesc = make_idesc(128, un.b_slabs * 64, 1, 1);
      const int mhalves = un.a_slabs / 2;
      for (int h = h_begin; h < h_end; ++h) {
        mbar_wait(bar0 + 8 * stage, phase);
        tc_fence_after();
        uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES), sb = sa + WG_A_BYTES;
        for (int mh = 0; mh < mhalves; ++mh) {
          // A: M = 128 out-features = 2 slabs; MN-major: LBO = 8192 (next 64 features), SBO = 1024 (next 8 rows)
          uint64_t da = make_smem_desc(sa + mh * 2 * 8192, 8192, 1024);
          uint64_t db = make_smem_desc(sb, 8192, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 rows per k-step = 2048 bytes
            umma_bf16(tmem_base + mh * 256, da + 128 * k, db + 128 * k, idesc, (h == h_begin && k == 0) ? 0u : 1u);
        }
        umma_commit(bar0 + 8 * (WG_STAGES + stage));
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(bar0 + 8 * 2 * WG_STAGES);
    }
  } else {
    // ------------------------------------------------------------------------------------------------------------
    // helper warps (256 threads): CUDA-core reductions straight from the staged tiles, then the accumulator flush.
    //   every unit : bias gradient = column sums of dY; thread = column PAIR (packed bf16x2), half of the 64 rows
    //   alpha == 1 : lower 128 threads do the bias sums (all rows), upper 128 threads accumulate
    //                d w_alpha[c] = sum_rows d_sigma[row] * H7[row][c] for a column pair (d_sigma from the staged draw rows)
    //   alpha == 2 : lower 128 threads: column sums of G9 per ray segment -> d b_views and the 27 view-direction
    //                columns of d W_views; upper 128 threads: d W_rgb[:, k] = sum_rows d_rgb[row] * h9[row][k]
    // ------------------------------------------------------------------------------------------------------------
    const int ht = threadIdx.x - kEpiWarp0 * 32;  // 0..255
    const bool upper = ht >= 128;
    const uint32_t hh = (uint32_t)ht & 127u;
    float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, gsum = 0.f, db0 = 0.f, db1 = 0.f, db2 = 0.f, db3 = 0.f;
    float av[27];
#pragma unroll
    for (int x = 0; x < 27; ++x) av[x] = 0.f;
    // byte offsets inside an 8-row group of a SWIZZLE_128B slab: column pair (2hh, 2hh+1) and single column hh
    uint32_t offp[8], offs[8];
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) {
      const uint32_t cp = (2u * hh) & 63u, c1 = hh & 63u;
      offp[j] = ((2u * hh) >> 6) * 8192u + j * 128u + ((((cp >> 3) ^ j) & 7u) << 4) + ((cp & 7u) << 1);
      offs[j] = (hh >> 6) * 8192u + j * 128u + ((((c1 >> 3) ^ j) & 7u) << 4) + ((c1 & 7u) << 1);
    }
    const bool pair_valid = 2 * (int)hh < un.a_slabs * 64;
    int rem = (int)(((int64_t)h_begin * 64) % p.S);  // position of the next row inside its ray (heads unit)
    const int hf = p.helper_flags;
    uint32_t stage = 0, phase = 0;
    const long long t_start = clock64();
    for (int h = h_begin; h < h_end; ++h) {
      mbar_wait(bar0 + 8 * stage, phase);
      const uint8_t *sa = smem + stage * WG_STAGE_BYTES;
      const uint8_t *sb = sa + WG_A_BYTES;
      const float4 *sd = reinterpret_cast<const float4 *>(smem + WG_OFF_DRAW + stage * 1024);
      const int64_t row0 = (int64_t)h * 64;
      const int nv = (int)((p.n - row0) < 64 ? (p.n - row0 < 0 ? 0 : p.n - row0) : 64);
      if (un.alpha == 0) {
        if (un.bias_off >= 0 && pair_valid && (hf & 1)) {  // rows [32*upper, 32*upper + 32)
#pragma unroll
          for (uint32_t g = 0; g < 4; ++g)
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
              const uint32_t w = *reinterpret_cast<const uint32_t *>(sa + ((upper ? 4u : 0u) + g) * 1024u + offp[j]);
              s0 += bf16_lo(w);
              s1 += bf16_hi(w);
            }
        }
      } else if (un.alpha == 1) {
        if (!upper) {
          if (hf & 1)
#pragma unroll
          for (uint32_t g = 0; g < 8; ++g)
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
              const uint32_t w = *reinterpret_cast<const uint32_t *>(sa + g * 1024u + offp[j]);
              s0 += bf16_lo(w);
              s1 += bf16_hi(w);
            }
        } else if (hf & 6) {
          const float xs = (hf & 2) ? 1.f : 0.f, dsel = (hf & 4) ? 1.f : 0.f;
#pragma unroll
          for (uint32_t g = 0; g < 8; ++g)
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
              const int rr = (int)(g * 8 + j);
              const float ds = rr < nv ? sd[rr].w : 0.f;
              const uint32_t w = *reinterpret_cast<const uint32_t *>(sb + g * 1024u + offp[j]);
              t0 = fmaf(ds * xs, bf16_lo(w), t0);
              t1 = fmaf(ds * xs, bf16_hi(w), t1);
              if (hh == 0) db3 += ds * dsel;
            }
        }
      } else {
        if (!upper) {
         if (hf & 1) {
          // the ray segment sum is carried across the consecutive half tiles of this CTA: it is folded with the ray's
          // 27 PE values only when the ray ends (every S rows) or at the CTA's last row
          const bool last_stage = (h == h_end - 1);
          auto fold = [&](int64_t row) {
            if (row < p.n) {
              const float *pe = p.dirpe + (row / p.S) * 32;
#pragma unroll
              for (int x = 0; x < 27; ++x) av[x] = fmaf(gsum, __ldg(pe + x), av[x]);
            }
            s0 += gsum;
            gsum = 0.f;
          };
          if (rem + 64 <= p.S) {  // no ray boundary strictly inside this half tile: straight 64-row sum
#pragma unroll
            for (uint32_t g = 0; g < 8; ++g)
#pragma unroll
              for (uint32_t j = 0; j < 8; ++j) gsum += bf16_at(sa + g * 1024u + offs[j]);
            rem += 64;
            if (rem == p.S || last_stage) {
              fold(row0 + 63);
              if (rem == p.S) rem = 0;
            }
          } else {  // general case (S not a multiple of 64): per-row boundary test, computed addresses
            const uint32_t c1 = hh & 63u, sl = (hh >> 6) * 8192u;
#pragma unroll 1
            for (uint32_t rr = 0; rr < 64; ++rr) {
              const uint32_t j = rr & 7u;
              gsum += bf16_at(sa + sl + (rr >> 3) * 1024u + j * 128u + ((((c1 >> 3) ^ j) & 7u) << 4) + ((c1 & 7u) << 1));
              if (++rem == p.S || (last_stage && rr == 63)) {
                fold(row0 + rr);
                if (rem == p.S) rem = 0;
              }
            }
          }
         }
        } else if (hf & 6) {
          const float xs = (hf & 2) ? 1.f : 0.f, dsel = (hf & 4) ? 1.f : 0.f;
#pragma unroll
          for (uint32_t g = 0; g < 8; ++g)
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
              const int rr = (int)(g * 8 + j);
              const float4 d = rr < nv ? sd[rr] : make_float4(0.f, 0.f, 0.f, 0.f);
              const float hv = bf16_at(sa + 16384 + g * 1024u + offs[j]) * xs;
              t0 = fmaf(d.x, hv, t0);
              t1 = fmaf(d.y, hv, t1);
              t2 = fmaf(d.z, hv, t2);
              if (hh == 0) { db0 += d.x * dsel; db1 += d.y * dsel; db2 += d.z * dsel; }
            }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (WG_STAGES + stage));
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
    const long long t_loop = clock64();
    if (h_end > h_begin) {
      if (un.alpha == 0) {
        if (un.bias_off >= 0 && pair_valid) {
          atomicAdd(p.G + un.bias_off + 2 * hh, s0);
          atomicAdd(p.G + un.bias_off + 2 * hh + 1, s1);
        }
      } else if (un.alpha == 1) {
        if (!upper) {
          atomicAdd(p.G + un.bias_off + 2 * hh, s0);
          atomicAdd(p.G + un.bias_off + 2 * hh + 1, s1);
        } else {
          atomicAdd(p.G + W_ALPHA + 2 * hh, t0);
          atomicAdd(p.G + W_ALPHA + 2 * hh + 1, t1);
          if (hh == 0) atomicAdd(p.G + B_ALPHA, db3);
        }
      } else {
        if (!upper) {
          atomicAdd(p.G + un.bias_off + hh, s0);
#pragma unroll
          for (int x = 0; x < 27; ++x) atomicAdd(p.G + W_VIEWS + hh * 283 + 256 + x, av[x]);
        } else {
          atomicAdd(p.G + W_RGB + hh, t0);
          atomicAdd(p.G + W_RGB + 128 + hh, t1);
          atomicAdd(p.G + W_RGB + 256 + hh, t2);
          if (hh == 0) {
            atomicAdd(p.G + B_RGB, db0);
            atomicAdd(p.G + B_RGB + 1, db1);
            atomicAdd(p.G + B_RGB + 2, db2);
          }
        }
      }
      // flush the accumulators: TMEM lane = out feature (within the 128-row M half), column = in feature.  A thread
      // owns a ROW, so direct atomics would touch 32 different lines per instruction; each warp transposes its
      // 32x32 block through the (now idle: all helpers are past their last stage) stage memory and issues
      // line-coalesced reductions instead.
      mbar_wait(bar0 + 8 * 2 * WG_STAGES, 0);
      tc_fence_after();
      named_bar_sync(1, 256);
      const int e = warp - kEpiWarp0;
      const uint32_t quarter = warp & 3;
      const int ncol = un.b_slabs * 64;
      const int mhalves = un.a_slabs / 2;
      float *tr = reinterpret_cast<float *>(smem) + e * (32 * 33);
      for (int mh = 0; mh < mhalves; ++mh) {
        const int out0 = mh * 128 + quarter * 32;
        // the two warps that share a lane quarter split the 32-column blocks between them
        for (int cb = (e >> 2); cb * 32 < ncol; cb += 2) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((quarter * 32) << 16) + mh * 256 + cb * 32, v);
          tmem_ld_wait(v);
#pragma unroll
          for (int i = 0; i < 32; ++i) tr[lane * 33 + i] = __uint_as_float(v[i]);
          __syncwarp();
          const int col = cb * 32 + lane;
          // PE units have in_pts (63 / 84) valid inputs: the zero-pad columns behind them must not be written
          const bool ok = !(un.b_kind == 1 && col >= in_pts);
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr)
            if (ok) atomicAdd(p.G + un.w_off + (size_t)(out0 + rr) * un.ldw + col, tr[rr * 33 + lane]);
          __syncwarp();
        }
      }
    }
    if (p.dbg && threadIdx.x == kEpiWarp0 * 32) {
      p.dbg[blockIdx.x * 4 + 0] = u;
      p.dbg[blockIdx.x * 4 + 1] = t_loop - t_start;
      p.dbg[blockIdx.x * 4 + 2] = clock64() - t_start;
      p.dbg[blockIdx.x * 4 + 3] = h_end - h_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// rgb head + view-direction columns of views_linears.0: CUDA cores (0.1% of the FLOPs)
//   dW_rgb[c][k] += sum_rows d_rgb[row][c] * h9[row][k];  db_rgb += sum d_rgb
//   dW_v[c][256+j] += sum_rows G9[row][c] * dirpe[ray(row)][j]
__global__ void __launch_bounds__(512) wgrad_small_kernel(const uint8_t *__restrict__ stash_act,
                                                          const uint8_t *__restrict__ dy,
                                                          const float *__restrict__ draw,
                                                          const float *__restrict__ dirpe, float *__restrict__ G,
                                                          int64_t n, int S, int n_tiles, int tiles_per_block) {
  __shared__ float red[33][128];  // partial sums of one row-quarter at a time
  const int k = threadIdx.x & 127;   // feature column 0..127
  const int qr = threadIdx.x >> 7;   // row quarter 0..3 (32 rows of every tile)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, gsum = 0.f;
  float av[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) av[j] = 0.f;
  const int t0 = blockIdx.x * tiles_per_block, t1 = min(n_tiles, t0 + tiles_per_block);
  const uint32_t slab = k >> 6, c = k & 63;
  auto flush_ray = [&](int64_t ray) {  // sum_s G9[row][k] over a ray segment, multiplied once with the ray's 27 PE values
    const float *pe = dirpe + ray * 32;
#pragma unroll
    for (int j = 0; j < 27; ++j) av[j] = fmaf(gsum, __ldg(pe + j), av[j]);
    gsum = 0.f;
  };
  for (int tile = t0; tile < t1; ++tile) {
    const uint8_t *h9 = stash_act + (size_t)tile * TILE_ACT_BYTES + (size_t)9 * 65536 + slab * SLAB_BYTES;
    const uint8_t *g9 = dy + (size_t)tile * TILE_ACT_BYTES + (size_t)9 * 65536 + slab * SLAB_BYTES;
    const int64_t row0 = (int64_t)tile * 128 + qr * 32;
    const int64_t seg_end = min(n, row0 + 32);
#pragma unroll
    for (uint32_t r8 = 0; r8 < 32; r8 += 8) {
      if (row0 + r8 >= seg_end) break;
      float hv[8], gv[8];
      float4 d[8];
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j) {  // 8 rows in flight
        const uint32_t off = ((uint32_t)qr * 4 + (r8 >> 3)) * 1024u + j * 128u + ((((c >> 3) ^ j) & 7u) << 4) + ((c & 7u) << 1);
        hv[j] = __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t *>(h9 + off)) << 16);
        gv[j] = __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t *>(g9 + off)) << 16);
        const int64_t row = row0 + r8 + j;
        d[j] = row < n ? __ldg(reinterpret_cast<const float4 *>(draw) + row) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j) {
        const int64_t row = row0 + r8 + j;
        if (row >= n) break;
        a0 = fmaf(d[j].x, hv[j], a0); a1 = fmaf(d[j].y, hv[j], a1); a2 = fmaf(d[j].z, hv[j], a2);
        if (k == (int)(r8 + j)) { b0 += d[j].x; b1 += d[j].y; b2 += d[j].z; }
        gsum += gv[j];
        if ((row + 1) % S == 0 || row + 1 == seg_end) flush_ray(row / S);
      }
    }
  }
  // reduce the four row-quarters (one at a time through shared memory), then one atomic per (column, output)
  for (int q = 1; q < 4; ++q) {
    if (qr == q) {
      float *dst = &red[0][k];
      dst[0] = a0; dst[128] = a1; dst[256] = a2; dst[3 * 128] = b0; dst[4 * 128] = b1; dst[5 * 128] = b2;
#pragma unroll
      for (int j = 0; j < 27; ++j) dst[(6 + j) * 128] = av[j];
    }
    __syncthreads();
    if (qr == 0) {
      const float *src = &red[0][k];
      a0 += src[0]; a1 += src[128]; a2 += src[256]; b0 += src[3 * 128]; b1 += src[4 * 128]; b2 += src[5 * 128];
#pragma unroll
      for (int j = 0; j < 27; ++j) av[j] += src[(6 + j) * 128];
    }
    __syncthreads();
  }
  if (qr == 0) {
    atomicAdd(G + W_RGB + k, a0);
    atomicAdd(G + W_RGB + 128 + k, a1);
    atomicAdd(G + W_RGB + 256 + k, a2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      b0 += __shfl_xor_sync(0xffffffffu, b0, o);
      b1 += __shfl_xor_sync(0xffffffffu, b1, o);
      b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    if ((k & 31) == 0) {
      atomicAdd(G + B_RGB, b0);
      atomicAdd(G + B_RGB + 1, b1);
      atomicAdd(G + B_RGB + 2, b2);
    }
#pragma unroll
    for (int j = 0; j < 27; ++j) atomicAdd(G + W_VIEWS + k * 283 + 256 + j, av[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// clusters of 2, one pair per CTA: even grid, at least 2 (a CTA without a pair of its own recomputes the last one)
static int pair_grid(int n_pairs, int sm_count) {
  int g = n_pairs < sm_count ? n_pairs : sm_count;
  g = (g + 1) & ~1;
  if (g > sm_count) g -= 2;
  return g < 2 ? 2 : g;
}
static bool g_tables_ready = false;
static int g_wgrad_grid = 0;

// FLNERF_TC_PROF=1: the forward / dgrad kernels count, per role, the cycles spent waiting on each barrier; the host
// prints the leader CTA 0 and the grid mean for every launch of >= 1024 pairs (development aid, off by default)
static long long *prof_buffer() {
  static long long *buf = nullptr;
  static int on = -1;
  if (on < 0) {
    on = getenv("FLNERF_TC_PROF") != nullptr;
    if (on && cudaMalloc(&buf, sizeof(long long) * PROF_SLOTS * 1024) != cudaSuccess) buf = nullptr;
  }
  return buf;
}
static void prof_report(const char *name, int grid, int n_pairs, cudaStream_t st) {
  if (n_pairs < 1024) return;
  static long long h[PROF_SLOTS * 1024];
  cudaStreamSynchronize(st);
  cudaMemcpy(h, prof_buffer(), sizeof(long long) * PROF_SLOTS * grid, cudaMemcpyDeviceToHost);
  static const char *lbl[18] = {"mma_total", "mma_wait_act0", "mma_wait_act1", "mma_wait_wfull", "prod_wait_empty", "prod_total",
                                "epi0_total", "epi0_wait_acc", "epi0_wait_store", "epi0_body", "epi0_tail",
                                "epi1_total", "epi1_wait_acc", "epi1_wait_store", "epi1_body", "epi1_tail",
                                "mma_wait_wfull_L0", "mma_wait_wfull_L5"};
  fprintf(stderr, "tcprof %s grid %d pairs %d:", name, grid, n_pairs);
  for (int k = 0; k < 18; ++k) {
    double m = 0;
    for (int i = 0; i < grid; ++i) m += (double)h[i * PROF_SLOTS + k];
    fprintf(stderr, " %s=%.0f(cta0 %lld)", lbl[k], m / grid, h[k]);
  }
  fprintf(stderr, "\n");
}

static NetDesc g_desc[kNetKinds];

static int setup_tables(int sm_count) {
  if (g_tables_ready) return 0;
  static ChunkSrc ch[kNetKinds][MAX_FWD_CHUNKS + DG_CHUNKS];
  static WUnit units[kNetKinds][kUnits];
  for (int kind = 0; kind < kNetKinds; ++kind) {
    const NetDesc d = make_desc(kind);
    g_desc[kind] = d;
    if (kind == 0 && (d.total != TOTAL || d.w_views != W_VIEWS || d.w_rgb != W_RGB || d.b_pts[5] != B_PTS[5] ||
                      d.fwd_bytes != FWD_BYTES)) return 1;     // the generic layout reproduces mlp_layout.h
    int ci = 0, off = 0;
    auto add = [&](int base, int row_mul, int col_mul, int kvalid, int nrows) {
      ChunkSrc &c = ch[kind][ci];
      c.base = base; c.row_mul = row_mul; c.col_mul = col_mul; c.kvalid = kvalid; c.nrows = nrows; c.byte_off = off;
      off += nrows * 128;
      ++ci;
    };
    const int P = d.pe_slabs;
    // forward: B operand rows = output feature, columns = 64 consecutive input features.  Chunk order: layer 0 = its P
    // positional slabs; layers 1..4; layer 5 = its P positional slabs (columns [0, in_pts) of the 256+in_pts wide row), then
    // its four activation slabs; layers 6, 7; feature_linear; views_linears.0 (128 rows: half-size chunks)
    for (int sl = 0; sl < P; ++sl) add(d.w_pts[0] + sl * 64, d.in_pts, 1, d.in_pts - sl * 64 < 64 ? d.in_pts - sl * 64 : 64, 256);
    for (int l = 1; l < 8; ++l) {
      if (l == 5) {
        for (int sl = 0; sl < P; ++sl) add(d.w_pts[5] + sl * 64, d.k5, 1, d.in_pts - sl * 64 < 64 ? d.in_pts - sl * 64 : 64, 256);
        for (int c = 0; c < 4; ++c) add(d.w_pts[5] + d.in_pts + c * 64, d.k5, 1, 64, 256);
      } else {
        for (int c = 0; c < 4; ++c) add(d.w_pts[l] + c * 64, 256, 1, 64, 256);
      }
    }
    for (int c = 0; c < 4; ++c) add(d.w_feat + c * 64, 256, 1, 64, 256);
    for (int c = 0; c < 4; ++c) add(d.w_views + c * 64, 283, 1, 64, 128);
    if (ci != d.fwd_full + 4 || (size_t)off != d.fwd_bytes) return 1;
    // dgrad: B operand rows = input feature j, columns = 64 consecutive output features: W[out][in_off + j]
    for (int c = 0; c < 2; ++c) add(d.w_views + c * 64 * 283, 1, 283, 64, 256);
    for (int c = 0; c < 4; ++c) add(d.w_feat + c * 64 * 256, 1, 256, 64, 256);
    for (int l = 7; l >= 1; --l) {
      const int ld = l == 5 ? d.k5 : 256, in_off = (l == 5) ? d.in_pts : 0;
      for (int c = 0; c < 4; ++c) add(d.w_pts[l] + in_off + c * 64 * ld, 1, ld, 64, 256);
    }
    if (ci != d.fwd_full + 4 + DG_CHUNKS || (size_t)off != d.packed_bytes) return 1;

    const WUnit un[kUnits] = {
        // a_slot a_slabs b_kind b_slot b_slabs w_off             ldw  bias_off   alpha splits
        {0, 4, 1, 0, P, d.w_pts[0], d.in_pts, d.b_pts[0], 0, 0},
        {1, 4, 0, 0, 4, d.w_pts[1], 256, d.b_pts[1], 0, 0},
        {2, 4, 0, 1, 4, d.w_pts[2], 256, d.b_pts[2], 0, 0},
        {3, 4, 0, 2, 4, d.w_pts[3], 256, d.b_pts[3], 0, 0},
        {4, 4, 0, 3, 4, d.w_pts[4], 256, d.b_pts[4], 0, 0},
        {5, 4, 0, 4, 4, d.w_pts[5] + d.in_pts, d.k5, d.b_pts[5], 0, 0},
        {5, 4, 1, 0, P, d.w_pts[5], d.k5, -1, 0, 0},
        {6, 4, 0, 5, 4, d.w_pts[6], 256, d.b_pts[6], 0, 0},
        {7, 4, 0, 6, 4, d.w_pts[7], 256, d.b_pts[7], 0, 0},
        {8, 4, 0, 7, 4, d.w_feat, 256, d.b_feat, 1, 0},
        {9, 2, 0, 8, 4, d.w_views, 283, d.b_views, 2, 0},
    };
    // split the row range of every unit in proportion to the bytes it streams, one work item per SM
    double cost[kUnits], total = 0;
    for (int i = 0; i < kUnits; ++i) {
      // relative time per half tile, measured per unit with FLNERF_WG_DEBUG (bytes streamed + CUDA-core helper work)
      cost[i] = un[i].b_kind == 1 ? (un[i].b_slabs == 1 ? 0.85 : 0.9) : (un[i].alpha ? 1.2 : 1.0);
      total += cost[i];
      units[kind][i] = un[i];
    }
    int used = 0;
    for (int i = 0; i < kUnits; ++i) {
      int sp = (int)(sm_count * cost[i] / total);
      if (sp < 1) sp = 1;
      units[kind][i].splits = sp;
      used += sp;
    }
    for (int i = 0; used < sm_count; i = (i + 1) % kUnits) {
      if (un[i].a_slabs == 4 && un[i].b_kind == 0 && un[i].b_slabs == 4) { ++units[kind][i].splits; ++used; }
    }
    g_wgrad_grid = used;       // = sm_count for both kinds
  }
  if (cudaMemcpyToSymbol(c_chunks, ch, sizeof(ch)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(c_units, units, sizeof(units)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(c_desc, g_desc, sizeof(g_desc)) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_dgrad_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayD::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_dgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayD::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_WG) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_gen<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_gen<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_gen<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_gen<3, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_gen<3, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_gen<3, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_fwd_gen<3, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF::SMEM) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(mlp_dgrad_x3, cudaFuncAttributeMaxDynamicSharedMemorySize, LayD::SMEM) != cudaSuccess) return 1;
  g_tables_ready = true;
  return 0;
}

}  // namespace tc

// ---- entry points used by api.cu -------------------------------------------------------------------
// kind: 0 = 63 position channels (nerf-ours NeRF / nerf++ foreground), 1 = 84 (nerf++ background)
// packed image of one net: [part 0 = hi: forward chunks | dgrad chunks][part 1 = lo: same layout]
size_t mlp_tc_packed_bytes(int kind) { return 2 * tc::make_desc(kind).packed_bytes; }
int64_t mlp_tc_param_count(int kind) { return tc::make_desc(kind).total; }
size_t mlp_tc_pe_tile_bytes(int kind, bool x3) { return (size_t)tc::make_desc(kind).pe_slabs * tc::PE_BYTES * (x3 ? 2 : 1); }

static size_t tc_vb_bytes(int64_t n, int S) {
  int64_t B = (n + S - 1) / S;
  return (size_t)((B * 128 * 4 + 1023) / 1024) * 1024;
}
// How many of the three terms dYhi^T Xhi + dYhi^T Xlo + dYlo^T Xhi the split-precision weight gradient carries.  The sum over
// the rows averages the bf16 rounding of its operands away: measured against the oracle (tools/x3_wgrad_passes.py,
// profiles/r02g_wgrad_passes.log) the gradient's rel-L2 at 393 216 rows (2048 rays x 192) is 3.3e-5 / 3.2e-5 / 3.0e-5 for
// 1 / 2 / 3 terms, while 16..24-ray batches (3..5 k rows) need all three to stay under 2e-3.  Default: ONE term from 65 536
// rows on (any training batch), three below; with one term the stash keeps only the hi images (half the bytes).
// $FLNERF_X3_WGRAD_PASSES = 1 | 2 | 3 forces a count.
static int x3_wgrad_passes(int64_t n) {
  static int forced = -1;
  if (forced < 0) {
    const char *e = getenv("FLNERF_X3_WGRAD_PASSES");
    forced = e ? atoi(e) : 0;
    if (forced < 0 || forced > 3) forced = 0;
  }
  return forced ? forced : (n >= 65536 ? 1 : 3);
}
static bool x3_stash_lo(bool x3, int64_t n) { return x3 && x3_wgrad_passes(n) > 1; }
static size_t tc_tile_act_bytes(bool x3, int64_t n) { return x3_stash_lo(x3, n) ? tc::TILE_ACT_BYTES_X3 : tc::TILE_ACT_BYTES; }
size_t mlp_tc_stash_bytes(int64_t n, int S, int training, bool x3) {
  size_t tiles = (size_t)(flnerf_padded_rows(n) / 128);
  return tc_vb_bytes(n, S) + (training ? tiles * (tc_tile_act_bytes(x3, n) + tc::TILE_MASK_BYTES) : 0);
}
size_t mlp_tc_bwd_workspace_bytes(int64_t n, bool x3) { return (size_t)(flnerf_padded_rows(n) / 128) * tc_tile_act_bytes(x3, n); }

int mlp_tc_pack_weights(flnerf_ctx *ctx, int kind, const float *params, void *packed, cudaStream_t st) {
  FL_REQUIRE(tc::setup_tables(ctx->sm_count) == 0, "mlp_tc: table setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  const tc::NetDesc &nd = tc::g_desc[kind];
  dim3 grid(8, nd.fwd_full + 4 + tc::DG_CHUNKS, 2);
  FL_LAUNCH(tc::pack_weights_kernel, grid, 256, 0, st, params, (uint8_t *)packed, kind, nd.packed_bytes);
  return 0;
}

int mlp_tc_forward(flnerf_ctx *ctx, bool x3, int kind, const float *params, const void *packed, int64_t n, int S,
                   const void *pe_tiles, const float *dirpe, float *raw, void *stash, int training, cudaStream_t st) {
  FL_REQUIRE(tc::setup_tables(ctx->sm_count) == 0, "mlp_tc: table setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  FL_REQUIRE(n % S == 0, "mlp_tc_forward: n=%lld is not a multiple of S=%d", (long long)n, S);
  const tc::NetDesc &nd = tc::g_desc[kind];
  int64_t B = n / S;
  float *vb = (float *)stash;
  FL_LAUNCH(tc::viewbias_kernel, (unsigned)ceil_div64(B, 32), 256, 0, st, B, params, dirpe, vb, nd.w_views, nd.b_views);
  tc::FwdParams p{};
  p.P = params; p.packed = (const uint8_t *)packed; p.pe_tiles = (const uint8_t *)pe_tiles; p.viewbias = vb;
  p.raw = raw; p.n = n; p.S = S; p.kind = kind; p.stash_lo = x3_stash_lo(x3, n) ? 1 : 0;
  int64_t n_pad = flnerf_padded_rows(n);
  p.n_pairs = (int)(n_pad / 256);
  if (training) {
    p.stash_act = (uint8_t *)stash + tc_vb_bytes(n, S);
    p.stash_mask = (uint32_t *)(p.stash_act + (size_t)(n_pad / 128) * tc_tile_act_bytes(x3, n));
  }
  FL_CHECK_CUDA(cudaMemsetAsync(raw, 0, (size_t)n * 4 * sizeof(float), st));  // column-half warps accumulate into it
  if (x3 || kind != 0) {
    // one 128-row tile per CTA, ring items consumed in order: any number of positional slabs, one or three MMA passes
    const int grid = tc::pair_grid(p.n_pairs * 2, ctx->sm_count);
    // split mode: activations travel to the next layer's MMA through TENSOR MEMORY (A-in-TMEM) unless the stash must keep
    // the lo images too (they are staged in shared memory then); $FLNERF_X3_ATMEM=0 selects the shared-memory variant
    static int atmem = -1;
    if (atmem < 0) { const char *e = getenv("FLNERF_X3_ATMEM"); atmem = e ? atoi(e) : 1; }
    const bool at = x3 && atmem && !(training && p.stash_lo);
    p.prof = (x3 && kind == 0) ? tc::prof_buffer() : nullptr;
    if (p.prof) {      // FLNERF_TC_PROF=1: per-role cycle accounting of the split forward kernel
      if (at) FL_LAUNCH((tc::mlp_fwd_gen<3, 1, true, true>), grid, tc::kThreads, tc::LayF::SMEM, st, p);
      else FL_LAUNCH((tc::mlp_fwd_gen<3, 1, false, true>), grid, tc::kThreads, tc::LayF::SMEM, st, p);
      tc::prof_report(at ? "fwd_x3_atmem" : "fwd_x3", grid, p.n_pairs, st);
      return 0;
    }
    if (at && kind == 0) FL_LAUNCH((tc::mlp_fwd_gen<3, 1, true>), grid, tc::kThreads, tc::LayF::SMEM, st, p);
    else if (at) FL_LAUNCH((tc::mlp_fwd_gen<3, 2, true>), grid, tc::kThreads, tc::LayF::SMEM, st, p);
    else if (x3 && kind == 0) FL_LAUNCH((tc::mlp_fwd_gen<3, 1>), grid, tc::kThreads, tc::LayF::SMEM, st, p);
    else if (x3) FL_LAUNCH((tc::mlp_fwd_gen<3, 2>), grid, tc::kThreads, tc::LayF::SMEM, st, p);
    else FL_LAUNCH((tc::mlp_fwd_gen<1, 2>), grid, tc::kThreads, tc::LayF::SMEM, st, p);
    return 0;
  }
  { const char *e = getenv("FLNERF_FWD_DBG"); p.dbg = e ? atoi(e) : 0; }
  p.prof = tc::prof_buffer();
  const int grid = tc::pair_grid(p.n_pairs, ctx->sm_count);
  if (p.prof) {
    FL_LAUNCH(tc::mlp_fwd_tc<true>, grid, tc::kThreads, tc::LayF::SMEM, st, p);
    tc::prof_report("fwd", grid, p.n_pairs, st);
  } else {
    FL_LAUNCH(tc::mlp_fwd_tc<false>, grid, tc::kThreads, tc::LayF::SMEM, st, p);
  }
  return 0;
}

int mlp_tc_backward(flnerf_ctx *ctx, bool x3, int kind, const float *params, const void *packed, int64_t n, int S,
                    const void *pe_tiles, const float *dirpe, const void *stash, const float *draw, float *grads, void *ws,
                    int stages, cudaStream_t st) {
  FL_REQUIRE(tc::setup_tables(ctx->sm_count) == 0, "mlp_tc: table setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  const tc::NetDesc &nd = tc::g_desc[kind];
  int64_t n_pad = flnerf_padded_rows(n);
  const uint8_t *stash_act = (const uint8_t *)stash + tc_vb_bytes(n, S);
  const uint32_t *stash_mask = (const uint32_t *)(stash_act + (size_t)(n_pad / 128) * tc_tile_act_bytes(x3, n));
  tc::DgradParams d{};
  d.P = params; d.packed_dg = (const uint8_t *)packed + nd.fwd_bytes; d.draw = draw; d.stash_mask = stash_mask;
  d.dy = (uint8_t *)ws; d.n = n; d.n_pairs = (int)(n_pad / 256); d.kind = kind; d.stash_lo = x3_stash_lo(x3, n) ? 1 : 0;
  if ((stages & 1) && x3) {
    const int grid = tc::pair_grid(d.n_pairs * 2, ctx->sm_count);
    FL_LAUNCH(tc::mlp_dgrad_x3, grid, tc::kThreads, tc::LayD::SMEM, st, d);
  } else if (stages & 1) {
    const int grid = tc::pair_grid(d.n_pairs, ctx->sm_count);
    d.prof = tc::prof_buffer();
    if (d.prof) {
      FL_LAUNCH(tc::mlp_dgrad_tc<true>, grid, tc::kThreads, tc::LayD::SMEM, st, d);
      tc::prof_report("dgrad", grid, d.n_pairs, st);
    } else {
      FL_LAUNCH(tc::mlp_dgrad_tc<false>, grid, tc::kThreads, tc::LayD::SMEM, st, d);
    }
  }
  tc::WgradParams w{};
  w.dy = (const uint8_t *)ws; w.stash_act = stash_act; w.pe_tiles = (const uint8_t *)pe_tiles; w.draw = draw;
  w.G = grads; w.n = n; w.n_tiles = (int)(n_pad / 128); w.dirpe = dirpe; w.S = S; w.kind = kind;
  w.tile_stride = tc_tile_act_bytes(x3, n); w.slot_stride = x3_stash_lo(x3, n) ? tc::SLOT_BYTES_X3 : 65536;
  w.helper_flags = 7;
  static long long *dbg = nullptr;
  const bool want_dbg = getenv("FLNERF_WG_DEBUG") != nullptr;
  if (want_dbg && !dbg) cudaMalloc(&dbg, sizeof(long long) * 4 * 1024);
  w.dbg = want_dbg ? dbg : nullptr;
  if ((stages & 2) && x3) {
    // dW = (dYhi + dYlo)^T (Xhi + Xlo) ~= dYhi^T Xhi [+ dYhi^T Xlo [+ dYlo^T Xhi]]: up to three passes of the same kernel over
    // the hi / lo images, each with the CUDA-core reductions that belong to its operands (see x3_wgrad_passes)
    const size_t pe_lo = (size_t)w.n_tiles * nd.pe_slabs * tc::PE_BYTES;
    const int passes = x3_wgrad_passes(n);
    FL_LAUNCH(tc::mlp_wgrad_tc, tc::g_wgrad_grid, tc::kWgThreads, tc::SMEM_WG, st, w);
    if (passes >= 2) {
      w.b_part = 65536; w.pe_part = pe_lo; w.helper_flags = 2;
      FL_LAUNCH(tc::mlp_wgrad_tc, tc::g_wgrad_grid, tc::kWgThreads, tc::SMEM_WG, st, w);
    }
    if (passes >= 3) {
      w.a_part = 65536; w.b_part = 0; w.pe_part = 0; w.helper_flags = 1;
      FL_LAUNCH(tc::mlp_wgrad_tc, tc::g_wgrad_grid, tc::kWgThreads, tc::SMEM_WG, st, w);
    }
  } else if (stages & 2) {
    FL_LAUNCH(tc::mlp_wgrad_tc, tc::g_wgrad_grid, tc::kWgThreads, tc::SMEM_WG, st, w);
  }
  if (want_dbg && (stages & 2)) {
    static int printed = 0;
    if (printed++ == 3) {  // a warmed-up launch
      long long h[4 * 1024];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, dbg, sizeof(long long) * 4 * tc::g_wgrad_grid, cudaMemcpyDeviceToHost);
      for (int i = 0; i < tc::g_wgrad_grid; ++i)
        fprintf(stderr, "wgdbg cta %d unit %lld loop %lld total %lld halves %lld\n", i, h[i * 4], h[i * 4 + 1], h[i * 4 + 2], h[i * 4 + 3]);
    }
  }
  const int tpb = 8;
  if ((stages & 8) && !x3 && kind == 0)
    FL_LAUNCH(tc::wgrad_small_kernel, (unsigned)((w.n_tiles + tpb - 1) / tpb), 512, 0, st, stash_act, w.dy, draw, dirpe,
              grads, n, S, w.n_tiles, tpb);
  return 0;
}
