// mlp_tc_x3.cuh -- FLNERF_MODE_BF16X3: the tensor-core MLP at fp32-grade precision (the mode that meets the
// reference's 1e-4 tolerance; model.py:38-63 is fp32 end to end).  Included by mlp_tc.cu inside namespace tc.
//
// Split precision: every fp32 operand x is carried as two bf16 numbers, hi = bf16(x) and lo = bf16(x - hi)
// (16 mantissa bits together), and every product A*B is evaluated as
//     Ahi*Bhi + Alo*Bhi + Ahi*Blo          (the dropped Alo*Blo term is <= 2^-18 relative)
// by THREE tcgen05.mma groups into the SAME fp32 TMEM accumulator.  Weights are split once per optimiser step by
// pack_weights_kernel (part 0 / part 1 of the packed image), activations are split by the epilogue that produces them
// (the fp32 value is in registers there), the positional-encoding tiles by encode_tc_x3_kernel.
//
// Kernel shape: the CTA-pair skeleton of mlp_fwd_tc / mlp_dgrad_tc (cta_group::2, M = 256 = 128 rows per CTA, bulk-copy
// ring, 16 epilogue warps), but a CTA carries ONE 128-row tile: its hi image lives where tile set A lived and its lo
// image where tile set B lived, so shared memory and TMEM use are unchanged.  The three-fold MMA time per layer
// (6144 cycles per tile-layer) makes the un-overlapped epilogue (~1500 cycles) a 20 % tax instead of the 50 % it would
// be at single precision; the ring items of a tile are consumed strictly in order (hi chunk, lo chunk, ...).
// The heads (alpha, rgb), biases, the per-ray view-direction bias and the ReLU masks are fp32 / exact as in the bf16 mode,
// but are fed the un-rounded fp32 activations.

constexpr int X3_DG_ITEMS = 2 * DG_CHUNKS;
constexpr size_t TILE_ACT_BYTES_X3 = 2 * TILE_ACT_BYTES;  // per slot: [hi 64 KB][lo 64 KB]
constexpr size_t SLOT_BYTES_X3 = 2 * 65536;

// TMEM map of the A-in-TMEM variants (kAT): fp32 accumulator in columns [0, 256), the NEXT layer's A operand as packed bf16
// pairs behind it -- hi image in [256, 384), lo image in [384, 512) (column 256 + k/2 holds features k, k+1): all 512 columns.
constexpr uint32_t TM_A_HI = 256, TM_A_LO = 384;

// 8 consecutive columns [c, c+8) of row r, split into hi / lo bf16.
//   kAT = false: both go to the shared-memory images (lo image = hi + ACT_BYTES) the next layer's MMA reads;
//   kAT = true : both go to TENSOR MEMORY (tcgen05.st; taddr = this thread's lane + TM_A_HI + c/2); the hi chunk is ALSO written
//                to shared memory when `smem_hi` (training: the stash bulk stores ship it from there).
template <bool kAT>
__device__ __forceinline__ void store_split8(uint8_t *row_hi, uint32_t pos, const float x[8], uint32_t taddr = 0, bool smem_hi = true) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = pack_bf16_fast(x[2 * e], x[2 * e + 1]);
    l[e] = pack_bf16_fast(x[2 * e] - bf16_lo(h[e]), x[2 * e + 1] - bf16_hi(h[e]));
  }
  if (kAT) {
    tmem_st4(taddr, h[0], h[1], h[2], h[3]);
    tmem_st4(taddr + (TM_A_LO - TM_A_HI), l[0], l[1], l[2], l[3]);
    if (smem_hi) *reinterpret_cast<uint4 *>(row_hi + pos) = make_uint4(h[0], h[1], h[2], h[3]);
  } else {
    *reinterpret_cast<uint4 *>(row_hi + pos) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(row_hi + ACT_BYTES + pos) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// forward: 32 accumulator columns [c0, c0+32) of row r -> +bias -> (relu) -> hi/lo -> act images.
// kType 0 relu, 1 relu + alpha head (fp32 activations), 2 linear.  Returns the ReLU mask word (see fwd_block).
template <int kType, bool kMask, bool kAT>
__device__ __forceinline__ uint32_t fwd_block_x3(const uint32_t v[32], const float *s_b, uint8_t *act_hi, uint32_t r,
                                                 uint32_t c0, const float *s_wa, float &alpha, uint32_t t_row, bool smem_hi) {
  uint32_t m = 0;
  uint8_t *row_hi = act_hi + (c0 >> 6) * SLAB_BYTES + (r >> 3) * 1024u + (r & 7u) * 128u;
  const uint32_t q0 = (c0 & 63u) >> 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float x[8];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float4 b = *reinterpret_cast<const float4 *>(s_b + c0 + 8 * g + 4 * e);
      x[4 * e + 0] = __uint_as_float(v[8 * g + 4 * e + 0]) + b.x;
      x[4 * e + 1] = __uint_as_float(v[8 * g + 4 * e + 1]) + b.y;
      x[4 * e + 2] = __uint_as_float(v[8 * g + 4 * e + 2]) + b.z;
      x[4 * e + 3] = __uint_as_float(v[8 * g + 4 * e + 3]) + b.w;
    }
    if (kMask) {
#pragma unroll
      for (int i = 0; i < 8; ++i) m = __funnelshift_l(__float_as_uint(x[i]), m, 1);
    }
    if (kType != 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
    }
    if (kType == 1) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float4 a = *reinterpret_cast<const float4 *>(s_wa + c0 + 8 * g + 4 * e);
        alpha = fmaf(x[4 * e + 0], a.x, alpha);
        alpha = fmaf(x[4 * e + 1], a.y, alpha);
        alpha = fmaf(x[4 * e + 2], a.z, alpha);
        alpha = fmaf(x[4 * e + 3], a.w, alpha);
      }
    }
    store_split8<kAT>(row_hi, ((q0 + g) ^ (r & 7u)) << 4, x, t_row + TM_A_HI + (c0 + 8 * g) / 2, smem_hi);
  }
  return m;
}

// t_row = TMEM address of this thread's lane at column 0 (kAT: where the next layer's A operand goes)
template <int kType, bool kMask, bool kAT>
__device__ __forceinline__ uint2 fwd_epilogue_x3(uint32_t tmem_rc, uint32_t cq, const float *s_bias, uint8_t *act_hi,
                                                 uint32_t r, const float *s_wa, float &alpha, uint32_t t_row, bool smem_hi) {
  uint32_t va[32];
  const uint32_t c0 = cq * 64;
  uint2 mk;
  tmem_ld32(tmem_rc, va);
  tmem_ld_wait(va);
  mk.x = fwd_block_x3<kType, kMask, kAT>(va, s_bias, act_hi, r, c0, s_wa, alpha, t_row, smem_hi);
  tmem_ld32(tmem_rc + 32, va);
  tmem_ld_wait(va);
  mk.y = fwd_block_x3<kType, kMask, kAT>(va, s_bias, act_hi, r, c0 + 32, s_wa, alpha, t_row, smem_hi);
  return mk;
}

// views_linears.0 (N=128: 64 columns per warp) + rgb_linear on CUDA cores from the fp32 activations
template <bool kAT, class Params>
__device__ __forceinline__ uint2 fwd_views_rgb_x3(const Params &p, uint32_t tmem_row, uint32_t ch, uint8_t *act_hi, uint32_t r,
                                                  int64_t row, bool live, const float *s_head, bool smem_hi) {
  const int64_t ray = (row < p.n ? row : p.n - 1) / p.S;
  const float4 *vb4 = reinterpret_cast<const float4 *>(p.viewbias + ray * 128);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f;
  uint32_t mk2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t cb = ch * 2 + i;
    uint32_t v[32], m = 0;
    tmem_ld32(tmem_row + cb * 32, v);
    tmem_ld_wait(v);
    uint8_t *row_hi = act_hi + (cb >> 1) * SLAB_BYTES + (r >> 3) * 1024u + (r & 7u) * 128u;
    const uint32_t q0 = (cb & 1u) * 4u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x[8];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float4 b = __ldg(vb4 + cb * 8 + 2 * g + e);
        x[4 * e + 0] = __uint_as_float(v[8 * g + 4 * e + 0]) + b.x;
        x[4 * e + 1] = __uint_as_float(v[8 * g + 4 * e + 1]) + b.y;
        x[4 * e + 2] = __uint_as_float(v[8 * g + 4 * e + 2]) + b.z;
        x[4 * e + 3] = __uint_as_float(v[8 * g + 4 * e + 3]) + b.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        m = __funnelshift_l(__float_as_uint(x[j]), m, 1);
        x[j] = fmaxf(x[j], 0.f);
        const int k = cb * 32 + 8 * g + j;
        c0 = fmaf(x[j], s_head[k], c0);
        c1 = fmaf(x[j], s_head[128 + k], c1);
        c2 = fmaf(x[j], s_head[256 + k], c2);
      }
      if (kAT) {   // h9 feeds no MMA: only its hi image is needed, in shared memory, for the stash
        if (smem_hi) {
          uint32_t h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = pack_bf16_fast(x[2 * e], x[2 * e + 1]);
          *reinterpret_cast<uint4 *>(row_hi + (((q0 + g) ^ (r & 7u)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
        }
      } else if (smem_hi) {   // h9 feeds no MMA: its images are only staged for the stash
        store_split8<false>(row_hi, ((q0 + g) ^ (r & 7u)) << 4, x);
      }
    }
    mk2[i] = m;
  }
  if (live && row < p.n) {
    const float bsel = ch == 0 ? 1.f : 0.f;
    atomicAdd(p.raw + row * 4 + 0, c0 + bsel * s_head[384]);
    atomicAdd(p.raw + row * 4 + 1, c1 + bsel * s_head[385]);
    atomicAdd(p.raw + row * 4 + 2, c2 + bsel * s_head[386]);
  }
  return make_uint2(mk2[0], mk2[1]);
}

// a warp ships its own 4 KB pieces (32 rows x one slab) of the hi and the lo image as ONE bulk group
__device__ __forceinline__ void warp_store_slab_x3(uint8_t *dst_slot, const uint8_t *act_hi, uint32_t quarter, uint32_t slab) {
  const uint32_t off = slab * SLAB_BYTES + quarter * 4096u;
  bulk_s2g(dst_slot + off, smem_u32(act_hi + off), 4096u);
  bulk_s2g(dst_slot + 65536 + off, smem_u32(act_hi + ACT_BYTES + off), 4096u);
  bulk_commit();
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// pair MMA over one 64-wide K chunk with A in tensor memory: 4 k-steps of 16 features = 8 packed columns each
__device__ __forceinline__ void issue_chunk_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_smem, uint32_t idesc, bool first) {
  const uint64_t db = make_smem_desc(b_smem, 0, 1024);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma2_bf16_ts(tmem_d, tmem_a + 8 * k, db + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
}

// =================================================================================================
// forward, one tile per CTA.  kPasses = 3: split precision (hi and lo images, three MMA groups per chunk); kPasses = 1: plain
// bf16 (hi image only) -- used for networks whose positional input does not fit the ping-pong kernel's ring program
// (kPeSlabs = 2: the 84-channel nerf++ background network, nerf++-ours/nerf_network.py:70-142).
// Ring items of one tile, in consumption order (each "x" below is one item when kPasses = 1, a (hi, lo) pair when 3):
//   L0     : for every positional slab s: PE_s, W0_s
//   L1..L4 : 4 x W
//   L5     : 4 x W (activation slabs), then for every positional slab s: Wpe_s, PE_s
//   L6..L9 : 4 x W
// =================================================================================================
template <int kPasses, int kPeSlabs, bool kAT = false, bool kProf = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mlp_fwd_gen(FwdParams p) {
  constexpr bool kX3 = kPasses == 3;
  const bool prof_on = kProf && p.prof != nullptr;      // per-role cycle accounting (FLNERF_TC_PROF), compiled out otherwise
  static_assert(!kAT || kX3, "A-in-TMEM is the split-precision kernel's variant (one tile per CTA leaves 256 TMEM columns free)");
  constexpr int kPer = kX3 ? 2 : 1;                                    // ring items per operand
  constexpr int kItems = kPer * (2 * kPeSlabs + 16 + 4 + 2 * kPeSlabs + 16);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t bar = smem_u32(smem + LayF::OFF_BAR);
  const uint32_t cr = cluster_ctarank();
  const uint32_t tmem_base = pair_setup<LayF>(smem, bar, cr, warp, lane, p.P, p.kind);
  const uint32_t s_act = smem_u32(smem + OFF_ACT), s_w = smem_u32(smem + OFF_W);
  const int n_tiles = p.n_pairs * 2;
  const int iters = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int F = c_desc[p.kind].fwd_full;
  const size_t part_bytes = c_desc[p.kind].packed_bytes;

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
      auto push_w = [&](int cc) {  // this CTA's half (N/2 rows) of forward chunk cc: hi part (, then lo part)
        const uint32_t half = cc < F ? 16384u : 8192u;
        const size_t off = cc < F ? (size_t)cc * 32768 : (size_t)F * 32768 + (size_t)(cc - F) * 16384;
        push(p.packed + off + (size_t)cr * half, half);
        if (kX3) push(p.packed + part_bytes + off + (size_t)cr * half, half);
      };
      const size_t pe_lo = (size_t)n_tiles * kPeSlabs * PE_BYTES;
      for (int it = 0; it < iters; ++it) {
        const int tile = min(n_tiles - 1, (int)blockIdx.x + it * (int)gridDim.x);
        const uint8_t *pe = p.pe_tiles + (size_t)tile * kPeSlabs * PE_BYTES;
        auto push_pe = [&](int sl) {
          push(pe + (size_t)sl * PE_BYTES, PE_BYTES);
          if (kX3) push(pe + pe_lo + (size_t)sl * PE_BYTES, PE_BYTES);
        };
        for (int sl = 0; sl < kPeSlabs; ++sl) { push_pe(sl); push_w(sl); }
        int ci = kPeSlabs;
        for (int L = 1; L < 10; ++L) {
          if (L == 5) {
            for (int c = 0; c < 4; ++c) push_w(ci + kPeSlabs + c);
            for (int sl = 0; sl < kPeSlabs; ++sl) { push_w(ci + sl); push_pe(sl); }
            ci += kPeSlabs + 4;
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
      const int total = iters * kItems;
      for (int q = 0; q < total; ++q) {
        mbar_wait(LayF::w_full(bar, ring.stage), ring.phase);
        mbar_arrive_cluster(remote_full + 8u * ring.stage);
        ring.next();
      }
    } else if (lane == 0) {
      // -------------------------------------------------------------- pair MMA issuer (leader)
      const uint32_t idesc256 = make_idesc(256, 256, 0, 0), idesc128 = make_idesc(256, 128, 0, 0);
      uint32_t q = 0, n_act = 0;
      long long wa = 0, ww = 0;
      const long long mt0 = prof_on ? clock64() : 0;
      auto wait_full = [&](uint32_t i) {
        const long long c0 = prof_on ? clock64() : 0;
        mbar_wait(LayF::w_full(bar, LayF::item_stage(i)), LayF::item_phase(i));
        if (prof_on) ww += clock64() - c0;
      };
      auto release = [&](uint32_t i) { umma2_commit_multicast(LayF::w_empty(bar, LayF::item_stage(i)), (uint16_t)3); };
      auto st = [&](uint32_t i) { return s_w + LayF::item_stage(i) * WSTAGE; };
      const uint32_t a_hi = s_act, a_lo = s_act + ACT_BYTES;
      // one positional slab: its PE item(s) and its weight item(s) sit at ring positions pe0.. and w0.. (kPer items each)
      auto pe_group = [&](uint32_t pe0, uint32_t w0, bool first) {
        for (uint32_t i = 0; i < 2u * kPer; ++i) wait_full(q + i);
        tc_fence_after();
        issue_chunk(tmem_base, st(pe0), st(w0), idesc256, first);
        if (kX3) {
          issue_chunk(tmem_base, st(pe0 + 1), st(w0), idesc256, false);
          issue_chunk(tmem_base, st(pe0), st(w0 + 1), idesc256, false);
        }
        for (uint32_t i = 0; i < 2u * kPer; ++i) release(q + i);
        q += 2 * kPer;
      };
      for (int it = 0; it < iters; ++it) {
        for (int L = 0; L < 10; ++L) {
          if (!(it == 0 && L == 0)) {  // both CTAs' epilogue warps: inputs written, accumulator drained
            const long long c0 = prof_on ? clock64() : 0;
            mbar_wait(LayF::act_ready(bar, 0), n_act & 1);
            if (prof_on) wa += clock64() - c0;
            ++n_act;
          }
          tc_fence_after();
          if (L == 0) {
            for (int sl = 0; sl < kPeSlabs; ++sl) pe_group(q, q + kPer, sl == 0);
          } else {
            const uint32_t idesc = L == 9 ? idesc128 : idesc256;
            for (uint32_t c = 0; c < 4; ++c) {
              wait_full(q);
              tc_fence_after();
              if (kAT) {   // A (hi, then lo) from tensor memory, B = the hi weight chunk; then A hi x the lo chunk
                issue_chunk_ts(tmem_base, tmem_base + TM_A_HI + c * 32, st(q), idesc, c == 0);
                issue_chunk_ts(tmem_base, tmem_base + TM_A_LO + c * 32, st(q), idesc, false);
                release(q);
                wait_full(q + 1);
                tc_fence_after();
                issue_chunk_ts(tmem_base, tmem_base + TM_A_HI + c * 32, st(q + 1), idesc, false);
                release(q + 1);
              } else {
                issue_chunk(tmem_base, a_hi + c * SLAB_BYTES, st(q), idesc, c == 0);
                if (kX3) {
                  issue_chunk(tmem_base, a_lo + c * SLAB_BYTES, st(q), idesc, false);
                  release(q);
                  wait_full(q + 1);
                  tc_fence_after();
                  issue_chunk(tmem_base, a_hi + c * SLAB_BYTES, st(q + 1), idesc, false);
                  release(q + 1);
                } else {
                  release(q);
                }
              }
              q += kPer;
            }
            if (L == 5)
              for (int sl = 0; sl < kPeSlabs; ++sl) pe_group(q + kPer, q, false);
          }
          umma2_commit_multicast(LayF::acc_full(bar, 0), (uint16_t)3);
        }
      }
      if (prof_on) {
        long long *o = p.prof + blockIdx.x * PROF_SLOTS;
        o[0] = clock64() - mt0; o[1] = wa; o[2] = 0; o[3] = ww; o[16] = 0; o[17] = 0;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 16 warps (both CTAs)
    const int e = warp - kEpiWarp0;
    const uint32_t cq = (uint32_t)e >> 2;    // column quarter: columns [64 cq, 64 cq + 64) = slab cq
    const uint32_t quarter = warp & 3;       // TMEM lane quarter this warp may access
    const uint32_t r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const float *s_bias = reinterpret_cast<const float *>(smem + LayF::OFF_VEC);
    const float *s_head = reinterpret_cast<const float *>(smem + LayF::OFF_HEAD);
    const float *s_wa = s_head + 388;
    float *s_alpha = reinterpret_cast<float *>(smem + LayF::OFF_ALPHA);
    uint8_t *act_hi = smem + OFF_ACT;
    const uint32_t tmem_rc = tmem_base + ((quarter * 32) << 16) + cq * 64;
    const bool prof = prof_on && e == 0 && lane == 0;
    long long e_acc = 0, e_st = 0, e_body = 0, e_tail = 0;
    const long long et0 = prof ? clock64() : 0;
    const bool stash_lo = kX3 && p.stash_lo;
    const size_t tile_bytes = stash_lo ? TILE_ACT_BYTES_X3 : TILE_ACT_BYTES, slot_bytes = stash_lo ? SLOT_BYTES_X3 : 65536;
    uint32_t n_layer = 0;
    bool store_pending = false;
    for (int it = 0; it < iters; ++it) {
      const int tile_raw = (int)blockIdx.x + it * (int)gridDim.x;
      const bool live = tile_raw < n_tiles;            // surplus iteration: compute, but write nothing
      const int64_t tile = live ? tile_raw : n_tiles - 1;
      const int64_t row = tile * 128 + r;
      uint8_t *stash_act = live ? p.stash_act : nullptr;
      uint32_t *stash_mask = live ? p.stash_mask : nullptr;
      for (int L = 0; L < 10; ++L, ++n_layer) {
        const bool has_cols = L < 9 || cq < 2;   // the views layer has 128 outputs: column quarters 0,1 only
        const long long c0 = prof ? clock64() : 0;
        mbar_wait(LayF::acc_full(bar, 0), n_layer & 1);
        tc_fence_after();
        const long long c1 = prof ? clock64() : 0;
        if (p.stash_act) {  // my store of the previous layer reads the piece this layer overwrites
          if (lane == 0 && store_pending) bulk_wait_read0();
          __syncwarp();
        }
        const long long c2 = prof ? clock64() : 0;
        float alpha = 0.f;
        uint2 mk = make_uint2(0u, 0u);
        const bool want_mask = stash_mask != nullptr;
        const uint32_t t_row = tmem_base + ((quarter * 32) << 16);
        const bool smem_hi = p.stash_act != nullptr;      // kAT: shared memory only stages the stash stores
        if (L <= 7) {
          const float *sb = s_bias + L * 256;
          if (kX3) {
            if (L < 7) {
              if (want_mask) mk = fwd_epilogue_x3<0, true, kAT>(tmem_rc, cq, sb, act_hi, r, s_wa, alpha, t_row, smem_hi);
              else fwd_epilogue_x3<0, false, kAT>(tmem_rc, cq, sb, act_hi, r, s_wa, alpha, t_row, smem_hi);
            } else {
              if (want_mask) mk = fwd_epilogue_x3<1, true, kAT>(tmem_rc, cq, sb, act_hi, r, s_wa, alpha, t_row, smem_hi);
              else fwd_epilogue_x3<1, false, kAT>(tmem_rc, cq, sb, act_hi, r, s_wa, alpha, t_row, smem_hi);
            }
          } else {
            if (L < 7) {
              if (want_mask) mk = fwd_epilogue_q<0, true>(tmem_rc, cq, sb, act_hi, r, s_wa, alpha);
              else fwd_epilogue_q<0, false>(tmem_rc, cq, sb, act_hi, r, s_wa, alpha);
            } else {
              if (want_mask) mk = fwd_epilogue_q<1, true>(tmem_rc, cq, sb, act_hi, r, s_wa, alpha);
              else fwd_epilogue_q<1, false>(tmem_rc, cq, sb, act_hi, r, s_wa, alpha);
            }
          }
          if (L == 7) {
            // alpha_linear: the four column quarters of a row are summed in a fixed order (bit-reproducible)
            s_alpha[cq * 128 + r] = alpha;
            named_bar_sync(1 + quarter, 128);
            if (cq == 0 && live && row < p.n) {
              const float *pa = s_alpha + r;
              p.raw[row * 4 + 3] = ((pa[0] + pa[128]) + (pa[256] + pa[384])) + s_head[387];
            }
          }
        } else if (L == 8) {
          if (kX3) fwd_epilogue_x3<2, false, kAT>(tmem_rc, cq, s_bias + 8 * 256, act_hi, r, s_wa, alpha, t_row, smem_hi);
          else fwd_epilogue_q<2, false>(tmem_rc, cq, s_bias + 8 * 256, act_hi, r, s_wa, alpha);
        } else if (has_cols) {
          if (kX3) mk = fwd_views_rgb_x3<kAT>(p, tmem_rc - cq * 64, cq, act_hi, r, row, live, s_head, smem_hi);
          else mk = fwd_views_rgb(p, tmem_rc - cq * 64, cq, act_hi, r, row, live, s_head);
        }
        if (kAT) tmem_st_wait();       // the A operand is in tensor memory before the issuer hears of it
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        const long long c3 = prof ? clock64() : 0;
        if (lane == 0) {
          mbar_arrive_cluster(mapa_cluster(LayF::act_ready(bar, 0), 0));  // the leader's barrier
          if (stash_act && has_cols) {
            uint8_t *dst = stash_act + (size_t)tile * tile_bytes + (size_t)L * slot_bytes;
            if (stash_lo) warp_store_slab_x3(dst, act_hi, quarter, cq);
            else warp_store_slabs(dst, act_hi, quarter, cq, 1);
            store_pending = true;
          }
        }
        if (want_mask && L != 8 && has_cols)
          *reinterpret_cast<uint2 *>(stash_mask + mask_word_offset(tile, L < 9 ? L : 8, cq, r)) = mk;
        if (prof) { e_acc += c1 - c0; e_st += c2 - c1; e_body += c3 - c2; e_tail += clock64() - c3; }
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
// backward data-gradient chain, split precision
// =================================================================================================
template <bool kAlpha, bool kUseMask>
__device__ __forceinline__ void dgrad_block_x3(const uint32_t v[32], uint32_t m, float dsig, const float *s_wa, uint8_t *act_hi,
                                               uint32_t r, uint32_t c0) {
  uint8_t *row_hi = act_hi + (c0 >> 6) * SLAB_BYTES + (r >> 3) * 1024u + (r & 7u) * 128u;
  const uint32_t q0 = (c0 & 63u) >> 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = 8 * g + j;
      float gv = __uint_as_float(v[i]);
      if (kAlpha) gv = fmaf(dsig, s_wa[c0 + i], gv);
      if (kUseMask) gv = (m & (0x80000000u >> i)) ? 0.f : gv;  // mask bit 31 - i set = output i was inactive
      x[j] = gv;
    }
    store_split8<false>(row_hi, ((q0 + g) ^ (r & 7u)) << 4, x);
  }
}

template <bool kAlpha, bool kUseMask>
__device__ __forceinline__ void dgrad_epilogue_x3(uint32_t tmem_rc, uint32_t cq, const uint2 mk, float dsig, const float *s_wa,
                                                  uint8_t *act_hi, uint32_t r) {
  uint32_t va[32];
  const uint32_t c0 = cq * 64;
  tmem_ld32(tmem_rc, va);
  tmem_ld_wait(va);
  dgrad_block_x3<kAlpha, kUseMask>(va, mk.x, dsig, s_wa, act_hi, r, c0);
  tmem_ld32(tmem_rc + 32, va);
  tmem_ld_wait(va);
  dgrad_block_x3<kAlpha, kUseMask>(va, mk.y, dsig, s_wa, act_hi, r, c0 + 32);
}

// stage -1: G9 = (d_rgb * W_rgb) masked by relu(h9) -> slabs 0,1 (64 columns per warp)
__device__ __forceinline__ void dgrad_g9_x3(const float4 dr, const uint2 mk, uint32_t ch, const float *s_head, uint8_t *act_hi,
                                            uint32_t r) {
  const uint32_t mw[2] = {mk.x, mk.y};
#pragma unroll
  for (int j2 = 0; j2 < 2; ++j2) {
    const uint32_t cb = ch * 2 + j2;
    uint8_t *row_hi = act_hi + (cb >> 1) * SLAB_BYTES + (r >> 3) * 1024u + (r & 7u) * 128u;
    const uint32_t q0 = (cb & 1u) * 4u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = 8 * g + j, k = cb * 32 + i;
        const float gv = dr.x * s_head[k] + dr.y * s_head[128 + k] + dr.z * s_head[256 + k];
        x[j] = (mw[j2] & (0x80000000u >> i)) ? 0.f : gv;
      }
      store_split8<false>(row_hi, ((q0 + g) ^ (r & 7u)) << 4, x);
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mlp_dgrad_x3(DgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t bar = smem_u32(smem + LayD::OFF_BAR);
  const uint32_t cr = cluster_ctarank();
  const uint32_t tmem_base = pair_setup<LayD>(smem, bar, cr, warp, lane, p.P, p.kind);
  const uint32_t s_act = smem_u32(smem + OFF_ACT), s_w = smem_u32(smem + OFF_W);
  const int n_tiles = p.n_pairs * 2;
  const int iters = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      LayD::Ring ring;
      for (int it = 0; it < iters; ++it)
        for (int ci = 0; ci < DG_CHUNKS; ++ci)
          for (int part = 0; part < 2; ++part) {
            mbar_wait(LayD::w_empty(bar, ring.stage), ring.phase ^ 1);
            mbar_arrive_expect_tx(LayD::w_full(bar, ring.stage), 16384u);
            bulk_g2s(s_w + ring.stage * WSTAGE, p.packed_dg + (size_t)part * c_desc[p.kind].packed_bytes + (size_t)ci * 32768 + (size_t)cr * 16384,
                     16384u, LayD::w_full(bar, ring.stage));
            ring.next();
          }
    }
  } else if (warp == 1) {
    if (lane == 0 && cr != 0) {
      LayD::Ring ring;
      const uint32_t remote_full = mapa_cluster(LayD::w_full(bar, 0), 0);
      const int total = iters * X3_DG_ITEMS;
      for (int q = 0; q < total; ++q) {
        mbar_wait(LayD::w_full(bar, ring.stage), ring.phase);
        mbar_arrive_cluster(remote_full + 8u * ring.stage);
        ring.next();
      }
    } else if (lane == 0) {
      const uint32_t idesc = make_idesc(256, 256, 0, 0);
      const uint32_t a_hi = s_act, a_lo = s_act + ACT_BYTES;
      uint32_t q = 0, n_act = 0;
      auto wait_full = [&](uint32_t i) { mbar_wait(LayD::w_full(bar, LayD::item_stage(i)), LayD::item_phase(i)); };
      auto release = [&](uint32_t i) { umma2_commit_multicast(LayD::w_empty(bar, LayD::item_stage(i)), (uint16_t)3); };
      auto st = [&](uint32_t i) { return s_w + LayD::item_stage(i) * WSTAGE; };
      for (int it = 0; it < iters; ++it) {
        for (int D = 0; D < 9; ++D) {  // D=0: dF = G9 * Wv (K=128); D>=1: K=256
          const uint32_t nch = (D == 0) ? 2u : 4u;
          mbar_wait(LayD::act_ready(bar, 0), n_act & 1);
          ++n_act;
          tc_fence_after();
          for (uint32_t c = 0; c < nch; ++c) {
            wait_full(q);
            tc_fence_after();
            issue_chunk(tmem_base, a_hi + c * SLAB_BYTES, st(q), idesc, c == 0);
            issue_chunk(tmem_base, a_lo + c * SLAB_BYTES, st(q), idesc, false);
            release(q);
            wait_full(q + 1);
            tc_fence_after();
            issue_chunk(tmem_base, a_hi + c * SLAB_BYTES, st(q + 1), idesc, false);
            release(q + 1);
            q += 2;
          }
          umma2_commit_multicast(LayD::acc_full(bar, 0), (uint16_t)3);
        }
      }
    }
  } else {
    const int e = warp - kEpiWarp0;
    const uint32_t cq = (uint32_t)e >> 2;
    const uint32_t quarter = warp & 3;
    const uint32_t r = quarter * 32 + lane;
    const float *s_head = reinterpret_cast<const float *>(smem + LayD::OFF_HEAD);
    const float *s_wa = s_head + 388;
    uint8_t *act_hi = smem + OFF_ACT;
    const uint32_t tmem_rc = tmem_base + ((quarter * 32) << 16) + cq * 64;
    uint32_t n_layer = 0;
    bool store_pending = false;
    for (int it = 0; it < iters; ++it) {
      const int tile_raw = (int)blockIdx.x + it * (int)gridDim.x;
      const bool live = tile_raw < n_tiles;
      const int64_t tile = live ? tile_raw : n_tiles - 1;
      const int64_t row = tile * 128 + r;
      // stage -1: G9 from d_rgb; stages 0..8 = tensor layers (see mlp_dgrad_tc for the mask slots)
      for (int D = -1; D < 9; ++D) {
        const bool has_cols = D >= 0 || cq < 2;  // G9 has 128 columns: column quarters 0,1 only
        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (D <= 1 && D != 0) {
          if (row < p.n) dr = __ldg(reinterpret_cast<const float4 *>(p.draw) + row);
        }
        uint2 mk = make_uint2(0u, 0u);
        if (D != 0 && has_cols)
          mk = __ldg(reinterpret_cast<const uint2 *>(p.stash_mask + mask_word_offset(tile, D < 0 ? 8 : 8 - D, cq, r)));
        if (D >= 0) {
          mbar_wait(LayD::acc_full(bar, 0), n_layer & 1);
          tc_fence_after();
        }
        if (lane == 0 && store_pending) bulk_wait_read0();  // my previous store reads the piece this stage overwrites
        __syncwarp();
        if (D < 0) {
          if (has_cols) dgrad_g9_x3(dr, mk, cq, s_head, act_hi, r);
        } else if (D == 0) {
          dgrad_epilogue_x3<false, false>(tmem_rc, cq, mk, 0.f, s_wa, act_hi, r);
        } else if (D == 1) {  // dH7 also receives d_sigma * w_alpha (alpha_linear reads H7)
          dgrad_epilogue_x3<true, true>(tmem_rc, cq, mk, dr.w, s_wa, act_hi, r);
        } else {
          dgrad_epilogue_x3<false, true>(tmem_rc, cq, mk, 0.f, s_wa, act_hi, r);
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (D < 8) mbar_arrive_cluster(mapa_cluster(LayD::act_ready(bar, 0), 0));  // the last stage feeds no further MMA
          if (live && has_cols) {
            const int slot = (D < 0) ? 9 : 8 - D;  // D=0 -> dF (8), D=1 -> dH7 (7) ... D=8 -> dH0 (0)
            if (p.stash_lo) warp_store_slab_x3(p.dy + (size_t)tile * TILE_ACT_BYTES_X3 + (size_t)slot * SLOT_BYTES_X3, act_hi, quarter, cq);
            else warp_store_slabs(p.dy + (size_t)tile * TILE_ACT_BYTES + (size_t)slot * 65536, act_hi, quarter, cq, 1);
            store_pending = true;
          }
        }
        if (D >= 0) ++n_layer;
      }
    }
    if (lane == 0 && store_pending) bulk_wait_all0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem_base, 512);
}
