// common.cuh -- shared host/device helpers for libflnerf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/flnerf.h"

struct flnerf_ctx {
  int device;
  int sm_count;
  int max_smem_optin;
  // device-side per-step scalars (flnerf_set_step_record): when set, the batch start, the Philox offsets and the Adam
  // step size come from device memory, so that a whole training step can be replayed as ONE CUDA graph
  const flnerf_step_record *step_rec = nullptr;
};

void flnerf_set_error(const char *fmt, ...);
extern long long g_flnerf_launches;

#define FL_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      flnerf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define FL_REQUIRE(cond, ...)                                                                \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      flnerf_set_error(__VA_ARGS__);                                                         \
      return 2;                                                                              \
    }                                                                                        \
  } while (0)

// every kernel launch goes through this so that gpu_launches in bench.py is a count, not a guess
#define FL_LAUNCH(kernel, grid, block, smem, stream, ...)                                    \
  do {                                                                                       \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                \
    ++g_flnerf_launches;                                                                     \
    FL_CHECK_CUDA(cudaGetLastError());                                                       \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (stateless; one call = 4 uniform u32)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// U[0,1) with 24 random bits, like torch.rand for fp32
__host__ __device__ inline float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------------------------------------
// SWIZZLE_128B K-major slab image: R rows x 64 bf16, 128 B per row; byte offset of element (r, c)
// (identical to what TMA SWIZZLE_128B / a UMMA SWIZZLE_128B descriptor expects: 16-byte chunk index
// XOR (row mod 8) inside every 1024-byte group of 8 rows)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t sw128_offset(uint32_t r, uint32_t c) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((c >> 3) ^ r) & 7u) << 4) + ((c & 7u) << 1);
}
