// Microbenchmark: how fast can ONE SM stream its shared memory to HBM?  (The MLP kernels stash 64 KB per tile-layer per SM.)
//   mode 0: cp.async.bulk shared -> global, `warps` issuing warps x one `piece`-byte copy each per round, wait_group.read 1
//   mode 1: st.global.v4 from registers, 512 threads, 512 contiguous bytes per warp instruction
// Every CTA writes its own write-once stream (no line is written twice).  Prints bytes / clock / SM and TB/s.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ub/store_bw tools/ub/store_bw.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../fast-learning-nerf_b200/csrc/tc_ptx.cuh"
using namespace tc;

__global__ void __launch_bounds__(512, 1) bulk_store(uint8_t *dst, int rounds, int warps, int piece, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
  fence_async_smem();
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  uint8_t *base = dst + (size_t)blockIdx.x * rounds * warps * piece;
  const long long t0 = clock64();
  if (warp < warps && (threadIdx.x & 31) == 0) {
    for (int r = 0; r < rounds; ++r) {
      bulk_s2g(base + ((size_t)r * warps + warp) * piece, smem_u32(smem + (warp * piece) % 65536), piece);
      bulk_commit();
      bulk_wait_read1();
    }
    bulk_wait_all0();
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

__global__ void __launch_bounds__(512, 1) stg_store(uint4 *dst, int rounds, long long *out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4 *base = dst + (size_t)blockIdx.x * rounds * 4096;      // 64 KB per round
  uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int k = 0; k < 8; ++k) base[(size_t)r * 4096 + (warp * 8 + k) * 32 + lane] = v;
    v.x += 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  int sms = 148;
  const size_t total = (size_t)4 << 30;
  uint8_t *dst; long long *out;
  cudaMalloc(&dst, total); cudaMalloc(&out, sizeof(long long) * sms);
  cudaFuncSetAttribute(bulk_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  auto report = [&](const char *name, size_t bytes_per_cta, float ms) {
    long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += (double)h[i];
    c /= sms;
    printf("%-34s %7.1f B/clk/SM   %6.2f TB/s   (%.3f ms)\n", name, bytes_per_cta / c, bytes_per_cta * (double)sms / (ms * 1e-3) / 1e12, ms);
  };
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfg[][2] = {{16, 4096}, {8, 8192}, {4, 16384}, {1, 16384}, {1, 65536}, {2, 32768}};
  for (auto &c : cfg) {
    const int warps = c[0], piece = c[1];
    const int rounds = (int)(total / sms / ((size_t)warps * piece));
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      bulk_store<<<sms, 512, 65536>>>(dst, rounds, warps, piece, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    char name[64]; snprintf(name, sizeof(name), "bulk %2d warps x %5d B", warps, piece);
    report(name, (size_t)rounds * warps * piece, ms);
  }
  {
    const int rounds = (int)(total / sms / 65536);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      stg_store<<<sms, 512>>>((uint4 *)dst, rounds, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    report("st.global.v4, 512 threads", (size_t)rounds * 65536, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
