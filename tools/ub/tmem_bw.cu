// Microbenchmark: TMEM read bandwidth (tcgen05.ld.32x32b.x32) per SM as a function of warps and loads in flight.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ub/tmem_bw tools/ub/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../fast-learning-nerf_b200/csrc/tc_ptx.cuh"
using namespace tc;

template <int kInFlight>
__global__ void __launch_bounds__(512, 1) tmem_read(int iters, int nwarps, long long *out, uint32_t *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (((uint32_t)(warp & 3) * 32u) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    uint32_t va[32], vb[32];
    for (int i = 0; i < iters; ++i) {
      const uint32_t col = (uint32_t)((i * 64 + (warp >> 2) * 128) & 448);
      tmem_ld32(base + col, va);
      if (kInFlight == 2) {
        tmem_ld32(base + col + 32, vb);
        tmem_ld_wait2(va, vb);
        acc ^= va[3] ^ vb[5];
      } else {
        tmem_ld_wait(va);
        acc ^= va[3];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long *out; uint32_t *sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 20000;
  for (int fl = 1; fl <= 2; ++fl)
    for (int nw : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (fl == 1) tmem_read<1><<<148, 512>>>(iters, nw, out, sink); else tmem_read<2><<<148, 512>>>(iters, nw, out, sink);
        cudaDeviceSynchronize();
      }
      long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = (double)iters * nw * fl * 32 * 32 * 4;
      printf("in_flight %d warps %2d: %lld cycles, %.1f B/clk/SM, %.1f clk per LDTM.x32 per warp  (%s)\n", fl, nw, h[0], bytes / h[0],
             (double)h[0] / (iters * fl), cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
