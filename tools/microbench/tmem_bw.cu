// Micro-benchmark: tcgen05.ld / tcgen05.st throughput (TMEM <-> registers) per SM, for 4 / 8 / 16 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../long-context-asr_b200/csrc/sm100_ptx.cuh"
using namespace lcasr::ptx;

template <int MODE>  // 0: ld x32, 1: st x32, 2: ld x32 + 32 FFMA on the values
__global__ void k(int iters, uint32_t* out, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (((uint32_t)(warp & 3) * 32) << 16);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = lane + i;
  // init TMEM so loads read defined data
  for (int c = 0; c < 512; c += 32) tmem_st_32x32b_x32(base + c, r);
  tmem_wait_st();
  __syncthreads();
  float acc = 0.f;
  uint32_t a32[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a32[i] = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t col = ((it * 4 + c) * 32 + (warp >> 2) * 128) & 511;
      if (MODE == 1) { tmem_st_32x32b_x32(base + col, r); }
      else {
        tmem_ld_32x32b_x32(base + col, r);
        if (MODE == 2) { tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) a32[i] ^= r[i]; }
      }
    }
    if (MODE == 1) tmem_wait_st(); else tmem_wait_ld();
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += r[i] + a32[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint32_t)acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int MODE> void run(const char* name, int warps, int iters) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, warps * 32>>>(iters, out, cyc);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(iters, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double bytes = (double)warps * iters * 4 * 32 * 32 * 4;  // per SM
  printf("%-10s warps=%2d: %lld cycles, %.1f B/clk/SM  (%s)\n", name, warps, h[0], bytes / h[0], cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) { run<0>("ld.x32", w, 2000); run<1>("st.x32", w, 2000); run<2>("ld+use", w, 2000); }
  return 0;
}
