// Micro-benchmark: tcgen05.mma dispatch rate from one thread for several shapes (cycles per MMA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../long-context-asr_b200/csrc/sm100_ptx.cuh"
using namespace lcasr::ptx;

// MODE 0: SS K-major SW128 ; 1: TS (A from TMEM), B K-major SW128 ; 2: TS, B MN-major SW64 (the Dh=32 PV form)
template <int N, int MODE>
__global__ void k(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, MODE == 2 ? 1 : 0);
    const uint64_t adesc = make_smem_desc_kmajor(base, 1024, kLayoutSW128);
    uint64_t bdesc = make_smem_desc_kmajor(base + 16384, 1024, kLayoutSW128);
    if (MODE == 2) bdesc = make_smem_desc_kmajor(base + 16384, 512, kLayoutSW64);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (MODE == 0) umma_f16_ss(tm, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1);
          else umma_f16_ts(tm, tm + 256 + kk * 8, bdesc + (MODE == 2 ? 64 * kk : 2 * kk), idesc, 1);
        }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int MODE> void run(const char* name) {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<N, MODE><<<148, 128, 64 * 1024>>>(iters, cyc); cudaDeviceSynchronize();
  k<N, MODE><<<148, 128, 64 * 1024>>>(iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d: %.1f cycles/MMA (ideal %d)  %s\n", name, N, (double)h[0] / (iters * 4), 128 * N / 256, cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  run<256, 0>("SS kmajor sw128"); run<128, 0>("SS kmajor sw128"); run<64, 0>("SS kmajor sw128"); run<32, 0>("SS kmajor sw128");
  run<128, 1>("TS, B kmajor sw128"); run<64, 1>("TS, B kmajor sw128"); run<32, 1>("TS, B kmajor sw128");
  run<32, 2>("TS, B mn-major sw64"); run<64, 2>("TS, B mn-major sw64");
  return 0;
}
