// Micro-benchmark: one softmax iteration of the attention kernel without any barriers:
// LDTM 128 cols -> row max (FMNMX3) -> 128 x (ffma, ex2, fadd) + 64 packs -> STTM 64 cols, per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../long-context-asr_b200/csrc/sm100_ptx.cuh"
using namespace lcasr::ptx;

template <int MODE>  // 0 full, 1 no STTM, 2 no LDTM (reuse regs), 3 exp only
__global__ void k(int iters, float* out, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (((uint32_t)(warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = __float_as_uint(-0.01f * (lane + i));
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_st_32x32b_x32(base + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
  tmem_wait_st();
  __syncthreads();
  float m_run = 0.f, l_run = 0.f;
  const float scale_log2 = 0.255f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE != 2 && MODE != 3) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(base + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_wait_ld();
    }
    if (MODE != 3) {
      float mxa[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 128; i += 8)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(mxa[u]) : "f"(__uint_as_float(s[i + 2 * u])), "f"(__uint_as_float(s[i + 2 * u + 1])));
      m_run = fmaxf(m_run, fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * scale_log2);
    }
    float sums[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float neg_m = -m_run;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[c * 64 + 2 * i]), scale_log2, neg_m));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[c * 64 + 2 * i + 1]), scale_log2, neg_m));
        sums[(2 * i) & 7] += p0; sums[(2 * i + 1) & 7] += p1;
        __nv_bfloat162 pp = __floats2bfloat162_rn(p0, p1);
        pk[i] = *reinterpret_cast<uint32_t*>(&pp);
      }
      if (MODE == 0 || MODE == 2) tmem_st_32x32b_x32(base + c * 32, pk);
      else { s[c * 64] ^= pk[0] & 1; s[c * 64 + 1] ^= pk[31] & 1; }
    }
    l_run += ((sums[0] + sums[1]) + (sums[2] + sums[3])) + ((sums[4] + sums[5]) + (sums[6] + sums[7]));
    if (MODE == 0 || MODE == 2) tmem_wait_st();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l_run + m_run + __uint_as_float(s[5]);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int MODE> void run(const char* name, int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, warps * 32>>>(iters, out, cyc); cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(iters, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-22s warps/SM=%2d: %.0f cycles per warp-iteration (128 exps) -> %.0f per SMSP  %s\n", name, warps, (double)h[0] / iters,
         (double)h[0] / iters / 1.0, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 12}) { run<0>("full", w); run<1>("no STTM", w); run<2>("no LDTM", w); run<3>("exp only", w); }
  return 0;
}
