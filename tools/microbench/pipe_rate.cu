// Micro-benchmark: per-SM-sub-partition issue cost (cycles per warp instruction) of the instructions in the
// attention softmax loop: MUFU.EX2, F2FP.BF16 pack, FFMA, FADD, FMNMX3, and the real per-element mix.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(int iters, float* out, long long* cyc) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 0.001f + i;
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = i;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float mx[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]);
      if (MODE == 1 && (i & 1) == 0) { __nv_bfloat162 p = __floats2bfloat162_rn(v[i], v[i + 1]); pk[i / 2] ^= *reinterpret_cast<uint32_t*>(&p); }
      if (MODE == 2) v[i] = fmaf(v[i], 1.0001f, 0.5f);
      if (MODE == 3) acc[i & 7] += v[i];
      if (MODE == 4 && (i & 1) == 0) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(mx[(i / 2) & 3]) : "f"(v[i]), "f"(v[i + 1]));
      if (MODE == 5) {  // the softmax element: ffma, ex2, fadd, (pack every 2)
        float p = ex2(fmaf(v[i], 1.0001f, -0.5f));
        acc[i & 7] += p;
        v[i] = p;
        if (i & 1) { __nv_bfloat162 q = __floats2bfloat162_rn(v[i - 1], v[i]); pk[i / 2] ^= *reinterpret_cast<uint32_t*>(&q); }
      }
      if (MODE == 6) {  // same without the fadd
        float p = ex2(fmaf(v[i], 1.0001f, -0.5f));
        v[i] = p;
        if (i & 1) { __nv_bfloat162 q = __floats2bfloat162_rn(v[i - 1], v[i]); pk[i / 2] ^= *reinterpret_cast<uint32_t*>(&q); }
      }
      if (MODE == 7) {  // ffma + ex2 only
        v[i] = ex2(fmaf(v[i], 1.0001f, -0.5f));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += pk[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  s += mx[0] + mx[1] + mx[2] + mx[3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int warps, double ops_per_iter) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  k<MODE><<<148, warps * 32>>>(iters, out, cyc); cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(iters, out, cyc); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  // cycles per warp-instruction per sub-partition: (cycles) / (ops per warp * warps per SMSP)
  double per = (double)h[0] / (iters * ops_per_iter * (warps / 4.0));
  printf("%-26s warps/SM=%2d: %8lld cyc, %.2f cycles per warp-op per SMSP\n", name, warps, h[0], per);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("MUFU.EX2", w, 32); run<1>("F2FP.BF16 pack", w, 16); run<2>("FFMA", w, 32); run<3>("FADD (8 chains)", w, 32);
    run<4>("FMNMX3 (4 chains)", w, 16); run<5>("mix ffma+ex2+fadd+pack/2", w, 32); run<6>("mix ffma+ex2+pack/2", w, 32);
    run<7>("mix ffma+ex2", w, 32);
  }
  return 0;
}
