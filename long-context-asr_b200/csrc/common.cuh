// Shared helpers for the lcasr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>
#include "../../include/lcasr_b200.h"

namespace lcasr {

// ------------------------------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI
// ------------------------------------------------------------------------------------------------
char* last_error_buf();
int set_error(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define LCASR_CHECK_ARG(cond, ...)                                                    \
  do {                                                                                \
    if (!(cond)) return ::lcasr::set_error(LCASR_E_BADARG, __VA_ARGS__);              \
  } while (0)

#define LCASR_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      return ::lcasr::set_error(LCASR_E_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                                cudaGetErrorString(_e), __FILE__, __LINE__);          \
  } while (0)

#define LCASR_LAUNCH_CHECK()                                                          \
  do {                                                                                \
    ::lcasr::count_launch();                                                          \
    cudaError_t _e = cudaPeekAtLastError();                                           \
    if (_e != cudaSuccess)                                                            \
      return ::lcasr::set_error(LCASR_E_CUDA, "kernel launch failed: %s (%s:%d)",     \
                                cudaGetErrorString(_e), __FILE__, __LINE__);          \
  } while (0)

#define LCASR_TRY(expr)                                                               \
  do {                                                                                \
    int _s = (expr);                                                                  \
    if (_s != 0) return _s;                                                           \
  } while (0)

using bf16 = __nv_bfloat16;

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember it per (call site, device), not per process
struct PerDeviceFlag {
  std::atomic<bool> done[64] = {};
  bool needs_set(int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    return dev < 0 || dev >= 64 || !done[dev].load(std::memory_order_acquire);
  }
  void mark(int dev) { if (dev >= 0 && dev < 64) done[dev].store(true, std::memory_order_release); }
};

constexpr int kNumSMs = 148;

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
inline size_t dtype_size(int dt) { return dt == LCASR_BF16 ? 2 : 4; }

// shared-memory tiled depthwise Conv1d (dwconv_tile.cu): bf16, d % 128 == 0, k in {3,5,7,9,11,15}
bool dwconv1d_tile_ok(int d, int ks);
int dwconv1d_tile_fwd(const void* in, int B, int64_t N, int d, int ks, const float* w, const float* b, void* out, double* sum,
                      double* sumsq, cudaStream_t st);
int dwconv1d_tile_bwd_data(const void* dout, int B, int64_t N, int d, int ks, const float* w, void* din, cudaStream_t st);
int dwconv1d_tile_bwd_weight(const void* x, const void* dout, int B, int64_t N, int d, int ks, float* dw, float* db,
                             cudaStream_t st);
int dwconv1d_tile_eval(const void* in, int B, int64_t N, int d, int ks, const float* w, const float* b, const float* rm,
                       const float* rs, const float* bw, const float* bb, void* out, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// x * sigmoid(x); expf keeps fp32-mode parity at the 1e-6 level
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
// bf16-output variants: one MUFU (tanh.approx, rel. error 2^-11 — below bf16's 2^-9) instead of ex2 + rcp
__device__ __forceinline__ float tanh_approx_f(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float silu_fast(float x) {  // x*sigmoid(x) = h + h*tanh(h), h = x/2
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx_f(h), h);
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx_f(0.5f * x), 0.5f); }
template <typename TOut> __device__ __forceinline__ float silu_for(float x) {
  if constexpr (sizeof(TOut) == 2) return silu_fast(x); else return silu_f(x);
}
template <typename TOut> __device__ __forceinline__ float sigmoid_for(float x) {
  if constexpr (sizeof(TOut) == 2) return sigmoid_fast(x); else return sigmoid_f(x);
}
// F.gelu(x, approximate='tanh')  (fused_dense.py:466)
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float apply_act(float y, int act) {
  if (act == LCASR_ACT_GELU_TANH) return gelu_tanh_f(y);
  if (act == LCASR_ACT_SILU) return silu_f(y);
  return y;
}

// 8-element vector load/store of a row segment, converting to/from fp32
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

// V (4 or 8) consecutive bf16 <-> fp32: 8- or 16-byte accesses.  The narrower form halves the per-thread register
// footprint of the stencil kernels that keep per-channel weights in registers (more resident warps, more loads in flight).
template <int V> struct VecB;
template <> struct VecB<8> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) { Vec8<bf16>::load(p, v); }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) { Vec8<bf16>::store(p, v); }
};
template <> struct VecB<4> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

}  // namespace lcasr
