// Conformer convolution module, memory-bound middle part (convolution.py:112-121):
//   depthwise Conv1d(k=9, pad 4, groups=d) + bias -> BatchRenorm1d (eval: per-channel affine,
//   batchrenorm.py:86-91, no eps) -> SiLU,  channels-last [B,N,d].
// HBM-bound: algorithmic bytes = B*N*d*(e_in + e_out).  Each thread owns 8 channels and slides a
// register window over TT consecutive tokens, so every input vector is loaded once per block
// (+ (k-1)/TT halo re-reads that hit L2).
#include "common.cuh"
#include <cstdlib>

namespace lcasr {

constexpr int kDwTT = 32;     // tokens per thread
constexpr int kDwMaxK = 15;   // kernel sizes up to 15 keep the window in registers

template <typename TIn, typename TOut, int KS>
__global__ void __launch_bounds__(128) dwconv_brn_silu_kernel(const TIn* __restrict__ in, int64_t N, int d,
                                                              const float* __restrict__ w, const float* __restrict__ b,
                                                              const float* __restrict__ rm, const float* __restrict__ rs,
                                                              const float* __restrict__ bw, const float* __restrict__ bb,
                                                              TOut* __restrict__ out) {
  constexpr int PAD = (KS - 1) / 2;
  const int cgroups = d / 8;
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg >= cgroups) return;
  const int64_t n0 = (int64_t)blockIdx.y * kDwTT;
  const int64_t batch = blockIdx.z;
  const int c0 = cg * 8;
  float wr[8][KS], scale[8], shift[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int k = 0; k < KS; ++k) wr[c][k] = w[(c0 + c) * KS + k];
    // ((acc + b) - mean) / std * bw + bb  ==  acc*scale + shift
    float sc = bw[c0 + c] / rs[c0 + c];
    scale[c] = sc;
    shift[c] = (b[c0 + c] - rm[c0 + c]) * sc + bb[c0 + c];
  }
  const TIn* base = in + batch * N * d + c0;
  float win[KS][8];  // win[j] = x[n - PAD + j]
#pragma unroll
  for (int j = 0; j < KS - 1; ++j) {
    int64_t n = n0 - PAD + j;
    if (n >= 0 && n < N) Vec8<TIn>::load(base + n * d, win[j + 1]);
    else {
#pragma unroll
      for (int c = 0; c < 8; ++c) win[j + 1][c] = 0.f;
    }
  }
#pragma unroll 4
  for (int t = 0; t < kDwTT; ++t) {
    const int64_t n = n0 + t;
    if (n >= N) break;
#pragma unroll
    for (int j = 0; j < KS - 1; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) win[j][c] = win[j + 1][c];
    const int64_t nn = n + PAD;
    if (nn < N) Vec8<TIn>::load(base + nn * d, win[KS - 1]);
    else {
#pragma unroll
      for (int c = 0; c < 8; ++c) win[KS - 1][c] = 0.f;
    }
    float y[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < KS; ++k) a = fmaf(wr[c][k], win[k][c], a);
      y[c] = silu_for<TOut>(fmaf(a, scale[c], shift[c]));
    }
    Vec8<TOut>::store(out + (batch * N + n) * d + c0, y);
  }
}

template <typename TIn, typename TOut>
static int launch_dwconv(const void* in, int B, int64_t N, int d, int ks, const float* w, const float* b,
                         const float* rm, const float* rs, const float* bw, const float* bb, void* out,
                         cudaStream_t st) {
  const int cgroups = d / 8;
  const int threads = cgroups < 128 ? ((cgroups + 31) / 32) * 32 : 128;
  dim3 grid((unsigned)ceil_div(cgroups, threads), (unsigned)ceil_div(N, kDwTT), (unsigned)B), block(threads);
#define LCASR_DW_CASE(KS)                                                                                  \
  case KS:                                                                                                 \
    dwconv_brn_silu_kernel<TIn, TOut, KS><<<grid, block, 0, st>>>((const TIn*)in, N, d, w, b, rm, rs, bw, bb, \
                                                                  (TOut*)out);                             \
    break;
  switch (ks) {
    LCASR_DW_CASE(3) LCASR_DW_CASE(5) LCASR_DW_CASE(7) LCASR_DW_CASE(9) LCASR_DW_CASE(11) LCASR_DW_CASE(15)
    default:
      return set_error(LCASR_E_UNSUPPORTED, "dwconv_brn_silu: conv_kernel_size=%d not in {3,5,7,9,11,15}", ks);
  }
#undef LCASR_DW_CASE
  LCASR_LAUNCH_CHECK();
  return 0;
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_dwconv_brn_silu(const void* in, int dtype, int B, int64_t N, int d, int ksize, const float* w,
                                     const float* b, const float* brn_mean, const float* brn_std, const float* brn_w,
                                     const float* brn_b, void* out, int out_dtype, void* stream) {
  LCASR_CHECK_ARG(in && out && w && b && brn_mean && brn_std && brn_w && brn_b, "dwconv_brn_silu: NULL argument");
  LCASR_CHECK_ARG(B > 0 && N > 0 && d > 0 && d % 8 == 0, "dwconv_brn_silu: bad shape (d=%d must be a multiple of 8)", d);
  LCASR_CHECK_ARG(dtype == out_dtype, "dwconv_brn_silu: in/out dtypes must match");
  LCASR_CHECK_ARG((int64_t)ceil_div(N, kDwTT) <= 65535, "dwconv_brn_silu: N=%lld too long for the grid", (long long)N);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool legacy = getenv("LCASR_DWCONV_LEGACY") != nullptr;  // A/B switch: the register-window kernel
  if (dtype == LCASR_BF16 && !legacy && dwconv1d_tile_ok(d, ksize))
    return dwconv1d_tile_eval(in, B, N, d, ksize, w, b, brn_mean, brn_std, brn_w, brn_b, out, st);
  if (dtype == LCASR_BF16)
    return launch_dwconv<bf16, bf16>(in, B, N, d, ksize, w, b, brn_mean, brn_std, brn_w, brn_b, out, st);
  return launch_dwconv<float, float>(in, B, N, d, ksize, w, b, brn_mean, brn_std, brn_w, brn_b, out, st);
}
