// 8x depthwise-striding subsampling stencils (lcasr/components/subsampling.py:277-323).
// Both kernels are HBM-bound and write channels-last so the 1x1 convolutions that follow are
// plain [rows, C] x [C, C] GEMMs for the tcgen05 kernel.
//   conv0 : reads B*F*T*4 bytes, writes B*T1*F1*C*e bytes  (the activation is 160x the input)
//   dwconv: reads B*Tin*Fin*C*e, writes B*Tout*Fout*C*e
#include "common.cuh"

namespace lcasr {

// ---- conv0: Conv2d(1->C,3x3,s2,p1) + SiLU ------------------------------------------------------
// Block = one (batch, tile of TT output frames); the (2*TT+1) x (F+2) input patch is staged in
// shared memory with its zero border; each thread owns 8 adjacent channels (weights in registers)
// and walks output positions, so a warp writes 32*8 contiguous channels (512 B in bf16) per store.
constexpr int kConv0TT = 16;

template <typename TOut>
__global__ void __launch_bounds__(256) subsample_conv0_kernel(const float* __restrict__ spec, const float* __restrict__ w,
                                                              const float* __restrict__ bias, int F, int64_t T, int C,
                                                              int64_t T1, int F1, TOut* __restrict__ out) {
  extern __shared__ float s_in[];  // [(2*TT+1)][F+2]  (time-major rows, freq fastest; col 0 = freq -1)
  const int FW = F + 2;
  const int b = blockIdx.y;
  const int64_t t1_0 = (int64_t)blockIdx.x * kConv0TT;
  const int64_t t_in0 = 2 * t1_0 - 1;
  const int rows = 2 * kConv0TT + 1;
  // cooperative, time-coalesced load: spec[b][f][t]
  for (int idx = threadIdx.x; idx < rows * FW; idx += blockDim.x) {
    int f = idx / rows - 1;  // -1 .. F
    int r = idx % rows;
    int64_t t = t_in0 + r;
    float v = 0.f;
    if (f >= 0 && f < F && t >= 0 && t < T) v = spec[((int64_t)b * F + f) * T + t];
    s_in[r * FW + (f + 1)] = v;
  }
  __syncthreads();
  const int cgroups = C / 8;
  const int cg = threadIdx.x % cgroups;
  const int pos_lane = threadIdx.x / cgroups;
  const int pos_stride = blockDim.x / cgroups;
  float wr[8][9], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = bias[cg * 8 + c];
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = w[(cg * 8 + c) * 9 + k];
  }
  const int npos = kConv0TT * F1;
  for (int p = pos_lane; p < npos; p += pos_stride) {
    int tt = p / F1, f1 = p % F1;
    int64_t t1 = t1_0 + tt;
    if (t1 >= T1) break;
    float in[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) in[i * 3 + j] = s_in[(2 * tt + i) * FW + (2 * f1 + j)];
    float y[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = br[c];
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(wr[c][k], in[k], a);
      y[c] = silu_for<TOut>(a);
    }
    TOut* o = out + (((int64_t)b * T1 + t1) * F1 + f1) * C + cg * 8;
    Vec8<TOut>::store(o, y);
  }
}

// ---- depthwise Conv2d(C,3x3,s2,p1,groups=C), channels-last ------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) subsample_dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                               const float* __restrict__ bias, int64_t Tin, int Fin,
                                                               int C, int64_t Tout, int Fout, int64_t total_vec,
                                                               T* __restrict__ out) {
  const int cgroups = C / 8;             // divides blockDim.x, so a thread's channel group is fixed
  const int cg = threadIdx.x % cgroups;
  float wr[8][9], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = bias[cg * 8 + c];
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = w[(cg * 8 + c) * 9 + k];
  }
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_vec;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t pos = idx / cgroups;
    int fo = (int)(pos % Fout);
    int64_t bt = pos / Fout;
    int64_t to = bt % Tout;
    int64_t b = bt / Tout;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = br[c];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      int64_t ti = 2 * to - 1 + i;
      if (ti < 0 || ti >= Tin) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        int fi = 2 * fo - 1 + j;
        if (fi < 0 || fi >= Fin) continue;
        float v[8];
        Vec8<T>::load(in + (((b * Tin + ti) * Fin + fi) * C + cg * 8), v);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(wr[c][i * 3 + j], v[c], acc[c]);
      }
    }
    Vec8<T>::store(out + (pos * C + cg * 8), acc);
  }
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_subsample_conv0(const float* spec, const float* w, const float* b, int B, int F, int64_t T, int C,
                                     void* out, int out_dtype, void* stream) {
  LCASR_CHECK_ARG(spec && w && b && out && B > 0 && F > 0 && T > 0, "subsample_conv0: bad arguments");
  LCASR_CHECK_ARG(C % 8 == 0 && C >= 8 && (C / 8) <= 256 && 256 % (C / 8) == 0,
                  "subsample_conv0: conv_channels=%d must be a multiple of 8 with 256 %% (C/8) == 0", C);
  const int64_t T1 = (T - 1) / 2 + 1;
  const int F1 = (F - 1) / 2 + 1;
  dim3 grid((unsigned)ceil_div(T1, kConv0TT), B), block(256);
  size_t smem = (size_t)(2 * kConv0TT + 1) * (F + 2) * sizeof(float);
  LCASR_CHECK_ARG(smem <= 48 * 1024, "subsample_conv0: feat_in=%d too large for the staged patch", F);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == LCASR_BF16)
    subsample_conv0_kernel<bf16><<<grid, block, smem, st>>>(spec, w, b, F, T, C, T1, F1, (bf16*)out);
  else
    subsample_conv0_kernel<float><<<grid, block, smem, st>>>(spec, w, b, F, T, C, T1, F1, (float*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_subsample_dwconv(const void* in, int dtype, const float* w, const float* b, int B, int64_t Tin,
                                      int Fin, int C, void* out, void* stream) {
  LCASR_CHECK_ARG(in && w && b && out && B > 0 && Tin > 0 && Fin > 0, "subsample_dwconv: bad arguments");
  LCASR_CHECK_ARG(C % 8 == 0 && 256 % (C / 8) == 0, "subsample_dwconv: C=%d must be a multiple of 8 with 256 %% (C/8) == 0", C);
  const int64_t Tout = (Tin - 1) / 2 + 1;
  const int Fout = (Fin - 1) / 2 + 1;
  const int64_t total_vec = (int64_t)B * Tout * Fout * (C / 8);
  int64_t blocks = ceil_div(total_vec, 256);
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LCASR_BF16)
    subsample_dwconv_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)in, w, b, Tin, Fin, C, Tout, Fout,
                                                                     total_vec, (bf16*)out);
  else
    subsample_dwconv_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)in, w, b, Tin, Fin, C, Tout, Fout,
                                                                      total_vec, (float*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}
