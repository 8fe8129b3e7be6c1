// Backward of the 8x depthwise-striding subsampling stencils (lcasr/components/subsampling.py:277-323), bf16
// channels-last activations, fp32 parameter gradients.  HBM-bound like their forward counterparts in subsample.cu:
//   dwconv data gradient : reads the [B,Tout,Fout,C] output gradient (each element up to 9/4 times, L2-resident
//                          neighbours), writes [B,Tin,Fin,C]
//   dwconv weight grad   : reads input + output gradient once, 10 fp32 atomics per channel and CTA
//   conv0 weight grad    : reads the spectrogram (x C/64 from L2) and the conv0 output gradient once; the conv0
//                          pre-activation is recomputed from the spectrogram instead of being stored
#include "common.cuh"

namespace lcasr {

__device__ __forceinline__ float silu_grad_sb(float x) {
  const float s = sigmoid_fast(x);
  return s * (1.0f + x * (1.0f - s));
}

// din[b,ti,fi,:] = sum_{i,j : ti = 2*to-1+i, fi = 2*fo-1+j} dout[b,to,fo,:] * w[:, i*3+j]
__global__ void __launch_bounds__(256) subsample_dwconv_bwd_data_kernel(const bf16* __restrict__ dout, const float* __restrict__ w,
                                                                        int64_t Tin, int Fin, int C, int64_t Tout, int Fout,
                                                                        int64_t total_vec, bf16* __restrict__ din) {
  const int cgroups = C / 8;
  const int cg = threadIdx.x % cgroups;
  float wr[8][9];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = w[(cg * 8 + c) * 9 + k];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_vec; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = idx / cgroups;
    const int fi = (int)(pos % Fin);
    const int64_t bt = pos / Fin;
    const int64_t ti = bt % Tin, b = bt / Tin;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int64_t t2 = ti + 1 - i;
      if (t2 < 0 || (t2 & 1)) continue;
      const int64_t to = t2 >> 1;
      if (to >= Tout) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int f2 = fi + 1 - j;
        if (f2 < 0 || (f2 & 1)) continue;
        const int fo = f2 >> 1;
        if (fo >= Fout) continue;
        float g[8];
        Vec8<bf16>::load(dout + (((b * Tout + to) * Fout + fo) * C + cg * 8), g);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(wr[c][i * 3 + j], g[c], acc[c]);
      }
    }
    Vec8<bf16>::store(din + (pos * C + cg * 8), acc);
  }
}

// dw[c, i*3+j] += sum dout[b,to,fo,c] * in[b,2to-1+i,2fo-1+j,c] ; db[c] += sum dout
__global__ void __launch_bounds__(256) subsample_dwconv_bwd_weight_kernel(const bf16* __restrict__ in, const bf16* __restrict__ dout,
                                                                          int64_t Tin, int Fin, int C, int64_t Tout, int Fout,
                                                                          int64_t total_vec, float* __restrict__ dw,
                                                                          float* __restrict__ db) {
  extern __shared__ float sacc[];  // [C][10]
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cgroups = C / 8;
  const int cg = threadIdx.x % cgroups;
  float aw[8][9], ab[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    ab[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) aw[c][k] = 0.f;
  }
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_vec; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = idx / cgroups;
    const int fo = (int)(pos % Fout);
    const int64_t bt = pos / Fout;
    const int64_t to = bt % Tout, b = bt / Tout;
    float g[8];
    Vec8<bf16>::load(dout + (pos * C + cg * 8), g);
#pragma unroll
    for (int c = 0; c < 8; ++c) ab[c] += g[c];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int64_t ti = 2 * to - 1 + i;
      if (ti < 0 || ti >= Tin) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int fi = 2 * fo - 1 + j;
        if (fi < 0 || fi >= Fin) continue;
        float v[8];
        Vec8<bf16>::load(in + (((b * Tin + ti) * Fin + fi) * C + cg * 8), v);
#pragma unroll
        for (int c = 0; c < 8; ++c) aw[c][i * 3 + j] = fmaf(g[c], v[c], aw[c][i * 3 + j]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(&sacc[(cg * 8 + c) * 10 + k], aw[c][k]);
    atomicAdd(&sacc[(cg * 8 + c) * 10 + 9], ab[c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) {
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, sacc[i]);
    else atomicAdd(db + c, sacc[i]);
  }
}

// conv0 (1->C, 3x3, s2, p1) + SiLU: gradients of w0 [C,9] and b0 [C] from ds1 = dL/d(silu output) [B,T1,F1,C].
// Same tiling as subsample_conv0_kernel: one (batch, 16 output frames) tile per CTA, the input patch in shared
// memory, each thread owns 8 channels.
constexpr int kC0bTT = 16;
__global__ void __launch_bounds__(256) subsample_conv0_bwd_kernel(const float* __restrict__ spec, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, const bf16* __restrict__ ds1,
                                                                  int F, int64_t T, int C, int64_t T1, int F1,
                                                                  float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float smem[];  // input patch [(2*TT+1)][F+2], then accumulators [C][10]
  const int FW = F + 2;
  const int rows = 2 * kC0bTT + 1;
  float* s_in = smem;
  float* sacc = smem + rows * FW;
  const int b = blockIdx.y;
  const int64_t t1_0 = (int64_t)blockIdx.x * kC0bTT;
  const int64_t t_in0 = 2 * t1_0 - 1;
  for (int idx = threadIdx.x; idx < rows * FW; idx += blockDim.x) {
    const int f = idx / rows - 1;
    const int r = idx % rows;
    const int64_t t = t_in0 + r;
    float v = 0.f;
    if (f >= 0 && f < F && t >= 0 && t < T) v = spec[((int64_t)b * F + f) * T + t];
    s_in[r * FW + (f + 1)] = v;
  }
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cgroups = C / 8;
  const int pos_stride = blockDim.x / cgroups;
  if ((int)threadIdx.x < pos_stride * cgroups) {
    const int cg = threadIdx.x % cgroups;
    const int pos_lane = threadIdx.x / cgroups;
    float wr[8][9], br[8], aw[8][9], ab[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      br[c] = bias[cg * 8 + c];
      ab[c] = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) { wr[c][k] = w[(cg * 8 + c) * 9 + k]; aw[c][k] = 0.f; }
    }
    const int npos = kC0bTT * F1;
    for (int p = pos_lane; p < npos; p += pos_stride) {
      const int tt = p / F1, f1 = p % F1;
      const int64_t t1 = t1_0 + tt;
      if (t1 >= T1) break;
      float in[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) in[i * 3 + j] = s_in[(2 * tt + i) * FW + (2 * f1 + j)];
      float g[8];
      Vec8<bf16>::load(ds1 + ((((int64_t)b * T1 + t1) * F1 + f1) * C + cg * 8), g);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float a = br[c];
#pragma unroll
        for (int k = 0; k < 9; ++k) a = fmaf(wr[c][k], in[k], a);
        const float ga = g[c] * silu_grad_sb(a);
        ab[c] += ga;
#pragma unroll
        for (int k = 0; k < 9; ++k) aw[c][k] = fmaf(ga, in[k], aw[c][k]);
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(&sacc[(cg * 8 + c) * 10 + k], aw[c][k]);
      atomicAdd(&sacc[(cg * 8 + c) * 10 + 9], ab[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) {
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, sacc[i]);
    else atomicAdd(db + c, sacc[i]);
  }
}

}  // namespace lcasr

using namespace lcasr;

static bool cgroups_ok(int C) { return C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0; }

extern "C" int lcasr_subsample_dwconv_bwd_data(const void* dout, const float* w, int B, int64_t Tin, int Fin, int C, void* din,
                                               void* stream) {
  LCASR_CHECK_ARG(dout && w && din && B > 0 && Tin > 0 && Fin > 0, "subsample_dwconv_bwd_data: bad arguments");
  LCASR_CHECK_ARG(cgroups_ok(C), "subsample_dwconv_bwd_data: C=%d: C/8 must divide 256", C);
  const int64_t Tout = (Tin - 1) / 2 + 1;
  const int Fout = (Fin - 1) / 2 + 1;
  const int64_t total = (int64_t)B * Tin * Fin * (C / 8);
  int64_t g = ceil_div(total, 256);
  if (g > (int64_t)kNumSMs * 32) g = (int64_t)kNumSMs * 32;
  subsample_dwconv_bwd_data_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, w, Tin, Fin, C, Tout, Fout,
                                                                                 total, (bf16*)din);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_subsample_dwconv_bwd_weight(const void* in, const void* dout, int B, int64_t Tin, int Fin, int C, float* dw,
                                                 float* db, void* stream) {
  LCASR_CHECK_ARG(in && dout && dw && db && B > 0 && Tin > 0 && Fin > 0, "subsample_dwconv_bwd_weight: bad arguments");
  LCASR_CHECK_ARG(cgroups_ok(C) && C * 40 <= 48 * 1024, "subsample_dwconv_bwd_weight: C=%d unsupported", C);
  const int64_t Tout = (Tin - 1) / 2 + 1;
  const int Fout = (Fin - 1) / 2 + 1;
  const int64_t total = (int64_t)B * Tout * Fout * (C / 8);
  int64_t g = ceil_div(total, 256);
  if (g > (int64_t)kNumSMs * 4) g = (int64_t)kNumSMs * 4;
  subsample_dwconv_bwd_weight_kernel<<<(unsigned)g, 256, (size_t)C * 40, (cudaStream_t)stream>>>(
      (const bf16*)in, (const bf16*)dout, Tin, Fin, C, Tout, Fout, total, dw, db);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_subsample_conv0_bwd(const float* spec, const float* w, const float* b, const void* ds1, int B, int F,
                                         int64_t T, int C, float* dw, float* db, void* stream) {
  LCASR_CHECK_ARG(spec && w && b && ds1 && dw && db && B > 0 && F > 0 && T > 0, "subsample_conv0_bwd: bad arguments");
  LCASR_CHECK_ARG(cgroups_ok(C), "subsample_conv0_bwd: C=%d: C/8 must divide 256", C);
  const int64_t T1 = (T - 1) / 2 + 1;
  const int F1 = (F - 1) / 2 + 1;
  const size_t smem = ((size_t)(2 * kC0bTT + 1) * (F + 2) + (size_t)C * 10) * sizeof(float);
  LCASR_CHECK_ARG(smem <= 48 * 1024, "subsample_conv0_bwd: F=%d, C=%d need too much shared memory", F, C);
  dim3 grid((unsigned)ceil_div(T1, kC0bTT), (unsigned)B);
  subsample_conv0_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(spec, w, b, (const bf16*)ds1, F, T, C, T1, F1, dw, db);
  LCASR_LAUNCH_CHECK();
  return 0;
}
