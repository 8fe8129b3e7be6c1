// Backward of the 8x depthwise-striding subsampling stencils (lcasr/components/subsampling.py:277-323), bf16
// channels-last activations, fp32 parameter gradients.  HBM-bound like their forward counterparts in subsample.cu:
//   dwconv data gradient : reads the [B,Tout,Fout,C] output gradient (each element up to 9/4 times, L2-resident
//                          neighbours), writes [B,Tin,Fin,C]
//   dwconv weight grad   : reads input + output gradient once, 10 fp32 atomics per channel and CTA
//   conv0 weight grad    : reads the spectrogram (x C/64 from L2) and the conv0 output gradient once; the conv0
//                          pre-activation is recomputed from the spectrogram instead of being stored
#include "common.cuh"
#include <cstdlib>

namespace lcasr {

__device__ __forceinline__ float silu_grad_sb(float x) {
  const float s = sigmoid_fast(x);
  return s * (1.0f + x * (1.0f - s));
}

// din[b,ti,fi,:] = sum_{i,j : ti = 2*to-1+i, fi = 2*fo-1+j} dout[b,to,fo,:] * w[:, i*3+j]
// Stride 2 makes the tap pattern depend only on the parity of (ti, fi), so a thread produces one 2x2 block of
// input positions (rows 2a, 2a+1; columns 2e, 2e+1) from the four output gradients G[a..a+1][e..e+1]:
//   din[2a  ][2e  ] = G00 w11                 din[2a  ][2e+1] = G00 w12 + G01 w10
//   din[2a+1][2e  ] = G00 w21 + G10 w01       din[2a+1][2e+1] = G00 w22 + G01 w20 + G10 w02 + G11 w00
// 4 unconditional 16-byte loads and 4 stores per thread (instead of 9 predicated gathers per output), one CTA per
// kBdTB row pairs of one recording, 32-bit index arithmetic.
constexpr int kBdTB = 2;
template <int V>  // channels per thread (8: 16-byte accesses; 4: half the registers, twice the resident warps)
__global__ void __launch_bounds__(256) subsample_dwconv_bwd_data_kernel(const bf16* __restrict__ dout, const float* __restrict__ w,
                                                                        int64_t Tin, int Fin, int C, int64_t Tout, int Fout,
                                                                        bf16* __restrict__ din) {
  const int cgroups = C / V;
  const int cg = threadIdx.x % cgroups;
  float wr[V][9];
#pragma unroll
  for (int c = 0; c < V; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = w[(cg * V + c) * 9 + k];
  const int64_t b = blockIdx.y;
  const int64_t a0 = (int64_t)blockIdx.x * kBdTB;
  const int Fh = (Fin + 1) / 2;            // column pairs
  const int per_row = Fh * cgroups;
  const bf16* doutb = dout + b * Tout * Fout * C;
  bf16* dinb = din + b * Tin * Fin * C;
  for (int idx = threadIdx.x; idx < kBdTB * per_row; idx += blockDim.x) {
    const int al = idx / per_row;
    const int e = (idx - al * per_row) / cgroups;
    const int64_t a = a0 + al;
    if (2 * a >= Tin) break;
    float g00[V], g01[V], g10[V], g11[V];
    const bool r1 = a + 1 < Tout, c1 = e + 1 < Fout, r0 = a < Tout, c0 = e < Fout;
    auto ld = [&](float (&g)[V], bool ok, int64_t to, int fo) {
      if (ok) VecB<V>::load(doutb + ((to * Fout + fo) * C + cg * V), g);
      else {
#pragma unroll
        for (int c = 0; c < V; ++c) g[c] = 0.f;
      }
    };
    ld(g00, r0 && c0, a, e); ld(g01, r0 && c1, a, e + 1); ld(g10, r1 && c0, a + 1, e); ld(g11, r1 && c1, a + 1, e + 1);
    float o[V];
    const int64_t ti = 2 * a;
    const int fi = 2 * e;
#pragma unroll
    for (int c = 0; c < V; ++c) o[c] = g00[c] * wr[c][4];
    VecB<V>::store(dinb + ((ti * Fin + fi) * C + cg * V), o);
    if (fi + 1 < Fin) {
#pragma unroll
      for (int c = 0; c < V; ++c) o[c] = fmaf(g00[c], wr[c][5], g01[c] * wr[c][3]);
      VecB<V>::store(dinb + ((ti * Fin + fi + 1) * C + cg * V), o);
    }
    if (ti + 1 < Tin) {
#pragma unroll
      for (int c = 0; c < V; ++c) o[c] = fmaf(g00[c], wr[c][7], g10[c] * wr[c][1]);
      VecB<V>::store(dinb + (((ti + 1) * Fin + fi) * C + cg * V), o);
      if (fi + 1 < Fin) {
#pragma unroll
        for (int c = 0; c < V; ++c)
          o[c] = fmaf(g00[c], wr[c][8], fmaf(g01[c], wr[c][6], fmaf(g10[c], wr[c][2], g11[c] * wr[c][0])));
        VecB<V>::store(dinb + (((ti + 1) * Fin + fi + 1) * C + cg * V), o);
      }
    }
  }
}

// dw[c, i*3+j] += sum dout[b,to,fo,c] * in[b,2to-1+i,2fo-1+j,c] ; db[c] += sum dout
// CTAs stride over blocks of kBwTB output rows (grid.x of them per recording); register accumulators, one
// shared-memory reduction and C*10 global atomics per CTA.
constexpr int kBwTB = 4;
template <int V>
__global__ void __launch_bounds__(256) subsample_dwconv_bwd_weight_kernel(const bf16* __restrict__ in, const bf16* __restrict__ dout,
                                                                          int64_t Tin, int Fin, int C, int64_t Tout, int Fout,
                                                                          float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float sacc[];  // [C][10]
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cgroups = C / V;
  const int cg = threadIdx.x % cgroups;
  float aw[V][9], ab[V];
#pragma unroll
  for (int c = 0; c < V; ++c) {
    ab[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) aw[c][k] = 0.f;
  }
  const int64_t b = blockIdx.y;
  const int per_row = Fout * cgroups;
  const bf16* inb = in + b * Tin * Fin * C;
  const bf16* doutb = dout + b * Tout * Fout * C;
  const int64_t nblk = (Tout + kBwTB - 1) / kBwTB;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t t0 = blk * kBwTB;
    for (int idx = threadIdx.x; idx < kBwTB * per_row; idx += blockDim.x) {
      const int tl = idx / per_row;
      const int fo = (idx - tl * per_row) / cgroups;
      const int64_t to = t0 + tl;
      if (to >= Tout) break;
      float g[V];
      VecB<V>::load(doutb + ((to * Fout + fo) * C + cg * V), g);
#pragma unroll
      for (int c = 0; c < V; ++c) ab[c] += g[c];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int64_t ti = 2 * to - 1 + i;
        if (ti < 0 || ti >= Tin) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int fi = 2 * fo - 1 + j;
          if (fi < 0 || fi >= Fin) continue;
          float v[V];
          VecB<V>::load(inb + ((ti * Fin + fi) * C + cg * V), v);
#pragma unroll
          for (int c = 0; c < V; ++c) aw[c][i * 3 + j] = fmaf(g[c], v[c], aw[c][i * 3 + j]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < V; ++c) {
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(&sacc[(cg * V + c) * 10 + k], aw[c][k]);
    atomicAdd(&sacc[(cg * V + c) * 10 + 9], ab[c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) {
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, sacc[i]);
    else atomicAdd(db + c, sacc[i]);
  }
}

// conv0 (1->C, 3x3, s2, p1) + SiLU: gradients of w0 [C,9] and b0 [C] from ds1 = dL/d(silu output) [B,T1,F1,C].
// Same tiling as subsample_conv0_kernel (16 output frames per tile, the input patch in shared memory, each thread
// owns 8 channels); a CTA walks several tiles and keeps its partial sums in registers, so the reduction costs one
// shared-memory pass and C*10 global atomics per CTA, not per tile.
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
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) sacc[i] = 0.f;
  const int cgroups = C / 8;
  const int pos_stride = blockDim.x / cgroups;
  const bool worker = (int)threadIdx.x < pos_stride * cgroups;
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
  const int64_t ntiles = (T1 + kC0bTT - 1) / kC0bTT;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t1_0 = tile * kC0bTT;
    const int64_t t_in0 = 2 * t1_0 - 1;
    __syncthreads();  // the previous tile's patch is no longer read
    for (int idx = threadIdx.x; idx < rows * FW; idx += blockDim.x) {
      const int f = idx / rows - 1;
      const int r = idx % rows;
      const int64_t t = t_in0 + r;
      float v = 0.f;
      if (f >= 0 && f < F && t >= 0 && t < T) v = spec[((int64_t)b * F + f) * T + t];
      s_in[r * FW + (f + 1)] = v;
    }
    __syncthreads();
    if (worker) {
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
    }
  }
  if (worker) {
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
  LCASR_CHECK_ARG(B <= 65535 && ceil_div(Tin, 2 * kBdTB) <= 0x7fffffff, "subsample_dwconv_bwd_data: grid too large");
  dim3 grid((unsigned)ceil_div(ceil_div(Tin, 2), kBdTB), (unsigned)B);
  if (C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0)
    subsample_dwconv_bwd_data_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, w, Tin, Fin, C, Tout, Fout, (bf16*)din);
  else
    subsample_dwconv_bwd_data_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, w, Tin, Fin, C, Tout, Fout, (bf16*)din);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_subsample_dwconv_bwd_weight(const void* in, const void* dout, int B, int64_t Tin, int Fin, int C, float* dw,
                                                 float* db, void* stream) {
  LCASR_CHECK_ARG(in && dout && dw && db && B > 0 && Tin > 0 && Fin > 0, "subsample_dwconv_bwd_weight: bad arguments");
  LCASR_CHECK_ARG(cgroups_ok(C) && C * 40 <= 48 * 1024, "subsample_dwconv_bwd_weight: C=%d unsupported", C);
  const int64_t Tout = (Tin - 1) / 2 + 1;
  const int Fout = (Fin - 1) / 2 + 1;
  LCASR_CHECK_ARG(B <= 65535, "subsample_dwconv_bwd_weight: batch too large");
  int64_t gx = ceil_div((int64_t)kNumSMs * 8, B);  // ~8 CTAs per SM in total
  if (gx > ceil_div(Tout, kBwTB)) gx = ceil_div(Tout, kBwTB);
  dim3 grid((unsigned)gx, (unsigned)B);
  static const bool narrow = getenv("LCASR_SUBSAMPLE_V4") != nullptr;  // measured slower (856 vs 766 us): kept for A/B runs
  if (narrow && C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0)
    subsample_dwconv_bwd_weight_kernel<4><<<grid, 256, (size_t)C * 40, (cudaStream_t)stream>>>(
        (const bf16*)in, (const bf16*)dout, Tin, Fin, C, Tout, Fout, dw, db);
  else
    subsample_dwconv_bwd_weight_kernel<8><<<grid, 256, (size_t)C * 40, (cudaStream_t)stream>>>(
        (const bf16*)in, (const bf16*)dout, Tin, Fin, C, Tout, Fout, dw, db);
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
  LCASR_CHECK_ARG(B <= 65535, "subsample_conv0_bwd: batch too large");
  int64_t gx = ceil_div((int64_t)kNumSMs * 4, B);  // ~4 CTAs per SM in total, each walking several tiles
  if (gx > ceil_div(T1, kC0bTT)) gx = ceil_div(T1, kC0bTT);
  dim3 grid((unsigned)gx, (unsigned)B);
  subsample_conv0_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(spec, w, b, (const bf16*)ds1, F, T, C, T1, F1, dw, db);
  LCASR_LAUNCH_CHECK();
  return 0;
}
