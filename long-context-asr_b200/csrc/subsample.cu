// 8x depthwise-striding subsampling stencils (lcasr/components/subsampling.py:277-323).
// Both kernels are HBM-bound and write channels-last so the 1x1 convolutions that follow are
// plain [rows, C] x [C, C] GEMMs for the tcgen05 kernel.
//   conv0 : reads B*F*T*4 bytes, writes B*T1*F1*C*e bytes  (the activation is 160x the input)
//   dwconv: reads B*Tin*Fin*C*e, writes B*Tout*Fout*C*e
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <cstdlib>

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
// One CTA = kDwTB output time rows of one recording; all index arithmetic is 32-bit (the flat 64-bit
// div/mod chain of a grid-stride loop is avoided; measured 775 -> 685 us for the level-1 tensor of cfg 5).
constexpr int kDwTB = 4;
template <typename T, int V> struct VecT;
template <> struct VecT<float, 8> : Vec8<float> {};
template <int V> struct VecT<bf16, V> : VecB<V> {};
template <typename T, int V>  // V channels per thread (bf16: 4 halves the register footprint -> more loads in flight)
__global__ void __launch_bounds__(256) subsample_dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                               const float* __restrict__ bias, int64_t Tin, int Fin,
                                                               int C, int64_t Tout, int Fout, T* __restrict__ out) {
  const int cgroups = C / V;             // divides blockDim.x, so a thread's channel group is fixed
  const int cg = threadIdx.x % cgroups;
  float wr[V][9], br[V];
#pragma unroll
  for (int c = 0; c < V; ++c) {
    br[c] = bias[cg * V + c];
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = w[(cg * V + c) * 9 + k];
  }
  const int64_t b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * kDwTB;
  const int per_row = Fout * cgroups;
  const T* inb = in + b * Tin * Fin * C;
  T* outb = out + b * Tout * Fout * C;
  for (int i = threadIdx.x; i < kDwTB * per_row; i += blockDim.x) {
    const int tl = i / per_row;
    const int fo = (i - tl * per_row) / cgroups;
    const int64_t to = t0 + tl;
    if (to >= Tout) break;
    float acc[V];
#pragma unroll
    for (int c = 0; c < V; ++c) acc[c] = br[c];
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
      const int64_t ti = 2 * to - 1 + ii;
      if (ti < 0 || ti >= Tin) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int fi = 2 * fo - 1 + j;
        if (fi < 0 || fi >= Fin) continue;
        float v[V];
        VecT<T, V>::load(inb + ((ti * Fin + fi) * C + cg * V), v);
#pragma unroll
        for (int c = 0; c < V; ++c) acc[c] = fmaf(wr[c][ii * 3 + j], v[c], acc[c]);
      }
    }
    VecT<T, V>::store(outb + ((to * Fout + fo) * C + cg * V), acc);
  }
}

// ---- fused conv0 + SiLU + first depthwise level (bf16 path) -------------------------------------
// The 1->C conv0 activation is 160x the input (1.3 GB in bf16 for a 20-minute recording): writing it
// and reading it back 2.25x for the stride-2 depthwise conv is what made the two separate kernels
// HBM-bound.  Here a CTA owns (batch, 8 output rows of the depthwise level, 64 channels): it stages the
// 35 x (F+2) spectrogram patch, computes the 17 x (F1+2) x 64 conv0+SiLU tile into shared memory as bf16
// (zeros where the depthwise conv pads), and reduces it to the 8 x F2 x 64 depthwise outputs.
// HBM traffic drops to: input (re-read by the 4-8 channel groups, L2-resident) + the depthwise output.
constexpr int kFuCG = 64;
constexpr int kSubTcDefault = 1;  // conv0 on the tensor cores: 0.985 vs 1.020 ms (cfg 3), 1.87 vs 1.97 (cfg 2), 5.33 vs 5.54 (cfg 4)

template <int kFuTT2, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) subsample_conv0_dw_fused_kernel(
    const float* __restrict__ spec, const float* __restrict__ w0, const float* __restrict__ b0,
    const float* __restrict__ w1, const float* __restrict__ b1, int F, int64_t T, int C, int64_t T1, int F1,
    int64_t T2, int F2, bf16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t fsm[];
  constexpr int A0R = 2 * kFuTT2 + 1;           // conv0 rows needed: 17
  constexpr int INR = 2 * A0R + 1;              // input frames needed: 35
  const int FW = max(F + 2, 4 * ((F1 + 1) / 2) + 1), A0W = F1 + 2;  // patch pitch: every 5-wide block read stays inside its row
  float* s_in = reinterpret_cast<float*>(fsm);                                   // [INR][FW], col 0 = freq -1, zero beyond F
  __nv_bfloat162* s_a0 = reinterpret_cast<__nv_bfloat162*>(fsm + ((INR * FW * 4 + 15) & ~15));  // [A0R][A0W][32]
  const int cgi = blockIdx.y, b = blockIdx.z;
  const int64_t t2_0 = (int64_t)blockIdx.x * kFuTT2;
  const int64_t a0_row0 = 2 * t2_0 - 1;         // global conv0 row of tile row 0
  const int64_t t_in0 = 2 * a0_row0 - 1;        // global input frame of patch row 0
  for (int idx = threadIdx.x; idx < INR * FW; idx += blockDim.x) {
    const int f = idx / INR - 1, r = idx % INR;
    const int64_t t = t_in0 + r;
    float v = 0.f;
    if (f >= 0 && f < F && t >= 0 && t < T) v = spec[((int64_t)b * F + f) * T + t];
    s_in[r * FW + (f + 1)] = v;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;   // lane -> channel pair, warp -> positions
  const int c0 = cgi * kFuCG + 2 * lane;
  // the lane's two channels as packed fp32 pairs: one FFMA2 per tap for both (the kernel is issue-bound on FFMA)
  uint64_t wa2[9], wd2[9];
  const uint64_t ba2 = ptx::pack2f(b0[c0], b0[c0 + 1]), bd2 = ptx::pack2f(b1[c0], b1[c0 + 1]);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    wa2[k] = ptx::pack2f(w0[c0 * 9 + k], w0[(c0 + 1) * 9 + k]);
    wd2[k] = ptx::pack2f(w1[c0 * 9 + k], w1[(c0 + 1) * 9 + k]);
  }
  __syncthreads();
  // phase 1: conv0 + SiLU tile (bf16x2 per lane), zero where the depthwise conv sees padding.
  // A warp owns one 2x2 block of conv0 positions (global rows 2a, 2a+1; columns 2e, 2e+1): the four positions share a
  // 5x5 input patch (25 shared-memory reads instead of 36) and one round of index arithmetic; blocks advance
  // incrementally (no div/mod).  Tile row 0 is a global ODD row: it is the lower half of block row 0, whose upper half
  // lies above the tile and is skipped (warp-uniform).
  {
    const int r_lo = (int)max((int64_t)0, -a0_row0);                       // tile rows below are conv0 row < 0
    const int r_hi = (int)min((int64_t)A0R, T1 - a0_row0);                 // tile rows from here on are >= T1
    const int EB = (F1 + 1) / 2;                                           // column blocks
    constexpr int NBR = kFuTT2 + 1;                                        // block rows: tile rows (2bi-1, 2bi)
    for (int i = threadIdx.x; i < A0R * 2 * 32; i += blockDim.x) {        // the two zero-padding columns of the tile
      const int r = i / 64, side = (i >> 5) & 1;
      s_a0[(r * A0W + (side ? A0W - 1 : 0)) * 32 + (i & 31)] = __floats2bfloat162_rn(0.f, 0.f);
    }
    int bi = 0, e = wid;
    while (e >= EB) { e -= EB; ++bi; }
#pragma unroll 1
    while (bi < NBR) {
      float in[5][5];
      const int prow0 = 4 * bi - 2;                                        // patch row of in[0] (>= 0 except for bi == 0)
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) in[i][j] = s_in[max(prow0 + i, 0) * FW + 4 * e + j];
#pragma unroll
      for (int pi = 0; pi < 2; ++pi) {
        const int r = 2 * bi - 1 + pi;                                     // tile row
        if (r < 0) continue;                                               // warp-uniform: upper half of block row 0
        const bool row_live = r >= r_lo && r < r_hi;
#pragma unroll
        for (int pj = 0; pj < 2; ++pj) {
          const int c = 2 * e + 1 + pj;                                    // tile column (conv0 column c - 1)
          uint64_t acc2 = ba2;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) ptx::ffma2_acc_bcast(acc2, wa2[i * 3 + j], in[2 * pi + i][2 * pj + j]);
          float a, bq;
          ptx::unpack2f(acc2, a, bq);
          const bool live = row_live && c <= F1;
          const float y0 = live ? silu_fast(a) : 0.f, y1 = live ? silu_fast(bq) : 0.f;
          if (c < A0W) s_a0[(r * A0W + c) * 32 + lane] = __floats2bfloat162_rn(y0, y1);
        }
      }
      e += 8;
      while (e >= EB) { e -= EB; ++bi; }
    }
  }
  __syncthreads();
  // phase 2: depthwise 3x3 stride 2 over the tile
  {
    const int tl_hi = (int)min((int64_t)kFuTT2, T2 - t2_0);
    int tl = 0, f2 = wid;
    while (f2 >= F2) { f2 -= F2; ++tl; }                                    // F2 may be < 8
    bf16* obase = out + (((int64_t)b * T2 + t2_0) * F2) * C + c0;
#pragma unroll 2
    for (int q = wid; q < kFuTT2 * F2; q += 8) {
      if (tl >= tl_hi) break;
      const __nv_bfloat162* a0p = s_a0 + ((2 * tl) * A0W + 2 * f2) * 32 + lane;
      uint64_t acc2 = bd2;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float2 v = __bfloat1622float2(a0p[(i * A0W + j) * 32]);
          ptx::ffma2_acc(acc2, wd2[i * 3 + j], ptx::pack2f(v.x, v.y));
        }
      float a, bq;
      ptx::unpack2f(acc2, a, bq);
      *reinterpret_cast<__nv_bfloat162*>(obase + (size_t)q * C) = __floats2bfloat162_rn(a, bq);
      f2 += 8;
      while (f2 >= F2) { f2 -= F2; ++tl; }
    }
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
  LCASR_CHECK_ARG(B <= 65535 && ceil_div(Tout, kDwTB) <= 0x7fffffff, "subsample_dwconv: grid too large");
  dim3 grid((unsigned)ceil_div(Tout, kDwTB), (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool narrow = getenv("LCASR_SUBSAMPLE_V4") != nullptr;  // measured slower (801 vs 685 us): kept for A/B runs
  if (narrow && dtype == LCASR_BF16 && C / 4 <= 256 && 256 % (C / 4) == 0)
    subsample_dwconv_kernel<bf16, 4><<<grid, 256, 0, st>>>((const bf16*)in, w, b, Tin, Fin, C, Tout, Fout, (bf16*)out);
  else if (dtype == LCASR_BF16)
    subsample_dwconv_kernel<bf16, 8><<<grid, 256, 0, st>>>((const bf16*)in, w, b, Tin, Fin, C, Tout, Fout, (bf16*)out);
  else
    subsample_dwconv_kernel<float, 8><<<grid, 256, 0, st>>>((const float*)in, w, b, Tin, Fin, C, Tout, Fout, (float*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

template <int TT2, int MINB>
static int launch_fused(const float* spec, const float* w0, const float* b0, const float* w1, const float* b1, int B, int F,
                        int64_t T, int C, void* out, cudaStream_t st) {
  const int64_t T1 = (T - 1) / 2 + 1, T2 = (T1 - 1) / 2 + 1;
  const int F1 = (F - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  const int FWp = std::max(F + 2, 4 * ((F1 + 1) / 2) + 1);
  const size_t smem = (((size_t)(2 * (2 * TT2 + 1) + 1) * FWp * 4 + 15) & ~(size_t)15) + (size_t)(2 * TT2 + 1) * (F1 + 2) * 32 * 4;
  LCASR_CHECK_ARG(smem <= 110 * 1024, "subsample_conv0_dw: feat_in=%d too large for the fused tile", F);
  LCASR_CHECK_ARG(ceil_div(T2, TT2) <= 0x7fffffff && B <= 65535, "subsample_conv0_dw: grid too large");
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(subsample_conv0_dw_fused_kernel<TT2, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    attr_set.mark(attr_dev);
  }
  dim3 grid((unsigned)ceil_div(T2, TT2), (unsigned)(C / kFuCG), (unsigned)B);
  subsample_conv0_dw_fused_kernel<TT2, MINB><<<grid, 256, smem, st>>>(spec, w0, b0, w1, b1, F, T, C, T1, F1, T2, F2, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

namespace lcasr {
int subsample_conv0_dw_tc_launch(const float* spec, const float* w0, const float* b0, const float* w1, const float* b1, int B, int F,
                                 int64_t T, int C, void* out, cudaStream_t st);
}

extern "C" int lcasr_subsample_conv0_dw(const float* spec, const float* w0, const float* b0, const float* w1,
                                        const float* b1, int B, int F, int64_t T, int C, void* out, void* stream) {
  LCASR_CHECK_ARG(spec && w0 && b0 && w1 && b1 && out && B > 0 && F > 0 && T > 0, "subsample_conv0_dw: bad arguments");
  LCASR_CHECK_ARG(C % kFuCG == 0, "subsample_conv0_dw: conv_channels=%d must be a multiple of %d (use the unfused kernels)", C, kFuCG);
  // conv0 on the tensor cores (subsample_tc.cu); LCASR_SUB_TC=0 selects the SIMT kernel below (A/B runs), which also serves
  // feature axes too wide for the tensor-core tile
  static const int use_tc = getenv("LCASR_SUB_TC") ? atoi(getenv("LCASR_SUB_TC")) : kSubTcDefault;
  if (use_tc) {
    const int st_tc = subsample_conv0_dw_tc_launch(spec, w0, b0, w1, b1, B, F, T, C, out, (cudaStream_t)stream);
    if (st_tc != LCASR_E_UNSUPPORTED) return st_tc;
  }
  static const int tt2 = getenv("LCASR_SUB_TT2") ? atoi(getenv("LCASR_SUB_TT2")) : 4;  // tuning knob: depthwise rows per CTA (4: 1.48 ms, 8: 1.63 ms at cfg3)
  if (tt2 == 4) return launch_fused<4, 3>(spec, w0, b0, w1, b1, B, F, T, C, out, (cudaStream_t)stream);
  return launch_fused<8, 2>(spec, w0, b0, w1, b1, B, F, T, C, out, (cudaStream_t)stream);
}
