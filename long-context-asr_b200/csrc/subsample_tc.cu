// Fused conv0 + SiLU + first depthwise level with conv0 ON THE TENSOR CORES (lcasr/components/subsampling.py:277-296:
// Conv2d(1->C, 3x3, s2, p1), SiLU, depthwise Conv2d(C, 3x3, s2, p1); run at :418).
//
// conv0 is 80 % of the subsampling FLOPs and it is a contraction: out[pos, c] = sum_tap w[c, tap] * patch[pos, tap] — an
// [positions, 9] x [9, C] product.  The SIMT kernel (subsample.cu) spends 36 FFMA2 + 25 shared-memory loads per 2x2 block of
// positions and channel pair on it and is issue-bound.  Here a CTA builds the im2col operand in shared memory and lets
// tcgen05 do the product:
//   K = 32 columns per position:  [ x_hi(9) | x_lo(9) | x_hi(9) | 0(5) ]        x = x_hi + x_lo (bf16 split of the fp32 input)
//   weights, per channel:         [ w_hi(9) | w_hi(9) | w_lo(9) | 0(5) ]        w = w_hi + w_lo
//   => acc = w_hi x_hi + w_hi x_lo + w_lo x_hi  (fp32 accumulate): the product to ~2^-16 relative — the bf16 path keeps its
//      error budget (the SIMT kernel multiplies fp32 by fp32), for 2 MMA k-steps instead of 1.
// CTA = (recording, 4 rows of the depthwise output, 64 channels): 9 conv0 rows x F1 columns = 360 positions (3 M=128 tiles,
// N = 64: 192 TMEM columns), operands in the canonical no-swizzle K-major layout (8-row x 16-byte core matrices), written by
// the threads (no TMA: the operand does not exist in global memory).  Epilogue: TMEM -> registers (lane = position), + bias,
// SiLU, zero where the depthwise conv pads, bf16 pairs into the shared-memory tile the depthwise stencil reads; then the
// stencil exactly as in the SIMT kernel.  The 160x-expanded activation still never leaves the SM.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <algorithm>
#include <cstdlib>

namespace lcasr {

using namespace ptx;

namespace {

constexpr int kTcTT2 = 4;                  // depthwise output rows per CTA
constexpr int kTcA0R = 2 * kTcTT2 + 1;     // conv0 rows per CTA (9)
constexpr int kTcINR = 2 * kTcA0R + 1;     // input frames per CTA (19)
constexpr int kTcCG = 64;                  // channels per CTA
constexpr int kTcK = 32;                   // padded contraction length
constexpr int kTcThreads = 256;
constexpr int kTcRowB = 144;                // bytes per position of the stencil tile: 32 bf16 pairs + 16 bytes of padding, so that the
                                           // epilogue's 16-byte stores of consecutive positions fall into different banks

// canonical K-major no-swizzle operand: core matrix = 8 rows x 16 bytes (8 bf16), stored as 128 contiguous bytes;
// core matrices adjacent in K are 128 bytes apart (LBO), 8-row groups are kTcK/8 * 128 = 512 bytes apart (SBO)
__device__ __forceinline__ uint32_t tc_elem_off(int row, int chunk) { return (uint32_t)((row >> 3) * 512 + chunk * 128 + (row & 7) * 16); }

__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((128 >> 4) & 0x3FFF) << 16;   // leading-dimension byte offset: next core matrix along K
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;   // stride-dimension byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  return d;                                      // layout type 0: no swizzle
}

__device__ __forceinline__ uint32_t bf16x2_bits(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// 27 values [a(9) | b(9) | c(9)] + the two bias columns (e0, e1) + 3 zeros as four 16-byte chunks of bf16
__device__ __forceinline__ void tc_store_row(uint8_t* base, int row, const float (&a)[9], const float (&b)[9], const float (&c)[9],
                                             float e0, float e1) {
  float k[32];
#pragma unroll
  for (int i = 0; i < 9; ++i) { k[i] = a[i]; k[9 + i] = b[i]; k[18 + i] = c[i]; }
  k[27] = e0; k[28] = e1;
#pragma unroll
  for (int i = 29; i < 32; ++i) k[i] = 0.f;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 v;
    v.x = bf16x2_bits(k[8 * ch + 0], k[8 * ch + 1]);
    v.y = bf16x2_bits(k[8 * ch + 2], k[8 * ch + 3]);
    v.z = bf16x2_bits(k[8 * ch + 4], k[8 * ch + 5]);
    v.w = bf16x2_bits(k[8 * ch + 6], k[8 * ch + 7]);
    *reinterpret_cast<uint4*>(base + tc_elem_off(row, ch)) = v;
  }
}

// F1C > 0: the conv0 frequency width as a compile-time constant (80 mel bins -> 40): the stencil offsets become immediates and
// the position -> (row, column) divisions constant divisions (ncu of the first version: 48 % of the 571 M warp instructions at
// cfg 3 were integer / address arithmetic).  The bias and the 1/2 of SiLU's tanh form ride in the product: A gets two columns
// of ones (zero for positions outside the conv0 range, whose accumulator is then exactly 0 = SiLU(0): no select), B holds
// w/2 and b/2 (hi, lo) — the epilogue is h = acc, y = h + h tanh(h): MUFU + FFMA per value.
template <int F1C>
__global__ void __launch_bounds__(kTcThreads, 2)
subsample_conv0_dw_tc_kernel(const float* __restrict__ spec, const float* __restrict__ w0, const float* __restrict__ b0,
                             const float* __restrict__ w1, const float* __restrict__ b1, int F, int64_t T, int C, int64_t T1,
                             int F1_rt, int64_t T2, int F2, int FW, int MT, bf16* __restrict__ out) {
  const int F1 = F1C > 0 ? F1C : F1_rt;
  // MT = M tiles of 128 positions (ceil(9 * F1 / 128)); FW = pitch of the input patch
  extern __shared__ __align__(16) uint8_t tsm[];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
    const int A0W = F1 + 2;
  const uint32_t raw = smem_u32(tsm);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* sm = tsm + (base - raw);
  uint8_t* sA = sm;                                            // [MT*128][32] bf16, canonical layout
  uint8_t* sB = sA + (size_t)MT * 128 * kTcK * 2;              // [64][32] bf16
  float* s_in = reinterpret_cast<float*>(sB + kTcCG * kTcK * 2);   // [INR][FW]
  uint32_t* s_a0 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s_in) + ((kTcINR * FW * 4 + 127) & ~127));  // [A0R][A0W] rows of kTcRowB bytes: 32 bf16 pairs + padding
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cgi = blockIdx.y, b = blockIdx.z;
  const int64_t t2_0 = (int64_t)blockIdx.x * kTcTT2;
  const int64_t a0_row0 = 2 * t2_0 - 1;    // global conv0 row of tile row 0
  const int64_t t_in0 = 2 * a0_row0 - 1;   // global input frame of patch row 0
  const int npos = kTcA0R * F1;

  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 256);
    tmem_relinquish();
  }
  if (tid == 32) {
    mbar_init(smem_u32(&mma_bar), 1);
    fence_barrier_init();
  }
  // input patch (time-coalesced), zero outside the recording
  for (int idx = tid; idx < kTcINR * FW; idx += kTcThreads) {
    const int f = idx / kTcINR - 1, r = idx % kTcINR;
    const int64_t t = t_in0 + r;
    float v = 0.f;
    if (f >= 0 && f < F && t >= 0 && t < T) v = spec[((int64_t)b * F + f) * T + t];
    s_in[r * FW + (f + 1)] = v;
  }
  // weights of this CTA's 64 channels: [w_hi | w_hi | w_lo]
  if (tid < kTcCG) {
    const int c = cgi * kTcCG + tid;
    float wh[9], wl[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float w = 0.5f * w0[c * 9 + k];  // exact: the split of w/2 is half the split of w
      wh[k] = __bfloat162float(__float2bfloat16_rn(w));
      wl[k] = w - wh[k];
    }
    const float bh = 0.5f * b0[c];
    const float bhh = __bfloat162float(__float2bfloat16_rn(bh));
    tc_store_row(sB, tid, wh, wh, wl, bhh, bh - bhh);
  }
  // zero padding columns of the conv0 tile (the depthwise conv's left / right padding)
  for (int i = tid; i < kTcA0R * 2 * 32; i += kTcThreads) {
    const int r = i / 64, side = (i >> 5) & 1;
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s_a0) + (size_t)(r * A0W + (side ? A0W - 1 : 0)) * kTcRowB + (i & 31) * 4) = 0u;
  }
  __syncthreads();
  // im2col rows: [x_hi | x_lo | x_hi]
  const int r_lo = (int)max((int64_t)0, -a0_row0), r_hi = (int)min((int64_t)kTcA0R, T1 - a0_row0);  // conv0 rows that exist
  for (int p = tid; p < MT * 128; p += kTcThreads) {
    float xh[9], xl[9];
    const int r = p / F1, col = p - r * F1;
    const bool live = p < npos && r >= r_lo && r < r_hi;
    if (live) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float x = s_in[(2 * r + i) * FW + (2 * col + j)];
          xh[i * 3 + j] = __bfloat162float(__float2bfloat16_rn(x));
          xl[i * 3 + j] = x - xh[i * 3 + j];
        }
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) { xh[k] = 0.f; xl[k] = 0.f; }
    }
    tc_store_row(sA, p, xh, xl, xh, live ? 1.f : 0.f, live ? 1.f : 0.f);
  }
  fence_proxy_async();   // generic-proxy writes of A / B -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kTcCG, 0);
      const uint32_t a0 = smem_u32(sA), b0s = smem_u32(sB);
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < kTcK / 16; ++ks)  // K = 16 per instruction = two core matrices = 256 bytes along K
          umma_f16_ss(tmem_base + mt * kTcCG, tc_desc(a0 + mt * 16 * 512 + ks * 256), tc_desc(b0s + ks * 256), idesc, ks != 0);
      umma_commit(smem_u32(&mma_bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&mma_bar), 0);
  tc_fence_after();
  // epilogue: TMEM (lane = position) -> + bias, SiLU, zero outside the conv0 range -> bf16 pairs into the stencil tile
  {
    const int quarter = warp & 3, half = warp >> 2;   // TMEM lane quarter; which 32 of the 64 channels
    for (int mt = 0; mt < MT; ++mt) {
      const int p = mt * 128 + quarter * 32 + lane;
      uint32_t acc[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + mt * kTcCG + half * 32, acc);
      tmem_wait_ld();
      if (p < npos) {
        const int r = p / F1, col = p - r * F1;
        const int pos = r * A0W + col + 1;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {  // acc = (conv0 + bias) / 2 =: h; SiLU = h + h tanh(h)
          const float h0 = __uint_as_float(acc[2 * j]), h1 = __uint_as_float(acc[2 * j + 1]);
          pk[j] = bf16x2_bits(fmaf(h0, tanh_approx_f(h0), h0), fmaf(h1, tanh_approx_f(h1), h1));
        }
        // the position's row holds 32 channel pairs in 144 bytes: lanes of a warp are consecutive positions, so the 8 lanes
        // of a 16-byte store phase fall into 8 different bank groups; the stencil's reads (all lanes in one row) are contiguous
        uint8_t* rowp = reinterpret_cast<uint8_t*>(s_a0) + (size_t)pos * kTcRowB + half * 64;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(rowp + q * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
  // depthwise 3x3 stride 2 over the tile: a lane owns one channel pair (packed FFMA2), a warp one output position
  {
    const int c0 = cgi * kTcCG + 2 * lane;
    uint64_t wd2[9];
    const uint64_t bd2 = pack2f(b1[c0], b1[c0 + 1]);
#pragma unroll
    for (int k = 0; k < 9; ++k) wd2[k] = pack2f(w1[c0 * 9 + k], w1[(c0 + 1) * 9 + k]);
    const int tl_hi = (int)min((int64_t)kTcTT2, T2 - t2_0);
    int tl = 0, f2 = warp;
    while (f2 >= F2) { f2 -= F2; ++tl; }
    bf16* obase = out + (((int64_t)b * T2 + t2_0) * F2) * C + c0;
#pragma unroll 2
    for (int q = warp; q < kTcTT2 * F2; q += 8) {
      if (tl >= tl_hi) break;
      uint64_t acc2 = bd2;
      const uint8_t* a0p = reinterpret_cast<const uint8_t*>(s_a0) + (size_t)((2 * tl) * A0W + 2 * f2) * kTcRowB + lane * 4;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint32_t bits = *reinterpret_cast<const uint32_t*>(a0p + (size_t)(i * A0W + j) * kTcRowB);
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bits));
          ffma2_acc(acc2, wd2[i * 3 + j], pack2f(v.x, v.y));
        }
      float a, bq;
      unpack2f(acc2, a, bq);
      *reinterpret_cast<__nv_bfloat162*>(obase + (size_t)q * C) = __floats2bfloat162_rn(a, bq);
      f2 += 8;
      while (f2 >= F2) { f2 -= F2; ++tl; }
    }
  }
}

}  // namespace

// same contract as lcasr_subsample_conv0_dw (bf16 output [B, T2, F2, C], C % 64 == 0); returns LCASR_E_UNSUPPORTED when the
// tile does not fit (very wide feature axes), so that the caller can take the SIMT kernel
int subsample_conv0_dw_tc_launch(const float* spec, const float* w0, const float* b0, const float* w1, const float* b1, int B, int F,
                                 int64_t T, int C, void* out, cudaStream_t st) {
  const int64_t T1 = (T - 1) / 2 + 1, T2 = (T1 - 1) / 2 + 1;
  const int F1 = (F - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  const int MT = (int)ceil_div(kTcA0R * F1, 128);
  const int FW = ((2 * F1 + 2) + 3) & ~3;  // patch columns 0 .. 2*F1 (col 0 = frequency -1)
  if (MT * kTcCG > 256 || C % kTcCG != 0) return set_error(LCASR_E_UNSUPPORTED, "subsample(tc): feat_in=%d / C=%d not supported", F, C);
  const size_t smem = 128 + (size_t)MT * 128 * kTcK * 2 + kTcCG * kTcK * 2 + (((size_t)kTcINR * FW * 4 + 127) & ~(size_t)127) +
                      (size_t)kTcA0R * (F1 + 2) * kTcRowB;
  if (smem > 100 * 1024) return set_error(LCASR_E_UNSUPPORTED, "subsample(tc): tile needs %zu bytes of shared memory", smem);
  LCASR_CHECK_ARG(ceil_div(T2, kTcTT2) <= 0x7fffffff && B <= 65535, "subsample(tc): grid too large");
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(subsample_conv0_dw_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    LCASR_CUDA(cudaFuncSetAttribute(subsample_conv0_dw_tc_kernel<40>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set.mark(attr_dev);
  }
  dim3 grid((unsigned)ceil_div(T2, kTcTT2), (unsigned)(C / kTcCG), (unsigned)B);
  if (F1 == 40)
    subsample_conv0_dw_tc_kernel<40><<<grid, kTcThreads, smem, st>>>(spec, w0, b0, w1, b1, F, T, C, T1, F1, T2, F2, FW, MT, (bf16*)out);
  else
    subsample_conv0_dw_tc_kernel<0><<<grid, kTcThreads, smem, st>>>(spec, w0, b0, w1, b1, F, T, C, T1, F1, T2, F2, FW, MT, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

}  // namespace lcasr
