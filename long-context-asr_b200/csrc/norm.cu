// LayerNorm / RMSNorm over the last dim of the fp32 residual stream.
// HBM-bound: algorithmic bytes per row = d*4 (read) + d*4 (fp32 out, if any) + d*e (operand copy).
// One warp per row; the row lives in registers (128-bit loads), mean and centred variance are
// warp-shuffle reductions (two-pass, like torch's CPU/CUDA LayerNorm, so fp32 parity holds).
#include "common.cuh"

namespace lcasr {

template <int NV, typename TLo>  // NV float4 per lane: d == NV*128
__global__ void __launch_bounds__(256) norm_rows_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, int64_t M, float eps, int kind,
                                                        float* __restrict__ out_f32, TLo* __restrict__ out_lo) {
  constexpr int d = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  float mean = 0.f, rstd;
  if (kind == LCASR_NORM_LAYERNORM) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    mean = warp_sum(s) * (1.0f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + e * e);
    }
    rstd = rsqrtf(warp_sum(q) * (1.0f / d) + eps);
  } else {  // lcasr RMSNorm: x / (||x|| * d^-1/2 + eps)
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    rstd = 1.0f / (sqrtf(warp_sum(q)) * rsqrtf((float)d) + eps);
  }
  const float4* wr = reinterpret_cast<const float4*>(w);
  const float4* br = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 ww = wr[lane + 32 * i];
    float4 bb = b ? br[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 y;
    y.x = (v[i].x - mean) * rstd * ww.x + bb.x;
    y.y = (v[i].y - mean) * rstd * ww.y + bb.y;
    y.z = (v[i].z - mean) * rstd * ww.z + bb.z;
    y.w = (v[i].w - mean) * rstd * ww.w + bb.w;
    if (out_f32) reinterpret_cast<float4*>(out_f32 + row * d)[lane + 32 * i] = y;
    if (out_lo) {
      TLo* o = out_lo + row * d + (lane + 32 * i) * 4;
      if constexpr (sizeof(TLo) == 2) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(y.x, y.y), p1 = __floats2bfloat162_rn(y.z, y.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(o) = u;
      } else {
        *reinterpret_cast<float4*>(o) = y;
      }
    }
  }
}

// Up to three norms applied back to back to a row that stays in registers (ConformerLayer.norm_out -> decoder.norm of the
// self-conditioning branch, sconformer_xl.py:371 + decoder.py:23; at the end norm_out -> decoder.norm -> decoder.norm with
// legasee_double_norm, :245-247): one read of the fp32 residual stream instead of one per norm.  out_f32 receives the result
// of stage `f32_stage` (or nothing when NULL), out_lo the result of the last stage.  Same arithmetic per stage as above.
struct NormChain { const float* w[3]; const float* b[3]; int n; int f32_stage; };

template <int NV, typename TLo>
__global__ void __launch_bounds__(256) norm_chain_kernel(const float* __restrict__ x, NormChain ch, int64_t M, float eps, int kind,
                                                         float* __restrict__ out_f32, TLo* __restrict__ out_lo) {
  constexpr int d = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  for (int sidx = 0; sidx < ch.n; ++sidx) {
    float mean = 0.f, rstd;
    if (kind == LCASR_NORM_LAYERNORM) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      mean = warp_sum(s) * (1.0f / d);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + bb * bb) + (c * c + e * e);
      }
      rstd = rsqrtf(warp_sum(q) * (1.0f / d) + eps);
    } else {
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      rstd = 1.0f / (sqrtf(warp_sum(q)) * rsqrtf((float)d) + eps);
    }
    const float4* wr = reinterpret_cast<const float4*>(ch.w[sidx]);
    const float4* br = reinterpret_cast<const float4*>(ch.b[sidx]);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 ww = wr[lane + 32 * i];
      const float4 bb = br ? br[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
      v[i].x = (v[i].x - mean) * rstd * ww.x + bb.x;
      v[i].y = (v[i].y - mean) * rstd * ww.y + bb.y;
      v[i].z = (v[i].z - mean) * rstd * ww.z + bb.z;
      v[i].w = (v[i].w - mean) * rstd * ww.w + bb.w;
      if (out_f32 && sidx == ch.f32_stage) reinterpret_cast<float4*>(out_f32 + row * d)[lane + 32 * i] = v[i];
    }
  }
  if (out_lo) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      TLo* o = out_lo + row * d + (lane + 32 * i) * 4;
      if constexpr (sizeof(TLo) == 2) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i].x, v[i].y), p1 = __floats2bfloat162_rn(v[i].z, v[i].w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(o) = u;
      } else {
        *reinterpret_cast<float4*>(o) = v[i];
      }
    }
  }
}

// any d (tiny test configs): same arithmetic, strided scalar accesses (rows re-read from L1/L2)
template <typename TLo>
__global__ void __launch_bounds__(256) norm_rows_generic_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ b, int64_t M, int d, float eps,
                                                                int kind, float* __restrict__ out_f32,
                                                                TLo* __restrict__ out_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * d;
  float mean = 0.f, rstd;
  if (kind == LCASR_NORM_LAYERNORM) {
    float s = 0.f;
    for (int i = lane; i < d; i += 32) s += xr[i];
    mean = warp_sum(s) / d;
    float q = 0.f;
    for (int i = lane; i < d; i += 32) { float a = xr[i] - mean; q += a * a; }
    rstd = rsqrtf(warp_sum(q) / d + eps);
  } else {
    float q = 0.f;
    for (int i = lane; i < d; i += 32) q += xr[i] * xr[i];
    rstd = 1.0f / (sqrtf(warp_sum(q)) * rsqrtf((float)d) + eps);
  }
  // all lanes have read the row before anyone overwrites it (in-place safe): values kept below
  for (int i0 = 0; i0 < d; i0 += 32) {
    int i = i0 + lane;
    float xv = i < d ? xr[i] : 0.f;
    __syncwarp();
    if (i < d) {
      float y = (xv - mean) * rstd * w[i] + (b ? b[i] : 0.f);
      if (out_f32) out_f32[row * d + i] = y;
      if (out_lo) out_lo[row * d + i] = from_f32<TLo>(y);
    }
  }
}

template <typename TLo>
static int launch_norm(const float* x, const float* w, const float* b, int64_t M, int d, float eps, int kind,
                       float* out_f32, TLo* out_lo, cudaStream_t st) {
  const int rows_per_block = 8;
  dim3 grid((unsigned)ceil_div(M, rows_per_block)), block(rows_per_block * 32);
#define LCASR_NORM_CASE(NV)                                                                          \
  case NV * 128:                                                                                     \
    norm_rows_kernel<NV, TLo><<<grid, block, 0, st>>>(x, w, b, M, eps, kind, out_f32, out_lo);      \
    break;
  switch (d) {
    LCASR_NORM_CASE(1) LCASR_NORM_CASE(2) LCASR_NORM_CASE(4) LCASR_NORM_CASE(6) LCASR_NORM_CASE(8)
    LCASR_NORM_CASE(12) LCASR_NORM_CASE(16)
    default:
      norm_rows_generic_kernel<TLo><<<grid, block, 0, st>>>(x, w, b, M, d, eps, kind, out_f32, out_lo);
  }
#undef LCASR_NORM_CASE
  LCASR_LAUNCH_CHECK();
  return 0;
}

template <typename TLo>
static int launch_norm_chain(const float* x, const NormChain& ch, int64_t M, int d, float eps, int kind, float* out_f32, TLo* out_lo,
                             cudaStream_t st) {
  const int rows_per_block = 8;
  dim3 grid((unsigned)ceil_div(M, rows_per_block)), block(rows_per_block * 32);
#define LCASR_NORMC_CASE(NV)                                                                           \
  case NV * 128:                                                                                       \
    norm_chain_kernel<NV, TLo><<<grid, block, 0, st>>>(x, ch, M, eps, kind, out_f32, out_lo);         \
    break;
  switch (d) {
    LCASR_NORMC_CASE(1) LCASR_NORMC_CASE(2) LCASR_NORMC_CASE(4) LCASR_NORMC_CASE(6) LCASR_NORMC_CASE(8)
    LCASR_NORMC_CASE(12) LCASR_NORMC_CASE(16)
    default:
      return set_error(LCASR_E_UNSUPPORTED, "layernorm_chain: d=%d is not a multiple of 128 in {128..2048}", d);
  }
#undef LCASR_NORMC_CASE
  LCASR_LAUNCH_CHECK();
  return 0;
}

bool norm_chain_ok(int d) { return d % 128 == 0 && (d / 128 == 1 || d / 128 == 2 || d / 128 == 4 || d / 128 == 6 || d / 128 == 8 || d / 128 == 12 || d / 128 == 16); }

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_layernorm_chain(const float* x, int n_stages, const float* const* weights, const float* const* biases,
                                     int64_t M, int d, float eps, int kind, int f32_stage, float* out_f32, void* out_lo,
                                     int lo_dtype, void* stream) {
  LCASR_CHECK_ARG(x && weights && n_stages >= 1 && n_stages <= 3 && M >= 0 && d > 0, "layernorm_chain: bad arguments");
  LCASR_CHECK_ARG(kind == LCASR_NORM_LAYERNORM || kind == LCASR_NORM_RMSNORM, "layernorm_chain: bad kind %d", kind);
  LCASR_CHECK_ARG(out_f32 || out_lo, "layernorm_chain: no output given");
  LCASR_CHECK_ARG(!out_f32 || (f32_stage >= 0 && f32_stage < n_stages), "layernorm_chain: bad f32_stage");
  if (M == 0) return 0;
  NormChain ch{};
  ch.n = n_stages; ch.f32_stage = f32_stage;
  for (int i = 0; i < n_stages; ++i) {
    LCASR_CHECK_ARG(weights[i], "layernorm_chain: NULL weight");
    ch.w[i] = weights[i];
    ch.b[i] = biases ? biases[i] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (out_lo && lo_dtype == LCASR_BF16) return launch_norm_chain<bf16>(x, ch, M, d, eps, kind, out_f32, (bf16*)out_lo, st);
  return launch_norm_chain<float>(x, ch, M, d, eps, kind, out_f32, (float*)out_lo, st);
}

extern "C" int lcasr_layernorm(const float* x, const float* weight, const float* bias, int64_t M, int d, float eps,
                               int kind, float* out_f32, void* out_lo, int lo_dtype, void* stream) {
  LCASR_CHECK_ARG(x && weight && M >= 0 && d > 0, "layernorm: bad arguments");
  LCASR_CHECK_ARG(kind == LCASR_NORM_LAYERNORM || kind == LCASR_NORM_RMSNORM, "layernorm: bad kind %d", kind);
  LCASR_CHECK_ARG(out_f32 || out_lo, "layernorm: no output given");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_lo && lo_dtype == LCASR_BF16)
    return launch_norm<bf16>(x, weight, bias, M, d, eps, kind, out_f32, (bf16*)out_lo, st);
  return launch_norm<float>(x, weight, bias, M, d, eps, kind, out_f32, (float*)out_lo, st);
}
