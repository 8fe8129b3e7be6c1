// SpecAugment on the device — lcasr/utils/augmentation.py:61-104 (time / frequency masking of the [B, F, T]
// spectrogram batch just before the encoder, exp/train.py:227).  The reference applies its n_time + n_freq masks one
// after the other, each a full torch.where pass (plus the mask construction) over the batch; here ONE pass applies all
// of them: the mask intervals are rebuilt per CTA from the uniform draws exactly like
// torchaudio.functional.mask_along_axis(_iid) does (fp32 products, truncation to integers), and the fill value (the
// mean over the un-padded frames, or 0) is read from a device scalar — no host synchronisation anywhere.
// HBM-bound: B*F*T*4 bytes in, the same out (+ one read for the mean).
#include "common.cuh"

namespace lcasr {

constexpr int kAugMaxMasks = 64;

// sum and element count over the valid region t < lengths[b] (whole rows when lengths == NULL); fp64 partials
__global__ void __launch_bounds__(256) specaug_sum_kernel(const float* __restrict__ x, int F, int64_t T,
                                                          const int32_t* __restrict__ lengths, double* __restrict__ acc) {
  __shared__ double red[8];
  const int b = blockIdx.z, f = blockIdx.y;
  const int64_t len = lengths ? min((int64_t)max(lengths[b], 0), T) : T;
  const float* row = x + ((int64_t)b * F + f) * T;
  double s = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < len; t += (int64_t)gridDim.x * blockDim.x) s += (double)row[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    atomicAdd(acc, tot);
    if (blockIdx.x == 0) atomicAdd(acc + 1, (double)len);
  }
}

// [start, end) of one mask from its two uniform draws — mask_along_axis_iid: value = u1 * mask_param;
// min_value = u2 * (size - value); start = trunc(min_value); end = start + trunc(value)   (all fp32, no contraction)
__device__ __forceinline__ void aug_interval(float u1, float u2, int mask_param, int64_t size, int64_t& lo, int64_t& hi) {
  const float value = __fmul_rn(u1, (float)mask_param);
  const float min_value = __fmul_rn(u2, __fsub_rn((float)size, value));
  lo = (int64_t)min_value;
  hi = lo + (int64_t)value;
}

template <int VEC>
__global__ void __launch_bounds__(256) specaug_apply_kernel(const float* __restrict__ x, int F, int64_t T, int n_time, int time_param,
                                                            const float* __restrict__ u_time, int n_freq, int freq_param,
                                                            const float* __restrict__ u_freq, int draws_per_mask,
                                                            const double* __restrict__ mean_acc, float* __restrict__ out) {
  __shared__ int64_t t_lo[kAugMaxMasks], t_hi[kAugMaxMasks];
  __shared__ int s_row_masked;
  const int b = blockIdx.z, f = blockIdx.y;
  const int d = draws_per_mask == 1 ? 0 : b;  // one shared interval per mask (iid_masks=False) or one per recording
  if (threadIdx.x == 0) s_row_masked = 0;
  __syncthreads();
  if (threadIdx.x < n_time) {
    const float* u = u_time + (int64_t)threadIdx.x * 2 * draws_per_mask;
    aug_interval(u[d], u[draws_per_mask + d], time_param, T, t_lo[threadIdx.x], t_hi[threadIdx.x]);
  } else if (threadIdx.x >= 64 && threadIdx.x - 64 < n_freq) {
    const float* u = u_freq + (int64_t)(threadIdx.x - 64) * 2 * draws_per_mask;
    int64_t lo, hi;
    aug_interval(u[d], u[draws_per_mask + d], freq_param, F, lo, hi);
    if (f >= lo && f < hi) s_row_masked = 1;
  }
  __syncthreads();
  const float fill = mean_acc ? (float)(mean_acc[0] / mean_acc[1]) : 0.0f;
  const bool row_masked = s_row_masked != 0;
  const int64_t base = ((int64_t)b * F + f) * T;
  const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (t0 >= T) return;
  float v[VEC];
  if constexpr (VEC == 4) {
    const float4 r = *reinterpret_cast<const float4*>(x + base + t0);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  } else {
    v[0] = x[base + t0];
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    bool m = row_masked;
    const int64_t t = t0 + j;
    for (int k = 0; k < n_time; ++k) m = m || (t >= t_lo[k] && t < t_hi[k]);
    if (m) v[j] = fill;
  }
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(out + base + t0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    out[base + t0] = v[0];
  }
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_specaug_mean(const float* spec, int B, int F, int64_t T, const int32_t* lengths, double* acc, void* stream) {
  LCASR_CHECK_ARG(spec && acc, "specaug_mean: NULL argument");
  LCASR_CHECK_ARG(B > 0 && B <= 65535 && F > 0 && F <= 65535 && T > 0, "specaug_mean: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  LCASR_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(T, 256 * 8), 64), (unsigned)F, (unsigned)B);
  specaug_sum_kernel<<<grid, 256, 0, st>>>(spec, F, T, lengths, acc);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_specaug_apply(const float* spec, int B, int F, int64_t T, int n_time, int time_param, const float* u_time,
                                   int n_freq, int freq_param, const float* u_freq, int draws_per_mask,
                                   const double* mean_acc, float* out, void* stream) {
  LCASR_CHECK_ARG(spec && out, "specaug_apply: NULL argument");
  LCASR_CHECK_ARG(B > 0 && B <= 65535 && F > 0 && F <= 65535 && T > 0, "specaug_apply: bad shape");
  LCASR_CHECK_ARG(n_time >= 0 && n_time <= kAugMaxMasks && n_freq >= 0 && n_freq <= kAugMaxMasks,
                  "specaug_apply: at most %d masks per axis", kAugMaxMasks);
  LCASR_CHECK_ARG((n_time == 0 || (u_time && time_param >= 1)) && (n_freq == 0 || (u_freq && freq_param >= 1)),
                  "specaug_apply: masks need their draws and a mask parameter >= 1");
  LCASR_CHECK_ARG(draws_per_mask == 1 || draws_per_mask == B, "specaug_apply: draws_per_mask is 1 (shared masks) or B (iid masks)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = T % 4 == 0 && ((uintptr_t)spec % 16 == 0) && ((uintptr_t)out % 16 == 0);
  if (vec) {
    dim3 grid((unsigned)ceil_div(T, 256 * 4), (unsigned)F, (unsigned)B);
    specaug_apply_kernel<4><<<grid, 256, 0, st>>>(spec, F, T, n_time, time_param, u_time, n_freq, freq_param, u_freq, draws_per_mask,
                                                  mean_acc, out);
  } else {
    dim3 grid((unsigned)ceil_div(T, 256), (unsigned)F, (unsigned)B);
    specaug_apply_kernel<1><<<grid, 256, 0, st>>>(spec, F, T, n_time, time_param, u_time, n_freq, freq_param, u_freq, draws_per_mask,
                                                  mean_acc, out);
  }
  LCASR_LAUNCH_CHECK();
  return 0;
}
