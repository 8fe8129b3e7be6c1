// GLU, rotary tables, qkv split + rotary.  All HBM-bound, 128-bit vectorised.
#include "common.cuh"

namespace lcasr {

// ---- GLU over channels (convolution.py:107): out[m,j] = in[m,j] * sigmoid(in[m,d+j]) ----------
// `lengths` (may be NULL): valid tokens per batch entry of N tokens; rows at or beyond it are written as
// zeros — the reference's `x.masked_fill(pad_mask, 0)` in front of the depthwise conv (convolution.py:109-110)
template <typename T>
__global__ void __launch_bounds__(256) glu_kernel(const T* __restrict__ in, int64_t M, int d, T* __restrict__ out,
                                                  const int32_t* __restrict__ lengths, int64_t N) {
  const int vec_per_row = d / 8;
  const int64_t total = M * vec_per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = idx / vec_per_row;
    int j = (int)(idx % vec_per_row) * 8;
    float a[8], g[8], y[8];
    Vec8<T>::load(in + m * 2 * d + j, a);
    Vec8<T>::load(in + m * 2 * d + d + j, g);
    const bool pad = lengths && (m % N) >= lengths[m / N];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = pad ? 0.f : a[i] * sigmoid_for<T>(g[i]);
    Vec8<T>::store(out + m * d + j, y);
  }
}

// ---- rotary tables (rotary_emb.py:52-56): fp32 t = pos / interp ; ang = t * inv_freq ----------
// TRANSPOSED: tables are written pair-major [half, N] (the layout the fused-rotary GEMM epilogue reads: a warp's 32 lanes
// are 32 consecutive token positions, so one pair index is one coalesced 128-byte line)
template <bool TRANSPOSED>
__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float interp, int64_t pos_offset, int64_t N,
                                  int half, float* __restrict__ cos_out, float* __restrict__ sin_out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * half) return;
  const int64_t n = TRANSPOSED ? idx % N : idx / half;
  const int j = (int)(TRANSPOSED ? idx / N : idx % half);
  float t = (float)(pos_offset + n) / interp;
  float ang = t * inv_freq[j];
  float s, c;
  sincosf(ang, &s, &c);  // full-range (Payne-Hanek) fp32 sin/cos: positions reach 45k * inv_freq<=1
  cos_out[idx] = c;
  sin_out[idx] = s;
}

// ---- qkv split + NeoX rotary (attention.py:485, rotary_emb.py:61-73) --------------------------
// qkv row = [q(h,dh) | k(h,dh) | v(h,dh)];  q' = q*cos + rotate_half(q)*sin with
// rotate_half(x)[j] = -x[j+half] (j < half), x[j-half] (j >= half).
// One thread handles 8 consecutive dh of the lower half and the matching 8 of the upper half.
template <typename T>
__global__ void __launch_bounds__(256) rope_split_kernel(const T* __restrict__ qkv, int64_t N, int H, int Dh,
                                                         const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                         int64_t total, T* __restrict__ q, T* __restrict__ k,
                                                         T* __restrict__ v, int v_transposed, int64_t Npad) {
  const int half = Dh / 2;
  const int vec_per_head = half / 8;       // per half
  const int d = H * Dh;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int jv = (int)(idx % vec_per_head);
    int64_t r = idx / vec_per_head;
    int h = (int)(r % H);
    int64_t m = r / H;          // b*N + n
    int64_t n = m % N;
    int64_t b = m / N;
    const int j = jv * 8;
    float c[8], s[8];
    if (cos_t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = cos_t[n * half + j + i]; s[i] = sin_t[n * half + j + i]; }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = 1.f; s[i] = 0.f; }
    }
    const T* row = qkv + m * 3 * d;
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // q then k
      float lo[8], hi[8], olo[8], ohi[8];
      Vec8<T>::load(row + which * d + h * Dh + j, lo);
      Vec8<T>::load(row + which * d + h * Dh + half + j, hi);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        olo[i] = lo[i] * c[i] - hi[i] * s[i];
        ohi[i] = hi[i] * c[i] + lo[i] * s[i];
      }
      T* dst = (which == 0 ? q : k) + m * d + h * Dh;
      Vec8<T>::store(dst + j, olo);
      Vec8<T>::store(dst + half + j, ohi);
    }
    float vlo[8], vhi[8];
    Vec8<T>::load(row + 2 * d + h * Dh + j, vlo);
    Vec8<T>::load(row + 2 * d + h * Dh + half + j, vhi);
    if (!v_transposed) {
      Vec8<T>::store(v + m * d + h * Dh + j, vlo);
      Vec8<T>::store(v + m * d + h * Dh + half + j, vhi);
    } else {  // [B,H,Dh,Npad]
      T* base = v + ((b * H + h) * Dh) * Npad + n;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        base[(int64_t)(j + i) * Npad] = from_f32<T>(vlo[i]);
        base[(int64_t)(half + j + i) * Npad] = from_f32<T>(vhi[i]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_f32_kernel(const float* __restrict__ in, int64_t n8, T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    Vec8<float>::load(in + i * 8, v);
    Vec8<T>::store(out + i * 8, v);
  }
}

static inline unsigned grid_for(int64_t total, int block) {
  int64_t blocks = ceil_div(total, block);
  int64_t cap = (int64_t)kNumSMs * 16;
  return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// zero the rows of padded tokens in place: x [B,N,d], rows n >= lengths[b]  (the masked_fill(pad_mask, 0) of
// attention.py:511,541 and its backward).  One warp per padded row segment; valid rows are not touched.
__global__ void __launch_bounds__(256) mask_rows_kernel(uint4* __restrict__ x, int64_t M, int64_t N, int vec_per_row,
                                                        const int32_t* __restrict__ lengths) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  if ((warp % N) < lengths[warp / N]) return;
  uint4* row = x + warp * vec_per_row;
  for (int i = lane; i < vec_per_row; i += 32) row[i] = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_mask_rows(void* x, int dtype, int B, int64_t N, int d, const int32_t* lengths, void* stream) {
  LCASR_CHECK_ARG(x && lengths && B > 0 && N > 0 && d > 0, "mask_rows: bad arguments");
  const size_t row_bytes = (size_t)d * dtype_size(dtype);
  LCASR_CHECK_ARG(row_bytes % 16 == 0 && (uintptr_t)x % 16 == 0, "mask_rows: rows must be multiples of 16 bytes");
  const int64_t M = (int64_t)B * N;
  mask_rows_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, (cudaStream_t)stream>>>((uint4*)x, M, N, (int)(row_bytes / 16), lengths);
  LCASR_LAUNCH_CHECK();
  return 0;
}

static int glu_launch(const void* in, int dtype, int64_t M, int d, void* out, const int32_t* lengths, int64_t N, void* stream) {
  LCASR_CHECK_ARG(in && out && M >= 0 && d > 0 && d % 8 == 0, "glu: bad arguments (d=%d must be a multiple of 8)", d);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = M * (d / 8);
  if (dtype == LCASR_BF16)
    glu_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)in, M, d, (bf16*)out, lengths, N);
  else
    glu_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)in, M, d, (float*)out, lengths, N);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_glu(const void* in, int dtype, int64_t M, int d, void* out, void* stream) {
  return glu_launch(in, dtype, M, d, out, nullptr, 1, stream);
}

extern "C" int lcasr_glu_masked(const void* in, int dtype, int B, int64_t N, int d, const int32_t* lengths, void* out,
                                void* stream) {
  LCASR_CHECK_ARG(B > 0 && N > 0, "glu_masked: bad shape");
  return glu_launch(in, dtype, (int64_t)B * N, d, out, lengths, N, stream);
}

extern "C" int lcasr_rope_table(const float* inv_freq, float interp, int64_t pos_offset, int64_t N, int half,
                                float* cos_out, float* sin_out, void* stream) {
  LCASR_CHECK_ARG(inv_freq && cos_out && sin_out && N > 0 && half > 0 && interp > 0.f, "rope_table: bad arguments");
  int64_t total = N * half;
  rope_table_kernel<false><<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(inv_freq, interp, pos_offset, N,
                                                                                              half, cos_out, sin_out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_rope_table_t(const float* inv_freq, float interp, int64_t pos_offset, int64_t N, int half,
                                  float* cos_out, float* sin_out, void* stream) {
  LCASR_CHECK_ARG(inv_freq && cos_out && sin_out && N > 0 && half > 0 && interp > 0.f, "rope_table_t: bad arguments");
  int64_t total = N * half;
  rope_table_kernel<true><<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(inv_freq, interp, pos_offset, N,
                                                                                             half, cos_out, sin_out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_rope_split(const void* qkv, int dtype, int B, int64_t N, int H, int Dh, const float* cos_t,
                                const float* sin_t, void* q, void* k, void* v, int v_transposed, int64_t Npad,
                                void* stream) {
  LCASR_CHECK_ARG(qkv && q && k && v && B > 0 && N > 0 && H > 0, "rope_split: bad arguments");
  LCASR_CHECK_ARG(Dh % 16 == 0, "rope_split: head_dim=%d must be a multiple of 16", Dh);
  LCASR_CHECK_ARG((cos_t == nullptr) == (sin_t == nullptr), "rope_split: cos/sin must both be given or both NULL");
  LCASR_CHECK_ARG(!v_transposed || Npad >= N, "rope_split: Npad < N");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)B * N * H * (Dh / 16);
  if (dtype == LCASR_BF16)
    rope_split_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)qkv, N, H, Dh, cos_t, sin_t, total,
                                                                   (bf16*)q, (bf16*)k, (bf16*)v, v_transposed, Npad);
  else
    rope_split_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)qkv, N, H, Dh, cos_t, sin_t, total,
                                                                    (float*)q, (float*)k, (float*)v, v_transposed, Npad);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_cast_f32(const float* in, int64_t n, void* out, int out_dtype, void* stream) {
  LCASR_CHECK_ARG(in && out && n >= 0 && n % 8 == 0, "cast_f32: bad arguments (n must be a multiple of 8)");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == LCASR_BF16) cast_f32_kernel<bf16><<<grid_for(n / 8, 256), 256, 0, st>>>(in, n / 8, (bf16*)out);
  else cast_f32_kernel<float><<<grid_for(n / 8, 256), 256, 0, st>>>(in, n / 8, (float*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}
