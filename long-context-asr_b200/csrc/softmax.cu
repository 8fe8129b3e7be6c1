// Row softmax over the V=4096 CTC classes (self-conditioning, sconformer_xl.py:242), the final
// log-softmax fused with the greedy argmax (decoder.py:25 + decoding/greedy.py:19), and the greedy
// collapse (unique_consecutive + blank removal, greedy.py:20-21).
// All HBM-bound: softmax M*V*(e_in+e_out); log-softmax M*V*8 + M*4; collapse ~ B*N*8 bytes.
#include "common.cuh"

namespace lcasr {

constexpr int kSmThreads = 256;

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < kSmThreads / 32; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kSmThreads / 32; ++i) r += red[i];
  __syncthreads();
  return r;
}

// One CTA per row; the row is cached in registers when V <= 8*256*NV8.
template <typename TIn, typename TOut, int NV8>
__global__ void __launch_bounds__(kSmThreads) softmax_rows_kernel(const TIn* __restrict__ in, int V,
                                                                  TOut* __restrict__ out) {
  __shared__ float red[kSmThreads / 32];
  const int64_t row = blockIdx.x;
  const TIn* x = in + row * V;
  float v[NV8][8];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    int col = (i * kSmThreads + threadIdx.x) * 8;
    if (col < V) {
      Vec8<TIn>::load(x + col, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, v[i][j]);
    }
  }
  mx = block_max(mx, red);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    int col = (i * kSmThreads + threadIdx.x) * 8;
    if (col < V) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[i][j] = __expf(v[i][j] - mx); s += v[i][j]; }
    }
  }
  s = block_sum(s, red);
  const float inv = 1.0f / s;
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    int col = (i * kSmThreads + threadIdx.x) * 8;
    if (col < V) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] *= inv;
      Vec8<TOut>::store(out + row * V + col, v[i]);
    }
  }
}

// log-softmax in place (fp32) + first-index argmax
template <int NV8>
__global__ void __launch_bounds__(kSmThreads) log_softmax_argmax_kernel(float* __restrict__ logits, int V,
                                                                        int32_t* __restrict__ argmax) {
  __shared__ float red[kSmThreads / 32];
  __shared__ int red_i[kSmThreads / 32];
  const int64_t row = blockIdx.x;
  float* x = logits + row * V;
  float v[NV8][8];
  float mx = -INFINITY;
  int mi = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    int col = (i * kSmThreads + threadIdx.x) * 8;
    if (col < V) {
      Vec8<float>::load(x + col, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[i][j] > mx) { mx = v[i][j]; mi = col + j; }  // ascending scan keeps the first max
    }
  }
  // (max, lowest index) reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, mx, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mx; red_i[threadIdx.x >> 5] = mi; }
  __syncthreads();
  mx = red[0]; mi = red_i[0];
#pragma unroll
  for (int i = 1; i < kSmThreads / 32; ++i)
    if (red[i] > mx || (red[i] == mx && red_i[i] < mi)) { mx = red[i]; mi = red_i[i]; }
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    int col = (i * kSmThreads + threadIdx.x) * 8;
    if (col < V) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s += expf(v[i][j] - mx);
    }
  }
  s = block_sum(s, red);
  const float lse = mx + logf(s);
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    int col = (i * kSmThreads + threadIdx.x) * 8;
    if (col < V) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] -= lse;
      Vec8<float>::store(x + col, v[i]);
    }
  }
  if (argmax && threadIdx.x == 0) argmax[row] = mi == 0x7fffffff ? 0 : mi;
}

// argmax over classes of rows that are already log-probs (GreedyCTCDecoder on a caller's tensor)
__global__ void __launch_bounds__(kSmThreads) argmax_rows_kernel(const float* __restrict__ x, int V,
                                                                 int32_t* __restrict__ argmax) {
  __shared__ float red[kSmThreads / 32];
  __shared__ int red_i[kSmThreads / 32];
  const float* r = x + (int64_t)blockIdx.x * V;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int c = threadIdx.x; c < V; c += kSmThreads) {
    float v = r[c];
    if (v > mx) { mx = v; mi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, mx, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mx; red_i[threadIdx.x >> 5] = mi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    mx = red[0]; mi = red_i[0];
    for (int i = 1; i < kSmThreads / 32; ++i)
      if (red[i] > mx || (red[i] == mx && red_i[i] < mi)) { mx = red[i]; mi = red_i[i]; }
    argmax[blockIdx.x] = mi == 0x7fffffff ? 0 : mi;
  }
}

// Greedy collapse: one CTA per recording; keep[t] = a[t] != a[t-1] && a[t] != blank; block scan.
__global__ void __launch_bounds__(1024) greedy_collapse_kernel(const int32_t* __restrict__ argmax, int64_t N,
                                                               const int32_t* __restrict__ lengths, int blank,
                                                               int32_t* __restrict__ tokens, int32_t* __restrict__ n_tokens) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int64_t len = lengths ? (lengths[b] < N ? lengths[b] : N) : N;
  const int32_t* a = argmax + (int64_t)b * N;
  int32_t* out = tokens + (int64_t)b * N;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < len; base += blockDim.x) {
    int64_t t = base + threadIdx.x;
    int cur = 0, keep = 0;
    if (t < len) {
      cur = a[t];
      int prev = t > 0 ? a[t - 1] : -1;
      keep = (cur != prev) && (cur != blank);
    }
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    int in_warp = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int off = carry;
    for (int i = 0; i < wid; ++i) off += warp_tot[i];
    if (keep) out[off + in_warp] = cur;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += warp_tot[i];
      carry += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_tokens[b] = carry;
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_softmax(const void* in, int in_dtype, int64_t M, int V, void* out, int out_dtype, void* stream) {
  LCASR_CHECK_ARG(in && out && M >= 0 && V > 0 && V % 8 == 0, "softmax: bad arguments (V=%d must be a multiple of 8)", V);
  LCASR_CHECK_ARG(V <= 8 * kSmThreads * 4, "softmax: V=%d > %d unsupported", V, 8 * kSmThreads * 4);
  LCASR_CHECK_ARG(in_dtype == out_dtype, "softmax: in/out dtypes must match");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int nv8 = (int)ceil_div(V, 8 * kSmThreads);
#define LCASR_SM_LAUNCH(T, NV)                                                                             \
  softmax_rows_kernel<T, T, NV><<<(unsigned)M, kSmThreads, 0, st>>>((const T*)in, V, (T*)out)
  if (in_dtype == LCASR_BF16) {
    if (nv8 <= 1) LCASR_SM_LAUNCH(bf16, 1); else if (nv8 == 2) LCASR_SM_LAUNCH(bf16, 2); else LCASR_SM_LAUNCH(bf16, 4);
  } else {
    if (nv8 <= 1) LCASR_SM_LAUNCH(float, 1); else if (nv8 == 2) LCASR_SM_LAUNCH(float, 2); else LCASR_SM_LAUNCH(float, 4);
  }
#undef LCASR_SM_LAUNCH
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_log_softmax_argmax(float* logits, int64_t M, int V, int32_t* argmax, void* stream) {
  LCASR_CHECK_ARG(logits && M >= 0 && V > 0 && V % 8 == 0, "log_softmax_argmax: bad arguments (V=%d)", V);
  LCASR_CHECK_ARG(V <= 8 * kSmThreads * 4, "log_softmax_argmax: V=%d > %d unsupported", V, 8 * kSmThreads * 4);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int nv8 = (int)ceil_div(V, 8 * kSmThreads);
  if (nv8 <= 1) log_softmax_argmax_kernel<1><<<(unsigned)M, kSmThreads, 0, st>>>(logits, V, argmax);
  else if (nv8 == 2) log_softmax_argmax_kernel<2><<<(unsigned)M, kSmThreads, 0, st>>>(logits, V, argmax);
  else log_softmax_argmax_kernel<4><<<(unsigned)M, kSmThreads, 0, st>>>(logits, V, argmax);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_greedy_collapse(const int32_t* argmax, int B, int64_t N, const int32_t* lengths, int blank,
                                     int32_t* tokens, int32_t* n_tokens, void* stream) {
  LCASR_CHECK_ARG(argmax && tokens && n_tokens && B > 0 && N > 0, "greedy_collapse: bad arguments");
  greedy_collapse_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(argmax, N, lengths, blank, tokens, n_tokens);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_argmax_rows(const float* x, int64_t M, int V, int32_t* argmax, void* stream) {
  LCASR_CHECK_ARG(x && argmax && M >= 0 && V > 0, "argmax_rows: bad arguments");
  if (M == 0) return 0;
  argmax_rows_kernel<<<(unsigned)M, kSmThreads, 0, (cudaStream_t)stream>>>(x, V, argmax);
  LCASR_LAUNCH_CHECK();
  return 0;
}
