// Long-form (moving-window) inference merge — lcasr/eval/utils.py:45-111 (`fetch_logits`): the recording is cut into
// windows of seq_len frames with stride seq_len - overlap, every window is run through the encoder, the class
// PROBABILITIES exp(log_probs) of overlapping windows are averaged per output frame and the logarithm is taken.
// The reference does this window by window on the host (batch 1, a [N,4096] device->host copy per window, "TODO: write
// batched version"); here all windows go through the encoder as one batch and ONE kernel does exp / sum over the covering
// windows / divide / log and the per-frame argmax, so only token ids need to leave the GPU.
// HBM-bound: reads every window row once (K*n*V*4 B), writes N_total*V*4 B (optional) + N_total*4 B.
#include "common.cuh"

namespace lcasr {

// One CTA per merged output frame p.  Windows are sorted by start position (non-decreasing); the covering set
// {k : pos_k <= p < pos_k + len_k} is found by a binary search for the last pos_k <= p and a walk to the left.
// The sum runs over k ascending — the order in which the reference accumulates its windows.
// MEAN=false is the buffered mode (lcasr/eval/buffered_transcription.py:74-90): every merged frame is covered by exactly
// one window (its central chunk) and the log-probabilities are copied bit for bit.
template <bool MEAN>
__global__ void __launch_bounds__(256) window_merge_kernel(const float* __restrict__ logp, int V, int K,
                                                           const int64_t* __restrict__ win_row0, const int32_t* __restrict__ win_len,
                                                           const int32_t* __restrict__ win_pos, int max_len,
                                                           float* __restrict__ out, int32_t* __restrict__ argmax) {
  __shared__ float red[8];
  __shared__ int red_i[8];
  const int64_t p = blockIdx.x;
  int lo = 0, hi = K - 1, last = -1;  // last window starting at or before p
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)win_pos[mid] <= p) { last = mid; lo = mid + 1; } else hi = mid - 1;
  }
  int first = last;
  while (first > 0 && (int64_t)win_pos[first - 1] + max_len > p) --first;  // candidates; coverage is re-checked below
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int c0 = threadIdx.x * 4; c0 < V; c0 += 256 * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (int k = first; k <= last; ++k) {
      const int64_t r = p - win_pos[k];
      if (r < 0 || r >= win_len[k]) continue;
      const float4 v = *reinterpret_cast<const float4*>(logp + (win_row0[k] + r) * V + c0);
      if constexpr (MEAN) {
        acc.x += expf(v.x); acc.y += expf(v.y); acc.z += expf(v.z); acc.w += expf(v.w);
      } else {
        acc = v;
      }
      ++cnt;
    }
    const float inv = cnt > 0 ? (float)cnt : 1.0f;
    float4 o;
    if constexpr (MEAN) {
      o.x = logf(acc.x / inv); o.y = logf(acc.y / inv); o.z = logf(acc.z / inv); o.w = logf(acc.w / inv);
    } else {
      o = acc;
    }
    if (out) *reinterpret_cast<float4*>(out + p * V + c0) = o;
    const float vals[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (vals[j] > mx) { mx = vals[j]; mi = c0 + j; }  // ascending scan keeps the first maximum (torch.argmax)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mx; red_i[threadIdx.x >> 5] = mi; }
  __syncthreads();
  if (threadIdx.x == 0 && argmax) {
    mx = red[0]; mi = red_i[0];
    for (int i = 1; i < 8; ++i)
      if (red[i] > mx || (red[i] == mx && red_i[i] < mi)) { mx = red[i]; mi = red_i[i]; }
    argmax[p] = mi;
  }
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_window_merge(const float* logp, int V, int K, const int64_t* win_row0, const int32_t* win_len,
                                  const int32_t* win_pos, int max_len, int64_t n_total, float* out, int32_t* argmax,
                                  void* stream) {
  LCASR_CHECK_ARG(logp && win_row0 && win_len && win_pos && (out || argmax), "window_merge: NULL argument");
  LCASR_CHECK_ARG(V > 0 && V % 4 == 0 && K > 0 && max_len > 0 && n_total > 0 && n_total < ((int64_t)1 << 31),
                  "window_merge: bad shape (V %% 4 == 0)");
  window_merge_kernel<true><<<(unsigned)n_total, 256, 0, (cudaStream_t)stream>>>(logp, V, K, win_row0, win_len, win_pos, max_len,
                                                                                out, argmax);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_window_concat(const float* logp, int V, int K, const int64_t* win_row0, const int32_t* win_len,
                                   const int32_t* win_pos, int max_len, int64_t n_total, float* out, int32_t* argmax,
                                   void* stream) {
  LCASR_CHECK_ARG(logp && win_row0 && win_len && win_pos && (out || argmax), "window_concat: NULL argument");
  LCASR_CHECK_ARG(V > 0 && V % 4 == 0 && K > 0 && max_len > 0 && n_total > 0 && n_total < ((int64_t)1 << 31),
                  "window_concat: bad shape (V %% 4 == 0)");
  window_merge_kernel<false><<<(unsigned)n_total, 256, 0, (cudaStream_t)stream>>>(logp, V, K, win_row0, win_len, win_pos, max_len,
                                                                                 out, argmax);
  LCASR_LAUNCH_CHECK();
  return 0;
}
