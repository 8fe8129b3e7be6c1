// fp32 flash-style attention on CUDA cores: the fp32 parity mode of Attention.forward
// (attention.py:541: SDPA, scale 1/sqrt(Dh), non-causal, no mask) and the on-device check of the
// tcgen05 kernel.  One CTA = 64 query rows of one (batch, head); K/V streamed in 64-key tiles
// through shared memory; online softmax (running max / sum) in fp32.
#include "common.cuh"

namespace lcasr {

constexpr int SA_BQ = 64, SA_BK = 64, SA_THREADS = 256;

template <typename T, int DH>
__global__ void __launch_bounds__(SA_THREADS) attn_simt_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                               const T* __restrict__ v, int64_t N, int64_t Nk,
                                                               const int32_t* __restrict__ kv_len, int H,
                                                               int v_transposed, int64_t Npad, float scale,
                                                               T* __restrict__ out, int win_left, int win_right) {  // N query rows, Nk keys per batch entry
  extern __shared__ float smem[];
  constexpr int LD = DH + 1;
  float* Qs = smem;                    // [64][DH+1]
  float* Ks = Qs + SA_BQ * LD;         // [64][DH+1]
  float* Vs = Ks + SA_BK * LD;         // [64][DH]   (row = key)
  float* Ss = Vs + SA_BK * DH;         // [64][65]
  float* row_m = Ss + SA_BQ * (SA_BK + 1);  // [64]
  float* row_l = row_m + SA_BQ;             // [64]
  float* row_a = row_l + SA_BQ;             // [64] rescale factor of this step

  const int tid = threadIdx.x;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t q0 = (int64_t)blockIdx.x * SA_BQ;
  const int d = H * DH;
  const T* qb = q + (b * N) * d + h * DH;
  const T* kb = k + (b * Nk) * d + h * DH;
  const int64_t Nkv = kv_len ? min((int64_t)kv_len[b], Nk) : Nk;  // valid keys (key-padding mask of ragged batches)

  for (int idx = tid; idx < SA_BQ * DH; idx += SA_THREADS) {
    int r = idx / DH, c = idx % DH;
    int64_t n = q0 + r;
    Qs[r * LD + c] = n < N ? to_f32<T>(qb[n * d + c]) * scale : 0.f;
  }
  if (tid < SA_BQ) { row_m[tid] = -INFINITY; row_l[tid] = 0.f; }

  const int ty = tid >> 4, tx = tid & 15;      // S: rows ty*4.., cols tx*4..
  constexpr int OC = DH / 16;                  // O cols per thread: tx + 16*i
  float o_acc[4][OC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < OC; ++j) o_acc[i][j] = 0.f;

  // local attention (win_* >= 0): query i sees keys [i - win_left, i + win_right]; tiles outside the band are skipped
  const int64_t q_first = (int64_t)blockIdx.x * SA_BQ;
  const int64_t k_begin = win_left >= 0 ? (max(q_first - win_left, (int64_t)0) / SA_BK) * SA_BK : 0;
  const int64_t k_end = win_right >= 0 ? min(Nkv, q_first + SA_BQ - 1 + win_right + 1) : Nkv;
  for (int64_t k0 = k_begin; k0 < k_end; k0 += SA_BK) {
    __syncthreads();
    for (int idx = tid; idx < SA_BK * DH; idx += SA_THREADS) {
      int r = idx / DH, c = idx % DH;
      int64_t n = k0 + r;
      float kv = 0.f, vv = 0.f;
      if (n < Nkv) {
        kv = to_f32<T>(kb[n * d + c]);
        vv = v_transposed ? to_f32<T>(v[((b * H + h) * DH + c) * Npad + n]) : to_f32<T>(v[(b * Nk + n) * d + h * DH + c]);
      }
      Ks[r * LD + c] = kv;
      Vs[r * DH + c] = vv;
    }
    __syncthreads();
    // S = Q K^T (4x4 per thread)
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int c = 0; c < DH; ++c) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[(ty * 4 + i) * LD + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Ks[(tx * 4 + j) * LD + c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], bb[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t n = k0 + tx * 4 + j;
        const int64_t qi = q_first + ty * 4 + i;
        const bool in_band = (win_left < 0 || n >= qi - win_left) && (win_right < 0 || n <= qi + win_right);
        Ss[(ty * 4 + i) * (SA_BK + 1) + tx * 4 + j] = (n < Nkv && in_band) ? s[i][j] : -INFINITY;
      }
    __syncthreads();
    // online softmax: 4 threads per row, 16 columns each
    {
      int r = tid >> 2, part = tid & 3;
      float* srow = Ss + r * (SA_BK + 1) + part * 16;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) mx = fmaxf(mx, srow[j]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float m_old = row_m[r];
      float m_new = fmaxf(fmaxf(m_old, mx), -1e30f);  // stays finite while every key seen so far is masked
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float p = expf(srow[j] - m_new);
        srow[j] = p;
        sum += p;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      __syncwarp();
      if (part == 0) {
        float a = expf(m_old - m_new);  // 0 when m_old = -inf
        row_a[r] = a;
        row_l[r] = row_l[r] * a + sum;
        row_m[r] = m_new;
      }
    }
    __syncthreads();
    // O = O*alpha + P V
    float al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) al[i] = row_a[ty * 4 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < OC; ++j) o_acc[i][j] *= al[i];
    for (int kk = 0; kk < SA_BK; ++kk) {
      float p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = Ss[(ty * 4 + i) * (SA_BK + 1) + kk];
#pragma unroll
      for (int j = 0; j < OC; ++j) {
        float vv = Vs[kk * DH + tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) o_acc[i][j] = fmaf(p[i], vv, o_acc[i][j]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = ty * 4 + i;
    int64_t n = q0 + r;
    if (n >= N) continue;
    float inv = 1.0f / row_l[r];
#pragma unroll
    for (int j = 0; j < OC; ++j) out[(b * N + n) * d + h * DH + tx + 16 * j] = from_f32<T>(o_acc[i][j] * inv);
  }
}

template <typename T, int DH>
static int launch_attn_simt(const void* q, const void* k, const void* v, int B, int64_t N, int64_t Nk, const int32_t* kv_len, int H, int vt, int64_t Npad,
                            void* out, cudaStream_t st, int wl, int wr) {
  size_t smem = sizeof(float) * (2 * SA_BQ * (DH + 1) + SA_BK * DH + SA_BQ * (SA_BK + 1) + 3 * SA_BQ);
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(attn_simt_kernel<T, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.mark(attr_dev);
  }
  dim3 grid((unsigned)ceil_div(N, SA_BQ), H, B);
  float scale = 1.0f / sqrtf((float)DH);
  attn_simt_kernel<T, DH><<<grid, SA_THREADS, smem, st>>>((const T*)q, (const T*)k, (const T*)v, N, Nk, kv_len, H, vt, Npad, scale,
                                                          (T*)out, wl, wr);
  LCASR_LAUNCH_CHECK();
  return 0;
}

int attn_simt_launch(const void* q, const void* k, const void* v, int dtype, int B, int64_t N, int64_t Nk, const int32_t* kv_len, int H, int Dh,
                     int v_transposed, int64_t Npad, void* out, cudaStream_t st, int wl, int wr) {
#define LCASR_SA(DHV)                                                                                     \
  case DHV:                                                                                               \
    return dtype == LCASR_BF16 ? launch_attn_simt<bf16, DHV>(q, k, v, B, N, Nk, kv_len, H, v_transposed, Npad, out, st, wl, wr) \
                               : launch_attn_simt<float, DHV>(q, k, v, B, N, Nk, kv_len, H, v_transposed, Npad, out, st, wl, wr);
  switch (Dh) {
    LCASR_SA(32) LCASR_SA(64) LCASR_SA(128)
    default:
      return set_error(LCASR_E_UNSUPPORTED, "attention(simt): head_dim=%d not in {32,64,128}", Dh);
  }
#undef LCASR_SA
}

}  // namespace lcasr
