// fp32-accumulate SIMT GEMM  out = epi(A[M,K] . W[N,K]^T): the fp32 parity mode of the path (and
// the independent on-device check of the tcgen05 kernel).  CUDA-core FFMA only: 128x128x16 CTA
// tile, 8x8 register micro-tile per thread, operands staged through shared memory transposed to
// k-major so the inner product reads are conflict-free float4s.
#include "common.cuh"

namespace lcasr {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256;

template <typename TIn>
__device__ __forceinline__ void load4(const TIn* p, bool ok, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, bool ok, float (&v)[4]) {
  if (ok) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else { v[0] = v[1] = v[2] = v[3] = 0.f; }
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, bool ok, float (&v)[4]) {
  if (ok) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else { v[0] = v[1] = v[2] = v[3] = 0.f; }
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SG_THREADS) gemm_simt_kernel(const TIn* __restrict__ A, const TIn* __restrict__ W,
                                                               int64_t M, int N, int K, const float* __restrict__ bias,
                                                               int act, const float* resid, float alpha, TOut* out) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM;
  const int n0 = blockIdx.x * SG_BN;
  // loader mapping: 128 rows x 16 k = 512 float4 -> 2 per thread
  const int lrow = tid >> 2;         // 0..63 (+64)
  const int lk = (tid & 3) * 4;      // 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;   // 16x16 threads, each 8x8 outputs (rows ty*4+{0..3}, 64+ty*4+{0..3})
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + 64 * h;
      float va[4], vb[4];
      bool kin = (k0 + lk) < K;  // K % 4 == 0 guaranteed by the host
      load4<TIn>(A + (m0 + r) * K + k0 + lk, kin && (m0 + r) < M, va);
      load4<TIn>(W + (int64_t)(n0 + r) * K + k0 + lk, kin && (n0 + r) < N, vb);
#pragma unroll
      for (int i = 0; i < 4; ++i) { As[lk + i][r] = va[i]; Bs[lk + i][r] = vb[i]; }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[8], b[8];
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float y = acc[i][j] + (bias ? bias[n] : 0.f);
      y = apply_act(y, act);
      if (resid) y = resid[m * N + n] + alpha * y;
      out[m * N + n] = from_f32<TOut>(y);
    }
  }
}

int gemm_simt_launch(const void* A, const void* W, int ab_dtype, int64_t M, int N, int K, const float* bias, int act,
                     const float* resid, float alpha, void* out, int out_dtype, cudaStream_t st) {
  LCASR_CHECK_ARG(K % 4 == 0, "gemm(simt): K=%d must be a multiple of 4", K);
  dim3 grid((unsigned)ceil_div(N, SG_BN), (unsigned)ceil_div(M, SG_BM));
  LCASR_CHECK_ARG(grid.y <= 65535, "gemm(simt): M=%lld too large for this kernel", (long long)M);
#define LCASR_SG(TI, TO) \
  gemm_simt_kernel<TI, TO><<<grid, SG_THREADS, 0, st>>>((const TI*)A, (const TI*)W, M, N, K, bias, act, resid, alpha, (TO*)out)
  if (ab_dtype == LCASR_F32) {
    if (out_dtype == LCASR_F32) LCASR_SG(float, float); else LCASR_SG(float, bf16);
  } else {
    if (out_dtype == LCASR_F32) LCASR_SG(bf16, float); else LCASR_SG(bf16, bf16);
  }
#undef LCASR_SG
  LCASR_LAUNCH_CHECK();
  return 0;
}

}  // namespace lcasr
