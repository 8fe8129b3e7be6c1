// Memory-bound kernels of the TRAINING step (cfg 5: forward + backward of the encoder):
// the backward of LayerNorm / RMSNorm, GLU, rotary, softmax / log-softmax, the depthwise Conv1d, BatchRenorm in
// training mode (batchrenorm.py:52-84) and the reductions that produce bias / affine-parameter gradients.
// Activations are bf16 (the reference trains under bf16 autocast, exp/train.py:225), the residual stream, its
// gradient and every parameter gradient are fp32.  All kernels are HBM-bound; each thread moves 16-byte vectors
// and rows are contiguous, reductions use warp shuffles + shared-memory atomics + one global atomic per CTA.
#include "common.cuh"
#include <cstdlib>

namespace lcasr {

__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanh_approx_f(u));
}
__device__ __forceinline__ float silu_grad_f(float x) {
  const float s = sigmoid_fast(x);
  return s * (1.0f + x * (1.0f - s));
}

static inline unsigned grid_1d(int64_t total, int block) {
  int64_t g = ceil_div(total, block);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---- out(bf16) = scale * in(fp32) : the cast in front of the first backward GEMM of a sub-layer --------------
__global__ void __launch_bounds__(256) scale_cast_kernel(const float* __restrict__ in, int64_t n8, float scale,
                                                         bf16* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    Vec8<float>::load(in + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= scale;
    Vec8<bf16>::store(out + i * 8, v);
  }
}

// ---- out = act(in), bf16 (training keeps the pre-activation for the backward) -------------------------------
__global__ void __launch_bounds__(256) act_fwd_kernel(const bf16* __restrict__ in, int64_t n8, int act,
                                                      bf16* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    Vec8<bf16>::load(in + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act == LCASR_ACT_GELU_TANH ? gelu_tanh_fast(v[j]) : silu_fast(v[j]);
    Vec8<bf16>::store(out + i * 8, v);
  }
}

// ---- dpre = dy * act'(pre), bf16 ------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_tanh_grad_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float t = tanh_approx_f(k0 * (x + k1 * x * x * x));
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x * x);
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const bf16* __restrict__ pre, const bf16* __restrict__ dy, int64_t n8,
                                                      int act, bf16* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float x[8], g[8];
    Vec8<bf16>::load(pre + i * 8, x);
    Vec8<bf16>::load(dy + i * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= act == LCASR_ACT_GELU_TANH ? gelu_tanh_grad_f(x[j]) : silu_grad_f(x[j]);
    Vec8<bf16>::store(out + i * 8, g);
  }
}

// ---- GLU backward (convolution.py:107): u = [a | b], g = a*sigmoid(b) ----------------------------------------
__global__ void __launch_bounds__(256) glu_bwd_kernel(const bf16* __restrict__ u, const bf16* __restrict__ dg, int64_t M,
                                                      int d, bf16* __restrict__ du) {
  const int vpr = d / 8;
  const int64_t total = M * vpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = idx / vpr;
    const int j = (int)(idx % vpr) * 8;
    float a[8], b[8], g[8], da[8], db[8];
    Vec8<bf16>::load(u + m * 2 * d + j, a);
    Vec8<bf16>::load(u + m * 2 * d + d + j, b);
    Vec8<bf16>::load(dg + m * d + j, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float s = sigmoid_fast(b[i]);
      da[i] = g[i] * s;
      db[i] = g[i] * a[i] * s * (1.0f - s);
    }
    Vec8<bf16>::store(du + m * 2 * d + j, da);
    Vec8<bf16>::store(du + m * 2 * d + d + j, db);
  }
}

// ---- rotary backward + merge: dq,dk (w.r.t. the ROTATED q,k), dv [B,N,H,Dh] -> dqkv [M, 3d] = [dq|dk|dv] -----
// the transpose of the rotation by +theta is the rotation by -theta (rotary_emb.py:61-73)
__global__ void __launch_bounds__(256) rope_bwd_merge_kernel(const bf16* __restrict__ dq, const bf16* __restrict__ dk,
                                                             const bf16* __restrict__ dv, int64_t N, int H, int Dh,
                                                             const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                             int64_t total, bf16* __restrict__ dqkv) {
  const int half = Dh / 2, vph = half / 8, d = H * Dh;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int jv = (int)(idx % vph);
    int64_t r = idx / vph;
    const int h = (int)(r % H);
    const int64_t m = r / H, n = m % N;
    const int j = jv * 8;
    float c[8], s[8];
    if (cos_t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = cos_t[n * half + j + i]; s[i] = sin_t[n * half + j + i]; }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = 1.f; s[i] = 0.f; }
    }
    bf16* row = dqkv + m * 3 * d;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const bf16* src = (which == 0 ? dq : dk) + m * d + h * Dh;
      float lo[8], hi[8], olo[8], ohi[8];
      Vec8<bf16>::load(src + j, lo);
      Vec8<bf16>::load(src + half + j, hi);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        olo[i] = lo[i] * c[i] + hi[i] * s[i];
        ohi[i] = hi[i] * c[i] - lo[i] * s[i];
      }
      Vec8<bf16>::store(row + which * d + h * Dh + j, olo);
      Vec8<bf16>::store(row + which * d + h * Dh + half + j, ohi);
    }
    float vlo[8], vhi[8];
    Vec8<bf16>::load(dv + m * d + h * Dh + j, vlo);
    Vec8<bf16>::load(dv + m * d + h * Dh + half + j, vhi);
    Vec8<bf16>::store(row + 2 * d + h * Dh + j, vlo);
    Vec8<bf16>::store(row + 2 * d + h * Dh + half + j, vhi);
  }
}

// ---- D[b,h,n] = sum_dh dO * O  (the softmax-backward row term of attention) ----------------------------------
__global__ void __launch_bounds__(256) rowdot_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, int64_t N,
                                                     int H, int Dh, int64_t total, float* __restrict__ out) {
  const int tpr = Dh / 8;  // threads per (token, head): a power of two <= 32
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = idx < total;
  float acc = 0.f;
  int64_t r = 0;
  if (ok) {
    r = idx / tpr;
    const int j = (int)(idx % tpr) * 8;
    float x[8], y[8];
    Vec8<bf16>::load(a + r * Dh + j, x);
    Vec8<bf16>::load(b + r * Dh + j, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(x[i], y[i], acc);
  }
  for (int o = tpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (ok && (idx % tpr) == 0) {
    const int h = (int)(r % H);
    const int64_t m = r / H, n = m % N, bb = m / N;
    out[(bb * H + h) * N + n] = acc;
  }
}

// ---- softmax backward (self-conditioning, sconformer_xl.py:242): dl = p o (dp - sum(p o dp)) -----------------
// ---- log-softmax backward (decoder.py:29): dl = scale * (dlp - exp(lp) * sum(dlp)) ----------------------------
template <int MODE>  // 0: softmax (p, dp bf16)   1: log-softmax (lp, dlp fp32)
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const void* __restrict__ y_, const void* __restrict__ dy_, int V,
                                                          float scale, bf16* __restrict__ dl) {
  __shared__ float red[8];
  const int64_t row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float acc = 0.f;
  if (MODE == 0) {
    const bf16* y = (const bf16*)y_ + row * V;
    const bf16* dy = (const bf16*)dy_ + row * V;
    for (int i = tid * 8; i < V; i += 256 * 8) {
      float a[8], b[8];
      Vec8<bf16>::load(y + i, a);
      Vec8<bf16>::load(dy + i, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(a[k], b[k], acc);
    }
  } else {
    const float* dy = (const float*)dy_ + row * V;
    for (int i = tid * 8; i < V; i += 256 * 8) {
      float b[8];
      Vec8<float>::load(dy + i, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += b[k];
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  if (MODE == 0) {
    const bf16* y = (const bf16*)y_ + row * V;
    const bf16* dy = (const bf16*)dy_ + row * V;
    for (int i = tid * 8; i < V; i += 256 * 8) {
      float a[8], b[8], o[8];
      Vec8<bf16>::load(y + i, a);
      Vec8<bf16>::load(dy + i, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = scale * a[k] * (b[k] - tot);
      Vec8<bf16>::store(dl + row * V + i, o);
    }
  } else {
    const float* y = (const float*)y_ + row * V;
    const float* dy = (const float*)dy_ + row * V;
    for (int i = tid * 8; i < V; i += 256 * 8) {
      float a[8], b[8], o[8];
      Vec8<float>::load(y + i, a);
      Vec8<float>::load(dy + i, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = scale * (b[k] - __expf(a[k]) * tot);
      Vec8<bf16>::store(dl + row * V + i, o);
    }
  }
}

// ---- column sums: out[d] += scale * sum_m in[m, :]  (bias gradients) -----------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ in, int64_t M, int d, int rows_per_cta,
                                                     float scale, float* __restrict__ out) {
  __shared__ float red[8][32][8];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int cg = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (cg * 8 < d) {
    int64_t r = r0 + ry;
    for (; r + 24 < r1; r += 32) {  // four independent 16-byte loads in flight per thread
      float v0[8], v1[8], v2[8], v3[8];
      Vec8<T>::load(in + r * d + cg * 8, v0);
      Vec8<T>::load(in + (r + 8) * d + cg * 8, v1);
      Vec8<T>::load(in + (r + 16) * d + cg * 8, v2);
      Vec8<T>::load(in + (r + 24) * d + cg * 8, v3);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
    }
    for (; r < r1; r += 8) {
      float v[8];
      Vec8<T>::load(in + r * d + cg * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ry][cx][i] = acc[i];
  __syncthreads();
  if (ry == 0 && cg * 8 < d) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += red[y][cx][i];
      atomicAdd(out + cg * 8 + i, scale * s);
    }
  }
}

// ---- LayerNorm / lcasr-RMSNorm backward ----------------------------------------------------------------------
// x fp32 [M,d] (the saved input), dy [M,d] (bf16 or fp32), w [d].  dx (fp32) = or += the input gradient;
// dw[d], db[d] += parameter gradients.  One warp per row (grid-stride), statistics recomputed from x.
template <typename TDy>
__global__ void __launch_bounds__(256) norm_bwd_kernel(const float* __restrict__ x, const TDy* __restrict__ dy,
                                                       const float* __restrict__ w, int64_t M, int d, float eps, int kind,
                                                       int accumulate, float* __restrict__ dx, float* __restrict__ dw,
                                                       float* __restrict__ db, bf16* __restrict__ cast_out, float cast_scale) {
  // cast_out (may be NULL): bf16 copy of cast_scale * (the new dx) — the dY operand of the NEXT sub-layer's backward GEMMs
  extern __shared__ float sacc[];  // [2][d]
  float* sdw = sacc;
  float* sdb = sacc + d;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (int64_t)gridDim.x * wpb) {
    const float* xr = x + row * d;
    const TDy* gr = dy + row * d;
    float* dxr = dx + row * d;
    if (kind == LCASR_NORM_LAYERNORM) {
      float s = 0.f;
      for (int i = lane; i < d; i += 32) s += xr[i];
      const float mean = warp_sum(s) / d;
      float q = 0.f;
      for (int i = lane; i < d; i += 32) { const float a = xr[i] - mean; q += a * a; }
      const float rstd = rsqrtf(warp_sum(q) / d + eps);
      float s1 = 0.f, s2 = 0.f;  // sum(g), sum(g*xhat), g = dy*w
      for (int i = lane; i < d; i += 32) {
        const float g = to_f32<TDy>(gr[i]) * w[i];
        s1 += g;
        s2 += g * (xr[i] - mean) * rstd;
      }
      s1 = warp_sum(s1) / d;
      s2 = warp_sum(s2) / d;
      for (int i = lane; i < d; i += 32) {
        const float gy = to_f32<TDy>(gr[i]);
        const float xh = (xr[i] - mean) * rstd;
        float v = rstd * (gy * w[i] - s1 - xh * s2);
        if (accumulate) v += dxr[i];
        dxr[i] = v;
        if (cast_out) cast_out[row * d + i] = __float2bfloat16_rn(cast_scale * v);
        atomicAdd(sdw + i, gy * xh);
        atomicAdd(sdb + i, gy);
      }
    } else {  // y = w * x / (rms + eps), rms = ||x|| / sqrt(d)   (normalisation.py:30-47)
      float q = 0.f;
      for (int i = lane; i < d; i += 32) q += xr[i] * xr[i];
      const float rms = sqrtf(warp_sum(q)) * rsqrtf((float)d);
      const float inv = 1.0f / (rms + eps);
      float s2 = 0.f;  // sum(g * x)
      for (int i = lane; i < d; i += 32) s2 += to_f32<TDy>(gr[i]) * w[i] * xr[i];
      s2 = warp_sum(s2);
      const float k = rms > 0.f ? s2 * inv * inv / ((float)d * rms) : 0.f;
      for (int i = lane; i < d; i += 32) {
        const float gy = to_f32<TDy>(gr[i]);
        float v = gy * w[i] * inv - k * xr[i];
        if (accumulate) v += dxr[i];
        dxr[i] = v;
        if (cast_out) cast_out[row * d + i] = __float2bfloat16_rn(cast_scale * v);
        atomicAdd(sdw + i, gy * xr[i] * inv);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(dw + i, sdw[i]);
    if (db) atomicAdd(db + i, sdb[i]);
  }
}

// d == NE*32 (NE a multiple of 4) variant: lane l owns columns {128*j + 4*l .. +3}, j < NE/4, for EVERY row it
// visits, so the row (x, dy) lives in registers (128-bit loads), and the parameter-gradient partial sums are
// per-lane register accumulators reduced through shared memory once per CTA (the generic kernel above spends its
// time in 2*d shared-memory atomics per row).  HBM-bound: reads x (4 B) + dy (e) and reads+writes dx (8 B)
// per element.
template <int NE, typename TDy>
__global__ void __launch_bounds__(256, 2) norm_bwd_reg_kernel(const float* __restrict__ x, const TDy* __restrict__ dy,
                                                           const float* __restrict__ w, int64_t M, float eps, int kind,
                                                           int accumulate, float* __restrict__ dx, float* __restrict__ dw,
                                                           float* __restrict__ db, bf16* __restrict__ cast_out, float cast_scale) {
  constexpr int d = NE * 32, NV = NE / 4;
  extern __shared__ float sacc[];  // [2][d]
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  // (the weight vector is re-read from L1 where needed instead of living in registers: 2 CTAs per SM need <= 128)
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4 aw[NV], ab[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    aw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += (int64_t)gridDim.x * wpb) {
    // The row's three phases (statistics of x, the two dy reductions, the output pass) each wait for their own loads and the
    // registers are full (128, spilling): 16 warps per SM keep too few bytes in flight for HBM (0.58 of the copy rate).
    // The NEXT row of this warp is therefore pulled into L2 now — a hint, no registers: its loads then cost L2 latency.
    {
      const int64_t nrow = row + (int64_t)gridDim.x * wpb;
      if (nrow < M) {
        if (lane < NE) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(x + nrow * d) + lane * 128));
          if (accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(dx + nrow * d) + lane * 128));
        }
        if (lane < (NE * (int)sizeof(TDy) + 3) / 4)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(dy + nrow * d) + lane * 128));
      }
    }
    // dy is read twice (reduction pass, output pass): the second read hits L1 / L2 and saves NV float4 registers
    auto load_g = [&](int j) -> float4 {
      if constexpr (sizeof(TDy) == 4) {
        return reinterpret_cast<const float4*>(dy + row * d)[lane + 32 * j];
      } else {
        const uint2 u = reinterpret_cast<const uint2*>(dy + row * d)[lane + 32 * j];
        const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        return make_float4(f0.x, f0.y, f1.x, f1.y);
      }
    };
    float4 xv[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) xv[j] = reinterpret_cast<const float4*>(x + row * d)[lane + 32 * j];
    float4* dxr = reinterpret_cast<float4*>(dx + row * d);
    if (kind == LCASR_NORM_LAYERNORM) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) s += (xv[j].x + xv[j].y) + (xv[j].z + xv[j].w);
      const float mean = warp_sum(s) * (1.0f / d);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        xv[j].x -= mean; xv[j].y -= mean; xv[j].z -= mean; xv[j].w -= mean;
        q += (xv[j].x * xv[j].x + xv[j].y * xv[j].y) + (xv[j].z * xv[j].z + xv[j].w * xv[j].w);
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / d) + eps);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {  // xv becomes xhat
        xv[j].x *= rstd; xv[j].y *= rstd; xv[j].z *= rstd; xv[j].w *= rstd;
        const float4 wj = __ldg(w4 + lane + 32 * j);
        const float4 gj = load_g(j);
        const float g0 = gj.x * wj.x, g1 = gj.y * wj.y, g2 = gj.z * wj.z, g3 = gj.w * wj.w;
        s1 += (g0 + g1) + (g2 + g3);
        s2 += (g0 * xv[j].x + g1 * xv[j].y) + (g2 * xv[j].z + g3 * xv[j].w);
      }
      s1 = warp_sum(s1) * (1.0f / d);
      s2 = warp_sum(s2) * (1.0f / d);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 wj = __ldg(w4 + lane + 32 * j);
        const float4 gj = load_g(j);
        float4 v;
        v.x = rstd * (gj.x * wj.x - s1 - xv[j].x * s2);
        v.y = rstd * (gj.y * wj.y - s1 - xv[j].y * s2);
        v.z = rstd * (gj.z * wj.z - s1 - xv[j].z * s2);
        v.w = rstd * (gj.w * wj.w - s1 - xv[j].w * s2);
        if (accumulate) {
          const float4 o = dxr[lane + 32 * j];
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        dxr[lane + 32 * j] = v;
        if (cast_out) {
          __nv_bfloat162 c0 = __floats2bfloat162_rn(cast_scale * v.x, cast_scale * v.y), c1 = __floats2bfloat162_rn(cast_scale * v.z, cast_scale * v.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&c0); u.y = *reinterpret_cast<uint32_t*>(&c1);
          reinterpret_cast<uint2*>(cast_out + row * d)[lane + 32 * j] = u;
        }
        aw[j].x = fmaf(gj.x, xv[j].x, aw[j].x); aw[j].y = fmaf(gj.y, xv[j].y, aw[j].y);
        aw[j].z = fmaf(gj.z, xv[j].z, aw[j].z); aw[j].w = fmaf(gj.w, xv[j].w, aw[j].w);
        ab[j].x += gj.x; ab[j].y += gj.y; ab[j].z += gj.z; ab[j].w += gj.w;
      }
    } else {
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) q += (xv[j].x * xv[j].x + xv[j].y * xv[j].y) + (xv[j].z * xv[j].z + xv[j].w * xv[j].w);
      const float rms = sqrtf(warp_sum(q)) * rsqrtf((float)d);
      const float inv = 1.0f / (rms + eps);
      float s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 wj = __ldg(w4 + lane + 32 * j);
        const float4 gj = load_g(j);
        s2 += (gj.x * wj.x * xv[j].x + gj.y * wj.y * xv[j].y) + (gj.z * wj.z * xv[j].z + gj.w * wj.w * xv[j].w);
      }
      s2 = warp_sum(s2);
      const float k = rms > 0.f ? s2 * inv * inv / ((float)d * rms) : 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 wj = __ldg(w4 + lane + 32 * j);
        const float4 gj = load_g(j);
        float4 v;
        v.x = gj.x * wj.x * inv - k * xv[j].x;
        v.y = gj.y * wj.y * inv - k * xv[j].y;
        v.z = gj.z * wj.z * inv - k * xv[j].z;
        v.w = gj.w * wj.w * inv - k * xv[j].w;
        if (accumulate) {
          const float4 o = dxr[lane + 32 * j];
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        dxr[lane + 32 * j] = v;
        if (cast_out) {
          __nv_bfloat162 c0 = __floats2bfloat162_rn(cast_scale * v.x, cast_scale * v.y), c1 = __floats2bfloat162_rn(cast_scale * v.z, cast_scale * v.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&c0); u.y = *reinterpret_cast<uint32_t*>(&c1);
          reinterpret_cast<uint2*>(cast_out + row * d)[lane + 32 * j] = u;
        }
        aw[j].x = fmaf(gj.x, xv[j].x * inv, aw[j].x); aw[j].y = fmaf(gj.y, xv[j].y * inv, aw[j].y);
        aw[j].z = fmaf(gj.z, xv[j].z * inv, aw[j].z); aw[j].w = fmaf(gj.w, xv[j].w * inv, aw[j].w);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (lane + 32 * j) * 4;
    atomicAdd(sacc + c, aw[j].x); atomicAdd(sacc + c + 1, aw[j].y); atomicAdd(sacc + c + 2, aw[j].z); atomicAdd(sacc + c + 3, aw[j].w);
    atomicAdd(sacc + d + c, ab[j].x); atomicAdd(sacc + d + c + 1, ab[j].y); atomicAdd(sacc + d + c + 2, ab[j].z);
    atomicAdd(sacc + d + c + 3, ab[j].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(dw + i, sacc[i]);
    if (db) atomicAdd(db + i, sacc[d + i]);
  }
}

// ---- depthwise Conv1d over tokens, channels-last [B,N,d] bf16, register sliding window ------------------------
// MODE 0: out = conv(in, w) + b, and (if sum != NULL) per-channel sum / sum of squares of out (BatchRenorm
//         training statistics, batchrenorm.py:67-68)
// MODE 1: data gradient: out[n] = sum_k in[n + PAD - k] * w[k]   (the same conv with the taps flipped, no bias)
// MODE 2: weight / bias gradient: dw[c,k] += sum_n dout[n] * x[n - PAD + k], db[c] += sum_n dout[n]
//         (in = x, in2 = dout)
template <int KS, int MODE>
__global__ void __launch_bounds__(128) dwconv1d_kernel(const bf16* __restrict__ in, const bf16* __restrict__ in2, int64_t N,
                                                       int d, int tt, const float* __restrict__ w, const float* __restrict__ b,
                                                       bf16* __restrict__ out, float* __restrict__ acc0, float* __restrict__ acc1) {
  constexpr int PAD = (KS - 1) / 2;
  const int cgroups = d / 8;
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg >= cgroups) return;
  const int64_t n0 = (int64_t)blockIdx.y * tt;
  const int64_t batch = blockIdx.z;
  const int c0 = cg * 8;
  float wr[8][KS], bias[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      if (MODE == 0) wr[c][k] = w[(c0 + c) * KS + k];
      else if (MODE == 1) wr[c][k] = w[(c0 + c) * KS + (KS - 1 - k)];
      else wr[c][k] = 0.f;  // MODE 2: accumulators
    }
    bias[c] = MODE == 0 ? b[c0 + c] : 0.f;
  }
  const bf16* base = in + batch * N * d + c0;
  float win[KS][8];
#pragma unroll
  for (int j = 0; j < KS - 1; ++j) {
    const int64_t n = n0 - PAD + j;
    if (n >= 0 && n < N) Vec8<bf16>::load(base + n * d, win[j + 1]);
    else {
#pragma unroll
      for (int c = 0; c < 8; ++c) win[j + 1][c] = 0.f;
    }
  }
  float s1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, s2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t = 0; t < tt; ++t) {
    const int64_t n = n0 + t;
    if (n >= N) break;
#pragma unroll
    for (int j = 0; j < KS - 1; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) win[j][c] = win[j + 1][c];
    const int64_t nn = n + PAD;
    if (nn < N) Vec8<bf16>::load(base + nn * d, win[KS - 1]);
    else {
#pragma unroll
      for (int c = 0; c < 8; ++c) win[KS - 1][c] = 0.f;
    }
    if (MODE == 2) {
      float g[8];
      Vec8<bf16>::load(in2 + (batch * N + n) * d + c0, g);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int k = 0; k < KS; ++k) wr[c][k] = fmaf(g[c], win[k][c], wr[c][k]);
        s1[c] += g[c];
      }
    } else {
      float y[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float a = bias[c];
#pragma unroll
        for (int k = 0; k < KS; ++k) a = fmaf(wr[c][k], win[k][c], a);
        y[c] = a;
      }
      Vec8<bf16>::store(out + (batch * N + n) * d + c0, y);
      if (MODE == 0 && acc0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // statistics of the ROUNDED activation (what the next kernels read)
          const float yr = __bfloat162float(__float2bfloat16_rn(y[c]));
          s1[c] += yr;
          s2[c] = fmaf(yr, yr, s2[c]);
        }
      }
    }
  }
  if (MODE == 0 && acc0) {  // cross-CTA sums in fp64: the order in which the atomics land no longer reaches the fp32 result
    double* q0 = reinterpret_cast<double*>(acc0);
    double* q1 = reinterpret_cast<double*>(acc1);
#pragma unroll
    for (int c = 0; c < 8; ++c) { atomicAdd(q0 + c0 + c, (double)s1[c]); atomicAdd(q1 + c0 + c, (double)s2[c]); }
  }
  if (MODE == 2) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int k = 0; k < KS; ++k) atomicAdd(acc0 + (c0 + c) * KS + k, wr[c][k]);
      atomicAdd(acc1 + c0 + c, s1[c]);
    }
  }
}

// ---- BatchRenorm1d, training mode (batchrenorm.py:52-84) ------------------------------------------------------
// From the per-channel sum / sum-of-squares of the depthwise-conv output c over all B*N tokens:
//   mu, sigma = std(unbiased=False) + eps, r = clamp(sigma / running_std, 1/rmax, rmax),
//   dd = clamp((mu - running_mean) / running_std, -dmax, dmax)            (r, dd are constants for autograd)
//   z = weight * ((c - mu) / sigma * r + dd) + bias = c * A + Bc
// and the running statistics are updated in place (momentum), exactly like the reference module.
// stats[5][d] receives mu, sigma, r, dd, s (= sigma - eps) for the backward.
__global__ void brn_train_stats_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, float count, int d,
                                       float* __restrict__ running_mean, float* __restrict__ running_std, float eps,
                                       float rmax, float dmax, float momentum, const float* __restrict__ weight,
                                       const float* __restrict__ bias, float* __restrict__ A, float* __restrict__ Bc,
                                       float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float mu = (float)sum[c] / count;
  const float var = fmaxf((float)sumsq[c] / count - mu * mu, 0.f);
  const float s = sqrtf(var);
  const float sigma = s + eps;
  const float rs = running_std[c], rm = running_mean[c];
  const float r = fminf(fmaxf(sigma / rs, 1.0f / rmax), rmax);
  const float dd = fminf(fmaxf((mu - rm) / rs, -dmax), dmax);
  const float a = weight[c] * r / sigma;
  A[c] = a;
  Bc[c] = weight[c] * dd - a * mu + bias[c];
  stats[c] = mu; stats[d + c] = sigma; stats[2 * d + c] = r; stats[3 * d + c] = dd; stats[4 * d + c] = s;
  running_mean[c] = rm + momentum * (mu - rm);
  running_std[c] = rs + momentum * (sigma - rs);
}

// y = silu(c * A + Bc)
__global__ void __launch_bounds__(256) affine_silu_kernel(const bf16* __restrict__ cin, int64_t M, int d,
                                                          const float* __restrict__ A, const float* __restrict__ Bc,
                                                          bf16* __restrict__ out) {
  const int vpr = d / 8;
  const int64_t total = M * vpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % vpr) * 8;
    float v[8];
    Vec8<bf16>::load(cin + idx * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = silu_fast(fmaf(v[i], A[j + i], Bc[j + i]));
    Vec8<bf16>::store(out + idx * 8, v);
  }
}

// backward pass 1: dz = dy * silu'(c*A + Bc) (stored bf16), S1[d] += sum dz, S2[d] += sum dz * xhat,
// xhat = (c - mu) / sigma.  Thread layout like colsum_kernel.
__global__ void __launch_bounds__(256) affine_silu_bwd_kernel(const bf16* __restrict__ cin, const bf16* __restrict__ dy,
                                                              int64_t M, int d, int rows_per_cta, const float* __restrict__ A,
                                                              const float* __restrict__ Bc, const float* __restrict__ stats,
                                                              bf16* __restrict__ dz, double* __restrict__ S1,
                                                              double* __restrict__ S2) {
  __shared__ float red[8][32][16];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int cg = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  float a1[8], a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a1[i] = a2[i] = 0.f;
  if (cg * 8 < d) {
    float Av[8], Bv[8], mu[8], isg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      Av[i] = A[cg * 8 + i]; Bv[i] = Bc[cg * 8 + i];
      mu[i] = stats[cg * 8 + i]; isg[i] = 1.0f / stats[d + cg * 8 + i];
    }
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      float c[8], g[8], o[8];
      Vec8<bf16>::load(cin + r * d + cg * 8, c);
      Vec8<bf16>::load(dy + r * d + cg * 8, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = fmaf(c[i], Av[i], Bv[i]);
        const float v = g[i] * silu_grad_f(z);
        o[i] = v;
        const float vr = __bfloat162float(__float2bfloat16_rn(v));
        a1[i] += vr;
        a2[i] = fmaf(vr, (c[i] - mu[i]) * isg[i], a2[i]);
      }
      Vec8<bf16>::store(dz + r * d + cg * 8, o);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[ry][cx][i] = a1[i]; red[ry][cx][8 + i] = a2[i]; }
  __syncthreads();
  if (ry == 0 && cg * 8 < d) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f, t = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) { s += red[y][cx][i]; t += red[y][cx][8 + i]; }
      atomicAdd(S1 + cg * 8 + i, (double)s);  // fp64 across CTAs: deterministic fp32 coefficients downstream
      atomicAdd(S2 + cg * 8 + i, (double)t);
    }
  }
}

// backward finalize (per channel): gradients of the BatchRenorm affine parameters and the three coefficient
// vectors of  dc = k1*dz + k2*c + k3 :
//   dxhat = dz*w*r ; dc = (1/sigma) [dxhat - mean(dxhat) - xhat * (sigma/s) * mean(dxhat*xhat)]
__global__ void brn_bwd_finalize_kernel(const double* __restrict__ S1, const double* __restrict__ S2, float count, int d,
                                        const float* __restrict__ weight, const float* __restrict__ stats,
                                        float* __restrict__ dweight, float* __restrict__ dbias, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float mu = stats[c], sigma = stats[d + c], r = stats[2 * d + c], dd = stats[3 * d + c], s = stats[4 * d + c];
  const float s1 = (float)S1[c], s2 = (float)S2[c];
  atomicAdd(dweight + c, r * s2 + dd * s1);  // sum dz * (r*xhat + dd)
  atomicAdd(dbias + c, s1);
  const float k = weight[c] * r / sigma;
  const float q = s > 0.f ? k * (s2 / count) / s : 0.f;
  coef[c] = k;
  coef[d + c] = -q;
  coef[2 * d + c] = q * mu - k * s1 / count;
}

// out = k1*a + k2*b + k3 per channel (the BatchRenorm input gradient)
__global__ void __launch_bounds__(256) affine3_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, int64_t M, int d,
                                                      const float* __restrict__ coef, bf16* __restrict__ out) {
  const int vpr = d / 8;
  const int64_t total = M * vpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % vpr) * 8;
    float x[8], y[8], o[8];
    Vec8<bf16>::load(a + idx * 8, x);
    Vec8<bf16>::load(b + idx * 8, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(coef[j + i], x[i], fmaf(coef[d + j + i], y[i], coef[2 * d + j + i]));
    Vec8<bf16>::store(out + idx * 8, o);
  }
}

// x(fp32) += a(bf16)   (gradient joins of the residual stream)
__global__ void __launch_bounds__(256) add_bf16_kernel(float* __restrict__ x, const bf16* __restrict__ a, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8], w[8];
    Vec8<float>::load(x + i * 8, v);
    Vec8<bf16>::load(a + i * 8, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += w[j];
    Vec8<float>::store(x + i * 8, v);
  }
}

}  // namespace lcasr

using namespace lcasr;

#define ST ((cudaStream_t)stream)

extern "C" int lcasr_scale_cast(const float* in, int64_t n, float scale, void* out, void* stream) {
  LCASR_CHECK_ARG(in && out && n >= 0 && n % 8 == 0, "scale_cast: n must be a multiple of 8");
  if (n == 0) return 0;
  scale_cast_kernel<<<grid_1d(n / 8, 256), 256, 0, ST>>>(in, n / 8, scale, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_act_fwd(const void* in, int64_t n, int act, void* out, void* stream) {
  LCASR_CHECK_ARG(in && out && n >= 0 && n % 8 == 0, "act_fwd: n must be a multiple of 8");
  LCASR_CHECK_ARG(act == LCASR_ACT_GELU_TANH || act == LCASR_ACT_SILU, "act_fwd: bad activation %d", act);
  if (n == 0) return 0;
  act_fwd_kernel<<<grid_1d(n / 8, 256), 256, 0, ST>>>((const bf16*)in, n / 8, act, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_act_bwd(const void* pre, const void* dy, int64_t n, int act, void* out, void* stream) {
  LCASR_CHECK_ARG(pre && dy && out && n >= 0 && n % 8 == 0, "act_bwd: n must be a multiple of 8");
  LCASR_CHECK_ARG(act == LCASR_ACT_GELU_TANH || act == LCASR_ACT_SILU, "act_bwd: bad activation %d", act);
  if (n == 0) return 0;
  act_bwd_kernel<<<grid_1d(n / 8, 256), 256, 0, ST>>>((const bf16*)pre, (const bf16*)dy, n / 8, act, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_glu_bwd(const void* u, const void* dg, int64_t M, int d, void* du, void* stream) {
  LCASR_CHECK_ARG(u && dg && du && M > 0 && d > 0 && d % 8 == 0, "glu_bwd: bad arguments");
  glu_bwd_kernel<<<grid_1d(M * (d / 8), 256), 256, 0, ST>>>((const bf16*)u, (const bf16*)dg, M, d, (bf16*)du);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_rope_bwd_merge(const void* dq, const void* dk, const void* dv, int B, int64_t N, int H, int Dh,
                                    const float* cos_t, const float* sin_t, void* dqkv, void* stream) {
  LCASR_CHECK_ARG(dq && dk && dv && dqkv && B > 0 && N > 0 && H > 0 && Dh % 16 == 0, "rope_bwd_merge: bad arguments");
  LCASR_CHECK_ARG((cos_t == nullptr) == (sin_t == nullptr), "rope_bwd_merge: cos and sin come together");
  const int64_t total = (int64_t)B * N * H * (Dh / 16);
  rope_bwd_merge_kernel<<<grid_1d(total, 256), 256, 0, ST>>>((const bf16*)dq, (const bf16*)dk, (const bf16*)dv, N, H, Dh,
                                                             cos_t, sin_t, total, (bf16*)dqkv);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_rowdot(const void* a, const void* b, int B, int64_t N, int H, int Dh, float* out, void* stream) {
  LCASR_CHECK_ARG(a && b && out && B > 0 && N > 0 && H > 0, "rowdot: bad arguments");
  LCASR_CHECK_ARG(Dh == 8 || Dh == 16 || Dh == 32 || Dh == 64 || Dh == 128 || Dh == 256, "rowdot: Dh=%d unsupported", Dh);
  const int64_t total = (int64_t)B * N * H * (Dh / 8);
  rowdot_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, ST>>>((const bf16*)a, (const bf16*)b, N, H, Dh, total, out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_softmax_bwd(const void* p, const void* dp, int64_t M, int V, float scale, void* dl, void* stream) {
  LCASR_CHECK_ARG(p && dp && dl && M > 0 && V > 0 && V % 8 == 0, "softmax_bwd: bad arguments (V %% 8 == 0)");
  LCASR_CHECK_ARG(M < ((int64_t)1 << 31), "softmax_bwd: too many rows");
  softmax_bwd_kernel<0><<<(unsigned)M, 256, 0, ST>>>(p, dp, V, scale, (bf16*)dl);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_log_softmax_bwd(const float* lp, const float* dlp, int64_t M, int V, float scale, void* dl,
                                     void* stream) {
  LCASR_CHECK_ARG(lp && dlp && dl && M > 0 && V > 0 && V % 8 == 0, "log_softmax_bwd: bad arguments (V %% 8 == 0)");
  LCASR_CHECK_ARG(M < ((int64_t)1 << 31), "log_softmax_bwd: too many rows");
  softmax_bwd_kernel<1><<<(unsigned)M, 256, 0, ST>>>(lp, dlp, V, scale, (bf16*)dl);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_colsum(const void* in, int dtype, int64_t M, int d, float scale, float* out, void* stream) {
  LCASR_CHECK_ARG(in && out && M > 0 && d > 0 && d % 8 == 0, "colsum: bad arguments (d %% 8 == 0)");
  const int rows_per_cta = 256;
  dim3 grid((unsigned)ceil_div(d / 8, 32), (unsigned)ceil_div(M, rows_per_cta));
  LCASR_CHECK_ARG(grid.y <= 65535, "colsum: too many rows");
  if (dtype == LCASR_BF16) colsum_kernel<bf16><<<grid, 256, 0, ST>>>((const bf16*)in, M, d, rows_per_cta, scale, out);
  else colsum_kernel<float><<<grid, 256, 0, ST>>>((const float*)in, M, d, rows_per_cta, scale, out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_layernorm_bwd_cast(const float* x, const void* dy, int dy_dtype, const float* weight, int64_t M, int d,
                                        float eps, int kind, int accumulate, float* dx, float* dweight, float* dbias,
                                        void* cast_out_, float cast_scale, void* stream);

extern "C" int lcasr_layernorm_bwd(const float* x, const void* dy, int dy_dtype, const float* weight, int64_t M, int d,
                                   float eps, int kind, int accumulate, float* dx, float* dweight, float* dbias,
                                   void* stream) {
  return lcasr_layernorm_bwd_cast(x, dy, dy_dtype, weight, M, d, eps, kind, accumulate, dx, dweight, dbias, nullptr, 1.0f, stream);
}

extern "C" int lcasr_layernorm_bwd_cast(const float* x, const void* dy, int dy_dtype, const float* weight, int64_t M, int d,
                                        float eps, int kind, int accumulate, float* dx, float* dweight, float* dbias,
                                        void* cast_out_, float cast_scale, void* stream) {
  bf16* cast_out = (bf16*)cast_out_;
  LCASR_CHECK_ARG(x && dy && weight && dx && dweight && M > 0 && d > 0, "layernorm_bwd: bad arguments");
  LCASR_CHECK_ARG(kind == LCASR_NORM_LAYERNORM || kind == LCASR_NORM_RMSNORM, "layernorm_bwd: bad kind %d", kind);
  LCASR_CHECK_ARG(kind == LCASR_NORM_RMSNORM || dbias, "layernorm_bwd: LayerNorm needs dbias");
  const size_t smem = 2 * (size_t)d * sizeof(float);
  LCASR_CHECK_ARG(smem <= 48 * 1024, "layernorm_bwd: d=%d too wide", d);
  const int64_t want = ceil_div(M, 8);
  const unsigned grid = (unsigned)(want < 2 * kNumSMs ? want : 2 * kNumSMs);
  const bool al16 = (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)weight | (uintptr_t)cast_out) & 15) == 0;
#define LCASR_NB_CASE(NE)                                                                                                  \
  case NE * 32:                                                                                                            \
    if (dy_dtype == LCASR_BF16)                                                                                            \
      norm_bwd_reg_kernel<NE, bf16><<<grid, 256, smem, ST>>>(x, (const bf16*)dy, weight, M, eps, kind, accumulate, dx, dweight, dbias, cast_out, cast_scale); \
    else                                                                                                                   \
      norm_bwd_reg_kernel<NE, float><<<grid, 256, smem, ST>>>(x, (const float*)dy, weight, M, eps, kind, accumulate, dx, dweight, dbias, cast_out, cast_scale); \
    LCASR_LAUNCH_CHECK();                                                                                                  \
    return 0;
  if (al16) {
    switch (d) {
      LCASR_NB_CASE(4) LCASR_NB_CASE(8) LCASR_NB_CASE(16) LCASR_NB_CASE(24) LCASR_NB_CASE(32)  // wider rows: generic kernel
      default: break;
    }
  }
#undef LCASR_NB_CASE
  if (dy_dtype == LCASR_BF16)
    norm_bwd_kernel<bf16><<<grid, 256, smem, ST>>>(x, (const bf16*)dy, weight, M, d, eps, kind, accumulate, dx, dweight, dbias, cast_out, cast_scale);
  else
    norm_bwd_kernel<float><<<grid, 256, smem, ST>>>(x, (const float*)dy, weight, M, d, eps, kind, accumulate, dx, dweight, dbias, cast_out, cast_scale);
  LCASR_LAUNCH_CHECK();
  return 0;
}

template <int MODE>
static int launch_dwconv1d(const void* in, const void* in2, int B, int64_t N, int d, int ks, int tt, const float* w,
                           const float* b, void* out, float* acc0, float* acc1, cudaStream_t st) {
  const int cgroups = d / 8;
  const int threads = cgroups < 128 ? ((cgroups + 31) / 32) * 32 : 128;
  dim3 grid((unsigned)ceil_div(cgroups, threads), (unsigned)ceil_div(N, tt), (unsigned)B), block(threads);
  if (grid.y > 65535) return set_error(LCASR_E_BADARG, "dwconv1d: N=%lld too long for the grid", (long long)N);
#define LCASR_DW1(KS)                                                                                              \
  case KS:                                                                                                         \
    dwconv1d_kernel<KS, MODE><<<grid, block, 0, st>>>((const bf16*)in, (const bf16*)in2, N, d, tt, w, b, (bf16*)out, \
                                                      acc0, acc1);                                                 \
    break;
  switch (ks) {
    LCASR_DW1(3) LCASR_DW1(5) LCASR_DW1(7) LCASR_DW1(9) LCASR_DW1(11) LCASR_DW1(15)
    default:
      return set_error(LCASR_E_UNSUPPORTED, "dwconv1d: conv_kernel_size=%d not in {3,5,7,9,11,15}", ks);
  }
#undef LCASR_DW1
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_dwconv1d_fwd(const void* in, int B, int64_t N, int d, int ksize, const float* w, const float* b,
                                  void* out, double* sum, double* sumsq, void* stream) {
  LCASR_CHECK_ARG(in && out && w && b && B > 0 && N > 0 && d > 0 && d % 8 == 0, "dwconv1d_fwd: bad arguments");
  LCASR_CHECK_ARG((sum == nullptr) == (sumsq == nullptr), "dwconv1d_fwd: sum and sumsq come together");
  static const bool legacy = getenv("LCASR_DWCONV_LEGACY") != nullptr;  // A/B switch: the register-window kernels
  if (!legacy && dwconv1d_tile_ok(d, ksize)) return dwconv1d_tile_fwd(in, B, N, d, ksize, w, b, out, sum, sumsq, ST);
  return launch_dwconv1d<0>(in, nullptr, B, N, d, ksize, 32, w, b, out, reinterpret_cast<float*>(sum), reinterpret_cast<float*>(sumsq), ST);
}

extern "C" int lcasr_dwconv1d_bwd_data(const void* dout, int B, int64_t N, int d, int ksize, const float* w, void* din,
                                       void* stream) {
  LCASR_CHECK_ARG(dout && din && w && B > 0 && N > 0 && d > 0 && d % 8 == 0, "dwconv1d_bwd_data: bad arguments");
  static const bool legacy = getenv("LCASR_DWCONV_LEGACY") != nullptr;
  if (!legacy && dwconv1d_tile_ok(d, ksize)) return dwconv1d_tile_bwd_data(dout, B, N, d, ksize, w, din, ST);
  return launch_dwconv1d<1>(dout, nullptr, B, N, d, ksize, 32, w, nullptr, din, nullptr, nullptr, ST);
}

extern "C" int lcasr_dwconv1d_bwd_weight(const void* x, const void* dout, int B, int64_t N, int d, int ksize, float* dw,
                                         float* db, void* stream) {
  LCASR_CHECK_ARG(x && dout && dw && db && B > 0 && N > 0 && d > 0 && d % 8 == 0, "dwconv1d_bwd_weight: bad arguments");
  static const bool legacy = getenv("LCASR_DWCONV_LEGACY") != nullptr;
  if (!legacy && dwconv1d_tile_ok(d, ksize)) return dwconv1d_tile_bwd_weight(x, dout, B, N, d, ksize, dw, db, ST);
  return launch_dwconv1d<2>(x, dout, B, N, d, ksize, 128, nullptr, nullptr, nullptr, dw, db, ST);
}

extern "C" int lcasr_brn_train_stats(const double* sum, const double* sumsq, int64_t count, int d, float* running_mean,
                                     float* running_std, float eps, float rmax, float dmax, float momentum,
                                     const float* weight, const float* bias, float* A, float* Bc, float* stats,
                                     void* stream) {
  LCASR_CHECK_ARG(sum && sumsq && running_mean && running_std && weight && bias && A && Bc && stats && count > 0 && d > 0,
                  "brn_train_stats: bad arguments");
  brn_train_stats_kernel<<<(unsigned)ceil_div(d, 128), 128, 0, ST>>>(sum, sumsq, (float)count, d, running_mean, running_std,
                                                                    eps, rmax, dmax, momentum, weight, bias, A, Bc, stats);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_affine_silu(const void* c, int64_t M, int d, const float* A, const float* Bc, void* out, void* stream) {
  LCASR_CHECK_ARG(c && A && Bc && out && M > 0 && d > 0 && d % 8 == 0, "affine_silu: bad arguments");
  affine_silu_kernel<<<grid_1d(M * (d / 8), 256), 256, 0, ST>>>((const bf16*)c, M, d, A, Bc, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_affine_silu_bwd(const void* c, const void* dy, int64_t M, int d, const float* A, const float* Bc,
                                     const float* stats, void* dz, double* S1, double* S2, void* stream) {
  LCASR_CHECK_ARG(c && dy && A && Bc && stats && dz && S1 && S2 && M > 0 && d > 0 && d % 8 == 0, "affine_silu_bwd: bad arguments");
  const int rows_per_cta = 128;
  dim3 grid((unsigned)ceil_div(d / 8, 32), (unsigned)ceil_div(M, rows_per_cta));
  LCASR_CHECK_ARG(grid.y <= 65535, "affine_silu_bwd: too many rows");
  affine_silu_bwd_kernel<<<grid, 256, 0, ST>>>((const bf16*)c, (const bf16*)dy, M, d, rows_per_cta, A, Bc, stats, (bf16*)dz,
                                               S1, S2);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_brn_bwd_finalize(const double* S1, const double* S2, int64_t count, int d, const float* weight,
                                      const float* stats, float* dweight, float* dbias, float* coef, void* stream) {
  LCASR_CHECK_ARG(S1 && S2 && weight && stats && dweight && dbias && coef && count > 0 && d > 0, "brn_bwd_finalize: bad arguments");
  brn_bwd_finalize_kernel<<<(unsigned)ceil_div(d, 128), 128, 0, ST>>>(S1, S2, (float)count, d, weight, stats, dweight, dbias, coef);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_affine3(const void* a, const void* b, int64_t M, int d, const float* coef, void* out, void* stream) {
  LCASR_CHECK_ARG(a && b && coef && out && M > 0 && d > 0 && d % 8 == 0, "affine3: bad arguments");
  affine3_kernel<<<grid_1d(M * (d / 8), 256), 256, 0, ST>>>((const bf16*)a, (const bf16*)b, M, d, coef, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_add_bf16(float* x, const void* a, int64_t n, void* stream) {
  LCASR_CHECK_ARG(x && a && n >= 0 && n % 8 == 0, "add_bf16: n must be a multiple of 8");
  if (n == 0) return 0;
  add_bf16_kernel<<<grid_1d(n / 8, 256), 256, 0, ST>>>(x, (const bf16*)a, n / 8);
  LCASR_LAUNCH_CHECK();
  return 0;
}
