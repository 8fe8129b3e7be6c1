// Depthwise Conv1d(k, pad (k-1)/2, groups=d) of the Conformer convolution module (convolution.py:112), shared-memory
// tiled — inference (fused with the BatchRenorm-eval affine and SiLU) and the three training forms (forward with the
// BatchRenorm batch statistics, data gradient, weight/bias gradient).  bf16 channels-last [B,N,d], d % 128 == 0.
//
// The first versions (convmod.cu, train.cu: one thread = 8 channels x 32..128 consecutive tokens, register sliding window)
// issued ONE dependent 16-byte load per token and thread and ran 3 warps per CTA: 45-96 us for a 25-50 MB problem
// (0.5-1.1 TB/s).  Here a CTA stages a [64 + k - 1] x 128-channel tile (32 tokens + their output gradients for the weight gradient) with cp.async (every load of the tile in flight at
// once, zero-filled outside the sequence), then a warp produces TPT consecutive tokens x 128 channels (4 per lane) from shared memory,
// walking the k + TPT - 1 window rows once.  Reductions (statistics, weight gradients) stay in registers across the tiles
// a CTA walks and leave through one shared-memory pass + one atomic per value and CTA.
#include "common.cuh"
#include <algorithm>

namespace lcasr {

constexpr int kDtCS = 128;   // channels per CTA: one warp covers a 256-byte row, a lane owns 4 channels (8 bytes)
constexpr int kDtTL = 8;     // time lanes per CTA = warps (256 threads)

enum { DT_FWD = 0, DT_BWD_DATA = 1, DT_BWD_WEIGHT = 2, DT_EVAL = 3 };

__device__ __forceinline__ void dt_cp16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

__device__ __forceinline__ void dt_unpack(const uint2& r, float (&v)[4]) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

struct DtParams {
  const bf16* in;        // x (FWD, EVAL, BWD_WEIGHT) or dout (BWD_DATA)
  const bf16* in2;       // dout (BWD_WEIGHT)
  bf16* out;             // FWD / BWD_DATA / EVAL
  const float* w;        // [d, KS]
  const float* b;        // conv bias (FWD, EVAL)
  const float *rm, *rs, *bw, *bb;  // EVAL: BatchRenorm running mean / std, weight, bias
  float* acc0;           // FWD: sum [d] (fp64 behind the pointer; may be NULL); BWD_WEIGHT: dw [d, KS]
  float* acc1;           // FWD: sum of squares [d];    BWD_WEIGHT: db [d]
  int64_t N;
  int d;
};

template <int KS, int MODE, int TPT>
__global__ void __launch_bounds__(256, 2) dwconv1d_tile_kernel(const DtParams p) {
  constexpr int PAD = (KS - 1) / 2;
  constexpr int TT = kDtTL * TPT;          // tokens per tile
  constexpr int XR = TT + KS - 1;          // staged input rows
  constexpr int NRED = MODE == DT_BWD_WEIGHT ? 4 * (KS + 1) : 8;   // values a thread contributes to the final reduction
  extern __shared__ __align__(16) uint8_t dt_sm[];
  constexpr int BUF = (XR + (MODE == DT_BWD_WEIGHT ? TT : 0)) * 32;   // uint2 elements of one stage: x rows, then dout rows
  uint2* s_buf = reinterpret_cast<uint2*>(dt_sm);                     // two stages (the next tile arrives while this one is used)
  const int lane = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * kDtCS + lane * 4;
  const int64_t batch = blockIdx.z;
  const int64_t N = p.N;
  const int d = p.d;
  const bf16* xin = p.in + batch * N * d + blockIdx.x * kDtCS;

  float wr[4][KS];    // taps (FWD/EVAL), flipped taps (BWD_DATA) or weight-gradient accumulators (BWD_WEIGHT)
  float e0[4], e1[4]; // EVAL: out = silu(acc * e0 + e1);  FWD: out = acc + e1
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      if (MODE == DT_BWD_WEIGHT) wr[c][k] = 0.f;
      else if (MODE == DT_BWD_DATA) wr[c][k] = p.w[(c0 + c) * KS + (KS - 1 - k)];
      else wr[c][k] = p.w[(c0 + c) * KS + k];
    }
    e0[c] = 0.f; e1[c] = 0.f;
    if (MODE == DT_FWD) e1[c] = p.b[c0 + c];
    if (MODE == DT_EVAL) {  // ((acc + b) - mean) / std * bw + bb  ==  acc*scale + shift   (batchrenorm.py:86-91)
      e0[c] = p.bw[c0 + c] / p.rs[c0 + c];
      e1[c] = (p.b[c0 + c] - p.rm[c0 + c]) * e0[c] + p.bb[c0 + c];
    }
  }
  float s1[4], s2[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
  const bool stats = MODE == DT_FWD && p.acc0 != nullptr;

  const int64_t ntiles = (N + TT - 1) / TT;
  auto stage = [&](int64_t tile_, int buf) {
    if (tile_ < ntiles) {
      const int64_t n0_ = tile_ * TT;
      uint4* dx = reinterpret_cast<uint4*>(s_buf + buf * BUF);
      for (int i = threadIdx.x; i < XR * 16; i += 256) {   // 16-byte chunks: 16 per row
        const int r = i >> 4, ch = i & 15;
        const int64_t n = n0_ - PAD + r;
        const bool ok = n >= 0 && n < N;
        dt_cp16(dx + i, ok ? xin + n * d + ch * 8 : xin, ok);
      }
      if (MODE == DT_BWD_WEIGHT) {
        const bf16* gin = p.in2 + batch * N * d + blockIdx.x * kDtCS;
        uint4* dg = dx + XR * 16;
        for (int i = threadIdx.x; i < TT * 16; i += 256) {
          const int r = i >> 4, ch = i & 15;
          const int64_t n = n0_ + r;
          const bool ok = n < N;
          dt_cp16(dg + i, ok ? gin + n * d + ch * 8 : gin, ok);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(blockIdx.y, 0);
  int it = 0;
  for (int64_t tile = blockIdx.y; tile < ntiles; tile += gridDim.y, ++it) {
    const int64_t n0 = tile * TT;
    stage(tile + gridDim.y, (it + 1) & 1);   // that stage's readers passed the barrier at the end of the previous iteration
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const uint2* s_x = s_buf + (it & 1) * BUF;
    const uint2* s_g = s_x + XR * 32;
    const int t0 = tl * TPT;  // first token of this warp inside the tile; window rows t0 .. t0 + TPT + KS - 2
    if (MODE == DT_BWD_WEIGHT) {
      float g[TPT][4];
#pragma unroll
      for (int u = 0; u < TPT; ++u) {
        dt_unpack(s_g[(t0 + u) * 32 + lane], g[u]);   // rows beyond N are zero-filled
#pragma unroll
        for (int c = 0; c < 4; ++c) s1[c] += g[u][c];
      }
#pragma unroll
      for (int r = 0; r < TPT + KS - 1; ++r) {
        float x[4];
        dt_unpack(s_x[(t0 + r) * 32 + lane], x);
#pragma unroll
        for (int u = 0; u < TPT; ++u) {
          const int k = r - u;                       // x row r is tap k of token u
          if (k >= 0 && k < KS) {
#pragma unroll
            for (int c = 0; c < 4; ++c) wr[c][k] = fmaf(g[u][c], x[c], wr[c][k]);
          }
        }
      }
    } else {
      float y[TPT][4];
#pragma unroll
      for (int u = 0; u < TPT; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) y[u][c] = 0.f;
#pragma unroll
      for (int r = 0; r < TPT + KS - 1; ++r) {
        float x[4];
        dt_unpack(s_x[(t0 + r) * 32 + lane], x);
#pragma unroll
        for (int u = 0; u < TPT; ++u) {
          const int k = r - u;
          if (k >= 0 && k < KS) {
#pragma unroll
            for (int c = 0; c < 4; ++c) y[u][c] = fmaf(wr[c][k], x[c], y[u][c]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < TPT; ++u) {
        const int64_t n = n0 + t0 + u;
        if (n < N) {
          float o[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (MODE == DT_EVAL) o[c] = silu_fast(fmaf(y[u][c], e0[c], e1[c]));
            else if (MODE == DT_FWD) o[c] = y[u][c] + e1[c];
            else o[c] = y[u][c];
          }
          // packed conversions (F2FP, not the quarter-rate scalar F2F); the statistics are those of the ROUNDED
          // activation — what the next kernels read
          const __nv_bfloat162 q0 = __floats2bfloat162_rn(o[0], o[1]), q1 = __floats2bfloat162_rn(o[2], o[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&q0);
          pk.y = *reinterpret_cast<const uint32_t*>(&q1);
          *reinterpret_cast<uint2*>(p.out + (batch * N + n) * d + c0) = pk;
          if (stats) {
            float yr[4];
            dt_unpack(pk, yr);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              s1[c] += yr[c];
              s2[c] = fmaf(yr[c], yr[c], s2[c]);
            }
          }
        }
      }
    }
    __syncthreads();  // this stage may be overwritten by the load issued in the next iteration
  }

  if (MODE == DT_BWD_WEIGHT || stats) {
    // the 8 warps (time lanes) hold partial sums for the same 128 channels: one pass through shared memory
    float* s_red = reinterpret_cast<float*>(dt_sm);   // [8 warps][NRED][32 lanes]
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // the last tile is no longer read, no copy is in flight
    auto put = [&](int idx, float v) { s_red[(tl * NRED + idx) * 32 + lane] = v; };
    if (MODE == DT_BWD_WEIGHT) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < KS; ++k) put(c * (KS + 1) + k, wr[c][k]);
        put(c * (KS + 1) + KS, s1[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) { put(c, s1[c]); put(4 + c, s2[c]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NRED * 32; i += 256) {
      const int idx = i >> 5, l32 = i & 31;
      float v = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) v += s_red[(w8 * NRED + idx) * 32 + l32];
      const int cbase = blockIdx.x * kDtCS + l32 * 4;
      if (MODE == DT_BWD_WEIGHT) {
        const int c = idx / (KS + 1), k = idx - c * (KS + 1);
        if (k < KS) atomicAdd(p.acc0 + (int64_t)(cbase + c) * KS + k, v);
        else atomicAdd(p.acc1 + cbase + c, v);
      } else {  // BatchRenorm batch statistics: fp64 across CTAs, so the fp32 mean / variance do not depend on the order
        if (idx < 4) atomicAdd(reinterpret_cast<double*>(p.acc0) + cbase + idx, (double)v);
        else atomicAdd(reinterpret_cast<double*>(p.acc1) + cbase + idx - 4, (double)v);
      }
    }
  }
}

template <int KS, int MODE, int TPT>
static int dt_launch_ks(const DtParams& p, int B, cudaStream_t st) {
  constexpr int TT = kDtTL * TPT;
  constexpr size_t tile_bytes = (size_t)2 * (TT + KS - 1 + (MODE == DT_BWD_WEIGHT ? TT : 0)) * 256;
  constexpr size_t red_bytes = MODE == DT_BWD_WEIGHT ? (size_t)8 * 4 * (KS + 1) * 32 * 4 : (MODE == DT_FWD ? 8 * 8 * 32 * 4 : 0);
  constexpr size_t smem = tile_bytes > red_bytes ? tile_bytes : red_bytes;
  static_assert(smem <= 100 * 1024, "dwconv1d tile: shared memory");
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev) && smem > 48 * 1024) {
    LCASR_CUDA(cudaFuncSetAttribute(dwconv1d_tile_kernel<KS, MODE, TPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.mark(attr_dev);
  }
  const int64_t ntiles = ceil_div(p.N, TT);
  const int slabs = p.d / kDtCS;
  // one wave of 2 CTAs per SM, each walking several tiles (double-buffered; partial sums of the reductions in registers)
  int64_t gy = std::max<int64_t>(1, ((int64_t)kNumSMs * 2) / ((int64_t)B * slabs));
  if (gy > ntiles) gy = ntiles;
  LCASR_CHECK_ARG(gy <= 65535 && B <= 65535, "dwconv1d: sequence / batch too long for the grid");
  dim3 grid((unsigned)slabs, (unsigned)gy, (unsigned)B);
  dwconv1d_tile_kernel<KS, MODE, TPT><<<grid, 256, smem, st>>>(p);
  LCASR_LAUNCH_CHECK();
  return 0;
}

// returns LCASR_E_UNSUPPORTED (without setting an error message the caller would report) when the shape is not covered
template <int MODE>
int dwconv1d_tile_launch(const DtParams& p, int B, int ks, cudaStream_t st) {
  constexpr int TPT = MODE == DT_BWD_WEIGHT ? 4 : 8;
  switch (ks) {
    case 3: return dt_launch_ks<3, MODE, TPT>(p, B, st);
    case 5: return dt_launch_ks<5, MODE, TPT>(p, B, st);
    case 7: return dt_launch_ks<7, MODE, TPT>(p, B, st);
    case 9: return dt_launch_ks<9, MODE, TPT>(p, B, st);
    case 11: return dt_launch_ks<11, MODE, TPT>(p, B, st);
    case 15: return dt_launch_ks<15, MODE, TPT>(p, B, st);
    default: return LCASR_E_UNSUPPORTED;
  }
}

bool dwconv1d_tile_ok(int d, int ks) {
  return d % kDtCS == 0 && (ks == 3 || ks == 5 || ks == 7 || ks == 9 || ks == 11 || ks == 15);
}

int dwconv1d_tile_fwd(const void* in, int B, int64_t N, int d, int ks, const float* w, const float* b, void* out, double* sum,
                      double* sumsq, cudaStream_t st) {
  DtParams p{(const bf16*)in, nullptr, (bf16*)out, w, b, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<float*>(sum),
             reinterpret_cast<float*>(sumsq), N, d};
  return dwconv1d_tile_launch<DT_FWD>(p, B, ks, st);
}
int dwconv1d_tile_bwd_data(const void* dout, int B, int64_t N, int d, int ks, const float* w, void* din, cudaStream_t st) {
  DtParams p{(const bf16*)dout, nullptr, (bf16*)din, w, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, N, d};
  return dwconv1d_tile_launch<DT_BWD_DATA>(p, B, ks, st);
}
int dwconv1d_tile_bwd_weight(const void* x, const void* dout, int B, int64_t N, int d, int ks, float* dw, float* db,
                             cudaStream_t st) {
  DtParams p{(const bf16*)x, (const bf16*)dout, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dw, db, N, d};
  return dwconv1d_tile_launch<DT_BWD_WEIGHT>(p, B, ks, st);
}
int dwconv1d_tile_eval(const void* in, int B, int64_t N, int d, int ks, const float* w, const float* b, const float* rm,
                       const float* rs, const float* bw, const float* bb, void* out, cudaStream_t st) {
  DtParams p{(const bf16*)in, nullptr, (bf16*)out, w, b, rm, rs, bw, bb, nullptr, nullptr, N, d};
  return dwconv1d_tile_launch<DT_EVAL>(p, B, ks, st);
}

}  // namespace lcasr
