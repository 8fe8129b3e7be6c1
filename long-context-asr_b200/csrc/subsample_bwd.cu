// Backward of the 8x depthwise-striding subsampling stencils (lcasr/components/subsampling.py:277-323), bf16
// channels-last activations, fp32 parameter gradients.  HBM-bound like their forward counterparts in subsample.cu:
//   dwconv data gradient : reads the [B,Tout,Fout,C] output gradient (each element up to 9/4 times, L2-resident
//                          neighbours), writes [B,Tin,Fin,C]
//   dwconv weight grad   : reads input + output gradient once, 10 fp32 atomics per channel and CTA
//   conv0 weight grad    : reads the spectrogram (x C/64 from L2) and the conv0 output gradient once; the conv0
//                          pre-activation is recomputed from the spectrogram instead of being stored
#include "common.cuh"
#include <cstdlib>
#include <algorithm>

namespace lcasr {

__device__ __forceinline__ float silu_grad_sb(float x) {
  const float s = sigmoid_fast(x);
  return s * (1.0f + x * (1.0f - s));
}

// din[b,ti,fi,:] = sum_{i,j : ti = 2*to-1+i, fi = 2*fo-1+j} dout[b,to,fo,:] * w[:, i*3+j]
// Stride 2 makes the tap pattern depend only on the parity of (ti, fi), so a thread produces one 2x2 block of
// input positions (rows 2a, 2a+1; columns 2e, 2e+1) from the four output gradients G[a..a+1][e..e+1]:
//   din[2a  ][2e  ] = G00 w11                 din[2a  ][2e+1] = G00 w12 + G01 w10
//   din[2a+1][2e  ] = G00 w21 + G10 w01       din[2a+1][2e+1] = G00 w22 + G01 w20 + G10 w02 + G11 w00
// 4 unconditional 16-byte loads and 4 stores per thread (instead of 9 predicated gathers per output), one CTA per
// kBdTB row pairs of one recording, 32-bit index arithmetic.
constexpr int kBdTB = 2;
template <int V>  // channels per thread (8: 16-byte accesses; 4: half the registers, twice the resident warps)
__global__ void __launch_bounds__(256) subsample_dwconv_bwd_data_kernel(const bf16* __restrict__ dout, const float* __restrict__ w,
                                                                        int64_t Tin, int Fin, int C, int64_t Tout, int Fout,
                                                                        bf16* __restrict__ din) {
  const int cgroups = C / V;
  const int cg = threadIdx.x % cgroups;
  float wr[V][9];
#pragma unroll
  for (int c = 0; c < V; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = w[(cg * V + c) * 9 + k];
  const int64_t b = blockIdx.y;
  const int64_t a0 = (int64_t)blockIdx.x * kBdTB;
  const int Fh = (Fin + 1) / 2;            // column pairs
  const int per_row = Fh * cgroups;
  const bf16* doutb = dout + b * Tout * Fout * C;
  bf16* dinb = din + b * Tin * Fin * C;
  for (int idx = threadIdx.x; idx < kBdTB * per_row; idx += blockDim.x) {
    const int al = idx / per_row;
    const int e = (idx - al * per_row) / cgroups;
    const int64_t a = a0 + al;
    if (2 * a >= Tin) break;
    float g00[V], g01[V], g10[V], g11[V];
    const bool r1 = a + 1 < Tout, c1 = e + 1 < Fout, r0 = a < Tout, c0 = e < Fout;
    auto ld = [&](float (&g)[V], bool ok, int64_t to, int fo) {
      if (ok) VecB<V>::load(doutb + ((to * Fout + fo) * C + cg * V), g);
      else {
#pragma unroll
        for (int c = 0; c < V; ++c) g[c] = 0.f;
      }
    };
    ld(g00, r0 && c0, a, e); ld(g01, r0 && c1, a, e + 1); ld(g10, r1 && c0, a + 1, e); ld(g11, r1 && c1, a + 1, e + 1);
    float o[V];
    const int64_t ti = 2 * a;
    const int fi = 2 * e;
#pragma unroll
    for (int c = 0; c < V; ++c) o[c] = g00[c] * wr[c][4];
    VecB<V>::store(dinb + ((ti * Fin + fi) * C + cg * V), o);
    if (fi + 1 < Fin) {
#pragma unroll
      for (int c = 0; c < V; ++c) o[c] = fmaf(g00[c], wr[c][5], g01[c] * wr[c][3]);
      VecB<V>::store(dinb + ((ti * Fin + fi + 1) * C + cg * V), o);
    }
    if (ti + 1 < Tin) {
#pragma unroll
      for (int c = 0; c < V; ++c) o[c] = fmaf(g00[c], wr[c][7], g10[c] * wr[c][1]);
      VecB<V>::store(dinb + (((ti + 1) * Fin + fi) * C + cg * V), o);
      if (fi + 1 < Fin) {
#pragma unroll
        for (int c = 0; c < V; ++c)
          o[c] = fmaf(g00[c], wr[c][8], fmaf(g01[c], wr[c][6], fmaf(g10[c], wr[c][2], g11[c] * wr[c][0])));
        VecB<V>::store(dinb + (((ti + 1) * Fin + fi + 1) * C + cg * V), o);
      }
    }
  }
}

// dw[c, i*3+j] += sum dout[b,to,fo,c] * in[b,2to-1+i,2fo-1+j,c] ; db[c] += sum dout
// CTAs stride over blocks of kBwTB output rows (grid.x of them per recording); register accumulators, one
// shared-memory reduction and C*10 global atomics per CTA.
constexpr int kBwTB = 4;
template <int V>
__global__ void __launch_bounds__(256) subsample_dwconv_bwd_weight_kernel(const bf16* __restrict__ in, const bf16* __restrict__ dout,
                                                                          int64_t Tin, int Fin, int C, int64_t Tout, int Fout,
                                                                          float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float sacc[];  // [C][10]
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cgroups = C / V;
  const int cg = threadIdx.x % cgroups;
  float aw[V][9], ab[V];
#pragma unroll
  for (int c = 0; c < V; ++c) {
    ab[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) aw[c][k] = 0.f;
  }
  const int64_t b = blockIdx.y;
  const int per_row = Fout * cgroups;
  const bf16* inb = in + b * Tin * Fin * C;
  const bf16* doutb = dout + b * Tout * Fout * C;
  const int64_t nblk = (Tout + kBwTB - 1) / kBwTB;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t t0 = blk * kBwTB;
    for (int idx = threadIdx.x; idx < kBwTB * per_row; idx += blockDim.x) {
      const int tl = idx / per_row;
      const int fo = (idx - tl * per_row) / cgroups;
      const int64_t to = t0 + tl;
      if (to >= Tout) break;
      float g[V];
      VecB<V>::load(doutb + ((to * Fout + fo) * C + cg * V), g);
#pragma unroll
      for (int c = 0; c < V; ++c) ab[c] += g[c];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int64_t ti = 2 * to - 1 + i;
        if (ti < 0 || ti >= Tin) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int fi = 2 * fo - 1 + j;
          if (fi < 0 || fi >= Fin) continue;
          float v[V];
          VecB<V>::load(inb + ((ti * Fin + fi) * C + cg * V), v);
#pragma unroll
          for (int c = 0; c < V; ++c) aw[c][i * 3 + j] = fmaf(g[c], v[c], aw[c][i * 3 + j]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < V; ++c) {
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(&sacc[(cg * V + c) * 10 + k], aw[c][k]);
    atomicAdd(&sacc[(cg * V + c) * 10 + 9], ab[c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) {
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, sacc[i]);
    else atomicAdd(db + c, sacc[i]);
  }
}

// conv0 (1->C, 3x3, s2, p1) + SiLU: gradients of w0 [C,9] and b0 [C] from ds1 = dL/d(silu output) [B,T1,F1,C].
// Same tiling as subsample_conv0_kernel (16 output frames per tile, the input patch in shared memory, each thread
// owns 8 channels); a CTA walks several tiles and keeps its partial sums in registers, so the reduction costs one
// shared-memory pass and C*10 global atomics per CTA, not per tile.
constexpr int kC0bTT = 16;
__global__ void __launch_bounds__(256) subsample_conv0_bwd_kernel(const float* __restrict__ spec, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, const bf16* __restrict__ ds1,
                                                                  int F, int64_t T, int C, int64_t T1, int F1,
                                                                  float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float smem[];  // input patch [(2*TT+1)][F+2], then accumulators [C][10]
  const int FW = F + 2;
  const int rows = 2 * kC0bTT + 1;
  float* s_in = smem;
  float* sacc = smem + rows * FW;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) sacc[i] = 0.f;
  const int cgroups = C / 8;
  const int pos_stride = blockDim.x / cgroups;
  const bool worker = (int)threadIdx.x < pos_stride * cgroups;
  const int cg = threadIdx.x % cgroups;
  const int pos_lane = threadIdx.x / cgroups;
  float wr[8][9], br[8], aw[8][9], ab[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = bias[cg * 8 + c];
    ab[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { wr[c][k] = w[(cg * 8 + c) * 9 + k]; aw[c][k] = 0.f; }
  }
  const int64_t ntiles = (T1 + kC0bTT - 1) / kC0bTT;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t1_0 = tile * kC0bTT;
    const int64_t t_in0 = 2 * t1_0 - 1;
    __syncthreads();  // the previous tile's patch is no longer read
    for (int idx = threadIdx.x; idx < rows * FW; idx += blockDim.x) {
      const int f = idx / rows - 1;
      const int r = idx % rows;
      const int64_t t = t_in0 + r;
      float v = 0.f;
      if (f >= 0 && f < F && t >= 0 && t < T) v = spec[((int64_t)b * F + f) * T + t];
      s_in[r * FW + (f + 1)] = v;
    }
    __syncthreads();
    if (worker) {
      const int npos = kC0bTT * F1;
      for (int p = pos_lane; p < npos; p += pos_stride) {
        const int tt = p / F1, f1 = p % F1;
        const int64_t t1 = t1_0 + tt;
        if (t1 >= T1) break;
        float in[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) in[i * 3 + j] = s_in[(2 * tt + i) * FW + (2 * f1 + j)];
        float g[8];
        Vec8<bf16>::load(ds1 + ((((int64_t)b * T1 + t1) * F1 + f1) * C + cg * 8), g);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float a = br[c];
#pragma unroll
          for (int k = 0; k < 9; ++k) a = fmaf(wr[c][k], in[k], a);
          const float ga = g[c] * silu_grad_sb(a);
          ab[c] += ga;
#pragma unroll
          for (int k = 0; k < 9; ++k) aw[c][k] = fmaf(ga, in[k], aw[c][k]);
        }
      }
    }
  }
  if (worker) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(&sacc[(cg * 8 + c) * 10 + k], aw[c][k]);
      atomicAdd(&sacc[(cg * 8 + c) * 10 + 9], ab[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) {
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, sacc[i]);
    else atomicAdd(db + c, sacc[i]);
  }
}

// ---- fused backward of conv0 + SiLU + first depthwise level -------------------------------------------------------
// The unfused chain writes the 160x-expanded conv0 activation s1 (1.3 GB in bf16 at cfg 5) in the forward, reads it twice
// (depthwise forward, depthwise weight gradient), writes its gradient ds1 and reads that back for the conv0 weight gradient:
// 6.7 GB of HBM traffic for 2.7 % of the step's FLOPs.  Nothing of it is needed: s1 = silu(conv0(spec)) is 9 FMAs and one
// tanh away from the 80-bin spectrogram, and ds1 is consumed where it is produced.  This kernel reads only the
// spectrogram (L2-resident patch in shared memory) and dd1 = dL/d(depthwise output) [B,T2,F2,C] and accumulates all four
// parameter gradients (conv0 w/b, depthwise w/b) in registers:
//   a warp owns one 2x2 block of s1 positions (rows 2a,2a+1; columns 2e,2e+1), a lane two channels.  Stride 2 makes the
//   tap pattern depend only on the parity of the position, so the block needs the four output gradients
//   G[a..a+1][e..e+1] and the 5x5 input patch around it — no divergence, no gathers:
//     ds1[2a  ][2e  ] = G00 w11                 ds1[2a  ][2e+1] = G00 w12 + G01 w10
//     ds1[2a+1][2e  ] = G00 w21 + G10 w01       ds1[2a+1][2e+1] = G00 w22 + G01 w20 + G10 w02 + G11 w00
//   and every (G, tap) pair above also feeds the depthwise weight gradient dw1[tap] += G * s1(position).
// A CTA (one recording, 64 channels) walks tiles of kL1R s1 rows; one shared-memory reduction and 64*20 atomics per CTA.
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc, bool valid) {  // zero-fills when !valid
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kL1R = 16;   // s1 rows per tile (even)
constexpr int kL1CG = 64;  // channels per CTA (2 per lane)

__global__ void __launch_bounds__(256, 2) subsample_l1_bwd_fused_kernel(
    const float* __restrict__ spec, const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
    const bf16* __restrict__ dd1, int F, int64_t T, int C, int64_t T1, int F1, int64_t T2, int F2, int FWp,
    float* __restrict__ dw0, float* __restrict__ db0, float* __restrict__ dw1, float* __restrict__ db1) {
  // two input patches [2*kL1R+1][FWp] (col 0 = freq -1, zero-filled up to FWp; double-buffered: the next tile's patch
  // arrives by cp.async while this tile is processed — the kernel used to idle on the staging loads), then accumulators [64][20]
  extern __shared__ float l1sm[];
  constexpr int INR = 2 * kL1R + 1;
  float* sacc = l1sm + 2 * INR * FWp;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kL1CG + 2 * lane;
  for (int i = threadIdx.x; i < kL1CG * 20; i += blockDim.x) sacc[i] = 0.f;
  float wa[2][9], ba[2], wd[2][9];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    ba[c] = b0[c0 + c];
#pragma unroll
    for (int k = 0; k < 9; ++k) { wa[c][k] = w0[(c0 + c) * 9 + k]; wd[c][k] = w1[(c0 + c) * 9 + k]; }
  }
  float g0w[2][9], g0b[2], g1w[2][9], g1b[2];  // gradient partial sums: conv0 w/b, depthwise w/b
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    g0b[c] = 0.f; g1b[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { g0w[c][k] = 0.f; g1w[c][k] = 0.f; }
  }
  const int EB = (F1 + 1) / 2;                 // column blocks
  const bf16* ddb = dd1 + (int64_t)b * T2 * F2 * C + c0;
  const int64_t ntiles = (T1 + kL1R - 1) / kL1R;
  auto stage = [&](int64_t tile_, float* dst) {
    if (tile_ < ntiles) {
      const int64_t t0_ = 2 * tile_ * kL1R - 1;  // input frame of patch row 0
      for (int idx = threadIdx.x; idx < INR * FWp; idx += blockDim.x) {
        const int f = idx / INR - 1, r = idx - (f + 1) * INR;
        const int64_t t = t0_ + r;
        const bool ok = f >= 0 && f < F && t >= 0 && t < T;
        cp_async_f32(dst + r * FWp + (f + 1), ok ? spec + ((int64_t)b * F + f) * T + t : spec, ok);
      }
    }
    cp_async_commit();
  };
  stage(blockIdx.x, l1sm);
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int64_t r0 = tile * kL1R;            // first s1 row of the tile (even)
    const float* s_in = l1sm + (it & 1) * INR * FWp;
    stage(tile + gridDim.x, l1sm + ((it + 1) & 1) * INR * FWp);   // that buffer's readers finished at the barrier below
    cp_async_wait<1>();
    __syncthreads();                           // this tile's patch is complete and visible
    const int a_base = (int)(r0 / 2);          // first block row == first dd1 row touched (T2 < 2^31 rows)
    const int nal = (int)min((int64_t)(kL1R / 2), (T1 - r0 + 1) / 2);   // block rows of this tile that exist
    const int T2i = (int)T2, T1rem = (int)min((int64_t)kL1R, T1 - r0); // s1 rows of this tile that exist
    int al = 0, e = wid;
    while (e >= EB) { e -= EB; ++al; }
    // raw bf16x2 output gradients of a block, always from a clamped (valid) address: loads never branch, the
    // out-of-range ones are zeroed by a select at the point of use.  The next block's four values are requested
    // before the current block's ~200 FMAs (the kernel was stalled on exactly these loads: long-scoreboard 2.7).
    auto ldg4 = [&](int al_, int e_, uint32_t (&g)[4]) {
      const int a_ = a_base + al_;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int ai = min(a_ + i, T2i - 1), ej = min(e_ + j, F2 - 1);
          g[i * 2 + j] = __ldg(reinterpret_cast<const uint32_t*>(ddb + ((int64_t)ai * F2 + ej) * C));
        }
    };
    uint32_t gcur[4], gnext[4];
    ldg4(al, e, gcur);
#pragma unroll 1
    for (; al < nal;) {
      const int a = a_base + al;
      int al2 = al, e2 = e + 8;
      while (e2 >= EB) { e2 -= EB; ++al2; }
      ldg4(al2, e2, gnext);
      float2 G[2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bool ok = (a + i < T2i) && (e + j < F2);
          const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gcur[i * 2 + j]));
          G[i][j] = make_float2(ok ? g.x : 0.f, ok ? g.y : 0.f);
        }
      // bias gradient of the depthwise level: every output gradient is G00 of exactly one block
      g1b[0] += G[0][0].x; g1b[1] += G[0][0].y;
      // 5x5 input patch: rows 4*al .. 4*al+4 of the tile patch, columns 4e .. 4e+4 (col 0 = freq -1; FWp >= 4*EB+1)
      float in[5][5];
      const float* ip = s_in + (4 * al) * FWp + 4 * e;
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) in[i][j] = ip[i * FWp + j];
#pragma unroll
      for (int pi = 0; pi < 2; ++pi) {
#pragma unroll
        for (int pj = 0; pj < 2; ++pj) {
          const bool live = (2 * al + pi < T1rem) && (2 * e + pj < F1);
          float acc[2] = {ba[0], ba[1]};
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) acc[c] = fmaf(wa[c][i * 3 + j], in[2 * pi + i][2 * pj + j], acc[c]);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float sg = sigmoid_fast(acc[c]);
            // the forward stored s1 as bf16: the depthwise weight gradient sees the same rounded activation
            const float y = live ? __bfloat162float(__float2bfloat16_rn(acc[c] * sg)) : 0.f;
            const float dsilu = sg * (1.0f + acc[c] * (1.0f - sg));
            float ds = 0.f;
            // (output offset i2, j2; tap) pairs of this parity class — see the table in the header comment
#define LCASR_L1_PAIR(i2, j2, tap)                                          \
  {                                                                         \
    const float g = c == 0 ? G[i2][j2].x : G[i2][j2].y;                     \
    ds = fmaf(g, wd[c][tap], ds);                                           \
    g1w[c][tap] = fmaf(g, y, g1w[c][tap]);                                  \
  }
            if (pi == 0 && pj == 0) { LCASR_L1_PAIR(0, 0, 4) }
            if (pi == 0 && pj == 1) { LCASR_L1_PAIR(0, 0, 5) LCASR_L1_PAIR(0, 1, 3) }
            if (pi == 1 && pj == 0) { LCASR_L1_PAIR(0, 0, 7) LCASR_L1_PAIR(1, 0, 1) }
            if (pi == 1 && pj == 1) { LCASR_L1_PAIR(0, 0, 8) LCASR_L1_PAIR(0, 1, 6) LCASR_L1_PAIR(1, 0, 2) LCASR_L1_PAIR(1, 1, 0) }
#undef LCASR_L1_PAIR
            const float ga = live ? ds * dsilu : 0.f;
            g0b[c] += ga;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) g0w[c][i * 3 + j] = fmaf(ga, in[2 * pi + i][2 * pj + j], g0w[c][i * 3 + j]);
          }
        }
      }
      al = al2; e = e2;
#pragma unroll
      for (int i = 0; i < 4; ++i) gcur[i] = gnext[i];
    }
    __syncthreads();                           // the patch is no longer read: the stage after next may overwrite it
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float* sa = sacc + (2 * lane + c) * 20;
#pragma unroll
    for (int k = 0; k < 9; ++k) { atomicAdd(sa + k, g0w[c][k]); atomicAdd(sa + 10 + k, g1w[c][k]); }
    atomicAdd(sa + 9, g0b[c]);
    atomicAdd(sa + 19, g1b[c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kL1CG * 20; i += blockDim.x) {
    const int c = blockIdx.y * kL1CG + i / 20, k = i % 20;
    const float v = sacc[i];
    if (k < 9) atomicAdd(dw0 + c * 9 + k, v);
    else if (k == 9) atomicAdd(db0 + c, v);
    else if (k < 19) atomicAdd(dw1 + c * 9 + (k - 10), v);
    else atomicAdd(db1 + c, v);
  }
}

}  // namespace lcasr

using namespace lcasr;

static bool cgroups_ok(int C) { return C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0; }

extern "C" int lcasr_subsample_dwconv_bwd_data(const void* dout, const float* w, int B, int64_t Tin, int Fin, int C, void* din,
                                               void* stream) {
  LCASR_CHECK_ARG(dout && w && din && B > 0 && Tin > 0 && Fin > 0, "subsample_dwconv_bwd_data: bad arguments");
  LCASR_CHECK_ARG(cgroups_ok(C), "subsample_dwconv_bwd_data: C=%d: C/8 must divide 256", C);
  const int64_t Tout = (Tin - 1) / 2 + 1;
  const int Fout = (Fin - 1) / 2 + 1;
  LCASR_CHECK_ARG(B <= 65535 && ceil_div(Tin, 2 * kBdTB) <= 0x7fffffff, "subsample_dwconv_bwd_data: grid too large");
  dim3 grid((unsigned)ceil_div(ceil_div(Tin, 2), kBdTB), (unsigned)B);
  if (C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0)
    subsample_dwconv_bwd_data_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, w, Tin, Fin, C, Tout, Fout, (bf16*)din);
  else
    subsample_dwconv_bwd_data_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, w, Tin, Fin, C, Tout, Fout, (bf16*)din);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_subsample_dwconv_bwd_weight(const void* in, const void* dout, int B, int64_t Tin, int Fin, int C, float* dw,
                                                 float* db, void* stream) {
  LCASR_CHECK_ARG(in && dout && dw && db && B > 0 && Tin > 0 && Fin > 0, "subsample_dwconv_bwd_weight: bad arguments");
  LCASR_CHECK_ARG(cgroups_ok(C) && C * 40 <= 48 * 1024, "subsample_dwconv_bwd_weight: C=%d unsupported", C);
  const int64_t Tout = (Tin - 1) / 2 + 1;
  const int Fout = (Fin - 1) / 2 + 1;
  LCASR_CHECK_ARG(B <= 65535, "subsample_dwconv_bwd_weight: batch too large");
  int64_t gx = ceil_div((int64_t)kNumSMs * 8, B);  // ~8 CTAs per SM in total
  if (gx > ceil_div(Tout, kBwTB)) gx = ceil_div(Tout, kBwTB);
  dim3 grid((unsigned)gx, (unsigned)B);
  static const bool narrow = getenv("LCASR_SUBSAMPLE_V4") != nullptr;  // measured slower (856 vs 766 us): kept for A/B runs
  if (narrow && C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0)
    subsample_dwconv_bwd_weight_kernel<4><<<grid, 256, (size_t)C * 40, (cudaStream_t)stream>>>(
        (const bf16*)in, (const bf16*)dout, Tin, Fin, C, Tout, Fout, dw, db);
  else
    subsample_dwconv_bwd_weight_kernel<8><<<grid, 256, (size_t)C * 40, (cudaStream_t)stream>>>(
        (const bf16*)in, (const bf16*)dout, Tin, Fin, C, Tout, Fout, dw, db);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_subsample_conv0_bwd(const float* spec, const float* w, const float* b, const void* ds1, int B, int F,
                                         int64_t T, int C, float* dw, float* db, void* stream) {
  LCASR_CHECK_ARG(spec && w && b && ds1 && dw && db && B > 0 && F > 0 && T > 0, "subsample_conv0_bwd: bad arguments");
  LCASR_CHECK_ARG(cgroups_ok(C), "subsample_conv0_bwd: C=%d: C/8 must divide 256", C);
  const int64_t T1 = (T - 1) / 2 + 1;
  const int F1 = (F - 1) / 2 + 1;
  const size_t smem = ((size_t)(2 * kC0bTT + 1) * (F + 2) + (size_t)C * 10) * sizeof(float);
  LCASR_CHECK_ARG(smem <= 48 * 1024, "subsample_conv0_bwd: F=%d, C=%d need too much shared memory", F, C);
  LCASR_CHECK_ARG(B <= 65535, "subsample_conv0_bwd: batch too large");
  int64_t gx = ceil_div((int64_t)kNumSMs * 4, B);  // ~4 CTAs per SM in total, each walking several tiles
  if (gx > ceil_div(T1, kC0bTT)) gx = ceil_div(T1, kC0bTT);
  dim3 grid((unsigned)gx, (unsigned)B);
  subsample_conv0_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(spec, w, b, (const bf16*)ds1, F, T, C, T1, F1, dw, db);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_subsample_l1_bwd(const float* spec, const float* w0, const float* b0, const float* w1, const void* dd1,
                                      int B, int F, int64_t T, int C, float* dw0, float* db0, float* dw1, float* db1,
                                      void* stream) {
  LCASR_CHECK_ARG(spec && w0 && b0 && w1 && dd1 && dw0 && db0 && dw1 && db1 && B > 0 && F > 0 && T > 0,
                  "subsample_l1_bwd: bad arguments");
  LCASR_CHECK_ARG(C % kL1CG == 0, "subsample_l1_bwd: conv_channels=%d must be a multiple of %d (use the unfused kernels)", C, kL1CG);
  const int64_t T1 = (T - 1) / 2 + 1, T2 = (T1 - 1) / 2 + 1;
  const int F1 = (F - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  const int FWp = std::max(F + 2, 4 * ((F1 + 1) / 2) + 1);  // patch pitch: every 5-wide block read stays inside the row
  const size_t smem = ((size_t)2 * (2 * kL1R + 1) * FWp + (size_t)kL1CG * 20) * sizeof(float);
  LCASR_CHECK_ARG(smem <= 48 * 1024, "subsample_l1_bwd: feat_in=%d too large for the staged patch", F);
  LCASR_CHECK_ARG(T2 < ((int64_t)1 << 30), "subsample_l1_bwd: too many frames");
  LCASR_CHECK_ARG(B <= 65535, "subsample_l1_bwd: batch too large");
  const int ncg = C / kL1CG;
  int64_t gx = ((int64_t)kNumSMs * 2) / ((int64_t)B * ncg);  // one wave of 2 resident CTAs per SM (rounded DOWN: a few extra CTAs
  if (gx < 1) gx = 1;                                        // would run as a second wave), each walking several tiles
  if (gx > ceil_div(T1, kL1R)) gx = ceil_div(T1, kL1R);
  dim3 grid((unsigned)gx, (unsigned)ncg, (unsigned)B);
  subsample_l1_bwd_fused_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(spec, w0, b0, w1, (const bf16*)dd1, F, T, C, T1, F1, T2, F2,
                                                                          FWp, dw0, db0, dw1, db1);
  LCASR_LAUNCH_CHECK();
  return 0;
}
