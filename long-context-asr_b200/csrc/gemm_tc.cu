// tcgen05 GEMM  out = epi(A[M,K] . W[N,K]^T), bf16 operands, fp32 accumulation in tensor memory.
//
// Persistent, warp-specialised (one CTA per SM, 192 threads):
//   warp 0   TMA producer   : cp.async.bulk.tensor 128x64 (A) and BNx64 (W) bf16 boxes, 128B swizzle,
//                             into a STAGES-deep shared-memory ring (mbarrier full/empty pairs)
//   warp 1   MMA issuer     : one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                             (M=128, N=BN, K=16) x4 per stage into one of two TMEM accumulators,
//                             tcgen05.commit releases the smem slot / publishes the accumulator
//   warps 2-5 epilogue      : tcgen05.ld 32 lanes x 32 columns -> bias/activation/residual -> global;
//                             overlaps the next tile's main loop through the 2nd TMEM accumulator
// Both operands are K-major (activations [M,K] and nn.Linear weights [N,K] are K-contiguous), so no
// transposes are needed anywhere on the path.  Out-of-range rows / K tail are zero-filled by TMA.
// Tensor-bound: algorithmic FLOPs = 2*M*N*K.
#include "common.cuh"
#include <cstdlib>
#include "sm100_ptx.cuh"

namespace lcasr {

using namespace ptx;

constexpr int TG_BM = 128, TG_BK = 64;
// four epilogue warps, one per TMEM lane quarter (eight — two per quarter, half of the columns each — measured neutral:
// the bound of the bf16-output GEMMs was the global store pattern, see tg_stage_chunk64)
// (CTA pairs + bf16 outputs: eight, two per quarter taking half of the columns each — with the stores gone to TMA the
// remaining epilogue cost is the per-warp chain TMEM load -> convert -> stage, which two warps per quarter overlap)
// fp32 outputs with a residual are bound by the memory-level parallelism of the residual loads (4 KB in flight per
// warp): CTA pairs (5-stage ring, room for 8 transpose tiles) run eight epilogue warps there as well.
// (W8: chosen for short-K fp32 GEMMs, whose tile time IS the epilogue; long-K ones measured 4 % faster with four)
template <typename TOut, int CG, bool W8 = true> struct TgWarps {
  static constexpr int EPI = (CG == 2 && (sizeof(TOut) == 2 || W8)) ? 8 : 4, THREADS = 64 + 32 * EPI;
};

// CG = CTAs per MMA (tcgen05 cta_group): 1, or 2 = a CTA pair computes a 256 x BN tile, each CTA staging its 128 rows
// of A and HALF of the B tile -> 2/3 of the shared-memory fill traffic per FLOP and a deeper ring
// EPI_BYTES: epilogue staging behind the operand ring — bf16 outputs: one 4 KB TMA-store box (32 rows x 64 columns) per
// epilogue warp; fp32 outputs: two 4 KB boxes (32 x 32 fp32) per warp, which alternate as TMA-load target of the residual
// chunk and, after the in-place add, TMA-store source.  The ring gives up stages where both do not fit.
template <int BN, int CG, int EPI_BYTES> struct TgCfgE {
  static constexpr int A_BYTES = TG_BM * TG_BK * 2;
  static constexpr int B_BYTES = (BN / CG) * TG_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int DEF_STAGES = BN == 256 ? (CG == 2 ? 5 : 4) : (CG == 2 ? 8 : 6);
  static constexpr int FIT_STAGES = (232448 - 1024 /*align slack*/ - (2 * BN * 4 + 512) /*static: bias, barriers*/ - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = DEF_STAGES < FIT_STAGES ? DEF_STAGES : FIT_STAGES;
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered fp32 accumulator (power of two: 256 / 512)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024;
  static_assert(STAGES >= 3, "gemm_tc: operand ring too shallow");
};
template <int BN, typename TOut, int CG, bool W8> using TgCfgFor =
    TgCfgE<BN, CG, TgWarps<TOut, CG, W8>::EPI * (sizeof(TOut) == 4 ? 8192 : 4096)>;

struct TgEpilogue {
  const float* bias;   // [N] or null
  const float* resid;  // [M,N] fp32 or null
  float alpha;
  int act;
  bf16* pre_out = nullptr;  // training: the pre-activation acc+bias [M,N] is stored as well (bf16 outputs only)
  int debug = 0;            // LCASR_GEMM_DEBUG (profiling only): 1 = skip the global stores, 2 = skip the whole epilogue, 3 / 4 = fp32: no stores / no residual
  // EPI == TG_EPI_ROPE (fused rotary, attention.py:499-507 / rotary_emb.py:61-73): output columns < rope_cols are rotated in
  // fp32 before the bf16 store.  The weight rows of every q / k head were interleaved on the host (new 2i <- old i,
  // new 2i+1 <- old i + Dh/2), so a rotation pair is two ADJACENT columns: y[2i] = x[2i] c_i - x[2i+1] s_i,
  // y[2i+1] = x[2i+1] c_i + x[2i] s_i with c/s = tables[i, pos] (pair-major, lcasr_rope_table_t), pos = row % rope_n.  (q.k is invariant under a common
  // permutation of the head dimension, so attention sees exactly the reference's scores.)
  const float* rope_cos = nullptr;
  const float* rope_sin = nullptr;
  int rope_cols = 0, rope_dh = 0;
  int64_t rope_n = 1;
};

enum { TG_EPI_STD = 0, TG_EPI_ROPE = 1, TG_EPI_GLU = 2 };

// Direct epilogue (bf16 outputs, no residual): lane == row, 64 contiguous bytes per lane and chunk.
// Measured faster than the transposed variant below for 2-byte outputs (923 vs 654 TFLOP/s on the
// 768->3072 GELU GEMM): the extra shared-memory round trip costs more than the partially filled lines.
__device__ __forceinline__ void tg_store_chunk_direct(const uint32_t (&r)[32], int64_t row, int col0, int64_t M, int N,
                                                      const TgEpilogue& ep, const float* __restrict__ bias_chunk,
                                                      bf16* __restrict__ out) {
  if (row >= M) return;
  const int ngroups = min(4, (N - col0) >> 3);  // 8 columns per group; N % 8 == 0
  float y[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(r[i]);
  if (ep.bias) {
    const float4* bp = reinterpret_cast<const float4*>(bias_chunk);  // shared memory (broadcast reads)
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < ngroups) {
        const float4 b0 = bp[2 * g], b1 = bp[2 * g + 1];
        y[8 * g + 0] += b0.x; y[8 * g + 1] += b0.y; y[8 * g + 2] += b0.z; y[8 * g + 3] += b0.w;
        y[8 * g + 4] += b1.x; y[8 * g + 5] += b1.y; y[8 * g + 6] += b1.z; y[8 * g + 7] += b1.w;
      }
  }
  if (ep.pre_out) {
    bf16* pp = ep.pre_out + row * N + col0;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < ngroups) Vec8<bf16>::store(pp + 8 * g, *reinterpret_cast<const float(*)[8]>(&y[8 * g]));
    // the activation below is applied to the ROUNDED pre-activation, i.e. exactly what the backward will see
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = __bfloat162float(__float2bfloat16_rn(y[i]));
  }
  if (ep.act == LCASR_ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float x = y[i];
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      y[i] = 0.5f * x * (1.0f + tanh_approx(u));
    }
  } else if (ep.act == LCASR_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = silu_fast(y[i]);
  }
  bf16* op = out + row * N + col0;
  if (ep.debug == 1 && y[0] != 12345.678f) return;  // profiling: everything but the stores (the test keeps y alive)
#pragma unroll
  for (int g = 0; g < 4; ++g)
    if (g < ngroups) Vec8<bf16>::store(op + 8 * g, *reinterpret_cast<const float(*)[8]>(&y[8 * g]));
}

// fp32 outputs: residual in and result out through TMA (run 83: the epilogue of these GEMMs cost 12.2 us per 256x256 pair
// tile whatever K — its coalesced residual loads (+14 us over the launch) and its stores (+10 us) each occupied the eight
// epilogue warps in turn; TMA stores alone brought the residual-free case to the main-loop bound, register loads of the
// residual in the accumulator's row layout made the other case worse).
// box: this warp's 32 x 32 fp32 shared-memory box (128-byte rows, 16-byte slots XOR-swizzled by row % 8 — the
// SWIZZLE_128B layout both TMA directions use).  With a residual the box holds its chunk (TMA-loaded) and is updated in
// place; lane == row, 8 lanes with distinct row % 8 cover 8 distinct slots per wavefront: conflict-free both ways.
__device__ __forceinline__ void tg_stage_chunk_f32(const uint32_t (&r)[32], int lane, uint32_t box, const TgEpilogue& ep,
                                                   const float* __restrict__ bias_chunk) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                           __uint_as_float(r[4 * q + 3]));
    if (ep.bias) {
      const float4 b = *reinterpret_cast<const float4*>(bias_chunk + 4 * q);  // shared memory (broadcast reads)
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (ep.act == LCASR_ACT_GELU_TANH) {
      float* pv = reinterpret_cast<float*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = pv[i];
        const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
        pv[i] = 0.5f * x * (1.0f + tanh_approx(u));
      }
    } else if (ep.act == LCASR_ACT_SILU) {
      float* pv = reinterpret_cast<float*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = __fdividef(pv[i], 1.0f + __expf(-pv[i]));
    }
    const uint32_t addr = box + lane * 128 + ((q ^ (lane & 7)) << 4);
    if (ep.resid) {
      float4 res;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(res.x), "=f"(res.y), "=f"(res.z), "=f"(res.w) : "r"(addr) : "memory");
      v.x = fmaf(ep.alpha, v.x, res.x); v.y = fmaf(ep.alpha, v.y, res.y);
      v.z = fmaf(ep.alpha, v.z, res.z); v.w = fmaf(ep.alpha, v.w, res.w);
    }
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

// bf16 outputs through TMA stores.  Measured on the 768 -> 3072 GELU GEMM: main loop alone 1520 TFLOP/s, + TMEM loads
// and the activation 1330, + per-lane 16-byte global stores (one row per lane: 32 partial lines per instruction) 944.
// Here a warp converts 64 columns of its 32 rows into a 128B-swizzled shared-memory box (conflict-free: 8 lanes with
// distinct row%8 cover 8 distinct 16-byte slots per wavefront) and one lane issues a single cp.async.bulk.tensor
// store for the box: full lines, no LSU work, rows / columns beyond the tensor clipped by the hardware.
template <int EPI>
__device__ __forceinline__ void tg_stage_chunk64(const uint32_t (&r0)[32], const uint32_t (&r1)[32], int lane, uint32_t stage_addr,
                                                 const TgEpilogue& ep, const float* __restrict__ bias_chunk, bool pre_pass,
                                                 int col0 = 0, const float* __restrict__ cos_row = nullptr,
                                                 const float* __restrict__ sin_row = nullptr) {
  // col0 (EPI == ROPE): global column of r0[0]; cos_row / sin_row: this lane's table rows (position of its token)
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    float y[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(hh == 0 ? r0[i] : r1[i]);
    if (ep.bias) {
      const float4* bp = reinterpret_cast<const float4*>(bias_chunk + 32 * hh);  // shared memory (broadcast reads)
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b = bp[g];
        y[4 * g] += b.x; y[4 * g + 1] += b.y; y[4 * g + 2] += b.z; y[4 * g + 3] += b.w;
      }
    }
    if constexpr (EPI == TG_EPI_ROPE) {
      const int colh = col0 + 32 * hh;
      if (colh < ep.rope_cols) {  // warp-uniform: a q / k column block (v passes through)
        // tables are pair-major [Dh/2][rope_n]: the warp's lanes are consecutive token positions, so each of the 16 pair
        // indices of this block is ONE coalesced line (position-major tables cost 32 lines per load: measured +60 us per GEMM)
        const int64_t j0 = (int64_t)((colh % ep.rope_dh) >> 1) * ep.rope_n;
        float cv[16], sv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          cv[i] = __ldg(cos_row + j0 + i * ep.rope_n);
          sv[i] = __ldg(sin_row + j0 + i * ep.rope_n);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = y[2 * i], b = y[2 * i + 1];
          y[2 * i] = fmaf(a, cv[i], -b * sv[i]);
          y[2 * i + 1] = fmaf(b, cv[i], a * sv[i]);
        }
      }
    }
    if (!pre_pass) {
      if (ep.pre_out) {  // the activation is applied to the ROUNDED pre-activation, i.e. what the backward will see
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __bfloat162float(__float2bfloat16_rn(y[i]));
      }
      if (ep.act == LCASR_ACT_GELU_TANH) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = y[i];
          const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
          y[i] = 0.5f * x * (1.0f + tanh_approx(u));
        }
      } else if (ep.act == LCASR_ACT_SILU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = silu_fast(y[i]);
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {  // 16-byte slot j = 4*hh + g of this lane's 128-byte row, XOR-swizzled by row % 8
      uint4 v;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) h2[i] = __floats2bfloat162_rn(y[8 * g + 2 * i], y[8 * g + 2 * i + 1]);
      const uint32_t addr = stage_addr + lane * 128 + (((4 * hh + g) ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
}

// EPI == TG_EPI_GLU (torch.nn.functional.glu fused into pointwise_conv1, convolution.py:105-107): the weight rows were
// packed on the host so that every 64-column block of the tile holds 32 value channels followed by THEIR 32 gate channels;
// out[:, ch] = (acc_v + b_v) * sigmoid(acc_g + b_g) — 32 bf16 outputs per block, two blocks fill one 64-column store box.
__device__ __forceinline__ void tg_stage_glu32(const uint32_t (&rv)[32], const uint32_t (&rg)[32], int lane, uint32_t stage_addr,
                                               int half_sel, const TgEpilogue& ep, const float* __restrict__ bias_chunk) {
  float y[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float v = __uint_as_float(rv[i]), g = __uint_as_float(rg[i]);
    if (ep.bias) { v += bias_chunk[i]; g += bias_chunk[32 + i]; }
    y[i] = v * sigmoid_fast(g);
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {  // 16-byte slot 4*half_sel + g of this lane's 128-byte row, XOR-swizzled by row % 8
    uint4 v;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h2[i] = __floats2bfloat162_rn(y[8 * g + 2 * i], y[8 * g + 2 * i + 1]);
    const uint32_t addr = stage_addr + lane * 128 + (((4 * half_sel + g) ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}

template <int BN, typename TOut, int CG, bool W8, int EPI = TG_EPI_STD>
__global__ void __launch_bounds__(TgWarps<TOut, CG, W8>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, int64_t M, int N, int K,
               TgEpilogue ep, TOut* out) {
  // tmC / tmC2 (bf16 outputs only): store maps of `out` / `ep.pre_out`, 64-column x 32-row boxes, 128B swizzle
  using Cfg = TgCfgFor<BN, TOut, CG, W8>;
  // CG == 2: launched as clusters of 2 CTAs; `rank` 0 is the leader (issues every MMA, owns the full / tempty barriers
  // that both CTAs signal), tiles are 256 x BN and indexed per cluster
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * Cfg::STAGES + 4];
  __shared__ uint32_t tmem_slot;
  constexpr int NEW = TgWarps<TOut, CG, W8>::EPI;
  __shared__ __align__(8) uint64_t rbars[sizeof(TOut) == 4 ? 2 * NEW : 1];  // fp32: residual-box "landed" barriers, two per epilogue warp
  __shared__ __align__(16) float bias_s[2][BN];          // the tile's bias slice, staged while the main loop still runs
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle atoms need 1024B alignment
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (K + TG_BK - 1) / TG_BK;
  const int tiles_n = (N + BN - 1) / BN;
  const int64_t tiles_m = (M + CG * TG_BM - 1) / (CG * TG_BM);
  const int64_t total_tiles = tiles_m * tiles_n;
  const int64_t tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;  // per cluster

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), CG * NEW); }
    if (sizeof(TOut) == 4)
      for (int i = 0; i < 2 * NEW; ++i) mbar_init(smem_u32(&rbars[i]), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) { tmem_alloc_cg2(smem_u32(&tmem_slot), Cfg::TMEM_COLS); tmem_relinquish_cg2(); }
    else { tmem_alloc(smem_u32(&tmem_slot), Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer ----------------
      int stage = 0; uint32_t phase = 0;
      const uint32_t full0 = CG == 2 ? mapa_shared(full_bar(0), 0) : full_bar(0);  // the leader's full barriers
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        const int m_idx = (int)(tile / tiles_n) * (CG * TG_BM) + (int)rank * TG_BM;
        const int n_idx = (int)(tile % tiles_n) * BN + (int)rank * (BN / CG);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          if constexpr (CG == 2) {
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes land here
            tma_load_2d_cg2(sa, &tmA, full0 + 8u * stage, kb * TG_BK, m_idx);
            tma_load_2d_cg2(sa + Cfg::A_BYTES, &tmB, full0 + 8u * stage, kb * TG_BK, n_idx);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            tma_load_2d(sa, &tmA, full_bar(stage), kb * TG_BK, m_idx);
            tma_load_2d(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * TG_BK, n_idx);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ---------------- MMA issuer (leader CTA only) ----------------
      constexpr uint32_t idesc = make_idesc_bf16(CG * TG_BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc_kmajor(sa, 1024, kLayoutSW128);
          const uint64_t bdesc = make_smem_desc_kmajor(sa + Cfg::A_BYTES, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < TG_BK / 16; ++k) {  // +32 bytes (= 2 in >>4 units) per K=16 step inside the swizzle row
            if constexpr (CG == 2) umma_f16_ss_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            else umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if constexpr (CG == 2) umma_commit_cg2(empty_bar(stage), 3);  // frees the slot in BOTH CTAs
          else umma_commit(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 2) umma_commit_cg2(tfull_bar(acc), 3);      // each CTA drains its own 128 accumulator rows
        else umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {  // ---------------- epilogue warps 2..5 ----------------
    const int lane_base = (warp & 3) * 32;  // TMEM lanes this warp may touch
    const int half = (warp - 2) >> 2;       // NEW == 8: which half of the tile's columns this warp drains
    constexpr int CPW = (BN / 32) / (NEW / 4);  // 32-column chunks per warp
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t rsel = 0, rphase = 0;  // fp32: box in use, parity bits of the two residual barriers
    const uint32_t tempty0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : 0;  // the leader's tempty barriers
    for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
      const int64_t m_idx = (tile / tiles_n) * (CG * TG_BM) + rank * TG_BM;
      const int n_idx = (int)(tile % tiles_n) * BN;
      if (ep.bias) {  // stage this tile's bias slice now: its latency hides behind the wait for the accumulator
        const int e = (warp - 2) * 32 + lane;
        for (int i = e; i < BN; i += 32 * NEW) {
          const int col = n_idx + i;
          bias_s[acc][i] = col < N ? __ldg(ep.bias + col) : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory");  // the epilogue warps only
      }
      const int64_t row0 = m_idx + lane_base;
      // fp32: this warp's two boxes alternate per chunk.  Chunk c's residual is requested one chunk ahead (at tile start for
      // the first one), as soon as the box's previous bulk store has been read out.
      const uint32_t box0 = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES + (warp - 2) * 8192;
      const uint32_t rbar0 = smem_u32(&rbars[sizeof(TOut) == 4 ? 2 * (warp - 2) : 0]);
      const int c_begin = half * CPW;
      if constexpr (sizeof(TOut) == 4) {
        // (the boxes bound the bytes in flight — one 4 KB chunk ahead per warp; L2 prefetches of the next tile's residual
        // boxes, cp.async.bulk.prefetch.tensor, were measured: no gain at K = 768, 6-12 % slower for long K)
        if (ep.resid && lane == 0 && n_idx + c_begin * 32 < N) {
          tma_store_wait_read();
          mbar_arrive_expect_tx(rbar0 + 8u * rsel, 4096);
          tma_load_2d(box0 + 4096u * rsel, &tmC2, rbar0 + 8u * rsel, n_idx + c_begin * 32, (int)row0);
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)lane_base << 16) + acc * BN;
#pragma unroll 1
      for (int c = half * CPW; c < (half + 1) * CPW; ++c) {
        if (n_idx + c * 32 >= N || ep.debug == 2) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_addr + c * 32, r);
        if constexpr (sizeof(TOut) == 2 && EPI == TG_EPI_GLU) {
          // chunks come in groups of four: (value, gate) (value, gate) -> one 64-column store box of the [M, N/2] output
          const uint32_t stg = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES + (warp - 2) * 4096;
          uint32_t r1[32];
          tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, r1);
          tmem_wait_ld();
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          tg_stage_glu32(r, r1, lane, stg, 0, ep, &bias_s[acc][c * 32]);
          if (n_idx + (c + 2) * 32 < N) {  // warp-uniform (N % 64 == 0: a block is complete or absent)
            tmem_ld_32x32b_x32(t_addr + (c + 2) * 32, r);
            tmem_ld_32x32b_x32(t_addr + (c + 3) * 32, r1);
            tmem_wait_ld();
            tg_stage_glu32(r, r1, lane, stg, 1, ep, &bias_s[acc][(c + 2) * 32]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {  // columns beyond N/2 are clipped by the tensor map
            tma_store_2d(&tmC, stg, (n_idx + c * 32) >> 1, (int)row0);
            tma_store_commit();
          }
          c += 3;  // four chunks per iteration (CPW and the warp's first chunk are multiples of 4)
        } else if constexpr (sizeof(TOut) == 2) {
          if (ep.debug || (c & 1)) {  // profiling variants keep the direct path; odd chunks are drained with their even twin
            tmem_wait_ld();
            if (ep.debug) tg_store_chunk_direct(r, row0 + lane, n_idx + c * 32, M, N, ep, &bias_s[acc][c * 32], out);
            continue;
          }
          uint32_t r1[32];
          tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, r1);  // BN is a multiple of 64: the twin chunk exists
          const float* cos_row = nullptr;
          const float* sin_row = nullptr;
          if constexpr (EPI == TG_EPI_ROPE) {  // this lane's token position -> its table rows (clamped for rows beyond M)
            const int64_t rr = row0 + lane < M ? row0 + lane : M - 1;
            const int64_t pos = rr % ep.rope_n;
            cos_row = ep.rope_cos + pos;  // pair-major tables: element (j, pos) at j * rope_n + pos
            sin_row = ep.rope_sin + pos;
          }
          tmem_wait_ld();
          const uint32_t stg = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES + (warp - 2) * 4096;
          for (int pass = ep.pre_out ? 0 : 1; pass < 2; ++pass) {
            if (lane == 0) tma_store_wait_read();  // the previous box has been read out of the staging tile
            __syncwarp();
            tg_stage_chunk64<EPI>(r, r1, lane, stg, ep, &bias_s[acc][c * 32], pass == 0, n_idx + c * 32, cos_row, sin_row);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(pass == 0 ? &tmC2 : &tmC, stg, n_idx + c * 32, (int)row0);
              tma_store_commit();
            }
          }
        } else {
          const uint32_t box = box0 + 4096u * rsel, other = box0 + 4096u * (rsel ^ 1);
          if (lane == 0) {
            tma_store_wait_read();  // the other box (stored from in the previous chunk) has been read out
            if (ep.resid && c + 1 < (half + 1) * CPW && n_idx + (c + 1) * 32 < N) {  // next chunk's residual, in flight during this one
              mbar_arrive_expect_tx(rbar0 + 8u * (rsel ^ 1), 4096);
              tma_load_2d(other, &tmC2, rbar0 + 8u * (rsel ^ 1), n_idx + (c + 1) * 32, (int)row0);
            }
          }
          if (ep.resid) {
            mbar_wait(rbar0 + 8u * rsel, (rphase >> rsel) & 1u);
            rphase ^= 1u << rsel;
          }
          tmem_wait_ld();
          __syncwarp();
          tg_stage_chunk_f32(r, lane, box, ep, &bias_s[acc][c * 32]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && ep.debug != 3) {
            tma_store_2d(&tmC, box, n_idx + c * 32, (int)row0);
            tma_store_commit();
          }
          rsel ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();  // bulk stores complete before the CTA retires
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its TMEM
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---- host side ----------------------------------------------------------------------------------

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

static int make_tmap_2d_any(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t rows, uint64_t cols,
                            uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  return make_tmap_2d_any(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rows, cols, row_pitch_bytes, box_rows, box_cols, swizzle);
}

static int make_tmap_2d_any(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t rows, uint64_t cols,
                            uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  if (!fn) return set_error(LCASR_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LCASR_E_CUDA, "cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu pitch=%llu box=%ux%u",
                     (int)r, base, (unsigned long long)rows, (unsigned long long)cols,
                     (unsigned long long)row_pitch_bytes, box_rows, box_cols);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes,
                      uint64_t pitch2_bytes, uint32_t box0, uint32_t box1, uint32_t box2, CUtensorMapSwizzle swizzle) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  if (!fn) return set_error(LCASR_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LCASR_E_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d): base=%p dims=%llu,%llu,%llu", (int)r, base,
                     (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
  return 0;
}

template <int BN, typename TOut, int CG, bool W8 = true, int EPI = TG_EPI_STD>
static int launch_tc(const void* A, const void* W, int64_t M, int N, int K, const TgEpilogue& ep, void* out,
                     cudaStream_t st) {
  using Cfg = TgCfgFor<BN, TOut, CG, W8>;
  CUtensorMap tmA, tmB;
  LCASR_TRY(make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)K * 2, TG_BM, TG_BK, CU_TENSOR_MAP_SWIZZLE_128B));
  LCASR_TRY(make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, BN / CG, TG_BK, CU_TENSOR_MAP_SWIZZLE_128B));
  CUtensorMap tmC = tmA, tmC2 = tmA;
  if (sizeof(TOut) == 4) {  // 32 x 32 fp32 boxes (128-byte rows)
    LCASR_TRY(make_tmap_2d_any(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, out, (uint64_t)M, (uint64_t)N, (uint64_t)N * 4, 32, 32,
                               CU_TENSOR_MAP_SWIZZLE_128B));
    if (ep.resid)
      LCASR_TRY(make_tmap_2d_any(&tmC2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ep.resid, (uint64_t)M, (uint64_t)N, (uint64_t)N * 4, 32, 32,
                                 CU_TENSOR_MAP_SWIZZLE_128B));
  } else {
    const uint64_t No = EPI == TG_EPI_GLU ? (uint64_t)N / 2 : (uint64_t)N;  // GLU halves the output width
    LCASR_TRY(make_tmap_2d_bf16(&tmC, out, (uint64_t)M, No, No * 2, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
    if (ep.pre_out)
      LCASR_TRY(make_tmap_2d_bf16(&tmC2, ep.pre_out, (uint64_t)M, (uint64_t)N, (uint64_t)N * 2, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, TOut, CG, W8, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set.mark(attr_dev);
  }
  const int64_t tiles = ceil_div(M, CG * TG_BM) * ceil_div(N, BN);
  const int64_t units = kNumSMs / CG;  // CTAs (CG == 1) or CTA pairs: one per SM / per TPC, persistent
  const int grid = (int)(tiles < units ? tiles : units) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TgWarps<TOut, CG, W8>::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  LCASR_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, TOut, CG, W8, EPI>, tmA, tmB, tmC, tmC2, M, N, K, ep, (TOut*)out));
  LCASR_LAUNCH_CHECK();
  return 0;
}

template <int BN, typename TOut, int EPI = TG_EPI_STD>
static int launch_tc_cg(const void* A, const void* W, int64_t M, int N, int K, const TgEpilogue& ep, void* out,
                        cudaStream_t st, int prefer_single = 0) {
  // CTA pairs (256-row tiles) once there is enough work to fill the 74 pairs; LCASR_GEMM_CG=1|2 forces a choice (A/B runs)
  static const int force = getenv("LCASR_GEMM_CG") ? atoi(getenv("LCASR_GEMM_CG")) : 0;
  const bool pair = force ? force == 2 : (!prefer_single && ceil_div(M, 2 * TG_BM) * ceil_div(N, BN) >= kNumSMs / 2);
  if (!pair) return launch_tc<BN, TOut, 1, true, EPI>(A, W, M, N, K, ep, out, st);
  if (sizeof(TOut) == 4 && K > 1536) return launch_tc<BN, TOut, 2, false, EPI>(A, W, M, N, K, ep, out, st);
  return launch_tc<BN, TOut, 2, true, EPI>(A, W, M, N, K, ep, out, st);
}

static int gemm_tc_check_common(const void* A, const void* W, int64_t M, int N, int K, const float* bias, const void* out) {
  LCASR_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "gemm(tcgen05): K=%d and N=%d must be multiples of 8", K, N);
  LCASR_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)out & 15) == 0,
                  "gemm(tcgen05): A, W and out must be 16-byte aligned");
  LCASR_CHECK_ARG(M < (int64_t)1 << 31, "gemm(tcgen05): M too large");
  LCASR_CHECK_ARG(((uintptr_t)bias & 15) == 0, "gemm(tcgen05): bias must be 16-byte aligned");
  return 0;
}

int gemm_tc_launch(const void* A, const void* W, int64_t M, int N, int K, const float* bias, int act, const float* resid,
                   float alpha, void* out, int out_dtype, cudaStream_t st, void* pre_out) {
  LCASR_CHECK_ARG(!pre_out || (out_dtype == LCASR_BF16 && !resid && ((uintptr_t)pre_out & 15) == 0),
                  "gemm(tcgen05): the pre-activation output needs a bf16, residual-free epilogue");
  LCASR_TRY(gemm_tc_check_common(A, W, M, N, K, bias, out));
  LCASR_CHECK_ARG(((uintptr_t)resid & 15) == 0, "gemm(tcgen05): resid must be 16-byte aligned");
  LCASR_CHECK_ARG(!resid || out_dtype == LCASR_F32, "gemm: a residual epilogue writes fp32");
  static const int debug = getenv("LCASR_GEMM_DEBUG") ? atoi(getenv("LCASR_GEMM_DEBUG")) : 0;
  TgEpilogue ep{bias, debug == 4 ? nullptr : resid, alpha, act, (bf16*)pre_out, debug};  // (4: profiling — no residual loads)
  bool wide = (N % 256 == 0) || N > 512;
  int single = 0;
  if (wide && N % 128 == 0) {
    // Few rows (a sequence-parallel rank's block, short recordings): pick the tile form that fills the SMs best.  E.g.
    // 2048 x 768 is 48 tiles of 128 x 256 for 148 SMs; 2048 x 3072 is 96 CTA-pair tiles for 74 pairs (two waves, the second
    // 30 % full) but 384 single-CTA tiles of 128 x 128 (2.6 waves).  Large problems keep the tuned choice (all forms > 0.9).
    auto fill = [](int64_t t, int64_t units) { return (double)t / (double)(ceil_div(t, units) * units); };
    const int64_t mt = ceil_div(M, TG_BM), mt2 = ceil_div(M, 2 * TG_BM), nt256 = ceil_div(N, 256), nt128 = ceil_div(N, 128);
    const bool pair_ok = mt2 * nt256 >= kNumSMs / 2;
    const double cur = pair_ok ? fill(mt2 * nt256, kNumSMs / 2) : fill(mt * nt256, kNumSMs);
    const double narrow = 0.9 * fill(mt * nt128, kNumSMs);  // 128-column tiles re-read A twice as often
    const double single256 = 0.97 * fill(mt * nt256, kNumSMs);
    if (narrow > 1.12 * cur && narrow >= single256) { wide = false; single = 1; }
    else if (pair_ok && single256 > 1.12 * cur) single = 1;
  }
  if (out_dtype == LCASR_BF16)
    return wide ? launch_tc_cg<256, bf16>(A, W, M, N, K, ep, out, st, single) : launch_tc_cg<128, bf16>(A, W, M, N, K, ep, out, st, single);
  return wide ? launch_tc_cg<256, float>(A, W, M, N, K, ep, out, st, single) : launch_tc_cg<128, float>(A, W, M, N, K, ep, out, st, single);
}

// qkv projection with the rotary embedding applied in the epilogue (see TgEpilogue): out [M, N] bf16, columns
// [0, rope_cols) rotated with tables cos/sin [rope_n, dh/2] (position = row % rope_n), the rest (v) stored as is.
int gemm_tc_launch_rope(const void* A, const void* W, int64_t M, int N, int K, const float* cos_t, const float* sin_t,
                        int64_t rope_n, int rope_cols, int dh, void* out, cudaStream_t st) {
  LCASR_TRY(gemm_tc_check_common(A, W, M, N, K, nullptr, out));
  LCASR_CHECK_ARG(cos_t && sin_t && rope_n > 0, "gemm_rope: NULL tables");
  LCASR_CHECK_ARG(dh >= 32 && dh % 32 == 0 && rope_cols % 64 == 0 && rope_cols <= N && N % 64 == 0,
                  "gemm_rope: head_dim %d / rotated columns %d / N %d not supported", dh, rope_cols, N);
  TgEpilogue ep{nullptr, nullptr, 0.f, LCASR_ACT_NONE, nullptr, 0};
  ep.rope_cos = cos_t; ep.rope_sin = sin_t; ep.rope_cols = rope_cols; ep.rope_dh = dh; ep.rope_n = rope_n;
  const bool wide = (N % 256 == 0) || N > 512;
  return wide ? launch_tc_cg<256, bf16, TG_EPI_ROPE>(A, W, M, N, K, ep, out, st)
              : launch_tc_cg<128, bf16, TG_EPI_ROPE>(A, W, M, N, K, ep, out, st);
}

// pointwise_conv1 + GLU: W [N = 2d, K] with rows packed in 64-row blocks (32 value channels, then their 32 gate channels),
// bias packed the same way; out [M, d] bf16.
int gemm_tc_launch_glu(const void* A, const void* W, int64_t M, int N, int K, const float* bias, void* out, cudaStream_t st) {
  LCASR_TRY(gemm_tc_check_common(A, W, M, N, K, bias, out));
  LCASR_CHECK_ARG(N % 64 == 0 && ((N % 256 == 0) || N > 512), "gemm_glu: N=%d needs the 256-column tile form (N %% 64 == 0, N > 512 or N %% 256 == 0)", N);
  TgEpilogue ep{bias, nullptr, 0.f, LCASR_ACT_NONE, nullptr, 0};
  return launch_tc_cg<256, bf16, TG_EPI_GLU>(A, W, M, N, K, ep, out, st);
}

}  // namespace lcasr
