// tcgen05 GEMM for the TRAINING path: the same persistent warp-specialised pipeline as gemm_tc.cu
// (TMA producer warp, single-thread tcgen05.mma issuer, 4 epilogue warps, double-buffered TMEM
// accumulator), generalised in the three ways the backward pass needs:
//
//   * either operand may be MN-major: the stored matrix is [K, M] (resp. [K, N]) row-major.  This is what
//     data-gradient (dX = dY . W, W stored [out,in] = [K,N]) and weight-gradient (dW = dY^T . X, both
//     operands stored token-major = [K,M] / [K,N]) GEMMs look like; the UMMA descriptors read such tiles
//     directly (a_major / b_major bits of the instruction descriptor), so no transposed copies exist.
//   * up to two batch dimensions with independent strides per tensor (4-D tensor maps): the five GEMMs of
//     the attention backward run over (head, recording) pairs in one launch, straight on the
//     [B,N,H,Dh] layout.
//   * split-K with fp32 atomic accumulation (weight gradients have K = all tokens of the batch and few
//     output tiles) and the backward epilogues: P = exp2(s*a - lse), dS = P o (dP - D) * a,
//     dH = dY o gelu'(h), plain scale.
//
// Tensor-bound: algorithmic FLOPs = 2*M*N*K per batch entry.
#include "common.cuh"
#include <cstdlib>
#include <mutex>
#include "sm100_ptx.cuh"

namespace lcasr {

using namespace ptx;

constexpr int TX_BM = 128, TX_BK = 64;
constexpr int TX_EPI_WARPS = 8;  // two epilogue warps per TMEM lane quarter, each draining half of the tile's columns
constexpr int TX_THREADS = 64 + 32 * TX_EPI_WARPS;

// CG = 2: CTA pairs (tcgen05 cta_group::2, see gemm_tc.cu): 256-row tiles, each CTA stages its 128 rows of A and
// HALF of the B tile; the leader issues the MMAs, commits are multicast, the peer's epilogue arrives remotely.
template <int BN, int CG = 1> struct TxCfg {
  static constexpr int STAGES = BN == 256 ? (CG == 2 ? 5 : 4) : (CG == 2 ? 7 : 6);
  static constexpr int A_BYTES = TX_BM * TX_BK * 2;
  static constexpr int B_BYTES = (BN / CG) * TX_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int OUT_STAGE_BYTES = TX_EPI_WARPS * 4096;  // bf16 outputs leave through TMA stores (see gemm_tc.cu)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_STAGE_BYTES + 1024;
};

struct TxParams {
  int64_t M;
  int N, K;
  int Nst;                      // columns stored per row: N rounded up to 8 (padding columns hold garbage)
  int nb1, nb2;                 // batch = b2 * nb1 + b1
  int ksplit, kper;             // K blocks per split
  int64_t ldo, so1, so2;        // output row pitch / batch strides (elements)
  const bf16* aux;              // bf16, indexed like the output (own pitch / strides)
  int aux32;                    // aux rows are 32-byte aligned (pointer, pitch, strides): 256-bit loads allowed
  int64_t ldx, sx1, sx2;
  const float* rowvec;          // fp32 per output row: rowvec[b2*sr2 + b1*sr1 + row]
  int64_t sr1, sr2;
  float alpha;
  int epi;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// kind::f16 instruction descriptor with both majorness bits ([15] a_major, [16] b_major; 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major operand tile in shared memory: [64 K rows][64 MN elements = 128 swizzled bytes] per 64-element
// MN sub-tile, sub-tiles 8192 B apart (LBO); 8 K rows = 1024 B (SBO); one K=16 step = 2048 B.
__device__ __forceinline__ uint64_t make_smem_desc_mnmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((8192u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)kLayoutSW128 << 61;
  return d;
}

__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  const float t = tanh_approx(u);
  const float du = k0 * (1.0f + 3.0f * k1 * x * x);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}
__device__ __forceinline__ float silu_grad(float x) {
  const float s = sigmoid_fast(x);
  return s * (1.0f + x * (1.0f - s));
}

// aux values (bf16, same indexing as the output) of one 32-column chunk of this lane's row: issued one chunk
// AHEAD of their use, so the strided 16-byte loads (one row per lane) overlap the previous chunk's work
__device__ __forceinline__ void tx_load_aux(uint4 (&ax)[4], int64_t row, int col0, const TxParams& p, int b1, int b2) {
  if (row >= p.M) return;
  const bf16* xp = p.aux + b2 * p.sx2 + b1 * p.sx1 + row * p.ldx + col0;
  if (p.aux32 && col0 + 32 <= p.Nst) {  // 32-byte aligned rows: two 256-bit loads (LDG.256) instead of four 128-bit ones
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(ax[0].x), "=r"(ax[0].y), "=r"(ax[0].z), "=r"(ax[0].w), "=r"(ax[1].x), "=r"(ax[1].y), "=r"(ax[1].z), "=r"(ax[1].w)
                 : "l"(xp));
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(ax[2].x), "=r"(ax[2].y), "=r"(ax[2].z), "=r"(ax[2].w), "=r"(ax[3].x), "=r"(ax[3].y), "=r"(ax[3].z), "=r"(ax[3].w)
                 : "l"(xp + 16));
    return;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g)
    if (col0 + 8 * g < p.Nst) ax[g] = *reinterpret_cast<const uint4*>(xp + 8 * g);
}

// epilogue arithmetic of one 32-column chunk of one accumulator row (lane == row); `ngroups` 8-column groups are live
__device__ __forceinline__ void tx_epi_math(const uint32_t (&r)[32], const uint4 (&ax)[4], float rv, const TxParams& p,
                                            int ngroups, float (&y)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(r[i]);
  const float a = p.alpha;
  if (p.epi == LCASR_EPI_SCALE) {
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] *= a;
  } else if (p.epi == LCASR_EPI_EXP2) {
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = ex2_approx(fmaf(y[i], a, -rv));
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < ngroups) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&ax[g]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = __bfloat1622float2(h[i]);
          float& v0 = y[8 * g + 2 * i];
          float& v1 = y[8 * g + 2 * i + 1];
          if (p.epi == LCASR_EPI_DS) { v0 = x.x * (v0 - rv) * a; v1 = x.y * (v1 - rv) * a; }
          else if (p.epi == LCASR_EPI_GELU_BWD) { v0 = a * v0 * gelu_tanh_grad(x.x); v1 = a * v1 * gelu_tanh_grad(x.y); }
          else { v0 = a * v0 * silu_grad(x.x); v1 = a * v1 * silu_grad(x.y); }  // LCASR_EPI_SILU_BWD
        }
      }
  }
}

// fp32 outputs ACCUMULATE with atomics (split-K partials, gradient accumulation into an existing buffer)
__device__ __forceinline__ void tx_store_chunk_f32(const uint32_t (&r)[32], const uint4 (&ax)[4], float rv, int64_t row, int col0,
                                                   const TxParams& p, int b1, int b2, float* __restrict__ out) {
  if (row >= p.M) return;
  const int ngroups = min(4, (p.Nst - col0) >> 3);  // 8 columns per group
  float y[32];
  tx_epi_math(r, ax, rv, p, ngroups, y);
  float* op = out + b2 * p.so2 + b1 * p.so1 + row * p.ldo + col0;
#pragma unroll
  for (int g = 0; g < 4; ++g)
    if (g < ngroups) {
      atomicAdd(reinterpret_cast<float4*>(op + 8 * g), make_float4(y[8 * g], y[8 * g + 1], y[8 * g + 2], y[8 * g + 3]));
      atomicAdd(reinterpret_cast<float4*>(op + 8 * g + 4), make_float4(y[8 * g + 4], y[8 * g + 5], y[8 * g + 6], y[8 * g + 7]));
    }
}

// 32 fp32 values of this lane's row -> bf16 -> 16-byte slots [4*hh, 4*hh+4) of its 128-byte row in the 128B-swizzled box
__device__ __forceinline__ void tx_stage_half(const float (&y)[32], int lane, uint32_t stage_addr, int hh) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 v;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h2[i] = __floats2bfloat162_rn(y[8 * g + 2 * i], y[8 * g + 2 * i + 1]);
    const uint32_t addr = stage_addr + lane * 128 + (((4 * hh + g) ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}

struct TxTile {
  int m_idx, n_idx, b1, b2, kb0, kb1;
};

template <int BN, int CG>
__device__ __forceinline__ TxTile tx_decode(int64_t tile, const TxParams& p, int tiles_n, int64_t tiles_m, int num_k) {
  TxTile t;
  const int ks = (int)(tile % p.ksplit);
  tile /= p.ksplit;
  t.n_idx = (int)(tile % tiles_n) * BN;
  tile /= tiles_n;
  t.m_idx = (int)(tile % tiles_m) * (CG * TX_BM);
  tile /= tiles_m;
  t.b1 = (int)(tile % p.nb1);
  t.b2 = (int)(tile / p.nb1);
  t.kb0 = ks * p.kper;
  t.kb1 = min(num_k, t.kb0 + p.kper);
  return t;
}

template <int BN, int A_MN, int B_MN, typename TOut, int CG>
__global__ void __launch_bounds__(TX_THREADS, 1)
gemm_tcx_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, TxParams p, TOut* out) {
  // tmC (bf16 outputs): 4-D store map of `out`, 64-column x 32-row boxes, 128B swizzle
  using Cfg = TxCfg<BN, CG>;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * Cfg::STAGES + 4];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (p.K + TX_BK - 1) / TX_BK;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int64_t tiles_m = (p.M + CG * TX_BM - 1) / (CG * TX_BM);
  const int64_t tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;  // per cluster
  const int64_t total_tiles = tiles_m * tiles_n * p.ksplit * p.nb1 * p.nb2;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), CG * TX_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) { tmem_alloc_cg2(smem_u32(&tmem_slot), Cfg::TMEM_COLS); tmem_relinquish_cg2(); }
    else { tmem_alloc(smem_u32(&tmem_slot), Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer ----------------
      int stage = 0; uint32_t phase = 0;
      const uint32_t full0 = CG == 2 ? mapa_shared(full_bar(0), 0) : full_bar(0);  // the leader's full barriers
      auto load = [&](uint32_t dst, const CUtensorMap* m, int st_, int c0, int c1, int c2, int c3) {
        if constexpr (CG == 2) tma_load_4d_cg2(dst, m, full0 + 8u * st_, c0, c1, c2, c3);
        else tma_load_4d(dst, m, full_bar(st_), c0, c1, c2, c3);
      };
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        const TxTile t = tx_decode<BN, CG>(tile, p, tiles_n, tiles_m, num_k);
        const int m_idx = t.m_idx + (int)rank * TX_BM, n_idx = t.n_idx + (int)rank * (BN / CG);  // this CTA's halves
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), CG * Cfg::STAGE_BYTES);  // both CTAs' bytes land here
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
          if constexpr (A_MN) {
#pragma unroll
            for (int i = 0; i < TX_BM / 64; ++i) load(sa + i * 8192, &tmA, stage, m_idx + i * 64, kb * TX_BK, t.b1, t.b2);
          } else {
            load(sa, &tmA, stage, kb * TX_BK, m_idx, t.b1, t.b2);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int i = 0; i < (BN / CG) / 64; ++i) load(sb + i * 8192, &tmB, stage, n_idx + i * 64, kb * TX_BK, t.b1, t.b2);
          } else {
            load(sb, &tmB, stage, kb * TX_BK, n_idx, t.b1, t.b2);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ---------------- MMA issuer (leader CTA only) ----------------
      constexpr uint32_t idesc = make_idesc_bf16_mn(CG * TX_BM, BN, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        const TxTile t = tx_decode<BN, CG>(tile, p, tiles_n, tiles_m, num_k);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = A_MN ? make_smem_desc_mnmajor(sa) : make_smem_desc_kmajor(sa, 1024, kLayoutSW128);
          const uint64_t bdesc = B_MN ? make_smem_desc_mnmajor(sb) : make_smem_desc_kmajor(sb, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < TX_BK / 16; ++k) {  // K-major: +32 B per K=16 step; MN-major: +16 rows * 128 B
            if constexpr (CG == 2)
              umma_f16_ss_cg2(d_tmem, adesc + (A_MN ? 128 : 2) * k, bdesc + (B_MN ? 128 : 2) * k, idesc, (kb != t.kb0 || k != 0) ? 1u : 0u);
            else
              umma_f16_ss(d_tmem, adesc + (A_MN ? 128 : 2) * k, bdesc + (B_MN ? 128 : 2) * k, idesc, (kb != t.kb0 || k != 0) ? 1u : 0u);
          }
          if constexpr (CG == 2) umma_commit_cg2(empty_bar(stage), 3);
          else umma_commit(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 2) umma_commit_cg2(tfull_bar(acc), 3);
        else umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {  // ---------------- epilogue warps 2..5 ----------------
    const int lane_base = (warp & 3) * 32;
    const int half = (warp - 2) >> 2;
    constexpr int CPW = (BN / 32) / (TX_EPI_WARPS / 4);  // 32-column chunks per warp
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t tempty0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : 0;  // the leader's tempty barriers
    for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
      const TxTile t = tx_decode<BN, CG>(tile, p, tiles_n, tiles_m, num_k);
      const int m_idx = t.m_idx + (int)rank * TX_BM;  // this CTA's 128 accumulator rows
      const int64_t row = (int64_t)m_idx + lane_base + lane;
      const bool has_aux = p.epi >= LCASR_EPI_DS;
      // operands of the epilogue are requested BEFORE the accumulator wait: their latency hides behind the main loop
      float rv = 0.f;
      if ((p.epi == LCASR_EPI_EXP2 || p.epi == LCASR_EPI_DS) && row < p.M) rv = p.rowvec[t.b2 * p.sr2 + t.b1 * p.sr1 + row];
      const int c_lo = half * CPW, c_hi = (half + 1) * CPW;
      if constexpr (sizeof(TOut) == 4) {
        uint4 ax_nxt[4] = {};
        if (has_aux && t.n_idx + c_lo * 32 < p.N) tx_load_aux(ax_nxt, row, t.n_idx + c_lo * 32, p, t.b1, t.b2);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)lane_base << 16) + acc * BN;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          if (t.n_idx + c * 32 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_addr + c * 32, r);
          uint4 ax_cur[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) ax_cur[g] = ax_nxt[g];
          if (has_aux && c + 1 < c_hi && t.n_idx + (c + 1) * 32 < p.N) tx_load_aux(ax_nxt, row, t.n_idx + (c + 1) * 32, p, t.b1, t.b2);
          tmem_wait_ld();
          tx_store_chunk_f32(r, ax_cur, rv, row, t.n_idx + c * 32, p, t.b1, t.b2, reinterpret_cast<float*>(out));
        }
      } else {  // bf16: 64 columns at a time through a swizzled shared-memory box and one TMA store per box
        uint4 ax0[4] = {}, ax1[4] = {};
        if (has_aux && t.n_idx + c_lo * 32 < p.N) {
          tx_load_aux(ax0, row, t.n_idx + c_lo * 32, p, t.b1, t.b2);
          tx_load_aux(ax1, row, t.n_idx + (c_lo + 1) * 32, p, t.b1, t.b2);
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)lane_base << 16) + acc * BN;
        const uint32_t stg = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES + (warp - 2) * 4096;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 2) {
          const int col0 = t.n_idx + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(t_addr + c * 32, r0);
          tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, r1);
          uint4 a0[4], a1[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) { a0[g] = ax0[g]; a1[g] = ax1[g]; }
          if (has_aux && c + 2 < c_hi && col0 + 64 < p.N) {  // next pair's aux, in flight during this pair's work
            tx_load_aux(ax0, row, col0 + 64, p, t.b1, t.b2);
            tx_load_aux(ax1, row, col0 + 96, p, t.b1, t.b2);
          }
          tmem_wait_ld();
          float y[32];
          if (lane == 0) tma_store_wait_read();  // the previous box has left the staging tile
          __syncwarp();
          tx_epi_math(r0, a0, rv, p, 4, y);
          tx_stage_half(y, lane, stg, 0);
          tx_epi_math(r1, a1, rv, p, 4, y);
          tx_stage_half(y, lane, stg, 1);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmC, stg, col0, m_idx + lane_base, t.b1, t.b2);  // rows >= M / columns >= N are clipped
            tma_store_commit();
          }
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
    if (sizeof(TOut) == 2 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---- attention backward, first half: P and dS of one (recording, head) in ONE pass --------------------------
//   S  = Q K^T,  dP = dO V^T          (two accumulators per tile, K = Dh)
//   P  = exp2(S * scale*log2e - lse2[row])                         -> bf16 [.., N, Np]
//   dS = P o (dP - D[row]) * scale                                 -> bf16 [.., N, Np]
// Same pipeline as gemm_tcx_kernel with 128x128 tiles; a stage holds the Q, K, dO and V tiles (64 KB), TMEM holds
// 2 x (S | dP) = 512 columns.  With K = Dh <= 128 the tile is epilogue-bound: computing both products per tile
// halves the number of epilogue passes and removes the re-read of P that a separate dS GEMM needs.
struct PdsParams {
  int64_t N;   // tokens (rows and keys)
  int Dh, Nst, H, nbatch;
  int64_t ldo, so1, so2;      // P / dS pitch and (head, recording) strides
  const float* lse;           // [nbatch, H, N] log2-domain
  const float* dvec;          // [nbatch, H, N]
  float alpha_p, scale;
  bf16* P;
  bf16* dS;
};
constexpr int PDS_STAGES = 3, PDS_STAGE_BYTES = 4 * 16384;
constexpr int PDS_SMEM = PDS_STAGES * PDS_STAGE_BYTES + TX_EPI_WARPS * 4096 + 1024;  // + one staging box per epilogue warp

__global__ void __launch_bounds__(TX_THREADS, 1)
attn_bwd_pds_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmDS, PdsParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * PDS_STAGES + 4];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PDS_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PDS_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PDS_STAGES + 2 + a); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (p.Dh + TX_BK - 1) / TX_BK;
  const int tiles_1d = (int)((p.N + 127) / 128);
  const int64_t total_tiles = (int64_t)tiles_1d * tiles_1d * p.H * p.nbatch;
  auto decode = [&](int64_t tile, int& m_idx, int& n_idx, int& h, int& b) {
    n_idx = (int)(tile % tiles_1d) * 128; tile /= tiles_1d;
    m_idx = (int)(tile % tiles_1d) * 128; tile /= tiles_1d;
    h = (int)(tile % p.H); b = (int)(tile / p.H);
  };
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmDO); prefetch_tensormap(&tmV);
    for (int s = 0; s < PDS_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TX_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_idx, n_idx, h, b;
        decode(tile, m_idx, n_idx, h, b);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_arrive_expect_tx(full_bar(stage), PDS_STAGE_BYTES);
          const uint32_t s0 = smem_base + stage * PDS_STAGE_BYTES;
          tma_load_4d(s0, &tmQ, full_bar(stage), kb * TX_BK, m_idx, h, b);
          tma_load_4d(s0 + 16384, &tmK, full_bar(stage), kb * TX_BK, n_idx, h, b);
          tma_load_4d(s0 + 32768, &tmDO, full_bar(stage), kb * TX_BK, m_idx, h, b);
          tma_load_4d(s0 + 49152, &tmV, full_bar(stage), kb * TX_BK, n_idx, h, b);
          if (++stage == PDS_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_mn(128, 128, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_s = tmem_base + acc * 256, d_dp = d_s + 128;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t s0 = smem_base + stage * PDS_STAGE_BYTES;
          const uint64_t qd = make_smem_desc_kmajor(s0, 1024, kLayoutSW128), kd = make_smem_desc_kmajor(s0 + 16384, 1024, kLayoutSW128);
          const uint64_t od = make_smem_desc_kmajor(s0 + 32768, 1024, kLayoutSW128), vd = make_smem_desc_kmajor(s0 + 49152, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < TX_BK / 16; ++k) umma_f16_ss(d_s, qd + 2 * k, kd + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
          for (int k = 0; k < TX_BK / 16; ++k) umma_f16_ss(d_dp, od + 2 * k, vd + 2 * k, idesc, (kb | k) != 0);
          umma_commit(empty_bar(stage));
          if (++stage == PDS_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int lane_base = (warp & 3) * 32;
    const int half = (warp - 2) >> 2;
    constexpr int CPW = 4 / (TX_EPI_WARPS / 4);
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_idx, n_idx, h, b;
      decode(tile, m_idx, n_idx, h, b);
      const int64_t row = (int64_t)m_idx + lane_base + lane;
      const bool row_ok = row < p.N;
      float lse = 0.f, dv = 0.f;
      if (row_ok) {
        lse = p.lse[((int64_t)b * p.H + h) * p.N + row];
        dv = p.dvec[((int64_t)b * p.H + h) * p.N + row];
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)lane_base << 16) + acc * 256;
      const uint32_t stg = smem_base + PDS_STAGES * PDS_STAGE_BYTES + (warp - 2) * 4096;
      {  // this warp's 64 columns [n_idx + 64*half, +64): P box, then dS box, one TMA store each
        const int c = half * CPW;
        const int col0 = n_idx + c * 32;
        if (col0 < p.N) {
          uint32_t rs0[32], rs1[32], rd0[32], rd1[32];
          tmem_ld_32x32b_x32(t_addr + c * 32, rs0);
          tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, rs1);
          tmem_ld_32x32b_x32(t_addr + 128 + c * 32, rd0);
          tmem_ld_32x32b_x32(t_addr + 128 + (c + 1) * 32, rd1);
          tmem_wait_ld();
          float pv[32];
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int i = 0; i < 32; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(hh ? rs1[i] : rs0[i]), p.alpha_p, -lse));
            tx_stage_half(pv, lane, stg, hh);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { tma_store_4d(&tmP, stg, col0, m_idx + lane_base, h, b); tma_store_commit(); tma_store_wait_read(); }
          __syncwarp();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              // dS uses the bf16-rounded P, the operand the dV product consumes
              const float pr = ex2_approx(fmaf(__uint_as_float(hh ? rs1[i] : rs0[i]), p.alpha_p, -lse));
              const float prr = __bfloat162float(__float2bfloat16_rn(pr));
              pv[i] = prr * (__uint_as_float(hh ? rd1[i] : rd0[i]) - dv) * p.scale;
            }
            tx_stage_half(pv, lane, stg, hh);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { tma_store_4d(&tmDS, stg, col0, m_idx + lane_base, h, b); tma_store_commit(); }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---- attention backward, flash style: no N x N tensor ever reaches memory -------------------------------------
// Two launches of one kernel template (7 tile products instead of the 5 of the materialised form, but no atomics and
// O(N) memory):
//   MODE 0 (dK, dV): a CTA owns 128 KEYS (K_j, V_j resident in shared memory) and streams 64-query chunks (Q_i, dO_i):
//       S^T = K_j Q_i^T, dP^T = V_j dO_i^T            (SS products, 128 x 64 fp32 accumulators, double-buffered)
//       P^T = exp2(S^T c - lse_i), dS^T = P^T o (dP^T - D_i) scale   -> bf16, written over the accumulators they came from
//       dV_j += P^T dO_i, dK_j += dS^T Q_i            (TS products: A from tensor memory, B = the same chunk tiles read
//                                                      as MN-major operands — no transposed copies)
//   MODE 1 (dQ):     a CTA owns 128 QUERIES and streams 64-key chunks (K_j, V_j):
//       S = Q_i K_j^T, dP = dO_i V_j^T, dS as above (lane = query: lse / D are per-thread scalars), dQ_i += dS K_j.
//       Q_i and dO_i live in TENSOR memory (copied there once by the softmax warps), so every product of this mode is a TS
//       product: an SS product with a 64-wide N re-reads the 128-row A tile from shared memory for every K = 16 step
//       (6 KB per 32 tensor cycles: shared-memory-bound, MODE 0 pays that for S^T / dP^T).
// The transposed formulation of MODE 0 is what lets P^T / dS^T be TS operands: tensor memory lanes must be the M
// dimension (keys) of the dV / dK products.  TMEM: accumulators [0, 2 Dh) (MODE 1: [0, Dh)), then 2 x 64 columns of
// S and 2 x 64 of dP (MODE 1: then Q_i and dO_i, Dh / 2 columns each).  Per-row statistics come packed as float2
// {lse2, D * scale} per (recording, head, token), padded to a multiple of 128 tokens with {+inf, 0} (P = dS = 0 for tokens
// that do not exist) by attn_bwd_prep_kernel.
// Warps 0-15: softmax (four per TMEM lane quarter, 16 columns each — the first version had 8 and its tensor pipe idled half
// the time: ncu showed 270 instructions per warp and chunk taking ~1300 cycles, pure latency), warp 16: TMA producer,
// warp 17: MMA issuer.
constexpr int FB_SOFT_WARPS = 16, FB_THREADS = 32 * FB_SOFT_WARPS + 64, FB_STAGES = 3, FB_CH = 64;
template <int DH, int MODE> struct FbCfg {
  static constexpr int SUB = DH / 64;
  static constexpr int X_BYTES = MODE == 0 ? SUB * 128 * 128 : 0;  // one resident tile: [128 rows][DH] as SUB swizzled boxes
  static constexpr int Y_BYTES = SUB * FB_CH * 128;                // one streamed tile: [64 rows][DH]
  static constexpr int STAGE_BYTES = 2 * Y_BYTES;
  static constexpr int VEC_BYTES = FB_CH * 8;
  static constexpr int SMEM = 2 * X_BYTES + FB_STAGES * STAGE_BYTES + FB_STAGES * VEC_BYTES + 1024;
};

__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct FbParams {
  int64_t N, Npad;        // valid tokens; padded length of the statistics rows
  int64_t tok_stride;     // elements between consecutive tokens of a [.., N, H, Dh] tensor (H * Dh)
  int64_t rec_stride;     // elements between recordings
  int H;
  float c_log2, scale;    // scale * log2(e), scale
  const float2* stats;    // [nb, H, Npad]
  const bf16* x1;         // MODE 1: q and dO (copied to tensor memory by the softmax warps)
  const bf16* x2;
  bf16* out1;             // MODE 0: dV, MODE 1: dQ
  bf16* out2;             // MODE 0: dK
};

template <int DH, int MODE>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_flash_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                      const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2, FbParams p) {
  using Cfg = FbCfg<DH, MODE>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * FB_STAGES + 2 + 2 + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x1_smem = smem_base, x2_smem = smem_base + Cfg::X_BYTES;
  auto y1_smem = [&](int s) { return smem_base + 2 * Cfg::X_BYTES + s * Cfg::STAGE_BYTES; };
  auto y2_smem = [&](int s) { return y1_smem(s) + Cfg::Y_BYTES; };
  const uint32_t vec0 = smem_base + 2 * Cfg::X_BYTES + FB_STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t x_full = bar0;
  auto full_bar = [&](int s) { return bar0 + 8u * (1 + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (1 + FB_STAGES + s); };
  auto s_full = [&](int u) { return bar0 + 8u * (1 + 2 * FB_STAGES + u); };      // S, dP of a chunk retired
  auto p_full = [&](int u) { return bar0 + 8u * (1 + 2 * FB_STAGES + 2 + u); };  // P / dS written (all softmax warps)
  const uint32_t done_bar = bar0 + 8u * (1 + 2 * FB_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int x_row0 = blockIdx.x * 128;
  const int n_chunks = (int)((p.N + FB_CH - 1) / FB_CH);
  constexpr int ACC_COLS = MODE == 0 ? 2 * DH : DH;
  constexpr int S_COL = ACC_COLS, DP_COL = ACC_COLS + 128;
  constexpr int X1_COL = ACC_COLS + 256, X2_COL = X1_COL + DH / 2;  // MODE 1: bf16 Q_i, dO_i
  static_assert(MODE == 0 ? ACC_COLS + 256 <= 512 : X2_COL + DH / 2 <= 512, "tensor memory budget");

  if (warp == FB_SOFT_WARPS && lane == 0) {
    if (MODE == 0) { prefetch_tensormap(&tmX1); prefetch_tensormap(&tmX2); }
    prefetch_tensormap(&tmY1); prefetch_tensormap(&tmY2);
    mbar_init(x_full, MODE == 0 ? 1 : FB_SOFT_WARPS);
    for (int s = 0; s < FB_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int u = 0; u < 2; ++u) { mbar_init(s_full(u), 1); mbar_init(p_full(u), FB_SOFT_WARPS); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == FB_SOFT_WARPS + 1) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  const float2* stats_row = p.stats + ((int64_t)b * p.H + h) * p.Npad;

  if (warp == FB_SOFT_WARPS) {
    if (lane == 0) {  // ------------------------- TMA producer -------------------------
      if (MODE == 0) {
        mbar_arrive_expect_tx(x_full, 2 * Cfg::X_BYTES);
#pragma unroll
        for (int i = 0; i < Cfg::SUB; ++i) {
          tma_load_4d(x1_smem + i * 16384, &tmX1, x_full, i * 64, x_row0, h, b);
          tma_load_4d(x2_smem + i * 16384, &tmX2, x_full, i * 64, x_row0, h, b);
        }
      }
      int stage = 0; uint32_t phase = 0;
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES + (MODE == 0 ? Cfg::VEC_BYTES : 0));
#pragma unroll
        for (int i = 0; i < Cfg::SUB; ++i) {
          tma_load_4d(y1_smem(stage) + i * 8192, &tmY1, full_bar(stage), i * 64, c * FB_CH, h, b);
          tma_load_4d(y2_smem(stage) + i * 8192, &tmY2, full_bar(stage), i * 64, c * FB_CH, h, b);
        }
        if (MODE == 0) bulk_load_1d(vec0 + stage * Cfg::VEC_BYTES, stats_row + (int64_t)c * FB_CH, Cfg::VEC_BYTES, full_bar(stage));
        if (++stage == FB_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == FB_SOFT_WARPS + 1) {
    // ------------------------- MMA issuer (converged warp, one elected lane issues) -------------------------
    constexpr uint32_t idesc_s = make_idesc_bf16_mn(128, FB_CH, 0, 0);
    constexpr uint32_t idesc_acc = make_idesc_bf16_mn(128, DH, 0, 1);  // A from tensor memory, B MN-major
    auto issue_s = [&](int stage, int u) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const int sub = kk / 4, within = kk % 4;
        const uint64_t bd = make_smem_desc_kmajor(y1_smem(stage) + sub * 8192, 1024, kLayoutSW128) + 2 * within;
        if (MODE == 0)
          umma_f16_ss(tmem_base + S_COL + u * FB_CH, make_smem_desc_kmajor(x1_smem + sub * 16384, 1024, kLayoutSW128) + 2 * within, bd,
                      idesc_s, kk != 0);
        else umma_f16_ts(tmem_base + S_COL + u * FB_CH, tmem_base + X1_COL + kk * 8, bd, idesc_s, kk != 0);
      }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const int sub = kk / 4, within = kk % 4;
        const uint64_t bd = make_smem_desc_kmajor(y2_smem(stage) + sub * 8192, 1024, kLayoutSW128) + 2 * within;
        if (MODE == 0)
          umma_f16_ss(tmem_base + DP_COL + u * FB_CH, make_smem_desc_kmajor(x2_smem + sub * 16384, 1024, kLayoutSW128) + 2 * within, bd,
                      idesc_s, kk != 0);
        else umma_f16_ts(tmem_base + DP_COL + u * FB_CH, tmem_base + X2_COL + kk * 8, bd, idesc_s, kk != 0);
      }
    };
    // bf16 operand columns of a chunk: the softmax warp of column quarter cq writes its 16 values as 8 packed columns at the
    // START of its own fp32 range (never over columns another warp may still have to read): k-step kk -> column kk * 16
    auto issue_acc = [&](int stage, int u, bool accumulate) {
      if (MODE == 0) {
#pragma unroll
        for (int kk = 0; kk < FB_CH / 16; ++kk)  // dV += P^T dO
          umma_f16_ts(tmem_base, tmem_base + S_COL + u * FB_CH + kk * 16, make_smem_desc_mnmajor(y2_smem(stage)) + kk * (2048 >> 4),
                      idesc_acc, (accumulate || kk != 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < FB_CH / 16; ++kk)  // dK += dS^T Q
          umma_f16_ts(tmem_base + DH, tmem_base + DP_COL + u * FB_CH + kk * 16, make_smem_desc_mnmajor(y1_smem(stage)) + kk * (2048 >> 4),
                      idesc_acc, (accumulate || kk != 0) ? 1u : 0u);
      } else {
#pragma unroll
        for (int kk = 0; kk < FB_CH / 16; ++kk)  // dQ += dS K
          umma_f16_ts(tmem_base, tmem_base + DP_COL + u * FB_CH + kk * 16, make_smem_desc_mnmajor(y1_smem(stage)) + kk * (2048 >> 4),
                      idesc_acc, (accumulate || kk != 0) ? 1u : 0u);
      }
    };
    mbar_wait(x_full, 0);
    for (int c = 0; c < 2 && c < n_chunks; ++c) {
      mbar_wait(full_bar(c), 0);
      tc_fence_after();
      if (elect_one()) { issue_s(c, c); umma_commit(s_full(c)); }
      __syncwarp();
    }
    for (int c = 0; c < n_chunks; ++c) {
      const int u = c & 1, stage = c % FB_STAGES;
      mbar_wait(p_full(u), (c >> 1) & 1);
      tc_fence_after();
      if (elect_one()) { issue_acc(stage, u, c > 0); umma_commit(empty_bar(stage)); }
      __syncwarp();
      if (c + 2 < n_chunks) {
        // S / dP of chunk c+2 reuse buffer u: the in-order tensor pipe runs them after the products that read P / dS(c)
        const int st2 = (c + 2) % FB_STAGES;
        mbar_wait(full_bar(st2), ((c + 2) / FB_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) { issue_s(st2, u); umma_commit(s_full(u)); }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // ------------------------- softmax warps -------------------------
    const int lq = warp & 3, cq = warp >> 2;  // TMEM lane quarter (== warp_id % 4), column quarter of a chunk
    const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
    const int64_t row = (int64_t)x_row0 + lq * 32 + lane;
    float lse_r = 0.f, dsc_r = 0.f;  // MODE 1: this lane's query row
    if (MODE == 1) {
      const float2 st = stats_row[row];  // padded to a multiple of 128 rows
      lse_r = st.x; dsc_r = st.y;
      // Q_i and dO_i of this lane's row -> tensor memory (bf16 pairs): this warp's quarter of the head dim
      constexpr int CW = DH / 4;  // elements per warp
      const int64_t off = (int64_t)b * p.rec_stride + row * p.tok_stride + (int64_t)h * DH + cq * CW;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const bf16* src = (t == 0 ? p.x1 : p.x2) + off;
        uint32_t w[CW / 2];
#pragma unroll
        for (int g = 0; g < CW / 8; ++g) {
          uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
          if (row < p.N) v4 = *reinterpret_cast<const uint4*>(src + 8 * g);
          w[4 * g] = v4.x; w[4 * g + 1] = v4.y; w[4 * g + 2] = v4.z; w[4 * g + 3] = v4.w;
        }
        const uint32_t dst = t_lane + (t == 0 ? X1_COL : X2_COL) + cq * (CW / 2);
        if constexpr (CW / 2 == 16) tmem_st_32x32b_x16(dst, *reinterpret_cast<uint32_t(*)[16]>(&w[0]));
        else tmem_st_32x32b_x8(dst, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_full);
    }
    for (int c = 0; c < n_chunks; ++c) {
      const int u = c & 1, stage = c % FB_STAGES;
      if (MODE == 0) mbar_wait(full_bar(stage), (c / FB_STAGES) & 1);  // the statistics of this chunk (already complete)
      mbar_wait(s_full(u), (c >> 1) & 1);
      tc_fence_after();
      uint32_t s[16], dp[16];
      tmem_ld_32x32b_x16(t_lane + S_COL + u * FB_CH + cq * 16, s);
      tmem_ld_32x32b_x16(t_lane + DP_COL + u * FB_CH + cq * 16, dp);
      tmem_wait_ld();
      uint32_t pk[8], dk[8];
      const float2* vec = reinterpret_cast<const float2*>(smem_raw + (vec0 - smem_u32(smem_raw)) + stage * Cfg::VEC_BYTES) + cq * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float l0 = lse_r, l1 = lse_r, d0 = dsc_r, d1 = dsc_r;
        if (MODE == 0) {
          const float4 v2 = *reinterpret_cast<const float4*>(vec + 2 * i);  // two queries' {lse, D scale}: broadcast read
          l0 = v2.x; d0 = v2.y; l1 = v2.z; d1 = v2.w;
        }
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), p.c_log2, -l0));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), p.c_log2, -l1));
        const float g0 = p0 * fmaf(__uint_as_float(dp[2 * i]), p.scale, -d0);
        const float g1 = p1 * fmaf(__uint_as_float(dp[2 * i + 1]), p.scale, -d1);
        __nv_bfloat162 pp = __floats2bfloat162_rn(p0, p1), gg = __floats2bfloat162_rn(g0, g1);
        pk[i] = *reinterpret_cast<uint32_t*>(&pp);
        dk[i] = *reinterpret_cast<uint32_t*>(&gg);
      }
      if (MODE == 0) tmem_st_32x32b_x8(t_lane + S_COL + u * FB_CH + cq * 16, pk);
      tmem_st_32x32b_x8(t_lane + DP_COL + u * FB_CH + cq * 16, dk);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(u));
    }
    // ---- epilogue: this warp's lane quarter, column quarter cq of every accumulator ----
    mbar_wait(done_bar, 0);
    tc_fence_after();
#pragma unroll
    for (int a = 0; a < (MODE == 0 ? 2 : 1); ++a) {
      bf16* outp = (a == 0 ? p.out1 : p.out2) + (int64_t)b * p.rec_stride + row * p.tok_stride + (int64_t)h * DH;
#pragma unroll
      for (int cc = 0; cc < DH / 64; ++cc) {
        const int col = cq * (DH / 4) + cc * 16;
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_lane + a * DH + col, r);
        tmem_wait_ld();
        if (row < p.N) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(r[8 * g + i]);
            Vec8<bf16>::store(outp + col + 8 * g, y);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FB_SOFT_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D = rowsum(dO o O) and the packed per-row statistics {lse2, D * scale} of the flash backward
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_out, const float* __restrict__ lse2,
                                     int nb, int64_t N, int64_t Npad, int64_t n_pitch, int64_t lse_pitch, int H, int Dh,
                                     float scale, float2* __restrict__ stats) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (b, n in [0, Npad), h), h fastest
  if (idx >= (int64_t)nb * Npad * H) return;
  const int h = (int)(idx % H);
  const int64_t n = (idx / H) % Npad, b = idx / ((int64_t)H * Npad);
  float2 st = make_float2(INFINITY, 0.f);
  if (n < N) {
    const int64_t off = ((b * n_pitch + n) * H + h) * Dh;
    float acc = 0.f;
    for (int i = 0; i < Dh; i += 8) {
      float a[8], g[8];
      Vec8<bf16>::load(o + off + i, a);
      Vec8<bf16>::load(d_out + off + i, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(a[j], g[j], acc);
    }
    st = make_float2(lse2[(b * H + h) * lse_pitch + n], acc * scale);
  }
  stats[(b * H + h) * Npad + n] = st;
}

// ---- host side ----------------------------------------------------------------------------------

typedef CUresult (*PFN_tmapEncodeTiledX)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiledX get_encode_fn_x() {
  static PFN_tmapEncodeTiledX fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiledX)ptr;
  }
  return fn;
}

// 4-D bf16 tensor map over a stored row-major matrix [rows, cols] (pitch elements) with two batch dims.
static int make_tmap_4d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, int64_t pitch, int nb1, int64_t s1,
                        int nb2, int64_t s2, uint32_t box_rows, uint32_t box_cols) {
  PFN_tmapEncodeTiledX fn = get_encode_fn_x();
  if (!fn) return set_error(LCASR_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // strides of size-1 dims are never used for addressing but must still be valid (multiple of 16 B, > 0)
  const uint64_t st1 = nb1 > 1 ? (uint64_t)s1 * 2 : (uint64_t)pitch * 2;
  const uint64_t st2 = nb2 > 1 ? (uint64_t)s2 * 2 : (uint64_t)pitch * 2;
  cuuint64_t dims[4] = {cols, rows, (cuuint64_t)nb1, (cuuint64_t)nb2};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, st1, st2};
  cuuint32_t box[4] = {box_cols, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LCASR_E_CUDA,
                     "cuTensorMapEncodeTiled(4d) failed (%d): base=%p rows=%llu cols=%llu pitch=%lld nb=%d,%d s=%lld,%lld box=%ux%u",
                     (int)r, base, (unsigned long long)rows, (unsigned long long)cols, (long long)pitch, nb1, nb2,
                     (long long)s1, (long long)s2, box_rows, box_cols);
  return 0;
}

template <int BN, int A_MN, int B_MN, typename TOut, int CG>
static int launch_tcx(const lcasr_gemm_ex_args& g, const TxParams& p, cudaStream_t st) {
  using Cfg = TxCfg<BN, CG>;
  CUtensorMap tmA, tmB;
  // stored matrices: K-major operand = [M or N rows, K cols]; MN-major operand = [K rows, M or N cols]
  if (A_MN) LCASR_TRY(make_tmap_4d(&tmA, g.A, (uint64_t)g.K, (uint64_t)g.M, g.lda, g.nb1, g.sa1, g.nb2, g.sa2, TX_BK, 64));
  else LCASR_TRY(make_tmap_4d(&tmA, g.A, (uint64_t)g.M, (uint64_t)g.K, g.lda, g.nb1, g.sa1, g.nb2, g.sa2, TX_BM, TX_BK));
  if (B_MN) LCASR_TRY(make_tmap_4d(&tmB, g.B, (uint64_t)g.K, (uint64_t)g.N, g.ldb, g.nb1, g.sb1, g.nb2, g.sb2, TX_BK, 64));
  else LCASR_TRY(make_tmap_4d(&tmB, g.B, (uint64_t)g.N, (uint64_t)g.K, g.ldb, g.nb1, g.sb1, g.nb2, g.sb2, BN / CG, TX_BK));
  CUtensorMap tmC = tmA;  // placeholder for fp32 outputs (never dereferenced)
  if (sizeof(TOut) == 2)
    LCASR_TRY(make_tmap_4d(&tmC, g.out, (uint64_t)g.M, (uint64_t)g.N, g.ldo, g.nb1, g.so1, g.nb2, g.so2, 32, 64));
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(gemm_tcx_kernel<BN, A_MN, B_MN, TOut, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES));
    attr_set.mark(attr_dev);
  }
  const int64_t tiles = ceil_div(g.M, CG * TX_BM) * ceil_div(g.N, BN) * p.ksplit * g.nb1 * g.nb2;
  const int64_t units = kNumSMs / CG;
  const int grid = (int)(tiles < units ? tiles : units) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TX_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  LCASR_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcx_kernel<BN, A_MN, B_MN, TOut, CG>, tmA, tmB, tmC, p, (TOut*)g.out));
  LCASR_LAUNCH_CHECK();
  return 0;
}

template <int BN, typename TOut, int CG>
static int dispatch_major(const lcasr_gemm_ex_args& g, const TxParams& p, cudaStream_t st) {
  if (g.a_mn) return g.b_mn ? launch_tcx<BN, 1, 1, TOut, CG>(g, p, st) : launch_tcx<BN, 1, 0, TOut, CG>(g, p, st);
  return g.b_mn ? launch_tcx<BN, 0, 1, TOut, CG>(g, p, st) : launch_tcx<BN, 0, 0, TOut, CG>(g, p, st);
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_gemm_ex(const lcasr_gemm_ex_args* gp, void* stream) {
  LCASR_CHECK_ARG(gp, "gemm_ex: NULL args");
  const lcasr_gemm_ex_args& g = *gp;
  LCASR_CHECK_ARG(g.A && g.B && g.out, "gemm_ex: NULL operand");
  LCASR_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0 && g.nb1 > 0 && g.nb2 > 0, "gemm_ex: bad shape");
  LCASR_CHECK_ARG(g.M < ((int64_t)1 << 31), "gemm_ex: M too large");
  // bf16 outputs may have any N when the row pitch leaves room for the 8-column store granularity: the
  // padding columns then receive unspecified values (consumers bound their K / M extent by the true size)
  const int Nst = (g.N + 7) / 8 * 8;
  LCASR_CHECK_ARG(g.N % 8 == 0 || (g.out_dtype == LCASR_BF16 && g.ldo >= Nst && (!g.aux || g.ldaux >= Nst)),
                  "gemm_ex: N=%d must be a multiple of 8 (or a bf16 output with pitch >= N rounded up to 8)", g.N);
  LCASR_CHECK_ARG(g.lda % 8 == 0 && g.ldb % 8 == 0 && g.sa1 % 8 == 0 && g.sa2 % 8 == 0 && g.sb1 % 8 == 0 && g.sb2 % 8 == 0,
                  "gemm_ex: operand pitches / batch strides must be multiples of 8 elements (TMA: 16 bytes)");
  LCASR_CHECK_ARG(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.B & 15) == 0 && ((uintptr_t)g.out & 15) == 0 &&
                      ((uintptr_t)g.aux & 15) == 0,
                  "gemm_ex: A, B, out, aux must be 16-byte aligned");
  LCASR_CHECK_ARG(g.out_dtype == LCASR_BF16 || g.out_dtype == LCASR_F32, "gemm_ex: bad out dtype");
  LCASR_CHECK_ARG(g.ldo % (g.out_dtype == LCASR_BF16 ? 8 : 4) == 0 && g.so1 % 8 == 0 && g.so2 % 8 == 0,
                  "gemm_ex: output pitch / strides must keep 16-byte alignment");
  LCASR_CHECK_ARG(g.epi >= LCASR_EPI_SCALE && g.epi <= LCASR_EPI_SILU_BWD, "gemm_ex: bad epilogue %d", g.epi);
  const bool needs_aux = g.epi == LCASR_EPI_DS || g.epi == LCASR_EPI_GELU_BWD || g.epi == LCASR_EPI_SILU_BWD;
  const bool needs_rv = g.epi == LCASR_EPI_EXP2 || g.epi == LCASR_EPI_DS;
  LCASR_CHECK_ARG(!needs_aux || (g.aux && g.ldaux % 8 == 0 && g.sx1 % 8 == 0 && g.sx2 % 8 == 0), "gemm_ex: epilogue %d needs aux", g.epi);
  LCASR_CHECK_ARG(!needs_rv || g.rowvec, "gemm_ex: epilogue %d needs rowvec", g.epi);
  LCASR_CHECK_ARG(g.out_dtype == LCASR_F32 || g.ksplit <= 1, "gemm_ex: split-K needs an fp32 (accumulating) output");
  LCASR_CHECK_ARG(g.out_dtype == LCASR_BF16 || g.epi == LCASR_EPI_SCALE, "gemm_ex: fp32 outputs accumulate alpha*acc only");
  const int num_k = (int)ceil_div(g.K, TX_BK);
  const bool wide = (g.N % 256 == 0) || g.N > 512;
  const int BN = wide ? 256 : 128;
  // CTA pairs (256-row tiles) when the un-split problem has at least one 256-row tile per pair-slot worth of work;
  // LCASR_GEMM_CG=1|2 forces a choice (A/B runs, tests)
  static const int force_cg = getenv("LCASR_GEMM_CG") ? atoi(getenv("LCASR_GEMM_CG")) : 0;
  const int64_t tiles2 = ceil_div(g.M, 2 * TX_BM) * ceil_div(g.N, BN) * g.nb1 * g.nb2;
  const bool pair = force_cg ? force_cg == 2 : (g.M >= 2 * TX_BM && tiles2 * num_k >= (int64_t)(kNumSMs / 2) * 8);
  const int CGr = pair ? 2 : 1;
  int ksplit = g.ksplit;
  if (ksplit <= 0) {  // auto (fp32 accumulating outputs: weight gradients have K = all tokens and few output tiles)
    ksplit = 1;
    if (g.out_dtype == LCASR_F32) {
      // pick the split whose tile count fills whole waves of the persistent CTAs (CTA pairs): efficiency =
      // waves / ceil(waves); a larger split must be 5 % better than a smaller one (fewer atomic passes over the output).
      // e.g. 72 tiles on 148 CTAs: k = 2 -> 144 tiles = 0.97 wave instead of k = 5 -> 2.43 waves -> 3 rounds (0.81).
      const int64_t tiles = ceil_div(g.M, CGr * TX_BM) * ceil_div(g.N, BN) * g.nb1 * g.nb2;
      const int64_t units = kNumSMs / CGr;
      double best = 0.0;
      for (int k = 1; k <= 128 && k * 4 <= num_k; ++k) {
        const double waves = (double)(tiles * k) / units;
        const double eff = waves / (double)ceil_div(tiles * k, units);
        if (eff > best * 1.05) { best = eff; ksplit = k; }
      }
    }
  }
  if (ksplit > num_k) ksplit = num_k;
  const int kper = (int)ceil_div(num_k, ksplit);
  ksplit = (int)ceil_div(num_k, kper);  // no empty splits
  TxParams p;
  p.M = g.M; p.N = g.N; p.Nst = Nst; p.K = g.K; p.nb1 = g.nb1; p.nb2 = g.nb2; p.ksplit = ksplit; p.kper = kper;
  p.ldo = g.ldo; p.so1 = g.so1; p.so2 = g.so2;
  p.aux = (const bf16*)g.aux; p.ldx = g.ldaux; p.sx1 = g.sx1; p.sx2 = g.sx2;
  p.aux32 = g.aux && ((uintptr_t)g.aux & 31) == 0 && g.ldaux % 16 == 0 && g.sx1 % 16 == 0 && g.sx2 % 16 == 0;
  p.rowvec = g.rowvec; p.sr1 = g.sr1; p.sr2 = g.sr2;
  p.alpha = g.alpha; p.epi = g.epi;
  cudaStream_t st = (cudaStream_t)stream;
  if (pair) {
    if (g.out_dtype == LCASR_BF16)
      return wide ? dispatch_major<256, bf16, 2>(g, p, st) : dispatch_major<128, bf16, 2>(g, p, st);
    return wide ? dispatch_major<256, float, 2>(g, p, st) : dispatch_major<128, float, 2>(g, p, st);
  }
  if (g.out_dtype == LCASR_BF16)
    return wide ? dispatch_major<256, bf16, 1>(g, p, st) : dispatch_major<128, bf16, 1>(g, p, st);
  return wide ? dispatch_major<256, float, 1>(g, p, st) : dispatch_major<128, float, 1>(g, p, st);
}

// P and dS of the attention backward in one pass (see attn_bwd_pds_kernel).  q, k, v, dO: bf16 [nb, N, H, Dh];
// lse2, dvec: fp32 [nb, H, N]; P, dS: bf16 [nb, H, N, Np] with Np = N rounded up to 8.
extern "C" int lcasr_attention_bwd_pds(const void* q, const void* k, const void* v, const void* d_out, const float* lse2,
                                       const float* dvec, int nb, int64_t N, int H, int Dh, void* P, void* dS, void* stream) {
  LCASR_CHECK_ARG(q && k && v && d_out && lse2 && dvec && P && dS, "attention_bwd_pds: NULL argument");
  LCASR_CHECK_ARG(nb > 0 && N > 0 && H > 0 && Dh > 0 && Dh % 8 == 0, "attention_bwd_pds: bad shape");
  LCASR_CHECK_ARG(N < ((int64_t)1 << 30), "attention_bwd_pds: N too large");
  const int64_t d = (int64_t)H * Dh, Np = (N + 7) / 8 * 8;
  CUtensorMap tmQ, tmK, tmDO, tmV;
  LCASR_TRY(make_tmap_4d(&tmQ, q, (uint64_t)N, (uint64_t)Dh, d, H, Dh, nb, N * d, 128, TX_BK));
  LCASR_TRY(make_tmap_4d(&tmK, k, (uint64_t)N, (uint64_t)Dh, d, H, Dh, nb, N * d, 128, TX_BK));
  LCASR_TRY(make_tmap_4d(&tmDO, d_out, (uint64_t)N, (uint64_t)Dh, d, H, Dh, nb, N * d, 128, TX_BK));
  LCASR_TRY(make_tmap_4d(&tmV, v, (uint64_t)N, (uint64_t)Dh, d, H, Dh, nb, N * d, 128, TX_BK));
  CUtensorMap tmP, tmDS;
  LCASR_TRY(make_tmap_4d(&tmP, P, (uint64_t)N, (uint64_t)N, Np, H, N * Np, nb, (int64_t)H * N * Np, 32, 64));
  LCASR_TRY(make_tmap_4d(&tmDS, dS, (uint64_t)N, (uint64_t)N, Np, H, N * Np, nb, (int64_t)H * N * Np, 32, 64));
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(attn_bwd_pds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PDS_SMEM));
    attr_set.mark(attr_dev);
  }
  PdsParams p;
  p.N = N; p.Dh = Dh; p.Nst = (int)Np; p.H = H; p.nbatch = nb;
  p.ldo = Np; p.so1 = N * Np; p.so2 = (int64_t)H * N * Np;
  p.lse = lse2; p.dvec = dvec;
  p.scale = 1.0f / sqrtf((float)Dh);
  p.alpha_p = p.scale * 1.4426950408889634f;
  p.P = (bf16*)P; p.dS = (bf16*)dS;
  const int64_t t1 = ceil_div(N, 128);
  const int64_t tiles = t1 * t1 * H * nb;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  attn_bwd_pds_kernel<<<grid, TX_THREADS, PDS_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmDO, tmV, tmP, tmDS, p);
  LCASR_LAUNCH_CHECK();
  return 0;
}

// Flash-style attention backward (attn_bwd_flash_kernel): q, k, v, o, dO, dq, dk, dv bf16 [nb, n_pitch, H, Dh] of which
// the first N tokens of every recording are valid; lse2 fp32 [nb, H, lse_pitch] (log2-domain, from the training forward);
// workspace: lcasr_attention_bwd_flash_workspace_bytes.  Gradient rows of tokens >= N are left untouched.
extern "C" int64_t lcasr_attention_bwd_flash_workspace_bytes(int nb, int64_t N, int H) {
  if (nb <= 0 || N <= 0 || H <= 0) return -1;
  return (int64_t)nb * H * round_up(N, (int64_t)128) * 8;
}

// one side stream + fork / join events per device (created on first use; LCASR_ATTN_BWD_ONE_STREAM=1: both launches in order)
struct FbSide { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static FbSide* fb_side() {
  static const bool off = getenv("LCASR_ATTN_BWD_ONE_STREAM") != nullptr;
  if (off) return nullptr;
  static FbSide sides[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  FbSide& s = sides[dev];
  if (!s.stream) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      s = FbSide();
      return nullptr;
    }
  }
  return &s;
}

template <int DH>
static int launch_attn_bwd_flash(const void* q, const void* k, const void* v, const void* d_out, int nb, int64_t N, int64_t n_pitch,
                                 int H, const float2* stats, int64_t Npad, void* dq, void* dk, void* dv, cudaStream_t st) {
  const int64_t d = (int64_t)H * DH;
  CUtensorMap kX, vX, qY, oY, qX, oX, kY, vY;
  LCASR_TRY(make_tmap_4d(&kX, k, (uint64_t)N, DH, d, H, DH, nb, n_pitch * d, 128, 64));
  LCASR_TRY(make_tmap_4d(&vX, v, (uint64_t)N, DH, d, H, DH, nb, n_pitch * d, 128, 64));
  LCASR_TRY(make_tmap_4d(&qY, q, (uint64_t)N, DH, d, H, DH, nb, n_pitch * d, FB_CH, 64));
  LCASR_TRY(make_tmap_4d(&oY, d_out, (uint64_t)N, DH, d, H, DH, nb, n_pitch * d, FB_CH, 64));
  LCASR_TRY(make_tmap_4d(&qX, q, (uint64_t)N, DH, d, H, DH, nb, n_pitch * d, 128, 64));
  LCASR_TRY(make_tmap_4d(&oX, d_out, (uint64_t)N, DH, d, H, DH, nb, n_pitch * d, 128, 64));
  LCASR_TRY(make_tmap_4d(&kY, k, (uint64_t)N, DH, d, H, DH, nb, n_pitch * d, FB_CH, 64));
  LCASR_TRY(make_tmap_4d(&vY, v, (uint64_t)N, DH, d, H, DH, nb, n_pitch * d, FB_CH, 64));
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(attn_bwd_flash_kernel<DH, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FbCfg<DH, 0>::SMEM));
    LCASR_CUDA(cudaFuncSetAttribute(attn_bwd_flash_kernel<DH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FbCfg<DH, 1>::SMEM));
    attr_set.mark(attr_dev);
  }
  FbParams p;
  p.N = N; p.Npad = Npad; p.tok_stride = d; p.rec_stride = n_pitch * d; p.H = H;
  p.scale = 1.0f / sqrtf((float)DH);
  p.c_log2 = p.scale * 1.4426950408889634f;
  p.stats = stats;
  p.x1 = (const bf16*)q; p.x2 = (const bf16*)d_out;
  const dim3 grid((unsigned)ceil_div(N, (int64_t)128), (unsigned)H, (unsigned)nb);
  // the two launches are independent: the dQ kernel goes to a side stream so that its CTAs fill the last wave of the dK / dV
  // kernel (768 CTAs on 148 SMs at the training config: 5.19 waves each, 10.4 together instead of 6 + 6)
  FbSide* side = fb_side();
  cudaStream_t st_b = st;
  if (side) {
    LCASR_CUDA(cudaEventRecord(side->fork, st));
    LCASR_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    st_b = side->stream;
  }
  p.out1 = (bf16*)dv; p.out2 = (bf16*)dk;
  attn_bwd_flash_kernel<DH, 0><<<grid, FB_THREADS, FbCfg<DH, 0>::SMEM, st>>>(kX, vX, qY, oY, p);
  LCASR_LAUNCH_CHECK();
  p.out1 = (bf16*)dq; p.out2 = nullptr;
  attn_bwd_flash_kernel<DH, 1><<<grid, FB_THREADS, FbCfg<DH, 1>::SMEM, st_b>>>(qX, oX, kY, vY, p);
  LCASR_LAUNCH_CHECK();
  if (side) {
    LCASR_CUDA(cudaEventRecord(side->join, side->stream));
    LCASR_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  }
  return 0;
}

extern "C" int lcasr_attention_bwd_flash(const void* q, const void* k, const void* v, const void* o, const void* d_out,
                                         const float* lse2, int nb, int64_t N, int64_t n_pitch, int64_t lse_pitch, int H, int Dh,
                                         void* dq, void* dk, void* dv, void* workspace, int64_t workspace_bytes, void* stream) {
  LCASR_CHECK_ARG(q && k && v && o && d_out && lse2 && dq && dk && dv && workspace, "attention_bwd_flash: NULL argument");
  LCASR_CHECK_ARG(nb > 0 && N > 0 && H > 0 && n_pitch >= N && lse_pitch >= N, "attention_bwd_flash: bad shape");
  LCASR_CHECK_ARG(Dh == 64 || Dh == 128, "attention_bwd_flash: head_dim %d not in {64, 128} (use the materialised form)", Dh);
  LCASR_CHECK_ARG(N < ((int64_t)1 << 30), "attention_bwd_flash: N too large");
  LCASR_CHECK_ARG(workspace_bytes >= lcasr_attention_bwd_flash_workspace_bytes(nb, N, H), "attention_bwd_flash: workspace too small");
  LCASR_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "attention_bwd_flash: workspace must be 16-byte aligned");
  const int64_t Npad = round_up(N, (int64_t)128);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)nb * Npad * H;
  attn_bwd_prep_kernel<<<(unsigned)ceil_div(total, (int64_t)256), 256, 0, st>>>((const bf16*)o, (const bf16*)d_out, lse2, nb, N, Npad,
                                                                              n_pitch, lse_pitch, H, Dh, 1.0f / sqrtf((float)Dh),
                                                                              (float2*)workspace);
  LCASR_LAUNCH_CHECK();
  if (Dh == 128)
    return launch_attn_bwd_flash<128>(q, k, v, d_out, nb, N, n_pitch, H, (const float2*)workspace, Npad, dq, dk, dv, st);
  return launch_attn_bwd_flash<64>(q, k, v, d_out, nb, N, n_pitch, H, (const float2*)workspace, Npad, dq, dk, dv, st);
}
