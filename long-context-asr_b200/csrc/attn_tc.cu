// tcgen05 flash attention forward: softmax(Q K^T / sqrt(Dh)) V, non-causal, no mask
// (Attention.forward, lcasr/components/attention.py:509-551; the reference calls flash-attn 2 here).
//
// One CTA = 128 query rows of one (batch, head); K/V streamed in 128-key tiles.
//   warp 0    TMA producer: Q once, then K/V tiles into a STAGES-deep shared-memory ring
//   warp 1    MMA issuer  : S = Q K^T  (tcgen05.mma SS, M=128 N=128 K=Dh) into one of two TMEM S buffers,
//                           O += P V   (tcgen05.mma TS: A = P read straight from tensor memory)
//   warps 2-5 softmax     : thread == query row.  tcgen05.ld S row -> online softmax in fp32 (exp2,
//                           lazy max: O is only rescaled when a row max grows by > 2^8) -> P packed to
//                           bf16 and written back over S with tcgen05.st -> epilogue O / l -> global
// TMEM map (512 columns): S0 [0,128) | S1 [128,256) | O [256,256+Dh).  P(j) aliases S(j) (64 columns).
// Tensor-bound for Dh=128 (512 MMA flops per score); MUFU(ex2)-bound for Dh=32 (SURVEY §7 hard part 2).
// Algorithmic FLOPs per launch = 4 * B * H * N^2 * Dh.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <cstdlib>

namespace lcasr {

using namespace ptx;

constexpr int FA_BQ = 128, FA_BK = 128, FA_THREADS = 192;

template <int DH> struct FaCfg {
  static constexpr int STAGES = DH == 128 ? 2 : (DH == 64 ? 3 : 4);
  static constexpr int ROW_BYTES = DH >= 64 ? 128 : 64;            // swizzle span of a Q/K row segment
  static constexpr int SUB = DH >= 64 ? DH / 64 : 1;                // 64-column sub-tiles per Q/K tile
  static constexpr int SUB_COLS = DH >= 64 ? 64 : DH;
  static constexpr uint32_t QK_LAYOUT = DH >= 64 ? kLayoutSW128 : kLayoutSW64;
  static constexpr int Q_BYTES = FA_BQ * DH * 2;
  static constexpr int K_BYTES = FA_BK * DH * 2;
  static constexpr int V_BYTES = FA_BK * DH * 2;
  static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int O_COL = 256;
};

// VT = true : v is [B,H,Dh,Npad] (keys contiguous)  -> K-major B operand, two [Dh x 64-key] SW128 sub-tiles
// VT = false: v is [B,N,H,Dh]   (natural layout)    -> MN-major B operand, [128 keys x 64 dh] sub-tiles
template <int DH, bool VT>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, int64_t N, int H, float scale_log2, bf16* __restrict__ out) {
  using Cfg = FaCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * Cfg::STAGES + 6];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int s) { return smem_base + Cfg::Q_BYTES + s * Cfg::STAGE_BYTES; };
  auto v_smem = [&](int s) { return smem_base + Cfg::Q_BYTES + s * Cfg::STAGE_BYTES + Cfg::K_BYTES; };
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0;
  auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (1 + Cfg::STAGES + s); };
  auto s_full = [&](int b) { return bar0 + 8u * (1 + 2 * Cfg::STAGES + b); };
  auto p_full = [&](int b) { return bar0 + 8u * (1 + 2 * Cfg::STAGES + 2 + b); };
  auto pv_done = [&](int b) { return bar0 + 8u * (1 + 2 * Cfg::STAGES + 4 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t q0 = (int64_t)blockIdx.x * FA_BQ;
  const int n_tiles = (int)((N + FA_BK - 1) / FA_BK);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(s_full(i), 1); mbar_init(p_full(i), 4); mbar_init(pv_done(i), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);

  if (warp == 0) {
    if (lane == 0) {  // ------------------------- TMA producer -------------------------
      const int row_q = (int)(b * N + q0);
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int i = 0; i < Cfg::SUB; ++i)
        tma_load_2d(q_smem + i * (FA_BQ * Cfg::ROW_BYTES), &tmQ, q_full, h * DH + i * Cfg::SUB_COLS, row_q);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(kv_empty(stage), phase ^ 1);
        mbar_arrive_expect_tx(kv_full(stage), Cfg::STAGE_BYTES);
        const int row_k = (int)(b * N + (int64_t)j * FA_BK);
#pragma unroll
        for (int i = 0; i < Cfg::SUB; ++i)
          tma_load_2d(k_smem(stage) + i * (FA_BK * Cfg::ROW_BYTES), &tmK, kv_full(stage), h * DH + i * Cfg::SUB_COLS, row_k);
        if constexpr (VT) {  // [ (b,h,dh) rows , key cols ]: two 64-key halves of DH rows x 128 B
#pragma unroll
          for (int i = 0; i < 2; ++i)
            tma_load_2d(v_smem(stage) + i * (DH * 128), &tmV, kv_full(stage), j * FA_BK + i * 64, (int)((b * H + h) * DH));
        } else {             // natural layout: 128 key rows x (64 or 32) dh columns per sub-tile
#pragma unroll
          for (int i = 0; i < Cfg::SUB; ++i)
            tma_load_2d(v_smem(stage) + i * (FA_BK * Cfg::ROW_BYTES), &tmV, kv_full(stage), h * DH + i * Cfg::SUB_COLS, row_k);
        }
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ------------------------- MMA issuer -------------------------
      constexpr uint32_t idesc_qk = make_idesc_bf16(FA_BQ, FA_BK, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(FA_BQ, DH, VT ? 0 : 1);
      constexpr uint32_t SBO_QK = 8 * Cfg::ROW_BYTES;
      auto issue_qk = [&](int stage, int buf) {
        const uint32_t d_tmem = tmem_base + buf * FA_BK;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          const int sub = (kk * 16) / Cfg::SUB_COLS, within = (kk * 16) % Cfg::SUB_COLS;
          const uint64_t ad = make_smem_desc_kmajor(q_smem + sub * (FA_BQ * Cfg::ROW_BYTES) + within * 2, SBO_QK, Cfg::QK_LAYOUT);
          const uint64_t bd = make_smem_desc_kmajor(k_smem(stage) + sub * (FA_BK * Cfg::ROW_BYTES) + within * 2, SBO_QK, Cfg::QK_LAYOUT);
          umma_f16_ss(d_tmem, ad, bd, idesc_qk, kk != 0);
        }
      };
      auto issue_pv = [&](int stage, int buf, bool accumulate) {
        const uint32_t d_tmem = tmem_base + Cfg::O_COL;
        const uint32_t a_tmem = tmem_base + buf * FA_BK;  // P: 128 bf16 per row = 64 columns
#pragma unroll
        for (int kk = 0; kk < FA_BK / 16; ++kk) {
          uint64_t bd;
          if constexpr (VT) {
            bd = make_smem_desc_kmajor(v_smem(stage) + (kk / 4) * (DH * 128) + (kk % 4) * 32, 1024, kLayoutSW128);
          } else {  // MN-major: 16 key rows per step; LBO = distance between 64-dh sub-tiles; SBO = 8 key rows
            uint64_t d = 0;
            d |= (uint64_t)(((v_smem(stage) + kk * 16 * Cfg::ROW_BYTES) >> 4) & 0x3FFF);
            d |= (uint64_t)(((FA_BK * Cfg::ROW_BYTES) >> 4) & 0x3FFF) << 16;
            d |= (uint64_t)(((8 * Cfg::ROW_BYTES) >> 4) & 0x3FFF) << 32;
            d |= (uint64_t)1 << 46;
            d |= (uint64_t)Cfg::QK_LAYOUT << 61;
            bd = d;
          }
          umma_f16_ts(d_tmem, a_tmem + kk * 8, bd, idesc_pv, (accumulate || kk != 0) ? 1u : 0u);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(kv_full(0), 0);
      tc_fence_after();
      issue_qk(0, 0);
      umma_commit(s_full(0));
      int stage = 0; uint32_t phase = 0;  // ring position of tile j
      for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        if (j + 1 < n_tiles) {
          int nstage = stage + 1; uint32_t nphase = phase;
          if (nstage == Cfg::STAGES) { nstage = 0; nphase ^= 1; }
          mbar_wait(kv_full(nstage), nphase);
          if (j >= 1) mbar_wait(pv_done(buf ^ 1), ((j - 1) >> 1) & 1);  // P(j-1) consumed: S[(j+1)&1] is free
          tc_fence_after();
          issue_qk(nstage, buf ^ 1);
          umma_commit(s_full(buf ^ 1));
        }
        mbar_wait(p_full(buf), (j >> 1) & 1);  // softmax wrote P(j) (and rescaled O if needed)
        tc_fence_after();
        issue_pv(stage, buf, j > 0);
        umma_commit(kv_empty(stage));
        umma_commit(pv_done(buf));
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {  // ------------------------- softmax / correction / epilogue warps -------------------------
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;  // query row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)lane_base << 16);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int buf = j & 1;
      mbar_wait(s_full(buf), (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[FA_BK];
#pragma unroll
      for (int c = 0; c < FA_BK / 32; ++c) tmem_ld_32x32b_x32(t_lane + buf * FA_BK + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_wait_ld();
      const int64_t valid = N - (int64_t)j * FA_BK;  // keys of this tile that exist
      float mx = -INFINITY;
      if (valid >= FA_BK) {
#pragma unroll
        for (int i = 0; i < FA_BK; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
      } else {
#pragma unroll
        for (int i = 0; i < FA_BK; ++i) {
          if (i >= valid) s[i] = 0xff800000u;  // -inf: masked key
          mx = fmaxf(mx, __uint_as_float(s[i]));
        }
      }
      mx *= scale_log2;
      if (j == 0) {
        m_run = mx;
      } else {
        const bool grow = mx > m_run + 8.0f;  // lazy rescale: P stays <= 2^8, exact after the final O / l
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(pv_done(buf ^ 1), ((j - 1) >> 1) & 1);  // O must be quiescent (PV(j-1) retired)
          tc_fence_after();
          const float m_new = grow ? mx : m_run;
          const float alpha = ex2_approx(m_run - m_new);
          l_run *= alpha;
          m_run = m_new;
#pragma unroll
          for (int c = 0; c < DH / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(t_lane + Cfg::O_COL + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(t_lane + Cfg::O_COL + c * 32, o);
          }
          tmem_wait_st();
        }
      }
      float sum = 0.f;
      uint32_t pk[FA_BK / 2];
#pragma unroll
      for (int i = 0; i < FA_BK; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), scale_log2, -m_run));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), scale_log2, -m_run));
        sum += p0 + p1;
        __nv_bfloat162 pp = __floats2bfloat162_rn(p0, p1);  // .x (low half) = even key
        pk[i / 2] = *reinterpret_cast<uint32_t*>(&pp);
      }
      l_run += sum;
#pragma unroll
      for (int c = 0; c < FA_BK / 64; ++c)
        tmem_st_32x32b_x32(t_lane + buf * FA_BK + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[c * 32]));
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(buf));
    }
    // epilogue: O / l
    const int lastbuf = (n_tiles - 1) & 1;
    mbar_wait(pv_done(lastbuf), ((n_tiles - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int64_t n = q0 + row;
    bf16* orow = out + ((b * N + n) * H + h) * DH;
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_lane + Cfg::O_COL + c * 32, o);
      tmem_wait_ld();
      if (n < N) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(o[g * 8 + i]) * inv_l;
          Vec8<bf16>::store(orow + c * 32 + g * 8, y);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// ================================================================================================
// Version 2 (default for the natural V layout): TWO query tiles per CTA, FA4-style ping-pong.
//   warps 0-3  softmax for query tile A (rows q0 .. q0+127)      warp 8  TMA producer
//   warps 4-7  softmax for query tile B (rows q0+128 .. q0+255)  warp 9  MMA issuer   (10,11 idle)
// TMEM: S_A [0,128) | S_B [128,256) | O_A [256,256+Dh) | O_B [256+Dh,256+2Dh).  P_t aliases S_t.
// The tensor pipe executes one thread's MMAs in issue order, so the issuer simply emits
//   PV_t(j) ; QK_t(j+1)      for t = A, B
// after P_t(j) is published: QK_t(j+1) may overwrite S_t/P_t because PV_t(j) precedes it in the pipe,
// and when s_full[t] fires for tile j+1, PV_t(j) has retired, so the softmax warps may rescale O_t
// without any extra barrier.  While tile A waits for its MMAs, tile B's softmax keeps the MUFU/FMA
// pipes busy (two softmax warps per SM sub-partition), and K/V smem traffic per FLOP is halved.
// ================================================================================================
constexpr int FA2_THREADS = 384;

// The MMA issuer polls several mbarriers.  A hot polling loop competes for issue slots with the two or
// three softmax warps on its SM sub-partition (and every tile waits for its slowest quarter), so the
// loop backs off for a few dozen cycles whenever a sweep found nothing to do.
#ifndef LCASR_POLL_SLEEP_NS
#define LCASR_POLL_SLEEP_NS 0
#endif
#define LCASR_POLL_BACKOFF() ((LCASR_POLL_SLEEP_NS) > 0 ? __nanosleep(LCASR_POLL_SLEEP_NS) : (void)0)

// Optional phase trace of CTA (0,0,0) (tools/trace_attn.py): clock64 stamps of the softmax warps
// (quarter 0 of each query tile) and of the MMA issuer.  Enabled by lcasr_debug_attn_trace(1).
constexpr int kTraceIters = 32;
__device__ long long g_attn_trace[2 * kTraceIters * 8 + 4 * kTraceIters * 4];
__device__ int g_attn_trace_on = 0;
#ifdef LCASR_ATTN_TRACE
#define TRACE_M0 const long long tm0 = clock64()
#define TRACE_M1(t, j) do { if (g_attn_trace_on && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (j) < kTraceIters) { \
    long long* tp = g_attn_trace + 2 * kTraceIters * 8 + ((t) * kTraceIters + (j)) * 4; tp[0] = tm0; tp[1] = clock64(); } } while (0)
#else
#define TRACE_M0 do { } while (0)
#define TRACE_M1(t, j) do { } while (0)
#endif

// PX = packed-exponential variant (POLY == 16, Dh = 32): see the softmax loop
template <int DH, bool PX = false> struct Fa2Cfg {
  static constexpr int STAGES = DH == 128 ? 2 : (DH == 64 ? 3 : 4);
  static constexpr int ROW_BYTES = DH >= 64 ? 128 : 64;
  static constexpr int SUB = DH >= 64 ? DH / 64 : 1;
  static constexpr int SUB_COLS = DH >= 64 ? 64 : DH;
  static constexpr uint32_t LAYOUT = DH >= 64 ? kLayoutSW128 : kLayoutSW64;
  static constexpr int Q_SUB_BYTES = 2 * FA_BQ * ROW_BYTES;   // one 64-col sub-tile of BOTH query tiles
  static constexpr int Q_BYTES = SUB * Q_SUB_BYTES;
  static constexpr int KV_SUB_BYTES = FA_BK * ROW_BYTES;
  static constexpr int K_BYTES = SUB * KV_SUB_BYTES;
  static constexpr int STAGE_BYTES = 2 * K_BYTES;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + 1024;
  // Dh <= 64: P gets its own TMEM columns, so QK(j+1) can be issued as soon as the softmax warps
  // have pulled S(j) into registers (the tensor-pipe round trip leaves the softmax critical path).
  // Dh = 128: no TMEM left for that (2x128 S + 2x128 O = 512): P(j) aliases S(j).
  static constexpr bool SEP_P = DH <= 64;
  static constexpr int P_COL = SEP_P ? 256 : 0;     // + t*64 (SEP_P) / + t*128 (aliased)
  static constexpr int P_STRIDE = SEP_P ? 64 : FA_BK;
  static constexpr int O_COL = SEP_P ? 384 : 256;   // + t*O_STRIDE
  // Dh = 32: the softmax row sums come out of the tensor core for free.  V gets 16 extra "columns" of
  // ones (a constant all-ones MN-atom in shared memory reached through the descriptor's LBO), so PV
  // produces O[:, 32..47] = sum_k P[:, k] — with the bf16-rounded P that the numerator uses — and the
  // softmax warps drop 128 FADDs per row tile (22% of their instructions; they are issue/latency bound).
  // (alone it measured 4 % slower; it is what makes the packed bf16x2 exponential possible, whose results never
  //  exist as fp32 values that could be summed on the CUDA cores)
  static constexpr bool MMA_SUM = PX && DH == 32;
  static constexpr int PV_N = MMA_SUM ? DH + 16 : DH;
  static constexpr int O_STRIDE = MMA_SUM ? 64 : DH;
  static constexpr int ONES_BYTES = MMA_SUM ? FA_BK * ROW_BYTES : 0;
  static constexpr int SMEM_TOTAL = SMEM_BYTES + ONES_BYTES;
};

template <int DH, int POLY, bool WIN>  // WIN: local-attention instantiation (keeps the band code out of the dense kernel)
__global__ void __launch_bounds__(FA2_THREADS, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, int64_t N, int64_t Nk, const int32_t* __restrict__ kv_len, int H,
                float scale_log2, bf16* __restrict__ out, float* __restrict__ lse, int win_left, int win_right,
                float* __restrict__ out32) {
  // out32 (may be NULL; sequence-parallel partial results): the normalised output O / l is written as fp32 [B,N,H,Dh] there
  // INSTEAD of bf16 to `out`; together with lse it is one term of the exact merge over key blocks (attn_merge_kernel)
  // win_left / win_right (-1 = unlimited; self-attention only): local attention, query i sees keys
  // [i - win_left, i + win_right] (flash-attn window_size, attention.py:466,527-530; eval/run.py:38-43).  Key tiles
  // outside the band of this CTA's 256 queries are never loaded; tiles crossing a band edge get an element mask.
  // lse (may be NULL; training): [B,H,N] fp32, log2-domain log-sum-exp of the SCALED scores of every query row
  // (lse2 = max*scale*log2e + log2(sum)), so that the backward recomputes P = exp2(s*scale*log2e - lse2)
  // N = query rows per batch entry, Nk = keys per batch entry (row pitch of k/v); kv_len (may be NULL): the
  // number of VALID keys of each batch entry — the key-padding mask of ragged batches (attention.py:511-547)
  using Cfg = Fa2Cfg<DH, POLY == 16>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * Cfg::STAGES + 9];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int s) { return smem_base + Cfg::Q_BYTES + s * Cfg::STAGE_BYTES; };
  auto v_smem = [&](int s) { return smem_base + Cfg::Q_BYTES + s * Cfg::STAGE_BYTES + Cfg::K_BYTES; };
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0;
  auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (1 + Cfg::STAGES + s); };
  auto s_full = [&](int t) { return bar0 + 8u * (1 + 2 * Cfg::STAGES + t); };       // QK(j) retired
  auto p_full = [&](int t) { return bar0 + 8u * (1 + 2 * Cfg::STAGES + 2 + t); };   // P(j) published
  auto pv_done = [&](int t) { return bar0 + 8u * (1 + 2 * Cfg::STAGES + 4 + t); };  // PV(j) retired
  const uint32_t stagger = bar0 + 8u * (1 + 2 * Cfg::STAGES + 6);  // tile A half-way through its first tile
  auto s_free = [&](int t) { return bar0 + 8u * (1 + 2 * Cfg::STAGES + 7 + t); };   // S(j) is in registers
  const uint32_t ones_smem = smem_base + Cfg::Q_BYTES + Cfg::STAGES * Cfg::STAGE_BYTES;
  if constexpr (Cfg::MMA_SUM) {  // all-ones bf16 tile (swizzle-invariant), visible to the async proxy before any MMA
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_raw + (ones_smem - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < Cfg::ONES_BYTES / 4; i += FA2_THREADS) ones[i] = 0x3F803F80u;
    fence_proxy_async();
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t q0 = (int64_t)blockIdx.x * (2 * FA_BQ);
  const int64_t Nkv = kv_len ? min((int64_t)kv_len[blockIdx.z], Nk) : Nk;  // valid keys of this batch entry
  int n_tiles = (int)((Nkv + FA_BK - 1) / FA_BK);
  int j_lo = 0;  // key tiles [j_lo, j_lo + n_tiles) intersect the band of queries [q0, q0 + 256)
  if constexpr (WIN) {  // (the dense instantiation keeps exactly the expressions it was tuned with)
    int j_hi = n_tiles;
    if (win_left >= 0) j_lo = (int)(max(q0 - win_left, (int64_t)0) / FA_BK);
    if (win_right >= 0) j_hi = (int)min((int64_t)n_tiles, (min(q0 + 2 * FA_BQ - 1, N - 1) + win_right) / FA_BK + 1);
    j_lo = min(j_lo, n_tiles - 1);
    n_tiles = max(j_hi - j_lo, 1);
  }

  if (warp == 8 && lane == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full(t), 1); mbar_init(p_full(t), 4); mbar_init(pv_done(t), 1); mbar_init(s_free(t), 4);
    }
    mbar_init(stagger, 4);
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);


  if (warp >= 8) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp == 8) {
    if (lane == 0) {  // ------------------------- TMA producer -------------------------
      const int row_q = (int)(b * N + q0);
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int i = 0; i < Cfg::SUB; ++i)  // 256-row box: tile A rows then tile B rows
        tma_load_2d(q_smem + i * Cfg::Q_SUB_BYTES, &tmQ, q_full, h * DH + i * Cfg::SUB_COLS, row_q);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(kv_empty(stage), phase ^ 1);
        mbar_arrive_expect_tx(kv_full(stage), Cfg::STAGE_BYTES);
        const int row_k = (int)(b * Nk + (int64_t)(WIN ? j_lo + j : j) * FA_BK);
#pragma unroll
        for (int i = 0; i < Cfg::SUB; ++i) {
          tma_load_2d(k_smem(stage) + i * Cfg::KV_SUB_BYTES, &tmK, kv_full(stage), h * DH + i * Cfg::SUB_COLS, row_k);
          tma_load_2d(v_smem(stage) + i * Cfg::KV_SUB_BYTES, &tmV, kv_full(stage), h * DH + i * Cfg::SUB_COLS, row_k);
        }
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ------------------------- MMA issuer -------------------------
    // The whole warp runs this loop converged (barrier polls are warp-uniform); one elected lane
    // issues the tcgen05 instructions.  Descriptors are a per-stage base + compile-time increments so
    // that each MMA costs one uniform add, not a descriptor rebuild (the issue latency of this thread
    // is on the critical path of every softmax iteration).
    constexpr uint32_t idesc_qk = make_idesc_bf16(FA_BQ, FA_BK, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(FA_BQ, Cfg::PV_N, 1);  // B (= V) is MN-major
    constexpr uint32_t SBO = 8 * Cfg::ROW_BYTES;
    constexpr uint32_t DESC_HI = ((SBO >> 4) & 0x3FFF) | (1u << 14) | (Cfg::LAYOUT << 29);  // SBO | version 1 | swizzle
    constexpr uint32_t LBO_K = 1u << 16;                                                   // unused for K-major
    constexpr uint32_t LBO_V = ((Cfg::KV_SUB_BYTES >> 4) & 0x3FFF) << 16;                  // next 64-dh sub-tile
    auto mk = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
    const uint32_t q_lo0 = (q_smem >> 4) | LBO_K;
    const uint32_t k_lo0 = (k_smem(0) >> 4) | LBO_K;
    // MMA_SUM: the second MN-atom of the B operand is the shared all-ones tile: LBO = ones - V(stage)
    const uint32_t v_lo0 = Cfg::MMA_SUM ? (v_smem(0) >> 4) | ((((ones_smem - v_smem(0)) >> 4) & 0x3FFF) << 16)
                                        : (v_smem(0) >> 4) | LBO_V;
    constexpr uint32_t V_STAGE_STEP = Cfg::MMA_SUM ? (Cfg::STAGE_BYTES >> 4) - ((Cfg::STAGE_BYTES >> 4) << 16)
                                                   : (Cfg::STAGE_BYTES >> 4);  // address up, LBO down
    auto issue_qk = [&](int stage, int t) {
      const uint32_t d_tmem = tmem_base + t * FA_BK;
      const uint32_t a_lo = q_lo0 + t * ((FA_BQ * Cfg::ROW_BYTES) >> 4);
      const uint32_t b_lo = k_lo0 + stage * (Cfg::STAGE_BYTES >> 4);
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const int sub = (kk * 16) / Cfg::SUB_COLS, within = (kk * 16) % Cfg::SUB_COLS;
        umma_f16_ss(d_tmem, mk(a_lo + ((sub * Cfg::Q_SUB_BYTES + within * 2) >> 4)),
                    mk(b_lo + ((sub * Cfg::KV_SUB_BYTES + within * 2) >> 4)), idesc_qk, kk != 0);
      }
    };
    auto issue_pv = [&](int stage, int t, bool accumulate) {
      const uint32_t d_tmem = tmem_base + Cfg::O_COL + t * Cfg::O_STRIDE;
      const uint32_t a_tmem = tmem_base + Cfg::P_COL + t * Cfg::P_STRIDE;
      const uint32_t b_lo = v_lo0 + stage * V_STAGE_STEP;
#pragma unroll
      for (int kk = 0; kk < FA_BK / 16; ++kk)  // 16 key rows per step
        umma_f16_ts(d_tmem, a_tmem + kk * 8, mk(b_lo + ((kk * 16 * Cfg::ROW_BYTES) >> 4)), idesc_pv,
                    (accumulate || kk != 0) ? 1u : 0u);
    };
    mbar_wait(q_full, 0);
    mbar_wait(kv_full(0), 0);
    tc_fence_after();
    if (elect_one()) {
      issue_qk(0, 0); umma_commit(s_full(0));
      issue_qk(0, 1); umma_commit(s_full(1));
    }
    __syncwarp();
    // Event-driven service of the two query tiles (they stay out of phase: one computes exponentials
    // while the other waits for the tensor pipe).  Tiles may drift apart by up to STAGES-1 K/V tiles.
    //   SEP_P : QK_t(j+1) is issued when S_t(j) has been read (s_free), PV_t(j) when P_t(j) is published
    //   else  : PV_t(j) ; QK_t(j+1) are issued together when P_t(j) is published (P aliases S; the
    //           in-order tensor pipe makes the overwrite safe)
    int qk_n[2] = {1, 1};  // QK tiles issued so far
    int jt[2] = {0, 0};    // PV tiles issued so far
    int full_upto = 0;     // highest K/V tile whose kv_full has been observed
    uint32_t idle = 0;
    uint64_t idle_t0 = 0;
    auto kv_landed = [&](int jj) {  // warp-uniform; jj <= full_upto + 1 always holds
      if (jj <= full_upto) return true;
      if (!mbar_test_wait(kv_full(jj % Cfg::STAGES), (jj / Cfg::STAGES) & 1)) return false;
      full_upto = jj;
      return true;
    };
    while (jt[0] < n_tiles || jt[1] < n_tiles) {
      bool progressed = false;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if constexpr (Cfg::SEP_P) {
          const int jq = qk_n[t];
          if (jq < n_tiles && mbar_test_wait(s_free(t), (jq - 1) & 1) && kv_landed(jq)) {
            tc_fence_after();
            if (elect_one()) {
              issue_qk(jq % Cfg::STAGES, t);
              umma_commit(s_full(t));
            }
            __syncwarp();
            qk_n[t] = jq + 1;
            progressed = true;
          }
          const int j = jt[t];
          if (j < n_tiles && mbar_test_wait(p_full(t), j & 1)) {
            tc_fence_after();
            const int stage = j % Cfg::STAGES;
            const bool release = jt[t ^ 1] > j;  // the other tile already consumed K/V(j)
            TRACE_M0;
            if (elect_one()) {
              issue_pv(stage, t, j > 0);
              if (release) umma_commit(kv_empty(stage));
              umma_commit(pv_done(t));
            }
            __syncwarp();
            TRACE_M1(t, j);
            jt[t] = j + 1;
            progressed = true;
          }
        } else {
          const int j = jt[t];
          if (j >= n_tiles || !mbar_test_wait(p_full(t), j & 1)) continue;
          const bool more = j + 1 < n_tiles;
          if (more && !kv_landed(j + 1)) continue;  // K(j+1) not landed yet
          tc_fence_after();
          const int stage = j % Cfg::STAGES;
          const bool release = jt[t ^ 1] > j;
          TRACE_M0;
          if (elect_one()) {
            issue_pv(stage, t, j > 0);
            if (release) umma_commit(kv_empty(stage));
            umma_commit(pv_done(t));
            if (more) {
              issue_qk((j + 1) % Cfg::STAGES, t);
              umma_commit(s_full(t));
            }
          }
          __syncwarp();
          TRACE_M1(t, j);
          jt[t] = j + 1;
          progressed = true;
        }
      }
      if (progressed) {
        idle = 0; idle_t0 = 0;
      } else if (LCASR_POLL_BACKOFF(), (++idle & 0xFFFF) == 0) {  // deadlock guard: trap instead of hanging the GPU
        uint64_t now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (idle_t0 == 0) idle_t0 = now;
        else if (now - idle_t0 > 4000000000ull) {
          if (lane == 0)
            printf("lcasr_b200: attention MMA issuer stalled (block %d,%d,%d tiles %d/%d of %d)\n", blockIdx.x, blockIdx.y,
                   blockIdx.z, jt[0], jt[1], n_tiles);
          asm volatile("trap;");
        }
      }
    }
   }
  } else {  // ------------------------- softmax warps -------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = warp >> 2;                 // query tile of this warp group
    const int lane_base = (warp & 3) * 32;   // TMEM lane quarter == warp_id % 4
    const int row = lane_base + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)lane_base << 16);
    const uint32_t s_addr = t_lane + t * FA_BK;
    const uint32_t p_addr = t_lane + Cfg::P_COL + t * Cfg::P_STRIDE;
    const uint32_t o_addr = t_lane + Cfg::O_COL + t * Cfg::O_STRIDE;
    float m_run = -INFINITY, l_run = 0.f;
    // de-synchronise the two query tiles: B starts when A is half-way through its first exp phase, so
    // that afterwards one tile's MUFU phase overlaps the other's MMA round trip / load / max phases
    // (An explicit MUFU ping-pong between the two warps of a sub-partition — named barriers, FA-3 style —
    //  was tried and measured neutral: one warp alone reaches only 9.5 cycles/exp, two overlapping 8.1.)
    if (t == 1) mbar_wait(stagger, 0);
    uint32_t s_ready = 0;
#ifdef LCASR_ATTN_TRACE  // compile with -DLCASR_ATTN_TRACE for tools/trace_attn.py (costs ~5% otherwise)
    const bool tr = g_attn_trace_on && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (warp & 3) == 0 && lane == 0;
#define TRACE_S(k) do { if (tr && j < kTraceIters) g_attn_trace[(t * kTraceIters + j) * 8 + (k)] = clock64(); } while (0)
#else
#define TRACE_S(k) do { } while (0)
#endif
    for (int j = 0; j < n_tiles; ++j) {
      TRACE_S(0);
      if (!s_ready) mbar_wait(s_full(t), j & 1);  // usually already observed during the previous iteration
      TRACE_S(1);
      tc_fence_after();
      // non-blocking probe now, consumed later: the barrier latency hides behind the TMEM load
      uint32_t pv_ok = 1;
      if constexpr (Cfg::SEP_P) { if (j > 0) pv_ok = mbar_test_wait(pv_done(t), (j - 1) & 1); }
      uint32_t s[FA_BK];
#pragma unroll
      for (int c = 0; c < FA_BK / 32; ++c) tmem_ld_32x32b_x32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_wait_ld();
      TRACE_S(2);
      if constexpr (Cfg::SEP_P) {  // S_t is in registers: the issuer may overwrite it with QK_t(j+1) now
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free(t));
      }
      const int64_t kbase = (int64_t)(WIN ? j_lo + j : j) * FA_BK;  // first key of this tile
      const int64_t valid = Nkv - kbase;
      if (valid < FA_BK) {
#pragma unroll
        for (int i = 0; i < FA_BK; ++i)
          if (i >= valid) s[i] = 0xff800000u;  // -inf: key does not exist
      }
      if constexpr (WIN) {  // band edges: warp-uniform test on the warp's 32 rows, element mask only in edge tiles
        const int64_t r_first = q0 + t * FA_BQ + lane_base;
        const bool edge = (win_left >= 0 && r_first + 31 - win_left > kbase) ||
                          (win_right >= 0 && r_first + win_right < kbase + FA_BK - 1);
        if (edge) {
          const int64_t r = r_first + lane;
          const int lo_i = win_left >= 0 ? (int)max(r - win_left - kbase, (int64_t)0) : 0;
          const int hi_i = win_right >= 0 ? (int)min(r + win_right - kbase, (int64_t)FA_BK - 1) : FA_BK - 1;
#pragma unroll
          for (int i = 0; i < FA_BK; ++i)
            if (i < lo_i || i > hi_i) s[i] = 0xff800000u;
        }
      }
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < FA_BK; i += 8) {  // FMNMX3: two values per instruction, four independent chains
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm("max.f32 %0, %0, %1, %2;" : "+f"(mxa[u]) : "f"(__uint_as_float(s[i + 2 * u])), "f"(__uint_as_float(s[i + 2 * u + 1])));
      }
      // (local attention: a row whose keys in this tile are all masked has mx = -inf; keep the running maximum finite)
      float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * scale_log2;
      if constexpr (WIN) mx = fmaxf(mx, -1e30f);
      if (j == 0) {
        m_run = mx;
      } else {
        // packed-exponential variant: the exponent argument is rounded to bf16, so it must stay <= 0 (running max exact)
        const bool grow = mx > m_run + (POLY == 16 ? 0.0f : 8.0f);
        if (__any_sync(0xffffffffu, grow)) {
          // O_t must be quiescent.  Aliased P: PV_t(j-1) precedes QK_t(j) in the tensor pipe, so it has
          // retired when s_full fired.  Separate P: wait for its commit.
          if constexpr (Cfg::SEP_P) { if (!pv_ok) { mbar_wait(pv_done(t), (j - 1) & 1); pv_ok = 1; } tc_fence_after(); }
          const float m_new = grow ? mx : m_run;
          const float alpha = ex2_approx(m_run - m_new);
          l_run *= alpha;
          m_run = m_new;
#pragma unroll
          for (int c = 0; c < DH / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(o_addr + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(o_addr + c * 32, o);
          }
          if constexpr (Cfg::MMA_SUM) {  // the row-sum columns live in O as well
            uint32_t o2[16];
            tmem_ld_32x32b_x16(o_addr + DH, o2);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o2[i] = __float_as_uint(__uint_as_float(o2[i]) * alpha);
            tmem_st_32x32b_x16(o_addr + DH, o2);
          }
          tmem_wait_st();
        }
      }
      TRACE_S(3);
      float sums[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // independent chains: no serial FADD latency
      const float neg_m = -m_run;
      // separate P buffer: PV_t(j-1) must have consumed P_t(j-1) before it is overwritten (long retired
      // in steady state — the wait is a formality, but it is what makes the reuse legal)
      if constexpr (Cfg::SEP_P) {
        if (!pv_ok) mbar_wait(pv_done(t), (j - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < FA_BK / 64; ++c) {
        uint32_t pk[32];
        if constexpr (POLY == 16) {
          // Packed exponential: P is needed as bf16 anyway, so two arguments are rounded to bf16x2 and ONE MUFU
          // instruction (ex2.approx.ftz.bf16x2) produces two probabilities already in the operand format — half the
          // MUFU work (the binding unit at Dh = 32: 74 % busy) and 4 instead of 7 issue slots per pair (no fp32 sums:
          // the row sums come from the tensor core, MMA_SUM; no separate pack).  Arguments are <= 0 (exact running
          // max), so their bf16 rounding error is <= 2^-9 for the terms that matter (|x| < 1: 0.14 % on p, the size
          // of P's own bf16 rounding).
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x0 = fmaf(__uint_as_float(s[c * 64 + 2 * i]), scale_log2, neg_m);
            const float x1 = fmaf(__uint_as_float(s[c * 64 + 2 * i + 1]), scale_log2, neg_m);
            uint32_t xx;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xx) : "f"(x1), "f"(x0));  // {hi, lo} = {x1, x0}
            asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(pk[i]) : "r"(xx));
          }
        } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x0 = fmaf(__uint_as_float(s[c * 64 + 2 * i]), scale_log2, neg_m);
          const float x1 = fmaf(__uint_as_float(s[c * 64 + 2 * i + 1]), scale_log2, neg_m);
          // POLY == 9 / 8: timing experiments only (no exponentials at all / all on the FMA pipes)
          const float p0 = POLY == 9 ? x0 : (POLY == 8 ? ex2_poly(x0) : ex2_approx(x0));
          const float p1 = POLY == 9 ? x1 : (POLY == 8 ? ex2_poly(x1)
                           : ((POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == 0) ? ex2_poly(x1) : ex2_approx(x1)));
          if constexpr (!Cfg::MMA_SUM) { sums[(2 * i) & 7] += p0; sums[(2 * i + 1) & 7] += p1; }
          __nv_bfloat162 pp = __floats2bfloat162_rn(p0, p1);
          pk[i] = *reinterpret_cast<uint32_t*>(&pp);
        }
        }
        tmem_st_32x32b_x32(p_addr + c * 32, pk);
        if (c == 0 && j == 0 && t == 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(stagger);
        }
      }
      if constexpr (Cfg::SEP_P)
        s_ready = (j + 1 < n_tiles) ? mbar_test_wait(s_full(t), (j + 1) & 1) : 0;  // probe next S early
      l_run += ((sums[0] + sums[1]) + (sums[2] + sums[3])) + ((sums[4] + sums[5]) + (sums[6] + sums[7]));
      TRACE_S(4);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(t));
      TRACE_S(5);
    }
#undef TRACE_S
    mbar_wait(pv_done(t), (n_tiles - 1) & 1);
    tc_fence_after();
    if constexpr (Cfg::MMA_SUM) {
      uint32_t o2[16];
      tmem_ld_32x32b_x16(o_addr + DH, o2);
      tmem_wait_ld();
      l_run = __uint_as_float(o2[0]);
    }
    const float inv_l = 1.0f / l_run;
    const int64_t n = q0 + t * FA_BQ + row;
    const int64_t o_off = ((b * N + n) * H + h) * DH;
    if (lse && n < N) lse[(b * H + h) * N + n] = m_run + log2f(l_run);
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_addr + c * 32, o);
      tmem_wait_ld();
      if (n < N) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(o[g * 8 + i]) * inv_l;
          if (out32) Vec8<float>::store(out32 + o_off + c * 32 + g * 8, y);
          else Vec8<bf16>::store(out + o_off + c * 32 + g * 8, y);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH, int POLY, bool WIN = false>
static int launch_attn_tc2(const void* q, const void* k, const void* v, int B, int64_t N, int64_t Nk, const int32_t* kv_len,
                           int H, void* out, float* lse, int wl, int wr, cudaStream_t st, float* out32, int64_t ldq, int64_t ldkv) {
  using Cfg = Fa2Cfg<DH, POLY == 16>;
  const uint64_t d = (uint64_t)H * DH;
  const CUtensorMapSwizzle sw = DH >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmQ, tmK, tmV;
  // ldq / ldkv: row pitch in elements (>= H*DH): q, k, v may be column blocks of a wider row-major matrix (the qkv projection)
  LCASR_TRY(make_tmap_2d_bf16(&tmQ, q, (uint64_t)B * N, d, (uint64_t)ldq * 2, 2 * FA_BQ, Cfg::SUB_COLS, sw));
  LCASR_TRY(make_tmap_2d_bf16(&tmK, k, (uint64_t)B * Nk, d, (uint64_t)ldkv * 2, FA_BK, Cfg::SUB_COLS, sw));
  LCASR_TRY(make_tmap_2d_bf16(&tmV, v, (uint64_t)B * Nk, d, (uint64_t)ldkv * 2, FA_BK, Cfg::SUB_COLS, sw));
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<DH, POLY, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_TOTAL));
    attr_set.mark(attr_dev);
  }
  dim3 grid((unsigned)ceil_div(N, 2 * FA_BQ), (unsigned)H, (unsigned)B);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)DH);
  attn_tc2_kernel<DH, POLY, WIN><<<grid, FA2_THREADS, Cfg::SMEM_TOTAL, st>>>(tmQ, tmK, tmV, N, Nk, kv_len, H, scale_log2, (bf16*)out, lse, wl, wr, out32);
  LCASR_LAUNCH_CHECK();
  return 0;
}


int attn_tc_available() { return 1; }
int attn_tc4_launch(const void* q, const void* k, const void* v, int B, int64_t N, int64_t Nk, const int32_t* kv_len, int H,
                    void* out, float* lse, cudaStream_t st, float* out32, int64_t ldq, int64_t ldkv);

}  // namespace lcasr
// debug hooks (not part of the public header): phase trace of the v2 attention kernel
extern "C" int lcasr_debug_attn_trace(int enable) {
  int v = enable;
  return cudaMemcpyToSymbol(lcasr::g_attn_trace_on, &v, sizeof(int)) == cudaSuccess ? 0 : -3;
}
extern "C" int lcasr_debug_attn_trace_read(long long* host, int n) {
  const int total = (int)(sizeof(lcasr::g_attn_trace) / sizeof(long long));
  if (n > total) n = total;
  return cudaMemcpyFromSymbol(host, lcasr::g_attn_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? n : -3;
}
namespace lcasr {

template <int DH, bool VT>
static int launch_attn_tc(const void* q, const void* k, const void* v, int B, int64_t N, int H, int64_t Npad, void* out,
                          cudaStream_t st) {
  using Cfg = FaCfg<DH>;
  const uint64_t d = (uint64_t)H * DH;
  const CUtensorMapSwizzle sw = DH >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmQ, tmK, tmV;
  LCASR_TRY(make_tmap_2d_bf16(&tmQ, q, (uint64_t)B * N, d, d * 2, FA_BQ, Cfg::SUB_COLS, sw));
  LCASR_TRY(make_tmap_2d_bf16(&tmK, k, (uint64_t)B * N, d, d * 2, FA_BK, Cfg::SUB_COLS, sw));
  if (VT) LCASR_TRY(make_tmap_2d_bf16(&tmV, v, (uint64_t)B * H * DH, (uint64_t)Npad, (uint64_t)Npad * 2, DH, 64, CU_TENSOR_MAP_SWIZZLE_128B));
  else LCASR_TRY(make_tmap_2d_bf16(&tmV, v, (uint64_t)B * N, d, d * 2, FA_BK, Cfg::SUB_COLS, sw));
  static PerDeviceFlag attr_set;
  int attr_dev = 0;
  if (attr_set.needs_set(&attr_dev)) {
    LCASR_CUDA(cudaFuncSetAttribute(attn_tc_kernel<DH, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set.mark(attr_dev);
  }
  dim3 grid((unsigned)ceil_div(N, FA_BQ), (unsigned)H, (unsigned)B);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)DH);
  attn_tc_kernel<DH, VT><<<grid, FA_THREADS, Cfg::SMEM_BYTES, st>>>(tmQ, tmK, tmV, N, H, scale_log2, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

int attn_tc_launch(const void* q, const void* k, const void* v, int B, int64_t N, int64_t Nk, const int32_t* kv_len, int H,
                   int Dh, int v_transposed, int64_t Npad, void* out, float* lse, cudaStream_t st, int wl, int wr, float* out32,
                   int64_t ldq, int64_t ldkv) {
  if (ldq <= 0) ldq = (int64_t)H * Dh;
  if (ldkv <= 0) ldkv = (int64_t)H * Dh;
  LCASR_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldq >= (int64_t)H * Dh && ldkv >= (int64_t)H * Dh, "attention(tcgen05): bad row pitch");
  LCASR_CHECK_ARG((!out32 && ldq == (int64_t)H * Dh && ldkv == (int64_t)H * Dh) || (!v_transposed && wl < 0 && wr < 0),
                  "attention(tcgen05): fp32 partial outputs / row pitches need the dense natural-layout kernel");
  LCASR_CHECK_ARG(!out32 || lse, "attention(tcgen05): a partial result needs its log-sum-exp output");
  LCASR_CHECK_ARG((wl < 0 && wr < 0) || (!v_transposed && N == Nk), "attention(tcgen05): a window needs self-attention and the natural V layout");
  LCASR_CHECK_ARG(!lse || !v_transposed, "attention(tcgen05): the log-sum-exp output needs the natural V layout");
  LCASR_CHECK_ARG(!kv_len || !v_transposed, "attention(tcgen05): key lengths need the natural V layout");
  LCASR_CHECK_ARG(Nk == N || !v_transposed, "attention(tcgen05): Nq != Nk needs the natural V layout");
  LCASR_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)out & 15) == 0,
                  "attention(tcgen05): q, k, v, out must be 16-byte aligned");
  LCASR_CHECK_ARG((int64_t)B * N < ((int64_t)1 << 31) && (int64_t)B * Nk < ((int64_t)1 << 31), "attention(tcgen05): B*N too large");
  LCASR_CHECK_ARG(!v_transposed || Npad % 8 == 0, "attention(tcgen05): Npad must be a multiple of 8");
  LCASR_CHECK_ARG(H <= 65535 && B <= 65535, "attention(tcgen05): too many heads / batch entries");
  // head dim 32: the four-stream kernel (attn_tc4.cu; two independent 64-key softmax streams per query tile)
  // (default for Dh = 32: 1.94 vs 2.25 ms at N = 16384, H = 24; LCASR_ATTN_V4=0 selects the two-stream kernel for A/B runs)
  static const int v4 = getenv("LCASR_ATTN_V4") ? atoi(getenv("LCASR_ATTN_V4")) : 1;
  if (v4 && Dh == 32 && !v_transposed && wl < 0 && wr < 0)
    return attn_tc4_launch(q, k, v, B, N, Nk, kv_len, H, out, lse, st, out32, ldq, ldkv);
  static const bool force_v1 = getenv("LCASR_ATTN_V1") != nullptr;  // debugging aid: one query tile per CTA
  // fraction of exponentials evaluated on the FMA pipes: POLY=p -> every p-th odd key, i.e. 1/(2p) of all
  static const int poly_env = getenv("LCASR_ATTN_POLY") ? atoi(getenv("LCASR_ATTN_POLY")) : 0;  // measured: the Dh=32 kernel is latency- not MUFU-bound, offload does not pay yet
  const int poly = (poly_env == 16 && Dh != 32) ? 0 : poly_env;  // the packed exponential exists for head dim 32 only
#define LCASR_FA(DHV)                                                                                             \
  case DHV:                                                                                                       \
    if (v_transposed) return launch_attn_tc<DHV, true>(q, k, v, B, N, H, Npad, out, st);                          \
    if (wl >= 0 || wr >= 0) return launch_attn_tc2<DHV, 4, true>(q, k, v, B, N, Nk, kv_len, H, out, lse, wl, wr, st, nullptr, ldq, ldkv);                 \
    if (force_v1 && Nk == N && !kv_len && !lse && !out32 && ldq == (int64_t)H * Dh && ldkv == ldq) return launch_attn_tc<DHV, false>(q, k, v, B, N, H, Npad, out, st);                  \
    switch (poly) {                                                                                               \
      case 0: return launch_attn_tc2<DHV, 0>(q, k, v, B, N, Nk, kv_len, H, out, lse, wl, wr, st, out32, ldq, ldkv);                                          \
      case 1: return launch_attn_tc2<DHV, 1>(q, k, v, B, N, Nk, kv_len, H, out, lse, wl, wr, st, out32, ldq, ldkv);                                          \
      case 2: return launch_attn_tc2<DHV, 2>(q, k, v, B, N, Nk, kv_len, H, out, lse, wl, wr, st, out32, ldq, ldkv);                                          \
      case 8: return launch_attn_tc2<DHV, 8>(q, k, v, B, N, Nk, kv_len, H, out, lse, wl, wr, st, out32, ldq, ldkv);                                          \
      case 9: return launch_attn_tc2<DHV, 9>(q, k, v, B, N, Nk, kv_len, H, out, lse, wl, wr, st, out32, ldq, ldkv);                                          \
      case 16: return launch_attn_tc2<DHV, 16>(q, k, v, B, N, Nk, kv_len, H, out, lse, wl, wr, st, out32, ldq, ldkv);                                          \
      default: return launch_attn_tc2<DHV, 4>(q, k, v, B, N, Nk, kv_len, H, out, lse, wl, wr, st, out32, ldq, ldkv);                                         \
    }
  switch (Dh) {
    LCASR_FA(32) LCASR_FA(64) LCASR_FA(128)
    default:
      return set_error(LCASR_E_UNSUPPORTED, "attention(tcgen05): head_dim=%d not in {32,64,128}", Dh);
  }
#undef LCASR_FA
}

}  // namespace lcasr
