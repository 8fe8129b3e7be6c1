// tcgen05 flash attention, head dim 32, FOUR softmax streams per CTA (Attention.forward, lcasr/components/attention.py:509-551).
//
// At Dh = 32 a 128x128 score tile costs 256 tensor-pipe cycles (QK^T + PV) but 1024 MUFU cycles (one ex2 per score at
// 16/clk/SM): the kernel is bound by the exponentials, and the two-stream kernel (attn_tc2_kernel: one 128-column row per
// thread, two warps per SM sub-partition) left the MUFU pipe idle a quarter of the time — a single warp in its exp phase
// cannot feed the pipe (dependent ffma -> ex2 -> fadd chains: 9.5 instead of 8 cycles per exponential) and its load /
// max / store phases have only ONE other warp to hide behind.
//
// Here every query tile is processed as TWO independent online-softmax streams, one per 64-key half of each key tile:
//   stream (t, h): query tile t in {A, B}, keys [64h, 64h + 64) of every 128-key tile, own running max / sum and its OWN
//   output accumulator O[t][h] in tensor memory; the two halves are merged exactly at the end (flash-decoding style:
//   O = (O_0 2^(m_0 - M) + O_1 2^(m_1 - M)) / (l_0 2^(m_0 - M) + l_1 2^(m_1 - M))).
// 16 softmax warps = 4 per SM sub-partition, each handling 32 rows x 64 columns per key tile.
//
//   warps 0-15  softmax: warp w -> TMEM lane quarter w % 4, stream w / 4 (t = stream / 2, h = stream % 2)
//   warp 16     TMA producer (Q once, K/V tiles into a 4-deep ring)
//   warp 17     QK^T issuer (both query tiles)       warps 18, 19   PV issuers (query tile A / B)
// TMEM (512 columns): S_A [0,128) | S_B [128,256) | P[t][h] 4 x 32 at 256 | O[t][h] 4 x 32 at 384.
// Per key tile the issuer emits QK_t (M128 N128 K32) when S_t has been read by all 8 warps of tile t, and
// PV_{t,h} (M128 N32 K64, A = P from tensor memory, B = V rows [64h, 64h+64) MN-major) when stream (t, h) published P.
// Algorithmic FLOPs per launch = 4 * B * H * Nq * Nk * 32.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <cstdlib>

namespace lcasr {

using namespace ptx;

namespace {

constexpr int F4_DH = 32, F4_BQ = 128, F4_BK = 128, F4_HK = 64;  // HK: keys per stream and tile
constexpr int F4_THREADS = 640, F4_STAGES = 4;  // 16 softmax warps + one warpgroup for the TMA producer and the MMA issuer
// Registers: the SM allocates them per 4 warps, so 20 warps launch with 96 each (61440 in all) and setmaxnreg can only
// REDISTRIBUTE that total: producer / issuer warpgroup down to 32, softmax warps up to 112 (128*32 + 512*112 = 61440).
constexpr int F4_ROW_BYTES = 64;                       // one 32-element bf16 row; SWIZZLE_64B
constexpr int F4_Q_BYTES = 2 * F4_BQ * F4_ROW_BYTES;   // both query tiles
constexpr int F4_K_BYTES = F4_BK * F4_ROW_BYTES;
constexpr int F4_STAGE_BYTES = 2 * F4_K_BYTES;
constexpr int F4_EXO_STRIDE = 33;                      // fp32 words per exchanged row (conflict-free)
constexpr int F4_EX_BYTES = 2 * F4_BQ * (F4_EXO_STRIDE + 2) * 4;
constexpr int F4_SMEM_BYTES = F4_Q_BYTES + F4_STAGES * F4_STAGE_BYTES + F4_EX_BYTES + 1024;
constexpr int F4_P_COL = 256, F4_O_COL = 384;
constexpr int kF4DefaultPoly = 0, kF4DefaultIpack = 0;
constexpr float kF4TruncBias = 0.7213475f / 256.f;            // mean relative loss of truncating fp32 -> bf16
constexpr float kF4TruncBiasLog2 = 1.4426950f * kF4TruncBias;  // log2(1 + b) ~ b / ln 2 (b = 2.8e-3: error 4e-6)

// POLY = p > 0: every p-th odd key's exponential is evaluated on the FMA / ALU pipes (ex2_poly: Cody-Waite split + cubic,
// relative error 1.6e-4, far below the bf16 rounding of P) instead of the MUFU pipe — 1/(2p) of all exponentials.
// IPACK: P is packed to bf16 with integer arithmetic (+0x8000 on the fp32 bits = round half up, then one byte permute per
// pair) instead of F2FP: ncu shows the conversion on the XU pipe next to MUFU.EX2 (81 % busy, 14 points of it not ex2).
template <int POLY, int IPACK>
__global__ void __launch_bounds__(F4_THREADS, 1)
attn_tc4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, int64_t N, int64_t Nk, const int32_t* __restrict__ kv_len, int H,
                float scale_log2, bf16* __restrict__ out, float* __restrict__ lse, float* __restrict__ out32) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * F4_STAGES + 2 + 2 + 4 + 4];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int s) { return smem_base + F4_Q_BYTES + s * F4_STAGE_BYTES; };
  auto v_smem = [&](int s) { return smem_base + F4_Q_BYTES + s * F4_STAGE_BYTES + F4_K_BYTES; };
  float* ex = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + F4_Q_BYTES + F4_STAGES * F4_STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0;
  auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (1 + F4_STAGES + s); };
  constexpr int BB = 1 + 2 * F4_STAGES;
  auto s_full = [&](int t) { return bar0 + 8u * (BB + t); };              // QK_t(j) retired
  auto s_free = [&](int t) { return bar0 + 8u * (BB + 2 + t); };          // S_t(j) is in the registers of all 8 warps of tile t
  auto p_full = [&](int st) { return bar0 + 8u * (BB + 4 + st); };        // P of stream st published
  auto pv_done = [&](int st) { return bar0 + 8u * (BB + 8 + st); };       // PV of stream st retired

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h_idx = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int64_t q0 = (int64_t)blockIdx.x * (2 * F4_BQ);
  const int64_t Nkv = kv_len ? min((int64_t)kv_len[blockIdx.z], Nk) : Nk;
  const int n_tiles = (int)((Nkv + F4_BK - 1) / F4_BK);

  if (warp == 16 && lane == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < F4_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2); }  // kv_empty: one commit per PV warp
    for (int t = 0; t < 2; ++t) { mbar_init(s_full(t), 1); mbar_init(s_free(t), 8); }
    for (int st = 0; st < 4; ++st) { mbar_init(p_full(st), 4); mbar_init(pv_done(st), 1); }
    fence_barrier_init();
  }
  if (warp == 17) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 16) {
      if (lane == 0) {  // ------------------------- TMA producer -------------------------
        const int row_q = (int)(b * N + q0);
        mbar_arrive_expect_tx(q_full, F4_Q_BYTES);
        tma_load_2d(q_smem, &tmQ, q_full, h_idx * F4_DH, row_q);  // 256-row box: tile A rows then tile B rows
        int stage = 0; uint32_t phase = 0;
        for (int j = 0; j < n_tiles; ++j) {
          mbar_wait(kv_empty(stage), phase ^ 1);
          mbar_arrive_expect_tx(kv_full(stage), F4_STAGE_BYTES);
          const int row_k = (int)(b * Nk + (int64_t)j * F4_BK);
          tma_load_2d(k_smem(stage), &tmK, kv_full(stage), h_idx * F4_DH, row_k);
          tma_load_2d(v_smem(stage), &tmV, kv_full(stage), h_idx * F4_DH, row_k);
          if (++stage == F4_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else {
      // ------------------------- MMA issuers: warp 17 issues QK^T of both query tiles, warps 18 / 19 issue the PV products
      // of query tile A / B (two streams each).  Three small polling loops instead of one that watches six barriers: the
      // reaction time from "P published" to "PV retired" is on the softmax warps' critical path, and one loop with all the
      // state did not fit the 32 registers these warps keep.  Whole warp converged, one elected lane issues. ---------------
      constexpr uint32_t idesc_qk = make_idesc_bf16(F4_BQ, F4_BK, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(F4_BQ, F4_DH, 1);  // B (= V) is MN-major
      constexpr uint32_t SBO = 8 * F4_ROW_BYTES;
      constexpr uint32_t DESC_HI = ((SBO >> 4) & 0x3FFF) | (1u << 14) | (kLayoutSW64 << 29);
      constexpr uint32_t LBO_K = 1u << 16;                                 // unused for swizzled K-major operands
      constexpr uint32_t LBO_V = ((F4_K_BYTES >> 4) & 0x3FFF) << 16;      // MN-major V: next 64-dh sub-tile (there is one)
      auto mk = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
      uint32_t idle = 0;
      uint64_t idle_t0 = 0;
      auto guard = [&](bool progressed) {  // deadlock guard: trap instead of hanging the GPU (no printf: 32 registers)
        if (progressed) { idle = 0; idle_t0 = 0; return; }
        if ((++idle & 0xFFFF) != 0) return;
        uint64_t now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (idle_t0 == 0) idle_t0 = now;
        else if (now - idle_t0 > 4000000000ull) asm volatile("trap;");
      };
      if (warp == 17) {
        const uint32_t q_lo0 = (q_smem >> 4) | LBO_K;
        const uint32_t k_lo0 = (k_smem(0) >> 4) | LBO_K;
        auto issue_qk = [&](int stage, int t) {
          const uint32_t d_tmem = tmem_base + t * F4_BK;
          const uint32_t a_lo = q_lo0 + t * ((F4_BQ * F4_ROW_BYTES) >> 4);
          const uint32_t b_lo = k_lo0 + stage * (F4_STAGE_BYTES >> 4);
#pragma unroll
          for (int kk = 0; kk < F4_DH / 16; ++kk)
            umma_f16_ss(d_tmem, mk(a_lo + ((kk * 32) >> 4)), mk(b_lo + ((kk * 32) >> 4)), idesc_qk, kk != 0);
        };
        mbar_wait(q_full, 0);
        mbar_wait(kv_full(0), 0);
        tc_fence_after();
        if (elect_one()) {
          issue_qk(0, 0); umma_commit(s_full(0));
          issue_qk(0, 1); umma_commit(s_full(1));
        }
        __syncwarp();
        int qk0 = 1, qk1 = 1;  // QK tiles issued per query tile
        int full_upto = 0;
        auto kv_landed = [&](int jj) {
          if (jj <= full_upto) return true;
          if (!mbar_test_wait(kv_full(jj % F4_STAGES), (jj / F4_STAGES) & 1)) return false;
          full_upto = jj;
          return true;
        };
        while (qk0 < n_tiles || qk1 < n_tiles) {
          bool progressed = false;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int jq = t == 0 ? qk0 : qk1;
            if (jq < n_tiles && mbar_test_wait(s_free(t), (jq - 1) & 1) && kv_landed(jq)) {
              tc_fence_after();
              if (elect_one()) {
                issue_qk(jq % F4_STAGES, t);
                umma_commit(s_full(t));
              }
              __syncwarp();
              if (t == 0) qk0 = jq + 1; else qk1 = jq + 1;
              progressed = true;
            }
          }
          guard(progressed);
        }
      } else {
        const int t = warp - 18;  // query tile whose two streams this warp serves
        const uint32_t v_lo0 = (v_smem(0) >> 4) | LBO_V;
        auto issue_pv = [&](int stage, int st, bool accumulate) {
          const uint32_t d_tmem = tmem_base + F4_O_COL + st * F4_DH;
          const uint32_t a_tmem = tmem_base + F4_P_COL + st * (F4_HK / 2);
          const uint32_t b_lo = v_lo0 + stage * (F4_STAGE_BYTES >> 4) + (((st & 1) * F4_HK * F4_ROW_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < F4_HK / 16; ++kk)  // 16 key rows per step
            umma_f16_ts(d_tmem, a_tmem + kk * 8, mk(b_lo + ((kk * 16 * F4_ROW_BYTES) >> 4)), idesc_pv, (accumulate || kk != 0) ? 1u : 0u);
        };
        int pv0 = 0, pv1 = 0;  // PV tiles issued for key half 0 / 1
        while (pv0 < n_tiles || pv1 < n_tiles) {
          bool progressed = false;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int j = hh == 0 ? pv0 : pv1;
            const int st = 2 * t + hh;
            if (j < n_tiles && mbar_test_wait(p_full(st), j & 1)) {
              tc_fence_after();
              const int stage = j % F4_STAGES;
              // this warp's share of K/V tile j is consumed once BOTH of its PV(j) have been issued (kv_empty counts the
              // commits of the two PV warps; QK_t(j) retired before P_t(j) could exist)
              const bool release = (hh == 0 ? pv1 : pv0) > j;
              if (elect_one()) {
                issue_pv(stage, st, j > 0);
                if (release) umma_commit(kv_empty(stage));
                umma_commit(pv_done(st));
              }
              __syncwarp();
              if (hh == 0) pv0 = j + 1; else pv1 = j + 1;
              progressed = true;
            }
          }
          guard(progressed);
        }
      }
    }
  } else {  // ------------------------- softmax warps -------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int st = warp >> 2;                // stream
    const int t = st >> 1, hh = st & 1;      // query tile, key half
    const int lane_base = (warp & 3) * 32;   // TMEM lane quarter == warp_id % 4
    const int row = lane_base + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)lane_base << 16);
    const uint32_t s_addr = t_lane + t * F4_BK + hh * F4_HK;
    const uint32_t p_addr = t_lane + F4_P_COL + st * (F4_HK / 2);
    const uint32_t o_addr = t_lane + F4_O_COL + st * F4_DH;
    float m_run = -1e30f, l_run = 0.f;
    // (Starting stream k only when stream k-1 is half-way through its first exp phase — the de-phasing that helps the
    //  two-stream kernel — measured SLOWER here: 2.21 vs 2.11 ms at N = 16384, H = 24.)
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full(t), j & 1);
      tc_fence_after();
      uint32_t pv_ok = 1;
      if (j > 0) pv_ok = mbar_test_wait(pv_done(st), (j - 1) & 1);
      uint32_t s[F4_HK];
      tmem_ld_32x32b_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_32x32b_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(t));  // the issuer may overwrite S_t with QK_t(j+1) once all 8 warps arrived
      const int64_t valid = Nkv - ((int64_t)j * F4_BK + hh * F4_HK);  // keys of this half that exist (may be <= 0)
      if (valid < F4_HK) {
#pragma unroll
        for (int i = 0; i < F4_HK; ++i)
          if (i >= valid) s[i] = 0xff800000u;  // -inf
      }
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < F4_HK; i += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm("max.f32 %0, %0, %1, %2;" : "+f"(mxa[u]) : "f"(__uint_as_float(s[i + 2 * u])), "f"(__uint_as_float(s[i + 2 * u + 1])));
      }
      // a half whose keys are all masked has mx = -inf: keep the running maximum finite (its probabilities are exp2(-inf) = 0)
      const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * scale_log2, -1e30f);
      if (j == 0) {
        m_run = mx;
      } else {
        const bool grow = mx > m_run + 8.0f;  // lazy rescale: P stays <= 2^8, exact after the final division
        if (__any_sync(0xffffffffu, grow)) {
          if (!pv_ok) { mbar_wait(pv_done(st), (j - 1) & 1); pv_ok = 1; }  // O of this stream must be quiescent
          tc_fence_after();
          const float m_new = grow ? mx : m_run;
          const float alpha = ex2_approx(m_run - m_new);
          l_run *= alpha;
          m_run = m_new;
          uint32_t o[32];
          tmem_ld_32x32b_x32(o_addr, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(o_addr, o);
          tmem_wait_st();
        }
      }
      float sums[4] = {0.f, 0.f, 0.f, 0.f};
      // (IPACK == 2, measured: 1.931 vs 1.934 ms at N = 16384, H = 24 — no gain, and rms error x1.6-2 with a -2e-3 relative
      //  bias on peaked rows: the bf16 pack on the XU pipe is NOT what limits the kernel.  Kept only as an experiment.)
      // IPACK == 2: P is TRUNCATED to bf16 by one byte permute (ALU pipe; F2FP shares the XU pipe with MUFU.EX2).  Truncation
      // loses on average 0.7213 * 2^-8 of a value (log-uniform mantissa), so the exponent is offset by log2(1 + that): the
      // row sum then carries the same factor as the mean of the truncated P, and O / l is unbiased; lse is corrected below.
      const float neg_m = IPACK == 2 ? kF4TruncBiasLog2 - m_run : -m_run;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // packed fp32 pairs: one FFMA2 (scale and subtract the running maximum) and one FADD2 (row sums) per TWO scores —
          // 5 instead of 7 issue slots per pair; the kernel is issue-bound, not MUFU-bound (moving exponentials to the FMA
          // pipes made it slower at every fraction)
          float x0, x1;
          ffma2(x0, x1, __uint_as_float(s[c * 32 + 2 * i]), __uint_as_float(s[c * 32 + 2 * i + 1]), scale_log2, neg_m);
          const float p0 = ex2_approx(x0);
          const float p1 = (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == 0) ? ex2_poly(x1) : ex2_approx(x1);
          if (i & 1) fadd2(sums[2], sums[3], p0, p1);
          else fadd2(sums[0], sums[1], p0, p1);
          if constexpr (IPACK == 2) {
            pk[i] = __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
          } else if constexpr (IPACK == 1) {  // p >= 0 and finite: no carry into the sign, ties (exact .5 ulp) round up instead of to even
            pk[i] = __byte_perm(__float_as_uint(p0) + 0x8000u, __float_as_uint(p1) + 0x8000u, 0x7632);
          } else {
            __nv_bfloat162 pp = __floats2bfloat162_rn(p0, p1);
            pk[i] = *reinterpret_cast<uint32_t*>(&pp);
          }
        }
        if (c == 0) {
          // PV(j-1) must have consumed P(j-1) before the buffer is rewritten.  Waiting HERE — after the first half of the
          // exponentials, not before them — hides the issuer's reaction time (p_full -> PV -> commit took longer than this
          // warp's load / max phase: 8 % of all stall samples sat on this barrier)
          if (!pv_ok) mbar_wait(pv_done(st), (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st_32x32b_x16(p_addr + c * 16, pk);
      }
      l_run += (sums[0] + sums[1]) + (sums[2] + sums[3]);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(st));
    }
    // ---- epilogue: this stream's (m, l, O) -> exact merge of the two key halves of the tile ----
    mbar_wait(pv_done(st), (n_tiles - 1) & 1);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_32x32b_x32(o_addr, o);
    tmem_wait_ld();
    float* exo = ex + (size_t)(t * F4_BQ + row) * F4_EXO_STRIDE;
    float* exm = ex + 2 * F4_BQ * F4_EXO_STRIDE + (t * F4_BQ + row) * 2;
    if (hh == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) exo[i] = __uint_as_float(o[i]);
      exm[0] = m_run; exm[1] = l_run;
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");  // the 8 warps of query tile t
    if (hh == 0) {
      const float m1 = exm[0], l1 = exm[1];
      const float M = fmaxf(m_run, m1);
      const float a0 = ex2_approx(m_run - M), a1 = ex2_approx(m1 - M);
      const float l = l_run * a0 + l1 * a1;
      const float inv_l = 1.0f / l;
      const float w0 = a0 * inv_l, w1 = a1 * inv_l;
      const int64_t n = q0 + t * F4_BQ + row;
      if (n < N) {
        const int64_t o_off = ((b * N + n) * H + h_idx) * F4_DH;
        if (lse) lse[(b * H + h_idx) * N + n] = M + log2f(l) - (IPACK == 2 ? kF4TruncBiasLog2 : 0.f);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(o[g * 8 + i]) * w0 + exo[g * 8 + i] * w1;
          if (out32) Vec8<float>::store(out32 + o_off + g * 8, y);
          else Vec8<bf16>::store(out + o_off + g * 8, y);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// q [B,N,H,32] / k, v [B,Nk,H,32] with row pitches ldq / ldkv (elements); same contract as the dense path of attn_tc_launch
int attn_tc4_launch(const void* q, const void* k, const void* v, int B, int64_t N, int64_t Nk, const int32_t* kv_len, int H,
                    void* out, float* lse, cudaStream_t st, float* out32, int64_t ldq, int64_t ldkv) {
  const uint64_t d = (uint64_t)H * F4_DH;
  CUtensorMap tmQ, tmK, tmV;
  LCASR_TRY(make_tmap_2d_bf16(&tmQ, q, (uint64_t)B * N, d, (uint64_t)ldq * 2, 2 * F4_BQ, F4_DH, CU_TENSOR_MAP_SWIZZLE_64B));
  LCASR_TRY(make_tmap_2d_bf16(&tmK, k, (uint64_t)B * Nk, d, (uint64_t)ldkv * 2, F4_BK, F4_DH, CU_TENSOR_MAP_SWIZZLE_64B));
  LCASR_TRY(make_tmap_2d_bf16(&tmV, v, (uint64_t)B * Nk, d, (uint64_t)ldkv * 2, F4_BK, F4_DH, CU_TENSOR_MAP_SWIZZLE_64B));
  static const int poly = getenv("LCASR_ATTN_POLY") ? atoi(getenv("LCASR_ATTN_POLY")) : kF4DefaultPoly;
  static const int ipack = getenv("LCASR_ATTN_IPACK") ? atoi(getenv("LCASR_ATTN_IPACK")) : kF4DefaultIpack;
  dim3 grid((unsigned)ceil_div(N, 2 * F4_BQ), (unsigned)H, (unsigned)B);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)F4_DH);
#define LCASR_F4(P, I)                                                                                                          \
  {                                                                                                                            \
    static PerDeviceFlag attr_set;                                                                                             \
    int attr_dev = 0;                                                                                                          \
    if (attr_set.needs_set(&attr_dev)) {                                                                                       \
      LCASR_CUDA(cudaFuncSetAttribute(attn_tc4_kernel<P, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM_BYTES));     \
      attr_set.mark(attr_dev);                                                                                                 \
    }                                                                                                                          \
    attn_tc4_kernel<P, I><<<grid, F4_THREADS, F4_SMEM_BYTES, st>>>(tmQ, tmK, tmV, N, Nk, kv_len, H, scale_log2, (bf16*)out, lse, out32); \
  }
  if (ipack == 2) {
    LCASR_F4(0, 2)
  } else if (ipack) {
    LCASR_F4(0, 1)
  } else {
    switch (poly) {
      case 2: LCASR_F4(2, 0) break;
      case 4: LCASR_F4(4, 0) break;
      default: LCASR_F4(0, 0) break;
    }
  }
#undef LCASR_F4
  LCASR_LAUNCH_CHECK();
  return 0;
}

}  // namespace lcasr
