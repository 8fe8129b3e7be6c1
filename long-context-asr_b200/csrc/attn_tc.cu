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

int attn_tc_available() { return 1; }

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
  static bool attr_set = false;
  if (!attr_set) {
    LCASR_CUDA(cudaFuncSetAttribute(attn_tc_kernel<DH, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(N, FA_BQ), (unsigned)H, (unsigned)B);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)DH);
  attn_tc_kernel<DH, VT><<<grid, FA_THREADS, Cfg::SMEM_BYTES, st>>>(tmQ, tmK, tmV, N, H, scale_log2, (bf16*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

int attn_tc_launch(const void* q, const void* k, const void* v, int B, int64_t N, int H, int Dh, int v_transposed,
                   int64_t Npad, void* out, cudaStream_t st) {
  LCASR_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)out & 15) == 0,
                  "attention(tcgen05): q, k, v, out must be 16-byte aligned");
  LCASR_CHECK_ARG((int64_t)B * N < ((int64_t)1 << 31), "attention(tcgen05): B*N too large");
  LCASR_CHECK_ARG(!v_transposed || Npad % 8 == 0, "attention(tcgen05): Npad must be a multiple of 8");
  LCASR_CHECK_ARG(H <= 65535 && B <= 65535, "attention(tcgen05): too many heads / batch entries");
#define LCASR_FA(DHV)                                                                                             \
  case DHV:                                                                                                       \
    return v_transposed ? launch_attn_tc<DHV, true>(q, k, v, B, N, H, Npad, out, st)                              \
                        : launch_attn_tc<DHV, false>(q, k, v, B, N, H, Npad, out, st);
  switch (Dh) {
    LCASR_FA(32) LCASR_FA(64) LCASR_FA(128)
    default:
      return set_error(LCASR_E_UNSUPPORTED, "attention(tcgen05): head_dim=%d not in {32,64,128}", Dh);
  }
#undef LCASR_FA
}

}  // namespace lcasr
