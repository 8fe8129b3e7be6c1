// CTC loss: torch.nn.CTCLoss(blank=V, reduction=...) as exp/train.py:104,249 calls it
// (ATen LossCTC: Graves et al. alpha/beta recursion in the log domain).
//
// Forward (alpha): one CTA per sample; the 2S+1 extended-label states are strided over the
// threads, alpha(t-1)/alpha(t) are double-buffered in shared memory (2*(2S+1)*4 bytes: 27001 states
// of a 1-hour recording = 216 KB, inside the 227 KB a CTA may own), one __syncthreads per frame;
// the log-prob gathers of frame t+1 are prefetched into registers while frame t is computed.
// Latency-bound (serial in time); algorithmic HBM bytes = N*V*4 only if every class were read —
// in fact it touches (S+1) values per frame.
// Backward: beta recursion run the same way backwards in time writing log(alpha*beta) in place,
// then a collect kernel (one CTA per frame) scatters into the class axis.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace lcasr {

constexpr int kCtcMaxSPT = 32;

// log(e^a + e^b + e^c).  The correction term log(sum of exp(x - max)) lies in [0, log 3]; the state values themselves are
// O(10^2..10^4) in magnitude, so their fp32 ulp (6e-5 at 1000) is orders above the 2^-22 error of ex2/lg2.approx:
// the fast intrinsics cost nothing in accuracy and take the recursion from ~70 to ~20 instructions per state
// (the frame loop is issue-bound on ONE SM per sample: 0.84 -> 0.66 us per frame at 1229 states together with the
// pre-gathered inputs and the hoisted index arithmetic below).
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(fmaxf(a, b), c);
  if (m == -INFINITY) return -INFINITY;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

// G[b, t, s] = log_probs[b, t, ext_label(s)] for every frame, into BOTH lattices (fully parallel pre-pass)
__global__ void __launch_bounds__(256) ctc_gather_kernel(const float* __restrict__ log_probs, int64_t N, int V,
                                                         const int64_t* __restrict__ targets, int64_t S_max,
                                                         const int32_t* __restrict__ input_lengths,
                                                         const int64_t* __restrict__ target_lengths, int blank,
                                                         float* __restrict__ ga, float* __restrict__ gb) {
  const int b = blockIdx.y;
  const int64_t t = blockIdx.x;
  const int64_t T = input_lengths ? min((int64_t)input_lengths[b], N) : N;
  if (t >= T) return;
  const int Lp = (int)(2 * target_lengths[b] + 1), Lp_max = (int)(2 * S_max + 1);
  const float* row = log_probs + ((int64_t)b * N + t) * V;
  const int64_t* tgt = targets + (int64_t)b * S_max;
  const int64_t o = ((int64_t)b * N + t) * Lp_max;
  const float bl = row[blank];
  for (int s = threadIdx.x; s < Lp; s += 256) {
    const float v = (s & 1) ? row[(int)tgt[(s - 1) >> 1]] : bl;
    ga[o + s] = v;
    gb[o + s] = v;
  }
}

// direction = +1: alpha (forward in time, labels as given); direction = -1: beta (time and label
// order reversed — the recursion is symmetric). `store` (may be NULL) receives the per-frame state
// vector in ORIGINAL (t, s) coordinates; for beta it is ADDED to what is there (alpha+beta).
template <int SPT>
__global__ void __launch_bounds__(1024) ctc_recursion_kernel(const float* __restrict__ log_probs, int64_t N, int V,
                                                             const int64_t* __restrict__ targets, int64_t S_max,
                                                             const int32_t* __restrict__ input_lengths,
                                                             const int64_t* __restrict__ target_lengths, int blank,
                                                             int direction, float* __restrict__ nll,
                                                             float* __restrict__ store, float* __restrict__ store_beta,
                                                             int batch, int pregathered) {
  // pregathered (direction == 0 only): both lattices were pre-filled with G[t, s] = log_probs[t, ext_label(s)] by
  // ctc_gather_kernel; the recursion then reads its per-frame inputs as one coalesced row (a few 128-byte lines)
  // instead of ~(S+1) scattered sectors per frame, which is what bounded the frame time (LSU: ~1 sector per clock),
  // and overwrites row t with alpha / beta after it has been consumed (reads run kPF frames ahead of the writes).
  // direction == 0: BOTH recursions in one launch (grid = 2*batch CTAs): CTA b < batch runs alpha into `store`,
  // CTA batch+b runs beta into `store_beta` (plain stores: the two are independent and run concurrently)
  extern __shared__ float sm_alpha[];  // [2][Lp_pad]
  const bool both = direction == 0;
  if (both) {
    direction = (int)blockIdx.x < batch ? +1 : -1;
    if (direction < 0) { store = store_beta; nll = nullptr; }
  }
  const bool beta_adds = !both;  // legacy two-launch form: beta is added onto the alpha already in `store`
  const int b = both ? (int)blockIdx.x % batch : (int)blockIdx.x;
  const int NT = blockDim.x;
  const int tid = threadIdx.x;
  const int64_t T = input_lengths ? min((int64_t)input_lengths[b], N) : N;
  const int64_t S = target_lengths[b];
  const int Lp = (int)(2 * S + 1);
  const int Lp_max = (int)(2 * S_max + 1);
  const int Lp_pad = SPT * NT;
  float* cur = sm_alpha;
  float* nxt = sm_alpha + Lp_pad;
  const float* lp = log_probs + (int64_t)b * N * V;
  const int64_t* tgt = targets + (int64_t)b * S_max;
  float* st = store ? store + (int64_t)b * N * Lp_max : nullptr;

  if (T <= 0 || (T < S)) {  // ATen: impossible alignment -> inf loss
    if (tid == 0 && nll) nll[b] = INFINITY;
    return;
  }

  // label of state s in the (possibly reversed) extended sequence, and the "skip" permission
  int lab[SPT];
  bool skip[SPT];
#pragma unroll
  for (int j = 0; j < SPT; ++j) {
    int s = tid + j * NT;
    lab[j] = blank;
    skip[j] = false;
    if (s < Lp && (s & 1)) {
      int64_t li = (s - 1) >> 1;                       // index in the (reversed) label sequence
      int64_t oi = direction > 0 ? li : (S - 1 - li);  // index in the original targets
      lab[j] = (int)tgt[oi];
      if (li >= 1) {
        int64_t po = direction > 0 ? oi - 1 : oi + 1;
        skip[j] = tgt[po] != tgt[oi];
      }
    }
  }
  auto frame = [&](int64_t step) -> int64_t { return direction > 0 ? step : (T - 1 - step); };

  // log-prob gathers run kPF frames AHEAD of their use (a register ring indexed at compile time by unrolling the
  // frame loop kPF times): one frame of recursion (~0.2 us) is far shorter than the ~0.6 us global-memory latency of
  // a gather, which a one-frame look-ahead exposed on every step (0.82 us/frame measured before).
  constexpr int kPF = SPT <= 4 ? 4 : (SPT <= 8 ? 2 : 1);  // the ring costs kPF*SPT registers (1024-thread CTAs: 64 in all)
  float lpq[kPF][SPT];
  {
    const float* row = lp + frame(0) * V;
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      int s = tid + j * NT;
      float a = -INFINITY;
      if (s == 0) a = row[blank];
      else if (s == 1 && Lp > 1) a = row[lab[j]];
      cur[s] = a;
      if (st && s < Lp) {
        int so = direction > 0 ? s : (Lp - 1 - s);
        float* p = st + frame(0) * Lp_max + so;
        *p = (direction > 0 || !beta_adds) ? a : (*p + a);
      }
    }
#pragma unroll
    for (int u = 0; u < kPF; ++u) {  // frames 1 .. kPF
      if (1 + u < T) {
        const float* rowu = lp + frame(1 + u) * V;
        const float* gu = st + frame(1 + u) * Lp_max;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
          const int s = tid + j * NT;
          if (pregathered) lpq[u][j] = s < Lp ? gu[direction > 0 ? s : (Lp - 1 - s)] : 0.f;
          else lpq[u][j] = rowu[lab[j]];
        }
      }
    }
  }
  __syncthreads();
  // The frame loop is issue-bound on one SM (ncu: 64 % issue-active, ~190 warp instructions per frame before this
  // form): everything that does not depend on the frame is hoisted — per-state validity / output index, and the
  // row pointers advance by a constant stride instead of being rebuilt from 64-bit products every frame.
  const int64_t dstride = (int64_t)direction * Lp_max, lstride = (int64_t)direction * V;
  bool live[SPT];
  int so_j[SPT];
#pragma unroll
  for (int j = 0; j < SPT; ++j) {
    const int s = tid + j * NT;
    live[j] = s < Lp;
    so_j[j] = direction > 0 ? s : (Lp - 1 - s);
    if (!live[j]) so_j[j] = 0;
  }
  float* strow = st ? st + frame(1) * Lp_max : nullptr;             // row of the frame being computed
  const float* pf_g = st ? st + frame(1 + kPF) * Lp_max : nullptr;   // rows kPF frames ahead (pre-gathered inputs)
  const float* pf_l = lp + frame(1 + kPF) * V;                        // (gather form)
  for (int64_t step0 = 1; step0 < T; step0 += kPF) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      const int64_t step = step0 + u;
      if (step >= T) break;  // block-uniform
      float lpv[SPT];
#pragma unroll
      for (int j = 0; j < SPT; ++j) lpv[j] = lpq[u][j];
      if (step + kPF < T) {  // refill this ring slot with the frame kPF steps ahead
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
          if (pregathered) lpq[u][j] = live[j] ? pf_g[so_j[j]] : 0.f;
          else lpq[u][j] = pf_l[lab[j]];
        }
      }
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        const int s = tid + j * NT;
        if (live[j]) {
          const float a0 = cur[s];
          const float a1 = s >= 1 ? cur[s - 1] : -INFINITY;
          const float a2 = skip[j] ? cur[s - 2] : -INFINITY;
          const float a = lse3(a0, a1, a2) + lpv[j];
          nxt[s] = a;
          if (strow) strow[so_j[j]] = (direction > 0 || !beta_adds) ? a : (strow[so_j[j]] + a);
        }
      }
      __syncthreads();
      float* tmp = cur; cur = nxt; nxt = tmp;
      if (strow) { strow += dstride; pf_g += dstride; }
      pf_l += lstride;
    }
  }
  if (tid == 0 && nll) {
    float l1 = cur[Lp - 1];
    float l2 = Lp > 1 ? cur[Lp - 2] : -INFINITY;
    float m = fmaxf(l1, l2);
    nll[b] = m == -INFINITY ? INFINITY : -(m + logf(expf(l1 - m) + expf(l2 - m)));
  }
}

// ---- long sequences: one thread-block CLUSTER per sample -------------------------------------
// A 1-hour recording has 45000 frames x 27001 states; on one SM that is ~26 us per frame (1.16 s,
// the same as ATen's kernel).  Here the states are split over the CTAs of a cluster (<= 8 SMs, 1024
// threads each, SPT <= 4 contiguous states per thread); alpha(t-1)/alpha(t) stay double-buffered in
// each CTA's shared memory, the two boundary states come from the left neighbour through distributed
// shared memory, and one cluster barrier per frame replaces __syncthreads.
namespace cg = cooperative_groups;

template <int SPT>
__global__ void __launch_bounds__(1024) ctc_cluster_kernel(const float* __restrict__ log_probs, int64_t N, int V,
                                                           const int64_t* __restrict__ targets, int64_t S_max,
                                                           const int32_t* __restrict__ input_lengths,
                                                           const int64_t* __restrict__ target_lengths, int blank,
                                                           int direction, float* __restrict__ nll,
                                                           float* __restrict__ store) {
  extern __shared__ float sm_alpha[];  // [2][CH]
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  const int b = blockIdx.x / csize;
  const int NT = blockDim.x, tid = threadIdx.x;
  const int CH = NT * SPT;
  const int64_t T = input_lengths ? min((int64_t)input_lengths[b], N) : N;
  const int64_t S = target_lengths[b];
  const int Lp = (int)(2 * S + 1), Lp_max = (int)(2 * S_max + 1);
  float* cur = sm_alpha;
  float* nxt = sm_alpha + CH;
  const float* lp = log_probs + (int64_t)b * N * V;
  const int64_t* tgt = targets + (int64_t)b * S_max;
  float* st = store ? store + (int64_t)b * N * Lp_max : nullptr;
  if (T <= 0 || T < S) {  // uniform over the cluster
    if (crank == 0 && tid == 0 && nll) nll[b] = INFINITY;
    return;
  }
  const int loc0 = tid * SPT, s0 = crank * CH + loc0;
  int lab[SPT];
  bool skip[SPT];
#pragma unroll
  for (int j = 0; j < SPT; ++j) {
    const int s = s0 + j;
    lab[j] = blank;
    skip[j] = false;
    if (s < Lp && (s & 1)) {
      const int64_t li = (s - 1) >> 1;
      const int64_t oi = direction > 0 ? li : (S - 1 - li);
      lab[j] = (int)tgt[oi];
      if (li >= 1) skip[j] = tgt[direction > 0 ? oi - 1 : oi + 1] != tgt[oi];
    }
  }
  auto frame = [&](int64_t step) -> int64_t { return direction > 0 ? step : (T - 1 - step); };
  float lpv[SPT], lpn[SPT];
  {
    const float* row = lp + frame(0) * V;
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const int s = s0 + j;
      float a = -INFINITY;
      if (s == 0) a = row[blank];
      else if (s == 1 && Lp > 1) a = row[lab[j]];
      cur[loc0 + j] = a;
      if (st && s < Lp) {
        float* p = st + frame(0) * Lp_max + (direction > 0 ? s : (Lp - 1 - s));
        *p = direction > 0 ? a : (*p + a);
      }
    }
    if (T > 1) {
      const float* row1 = lp + frame(1) * V;
#pragma unroll
      for (int j = 0; j < SPT; ++j) lpv[j] = row1[lab[j]];
    }
  }
  // the left neighbour's two buffers (distributed shared memory)
  const float* left_cur = crank > 0 ? cluster.map_shared_rank(cur, crank - 1) : nullptr;
  const float* left_nxt = crank > 0 ? cluster.map_shared_rank(nxt, crank - 1) : nullptr;
  cluster.sync();
  for (int64_t step = 1; step < T; ++step) {
    if (step + 1 < T) {
      const float* rown = lp + frame(step + 1) * V;
#pragma unroll
      for (int j = 0; j < SPT; ++j) lpn[j] = rown[lab[j]];
    }
    float* strow = st ? st + frame(step) * Lp_max : nullptr;
    float prev1, prev2;  // alpha(t-1) of states s0-1, s0-2
    if (loc0 >= 2) { prev1 = cur[loc0 - 1]; prev2 = cur[loc0 - 2]; }
    else if (crank > 0) { prev1 = left_cur[CH - 1]; prev2 = left_cur[CH - 2]; }  // only thread 0 of a CTA (SPT >= 2) ...
    else { prev1 = -INFINITY; prev2 = -INFINITY; }
    if (SPT == 1 && loc0 == 1) { prev1 = cur[0]; prev2 = crank > 0 ? left_cur[CH - 1] : -INFINITY; }  // ... or threads 0,1 (SPT == 1)
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const int s = s0 + j;
      const float a0 = cur[loc0 + j];
      if (s < Lp) {
        const float a2 = skip[j] ? prev2 : -INFINITY;
        const float m = fmaxf(fmaxf(a0, prev1), a2);
        float a = -INFINITY;
        if (m != -INFINITY) a = m + __logf(__expf(a0 - m) + __expf(prev1 - m) + __expf(a2 - m)) + lpv[j];
        nxt[loc0 + j] = a;
        if (strow) {
          const int so = direction > 0 ? s : (Lp - 1 - s);
          strow[so] = direction > 0 ? a : (strow[so] + a);
        }
      }
      prev2 = prev1;
      prev1 = a0;
    }
    cluster.sync();
    { float* tmp = cur; cur = nxt; nxt = tmp; }
    { const float* tmp = left_cur; left_cur = left_nxt; left_nxt = tmp; }
#pragma unroll
    for (int j = 0; j < SPT; ++j) lpv[j] = lpn[j];
  }
  if (nll) {  // the thread that owns state Lp-1 finishes the sample
    const int sl = Lp - 1;
    if (sl >= s0 && sl < s0 + SPT) {
      const int loc = sl - crank * CH;
      const float l1 = cur[loc];
      float l2 = -INFINITY;
      if (Lp > 1) l2 = loc >= 1 ? cur[loc - 1] : left_cur[CH - 1];
      const float m = fmaxf(l1, l2);
      nll[b] = m == -INFINITY ? INFINITY : -(m + logf(expf(l1 - m) + expf(l2 - m)));
    }
  }
  cluster.sync();  // nobody exits while a neighbour may still read its shared memory
}

template <int SPT>
static int launch_cluster(const float* lp, int B, int64_t N, int V, const int64_t* tg, int64_t S_max, const int32_t* il,
                          const int64_t* tl, int blank, int dir, float* nll, float* store, int nt, int csize,
                          cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(B * csize));
  cfg.blockDim = dim3((unsigned)nt);
  cfg.dynamicSmemBytes = (size_t)2 * SPT * nt * sizeof(float);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LCASR_CUDA(cudaLaunchKernelEx(&cfg, ctc_cluster_kernel<SPT>, lp, N, V, tg, S_max, il, tl, blank, dir, nll, store));
  count_launch();
  return 0;
}

// ---- time-skewed wavefront over the whole GPU ---------------------------------------------------
// The recursion is serial in time, but state s at frame t only needs states s, s-1, s-2 at frame t-1.  The extended
// states of one lattice are cut into G contiguous chunks, one persistent CTA each (cooperative launch: all resident);
// chunk g runs about two batches of kWfTB frames BEHIND chunk g-1 and receives the alpha values of its left neighbour's
// last two states through a global hand-off buffer.  No grid- or cluster-wide barrier remains: the per-frame cost is one
// CTA's dependent chain (shared-memory neighbours, one branch-free lse3, __syncthreads over <= 1024 threads) instead of
// 2.1 us for a barrier across 8 SMs.
// Hand-off without flags or fences: the record of (chunk, step) is ONE aligned 16-byte store
//   {alpha[last-1], alpha[last], step + 1, launch epoch}
// written by the thread that owns the chunk's last state (it reads its left neighbour's previous value anyway), and the
// consumer checks the tag of every record it loads (L2 loads, 16 records = one 256-byte request by half a warp); stale
// data of an earlier launch carries another epoch.  Records of the NEXT batch are fetched while the current one is being
// computed, so in steady state nobody waits.  Log-prob gathers run one batch ahead in registers.
// Same arithmetic per state as the kernels above (bit-identical results).
constexpr int kWfTBDefault = 16;

// branch-free lse3: an all -inf input gives exp(-inf) = 0 -> log(0) = -inf, as the branchy form returns
// ex2 / lg2 in their .ftz forms: __expf / __logf wrap every MUFU in a compare and two predicated multiplies for denormal
// results / arguments (a third of the recursion's instructions, all on its dependent chain).  Here they cannot matter: the
// largest term is exp(0) = 1, a denormal term added to >= 1 does not change the fp32 sum, and lg2's argument lies in [1, 3]
// (or is exactly 0) — the results stay bit-identical to lse3 above (tests/test_gpu_ops.py checks exactly that).
__device__ __forceinline__ float lse3_nb(float a, float b, float c) {
  const float m = fmaxf(fmaxf(fmaxf(a, b), c), -1e30f);
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  auto e2 = [](float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
  const float sum = e2((a - m) * kLog2e) + e2((b - m) * kLog2e) + e2((c - m) * kLog2e);
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(sum));
  return m + l * kLn2;
}

template <int kWfTB, bool STORE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
ctc_wavefront_kernel(const float* __restrict__ log_probs, int64_t N, int V, const int64_t* __restrict__ targets, int64_t S_max,
                     const int32_t* __restrict__ input_lengths, const int64_t* __restrict__ target_lengths, int blank,
                     int direction, int G, int batch, float* __restrict__ nll, float* __restrict__ store,
                     float* __restrict__ store_beta, int pregathered, float4* __restrict__ bnd, unsigned epoch) {
  // grid = units * G CTAs, unit = (sample) for direction +-1, (sample, direction) for direction 0 (alpha units first).
  // bnd [units*G][N]: one record per (chunk, step).
  static_assert(kWfTB <= 32, "one lane per record of a batch");
  extern __shared__ float wf_sm[];  // [2][2 + CH] state vectors with a 2-slot left halo, then bl [2][kWfTB][2]
  const int NT = blockDim.x, tid = threadIdx.x, CH = NT;
  const int unit = (int)blockIdx.x / G, g = (int)blockIdx.x % G;
  const bool both = direction == 0;
  int b = unit;
  if (both) {
    direction = unit < batch ? +1 : -1;
    b = unit % batch;
    if (direction < 0) { store = store_beta; nll = nullptr; }
  }
  const bool beta_adds = !both;
  const int64_t T = input_lengths ? min((int64_t)input_lengths[b], N) : N;
  const int64_t S = target_lengths[b];
  const int Lp = (int)(2 * S + 1), Lp_max = (int)(2 * S_max + 1);
  if (T <= 0 || T < S) {  // uniform over the unit: nobody waits for anybody
    if (g == 0 && tid == 0 && nll) nll[b] = INFINITY;
    return;
  }
  const int s0 = g * CH;
  if (s0 >= Lp) return;  // dead chunk: everything to its right is dead as well
  float* cur = wf_sm + 2;              // cur[-2], cur[-1] = left neighbour's last two states
  float* nxt = wf_sm + 2 + (CH + 2);
  float* bl = wf_sm + 2 * (CH + 2);    // [2][kWfTB][2]
  const float* lp = log_probs + (int64_t)b * N * V;
  const int64_t* tgt = targets + (int64_t)b * S_max;
  float* st = (STORE && store) ? store + (int64_t)b * N * Lp_max : nullptr;
  const int s = s0 + tid;
  const bool live = s < Lp;
  int lab = blank;
  bool skip = false;
  if (live && (s & 1)) {
    const int64_t li = (s - 1) >> 1;
    const int64_t oi = direction > 0 ? li : (S - 1 - li);
    lab = (int)tgt[oi];
    if (li >= 1) skip = tgt[direction > 0 ? oi - 1 : oi + 1] != tgt[oi];
  }
  const int so = live ? (direction > 0 ? s : (Lp - 1 - s)) : 0;
  auto frame = [&](int64_t step) -> int64_t { return direction > 0 ? step : (T - 1 - step); };
  auto fetch = [&](int64_t step) -> float {  // the log-prob this state consumes at `step`
    if (!live || step >= T) return 0.f;
    if (STORE && pregathered) return __ldcg(st + frame(step) * Lp_max + so);
    return __ldg(lp + frame(step) * V + lab);
  };
  float4* my_bnd = bnd + (int64_t)blockIdx.x * N;
  const float4* left_bnd = g > 0 ? bnd + (int64_t)(blockIdx.x - 1) * N : nullptr;
  const bool pub = tid == CH - 1;  // owner of the chunk's last state: publishes one record per step
  const float epoch_f = __uint_as_float(epoch);

  // warp 0, lanes < cnt: records of steps [first, first + cnt) of the left chunk -> bl[slot]; true once all tags match
  auto try_fetch = [&](int64_t first, int cnt, int slot) -> bool {
    bool ok = true;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < cnt) {
      r = __ldcg(left_bnd + first + tid);
      ok = __float_as_uint(r.z) == (unsigned)(first + tid + 1) && __float_as_uint(r.w) == epoch;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (ok && tid < cnt) { bl[(slot * kWfTB + tid) * 2] = r.x; bl[(slot * kWfTB + tid) * 2 + 1] = r.y; }
    __syncwarp();
    return ok;
  };
  auto fetch_blocking = [&](int64_t first, int cnt, int slot) {  // warp 0 only; deadlock guard: trap instead of hanging the GPU
    unsigned spins = 0;
    uint64_t t_first = 0;
    while (!try_fetch(first, cnt, slot)) {
      __nanosleep(64);
      if ((++spins & 0x3FFF) == 0) {
        uint64_t now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t_first == 0) t_first = now;
        else if (now - t_first > 4000000000ull) {
          if (tid == 0) printf("lcasr_b200: ctc wavefront stalled (cta %d waits for steps %lld..)\n", (int)blockIdx.x, (long long)first);
          asm volatile("trap;");
        }
      }
    }
  };

  // ---- step 0 ----
  float lpa[kWfTB], lpb[kWfTB];
  {
    float a = -INFINITY;
    if (live) {
      const float* row = lp + frame(0) * V;
      if (s == 0) a = row[blank];
      else if (s == 1) a = row[lab];
    }
    cur[tid] = a;
    if (tid < 2) { cur[tid - 2] = -INFINITY; nxt[tid - 2] = -INFINITY; }
    if (st && live) {
      float* p = st + frame(0) * Lp_max + so;
      *p = (direction > 0 || !beta_adds) ? a : (*p + a);
    }
#pragma unroll
    for (int u = 0; u < kWfTB; ++u) lpa[u] = fetch(1 + u);
  }
  __syncthreads();
  // records of steps t0-1 .. t1-2 serve batch [t0, t1); batch 0 = [1, ..): fetched before anything else
  int slot = 0;
  bool next_ready = false;
  if (g > 0 && tid < 32 && T > 1) fetch_blocking(0, (int)min((int64_t)kWfTB, T - 1), 0);

  // one batch of up to kWfTB steps [t0, t1): inputs from `lq`, next batch's inputs prefetched into `lnext`
  auto run_batch = [&](int64_t t0, float (&lq)[kWfTB], float (&lnext)[kWfTB]) {
    const int64_t t1 = min(T, t0 + kWfTB);
    const int64_t n1 = min(T, t1 + kWfTB);  // next batch = [t1, n1)
#pragma unroll
    for (int u = 0; u < kWfTB; ++u) lnext[u] = fetch(t0 + kWfTB + u);
    if (g > 0 && tid < 32) {
      __syncwarp();
      if (tid < 2) cur[tid - 2] = bl[(slot * kWfTB) * 2 + tid];  // boundary of step t0-1 (same warp wrote bl: no CTA barrier)
      next_ready = n1 > t1 ? try_fetch(t1 - 1, (int)(n1 - t1), slot ^ 1) : true;  // usually already there: the left chunk is ahead
    }
    // (no CTA barrier: the halo slots are written and read by lanes 0 / 1 of warp 0 only)
#pragma unroll
    for (int u = 0; u < kWfTB; ++u) {
      const int64_t step = t0 + u;
      if (step < t1) {  // block-uniform
        const float a0 = cur[tid], a1 = cur[tid - 1], a2 = skip ? cur[tid - 2] : -INFINITY;
        if (pub) {  // a0 / a1 are the chunk's last two states at step-1: their record (one 16-byte store)
          float4 rec = make_float4(a1, a0, __uint_as_float((unsigned)step), epoch_f);
          __stcg(my_bnd + (step - 1), rec);
        }
        const float a = lse3_nb(a0, a1, a2) + lq[u];
        nxt[tid] = a;
        if (g > 0 && tid < 2 && step + 1 < t1) nxt[tid - 2] = bl[(slot * kWfTB + u + 1) * 2 + tid];
        if (STORE && st && live) {
          float* p = st + frame(step) * Lp_max + so;
          *p = (direction > 0 || !beta_adds) ? a : (*p + a);
        }
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
      }
    }
    if (g > 0 && tid < 32) {
      if (!next_ready && n1 > t1) fetch_blocking(t1 - 1, (int)(n1 - t1), slot ^ 1);
    }
    slot ^= 1;
  };
  for (int64_t t0 = 1; t0 < T; t0 += 2 * kWfTB) {
    run_batch(t0, lpa, lpb);
    if (t0 + kWfTB < T) run_batch(t0 + kWfTB, lpb, lpa);
  }
  if (pub) {  // record of the final step
    float4 rec = make_float4(cur[tid - 1], cur[tid], __uint_as_float((unsigned)T), epoch_f);
    __stcg(my_bnd + (T - 1), rec);
  }
  if (nll) {  // the chunk that owns state Lp-1 finishes the sample
    const int sl = Lp - 1;
    if (sl >= s0 && sl < s0 + CH) {
      const int loc = sl - s0;
      if (loc == 0 && g > 0) {  // alpha_{T-1}[Lp-2] lives in the left chunk
        if (tid < 32) {
          fetch_blocking(T - 1, 1, 0);
          if (tid == 0) cur[-1] = bl[1];
        }
        __syncthreads();
      }
      if (tid == loc) {
        const float l1 = cur[loc];
        const float l2 = Lp > 1 ? cur[loc - 1] : -INFINITY;
        const float m = fmaxf(l1, l2);
        nll[b] = m == -INFINITY ? INFINITY : -(m + logf(expf(l1 - m) + expf(l2 - m)));
      }
    }
  }
}


// ---- wavefront, third form: ONE warp runs a chunk's recursion out of registers ----------------------------------------
// ctc_wavefront_kernel above keeps one state per thread: per frame three shared-memory loads, a store and a __syncthreads
// over six warps — ~520 cycles per frame at 183 states per chunk (12.3 ms for the 1-hour lattice), of which the arithmetic
// is a fraction.  Here lane l of the compute warp owns SPL consecutive states in registers: the neighbours s-1 / s-2 are
// registers of the same lane or two warp shuffles from lane l-1; no shared-memory round trip and no CTA barrier remain in
// the frame loop.  Seven helper warps gather the log-probs of the next batch of frames into a shared-memory ring
// (lane-major layout: conflict-free) behind a pair of mbarriers, so the compute warp never waits for a global load either.
// Hand-off between chunks: the same tagged 16-byte records, same arithmetic per state: bit-identical results.
constexpr int kWwHelpers = 7, kWwThreads = 32 * (1 + kWwHelpers);

template <int kTB, bool STORE, int SPL>
__global__ void __launch_bounds__(kWwThreads, 1)
ctc_wavefront_warp_kernel(const float* __restrict__ log_probs, int64_t N, int V, const int64_t* __restrict__ targets, int64_t S_max,
                          const int32_t* __restrict__ input_lengths, const int64_t* __restrict__ target_lengths, int blank,
                          int direction, int G, int batch, float* __restrict__ nll, float* __restrict__ store,
                          float* __restrict__ store_beta, int pregathered, float4* __restrict__ bnd, unsigned epoch) {
  using namespace ptx;
  static_assert(kTB <= 32, "one lane per record of a batch");
  constexpr int CH = 32 * SPL;
  __shared__ float lps[2][kTB][CH];        // log-prob ring: [slot][frame of the batch][j * 32 + lane]
  __shared__ float bl[2][kTB][2];          // records of the left chunk
  __shared__ float fin[CH + 2];            // final state vector (for the loss)
  __shared__ int labs[CH];
  __shared__ __align__(8) uint64_t bars[4];  // full[2], empty[2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int unit = (int)blockIdx.x / G, g = (int)blockIdx.x % G;
  const bool both = direction == 0;
  int b = unit;
  if (both) {
    direction = unit < batch ? +1 : -1;
    b = unit % batch;
    if (direction < 0) { store = store_beta; nll = nullptr; }
  }
  const bool beta_adds = !both;
  const int64_t T = input_lengths ? min((int64_t)input_lengths[b], N) : N;
  const int64_t S = target_lengths[b];
  const int Lp = (int)(2 * S + 1), Lp_max = (int)(2 * S_max + 1);
  if (T <= 0 || T < S) {
    if (g == 0 && tid == 0 && nll) nll[b] = INFINITY;
    return;
  }
  const int s0 = g * CH;
  if (s0 >= Lp) return;  // dead chunk
  const float* lp = log_probs + (int64_t)b * N * V;
  const int64_t* tgt = targets + (int64_t)b * S_max;
  float* st = (STORE && store) ? store + (int64_t)b * N * Lp_max : nullptr;
  auto frame = [&](int64_t step) -> int64_t { return direction > 0 ? step : (T - 1 - step); };
  auto label_of = [&](int s_) -> int {  // extended label of state s_ (odd states carry targets, in recursion order)
    if (!(s_ & 1) || s_ >= Lp) return blank;
    const int64_t li = (s_ - 1) >> 1;
    return (int)tgt[direction > 0 ? li : (S - 1 - li)];
  };
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int sl) { return bar0 + 8u * sl; };
  auto empty_bar = [&](int sl) { return bar0 + 8u * (2 + sl); };
  for (int i = tid; i < CH; i += kWwThreads) labs[i] = label_of(s0 + i);
  if (tid == 0) {
    mbar_init(full_bar(0), kWwHelpers); mbar_init(full_bar(1), kWwHelpers);
    mbar_init(empty_bar(0), 1); mbar_init(empty_bar(1), 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int64_t n_batches = (T - 1 + kTB - 1) / kTB;  // steps 1 .. T-1

  if (warp > 0) {
    // ------------------------- helper warps: gather the log-probs of batch k into slot k & 1 -------------------------
    const int ht = tid - 32;
    for (int64_t k = 0; k < n_batches; ++k) {
      const int slot = (int)(k & 1);
      mbar_wait(empty_bar(slot), (uint32_t)(((k >> 1) & 1) ^ 1));
      const int64_t t0 = 1 + k * kTB;
      // all loads of this thread first (independent: their latencies overlap), then the shared-memory stores
      constexpr int NIT = (kTB * CH + 32 * kWwHelpers - 1) / (32 * kWwHelpers);
      float v[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int idx = ht + it * 32 * kWwHelpers;
        const int u = idx / CH, sl = idx - u * CH;
        const int64_t step = t0 + u;
        const int s_ = s0 + sl;
        v[it] = 0.f;
        if (idx < kTB * CH && s_ < Lp && step < T) {
          if (STORE && pregathered) v[it] = __ldcg(st + frame(step) * Lp_max + (direction > 0 ? s_ : (Lp - 1 - s_)));
          else v[it] = __ldg(lp + frame(step) * V + labs[sl]);
        }
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int idx = ht + it * 32 * kWwHelpers;
        const int u = idx / CH, sl = idx - u * CH;
        if (idx < kTB * CH) lps[slot][u][(sl % SPL) * 32 + sl / SPL] = v[it];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(slot));
    }
    return;
  }

  // ------------------------- compute warp -------------------------
  float4* my_bnd = bnd + (int64_t)blockIdx.x * N;
  const float4* left_bnd = g > 0 ? bnd + (int64_t)(blockIdx.x - 1) * N : nullptr;
  const float epoch_f = __uint_as_float(epoch);
  auto try_fetch = [&](int64_t first, int cnt, int slot) -> bool {
    bool ok = true;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < cnt) {
      r = __ldcg(left_bnd + first + lane);
      ok = __float_as_uint(r.z) == (unsigned)(first + lane + 1) && __float_as_uint(r.w) == epoch;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (ok && lane < cnt) { bl[slot][lane][0] = r.x; bl[slot][lane][1] = r.y; }
    __syncwarp();
    return ok;
  };
  auto fetch_blocking = [&](int64_t first, int cnt, int slot) {
    unsigned spins = 0;
    uint64_t t_first = 0;
    while (!try_fetch(first, cnt, slot)) {
      __nanosleep(64);
      if ((++spins & 0x3FFF) == 0) {
        uint64_t now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t_first == 0) t_first = now;
        else if (now - t_first > 4000000000ull) {
          if (lane == 0) printf("lcasr_b200: ctc wavefront (warp form) stalled (cta %d waits for steps %lld..)\n", (int)blockIdx.x, (long long)first);
          asm volatile("trap;");
        }
      }
    }
  };

  float a[SPL];
  bool skip[SPL], live[SPL];
  int so[SPL];
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s_ = s0 + lane * SPL + j;
    live[j] = s_ < Lp;
    skip[j] = false;
    if (live[j] && (s_ & 1)) {
      const int64_t li = (s_ - 1) >> 1;
      const int64_t oi = direction > 0 ? li : (S - 1 - li);
      if (li >= 1) skip[j] = tgt[direction > 0 ? oi - 1 : oi + 1] != tgt[oi];
    }
    so[j] = live[j] ? (direction > 0 ? s_ : (Lp - 1 - s_)) : 0;
    // ---- step 0 ----
    float v = -INFINITY;
    if (live[j]) {
      const float* row = lp + frame(0) * V;
      if (s_ == 0) v = row[blank];
      else if (s_ == 1) v = row[labs[lane * SPL + j]];
    }
    a[j] = v;
    if (STORE && st && live[j]) {
      float* p = st + frame(0) * Lp_max + so[j];
      *p = (direction > 0 || !beta_adds) ? v : (*p + v);
    }
  }
  int rslot = 0;
  if (g > 0 && T > 1) fetch_blocking(0, (int)min((int64_t)kTB, T - 1), 0);
  for (int64_t k = 0; k < n_batches; ++k) {
    const int slot = (int)(k & 1);
    const int64_t t0 = 1 + k * kTB, t1 = min(T, t0 + kTB), n1 = min(T, t1 + kTB);
    bool next_ready = true;
    if (g > 0 && n1 > t1) next_ready = try_fetch(t1 - 1, (int)(n1 - t1), rslot ^ 1);  // usually there: the left chunk is ahead
    mbar_wait(full_bar(slot), (uint32_t)((k >> 1) & 1));
#pragma unroll
    for (int u = 0; u < kTB; ++u) {
      const int64_t step = t0 + u;
      if (step < t1) {  // warp-uniform
        float lq[SPL];
#pragma unroll
        for (int j = 0; j < SPL; ++j) lq[j] = lps[slot][u][j * 32 + lane];
        // the left neighbour's last two states at step-1: lane-1's registers, or the left chunk's record for lane 0
        float p1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1), p2 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
        if (lane == 0) {
          p1 = g > 0 ? bl[rslot][u][1] : -INFINITY;
          p2 = g > 0 ? bl[rslot][u][0] : -INFINITY;
        }
        if (lane == 31) __stcg(my_bnd + (step - 1), make_float4(a[SPL - 2], a[SPL - 1], __uint_as_float((unsigned)step), epoch_f));
        float nw[SPL];
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
          const float x1 = j >= 1 ? a[j - 1] : p1;
          const float x2 = j >= 2 ? a[j - 2] : (j == 1 ? p1 : p2);
          nw[j] = lse3_nb(a[j], x1, skip[j] ? x2 : -INFINITY) + lq[j];
        }
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
          a[j] = nw[j];
          if (STORE && st && live[j]) {
            float* p = st + frame(step) * Lp_max + so[j];
            *p = (direction > 0 || !beta_adds) ? nw[j] : (*p + nw[j]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(slot));
    if (g > 0 && !next_ready && n1 > t1) fetch_blocking(t1 - 1, (int)(n1 - t1), rslot ^ 1);
    rslot ^= 1;
  }
  if (lane == 31) __stcg(my_bnd + (T - 1), make_float4(a[SPL - 2], a[SPL - 1], __uint_as_float((unsigned)T), epoch_f));
  if (nll) {  // the chunk that owns state Lp-1 finishes the sample
    const int sl = Lp - 1;
    if (sl >= s0 && sl < s0 + CH) {
      const int loc = sl - s0;
#pragma unroll
      for (int j = 0; j < SPL; ++j) fin[2 + lane * SPL + j] = a[j];
      if (loc == 0 && g > 0) {  // alpha_{T-1}[Lp-2] lives in the left chunk
        fetch_blocking(T - 1, 1, 0);
        if (lane == 0) fin[1] = bl[0][0][1];
      }
      __syncwarp();
      if (lane == 0) {
        const float l1 = fin[2 + loc];
        const float l2 = Lp > 1 ? fin[2 + loc - 1] : -INFINITY;
        const float m = fmaxf(l1, l2);
        nll[b] = m == -INFINITY ? INFINITY : -(m + logf(expf(l1 - m) + expf(l2 - m)));
      }
    }
  }
}

// states per lane and chunks of the warp form (0: does not apply — the thread-per-state kernel takes the lattice)
static int ww_plan(int units, int64_t Lp_max, int* G_out) {
  // MEASURED SLOWER than the thread-per-state form (1-hour lattice 15.7 vs 10.6 ms, N = 16384: 4.9 vs 3.5, cfg 5: 0.62 vs
  // 0.45): one warp issues ~250 instructions per frame for its six states at an IPC of 0.38 — the six dependent chains do
  // not overlap as well as six warps do.  Kept behind LCASR_CTC_WF_WARP=1 as the record of the experiment.
  static const bool on = getenv("LCASR_CTC_WF_WARP") && atoi(getenv("LCASR_CTC_WF_WARP")) == 1;
  if (!on || units < 1 || units > kNumSMs) return 0;
  const int G_max = kNumSMs / units;
  for (int spl : {2, 4, 6, 8}) {
    const int64_t G = ceil_div(Lp_max, (int64_t)32 * spl);
    if (G <= G_max && G >= 2) { *G_out = (int)G; return spl; }
  }
  return 0;
}

static int wf_tb() {
  static const int tb = getenv("LCASR_CTC_WF_TB") ? atoi(getenv("LCASR_CTC_WF_TB")) : kWfTBDefault;
  return tb == 8 || tb == 32 ? tb : 16;
}

static int wf_chunks(int units, int64_t Lp_max) {
  if (units < 1 || units > kNumSMs) return 0;
  int G = kNumSMs / units;
  const int64_t min_chunk = 64;  // below this the hand-off costs more than the shorter chain saves
  if ((int64_t)G > ceil_div(Lp_max, min_chunk)) G = (int)ceil_div(Lp_max, min_chunk);
  if (G < 1) G = 1;
  if (ceil_div(Lp_max, (int64_t)G) > 1024) return 0;
  return G;
}

int64_t ctc_wavefront_workspace_bytes(int units, int64_t N, int64_t Lp_max) {
  const int G = wf_chunks(units, Lp_max);
  if (G < 2) return 0;  // one chunk per lattice: the plain one-CTA recursion is the same thing
  return (int64_t)units * G * N * 16;
}

static int launch_wavefront(const float* lp, int B, int64_t N, int V, const int64_t* tg, int64_t S_max, const int32_t* il,
                            const int64_t* tl, int blank, int dir, float* nll, float* store, float* store_beta, int pregathered,
                            void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const int units = dir == 0 ? 2 * B : B;
  const int64_t Lp_max = 2 * S_max + 1;
  const int G = wf_chunks(units, Lp_max);
  LCASR_CHECK_ARG(G > 0, "ctc wavefront: %d lattices of %lld states do not fit this GPU", units, (long long)Lp_max);
  LCASR_CHECK_ARG(workspace && workspace_bytes >= ctc_wavefront_workspace_bytes(units, N, Lp_max) && ((uintptr_t)workspace & 15) == 0,
                  "ctc wavefront: workspace too small or misaligned");
  LCASR_CHECK_ARG(N < ((int64_t)1 << 31), "ctc wavefront: too many frames");
  int nt = (int)round_up(ceil_div(Lp_max, (int64_t)G), 32);
  float4* bnd = (float4*)workspace;
  // every launch tags its records with a fresh epoch: records left in the workspace by earlier launches never match
  static std::atomic<unsigned> g_epoch{0x5eed0000u};
  unsigned epoch = g_epoch.fetch_add(1, std::memory_order_relaxed) + 1;
  const int tb = wf_tb();
  int batch = B;
  void* args[] = {(void*)&lp, (void*)&N, (void*)&V, (void*)&tg, (void*)&S_max, (void*)&il, (void*)&tl, (void*)&blank, (void*)&dir,
                  (void*)&G, (void*)&batch, (void*)&nll, (void*)&store, (void*)&store_beta, (void*)&pregathered, (void*)&bnd,
                  (void*)&epoch};
  const bool has_store_w = store != nullptr || store_beta != nullptr;
  int Gw = 0;
  const int spl = ww_plan(units, Lp_max, &Gw);
  if (spl > 0 && Gw <= G) {  // the records fit the workspace sized for G chunks
    void* wargs[] = {(void*)&lp, (void*)&N, (void*)&V, (void*)&tg, (void*)&S_max, (void*)&il, (void*)&tl, (void*)&blank, (void*)&dir,
                     (void*)&Gw, (void*)&batch, (void*)&nll, (void*)&store, (void*)&store_beta, (void*)&pregathered, (void*)&bnd,
                     (void*)&epoch};
    const void* wfn = nullptr;
#define LCASR_WW(S_)                                                                                                     \
  case S_:                                                                                                               \
    wfn = has_store_w ? (const void*)ctc_wavefront_warp_kernel<16, true, S_> : (const void*)ctc_wavefront_warp_kernel<16, false, S_>; \
    break;
    switch (spl) { LCASR_WW(2) LCASR_WW(4) LCASR_WW(6) LCASR_WW(8) default: break; }
#undef LCASR_WW
    LCASR_CUDA(cudaLaunchCooperativeKernel(wfn, dim3((unsigned)(units * Gw)), dim3((unsigned)kWwThreads), wargs, 0, st));
    count_launch();
    return 0;
  }
  // cooperative launch: every CTA is resident, so waiting for a neighbour's records always makes progress
  const void* fn = nullptr;
  const bool small = nt <= 256;  // chunks of <= 256 states: up to 255 registers per thread for the look-ahead ring
  const bool has_store = store != nullptr || store_beta != nullptr;
#define LCASR_WF(TB)                                                                                                          \
  fn = has_store ? (small ? (const void*)ctc_wavefront_kernel<TB, true, 256> : (const void*)ctc_wavefront_kernel<TB, true, 1024>)   \
                 : (small ? (const void*)ctc_wavefront_kernel<TB, false, 256> : (const void*)ctc_wavefront_kernel<TB, false, 1024>)
  int tb_used = 16;
  if (tb == 8) { LCASR_WF(8); tb_used = 8; } else if (tb == 32 && small) { LCASR_WF(32); tb_used = 32; } else { LCASR_WF(16); }
#undef LCASR_WF
  const size_t smem = (size_t)(2 * (nt + 2) + 2 * 2 * tb_used) * sizeof(float);
  LCASR_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)(units * G)), dim3((unsigned)nt), args, smem, st));
  count_launch();
  return 0;
}

// grad[b,t,c] = exp(lp) - exp(log(sum_{s:ext[s]=c} exp(ab[t,s])) + nll - lp)     (t < input_length)
// ab = alpha+beta (log), both including lp[t,ext[s]] (ATen convention).  One CTA per (t, b).
__global__ void __launch_bounds__(256) ctc_grad_collect_kernel(const float* __restrict__ log_probs, int64_t N, int V,
                                                               const int64_t* __restrict__ targets, int64_t S_max,
                                                               const int32_t* __restrict__ input_lengths,
                                                               const int64_t* __restrict__ target_lengths, int blank,
                                                               const float* __restrict__ nll,
                                                               const float* __restrict__ grad_nll,
                                                               const float* __restrict__ ab, const float* __restrict__ ab2,
                                                               float* __restrict__ grad) {
  // ab2 == NULL: `ab` holds alpha+beta; else ab = alpha, ab2 = beta (kept apart by the concurrent recursion)
  // [V] linear-domain sums relative to the frame max, as 64-bit FIXED-POINT numbers (2^-40 steps): integer atomics are
  // associative, so the sum over the lattice states of a repeated label does not depend on the order in which the
  // threads' atomics land — the CTC gradient feeds every other gradient of the step and is now reproducible.
  extern __shared__ unsigned long long acc[];
  __shared__ float red[8];
  constexpr float kFix = 1099511627776.0f;  // 2^40; every term is exp(v - max) <= 1, at most 2S+1 < 2^23 of them
  const int64_t t = blockIdx.x;
  const int b = blockIdx.y;
  const int64_t T = input_lengths ? min((int64_t)input_lengths[b], N) : N;
  float* g = grad + ((int64_t)b * N + t) * V;
  const int64_t S = target_lengths[b];
  if (t >= T || T < S) {
    for (int c = threadIdx.x; c < V; c += blockDim.x) g[c] = 0.f;
    return;
  }
  const int Lp = (int)(2 * S + 1), Lp_max = (int)(2 * S_max + 1);
  const float* abr = ab + ((int64_t)b * N + t) * Lp_max;
  const float* abr2 = ab2 ? ab2 + ((int64_t)b * N + t) * Lp_max : nullptr;
  const int64_t* tgt = targets + (int64_t)b * S_max;
  const float* lp = log_probs + ((int64_t)b * N + t) * V;
  for (int c = threadIdx.x; c < V; c += blockDim.x) acc[c] = 0ull;
  float mx = -INFINITY;
  for (int s = threadIdx.x; s < Lp; s += blockDim.x) mx = fmaxf(mx, abr2 ? abr[s] + abr2[s] : abr[s]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  if (mx != -INFINITY) {
    unsigned long long blank_sum = 0ull;  // the S+1 blank states would serialise on one shared-memory atomic
    for (int s = threadIdx.x; s < Lp; s += blockDim.x) {
      float v = abr2 ? abr[s] + abr2[s] : abr[s];
      if (v == -INFINITY) continue;
      float e = expf(v - mx);
      const unsigned long long q = (unsigned long long)(e * kFix);
      if (s & 1) atomicAdd(&acc[(int)tgt[(s - 1) >> 1]], q);
      else blank_sum += q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) blank_sum += __shfl_xor_sync(0xffffffffu, blank_sum, o);
    if ((threadIdx.x & 31) == 0 && blank_sum != 0ull) atomicAdd(&acc[blank], blank_sum);
  }
  __syncthreads();
  const float nl = nll[b], gn = grad_nll ? grad_nll[b] : 1.f;
  const bool bad = isinf(nl);  // zero_infinity=False: grads of an inf loss are NaN in ATen; we emit 0*gn
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    float l = lp[c];
    const float a = (float)acc[c] * (1.0f / kFix);
    float occ = a > 0.f ? expf(logf(a) + mx + nl - l) : 0.f;
    g[c] = bad ? 0.f : (expf(l) - occ) * gn;
  }
}

template <int SPT>
static int launch_rec(const float* lp, int B, int64_t N, int V, const int64_t* tg, int64_t S_max, const int32_t* il,
                      const int64_t* tl, int blank, int dir, float* nll, float* store, int nt, cudaStream_t st,
                      float* store_beta = nullptr, int pregathered = 0) {
  size_t smem = (size_t)2 * SPT * nt * sizeof(float);
    if (smem > 48 * 1024) {  // per device and size-dependent: set on every call (cheap)
    LCASR_CUDA(cudaFuncSetAttribute(ctc_recursion_kernel<SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  ctc_recursion_kernel<SPT><<<dir == 0 ? 2 * B : B, nt, smem, st>>>(lp, N, V, tg, S_max, il, tl, blank, dir, nll, store,
                                                                    store_beta, B, pregathered);
  LCASR_LAUNCH_CHECK();
  return 0;
}

static int ctc_recursion(const float* lp, int B, int64_t N, int V, const int64_t* tg, int64_t S_max, const int32_t* il,
                         const int64_t* tl, int blank, int dir, float* nll, float* store, cudaStream_t st,
                         float* store_beta = nullptr, int pregathered = 0, void* workspace = nullptr, int64_t workspace_bytes = 0) {
  const int64_t Lp = 2 * S_max + 1;
  static const bool no_wavefront = getenv("LCASR_CTC_NO_WAVEFRONT") != nullptr;  // A/B switch
  if (workspace && !no_wavefront) {  // time-skewed wavefront over the whole GPU whenever a lattice can be cut into >= 2 chunks
    const int units = dir == 0 ? 2 * B : B;
    if (wf_chunks(units, Lp) >= 2)
      return launch_wavefront(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, store_beta, pregathered, workspace,
                              workspace_bytes, st);
  }
  LCASR_CHECK_ARG(dir != 0 || Lp <= 4096, "ctc_loss: the concurrent alpha/beta form covers up to 4096 extended states");
  static const bool no_cluster = getenv("LCASR_CTC_NO_CLUSTER") != nullptr;  // A/B switch for profiling
  if (Lp > 4096 && !no_cluster) {  // spread the states over a cluster of up to 8 SMs (below that the cluster
                                    // barrier costs more than it saves: measured 2.7 vs 1.45 ms at 1229 states)
    LCASR_CHECK_ARG(Lp <= 8 * 1024 * 4, "ctc_loss: %lld extended states exceed one cluster (max 32768)", (long long)Lp);
    int spt = 1;
    while ((int64_t)8 * 1024 * spt < Lp) spt *= 2;          // fewest states per thread that still fits 8 CTAs
    const int csize = (int)ceil_div(Lp, (int64_t)1024 * spt);
    const int nt = 1024;
    switch (spt) {
      case 1: return launch_cluster<1>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, csize, st);
      case 2: return launch_cluster<2>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, csize, st);
      default: return launch_cluster<4>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, csize, st);
    }
  }
  LCASR_CHECK_ARG(Lp <= (int64_t)kCtcMaxSPT * 1024 && Lp * 8 <= 227 * 1024 - 64,
                  "ctc_loss: %lld extended states do not fit one CTA's shared memory (max 29048)", (long long)Lp);
  int spt = 1;
  while ((int64_t)spt * 1024 < Lp) spt *= 2;
  if (Lp <= 512) spt = 1;
  int nt = (int)round_up(ceil_div(Lp, spt), 32);
  if (nt > 1024) nt = 1024;
  // shrink padding for the big cases so 2*SPT*nt*4 stays within 227 KB
  while ((size_t)2 * spt * nt * 4 > 227 * 1024 && nt > 32) nt -= 32;
  LCASR_CHECK_ARG((int64_t)spt * nt >= Lp, "ctc_loss: internal sizing error");
  switch (spt) {
    case 1: return launch_rec<1>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, st, store_beta, pregathered);
    case 2: return launch_rec<2>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, st, store_beta, pregathered);
    case 4: return launch_rec<4>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, st, store_beta, pregathered);
    case 8: return launch_rec<8>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, st, store_beta, pregathered);
    case 16: return launch_rec<16>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, st, store_beta, pregathered);
    default: return launch_rec<32>(lp, B, N, V, tg, S_max, il, tl, blank, dir, nll, store, nt, st, store_beta, pregathered);
  }
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int64_t lcasr_ctc_workspace_bytes(int B, int64_t N, int64_t S_max, int both_directions) {
  if (B <= 0 || N <= 0 || S_max < 0) return -1;
  return ctc_wavefront_workspace_bytes(both_directions ? 2 * B : B, N, 2 * S_max + 1);
}

extern "C" int lcasr_ctc_loss_fwd_ws(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                     int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                     int blank, float* nll, float* alpha_ws, void* workspace, int64_t workspace_bytes,
                                     void* stream) {
  LCASR_CHECK_ARG(log_probs && targets && target_lengths && nll, "ctc_loss_fwd: NULL argument");
  LCASR_CHECK_ARG(B > 0 && N > 0 && V > 1 && S_max >= 0 && blank >= 0 && blank < V, "ctc_loss_fwd: bad shape");
  return ctc_recursion(log_probs, B, N, V, targets, S_max, input_lengths, target_lengths, blank, +1, nll, alpha_ws,
                       (cudaStream_t)stream, nullptr, 0, workspace, workspace_bytes);
}

extern "C" int lcasr_ctc_loss_fwd(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                  int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                  int blank, float* nll, float* alpha_ws, void* stream) {
  return lcasr_ctc_loss_fwd_ws(log_probs, B, N, V, targets, S_max, input_lengths, target_lengths, blank, nll, alpha_ws, nullptr, 0,
                               stream);
}

extern "C" int lcasr_ctc_loss_bwd_ws(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                     int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                     int blank, const float* nll, const float* grad_nll, const float* alpha_ws,
                                     float* beta_ws, float* grad, void* workspace, int64_t workspace_bytes, void* stream);

extern "C" int lcasr_ctc_loss_bwd(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                  int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                  int blank, const float* nll, const float* grad_nll, const float* alpha_ws,
                                  float* beta_ws, float* grad, void* stream) {
  return lcasr_ctc_loss_bwd_ws(log_probs, B, N, V, targets, S_max, input_lengths, target_lengths, blank, nll, grad_nll, alpha_ws,
                               beta_ws, grad, nullptr, 0, stream);
}

extern "C" int lcasr_ctc_loss_bwd_ws(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                     int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                     int blank, const float* nll, const float* grad_nll, const float* alpha_ws,
                                     float* beta_ws, float* grad, void* workspace, int64_t workspace_bytes, void* stream) {
  LCASR_CHECK_ARG(log_probs && targets && target_lengths && nll && alpha_ws && beta_ws && grad,
                  "ctc_loss_bwd: NULL argument");
  LCASR_CHECK_ARG(B > 0 && N > 0 && V > 1 && S_max >= 0 && blank >= 0 && blank < V, "ctc_loss_bwd: bad shape");
  LCASR_CHECK_ARG((size_t)V * 4 <= 200 * 1024, "ctc_loss_bwd: V=%d too large", V);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t Lp_max = 2 * S_max + 1;
  // beta_ws <- alpha, then the reversed recursion adds beta - lp[t, ext[s]] ... we store alpha+beta
  LCASR_CUDA(cudaMemcpyAsync(beta_ws, alpha_ws, (size_t)B * N * Lp_max * sizeof(float), cudaMemcpyDeviceToDevice, st));
  LCASR_TRY(ctc_recursion(log_probs, B, N, V, targets, S_max, input_lengths, target_lengths, blank, -1, nullptr,
                          beta_ws, st, nullptr, 0, workspace, workspace_bytes));
  size_t smem = (size_t)V * sizeof(unsigned long long);
    if (smem > 48 * 1024) {  // per device and size-dependent: set on every call (cheap)
    LCASR_CUDA(cudaFuncSetAttribute(ctc_grad_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  LCASR_CHECK_ARG(B <= 65535, "ctc_loss_bwd: batch too large");
  dim3 grid((unsigned)N, (unsigned)B);
  ctc_grad_collect_kernel<<<grid, 256, smem, st>>>(log_probs, N, V, targets, S_max, input_lengths, target_lengths, blank,
                                                   nll, grad_nll, beta_ws, nullptr, grad);
  LCASR_LAUNCH_CHECK();
  return 0;
}

// Training form: alpha and beta recursions in ONE launch (2*B CTAs, independent, concurrent) ...
extern "C" int lcasr_ctc_loss_fwd_ab_ws(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                        int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                        int blank, float* nll, float* alpha_ws, float* beta_ws, void* workspace,
                                        int64_t workspace_bytes, void* stream);

extern "C" int lcasr_ctc_loss_fwd_ab(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                     int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                     int blank, float* nll, float* alpha_ws, float* beta_ws, void* stream) {
  return lcasr_ctc_loss_fwd_ab_ws(log_probs, B, N, V, targets, S_max, input_lengths, target_lengths, blank, nll, alpha_ws, beta_ws,
                                  nullptr, 0, stream);
}

extern "C" int lcasr_ctc_loss_fwd_ab_ws(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                        int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                        int blank, float* nll, float* alpha_ws, float* beta_ws, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
  LCASR_CHECK_ARG(log_probs && targets && target_lengths && nll && alpha_ws && beta_ws, "ctc_loss_fwd_ab: NULL argument");
  LCASR_CHECK_ARG(B > 0 && N > 0 && V > 1 && S_max >= 0 && blank >= 0 && blank < V, "ctc_loss_fwd_ab: bad shape");
  LCASR_CHECK_ARG(B <= 65535 && N < ((int64_t)1 << 31), "ctc_loss_fwd_ab: batch / length too large");
  dim3 grid((unsigned)N, (unsigned)B);
  ctc_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(log_probs, N, V, targets, S_max, input_lengths, target_lengths, blank,
                                                            alpha_ws, beta_ws);
  LCASR_LAUNCH_CHECK();
  return ctc_recursion(log_probs, B, N, V, targets, S_max, input_lengths, target_lengths, blank, 0, nll, alpha_ws,
                       (cudaStream_t)stream, beta_ws, 1, workspace, workspace_bytes);
}

// ... and the gradient from the two state lattices (no recursion left in the backward).
extern "C" int lcasr_ctc_loss_grad(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                                   int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                                   int blank, const float* nll, const float* grad_nll, const float* alpha_ws,
                                   const float* beta_ws, float* grad, void* stream) {
  LCASR_CHECK_ARG(log_probs && targets && target_lengths && nll && alpha_ws && beta_ws && grad, "ctc_loss_grad: NULL argument");
  LCASR_CHECK_ARG(B > 0 && B <= 65535 && N > 0 && V > 1 && S_max >= 0 && blank >= 0 && blank < V, "ctc_loss_grad: bad shape");
  LCASR_CHECK_ARG((size_t)V * 4 <= 200 * 1024, "ctc_loss_grad: V=%d too large", V);
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = (size_t)V * sizeof(unsigned long long);
    if (smem > 48 * 1024) {  // per device and size-dependent: set on every call (cheap)
    LCASR_CUDA(cudaFuncSetAttribute(ctc_grad_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((unsigned)N, (unsigned)B);
  ctc_grad_collect_kernel<<<grid, 256, smem, st>>>(log_probs, N, V, targets, S_max, input_lengths, target_lengths, blank,
                                                   nll, grad_nll, alpha_ws, beta_ws, grad);
  LCASR_LAUNCH_CHECK();
  return 0;
}
