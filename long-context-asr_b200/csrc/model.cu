// Host-side orchestration of SCConformerXL.forward (lcasr/models/sconformer_xl.py:162-252,
// ConformerLayer.forward :346-372) as a fixed sequence of kernel launches on one stream.
// No host synchronisation, no allocation: every intermediate lives in the caller's workspace.
//
// HBM layout (all row-major, "M" = B*N tokens, e = bytes of the compute dtype):
//   x      [M, d]            fp32   residual stream (kept fp32 in both precisions, >= the reference)
//   a      [M, d]            e      normalised GEMM operand / attention output
//   wide   [M, max(4d,V1)]   e      FFN hidden | qkv | pw1 pre-GLU | self-conditioning logits/probs
//   q,k,v  [M, d]            e      (g / c of the conv module alias q / k)
//   subsampling scratch (dead once x is produced) aliases a..v:
//     s1 [B,T1,F1,C]  s2 [B,T2,F2,C] x2   s3 [B,N,F3,C] x2     channels-last
#include "common.cuh"
#include <vector>
#include <new>
#include <cstdlib>

namespace lcasr {
bool norm_chain_ok(int d);
int attn_tc_available();
int gemm_tc_launch_rope(const void* A, const void* W, int64_t M, int N, int K, const float* cos_t, const float* sin_t,
                        int64_t rope_n, int rope_cols, int dh, void* out, cudaStream_t st);
int gemm_tc_launch_glu(const void* A, const void* W, int64_t M, int N, int K, const float* bias, void* out, cudaStream_t st);
int attn_tc_launch(const void* q, const void* k, const void* v, int B, int64_t N, int64_t Nk, const int32_t* kv_len, int H,
                   int Dh, int v_transposed, int64_t Npad, void* out, float* lse, cudaStream_t st, int wl = -1, int wr = -1,
                   float* out32 = nullptr, int64_t ldq = 0, int64_t ldkv = 0);
int attn_merge_launch(const float* parts, const float* lses, int P, int64_t n, int H, int Dh, int64_t part_stride,
                      int64_t lse_stride, void* out, int out_dtype, cudaStream_t st);
}

using namespace lcasr;

#include "model_internal.cuh"
#include <algorithm>

namespace {

struct Plan {
  int B; int64_t T, T1, T2, N, M; int F1, F2, F3;
  size_t off_x, off_cos, off_sin, off_a, off_wide, off_q, off_k, off_v;
  size_t off_s1, off_s2a, off_s2b, off_s3a, off_s3b;
  size_t total;
  int64_t Npad;
};

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

Plan make_plan(const lcasr_config& c, int B, int64_t T) {
  Plan p{};
  p.B = B; p.T = T;
  p.T1 = (T - 1) / 2 + 1; p.T2 = (p.T1 - 1) / 2 + 1; p.N = (p.T2 - 1) / 2 + 1;
  p.F1 = (c.feat_in - 1) / 2 + 1; p.F2 = (p.F1 - 1) / 2 + 1; p.F3 = (p.F2 - 1) / 2 + 1;
  p.M = (int64_t)B * p.N;
  p.Npad = round_up(p.N, 128);
  const size_t e = dtype_size(c.compute_dtype);
  const int d = c.d_model, C = c.conv_channels, V1 = c.num_classes;
  size_t o = 0;
  p.off_x = o; o += al((size_t)p.M * d * 4);
  p.off_cos = o; o += al((size_t)p.N * (c.head_dim / 2) * 4);
  p.off_sin = o; o += al((size_t)p.N * (c.head_dim / 2) * 4);
  const size_t scratch0 = o;
  // layer scratch
  p.off_a = o; o += al((size_t)p.M * d * e);
  size_t wide_cols = (size_t)4 * d;
  if ((size_t)V1 > wide_cols) wide_cols = V1;
  p.off_wide = o; o += al((size_t)p.M * wide_cols * e);
  p.off_q = o; o += al((size_t)p.M * d * e);
  p.off_k = o; o += al((size_t)p.M * d * e);
  p.off_v = o; o += al((size_t)B * c.n_heads * c.head_dim * p.Npad * e);  // >= M*d: also holds the transposed layout
  const size_t layer_end = o;
  // subsampling scratch aliases the layer scratch
  o = scratch0;
  p.off_s1 = o; o += al((size_t)B * p.T1 * p.F1 * C * e);
  p.off_s2a = o; o += al((size_t)B * p.T2 * p.F2 * C * e);
  p.off_s2b = o; o += al((size_t)B * p.T2 * p.F2 * C * e);
  p.off_s3a = o; o += al((size_t)B * p.N * p.F3 * C * e);
  p.off_s3b = o; o += al((size_t)B * p.N * p.F3 * C * e);
  p.total = o > layer_end ? o : layer_end;
  return p;
}

// ---- wave quantisation of the attention launch ----
// One CTA of the flash kernels owns 256 queries of one head and walks ALL keys, one CTA per SM: the launch is U = B*H*ceil(N/256)
// equal units on W SMs and takes ceil(U / W) unit-times — 1536 units on 148 SMs are 10.38 waves of work in 11 (cfg 3), 2816
// units 19.03 in 20 (cfg 4), 768 units 5.19 in 6 (cfg 2).  The last t query-tile pairs of the last recording are therefore
// computed as P key-range pieces (exact partial results + the merge kernel of the sequence-parallel path) on side streams:
// the short CTAs fill the SMs the last wave leaves idle.  Greedy list-schedule model, all candidates (t, P) tried.
double attention_makespan(int W, int64_t full, double full_cost, int64_t pieces, double piece_cost) {
  // the full-length units are equal, so their greedy assignment is known in closed form (full / W each, the first full % W
  // SMs one more); only the few pieces are list-scheduled on the heap of SM-free times — planning stays O(pieces log W)
  // however large the batch is
  std::vector<double> sm((size_t)W);
  const int64_t base = full / W, extra = full % W;
  for (int i = 0; i < W; ++i) sm[(size_t)i] = (double)(base + (i < extra ? 1 : 0)) * full_cost;
  auto cmp = [](double a, double b) { return a > b; };
  std::make_heap(sm.begin(), sm.end(), cmp);
  for (int64_t i = 0; i < pieces; ++i) {
    std::pop_heap(sm.begin(), sm.end(), cmp);
    sm.back() += piece_cost;
    std::push_heap(sm.begin(), sm.end(), cmp);
  }
  return *std::max_element(sm.begin(), sm.end());
}

lcasr_model::TailPlan plan_attention_tail(int B, int64_t N, int H, int W) {
  lcasr_model::TailPlan best;
  best.B = B; best.N = N;
  static const bool off = getenv("LCASR_ATTN_TAIL") && atoi(getenv("LCASR_ATTN_TAIL")) == 0;
  const int64_t nq = ceil_div(N, (int64_t)256), nkt = ceil_div(N, (int64_t)128);
  const int64_t U = (int64_t)B * H * nq;
  if (off || U <= W || nkt < 16) return best;
  const double ovh = 1.5 / (double)nkt;  // prologue + epilogue of a CTA in unit-times
  const double base = attention_makespan(W, U, 1.0 + ovh, 0, 0.0);
  double best_ms = base * 0.985;  // worth two more launches and a merge only beyond 1.5 %
  for (int t = 1; t <= 8 && t <= nq; ++t)
    for (int P = 2; P <= 4; ++P) {
      if (nkt / P < 4) continue;
      const double piece = (double)ceil_div(nkt, (int64_t)P) / (double)nkt + ovh;
      const double ms = attention_makespan(W, U - (int64_t)t * H, 1.0 + ovh, (int64_t)t * H * P, piece);
      if (ms < best_ms) { best_ms = ms; best.t = t; best.P = P; }
    }
  return best;
}

}  // namespace

extern "C" int lcasr_model_create(const lcasr_config* cfg, const lcasr_weights* w, lcasr_model** out) {
  LCASR_CHECK_ARG(cfg && w && out, "model_create: NULL argument");
  LCASR_CHECK_ARG(cfg->abi_version == LCASR_ABI_VERSION, "model_create: ABI version %d != %d", cfg->abi_version,
                  LCASR_ABI_VERSION);
  LCASR_CHECK_ARG(cfg->n_layers > 0 && cfg->d_model > 0 && cfg->n_heads > 0 && cfg->head_dim > 0, "model_create: bad dims");
  LCASR_CHECK_ARG(cfg->n_heads * cfg->head_dim == cfg->d_model,
                  "model_create: n_heads*head_dim (%d) != d_model (%d) is not supported", cfg->n_heads * cfg->head_dim,
                  cfg->d_model);
  LCASR_CHECK_ARG(cfg->d_model % 8 == 0 && cfg->conv_channels % 8 == 0 && cfg->num_classes % 8 == 0,
                  "model_create: d_model, conv_channels and num_classes must be multiples of 8");
  LCASR_CHECK_ARG(cfg->head_dim == 32 || cfg->head_dim == 64 || cfg->head_dim == 128,
                  "model_create: head_dim=%d not in {32,64,128}", cfg->head_dim);
  LCASR_CHECK_ARG(cfg->compute_dtype == LCASR_F32 || cfg->compute_dtype == LCASR_BF16, "model_create: bad compute dtype");
  LCASR_CHECK_ARG(w->layers_host, "model_create: layers_host is NULL");
  LCASR_CHECK_ARG(!cfg->use_rotary || w->inv_freq, "model_create: use_rotary without inv_freq");
  lcasr_model* m = new (std::nothrow) lcasr_model();
  if (!m) return set_error(LCASR_E_NOMEM, "model_create: out of host memory");
  m->cfg = *cfg;
  m->w = *w;
  m->layers.assign(w->layers_host, w->layers_host + cfg->n_layers);
  m->w.layers_host = m->layers.data();
  // side streams / events of the attention tail split, on the device that holds the weights (not necessarily the caller's
  // current device)
  int prev_dev = 0, w_dev = 0;
  cudaGetDevice(&prev_dev);
  w_dev = prev_dev;
  cudaPointerAttributes pa;
  if (w->conv0_w && cudaPointerGetAttributes(&pa, w->conv0_w) == cudaSuccess && pa.type == cudaMemoryTypeDevice) w_dev = pa.device;
  else cudaGetLastError();
  if (w_dev != prev_dev) cudaSetDevice(w_dev);
  bool ok = cudaEventCreateWithFlags(&m->tail_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < lcasr_model::kTailStreams; ++i)
    ok = cudaStreamCreateWithFlags(&m->tail_streams[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&m->tail_done[i], cudaEventDisableTiming) == cudaSuccess;
  if (w_dev != prev_dev) cudaSetDevice(prev_dev);
  if (!ok) {
    cudaGetLastError();
    delete m;
    return set_error(LCASR_E_CUDA, "model_create: cannot create the side streams of the attention tail split");
  }
  *out = m;
  return 0;
}

extern "C" void lcasr_model_destroy(lcasr_model* m) { delete m; }

extern "C" int lcasr_model_set_impl(lcasr_model* m, int gemm_impl, int attn_impl) {
  LCASR_CHECK_ARG(m, "model_set_impl: NULL model");
  m->gemm_impl = gemm_impl;
  m->attn_impl = attn_impl;
  return 0;
}

// host-only: the plan lcasr_model_forward would choose for (B recordings, N tokens, H heads) on a GPU with `sms` SMs
extern "C" int lcasr_attention_tail_plan(int B, int64_t N, int H, int sms, int* tail_pairs, int* key_pieces) {
  LCASR_CHECK_ARG(B > 0 && N > 0 && H > 0 && sms > 0 && tail_pairs && key_pieces, "attention_tail_plan: bad arguments");
  const lcasr_model::TailPlan tp = plan_attention_tail(B, N, H, sms);
  *tail_pairs = tp.t;
  *key_pieces = tp.P;
  return 0;
}

extern "C" int lcasr_model_set_attention_tail(lcasr_model* m, int tail_pairs, int key_pieces) {
  LCASR_CHECK_ARG(m, "model_set_attention_tail: NULL model");
  LCASR_CHECK_ARG(tail_pairs <= 0 || (key_pieces >= 2 && key_pieces <= 4), "model_set_attention_tail: key_pieces %d not in 2..4",
                  key_pieces);
  m->tail_force_t = tail_pairs;
  m->tail_force_p = key_pieces;
  m->tail_plan = lcasr_model::TailPlan();
  return 0;
}

extern "C" int64_t lcasr_model_workspace_bytes(const lcasr_model* m, int B, int64_t T) {
  if (!m || B <= 0 || T <= 0) return -1;
  return (int64_t)make_plan(m->cfg, B, T).total;
}

extern "C" int lcasr_model_forward_lengths(lcasr_model* m, const float* spec, int B, int64_t T, const int32_t* tok_len,
                                           float* out, int32_t* argmax, int return_logits, void* workspace,
                                           int64_t workspace_bytes, void* stream);

extern "C" int lcasr_model_forward(lcasr_model* m, const float* spec, int B, int64_t T, float* out, int32_t* argmax,
                                   int return_logits, void* workspace, int64_t workspace_bytes, void* stream) {
  return lcasr_model_forward_lengths(m, spec, B, T, nullptr, out, argmax, return_logits, workspace, workspace_bytes, stream);
}

// tok_len (device int32[B], may be NULL): valid TOKENS per recording after subsampling (ragged batch,
// sconformer_xl.py:204-213).  Keys beyond it are masked in attention (attention.py:511-547) and the conv
// module's input is zeroed there (convolution.py:109-110); outputs of padded frames are unspecified.
extern "C" int lcasr_model_forward_lengths(lcasr_model* m, const float* spec, int B, int64_t T, const int32_t* tok_len,
                                           float* out, int32_t* argmax, int return_logits, void* workspace,
                                           int64_t workspace_bytes, void* stream) {
  LCASR_CHECK_ARG(m && spec && out && workspace, "model_forward: NULL argument");
  LCASR_CHECK_ARG(B > 0 && T > 0, "model_forward: bad shape B=%d T=%lld", B, (long long)T);
  const lcasr_config& c = m->cfg;
  const lcasr_weights& w = m->w;
  const Plan p = make_plan(c, B, T);
  LCASR_CHECK_ARG(p.N > 0, "model_forward: input too short");
  LCASR_CHECK_ARG((size_t)workspace_bytes >= p.total, "model_forward: workspace %lld < required %lld bytes",
                  (long long)workspace_bytes, (long long)p.total);
  LCASR_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "model_forward: workspace must be 256-byte aligned");
  char* ws = (char*)workspace;
  const int cd = c.compute_dtype;
  const int d = c.d_model, C = c.conv_channels, V1 = c.num_classes, H = c.n_heads, Dh = c.head_dim;
  const int64_t M = p.M, N = p.N;
  float* x = (float*)(ws + p.off_x);
  float* cos_t = (float*)(ws + p.off_cos);
  float* sin_t = (float*)(ws + p.off_sin);
  void* a = ws + p.off_a; void* wide = ws + p.off_wide;
  void* q = ws + p.off_q; void* k = ws + p.off_k; void* v = ws + p.off_v;
  const int gi = m->gemm_impl;
  int ai = m->attn_impl;
  if (ai == LCASR_ATTN_AUTO) ai = (cd == LCASR_BF16 && attn_tc_available()) ? LCASR_ATTN_TCGEN05 : LCASR_ATTN_SIMT;
  void* const a2 = a;  // attention output (the LayerNorm output in `a` is dead once the qkv GEMM has run)
  static const bool no_fuse = getenv("LCASR_NO_FUSED_EPILOGUES") != nullptr;  // A/B switch: separate rope_split / glu passes
  const int vt = 0;  // natural [B,N,H,Dh] V: the tcgen05 kernel consumes it as an MN-major B operand (no transpose pass)

  cudaStream_t cst = (cudaStream_t)stream;
  auto timed = [&](int cat, int status) {  // closes the span opened by begin()
    if (m->timing && !m->spans.empty() && m->spans.back().e1 == (size_t)-1) {
      m->spans.back().e1 = m->ev_used;
      cudaEventRecord(m->next_event(), cst);
    }
    (void)cat;
    return status;
  };
  auto begin = [&](int cat) {
    if (m->timing) {
      m->spans.push_back({cat, m->ev_used, (size_t)-1});
      cudaEventRecord(m->next_event(), cst);
    }
    return cat;
  };
#define OP(cat, expr)            \
  do {                           \
    begin(cat);                  \
    int _st = (expr);            \
    timed(cat, _st);             \
    if (_st != 0) return _st;    \
  } while (0)
  auto gemm = [&](const void* A, const void* W, int64_t rows, int n, int kk, const float* bias, int act,
                  const float* resid, float alpha, void* o, int odt) {
    begin(CAT_GEMM);
    return timed(CAT_GEMM, lcasr_gemm(A, W, cd, rows, n, kk, bias, act, resid, alpha, o, odt, gi, stream));
  };
  auto norm = [&](const float* nw, const float* nb, float* o32, void* olo) {
    begin(CAT_NORM);
    return timed(CAT_NORM, lcasr_layernorm(x, nw, nb, M, d, c.norm_eps, c.norm_kind, o32, olo, cd, stream));
  };

  // ---- subsampling (subsampling.py:384-428) ----
  {
    void* s1 = ws + p.off_s1; void* s2a = ws + p.off_s2a; void* s2b = ws + p.off_s2b;
    void* s3a = ws + p.off_s3a; void* s3b = ws + p.off_s3b;
    if (cd == LCASR_BF16 && C % 64 == 0) {  // conv0 + SiLU + depthwise level 1 fused: the 160x activation stays on chip
      OP(CAT_SUBSAMPLE, lcasr_subsample_conv0_dw(spec, w.conv0_w, w.conv0_b, w.dw1_w, w.dw1_b, B, c.feat_in, T, C, s2a, stream));
    } else {
      OP(CAT_SUBSAMPLE, lcasr_subsample_conv0(spec, w.conv0_w, w.conv0_b, B, c.feat_in, T, C, s1, cd, stream));
      OP(CAT_SUBSAMPLE, lcasr_subsample_dwconv(s1, cd, w.dw1_w, w.dw1_b, B, p.T1, p.F1, C, s2a, stream));
    }
    LCASR_TRY(gemm(s2a, w.pw1_w, (int64_t)B * p.T2 * p.F2, C, C, w.pw1_b, LCASR_ACT_SILU, nullptr, 0.f, s2b, cd));
    OP(CAT_SUBSAMPLE, lcasr_subsample_dwconv(s2b, cd, w.dw2_w, w.dw2_b, B, p.T2, p.F2, C, s3a, stream));
    LCASR_TRY(gemm(s3a, w.pw2_w, (int64_t)B * N * p.F3, C, C, w.pw2_b, LCASR_ACT_SILU, nullptr, 0.f, s3b, cd));
    LCASR_TRY(gemm(s3b, w.sub_out_w, M, d, p.F3 * C, nullptr, LCASR_ACT_NONE, nullptr, 0.f, x, LCASR_F32));
  }
  // fused rotary (bf16 tensor-core path) reads pair-major tables; the split / rotate kernel position-major ones
  const bool fused_rope_tables = !no_fuse && cd == LCASR_BF16 && gi != LCASR_GEMM_SIMT && ai == LCASR_ATTN_TCGEN05 && d % 32 == 0 &&
                                 c.attn_window_left < 0 && c.attn_window_right < 0 && m->layers[0].qkv_w_il;
  if (c.use_rotary) {
    if (fused_rope_tables) OP(CAT_ROPE, lcasr_rope_table_t(w.inv_freq, c.rotary_interp, 0, N, Dh / 2, cos_t, sin_t, stream));
    else OP(CAT_ROPE, lcasr_rope_table(w.inv_freq, c.rotary_interp, 0, N, Dh / 2, cos_t, sin_t, stream));
  }

  auto ffn = [&](const float* nw, const float* nb, const void* fc1, const float* b1, const void* fc2, const float* b2) {
    LCASR_TRY(norm(nw, nb, nullptr, a));
    LCASR_TRY(gemm(a, fc1, M, 4 * d, d, b1, LCASR_ACT_GELU_TANH, nullptr, 0.f, wide, cd));
    LCASR_TRY(gemm(wide, fc2, M, d, 4 * d, b2, LCASR_ACT_NONE, x, 0.5f, x, LCASR_F32));  // x += 0.5*ffn
    return 0;
  };

  // transposed-V layout: the key padding columns [N, Npad) are read by TMA (times P == 0): keep them finite
  if (vt) LCASR_CUDA(cudaMemsetAsync(v, 0, (size_t)B * H * Dh * p.Npad * dtype_size(cd), cst));

  // attention of the fused path with the last wave filled (plan_attention_tail); the partial results live in the q / k / v
  // buffers, which the fused path does not use
  if (m->tail_plan.B != B || m->tail_plan.N != N) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    m->tail_plan = plan_attention_tail(B, N, H, sms);
    if (m->tail_force_t >= 0) {
      m->tail_plan.t = (int)std::min<int64_t>(m->tail_force_t, ceil_div(N, (int64_t)256));
      m->tail_plan.P = m->tail_force_p;
      if (m->tail_plan.P > ceil_div(N, (int64_t)128)) m->tail_plan.t = 0;
    }
  }
  lcasr_model::TailPlan tail = m->tail_plan;
  {
    const int64_t rows_tail = N - std::max<int64_t>(0, ceil_div(N, (int64_t)256) - tail.t) * 256;
    const size_t need = (size_t)tail.P * rows_tail * (d + H) * 4;
    if (cd != LCASR_BF16 || need > (size_t)(p.off_v - p.off_q) + (size_t)M * d * 2) tail.t = 0;
  }
  auto attention_qkv_tail_split = [&]() -> int {
    const int64_t ld = 3 * (int64_t)d;
    const int64_t n_main = (ceil_div(N, (int64_t)256) - tail.t) * 256, rows_tail = N - n_main;
    const char* qb = (const char*)wide + (size_t)(B - 1) * N * ld * 2;  // the last recording
    const char* kb = qb + (size_t)d * 2;
    const char* vb = qb + (size_t)2 * d * 2;
    char* ob = (char*)a2 + (size_t)(B - 1) * N * d * 2;
    float* parts = (float*)q;
    float* lses = parts + (size_t)tail.P * rows_tail * d;
    cudaStream_t* ss = m->tail_streams;
    LCASR_CUDA(cudaEventRecord(m->tail_fork, cst));
    const int n_side = tail.P + 1;
    for (int i = (B > 1 ? 0 : 1); i < n_side; ++i) LCASR_CUDA(cudaStreamWaitEvent(ss[i], m->tail_fork, 0));
    // full-length CTAs first (launch order = dispatch order): the other recordings, then the last recording's leading rows —
    // on their own stream when both exist, so that neither waits for the other's last wave
    if (B > 1) {
      const char* w0 = (const char*)wide;
      LCASR_TRY(attn_tc_launch(w0, w0 + (size_t)d * 2, w0 + (size_t)2 * d * 2, B - 1, N, N, nullptr, H, Dh, 0, 0, a2, nullptr, cst, -1,
                               -1, nullptr, ld, ld));
    }
    if (n_main > 0)
      LCASR_TRY(attn_tc_launch(qb, kb, vb, 1, n_main, N, nullptr, H, Dh, 0, 0, ob, nullptr, B > 1 ? ss[0] : cst, -1, -1, nullptr, ld,
                               ld));
    const int64_t nkt = ceil_div(N, (int64_t)128), tiles = ceil_div(nkt, (int64_t)tail.P);
    int n_parts = 0;
    for (int s = 0; s < tail.P; ++s) {
      const int64_t k0 = s * tiles * 128, k1 = std::min<int64_t>(N, k0 + tiles * 128);
      if (k1 <= k0) break;
      LCASR_TRY(attn_tc_launch(qb + (size_t)n_main * ld * 2, kb + (size_t)k0 * ld * 2, vb + (size_t)k0 * ld * 2, 1, rows_tail, k1 - k0,
                               nullptr, H, Dh, 0, 0, nullptr, lses + (size_t)s * H * rows_tail, ss[1 + s], -1, -1,
                               parts + (size_t)s * rows_tail * d, ld, ld));
      ++n_parts;
    }
    for (int s = 1; s < n_parts; ++s) {
      LCASR_CUDA(cudaEventRecord(m->tail_done[1 + s], ss[1 + s]));
      LCASR_CUDA(cudaStreamWaitEvent(ss[1], m->tail_done[1 + s], 0));
    }
    LCASR_TRY(attn_merge_launch(parts, lses, n_parts, rows_tail, H, Dh, rows_tail * d, (int64_t)H * rows_tail,
                                ob + (size_t)n_main * d * 2, LCASR_BF16, ss[1]));
    LCASR_CUDA(cudaEventRecord(m->tail_done[1], ss[1]));
    LCASR_CUDA(cudaStreamWaitEvent(cst, m->tail_done[1], 0));
    if (B > 1 && n_main > 0) {
      LCASR_CUDA(cudaEventRecord(m->tail_done[0], ss[0]));
      LCASR_CUDA(cudaStreamWaitEvent(cst, m->tail_done[0], 0));
    }
    return 0;
  };

  bool final_operand_ready = false;
  for (int l = 0; l < c.n_layers; ++l) {
    const lcasr_layer_weights& L = m->layers[l];
    LCASR_TRY(ffn(L.ff1_norm_w, L.ff1_norm_b, L.ff1_fc1_w, L.ff1_fc1_b, L.ff1_fc2_w, L.ff1_fc2_b));
    // attention (attention.py:509-551)
    LCASR_TRY(norm(L.attn_norm_w, L.attn_norm_b, nullptr, a));
    // bf16 tensor-core path: rotary in the epilogue of the qkv GEMM (q / k rotated in fp32 before the bf16 store) and
    // attention reading q, k, v as column blocks of the projection — the split / rotate pass (3 reads + 3 writes of M*d) is gone
    const bool fused_qkv = !no_fuse && cd == LCASR_BF16 && gi != LCASR_GEMM_SIMT && ai == LCASR_ATTN_TCGEN05 && d % 32 == 0 &&
                           c.attn_window_left < 0 && c.attn_window_right < 0 && (!c.use_rotary || (L.qkv_w_il && fused_rope_tables));
    if (fused_qkv) {
      if (c.use_rotary) {
        begin(CAT_GEMM);
        LCASR_TRY(timed(CAT_GEMM, gemm_tc_launch_rope(a, L.qkv_w_il, M, 3 * d, d, cos_t, sin_t, N, 2 * d, Dh, wide, cst)));
      } else {
        LCASR_TRY(gemm(a, L.qkv_w, M, 3 * d, d, nullptr, LCASR_ACT_NONE, nullptr, 0.f, wide, cd));
      }
      if (tail.t > 0 && !tok_len) OP(CAT_ATTN, attention_qkv_tail_split());
      else OP(CAT_ATTN, lcasr_attention_qkv(wide, B, N, tok_len, H, Dh, a2, stream));
    } else {
    LCASR_TRY(gemm(a, L.qkv_w, M, 3 * d, d, nullptr, LCASR_ACT_NONE, nullptr, 0.f, wide, cd));
    OP(CAT_ROPE, lcasr_rope_split(wide, cd, B, N, H, Dh, c.use_rotary ? cos_t : nullptr, c.use_rotary ? sin_t : nullptr, q, k,
                               v, vt, p.Npad, stream));
    if (c.attn_window_left >= 0 || c.attn_window_right >= 0)
      OP(CAT_ATTN, lcasr_attention_window(q, k, v, cd, B, N, tok_len, H, Dh, c.attn_window_left, c.attn_window_right, a2, ai, stream));
    else if (tok_len) OP(CAT_ATTN, lcasr_attention_masked(q, k, v, cd, B, N, N, tok_len, H, Dh, a2, ai, stream));
    else OP(CAT_ATTN, lcasr_attention(q, k, v, cd, B, N, H, Dh, vt, p.Npad, a2, ai, stream));
    }
    LCASR_TRY(gemm(a2, L.out_w, M, d, d, nullptr, LCASR_ACT_NONE, x, 1.0f, x, LCASR_F32));
    // convolution module (convolution.py:103-124)
    LCASR_TRY(norm(L.conv_norm_w, L.conv_norm_b, nullptr, a));
    const bool fused_glu = !no_fuse && cd == LCASR_BF16 && gi != LCASR_GEMM_SIMT && !tok_len && L.pw1_w_glu && L.pw1_b_glu &&
                           (2 * d) % 64 == 0 && ((2 * d) % 256 == 0 || 2 * d > 512);
    if (fused_glu) {  // GLU in the epilogue of pointwise_conv1: the [M, 2d] pre-activation never reaches HBM
      begin(CAT_GEMM);
      LCASR_TRY(timed(CAT_GEMM, gemm_tc_launch_glu(a, L.pw1_w_glu, M, 2 * d, d, L.pw1_b_glu, q, cst)));
    } else {
      LCASR_TRY(gemm(a, L.pw1_w, M, 2 * d, d, L.pw1_b, LCASR_ACT_NONE, nullptr, 0.f, wide, cd));
      if (tok_len) OP(CAT_CONVMOD, lcasr_glu_masked(wide, cd, B, N, d, tok_len, q, stream));
      else OP(CAT_CONVMOD, lcasr_glu(wide, cd, M, d, q, stream));
    }
    OP(CAT_CONVMOD, lcasr_dwconv_brn_silu(q, cd, B, N, d, c.conv_kernel_size, L.dw_w, L.dw_b, L.brn_mean, L.brn_std, L.brn_w,
                                    L.brn_b, k, cd, stream));
    LCASR_TRY(gemm(k, L.pw2_w, M, d, d, L.pw2_b, LCASR_ACT_NONE, x, 1.0f, x, LCASR_F32));
    LCASR_TRY(ffn(L.ff2_norm_w, L.ff2_norm_b, L.ff2_fc1_w, L.ff2_fc1_b, L.ff2_fc2_w, L.ff2_fc2_b));
    const bool last = l == c.n_layers - 1;
    const bool chain = !no_fuse && c.decoder_norm && norm_chain_ok(d);
    if (chain && !last && c.self_conditioning) {
      // norm_out (-> residual stream, fp32) and decoder.norm (-> GEMM operand) in one pass over the row
      const float* cw[2] = {L.norm_out_w, w.dec_norm_w};
      const float* cb[2] = {L.norm_out_b, w.dec_norm_b};
      OP(CAT_NORM, lcasr_layernorm_chain(x, 2, cw, cb, M, d, c.norm_eps, c.norm_kind, 0, x, a, cd, stream));
    } else if (chain && last) {
      // end of the encoder: norm_out -> [decoder.norm (legasee double norm)] -> decoder.norm, only the operand is kept
      const float* cw[3] = {L.norm_out_w, w.dec_norm_w, w.dec_norm_w};
      const float* cb[3] = {L.norm_out_b, w.dec_norm_b, w.dec_norm_b};
      OP(CAT_NORM, lcasr_layernorm_chain(x, c.legasee_double_norm ? 3 : 2, cw, cb, M, d, c.norm_eps, c.norm_kind, 0, nullptr, a,
                                         cd, stream));
      final_operand_ready = true;
    } else {
      LCASR_TRY(norm(L.norm_out_w, L.norm_out_b, x, nullptr));
    }
    if (l != c.n_layers - 1 && c.self_conditioning) {  // sconformer_xl.py:241-243
      if (chain) { /* operand produced by the chained norm above */ }
      else if (c.decoder_norm) LCASR_TRY(norm(w.dec_norm_w, w.dec_norm_b, nullptr, a));
      else OP(CAT_NORM, lcasr_cast_f32(x, M * d, a, cd, stream));
      LCASR_TRY(gemm(a, w.dec_ff_w, M, V1, d, w.dec_ff_b, LCASR_ACT_NONE, nullptr, 0.f, wide, cd));
      OP(CAT_SOFTMAX, lcasr_softmax(wide, cd, M, V1, wide, cd, stream));
      LCASR_TRY(gemm(wide, w.dec_rep_w, M, d, V1, w.dec_rep_b, LCASR_ACT_NONE, x, 1.0f, x, LCASR_F32));
    }
  }
  if (!final_operand_ready) {
    if (c.legasee_double_norm && c.decoder_norm) LCASR_TRY(norm(w.dec_norm_w, w.dec_norm_b, x, nullptr));
    if (c.decoder_norm) LCASR_TRY(norm(w.dec_norm_w, w.dec_norm_b, nullptr, a));
    else OP(CAT_NORM, lcasr_cast_f32(x, M * d, a, cd, stream));
  }
  LCASR_TRY(gemm(a, w.dec_ff_w, M, V1, d, w.dec_ff_b, LCASR_ACT_NONE, nullptr, 0.f, out, LCASR_F32));
  if (!return_logits) OP(CAT_SOFTMAX, lcasr_log_softmax_argmax(out, M, V1, argmax, stream));
#undef OP
  return 0;
}

extern "C" int lcasr_model_set_timing(lcasr_model* m, int enable) {
  LCASR_CHECK_ARG(m, "model_set_timing: NULL model");
  m->timing = enable != 0;
  m->spans.clear();
  m->ev_used = 0;
  return 0;
}

// Synchronises the device, sums the recorded spans per category and resets the recorder.
extern "C" int lcasr_model_get_timing(lcasr_model* m, float* ms_by_cat, int32_t* launches_by_cat, int ncat) {
  LCASR_CHECK_ARG(m && ms_by_cat && launches_by_cat && ncat >= CAT_COUNT, "model_get_timing: bad arguments");
  LCASR_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < ncat; ++i) { ms_by_cat[i] = 0.f; launches_by_cat[i] = 0; }
  for (const auto& sp : m->spans) {
    if (sp.e1 == (size_t)-1) continue;
    float ms = 0.f;
    LCASR_CUDA(cudaEventElapsedTime(&ms, m->ev_pool[sp.e0], m->ev_pool[sp.e1]));
    ms_by_cat[sp.cat] += ms;
    launches_by_cat[sp.cat] += 1;
  }
  m->spans.clear();
  m->ev_used = 0;
  return 0;
}

extern "C" int lcasr_model_transcribe_host(lcasr_model* m, const float* spec_host, int B, int64_t T,
                                           int32_t* tokens_host, int32_t* n_tokens_host, float* logp_dev,
                                           void* workspace, int64_t workspace_bytes, void* stream) {
  LCASR_CHECK_ARG(m && spec_host && tokens_host && n_tokens_host && logp_dev && workspace,
                  "transcribe_host: NULL argument (logp_dev must hold B*N*num_classes floats + staging, see docs)");
  const lcasr_config& c = m->cfg;
  const int64_t N = lcasr_out_length(T);
  cudaStream_t st = (cudaStream_t)stream;
  // staging carved from the END of the caller's workspace: spec [B,F,T] fp32, argmax/tokens [B,N] i32, counts [B]
  const size_t need_fwd = (size_t)lcasr_model_workspace_bytes(m, B, T);
  const size_t spec_bytes = al((size_t)B * c.feat_in * T * 4), tok_bytes = al((size_t)B * N * 4), cnt_bytes = al((size_t)B * 4);
  LCASR_CHECK_ARG((size_t)workspace_bytes >= need_fwd + spec_bytes + 2 * tok_bytes + cnt_bytes,
                  "transcribe_host: workspace %lld < %lld bytes", (long long)workspace_bytes,
                  (long long)(need_fwd + spec_bytes + 2 * tok_bytes + cnt_bytes));
  char* ws = (char*)workspace;
  float* spec_dev = (float*)(ws + need_fwd);
  int32_t* am = (int32_t*)(ws + need_fwd + spec_bytes);
  int32_t* tok = (int32_t*)(ws + need_fwd + spec_bytes + tok_bytes);
  int32_t* cnt = (int32_t*)(ws + need_fwd + spec_bytes + 2 * tok_bytes);
  LCASR_CUDA(cudaMemcpyAsync(spec_dev, spec_host, (size_t)B * c.feat_in * T * 4, cudaMemcpyHostToDevice, st));
  LCASR_TRY(lcasr_model_forward(m, spec_dev, B, T, logp_dev, am, 0, workspace, (int64_t)need_fwd, stream));
  LCASR_TRY(lcasr_greedy_collapse(am, B, N, nullptr, c.num_classes - 1, tok, cnt, stream));
  LCASR_CUDA(cudaMemcpyAsync(tokens_host, tok, (size_t)B * N * 4, cudaMemcpyDeviceToHost, st));
  LCASR_CUDA(cudaMemcpyAsync(n_tokens_host, cnt, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  LCASR_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int64_t lcasr_model_transcribe_workspace_bytes(const lcasr_model* m, int B, int64_t T) {
  if (!m || B <= 0 || T <= 0) return -1;
  const int64_t N = lcasr_out_length(T);
  return lcasr_model_workspace_bytes(m, B, T) + (int64_t)al((size_t)B * m->cfg.feat_in * T * 4) +
         2 * (int64_t)al((size_t)B * N * 4) + (int64_t)al((size_t)B * 4);
}
