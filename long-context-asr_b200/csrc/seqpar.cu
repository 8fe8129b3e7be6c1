// Sequence-parallel forward of SCConformerXL (BASELINE configs 3 and 4): ONE recording, its N tokens split
// into P contiguous blocks, one block per rank (one process per GPU), driven entirely from C++ (no Python per op).
//
// What crosses ranks — everything else is token-local:
//   attention   K/V blocks travel over NVLink through NCCL on a side stream (one in-place all-gather pair per layer, or
//               ncclSend/ncclRecv pairs in ring order: step s sends the own block to rank r+s and receives the block of
//               rank r-s — see LCASR_SP_KV_MODE below) while the attention kernel already works on the own block.  Each
//               (query block, key block) launch produces an exact partial result (O_s / l_s in fp32 + log2-domain
//               log-sum-exp) and attn_merge_kernel combines the P partials: softmax over the union of the key blocks, no
//               approximation, independent launches (so consecutive blocks overlap on two streams and the 1.3-wave grids
//               of a 2048-token block fill each other's tails);
//   conv module the k=9 depthwise conv needs (k-1)/2 rows of the post-GLU tensor from each NEIGHBOUR: two small
//               ncclSend/ncclRecv pairs per layer (zeros at the true sequence ends);
//   decode      per-frame argmax ids are exchanged so that the greedy collapse sees the seams.
// The 8x subsampling needs no communication: a rank reads its slice of the spectrogram plus 8 frames (one token) of
// left context and drops the one contaminated token (proved in tests/test_seqpar_host.py).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 PyTorch already loaded): the library itself has no link-time
// dependency on it and single-GPU users never touch it.  `lcasr_model_forward_seqpar_emulated` runs the same per-rank
// phases for all P ranks in ONE process on one GPU with device-to-device copies in place of the transfers — the
// parity test of the single-GPU test box.
#include "model_internal.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <new>
#include <string>
#include <cstdlib>
#include <climits>

namespace lcasr {
int attn_tc_launch(const void* q, const void* k, const void* v, int B, int64_t N, int64_t Nk, const int32_t* kv_len, int H,
                   int Dh, int v_transposed, int64_t Npad, void* out, float* lse, cudaStream_t st, int wl = -1, int wr = -1,
                   float* out32 = nullptr, int64_t ldq = 0, int64_t ldkv = 0);
int attn_simt_launch(const void* q, const void* k, const void* v, int dtype, int B, int64_t N, int64_t Nk, const int32_t* kv_len,
                     int H, int Dh, int v_transposed, int64_t Npad, void* out, cudaStream_t st, int wl = -1, int wr = -1);
}  // namespace lcasr

using namespace lcasr;

// ------------------------------------------------------------------------------------------------
// exact merge of per-key-block attention partials
//   parts [P][n, H*Dh] fp32 (normalised per block), lses [P][H, n] fp32 (log2-domain) -> out [n, H*Dh]
// ------------------------------------------------------------------------------------------------
namespace {

template <typename TOut>
__global__ void attn_merge_kernel(const float* __restrict__ parts, const float* __restrict__ lses, int P, int64_t n, int H,
                                  int Dh, int64_t part_stride, int64_t lse_stride, TOut* __restrict__ out) {
  const int d = H * Dh;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (row, group of 8 columns)
  const int groups = d / 8;
  if (idx >= n * groups) return;
  const int64_t row = idx / groups;
  const int c0 = (int)(idx % groups) * 8;
  const int h = c0 / Dh;
  float mx = -INFINITY;
  for (int s = 0; s < P; ++s) mx = fmaxf(mx, lses[s * lse_stride + (int64_t)h * n + row]);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float wsum = 0.f;
  for (int s = 0; s < P; ++s) {
    const float w = exp2f(lses[s * lse_stride + (int64_t)h * n + row] - mx);
    wsum += w;
    float v[8];
    Vec8<float>::load(parts + s * part_stride + row * d + c0, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, v[i], acc[i]);
  }
  const float inv = 1.0f / wsum;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= inv;
  Vec8<TOut>::store(out + row * d + c0, acc);
}

int merge_launch(const float* parts, const float* lses, int P, int64_t n, int H, int Dh, int64_t part_stride, int64_t lse_stride,
                 void* out, int out_dtype, cudaStream_t st) {
  const int64_t total = n * (H * Dh / 8);
  const int threads = 256;
  const unsigned blocks = (unsigned)ceil_div(total, threads);
  if (out_dtype == LCASR_BF16)
    attn_merge_kernel<bf16><<<blocks, threads, 0, st>>>(parts, lses, P, n, H, Dh, part_stride, lse_stride, (bf16*)out);
  else
    attn_merge_kernel<float><<<blocks, threads, 0, st>>>(parts, lses, P, n, H, Dh, part_stride, lse_stride, (float*)out);
  LCASR_LAUNCH_CHECK();
  return 0;
}

}  // namespace

namespace lcasr {
int attn_merge_launch(const float* parts, const float* lses, int P, int64_t n, int H, int Dh, int64_t part_stride,
                      int64_t lse_stride, void* out, int out_dtype, cudaStream_t st) {
  return merge_launch(parts, lses, P, n, H, Dh, part_stride, lse_stride, out, out_dtype, st);
}
}  // namespace lcasr

extern "C" int lcasr_attention_partial(const void* q, const void* k, const void* v, int B, int64_t Nq, int64_t Nk, int H, int Dh,
                                       float* out32, float* lse, void* stream) {
  LCASR_CHECK_ARG(q && k && v && out32 && lse, "attention_partial: NULL operand");
  LCASR_CHECK_ARG(B > 0 && Nq > 0 && Nk > 0 && H > 0 && Dh > 0, "attention_partial: bad shape");
  return attn_tc_launch(q, k, v, B, Nq, Nk, nullptr, H, Dh, 0, 0, nullptr, lse, (cudaStream_t)stream, -1, -1, out32);
}

extern "C" int lcasr_attention_merge(const float* parts, const float* lses, int P, int64_t rows, int H, int Dh, void* out,
                                     int out_dtype, void* stream) {
  LCASR_CHECK_ARG(parts && lses && out && P > 0 && rows > 0 && H > 0 && Dh % 8 == 0, "attention_merge: bad arguments");
  LCASR_CHECK_ARG(out_dtype == LCASR_BF16 || out_dtype == LCASR_F32, "attention_merge: bad output dtype");
  return merge_launch(parts, lses, P, rows, H, Dh, rows * H * Dh, (int64_t)H * rows, out, out_dtype, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ------------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;

  bool load(const char* path) {
    if (handle) return true;
    const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);  // the copy PyTorch already mapped, if any
      if (!handle) handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) { error = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?"); return false; }
#define LCASR_NCCL_SYM(field, sym)                                         \
  *(void**)(&field) = dlsym(handle, sym);                                  \
  if (!field) { error = std::string("NCCL symbol missing: ") + sym; handle = nullptr; return false; }
    LCASR_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    LCASR_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    LCASR_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    LCASR_NCCL_SYM(Send, "ncclSend");
    LCASR_NCCL_SYM(Recv, "ncclRecv");
    LCASR_NCCL_SYM(AllGather, "ncclAllGather");
    LCASR_NCCL_SYM(GroupStart, "ncclGroupStart");
    LCASR_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    LCASR_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef LCASR_NCCL_SYM
    return true;
  }
};

NcclApi g_nccl;

#define LCASR_NCCL(expr)                                                                          \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess)                                                                        \
      return set_error(LCASR_E_CUDA, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

}  // namespace

struct lcasr_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t cs = nullptr;   // communication stream (K/V blocks, halos)
  cudaStream_t aux = nullptr;  // second attention stream
  cudaStream_t aux2 = nullptr, aux3 = nullptr;  // further attention streams: the pieces of one layer run concurrently
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  cudaEvent_t next_event() {
    if (ev_used == ev.size()) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      ev.push_back(e);
    }
    return ev[ev_used++];
  }
};

extern "C" int lcasr_comm_unique_id(char* id_out_host, const char* nccl_path_or_null) {
  LCASR_CHECK_ARG(id_out_host, "comm_unique_id: NULL");
  if (!g_nccl.load(nccl_path_or_null)) return set_error(LCASR_E_UNSUPPORTED, "%s", g_nccl.error.c_str());
  ncclUniqueId id;
  LCASR_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out_host, id.internal, NCCL_UNIQUE_ID_BYTES);
  return 0;
}

extern "C" int lcasr_comm_create(const char* id_host, int rank, int world, const char* nccl_path_or_null, lcasr_comm** out) {
  LCASR_CHECK_ARG(id_host && out && world >= 1 && rank >= 0 && rank < world, "comm_create: bad arguments");
  if (!g_nccl.load(nccl_path_or_null)) return set_error(LCASR_E_UNSUPPORTED, "%s", g_nccl.error.c_str());
  lcasr_comm* c = new (std::nothrow) lcasr_comm();
  if (!c) return set_error(LCASR_E_NOMEM, "comm_create: out of host memory");
  c->rank = rank; c->world = world;
  ncclUniqueId id;
  memcpy(id.internal, id_host, NCCL_UNIQUE_ID_BYTES);
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return set_error(LCASR_E_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
  }
  LCASR_CUDA(cudaStreamCreateWithFlags(&c->cs, cudaStreamNonBlocking));
  LCASR_CUDA(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
  LCASR_CUDA(cudaStreamCreateWithFlags(&c->aux2, cudaStreamNonBlocking));
  LCASR_CUDA(cudaStreamCreateWithFlags(&c->aux3, cudaStreamNonBlocking));
  *out = c;
  return 0;
}

extern "C" void lcasr_comm_destroy(lcasr_comm* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  if (c->cs) cudaStreamDestroy(c->cs);
  if (c->aux) cudaStreamDestroy(c->aux);
  if (c->aux2) cudaStreamDestroy(c->aux2);
  if (c->aux3) cudaStreamDestroy(c->aux3);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
}

// ------------------------------------------------------------------------------------------------
// per-rank plan and phases
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxOtherPieces = 8;  // partial-attention launches over the keys of the other ranks (+ 1 for the own block)

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct SpPlan {
  int P = 1, r = 0;
  int64_t N = 0;                 // tokens of the whole recording
  std::vector<int64_t> start, cnt;
  int64_t n = 0, s0 = 0, nmax = 0;  // own block
  int drop = 0;                  // context tokens in front of the own slice (1 unless the rank owns token 0)
  int halo = 0;
  int64_t Ts = 0, T1 = 0, T2 = 0, Nl = 0;  // slice frames and the subsampling pyramid of the slice (Nl = n + drop)
  int F1 = 0, F2 = 0, F3 = 0;
  size_t off_spec, off_x, off_cos, off_sin, off_a, off_wide, off_q, off_K, off_V, off_parts, off_lse, off_ext, off_cb, off_am;
  size_t off_s1, off_s2a, off_s2b, off_s3a, off_s3b;
  size_t total = 0;
};

int make_sp_plan(const lcasr_config& c, int P, int r, int64_t T, SpPlan* out) {
  LCASR_CHECK_ARG(P >= 1 && r >= 0 && r < P, "seqpar: bad world / rank");
  LCASR_CHECK_ARG(T > 0 && T % 8 == 0, "seqpar: the frame count must be a positive multiple of 8 (got %lld)", (long long)T);
  SpPlan p;
  p.P = P; p.r = r; p.N = T / 8;
  p.halo = (c.conv_kernel_size - 1) / 2;
  const int64_t base = p.N / P, rem = p.N % P;  // contiguous blocks, sizes differing by at most one (larger first)
  LCASR_CHECK_ARG(base >= p.halo && base >= 1, "seqpar: %lld tokens are too few for %d ranks", (long long)p.N, P);
  p.start.resize(P); p.cnt.resize(P);
  int64_t s = 0;
  for (int i = 0; i < P; ++i) { p.cnt[i] = base + (i < rem ? 1 : 0); p.start[i] = s; s += p.cnt[i]; }
  p.n = p.cnt[r]; p.s0 = p.start[r]; p.nmax = base + (rem ? 1 : 0);
  p.drop = p.s0 > 0 ? 1 : 0;
  p.Nl = p.n + p.drop;
  p.Ts = 8 * p.Nl;
  p.T1 = (p.Ts - 1) / 2 + 1; p.T2 = (p.T1 - 1) / 2 + 1;
  p.F1 = (c.feat_in - 1) / 2 + 1; p.F2 = (p.F1 - 1) / 2 + 1; p.F3 = (p.F2 - 1) / 2 + 1;
  const size_t e = dtype_size(c.compute_dtype);
  const int d = c.d_model, C = c.conv_channels, V1 = c.num_classes;
  size_t o = 0;
  p.off_spec = o; o += al256((size_t)c.feat_in * p.Ts * 4);
  p.off_x = o; o += al256((size_t)p.Nl * d * 4);
  p.off_cos = o; o += al256((size_t)p.n * (c.head_dim / 2) * 4);
  p.off_sin = o; o += al256((size_t)p.n * (c.head_dim / 2) * 4);
  p.off_K = o; o += al256((size_t)p.N * d * e);
  p.off_V = o; o += al256((size_t)p.N * d * e);
  const int slots = P > kMaxOtherPieces + 1 ? P : kMaxOtherPieces + 1;  // per-step mode: P partials; gathered mode: own + pieces
  p.off_parts = o; o += al256((size_t)slots * p.n * d * 4);
  p.off_lse = o; o += al256((size_t)slots * c.n_heads * p.n * 4);
  p.off_ext = o; o += al256((size_t)(p.n + 2 * p.halo) * d * e);
  p.off_cb = o; o += al256((size_t)(p.n + 2 * p.halo) * d * e);
  p.off_am = o; o += al256((size_t)p.N * 4);
  const size_t scratch0 = o;
  p.off_a = o; o += al256((size_t)p.Nl * d * e);
  size_t wide_cols = (size_t)4 * d;
  if ((size_t)V1 > wide_cols) wide_cols = V1;
  p.off_wide = o; o += al256((size_t)p.Nl * wide_cols * e);
  p.off_q = o; o += al256((size_t)p.Nl * d * e);
  const size_t layer_end = o;
  o = scratch0;  // subsampling scratch aliases the layer scratch (dead once x exists)
  p.off_s1 = o; o += al256((size_t)p.T1 * p.F1 * C * e);
  p.off_s2a = o; o += al256((size_t)p.T2 * p.F2 * C * e);
  p.off_s2b = o; o += al256((size_t)p.T2 * p.F2 * C * e);
  p.off_s3a = o; o += al256((size_t)p.Nl * p.F3 * C * e);
  p.off_s3b = o; o += al256((size_t)p.Nl * p.F3 * C * e);
  p.total = o > layer_end ? o : layer_end;
  *out = p;
  return 0;
}

// One rank's state for one forward: pointers into its workspace + the phases between communication points.
struct SpRank {
  lcasr_model* m;
  SpPlan p;
  char* ws;
  cudaStream_t st;
  float* out;        // [n, V1] fp32: this rank's rows of the result
  int32_t* am_full;  // [N] int32 (this rank's copy of the gathered argmax ids; own block written by post())

  const lcasr_config& c() const { return m->cfg; }
  size_t e() const { return dtype_size(m->cfg.compute_dtype); }
  float* x() const { return (float*)(ws + p.off_x) + (size_t)p.drop * m->cfg.d_model; }
  char* K(int64_t tok) const { return ws + p.off_K + (size_t)tok * m->cfg.d_model * e(); }
  char* V(int64_t tok) const { return ws + p.off_V + (size_t)tok * m->cfg.d_model * e(); }
  char* ext(int64_t row) const { return ws + p.off_ext + (size_t)row * m->cfg.d_model * e(); }
  size_t block_bytes(int j) const { return (size_t)p.cnt[j] * m->cfg.d_model * e(); }
  size_t halo_bytes() const { return (size_t)p.halo * m->cfg.d_model * e(); }

  int gemm(const void* A, const void* W, int64_t rows, int nn, int kk, const float* bias, int act, const float* resid,
           float alpha, void* o, int odt) {
    return lcasr_gemm(A, W, c().compute_dtype, rows, nn, kk, bias, act, resid, alpha, o, odt, m->gemm_impl, st);
  }
  int norm(const float* nw, const float* nb, float* o32, void* olo) {
    return lcasr_layernorm(x(), nw, nb, p.n, c().d_model, c().norm_eps, c().norm_kind, o32, olo, c().compute_dtype, st);
  }
  int ffn(const float* nw, const float* nb, const void* fc1, const float* b1, const void* fc2, const float* b2) {
    const int d = c().d_model;
    void* a = ws + p.off_a; void* wide = ws + p.off_wide;
    LCASR_TRY(norm(nw, nb, nullptr, a));
    LCASR_TRY(gemm(a, fc1, p.n, 4 * d, d, b1, LCASR_ACT_GELU_TANH, nullptr, 0.f, wide, c().compute_dtype));
    LCASR_TRY(gemm(wide, fc2, p.n, d, 4 * d, b2, LCASR_ACT_NONE, x(), 0.5f, x(), LCASR_F32));
    return 0;
  }

  // slice of the spectrogram (+ one token of left context) -> subsampling -> residual stream; rotary tables of the block
  int pre(const float* spec_full, int64_t T) {
    const lcasr_weights& w = m->w;
    const int cd = c().compute_dtype, C = c().conv_channels, d = c().d_model, F = c().feat_in;
    float* spec = (float*)(ws + p.off_spec);
    const int64_t f0 = 8 * (p.s0 - p.drop);
    LCASR_CUDA(cudaMemcpy2DAsync(spec, (size_t)p.Ts * 4, spec_full + f0, (size_t)T * 4, (size_t)p.Ts * 4, F,
                                 cudaMemcpyDeviceToDevice, st));
    void* s1 = ws + p.off_s1; void* s2a = ws + p.off_s2a; void* s2b = ws + p.off_s2b;
    void* s3a = ws + p.off_s3a; void* s3b = ws + p.off_s3b;
    if (cd == LCASR_BF16 && C % 64 == 0) {
      LCASR_TRY(lcasr_subsample_conv0_dw(spec, w.conv0_w, w.conv0_b, w.dw1_w, w.dw1_b, 1, F, p.Ts, C, s2a, st));
    } else {
      LCASR_TRY(lcasr_subsample_conv0(spec, w.conv0_w, w.conv0_b, 1, F, p.Ts, C, s1, cd, st));
      LCASR_TRY(lcasr_subsample_dwconv(s1, cd, w.dw1_w, w.dw1_b, 1, p.T1, p.F1, C, s2a, st));
    }
    LCASR_TRY(gemm(s2a, w.pw1_w, p.T2 * p.F2, C, C, w.pw1_b, LCASR_ACT_SILU, nullptr, 0.f, s2b, cd));
    LCASR_TRY(lcasr_subsample_dwconv(s2b, cd, w.dw2_w, w.dw2_b, 1, p.T2, p.F2, C, s3a, st));
    LCASR_TRY(gemm(s3a, w.pw2_w, p.Nl * p.F3, C, C, w.pw2_b, LCASR_ACT_SILU, nullptr, 0.f, s3b, cd));
    LCASR_TRY(gemm(s3b, w.sub_out_w, p.Nl, d, p.F3 * C, nullptr, LCASR_ACT_NONE, nullptr, 0.f, ws + p.off_x, LCASR_F32));
    if (c().use_rotary)
      LCASR_TRY(lcasr_rope_table(w.inv_freq, c().rotary_interp, p.s0, p.n, c().head_dim / 2, (float*)(ws + p.off_cos),
                                 (float*)(ws + p.off_sin), st));
    // zero halos at the true ends of the recording (never overwritten: only interior neighbours send)
    if (p.r == 0) LCASR_CUDA(cudaMemsetAsync(ext(0), 0, halo_bytes(), st));
    if (p.r == p.P - 1) LCASR_CUDA(cudaMemsetAsync(ext(p.halo + p.n), 0, halo_bytes(), st));
    return 0;
  }

  // first half-FFN, attention norm, qkv projection, rotary: q and the own K/V block (in place in the gathered K/V)
  int layer_a(int l) {
    const lcasr_layer_weights& L = m->layers[l];
    const int d = c().d_model, H = c().n_heads, Dh = c().head_dim, cd = c().compute_dtype;
    void* a = ws + p.off_a; void* wide = ws + p.off_wide;
    LCASR_TRY(ffn(L.ff1_norm_w, L.ff1_norm_b, L.ff1_fc1_w, L.ff1_fc1_b, L.ff1_fc2_w, L.ff1_fc2_b));
    LCASR_TRY(norm(L.attn_norm_w, L.attn_norm_b, nullptr, a));
    LCASR_TRY(gemm(a, L.qkv_w, p.n, 3 * d, d, nullptr, LCASR_ACT_NONE, nullptr, 0.f, wide, cd));
    LCASR_TRY(lcasr_rope_split(wide, cd, 1, p.n, H, Dh, c().use_rotary ? (float*)(ws + p.off_cos) : nullptr,
                               c().use_rotary ? (float*)(ws + p.off_sin) : nullptr, ws + p.off_q, K(p.s0), V(p.s0), 0, 0, st));
    return 0;
  }

  bool partial_mode() const { return c().compute_dtype == LCASR_BF16 && m->attn_impl != LCASR_ATTN_SIMT; }

  int n_parts = 1;  // partial results the merge of this layer combines

  // bf16: own queries x keys [tok0, tok0 + ntok) -> partial slot `slot` (may run on another stream)
  int attn_range(int slot, int64_t tok0, int64_t ntok, cudaStream_t stream) {
    const int d = c().d_model, H = c().n_heads, Dh = c().head_dim;
    float* part = (float*)(ws + p.off_parts) + (size_t)slot * p.n * d;
    float* lse = (float*)(ws + p.off_lse) + (size_t)slot * H * p.n;
    return attn_tc_launch(ws + p.off_q, K(tok0), V(tok0), 1, p.n, ntok, nullptr, H, Dh, 0, 0, nullptr, lse, stream, -1, -1, part);
  }
  int attn_block(int slot, int j, cudaStream_t stream) { return attn_range(slot, p.start[j], p.cnt[j], stream); }
  // Everything that is not the own block: the keys in front of and behind it, cut into pieces of equal length so that the
  // CTAs of all pieces (launched on two streams, they fill each other's tails) tile the 148 SMs with little left over.
  // One launch per side is 192 CTAs of up to 112 key tiles at 8 ranks of the 20-minute context: 2 waves for 1.3 waves of
  // work; three pieces are 576 CTAs of 38 tiles: 4 waves x 38 = 152 tile-times against an ideal of 145.
  int attn_others(cudaStream_t s_a, cudaStream_t s_b, cudaStream_t s_c = nullptr, cudaStream_t s_d = nullptr) {
    static const int max_pieces_env = getenv("LCASR_SP_MAX_PIECES") ? atoi(getenv("LCASR_SP_MAX_PIECES")) : kMaxOtherPieces;
    const int max_pieces = max_pieces_env < 2 ? 2 : (max_pieces_env > kMaxOtherPieces ? kMaxOtherPieces : max_pieces_env);
    cudaStream_t streams[4] = {s_a, s_b, s_c ? s_c : s_a, s_d ? s_d : s_b};
    const int64_t ranges[2][2] = {{0, p.s0}, {p.s0 + p.n, p.N}};
    const int64_t tiles_total = ceil_div(ranges[0][1] - ranges[0][0], 128) + ceil_div(ranges[1][1] - ranges[1][0], 128);
    int slot = 1;
    if (tiles_total > 0) {
      const int64_t ctas = ceil_div(p.n, 256) * c().n_heads;  // CTAs per launch (two 128-row query tiles each)
      int64_t best_c = tiles_total, best_cost = INT64_MAX;
      for (int pieces = 1; pieces <= max_pieces - 1; ++pieces) {
        const int64_t cc = ceil_div(tiles_total, pieces);  // tiles per piece
        const int64_t np = ceil_div(ceil_div(ranges[0][1] - ranges[0][0], 128), cc) + ceil_div(ceil_div(ranges[1][1] - ranges[1][0], 128), cc);
        if (np > max_pieces) continue;
        const int64_t cost = ceil_div(ctas * np, kNumSMs) * (cc + 2);  // waves x (tiles per CTA + fixed cost of a CTA)
        if (cost < best_cost) { best_cost = cost; best_c = cc; }
      }
      int k = 0;
      for (int side = 0; side < 2; ++side)
        for (int64_t t0 = ranges[side][0]; t0 < ranges[side][1]; t0 += best_c * 128) {
          const int64_t cnt = ranges[side][1] - t0 < best_c * 128 ? ranges[side][1] - t0 : best_c * 128;
          LCASR_TRY(attn_range(slot++, t0, cnt, streams[k++ & 3]));
        }
    }
    n_parts = slot;
    return 0;
  }

  // merge (bf16) or the single fp32 launch over the gathered K/V, out-projection, conv-module front: norm, pw1, GLU
  int layer_b(int l) {
    const lcasr_layer_weights& L = m->layers[l];
    const int d = c().d_model, H = c().n_heads, Dh = c().head_dim, cd = c().compute_dtype;
    void* a = ws + p.off_a; void* wide = ws + p.off_wide;
    if (partial_mode()) {
      LCASR_TRY(merge_launch((float*)(ws + p.off_parts), (float*)(ws + p.off_lse), n_parts, p.n, H, Dh, (int64_t)p.n * d,
                             (int64_t)H * p.n, a, cd, st));
    } else {  // fp32 parity mode: one launch over all keys in global order — bit-identical to the single-GPU forward
      if (cd == LCASR_BF16) LCASR_TRY(attn_tc_launch(ws + p.off_q, K(0), V(0), 1, p.n, p.N, nullptr, H, Dh, 0, 0, a, nullptr, st));
      else LCASR_TRY(attn_simt_launch(ws + p.off_q, K(0), V(0), cd, 1, p.n, p.N, nullptr, H, Dh, 0, 0, a, st));
    }
    LCASR_TRY(gemm(a, L.out_w, p.n, d, d, nullptr, LCASR_ACT_NONE, x(), 1.0f, x(), LCASR_F32));
    LCASR_TRY(norm(L.conv_norm_w, L.conv_norm_b, nullptr, a));
    LCASR_TRY(gemm(a, L.pw1_w, p.n, 2 * d, d, L.pw1_b, LCASR_ACT_NONE, nullptr, 0.f, wide, cd));
    LCASR_TRY(lcasr_glu(wide, cd, p.n, d, ext(p.halo), st));
    return 0;
  }

  // depthwise conv over [left halo | own rows | right halo], pw2, second half-FFN, norm_out, self-conditioning
  int layer_c(int l) {
    const lcasr_layer_weights& L = m->layers[l];
    const lcasr_weights& w = m->w;
    const int d = c().d_model, V1 = c().num_classes, cd = c().compute_dtype;
    void* a = ws + p.off_a; void* wide = ws + p.off_wide;
    char* cb = ws + p.off_cb;
    LCASR_TRY(lcasr_dwconv_brn_silu(ext(0), cd, 1, p.n + 2 * p.halo, d, c().conv_kernel_size, L.dw_w, L.dw_b, L.brn_mean, L.brn_std,
                                    L.brn_w, L.brn_b, cb, cd, st));
    LCASR_TRY(gemm(cb + halo_bytes(), L.pw2_w, p.n, d, d, L.pw2_b, LCASR_ACT_NONE, x(), 1.0f, x(), LCASR_F32));
    LCASR_TRY(ffn(L.ff2_norm_w, L.ff2_norm_b, L.ff2_fc1_w, L.ff2_fc1_b, L.ff2_fc2_w, L.ff2_fc2_b));
    LCASR_TRY(norm(L.norm_out_w, L.norm_out_b, x(), nullptr));
    if (l != c().n_layers - 1 && c().self_conditioning) {
      if (c().decoder_norm) LCASR_TRY(norm(w.dec_norm_w, w.dec_norm_b, nullptr, a));
      else LCASR_TRY(lcasr_cast_f32(x(), p.n * d, a, cd, st));
      LCASR_TRY(gemm(a, w.dec_ff_w, p.n, V1, d, w.dec_ff_b, LCASR_ACT_NONE, nullptr, 0.f, wide, cd));
      LCASR_TRY(lcasr_softmax(wide, cd, p.n, V1, wide, cd, st));
      LCASR_TRY(gemm(wide, w.dec_rep_w, p.n, d, V1, w.dec_rep_b, LCASR_ACT_NONE, x(), 1.0f, x(), LCASR_F32));
    }
    return 0;
  }

  int post(int return_logits) {
    const lcasr_weights& w = m->w;
    const int d = c().d_model, V1 = c().num_classes, cd = c().compute_dtype;
    void* a = ws + p.off_a;
    if (c().legasee_double_norm && c().decoder_norm) LCASR_TRY(norm(w.dec_norm_w, w.dec_norm_b, x(), nullptr));
    if (c().decoder_norm) LCASR_TRY(norm(w.dec_norm_w, w.dec_norm_b, nullptr, a));
    else LCASR_TRY(lcasr_cast_f32(x(), p.n * d, a, cd, st));
    LCASR_TRY(gemm(a, w.dec_ff_w, p.n, V1, d, w.dec_ff_b, LCASR_ACT_NONE, nullptr, 0.f, out, LCASR_F32));
    if (!return_logits) LCASR_TRY(lcasr_log_softmax_argmax(out, p.n, V1, am_full ? am_full + p.s0 : nullptr, st));
    return 0;
  }
};

int check_model_for_sp(const lcasr_model* m) {
  LCASR_CHECK_ARG(m->cfg.attn_window_left < 0 && m->cfg.attn_window_right < 0,
                  "seqpar: windowed attention is a single-GPU evaluation mode");
  return 0;
}

}  // namespace

extern "C" int64_t lcasr_model_seqpar_workspace_bytes(const lcasr_model* m, int world, int rank, int64_t T) {
  if (!m) return -1;
  SpPlan p;
  if (make_sp_plan(m->cfg, world, rank, T, &p) != 0) return -1;
  return (int64_t)p.total;
}

extern "C" int lcasr_model_seqpar_block(const lcasr_model* m, int world, int rank, int64_t T, int64_t* start_tok, int64_t* n_tok) {
  LCASR_CHECK_ARG(m && start_tok && n_tok, "seqpar_block: NULL argument");
  SpPlan p;
  LCASR_TRY(make_sp_plan(m->cfg, world, rank, T, &p));
  *start_tok = p.s0; *n_tok = p.n;
  return 0;
}

// One rank of the sequence-parallel forward.  spec_full [1, feat_in, T] fp32 (device; only this rank's slice is read),
// out_local [n_r, num_classes] fp32 (log-probs, or logits), argmax_full [N] int32 or NULL: per-frame argmax ids of the
// WHOLE recording on every rank (exchanged at the end).  Asynchronous on `stream` (+ the communicator's side streams,
// joined back into `stream` before return).
extern "C" int lcasr_model_forward_seqpar(lcasr_model* m, lcasr_comm* comm, const float* spec_full, int64_t T, float* out_local,
                                          int32_t* argmax_full, int return_logits, void* workspace, int64_t workspace_bytes,
                                          void* stream) {
  LCASR_CHECK_ARG(m && comm && spec_full && out_local && workspace, "forward_seqpar: NULL argument");
  LCASR_TRY(check_model_for_sp(m));
  SpRank R;
  R.m = m;
  LCASR_TRY(make_sp_plan(m->cfg, comm->world, comm->rank, T, &R.p));
  LCASR_CHECK_ARG((size_t)workspace_bytes >= R.p.total, "forward_seqpar: workspace %lld < required %lld bytes",
                  (long long)workspace_bytes, (long long)R.p.total);
  LCASR_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "forward_seqpar: workspace must be 256-byte aligned");
  R.ws = (char*)workspace; R.st = (cudaStream_t)stream; R.out = out_local;
  R.am_full = argmax_full ? argmax_full : (int32_t*)(R.ws + R.p.off_am);
  const SpPlan& p = R.p;
  const int P = p.P, r = p.r;
  cudaStream_t st = R.st, cs = comm->cs, aux = comm->aux;
  comm->ev_used = 0;
  const bool partial = R.partial_mode();

  LCASR_TRY(R.pre(spec_full, T));
  for (int l = 0; l < m->cfg.n_layers; ++l) {
    LCASR_TRY(R.layer_a(l));
    if (P > 1) {
      // the own K/V block is final (and, transitively, every attention launch of the previous layer has been issued
      // before this point on `st`, so peers' blocks may be overwritten by this layer's transfers)
      cudaEvent_t ev_kv = comm->next_event();
      LCASR_CUDA(cudaEventRecord(ev_kv, st));
      LCASR_CUDA(cudaStreamWaitEvent(cs, ev_kv, 0));
      LCASR_CUDA(cudaStreamWaitEvent(aux, ev_kv, 0));
      // How the other ranks' K/V blocks arrive (LCASR_SP_KV_MODE):
      //   2 (default when the blocks are equal): ONE in-place ncclAllGather pair per layer — NVSwitch-optimised, 2 NCCL
      //     kernels; the own block's attention overlaps it, the other blocks follow its completion;
      //   1: ONE group of all P-1 ncclSend/ncclRecv pairs (ring-ordered peers; NCCL runs them concurrently over its
      //     channels) — the form for unequal blocks;
      //   0: P-1 separate groups, one event each, so that attention on block r-s can start as soon as that block landed.
      //     Finest overlap on paper, but every group is a separate NCCL kernel over the one or two channels NCCL gives a peer
      //     pair: measured 136 us per step at 8 ranks (6 MB per step) — 2.2x instead of 5x+ for the 20-minute context.
      static const int kv_mode_env = getenv("LCASR_SP_KV_MODE") ? atoi(getenv("LCASR_SP_KV_MODE")) : -1;
      bool equal_blocks = true;
      for (int j = 1; j < P; ++j) equal_blocks = equal_blocks && p.cnt[j] == p.cnt[0];
      int kv_mode = kv_mode_env >= 0 ? kv_mode_env : 2;
      if (kv_mode == 2 && !equal_blocks) kv_mode = 1;
      std::vector<cudaEvent_t> ev_blk(P, nullptr);
      if (kv_mode == 0) {
        for (int s = 1; s < P; ++s) {  // ring order: step s sends to r+s and receives the block of r-s
          const int to = (r + s) % P, from = (r - s + P) % P;
          LCASR_NCCL(g_nccl.GroupStart());
          LCASR_NCCL(g_nccl.Send(R.K(p.s0), R.block_bytes(r), ncclChar, to, comm->comm, cs));
          LCASR_NCCL(g_nccl.Send(R.V(p.s0), R.block_bytes(r), ncclChar, to, comm->comm, cs));
          LCASR_NCCL(g_nccl.Recv(R.K(p.start[from]), R.block_bytes(from), ncclChar, from, comm->comm, cs));
          LCASR_NCCL(g_nccl.Recv(R.V(p.start[from]), R.block_bytes(from), ncclChar, from, comm->comm, cs));
          LCASR_NCCL(g_nccl.GroupEnd());
          ev_blk[s] = comm->next_event();
          LCASR_CUDA(cudaEventRecord(ev_blk[s], cs));
        }
      } else {
        LCASR_NCCL(g_nccl.GroupStart());
        if (kv_mode == 2) {  // in place: this rank's block already sits at its slot of the gathered tensors
          LCASR_NCCL(g_nccl.AllGather(R.K(p.s0), R.K(0), R.block_bytes(r), ncclChar, comm->comm, cs));
          LCASR_NCCL(g_nccl.AllGather(R.V(p.s0), R.V(0), R.block_bytes(r), ncclChar, comm->comm, cs));
        } else {
          for (int s = 1; s < P; ++s) {
            const int to = (r + s) % P, from = (r - s + P) % P;
            LCASR_NCCL(g_nccl.Send(R.K(p.s0), R.block_bytes(r), ncclChar, to, comm->comm, cs));
            LCASR_NCCL(g_nccl.Send(R.V(p.s0), R.block_bytes(r), ncclChar, to, comm->comm, cs));
            LCASR_NCCL(g_nccl.Recv(R.K(p.start[from]), R.block_bytes(from), ncclChar, from, comm->comm, cs));
            LCASR_NCCL(g_nccl.Recv(R.V(p.start[from]), R.block_bytes(from), ncclChar, from, comm->comm, cs));
          }
        }
        LCASR_NCCL(g_nccl.GroupEnd());
        cudaEvent_t ev_all = comm->next_event();
        LCASR_CUDA(cudaEventRecord(ev_all, cs));
        for (int s = 1; s < P; ++s) ev_blk[s] = ev_all;
      }
      if (partial) {
        LCASR_TRY(R.attn_block(0, r, st));  // the own block needs no transfer: it overlaps the exchange
        if (kv_mode == 0) {
          for (int s = 1; s < P; ++s) {
            cudaStream_t as = (s & 1) ? aux : st;  // independent partials: alternate streams so that tails overlap
            LCASR_CUDA(cudaStreamWaitEvent(as, ev_blk[s], 0));
            LCASR_TRY(R.attn_block(s, (r - s + P) % P, as));
          }
          R.n_parts = P;
        } else {  // everything arrives at once: two long launches (keys in front of / behind the own block) on two streams
          cudaStream_t side[3] = {aux, comm->aux2, comm->aux3};
          LCASR_CUDA(cudaStreamWaitEvent(st, ev_blk[1], 0));
          for (cudaStream_t ss : side) {
            LCASR_CUDA(cudaStreamWaitEvent(ss, ev_kv, 0));
            LCASR_CUDA(cudaStreamWaitEvent(ss, ev_blk[1], 0));
          }
          LCASR_TRY(R.attn_others(aux, comm->aux2, comm->aux3, st));
          for (int i = 1; i < 3; ++i) {  // join the extra streams
            cudaEvent_t ev_j = comm->next_event();
            LCASR_CUDA(cudaEventRecord(ev_j, side[i]));
            LCASR_CUDA(cudaStreamWaitEvent(st, ev_j, 0));
          }
        }
        cudaEvent_t ev_aux = comm->next_event();
        LCASR_CUDA(cudaEventRecord(ev_aux, aux));
        LCASR_CUDA(cudaStreamWaitEvent(st, ev_aux, 0));
      } else {
        LCASR_CUDA(cudaStreamWaitEvent(st, ev_blk[P - 1], 0));  // in-order side stream: the last block implies all
      }
    } else if (partial) {
      LCASR_TRY(R.attn_block(0, 0, st));
      R.n_parts = 1;
    }
    LCASR_TRY(R.layer_b(l));
    if (P > 1) {  // halo rows of the post-GLU tensor to / from the two neighbours
      cudaEvent_t ev_g = comm->next_event();
      LCASR_CUDA(cudaEventRecord(ev_g, st));
      LCASR_CUDA(cudaStreamWaitEvent(cs, ev_g, 0));
      LCASR_NCCL(g_nccl.GroupStart());
      if (r > 0) {
        LCASR_NCCL(g_nccl.Send(R.ext(p.halo), R.halo_bytes(), ncclChar, r - 1, comm->comm, cs));       // my first rows
        LCASR_NCCL(g_nccl.Recv(R.ext(0), R.halo_bytes(), ncclChar, r - 1, comm->comm, cs));            // its last rows
      }
      if (r < P - 1) {
        LCASR_NCCL(g_nccl.Send(R.ext(p.n), R.halo_bytes(), ncclChar, r + 1, comm->comm, cs));          // my last rows
        LCASR_NCCL(g_nccl.Recv(R.ext(p.halo + p.n), R.halo_bytes(), ncclChar, r + 1, comm->comm, cs)); // its first rows
      }
      LCASR_NCCL(g_nccl.GroupEnd());
      cudaEvent_t ev_halo = comm->next_event();
      LCASR_CUDA(cudaEventRecord(ev_halo, cs));
      LCASR_CUDA(cudaStreamWaitEvent(st, ev_halo, 0));
    }
    LCASR_TRY(R.layer_c(l));
  }
  LCASR_TRY(R.post(return_logits));
  if (P > 1 && !return_logits) {  // everyone gets everyone's argmax ids (tiny): the greedy collapse needs the seams
    LCASR_NCCL(g_nccl.GroupStart());
    for (int j = 0; j < P; ++j) {
      if (j == r) continue;
      LCASR_NCCL(g_nccl.Send(R.am_full + p.s0, (size_t)p.n * 4, ncclChar, j, comm->comm, st));
      LCASR_NCCL(g_nccl.Recv(R.am_full + p.start[j], (size_t)p.cnt[j] * 4, ncclChar, j, comm->comm, st));
    }
    LCASR_NCCL(g_nccl.GroupEnd());
  }
  return 0;
}

// All P ranks of the same algorithm in ONE process on one GPU (device-to-device copies instead of NCCL transfers, one
// stream): the single-GPU parity test of the sequence-parallel path.  out_full [N, num_classes], argmax_full [N] or NULL.
// workspace: sum over ranks of lcasr_model_seqpar_workspace_bytes(m, world, rank, T) (each 256-byte aligned).
extern "C" int lcasr_model_forward_seqpar_emulated(lcasr_model* m, int world, const float* spec_full, int64_t T, float* out_full,
                                                   int32_t* argmax_full, int return_logits, void* workspace,
                                                   int64_t workspace_bytes, void* stream) {
  LCASR_CHECK_ARG(m && spec_full && out_full && workspace && world >= 1, "forward_seqpar_emulated: bad argument");
  LCASR_TRY(check_model_for_sp(m));
  std::vector<SpRank> R(world);
  size_t off = 0;
  for (int r = 0; r < world; ++r) {
    R[r].m = m;
    LCASR_TRY(make_sp_plan(m->cfg, world, r, T, &R[r].p));
    R[r].ws = (char*)workspace + off;
    off += al256(R[r].p.total);
    R[r].st = (cudaStream_t)stream;
    R[r].out = out_full + (size_t)R[r].p.s0 * m->cfg.num_classes;
    R[r].am_full = argmax_full;  // all ranks write their block into the same gathered buffer
  }
  LCASR_CHECK_ARG((size_t)workspace_bytes >= off, "forward_seqpar_emulated: workspace %lld < required %lld bytes",
                  (long long)workspace_bytes, (long long)off);
  LCASR_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "forward_seqpar_emulated: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const bool partial = R[0].partial_mode();
  for (int r = 0; r < world; ++r) LCASR_TRY(R[r].pre(spec_full, T));
  for (int l = 0; l < m->cfg.n_layers; ++l) {
    for (int r = 0; r < world; ++r) LCASR_TRY(R[r].layer_a(l));
    for (int r = 0; r < world; ++r)
      for (int j = 0; j < world; ++j) {
        if (j == r) continue;
        const SpPlan& p = R[r].p;
        LCASR_CUDA(cudaMemcpyAsync(R[r].K(p.start[j]), R[j].K(p.start[j]), R[r].block_bytes(j), cudaMemcpyDeviceToDevice, st));
        LCASR_CUDA(cudaMemcpyAsync(R[r].V(p.start[j]), R[j].V(p.start[j]), R[r].block_bytes(j), cudaMemcpyDeviceToDevice, st));
      }
    if (partial)
      for (int r = 0; r < world; ++r) {  // the same launches as the real driver's default exchange mode
        LCASR_TRY(R[r].attn_block(0, r, st));
        LCASR_TRY(R[r].attn_others(st, st));
      }
    for (int r = 0; r < world; ++r) LCASR_TRY(R[r].layer_b(l));
    for (int r = 0; r < world; ++r) {
      const SpPlan& p = R[r].p;
      if (r > 0)
        LCASR_CUDA(cudaMemcpyAsync(R[r].ext(0), R[r - 1].ext(R[r - 1].p.n), R[r].halo_bytes(), cudaMemcpyDeviceToDevice, st));
      if (r < world - 1)
        LCASR_CUDA(cudaMemcpyAsync(R[r].ext(p.halo + p.n), R[r + 1].ext(p.halo), R[r].halo_bytes(), cudaMemcpyDeviceToDevice, st));
    }
    for (int r = 0; r < world; ++r) LCASR_TRY(R[r].layer_c(l));
  }
  for (int r = 0; r < world; ++r) LCASR_TRY(R[r].post(return_logits));
  return 0;
}

extern "C" int64_t lcasr_model_seqpar_emulated_workspace_bytes(const lcasr_model* m, int world, int64_t T) {
  if (!m || world < 1) return -1;
  size_t off = 0;
  for (int r = 0; r < world; ++r) {
    SpPlan p;
    if (make_sp_plan(m->cfg, world, r, T, &p) != 0) return -1;
    off += al256(p.total);
  }
  return (int64_t)off;
}
