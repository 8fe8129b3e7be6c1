// C-ABI plumbing: error reporting, launch counter, GEMM / attention dispatch.
#include "common.cuh"
#include <cstring>

namespace lcasr {

std::atomic<int64_t> g_launch_count{0};

char* last_error_buf() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

int gemm_simt_launch(const void* A, const void* W, int ab_dtype, int64_t M, int N, int K, const float* bias, int act,
                     const float* resid, float alpha, void* out, int out_dtype, cudaStream_t st);
int gemm_tc_launch(const void* A, const void* W, int64_t M, int N, int K, const float* bias, int act, const float* resid,
                   float alpha, void* out, int out_dtype, cudaStream_t st, void* pre_out = nullptr);
int gemm_tc_launch_rope(const void* A, const void* W, int64_t M, int N, int K, const float* cos_t, const float* sin_t,
                        int64_t rope_n, int rope_cols, int dh, void* out, cudaStream_t st);
int gemm_tc_launch_glu(const void* A, const void* W, int64_t M, int N, int K, const float* bias, void* out, cudaStream_t st);
int attn_simt_launch(const void* q, const void* k, const void* v, int dtype, int B, int64_t N, int64_t Nk, const int32_t* kv_len,
                     int H, int Dh, int v_transposed, int64_t Npad, void* out, cudaStream_t st, int wl = -1, int wr = -1);
int attn_tc_launch(const void* q, const void* k, const void* v, int B, int64_t N, int64_t Nk, const int32_t* kv_len, int H,
                   int Dh, int v_transposed, int64_t Npad, void* out, float* lse, cudaStream_t st, int wl = -1, int wr = -1,
                   float* out32 = nullptr, int64_t ldq = 0, int64_t ldkv = 0);

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_abi_version(void) { return LCASR_ABI_VERSION; }
extern "C" const char* lcasr_last_error(void) { return last_error_buf(); }
extern "C" int64_t lcasr_launch_count(void) { return g_launch_count.load(); }
extern "C" void lcasr_reset_launch_count(void) { g_launch_count.store(0); }

extern "C" int64_t lcasr_out_length(int64_t T) {
  // calc_length (subsampling.py:557-567): floor((L + 2 - 3) / 2 + 1) three times == (L-1)/2 + 1 for L >= 1
  for (int i = 0; i < 3; ++i) T = T >= 1 ? (T - 1) / 2 + 1 : 0;
  return T;
}

extern "C" int lcasr_gemm(const void* A, const void* W, int ab_dtype, int64_t M, int N, int K, const float* bias,
                          int act, const float* resid, float alpha, void* out, int out_dtype, int impl, void* stream) {
  LCASR_CHECK_ARG(A && W && out, "gemm: NULL operand");
  LCASR_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  LCASR_CHECK_ARG(ab_dtype == LCASR_F32 || ab_dtype == LCASR_BF16, "gemm: bad operand dtype %d", ab_dtype);
  LCASR_CHECK_ARG(out_dtype == LCASR_F32 || out_dtype == LCASR_BF16, "gemm: bad output dtype %d", out_dtype);
  LCASR_CHECK_ARG(act >= LCASR_ACT_NONE && act <= LCASR_ACT_SILU, "gemm: bad activation %d", act);
  LCASR_CHECK_ARG(!resid || out_dtype == LCASR_F32, "gemm: a residual epilogue writes fp32");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == LCASR_GEMM_AUTO) impl = ab_dtype == LCASR_BF16 ? LCASR_GEMM_TCGEN05 : LCASR_GEMM_SIMT;
  if (impl == LCASR_GEMM_TCGEN05) {
    LCASR_CHECK_ARG(ab_dtype == LCASR_BF16, "gemm: the tcgen05 kernel takes bf16 operands");
    return gemm_tc_launch(A, W, M, N, K, bias, act, resid, alpha, out, out_dtype, st);
  }
  LCASR_CHECK_ARG(impl == LCASR_GEMM_SIMT, "gemm: bad impl %d", impl);
  return gemm_simt_launch(A, W, ab_dtype, M, N, K, bias, act, resid, alpha, out, out_dtype, st);
}

static int attention_dispatch(const void* q, const void* k, const void* v, int dtype, int B, int64_t N, int64_t Nk,
                              const int32_t* kv_len, int H, int Dh, int v_transposed, int64_t Npad, void* out, int impl,
                              void* stream, int wl = -1, int wr = -1) {
  LCASR_CHECK_ARG(q && k && v && out, "attention: NULL operand");
  LCASR_CHECK_ARG(B > 0 && N > 0 && Nk > 0 && H > 0 && Dh > 0, "attention: bad shape");
  LCASR_CHECK_ARG(dtype == LCASR_F32 || dtype == LCASR_BF16, "attention: bad dtype %d", dtype);
  LCASR_CHECK_ARG(!v_transposed || Npad >= Nk, "attention: Npad < N");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == LCASR_ATTN_AUTO) impl = dtype == LCASR_BF16 ? LCASR_ATTN_TCGEN05 : LCASR_ATTN_SIMT;
  if (impl == LCASR_ATTN_TCGEN05) {
    LCASR_CHECK_ARG(dtype == LCASR_BF16, "attention: the tcgen05 kernel takes bf16 operands");
    return attn_tc_launch(q, k, v, B, N, Nk, kv_len, H, Dh, v_transposed, Npad, out, nullptr, st, wl, wr);
  }
  LCASR_CHECK_ARG(impl == LCASR_ATTN_SIMT, "attention: bad impl %d", impl);
  return attn_simt_launch(q, k, v, dtype, B, N, Nk, kv_len, H, Dh, v_transposed, Npad, out, st, wl, wr);
}

extern "C" int lcasr_attention(const void* q, const void* k, const void* v, int dtype, int B, int64_t N, int H, int Dh,
                               int v_transposed, int64_t Npad, void* out, int impl, void* stream) {
  return attention_dispatch(q, k, v, dtype, B, N, N, nullptr, H, Dh, v_transposed, Npad, out, impl, stream);
}

extern "C" int lcasr_attention_cross(const void* q, const void* k, const void* v, int dtype, int B, int64_t Nq, int64_t Nk,
                                     int H, int Dh, void* out, int impl, void* stream) {
  return attention_dispatch(q, k, v, dtype, B, Nq, Nk, nullptr, H, Dh, 0, 0, out, impl, stream);
}

extern "C" int lcasr_attention_masked(const void* q, const void* k, const void* v, int dtype, int B, int64_t Nq, int64_t Nk,
                                      const int32_t* kv_len, int H, int Dh, void* out, int impl, void* stream) {
  return attention_dispatch(q, k, v, dtype, B, Nq, Nk, kv_len, H, Dh, 0, 0, out, impl, stream);
}

extern "C" int lcasr_gemm_rope(const void* A, const void* W, int64_t M, int N, int K, const float* cos_t, const float* sin_t,
                               int64_t rope_n, int rope_cols, int Dh, void* out, void* stream) {
  LCASR_CHECK_ARG(A && W && out, "gemm_rope: NULL operand");
  LCASR_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm_rope: bad shape");
  if (M == 0) return 0;
  return gemm_tc_launch_rope(A, W, M, N, K, cos_t, sin_t, rope_n, rope_cols, Dh, out, (cudaStream_t)stream);
}

extern "C" int lcasr_gemm_glu(const void* A, const void* W, int64_t M, int N, int K, const float* bias, void* out, void* stream) {
  LCASR_CHECK_ARG(A && W && out, "gemm_glu: NULL operand");
  LCASR_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm_glu: bad shape");
  if (M == 0) return 0;
  return gemm_tc_launch_glu(A, W, M, N, K, bias, out, (cudaStream_t)stream);
}

extern "C" int lcasr_attention_qkv(const void* qkv, int B, int64_t N, const int32_t* kv_len, int H, int Dh, void* out, void* stream) {
  LCASR_CHECK_ARG(qkv && out && B > 0 && N > 0 && H > 0 && Dh > 0, "attention_qkv: bad arguments");
  const int64_t d = (int64_t)H * Dh;
  const char* base = (const char*)qkv;
  return attn_tc_launch(base, base + d * 2, base + 2 * d * 2, B, N, N, kv_len, H, Dh, 0, 0, out, nullptr, (cudaStream_t)stream, -1, -1,
                        nullptr, 3 * d, 3 * d);
}

// training forward: bf16 tcgen05 attention that also returns the per-row log-sum-exp the backward needs
extern "C" int lcasr_attention_train(const void* q, const void* k, const void* v, int B, int64_t N, int H, int Dh, void* out,
                                     float* lse, void* stream) {
  LCASR_CHECK_ARG(q && k && v && out && lse, "attention_train: NULL operand");
  LCASR_CHECK_ARG(B > 0 && N > 0 && H > 0 && Dh > 0, "attention_train: bad shape");
  return attn_tc_launch(q, k, v, B, N, N, nullptr, H, Dh, 0, 0, out, lse, (cudaStream_t)stream);
}

// the same for a padded batch: keys at or beyond kv_len[b] are masked (rows of padded queries: see lcasr_attention_masked)
extern "C" int lcasr_attention_train_masked(const void* q, const void* k, const void* v, int B, int64_t N, const int32_t* kv_len,
                                            int H, int Dh, void* out, float* lse, void* stream) {
  LCASR_CHECK_ARG(q && k && v && out && lse, "attention_train_masked: NULL operand");
  LCASR_CHECK_ARG(B > 0 && N > 0 && H > 0 && Dh > 0, "attention_train_masked: bad shape");
  return attn_tc_launch(q, k, v, B, N, N, kv_len, H, Dh, 0, 0, out, lse, (cudaStream_t)stream);
}

// training forward: out = act(pre), pre = A.W^T + bias, both stored as bf16 (the backward needs the pre-activation)
extern "C" int lcasr_gemm_act_pre(const void* A, const void* W, int64_t M, int N, int K, const float* bias, int act, void* out,
                                  void* pre_out, void* stream) {
  LCASR_CHECK_ARG(A && W && out && pre_out, "gemm_act_pre: NULL operand");
  LCASR_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_act_pre: bad shape");
  LCASR_CHECK_ARG(act == LCASR_ACT_GELU_TANH || act == LCASR_ACT_SILU, "gemm_act_pre: bad activation %d", act);
  return gemm_tc_launch(A, W, M, N, K, bias, act, nullptr, 0.f, out, LCASR_BF16, (cudaStream_t)stream, pre_out);
}

// local (windowed) self-attention: query i attends to keys [i - win_left, i + win_right] (-1 = unlimited on that side)
extern "C" int lcasr_attention_window(const void* q, const void* k, const void* v, int dtype, int B, int64_t N, const int32_t* kv_len,
                                      int H, int Dh, int win_left, int win_right, void* out, int impl, void* stream) {
  LCASR_CHECK_ARG(win_left >= -1 && win_right >= -1, "attention_window: bad window (%d, %d)", win_left, win_right);
  return attention_dispatch(q, k, v, dtype, B, N, N, kv_len, H, Dh, 0, 0, out, impl, stream, win_left, win_right);
}
