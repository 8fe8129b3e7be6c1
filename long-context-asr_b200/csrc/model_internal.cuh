// Shared between model.cu (single-GPU forward) and seqpar.cu (sequence-parallel forward): the model handle.
#pragma once
#include "common.cuh"
#include <vector>

// kernel categories for the in-step CUDA-event timing (bench.py's roofline numbers)
enum { CAT_SUBSAMPLE = 0, CAT_NORM, CAT_GEMM, CAT_ATTN, CAT_ROPE, CAT_CONVMOD, CAT_SOFTMAX, CAT_COUNT };

struct lcasr_model {
  lcasr_config cfg;
  lcasr_weights w;
  std::vector<lcasr_layer_weights> layers;
  int attn_impl = LCASR_ATTN_AUTO;
  int gemm_impl = LCASR_GEMM_AUTO;
  // optional per-op timing: (start,end) event pairs recorded on the launch stream
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;
  struct Span { int cat; size_t e0, e1; };
  std::vector<Span> spans;
  size_t ev_used = 0;
  cudaEvent_t next_event() {
    if (ev_used == ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev_pool.push_back(e);
    }
    return ev_pool[ev_used++];
  }
  // wave-quantisation fix of the attention launch (model.cu: attention_tail_split): side streams + fork / join events
  static constexpr int kTailStreams = 5;
  cudaStream_t tail_streams[kTailStreams] = {};
  cudaEvent_t tail_fork = nullptr, tail_done[kTailStreams] = {};
  struct TailPlan { int B = -1; int64_t N = -1; int t = 0, P = 0; };  // last t query-tile pairs of the last recording in P key pieces
  TailPlan tail_plan;
  int tail_force_t = -1, tail_force_p = 0;  // lcasr_model_set_attention_tail
  ~lcasr_model() {
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
    for (cudaStream_t s : tail_streams) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : tail_done) if (e) cudaEventDestroy(e);
    if (tail_fork) cudaEventDestroy(tail_fork);
  }
};

