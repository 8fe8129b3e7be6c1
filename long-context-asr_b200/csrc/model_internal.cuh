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
  ~lcasr_model() {
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
  }
};

