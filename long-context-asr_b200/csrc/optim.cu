// Optimizer step of the training loop: global-norm gradient clipping (exp/train.py:54-55,
// torch.nn.utils.clip_grad_norm_) and MADGRAD (lcasr/optim/madgrad.py:81-212, dense branch) as multi-tensor kernels over
// a device table of (parameter, gradient, state) pointers.  The reference issues ~12 elementwise torch kernels per
// parameter tensor and step (pow, addcmul, addcdiv, ... x 330 tensors); here one reduction launch produces sum(g^2) and one
// launch updates every parameter, with the clip coefficient read from device memory (no host synchronisation).
// HBM-bound: per element reads p, g, grad_sum_sq, s, x0 (20 B) and writes p, grad_sum_sq, s (12 B).
#include "common.cuh"

namespace lcasr {

constexpr int kOptChunk = 256 * 16;  // elements per CTA pass

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const lcasr_opt_tensor* __restrict__ tensors,
                                                         const int32_t* __restrict__ chunk_tensor,
                                                         const int32_t* __restrict__ chunk_index, float* __restrict__ sumsq) {
  __shared__ float red[8];
  const lcasr_opt_tensor t = tensors[chunk_tensor[blockIdx.x]];
  const int64_t i0 = (int64_t)chunk_index[blockIdx.x] * kOptChunk;
  float acc = 0.f;
  if (t.g) {
    for (int64_t i = i0 + threadIdx.x; i < i0 + kOptChunk && i < t.n; i += 256) {
      const float g = t.g[i];
      acc = fmaf(g, g, acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    if (s != 0.f) atomicAdd(sumsq, s);
  }
}

// g *= min(1, max_norm / (sqrt(sumsq) + 1e-6))   (clip_grad_norm_, in place)
__global__ void __launch_bounds__(256) grad_scale_kernel(const lcasr_opt_tensor* __restrict__ tensors,
                                                         const int32_t* __restrict__ chunk_tensor,
                                                         const int32_t* __restrict__ chunk_index, const float* __restrict__ sumsq,
                                                         float max_norm) {
  const lcasr_opt_tensor t = tensors[chunk_tensor[blockIdx.x]];
  if (!t.g) return;
  const float coef = fminf(max_norm / (sqrtf(*sumsq) + 1e-6f), 1.0f);
  const int64_t i0 = (int64_t)chunk_index[blockIdx.x] * kOptChunk;
  float* g = const_cast<float*>(t.g);
  for (int64_t i = i0 + threadIdx.x; i < i0 + kOptChunk && i < t.n; i += 256) g[i] *= coef;
}

struct MadgradHyper {
  float lr, lamb, eps, decay, momentum, max_norm;
  int decouple_decay;
};

__global__ void __launch_bounds__(256) madgrad_step_kernel(const lcasr_opt_tensor* __restrict__ tensors,
                                                           const int32_t* __restrict__ chunk_tensor,
                                                           const int32_t* __restrict__ chunk_index, const float* __restrict__ sumsq,
                                                           MadgradHyper h) {
  const lcasr_opt_tensor t = tensors[chunk_tensor[blockIdx.x]];
  if (!t.g) return;  // parameter without gradient: skipped like `if p.grad is None: continue`
  const float coef = (sumsq && h.max_norm > 0.f) ? fminf(h.max_norm / (sqrtf(*sumsq) + 1e-6f), 1.0f) : 1.0f;
  const float ck = 1.0f - h.momentum;
  const int64_t i0 = (int64_t)chunk_index[blockIdx.x] * kOptChunk;
  for (int64_t i = i0 + threadIdx.x; i < i0 + kOptChunk && i < t.n; i += 256) {
    float p = t.p[i];
    float g = t.g[i] * coef;
    float gss = t.gss[i], s = t.s[i];
    if (h.decay != 0.f && !h.decouple_decay) g = fmaf(h.decay, p, g);
    float x0;
    if (h.momentum == 0.f) {  // x0 from the other known quantities (madgrad.py:168-171)
      const float rms0 = cbrtf(gss) + h.eps;
      x0 = p + s / rms0;
    } else {
      x0 = t.x0[i];
    }
    gss = fmaf(h.lamb * g, g, gss);
    float rms = cbrtf(gss) + h.eps;
    if (h.eps == 0.f && rms == 0.f) rms = INFINITY;
    s = fmaf(h.lamb, g, s);
    const float z = x0 - s / rms;
    const float p_old = p;
    p = h.momentum == 0.f ? z : fmaf(ck, z, (1.0f - ck) * p);
    if (h.decay != 0.f && h.decouple_decay) p -= h.lr * h.decay * p_old;
    t.p[i] = p; t.gss[i] = gss; t.s[i] = s;
  }
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int lcasr_grad_sumsq(const lcasr_opt_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                                int n_chunks, float* sumsq, void* stream) {
  LCASR_CHECK_ARG(tensors && chunk_tensor && chunk_index && sumsq && n_chunks > 0, "grad_sumsq: bad arguments");
  LCASR_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float), (cudaStream_t)stream));
  grad_sumsq_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tensors, chunk_tensor, chunk_index, sumsq);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_grad_scale(const lcasr_opt_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                                int n_chunks, const float* sumsq, float max_norm, void* stream) {
  LCASR_CHECK_ARG(tensors && chunk_tensor && chunk_index && sumsq && n_chunks > 0 && max_norm > 0.f, "grad_scale: bad arguments");
  grad_scale_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tensors, chunk_tensor, chunk_index, sumsq, max_norm);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_madgrad_step(const lcasr_opt_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                                  int n_chunks, const float* sumsq, float max_norm, float lr, float lamb, float eps,
                                  float weight_decay, float momentum, int decouple_decay, void* stream) {
  LCASR_CHECK_ARG(tensors && chunk_tensor && chunk_index && n_chunks > 0, "madgrad_step: bad arguments");
  LCASR_CHECK_ARG(momentum >= 0.f && momentum < 1.f && eps >= 0.f && weight_decay >= 0.f, "madgrad_step: bad hyper-parameters");
  MadgradHyper h{lr, lamb, eps, weight_decay, momentum, max_norm, decouple_decay};
  madgrad_step_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tensors, chunk_tensor, chunk_index, sumsq, h);
  LCASR_LAUNCH_CHECK();
  return 0;
}

extern "C" int lcasr_opt_chunk_elems(void) { return kOptChunk; }
