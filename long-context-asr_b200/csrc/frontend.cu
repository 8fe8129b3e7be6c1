// Front-end of the path: the 80-bin mel power spectrogram the encoder consumes —
// lcasr/utils/audio_tools.py:44-57 `to_spectogram` = torchaudio.transforms.MelSpectrogram(sample_rate 16 kHz, n_fft 512,
// win_length 400 (periodic Hann, zero-padded to 512 centred), hop 160, center=True / reflect padding, power 2, HTK mel
// scale, no filterbank normalisation, NO log) followed by per-bin standardisation over time ((x - mean) / std, unbiased).
// One hour of audio is 360 000 frames x 512-point DFTs = 0.19 TFLOP in fp32: the direct DFT against a window-premultiplied
// twiddle table (1 MB, L2-resident) is cheaper than staging an FFT, and its summation is as accurate as the fp32 FFT the
// reference runs.  Kernel 1: frames -> |DFT|^2 -> mel filterbank -> out[mel][frame] + per-mel sum / sum of squares
// (fp64 atomics).  Kernel 2: standardise in place.
#include "common.cuh"

namespace lcasr {

constexpr int kFeFT = 16;          // frames per CTA
constexpr int kFeNfft = 512, kFeBins = 257, kFeHop = 160, kFeThreads = 288;

__global__ void __launch_bounds__(kFeThreads) melspec_kernel(const float* __restrict__ wave, int64_t L,
                                                             const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                                             const float* __restrict__ fb, int n_mels, int64_t n_frames,
                                                             float* __restrict__ out, double* __restrict__ sums) {
  __shared__ float s_x[(kFeFT - 1) * kFeHop + kFeNfft];
  __shared__ float s_pow[kFeFT][kFeBins + 3];
  const int64_t f0 = (int64_t)blockIdx.x * kFeFT;
  const int64_t b = blockIdx.y;
  const float* w = wave + b * L;
  // frame f covers padded samples [f*hop, f*hop + n_fft) of the reflect-padded signal (pad n_fft/2 on both sides)
  const int span = (kFeFT - 1) * kFeHop + kFeNfft;
  for (int i = threadIdx.x; i < span; i += kFeThreads) {
    int64_t t = f0 * kFeHop + i - kFeNfft / 2;  // index into the unpadded waveform
    if (t < 0) t = -t;                           // reflect (no edge repeat), torch.stft pad_mode='reflect'
    if (t >= L) t = 2 * (L - 1) - t;
    s_x[i] = (t >= 0 && t < L) ? w[t] : 0.f;
  }
  __syncthreads();
  const int k = threadIdx.x;
  if (k < kFeBins) {
    float re[kFeFT], im[kFeFT];
#pragma unroll
    for (int f = 0; f < kFeFT; ++f) re[f] = im[f] = 0.f;
    for (int n = 0; n < kFeNfft; ++n) {
      const float c = cos_tab[n * kFeBins + k], s = sin_tab[n * kFeBins + k];
#pragma unroll
      for (int f = 0; f < kFeFT; ++f) {
        const float x = s_x[f * kFeHop + n];
        re[f] = fmaf(x, c, re[f]);
        im[f] = fmaf(x, s, im[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < kFeFT; ++f) s_pow[f][k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kFeFT * n_mels; idx += kFeThreads) {
    const int m = idx / kFeFT, f = idx % kFeFT;  // consecutive threads -> consecutive frames of one mel bin (coalesced store)
    const int64_t frame = f0 + f;
    if (frame >= n_frames) continue;
    float acc = 0.f;
    for (int kk = 0; kk < kFeBins; ++kk) acc = fmaf(fb[kk * n_mels + m], s_pow[f][kk], acc);
    out[(b * n_mels + m) * n_frames + frame] = acc;
    if (sums) {
      atomicAdd(&sums[(b * n_mels + m) * 2], (double)acc);
      atomicAdd(&sums[(b * n_mels + m) * 2 + 1], (double)acc * (double)acc);
    }
  }
}

__global__ void __launch_bounds__(256) melspec_normalize_kernel(float* __restrict__ x, int64_t n_frames,
                                                                const double* __restrict__ sums) {
  const int64_t row = blockIdx.y;  // (batch, mel)
  const double n = (double)n_frames;
  const double mean = sums[row * 2] / n;
  const double var = (sums[row * 2 + 1] - n * mean * mean) / (n - 1.0);  // torch.std: unbiased
  const float mu = (float)mean, inv = (float)(1.0 / sqrt(var));
  float* xr = x + row * n_frames;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_frames; i += (int64_t)gridDim.x * blockDim.x)
    xr[i] = (xr[i] - mu) * inv;
}

}  // namespace lcasr

using namespace lcasr;

extern "C" int64_t lcasr_melspec_frames(int64_t n_samples) { return 1 + n_samples / kFeHop; }

extern "C" int lcasr_melspec(const float* wave, int B, int64_t n_samples, const float* cos_tab, const float* sin_tab,
                             const float* fb, int n_mels, float* out, double* sums, int normalise, void* stream) {
  LCASR_CHECK_ARG(wave && cos_tab && sin_tab && fb && out, "melspec: NULL argument");
  LCASR_CHECK_ARG(B > 0 && B <= 65535 && n_samples > kFeNfft / 2 && n_mels > 0, "melspec: bad shape (needs more than %d samples)", kFeNfft / 2);
  LCASR_CHECK_ARG(!normalise || sums, "melspec: normalisation needs the sums scratch [B, n_mels, 2] (fp64)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_frames = lcasr_melspec_frames(n_samples);
  if (sums) LCASR_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * n_mels * 2 * sizeof(double), st));
  dim3 grid((unsigned)ceil_div(n_frames, kFeFT), (unsigned)B);
  melspec_kernel<<<grid, kFeThreads, 0, st>>>(wave, n_samples, cos_tab, sin_tab, fb, n_mels, n_frames, out, sums);
  LCASR_LAUNCH_CHECK();
  if (normalise) {
    LCASR_CHECK_ARG(n_frames > 1, "melspec: standardisation needs at least 2 frames");
    dim3 g2((unsigned)min((int64_t)64, ceil_div(n_frames, 256)), (unsigned)(B * n_mels));
    melspec_normalize_kernel<<<g2, 256, 0, st>>>(out, n_frames, sums);
    LCASR_LAUNCH_CHECK();
  }
  return 0;
}
