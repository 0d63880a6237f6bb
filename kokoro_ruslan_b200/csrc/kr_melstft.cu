// kr_melstft.cu — log-mel feature extraction (reference data/dataset.py:162-178, 672, 694-697:
// torchaudio MelSpectrogram(sr 22050, n_fft 1024, win 1024, hop 256, f 0-8000, 80 mels, power 2,
// periodic Hann, center=True / reflect pad, HTK scale, norm=None) followed by log(x + 1e-9);
// SURVEY.md §9 S6).  HBM-bound: 1 KB of unique waveform in, 320 B out per frame.
//
// One CTA per frame: windowed, reflect-padded samples -> 1024-point in-place radix-4 FFT in shared
// memory (fp32, quarter-wave twiddle table) -> |X|^2 for the 513 one-sided bins -> HTK triangular
// filterbank (dense [n_mels, 513] weights, each warp reduces its filters with shuffles) -> log.
#include "kr_common.cuh"
#include "kr_features_core.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace {
using namespace kr;

constexpr int NFFT = 1024, HOP = 256, NBINS = NFFT / 2 + 1, THREADS = 256;

__global__ void wave_peak_kernel(const float* __restrict__ wav, const long long* __restrict__ lengths, long long n_max,
                                 unsigned int* __restrict__ peak_bits) {
  kr::pdl_entry();
  __shared__ float red[32];
  const int b = blockIdx.y;
  const long long n = lengths != nullptr ? lengths[b] : n_max;
  const float* x = wav + (long long)b * n_max;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = m;
  __syncthreads();
  if (w == 0) {
    m = l < (blockDim.x >> 5) ? red[l] : 0.f;
    m = warp_max(m);
    if (l == 0) atomicMax(peak_bits + b, __float_as_uint(m));   // non-negative floats order like uints
  }
}

// One CTA per frame; body in kr_features_core.cuh (mel_frame_body: in-place radix-4 transform shared with the pitch
// kernel, 5 shared-memory passes, 12 KB of shared memory), also compiled as a host emulation by the CPU tests.
// Measured on B200 (round 2, 8 x 800 frames): 174.7 us against 186.8 us for the radix-2 Stockham kernel it replaced.
__global__ void __launch_bounds__(THREADS)
mel_stft_kernel(const float* __restrict__ wav, const long long* __restrict__ lengths, const float* __restrict__ peak,
                const float* __restrict__ fb_t, const int* __restrict__ fb_ranges, float* __restrict__ out, long long n_max,
                int frames_max, int n_mels, float log_eps) {
  kr::pdl_entry();
  __shared__ float2 z[krf::MEL_NFFT];
  __shared__ float2 qw[krf::MEL_NFFT / 4 + 1];
  __shared__ float pw[krf::MEL_BINS + 3];
  const int f = blockIdx.x, b = blockIdx.y;
  const long long n = lengths != nullptr ? lengths[b] : n_max;
  float* orow = out + ((long long)b * n_mels) * frames_max + f;
  if (f >= 1 + (int)(n / HOP)) {
    for (int m = threadIdx.x; m < n_mels; m += THREADS) orow[(long long)m * frames_max] = 0.f;
    return;
  }
  const float gain = peak != nullptr ? 1.f / (peak[b] + 1e-9f) : 1.f;
  krf::mel_frame_body(wav + (long long)b * n_max, n, f, gain, fb_t, fb_ranges, n_mels, frames_max, log_eps, z, qw, pw, orow);
}

}  // namespace

extern "C" int kr_wave_peak(const float* wav, const long long* lengths, float* peak, int B, long long n_max, void* stream) {
  if (B <= 0 || n_max <= 0) return KR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(peak, 0, sizeof(float) * B, st);
  int bx = (int)((n_max + 256 * 8 - 1) / (256 * 8));
  if (bx > 64) bx = 64;
  kr::launch(wave_peak_kernel, dim3(bx, B), 256, 0, st, wav, lengths, n_max, reinterpret_cast<unsigned int*>(peak));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_mel_stft(const float* wav, const long long* lengths, const float* peak, const float* fb_t,
                           const int* fb_ranges, float* out, int B, long long n_max, int frames_max, int n_mels, int n_fft,
                           int hop, float log_eps, void* stream) {
  if (n_fft != NFFT || hop != HOP) { kr_set_error("kr_mel_stft: built for n_fft 1024 / hop 256"); return KR_ERR_UNSUPPORTED; }
  if (B <= 0 || frames_max <= 0) return KR_OK;
  if (n_max < NFFT / 2 + 1) { kr_set_error("kr_mel_stft: waveform shorter than the reflect padding"); return KR_ERR_ARG; }
  kr::launch(mel_stft_kernel, dim3(frames_max, B), THREADS, 0, (cudaStream_t)stream, wav, lengths, peak, fb_t, fb_ranges,
             out, n_max, frames_max, n_mels, log_eps);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
