// kr_resample.cu — windowed-sinc resampler for the speed perturbation of the feature pipeline (SURVEY.md §8(f) N1;
// reference data/dataset.py:674-684 -> torchaudio.functional.resample).  Body and design notes: kr_resample_core.cuh.
#include "kr_common.cuh"
#include "kr_resample_core.cuh"

namespace {
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, const long long* __restrict__ lengths, float* __restrict__ y,
                long long n_max, long long m_max, krr::Plan pl) {
  kr::pdl_entry();
  const int b = blockIdx.y;
  const long long n = lengths != nullptr ? lengths[b] : n_max;
  const long long m = krr::out_length(pl, n);
  const float* xb = x + (long long)b * n_max;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m_max; j += (long long)gridDim.x * blockDim.x)
    y[(long long)b * m_max + j] = j < m ? krr::resample_sample(pl, xb, n, j) : 0.f;
}

long long gcd_ll(long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; }
}  // namespace

// y[B, m_max] = resample(x[B, n_max]) row by row; rows are zero beyond ceil(new_freq * len[b] / orig_freq).
// m_max must be >= that length for n_max (kr_resample_length).
extern "C" long long kr_resample_length(long long n, int orig_freq, int new_freq) {
  if (orig_freq <= 0 || new_freq <= 0 || n < 0) return -1;
  const long long g = gcd_ll(orig_freq, new_freq);
  return (n * (new_freq / g) + (orig_freq / g) - 1) / (orig_freq / g);
}

extern "C" int kr_resample(const float* x, const long long* lengths, float* y, int B, long long n_max, long long m_max,
                           int orig_freq, int new_freq, int lowpass_filter_width, float rolloff, void* stream) {
  if (B <= 0 || m_max <= 0) return KR_OK;
  if (orig_freq <= 0 || new_freq <= 0 || lowpass_filter_width <= 0 || !(rolloff > 0.f)) {
    kr_set_error("kr_resample: rates, filter width and rolloff must be positive");
    return KR_ERR_ARG;
  }
  const long long g = gcd_ll(orig_freq, new_freq);
  krr::Plan pl;
  pl.orig = (int)(orig_freq / g);
  pl.neu = (int)(new_freq / g);
  pl.lpw = lowpass_filter_width;
  pl.base_freq = (double)(pl.orig < pl.neu ? pl.orig : pl.neu) * (double)rolloff;
  pl.width = (int)ceil((double)lowpass_filter_width * (double)pl.orig / pl.base_freq);
  if (m_max < kr_resample_length(n_max, orig_freq, new_freq)) { kr_set_error("kr_resample: output row too short"); return KR_ERR_ARG; }
  long long blocks = (m_max + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  kr::launch(resample_kernel, dim3((unsigned)blocks, B), 256, 0, (cudaStream_t)stream, x, lengths, y, n_max, m_max, pl);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
