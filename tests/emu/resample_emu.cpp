// TEST INFRASTRUCTURE — host emulation of kr_resample (kokoro_ruslan_b200/csrc/kr_resample.cu): the kernel's own body
// (kr_resample_core.cuh) compiled with -DKR_HOST_EMU.  Used by tests/test_resample_emu_cpu.py only.
#define KR_HOST_EMU 1
#include "kr_resample_core.cuh"

static long long gcd_ll(long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; }

extern "C" long long emu_resample_length(long long n, int orig_freq, int new_freq) {
  const long long g = gcd_ll(orig_freq, new_freq);
  return (n * (new_freq / g) + (orig_freq / g) - 1) / (orig_freq / g);
}

extern "C" int emu_resample(const float* x, const long long* lengths, float* y, int B, long long n_max, long long m_max,
                            int orig_freq, int new_freq, int lowpass_filter_width, float rolloff) {
  const long long g = gcd_ll(orig_freq, new_freq);
  krr::Plan pl;
  pl.orig = (int)(orig_freq / g);
  pl.neu = (int)(new_freq / g);
  pl.lpw = lowpass_filter_width;
  pl.base_freq = (double)(pl.orig < pl.neu ? pl.orig : pl.neu) * (double)rolloff;
  pl.width = (int)ceil((double)lowpass_filter_width * (double)pl.orig / pl.base_freq);
  for (int b = 0; b < B; ++b) {
    const long long n = lengths ? lengths[b] : n_max;
    const long long m = krr::out_length(pl, n);
    for (long long j = 0; j < m_max; ++j)
      y[(long long)b * m_max + j] = j < m ? krr::resample_sample(pl, x + (long long)b * n_max, n, j) : 0.f;
  }
  return 0;
}
