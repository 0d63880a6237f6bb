// kr_resample_core.cuh — body of the speed-perturbation resampler (third item): the reference's
// torchaudio.functional.resample(audio, orig_freq, new_freq) call in Dataset.__getitem__ (data/dataset.py:674-684) with
// torchaudio's defaults (windowed-sinc interpolation, Hann window, lowpass_filter_width 6, rolloff 0.99).
//
// torchaudio materialises a [new, 1, 2 * width + orig] filter bank after reducing the two rates by their gcd — for a
// speed factor such as 0.93 that is 10253 x 11039 taps (450 MB) of which about 14 per output sample are non-zero — and
// runs a strided conv1d.  Here every output sample evaluates its own ~14 taps on the fly (float64 like torchaudio's
// kernel construction, rounded to float32 before the multiply-add, including its float32 phase-offset quirk, so the taps are the same numbers) and reads ~14 input
// samples: HBM-bound, 4 B in + 4 B out per sample.
// DUAL-COMPILED like kr_features_core.cuh (g++ -DKR_HOST_EMU -> tests/emu/resample_emu.cpp).
#pragma once

#ifdef KR_HOST_EMU
#include <math.h>
#define KRR_DEV static inline
#else
#define KRR_DEV __device__ __forceinline__
#endif

namespace krr {

struct Plan {               // torchaudio/functional/functional.py: _get_sinc_resample_kernel
  int orig, neu;            // rates divided by their gcd
  int width;                // ceil(lowpass_filter_width * orig / base_freq)
  int lpw;                  // lowpass_filter_width
  double base_freq;         // min(orig, neu) * rolloff
};

// tap k (0 <= k < 2 * width + orig) of output phase p, as the float32 value torchaudio's filter bank holds
KRR_DEV float tap(const Plan& pl, int p, int k) {
  const double kPi = 3.14159265358979323846;
  const double idx = (double)(k - pl.width) / (double)pl.orig;
  // torchaudio builds the phase offsets as `arange(0, -new, -1) / new` on an int64 tensor, i.e. in FLOAT32, and only then
  // adds the float64 tap positions: the rounding of p / new is multiplied by base_freq (~1e4 for awkward ratios) and moves
  // the taps by up to 1e-4 — reproduced, since the filter bank is what defines the reference output
  double t = ((double)((float)(-p) / (float)pl.neu) + idx) * pl.base_freq;
  t = t < -(double)pl.lpw ? -(double)pl.lpw : (t > (double)pl.lpw ? (double)pl.lpw : t);
  const double c = cos(t * kPi / (double)pl.lpw / 2.0);
  const double window = c * c;
  t *= kPi;
  const double scale = pl.base_freq / (double)pl.orig;
  const double s = t == 0.0 ? 1.0 : sin(t) / t;
  return (float)(s * (window * scale));
}

// Output sample j of an utterance of n input samples (zero outside [0, n)): frame i = j / neu of the strided
// convolution, phase p = j % neu; only the taps inside the window's support are visited.
KRR_DEV float resample_sample(const Plan& pl, const float* x, long long n, long long j) {
  const long long i = j / pl.neu;
  const int p = (int)(j - i * pl.neu);
  // |(-p / neu + (k - width) / orig) * base_freq| < lpw   <=>   k in (centre - half, centre + half)
  const double centre = (double)pl.width + (double)pl.orig * (double)p / (double)pl.neu;
  const double half = (double)pl.orig * (double)pl.lpw / pl.base_freq;
  int k_lo = (int)floor(centre - half) - 1, k_hi = (int)ceil(centre + half) + 1;   // +-1: the float32 phase rounding
  if (k_lo < 0) k_lo = 0;
  if (k_hi > 2 * pl.width + pl.orig - 1) k_hi = 2 * pl.width + pl.orig - 1;
  float acc = 0.f;
  for (int k = k_lo; k <= k_hi; ++k) {
    const long long m = i * pl.orig + k - pl.width;        // index into the un-padded input
    if (m < 0 || m >= n) continue;
    acc += x[m] * tap(pl, p, k);
  }
  return acc;
}

// ceil(neu * n / orig), torchaudio's target_length
KRR_DEV long long out_length(const Plan& pl, long long n) { return (n * pl.neu + pl.orig - 1) / pl.orig; }

}  // namespace krr
