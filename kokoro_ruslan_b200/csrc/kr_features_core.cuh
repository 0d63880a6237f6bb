// kr_features_core.cuh — bodies of the variance FEATURE extraction kernels: the reference's
// PitchExtractor.extract_pitch (model/variance_predictor.py:448-625, YIN-style CMND on 2048-sample Hann frames through a
// 4096-point FFT autocorrelation) and EnergyExtractor.extract_energy_from_mel (:633-688), which the reference runs on the
// CPU per utterance inside Dataset.__getitem__ (data/dataset.py:793-815).
//
// DUAL-COMPILED.  nvcc builds these bodies into the kernels of kr_features.cu.  The same source, compiled by g++ with
// -DKR_HOST_EMU, becomes a single-"thread" emulation (tests/emu/features_emu.cpp): thread id 0, block size 1, lane
// count 1, barriers and shuffles as identities — every `for (i = tid; i < n; i += nthreads)` loop then simply covers
// its whole range in order.  That lets the CPU test-suite check the index arithmetic, the in-place FFT and every
// thresholded decision of the kernels against the live-reference fixtures without a GPU.  What the emulation cannot
// see are missing barriers; hence the rule followed below: a phase only READS what an earlier phase wrote, and every
// phase that writes shared memory ends in KRF_SYNC().
#pragma once

#ifdef KR_HOST_EMU
#include <math.h>
#include <stdint.h>
#define KRF_DEV static inline
#define KRF_HD static inline
#ifdef KR_HOST_EMU_SIMT            // one host thread per CUDA thread, real block / warp geometry (tests/emu/emu_simt.h)
#include "emu_simt.h"
#define KRF_TID (emu::tid())
#define KRF_NT (emu::nthreads())
#define KRF_LANE (emu::tid() & 31)
#define KRF_NLANES 32
#define KRF_WARP (emu::tid() >> 5)
#define KRF_NWARPS (emu::nthreads() >> 5)
#define KRF_SYNC() emu::syncthreads()
#define KRF_WARP_SYNC() do { } while (0)
KRF_DEV float krf_warp_sum(float v) { return emu::warp_sum(v); }
KRF_DEV float krf_warp_min(float v) { return emu::warp_min(v); }
KRF_DEV float krf_warp_max(float v) { return emu::warp_max(v); }
KRF_DEV int krf_warp_min_int(int v) { return emu::warp_min_int(v); }
KRF_DEV float krf_block_sum(float v, float* red) { return emu::block_sum(v, red); }
KRF_DEV float krf_block_max(float v, float* red) { return emu::block_max(v, red); }
KRF_DEV int krf_block_sum_int(int v, int* red) { return emu::block_sum_int(v, red); }
#else                              // one sequential "thread" per block
#define KRF_TID 0
#define KRF_NT 1
#define KRF_LANE 0
#define KRF_NLANES 1
#define KRF_WARP 0
#define KRF_NWARPS 1
#define KRF_SYNC() do { } while (0)
#define KRF_WARP_SYNC() do { } while (0)
KRF_DEV float krf_warp_sum(float v) { return v; }
KRF_DEV float krf_warp_min(float v) { return v; }
KRF_DEV float krf_warp_max(float v) { return v; }
KRF_DEV int krf_warp_min_int(int v) { return v; }
KRF_DEV float krf_block_sum(float v, float*) { return v; }
KRF_DEV float krf_block_max(float v, float*) { return v; }
KRF_DEV int krf_block_sum_int(int v, int*) { return v; }
#endif
struct krf_float2 { float x, y; };
KRF_DEV krf_float2 krf_make2(float x, float y) { krf_float2 r; r.x = x; r.y = y; return r; }
KRF_DEV void krf_sincospi(float x, float* s, float* c) { const double a = M_PI * (double)x; *s = (float)sin(a); *c = (float)cos(a); }
KRF_DEV float krf_cospi(float x) { return (float)cos(M_PI * (double)x); }
KRF_DEV float krf_mul(float a, float b) { volatile float r = a * b; return r; }      // no FMA contraction
KRF_DEV float krf_ldg(const float* p) { return *p; }
KRF_DEV float krf_exp(float x) { return expf(x); }
KRF_DEV float krf_log1p(float x) { return log1pf(x); }
#else
#define KRF_DEV __device__ __forceinline__
#define KRF_HD __host__ __device__ inline
#define KRF_TID ((int)threadIdx.x)
#define KRF_NT ((int)blockDim.x)
#define KRF_LANE ((int)(threadIdx.x & 31))
#define KRF_NLANES 32
#define KRF_WARP ((int)(threadIdx.x >> 5))
#define KRF_NWARPS ((int)(blockDim.x >> 5))
#define KRF_SYNC() __syncthreads()
#define KRF_WARP_SYNC() __syncwarp()
typedef float2 krf_float2;
KRF_DEV krf_float2 krf_make2(float x, float y) { return make_float2(x, y); }
KRF_DEV float krf_warp_sum(float v) { return kr::warp_sum(v); }
KRF_DEV float krf_warp_max(float v) { return kr::warp_max(v); }
KRF_DEV float krf_warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
KRF_DEV int krf_warp_min_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
KRF_DEV float krf_block_sum(float v, float* red) { return kr::block_sum(v, red); }
// block-wide max / integer sum; `red` holds >= 32 words of shared memory; all threads get the result
KRF_DEV float krf_block_max(float v, float* red) {
  v = kr::warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : -3.402823466e38f;
  return kr::warp_max(r);
}
KRF_DEV int krf_block_sum_int(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  int r = (l < nw) ? red[l] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}
KRF_DEV void krf_sincospi(float x, float* s, float* c) { sincospif(x, s, c); }
KRF_DEV float krf_cospi(float x) { return cospif(x); }
KRF_DEV float krf_mul(float a, float b) { return __fmul_rn(a, b); }
KRF_DEV float krf_ldg(const float* p) { return __ldg(p); }
KRF_DEV float krf_exp(float x) { return expf(x); }
KRF_DEV float krf_log1p(float x) { return log1pf(x); }
#endif

namespace krf {

constexpr int WIN = 2048;          // max(2048, hop * 8) with hop 256, variance_predictor.py:492
constexpr int HOP = 256;
constexpr int NFFT = 2 * WIN;      // zero-padded autocorrelation length, :517
constexpr int LOG4_NFFT = 6;
constexpr int TW = NFFT / 4 + 1;   // quarter-wave twiddle table: exp(-2 pi i j / 4096), j <= 1024
constexpr int MAX_LAGS = 512;      // lag_max - lag_min + 1 (415 for 50..800 Hz at 22.05 kHz)

// Number of analysis frames of an utterance of n samples: the signal is zero-padded to WIN, reflect-padded by
// WIN/2 on both sides and framed with hop HOP (:495-509).
KRF_HD int pitch_num_frames(long long n) {
  const long long L = n < WIN ? WIN : n;
  return (int)(L / HOP) + 1;
}

// Candidate lag range (:532-533); false when it does not fit the shared-memory CMND row.
KRF_HD bool pitch_lag_range(int sample_rate, float fmin, float fmax, int* lag_min, int* lag_max) {
  int lo = (int)((float)sample_rate / fmax);
  if (lo < 2) lo = 2;
  int hi = (int)((float)sample_rate / fmin);
  if (hi < lo + 1) hi = lo + 1;
  if (hi > WIN - 2) hi = WIN - 2;
  *lag_min = lo;
  *lag_max = hi;
  return hi - lo + 1 <= MAX_LAGS;
}

// ------------------------------------------------------------------------------------------------------------------
// In-place radix-4 FFT pair in shared memory, N = 4^m points, block-cooperative (N / 4 butterflies per stage spread over
// the block, one barrier per stage).  fft4_forward is decimation-in-frequency: natural order in, base-4 digit-reversed
// order out.  fft4_inverse is its exact mirror (decimation-in-time, conjugate twiddles, stages in reverse order):
// digit-reversed in, natural order out, unnormalised.  A pointwise operation between the two (the power spectrum of the
// autocorrelation) therefore needs no permutation at all; digit_reverse4() maps a natural bin index to its position for
// consumers that do (the mel filterbank).  Twiddles come from a QUARTER-wave table qw[j] = exp(-2 pi i j / N),
// j <= N / 4, rotated by multiples of -i for the other quadrants (exponents reach 3 N / 4).
// Half the shared-memory passes and barriers of a radix-2 transform of the same size.
// ------------------------------------------------------------------------------------------------------------------
KRF_DEV void fft4_fill_twiddles(krf_float2* qw, int N) {
  for (int j = KRF_TID; j <= N / 4; j += KRF_NT) {
    float s, c;
    krf_sincospi(-2.f * (float)j / (float)N, &s, &c);
    qw[j] = krf_make2(c, s);
  }
}

KRF_DEV krf_float2 fft4_twiddle(const krf_float2* qw, int e, int N) {      // exp(-2 pi i e / N), 0 <= e < N
  const int quarter = N / 4, quad = e / quarter;
  const krf_float2 t = qw[e - quad * quarter];
  if (quad == 0) return t;
  if (quad == 1) return krf_make2(t.y, -t.x);                              // * (-i)
  if (quad == 2) return krf_make2(-t.x, -t.y);
  return krf_make2(-t.y, t.x);                                             // * (+i)
}

KRF_DEV krf_float2 cmul(krf_float2 a, krf_float2 b) { return krf_make2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
KRF_DEV krf_float2 cmul_conj(krf_float2 a, krf_float2 b) {                 // a * conj(b)
  return krf_make2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

KRF_DEV void fft4_forward(krf_float2* z, const krf_float2* qw, int N, int log4n) {
  for (int st = log4n - 1; st >= 0; --st) {
    const int lq = 2 * st, q = 1 << lq;                                    // quarter size of this stage's blocks
    const int tw_step = N >> (lq + 2);                                     // N / (4 q)
    for (int i = KRF_TID; i < N / 4; i += KRF_NT) {
      const int k = i & (q - 1);
      const int base = ((i >> lq) << (lq + 2)) + k;
      const krf_float2 a0 = z[base], a1 = z[base + q], a2 = z[base + 2 * q], a3 = z[base + 3 * q];
      const krf_float2 t0 = krf_make2(a0.x + a2.x, a0.y + a2.y), t1 = krf_make2(a0.x - a2.x, a0.y - a2.y);
      const krf_float2 t2 = krf_make2(a1.x + a3.x, a1.y + a3.y);
      const krf_float2 t3 = krf_make2(a1.y - a3.y, -(a1.x - a3.x));        // -i (a1 - a3)
      const int e = k * tw_step;
      z[base] = krf_make2(t0.x + t2.x, t0.y + t2.y);
      z[base + q] = cmul(krf_make2(t1.x + t3.x, t1.y + t3.y), fft4_twiddle(qw, e, N));
      z[base + 2 * q] = cmul(krf_make2(t0.x - t2.x, t0.y - t2.y), fft4_twiddle(qw, 2 * e, N));
      z[base + 3 * q] = cmul(krf_make2(t1.x - t3.x, t1.y - t3.y), fft4_twiddle(qw, 3 * e, N));
    }
    KRF_SYNC();
  }
}

KRF_DEV void fft4_inverse(krf_float2* z, const krf_float2* qw, int N, int log4n) {
  for (int st = 0; st < log4n; ++st) {
    const int lq = 2 * st, q = 1 << lq;
    const int tw_step = N >> (lq + 2);
    for (int i = KRF_TID; i < N / 4; i += KRF_NT) {
      const int k = i & (q - 1);
      const int base = ((i >> lq) << (lq + 2)) + k;
      const int e = k * tw_step;
      const krf_float2 b0 = z[base];
      const krf_float2 b1 = cmul_conj(z[base + q], fft4_twiddle(qw, e, N));
      const krf_float2 b2 = cmul_conj(z[base + 2 * q], fft4_twiddle(qw, 2 * e, N));
      const krf_float2 b3 = cmul_conj(z[base + 3 * q], fft4_twiddle(qw, 3 * e, N));
      const krf_float2 t0 = krf_make2(b0.x + b2.x, b0.y + b2.y), t1 = krf_make2(b0.x - b2.x, b0.y - b2.y);
      const krf_float2 t2 = krf_make2(b1.x + b3.x, b1.y + b3.y);
      const krf_float2 t3 = krf_make2(-(b1.y - b3.y), b1.x - b3.x);        // +i (b1 - b3)
      z[base] = krf_make2(t0.x + t2.x, t0.y + t2.y);
      z[base + q] = krf_make2(t1.x + t3.x, t1.y + t3.y);
      z[base + 2 * q] = krf_make2(t0.x - t2.x, t0.y - t2.y);
      z[base + 3 * q] = krf_make2(t1.x - t3.x, t1.y - t3.y);
    }
    KRF_SYNC();
  }
}

// position of natural bin k in the output of fft4_forward: reverse the log4n base-4 digits of k
KRF_DEV int digit_reverse4(int k, int log4n) {
  int r = 0;
  for (int d = 0; d < log4n; ++d) { r = (r << 2) | (k & 3); k >>= 2; }
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// Log-mel frame on the radix-4 transform (kr_mel_stft): reflect pad 512, periodic Hann 1024, 1024-point FFT (5 radix-4 stages
// instead of 10 radix-2 ones), |X|^2 of the 513 one-sided bins read through the digit reversal, HTK filterbank (one warp
// per filter), log.  Reference data/dataset.py:162-178, 694-697.  Shared memory: z[1024], qw[257], pw[513].
// ------------------------------------------------------------------------------------------------------------------
constexpr int MEL_NFFT = 1024, MEL_LOG4 = 5, MEL_BINS = MEL_NFFT / 2 + 1;

// fb_ranges (optional, [n_mels][2] = first non-zero bin, one past the last): the HTK filters are triangles, i.e. ~13 of
// the 513 weights of a row are non-zero on average; the dense dot products were 40 % of the frame's work.  A lane keeps
// the bins it has in the dense loop (bin % 32 == lane), so skipping exact zeros leaves every partial sum bit-identical.
KRF_DEV void mel_frame_body(const float* x, long long n, int f, float gain, const float* fb_t, const int* fb_ranges,
                            int n_mels, long long out_stride, float log_eps, krf_float2* z, krf_float2* qw, float* pw,
                            float* orow) {
  const int tid = KRF_TID, nt = KRF_NT;
  fft4_fill_twiddles(qw, MEL_NFFT);
  for (int i = tid; i < MEL_NFFT; i += nt) {
    long long j = (long long)f * HOP + i - MEL_NFFT / 2;
    if (j < 0) j = -j;
    if (j >= n) j = 2 * (n - 1) - j;
    j = j < 0 ? 0 : j;
    const float w = 0.5f - 0.5f * krf_cospi(2.f * (float)i / (float)MEL_NFFT);
    z[i] = krf_make2(krf_ldg(x + j) * gain * w, 0.f);
  }
  KRF_SYNC();
  fft4_forward(z, qw, MEL_NFFT, MEL_LOG4);
  for (int k = tid; k < MEL_BINS; k += nt) {
    const krf_float2 u = z[digit_reverse4(k, MEL_LOG4)];
    pw[k] = u.x * u.x + u.y * u.y;
  }
  KRF_SYNC();
  for (int m = KRF_WARP; m < n_mels; m += KRF_NWARPS) {
    const float* frow = fb_t + (long long)m * MEL_BINS;
    float acc = 0.f;
    int lo = 0, hi = MEL_BINS;
    if (fb_ranges != nullptr) { lo = fb_ranges[2 * m]; hi = fb_ranges[2 * m + 1]; }
    for (int i = (lo / KRF_NLANES) * KRF_NLANES + KRF_LANE; i < hi; i += KRF_NLANES)
      if (i >= lo) acc = fmaf(pw[i], krf_ldg(frow + i), acc);
    acc = krf_warp_sum(acc);
    if (KRF_LANE == 0) orow[(long long)m * out_stride] = logf(acc + log_eps);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Per-frame analysis.  One block per (frame f, utterance b).  Shared memory: z[NFFT] complex, tw[TW] complex (quarter wave),
// cm[MAX_LAGS] floats, red[32] floats.  Outputs (one float each per frame): the frequency candidate
// before any voicing decision, the autocorrelation peak in the lag range, the mean energy of the windowed frame.
// ------------------------------------------------------------------------------------------------------------------
KRF_DEV void pitch_frame_body(const float* x, long long n, int f, int lag_min, int lag_max, float sample_rate,
                              krf_float2* z, krf_float2* tw, float* cm, float* red,
                              float* cand_out, float* acmax_out, float* energy_out) {
  const int tid = KRF_TID, nt = KRF_NT;
  const long long L = n < WIN ? WIN : n;
  // phase 0: quarter-wave twiddle table
  fft4_fill_twiddles(tw, NFFT);
  // phase 1: pre-emphasis (:499-503, applied AFTER the zero padding to WIN), reflect padding (:505-506),
  // periodic Hann (:512); upper half of the FFT buffer is the zero padding of the autocorrelation
  float e_part = 0.f;
  for (int i = tid; i < WIN; i += nt) {
    long long j = (long long)f * HOP + i - WIN / 2;
    if (j < 0) j = -j;
    if (j >= L) j = 2 * (L - 1) - j;
    const float x0 = j < n ? krf_ldg(x + j) : 0.f;
    float v = x0;
    if (j > 0) {
      const float x1 = (j - 1) < n ? krf_ldg(x + j - 1) : 0.f;
      v = x0 - krf_mul(0.97f, x1);
    }
    const float w = 0.5f - 0.5f * krf_cospi(2.f * (float)i / (float)WIN);
    v = krf_mul(v, w);
    z[i] = krf_make2(v, 0.f);
    z[i + WIN] = krf_make2(0.f, 0.f);
    e_part += krf_mul(v, v);
  }
  const float e_sum = krf_block_sum(e_part, red);      // (contains the barriers that publish z and tw)
  KRF_SYNC();
  // phase 2: forward radix-4 FFT (digit-reversed order out)
  fft4_forward(z, tw, NFFT, LOG4_NFFT);
  // phase 3: power spectrum (pointwise, so the digit-reversed order does not matter)
  for (int i = tid; i < NFFT; i += nt) {
    const krf_float2 u = z[i];
    z[i] = krf_make2(u.x * u.x + u.y * u.y, 0.f);
  }
  KRF_SYNC();
  // phase 4: inverse radix-4 FFT (digit-reversed in, natural order out, unnormalised)
  fft4_inverse(z, tw, NFFT, LOG4_NFFT);
  // phase 5: cumulative mean normalised difference over the lag range (:522-529).  diff(tau) = 2 acf(0) - 2 acf(tau);
  // cmnd(tau) = diff(tau) / (cumsum(diff)[tau] / tau + 1e-8).  The running sum over tau < lag_min is one short serial
  // loop (torch accumulates a float cumsum in double on the CPU; so does this).
  const float inv_n = 1.f / (float)NFFT;
  const float acf0 = z[0].x * inv_n;
  if (tid == 0) {
    double cs = 0.0;
    for (int tau = 1; tau <= lag_max; ++tau) {
      const float d = 2.f * acf0 - 2.f * (z[tau].x * inv_n);
      cs += (double)d;
      if (tau >= lag_min) cm[tau - lag_min] = d / ((float)cs / (float)tau + 1e-8f);
    }
  }
  KRF_SYNC();
  // phase 6 (one warp): autocorrelation peak (:539-541), first dip below 0.15 else the global minimum (:544-550),
  // parabolic interpolation around it (:553-563)
  if (KRF_WARP == 0) {
    const int n_lags = lag_max - lag_min + 1;
    const float zden = fmaxf(acf0, 1e-8f);
    float ac_max = -3.402823466e38f, c_min = 3.402823466e38f;
    int first_dip = 0x7fffffff, arg_min = 0x7fffffff;
    for (int i = KRF_LANE; i < n_lags; i += KRF_NLANES) {
      ac_max = fmaxf(ac_max, (z[lag_min + i].x * inv_n) / zden);
      const float c = cm[i];
      if (c < 0.15f && i < first_dip) first_dip = i;
      if (c < c_min) { c_min = c; arg_min = i; }
    }
    ac_max = krf_warp_max(ac_max);
    first_dip = krf_warp_min_int(first_dip);
    const float all_min = krf_warp_min(c_min);
    arg_min = krf_warp_min_int(c_min == all_min ? arg_min : 0x7fffffff);
    if (KRF_LANE == 0) {
      const int best = first_dip != 0x7fffffff ? first_dip : (arg_min != 0x7fffffff ? arg_min : 0);
      const int pi = best > 0 ? best - 1 : 0, ni = best + 1 < n_lags ? best + 1 : n_lags - 1;
      const float a = cm[pi], bb = cm[best], g = cm[ni];
      const float denom = fmaxf(a - 2.f * bb + g, 1e-8f);
      float off = 0.5f * (a - g) / denom;
      off = fminf(fmaxf(off, -1.f), 1.f);
      const float best_lag = fmaxf((float)(lag_min + best) + off, 1.f);
      *cand_out = sample_rate / best_lag;
      *acmax_out = ac_max;
      *energy_out = e_sum / (float)WIN;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Order statistics by rank counting: out[r] = sorted(x[0..T))[ranks[r]] for up to 4 ranks, one block.  Every element's
// rank is unique (ties broken by index), so exactly one thread writes each result.  O(T^2 / threads): T is an
// utterance's frame count (a few hundred to 2000), the whole thing is a few microseconds.
// ------------------------------------------------------------------------------------------------------------------
KRF_DEV void select_ranks(const float* x, int T, const int* ranks, int n_ranks, float* out) {
  for (int i = KRF_TID; i < T; i += KRF_NT) {
    const float xi = x[i];
    int rank = 0;
    for (int j = 0; j < T; ++j) {
      const float xj = x[j];
      rank += (xj < xi || (xj == xi && j < i)) ? 1 : 0;
    }
    for (int r = 0; r < n_ranks; ++r)
      if (rank == ranks[r]) out[r] = xi;
  }
  KRF_SYNC();
}

// torch.quantile(x, q) with linear interpolation: position q * (T - 1) in fp32, torch.lerp between the neighbours
KRF_DEV void quantile_ranks(float q, int T, int* lo, int* hi, float* w) {
  const float pos = q * (float)(T - 1);
  int l = (int)floorf(pos);
  if (l > T - 1) l = T - 1;
  *lo = l;
  *hi = l + 1 < T ? l + 1 : T - 1;
  *w = pos - (float)l;
}
KRF_DEV float lerp_torch(float a, float b, float w) { return w < 0.5f ? a + w * (b - a) : b - (b - a) * (1.f - w); }

// Gap-filled frequency at frame t (:582-606): an unvoiced frame between two voiced ones at most 5 unvoiced frames apart
// is interpolated linearly; `v` holds the voiced-gated frequencies (0 = unvoiced).
KRF_DEV float gap_filled(const float* v, int T, int t) {
  const float cur = v[t];
  if (cur > 0.f) return cur;
  int p = -1, q = -1;
  for (int d = 1; d <= 5 && p < 0; ++d)
    if (t - d >= 0 && v[t - d] > 0.f) p = t - d;
  for (int d = 1; d <= 5 && q < 0; ++d)
    if (t + d < T && v[t + d] > 0.f) q = t + d;
  if (p < 0 || q < 0 || q - p - 1 > 5) return cur;
  const float dd = fmaxf((float)(q - p), 1.f);
  const float tt = (float)(t - p) / dd;
  return v[p] * (1.f - tt) + v[q] * tt;
}

KRF_DEV float median5(float a, float b, float c, float d, float e) {
  float s[5] = {a, b, c, d, e};
  for (int i = 1; i < 5; ++i) {           // insertion sort of 5
    const float key = s[i];
    int j = i - 1;
    while (j >= 0 && s[j] > key) { s[j + 1] = s[j]; --j; }
    s[j + 1] = key;
  }
  return s[2];
}

// ------------------------------------------------------------------------------------------------------------------
// Per-utterance tracking: adaptive voicing threshold from the 25 % quantile of the autocorrelation peaks (:566-567),
// energy gate from the (lower) median frame energy (:569-572), range gate (:577-579), gap interpolation, median-5
// (reflect padding), normalisation to [0, 1] with 0 = unvoiced (:609-615).  One block per utterance; `work` is a
// T-float scratch row in global memory, `sel` 4 floats of shared memory.
// ------------------------------------------------------------------------------------------------------------------
KRF_DEV void pitch_track_body(const float* cand, const float* acmax, const float* energy, int T, int T_max,
                              float fmin, float fmax, float* work, float* sel, float* out) {
  const int tid = KRF_TID, nt = KRF_NT;
  int rk[3];
  float wq;
  quantile_ranks(0.25f, T, &rk[0], &rk[1], &wq);
  select_ranks(acmax, T, rk, 2, sel);
  rk[2] = (T - 1) / 2;
  select_ranks(energy, T, rk + 2, 1, sel + 2);
  float thr = lerp_torch(sel[0], sel[1], wq) * 0.8f;
  thr = fminf(fmaxf(thr, 0.15f), 0.35f);
  const float e_thr = fmaxf(sel[2] * 0.05f, 1e-9f);
  for (int t = tid; t < T; t += nt) {
    float fr = cand[t];
    if (acmax[t] < thr || energy[t] < e_thr) fr = 0.f;
    if (fr < fmin || fr > fmax) fr = 0.f;
    work[t] = fr;
  }
  KRF_SYNC();
  const float scale = (float)((double)fmax - (double)fmin + 1e-8);
  for (int t = tid; t < T_max; t += nt) {
    if (t >= T) { out[t] = 0.f; continue; }
    float m[5];
    for (int k = -2; k <= 2; ++k) {
      int i = t + k;
      if (i < 0) i = -i;
      if (i >= T) i = 2 * (T - 1) - i;
      m[k + 2] = gap_filled(work, T, i);
    }
    const float med = median5(m[0], m[1], m[2], m[3], m[4]);
    float o = (med - fmin) / scale;
    o = fminf(fmaxf(o, 0.f), 1.f);
    out[t] = med == 0.f ? 0.f : o;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Energy: per-frame mean over the mel bins (:664 log-domain input; :667-668 log1p of the clamped mean for linear power,
// which is what the dataset passes, data/dataset.py:813), then 5 / 95 percentile normalisation per utterance (:675-687).
// exp_input = 1: the input is the log-mel the mel-STFT kernel wrote and the LINEAR-power semantics are wanted.
// ------------------------------------------------------------------------------------------------------------------
KRF_DEV float energy_finish(float mean, int log_domain) {
  return log_domain ? mean : krf_log1p(fmaxf(mean, 0.f));
}

KRF_DEV void energy_norm_body(const float* e, int T, int T_max, float* sel, float* red, float* out) {
  const int tid = KRF_TID, nt = KRF_NT;
  float lo_v, hi_v;
  if (T < 3) {
    float mx = -3.402823466e38f, mn = -3.402823466e38f;
    for (int t = tid; t < T; t += nt) { mx = fmaxf(mx, e[t]); mn = fmaxf(mn, -e[t]); }
    hi_v = krf_block_max(mx, red);
    lo_v = -krf_block_max(mn, red);
  } else {
    int rk[4];
    float w0, w1;
    quantile_ranks(0.05f, T, &rk[0], &rk[1], &w0);
    quantile_ranks(0.95f, T, &rk[2], &rk[3], &w1);
    select_ranks(e, T, rk, 4, sel);
    lo_v = lerp_torch(sel[0], sel[1], w0);
    hi_v = lerp_torch(sel[2], sel[3], w1);
  }
  const float den = fmaxf(hi_v - lo_v, 1e-8f);
  for (int t = tid; t < T_max; t += nt) {
    if (t >= T) { out[t] = 0.f; continue; }
    const float o = (e[t] - lo_v) / den;
    out[t] = fminf(fmaxf(o, 0.f), 1.f);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Trailing-silence trim of a generated mel before vocoding (reference inference/inference.py:590-621): from the per-frame
// means e[0..T) (kr_energy_frames, log-domain branch) the threshold clamp(0.5 * (q10 + q20), -9.8, -9.2), the last frame
// above it, and t_end = min(T, max(60, last + 24 + 1)); T when no frame is above the threshold.  One block per utterance.
// ------------------------------------------------------------------------------------------------------------------
KRF_DEV void trim_end_body(const float* e, int T, float* sel, float* red, int* t_end) {
  if (T <= 0) { if (KRF_TID == 0) *t_end = 0; return; }
  int rk[4];
  float w0, w1;
  quantile_ranks(0.10f, T, &rk[0], &rk[1], &w0);
  quantile_ranks(0.20f, T, &rk[2], &rk[3], &w1);
  select_ranks(e, T, rk, 4, sel);
  const float q10 = lerp_torch(sel[0], sel[1], w0), q20 = lerp_torch(sel[2], sel[3], w1);
  const float thr = fmaxf(-9.8f, fminf(-9.2f, 0.5f * (q10 + q20)));
  float last = -1.f;                                       // frame indices are exact in float (T < 2^24)
  for (int t = KRF_TID; t < T; t += KRF_NT)
    if (e[t] > thr) last = fmaxf(last, (float)t);
  last = krf_block_max(last, red);
  if (KRF_TID == 0) {
    int end = T;
    if (last >= 0.f) {
      const int proposed = (int)last + 24 + 1 < T ? (int)last + 24 + 1 : T;
      end = proposed > 60 ? proposed : 60;
      end = end < T ? end : T;
    }
    *t_end = end;
  }
}

}  // namespace krf
