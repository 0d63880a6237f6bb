// TEST INFRASTRUCTURE — host emulation of the feature-extraction kernels (kokoro_ruslan_b200/csrc/kr_features.cu).
// Compiles the kernels' own bodies (kr_features_core.cuh) with -DKR_HOST_EMU, i.e. as one "thread" per block, and
// loops over the grid exactly as the launch wrappers do.  Built and called by tests/test_features_emu_cpu.py only.
#define KR_HOST_EMU 1
#include "kr_features_core.cuh"
#include <vector>

// Block execution: one sequential "thread" (default) or, with -DKR_HOST_EMU_SIMT, the real block size on a pool of host
// threads (tests/emu/emu_simt.h).  EMU_BLOCK(n, stmt) runs `stmt` as one thread block of n threads.
#ifdef KR_HOST_EMU_SIMT
#include <map>
#include <memory>
static emu::Pool& emu_pool(int n) {
  static std::map<int, std::unique_ptr<emu::Pool>> pools;
  auto& p = pools[n];
  if (!p) p.reset(new emu::Pool(n));
  return *p;
}
#define EMU_BLOCK(n, stmt) emu_pool(n).run([&] { stmt; })
#else
#define EMU_BLOCK(n, stmt) do { stmt; } while (0)
#endif

extern "C" int emu_pitch_num_frames(long long n) { return krf::pitch_num_frames(n); }

extern "C" int emu_pitch_frames(const float* wav, const long long* lengths, float* cand, float* acmax, float* energy,
                                int B, long long n_max, int frames_max, int sample_rate, float fmin, float fmax) {
  int lag_min, lag_max;
  if (!krf::pitch_lag_range(sample_rate, fmin, fmax, &lag_min, &lag_max)) return -4;
  std::vector<krf_float2> z(krf::NFFT), tw(krf::TW);
  std::vector<float> cm(krf::MAX_LAGS), red(32);
  for (int b = 0; b < B; ++b) {
    const long long n = lengths ? lengths[b] : n_max;
    for (int f = 0; f < frames_max; ++f) {
      if (f >= krf::pitch_num_frames(n)) continue;
      const long long o = (long long)b * frames_max + f;
      EMU_BLOCK(512, krf::pitch_frame_body(wav + (long long)b * n_max, n, f, lag_min, lag_max, (float)sample_rate, z.data(),
                                           tw.data(), cm.data(), red.data(), cand + o, acmax + o, energy + o));
    }
  }
  return 0;
}

extern "C" int emu_pitch_track(const float* cand, const float* acmax, const float* energy, const long long* lengths,
                               float* work, float* out, int B, long long n_max, int frames_max, float fmin, float fmax) {
  float sel[4];
  for (int b = 0; b < B; ++b) {
    const long long n = lengths ? lengths[b] : n_max;
    int T = krf::pitch_num_frames(n);
    T = T < frames_max ? T : frames_max;
    const long long o = (long long)b * frames_max;
    EMU_BLOCK(256, krf::pitch_track_body(cand + o, acmax + o, energy + o, T, frames_max, fmin, fmax, work + o, sel, out + o));
  }
  return 0;
}

extern "C" int emu_energy_frames(const float* mel, float* e, int B, int T, int n_mels, int time_major, int exp_input,
                                 int log_domain) {
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t) {
      float acc = 0.f;
      for (int m = 0; m < n_mels; ++m) {
        const float v = time_major ? mel[((long long)b * T + t) * n_mels + m] : mel[((long long)b * n_mels + m) * T + t];
        acc += exp_input ? krf_exp(v) : v;
      }
      e[(long long)b * T + t] = krf::energy_finish(acc / (float)n_mels, log_domain);
    }
  return 0;
}

extern "C" int emu_energy_norm(const float* e, const long long* frames, float* out, int B, int T_max) {
  float sel[4], red[32];
  for (int b = 0; b < B; ++b) {
    long long T = frames ? frames[b] : T_max;
    T = T < 0 ? 0 : (T < T_max ? T : T_max);
    EMU_BLOCK(256, krf::energy_norm_body(e + (long long)b * T_max, (int)T, T_max, sel, red, out + (long long)b * T_max));
  }
  return 0;
}

// kr_mel_stft, KR_MELSTFT_R4=1 variant (mel_stft_r4_kernel in kr_melstft.cu)
extern "C" int emu_mel_stft_r4(const float* wav, const long long* lengths, const float* peak, const float* fb_t,
                               const int* fb_ranges, float* out, int B, long long n_max, int frames_max, int n_mels,
                               float log_eps) {
  std::vector<krf_float2> z(krf::MEL_NFFT), qw(krf::MEL_NFFT / 4 + 1);
  std::vector<float> pw(krf::MEL_BINS + 3);
  for (int b = 0; b < B; ++b) {
    const long long n = lengths ? lengths[b] : n_max;
    const float gain = peak ? 1.f / (peak[b] + 1e-9f) : 1.f;
    for (int f = 0; f < frames_max; ++f) {
      float* orow = out + ((long long)b * n_mels) * frames_max + f;
      if (f >= 1 + (int)(n / krf::HOP)) {
        for (int m = 0; m < n_mels; ++m) orow[(long long)m * frames_max] = 0.f;
        continue;
      }
      EMU_BLOCK(256, krf::mel_frame_body(wav + (long long)b * n_max, n, f, gain, fb_t, fb_ranges, n_mels, frames_max, log_eps, z.data(),
                                         qw.data(), pw.data(), orow));
    }
  }
  return 0;
}

extern "C" int emu_trim_end(const float* e, const long long* frames, int* t_end, int B, int T_max) {
  float sel[4], red[32];
  for (int b = 0; b < B; ++b) {
    long long T = frames ? frames[b] : T_max;
    T = T < 0 ? 0 : (T < T_max ? T : T_max);
    EMU_BLOCK(256, krf::trim_end_body(e + (long long)b * T_max, (int)T, sel, red, t_end + b));
  }
  return 0;
}
