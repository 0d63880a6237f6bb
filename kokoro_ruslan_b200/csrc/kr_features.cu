// kr_features.cu — on-device pitch / energy feature extraction, the row after the mel-STFT:
// PitchExtractor.extract_pitch (reference model/variance_predictor.py:448-625) and
// EnergyExtractor.extract_energy_from_mel (:633-688) as the dataset uses them (data/dataset.py:793-815).
//
// HBM-bound integer/float work, no tensor cores: per analysis frame 8 KB of waveform is read once (L2-shared with the
// 7 neighbouring frames that overlap it) and 12 bytes leave the SM; the 4096-point autocorrelation FFT pair lives in
// shared memory.  The kernel bodies are in kr_features_core.cuh, which the CPU test-suite also compiles as a host
// emulation (tests/emu/) — see the header for what that does and does not prove.
//
//   kr_pitch_frames   grid (frames_max, B) x 512 threads: one CTA per analysis frame
//   kr_pitch_track    grid B x 256 threads: per-utterance thresholds (rank-counting order statistics), gap fill, median
//   kr_energy_frames  one warp per frame (time-major) or one thread per frame (channel-major, coalesced along time)
//   kr_energy_norm    grid B x 256 threads: 5 / 95 percentile normalisation
#include "kr_common.cuh"
#include "kr_features_core.cuh"

namespace {
using namespace kr;

constexpr int FRAME_THREADS = 512, TRACK_THREADS = 256;

__global__ void __launch_bounds__(FRAME_THREADS)
pitch_frames_kernel(const float* __restrict__ wav, const long long* __restrict__ lengths, float* __restrict__ cand,
                    float* __restrict__ acmax, float* __restrict__ energy, long long n_max, int frames_max,
                    int lag_min, int lag_max, float sample_rate) {
  kr::pdl_entry();
  __shared__ float2 z[krf::NFFT];          // 32 KB
  __shared__ float2 tw[krf::TW];           // 8 KB
  __shared__ float cm[krf::MAX_LAGS];
  __shared__ float red[32];
  const int f = blockIdx.x, b = blockIdx.y;
  const long long n = lengths != nullptr ? lengths[b] : n_max;
  if (f >= krf::pitch_num_frames(n)) return;            // padding frames of a ragged batch: the tracker zero-fills
  const long long o = (long long)b * frames_max + f;
  krf::pitch_frame_body(wav + (long long)b * n_max, n, f, lag_min, lag_max, sample_rate, z, tw, cm, red,
                        cand + o, acmax + o, energy + o);
}

__global__ void __launch_bounds__(TRACK_THREADS)
pitch_track_kernel(const float* __restrict__ cand, const float* __restrict__ acmax, const float* __restrict__ energy,
                   const long long* __restrict__ lengths, float* __restrict__ work, float* __restrict__ out,
                   long long n_max, int frames_max, float fmin, float fmax) {
  kr::pdl_entry();
  __shared__ float sel[4];
  const int b = blockIdx.x;
  const long long n = lengths != nullptr ? lengths[b] : n_max;
  int T = krf::pitch_num_frames(n);
  T = T < frames_max ? T : frames_max;
  const long long o = (long long)b * frames_max;
  krf::pitch_track_body(cand + o, acmax + o, energy + o, T, frames_max, fmin, fmax, work + o, sel, out + o);
}

// mel is (B, n_mels, T) [time_major = 0] or (B, T, n_mels) [time_major = 1]
__global__ void energy_frames_kernel(const float* __restrict__ mel, float* __restrict__ e, int T, int n_mels,
                                     int time_major, int exp_input, int log_domain) {
  kr::pdl_entry();
  const int b = blockIdx.y;
  if (time_major) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * warps + (threadIdx.x >> 5);
    if (t >= T) return;
    const float* row = mel + ((long long)b * T + t) * n_mels;
    float acc = 0.f;
    for (int m = lane; m < n_mels; m += 32) {
      const float v = __ldg(row + m);
      acc += exp_input ? krf_exp(v) : v;
    }
    acc = warp_sum(acc);
    if (lane == 0) e[(long long)b * T + t] = krf::energy_finish(acc / (float)n_mels, log_domain);
  } else {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float* col = mel + (long long)b * n_mels * T + t;
    float acc = 0.f;
    for (int m = 0; m < n_mels; ++m) {
      const float v = __ldg(col + (long long)m * T);
      acc += exp_input ? krf_exp(v) : v;
    }
    e[(long long)b * T + t] = krf::energy_finish(acc / (float)n_mels, log_domain);
  }
}

__global__ void __launch_bounds__(TRACK_THREADS)
energy_norm_kernel(const float* __restrict__ e, const long long* __restrict__ frames, float* __restrict__ out, int T_max) {
  kr::pdl_entry();
  __shared__ float sel[4];
  __shared__ float red[32];
  const int b = blockIdx.x;
  long long T = frames != nullptr ? frames[b] : T_max;
  T = T < 0 ? 0 : (T < T_max ? T : T_max);
  const long long o = (long long)b * T_max;
  krf::energy_norm_body(e + o, (int)T, T_max, sel, red, out + o);
}

__global__ void __launch_bounds__(TRACK_THREADS)
trim_end_kernel(const float* __restrict__ e, const long long* __restrict__ frames, int* __restrict__ t_end, int T_max) {
  kr::pdl_entry();
  __shared__ float sel[4];
  __shared__ float red[32];
  const int b = blockIdx.x;
  long long T = frames != nullptr ? frames[b] : T_max;
  T = T < 0 ? 0 : (T < T_max ? T : T_max);
  krf::trim_end_body(e + (long long)b * T_max, (int)T, sel, red, t_end + b);
}

}  // namespace

// cand / acmax / energy: [B, frames_max] fp32 per-frame intermediates (entries of frames beyond an utterance's own
// 1 + max(len, 2048) // 256 are left untouched and never read).
extern "C" int kr_pitch_frames(const float* wav, const long long* lengths, float* cand, float* acmax, float* energy,
                               int B, long long n_max, int frames_max, int sample_rate, int hop, float fmin, float fmax,
                               void* stream) {
  if (hop != krf::HOP) { kr_set_error("kr_pitch_frames: built for hop 256 (analysis window 2048)"); return KR_ERR_UNSUPPORTED; }
  if (B <= 0 || frames_max <= 0) return KR_OK;
  if (n_max <= 0 || sample_rate <= 0 || !(fmin > 0.f) || !(fmax > fmin)) { kr_set_error("kr_pitch_frames: bad arguments"); return KR_ERR_ARG; }
  int lag_min, lag_max;
  if (!krf::pitch_lag_range(sample_rate, fmin, fmax, &lag_min, &lag_max)) {
    kr_set_error("kr_pitch_frames: more than 512 candidate lags");
    return KR_ERR_UNSUPPORTED;
  }
  kr::launch(pitch_frames_kernel, dim3(frames_max, B), FRAME_THREADS, 0, (cudaStream_t)stream, wav, lengths, cand, acmax,
             energy, n_max, frames_max, lag_min, lag_max, (float)sample_rate);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// work: [B, frames_max] fp32 scratch; out: [B, frames_max] normalised pitch in [0, 1], 0 = unvoiced / padding.
extern "C" int kr_pitch_track(const float* cand, const float* acmax, const float* energy, const long long* lengths,
                              float* work, float* out, int B, long long n_max, int frames_max, float fmin, float fmax,
                              void* stream) {
  if (B <= 0 || frames_max <= 0) return KR_OK;
  kr::launch(pitch_track_kernel, dim3(B), TRACK_THREADS, 0, (cudaStream_t)stream, cand, acmax, energy, lengths, work, out,
             n_max, frames_max, fmin, fmax);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_energy_frames(const float* mel, float* e, int B, int T, int n_mels, int time_major, int exp_input,
                                int log_domain, void* stream) {
  if (B <= 0 || T <= 0) return KR_OK;
  if (n_mels <= 0) { kr_set_error("kr_energy_frames: n_mels must be positive"); return KR_ERR_ARG; }
  const dim3 grid(time_major ? (T + 7) / 8 : (T + 255) / 256, B);
  kr::launch(energy_frames_kernel, grid, 256, 0, (cudaStream_t)stream, mel, e, T, n_mels, time_major, exp_input, log_domain);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// frames[b] = valid frames of utterance b (nullptr: all T); out beyond them is zero.
extern "C" int kr_energy_norm(const float* e, const long long* frames, float* out, int B, int T, void* stream) {
  if (B <= 0 || T <= 0) return KR_OK;
  kr::launch(energy_norm_kernel, dim3(B), TRACK_THREADS, 0, (cudaStream_t)stream, e, frames, out, T);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// e: [B, T] per-frame mel means (kr_energy_frames with log_domain = 1); t_end[b] = frames to keep of utterance b.
extern "C" int kr_trim_end(const float* e, const long long* frames, int* t_end, int B, int T, void* stream) {
  if (B <= 0) return KR_OK;
  if (T <= 0) { kr_set_error("kr_trim_end: no frames"); return KR_ERR_ARG; }
  kr::launch(trim_end_kernel, dim3(B), TRACK_THREADS, 0, (cudaStream_t)stream, e, frames, t_end, T);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
