// kr_decode.cu — autoregressive decode-step kernels (reference model/generator.py:24-127,
// model/transformers.py:237-277, model/model.py:675-779).  Bodies and the design notes are in kr_decode_core.cuh.
//
// A decode step is latency- and weight-bandwidth-bound (<= 16 rows against 27 M decoder weights that stay in L2): no
// tensor-core shape here.  The step counter, the KV-cache fill level and the generator's stop rules live in a device
// struct (krd::DecState), so a step has no host round trip and no per-step pointer arithmetic on the host: the whole
// step (these kernels + the shared training-path kernels) is ONE CUDA graph replayed per frame, and the host only looks
// at `done` every few dozen frames.  After `done` every kernel here returns without touching memory.
//
//   kr_dec_feed    grid B          x 256   decoder input of frame t
//   kr_dec_attn    grid (H, B)     x 128   one (utterance, head) per CTA, 4 warps stride the cached keys
//   kr_dec_finish  grid 1          x 256   output heads + stop rules + t += 1
#include "kr_common.cuh"
#include "kr_decode_core.cuh"

namespace {
using namespace kr;

constexpr int ATTN_THREADS = 128, FEED_THREADS = 256, FINISH_THREADS = 256, MAX_MELS = 128;

__global__ void __launch_bounds__(FEED_THREADS)
dec_feed_kernel(const krd::DecState* __restrict__ st, const float* __restrict__ prev, const float* __restrict__ forced,
                int forced_T, const float* __restrict__ w_in, const float* __restrict__ b_in,
                const float* __restrict__ pe, float* __restrict__ x, int D, int n_mels) {
  kr::pdl_entry();
  if (st->done) return;
  const int b = blockIdx.x;
  const float* frame = forced != nullptr ? forced + ((long long)b * forced_T + st->t) * n_mels : prev + (long long)b * n_mels;
  krd::dec_feed_body(st, frame, w_in, b_in, pe, D, n_mels, x + (long long)b * D);
}

__global__ void __launch_bounds__(ATTN_THREADS)
dec_attn_kernel(const krd::DecState* __restrict__ st, const bf16* __restrict__ q, long long ld_q,
                const bf16* __restrict__ k_raw, const bf16* __restrict__ v_raw, long long ld_kv,
                const float* __restrict__ gq, const float* __restrict__ gk, const float* __restrict__ gv,
                const float* __restrict__ cos_t, const float* __restrict__ sin_t, bf16* kc, bf16* vc, long long cache_ld,
                long long cache_bs, int n_keys, const unsigned char* __restrict__ mask, bf16* __restrict__ o,
                long long ld_o, float scale, int rotate_q) {
  kr::pdl_entry();
  __shared__ float qs[krd::DK], wm[krd::MAX_WARPS], wl[krd::MAX_WARPS], wacc[krd::MAX_WARPS * krd::DK];
  if (st->done) return;
  const int h = blockIdx.x, b = blockIdx.y, col = h * krd::DK;
  const bool self = n_keys < 0;
  const int t = st->t;
  krd::dec_attn_body(q + (long long)b * ld_q + col, gq,
                     self ? k_raw + (long long)b * ld_kv + col : nullptr, gk,
                     self ? v_raw + (long long)b * ld_kv + col : nullptr, gv,
                     self ? cos_t + (long long)t * (krd::DK / 2) : nullptr, self ? sin_t + (long long)t * (krd::DK / 2) : nullptr,
                     kc + (long long)b * cache_bs + col, vc + (long long)b * cache_bs + col, cache_ld,
                     self ? t + 1 : n_keys, self ? t : -1, mask != nullptr ? mask + (long long)b * n_keys : nullptr,
                     scale, 1.1920929e-7f, rotate_q, qs, wm, wl, wacc, o + (long long)b * ld_o + col);
}

__global__ void __launch_bounds__(FINISH_THREADS)
dec_finish_kernel(krd::DecState* st, const float* __restrict__ y, const float* __restrict__ ln_g,
                  const float* __restrict__ ln_b, const float* __restrict__ w_out, const float* __restrict__ b_out,
                  const float* __restrict__ w_stop, const float* __restrict__ b_stop, float* __restrict__ mel_out,
                  float* __restrict__ next_frame, float* __restrict__ probs, int B, int D, int n_mels, int t_cap) {
  kr::pdl_entry();
  __shared__ float stats[2 * krd::MAX_B], vals[krd::MAX_B * (MAX_MELS + 1)], red[32];
  krd::dec_finish_body(st, y, ln_g, ln_b, w_out, b_out, w_stop, b_stop, B, D, n_mels, t_cap, stats, vals, red, mel_out,
                       next_frame, probs);
}

constexpr int GEMV_THREADS = 256, GEMV_MAX_K = 1536;

__global__ void __launch_bounds__(GEMV_THREADS)
dec_gemv_kernel(const krd::DecState* __restrict__ st, const bf16* __restrict__ x, const float* __restrict__ x_f32,
                long long ld_x, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                const bf16* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ resid, long long ld_r,
                void* __restrict__ out, long long ld_o, int out_f32, int glu, int B, int N, int K) {
  kr::pdl_entry();
  __shared__ __align__(16) bf16 xs[krd::GEMV_MAX_B * GEMV_MAX_K];       // 24 KB
  if (st != nullptr && st->done) return;
  const int warps = GEMV_THREADS / 32;
  krd::dec_gemv_body(x, x_f32, ld_x, ln_g, ln_b, w, bias, resid, ld_r, out, ld_o, out_f32, glu, B, N, K,
                     blockIdx.x * warps, gridDim.x * warps, xs);
}

}  // namespace

// out[b, n] (bf16 or fp32, row stride ld_o) = x[b, :K] . w[n, :K] (bf16, dense) + bias[n] + resid[b, n] (fp32, row stride
// ld_r); B <= 8, K % 8 == 0, K <= 1536.  The activation rows are x (bf16, row stride ld_x) or, when x_f32 is given,
// LayerNorm(x_f32[b]; ln_g, ln_b) computed in the kernel's prologue.  glu != 0: w is [2 * N, K], bias [2 * N], and
// out[b, n] = gelu_erf(gate_n) * lin_n (bf16, no residual).  `state` (optional): skipped once done.
extern "C" int kr_dec_gemv(const void* state, const void* x, const float* x_f32, long long ld_x, const float* ln_g,
                           const float* ln_b, const void* w, const float* bias, const float* resid, long long ld_r,
                           void* out, long long ld_o, int out_f32, int glu, int B, int N, int K, void* stream) {
  if (B <= 0 || N <= 0) return KR_OK;
  if (B > krd::GEMV_MAX_B || K <= 0 || K % 8 != 0 || K > GEMV_MAX_K) {
    kr_set_error("kr_dec_gemv: needs B <= 8, K % 8 == 0, K <= 1536");
    return KR_ERR_UNSUPPORTED;
  }
  if ((x == nullptr) == (x_f32 == nullptr) || (x_f32 != nullptr && (ln_g == nullptr || ln_b == nullptr)) ||
      (glu && (resid != nullptr || out_f32))) {
    kr_set_error("kr_dec_gemv: inconsistent operand combination");
    return KR_ERR_ARG;
  }
  const int warps = GEMV_THREADS / 32;
  int blocks = (N + warps - 1) / warps;
  if (blocks > 2 * kr::kNumSMs) blocks = 2 * kr::kNumSMs;
  kr::launch(dec_gemv_kernel, dim3(blocks), GEMV_THREADS, 0, (cudaStream_t)stream, (const krd::DecState*)state,
             (const bf16*)x, x_f32, ld_x, ln_g, ln_b, (const bf16*)w, bias, resid, ld_r, out, ld_o, out_f32, glu, B, N, K);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_dec_state_size(void) { return (int)sizeof(krd::DecState); }

// x[b, :D] (fp32 residual stream rows) = w_in[D, n_mels] . frame_b + b_in + pe[t]; frame_b = forced[b, t, :] when
// `forced` [B, forced_T, n_mels] is given (teacher-forced decode), else prev[b, :] (the previous output frame).
extern "C" int kr_dec_feed(const void* state, const float* prev, const float* forced, int forced_T, const float* w_in,
                           const float* b_in, const float* pe, float* x, int B, int D, int n_mels, void* stream) {
  if (B <= 0) return KR_OK;
  kr::launch(dec_feed_kernel, dim3(B), FEED_THREADS, 0, (cudaStream_t)stream, (const krd::DecState*)state, prev, forced,
             forced_T, w_in, b_in, pe, x, D, n_mels);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// n_keys < 0: self-attention of the new frame — k_raw / v_raw (this step's projections, row stride ld_kv) are normalised,
// rotated (key) and appended to the caches at row t, scores over rows 0..t.  n_keys >= 0: cross-attention over n_keys
// pre-normalised memory keys / values with the key-padding mask [B, n_keys] (1 = masked).  rotate_q != 0 (self-attention):
// rotate the new query to position t like training does, instead of the reference decode's position 0.  Caches: bf16, key j of
// utterance b at kc + b * cache_bs + j * cache_ld (+ head * 64).
extern "C" int kr_dec_attn(const void* state, const void* q, long long ld_q, const void* k_raw, const void* v_raw,
                           long long ld_kv, const float* gq, const float* gk, const float* gv, const float* cos_t,
                           const float* sin_t, void* kc, void* vc, long long cache_ld, long long cache_bs, int n_keys,
                           const unsigned char* mask, void* o, long long ld_o, int B, int H, float scale, int rotate_q,
                           void* stream) {
  if (B <= 0 || H <= 0) return KR_OK;
  if (n_keys < 0 && (k_raw == nullptr || v_raw == nullptr || cos_t == nullptr || sin_t == nullptr)) {
    kr_set_error("kr_dec_attn: self-attention needs the new key / value projections and the RoPE tables");
    return KR_ERR_ARG;
  }
  kr::launch(dec_attn_kernel, dim3(H, B), ATTN_THREADS, 0, (cudaStream_t)stream, (const krd::DecState*)state,
             (const bf16*)q, ld_q, (const bf16*)k_raw, (const bf16*)v_raw, ld_kv, gq, gk, gv, cos_t, sin_t, (bf16*)kc,
             (bf16*)vc, cache_ld, cache_bs, n_keys, mask, (bf16*)o, ld_o, scale, rotate_q);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// mel_out [B, t_cap, n_mels] (clamped frames), next_frame [B, n_mels] (un-clamped feedback), probs [t_cap].
extern "C" int kr_dec_finish(void* state, const float* y, const float* ln_g, const float* ln_b, const float* w_out,
                             const float* b_out, const float* w_stop, const float* b_stop, float* mel_out,
                             float* next_frame, float* probs, int B, int D, int n_mels, int t_cap, void* stream) {
  if (B <= 0) return KR_OK;
  if (B > krd::MAX_B || n_mels > MAX_MELS) { kr_set_error("kr_dec_finish: at most 16 utterances / 128 mel bins"); return KR_ERR_UNSUPPORTED; }
  kr::launch(dec_finish_kernel, dim3(1), FINISH_THREADS, 0, (cudaStream_t)stream, (krd::DecState*)state, y, ln_g, ln_b,
             w_out, b_out, w_stop, b_stop, mel_out, next_frame, probs, B, D, n_mels, t_cap);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
