// TEST INFRASTRUCTURE — host emulation of the decode-step kernels (kokoro_ruslan_b200/csrc/kr_decode.cu): the kernels'
// own bodies (kr_decode_core.cuh) compiled with -DKR_HOST_EMU, looped over the grid like the launch wrappers do.
// Built and called by tests/test_decode_emu_cpu.py only.
#define KR_HOST_EMU 1
#include "kr_decode_core.cuh"
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

extern "C" int emu_dec_state_size(void) { return (int)sizeof(krd::DecState); }

extern "C" int emu_dec_feed(const void* state, const float* prev, const float* forced, int forced_T, const float* w_in,
                            const float* b_in, const float* pe, float* x, int B, int D, int n_mels) {
  const krd::DecState* st = (const krd::DecState*)state;
  if (st->done) return 0;
  for (int b = 0; b < B; ++b) {
    const float* frame = forced ? forced + ((long long)b * forced_T + st->t) * n_mels : prev + (long long)b * n_mels;
    EMU_BLOCK(256, krd::dec_feed_body(st, frame, w_in, b_in, pe, D, n_mels, x + (long long)b * D));
  }
  return 0;
}

extern "C" int emu_dec_attn(const void* state, const uint16_t* q, long long ld_q, const uint16_t* k_raw,
                            const uint16_t* v_raw, long long ld_kv, const float* gq, const float* gk, const float* gv,
                            const float* cos_t, const float* sin_t, uint16_t* kc, uint16_t* vc, long long cache_ld,
                            long long cache_bs, int n_keys, const unsigned char* mask, uint16_t* o, long long ld_o, int B,
                            int H, float scale, int rotate_q) {
  const krd::DecState* st = (const krd::DecState*)state;
  if (st->done) return 0;
  float qs[krd::DK], wm[krd::MAX_WARPS], wl[krd::MAX_WARPS], wacc[krd::MAX_WARPS * krd::DK];
  const bool self = n_keys < 0;
  const int t = st->t;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h) {
      const int col = h * krd::DK;
      EMU_BLOCK(128, krd::dec_attn_body(q + (long long)b * ld_q + col, gq, self ? k_raw + (long long)b * ld_kv + col : nullptr, gk,
                         self ? v_raw + (long long)b * ld_kv + col : nullptr, gv,
                         self ? cos_t + (long long)t * (krd::DK / 2) : nullptr,
                         self ? sin_t + (long long)t * (krd::DK / 2) : nullptr, kc + (long long)b * cache_bs + col,
                         vc + (long long)b * cache_bs + col, cache_ld, self ? t + 1 : n_keys, self ? t : -1,
                         mask ? mask + (long long)b * n_keys : nullptr, scale, 1.1920929e-7f, rotate_q, qs, wm, wl, wacc,
                         o + (long long)b * ld_o + col));
    }
  return 0;
}

extern "C" int emu_dec_finish(void* state, const float* y, const float* ln_g, const float* ln_b, const float* w_out,
                              const float* b_out, const float* w_stop, const float* b_stop, float* mel_out,
                              float* next_frame, float* probs, int B, int D, int n_mels, int t_cap) {
  if (B > krd::MAX_B || n_mels > 128) return -4;
  std::vector<float> stats(2 * krd::MAX_B), vals(krd::MAX_B * 129), red(32);
  EMU_BLOCK(256, krd::dec_finish_body((krd::DecState*)state, y, ln_g, ln_b, w_out, b_out, w_stop, b_stop, B, D, n_mels, t_cap,
                                      stats.data(), vals.data(), red.data(), mel_out, next_frame, probs));
  return 0;
}

extern "C" int emu_dec_gemv(const void* state, const uint16_t* x, const float* x_f32, long long ld_x, const float* ln_g,
                            const float* ln_b, const uint16_t* w, const float* bias, const float* resid, long long ld_r,
                            void* out, long long ld_o, int out_f32, int glu, int B, int N, int K) {
  if (B > krd::GEMV_MAX_B || K % 8 != 0) return -4;
  if ((x == nullptr) == (x_f32 == nullptr) || (glu && (resid != nullptr || out_f32))) return -1;
  if (state && ((const krd::DecState*)state)->done) return 0;
  std::vector<uint16_t> xs((size_t)B * K);
#ifdef KR_HOST_EMU_SIMT
  // like the launch wrapper: blocks of 8 warps, block i starts at feature 8 i, stride = blocks * 8
  const int blocks = (N + 7) / 8 < 4 ? (N + 7) / 8 : 4;
  for (int blk = 0; blk < blocks; ++blk)
    EMU_BLOCK(256, krd::dec_gemv_body(x, x_f32, ld_x, ln_g, ln_b, w, bias, resid, ld_r, out, ld_o, out_f32, glu, B, N, K,
                                      blk * 8, blocks * 8, xs.data()));
#else
  // the launch wrapper runs `blocks` blocks of 8 warps; one emulated block with n_step = 1 covers every feature once
  krd::dec_gemv_body(x, x_f32, ld_x, ln_g, ln_b, w, bias, resid, ld_r, out, ld_o, out_f32, glu, B, N, K, 0, 1, xs.data());
#endif
  return 0;
}
