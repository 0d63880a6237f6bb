// TEST INFRASTRUCTURE — ThreadSanitizer run of the SIMT emulation (tests/emu/emu_simt.h): every dual-compiled kernel body
// executes once or twice with its real block geometry on host threads, built with -fsanitize=thread.  pthread barriers
// are synchronisation TSan understands, so a shared-memory (or scratch-buffer) hand-off that lacks a __syncthreads() in
// the kernel source shows up as a reported data race here — the host-side counterpart of compute-sanitizer racecheck.
// Built and run by tests/test_emu_tsan_cpu.py; exit code 66 = race reported (TSAN_OPTIONS), 0 = clean.
#define KR_HOST_EMU 1
#define KR_HOST_EMU_SIMT 1
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "decode_emu.cpp"
#undef EMU_BLOCK
#define emu_pool emu_pool_features
#include "features_emu.cpp"
#undef EMU_BLOCK
#undef emu_pool
#define emu_pool emu_pool_metrics
#include "metrics_emu.cpp"
#undef EMU_BLOCK
#undef emu_pool
#define emu_pool emu_pool_lengths
#define gcd_ll gcd_ll_unused
#include "lengths_emu.cpp"

static float frand() { return (float)rand() / (float)RAND_MAX - 0.5f; }

int main() {
  srand(1);
  // ---- features: 3 pitch frames of a noisy tone, tracker, energy, trim, radix-4 mel frame
  const long long n = 2600;
  std::vector<float> wav(n);
  for (long long i = 0; i < n; ++i) wav[i] = 0.4f * sinf(6.2831853f * 180.f * (float)i / 22050.f) + 0.01f * frand();
  const int T = emu_pitch_num_frames(n);
  std::vector<float> cand(T), ac(T), en(T), work(T), pitch(T);
  emu_pitch_frames(wav.data(), nullptr, cand.data(), ac.data(), en.data(), 1, n, T, 22050, 50.f, 800.f);
  emu_pitch_track(cand.data(), ac.data(), en.data(), nullptr, work.data(), pitch.data(), 1, n, T, 50.f, 800.f);
  std::vector<float> mel(40 * 80), e(40), eo(40);
  for (auto& v : mel) v = -6.f + 3.f * frand();
  emu_energy_frames(mel.data(), e.data(), 1, 40, 80, 1, 0, 1);
  emu_energy_norm(e.data(), nullptr, eo.data(), 1, 40);
  long long two = 2;
  emu_energy_norm(e.data(), &two, eo.data(), 1, 40);          // the min / max branch for fewer than 3 frames
  int t_end = -1;
  emu_trim_end(e.data(), nullptr, &t_end, 1, 40);
  std::vector<float> fb(80 * 513, 0.001f), lm(80 * 3);
  emu_mel_stft_r4(wav.data(), nullptr, nullptr, fb.data(), nullptr, lm.data(), 1, n, 3, 80, 1e-9f);
  // ---- decode: two self-attention steps with cache append, a masked cross-attention, fused GEMVs, finish
  const int D = 128, H = 2, B = 2, cap = 8, Tp = 5, M = 80;
  krd::DecState st = {};
  st.lo = 1; st.hi = 6; st.expected = 3; st.stop_thr = 2.f; st.post_thr = 2.f;
  std::vector<uint16_t> qkv(B * 3 * D), kc(B * cap * D), vc(B * cap * D), o(B * D), xk(B * Tp * 2 * D), w(2 * D * D), u(B * D);
  auto fill16 = [](std::vector<uint16_t>& v) { for (auto& x : v) x = krd_f2b(frand()); };
  fill16(qkv); fill16(xk); fill16(w);
  std::vector<float> g(64, 1.f), cs(16 * 32, 0.8f), sn(16 * 32, 0.6f), x(B * D), y(B * D), ln(D, 1.f), lb(D, 0.f), bias(2 * D, 0.1f);
  for (auto& v : x) v = frand();
  std::vector<unsigned char> mask(B * Tp, 0);
  mask[Tp - 1] = 1;
  std::vector<float> w_in(D * M), b_in(D, 0.f), pe(16 * D, 0.f), prev(B * M, 0.f), w_out(M * D), b_out(M, -3.f), w_stop(D), b_stop(1, 0.f),
      mel_out(B * cap * M), nxt(B * M), probs(cap);
  for (auto& v : w_in) v = 0.1f * frand();
  for (auto& v : w_out) v = 0.1f * frand();
  for (auto& v : w_stop) v = 0.1f * frand();
  for (int step = 0; step < 2; ++step) {
    emu_dec_feed(&st, prev.data(), nullptr, 0, w_in.data(), b_in.data(), pe.data(), x.data(), B, D, M);
    emu_dec_gemv(&st, nullptr, x.data(), D, ln.data(), lb.data(), w.data(), nullptr, nullptr, 0, qkv.data(), 3 * D, 0, 0, B, 3 * D > 2 * D ? 2 * D : 3 * D, D);
    emu_dec_attn(&st, qkv.data(), 3 * D, qkv.data() + D, qkv.data() + 2 * D, 3 * D, g.data(), g.data(), g.data(), cs.data(), sn.data(),
                 kc.data(), vc.data(), D, (long long)cap * D, -1, nullptr, o.data(), D, B, H, 0.125f, step);
    emu_dec_attn(&st, qkv.data(), 3 * D, nullptr, nullptr, 0, g.data(), nullptr, nullptr, nullptr, nullptr, xk.data(), xk.data() + D,
                 2 * D, (long long)Tp * 2 * D, Tp, mask.data(), o.data(), D, B, H, 0.125f, 0);
    emu_dec_gemv(&st, o.data(), nullptr, D, nullptr, nullptr, w.data(), bias.data(), x.data(), D, y.data(), D, 1, 0, B, D, D);
    emu_dec_gemv(&st, nullptr, y.data(), D, ln.data(), lb.data(), w.data(), bias.data(), nullptr, 0, u.data(), D, 0, 1, B, D, D);
    emu_dec_finish(&st, y.data(), ln.data(), lb.data(), w_out.data(), b_out.data(), w_stop.data(), b_stop.data(), mel_out.data(),
                   nxt.data(), probs.data(), B, D, M, cap);
  }
  // ---- validation metrics, average_by_duration
  std::vector<float> mp(2 * 30 * 80), mt(2 * 30 * 80), pp(2 * 30), pt(2 * 30), acc(emu_val_metrics_acc_floats(), 0.f);
  for (auto& v : mp) v = frand();
  for (auto& v : mt) v = frand();
  long long lens[2] = {30, 11};
  emu_val_metrics(mp.data(), mt.data(), pp.data(), pt.data(), lens, acc.data(), 2, 30, 30, 80);
  long long dur[2 * 6] = {3, 0, 5, 9, 2, 30, 1, 1, 1, 1, 1, 1};
  std::vector<int> label(2 * 30);
  std::vector<float> avg(2 * 6);
  emu_average_by_duration(pp.data(), dur, nullptr, label.data(), avg.data(), 2, 6, 30);
  unsigned char tok_mask[2 * 6] = {0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
  emu_average_by_duration(pp.data(), dur, tok_mask, label.data(), avg.data(), 2, 6, 30);
  printf("simt emulation done: t=%d frames=%d t_end=%d acc=%.3f pitch0=%.3f\n", st.t, T, t_end, acc[1], pitch[1]);
  return 0;
}
