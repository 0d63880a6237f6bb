// TEST INFRASTRUCTURE — host emulation of kr_val_metrics (kokoro_ruslan_b200/csrc/kr_metrics.cu): the kernel's own body
// (kr_metrics_core.cuh) compiled with -DKR_HOST_EMU, blocks run in order.  Used by tests/test_metrics_emu_cpu.py only.
#define KR_HOST_EMU 1
#include "kr_metrics_core.cuh"

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

extern "C" int emu_val_metrics_acc_floats(void) { return krm::ACC_FLOATS; }

extern "C" int emu_val_metrics(const float* mel_pred, const float* mel_tgt, const float* pitch_pred, const float* pitch_tgt,
                               const long long* mel_lengths, float* acc, int B, int T, int Tp, int C) {
  if (B > krm::MAX_B) return -4;
  float red[32];
  for (int b = 0; b < B; ++b) {
    const long long o = (long long)b * T * C;
    EMU_BLOCK(512, krm::utterance_metrics(mel_pred + o, mel_tgt + o, pitch_pred ? pitch_pred + (long long)b * Tp : nullptr,
                                          pitch_tgt ? pitch_tgt + (long long)b * T : nullptr, mel_lengths[b], T, Tp, C, red,
                                          acc + krm::ACC_HEAD + 2 * b));
    const unsigned ticket = krm_arrive(reinterpret_cast<unsigned*>(acc + 4));
    if (ticket == (unsigned)(B - 1)) {
      krm::fold_batch(acc, B);
      *reinterpret_cast<unsigned*>(acc + 4) = 0u;
    }
  }
  return 0;
}
