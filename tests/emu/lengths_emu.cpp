// TEST INFRASTRUCTURE — host emulation of kr_average_by_duration (kokoro_ruslan_b200/csrc/kr_lengths.cu): the kernel's
// own body (kr_lengths_core.cuh) compiled with -DKR_HOST_EMU.  Used by tests/test_lengths_emu_cpu.py only.
#define KR_HOST_EMU 1
#include "kr_lengths_core.cuh"

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

extern "C" int emu_average_by_duration(const float* values, const long long* durations, const unsigned char* mask,
                                       int* label, float* out, int B, int P, int T) {
  for (int b = 0; b < B; ++b)
    EMU_BLOCK(256, krl::average_by_duration_body(values + (long long)b * T, durations + (long long)b * P,
                                                 mask ? mask + (long long)b * P : nullptr, P, T, label + (long long)b * T,
                                                 out + (long long)b * P));
  return 0;
}
