// TEST INFRASTRUCTURE — host emulation of kr_average_by_duration (kokoro_ruslan_b200/csrc/kr_lengths.cu): the kernel's
// own body (kr_lengths_core.cuh) compiled with -DKR_HOST_EMU.  Used by tests/test_lengths_emu_cpu.py only.
#define KR_HOST_EMU 1
#include "kr_lengths_core.cuh"

extern "C" int emu_average_by_duration(const float* values, const long long* durations, const unsigned char* mask,
                                       int* label, float* out, int B, int P, int T) {
  for (int b = 0; b < B; ++b)
    krl::average_by_duration_body(values + (long long)b * T, durations + (long long)b * P,
                                  mask ? mask + (long long)b * P : nullptr, P, T, label + (long long)b * T,
                                  out + (long long)b * P);
  return 0;
}
