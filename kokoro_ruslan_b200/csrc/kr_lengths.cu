// kr_lengths.cu — average_by_duration on the device (reference utils/lengths.py:156-208).  Body and the reference
// quirks it reproduces: kr_lengths_core.cuh.  (The LengthRegulator index / gather kernels of the same reference file
// live in kr_variance.cu next to their consumers.)
#include "kr_common.cuh"
#include "kr_lengths_core.cuh"

namespace {
__global__ void __launch_bounds__(256)
average_by_duration_kernel(const float* __restrict__ values, const long long* __restrict__ dur,
                           const unsigned char* __restrict__ mask, int P, int T, int* label, float* __restrict__ out) {
  kr::pdl_entry();
  const int b = blockIdx.x;
  krl::average_by_duration_body(values + (long long)b * T, dur + (long long)b * P,
                                mask != nullptr ? mask + (long long)b * P : nullptr, P, T, label + (long long)b * T,
                                out + (long long)b * P);
}
}  // namespace

// values [B, T] fp32, durations [B, P] int64, mask [B, P] bytes (1 = padded token) or null, label [B, T] int32 scratch,
// out [B, P] fp32.
extern "C" int kr_average_by_duration(const float* values, const long long* durations, const unsigned char* mask,
                                      int* label, float* out, int B, int P, int T, void* stream) {
  if (B <= 0 || P <= 0) return KR_OK;
  if (T <= 0) { kr_set_error("kr_average_by_duration: no frames"); return KR_ERR_ARG; }
  kr::launch(average_by_duration_kernel, dim3(B), 256, 0, (cudaStream_t)stream, values, durations, mask, P, T, label, out);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
