// kr_lengths_core.cuh — body of average_by_duration (reference utils/lengths.py:156-208): frame-level values averaged to
// token level through the durations — the inverse direction of the LengthRegulator, used by the loss code for
// token-level pitch / energy targets (training/losses.py:19, model/model.py:415-431).
//
// An exact restatement of the reference's scatter formulation, quirks included: starts are clamped to T - 1 and ends to
// T, +j / -j are scattered at the clamped starts / ends, the running sum is the frame's token label.  Consequences
// that a "segment mean" would get wrong: frames not covered by any token carry label 0 and are averaged into token 0;
// tokens that start beyond the last frame all pile their labels onto frame T - 1, whose label then is a sum of indices.
// Sums run in ascending frame order per token (the order of the CPU scatter_add_), so the result is bit-identical.
// DUAL-COMPILED like kr_features_core.cuh (g++ -DKR_HOST_EMU -> tests/emu/lengths_emu.cpp).
#pragma once

#ifdef KR_HOST_EMU
#define KRL_DEV static inline
#ifdef KR_HOST_EMU_SIMT            // one host thread per CUDA thread (tests/emu/emu_simt.h)
#include "emu_simt.h"
#define KRL_TID (emu::tid())
#define KRL_NT (emu::nthreads())
#define KRL_SYNC() emu::syncthreads()
#else
#define KRL_TID 0
#define KRL_NT 1
#define KRL_SYNC() do { } while (0)
#endif
#else
#define KRL_DEV __device__ __forceinline__
#define KRL_TID ((int)threadIdx.x)
#define KRL_NT ((int)blockDim.x)
#define KRL_SYNC() __syncthreads()
#endif

namespace krl {

// One block per utterance.  label: T ints of scratch (global memory), written by thread 0 and read by all after the
// barrier.  dur: P int64, mask: P bytes or nullptr, values: T floats, out: P floats.
KRL_DEV void average_by_duration_body(const float* values, const long long* dur, const unsigned char* mask, int P, int T,
                                      int* label, float* out) {
  if (KRL_TID == 0) {
    // labels = cumsum over frames of the scattered +-j; done as one ordered sweep over the (sorted) token boundaries:
    // every token adds j on [min(start, T - 1), min(end, T)).  Integer arithmetic — the reference's float32 sums of
    // integers below 2^24 are exact, so this is the same number.
    for (int f = 0; f < T; ++f) label[f] = 0;
    long long end = 0;
    for (int j = 0; j < P; ++j) {
      long long d = dur[j] < 0 ? 0 : dur[j];
      long long start = end;
      end += d;
      long long s = start > T - 1 ? T - 1 : start;
      long long e = end > T ? T : end;
      for (long long f = s; f < e; ++f) label[f] += j;          // a difference array would do; ranges are disjoint
    }                                                            // except on frame T - 1, so this is O(T + P) anyway
  }
  KRL_SYNC();
  for (int j = KRL_TID; j < P; j += KRL_NT) {
    float sum = 0.f, cnt = 0.f;
    for (int f = 0; f < T; ++f) {
      int l = label[f];
      if (l >= P) continue;                                      // valid_frame_mask, :194 (adds exact zeros)
      if (l == j) { sum += values[f]; cnt += 1.f; }
    }
    float o = sum / (cnt < 1.f ? 1.f : cnt);
    const long long d = dur[j] < 0 ? 0 : dur[j];
    if (mask != nullptr ? mask[j] != 0 : d == 0) o = 0.f;         // :203-206
    out[j] = o;
  }
}

}  // namespace krl
