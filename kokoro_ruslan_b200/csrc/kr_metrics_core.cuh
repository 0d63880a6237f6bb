// kr_metrics_core.cuh — body of the validation-metrics kernel: the per-batch spectral convergence
// and frame-level F0 RMSE of KokoroTrainer.validate_epoch (reference training/trainer.py:1868-1916), which the reference
// computes with a Python loop over the batch and four `.item()` host syncs per utterance.  Here: one block per
// utterance reduces its sums, the last block to finish folds the batch means into epoch accumulators on the device; the
// host reads four floats once per epoch.  DUAL-COMPILED like kr_features_core.cuh (g++ -DKR_HOST_EMU -> tests/emu/).
//
// acc layout (floats): [0] sum over batches of the batch-mean spectral convergence, [1] number of such batches,
// [2] / [3] the same for the F0 RMSE, [4] blocks-done counter (as unsigned), [8 + 2b] / [9 + 2b] per-utterance scratch.
#pragma once

#ifdef KR_HOST_EMU
#include <math.h>
#define KRM_DEV static inline
#ifdef KR_HOST_EMU_SIMT            // one host thread per CUDA thread (tests/emu/emu_simt.h)
#include "emu_simt.h"
#define KRM_TID (emu::tid())
#define KRM_NT (emu::nthreads())
KRM_DEV float krm_block_sum(float v, float* red) { return emu::block_sum(v, red); }
#else
#define KRM_TID 0
#define KRM_NT 1
KRM_DEV float krm_block_sum(float v, float*) { return v; }
#endif
KRM_DEV unsigned krm_arrive(unsigned* counter) { return (*counter)++; }      // blocks run in order: the last one folds
#else
#define KRM_DEV __device__ __forceinline__
#define KRM_TID ((int)threadIdx.x)
#define KRM_NT ((int)blockDim.x)
KRM_DEV float krm_block_sum(float v, float* red) { return kr::block_sum(v, red); }
KRM_DEV unsigned krm_arrive(unsigned* counter) {         // release our scratch writes, then count ourselves in
  __threadfence();
  return atomicAdd(counter, 1u);
}
#endif

namespace krm {

constexpr int ACC_HEAD = 8, MAX_B = 60, ACC_FLOATS = ACC_HEAD + 2 * MAX_B;

// Block b of n_blocks.  mel_*: this utterance's (T, C) rows; pitch_*: its (T) / (Tp) rows or nullptr.
// Returns (through *last) whether this block arrived last, in which case the caller folds with fold_batch().
KRM_DEV void utterance_metrics(const float* mel_pred, const float* mel_tgt, const float* pitch_pred,
                               const float* pitch_tgt, long long L, int T, int Tp, int C, float* red, float* out2) {
  if (L > T) L = T;
  float num = 0.f, den = 0.f, pe = 0.f;
  const long long n = L > 0 ? L * C : 0;
  for (long long i = KRM_TID; i < n; i += KRM_NT) {
    const float r = mel_tgt[i], d = r - mel_pred[i];
    num += d * d;
    den += r * r;
  }
  const bool f0_ok = pitch_pred != nullptr && pitch_tgt != nullptr && L > 0 && L <= Tp;
  if (f0_ok)
    for (long long t = KRM_TID; t < L; t += KRM_NT) { const float d = pitch_tgt[t] - pitch_pred[t]; pe += d * d; }
  num = krm_block_sum(num, red);
  den = krm_block_sum(den, red);
  pe = krm_block_sum(pe, red);
  if (KRM_TID == 0) {
    out2[0] = (L > 0 && den > 0.f) ? sqrtf(num) / sqrtf(den) : -1.f;     // ||ref - pred||_F / ||ref||_F, trainer.py:1879-1884
    out2[1] = f0_ok ? sqrtf(pe / (float)L) : -1.f;                       // sqrt(mean((tgt - pred)^2)), :1905-1907
  }
}

// Batch means of the valid utterances -> epoch accumulators (trainer.py:1887-1889, 1910-1912); one thread.
KRM_DEV void fold_batch(volatile float* acc, int B) {       // volatile: the scratch rows come from other SMs (L2)
  float sc = 0.f, f0 = 0.f;
  int n_sc = 0, n_f0 = 0;
  for (int b = 0; b < B; ++b) {
    const float s = acc[ACC_HEAD + 2 * b], r = acc[ACC_HEAD + 2 * b + 1];
    if (s >= 0.f) { sc += s; ++n_sc; }
    if (r >= 0.f) { f0 += r; ++n_f0; }
  }
  if (n_sc > 0) { acc[0] += sc / (float)n_sc; acc[1] += 1.f; }
  if (n_f0 > 0) { acc[2] += f0 / (float)n_f0; acc[3] += 1.f; }
}

}  // namespace krm
