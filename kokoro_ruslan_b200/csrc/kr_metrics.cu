// kr_metrics.cu — epoch-level validation metrics as device reductions (reference
// training/trainer.py:1868-1916).  HBM-bound: reads the predicted and target mel once (2 x B x T x 80 x 4 B).
// Body in kr_metrics_core.cuh (also compiled as a host emulation by the CPU tests).
#include "kr_common.cuh"
#include "kr_metrics_core.cuh"

namespace {
constexpr int THREADS = 512;

__global__ void __launch_bounds__(THREADS)
val_metrics_kernel(const float* __restrict__ mel_pred, const float* __restrict__ mel_tgt,
                   const float* __restrict__ pitch_pred, const float* __restrict__ pitch_tgt,
                   const long long* __restrict__ mel_lengths, float* acc, int T, int Tp, int C) {
  kr::pdl_entry();
  __shared__ float red[32];
  __shared__ unsigned ticket;
  const int b = blockIdx.x, B = gridDim.x;
  const long long o = (long long)b * T * C;
  krm::utterance_metrics(mel_pred + o, mel_tgt + o, pitch_pred != nullptr ? pitch_pred + (long long)b * Tp : nullptr,
                         pitch_tgt != nullptr ? pitch_tgt + (long long)b * T : nullptr, mel_lengths[b], T, Tp, C, red,
                         acc + krm::ACC_HEAD + 2 * b);
  if (threadIdx.x == 0) ticket = krm_arrive(reinterpret_cast<unsigned*>(acc + 4));
  __syncthreads();
  if (ticket == (unsigned)(B - 1) && threadIdx.x == 0) {
    __threadfence();                                        // acquire the other blocks' scratch rows
    krm::fold_batch(acc, B);
    *reinterpret_cast<unsigned*>(acc + 4) = 0u;             // ready for the next batch
  }
}
}  // namespace

extern "C" int kr_val_metrics_acc_floats(void) { return krm::ACC_FLOATS; }

// acc: krm::ACC_FLOATS floats, zeroed by the caller at the start of a validation epoch; after any number of calls
// acc[0] / acc[1] = mean spectral convergence, acc[2] / acc[3] = mean F0 RMSE (a zero count = metric undefined).
// mel_pred / mel_tgt [B, T, C], pitch_pred [B, Tp] (or null), pitch_tgt [B, T] (or null), mel_lengths [B] int64.
extern "C" int kr_val_metrics(const float* mel_pred, const float* mel_tgt, const float* pitch_pred, const float* pitch_tgt,
                              const long long* mel_lengths, float* acc, int B, int T, int Tp, int C, void* stream) {
  if (B <= 0) return KR_OK;
  if (B > krm::MAX_B) { kr_set_error("kr_val_metrics: at most 60 utterances per batch"); return KR_ERR_UNSUPPORTED; }
  kr::launch(val_metrics_kernel, dim3(B), THREADS, 0, (cudaStream_t)stream, mel_pred, mel_tgt, pitch_pred, pitch_tgt,
             mel_lengths, acc, T, Tp, C);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
