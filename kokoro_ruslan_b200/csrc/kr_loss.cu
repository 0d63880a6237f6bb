// kr_loss.cu — stop-token head and the fused training losses + their gradients (HBM-bound,
// warp-shuffle reductions, no host synchronisation).
//   reference: training/losses.py:9-216 (masked L1 mel, Huber(1.0) on log1p durations,
//   BCE-with-logits pos_weight, Huber(0.05) pitch/energy, clamps 100/100/100/10/10, weighted sum),
//   criteria training/trainer.py:410-444, stop head on a detached input model/model.py:561-562.
//   SURVEY.md §9 S4.
#include "kr_common.cuh"

namespace {
using namespace kr;
constexpr int WARPS = 8;

__device__ __forceinline__ float softplus(float z) { return z > 0.f ? z + log1pf(__expf(-z)) : log1pf(__expf(z)); }
__device__ __forceinline__ float sigmoidf(float z) { return 1.f / (1.f + __expf(-z)); }
__device__ __forceinline__ float huber(float e, float d) {
  const float a = fabsf(e);
  return a <= d ? 0.5f * e * e : d * (a - 0.5f * d);
}
__device__ __forceinline__ float huber_grad(float e, float d) { return fminf(fmaxf(e, -d), d); }

struct LossParams {
  const float* mel_pred; const float* mel_tgt;        // [B,T,C]
  const float* dur_pred; const long long* dur_tgt;    // [B,P]
  const float* stop_pred; const float* stop_tgt;      // [B,T]
  const float* pitch_pred; const float* pitch_tgt;    // pred [B,Tp], tgt [B,Tt]
  const float* energy_pred; const float* energy_tgt;
  const long long* mel_len; const long long* ph_len;  // [B]
  int B, T, P, C, Tp, Tt;
  float w_dur, w_stop, w_pitch, w_energy, pos_weight, delta_var;
  double* acc;        // [10] (sum,count) x (mel,dur,stop,pitch,energy)
  float* losses;      // [6] total, mel, dur, stop, pitch, energy
  const float* loss_scale;  // device scalar multiplied into every gradient (may be null -> 1)
  bf16* dmel;         // [B,T,C]
  float* ddur;        // [B,P]
  float* dstop;       // [B,T]
  float* dpitch; float* denergy;  // [B,Tp]
};

__device__ __forceinline__ void block_accumulate(float s, float n, double* dst) {
  __shared__ float red[2][32];
  s = warp_sum(s); n = warp_sum(n);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) { red[0][w] = s; red[1][w] = n; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    float a = l < nw ? red[0][l] : 0.f, b = l < nw ? red[1][l] : 0.f;
    a = warp_sum(a); b = warp_sum(b);
    if (l == 0 && b > 0.f) { atomicAdd(dst, (double)a); atomicAdd(dst + 1, (double)b); }
  }
}

__global__ void loss_reduce_kernel(const LossParams p) {
  kr::pdl_entry();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  float s = 0.f, n = 0.f;
  for (long long i = tid; i < (long long)p.B * p.T * p.C; i += nth) {
    const long long bt = i / p.C;
    const int b = (int)(bt / p.T), t = (int)(bt % p.T);
    const float l = fabsf(p.mel_pred[i] - p.mel_tgt[i]);
    if (t < p.mel_len[b] && isfinite(l)) { s += l; n += 1.f; }
  }
  block_accumulate(s, n, p.acc + 0);
  s = 0.f; n = 0.f;
  for (long long i = tid; i < (long long)p.B * p.P; i += nth) {
    const int b = (int)(i / p.P), j = (int)(i % p.P);
    const long long d = p.dur_tgt[i];
    const float l = huber(p.dur_pred[i] - logf((float)d + 1.f), 1.f);
    if (j < p.ph_len[b] && d > 0 && isfinite(l)) { s += l; n += 1.f; }
  }
  block_accumulate(s, n, p.acc + 2);
  s = 0.f; n = 0.f;
  float sp = 0.f, np = 0.f, se = 0.f, ne = 0.f;
  for (long long i = tid; i < (long long)p.B * p.T; i += nth) {
    const int b = (int)(i / p.T), t = (int)(i % p.T);
    const bool in = t < p.mel_len[b];
    const float z = p.stop_pred[i], y = p.stop_tgt[i];
    const float l = p.pos_weight * y * softplus(-z) + (1.f - y) * softplus(z);
    if (in && isfinite(l)) { s += l; n += 1.f; }
    if (p.pitch_pred != nullptr) {
      const float lp = huber(p.pitch_pred[(long long)b * p.Tp + t] - p.pitch_tgt[(long long)b * p.Tt + t], p.delta_var);
      const float le = huber(p.energy_pred[(long long)b * p.Tp + t] - p.energy_tgt[(long long)b * p.Tt + t], p.delta_var);
      if (in && isfinite(lp)) { sp += lp; np += 1.f; }
      if (in && isfinite(le)) { se += le; ne += 1.f; }
    }
  }
  block_accumulate(s, n, p.acc + 4);
  block_accumulate(sp, np, p.acc + 6);
  block_accumulate(se, ne, p.acc + 8);
}

__global__ void loss_grad_kernel(const LossParams p) {
  kr::pdl_entry();
  // every thread derives the five means / clamp gates from the 10 accumulators
  float L[5], f[5];
  const float cl[5] = {100.f, 100.f, 100.f, 10.f, 10.f};
  const float w[5] = {1.f, p.w_dur, p.w_stop, p.w_pitch, p.w_energy};
  const float ls = p.loss_scale != nullptr ? *p.loss_scale : 1.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const double cnt = p.acc[2 * i + 1];
    const float m = cnt > 0 ? (float)(p.acc[2 * i] / cnt) : 0.f;
    L[i] = fminf(m, cl[i]);
    f[i] = (cnt > 0 && m <= cl[i]) ? w[i] * ls / (float)cnt : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.losses[0] = L[0] + w[1] * L[1] + w[2] * L[2] + w[3] * L[3] + w[4] * L[4];
    for (int i = 0; i < 5; ++i) p.losses[1 + i] = L[i];
  }
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < (long long)p.B * p.T * p.C; i += nth) {
    const long long bt = i / p.C;
    const int b = (int)(bt / p.T), t = (int)(bt % p.T);
    const float e = p.mel_pred[i] - p.mel_tgt[i];
    float g = 0.f;
    if (t < p.mel_len[b] && isfinite(e)) g = (e > 0.f ? f[0] : (e < 0.f ? -f[0] : 0.f));
    p.dmel[i] = __float2bfloat16(g);
  }
  for (long long i = tid; i < (long long)p.B * p.P; i += nth) {
    const int b = (int)(i / p.P), j = (int)(i % p.P);
    const long long d = p.dur_tgt[i];
    const float e = p.dur_pred[i] - logf((float)d + 1.f);
    p.ddur[i] = (j < p.ph_len[b] && d > 0 && isfinite(e)) ? huber_grad(e, 1.f) * f[1] : 0.f;
  }
  for (long long i = tid; i < (long long)p.B * p.T; i += nth) {
    const int b = (int)(i / p.T), t = (int)(i % p.T);
    const bool in = t < p.mel_len[b];
    const float z = p.stop_pred[i], y = p.stop_tgt[i];
    const float sg = sigmoidf(z);
    p.dstop[i] = (in && isfinite(z)) ? ((1.f - y) * sg - p.pos_weight * y * (1.f - sg)) * f[2] : 0.f;
  }
  if (p.pitch_pred != nullptr) {
    for (long long i = tid; i < (long long)p.B * p.Tp; i += nth) {
      const int b = (int)(i / p.Tp), t = (int)(i % p.Tp);
      float gp = 0.f, ge = 0.f;
      if (t < p.T && t < p.mel_len[b]) {
        const float ep = p.pitch_pred[i] - p.pitch_tgt[(long long)b * p.Tt + t];
        const float ee = p.energy_pred[i] - p.energy_tgt[(long long)b * p.Tt + t];
        if (isfinite(ep)) gp = huber_grad(ep, p.delta_var) * f[3];
        if (isfinite(ee)) ge = huber_grad(ee, p.delta_var) * f[4];
      }
      p.dpitch[i] = gp;
      p.denergy[i] = ge;
    }
  }
}

// z[n] = x[n,:] . w + b   (x bf16 [N,D])
__global__ void stop_head_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ bias, float* __restrict__ z, int N, int D) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  for (int n = blockIdx.x * WARPS + (threadIdx.x >> 5); n < N; n += gridDim.x * WARPS) {
    float s = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + (long long)n * D + c);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16(u[k]);
        s += f.x * w[c + 2 * k] + f.y * w[c + 2 * k + 1];
      }
    }
    s = warp_sum(s);
    if (lane == 0) z[n] = s + bias[0];
  }
}

// dw += sum_n dz[n] x[n,:], db += sum_n dz[n]   (D <= 1024)
__global__ void stop_head_bwd_kernel(const float* __restrict__ dz, const bf16* __restrict__ x,
                                     float* __restrict__ dw, float* __restrict__ db, int N, int D) {
  kr::pdl_entry();
  __shared__ float sm[WARPS][1024];
  __shared__ float sb[WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  float ab = 0.f;
  for (int n = blockIdx.x * WARPS + warp; n < N; n += gridDim.x * WARPS) {
    const float g = dz[n];
    ab += g;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane * 8 + 256 * i;
      if (c < D) {
        const uint4 v = *reinterpret_cast<const uint4*>(x + (long long)n * D + c);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(u[k]);
          acc[i * 8 + 2 * k] += g * f.x;
          acc[i * 8 + 2 * k + 1] += g * f.y;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[warp][lane * 8 + 256 * i + k] = acc[i * 8 + k];
  if (lane == 0) sb[warp] = ab;
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < WARPS; ++k) a += sm[k][c];
    atomicAdd(dw + c, a);
  }
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int k = 0; k < WARPS; ++k) a += sb[k];
    atomicAdd(db, a);
  }
}

}  // namespace

extern "C" int kr_stop_head_fwd(const void* x, const float* w, const float* bias, float* z, int N, int D,
                                void* stream) {
  if (N <= 0) return KR_OK;
  if (D % 8) { kr_set_error("kr_stop_head: D % 8 != 0"); return KR_ERR_ARG; }
  const int blocks = min((N + WARPS - 1) / WARPS, kNumSMs * 8);
  kr::launch(stop_head_fwd_kernel, blocks, WARPS * 32, 0, (cudaStream_t)stream, (const bf16*)x, w, bias, z, N, D);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_stop_head_bwd(const float* dz, const void* x, float* dw, float* db, int N, int D,
                                void* stream) {
  if (N <= 0) return KR_OK;
  if ((D % 8) || D > 1024) { kr_set_error("kr_stop_head: D % 8 != 0 or D > 1024"); return KR_ERR_ARG; }
  const int blocks = min((N + WARPS - 1) / WARPS, kNumSMs);
  kr::launch(stop_head_bwd_kernel, blocks, WARPS * 32, 0, (cudaStream_t)stream, dz, (const bf16*)x, dw, db, N, D);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// Fused losses + gradients.  `acc` (10 doubles) is scratch; `losses` receives
// [total, mel, dur, stop, pitch, energy]; gradients are d(total * loss_scale)/d(prediction).
extern "C" int kr_losses_fwd_bwd(const float* mel_pred, const float* mel_tgt, const float* dur_pred,
                                 const long long* dur_tgt, const float* stop_pred, const float* stop_tgt,
                                 const float* pitch_pred, const float* pitch_tgt, const float* energy_pred,
                                 const float* energy_tgt, const long long* mel_len, const long long* ph_len,
                                 int B, int T, int P, int C, int Tp, int Tt, float w_dur, float w_stop,
                                 float w_pitch, float w_energy, float pos_weight, float delta_var,
                                 const float* loss_scale, double* acc, float* losses, void* dmel_bf16,
                                 float* ddur, float* dstop, float* dpitch, float* denergy, void* stream) {
  if (B <= 0) return KR_OK;
  if (pitch_pred != nullptr && (Tp < T || Tt < T)) {
    kr_set_error("kr_losses: frame-level pitch/energy predictions and targets must cover the mel length");
    return KR_ERR_ARG;
  }
  LossParams p;
  p.mel_pred = mel_pred; p.mel_tgt = mel_tgt; p.dur_pred = dur_pred; p.dur_tgt = dur_tgt;
  p.stop_pred = stop_pred; p.stop_tgt = stop_tgt; p.pitch_pred = pitch_pred; p.pitch_tgt = pitch_tgt;
  p.energy_pred = energy_pred; p.energy_tgt = energy_tgt; p.mel_len = mel_len; p.ph_len = ph_len;
  p.B = B; p.T = T; p.P = P; p.C = C; p.Tp = Tp; p.Tt = Tt;
  p.w_dur = w_dur; p.w_stop = w_stop; p.w_pitch = w_pitch; p.w_energy = w_energy;
  p.pos_weight = pos_weight; p.delta_var = delta_var; p.acc = acc; p.losses = losses;
  p.loss_scale = loss_scale; p.dmel = (bf16*)dmel_bf16; p.ddur = ddur; p.dstop = dstop;
  p.dpitch = dpitch; p.denergy = denergy;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(acc, 0, 10 * sizeof(double), st) != cudaSuccess) { kr_set_error("memset failed"); return KR_ERR_CUDA; }
  const long long n = (long long)B * T * C;
  const int blocks = (int)((n + 255) / 256 < kNumSMs * 4 ? (n + 255) / 256 : kNumSMs * 4);
  kr::launch(loss_reduce_kernel, blocks, 256, 0, st, p);
  KR_CHECK_LAUNCH();
  kr::launch(loss_grad_kernel, blocks, 256, 0, st, p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
