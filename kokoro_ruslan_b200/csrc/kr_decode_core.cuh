// kr_decode_core.cuh — bodies of the autoregressive DECODE-step kernels: the reference's
// KokoroGenerator.generate loop (model/generator.py:24-127) with MultiHeadAttentionImproved's KV cache
// (model/transformers.py:237-277, 393-437).  Three small kernels are new; everything else of a decode step (LayerNorm,
// the projections, GLU, the FFN output RMSNorm) is the training path's kernels run on a 128-row padded batch.
//
//   dec_feed_body     x0[b,:] = W_in . frame[b,:] + b_in + PE[t]                        (model.py:519-531, eval mode)
//   dec_attn_body     one (batch, head): RMSNorm(q) (+ RMSNorm/RoPE of the new key, RMSNorm of the new value, both
//                     appended to the cache) and softmax(q K^T / 8) V over the cached keys; cross-attention = the same
//                     without the append, with the memory's key-padding mask
//   dec_finish_body   decoder.norm -> mel frame + stop logit, the generator's stop rules ON THE DEVICE
//                     (generator.py:66-86), the next input frame, t += 1
//
// Cache contents: the reference keeps RAW key projections and re-applies k-norm + RoPE to the whole cache every step with
// positions 0..t; key j therefore always carries position j, and caching the normalised, rotated key once is the same
// function.  The single new QUERY is rotated as position 0 every step (q_offset = 0, transformers.py:276-277) — the
// identity rotation — which is the reference's train / inference mismatch, reproduced here (oracle/inference.py).
//
// DUAL-COMPILED like kr_features_core.cuh: g++ -DKR_HOST_EMU turns a block into one sequential "thread"
// (tests/emu/decode_emu.cpp), so the CPU suite runs the whole generation loop through these bodies against the oracle.
#pragma once

#ifdef KR_HOST_EMU
#include <math.h>
#include <stdint.h>
#include <string.h>
#define KRD_DEV static inline
#ifdef KR_HOST_EMU_SIMT            // one host thread per CUDA thread, real block / warp geometry (tests/emu/emu_simt.h)
#include "emu_simt.h"
#define KRD_TID (emu::tid())
#define KRD_NT (emu::nthreads())
#define KRD_LANE (emu::tid() & 31)
#define KRD_NLANES 32
#define KRD_WARP (emu::tid() >> 5)
#define KRD_NWARPS (emu::nthreads() >> 5)
#define KRD_SYNC() emu::syncthreads()
KRD_DEV float krd_warp_sum(float v) { return emu::warp_sum(v); }
KRD_DEV float krd_warp_max(float v) { return emu::warp_max(v); }
KRD_DEV void krd_warp_sum_vec64(float* v) { emu::warp_sum_vec(v, 64); }
KRD_DEV float krd_block_sum(float v, float* red) { return emu::block_sum(v, red); }
#else                              // one sequential "thread" per block
#define KRD_TID 0
#define KRD_NT 1
#define KRD_LANE 0
#define KRD_NLANES 1
#define KRD_WARP 0
#define KRD_NWARPS 1
#define KRD_SYNC() do { } while (0)
KRD_DEV float krd_warp_sum(float v) { return v; }
KRD_DEV float krd_warp_max(float v) { return v; }
KRD_DEV void krd_warp_sum_vec64(float*) { }
KRD_DEV float krd_block_sum(float v, float*) { return v; }
#endif
typedef uint16_t krd_bf16;
KRD_DEV float krd_b2f(krd_bf16 v) { uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4); return f; }
KRD_DEV krd_bf16 krd_f2b(float f) {                    // round to nearest even, like __float2bfloat16_rn
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (krd_bf16)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (krd_bf16)(u >> 16);
}
KRD_DEV float krd_rsqrt(float x) { return 1.f / sqrtf(x); }
KRD_DEV void krd_load8(const krd_bf16* p, float* v) { for (int i = 0; i < 8; ++i) v[i] = krd_b2f(p[i]); }
template <int N> KRD_DEV void krd_loadn(const krd_bf16* p, float* v) { for (int i = 0; i < N; ++i) v[i] = krd_b2f(p[i]); }
#else
#define KRD_DEV __device__ __forceinline__
#define KRD_TID ((int)threadIdx.x)
#define KRD_NT ((int)blockDim.x)
#define KRD_LANE ((int)(threadIdx.x & 31))
#define KRD_NLANES 32
#define KRD_WARP ((int)(threadIdx.x >> 5))
#define KRD_NWARPS ((int)(blockDim.x >> 5))
#define KRD_SYNC() __syncthreads()
typedef __nv_bfloat16 krd_bf16;
KRD_DEV float krd_b2f(krd_bf16 v) { return __bfloat162float(v); }
KRD_DEV krd_bf16 krd_f2b(float f) { return __float2bfloat16_rn(f); }
KRD_DEV float krd_warp_sum(float v) { return kr::warp_sum(v); }
KRD_DEV float krd_warp_max(float v) { return kr::warp_max(v); }
KRD_DEV void krd_warp_sum_vec64(float* v) {                     // 64 butterfly reductions; every lane gets every sum
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = kr::warp_sum(v[i]);
}
KRD_DEV float krd_block_sum(float v, float* red) { return kr::block_sum(v, red); }
KRD_DEV float krd_rsqrt(float x) { return rsqrtf(x); }
KRD_DEV void krd_load8(const krd_bf16* p, float* v) {          // one 16-byte load of 8 bf16 (p is 16-byte aligned)
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
// N consecutive bf16; the decode attention uses N = 2: one 4-byte load per lane, 128 contiguous bytes per warp and key
template <int N> KRD_DEV void krd_loadn(const krd_bf16* p, float* v) {
  static_assert(N == 2, "a lane owns two consecutive head dimensions");
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  v[0] = __uint_as_float(u << 16);
  v[1] = __uint_as_float(u & 0xffff0000u);
}
#endif

namespace krd {

constexpr int DK = 64;             // head dimension (hidden_dim / n_heads on every configuration of the path)
constexpr int RING = 30;           // generator.py:81 looks at the last 30 frames
constexpr int MAX_WARPS = 8;
constexpr int MAX_B = 16;          // utterances decoded together by one dec_finish block
constexpr float NEG_INF = -3.402823466e38f;

// Device-resident generation state, 64 words.  Written by the host before the loop, advanced by dec_finish_body only.
struct DecState {
  int t;              // index of the frame the current step produces
  int done;           // set by the stop rules or when t reaches hi
  int n_frames;       // frames generated when done was set
  int lo, hi;         // model.py:737-745 bounds: no stop test before lo, hard stop at hi
  int expected;       // expanded memory length: beyond it the stop threshold drops (generator.py:70-73)
  float stop_thr, post_thr;
  float ring[RING];   // mean of each of the last 30 un-clamped output frames
  int reserved[26];
};

KRD_DEV float rms_scale(float sumsq, float eps) { return krd_rsqrt(sumsq * (1.f / (float)DK) + eps); }

// ------------------------------------------------------------------------------------------------------------------
// Decoder input of frame t: mel_projection_in on the previous output frame (zeros at t = 0; `forced` substitutes given
// input frames — the teacher-forced mode the parity tests use) + bias + sinusoidal PE row t.  One block per utterance.
// ------------------------------------------------------------------------------------------------------------------
KRD_DEV void dec_feed_body(const DecState* st, const float* frame, const float* w_in, const float* b_in, const float* pe,
                           int D, int n_mels, float* x_row) {
  const int t = st->t;
  for (int n = KRD_TID; n < D; n += KRD_NT) {
    float acc = b_in[n] + pe[(long long)t * D + n];
    const float* w = w_in + (long long)n * n_mels;
    for (int k = 0; k < n_mels; ++k) acc += w[k] * frame[k];
    x_row[n] = acc;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// One (utterance, head) of a decode-step attention.  q_raw / k_raw / v_raw: this head's 64 bf16 values of the projection
// outputs.  kc / vc: this (utterance, head)'s cache columns, key j at kc + j * ld.  append_at >= 0 (self-attention):
// the new key / value are normalised, the key rotated to position append_at (cos_row / sin_row = that row of the RoPE
// tables) and both stored at row append_at before the scores are taken over n_keys = append_at + 1 rows.
// rotate_q != 0 (self-attention only): the query is rotated to position append_at as well instead of position 0.
// mask (cross-attention): 1 = padded memory frame.  Shared memory: qs[64], wm[8], wl[8], wacc[8 * 64].
// ------------------------------------------------------------------------------------------------------------------
KRD_DEV void dec_attn_body(const krd_bf16* q_raw, const float* gq, const krd_bf16* k_raw, const float* gk,
                           const krd_bf16* v_raw, const float* gv, const float* cos_row, const float* sin_row,
                           krd_bf16* kc, krd_bf16* vc, long long ld, int n_keys, int append_at,
                           const unsigned char* mask, float scale, float eps, int rotate_q, float* qs, float* wm,
                           float* wl, float* wacc, krd_bf16* o_out) {
  const int lane = KRD_LANE, warp = KRD_WARP, nw = KRD_NWARPS;
  constexpr int HALF = DK / 2;
  // phase A: the three per-head RMSNorms, each by one warp (the same warp when the block has fewer)
  if (warp == 0) {
    float ss = 0.f;
    for (int d = lane; d < DK; d += KRD_NLANES) { const float v = krd_b2f(q_raw[d]); ss += v * v; }
    const float r = rms_scale(krd_warp_sum(ss), eps);
    if (rotate_q && append_at >= 0) {
      // opt-in FIX of the reference's quirk: rotate the new query to its true position (what training does)
      for (int d = lane; d < HALF; d += KRD_NLANES) {
        const float a = krd_b2f(q_raw[d]) * r * gq[d], b = krd_b2f(q_raw[d + HALF]) * r * gq[d + HALF];
        const float c = cos_row[d], s = sin_row[d];
        qs[d] = a * c - b * s;
        qs[d + HALF] = b * c + a * s;
      }
    } else {
      for (int d = lane; d < DK; d += KRD_NLANES) qs[d] = krd_b2f(q_raw[d]) * r * gq[d];   // RoPE at position 0 = identity
    }
  }
  if (append_at >= 0 && warp == 1 % nw) {
    float ss = 0.f;
    for (int d = lane; d < DK; d += KRD_NLANES) { const float v = krd_b2f(k_raw[d]); ss += v * v; }
    const float r = rms_scale(krd_warp_sum(ss), eps);
    for (int d = lane; d < HALF; d += KRD_NLANES) {          // rotate-half pairs (d, d + 32), positional_encoding.py:196-209
      const float a = krd_b2f(k_raw[d]) * r * gk[d], b = krd_b2f(k_raw[d + HALF]) * r * gk[d + HALF];
      const float c = cos_row[d], s = sin_row[d];
      kc[(long long)append_at * ld + d] = krd_f2b(a * c - b * s);
      kc[(long long)append_at * ld + d + HALF] = krd_f2b(b * c + a * s);
    }
  }
  if (append_at >= 0 && warp == 2 % nw) {
    float ss = 0.f;
    for (int d = lane; d < DK; d += KRD_NLANES) { const float v = krd_b2f(v_raw[d]); ss += v * v; }
    const float r = rms_scale(krd_warp_sum(ss), eps);
    for (int d = lane; d < DK; d += KRD_NLANES) vc[(long long)append_at * ld + d] = krd_f2b(krd_b2f(v_raw[d]) * r * gv[d]);
  }
  KRD_SYNC();
  // phase B: ONE KEY PER LANE.  Lane l of warp w takes keys w * 32 + l, + 32 * nw, ...: it reads its key / value rows
  // with 16-byte loads (8 per row), the query from shared memory (broadcast), and keeps a private online softmax
  // (m, l, acc[64]) — no cross-lane traffic and no dependent shuffle chain inside the loop, which is what bounds a
  // one-key-per-warp formulation at long contexts.  The 32 partial softmaxes of a warp are merged once at the end.
  float m = NEG_INF, l = 0.f, acc[DK];
#ifndef KR_HOST_EMU
#pragma unroll
#endif
  for (int d = 0; d < DK; ++d) acc[d] = 0.f;
  for (int j = warp * KRD_NLANES + lane; j < n_keys; j += nw * KRD_NLANES) {
    if (mask != nullptr && mask[j]) continue;
    const krd_bf16* krow = kc + (long long)j * ld;
    const krd_bf16* vrow = vc + (long long)j * ld;
    float dot = 0.f;
#ifndef KR_HOST_EMU
#pragma unroll
#endif
    for (int c = 0; c < DK / 8; ++c) {
      float kv[8];
      krd_load8(krow + c * 8, kv);
#ifndef KR_HOST_EMU
#pragma unroll
#endif
      for (int i = 0; i < 8; ++i) dot += qs[c * 8 + i] * kv[i];
    }
    const float s = dot * scale;
    const float m_new = fmaxf(m, s);
    const float corr = expf(m - m_new), p = expf(s - m_new);
    l = l * corr + p;
#ifndef KR_HOST_EMU
#pragma unroll
#endif
    for (int c = 0; c < DK / 8; ++c) {
      float vv[8];
      krd_load8(vrow + c * 8, vv);
#ifndef KR_HOST_EMU
#pragma unroll
#endif
      for (int i = 0; i < 8; ++i) acc[c * 8 + i] = acc[c * 8 + i] * corr + p * vv[i];
    }
    m = m_new;
  }
  // merge the lanes of this warp (all 32 lanes take part: the loop above has no early exit, only skipped iterations)
  const float m_w = krd_warp_max(m);
  const float f = l > 0.f ? expf(m - m_w) : 0.f;
  const float l_w = krd_warp_sum(l * f);
#ifndef KR_HOST_EMU
#pragma unroll
#endif
  for (int d = 0; d < DK; ++d) acc[d] *= f;
  krd_warp_sum_vec64(acc);
#ifndef KR_HOST_EMU
#pragma unroll
#endif
  for (int d = 0; d < DK; ++d)
    if ((d % KRD_NLANES) == lane) wacc[warp * DK + d] = acc[d];
  if (lane == 0) { wm[warp] = m_w; wl[warp] = l_w; }
  KRD_SYNC();
  // phase C: merge the warps' partial softmaxes
  for (int d = KRD_TID; d < DK; d += KRD_NT) {
    float M = NEG_INF;
    for (int w = 0; w < nw; ++w) M = fmaxf(M, wm[w]);
    float L = 0.f, o = 0.f;
    for (int w = 0; w < nw; ++w) {
      if (wl[w] == 0.f) continue;                          // warp saw no (unmasked) key
      const float f = expf(wm[w] - M);
      L += wl[w] * f;
      o += wacc[w * DK + d] * f;
    }
    o_out[d] = krd_f2b(L > 0.f ? o / L : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// End of a decode step, ONE block for the whole batch: decoder.norm (LayerNorm, eps 1e-5), mel_projection_out and the
// stop head on it (model.py:547-563), then generator.py:58-103: append the frame, stop when
//   t >= lo and mean_b sigmoid(stop) > (t < expected ? stop_thr : min(stop_thr, post_thr)),   or
//   t >= lo and 30 frames exist and the mean of the last 30 frames < -9.5,                     or   t + 1 == hi.
// The un-clamped frame is the next step's input; the stored output is clamped to [-11.5, 2] (model.py:775-777).
// Shared memory: stats[2 * MAX_B], vals[MAX_B * (n_mels + 1)], red[32].
// ------------------------------------------------------------------------------------------------------------------
KRD_DEV void dec_finish_body(DecState* st, const float* y, const float* ln_g, const float* ln_b, const float* w_out,
                             const float* b_out, const float* w_stop, const float* b_stop, int B, int D, int n_mels,
                             int t_cap, float* stats, float* vals, float* red, float* mel_out, float* next_frame,
                             float* probs) {
  if (st->done) return;                                   // uniform: every thread reads the same word
  const int t = st->t, lane = KRD_LANE, warp = KRD_WARP, nw = KRD_NWARPS;
  for (int b = warp; b < B; b += nw) {                    // LayerNorm statistics, one warp per row
    const float* row = y + (long long)b * D;
    float s = 0.f;
    for (int d = lane; d < D; d += KRD_NLANES) s += row[d];
    const float mean = krd_warp_sum(s) / (float)D;
    float v = 0.f;
    for (int d = lane; d < D; d += KRD_NLANES) { const float c = row[d] - mean; v += c * c; }
    const float rstd = krd_rsqrt(krd_warp_sum(v) / (float)D + 1e-5f);
    if (lane == 0) { stats[2 * b] = mean; stats[2 * b + 1] = rstd; }
  }
  KRD_SYNC();
  const int n_out = n_mels + 1;                           // 80 mel bins + the stop logit
  for (int o = warp; o < B * n_out; o += nw) {
    const int b = o / n_out, n = o % n_out;
    const float* row = y + (long long)b * D;
    const float* w = n < n_mels ? w_out + (long long)n * D : w_stop;
    const float mean = stats[2 * b], rstd = stats[2 * b + 1];
    float acc = 0.f;
    for (int d = lane; d < D; d += KRD_NLANES) acc += ((row[d] - mean) * rstd * ln_g[d] + ln_b[d]) * w[d];
    acc = krd_warp_sum(acc);
    if (lane == 0) vals[o] = acc + (n < n_mels ? b_out[n] : b_stop[0]);
  }
  KRD_SYNC();
  float fsum = 0.f, psum = 0.f;
  for (int i = KRD_TID; i < B * n_mels; i += KRD_NT) {
    const int b = i / n_mels, n = i % n_mels;
    const float v = vals[b * n_out + n];
    fsum += v;
    next_frame[i] = v;
    mel_out[((long long)b * t_cap + t) * n_mels + n] = fminf(fmaxf(v, -11.5f), 2.f);
  }
  for (int b = KRD_TID; b < B; b += KRD_NT) psum += 1.f / (1.f + expf(-vals[b * n_out + n_mels]));
  fsum = krd_block_sum(fsum, red);
  psum = krd_block_sum(psum, red);
  KRD_SYNC();
  if (KRD_TID == 0) {
    const float p = psum / (float)B;
    probs[t] = p;
    st->ring[t % RING] = fsum / (float)(B * n_mels);
    int stop = 0;
    if (t >= st->lo) {
      const float thr = t < st->expected ? st->stop_thr : fminf(st->stop_thr, st->post_thr);
      if (p > thr) stop = 1;
      if (!stop && t + 1 >= RING) {
        float r = 0.f;
        for (int i = 0; i < RING; ++i) r += st->ring[i];
        if (r / (float)RING < -9.5f) stop = 1;
      }
    }
    if (t + 1 >= st->hi || t + 1 >= t_cap) stop = 1;
    st->t = t + 1;
    if (stop) { st->n_frames = t + 1; st->done = 1; }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Skinny projection of a decode step: out[b, n] = sum_k x[b, k] * W[n, k] (+ bias[n]) (+ resid[b, n]) for B <= 8 rows.
// The weight matrix is streamed ONCE by the whole grid (a warp owns output feature n and reads its K bf16 weights with
// 16-byte loads), the B activation rows sit in shared memory as bf16; fp32 accumulation.  This is the weight-bandwidth
// shape of a decode step — the padded 128-row tensor-core GEMM it replaces touches the same bytes from 2 - 12 CTAs only.
// xs: shared memory, B * K bf16 (filled here).  K % 8 == 0.
// ------------------------------------------------------------------------------------------------------------------
constexpr int GEMV_MAX_B = 8;

KRD_DEV float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// Fusions (each removes one launch per use from the 13-kernel decoder layer):
//   x_f32 != nullptr  LayerNorm prologue: the activation rows are LN(x_f32[b]; ln_g, ln_b, eps 1e-5) rounded to bf16
//                     (what kr_layernorm_fwd would have written), computed redundantly by every block — B <= 8 rows;
//   glu != 0          GLU epilogue (model/transformers.py:105-108): W is linear1 [2 * N, K], a warp computes the gate
//                     feature n and the linear feature n + N together and writes gelu_erf(gate + b) * (lin + b) as bf16.
KRD_DEV void dec_gemv_body(const krd_bf16* x, const float* x_f32, long long ld_x, const float* ln_g, const float* ln_b,
                           const krd_bf16* w, const float* bias, const float* resid, long long ld_r, void* out,
                           long long ld_o, int out_f32, int glu, int B, int N, int K, int n_first, int n_step,
                           krd_bf16* xs) {
  const int lane = KRD_LANE;
  if (x_f32 != nullptr) {
    for (int b = KRD_WARP; b < B; b += KRD_NWARPS) {              // one warp per row
      const float* row = x_f32 + (long long)b * ld_x;
      float s = 0.f;
      for (int k = lane; k < K; k += KRD_NLANES) s += row[k];
      const float mean = krd_warp_sum(s) / (float)K;
      float v = 0.f;
      for (int k = lane; k < K; k += KRD_NLANES) { const float c = row[k] - mean; v += c * c; }
      const float rstd = krd_rsqrt(krd_warp_sum(v) / (float)K + 1e-5f);
      for (int k = lane; k < K; k += KRD_NLANES) xs[b * K + k] = krd_f2b((row[k] - mean) * rstd * ln_g[k] + ln_b[k]);
    }
  } else {
    for (int i = KRD_TID; i < B * K; i += KRD_NT) xs[i] = x[(long long)(i / K) * ld_x + (i % K)];
  }
  KRD_SYNC();
  const int chunks = K / 8;
  for (int n = n_first + KRD_WARP; n < N; n += n_step) {
    float acc[GEMV_MAX_B], acc2[GEMV_MAX_B];
    for (int b = 0; b < GEMV_MAX_B; ++b) acc[b] = acc2[b] = 0.f;
    const krd_bf16* wrow = w + (long long)n * K;
    const krd_bf16* wrow2 = w + (long long)(n + N) * K;           // the "lin" half of linear1 (GLU only)
    for (int c = lane; c < chunks; c += KRD_NLANES) {
      float wv[8], wv2[8], xv[8];
      krd_load8(wrow + c * 8, wv);
      if (glu) krd_load8(wrow2 + c * 8, wv2);
      for (int b = 0; b < B; ++b) {
        krd_load8(xs + b * K + c * 8, xv);
        float s = 0.f, s2 = 0.f;
        for (int i = 0; i < 8; ++i) s += wv[i] * xv[i];
        if (glu) for (int i = 0; i < 8; ++i) s2 += wv2[i] * xv[i];
        acc[b] += s;
        acc2[b] += s2;
      }
    }
    for (int b = 0; b < B; ++b) {
      const float v = krd_warp_sum(acc[b]);
      const float v2 = glu ? krd_warp_sum(acc2[b]) : 0.f;
      if (lane == 0) {
        float o;
        if (glu) {
          o = gelu_erf(v + (bias != nullptr ? bias[n] : 0.f)) * (v2 + (bias != nullptr ? bias[n + N] : 0.f));
        } else {
          o = v + (bias != nullptr ? bias[n] : 0.f) + (resid != nullptr ? resid[(long long)b * ld_r + n] : 0.f);
        }
        if (out_f32) reinterpret_cast<float*>(out)[(long long)b * ld_o + n] = o;
        else reinterpret_cast<krd_bf16*>(out)[(long long)b * ld_o + n] = krd_f2b(o);
      }
    }
  }
}

}  // namespace krd
