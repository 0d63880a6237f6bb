// kr_variance.cu — variance adaptor kernels (HBM-bound; the k=3 convolutions themselves run on the
// tcgen05 GEMM over an overlapping-row view of a zero-padded activation buffer).
//   * LengthRegulator index: per-row inclusive scan of durations + upper-bound search, int32
//     indices bit-exact with reference utils/lengths.py:16-96
//   * expansion gather + pitch/energy bucketize + embedding add + frame mask
//     (reference model/variance_predictor.py:196-218,345-437)
//   * GroupNorm(1 group) + ReLU over independent 512-frame chunks, fwd/bwd
//     (model/variance_predictor.py:70-115; statistics per (sample, chunk) over C x L_chunk, padded
//     frames included)
//   * Linear(F->1) head + mask, fwd/bwd
//
// "padded layout": the rows of one predictor input are laid out chunk by chunk with one zero row
// before and after every chunk (the conv's zero padding), plus one guard row at each end of the
// buffer.  row_group[r] = (sample, chunk) id of padded row r or -1 for a zero row.
#include "kr_common.cuh"
#include "kokoro_b200.h"

namespace {
using namespace kr;
constexpr int WARPS = 8;

// ---------------------------------------------------------------------------------------------
// length regulator
// ---------------------------------------------------------------------------------------------
// pad_mask == nullptr: LengthRegulator semantics (durations clamped at 0, utils/lengths.py:38-41).
// pad_mask != nullptr: the `length_regulate` fallback (utils/lengths.py:108-153): padded tokens are skipped and
// every remaining duration is clamped to >= 1.
__global__ void lr_index_kernel(const long long* __restrict__ dur, const unsigned char* __restrict__ pad_mask,
                                int* __restrict__ idx, int* __restrict__ lengths, int P, int Tp) {
  kr::pdl_entry();
  extern __shared__ int cs[];  // P ints (double-buffered scan: 2*P)
  const int b = blockIdx.x;
  int* a = cs;
  int* t = cs + P;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const long long d = dur[(long long)b * P + i];
    if (pad_mask != nullptr) a[i] = pad_mask[(long long)b * P + i] ? 0 : (d > 1 ? (int)d : 1);
    else a[i] = d > 0 ? (int)d : 0;
  }
  __syncthreads();
  for (int off = 1; off < P; off <<= 1) {
    for (int i = threadIdx.x; i < P; i += blockDim.x) t[i] = a[i] + (i >= off ? a[i - off] : 0);
    __syncthreads();
    int* s = a; a = t; t = s;
  }
  const int L = a[P - 1];
  if (threadIdx.x == 0) lengths[b] = L;
  for (int f = threadIdx.x; f < Tp; f += blockDim.x) {
    int r = -1;
    if (f < L) {  // first j with cs[j] > f
      int lo = 0, hi = P - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] > f) hi = mid; else lo = mid + 1;
      }
      r = lo;
    }
    idx[(long long)b * Tp + f] = r;
  }
}

// out[b, f, :] = x[b, idx[b, f], :] (0 where idx < 0); mask[b, f] = idx < 0   (fp32, exact copy)
__global__ void expand_rows_kernel(const float* __restrict__ x, const int* __restrict__ idx, float* __restrict__ out,
                                   unsigned char* __restrict__ mask, int B, int P, int Tp, int D) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < (long long)B * Tp;
       r += (long long)gridDim.x * WARPS) {
    const int b = (int)(r / Tp);
    const int j = idx[r];
    if (lane == 0 && mask != nullptr) mask[r] = j < 0 ? 1 : 0;
    const float4* src = j >= 0 ? reinterpret_cast<const float4*>(x + ((long long)b * P + j) * D) : nullptr;
    float4* dst = reinterpret_cast<float4*>(out + r * D);
    for (int c = lane; c < D / 4; c += 32) dst[c] = src != nullptr ? src[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// dx[b, j, :] = sum over the frames f with idx[b, f] == j of dout[b, f, :]: frames of a token are one contiguous
// run, found by binary search in the (sorted) index row -> deterministic segment sums, no atomics.
__global__ void expand_rows_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ idx,
                                       const int* __restrict__ lengths, float* __restrict__ dx, int B, int P, int Tp,
                                       int D) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  for (long long t = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); t < (long long)B * P;
       t += (long long)gridDim.x * WARPS) {
    const int b = (int)(t / P), j = (int)(t % P);
    const int* row = idx + (long long)b * Tp;
    const int L = min(lengths[b], Tp);
    int lo = 0, hi = L;                     // first f with row[f] >= j
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid] >= j) hi = mid; else lo = mid + 1; }
    const int f0 = lo;
    hi = L;                                 // first f with row[f] > j
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid] > j) hi = mid; else lo = mid + 1; }
    const int f1 = lo;
    for (int c = lane; c < D / 4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int f = f0; f < f1; ++f) {
        const float4 v = *reinterpret_cast<const float4*>(dout + ((long long)b * Tp + f) * D + 4 * c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(dx + t * D + 4 * c) = acc;
    }
  }
}

__global__ void range_flag_kernel(const float* __restrict__ x, long long n, int* flag) {
  kr::pdl_entry();
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    bad |= (v > 1.f) || (v < 0.f);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__device__ __forceinline__ int bucketize(const float* bins, int nb, float v) {
  int lo = 0, hi = nb;  // count of bins[i] < v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bins[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct AdaptParams {
  const float* enc;      // [B,P,D]
  const int* idx;        // [B,Tp]
  const int* lengths;    // [B]
  const float* pitch;    // [B,Tt]
  const float* energy;   // [B,Tt]
  const int* flags;      // [2] range flags (1 -> normalise+clamp)
  const float* pbins; const float* ebins;   // [nb]
  const float* pemb; const float* eemb;     // [nb+1, D]
  const int* row_of_tok; // [B*Tp] padded row of (b,f)
  bf16* xpad;            // padded predictor input (pointer past the guard row)
  bf16* mem;             // [B,T,D]
  int* p_idx; int* e_idx;  // [B,T]
  unsigned char* fmask_t;  // [B,T]  cross-attention key mask
  unsigned char* fmask_p;  // [B,Tp] predictor mask
  int B, P, D, Tp, T, Tt, nb;
};

__global__ void expand_adapt_kernel(const AdaptParams p) {
  kr::pdl_entry();
  __shared__ float sb[2][256];
  for (int i = threadIdx.x; i < p.nb; i += blockDim.x) { sb[0][i] = p.pbins[i]; sb[1][i] = p.ebins[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int Tmax = max(p.Tp, p.T);
  const long long rows = (long long)p.B * Tmax;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * WARPS) {
    const int b = (int)(r / Tmax), f = (int)(r % Tmax);
    const int L = p.lengths[b];
    const bool in_p = f < p.Tp, in_t = f < p.T;
    const bool valid = in_p && f < L;
    const int src = valid ? p.idx[(long long)b * p.Tp + f] : -1;
    int pi = 0, ei = 0;
    if (valid && in_t) {
      float pv = f < p.Tt ? p.pitch[(long long)b * p.Tt + f] : 0.f;
      float ev = f < p.Tt ? p.energy[(long long)b * p.Tt + f] : 0.f;
      if (p.flags[0]) pv = fminf(fmaxf(pv / (1.f + 1e-8f), 0.f), 1.f);
      if (p.flags[1]) ev = fminf(fmaxf(ev / (1.f + 1e-8f), 0.f), 1.f);
      pi = bucketize(sb[0], p.nb, pv);
      ei = bucketize(sb[1], p.nb, ev);
    }
    if (lane == 0) {
      if (in_t) {
        p.fmask_t[(long long)b * p.T + f] = valid ? 0 : 1;
        p.p_idx[(long long)b * p.T + f] = valid ? pi : -1;
        p.e_idx[(long long)b * p.T + f] = valid ? ei : -1;
      }
      if (in_p) p.fmask_p[(long long)b * p.Tp + f] = valid ? 0 : 1;
    }
    const float* er = valid ? p.enc + ((long long)b * p.P + src) * p.D : nullptr;
    bf16* xr = in_p ? p.xpad + (long long)p.row_of_tok[(long long)b * p.Tp + f] * p.D : nullptr;
    bf16* mr = in_t ? p.mem + ((long long)b * p.T + f) * p.D : nullptr;
    for (int c = lane * 4; c < p.D; c += 128) {
      float4 v = make_float4(0, 0, 0, 0);
      if (valid) v = *reinterpret_cast<const float4*>(er + c);
      if (xr != nullptr) {
        uint2 u; u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(xr + c) = u;
      }
      if (mr != nullptr) {
        if (valid) {
          const float4 a = *reinterpret_cast<const float4*>(p.pemb + (long long)pi * p.D + c);
          const float4 e = *reinterpret_cast<const float4*>(p.eemb + (long long)ei * p.D + c);
          v.x += a.x + e.x; v.y += a.y + e.y; v.z += a.z + e.z; v.w += a.w + e.w;
        }
        uint2 u; u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(mr + c) = u;
      }
    }
  }
}

// d pitch_emb[p_idx] += dmem ; d energy_emb[e_idx] += dmem   (only source of gradient through the
// detached expansion, SURVEY.md §8 backward sub-graph (i))
__global__ void adapt_bwd_kernel(const float* __restrict__ dmem, const int* __restrict__ p_idx,
                                 const int* __restrict__ e_idx, float* __restrict__ dpemb,
                                 float* __restrict__ deemb, long long rows, int D) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * WARPS) {
    const int pi = p_idx[r], ei = e_idx[r];
    if (pi < 0) continue;
    for (int c = lane; c < D; c += 32) {
      const float g = dmem[r * D + c];
      atomicAdd(dpemb + (long long)pi * D + c, g);
      atomicAdd(deemb + (long long)ei * D + c, g);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm(1) + ReLU on the padded layout
// ---------------------------------------------------------------------------------------------
// Group sums are accumulated per WARP over a contiguous run of rows and flushed (two double atomics) only when the group
// changes or the run ends: a group is one (utterance, 512-frame chunk), i.e. hundreds of consecutive rows, and one atomic
// pair per ROW onto the ~16 group addresses serialised in the L2 (12 us for 6400 rows, 26 us in the backward pass).
struct GroupRun {
  int g = -1;
  double s = 0.0, q = 0.0;
  __device__ __forceinline__ void add(int gr, float a, float b, double* out, int lane) {
    if (gr != g) { flush(out, lane); g = gr; }
    s += (double)a; q += (double)b;
  }
  __device__ __forceinline__ void flush(double* out, int lane) {
    if (g >= 0 && lane == 0) { atomicAdd(out + 2 * g, s); atomicAdd(out + 2 * g + 1, q); }
    s = 0.0; q = 0.0;
  }
};
// contiguous rows [r0, r1) of warp `gw` out of `nw` warps
__device__ __forceinline__ void warp_row_range(int R, int gw, int nw, int& r0, int& r1) {
  const int per = (R + nw - 1) / nw;
  r0 = min(R, gw * per);
  r1 = min(R, r0 + per);
}

__global__ void gn_stats_kernel(const float* __restrict__ x, const int* __restrict__ row_group,
                                double* __restrict__ stats, int R, int C) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  int r0, r1;
  warp_row_range(R, blockIdx.x * WARPS + (threadIdx.x >> 5), gridDim.x * WARPS, r0, r1);
  GroupRun run;
  for (int r = r0; r < r1; ++r) {
    const int g = row_group[r];
    if (g < 0) continue;
    float s = 0.f, q = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = x[(long long)r * C + c]; s += v; q += v * v; }
    s = warp_sum(s); q = warp_sum(q);
    run.add(g, s, q, stats, lane);
  }
  run.flush(stats, lane);
}

__device__ __forceinline__ void gn_mean_rstd(const double* stats, const int* group_rows, int g, int C,
                                             float& mean, float& rstd) {
  const double n = (double)group_rows[g] * C;
  const double m = stats[2 * g] / n;
  double var = stats[2 * g + 1] / n - m * m;
  if (var < 0) var = 0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + 1e-5));
}

__global__ void gn_apply_relu_kernel(const float* __restrict__ x, const int* __restrict__ row_group,
                                     const double* __restrict__ stats, const int* __restrict__ group_rows,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     bf16* __restrict__ out, int R, int C, const DropSpec drop) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  DropCtx dc{};
  if (drop.state != nullptr) dc = drop_ctx(drop);
  for (int r = blockIdx.x * WARPS + (threadIdx.x >> 5); r < R; r += gridDim.x * WARPS) {
    const int g = row_group[r];
    float mean = 0.f, rstd = 0.f;
    if (g >= 0) gn_mean_rstd(stats, group_rows, g, C, mean, rstd);
    for (int c = lane; c < C; c += 32) {
      float y = 0.f;
      if (g >= 0) y = fmaxf((x[(long long)r * C + c] - mean) * rstd * gamma[c] + beta[c], 0.f);
      if (drop.state != nullptr) y *= drop_one(dc, (long long)r * C + c);   // variance_predictor.py:106
      out[(long long)r * C + c] = __float2bfloat16(y);
    }
  }
}

// pass 1 of the backward: per-group sums of g*dy' and g*dy'*xhat, per-channel dgamma/dbeta
__global__ void gn_bwd_stats_kernel(const bf16* __restrict__ dy, const float* __restrict__ x,
                                    const int* __restrict__ row_group, const double* __restrict__ stats,
                                    const int* __restrict__ group_rows, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, double* __restrict__ gsum,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int R, int C,
                                    const DropSpec drop) {
  kr::pdl_entry();
  __shared__ float sm[2][WARPS][256];
  DropCtx dc{};
  if (drop.state != nullptr) dc = drop_ctx(drop);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  int r0, r1;
  warp_row_range(R, blockIdx.x * WARPS + warp, gridDim.x * WARPS, r0, r1);
  GroupRun run;
  int g_cached = -1;
  float mean = 0.f, rstd = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int g = row_group[r];
    if (g < 0) continue;
    if (g != g_cached) { gn_mean_rstd(stats, group_rows, g, C, mean, rstd); g_cached = g; }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float xh = (x[(long long)r * C + c] - mean) * rstd;
        const float y = xh * gamma[c] + beta[c];
        float d = y > 0.f ? __bfloat162float(dy[(long long)r * C + c]) : 0.f;
        if (drop.state != nullptr) d *= drop_one(dc, (long long)r * C + c);
        ag[i] += d * xh; ab[i] += d;
        s1 += d * gamma[c]; s2 += d * gamma[c] * xh;
      }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    run.add(g, s1, s2, gsum, lane);
  }
  run.flush(gsum, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) { sm[0][warp][lane + 32 * i] = ag[i]; sm[1][warp][lane + 32 * i] = ab[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { a += sm[0][w][c]; b += sm[1][w][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

__global__ void gn_bwd_apply_kernel(const bf16* __restrict__ dy, const float* __restrict__ x,
                                    const int* __restrict__ row_group, const double* __restrict__ stats,
                                    const int* __restrict__ group_rows, const double* __restrict__ gsum,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    bf16* __restrict__ dx, int R, int C, const DropSpec drop) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  DropCtx dc{};
  if (drop.state != nullptr) dc = drop_ctx(drop);
  for (int r = blockIdx.x * WARPS + (threadIdx.x >> 5); r < R; r += gridDim.x * WARPS) {
    const int g = row_group[r];
    float mean = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
    if (g >= 0) {
      gn_mean_rstd(stats, group_rows, g, C, mean, rstd);
      const double n = (double)group_rows[g] * C;
      m1 = (float)(gsum[2 * g] / n);
      m2 = (float)(gsum[2 * g + 1] / n);
    }
    for (int c = lane; c < C; c += 32) {
      float o = 0.f;
      if (g >= 0) {
        const float xh = (x[(long long)r * C + c] - mean) * rstd;
        const float y = xh * gamma[c] + beta[c];
        float d = y > 0.f ? __bfloat162float(dy[(long long)r * C + c]) : 0.f;
        if (drop.state != nullptr) d *= drop_one(dc, (long long)r * C + c);
        o = rstd * (d * gamma[c] - m1 - xh * m2);
      }
      dx[(long long)r * C + c] = __float2bfloat16(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Linear(F -> 1) head
// ---------------------------------------------------------------------------------------------
__global__ void vp_head_fwd_kernel(const bf16* __restrict__ h, const int* __restrict__ row_of_tok,
                                   const float* __restrict__ w, const float* __restrict__ bias,
                                   const unsigned char* __restrict__ mask, float* __restrict__ out,
                                   int n_tok, int L, int F, int chunk) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  for (int t = blockIdx.x * WARPS + (threadIdx.x >> 5); t < n_tok; t += gridDim.x * WARPS) {
    const int pos = t % L;
    const bool dead = (mask != nullptr && mask[t]) || ((L % chunk) == 1 && pos == L - 1);
    float s = 0.f;
    if (!dead) {
      const bf16* hr = h + (long long)row_of_tok[t] * F;
      for (int c = lane; c < F; c += 32) s += __bfloat162float(hr[c]) * w[c];
    }
    s = warp_sum(s);
    if (lane == 0) out[t] = dead ? 0.f : s + bias[0];
  }
}

__global__ void vp_head_bwd_kernel(const float* __restrict__ dout, const bf16* __restrict__ h,
                                   const int* __restrict__ tok_of_row, const float* __restrict__ w,
                                   const unsigned char* __restrict__ mask, bf16* __restrict__ dh,
                                   float* __restrict__ dw, float* __restrict__ dbias, int R, int L, int F,
                                   int chunk) {
  kr::pdl_entry();
  __shared__ float sm[WARPS][256];
  __shared__ float sb[WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) aw[i] = 0.f;
  float abias = 0.f;
  for (int r = blockIdx.x * WARPS + warp; r < R; r += gridDim.x * WARPS) {
    const int t = tok_of_row[r];
    float g = 0.f;
    if (t >= 0) {
      const int pos = t % L;
      const bool dead = (mask != nullptr && mask[t]) || ((L % chunk) == 1 && pos == L - 1);
      g = dead ? 0.f : dout[t];
    }
    abias += g;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < F) {
        if (g != 0.f) aw[i] += g * __bfloat162float(h[(long long)r * F + c]);
        dh[(long long)r * F + c] = __float2bfloat16(g * w[c]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[warp][lane + 32 * i] = aw[i];
  if (lane == 0) sb[warp] = abias;
  __syncthreads();
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < WARPS; ++k) a += sm[k][c];
    atomicAdd(dw + c, a);
  }
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int k = 0; k < WARPS; ++k) a += sb[k];
    atomicAdd(dbias, a);
  }
}

// Wd[c, j*Co + o] = W2[o, (2-j)*Ci + c]  (W2 = master conv weight in [Co, 3, Ci] layout)
__global__ void conv_dgrad_shadow_kernel(const float* __restrict__ w2, bf16* __restrict__ wd, int Co, int Ci) {
  kr::pdl_entry();
  const long long total = (long long)Ci * 3 * Co;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % Co);
    const int j = (int)((i / Co) % 3);
    const int c = (int)(i / (3LL * Co));
    wd[i] = __float2bfloat16(w2[((long long)o * 3 + (2 - j)) * Ci + c]);
  }
}

// the same for up to 8 convs of one shape in ONE launch (blockIdx.y = conv): the optimizer step refreshes the four dgrad
// shadows of the variance predictors at its very end, where four dependent ~10 us launches were 40 us of the step
struct ShadowBatch { const float* w2[8]; bf16* wd[8]; };
__global__ void conv_dgrad_shadow_multi_kernel(const ShadowBatch b, int Co, int Ci) {
  kr::pdl_entry();
  const float* __restrict__ w2 = b.w2[blockIdx.y];
  bf16* __restrict__ wd = b.wd[blockIdx.y];
  const long long total = (long long)Ci * 3 * Co;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % Co);
    const int j = (int)((i / Co) % 3);
    const int c = (int)(i / (3LL * Co));
    wd[i] = __float2bfloat16(w2[((long long)o * 3 + (2 - j)) * Ci + c]);
  }
}

inline int warp_blocks(long long rows, int cap_mult = 8) {
  long long b = (rows + WARPS - 1) / WARPS;
  const long long cap = (long long)kNumSMs * cap_mult;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int kr_lr_index(const long long* dur, int* idx, int* lengths, int B, int P, int Tp, void* stream) {
  if (B <= 0) return KR_OK;
  if (P > 4096) { kr_set_error("kr_lr_index: P > 4096 unsupported"); return KR_ERR_UNSUPPORTED; }
  kr::launch(lr_index_kernel, B, 256, 2 * P * sizeof(int), (cudaStream_t)stream, dur, (const unsigned char*)nullptr, idx,
             lengths, P, Tp);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_lr_index_masked(const long long* dur, const unsigned char* pad_mask, int* idx, int* lengths, int B,
                                  int P, int Tp, void* stream) {
  if (B <= 0) return KR_OK;
  if (P > 4096 || pad_mask == nullptr) { kr_set_error("kr_lr_index_masked: P <= 4096 and a padding mask are required"); return KR_ERR_ARG; }
  kr::launch(lr_index_kernel, B, 256, 2 * P * sizeof(int), (cudaStream_t)stream, dur, pad_mask, idx, lengths, P, Tp);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_expand_rows_fwd(const float* x, const int* idx, float* out, unsigned char* frame_mask, int B, int P,
                                  int Tp, int D, void* stream) {
  if (B <= 0 || Tp <= 0) return KR_OK;
  if (D % 4) { kr_set_error("kr_expand_rows: D % 4 == 0 required"); return KR_ERR_ARG; }
  long long nb = ((long long)B * Tp + WARPS - 1) / WARPS;
  if (nb > kNumSMs * 16) nb = kNumSMs * 16;
  kr::launch(expand_rows_kernel, (int)nb, WARPS * 32, 0, (cudaStream_t)stream, x, idx, out, frame_mask, B, P, Tp, D);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_expand_rows_bwd(const float* dout, const int* idx, const int* lengths, float* dx, int B, int P, int Tp,
                                  int D, void* stream) {
  if (B <= 0 || P <= 0) return KR_OK;
  if (D % 4) { kr_set_error("kr_expand_rows: D % 4 == 0 required"); return KR_ERR_ARG; }
  long long nb = ((long long)B * P + WARPS - 1) / WARPS;
  if (nb > kNumSMs * 16) nb = kNumSMs * 16;
  kr::launch(expand_rows_bwd_kernel, (int)nb, WARPS * 32, 0, (cudaStream_t)stream, dout, idx, lengths, dx, B, P, Tp, D);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_range_flag(const float* x, long long n, int* flag, void* stream) {
  if (n <= 0) return KR_OK;
  long long b = (n + 255) / 256;
  kr::launch(range_flag_kernel, (int)(b < 592 ? b : 592), 256, 0, (cudaStream_t)stream, x, n, flag);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_expand_adapt(const float* enc, const int* idx, const int* lengths, const float* pitch,
                               const float* energy, const int* flags, const float* pbins,
                               const float* ebins, const float* pemb, const float* eemb,
                               const int* row_of_tok, void* xpad, void* mem, int* p_idx, int* e_idx,
                               unsigned char* fmask_t, unsigned char* fmask_p, int B, int P, int D, int Tp,
                               int T, int Tt, int nb, void* stream) {
  if (B <= 0) return KR_OK;
  if (nb > 256 || (D % 4)) { kr_set_error("kr_expand_adapt: nb <= 256 and D % 4 == 0 required"); return KR_ERR_ARG; }
  AdaptParams p;
  p.enc = enc; p.idx = idx; p.lengths = lengths; p.pitch = pitch; p.energy = energy; p.flags = flags;
  p.pbins = pbins; p.ebins = ebins; p.pemb = pemb; p.eemb = eemb; p.row_of_tok = row_of_tok;
  p.xpad = (bf16*)xpad; p.mem = (bf16*)mem; p.p_idx = p_idx; p.e_idx = e_idx; p.fmask_t = fmask_t;
  p.fmask_p = fmask_p; p.B = B; p.P = P; p.D = D; p.Tp = Tp; p.T = T; p.Tt = Tt; p.nb = nb;
  const long long rows = (long long)B * (Tp > T ? Tp : T);
  kr::launch(expand_adapt_kernel, warp_blocks(rows, 16), WARPS * 32, 0, (cudaStream_t)stream, p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_adapt_bwd(const float* dmem, const int* p_idx, const int* e_idx, float* dpemb,
                            float* deemb, long long rows, int D, void* stream) {
  if (rows <= 0) return KR_OK;
  kr::launch(adapt_bwd_kernel, warp_blocks(rows, 16), WARPS * 32, 0, (cudaStream_t)stream, dmem, p_idx, e_idx, dpemb, deemb, rows, D);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_gn_fwd(const float* x, const int* row_group, const int* group_rows, double* stats,
                         const float* gamma, const float* beta, void* out_bf16, int R, int C, int G,
                         const kr_drop_spec* drop, void* stream) {
  if (R <= 0) return KR_OK;
  if (C > 256 || (C % 32)) { kr_set_error("kr_gn: C must be a multiple of 32, <= 256"); return KR_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * G, st) != cudaSuccess) { kr_set_error("memset failed"); return KR_ERR_CUDA; }
  kr::launch(gn_stats_kernel, warp_blocks(R, 2), WARPS * 32, 0, st, x, row_group, stats, R, C);
  KR_CHECK_LAUNCH();
  kr::launch(gn_apply_relu_kernel, warp_blocks(R), WARPS * 32, 0, st, x, row_group, stats, group_rows, gamma, beta, (bf16*)out_bf16, R, C,
             kr_drop_to_device(drop));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_gn_bwd(const void* dy_bf16, const float* x, const int* row_group, const int* group_rows,
                         const double* stats, double* gsum, const float* gamma, const float* beta,
                         void* dx_bf16, float* dgamma, float* dbeta, int R, int C, int G,
                         const kr_drop_spec* drop, void* stream) {
  if (R <= 0) return KR_OK;
  if (C > 256 || (C % 32)) { kr_set_error("kr_gn: C must be a multiple of 32, <= 256"); return KR_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(gsum, 0, sizeof(double) * 2 * G, st) != cudaSuccess) { kr_set_error("memset failed"); return KR_ERR_CUDA; }
  kr::launch(gn_bwd_stats_kernel, warp_blocks(R, 2), WARPS * 32, 0, st, (const bf16*)dy_bf16, x, row_group, stats, group_rows, gamma, beta, gsum, dgamma, dbeta, R, C,
             kr_drop_to_device(drop));
  KR_CHECK_LAUNCH();
  kr::launch(gn_bwd_apply_kernel, warp_blocks(R), WARPS * 32, 0, st, (const bf16*)dy_bf16, x, row_group, stats, group_rows, gsum, gamma, beta, (bf16*)dx_bf16, R, C,
             kr_drop_to_device(drop));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_vp_head_fwd(const void* h, const int* row_of_tok, const float* w, const float* bias,
                              const unsigned char* mask, float* out, int n_tok, int L, int F, int chunk,
                              void* stream) {
  if (n_tok <= 0) return KR_OK;
  kr::launch(vp_head_fwd_kernel, warp_blocks(n_tok), WARPS * 32, 0, (cudaStream_t)stream, (const bf16*)h, row_of_tok, w, bias, mask, out, n_tok, L, F, chunk);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_vp_head_bwd(const float* dout, const void* h, const int* tok_of_row, const float* w,
                              const unsigned char* mask, void* dh, float* dw, float* dbias, int R, int L,
                              int F, int chunk, void* stream) {
  if (R <= 0) return KR_OK;
  if (F > 256) { kr_set_error("kr_vp_head: F <= 256 required"); return KR_ERR_ARG; }
  kr::launch(vp_head_bwd_kernel, warp_blocks(R, 2), WARPS * 32, 0, (cudaStream_t)stream, dout, (const bf16*)h, tok_of_row, w, mask, (bf16*)dh, dw, dbias, R, L, F, chunk);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_conv_dgrad_shadow(const float* w2, void* wd, int Co, int Ci, void* stream) {
  const long long total = (long long)Ci * 3 * Co;
  long long b = (total + 255) / 256;
  kr::launch(conv_dgrad_shadow_kernel, (int)(b < 1184 ? b : 1184), 256, 0, (cudaStream_t)stream, w2, (bf16*)wd, Co, Ci);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_conv_dgrad_shadow_multi(const float* const* w2, void* const* wd, int n, int Co, int Ci, void* stream) {
  if (n <= 0) return KR_OK;
  if (n > 8) { kr_set_error("kr_conv_dgrad_shadow_multi: at most 8 convs per launch"); return KR_ERR_ARG; }
  ShadowBatch b{};
  for (int i = 0; i < n; ++i) { b.w2[i] = w2[i]; b.wd[i] = reinterpret_cast<bf16*>(wd[i]); }
  const long long total = (long long)Ci * 3 * Co;
  long long blocks = (total + 255) / 256;
  if (blocks > 296) blocks = 296;
  kr::launch(conv_dgrad_shadow_multi_kernel, dim3((unsigned)blocks, (unsigned)n), 256, 0, (cudaStream_t)stream, b, Co, Ci);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// ---------------------------------------------------------------------------------------------
// SpecAugment on the cross-attention memory (reference training/trainer.py:1578-1604, hooked in at
// model/model.py:636-639): per sample, n_time spans of frames and n_feat spans of hidden dims are
// zeroed.  The spans are drawn on the host (same torch.randint sequence as the reference) and live in
// a small device table spans[B, n_time + n_feat, 2] = (start, length); the same call masks the memory
// gradient in the backward pass (is_f32 = 1).
// ---------------------------------------------------------------------------------------------
namespace {
template <typename T>
__global__ void spec_augment_kernel(T* __restrict__ x, const int* __restrict__ spans, int B, int Tn, int D, int n_time,
                                    int n_feat) {
  kr::pdl_entry();
  const int b = blockIdx.y;
  const int* sp = spans + (long long)b * (n_time + n_feat) * 2;
  T* xb = x + (long long)b * Tn * D;
  const T zero = T(0.f);
  // time spans: whole rows
  for (int k = 0; k < n_time; ++k) {
    const int t0 = sp[2 * k], tl = sp[2 * k + 1];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)tl * D;
         e += (long long)gridDim.x * blockDim.x) {
      const int t = t0 + (int)(e / D);
      if (t < Tn) xb[(long long)t * D + (e % D)] = zero;
    }
  }
  // feature spans: a few columns of every row
  for (int k = 0; k < n_feat; ++k) {
    const int f0 = sp[2 * (n_time + k)], fl = sp[2 * (n_time + k) + 1];
    if (fl <= 0) continue;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)Tn * fl;
         e += (long long)gridDim.x * blockDim.x) {
      const int t = (int)(e / fl), f = f0 + (int)(e % fl);
      if (f < D) xb[(long long)t * D + f] = zero;
    }
  }
}
}  // namespace

extern "C" int kr_spec_augment(void* x, int is_f32, const int* spans, int B, int T, int D, int n_time, int n_feat,
                               void* stream) {
  if (B <= 0 || T <= 0 || (n_time + n_feat) <= 0) return KR_OK;
  dim3 grid(8, B);
  if (is_f32) kr::launch(spec_augment_kernel<float>, grid, 256, 0, (cudaStream_t)stream, (float*)x, spans, B, T, D, n_time, n_feat);
  else kr::launch(spec_augment_kernel<bf16>, grid, 256, 0, (cudaStream_t)stream, (bf16*)x, spans, B, T, D, n_time, n_feat);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
