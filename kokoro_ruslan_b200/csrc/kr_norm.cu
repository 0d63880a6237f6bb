// kr_norm.cu — HBM-bound normalisation kernels of the transformer blocks (warp-per-row, float4
// loads, warp-shuffle reductions; column gradients reduced per block then one atomic per column).
//   * LayerNorm fwd/bwd            (reference model/transformers.py:478,485,564,572,581,660; eps 1e-5)
//   * FFN output RMSNorm + residual (transformers.py:94,105-111; nn.RMSNorm(eps=None) -> finfo.eps)
//   * per-head Q/K/V RMSNorm with learned gain + rotate-half RoPE, fwd/bwd
//     (transformers.py:145-148,260-272; positional_encoding.py:196-209)
#include "kr_common.cuh"
#include "kokoro_b200.h"
#include <float.h>
#include <stdlib.h>

namespace {
using namespace kr;

constexpr int WARPS = 8;  // warps per block for the row kernels

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_bf16x4(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16(v.x, v.y);
  u.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward: x f32 [N,D] -> y bf16 [N,D] (+ optional f32 copy), mean/rstd [N]
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ beta, bf16* __restrict__ y,
                              float* __restrict__ y32, float* __restrict__ mean_out,
                              float* __restrict__ rstd_out, int N, float eps) {
  kr::pdl_entry();
  constexpr int D = NV * 128;
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* xr = x + (long long)row * D;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = ld4(xr + i * 128 + lane * 4);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = i * 128 + lane * 4;
    const float4 g = ld4(gamma + c0), b = ld4(beta + c0);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (y != nullptr) st_bf16x4(y + (long long)row * D + c0, o);
    if (y32 != nullptr) st4(y32 + (long long)row * D + c0, o);
  }
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward: dx = dres + rstd*(g*dy - mean(g*dy) - xhat*mean(g*dy*xhat))
// ---------------------------------------------------------------------------------------------
template <int NV, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                              const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                              const float* __restrict__ gamma, const float* dres, float* dx,
                              bf16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, int N, const DropSpec drop,
                              float* __restrict__ dcol_bf16) {
  kr::pdl_entry();
  constexpr int D = NV * 128;
  __shared__ float sm[WARPS][D];
  DropCtx dc{};
  if (drop.state != nullptr) dc = drop_ctx(drop);
  float4 ao[NV];      // column sums of what is written to dx_bf16 (= the bias gradient of the Linear it feeds)
#pragma unroll
  for (int i = 0; i < NV; ++i) ao[i] = make_float4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
  for (int row = blockIdx.x * WARPS + warp; row < N; row += gridDim.x * WARPS) {
    const long long off = (long long)row * D;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[NV], g[NV], rs[NV];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)          // residual gradient: issue its loads together with x / dy
      rs[i] = dres != nullptr ? ld4(dres + off + i * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = i * 128 + lane * 4;
      const float4 xv = ld4(x + off + c0), d = ld4(dy + off + c0), gm = ld4(gamma + c0);
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      c1 += g[i].x + g[i].y + g[i].z + g[i].w;
      c2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
      ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
    }
    c1 = warp_sum(c1) * (1.f / D);
    c2 = warp_sum(c2) * (1.f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = i * 128 + lane * 4;
      float4 o;
      o.x = rstd * (g[i].x - c1 - xh[i].x * c2);
      o.y = rstd * (g[i].y - c1 - xh[i].y * c2);
      o.z = rstd * (g[i].z - c1 - xh[i].z * c2);
      o.w = rstd * (g[i].w - c1 - xh[i].w * c2);
      o.x += rs[i].x; o.y += rs[i].y; o.z += rs[i].z; o.w += rs[i].w;
      st4(dx + off + c0, o);
      if (dx_bf16 != nullptr) {
        if (drop.state != nullptr) {   // the bf16 copy feeds the backward of a dropped residual branch
          const float rsf = drop_row_scale(drop, row);
          const float4 f = drop_quad(dc, off + c0);
          o.x *= f.x * rsf; o.y *= f.y * rsf; o.z *= f.z * rsf; o.w *= f.w * rsf;
        }
        st_bf16x4(dx_bf16 + off + c0, o);
        ao[i].x += o.x; ao[i].y += o.y; ao[i].z += o.z; ao[i].w += o.w;
      }
    }
  }
  // column gradients: warps -> smem -> one atomic per column per block
  const int n_pass = dcol_bf16 != nullptr ? 3 : 2;
  for (int pass = 0; pass < n_pass; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) st4(&sm[warp][i * 128 + lane * 4], pass == 0 ? ag[i] : (pass == 1 ? ab[i] : ao[i]));
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) t += sm[w][c];
      atomicAdd((pass == 0 ? dgamma : (pass == 1 ? dbeta : dcol_bf16)) + c, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FFN output RMSNorm + residual: out = resid + y * rsqrt(mean(y^2)+eps) * g
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void rms_resid_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gain,
                                     const float* resid, float* out, int N, float eps, const DropSpec drop) {
  kr::pdl_entry();
  constexpr int D = NV * 128;
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const long long off = (long long)row * D;
  DropCtx dc{};
  float rsf = 1.f;
  if (drop.state != nullptr) { dc = drop_ctx(drop); rsf = drop_row_scale(drop, row); }
  float4 v[NV];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = ld4(y + off + i * 128 + lane * 4);
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = i * 128 + lane * 4;
    const float4 g = ld4(gain + c0), r = ld4(resid + off + c0);
    float4 f = make_float4(1.f, 1.f, 1.f, 1.f);
    if (drop.state != nullptr) {     // dropout(s) + stochastic depth on the branch, transformers.py:111,486-487,581
      f = drop_quad(dc, off + c0);
      f.x *= rsf; f.y *= rsf; f.z *= rsf; f.w *= rsf;
    }
    st4(out + off + c0, make_float4(r.x + v[i].x * rstd * g.x * f.x, r.y + v[i].y * rstd * g.y * f.y,
                                    r.z + v[i].z * rstd * g.z * f.z, r.w + v[i].w * rstd * g.w * f.w));
  }
}

// ---------------------------------------------------------------------------------------------
// Sub-layer tails fused with the LayerNorm of the sub-layer that FOLLOWS (the forward is one dependent chain with a single
// kernel in flight, tools/step_timeline.py: a launch removed from it comes off the step):
//   resid_drop_ln_fwd   out = resid + dropout(y) (y = the attention out-projection incl. bias, transformers.py:482-483) and
//                       h = LayerNorm(out) — replaces the dropout / residual epilogue of the out-projection GEMM (18.8 us at
//                       6400 x 512 x 512 against 10.4 us plain: the epilogue of the CTA's only tile is exposed) AND ln_fwd
//   rms_resid_ln_fwd    rmsnorm_resid_fwd AND ln_fwd
// Same arithmetic, in the same order, as the kernels they replace: bit-identical residual stream and LayerNorm outputs.
// ---------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void ln_tail(const float4 (&v)[NV], int row, int lane, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, bf16* __restrict__ h, float* __restrict__ h32,
                                        float* __restrict__ mean_out, float* __restrict__ rstd_out, float eps) {
  constexpr int D = NV * 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = i * 128 + lane * 4;
    const float4 g = ld4(gamma + c0), b = ld4(beta + c0);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (h != nullptr) st_bf16x4(h + (long long)row * D + c0, o);
    if (h32 != nullptr) st4(h32 + (long long)row * D + c0, o);
  }
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

template <int NV>
__global__ void resid_drop_ln_fwd_kernel(const float* __restrict__ y, const float* resid, float* out, int N,
                                         const DropSpec drop, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, bf16* __restrict__ h, float* __restrict__ h32,
                                         float* __restrict__ mean_out, float* __restrict__ rstd_out, float eps) {
  kr::pdl_entry();
  constexpr int D = NV * 128;
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const long long off = (long long)row * D;
  DropCtx dc{};
  float rsf = 1.f;
  if (drop.state != nullptr) { dc = drop_ctx(drop); rsf = drop_row_scale(drop, row); }
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = i * 128 + lane * 4;
    float4 a = ld4(y + off + c0);
    const float4 r = ld4(resid + off + c0);
    if (drop.state != nullptr) {              // the GEMM epilogue's order: v *= f * row factor, then + resid
      const float4 f = drop_quad(dc, off + c0);
      a.x *= f.x * rsf; a.y *= f.y * rsf; a.z *= f.z * rsf; a.w *= f.w * rsf;
    }
    v[i] = make_float4(a.x + r.x, a.y + r.y, a.z + r.z, a.w + r.w);
    st4(out + off + c0, v[i]);
  }
  ln_tail<NV>(v, row, lane, gamma, beta, h, h32, mean_out, rstd_out, eps);
}

template <int NV>
__global__ void rms_resid_ln_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gain, const float* resid,
                                        float* out, int N, float eps_rms, const DropSpec drop,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        bf16* __restrict__ h, float* __restrict__ h32, float* __restrict__ mean_out,
                                        float* __restrict__ rstd_out, float eps) {
  kr::pdl_entry();
  constexpr int D = NV * 128;
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const long long off = (long long)row * D;
  DropCtx dc{};
  float rsf = 1.f;
  if (drop.state != nullptr) { dc = drop_ctx(drop); rsf = drop_row_scale(drop, row); }
  float4 v[NV];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = ld4(y + off + i * 128 + lane * 4);
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps_rms);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = i * 128 + lane * 4;
    const float4 g = ld4(gain + c0), r = ld4(resid + off + c0);
    float4 f = make_float4(1.f, 1.f, 1.f, 1.f);
    if (drop.state != nullptr) {
      f = drop_quad(dc, off + c0);
      f.x *= rsf; f.y *= rsf; f.z *= rsf; f.w *= rsf;
    }
    v[i] = make_float4(r.x + v[i].x * rstd * g.x * f.x, r.y + v[i].y * rstd * g.y * f.y,
                       r.z + v[i].z * rstd * g.z * f.z, r.w + v[i].w * rstd * g.w * f.w);
    st4(out + off + c0, v[i]);
  }
  ln_tail<NV>(v, row, lane, gamma, beta, h, h32, mean_out, rstd_out, eps);
}

template <int NV>
__global__ void __launch_bounds__(WARPS * 32, 3) rms_resid_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                     const float* __restrict__ gain, bf16* __restrict__ dy,
                                     float* __restrict__ dgain, int N, float eps, const DropSpec drop,
                                     float* __restrict__ dcol) {
  kr::pdl_entry();
  constexpr int D = NV * 128;
  __shared__ float sm[WARPS][D];
  DropCtx dc{};
  if (drop.state != nullptr) dc = drop_ctx(drop);
  float4 ao[NV];      // column sums of dy (= the bias gradient of linear2)
#pragma unroll
  for (int i = 0; i < NV; ++i) ao[i] = make_float4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 ag[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = make_float4(0, 0, 0, 0);
  for (int row = blockIdx.x * WARPS + warp; row < N; row += gridDim.x * WARPS) {
    const long long off = (long long)row * D;
    float4 v[NV], d[NV];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = ld4(y + off + i * 128 + lane * 4);
      q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
    float c = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = i * 128 + lane * 4;
      const float4 g = ld4(gain + c0);
      float4 o = ld4(dout + off + c0);
      if (drop.state != nullptr) {
        const float rsf = drop_row_scale(drop, row);
        const float4 f = drop_quad(dc, off + c0);
        o.x *= f.x * rsf; o.y *= f.y * rsf; o.z *= f.z * rsf; o.w *= f.w * rsf;
      }
      v[i] = make_float4(v[i].x * rstd, v[i].y * rstd, v[i].z * rstd, v[i].w * rstd);  // yhat
      ag[i].x += o.x * v[i].x; ag[i].y += o.y * v[i].y; ag[i].z += o.z * v[i].z; ag[i].w += o.w * v[i].w;
      d[i] = make_float4(o.x * g.x, o.y * g.y, o.z * g.z, o.w * g.w);
      c += d[i].x * v[i].x + d[i].y * v[i].y + d[i].z * v[i].z + d[i].w * v[i].w;
    }
    c = warp_sum(c) * (1.f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = i * 128 + lane * 4;
      const float4 o = make_float4(rstd * (d[i].x - v[i].x * c), rstd * (d[i].y - v[i].y * c),
                                   rstd * (d[i].z - v[i].z * c), rstd * (d[i].w - v[i].w * c));
      st_bf16x4(dy + off + c0, o);
      ao[i].x += o.x; ao[i].y += o.y; ao[i].z += o.z; ao[i].w += o.w;
    }
  }
  const int n_pass = dcol != nullptr ? 2 : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) st4(&sm[warp][i * 128 + lane * 4], pass == 0 ? ag[i] : ao[i]);
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) t += sm[w][c];
      atomicAdd((pass == 0 ? dgain : dcol) + c, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Q/K/V head prep: per-head RMSNorm (d_k = 64, learned gain shared by the heads) + optional RoPE.
// One warp handles one (token, part) = 8 heads x 64 per iteration; 4 lanes per head, 16 elements
// per lane; the rotate-half partner (i, i+32) lives in lane^2.
// ---------------------------------------------------------------------------------------------
struct PrepPart {
  const void* in;     // raw projection output (bf16) — fwd input / bwd saved input
  void* out;          // fwd: normalised bf16; bwd: d(raw) bf16
  const void* grad;   // bwd: incoming gradient (f32 if grad_f32 else bf16)
  const float* gain;  // [64]
  float* dgain;       // [64] atomics (bwd)
  long long ld_in, ld_out, ld_grad;
  int rope, grad_f32;
};
struct PrepParams {
  PrepPart part[3];
  int n_parts, N, S, H;
  const float* cos_t;  // [S_max, 32]
  const float* sin_t;
  float eps;
};

__device__ __forceinline__ void load16_bf16(const bf16* p, float* v) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 8);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float2 f = unpack_bf16(w[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store16_bf16(bf16* p, const float* v) {
  uint4 a, b;
  a.x = pack_bf16(v[0], v[1]); a.y = pack_bf16(v[2], v[3]); a.z = pack_bf16(v[4], v[5]); a.w = pack_bf16(v[6], v[7]);
  b.x = pack_bf16(v[8], v[9]); b.y = pack_bf16(v[10], v[11]); b.z = pack_bf16(v[12], v[13]); b.w = pack_bf16(v[14], v[15]);
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + 8) = b;
}

// unit = (row, head) of ONE part (blockIdx.y); a warp handles 8 consecutive units per iteration
// (4 lanes each) = one full 512-wide row when H = 8, i.e. 1 KB coalesced bf16 per warp access.
struct PrepUnit { int row, col; bool valid; };
__device__ __forceinline__ PrepUnit prep_unit(const PrepParams& p, long long w, int lane, long long total) {
  // 32-bit arithmetic (the launchers refuse N * H >= 2^31): two 64-bit divisions per 16 elements were ~a quarter of the
  // instructions of these issue-bound kernels
  unsigned u = (unsigned)w * 8u + (unsigned)(lane >> 2);
  PrepUnit r;
  r.valid = u < (unsigned)total;
  if (!r.valid) u = (unsigned)total - 1u;
  const unsigned row = u / (unsigned)p.H;
  const int head = (int)(u - row * (unsigned)p.H);
  r.row = (int)row;
  r.col = head * 64 + (lane & 3) * 16;
  return r;
}

__device__ __forceinline__ void load16_f32(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p) + i);
    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
  }
}

__global__ void __launch_bounds__(WARPS * 32) qkv_prep_fwd_kernel(const PrepParams p) {
  kr::pdl_entry();
  const int lane = threadIdx.x & 31;
  const int sub = lane & 3;                // 16-element slice of the head
  const PrepPart& pp = p.part[blockIdx.y];
  const long long total = (long long)p.N * p.H;
  const long long n_iter = (total + 7) / 8;
  float gn[16];
  load16_f32(pp.gain + sub * 16, gn);
  const float sgn = (sub < 2) ? -1.f : 1.f;   // rot(x)[i] = -x[i+32] (i<32), +x[i-32] (i>=32)
  for (long long w = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); w < n_iter;
       w += (long long)gridDim.x * WARPS) {
    const PrepUnit un = prep_unit(p, w, lane, total);
    float v[16];
    load16_bf16(reinterpret_cast<const bf16*>(pp.in) + (long long)un.row * pp.ld_in + un.col, v);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) q += v[i] * v[i];
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    const float rstd = rsqrtf(q * (1.f / 64.f) + p.eps);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] * rstd * gn[i];
    if (pp.rope) {
      const int pos = un.row % p.S;
      float ct[16], st[16];
      load16_f32(p.cos_t + (long long)pos * 32 + (sub & 1) * 16, ct);
      load16_f32(p.sin_t + (long long)pos * 32 + (sub & 1) * 16, st);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float other = __shfl_xor_sync(0xffffffffu, v[i], 2);
        v[i] = v[i] * ct[i] + sgn * other * st[i];
      }
    }
    if (un.valid)
      store16_bf16(reinterpret_cast<bf16*>(pp.out) + (long long)un.row * pp.ld_out + un.col, v);
  }
}

template <int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) qkv_prep_bwd_kernel(const PrepParams p) {
  kr::pdl_entry();
  __shared__ float sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane & 3;
  const PrepPart& pp = p.part[blockIdx.y];
  float dg[16], gn[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dg[i] = 0.f;
  load16_f32(pp.gain + sub * 16, gn);
  const float sgn = (sub < 2) ? 1.f : -1.f;
  const long long total = (long long)p.N * p.H;
  const long long n_iter = (total + 7) / 8;
  for (long long w = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); w < n_iter;
       w += (long long)gridDim.x * WARPS) {
    const PrepUnit un = prep_unit(p, w, lane, total);
    float x[16], g[16];
    load16_bf16(reinterpret_cast<const bf16*>(pp.in) + (long long)un.row * pp.ld_in + un.col, x);
    if (pp.grad_f32)
      load16_f32(reinterpret_cast<const float*>(pp.grad) + (long long)un.row * pp.ld_grad + un.col, g);
    else
      load16_bf16(reinterpret_cast<const bf16*>(pp.grad) + (long long)un.row * pp.ld_grad + un.col, g);
    if (pp.rope) {  // transpose of the rotation: dz[i] = dy[i] cos_i + (i<32 ? +dy[i+32] : -dy[i-32]) sin_i
      const int pos = un.row % p.S;
      float ct[16], st[16];
      load16_f32(p.cos_t + (long long)pos * 32 + (sub & 1) * 16, ct);
      load16_f32(p.sin_t + (long long)pos * 32 + (sub & 1) * 16, st);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float other = __shfl_xor_sync(0xffffffffu, g[i], 2);
        g[i] = g[i] * ct[i] + sgn * other * st[i];
      }
    }
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) q += x[i] * x[i];
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    const float rstd = rsqrtf(q * (1.f / 64.f) + p.eps);
    float c = 0.f;
    const float vf = un.valid ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float xh = x[i] * rstd;
      dg[i] += g[i] * xh * vf;                         // d gain
      g[i] *= gn[i];                                   // d xhat
      c += g[i] * xh;
      x[i] = xh;
    }
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c *= (1.f / 64.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] = rstd * (g[i] - x[i] * c);
    if (un.valid)
      store16_bf16(reinterpret_cast<bf16*>(pp.out) + (long long)un.row * pp.ld_out + un.col, g);
  }
  // lanes with equal `sub` hold the same 16 gain slots: fold across the 8 units, then block, then global
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float t = dg[i];
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    if (lane < 4) atomicAdd(&sm[sub * 16 + i], t);
  }
  __syncthreads();
  if (threadIdx.x < 64 && pp.dgain != nullptr) atomicAdd(pp.dgain + threadIdx.x, sm[threadIdx.x]);
}

// Tried and dropped (round 2): ln_bwd with the next row's x / dy prefetched in registers and the three column accumulators
// in per-warp shared memory rows (to free the registers): 14.9 vs 9.1 us at 6400 x 512 — the shared-memory read-modify-write
// per row and 184 bytes of spills cost more than the overlapped load latency gains.
// Also dropped: qkv_prep_fwd with gain / cos / sin fetched four values at a time at their use (60 registers, 4 blocks per SM,
// instead of 80 / 3): 10.7 vs 9.9 us for the three-part self-attention call; at 48 / 40 registers the spills make it 14 - 17 us.
// Launch geometry, fixed by sweeps on B200 (round 2, bench shape; the sweep knobs are gone):
//   * ln_bwd runs at 2 blocks / SM (launch bounds): 3 blocks / SM would cap it at 80 registers and spill ~260 bytes since
//     the dropout specs were added;
//   * persistent backward kernels (their column-gradient epilogue costs D atomics per block): 2 blocks per SM beat 1 and 4
//     (ln_bwd 9.1 vs 10.7 us);
//   * qkv_prep: 8 blocks per SM forward, 2 backward (18.1 vs 22.3 us at 6).
int row_blocks(int N) { return (N + WARPS - 1) / WARPS; }
int persistent_blocks(int N) { return min(row_blocks(N), kNumSMs * 2); }
int prep_blocks_per_sm(bool bwd) { return bwd ? 2 : 8; }

}  // namespace

#define DISPATCH_NV(D, CALL)                                          \
  switch (D) {                                                        \
    case 128: { constexpr int NV = 1; CALL; break; }                  \
    case 256: { constexpr int NV = 2; CALL; break; }                  \
    case 512: { constexpr int NV = 4; CALL; break; }                  \
    default: kr_set_error("hidden dim must be 128, 256 or 512"); return KR_ERR_UNSUPPORTED; \
  }

extern "C" int kr_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16,
                                float* y_f32, float* mean, float* rstd, int N, int D, float eps,
                                void* stream) {
  if (N <= 0) return KR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_NV(D, (kr::launch(ln_fwd_kernel<NV>, row_blocks(N), WARPS * 32, 0, st, 
                     x, gamma, beta, reinterpret_cast<bf16*>(y_bf16), y_f32, mean, rstd, N, eps)));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                                const float* gamma, const float* dres, float* dx, void* dx_bf16,
                                float* dgamma, float* dbeta, int N, int D, const kr_drop_spec* drop_bf16,
                                float* dcol_bf16, void* stream) {
  if (N <= 0) return KR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_NV(D, (kr::launch(ln_bwd_kernel<NV, 2>, persistent_blocks(N), WARPS * 32, 0, st,
                     dy, x, mean, rstd, gamma, dres, dx, reinterpret_cast<bf16*>(dx_bf16), dgamma,
                     dbeta, N, kr_drop_to_device(drop_bf16), dcol_bf16)));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_rmsnorm_resid_fwd(const float* y, const float* gain, const float* resid, float* out,
                                    int N, int D, const kr_drop_spec* drop, void* stream) {
  if (N <= 0) return KR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_NV(D, (kr::launch(rms_resid_fwd_kernel<NV>, row_blocks(N), WARPS * 32, 0, st, y, gain, resid, out, N,
                                                                                FLT_EPSILON, kr_drop_to_device(drop))));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_resid_drop_ln_fwd(const float* y, const float* resid, float* out, int N, int D, const kr_drop_spec* drop,
                                    const float* ln_gamma, const float* ln_beta, void* h_bf16, float* h_f32, float* mean,
                                    float* rstd, float ln_eps, void* stream) {
  if (N <= 0) return KR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_NV(D, (kr::launch(resid_drop_ln_fwd_kernel<NV>, row_blocks(N), WARPS * 32, 0, st, y, resid, out, N,
                             kr_drop_to_device(drop), ln_gamma, ln_beta, reinterpret_cast<bf16*>(h_bf16), h_f32, mean, rstd,
                             ln_eps)));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_rmsnorm_resid_ln_fwd(const float* y, const float* gain, const float* resid, float* out, int N, int D,
                                       const kr_drop_spec* drop, const float* ln_gamma, const float* ln_beta, void* h_bf16,
                                       float* h_f32, float* mean, float* rstd, float ln_eps, void* stream) {
  if (N <= 0) return KR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_NV(D, (kr::launch(rms_resid_ln_fwd_kernel<NV>, row_blocks(N), WARPS * 32, 0, st, y, gain, resid, out, N,
                             FLT_EPSILON, kr_drop_to_device(drop), ln_gamma, ln_beta, reinterpret_cast<bf16*>(h_bf16), h_f32,
                             mean, rstd, ln_eps)));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_rmsnorm_resid_bwd(const float* dout, const float* y, const float* gain, void* dy_bf16,
                                    float* dgain, int N, int D, const kr_drop_spec* drop, float* dcol,
                                    void* stream) {
  if (N <= 0) return KR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_NV(D, (kr::launch(rms_resid_bwd_kernel<NV>, persistent_blocks(N), WARPS * 32, 0, st, 
                     dout, y, gain, reinterpret_cast<bf16*>(dy_bf16), dgain, N, FLT_EPSILON,
                     kr_drop_to_device(drop), dcol)));
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// parts: up to 3 column blocks of H*64 each (e.g. q|k|v of one fused projection output).
// in/out/grad pointers and leading dimensions per part; rope_mask bit i = apply RoPE to part i;
// grad_f32_mask bit i = incoming gradient of part i is fp32 (else bf16).  Position = row % S.
extern "C" int kr_qkv_prep_fwd(const void* in0, const void* in1, const void* in2, void* out0, void* out1,
                               void* out2, long long ld_in, long long ld_out, const float* gain0,
                               const float* gain1, const float* gain2, int n_parts, int rope_mask,
                               const float* cos_t, const float* sin_t, int N, int S, int H,
                               void* stream) {
  if (N <= 0) return KR_OK;
  PrepParams p{};
  const void* ins[3] = {in0, in1, in2};
  void* outs[3] = {out0, out1, out2};
  const float* gains[3] = {gain0, gain1, gain2};
  for (int i = 0; i < n_parts; ++i) {
    p.part[i].in = ins[i]; p.part[i].out = outs[i]; p.part[i].gain = gains[i];
    p.part[i].ld_in = ld_in; p.part[i].ld_out = ld_out; p.part[i].rope = (rope_mask >> i) & 1;
  }
  p.n_parts = n_parts; p.N = N; p.S = S; p.H = H; p.cos_t = cos_t; p.sin_t = sin_t; p.eps = FLT_EPSILON;
  if ((long long)N * H >= (1LL << 31) - 8) { kr_set_error("kr_qkv_prep: N * H must stay below 2^31"); return KR_ERR_UNSUPPORTED; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = ((long long)N * H + 7) / 8;
  const long long nb_ = (total + WARPS - 1) / WARPS;
  const int per_part = kNumSMs * prep_blocks_per_sm(false) / n_parts;
  const int blocks = (int)(nb_ < per_part ? nb_ : per_part);
  kr::launch(qkv_prep_fwd_kernel, dim3(blocks, n_parts), WARPS * 32, 0, st, p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_qkv_prep_bwd(const void* in0, const void* in1, const void* in2, const void* grad0,
                               const void* grad1, const void* grad2, long long ld_grad0,
                               long long ld_grad1, long long ld_grad2, void* out0, void* out1, void* out2,
                               long long ld_in, long long ld_out, const float* gain0, const float* gain1,
                               const float* gain2, float* dgain0, float* dgain1, float* dgain2,
                               int n_parts, int rope_mask, int grad_f32_mask, const float* cos_t,
                               const float* sin_t, int N, int S, int H, void* stream) {
  if (N <= 0) return KR_OK;
  PrepParams p{};
  const void* ins[3] = {in0, in1, in2};
  const void* grads[3] = {grad0, grad1, grad2};
  const long long ldg[3] = {ld_grad0, ld_grad1, ld_grad2};
  void* outs[3] = {out0, out1, out2};
  const float* gains[3] = {gain0, gain1, gain2};
  float* dgains[3] = {dgain0, dgain1, dgain2};
  for (int i = 0; i < n_parts; ++i) {
    p.part[i].in = ins[i]; p.part[i].out = outs[i]; p.part[i].grad = grads[i]; p.part[i].gain = gains[i];
    p.part[i].dgain = dgains[i]; p.part[i].ld_in = ld_in; p.part[i].ld_out = ld_out;
    p.part[i].ld_grad = ldg[i]; p.part[i].rope = (rope_mask >> i) & 1;
    p.part[i].grad_f32 = (grad_f32_mask >> i) & 1;
  }
  p.n_parts = n_parts; p.N = N; p.S = S; p.H = H; p.cos_t = cos_t; p.sin_t = sin_t; p.eps = FLT_EPSILON;
  if ((long long)N * H >= (1LL << 31) - 8) { kr_set_error("kr_qkv_prep: N * H must stay below 2^31"); return KR_ERR_UNSUPPORTED; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = ((long long)N * H + 7) / 8;
  const long long nb_ = (total + WARPS - 1) / WARPS;
  const int per_part = kNumSMs * prep_blocks_per_sm(true) / n_parts;
  const int blocks = (int)(nb_ < per_part ? nb_ : per_part);
  kr::launch(qkv_prep_bwd_kernel<2>, dim3(blocks, n_parts), WARPS * 32, 0, st, p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
