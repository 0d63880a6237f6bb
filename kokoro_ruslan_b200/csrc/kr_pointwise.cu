// kr_pointwise.cu — vectorised elementwise / gather / column-reduction kernels (all HBM-bound).
//   * GLU gate  u = gelu_erf(gate) * lin, fwd/bwd        (reference model/transformers.py:105-111)
//   * encoder input: emb[idx]*sqrt(D) + stress_emb[s] + PE, fwd/bwd (model/model.py:375-378,
//     model/positional_encoding.py:66-74; stress row 0 is padding_idx -> no gradient)
//   * decoder input shift-right + cast                     (model/model.py:519)
//   * column sums (bias gradients), f32->bf16 casts, row gather/scatter helpers
#include "kr_common.cuh"
#include "kokoro_b200.h"

namespace {
using namespace kr;

// GELU (erf form, the reference's nn.GELU() default) and its derivative for the GLU kernels, which are bound by
// instruction issue, not by memory (6400 x 1536 gates, operands in L2: erff alone is ~40 issue slots per element with both
// of its branches predicated): Abramowitz-Stegun 7.1.26, erf(u) = 1 - (a1 t + ... + a5 t^5) exp(-u^2), t = 1 / (1 + p u),
// |error| < 5e-7 in fp32 — three decimal orders below the bf16 resolution of the gated product — with ONE exponential shared
// by the CDF and the PDF (exp(-u^2) = exp(-x^2 / 2)).
struct GeluParts { float cdf, e; };      // cdf = Phi(x), e = exp(-x^2 / 2)
__device__ __forceinline__ GeluParts gelu_parts(float x) {
  const float u = fabsf(x) * 0.70710678118654752f;
  float t;                                                          // 1 / (1 + p u): one MUFU.RCP (>= 22 good bits)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.f)));
  const float e = __expf(-u * u);
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.f);                 // erf(|x| / sqrt 2)
  GeluParts r;
  r.cdf = 0.5f * (1.f + copysignf(erf_abs, x));
  r.e = e;
  return r;
}
__device__ __forceinline__ float gelu_erf(float x) { return x * gelu_parts(x).cdf; }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const GeluParts g = gelu_parts(x);
  return fmaf(x * 0.3989422804014327f, g.e, g.cdf);              // cdf + x * pdf
}

// dropout factors of the 8 consecutive elements starting at e0 (all 1 when disabled)
__device__ __forceinline__ void drop8(const DropSpec& d, const DropCtx& c, long long e0, float* f) {
  if (d.state != nullptr) {
    const float4 a = drop_quad(c, e0), b = drop_quad(c, e0 + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 1.f;
  }
}

__global__ void glu_fwd_kernel(const bf16* __restrict__ h, bf16* __restrict__ u, long long n_vec, int FF,
                               const DropSpec d) {
  kr::pdl_entry();
  DropCtx dc{};
  if (d.state != nullptr) dc = drop_ctx(d);
  // one thread = 8 consecutive output columns; a block walks whole rows (threadIdx.x = the column vector, no per-element
  // 64-bit division: it was ~5 of the ~50 issue slots per element of this instruction-bound kernel)
  const int vec_per_row = FF / 8;
  const long long n_rows = n_vec / vec_per_row;
  if ((int)threadIdx.x >= vec_per_row) return;
  const int c = (int)threadIdx.x * 8;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const long long i = row * vec_per_row + threadIdx.x;
    const uint4 g = *reinterpret_cast<const uint4*>(h + row * 2 * FF + c);
    const uint4 l = *reinterpret_cast<const uint4*>(h + row * 2 * FF + FF + c);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, lw[4] = {l.x, l.y, l.z, l.w};
    uint32_t o[4];
    float f[8];
    drop8(d, dc, 8 * i, f);          // dropout on the gated product, transformers.py:108
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = unpack_bf16(gw[k]), b = unpack_bf16(lw[k]);
      o[k] = pack_bf16(gelu_erf(a.x) * b.x * f[2 * k], gelu_erf(a.y) * b.y * f[2 * k + 1]);
    }
    *reinterpret_cast<uint4*>(u + row * FF + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void glu_bwd_kernel(const bf16* __restrict__ du, const bf16* __restrict__ h,
                               bf16* __restrict__ dh, long long n_vec, int FF, const DropSpec drop) {
  kr::pdl_entry();
  DropCtx dc{};
  if (drop.state != nullptr) dc = drop_ctx(drop);
  const int vec_per_row = FF / 8;
  const long long n_rows = n_vec / vec_per_row;
  if ((int)threadIdx.x >= vec_per_row) return;
  const int c = (int)threadIdx.x * 8;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const long long i = row * vec_per_row + threadIdx.x;
    const uint4 g = *reinterpret_cast<const uint4*>(h + row * 2 * FF + c);
    const uint4 l = *reinterpret_cast<const uint4*>(h + row * 2 * FF + FF + c);
    const uint4 d = *reinterpret_cast<const uint4*>(du + row * FF + c);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, lw[4] = {l.x, l.y, l.z, l.w}, dw[4] = {d.x, d.y, d.z, d.w};
    uint32_t og[4], ol[4];
    float f[8];
    drop8(drop, dc, 8 * i, f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = unpack_bf16(gw[k]), b = unpack_bf16(lw[k]);
      float2 e = unpack_bf16(dw[k]);
      e.x *= f[2 * k]; e.y *= f[2 * k + 1];
      const GeluParts px = gelu_parts(a.x), py = gelu_parts(a.y);
      og[k] = pack_bf16(e.x * b.x * fmaf(a.x * 0.3989422804014327f, px.e, px.cdf),
                        e.y * b.y * fmaf(a.y * 0.3989422804014327f, py.e, py.cdf));
      ol[k] = pack_bf16(e.x * a.x * px.cdf, e.y * a.y * py.cdf);
    }
    *reinterpret_cast<uint4*>(dh + row * 2 * FF + c) = make_uint4(og[0], og[1], og[2], og[3]);
    *reinterpret_cast<uint4*>(dh + row * 2 * FF + FF + c) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

// out[c] += sum_n x[n, c]; block (32, 8) covers 256 columns x ROWS_PER_BLOCK rows.
constexpr int CS_ROWS = 128;
__global__ void colsum_bf16_kernel(const bf16* __restrict__ x, long long ld, float* __restrict__ out,
                                   int N, int C) {
  kr::pdl_entry();
  __shared__ float sm[8][256 + 8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c0 = blockIdx.x * 256 + tx * 8;
  const int r0 = blockIdx.y * CS_ROWS;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (c0 < C) {
    const int r1 = min(r0 + CS_ROWS, N);
    for (int r = r0 + ty; r < r1; r += 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + (long long)r * ld + c0);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16(w[k]);
        acc[2 * k] += f.x; acc[2 * k + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  const int t = ty * 32 + tx;  // 256 threads -> 256 columns
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += sm[k][t];
  const int c = blockIdx.x * 256 + t;
  if (c < C) atomicAdd(out + c, s);
}

__global__ void embed_fwd_kernel(const long long* __restrict__ idx, const long long* __restrict__ stress,
                                 const float* __restrict__ emb, const float* __restrict__ semb,
                                 const float* __restrict__ pe, float* __restrict__ x, int N, int P, int D,
                                 float scale, const DropSpec d) {
  kr::pdl_entry();
  DropCtx dc{};
  if (d.state != nullptr) dc = drop_ctx(d);
  const int vec_per_row = D / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * vec_per_row;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / vec_per_row), c = (int)(i % vec_per_row) * 4;
    const float4 e = *reinterpret_cast<const float4*>(emb + idx[n] * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pe + (long long)(n % P) * D + c);
    float4 o = make_float4(e.x * scale + p.x, e.y * scale + p.y, e.z * scale + p.z, e.w * scale + p.w);
    if (stress != nullptr) {
      const float4 s = *reinterpret_cast<const float4*>(semb + stress[n] * D + c);
      o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
    }
    if (d.state != nullptr) {          // PositionalEncoding's dropout, positional_encoding.py:74
      const float4 f = drop_quad(dc, 4 * i);
      o.x *= f.x; o.y *= f.y; o.z *= f.z; o.w *= f.w;
    }
    *reinterpret_cast<float4*>(x + (long long)n * D + c) = o;
  }
}

__global__ void embed_bwd_kernel(const float* __restrict__ dx, const long long* __restrict__ idx,
                                 const long long* __restrict__ stress, float* __restrict__ demb,
                                 float* __restrict__ dsemb, int N, int D, float scale, const DropSpec d) {
  kr::pdl_entry();
  DropCtx dc{};
  if (d.state != nullptr) dc = drop_ctx(d);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * D;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / D), c = (int)(i % D);
    float g = dx[i];
    if (d.state != nullptr) {
      float f0, f1;
      drop_pair(dc, (uint32_t)(i >> 1), f0, f1);
      g *= (i & 1) ? f1 : f0;
    }
    atomicAdd(demb + idx[n] * D + c, g * scale);
    if (stress != nullptr && stress[n] != 0) atomicAdd(dsemb + stress[n] * D + c, g);
  }
}

__global__ void shift_cast_kernel(const float* __restrict__ mel, bf16* __restrict__ out, int B, int T, int C) {
  kr::pdl_entry();
  const long long total = (long long)B * T * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)((i / C) % T);
    out[i] = __float2bfloat16(t > 0 ? mel[i - C] : 0.f);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  kr::pdl_entry();
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(in + 4 * i);
    uint2 u;
    u.x = pack_bf16(v.x, v.y);
    u.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + 4 * i) = u;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[n4 * 4 + threadIdx.x] = __float2bfloat16(in[n4 * 4 + threadIdx.x]);
}

// dst_row[map[r]] = src_row[r]  (f32 -> bf16 or f32 -> f32); map[r] < 0 skips the row.
template <typename TO>
__global__ void scatter_rows_kernel(const float* __restrict__ src, const int* __restrict__ map,
                                    TO* __restrict__ dst, int R, int C) {
  kr::pdl_entry();
  const int vec = C / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)R * vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec), c = (int)(i % vec) * 4;
    const int d = map[r];
    if (d < 0) continue;
    const float4 v = *reinterpret_cast<const float4*>(src + (long long)r * C + c);
    if (sizeof(TO) == 2) {
      uint2 u;
      u.x = pack_bf16(v.x, v.y);
      u.y = pack_bf16(v.z, v.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(dst) + (long long)d * C + c) = u;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + (long long)d * C + c) = v;
    }
  }
}
// dst_row[r] = src_row[map[r]]  (f32 -> f32); map[r] < 0 writes zeros.
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ map,
                                   float* __restrict__ dst, int R, int C) {
  kr::pdl_entry();
  const int vec = C / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)R * vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec), c = (int)(i % vec) * 4;
    const int s = map[r];
    float4 v = make_float4(0, 0, 0, 0);
    if (s >= 0) v = *reinterpret_cast<const float4*>(src + (long long)s * C + c);
    *reinterpret_cast<float4*>(dst + (long long)r * C + c) = v;
  }
}

// GLU kernels: one block walks rows, threadIdx.x = column vector (FF / 8 of them, rounded up to a warp multiple)
inline int glu_threads(int FF) { return ((FF / 8) + 31) / 32 * 32; }
inline int glu_blocks(int N) { const int cap = kNumSMs * 8; return N < cap ? (N > 0 ? N : 1) : cap; }

inline int ew_blocks(long long n, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int kr_glu_fwd(const void* h, void* u, int N, int FF, const kr_drop_spec* drop, void* stream) {
  if (N <= 0) return KR_OK;
  if (FF % 8 || FF > 8192) { kr_set_error("kr_glu: FF must be a multiple of 8, at most 8192"); return KR_ERR_ARG; }
  const long long n_vec = (long long)N * FF / 8;
  kr::launch(glu_fwd_kernel, glu_blocks(N), glu_threads(FF), 0, (cudaStream_t)stream, (const bf16*)h, (bf16*)u, n_vec, FF,
             kr_drop_to_device(drop));
  KR_CHECK_LAUNCH();
  return KR_OK;
}
extern "C" int kr_glu_bwd(const void* du, const void* h, void* dh, int N, int FF, const kr_drop_spec* drop,
                          void* stream) {
  if (N <= 0) return KR_OK;
  if (FF % 8 || FF > 8192) { kr_set_error("kr_glu: FF must be a multiple of 8, at most 8192"); return KR_ERR_ARG; }
  const long long n_vec = (long long)N * FF / 8;
  kr::launch(glu_bwd_kernel, glu_blocks(N), glu_threads(FF), 0, (cudaStream_t)stream, (const bf16*)du, (const bf16*)h, (bf16*)dh, n_vec, FF,
             kr_drop_to_device(drop));
  KR_CHECK_LAUNCH();
  return KR_OK;
}
extern "C" int kr_colsum_bf16(const void* x, long long ld, float* out, int N, int C, void* stream) {
  if (N <= 0 || C <= 0) return KR_OK;
  if ((C % 8) || (ld % 8)) { kr_set_error("kr_colsum_bf16: C and ld must be multiples of 8"); return KR_ERR_ARG; }
  dim3 grid((C + 255) / 256, (N + CS_ROWS - 1) / CS_ROWS), block(32, 8);
  kr::launch(colsum_bf16_kernel, grid, block, 0, (cudaStream_t)stream, (const bf16*)x, ld, out, N, C);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
extern "C" int kr_embed_fwd(const long long* idx, const long long* stress, const float* emb,
                            const float* stress_emb, const float* pe, float* x, int N, int P, int D,
                            const kr_drop_spec* drop, void* stream) {
  if (N <= 0) return KR_OK;
  kr::launch(embed_fwd_kernel, ew_blocks((long long)N * D / 4), 256, 0, (cudaStream_t)stream, 
      idx, stress, emb, stress_emb, pe, x, N, P, D, sqrtf((float)D), kr_drop_to_device(drop));
  KR_CHECK_LAUNCH();
  return KR_OK;
}
extern "C" int kr_embed_bwd(const float* dx, const long long* idx, const long long* stress, float* demb,
                            float* dstress_emb, int N, int D, const kr_drop_spec* drop, void* stream) {
  if (N <= 0) return KR_OK;
  kr::launch(embed_bwd_kernel, ew_blocks((long long)N * D), 256, 0, (cudaStream_t)stream, 
      dx, idx, stress, demb, dstress_emb, N, D, sqrtf((float)D), kr_drop_to_device(drop));
  KR_CHECK_LAUNCH();
  return KR_OK;
}
extern "C" int kr_shift_cast(const float* mel, void* out, int B, int T, int C, void* stream) {
  if (B <= 0 || T <= 0) return KR_OK;
  kr::launch(shift_cast_kernel, ew_blocks((long long)B * T * C), 256, 0, (cudaStream_t)stream, mel, (bf16*)out, B, T, C);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
extern "C" int kr_cast_bf16(const float* in, void* out, long long n, void* stream) {
  if (n <= 0) return KR_OK;
  kr::launch(cast_bf16_kernel, ew_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream, in, (bf16*)out, n);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
extern "C" int kr_scatter_rows(const float* src, const int* map, void* dst, int R, int C, int dst_bf16,
                               void* stream) {
  if (R <= 0) return KR_OK;
  if (C % 4) { kr_set_error("kr_scatter_rows: C must be a multiple of 4"); return KR_ERR_ARG; }
  const int blocks = ew_blocks((long long)R * C / 4);
  if (dst_bf16) kr::launch(scatter_rows_kernel<bf16>, blocks, 256, 0, (cudaStream_t)stream, src, map, (bf16*)dst, R, C);
  else          kr::launch(scatter_rows_kernel<float>, blocks, 256, 0, (cudaStream_t)stream, src, map, (float*)dst, R, C);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
extern "C" int kr_gather_rows(const float* src, const int* map, float* dst, int R, int C, void* stream) {
  if (R <= 0) return KR_OK;
  if (C % 4) { kr_set_error("kr_gather_rows: C must be a multiple of 4"); return KR_ERR_ARG; }
  kr::launch(gather_rows_kernel, ew_blocks((long long)R * C / 4), 256, 0, (cudaStream_t)stream, src, map, dst, R, C);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

namespace {
__global__ void eq_mask_kernel(const long long* __restrict__ idx, long long value, unsigned char* __restrict__ out, long long n) {
  kr::pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = idx[i] == value ? 1 : 0;
}
// any non-finite value in x -> flag |= bit
__global__ void nonfinite_flag_kernel(const float* __restrict__ x, long long n, int* flag, int bit) {
  kr::pdl_entry();
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    bad |= !isfinite(x[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, bit);
}
}  // namespace

// out[i] = (idx[i] == value)  — the reference's text padding mask `phoneme_indices == 0` (model/model.py:587)
extern "C" int kr_eq_mask_i64(const long long* idx, long long value, unsigned char* out, long long n, void* stream) {
  if (n <= 0) return KR_OK;
  kr::launch(eq_mask_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, idx, value, out, n);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
// finite-output guard without a host sync per tensor (reference training/trainer.py:3233-3256)
extern "C" int kr_nonfinite_flag(const float* x, long long n, int* flag, int bit, void* stream) {
  if (n <= 0) return KR_OK;
  kr::launch(nonfinite_flag_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, x, n, flag, bit);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
