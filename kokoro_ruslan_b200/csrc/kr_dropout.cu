// kr_dropout.cu — step-level plumbing of the counter-based dropout / stochastic-depth RNG
// (kr_common.cuh "Dropout / stochastic depth").
//   * kr_drop_begin        advances the {seed, step} state once per training forward and fills the
//                          per-sample stochastic-depth factor table of every residual branch
//                          (reference drop_path, model/transformers.py:16-40; rates model/model.py:99-107)
//   * kr_dec_in_drop       decoder input: dropout(dropout(proj, p_in) + PE, p_pe)
//                          (model/model.py:525-531, model/positional_encoding.py:72-74)
//   * kr_drop_export_mask  the keep mask of one site as bytes — test / debugging aid: lets the CPU
//                          oracle apply EXACTLY the masks the fused kernels regenerate
#include "kr_common.cuh"
#include "kokoro_b200.h"

namespace {
using namespace kr;

__global__ void drop_begin_kernel(unsigned long long* state, const int* __restrict__ path_site,
                                  const float* __restrict__ path_p, float* __restrict__ table, int n_sites, int B) {
  kr::pdl_entry();
  if (threadIdx.x == 0) state[1] += 1ull;
  __syncthreads();
  for (int i = threadIdx.x; i < n_sites * B; i += blockDim.x) {
    const int s = i / B, b = i % B;
    const float p = path_p[s];
    float f = 1.f;
    if (p > 0.f) {
      const uint32_t thr = (uint32_t)(p * 65536.f + 0.5f);
      const uint2 k = drop_key(state, (uint32_t)path_site[s]);
      const uint32_t x = drop_hash((uint32_t)b >> 1, k);
      const uint32_t lane16 = (b & 1) ? (x >> 16) : (x & 0xffffu);
      f = lane16 >= thr ? 65536.f / (65536.f - (float)thr) : 0.f;
    }
    table[i] = f;
  }
}

__global__ void dec_in_drop_kernel(const float* __restrict__ t, const float* __restrict__ pe, float* __restrict__ y,
                                   long long n4, int T, int D, const DropSpec d, float scale_a) {
  kr::pdl_entry();
  DropCtx ca, cb;
  ca.ka = drop_key(d.state, d.site_a); ca.kb = make_uint2(0u, 0u); ca.thr_a = d.thr_a; ca.thr_b = 0; ca.scale = scale_a;
  cb.ka = drop_key(d.state, d.site_b); cb.kb = make_uint2(0u, 0u); cb.thr_a = d.thr_b; cb.thr_b = 0;
  cb.scale = d.scale / scale_a;
  const int vec_per_row = D / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec_per_row;
    const int c = (int)(i % vec_per_row) * 4;
    const float4 v = *reinterpret_cast<const float4*>(t + 4 * i);
    const float4 p = *reinterpret_cast<const float4*>(pe + (row % T) * D + c);
    const float4 fa = drop_quad(ca, 4 * i), fb = drop_quad(cb, 4 * i);
    *reinterpret_cast<float4*>(y + 4 * i) = make_float4(fb.x * (fa.x * v.x + p.x), fb.y * (fa.y * v.y + p.y),
                                                       fb.z * (fa.z * v.z + p.z), fb.w * (fa.w * v.w + p.w));
  }
}

__global__ void drop_export_kernel(const unsigned long long* state, uint32_t site, uint32_t thr, long long rows,
                                   int cols, long long ld, unsigned char* __restrict__ out, int byte_lanes) {
  kr::pdl_entry();
  const uint2 k = drop_key(state, site);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i % cols;
    const long long e = r * ld + c;
    if (byte_lanes) {      // attention sites: 4 elements per hash, 8-bit lanes, threshold thr >> 8
      const uint32_t x = drop_hash((uint32_t)(e >> 2), k);
      out[i] = ((x >> (8 * (e & 3))) & 0xffu) >= (thr >> 8) ? 1 : 0;
    } else {
      const uint32_t x = drop_hash((uint32_t)(e >> 1), k);
      const uint32_t lane16 = (e & 1) ? (x >> 16) : (x & 0xffffu);
      out[i] = lane16 >= thr ? 1 : 0;
    }
  }
}

}  // namespace

kr::DropSpec kr_drop_to_device(const kr_drop_spec* s) {
  kr::DropSpec d{};
  if (s != nullptr && s->state != nullptr) {
    d.state = s->state; d.site_a = s->site_a; d.thr_a = s->thr_a; d.site_b = s->site_b; d.thr_b = s->thr_b;
    d.scale = s->scale; d.row_scale = s->row_scale; d.rows_per_sample = s->rows_per_sample > 0 ? s->rows_per_sample : 1;
  }
  return d;
}

extern "C" int kr_drop_begin(unsigned long long* state, const int* path_site, const float* path_p, float* table,
                             int n_sites, int B, void* stream) {
  if (state == nullptr) { kr_set_error("kr_drop_begin: null state"); return KR_ERR_ARG; }
  kr::launch(drop_begin_kernel, 1, 256, 0, (cudaStream_t)stream, state, path_site, path_p, table, n_sites, B);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_dec_in_drop(const float* t, const float* pe, float* y, int N, int T, int D, float scale_a,
                              const kr_drop_spec* drop, void* stream) {
  if (N <= 0) return KR_OK;
  if (drop == nullptr || drop->state == nullptr || (D % 4)) { kr_set_error("kr_dec_in_drop: needs a drop spec and D % 4 == 0"); return KR_ERR_ARG; }
  const long long n4 = (long long)N * D / 4;
  long long b = (n4 + 255) / 256;
  const long long cap = (long long)kr::kNumSMs * 16;
  kr::launch(dec_in_drop_kernel, (int)(b < cap ? b : cap), 256, 0, (cudaStream_t)stream, t, pe, y, n4, T, D,
             kr_drop_to_device(drop), scale_a);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_drop_export_mask(const unsigned long long* state, unsigned int site, unsigned int thr,
                                   long long rows, int cols, long long ld, unsigned char* out, int byte_lanes,
                                   void* stream) {
  if (rows <= 0 || cols <= 0) return KR_OK;
  long long b = (rows * cols + 255) / 256;
  const long long cap = (long long)kr::kNumSMs * 16;
  kr::launch(drop_export_kernel, (int)(b < cap ? b : cap), 256, 0, (cudaStream_t)stream, state, site, thr, rows, cols, ld, out,
             byte_lanes);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
