// kr_hifigan.cu — the two non-GEMM ends of the HiFi-GAN generator (reference
// inference/hifigan_vocoder.py:110-133); everything in between is kr_gemm_ex in conv mode.
//   kr_hifi_pack_mel  : (B,80,T) or (B,T,80) fp32 mel -> channels-last bf16 [B, T + 2*halo, c_phys]
//                       (interior rows only; halo rows / padded channels stay zero) — the input layout
//                       handling of hifigan_vocoder.py:112-117 without a transposed copy in fp32.
//   kr_hifi_post_tanh : conv_post (C -> 1, k = 7, pad 3) + tanh on the channels-last activation
//                       (hifigan_vocoder.py:131-132).  N = 1 is a GEMV: HBM-bound, CUDA cores.
#include "kr_common.cuh"

namespace {
using namespace kr;

// block = 32 time steps of one item; smem tile [c][33]
__global__ void pack_mel_kernel(const float* __restrict__ mel, int time_major, bf16* __restrict__ out, int B, int T,
                                int n_mels, int halo, int c_phys) {
  kr::pdl_entry();
  extern __shared__ float tile[];  // [n_mels][33]
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (!time_major) {
    for (int c = warp; c < n_mels; c += nw) {
      const int t = t0 + lane;
      tile[c * 33 + lane] = t < T ? mel[((long long)b * n_mels + c) * T + t] : 0.f;
    }
  } else {
    for (int e = threadIdx.x; e < 32 * n_mels; e += blockDim.x) {
      const int tt = e / n_mels, c = e - tt * n_mels;
      const int t = t0 + tt;
      tile[c * 33 + tt] = t < T ? mel[((long long)b * T + t) * n_mels + c] : 0.f;
    }
  }
  __syncthreads();
  const long long Lp = T + 2 * halo;
  for (int e = threadIdx.x; e < 32 * (c_phys / 2); e += blockDim.x) {
    const int tt = e / (c_phys / 2), c2 = (e - tt * (c_phys / 2)) * 2;
    const int t = t0 + tt;
    if (t >= T) continue;
    const float a = c2 < n_mels ? tile[c2 * 33 + tt] : 0.f;
    const float bb = c2 + 1 < n_mels ? tile[(c2 + 1) * 33 + tt] : 0.f;
    *reinterpret_cast<uint32_t*>(out + ((long long)b * Lp + halo + t) * c_phys + c2) = pack_bf16(a, bb);
  }
}

// one thread per output sample; weights [taps][C] fp32 in smem
template <int TAPS>
__global__ void post_tanh_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ out, int B, long long L, int halo, int C, int c_phys) {
  kr::pdl_entry();
  extern __shared__ float sw[];  // [TAPS][C]
  for (int i = threadIdx.x; i < TAPS * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * L) return;
  const int b = (int)(idx / L);
  const long long t = idx - (long long)b * L;
  const long long Lp = L + 2 * halo;
  const bf16* base = x + ((long long)b * Lp + halo + t - TAPS / 2) * c_phys;
  float acc = bias[0];
#pragma unroll
  for (int j = 0; j < TAPS; ++j) {
    const bf16* row = base + (long long)j * c_phys;
    for (int c = 0; c < C; c += 8) {
      const uint4 q = *reinterpret_cast<const uint4*>(row + c);
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16(u[k]);
        acc = fmaf(f.x, sw[j * C + c + 2 * k], acc);
        acc = fmaf(f.y, sw[j * C + c + 2 * k + 1], acc);
      }
    }
  }
  out[idx] = tanhf(acc);
}

}  // namespace

extern "C" int kr_hifi_pack_mel(const float* mel, int time_major, void* out, int B, int T, int n_mels, int halo,
                                int c_phys, void* stream) {
  if (B <= 0 || T <= 0) { kr_set_error("kr_hifi_pack_mel: empty input"); return KR_ERR_ARG; }
  if (c_phys < n_mels || (c_phys & 1)) { kr_set_error("kr_hifi_pack_mel: c_phys must be even and >= n_mels"); return KR_ERR_ARG; }
  dim3 grid((T + 31) / 32, B);
  kr::launch(pack_mel_kernel, grid, 256, n_mels * 33 * sizeof(float), (cudaStream_t)stream, 
      mel, time_major, reinterpret_cast<bf16*>(out), B, T, n_mels, halo, c_phys);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_hifi_post_tanh(const void* x, const float* w, const float* bias, float* out, int B, long long L,
                                 int halo, int C, int c_phys, int taps, void* stream) {
  if (taps != 7 || (C & 7) || halo < 3) { kr_set_error("kr_hifi_post_tanh: needs taps == 7, C % 8 == 0, halo >= 3"); return KR_ERR_UNSUPPORTED; }
  const long long n = (long long)B * L;
  if (n <= 0) return KR_OK;
  const int threads = 256;
  kr::launch(post_tanh_kernel<7>, (unsigned)((n + threads - 1) / threads), threads, 7 * C * sizeof(float), (cudaStream_t)stream, 
      reinterpret_cast<const bf16*>(x), w, bias, out, B, L, halo, C, c_phys);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
