// kr_gemm.cu — tcgen05 bf16 GEMM for every Linear on the acoustic-model hot path
// (reference call sites: model/transformers.py:228,258-259,434 Q/K/V/out projections;
//  transformers.py:105-111 GLU feed-forward; model/model.py:519-531,561 mel in/out projections).
//
//   C[b][M,N] (+)= alpha * sum_k A[b][m,k] * B[b][n,k]  (+ bias[n]) (+ resid[m % resid_mod, n])
//
// One CTA computes one 128 x BLOCK_N tile: warp 0 = TMA producer, warp 1 = single-thread
// tcgen05.mma issuer (accumulator in TMEM), warps 2..5 = epilogue (tcgen05.ld -> registers ->
// global).  Operands are staged by TMA into 128B-swizzled shared-memory tiles through a
// KR_GEMM_STAGES-deep mbarrier ring.  A and B may each be K-major ([rows, K], K contiguous) or
// MN-major ([K, rows], rows contiguous): the second form is what the weight-gradient
// (dW = dY^T X) and data-gradient GEMMs read, so no transposed copies are materialised.
// Split-K (grid.z) with an fp32 vector-atomic epilogue serves the weight gradients, which are
// accumulated into the flat gradient buffer anyway.
#include "kr_common.cuh"
#include <stdio.h>

namespace {

using namespace kr;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

enum EpiMode { EPI_BF16 = 0, EPI_F32 = 1, EPI_ATOMIC_F32 = 2 };

struct GemmParams {
  int M, N, K;
  int batch, splits, kb_per_split;
  void* C;
  long long ldc, c_batch_stride;
  const float* bias;
  const float* resid;
  long long ldr, r_batch_stride;
  int resid_mod;
  float alpha;
};

template <int BLOCK_N>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N >= 128) ? 3 : 4;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
};

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
kr_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p) {
  using L = SmemLayout<BLOCK_N>;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BLOCK_N;
  const int bz = blockIdx.z / p.splits;
  const int split = blockIdx.z % p.splits;
  const int total_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, total_kb);
  const int num_kb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
        const int k = (kb0 + i) * BLOCK_K;
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < BLOCK_M / 64; ++j)
            tma_load_3d(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[s], m0 + j * 64, k, bz);
        } else {
          tma_load_3d(sa, &tmap_a, &full_bar[s], k, m0, bz);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_3d(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[s], n0 + j * 64, k, bz);
        } else {
          tma_load_3d(sb, &tmap_b, &full_bar[s], k, n0, bz);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      // K-major: 8-row groups 1024 B apart, +32 B per UMMA_K.  MN-major: 8-k groups 1024 B apart,
      // 64-wide MN chunks BLOCK_K*128 B apart, +2048 B per UMMA_K.
      constexpr uint32_t a_lbo = A_MN ? BLOCK_K * 128 : 16, b_lbo = B_MN ? BLOCK_K * 128 : 16;
      constexpr uint32_t a_kstep = A_MN ? 2048 : 32, b_kstep = B_MN ? 2048 : 32;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint64_t da = make_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
          const uint64_t db = make_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
          umma_bf16_ss(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 -> TMEM lane groups (warp % 4) =====
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int m = m0 + row;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const bool row_ok = (m < p.M) && (num_kb > 0);
    const long long c_off = (long long)bz * p.c_batch_stride + (long long)m * p.ldc;
    const float* rrow = nullptr;
    if (p.resid != nullptr && row_ok) {
      const int rm = p.resid_mod > 0 ? (m % p.resid_mod) : m;
      rrow = p.resid + (long long)bz * p.r_batch_stride + (long long)rm * p.ldr;
    }
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int nb = n0 + c * 32;
      if (row_ok && nb < p.N) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
      const bool full = (nb + 32 <= p.N);
      if (p.bias != nullptr && (EPI != EPI_ATOMIC_F32 || split == 0)) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (full || nb + i < p.N) v[i] += __ldg(p.bias + nb + i);
      }
      if (rrow != nullptr) {
        if (full) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(rrow + nb + 4 * i);
            v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
          }
        } else {
          for (int i = 0; i < 32; ++i)
            if (nb + i < p.N) v[i] += rrow[nb + i];
        }
      }
      if (EPI == EPI_BF16) {
        bf16* crow = reinterpret_cast<bf16*>(p.C) + c_off + nb;
        if (full) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 q;
            q.x = pack_bf16(v[8 * i], v[8 * i + 1]);
            q.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
            q.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
            q.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
            *reinterpret_cast<uint4*>(crow + 8 * i) = q;
          }
        } else {
          for (int i = 0; i < 32; ++i)
            if (nb + i < p.N) crow[i] = __float2bfloat16(v[i]);
        }
      } else if (EPI == EPI_F32) {
        float* crow = reinterpret_cast<float*>(p.C) + c_off + nb;
        if (full) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(crow + 4 * i) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          for (int i = 0; i < 32; ++i)
            if (nb + i < p.N) crow[i] = v[i];
        }
      } else {
        float* crow = reinterpret_cast<float*>(p.C) + c_off + nb;
        if (full) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            atomicAdd(reinterpret_cast<float4*>(crow + 4 * i),
                      make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
        } else {
          for (int i = 0; i < 32; ++i)
            if (nb + i < p.N) atomicAdd(crow + i, v[i]);
        }
      }
      }
      __syncwarp();  // reconverge before the next warp-aligned tcgen05.ld
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  using L = SmemLayout<BLOCK_N>;
  auto kern = kr_gemm_kernel<BLOCK_N, A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) { kr_set_error(cudaGetErrorString(e)); return KR_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid((p.M + BLOCK_M - 1) / BLOCK_M, (p.N + BLOCK_N - 1) / BLOCK_N, p.batch * p.splits);
  kern<<<grid, GEMM_THREADS, L::TOTAL, st>>>(ta, tb, p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN>
int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                 cudaStream_t st) {
  switch (epi) {
    case EPI_BF16: return launch_gemm<BLOCK_N, A_MN, B_MN, EPI_BF16>(ta, tb, p, st);
    case EPI_F32: return launch_gemm<BLOCK_N, A_MN, B_MN, EPI_F32>(ta, tb, p, st);
    case EPI_ATOMIC_F32: return launch_gemm<BLOCK_N, A_MN, B_MN, EPI_ATOMIC_F32>(ta, tb, p, st);
  }
  kr_set_error("kr_gemm_bf16: bad epilogue mode");
  return KR_ERR_ARG;
}

template <int BLOCK_N>
int dispatch_major(int a_mn, int b_mn, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                   const GemmParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn) return dispatch_epi<BLOCK_N, false, false>(epi, ta, tb, p, st);
  if (!a_mn && b_mn) return dispatch_epi<BLOCK_N, false, true>(epi, ta, tb, p, st);
  if (a_mn && !b_mn) return dispatch_epi<BLOCK_N, true, false>(epi, ta, tb, p, st);
  return dispatch_epi<BLOCK_N, true, true>(epi, ta, tb, p, st);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int kr_make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows,
                         uint64_t batch, uint64_t row_stride_elems, uint64_t batch_stride_elems,
                         uint32_t box_inner, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) { kr_set_error("cuTensorMapEncodeTiled entry point unavailable"); return KR_ERR_TMAP; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_elems & 7) || (batch_stride_elems & 7)) {
    kr_set_error("tensor map: base / strides must be 16-byte aligned");
    return KR_ERR_ARG;
  }
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg),
             "cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu batch=%llu ld=%llu bs=%llu",
             (int)r, (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)batch,
             (unsigned long long)row_stride_elems, (unsigned long long)batch_stride_elems);
    kr_set_error(msg);
    return KR_ERR_TMAP;
  }
  return KR_OK;
}

int kr_make_tmap_bf16_heads(CUtensorMap* out, const void* base, uint64_t heads, uint64_t seq,
                            uint64_t batch, uint64_t head_stride, uint64_t seq_stride,
                            uint64_t batch_stride, uint32_t box_seq) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) { kr_set_error("cuTensorMapEncodeTiled entry point unavailable"); return KR_ERR_TMAP; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (head_stride & 7) || (seq_stride & 7) || (batch_stride & 7)) {
    kr_set_error("tensor map (heads): base / strides must be 16-byte aligned");
    return KR_ERR_ARG;
  }
  cuuint64_t dims[4] = {64, heads, seq, batch};
  cuuint64_t strides[3] = {head_stride * 2, seq_stride * 2, batch_stride * 2};
  cuuint32_t box[4] = {64, 1, box_seq, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(heads) failed (%d): heads=%llu seq=%llu batch=%llu",
             (int)r, (unsigned long long)heads, (unsigned long long)seq, (unsigned long long)batch);
    kr_set_error(msg);
    return KR_ERR_TMAP;
  }
  return KR_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int kr_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int batch,
                            long long lda, long long ldb, long long ldc, long long stride_a,
                            long long stride_b, long long stride_c, int a_mn_major, int b_mn_major,
                            int epi_mode, const float* bias, const float* resid, long long ldr,
                            long long stride_r, int resid_mod, float alpha, int splits,
                            void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) { kr_set_error("kr_gemm_bf16: empty problem"); return KR_ERR_ARG; }
  // K needs no alignment: TMA zero-fills the out-of-bounds tail of the last K block (row strides
  // are validated when the tensor maps are encoded).
  const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  if (splits > 1 && epi_mode != EPI_ATOMIC_F32) { kr_set_error("kr_gemm_bf16: split-K needs the atomic epilogue"); return KR_ERR_ARG; }
  int kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per_split - 1) / kb_per_split;  // no empty splits

  const int block_n = (N > 64) ? 128 : 64;
  const uint64_t bstride_a = batch > 1 ? (uint64_t)stride_a : (uint64_t)lda * (a_mn_major ? K : M);
  const uint64_t bstride_b = batch > 1 ? (uint64_t)stride_b : (uint64_t)ldb * (b_mn_major ? K : N);
  CUtensorMap ta, tb;
  int rc;
  if (a_mn_major) rc = kr_make_tmap_bf16_3d(&ta, A, M, K, batch, lda, bstride_a, 64, BLOCK_K);
  else            rc = kr_make_tmap_bf16_3d(&ta, A, K, M, batch, lda, bstride_a, BLOCK_K, BLOCK_M);
  if (rc != KR_OK) return rc;
  if (b_mn_major) rc = kr_make_tmap_bf16_3d(&tb, B, N, K, batch, ldb, bstride_b, 64, BLOCK_K);
  else            rc = kr_make_tmap_bf16_3d(&tb, B, K, N, batch, ldb, bstride_b, BLOCK_K, block_n);
  if (rc != KR_OK) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.splits = splits; p.kb_per_split = kb_per_split;
  p.C = C; p.ldc = ldc; p.c_batch_stride = stride_c;
  p.bias = bias; p.resid = resid; p.ldr = ldr; p.r_batch_stride = stride_r; p.resid_mod = resid_mod;
  p.alpha = alpha;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (block_n == 128) return dispatch_major<128>(a_mn_major, b_mn_major, epi_mode, ta, tb, p, st);
  return dispatch_major<64>(a_mn_major, b_mn_major, epi_mode, ta, tb, p, st);
}
