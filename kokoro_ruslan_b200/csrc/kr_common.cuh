// kr_common.cuh — sm_100a building blocks shared by every kernel in libkokoro_b200.so.
//
// Thin inline-PTX wrappers for the Blackwell primitives the hot path uses:
//   * mbarrier (init / expect_tx / try_wait with a watchdog),
//   * TMA tiled loads (cp.async.bulk.tensor) + host-side tensor-map encoding,
//   * tcgen05 (TMEM alloc/dealloc, UMMA smem/instruction descriptors, mma, commit, ld/st),
//   * warp/block reductions for the HBM-bound kernels.
// No CUTLASS/CuTe includes: the bit layouts below follow the PTX ISA tables for
// tcgen05 matrix/instruction descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define KR_OK 0
#define KR_ERR_ARG (-1)
#define KR_ERR_CUDA (-2)
#define KR_ERR_TMAP (-3)
#define KR_ERR_UNSUPPORTED (-4)

#define KR_CHECK_LAUNCH()                                   \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) { kr_set_error(cudaGetErrorString(e__)); return KR_ERR_CUDA; } \
    kr_count_launch();                                      \
  } while (0)

void kr_set_error(const char* msg);
void kr_count_launch();
int kr_pdl_enabled();   // programmatic dependent launch (always on, see kr_api.cu)

typedef __nv_bfloat16 bf16;

namespace kr {

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library starts with
//   pdl_launch_dependents();  ... prologue that touches no global memory ...  pdl_wait();
// and is launched with programmaticStreamSerialization, so the launch latency and the prologue
// (barrier init, TMEM allocation, tensor-map prefetch) of kernel N+1 overlap the tail of kernel N.
// griddepcontrol.wait returns only when the preceding grid has COMPLETED and flushed its memory, so
// ordering is exactly that of a normal stream; a step is ~540 short dependent kernels, which makes the
// per-launch bubble (~2 us) a first-order cost.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }

// Every kernel of the library asks for the MAXIMUM shared-memory carve-out, including the HBM kernels
// that use no shared memory at all: the tcgen05 GEMM / attention CTAs need ~200 KB, and an SM has to
// drain before it can change its L1/shared split, so alternating small-smem and large-smem kernels
// (which is what a training step is) would otherwise pay a reconfiguration bubble on every GEMM launch.
void kr_prefer_max_smem(const void* kernel);

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  kr_prefer_max_smem(reinterpret_cast<const void*>(kernel));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kr_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// generic helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` must hold >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// ---------------------------------------------------------------------------------------------
// Dropout / stochastic depth (reference nn.Dropout sites of model/transformers.py:108-111,396,482-487,
// 569-581, model/positional_encoding.py:74, model/model.py:525, model/variance_predictor.py:106 and
// drop_path :16-40).  Counter-based: the keep decision of element e at dropout site s in optimizer
// step t is a pure function of (seed, t, s, e), so the backward kernels REGENERATE the masks instead
// of storing them, and tests can export exactly the mask a fused kernel used (kr_drop_export_mask).
//   key(seed, t, s)   = splitmix64 finaliser -> two 32-bit words
//   hash(pair, key)   = two multiply-xorshift rounds keyed before each multiply; 32 bits serve the
//                       element pair (2*pair, 2*pair + 1) as two 16-bit lanes
//   keep(e)           = lane16 >= thr,   thr = round(p * 65536)      (P(keep) = 1 - thr / 65536)
//   byte-lane mode (attention-probability sites only, kr_attn.cu): one hash serves the 4 elements 4*quad .. 4*quad+3
//                       as 8-bit lanes, keep iff lane8 >= thr >> 8 with thr = round(p * 256) * 256
// A spec can carry a second independent mask (site_b / thr_b: consecutive reference dropouts) and a
// per-sample factor table (stochastic depth: 0 or 1 / (1 - p_path)), see kr_drop_spec.
// ---------------------------------------------------------------------------------------------
struct DropSpec {               // device-side copy of kr_drop_spec (include/kokoro_b200.h)
  const unsigned long long* state;   // {seed, step}; nullptr = dropout disabled
  uint32_t site_a, thr_a, site_b, thr_b;
  float scale;                       // factor of kept elements
  const float* row_scale;            // optional [n_samples]
  int rows_per_sample;
};
struct DropCtx { uint2 ka, kb; uint32_t thr_a, thr_b; float scale; };

__device__ __forceinline__ uint2 drop_key(const unsigned long long* state, uint32_t site) {
  unsigned long long z = state[0] + 0x9E3779B97F4A7C15ull * (state[1] + 1ull) + 0xD1B54A32D192ED03ull * (site + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return make_uint2(static_cast<uint32_t>(z), static_cast<uint32_t>(z >> 32));
}
__device__ __forceinline__ uint32_t drop_hash(uint32_t pair, uint2 k) {
  uint32_t x = pair ^ k.x;
  x *= 0x7feb352du; x ^= x >> 15;
  x ^= k.y;
  x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ DropCtx drop_ctx(const DropSpec& d) {
  DropCtx c;
  c.ka = drop_key(d.state, d.site_a);
  c.kb = d.thr_b ? drop_key(d.state, d.site_b) : make_uint2(0u, 0u);
  c.thr_a = d.thr_a; c.thr_b = d.thr_b; c.scale = d.scale;
  return c;
}
// factors (0 or scale) of elements 2*pair and 2*pair + 1
__device__ __forceinline__ void drop_pair(const DropCtx& c, uint32_t pair, float& f0, float& f1) {
  const uint32_t x = drop_hash(pair, c.ka);
  bool k0 = (x & 0xffffu) >= c.thr_a, k1 = (x >> 16) >= c.thr_a;
  if (c.thr_b) {
    const uint32_t y = drop_hash(pair, c.kb);
    k0 = k0 && (y & 0xffffu) >= c.thr_b;
    k1 = k1 && (y >> 16) >= c.thr_b;
  }
  f0 = k0 ? c.scale : 0.f;
  f1 = k1 ? c.scale : 0.f;
}
// factors of the four consecutive elements starting at e0 (e0 % 4 == 0)
__device__ __forceinline__ float4 drop_quad(const DropCtx& c, long long e0) {
  float4 f;
  const uint32_t pr = static_cast<uint32_t>(e0 >> 1);
  drop_pair(c, pr, f.x, f.y);
  drop_pair(c, pr + 1u, f.z, f.w);
  return f;
}
// factor of the single element e
__device__ __forceinline__ float drop_one(const DropCtx& c, long long e) {
  float f0, f1;
  drop_pair(c, static_cast<uint32_t>(e >> 1), f0, f1);
  return (e & 1) ? f1 : f0;
}
__device__ __forceinline__ float drop_row_scale(const DropSpec& d, int row) {
  return d.row_scale != nullptr ? d.row_scale[row / d.rows_per_sample] : 1.f;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  // generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait with a watchdog: a pipeline bug traps (-> cudaErrorLaunchFailure) instead of hanging the box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 1.9 GHz
      printf("kr: mbarrier watchdog fired (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrive on `bar` when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M lanes x K/2 packed-bf16x2 columns, K-major only) is read
// from tensor memory, so only B costs shared-memory bandwidth.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns, registers -> TMEM (thread i of the warp writes lane 32*(warp%4)+i).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
      " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns, registers -> TMEM.
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
//   [4,6) D format (1 = f32) | [7,10) A format (1 = bf16) | [10,13) B format (1 = bf16)
//   [15] A major (0 = K, 1 = MN) | [16] B major | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                       uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Shared-memory matrix descriptor, 128-byte swizzle.
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SW128)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// Same with an explicit swizzle layout: 2 = 128-byte swizzle (64-element bf16 K blocks), 4 = 64-byte swizzle
// (32-element K blocks: rows are 64 B, the pattern repeats every 8 rows = 512 B).
__device__ __forceinline__ uint64_t make_smem_desc_sw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                      uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives TMEM lane
// (32*(warp_id%4) + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace kr

// host: kr_drop_spec (include/kokoro_b200.h) -> kernel parameter; NULL / disabled -> state == nullptr
struct kr_drop_spec;
kr::DropSpec kr_drop_to_device(const kr_drop_spec* s);

// ---------------------------------------------------------------------------------------------
// host: tensor-map encoding through the driver entry point (no link-time libcuda dependency, so
// the library still dlopen()s on a CPU-only box for the symbol-export test).
// ---------------------------------------------------------------------------------------------
// 3-D bf16 tensor map: dims (inner, rows, batch), strides in ELEMENTS for rows / batch, box
// (box_inner, box_rows, 1), 128-byte swizzle (box_inner must be 64 bf16 = 128 B).
int kr_make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows,
                         uint64_t batch, uint64_t row_stride_elems, uint64_t batch_stride_elems,
                         uint32_t box_inner, uint32_t box_rows, int swizzle64 = 0);
// 4-D bf16 tensor map over a [batch, seq, heads, 64] token-major activation viewed as
// (d=64, head, seq, batch); strides in ELEMENTS; box (64, 1, box_seq, 1), 128-byte swizzle.
int kr_make_tmap_bf16_heads(CUtensorMap* out, const void* base, uint64_t heads, uint64_t seq,
                            uint64_t batch, uint64_t head_stride, uint64_t seq_stride,
                            uint64_t batch_stride, uint32_t box_seq);
