// kr_gemm.cu — persistent tcgen05 bf16 GEMM / implicit-GEMM conv1d for every dense contraction on
// the hot path.
//   acoustic model: model/transformers.py:228,258-259,434 (Q/K/V/out projections), :105-111 (GLU FFN),
//                   model/model.py:519-531,561 (mel in/out projections), model/variance_predictor.py:46
//                   (k=3 Conv1d) and all of their weight- / data-gradient GEMMs;
//   vocoder:        inference/hifigan_vocoder.py:31-133 (dilated Conv1d / polyphase ConvTranspose1d).
//
//   v      = alpha * sum_k A[b][m,k] * B[b][n,k] + bias[n] + resid[b][m % resid_mod, n]
//   v      = v * beta + resid2[b][m, n]        (beta is applied with or without resid2)
//   C [b][m,n] = v                      (bf16 | fp32 | fp32 atomic-add | not stored)
//   C2[b][m,n] = bf16(leaky_relu(v, act_slope))          (optional second output)
//
// Design (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: A/B tiles -> 128B-swizzled smem ring (STAGES deep, runs ahead across
//               tile boundaries);
//   warp 1      single-thread tcgen05.mma issuer; the fp32 accumulator lives in TMEM and is
//               DOUBLE-BUFFERED (2 x BLOCK_N columns) so the MMAs of tile i+1 overlap the epilogue
//               of tile i;
//   warps 2..5  epilogue: tcgen05.ld -> registers -> fused bias / residual / activation -> global.
// Operands may be K-major ([rows, K]) or MN-major ([K, rows]); the latter is what the weight- and
// data-gradient GEMMs read in place.  Split-K uses the fp32 atomic epilogue.  In conv mode the A
// operand is a channels-last activation [batch, rows(+zero halos), C_in]: K-block kb = (tap, 64-ch
// chunk) is fetched at row offset tap*dilation — the im2col matrix is never materialised.
#include "kr_common.cuh"
#include "kokoro_b200.h"
#include <stdio.h>
#include <stdlib.h>

namespace {

using namespace kr;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;                       // two warps per TMEM lane group, alternating column chunks
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;   // producer warp + MMA warp + epilogue warps
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;         // per-warp 32x32 fp32 transpose buffer

enum CMode { C_BF16 = 0, C_F32 = 1, C_ATOMIC_F32 = 2, C_NONE = 3 };

struct GemmParams {
  int M, N, K;
  int batch, splits, kb_per_split, total_kb;
  int m_tiles, n_tiles, total_tiles;
  int conv_taps, conv_dil, conv_row0, conv_cin_blocks;
  int conv_slab, slab_rows;   // slab mode: one [slab_rows x 64ch] load per (tile, chunk); taps = row offsets into it
  // slab-stream mode (conv whose weights do NOT fit in smem): the activation slab of a tile is still fetched once per
  // 64-channel block, the weight K blocks stream through the ring, and every weight block feeds `msub` 128-row sub-tiles
  // (msub accumulators in TMEM) — the L2 -> SM traffic per MMA is what bounds these convs
  int slab_stream, msub, slab_bufs, slab_buf_bytes, slab_box_rows, tmem_cols;
  // slab-stream with swapped operands (BLOCK_N = 128, msub = 2): D^T[128 output channels x 256 rows] = W_blk . slab^T — ONE
  // N = 256 MMA per K step instead of two N = 128 ones.  A 1-CTA N = 128 MMA re-reads 8 KB of operands per 64 clocks, the
  // whole shared-memory bandwidth (measured cap 0.85 PFLOP/s, also at 8192^3 with BLOCK_N = 128); N = 256 reads 12 KB per
  // 128 clocks.  The accumulator then holds channels on the TMEM lanes and rows in the columns; the epilogue's staging
  // tile transposes it back, everything after the staging tile is unchanged.
  int swap;
  void* C; long long ldc, c_batch_stride; int c_mode;
  bf16* C2; long long ldc2, c2_batch_stride; float act_slope;
  const float* bias;
  const void* resid; int resid_bf16; long long ldr, r_batch_stride; int resid_mod;
  const void* resid2; int resid2_bf16; long long ldr2, r2_batch_stride;
  float alpha, beta;
  int stages, stage_bytes, b_resident, b_res_bytes;   // smem ring geometry (host-chosen)
  int b_shared;   // B has no batch dimension (weights shared by every batch item)
  DropSpec drop;  // dropout / stochastic depth on v = alpha*acc + bias, before the residual adds (state == nullptr: off)
};

// Dynamic shared memory map (offsets from the 1024-aligned base):
//   [0, 1024)                      mbarriers + TMEM slot
//   [1024, 1024 + 32 KB)           epilogue transpose buffers (one 32x32 fp32 tile per epilogue warp)
//   [.., + b_res_bytes)            RESIDENT B operand (all K blocks) when p.b_resident, else empty
//   [.., + stages * stage_bytes)   TMA ring: A tile (+ B tile unless resident) per stage
constexpr int MAX_STAGES = 12;
constexpr int BAR_BYTES = 1024;
constexpr int RING_OFFSET0 = BAR_BYTES + EPI_WARPS * EPI_STAGE_BYTES;   // 33 KB, 1024-aligned
constexpr int SMEM_TOTAL = 227 * 1024 - 1024;                            // leave slack for the base alignment

template <int BLOCK_N, int BK = BLOCK_K>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BK * 2;
  static constexpr int B_BYTES = BLOCK_N * BK * 2;
  static constexpr int TMEM_COLS = BLOCK_N <= 64 ? 128 : (BLOCK_N <= 128 ? 256 : 512);  // power of two >= 2*BLOCK_N
};

struct TileCoord { int m_blk, n_blk, bz, split; };
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
  TileCoord t;
  t.m_blk = tile % p.m_tiles;
  int r = tile / p.m_tiles;
  t.n_blk = r % p.n_tiles;
  r /= p.n_tiles;
  t.split = r % p.splits;
  t.bz = r / p.splits;
  return t;
}

__device__ __forceinline__ float4 ld4any(const void* base, int is_bf16, long long off) {
  if (is_bf16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(base) + off);
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y);
    return make_float4(f0.x, f0.y, f1.x, f1.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
}

__device__ __forceinline__ float4 add4(float4 v, const void* base, int is_bf16, long long off) {
  if (is_bf16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(base) + off);
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y);
    v.x += f0.x; v.y += f0.y; v.z += f1.x; v.w += f1.y;
  } else {
    const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    v.x += f.x; v.y += f.y; v.z += f.z; v.w += f.w;
  }
  return v;
}

// Rare path (N or a leading dimension not a multiple of 4, e.g. N = 80 with ld 80 is fine but N = 1
// or odd strides are not): one element at a time, deliberately not inlined / unrolled so that it does
// not bloat the instruction footprint of the hot epilogue.
__device__ __noinline__ void epi_chunk_scalar(const GemmParams& p, const float* stg, int bz, int mw0, int nb,
                                              int lane, bool use_bias, bool use_resid) {
#pragma unroll 1
  for (int e = lane; e < 32 * 32; e += 32) {
    const int rr = e >> 5, cc = e & 31;
    const int m = mw0 + rr, n = nb + cc;
    if (m >= p.M || n >= p.N) continue;
    float v = stg[rr * 32 + ((((cc >> 2) ^ (rr & 7)) << 2) | (cc & 3))] * p.alpha;
    if (use_bias) v += __ldg(p.bias + n);
    if (use_resid) {
      const int rm = p.resid_mod > 0 ? (m % p.resid_mod) : m;
      const long long off = (long long)bz * p.r_batch_stride + (long long)rm * p.ldr + n;
      v += p.resid_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p.resid)[off])
                        : reinterpret_cast<const float*>(p.resid)[off];
    }
    v *= p.beta;
    if (p.resid2 != nullptr) {
      const long long off = (long long)bz * p.r2_batch_stride + (long long)m * p.ldr2 + n;
      v += p.resid2_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p.resid2)[off])
                         : reinterpret_cast<const float*>(p.resid2)[off];
    }
    const long long c_off = (long long)bz * p.c_batch_stride + (long long)m * p.ldc + n;
    if (p.c_mode == C_BF16) reinterpret_cast<bf16*>(p.C)[c_off] = __float2bfloat16(v);
    else if (p.c_mode == C_F32) reinterpret_cast<float*>(p.C)[c_off] = v;
    else if (p.c_mode == C_ATOMIC_F32) atomicAdd(reinterpret_cast<float*>(p.C) + c_off, v);
    if (p.C2 != nullptr)
      p.C2[(long long)bz * p.c2_batch_stride + (long long)m * p.ldc2 + n] =
          __float2bfloat16(v > 0.f ? v : v * p.act_slope);
  }
}

// BK = 64: 128-byte swizzled K blocks (everything).  BK = 32: 64-byte swizzled K blocks, K-major operands only —
// the 32-channel HiFi-GAN stage, whose K blocks would otherwise be half zero padding.
// DROP = the dropout epilogue is compiled in (K-major operands only): a template parameter, not a runtime branch,
// because the epilogue's instruction footprint matters (an earlier 240 KB kernel thrashed the instruction cache, and
// the runtime-branch version of the dropout code cost the HiFi-GAN convs 8 % although they never take it).
template <int BLOCK_N, bool A_MN, bool B_MN, int BK = BLOCK_K, bool DROP = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
kr_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p) {
  static_assert(BK == 64 || (BK == 32 && !A_MN && !B_MN), "32-wide K blocks are K-major only");
  using L = SmemLayout<BLOCK_N, BK>;
  constexpr uint32_t ROWB = BK * 2;                 // bytes per K-major smem row
  constexpr uint32_t SBO = 8 * ROWB;                // 8-row swizzle atom
  constexpr uint32_t LAYOUT = BK == 64 ? 2u : 4u;   // SWIZZLE_128B / SWIZZLE_64B
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]
  uint64_t* bres_bar = tmem_empty_bar + 2;
  uint64_t* slab_full_bar = bres_bar + 1;             // [4] slab-stream mode
  uint64_t* slab_empty_bar = slab_full_bar + 4;       // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slab_empty_bar + 4);
  uint8_t* b_res = smem + RING_OFFSET0;
  uint8_t* ring = b_res + p.b_res_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();   // the next kernel may start its own prologue now

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(bres_bar, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&slab_full_bar[s], 1); mbar_init(&slab_empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      // one arrive per epilogue warp that works on the buffer: all 8, or — narrow outputs, see the epilogue — the 4 of
      // the warp set that owns it
      mbar_init(&tmem_empty_bar[s], p.N <= 32 ? EPI_WARPS / 2 : EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                // everything below reads / writes global memory of the preceding kernels

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int ring_s = 0; uint32_t ring_ph = 0;  // ring position (stage, phase) — no runtime division
      if (p.b_resident) {     // the whole [N, K] weight operand stays in smem for the life of the CTA
        mbar_arrive_expect_tx(bres_bar, (uint32_t)p.b_res_bytes);
        for (int kb = 0; kb < p.total_kb; ++kb) {
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_3d(b_res + kb * L::B_BYTES + j * (BLOCK_K * 128), &tmap_b, bres_bar, j * 64, kb * BK, 0);
          } else {
            tma_load_3d(b_res + kb * L::B_BYTES, &tmap_b, bres_bar, kb * BK, 0, 0);
          }
        }
      }
      if (p.slab_stream) {
        // units u = (local tile, 64-channel block); the slab of unit u + 1 is requested BEFORE the weight blocks of unit u
        // (its buffer was released by unit u + 1 - slab_bufs, two units behind the MMA front when slab_bufs == 3)
        const int cinb = p.conv_cin_blocks;
        const int n_local = blockIdx.x < p.total_tiles ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int n_units = n_local * cinb;
        uint8_t* ring_b = ring + p.slab_bufs * p.slab_buf_bytes;
        auto slab_issue = [&](int u) {
          if (u >= n_units) return;
          const int lt = u / cinb, cb = u - lt * cinb;
          const TileCoord tc = decode_tile(p, blockIdx.x + lt * gridDim.x);
          const int sbuf = u % p.slab_bufs;
          mbar_wait(&slab_empty_bar[sbuf], (((uint32_t)(u / p.slab_bufs)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&slab_full_bar[sbuf], (uint32_t)p.slab_buf_bytes);
          const int row = p.conv_row0 + tc.m_blk * BLOCK_M * p.msub;
          for (int r = 0; r < p.slab_rows; r += p.slab_box_rows)
            tma_load_3d(ring + sbuf * p.slab_buf_bytes + r * ROWB, &tmap_a, &slab_full_bar[sbuf], cb * BK, row + r, tc.bz);
        };
        slab_issue(0);
        for (int u = 0; u < n_units; ++u) {
          slab_issue(u + 1);
          const int lt = u / cinb, cb = u - lt * cinb;
          const TileCoord tc = decode_tile(p, blockIdx.x + lt * gridDim.x);
          for (int tap = 0; tap < p.conv_taps; ++tap) {
            const int s = ring_s;
            const uint32_t ph = ring_ph;
            if (++ring_s == STAGES) { ring_s = 0; ring_ph ^= 1u; }
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)L::B_BYTES);
            tma_load_3d(ring_b + s * L::B_BYTES, &tmap_b, &full_bar[s], (tap * cinb + cb) * BK, tc.n_blk * BLOCK_N,
                        p.b_shared ? 0 : tc.bz);
          }
        }
      } else
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int m0 = tc.m_blk * BLOCK_M, n0 = tc.n_blk * BLOCK_N;
        const int kb0 = tc.split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.total_kb);
        const int bzb = p.b_shared ? 0 : tc.bz;
        if (p.conv_slab) {
          // implicit GEMM proper: every activation row of the tile (+ the (taps-1)*dil halo rows) is fetched
          // ONCE per 64-channel chunk; the taps are row-shifted views of the same smem slab.
          const int s = ring_s;
          const uint32_t ph = ring_ph;
          if (++ring_s == STAGES) { ring_s = 0; ring_ph ^= 1u; }
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = ring + s * p.stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.stage_bytes);
          for (int cb = 0; cb < p.conv_cin_blocks; ++cb)
            tma_load_3d(sa + cb * p.slab_rows * ROWB, &tmap_a, &full_bar[s], cb * BK, p.conv_row0 + m0, tc.bz);
          continue;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          const int s = ring_s;
          const uint32_t ph = ring_ph;
          if (++ring_s == STAGES) { ring_s = 0; ring_ph ^= 1u; }
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = ring + s * p.stage_bytes;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.stage_bytes);
          const int k = kb * BK;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_3d(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[s], m0 + j * 64, k, tc.bz);
          } else if (p.conv_taps > 0) {
            const int tap = kb / p.conv_cin_blocks, cb = kb - tap * p.conv_cin_blocks;
            tma_load_3d(sa, &tmap_a, &full_bar[s], cb * BK, p.conv_row0 + m0 + tap * p.conv_dil, tc.bz);
          } else {
            tma_load_3d(sa, &tmap_a, &full_bar[s], k, m0, tc.bz);
          }
          if (p.b_resident) {
            // nothing: B is already in smem
          } else if (B_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_3d(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[s], n0 + j * 64, k, bzb);
          } else {
            tma_load_3d(sb, &tmap_b, &full_bar[s], k, n0, bzb);
          }
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
      int ring_s = 0, lt = 0;  // ring stage, local tile counter
      uint32_t ring_ph = 0;
      if (p.b_resident) { mbar_wait(bres_bar, 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const TileCoord tc = decode_tile(p, tile);
        const int kb0 = tc.split * p.kb_per_split;
        const int num_kb = min(kb0 + p.kb_per_split, p.total_kb) - kb0;
        const int buf = lt & 1;
        const uint32_t use = (uint32_t)(lt >> 1);
        mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * p.msub * BLOCK_N;
        if (p.slab_stream) {
          const uint8_t* ring_b = ring + p.slab_bufs * p.slab_buf_bytes;
          for (int cb = 0; cb < p.conv_cin_blocks; ++cb) {
            const int u = lt * p.conv_cin_blocks + cb;
            const int sbuf = u % p.slab_bufs;
            mbar_wait(&slab_full_bar[sbuf], ((uint32_t)(u / p.slab_bufs)) & 1u);
            tc_fence_after();
            const uint32_t slab = smem_u32(ring + sbuf * p.slab_buf_bytes);
            for (int tap = 0; tap < p.conv_taps; ++tap) {
              const int s = ring_s;
              const uint32_t ph = ring_ph;
              if (++ring_s == STAGES) { ring_s = 0; ring_ph ^= 1u; }
              mbar_wait(&full_bar[s], ph);
              tc_fence_after();
              const uint32_t sb = smem_u32(ring_b + s * L::B_BYTES);
              const uint32_t acc = (cb > 0 || tap > 0) ? 1u : 0u;
              if (p.swap) {
                constexpr uint32_t idesc_t = make_idesc_bf16(128, 256, 0, 0);
                const uint32_t sx = slab + tap * p.conv_dil * ROWB;      // 256 consecutive slab rows = the N operand
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                  umma_bf16_ss(tmem_d, make_smem_desc_sw(sb + k * 32, 16, SBO, LAYOUT), make_smem_desc_sw(sx + k * 32, 16, SBO, LAYOUT),
                               idesc_t, (acc | (k > 0)) ? 1u : 0u);
              } else
              for (int sub = 0; sub < p.msub; ++sub) {
                const uint32_t sa = slab + (tap * p.conv_dil + sub * BLOCK_M) * ROWB;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                  umma_bf16_ss(tmem_d + sub * BLOCK_N, make_smem_desc_sw(sa + k * 32, 16, SBO, LAYOUT),
                               make_smem_desc_sw(sb + k * b_kstep, b_lbo, SBO, LAYOUT), idesc, (acc | (k > 0)) ? 1u : 0u);
              }
              umma_commit(&empty_bar[s]);
            }
            umma_commit(&slab_empty_bar[sbuf]);
          }
          umma_commit(&tmem_full_bar[buf]);
          continue;
        }
        if (p.conv_slab) {
          const int s = ring_s;
          const uint32_t ph = ring_ph;
          if (++ring_s == STAGES) { ring_s = 0; ring_ph ^= 1u; }
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t slab = smem_u32(ring + s * p.stage_bytes);
          bool first = true;
          for (int tap = 0; tap < p.conv_taps; ++tap) {
            for (int cb = 0; cb < p.conv_cin_blocks; ++cb) {
              // rows [tap*dil, tap*dil + 128) of chunk cb's slab.  The start address is row-shifted (not 1024-byte
              // aligned); the 128B swizzle is a function of the absolute smem address on sm_100a, so TMA's
              // placement and the UMMA read agree with the descriptor's base-offset field left at 0 (measured:
              // bit-identical to re-fetching every tap; setting base_offset = (addr >> 7) & 7 is WRONG here)
              const uint32_t sa = slab + cb * p.slab_rows * ROWB + tap * p.conv_dil * ROWB;
              const uint32_t sb = smem_u32(b_res + (tap * p.conv_cin_blocks + cb) * L::B_BYTES);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t da = make_smem_desc_sw(sa + k * 32, 16, SBO, LAYOUT);
                const uint64_t db = make_smem_desc_sw(sb + k * b_kstep, b_lbo, SBO, LAYOUT);
                umma_bf16_ss(tmem_d, da, db, idesc, first ? 0u : 1u);
                first = false;
              }
            }
          }
          umma_commit(&empty_bar[s]);
          umma_commit(&tmem_full_bar[buf]);
          continue;
        }
        for (int i = 0; i < num_kb; ++i) {
          const int s = ring_s;
          const uint32_t ph = ring_ph;
          if (++ring_s == STAGES) { ring_s = 0; ring_ph ^= 1u; }
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + s * p.stage_bytes);
          const uint32_t sb = p.b_resident ? smem_u32(b_res + (kb0 + i) * L::B_BYTES) : sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw(sa + k * a_kstep, a_lbo, A_MN ? 1024u : SBO, LAYOUT);
            const uint64_t db = make_smem_desc_sw(sb + k * b_kstep, b_lbo, B_MN ? 1024u : SBO, LAYOUT);
            umma_bf16_ss(tmem_d, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
        }
        umma_commit(&tmem_full_bar[buf]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane group = warp % 4, two warps per group split the column chunks =====
    // Accumulator rows live one per thread (TMEM lane); every chunk of 32 columns is transposed through a
    // per-warp swizzled smem buffer so that ALL global traffic (residual loads, output stores) is issued
    // as 4 rows x 128 contiguous bytes per warp instruction.
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stg = reinterpret_cast<float*>(smem + BAR_BYTES + (warp - 2) * EPI_STAGE_BYTES);
    const int crow = lane >> 3;          // coalesced layout: this lane handles row 4*i + crow ...
    const int ccol = (lane & 7) * 4;     // ... columns [ccol, ccol + 4) of the chunk
    const bool vec_ok = ((p.N & 3) == 0) && ((p.ldc & 3) == 0) && ((p.ldr & 3) == 0) && ((p.ldr2 & 3) == 0) &&
                        ((p.ldc2 & 3) == 0);
    // Narrow outputs (N <= 32: the 32-channel HiFi-GAN stage): a tile has ONE 32-column chunk, which would leave the
    // second warp of every lane group idle while the first runs a latency chain (TMEM load, smem transpose, global
    // stores) per tile.  The two warp sets then alternate TILES instead of column chunks: set h owns accumulator
    // buffer h (= tiles with lt % 2 == h), so two tiles' epilogues are in flight per CTA.
    const bool tile_alt = p.N <= 32;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      if (tile_alt && (lt & 1) != half) continue;
      const TileCoord tc = decode_tile(p, tile);
      for (int sub = 0; sub < p.msub; ++sub) {        // 128-row sub-tiles of the tile (1 except in slab-stream mode)
      const int mw0_t = (tc.m_blk * p.msub + sub) * BLOCK_M + lg * 32;   // first row of this warp's 32-row slab
      const int n0 = tc.n_blk * BLOCK_N;
      const int buf = lt & 1;
      const uint32_t use = (uint32_t)(lt >> 1);
      if (sub == 0) { mbar_wait(&tmem_full_bar[buf], use & 1); tc_fence_after(); }
      const bool first_split = (tc.split == 0);
      const bool use_bias = p.bias != nullptr && first_split;
      const bool use_resid = p.resid != nullptr && first_split;
      // dropout epilogue (attention out-projection): per-row stochastic-depth factors of this lane's 8 rows
      constexpr bool use_drop = DROP;
      DropCtx dc{};
      float rowf[8];
      if (DROP) {
        dc = drop_ctx(p.drop);
#pragma unroll
        for (int i = 0; i < 8; ++i) rowf[i] = drop_row_scale(p.drop, min(mw0_t + (lane >> 3) + 4 * i, p.M - 1));
      }
#pragma unroll 1
      for (int c = tile_alt ? 0 : half; c < BLOCK_N / 32; c += 2) {
        // swapped slab-stream tiles: TMEM lanes = output channels n0 + 32 lg .., column chunk (sub, c) = 32 output rows
        const int nb = p.swap ? n0 + lg * 32 : n0 + c * 32;
        const int mw0 = p.swap ? tc.m_blk * (2 * BLOCK_M) + (sub * (BLOCK_N / 32) + c) * 32 : mw0_t;
        if (nb >= p.N) break;                          // warp-uniform
        // The epilogue is ISSUE-bound for the short-K GEMMs / convs of this model (it was ~870 warp instructions
        // per 32x32 chunk = 3500 issue cycles per 128x128 tile vs ~2000 MMA cycles), so everything below keeps
        // the per-row work to a handful of instructions: row pointers advance by a constant stride, optional
        // operands are handled by warp-uniform branches OUTSIDE the row loops, interior tiles skip predicates.
        const int n = nb + ccol;
        const bool col_ok = n < p.N;          // N % 4 == 0 on this path: a 4-group is all in or all out
        const int m_first = mw0 + crow;       // this lane's rows: m_first + 4*i
        const bool interior = (mw0 + 32 <= p.M) && (nb + 32 <= p.N);      // warp-uniform
        uint32_t okm = 0;                     // bit i: row i of this lane is inside the matrix
        if (col_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) okm |= (m_first + 4 * i < p.M) ? (1u << i) : 0u;
        }
        // (1) residual operands are requested FIRST, in the coalesced layout, so that their HBM / L2 latency
        //     overlaps the TMEM load and the smem transpose
        float4 r1[8], r2[8];
        if (vec_ok && use_resid) {
          if (p.resid_mod > 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int m = m_first + 4 * i;
              r1[i] = (okm >> i) & 1u ? ld4any(p.resid, p.resid_bf16, (long long)tc.bz * p.r_batch_stride +
                                                                      (long long)(m % p.resid_mod) * p.ldr + n)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          } else {
            const long long off0 = (long long)tc.bz * p.r_batch_stride + (long long)m_first * p.ldr + n;
            const long long step = 4 * p.ldr;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              r1[i] = (interior || ((okm >> i) & 1u)) ? ld4any(p.resid, p.resid_bf16, off0 + i * step)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (vec_ok && p.resid2 != nullptr) {
          const long long off0 = (long long)tc.bz * p.r2_batch_stride + (long long)m_first * p.ldr2 + n;
          const long long step = 4 * p.ldr2;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            r2[i] = (interior || ((okm >> i) & 1u)) ? ld4any(p.resid2, p.resid2_bf16, off0 + i * step)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // (2) accumulator chunk: TMEM -> registers (row layout) -> smem (16-byte slots XOR-swizzled by row:
        //     conflict-free both ways)
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (buf * p.msub + sub) * BLOCK_N + (static_cast<uint32_t>(lg * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        if (p.swap) {      // register t of lane l = (row t, channel l) of the 32 x 32 block: transposed, conflict-free both ways
#pragma unroll
          for (int t = 0; t < 32; ++t) stg[t * 32 + lane] = __uint_as_float(r[t]);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
                make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
        __syncwarp();
        if (vec_ok) {
          // (3) back in the coalesced layout: v[i] = row m_first + 4*i, columns [n, n+4)
          float4 v[8];
          const float* sp = stg + crow * 32;
          const int xs = lane & 7;
          if (p.swap) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(sp + i * 128 + (xs << 2));
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)   // row 4*i + crow: (row & 7) = ((i & 1) << 2) | crow
              v[i] = *reinterpret_cast<const float4*>(sp + i * 128 + ((xs ^ (((i & 1) << 2) | crow)) << 2));
          }
          if (p.alpha != 1.f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i].x *= p.alpha; v[i].y *= p.alpha; v[i].z *= p.alpha; v[i].w *= p.alpha; }
          }
          if (use_bias) {
            const float4 b4 = col_ok ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i].x += b4.x; v[i].y += b4.y; v[i].z += b4.z; v[i].w += b4.w; }
          }
          if (use_drop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 f = drop_quad(dc, (long long)(m_first + 4 * i) * p.N + n);
              const float rf = rowf[i];
              v[i].x *= f.x * rf; v[i].y *= f.y * rf; v[i].z *= f.z * rf; v[i].w *= f.w * rf;
            }
          }
          if (use_resid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i].x += r1[i].x; v[i].y += r1[i].y; v[i].z += r1[i].z; v[i].w += r1[i].w; }
          }
          if (p.beta != 1.f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i].x *= p.beta; v[i].y *= p.beta; v[i].z *= p.beta; v[i].w *= p.beta; }
          }
          if (p.resid2 != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i].x += r2[i].x; v[i].y += r2[i].y; v[i].z += r2[i].z; v[i].w += r2[i].w; }
          }
          const uint32_t okw = interior ? 0xffu : okm;
          if (p.c_mode == C_BF16) {
            bf16* cp = reinterpret_cast<bf16*>(p.C) + (long long)tc.bz * p.c_batch_stride + (long long)m_first * p.ldc + n;
            const long long step = 4 * p.ldc;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if ((okw >> i) & 1u)
                *reinterpret_cast<uint2*>(cp + i * step) = make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
          } else if (p.c_mode == C_F32) {
            float* cp = reinterpret_cast<float*>(p.C) + (long long)tc.bz * p.c_batch_stride + (long long)m_first * p.ldc + n;
            const long long step = 4 * p.ldc;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if ((okw >> i) & 1u) *reinterpret_cast<float4*>(cp + i * step) = v[i];
          } else if (p.c_mode == C_ATOMIC_F32) {
            float* cp = reinterpret_cast<float*>(p.C) + (long long)tc.bz * p.c_batch_stride + (long long)m_first * p.ldc + n;
            const long long step = 4 * p.ldc;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if ((okw >> i) & 1u) atomicAdd(reinterpret_cast<float4*>(cp + i * step), v[i]);
          }
          if (p.C2 != nullptr) {
            const float sl = p.act_slope;
            bf16* cp = p.C2 + (long long)tc.bz * p.c2_batch_stride + (long long)m_first * p.ldc2 + n;
            const long long step = 4 * p.ldc2;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a = v[i];
              if ((okw >> i) & 1u)
                *reinterpret_cast<uint2*>(cp + i * step) =
                    make_uint2(pack_bf16(fmaxf(a.x, a.x * sl), fmaxf(a.y, a.y * sl)),
                               pack_bf16(fmaxf(a.z, a.z * sl), fmaxf(a.w, a.w * sl)));
            }
          }
        } else {
          epi_chunk_scalar(p, stg, tc.bz, mw0, nb, lane, use_bias, use_resid);
        }
        __syncwarp();  // staging buffer is reused by the next chunk; also reconverges for tcgen05.ld
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[lt & 1]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN, int BK = BLOCK_K, bool DROP = false>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& ta_slab, const CUtensorMap& tb, GemmParams p, bool want_resident,
                int want_slab_rows, cudaStream_t st) {
  using L = SmemLayout<BLOCK_N, BK>;
  auto kern = kr_gemm_kernel<BLOCK_N, A_MN, B_MN, BK, DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL + 1024);
    if (e != cudaSuccess) { kr_set_error(cudaGetErrorString(e)); return KR_ERR_CUDA; }
    attr_set = true;
  }
  {
    int cols = L::TMEM_COLS;
    while (cols < 2 * p.msub * BLOCK_N) cols *= 2;
    p.tmem_cols = cols;                                    // <= 512: kr_gemm_ex only picks msub = 2 for BLOCK_N <= 128
  }
  if (p.slab_stream) {   // geometry chosen by kr_gemm_ex: slab buffers first, then p.stages weight stages
    p.b_resident = 0; p.b_res_bytes = 0; p.conv_slab = 0; p.stage_bytes = L::B_BYTES;
    const int smem_ss = RING_OFFSET0 + p.slab_bufs * p.slab_buf_bytes + p.stages * L::B_BYTES + 1024;
    const int grid_ss = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    kr::launch(kern, grid_ss, GEMM_THREADS, smem_ss, st, ta_slab, tb, p);
    KR_CHECK_LAUNCH();
    return KR_OK;
  }
  // ring geometry: keep the weight operand resident when it fits next to >= 4 A-only stages
  const int avail = SMEM_TOTAL - RING_OFFSET0;
  p.b_resident = 0; p.b_res_bytes = 0; p.stage_bytes = L::A_BYTES + L::B_BYTES;
  const long long bres = (long long)p.total_kb * L::B_BYTES;
  if (want_resident && bres + 4LL * L::A_BYTES <= avail) {
    p.b_resident = 1; p.b_res_bytes = (int)bres; p.stage_bytes = L::A_BYTES;
  }
  p.conv_slab = 0; p.slab_rows = 0;
  if (p.b_resident && !A_MN && p.conv_taps > 1 && want_slab_rows > 0) {
    const int slab_bytes = p.conv_cin_blocks * want_slab_rows * BK * 2;
    if (p.b_res_bytes + 2LL * slab_bytes <= avail) {
      p.conv_slab = 1; p.slab_rows = want_slab_rows; p.stage_bytes = slab_bytes;
    }
  }
  int stages = (avail - p.b_res_bytes) / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > 8 && !p.b_resident) stages = 8;
  p.stages = stages;
  const int smem_bytes = RING_OFFSET0 + p.b_res_bytes + stages * p.stage_bytes + 1024;
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  kr::launch(kern, grid, GEMM_THREADS, smem_bytes, st, p.conv_slab ? ta_slab : ta, tb, p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

template <int BLOCK_N>
int dispatch_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& ta_slab, const CUtensorMap& tb,
                   const GemmParams& p, bool res, int slab_rows, cudaStream_t st) {
  if (!a_mn && !b_mn && p.drop.state != nullptr)
    return launch_gemm<BLOCK_N, false, false, BLOCK_K, true>(ta, ta_slab, tb, p, res, slab_rows, st);
  if (!a_mn && !b_mn) return launch_gemm<BLOCK_N, false, false>(ta, ta_slab, tb, p, res, slab_rows, st);
  if (!a_mn && b_mn) return launch_gemm<BLOCK_N, false, true>(ta, ta_slab, tb, p, res, 0, st);
  if (a_mn && !b_mn) return launch_gemm<BLOCK_N, true, false>(ta, ta_slab, tb, p, res, 0, st);
  return launch_gemm<BLOCK_N, true, true>(ta, ta_slab, tb, p, res, 0, st);
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
                         uint32_t box_inner, uint32_t box_rows, int swizzle64) {
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
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
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
extern "C" int kr_gemm_ex(const kr_gemm_args* a, void* stream) {
  if (a == nullptr) { kr_set_error("kr_gemm_ex: null args"); return KR_ERR_ARG; }
  const int M = a->M, N = a->N, K = a->K, batch = a->batch;
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) { kr_set_error("kr_gemm_ex: empty problem"); return KR_ERR_ARG; }
  const bool conv = a->conv_taps > 0;
  const int bk = (conv && a->conv_cin == 32 && !a->a_mn_major && !a->b_mn_major && N <= 64) ? 32 : BLOCK_K;
  if (conv && (a->a_mn_major || (a->conv_cin % bk) != 0 || K != a->conv_taps * a->conv_cin)) {
    kr_set_error("kr_gemm_ex: conv mode needs K-major A, K == taps * C_in and C_in % 64 == 0 (or C_in == 32 with N <= 64)");
    return KR_ERR_ARG;
  }
  const int total_kb = (K + bk - 1) / bk;
  int splits = a->splits < 1 ? 1 : a->splits;
  if (splits > total_kb) splits = total_kb;
  if (splits > 1 && a->c_mode != C_ATOMIC_F32) { kr_set_error("kr_gemm_ex: split-K needs the atomic epilogue"); return KR_ERR_ARG; }
  if (splits > 1 && (a->resid2 != nullptr || a->C2 != nullptr)) { kr_set_error("kr_gemm_ex: split-K cannot use resid2 / C2"); return KR_ERR_ARG; }
  const int kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per_split - 1) / kb_per_split;  // no empty splits

  // tile width: maximise (wave quantisation efficiency) x (MMA feed efficiency: narrow tiles re-read
  // the 128-row A tile from shared memory more often per flop)
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  int block_n = 64;
  {
    const int cand[4] = {64, 128, 192, 256};
    const double feed[4] = {0.67, 0.92, 0.97, 1.0};
    double best = -1.0;
    for (int i = 0; i < 4; ++i) {
      const int bn = cand[i];
      if (bn > 64 && bn - 64 >= N) continue;                 // more than one empty 64-column group
      const long long tiles = (long long)m_tiles * ((N + bn - 1) / bn) * batch * splits;
      const long long rounds = (tiles + kNumSMs - 1) / kNumSMs;
      const double used = (double)N / (double)(((N + bn - 1) / bn) * bn);
      const double eff = (double)tiles / (double)(rounds * kNumSMs) * feed[i] * used;
      if (eff > best + 1e-9) { best = eff; block_n = bn; }
    }
  }
  if (a->force_block_n == 64 || a->force_block_n == 128 || a->force_block_n == 192 || a->force_block_n == 256)
    block_n = a->force_block_n;

  const bool b_shared = batch > 1 && a->stride_b == 0;
  // Convs whose weights cannot stay resident even as a 128-row tile (the 128 / 256-channel HiFi-GAN stages): 128-column
  // tiles, so that slab-stream mode below runs with swapped operands (N = 256 MMAs, see GemmParams::swap).  (Also taking the
  // convs whose weights WOULD fit — 128 channels, k = 3 — away from the resident-weight path gained 0.05 of 13.7 ms: not done.)
  {
    const long long tiles2 = (long long)((m_tiles + 1) / 2) * ((N + 127) / 128) * batch;
    const bool vec = !(N & 3) && !(a->ldc & 3) && !(a->ldr & 3) && !(a->ldr2 & 3) && !(a->ldc2 & 3);
    if (conv && a->conv_taps > 1 && !a->no_slab && bk == BLOCK_K && !a->b_mn_major && splits == 1 && (batch == 1 || b_shared) &&
        (N % 128) == 0 && vec && a->force_block_n == 0 && tiles2 >= kNumSMs &&
        (long long)total_kb * 128 * bk * 2 + 4LL * BLOCK_M * bk * 2 > SMEM_TOTAL - RING_OFFSET0)
      block_n = 128;
  }
  const uint64_t bstride_a = batch > 1 ? (uint64_t)a->stride_a
                                       : (uint64_t)a->lda * (a->a_mn_major ? K : (conv ? a->a_rows : M));
  const uint64_t bstride_b = (batch > 1 && !b_shared) ? (uint64_t)a->stride_b
                                                      : (uint64_t)a->ldb * (a->b_mn_major ? K : N);
  const int b_batch = b_shared ? 1 : batch;
  CUtensorMap ta, tb;
  int rc;
  if (a->a_mn_major) rc = kr_make_tmap_bf16_3d(&ta, a->A, M, K, batch, a->lda, bstride_a, 64, BLOCK_K);
  else if (conv)     rc = kr_make_tmap_bf16_3d(&ta, a->A, a->conv_cin, a->a_rows, batch, a->lda, bstride_a, bk, BLOCK_M, bk == 32);
  else               rc = kr_make_tmap_bf16_3d(&ta, a->A, K, M, batch, a->lda, bstride_a, BLOCK_K, BLOCK_M);
  if (rc != KR_OK) return rc;
  if (a->b_mn_major) rc = kr_make_tmap_bf16_3d(&tb, a->B, N, K, b_batch, a->ldb, bstride_b, 64, BLOCK_K);
  else               rc = kr_make_tmap_bf16_3d(&tb, a->B, K, N, b_batch, a->ldb, bstride_b, bk, block_n, bk == 32);
  if (rc != KR_OK) return rc;

  GemmParams p{};
  p.msub = 1;
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.splits = splits; p.kb_per_split = kb_per_split;
  p.total_kb = total_kb; p.m_tiles = m_tiles; p.n_tiles = (N + block_n - 1) / block_n;
  p.total_tiles = p.m_tiles * p.n_tiles * batch * splits;
  p.conv_taps = a->conv_taps; p.conv_dil = a->conv_dil; p.conv_row0 = a->conv_row0;
  p.conv_cin_blocks = conv ? a->conv_cin / bk : 0;
  p.C = a->C; p.ldc = a->ldc; p.c_batch_stride = a->stride_c; p.c_mode = a->C != nullptr ? a->c_mode : C_NONE;
  p.C2 = reinterpret_cast<bf16*>(a->C2); p.ldc2 = a->ldc2; p.c2_batch_stride = a->stride_c2; p.act_slope = a->act_slope;
  p.bias = a->bias;
  p.resid = a->resid; p.resid_bf16 = a->resid_dtype; p.ldr = a->ldr; p.r_batch_stride = a->stride_r; p.resid_mod = a->resid_mod;
  p.resid2 = a->resid2; p.resid2_bf16 = a->resid2_dtype; p.ldr2 = a->ldr2; p.r2_batch_stride = a->stride_r2;
  p.alpha = a->alpha; p.beta = a->beta;
  p.b_shared = b_shared ? 1 : 0;
  p.drop = kr_drop_to_device(a->drop);
  if (p.drop.state != nullptr && (batch != 1 || splits != 1 || (N & 3) || (a->ldc & 3) || (a->ldr & 3) ||
                                  a->a_mn_major || a->b_mn_major || conv)) {
    kr_set_error("kr_gemm_ex: the dropout epilogue needs a plain K-major GEMM, batch == 1, no split-K and N / leading dimensions % 4 == 0");
    return KR_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // resident weights: one N tile, no split-K, weights shared by the batch, and enough tiles per CTA to pay off
  const bool res = p.n_tiles == 1 && splits == 1 && (batch == 1 || b_shared) && p.total_tiles >= 2 * kNumSMs;
  // slab tensor map: same tensor, box = all rows a tile touches (128 + (taps-1)*dil, rounded up to 8)
  CUtensorMap ta_slab = ta;
  int slab_rows = 0;
  // weights that cannot stay resident (launch_gemm's rule): slab-stream mode — activation slab per (tile, 64-channel
  // block), streamed weight blocks, two 128-row sub-tiles per weight block when two double-buffered accumulator pairs fit
  // in TMEM.  Measured on B200 (HiFi-GAN C = 128 / 256 stages): the per-tap path moves A + B tiles = 32 KB from L2 per 4
  // MMAs, ~3x the ~42 B/clk/SM the L2 delivers chip-wide, and ran at 0.55 - 0.9 PFLOP/s.
  {
    const int avail = SMEM_TOTAL - RING_OFFSET0;
    const long long b_bytes = (long long)block_n * bk * 2, a_bytes = (long long)BLOCK_M * bk * 2;
    const bool resident_fits = res && (long long)total_kb * b_bytes + 4 * a_bytes <= avail;
    if (conv && a->conv_taps > 1 && !a->no_slab && !resident_fits && bk == BLOCK_K && !a->b_mn_major && splits == 1 &&
        (batch == 1 || b_shared)) {
      // two sub-tiles per weight block only while the halved tile count still fills the machine
      const long long tiles2 = (long long)((m_tiles + 1) / 2) * p.n_tiles * batch;
      for (int msub = (block_n <= 128 && tiles2 >= kNumSMs) ? 2 : 1; msub >= 1 && !p.slab_stream; --msub) {
        const int rows = (BLOCK_M * msub + (a->conv_taps - 1) * a->conv_dil + 15) / 16 * 16;
        if (rows > 512) continue;
        const int bytes = rows * BLOCK_K * 2;
        for (int bufs = 3; bufs >= 2 && !p.slab_stream; --bufs) {
          long long stages = (avail - (long long)bufs * bytes) / b_bytes;
          if (stages < 3) continue;
          p.slab_stream = 1; p.msub = msub; p.slab_bufs = bufs; p.slab_buf_bytes = bytes; p.slab_rows = rows;
          p.swap = (msub == 2 && block_n == 128 && (N % 128) == 0 && !(N & 3) && !(a->ldc & 3) && !(a->ldr & 3) &&
                    !(a->ldr2 & 3) && !(a->ldc2 & 3)) ? 1 : 0;
          p.slab_box_rows = rows > 256 ? rows / 2 : rows;
          p.stages = (int)(stages > MAX_STAGES ? MAX_STAGES : stages);
        }
      }
      if (p.slab_stream) {
        p.m_tiles = (M + BLOCK_M * p.msub - 1) / (BLOCK_M * p.msub);
        p.total_tiles = p.m_tiles * p.n_tiles * batch * splits;
        rc = kr_make_tmap_bf16_3d(&ta_slab, a->A, a->conv_cin, a->a_rows, batch, a->lda, bstride_a, bk, p.slab_box_rows);
        if (rc != KR_OK) return rc;
      }
    }
  }
  if (!p.slab_stream && conv && res && a->conv_taps > 1 && !a->no_slab) {
    slab_rows = (BLOCK_M + (a->conv_taps - 1) * a->conv_dil + 7) / 8 * 8;
    if (slab_rows <= 256) {
      rc = kr_make_tmap_bf16_3d(&ta_slab, a->A, a->conv_cin, a->a_rows, batch, a->lda, bstride_a, bk, slab_rows, bk == 32);
      if (rc != KR_OK) return rc;
    } else {
      slab_rows = 0;
    }
  }
  if (bk == 32) return launch_gemm<64, false, false, 32>(ta, ta_slab, tb, p, res, slab_rows, st);
  if (block_n == 256) return dispatch_major<256>(a->a_mn_major, a->b_mn_major, ta, ta_slab, tb, p, res, slab_rows, st);
  if (block_n == 192) return dispatch_major<192>(a->a_mn_major, a->b_mn_major, ta, ta_slab, tb, p, res, slab_rows, st);
  if (block_n == 128) return dispatch_major<128>(a->a_mn_major, a->b_mn_major, ta, ta_slab, tb, p, res, slab_rows, st);
  return dispatch_major<64>(a->a_mn_major, a->b_mn_major, ta, ta_slab, tb, p, res, slab_rows, st);
}

extern "C" int kr_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int batch,
                            long long lda, long long ldb, long long ldc, long long stride_a,
                            long long stride_b, long long stride_c, int a_mn_major, int b_mn_major,
                            int epi_mode, const float* bias, const float* resid, long long ldr,
                            long long stride_r, int resid_mod, float alpha, int splits,
                            void* stream) {
  kr_gemm_args a{};
  a.A = A; a.B = B; a.M = M; a.N = N; a.K = K; a.batch = batch;
  a.lda = lda; a.ldb = ldb; a.stride_a = stride_a; a.stride_b = stride_b;
  a.a_mn_major = a_mn_major; a.b_mn_major = b_mn_major;
  a.alpha = alpha; a.beta = 1.f; a.bias = bias;
  a.resid = resid; a.resid_dtype = 0; a.ldr = ldr; a.stride_r = stride_r; a.resid_mod = resid_mod;
  a.C = C; a.c_mode = epi_mode; a.ldc = ldc; a.stride_c = stride_c;
  a.splits = splits;
  return kr_gemm_ex(&a, stream);
}
