// kr_hifi_resblock.cu — one HiFi-GAN ResBlock step in ONE kernel (reference inference/hifigan_vocoder.py:31-83:
// xt = c1(lrelu(x)); xt = c2(lrelu(xt)); x = xt + x):
//
//   t        = lrelu(conv1(x_act) + b1, 0.1)                      bf16, lives in SHARED MEMORY only
//   v        = conv2(t) + b2 + resid                              (resid = the fp32 residual stream x)
//   v        = v * beta + resid2                                  (optional: MRF accumulation xs += x / num_kernels)
//   out      = v (fp32, optional),  out_act = bf16(lrelu(v, slope)) (optional: the next conv's operand)
//
// With two kr_gemm_ex launches the intermediate t costs a bf16 write and a bf16 read of the whole activation per step —
// 26 % of the step's DRAM traffic on the HBM-bound 64-channel stage, and a second kernel's worth of per-tile latency on
// the narrow one.  Here both implicit GEMMs run back to back on the tensor cores, with the weights of BOTH convs resident
// in shared memory (streaming a weight block per few MMAs would need the SM's whole 64 B / cycle L2 port: the kernel
// refuses shapes whose weights do not fit; the caller keeps the two-launch path for those).
//
// Both convs are given as lists of K-HALF BLOCKS: block i = a [64 output channels x 32 input channels] bf16 weight block
// applied to the activation row at offset off[i] (relative to the output row) and input-channel half kh[i] (0 / 1).  A
// plain 64-channel tap is two such blocks; the TIME-FOLDED convs of the 32-channel stage (hifigan.py: [L, 32] viewed as
// [L/2, 64]) are block-sparse — their zero halves are simply not listed, which halves the MMA work and the shared memory
// of the folded weights (all nine steps of that stage then fit, including k = 11 with dilation 3 / 5).
//
//   tile = 128 - 2*h2 output rows (h2 = conv2's reach): conv1 produces exactly the 128 rows of t that conv2 needs for them
//   warp 0    TMA producer: weights once (two half blocks per 128B-swizzled 8 KB tile); per tile the x_act slab (128 + conv1's span rows,
//             fetched ONCE — the taps are row-shifted views of it) into a double-buffered slot
//   warp 1    tcgen05.mma issuer: GEMM1 -> acc1 (TMEM), GEMM2 (A = the t tile in smem, taps = row shifts) -> acc2 (TMEM);
//             two K = 16 MMAs per half block (A descriptor: 128B-swizzled row + 64 * kh bytes, B descriptor: the block's half of its tile)
//   warps 2-9 two independent TILE SLOTS of four epilogue warps each (even / odd tiles; own acc1, acc2, t tile, barriers):
//             epilogue 1: acc1 -> +b1 -> lrelu -> zero outside [0, L) (conv2's zero padding applies to t) -> bf16 ->
//             128B-swizzled smem tile;  epilogue 2: residual operands requested BEFORE the wait for GEMM2, then
//             acc2 -> fused bias / residual / MRF / activation -> global, in a coalesced layout through a staged transpose
// Issue order  G1(0), G1(1), G2(0), G1(2), G2(1), ...: while one slot converts its t tile or writes its outputs, the tensor
// pipe works for the other slot.  Weights that only fit next to ONE slot (k = 11 on 64 channels: 176 KB) run in one-slot mode:
// the same issue order with a single slab, t tile and acc2 and four epilogue warps — acc1 stays double-buffered, so GEMM1 of
// the next tile runs under the epilogues of the current one.  (Measured and dropped for that mode: a second MMA-issuing warp
// for GEMM2 — 568 vs 510 us per k = 11 step, and 290 vs 266 us on the two-slot k = 3 steps; converting the t tile with both
// warp sets — the second set has to wait for epilogue 2 of the previous tile, which stages in the t tile: 490 us.)
// Channels-last activations [B, L + 2*halo, 64] with zero halos (halo >= h2 + conv1's reach); weights [64, n * 32] bf16.
#include "kr_common.cuh"
#include "kokoro_b200.h"

namespace {
using namespace kr;

constexpr int RB_THREADS = 320;
constexpr int RB_EPI_WARPS = 8;
constexpr int RB_TROWS = 144;            // t tile: 128 rows + 16 zero rows read by the shifted taps of the invalid output rows
constexpr int RB_MAX_SLAB = 184;         // 128 + 2 * 25 = 178 rows, rounded up to the 8-row swizzle atom
constexpr int RB_SMEM_MAX = 227 * 1024 - 1024;

constexpr int RB_C = 64;                 // physical channels of the activations
constexpr int RB_MAX_BLOCKS = 48;        // half blocks per conv
constexpr int RB_PAIR_BYTES = RB_C * 128; // two half blocks side by side: one 128B-swizzled [64 x 64] bf16 tile (8 KB)

struct RbParams {
  int B, L, halo;                        // activations [B, L + 2*halo, 64]
  int n1, n2, lead1, h2;                 // half blocks of conv1 / conv2; lead1 = -min conv1 offset; h2 = conv2's reach
  int slots;                             // 2: two tile slots ping-pong; 1: one slot (its slab / t tile make room for more weights)
  // block -> start of its A operand inside the slab (conv1) / the t tile (conv2), in 64-byte units: 2 * row offset + half
  unsigned short a1[RB_MAX_BLOCKS], a2[RB_MAX_BLOCKS];
  int slab_rows, rows_out, tiles_per_item, total_tiles;
  const float* b1; const float* b2;
  const float* resid; long long r_ld, r_bs;         // fp32, pointing at time 0 of item 0
  const float* resid2; long long r2_ld, r2_bs;
  float beta, slope;
  float* out; long long o_ld, o_bs;
  bf16* out_act; long long a_ld, a_bs;
};

__global__ void __launch_bounds__(RB_THREADS, 1)
hifi_resblock_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const RbParams p) {
  constexpr int C = RB_C;
  constexpr int CB = 1;
  constexpr int SLAB_BYTES = CB * RB_MAX_SLAB * 128;
  constexpr int T_BYTES = CB * RB_TROWS * 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* wres_bar = bars;
  uint64_t* slab_full = bars + 1;               // [2]
  uint64_t* slab_empty = slab_full + 2;         // [2]
  uint64_t* acc1_full = slab_empty + 2;         // [2] per tile slot
  uint64_t* t_full = acc1_full + 2;             // [2] 4 arrivals (one per epilogue warp of the slot)
  uint64_t* acc2_full = t_full + 2;             // [2]
  uint64_t* acc2_empty = acc2_full + 2;         // [2] 4 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 2);
  uint8_t* slab = smem + 1024;                  // [2][CB][RB_MAX_SLAB rows][128 B]
  const int sh = p.slots - 1;                   // local tile lt -> slot lt & sh, use count lt >> sh
  uint8_t* tt = slab + p.slots * SLAB_BYTES;    // [slots][RB_TROWS rows][128 B]
  // weights: half blocks 2t, 2t+1 of a conv share the 128B-swizzled tile t ([64 rows][128 B]); conv2's tiles follow conv1's
  // (a 64-byte-swizzled 4 KB tile per half block was measured 25 % slower on the MMA-bound k = 7 steps)
  uint8_t* wres = tt + p.slots * T_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h2 = p.h2;
  const int n1 = p.n1, n2 = p.n2;                             // half blocks of the two GEMMs
  const int t1 = (n1 + 1) >> 1, t2 = (n2 + 1) >> 1;           // ... and their weight tiles
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_w1); tma_prefetch_desc(&tm_w2);
    mbar_init(wres_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&slab_full[s], 1); mbar_init(&slab_empty[s], 1);
      mbar_init(&acc1_full[s], 1); mbar_init(&t_full[s], RB_EPI_WARPS / 2);
      mbar_init(&acc2_full[s], 1); mbar_init(&acc2_empty[s], RB_EPI_WARPS / 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 4 * C); tmem_relinquish(); }
  // the 16 spare rows of both t tiles stay zero for the whole kernel
  for (int i = threadIdx.x; i < p.slots * CB * 16 * 8; i += RB_THREADS) {
    const int blk = i / (16 * 8), r = (i / 8) % 16, q = i % 8;        // blk = slot * CB + cb
    *reinterpret_cast<uint4*>(tt + blk * (RB_TROWS * 128) + (128 + r) * 128 + q * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(wres_bar, (uint32_t)((t1 + t2) * RB_PAIR_BYTES));   // (columns beyond n * 32 are zero-filled)
      for (int t = 0; t < t1 + t2; ++t) {
        if (t < t1) tma_load_3d(wres + t * RB_PAIR_BYTES, &tm_w1, wres_bar, t * 64, 0, 0);
        else        tma_load_3d(wres + t * RB_PAIR_BYTES, &tm_w2, wres_bar, (t - t1) * 64, 0, 0);
      }
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int bz = tile / p.tiles_per_item, m0 = (tile % p.tiles_per_item) * p.rows_out;
        const int sb = lt & sh;
        mbar_wait(&slab_empty[sb], ((lt >> sh) & 1) ^ 1);
        mbar_arrive_expect_tx(&slab_full[sb], (uint32_t)(CB * p.slab_rows * 128));
        // t row j <-> time m0 - h2 + j; conv1 block i reads time m0 - h2 + j + (its row offset) - lead1
        const int row0 = p.halo + m0 - h2 - p.lead1;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb)
          tma_load_3d(slab + sb * SLAB_BYTES + cb * (RB_MAX_SLAB * 128), &tm_x, &slab_full[sb], cb * 64, row0, bz);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, C, 0, 0);
      mbar_wait(wres_bar, 0);
      tc_fence_after();
      // GEMM2 of local tile `g`: acc2[slot] = sum over the half blocks of t[slot][rows shifted by off2, half kh2] . W2 block
      auto gemm2 = [&](int g) {
        const int s = g & sh;
        const uint32_t ph = (uint32_t)(g >> sh) & 1u;
        mbar_wait(&t_full[s], ph);                  // the slot's epilogue warps wrote the t tile (generic proxy, fenced)
        mbar_wait(&acc2_empty[s], ph ^ 1u);         // ... and finished reading acc2 of the slot's previous tile
        tc_fence_after();
        const uint32_t acc2 = tmem_base + 2 * C + s * C;
        const uint32_t a_base = smem_u32(tt + s * T_BYTES), b_base = smem_u32(wres + t1 * RB_PAIR_BYTES);
        // (unrolled: the block table reads of the next blocks are issued ahead — one thread feeds the tensor pipe, and a
        //  dependent constant load per two 32-clock MMAs made the k = 7 steps issue-bound: 361 vs 285 us)
        //  and the descriptors are one 64-bit add each: the start-address field counts 16-byte units and cannot carry)
        const uint64_t da0 = make_smem_desc_sw128(a_base, 16, 1024), db0 = make_smem_desc_sw128(b_base, 16, 1024);
#pragma unroll 4
        for (int i = 0; i < n2; ++i) {
          const uint64_t da = da0 + (uint32_t)p.a2[i] * 4u;
          const uint64_t db = db0 + (uint32_t)((i >> 1) * (RB_PAIR_BYTES / 16) + (i & 1) * 4);
          umma_bf16_ss(acc2, da, db, idesc, i > 0 ? 1u : 0u);
          umma_bf16_ss(acc2, da + 2, db + 2, idesc, 1u);
        }
        umma_commit(&acc2_full[s]);
      };
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int s = lt & sh;
        // ---- GEMM1: acc1[slot] = sum over the half blocks of slab[rows shifted by off1, channel half kh1] . W1 block ----
        // (acc1[slot] is free: GEMM2 of the slot's previous tile was issued, i.e. its t tile — read from acc1 — was complete)
        mbar_wait(&slab_full[s], (lt >> sh) & 1);
        tc_fence_after();
        // acc1 is double-buffered by tile parity in BOTH modes (in one-slot mode only the slab, the t tile and acc2 are single):
        // acc1[lt & 1] was last read by epilogue 1 of tile lt - 2, which finished before GEMM2(lt - 2) was issued
        const uint32_t acc1 = tmem_base + (lt & 1) * C;
        const uint32_t a_base = smem_u32(slab + s * SLAB_BYTES), b_base = smem_u32(wres);
        const uint64_t da0 = make_smem_desc_sw128(a_base, 16, 1024), db0 = make_smem_desc_sw128(b_base, 16, 1024);
#pragma unroll 4
        for (int i = 0; i < n1; ++i) {
          const uint64_t da = da0 + (uint32_t)p.a1[i] * 4u;
          const uint64_t db = db0 + (uint32_t)((i >> 1) * (RB_PAIR_BYTES / 16) + (i & 1) * 4);
          umma_bf16_ss(acc1, da, db, idesc, i > 0 ? 1u : 0u);
          umma_bf16_ss(acc1, da + 2, db + 2, idesc, 1u);
        }
        umma_commit(&slab_empty[s]);
        umma_commit(&acc1_full[lt & 1]);
        if (lt > 0) gemm2(lt - 1);
      }
      if (lt > 0) gemm2(lt - 1);
    }
  } else {
    // ===== epilogue warps 2..9: slot = (warp - 2) / 4 handles the local tiles with lt % 2 == slot; TMEM lane group = warp % 4
    const int lg = warp & 3, slot = (warp - 2) >> 2;
    const int j = lg * 32 + lane;                                       // row of the tile owned by this thread
    const uint32_t t_lane = static_cast<uint32_t>(lg * 32) << 16;
    const uint32_t acc2 = tmem_base + 2 * C + slot * C + t_lane;
    uint8_t* tts = tt + slot * T_BYTES;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      if (slot > sh || (lt & sh) != slot) continue;             // (one-slot mode: warps 6-9 only keep the block barriers)
      const uint32_t ph = (uint32_t)(lt >> sh) & 1u;
      const int bz = tile / p.tiles_per_item, m0 = (tile % p.tiles_per_item) * p.rows_out;
      // ---- epilogue 1: t row j (time m0 - h2 + j) ----
      const uint32_t acc1 = tmem_base + (lt & 1) * C + t_lane;
      mbar_wait(&acc1_full[lt & 1], (uint32_t)(lt >> 1) & 1u);
      tc_fence_after();
      {
        const int time = m0 - h2 + j;
        const bool inside = time >= 0 && time < p.L;                    // conv2 pads t with ZEROS outside the signal
#pragma unroll 1
        for (int c = 0; c < C / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(acc1 + c * 32, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1 + c * 32) + q);
            float a0 = __uint_as_float(r[4 * q]) + bb.x, a1 = __uint_as_float(r[4 * q + 1]) + bb.y;
            float a2 = __uint_as_float(r[4 * q + 2]) + bb.z, a3 = __uint_as_float(r[4 * q + 3]) + bb.w;
            a0 = fmaxf(a0, 0.1f * a0); a1 = fmaxf(a1, 0.1f * a1); a2 = fmaxf(a2, 0.1f * a2); a3 = fmaxf(a3, 0.1f * a3);
            pk[2 * q] = inside ? pack_bf16(a0, a1) : 0u;
            pk[2 * q + 1] = inside ? pack_bf16(a2, a3) : 0u;
          }
          // columns [32c, 32c+32) of row j: 16-byte pieces u0 .. u0+3 of K block cb, XOR-swizzled by the row
          const int cb = c >> 1, u0 = (c & 1) * 4;
          uint8_t* rowp = tts + cb * (RB_TROWS * 128) + j * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(rowp + (((u0 + q) ^ (j & 7)) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      fence_proxy_async_smem();                                         // generic-proxy smem writes -> visible to UMMA
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[slot]);
      // ---- epilogue 2: output rows (time m0 + row), valid for row < rows_out and time < L ----
      // All global traffic is issued in a COALESCED layout: lane -> (row 4*i + crow, columns [4*xs, 4*xs + 4)) of a 32 x 32
      // chunk, i.e. 4 rows x 128 contiguous bytes per warp instruction; the accumulator chunk gets there from the TMEM row
      // layout through a swizzled 4 KB staging tile, for which the warp borrows its own 32 rows of the slot's t tile (GEMM2
      // has finished reading it when acc2 is full, and epilogue 1 of the slot's next tile rewrites it completely).
      // (A row-per-thread epilogue was measured first: 32 separate 16-byte requests per instruction made the step 40 % slower
      // than two kr_gemm_ex launches.)
      {
        const int crow = lane >> 3, xs = lane & 7;
        float* stg = reinterpret_cast<float*>(tts + lg * 4096);       // rows [32 lg, 32 lg + 32) of K block 0 of the t tile
        uint32_t okm = 0;                                             // bit i: row 4*i + crow of this warp's slab is valid
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = lg * 32 + 4 * i + crow;
          okm |= (row < p.rows_out && m0 + row < p.L) ? (1u << i) : 0u;
        }
        const long long t0 = (long long)(m0 + lg * 32 + crow);        // time of this lane's first row; rows advance by 4
        const float* rp = p.resid + (long long)bz * p.r_bs + t0 * p.r_ld + 4 * xs;
        const float* r2p = p.resid2 != nullptr ? p.resid2 + (long long)bz * p.r2_bs + t0 * p.r2_ld + 4 * xs : nullptr;
        float* op = p.out != nullptr ? p.out + (long long)bz * p.o_bs + t0 * p.o_ld + 4 * xs : nullptr;
        bf16* ap = p.out_act != nullptr ? p.out_act + (long long)bz * p.a_bs + t0 * p.a_ld + 4 * xs : nullptr;
        // the residual operands of the FIRST chunk are requested before the wait for GEMM2
        float4 rr[8], r2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool okr = (okm >> i) & 1u;
          rr[i] = okr ? *reinterpret_cast<const float4*>(rp + (long long)(4 * i) * p.r_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
          r2[i] = (okr && r2p != nullptr) ? *reinterpret_cast<const float4*>(r2p + (long long)(4 * i) * p.r2_ld)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&acc2_full[slot], ph);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < C / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(acc2 + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
                make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
          __syncwarp();
          const int n = c * 32;
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + n + 4 * xs));
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = *reinterpret_cast<const float4*>(stg + (4 * i + crow) * 32 + ((xs ^ (((i & 1) << 2) | crow)) << 2));
            v[i].x = (v[i].x + bb.x + rr[i].x) * p.beta + r2[i].x; v[i].y = (v[i].y + bb.y + rr[i].y) * p.beta + r2[i].y;
            v[i].z = (v[i].z + bb.z + rr[i].z) * p.beta + r2[i].z; v[i].w = (v[i].w + bb.w + rr[i].w) * p.beta + r2[i].w;
          }
          __syncwarp();                                 // the staging tile is rewritten by the next chunk
          if (c + 1 < C / 32) {                         // next chunk's residuals fly under this chunk's stores
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool okr = (okm >> i) & 1u;
              rr[i] = okr ? *reinterpret_cast<const float4*>(rp + (long long)(4 * i) * p.r_ld + n + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
              r2[i] = (okr && r2p != nullptr) ? *reinterpret_cast<const float4*>(r2p + (long long)(4 * i) * p.r2_ld + n + 32)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (op != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if ((okm >> i) & 1u) *reinterpret_cast<float4*>(op + (long long)(4 * i) * p.o_ld + n) = v[i];
          }
          if (ap != nullptr) {
            const float sl = p.slope;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if ((okm >> i) & 1u)
                *reinterpret_cast<uint2*>(ap + (long long)(4 * i) * p.a_ld + n) =
                    make_uint2(pack_bf16(fmaxf(v[i].x, v[i].x * sl), fmaxf(v[i].y, v[i].y * sl)),
                               pack_bf16(fmaxf(v[i].z, v[i].z * sl), fmaxf(v[i].w, v[i].w * sl)));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[slot]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 4 * C); }
}

constexpr int rb_fixed_smem(int slots) { return 1024 + slots * (RB_MAX_SLAB * 128 + RB_TROWS * 128); }

int launch_rb(const CUtensorMap& tx, const CUtensorMap& t1, const CUtensorMap& t2, const RbParams& p, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(hifi_resblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM_MAX + 1024);
    if (e != cudaSuccess) { kr_set_error(cudaGetErrorString(e)); return KR_ERR_CUDA; }
    attr = true;
  }
  const int smem = rb_fixed_smem(p.slots) + ((p.n1 + 1) / 2 + (p.n2 + 1) / 2) * RB_PAIR_BYTES + 1024;
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  kr::launch(hifi_resblock_kernel, grid, RB_THREADS, smem, st, tx, t1, t2, p);
  return KR_OK;
}

// tile slots the kernel can run with for (n1, n2) half blocks: 2 (ping-pong), 1 (k = 11 of the 64-channel stage: 44 blocks),
// 0 = the weights do not fit
int rb_slots(int n1, int n2) {
  if (n1 < 1 || n2 < 1 || n1 > RB_MAX_BLOCKS || n2 > RB_MAX_BLOCKS) return 0;
  for (int slots = 2; slots >= 1; --slots)
    if (rb_fixed_smem(slots) + ((n1 + 1) / 2 + (n2 + 1) / 2) * RB_PAIR_BYTES <= RB_SMEM_MAX) return slots;
  return 0;
}

}  // namespace

extern "C" int kr_hifi_resblock(const void* x_act, int B, long long L, int halo, const void* w1, int n1, const int* off1,
                                const int* kh1, const float* b1, const void* w2, int n2, const int* off2, const int* kh2,
                                const float* b2, const float* resid, long long r_ld, long long r_bs, const float* resid2,
                                long long r2_ld, long long r2_bs, float beta, float* out, long long o_ld, long long o_bs,
                                void* out_act, long long a_ld, long long a_bs, float slope, void* stream) {
  if (B <= 0 || L <= 0) return KR_OK;
  if (off1 == nullptr || kh1 == nullptr || off2 == nullptr || kh2 == nullptr) { kr_set_error("kr_hifi_resblock: null block lists"); return KR_ERR_ARG; }
  if (rb_slots(n1, n2) == 0) {
    kr_set_error("kr_hifi_resblock: 1 .. 48 half blocks per conv, and the weights of both convs must fit in shared memory (see kr_hifi_resblock_resident)");
    return KR_ERR_UNSUPPORTED;
  }
  int lo1 = 0, hi1 = 0, h2 = 0;
  for (int i = 0; i < n1; ++i) {
    if (kh1[i] < 0 || kh1[i] > 1) { kr_set_error("kr_hifi_resblock: channel half must be 0 or 1"); return KR_ERR_ARG; }
    lo1 = off1[i] < lo1 ? off1[i] : lo1; hi1 = off1[i] > hi1 ? off1[i] : hi1;
  }
  for (int i = 0; i < n2; ++i) {
    if (kh2[i] < 0 || kh2[i] > 1) { kr_set_error("kr_hifi_resblock: channel half must be 0 or 1"); return KR_ERR_ARG; }
    const int r = off2[i] < 0 ? -off2[i] : off2[i];
    h2 = r > h2 ? r : h2;
  }
  const int lead = -lo1, reach = lead > hi1 ? lead : hi1;
  const int slab_rows = (128 + lead + hi1 + 7) / 8 * 8;
  if (h2 > 8 || slab_rows > RB_MAX_SLAB || halo < h2 + reach) {
    kr_set_error("kr_hifi_resblock: receptive field too large (conv2 reach <= 8, 128 + conv1 span <= 184 rows, halo >= conv2 reach + conv1 reach)");
    return KR_ERR_UNSUPPORTED;
  }
  if (resid == nullptr || (out == nullptr && out_act == nullptr)) { kr_set_error("kr_hifi_resblock: needs resid and an output"); return KR_ERR_ARG; }
  if ((r_ld & 3) || (r2_ld & 3) || (o_ld & 3) || (a_ld & 3) || (r_bs & 3) || (r2_bs & 3) || (o_bs & 3) || (a_bs & 3)) {
    kr_set_error("kr_hifi_resblock: leading dimensions / batch strides must be multiples of 4 elements"); return KR_ERR_ARG;
  }
  RbParams p{};
  p.B = B; p.L = (int)L; p.halo = halo; p.n1 = n1; p.n2 = n2; p.lead1 = lead; p.h2 = h2; p.slots = rb_slots(n1, n2);
  for (int i = 0; i < n1; ++i) p.a1[i] = (unsigned short)(2 * (off1[i] + lead) + kh1[i]);
  for (int i = 0; i < n2; ++i) p.a2[i] = (unsigned short)(2 * (off2[i] + h2) + kh2[i]);
  p.slab_rows = slab_rows; p.rows_out = 128 - 2 * h2;
  p.tiles_per_item = (int)((L + p.rows_out - 1) / p.rows_out);
  p.total_tiles = p.tiles_per_item * B;
  p.b1 = b1; p.b2 = b2; p.resid = resid; p.r_ld = r_ld; p.r_bs = r_bs;
  p.resid2 = resid2; p.r2_ld = r2_ld; p.r2_bs = r2_bs; p.beta = beta; p.slope = slope;
  p.out = out; p.o_ld = o_ld; p.o_bs = o_bs; p.out_act = reinterpret_cast<bf16*>(out_act); p.a_ld = a_ld; p.a_bs = a_bs;
  const long long rows_phys = L + 2LL * halo;
  CUtensorMap tx, t1, t2;
  int rc;
  // activation: (64, rows_phys, B), box (64 channels, slab_rows, 1); rows beyond the tensor are zero-filled by TMA
  if ((rc = kr_make_tmap_bf16_3d(&tx, x_act, RB_C, rows_phys, B, RB_C, rows_phys * RB_C, 64, slab_rows)) != KR_OK) return rc;
  // weights: (K = n * 32, 64 rows, 1), box (64 k = two half blocks, 64 rows), 128-byte swizzle
  if ((rc = kr_make_tmap_bf16_3d(&t1, w1, (unsigned long long)n1 * 32, RB_C, 1, (unsigned long long)n1 * 32, (unsigned long long)n1 * 32 * RB_C, 64, RB_C)) != KR_OK) return rc;
  if ((rc = kr_make_tmap_bf16_3d(&t2, w2, (unsigned long long)n2 * 32, RB_C, 1, (unsigned long long)n2 * 32, (unsigned long long)n2 * 32 * RB_C, 64, RB_C)) != KR_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  rc = launch_rb(tx, t1, t2, p, st);
  if (rc != KR_OK) return rc;
  KR_CHECK_LAUNCH();
  return KR_OK;
}

// Non-zero if the weights of both convs of a ResBlock step (n1 / n2 half blocks) fit in shared memory next to the
// activation slabs: the number of tile slots the kernel will run with (2 = ping-pong, 1); 0 = the caller keeps the
// two-launch path.
extern "C" int kr_hifi_resblock_resident(int n1, int n2) { return rb_slots(n1, n2); }
