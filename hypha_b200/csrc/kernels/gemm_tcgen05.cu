// Persistent warp-specialised bf16 GEMM on the 5th-gen tensor cores (sm_100a).
//
//   D[M,N] (+)= A . B,  fp32 accumulation in TMEM, bf16 output.
//
//  * operands staged global -> shared by TMA (cp.async.bulk.tensor, 128B swizzle)
//    through a 4-stage full/empty mbarrier ring,
//  * tcgen05.mma (128x256x16, kind::f16) issued by ONE thread, accumulators live
//    in TMEM, double-buffered (2 x 256 columns) so the epilogue of tile i overlaps
//    the mainloop of tile i+1,
//  * epilogue warps: tcgen05.ld -> bf16 -> swizzled shared -> TMA store,
//  * persistent grid (one CTA per SM) with grouped-M rasterisation for L2 reuse,
//  * A and B may each be K-major or MN-major (UMMA descriptor major bits), so
//    forward (x.W^T), dgrad (dy.W) and wgrad (dy^T.x) of a Linear layer all run
//    without materialising transposes.
//
// This is the hot op behind every transformer GEMM the reference reaches through
// `model(**batch)` / `accelerator.backward(loss)` (training.py:108-112).
#include <mutex>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace hb {

namespace {

using namespace sm100;

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kAStage = BM * BK * 2;       // 16 KB
constexpr int kBStage = BN * BK * 2;       // 32 KB
constexpr int kEpiWarpBuf = 32 * 128;      // 32 rows x 64 bf16, per buffer
constexpr int kEpiBytes = 4 * 2 * kEpiWarpBuf;
constexpr int kTmemCols = 512;             // 2 accumulator stages x 256 fp32 columns
constexpr int kGroupM = 16;
constexpr int kThreads = 256;

struct SmemLayout {
  static constexpr int a_off = 0;
  static constexpr int b_off = a_off + kStages * kAStage;
  static constexpr int epi_off = b_off + kStages * kBStage;
  static constexpr int bar_off = epi_off + kEpiBytes;
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
  static constexpr int total = bar_off + (2 * kStages + 4) * 8 + 16;
};
constexpr int kSmemBytes = SmemLayout::total + 1024;  // slack for manual 1 KB alignment

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;      // 1 => operand is MN-major
  int accumulate;
  const bf16* D;
  int ldd;
  int tiles_m, tiles_n, num_tiles;
  int a_fmt, b_fmt;        // fp8 only: 0 = e4m3, 1 = e5m2
  const float* scale;      // fp8 only: device dequantisation factor of A
  const float* scale_b;    // fp8 only: device dequantisation factor of B
  int swiglu;              // 2-CTA kernel: the output columns interleave gate/up; ALSO store silu(gate)*up through tma_e
  const bf16* gu;          // 2-CTA kernel, dswiglu epilogue: D is NOT written; instead the [M, N] accumulator is taken as
  int ld_gu;               // d(act) and d(gate|up)[M, 2N] = swiglu'(gu) * d(act) leaves through tma_e (interleaved columns)
};

__device__ __forceinline__ void tile_coords(int t, const GemmParams& p, int& tm, int& tn) {
  const int per_group = kGroupM * p.tiles_n;
  const int g = t / per_group;
  const int first_m = g * kGroupM;
  const int gsz = min(p.tiles_m - first_m, kGroupM);
  const int r = t - g * per_group;
  tm = first_m + (r % gsz);
  tn = r / gsz;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                         const __grid_constant__ CUtensorMap tma_b,
                         const __grid_constant__ CUtensorMap tma_d, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t a_smem = smem_base + SmemLayout::a_off;
  const uint32_t b_smem = smem_base + SmemLayout::b_off;
  const uint32_t bars = smem_base + SmemLayout::bar_off;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_ptr_smem = bars + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + SmemLayout::bar_off + 8 * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_d);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0 && lane == 0) {
    // ============================ TMA producer ============================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int tm, tn;
      tile_coords(t, p, tm, tn);
      const int m0 = tm * BM, n0 = tn * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), kAStage + kBStage);
        const int k0 = kb * BK;
        const uint32_t sa = a_smem + stage * kAStage;
        const uint32_t sb = b_smem + stage * kBStage;
        if (!p.a_mn) {
          tma_load_2d(sa, &tma_a, full_bar(stage), k0, m0);  // box {64 k, 128 m}
        } else {
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)                     // box {64 m, 64 k}
            tma_load_2d(sa + i * (BK * 128), &tma_a, full_bar(stage), m0 + i * 64, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(sb, &tma_b, full_bar(stage), k0, n0);  // box {64 k, 256 n}
        } else {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)                     // box {64 n, 64 k}
            tma_load_2d(sb + i * (BK * 128), &tma_b, full_bar(stage), n0 + i * 64, k0);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ============================ MMA issuer ==============================
    const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
    // K-major : rows at 128 B pitch, 8-row groups 1024 B apart; advance 32 B per UMMA_K.
    // MN-major: 64-element MN chunks BK*128 B apart (LBO), 8-k groups 1024 B apart;
    //           advance 16 k-rows = 2048 B per UMMA_K.
    const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, b_lbo = p.b_mn ? BK * 128 : 16;
    const uint32_t a_kadv = p.a_mn ? UMMA_K * 128 : UMMA_K * 2;
    const uint32_t b_kadv = p.b_mn ? UMMA_K * 128 : UMMA_K * 2;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = a_smem + stage * kAStage;
        const uint32_t sb = b_smem + stage * kBStage;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t ad = umma_smem_desc(sa + k * a_kadv, a_lbo, 1024);
          const uint64_t bd = umma_smem_desc(sb + k * b_kadv, b_lbo, 1024);
          umma_bf16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tfull_bar(acc));      // accumulator complete -> epilogue
    }
  } else if (warp >= 4) {
    // ============================ epilogue ================================
    const int ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    const uint32_t epi_smem = smem_base + SmemLayout::epi_off + ew * 2 * kEpiWarpBuf;
    int it = 0;
    int buf = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      int tm, tn;
      tile_coords(t, p, tm, tn);
      const int m0 = tm * BM, n0 = tn * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      const int row = m0 + ew * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_addr + c * 64, r0);
        tmem_ld_32x32(t_addr + c * 64 + 32, r1);
        tmem_ld_wait();
        if (c == BN / 64 - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        const int col0 = n0 + c * 64;
        uint32_t packed[32];
        if (p.accumulate) {
          const bool row_ok = row < p.M;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 prev = make_uint4(0, 0, 0, 0);
            if (row_ok && col0 + j * 8 + 8 <= p.N)
              prev = *reinterpret_cast<const uint4*>(p.D + (size_t)row * p.ldd + col0 + j * 8);
            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = j * 8 + e * 2;
              const float2 o = __bfloat1622float2(pv[e]);
              const float x0 = __uint_as_float(i < 32 ? r0[i & 31] : r1[i & 31]) + o.x;
              const float x1 = __uint_as_float(i + 1 < 32 ? r0[(i + 1) & 31] : r1[(i + 1) & 31]) + o.y;
              __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
              packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(r0[2 * i]),
                                                      __uint_as_float(r0[2 * i + 1]));
            __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(r1[2 * i]),
                                                      __uint_as_float(r1[2 * i + 1]));
            packed[i] = *reinterpret_cast<uint32_t*>(&h0);
            packed[16 + i] = *reinterpret_cast<uint32_t*>(&h1);
          }
        }
        // the buffer we are about to overwrite was handed to TMA two chunks ago
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const uint32_t sbuf = epi_smem + buf * kEpiWarpBuf;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dst = sbuf + lane * 128 + ((j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(packed[4 * j]),
                       "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (col0 < p.N && m0 + ew * 32 < p.M) tma_store_2d(&tma_d, sbuf, col0, m0 + ew * 32);
          tma_store_commit();
        }
        buf ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// =====================================================================================
// 2-CTA variant: a cluster of two CTAs (one SM pair) computes a 256x256 tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256).  Each CTA stages ITS half of A (128 rows) and
// its half of B (128 of the 256 N rows), so per-SM shared-memory fill traffic for B is
// halved and 6 pipeline stages fit.  The leader CTA (rank 0) issues the MMAs and owns the
// barriers the MMA thread waits on (full / tmem_empty); MMA completion is multicast to both
// CTAs' empty / tmem_full barriers; each CTA's epilogue drains its own 128 TMEM lanes.
// =====================================================================================
constexpr int kGroupM2 = 8;

// kMode 0: bf16 operands (kind::f16, 64-element K blocks, UMMA_K = 16), 256x256 tiles
// kMode 1: fp8 e4m3/e5m2 operands, per-tensor scales (kind::f8f6f4, 128-element K blocks, UMMA_K = 32);
//          the fp32 accumulator is multiplied by the two dequantisation factors in the epilogue
// kMode 2: MXFP8 -- fp8 operands with one UE8M0 scale per 32 K-elements of every row of A and B
//          (kind::mxf8f6f4.block_scale).  The scale factors of a K block travel with the operands (one more
//          TMA box per operand and stage), are copied shared -> tensor memory by tcgen05.cp from a thread of their own
//          (in order with the MMAs the copies cost the tensor pipe a third of its time), and are applied inside the
//          tensor core.  Both operands K-major.
//          256x256 tiles like the other modes (a 256x128 tile measured 1.5 PFLOP/s against 2.8 of the per-tensor
//          kernel: its 171 FLOP per operand byte runs into the ~6.3 KB/clk the L2 can feed the SMs).  Two 256-column
//          accumulators fill the 512 TMEM columns, so the scale-factor columns of tile i (one 12-column slot per stage) live in
//          the first 64 columns of the OTHER accumulator: the epilogue of tile i-1 reads those columns first and says so (sf_free), the rest
//          of its drain still overlaps the main loop of tile i.
template <int kMode>
struct Cfg2 {
  static constexpr bool kFp8 = kMode != 0;
  static constexpr bool kMx = kMode == 2;
  static constexpr int TN = 256;                        // tile columns
  // MXFP8: the scale factors cost 1.5 KB per stage -> 5 stages; TMEM slot s (12 columns) belongs to stage s
  static constexpr int stages = kMx ? 5 : 6;
  static constexpr int epi_bufs = 2;
  static constexpr int a_stage = 128 * BK * 2;          // 16 KB: this CTA's 128 rows of A
  static constexpr int b_stage = (TN / 2) * BK * 2;     // this CTA's half of the B rows
  static constexpr int sf_stage = kMx ? 1536 : 0;       // SFA (this CTA's 128 rows, 512 B) + SFB (all 256 columns, 1024 B)
  static constexpr int a_off = 0;
  static constexpr int b_off = a_off + stages * a_stage;
  static constexpr int sf_off = b_off + stages * b_stage;
  // the epilogue staging buffers are written with the 128-byte swizzle pattern of a 1024-byte aligned tile
  static constexpr int epi_off = (sf_off + stages * sf_stage + 1023) / 1024 * 1024;
  static constexpr int bar_off = epi_off + 4 * epi_bufs * kEpiWarpBuf;
  static constexpr int total = bar_off + (3 * stages + 6) * 8 + 16;
  static constexpr int smem_bytes = total + 1024;
};
static_assert(Cfg2<0>::smem_bytes <= 232448 && Cfg2<2>::smem_bytes <= 232448, "shared memory budget");
static_assert(Cfg2<0>::epi_off % 1024 == 0 && Cfg2<1>::epi_off % 1024 == 0 && Cfg2<2>::epi_off % 1024 == 0 &&
              Cfg2<2>::b_off % 1024 == 0, "swizzled tiles need 1024-byte alignment");

__device__ __forceinline__ void tile_coords2(int t, const GemmParams& p, int& tm, int& tn) {
  const int per_group = kGroupM2 * p.tiles_n;
  const int g = t / per_group;
  const int first_m = g * kGroupM2;
  const int gsz = min(p.tiles_m - first_m, kGroupM2);
  const int r = t - g * per_group;
  tm = first_m + (r % gsz);
  tn = r / gsz;
}

template <int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tma_a,
                              const __grid_constant__ CUtensorMap tma_b,
                              const __grid_constant__ CUtensorMap tma_d,
                              const __grid_constant__ CUtensorMap tma_e,      // second output [M, N/2] (swiglu epilogue)
                              const __grid_constant__ CUtensorMap tma_sfa,    // MXFP8: scale factors of A, [row block x K block] x 512 B
                              const __grid_constant__ CUtensorMap tma_sfb,    // MXFP8: scale factors of B
                              const GemmParams p) {
  using C2 = Cfg2<kMode>;
  constexpr bool kFp8 = C2::kFp8;
  constexpr bool kMx = C2::kMx;
  constexpr int TN = C2::TN;
  constexpr int kStages2 = C2::stages;
  constexpr int kAStage2 = C2::a_stage;
  constexpr int kBStage2 = C2::b_stage;
  constexpr int kEB = C2::epi_bufs;               // staging buffers per epilogue warp
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t a_smem = smem_base + C2::a_off;
  const uint32_t b_smem = smem_base + C2::b_off;
  const uint32_t sf_smem = smem_base + C2::sf_off;
  const uint32_t bars = smem_base + C2::bar_off;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages2 + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages2 + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages2 + 2 + s); };
  auto sffree_bar = [&](int s) { return bars + 8u * (2 * kStages2 + 4 + s); };
  auto sfready_bar = [&](int s) { return bars + 8u * (2 * kStages2 + 6 + s); };
  const uint32_t tmem_ptr_smem = bars + 8u * (3 * kStages2 + 6);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + C2::bar_off + 8 * (3 * kStages2 + 6));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  constexpr int kElem = kFp8 ? 1 : 2;            // bytes per operand element
  constexpr int BKE = 128 / kElem;               // K elements per 128-byte swizzle row
  constexpr int UK = 32 / kElem;                 // K elements per tcgen05.mma
  constexpr int kChunk = BKE * 128;              // bytes of one MN-major chunk (BKE k-rows x 128 B)
  constexpr int kMnPerChunk = 128 / kElem;       // MN elements per 128-byte row
  const int num_kb = (p.K + BKE - 1) / BKE;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_d);
    if constexpr (kMx) {
      tma_prefetch_desc(&tma_sfa);
      tma_prefetch_desc(&tma_sfb);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(full_bar(s), 1);    // leader's arrive.expect_tx (bytes of BOTH CTAs land here)
      mbar_init(empty_bar(s), 1);   // one multicast tcgen05.commit per phase
      mbar_init(sfready_bar(s), 1); // MXFP8: the scale factors of this stage are in tensor memory (leader only)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);  // 4 epilogue warps x 2 CTAs (used on the leader only)
      mbar_init(sffree_bar(s), 8);  // MXFP8: the first columns of accumulator s have been read
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0 && lane == 0) {
    // ============================ TMA producer (both CTAs) =================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
      int tm, tn;
      tile_coords2(t, p, tm, tn);
      const int m0 = tm * 256 + (int)cta_rank * 128;
      const int n0 = tn * TN + (int)cta_rank * (TN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * (kAStage2 + kBStage2 + C2::sf_stage));
        const uint32_t lbar = mapa_shared(full_bar(stage), 0);   // the leader's barrier
        const int k0 = kb * BKE;
        const uint32_t sa = a_smem + stage * kAStage2;
        const uint32_t sb = b_smem + stage * kBStage2;
        if (!p.a_mn) {
          tma_load_2d_2cta(sa, &tma_a, lbar, k0, m0);            // box {128 B of k, 128 m}
        } else {
#pragma unroll
          for (int i = 0; i < 128 / kMnPerChunk; ++i)            // box {128 B of m, BKE k}
            tma_load_2d_2cta(sa + i * kChunk, &tma_a, lbar, m0 + i * kMnPerChunk, k0);
        }
        if (!p.b_mn) {
          tma_load_2d_2cta(sb, &tma_b, lbar, k0, n0);            // box {128 B of k, TN / 2 n}
        } else {
#pragma unroll
          for (int i = 0; i < 128 / kMnPerChunk; ++i)
            tma_load_2d_2cta(sb + i * kChunk, &tma_b, lbar, n0 + i * kMnPerChunk, k0);
        }
        if (++stage == kStages2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (kMx && warp == 3 && lane == 0) {
    // ============================ scale-factor producer (both CTAs) =========
    // A second TMA-issuing thread: with the three extra boxes per stage on the operand producer, the single thread's
    // issue rate (5 TMA instructions per 512 tensor-core cycles) was what the block-scaled kernel ran at.
    // Tables are [K block][128-row block][512 B]: this CTA's 128 rows of A, and the two adjacent blocks of the tile's
    // 256 columns of B in one box.  The bytes land on the same full barrier the leader's operand producer armed.
    const int mblocks = (p.M + 127) >> 7, nblocks = (p.N + 127) >> 7;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
      int tm, tn;
      tile_coords2(t, p, tm, tn);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t lbar = mapa_shared(full_bar(stage), 0);
        const uint32_t ssf = sf_smem + stage * C2::sf_stage;
        tma_load_2d_2cta(ssf, &tma_sfa, lbar, 0, kb * mblocks + tm * 2 + (int)cta_rank);
        tma_load_2d_2cta(ssf + 512, &tma_sfb, lbar, 0, kb * nblocks + tn * 2);
        if (++stage == kStages2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (kMx && warp == 2 && lane == 0 && leader) {
    // ============================ scale-factor copier (leader CTA) ==========
    // tcgen05.cp issued by the MMA thread executes in order WITH its MMAs: ncu showed the tensor pipe 64 % active
    // against 96 % of the unscaled kernel, ~100 idle cycles per copy and three copies per 512 cycles of MMAs.  Issued
    // from another thread the copies are ordered with the MMAs only through sfready (tcgen05.commit -> mbarrier).
    // Slot s = 12 TMEM columns of stage s, inside the first 64-column chunk of the accumulator NOT being written.
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = cluster_id; t < p.num_tiles; t += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t tsf0 = tmem_base + (uint32_t)((acc ^ 1) * TN);
      if (it > 0) {                        // tile it-1's epilogue reads those columns first and says so
        mbar_wait(sffree_bar(acc ^ 1), (uint32_t)((it - 1) >> 1) & 1u);
        tc_fence_after();
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);          // operands AND scale factors of both CTAs have landed
        tc_fence_after();
        const uint32_t ssf = sf_smem + stage * C2::sf_stage;
        const uint32_t ta = tsf0 + (uint32_t)(stage * 12);
        tmem_cp_sf_2cta(ta, umma_sf_desc(ssf));
        tmem_cp_sf_2cta(ta + 4u, umma_sf_desc(ssf + 512));
        tmem_cp_sf_2cta(ta + 8u, umma_sf_desc(ssf + 1024));
        umma_commit_1of2(sfready_bar(stage));       // arrives on the leader's barrier when the copies are done
        if (++stage == kStages2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ============================ MMA issuer (leader CTA) ==================
    const uint32_t idesc = kMx ? umma_idesc_mxfp8(256, TN, p.a_fmt, p.b_fmt)
                           : kFp8 ? umma_idesc_fp8(256, TN, p.a_mn, p.b_mn, p.a_fmt, p.b_fmt)
                                  : umma_idesc_bf16(256, TN, p.a_mn, p.b_mn);

    const uint32_t a_lbo = p.a_mn ? kChunk : 16, b_lbo = p.b_mn ? kChunk : 16;
    const uint32_t a_kadv = p.a_mn ? UK * 128 : 32;   // MN-major: UK k-rows of 128 B; K-major: 32 B
    const uint32_t b_kadv = p.b_mn ? UK * 128 : 32;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = cluster_id; t < p.num_tiles; t += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TN);
      // MXFP8: this tile's scale factors sit in the first 64 columns of the other accumulator (see the copier)
      const uint32_t tsf0 = tmem_base + (uint32_t)((acc ^ 1) * TN);
      for (int kb = 0; kb < num_kb; ++kb) {
        if constexpr (kMx) mbar_wait(sfready_bar(stage), phase);     // implies full_bar(stage)
        else mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = a_smem + stage * kAStage2;
        const uint32_t sb = b_smem + stage * kBStage2;
        const uint32_t tsfa = tsf0 + (uint32_t)(stage * 12), tsfb = tsfa + 4u;
#pragma unroll
        for (int k = 0; k < BKE / UK; ++k) {
          const uint64_t ad = umma_smem_desc(sa + k * a_kadv, a_lbo, 1024);
          const uint64_t bd = umma_smem_desc(sb + k * b_kadv, b_lbo, 1024);
          if constexpr (kMx) umma_mxfp8_ss_2cta(d_tmem, ad, bd, idesc | ((uint32_t)k << 29) | ((uint32_t)k << 4), tsfa, tsfb,
                                                (kb | k) != 0 ? 1u : 0u);
          else if constexpr (kFp8) umma_fp8_ss_2cta(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          else umma_bf16_ss_2cta(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit_2cta(empty_bar(stage), 0x3);   // frees the slot in BOTH CTAs
        if (++stage == kStages2) { stage = 0; phase ^= 1u; }
      }
      umma_commit_2cta(tfull_bar(acc), 0x3);       // accumulators ready in BOTH CTAs
    }
  } else if (warp >= 4) {
    // ============================ epilogue (both CTAs) =====================
    const int ew = warp - 4;
    const uint32_t epi_smem = smem_base + C2::epi_off + ew * kEB * kEpiWarpBuf;
    const float out_scale = (kFp8 && !kMx && p.scale != nullptr) ? __ldg(p.scale) * __ldg(p.scale_b) : 1.f;
    int it = 0;
    int buf = 0;
    for (int t = cluster_id; t < p.num_tiles; t += num_clusters, ++it) {
      int tm, tn;
      tile_coords2(t, p, tm, tn);
      const int m0 = tm * 256 + (int)cta_rank * 128, n0 = tn * TN;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * TN);
      const int row = m0 + ew * 32 + lane;
uint32_t actp[32];
      #pragma unroll 2
      for (int c = 0; c < TN / 64; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_addr + c * 64, r0);
        tmem_ld_32x32(t_addr + c * 64 + 32, r1);
        tmem_ld_wait();
        if (kMx && c == 0) {
          // the next tile's scale factors go into columns [0, 64) of this accumulator: they are in registers now
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_shared(sffree_bar(acc), 0));
        }
        if (c == TN / 64 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_shared(tempty_bar(acc), 0));
        }
        const int col0 = n0 + c * 64;
        if (p.gu != nullptr) {
          // Fused SwiGLU backward (down-projection dgrad): this chunk holds d(act) for 64 activations of 32 rows;
          // with the (gate, up) pairs read straight from gu (256 contiguous bytes per row and chunk) it becomes
          // 128 interleaved d(gate), d(up) columns -- two 64-column stores through tma_e.  d(act) is rounded to
          // bf16 first, so the result equals the stand-alone `swiglu_bwd_kernel` applied to the bf16 dgrad.
          const bool row_ok = row < p.M;
          const bf16* gp = p.gu + (size_t)row * p.ld_gu + 2 * (size_t)col0;
          if (row_ok && c + 1 < TN / 64 && 2 * (col0 + 64) + 128 <= 2 * p.N) {
            // the next chunk's (gate, up) pairs: start them on their way from HBM to L2 now
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + 128));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + 192));
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t outp[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {                       // 8 x 16 bytes of gu = 4 pairs each
              float a[8];
              bf16x8 gv = bf16x8{};
              if (row_ok && 2 * col0 + hh * 64 + j * 8 + 8 <= 2 * p.N) gv = ld_bf16x8(gp + hh * 64 + j * 8);
              unpack8(gv, a);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = hh * 32 + j * 4 + e;              // activation index inside the chunk (0..63)
                const float dacc = __uint_as_float(i < 32 ? r0[i & 31] : r1[i & 31]) * out_scale;
                const float d = __bfloat162float(__float2bfloat16_rn(dacc));
                const float g = a[2 * e], u = a[2 * e + 1];
                const float sg = __fdividef(1.f, 1.f + __expf(-g));
                const float silu = g * sg;
                __nv_bfloat162 h = __floats2bfloat162_rn(d * u * (sg + silu * (1.f - sg)), d * silu);
                outp[j * 4 + e] = *reinterpret_cast<uint32_t*>(&h);
              }
            }
            if (lane == 0) tma_store_wait_read<kEB - 1>();
            __syncwarp();
            const uint32_t sbuf = epi_smem + buf * kEpiWarpBuf;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t dst = sbuf + lane * 128 + ((j ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(outp[4 * j]),
                           "r"(outp[4 * j + 1]), "r"(outp[4 * j + 2]), "r"(outp[4 * j + 3])
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (col0 < p.N && m0 + ew * 32 < p.M) tma_store_2d(&tma_e, sbuf, 2 * col0 + hh * 64, m0 + ew * 32);
              tma_store_commit();
            }
            buf = (kEB == 2) ? (buf ^ 1) : 0;
          }
          continue;
        }
        uint32_t packed[32];
        if (p.accumulate) {
          const bool row_ok = row < p.M;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 prev = make_uint4(0, 0, 0, 0);
            if (row_ok && col0 + j * 8 + 8 <= p.N)
              prev = *reinterpret_cast<const uint4*>(p.D + (size_t)row * p.ldd + col0 + j * 8);
            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = j * 8 + e * 2;
              const float2 o = __bfloat1622float2(pv[e]);
              const float x0 = __uint_as_float(i < 32 ? r0[i & 31] : r1[i & 31]) * out_scale + o.x;
              const float x1 = __uint_as_float(i + 1 < 32 ? r0[(i + 1) & 31] : r1[(i + 1) & 31]) * out_scale + o.y;
              __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
              packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(r0[2 * i]) * out_scale,
                                                      __uint_as_float(r0[2 * i + 1]) * out_scale);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(r1[2 * i]) * out_scale,
                                                      __uint_as_float(r1[2 * i + 1]) * out_scale);
            packed[i] = *reinterpret_cast<uint32_t*>(&h0);
            packed[16 + i] = *reinterpret_cast<uint32_t*>(&h1);
          }
        }
        if (lane == 0) tma_store_wait_read<kEB - 1>();
        __syncwarp();
        const uint32_t sbuf = epi_smem + buf * kEpiWarpBuf;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dst = sbuf + lane * 128 + ((j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(packed[4 * j]),
                       "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (col0 < p.N && m0 + ew * 32 < p.M) tma_store_2d(&tma_d, sbuf, col0, m0 + ew * 32);
          tma_store_commit();
        }
        buf = (kEB == 2) ? (buf ^ 1) : 0;
        if (p.swiglu) {
          // Fused SwiGLU: the 64 columns of this chunk are 32 (gate, up) pairs.  The activation is computed from
          // the bf16-ROUNDED pair, i.e. bit-for-bit what `swiglu_fwd_kernel` re-derives from `gu` in backward.
          // Two chunks make one 64-column (128-byte) row of the second output; it leaves through the same
          // two staging buffers as the main output (a buffer is rewritten once the store before last has read it).
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&packed[i]));
            const float2 b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&packed[i + 1]));
            const float o0 = a.x * __fdividef(1.f, 1.f + __expf(-a.x)) * a.y;
            const float o1 = b2.x * __fdividef(1.f, 1.f + __expf(-b2.x)) * b2.y;
            __nv_bfloat162 h = __floats2bfloat162_rn(o0, o1);
            actp[(c & 1) * 16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&h);
          }
          if (c & 1) {
            if (lane == 0) tma_store_wait_read<kEB - 1>();
            __syncwarp();
            const uint32_t abuf = epi_smem + buf * kEpiWarpBuf;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t dst = abuf + lane * 128 + ((j ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(actp[4 * j]),
                           "r"(actp[4 * j + 1]), "r"(actp[4 * j + 2]), "r"(actp[4 * j + 3])
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int acol = (n0 + (c - 1) * 64) >> 1;
              if (2 * acol < p.N && m0 + ew * 32 < p.M) tma_store_2d(&tma_e, abuf, acol, m0 + ew * 32);
              tma_store_commit();
            }
            buf = (kEB == 2) ? (buf ^ 1) : 0;
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::string g_err;
int g_mode = 2;   // 2 = cta_group::2 CTA-pair kernel (default, faster), 1 = single-CTA kernel
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;
int g_num_sms = 0;

void init_once() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) ==
          cudaSuccess && q == cudaDriverEntryPointSuccess) {
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       kSmemBytes);
  cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       Cfg2<0>::smem_bytes);
  cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       Cfg2<1>::smem_bytes);
  cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       Cfg2<2>::smem_bytes);
  if (const char* e = getenv("HYPHA_GEMM_CTAS")) g_mode = (e[0] == '1') ? 1 : 2;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of `ld` elements.
bool make_tmap(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
               uint32_t box_inner, uint32_t box_outer, int elem_bytes = 2) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                        2, const_cast<void*>(ptr), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
    return false;
  }
  return true;
}

// Scale-factor table of an MXFP8 operand: `rows` = (128-row blocks) x (128-element K blocks), 512 bytes each.
bool make_tmap_sf(CUtensorMap* m, const void* ptr, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {128, rows};
  cuuint64_t strides[1] = {512};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_err = "cuTensorMapEncodeTiled (scale factors) failed with code " + std::to_string((int)r);
    return false;
  }
  return true;
}

}  // namespace

const char* gemm_last_error() { return g_err.c_str(); }
void gemm_set_cta_mode(int mode) { std::call_once(g_once, init_once); g_mode = (mode == 1) ? 1 : 2; }
int gemm_get_cta_mode() { std::call_once(g_once, init_once); return g_mode; }

int gemm_bf16_tcgen05(const bf16* A, const bf16* B, bf16* D, int M, int N, int K, int lda,
                      int ldb, int ldd, int a_kmajor, int b_kmajor, int accumulate,
                      cudaStream_t s, bf16* act, int ld_act, const bf16* gu, int ld_gu) {
  std::call_once(g_once, init_once);
  if (g_encode == nullptr) { g_err = "cuTensorMapEncodeTiled unavailable"; return 1; }
  if (M <= 0 || N <= 0 || K <= 0) { g_err = "empty problem"; return 2; }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al16(A) || !al16(B) || !al16(D) || (lda % 8) || (ldb % 8) || (ldd % 8) || (N % 8)) {
    g_err = "operands must be 16-byte aligned with leading dimensions and N multiples of 8";
    return 3;
  }
  if ((a_kmajor || b_kmajor) && (K % 8)) { g_err = "K must be a multiple of 8"; return 3; }
  if (!a_kmajor && (M % 8)) { g_err = "MN-major A needs M % 8 == 0"; return 3; }

  if (act != nullptr && gu == nullptr && (accumulate || !al16(act) || (ld_act % 8) || (N % 128))) {
    g_err = "swiglu epilogue: needs N % 128 == 0, a 16-byte aligned second output and no accumulation";
    return 3;
  }
  if (gu != nullptr && (act == nullptr || accumulate || !al16(gu) || (ld_gu % 8) || (N % 64))) {
    g_err = "dswiglu epilogue: needs the [M, 2N] output in `act`, N % 64 == 0, 16-byte aligned gu and no accumulation";
    return 3;
  }
  const bool two = (g_mode == 2) || act != nullptr;       // the fused epilogues live in the 2-CTA kernel
  const uint32_t b_rows = two ? 128 : BN;       // rows of the K-major B box each CTA loads
  CUtensorMap ta, tb, td, te;
  bool ok = a_kmajor ? make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)
                     : make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  ok = ok && (b_kmajor ? make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, b_rows)
                       : make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK));
  ok = ok && make_tmap(&td, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd, 64, 32);
  ok = ok && (act != nullptr ? make_tmap(&te, act, (uint64_t)(gu != nullptr ? 2 * N : N / 2), (uint64_t)M, (uint64_t)ld_act, 64, 32)
                             : make_tmap(&te, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd, 64, 32));
  if (!ok) return 4;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_kmajor ? 0 : 1;
  p.b_mn = b_kmajor ? 0 : 1;
  p.accumulate = accumulate;
  p.D = D; p.ldd = ldd;
  p.swiglu = (act != nullptr && gu == nullptr) ? 1 : 0;
  p.gu = gu; p.ld_gu = ld_gu;
  p.a_fmt = p.b_fmt = 0; p.scale = nullptr; p.scale_b = nullptr;
  p.tiles_n = (N + BN - 1) / BN;
  if (two) {
    p.tiles_m = (M + 255) / 256;
    p.num_tiles = p.tiles_m * p.tiles_n;
    int clusters = g_num_sms / 2;
    if (p.num_tiles < clusters) clusters = p.num_tiles;
    gemm_tcgen05_2cta_kernel<0><<<clusters * 2, kThreads, Cfg2<0>::smem_bytes, s>>>(ta, tb, td, te, td, td, p);
  } else {
    p.tiles_m = (M + BM - 1) / BM;
    p.num_tiles = p.tiles_m * p.tiles_n;
    const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
    gemm_bf16_tcgen05_kernel<<<grid, kThreads, kSmemBytes, s>>>(ta, tb, td, p);
  }
  count_launch();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return 5; }
  return 0;
}


// D[M,N] (+)= scale * (A . B) with fp8 (e4m3 / e5m2) operands, fp32 accumulation, bf16 output.
int gemm_fp8_tcgen05(const uint8_t* A, const uint8_t* B, bf16* D, int M, int N, int K, int lda,
                     int ldb, int ldd, int a_kmajor, int b_kmajor, int a_fmt, int b_fmt,
                     const float* dequant_a, const float* dequant_b, int accumulate, cudaStream_t s,
                     bf16* act, int ld_act) {
  std::call_once(g_once, init_once);
  if (g_encode == nullptr) { g_err = "cuTensorMapEncodeTiled unavailable"; return 1; }
  if (M <= 0 || N <= 0 || K <= 0) { g_err = "empty problem"; return 2; }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al16(A) || !al16(B) || !al16(D) || (lda % 16) || (ldb % 16) || (ldd % 8) || (N % 16) ||
      (K % 16) || (!a_kmajor && (M % 16))) {
    g_err = "fp8 operands need 16-byte aligned pointers / leading dimensions and M,N,K % 16 == 0";
    return 3;
  }
  if (act != nullptr && (accumulate || !al16(act) || (ld_act % 8) || (N % 128))) {
    g_err = "swiglu epilogue: needs N % 128 == 0, a 16-byte aligned second output and no accumulation";
    return 3;
  }
  CUtensorMap ta, tb, td, te;
  bool ok = a_kmajor ? make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 128, 128, 1)
                     : make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 128, 1);
  ok = ok && (b_kmajor ? make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 128, 128, 1)
                       : make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 128, 128, 1));
  ok = ok && make_tmap(&td, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd, 64, 32);
  ok = ok && (act != nullptr ? make_tmap(&te, act, (uint64_t)(N / 2), (uint64_t)M, (uint64_t)ld_act, 64, 32)
                             : make_tmap(&te, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd, 64, 32));
  if (!ok) return 4;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_kmajor ? 0 : 1;
  p.b_mn = b_kmajor ? 0 : 1;
  p.accumulate = accumulate;
  p.D = D; p.ldd = ldd;
  p.swiglu = act != nullptr ? 1 : 0;
  p.gu = nullptr; p.ld_gu = 0;
  p.a_fmt = a_fmt; p.b_fmt = b_fmt; p.scale = dequant_a; p.scale_b = dequant_b;
  p.tiles_n = (N + BN - 1) / BN;
  p.tiles_m = (M + 255) / 256;
  p.num_tiles = p.tiles_m * p.tiles_n;
  int clusters = g_num_sms / 2;
  if (p.num_tiles < clusters) clusters = p.num_tiles;
  gemm_tcgen05_2cta_kernel<1><<<clusters * 2, kThreads, Cfg2<1>::smem_bytes, s>>>(ta, tb, td, te, td, td, p);
  count_launch();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return 5; }
  return 0;
}


// D[M,N] (+)= A . B^T with MXFP8 operands: A [M,K] and B [N,K] fp8 (K contiguous), one UE8M0 scale per 32 consecutive
// K-elements of every row, the scales in the tensor-core atom layout (see `quantize_mxfp8`): table[K block of 128]
// [row block of 128][(r % 32) * 16 + (r / 32 % 4) * 4 + (k / 32 % 4)].  fp32 accumulation, bf16 output.
int gemm_mxfp8_tcgen05(const uint8_t* A, const uint8_t* sfa, const uint8_t* B, const uint8_t* sfb, bf16* D, int M, int N,
                       int K, int lda, int ldb, int ldd, int a_fmt, int b_fmt, int accumulate, cudaStream_t s,
                       bf16* act, int ld_act, const bf16* gu, int ld_gu) {
  std::call_once(g_once, init_once);
  if (g_encode == nullptr) { g_err = "cuTensorMapEncodeTiled unavailable"; return 1; }
  if (M <= 0 || N <= 0 || K <= 0) { g_err = "empty problem"; return 2; }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al16(A) || !al16(B) || !al16(D) || !al16(sfa) || !al16(sfb) || (lda % 16) || (ldb % 16) || (ldd % 8) ||
      (M % 128) || (N % 128) || (K % 128)) {
    g_err = "mxfp8 operands need 16-byte aligned pointers / leading dimensions and M, N, K % 128 == 0";
    return 3;
  }
  if (act != nullptr && gu == nullptr && (accumulate || !al16(act) || (ld_act % 8))) {
    g_err = "swiglu epilogue: needs a 16-byte aligned second output and no accumulation";
    return 3;
  }
  if (gu != nullptr && (act == nullptr || accumulate || !al16(gu) || (ld_gu % 8))) {
    g_err = "dswiglu epilogue: needs the [M, 2N] output in `act`, 16-byte aligned gu and no accumulation";
    return 3;
  }
  constexpr int TN = Cfg2<2>::TN;
  const int num_kb = K / 128;
  CUtensorMap ta, tb, td, te, tsa, tsb;
  bool ok = make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 128, 128, 1);
  ok = ok && make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 128, TN / 2, 1);
  ok = ok && make_tmap(&td, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd, 64, 32);
  ok = ok && (act != nullptr ? make_tmap(&te, act, (uint64_t)(gu != nullptr ? 2 * N : N / 2), (uint64_t)M, (uint64_t)ld_act, 64, 32)
                             : make_tmap(&te, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd, 64, 32));
  ok = ok && make_tmap_sf(&tsa, sfa, (uint64_t)(M / 128) * num_kb, 1) && make_tmap_sf(&tsb, sfb, (uint64_t)(N / 128) * num_kb, 2);
  if (!ok) return 4;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = 0; p.b_mn = 0;
  p.accumulate = accumulate;
  p.D = D; p.ldd = ldd;
  p.swiglu = (act != nullptr && gu == nullptr) ? 1 : 0;
  p.gu = gu; p.ld_gu = ld_gu;
  p.a_fmt = a_fmt; p.b_fmt = b_fmt; p.scale = nullptr; p.scale_b = nullptr;
  p.tiles_n = (N + TN - 1) / TN;
  p.tiles_m = (M + 255) / 256;
  p.num_tiles = p.tiles_m * p.tiles_n;
  int clusters = g_num_sms / 2;
  if (p.num_tiles < clusters) clusters = p.num_tiles;
  gemm_tcgen05_2cta_kernel<2><<<clusters * 2, kThreads, Cfg2<2>::smem_bytes, s>>>(ta, tb, td, te, tsa, tsb, p);
  count_launch();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return 5; }
  return 0;
}

}  // namespace hb
