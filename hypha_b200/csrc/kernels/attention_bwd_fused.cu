// Causal GQA flash-attention BACKWARD, single fused kernel (sm_100a, head_dim 128): 5 GEMMs per tile pair.
//
//   P^T  = exp2(scale_log2 * K Q^T - LSE2)        dV += P^T dO
//   dP^T = V dO^T                                  dS^T = P^T o (dP^T - D) * scale
//   dK  += dS^T Q                                  dQ  += dS K      (fp32, reduced across key tiles in L2)
//
// The two-kernel variant in attention_bwd_tcgen05.cu recomputes S and dP twice (7 GEMMs) to stay
// atomics-free; this kernel computes them once per (key tile, q tile) pair and hands the dQ partial to the
// memory system instead: the accumulator tile leaves TMEM through registers and shared memory as a
// `cp.reduce.async.bulk.tensor ... .add` (TMA reduce) into an fp32 buffer that a tiny epilogue kernel
// converts into the fused d(qkv).  Everything is 128x128 tiles, so every MMA is 128x128x16.
//
// One CTA = one 128-key tile j of one (batch, kv head); it streams the 128-row Q_i / dO_i tiles (i >= j) of
// all q heads of the GQA group.  TMEM (512 columns): dK 0..127 | dV 128..255 | dP / dQ 256..383 | S / P 384..511.
//
//   warps 0-3   dQ drain: tcgen05.ld 128 columns -> release TMEM -> 4 x (st.shared 128B-swizzled -> TMA reduce-add)
//   warps 4-11  softmax: thread = key row (TMEM lane), warpgroup = 64 of the 128 q columns.
//               S -> P (fp32 kept in registers) -> bf16 P^T back into TMEM over S (A operand of dV, TS form)
//               dP -> dS^T = P (dP - D) scale -> bf16 -> swizzled smem (K-major A of dK, MN-major A of dQ)
//   warp 12     tcgen05.mma issuer (one thread).  Issue order per q tile (tensor pipe executes in order):
//                  S_i | dQ_{i-1} | dK_{i-1} | dP_i | dV_i
//               so the exponentials of tile i run under dQ/dK of tile i-1 and the dS arithmetic under dV_i / S_{i+1}
//   warp 13     TMA producer: K_j, V_j once; Q_i (+ LSE2_i, D_i) 2 stages; dO_i 1 stage
//
// Replaces the stock attention backward the reference reaches through `accelerator.backward(loss)`
// (executors/accelerate/src/hypha/accelerate_executor/training.py:110-112).
#include <mutex>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace hb {
namespace {
using namespace sm100;

constexpr int kD = 128;
constexpr int kT = 128;                        // tile edge (keys per CTA, q rows per iteration)
constexpr int kTileBytes = kT * kD * 2;        // 32 KB: two 64-column swizzle atoms of [128 rows x 128 B]
constexpr int kAtomBytes = kT * 128;           // 16 KB
constexpr int kStageBytes = kT * 32 * 4;       // 16 KB: 128 rows x 32 fp32 columns of dQ
constexpr int kFusedThreads = 512;

struct FSmem {
  static constexpr int k_off = 0;
  static constexpr int v_off = k_off + kTileBytes;
  static constexpr int q_off = v_off + kTileBytes;            // 2 stages
  static constexpr int do_off = q_off + 2 * kTileBytes;       // 1 stage
  static constexpr int ds_off = do_off + kTileBytes;
  static constexpr int stg_off = ds_off + kTileBytes;         // 2 x 16 KB dQ staging
  static constexpr int vec_off = stg_off + 2 * kStageBytes;   // 2 stages x (128 lse2 + 128 D) floats
  static constexpr int bar_off = vec_off + 2 * 1024;
  static constexpr int total = bar_off + 128;
};
constexpr int kMaxDynSmem = 232448;            // 227 KB
constexpr int kFusedSmemBytes = FSmem::total + 768 <= kMaxDynSmem ? FSmem::total + 768 : kMaxDynSmem;
static_assert(FSmem::total <= kMaxDynSmem, "fused attention backward: shared memory plan does not fit");

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void sts16(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// smem tile -> global, element-wise fp32 add performed by the memory system (TMA reduce)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct FusedParams {
  int batch, seq, n_heads, n_kv;
  int dqkv_stride;
  float scale, scale_log2;
  const float* lse2;      // [batch, n_heads, seq]
  const float* delta;     // [batch, n_heads, seq]
  bf16* dqkv;
};

// D = rowsum(dO o O), lse2 = lse * log2(e), and the fp32 dQ accumulator is zeroed.  Eight lanes per (token, head)
// row: 32 bytes of O and dO per lane (two 16-byte loads each, 4 rows in flight per warp) and 64 bytes of zeros.
__global__ void __launch_bounds__(256) attn_bwd_fused_prep_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout,
                                                                  const float* __restrict__ lse, float* __restrict__ delta,
                                                                  float* __restrict__ lse2, float* __restrict__ dq_acc,
                                                                  int tokens, int seq, int n_heads, int out_stride,
                                                                  int dout_stride) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid >> 3;                                    // (token, head)
  const int sub = (int)(gid & 7);                                  // 16 of the 128 columns
  const bool live = row < (int64_t)tokens * n_heads;
  float acc = 0.f;
  int t = 0, h = 0;
  if (live) {
    t = (int)(row / n_heads);
    h = (int)(row % n_heads);
    const bf16* o = out + (size_t)t * out_stride + h * kD + sub * 16;
    const bf16* g = dout + (size_t)t * dout_stride + h * kD + sub * 16;
    const bf16x8 o0 = ld_stream_bf16x8(o), o1 = ld_stream_bf16x8(o + 8);
    const bf16x8 g0 = ld_stream_bf16x8(g), g1 = ld_stream_bf16x8(g + 8);
    float fo[16], fg[16];
    unpack8(o0, fo); unpack8(o1, fo + 8);
    unpack8(g0, fg); unpack8(g1, fg + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc = fmaf(fo[i], fg[i], acc);
    float4* z = reinterpret_cast<float4*>(dq_acc + (size_t)row * kD + sub * 16);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    z[0] = zero; z[1] = zero; z[2] = zero; z[3] = zero;
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (live && sub == 0) {
    const int b = t / seq, pos = t % seq;
    const size_t idx = ((size_t)b * n_heads + h) * seq + pos;
    delta[idx] = acc;
    lse2[idx] = lse[idx] * 1.4426950408889634f;
  }
}

// fp32 dQ accumulator -> bf16 q columns of d(qkv)
__global__ void __launch_bounds__(256) attn_bwd_dq_convert_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv,
                                                                  int64_t n8, int row_elems8, int dqkv_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int64_t row = i / row_elems8;
  const int c8 = (int)(i % row_elems8);
  const float4 a = ld_stream_f32x4(dq_acc + i * 8), b = ld_stream_f32x4(dq_acc + i * 8 + 4);
  const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  st_bf16x8(dqkv + (size_t)row * dqkv_stride + c8 * 8, pack8(f));
}

__global__ void __launch_bounds__(kFusedThreads, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tma_qkv,     // qkv bf16, box {64, 128}
                      const __grid_constant__ CUtensorMap tma_do,      // dO  bf16, box {64, 128}
                      const __grid_constant__ CUtensorMap tma_dq,      // dQ accumulator fp32, box {32, 128}
                      const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (smem_base - smem_u32(smem_raw) + FSmem::total > (uint32_t)kFusedSmemBytes) __trap();   // base less aligned than planned
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t k_smem = smem_base + FSmem::k_off, v_smem = smem_base + FSmem::v_off;
  const uint32_t q_smem = smem_base + FSmem::q_off, do_smem = smem_base + FSmem::do_off;
  const uint32_t ds_smem = smem_base + FSmem::ds_off, stg_smem = smem_base + FSmem::stg_off;
  const uint32_t vec_smem = smem_base + FSmem::vec_off;
  const float* vec_gen = reinterpret_cast<const float*>(smem_gen + FSmem::vec_off);
  const uint32_t bars = smem_base + FSmem::bar_off;
  const uint32_t kv_full = bars;
  auto q_full = [&](int s) { return bars + 8u * (1 + s); };
  auto q_empty = [&](int s) { return bars + 8u * (3 + s); };
  const uint32_t do_full = bars + 8u * 5, do_empty = bars + 8u * 6;
  const uint32_t s_full = bars + 8u * 7, p_full = bars + 8u * 8, dp_full = bars + 8u * 9;
  const uint32_t ds_full = bars + 8u * 10, ds_empty = bars + 8u * 11;
  const uint32_t dq_full = bars + 8u * 12, dq_free = bars + 8u * 13, acc_full = bars + 8u * 14;
  const uint32_t tmem_ptr_smem = bars + 8u * 15;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + FSmem::bar_off + 8 * 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq_tiles = p.seq / kT;
  const int per_tile = p.batch * p.n_kv;
  const int j = (int)(blockIdx.x / per_tile);                      // early key tiles (most q tiles) first
  const int bh = (int)(blockIdx.x % per_tile);
  const int b = bh / p.n_kv, hk = bh % p.n_kv;
  const int G = p.n_heads / p.n_kv;
  const int n_q = nq_tiles - j;
  const int n_it = G * n_q;
  const int krow0 = b * p.seq + j * kT;
  const int k_col = (p.n_heads + hk) * kD, v_col = (p.n_heads + p.n_kv + hk) * kD;

  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    tma_prefetch_desc(&tma_dq);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    mbar_init(do_full, 1); mbar_init(do_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(dp_full, 1);
    mbar_init(ds_full, 8); mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1); mbar_init(dq_free, 4); mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t tmem_dk = tmem_base, tmem_dv = tmem_base + 128, tmem_dp = tmem_base + 256, tmem_s = tmem_base + 384;

  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 13 && lane == 0) {
      // ====================== TMA producer ======================
      mbar_arrive_expect_tx(kv_full, 2 * kTileBytes);
      tma_load_2d(k_smem, &tma_qkv, kv_full, k_col, krow0);
      tma_load_2d(k_smem + kAtomBytes, &tma_qkv, kv_full, k_col + 64, krow0);
      tma_load_2d(v_smem, &tma_qkv, kv_full, v_col, krow0);
      tma_load_2d(v_smem + kAtomBytes, &tma_qkv, kv_full, v_col + 64, krow0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const int h = hk * G + it / n_q, i = j + it % n_q;
        const int qrow = b * p.seq + i * kT;
        const size_t vbase = ((size_t)b * p.n_heads + h) * p.seq + (size_t)i * kT;
        mbar_wait(q_empty(st), ((uint32_t)(it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full(st), kTileBytes + 2 * kT * 4);
        bulk_load_1d(vec_smem + st * 1024, p.lse2 + vbase, kT * 4, q_full(st));
        bulk_load_1d(vec_smem + st * 1024 + 512, p.delta + vbase, kT * 4, q_full(st));
        tma_load_2d(q_smem + st * kTileBytes, &tma_qkv, q_full(st), h * kD, qrow);
        tma_load_2d(q_smem + st * kTileBytes + kAtomBytes, &tma_qkv, q_full(st), h * kD + 64, qrow);
        mbar_wait(do_empty, ((uint32_t)it & 1u) ^ 1u);
        mbar_arrive_expect_tx(do_full, kTileBytes);
        tma_load_2d(do_smem, &tma_do, do_full, h * kD, qrow);
        tma_load_2d(do_smem + kAtomBytes, &tma_do, do_full, h * kD + 64, qrow);
      }
    } else if (warp == 12) {
      // ====================== MMA issuer: the whole warp walks the loop, one elected lane issues ========================
      const uint32_t idesc_kk = umma_idesc_bf16(kT, kT, 0, 0);     // S^T = K Q^T, dP^T = V dO^T : K-major x K-major
      const uint32_t idesc_kn = umma_idesc_bf16(kT, kD, 0, 1);     // dV += P^T dO, dK += dS^T Q : B MN-major
      const uint32_t idesc_nn = umma_idesc_bf16(kT, kD, 1, 1);     // dQ = dS K : A (dS^T read transposed) and B MN-major
      // descriptors are built once; per MMA only the 14-bit start-address field (16-byte units) moves
      const uint64_t dK_k = umma_smem_desc(k_smem, 16, 1024), dK_n = umma_smem_desc(k_smem, kAtomBytes, 1024);
      const uint64_t dV_k = umma_smem_desc(v_smem, 16, 1024);
      const uint64_t dDO_k = umma_smem_desc(do_smem, 16, 1024), dDO_n = umma_smem_desc(do_smem, kAtomBytes, 1024);
      const uint64_t dDS_k = umma_smem_desc(ds_smem, 16, 1024), dDS_n = umma_smem_desc(ds_smem, kAtomBytes, 1024);
      const uint64_t dQ_k0 = umma_smem_desc(q_smem, 16, 1024), dQ_n0 = umma_smem_desc(q_smem, kAtomBytes, 1024);
      auto koff = [](int kk) { return (uint64_t)(((kk >> 2) * kAtomBytes + (kk & 3) * 32) >> 4); };   // K-major 16-element k-step
      auto noff = [](int kk) { return (uint64_t)((kk * 16 * 128) >> 4); };                               // MN-major: 16 k-rows
      auto issue_dq_dk = [&](int it) {                             // consumes dS_it, Q_it, K_j
        const uint64_t qn = dQ_n0 + (uint64_t)((it & 1) * (kTileBytes >> 4));
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kT / 16; ++kk)
            umma_bf16_ss(tmem_dp, dDS_n + noff(kk), dK_n + noff(kk), idesc_nn, kk != 0 ? 1u : 0u);
          umma_commit(dq_full);
#pragma unroll
          for (int kk = 0; kk < kT / 16; ++kk)
            umma_bf16_ss(tmem_dk, dDS_k + koff(kk), qn + noff(kk), idesc_kn, (it != 0 || kk != 0) ? 1u : 0u);
          umma_commit(q_empty(it & 1));
          umma_commit(ds_empty);
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const uint64_t qk = dQ_k0 + (uint64_t)(st * (kTileBytes >> 4));
        mbar_wait(q_full(st), (uint32_t)(it >> 1) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk)
            umma_bf16_ss(tmem_s, dK_k + koff(kk), qk + koff(kk), idesc_kk, kk != 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
        if (it > 0) {
          mbar_wait(ds_full, (uint32_t)(it - 1) & 1u);
          tc_fence_after();
          issue_dq_dk(it - 1);
          mbar_wait(dq_free, (uint32_t)(it - 1) & 1u);             // dQ_{it-1} has left TMEM: dP_it may overwrite it
        }
        mbar_wait(do_full, (uint32_t)it & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk)
            umma_bf16_ss(tmem_dp, dV_k + koff(kk), dDO_k + koff(kk), idesc_kk, kk != 0 ? 1u : 0u);
          umma_commit(dp_full);
        }
        __syncwarp();
        mbar_wait(p_full, (uint32_t)it & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kT / 16; ++kk) {
            // P^T (bf16 pairs) sits in the S columns each warpgroup just consumed: q 0..63 -> cols 0..31, q 64..127 -> cols 64..95
            const uint32_t a_tmem = tmem_s + (uint32_t)(kk >> 2) * 64 + (uint32_t)(kk & 3) * 8;
            umma_bf16_ts(tmem_dv, a_tmem, dDO_n + noff(kk), idesc_kn, (it != 0 || kk != 0) ? 1u : 0u);
          }
          umma_commit(do_empty);
        }
        __syncwarp();
      }
      mbar_wait(ds_full, (uint32_t)(n_it - 1) & 1u);
      tc_fence_after();
      issue_dq_dk(n_it - 1);
      if (elect_one_sync()) umma_commit(acc_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ====================== softmax warpgroups ==========
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = q4 * 32 + lane;                                   // key row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const uint32_t ds_row = ds_smem + (uint32_t)half * kAtomBytes + (uint32_t)r * 128;
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const bool diag = (it % n_q) == 0;                            // q tile i == key tile j: causal boundary
      const float4* lse4 = reinterpret_cast<const float4*>(vec_gen + st * 256 + half * 64);
      const float4* d4 = reinterpret_cast<const float4*>(vec_gen + st * 256 + 128 + half * 64);
      mbar_wait(q_full(st), (uint32_t)(it >> 1) & 1u);              // lse2 / D of this q tile staged
      mbar_wait(s_full, (uint32_t)it & 1u);
      tc_fence_after();
      float pr[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32(tmem_s + lane_addr + half * 64 + c * 32, sv);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 l = lse4[(c * 32 + e) >> 2];
          const float ls[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            float pe = ex2f(fmaf(__uint_as_float(sv[e + x]), p.scale_log2, -ls[x]));
            if (diag && r > half * 64 + c * 32 + e + x) pe = 0.f;
            pr[c * 32 + e + x] = pe;
          }
          pk[(e >> 1)] = pack2(pr[c * 32 + e], pr[c * 32 + e + 1]);
          pk[(e >> 1) + 1] = pack2(pr[c * 32 + e + 2], pr[c * 32 + e + 3]);
        }
        tmem_st_32x16(tmem_s + lane_addr + half * 64 + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);

      mbar_wait(dp_full, (uint32_t)it & 1u);
      tc_fence_after();
      if (it > 0) mbar_wait(ds_empty, (uint32_t)(it - 1) & 1u);     // dQ/dK MMAs of tile it-1 have read dS
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t dv[32];
        tmem_ld_32x32(tmem_dp + lane_addr + half * 64 + c * 32, dv);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          const float4 da = d4[(c * 32 + e) >> 2], db = d4[((c * 32 + e) >> 2) + 1];
          const float dd[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
          uint32_t o[4];
#pragma unroll
          for (int x = 0; x < 8; x += 2) {
            const float a0 = pr[c * 32 + e + x] * (__uint_as_float(dv[e + x]) - dd[x]) * p.scale;
            const float a1 = pr[c * 32 + e + x + 1] * (__uint_as_float(dv[e + x + 1]) - dd[x + 1]) * p.scale;
            o[x >> 1] = pack2(a0, a1);
          }
          const uint32_t chunk = (uint32_t)(c * 4 + (e >> 3));
          sts16(ds_row + ((chunk ^ (uint32_t)(r & 7)) << 4), o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    // ---- epilogue: dK, dV -> bf16 -> d(qkv); each warpgroup owns 64 of the 128 columns of both ----
    mbar_wait(acc_full, 0);
    tc_fence_after();
    bf16* row = p.dqkv + (size_t)(krow0 + r) * p.dqkv_stride;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        bf16* dst = row + (which == 0 ? v_col : k_col) + half * 64 + c * 32;
        uint32_t t[32];
        tmem_ld_32x32((which == 0 ? tmem_dv : tmem_dk) + lane_addr + half * 64 + c * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          float f[8];
#pragma unroll
          for (int x = 0; x < 8; ++x) f[x] = __uint_as_float(t[e + x]);
          st_bf16x8(dst + e, pack8(f));
        }
      }
    }
  } else {
    // ====================== dQ drain warpgroup ==========
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int r = warp * 32 + lane;                                 // q row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int it = 0; it < n_it; ++it) {
      const int h = hk * G + it / n_q, i = j + it % n_q;
      mbar_wait(dq_full, (uint32_t)it & 1u);
      tc_fence_after();
      uint32_t acc[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_dp + lane_addr + c * 32, acc[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t buf = stg_smem + (uint32_t)(c & 1) * kStageBytes;
        if (threadIdx.x == 0) tma_store_wait_read<1>();             // the reduce that last used this buffer has read it
        named_bar_sync(1, 128);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sts16(buf + (uint32_t)r * 128 + (((uint32_t)k ^ (uint32_t)(r & 7)) << 4), acc[c][4 * k], acc[c][4 * k + 1],
                acc[c][4 * k + 2], acc[c][4 * k + 3]);
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (threadIdx.x == 0) {
          tma_reduce_add_2d(&tma_dq, buf, h * kD + c * 32, b * p.seq + i * kT);
          tma_store_commit();
        }
      }
    }
    if (threadIdx.x == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;
std::once_flag g_once;
std::string g_err;

void fused_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_enc = reinterpret_cast<EncodeTiledFn>(fn);
  cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes);
}

bool tmap2d(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld,
            uint32_t box_cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

const char* attention_bwd_fused_last_error() { return g_err.c_str(); }

size_t attention_bwd_fused_ws_floats(int batch, int seq, int n_heads) {
  return (size_t)2 * batch * n_heads * seq + (size_t)batch * seq * n_heads * kD;
}

// Same contract as attention_bwd_tcgen05; `ws` holds attention_bwd_fused_ws_floats() floats
// (D, lse2, then the [tokens, n_heads*128] fp32 dQ accumulator).
int attention_bwd_fused_tcgen05(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, float* ws,
                                int batch, int seq, int n_heads, int n_kv, int head_dim, int qkv_stride, int out_stride,
                                int dout_stride, int dqkv_stride, float scale, cudaStream_t s) {
  std::call_once(g_once, fused_init);
  if (g_enc == nullptr) { g_err = "cuTensorMapEncodeTiled unavailable"; return 1; }
  if (head_dim != kD || seq % kT != 0 || n_heads % n_kv != 0 || (qkv_stride % 8) || (out_stride % 8) ||
      (dout_stride % 8) || (dqkv_stride % 8)) {
    g_err = "attention_bwd_fused_tcgen05 needs head_dim 128, seq % 128 == 0, n_heads % n_kv == 0, 16-byte aligned rows";
    return 2;
  }
  const uint64_t cols = (uint64_t)(n_heads + 2 * n_kv) * kD, rows = (uint64_t)batch * seq;
  const uint64_t ocols = (uint64_t)n_heads * kD;
  const size_t nvec = (size_t)batch * n_heads * seq;
  float* delta = ws;
  float* lse2 = ws + nvec;
  float* dq_acc = ws + 2 * nvec;
  CUtensorMap t_qkv, t_do, t_dq;
  if (!tmap2d(&t_qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, cols, rows, qkv_stride, 64, kT) ||
      !tmap2d(&t_do, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dout, ocols, rows, dout_stride, 64, kT) ||
      !tmap2d(&t_dq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_acc, ocols, rows, ocols, 32, kT)) {
    g_err = "cuTensorMapEncodeTiled failed";
    return 3;
  }
  {
    const int64_t threads = (int64_t)rows * n_heads * 8;
    const int blocks = (int)((threads + 255) / 256);
    attn_bwd_fused_prep_kernel<<<blocks, 256, 0, s>>>(out, dout, lse, delta, lse2, dq_acc, (int)rows, seq, n_heads,
                                                      out_stride, dout_stride);
    count_launch();
  }
  FusedParams p;
  p.batch = batch; p.seq = seq; p.n_heads = n_heads; p.n_kv = n_kv;
  p.dqkv_stride = dqkv_stride;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2; p.delta = delta; p.dqkv = dqkv;
  attn_bwd_fused_kernel<<<(seq / kT) * batch * n_kv, kFusedThreads, kFusedSmemBytes, s>>>(t_qkv, t_do, t_dq, p);
  count_launch();
  {
    const int64_t n8 = (int64_t)rows * ocols / 8;
    attn_bwd_dq_convert_kernel<<<(int)((n8 + 255) / 256), 256, 0, s>>>(dq_acc, dqkv, n8, (int)(ocols / 8), dqkv_stride);
    count_launch();
  }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return 4; }
  return 0;
}

}  // namespace hb
