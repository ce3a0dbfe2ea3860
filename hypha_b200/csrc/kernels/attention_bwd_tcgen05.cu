// Causal GQA flash-attention BACKWARD on the 5th-gen tensor cores (sm_100a), head_dim 128.
//
//   P  = exp(scale * Q K^T - LSE)            (recomputed tile by tile, never stored)
//   dV = P^T dO          dP = dO V^T         dS = P o (dP - D) * scale,  D = rowsum(dO o O)
//   dK = dS^T Q          dQ = dS K
//
// Three launches, no atomics, every GEMM on tcgen05 with fp32 accumulators in TMEM:
//   attn_bwd_prep_kernel   D[b,h,t] = sum_d dO*O  and LSE in log2 units (one warp per (token, head))
//   attn_bwd_dkv_kernel    one CTA = one 128-key tile of one (batch, kv head).  It keeps K_j, V_j in
//                          smem and dK_j, dV_j in TMEM (2 x 128 columns) while it streams the 64-row
//                          Q_i / dO_i tiles of ALL q heads of the GQA group (i >= causal start):
//                            S^T = K_j Q_i^T, dP^T = V_j dO_i^T   (128x64x128, TMEM, double-buffered)
//                            softmax warpgroup: thread = key row; P^T, dS^T -> bf16 -> swizzled smem
//                            dV_j += P^T dO_i,  dK_j += dS^T Q_i  (128x128x64; Q_i/dO_i are the
//                            MN-major B operands, so no transposed copy of anything is ever made)
//   attn_bwd_dq_kernel     one CTA = one 128-row Q tile of one (batch, head); streams 64-key K_j/V_j:
//                            S = Q K_j^T, dP = dO V_j^T -> dS (thread = q row) -> dQ += dS K_j (TMEM)
// Roles per CTA: warp 0 = TMA producer, warp 1 = MMA issuer, warps 4-19 = FOUR softmax warpgroups.
// The backward needs no row reductions (LSE and D are inputs), so the 64 columns of every S/dP tile
// are split 16 per warpgroup: all four read the same TMEM lanes (lane = warp%4 quarter) at different
// column offsets, which cuts the exp/convert latency that sits between the two MMA groups of a tile
// by 4x (ncu on the 1-warpgroup version: tensor pipe 25-30 %, issue slots 24 % busy -> latency-bound).
// The S/dP MMAs of tile i+1 are issued before the dV/dK (dQ) MMAs of tile i, so the tensor pipe works
// under the exponentials.
// Results are written straight into the fused d(qkv) activation gradient that the qkv projection's
// dgrad/wgrad GEMMs consume.
//
// Replaces the stock attention backward the reference reaches through `accelerator.backward(loss)`
// (training.py:110-112).
#include <mutex>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace hb {
namespace {
using namespace sm100;

constexpr int kD = 128;
constexpr int kWG = 4;                       // softmax warpgroups; each owns 64 / kWG columns of a tile
constexpr int kColsPerWG = 64 / kWG;         // 16
constexpr int kThreads = 128 + 128 * kWG;    // 640
constexpr int kTmemCols = 512;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_16(uint32_t dst, const uint32_t* v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct BwdParams {
  int batch, seq, n_heads, n_kv;
  int dqkv_stride;        // elements per token row of dqkv
  float scale, scale_log2;
  const float* lse2;      // [batch, n_heads, seq]  log-sum-exp in log2 units
  const float* delta;     // [batch, n_heads, seq]  rowsum(dO o O)
  bf16* dqkv;             // [tokens, (Hq + 2 Hkv) * 128]
};

// ---------------------------------------------------------------------------------------------
// D = rowsum(dO o O), lse2 = lse * log2(e)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout,
                                                            const float* __restrict__ lse, float* __restrict__ delta,
                                                            float* __restrict__ lse2, int tokens, int seq, int n_heads,
                                                            int out_stride, int dout_stride) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // (token, head)
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)tokens * n_heads) return;
  const int t = (int)(w / n_heads), h = (int)(w % n_heads);
  const bf16* o = out + (size_t)t * out_stride + h * kD + lane * 4;
  const bf16* g = dout + (size_t)t * dout_stride + h * kD + lane * 4;
  const uint2 ov = *reinterpret_cast<const uint2*>(o), gv = *reinterpret_cast<const uint2*>(g);
  const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
  const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gv);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float2 a = __bfloat1622float2(o2[i]), b = __bfloat1622float2(g2[i]);
    acc += a.x * b.x + a.y * b.y;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    const int b = t / seq, pos = t % seq;
    const size_t idx = ((size_t)b * n_heads + h) * seq + pos;
    delta[idx] = acc;
    lse2[idx] = lse[idx] * 1.4426950408889634f;
  }
}

// ---------------------------------------------------------------------------------------------
// dK, dV
// ---------------------------------------------------------------------------------------------
constexpr int kKV = 128;             // keys per CTA
constexpr int kBQ = 64;              // q rows per iteration
constexpr int kKVTileBytes = kKV * kD * 2;     // 32 KB
constexpr int kQTileBytes = kBQ * kD * 2;      // 16 KB
constexpr int kPTBytes = kKV * kBQ * 2;        // 16 KB  ([128 keys][64 q] bf16, one swizzle atom wide)
// Q_i/dO_i ring depth.  A stage is only released when the dV/dK MMAs that read it have completed, so
// with 2 stages every iteration paid a full TMA round trip (measured 2.0 us/iteration vs 0.6 us of
// tensor work); 3 stages is what fits next to K_j, V_j, P^T and dS^T in 227 KB.
constexpr int kQStages = 3;

struct DkvSmem {
  static constexpr int k_off = 0;
  static constexpr int v_off = k_off + kKVTileBytes;
  static constexpr int q_off = v_off + kKVTileBytes;              // kQStages stages
  static constexpr int do_off = q_off + kQStages * kQTileBytes;   // kQStages stages
  static constexpr int pt_off = do_off + kQStages * kQTileBytes;
  static constexpr int ds_off = pt_off + kPTBytes;
  static constexpr int vec_off = ds_off + kPTBytes;               // lse2/D staging: kQStages x 128 floats
  static constexpr int bar_off = vec_off + kQStages * 128 * 4;
  static constexpr int total = bar_off + 24 * 8;
};
constexpr int kDkvSmemBytes = DkvSmem::total + 1024;

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tma_qkv128,    // qkv, box {64, 128}
                    const __grid_constant__ CUtensorMap tma_qkv64,     // qkv, box {64, 64}
                    const __grid_constant__ CUtensorMap tma_do64,      // dO,  box {64, 64}
                    const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t k_smem = smem_base + DkvSmem::k_off, v_smem = smem_base + DkvSmem::v_off;
  const uint32_t q_smem = smem_base + DkvSmem::q_off, do_smem = smem_base + DkvSmem::do_off;
  const uint32_t pt_smem = smem_base + DkvSmem::pt_off, ds_smem = smem_base + DkvSmem::ds_off;
  const float* vec_gen = reinterpret_cast<const float*>(smem_gen + DkvSmem::vec_off);
  const uint32_t vec_smem = smem_base + DkvSmem::vec_off;
  const uint32_t bars = smem_base + DkvSmem::bar_off;
  const uint32_t kv_full = bars;
  auto qdo_full = [&](int s) { return bars + 8u * (1 + s); };                 // 1..4
  auto qdo_empty = [&](int s) { return bars + 8u * (5 + s); };                // 5..8
  auto sdp_full = [&](int s) { return bars + 8u * (9 + s); };                 // 9,10
  auto sdp_free = [&](int s) { return bars + 8u * (11 + s); };                // 11,12
  const uint32_t pds_full = bars + 8u * 13, pds_free = bars + 8u * 14, acc_full = bars + 8u * 15;
  const uint32_t tmem_ptr_smem = bars + 8u * 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + DkvSmem::bar_off + 8 * 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv_tiles = p.seq / kKV;
  const int per_tile = p.batch * p.n_kv;
  const int j = (int)(blockIdx.x / per_tile);                      // early key tiles (most work) first
  const int bh = (int)(blockIdx.x % per_tile);
  const int b = bh / p.n_kv, hk = bh % p.n_kv;
  const int G = p.n_heads / p.n_kv;
  const int i0 = j * (kKV / kBQ);                                  // first 64-row q tile that sees this key tile
  const int n_q = p.seq / kBQ - i0;
  const int n_it = G * n_q;
  const int krow0 = b * p.seq + j * kKV;
  const int k_col = (p.n_heads + hk) * kD, v_col = (p.n_heads + p.n_kv + hk) * kD;
  (void)kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_qkv128);
    tma_prefetch_desc(&tma_qkv64);
    tma_prefetch_desc(&tma_do64);
    mbar_init(kv_full, 1);
    for (int s = 0; s < kQStages; ++s) { mbar_init(qdo_full(s), 1); mbar_init(qdo_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(sdp_full(s), 1); mbar_init(sdp_free(s), 4 * kWG); }
    mbar_init(pds_full, 4 * kWG); mbar_init(pds_free, 1); mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  auto tmem_s = [&](int buf) { return tmem_base + 64u * buf; };
  auto tmem_dp = [&](int buf) { return tmem_base + 128u + 64u * buf; };
  const uint32_t tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 384;

  if (warp < 4) {
    if (warp == 0 && lane == 0) {
      // ====================== TMA producer ======================
      mbar_arrive_expect_tx(kv_full, 2 * kKVTileBytes);
      tma_load_2d(k_smem, &tma_qkv128, kv_full, k_col, krow0);
      tma_load_2d(k_smem + kKV * 128, &tma_qkv128, kv_full, k_col + 64, krow0);
      tma_load_2d(v_smem, &tma_qkv128, kv_full, v_col, krow0);
      tma_load_2d(v_smem + kKV * 128, &tma_qkv128, kv_full, v_col + 64, krow0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % kQStages;
        const int h = hk * G + it / n_q, i = i0 + it % n_q;
        const int qrow = b * p.seq + i * kBQ;
        mbar_wait(qdo_empty(st), ((uint32_t)(it / kQStages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(qdo_full(st), 2 * kQTileBytes + 2 * kBQ * 4);
        {  // lse2 and D of the 64 q rows ride on the same barrier
          const size_t vbase = ((size_t)b * p.n_heads + h) * p.seq + (size_t)i * kBQ;
          bulk_load_1d(vec_smem + st * 512, p.lse2 + vbase, kBQ * 4, qdo_full(st));
          bulk_load_1d(vec_smem + st * 512 + 256, p.delta + vbase, kBQ * 4, qdo_full(st));
        }
        tma_load_2d(q_smem + st * kQTileBytes, &tma_qkv64, qdo_full(st), h * kD, qrow);
        tma_load_2d(q_smem + st * kQTileBytes + kBQ * 128, &tma_qkv64, qdo_full(st), h * kD + 64, qrow);
        tma_load_2d(do_smem + st * kQTileBytes, &tma_do64, qdo_full(st), h * kD, qrow);
        tma_load_2d(do_smem + st * kQTileBytes + kBQ * 128, &tma_do64, qdo_full(st), h * kD + 64, qrow);
      }
    } else if (warp == 1 && lane == 0) {
      // ====================== MMA issuer ========================
      const uint32_t idesc_s = umma_idesc_bf16(kKV, kBQ, 0, 0);    // S^T = K Q^T, dP^T = V dO^T : K-major x K-major
      const uint32_t idesc_acc = umma_idesc_bf16(kKV, kD, 0, 1);   // dV += P^T dO, dK += dS^T Q : B is MN-major
      auto issue_sdp = [&](int it) {
        const int st = it % kQStages, tb = it & 1;                  // smem stage, TMEM S/dP buffer
        mbar_wait(qdo_full(st), (uint32_t)(it / kQStages) & 1u);
        if (it >= 2) mbar_wait(sdp_free(tb), (uint32_t)((it >> 1) - 1) & 1u);
        tc_fence_after();
        const uint32_t qb = q_smem + st * kQTileBytes, dob = do_smem + st * kQTileBytes;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(k_smem + (kk >> 2) * (kKV * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(qb + (kk >> 2) * (kBQ * 128) + (kk & 3) * 32, 16, 1024);
          umma_bf16_ss(tmem_s(tb), ad, bd, idesc_s, kk != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(v_smem + (kk >> 2) * (kKV * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(dob + (kk >> 2) * (kBQ * 128) + (kk & 3) * 32, 16, 1024);
          umma_bf16_ss(tmem_dp(tb), ad, bd, idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(sdp_full(tb));
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_sdp(it + 1);                       // overlaps the softmax of tile `it`
        const int st = it % kQStages;
        mbar_wait(pds_full, (uint32_t)it & 1u);
        tc_fence_after();
        const uint32_t qb = q_smem + st * kQTileBytes, dob = do_smem + st * kQTileBytes;
#pragma unroll
        for (int kk = 0; kk < kBQ / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(pt_smem + kk * 32, 16, 1024);               // P^T: K-major, one atom
          const uint64_t bd = umma_smem_desc(dob + kk * (16 * 128), kBQ * 128, 1024);    // dO: MN-major, 2 chunks
          umma_bf16_ss(tmem_dv, ad, bd, idesc_acc, (it != 0 || kk != 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < kBQ / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(ds_smem + kk * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(qb + kk * (16 * 128), kBQ * 128, 1024);
          umma_bf16_ss(tmem_dk, ad, bd, idesc_acc, (it != 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit(pds_free);
        umma_commit(qdo_empty(st));
        if (it == n_it - 1) umma_commit(acc_full);
      }
    }
  } else {
    // ====================== softmax warpgroups: thread = key row, 16 q columns each ==========
    const int q4 = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int r_local = q4 * 32 + lane;                             // key row inside the tile == TMEM lane
    const int k_pos = j * kKV + r_local;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const int col0 = wg * kColsPerWG;
    for (int it = 0; it < n_it; ++it) {
      const int ss = it % kQStages, st = it & 1;                    // smem stage, TMEM S/dP buffer
      const int i = i0 + it % n_q;
      const float* lse_s = vec_gen + ss * 128;
      const float* d_s = lse_s + 64;
      mbar_wait(qdo_full(ss), (uint32_t)(it / kQStages) & 1u);      // lse2 / D staged by the producer
      mbar_wait(sdp_full(st), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const bool diag = i * kBQ < j * kKV + kKV;                    // some (key, q) pairs are masked
      const int q_pos0 = i * kBQ;
      uint32_t sv[16], dp[16];
      tmem_ld_32x16(tmem_s(st) + lane_addr + col0, sv);
      tmem_ld_32x16(tmem_dp(st) + lane_addr + col0, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free(st));
      uint32_t pk[8], dk[8];
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        float pv[2], dsv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = col0 + c + e;
          float pe = ex2(fmaf(__uint_as_float(sv[c + e]), p.scale_log2, -lse_s[col]));
          if (diag && k_pos > q_pos0 + col) pe = 0.f;
          pv[e] = pe;
          dsv[e] = pe * (__uint_as_float(dp[c + e]) - d_s[col]) * p.scale;
        }
        pk[c >> 1] = pack_bf16x2(pv[0], pv[1]);
        dk[c >> 1] = pack_bf16x2(dsv[0], dsv[1]);
      }
      if (it > 0) mbar_wait(pds_free, (uint32_t)(it - 1) & 1u);     // dV/dK MMAs of tile it-1 done
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const uint32_t sw = (uint32_t)(((wg * 2 + ch) ^ (r_local & 7)) << 4);
        st_shared_16(pt_smem + r_local * 128 + sw, pk + 4 * ch);
        st_shared_16(ds_smem + r_local * 128 + sw, dk + 4 * ch);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    // ---- epilogue: dV, dK -> bf16 -> d(qkv); warpgroup g owns columns [32g, 32g+32) of both ----
    mbar_wait(acc_full, 0);
    tc_fence_after();
    bf16* row = p.dqkv + (size_t)(krow0 + r_local) * p.dqkv_stride;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      for (int c = wg; c < kD / 32; c += kWG) {
        bf16* dst = row + (which == 0 ? v_col : k_col) + c * 32;
        uint32_t t[32];
        tmem_ld_32x32((which == 0 ? tmem_dv : tmem_dk) + lane_addr + c * 32, t);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------
constexpr int kBM = 128;             // q rows per CTA
constexpr int kBN = 64;              // keys per iteration
constexpr int kQBytes = kBM * kD * 2;          // 32 KB
constexpr int kKBytes = kBN * kD * 2;          // 16 KB
constexpr int kDSBytes = kBM * kBN * 2;        // 16 KB
constexpr int kKVStages = 4;                   // K_j/V_j ring (a stage is held until its dQ MMAs completed)

struct DqSmem {
  static constexpr int q_off = 0;
  static constexpr int do_off = q_off + kQBytes;
  static constexpr int k_off = do_off + kQBytes;                  // kKVStages stages
  static constexpr int v_off = k_off + kKVStages * kKBytes;       // kKVStages stages
  static constexpr int ds_off = v_off + kKVStages * kKBytes;
  static constexpr int bar_off = ds_off + kDSBytes;
  static constexpr int total = bar_off + 24 * 8;
};
constexpr int kDqSmemBytes = DqSmem::total + 1024;

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tma_qkv128,     // qkv, box {64, 128}
                   const __grid_constant__ CUtensorMap tma_qkv64,      // qkv, box {64, 64}
                   const __grid_constant__ CUtensorMap tma_do128,      // dO,  box {64, 128}
                   const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_smem = smem_base + DqSmem::q_off, do_smem = smem_base + DqSmem::do_off;
  const uint32_t k_smem = smem_base + DqSmem::k_off, v_smem = smem_base + DqSmem::v_off;
  const uint32_t ds_smem = smem_base + DqSmem::ds_off;
  const uint32_t bars = smem_base + DqSmem::bar_off;
  const uint32_t qdo_full = bars;
  auto kv_full = [&](int s) { return bars + 8u * (1 + s); };                  // 1..4
  auto kv_empty = [&](int s) { return bars + 8u * (5 + s); };                 // 5..8
  auto sdp_full = [&](int s) { return bars + 8u * (9 + s); };                 // 9,10
  auto sdp_free = [&](int s) { return bars + 8u * (11 + s); };                // 11,12
  const uint32_t ds_full = bars + 8u * 13, ds_free = bars + 8u * 14, acc_full = bars + 8u * 15;
  const uint32_t tmem_ptr_smem = bars + 8u * 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + DqSmem::bar_off + 8 * 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tiles = p.seq / kBM;
  const int per_tile = p.batch * p.n_heads;
  const int qt = q_tiles - 1 - (int)(blockIdx.x / per_tile);       // late q tiles (most keys) first
  const int bh = (int)(blockIdx.x % per_tile);
  const int b = bh / p.n_heads, h = bh % p.n_heads;
  const int hk = h / (p.n_heads / p.n_kv);
  const int n_tiles = 2 * qt + 2;
  const int row0 = b * p.seq + qt * kBM;
  const int q_col = h * kD, k_col = (p.n_heads + hk) * kD, v_col = (p.n_heads + p.n_kv + hk) * kD;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_qkv128);
    tma_prefetch_desc(&tma_qkv64);
    tma_prefetch_desc(&tma_do128);
    mbar_init(qdo_full, 1);
    for (int s = 0; s < kKVStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(sdp_full(s), 1); mbar_init(sdp_free(s), 4 * kWG); }
    mbar_init(ds_full, 4 * kWG); mbar_init(ds_free, 1); mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  auto tmem_s = [&](int buf) { return tmem_base + 64u * buf; };
  auto tmem_dp = [&](int buf) { return tmem_base + 128u + 64u * buf; };
  const uint32_t tmem_dq = tmem_base + 256;

  if (warp < 4) {
    if (warp == 0 && lane == 0) {
      // ====================== TMA producer ======================
      mbar_arrive_expect_tx(qdo_full, 2 * kQBytes);
      tma_load_2d(q_smem, &tma_qkv128, qdo_full, q_col, row0);
      tma_load_2d(q_smem + kBM * 128, &tma_qkv128, qdo_full, q_col + 64, row0);
      tma_load_2d(do_smem, &tma_do128, qdo_full, q_col, row0);
      tma_load_2d(do_smem + kBM * 128, &tma_do128, qdo_full, q_col + 64, row0);
      for (int jj = 0; jj < n_tiles; ++jj) {
        const int st = jj % kKVStages;
        const int krow = b * p.seq + jj * kBN;
        mbar_wait(kv_empty(st), ((uint32_t)(jj / kKVStages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(kv_full(st), 2 * kKBytes);
        tma_load_2d(k_smem + st * kKBytes, &tma_qkv64, kv_full(st), k_col, krow);
        tma_load_2d(k_smem + st * kKBytes + kBN * 128, &tma_qkv64, kv_full(st), k_col + 64, krow);
        tma_load_2d(v_smem + st * kKBytes, &tma_qkv64, kv_full(st), v_col, krow);
        tma_load_2d(v_smem + st * kKBytes + kBN * 128, &tma_qkv64, kv_full(st), v_col + 64, krow);
      }
    } else if (warp == 1 && lane == 0) {
      // ====================== MMA issuer ========================
      const uint32_t idesc_s = umma_idesc_bf16(kBM, kBN, 0, 0);    // S = Q K^T, dP = dO V^T
      const uint32_t idesc_dq = umma_idesc_bf16(kBM, kD, 0, 1);    // dQ += dS K : K_j is the MN-major B operand
      auto issue_sdp = [&](int jj) {
        const int st = jj % kKVStages, tb = jj & 1;                 // smem stage, TMEM S/dP buffer
        mbar_wait(kv_full(st), (uint32_t)(jj / kKVStages) & 1u);
        if (jj >= 2) mbar_wait(sdp_free(tb), (uint32_t)((jj >> 1) - 1) & 1u);
        tc_fence_after();
        const uint32_t kb = k_smem + st * kKBytes, vb = v_smem + st * kKBytes;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(q_smem + (kk >> 2) * (kBM * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(kb + (kk >> 2) * (kBN * 128) + (kk & 3) * 32, 16, 1024);
          umma_bf16_ss(tmem_s(tb), ad, bd, idesc_s, kk != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(do_smem + (kk >> 2) * (kBM * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(vb + (kk >> 2) * (kBN * 128) + (kk & 3) * 32, 16, 1024);
          umma_bf16_ss(tmem_dp(tb), ad, bd, idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(sdp_full(tb));
      };
      mbar_wait(qdo_full, 0);
      issue_sdp(0);
      for (int jj = 0; jj < n_tiles; ++jj) {
        if (jj + 1 < n_tiles) issue_sdp(jj + 1);
        const int st = jj % kKVStages;
        mbar_wait(ds_full, (uint32_t)jj & 1u);
        tc_fence_after();
        const uint32_t kb = k_smem + st * kKBytes;
#pragma unroll
        for (int kk = 0; kk < kBN / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(ds_smem + kk * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(kb + kk * (16 * 128), kBN * 128, 1024);
          umma_bf16_ss(tmem_dq, ad, bd, idesc_dq, (jj != 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit(ds_free);
        umma_commit(kv_empty(st));
        if (jj == n_tiles - 1) umma_commit(acc_full);
      }
    }
  } else {
    // ====================== softmax warpgroups: thread = q row, 16 key columns each ==========
    const int q4 = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int r_local = q4 * 32 + lane;
    const int q_pos = qt * kBM + r_local;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const int col0 = wg * kColsPerWG;
    const size_t vidx = ((size_t)b * p.n_heads + h) * p.seq + q_pos;
    const float lse_r = p.lse2[vidx], d_r = p.delta[vidx];
    for (int jj = 0; jj < n_tiles; ++jj) {
      const int st = jj & 1;
      mbar_wait(sdp_full(st), (uint32_t)(jj >> 1) & 1u);
      tc_fence_after();
      const int key0 = jj * kBN;
      const bool diag = key0 + kBN - 1 > qt * kBM;
      uint32_t sv[16], dp[16];
      tmem_ld_32x16(tmem_s(st) + lane_addr + col0, sv);
      tmem_ld_32x16(tmem_dp(st) + lane_addr + col0, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free(st));
      uint32_t dk[8];
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        float dsv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = col0 + c + e;
          float pe = ex2(fmaf(__uint_as_float(sv[c + e]), p.scale_log2, -lse_r));
          if (diag && key0 + col > q_pos) pe = 0.f;
          dsv[e] = pe * (__uint_as_float(dp[c + e]) - d_r) * p.scale;
        }
        dk[c >> 1] = pack_bf16x2(dsv[0], dsv[1]);
      }
      if (jj > 0) mbar_wait(ds_free, (uint32_t)(jj - 1) & 1u);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const uint32_t sw = (uint32_t)(((wg * 2 + ch) ^ (r_local & 7)) << 4);
        st_shared_16(ds_smem + r_local * 128 + sw, dk + 4 * ch);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    bf16* dst = p.dqkv + (size_t)(row0 + r_local) * p.dqkv_stride + q_col;
    for (int c = wg; c < kD / 32; c += kWG) {
      uint32_t t[32];
      tmem_ld_32x32(tmem_dq + lane_addr + c * 32, t);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        float f[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) f[x] = __uint_as_float(t[e + x]);
        st_bf16x8(dst + c * 32 + e, pack8(f));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;
std::once_flag g_once;
std::string g_err;

void bwd_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_enc = reinterpret_cast<EncodeTiledFn>(fn);
  cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmemBytes);
  cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmemBytes);
}

bool tmap(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

const char* attention_bwd_last_error() { return g_err.c_str(); }

// qkv  : [batch*seq, (n_heads + 2 n_kv) * 128] bf16, RoPE'd (what the forward consumed)
// out  : [batch*seq, n_heads*128] bf16 forward output;   dout: its gradient
// lse  : [batch, n_heads, seq] fp32 natural-log LSE from the forward
// dqkv : [batch*seq, (n_heads + 2 n_kv) * 128] bf16, fully overwritten
// ws   : fp32 workspace of 2 * batch * n_heads * seq floats (D and lse2)
int attention_bwd_tcgen05(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, float* ws,
                          int batch, int seq, int n_heads, int n_kv, int head_dim, int qkv_stride, int out_stride,
                          int dout_stride, int dqkv_stride, float scale, cudaStream_t s) {
  std::call_once(g_once, bwd_init);
  if (g_enc == nullptr) { g_err = "cuTensorMapEncodeTiled unavailable"; return 1; }
  if (head_dim != kD || seq % 128 != 0 || n_heads % n_kv != 0 || (qkv_stride % 8) || (out_stride % 8) ||
      (dout_stride % 8) || (dqkv_stride % 8)) {
    g_err = "attention_bwd_tcgen05 needs head_dim 128, seq % 128 == 0, n_heads % n_kv == 0, 16-byte aligned rows";
    return 2;
  }
  const uint64_t cols = (uint64_t)(n_heads + 2 * n_kv) * kD, rows = (uint64_t)batch * seq;
  const uint64_t ocols = (uint64_t)n_heads * kD;
  CUtensorMap t_qkv128, t_qkv64, t_do128, t_do64;
  if (!tmap(&t_qkv128, qkv, cols, rows, qkv_stride, 128) || !tmap(&t_qkv64, qkv, cols, rows, qkv_stride, 64) ||
      !tmap(&t_do128, dout, ocols, rows, dout_stride, 128) || !tmap(&t_do64, dout, ocols, rows, dout_stride, 64)) {
    g_err = "cuTensorMapEncodeTiled failed";
    return 3;
  }
  const size_t nvec = (size_t)batch * n_heads * seq;
  float* delta = ws;
  float* lse2 = ws + nvec;
  {
    const int64_t warps = (int64_t)rows * n_heads;
    const int blocks = (int)((warps * 32 + 255) / 256);
    attn_bwd_prep_kernel<<<blocks, 256, 0, s>>>(out, dout, lse, delta, lse2, (int)rows, seq, n_heads, out_stride,
                                                dout_stride);
    count_launch();
  }
  BwdParams p;
  p.batch = batch; p.seq = seq; p.n_heads = n_heads; p.n_kv = n_kv;
  p.dqkv_stride = dqkv_stride;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2; p.delta = delta; p.dqkv = dqkv;
  attn_bwd_dkv_kernel<<<(seq / kKV) * batch * n_kv, kThreads, kDkvSmemBytes, s>>>(t_qkv128, t_qkv64, t_do64, p);
  count_launch();
  attn_bwd_dq_kernel<<<(seq / kBM) * batch * n_heads, kThreads, kDqSmemBytes, s>>>(t_qkv128, t_qkv64, t_do128, p);
  count_launch();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return 4; }
  return 0;
}

}  // namespace hb
