// Causal GQA flash-attention FORWARD on the 5th-gen tensor cores (sm_100a), head_dim 128.
//
//   O = softmax(scale * Q K^T + causal mask) V      (online softmax, never materialises S)
//
// One CTA = one 128-row Q tile of one (batch, head); KV is streamed in 64-key tiles.
//   warp 0      TMA producer: Q once, then K_j / V_j into a 2-stage ring (cp.async.bulk.tensor,
//               128B swizzle, straight out of the fused [tokens, (Hq+2Hkv)*D] qkv activation)
//   warp 1      tcgen05.mma issuer:  S = Q K_j^T  (128x64x128, accumulator in TMEM cols 0..63)
//                                    O_j = P_j V_j (128x128x64, TMEM cols 64..191), V is the
//                                    MN-major B operand so no transpose is ever made
//   warps 4-7   softmax warpgroup: thread t owns row t; tcgen05.ld S -> scale/mask/exp2 ->
//               running (max, sum) -> P (bf16) into TMEM (cols 192..223) as the A operand of the PV
//               MMA; O accumulates in TMEM across tiles and is rescaled in place (tcgen05.ld ->
//               scale -> tcgen05.st) only when the row max grew by more than 2^8 (lazy rescale)
//               -> final tcgen05.ld, normalise, store O and log-sum-exp
// S_{j+1} is issued as soon as S_j has been read out of TMEM, so the QK^T of the next tile
// overlaps the exponentials of the current one; two CTAs are co-resident per SM (96 KB smem,
// 256 TMEM columns each) so one CTA's softmax hides under the other's MMAs.  `setmaxnreg`
// moves registers from the producer/MMA warpgroup to the softmax warpgroup (128-float O row).
//
// Replaces the stock attention the reference reaches through HF `model(**batch)`
// (training.py:108-112); the backward pass still uses the library kernel.
#include <mutex>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace hb {
namespace {
using namespace sm100;

constexpr int kD = 128;          // head dim
constexpr int kBM = 128;         // q rows per CTA
constexpr int kBN = 64;          // keys per tile
constexpr int kKVStages = 2;         // K ring depth
constexpr int kVStages = 2;          // P lives in TMEM, so its 16 KB of smem buy a second V stage
constexpr int kQBytes = kBM * kD * 2;          // 32 KB (two 64-column swizzle atoms)
constexpr int kKBytes = kBN * kD * 2;          // 16 KB
constexpr int kVBytes = kBN * kD * 2;          // 16 KB
constexpr int kAttnTmemCols = 256;             // S: 0..63, O_j: 64..191
constexpr int kAttnThreads = 256;

struct AttnSmem {
  static constexpr int q_off = 0;
  static constexpr int k_off = q_off + kQBytes;
  static constexpr int v_off = k_off + kKVStages * kKBytes;
  static constexpr int bar_off = v_off + kVStages * kVBytes;
  // q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full, s_free, p_full, o_full, (spare), tmem ptr
  static constexpr int total = bar_off + 16 * 8;
};
constexpr int kAttnSmemBytes = AttnSmem::total + 1024;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU op; exp2f() adds range fix-ups
  return y;
}

struct AttnParams {
  int batch, seq, n_heads, n_kv, q_tiles;
  int out_stride;        // elements per token row of `out`
  float scale_log2;      // softmax scale * log2(e)
  bf16* out;
  float* lse;            // [batch, n_heads, seq]
};

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_q,      // box {64, 128}
                        const __grid_constant__ CUtensorMap tma_kv,     // box {64, 64}
                        const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_smem = smem_base + AttnSmem::q_off;
  const uint32_t k_smem = smem_base + AttnSmem::k_off;
  const uint32_t v_smem = smem_base + AttnSmem::v_off;
  const uint32_t bars = smem_base + AttnSmem::bar_off;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (3 + s); };
  auto v_full = [&](int s) { return bars + 8u * (5 + s); };
  auto v_empty = [&](int s) { return bars + 8u * (7 + s); };
  const uint32_t s_full = bars + 8u * 9, s_free = bars + 8u * 10, p_full = bars + 8u * 11;
  const uint32_t o_full = bars + 8u * 12;
  const uint32_t tmem_ptr_smem = bars + 8u * 14;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + AttnSmem::bar_off + 8 * 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy (late) q tiles first: the causal triangle is load-balanced by launch order
  const int per_tile = p.batch * p.n_heads;
  const int qt = p.q_tiles - 1 - (int)(blockIdx.x / per_tile);
  const int bh = (int)(blockIdx.x % per_tile);
  const int b = bh / p.n_heads, h = bh % p.n_heads;
  const int hk = h / (p.n_heads / p.n_kv);
  const int n_tiles = 2 * qt + 2;                                  // keys 0 .. qt*128+127
  const int row0 = b * p.seq + qt * kBM;                           // first token row of the q tile
  const int q_col = h * kD, k_col = (p.n_heads + hk) * kD, v_col = (p.n_heads + p.n_kv + hk) * kD;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, 4); mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, kAttnTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 64, tmem_p = tmem_base + 192;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ====================== TMA producer ======================
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_2d(q_smem, &tma_q, q_full, q_col, row0);
      tma_load_2d(q_smem + kBM * 128, &tma_q, q_full, q_col + 64, row0);
      auto load_k = [&](int j) {
        const int st = j & 1;
        const int krow = b * p.seq + j * kBN;
        mbar_wait(k_empty(st), ((uint32_t)(j >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(k_full(st), kKBytes);
        tma_load_2d(k_smem + st * kKBytes, &tma_kv, k_full(st), k_col, krow);
        tma_load_2d(k_smem + st * kKBytes + kBN * 128, &tma_kv, k_full(st), k_col + 64, krow);
      };
      load_k(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) load_k(j + 1);                          // K runs one tile ahead of V
        const int krow = b * p.seq + j * kBN;
        const int vs = j & 1;
        mbar_wait(v_empty(vs), ((uint32_t)(j >> 1) & 1u) ^ 1u);       // PV_{j-2} has consumed this V stage
        mbar_arrive_expect_tx(v_full(vs), kVBytes);
        tma_load_2d(v_smem + vs * kVBytes, &tma_kv, v_full(vs), v_col, krow);
        tma_load_2d(v_smem + vs * kVBytes + kBN * 128, &tma_kv, v_full(vs), v_col + 64, krow);
      }
    } else if (warp == 1 && lane == 0) {
      // ====================== MMA issuer ========================
      const uint32_t idesc_s = umma_idesc_bf16(kBM, kBN, 0, 0);    // S = Q K^T : both K-major
      const uint32_t idesc_o = umma_idesc_bf16(kBM, kD, 0, 1);     // O = P V   : V is MN-major
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(k_full(st), (uint32_t)(j >> 1) & 1u);
        if (j > 0) mbar_wait(s_free, (uint32_t)(j - 1) & 1u);      // S_{j-1} has been read out
        tc_fence_after();
        const uint32_t kb = k_smem + st * kKBytes;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(q_smem + (kk >> 2) * (kBM * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(kb + (kk >> 2) * (kBN * 128) + (kk & 3) * 32, 16, 1024);
          umma_bf16_ss(tmem_s, ad, bd, idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(k_empty(st));
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);                        // overlaps softmax_j
        const int st = j & 1;
        mbar_wait(p_full, (uint32_t)j & 1u);          // also orders any rescale of O before this MMA
        mbar_wait(v_full(st), (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        const uint32_t vb = v_smem + st * kVBytes;
#pragma unroll
        for (int kk = 0; kk < kBN / 16; ++kk) {
          const uint64_t bd = umma_smem_desc(vb + kk * (16 * 128), kBN * 128, 1024);     // V: MN-major, 2 chunks
          umma_bf16_ts(tmem_o, tmem_p + kk * 8, bd, idesc_o, (j != 0 || kk != 0) ? 1u : 0u);   // P from TMEM
        }
        umma_commit(o_full);
        umma_commit(v_empty(st));
      }
    }
  } else {
    // ====================== softmax / epilogue warpgroup ==========
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int q4 = warp & 3;                                        // TMEM lane quarter of this warp
    const int r_local = q4 * 32 + lane;                             // row inside the q tile
    const int q_pos = qt * kBM + r_local;                           // position in the sequence
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    // O stays in TMEM for the whole KV loop.  The running max used for the exponentials (m_used) is
    // only advanced when the true max has grown by more than 2^8 since the last advance (lazy
    // rescale): P <= 256 is harmless in bf16 / fp32 accumulation, and the TMEM read-modify-write of
    // the 128-column O row then happens a handful of times per row instead of once per tile.
    float m_used = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(tmem_s + lane_addr, s0);
      tmem_ld_32x32(tmem_s + lane_addr + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      const int key0 = j * kBN;
      const bool diag = key0 + kBN - 1 > qt * kBM;                  // tile touches the causal boundary
      float x[64];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float v = __uint_as_float(c < 32 ? s0[c & 31] : s1[c & 31]) * p.scale_log2;
        if (diag && key0 + c > q_pos) v = -INFINITY;
        x[c] = v;
        mx = fmaxf(mx, v);
      }
      const bool need = mx > m_used + 8.f;                          // true on the first tile (m_used = -inf)
      const float m_next = need ? mx : m_used;
      const float alpha = need ? fast_exp2(m_used - m_next) : 1.f;  // 0 on the first tile
      const float m_safe = (m_next == -INFINITY) ? 0.f : m_next;
      float sum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        const float p0 = fast_exp2(x[c] - m_safe), p1 = fast_exp2(x[c + 1] - m_safe);
        sum += p0 + p1;
        __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
        pk[c >> 1] = *reinterpret_cast<uint32_t*>(&hh);
      }
      l_run = l_run * alpha + sum;
      m_used = m_next;
      // PV_{j-1} must have finished before P is overwritten or O is rescaled
      if (j > 0) {
        mbar_wait(o_full, (uint32_t)(j - 1) & 1u);
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
#pragma unroll
          for (int c = 0; c < kD / 32; ++c) {
            uint32_t t[32];
            tmem_ld_32x32(tmem_o + lane_addr + c * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st_32x32(tmem_o + lane_addr + c * 32, t);
          }
          tmem_st_wait();
        }
      }
      // P (bf16 pairs) goes to TMEM columns 192..223 of this thread's lane: the PV MMA takes it as its
      // A operand from tensor memory, so P costs neither an smem round trip nor MMA operand bandwidth
      tmem_st_32x32(tmem_p + lane_addr, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, (uint32_t)(n_tiles - 1) & 1u);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const float m_run = m_used;
    bf16* orow = p.out + (size_t)(row0 + r_local) * p.out_stride + h * kD;
#pragma unroll
    for (int c = 0; c < kD / 32; ++c) {
      uint32_t t[32];
      tmem_ld_32x32(tmem_o + lane_addr + c * 32, t);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(t[e + i]) * inv;
        st_bf16x8(orow + c * 32 + e, pack8(f));
      }
    }
    p.lse[((size_t)b * p.n_heads + h) * p.seq + q_pos] =
        (l_run > 0.f) ? (m_run + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}


// =====================================================================================================
// Forward v2: TWO 128-row Q tiles per CTA ping-pong on one MMA issuer (one CTA per SM), 128-key tiles.
//
//   TMEM (512 columns): S0 0..127 | S1 128..255 | O0 256..383 | O1 384..511; P_w (bf16 pairs) overwrites the
//   first 64 columns of S_w and is the TMEM A operand of O_w += P_w V (TS form).
//   warps 0-3 / 4-7  softmax warpgroup of Q tile 0 / 1: thread = q row, all 128 key columns of the tile in
//                    registers; row max with 3-input max, scale-and-subtract with packed fp32x2 FMAs, ex2,
//                    lazy rescale of O_w in TMEM (only when the max grew by > 2^8), final normalise + store.
//   warp 8           tcgen05.mma issuer.  Per key tile j the tensor pipe runs
//                        O0 += P0_j V_j | S0_{j+1} = Q0 K_{j+1}^T | O1 += P1_j V_j | S1_{j+1} = Q1 K_{j+1}^T
//                    so one tile's exponentials hide under the other tile's MMAs, and K_j / V_j are fetched
//                    once for 256 q rows.  S_{j+1} of a tile is issued right behind its O += P_j V, so a
//                    completed S_{j+1} also means the O accumulator is quiescent: no extra barrier for the rescale.
//   warp 9           TMA producer: Q (256 rows) once, K and V in separate 2-stage rings.
// =====================================================================================================
constexpr int kBN2 = 128;
constexpr int kQ2Bytes = 2 * kBM * kD * 2;         // 64 KB: two atoms of [256 rows x 128 B]
constexpr int kKV2Bytes = kBN2 * kD * 2;           // 32 KB
constexpr int kAtom2 = kBN2 * 128;                 // 16 KB: one 64-column atom of a 128-row tile
constexpr int kFwd2Threads = 384;

struct Attn2Smem {
  static constexpr int q_off = 0;
  static constexpr int k_off = q_off + kQ2Bytes;
  static constexpr int v_off = k_off + 2 * kKV2Bytes;
  static constexpr int bar_off = v_off + 2 * kKV2Bytes;
  static constexpr int total = bar_off + 24 * 8;
};
constexpr int kAttn2SmemBytes = Attn2Smem::total + 1024;

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// (a.x*b + c, a.y*b + c) with one packed FFMA2
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long ua, ub, uc, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long ua, ub, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}
// 2^x for a pair on the FMA pipe (no MUFU): round-to-nearest split x = n + r, |r| <= 0.5, degree-3 minimax
// polynomial for 2^r (max relative error 7.5e-5, far below the bf16 rounding of P), exponent patched in with
// an integer add.  The softmax is MUFU-bound (16 ex2/clk/SM against 8192 MMA FLOP/clk/SM), so a fraction of
// every tile's exponentials is moved here.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 j = add2(x, make_float2(12582912.f, 12582912.f));            // 1.5 * 2^23: integer part in the low bits
  const float2 t = add2(j, make_float2(-12582912.f, -12582912.f));
  const float2 r = fma2(t, make_float2(-1.f, -1.f), x);
  float2 q = fma2(r, make_float2(0.05517074465751648f, 0.05517074465751648f),
                  make_float2(0.24260912835597992f, 0.24260912835597992f));
  q = fma2(q, r, make_float2(0.6932609677314758f, 0.6932609677314758f));
  q = fma2(q, r, make_float2(0.9999281764030457f, 0.9999281764030457f));
  q.x = __uint_as_float(__float_as_uint(q.x) + (__float_as_uint(j.x) << 23));
  q.y = __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(j.y) << 23));
  return q;
}

template <int kPolyOf4>      // of every 4 element pairs, this many take the polynomial instead of MUFU.EX2
__global__ void __launch_bounds__(kFwd2Threads, 1)
attn_fwd2_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_q,      // box {64, 256}
                         const __grid_constant__ CUtensorMap tma_kv,     // box {64, 128}
                         const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_smem = smem_base + Attn2Smem::q_off;
  const uint32_t k_smem = smem_base + Attn2Smem::k_off;
  const uint32_t v_smem = smem_base + Attn2Smem::v_off;
  const uint32_t bars = smem_base + Attn2Smem::bar_off;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (3 + s); };
  auto v_full = [&](int s) { return bars + 8u * (5 + s); };
  auto v_empty = [&](int s) { return bars + 8u * (7 + s); };
  auto s_full = [&](int w) { return bars + 8u * (9 + w); };
  auto p_full = [&](int w) { return bars + 8u * (11 + w); };
  auto o_full = [&](int w) { return bars + 8u * (13 + w); };
  const uint32_t tmem_ptr_smem = bars + 8u * 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Attn2Smem::bar_off + 8 * 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_tile = p.batch * p.n_heads;
  const int pairs = p.seq / (2 * kBM);
  const int t = pairs - 1 - (int)(blockIdx.x / per_tile);          // heavy (late) tile pairs first
  const int bh = (int)(blockIdx.x % per_tile);
  const int b = bh / p.n_heads, h = bh % p.n_heads;
  const int hk = h / (p.n_heads / p.n_kv);
  const int n0 = 2 * t + 1, n1 = 2 * t + 2;                         // key tiles seen by Q tile 0 / 1
  const int row0 = b * p.seq + t * 2 * kBM;
  const int q_col = h * kD, k_col = (p.n_heads + hk) * kD, v_col = (p.n_heads + p.n_kv + hk) * kD;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1); mbar_init(p_full(s), 4); mbar_init(o_full(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 9 && lane == 0) {
      // ====================== TMA producer ======================
      mbar_arrive_expect_tx(q_full, kQ2Bytes);
      tma_load_2d(q_smem, &tma_q, q_full, q_col, row0);
      tma_load_2d(q_smem + 2 * kAtom2, &tma_q, q_full, q_col + 64, row0);
      for (int j = 0; j < n1; ++j) {
        const int st = j & 1;
        const uint32_t par = (((uint32_t)(j >> 1)) & 1u) ^ 1u;
        const int krow = b * p.seq + j * kBN2;
        mbar_wait(k_empty(st), par);
        mbar_arrive_expect_tx(k_full(st), kKV2Bytes);
        tma_load_2d(k_smem + st * kKV2Bytes, &tma_kv, k_full(st), k_col, krow);
        tma_load_2d(k_smem + st * kKV2Bytes + kAtom2, &tma_kv, k_full(st), k_col + 64, krow);
        mbar_wait(v_empty(st), par);
        mbar_arrive_expect_tx(v_full(st), kKV2Bytes);
        tma_load_2d(v_smem + st * kKV2Bytes, &tma_kv, v_full(st), v_col, krow);
        tma_load_2d(v_smem + st * kKV2Bytes + kAtom2, &tma_kv, v_full(st), v_col + 64, krow);
      }
    } else if (warp == 8) {
      // ====================== MMA issuer: the whole warp walks the loop, one elected lane issues ========================
      const uint32_t idesc_s = umma_idesc_bf16(kBM, kBN2, 0, 0);   // S = Q K^T
      const uint32_t idesc_o = umma_idesc_bf16(kBM, kD, 0, 1);     // O += P V : V MN-major
      const uint64_t dq0 = umma_smem_desc(q_smem, 16, 1024), dq1 = umma_smem_desc(q_smem + kAtom2, 16, 1024);
      const uint64_t dk0 = umma_smem_desc(k_smem, 16, 1024), dv0 = umma_smem_desc(v_smem, kAtom2, 1024);
      auto issue_s = [&](int w, int j) {                           // S_w = Q_w K_j^T into TMEM columns 128w..
        const uint64_t kd = dk0 + (uint64_t)((j & 1) * (kKV2Bytes >> 4));
        const uint64_t qd = w ? dq1 : dq0;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint64_t ad = qd + (uint64_t)((((kk >> 2) * (2 * kAtom2)) + (kk & 3) * 32) >> 4);
            const uint64_t bd = kd + (uint64_t)((((kk >> 2) * kAtom2) + (kk & 3) * 32) >> 4);
            umma_bf16_ss(tmem_base + 128u * w, ad, bd, idesc_s, kk != 0 ? 1u : 0u);
          }
          umma_commit(s_full(w));
        }
        __syncwarp();
      };
      auto issue_pv = [&](int w, int j) {
        const uint64_t vd = dv0 + (uint64_t)((j & 1) * (kKV2Bytes >> 4));
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kBN2 / 16; ++kk)
            umma_bf16_ts(tmem_base + 256u + 128u * w, tmem_base + 128u * w + kk * 8, vd + (uint64_t)((kk * 16 * 128) >> 4),
                         idesc_o, (j != 0 || kk != 0) ? 1u : 0u);
        }
        __syncwarp();
      };
      auto commit = [&](uint32_t bar) {
        if (elect_one_sync()) umma_commit(bar);
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(k_full(0), 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      commit(k_empty(0));
      for (int j = 0; j < n1; ++j) {
        const int st = j & 1, nst = (j + 1) & 1;
        mbar_wait(v_full(st), (uint32_t)(j >> 1) & 1u);
        if (j < n0) {
          mbar_wait(p_full(0), (uint32_t)j & 1u);
          tc_fence_after();
          issue_pv(0, j);
          if (j == n0 - 1) commit(o_full(0));
        }
        if (j + 1 < n0) {
          mbar_wait(k_full(nst), (uint32_t)((j + 1) >> 1) & 1u);
          tc_fence_after();
          issue_s(0, j + 1);
        }
        mbar_wait(p_full(1), (uint32_t)j & 1u);
        tc_fence_after();
        issue_pv(1, j);
        commit(v_empty(st));
        if (j == n1 - 1) commit(o_full(1));
        if (j + 1 < n1) {
          mbar_wait(k_full(nst), (uint32_t)((j + 1) >> 1) & 1u);
          tc_fence_after();
          issue_s(1, j + 1);
          commit(k_empty(nst));
        }
      }
    }
  } else {
    // ====================== softmax warpgroups (w = 0, 1) ==========
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int w = warp >> 2;
    const int q4 = warp & 3;
    const int r_local = q4 * 32 + lane;
    const int q_pos = t * 2 * kBM + w * kBM + r_local;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const uint32_t tmem_s = tmem_base + 128u * w, tmem_o = tmem_base + 256u + 128u * w;
    const int n_w = w == 0 ? n0 : n1;
    const float sl2 = p.scale_log2;
    float m_used = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_w; ++j) {
      mbar_wait(s_full(w), (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t sv[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv[c]);
      tmem_ld_wait();
      if (j == n_w - 1) {                                           // diagonal tile: key c is visible iff c <= row
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e > r_local) sv[c][e] = 0xff800000u;       // -inf
      }
      if (j == 0) {                                                 // first tile: the reference max is this tile's max
        float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          a0 = max3(a0, __uint_as_float(sv[0][e]), __uint_as_float(sv[0][e + 1]));
          a1 = max3(a1, __uint_as_float(sv[1][e]), __uint_as_float(sv[1][e + 1]));
          a2 = max3(a2, __uint_as_float(sv[2][e]), __uint_as_float(sv[2][e + 1]));
          a3 = max3(a3, __uint_as_float(sv[3][e]), __uint_as_float(sv[3][e + 1]));
        }
        m_used = max3(a0, a1, fmaxf(a2, a3)) * sl2;
      }
      // Optimistic pass: exponentials against the CURRENT reference max while the tile max is tracked on the
      // side (four independent max chains, off the MUFU-bound critical path).  The reference only moves when
      // the max grew by more than 2^8 -- rare after the first tiles -- and then the tile is redone below.
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      float sum0 = 0.f, sum1 = 0.f;
      {
        const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-m_used, -m_used);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float s0 = __uint_as_float(sv[c][e]), s1 = __uint_as_float(sv[c][e + 1]);
            mxa[c] = max3(mxa[c], s0, s1);
            const float2 x = fma2(make_float2(s0, s1), sc2, nm2);
            float p0, p1;
            if (((e >> 1) & 3) < kPolyOf4) {
              const float2 pp = exp2_poly2(x);
              p0 = pp.x;
              p1 = pp.y;
            } else {
              p0 = fast_exp2(x.x);
              p1 = fast_exp2(x.y);
            }
            sum0 += p0;
            sum1 += p1;
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
          tmem_st_32x16(tmem_s + lane_addr + c * 16, pk);
        }
      }
      const float mx2 = max3(mxa[0], mxa[1], fmaxf(mxa[2], mxa[3])) * sl2;
      const bool need = mx2 > m_used + 8.f;
      if (__any_sync(0xffffffffu, need)) {
        // slow path: move the reference, redo the exponentials of this tile from the S values still in
        // registers, rescale the running sum and (after the first tile) the O accumulator.  PV_{j-1} of this
        // tile retired before S_j was produced (in-order tensor pipe), so O is quiescent here.
        const float m_next = need ? mx2 : m_used;
        const float alpha = need ? fast_exp2(m_used - m_next) : 1.f;
        m_used = m_next;
        l_run *= alpha;
        sum0 = 0.f;
        sum1 = 0.f;
        const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-m_used, -m_used);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 x = fma2(make_float2(__uint_as_float(sv[c][e]), __uint_as_float(sv[c][e + 1])), sc2, nm2);
            const float p0 = fast_exp2(x.x), p1 = fast_exp2(x.y);
            sum0 += p0;
            sum1 += p1;
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
          tmem_st_32x16(tmem_s + lane_addr + c * 16, pk);
        }
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < kD / 16; ++c) {
            uint32_t tt[16];
            tmem_ld_32x16(tmem_o + lane_addr + c * 16, tt);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) tt[i] = __float_as_uint(__uint_as_float(tt[i]) * alpha);
            tmem_st_32x16(tmem_o + lane_addr + c * 16, tt);
          }
        }
      }
      l_run += sum0 + sum1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(w));
    }
    mbar_wait(o_full(w), 0);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    bf16* orow = p.out + (size_t)(row0 + w * kBM + r_local) * p.out_stride + h * kD;
#pragma unroll
    for (int c = 0; c < kD / 32; ++c) {
      uint32_t tt[32];
      tmem_ld_32x32(tmem_o + lane_addr + c * 32, tt);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(tt[e + i]) * inv;
        st_bf16x8(orow + c * 32 + e, pack8(f));
      }
    }
    p.lse[((size_t)b * p.n_heads + h) * p.seq + q_pos] =
        (l_run > 0.f) ? (m_used + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================
// Forward v3 = v2 made PERSISTENT: one CTA per SM walks a heavy-first, snake-ordered list of (tile pair,
// batch, head) work items, so TMEM allocation, barrier set-up, the Q load and the O write-back of an item
// overlap the MMAs / exponentials of its neighbours instead of costing ~16 us per item (v2 measured
// time = 29.6 kclk + 3.3 kclk * key tiles per CTA).  A third warpgroup takes the epilogue off the softmax
// warpgroups: it drains O_w from TMEM, normalises with the row statistics the softmax warpgroup left in
// shared memory, and writes O and LSE while the softmax warpgroups are already on the next item.
//
//   warps 0-3 / 4-7  softmax of Q tile 0 / 1 (as v2)          warps 8-11  epilogue / write-back
//   warp 12          tcgen05.mma issuer                          warp 13     TMA producer
// =====================================================================================================
constexpr int kFwd3Threads = 512;
struct Attn3Smem {
  static constexpr int q_off = 0;
  static constexpr int k_off = q_off + kQ2Bytes;
  static constexpr int v_off = k_off + 2 * kKV2Bytes;
  static constexpr int stat_off = v_off + 2 * kKV2Bytes;       // [2 tiles][128 rows] x (1/l, lse)
  static constexpr int bar_off = stat_off + 2 * 128 * 8;
  static constexpr int total = bar_off + 32 * 8;
};
constexpr int kAttn3SmemBytes = Attn3Smem::total + 1024;

template <int kPolyOf4>
__global__ void __launch_bounds__(kFwd3Threads, 1)
attn_fwd3_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_q,      // box {64, 256}
                         const __grid_constant__ CUtensorMap tma_kv,     // box {64, 128}
                         const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_smem = smem_base + Attn3Smem::q_off;
  const uint32_t k_smem = smem_base + Attn3Smem::k_off;
  const uint32_t v_smem = smem_base + Attn3Smem::v_off;
  float2* stats = reinterpret_cast<float2*>(smem_gen + Attn3Smem::stat_off);
  const uint32_t bars = smem_base + Attn3Smem::bar_off;
  const uint32_t q_full = bars, q_empty = bars + 8u;
  auto k_full = [&](int s) { return bars + 8u * (2 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (4 + s); };
  auto v_full = [&](int s) { return bars + 8u * (6 + s); };
  auto v_empty = [&](int s) { return bars + 8u * (8 + s); };
  auto s_full = [&](int w) { return bars + 8u * (10 + w); };
  auto p_full = [&](int w) { return bars + 8u * (12 + w); };
  auto o_full = [&](int w) { return bars + 8u * (14 + w); };
  auto o_free = [&](int w) { return bars + 8u * (16 + w); };
  auto st_full = [&](int w) { return bars + 8u * (18 + w); };
  const uint32_t tmem_ptr_smem = bars + 8u * 24;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Attn3Smem::bar_off + 8 * 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_tile = p.batch * p.n_heads;
  const int pairs = p.seq / (2 * kBM);
  const int n_items = pairs * per_tile;
  const int G = (int)gridDim.x;
  // item of round k for this CTA: snake order over the heavy-first list keeps the per-CTA totals within ~2 %
  auto item_of = [&](int k) { return k * G + ((k & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x); };

  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1); mbar_init(p_full(s), 4); mbar_init(o_full(s), 1);
      mbar_init(o_free(s), 4); mbar_init(st_full(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == 13 && lane == 0) {
      // ====================== TMA producer ======================
      uint32_t kc = 0;                                             // running K/V tile count (ring position)
      for (int k = 0;; ++k) {
        const int item = item_of(k);
        if (item >= n_items) break;
        const int t = pairs - 1 - item / per_tile, bh = item % per_tile;
        const int b = bh / p.n_heads, h = bh % p.n_heads, hk = h / (p.n_heads / p.n_kv);
        const int row0 = b * p.seq + t * 2 * kBM;
        const int q_col = h * kD, k_col = (p.n_heads + hk) * kD, v_col = (p.n_heads + p.n_kv + hk) * kD;
        const int n1 = 2 * t + 2;
        mbar_wait(q_empty, ((uint32_t)k & 1u) ^ 1u);               // the S MMAs of the previous item have read Q
        mbar_arrive_expect_tx(q_full, kQ2Bytes);
        tma_load_2d(q_smem, &tma_q, q_full, q_col, row0);
        tma_load_2d(q_smem + 2 * kAtom2, &tma_q, q_full, q_col + 64, row0);
        for (int j = 0; j < n1; ++j, ++kc) {
          const int st = kc & 1;
          const uint32_t par = ((kc >> 1) & 1u) ^ 1u;
          const int krow = b * p.seq + j * kBN2;
          mbar_wait(k_empty(st), par);
          mbar_arrive_expect_tx(k_full(st), kKV2Bytes);
          tma_load_2d(k_smem + st * kKV2Bytes, &tma_kv, k_full(st), k_col, krow);
          tma_load_2d(k_smem + st * kKV2Bytes + kAtom2, &tma_kv, k_full(st), k_col + 64, krow);
          mbar_wait(v_empty(st), par);
          mbar_arrive_expect_tx(v_full(st), kKV2Bytes);
          tma_load_2d(v_smem + st * kKV2Bytes, &tma_kv, v_full(st), v_col, krow);
          tma_load_2d(v_smem + st * kKV2Bytes + kAtom2, &tma_kv, v_full(st), v_col + 64, krow);
        }
      }
    } else if (warp == 12) {
      // ====================== MMA issuer: the whole warp walks the loop, one elected lane issues ========================
      const uint32_t idesc_s = umma_idesc_bf16(kBM, kBN2, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(kBM, kD, 0, 1);
      // descriptors are built once; per MMA only the 14-bit start-address field (16-byte units) moves
      const uint64_t dq0 = umma_smem_desc(q_smem, 16, 1024), dq1 = umma_smem_desc(q_smem + kAtom2, 16, 1024);
      const uint64_t dk0 = umma_smem_desc(k_smem, 16, 1024), dv0 = umma_smem_desc(v_smem, kAtom2, 1024);
      auto issue_s = [&](int w, uint32_t kcnt) {
        const uint64_t kd = dk0 + (uint64_t)((kcnt & 1) * (kKV2Bytes >> 4));
        const uint64_t qd = w ? dq1 : dq0;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint64_t ad = qd + (uint64_t)((((kk >> 2) * (2 * kAtom2)) + (kk & 3) * 32) >> 4);
            const uint64_t bd = kd + (uint64_t)((((kk >> 2) * kAtom2) + (kk & 3) * 32) >> 4);
            umma_bf16_ss(tmem_base + 128u * w, ad, bd, idesc_s, kk != 0 ? 1u : 0u);
          }
          umma_commit(s_full(w));
        }
        __syncwarp();
      };
      auto commit = [&](uint32_t bar) {
        if (elect_one_sync()) umma_commit(bar);
        __syncwarp();
      };
      auto issue_pv = [&](int w, uint32_t kcnt, bool first) {
        const uint64_t vd = dv0 + (uint64_t)((kcnt & 1) * (kKV2Bytes >> 4));
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kBN2 / 16; ++kk) {
            umma_bf16_ts(tmem_base + 256u + 128u * w, tmem_base + 128u * w + kk * 8, vd + (uint64_t)((kk * 16 * 128) >> 4),
                         idesc_o, (!first || kk != 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      uint32_t kc = 0;                                             // K/V tiles consumed so far (ring position of tile 0 of the item)
      uint32_t pc0 = 0, pc1 = 0;                                   // P tiles consumed per warpgroup (p_full phase)
      for (int k = 0;; ++k) {
        const int item = item_of(k);
        if (item >= n_items) break;
        const int t = pairs - 1 - item / per_tile;
        const int n0 = 2 * t + 1, n1 = 2 * t + 2;
        mbar_wait(q_full, (uint32_t)k & 1u);
        mbar_wait(k_full(kc & 1), (kc >> 1) & 1u);
        tc_fence_after();
        issue_s(0, kc);
        issue_s(1, kc);
        commit(k_empty(kc & 1));
        for (int j = 0; j < n1; ++j) {
          const uint32_t cur = kc + j, nxt = cur + 1;
          mbar_wait(v_full(cur & 1), (cur >> 1) & 1u);
          if (j < n0) {
            if (j == 0 && k > 0) mbar_wait(o_free(0), (uint32_t)(k - 1) & 1u);   // O0 of the previous item written back
            mbar_wait(p_full(0), pc0 & 1u);
            ++pc0;
            tc_fence_after();
            issue_pv(0, cur, j == 0);
            if (j == n0 - 1) commit(o_full(0));
          }
          if (j + 1 < n0) {
            mbar_wait(k_full(nxt & 1), (nxt >> 1) & 1u);
            tc_fence_after();
            issue_s(0, nxt);
          }
          if (j == 0 && k > 0) mbar_wait(o_free(1), (uint32_t)(k - 1) & 1u);
          mbar_wait(p_full(1), pc1 & 1u);
          ++pc1;
          tc_fence_after();
          issue_pv(1, cur, j == 0);
          commit(v_empty(cur & 1));
          if (j == n1 - 1) commit(o_full(1));
          if (j + 1 < n1) {
            mbar_wait(k_full(nxt & 1), (nxt >> 1) & 1u);
            tc_fence_after();
            issue_s(1, nxt);
            commit(k_empty(nxt & 1));
            if (j + 1 == n1 - 1) commit(q_empty);             // last S MMAs of this item: Q may be replaced
          }
        }
        kc += (uint32_t)n1;
      }
    }
  } else if (warp >= 8) {
    // ====================== epilogue warpgroup: O_w / l -> bf16 -> global, LSE ==========
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int q4 = warp & 3;
    const int r_local = q4 * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    for (int k = 0;; ++k) {
      const int item = item_of(k);
      if (item >= n_items) break;
      const int t = pairs - 1 - item / per_tile, bh = item % per_tile;
      const int b = bh / p.n_heads, h = bh % p.n_heads;
      const int row0 = b * p.seq + t * 2 * kBM;
#pragma unroll 1
      for (int w = 0; w < 2; ++w) {
        mbar_wait(st_full(w), (uint32_t)k & 1u);
        mbar_wait(o_full(w), (uint32_t)k & 1u);
        tc_fence_after();
        const float2 st = stats[w * 128 + r_local];
        bf16* orow = p.out + (size_t)(row0 + w * kBM + r_local) * p.out_stride + h * kD;
#pragma unroll 1
        for (int c = 0; c < kD / 32; ++c) {
          uint32_t tt[32];
          tmem_ld_32x32(tmem_base + 256u + 128u * w + lane_addr + c * 32, tt);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(tt[e + i]) * st.x;
            st_bf16x8(orow + c * 32 + e, pack8(f));
          }
        }
        p.lse[((size_t)b * p.n_heads + h) * p.seq + t * 2 * kBM + w * kBM + r_local] = st.y;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free(w));
      }
    }
  } else {
    // ====================== softmax warpgroups (w = 0, 1) ==========
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
    const int w = warp >> 2;
    const int q4 = warp & 3;
    const int r_local = q4 * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const uint32_t tmem_s = tmem_base + 128u * w, tmem_o = tmem_base + 256u + 128u * w;
    const float sl2 = p.scale_log2;
    uint32_t sc = 0;                                               // S tiles consumed so far (s_full phase)
    for (int k = 0;; ++k) {
      const int item = item_of(k);
      if (item >= n_items) break;
      const int t = pairs - 1 - item / per_tile;
      const int n_w = w == 0 ? 2 * t + 1 : 2 * t + 2;
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_w; ++j, ++sc) {
        mbar_wait(s_full(w), sc & 1u);
        tc_fence_after();
        uint32_t sv[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv[c]);
        tmem_ld_wait();
        if (j == n_w - 1) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c * 32 + e > r_local) sv[c][e] = 0xff800000u;
        }
        if (j == 0) {
          float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            a0 = max3(a0, __uint_as_float(sv[0][e]), __uint_as_float(sv[0][e + 1]));
            a1 = max3(a1, __uint_as_float(sv[1][e]), __uint_as_float(sv[1][e + 1]));
            a2 = max3(a2, __uint_as_float(sv[2][e]), __uint_as_float(sv[2][e + 1]));
            a3 = max3(a3, __uint_as_float(sv[3][e]), __uint_as_float(sv[3][e + 1]));
          }
          m_used = max3(a0, a1, fmaxf(a2, a3)) * sl2;
        }
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        float sum0 = 0.f, sum1 = 0.f;
        {
          const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-m_used, -m_used);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float s0 = __uint_as_float(sv[c][e]), s1 = __uint_as_float(sv[c][e + 1]);
              mxa[c] = max3(mxa[c], s0, s1);
              const float2 x = fma2(make_float2(s0, s1), sc2, nm2);
              float p0, p1;
              if (((e >> 1) & 3) < kPolyOf4) {
                const float2 pp = exp2_poly2(x);
                p0 = pp.x;
                p1 = pp.y;
              } else {
                p0 = fast_exp2(x.x);
                p1 = fast_exp2(x.y);
              }
              sum0 += p0;
              sum1 += p1;
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              pk[e >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
            tmem_st_32x16(tmem_s + lane_addr + c * 16, pk);
          }
        }
        const float mx2 = max3(mxa[0], mxa[1], fmaxf(mxa[2], mxa[3])) * sl2;
        const bool need = mx2 > m_used + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_next = need ? mx2 : m_used;
          const float alpha = need ? fast_exp2(m_used - m_next) : 1.f;
          m_used = m_next;
          l_run *= alpha;
          sum0 = 0.f;
          sum1 = 0.f;
          const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-m_used, -m_used);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float2 x = fma2(make_float2(__uint_as_float(sv[c][e]), __uint_as_float(sv[c][e + 1])), sc2, nm2);
              const float p0 = fast_exp2(x.x), p1 = fast_exp2(x.y);
              sum0 += p0;
              sum1 += p1;
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              pk[e >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
            tmem_st_32x16(tmem_s + lane_addr + c * 16, pk);
          }
          if (j > 0) {
            // PV_{j-1} of this tile retired before S_j was produced (in-order tensor pipe): O is quiescent; the
            // epilogue warpgroup only touches O after o_full, i.e. after this item's last PV
#pragma unroll 1
            for (int c = 0; c < kD / 16; ++c) {
              uint32_t tt[16];
              tmem_ld_32x16(tmem_o + lane_addr + c * 16, tt);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) tt[i] = __float_as_uint(__uint_as_float(tt[i]) * alpha);
              tmem_st_32x16(tmem_o + lane_addr + c * 16, tt);
            }
          }
        }
        l_run += sum0 + sum1;
        if (j == n_w - 1) {
          // row statistics for the epilogue warpgroup; the slot is free once the previous item's O_w was written back
          if (k > 0) mbar_wait(o_free(w), (uint32_t)(k - 1) & 1u);
          stats[w * 128 + r_local] = make_float2(l_run > 0.f ? 1.f / l_run : 0.f,
                                                 l_run > 0.f ? (m_used + log2f(l_run)) * 0.6931471805599453f : -INFINITY);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (j == n_w - 1) mbar_arrive(st_full(w));
          mbar_arrive(p_full(w));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================
// Forward v4 = v2 with TWO softmax warpgroups per Q tile (each owns 64 of the 128 key columns of a tile).
// ncu on v2/v3: tensor pipe 41 %, MUFU 30 %, softmax warps at ~0.3 IPC -- one warp per scheduler cannot hide
// the 4-cycle ALU / MUFU latencies of its own dependency chains, so the exponentials of a tile took
// ~2200 cycles against 1024 cycles of MMAs.  With 16 softmax warps every scheduler interleaves four.
// The two halves of a row agree on the reference max through shared memory (one named barrier per tile);
// P stays in registers until they have agreed, so the rare "reference moved" path re-reads S from TMEM.
//   warps 0-15  softmax: warpgroup g -> Q tile g>>1, column half g&1        warp 16  MMA issuer    warp 17  TMA
// =====================================================================================================
constexpr int kFwd4Threads = 640;
struct Attn4Smem {
  static constexpr int q_off = 0;
  static constexpr int k_off = q_off + kQ2Bytes;
  static constexpr int v_off = k_off + 2 * kKV2Bytes;
  static constexpr int xch_off = v_off + 2 * kKV2Bytes;        // [2 tiles][2 halves][128 rows] floats
  static constexpr int bar_off = xch_off + 2 * 2 * 128 * 4;
  static constexpr int total = bar_off + 24 * 8;
};
constexpr int kAttn4SmemBytes = Attn4Smem::total + 1024;

template <int kPolyOf4>
__global__ void __launch_bounds__(kFwd4Threads, 1)
attn_fwd4_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_q,      // box {64, 256}
                         const __grid_constant__ CUtensorMap tma_kv,     // box {64, 128}
                         const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_smem = smem_base + Attn4Smem::q_off;
  const uint32_t k_smem = smem_base + Attn4Smem::k_off;
  const uint32_t v_smem = smem_base + Attn4Smem::v_off;
  volatile float* xch = reinterpret_cast<volatile float*>(smem_gen + Attn4Smem::xch_off);
  const uint32_t bars = smem_base + Attn4Smem::bar_off;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (3 + s); };
  auto v_full = [&](int s) { return bars + 8u * (5 + s); };
  auto v_empty = [&](int s) { return bars + 8u * (7 + s); };
  auto s_full = [&](int w) { return bars + 8u * (9 + w); };
  auto p_full = [&](int w) { return bars + 8u * (11 + w); };
  auto o_full = [&](int w) { return bars + 8u * (13 + w); };
  const uint32_t tmem_ptr_smem = bars + 8u * 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Attn4Smem::bar_off + 8 * 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_tile = p.batch * p.n_heads;
  const int pairs = p.seq / (2 * kBM);
  const int t = pairs - 1 - (int)(blockIdx.x / per_tile);
  const int bh = (int)(blockIdx.x % per_tile);
  const int b = bh / p.n_heads, h = bh % p.n_heads;
  const int hk = h / (p.n_heads / p.n_kv);
  const int n0 = 2 * t + 1, n1 = 2 * t + 2;
  const int row0 = b * p.seq + t * 2 * kBM;
  const int q_col = h * kD, k_col = (p.n_heads + hk) * kD, v_col = (p.n_heads + p.n_kv + hk) * kD;

  if (warp == 17 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1); mbar_init(p_full(s), 8); mbar_init(o_full(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 17 && lane == 0) {
      // ====================== TMA producer ======================
      mbar_arrive_expect_tx(q_full, kQ2Bytes);
      tma_load_2d(q_smem, &tma_q, q_full, q_col, row0);
      tma_load_2d(q_smem + 2 * kAtom2, &tma_q, q_full, q_col + 64, row0);
      for (int j = 0; j < n1; ++j) {
        const int st = j & 1;
        const uint32_t par = (((uint32_t)(j >> 1)) & 1u) ^ 1u;
        const int krow = b * p.seq + j * kBN2;
        mbar_wait(k_empty(st), par);
        mbar_arrive_expect_tx(k_full(st), kKV2Bytes);
        tma_load_2d(k_smem + st * kKV2Bytes, &tma_kv, k_full(st), k_col, krow);
        tma_load_2d(k_smem + st * kKV2Bytes + kAtom2, &tma_kv, k_full(st), k_col + 64, krow);
        mbar_wait(v_empty(st), par);
        mbar_arrive_expect_tx(v_full(st), kKV2Bytes);
        tma_load_2d(v_smem + st * kKV2Bytes, &tma_kv, v_full(st), v_col, krow);
        tma_load_2d(v_smem + st * kKV2Bytes + kAtom2, &tma_kv, v_full(st), v_col + 64, krow);
      }
    } else if (warp == 16) {
      // ====================== MMA issuer: the whole warp walks the loop, one elected lane issues ========================
      const uint32_t idesc_s = umma_idesc_bf16(kBM, kBN2, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(kBM, kD, 0, 1);
      const uint64_t dq0 = umma_smem_desc(q_smem, 16, 1024), dq1 = umma_smem_desc(q_smem + kAtom2, 16, 1024);
      const uint64_t dk0 = umma_smem_desc(k_smem, 16, 1024), dv0 = umma_smem_desc(v_smem, kAtom2, 1024);
      auto issue_s = [&](int w, int j) {
        const uint64_t kd = dk0 + (uint64_t)((j & 1) * (kKV2Bytes >> 4));
        const uint64_t qd = w ? dq1 : dq0;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint64_t ad = qd + (uint64_t)((((kk >> 2) * (2 * kAtom2)) + (kk & 3) * 32) >> 4);
            const uint64_t bd = kd + (uint64_t)((((kk >> 2) * kAtom2) + (kk & 3) * 32) >> 4);
            umma_bf16_ss(tmem_base + 128u * w, ad, bd, idesc_s, kk != 0 ? 1u : 0u);
          }
          umma_commit(s_full(w));
        }
        __syncwarp();
      };
      auto commit = [&](uint32_t bar) {
        if (elect_one_sync()) umma_commit(bar);
        __syncwarp();
      };
      auto issue_pv = [&](int w, int j) {
        const uint64_t vd = dv0 + (uint64_t)((j & 1) * (kKV2Bytes >> 4));
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kBN2 / 16; ++kk) {
            // P (bf16 pairs): keys 0..63 in S columns 0..31, keys 64..127 in S columns 64..95
            const uint32_t a_tmem = tmem_base + 128u * w + (uint32_t)(kk >> 2) * 64u + (uint32_t)(kk & 3) * 8u;
            umma_bf16_ts(tmem_base + 256u + 128u * w, a_tmem, vd + (uint64_t)((kk * 16 * 128) >> 4), idesc_o,
                         (j != 0 || kk != 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(k_full(0), 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      commit(k_empty(0));
      for (int j = 0; j < n1; ++j) {
        const int st = j & 1, nst = (j + 1) & 1;
        mbar_wait(v_full(st), (uint32_t)(j >> 1) & 1u);
        if (j < n0) {
          mbar_wait(p_full(0), (uint32_t)j & 1u);
          tc_fence_after();
          issue_pv(0, j);
          if (j == n0 - 1) commit(o_full(0));
        }
        if (j + 1 < n0) {
          mbar_wait(k_full(nst), (uint32_t)((j + 1) >> 1) & 1u);
          tc_fence_after();
          issue_s(0, j + 1);
        }
        mbar_wait(p_full(1), (uint32_t)j & 1u);
        tc_fence_after();
        issue_pv(1, j);
        commit(v_empty(st));
        if (j == n1 - 1) commit(o_full(1));
        if (j + 1 < n1) {
          mbar_wait(k_full(nst), (uint32_t)((j + 1) >> 1) & 1u);
          tc_fence_after();
          issue_s(1, j + 1);
          commit(k_empty(nst));
        }
      }
    }
  } else {
    // ====================== softmax warpgroups: g -> (tile w, column half hf) ==========
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int g = warp >> 2;
    const int w = g >> 1, hf = g & 1;
    const int q4 = warp & 3;
    const int r_local = q4 * 32 + lane;
    const int q_pos = t * 2 * kBM + w * kBM + r_local;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const uint32_t tmem_s = tmem_base + 128u * w + 64u * hf;        // this half's 64 S columns (and its P, 32 columns)
    const uint32_t tmem_o = tmem_base + 256u + 128u * w + 64u * hf; // this half's 64 O columns
    volatile float* x_mine = xch + (w * 2 + hf) * 128 + r_local;
    volatile float* x_other = xch + (w * 2 + (hf ^ 1)) * 128 + r_local;
    const int bar_id = 1 + w;                                        // named barrier of the two halves of tile w
    const int n_w = w == 0 ? n0 : n1;
    const float sl2 = p.scale_log2;
    float m_used = -INFINITY, l_part = 0.f;
    for (int j = 0; j < n_w; ++j) {
      mbar_wait(s_full(w), (uint32_t)j & 1u);
      tc_fence_after();
      const bool diag = j == n_w - 1;
      uint32_t pk[32];
      float mloc = -INFINITY, sum0 = 0.f, sum1 = 0.f;
      if (j == 0) {
        // first tile: the reference is the true row max of this tile -> agree on it before any exponential
        float a0 = -INFINITY, a1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float s0 = __uint_as_float(sv[e]), s1 = __uint_as_float(sv[e + 1]);
            if (diag && hf * 64 + c * 32 + e > r_local) s0 = -INFINITY;
            if (diag && hf * 64 + c * 32 + e + 1 > r_local) s1 = -INFINITY;
            if (c == 0) a0 = max3(a0, s0, s1); else a1 = max3(a1, s0, s1);
          }
        }
        *x_mine = fmaxf(a0, a1);
        named_bar_sync(bar_id, 256);
        m_used = fmaxf(fmaxf(a0, a1), *x_other) * sl2;
        named_bar_sync(bar_id, 256);                                 // the slot is rewritten below
      }
      {
        const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-m_used, -m_used);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float s0 = __uint_as_float(sv[e]), s1 = __uint_as_float(sv[e + 1]);
            if (diag && hf * 64 + c * 32 + e > r_local) s0 = -INFINITY;
            if (diag && hf * 64 + c * 32 + e + 1 > r_local) s1 = -INFINITY;
            mloc = max3(mloc, s0, s1);
            const float2 x = fma2(make_float2(s0, s1), sc2, nm2);
            float p0, p1;
            if (((e >> 1) & 3) < kPolyOf4) {
              const float2 pp = exp2_poly2(x);
              p0 = pp.x;
              p1 = pp.y;
            } else {
              p0 = fast_exp2(x.x);
              p1 = fast_exp2(x.y);
            }
            sum0 += p0;
            sum1 += p1;
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            pk[c * 16 + (e >> 1)] = *reinterpret_cast<uint32_t*>(&hh);
          }
        }
      }
      // agree with the other half of the row on whether the reference max has to move (> 2^8 growth)
      *x_mine = mloc;
      named_bar_sync(bar_id, 256);
      const float mx2 = fmaxf(mloc, *x_other) * sl2;
      const bool need = mx2 > m_used + 8.f;
      const bool slow = __any_sync(0xffffffffu, need);               // same rows in both halves -> same decision
      if (slow) {
        const float m_next = need ? mx2 : m_used;
        const float alpha = need ? fast_exp2(m_used - m_next) : 1.f;
        m_used = m_next;
        l_part *= alpha;
        sum0 = 0.f;
        sum1 = 0.f;
        const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-m_used, -m_used);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv);            // S is still intact: P has not been stored yet
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float s0 = __uint_as_float(sv[e]), s1 = __uint_as_float(sv[e + 1]);
            if (diag && hf * 64 + c * 32 + e > r_local) s0 = -INFINITY;
            if (diag && hf * 64 + c * 32 + e + 1 > r_local) s1 = -INFINITY;
            const float2 x = fma2(make_float2(s0, s1), sc2, nm2);
            const float p0 = fast_exp2(x.x), p1 = fast_exp2(x.y);
            sum0 += p0;
            sum1 += p1;
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            pk[c * 16 + (e >> 1)] = *reinterpret_cast<uint32_t*>(&hh);
          }
        }
        if (j > 0) {
          // PV_{j-1} of this tile retired before S_j was produced (in-order tensor pipe): O is quiescent.
          // Each half rescales its own 64 of the 128 O columns.
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t tt[16];
            tmem_ld_32x16(tmem_o + lane_addr + c * 16, tt);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) tt[i] = __float_as_uint(__uint_as_float(tt[i]) * alpha);
            tmem_st_32x16(tmem_o + lane_addr + c * 16, tt);
          }
        }
      }
      l_part += sum0 + sum1;
      named_bar_sync(bar_id, 256);          // both halves have read the exchange slots (and all S reads of the tile are done)
      {
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { lo[i] = pk[i]; hi[i] = pk[16 + i]; }
        tmem_st_32x16(tmem_s + lane_addr, lo);
        tmem_st_32x16(tmem_s + lane_addr + 16, hi);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(w));
    }
    // ---- epilogue: total row sum = both halves; each half normalises and stores its 64 O columns ----
    *x_mine = l_part;
    named_bar_sync(bar_id, 256);
    const float l_run = l_part + *x_other;
    mbar_wait(o_full(w), 0);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    bf16* orow = p.out + (size_t)(row0 + w * kBM + r_local) * p.out_stride + h * kD + hf * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t tt[32];
      tmem_ld_32x32(tmem_o + lane_addr + c * 32, tt);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(tt[e + i]) * inv;
        st_bf16x8(orow + c * 32 + e, pack8(f));
      }
    }
    if (hf == 0)
      p.lse[((size_t)b * p.n_heads + h) * p.seq + q_pos] =
          (l_run > 0.f) ? (m_used + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;
std::once_flag g_attn_once;
int g_fwd_poly = 1;          // exponentials evaluated on the FMA pipe: this many of every 4 pairs (0, 1 or 2)
int g_fwd_version = 2;       // 2 = two Q tiles per CTA, one softmax warpgroup per tile (default: fastest measured);
                             // 3 = 2 made persistent with an epilogue warpgroup; 4 = 2 with two softmax warpgroups per tile;
                             // 1 = one Q tile per CTA, 64-key tiles, two CTAs per SM
std::string g_attn_err;

void attn_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_enc = reinterpret_cast<EncodeTiledFn>(fn);
  cudaFuncSetAttribute(attn_fwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
  cudaFuncSetAttribute(attn_fwd4_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn4SmemBytes);
  cudaFuncSetAttribute(attn_fwd4_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn4SmemBytes);
  cudaFuncSetAttribute(attn_fwd4_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn4SmemBytes);
  cudaFuncSetAttribute(attn_fwd3_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn3SmemBytes);
  cudaFuncSetAttribute(attn_fwd3_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn3SmemBytes);
  cudaFuncSetAttribute(attn_fwd3_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn3SmemBytes);
  cudaFuncSetAttribute(attn_fwd2_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn2SmemBytes);
  cudaFuncSetAttribute(attn_fwd2_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn2SmemBytes);
  cudaFuncSetAttribute(attn_fwd2_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn2SmemBytes);
}

bool attn_tmap(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

const char* attention_last_error() { return g_attn_err.c_str(); }
void attention_fwd_set_version(int v) { g_fwd_version = (v >= 1 && v <= 4) ? v : 2; }
int attention_fwd_get_version() { return g_fwd_version; }
void attention_fwd_set_poly(int n) { g_fwd_poly = n < 0 ? 0 : (n > 2 ? 2 : n); }

// qkv: [batch*seq, (n_heads + 2*n_kv) * 128] bf16 (row stride qkv_stride), already RoPE'd.
// out: [batch*seq, n_heads*128] bf16 (row stride out_stride); lse: [batch, n_heads, seq] fp32.
int attention_fwd_tcgen05(const bf16* qkv, bf16* out, float* lse, int batch, int seq, int n_heads, int n_kv,
                          int head_dim, int qkv_stride, int out_stride, float scale, cudaStream_t s) {
  std::call_once(g_attn_once, attn_init);
  if (g_enc == nullptr) { g_attn_err = "cuTensorMapEncodeTiled unavailable"; return 1; }
  if (head_dim != kD || seq % kBM != 0 || n_heads % n_kv != 0 || (qkv_stride % 8) || (out_stride % 8)) {
    g_attn_err = "attention_fwd_tcgen05 needs head_dim 128, seq % 128 == 0, n_heads % n_kv == 0";
    return 2;
  }
  const uint64_t cols = (uint64_t)(n_heads + 2 * n_kv) * kD, rows = (uint64_t)batch * seq;
  CUtensorMap tq, tkv;
  if (!attn_tmap(&tq, qkv, cols, rows, qkv_stride, kBM) || !attn_tmap(&tkv, qkv, cols, rows, qkv_stride, kBN)) {
    g_attn_err = "cuTensorMapEncodeTiled failed";
    return 3;
  }
  AttnParams p;
  p.batch = batch; p.seq = seq; p.n_heads = n_heads; p.n_kv = n_kv; p.q_tiles = seq / kBM;
  p.out_stride = out_stride;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.lse = lse;
  if (g_fwd_version >= 2 && seq % (2 * kBM) == 0) {
    CUtensorMap tq2, tkv2;
    if (!attn_tmap(&tq2, qkv, cols, rows, qkv_stride, 2 * kBM) || !attn_tmap(&tkv2, qkv, cols, rows, qkv_stride, kBN2)) {
      g_attn_err = "cuTensorMapEncodeTiled failed";
      return 3;
    }
    const int grid2 = (seq / (2 * kBM)) * batch * n_heads;
    if (g_fwd_version == 4) {
      if (g_fwd_poly == 0) attn_fwd4_tcgen05_kernel<0><<<grid2, kFwd4Threads, kAttn4SmemBytes, s>>>(tq2, tkv2, p);
      else if (g_fwd_poly == 1) attn_fwd4_tcgen05_kernel<1><<<grid2, kFwd4Threads, kAttn4SmemBytes, s>>>(tq2, tkv2, p);
      else attn_fwd4_tcgen05_kernel<2><<<grid2, kFwd4Threads, kAttn4SmemBytes, s>>>(tq2, tkv2, p);
    } else if (g_fwd_version == 3) {
      const int grid3 = grid2 < kNumSMs ? grid2 : kNumSMs;
      if (g_fwd_poly == 0) attn_fwd3_tcgen05_kernel<0><<<grid3, kFwd3Threads, kAttn3SmemBytes, s>>>(tq2, tkv2, p);
      else if (g_fwd_poly == 1) attn_fwd3_tcgen05_kernel<1><<<grid3, kFwd3Threads, kAttn3SmemBytes, s>>>(tq2, tkv2, p);
      else attn_fwd3_tcgen05_kernel<2><<<grid3, kFwd3Threads, kAttn3SmemBytes, s>>>(tq2, tkv2, p);
    } else if (g_fwd_poly == 0) attn_fwd2_tcgen05_kernel<0><<<grid2, kFwd2Threads, kAttn2SmemBytes, s>>>(tq2, tkv2, p);
    else if (g_fwd_poly == 1) attn_fwd2_tcgen05_kernel<1><<<grid2, kFwd2Threads, kAttn2SmemBytes, s>>>(tq2, tkv2, p);
    else attn_fwd2_tcgen05_kernel<2><<<grid2, kFwd2Threads, kAttn2SmemBytes, s>>>(tq2, tkv2, p);
  } else {
    const int grid = p.q_tiles * batch * n_heads;
    attn_fwd_tcgen05_kernel<<<grid, kAttnThreads, kAttnSmemBytes, s>>>(tq, tkv, p);
  }
  count_launch();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_attn_err = cudaGetErrorString(e); return 4; }
  return 0;
}

}  // namespace hb
