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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;
std::once_flag g_attn_once;
std::string g_attn_err;

void attn_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_enc = reinterpret_cast<EncodeTiledFn>(fn);
  cudaFuncSetAttribute(attn_fwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
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
  const int grid = p.q_tiles * batch * n_heads;
  attn_fwd_tcgen05_kernel<<<grid, kAttnThreads, kAttnSmemBytes, s>>>(tq, tkv, p);
  count_launch();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_attn_err = cudaGetErrorString(e); return 4; }
  return 0;
}

}  // namespace hb
