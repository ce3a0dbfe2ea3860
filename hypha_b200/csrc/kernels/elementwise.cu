// Memory-bound transformer ops for sm_100a: fused (residual+)RMSNorm fwd/bwd,
// RoPE, SwiGLU, fused softmax-cross-entropy fwd+bwd, embedding gather/scatter.
// All I/O is bf16 with 16-byte vector accesses; math is fp32.
//
// These replace the stock ATen ops the reference reaches through
// `model(**batch)` / `accelerator.backward(loss)`
// (reference: executors/accelerate/src/hypha/accelerate_executor/training.py:108-112).
#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"

namespace hb {

unsigned long long g_launch_count = 0;
unsigned long long launch_count() { return g_launch_count; }
void reset_launch_count() { g_launch_count = 0; }

// =============================================================================
// RMSNorm forward (optionally fused with the residual add that precedes it)
//   h = x (+ res_in) ; res_out = h ; y = h * rsqrt(mean(h^2)+eps) * w
// One CTA per row; the row stays in registers between the two passes.
// =============================================================================
template <int kThr, int kMaxVec>
__global__ void __launch_bounds__(kThr) rmsnorm_fwd_kernel(
    const bf16* __restrict__ x, const bf16* __restrict__ res_in, bf16* __restrict__ res_out,
    const bf16* __restrict__ w, bf16* __restrict__ y, float* __restrict__ rstd_out, int cols,
    float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int nvec = cols >> 3;
  const size_t base = (size_t)row * cols;
  float h[kMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kThr;
    if (i < nvec) {
      bf16x8 v = ld_bf16x8(x + base + (size_t)i * 8);
      unpack8(v, h[j]);
      if (res_in != nullptr) {
        float r[8];
        unpack8(ld_bf16x8(res_in + base + (size_t)i * 8), r);
#pragma unroll
        for (int e = 0; e < 8; ++e) h[j][e] += r[e];
        bf16x8 hv = pack8(h[j]);
        st_bf16x8(res_out + base + (size_t)i * 8, hv);
        unpack8(hv, h[j]);  // normalise exactly what backward will re-read
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += h[j][e] * h[j][e];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / (float)cols + eps);
  if (threadIdx.x == 0 && rstd_out != nullptr) rstd_out[row] = rstd;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kThr;
    if (i < nvec) {
      float wv[8], o[8];
      unpack8(ld_bf16x8(w + (size_t)i * 8), wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = h[j][e] * rstd * wv[e];
      st_bf16x8(y + base + (size_t)i * 8, pack8(o));
    }
  }
}

// -----------------------------------------------------------------------------------------------------------
// Row-streaming variants (the default for hidden sizes that fit: 16-byte aligned rows of <= 16 KB): rows arrive in
// a 4-deep shared-memory ring through `cp.async.bulk` (the TMA engine, completion on an mbarrier), so HBM reads of
// the next rows are in flight while the block reduces / writes the current one; every thread owns fixed columns
// (weights and, in backward, the dW accumulators stay in registers).  The block-per-row kernels above leave the
// memory system idle during their reduction and ran at 0.40-0.67 of the measured HBM copy rate.
// -----------------------------------------------------------------------------------------------------------
constexpr int kNormStages = 4;
constexpr int kNormThr = 256;

__device__ __forceinline__ bf16x8 lds_bf16x8(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return *reinterpret_cast<bf16x8*>(&r);
}

template <int kVec, bool kRes>
__global__ void __launch_bounds__(kNormThr, 2) rmsnorm_fwd_stream_kernel(
    const bf16* __restrict__ x, const bf16* __restrict__ res_in, bf16* __restrict__ res_out,
    const bf16* __restrict__ w, bf16* __restrict__ y, float* __restrict__ rstd_out, int rows, int cols, float eps) {
  using namespace sm100;
  extern __shared__ __align__(128) uint8_t norm_smem[];
  __shared__ float red[2][kNormThr / 32];
  __shared__ __align__(8) uint64_t bars[kNormStages];
  const uint32_t row_bytes = (uint32_t)cols * 2;
  const uint32_t stage_bytes = row_bytes * (kRes ? 2 : 1);
  const uint32_t smem0 = smem_u32(norm_smem);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nvec = cols >> 3;
  const int my_rows = (rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int k) {                      // thread 0: fetch this block's k-th row into stage k % kNormStages
    const int st = k % kNormStages;
    const size_t off = ((size_t)blockIdx.x + (size_t)k * gridDim.x) * cols;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_arrive_expect_tx(bar, stage_bytes);
    bulk_load_1d(smem0 + st * stage_bytes, x + off, row_bytes, bar);
    if (kRes) bulk_load_1d(smem0 + st * stage_bytes + row_bytes, res_in + off, row_bytes, bar);
  };
  if (tid == 0) {
    for (int st = 0; st < kNormStages; ++st) mbar_init(smem_u32(&bars[st]), 1);
    fence_barrier_init();
    for (int k = 0; k < kNormStages && k < my_rows; ++k) issue(k);
  }
  float wv[kVec][8];
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const int i = tid + j * kNormThr;
#pragma unroll
    for (int e = 0; e < 8; ++e) wv[j][e] = 0.f;
    if (i < nvec) unpack8(ld_bf16x8(w + (size_t)i * 8), wv[j]);
  }
  __syncthreads();
  for (int k = 0; k < my_rows; ++k) {
    const int st = k % kNormStages;
    const size_t base = ((size_t)blockIdx.x + (size_t)k * gridDim.x) * cols;
    mbar_wait(smem_u32(&bars[st]), (uint32_t)(k / kNormStages) & 1u);
    float h[kVec][8];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const int i = tid + j * kNormThr;
      if (i < nvec) {
        unpack8(lds_bf16x8(smem0 + st * stage_bytes + (uint32_t)i * 16), h[j]);
        if (kRes) {
          float r[8];
          unpack8(lds_bf16x8(smem0 + st * stage_bytes + row_bytes + (uint32_t)i * 16), r);
#pragma unroll
          for (int e = 0; e < 8; ++e) h[j][e] += r[e];
          const bf16x8 hv = pack8(h[j]);
          st_bf16x8(res_out + base + (size_t)i * 8, hv);
          unpack8(hv, h[j]);                     // normalise exactly what backward will re-read
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += h[j][e] * h[j][e];
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) red[k & 1][wid] = ss;
    __syncthreads();                             // also: every thread has read stage st -> it can be refilled
    if (tid == 0 && k + kNormStages < my_rows) issue(k + kNormStages);
    float tot = lane < kNormThr / 32 ? red[k & 1][lane] : 0.f;
    tot = warp_sum(tot);
    const float rstd = rsqrtf(tot / (float)cols + eps);
    if (tid == 0 && rstd_out != nullptr) rstd_out[(size_t)blockIdx.x + (size_t)k * gridDim.x] = rstd;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const int i = tid + j * kNormThr;
      if (i < nvec) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = h[j][e] * rstd * wv[j][e];
        st_bf16x8(y + base + (size_t)i * 8, pack8(o));
      }
    }
  }
}

static bool norm_stream_ok(int cols, const void* a, const void* b) {
  return cols % 8 == 0 && cols / 8 <= kNormThr * 4 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}
static int norm_stream_grid(int rows) {
  const int cap = kNumSMs * 2;
  return rows < cap ? rows : cap;
}

void rmsnorm_fwd(const bf16* x, const bf16* res_in, bf16* res_out, const bf16* w, bf16* y,
                 float* rstd, int rows, int cols, float eps, cudaStream_t s) {
  if (rows == 0) return;
  const int nvec = cols / 8;
  if (norm_stream_ok(cols, x, res_in)) {
    const int grid = norm_stream_grid(rows);
    const int kv = (nvec + kNormThr - 1) / kNormThr;
    const size_t smem = (size_t)kNormStages * cols * 2 * (res_in != nullptr ? 2 : 1);
#define HB_FWD_STREAM(KV, RES)                                                                                  \
    do {                                                                                                        \
      static bool attr = false;                                                                                 \
      if (!attr) { cudaFuncSetAttribute(rmsnorm_fwd_stream_kernel<KV, RES>,                                     \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 * 2); attr = true; } \
      rmsnorm_fwd_stream_kernel<KV, RES><<<grid, kNormThr, smem, s>>>(x, res_in, res_out, w, y, rstd, rows, cols, eps); \
    } while (0)
    if (res_in != nullptr) {
      if (kv == 1) HB_FWD_STREAM(1, true); else if (kv == 2) HB_FWD_STREAM(2, true); else HB_FWD_STREAM(4, true);
    } else {
      if (kv == 1) HB_FWD_STREAM(1, false); else if (kv == 2) HB_FWD_STREAM(2, false); else HB_FWD_STREAM(4, false);
    }
#undef HB_FWD_STREAM
    count_launch();
    return;
  }
  // one 16-byte vector per thread whenever possible: small register footprint -> full occupancy
  if (nvec <= 256) {
    rmsnorm_fwd_kernel<256, 1><<<rows, 256, 0, s>>>(x, res_in, res_out, w, y, rstd, cols, eps);
  } else if (nvec <= 512) {
    rmsnorm_fwd_kernel<512, 1><<<rows, 512, 0, s>>>(x, res_in, res_out, w, y, rstd, cols, eps);
  } else if (nvec <= 1024) {
    rmsnorm_fwd_kernel<512, 2><<<rows, 512, 0, s>>>(x, res_in, res_out, w, y, rstd, cols, eps);
  } else {
    rmsnorm_fwd_kernel<512, 4><<<rows, 512, 0, s>>>(x, res_in, res_out, w, y, rstd, cols, eps);
  }
  count_launch();
}

// =============================================================================
// RMSNorm backward. Persistent CTAs stride over rows; each thread owns fixed
// columns so dW accumulates in registers and is written once per CTA as an
// fp32 partial row (reduced by colsum_partials).
//   xhat = x*rstd ; g = dy*w ; c = mean(g*xhat)
//   dx = rstd*(g - xhat*c) (+ dres_in) ; dW += dy*xhat
// =============================================================================
// kRows rows are in flight per block iteration: their loads are all issued before the (single, kRows-wide)
// block reduction, so one DRAM round trip and one pair of __syncthreads is paid per kRows rows instead of per
// row (round 1: one row per iteration ran at 0.40 of the measured HBM copy rate).
template <int kThr, int kMaxVec, int kRows>
__global__ void __launch_bounds__(kThr) rmsnorm_bwd_kernel(
    const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
    const float* __restrict__ rstd, const bf16* __restrict__ dres_in, bf16* __restrict__ dx,
    float* __restrict__ dw_partial, int rows, int cols) {
  __shared__ float red[kRows][32];
  const int nvec = cols >> 3;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int kWarps = kThr / 32;
  float wv[kMaxVec][8], dwacc[kMaxVec][8];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kThr;
#pragma unroll
    for (int e = 0; e < 8; ++e) { dwacc[j][e] = 0.f; wv[j][e] = 0.f; }
    if (i < nvec) unpack8(ld_bf16x8(w + (size_t)i * 8), wv[j]);
  }
  for (int row0 = blockIdx.x * kRows; row0 < rows; row0 += gridDim.x * kRows) {
    bf16x8 dyr[kRows][kMaxVec], xr[kRows][kMaxVec];
    float rs[kRows], dot[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int row = row0 + r;
      rs[r] = row < rows ? rstd[row] : 0.f;
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        const int i = threadIdx.x + j * kThr;
        if (row < rows && i < nvec) {
          dyr[r][j] = ld_stream_bf16x8(dy + (size_t)row * cols + (size_t)i * 8);
          xr[r][j] = ld_stream_bf16x8(x + (size_t)row * cols + (size_t)i * 8);
        } else {
          dyr[r][j] = bf16x8{};
          xr[r][j] = bf16x8{};
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        float dyv[8], xv[8];
        unpack8(dyr[r][j], dyv);
        unpack8(xr[r][j], xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = xv[e] * rs[r];
          d += dyv[e] * wv[j][e] * xh;
          dwacc[j][e] += dyv[e] * xh;
        }
      }
      d = warp_sum(d);
      if (lane == 0) red[r][wid] = d;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float v = lane < kWarps ? red[r][lane] : 0.f;
      dot[r] = warp_sum(v) / (float)cols;
    }
    __syncthreads();                                   // `red` is rewritten by the next iteration
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int row = row0 + r;
      if (row >= rows) break;
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        const int i = threadIdx.x + j * kThr;
        if (i < nvec) {
          float dyv[8], xv[8], o[8];
          unpack8(dyr[r][j], dyv);
          unpack8(xr[r][j], xv);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = rs[r] * (dyv[e] * wv[j][e] - xv[e] * rs[r] * dot[r]);
          if (dres_in != nullptr) {
            float rr[8];
            unpack8(ld_stream_bf16x8(dres_in + (size_t)row * cols + (size_t)i * 8), rr);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += rr[e];
          }
          st_bf16x8(dx + (size_t)row * cols + (size_t)i * 8, pack8(o));
        }
      }
    }
  }
  float* out = dw_partial + (size_t)blockIdx.x * cols;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kThr;
    if (i < nvec) {
      *reinterpret_cast<float4*>(out + (size_t)i * 8) =
          make_float4(dwacc[j][0], dwacc[j][1], dwacc[j][2], dwacc[j][3]);
      *reinterpret_cast<float4*>(out + (size_t)i * 8 + 4) =
          make_float4(dwacc[j][4], dwacc[j][5], dwacc[j][6], dwacc[j][7]);
    }
  }
}

template <int kVec, bool kRes>
__global__ void __launch_bounds__(kNormThr, 2) rmsnorm_bwd_stream_kernel(
    const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
    const float* __restrict__ rstd, const bf16* __restrict__ dres_in, bf16* __restrict__ dx,
    float* __restrict__ dw_partial, int rows, int cols) {
  using namespace sm100;
  extern __shared__ __align__(128) uint8_t norm_smem[];
  __shared__ float red[2][kNormThr / 32];
  __shared__ __align__(8) uint64_t bars[kNormStages];
  const uint32_t row_bytes = (uint32_t)cols * 2;
  const uint32_t stage_bytes = row_bytes * (kRes ? 3 : 2);
  const uint32_t smem0 = smem_u32(norm_smem);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nvec = cols >> 3;
  const int my_rows = (rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int k) {
    const int st = k % kNormStages;
    const size_t off = ((size_t)blockIdx.x + (size_t)k * gridDim.x) * cols;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_arrive_expect_tx(bar, stage_bytes);
    bulk_load_1d(smem0 + st * stage_bytes, dy + off, row_bytes, bar);
    bulk_load_1d(smem0 + st * stage_bytes + row_bytes, x + off, row_bytes, bar);
    if (kRes) bulk_load_1d(smem0 + st * stage_bytes + 2 * row_bytes, dres_in + off, row_bytes, bar);
  };
  if (tid == 0) {
    for (int st = 0; st < kNormStages; ++st) mbar_init(smem_u32(&bars[st]), 1);
    fence_barrier_init();
    for (int k = 0; k < kNormStages && k < my_rows; ++k) issue(k);
  }
  float wv[kVec][8], dwacc[kVec][8];
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const int i = tid + j * kNormThr;
#pragma unroll
    for (int e = 0; e < 8; ++e) { dwacc[j][e] = 0.f; wv[j][e] = 0.f; }
    if (i < nvec) unpack8(ld_bf16x8(w + (size_t)i * 8), wv[j]);
  }
  __syncthreads();
  for (int k = 0; k < my_rows; ++k) {
    const int st = k % kNormStages;
    const size_t row = (size_t)blockIdx.x + (size_t)k * gridDim.x;
    const size_t base = row * cols;
    const float rs = rstd[row];
    mbar_wait(smem_u32(&bars[st]), (uint32_t)(k / kNormStages) & 1u);
    float g[kVec][8], xh[kVec][8], rr[kVec][8];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const int i = tid + j * kNormThr;
      if (i < nvec) {
        float dyv[8], xv[8];
        unpack8(lds_bf16x8(smem0 + st * stage_bytes + (uint32_t)i * 16), dyv);
        unpack8(lds_bf16x8(smem0 + st * stage_bytes + row_bytes + (uint32_t)i * 16), xv);
        if (kRes) unpack8(lds_bf16x8(smem0 + st * stage_bytes + 2 * row_bytes + (uint32_t)i * 16), rr[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          xh[j][e] = xv[e] * rs;
          g[j][e] = dyv[e] * wv[j][e];
          dot += g[j][e] * xh[j][e];
          dwacc[j][e] += dyv[e] * xh[j][e];
        }
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) red[k & 1][wid] = dot;
    __syncthreads();                             // every thread has read stage st: refill it
    if (tid == 0 && k + kNormStages < my_rows) issue(k + kNormStages);
    float tot = lane < kNormThr / 32 ? red[k & 1][lane] : 0.f;
    const float c = warp_sum(tot) / (float)cols;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const int i = tid + j * kNormThr;
      if (i < nvec) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = rs * (g[j][e] - xh[j][e] * c);
          if (kRes) o[e] += rr[j][e];
        }
        st_bf16x8(dx + base + (size_t)i * 8, pack8(o));
      }
    }
  }
  float* out = dw_partial + (size_t)blockIdx.x * cols;
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const int i = tid + j * kNormThr;
    if (i < nvec) {
      *reinterpret_cast<float4*>(out + (size_t)i * 8) = make_float4(dwacc[j][0], dwacc[j][1], dwacc[j][2], dwacc[j][3]);
      *reinterpret_cast<float4*>(out + (size_t)i * 8 + 4) = make_float4(dwacc[j][4], dwacc[j][5], dwacc[j][6], dwacc[j][7]);
    }
  }
}

int rmsnorm_bwd_num_partials(int rows) {
  const int cap = kNumSMs * 2;    // both backward variants launch at most this many blocks (one partial row each)
  return rows < cap ? (rows > 0 ? rows : 1) : cap;
}

void rmsnorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* rstd,
                 const bf16* dres_in, bf16* dx, float* dw_partial, int rows, int cols,
                 cudaStream_t s) {
  const int grid = rmsnorm_bwd_num_partials(rows);
  const int nvec = cols / 8;
  if (norm_stream_ok(cols, dy, x) && (reinterpret_cast<uintptr_t>(dres_in) & 15) == 0 && nvec <= kNormThr * 2) {
    const int kv = (nvec + kNormThr - 1) / kNormThr;
    const size_t smem = (size_t)kNormStages * cols * 2 * (dres_in != nullptr ? 3 : 2);
#define HB_BWD_STREAM(KV, RES)                                                                                  \
    do {                                                                                                        \
      static bool attr = false;                                                                                 \
      if (!attr) { cudaFuncSetAttribute(rmsnorm_bwd_stream_kernel<KV, RES>,                                     \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192 * 2 * 3); attr = true; } \
      rmsnorm_bwd_stream_kernel<KV, RES><<<grid, kNormThr, smem, s>>>(dy, x, w, rstd, dres_in, dx, dw_partial, rows, cols); \
    } while (0)
    if (dres_in != nullptr) { if (kv == 1) HB_BWD_STREAM(1, true); else HB_BWD_STREAM(2, true); }
    else { if (kv == 1) HB_BWD_STREAM(1, false); else HB_BWD_STREAM(2, false); }
#undef HB_BWD_STREAM
    count_launch();
    return;
  }
  if (nvec <= 256) {
    rmsnorm_bwd_kernel<256, 1, 4><<<grid, 256, 0, s>>>(dy, x, w, rstd, dres_in, dx, dw_partial, rows, cols);
  } else if (nvec <= 512) {
    rmsnorm_bwd_kernel<512, 1, 4><<<grid, 512, 0, s>>>(dy, x, w, rstd, dres_in, dx, dw_partial, rows, cols);
  } else if (nvec <= 1024) {
    rmsnorm_bwd_kernel<512, 2, 2><<<grid, 512, 0, s>>>(dy, x, w, rstd, dres_in, dx, dw_partial, rows, cols);
  } else {
    rmsnorm_bwd_kernel<512, 4, 1><<<grid, 512, 0, s>>>(dy, x, w, rstd, dres_in, dx, dw_partial, rows, cols);
  }
  count_launch();
}

// out[c] (+)= sum_p partial[p, c].  One block = 32 columns x 8 warps: every warp reads whole 128-byte lines of
// every 8th partial row, the 8 warp sums meet in shared memory (round 1: cols/256 blocks -> 16 blocks for a
// 4096-wide norm took as long as the backward kernel's own tail).
__global__ void __launch_bounds__(256) colsum_partials_kernel(const float* __restrict__ partial,
                                                              int nparts, int cols,
                                                              bf16* __restrict__ out,
                                                              int accumulate) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < cols)
    for (int p = wid; p < nparts; p += 8) acc += partial[(size_t)p * cols + c];
  part[wid][lane] = acc;
  __syncthreads();
  if (wid == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    if (accumulate) t += __bfloat162float(out[c]);
    out[c] = __float2bfloat16_rn(t);
  }
}
void colsum_partials(const float* partial, int nparts, int cols, bf16* out, int accumulate,
                     cudaStream_t s) {
  colsum_partials_kernel<<<(cols + 31) / 32, 256, 0, s>>>(partial, nparts, cols, out, accumulate);
  count_launch();
}

// =============================================================================
// RoPE (HF "rotate_half" convention), in place on the fused QKV activation.
// qkv row layout: [q heads | k heads | v heads] x head_dim; the first
// n_rot_heads heads (q and k) are rotated, v is untouched.
// =============================================================================
__global__ void __launch_bounds__(256) rope_kernel(bf16* __restrict__ qkv,
                                                   const float* __restrict__ cos_t,
                                                   const float* __restrict__ sin_t, int tokens,
                                                   int seq_len, int n_rot_heads, int head_dim,
                                                   int row_stride, float sign) {
  const int half = head_dim >> 1;
  const int vec_per_head = half >> 3;  // 8-element vectors in one half
  const int64_t total = (int64_t)tokens * n_rot_heads * vec_per_head;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vec_per_head);
    const int64_t th = idx / vec_per_head;
    const int head = (int)(th % n_rot_heads);
    const int64_t tok = th / n_rot_heads;
    const int pos = (int)(tok % seq_len);
    bf16* p1 = qkv + tok * row_stride + (int64_t)head * head_dim + v * 8;
    bf16* p2 = p1 + half;
    float a[8], b[8], c[8], sn[8], o1[8], o2[8];
    unpack8(ld_bf16x8(p1), a);
    unpack8(ld_bf16x8(p2), b);
    const float* cp = cos_t + (size_t)pos * half + v * 8;
    const float* sp = sin_t + (size_t)pos * half + v * 8;
    *reinterpret_cast<float4*>(c) = *reinterpret_cast<const float4*>(cp);
    *reinterpret_cast<float4*>(c + 4) = *reinterpret_cast<const float4*>(cp + 4);
    *reinterpret_cast<float4*>(sn) = *reinterpret_cast<const float4*>(sp);
    *reinterpret_cast<float4*>(sn + 4) = *reinterpret_cast<const float4*>(sp + 4);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float s_ = sn[e] * sign;
      o1[e] = a[e] * c[e] - b[e] * s_;
      o2[e] = b[e] * c[e] + a[e] * s_;
    }
    st_bf16x8(p1, pack8(o1));
    st_bf16x8(p2, pack8(o2));
  }
}
void rope_inplace(bf16* qkv, const float* cos, const float* sin, int tokens, int seq_len,
                  int n_rot_heads, int head_dim, int row_stride, int backward, cudaStream_t s) {
  const int64_t total = (int64_t)tokens * n_rot_heads * (head_dim / 16);
  if (total == 0) return;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  rope_kernel<<<grid, 256, 0, s>>>(qkv, cos, sin, tokens, seq_len, n_rot_heads, head_dim,
                                   row_stride, backward ? -1.f : 1.f);
  count_launch();
}

// =============================================================================
// SwiGLU on the fused gate|up activation: out = silu(g) * u
// =============================================================================
// gate/up are INTERLEAVED along the row: gu[r, 2i] = gate_i, gu[r, 2i+1] = up_i (the rows of the fused gate_up
// weight are interleaved the same way), so that one 256-wide GEMM output tile holds both halves of 128 pairs and
// the GEMM epilogue can emit silu(gate) * up itself (gemm_tcgen05.cu, `swiglu` epilogue).
__device__ __forceinline__ float silu_mul(float g, float u) {
  const float sg = __fdividef(1.f, 1.f + __expf(-g));
  return g * sg * u;
}
// One thread = one 16-byte chunk of gu (4 (gate, up) pairs) and the matching 8 bytes of the [rows, inter] side:
// consecutive lanes touch consecutive chunks, so every load / store instruction of a warp covers whole 32-byte
// sectors (a 32-byte-per-thread split made each instruction use half of every sector it touched: 611 us instead of
// 180 us for the backward at 8192 x 14336).
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const bf16* __restrict__ gu,
                                                         bf16* __restrict__ out, int rows,
                                                         int inter) {
  const int chunks = inter >> 2;                       // 16-byte chunks per gu row
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= chunks) return;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    float a[8];
    unpack8(ld_stream_bf16x8(gu + (int64_t)r * 2 * inter + (int64_t)c * 8), a);      // g0 u0 g1 u1 g2 u2 g3 u3
    __nv_bfloat162 o[2];
    o[0] = __floats2bfloat162_rn(silu_mul(a[0], a[1]), silu_mul(a[2], a[3]));
    o[1] = __floats2bfloat162_rn(silu_mul(a[4], a[5]), silu_mul(a[6], a[7]));
    *reinterpret_cast<uint2*>(out + (int64_t)r * inter + (int64_t)c * 4) = *reinterpret_cast<uint2*>(o);
  }
}
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const bf16* __restrict__ dout,
                                                         const bf16* __restrict__ gu,
                                                         bf16* __restrict__ dgu, int rows,
                                                         int inter) {
  const int chunks = inter >> 2;
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= chunks) return;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    float a[8], o[8];
    unpack8(ld_stream_bf16x8(gu + (int64_t)r * 2 * inter + (int64_t)c * 8), a);
    const uint2 dv = *reinterpret_cast<const uint2*>(dout + (int64_t)r * inter + (int64_t)c * 4);
    const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dv);
    const float2 d01 = __bfloat1622float2(d2[0]), d23 = __bfloat1622float2(d2[1]);
    const float d[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float g = a[2 * e], u = a[2 * e + 1];
      const float sg = __fdividef(1.f, 1.f + __expf(-g));
      const float silu = g * sg;
      o[2 * e] = d[e] * u * (sg + silu * (1.f - sg));
      o[2 * e + 1] = d[e] * silu;
    }
    st_bf16x8(dgu + (int64_t)r * 2 * inter + (int64_t)c * 8, pack8(o));
  }
}
static int ew_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}
static dim3 swiglu_grid(int rows, int inter) {
  const int gx = (inter / 4 + 255) / 256;
  int gy = rows < 8192 ? rows : 8192;
  if (gy < 1) gy = 1;
  return dim3(gx, gy);
}
void swiglu_fwd(const bf16* gu, bf16* out, int rows, int inter, cudaStream_t s) {
  swiglu_fwd_kernel<<<swiglu_grid(rows, inter), 256, 0, s>>>(gu, out, rows, inter);
  count_launch();
}
void swiglu_bwd(const bf16* dout, const bf16* gu, bf16* dgu, int rows, int inter,
                cudaStream_t s) {
  swiglu_bwd_kernel<<<swiglu_grid(rows, inter), 256, 0, s>>>(dout, gu, dgu, rows, inter);
  count_launch();
}

// =============================================================================
// Fused softmax cross-entropy forward + backward, in place:
//   loss[r] = logsumexp(z_r) - z_r[target]; z_r <- (softmax(z_r) - onehot) * scale
// One CTA per row; pass 1 is an online (max,sum) reduction, pass 2 rewrites the
// row (second read hits L2: a 128k-vocab row is 250 KB).
// =============================================================================
__global__ void __launch_bounds__(512) cross_entropy_kernel(bf16* __restrict__ logits,
                                                            const int64_t* __restrict__ targets,
                                                            float* __restrict__ loss_rows,
                                                            int vocab, int row_stride,
                                                            float grad_scale,
                                                            int64_t ignore_index) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  bf16* z = logits + (size_t)row * row_stride;
  const int64_t tgt = targets[row];
  const int nvec = vocab >> 3;
  float m = -INFINITY, ssum = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float f[8];
    unpack8(ld_bf16x8(z + (size_t)i * 8), f);
    float lm = f[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) lm = fmaxf(lm, f[e]);
    const float nm = fmaxf(m, lm);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += __expf(f[e] - nm);
    ssum = ssum * __expf(m - nm) + acc;
    m = nm;
  }
  const float gm = block_max(m, red);
  ssum = (m == -INFINITY) ? 0.f : ssum * __expf(m - gm);
  const float gs = block_sum(ssum, red);
  const bool ignored = (tgt == ignore_index) || tgt < 0 || tgt >= vocab;
  if (threadIdx.x == 0) {
    float l = 0.f;
    if (!ignored) l = __logf(gs) + gm - __bfloat162float(z[tgt]);
    loss_rows[row] = l;
  }
  __syncthreads();  // z[tgt] read above must precede the in-place rewrite
  const float inv = 1.f / gs;
  const float sc = ignored ? 0.f : grad_scale;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float f[8];
    unpack8(ld_bf16x8(z + (size_t)i * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float p = __expf(f[e] - gm) * inv;
      if ((int64_t)i * 8 + e == tgt) p -= 1.f;
      f[e] = p * sc;
    }
    st_bf16x8(z + (size_t)i * 8, pack8(f));
  }
}
void cross_entropy_fwd_bwd(bf16* logits, const int64_t* targets, float* loss_rows, int rows,
                           int vocab, int row_stride, float grad_scale, int64_t ignore_index,
                           cudaStream_t s) {
  if (rows == 0) return;
  cross_entropy_kernel<<<rows, 512, 0, s>>>(logits, targets, loss_rows, vocab, row_stride,
                                            grad_scale, ignore_index);
  count_launch();
}

// =============================================================================
// Embedding gather / scatter-add (bf16x2 red.global atomics into the bf16
// gradient table; the table region must be zeroed by the caller beforehand).
// =============================================================================
__global__ void __launch_bounds__(256) embedding_fwd_kernel(const int64_t* __restrict__ ids,
                                                            const bf16* __restrict__ table,
                                                            bf16* __restrict__ out, int tokens,
                                                            int dim) {
  const int vec_per_row = dim >> 3;
  const int64_t total = (int64_t)tokens * vec_per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / vec_per_row;
    const int v = (int)(idx % vec_per_row);
    st_bf16x8(out + t * dim + (int64_t)v * 8, ld_bf16x8(table + ids[t] * dim + (int64_t)v * 8));
  }
}
__global__ void __launch_bounds__(256) embedding_bwd_kernel(const int64_t* __restrict__ ids,
                                                            const bf16* __restrict__ dout,
                                                            bf16* __restrict__ dtable,
                                                            int tokens, int dim) {
  const int vec_per_row = dim >> 3;
  const int64_t total = (int64_t)tokens * vec_per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / vec_per_row;
    const int v = (int)(idx % vec_per_row);
    bf16x8 d = ld_stream_bf16x8(dout + t * dim + (int64_t)v * 8);
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(dtable + ids[t] * dim + (int64_t)v * 8);
#pragma unroll
    for (int e = 0; e < 4; ++e) atomicAdd(dst + e, d.v[e]);
  }
}
void embedding_fwd(const int64_t* ids, const bf16* table, bf16* out, int tokens, int dim,
                   cudaStream_t s) {
  embedding_fwd_kernel<<<ew_grid((int64_t)tokens * (dim / 8)), 256, 0, s>>>(ids, table, out, tokens, dim);
  count_launch();
}
void embedding_bwd(const int64_t* ids, const bf16* dout, bf16* dtable, int tokens, int dim,
                   cudaStream_t s) {
  embedding_bwd_kernel<<<ew_grid((int64_t)tokens * (dim / 8)), 256, 0, s>>>(ids, dout, dtable, tokens, dim);
  count_launch();
}

__global__ void __launch_bounds__(256) add_f32_into_bf16_kernel(const float* __restrict__ src,
                                                                bf16* __restrict__ dst,
                                                                int64_t n, int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = src[i];
    if (accumulate) v += __bfloat162float(dst[i]);
    dst[i] = __float2bfloat16_rn(v);
  }
}
void add_f32_into_bf16(const float* src, bf16* dst, int64_t n, int accumulate, cudaStream_t s) {
  add_f32_into_bf16_kernel<<<ew_grid(n), 256, 0, s>>>(src, dst, n, accumulate);
  count_launch();
}

}  // namespace hb
