// fp8 quantisation for the tcgen05 kind::f8f6f4 GEMM path (BASELINE config 3: fp8 inner steps).
// Per-tensor *current* scaling: amax(x) is reduced first, then x is scaled by fmax/amax and
// converted with saturation; the dequantisation factor amax/fmax is left on the device for the
// GEMM epilogue, so no host synchronisation is involved.
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.h"

namespace hb {

__global__ void __launch_bounds__(256) amax_bf16_kernel(const bf16* __restrict__ x, int64_t nvec,
                                                        float* __restrict__ amax) {
  __shared__ float red[32];
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(ld_bf16x8(x + (i << 3)), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) m = fmaxf(m, fabsf(f[e]));
  }
  m = block_max(m, red);
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));  // m >= 0
}

// out = fp8(x * fmax/amax); dequant[0] = amax/fmax.  fmt: 0 = e4m3 (fmax 448), 1 = e5m2 (fmax 57344).
template <int kFmt>
__global__ void __launch_bounds__(256) quant_fp8_kernel(const bf16* __restrict__ x,
                                                        uint8_t* __restrict__ out, int64_t nvec,
                                                        const float* __restrict__ amax,
                                                        float* __restrict__ dequant) {
  constexpr float kMax = kFmt == 0 ? 448.f : 57344.f;
  const float am = fmaxf(*amax, 1e-12f);
  const float scale = kMax / am;
  if (blockIdx.x == 0 && threadIdx.x == 0) *dequant = am / kMax;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(ld_stream_bf16x8(x + (i << 3)), f);
    uint32_t w[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t acc = 0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float2 v = make_float2(f[4 * h + 2 * q] * scale, f[4 * h + 2 * q + 1] * scale);
        const __nv_fp8x2_storage_t p2 =
            __nv_cvt_float2_to_fp8x2(v, __NV_SATFINITE, kFmt == 0 ? __NV_E4M3 : __NV_E5M2);
        acc |= (uint32_t)p2 << (16 * q);
      }
      w[h] = acc;
    }
    *reinterpret_cast<uint2*>(out + (i << 3)) = make_uint2(w[0], w[1]);
  }
}

static int q_grid(int64_t nvec) {
  int64_t b = (nvec + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

void amax_bf16(const bf16* x, int64_t n, float* amax, cudaStream_t s) {
  cudaMemsetAsync(amax, 0, sizeof(float), s);
  amax_bf16_kernel<<<q_grid(n / 8), 256, 0, s>>>(x, n / 8, amax);
  count_launch();
}

void quantize_fp8(const bf16* x, uint8_t* out, int64_t n, const float* amax, float* dequant, int fmt,
                  cudaStream_t s) {
  if (fmt == 0) quant_fp8_kernel<0><<<q_grid(n / 8), 256, 0, s>>>(x, out, n / 8, amax, dequant);
  else quant_fp8_kernel<1><<<q_grid(n / 8), 256, 0, s>>>(x, out, n / 8, amax, dequant);
  count_launch();
}

// ---------------------------------------------------------------------------------------------------------------
// MXFP8 (OCP microscaling): e4m3 values with one power-of-two (UE8M0) scale per 32 consecutive elements along the
// contraction dimension, consumed by tcgen05.mma.kind::mxf8f6f4.block_scale.
//
// One pass over a bf16 matrix x[R, C] produces up to two operands:
//   q  [R, C]  scaled along C (the operand when the GEMM contracts over C), scale table sf
//   qT [C, R]  = x^T scaled along R (the operand when the GEMM contracts over R), scale table sfT
// so every GEMM of a Linear layer (fwd: x, W; dgrad: dy, W^T; wgrad: dy^T, x^T) reads K-major operands.
// A scale table is laid out the way tcgen05.cp wants to find it in shared memory -- per (128-element K block, 128-row
// block) one 512-byte atom with byte (r % 32) * 16 + (r / 32) * 4 + (k / 32), K blocks outermost -- so the GEMM
// fetches one contiguous box per operand and stage (the two row blocks of a 256-wide tile are adjacent) and needs no
// shuffle.
//
// Scale of a block: 2^e with e = ceil(log2(amax / 448)), i.e. the smallest power of two that keeps the block inside
// the e4m3 range without saturating; stored biased (e + 127), amax = 0 stores 0.
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kQT = 128;                     // tile edge
constexpr int kQPitch = kQT * 2 + 16;        // bytes per staged bf16 row (+16: column reads hit 32 distinct banks)

__device__ __forceinline__ uint32_t mx_scale_byte(float amax) {
  const uint32_t u = __float_as_uint(amax * (1.f / 448.f));
  uint32_t e = (u >> 23) & 0xffu;
  if (u & 0x7fffffu) e += 1u;
  return e > 254u ? 254u : e;                 // amax == 0 -> 0 (the values are 0 too)
}
__device__ __forceinline__ float mx_inv_scale(uint32_t e) {   // 2^(127 - e)
  return __uint_as_float((254u - e) << 23);
}
__device__ __forceinline__ uint32_t cvt4_e4m3(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}

template <bool kRow, bool kCol>
__global__ void __launch_bounds__(256) quant_mxfp8_kernel(const bf16* __restrict__ x, int ldx, uint8_t* __restrict__ q,
                                                          int ldq, uint8_t* __restrict__ sf, uint8_t* __restrict__ qT,
                                                          int ldqT, uint8_t* __restrict__ sfT, int R, int C) {
  __shared__ __align__(16) uint8_t tile[kCol ? kQT * kQPitch : 16];
  __shared__ __align__(16) uint8_t sf_row[512];
  __shared__ __align__(16) uint8_t sf_col[512];
  const int tc = blockIdx.x, tr = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cl = (lane & 15) * 8;                              // first of this lane's 8 columns inside the tile
#pragma unroll 4
  for (int it = 0; it < 8; ++it) {
    const int rl = warp * 16 + it * 2 + (lane >> 4);           // row inside the tile
    const bf16x8 v = ld_stream_bf16x8(x + (size_t)(tr * kQT + rl) * ldx + tc * kQT + cl);
    if constexpr (kCol) *reinterpret_cast<bf16x8*>(tile + rl * kQPitch + cl * 2) = v;
    if constexpr (kRow) {
      float f[8];
      unpack8(v, f);
      float m = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) m = fmaxf(m, fabsf(f[e]));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));        // 4 lanes = one 32-element block
      const uint32_t e = mx_scale_byte(m);
      const float inv = mx_inv_scale(e);
      const uint2 o = make_uint2(cvt4_e4m3(f[0] * inv, f[1] * inv, f[2] * inv, f[3] * inv),
                                 cvt4_e4m3(f[4] * inv, f[5] * inv, f[6] * inv, f[7] * inv));
      *reinterpret_cast<uint2*>(q + (size_t)(tr * kQT + rl) * ldq + tc * kQT + cl) = o;
      if ((lane & 3) == 0) sf_row[(rl & 31) * 16 + (rl >> 5) * 4 + ((lane & 15) >> 2)] = (uint8_t)e;
    }
  }
  __syncthreads();
  if constexpr (kCol) {
    // thread = (column pair, 32-row block): one 32-bit shared-memory word per row, conflict-free across the warp
    const int cp = threadIdx.x & 63, rb = threadIdx.x >> 6;
    uint32_t w[32];
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      w[r] = *reinterpret_cast<const uint32_t*>(tile + (rb * 32 + r) * kQPitch + cp * 4);
      const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[r]));
      m0 = fmaxf(m0, fabsf(t.x));
      m1 = fmaxf(m1, fabsf(t.y));
    }
    const uint32_t e0 = mx_scale_byte(m0), e1 = mx_scale_byte(m1);
    const float i0 = mx_inv_scale(e0), i1 = mx_inv_scale(e1);
    uint32_t o0[8], o1[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float a[4], b[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[g * 4 + e]));
        a[e] = t.x * i0;
        b[e] = t.y * i1;
      }
      o0[g] = cvt4_e4m3(a[0], a[1], a[2], a[3]);
      o1[g] = cvt4_e4m3(b[0], b[1], b[2], b[3]);
    }
    // Transposed tile through shared memory (the bf16 staging tile is dead once every thread holds its words), so
    // that a warp stores four 128-byte rows of qT per instruction instead of 32 scattered 16-byte pieces (the direct
    // stores ran at 0.46 of the HBM rate).  Row pitch 144 B keeps the 16-byte column writes at 2-way conflicts.
    __syncthreads();
    constexpr int kTP = kQT + 16;
    {
      uint8_t* r0p = tile + (2 * cp) * kTP + rb * 32;
      uint8_t* r1p = r0p + kTP;
      *reinterpret_cast<uint4*>(r0p) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
      *reinterpret_cast<uint4*>(r0p + 16) = make_uint4(o0[4], o0[5], o0[6], o0[7]);
      *reinterpret_cast<uint4*>(r1p) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
      *reinterpret_cast<uint4*>(r1p + 16) = make_uint4(o1[4], o1[5], o1[6], o1[7]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = threadIdx.x + 256 * i, row = j >> 3, part = j & 7;
      const uint4 vq = *reinterpret_cast<const uint4*>(tile + row * kTP + part * 16);
      *reinterpret_cast<uint4*>(qT + (size_t)(tc * kQT + row) * ldqT + tr * kQT + part * 16) = vq;
    }
    const int c0 = 2 * cp, c1 = 2 * cp + 1;
    sf_col[(c0 & 31) * 16 + (c0 >> 5) * 4 + rb] = (uint8_t)e0;
    sf_col[(c1 & 31) * 16 + (c1 >> 5) * 4 + rb] = (uint8_t)e1;
    __syncthreads();
  }
  // tables are [K block][128-row block][512]: for q the K blocks run along C (tc), for qT along R (tr)
  if (kRow && threadIdx.x < 32)
    reinterpret_cast<uint4*>(sf + ((size_t)tc * gridDim.y + tr) * 512)[threadIdx.x] =
        reinterpret_cast<const uint4*>(sf_row)[threadIdx.x];
  if (kCol && threadIdx.x >= 32 && threadIdx.x < 64)
    reinterpret_cast<uint4*>(sfT + ((size_t)tr * gridDim.x + tc) * 512)[threadIdx.x - 32] =
        reinterpret_cast<const uint4*>(sf_col)[threadIdx.x - 32];
}

}  // namespace

// x[R, C] bf16 (row pitch ldx) -> q[R, C] + sf and / or qT[C, R] + sfT (either pair may be null); R, C % 128 == 0.
int quantize_mxfp8(const bf16* x, int ldx, uint8_t* q, int ldq, uint8_t* sf, uint8_t* qT, int ldqT, uint8_t* sfT, int R,
                   int C, cudaStream_t s) {
  if (R <= 0 || C <= 0 || (R % kQT) || (C % kQT) || (ldx % 8)) return 1;
  const bool row = q != nullptr, col = qT != nullptr;
  if ((row && ((ldq % 16) || sf == nullptr)) || (col && ((ldqT % 16) || sfT == nullptr)) || (!row && !col)) return 2;
  const dim3 grid(C / kQT, R / kQT);
  if (row && col) quant_mxfp8_kernel<true, true><<<grid, 256, 0, s>>>(x, ldx, q, ldq, sf, qT, ldqT, sfT, R, C);
  else if (row) quant_mxfp8_kernel<true, false><<<grid, 256, 0, s>>>(x, ldx, q, ldq, sf, qT, ldqT, sfT, R, C);
  else quant_mxfp8_kernel<false, true><<<grid, 256, 0, s>>>(x, ldx, q, ldq, sf, qT, ldqT, sfT, R, C);
  count_launch();
  return 0;
}

}  // namespace hb
