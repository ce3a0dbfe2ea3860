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

}  // namespace hb
