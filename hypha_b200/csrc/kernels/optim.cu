// Fused flat ("multi-tensor by construction") AdamW for sm_100a.
//
// All parameters of a model live in ONE flat bf16 buffer with matching flat
// gradient / fp32-master / moment buffers, so a single launch updates the whole
// model: 1 read of g, 1 RMW of master/m/v, 1 write of the bf16 compute copy.
// Moments may be stored in bf16 with stochastic rounding (halves optimizer
// memory so Llama-3-8B + outer DiLoCo state fits one 192 GB B200).
//
// Semantics match torch.optim.AdamW (decoupled weight decay), which is what the
// reference instantiates (executors/accelerate/.../utils.py:56-64).
#include "common.cuh"
#include "kernels.h"

namespace hb {

template <bool kStateBf16>
__global__ void __launch_bounds__(256) adamw_flat_kernel(
    bf16* __restrict__ p, const bf16* __restrict__ g, float* __restrict__ master,
    void* __restrict__ m_, void* __restrict__ v_, int64_t nvec, float lr, float beta1,
    float beta2, float eps, float decay, float inv_bc1, float inv_sqrt_bc2, float grad_scale,
    uint32_t seed) {
  const float step_size = lr * inv_bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = i * 8;
    float gf[8], w[8], m[8], v[8];
    unpack8(ld_stream_bf16x8(g + off), gf);
    *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(master + off);
    *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(master + off + 4);
    if constexpr (kStateBf16) {
      unpack8(ld_bf16x8(reinterpret_cast<const bf16*>(m_) + off), m);
      unpack8(ld_bf16x8(reinterpret_cast<const bf16*>(v_) + off), v);
    } else {
      const float* mp = reinterpret_cast<const float*>(m_) + off;
      const float* vp = reinterpret_cast<const float*>(v_) + off;
      *reinterpret_cast<float4*>(m) = *reinterpret_cast<const float4*>(mp);
      *reinterpret_cast<float4*>(m + 4) = *reinterpret_cast<const float4*>(mp + 4);
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(vp);
      *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(vp + 4);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float ge = gf[e] * grad_scale;
      m[e] = beta1 * m[e] + (1.f - beta1) * ge;
      v[e] = beta2 * v[e] + (1.f - beta2) * ge * ge;
      // 2 MUFU ops / element (sqrt.approx + rcp.approx): the IEEE sqrtf + division made this
      // kernel issue-bound (5.1 ms / 2^30 elements) instead of HBM-bound.
      float sq;
      asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v[e]));
      const float denom = fmaf(sq, inv_sqrt_bc2, eps);
      w[e] = fmaf(w[e], decay, -step_size * __fdividef(m[e], denom));
    }
    *reinterpret_cast<float4*>(master + off) = *reinterpret_cast<float4*>(w);
    *reinterpret_cast<float4*>(master + off + 4) = *reinterpret_cast<float4*>(w + 4);
    st_bf16x8(p + off, pack8(w));
    if constexpr (kStateBf16) {
      bf16x8 mo, vo;
      bf16* mo_e = reinterpret_cast<bf16*>(&mo);
      bf16* vo_e = reinterpret_cast<bf16*>(&vo);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t r = mix32((uint32_t)(off + e) * 0x9E3779B9u ^ seed);
        mo_e[e] = f2bf_sr(m[e], r & 0xffffu);
        vo_e[e] = f2bf_sr(v[e], r >> 16);
      }
      st_bf16x8(reinterpret_cast<bf16*>(m_) + off, mo);
      st_bf16x8(reinterpret_cast<bf16*>(v_) + off, vo);
    } else {
      float* mp = reinterpret_cast<float*>(m_) + off;
      float* vp = reinterpret_cast<float*>(v_) + off;
      *reinterpret_cast<float4*>(mp) = *reinterpret_cast<float4*>(m);
      *reinterpret_cast<float4*>(mp + 4) = *reinterpret_cast<float4*>(m + 4);
      *reinterpret_cast<float4*>(vp) = *reinterpret_cast<float4*>(v);
      *reinterpret_cast<float4*>(vp + 4) = *reinterpret_cast<float4*>(v + 4);
    }
  }
}

void adamw_flat(bf16* p, const bf16* g, float* master, void* m, void* v, int state_bf16,
                int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                float bias_corr1, float bias_corr2, float grad_scale, uint32_t seed,
                cudaStream_t s) {
  const int64_t nvec = n / 8;
  if (nvec == 0) return;
  int64_t blocks = (nvec + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
  const float decay = 1.f - lr * weight_decay;
  const float inv_bc1 = 1.f / bias_corr1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bias_corr2);
  if (state_bf16) {
    adamw_flat_kernel<true><<<grid, 256, 0, s>>>(p, g, master, m, v, nvec, lr, beta1, beta2, eps,
                                                 decay, inv_bc1, inv_sqrt_bc2, grad_scale, seed);
  } else {
    adamw_flat_kernel<false><<<grid, 256, 0, s>>>(p, g, master, m, v, nvec, lr, beta1, beta2, eps,
                                                  decay, inv_bc1, inv_sqrt_bc2, grad_scale, seed);
  }
  count_launch();
}

}  // namespace hb
