// Memory-bound transformer ops for sm_100a: fused (residual+)RMSNorm fwd/bwd,
// RoPE, SwiGLU, fused softmax-cross-entropy fwd+bwd, embedding gather/scatter.
// All I/O is bf16 with 16-byte vector accesses; math is fp32.
//
// These replace the stock ATen ops the reference reaches through
// `model(**batch)` / `accelerator.backward(loss)`
// (reference: executors/accelerate/src/hypha/accelerate_executor/training.py:108-112).
#include "common.cuh"
#include "kernels.h"

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

void rmsnorm_fwd(const bf16* x, const bf16* res_in, bf16* res_out, const bf16* w, bf16* y,
                 float* rstd, int rows, int cols, float eps, cudaStream_t s) {
  if (rows == 0) return;
  const int nvec = cols / 8;
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

int rmsnorm_bwd_num_partials(int rows) {
  const int cap = kNumSMs;        // the kernel holds kRows rows in registers: one 512-thread block per SM
  return rows < cap ? (rows > 0 ? rows : 1) : cap;
}

void rmsnorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* rstd,
                 const bf16* dres_in, bf16* dx, float* dw_partial, int rows, int cols,
                 cudaStream_t s) {
  const int grid = rmsnorm_bwd_num_partials(rows);
  const int nvec = cols / 8;
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
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const bf16* __restrict__ gu,
                                                         bf16* __restrict__ out, int rows,
                                                         int inter) {
  // grid.y strides over rows, grid.x over 8-element vectors of a row: no per-element div/mod
  const int vec_per_row = inter >> 3;
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v >= vec_per_row) return;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const bf16* gp = gu + (int64_t)r * 2 * inter + (int64_t)v * 8;
    float g[8], u[8], o[8];
    unpack8(ld_stream_bf16x8(gp), g);
    unpack8(ld_stream_bf16x8(gp + inter), u);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = __fdividef(1.f, 1.f + __expf(-g[e]));
      o[e] = g[e] * sg * u[e];
    }
    st_bf16x8(out + (int64_t)r * inter + (int64_t)v * 8, pack8(o));
  }
}
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const bf16* __restrict__ dout,
                                                         const bf16* __restrict__ gu,
                                                         bf16* __restrict__ dgu, int rows,
                                                         int inter) {
  const int vec_per_row = inter >> 3;
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v >= vec_per_row) return;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const bf16* gp = gu + (int64_t)r * 2 * inter + (int64_t)v * 8;
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(ld_stream_bf16x8(gp), g);
    unpack8(ld_stream_bf16x8(gp + inter), u);
    unpack8(ld_stream_bf16x8(dout + (int64_t)r * inter + (int64_t)v * 8), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = __fdividef(1.f, 1.f + __expf(-g[e]));
      const float silu = g[e] * sg;
      du[e] = d[e] * silu;
      dg[e] = d[e] * u[e] * (sg + silu * (1.f - sg));
    }
    bf16* op = dgu + (int64_t)r * 2 * inter + (int64_t)v * 8;
    st_bf16x8(op, pack8(dg));
    st_bf16x8(op + inter, pack8(du));
  }
}
static int ew_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}
static dim3 swiglu_grid(int rows, int inter) {
  const int gx = (inter / 8 + 255) / 256;
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
