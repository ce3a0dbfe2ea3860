// DiLoCo outer step as ONE persistent sm_100a kernel over NVLink/NVSwitch peer
// memory -- no NCCL on this path.
//
// Replaces the reference's parameter-server round trip
//   worker:  delta = theta_local - theta_prev              (utils.py:118-123, on CPU)
//   PS:      avg  = running (a+b)/2 over received files     (parameter_server.rs:182-211)
//            m    = mu*m + avg  (m0 = avg)                  (parameter_server.rs:386-416)
//            u    = lr*(mu*m + avg)                         (parameter_server.rs:418-446)
//   worker:  theta = theta_prev + u                         (utils.py:105-115)
// with a true 1/N mean (the reference's pairwise average is only right for N=2,
// SURVEY appendix B Q1).
//
// Layout.  Every rank r owns shard r of the flat parameter vector and keeps the
// fp32 outer parameters theta_outer[shard] and momentum[shard] (ZeRO-1 style).
// Every rank also keeps `anchor` = the bf16 weights all ranks started the round
// from.  Using  mean_r(master_r) - theta_outer = mean_r(master_r - anchor) +
// (anchor - theta_outer)  only bf16 deltas relative to the common anchor cross
// NVLink while the owner's fp32 theta_outer absorbs the exact update (error
// feedback), so no precision is lost to the 16-bit wire format.
//
// Phases (all inside one cooperative launch per rank; software-pipelined over chunks for large n):
//   A  local : delta = bf16(master - anchor)                 -> symmetric buffer
//   -- cross-GPU flag barrier (st.release.sys / ld.acquire.sys, epoch-stamped) --
//   B  owner : g = mean_r delta_r[shard] (P2P 16-byte loads from every peer)
//              + (anchor - theta_outer); Nesterov; theta_outer += u;
//              P2P-store bf16(theta_outer) into every member's parameter buffer
//   -- cross-GPU flag barrier --
//   C  local : master = float(param); anchor = param
// Barriers carry a timeout: a dead peer yields status != 0 instead of a hang so
// the host can shrink the membership and retry.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace hb {

namespace {

struct LocalSync {
  unsigned int grid_bar;   // local grid-barrier counter (zeroed before launch)
  unsigned int status;     // 0 ok, 1 = barrier-1 timeout, 2 = barrier-2 timeout
  unsigned int missing;    // bitmask of ranks that never arrived
  unsigned int pad;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 ld_peer_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_peer_16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// NVSwitch in-fabric reduction: one 16-byte load returns the SUM over all devices bound to the
// multicast object (8 bf16 lanes, accumulated in fp32 inside the switch).
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}
// One 16-byte store replicated by the switch into every device's copy.
__device__ __forceinline__ void multimem_st_bf16x8(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Local grid barrier over a monotonically increasing counter.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(bar, 1u);
    const unsigned int target = (old / nblocks + 1u) * nblocks;
    while (ld_acquire_gpu(bar) < target) { __nanosleep(32); }
  }
  __syncthreads();
}

// Cross-GPU barrier: CTA 0 / thread r signals every member, then waits for all.
// `slot` separates the barrier instances of one launch; tokens are epoch-unique.
__device__ void cross_gpu_barrier(const OuterStepArgs& a, LocalSync* ls, int slot,
                                  uint32_t token, uint64_t timeout_ns, unsigned int err_code) {
  grid_barrier(&ls->grid_bar, gridDim.x);
  if (blockIdx.x == 0) {
    if (threadIdx.x < 8 && ((a.member_mask >> threadIdx.x) & 1u)) {
      const int peer = threadIdx.x;
      st_release_sys(a.flag_peers[peer] + slot * 8 + a.rank, token);
      const uint32_t* mine = a.flag_peers[a.rank] + slot * 8 + peer;
      const uint64_t t0 = globaltimer_ns();
      while ((int32_t)(ld_acquire_sys(mine) - token) < 0) {
        __nanosleep(64);
        if (globaltimer_ns() - t0 > timeout_ns) {
          atomicExch(&ls->status, err_code);
          atomicOr(&ls->missing, 1u << peer);
          break;
        }
      }
    }
  }
  grid_barrier(&ls->grid_bar, gridDim.x);
}

// ---- Phase A: local pseudo-gradient relative to the common anchor, elements [base+8*lo, base+8*hi)
// (two 16-byte vectors per thread per trip: all loads are issued before the first use)
__device__ __forceinline__ void phase_a(const OuterStepArgs& a, int64_t base, int64_t lo, int64_t hi,
                                        int64_t tid, int64_t nthreads) {
  bf16* my_delta = const_cast<bf16*>(a.delta_peers[a.rank]) + base;
  const bf16* anchor = a.anchor_peers[a.rank] + base;
  const float* master = a.master + base;
  for (int64_t i = lo + tid; i < hi; i += 2 * nthreads) {
    const int64_t i1 = i + nthreads;
    const bool two = i1 < hi;
    const int64_t off0 = i << 3, off1 = (two ? i1 : i) << 3;
    float w0[8], w1[8], x0[8], x1[8];
    *reinterpret_cast<float4*>(w0) = ld_stream_f32x4(master + off0);
    *reinterpret_cast<float4*>(w0 + 4) = ld_stream_f32x4(master + off0 + 4);
    *reinterpret_cast<float4*>(w1) = ld_stream_f32x4(master + off1);
    *reinterpret_cast<float4*>(w1 + 4) = ld_stream_f32x4(master + off1 + 4);
    const bf16x8 a0 = ld_stream_bf16x8(anchor + off0), a1 = ld_stream_bf16x8(anchor + off1);
    unpack8(a0, x0);
    unpack8(a1, x1);
#pragma unroll
    for (int e = 0; e < 8; ++e) { w0[e] -= x0[e]; w1[e] -= x1[e]; }
    st_bf16x8(my_delta + off0, pack8(w0));
    if (two) st_bf16x8(my_delta + off1, pack8(w1));
  }
}

// ---- Phase C: adopt the new global weights locally -------------------------------------------
__device__ __forceinline__ void phase_c(const OuterStepArgs& a, int64_t base, int64_t lo, int64_t hi,
                                        int64_t tid, int64_t nthreads) {
  const bf16* param = a.param_peers[a.rank] + base;
  bf16* anchor = a.anchor_peers[a.rank] + base;
  float* master = a.master + base;
  for (int64_t i = lo + tid; i < hi; i += 2 * nthreads) {
    const int64_t i1 = i + nthreads;
    const bool two = i1 < hi;
    const int64_t off0 = i << 3, off1 = (two ? i1 : i) << 3;
    const uint4 r0 = ld_peer_16(param + off0);   // written by remote owners: bypass L1
    const uint4 r1 = ld_peer_16(param + off1);
    float w[8];
    unpack8(*reinterpret_cast<const bf16x8*>(&r0), w);
    *reinterpret_cast<float4*>(master + off0) = *reinterpret_cast<float4*>(w);
    *reinterpret_cast<float4*>(master + off0 + 4) = *reinterpret_cast<float4*>(w + 4);
    *reinterpret_cast<uint4*>(anchor + off0) = r0;
    if (two) {
      unpack8(*reinterpret_cast<const bf16x8*>(&r1), w);
      *reinterpret_cast<float4*>(master + off1) = *reinterpret_cast<float4*>(w);
      *reinterpret_cast<float4*>(master + off1 + 4) = *reinterpret_cast<float4*>(w + 4);
      *reinterpret_cast<uint4*>(anchor + off1) = r1;
    }
  }
}

// Nesterov on 8 lanes given the summed delta; returns the packed bf16 weights.
__device__ __forceinline__ uint4 nesterov8(const OuterStepArgs& a, const uint4& sum_raw,
                                           const bf16* anchor_at, float* theta_at, float* mom_at) {
  float acc[8], x[8], th[8], mo[8];
  unpack8(*reinterpret_cast<const bf16x8*>(&sum_raw), acc);
  unpack8(ld_stream_bf16x8(anchor_at), x);
  *reinterpret_cast<float4*>(th) = *reinterpret_cast<const float4*>(theta_at);
  *reinterpret_cast<float4*>(th + 4) = *reinterpret_cast<const float4*>(theta_at + 4);
  *reinterpret_cast<float4*>(mo) = *reinterpret_cast<const float4*>(mom_at);
  *reinterpret_cast<float4*>(mo + 4) = *reinterpret_cast<const float4*>(mom_at + 4);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float g = acc[e] * a.inv_world + (x[e] - th[e]);
    mo[e] = a.first_round ? g : (a.mu * mo[e] + g);
    th[e] += a.lr * (a.mu * mo[e] + g);
  }
  *reinterpret_cast<float4*>(theta_at) = *reinterpret_cast<float4*>(th);
  *reinterpret_cast<float4*>(theta_at + 4) = *reinterpret_cast<float4*>(th + 4);
  *reinterpret_cast<float4*>(mom_at) = *reinterpret_cast<float4*>(mo);
  *reinterpret_cast<float4*>(mom_at + 4) = *reinterpret_cast<float4*>(mo + 4);
  bf16x8 pn = pack8(th);
  return *reinterpret_cast<uint4*>(&pn);
}

// ---- Phase B: owner-shard reduce over peers + Nesterov + broadcast, vectors [lo,hi) of a segment
// P2P variant: two vectors per thread per trip, i.e. up to 16 peer loads of 16 B in flight per
// thread before the first one is consumed (NVLink latency ~2 us needs deep memory-level
// parallelism).  NVLS variant: the switch reduces the N deltas (multimem.ld_reduce) and replicates
// the new weights (multimem.st), so each byte crosses this GPU's links once per direction.
template <bool kNvls>
__device__ __forceinline__ void phase_b(const OuterStepArgs& a, int seg, int64_t lo, int64_t hi,
                                        int64_t tid, int64_t nthreads) {
  const int64_t seg_begin = a.seg_begin[seg];
  const bf16* anchor = a.anchor_peers[a.rank] + seg_begin;
  float* theta_outer = a.seg_theta[seg];
  float* momentum = a.seg_mom[seg];
  for (int64_t i = lo + tid; i < hi; i += 2 * nthreads) {
    const int64_t iv[2] = {i, i + nthreads};
    const int cnt = iv[1] < hi ? 2 : 1;
    if constexpr (kNvls) {
      uint4 red[2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (u < cnt) red[u] = multimem_ld_reduce_bf16x8(a.mc_delta + seg_begin + (iv[u] << 3));
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u < cnt) {
          const int64_t loc = iv[u] << 3;
          const uint4 pv = nesterov8(a, red[u], anchor + loc, theta_outer + loc, momentum + loc);
          multimem_st_bf16x8(a.mc_param + seg_begin + loc, pv);
        }
      }
    } else {
      uint4 raw[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u < cnt) {
          const int64_t off = seg_begin + (iv[u] << 3);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if ((a.member_mask >> r) & 1u) raw[u][r] = ld_peer_16(a.delta_peers[r] + off);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u < cnt) {
          const int64_t loc = iv[u] << 3;
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if ((a.member_mask >> r) & 1u) {
              float d[8];
              unpack8(*reinterpret_cast<bf16x8*>(&raw[u][r]), d);
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[e] += d[e];
            }
          }
          // the fp32 sum is handed over exactly: nesterov8 takes the sum in bf16 only for NVLS
          float x[8], th[8], mo[8];
          unpack8(ld_stream_bf16x8(anchor + loc), x);
          *reinterpret_cast<float4*>(th) = *reinterpret_cast<const float4*>(theta_outer + loc);
          *reinterpret_cast<float4*>(th + 4) = *reinterpret_cast<const float4*>(theta_outer + loc + 4);
          *reinterpret_cast<float4*>(mo) = *reinterpret_cast<const float4*>(momentum + loc);
          *reinterpret_cast<float4*>(mo + 4) = *reinterpret_cast<const float4*>(momentum + loc + 4);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float g = acc[e] * a.inv_world + (x[e] - th[e]);
            mo[e] = a.first_round ? g : (a.mu * mo[e] + g);
            th[e] += a.lr * (a.mu * mo[e] + g);
          }
          *reinterpret_cast<float4*>(theta_outer + loc) = *reinterpret_cast<float4*>(th);
          *reinterpret_cast<float4*>(theta_outer + loc + 4) = *reinterpret_cast<float4*>(th + 4);
          *reinterpret_cast<float4*>(momentum + loc) = *reinterpret_cast<float4*>(mo);
          *reinterpret_cast<float4*>(momentum + loc + 4) = *reinterpret_cast<float4*>(mo + 4);
          bf16x8 pn = pack8(th);
          const uint4 pv = *reinterpret_cast<uint4*>(&pn);
          const int64_t off = seg_begin + loc;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if ((a.member_mask >> r) & 1u) st_peer_16(a.param_peers[r] + off, pv);
          }
        }
      }
    }
  }
}

// Vector range [lo,hi) of chunk c out of K of a shard of `len` elements.
__device__ __forceinline__ void chunk_slice(int64_t len, int K, int c, int64_t& lo, int64_t& hi) {
  const int64_t nv = len >> 3;
  const int64_t per = (nv + K - 1) / K;
  lo = c * per < nv ? c * per : nv;
  hi = lo + per < nv ? lo + per : nv;
}

// Software pipeline over K chunks (chunk c = the c-th 1/K slice of EVERY owner's shard):
//   stage s :  A(chunk s)  |  B(chunk s-1)  |  C(chunk s-2)        -- cross-GPU barrier --
// A and C are local HBM streams, B is NVLink traffic, so within a stage the CTAs are split by role
// (`b_ctas` CTAs drive the links, the rest stream HBM) and the two resources are busy at the same
// time.  K = 1 degenerates to the plain A | B | C schedule with two barriers.
__global__ void __launch_bounds__(512, 1) diloco_outer_step_kernel(OuterStepArgs a,
                                                                   LocalSync* ls,
                                                                   uint64_t timeout_ns) {
  const bool nvls = a.mc_delta != nullptr && a.mc_param != nullptr;
  const int K = a.num_chunks < 1 ? 1 : a.num_chunks;
  for (int stage = 0; stage < K + 2; ++stage) {
    const bool has_a = stage < K, has_b = stage >= 1 && stage <= K, has_c = stage >= 2;
    const bool has_local = has_a || has_c;
    int nb = 0;                                   // CTAs [0, nb) drive phase B in this stage
    if (has_b) nb = has_local ? min(max(a.b_ctas, 1), (int)gridDim.x - 1) : (int)gridDim.x;
    if ((int)blockIdx.x < nb) {
      const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      const int64_t nthreads = (int64_t)nb * blockDim.x;
      for (int seg = 0; seg < a.num_segs; ++seg) {
        int64_t lo, hi;
        chunk_slice(a.seg_len[seg], K, stage - 1, lo, hi);
        if (nvls) phase_b<true>(a, seg, lo, hi, tid, nthreads);
        else phase_b<false>(a, seg, lo, hi, tid, nthreads);
      }
    } else {
      const int64_t tid = (int64_t)(blockIdx.x - nb) * blockDim.x + threadIdx.x;
      const int64_t nthreads = (int64_t)(gridDim.x - nb) * blockDim.x;
      for (int r = 0; r < a.lay_n; ++r) {
        int64_t lo, hi;
        if (has_a) {
          chunk_slice(a.lay_len[r], K, stage, lo, hi);
          phase_a(a, a.lay_begin[r], lo, hi, tid, nthreads);
        }
        if (has_c) {
          chunk_slice(a.lay_len[r], K, stage - 2, lo, hi);
          phase_c(a, a.lay_begin[r], lo, hi, tid, nthreads);
        }
      }
    }
    if (stage < K + 1) {
      // a peer that never shows up is detected at the first barrier, before any state changed
      cross_gpu_barrier(a, ls, stage & 1, a.epoch * 32u + (uint32_t)stage + 1u, timeout_ns,
                        stage == 0 ? 1u : 2u);
      if (ld_acquire_gpu(&ls->status) != 0u) return;
    }
  }
}

LocalSync* g_local_sync[16] = {nullptr};

LocalSync* local_sync_for_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (g_local_sync[dev] == nullptr) {
    cudaMalloc(&g_local_sync[dev], sizeof(LocalSync));
    cudaMemset(g_local_sync[dev], 0, sizeof(LocalSync));
  }
  return g_local_sync[dev];
}

}  // namespace

static uint64_t g_outer_timeout_ns = 30ull * 1000ull * 1000ull * 1000ull;
void set_outer_timeout_ms(uint64_t ms) { g_outer_timeout_ns = ms * 1000000ull; }

int diloco_outer_step_status(unsigned int* missing_mask) {
  LocalSync h;
  cudaMemcpy(&h, local_sync_for_device(), sizeof(LocalSync), cudaMemcpyDeviceToHost);
  if (missing_mask) *missing_mask = h.missing;
  return (int)h.status;
}

void diloco_outer_step_fused(const OuterStepArgs& a, int num_ctas, cudaStream_t s) {
  LocalSync* ls = local_sync_for_device();
  cudaMemsetAsync(ls, 0, sizeof(LocalSync), s);
  if (num_ctas <= 0 || num_ctas > kNumSMs) num_ctas = kNumSMs;
  OuterStepArgs args = a;
  uint64_t timeout = g_outer_timeout_ns;
  void* params[] = {&args, &ls, &timeout};
  cudaLaunchCooperativeKernel((void*)diloco_outer_step_kernel, dim3(num_ctas), dim3(512), params,
                              0, s);
  count_launch();
}

// -----------------------------------------------------------------------------
// Local helpers
// -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) make_delta_kernel(const float* __restrict__ master,
                                                         const bf16* __restrict__ anchor,
                                                         bf16* __restrict__ delta, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = i << 3;
    float w[8], x[8];
    *reinterpret_cast<float4*>(w) = ld_stream_f32x4(master + off);
    *reinterpret_cast<float4*>(w + 4) = ld_stream_f32x4(master + off + 4);
    unpack8(ld_stream_bf16x8(anchor + off), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] -= x[e];
    st_bf16x8(delta + off, pack8(w));
  }
}
static int flat_grid(int64_t nvec) {
  int64_t b = (nvec + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}
void diloco_make_delta(const float* master, const bf16* anchor, bf16* delta, int64_t n,
                       cudaStream_t s) {
  make_delta_kernel<<<flat_grid(n / 8), 256, 0, s>>>(master, anchor, delta, n / 8);
  count_launch();
}

// g = (delta_sum*inv_world + anchor - theta)   [delta_sum != null, shard of a reduced delta]
//   or (master - theta)                         [delta_sum == null, single worker, exact fp32]
__global__ void __launch_bounds__(256) outer_local_kernel(
    const float* __restrict__ master_in, const bf16* __restrict__ delta_sum,
    const bf16* __restrict__ anchor, float* __restrict__ theta, float* __restrict__ mom,
    float* __restrict__ master_out, bf16* __restrict__ param_out, int64_t nvec, float lr,
    float mu, int first_round, float inv_world) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = i << 3;
    float th[8], mo[8], g[8];
    *reinterpret_cast<float4*>(th) = *reinterpret_cast<const float4*>(theta + off);
    *reinterpret_cast<float4*>(th + 4) = *reinterpret_cast<const float4*>(theta + off + 4);
    *reinterpret_cast<float4*>(mo) = *reinterpret_cast<const float4*>(mom + off);
    *reinterpret_cast<float4*>(mo + 4) = *reinterpret_cast<const float4*>(mom + off + 4);
    if (delta_sum != nullptr) {
      float d[8], x[8];
      unpack8(ld_stream_bf16x8(delta_sum + off), d);
      unpack8(ld_stream_bf16x8(anchor + off), x);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = d[e] * inv_world + (x[e] - th[e]);
    } else {
      float w[8];
      *reinterpret_cast<float4*>(w) = ld_stream_f32x4(master_in + off);
      *reinterpret_cast<float4*>(w + 4) = ld_stream_f32x4(master_in + off + 4);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = w[e] - th[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mo[e] = first_round ? g[e] : (mu * mo[e] + g[e]);
      th[e] += lr * (mu * mo[e] + g[e]);
    }
    *reinterpret_cast<float4*>(theta + off) = *reinterpret_cast<float4*>(th);
    *reinterpret_cast<float4*>(theta + off + 4) = *reinterpret_cast<float4*>(th + 4);
    *reinterpret_cast<float4*>(mom + off) = *reinterpret_cast<float4*>(mo);
    *reinterpret_cast<float4*>(mom + off + 4) = *reinterpret_cast<float4*>(mo + 4);
    if (master_out != nullptr) {
      *reinterpret_cast<float4*>(master_out + off) = *reinterpret_cast<float4*>(th);
      *reinterpret_cast<float4*>(master_out + off + 4) = *reinterpret_cast<float4*>(th + 4);
    }
    if (param_out != nullptr) st_bf16x8(param_out + off, pack8(th));
  }
}
void diloco_outer_local(const float* master_in, const bf16* delta_sum, const bf16* anchor,
                        float* theta_outer, float* momentum, float* master_out, bf16* param_out,
                        int64_t n, float lr, float mu, int first_round, float inv_world,
                        cudaStream_t s) {
  outer_local_kernel<<<flat_grid(n / 8), 256, 0, s>>>(master_in, delta_sum, anchor, theta_outer,
                                                      momentum, master_out, param_out, n / 8, lr,
                                                      mu, first_round, inv_world);
  count_launch();
}

__global__ void __launch_bounds__(512) peer_pull_kernel(const uint4* __restrict__ src,
                                                        uint4* __restrict__ dst, int64_t nvec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = tid;
  for (; i + 3 * stride < nvec; i += 4 * stride) {  // 4 x 16 B in flight per thread
    uint4 v0 = ld_peer_16(src + i), v1 = ld_peer_16(src + i + stride);
    uint4 v2 = ld_peer_16(src + i + 2 * stride), v3 = ld_peer_16(src + i + 3 * stride);
    dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
  }
  for (; i < nvec; i += stride) dst[i] = ld_peer_16(src + i);
}
void peer_pull_copy(const void* src_peer, void* dst_local, int64_t bytes, int num_ctas,
                    cudaStream_t s) {
  if (num_ctas <= 0) num_ctas = kNumSMs * 2;
  peer_pull_kernel<<<num_ctas, 512, 0, s>>>(reinterpret_cast<const uint4*>(src_peer),
                                            reinterpret_cast<uint4*>(dst_local), bytes / 16);
  count_launch();
}

__global__ void __launch_bounds__(256) bf16_to_master_kernel(const bf16* __restrict__ p,
                                                             float* __restrict__ master,
                                                             int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float w[8];
    unpack8(ld_stream_bf16x8(p + (i << 3)), w);
    *reinterpret_cast<float4*>(master + (i << 3)) = *reinterpret_cast<float4*>(w);
    *reinterpret_cast<float4*>(master + (i << 3) + 4) = *reinterpret_cast<float4*>(w + 4);
  }
}
void bf16_to_master(const bf16* p, float* master, int64_t n, cudaStream_t s) {
  bf16_to_master_kernel<<<flat_grid(n / 8), 256, 0, s>>>(p, master, n / 8);
  count_launch();
}

}  // namespace hb
