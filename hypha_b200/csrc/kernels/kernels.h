// Host-side launch API of the hypha_b200 sm_100a kernels (raw pointers +
// stream; the Python layer validates dtype/shape/contiguity).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {
using bf16 = __nv_bfloat16;

// ---- normalisation / elementwise (elementwise.cu) ---------------------------
void rmsnorm_fwd(const bf16* x, const bf16* res_in, bf16* res_out, const bf16* w, bf16* y,
                 float* rstd, int rows, int cols, float eps, cudaStream_t s);
int rmsnorm_bwd_num_partials(int rows);
void rmsnorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* rstd,
                 const bf16* dres_in, bf16* dx, float* dw_partial, int rows, int cols,
                 cudaStream_t s);
void colsum_partials(const float* partial, int nparts, int cols, bf16* out, int accumulate,
                     cudaStream_t s);
void rope_inplace(bf16* qkv, const float* cos, const float* sin, int tokens, int seq_len,
                  int n_rot_heads, int head_dim, int row_stride, int backward, cudaStream_t s);
void swiglu_fwd(const bf16* gu, bf16* out, int rows, int inter, cudaStream_t s);
void swiglu_bwd(const bf16* dout, const bf16* gu, bf16* dgu, int rows, int inter, cudaStream_t s);
void cross_entropy_fwd_bwd(bf16* logits, const int64_t* targets, float* loss_rows, int rows,
                           int vocab, int row_stride, float grad_scale, int64_t ignore_index,
                           cudaStream_t s);
void embedding_fwd(const int64_t* ids, const bf16* table, bf16* out, int tokens, int dim,
                   cudaStream_t s);
void embedding_bwd(const int64_t* ids, const bf16* dout, bf16* dtable, int tokens, int dim,
                   cudaStream_t s);
void add_f32_into_bf16(const float* src, bf16* dst, int64_t n, int accumulate, cudaStream_t s);

// ---- optimizers (optim.cu) --------------------------------------------------
void adamw_flat(bf16* p, const bf16* g, float* master, void* m, void* v, int state_bf16,
                int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                float bias_corr1, float bias_corr2, float grad_scale, uint32_t seed,
                cudaStream_t s);

// ---- DiLoCo outer step (outer_step.cu) --------------------------------------
// Local part: delta[i] = bf16(master[i] - float(anchor[i]))
void diloco_make_delta(const float* master, const bf16* anchor, bf16* delta, int64_t n,
                       cudaStream_t s);
struct OuterStepArgs {
  // per-rank symmetric buffers (device pointers valid on THIS rank)
  const bf16* delta_peers[8];   // peer-mapped delta buffers, index = rank
  bf16* param_peers[8];         // peer-mapped bf16 parameter buffers
  bf16* anchor_peers[8];        // peer-mapped bf16 anchor (round-start weights)
  uint32_t* flag_peers[8];      // peer-mapped barrier flag words [>= 64 u32 each]
  float* master;                // local fp32 master weights (full)
  // Segments of the parameter vector this rank currently owns: its own shard plus any
  // orphaned shards adopted from ranks that dropped out (fault tolerance).
  int num_segs;
  int64_t seg_begin[8];         // first element of the segment
  int64_t seg_len[8];           // elements in the segment
  float* seg_theta[8];          // fp32 outer params of the segment
  float* seg_mom[8];            // fp32 outer momentum of the segment
  int64_t n;                    // total elements
  int world;                    // participating ranks this round
  int rank;                     // my index among participants
  uint32_t member_mask;         // bit r set => rank r participates
  uint32_t epoch;               // monotonically increasing round id
  float lr, mu;
  int first_round;              // momentum := g on the first round
  float inv_world;              // 1 / number of contributing ranks
  // NVLS (NVSwitch multicast) path: multicast addresses of the delta and parameter buffers.
  // When set and every rank participates, phase B uses multimem.ld_reduce / multimem.st.
  const bf16* mc_delta;
  bf16* mc_param;
  // Shard layout of the whole vector (independent of who currently owns which shard) and the
  // software-pipeline shape: chunk c = the c-th 1/num_chunks slice of every shard.
  int lay_n;
  int64_t lay_begin[8];
  int64_t lay_len[8];
  int num_chunks;               // 1 = plain A | B | C
  int b_ctas;                   // CTAs that drive NVLink (phase B) while the rest stream HBM
};
void diloco_outer_step_fused(const OuterStepArgs& a, int num_ctas, cudaStream_t s);
// Local Nesterov step.  delta_sum == null: single worker, exact fp32
// g = master_in - theta.  Otherwise g = delta_sum*inv_world + (anchor - theta)
// on a shard whose bf16 delta has already been reduced (NCCL fallback path).
void diloco_outer_local(const float* master_in, const bf16* delta_sum, const bf16* anchor,
                        float* theta_outer, float* momentum, float* master_out, bf16* param_out,
                        int64_t n, float lr, float mu, int first_round, float inv_world,
                        cudaStream_t s);
void set_outer_timeout_ms(uint64_t ms);
int diloco_outer_step_status(unsigned int* missing_mask);
// peer-memory broadcast: copy bf16 params from a live rank (pull over NVLink)
void peer_pull_copy(const void* src_peer, void* dst_local, int64_t bytes, int num_ctas,
                    cudaStream_t s);
void bf16_to_master(const bf16* p, float* master, int64_t n, cudaStream_t s);

// ---- tcgen05 GEMM (gemm_tcgen05.cu) ------------------------------------------
// D[M,N] (+)= A * B with bf16 inputs, fp32 TMEM accumulation, bf16 output.
//  a_kmajor: A stored [M,K] (K contiguous) else stored [K,M]
//  b_kmajor: B stored [N,K] (K contiguous) else stored [K,N]
// lda/ldb/ldd are row strides (elements) of the stored matrices.
// `act` (optional): SwiGLU epilogue -- the N output columns interleave (gate, up) pairs and the kernel ALSO writes
// silu(gate) * up to act[M, N/2] (row stride ld_act), computed from the bf16-rounded pair.
int gemm_bf16_tcgen05(const bf16* A, const bf16* B, bf16* D, int M, int N, int K, int lda,
                      int ldb, int ldd, int a_kmajor, int b_kmajor, int accumulate,
                      cudaStream_t s, bf16* act = nullptr, int ld_act = 0, const bf16* gu = nullptr, int ld_gu = 0);
// `gu` (with `act`): SwiGLU-BACKWARD epilogue -- the [M, N] product is d(act); D is not written, act[M, 2N] receives
// d(gate|up) = swiglu'(gu) * d(act) with interleaved columns (gu[M, 2N] is the forward's gate|up activation).
// fp8 (kind::f8f6f4): D (+)= (*dequant_a * *dequant_b) * (A . B); formats 0 = e4m3, 1 = e5m2.
int gemm_fp8_tcgen05(const uint8_t* A, const uint8_t* B, bf16* D, int M, int N, int K, int lda,
                     int ldb, int ldd, int a_kmajor, int b_kmajor, int a_fmt, int b_fmt,
                     const float* dequant_a, const float* dequant_b, int accumulate, cudaStream_t s,
                     bf16* act = nullptr, int ld_act = 0);
// MXFP8 (kind::mxf8f6f4.block_scale): both operands K-major, UE8M0 scales per 32 K-elements in the atom layout
// `quantize_mxfp8` writes; optional fused SwiGLU / SwiGLU-backward epilogues as in gemm_bf16_tcgen05.
int gemm_mxfp8_tcgen05(const uint8_t* A, const uint8_t* sfa, const uint8_t* B, const uint8_t* sfb, bf16* D, int M, int N,
                       int K, int lda, int ldb, int ldd, int a_fmt, int b_fmt, int accumulate, cudaStream_t s,
                       bf16* act = nullptr, int ld_act = 0, const bf16* gu = nullptr, int ld_gu = 0);
int quantize_mxfp8(const bf16* x, int ldx, uint8_t* q, int ldq, uint8_t* sf, uint8_t* qT, int ldqT, uint8_t* sfT, int R,
                   int C, cudaStream_t s);
void amax_bf16(const bf16* x, int64_t n, float* amax, cudaStream_t s);
void quantize_fp8(const bf16* x, uint8_t* out, int64_t n, const float* amax, float* dequant, int fmt,
                  cudaStream_t s);
const char* gemm_last_error();
void gemm_set_cta_mode(int mode);   // 1 = single-CTA tiles, 2 = cta_group::2 CTA pairs
int gemm_get_cta_mode();

// ---- tcgen05 flash-attention forward (attention_tcgen05.cu) ------------------------------------
int attention_fwd_tcgen05(const bf16* qkv, bf16* out, float* lse, int batch, int seq, int n_heads, int n_kv,
                          int head_dim, int qkv_stride, int out_stride, float scale, cudaStream_t s);
const char* attention_last_error();
void attention_fwd_set_version(int v);   // 2 = two-Q-tile ping-pong kernel (default), 1 = one Q tile per CTA
int attention_fwd_get_version();
void attention_fwd_set_poly(int n);      // 0..2 of every 4 exponential pairs on the FMA pipe instead of MUFU
// ---- tcgen05 flash-attention backward (attention_bwd_tcgen05.cu) ---------------------------------
// ws: fp32 workspace of 2 * batch * n_heads * seq floats.  dqkv is fully overwritten.
int attention_bwd_tcgen05(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, float* ws,
                          int batch, int seq, int n_heads, int n_kv, int head_dim, int qkv_stride, int out_stride,
                          int dout_stride, int dqkv_stride, float scale, cudaStream_t s);
const char* attention_bwd_last_error();
// ---- fused single-kernel backward, 5 GEMMs per tile pair, dQ through TMA reduce-add (attention_bwd_fused.cu) ----
// ws: attention_bwd_fused_ws_floats() floats (D, lse2, fp32 dQ accumulator).  dqkv is fully overwritten.
size_t attention_bwd_fused_ws_floats(int batch, int seq, int n_heads);
int attention_bwd_fused_tcgen05(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, float* ws,
                                int batch, int seq, int n_heads, int n_kv, int head_dim, int qkv_stride, int out_stride,
                                int dout_stride, int dqkv_stride, float scale, cudaStream_t s);
const char* attention_bwd_fused_last_error();

unsigned long long launch_count();
void reset_launch_count();
}  // namespace hb
