// pybind11 surface of the sm_100a kernel library + peer-memory transport.
// Pointers and streams cross the boundary as integers (tensor.data_ptr(),
// torch.cuda.current_stream().cuda_stream); hypha_b200/ops validates shapes.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stdexcept>
#include <vector>

#include "comm/peer_mem.h"
#include "comm/vmm.h"
#include "kernels/kernels.h"

namespace py = pybind11;
using hb::bf16;
using u64 = uintptr_t;

template <class T>
static T* P(u64 v) { return reinterpret_cast<T*>(v); }
static cudaStream_t S(u64 v) { return reinterpret_cast<cudaStream_t>(v); }

static void check_last(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
  }
}

PYBIND11_MODULE(_C, m) {
  m.doc() = "hypha_b200 sm_100a kernels";
  m.attr("ARCH") = "sm_100a";

  m.def("launch_count", &hb::launch_count);
  m.def("reset_launch_count", &hb::reset_launch_count);

  m.def("rmsnorm_fwd", [](u64 x, u64 res_in, u64 res_out, u64 w, u64 y, u64 rstd, int rows,
                          int cols, float eps, u64 s) {
    hb::rmsnorm_fwd(P<const bf16>(x), P<const bf16>(res_in), P<bf16>(res_out), P<const bf16>(w),
                    P<bf16>(y), P<float>(rstd), rows, cols, eps, S(s));
    check_last("rmsnorm_fwd");
  });
  m.def("rmsnorm_bwd_num_partials", &hb::rmsnorm_bwd_num_partials);
  m.def("rmsnorm_bwd", [](u64 dy, u64 x, u64 w, u64 rstd, u64 dres_in, u64 dx, u64 dw_partial,
                          int rows, int cols, u64 s) {
    hb::rmsnorm_bwd(P<const bf16>(dy), P<const bf16>(x), P<const bf16>(w), P<const float>(rstd),
                    P<const bf16>(dres_in), P<bf16>(dx), P<float>(dw_partial), rows, cols, S(s));
    check_last("rmsnorm_bwd");
  });
  m.def("colsum_partials", [](u64 partial, int nparts, int cols, u64 out, int accumulate, u64 s) {
    hb::colsum_partials(P<const float>(partial), nparts, cols, P<bf16>(out), accumulate, S(s));
    check_last("colsum_partials");
  });
  m.def("rope_inplace", [](u64 qkv, u64 cos, u64 sin, int tokens, int seq_len, int n_rot_heads,
                           int head_dim, int row_stride, int backward, u64 s) {
    hb::rope_inplace(P<bf16>(qkv), P<const float>(cos), P<const float>(sin), tokens, seq_len,
                     n_rot_heads, head_dim, row_stride, backward, S(s));
    check_last("rope_inplace");
  });
  m.def("swiglu_fwd", [](u64 gu, u64 out, int rows, int inter, u64 s) {
    hb::swiglu_fwd(P<const bf16>(gu), P<bf16>(out), rows, inter, S(s));
    check_last("swiglu_fwd");
  });
  m.def("swiglu_bwd", [](u64 dout, u64 gu, u64 dgu, int rows, int inter, u64 s) {
    hb::swiglu_bwd(P<const bf16>(dout), P<const bf16>(gu), P<bf16>(dgu), rows, inter, S(s));
    check_last("swiglu_bwd");
  });
  m.def("cross_entropy_fwd_bwd", [](u64 logits, u64 targets, u64 loss_rows, int rows, int vocab,
                                    int row_stride, float grad_scale, int64_t ignore_index,
                                    u64 s) {
    hb::cross_entropy_fwd_bwd(P<bf16>(logits), P<const int64_t>(targets), P<float>(loss_rows),
                              rows, vocab, row_stride, grad_scale, ignore_index, S(s));
    check_last("cross_entropy_fwd_bwd");
  });
  m.def("embedding_fwd", [](u64 ids, u64 table, u64 out, int tokens, int dim, u64 s) {
    hb::embedding_fwd(P<const int64_t>(ids), P<const bf16>(table), P<bf16>(out), tokens, dim, S(s));
    check_last("embedding_fwd");
  });
  m.def("embedding_bwd", [](u64 ids, u64 dout, u64 dtable, int tokens, int dim, u64 s) {
    hb::embedding_bwd(P<const int64_t>(ids), P<const bf16>(dout), P<bf16>(dtable), tokens, dim, S(s));
    check_last("embedding_bwd");
  });
  m.def("add_f32_into_bf16", [](u64 src, u64 dst, int64_t n, int accumulate, u64 s) {
    hb::add_f32_into_bf16(P<const float>(src), P<bf16>(dst), n, accumulate, S(s));
    check_last("add_f32_into_bf16");
  });

  m.def("adamw_flat", [](u64 p, u64 g, u64 master, u64 mm, u64 vv, int state_bf16, int64_t n,
                         float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                         float grad_scale, uint32_t seed, u64 s) {
    hb::adamw_flat(P<bf16>(p), P<const bf16>(g), P<float>(master), P<void>(mm), P<void>(vv),
                   state_bf16, n, lr, b1, b2, eps, wd, bc1, bc2, grad_scale, seed, S(s));
    check_last("adamw_flat");
  });

  m.def("diloco_make_delta", [](u64 master, u64 anchor, u64 delta, int64_t n, u64 s) {
    hb::diloco_make_delta(P<const float>(master), P<const bf16>(anchor), P<bf16>(delta), n, S(s));
    check_last("diloco_make_delta");
  });
  m.def("diloco_outer_local", [](u64 master_in, u64 delta_sum, u64 anchor, u64 theta, u64 mom,
                                 u64 master_out, u64 param_out, int64_t n, float lr, float mu,
                                 int first_round, float inv_world, u64 s) {
    hb::diloco_outer_local(P<const float>(master_in), P<const bf16>(delta_sum),
                           P<const bf16>(anchor), P<float>(theta), P<float>(mom),
                           P<float>(master_out), P<bf16>(param_out), n, lr, mu, first_round,
                           inv_world, S(s));
    check_last("diloco_outer_local");
  });
  m.def("diloco_outer_step_fused",
        [](std::vector<u64> delta_peers, std::vector<u64> param_peers, std::vector<u64> anchor_peers,
           std::vector<u64> flag_peers, u64 master, std::vector<int64_t> seg_begin,
           std::vector<int64_t> seg_len, std::vector<u64> seg_theta, std::vector<u64> seg_mom, int64_t n,
           int rank, uint32_t member_mask, uint32_t epoch, float lr, float mu, int first_round,
           int num_ctas, u64 mc_delta, u64 mc_param, std::vector<int64_t> lay_begin,
           std::vector<int64_t> lay_len, int num_chunks, int b_ctas, u64 s) {
          hb::OuterStepArgs a{};
          if (delta_peers.size() > 8) throw std::runtime_error("at most 8 ranks per NVSwitch box");
          if (seg_begin.size() > 8 || seg_begin.size() != seg_len.size() ||
              seg_begin.size() != seg_theta.size() || seg_begin.size() != seg_mom.size())
            throw std::runtime_error("bad segment lists");
          int world = 0;
          for (size_t r = 0; r < delta_peers.size(); ++r) {
            a.delta_peers[r] = P<const bf16>(delta_peers[r]);
            a.param_peers[r] = P<bf16>(param_peers[r]);
            a.anchor_peers[r] = P<bf16>(anchor_peers[r]);
            a.flag_peers[r] = P<uint32_t>(flag_peers[r]);
            if ((member_mask >> r) & 1u) ++world;
          }
          a.master = P<float>(master);
          a.num_segs = (int)seg_begin.size();
          for (size_t i = 0; i < seg_begin.size(); ++i) {
            a.seg_begin[i] = seg_begin[i];
            a.seg_len[i] = seg_len[i];
            a.seg_theta[i] = P<float>(seg_theta[i]);
            a.seg_mom[i] = P<float>(seg_mom[i]);
          }
          a.n = n;
          a.world = world; a.rank = rank; a.member_mask = member_mask; a.epoch = epoch;
          a.lr = lr; a.mu = mu; a.first_round = first_round;
          a.inv_world = 1.0f / (float)(world > 0 ? world : 1);
          a.mc_delta = P<const bf16>(mc_delta);
          a.mc_param = P<bf16>(mc_param);
          if (lay_begin.empty()) { lay_begin = {0}; lay_len = {n}; }
          if (lay_begin.size() > 8 || lay_begin.size() != lay_len.size())
            throw std::runtime_error("bad shard layout");
          a.lay_n = (int)lay_begin.size();
          for (size_t i = 0; i < lay_begin.size(); ++i) { a.lay_begin[i] = lay_begin[i]; a.lay_len[i] = lay_len[i]; }
          if (num_chunks < 1) num_chunks = 1;
          if (num_chunks > 30) num_chunks = 30;   // barrier tokens: epoch*32 + stage + 1
          a.num_chunks = num_chunks;
          a.b_ctas = b_ctas;
          hb::diloco_outer_step_fused(a, num_ctas, S(s));
          check_last("diloco_outer_step_fused");
        });
  m.def("diloco_outer_step_status", []() {
    unsigned int missing = 0;
    int st = hb::diloco_outer_step_status(&missing);
    return py::make_tuple(st, missing);
  });
  m.def("set_outer_timeout_ms", &hb::set_outer_timeout_ms);
  m.def("peer_pull_copy", [](u64 src, u64 dst, int64_t bytes, int num_ctas, u64 s) {
    hb::peer_pull_copy(P<const void>(src), P<void>(dst), bytes, num_ctas, S(s));
    check_last("peer_pull_copy");
  });
  m.def("bf16_to_master", [](u64 p, u64 master, int64_t n, u64 s) {
    hb::bf16_to_master(P<const bf16>(p), P<float>(master), n, S(s));
    check_last("bf16_to_master");
  });

  m.def("gemm_bf16", [](u64 A, u64 B, u64 D, int M, int N, int K, int lda, int ldb, int ldd,
                        int a_kmajor, int b_kmajor, int accumulate, u64 s, u64 act, int ld_act, u64 gu, int ld_gu) {
    int rc = hb::gemm_bf16_tcgen05(P<const bf16>(A), P<const bf16>(B), P<bf16>(D), M, N, K, lda,
                                   ldb, ldd, a_kmajor, b_kmajor, accumulate, S(s), P<bf16>(act), ld_act,
                                   P<const bf16>(gu), ld_gu);
    if (rc != 0) throw std::runtime_error(std::string("gemm_bf16_tcgen05: ") + hb::gemm_last_error());
  }, py::arg("A"), py::arg("B"), py::arg("D"), py::arg("M"), py::arg("N"), py::arg("K"), py::arg("lda"),
     py::arg("ldb"), py::arg("ldd"), py::arg("a_kmajor"), py::arg("b_kmajor"), py::arg("accumulate"), py::arg("s"),
     py::arg("act") = 0, py::arg("ld_act") = 0, py::arg("gu") = 0, py::arg("ld_gu") = 0);

  m.def("attention_fwd", [](u64 qkv, u64 out, u64 lse, int batch, int seq, int n_heads, int n_kv, int head_dim,
                            int qkv_stride, int out_stride, float scale, u64 s) {
    int rc = hb::attention_fwd_tcgen05(P<const bf16>(qkv), P<bf16>(out), P<float>(lse), batch, seq, n_heads, n_kv,
                                       head_dim, qkv_stride, out_stride, scale, S(s));
    if (rc != 0) throw std::runtime_error(std::string("attention_fwd_tcgen05: ") + hb::attention_last_error());
  });
  m.def("attention_bwd", [](u64 qkv, u64 out, u64 dout, u64 lse, u64 dqkv, u64 ws, int batch, int seq, int n_heads,
                            int n_kv, int head_dim, int qkv_stride, int out_stride, int dout_stride, int dqkv_stride,
                            float scale, u64 s) {
    int rc = hb::attention_bwd_tcgen05(P<const bf16>(qkv), P<const bf16>(out), P<const bf16>(dout), P<const float>(lse),
                                       P<bf16>(dqkv), P<float>(ws), batch, seq, n_heads, n_kv, head_dim, qkv_stride,
                                       out_stride, dout_stride, dqkv_stride, scale, S(s));
    if (rc != 0) throw std::runtime_error(std::string("attention_bwd_tcgen05: ") + hb::attention_bwd_last_error());
  });
  m.def("attention_fwd_set_version", &hb::attention_fwd_set_version);
  m.def("attention_fwd_get_version", &hb::attention_fwd_get_version);
  m.def("attention_fwd_set_poly", &hb::attention_fwd_set_poly);
  m.def("attention_bwd_fused_ws_floats", [](int batch, int seq, int n_heads) {
    return (unsigned long long)hb::attention_bwd_fused_ws_floats(batch, seq, n_heads);
  });
  m.def("attention_bwd_fused", [](u64 qkv, u64 out, u64 dout, u64 lse, u64 dqkv, u64 ws, int batch, int seq, int n_heads,
                                  int n_kv, int head_dim, int qkv_stride, int out_stride, int dout_stride,
                                  int dqkv_stride, float scale, u64 s) {
    int rc = hb::attention_bwd_fused_tcgen05(P<const bf16>(qkv), P<const bf16>(out), P<const bf16>(dout),
                                             P<const float>(lse), P<bf16>(dqkv), P<float>(ws), batch, seq, n_heads, n_kv,
                                             head_dim, qkv_stride, out_stride, dout_stride, dqkv_stride, scale, S(s));
    if (rc != 0)
      throw std::runtime_error(std::string("attention_bwd_fused_tcgen05: ") + hb::attention_bwd_fused_last_error());
  });
  m.def("gemm_fp8", [](u64 A, u64 B, u64 D, int M, int N, int K, int lda, int ldb, int ldd, int a_kmajor,
                       int b_kmajor, int a_fmt, int b_fmt, u64 dq_a, u64 dq_b, int accumulate, u64 s, u64 act, int ld_act) {
    int rc = hb::gemm_fp8_tcgen05(P<const uint8_t>(A), P<const uint8_t>(B), P<bf16>(D), M, N, K, lda, ldb, ldd,
                                  a_kmajor, b_kmajor, a_fmt, b_fmt, P<const float>(dq_a), P<const float>(dq_b),
                                  accumulate, S(s), P<bf16>(act), ld_act);
    if (rc != 0) throw std::runtime_error(std::string("gemm_fp8_tcgen05: ") + hb::gemm_last_error());
  }, py::arg("A"), py::arg("B"), py::arg("D"), py::arg("M"), py::arg("N"), py::arg("K"), py::arg("lda"),
     py::arg("ldb"), py::arg("ldd"), py::arg("a_kmajor"), py::arg("b_kmajor"), py::arg("a_fmt"), py::arg("b_fmt"),
     py::arg("dq_a"), py::arg("dq_b"), py::arg("accumulate"), py::arg("s"), py::arg("act") = 0, py::arg("ld_act") = 0);
  m.def("gemm_mxfp8", [](u64 A, u64 sfa, u64 B, u64 sfb, u64 D, int M, int N, int K, int lda, int ldb, int ldd, int a_fmt,
                         int b_fmt, int accumulate, u64 s, u64 act, int ld_act, u64 gu, int ld_gu) {
    int rc = hb::gemm_mxfp8_tcgen05(P<const uint8_t>(A), P<const uint8_t>(sfa), P<const uint8_t>(B), P<const uint8_t>(sfb),
                                    P<bf16>(D), M, N, K, lda, ldb, ldd, a_fmt, b_fmt, accumulate, S(s), P<bf16>(act),
                                    ld_act, P<const bf16>(gu), ld_gu);
    if (rc != 0) throw std::runtime_error(std::string("gemm_mxfp8_tcgen05: ") + hb::gemm_last_error());
    check_last("gemm_mxfp8");
  });
  m.def("quantize_mxfp8", [](u64 x, int ldx, u64 q, int ldq, u64 sf, u64 qT, int ldqT, u64 sfT, int R, int C, u64 s) {
    int rc = hb::quantize_mxfp8(P<const bf16>(x), ldx, P<uint8_t>(q), ldq, P<uint8_t>(sf), P<uint8_t>(qT), ldqT,
                                P<uint8_t>(sfT), R, C, S(s));
    if (rc != 0) throw std::runtime_error("quantize_mxfp8: needs R, C % 128 == 0, 16-byte aligned pitches, an output pair");
    check_last("quantize_mxfp8");
  });
  m.def("amax_bf16", [](u64 x, int64_t n, u64 amax, u64 s) {
    hb::amax_bf16(P<const bf16>(x), n, P<float>(amax), S(s));
    check_last("amax_bf16");
  });
  m.def("quantize_fp8", [](u64 x, u64 out, int64_t n, u64 amax, u64 dequant, int fmt, u64 s) {
    hb::quantize_fp8(P<const bf16>(x), P<uint8_t>(out), n, P<const float>(amax), P<float>(dequant), fmt, S(s));
    check_last("quantize_fp8");
  });
  m.def("gemm_set_cta_mode", &hb::gemm_set_cta_mode);
  m.def("gemm_get_cta_mode", &hb::gemm_get_cta_mode);

  // ---- peer memory ------------------------------------------------------------
  m.def("sym_alloc", &hb::comm::sym_alloc, py::arg("bytes"), py::arg("zero") = true);
  m.def("sym_free", &hb::comm::sym_free);
  m.def("ipc_export", [](u64 ptr) { return py::bytes(hb::comm::ipc_export(ptr)); });
  m.def("ipc_open", [](py::bytes h) { return hb::comm::ipc_open(std::string(h)); });
  m.def("ipc_close", &hb::comm::ipc_close);
  m.def("enable_peer_access", &hb::comm::enable_peer_access);
  m.def("sym_bytes_allocated", &hb::comm::bytes_allocated);
  m.def("memcpy_d2d_async", &hb::comm::memcpy_d2d_async);
  m.def("memset_async", &hb::comm::memset_async);

  // ---- VMM / NVSwitch multicast / descriptor passing / membership epochs (csrc/comm/vmm.cpp) ----
  {
    namespace vm = hb::comm::vmm;
    static py::exception<vm::Unavailable> unavailable(m, "VmmUnavailable", PyExc_RuntimeError);
    m.def("vmm_driver_available", &vm::driver_available);
    m.def("vmm_multicast_supported", &vm::multicast_supported);
    m.def("vmm_usable", &vm::usable, py::call_guard<py::gil_scoped_release>());
    m.def("vmm_granularity", &vm::granularity);
    m.def("vmm_create", [](int device, size_t bytes, int world) {
      vm::AllocInfo a = vm::create(device, bytes, world);
      return py::make_tuple(a.id, a.size, a.gran);
    });
    m.def("vmm_export_fd", &vm::export_fd);
    m.def("vmm_import_fd", &vm::import_fd, py::arg("fd"), py::arg("multicast") = false);
    m.def("vmm_map", &vm::map);
    m.def("vmm_unmap", &vm::unmap);
    m.def("vmm_release", &vm::release);
    m.def("vmm_live_handles", &vm::live_handles);
    m.def("vmm_mc_create", &vm::mc_create);
    m.def("vmm_mc_add_device", &vm::mc_add_device);
    m.def("vmm_mc_bind", &vm::mc_bind);
    m.def("uds_listen", &vm::uds_listen);
    m.def("uds_close", &vm::uds_close);
    m.def("uds_send_fds", &vm::uds_send_fds, py::call_guard<py::gil_scoped_release>());
    m.def("uds_recv_fds", &vm::uds_recv_fds, py::call_guard<py::gil_scoped_release>());
    m.def("epoch_stress", &vm::epoch_stress, py::call_guard<py::gil_scoped_release>());
    py::class_<vm::EpochTable>(m, "EpochTable")
        .def(py::init<int>())
        .def("snapshot", [](const vm::EpochTable& t) {
          const vm::EpochSnapshot s = t.snapshot();
          py::dict d;
          d["epoch"] = s.epoch;
          d["flag_epoch"] = s.flag_epoch;
          d["mask"] = s.mask;
          d["world"] = s.world;
          d["owner"] = std::vector<int>(s.owner, s.owner + s.world);
          return d;
        })
        .def("drop", &vm::EpochTable::drop, py::arg("rank"), py::arg("adopter") = -1)
        .def("set_mask", &vm::EpochTable::set_mask)
        .def("set_owner", &vm::EpochTable::set_owner)
        .def("admit", &vm::EpochTable::admit, py::arg("rank"), py::arg("take_back_shard") = true)
        .def("set_flag_epoch", &vm::EpochTable::set_flag_epoch);
  }
}
