// Host parameter-server round over memory-mapped safetensors files, one pass.
//
// What the reference's parameter server does per round (crates/worker/src/executor/parameter_server.rs:
// apply_tensor_op :331-384, average_op :194-197, update_momentum :386-416, nesterov :418-446) is four
// passes over mmapped files with a spill file per tensor per pass.  Here every worker file, the previous
// momentum and the previous cumulative update are mapped once and a pool of threads streams through them:
//
//     g   = (1/N) * sum_k delta_k[i]                      (true mean)
//     m   = first round ? g : mu * m_prev[i] + g
//     u   = lr * (mu * m + g)
//     cum = cum_prev[i] + u                               (what a joining worker needs: theta = theta_0 + cum)
//
// and writes m, u and cum straight into the (pre-sized, header already written) output files.  The caller
// parses / writes the safetensors JSON headers and passes absolute byte offsets; only F32 tensors take this
// path (the Python fallback handles everything else).  On GPUs the same arithmetic is the fused NVLink
// kernel (kernels/outer_step.cu); this is the CPU / cross-host aggregation path.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace hypha {

// safetensors data is only guaranteed byte-aligned relative to the file start
typedef float unaligned_f32 __attribute__((aligned(1), may_alias));

class MappedFile {
 public:
  MappedFile() = default;
  MappedFile(const std::string& path, bool writable) { open(path, writable); }
  MappedFile(const MappedFile&) = delete;
  MappedFile& operator=(const MappedFile&) = delete;
  ~MappedFile() { close(); }

  void open(const std::string& path, bool writable) {
    close();
    int fd = ::open(path.c_str(), writable ? O_RDWR : O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + path + ": " + std::strerror(errno));
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); throw std::runtime_error("cannot stat " + path); }
    size_ = static_cast<uint64_t>(st.st_size);
    if (size_ > 0) {
      void* p = ::mmap(nullptr, size_, writable ? (PROT_READ | PROT_WRITE) : PROT_READ, MAP_SHARED, fd, 0);
      if (p == MAP_FAILED) { ::close(fd); throw std::runtime_error("cannot mmap " + path + ": " + std::strerror(errno)); }
      base_ = static_cast<uint8_t*>(p);
      if (!writable) ::madvise(p, size_, MADV_SEQUENTIAL);
    }
    ::close(fd);
  }
  void close() {
    if (base_) ::munmap(base_, size_);
    base_ = nullptr; size_ = 0;
  }
  bool mapped() const { return base_ != nullptr; }
  uint64_t size() const { return size_; }
  uint8_t* at(uint64_t off, uint64_t bytes) const {
    if (off + bytes > size_) throw std::runtime_error("tensor range beyond the end of the file");
    return base_ + off;
  }

 private:
  uint8_t* base_ = nullptr;
  uint64_t size_ = 0;
};

struct PsRoundStats {
  uint64_t elements = 0;
  uint64_t bytes_read = 0;
  uint64_t bytes_written = 0;
  int threads = 0;
  double seconds = 0.0;
};

// in_offsets[k][t]: absolute byte offset of tensor t in worker file k; numels[t] elements each.
// prev_m / prev_cum: "" = none (first round / nothing accumulated yet); their offsets are per tensor.
// out_*: pre-sized files; out_offsets[t] applies to all three (same header).
inline PsRoundStats ps_fused_round(const std::vector<std::string>& in_paths,
                                   const std::vector<std::vector<uint64_t>>& in_offsets,
                                   const std::vector<uint64_t>& numels,
                                   const std::string& prev_m_path, const std::vector<uint64_t>& prev_m_offsets,
                                   const std::string& prev_cum_path, const std::vector<uint64_t>& prev_cum_offsets,
                                   const std::string& out_u_path, const std::string& out_m_path,
                                   const std::string& out_cum_path, const std::vector<uint64_t>& out_offsets,
                                   double lr, double mu, int threads,
                                   const std::vector<double>& weights = {}) {
  // weights: empty = true mean (every input counts 1/N); else one weight per input, normalised by the caller
  // (sample-weighted mean: w_k = samples_k / sum of samples)
  const size_t N = in_paths.size(), T = numels.size();
  if (!weights.empty() && weights.size() != N) throw std::invalid_argument("weights: one per worker file");
  if (N == 0) throw std::invalid_argument("no worker files");
  if (in_offsets.size() != N) throw std::invalid_argument("in_offsets: one list per worker file");
  for (auto& o : in_offsets) if (o.size() != T) throw std::invalid_argument("in_offsets: one offset per tensor");
  if (out_offsets.size() != T) throw std::invalid_argument("out_offsets: one offset per tensor");
  const bool have_m = !prev_m_path.empty(), have_c = !prev_cum_path.empty();
  if (have_m && prev_m_offsets.size() != T) throw std::invalid_argument("prev_m_offsets: one offset per tensor");
  if (have_c && prev_cum_offsets.size() != T) throw std::invalid_argument("prev_cum_offsets: one offset per tensor");

  std::vector<MappedFile> ins(N);
  for (size_t k = 0; k < N; ++k) ins[k].open(in_paths[k], false);
  MappedFile pm, pc, ou(out_u_path, true), om(out_m_path, true), oc(out_cum_path, true);
  if (have_m) pm.open(prev_m_path, false);
  if (have_c) pc.open(prev_cum_path, false);

  // resolve + bounds-check every tensor once, then cut the element space into chunks
  struct Piece { std::vector<const uint8_t*> in; const unaligned_f32 *m, *c; unaligned_f32 *ou, *om, *oc; uint64_t n; };
  constexpr uint64_t kChunk = 1u << 18;                      // 1 MiB of fp32 per stream and chunk
  std::vector<Piece> pieces;
  uint64_t total = 0;
  for (size_t t = 0; t < T; ++t) {
    const uint64_t n = numels[t], bytes = n * 4;
    if (n == 0) continue;
    std::vector<const uint8_t*> src(N);
    for (size_t k = 0; k < N; ++k) src[k] = ins[k].at(in_offsets[k][t], bytes);
    const unaligned_f32* m = have_m ? reinterpret_cast<const unaligned_f32*>(pm.at(prev_m_offsets[t], bytes)) : nullptr;
    const unaligned_f32* c = have_c ? reinterpret_cast<const unaligned_f32*>(pc.at(prev_cum_offsets[t], bytes)) : nullptr;
    unaligned_f32* u_out = reinterpret_cast<unaligned_f32*>(ou.at(out_offsets[t], bytes));
    unaligned_f32* m_out = reinterpret_cast<unaligned_f32*>(om.at(out_offsets[t], bytes));
    unaligned_f32* c_out = reinterpret_cast<unaligned_f32*>(oc.at(out_offsets[t], bytes));
    for (uint64_t a = 0; a < n; a += kChunk) {
      Piece p;
      p.n = std::min<uint64_t>(kChunk, n - a);
      p.in.resize(N);
      for (size_t k = 0; k < N; ++k) p.in[k] = src[k] + 4 * a;
      p.m = m ? m + a : nullptr; p.c = c ? c + a : nullptr;
      p.ou = u_out + a; p.om = m_out + a; p.oc = c_out + a;
      pieces.push_back(std::move(p));
    }
    total += n;
  }

  const float inv = 1.0f / static_cast<float>(N), fmu = static_cast<float>(mu), flr = static_cast<float>(lr);
  const bool weighted = !weights.empty();
  std::vector<float> w(N, 1.0f);
  for (size_t k = 0; weighted && k < N; ++k) w[k] = static_cast<float>(weights[k]);
  std::atomic<size_t> next{0};
  auto work = [&]() {
    float acc[1024];
    for (;;) {
      size_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= pieces.size()) return;
      const Piece& p = pieces[i];
      for (uint64_t a = 0; a < p.n; a += 1024) {
        const uint64_t b = std::min<uint64_t>(1024, p.n - a);
        const unaligned_f32* s0 = reinterpret_cast<const unaligned_f32*>(p.in[0]) + a;
        if (weighted) {
          const float w0 = w[0];
          for (uint64_t j = 0; j < b; ++j) acc[j] = w0 * s0[j];
          for (size_t k = 1; k < N; ++k) {
            const unaligned_f32* s = reinterpret_cast<const unaligned_f32*>(p.in[k]) + a;
            const float wk = w[k];
            for (uint64_t j = 0; j < b; ++j) acc[j] += wk * s[j];
          }
        } else {
          for (uint64_t j = 0; j < b; ++j) acc[j] = s0[j];
          for (size_t k = 1; k < N; ++k) {
            const unaligned_f32* s = reinterpret_cast<const unaligned_f32*>(p.in[k]) + a;
            for (uint64_t j = 0; j < b; ++j) acc[j] += s[j];
          }
        }
        const float scale = weighted ? 1.0f : inv;
        for (uint64_t j = 0; j < b; ++j) {
          const float g = acc[j] * scale;
          const float m = p.m ? fmu * p.m[a + j] + g : g;
          const float u = flr * (fmu * m + g);
          p.om[a + j] = m;
          p.ou[a + j] = u;
          p.oc[a + j] = p.c ? p.c[a + j] + u : u;
        }
      }
    }
  };
  int nt = threads > 0 ? threads : static_cast<int>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
  nt = static_cast<int>(std::min<size_t>(static_cast<size_t>(nt), std::max<size_t>(1, pieces.size())));
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int i = 1; i < nt; ++i) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  PsRoundStats s;
  s.elements = total;
  s.bytes_read = total * 4 * (N + (have_m ? 1 : 0) + (have_c ? 1 : 0));
  s.bytes_written = total * 4 * 3;
  s.threads = nt;
  s.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return s;
}

}  // namespace hypha
