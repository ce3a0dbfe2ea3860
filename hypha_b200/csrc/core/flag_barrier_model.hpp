// Host model of the cross-GPU flag barrier that diloco_outer_step_kernel runs on the device
// (csrc/kernels/outer_step.cu: cross_gpu_barrier).  One std::thread per "rank", one flag array per
// rank ([2 slots][8 peers] of 32-bit tokens), release stores into the PEER's array, acquire polling of
// the OWN array -- the same data flow as st.release.sys / ld.acquire.sys over peer-mapped memory.
//
//   token(epoch, stage) = epoch * 32 + stage + 1      slot(stage) = stage & 1
//   a rank passes barrier (epoch, stage) once every member's entry in its own array is >= token
//   (signed distance, so the 32-bit counter may wrap), or gives up after `timeout` and reports who is
//   missing.
//
// The stress test drives many epochs of the K+2-stage software pipeline with random skew and checks
// the property the kernel relies on: no rank leaves barrier (e, s) before every member has entered it,
// although only two slots are ever used.  Built into _core (unit test) and as a standalone
// ThreadSanitizer binary (tools/sanitize.sh tsan).  The reference has no race tooling.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>

namespace hypha {

struct FlagBarrierReport {
  bool ok = true;
  uint64_t barriers_passed = 0;     // summed over ranks
  uint64_t violations = 0;          // a rank left a barrier before all members had entered it
  uint32_t timeouts = 0;            // barrier instances that timed out (summed over ranks)
  uint32_t missing_union = 0;       // union of the missing masks reported by the survivors
};

class FlagBarrierModel {
 public:
  FlagBarrierModel(int world, uint32_t member_mask)
      : world_(world), mask_(member_mask), flags_(static_cast<size_t>(world) * 16), entered_(kTrack * 8) {
    for (auto& f : flags_) f.store(0, std::memory_order_relaxed);
    for (auto& e : entered_) e.store(0, std::memory_order_relaxed);
  }

  // Returns the mask of members that never arrived (0 = everyone did).
  uint32_t barrier(int rank, uint32_t epoch, int stage, std::chrono::microseconds timeout) {
    const uint32_t token = epoch * 32u + static_cast<uint32_t>(stage) + 1u;
    const int slot = stage & 1;
    for (int peer = 0; peer < world_; ++peer)
      if ((mask_ >> peer) & 1u) flag(peer, slot, rank).store(token, std::memory_order_release);
    uint32_t missing = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int peer = 0; peer < world_; ++peer) {
      if (!((mask_ >> peer) & 1u)) continue;
      while (static_cast<int32_t>(flag(rank, slot, peer).load(std::memory_order_acquire) - token) < 0) {
        if (std::chrono::steady_clock::now() - t0 > timeout) { missing |= 1u << peer; break; }
        std::this_thread::yield();
      }
    }
    return missing;
  }

  // Run `epochs` launches of a (stages)-barrier pipeline on every member except `dead_rank`.
  FlagBarrierReport stress(int epochs, int stages, uint32_t seed, int dead_rank = -1, int timeout_ms = 2000,
                           uint32_t first_epoch = 1) {
    FlagBarrierReport rep;
    // flags as the previous launch left them (a fresh zero flag only works while tokens are < 2^31,
    // i.e. for the first 2^26 launches -- which is how the device buffers start out)
    if (first_epoch > 1)
      for (auto& f : flags_) f.store(first_epoch * 32u, std::memory_order_relaxed);
    std::atomic<uint64_t> passed{0}, violations{0};
    std::atomic<uint32_t> timeouts{0}, missing_union{0};
    int members = 0;
    for (int r = 0; r < world_; ++r) members += (mask_ >> r) & 1u;
    std::vector<std::thread> th;
    for (int r = 0; r < world_; ++r) {
      if (!((mask_ >> r) & 1u) || r == dead_rank) continue;
      th.emplace_back([&, r] {
        std::mt19937 rng(seed * 977u + static_cast<uint32_t>(r));
        for (int e = 0; e < epochs; ++e) {
          const uint32_t epoch = first_epoch + static_cast<uint32_t>(e);
          bool aborted = false;
          for (int s = 0; s < stages && !aborted; ++s) {
            if ((rng() & 7u) == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 200));
            const uint64_t id = static_cast<uint64_t>(epoch) * 32u + static_cast<uint64_t>(s);
            entered(id, r).store(id + 1, std::memory_order_release);          // "phase work done, entering"
            const uint32_t miss = barrier(r, epoch, s, std::chrono::milliseconds(timeout_ms));
            if (miss) {
              timeouts.fetch_add(1);
              missing_union.fetch_or(miss);
              aborted = true;                                                 // the kernel returns status != 0
              break;
            }
            for (int p = 0; p < world_; ++p)                                  // everyone must have ENTERED (e, s)
              if (((mask_ >> p) & 1u) && entered(id, p).load(std::memory_order_acquire) < id + 1)
                violations.fetch_add(1);
            passed.fetch_add(1);
          }
          if (aborted) break;
        }
      });
    }
    for (auto& t : th) t.join();
    rep.barriers_passed = passed.load();
    rep.violations = violations.load();
    rep.timeouts = timeouts.load();
    rep.missing_union = missing_union.load();
    rep.ok = rep.violations == 0 && (dead_rank >= 0 ? rep.timeouts > 0 : rep.timeouts == 0);
    (void)members;
    return rep;
  }

 private:
  static constexpr size_t kTrack = 4096;                    // ring of barrier ids for the invariant check
  std::atomic<uint32_t>& flag(int owner, int slot, int from) {
    return flags_[static_cast<size_t>(owner) * 16 + static_cast<size_t>(slot) * 8 + static_cast<size_t>(from)];
  }
  // entered(id, r) holds the largest id+1 rank r has entered among ids congruent to id (mod kTrack);
  // ranks are never more than one barrier apart, so the ring cannot alias within a test.
  std::atomic<uint64_t>& entered(uint64_t id, int r) {
    return entered_[(id % kTrack) * 8 + static_cast<size_t>(r)];
  }
  int world_;
  uint32_t mask_;
  std::vector<std::atomic<uint32_t>> flags_;
  std::vector<std::atomic<uint64_t>> entered_;
};

}  // namespace hypha
