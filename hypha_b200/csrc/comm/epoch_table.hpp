// Membership-epoch table of the fused outer step (header-only: no CUDA dependency, so the ThreadSanitizer target
// `tools/tsan_epoch_table.cpp` builds it on its own).  Writers -- the control plane: drop / admit / owner change /
// flag-epoch resync -- are serialised by a mutex and publish under a sequence lock; a reader (every kernel launch)
// never blocks and never sees half of an update.  Every field is an atomic, so the optimistic read is race-free in
// the C++ memory model, not just in practice.
#pragma once
#include <atomic>
#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

namespace hb {
namespace comm {
namespace vmm {

constexpr int kMaxRanks = 8;   // one NVSwitch box; `outer_step.cu` carries 8 peer pointers

struct EpochSnapshot {
  uint64_t epoch = 0;        // bumps on every membership / ownership / flag-epoch change
  uint64_t flag_epoch = 0;   // token base of the device-side flag barriers
  uint32_t mask = 0;         // members
  int world = 0;
  int owner[kMaxRanks] = {}; // shard -> rank that holds its fp32 theta_outer + momentum
};

class EpochTable {
 public:
  explicit EpochTable(int world);
  EpochSnapshot snapshot() const;                    // lock-free, consistent
  // member leaves; every shard it held goes to `adopter`, or (adopter < 0) to the next live rank after the shard
  uint64_t drop(int rank, int adopter = -1);
  uint64_t set_mask(uint32_t mask);                  // membership as decided by the control plane
  uint64_t set_owner(int shard, int rank);
  uint64_t admit(int rank, bool take_back_shard);    // member (re)joins; optionally owns its own shard again
  uint64_t set_flag_epoch(uint64_t v);

 private:
  void publish_begin();
  void publish_end();
  mutable std::mutex mu_;
  std::atomic<uint64_t> seq_{0};
  std::atomic<uint64_t> epoch_{0}, flag_epoch_{0};
  std::atomic<uint32_t> mask_{0};
  std::atomic<int> world_{0};
  std::atomic<int> owner_[kMaxRanks];
};

// ---- membership epochs ---------------------------------------------------------------------------------------------------
// Writers (the control plane: drop / admit / owner change / flag-epoch resync) are serialised by a mutex and publish
// under a sequence lock; a reader never blocks and never sees half of an update.
inline void EpochTable::publish_begin() { seq_.fetch_add(1, std::memory_order_acq_rel); }
inline void EpochTable::publish_end() { seq_.fetch_add(1, std::memory_order_release); }

inline EpochTable::EpochTable(int world) {
  if (world < 1 || world > kMaxRanks) throw std::runtime_error("EpochTable: world must be 1..8");
  world_.store(world, std::memory_order_relaxed);
  mask_.store((1u << world) - 1u, std::memory_order_relaxed);
  for (int r = 0; r < kMaxRanks; ++r) owner_[r].store(r < world ? r : -1, std::memory_order_relaxed);
}

inline EpochSnapshot EpochTable::snapshot() const {
  EpochSnapshot s;
  for (;;) {
    const uint64_t a = seq_.load(std::memory_order_acquire);
    if (a & 1u) {
      std::this_thread::yield();
      continue;
    }
    // acquire loads: none of them may be satisfied after the re-read of the sequence number below
    s.epoch = epoch_.load(std::memory_order_acquire);
    s.flag_epoch = flag_epoch_.load(std::memory_order_acquire);
    s.mask = mask_.load(std::memory_order_acquire);
    s.world = world_.load(std::memory_order_acquire);
    for (int r = 0; r < kMaxRanks; ++r) s.owner[r] = owner_[r].load(std::memory_order_acquire);
    if (seq_.load(std::memory_order_relaxed) == a) return s;
  }
}

inline uint64_t EpochTable::drop(int rank, int adopter) {
  std::lock_guard<std::mutex> lk(mu_);
  const uint32_t m = mask_.load(std::memory_order_relaxed);
  const int world = world_.load(std::memory_order_relaxed);
  if (rank < 0 || rank >= kMaxRanks || !(m & (1u << rank))) throw std::runtime_error("EpochTable::drop: not a member");
  const uint32_t left = m & ~(1u << rank);
  if (left == 0) throw std::runtime_error("EpochTable::drop: the last member cannot leave");
  if (adopter >= 0 && (adopter >= kMaxRanks || adopter == rank || !(m & (1u << adopter))))
    throw std::runtime_error("EpochTable::drop: adopter must be a remaining member");
  publish_begin();
  mask_.store(left, std::memory_order_relaxed);
  for (int s = 0; s < world; ++s) {                   // every shard the leaver held (its own and adopted ones)
    if (owner_[s].load(std::memory_order_relaxed) != rank) continue;
    int to = adopter;
    for (int d = 1; to < 0 && d <= world; ++d)        // default policy: the next live rank after the SHARD, cyclically
      if (left & (1u << ((s + d) % world))) to = (s + d) % world;
    owner_[s].store(to, std::memory_order_relaxed);
  }
  const uint64_t e = epoch_.load(std::memory_order_relaxed) + 1;
  epoch_.store(e, std::memory_order_relaxed);
  publish_end();
  return e;
}

inline uint64_t EpochTable::set_mask(uint32_t mask) {
  std::lock_guard<std::mutex> lk(mu_);
  const int world = world_.load(std::memory_order_relaxed);
  if (mask >> world) throw std::runtime_error("EpochTable::set_mask: rank outside the world");
  publish_begin();
  mask_.store(mask, std::memory_order_relaxed);
  const uint64_t e = epoch_.load(std::memory_order_relaxed) + 1;
  epoch_.store(e, std::memory_order_relaxed);
  publish_end();
  return e;
}

inline uint64_t EpochTable::set_owner(int shard, int rank) {
  std::lock_guard<std::mutex> lk(mu_);
  const int world = world_.load(std::memory_order_relaxed);
  if (shard < 0 || shard >= world || rank < 0 || rank >= world)
    throw std::runtime_error("EpochTable::set_owner: shard / rank outside the world");
  publish_begin();
  owner_[shard].store(rank, std::memory_order_relaxed);
  const uint64_t e = epoch_.load(std::memory_order_relaxed) + 1;
  epoch_.store(e, std::memory_order_relaxed);
  publish_end();
  return e;
}

inline uint64_t EpochTable::admit(int rank, bool take_back_shard) {
  std::lock_guard<std::mutex> lk(mu_);
  const uint32_t m = mask_.load(std::memory_order_relaxed);
  if (rank < 0 || rank >= world_.load(std::memory_order_relaxed) || (m & (1u << rank)))
    throw std::runtime_error("EpochTable::admit: rank is a member already or out of range");
  publish_begin();
  mask_.store(m | (1u << rank), std::memory_order_relaxed);
  if (take_back_shard) owner_[rank].store(rank, std::memory_order_relaxed);
  const uint64_t e = epoch_.load(std::memory_order_relaxed) + 1;
  epoch_.store(e, std::memory_order_relaxed);
  publish_end();
  return e;
}

inline uint64_t EpochTable::set_flag_epoch(uint64_t v) {
  std::lock_guard<std::mutex> lk(mu_);
  publish_begin();
  flag_epoch_.store(v, std::memory_order_relaxed);
  const uint64_t e = epoch_.load(std::memory_order_relaxed) + 1;
  epoch_.store(e, std::memory_order_relaxed);
  publish_end();
  return e;
}

// Re-bind under concurrent launches: `readers` threads take snapshots back to back (what a launch does) while one
// writer drops and re-admits ranks `iters` times.  A snapshot is consistent iff every shard's owner is a member and
// the epoch never runs backwards.  Returns the number of inconsistent snapshots (0 = pass).
inline uint64_t epoch_stress(int world, int readers, int iters) {
  EpochTable t(world);
  std::atomic<bool> stop{false};
  std::atomic<uint64_t> bad{0}, taken{0};
  std::vector<std::thread> th;
  for (int i = 0; i < readers; ++i) {
    th.emplace_back([&] {
      uint64_t last = 0;
      while (!stop.load(std::memory_order_relaxed)) {
        const EpochSnapshot s = t.snapshot();
        bool ok = s.epoch >= last;
        last = s.epoch;
        for (int r = 0; r < s.world; ++r) ok = ok && s.owner[r] >= 0 && (s.mask & (1u << s.owner[r]));
        // one drop + one admit per writer round: a consistent epoch has parity == (a rank is out)
        ok = ok && ((s.epoch & 1u) == (uint64_t)(__builtin_popcount(s.mask) != s.world));
        if (!ok) bad.fetch_add(1, std::memory_order_relaxed);
        taken.fetch_add(1, std::memory_order_relaxed);
      }
    });
  }
  for (int i = 0; i < iters; ++i) {
    const int victim = 1 + i % (world - 1 > 0 ? world - 1 : 1);
    if (world < 2) break;
    t.drop(victim, (victim + 1) % world == victim ? 0 : (victim + 1) % world);
    t.admit(victim, true);
  }
  while (taken.load(std::memory_order_relaxed) < (uint64_t)readers) std::this_thread::yield();
  stop.store(true);
  for (auto& x : th) x.join();
  return bad.load();
}

}  // namespace vmm
}  // namespace comm
}  // namespace hb
