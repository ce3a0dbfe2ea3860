// ThreadSanitizer driver for the membership-epoch table (csrc/comm/epoch_table.hpp): reader threads take snapshots
// back to back -- what every fused-outer-step launch does -- while one writer drops and re-admits ranks, changes
// owners and resyncs the flag epoch.  Fails (exit 1) on an inconsistent snapshot; TSAN fails it on a data race.
#include <cstdio>

#include "comm/epoch_table.hpp"

int main() {
  using namespace hb::comm::vmm;
  for (int world : {2, 4, 8}) {
    const uint64_t bad = epoch_stress(world, 4, 20000);
    std::printf("epoch_stress(world=%d, readers=4, iters=20000): %llu inconsistent snapshots\n", world,
                (unsigned long long)bad);
    if (bad != 0) return 1;
  }
  // mixed writers: membership through drop/admit on one thread, flag epochs on another, readers throughout
  EpochTable t(8);
  std::atomic<bool> stop{false};
  std::atomic<uint64_t> bad{0};
  std::thread flags([&] {
    for (uint64_t i = 1; i <= 20000; ++i) t.set_flag_epoch(i);
  });
  std::thread members([&] {
    for (int i = 0; i < 20000; ++i) {
      const int v = 1 + i % 7;
      t.drop(v);
      t.admit(v, true);
    }
  });
  std::vector<std::thread> readers;
  for (int r = 0; r < 4; ++r)
    readers.emplace_back([&] {
      uint64_t last_epoch = 0, last_flag = 0;
      while (!stop.load(std::memory_order_relaxed)) {
        const EpochSnapshot s = t.snapshot();
        bool ok = s.epoch >= last_epoch && s.flag_epoch >= last_flag;
        last_epoch = s.epoch;
        last_flag = s.flag_epoch;
        for (int k = 0; k < s.world; ++k) ok = ok && s.owner[k] >= 0 && (s.mask & (1u << s.owner[k]));
        if (!ok) bad.fetch_add(1);
      }
    });
  flags.join();
  members.join();
  stop.store(true);
  for (auto& x : readers) x.join();
  std::printf("mixed writers: %llu inconsistent snapshots, final epoch %llu\n", (unsigned long long)bad.load(),
              (unsigned long long)t.snapshot().epoch);
  return bad.load() == 0 ? 0 : 1;
}
