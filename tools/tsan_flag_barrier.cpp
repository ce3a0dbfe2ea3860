// ThreadSanitizer driver for the flag-barrier protocol model (tools/sanitize.sh tsan).
#include <cstdio>

#include "core/flag_barrier_model.hpp"

int main() {
  using hypha::FlagBarrierModel;
  int bad = 0;
  for (int world : {2, 4, 8}) {
    FlagBarrierModel m(world, (1u << world) - 1u);
    auto r = m.stress(/*epochs=*/200, /*stages=*/10, /*seed=*/world);
    std::printf("world %d: passed %llu violations %llu timeouts %u -> %s\n", world,
                (unsigned long long)r.barriers_passed, (unsigned long long)r.violations, r.timeouts, r.ok ? "ok" : "FAIL");
    bad += !r.ok;
  }
  {  // epoch counter close to wrap-around
    FlagBarrierModel m(4, 0xF);
    auto r = m.stress(64, 18, 5, -1, 2000, /*first_epoch=*/(0xFFFFFFFFu / 32u) - 20u);
    std::printf("wrap: passed %llu violations %llu -> %s\n", (unsigned long long)r.barriers_passed,
                (unsigned long long)r.violations, r.ok ? "ok" : "FAIL");
    bad += !r.ok;
  }
  {  // a dead member is reported by every survivor instead of hanging them
    FlagBarrierModel m(4, 0xF);
    auto r = m.stress(3, 4, 9, /*dead_rank=*/2, /*timeout_ms=*/100);
    std::printf("dead rank 2: timeouts %u missing mask 0x%x -> %s\n", r.timeouts, r.missing_union,
                (r.ok && r.missing_union == 0x4) ? "ok" : "FAIL");
    bad += !(r.ok && r.missing_union == 0x4);
  }
  return bad;
}
