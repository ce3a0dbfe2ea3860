// VMM / multicast / descriptor-passing / membership-epoch layer of the peer-memory transport (see vmm.cpp).
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace hb {
namespace comm {
namespace vmm {

// The NVLS path cannot be set up here (no driver, no multicast, a peer went away): callers fall back to CUDA IPC.
struct Unavailable : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct AllocInfo {
  uint64_t id = 0;     // registry id of the physical allocation
  size_t size = 0;     // bytes, rounded up to `gran`
  size_t gran = 0;     // max(allocation granularity, multicast granularity for the team size)
};

bool driver_available();
bool multicast_supported(int device);
bool usable(int device, int world);
size_t granularity(int device, int world, size_t bytes);

AllocInfo create(int device, size_t bytes, int world);               // cuMemCreate, exportable as a POSIX fd
int export_fd(uint64_t id);                                          // caller closes the descriptor
uint64_t import_fd(int fd, bool multicast = false);                  // handle only; `map` makes it addressable
uintptr_t map(uint64_t id, size_t size, size_t gran, int device);    // reserve VA + map + read/write access
void unmap(uint64_t id);
void release(uint64_t id) noexcept;                                  // unmap (if mapped) + cuMemRelease; idempotent
size_t live_handles();

uint64_t mc_create(int world, size_t size);
void mc_add_device(uint64_t mc, int device);
void mc_bind(uint64_t mc, uint64_t mem, size_t size);

int uds_listen(const std::string& path, int backlog);
void uds_close(int fd, const std::string& path);
void uds_send_fds(const std::string& path, int rank, const std::vector<int>& fds, double connect_timeout_s,
                  double io_timeout_s);
std::map<int, std::vector<int>> uds_recv_fds(int server_fd, int count, int nfds, double timeout_s);

}  // namespace vmm
}  // namespace comm
}  // namespace hb

#include "epoch_table.hpp"
