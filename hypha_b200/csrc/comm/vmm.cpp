// Native substrate of the NVLS / peer-memory transport (SURVEY 7.1 `csrc/comm`):
//
//   * VMM allocations (cuMemCreate) exportable as POSIX file descriptors, import + map of a peer's allocation,
//   * NVSwitch multicast objects (cuMulticastCreate / AddDevice / BindMem) and their mapped view, which is what
//     `multimem.ld_reduce` / `multimem.st` in `kernels/outer_step.cu` address,
//   * descriptor passing between the rank processes over AF_UNIX sockets (SCM_RIGHTS),
//   * the membership-epoch table the fused outer step is launched against: readers (kernel launches) take a
//     consistent snapshot without a lock while a re-bind (drop / admit / owner change) publishes a new epoch.
//
// The driver API is resolved at run time (cudaGetDriverEntryPoint), so the extension imports on a machine without
// libcuda.so.1 and the CPU test-suite can exercise the socket and epoch code.  Python (`parallel/nvls.py`) only
// sequences the stages between control-plane exchanges.
//
// Replaces, at the transport level, the reference's tensor streams and parameter-server fan-in / fan-out
// (crates/network/src/stream_push.rs:26-90, crates/worker/src/executor/parameter_server.rs:101-283).
#include "vmm.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace hb {
namespace comm {
namespace vmm {

namespace {

// ---- driver entry points ---------------------------------------------------------------------------------------
struct Driver {
  bool ok = false;
  std::string why;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
};

template <class F>
bool resolve(const char* name, F*& fn, std::string& why) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess ||
      p == nullptr) {
    cudaGetLastError();
    why = std::string("driver entry point ") + name + " unavailable";
    return false;
  }
  fn = reinterpret_cast<F*>(p);
  return true;
}

Driver& drv() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    std::string& w = d.why;
    d.ok = resolve("cuDeviceGet", d.DeviceGet, w) && resolve("cuDeviceGetAttribute", d.DeviceGetAttribute, w) &&
           resolve("cuGetErrorString", d.GetErrorString, w) && resolve("cuMemCreate", d.MemCreate, w) &&
           resolve("cuMemRelease", d.MemRelease, w) &&
           resolve("cuMemExportToShareableHandle", d.MemExportToShareableHandle, w) &&
           resolve("cuMemImportFromShareableHandle", d.MemImportFromShareableHandle, w) &&
           resolve("cuMemAddressReserve", d.MemAddressReserve, w) && resolve("cuMemAddressFree", d.MemAddressFree, w) &&
           resolve("cuMemMap", d.MemMap, w) && resolve("cuMemUnmap", d.MemUnmap, w) &&
           resolve("cuMemSetAccess", d.MemSetAccess, w) &&
           resolve("cuMemGetAllocationGranularity", d.MemGetAllocationGranularity, w) &&
           resolve("cuMulticastCreate", d.MulticastCreate, w) && resolve("cuMulticastAddDevice", d.MulticastAddDevice, w) &&
           resolve("cuMulticastBindMem", d.MulticastBindMem, w) &&
           resolve("cuMulticastGetGranularity", d.MulticastGetGranularity, w);
  });
  return d;
}

Driver& need_driver() {
  Driver& d = drv();
  if (!d.ok) throw Unavailable(d.why);
  return d;
}

void ck(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* s = nullptr;
  Driver& d = drv();
  if (d.GetErrorString) d.GetErrorString(r, &s);
  throw Unavailable(std::string(what) + " failed: " + (s ? s : "CUresult ") + " (" + std::to_string((int)r) + ")");
}

constexpr CUmemAllocationHandleType kFd = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;

CUmemAllocationProp mem_prop(int device) {
  CUmemAllocationProp p;
  std::memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = kFd;
  return p;
}

CUmulticastObjectProp mc_prop(int world, size_t size) {
  CUmulticastObjectProp p;
  std::memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned)(world < 1 ? 1 : world);
  p.handleTypes = kFd;
  p.size = size;
  return p;
}

// ---- registry ------------------------------------------------------------------------------------------------------
struct Entry {
  CUmemGenericAllocationHandle h = 0;
  CUdeviceptr va = 0;
  size_t mapped = 0;        // bytes mapped at va (0: handle only)
  bool multicast = false;
};

std::mutex g_mu;
std::unordered_map<uint64_t, Entry> g_entries;
uint64_t g_next = 1;

uint64_t put(const Entry& e) {
  std::lock_guard<std::mutex> lk(g_mu);
  const uint64_t id = g_next++;
  g_entries[id] = e;
  return id;
}

Entry get(uint64_t id) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_entries.find(id);
  if (it == g_entries.end()) throw std::runtime_error("vmm: unknown handle id " + std::to_string(id));
  return it->second;
}

void set(uint64_t id, const Entry& e) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_entries[id] = e;
}

}  // namespace

bool driver_available() { return drv().ok; }

bool multicast_supported(int device) {
  Driver& d = drv();
  if (!d.ok) return false;
  CUdevice dev;
  int val = 0;
  if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return false;
  if (d.DeviceGetAttribute(&val, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return false;
  return val != 0;
}

size_t granularity(int device, int world, size_t bytes) {
  Driver& d = need_driver();
  size_t g_mem = 0, g_mc = 0;
  CUmemAllocationProp p = mem_prop(device);
  ck(d.MemGetAllocationGranularity(&g_mem, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  if (world > 0) {
    CUmulticastObjectProp mp = mc_prop(world, bytes ? bytes : 1);
    ck(d.MulticastGetGranularity(&g_mc, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
  }
  return g_mem > g_mc ? g_mem : g_mc;
}

AllocInfo create(int device, size_t bytes, int world) {
  Driver& d = need_driver();
  AllocInfo info;
  info.gran = granularity(device, world, bytes);
  info.size = (bytes + info.gran - 1) / info.gran * info.gran;
  CUmemAllocationProp p = mem_prop(device);
  Entry e;
  ck(d.MemCreate(&e.h, info.size, &p, 0), "cuMemCreate");
  info.id = put(e);
  return info;
}

int export_fd(uint64_t id) {
  Driver& d = need_driver();
  const Entry e = get(id);
  int fd = -1;
  ck(d.MemExportToShareableHandle(&fd, e.h, kFd, 0), "cuMemExportToShareableHandle");
  return fd;
}

uint64_t import_fd(int fd, bool multicast) {
  Driver& d = need_driver();
  Entry e;
  e.multicast = multicast;
  ck(d.MemImportFromShareableHandle(&e.h, (void*)(uintptr_t)fd, kFd), "cuMemImportFromShareableHandle");
  return put(e);
}

uintptr_t map(uint64_t id, size_t size, size_t gran, int device) {
  Driver& d = need_driver();
  Entry e = get(id);
  if (e.mapped) return (uintptr_t)e.va;
  CUdeviceptr va = 0;
  ck(d.MemAddressReserve(&va, size, gran, 0, 0), "cuMemAddressReserve");
  CUresult r = d.MemMap(va, size, 0, e.h, 0);
  if (r != CUDA_SUCCESS) {
    d.MemAddressFree(va, size);
    ck(r, "cuMemMap");
  }
  CUmemAccessDesc a;
  std::memset(&a, 0, sizeof(a));
  a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  a.location.id = device;
  a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.MemSetAccess(va, size, &a, 1);
  if (r != CUDA_SUCCESS) {
    d.MemUnmap(va, size);
    d.MemAddressFree(va, size);
    ck(r, "cuMemSetAccess");
  }
  e.va = va;
  e.mapped = size;
  set(id, e);
  return (uintptr_t)va;
}

void unmap(uint64_t id) {
  Driver& d = need_driver();
  Entry e = get(id);
  if (!e.mapped) return;
  d.MemUnmap(e.va, e.mapped);
  d.MemAddressFree(e.va, e.mapped);
  e.va = 0;
  e.mapped = 0;
  set(id, e);
}

void release(uint64_t id) noexcept {
  Driver& d = drv();
  Entry e;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_entries.find(id);
    if (it == g_entries.end()) return;
    e = it->second;
    g_entries.erase(it);
  }
  if (!d.ok) return;
  if (e.mapped) {
    d.MemUnmap(e.va, e.mapped);
    d.MemAddressFree(e.va, e.mapped);
  }
  d.MemRelease(e.h);
}

size_t live_handles() {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_entries.size();
}

uint64_t mc_create(int world, size_t size) {
  Driver& d = need_driver();
  CUmulticastObjectProp p = mc_prop(world, size);
  Entry e;
  e.multicast = true;
  ck(d.MulticastCreate(&e.h, &p), "cuMulticastCreate");
  return put(e);
}

void mc_add_device(uint64_t mc, int device) {
  Driver& d = need_driver();
  CUdevice dev;
  ck(d.DeviceGet(&dev, device), "cuDeviceGet");
  ck(d.MulticastAddDevice(get(mc).h, dev), "cuMulticastAddDevice");
}

void mc_bind(uint64_t mc, uint64_t mem, size_t size) {
  Driver& d = need_driver();
  ck(d.MulticastBindMem(get(mc).h, 0, get(mem).h, 0, size, 0), "cuMulticastBindMem");
}

bool usable(int device, int world) {
  // Rank-local pre-flight without any inter-rank traffic: can this process create an exportable allocation and a
  // multicast object for `world` devices and add its own device?  (fabric manager down, multicast not permitted
  // in the container, out of handles -> false, and every rank falls back to CUDA IPC before the collective set-up.)
  if (!multicast_supported(device)) return false;
  uint64_t mem = 0, mc = 0;
  int fd1 = -1, fd2 = -1;
  bool ok = false;
  try {
    AllocInfo a = create(device, 1, world);
    mem = a.id;
    fd1 = export_fd(mem);
    mc = mc_create(world, a.size);
    fd2 = export_fd(mc);
    mc_add_device(mc, device);
    ok = true;
  } catch (const std::exception&) {
    ok = false;
  }
  if (fd1 >= 0) ::close(fd1);
  if (fd2 >= 0) ::close(fd2);
  if (mc) release(mc);
  if (mem) release(mem);
  return ok;
}

// ---- descriptor passing over AF_UNIX ---------------------------------------------------------------------------------
namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool wait_fd(int fd, short events, double timeout_s) {
  struct pollfd p;
  p.fd = fd;
  p.events = events;
  p.revents = 0;
  const int ms = timeout_s < 0 ? -1 : (int)(timeout_s * 1000.0);
  int r;
  do {
    r = ::poll(&p, 1, ms);
  } while (r < 0 && errno == EINTR);
  return r > 0;
}

sockaddr_un make_addr(const std::string& path) {
  sockaddr_un a;
  std::memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  if (path.size() >= sizeof(a.sun_path)) throw std::runtime_error("unix socket path too long: " + path);
  std::memcpy(a.sun_path, path.data(), path.size());
  return a;
}

constexpr int kMaxFds = 16;

}  // namespace

int uds_listen(const std::string& path, int backlog) {
  ::unlink(path.c_str());
  const int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) throw std::runtime_error(std::string("socket: ") + std::strerror(errno));
  sockaddr_un a = make_addr(path);
  if (::bind(fd, (sockaddr*)&a, sizeof(a)) != 0 || ::listen(fd, backlog) != 0) {
    const std::string err = std::strerror(errno);
    ::close(fd);
    throw std::runtime_error("bind/listen " + path + ": " + err);
  }
  return fd;
}

void uds_close(int fd, const std::string& path) {
  if (fd >= 0) ::close(fd);
  if (!path.empty()) ::unlink(path.c_str());
}

void uds_send_fds(const std::string& path, int rank, const std::vector<int>& fds, double connect_timeout_s,
                  double io_timeout_s) {
  if (fds.empty() || (int)fds.size() > kMaxFds) throw std::runtime_error("uds_send_fds: 1..16 descriptors");
  const double deadline = now_s() + connect_timeout_s;
  int s = -1;
  for (;;) {
    s = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (s < 0) throw std::runtime_error(std::string("socket: ") + std::strerror(errno));
    sockaddr_un a = make_addr(path);
    if (::connect(s, (sockaddr*)&a, sizeof(a)) == 0) break;
    const int e = errno;
    ::close(s);
    if ((e != ENOENT && e != ECONNREFUSED) || now_s() > deadline)
      throw Unavailable("peer socket " + path + " never appeared: " + std::strerror(e));
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
  uint32_t tag = (uint32_t)rank;
  struct iovec iov;
  iov.iov_base = &tag;
  iov.iov_len = sizeof(tag);
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int) * kMaxFds)];
  std::memset(ctrl, 0, sizeof(ctrl));
  struct msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = CMSG_SPACE(sizeof(int) * fds.size());
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int) * fds.size());
  std::memcpy(CMSG_DATA(c), fds.data(), sizeof(int) * fds.size());
  bool ok = ::sendmsg(s, &msg, MSG_NOSIGNAL) == (ssize_t)sizeof(tag);
  char ack = 0;
  // the receiver acknowledges once it owns its duplicates; a receiver that died before accepting must not hang us
  ok = ok && wait_fd(s, POLLIN, io_timeout_s) && ::recv(s, &ack, 1, 0) == 1;
  ::close(s);
  if (!ok) throw Unavailable("peer at " + path + " did not take the memory handles");
}

std::map<int, std::vector<int>> uds_recv_fds(int server_fd, int count, int nfds, double timeout_s) {
  if (nfds < 1 || nfds > kMaxFds) throw std::runtime_error("uds_recv_fds: 1..16 descriptors");
  std::map<int, std::vector<int>> got;
  auto close_all = [&] {
    for (auto& kv : got)
      for (int fd : kv.second) ::close(fd);
    got.clear();
  };
  for (int i = 0; i < count; ++i) {
    if (!wait_fd(server_fd, POLLIN, timeout_s)) {
      close_all();
      throw Unavailable("timed out waiting for a peer's memory handle");
    }
    const int conn = ::accept4(server_fd, nullptr, nullptr, SOCK_CLOEXEC);
    if (conn < 0) {
      close_all();
      throw std::runtime_error(std::string("accept: ") + std::strerror(errno));
    }
    uint32_t tag = 0;
    struct iovec iov;
    iov.iov_base = &tag;
    iov.iov_len = sizeof(tag);
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int) * kMaxFds)];
    struct msghdr msg;
    std::memset(&msg, 0, sizeof(msg));
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    ssize_t n = -1;
    if (wait_fd(conn, POLLIN, timeout_s)) n = ::recvmsg(conn, &msg, MSG_CMSG_CLOEXEC);
    std::vector<int> fds;
    if (n == (ssize_t)sizeof(tag)) {
      for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c != nullptr; c = CMSG_NXTHDR(&msg, c)) {
        if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
          const size_t k = (c->cmsg_len - CMSG_LEN(0)) / sizeof(int);
          fds.resize(k);
          std::memcpy(fds.data(), CMSG_DATA(c), k * sizeof(int));
        }
      }
    }
    if (fds.empty() || (int)fds.size() > nfds) {
      for (int fd : fds) ::close(fd);
      ::close(conn);
      close_all();
      throw Unavailable("malformed memory-handle message from a peer");
    }
    got[(int)tag] = fds;
    const char ack = 'k';
    (void)!::send(conn, &ack, 1, MSG_NOSIGNAL);
    ::close(conn);
  }
  return got;
}

}  // namespace vmm
}  // namespace comm
}  // namespace hb
