// Peer-memory transport: symmetric device buffers shared between the per-GPU
// worker processes of one NVSwitch box.
//
// This is the B200-native replacement for the reference's tensor streams
// (`/hypha-tensor-stream/push`, crates/network/src/stream_push.rs:16-90 and the
// worker connector's send/receive, crates/worker/src/connector/mod.rs:318-433):
// instead of serialising safetensors files through TCP+TLS, every rank exports
// its buffers once at job start and peers map them; kernels then move data with
// plain ld/st over NVLink.
//
// Two mapping back-ends:
//   * legacy CUDA IPC   (cudaIpcGetMemHandle / cudaIpcOpenMemHandle)
//   * same-process peer access (cudaDeviceEnablePeerAccess), used by tests that
//     drive several GPUs from one process.
#include "peer_mem.h"

#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace hb {
namespace comm {

namespace {
std::mutex g_mu;
std::unordered_map<uintptr_t, size_t> g_allocs;   // ptr -> bytes (owned)
std::unordered_map<uintptr_t, int> g_imports;     // ptr -> refcount (imported)
unsigned long long g_bytes_alloc = 0;

void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
  }
}
}  // namespace

uintptr_t sym_alloc(size_t bytes, bool zero) {
  void* p = nullptr;
  check(cudaMalloc(&p, bytes), "cudaMalloc(symmetric)");
  if (zero) check(cudaMemset(p, 0, bytes), "cudaMemset(symmetric)");
  std::lock_guard<std::mutex> lk(g_mu);
  g_allocs[(uintptr_t)p] = bytes;
  g_bytes_alloc += bytes;
  return (uintptr_t)p;
}

void sym_free(uintptr_t ptr) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_allocs.find(ptr);
    if (it == g_allocs.end()) throw std::runtime_error("sym_free: unknown pointer");
    g_bytes_alloc -= it->second;
    g_allocs.erase(it);
  }
  check(cudaFree((void*)ptr), "cudaFree(symmetric)");
}

std::string ipc_export(uintptr_t ptr) {
  cudaIpcMemHandle_t h;
  check(cudaIpcGetMemHandle(&h, (void*)ptr), "cudaIpcGetMemHandle");
  return std::string(reinterpret_cast<const char*>(&h), sizeof(h));
}

uintptr_t ipc_open(const std::string& handle) {
  if (handle.size() != sizeof(cudaIpcMemHandle_t)) {
    throw std::runtime_error("ipc_open: bad handle size");
  }
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  std::lock_guard<std::mutex> lk(g_mu);
  g_imports[(uintptr_t)p] += 1;
  return (uintptr_t)p;
}

void ipc_close(uintptr_t ptr) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_imports.find(ptr);
    if (it == g_imports.end()) throw std::runtime_error("ipc_close: unknown pointer");
    if (--it->second > 0) return;
    g_imports.erase(it);
  }
  check(cudaIpcCloseMemHandle((void*)ptr), "cudaIpcCloseMemHandle");
}

bool enable_peer_access(int device, int peer) {
  int can = 0;
  check(cudaDeviceCanAccessPeer(&can, device, peer), "cudaDeviceCanAccessPeer");
  if (!can) return false;
  int prev = 0;
  cudaGetDevice(&prev);
  check(cudaSetDevice(device), "cudaSetDevice");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    e = cudaSuccess;
  }
  cudaSetDevice(prev);
  check(e, "cudaDeviceEnablePeerAccess");
  return true;
}

unsigned long long bytes_allocated() {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_bytes_alloc;
}

void memcpy_d2d_async(uintptr_t dst, uintptr_t src, size_t bytes, uintptr_t stream) {
  check(cudaMemcpyAsync((void*)dst, (const void*)src, bytes, cudaMemcpyDeviceToDevice,
                        (cudaStream_t)stream),
        "cudaMemcpyAsync(d2d)");
}

void memset_async(uintptr_t dst, int value, size_t bytes, uintptr_t stream) {
  check(cudaMemsetAsync((void*)dst, value, bytes, (cudaStream_t)stream), "cudaMemsetAsync");
}

}  // namespace comm
}  // namespace hb
