// Peer-memory transport API (see peer_mem.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>

namespace hb {
namespace comm {
uintptr_t sym_alloc(size_t bytes, bool zero);
void sym_free(uintptr_t ptr);
std::string ipc_export(uintptr_t ptr);
uintptr_t ipc_open(const std::string& handle);
void ipc_close(uintptr_t ptr);
bool enable_peer_access(int device, int peer);
unsigned long long bytes_allocated();
void memcpy_d2d_async(uintptr_t dst, uintptr_t src, size_t bytes, uintptr_t stream);
void memset_async(uintptr_t dst, int value, size_t bytes, uintptr_t stream);
}  // namespace comm
}  // namespace hb
