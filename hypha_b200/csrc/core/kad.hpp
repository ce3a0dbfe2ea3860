// Kademlia routing table of the fabric's DHT (native counterpart of crates/network/src/kad.rs:394-432,482-737).
//
//   ids        sha256(peer id) / sha256(key), 256 bits
//   distance   XOR, compared as big-endian integers
//   buckets    256, indexed by the length of the common prefix with our own id; K entries each in
//              least-recently-seen order; a full bucket keeps its live entries and gives way only for a stale one
//   closest    the n known servers nearest to a key -- where records are stored and where lookups start
//
// Address filtering (CIDR) is a policy of the caller (`transport/kad.py`); this table stores what it is given.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <list>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace hb {
namespace kad {

using Id = std::array<uint8_t, 32>;

// ---- SHA-256 (FIPS 180-4) ------------------------------------------------------------------------------------
inline Id sha256(const std::string& msg) {
  static const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
      0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
      0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
      0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
      0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
      0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
      0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
      0xc67178f2};
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  std::string m = msg;
  const uint64_t bits = (uint64_t)msg.size() * 8;
  m.push_back((char)0x80);
  while (m.size() % 64 != 56) m.push_back('\0');
  for (int i = 7; i >= 0; --i) m.push_back((char)((bits >> (8 * i)) & 0xff));
  auto rotr = [](uint32_t x, int n) { return (x >> n) | (x << (32 - n)); };
  for (size_t off = 0; off < m.size(); off += 64) {
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
      w[i] = ((uint32_t)(uint8_t)m[off + 4 * i] << 24) | ((uint32_t)(uint8_t)m[off + 4 * i + 1] << 16) |
             ((uint32_t)(uint8_t)m[off + 4 * i + 2] << 8) | (uint32_t)(uint8_t)m[off + 4 * i + 3];
    for (int i = 16; i < 64; ++i) {
      const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
      const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
      const uint32_t ch = (e & f) ^ (~e & g);
      const uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
      const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      const uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  Id out;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t)(h[i] >> (24 - 8 * j));
  return out;
}

inline Id xor_id(const Id& a, const Id& b) {
  Id d;
  for (int i = 0; i < 32; ++i) d[i] = a[i] ^ b[i];
  return d;
}

// number of leading zero bits of a distance (256 for the zero distance)
inline int leading_zeros(const Id& d) {
  for (int i = 0; i < 32; ++i) {
    if (d[i] == 0) continue;
    int n = 0;
    for (uint8_t v = d[i]; !(v & 0x80); v <<= 1) ++n;
    return 8 * i + n;
  }
  return 256;
}

struct Entry {
  std::string peer;
  Id id;
  std::vector<std::string> addrs;
  bool connected = true;
};

class RoutingTable {
 public:
  RoutingTable(const std::string& self_id, int k) : self_(self_id), self_hash_(sha256(self_id)), k_(k < 1 ? 1 : k) {}

  int bucket_index(const std::string& peer) const {
    const int lz = leading_zeros(xor_id(sha256(peer), self_hash_));
    return lz >= 256 ? 255 : lz;       // the bucket of the longest common prefix
  }

  // Learn or refresh a server.  `keep_addrs`: an empty address list leaves a known peer's addresses as they are.
  // Returns false when the bucket is full of live entries (the newcomer is ignored).
  bool add(const std::string& peer, const std::vector<std::string>& addrs, bool connected) {
    if (peer == self_) return false;
    const int b = bucket_index(peer);
    auto& bucket = buckets_[b];
    auto it = index_.find(peer);
    if (it != index_.end()) {
      Entry e = *it->second;
      bucket.erase(it->second);
      if (!addrs.empty()) e.addrs = addrs;
      e.connected = connected;
      bucket.push_back(std::move(e));
      index_[peer] = std::prev(bucket.end());
      return true;
    }
    if ((int)bucket.size() >= k_) {
      auto stale = std::find_if(bucket.begin(), bucket.end(), [](const Entry& e) { return !e.connected; });
      if (stale == bucket.end()) return false;
      index_.erase(stale->peer);
      bucket.erase(stale);
    }
    bucket.push_back(Entry{peer, sha256(peer), addrs, connected});
    index_[peer] = std::prev(bucket.end());
    return true;
  }

  void mark_disconnected(const std::string& peer) {
    auto it = index_.find(peer);
    if (it != index_.end()) it->second->connected = false;
  }

  void remove(const std::string& peer) {
    auto it = index_.find(peer);
    if (it == index_.end()) return;
    buckets_[bucket_index(peer)].erase(it->second);
    index_.erase(it);
  }

  bool contains(const std::string& peer) const { return index_.count(peer) != 0; }
  size_t size() const { return index_.size(); }

  std::vector<std::string> addrs(const std::string& peer) const {
    auto it = index_.find(peer);
    return it == index_.end() ? std::vector<std::string>{} : it->second->addrs;
  }

  std::vector<std::string> peers() const {              // bucket order, least recently seen first inside a bucket
    std::vector<std::string> out;
    out.reserve(index_.size());
    for (const auto& b : buckets_)
      for (const auto& e : b) out.push_back(e.peer);
    return out;
  }

  std::vector<std::string> bucket(int i) const {
    std::vector<std::string> out;
    if (i >= 0 && i < 256)
      for (const auto& e : buckets_[i]) out.push_back(e.peer);
    return out;
  }

  std::vector<std::string> closest(const std::string& key, int n, bool include_self) const {
    const Id t = sha256(key);
    std::vector<std::pair<Id, const std::string*>> cand;
    cand.reserve(index_.size() + 1);
    for (const auto& b : buckets_)
      for (const auto& e : b) cand.emplace_back(xor_id(e.id, t), &e.peer);
    if (include_self) cand.emplace_back(xor_id(self_hash_, t), &self_);
    const size_t take = std::min<size_t>(cand.size(), (size_t)std::max(n, 0));
    std::partial_sort(cand.begin(), cand.begin() + take, cand.end(),
                      [](const auto& a, const auto& b) { return a.first < b.first; });   // big-endian byte order
    std::vector<std::string> out;
    out.reserve(take);
    for (size_t i = 0; i < take; ++i) out.push_back(*cand[i].second);
    return out;
  }

 private:
  std::string self_;
  Id self_hash_;
  int k_;
  std::array<std::list<Entry>, 256> buckets_;
  std::unordered_map<std::string, std::list<Entry>::iterator> index_;
};

}  // namespace kad
}  // namespace hb
