// hypha_b200 control-plane core (C++17, header-only, no CUDA).
//
// Native equivalents of the reference's Rust building blocks, re-designed around
// an explicit clock so every component is deterministic under test:
//   Resources / WeightedResourceEvaluator   (crates/resources/src/lib.rs:10-199)
//   Ledger                                  (crates/leases/src/lib.rs:20-131)
//   StaticResourceManager                   (crates/worker/src/resources.rs:38-92)
//   ResourceLeaseManager                    (crates/worker/src/lease_manager.rs:71-185)
//   OfferAggregator (greedy, diversity)     (crates/scheduler/src/allocator.rs:169-418)
//   simulate_sync_point                     (crates/scheduler/src/simulation.rs:16-68)
//   RunningMean                             (crates/scheduler/src/statistics.rs:1-44)
//   WorkerTracker / ProgressTracker         (crates/scheduler/src/tracker/{worker,progress}.rs)
//   BatchScheduler::handle                  (crates/scheduler/src/scheduling/batch_scheduler.rs:54-163)
//   SliceTracker                            (crates/scheduler/src/tracker/slice.rs:35-114)
//   BatchWindow                             (crates/network/src/utils.rs:50-141)
// Times are doubles in seconds (wall clock unless the caller injects `now`);
// tracker times are integer milliseconds like the reference's statistics.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <limits>
#include <map>
#include <mutex>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace hypha {

inline double wall_now() {
  using namespace std::chrono;
  return duration<double>(system_clock::now().time_since_epoch()).count();
}
inline double resolve_now(double now) { return now >= 0.0 ? now : wall_now(); }

inline std::string uuid4() {
  static thread_local std::mt19937_64 rng{std::random_device{}()};
  uint64_t a = rng(), b = rng();
  a = (a & 0xFFFFFFFFFFFF0FFFull) | 0x0000000000004000ull;  // version 4
  b = (b & 0x3FFFFFFFFFFFFFFFull) | 0x8000000000000000ull;  // variant 1
  static const char* hex = "0123456789abcdef";
  std::string s(36, '-');
  int pos = 0;
  auto put = [&](uint64_t v, int nibbles) {
    for (int i = nibbles - 1; i >= 0; --i) s[pos++] = hex[(v >> (4 * i)) & 0xF];
  };
  put(a >> 32, 8); ++pos; put((a >> 16) & 0xFFFF, 4); ++pos; put(a & 0xFFFF, 4); ++pos;
  put(b >> 48, 4); ++pos; put(b & 0xFFFFFFFFFFFFull, 12);
  return s;
}

// ------------------------------------------------------------------------------------
// Resources: 4-vector with a component-wise PARTIAL order.
// ------------------------------------------------------------------------------------
enum class Ordering { Less = -1, Equal = 0, Greater = 1, Incomparable = 2 };

struct Resources {
  double gpu = 0, cpu = 0, storage = 0, memory = 0;

  Resources operator+(const Resources& o) const { return {gpu + o.gpu, cpu + o.cpu, storage + o.storage, memory + o.memory}; }
  Resources operator-(const Resources& o) const { return {gpu - o.gpu, cpu - o.cpu, storage - o.storage, memory - o.memory}; }
  Resources& operator+=(const Resources& o) { return *this = *this + o; }
  Resources& operator-=(const Resources& o) { return *this = *this - o; }
  bool operator==(const Resources& o) const { return gpu == o.gpu && cpu == o.cpu && storage == o.storage && memory == o.memory; }

  Ordering compare(const Resources& o) const {
    const double a[4] = {cpu, memory, gpu, storage};
    const double b[4] = {o.cpu, o.memory, o.gpu, o.storage};
    Ordering acc = Ordering::Equal;
    for (int i = 0; i < 4; ++i) {
      if (a[i] != a[i] || b[i] != b[i]) return Ordering::Incomparable;  // NaN
      Ordering f = a[i] < b[i] ? Ordering::Less : (a[i] > b[i] ? Ordering::Greater : Ordering::Equal);
      if (f == Ordering::Equal || f == acc) continue;
      if (acc == Ordering::Equal) { acc = f; continue; }
      return Ordering::Incomparable;
    }
    return acc;
  }
  bool ge(const Resources& o) const { auto c = compare(o); return c == Ordering::Greater || c == Ordering::Equal; }
  bool gt(const Resources& o) const { return compare(o) == Ordering::Greater; }
  bool le(const Resources& o) const { auto c = compare(o); return c == Ordering::Less || c == Ordering::Equal; }
  bool lt(const Resources& o) const { return compare(o) == Ordering::Less; }
};

struct WeightedResourceEvaluator {
  double gpu = 25.0, cpu = 1.0, memory = 0.1, storage = 0.01;
  double evaluate(double price, const Resources& r) const {
    const double units = gpu * r.gpu + cpu * r.cpu + memory * r.memory + storage * r.storage;
    return units > 0.0 ? price / units : 0.0;
  }
};

// ------------------------------------------------------------------------------------
// Ledger: lease table with wall-clock expiry; renew sets timeout = now + d.
// ------------------------------------------------------------------------------------
template <class T>
struct Lease {
  std::string id;
  T leasable;
  double timeout = 0;
  bool is_expired(double now = -1) const { return resolve_now(now) >= timeout; }
};

struct LeaseNotFound : std::runtime_error {
  explicit LeaseNotFound(const std::string& id) : std::runtime_error("Lease not found: " + id) {}
};

template <class T>
class Ledger {
 public:
  Lease<T> insert(T leasable, double duration, double now = -1) {
    std::lock_guard<std::mutex> lk(mu_);
    Lease<T> l{uuid4(), std::move(leasable), resolve_now(now) + duration};
    leases_[l.id] = l;
    return l;
  }
  Lease<T> get(const std::string& id) const {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = leases_.find(id);
    if (it == leases_.end()) throw LeaseNotFound(id);
    return it->second;
  }
  Lease<T> remove(const std::string& id) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = leases_.find(id);
    if (it == leases_.end()) throw LeaseNotFound(id);
    Lease<T> l = it->second;
    leases_.erase(it);
    return l;
  }
  Lease<T> renew(const std::string& id, double duration, double now = -1) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = leases_.find(id);
    if (it == leases_.end()) throw LeaseNotFound(id);
    it->second.timeout = resolve_now(now) + duration;
    return it->second;
  }
  std::vector<Lease<T>> list() const {
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<Lease<T>> out;
    for (auto& kv : leases_) out.push_back(kv.second);
    return out;
  }
  std::vector<Lease<T>> list_expired(double now = -1) const {
    const double t = resolve_now(now);
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<Lease<T>> out;
    for (auto& kv : leases_)
      if (kv.second.timeout <= t) out.push_back(kv.second);
    return out;
  }
  size_t size() const { std::lock_guard<std::mutex> lk(mu_); return leases_.size(); }

 private:
  mutable std::mutex mu_;
  std::map<std::string, Lease<T>> leases_;
};

// ------------------------------------------------------------------------------------
// Capacity accounting
// ------------------------------------------------------------------------------------
struct InsufficientResources : std::runtime_error {
  InsufficientResources() : std::runtime_error("Insufficient resources") {}
};
struct UnknownReservation : std::runtime_error {
  UnknownReservation() : std::runtime_error("Unknown resource reservation") {}
};

class StaticResourceManager {
 public:
  explicit StaticResourceManager(Resources capacity) : capacity_(capacity) {}
  Resources capacity() const { return capacity_; }
  Resources available() const {
    std::lock_guard<std::mutex> lk(mu_);
    return available_locked();
  }
  std::string reserve(const Resources& r) {
    std::lock_guard<std::mutex> lk(mu_);
    if (!available_locked().ge(r)) throw InsufficientResources();
    std::string id = uuid4();
    reservations_[id] = r;
    return id;
  }
  Resources release(const std::string& id) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = reservations_.find(id);
    if (it == reservations_.end()) throw UnknownReservation();
    Resources r = it->second;
    reservations_.erase(it);
    return r;
  }

 private:
  Resources available_locked() const {
    Resources used;
    for (auto& kv : reservations_) used += kv.second;
    return capacity_ - used;
  }
  mutable std::mutex mu_;
  Resources capacity_;
  std::map<std::string, Resources> reservations_;
};

struct ResourceLease {
  std::string peer_id;
  std::string reservation;
  Resources resources;
};

class ResourceLeaseManager {
 public:
  explicit ResourceLeaseManager(Resources capacity) : rm_(capacity) {}
  StaticResourceManager& resources() { return rm_; }

  Lease<ResourceLease> request(const std::string& peer, const Resources& r, double duration, double now = -1) {
    std::string reservation = rm_.reserve(r);
    return ledger_.insert(ResourceLease{peer, reservation, r}, duration, now);
  }
  Lease<ResourceLease> get(const std::string& id) const { return ledger_.get(id); }
  // Reference quirk Q7 fixed: prefer the peer's live lease with the latest timeout.
  Lease<ResourceLease> get_by_peer(const std::string& peer, double now = -1) const {
    const double t = resolve_now(now);
    std::optional<Lease<ResourceLease>> best;
    for (auto& l : ledger_.list()) {
      if (l.leasable.peer_id != peer) continue;
      const bool live = l.timeout > t;
      if (!best || (live && !(best->timeout > t)) || (live == (best->timeout > t) && l.timeout > best->timeout)) best = l;
    }
    if (!best) throw LeaseNotFound("<peer " + peer + ">");
    return *best;
  }
  Lease<ResourceLease> remove(const std::string& id) {
    auto l = ledger_.remove(id);
    rm_.release(l.leasable.reservation);
    return l;
  }
  Lease<ResourceLease> renew(const std::string& id, double duration, double now = -1) {
    return ledger_.renew(id, duration, now);
  }
  // Drop every expired lease, releasing its reservation; returns what was removed.
  std::vector<Lease<ResourceLease>> proceed(double now = -1) {
    std::vector<Lease<ResourceLease>> removed;
    for (auto& l : ledger_.list_expired(now)) {
      try { removed.push_back(remove(l.id)); } catch (const LeaseNotFound&) {}
    }
    return removed;
  }
  std::vector<Lease<ResourceLease>> list() const { return ledger_.list(); }

 private:
  StaticResourceManager rm_;
  Ledger<ResourceLease> ledger_;
};

// ------------------------------------------------------------------------------------
// Greedy offer aggregation (scheduler side of the allocation protocol)
// ------------------------------------------------------------------------------------
struct Offer {
  std::string peer_id;
  std::string lease_id;     // worker-side lease id carried by the offer
  std::string request_id;
  double price = 0;
  Resources resources;
  double timeout = 0;       // absolute expiry of the tentative lease
  double score = 0;
};

enum class OfferVerdict { Accepted = 0, RejectedPrice = 1, RejectedWorse = 2, Full = 3, Closed = 4 };

class OfferAggregator {
 public:
  OfferAggregator(double now, double deadline_s, size_t desired, bool diversity, double upper_price,
                  WeightedResourceEvaluator ev = {}, long max_offers = -1)
      : capacity_(std::max<size_t>(desired, 1)), diversity_(diversity), upper_price_(upper_price), ev_(ev),
        hard_deadline_(now + deadline_s), deadline_(now + deadline_s), max_offers_(max_offers) {}

  OfferVerdict push(Offer o, double now) {
    if (closed_) return OfferVerdict::Closed;
    ++received_;
    OfferVerdict v;
    if (o.price > upper_price_) {
      v = OfferVerdict::RejectedPrice;
    } else {
      o.score = ev_.evaluate(o.price, o.resources);
      v = try_insert(o);
      if (v == OfferVerdict::Accepted) {
        double d = hard_deadline_;
        for (auto& c : offers_) {
          if (c.timeout >= now) d = std::min(d, now + std::max(0.0, c.timeout - now - 0.1));
        }
        deadline_ = d;
        if (offers_.size() >= capacity_) closed_ = true;  // early return
      }
    }
    if (max_offers_ >= 0 && received_ >= max_offers_) closed_ = true;
    return v;
  }
  bool done(double now) const { return closed_ || now >= deadline_; }
  double deadline() const { return deadline_; }
  const std::vector<Offer>& offers() const { return offers_; }
  std::vector<Offer> finish() { closed_ = true; return offers_; }
  long received() const { return received_; }

 private:
  OfferVerdict try_insert(const Offer& c) {
    if (diversity_) {
      for (auto& e : offers_) {
        if (e.peer_id == c.peer_id) {
          if (c.score < e.score) { e = c; sort(); return OfferVerdict::Accepted; }
          return OfferVerdict::RejectedWorse;
        }
      }
    }
    if (offers_.size() < capacity_) { offers_.push_back(c); sort(); return OfferVerdict::Accepted; }
    if (!offers_.empty() && c.score < offers_.back().score) { offers_.back() = c; sort(); return OfferVerdict::Accepted; }
    return OfferVerdict::Full;
  }
  void sort() { std::stable_sort(offers_.begin(), offers_.end(), [](const Offer& a, const Offer& b) { return a.score < b.score; }); }

  size_t capacity_;
  bool diversity_;
  double upper_price_;
  WeightedResourceEvaluator ev_;
  double hard_deadline_, deadline_;
  long max_offers_;
  long received_ = 0;
  bool closed_ = false;
  std::vector<Offer> offers_;
};

// Worker-side request ranking (arbiter.rs:328-381): filter then sort by score, best first.
struct WorkerRequestView {
  size_t index = 0;        // position in the caller's batch
  double bid = 0;
  Resources resources;
  bool executors_supported = true;
  double score = 0;
};
inline std::vector<WorkerRequestView> rank_requests(std::vector<WorkerRequestView> reqs, const Resources& capacity,
                                                    double floor, const WeightedResourceEvaluator& ev) {
  std::vector<WorkerRequestView> out;
  for (auto& r : reqs) {
    if (!r.executors_supported) continue;
    if (r.bid < floor) continue;
    if (r.resources.gt(capacity)) continue;  // partial order: only strictly-greater requests are dropped here
    r.score = ev.evaluate(r.bid, r.resources);
    out.push_back(r);
  }
  std::stable_sort(out.begin(), out.end(), [](const WorkerRequestView& a, const WorkerRequestView& b) { return a.score > b.score; });
  return out;
}

// ------------------------------------------------------------------------------------
// Synchronisation-point simulation + statistics
// ------------------------------------------------------------------------------------
struct Projection {
  uint64_t time = 0;
  uint32_t remaining = 0;
  std::vector<uint32_t> updates;
  bool capped = false;
};

inline Projection simulate_sync_point(const std::vector<uint64_t>& last_update, const std::vector<uint32_t>& batch_sizes,
                                      const std::vector<uint64_t>& mean_latency, uint32_t data_points_left,
                                      uint64_t time_cap, uint32_t steps_cap) {
  const size_t n = batch_sizes.size();
  if (last_update.size() != n || mean_latency.size() != n) throw std::invalid_argument("simulate: length mismatch");
  auto sat_add = [](uint64_t a, uint64_t b) { return a > std::numeric_limits<uint64_t>::max() - b ? std::numeric_limits<uint64_t>::max() : a + b; };
  Projection p;
  p.updates.assign(n, 0);
  p.remaining = data_points_left;
  std::vector<uint64_t> next(n);
  for (size_t i = 0; i < n; ++i) next[i] = sat_add(last_update[i], mean_latency[i]);
  while (p.remaining > 0 && n > 0) {
    const uint64_t t = *std::min_element(next.begin(), next.end());
    if (t >= time_cap) { p.capped = true; break; }
    p.time = t;
    bool hit_steps = false;
    for (size_t i = 0; i < n; ++i) {
      if (next[i] != t) continue;  // every worker finishing at the same instant is processed together
      p.remaining = p.remaining > batch_sizes[i] ? p.remaining - batch_sizes[i] : 0;
      if (p.updates[i] != std::numeric_limits<uint32_t>::max()) ++p.updates[i];
      if (p.updates[i] >= steps_cap) hit_steps = true;
      next[i] = sat_add(next[i], mean_latency[i]);
    }
    if (hit_steps) { p.capped = true; break; }
  }
  return p;
}

class RunningMean {
 public:
  void update(uint64_t x) {
    if (samples_ == 0) { mean_ = x; samples_ = 1; return; }
    ++samples_;
    const int64_t m = (int64_t)mean_;
    mean_ = (uint64_t)(m + ((int64_t)x - m) / (int64_t)samples_);
  }
  uint64_t value() const { return mean_; }
  uint64_t samples() const { return samples_; }

 private:
  uint64_t mean_ = std::numeric_limits<uint64_t>::max();
  uint64_t samples_ = 0;
};

enum class WorkerState { Training = 0, UpdateScheduled = 1, Updating = 2, Done = 3 };

struct PeerNotFound : std::runtime_error {
  PeerNotFound() : std::runtime_error("Peer not found") {}
};

class WorkerTracker {
 public:
  size_t position(const std::string& peer) const {
    for (size_t i = 0; i < peers_.size(); ++i) if (peers_[i] == peer) return i;
    throw PeerNotFound();
  }
  bool contains(const std::string& peer) const { return std::find(peers_.begin(), peers_.end(), peer) != peers_.end(); }
  void add_worker(const std::string& peer, uint32_t batch_size) {
    peers_.push_back(peer); batch_.push_back(batch_size); last_.push_back(0);
    state_.push_back(WorkerState::Training); stats_.emplace_back();
  }
  void remove_worker(const std::string& peer) {
    const size_t i = position(peer);
    peers_.erase(peers_.begin() + i); batch_.erase(batch_.begin() + i); last_.erase(last_.begin() + i);
    state_.erase(state_.begin() + i); stats_.erase(stats_.begin() + i);
  }
  void update(const std::string& peer, uint64_t now_ms) {
    const size_t i = position(peer);
    stats_[i].update(now_ms >= last_[i] ? now_ms - last_[i] : 0);
    last_[i] = now_ms;
  }
  const std::vector<uint32_t>& batch_sizes() const { return batch_; }
  const std::vector<uint64_t>& last_updates() const { return last_; }
  std::vector<uint64_t> estimates() const { std::vector<uint64_t> e; for (auto& s : stats_) e.push_back(s.value()); return e; }
  WorkerState state(const std::string& peer) const { return state_[position(peer)]; }
  void set_state(const std::string& peer, WorkerState s) { state_[position(peer)] = s; }
  const std::vector<std::string>& workers() const { return peers_; }
  const std::vector<WorkerState>& states() const { return state_; }
  void new_round() { std::fill(last_.begin(), last_.end(), 0); std::fill(state_.begin(), state_.end(), WorkerState::Training); }
  void done() { std::fill(state_.begin(), state_.end(), WorkerState::Done); }
  size_t size() const { return peers_.size(); }

 private:
  std::vector<std::string> peers_;
  std::vector<uint32_t> batch_;
  std::vector<uint64_t> last_;
  std::vector<WorkerState> state_;
  std::vector<RunningMean> stats_;
};

class ProgressTracker {
 public:
  ProgressTracker(std::string parameter_server, uint32_t update_target, uint32_t update_epochs, uint64_t now_ms = 0)
      : ps_(std::move(parameter_server)), counter_(update_target), target_(update_target), epochs_(update_epochs), round_start_ms_(now_ms) {}
  void update(const std::string& peer, uint32_t count, uint64_t now_ms) {
    counter_ = counter_ > count ? counter_ - count : 0;
    workers.update(peer, now_ms >= round_start_ms_ ? now_ms - round_start_ms_ : 0);
  }
  void next_round(uint64_t now_ms) { counter_ = target_; round_start_ms_ = now_ms; ++round_; workers.new_round(); }
  uint32_t count() const { return counter_; }
  uint32_t round() const { return round_; }
  bool training_finished() const { return round_ == epochs_; }
  const std::string& parameter_server() const { return ps_; }
  void set_parameter_server(std::string ps) { ps_ = std::move(ps); }
  WorkerTracker workers;

 private:
  std::string ps_;
  uint32_t counter_, target_, epochs_, round_ = 0;
  uint64_t round_start_ms_;
};

// Progress state machine (batch_scheduler.rs:54-163).
enum class ProgressKind { Status = 0, Metrics = 1, Update = 2, Updated = 3, UpdateReceived = 4 };
enum class ResponseKind { Ok = 0, Continue = 1, ScheduleUpdate = 2, Done = 3, Error = 4 };
struct ProgressResponse {
  ResponseKind kind = ResponseKind::Ok;
  uint32_t counter = 0;
  bool all_done = false;  // every worker reported UpdateReceived after the final round
};

class BatchScheduler {
 public:
  BatchScheduler(ProgressTracker tracker, uint64_t time_cap_ms = 10000, uint32_t update_cap = 3)
      : tracker(std::move(tracker)), time_cap_(time_cap_ms), update_cap_(update_cap) {}

  ProgressResponse handle(const std::string& peer, ProgressKind kind, uint32_t batch_size, uint64_t now_ms) {
    ProgressResponse r;
    switch (kind) {
      case ProgressKind::Metrics:
        r.kind = ResponseKind::Ok;
        return r;
      case ProgressKind::Status: {
        if (!tracker.workers.contains(peer)) { r.kind = ResponseKind::Error; return r; }
        tracker.update(peer, batch_size, now_ms);  // counted in every state, like the reference
        const WorkerState st = tracker.workers.state(peer);
        if (st == WorkerState::Training) {
          const Projection p = simulate_sync_point(tracker.workers.last_updates(), tracker.workers.batch_sizes(),
                                                   tracker.workers.estimates(), tracker.count(), time_cap_, update_cap_);
          if (p.remaining == 0 && !p.capped) {
            tracker.workers.set_state(peer, WorkerState::UpdateScheduled);
            r.kind = ResponseKind::ScheduleUpdate;
            r.counter = p.updates[tracker.workers.position(peer)];
          } else {
            r.kind = ResponseKind::Continue;
          }
        } else if (st == WorkerState::UpdateScheduled) {
          r.kind = ResponseKind::Continue;
        } else {
          r.kind = ResponseKind::Error;
        }
        return r;
      }
      case ProgressKind::Update:
        if (!tracker.workers.contains(peer)) { r.kind = ResponseKind::Error; return r; }
        tracker.workers.set_state(peer, WorkerState::Updating);
        round_advanced_ = false;
        r.kind = ResponseKind::Ok;
        return r;
      case ProgressKind::Updated:
        // Idempotent within one synchronisation: in collective mode every rank may report `updated`
        // (no single aggregate worker whose death would stall the round); only the first one after
        // the workers' `update` messages advances the round.
        if (!round_advanced_) {
          tracker.next_round(now_ms);
          round_advanced_ = true;
        }
        r.kind = tracker.training_finished() ? ResponseKind::Done : ResponseKind::Ok;
        return r;
      case ProgressKind::UpdateReceived:
        if (tracker.training_finished()) {
          if (tracker.workers.contains(peer)) tracker.workers.set_state(peer, WorkerState::Done);
          r.kind = ResponseKind::Done;
          r.all_done = true;
          for (auto s : tracker.workers.states()) if (s != WorkerState::Done) r.all_done = false;
        } else {
          if (tracker.workers.contains(peer)) tracker.workers.set_state(peer, WorkerState::Training);
          r.kind = ResponseKind::Continue;
        }
        return r;
    }
    r.kind = ResponseKind::Error;
    return r;
  }
  ProgressTracker tracker;

 private:
  uint64_t time_cap_;
  uint32_t update_cap_;
  bool round_advanced_ = false;
};

// ------------------------------------------------------------------------------------
// Dataset slice hand-out with per-worker affinity and stealing
// ------------------------------------------------------------------------------------
class SliceTracker {
 public:
  explicit SliceTracker(uint64_t num_slices) : slices_(num_slices) {
    if (num_slices == 0) throw std::invalid_argument("SliceTracker needs at least one slice");
  }
  uint64_t next(const std::string& peer) {
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (auto s = find_open(peer)) return claim(*s, peer);
      // steal from the peer with the fewest open slices (reference behaviour, quirk Q6;
      // ties broken by peer id so the choice is deterministic)
      std::map<std::string, uint32_t> counts;
      for (auto& s : slices_) if (!s.processed) counts[s.peer] += 1;
      if (!counts.empty()) {
        auto victim = std::min_element(counts.begin(), counts.end(), [](auto& a, auto& b) { return a.second < b.second; });
        if (auto s = find_open(victim->first)) return claim(*s, peer);
      }
      ++rounds_;  // everything processed: new epoch, ownership (cache affinity) sticks
      for (auto& s : slices_) s.processed = false;
    }
    throw std::logic_error("SliceTracker: no slice available");
  }
  void remove_worker(const std::string& peer) {
    for (auto& s : slices_) if (s.owned && s.peer == peer) { s.owned = false; s.peer.clear(); }
    auto it = processing_.find(peer);
    if (it != processing_.end()) { slices_[it->second].processed = false; processing_.erase(it); }
  }
  uint32_t rounds() const { return rounds_; }
  uint64_t num_slices() const { return slices_.size(); }
  // (owner or "", processed) per slice -- checkpointable state
  std::vector<std::pair<std::string, bool>> snapshot() const {
    std::vector<std::pair<std::string, bool>> out;
    for (auto& s : slices_) out.emplace_back(s.owned ? s.peer : std::string(), s.processed);
    return out;
  }
  void restore(const std::vector<std::pair<std::string, bool>>& st, uint32_t rounds) {
    if (st.size() != slices_.size()) throw std::invalid_argument("restore: size mismatch");
    for (size_t i = 0; i < st.size(); ++i) { slices_[i].owned = !st[i].first.empty(); slices_[i].peer = st[i].first; slices_[i].processed = st[i].second; }
    rounds_ = rounds;
    processing_.clear();
  }

 private:
  struct Slice { bool owned = false; std::string peer; bool processed = false; };
  std::optional<size_t> find_open(const std::string& peer) const {
    for (size_t i = 0; i < slices_.size(); ++i)
      if (!slices_[i].processed && (!slices_[i].owned || slices_[i].peer == peer)) return i;
    return std::nullopt;
  }
  uint64_t claim(size_t i, const std::string& peer) {
    slices_[i].processed = true; slices_[i].owned = true; slices_[i].peer = peer;
    processing_[peer] = i;
    return i;
  }
  std::vector<Slice> slices_;
  std::map<std::string, size_t> processing_;
  uint32_t rounds_ = 0;
};

// ------------------------------------------------------------------------------------
// Count-or-time window over a stream of items (arbiter ad batching: 100 msgs / 200 ms)
// ------------------------------------------------------------------------------------
template <class T>
class BatchWindow {
 public:
  BatchWindow(size_t limit, double window_s) : limit_(limit), window_(window_s) {}
  // Returns a full batch when the count limit is hit, otherwise nothing.
  std::optional<std::vector<T>> push(T item, double now) {
    if (items_.empty()) opened_ = now;
    items_.push_back(std::move(item));
    if (items_.size() >= limit_) return take();
    return std::nullopt;
  }
  // Returns the pending batch once the window has elapsed.
  std::optional<std::vector<T>> poll(double now) {
    if (!items_.empty() && now - opened_ >= window_) return take();
    return std::nullopt;
  }
  std::optional<std::vector<T>> flush() { if (items_.empty()) return std::nullopt; return take(); }
  double next_deadline() const { return items_.empty() ? -1.0 : opened_ + window_; }
  size_t pending() const { return items_.size(); }

 private:
  std::vector<T> take() { std::vector<T> out; out.swap(items_); return out; }
  size_t limit_;
  double window_;
  double opened_ = 0;
  std::vector<T> items_;
};

}  // namespace hypha
