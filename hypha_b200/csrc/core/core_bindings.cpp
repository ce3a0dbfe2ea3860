// pybind11 surface of the control-plane core (hypha_b200._core).
#include <pybind11/functional.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "core.hpp"
#include "kad.hpp"
#include "flag_barrier_model.hpp"
#include "ps_reduce.hpp"

namespace py = pybind11;
using namespace hypha;

PYBIND11_MODULE(_core, m) {
  m.doc() = "hypha_b200 control-plane core (C++)";
  m.def("uuid4", &uuid4);
  m.def("wall_now", &wall_now);

  py::register_exception<LeaseNotFound>(m, "LeaseNotFound", PyExc_KeyError);
  py::register_exception<InsufficientResources>(m, "InsufficientResources", PyExc_RuntimeError);
  py::register_exception<UnknownReservation>(m, "UnknownReservation", PyExc_KeyError);
  py::register_exception<PeerNotFound>(m, "PeerNotFound", PyExc_KeyError);

  py::enum_<Ordering>(m, "Ordering")
      .value("Less", Ordering::Less).value("Equal", Ordering::Equal)
      .value("Greater", Ordering::Greater).value("Incomparable", Ordering::Incomparable);

  py::class_<Resources>(m, "Resources")
      .def(py::init([](double gpu, double cpu, double storage, double memory) { return Resources{gpu, cpu, storage, memory}; }),
           py::arg("gpu") = 0.0, py::arg("cpu") = 0.0, py::arg("storage") = 0.0, py::arg("memory") = 0.0)
      .def_readwrite("gpu", &Resources::gpu).def_readwrite("cpu", &Resources::cpu)
      .def_readwrite("storage", &Resources::storage).def_readwrite("memory", &Resources::memory)
      .def("__add__", [](const Resources& a, const Resources& b) { return a + b; })
      .def("__sub__", [](const Resources& a, const Resources& b) { return a - b; })
      .def("__eq__", [](const Resources& a, const Resources& b) { return a == b; })
      .def("__ge__", &Resources::ge).def("__gt__", &Resources::gt)
      .def("__le__", &Resources::le).def("__lt__", &Resources::lt)
      .def("compare", &Resources::compare)
      .def("to_dict", [](const Resources& r) {
        py::dict d; d["gpu"] = r.gpu; d["cpu"] = r.cpu; d["storage"] = r.storage; d["memory"] = r.memory; return d; })
      .def("__repr__", [](const Resources& r) {
        return "Resources(gpu=" + std::to_string(r.gpu) + ", cpu=" + std::to_string(r.cpu) + ", storage=" +
               std::to_string(r.storage) + ", memory=" + std::to_string(r.memory) + ")"; });

  py::class_<WeightedResourceEvaluator>(m, "WeightedResourceEvaluator")
      .def(py::init([](double gpu, double cpu, double memory, double storage) { return WeightedResourceEvaluator{gpu, cpu, memory, storage}; }),
           py::arg("gpu") = 25.0, py::arg("cpu") = 1.0, py::arg("memory") = 0.1, py::arg("storage") = 0.01)
      .def_readwrite("gpu", &WeightedResourceEvaluator::gpu).def_readwrite("cpu", &WeightedResourceEvaluator::cpu)
      .def_readwrite("memory", &WeightedResourceEvaluator::memory).def_readwrite("storage", &WeightedResourceEvaluator::storage)
      .def("evaluate", &WeightedResourceEvaluator::evaluate)
      .def("score", &WeightedResourceEvaluator::evaluate);

  using PyLease = Lease<py::object>;
  py::class_<PyLease>(m, "Lease")
      .def_readonly("id", &PyLease::id).def_readonly("leasable", &PyLease::leasable)
      .def_readonly("timeout", &PyLease::timeout)
      .def("is_expired", &PyLease::is_expired, py::arg("now") = -1.0);
  py::class_<Ledger<py::object>>(m, "Ledger")
      .def(py::init<>())
      .def("insert", &Ledger<py::object>::insert, py::arg("leasable"), py::arg("duration"), py::arg("now") = -1.0)
      .def("get", &Ledger<py::object>::get).def("remove", &Ledger<py::object>::remove)
      .def("renew", &Ledger<py::object>::renew, py::arg("id"), py::arg("duration"), py::arg("now") = -1.0)
      .def("list", &Ledger<py::object>::list)
      .def("list_expired", &Ledger<py::object>::list_expired, py::arg("now") = -1.0)
      .def("__len__", &Ledger<py::object>::size);

  py::class_<StaticResourceManager>(m, "StaticResourceManager")
      .def(py::init<Resources>())
      .def("capacity", &StaticResourceManager::capacity)
      .def("available", &StaticResourceManager::available)
      .def("compute_resources", &StaticResourceManager::available)
      .def("reserve", &StaticResourceManager::reserve).def("release", &StaticResourceManager::release);

  py::class_<ResourceLease>(m, "ResourceLease")
      .def_readonly("peer_id", &ResourceLease::peer_id).def_readonly("reservation", &ResourceLease::reservation)
      .def_readonly("resources", &ResourceLease::resources);
  using RLease = Lease<ResourceLease>;
  py::class_<RLease>(m, "ResourceLeaseEntry")
      .def_readonly("id", &RLease::id).def_readonly("leasable", &RLease::leasable)
      .def_readonly("timeout", &RLease::timeout)
      .def("is_expired", &RLease::is_expired, py::arg("now") = -1.0);
  py::class_<ResourceLeaseManager>(m, "ResourceLeaseManager")
      .def(py::init<Resources>())
      .def("request", &ResourceLeaseManager::request, py::arg("peer"), py::arg("resources"), py::arg("duration"), py::arg("now") = -1.0)
      .def("get", &ResourceLeaseManager::get)
      .def("get_by_peer", &ResourceLeaseManager::get_by_peer, py::arg("peer"), py::arg("now") = -1.0)
      .def("remove", &ResourceLeaseManager::remove)
      .def("renew", &ResourceLeaseManager::renew, py::arg("id"), py::arg("duration"), py::arg("now") = -1.0)
      .def("proceed", &ResourceLeaseManager::proceed, py::arg("now") = -1.0)
      .def("list", &ResourceLeaseManager::list)
      .def("available", [](ResourceLeaseManager& s) { return s.resources().available(); })
      .def("capacity", [](ResourceLeaseManager& s) { return s.resources().capacity(); });

  py::class_<Offer>(m, "Offer")
      .def(py::init([](std::string peer, std::string lease_id, std::string request_id, double price, Resources r, double timeout) {
             Offer o; o.peer_id = std::move(peer); o.lease_id = std::move(lease_id); o.request_id = std::move(request_id);
             o.price = price; o.resources = r; o.timeout = timeout; return o; }),
           py::arg("peer_id"), py::arg("lease_id"), py::arg("request_id"), py::arg("price"), py::arg("resources"), py::arg("timeout"))
      .def_readonly("peer_id", &Offer::peer_id).def_readonly("lease_id", &Offer::lease_id)
      .def_readonly("request_id", &Offer::request_id).def_readonly("price", &Offer::price)
      .def_readonly("resources", &Offer::resources).def_readonly("timeout", &Offer::timeout)
      .def_readonly("score", &Offer::score);
  py::enum_<OfferVerdict>(m, "OfferVerdict")
      .value("Accepted", OfferVerdict::Accepted).value("RejectedPrice", OfferVerdict::RejectedPrice)
      .value("RejectedWorse", OfferVerdict::RejectedWorse).value("Full", OfferVerdict::Full)
      .value("Closed", OfferVerdict::Closed);
  py::class_<OfferAggregator>(m, "OfferAggregator")
      .def(py::init<double, double, size_t, bool, double, WeightedResourceEvaluator, long>(), py::arg("now"),
           py::arg("deadline_s"), py::arg("desired"), py::arg("diversity") = true,
           py::arg("upper_price") = std::numeric_limits<double>::infinity(),
           py::arg("evaluator") = WeightedResourceEvaluator{}, py::arg("max_offers") = -1)
      .def("push", &OfferAggregator::push).def("done", &OfferAggregator::done)
      .def("deadline", &OfferAggregator::deadline).def("offers", &OfferAggregator::offers)
      .def("finish", &OfferAggregator::finish).def("received", &OfferAggregator::received);

  py::class_<WorkerRequestView>(m, "WorkerRequestView")
      .def(py::init([](size_t index, double bid, Resources r, bool supported) {
             WorkerRequestView v; v.index = index; v.bid = bid; v.resources = r; v.executors_supported = supported; return v; }),
           py::arg("index"), py::arg("bid"), py::arg("resources"), py::arg("executors_supported") = true)
      .def_readonly("index", &WorkerRequestView::index).def_readonly("bid", &WorkerRequestView::bid)
      .def_readonly("resources", &WorkerRequestView::resources).def_readonly("score", &WorkerRequestView::score);
  m.def("rank_requests", &rank_requests, py::arg("requests"), py::arg("capacity"), py::arg("floor"),
        py::arg("evaluator") = WeightedResourceEvaluator{});

  py::class_<Projection>(m, "Projection")
      .def_readonly("time", &Projection::time).def_readonly("remaining", &Projection::remaining)
      .def_readonly("updates", &Projection::updates).def_readonly("capped", &Projection::capped)
      .def("astuple", [](const Projection& p) { return py::make_tuple(p.time, p.remaining, p.updates, p.capped); });
  m.def("simulate_sync_point", &simulate_sync_point, py::arg("last_update"), py::arg("batch_sizes"),
        py::arg("mean_latency"), py::arg("data_points_left"), py::arg("time_cap"), py::arg("steps_cap"));

  py::class_<RunningMean>(m, "RunningMean")
      .def(py::init<>()).def("update", &RunningMean::update).def("value", &RunningMean::value)
      .def("samples", &RunningMean::samples);

  py::enum_<WorkerState>(m, "WorkerState")
      .value("Training", WorkerState::Training).value("UpdateScheduled", WorkerState::UpdateScheduled)
      .value("Updating", WorkerState::Updating).value("Done", WorkerState::Done);
  py::class_<WorkerTracker>(m, "WorkerTracker")
      .def(py::init<>())
      .def("worker_position", &WorkerTracker::position).def("contains", &WorkerTracker::contains)
      .def("add_worker", &WorkerTracker::add_worker).def("remove_worker", &WorkerTracker::remove_worker)
      .def("update", &WorkerTracker::update).def("batch_sizes", &WorkerTracker::batch_sizes)
      .def("last_updates", &WorkerTracker::last_updates).def("estimates", &WorkerTracker::estimates)
      .def("worker_state", &WorkerTracker::state).def("update_worker_state", &WorkerTracker::set_state)
      .def("worker", &WorkerTracker::workers).def("states", &WorkerTracker::states)
      .def("new_round", &WorkerTracker::new_round).def("done", &WorkerTracker::done)
      .def("__len__", &WorkerTracker::size);
  py::class_<ProgressTracker>(m, "ProgressTracker")
      .def(py::init<std::string, uint32_t, uint32_t, uint64_t>(), py::arg("parameter_server"), py::arg("update_target"),
           py::arg("update_epochs"), py::arg("now_ms") = 0)
      .def("update", &ProgressTracker::update).def("next_round", &ProgressTracker::next_round)
      .def("count", &ProgressTracker::count).def("round", &ProgressTracker::round)
      .def("training_finished", &ProgressTracker::training_finished)
      .def("parameter_server", &ProgressTracker::parameter_server)
      .def("update_parameter_server", &ProgressTracker::set_parameter_server)
      .def_readwrite("worker_tracker", &ProgressTracker::workers, py::return_value_policy::reference_internal);

  py::enum_<ProgressKind>(m, "ProgressKind")
      .value("Status", ProgressKind::Status).value("Metrics", ProgressKind::Metrics)
      .value("Update", ProgressKind::Update).value("Updated", ProgressKind::Updated)
      .value("UpdateReceived", ProgressKind::UpdateReceived);
  py::enum_<ResponseKind>(m, "ResponseKind")
      .value("Ok", ResponseKind::Ok).value("Continue", ResponseKind::Continue)
      .value("ScheduleUpdate", ResponseKind::ScheduleUpdate).value("Done", ResponseKind::Done)
      .value("Error", ResponseKind::Error);
  py::class_<ProgressResponse>(m, "ProgressResponse")
      .def_readonly("kind", &ProgressResponse::kind).def_readonly("counter", &ProgressResponse::counter)
      .def_readonly("all_done", &ProgressResponse::all_done);
  py::class_<BatchScheduler>(m, "BatchScheduler")
      .def(py::init<ProgressTracker, uint64_t, uint32_t>(), py::arg("tracker"), py::arg("time_cap_ms") = 10000,
           py::arg("update_cap") = 3)
      .def("handle", &BatchScheduler::handle, py::arg("peer"), py::arg("kind"), py::arg("batch_size"), py::arg("now_ms"))
      .def_readwrite("tracker", &BatchScheduler::tracker, py::return_value_policy::reference_internal);

  py::class_<SliceTracker>(m, "SliceTracker")
      .def(py::init<uint64_t>())
      .def("next", &SliceTracker::next).def("remove_worker", &SliceTracker::remove_worker)
      .def("rounds", &SliceTracker::rounds).def("num_slices", &SliceTracker::num_slices)
      .def("snapshot", &SliceTracker::snapshot).def("restore", &SliceTracker::restore);

  using PyWindow = BatchWindow<py::object>;
  py::class_<PyWindow>(m, "BatchWindow")
      .def(py::init<size_t, double>(), py::arg("limit"), py::arg("window_s"))
      .def("push", &PyWindow::push).def("poll", &PyWindow::poll).def("flush", &PyWindow::flush)
      .def("next_deadline", &PyWindow::next_deadline).def("pending", &PyWindow::pending);

  py::class_<FlagBarrierReport>(m, "FlagBarrierReport")
      .def_readonly("ok", &FlagBarrierReport::ok)
      .def_readonly("barriers_passed", &FlagBarrierReport::barriers_passed)
      .def_readonly("violations", &FlagBarrierReport::violations)
      .def_readonly("timeouts", &FlagBarrierReport::timeouts)
      .def_readonly("missing_union", &FlagBarrierReport::missing_union);
  // ---- Kademlia routing table (kad.hpp) ----
  m.def("sha256", [](const std::string& s) {
    const hb::kad::Id d = hb::kad::sha256(s);
    return py::bytes(reinterpret_cast<const char*>(d.data()), d.size());
  });
  py::class_<hb::kad::RoutingTable>(m, "KadTable")
      .def(py::init<const std::string&, int>(), py::arg("self_id"), py::arg("k") = 20)
      .def("add", &hb::kad::RoutingTable::add, py::arg("peer"), py::arg("addrs"), py::arg("connected") = true)
      .def("mark_disconnected", &hb::kad::RoutingTable::mark_disconnected)
      .def("remove", &hb::kad::RoutingTable::remove)
      .def("contains", &hb::kad::RoutingTable::contains)
      .def("addrs", &hb::kad::RoutingTable::addrs)
      .def("peers", &hb::kad::RoutingTable::peers)
      .def("bucket", &hb::kad::RoutingTable::bucket)
      .def("bucket_index", &hb::kad::RoutingTable::bucket_index)
      .def("closest", &hb::kad::RoutingTable::closest, py::arg("key"), py::arg("n") = 20, py::arg("include_self") = false)
      .def("__len__", &hb::kad::RoutingTable::size);

  m.def("flag_barrier_stress",
        [](int world, uint32_t member_mask, int epochs, int stages, uint32_t seed, int dead_rank, int timeout_ms,
           uint32_t first_epoch) {
          py::gil_scoped_release nogil;
          FlagBarrierModel model(world, member_mask);
          return model.stress(epochs, stages, seed, dead_rank, timeout_ms, first_epoch);
        },
        py::arg("world"), py::arg("member_mask"), py::arg("epochs"), py::arg("stages"), py::arg("seed") = 1,
        py::arg("dead_rank") = -1, py::arg("timeout_ms") = 2000, py::arg("first_epoch") = 1,
        "Host stress test of the device flag-barrier protocol (outer_step.cu: cross_gpu_barrier).");

  m.def("ps_fused_round",
        [](const std::vector<std::string>& in_paths, const std::vector<std::vector<uint64_t>>& in_offsets,
           const std::vector<uint64_t>& numels, const std::string& prev_m_path,
           const std::vector<uint64_t>& prev_m_offsets, const std::string& prev_cum_path,
           const std::vector<uint64_t>& prev_cum_offsets, const std::string& out_u_path,
           const std::string& out_m_path, const std::string& out_cum_path,
           const std::vector<uint64_t>& out_offsets, double lr, double mu, int threads,
           const std::vector<double>& weights) {
          PsRoundStats s;
          {
            py::gil_scoped_release nogil;
            s = ps_fused_round(in_paths, in_offsets, numels, prev_m_path, prev_m_offsets, prev_cum_path,
                               prev_cum_offsets, out_u_path, out_m_path, out_cum_path, out_offsets, lr, mu, threads,
                               weights);
          }
          py::dict d;
          d["elements"] = s.elements; d["bytes_read"] = s.bytes_read; d["bytes_written"] = s.bytes_written;
          d["threads"] = s.threads; d["seconds"] = s.seconds;
          return d;
        },
        py::arg("in_paths"), py::arg("in_offsets"), py::arg("numels"), py::arg("prev_m_path"),
        py::arg("prev_m_offsets"), py::arg("prev_cum_path"), py::arg("prev_cum_offsets"), py::arg("out_u_path"),
        py::arg("out_m_path"), py::arg("out_cum_path"), py::arg("out_offsets"), py::arg("lr"), py::arg("mu"),
        py::arg("threads") = 0, py::arg("weights") = std::vector<double>{},
        "One parameter-server round (mean, momentum, Nesterov, cumulative update) over mmapped safetensors files.");
}
