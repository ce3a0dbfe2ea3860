"""Reference test vectors for the C++ control-plane core."""
import pytest

from hypha_b200.ops.native import load_core

core = load_core()


def test_simulation_hits_time_cap():   # simulation.rs:76-92
    p = core.simulate_sync_point([0, 0, 0], [10, 5, 1], [20, 30, 50], 9999, 101, 999)
    assert p.astuple() == (100, 9999 - 67, [5, 3, 2], True)


def test_simulation_hits_steps_cap():  # simulation.rs:94-110
    p = core.simulate_sync_point([0, 0, 0], [10, 5, 1], [20, 30, 50], 9999, 999, 3)
    assert p.astuple() == (60, 9999 - 41, [3, 2, 1], True)


def test_simulation_hits_update_target():  # simulation.rs:112-128
    p = core.simulate_sync_point([0, 0, 0], [10, 5, 1], [20, 30, 50], 50, 999, 999)
    assert p.astuple() == (80, 0, [4, 2, 1], False)


def test_running_mean_update():  # statistics.rs:50-69
    rm = core.RunningMean()
    for x, want in ((1050, 1050), (1000, 1025), (1025, 1025), (2050, 1281)):
        rm.update(x)
        assert rm.value() == want
    assert rm.samples() == 4


def test_slice_simple_scheduling():  # tracker/slice.rs:123-141
    t = core.SliceTracker(7)
    seq = [("a", 0), ("b", 1), ("a", 2), ("b", 3), ("a", 4), ("b", 5), ("a", 6), ("b", 1), ("a", 0)]
    for peer, want in seq:
        assert t.next(peer) == want


def test_slice_fast_slow_scheduling():  # tracker/slice.rs:143-171
    t = core.SliceTracker(7)
    for peer, want in [("a", 0), ("a", 1), ("a", 2), ("b", 3), ("a", 4), ("a", 5), ("a", 6), ("b", 3),
                       ("a", 0), ("a", 1), ("a", 2)]:
        assert t.next(peer) == want
    left = [4, 5, 6]
    s = t.next("b")
    assert s in left
    left.remove(s)
    assert t.next("a") in left
    assert t.next("a") in left
    assert t.next("a") == 0
    assert t.next("b") == 3
    assert t.rounds() == 2


def test_slice_peer_removed_scheduling():  # tracker/slice.rs:173-202
    t = core.SliceTracker(7)
    for peer, want in [("a", 0), ("b", 1), ("a", 2), ("b", 3), ("a", 4), ("b", 5), ("a", 6)]:
        assert t.next(peer) == want
    t.remove_worker("a")
    for peer, want in [("c", 6), ("b", 0), ("c", 2), ("b", 1), ("c", 4), ("b", 3), ("c", 6), ("b", 5)]:
        assert t.next(peer) == want


def test_slice_snapshot_restore():
    t = core.SliceTracker(4)
    t.next("a"); t.next("b")
    snap, rounds = t.snapshot(), t.rounds()
    u = core.SliceTracker(4)
    u.restore(snap, rounds)
    assert u.next("a") == 2 and u.snapshot()[0] == ("a", True)


def test_worker_position_and_remove():  # tracker/progress.rs worker_position
    w = core.WorkerTracker()
    w.add_worker("x", 12); w.add_worker("p", 2); w.add_worker("y", 12)
    assert w.worker_position("p") == 1
    w.remove_worker("p")
    with pytest.raises(KeyError):
        w.worker_position("nobody")


def test_progress_tracker_peer_info():  # tracker/progress.rs peer_info
    t = core.ProgressTracker("ps", 1024, 1, 0)
    t.worker_tracker.add_worker("p1", 2)
    t.worker_tracker.add_worker("p2", 12)
    assert t.worker_tracker.worker() == ["p1", "p2"]
    assert t.worker_tracker.batch_sizes() == [2, 12]
    t.update("p1", 2, 12)
    assert t.count() == 1022
    assert t.worker_tracker.last_updates() == [12, 0]
    t.update("p2", 12, 50)
    assert t.worker_tracker.estimates() == [12, 50]
    t.next_round(50)
    assert t.worker_tracker.last_updates() == [0, 0]
    assert t.count() == 1024 and t.training_finished()
    t.worker_tracker.done()
    assert t.worker_tracker.states() == [core.WorkerState.Done] * 2


def _run_script(sched, steps, start=0):
    now = start
    for i, (peer, kind, bs, want_kind, want_counter, dt) in enumerate(steps):
        now += dt
        r = sched.handle(peer, kind, bs, now)
        assert r.kind == want_kind, f"step {i}: {peer} {kind} -> {r.kind}, want {want_kind}"
        if want_counter is not None:
            assert r.counter == want_counter, f"step {i}: counter {r.counter} != {want_counter}"
    return now


def test_batch_scheduler_two_rounds():
    """Scripted table of batch_scheduler.rs:346-447 (`handle_simple_two_rounds`)."""
    K, R = core.ProgressKind, core.ResponseKind
    tr = core.ProgressTracker("ps", 800, 2, 0)
    for w, bs in (("w1", 150), ("w2", 100), ("w3", 50)):
        tr.worker_tracker.add_worker(w, bs)
    s = core.BatchScheduler(tr)
    steps = [
        ("w3", K.Status, 50, R.Continue, None, 500),
        ("w2", K.Status, 100, R.Continue, None, 300),
        ("w1", K.Status, 150, R.Continue, None, 150),
        ("w3", K.Status, 50, R.ScheduleUpdate, 2, 50),
        ("w3", K.Status, 50, R.Continue, None, 500),
        ("w2", K.Status, 100, R.ScheduleUpdate, 1, 100),
        ("w1", K.Status, 150, R.ScheduleUpdate, 0, 300),
        ("w3", K.Status, 50, R.Continue, None, 100),
        ("w2", K.Status, 100, R.Continue, None, 100),
        ("w1", K.Update, 0, R.Ok, None, 0), ("w2", K.Update, 0, R.Ok, None, 0), ("w3", K.Update, 0, R.Ok, None, 0),
        ("ps", K.Updated, 0, R.Ok, None, 0),
        ("w1", K.UpdateReceived, 0, R.Continue, None, 0), ("w2", K.UpdateReceived, 0, R.Continue, None, 0),
        ("w3", K.UpdateReceived, 0, R.Continue, None, 0),
    ]
    end = _run_script(s, steps)
    assert s.tracker.worker_tracker.states() == [core.WorkerState.Training] * 3
    assert s.tracker.count() == 800 and s.tracker.round() == 1
    # second (final) round ends with Done and the all-done notification
    _run_script(s, steps[:12] + [("ps", K.Updated, 0, R.Done, None, 0)], start=end)
    for i, w in enumerate(("w1", "w2", "w3")):
        r = s.handle(w, K.UpdateReceived, 0, 10_000)
        assert r.kind == R.Done and r.all_done == (i == 2)


def test_batch_scheduler_single_worker_responding():
    """batch_scheduler.rs:449-482."""
    K, R = core.ProgressKind, core.ResponseKind
    tr = core.ProgressTracker("ps", 200, 2, 0)
    tr.worker_tracker.add_worker("w1", 150)
    tr.worker_tracker.add_worker("w2", 100)
    s = core.BatchScheduler(tr)
    _run_script(s, [("w1", K.Status, 100, R.ScheduleUpdate, 1, 100), ("w1", K.Status, 100, R.Continue, None, 100)])
    assert s.tracker.worker_tracker.states() == [core.WorkerState.UpdateScheduled, core.WorkerState.Training]


def test_batch_scheduler_wrong_state_is_error():
    K, R = core.ProgressKind, core.ResponseKind
    tr = core.ProgressTracker("ps", 100, 1, 0)
    tr.worker_tracker.add_worker("w", 10)
    s = core.BatchScheduler(tr)
    assert s.handle("w", K.Update, 0, 1).kind == R.Ok
    assert s.handle("w", K.Status, 10, 2).kind == R.Error
    assert s.handle("stranger", K.Status, 10, 2).kind == R.Error


def test_batch_scheduler_metrics_ok():
    tr = core.ProgressTracker("ps", 10, 1, 0)
    tr.worker_tracker.add_worker("w", 1)
    s = core.BatchScheduler(tr)
    assert s.handle("w", core.ProgressKind.Metrics, 0, 5).kind == core.ResponseKind.Ok


def test_resources_partial_order():
    R = core.Resources
    a, b = R(gpu=2, cpu=4, storage=10, memory=8), R(gpu=1, cpu=4, storage=10, memory=8)
    assert a >= b and a > b and not (b >= a) and b < a
    c = R(gpu=1, cpu=8, storage=10, memory=8)
    assert a.compare(c) == core.Ordering.Incomparable
    assert not (a >= c) and not (c >= a) and not (a > c)
    assert (a + b).gpu == 3 and (a - b).gpu == 1 and a == R(gpu=2, cpu=4, storage=10, memory=8)


def test_weighted_evaluator_defaults():
    ev = core.WeightedResourceEvaluator()
    r = core.Resources(gpu=1, cpu=2, storage=100, memory=10)
    assert ev.evaluate(100.0, r) == pytest.approx(100.0 / (25 + 2 + 1 + 1))
    assert ev.evaluate(5.0, core.Resources()) == 0.0


def test_ledger_semantics():
    led = core.Ledger()
    l = led.insert({"x": 1}, 10.0, now=100.0)
    assert l.timeout == 110.0 and led.get(l.id).leasable == {"x": 1}
    assert led.list_expired(now=105.0) == []
    r = led.renew(l.id, 10.0, now=108.0)       # renew = now + d, not extend
    assert r.timeout == 118.0
    assert [x.id for x in led.list_expired(now=118.0)] == [l.id]
    led.remove(l.id)
    with pytest.raises(KeyError):
        led.get(l.id)
    with pytest.raises(KeyError):
        led.renew("nope", 1.0)


def test_resource_lease_manager():
    R = core.Resources
    m = core.ResourceLeaseManager(R(gpu=4, cpu=8, storage=100, memory=64))
    l1 = m.request("s1", R(gpu=3, cpu=1, storage=1, memory=1), 0.5, now=0.0)
    with pytest.raises(RuntimeError):
        m.request("s2", R(gpu=2, cpu=1, storage=1, memory=1), 0.5, now=0.0)
    assert m.available().gpu == 1
    m.renew(l1.id, 10.0, now=0.2)
    assert m.proceed(now=1.0) == []
    l2 = m.request("s2", R(gpu=1), 0.5, now=1.0)
    gone = m.proceed(now=2.0)
    assert [g.id for g in gone] == [l2.id] and m.available().gpu == 1
    assert m.get_by_peer("s1", now=2.0).id == l1.id
    m.remove(l1.id)
    assert m.available().gpu == 4
    with pytest.raises(KeyError):
        m.get_by_peer("s1")


def test_offer_aggregator_diversity_price_and_early_return():
    R = core.Resources
    agg = core.OfferAggregator(now=0.0, deadline_s=5.0, desired=2, diversity=True, upper_price=100.0)
    O = lambda peer, price, timeout=10.0: core.Offer(peer, f"lease-{peer}-{price}", "req", price, R(gpu=1), timeout)
    V = core.OfferVerdict
    assert agg.push(O("a", 150.0), 0.1) == V.RejectedPrice
    assert agg.push(O("a", 80.0, timeout=0.6), 0.1) == V.Accepted
    assert agg.deadline() == pytest.approx(0.5)          # earliest expiry - 100 ms
    assert agg.push(O("a", 90.0), 0.2) == V.RejectedWorse
    assert agg.push(O("a", 70.0), 0.2) == V.Accepted     # better offer from same peer replaces
    assert not agg.done(0.2)
    assert agg.push(O("b", 60.0), 0.3) == V.Accepted
    assert agg.done(0.3)                                   # early return at desired count
    assert [o.peer_id for o in agg.finish()] == ["b", "a"]
    assert agg.push(O("c", 1.0), 0.4) == V.Closed


def test_offer_aggregator_deadline_and_full():
    R = core.Resources
    agg = core.OfferAggregator(now=0.0, deadline_s=5.0, desired=1, diversity=False)
    assert not agg.done(4.9) and agg.done(5.0)
    agg2 = core.OfferAggregator(now=0.0, deadline_s=5.0, desired=3, diversity=False, max_offers=2)
    agg2.push(core.Offer("a", "l", "r", 10.0, R(gpu=1), 100.0), 0.0)
    agg2.push(core.Offer("b", "l", "r", 20.0, R(gpu=1), 100.0), 0.0)
    assert agg2.done(0.0) and len(agg2.offers()) == 2


def test_rank_requests():
    R = core.Resources
    cap = R(gpu=2, cpu=8, storage=100, memory=64)
    reqs = [core.WorkerRequestView(0, 50.0, R(gpu=1)), core.WorkerRequestView(1, 200.0, R(gpu=1)),
            core.WorkerRequestView(2, 500.0, R(gpu=4, cpu=16, storage=200, memory=128)),   # > capacity
            core.WorkerRequestView(3, 150.0, R(gpu=1), False),                               # unsupported
            core.WorkerRequestView(4, 120.0, R(gpu=1, cpu=64))]                              # incomparable: kept
    ranked = core.rank_requests(reqs, cap, 100.0)
    assert [r.index for r in ranked] == [1, 4]


def test_batch_window():
    w = core.BatchWindow(3, 0.2)
    assert w.push("a", 0.0) is None and w.poll(0.1) is None
    assert w.poll(0.2) == ["a"]
    assert w.push("b", 1.0) is None and w.push("c", 1.01) is None
    assert w.push("d", 1.02) == ["b", "c", "d"]
    assert w.pending() == 0 and w.flush() is None


# ---------------------------------------------------------------------------------------------
# host model of the device flag barrier (csrc/core/flag_barrier_model.hpp)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 4, 8])
def test_flag_barrier_protocol_stress(world):
    """Two slots + epoch-stamped tokens are enough for a K+2-stage pipeline: nobody leaves a barrier
    before every member entered it, for any interleaving the scheduler produces."""
    core = load_core()
    rep = core.flag_barrier_stress(world, (1 << world) - 1, epochs=60, stages=10, seed=world)
    assert rep.ok and rep.violations == 0 and rep.timeouts == 0
    assert rep.barriers_passed == world * 60 * 10


def test_flag_barrier_token_wraparound():
    core = load_core()
    rep = core.flag_barrier_stress(4, 0xF, epochs=64, stages=18, seed=5, first_epoch=(0xFFFFFFFF // 32) - 20)
    assert rep.ok and rep.violations == 0


def test_flag_barrier_reports_dead_rank():
    """A member that never arrives costs a timeout and is named in the missing mask (the kernel then
    returns status != 0 and the host shrinks the membership) instead of hanging the survivors."""
    core = load_core()
    rep = core.flag_barrier_stress(4, 0xF, epochs=3, stages=4, seed=9, dead_rank=2, timeout_ms=100)
    assert rep.ok and rep.timeouts == 3 and rep.missing_union == 0b0100
    # with the membership shrunk to the survivors the same ranks synchronise again
    rep2 = core.flag_barrier_stress(4, 0b1011, epochs=20, stages=4, seed=9)
    assert rep2.ok and rep2.barriers_passed == 3 * 20 * 4


def test_flag_barrier_model_under_thread_sanitizer():
    """`tools/sanitize.sh tsan`: the protocol model is data-race-free under ThreadSanitizer."""
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cxx = os.environ.get("HYPHA_SANITIZE_CXX", "/usr/bin/g++")
    if not os.path.exists(cxx) or shutil.which("sh") is None:
        pytest.skip("no sanitizer-capable compiler")
    probe = subprocess.run([cxx, "-print-file-name=libtsan.so"], capture_output=True, text=True)
    if not os.path.isabs(probe.stdout.strip()):
        pytest.skip("libtsan not installed")
    r = subprocess.run(["sh", os.path.join(root, "tools", "sanitize.sh"), "tsan"], capture_output=True, text=True,
                       timeout=300, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout and "ThreadSanitizer" not in r.stderr


def test_updated_is_idempotent_within_one_synchronisation():
    """Collective jobs: every rank reports `updated`; only the first one after the workers' `update`
    messages advances the round (the reference has a single aggregate worker and no such rule)."""
    core = load_core()
    tr = core.ProgressTracker("w0", 8, 2, 0)
    for w in ("w0", "w1"):
        tr.worker_tracker.add_worker(w, 4)
    sched = core.BatchScheduler(tr, 10000, 3)
    K, R = core.ProgressKind, core.ResponseKind
    t = 0
    for rnd in range(2):
        for w in ("w0", "w1"):
            t += 10
            assert sched.handle(w, K.Update, 0, t).kind == R.Ok
        want = R.Done if rnd == 1 else R.Ok
        for w in ("w0", "w1"):                      # both ranks report `updated`
            t += 1
            assert sched.handle(w, K.Updated, 0, t).kind == want
        assert sched.tracker.round() == rnd + 1     # advanced exactly once
        acks = [sched.handle(w, K.UpdateReceived, 0, t + 5) for w in ("w0", "w1")]
        assert all(a.kind == (R.Done if rnd == 1 else R.Continue) for a in acks)
    assert acks[-1].all_done and sched.tracker.training_finished()


def test_native_kad_table_matches_the_xor_metric():
    """`_core.KadTable` (csrc/core/kad.hpp): native SHA-256 ids, XOR ordering, bucket index = shared prefix length,
    least-recently-seen order, live entries win -- checked against hashlib / big-integer arithmetic."""
    import hashlib

    from hypha_b200.ops.native import load_core

    core = load_core()
    for m in ("", "abc", "a" * 55, "b" * 56, "c" * 64, "d" * 1000):
        assert core.sha256(m) == hashlib.sha256(m.encode()).digest()

    def h(s):
        return int.from_bytes(hashlib.sha256(s.encode()).digest(), "big")

    t = core.KadTable("self", 3)
    peers = [f"peer-{i}" for i in range(300)]
    accepted = [p for p in peers if t.add(p, [f"tcp://1.2.3.{i % 250}:1" for i in range(2)], True)]
    assert len(t) == len(accepted) and not t.add("self", [], True)
    for p in peers[:40]:
        assert t.bucket_index(p) == 256 - (h(p) ^ h("self")).bit_length()
    assert all(len(t.bucket(i)) <= 3 for i in range(256)) and sum(len(t.bucket(i)) for i in range(256)) == len(t)
    for key in ("k1", "dataset", "another key"):
        want = sorted(accepted, key=lambda p: h(p) ^ h(key))
        assert t.closest(key, 7, False) == want[:7]
        with_self = sorted(accepted + ["self"], key=lambda p: h(p) ^ h(key))
        assert t.closest(key, len(accepted) + 5, True) == with_self
    assert t.closest("k", 0, False) == []
    full = next(i for i in range(256) if len(t.bucket(i)) == 3)
    rejected = next(p for p in peers if p not in accepted and t.bucket_index(p) == full)
    assert not t.add(rejected, ["tcp://9.9.9.9:1"], True)
    victim = t.bucket(full)[1]
    t.mark_disconnected(victim)
    assert t.add(rejected, ["tcp://9.9.9.9:1"], True) and not t.contains(victim) and t.bucket(full)[-1] == rejected
    t.remove(rejected)
    assert not t.contains(rejected) and t.addrs(rejected) == [] and len(t.bucket(full)) == 2
