// _fedcore: pybind11 surface of the native party runtime.
//   Ref             one-shot future (replaces ray.ObjectRef)
//   Executor        task scheduler + actor mailboxes (replaces Ray core worker)
//   Mailbox         (upstream, downstream) rendezvous table (replaces the
//                   receiver proxy actor's event tables)
//   SocketTransport framed TCP/Unix transport with pipelined acks
//   SeqCounter      the deterministic sequence-id mint
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "executor.h"
#include "mailbox.h"
#include "ref.h"
#include "send_tracker.h"
#include "transport.h"

namespace py = pybind11;
using namespace fedcore;

namespace {

[[noreturn]] void raise_timeout(const char* what) {
  PyErr_SetString(PyExc_TimeoutError, what);
  throw py::error_already_set();
}

class SeqCounter {
 public:
  uint64_t next() { return v_.fetch_add(1, std::memory_order_relaxed) + 1; }
  uint64_t value() const { return v_.load(std::memory_order_relaxed); }
  void reset(uint64_t v) { v_.store(v, std::memory_order_relaxed); }

 private:
  std::atomic<uint64_t> v_{0};
};

// Block on a Ref without starving Python signal handlers: the wait is sliced and
// PyErr_CheckSignals() runs between slices, so SIGINT (the exit-on-failure path,
// fed/cleanup.py:112-128) interrupts a blocked fed.get like it interrupts ray.get.
bool wait_interruptible(Ref& r, double timeout_s) {
  if (r.done()) return true;
  const double slice = 0.05;
  auto start = std::chrono::steady_clock::now();
  for (;;) {
    double remaining = slice;
    if (timeout_s >= 0) {
      double left = timeout_s - std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
      if (left <= 0) return r.done();
      remaining = std::min(slice, left);
    }
    bool ok;
    {
      py::gil_scoped_release rel;
      NativeWaitScope parked;
      ok = r.wait(remaining);
    }
    if (ok) return true;
    if (PyErr_CheckSignals() != 0) throw py::error_already_set();
  }
}

// ray.wait analogue: block until `num_ready` of `refs` resolved or timeout.
std::pair<std::vector<RefPtr>, std::vector<RefPtr>> wait_refs(const std::vector<RefPtr>& refs, size_t num_ready,
                                                              double timeout_s) {
  struct Shared {
    std::mutex mu;
    std::condition_variable cv;
    size_t n = 0;
    size_t need = 0;
  };
  auto sh = std::make_shared<Shared>();
  sh->need = std::min(num_ready, refs.size());
  for (auto& r : refs) {
    r->on_done([sh]() {
      bool wake;
      {
        std::lock_guard<std::mutex> lk(sh->mu);
        wake = ++sh->n >= sh->need;  // the waiter is woken once, when enough of them resolved
      }
      if (wake) sh->cv.notify_all();
    });
  }
  if (num_ready > refs.size()) num_ready = refs.size();
  {
    py::gil_scoped_release rel;
    NativeWaitScope parked;
    std::unique_lock<std::mutex> lk(sh->mu);
    auto pred = [&] { return sh->n >= num_ready; };
    if (timeout_s < 0)
      sh->cv.wait(lk, pred);
    else
      sh->cv.wait_for(lk, std::chrono::duration<double>(timeout_s), pred);
  }
  std::vector<RefPtr> ready, rest;
  for (auto& r : refs) {
    if (r->done() && ready.size() < num_ready)
      ready.push_back(r);
    else
      rest.push_back(r);
  }
  return {ready, rest};
}

}  // namespace

PYBIND11_MODULE(_fedcore, m) {
  m.doc() = "rayfed_b200 native party runtime (scheduler, futures, mailbox, socket transport)";

  py::class_<Ref, RefPtr>(m, "Ref")
      .def(py::init<>())
      .def("done", &Ref::done)
      .def("is_error", &Ref::is_error)
      .def(
          "wait",
          [](Ref& r, double timeout) { return wait_interruptible(r, timeout); },
          py::arg("timeout") = -1.0)
      .def(
          "result",
          [](RefPtr r, double timeout) {
            if (!wait_interruptible(*r, timeout)) raise_timeout("timed out waiting for a fed Ref");
            return r->result();
          },
          py::arg("timeout") = -1.0)
      .def(
          "exception",
          [](RefPtr r, double timeout) {
            if (!wait_interruptible(*r, timeout)) raise_timeout("timed out waiting for a fed Ref");
            return r->exception();
          },
          py::arg("timeout") = -1.0)
      .def("set_value", &Ref::set_value)
      .def("set_error", &Ref::set_error)
      .def(
          "set_raw", [](Ref& r, py::bytes b, bool is_error) { return r.set_raw(std::string(b), is_error); },
          py::arg("data"), py::arg("is_error") = false)
      // ---- device ordering (device_hooks.h) ----
      .def_property_readonly("cuda_event",
                             [](Ref& r) {
                               DeviceEventPtr e = r.device_event();
                               return e ? e->event : (uint64_t)0;
                             })
      .def_property_readonly("cuda_stream",
                             [](Ref& r) {
                               DeviceEventPtr e = r.device_event();
                               return e ? e->stream : (uint64_t)0;
                             })
      .def(
          "set_cuda_event",
          [](Ref& r, uint64_t event, uint64_t stream, py::object owner) {
            auto e = std::make_shared<DeviceEvent>();
            e->event = event;
            e->stream = stream;
            e->owner = owner.is_none() ? py::object() : owner;
            r.set_device_event(std::move(e));
          },
          py::arg("event"), py::arg("stream") = 0, py::arg("owner") = py::none())
      .def("device_sync",
           [](RefPtr r) {
             py::gil_scoped_release rel;
             if (!r->device_sync()) throw std::runtime_error("cudaEventSynchronize failed on a fed Ref");
           })
      .def("set_decoder", &Ref::set_decoder)
      .def("set_decoded_exc_is_error", &Ref::set_decoded_exc_is_error)
      .def("set_ack_mode", &Ref::set_ack_mode, py::arg("on") = true)
      .def("forward_to", &Ref::forward_to, py::arg("dst"),
           "complete `dst` with this Ref's outcome (native continuation; raw outcomes never take the GIL)")
      .def("add_done_callback", &Ref::add_done_callback);

  py::class_<SendTracker, std::shared_ptr<SendTracker>>(m, "SendTracker")
      .def(py::init<py::object>(), py::arg("handler"))
      .def("watch", &SendTracker::watch, py::arg("ref"), py::arg("pack"), py::arg("is_error") = false)
      .def("retire", &SendTracker::retire, py::arg("is_error") = false)
      .def("pending", &SendTracker::pending)
      .def("pending_errors", &SendTracker::pending_errors)
      .def("fast_retired", &SendTracker::fast_retired)
      .def("slow_settled", &SendTracker::slow_settled)
      .def("set_draining", &SendTracker::set_draining)
      .def("settling", &SendTracker::settling)
      .def("bury", &SendTracker::bury)
      .def(
          "wait_idle",
          [](SendTracker& t, double timeout, bool errors_only) {
            py::gil_scoped_release rel;
            return t.wait_idle(timeout, errors_only);
          },
          py::arg("timeout") = -1.0, py::arg("errors_only") = false);

  py::class_<DeviceValue>(m, "DeviceValue",
                          "value + the CUDA event (and stream) that orders it; a Ref completed with one "
                          "keeps the plain value and adopts the event")
      .def(py::init([](py::object value, uint64_t event, uint64_t stream, py::object owner) {
             DeviceValue dv;
             dv.value = std::move(value);
             dv.event = event;
             dv.stream = stream;
             dv.owner = owner.is_none() ? py::object() : owner;
             return dv;
           }),
           py::arg("value"), py::arg("event") = 0, py::arg("stream") = 0, py::arg("owner") = py::none())
      .def_readwrite("value", &DeviceValue::value)
      .def_readwrite("event", &DeviceValue::event)
      .def_readwrite("stream", &DeviceValue::stream)
      .def_readwrite("owner", &DeviceValue::owner);

  // The CUDA calls the scheduler needs, published by _fedcuda as a capsule (no link dependency).
  m.def("set_device_hooks", [](py::object capsule) {
    const DeviceHooks* old = device_hooks();
    const DeviceHooks* h = nullptr;
    if (!capsule.is_none()) {
      h = static_cast<const DeviceHooks*>(PyCapsule_GetPointer(capsule.ptr(), "rayfed_b200.device_hooks"));
      if (!h) throw py::error_already_set();
    }
    if (h == old) return;
    EventPool::get().drop_all(old);
    g_device_hooks.store(h, std::memory_order_release);
  });
  m.def("install_fake_device_hooks", [] {
    const DeviceHooks* old = device_hooks();
    if (old == fake_device_hooks()) return;
    EventPool::get().drop_all(old);
    g_device_hooks.store(fake_device_hooks(), std::memory_order_release);
  });
  m.def("has_device_hooks", [] { return device_hooks() != nullptr; });
  m.def("fake_device_log", [](bool clear) {
    FakeDeviceLog& l = fake_device_log();
    std::lock_guard<std::mutex> lk(l.mu);
    std::vector<std::string> out = l.lines;
    if (clear) l.lines.clear();
    return out;
  }, py::arg("clear") = false);

  m.def("set_spin_policy", [](int cap_us, int max_spinners) {
    g_adaptive_cap_us.store(cap_us, std::memory_order_relaxed);
    g_max_long_spinners.store(max_spinners, std::memory_order_relaxed);
  }, py::arg("cap_us"), py::arg("max_spinners"), "explicit adaptive-spin policy (environment variables still win)");
  m.def("set_spin_policy_auto", &set_spin_policy_auto, py::arg("local_parties") = 1,
        "derive the adaptive-spin policy from the usable CPUs and the number of parties on this host");
  m.def("spin_info", [] {
    py::dict d;
    d["usable_cpus"] = usable_cpus();
    d["base_spin_us"] = spin_budget_us();
    d["adaptive_cap_us"] = adaptive_spin_cap_us();
    d["max_long_spinners"] = max_long_spinners();
    d["long_spinners_now"] = g_long_spinners.load();
    d["preemptions_seen"] = g_long_spin_preemptions.load();
    d["pause_ms"] = g_long_spin_pause_ms.load();
    return d;
  }, "adaptive-spin configuration and counters of this process");
  m.def("wait_refs", &wait_refs, py::arg("refs"), py::arg("num_ready") = 1, py::arg("timeout") = -1.0);

  py::class_<SeqCounter>(m, "SeqCounter")
      .def(py::init<>())
      .def("next", &SeqCounter::next)
      .def("value", &SeqCounter::value)
      .def("reset", &SeqCounter::reset, py::arg("value") = 0);

  py::class_<Executor, std::shared_ptr<Executor>>(m, "Executor")
      .def(py::init<int, py::object>(), py::arg("pool_threads") = 4, py::arg("wrap_error") = py::none())
      .def("create_lane", &Executor::create_lane, py::arg("name"), py::arg("concurrency") = 1,
           py::arg("thread_init") = py::none())
      .def("set_lane_thread_init", &Executor::set_lane_thread_init, py::arg("lane"), py::arg("thread_init"))
      .def("submit", &Executor::submit, py::arg("lane"), py::arg("fn"), py::arg("args"), py::arg("kwargs"),
           py::arg("num_returns") = 1)
      .def("kill_lane", &Executor::kill_lane)
      .def(
          "shutdown",
          [](Executor& e, bool wait, double timeout) {
            py::gil_scoped_release rel;
            e.shutdown(wait, timeout);
          },
          py::arg("wait") = true, py::arg("timeout") = 5.0)
      .def("lane_backlog", &Executor::lane_backlog)
      .def("num_lanes", &Executor::num_lanes)
      .def("stats", [](Executor& e) {
        py::dict d;
        d["submitted"] = e.submitted();
        d["completed"] = e.completed();
        d["failed"] = e.failed();
        return d;
      });

  py::class_<Mailbox, std::shared_ptr<Mailbox>>(m, "Mailbox")
      .def(py::init<>())
      .def("claim", &Mailbox::claim)
      .def(
          "claim_many",
          [](Mailbox& mb, const std::vector<std::string>& ups, const std::string& down, py::object decoder) {
            std::vector<RefPtr> refs = mb.claim_many(ups, down);
            if (!decoder.is_none())
              for (auto& r : refs) r->set_decoder(decoder);
            return refs;
          },
          py::arg("ups"), py::arg("down"), py::arg("decoder") = py::none())
      .def("deliver_obj", &Mailbox::deliver_obj, py::arg("up"), py::arg("down"), py::arg("obj"),
           py::arg("is_error") = false)
      .def(
          "deliver_raw",
          [](Mailbox& mb, const std::string& up, const std::string& down, py::bytes b, bool is_error) {
            mb.deliver_raw(up, down, std::string(b), is_error);
          },
          py::arg("up"), py::arg("down"), py::arg("data"), py::arg("is_error") = false)
      .def("pending", &Mailbox::pending)
      .def(
          "snapshot",
          [](Mailbox& mb, size_t limit) {
            std::vector<std::pair<std::string, std::string>> waiting, unclaimed;
            mb.snapshot(&waiting, &unclaimed, limit);
            return py::make_tuple(waiting, unclaimed);
          },
          py::arg("limit") = 32)
      .def("clear", &Mailbox::clear)
      .def("stats", [](Mailbox& mb) {
        py::dict d;
        d["delivered"] = mb.delivered();
        d["claimed"] = mb.claimed();
        d["bytes"] = mb.bytes();
        return d;
      });

  py::class_<SocketTransport, std::shared_ptr<SocketTransport>>(m, "SocketTransport")
      .def(py::init<std::string, std::string, std::shared_ptr<Mailbox>>(), py::arg("party"), py::arg("job"),
           py::arg("mailbox"))
      .def("listen",
           [](SocketTransport& t, const std::string& addr) {
             py::gil_scoped_release rel;
             return t.listen(addr);
           })
      .def(
          "add_peer",
          [](SocketTransport& t, const std::string& name, const std::string& addr, int max_attempts,
             double initial_backoff_s, double max_backoff_s, double multiplier, double timeout_s, int max_resends) {
            RetryPolicy rp;
            rp.max_resends = max_resends < 0 ? 0 : max_resends;
            rp.max_attempts = max_attempts;
            rp.initial_backoff_s = initial_backoff_s;
            rp.max_backoff_s = max_backoff_s;
            rp.multiplier = multiplier;
            t.add_peer(name, addr, rp, timeout_s);
          },
          py::arg("name"), py::arg("addr"), py::arg("max_attempts") = 5, py::arg("initial_backoff_s") = 5.0,
          py::arg("max_backoff_s") = 30.0, py::arg("multiplier") = 2.0, py::arg("timeout_s") = 60.0,
          py::arg("max_resends") = 0)
      .def(
          "send",
          [](SocketTransport& t, const std::string& dest, const std::string& up, const std::string& down,
             py::bytes data, bool is_error, py::object decoder) {
            auto ack = std::make_shared<Ref>();
            if (!decoder.is_none()) {
              ack->set_decoder(decoder);
            } else {
              ack->set_ack_mode(true);  // True / RuntimeError(text), decoded without a Python object
            }
            std::string payload(data);
            bool ok;
            {
              py::gil_scoped_release rel;
              ok = t.send(dest, up, down, std::move(payload), is_error, ack);
            }
            if (!ok) throw py::key_error("unknown destination party '" + dest + "'");
            return ack;
          },
          py::arg("dest"), py::arg("up"), py::arg("down"), py::arg("data"), py::arg("is_error") = false,
          py::arg("decoder") = py::none())
      .def(
          "send_many",
          // one payload to several parties (fed.get's broadcast): the GIL is dropped ONCE around all
          // the writes, so the driver thread can go on while the frames leave
          [](SocketTransport& t, const std::vector<std::string>& dests, const std::string& up,
             const std::string& down, py::bytes data, bool is_error) {
            std::vector<RefPtr> acks;
            acks.reserve(dests.size());
            for (size_t i = 0; i < dests.size(); ++i) {
              acks.push_back(std::make_shared<Ref>());
              acks.back()->set_ack_mode(true);
            }
            std::string payload(data);
            std::string missing;
            {
              py::gil_scoped_release rel;
              DoorbellBatch batch;  // sleeping receivers are woken together, with ONE futex call
              for (size_t i = 0; i < dests.size(); ++i) {
                std::string copy = i + 1 == dests.size() ? std::move(payload) : payload;
                if (!t.send(dests[i], up, down, std::move(copy), is_error, acks[i])) {
                  if (missing.empty()) missing = dests[i];
                  acks[i]->set_raw("unknown destination party '" + dests[i] + "'", true);
                }
              }
            }
            if (!missing.empty()) throw py::key_error("unknown destination party '" + missing + "'");
            return acks;
          },
          py::arg("dests"), py::arg("up"), py::arg("down"), py::arg("data"), py::arg("is_error") = false)
      .def("set_max_message_bytes", &SocketTransport::set_max_message_bytes)
      .def("max_message_bytes", &SocketTransport::max_message_bytes)
      .def(
          "drain",
          [](SocketTransport& t, double timeout) {
            py::gil_scoped_release rel;
            return t.drain(timeout);
          },
          py::arg("timeout") = 60.0)
      .def("stop",
           [](SocketTransport& t) {
             py::gil_scoped_release rel;
             t.stop();
           })
      .def("stats", [](SocketTransport& t) {
        const TransportStats& s = t.stats();
        py::dict d;
        d["send_op_count"] = s.send_op_count.load();
        d["receive_op_count"] = s.receive_op_count.load();
        d["bytes_sent"] = s.bytes_sent.load();
        d["bytes_received"] = s.bytes_received.load();
        d["rejected"] = s.rejected.load();
        d["shm_conns"] = s.shm_conns.load();
        d["paged_conns"] = s.paged_conns.load();
        return d;
      });
}
