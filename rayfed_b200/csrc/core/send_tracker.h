// SendTracker — bookkeeping behind the CleanupManager (fed/cleanup.py:29-203 in the reference,
// where one polling thread calls ray.get on every send).
//
// Every cross-party send is watched until its outcome is known.  The common outcome — the
// peer acknowledged — arrives as a raw frame on a transport thread; it is retired right there
// with two atomic operations and NO interpreter involvement: a transport thread that takes the
// GIL per acknowledgement steals it from the driver thread in the middle of its `.remote()`
// calls, which is what limited `many_tiny_tasks` at N parties.  Only failures (and outcomes
// that are Python objects: generic proxies) call back into Python.
#pragma once

#include <pybind11/pybind11.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "ref.h"

namespace fedcore {

class SendTracker : public std::enable_shared_from_this<SendTracker> {
 public:
  // handler(ref, pack, is_error): called with the GIL for every send that did not resolve to a
  // plain acknowledgement; it decides (and later calls retire()).
  explicit SendTracker(py::object handler) : handler_(std::move(handler)) {}
  ~SendTracker() {
    safe_release(handler_);
    for (auto& o : graveyard_) safe_release(o);
  }

  // GIL held.  `pack` is opaque (what the failure path needs: ref, destination, seq ids).
  void watch(RefPtr ref, py::object pack, bool is_error) {
    bury();  // drop what the GIL-free path could not
    auto e = std::make_shared<Entry>();
    e->ref = ref;
    e->pack = std::move(pack);
    e->is_error = is_error;
    pending_.fetch_add(1, std::memory_order_acq_rel);
    if (is_error) pending_err_.fetch_add(1, std::memory_order_acq_rel);
    {
      std::lock_guard<std::mutex> lk(mu_);
      e->id = next_id_++;
      watched_[e->id] = ref;
    }
    auto self = shared_from_this();
    ref->on_done([self, e] { self->settled(e); });
  }

  // A send whose outcome was handled (success on the slow path, or failure processed).
  void retire(bool is_error) {
    pending_.fetch_sub(1, std::memory_order_acq_rel);
    if (is_error) pending_err_.fetch_sub(1, std::memory_order_acq_rel);
    if (draining_.load(std::memory_order_acquire)) {
      std::lock_guard<std::mutex> lk(mu_);
      cv_.notify_all();
    }
  }

  long pending() const { return pending_.load(std::memory_order_acquire); }
  long pending_errors() const { return pending_err_.load(std::memory_order_acquire); }
  long fast_retired() const { return fast_.load(std::memory_order_relaxed); }
  long slow_settled() const { return slow_.load(std::memory_order_relaxed); }

  // Somebody waits from now on: completions must notify.
  void set_draining(bool v) { draining_.store(v, std::memory_order_release); }

  // True when (all | error) sends are retired.  GIL released by the binding.
  bool wait_idle(double timeout_s, bool errors_only) {
    auto idle = [&] { return (errors_only ? pending_errors() : pending()) <= 0; };
    std::unique_lock<std::mutex> lk(mu_);
    if (timeout_s < 0) {
      cv_.wait(lk, idle);
      return true;
    }
    return cv_.wait_for(lk, std::chrono::duration<double>(timeout_s), idle);
  }

  // A watched future that is resolved while its settle step has not finished yet (the window
  // between "the future resolved" and "its failure was queued").
  bool settling() {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : watched_)
      if (kv.second->done()) return true;
    return false;
  }

  // GIL held: release deferred Python references.
  void bury() {
    std::vector<py::object> dead;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dead.swap(graveyard_);
    }
    dead.clear();
  }

 private:
  struct Entry {
    uint64_t id = 0;
    RefPtr ref;
    py::object pack;
    bool is_error = false;
    ~Entry() { safe_release(pack); }
  };

  void unwatch(uint64_t id) {
    std::lock_guard<std::mutex> lk(mu_);
    watched_.erase(id);
  }

  // Any thread; no GIL assumed.
  void settled(const std::shared_ptr<Entry>& e) {
    if (e->ref->raw_ack_ok()) {
      {
        std::lock_guard<std::mutex> lk(mu_);
        watched_.erase(e->id);
        graveyard_.push_back(std::move(e->pack));  // decref later, under somebody's GIL
      }
      fast_.fetch_add(1, std::memory_order_relaxed);
      retire(e->is_error);
      return;
    }
    slow_.fetch_add(1, std::memory_order_relaxed);
    if (!interpreter_alive()) {
      unwatch(e->id);
      retire(e->is_error);
      return;
    }
    py::gil_scoped_acquire gil;
    py::object pack = std::move(e->pack);
    try {
      handler_(py::cast(e->ref), pack, e->is_error);
    } catch (py::error_already_set& err) {
      err.discard_as_unraisable("fedcore.SendTracker handler");
      retire(e->is_error);
    }
    unwatch(e->id);  // after the handler: stop() treats "resolved but still watched" as settling
  }

  py::object handler_;
  std::atomic<long> pending_{0}, pending_err_{0}, fast_{0}, slow_{0};
  std::atomic<bool> draining_{false};
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t next_id_ = 1;
  std::unordered_map<uint64_t, RefPtr> watched_;
  std::vector<py::object> graveyard_;
};

}  // namespace fedcore
