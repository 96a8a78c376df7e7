// fedcore::Ref — the in-process future that replaces Ray's ObjectRef.
//
// The reference leans on Ray ObjectRefs for every DAG edge
// (fed/fed_object.py:55, fed/api.py:597).  Here an edge is a `Ref`: a
// one-shot promise/future owned by the party's own runtime.  A Ref resolves to
//   * a Python object (task result),
//   * a raw byte string delivered by a transport thread, decoded lazily by the
//     consumer thread (so unpickling / device-side unpack runs on the thread —
//     and CUDA stream — that consumes the value), or
//   * an error (Python exception instance).
// Continuations are C++ callbacks, so dependency resolution never needs the
// GIL until a task body actually runs.
#pragma once
#include <pybind11/pybind11.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <thread>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "device_hooks.h"

namespace py = pybind11;

namespace fedcore {

enum class RefState : int { kPending = 0, kValue = 1, kRaw = 2, kError = 3 };

// Adaptive wait: spin briefly before parking on a condition variable. A futex
// wake costs tens of microseconds; a tiny fed task costs less than that.
inline int spin_budget_us() {
  static int v = [] {
    const char* e = std::getenv("RAYFED_B200_SPIN_US");
    return e ? std::atoi(e) : 40;
  }();
  return v;
}

// ---- GIL hand-off heuristic ------------------------------------------------------------
// A worker that is notified while the driver thread is still executing Python (between its
// `.remote()` calls and the `fed.get` that follows) would block in take_gil(), i.e. go to
// sleep in the kernel on the GIL condition variable, and pay a futex wake-up (10-30 us on a
// VM) when the driver finally parks.  We know when Python threads are parked in OUR waits
// (they release the GIL there): `g_native_waiters` counts them.  Before taking the GIL a
// worker spins in user space for a bounded time while nobody is parked; the moment the
// driver enters its wait the GIL is free and the acquisition is the uncontended fast path.
// The spin adapts: after repeated timeouts (the driver is busy elsewhere: sleeping, I/O,
// long Python compute) it is skipped except for an occasional probe.
inline std::atomic<int> g_native_waiters{0};
inline std::atomic<int> g_gil_spin_misses{0};

struct NativeWaitScope {
  NativeWaitScope() { g_native_waiters.fetch_add(1, std::memory_order_release); }
  ~NativeWaitScope() { g_native_waiters.fetch_sub(1, std::memory_order_release); }
};

inline int gil_spin_budget_us() {
  static int v = [] {
    const char* e = std::getenv("RAYFED_B200_GIL_SPIN_US");
    return e ? std::atoi(e) : 80;
  }();
  return v;
}

inline void cpu_relax();

inline void pre_gil_spin() {
  if (g_native_waiters.load(std::memory_order_acquire) > 0) return;
  const int budget = gil_spin_budget_us();
  if (budget <= 0) return;
  int misses = g_gil_spin_misses.load(std::memory_order_relaxed);
  if (misses > 8 && (misses & 63) != 0) {  // heuristic is not paying off: only probe now and then
    g_gil_spin_misses.store(misses + 1, std::memory_order_relaxed);
    return;
  }
  auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    for (int i = 0; i < 64; ++i) {
      if (g_native_waiters.load(std::memory_order_acquire) > 0) {
        g_gil_spin_misses.store(0, std::memory_order_relaxed);
        return;
      }
      cpu_relax();
    }
    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(budget)) break;
  }
  g_gil_spin_misses.store(misses + 1, std::memory_order_relaxed);
}

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

template <typename Pred>
inline bool spin_until(Pred&& pred) {
  int budget = spin_budget_us();
  if (budget <= 0) return pred();
  auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    if (pred()) return true;
    for (int i = 0; i < 32; ++i) cpu_relax();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(budget)) return pred();
  }
}

// Adaptive spinning for waits on a Ref (the driver in fed.get, an actor lane parked on its
// inputs): when this thread's recent waits were short, spin about that long before falling back
// to the futex — the kernel wake-up it avoids costs more than the wait itself in tiny-task
// loops.  Long waits (GPU work, slow peers) push the average up and the spin falls back to the
// base budget, so nothing burns a core while real work is pending.  OFF by default
// (RAYFED_B200_ADAPTIVE_SPIN_US = cap in microseconds, 0 = off): to be tuned on the target host.
inline int adaptive_spin_cap_us() {
  static int v = [] {
    const char* e = std::getenv("RAYFED_B200_ADAPTIVE_SPIN_US");
    return e ? std::atoi(e) : 0;
  }();
  return v;
}

template <typename Pred>
inline bool spin_for(Pred&& pred, int budget_us) {
  if (budget_us <= 0) return pred();
  auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    if (pred()) return true;
    for (int i = 0; i < 32; ++i) cpu_relax();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(budget_us)) return pred();
  }
}

inline bool interpreter_alive() {
  if (!Py_IsInitialized()) return false;
#if PY_VERSION_HEX >= 0x030D0000
  return !Py_IsFinalizing();
#else
  return !_Py_IsFinalizing();
#endif
}

// Drop a Python reference from a thread that may not hold the GIL; leak it if
// the interpreter is already going away (worker threads can outlive main).
inline void safe_release(py::object& o) {
  if (!o) return;
  if (!interpreter_alive()) {
    o.release();
    return;
  }
  py::gil_scoped_acquire gil;
  o = py::object();
}

// The device-side completion of a value: "everything the producer enqueued on `stream`
// before `event` was recorded".  Shared by all outputs of one task.
struct DeviceEvent {
  uint64_t event = 0;
  uint64_t stream = 0;  // the producer's stream (a consumer on the same stream needs no wait)
  bool pooled = false;  // ours (goes back to the pool) vs borrowed (`owner` keeps it alive)
  py::object owner;     // e.g. the torch.cuda.Event a transport recorded
  ~DeviceEvent() {
    if (pooled && event) EventPool::get().release(event);
    safe_release(owner);
  }
};
using DeviceEventPtr = std::shared_ptr<DeviceEvent>;

// Python-visible carrier "value + the CUDA event that orders it": what a transport's decoder
// returns for a payload that landed through the device channel, and what the sender hand-off
// receives.  A Ref completed with one keeps the plain value and adopts the event.
struct DeviceValue {
  py::object value;
  uint64_t event = 0;
  uint64_t stream = 0;
  py::object owner;
};

class Ref : public std::enable_shared_from_this<Ref> {
 public:
  Ref() = default;
  ~Ref() {
    // Python objects may only be released with the GIL held.
    safe_release(value_);
    safe_release(decoder_);
    for (auto& cb : py_callbacks_) safe_release(cb);
  }
  Ref(const Ref&) = delete;
  Ref& operator=(const Ref&) = delete;

  bool done() const { return state_.load(std::memory_order_acquire) != (int)RefState::kPending; }
  bool is_error() const {
    int s = state_.load(std::memory_order_acquire);
    return s == (int)RefState::kError || (s == (int)RefState::kRaw && raw_is_error_);
  }

  // ---- completion (callable from any thread) -------------------------------
  // GIL must be held by the caller for the py::object flavours.
  bool set_value(py::object v) {
    adopt_device_value(v);
    return complete(RefState::kValue, std::move(v), std::string(), false);
  }
  bool set_error(py::object exc) { return complete(RefState::kError, std::move(exc), std::string(), true); }
  // No GIL needed: transport threads hand over bytes.
  bool set_raw(std::string bytes, bool is_error) {
    return complete(RefState::kRaw, py::object(), std::move(bytes), is_error);
  }

  // Decoder applied (once, with the GIL) to raw bytes on first read.
  void set_decoder(py::object dec) { decoder_ = std::move(dec); }

  // ---- waiting --------------------------------------------------------------
  // Returns true when resolved. timeout < 0 => wait forever. Caller must NOT
  // hold the GIL (bindings release it).
  bool wait(double timeout_s) {
    if (done()) return true;
    const int cap = adaptive_spin_cap_us();
    if (cap <= 0) {
      if (timeout_s != 0 && spin_until([&] { return done(); })) return true;
      return block(timeout_s);
    }
    // adaptive: spin ~2x this thread's recent wait time, bounded by [base budget, cap]
    static thread_local double ema_us = 1e12;
    auto t0 = std::chrono::steady_clock::now();
    int budget = spin_budget_us();
    if (ema_us < (double)cap) budget = std::max(budget, std::min(cap, (int)(2.0 * ema_us) + 1));
    bool ok = timeout_s != 0 && spin_for([&] { return done(); }, budget);
    if (!ok) ok = block(timeout_s);
    double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    ema_us = ema_us > 1e11 ? us : 0.75 * ema_us + 0.25 * us;
    return ok;
  }

  bool block(double timeout_s) {
    std::unique_lock<std::mutex> lk(mu_);
    if (timeout_s < 0) {
      cv_.wait(lk, [&] { return done(); });
      return true;
    }
    return cv_.wait_for(lk, std::chrono::duration<double>(timeout_s), [&] { return done(); });
  }

  // C++ continuation; runs immediately if already resolved. Must not assume GIL.
  void on_done(std::function<void()> cb) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!done()) {
        callbacks_.push_back(std::move(cb));
        return;
      }
    }
    cb();
  }

  // Python continuation `cb(ref)`; GIL held by caller.
  void add_done_callback(py::object cb) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!done()) {
        py_callbacks_.push_back(std::move(cb));
        return;
      }
    }
    cb(py::cast(shared_from_this()));
  }

  // ---- reading (GIL held) ---------------------------------------------------
  // Materialise raw bytes into a Python value via the decoder (idempotent).
  void materialize() {
    if (state_.load(std::memory_order_acquire) != (int)RefState::kRaw) return;
    // Take decode_mu_ with the GIL dropped: the holder runs Python (the decoder)
    // and may yield the GIL to us, so blocking here with the GIL would deadlock.
    std::unique_lock<std::mutex> lk(decode_mu_, std::defer_lock);
    {
      py::gil_scoped_release rel;
      lk.lock();
    }
    if (state_.load(std::memory_order_acquire) != (int)RefState::kRaw) return;
    py::bytes b(raw_);
    py::object out;
    bool err = raw_is_error_;
    try {
      out = decoder_ ? decoder_(b, err) : py::object(b);
    } catch (py::error_already_set& e) {
      // decoding failed (e.g. whitelist violation): surface as the error value
      out = py::reinterpret_borrow<py::object>(e.value());
      err = true;
    }
    // A decoder may itself hand back an exception instance for error frames.
    if (!err && PyExceptionInstance_Check(out.ptr()) && decoded_exc_is_error_) err = true;
    if (!err) adopt_device_value(out);
    value_ = std::move(out);
    std::string().swap(raw_);
    state_.store(err ? (int)RefState::kError : (int)RefState::kValue, std::memory_order_release);
  }

  // Returns value or raises the stored exception.
  py::object result() {
    materialize();
    int s = state_.load(std::memory_order_acquire);
    if (s == (int)RefState::kError) {
      py::object exc = value_;
      if (PyExceptionInstance_Check(exc.ptr())) {
        PyErr_SetObject((PyObject*)Py_TYPE(exc.ptr()), exc.ptr());
      } else {
        PyErr_SetObject(PyExc_RuntimeError, exc.ptr());
      }
      throw py::error_already_set();
    }
    if (s != (int)RefState::kValue) throw std::runtime_error("Ref is not resolved");
    return value_;
  }

  // The stored exception (or None).
  py::object exception() {
    materialize();
    if (state_.load(std::memory_order_acquire) == (int)RefState::kError) return value_;
    return py::none();
  }

  void set_decoded_exc_is_error(bool v) { decoded_exc_is_error_ = v; }

  // ---- device ordering -------------------------------------------------------
  // Attach BEFORE completing the Ref (the executor does) or from the decoder.
  void set_device_event(DeviceEventPtr e) { std::atomic_store(&dev_event_, std::move(e)); }
  DeviceEventPtr device_event() const { return std::atomic_load(&dev_event_); }
  // Host-side read barrier: block until the producer's device work is complete. No GIL needed.
  bool device_sync() {
    DeviceEventPtr e = device_event();
    const DeviceHooks* h = device_hooks();
    if (!e || !e->event || !h) return true;
    return h->event_synchronize(e->event) == 0;
  }

 private:
  bool complete(RefState st, py::object v, std::string raw, bool is_err) {
    std::vector<std::function<void()>> cbs;
    std::vector<py::object> pycbs;
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (done()) return false;
      value_ = std::move(v);
      raw_ = std::move(raw);
      raw_is_error_ = is_err;
      state_.store((int)st, std::memory_order_release);
      cbs.swap(callbacks_);
      pycbs.swap(py_callbacks_);
    }
    cv_.notify_all();
    for (auto& cb : cbs) cb();
    if (!pycbs.empty()) {
      py::gil_scoped_acquire gil;
      py::object self = py::cast(shared_from_this());
      for (auto& cb : pycbs) {
        try {
          cb(self);
        } catch (py::error_already_set& e) {
          e.discard_as_unraisable("fedcore.Ref done-callback");
        }
      }
      pycbs.clear();
    }
    return true;
  }

  // GIL held. `v` may be a DeviceValue: keep the plain value, adopt the event.
  void adopt_device_value(py::object& v) {
    if (!v || !py::isinstance<DeviceValue>(v)) return;
    DeviceValue& dv = v.cast<DeviceValue&>();
    if (dv.event) {
      auto e = std::make_shared<DeviceEvent>();
      e->event = dv.event;
      e->stream = dv.stream;
      e->owner = dv.owner;
      set_device_event(std::move(e));
    }
    py::object inner = dv.value;
    v = std::move(inner);
  }

  std::atomic<int> state_{(int)RefState::kPending};
  DeviceEventPtr dev_event_;
  std::mutex mu_;
  std::mutex decode_mu_;
  std::condition_variable cv_;
  py::object value_;
  py::object decoder_;
  std::string raw_;
  bool raw_is_error_ = false;
  bool decoded_exc_is_error_ = false;
  std::vector<std::function<void()>> callbacks_;
  std::vector<py::object> py_callbacks_;
};

using RefPtr = std::shared_ptr<Ref>;

}  // namespace fedcore
