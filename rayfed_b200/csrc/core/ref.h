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

#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
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

// Adaptive spinning.  A futex / poll wake-up costs 10-50 us — more than a tiny fed task — and
// the thread that has to ring the sleeper pays a system call on top.  So a waiter whose recent
// waits were SHORT keeps spinning for about that long before it parks: the driver in fed.get, an
// actor lane between two calls, a shared-memory ring reader between two frames.  Long waits (GPU
// work, slow peers, a training step between messages) push the average above the cap and the
// waiter falls back to the base budget, so nothing burns a core unless messages really arrive
// every few hundred microseconds.  The number of threads that may spin long at once is bounded
// per process.  RAYFED_B200_ADAPTIVE_SPIN_US = cap in microseconds (0 = off) and
// RAYFED_B200_MAX_SPINNERS = concurrent long spinners per process override the policy below.

// CPUs this process may really use: the affinity mask, cut down by a cgroup CPU quota if any.
inline int usable_cpus() {
  static int v = [] {
    int n = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
      int c = CPU_COUNT(&set);
      if (c > 0) n = c;
    }
    // cgroup v2: "max 100000" or "<quota> <period>"; cgroup v1: cfs_quota_us / cfs_period_us
    auto quota_cpus = [](const char* path_q, const char* path_p) -> double {
      FILE* f = std::fopen(path_q, "r");
      if (!f) return -1.0;
      char a[64] = {0}, b[64] = {0};
      int got = std::fscanf(f, "%63s %63s", a, b);
      std::fclose(f);
      if (got < 1 || a[0] == 'm') return -1.0;  // "max"
      double q = std::atof(a), p = 0;
      if (path_p) {
        FILE* g = std::fopen(path_p, "r");
        if (!g) return -1.0;
        if (std::fscanf(g, "%63s", b) != 1) b[0] = 0;
        std::fclose(g);
      }
      p = std::atof(b);
      return (q > 0 && p > 0) ? q / p : -1.0;
    };
    double q = quota_cpus("/sys/fs/cgroup/cpu.max", nullptr);
    if (q < 0) q = quota_cpus("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us");
    if (q > 0 && q < n) n = std::max(1, (int)q);
#endif
    return std::max(1, n);
  }();
  return v;
}

// Policy.  Environment variables win; otherwise long spins are OFF until somebody who knows the
// deployment switches them on: `fed.init` does (set_spin_policy), from the CPUs this process may
// use and the number of parties that share the host.  Measured on the B200 host (128 hardware
// threads behind a 16-CPU cgroup quota), many_tiny_tasks per iteration, off -> on:
// 2 parties 164-193 -> 44-80 us, 4 parties 101-122 -> 69-113 us, 8 parties 181-186 -> 222-241 us
// (16 spinning threads eat the quota) — hence "at least two spinners' worth of CPU per party".
// Spinning also keeps the 2-party case out of the deep-idle lottery: with sleeping threads the
// same binary ran at 78 and at 521 us per iteration (profiles/README.md R2.2b).
inline std::atomic<int> g_adaptive_cap_us{-1};   // -1: not configured
inline std::atomic<int> g_max_long_spinners{-1};

inline int adaptive_spin_cap_us() {
  static const int env = [] {
    const char* e = std::getenv("RAYFED_B200_ADAPTIVE_SPIN_US");
    return e ? std::max(0, std::atoi(e)) : -1;
  }();
  if (env >= 0) return env;
  int v = g_adaptive_cap_us.load(std::memory_order_relaxed);
  return v < 0 ? 0 : v;
}

inline int max_long_spinners() {
  static const int env = [] {
    const char* e = std::getenv("RAYFED_B200_MAX_SPINNERS");
    return e ? std::max(0, std::atoi(e)) : -1;
  }();
  if (env >= 0) return env;
  int v = g_max_long_spinners.load(std::memory_order_relaxed);
  return v < 0 ? 2 : v;
}

// `local_parties`: fed parties (processes) that share this host.  Each gets an equal share of
// half the usable CPUs for long spinners; with less than two per party the feature stays off.
inline void set_spin_policy_auto(int local_parties) {
  if (local_parties < 1) local_parties = 1;
  int per_party = usable_cpus() / (2 * local_parties);
  if (per_party >= 2) {
    g_max_long_spinners.store(std::min(8, per_party), std::memory_order_relaxed);
    g_adaptive_cap_us.store(500, std::memory_order_relaxed);
  } else {
    g_adaptive_cap_us.store(0, std::memory_order_relaxed);
  }
}

inline std::atomic<int> g_long_spinners{0};

// Self-limiting: a long spin only pays when there is a core to burn.  A spinner that notices a
// hole in its own time line (it was descheduled: more runnable threads than cores, or a cgroup
// CPU quota throttling the process) switches long spins off for the whole process for a while;
// holes in quick succession double the pause (2 ms .. 256 ms), an isolated one costs 2 ms.
inline std::atomic<int64_t> g_long_spin_off_until_ns{0};
inline std::atomic<int> g_long_spin_pause_ms{2};
inline std::atomic<long> g_long_spin_preemptions{0};

inline int64_t mono_ns() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

inline std::atomic<int64_t> g_long_spin_last_hole_ns{0};

inline void long_spin_preempted() {
  const int64_t now = mono_ns();
  const int64_t prev = g_long_spin_last_hole_ns.exchange(now, std::memory_order_relaxed);
  int pause = g_long_spin_pause_ms.load(std::memory_order_relaxed);
  // an isolated hole (hypervisor steal, a stray interrupt storm) costs a short pause; holes in
  // quick succession mean the cores are oversubscribed: escalate
  pause = (now - prev < (int64_t)20 * pause * 1000000) ? std::min(256, pause * 2) : 2;
  g_long_spin_pause_ms.store(pause, std::memory_order_relaxed);
  g_long_spin_off_until_ns.store(now + (int64_t)pause * 1000000, std::memory_order_relaxed);
  g_long_spin_preemptions.fetch_add(1, std::memory_order_relaxed);
}

// RAII claim on one of the process's long-spin slots.
struct LongSpinSlot {
  bool ok;
  LongSpinSlot() {
    ok = g_long_spinners.fetch_add(1, std::memory_order_acq_rel) < max_long_spinners() &&
         mono_ns() >= g_long_spin_off_until_ns.load(std::memory_order_relaxed);
    if (!ok) g_long_spinners.fetch_sub(1, std::memory_order_acq_rel);
  }
  ~LongSpinSlot() {
    if (ok) g_long_spinners.fetch_sub(1, std::memory_order_acq_rel);
  }
  LongSpinSlot(const LongSpinSlot&) = delete;
  LongSpinSlot& operator=(const LongSpinSlot&) = delete;
};

// Spin on `pred` for up to `budget_us`, giving up (and pausing long spins process-wide) when this
// thread was descheduled in between.  Returns pred's last value.
template <typename Pred>
inline bool spin_long(Pred&& pred, int budget_us) {
  constexpr int64_t kHoleNs = 200 * 1000;  // above any interrupt, below any scheduler time slice
  const int64_t t0 = mono_ns();
  int64_t last = t0;
  for (;;) {
    if (pred()) return true;
    for (int i = 0; i < 32; ++i) cpu_relax();
    int64_t now = mono_ns();
    if (now - last > kHoleNs) {
      long_spin_preempted();
      return pred();
    }
    last = now;
    if (now - t0 > (int64_t)budget_us * 1000) return pred();
  }
}

// Per-waiter state (one thread at a time): exponential average of recent wait durations.
struct AdaptiveSpin {
  double ema_us = 1e12;
  // How long to spin before parking, given the history.
  int budget_us() const {
    const int base = spin_budget_us();
    const int cap = adaptive_spin_cap_us();
    if (cap <= 0 || base <= 0 || ema_us >= (double)cap) return base;
    return std::max(base, std::min(cap, (int)(2.0 * ema_us) + 1));
  }
  void observe(double waited_us) { ema_us = ema_us > 1e11 ? waited_us : 0.75 * ema_us + 0.25 * waited_us; }
};

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

// spin (adaptively) on `pred`, then `park()` — which must return pred's final value.
template <typename Pred, typename Park>
inline bool adaptive_wait(AdaptiveSpin& hist, Pred&& pred, Park&& park) {
  const int base = spin_budget_us();
  int budget = hist.budget_us();
  auto t0 = std::chrono::steady_clock::now();
  bool ok;
  if (budget > base) {
    LongSpinSlot slot;
    ok = slot.ok ? spin_long(pred, budget) : spin_for(pred, base);
  } else {
    ok = spin_for(pred, budget);
  }
  if (!ok) ok = park();
  hist.observe(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
  return ok;
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
    if (timeout_s == 0) return false;
    static thread_local AdaptiveSpin hist;
    return adaptive_wait(hist, [&] { return done(); }, [&] { return block(timeout_s); });
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
      if (!decoder_ && ack_mode_) {
        // transport acknowledgement, decoded natively: True, or the failure text as an error
        if (err) {
          out = py::reinterpret_steal<py::object>(PyObject_CallOneArg(PyExc_RuntimeError, py::str(raw_).ptr()));
          if (!out) throw py::error_already_set();
        } else {
          out = py::bool_(true);
        }
      } else {
        out = decoder_ ? decoder_(b, err) : py::object(b);
      }
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

  // Acknowledgement futures of the native transport: the raw frame decodes to True (or to a
  // RuntimeError carrying the failure text) without a Python decoder object — so such a Ref can
  // be completed, forwarded, watched and destroyed on transport threads without the GIL.
  void set_ack_mode(bool v) { ack_mode_ = v; }
  bool ack_mode() const { return ack_mode_; }
  // No GIL needed. True: resolved to an un-decoded, non-error acknowledgement frame.
  bool raw_ack_ok() const {
    return ack_mode_ && state_.load(std::memory_order_acquire) == (int)RefState::kRaw && !raw_is_error_;
  }

  // Complete `dst` with this Ref's outcome as soon as there is one.  A native continuation: a raw
  // outcome (bytes from a transport thread) is handed over without touching the interpreter; only
  // Python-object outcomes take the GIL.  Call with the GIL held (the decoder is shared here).
  void forward_to(std::shared_ptr<Ref> dst) {
    if (!dst || dst.get() == this) return;
    if (decoder_) dst->decoder_ = decoder_;
    dst->ack_mode_ = ack_mode_;
    dst->decoded_exc_is_error_ = decoded_exc_is_error_;
    auto self = shared_from_this();
    on_done([self, dst] { self->copy_outcome_to(*dst); });
  }

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

  // Any thread, GIL held or not.
  void copy_outcome_to(Ref& dst) {
    {
      // vs. a concurrent materialize(): it holds decode_mu_ while running Python, so never
      // block on it with the GIL held
      std::unique_lock<std::mutex> lk(decode_mu_, std::defer_lock);
      if (Py_IsInitialized() && PyGILState_Check()) {
        py::gil_scoped_release rel;
        lk.lock();
      } else {
        lk.lock();
      }
      if (state_.load(std::memory_order_acquire) == (int)RefState::kRaw) {
        std::string raw = raw_;
        bool e = raw_is_error_;
        lk.unlock();
        dst.set_raw(std::move(raw), e);
        return;
      }
    }
    if (!interpreter_alive()) return;
    py::gil_scoped_acquire gil;
    if (DeviceEventPtr ev = device_event()) dst.set_device_event(ev);
    if (state_.load(std::memory_order_acquire) == (int)RefState::kError) {
      dst.set_error(value_);
    } else {
      dst.complete(RefState::kValue, value_, std::string(), false);
    }
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
  bool ack_mode_ = false;
  std::vector<std::function<void()>> callbacks_;
  std::vector<py::object> py_callbacks_;
};

using RefPtr = std::shared_ptr<Ref>;

}  // namespace fedcore
