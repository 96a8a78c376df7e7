// fedcore::Executor — the per-party task scheduler that replaces what the
// reference borrowed from Ray core: task submission with dependency gating
// (an ObjectRef passed as a top-level argument delays the task until it
// resolves — fed/api.py:413-416, fed/proxy/barriers.py:478-483), ordered actor
// mailboxes with `max_concurrency` (fed/_private/fed_actor.py:86-112) and
// `ray.kill` (fed/api.py:622-623).
//
// Design: every task is a node whose inputs are `Ref`s.  Stateless tasks are
// released to the shared pool lane the moment their last input resolves
// (continuations are C++ callbacks — no polling thread, no GIL).  Actor tasks
// go to the actor's own FIFO lane at submit time so mailbox order is the
// program order; the lane's thread parks on the head task's inputs without the
// GIL.  The GIL is only taken to run the Python body.
#pragma once
#include <pybind11/pybind11.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "ref.h"

namespace fedcore {

struct Task {
  py::object fn;
  py::object args;    // tuple
  py::object kwargs;  // dict or None
  std::vector<RefPtr> outs;
  std::vector<RefPtr> deps;
  std::atomic<int> pending{0};
  ~Task() {
    safe_release(fn);
    safe_release(args);
    safe_release(kwargs);
  }
};
using TaskPtr = std::shared_ptr<Task>;

struct ExecStats {
  std::atomic<uint64_t> submitted{0};
  std::atomic<uint64_t> completed{0};
  std::atomic<uint64_t> failed{0};
};

class Executor;

// One FIFO of runnable (pool) or submitted (actor) tasks plus its threads.
struct Lane : public std::enable_shared_from_this<Lane> {
  std::string name;
  bool ordered = false;  // actor mailbox: wait for inputs in FIFO order
  std::mutex mu;
  std::condition_variable cv;
  std::deque<TaskPtr> q;
  bool stopping = false;
  bool killed = false;
  py::object kill_error;  // error handed to tasks submitted after kill
  std::atomic<int> live_threads{0};
  std::atomic<int> busy{0};
  std::atomic<int> qsize{0};     // mirrors q.size() for lock-free spinning
  std::atomic<bool> wake{false};  // set on stop/kill
  // Device lanes: `thread_init()` runs once per worker thread (GIL held) — it binds the thread to
  // the party's CUDA device, makes a stream of its own the thread's current stream and returns the
  // stream handle.  Bumping `init_gen` makes running workers re-initialise before their next task
  // (fed.init sets the device after the pool threads exist).
  py::object thread_init;
  std::atomic<uint64_t> init_gen{0};
  ~Lane() {
    safe_release(kill_error);
    safe_release(thread_init);
  }
};
using LanePtr = std::shared_ptr<Lane>;

class Executor : public std::enable_shared_from_this<Executor> {
 public:
  Executor(int pool_threads, py::object wrap_error) : wrap_error_(std::move(wrap_error)) {
    auto pool = std::make_shared<Lane>();
    pool->name = "pool";
    pool->ordered = false;
    lanes_.push_back(pool);
    stats_ = std::make_shared<ExecStats>();
    spawn(pool, pool_threads);
  }
  ~Executor() { safe_release(wrap_error_); }

  int create_lane(const std::string& name, int concurrency, py::object thread_init = py::none()) {
    auto lane = std::make_shared<Lane>();
    lane->name = name;
    lane->ordered = true;
    if (!thread_init.is_none()) {
      lane->thread_init = std::move(thread_init);
      lane->init_gen.store(1);
    }
    int id;
    {
      std::lock_guard<std::mutex> lk(mu_);
      lanes_.push_back(lane);
      id = (int)lanes_.size() - 1;
    }
    spawn(lane, concurrency < 1 ? 1 : concurrency);
    return id;
  }

  // GIL held. Returns `num_returns` Refs (>=1).
  std::vector<RefPtr> submit(int lane_id, py::object fn, py::tuple args, py::object kwargs,
                             int num_returns) {
    LanePtr lane = get_lane(lane_id);
    auto t = std::make_shared<Task>();
    t->fn = std::move(fn);
    t->args = args;
    t->kwargs = std::move(kwargs);
    if (num_returns < 1) num_returns = 1;
    for (int i = 0; i < num_returns; ++i) t->outs.push_back(std::make_shared<Ref>());
    for (auto item : args) collect_dep(item, t->deps);
    if (t->kwargs && !t->kwargs.is_none()) {
      for (auto kv : py::reinterpret_borrow<py::dict>(t->kwargs)) collect_dep(kv.second, t->deps);
    }
    stats_->submitted.fetch_add(1, std::memory_order_relaxed);
    std::vector<RefPtr> outs = t->outs;

    {
      std::unique_lock<std::mutex> lk(lane->mu);
      if (lane->killed || lane->stopping) {
        py::object err = lane->kill_error;
        lk.unlock();
        fail_task(t, err ? err : py::cast(std::string("lane '") + lane->name + "' is not running"));
        return outs;
      }
      if (lane->ordered) {
        lane->q.push_back(t);
        lane->qsize.fetch_add(1, std::memory_order_release);
        lk.unlock();
        lane->cv.notify_one();
        return outs;
      }
    }
    // Pool lane: release when the last input resolves.
    int n = (int)t->deps.size();
    if (n == 0) {
      enqueue(lane, t);
    } else {
      t->pending.store(n, std::memory_order_relaxed);
      for (auto& d : t->deps) {
        std::weak_ptr<Lane> wl = lane;
        d->on_done([t, wl]() {
          if (t->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            if (auto l = wl.lock()) enqueue(l, t);
          }
        });
      }
    }
    return outs;
  }

  // GIL held. (Re)bind a lane's worker threads to a device stream; None detaches.
  void set_lane_thread_init(int lane_id, py::object thread_init) {
    LanePtr lane = get_lane(lane_id);
    std::lock_guard<std::mutex> lk(lane->mu);
    lane->thread_init = thread_init.is_none() ? py::object() : std::move(thread_init);
    lane->init_gen.fetch_add(1);
  }

  // Drop an actor's mailbox: queued tasks fail with `error`; later submits too.
  void kill_lane(int lane_id, py::object error) {
    LanePtr lane = get_lane(lane_id);
    std::deque<TaskPtr> drained;
    {
      std::lock_guard<std::mutex> lk(lane->mu);
      lane->killed = true;
      lane->wake.store(true);
      lane->kill_error = error;
      drained.swap(lane->q);
      lane->qsize.store(0);
    }
    lane->cv.notify_all();
    for (auto& t : drained) fail_task(t, error);
  }

  // GIL must be released by the caller when wait=true.
  void shutdown(bool wait, double timeout_s) {
    std::vector<LanePtr> lanes;
    {
      std::lock_guard<std::mutex> lk(mu_);
      lanes = lanes_;
    }
    for (auto& l : lanes) {
      {
        std::lock_guard<std::mutex> lk(l->mu);
        l->stopping = true;
        l->wake.store(true);
      }
      l->cv.notify_all();
    }
    if (!wait) return;
    auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s);
    for (auto& l : lanes) {
      while (l->live_threads.load() > 0 && std::chrono::steady_clock::now() < deadline)
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
  }

  uint64_t submitted() const { return stats_->submitted.load(); }
  uint64_t completed() const { return stats_->completed.load(); }
  uint64_t failed() const { return stats_->failed.load(); }
  int num_lanes() {
    std::lock_guard<std::mutex> lk(mu_);
    return (int)lanes_.size();
  }
  int lane_backlog(int lane_id) {
    LanePtr lane = get_lane(lane_id);
    std::lock_guard<std::mutex> lk(lane->mu);
    return (int)lane->q.size() + lane->busy.load();
  }

 private:
  static void collect_dep(py::handle item, std::vector<RefPtr>& deps) {
    if (py::isinstance<Ref>(item)) deps.push_back(item.cast<RefPtr>());
  }

  static void enqueue(const LanePtr& lane, const TaskPtr& t) {
    {
      std::lock_guard<std::mutex> lk(lane->mu);
      lane->q.push_back(t);
      lane->qsize.fetch_add(1, std::memory_order_release);
    }
    lane->cv.notify_one();
  }

  LanePtr get_lane(int id) {
    std::lock_guard<std::mutex> lk(mu_);
    if (id < 0 || id >= (int)lanes_.size()) throw std::out_of_range("no such lane");
    return lanes_[id];
  }

  // GIL held.
  void fail_task(const TaskPtr& t, py::object err) {
    for (auto& o : t->outs) o->set_error(err);
    stats_->failed.fetch_add(1, std::memory_order_relaxed);
  }

  void spawn(const LanePtr& lane, int n) {
    for (int i = 0; i < n; ++i) {
      lane->live_threads.fetch_add(1);
      std::shared_ptr<ExecStats> stats = stats_;
      // The wrap_error callable is borrowed for the executor's life; threads keep
      // their own strong ref so a late task never touches a dead Executor.
      py::object wrap = wrap_error_;
      std::thread([lane, stats, wrap]() mutable {
        worker(lane, stats, wrap);
        safe_release(wrap);
        lane->live_threads.fetch_sub(1);
      }).detach();
    }
  }

  static void worker(LanePtr lane, std::shared_ptr<ExecStats> stats, py::object& wrap) {
    // One persistent Python thread state per worker: creating/destroying a
    // tstate around every task body costs far more than the body of a tiny task.
    if (!interpreter_alive()) return;
    PyGILState_STATE gstate = PyGILState_Ensure();
    PyThreadState* saved = PyEval_SaveThread();
    worker_loop(lane, stats, wrap);
    if (interpreter_alive()) {
      PyEval_RestoreThread(saved);
      PyGILState_Release(gstate);
    }
  }

  // Block until every input is resolved.  A fan-in (the aggregator of N parties) used to wait for
  // its inputs one after the other — up to N sleeps and N wake-ups when the arrivals are spread
  // out; with several inputs outstanding the lane thread now sleeps ONCE, on a countdown latch the
  // inputs' native continuations release.
  static void wait_inputs(const std::vector<RefPtr>& deps) {
    size_t outstanding = 0;
    for (auto& d : deps)
      if (!d->done()) ++outstanding;
    if (outstanding == 0) return;
    if (outstanding == 1) {
      for (auto& d : deps) d->wait(-1.0);
      return;
    }
    struct Latch {
      std::atomic<int> left{0};
      std::mutex mu;
      std::condition_variable cv;
    };
    auto latch = std::make_shared<Latch>();
    latch->left.store((int)deps.size(), std::memory_order_relaxed);
    for (auto& d : deps) {
      d->on_done([latch] {
        if (latch->left.fetch_sub(1, std::memory_order_acq_rel) == 1) {
          std::lock_guard<std::mutex> lk(latch->mu);  // pairs with the predicate check of the waiter
          latch->cv.notify_all();
        }
      });
    }
    static thread_local AdaptiveSpin hist;
    adaptive_wait(
        hist, [&] { return latch->left.load(std::memory_order_acquire) <= 0; },
        [&] {
          std::unique_lock<std::mutex> lk(latch->mu);
          latch->cv.wait(lk, [&] { return latch->left.load(std::memory_order_acquire) <= 0; });
          return true;
        });
  }

  static void worker_loop(LanePtr& lane, std::shared_ptr<ExecStats>& stats, py::object& wrap) {
    AdaptiveSpin idle;  // a lane that is called every few hundred us stays awake between calls
    for (;;) {
      TaskPtr t;
      auto idle_t0 = std::chrono::steady_clock::now();
      {
        auto pred = [&] { return lane->qsize.load(std::memory_order_acquire) > 0 || lane->wake.load(std::memory_order_relaxed); };
        const int base = spin_budget_us(), budget = idle.budget_us();
        if (budget > base) {
          LongSpinSlot slot;
          slot.ok ? spin_long(pred, budget) : spin_for(pred, base);
        } else {
          spin_for(pred, budget);
        }
      }
      {
        std::unique_lock<std::mutex> lk(lane->mu);
        lane->cv.wait(lk, [&] { return lane->stopping || lane->killed || !lane->q.empty(); });
        // the whole idle period (spin + park) drives the next spin budget
        idle.observe(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - idle_t0).count());
        if (lane->killed) return;
        if (lane->q.empty()) {
          if (lane->stopping) return;
          continue;
        }
        t = lane->q.front();
        lane->q.pop_front();
        lane->qsize.fetch_sub(1, std::memory_order_release);
        lane->busy.fetch_add(1);
      }
      if (lane->ordered) {
        // Actor mailbox: park on the inputs (no GIL held here).
        wait_inputs(t->deps);
      }
      if (!interpreter_alive()) return;
      if (lane->init_gen.load(std::memory_order_acquire) != tl_init_gen()) bind_thread(*lane);
      run(*t, *stats, wrap);
      lane->busy.fetch_sub(1);
      t.reset();
    }
  }

  // ---- device ordering (device_hooks.h) ---------------------------------------------------
  static uint64_t& tl_stream() {
    static thread_local uint64_t s = 0;
    return s;
  }
  static uint64_t& tl_init_gen() {
    static thread_local uint64_t g = 0;
    return g;
  }

  static void bind_thread(Lane& lane) {
    py::gil_scoped_acquire gil;
    py::object init;
    uint64_t gen;
    {
      std::lock_guard<std::mutex> lk(lane.mu);
      init = lane.thread_init;
      gen = lane.init_gen.load();
    }
    uint64_t stream = 0;
    if (init) {
      try {
        stream = init().cast<uint64_t>();
      } catch (py::error_already_set& e) {
        e.discard_as_unraisable("fedcore lane thread_init");
      }
    }
    tl_stream() = stream;
    tl_init_gen() = gen;
  }

  // Make this worker's stream wait for the device completion of every input that has one.
  // `seen` remembers which inputs were handled so the post-decode pass only adds new ones.
  static void wait_device_inputs(const std::vector<RefPtr>& deps, uint64_t stream, std::vector<const Ref*>& seen) {
    const DeviceHooks* h = device_hooks();
    if (!h || !stream) return;
    for (auto& d : deps) {
      if (std::find(seen.begin(), seen.end(), d.get()) != seen.end()) continue;
      DeviceEventPtr e = d->device_event();
      if (!e || !e->event) continue;
      seen.push_back(d.get());
      if (e->stream == stream) continue;  // same stream: already ordered
      h->stream_wait_event(stream, e->event);
    }
  }

  static void run(Task& t, ExecStats& stats, py::object& wrap) {
    // Device dependencies first, and without the GIL: the consumer stream is ordered after the
    // producers' events by the scheduler itself (no Python trampoline, no host blocking).
    const uint64_t stream = tl_stream();
    std::vector<const Ref*> dev_seen;
    if (stream) wait_device_inputs(t.deps, stream, dev_seen);
    pre_gil_spin();  // see ref.h: avoid sleeping on the GIL while the driver is mid-submit
    py::gil_scoped_acquire gil;
    py::object dep_err;
    py::tuple in_args = py::reinterpret_borrow<py::tuple>(t.args);
    py::tuple call_args(in_args.size());
    for (size_t i = 0; i < in_args.size() && !dep_err; ++i) {
      py::object item = in_args[i];
      if (py::isinstance<Ref>(item)) {
        RefPtr r = item.cast<RefPtr>();
        py::object e = r->exception();
        if (!e.is_none()) {
          dep_err = e;
          break;
        }
        call_args[i] = r->result();
      } else {
        call_args[i] = item;
      }
    }
    py::dict call_kwargs;
    if (!dep_err && t.kwargs && !t.kwargs.is_none()) {
      for (auto kv : py::reinterpret_borrow<py::dict>(t.kwargs)) {
        if (py::isinstance<Ref>(kv.second)) {
          RefPtr r = kv.second.cast<RefPtr>();
          py::object e = r->exception();
          if (!e.is_none()) {
            dep_err = e;
            break;
          }
          call_kwargs[kv.first] = r->result();
        } else {
          call_kwargs[kv.first] = kv.second;
        }
      }
    }
    // Outcome is computed first and PUBLISHED LAST: the waiter (usually the driver spinning
    // in fed.get) wakes the moment a Ref completes and immediately wants the GIL, so every
    // microsecond this thread keeps holding it after publishing is a chance for the waiter to
    // fall asleep on the GIL instead of running.
    std::vector<py::object> values;
    py::object error;
    DeviceEventPtr produced;
    if (dep_err) {
      // An upstream task failed: the failure flows along the edge (Ray semantics).
      error = dep_err;
    } else {
      // inputs that were still raw bytes above got their event from the decoder just now
      if (stream) wait_device_inputs(t.deps, stream, dev_seen);
      try {
        py::object res = t.fn(*call_args, **call_kwargs);
        if (stream) {
          // one event after the body orders every output of this task on the device
          if (const DeviceHooks* h = device_hooks()) {
            uint64_t ev = EventPool::get().acquire(h);
            if (ev && h->event_record(ev, stream) == 0) {
              produced = std::make_shared<DeviceEvent>();
              produced->event = ev;
              produced->stream = stream;
              produced->pooled = true;
            } else if (ev) {
              EventPool::get().release(ev);
            }
          }
        }
        size_t n = t.outs.size();
        if (n == 1) {
          values.push_back(std::move(res));
        } else {
          py::list parts(res);
          if (parts.size() != n) {
            std::string msg = "task declared num_returns=" + std::to_string(n) + " but returned " +
                              std::to_string(parts.size()) + " values";
            PyErr_SetString(PyExc_ValueError, msg.c_str());
            throw py::error_already_set();
          }
          for (size_t i = 0; i < n; ++i) values.push_back(parts[i]);
        }
      } catch (py::error_already_set& e) {
        values.clear();
        try {
          e.restore();
          PyObject *ty, *val, *tb;
          PyErr_Fetch(&ty, &val, &tb);
          PyErr_NormalizeException(&ty, &val, &tb);
          if (tb && val) PyException_SetTraceback(val, tb);
          py::object v = py::reinterpret_steal<py::object>(val);
          Py_XDECREF(ty);
          Py_XDECREF(tb);
          error = wrap ? wrap(v) : v;
        } catch (py::error_already_set& e2) {
          error = py::reinterpret_borrow<py::object>(e2.value());
        }
      } catch (std::exception& e) {
        values.clear();
        error = py::reinterpret_steal<py::object>(PyObject_CallFunction(PyExc_RuntimeError, "s", e.what()));
      }
    }
    // release Python payloads while we still hold the GIL — before anyone is woken
    call_args = py::tuple();
    call_kwargs = py::dict();
    in_args = py::tuple();
    t.fn = py::object();
    t.args = py::object();
    t.kwargs = py::object();
    t.deps.clear();
    if (error) {
      stats.failed.fetch_add(1, std::memory_order_relaxed);
      for (auto& o : t.outs) o->set_error(error);
    } else {
      stats.completed.fetch_add(1, std::memory_order_relaxed);
      for (size_t i = 0; i < t.outs.size(); ++i) {
        if (produced) t.outs[i]->set_device_event(produced);
        t.outs[i]->set_value(std::move(values[i]));
      }
    }
  }

  std::mutex mu_;
  std::vector<LanePtr> lanes_;
  std::shared_ptr<ExecStats> stats_;
  py::object wrap_error_;
};

}  // namespace fedcore
