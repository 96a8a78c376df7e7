// Device ordering for the scheduler ("dependency resolution in C++ over CUDA
// streams").  The reference gets DAG ordering from Ray's ObjectRef resolution on the host
// (fed/api.py:413-416, fed/proxy/barriers.py:478-483).  Here an edge may also carry a *device*
// ordering constraint: a `Ref` produced by a task that ran on a lane with a CUDA stream owns a
// CUDA event recorded after the task body; a consumer lane makes its own stream wait for that
// event (cudaStreamWaitEvent) — issued by the executor itself, before it takes the GIL — so the
// host never blocks for a device->device dependency and local tasks overlap in-flight pushes.
//
// `_fedcore` stays free of any CUDA link dependency (it must import on CPU-only machines):
// the five calls it needs come through this table, which `_fedcuda` publishes as a PyCapsule
// (`_fedcuda.device_hooks()`) and `_fedcore.set_device_hooks()` installs.  A recording fake
// (`_fedcore.install_fake_device_hooks()`) lets the ordering logic be tested without a GPU.
#pragma once
#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace fedcore {

struct DeviceHooks {
  // every function returns 0 on success, a CUDA error code otherwise
  int (*event_create)(uint64_t* ev);  // cudaEventDisableTiming
  int (*event_destroy)(uint64_t ev);
  int (*event_record)(uint64_t ev, uint64_t stream);
  int (*stream_wait_event)(uint64_t stream, uint64_t ev);
  int (*event_synchronize)(uint64_t ev);
  int (*event_query)(uint64_t ev);  // 0 = complete, 1 = not yet
};

inline std::atomic<const DeviceHooks*> g_device_hooks{nullptr};
inline const DeviceHooks* device_hooks() { return g_device_hooks.load(std::memory_order_acquire); }

// Free list of CUDA events: create/destroy costs microseconds and takes a driver lock.
class EventPool {
 public:
  static EventPool& get() {
    static EventPool* p = new EventPool();  // leaked on purpose: events outlive static destruction order
    return *p;
  }
  uint64_t acquire(const DeviceHooks* h) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!free_.empty()) {
        uint64_t e = free_.back();
        free_.pop_back();
        return e;
      }
    }
    uint64_t e = 0;
    if (h->event_create(&e) != 0) return 0;
    created_.fetch_add(1, std::memory_order_relaxed);
    return e;
  }
  void release(uint64_t e) {
    if (!e) return;
    std::lock_guard<std::mutex> lk(mu_);
    free_.push_back(e);
  }
  // Called when the hooks change (tests swap the fake in and out): events of the old
  // provider must not be handed to the new one.
  void drop_all(const DeviceHooks* h) {
    std::vector<uint64_t> v;
    {
      std::lock_guard<std::mutex> lk(mu_);
      v.swap(free_);
    }
    if (h)
      for (uint64_t e : v) h->event_destroy(e);
  }
  uint64_t created() const { return created_.load(); }

 private:
  std::mutex mu_;
  std::vector<uint64_t> free_;
  std::atomic<uint64_t> created_{0};
};

// ---- recording fake (CPU tests) ----------------------------------------------------------
struct FakeDeviceLog {
  std::mutex mu;
  std::vector<std::string> lines;
  std::atomic<uint64_t> next_event{1000};
  void add(std::string s) {
    std::lock_guard<std::mutex> lk(mu);
    lines.push_back(std::move(s));
  }
};
inline FakeDeviceLog& fake_device_log() {
  static FakeDeviceLog* l = new FakeDeviceLog();
  return *l;
}
inline const DeviceHooks* fake_device_hooks() {
  static DeviceHooks h = {
      [](uint64_t* ev) -> int {
        *ev = fake_device_log().next_event.fetch_add(1);
        return 0;
      },
      [](uint64_t) -> int { return 0; },
      [](uint64_t ev, uint64_t stream) -> int {
        fake_device_log().add("record e" + std::to_string(ev) + " s" + std::to_string(stream));
        return 0;
      },
      [](uint64_t stream, uint64_t ev) -> int {
        fake_device_log().add("wait s" + std::to_string(stream) + " e" + std::to_string(ev));
        return 0;
      },
      [](uint64_t ev) -> int {
        fake_device_log().add("sync e" + std::to_string(ev));
        return 0;
      },
      [](uint64_t) -> int { return 0; },
  };
  return &h;
}

}  // namespace fedcore
