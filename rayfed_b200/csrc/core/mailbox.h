// fedcore::Mailbox — the receive-side rendezvous table.
//
// Behavioural contract of the reference receiver (fed/proxy/grpc/grpc_proxy.py
// :272-293 `get_data` vs :309-342 `SendData`): whichever of {message arrival,
// consumer request} happens first creates the entry keyed
// (upstream_seq_id, downstream_seq_id); the other side finds it; the entry is
// popped once both have happened.  The reference does this with two dicts of
// asyncio.Events under a threading.Lock inside an actor process; here it is a
// hash map of Ref promises completed directly by the transport thread.
#pragma once
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "ref.h"

namespace fedcore {

struct MailboxStats {
  std::atomic<uint64_t> delivered{0};
  std::atomic<uint64_t> claimed{0};
  std::atomic<uint64_t> bytes{0};
};

class Mailbox {
 public:
  static std::string make_key(const std::string& up, const std::string& down) {
    std::string k;
    k.reserve(up.size() + down.size() + 1);
    k.append(up).push_back('\x1f');
    k.append(down);
    return k;
  }

  // Consumer side: the Ref resolves when the message lands (maybe already has).
  RefPtr claim(const std::string& up, const std::string& down) {
    std::string key = make_key(up, down);
    std::lock_guard<std::mutex> lk(mu_);
    stats_.claimed.fetch_add(1, std::memory_order_relaxed);
    auto it = slots_.find(key);
    if (it != slots_.end()) {
      Slot s = it->second;
      if (s.filled) {
        slots_.erase(it);  // both halves met: pop
        return s.ref;
      }
      // A second claim on a pending key shares the same future.
      return s.ref;
    }
    Slot s;
    s.ref = std::make_shared<Ref>();
    s.claimed = true;
    slots_.emplace(std::move(key), s);
    return s.ref;
  }

  // Several claims that share the downstream id (the arguments of one task): one lock, one call.
  std::vector<RefPtr> claim_many(const std::vector<std::string>& ups, const std::string& down) {
    std::vector<RefPtr> out;
    out.reserve(ups.size());
    std::lock_guard<std::mutex> lk(mu_);
    for (const auto& up : ups) {
      stats_.claimed.fetch_add(1, std::memory_order_relaxed);
      std::string key = make_key(up, down);
      auto it = slots_.find(key);
      if (it != slots_.end()) {
        RefPtr r = it->second.ref;
        if (it->second.filled) slots_.erase(it);
        out.push_back(std::move(r));
        continue;
      }
      Slot s;
      s.ref = std::make_shared<Ref>();
      s.claimed = true;
      out.push_back(s.ref);
      slots_.emplace(std::move(key), std::move(s));
    }
    return out;
  }

  // Transport side (no GIL): bytes arrive.
  void deliver_raw(const std::string& up, const std::string& down, std::string bytes, bool is_error) {
    RefPtr ref = take_for_delivery(up, down);
    stats_.bytes.fetch_add(bytes.size(), std::memory_order_relaxed);
    ref->set_raw(std::move(bytes), is_error);
  }

  // In-process loopback / device transports hand over a ready Python object (GIL held).
  void deliver_obj(const std::string& up, const std::string& down, py::object obj, bool is_error) {
    RefPtr ref = take_for_delivery(up, down);
    if (is_error)
      ref->set_error(std::move(obj));
    else
      ref->set_value(std::move(obj));
  }

  size_t pending() {
    std::lock_guard<std::mutex> lk(mu_);
    return slots_.size();
  }
  // Diagnostics: keys a consumer is waiting for vs. messages nobody asked for.  A growing
  // second list while the first one stalls is the signature of controller divergence
  // (the parties' scripts issued different numbers of fed calls).
  void snapshot(std::vector<std::pair<std::string, std::string>>* waiting,
                std::vector<std::pair<std::string, std::string>>* unclaimed, size_t limit) {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : slots_) {
      auto sep = kv.first.find('\x1f');
      std::pair<std::string, std::string> k{kv.first.substr(0, sep),
                                            sep == std::string::npos ? "" : kv.first.substr(sep + 1)};
      if (kv.second.claimed && !kv.second.filled) {
        if (waiting->size() < limit) waiting->push_back(std::move(k));
      } else if (kv.second.filled && !kv.second.claimed) {
        if (unclaimed->size() < limit) unclaimed->push_back(std::move(k));
      }
    }
  }
  void clear() {
    std::unordered_map<std::string, Slot> dead;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dead.swap(slots_);
    }
  }
  uint64_t delivered() const { return stats_.delivered.load(); }
  uint64_t claimed() const { return stats_.claimed.load(); }
  uint64_t bytes() const { return stats_.bytes.load(); }

 private:
  struct Slot {
    RefPtr ref;
    bool filled = false;
    bool claimed = false;
  };

  RefPtr take_for_delivery(const std::string& up, const std::string& down) {
    std::string key = make_key(up, down);
    std::lock_guard<std::mutex> lk(mu_);
    stats_.delivered.fetch_add(1, std::memory_order_relaxed);
    auto it = slots_.find(key);
    if (it != slots_.end() && it->second.claimed && !it->second.filled) {
      RefPtr r = it->second.ref;
      slots_.erase(it);  // consumer already holds the future: pop
      return r;
    }
    // Nobody asked yet (or a never-claimed duplicate such as the init ping,
    // which the reference also leaves in its table): park a filled slot.
    Slot s;
    s.ref = std::make_shared<Ref>();
    s.filled = true;
    if (it != slots_.end())
      it->second = s;
    else
      slots_.emplace(std::move(key), s);
    return s.ref;
  }

  std::mutex mu_;
  std::unordered_map<std::string, Slot> slots_;
  MailboxStats stats_;
};

}  // namespace fedcore
