// fedcore::Conn — the byte channel under one transport connection.
//
// Every connection is a stream socket (TCP loopback / Unix): it provides connect-retry, the
// connection life cycle and peer-death detection (EOF).  When both ends run on one machine and
// `RAYFED_B200_SHM=1`, the DATA moves through a pair of single-producer/single-consumer byte
// rings in a POSIX shared-memory segment instead (this is the "HostRing" of SURVEY §7.1): a
// frame costs two memcpys and no system call, and a spinning reader sees it without a kernel
// wake-up.  A consumer that goes to sleep announces it in the ring header and blocks on that word
// (cross-process futex); the producer that sees the announcement wakes it.  The socket then only
// carries (a) the one-frame handshake that hands the segment name to the server and (b) EOF when
// a side goes away.  The creator unlinks the segment name as
// soon as the peer has mapped it, so nothing is left in /dev/shm even after a crash.
#pragma once
#include <fcntl.h>
#include <linux/futex.h>
#include <poll.h>
#include <sys/syscall.h>
#include <sys/mman.h>
#include <time.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include "ref.h"

namespace fedcore {

static constexpr uint32_t kShmMagic = 0x52424653;  // "SFBR"

struct ShmRing {  // in shared memory
  alignas(64) std::atomic<uint64_t> head;             // bytes produced
  alignas(64) std::atomic<uint64_t> tail;             // bytes consumed
  alignas(64) std::atomic<uint32_t> consumer_asleep;  // consumer is (about to be) blocked on this word (futex)
};

struct ShmHeader {
  uint32_t magic;
  uint32_t capacity;  // bytes per ring, power of two
  ShmRing c2s;        // client -> server
  ShmRing s2c;        // server -> client
};
static constexpr size_t kShmHeaderBytes = 4096;
static_assert(sizeof(ShmHeader) <= kShmHeaderBytes, "header must fit its page");

// Read per connection attempt (not cached): cheap, and lets one process hold both kinds.
// Default ON for same-host peers (the sender only offers it to loopback / unix addresses, and
// either side may refuse): a tiny fed message costs a ring memcpy + a futex wake instead of a
// TCP loopback round trip, which is what lets the N-1 sends of a fed.get broadcast and the N-1
// receives of a fan-in overlap instead of queueing behind syscalls.  RAYFED_B200_SHM=0 disables.
inline bool shm_enabled() {
  const char* e = std::getenv("RAYFED_B200_SHM");
  return !(e && e[0] == '0');
}
inline uint32_t shm_ring_bytes() {
  const char* e = std::getenv("RAYFED_B200_SHM_RING_KB");
  uint32_t kb = e ? (uint32_t)std::atoi(e) : 1024;
  uint32_t b = 64 * 1024;
  while (b < kb * 1024u && b < (1u << 30)) b <<= 1;
  return b;
}

inline bool fd_write_all(int fd, const void* buf, size_t n) {
  const char* p = (const char*)buf;
  while (n > 0) {
    ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += w;
    n -= (size_t)w;
  }
  return true;
}
inline bool fd_read_all(int fd, void* buf, size_t n) {
  char* p = (char*)buf;
  while (n > 0) {
    ssize_t r = ::recv(fd, p, n, 0);
    if (r == 0) return false;
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += r;
    n -= (size_t)r;
  }
  return true;
}

// Cross-process futex on a word in the shared segment (no FUTEX_PRIVATE_FLAG): the doorbell of a
// ring.  A wake is ~1-2 us; the one-byte TCP/Unix send it replaces cost 7-8 us on the B200 host,
// paid by the PRODUCER on the critical path of every hop (and 7 times in a row by a fed.get
// broadcast at 8 parties).
inline int futex_wait_shared(std::atomic<uint32_t>* addr, uint32_t expected, int timeout_ms) {
  timespec ts{timeout_ms / 1000, (long)(timeout_ms % 1000) * 1000000L};
  return (int)::syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), FUTEX_WAIT, expected,
                        timeout_ms >= 0 ? &ts : nullptr, nullptr, 0);
}
inline void futex_wake_shared(std::atomic<uint32_t>* addr) {
  ::syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), FUTEX_WAKE, 1, nullptr, nullptr, 0);
}

// Absolute-deadline wait with a wake mask: several consumers (one per receiving party) sleep on
// ONE word of the sender's doorbell page, each with its own bit; the sender wakes any subset of
// them with a single system call (a fed.get broadcast to N-1 parties: 1 wake instead of N-1).
inline int futex_wait_bitset_shared(std::atomic<uint32_t>* addr, uint32_t expected, int timeout_ms, uint32_t mask) {
  timespec ts{};
  ::clock_gettime(CLOCK_MONOTONIC, &ts);  // FUTEX_WAIT_BITSET takes an ABSOLUTE CLOCK_MONOTONIC deadline
  ts.tv_sec += timeout_ms / 1000;
  ts.tv_nsec += (long)(timeout_ms % 1000) * 1000000L;
  if (ts.tv_nsec >= 1000000000L) {
    ts.tv_sec += 1;
    ts.tv_nsec -= 1000000000L;
  }
  return (int)::syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), FUTEX_WAIT_BITSET, expected,
                        timeout_ms >= 0 ? &ts : nullptr, nullptr, mask);
}
inline void futex_wake_bitset_shared(std::atomic<uint32_t>* addr, uint32_t mask) {
  ::syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), FUTEX_WAKE_BITSET, INT32_MAX, nullptr, nullptr, mask);
}

// One page per sending transport, mapped by every same-host receiver of that sender.
static constexpr uint32_t kPageMagic = 0x50424653;  // "SFBP"
struct DoorbellPage {
  uint32_t magic;
  alignas(64) std::atomic<uint32_t> seq;  // bumped before every wake: a consumer that read the old
                                          // value and then calls futex_wait returns at once
};
static constexpr size_t kDoorbellPageBytes = 4096;

class DoorbellPageOwner {
 public:
  DoorbellPageOwner() = default;
  ~DoorbellPageOwner() {
    if (map_) ::munmap(map_, kDoorbellPageBytes);
    if (fd_ >= 0) ::close(fd_);
  }
  DoorbellPageOwner(const DoorbellPageOwner&) = delete;
  DoorbellPageOwner& operator=(const DoorbellPageOwner&) = delete;
  // "<pid>:<fd>" of an anonymous memory file (memfd): a same-host peer of the same user maps it
  // through /proc/<pid>/fd/<fd> during the handshake.  Nothing is ever visible in /dev/shm, so
  // nothing can be left behind by a crash.  "" when unavailable: the rings use their own words.
  const std::string& ensure() {
    std::lock_guard<std::mutex> lk(mu_);
    if (map_ || failed_) return desc_;
    int fd = (int)::syscall(SYS_memfd_create, "rfb.doorbell", 1u /* MFD_CLOEXEC */);
    if (fd < 0 || ::ftruncate(fd, (off_t)kDoorbellPageBytes) != 0) {
      if (fd >= 0) ::close(fd);
      failed_ = true;
      return desc_;
    }
    void* m = ::mmap(nullptr, kDoorbellPageBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) {
      ::close(fd);
      failed_ = true;
      return desc_;
    }
    auto* pg = new (m) DoorbellPage();
    pg->seq.store(0);
    pg->magic = kPageMagic;
    map_ = m;
    fd_ = fd;
    desc_ = std::to_string((long)getpid()) + ":" + std::to_string(fd);
    return desc_;
  }
  std::atomic<uint32_t>* seq() { return map_ ? &reinterpret_cast<DoorbellPage*>(map_)->seq : nullptr; }

 private:
  std::mutex mu_;
  std::string desc_;
  void* map_ = nullptr;
  int fd_ = -1;
  bool failed_ = false;
};

// Thread-local collector: while one is active, doorbells that go through the given page are not
// rung one by one but OR-ed into a mask and rung together when the scope ends.
struct DoorbellBatch {
  std::atomic<uint32_t>* seq = nullptr;
  uint32_t mask = 0;
  static DoorbellBatch*& current() {
    static thread_local DoorbellBatch* cur = nullptr;
    return cur;
  }
  DoorbellBatch() { current() = this; }
  ~DoorbellBatch() {
    current() = nullptr;
    flush();
  }
  void flush() {
    if (seq && mask) {
      seq->fetch_add(1, std::memory_order_seq_cst);
      futex_wake_bitset_shared(seq, mask);
    }
    mask = 0;
  }
  DoorbellBatch(const DoorbellBatch&) = delete;
  DoorbellBatch& operator=(const DoorbellBatch&) = delete;
};

class Conn {
 public:
  explicit Conn(int fd) : fd_(fd) {}
  ~Conn() {
    if (map_) ::munmap(map_, map_len_);
    if (rx_page_map_) ::munmap(rx_page_map_, kDoorbellPageBytes);
  }
  Conn(const Conn&) = delete;
  Conn& operator=(const Conn&) = delete;

  int fd() const { return fd_; }
  bool uses_shm() const { return tx_ != nullptr; }
  void mark_dead() { dead_.store(true, std::memory_order_release); }
  bool dead() const { return dead_.load(std::memory_order_acquire); }

  // ---- client side: create a segment, returns its name ("" on failure) ------------------
  std::string create_segment() {
    static std::atomic<uint64_t> seq{0};
    std::string name = "/rfb." + std::to_string((long)getpid()) + "." + std::to_string(seq.fetch_add(1));
    int sfd = ::shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (sfd < 0) return "";
    const uint32_t cap = shm_ring_bytes();
    size_t len = kShmHeaderBytes + 2 * (size_t)cap;
    if (::ftruncate(sfd, (off_t)len) != 0) {
      ::close(sfd);
      ::shm_unlink(name.c_str());
      return "";
    }
    void* m = ::mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, sfd, 0);
    ::close(sfd);
    if (m == MAP_FAILED) {
      ::shm_unlink(name.c_str());
      return "";
    }
    auto* h = new (m) ShmHeader();
    h->capacity = cap;
    h->c2s.head.store(0);
    h->c2s.tail.store(0);
    h->c2s.consumer_asleep.store(0);
    h->s2c.head.store(0);
    h->s2c.tail.store(0);
    h->s2c.consumer_asleep.store(0);
    h->magic = kShmMagic;
    pending_map_ = m;
    pending_len_ = len;
    pending_name_ = name;
    return name;
  }
  // handshake answered: switch to the rings (client view) or drop the segment
  void finish_create(bool accepted) {
    if (!pending_map_) return;
    ::shm_unlink(pending_name_.c_str());  // both sides have it mapped (or nobody wants it)
    if (accepted) {
      adopt(pending_map_, pending_len_, /*is_client=*/true);
    } else {
      ::munmap(pending_map_, pending_len_);
    }
    pending_map_ = nullptr;
  }
  // ---- server side: map the client's segment ---------------------------------------------------
  bool attach_segment(const std::string& name) {
    if (name.size() < 2 || name[0] != '/' || name.find('/', 1) != std::string::npos) return false;
    int sfd = ::shm_open(name.c_str(), O_RDWR, 0600);
    if (sfd < 0) return false;
    struct stat st {};
    if (::fstat(sfd, &st) != 0 || (size_t)st.st_size < kShmHeaderBytes + 2 * 64 * 1024) {
      ::close(sfd);
      return false;
    }
    size_t len = (size_t)st.st_size;
    void* m = ::mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, sfd, 0);
    ::close(sfd);
    if (m == MAP_FAILED) return false;
    auto* h = reinterpret_cast<ShmHeader*>(m);
    uint32_t cap = h->capacity;
    if (h->magic != kShmMagic || cap < 64 * 1024 || (cap & (cap - 1)) != 0 ||
        kShmHeaderBytes + 2 * (size_t)cap != len) {
      ::munmap(m, len);
      return false;
    }
    adopt(m, len, /*is_client=*/false);
    return true;
  }

  // ---- doorbell page (optional; both sides fall back to the ring's own futex word) ---------------
  // server side: map the sender's page; this connection's reader then sleeps on its word with `bit`
  bool attach_rx_page(const std::string& desc, int slot) {
    if (slot < 0 || slot > 31) return false;
    // "<pid>:<fd>", digits only: the path below is built from nothing else
    size_t colon = desc.find(':');
    if (colon == std::string::npos || colon == 0 || colon + 1 >= desc.size() || desc.size() > 24) return false;
    if (desc.find_first_not_of("0123456789:") != std::string::npos || desc.find(':', colon + 1) != std::string::npos)
      return false;
    std::string path = "/proc/" + desc.substr(0, colon) + "/fd/" + desc.substr(colon + 1);
    // only an anonymous memory file created by DoorbellPageOwner is acceptable — never a path the
    // peer could have picked in the file system
    char target[64] = {0};
    ssize_t tl = ::readlink(path.c_str(), target, sizeof(target) - 1);
    if (tl <= 0 || std::strncmp(target, "/memfd:rfb.doorbell", 19) != 0) return false;
    int sfd = ::open(path.c_str(), O_RDWR | O_CLOEXEC);
    if (sfd < 0) return false;  // another pid namespace / another user: the ring's own word is used
    struct stat st {};
    if (::fstat(sfd, &st) != 0 || !S_ISREG(st.st_mode) || (size_t)st.st_size != kDoorbellPageBytes) {
      ::close(sfd);
      return false;
    }
    void* m = ::mmap(nullptr, kDoorbellPageBytes, PROT_READ | PROT_WRITE, MAP_SHARED, sfd, 0);
    ::close(sfd);
    if (m == MAP_FAILED) return false;
    if (reinterpret_cast<DoorbellPage*>(m)->magic != kPageMagic) {
      ::munmap(m, kDoorbellPageBytes);
      return false;
    }
    rx_page_map_ = m;
    rx_bit_ = 1u << slot;
    rx_page_seq_.store(&reinterpret_cast<DoorbellPage*>(m)->seq, std::memory_order_release);
    return true;
  }
  // client side: the server sleeps on our page; ring its doorbell there
  void set_tx_page(std::atomic<uint32_t>* seq, int slot) {
    tx_bit_ = 1u << slot;
    tx_page_seq_ = seq;
  }

  // ---- byte stream -------------------------------------------------------------------------------
  bool write_all(const void* buf, size_t n) {
    if (!tx_) return fd_write_all(fd_, buf, n);
    const char* p = (const char*)buf;
    const uint64_t cap = cap_;
    int idle = 0;
    while (n > 0) {
      uint64_t head = tx_->head.load(std::memory_order_relaxed);
      uint64_t tail = tx_->tail.load(std::memory_order_acquire);
      if (head - tail > cap) return false;  // the peer scribbled over its counter: never index past the ring
      uint64_t space = cap - (head - tail);
      if (space == 0) {
        if (dead()) return false;
        ring_doorbell(/*now=*/true);  // the consumer may be asleep with a full ring in front of it
        if (++idle < 2000) {
          cpu_relax();
        } else {
          if (peer_closed()) return false;
          std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
        continue;
      }
      idle = 0;
      size_t chunk = (size_t)std::min<uint64_t>(space, n);
      size_t off = (size_t)(head & (cap - 1));
      size_t first = std::min(chunk, (size_t)(cap - off));
      std::memcpy(tx_data_ + off, p, first);
      if (chunk > first) std::memcpy(tx_data_, p + first, chunk - first);
      // seq_cst, not just release: Dekker with the consumer's sleep announcement (ring_doorbell)
      tx_->head.store(head + chunk, std::memory_order_seq_cst);
      p += chunk;
      n -= chunk;
    }
    ring_doorbell();
    return true;
  }

  bool read_all(void* buf, size_t n) {
    if (!rx_) return fd_read_all(fd_, buf, n);
    char* p = (char*)buf;
    const uint64_t cap = cap_;
    while (n > 0) {
      uint64_t tail = rx_->tail.load(std::memory_order_relaxed);
      uint64_t head = rx_->head.load(std::memory_order_acquire);
      uint64_t avail = head - tail;
      if (avail > cap) return false;  // a corrupt producer counter must not turn into an out-of-bounds copy
      if (avail == 0) {
        int w = wait_readable(-1);
        if (w < 0) return false;
        continue;
      }
      size_t chunk = (size_t)std::min<uint64_t>(avail, n);
      size_t off = (size_t)(tail & (cap - 1));
      size_t first = std::min(chunk, (size_t)(cap - off));
      std::memcpy(p, rx_data_ + off, first);
      if (chunk > first) std::memcpy(p + first, rx_data_, chunk - first);
      rx_->tail.store(tail + chunk, std::memory_order_release);
      p += chunk;
      n -= chunk;
    }
    return true;
  }

  // 1: data available, 0: timeout, -1: connection closed / error.  timeout_ms < 0: no limit.
  int wait_readable(int timeout_ms) {
    if (!rx_) return wait_readable_impl(timeout_ms);
    auto t0 = std::chrono::steady_clock::now();
    int r = wait_readable_impl(timeout_ms);
    if (r == 1) lazy_polls_ = 0;
    if (r == 1)  // how long this side sat idle before the frame came: drives the adaptive spin
      idle_.observe(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    return r;
  }

  // Receivers nobody waits for (acknowledgement readers) never spin longer than the base budget.
  void set_latency_critical(bool v) { latency_critical_ = v; }

 private:
  int wait_readable_impl(int timeout_ms) {
    if (!rx_) {
      for (;;) {
        pollfd p{fd_, POLLIN, 0};
        int pr = ::poll(&p, 1, timeout_ms < 0 ? 100 : timeout_ms);
        if (pr > 0) return 1;  // readable or hung up: the following read tells which
        if (pr < 0 && errno != EINTR) return -1;
        if (timeout_ms >= 0) return 0;
        if (dead()) return -1;
      }
    }
    auto ready = [&] { return rx_->head.load(std::memory_order_acquire) != rx_->tail.load(std::memory_order_relaxed); };
    auto start = std::chrono::steady_clock::now();
    for (;;) {
      if (ready()) return 1;
      if (dead()) return -1;
      // bounded user-space spin (same budget as every other wait of the runtime); a receive
      // side whose frames arrive every few hundred us keeps spinning between them (adaptive):
      // the producer then never pays the doorbell system call and the frame is seen at once
      const int base = latency_critical_ ? spin_budget_us() : 0;  // acks: nobody waits, do not spin
      int budget = latency_critical_ ? idle_.budget_us() : 0;
      if (budget > base) {
        LongSpinSlot slot;
        if (slot.ok ? spin_long(ready, budget) : spin_for(ready, base)) return 1;
      } else if (spin_for(ready, budget)) {
        return 1;
      }
      if (!latency_critical_ && lazy_polls_ < kLazyPolls) {
        // Nobody waits for what this side receives (acknowledgements): do not announce a sleep —
        // the producer then never pays a wake-up — and look again in a moment.  After a quiet
        // while fall through to the announced sleep so that an idle connection costs nothing.
        ++lazy_polls_;
        std::this_thread::sleep_for(std::chrono::microseconds(kLazyPollUs));
        if (ready()) {
          lazy_polls_ = 0;
          return 1;
        }
        if (timeout_ms >= 0) {
          auto spent = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - start).count();
          if (spent >= timeout_ms) return 0;
        }
        continue;
      }
      // announce the sleep, re-check (Dekker), then block on the futex word: the producer clears
      // it and wakes us when it sees the announcement.  Peer death is noticed on the socket (EOF)
      // between slices; our own side calls wake_reader() when it closes the connection.
      int slice = 100;
      if (timeout_ms >= 0) {
        auto spent = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - start).count();
        if (spent >= timeout_ms) return ready() ? 1 : 0;
        slice = (int)std::min<long long>(100, timeout_ms - spent);
      }
      std::atomic<uint32_t>* page = rx_page_seq_.load(std::memory_order_acquire);
      rx_->consumer_asleep.store(1, std::memory_order_seq_cst);
      // read the page word BEFORE the last look at the ring: a producer that bumps it after this
      // point makes the wait below return at once; one that bumped it before has published its
      // frame before that, so the look sees the frame
      const uint32_t seen = page ? page->load(std::memory_order_seq_cst) : 0;
      if (ready()) {
        rx_->consumer_asleep.store(0, std::memory_order_seq_cst);
        lazy_polls_ = 0;
        return 1;
      }
      if (page)
        futex_wait_bitset_shared(page, seen, slice, rx_bit_);
      else
        futex_wait_shared(&rx_->consumer_asleep, 1, slice);
      rx_->consumer_asleep.store(0, std::memory_order_seq_cst);
      if (ready()) {
        lazy_polls_ = 0;
        return 1;
      }
      if (dead()) return -1;
      if (peer_closed()) return ready() ? 1 : -1;  // EOF: drain what is already in the ring first
    }
  }

 public:
  // Our side is closing the connection: get a reader that sleeps on the ring out of its wait.
  void wake_reader() {
    mark_dead();
    if (rx_) {
      rx_->consumer_asleep.store(0, std::memory_order_seq_cst);
      if (std::atomic<uint32_t>* page = rx_page_seq_.load(std::memory_order_acquire)) {
        page->fetch_add(1, std::memory_order_seq_cst);
        futex_wake_bitset_shared(page, rx_bit_);
      } else {
        futex_wake_shared(&rx_->consumer_asleep);
      }
    }
  }

 private:
  void adopt(void* m, size_t len, bool is_client) {
    map_ = m;
    map_len_ = len;
    auto* h = reinterpret_cast<ShmHeader*>(m);
    cap_ = h->capacity;
    char* base = reinterpret_cast<char*>(m) + kShmHeaderBytes;
    if (is_client) {
      tx_ = &h->c2s;
      tx_data_ = base;
      rx_ = &h->s2c;
      rx_data_ = base + cap_;
    } else {
      tx_ = &h->s2c;
      tx_data_ = base + cap_;
      rx_ = &h->c2s;
      rx_data_ = base;
    }
  }
  // `now`: never defer into a DoorbellBatch (the caller is about to wait for this consumer).
  void ring_doorbell(bool now = false) {
    // Dekker: our `head` store must be globally visible before we read the consumer's announcement,
    // or both sides can miss each other — the consumer sleeps on a frame that is already there and
    // nobody wakes it until its 100 ms slice ends.  Both are seq_cst operations (a release store is
    // a plain store on x86 and may be reordered after the following load).
    // exchange: one wake per sleep, however many writes follow before the consumer is back
    if (!tx_->consumer_asleep.load(std::memory_order_seq_cst) ||
        tx_->consumer_asleep.exchange(0, std::memory_order_seq_cst) != 1)
      return;
    if (tx_page_seq_) {
      DoorbellBatch* batch = now ? nullptr : DoorbellBatch::current();
      if (batch && (batch->seq == nullptr || batch->seq == tx_page_seq_)) {
        batch->seq = tx_page_seq_;  // rung with the others when the batch ends
        batch->mask |= tx_bit_;
        return;
      }
      tx_page_seq_->fetch_add(1, std::memory_order_seq_cst);
      futex_wake_bitset_shared(tx_page_seq_, tx_bit_);
    } else {
      futex_wake_shared(&tx_->consumer_asleep);
    }
  }
  bool peer_closed() {
    // POLLRDHUP: the peer shut its side down (FIN) — a server that refused a frame and left, a
    // process that died — while our descriptor is still open; POLLHUP alone needs both directions.
    pollfd p{fd_, (short)(POLLIN | POLLRDHUP), 0};
    if (::poll(&p, 1, 0) > 0 && (p.revents & (POLLHUP | POLLRDHUP | POLLERR | POLLNVAL))) return true;
    return false;
  }

  int fd_;
  AdaptiveSpin idle_;              // reader thread only
  bool latency_critical_ = true;  // reader thread only
  int lazy_polls_ = 0;            // reader thread only
  static constexpr int kLazyPolls = 50;     // x kLazyPollUs of silence before the announced sleep
  static constexpr int kLazyPollUs = 200;
  std::atomic<bool> dead_{false};
  void* map_ = nullptr;
  size_t map_len_ = 0;
  uint64_t cap_ = 0;
  ShmRing* tx_ = nullptr;
  char* tx_data_ = nullptr;
  ShmRing* rx_ = nullptr;
  char* rx_data_ = nullptr;
  std::atomic<std::atomic<uint32_t>*> rx_page_seq_{nullptr};  // set once by the serve thread, read by wake_reader()
  void* rx_page_map_ = nullptr;
  uint32_t rx_bit_ = 0;
  std::atomic<uint32_t>* tx_page_seq_ = nullptr;  // set before the connection is published
  uint32_t tx_bit_ = 0;
  void* pending_map_ = nullptr;
  size_t pending_len_ = 0;
  std::string pending_name_;
};

}  // namespace fedcore
