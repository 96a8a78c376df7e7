// fedcore::SocketTransport — native host-side message transport for one party.
//
// Plays the role of the reference's gRPC sender/receiver pair
// (fed/proxy/grpc/grpc_proxy.py:102-220 sender, :223-381 receiver) for parties
// that share a box: one unary "SendData"-shaped exchange per object
// {data, upstream_seq_id, downstream_seq_id, job_name} -> {code, result}
// (schema: fed/grpc/fed.proto:5-19), but as length-prefixed frames over a
// TCP/Unix stream socket with pipelined acks, delivered by a C++ thread straight
// into the Mailbox (no actor process, no asyncio loop, no protobuf copy).
// Semantics kept: job-name check -> 417 (grpc_proxy.py:310-320), 4xx => send
// failure (:179-190), per-message timeout, connect retry with exponential
// backoff (grpc_options.py:19-25) so asynchronous party start-up works.
#pragma once
#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mailbox.h"
#include "ref.h"
#include "shm_conn.h"

namespace fedcore {

static constexpr uint32_t kMagic = 0x42444546;  // "FEDB"
enum : uint8_t { kFrameData = 1, kFrameAck = 2, kFrameHello = 3 };  // Hello: shm upgrade offer (shm_conn.h)

#pragma pack(push, 1)
struct FrameHeader {
  uint32_t magic;
  uint8_t type;
  uint8_t flags;  // bit0: payload is an error object
  uint16_t reserved;
  uint64_t msg_id;
  uint32_t job_len;
  uint32_t up_len;
  uint32_t down_len;
  int32_t code;  // acks only
  uint64_t data_len;
};
#pragma pack(pop)

// Opt-in on BOTH sides (RAYFED_B200_DOORBELL_PAGE=1).  One FUTEX_WAKE_BITSET for N sleepers saves
// N-1 system calls, but the kernel's work per woken task stays: on the development VM a 7-receiver
// broadcast costs 125 us either way, and the B200 host was not measured with it.
inline bool doorbell_page_enabled() {
  static const bool v = [] {
    const char* e = std::getenv("RAYFED_B200_DOORBELL_PAGE");
    return e && e[0] == '1';
  }();
  return v;
}

struct RetryPolicy {
  // `proxy_max_restarts` of the reference (fed/proxy/barriers.py:301-307: the sender proxy actor is
  // restarted once and its in-flight sends are retried N times): a message that was written but
  // not yet acknowledged when the connection died is re-queued up to this many times on the
  // re-established connection instead of failing its send.  At-least-once, like Ray's task retry.
  int max_resends = 0;
  int max_attempts = 5;
  double initial_backoff_s = 5.0;
  double max_backoff_s = 30.0;
  double multiplier = 2.0;
};

struct ParsedAddr {
  bool is_unix = false;
  std::string host;
  std::string port;
  std::string path;
};

inline ParsedAddr parse_addr(std::string a) {
  ParsedAddr out;
  if (a.rfind("unix:", 0) == 0) {
    out.is_unix = true;
    out.path = a.substr(5);
    return out;
  }
  auto strip = [&](const char* pre) {
    size_t n = strlen(pre);
    if (a.rfind(pre, 0) == 0) a = a.substr(n);
  };
  strip("http://");
  strip("https://");
  size_t slash = a.find('/');
  if (slash != std::string::npos) a = a.substr(0, slash);
  size_t colon = a.rfind(':');
  if (colon == std::string::npos) {
    out.host = a;
    out.port = "80";
  } else {
    out.host = a.substr(0, colon);
    out.port = a.substr(colon + 1);
  }
  if (out.host == "local" || out.host == "localhost" || out.host.empty()) out.host = "127.0.0.1";
  return out;
}

struct TransportStats {
  std::atomic<uint64_t> send_op_count{0};
  std::atomic<uint64_t> receive_op_count{0};
  std::atomic<uint64_t> bytes_sent{0};
  std::atomic<uint64_t> bytes_received{0};
  std::atomic<uint64_t> rejected{0};
  std::atomic<uint64_t> shm_conns{0};  // connections whose data path is the shared-memory rings
  std::atomic<uint64_t> paged_conns{0};  // ... whose reader sleeps on the sender's doorbell page
};

// ---------------------------------------------------------------------------
class Peer {
 public:
  struct Outgoing {
    uint64_t id;
    std::string up, down, data;
    bool is_error;
    RefPtr ack;
    std::chrono::steady_clock::time_point deadline;
    int resends = 0;  // times this message was re-queued after a lost connection
  };

  Peer(std::string name, std::string addr, std::string job, RetryPolicy retry, double timeout_s,
       TransportStats* stats, DoorbellPageOwner* page = nullptr, int slot = -1)
      : name_(std::move(name)), addr_(std::move(addr)), job_(std::move(job)), retry_(retry),
        timeout_s_(timeout_s), stats_(stats), page_(page), slot_(slot) {
    writer_ = std::thread([this] {
      try {
        writer_loop();
      } catch (...) {  // nothing may escape a std::thread
      }
    });
  }
  ~Peer() { stop(); }

  void push(std::string up, std::string down, std::string data, bool is_error, RefPtr ack) {
    auto m = std::make_shared<Outgoing>();
    m->up = std::move(up);
    m->down = std::move(down);
    m->data = std::move(data);
    m->is_error = is_error;
    m->ack = std::move(ack);
    m->deadline = std::chrono::steady_clock::now() + std::chrono::duration_cast<std::chrono::steady_clock::duration>(
                                                         std::chrono::duration<double>(timeout_s_));
    bool inline_send = false;
    {
      std::lock_guard<std::mutex> lk(mu_);
      m->id = next_id_++;
      // Fast path: connection is up and nothing is queued -> write from the
      // caller's thread and save a thread wake-up per message.
      if (!stopping_ && fd_.load() >= 0 && q_.empty() && !writing_) {
        writing_ = true;
        inline_send = true;
      } else {
        q_.push_back(std::move(m));
      }
    }
    if (inline_send) {
      transmit(std::move(m));
      {
        std::lock_guard<std::mutex> lk(mu_);
        writing_ = false;
      }
      cv_.notify_all();
    } else {
      cv_.notify_one();
    }
  }

  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (stopping_) return;
      stopping_ = true;
    }
    cv_.notify_all();
    close_conn();
    if (writer_.joinable()) writer_.join();
    {
      // an inline sender (push() on a caller thread) may still be inside write(); the
      // shutdown above makes it fail fast — let it leave before the descriptor goes away
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait_for(lk, std::chrono::seconds(1), [&] { return !writing_; });
    }
    reap_reader();
    fail_all("transport stopped");
  }

  // Wait until the queue and all in-flight acks drain (GIL released by caller).
  bool drain(double timeout_s) {
    auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s);
    for (;;) {
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (q_.empty() && inflight_.empty() && !writing_) return true;
      }
      if (std::chrono::steady_clock::now() > deadline) return false;
      std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
  }

 private:
  int connect_once() {
    ParsedAddr pa = parse_addr(addr_);
    int fd = -1;
    if (pa.is_unix) {
      fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
      if (fd < 0) return -1;
      sockaddr_un sa{};
      sa.sun_family = AF_UNIX;
      strncpy(sa.sun_path, pa.path.c_str(), sizeof(sa.sun_path) - 1);
      if (::connect(fd, (sockaddr*)&sa, sizeof(sa)) != 0) {
        ::close(fd);
        return -1;
      }
      return fd;
    }
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(pa.host.c_str(), pa.port.c_str(), &hints, &res) != 0) return -1;
    for (addrinfo* ai = res; ai; ai = ai->ai_next) {
      fd = ::socket(ai->ai_family, ai->ai_socktype, ai->ai_protocol);
      if (fd < 0) continue;
      if (::connect(fd, ai->ai_addr, ai->ai_addrlen) == 0) {
        int one = 1;
        setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
        break;
      }
      ::close(fd);
      fd = -1;
    }
    freeaddrinfo(res);
    return fd;
  }

  // Retry with exponential backoff; gives up after max_attempts.
  bool ensure_connected() {
    if (fd_.load() >= 0) return true;
    reap_reader();
    double backoff = retry_.initial_backoff_s;
    for (int attempt = 1; attempt <= retry_.max_attempts; ++attempt) {
      int fd = connect_once();
      if (fd >= 0) {
        auto c = std::make_shared<Conn>(fd);
        if (shm_enabled() && shm_ok_ && is_local_addr() && !offer_shm(*c)) {
          // the handshake broke the stream: start over on a fresh socket, without the offer
          shm_ok_ = false;
          ::close(fd);
          --attempt;
          continue;
        }
        std::atomic_store(&conn_, c);
        reader_fd_ = fd;
        fd_.store(fd);
        reader_ = std::thread([this, c] {
          try {
            reader_loop(c);
          } catch (...) {
          }
        });
        return true;
      }
      if (attempt == retry_.max_attempts) break;
      std::unique_lock<std::mutex> lk(mu_);
      if (cv_.wait_for(lk, std::chrono::duration<double>(backoff), [&] { return stopping_; })) return false;
      backoff = std::min(backoff * retry_.multiplier, retry_.max_backoff_s);
    }
    return false;
  }

  // Descriptor lifecycle: any thread may *retire* the connection (fd_ -> -1 + shutdown,
  // which wakes the reader and fails writers); the descriptor itself is closed only by
  // whoever joins the reader thread (reap_reader), so no thread ever sits in a syscall
  // on a number that was closed — and possibly reused — under it.
  void close_conn() {
    int fd = fd_.exchange(-1);
    if (auto c = std::atomic_load(&conn_)) c->wake_reader();  // marks it dead; a ring reader sleeps on a futex
    if (fd >= 0) ::shutdown(fd, SHUT_RDWR);
  }

  bool is_local_addr() const {
    ParsedAddr pa = parse_addr(addr_);
    return pa.is_unix || pa.host == "127.0.0.1" || pa.host == "::1";
  }

  // Offer the server a shared-memory data path (see shm_conn.h).  Returns false only when the
  // socket stream itself can no longer be trusted; a polite refusal keeps the socket path.
  bool offer_shm(Conn& c) {
    std::string name = c.create_segment();
    if (name.empty()) return true;  // no /dev/shm: plain socket
    // "<ring segment>\n<doorbell page>\n<slot>": the page is shared by all receivers of this
    // transport, so that one wake reaches any subset of them (shm_conn.h, DoorbellPage)
    std::string hello = name;
    const bool with_page = doorbell_page_enabled() && page_ && slot_ >= 0 && slot_ < 32 && !page_->ensure().empty();
    if (with_page) hello += "\n" + page_->ensure() + "\n" + std::to_string(slot_);
    FrameHeader h{};
    h.magic = kMagic;
    h.type = kFrameHello;
    h.data_len = hello.size();
    std::string out((const char*)&h, sizeof(h));
    out += hello;
    bool io_ok = fd_write_all(c.fd(), out.data(), out.size());
    bool accepted = false;
    if (io_ok) {
      pollfd p{c.fd(), POLLIN, 0};
      io_ok = ::poll(&p, 1, 3000) > 0;
      FrameHeader a{};
      if (io_ok) io_ok = fd_read_all(c.fd(), &a, sizeof(a)) && a.magic == kMagic && a.type == kFrameAck;
      if (io_ok && a.data_len) {
        std::string skip(a.data_len, '\0');
        io_ok = a.data_len <= 4096 && fd_read_all(c.fd(), &skip[0], a.data_len);
      }
      accepted = io_ok && (a.code == 200 || a.code == 201);
      if (accepted && a.code == 201 && with_page) {
        c.set_tx_page(page_->seq(), slot_);  // the server sleeps on our page
        stats_->paged_conns.fetch_add(1, std::memory_order_relaxed);
      }
    }
    c.finish_create(accepted);
    if (accepted) stats_->shm_conns.fetch_add(1, std::memory_order_relaxed);
    return io_ok;
  }

  // Callers: the writer thread while it holds the `writing_` token, or stop() after
  // it joined the writer — never concurrently.
  void reap_reader() {
    if (reader_.joinable()) reader_.join();
    if (reader_fd_ >= 0) {
      ::close(reader_fd_);
      reader_fd_ = -1;
    }
  }

  void writer_loop() {
    for (;;) {
      std::shared_ptr<Outgoing> m;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stopping_ || (!q_.empty() && !writing_); });
        if (stopping_) return;
        m = std::move(q_.front());
        q_.pop_front();
        writing_ = true;
      }
      if (!ensure_connected()) {
        m->ack->set_raw("failed to connect to party '" + name_ + "' at " + addr_ + " after " +
                            std::to_string(retry_.max_attempts) + " attempts",
                        true);
      } else {
        transmit(std::move(m));
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        writing_ = false;
      }
      cv_.notify_all();
    }
  }

  // Writes one frame on the live connection. Caller owns the `writing_` token.
  void transmit(std::shared_ptr<Outgoing> m) {
    FrameHeader h{};
    h.magic = kMagic;
    h.type = kFrameData;
    h.flags = m->is_error ? 1 : 0;
    h.msg_id = m->id;
    h.job_len = (uint32_t)job_.size();
    h.up_len = (uint32_t)m->up.size();
    h.down_len = (uint32_t)m->down.size();
    h.data_len = m->data.size();
    size_t nbytes = m->data.size();
    std::string head;
    head.reserve(sizeof(h) + job_.size() + m->up.size() + m->down.size() + (nbytes <= 4096 ? nbytes : 0));
    head.append((const char*)&h, sizeof(h));
    head += job_;
    head += m->up;
    head += m->down;
    if (nbytes <= 4096) head += m->data;  // one write (one syscall on the socket path) for small messages
    RefPtr ack = m->ack;
    uint64_t id = m->id;
    {
      std::lock_guard<std::mutex> lk(mu_);
      inflight_[id] = InFlight{ack, m->deadline, retry_.max_resends > 0 ? m : nullptr};
    }
    auto c = std::atomic_load(&conn_);
    bool ok = c && fd_.load() == c->fd() && c->write_all(head.data(), head.size()) &&
              (nbytes <= 4096 || c->write_all(m->data.data(), nbytes));
    if (!ok) {
      bool requeued = false;
      {
        std::lock_guard<std::mutex> lk(mu_);
        inflight_.erase(id);
        if (m->resends < retry_.max_resends && !stopping_) {
          m->resends++;
          q_.push_front(m);
          requeued = true;
        }
      }
      close_conn();
      if (!requeued) ack->set_raw("connection to party '" + name_ + "' lost while sending", true);
    } else {
      stats_->send_op_count.fetch_add(1, std::memory_order_relaxed);
      stats_->bytes_sent.fetch_add(nbytes, std::memory_order_relaxed);
    }
  }

  void reader_loop(std::shared_ptr<Conn> c) {
    const int fd = c->fd();
    c->set_latency_critical(false);  // nobody waits for an acknowledgement on the hot path
    for (;;) {
      int pr = c->wait_readable(100);
      if (pr < 0) break;
      sweep_timeouts();
      if (pr == 0) {
        if (fd_.load() != fd) break;
        continue;
      }
      FrameHeader h{};
      if (!c->read_all(&h, sizeof(h)) || h.magic != kMagic || h.type != kFrameAck) break;
      if (h.data_len > (1u << 20)) break;  // an ack carries a short status string, never a payload
      std::string result(h.data_len, '\0');
      if (h.data_len && !c->read_all(&result[0], h.data_len)) break;
      RefPtr ack;
      {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = inflight_.find(h.msg_id);
        if (it != inflight_.end()) {
          ack = it->second.first;
          inflight_.erase(it);
        }
      }
      if (!ack) continue;
      if (h.code == 200)
        ack->set_raw(std::move(result), false);
      else
        ack->set_raw(std::move(result), true);
    }
    c->mark_dead();
    if (fd_.load() == fd) close_conn();
    fail_inflight("connection to party '" + name_ + "' closed before the message was acknowledged");
  }

  void sweep_timeouts() {
    auto now = std::chrono::steady_clock::now();
    std::vector<RefPtr> dead;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto it = inflight_.begin(); it != inflight_.end();) {
        if (it->second.second < now) {
          dead.push_back(it->second.first);
          it = inflight_.erase(it);
        } else {
          ++it;
        }
      }
    }
    for (auto& r : dead) r->set_raw("timed out waiting for party '" + name_ + "' to acknowledge", true);
  }

  void fail_inflight(const std::string& why, bool allow_resend = true) {
    std::map<uint64_t, InFlight> dead;
    bool requeued = false;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dead.swap(inflight_);
      if (allow_resend && !stopping_) {
        // un-acknowledged messages go back to the head of the queue in their original order
        for (auto it = dead.rbegin(); it != dead.rend();) {
          auto& f = it->second;
          if (f.msg && f.msg->resends < retry_.max_resends) {
            f.msg->resends++;
            q_.push_front(f.msg);
            requeued = true;
            it = decltype(it)(dead.erase(std::next(it).base()));
          } else {
            ++it;
          }
        }
      }
    }
    if (requeued) cv_.notify_all();
    for (auto& kv : dead) kv.second.first->set_raw(why, true);
  }

  void fail_all(const std::string& why) {
    fail_inflight(why, false);
    std::deque<std::shared_ptr<Outgoing>> dead;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dead.swap(q_);
    }
    for (auto& m : dead) m->ack->set_raw(why, true);
  }

  std::string name_, addr_, job_;
  RetryPolicy retry_;
  double timeout_s_;
  TransportStats* stats_;
  DoorbellPageOwner* page_ = nullptr;  // the owning transport's doorbell page (outlives the peer)
  int slot_ = -1;                      // this peer's bit in it
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::shared_ptr<Outgoing>> q_;
  struct InFlight {
    RefPtr first;                                   // the ack future
    std::chrono::steady_clock::time_point second;   // deadline
    std::shared_ptr<Outgoing> msg;                  // kept only when resends are enabled
  };
  std::map<uint64_t, InFlight> inflight_;
  uint64_t next_id_ = 1;
  bool stopping_ = false;
  bool writing_ = false;
  std::atomic<int> fd_{-1};
  std::shared_ptr<Conn> conn_;  // byte channel of the live connection (atomic_load/atomic_store)
  bool shm_ok_ = true;          // cleared when a shm offer broke a connection attempt
  int reader_fd_ = -1;  // descriptor the running reader_ was started on (see reap_reader)
  std::thread writer_, reader_;
};

// ---------------------------------------------------------------------------
class SocketTransport {
 public:
  SocketTransport(std::string party, std::string job, std::shared_ptr<Mailbox> mailbox)
      : party_(std::move(party)), job_(std::move(job)), mailbox_(std::move(mailbox)) {}
  ~SocketTransport() { stop(); }

  // Receive-side payload cap (bytes); frames announcing more are refused with code 413.
  void set_max_message_bytes(uint64_t n) { max_message_bytes_.store(n ? n : kDefaultMaxMessageBytes); }
  uint64_t max_message_bytes() const { return max_message_bytes_.load(); }

  // Returns "" on success, else an error description (port in use, ...).
  std::string listen(const std::string& addr) {
    ParsedAddr pa = parse_addr(addr);
    int fd;
    if (pa.is_unix) {
      fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
      sockaddr_un sa{};
      sa.sun_family = AF_UNIX;
      strncpy(sa.sun_path, pa.path.c_str(), sizeof(sa.sun_path) - 1);
      ::unlink(pa.path.c_str());
      if (::bind(fd, (sockaddr*)&sa, sizeof(sa)) != 0) {
        std::string e = std::string("bind(") + pa.path + "): " + strerror(errno);
        ::close(fd);
        return e;
      }
      unix_path_ = pa.path;
    } else {
      fd = ::socket(AF_INET, SOCK_STREAM, 0);
      int one = 1;
      setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
      sockaddr_in sa{};
      sa.sin_family = AF_INET;
      sa.sin_addr.s_addr = htonl(INADDR_ANY);
      sa.sin_port = htons((uint16_t)atoi(pa.port.c_str()));
      if (::bind(fd, (sockaddr*)&sa, sizeof(sa)) != 0) {
        std::string e = "Failed to listen on port " + pa.port + ": " + strerror(errno);
        ::close(fd);
        return e;
      }
    }
    if (::listen(fd, 128) != 0) {
      std::string e = std::string("listen: ") + strerror(errno);
      ::close(fd);
      return e;
    }
    listen_fd_.store(fd);
    accept_thread_ = std::thread([this, fd] {
      try {
        accept_loop(fd);
      } catch (...) {
      }
    });
    return "";
  }

  void add_peer(const std::string& name, const std::string& addr, RetryPolicy retry, double timeout_s) {
    std::lock_guard<std::mutex> lk(mu_);
    int slot = next_slot_ < 32 ? next_slot_++ : -1;  // one wake bit per peer; beyond 32 the ring's own word
    peers_[name] = std::make_unique<Peer>(name, addr, job_, retry, timeout_s, &stats_, &page_, slot);
  }

  // GIL-free hot path; `ack` resolves raw: b"OK" or error text.
  bool send(const std::string& dest, std::string up, std::string down, std::string data, bool is_error,
            RefPtr ack) {
    Peer* p = nullptr;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = peers_.find(dest);
      if (it != peers_.end()) p = it->second.get();
    }
    if (!p) return false;
    p->push(std::move(up), std::move(down), std::move(data), is_error, std::move(ack));
    return true;
  }

  bool drain(double timeout_s) {
    std::vector<Peer*> ps;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& kv : peers_) ps.push_back(kv.second.get());
    }
    bool ok = true;
    for (auto* p : ps) ok = p->drain(timeout_s) && ok;
    return ok;
  }

  void stop() {
    if (stopped_.exchange(true)) return;
    int lfd = listen_fd_.exchange(-1);
    if (lfd >= 0) ::shutdown(lfd, SHUT_RDWR);  // wakes accept(); close only after the join
    if (accept_thread_.joinable()) accept_thread_.join();
    if (lfd >= 0) ::close(lfd);
    std::map<std::string, std::unique_ptr<Peer>> peers;
    std::vector<ServerConn> conns;
    {
      std::lock_guard<std::mutex> lk(mu_);
      peers.swap(peers_);
      conns.swap(conns_);
    }
    for (auto& kv : peers) kv.second->stop();
    for (auto& c : conns) {
      if (c.conn) c.conn->wake_reader();
      ::shutdown(c.fd, SHUT_RDWR);
    }
    for (auto& c : conns) {
      if (c.th.joinable()) c.th.join();
      ::close(c.fd);
    }
    if (!unix_path_.empty()) ::unlink(unix_path_.c_str());
  }

  const TransportStats& stats() const { return stats_; }
  const std::string& job() const { return job_; }

 private:
  void accept_loop(int lfd) {
    for (;;) {
      pollfd p{lfd, POLLIN, 0};
      int pr = ::poll(&p, 1, 100);  // bounded wait: stop() does not rely on shutdown() waking accept()
      if (stopped_.load()) return;
      if (pr < 0 && errno != EINTR) return;
      if (pr <= 0) continue;
      if (p.revents & (POLLERR | POLLHUP | POLLNVAL)) return;
      int cfd = ::accept(lfd, nullptr, nullptr);
      if (cfd < 0) {
        if (errno == EINTR) continue;
        return;
      }
      if (stopped_.load()) {
        ::close(cfd);
        return;
      }
      int one = 1;
      setsockopt(cfd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      std::lock_guard<std::mutex> lk(mu_);
      // reap connections whose peer went away (a party that restarts reconnects: without this
      // every reconnect would leave a thread and a descriptor behind until stop())
      for (auto it = conns_.begin(); it != conns_.end();) {
        if (it->done->load(std::memory_order_acquire)) {
          if (it->th.joinable()) it->th.join();
          ::close(it->fd);
          it = conns_.erase(it);
        } else {
          ++it;
        }
      }
      ServerConn c;
      c.fd = cfd;
      c.done = std::make_shared<std::atomic<bool>>(false);
      c.conn = std::make_shared<Conn>(cfd);
      auto done = c.done;
      auto conn = c.conn;
      c.th = std::thread([this, cfd, done, conn] {
        try {
          serve(*conn);
        } catch (...) {  // nothing may escape a std::thread (std::terminate): drop the connection
          stats_.rejected.fetch_add(1, std::memory_order_relaxed);
        }
        // tell the peer now (its writer may be parked on a full shared-memory ring); the descriptor
        // itself is closed by the accept loop's reaper
        ::shutdown(cfd, SHUT_RDWR);
        done->store(true, std::memory_order_release);
      });
      conns_.push_back(std::move(c));
    }
  }

  void serve(Conn& c) {
    const int fd = c.fd();
    for (;;) {
      FrameHeader h{};
      if (!c.read_all(&h, sizeof(h))) return;
      if (h.magic != kMagic) return;
      if (h.type == kFrameHello) {
        // shm upgrade offer: data = segment name.  Answered on the socket; afterwards both
        // directions of this connection use the rings (or stay on the socket when refused).
        if (h.data_len == 0 || h.data_len > 255 || c.uses_shm()) return;
        std::string hello(h.data_len, '\0');
        if (!c.read_all(&hello[0], h.data_len)) return;
        // "<ring segment>[\n<doorbell page>\n<slot>]"
        std::string name = hello, page;
        int slot = -1;
        size_t nl = hello.find('\n');
        if (nl != std::string::npos) {
          name = hello.substr(0, nl);
          size_t nl2 = hello.find('\n', nl + 1);
          if (nl2 != std::string::npos) {
            page = hello.substr(nl + 1, nl2 - nl - 1);
            const std::string s = hello.substr(nl2 + 1);
            if (!s.empty() && s.size() <= 2 && s.find_first_not_of("0123456789") == std::string::npos) slot = std::atoi(s.c_str());
          }
        }
        FrameHeader a{};
        a.magic = kMagic;
        a.type = kFrameAck;
        bool ok = shm_enabled();
        // reply BEFORE switching the receive side (the client reads the answer on the socket)
        ok = ok && c.attach_segment(name);
        if (ok) stats_.shm_conns.fetch_add(1, std::memory_order_relaxed);
        const bool paged = ok && !page.empty() && doorbell_page_enabled() && c.attach_rx_page(page, slot);
        if (paged) stats_.paged_conns.fetch_add(1, std::memory_order_relaxed);
        a.code = ok ? (paged ? 201 : 200) : 501;
        if (!fd_write_all(fd, &a, sizeof(a))) return;
        continue;
      }
      if (h.type != kFrameData) return;
      // Every length comes from the peer: bound them BEFORE allocating (an unauthenticated
      // 40-byte header must not be able to abort or exhaust the party).  The payload cap is
      // messages_max_size_in_bytes (default 500 MiB, fed/proxy/grpc/grpc_options.py:28-29).
      if (h.job_len > kMaxKeyLen || h.up_len > kMaxKeyLen || h.down_len > kMaxKeyLen ||
          h.data_len > max_message_bytes_.load(std::memory_order_relaxed)) {
        stats_.rejected.fetch_add(1, std::memory_order_relaxed);
        FrameHeader a{};
        a.magic = kMagic;
        a.type = kFrameAck;
        a.msg_id = h.msg_id;
        a.code = 413;
        std::string why = "message exceeds the receiver's limits (messages_max_size_in_bytes)";
        a.data_len = why.size();
        std::string out((const char*)&a, sizeof(a));
        out += why;
        c.write_all(out.data(), out.size());
        return;  // the stream cannot be resynchronised past an oversize frame: drop the connection
      }
      std::string job(h.job_len, '\0'), up(h.up_len, '\0'), down(h.down_len, '\0');
      if (h.job_len && !c.read_all(&job[0], h.job_len)) return;
      if (h.up_len && !c.read_all(&up[0], h.up_len)) return;
      if (h.down_len && !c.read_all(&down[0], h.down_len)) return;
      std::string data(h.data_len, '\0');
      if (h.data_len && !c.read_all(&data[0], h.data_len)) return;

      FrameHeader a{};
      a.magic = kMagic;
      a.type = kFrameAck;
      a.msg_id = h.msg_id;
      std::string result;
      if (!job_.empty() && job != job_) {
        // Same contract as the reference receiver: refuse foreign-job traffic.
        a.code = 417;
        result = "JobName mis-match, expected " + job_ + ", got " + job + ".";
        stats_.rejected.fetch_add(1, std::memory_order_relaxed);
      } else {
        stats_.receive_op_count.fetch_add(1, std::memory_order_relaxed);
        stats_.bytes_received.fetch_add(data.size(), std::memory_order_relaxed);
        mailbox_->deliver_raw(up, down, std::move(data), (h.flags & 1) != 0);
        a.code = 200;
        result = "OK";
      }
      a.data_len = result.size();
      std::string out((const char*)&a, sizeof(a));
      out += result;
      if (!c.write_all(out.data(), out.size())) return;
    }
  }

  static constexpr uint32_t kMaxKeyLen = 4096;                       // job name / seq-id strings
  static constexpr uint64_t kDefaultMaxMessageBytes = 500ull << 20;  // the reference's gRPC default
  std::atomic<uint64_t> max_message_bytes_{kDefaultMaxMessageBytes};
  std::string party_, job_;
  std::shared_ptr<Mailbox> mailbox_;
  std::mutex mu_;
  std::map<std::string, std::unique_ptr<Peer>> peers_;
  struct ServerConn {
    int fd = -1;
    std::thread th;
    std::shared_ptr<std::atomic<bool>> done;
    std::shared_ptr<Conn> conn;  // stop() wakes a reader that sleeps on its ring
  };
  std::vector<ServerConn> conns_;
  DoorbellPageOwner page_;  // sender role: the word all same-host receivers of this transport sleep on
  int next_slot_ = 0;
  std::atomic<int> listen_fd_{-1};
  std::atomic<bool> stopped_{false};
  std::thread accept_thread_;
  std::string unix_path_;
  TransportStats stats_;
};

}  // namespace fedcore
