// Swarm rendezvous / membership board (C ABI): a small TCP key-value service with atomic counters and peer heartbeats.
//
// The reference discovers its DiLoCo workers through hivemind's DHT, whose transport is the Go libp2p daemon `p2pd`
// (native-code census).  On one NVLink box the workers only need (a) a place to publish progress
// records, (b) arrival counters for the outer-step handshake and (c) liveness with expiry - this file is that service,
// independent of torch.distributed so that launchers, monitors and workers of different torchrun jobs can share it
// (`run_training.sh`: "initial peer").  One server thread multiplexes all clients with poll(); requests are tiny and
// handled inline.
//
// Wire format (little endian):   request  = u8 op | u32 klen | u32 vlen | key | value
//                                response = u8 status (0 ok, 1 not found, 2 bad request) | u32 len | payload
#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <stdint.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define ODB_API extern "C" __attribute__((visibility("default")))

namespace {

enum Op : uint8_t { kSet = 1, kGet = 2, kAdd = 3, kBeat = 4, kPeers = 5, kDel = 6, kCount = 7 };
constexpr uint32_t kMaxKey = 1u << 12, kMaxVal = 1u << 24;

int64_t now_ms() {
  return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool read_all(int fd, void* buf, size_t n) {
  uint8_t* p = static_cast<uint8_t*>(buf);
  while (n) {
    const ssize_t r = ::recv(fd, p, n, 0);
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

bool write_all(int fd, const void* buf, size_t n) {
  const uint8_t* p = static_cast<const uint8_t*>(buf);
  while (n) {
    const ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += r;
    n -= (size_t)r;
  }
  return true;
}

bool send_reply(int fd, uint8_t status, const std::string& payload) {
  uint8_t hdr[5];
  const uint32_t len = (uint32_t)payload.size();
  hdr[0] = status;
  memcpy(hdr + 1, &len, 4);
  return write_all(fd, hdr, 5) && (len == 0 || write_all(fd, payload.data(), len));
}

struct Server {
  int listen_fd = -1;
  int port = 0;
  std::thread thread;
  std::atomic<bool> stop{false};
  std::mutex mu;                                   // the maps are also read by odb_rdv_server_num_keys
  std::map<std::string, std::string> kv;
  std::map<std::string, int64_t> beat_deadline;    // peer id -> time (ms) after which it no longer counts as alive

  // one request from `fd`; false = drop the connection
  bool serve_one(int fd) {
    uint8_t hdr[9];
    if (!read_all(fd, hdr, 9)) return false;
    const uint8_t op = hdr[0];
    uint32_t klen, vlen;
    memcpy(&klen, hdr + 1, 4);
    memcpy(&vlen, hdr + 5, 4);
    if (klen > kMaxKey || vlen > kMaxVal) return false;
    std::string key(klen, '\0'), val(vlen, '\0');
    if (klen && !read_all(fd, &key[0], klen)) return false;
    if (vlen && !read_all(fd, &val[0], vlen)) return false;
    std::lock_guard<std::mutex> g(mu);
    switch (op) {
      case kSet:
        kv[key] = std::move(val);
        return send_reply(fd, 0, "");
      case kGet: {
        auto it = kv.find(key);
        return it == kv.end() ? send_reply(fd, 1, "") : send_reply(fd, 0, it->second);
      }
      case kAdd: {
        if (vlen != 8) return send_reply(fd, 2, "");
        int64_t delta, cur = 0;
        memcpy(&delta, val.data(), 8);
        auto it = kv.find(key);
        if (it != kv.end() && it->second.size() == 8) memcpy(&cur, it->second.data(), 8);
        cur += delta;
        kv[key] = std::string(reinterpret_cast<const char*>(&cur), 8);
        return send_reply(fd, 0, kv[key]);
      }
      case kBeat: {
        if (vlen != 4) return send_reply(fd, 2, "");
        uint32_t ttl;
        memcpy(&ttl, val.data(), 4);
        beat_deadline[key] = now_ms() + (int64_t)ttl;
        return send_reply(fd, 0, "");
      }
      case kPeers: {
        const int64_t t = now_ms();
        std::string out;
        for (auto it = beat_deadline.begin(); it != beat_deadline.end();) {
          if (it->second < t) {
            it = beat_deadline.erase(it);          // expired: the peer stopped beating
          } else {
            if (!out.empty()) out.push_back('\n');
            out += it->first;
            ++it;
          }
        }
        return send_reply(fd, 0, out);
      }
      case kDel:
        return send_reply(fd, kv.erase(key) ? 0 : 1, "");
      case kCount: {                               // number of keys that start with `key`
        int64_t n = 0;
        for (auto it = kv.lower_bound(key); it != kv.end() && it->first.compare(0, key.size(), key) == 0; ++it) ++n;
        return send_reply(fd, 0, std::string(reinterpret_cast<const char*>(&n), 8));
      }
      default:
        return send_reply(fd, 2, "");
    }
  }

  void run() {
    std::vector<pollfd> fds;
    fds.push_back({listen_fd, POLLIN, 0});
    while (!stop.load()) {
      const int n = ::poll(fds.data(), (nfds_t)fds.size(), 100);
      if (n <= 0) continue;
      if (fds[0].revents & POLLIN) {
        const int c = ::accept(listen_fd, nullptr, nullptr);
        if (c >= 0) {
          int one = 1;
          ::setsockopt(c, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
          // a client that stalls in the middle of a request must not freeze the single-threaded board: bound every
          // blocking recv / send on the connection; a timed-out request drops that connection only
          struct timeval tv;
          tv.tv_sec = 2;
          tv.tv_usec = 0;
          ::setsockopt(c, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
          ::setsockopt(c, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
          fds.push_back({c, POLLIN, 0});
        }
      }
      for (size_t i = 1; i < fds.size();) {
        bool keep = true;
        if (fds[i].revents & (POLLERR | POLLHUP | POLLNVAL)) keep = false;
        else if (fds[i].revents & POLLIN) keep = serve_one(fds[i].fd);
        if (!keep) {
          ::close(fds[i].fd);
          fds.erase(fds.begin() + (long)i);
        } else {
          ++i;
        }
      }
    }
    for (size_t i = 1; i < fds.size(); ++i) ::close(fds[i].fd);
  }
};

struct Client {
  int fd = -1;
  std::mutex mu;

  // returns status (0/1/2) or -1 on I/O failure; payload in `out`
  int call(uint8_t op, const void* key, uint32_t klen, const void* val, uint32_t vlen, std::string* out) {
    std::lock_guard<std::mutex> g(mu);
    uint8_t hdr[9];
    hdr[0] = op;
    memcpy(hdr + 1, &klen, 4);
    memcpy(hdr + 5, &vlen, 4);
    if (!write_all(fd, hdr, 9) || (klen && !write_all(fd, key, klen)) || (vlen && !write_all(fd, val, vlen))) return -1;
    uint8_t rh[5];
    if (!read_all(fd, rh, 5)) return -1;
    uint32_t len;
    memcpy(&len, rh + 1, 4);
    if (len > kMaxVal) return -1;
    std::string payload(len, '\0');
    if (len && !read_all(fd, &payload[0], len)) return -1;
    if (out) *out = std::move(payload);
    return rh[0];
  }
};

long long copy_out(const std::string& s, void* buf, long long cap) {
  if ((long long)s.size() > cap) return -(long long)s.size() - 2;     // caller's buffer is too small: -(needed) - 2
  if (!s.empty()) memcpy(buf, s.data(), s.size());
  return (long long)s.size();
}

}  // namespace

// ------------------------------------------------------------------------------------------------ server
// port 0 = ephemeral; the bound port is written to *bound_port.  Returns an opaque handle (nullptr on failure).
ODB_API void* odb_rdv_server_start(int port, int* bound_port) {
  const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) return nullptr;
  int one = 1;
  ::setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_ANY);
  addr.sin_port = htons((uint16_t)port);
  if (::bind(fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || ::listen(fd, 128) != 0) {
    ::close(fd);
    return nullptr;
  }
  socklen_t len = sizeof(addr);
  ::getsockname(fd, reinterpret_cast<sockaddr*>(&addr), &len);
  auto* s = new Server();
  s->listen_fd = fd;
  s->port = ntohs(addr.sin_port);
  if (bound_port) *bound_port = s->port;
  s->thread = std::thread([s] { s->run(); });
  return s;
}

ODB_API long long odb_rdv_server_num_keys(void* h) {
  auto* s = static_cast<Server*>(h);
  std::lock_guard<std::mutex> g(s->mu);
  return (long long)s->kv.size();
}

ODB_API void odb_rdv_server_stop(void* h) {
  auto* s = static_cast<Server*>(h);
  if (!s) return;
  s->stop.store(true);
  if (s->thread.joinable()) s->thread.join();
  ::close(s->listen_fd);
  delete s;
}

// ------------------------------------------------------------------------------------------------ client
ODB_API void* odb_rdv_connect(const char* host, int port, int timeout_ms) {
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  const std::string p = std::to_string(port);
  if (::getaddrinfo(host, p.c_str(), &hints, &res) != 0 || !res) return nullptr;
  const int64_t deadline = now_ms() + timeout_ms;
  int fd = -1;
  do {                                             // the server of a fresh job may not be listening yet: retry
    fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (fd >= 0 && ::connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
    if (fd >= 0) ::close(fd);
    fd = -1;
    ::usleep(20 * 1000);
  } while (now_ms() < deadline);
  ::freeaddrinfo(res);
  if (fd < 0) return nullptr;
  int one = 1;
  ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  auto* c = new Client();
  c->fd = fd;
  return c;
}

ODB_API void odb_rdv_close(void* h) {
  auto* c = static_cast<Client*>(h);
  if (!c) return;
  ::close(c->fd);
  delete c;
}

ODB_API int odb_rdv_set(void* h, const void* key, int klen, const void* val, long long vlen) {
  return static_cast<Client*>(h)->call(kSet, key, (uint32_t)klen, val, (uint32_t)vlen, nullptr);
}

// returns the value length (copied into buf), -1 if the key does not exist, -2 on I/O failure, -(needed)-2 if cap is too small
ODB_API long long odb_rdv_get(void* h, const void* key, int klen, void* buf, long long cap) {
  std::string out;
  const int st = static_cast<Client*>(h)->call(kGet, key, (uint32_t)klen, nullptr, 0, &out);
  if (st < 0) return -2;
  if (st == 1) return -1;
  return copy_out(out, buf, cap);
}

// atomic fetch-add on an int64 counter (created at 0); returns the new value through *result
ODB_API int odb_rdv_add(void* h, const void* key, int klen, long long delta, long long* result) {
  std::string out;
  const int64_t d = delta;
  const int st = static_cast<Client*>(h)->call(kAdd, key, (uint32_t)klen, &d, 8, &out);
  if (st != 0 || out.size() != 8) return st < 0 ? -2 : 2;
  int64_t v;
  memcpy(&v, out.data(), 8);
  *result = v;
  return 0;
}

ODB_API int odb_rdv_del(void* h, const void* key, int klen) {
  return static_cast<Client*>(h)->call(kDel, key, (uint32_t)klen, nullptr, 0, nullptr);
}

// number of keys with the given prefix (e.g. arrivals of one outer step)
ODB_API long long odb_rdv_count(void* h, const void* prefix, int plen) {
  std::string out;
  const int st = static_cast<Client*>(h)->call(kCount, prefix, (uint32_t)plen, nullptr, 0, &out);
  if (st != 0 || out.size() != 8) return -2;
  int64_t v;
  memcpy(&v, out.data(), 8);
  return v;
}

// "I am alive for the next ttl_ms milliseconds"
ODB_API int odb_rdv_beat(void* h, const void* peer, int plen, int ttl_ms) {
  const uint32_t ttl = (uint32_t)ttl_ms;
  return static_cast<Client*>(h)->call(kBeat, peer, (uint32_t)plen, &ttl, 4, nullptr);
}

// newline-separated ids of the peers whose heartbeat has not expired; same return convention as odb_rdv_get
ODB_API long long odb_rdv_peers(void* h, void* buf, long long cap) {
  std::string out;
  const int st = static_cast<Client*>(h)->call(kPeers, nullptr, 0, nullptr, 0, &out);
  if (st != 0) return -2;
  return copy_out(out, buf, cap);
}
