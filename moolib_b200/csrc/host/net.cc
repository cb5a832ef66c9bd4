// RpcCore: the message layer of the minimal control plane (see control.h).  One IO thread per Rpc, poll()-driven
// reads, blocking writes under a per-connection mutex, length-prefixed frames.
#include "control.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cerrno>
#include <random>
#include <sstream>
#include <stdexcept>

namespace mbh {

void trace_phase(const char* phase);
static void mbh_trace_phase_hook(const char* p) { trace_phase(p); }

namespace {
constexpr uint32_t kHello = 1, kRoute = 2;
constexpr size_t kMaxFrame = 1u << 30;

void setNoDelay(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
}

bool parseAddress(const std::string& address, std::string* host, std::string* port) {
  auto pos = address.rfind(':');
  if (pos == std::string::npos) return false;
  *host = address.substr(0, pos);
  *port = address.substr(pos + 1);
  if (host->empty()) *host = "0.0.0.0";
  return !port->empty();
}
}  // namespace

std::string randomName() {
  static const char* alphabet = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789";
  std::random_device rd;
  std::mt19937_64 g(((uint64_t)rd() << 32) ^ rd() ^ (uint64_t)Clock::now().time_since_epoch().count());
  std::string s;
  for (int i = 0; i < 16; ++i) s += alphabet[g() % 62];
  return s;
}

struct RpcCore::Conn {
  int fd = -1;
  std::string name;  // hub side: the peer's announced name
  std::string rbuf;
  std::mutex wmu;
  bool dead = false;
  ~Conn() {
    if (fd >= 0) ::close(fd);
  }
};

RpcCore::RpcCore() : name_(randomName()) {
  if (::pipe(wakeFd_) != 0) throw std::runtime_error("moolib_b200 Rpc: pipe() failed");
  fcntl(wakeFd_[0], F_SETFL, O_NONBLOCK);
  fcntl(wakeFd_[1], F_SETFL, O_NONBLOCK);
}

RpcCore::~RpcCore() { close(); }

void RpcCore::close() {
  if (stop_.exchange(true)) return;
  wake();
  if (threadStarted_ && thread_.joinable()) {
    if (std::this_thread::get_id() == thread_.get_id()) thread_.detach();
    else thread_.join();
  }
  std::lock_guard<std::mutex> l(mu_);
  conns_.clear();
  byName_.clear();
  hub_.reset();
  if (listenFd_ >= 0) ::close(listenFd_);
  listenFd_ = -1;
  for (int i = 0; i < 2; ++i)
    if (wakeFd_[i] >= 0) ::close(wakeFd_[i]), wakeFd_[i] = -1;
}

void RpcCore::wake() {
  if (wakeFd_[1] >= 0) {
    char c = 1;
    ssize_t r = ::write(wakeFd_[1], &c, 1);
    (void)r;
  }
}

void RpcCore::ensureThread() {
  if (!threadStarted_) {
    threadStarted_ = true;
    thread_ = std::thread([this] { ioLoop(); });
  }
}

void RpcCore::setName(const std::string& name) {
  std::shared_ptr<Conn> hub;
  {
    std::lock_guard<std::mutex> l(mu_);
    name_ = name;
    nameSet_ = true;
    hub = hub_;
  }
  if (hub) {
    Writer w;
    w.u32(kHello);
    w.str(name);
    writeFrame(*hub, w.b);
  }
}

std::string RpcCore::getName() {
  std::lock_guard<std::mutex> l(mu_);
  return name_;
}

void RpcCore::listen(const std::string& address) {
  std::string host, port;
  if (!parseAddress(address, &host, &port)) throw std::runtime_error("Rpc::listen: bad address '" + address + "'");
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  hints.ai_flags = AI_PASSIVE;
  if (getaddrinfo(host.c_str(), port.c_str(), &hints, &res) != 0 || !res)
    throw std::runtime_error("Rpc::listen: cannot resolve '" + address + "'");
  int fd = ::socket(res->ai_family, res->ai_socktype, res->ai_protocol);
  int one = 1;
  setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  if (fd < 0 || ::bind(fd, res->ai_addr, res->ai_addrlen) != 0 || ::listen(fd, 256) != 0) {
    int e = errno;
    freeaddrinfo(res);
    if (fd >= 0) ::close(fd);
    throw std::runtime_error("Rpc::listen: cannot listen on '" + address + "': " + std::strerror(e));
  }
  freeaddrinfo(res);
  fcntl(fd, F_SETFL, O_NONBLOCK);
  std::lock_guard<std::mutex> l(mu_);
  listenFd_ = fd;
  listening_ = true;
  ensureThread();
  wake();
}

namespace {
int tryConnect(const std::string& address) {
  std::string host, port;
  if (!parseAddress(address, &host, &port)) return -1;
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  if (getaddrinfo(host.c_str(), port.c_str(), &hints, &res) != 0 || !res) return -1;
  int fd = ::socket(res->ai_family, res->ai_socktype, res->ai_protocol);
  if (fd >= 0 && ::connect(fd, res->ai_addr, res->ai_addrlen) != 0) {
    ::close(fd);
    fd = -1;
  }
  freeaddrinfo(res);
  if (fd >= 0) setNoDelay(fd);
  return fd;
}
}  // namespace

void RpcCore::connect(const std::string& address) {
  std::string host, port;
  if (!parseAddress(address, &host, &port)) throw std::runtime_error("Rpc::connect: bad address '" + address + "'");
  // The reference connects lazily and keeps retrying (src/rpc.cc:628-638); so do we, from a helper thread so that
  // connect() never blocks the caller.
  auto self = shared_from_this();
  {
    std::lock_guard<std::mutex> l(mu_);
    ensureThread();
  }
  std::thread([self, address] {
    while (!self->stop_) {
      int fd = tryConnect(address);
      if (fd >= 0) {
        auto c = std::make_shared<Conn>();
        c->fd = fd;
        std::deque<Parked> pending;
        std::string name;
        {
          std::lock_guard<std::mutex> l(self->mu_);
          self->conns_.push_back(c);
          self->hub_ = c;
          pending.swap(self->outbox_);
          name = self->name_;
        }
        Writer w;
        w.u32(kHello);
        w.str(name);
        self->writeFrame(*c, w.b);
        for (auto& m : pending) self->route(m.dst, m.src, m.service, m.payload);
        self->wake();
        return;
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
  }).detach();
}

void RpcCore::handle(const std::string& service, Handler h) {
  std::unique_lock<std::shared_mutex> l(hmu_);
  handlers_[service] = std::move(h);
}

void RpcCore::unhandle(const std::string& service) {
  std::unique_lock<std::shared_mutex> l(hmu_);
  handlers_.erase(service);
}

bool RpcCore::writeFrame(Conn& c, const Bytes& body) {
  std::lock_guard<std::mutex> l(c.wmu);
  if (std::this_thread::get_id() != thread_.get_id()) mbh_trace_phase_hook("RpcCore::writeFrame(caller thread)");
  if (c.dead) return false;
  uint32_t len = (uint32_t)body.size();
  iovec iov[2] = {{&len, 4}, {const_cast<char*>(body.data()), body.size()}};
  size_t total = 4 + body.size(), done = 0;
  while (done < total) {
    msghdr mh{};
    iovec cur[2];
    int n = 0;
    size_t skip = done;
    for (auto& v : iov) {
      if (skip >= v.iov_len) {
        skip -= v.iov_len;
        continue;
      }
      cur[n].iov_base = static_cast<char*>(v.iov_base) + skip;
      cur[n].iov_len = v.iov_len - skip;
      skip = 0;
      ++n;
    }
    mh.msg_iov = cur;
    mh.msg_iovlen = n;
    ssize_t r = ::sendmsg(c.fd, &mh, MSG_NOSIGNAL);
    if (r < 0) {
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) {
        pollfd p{c.fd, POLLOUT, 0};
        ::poll(&p, 1, 100);
        continue;
      }
      c.dead = true;
      return false;
    }
    done += (size_t)r;
  }
  ++sent_;
  return true;
}

void RpcCore::send(const std::string& dst, const std::string& service, const Bytes& payload) {
  route(dst, getName(), service, payload);
}

void RpcCore::deliverLocal(const std::string& src, const std::string& service, const Bytes& payload) {
  std::shared_lock<std::shared_mutex> l(hmu_);  // held while the handler runs (see control.h)
  auto i = handlers_.find(service);
  if (i == handlers_.end()) return;
  ++received_;
  try {
    i->second(src, payload);
  } catch (const std::exception&) {
    // a malformed control message must not take the IO thread down
  }
}

void RpcCore::route(const std::string& dst, const std::string& src, const std::string& service, const Bytes& payload) {
  std::shared_ptr<Conn> target;
  bool local = false;
  {
    std::lock_guard<std::mutex> l(mu_);
    if (dst == name_) {
      local = true;
    } else if (listening_) {
      auto i = byName_.find(dst);
      if (i != byName_.end() && !i->second->dead) {
        target = i->second;
      } else {
        parked_.push_back(Parked{Clock::now(), dst, src, service, payload});
        return;
      }
    } else if (hub_ && !hub_->dead) {
      target = hub_;
    } else {
      outbox_.push_back(Parked{Clock::now(), dst, src, service, payload});
      return;
    }
  }
  if (local) {
    deliverLocal(src, service, payload);
    return;
  }
  Writer w;
  w.u32(kRoute);
  w.str(dst);
  w.str(src);
  w.str(service);
  w.str(payload);
  writeFrame(*target, w.b);
}

void RpcCore::onFrame(Conn& c, const char* data, size_t len) {
  Reader r(data, len);
  uint32_t type = r.u32();
  if (type == kHello) {
    std::string name = r.str();
    std::vector<Parked> flush;
    {
      std::lock_guard<std::mutex> l(mu_);
      if (!c.name.empty()) {
        auto i = byName_.find(c.name);
        if (i != byName_.end() && i->second.get() == &c) byName_.erase(i);
      }
      c.name = name;
      for (auto& sp : conns_)
        if (sp.get() == &c) byName_[name] = sp;
      for (auto i = parked_.begin(); i != parked_.end();) {
        if (i->dst == name) {
          flush.push_back(std::move(*i));
          i = parked_.erase(i);
        } else {
          ++i;
        }
      }
    }
    for (auto& m : flush) route(m.dst, m.src, m.service, m.payload);
  } else if (type == kRoute) {
    std::string dst = r.str(), src = r.str(), service = r.str();
    Bytes payload = r.str();
    route(dst, src, service, payload);
  }
}

void RpcCore::ioLoop() {
  std::vector<pollfd> pfds;
  std::vector<std::shared_ptr<Conn>> snapshot;
  while (!stop_) {
    pfds.clear();
    snapshot.clear();
    int lfd;
    {
      std::lock_guard<std::mutex> l(mu_);
      lfd = listenFd_;
      // drop dead connections
      for (auto i = conns_.begin(); i != conns_.end();) {
        if ((*i)->dead) {
          if (!(*i)->name.empty()) {
            auto j = byName_.find((*i)->name);
            if (j != byName_.end() && j->second == *i) byName_.erase(j);
          }
          if (hub_ == *i) hub_.reset();
          i = conns_.erase(i);
        } else {
          ++i;
        }
      }
      snapshot = conns_;
      // expire parked messages
      auto now = Clock::now();
      auto ttl = std::chrono::duration<double>(timeoutSeconds_);
      while (!parked_.empty() && now - parked_.front().t > ttl) parked_.pop_front();
    }
    pfds.push_back({wakeFd_[0], POLLIN, 0});
    if (lfd >= 0) pfds.push_back({lfd, POLLIN, 0});
    for (auto& c : snapshot) pfds.push_back({c->fd, POLLIN, 0});
    int rc = ::poll(pfds.data(), (nfds_t)pfds.size(), 200);
    if (rc < 0 && errno != EINTR) break;
    if (stop_) break;
    size_t k = 0;
    if (pfds[k].revents & POLLIN) {
      char buf[64];
      while (::read(wakeFd_[0], buf, sizeof(buf)) > 0) {
      }
    }
    ++k;
    if (lfd >= 0) {
      if (pfds[k].revents & POLLIN) {
        while (true) {
          int fd = ::accept(lfd, nullptr, nullptr);
          if (fd < 0) break;
          setNoDelay(fd);
          auto c = std::make_shared<Conn>();
          c->fd = fd;
          std::lock_guard<std::mutex> l(mu_);
          conns_.push_back(c);
        }
      }
      ++k;
    }
    for (size_t i = 0; i < snapshot.size(); ++i, ++k) {
      auto& c = *snapshot[i];
      if (!(pfds[k].revents & (POLLIN | POLLHUP | POLLERR))) continue;
      char buf[65536];
      while (true) {
        ssize_t n = ::recv(c.fd, buf, sizeof(buf), MSG_DONTWAIT);
        if (n > 0) {
          c.rbuf.append(buf, (size_t)n);
          if ((size_t)n < sizeof(buf)) break;
        } else if (n == 0) {
          c.dead = true;
          break;
        } else {
          if (errno == EINTR) continue;
          if (errno != EAGAIN && errno != EWOULDBLOCK) c.dead = true;
          break;
        }
      }
      size_t off = 0;
      while (c.rbuf.size() - off >= 4) {
        uint32_t len;
        std::memcpy(&len, c.rbuf.data() + off, 4);
        if (len > kMaxFrame) {
          c.dead = true;
          break;
        }
        if (c.rbuf.size() - off - 4 < len) break;
        try {
          onFrame(c, c.rbuf.data() + off + 4, len);
        } catch (const std::exception&) {
          c.dead = true;
          break;
        }
        off += 4 + (size_t)len;
      }
      if (off) c.rbuf.erase(0, off);
    }
  }
}

std::string RpcCore::debugInfo() {
  std::lock_guard<std::mutex> l(mu_);
  std::ostringstream os;
  os << "Rpc '" << name_ << "' " << (listening_ ? "hub" : "client") << " connections=" << conns_.size()
     << " sent=" << sent_ << " received=" << received_ << " parked=" << parked_.size();
  return os.str();
}

// ---- FutureState ---------------------------------------------------------------------------------------------------
void FutureState::setResult(Bytes v) {
  std::function<void()> cb;
  {
    std::lock_guard<std::mutex> l(mu);
    if (flags) return;
    value = std::move(v);
    flags |= 1;
    cb.swap(onDone);
  }
  cv.notify_all();
  if (cb) cb();
}
void FutureState::setError(std::string e) {
  std::function<void()> cb;
  {
    std::lock_guard<std::mutex> l(mu);
    if (flags) return;
    error = std::move(e);
    flags |= 2;
    cb.swap(onDone);
  }
  cv.notify_all();
  if (cb) cb();
}
void FutureState::cancel() {
  std::function<void()> cb;
  {
    std::lock_guard<std::mutex> l(mu);
    if (flags) return;
    flags |= 4;
    cb.swap(onDone);
  }
  cv.notify_all();
  if (cb) cb();
}
bool FutureState::done() {
  std::lock_guard<std::mutex> l(mu);
  return flags != 0;
}
bool FutureState::wait(double seconds) {
  std::unique_lock<std::mutex> l(mu);
  if (seconds < 0) {
    cv.wait(l, [&] { return flags != 0; });
    return true;
  }
  return cv.wait_for(l, std::chrono::duration<double>(seconds), [&] { return flags != 0; });
}

}  // namespace mbh
