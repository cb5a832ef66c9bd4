// Minimal control plane for the host layer: name-addressed message passing (Rpc), group membership (Broker /
// GroupService) and small-value collectives.  It carries membership epochs, counters, leader election, CUDA-IPC
// handles and late-joiner model state -- never gradients of CUDA models, which move over NVLink in the K-A2 kernel.
//
// This is deliberately NOT a port of the reference's RPC stack (src/rpc.cc, src/transports/*: 5k lines, out of scope
// per SURVEY.md section 2 rows 8-12).  It keeps the reference's protocol SHAPE so that the Python API behaves the same:
//   ping / resync / sync / update with syncId epochs          (src/broker.h:99-237, src/group.h:330-491)
//   allreduce ops keyed "<syncId>.<group>::<name>", early arrivals parked, cancel on regroup, timeout
//                                                              (src/group.h:508-788)
// Topology: the Rpc that listen()s is a hub; every other Rpc connect()s to it and all traffic is relayed (one box).
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <set>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace mbh {

using Clock = std::chrono::steady_clock;
using Bytes = std::string;

// ---- tiny serializer -----------------------------------------------------------------------------------------------
struct Writer {
  Bytes b;
  void u32(uint32_t v) { b.append(reinterpret_cast<const char*>(&v), 4); }
  void u64(uint64_t v) { b.append(reinterpret_cast<const char*>(&v), 8); }
  void i32(int32_t v) { b.append(reinterpret_cast<const char*>(&v), 4); }
  void i64(int64_t v) { b.append(reinterpret_cast<const char*>(&v), 8); }
  void str(const std::string& s) {
    u32((uint32_t)s.size());
    b.append(s);
  }
};
struct Reader {
  const char* p;
  const char* e;
  explicit Reader(const Bytes& b) : p(b.data()), e(b.data() + b.size()) {}
  Reader(const char* p, size_t n) : p(p), e(p + n) {}
  void need(size_t n) {
    if ((size_t)(e - p) < n) throw std::runtime_error("moolib_b200 control plane: truncated message");
  }
  uint32_t u32() {
    need(4);
    uint32_t v;
    std::memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  uint64_t u64() {
    need(8);
    uint64_t v;
    std::memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  int32_t i32() { return (int32_t)u32(); }
  int64_t i64() { return (int64_t)u64(); }
  std::string str() {
    uint32_t n = u32();
    need(n);
    std::string s(p, n);
    p += n;
    return s;
  }
  bool done() const { return p == e; }
};

// ---- Rpc: name-addressed datagrams over TCP, relayed through the listening hub --------------------------------------
class RpcCore : public std::enable_shared_from_this<RpcCore> {
 public:
  using Handler = std::function<void(const std::string& src, const Bytes& payload)>;

  RpcCore();
  ~RpcCore();

  void setName(const std::string& name);
  std::string getName();
  void listen(const std::string& address);
  void connect(const std::string& address);
  void setTimeout(double seconds) { timeoutSeconds_ = seconds; }
  double getTimeout() const { return timeoutSeconds_; }

  // Register the handler of a service; called on the IO thread.
  void handle(const std::string& service, Handler h);
  void unhandle(const std::string& service);
  // Fire-and-forget (the hub parks messages for names it has not seen yet, up to the timeout).
  void send(const std::string& dst, const std::string& service, const Bytes& payload);
  bool isHub() const { return listening_; }
  std::string debugInfo();
  void close();

 private:
  struct Conn;
  void ioLoop();
  void ensureThread();
  void onFrame(Conn& c, const char* data, size_t len);
  void deliverLocal(const std::string& src, const std::string& service, const Bytes& payload);
  void route(const std::string& dst, const std::string& src, const std::string& service, const Bytes& payload);
  bool writeFrame(Conn& c, const Bytes& frame);
  void wake();

  std::mutex mu_;
  std::string name_;
  bool nameSet_ = false;
  std::atomic<bool> listening_{false};
  std::atomic<bool> stop_{false};
  double timeoutSeconds_ = 10.0;
  int listenFd_ = -1;
  int wakeFd_[2] = {-1, -1};
  std::vector<std::shared_ptr<Conn>> conns_;
  std::shared_ptr<Conn> hub_;  // client side: connection to the hub
  std::unordered_map<std::string, std::shared_ptr<Conn>> byName_;  // hub side
  struct Parked {
    Clock::time_point t;
    std::string dst, src, service;
    Bytes payload;
  };
  std::deque<Parked> parked_;
  std::deque<Parked> outbox_;  // client side: sends issued before connect() completed
  // Handlers run under a shared lock; handle()/unhandle() take it exclusively, so once unhandle() returns no call of
  // that handler is still executing (services unregister in their destructors).
  std::shared_mutex hmu_;
  std::unordered_map<std::string, Handler> handlers_;
  std::thread thread_;
  bool threadStarted_ = false;
  uint64_t sent_ = 0, received_ = 0;
};

// ---- Future ------------------------------------------------------------------------------------------------------
// Completed from the IO thread, awaited from Python (GIL released while waiting).  `value` is opaque bytes; the
// binding layer converts (e.g. unpickles) on the Python thread.
struct FutureState {
  std::mutex mu;
  std::condition_variable cv;
  int flags = 0;  // 1 result, 2 error, 4 cancelled
  Bytes value;
  std::string error;
  std::function<void()> onDone;  // optional, run once outside the lock
  void setResult(Bytes v);
  void setError(std::string e);
  void cancel();
  bool done();
  bool wait(double seconds);  // < 0: forever
  // Snapshot (flags, value, error) without keeping the lock: callers must not hold `mu` while they start new operations
  // (GroupService::allReduce inspects the previous op of the same name, i.e. possibly this very future).
  int snapshot(Bytes* v, std::string* e) {
    std::lock_guard<std::mutex> l(mu);
    if (v && (flags & 1)) *v = value;
    if (e) *e = error;
    return flags;
  }
};

// ---- group membership -------------------------------------------------------------------------------------------
struct GroupInfo {
  std::mutex mutex;
  std::string name;
  std::string brokerName = "broker";
  int32_t sortOrder = 0;
  std::atomic<uint32_t> syncId{0};
  std::vector<std::string> members;
  // resync state (src/group.h:167-193)
  bool isResyncing = false, haveUpdate = false, hasPinged = false;
  std::atomic<bool> wantsResync{false};
  uint32_t newSyncId = 0;
  std::vector<std::string> newMembers;
  Clock::time_point lastPing{}, lastPingResponse{};
  bool pingOutstanding = false;
  bool brokerConnectionIsActive = false;
  std::optional<uint32_t> pingResponse;
  std::vector<std::weak_ptr<struct SmallReduce>> activeAllReductions;
};

// One small-value allreduce in flight (star through member 0, reduced in member order).
struct SmallReduce {
  std::string opName;  // "<syncId hex>.<group>::<name>"
  uint32_t syncId = 0;
  std::vector<std::string> peers;
  size_t myIndex = 0;
  Clock::time_point timestamp;
  std::function<Bytes(const Bytes& a, const Bytes& b)> op;  // runs at member 0 only: a = lower index, b = higher
  std::shared_ptr<FutureState> future;
  // member 0 state
  std::mutex mu;
  std::map<size_t, Bytes> got;
  bool finished = false;
};

class GroupService {
 public:
  explicit GroupService(std::shared_ptr<RpcCore> rpc);
  ~GroupService();
  std::shared_ptr<GroupInfo> group(const std::string& name);
  bool update(GroupInfo& g, int32_t sortOrder, uint32_t timeoutMs);
  void resync(GroupInfo& g);

  // Start a small allreduce over the current members; the returned future completes with the reduced bytes.
  std::shared_ptr<SmallReduce> allReduce(std::shared_ptr<GroupInfo> g, const std::string& name, Bytes value,
                                         std::function<Bytes(const Bytes&, const Bytes&)> op);
  std::shared_ptr<RpcCore> rpc() { return rpc_; }

 private:
  void onContribution(const std::string& src, const Bytes& payload);
  void onResult(const std::string& src, const Bytes& payload);
  void tryFinish(const std::shared_ptr<SmallReduce>& r);
  bool feed(const std::string& opName, uint32_t syncId, size_t index, const Bytes& value);
  void feedOp(const std::shared_ptr<SmallReduce>& r, size_t index, const Bytes& value);
  std::shared_ptr<SmallReduce> liveOpLocked(const std::string& opName, uint32_t syncId);

  std::shared_ptr<RpcCore> rpc_;
  std::mutex mu_;
  std::unordered_map<std::string, std::shared_ptr<GroupInfo>> groups_;
  std::unordered_map<std::string, std::weak_ptr<SmallReduce>> ops_;
  struct Early {
    Clock::time_point t;
    std::string opName;
    uint32_t syncId;
    size_t index;
    Bytes value;
  };
  std::vector<Early> early_;
};

class BrokerService {
 public:
  explicit BrokerService(std::shared_ptr<RpcCore> rpc);
  ~BrokerService();
  void update();

 private:
  struct Peer {
    std::string name;
    Clock::time_point lastPing;
    std::chrono::milliseconds timeout{10000};
    bool active = false;
    uint64_t creationOrder = 0;
    int32_t sortOrder = 0;
    bool syncReplied = false;
  };
  struct Grp {
    std::string name;
    std::map<std::string, Peer> peers;
    uint32_t syncId = 0;
    bool needsUpdate = false;
    bool syncing = false;
    Clock::time_point lastUpdate{};
    std::vector<std::string> active;
  };
  std::shared_ptr<RpcCore> rpc_;
  std::mutex mu_;
  std::map<std::string, Grp> groups_;
  uint64_t creationCounter_ = 0;
  uint32_t nextSyncId_;
  Clock::time_point lastCheck_{};
};

// Per-Rpc singletons (the reference's rpc->getService<T>(), src/rpc.h:146-193)
std::shared_ptr<GroupService> groupServiceFor(const std::shared_ptr<RpcCore>& rpc);

std::string randomName();

}  // namespace mbh
