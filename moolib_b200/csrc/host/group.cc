// Membership + small-value collectives of the control plane (see control.h) and the Python classes Rpc / Broker /
// Group / AllReduce (reference bindings: src/moolib.cc:1984-2164, 2210-2284).
#include "common.h"
#include "control.h"
#include "device_reduce.h"

#include <algorithm>
#include <random>
#include <sstream>

namespace mbh {

namespace {

std::string hex32(uint32_t v) {
  char buf[16];
  snprintf(buf, sizeof(buf), "%#x", v);
  return buf;
}

constexpr auto kBrokerCheckInterval = std::chrono::milliseconds(100);  // reference: 500 ms (src/broker.h:190)
constexpr auto kBrokerMinUpdateGap = std::chrono::milliseconds(500);   // reference: 2 s   (src/broker.h:216)

}  // namespace

// =====================================================================================================================
// GroupService (client side)                                                      reference: src/group.h:330-491
// =====================================================================================================================

GroupService::GroupService(std::shared_ptr<RpcCore> rpc) : rpc_(std::move(rpc)) {
  rpc_->handle("Group::pong", [this](const std::string&, const Bytes& p) {
    Reader r(p);
    std::string gname = r.str();
    uint32_t syncId = r.u32();
    auto g = group(gname);
    std::lock_guard<std::mutex> l(g->mutex);
    g->pingResponse = syncId;
    g->pingOutstanding = false;
  });
  rpc_->handle("Group::sync", [this](const std::string& src, const Bytes& p) {
    Reader r(p);
    std::string gname = r.str();
    uint32_t syncId = r.u32();
    std::shared_ptr<GroupInfo> g;
    {
      std::lock_guard<std::mutex> l(mu_);
      auto i = groups_.find(gname);
      if (i != groups_.end()) g = i->second;
    }
    Writer w;
    w.str(gname);
    w.str(rpc_->getName());
    if (g) {
      std::lock_guard<std::mutex> l(g->mutex);
      g->isResyncing = true;
      g->haveUpdate = false;
      g->newSyncId = syncId;
      w.u32(syncId);
      w.i32(g->sortOrder);
    } else {
      w.u32(0xffffffffu);
      w.i32(-1);
    }
    rpc_->send(src, "Broker::syncReply", w.b);
  });
  rpc_->handle("Group::update", [this](const std::string&, const Bytes& p) {
    Reader r(p);
    std::string gname = r.str();
    uint32_t syncId = r.u32();
    uint32_t n = r.u32();
    std::vector<std::string> members(n);
    for (auto& m : members) m = r.str();
    std::shared_ptr<GroupInfo> g;
    {
      std::lock_guard<std::mutex> l(mu_);
      auto i = groups_.find(gname);
      if (i != groups_.end()) g = i->second;
    }
    if (!g) return;
    std::lock_guard<std::mutex> l(g->mutex);
    if (syncId == g->newSyncId) {
      g->newMembers = std::move(members);
      g->haveUpdate = true;
    }
  });
  rpc_->handle("AR::contrib", [this](const std::string& src, const Bytes& p) { onContribution(src, p); });
  rpc_->handle("AR::result", [this](const std::string& src, const Bytes& p) { onResult(src, p); });
}

GroupService::~GroupService() {
  unhandleAll(*rpc_, {"Group::pong", "Group::sync", "Group::update", "AR::contrib", "AR::result"});
}

std::shared_ptr<GroupInfo> GroupService::group(const std::string& name) {
  std::lock_guard<std::mutex> l(mu_);
  auto& g = groups_[name];
  if (!g) {
    g = std::make_shared<GroupInfo>();
    g->name = name;
  }
  return g;
}

void GroupService::resync(GroupInfo& g) {
  if (g.wantsResync.load() || g.isResyncing) return;
  g.wantsResync = true;
  Writer w;
  w.str(g.name);
  rpc_->send(g.brokerName, "Broker::resync", w.b);
}

// Mirrors GroupService::update (src/group.h:393-490): ping cadence, broker-silence detection, adoption of a pushed
// member list, cancellation / timeout of in-flight allreduces.
bool GroupService::update(GroupInfo& g, int32_t sortOrder, uint32_t timeoutMs) {
  MBH_PHASE("GroupService::update");
  auto now = Clock::now();
  std::vector<std::shared_ptr<SmallReduce>> cancelled, timedOut;
  bool updated;
  {
    std::lock_guard<std::mutex> l(g.mutex);
    auto pingNow = [&] {
      g.lastPing = now;
      g.hasPinged = true;
      g.pingOutstanding = true;
      g.pingResponse.reset();
      Writer w;
      w.str(g.name);
      w.str(rpc_->getName());
      w.u32(timeoutMs);
      rpc_->send(g.brokerName, "Broker::ping", w.b);
    };
    g.sortOrder = sortOrder;
    updated = g.isResyncing && g.haveUpdate;
    if (updated) {
      g.wantsResync = false;
      g.isResyncing = false;
      g.haveUpdate = false;
      g.syncId = g.newSyncId;
      g.members = std::move(g.newMembers);
      pingNow();
    }
    auto interval = std::min(std::chrono::milliseconds(4000), std::chrono::milliseconds(timeoutMs) / 2);
    if (!g.hasPinged || updated || now - g.lastPing >= interval) {
      if (g.hasPinged && g.pingOutstanding) {
        if (g.brokerConnectionIsActive &&
            now - g.lastPingResponse >= std::chrono::seconds(1) + std::chrono::milliseconds(timeoutMs)) {
          g.brokerConnectionIsActive = false;
          g.members.clear();
          g.syncId = 0;
          resync(g);
          updated = true;
        }
      } else {
        g.brokerConnectionIsActive = true;
        g.lastPingResponse = now;
      }
      if (g.pingResponse && *g.pingResponse != g.syncId.load()) {
        g.members.clear();
        g.syncId = 0;
        resync(g);
        updated = true;
      }
      pingNow();
    }
    if (updated) {
      for (auto& wh : g.activeAllReductions)
        if (auto h = wh.lock()) cancelled.push_back(h);
      g.activeAllReductions.clear();
    } else {
      auto timeout = std::chrono::duration<double>(rpc_->getTimeout());
      bool shouldResync = false;
      auto& v = g.activeAllReductions;
      v.erase(std::remove_if(v.begin(), v.end(),
                             [&](std::weak_ptr<SmallReduce>& wh) {
                               auto h = wh.lock();
                               if (!h || h->future->done()) return true;
                               if (now >= h->timestamp + timeout) {
                                 shouldResync = true;
                                 timedOut.push_back(h);
                                 return true;
                               }
                               return false;
                             }),
              v.end());
      if (shouldResync) resync(g);
    }
  }
  for (auto& h : cancelled) h->future->setError("AllReduce operation cancelled due to a group change");
  for (auto& h : timedOut) h->future->setError("AllReduce operation timed out");
  MBH_PHASE("idle");
  return updated;
}

// ---- small allreduce (star through member 0) --------------------------------------------------------------------

std::shared_ptr<SmallReduce> GroupService::allReduce(std::shared_ptr<GroupInfo> g, const std::string& name, Bytes value,
                                                     std::function<Bytes(const Bytes&, const Bytes&)> op) {
  MBH_PHASE("GroupService::allReduce");
  auto r = std::make_shared<SmallReduce>();
  r->future = std::make_shared<FutureState>();
  r->op = std::move(op);
  r->timestamp = Clock::now();
  bool start;
  {
    std::lock_guard<std::mutex> l(g->mutex);
    r->syncId = g->syncId;
    r->opName = hex32(r->syncId) + "." + g->name + "::" + name;
    r->peers = g->members;
    auto i = std::find(r->peers.begin(), r->peers.end(), rpc_->getName());
    if (i == r->peers.end()) throw std::runtime_error("AllReduce: local peer is not a member of the specified group!");
    r->myIndex = (size_t)(i - r->peers.begin());
    g->activeAllReductions.push_back(r);
    start = !(g->wantsResync.load() || g->isResyncing);
  }
  {
    std::lock_guard<std::mutex> l(mu_);
    auto& slot = ops_[r->opName];
    if (auto prev = slot.lock()) {
      if (!prev->future->done())
        throw std::runtime_error("Attempt to all-reduce twice concurrently with the name '" + name + "'");
    }
    slot = r;
  }
  if (!start) return r;  // will be cancelled by the pending group change (src/group.h:736)
  if (r->peers.size() == 1) {
    r->future->setResult(std::move(value));
    return r;
  }
  if (r->myIndex == 0) {
    feed(r->opName, r->syncId, 0, value);
    // contributions that arrived before this op existed (src/group.h:771-783)
    std::vector<Early> mine;
    {
      std::lock_guard<std::mutex> l(mu_);
      auto now = Clock::now();
      auto ttl = std::chrono::duration<double>(rpc_->getTimeout());
      for (auto i = early_.begin(); i != early_.end();) {
        if (i->opName == r->opName && i->syncId == r->syncId) {
          mine.push_back(std::move(*i));
          i = early_.erase(i);
        } else if (now - i->t > ttl) {
          i = early_.erase(i);
        } else {
          ++i;
        }
      }
    }
    for (auto& e : mine) feed(e.opName, e.syncId, e.index, e.value);
  } else {
    Writer w;
    w.str(r->opName);
    w.u32(r->syncId);
    w.u64(r->myIndex);
    w.str(value);
    rpc_->send(r->peers[0], "AR::contrib", w.b);
  }
  return r;
}

// Look the operation up; returns it only if a contribution for (opName, syncId) can be fed into it right now.
// Caller holds mu_.
std::shared_ptr<SmallReduce> GroupService::liveOpLocked(const std::string& opName, uint32_t syncId) {
  auto i = ops_.find(opName);
  if (i == ops_.end()) return nullptr;
  std::shared_ptr<SmallReduce> r = i->second.lock();
  if (!r || r->syncId != syncId || r->future->done() || r->myIndex != 0) return nullptr;
  return r;
}

bool GroupService::feed(const std::string& opName, uint32_t syncId, size_t index, const Bytes& value) {
  std::shared_ptr<SmallReduce> r;
  {
    std::lock_guard<std::mutex> l(mu_);
    r = liveOpLocked(opName, syncId);
  }
  if (!r) return false;
  feedOp(r, index, value);
  return true;
}

void GroupService::feedOp(const std::shared_ptr<SmallReduce>& r, size_t index, const Bytes& value) {
  {
    std::lock_guard<std::mutex> l(r->mu);
    if (r->finished || index >= r->peers.size() || r->got.count(index)) return;
    r->got[index] = value;
  }
  tryFinish(r);
}

void GroupService::tryFinish(const std::shared_ptr<SmallReduce>& r) {
  // Take the contributions out under the lock, reduce WITHOUT it: a Python `op` needs the GIL, and the thread that
  // holds the GIL may be inside allReduce() waiting for r->mu (feed -> tryFinish) -- reducing under r->mu deadlocked.
  std::map<size_t, Bytes> got;
  {
    std::lock_guard<std::mutex> l(r->mu);
    if (r->finished || r->got.size() != r->peers.size()) return;
    r->finished = true;
    got.swap(r->got);
  }
  // fixed member order: ((v0 op v1) op v2) ...
  Bytes acc = std::move(got[0]);
  for (size_t i = 1; i < r->peers.size(); ++i) acc = r->op(acc, got[i]);
  Writer w;
  w.str(r->opName);
  w.u32(r->syncId);
  w.str(acc);
  for (size_t i = 1; i < r->peers.size(); ++i) rpc_->send(r->peers[i], "AR::result", w.b);
  r->future->setResult(std::move(acc));
}

void GroupService::onContribution(const std::string&, const Bytes& p) {
  Reader rd(p);
  std::string opName = rd.str();
  uint32_t syncId = rd.u32();
  size_t index = (size_t)rd.u64();
  Bytes value = rd.str();
  std::shared_ptr<SmallReduce> r;
  {
    // ONE critical section for "is the operation there?" and "park the contribution": allReduce() registers the op and
    // scans early_ under the same mutex, so a contribution is either fed or found by that scan -- never stranded.
    std::lock_guard<std::mutex> l(mu_);
    r = liveOpLocked(opName, syncId);
    if (!r) {
      early_.push_back(Early{Clock::now(), std::move(opName), syncId, index, std::move(value)});
      return;
    }
  }
  feedOp(r, index, value);
}

void GroupService::onResult(const std::string&, const Bytes& p) {
  Reader rd(p);
  std::string opName = rd.str();
  uint32_t syncId = rd.u32();
  Bytes value = rd.str();
  std::shared_ptr<SmallReduce> r;
  {
    std::lock_guard<std::mutex> l(mu_);
    auto i = ops_.find(opName);
    if (i != ops_.end()) r = i->second.lock();
  }
  if (r && r->syncId == syncId) r->future->setResult(std::move(value));
}

// =====================================================================================================================
// BrokerService                                                                   reference: src/broker.h:31-237
// =====================================================================================================================

BrokerService::BrokerService(std::shared_ptr<RpcCore> rpc) : rpc_(std::move(rpc)) {
  std::random_device rd;
  nextSyncId_ = rd() | 1u;
  rpc_->handle("Broker::ping", [this](const std::string& src, const Bytes& p) {
    Reader r(p);
    std::string gname = r.str(), name = r.str();
    uint32_t timeoutMs = r.u32();
    uint32_t syncId;
    {
      std::lock_guard<std::mutex> l(mu_);
      auto& g = groups_[gname];
      g.name = gname;
      auto it = g.peers.find(name);
      if (it == g.peers.end()) {
        it = g.peers.emplace(name, Peer{}).first;
        it->second.name = name;
        it->second.creationOrder = creationCounter_++;
      }
      auto& peer = it->second;
      peer.lastPing = Clock::now();
      peer.timeout = std::chrono::milliseconds(timeoutMs);
      if (!peer.active) g.needsUpdate = true;
      syncId = g.syncId;
    }
    Writer w;
    w.str(gname);
    w.u32(syncId);
    rpc_->send(src, "Group::pong", w.b);
  });
  rpc_->handle("Broker::resync", [this](const std::string&, const Bytes& p) {
    Reader r(p);
    std::string gname = r.str();
    std::lock_guard<std::mutex> l(mu_);
    groups_[gname].name = gname;
    groups_[gname].needsUpdate = true;
  });
  rpc_->handle("Broker::syncReply", [this](const std::string&, const Bytes& p) {
    Reader r(p);
    std::string gname = r.str(), name = r.str();
    uint32_t syncId = r.u32();
    int32_t sortOrder = r.i32();
    std::lock_guard<std::mutex> l(mu_);
    auto gi = groups_.find(gname);
    if (gi == groups_.end()) return;
    auto pi = gi->second.peers.find(name);
    if (pi == gi->second.peers.end() || syncId != gi->second.syncId) return;
    pi->second.syncReplied = true;
    pi->second.sortOrder = sortOrder;
  });
}

BrokerService::~BrokerService() { unhandleAll(*rpc_, {"Broker::ping", "Broker::resync", "Broker::syncReply"}); }

void BrokerService::update() {
  MBH_PHASE("BrokerService::update");
  auto now = Clock::now();
  struct Push {
    std::string peer, service;
    Bytes payload;
  };
  std::vector<Push> out;
  {
    std::lock_guard<std::mutex> l(mu_);
    // finish syncs: every peer answered, or 1 s passed (src/broker.h:150-186)
    for (auto& [gname, g] : groups_) {
      if (!g.syncing) continue;
      size_t total = g.peers.size(), ready = 0;
      for (auto& [pn, p] : g.peers) ready += p.syncReplied;
      if (ready >= total || now - g.lastUpdate >= std::chrono::seconds(1)) {
        std::vector<Peer*> act;
        for (auto& [pn, p] : g.peers) {
          p.active = p.syncReplied;
          if (p.active) act.push_back(&p);
        }
        std::sort(act.begin(), act.end(), [](Peer* a, Peer* b) {
          if (a->sortOrder == b->sortOrder) return a->creationOrder < b->creationOrder;
          return a->sortOrder < b->sortOrder;
        });
        g.active.clear();
        for (auto* p : act) g.active.push_back(p->name);
        Writer w;
        w.str(gname);
        w.u32(g.syncId);
        w.u32((uint32_t)g.active.size());
        for (auto& n : g.active) w.str(n);
        for (auto* p : act) out.push_back({p->name, "Group::update", w.b});
        g.syncing = false;
      }
    }
    if (now - lastCheck_ >= kBrokerCheckInterval) {
      lastCheck_ = now;
      for (auto& [gname, g] : groups_) {
        for (auto i = g.peers.begin(); i != g.peers.end();) {
          if (now - i->second.lastPing >= i->second.timeout) {
            if (i->second.active) g.needsUpdate = true;
            i = g.peers.erase(i);
          } else {
            ++i;
          }
        }
        if (g.needsUpdate && !g.syncing && now - g.lastUpdate >= kBrokerMinUpdateGap) {
          g.lastUpdate = now;
          g.needsUpdate = false;
          uint32_t syncId = nextSyncId_++;
          if (syncId == 0) syncId = nextSyncId_++;
          g.syncId = syncId;
          Writer w;
          w.str(gname);
          w.u32(syncId);
          for (auto& [pn, p] : g.peers) {
            p.syncReplied = false;
            out.push_back({pn, "Group::sync", w.b});
          }
          g.syncing = true;
        }
      }
    }
  }
  for (auto& m : out) rpc_->send(m.peer, m.service, m.payload);
}

// ---- per-Rpc service registry ----------------------------------------------------------------------------------
namespace {
std::mutex g_regMu;
std::map<RpcCore*, std::weak_ptr<GroupService>> g_groupServices;
}  // namespace

std::shared_ptr<GroupService> groupServiceFor(const std::shared_ptr<RpcCore>& rpc) {
  std::lock_guard<std::mutex> l(g_regMu);
  auto& w = g_groupServices[rpc.get()];
  auto s = w.lock();
  if (!s) {
    s = std::make_shared<GroupService>(rpc);
    w = s;
  }
  return s;
}

// =====================================================================================================================
// Python classes
// =====================================================================================================================

py::object pickleLoads(const Bytes& b) {
  static py::object loads = py::module_::import("pickle").attr("loads");
  return loads(py::bytes(b));
}
Bytes pickleDumps(const py::handle& o) {
  static py::object dumps = py::module_::import("pickle").attr("dumps");
  return dumps(o).cast<std::string>();
}

Bytes packTensor(const torch::Tensor& t0) {
  torch::Tensor t = t0.contiguous();
  Writer w;
  w.i32((int32_t)t.scalar_type());
  w.u32((uint32_t)t.dim());
  for (auto s : t.sizes()) w.i64(s);
  w.str(std::string(static_cast<const char*>(t.data_ptr()), t.nbytes()));
  return w.b;
}
torch::Tensor unpackTensor(const Bytes& b) {
  Reader r(b);
  auto dt = (c10::ScalarType)r.i32();
  uint32_t nd = r.u32();
  std::vector<int64_t> sizes(nd);
  for (auto& s : sizes) s = r.i64();
  std::string raw = r.str();
  torch::Tensor t = torch::empty(sizes, torch::TensorOptions().dtype(dt));
  if (t.nbytes() != raw.size()) throw std::runtime_error("tensor payload size mismatch");
  std::memcpy(t.data_ptr(), raw.data(), raw.size());
  return t;
}

// reference: QueueWrapper (src/moolib.cc:433-576) -- calls to a function defined with define_queue() are parked here and
// handed to the user as (return_callback, args, kwargs); with dynamic batching up to batch_size parked calls are merged
// with utils::stackFields (src/moolib.cc:492) and the single result is split again with unstackFields (:499).
struct PyQueue {
  struct Entry {
    std::string src;
    uint64_t id;
    py::object args, kwargs;
    Clock::time_point t;
  };
  std::weak_ptr<RpcCore> core;
  int64_t batchSize = 0;
  bool dynamicBatching = false;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Entry> q;

  ~PyQueue() {
    py::gil_scoped_acquire gil;
    q.clear();
  }

  static void reply(const std::weak_ptr<RpcCore>& wc, const std::string& src, uint64_t id, const py::handle& result) {
    auto c = wc.lock();
    if (!c) return;
    Writer w;
    w.u64(id);
    try {
      Bytes body = pickleDumps(result);
      w.u32(0);
      w.str(body);
    } catch (const std::exception& e) {
      w.u32(1);
      w.str(std::string("could not serialise the result: ") + e.what());
    }
    c->send(src, "Rpc::reply", w.b);
  }

  void push(Entry e) {
    {
      std::lock_guard<std::mutex> l(mu);
      q.push_back(std::move(e));
    }
    cv.notify_one();
  }
  size_t size() {
    std::lock_guard<std::mutex> l(mu);
    return q.size();
  }

  // Non-blocking: None when nothing is parked.
  py::object tryGet() {
    std::vector<Entry> batch;
    {
      std::lock_guard<std::mutex> l(mu);
      if (q.empty()) return py::none();
      size_t n = batchSize > 0 ? std::min<size_t>((size_t)batchSize, dynamicBatching ? q.size() : (size_t)batchSize) : 1;
      if (batchSize > 0 && !dynamicBatching && q.size() < n) return py::none();  // static batching waits for a full batch
      for (size_t i = 0; i < n; ++i) {
        batch.push_back(std::move(q.front()));
        q.pop_front();
      }
    }
    auto wc = core;
    if (batchSize <= 0) {
      Entry& e = batch[0];
      std::string src = e.src;
      uint64_t id = e.id;
      py::cpp_function ret([wc, src, id](py::object result) { reply(wc, src, id, result); });
      return py::make_tuple(ret, e.args, e.kwargs);
    }
    const int64_t n = (int64_t)batch.size();
    py::tuple srcs(n);
    std::vector<std::pair<std::string, uint64_t>> dests;
    for (int64_t i = 0; i < n; ++i) {
      srcs[i] = py::make_tuple(batch[i].args, batch[i].kwargs);
      dests.emplace_back(batch[i].src, batch[i].id);
    }
    py::tuple stacked = py::reinterpret_borrow<py::tuple>(stackFields(srcs, 0));
    py::cpp_function ret([wc, dests, n](py::object result) {
      if (result.is_none()) {
        for (auto& d : dests) reply(wc, d.first, d.second, py::none());
        return;
      }
      py::tuple parts = unstackFields(result, n, 0);
      for (int64_t i = 0; i < n; ++i) reply(wc, dests[i].first, dests[i].second, parts[i]);
    });
    return py::make_tuple(ret, stacked[0], stacked[1]);
  }

  py::object get(std::optional<double> timeout) {
    auto deadline = Clock::now() + std::chrono::duration<double>(timeout ? *timeout : 1e9);
    while (true) {
      py::object r = tryGet();
      if (!r.is_none()) return r;
      if (Clock::now() >= deadline) throw std::runtime_error("Queue.get timed out");
      {
        py::gil_scoped_release nogil;
        std::unique_lock<std::mutex> l(mu);
        cv.wait_for(l, std::chrono::milliseconds(20), [&] { return !q.empty(); });
      }
      if (PyErr_CheckSignals() != 0) throw py::error_already_set();
    }
  }
};

struct PyRpc {
  std::shared_ptr<RpcCore> core = std::make_shared<RpcCore>();
  std::map<std::string, std::shared_ptr<PyQueue>> queues;
  std::mutex mu;
  std::map<std::string, py::object> functions;
  std::map<uint64_t, std::shared_ptr<FutureState>> calls;
  uint64_t nextCall = 1;
  bool servicesUp = false;

  ~PyRpc() {
    unhandleAll(*core, {"Rpc::call", "Rpc::reply"});
    {
      py::gil_scoped_release nogil;
      core->close();
    }
    functions.clear();
    queues.clear();
  }

  void setupServices() {
    if (servicesUp) return;
    servicesUp = true;
    // user-defined functions (reference: Rpc::define / async_, src/moolib.cc:1024-1290): pickled args, executed on
    // the IO thread under the GIL, reply routed back to the caller
    core->handle("Rpc::call", [this](const std::string& src, const Bytes& p) {
      Reader r(p);
      uint64_t id = r.u64();
      std::string fname = r.str();
      Bytes args = r.str();
      Writer w;
      w.u64(id);
      {
        py::gil_scoped_acquire gil;
        py::object fn;
        std::shared_ptr<PyQueue> queue;
        {
          std::lock_guard<std::mutex> l(mu);
          auto i = functions.find(fname);
          if (i != functions.end()) fn = i->second;
          auto qi = queues.find(fname);
          if (qi != queues.end()) queue = qi->second;
        }
        if (queue) {
          try {
            py::tuple ak = pickleLoads(args);
            queue->push(PyQueue::Entry{src, id, ak[0], ak[1], Clock::now()});
            return;  // the reply is sent when the user calls the return callback
          } catch (const std::exception& e) {
            w.u32(1);
            w.str(e.what());
          }
        } else if (!fn) {
          w.u32(1);
          w.str("RPC function '" + fname + "' does not exist on peer '" + core->getName() + "'");
        } else {
          try {
            py::tuple ak = pickleLoads(args);
            py::object res = fn(*py::reinterpret_borrow<py::tuple>(ak[0]), **py::reinterpret_borrow<py::dict>(ak[1]));
            w.u32(0);
            w.str(pickleDumps(res));
          } catch (py::error_already_set& e) {
            w.u32(1);
            w.str(std::string("Python exception in remote function '") + fname + "': " + e.what());
          } catch (const std::exception& e) {
            w.u32(1);
            w.str(e.what());
          }
        }
      }
      core->send(src, "Rpc::reply", w.b);
    });
    core->handle("Rpc::reply", [this](const std::string&, const Bytes& p) {
      Reader r(p);
      uint64_t id = r.u64();
      uint32_t err = r.u32();
      Bytes body = r.str();
      std::shared_ptr<FutureState> f;
      {
        std::lock_guard<std::mutex> l(mu);
        auto i = calls.find(id);
        if (i == calls.end()) return;
        f = i->second;
        calls.erase(i);
      }
      if (err) f->setError(body);
      else f->setResult(std::move(body));
    });
  }

  void define(const std::string& name, py::object fn) {
    setupServices();
    std::lock_guard<std::mutex> l(mu);
    functions[name] = std::move(fn);
  }
  void undefine(const std::string& name) {
    std::lock_guard<std::mutex> l(mu);
    functions.erase(name);
    queues.erase(name);
  }
  std::shared_ptr<PyQueue> defineQueue(const std::string& name, py::kwargs kwargs) {
    setupServices();
    auto q = std::make_shared<PyQueue>();
    q->core = core;
    if (kwargs.contains("batch_size") && !kwargs["batch_size"].is_none()) q->batchSize = kwargs["batch_size"].cast<int64_t>();
    if (kwargs.contains("dynamic_batching")) q->dynamicBatching = kwargs["dynamic_batching"].cast<bool>();
    std::lock_guard<std::mutex> l(mu);
    queues[name] = q;
    return q;
  }
  std::shared_ptr<PyFuture> asyncCall(const std::string& peer, const std::string& fname, py::args args,
                                      py::kwargs kwargs) {
    setupServices();
    auto fut = std::make_shared<PyFuture>();
    fut->state = std::make_shared<FutureState>();
    fut->decode = [](const Bytes& b) { return pickleLoads(b); };
    uint64_t id;
    {
      std::lock_guard<std::mutex> l(mu);
      id = nextCall++;
      calls[id] = fut->state;
    }
    Writer w;
    w.u64(id);
    w.str(fname);
    w.str(pickleDumps(py::make_tuple(args, kwargs)));
    core->send(peer, "Rpc::call", w.b);
    return fut;
  }
  py::object syncCall(const std::string& peer, const std::string& fname, py::args args, py::kwargs kwargs) {
    auto f = asyncCall(peer, fname, std::move(args), std::move(kwargs));
    double timeout = core->getTimeout();
    f->wait(timeout);
    if (!f->done()) throw std::runtime_error("Call (" + peer + "::" + fname + ") timed out");
    return f->get();
  }
};

struct PyBroker {
  std::shared_ptr<PyRpc> rpc;
  std::shared_ptr<BrokerService> service;
  explicit PyBroker(std::shared_ptr<PyRpc> r) : rpc(std::move(r)) {
    if (!rpc) {
      rpc = std::make_shared<PyRpc>();
      rpc->core->setName("broker");
    }
    service = std::make_shared<BrokerService>(rpc->core);
  }
  void update() {
    py::gil_scoped_release nogil;
    service->update();
  }
};

struct PyGroup {
  std::shared_ptr<PyRpc> rpc;
  std::shared_ptr<GroupService> service;
  std::shared_ptr<GroupInfo> info;
  std::string groupName;
  uint32_t timeoutMs = 10 * 1000;
  int32_t sortOrder = 0;
  std::shared_ptr<DeviceReducerSet> reducers;  // K-A2 contexts of this group (CUDA tensors)

  PyGroup(std::shared_ptr<PyRpc> r, std::string name) : rpc(std::move(r)), groupName(std::move(name)) {
    if (!rpc) throw std::runtime_error("Group: rpc is None");
    service = groupServiceFor(rpc->core);
    info = service->group(groupName);
    reducers = std::make_shared<DeviceReducerSet>(service, info);
  }
  bool update() {
    // device-side all_reduces advance here too, not only when their own future is polled: a loop that waits on one
    // future (all(f.done() for f in futures) stops at the first unfinished one) must not starve the operations the
    // peers are waiting for
    reducers->progressAll();
    py::gil_scoped_release nogil;
    return service->update(*info, sortOrder, timeoutMs);
  }
  std::vector<std::string> members() {
    std::lock_guard<std::mutex> l(info->mutex);
    return info->members;
  }
  bool active() {
    std::lock_guard<std::mutex> l(info->mutex);
    return !info->members.empty();
  }

  // reference: GroupWrapper::allReduce, src/moolib.cc:1305-1365
  std::shared_ptr<PyFuture> allReduce(const std::string& name, py::object data, py::kwargs kwargs) {
    py::object op = py::none();
    if (kwargs.contains("op")) op = kwargs["op"];
    auto fut = std::make_shared<PyFuture>();
    if (!op.is_none()) {
      // the callable may be released from the IO thread: its reference must be dropped under the GIL
      std::shared_ptr<py::object> pop(new py::object(op), [](py::object* o) {
        py::gil_scoped_acquire gil;
        delete o;
      });
      auto red = service->allReduce(info, name, pickleDumps(data), [pop](const Bytes& a, const Bytes& b) {
        py::gil_scoped_acquire gil;
        return pickleDumps((*pop)(pickleLoads(a), pickleLoads(b)));
      });
      fut->state = red->future;
      fut->keep = red;
      fut->decode = [](const Bytes& b) { return pickleLoads(b); };
      return fut;
    }
    if (!is_tensor(data)) {
      throw std::runtime_error(
          "all_reduce can only use the default operator on Tensor data. Please specify an operator function");
    }
    torch::Tensor t = to_tensor(data);
    if (t.is_cuda()) {
      // HP-A, A8: the payload moves over NVLink in the K-A2 kernel; only the "everyone is here" gate is a message
      return reducers->allReduceTensor(name, t, data);
    }
    auto red = service->allReduce(info, name, packTensor(t), [](const Bytes& a, const Bytes& b) {
      torch::Tensor x = unpackTensor(a);
      x += unpackTensor(b);  // ReduceSum, src/group.h:243-247
      return packTensor(x);
    });
    fut->state = red->future;
    fut->keep = red;
    // the reference reduces in place ("the result is available in self.tensor", test/test_reduce.py:56)
    fut->decode = [t, data](const Bytes& b) mutable {
      t.copy_(unpackTensor(b).view_as(t));
      return data;
    };
    return fut;
  }
};

void bind_rpc(py::module_& m) {
  py::class_<PyFuture, std::shared_ptr<PyFuture>>(m, "Future")
      .def("result", &PyFuture::result, py::arg("timeout") = py::none())
      .def("done", &PyFuture::done)
      .def("exception", &PyFuture::exception)
      .def("cancel", &PyFuture::cancel)
      .def("wait", [](PyFuture& f, std::optional<double> t) { f.wait(t ? *t : -1.0); }, py::arg("timeout") = py::none());
  m.attr("AllReduce") = m.attr("Future");

  py::class_<PyQueue, std::shared_ptr<PyQueue>>(m, "Queue")
      .def("get", &PyQueue::get, py::arg("timeout") = py::none())
      .def("try_get", &PyQueue::tryGet)
      .def("size", &PyQueue::size);

  py::class_<PyRpc, std::shared_ptr<PyRpc>>(m, "Rpc")
      .def(py::init<>())
      .def("set_name", [](PyRpc& r, const std::string& n) { r.core->setName(n); })
      .def("get_name", [](PyRpc& r) { return r.core->getName(); })
      .def("listen", [](PyRpc& r, const std::string& a) { r.core->listen(a); })
      .def("connect", [](PyRpc& r, const std::string& a) { r.core->connect(a); })
      .def("set_timeout", [](PyRpc& r, double s) { r.core->setTimeout(s); })
      .def("set_transports", [](PyRpc&, py::object) { /* one transport: loopback TCP through the hub */ })
      .def("debug_info", [](PyRpc& r) { py::print(r.core->debugInfo()); })
      .def("define", [](PyRpc& r, const std::string& n, py::object fn, py::kwargs) { r.define(n, std::move(fn)); })
      .def("undefine", &PyRpc::undefine)
      .def("define_queue", &PyRpc::defineQueue, py::arg("name"))
      .def("async_", &PyRpc::asyncCall)
      .def("sync", &PyRpc::syncCall);

  py::class_<PyBroker, std::shared_ptr<PyBroker>>(m, "Broker")
      .def(py::init<std::shared_ptr<PyRpc>>(), py::arg("rpc") = nullptr)
      .def("set_name", [](PyBroker& b, const std::string& n) { b.rpc->core->setName(n); })
      .def("listen", [](PyBroker& b, const std::string& a) { b.rpc->core->listen(a); })
      .def("update", &PyBroker::update);

  py::class_<PyGroup, std::shared_ptr<PyGroup>>(m, "Group")
      .def(py::init<std::shared_ptr<PyRpc>, std::string>(), py::arg("rpc"), py::arg("name"))
      .def("update", &PyGroup::update)
      .def("set_broker_name", [](PyGroup& g, const std::string& n) {
        std::lock_guard<std::mutex> l(g.info->mutex);
        g.info->brokerName = n;
      })
      .def("set_timeout", [](PyGroup& g, double s) { g.timeoutMs = (uint32_t)(s * 1000); })
      .def("set_sort_order", [](PyGroup& g, int32_t o) { g.sortOrder = o; })
      .def("members", &PyGroup::members)
      .def("sync_id", [](PyGroup& g) { return g.info->syncId.load(); })
      .def("name", [](PyGroup& g) { return g.groupName; })
      .def("active", &PyGroup::active)
      .def("all_reduce", &PyGroup::allReduce, py::arg("name"), py::arg("data"));

  m.def("create_uid", [] { return randomName(); });
  m.def("set_log_level", [](py::object) {});
  m.def("set_logging", [](py::object) {});
  m.def("set_max_threads", [](int) {});
}

// accessors used by accumulator.cc
std::shared_ptr<RpcCore> rpcCoreOf(const py::handle& pyRpc) { return pyRpc.cast<std::shared_ptr<PyRpc>>()->core; }

GroupParts groupPartsOf(const py::handle& pyGroup) {
  auto g = pyGroup.cast<std::shared_ptr<PyGroup>>();
  return GroupParts{g->rpc->core, g->service, g->info, g->reducers, g};
}

py::object makeOwnGroup(const std::string& groupName) {
  auto rpc = std::make_shared<PyRpc>();
  auto g = std::make_shared<PyGroup>(rpc, groupName);
  return py::cast(g);
}

std::shared_ptr<PyFuture> makeReadyFuture(std::shared_ptr<FutureState> st, py::object value) {
  auto f = std::make_shared<PyFuture>();
  f->state = std::move(st);
  f->ready = std::move(value);
  return f;
}

}  // namespace mbh
