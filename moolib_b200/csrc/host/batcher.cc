// HP-B sink side: Batcher, UnrollBatcher, stack_fields / unstack_fields, to_device.
//
// API and observable behaviour follow moolib.Batcher (reference: src/moolib.cc:595-889 Batcher<T>, :1411-1488
// BatcherWrapper, bound at :1867-1935) and utils::stackFields / unstackFields (src/batch_utils.cc:259-325): same
// constructor, methods, nesting rules (dict / list / tuple / tensor / pass-through objects), error strings and cat carry.
//
// Design (not the reference's): an item is never walked twice and never produces tensor views.
//   * The first item of a batch is compiled into a NestPlan -- the flattened container tree with the leaf order, the
//     dict keys and the pass-through objects.  Every later item is only walked ALONG the plan to collect its tensor
//     leaves (type mismatches are found on the way).
//   * A batch is a flat vector of output leaves plus one LeafGeom per leaf (rows / bytes-per-index / pitches).  Stacking
//     item k or concatenating n columns is arithmetic on that geometry: one mb_copy_job per leaf, the whole item -- and
//     every batch it completes -- goes out as ONE launch of the pitched-copy kernels (include/moolib_b200.h).  The
//     reference issues one `select(dim,k).copy_()` / `narrow().copy_()` per leaf (src/moolib.cc:676,745-751).
//   * UnrollBatcher fuses `Batcher(T, dim=0).stack` x T with `Batcher(B', dim=1).cat`: the T items of an unroll are only
//     retained, and when the unroll is complete every (step, leaf, learner batch) piece is gathered straight into its
//     final [T, B', ...] place by ONE launch over a device-resident job table -- each observation byte moves once
//     instead of twice, in one big launch instead of T small ones and a re-tiling pass.
// Leaves the kernels cannot take (CPU batchers, dtype/shape-converting or pageable-host sources) use at::copy_ with the
// reference's semantics; they are API parity, not the hot path.
#include "common.h"

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <optional>

namespace mbh {

namespace {

[[noreturn]] void typeMismatch() { throw std::runtime_error("type mismatch in batch operation"); }

// ---------------------------------------------------------------------------------------------------------------------
// NestPlan
// ---------------------------------------------------------------------------------------------------------------------
class NestPlan {
 public:
  enum Kind : uint8_t { kDict, kList, kTuple, kTensor, kOther };

  void compile(const py::handle& item) {
    nodes_.clear();
    others_.clear();
    nTensors_ = 0;
    add(item);
  }
  bool compiled() const { return !nodes_.empty(); }
  size_t tensors() const { return nTensors_; }
  size_t others() const { return others_.size(); }

  // Walk `item` along the plan; tensor leaves are appended to `tensors`, pass-through leaves (if wanted) to `objects`.
  void collect(const py::handle& item, std::vector<torch::Tensor>& tensors, std::vector<py::object>* objects = nullptr) const {
    size_t cur = 0;
    walk(item, cur, tensors, objects);
  }

  // Rebuild the nest: tensor leaf i becomes tensorLeaf(i), pass-through leaf j becomes otherLeaf(j).
  template <class FT, class FO>
  py::object build(FT&& tensorLeaf, FO&& otherLeaf) const {
    size_t cur = 0;
    return make(cur, tensorLeaf, otherLeaf);
  }
  // ... around `leaves`, with the pass-through objects of the item the plan was compiled from (src/moolib.cc:689)
  py::object build(const std::vector<torch::Tensor>& leaves) const {
    return build([&](size_t i) { return to_python(leaves[i]); }, [&](size_t j) { return others_[j]; });
  }

 private:
  struct Node {
    Kind kind = kOther;
    uint32_t count = 0;  // children of a container
    uint32_t leaf = 0;   // kTensor: tensor index, kOther: pass-through index
    py::object key;      // set on the children of a dict
  };
  std::vector<Node> nodes_;  // pre-order
  std::vector<py::object> others_;
  size_t nTensors_ = 0;

  size_t add(const py::handle& v) {
    const size_t me = nodes_.size();
    nodes_.emplace_back();
    if (py::isinstance<py::dict>(v)) {
      uint32_t n = 0;
      for (auto kv : py::reinterpret_borrow<py::dict>(v)) {
        const size_t child = add(kv.second);
        nodes_[child].key = py::reinterpret_borrow<py::object>(kv.first);
        ++n;
      }
      nodes_[me].kind = kDict;
      nodes_[me].count = n;
    } else if (py::isinstance<py::list>(v) || py::isinstance<py::tuple>(v)) {
      const bool isList = py::isinstance<py::list>(v);
      uint32_t n = 0;
      for (auto x : py::reinterpret_borrow<py::sequence>(v)) {
        add(x);
        ++n;
      }
      nodes_[me].kind = isList ? kList : kTuple;
      nodes_[me].count = n;
    } else if (is_tensor(v)) {
      nodes_[me].kind = kTensor;
      nodes_[me].leaf = (uint32_t)nTensors_++;
    } else {
      nodes_[me].kind = kOther;
      nodes_[me].leaf = (uint32_t)others_.size();
      others_.push_back(py::reinterpret_borrow<py::object>(v));
    }
    return me;
  }

  void walk(const py::handle& v, size_t& cur, std::vector<torch::Tensor>& tensors, std::vector<py::object>* objects) const {
    const Node& nd = nodes_[cur++];
    switch (nd.kind) {
      case kDict: {
        if (!py::isinstance<py::dict>(v)) typeMismatch();
        for (uint32_t i = 0; i < nd.count; ++i) {
          PyObject* child = PyDict_GetItemWithError(v.ptr(), nodes_[cur].key.ptr());  // borrowed
          if (!child) {
            if (PyErr_Occurred()) throw py::error_already_set();
            throw py::key_error(py::repr(nodes_[cur].key).cast<std::string>());
          }
          walk(py::handle(child), cur, tensors, objects);
        }
        break;
      }
      case kList:
      case kTuple: {
        const bool ok = nd.kind == kList ? py::isinstance<py::list>(v) : py::isinstance<py::tuple>(v);
        if (!ok) typeMismatch();
        if ((uint32_t)PySequence_Fast_GET_SIZE(v.ptr()) < nd.count)
          throw py::index_error(nd.kind == kList ? "list index out of range" : "tuple index out of range");
        for (uint32_t i = 0; i < nd.count; ++i) walk(py::handle(PySequence_Fast_GET_ITEM(v.ptr(), i)), cur, tensors, objects);
        break;
      }
      case kTensor:
        if (!is_tensor(v)) typeMismatch();
        tensors.push_back(to_tensor(v));
        break;
      case kOther:
        if (objects) objects->push_back(py::reinterpret_borrow<py::object>(v));
        break;
    }
  }

  template <class FT, class FO>
  py::object make(size_t& cur, FT& tensorLeaf, FO& otherLeaf) const {
    const Node& nd = nodes_[cur++];
    switch (nd.kind) {
      case kDict: {
        py::dict d;
        for (uint32_t i = 0; i < nd.count; ++i) {
          py::object key = nodes_[cur].key;
          d[key] = make(cur, tensorLeaf, otherLeaf);
        }
        return std::move(d);
      }
      case kList: {
        py::list l(nd.count);
        for (uint32_t i = 0; i < nd.count; ++i) l[i] = make(cur, tensorLeaf, otherLeaf);
        return std::move(l);
      }
      case kTuple: {
        py::tuple t(nd.count);
        for (uint32_t i = 0; i < nd.count; ++i) t[i] = make(cur, tensorLeaf, otherLeaf);
        return std::move(t);
      }
      case kTensor:
        return tensorLeaf((size_t)nd.leaf);
      default:
        return otherLeaf((size_t)nd.leaf);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// CopyQueue: the pitched copies of one call, launched together
// ---------------------------------------------------------------------------------------------------------------------
class CopyQueue {
 public:
  ~CopyQueue() {
    if (ctx_) mb_copy_ctx_destroy(ctx_);
  }

  // Can the kernels read `src` as it is?  (contiguous; on `device`, or pinned host memory the device can address)
  static bool readable(const torch::Tensor& src, int device, bool* hostSrc) {
    if (!src.is_contiguous()) return false;  // a .contiguous() temporary of a host tensor would be pageable
    if (src.is_cuda()) {
      *hostSrc = false;
      return src.get_device() == device;
    }
    *hostSrc = true;
    return src.device().is_cpu() && src.is_pinned();
  }

  void add(const mb_copy_job& j, bool hostSrc, int device) {
    device_ = device;
    (hostSrc ? host_ : dev_).push_back(j);
  }

  void launch() {
    if (dev_.empty() && host_.empty()) return;
    c10::cuda::CUDAGuard g(device_);
    run(dev_, MB_SRC_DEVICE);
    run(host_, MB_SRC_HOST_MAPPED);
  }

  void discard() {
    dev_.clear();
    host_.clear();
  }

 private:
  std::vector<mb_copy_job> dev_, host_;
  int device_ = 0;
  mb_copy_ctx* ctx_ = nullptr;
  int ctxDevice_ = -1;

  void run(std::vector<mb_copy_job>& jobs, int kind) {
    if (jobs.empty()) return;
    int n;
    if (jobs.size() <= MB_COPY_MAX_INLINE_JOBS) {
      n = check(mb_copy2d_batch_ex(jobs.data(), (int)jobs.size(), kind, current_stream(device_)), "Batcher");
    } else {
      if (ctx_ && ctxDevice_ != device_) {
        mb_copy_ctx_destroy(ctx_);
        ctx_ = nullptr;
      }
      if (!ctx_) {
        check(mb_copy_ctx_create(device_, 16384, &ctx_), "mb_copy_ctx_create");
        ctxDevice_ = device_;
      }
      n = check(mb_copy2d_table(ctx_, jobs.data(), (int)jobs.size(), kind, current_stream(device_)), "Batcher");
    }
    launch_counter() += (uint64_t)n;
    jobs.clear();
  }
};

// Jobs that were described but not launched (an exception on the way) must not survive into the next call: their
// sources may be gone by then.
struct LaunchGuard {
  CopyQueue& q;
  bool launched = false;
  explicit LaunchGuard(CopyQueue& queue) : q(queue) {}
  void launch() {
    q.launch();
    launched = true;
  }
  ~LaunchGuard() {
    if (!launched) q.discard();
  }
};

// How one leaf of an item sits inside its batch tensor, in bytes.  `outer` rows; one index of the batch dimension is
// `inner` bytes; the batch tensor holds `dstCount` indices per row.
struct LeafGeom {
  int64_t outer = 1;
  int64_t inner = 1;
  int64_t dstCount = 1;
  bool kernel = false;  // destination is a contiguous CUDA tensor
  int device = 0;
};

int64_t prod(c10::IntArrayRef s, size_t from, size_t to) {
  int64_t p = 1;
  for (size_t i = from; i < to; ++i) p *= s[i];
  return p;
}

std::string dimsError(const char* op, size_t ndim, int64_t dim) {
  return "Given input tensor with " + std::to_string(ndim) + " dimensions, cannot " + op + " in dimension " +
         std::to_string(dim);
}

// dst.select(dim, k) <- src          (rows = outer, one slot of `inner` bytes per row)
void addStackCopy(CopyQueue& q, const torch::Tensor& dst, const LeafGeom& g, int64_t dim, int64_t k, const torch::Tensor& src,
                  c10::IntArrayRef expectSizes) {
  bool hostSrc = false;
  if (g.kernel && src.scalar_type() == dst.scalar_type() && src.sizes() == expectSizes &&
      CopyQueue::readable(src, g.device, &hostSrc)) {
    if (src.numel() == 0) return;
    mb_copy_job j;
    j.src = src.data_ptr();
    j.dst = static_cast<char*>(dst.data_ptr()) + k * g.inner;
    j.rows = (uint64_t)g.outer;
    j.row_bytes = (uint64_t)g.inner;
    j.src_pitch = g.inner;
    j.dst_pitch = g.dstCount * g.inner;
    q.add(j, hostSrc, g.device);
  } else {
    dst.select(dim, k).copy_(src, /*non_blocking=*/true);  // ATen semantics (conversion, broadcast, errors)
  }
}

// dst.narrow(dim, dstOff, n) <- src.narrow(dim, srcOff, n), where src has srcCount indices along dim
void addCatCopy(CopyQueue& q, const torch::Tensor& dst, const LeafGeom& g, int64_t dim, int64_t dstOff, const torch::Tensor& src,
                int64_t srcOff, int64_t n) {
  if (n == 0) return;
  bool hostSrc = false;
  bool same = src.dim() == dst.dim() && src.scalar_type() == dst.scalar_type();
  for (int64_t i = 0; same && i < dst.dim(); ++i) same = i == dim || src.size(i) == dst.size(i);
  if (g.kernel && same && CopyQueue::readable(src, g.device, &hostSrc)) {
    if (g.outer * g.inner == 0) return;
    const int64_t srcCount = src.size(dim);
    mb_copy_job j;
    j.src = static_cast<const char*>(src.data_ptr()) + srcOff * g.inner;
    j.dst = static_cast<char*>(dst.data_ptr()) + dstOff * g.inner;
    j.rows = (uint64_t)g.outer;
    j.row_bytes = (uint64_t)(n * g.inner);
    j.src_pitch = srcCount * g.inner;
    j.dst_pitch = g.dstCount * g.inner;
    q.add(j, hostSrc, g.device);
  } else {
    dst.narrow(dim, dstOff, n).copy_(src.narrow(dim, srcOff, n), /*non_blocking=*/true);
  }
}

LeafGeom geomOf(const torch::Tensor& batch, int64_t dim) {
  LeafGeom g;
  g.outer = prod(batch.sizes(), 0, (size_t)dim);
  g.inner = batch.element_size() * prod(batch.sizes(), (size_t)dim + 1, (size_t)batch.dim());
  g.dstCount = batch.size(dim);
  g.kernel = batch.is_cuda() && batch.is_contiguous();
  g.device = batch.is_cuda() ? batch.get_device() : 0;
  return g;
}

// ---------------------------------------------------------------------------------------------------------------------
// Batcher
// ---------------------------------------------------------------------------------------------------------------------
class Batcher {
 public:
  Batcher(int64_t size, const std::string& device, int64_t dim) : size_(size), dim_(dim), device_(device) {}

  // Batcher.stack (src/moolib.cc:813-845): returns the finished batch when this item completes it
  std::optional<py::object> stack(const py::object& item) {
    std::lock_guard<std::mutex> l(mu_);
    if (open_ && cat_)
      throw std::runtime_error("Batcher.stack: Previously called with cat; cannot mix cat/stack within the same batch");
    NestPlan fresh;
    const NestPlan& plan = open_ ? plan_ : (fresh.compile(item), fresh);
    leaves_.clear();
    plan.collect(item, leaves_);
    for (auto& t : leaves_)
      if ((int64_t)t.dim() < dim_) throw std::runtime_error(dimsError("stack", (size_t)t.dim(), dim_));
    if (!open_) {
      // the batch: every leaf gets `size` slots inserted at dim (src/moolib.cc:653-659)
      out_.clear();
      geom_.clear();
      itemSizes_.clear();
      for (auto& t : leaves_) {
        std::vector<int64_t> s(t.sizes().begin(), t.sizes().end());
        itemSizes_.push_back(s);
        s.insert(s.begin() + dim_, size_);
        out_.push_back(torch::empty(s, t.options().device(device_)));
        geom_.push_back(geomOf(out_.back(), dim_));
      }
      plan_ = std::move(fresh);
      open_ = true;
      cat_ = false;
      fill_ = 0;
    }
    LaunchGuard guard(q_);
    for (size_t i = 0; i < leaves_.size(); ++i) addStackCopy(q_, out_[i], geom_[i], dim_, fill_, leaves_[i], itemSizes_[i]);
    guard.launch();
    leaves_.clear();
    if (++fill_ == size_) return close();
    return {};
  }

  // Batcher.cat (src/moolib.cc:767-811): an item may finish several batches and start another; the remainder is
  // carried.  Everything it touches goes out as ONE launch, and the finished batches are handed over only after that
  // launch is enqueued, so a consumer can never get ahead of the copy on the stream.
  template <class Emit>
  void cat(const py::object& item, Emit&& emit) {
    std::vector<py::object> finished;
    {
      std::lock_guard<std::mutex> l(mu_);
      if (open_ && !cat_)
        throw std::runtime_error("Batcher.cat: Previously called with stack; cannot mix cat/stack within the same batch");
      NestPlan fresh;
      const NestPlan& plan = open_ ? plan_ : (fresh.compile(item), fresh);
      leaves_.clear();
      plan.collect(item, leaves_);
      int64_t n = 0;
      for (size_t i = 0; i < leaves_.size(); ++i) {
        const torch::Tensor& t = leaves_[i];
        if ((int64_t)t.dim() <= dim_) throw std::runtime_error(dimsError("cat", (size_t)t.dim(), dim_));
        if (i == 0) n = t.size(dim_);
        else if (t.size(dim_) != n)
          throw std::runtime_error(
              "Batch dimension size mismatch; during a cat operation, all tensors must have the same size in the batch "
              "dimension (" + std::to_string(dim_) + "). Got " + std::to_string(n) + " and " + std::to_string(t.size(dim_)));
      }
      int64_t taken = 0;
      LaunchGuard guard(q_);
      while (true) {
        if (!open_) {
          out_.clear();
          geom_.clear();
          for (auto& t : leaves_) {
            std::vector<int64_t> s(t.sizes().begin(), t.sizes().end());
            s[dim_] = size_;
            out_.push_back(torch::empty(s, t.options().device(device_)));
            geom_.push_back(geomOf(out_.back(), dim_));
          }
          // pass-through objects of a batch come from the item that opened it (src/moolib.cc:689)
          if (fresh.compiled()) plan_ = std::move(fresh);
          else plan_.compile(item);
          fresh = NestPlan();
          open_ = true;
          cat_ = true;
          fill_ = 0;
        }
        const int64_t take = std::min(n - taken, size_ - fill_);
        for (size_t i = 0; i < leaves_.size(); ++i) addCatCopy(q_, out_[i], geom_[i], dim_, fill_, leaves_[i], taken, take);
        fill_ += take;
        taken += take;
        if (fill_ == size_) finished.push_back(close());
        if (taken >= n) break;
      }
      guard.launch();
      leaves_.clear();
    }
    for (auto& r : finished) emit(std::move(r));
  }

  void reset() {
    std::lock_guard<std::mutex> l(mu_);
    out_.clear();
    leaves_.clear();
    plan_ = NestPlan();
    open_ = false;
  }

 private:
  const int64_t size_, dim_;
  const torch::Device device_;
  std::mutex mu_;
  bool open_ = false, cat_ = false;
  int64_t fill_ = 0;  // stack: items in the batch; cat: filled indices of the batch dimension
  NestPlan plan_;
  std::vector<torch::Tensor> out_, leaves_;
  std::vector<LeafGeom> geom_;
  std::vector<std::vector<int64_t>> itemSizes_;
  CopyQueue q_;

  py::object close() {
    py::object r = plan_.build(out_);
    out_.clear();
    open_ = false;
    return r;
  }
};

// Finished batches wait here for get() (reference: BatcherWrapper's queue + blocking get, src/moolib.cc:1411-1488)
class BatchQueue {
 public:
  ~BatchQueue() {
    py::gil_scoped_acquire gil;
    queue_.clear();
  }
  void push(py::object v) {
    {
      std::lock_guard<std::mutex> l(mu_);
      queue_.push_back(std::move(v));
    }
    cv_.notify_one();
  }
  bool empty() {
    std::lock_guard<std::mutex> l(mu_);
    return queue_.empty();
  }
  size_t size() {
    std::lock_guard<std::mutex> l(mu_);
    return queue_.size();
  }
  // blocks with the GIL released (another thread may be stacking), src/moolib.cc:296-314
  py::object get() {
    while (true) {
      {
        std::unique_lock<std::mutex> l(mu_);
        if (!queue_.empty()) {
          py::object r = std::move(queue_.front());
          queue_.pop_front();
          return r;
        }
      }
      {
        py::gil_scoped_release nogil;
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait_for(l, std::chrono::milliseconds(50), [&] { return !queue_.empty(); });
      }
      if (PyErr_CheckSignals() != 0) throw py::error_already_set();
    }
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<py::object> queue_;
};

struct PyBatcher {
  Batcher batcher;
  BatchQueue queue;
  PyBatcher(int64_t size, std::string device, int64_t dim) : batcher(size, device, dim) {
    if (size <= 0) throw std::runtime_error("Batcher: size must be positive");
  }
  ~PyBatcher() {
    py::gil_scoped_acquire gil;
    batcher.reset();
  }
  void stack(py::object data) {
    auto r = batcher.stack(data);
    if (r) queue.push(std::move(*r));
  }
  void cat(py::object data) {
    batcher.cat(data, [this](py::object v) { queue.push(std::move(v)); });
  }
  bool empty() { return queue.empty(); }
  size_t size() { return queue.size(); }
  py::object get() { return queue.get(); }
};

// ---------------------------------------------------------------------------------------------------------------------
// UnrollBatcher: Batcher(T, dim=0).stack x T fused with Batcher(batch_size, dim=cat_dim).cat
// ---------------------------------------------------------------------------------------------------------------------
class UnrollBatcher {
 public:
  UnrollBatcher(int64_t unroll, int64_t batchSize, const std::string& device, int64_t catDim)
      : T_(unroll), B_(batchSize), catDim_(catDim), device_(device) {
    if (unroll <= 0 || batchSize <= 0) throw std::runtime_error("UnrollBatcher: sizes must be positive");
    if (catDim < 1)
      throw std::runtime_error("UnrollBatcher: items are stacked along dimension 0, so cat_dim must be >= 1 (for other "
                               "layouts compose Batcher.stack and Batcher.cat)");
  }
  ~UnrollBatcher() {
    py::gil_scoped_acquire gil;
    held_.clear();
    out_.clear();
    xout_.clear();
    extra_.reset();
    plan_ = NestPlan();
    xplan_ = NestPlan();
  }

  // Retains the item's tensors (they must not be modified until the unroll is emitted -- the T-th stack() call).
  void stack(const py::object& item) {
    std::vector<py::object> finished;
    {
      std::lock_guard<std::mutex> l(mu_);
      if (held_.empty()) plan_.compile(item);
      std::vector<torch::Tensor> leaves;
      plan_.collect(item, leaves);
      for (auto& t : leaves)
        if ((int64_t)t.dim() < catDim_)
          throw std::runtime_error(dimsError("cat", (size_t)t.dim() + 1, catDim_));
      if (!held_.empty())
        for (size_t i = 0; i < leaves.size(); ++i)
          if (leaves[i].sizes() != held_[0][i].sizes() || leaves[i].scalar_type() != held_[0][i].scalar_type())
            throw std::runtime_error("UnrollBatcher.stack: every item of an unroll must have the same shapes and dtypes");
      held_.push_back(std::move(leaves));
      if ((int64_t)held_.size() == T_) emit(finished);
    }
    for (auto& r : finished) queue.push(std::move(r));
  }

  // A nest whose tensors are concatenated along cat_dim together with the NEXT emitted unroll and returned under `key`
  // of the (dict) result -- the `data["initial_core_state"] = ...` line of examples/vtrace/experiment.py:516-518.
  void setExtra(py::object key, py::object nest) {
    std::lock_guard<std::mutex> l(mu_);
    extraKey_ = std::move(key);
    extra_ = std::move(nest);
  }

  BatchQueue queue;

 private:
  const int64_t T_, B_, catDim_;
  const torch::Device device_;
  std::mutex mu_;
  NestPlan plan_, xplan_;
  std::vector<std::vector<torch::Tensor>> held_;  // [t][leaf]
  std::optional<py::object> extra_;
  py::object extraKey_;
  // the learner batch being filled (carried across unrolls when the actor batch is not a multiple of batch_size)
  bool open_ = false;
  int64_t fill_ = 0;
  std::vector<torch::Tensor> out_, xout_;
  std::vector<LeafGeom> geom_, xgeom_;
  NestPlan outPlan_, xoutPlan_;
  std::optional<py::object> outKey_;
  CopyQueue q_;
  std::vector<uint8_t> srcOk_;  // [t][leaf]: 0 = ATen copy, 1 = device source, 2 = pinned host source

  void emit(std::vector<py::object>& finished) {
    const std::vector<torch::Tensor>& first = held_[0];
    const int64_t ci = catDim_ - 1;  // the cat dimension inside one item
    int64_t n = 0;
    for (size_t i = 0; i < first.size(); ++i) {
      if (i == 0) n = first[i].size(ci);
      else if (first[i].size(ci) != n)
        throw std::runtime_error(
            "Batch dimension size mismatch; during a cat operation, all tensors must have the same size in the batch "
            "dimension (" + std::to_string(catDim_) + "). Got " + std::to_string(n) + " and " + std::to_string(first[i].size(ci)));
    }
    std::vector<torch::Tensor> xleaves;
    if (extra_) {
      xplan_.compile(*extra_);
      xplan_.collect(*extra_, xleaves);
      for (auto& t : xleaves) {
        if ((int64_t)t.dim() <= catDim_) throw std::runtime_error(dimsError("cat", (size_t)t.dim(), catDim_));
        if (first.empty() && &t == &xleaves[0]) n = t.size(catDim_);
        if (t.size(catDim_) != n)
          throw std::runtime_error(
              "Batch dimension size mismatch; during a cat operation, all tensors must have the same size in the batch "
              "dimension (" + std::to_string(catDim_) + "). Got " + std::to_string(n) + " and " + std::to_string(t.size(catDim_)));
      }
    }
    // an exception below (allocation failure, set_extra on a non-dict item, ...) abandons this unroll cleanly
    struct Abandon {
      UnrollBatcher& u;
      bool done = false;
      ~Abandon() {
        if (done) return;
        u.q_.discard();
        u.held_.clear();
        u.extra_.reset();
        u.out_.clear();
        u.xout_.clear();
        u.open_ = false;
      }
    } abandon{*this};
    // what the kernels may read directly, decided once per (step, leaf) -- not once per piece
    const size_t L = first.size();
    srcOk_.assign((size_t)T_ * L, 0);
    for (int64_t t = 0; t < T_; ++t)
      for (size_t i = 0; i < L; ++i) {
        bool hostSrc = false;
        const torch::Tensor& src = held_[(size_t)t][i];
        const int dev = device_.is_cuda() ? (device_.has_index() ? (int)device_.index() : (int)c10::cuda::current_device()) : -1;
        if (dev >= 0 && CopyQueue::readable(src, dev, &hostSrc)) srcOk_[(size_t)t * L + i] = hostSrc ? 2 : 1;
      }
    int64_t taken = 0;
    if (!open_ && n > 0 && n % B_ == 0 && alignedOk(first)) {
      // Aligned unroll (the actor batch is a whole number K of learner batches and the batch dimension is outermost):
      // the K batches of a leaf are slices of ONE allocation [K, T, B', ...], so a whole step of a leaf is ONE pitched
      // copy with K rows -- T x leaves jobs instead of T x leaves x K.
      const int64_t K = n / B_;
      std::vector<torch::Tensor> big;
      for (size_t i = 0; i < L; ++i) {
        std::vector<int64_t> sz(first[i].sizes().begin(), first[i].sizes().end());
        sz[0] = B_;
        sz.insert(sz.begin(), {K, T_});
        big.push_back(torch::empty(sz, first[i].options().device(device_)));
        const int64_t inner = big[i].element_size() * prod(big[i].sizes(), 3, (size_t)big[i].dim());
        const int dev = big[i].get_device();
        mb_copy_job j;
        j.rows = (uint64_t)K;
        j.row_bytes = (uint64_t)(B_ * inner);
        j.src_pitch = B_ * inner;
        j.dst_pitch = T_ * B_ * inner;
        for (int64_t t = 0; t < T_; ++t) {
          const torch::Tensor& src = held_[(size_t)t][i];
          if (j.row_bytes == 0) continue;
          j.src = src.data_ptr();
          j.dst = static_cast<char*>(big[i].data_ptr()) + t * B_ * inner;
          q_.add(j, srcOk_[(size_t)t * L + i] == 2, dev);  // alignedOk(): every source is kernel-readable
        }
      }
      for (int64_t k = 0; k < K; ++k) {
        openBatch(first, xleaves, /*allocate=*/false);
        for (size_t i = 0; i < L; ++i) out_.push_back(big[i].select(0, k));
        for (size_t i = 0; i < xleaves.size(); ++i) addCatCopy(q_, xout_[i], xgeom_[i], catDim_, 0, xleaves[i], k * B_, B_);
        finished.push_back(closeBatch());
      }
      taken = n;
    }
    while (taken < n) {
      if (!open_) openBatch(first, xleaves);
      const int64_t take = std::min(n - taken, B_ - fill_);
      for (size_t i = 0; i < L; ++i) {
        // step t of leaf i: out[i][t].narrow(ci, fill, take) <- item_t.narrow(ci, taken, take)
        const LeafGeom& g = geom_[i];
        const int64_t stepBytes = g.outer * g.dstCount * g.inner;
        char* const dstBase = static_cast<char*>(out_[i].data_ptr()) + fill_ * g.inner;
        mb_copy_job j;
        j.rows = (uint64_t)g.outer;
        j.row_bytes = (uint64_t)(take * g.inner);
        j.src_pitch = n * g.inner;
        j.dst_pitch = g.dstCount * g.inner;
        for (int64_t t = 0; t < T_; ++t) {
          const torch::Tensor& src = held_[(size_t)t][i];
          const uint8_t ok = srcOk_[(size_t)t * L + i];
          if (g.kernel && ok) {
            if (j.rows * j.row_bytes == 0) continue;
            j.src = static_cast<const char*>(src.data_ptr()) + taken * g.inner;
            j.dst = dstBase + t * stepBytes;
            q_.add(j, ok == 2, g.device);
          } else {
            out_[i].select(0, t).narrow(ci, fill_, take).copy_(src.narrow(ci, taken, take), /*non_blocking=*/true);
          }
        }
      }
      for (size_t i = 0; i < xleaves.size(); ++i) addCatCopy(q_, xout_[i], xgeom_[i], catDim_, fill_, xleaves[i], taken, take);
      fill_ += take;
      taken += take;
      if (fill_ == B_) finished.push_back(closeBatch());
    }
    q_.launch();  // ONE launch for the whole unroll: T x leaves (x pieces) pitched copies
    abandon.done = true;
    held_.clear();
    extra_.reset();
  }

  // the aligned fast path needs: a CUDA destination, the batch dimension outermost in every item, kernel-readable sources
  bool alignedOk(const std::vector<torch::Tensor>& first) const {
    if (catDim_ != 1 || !device_.is_cuda() || first.empty()) return false;
    for (uint8_t ok : srcOk_)
      if (!ok) return false;
    return true;
  }

  void openBatch(const std::vector<torch::Tensor>& first, const std::vector<torch::Tensor>& xleaves, bool allocate = true) {
    out_.clear();
    geom_.clear();
    xout_.clear();
    xgeom_.clear();
    for (auto& t : first) {
      if (!allocate) break;
      std::vector<int64_t> s(t.sizes().begin(), t.sizes().end());
      s[catDim_ - 1] = B_;
      s.insert(s.begin(), T_);
      out_.push_back(torch::empty(s, t.options().device(device_)));
      LeafGeom g = geomOf(out_.back(), catDim_);
      g.outer = prod(t.sizes(), 0, (size_t)catDim_ - 1);  // rows of ONE step; the step index is added by hand
      geom_.push_back(g);
    }
    for (auto& t : xleaves) {
      std::vector<int64_t> s(t.sizes().begin(), t.sizes().end());
      s[catDim_] = B_;
      xout_.push_back(torch::empty(s, t.options().device(device_)));
      xgeom_.push_back(geomOf(xout_.back(), catDim_));
    }
    outPlan_ = plan_;
    xoutPlan_ = xplan_;
    outKey_.reset();
    if (extra_) outKey_ = extraKey_;
    open_ = true;
    fill_ = 0;
  }

  py::object closeBatch() {
    py::object r = outPlan_.build(out_);
    if (outKey_) {
      if (!py::isinstance<py::dict>(r))
        throw std::runtime_error("UnrollBatcher: set_extra needs dict items (the extra nest is returned under its key)");
      py::reinterpret_borrow<py::dict>(r)[*outKey_] = xoutPlan_.build(xout_);
    }
    out_.clear();
    xout_.clear();
    open_ = false;
    return r;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// stack_fields / unstack_fields (reference behaviour: src/batch_utils.cc:246-325)
// ---------------------------------------------------------------------------------------------------------------------
bool isNode(const py::handle& h) {
  return py::isinstance<py::tuple>(h) || py::isinstance<py::list>(h) || py::isinstance<py::dict>(h) || is_tensor(h);
}

// N == 1: nothing is copied; tensors gain / lose the batch dimension as views, pass-through leaves are wrapped in /
// unwrapped from a 1-tuple (src/batch_utils.cc:261-263, 318-320)
py::object unsqueezeNest(const py::handle& v, int64_t dim) {
  NestPlan plan;
  plan.compile(v);
  std::vector<torch::Tensor> ts;
  std::vector<py::object> os;
  plan.collect(v, ts, &os);
  return plan.build([&](size_t i) { return to_python(ts[i].unsqueeze(dim)); },
                    [&](size_t j) { return py::object(py::make_tuple(os[j])); });
}

py::object squeezeNest(const py::handle& v, int64_t dim) {
  if (py::isinstance<py::tuple>(v)) {
    py::tuple src = py::reinterpret_borrow<py::tuple>(v);
    bool wraps = src.size() == 1;
    for (auto x : src) wraps = wraps && !isNode(x);
    if (wraps) return py::reinterpret_borrow<py::object>(src[0]);  // the 1-tuple around a pass-through leaf
    py::tuple dst(src.size());
    for (size_t i = 0; i < src.size(); ++i) dst[i] = squeezeNest(src[i], dim);
    return std::move(dst);
  }
  if (py::isinstance<py::list>(v)) {
    py::list src = py::reinterpret_borrow<py::list>(v), dst(src.size());
    for (size_t i = 0; i < src.size(); ++i) dst[i] = squeezeNest(src[i], dim);
    return std::move(dst);
  }
  if (py::isinstance<py::dict>(v)) {
    py::dict dst;
    for (auto kv : py::reinterpret_borrow<py::dict>(v)) dst[kv.first] = squeezeNest(kv.second, dim);
    return std::move(dst);
  }
  if (is_tensor(v)) return to_python(to_tensor(v).squeeze(dim));
  return py::reinterpret_borrow<py::object>(v);
}

// One value per batch element for the sub-nest `v`.
std::vector<py::object> unstackNest(const py::handle& v, int64_t n, int64_t dim) {
  std::vector<py::object> r((size_t)n);
  if (py::isinstance<py::tuple>(v)) {
    py::tuple src = py::reinterpret_borrow<py::tuple>(v);
    bool perItem = true;  // a tuple of plain objects IS the batch of a pass-through leaf (src/batch_utils.cc:288)
    for (auto x : src) perItem = perItem && !isNode(x);
    if (perItem) {
      if ((int64_t)src.size() < n) throw py::index_error("tuple index out of range");
      for (int64_t i = 0; i < n; ++i) r[(size_t)i] = py::reinterpret_borrow<py::object>(src[(size_t)i]);
      return r;
    }
    std::vector<std::vector<py::object>> ch;
    for (auto x : src) ch.push_back(unstackNest(x, n, dim));
    for (int64_t i = 0; i < n; ++i) {
      py::tuple t(ch.size());
      for (size_t j = 0; j < ch.size(); ++j) t[j] = ch[j][(size_t)i];
      r[(size_t)i] = std::move(t);
    }
    return r;
  }
  if (py::isinstance<py::list>(v)) {
    py::list src = py::reinterpret_borrow<py::list>(v);
    std::vector<std::vector<py::object>> ch;
    for (auto x : src) ch.push_back(unstackNest(x, n, dim));
    for (int64_t i = 0; i < n; ++i) {
      py::list t(ch.size());
      for (size_t j = 0; j < ch.size(); ++j) t[j] = ch[j][(size_t)i];
      r[(size_t)i] = std::move(t);
    }
    return r;
  }
  if (py::isinstance<py::dict>(v)) {
    for (int64_t i = 0; i < n; ++i) r[(size_t)i] = py::dict();
    for (auto kv : py::reinterpret_borrow<py::dict>(v)) {
      std::vector<py::object> ch = unstackNest(kv.second, n, dim);
      for (int64_t i = 0; i < n; ++i) py::reinterpret_borrow<py::dict>(r[(size_t)i])[kv.first] = ch[(size_t)i];
    }
    return r;
  }
  if (is_tensor(v)) {
    std::vector<torch::Tensor> parts = to_tensor(v).unbind(dim);  // views, as the reference (batch_utils.cc:233)
    if ((int64_t)parts.size() < n) throw py::index_error("unstack_fields: tensor is smaller than batch_size along dim");
    for (int64_t i = 0; i < n; ++i) r[(size_t)i] = to_python(parts[(size_t)i]);
    return r;
  }
  throw std::runtime_error("unstack_fields: a pass-through leaf must be a tuple of batch_size values");
}

}  // namespace

py::object stackFields(const py::tuple& input, int64_t dim) {
  const size_t n = input.size();
  if (n == 0) throw std::runtime_error("stack_fields: empty input");
  if (n == 1) return unsqueezeNest(input[0], dim);
  NestPlan plan;
  plan.compile(input[0]);
  std::vector<std::vector<torch::Tensor>> ts(n);
  std::vector<std::vector<py::object>> os(n);
  for (size_t i = 0; i < n; ++i) plan.collect(input[i], ts[i], &os[i]);
  // every tensor leaf: torch.stack of the n inputs -- all leaves of all inputs in ONE gather launch (K-B4)
  const size_t nl = plan.tensors();
  std::vector<torch::Tensor> stacked(nl);
  CopyQueue q;
  std::vector<torch::Tensor> column(n);
  for (size_t l = 0; l < nl; ++l) {
    const torch::Tensor& a = ts[0][l];
    bool uniform = a.is_cuda();
    for (size_t i = 0; i < n; ++i) {
      column[i] = ts[i][l];
      uniform = uniform && column[i].is_cuda() && column[i].get_device() == a.get_device() &&
                column[i].scalar_type() == a.scalar_type() && column[i].sizes() == a.sizes();
    }
    const int64_t d = dim < 0 ? dim + a.dim() + 1 : dim;
    if (!uniform || d < 0 || d > a.dim()) {
      stacked[l] = torch::stack(column, dim);  // CPU / mixed leaves, and ATen's own error for a bad dim
      continue;
    }
    std::vector<int64_t> s(a.sizes().begin(), a.sizes().end());
    s.insert(s.begin() + d, (int64_t)n);
    stacked[l] = torch::empty(s, a.options());
    const LeafGeom g = geomOf(stacked[l], d);
    for (size_t i = 0; i < n; ++i) addStackCopy(q, stacked[l], g, d, (int64_t)i, column[i], a.sizes());
  }
  q.launch();
  return plan.build([&](size_t l) { return to_python(stacked[l]); },
                    [&](size_t j) {
                      py::tuple t(n);
                      for (size_t i = 0; i < n; ++i) t[i] = os[i][j];
                      return py::object(std::move(t));
                    });
}

py::tuple unstackFields(const py::handle& input, int64_t batchSize, int64_t dim) {
  if (batchSize == 1) return py::make_tuple(squeezeNest(input, dim));
  std::vector<py::object> parts = unstackNest(input, batchSize, dim);
  py::tuple r((size_t)batchSize);
  for (int64_t i = 0; i < batchSize; ++i) r[(size_t)i] = std::move(parts[(size_t)i]);
  return r;
}

// Every tensor of `nest` on `device`: tensors already there are passed through, pinned host tensors are read by ONE
// launch of the copy kernel (host-mapped sources), anything else goes through at::to.  This is what
// EnvStepperFuture.result(device=...) returns (reference: src/env.cc:389-401 from_blob views followed by one
// `.to(device)` per key in examples/vtrace/experiment.py:492-494).
py::object nestToDevice(const py::handle& nest, const std::string& device) {
  const torch::Device dev(device);
  NestPlan plan;
  plan.compile(nest);
  std::vector<torch::Tensor> ts;
  plan.collect(nest, ts);
  if (!dev.is_cuda()) {
    for (auto& t : ts) t = t.to(dev);
    return plan.build(ts);
  }
  const int index = dev.has_index() ? dev.index() : c10::cuda::current_device();
  CopyQueue q;
  for (auto& t : ts) {
    bool hostSrc = false;
    if (t.is_cuda() && t.get_device() == index) continue;
    if (CopyQueue::readable(t, index, &hostSrc) && hostSrc && t.numel() > 0) {
      torch::Tensor d = torch::empty(t.sizes(), t.options().device(torch::kCUDA, index));
      mb_copy_job j;
      j.src = t.data_ptr();
      j.dst = d.data_ptr();
      j.rows = 1;
      j.row_bytes = (uint64_t)t.nbytes();
      j.src_pitch = j.dst_pitch = (int64_t)t.nbytes();
      q.add(j, true, index);
      t = d;
    } else {
      t = t.to(torch::Device(torch::kCUDA, index), /*non_blocking=*/true);
    }
  }
  q.launch();
  return plan.build(ts);
}

void bind_batcher(py::module_& m) {
  py::class_<PyBatcher>(m, "Batcher",
                        "Batches nested tensor structures along a dimension (moolib.Batcher API); device batches are "
                        "assembled by the sm_100a pitched-copy kernels, one launch per item.")
      .def(py::init<int64_t, std::string, int64_t>(), py::arg("size"), py::arg("device") = "cpu", py::arg("dim") = 0)
      .def("stack", &PyBatcher::stack, py::arg("tensors"))
      .def("cat", &PyBatcher::cat, py::arg("tensors"))
      .def("empty", &PyBatcher::empty)
      .def("size", &PyBatcher::size)
      .def("get", &PyBatcher::get);
  py::class_<UnrollBatcher>(m, "UnrollBatcher",
                            "Batcher(unroll, dim=0).stack fused with Batcher(batch_size, dim=cat_dim).cat: the items of "
                            "an unroll are retained and gathered straight into [unroll, batch_size, ...] learner batches "
                            "by one kernel launch when the unroll is complete (each byte moves once).")
      .def(py::init<int64_t, int64_t, std::string, int64_t>(), py::arg("unroll"), py::arg("batch_size"),
           py::arg("device"), py::arg("cat_dim") = 1)
      .def("stack", &UnrollBatcher::stack, py::arg("tensors"))
      .def("set_extra", &UnrollBatcher::setExtra, py::arg("key"), py::arg("tensors"))
      .def("empty", [](UnrollBatcher& u) { return u.queue.empty(); })
      .def("size", [](UnrollBatcher& u) { return u.queue.size(); })
      .def("get", [](UnrollBatcher& u) { return u.queue.get(); });
  m.def("stack_fields", &stackFields, py::arg("input"), py::arg("dim") = 0,
        "utils::stackFields (src/batch_utils.cc:259): stack N nested inputs leaf by leaf (one gather launch)");
  m.def("unstack_fields", &unstackFields, py::arg("input"), py::arg("batch_size"), py::arg("dim") = 0,
        "utils::unstackFields (src/batch_utils.cc:317)");
  m.def("to_device", &nestToDevice, py::arg("tensors"), py::arg("device"),
        "Move every tensor of a nest to `device`; pinned host tensors are read by one launch of the copy kernel");
  m.def("kernel_launches", [] { return launch_counter(); },
        "number of moolib_b200 kernels launched by this process through the host layer");
}

}  // namespace mbh
