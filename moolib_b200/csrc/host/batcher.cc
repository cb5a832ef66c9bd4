// Batcher (HP-B sink side) and stack/unstack of nested fields.
//
// Mirrors moolib.Batcher (reference: src/moolib.cc:595-889 Batcher<T>, :1411-1488 BatcherWrapper, bound at :1867-1935):
// same constructor, methods, nesting rules, error strings and carry semantics.  What changes is HOW the bytes move:
// the reference issues one `select(dim,k).copy_()` / `narrow().copy_()` per tensor leaf (src/moolib.cc:676,745-751);
// here every leaf of an item becomes one mb_copy_job and the whole item is ONE mb_copy2d_batch launch on the current
// stream.  CPU batchers (device="cpu") keep using at::copy_: they are API parity, not the hot path.
#include "common.h"

#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <thread>
#include <deque>
#include <mutex>
#include <optional>

namespace mbh {

uint64_t& launch_counter() {
  static uint64_t n = 0;
  return n;
}

namespace {
std::atomic<const char*> g_phase{"idle"};
std::atomic<int64_t> g_phase_ns{0};
}  // namespace

void trace_phase(const char* phase) {
  static const bool enabled = [] {
    const char* e = std::getenv("MOOLIB_B200_TRACE");
    if (!e || !*e || *e == '0') return false;
    std::thread([] {
      const char* last = nullptr;
      while (true) {
        std::this_thread::sleep_for(std::chrono::seconds(1));
        const char* ph = g_phase.load();
        int64_t age = std::chrono::steady_clock::now().time_since_epoch().count() - g_phase_ns.load();
        if (age > 3000000000ll && ph != last) {
          fprintf(stderr, "[moolib_b200 trace pid %d] stuck %.1f s in phase '%s'\n", (int)getpid(), age / 1e9, ph);
          fflush(stderr);
          last = ph;
        } else if (age <= 3000000000ll) {
          last = nullptr;
        }
      }
    }).detach();
    return true;
  }();
  if (!enabled) return;
  g_phase.store(phase);
  g_phase_ns.store(std::chrono::steady_clock::now().time_since_epoch().count());
}

namespace {

std::string fmt_dims(const char* op, size_t ndim, int64_t dim) {
  return "Given input tensor with " + std::to_string(ndim) + " dimensions, cannot " + op + " in dimension " +
         std::to_string(dim);
}

// Collects the pitched copies of one item and launches them together.
struct CopyBatch {
  std::vector<mb_copy_job> jobs;
  std::vector<torch::Tensor> keepalive;  // contiguous temporaries: stay alive until the launch is enqueued
  int device = -1;

  // dst_full: freshly allocated contiguous batch tensor; copies src into dst_full.narrow(dim, off, n) (n == -1:
  // dst_full.select(dim, off)), reading src.narrow(dim, src_off, n) when n >= 0.
  void add(const torch::Tensor& dst_full, int64_t dim, int64_t off, int64_t n, const torch::Tensor& src,
           int64_t src_off) {
    const bool select = n < 0;
    torch::Tensor dview = select ? dst_full.select(dim, off) : dst_full.narrow(dim, off, n);
    torch::Tensor sview = select ? src : ((src_off == 0 && n == src.size(dim)) ? src : src.narrow(dim, src_off, n));
    const bool kernel_ok = dst_full.is_cuda() && dst_full.is_contiguous() && sview.scalar_type() == dview.scalar_type() &&
                           sview.sizes() == dview.sizes() &&
                           ((sview.is_cuda() && sview.get_device() == dst_full.get_device()) ||
                            (!sview.is_cuda() && sview.is_pinned()));
    if (!kernel_ok) {
      // dtype/shape-converting or pageable-host copies keep the reference's copy_ semantics
      dview.copy_(sview, /*non_blocking=*/true);
      return;
    }
    if (sview.numel() == 0) return;
    torch::Tensor s = src.is_contiguous() ? src : src.contiguous();
    if (!src.is_contiguous()) keepalive.push_back(s);
    const int64_t esz = dst_full.element_size();
    int64_t outer = 1, inner = esz;
    for (int64_t i = 0; i < dim; ++i) outer *= dst_full.size(i);
    for (int64_t i = dim + 1; i < dst_full.dim(); ++i) inner *= dst_full.size(i);
    mb_copy_job j;
    if (select) {
      j.src = s.data_ptr();
      j.dst = static_cast<char*>(dst_full.data_ptr()) + off * inner;
      j.rows = (uint64_t)outer;
      j.row_bytes = (uint64_t)inner;
      j.src_pitch = inner;
      j.dst_pitch = dst_full.size(dim) * inner;
    } else {
      j.src = static_cast<const char*>(s.data_ptr()) + src_off * inner;
      j.dst = static_cast<char*>(dst_full.data_ptr()) + off * inner;
      j.rows = (uint64_t)outer;
      j.row_bytes = (uint64_t)(n * inner);
      j.src_pitch = s.size(dim) * inner;
      j.dst_pitch = dst_full.size(dim) * inner;
    }
    device = dst_full.get_device();
    jobs.push_back(j);
  }

  void launch() {
    if (jobs.empty()) return;
    c10::cuda::CUDAGuard g(device);
    int n = check(mb_copy2d_batch(jobs.data(), (int)jobs.size(), current_stream(device)), "Batcher");
    launch_counter() += (uint64_t)n;
    jobs.clear();
    keepalive.clear();
  }
};

struct Batcher {
  std::optional<py::object> target;
  int64_t nextStackIndex = 0;
  int64_t batchSize = 0;
  int64_t batchDimension = 0;
  torch::Device device{torch::kCPU};
  int nTensors = 0;
  int currentTensor = 0;
  int64_t catBatchInputOffset = 0;
  int64_t catBatchInputSize = 0;
  int64_t catBatchOutputOffset = 0;
  bool isDoingCat = false;
  std::vector<int64_t> sizes;
  CopyBatch copies;
  std::mutex batchMutex;

  Batcher(int64_t batchSize, const std::string& dev, int64_t dim)
      : batchSize(batchSize), batchDimension(dim), device(dev) {}

  // reference: src/moolib.cc:619-691
  template <bool cat>
  py::object prepareForBatchCopy(const py::handle& v) {
    if (py::isinstance<py::dict>(v)) {
      py::dict newdict;
      for (auto item : py::reinterpret_borrow<py::dict>(v)) newdict[item.first] = prepareForBatchCopy<cat>(item.second);
      return std::move(newdict);
    } else if (py::isinstance<py::list>(v)) {
      py::list list = py::reinterpret_borrow<py::list>(v);
      size_t n = list.size();
      py::list newlist(n);
      for (size_t i = 0; i != n; ++i) newlist[i] = prepareForBatchCopy<cat>(list[i]);
      return std::move(newlist);
    } else if (is_tensor(v)) {
      torch::Tensor t = to_tensor(v);
      auto s = t.sizes();
      if ((int64_t)s.size() <= (cat ? batchDimension : batchDimension - 1)) {
        throw std::runtime_error(fmt_dims(cat ? "cat" : "stack", s.size(), batchDimension));
      }
      if (cat) {
        sizes.assign(s.begin(), s.end());
        sizes[batchDimension] = batchSize;
      } else {
        sizes.resize(1 + s.size());
        std::copy(s.begin(), s.begin() + batchDimension, sizes.begin());
        std::copy(s.begin() + batchDimension, s.end(), sizes.begin() + batchDimension + 1);
        sizes[batchDimension] = batchSize;
      }
      torch::Tensor tensor = torch::empty(sizes, t.options().device(device));
      if (cat) {
        int64_t offset = catBatchInputOffset;
        int64_t n = s[batchDimension];
        if (offset > n) throw std::runtime_error("Batch internal error: offset > n");
        if (nTensors == 0) {
          catBatchInputSize = n;
        } else if (n != catBatchInputSize) {
          throw std::runtime_error(
              "Batch dimension size mismatch; during a cat operation, all tensors must have the same size in the "
              "batch dimension (" + std::to_string(batchDimension) + "). Got " + std::to_string(catBatchInputSize) +
              " and " + std::to_string(n));
        }
        n -= offset;
        n = std::min(n, batchSize);
        copies.add(tensor, batchDimension, 0, n, t, offset);
      } else {
        copies.add(tensor, batchDimension, 0, -1, t, 0);
      }
      ++nTensors;
      return to_python(tensor);
    } else if (py::isinstance<py::tuple>(v)) {
      py::tuple tuple = py::reinterpret_borrow<py::tuple>(v);
      size_t n = tuple.size();
      py::tuple newtuple(n);
      for (size_t i = 0; i != n; ++i) newtuple[i] = prepareForBatchCopy<cat>(tuple[i]);
      return std::move(newtuple);
    } else {
      return py::reinterpret_borrow<py::object>(v);
    }
  }

  // reference: src/moolib.cc:693-765
  template <bool cat>
  void visit(const py::handle& dest, const py::handle& source) {
    if (py::isinstance<py::dict>(dest)) {
      if (!py::isinstance<py::dict>(source)) throw std::runtime_error("type mismatch in batch operation");
      py::dict sourceDict = py::reinterpret_borrow<py::dict>(source);
      for (auto item : py::reinterpret_borrow<py::dict>(dest)) visit<cat>(item.second, sourceDict[item.first]);
    } else if (py::isinstance<py::list>(dest)) {
      if (!py::isinstance<py::list>(source)) throw std::runtime_error("type mismatch in batch operation");
      py::list sourceList = py::reinterpret_borrow<py::list>(source);
      py::list destList = py::reinterpret_borrow<py::list>(dest);
      size_t n = destList.size();
      for (size_t i = 0; i != n; ++i) visit<cat>(destList[i], sourceList[i]);
    } else if (is_tensor(dest)) {
      if (!is_tensor(source)) throw std::runtime_error("type mismatch in batch operation");
      torch::Tensor destT = to_tensor(dest);
      torch::Tensor sourceT = to_tensor(source);
      auto s = sourceT.sizes();
      if ((int64_t)s.size() <= (cat ? batchDimension : batchDimension - 1)) {
        throw std::runtime_error(fmt_dims(cat ? "cat" : "stack", s.size(), batchDimension));
      }
      if (cat) {
        int64_t inputOffset = catBatchInputOffset;
        int64_t n = s[batchDimension];
        if (inputOffset > n) throw std::runtime_error("Batch internal error: offset > n");
        if (currentTensor == 0) {
          catBatchInputSize = n;
        } else if (n != catBatchInputSize) {
          throw std::runtime_error(
              "Batch dimension size mismatch; during a cat operation, all tensors must have the same size in the "
              "batch dimension (" + std::to_string(batchDimension) + "). Got " + std::to_string(catBatchInputSize) +
              " and " + std::to_string(n));
        }
        int64_t outputOffset = catBatchOutputOffset;
        int64_t left = batchSize - outputOffset;
        n -= inputOffset;
        n = std::min(n, left);
        copies.add(destT, batchDimension, outputOffset, n, sourceT, inputOffset);
      } else {
        copies.add(destT, batchDimension, nextStackIndex, -1, sourceT, 0);
      }
      ++currentTensor;
    } else if (py::isinstance<py::tuple>(dest)) {
      if (!py::isinstance<py::tuple>(source)) throw std::runtime_error("type mismatch in batch operation");
      py::tuple sourceTuple = py::reinterpret_borrow<py::tuple>(source);
      py::tuple destTuple = py::reinterpret_borrow<py::tuple>(dest);
      size_t n = destTuple.size();
      for (size_t i = 0; i != n; ++i) visit<cat>(destTuple[i], sourceTuple[i]);
    }
  }

  // reference: src/moolib.cc:767-811.  Same carry loop; the copies of ALL batches this item completes (plus the
  // partial one it starts) go out as ONE launch, and the finished batches are handed over only after that launch is
  // enqueued, so a consumer can never get ahead of the copy on the stream.
  template <typename Callback>
  void cat(py::object value, Callback&& callback) {
    int64_t localInputOffset = 0;
    std::vector<py::object> finished;
    std::exception_ptr error;
    {
      std::unique_lock<std::mutex> l(batchMutex);
      try {
      while (true) {
        catBatchInputOffset = localInputOffset;
        catBatchInputSize = 0;
        if (!target) {
          catBatchOutputOffset = 0;
          nTensors = 0;
          target = prepareForBatchCopy<true>(value);
          isDoingCat = true;
        } else {
          if (!isDoingCat) {
            throw std::runtime_error(
                "Batcher.cat: Previously called with stack; cannot mix cat/stack within the same batch");
          }
          currentTensor = 0;
          visit<true>(*target, value);
          if (currentTensor != nTensors) {
            throw std::runtime_error("num tensors mismatch in batch operation; got " + std::to_string(currentTensor) +
                                     " tensors, batch has " + std::to_string(nTensors));
          }
        }
        int64_t inputSize = catBatchInputSize - localInputOffset;
        int64_t left = batchSize - catBatchOutputOffset;
        if (inputSize >= left) {
          finished.push_back(std::move(*target));
          target.reset();
          if (inputSize == left) break;
          localInputOffset += left;
        } else {
          catBatchOutputOffset += inputSize;
          break;
        }
      }
      } catch (...) {
        error = std::current_exception();  // the copies already described are still issued, as the reference's were
      }
      copies.launch();
    }
    for (auto& r : finished) callback(std::move(r));
    if (error) std::rethrow_exception(error);
  }

  // reference: src/moolib.cc:813-845
  std::optional<py::object> stack(py::object value) {
    std::lock_guard<std::mutex> l(batchMutex);
    if (!target) {
      nTensors = 0;
      try {
        target = prepareForBatchCopy<false>(value);
      } catch (...) {
        copies.launch();
        throw;
      }
      nextStackIndex = 1;
      isDoingCat = false;
    } else {
      if (isDoingCat) {
        throw std::runtime_error(
            "Batcher.stack: Previously called with cat; cannot mix cat/stack within the same batch");
      }
      currentTensor = 0;
      try {
        visit<false>(*target, value);
      } catch (...) {
        copies.launch();
        throw;
      }
      if (currentTensor != nTensors) {
        copies.launch();
        throw std::runtime_error("num tensors mismatch in batch operation; got " + std::to_string(currentTensor) +
                                 " tensors, batch has " + std::to_string(nTensors));
      }
      ++nextStackIndex;
    }
    copies.launch();
    if (nextStackIndex == batchSize) {
      py::object r = std::move(*target);
      target.reset();
      return r;
    }
    return {};
  }
};

// reference: src/moolib.cc:1411-1488 BatcherWrapper (queue of finished batches, blocking get)
struct BatcherWrapper {
  Batcher batcher;
  std::mutex mutex;
  std::condition_variable cv;
  std::deque<py::object> queue;

  BatcherWrapper(int64_t size, std::string device, int64_t dim) : batcher(size, device, dim) {
    if (size <= 0) throw std::runtime_error("Batcher: size must be positive");
  }
  ~BatcherWrapper() {
    py::gil_scoped_acquire gil;
    queue.clear();
    batcher.target.reset();
  }

  void enqueue(py::object value) {
    {
      std::lock_guard<std::mutex> l(mutex);
      queue.push_back(std::move(value));
    }
    cv.notify_one();
  }
  bool empty() {
    std::lock_guard<std::mutex> l(mutex);
    return queue.empty();
  }
  size_t size() {
    std::lock_guard<std::mutex> l(mutex);
    return queue.size();
  }
  py::object get() {
    {
      std::unique_lock<std::mutex> l(mutex);
      if (!queue.empty()) {
        py::object r = std::move(queue.front());
        queue.pop_front();
        return r;
      }
    }
    // blocking wait with the GIL released (another thread may be stacking), src/moolib.cc:296-314
    while (true) {
      {
        py::gil_scoped_release nogil;
        std::unique_lock<std::mutex> l(mutex);
        cv.wait_for(l, std::chrono::milliseconds(50), [&] { return !queue.empty(); });
      }
      std::unique_lock<std::mutex> l(mutex);
      if (!queue.empty()) {
        py::object r = std::move(queue.front());
        queue.pop_front();
        return r;
      }
      l.unlock();
      if (PyErr_CheckSignals() != 0) throw py::error_already_set();
    }
  }
  void stack(py::object data) {
    auto r = batcher.stack(std::move(data));
    if (r) enqueue(std::move(*r));
  }
  void cat(py::object data) {
    batcher.cat(std::move(data), [this](py::object v) { enqueue(std::move(v)); });
  }
};

// ---- stack_fields / unstack_fields (reference: src/batch_utils.cc:246-325) ---------------------------------------

template <class F>
void visitNested(F&& f, const py::handle& in) {
  if (py::isinstance<py::tuple>(in) || py::isinstance<py::list>(in)) {
    for (auto x : py::reinterpret_borrow<py::sequence>(in)) visitNested(f, x);
  } else if (py::isinstance<py::dict>(in)) {
    for (auto kv : py::reinterpret_borrow<py::dict>(in)) visitNested(f, kv.second);
  } else {
    f(in);
  }
}

template <class F>
py::object mapNested(F&& f, const py::handle& in) {
  if (py::isinstance<py::tuple>(in)) {
    py::tuple src = py::reinterpret_borrow<py::tuple>(in);
    py::tuple dst(src.size());
    for (size_t i = 0; i < src.size(); ++i) dst[i] = mapNested(f, src[i]);
    return std::move(dst);
  }
  if (py::isinstance<py::list>(in)) {
    py::list src = py::reinterpret_borrow<py::list>(in);
    py::list dst(src.size());
    for (size_t i = 0; i < src.size(); ++i) dst[i] = mapNested(f, src[i]);
    return std::move(dst);
  }
  if (py::isinstance<py::dict>(in)) {
    py::dict dst;
    for (auto kv : py::reinterpret_borrow<py::dict>(in)) dst[kv.first] = mapNested(f, kv.second);
    return std::move(dst);
  }
  return f(in);
}

// torch::stack of N same-shape leaves as ONE gather launch (K-B4), falling back to at::stack when the leaves are not
// same-device contiguous CUDA tensors of one dtype/shape.
torch::Tensor stackLeaves(const std::vector<torch::Tensor>& ts, int64_t dim) {
  const torch::Tensor& a = ts[0];
  bool ok = a.is_cuda();
  for (auto& t : ts)
    ok = ok && t.is_cuda() && t.get_device() == a.get_device() && t.scalar_type() == a.scalar_type() &&
         t.sizes() == a.sizes();
  if (!ok) return torch::stack(ts, dim);
  int64_t d = dim < 0 ? dim + a.dim() + 1 : dim;
  if (d < 0 || d > a.dim()) return torch::stack(ts, dim);  // let ATen raise its own error
  std::vector<int64_t> sizes(a.sizes().begin(), a.sizes().end());
  sizes.insert(sizes.begin() + d, (int64_t)ts.size());
  torch::Tensor out = torch::empty(sizes, a.options());
  CopyBatch cb;
  for (size_t i = 0; i < ts.size(); ++i) cb.add(out, d, (int64_t)i, -1, ts[i], 0);
  cb.launch();
  return out;
}

py::object unsqueezeFields(const py::handle& input, int64_t dim) {
  return mapNested(
      [dim](const py::handle& h) -> py::object {
        return is_tensor(h) ? to_python(to_tensor(h).unsqueeze(dim)) : py::object(py::make_tuple(h));
      },
      input);
}

std::pair<py::object, bool> squeezeFieldsImpl(const py::handle& input, int64_t dim) {
  if (py::isinstance<py::tuple>(input)) {
    py::tuple src = py::reinterpret_borrow<py::tuple>(input);
    const int64_t n = src.size();
    py::tuple dst(n);
    bool anyNode = false;
    for (int64_t i = 0; i < n; ++i) {
      auto [cur, tag] = squeezeFieldsImpl(src[i], dim);
      dst[i] = std::move(cur);
      anyNode |= tag;
    }
    if (n == 1 && !anyNode) return {py::object(dst[0]), true};
    return {std::move(dst), true};
  }
  if (py::isinstance<py::list>(input)) {
    py::list src = py::reinterpret_borrow<py::list>(input);
    py::list dst(src.size());
    for (size_t i = 0; i < src.size(); ++i) dst[i] = squeezeFieldsImpl(src[i], dim).first;
    return {std::move(dst), true};
  }
  if (py::isinstance<py::dict>(input)) {
    py::dict dst;
    for (auto kv : py::reinterpret_borrow<py::dict>(input)) dst[kv.first] = squeezeFieldsImpl(kv.second, dim).first;
    return {std::move(dst), true};
  }
  if (is_tensor(input)) return {to_python(to_tensor(input).squeeze(dim)), true};
  return {py::reinterpret_borrow<py::object>(input), false};
}

}  // namespace

// reference: src/batch_utils.cc:259-315
py::object stackFields(const py::tuple& input, int64_t dim) {
  if (input.size() == 0) throw std::runtime_error("stack_fields: empty input");
  if (input.size() == 1) return unsqueezeFields(input[0], dim);
  const int64_t batchSize = input.size();
  std::vector<std::vector<torch::Tensor>> tensors;
  std::vector<py::tuple> objects;
  size_t tensorIndex = 0, objectIndex = 0;
  for (int64_t i = 0; i < batchSize; ++i) {
    tensorIndex = 0;
    objectIndex = 0;
    visitNested(
        [&](const py::handle& h) {
          if (is_tensor(h)) {
            if (tensorIndex >= tensors.size()) tensors.emplace_back(batchSize);
            tensors[tensorIndex][i] = to_tensor(h);
            ++tensorIndex;
          } else {
            if (objectIndex >= objects.size()) objects.emplace_back(batchSize);
            objects[objectIndex][i] = py::reinterpret_borrow<py::object>(h);
            ++objectIndex;
          }
        },
        input[i]);
  }
  std::vector<torch::Tensor> stacked;
  stacked.reserve(tensors.size());
  for (auto& cur : tensors) stacked.push_back(stackLeaves(cur, dim));
  tensorIndex = 0;
  objectIndex = 0;
  return mapNested(
      [&](const py::handle& h) -> py::object {
        if (is_tensor(h)) return to_python(stacked[tensorIndex++]);
        return std::move(objects[objectIndex++]);
      },
      input[0]);
}

namespace {

bool prepareForUnstack(const py::handle& input, std::vector<bool>& batchTuple) {
  if (py::isinstance<py::tuple>(input)) {
    const size_t cur = batchTuple.size();
    batchTuple.push_back(false);
    bool anyNode = false;
    for (auto x : py::reinterpret_borrow<py::tuple>(input)) anyNode |= prepareForUnstack(x, batchTuple);
    batchTuple[cur] = !anyNode;
    return true;
  }
  if (py::isinstance<py::list>(input)) {
    for (auto x : py::reinterpret_borrow<py::list>(input)) prepareForUnstack(x, batchTuple);
    return true;
  }
  if (py::isinstance<py::dict>(input)) {
    for (auto kv : py::reinterpret_borrow<py::dict>(input)) prepareForUnstack(kv.second, batchTuple);
    return true;
  }
  return is_tensor(input);
}

template <class Sequence>
py::tuple unstackSequence(int64_t batchSize, std::vector<py::tuple>& src) {
  py::tuple dst(batchSize);
  const int64_t inner = src.size();
  for (int64_t i = 0; i < batchSize; ++i) {
    Sequence cur(inner);
    for (int64_t j = 0; j < inner; ++j) cur[j] = src[j][i];
    dst[i] = std::move(cur);
  }
  return dst;
}

py::tuple unstackFieldsImpl(const py::handle& input, int64_t batchSize, int64_t dim, const std::vector<bool>& batchTuple,
                            size_t& tupleIndex) {
  if (py::isinstance<py::tuple>(input)) {
    py::tuple src = py::reinterpret_borrow<py::tuple>(input);
    if (batchTuple[tupleIndex++]) return src;
    std::vector<py::tuple> children(src.size());
    for (size_t i = 0; i < src.size(); ++i) children[i] = unstackFieldsImpl(src[i], batchSize, dim, batchTuple, tupleIndex);
    return unstackSequence<py::tuple>(batchSize, children);
  }
  if (py::isinstance<py::list>(input)) {
    py::list src = py::reinterpret_borrow<py::list>(input);
    std::vector<py::tuple> children(src.size());
    for (size_t i = 0; i < src.size(); ++i) children[i] = unstackFieldsImpl(src[i], batchSize, dim, batchTuple, tupleIndex);
    return unstackSequence<py::list>(batchSize, children);
  }
  if (py::isinstance<py::dict>(input)) {
    py::tuple dst(batchSize);
    for (int64_t i = 0; i < batchSize; ++i) dst[i] = py::dict();
    for (auto kv : py::reinterpret_borrow<py::dict>(input)) {
      py::tuple cur = unstackFieldsImpl(kv.second, batchSize, dim, batchTuple, tupleIndex);
      for (int64_t i = 0; i < batchSize; ++i) py::reinterpret_borrow<py::dict>(dst[i])[kv.first] = cur[i];
    }
    return dst;
  }
  if (is_tensor(input)) {
    py::tuple dst(batchSize);
    std::vector<torch::Tensor> parts = to_tensor(input).unbind(dim);  // views, as the reference (batch_utils.cc:233)
    for (int64_t i = 0; i < batchSize; ++i) dst[i] = to_python(parts[i]);
    return dst;
  }
  return py::tuple();
}

}  // namespace

// reference: src/batch_utils.cc:317-325
py::tuple unstackFields(const py::handle& input, int64_t batchSize, int64_t dim) {
  if (batchSize == 1) return py::make_tuple(squeezeFieldsImpl(input, dim).first);
  std::vector<bool> batchTuple;
  prepareForUnstack(input, batchTuple);
  size_t tupleIndex = 0;
  return unstackFieldsImpl(input, batchSize, dim, batchTuple, tupleIndex);
}

void bind_batcher(py::module_& m) {
  py::class_<BatcherWrapper>(m, "Batcher",
                             "Batches nested tensor structures along a dimension (moolib.Batcher API); device batches "
                             "are assembled by the sm_100a pitched-copy kernels, one launch per item.")
      .def(py::init<int64_t, std::string, int64_t>(), py::arg("size"), py::arg("device") = "cpu", py::arg("dim") = 0)
      .def("stack", &BatcherWrapper::stack, py::arg("tensors"))
      .def("cat", &BatcherWrapper::cat, py::arg("tensors"))
      .def("empty", &BatcherWrapper::empty)
      .def("size", &BatcherWrapper::size)
      .def("get", &BatcherWrapper::get);
  m.def("stack_fields", &stackFields, py::arg("input"), py::arg("dim") = 0,
        "utils::stackFields (src/batch_utils.cc:259): stack N nested inputs leaf by leaf");
  m.def("unstack_fields", &unstackFields, py::arg("input"), py::arg("batch_size"), py::arg("dim") = 0,
        "utils::unstackFields (src/batch_utils.cc:317)");
  m.def("kernel_launches", [] { return launch_counter(); },
        "number of moolib_b200 kernels launched by this process through the host layer");
}

}  // namespace mbh
