// Device-side reducers of a Group: owns the mb_ar_ctx objects (symmetric staging + flags, include/moolib_b200.h),
// exchanges their handles over the control plane whenever the group's syncId changes and drives the K-A1/K-A2 kernels.
#pragma once

#include "common.h"
#include "control.h"

#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime_api.h>

namespace mbh {

// Python-visible future (reference: FutureWrapper / AllReduceWrapper, src/moolib.cc:201-393, 1292-1303)
struct PyFuture {
  std::shared_ptr<FutureState> state;
  std::function<py::object(const Bytes&)> decode;  // runs on the Python thread with the GIL
  std::shared_ptr<void> keep;                      // keeps the underlying operation alive
  std::optional<py::object> ready;                 // result object, when it is not decoded from bytes
  std::function<void()> progress;                  // optional: advances a device-side operation (Python thread)

  bool done();
  void wait(double timeout);
  py::object get();
  py::object result(std::optional<double> timeout);
  py::object exception();
  void cancel();
};

// One mb_ar_ctx bound to the group's current epoch.
class DeviceReducer {
 public:
  DeviceReducer(std::shared_ptr<GroupService> service, std::shared_ptr<GroupInfo> info, std::string tag, int device,
                uint64_t maxBytes, int nslots);
  ~DeviceReducer();

  // Non-blocking: (re)starts the handle exchange when the group epoch changed, imports peers when it completes.
  // Returns true when the context is connected for the group's current syncId.
  bool poll();
  bool failed() const { return failed_; }
  const std::string& error() const { return error_; }
  mb_ar_ctx* ctx() { return ctx_; }
  uint32_t syncId() const { return syncId_; }
  int world() const { return world_; }
  int rank() const { return rank_; }
  int device() const { return device_; }
  uint64_t maxBytes() const { return maxBytes_; }
  // The context's own stream: kernels that wait for peers (K-A0) must not sit in front of unrelated work -- or of
  // another context's gate -- on the caller's stream.
  cudaStream_t stream();

 private:
  std::shared_ptr<GroupService> service_;
  std::shared_ptr<GroupInfo> info_;
  std::string tag_;
  int device_;
  uint64_t maxBytes_;
  int nslots_;
  mb_ar_ctx* ctx_ = nullptr;
  uint32_t syncId_ = 0;  // epoch the context is (being) connected for
  int world_ = 0, rank_ = 0;
  bool connected_ = false, failed_ = false;
  std::string error_;
  std::shared_ptr<SmallReduce> exchange_;
  cudaStream_t stream_ = nullptr;
};

class DeviceReducerSet {
 public:
  DeviceReducerSet(std::shared_ptr<GroupService> service, std::shared_ptr<GroupInfo> info)
      : service_(std::move(service)), info_(std::move(info)) {}
  std::shared_ptr<DeviceReducer> get(const std::string& tag, int device, uint64_t maxBytes, int nslots);
  // group.all_reduce(name, cuda_tensor): in-place sum over the members (A8)
  std::shared_ptr<PyFuture> allReduceTensor(const std::string& name, torch::Tensor t, py::object pyTensor);
  // advance every device all_reduce of this group that is still in flight (called from Group.update())
  void progressAll();

 private:
  std::shared_ptr<GroupService> service_;
  std::shared_ptr<GroupInfo> info_;
  std::mutex mu_;
  std::map<std::string, std::shared_ptr<DeviceReducer>> reducers_;
  std::map<std::string, std::weak_ptr<FutureState>> inflight_;  // all_reduce on CUDA tensors, by operation name
  std::vector<std::function<bool()>> pending_;                   // step functions; true = finished
};

struct GroupParts {
  std::shared_ptr<RpcCore> rpc;
  std::shared_ptr<GroupService> service;
  std::shared_ptr<GroupInfo> info;
  std::shared_ptr<DeviceReducerSet> reducers;
  std::shared_ptr<void> keep;
};
GroupParts groupPartsOf(const py::handle& pyGroup);
py::object makeOwnGroup(const std::string& groupName);
std::shared_ptr<RpcCore> rpcCoreOf(const py::handle& pyRpc);

}  // namespace mbh
