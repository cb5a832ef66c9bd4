// Learner-side ops next to the hot paths: V-trace targets and the uint8 observation
// normalisation as one kernel launch each, behind torch tensors.  CUDA only: there is no CPU fallback.
#include "common.h"

#include <optional>

namespace mbh {

namespace {

torch::Tensor f32Contig(const torch::Tensor& t, const char* what, int device) {
  if (!t.is_cuda() || t.get_device() != device)
    throw std::runtime_error(std::string("moolib_b200.vtrace: ") + what + " must be a CUDA tensor on the same device");
  if (t.scalar_type() != torch::kFloat32) throw std::runtime_error(std::string("moolib_b200.vtrace: ") + what + " must be float32");
  return t.contiguous();
}

// reference: from_importance_weights, examples/common/vtrace.py:156-242 -> (vs, pg_advantages)
py::tuple vtraceFromImportanceWeights(const torch::Tensor& logRhos, const torch::Tensor& discounts, const torch::Tensor& rewards,
                                      const torch::Tensor& values, const torch::Tensor& bootstrapValue,
                                      std::optional<double> clipRho, std::optional<double> clipPgRho) {
  if (!logRhos.is_cuda()) throw std::runtime_error("moolib_b200.vtrace: the kernel runs on CUDA tensors (no CPU fallback)");
  const int dev = logRhos.get_device();
  torch::NoGradGuard ng;
  torch::Tensor lr = f32Contig(logRhos, "log_rhos", dev), d = f32Contig(discounts, "discounts", dev),
                r = f32Contig(rewards, "rewards", dev), v = f32Contig(values, "values", dev),
                b = f32Contig(bootstrapValue, "bootstrap_value", dev);
  if (lr.dim() < 1 || d.sizes() != lr.sizes() || r.sizes() != lr.sizes() || v.sizes() != lr.sizes())
    throw std::runtime_error("moolib_b200.vtrace: log_rhos, discounts, rewards and values must have the same [T, B, ...] shape");
  const int64_t T = lr.size(0);
  const int64_t B = T > 0 ? lr.numel() / T : 0;
  if (b.numel() != B) throw std::runtime_error("moolib_b200.vtrace: bootstrap_value must have the shape of one time step");
  torch::Tensor vs = torch::empty_like(lr), pg = torch::empty_like(lr);
  c10::cuda::CUDAGuard g(dev);
  launch_counter() += (uint64_t)check(
      mb_vtrace_f32(lr.data_ptr<float>(), d.data_ptr<float>(), r.data_ptr<float>(), v.data_ptr<float>(), b.data_ptr<float>(),
                    clipRho ? 1 : 0, clipRho ? (float)*clipRho : 0.f, clipPgRho ? 1 : 0, clipPgRho ? (float)*clipPgRho : 0.f,
                    (uint64_t)T, (uint64_t)B, vs.data_ptr<float>(), pg.data_ptr<float>(), current_stream(dev)),
      "vtrace");
  return py::make_tuple(to_python(vs), to_python(pg));
}

// reference: `x.float() / 255.0`, examples/atari/models.py:94
torch::Tensor u8ToFloat(const torch::Tensor& x, double scale) {
  if (!x.is_cuda()) throw std::runtime_error("moolib_b200.u8_to_float: the kernel runs on CUDA tensors (no CPU fallback)");
  if (x.scalar_type() != torch::kUInt8) throw std::runtime_error("moolib_b200.u8_to_float: expected a uint8 tensor");
  torch::NoGradGuard ng;
  torch::Tensor s = x.contiguous();
  torch::Tensor out = torch::empty(s.sizes(), s.options().dtype(torch::kFloat32));
  c10::cuda::CUDAGuard g(x.get_device());
  launch_counter() += (uint64_t)check(mb_u8_to_f32(s.data_ptr<uint8_t>(), out.data_ptr<float>(), (uint64_t)s.numel(), (float)scale,
                                                   current_stream(x.get_device())),
                                      "u8_to_float");
  return out;
}

}  // namespace

void bind_learner_ops(py::module_& m) {
  m.def("vtrace_from_importance_weights", &vtraceFromImportanceWeights, py::arg("log_rhos"), py::arg("discounts"),
        py::arg("rewards"), py::arg("values"), py::arg("bootstrap_value"), py::arg("clip_rho_threshold") = 1.0,
        py::arg("clip_pg_rho_threshold") = 1.0,
        "V-trace targets (vs, pg_advantages) from log importance weights in one kernel launch "
        "(examples/common/vtrace.py:156 from_importance_weights)");
  m.def("u8_to_float", &u8ToFloat, py::arg("x"), py::arg("scale") = (double)(1.0f / 255.0f),
        "x.float() * scale for uint8 observations in one pass (examples/atari/models.py:94 `x.float() / 255.0`)");
}

}  // namespace mbh
