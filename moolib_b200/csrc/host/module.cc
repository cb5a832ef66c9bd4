// pybind11 module `moolib_b200._C` -- same names as the reference's `moolib._C` for the hot-path classes
// (reference: src/moolib.cc:1521-2285 PYBIND11_MODULE(_C, m)).
#include "common.h"

PYBIND11_MODULE(_C, m) {
  m.doc() = "moolib_b200 host layer: moolib-compatible Batcher / Accumulator / Group / EnvPool over sm_100a kernels";
  m.attr("__c_abi_version__") = mb_version();
  mbh::bind_batcher(m);
  mbh::bind_rpc(m);
  mbh::bind_accumulator(m);
  mbh::bind_envpool(m);
  mbh::bind_learner_ops(m);
}
