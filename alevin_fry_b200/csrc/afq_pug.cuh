// afq_pug.cuh — parsimony (PUG) and EM kernels. (stub until the kernels land)
#pragma once
#include <functional>
#include <string>
#include "../../include/afq.h"
#include "afq_kernels.cuh"

namespace afq {
struct PugWork { void release() {} };
inline int pug_em_setup(int, std::string&) { return AFQ_OK; }
inline int run_pug_em_pipeline(const afq_config&, int, int, KArgs&, PugWork&, const afq_batch&,
                               cudaStream_t, std::function<void(int)>, std::string& err) {
  err = "resolution not yet implemented on the CUDA path";
  return AFQ_ERR_UNSUPPORTED;
}
}  // namespace afq
