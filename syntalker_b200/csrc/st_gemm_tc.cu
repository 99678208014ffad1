// tcgen05 split-fp16 GEMM engine (placeholder until the kernel lands; the dispatcher then never selects it).
#include "st_internal.cuh"

namespace st {
bool tc_supported(const GemmP&) { return false; }
int gemm_tc(const GemmP&, cudaStream_t) {
  set_error("tcgen05 engine not built");
  return ST_EUNSUPPORTED;
}
}  // namespace st
