// tcgen05 / TMA GEMM -- placeholder until the kernel lands in this file.
#include "internal.h"
namespace odevit {
bool gemm_tc_supports(const GemmArgs&) { return false; }
int gemm_tc(const GemmArgs&, cudaStream_t) { return set_error(ODEVIT_ERR_UNSUPPORTED, "gemm_tc: not built"); }
}  // namespace odevit
