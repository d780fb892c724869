// hd_tc.cu - tensor-core engines (placeholder until the tcgen05 kernels land)
#include "hd_common.cuh"
namespace hd {
bool tc_available() { return false; }
int tc_gcl(const FwdCtx&, int, float*, const float*, const float*, int) {
  set_error("tensor-core engine not built");
  return HD_E_UNSUPPORTED;
}
int tc_equiv(const FwdCtx&, int, const float*, const float*, const float*, float*, int) {
  set_error("tensor-core engine not built");
  return HD_E_UNSUPPORTED;
}
int tc_edge_only(const FwdCtx&, int, const float*, const float*, int) {
  set_error("tensor-core engine not built");
  return HD_E_UNSUPPORTED;
}
}  // namespace hd
