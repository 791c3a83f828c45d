// mlp_tc.cu — tcgen05/TMEM tensor-core path of the fused MLP kernels (placeholder until the tcgen05 kernels land:
// every entry fails loudly; nothing falls back to another path).
#include "common.cuh"
namespace fvgn {
int tc_mlp_rows(const fvgn_mlp_t*, const float*, int, int, float*, int, int, cudaStream_t) {
  set_error("tensor-core math modes are not built into this library yet");
  return FVGN_E_UNSUPPORTED;
}
int tc_cell_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, int, cudaStream_t) {
  set_error("tensor-core math modes are not built into this library yet");
  return FVGN_E_UNSUPPORTED;
}
int tc_edge_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, int, cudaStream_t) {
  set_error("tensor-core math modes are not built into this library yet");
  return FVGN_E_UNSUPPORTED;
}
}  // namespace fvgn
