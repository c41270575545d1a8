// gemm_tc.cu -- tcgen05 (5th-gen tensor core) implicit-GEMM path.  [stub: filled in next]
#include "common.cuh"
namespace spc {
bool tc_supported(const spc_conv_desc*, int) { return false; }
size_t tc_workspace_bytes(const spc_conv_desc*, int) { return 0; }
int tc_conv_fwd(const spc_conv_desc*, const void*, const void*, const void*, void*, void*, size_t, cudaStream_t) { return SPC_EUNSUPPORTED; }
int tc_conv_dgrad(const spc_conv_desc*, const void*, const void*, void*, void*, size_t, cudaStream_t) { return SPC_EUNSUPPORTED; }
int tc_conv_wgrad(const spc_conv_desc*, const void*, const void*, float*, int, void*, size_t, cudaStream_t) { return SPC_EUNSUPPORTED; }
}  // namespace spc
