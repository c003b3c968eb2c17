// tc_kernels.cu -- tcgen05 / TMA kernels (placeholder until the SIMT path is validated).
#include "engine.cuh"
namespace rvsr {
bool tc_conv_supported(const ConvOp &) { return false; }
int launch_conv_tc(const ConvOp &, cudaStream_t) { set_error("tc conv not built"); return RVSR_E_UNSUPPORTED; }
size_t tc_conv_weight_bytes(int, int, int) { return 0; }
int pack_weight_tc(const float *, void *, int, int, int, int, cudaStream_t) { return RVSR_OK; }
bool tc_dcn_supported(const DcnOp &) { return false; }
int launch_dcn_tc(const DcnOp &, cudaStream_t) { set_error("tc dcn not built"); return RVSR_E_UNSUPPORTED; }
size_t tc_dcn_weight_bytes(int, int, int) { return 0; }
int pack_weight_dcn_tc(const float *, void *, int, int, int, cudaStream_t) { return RVSR_OK; }
}  // namespace rvsr
