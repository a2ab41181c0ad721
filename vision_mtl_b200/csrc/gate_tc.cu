// placeholder until the tcgen05 kernels land
#include "gate_internal.cuh"
namespace vmtl {
int gate_tc_fwd_gemm(const float*, const float*, const float*, int64_t, int, int, int, float*, float*, int, int*, cudaStream_t) { return VMTL_EUNSUPPORTED; }
int gate_tc_fwd_eval(const float*, const float*, const float*, const float*, const float*, const float*, int64_t, int, int, int, float*, cudaStream_t) { return VMTL_EUNSUPPORTED; }
int gate_tc_bwd_gemm(const float*, const float*, const float*, const float*, const float*, const GateWs&, const float*, int64_t, int, int, int, float*, float*, int, int*, float*, cudaStream_t) { return VMTL_EUNSUPPORTED; }
}
