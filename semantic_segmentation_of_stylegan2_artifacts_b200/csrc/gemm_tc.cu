// tcgen05 GEMM path (placeholder until the TMA/TMEM kernel lands): reports "pattern unsupported".
#include "common.cuh"
namespace msu {
int gemm_tc(const MsuOperand*, const MsuOperand*, const MsuEpilogue*, int64_t, int64_t, int64_t, float*, int64_t,
            cudaStream_t) {
    return 1;
}
}  // namespace msu
