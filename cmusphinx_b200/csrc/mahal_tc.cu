// mahal_tc.cu -- placeholder (filled in by the tcgen05 kernel).
#include "gmm_dev.cuh"
namespace b200 {
struct TcPlan { int dummy; };
bool tc_shape_supported(const GmmDev &) { return false; }
TcPlan *tc_plan_create(const GmmDev &, const float *, const float *, const float *, const uint8_t *, int) { return nullptr; }
void tc_plan_free(TcPlan *p) { delete p; }
int tc_score(TcPlan *, const GmmDev &, const float *, int, int16_t *, cudaStream_t, cudaEvent_t *) {
    set_error("tensor-core path not built");
    return B200_ERR_UNSUP;
}
}  // namespace b200
