// placeholder until the table-driven kernel lands (next commit)
#include "common.cuh"
extern "C" {
int64_t gnb_mc_workspace_bytes(int32_t, int32_t, int32_t) { return 0; }
int32_t gnb_mc_count(const float*, int32_t, int32_t, int32_t, float, void*, int64_t*, void*) {
    gnb::set_error("gnb_mc_count: not built yet");
    return GNB_ERR_UNSUPPORTED;
}
int32_t gnb_mc_emit(const float*, int32_t, int32_t, int32_t, float, const double*, int32_t, const float*, void*, float*,
                    int32_t*, float*, float*, float*, void*) {
    gnb::set_error("gnb_mc_emit: not built yet");
    return GNB_ERR_UNSUPPORTED;
}
}
