// K4 placeholder (implemented next)
#include "hm_common.cuh"
extern "C" int64_t hm_pareto_workspace_bytes(int64_t, int) { return 256; }
extern "C" int hm_pareto_filter(const double*, int64_t, int, uint8_t*, void*, int64_t*, void*) {
    hm_set_error("hm_pareto_filter: not implemented yet");
    return HM_E_UNSUPPORTED;
}
extern "C" int hm_pareto_pass1(const double*, int64_t, int, uint8_t*, void*, int64_t*, void*) {
    hm_set_error("hm_pareto_pass1: not implemented yet");
    return HM_E_UNSUPPORTED;
}
