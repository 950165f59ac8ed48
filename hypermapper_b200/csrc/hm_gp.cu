// K2 placeholder (implemented next): GP posterior mean / variance
#include "hm_common.cuh"
extern "C" int hm_gp_create(int, int, const double*, const double*, const double*, const double*, double, double, double,
                            double, int, hm_gp_t** out) {
    if (out) *out = nullptr;
    hm_set_error("hm_gp_create: not implemented yet");
    return HM_E_UNSUPPORTED;
}
extern "C" void hm_gp_destroy(hm_gp_t*) {}
extern "C" int hm_gp_mean_var(const hm_gp_t*, const double*, int64_t, int64_t, double*, double*, void*) {
    hm_set_error("hm_gp_mean_var: not implemented yet");
    return HM_E_UNSUPPORTED;
}
