// host-buffer entry points placeholder (implemented next)
#include "hm_common.cuh"
extern "C" int hm_host_ctx_create(int64_t, int, hm_host_ctx_t** out) {
    if (out) *out = nullptr;
    hm_set_error("hm_host_ctx_create: not implemented yet");
    return HM_E_UNSUPPORTED;
}
extern "C" void hm_host_ctx_destroy(hm_host_ctx_t*) {}
extern "C" int hm_rf_score_host(hm_host_ctx_t*, const hm_forest_t* const*, int, const float*, int64_t, int64_t,
                                const hm_acq_params_t*, double*, int, double*, int64_t*) {
    hm_set_error("hm_rf_score_host: not implemented yet");
    return HM_E_UNSUPPORTED;
}
