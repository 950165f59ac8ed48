// Host-buffer entry point: what a reference-side binding calls when the candidates live in host memory
// (numpy arrays).  Candidates are streamed H2D in column-major chunks on two streams so that the copy of chunk
// c+1 overlaps the fused forest/acquisition kernel of chunk c; scores stream back D2H the same way; the stable
// top-k runs once over the device-resident score vector.
#include <algorithm>
#include <new>
#include "hm_common.cuh"

// Error paths: the two non-blocking streams may still have asynchronous D2H copies in flight into CALLER-owned host
// buffers (a numpy array that can be freed as soon as the binding raises), so both streams are drained before any
// error code is returned.
#define HM_HOST_FAIL(c, code)                  \
    do {                                       \
        cudaStreamSynchronize((c)->st[0]);     \
        cudaStreamSynchronize((c)->st[1]);     \
        return (code);                         \
    } while (0)
#define HM_HOST_TRY(c, x)                                               \
    do {                                                                \
        cudaError_t e__ = (x);                                          \
        if (e__ != cudaSuccess) {                                       \
            hm_set_error("%s: %s", #x, cudaGetErrorString(e__));        \
            HM_HOST_FAIL(c, (int)e__);                                  \
        }                                                               \
    } while (0)

struct hm_host_ctx {
    int64_t chunk;          // candidates per chunk
    int max_features;
    cudaStream_t st[2];
    cudaEvent_t ev[2];
    float* d_x[2];          // [max_features][chunk]
    double* d_scores;       // [cap_scores]
    int64_t cap_scores;
    void* d_topk_ws;
    int64_t topk_ws_bytes;
    double* d_topk_vals;
    int64_t* d_topk_idx;
    double* d_raw[2];       // [chunk][raw_params] staging of raw parameter values (hm_rf_score_raw_host), lazily sized
    int raw_params;
    int32_t* d_invalid;
    double* d_mv[2];        // [2 M][chunk] posterior mean / variance of a chunk (hm_gp_score_host), lazily sized
    int mv_objectives;
    uint64_t* d_codes[2];   // [chunk] packed codes of a chunk (hm_rf_score_packed_host), lazily allocated
};

extern "C" int hm_host_ctx_create(int64_t max_candidates_per_chunk, int max_features, hm_host_ctx_t** out) {
    HM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HM_REQUIRE(max_candidates_per_chunk > 0 && max_features > 0, "chunk/features > 0");
    hm_host_ctx* c = new (std::nothrow) hm_host_ctx();
    if (!c) return HM_E_NOMEM;
    memset(c, 0, sizeof(*c));
    c->chunk = (max_candidates_per_chunk + 1023) / 1024 * 1024;
    c->max_features = max_features;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_x[i], (size_t)c->chunk * max_features * sizeof(float));
    }
    c->topk_ws_bytes = hm_topk_workspace_bytes(1024);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_topk_ws, (size_t)c->topk_ws_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_topk_vals, 1024 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_topk_idx, 1024 * sizeof(int64_t));
    if (e != cudaSuccess) {
        hm_set_error("hm_host_ctx_create: %s", cudaGetErrorString(e));
        hm_host_ctx_destroy(c);
        return (int)e;
    }
    *out = c;
    return 0;
}

extern "C" void hm_host_ctx_destroy(hm_host_ctx_t* c) {
    if (!c) return;
    for (int i = 0; i < 2; ++i) {
        if (c->d_x[i]) cudaFree(c->d_x[i]);
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
        if (c->st[i]) cudaStreamDestroy(c->st[i]);
    }
    for (int i = 0; i < 2; ++i)
        if (c->d_raw[i]) cudaFree(c->d_raw[i]);
    for (int i = 0; i < 2; ++i)
        if (c->d_mv[i]) cudaFree(c->d_mv[i]);
    for (int i = 0; i < 2; ++i)
        if (c->d_codes[i]) cudaFree(c->d_codes[i]);
    if (c->d_invalid) cudaFree(c->d_invalid);
    if (c->d_scores) cudaFree(c->d_scores);
    if (c->d_topk_ws) cudaFree(c->d_topk_ws);
    if (c->d_topk_vals) cudaFree(c->d_topk_vals);
    if (c->d_topk_idx) cudaFree(c->d_topk_idx);
    delete c;
}

extern "C" int hm_rf_score_host(hm_host_ctx_t* c, const hm_forest_t* const* forests, int M, const float* X_host,
                                int64_t ldx, int64_t N, const hm_acq_params_t* acq, double* score_out_host, int k,
                                double* topk_vals_host, int64_t* topk_idx_host) {
    HM_REQUIRE(c != nullptr && forests != nullptr && acq != nullptr, "ctx/forests/acq");
    HM_REQUIRE(N >= 0 && ldx >= N, "ldx >= N >= 0");
    HM_REQUIRE(k >= 0 && k <= 1024, "0 <= k <= 1024");
    HM_REQUIRE(k == 0 || (topk_vals_host && topk_idx_host), "top-k outputs");
    const int D = hm_forest_n_features(forests[0]);
    HM_REQUIRE(D > 0 && D <= c->max_features, "n_features exceeds the context's max_features");
    if (N > c->cap_scores) {
        if (c->d_scores) cudaFree(c->d_scores);
        c->d_scores = nullptr;
        c->cap_scores = 0;
        HM_CUDA_TRY(cudaMalloc((void**)&c->d_scores, (size_t)N * sizeof(double)));
        c->cap_scores = N;
    }
    int n_chunks = 0;
    for (int64_t lo = 0; lo < N; lo += c->chunk, ++n_chunks) {
        const int b = n_chunks & 1;
        const int64_t n = std::min<int64_t>(c->chunk, N - lo);
        cudaStream_t st = c->st[b];
        // [D][n] slab of the column-major host matrix -> dense [D][n] device buffer
        HM_HOST_TRY(c, cudaMemcpy2DAsync(c->d_x[b], (size_t)n * sizeof(float), X_host + lo, (size_t)ldx * sizeof(float),
                                      (size_t)n * sizeof(float), (size_t)D, cudaMemcpyHostToDevice, st));
        int rc = hm_rf_eval(forests, M, c->d_x[b], n, n, nullptr, nullptr, nullptr, nullptr, acq, nullptr,
                            c->d_scores + lo, (void*)st);
        if (rc) HM_HOST_FAIL(c, rc);
        if (score_out_host)
            HM_HOST_TRY(c, cudaMemcpyAsync(score_out_host + lo, c->d_scores + lo, (size_t)n * sizeof(double),
                                        cudaMemcpyDeviceToHost, st));
    }
    if (k > 0) {
        // stream 0 waits for stream 1's kernels, then selects over the whole score vector
        if (n_chunks > 1) {
            HM_HOST_TRY(c, cudaEventRecord(c->ev[1], c->st[1]));
            HM_HOST_TRY(c, cudaStreamWaitEvent(c->st[0], c->ev[1], 0));
        }
        int rc = hm_topk(c->d_scores, N, 0, k, c->d_topk_vals, c->d_topk_idx, c->d_topk_ws, (void*)c->st[0]);
        if (rc) HM_HOST_FAIL(c, rc);
        HM_HOST_TRY(c, cudaMemcpyAsync(topk_vals_host, c->d_topk_vals, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost,
                                    c->st[0]));
        HM_HOST_TRY(c, cudaMemcpyAsync(topk_idx_host, c->d_topk_idx, (size_t)k * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                    c->st[0]));
    }
    HM_HOST_TRY(c, cudaStreamSynchronize(c->st[0]));
    HM_HOST_TRY(c, cudaStreamSynchronize(c->st[1]));
    return 0;
}


// Packed candidates (all-discrete spaces, hm_pack.cuh): 8 bytes per candidate over PCIe instead of 4 * D_enc.  Per chunk
// H2D of the codes -> fused forest + acquisition kernel on the codes -> D2H of the scores, double-buffered.
extern "C" int hm_rf_score_packed_host(hm_host_ctx_t* c, const hm_forest_t* const* forests, int M,
                                       const uint64_t* codes_host, int64_t N, const hm_acq_params_t* acq,
                                       double* score_out_host, int k, double* topk_vals_host, int64_t* topk_idx_host) {
    HM_REQUIRE(c != nullptr && forests != nullptr && acq != nullptr, "ctx/forests/acq");
    HM_REQUIRE(N >= 0 && (N == 0 || codes_host != nullptr), "codes_host");
    HM_REQUIRE(k >= 0 && k <= 1024, "0 <= k <= 1024");
    HM_REQUIRE(k == 0 || (topk_vals_host && topk_idx_host), "top-k outputs");
    for (int i = 0; i < 2; ++i)
        if (!c->d_codes[i]) HM_CUDA_TRY(cudaMalloc((void**)&c->d_codes[i], (size_t)c->chunk * sizeof(uint64_t)));
    if (N > c->cap_scores) {
        if (c->d_scores) cudaFree(c->d_scores);
        c->d_scores = nullptr;
        c->cap_scores = 0;
        HM_CUDA_TRY(cudaMalloc((void**)&c->d_scores, (size_t)N * sizeof(double)));
        c->cap_scores = N;
    }
    int n_chunks = 0;
    for (int64_t lo = 0; lo < N; lo += c->chunk, ++n_chunks) {
        const int b = n_chunks & 1;
        const int64_t n = std::min<int64_t>(c->chunk, N - lo);
        cudaStream_t st = c->st[b];
        HM_HOST_TRY(c, cudaMemcpyAsync(c->d_codes[b], codes_host + lo, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        int rc = hm_rf_eval_packed(forests, M, c->d_codes[b], n, nullptr, nullptr, nullptr, nullptr, acq, nullptr,
                                   c->d_scores + lo, (void*)st);
        if (rc) HM_HOST_FAIL(c, rc);
        if (score_out_host)
            HM_HOST_TRY(c, cudaMemcpyAsync(score_out_host + lo, c->d_scores + lo, (size_t)n * sizeof(double),
                                        cudaMemcpyDeviceToHost, st));
    }
    if (k > 0) {
        if (n_chunks > 1) {
            HM_HOST_TRY(c, cudaEventRecord(c->ev[1], c->st[1]));
            HM_HOST_TRY(c, cudaStreamWaitEvent(c->st[0], c->ev[1], 0));
        }
        int rc = hm_topk(c->d_scores, N, 0, k, c->d_topk_vals, c->d_topk_idx, c->d_topk_ws, (void*)c->st[0]);
        if (rc) HM_HOST_FAIL(c, rc);
        HM_HOST_TRY(c, cudaMemcpyAsync(topk_vals_host, c->d_topk_vals, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost,
                                    c->st[0]));
        HM_HOST_TRY(c, cudaMemcpyAsync(topk_idx_host, c->d_topk_idx, (size_t)k * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                    c->st[0]));
    }
    HM_HOST_TRY(c, cudaStreamSynchronize(c->st[0]));
    HM_HOST_TRY(c, cudaStreamSynchronize(c->st[1]));
    return 0;
}


// The same from RAW parameter values (what the reference API is handed): per chunk  H2D of the raw rows -> hm_encode
// (K5) -> fused forest/acquisition kernel -> D2H of the scores, double-buffered across the two streams.
extern "C" int hm_rf_score_raw_host(hm_host_ctx_t* c, const hm_space_t* space, const hm_forest_t* const* forests, int M,
                                    const double* raw_host, int64_t N, const hm_acq_params_t* acq,
                                    double* score_out_host, int k, double* topk_vals_host, int64_t* topk_idx_host,
                                    int32_t* n_invalid_out_host) {
    HM_REQUIRE(c != nullptr && space != nullptr && forests != nullptr && acq != nullptr, "ctx/space/forests/acq");
    HM_REQUIRE(N >= 0 && (N == 0 || raw_host != nullptr), "raw_host");
    HM_REQUIRE(k >= 0 && k <= 1024, "0 <= k <= 1024");
    HM_REQUIRE(k == 0 || (topk_vals_host && topk_idx_host), "top-k outputs");
    const int D = hm_forest_n_features(forests[0]);
    const int Dr = hm_space_n_params(space);
    HM_REQUIRE(D > 0 && D <= c->max_features && D == hm_space_n_encoded(space),
               "encoded width of the space must equal the forests' n_features (and fit the context)");
    if (Dr > c->raw_params) {
        for (int i = 0; i < 2; ++i) {
            if (c->d_raw[i]) cudaFree(c->d_raw[i]);
            c->d_raw[i] = nullptr;
        }
        c->raw_params = 0;
        for (int i = 0; i < 2; ++i) HM_CUDA_TRY(cudaMalloc((void**)&c->d_raw[i], (size_t)c->chunk * Dr * sizeof(double)));
        c->raw_params = Dr;
    }
    if (!c->d_invalid) HM_CUDA_TRY(cudaMalloc((void**)&c->d_invalid, sizeof(int32_t)));
    if (N > c->cap_scores) {
        if (c->d_scores) cudaFree(c->d_scores);
        c->d_scores = nullptr;
        c->cap_scores = 0;
        HM_CUDA_TRY(cudaMalloc((void**)&c->d_scores, (size_t)N * sizeof(double)));
        c->cap_scores = N;
    }
    HM_CUDA_TRY(cudaMemsetAsync(c->d_invalid, 0, sizeof(int32_t), c->st[0]));
    HM_CUDA_TRY(cudaEventRecord(c->ev[0], c->st[0]));
    HM_CUDA_TRY(cudaStreamWaitEvent(c->st[1], c->ev[0], 0));
    int n_chunks = 0;
    for (int64_t lo = 0; lo < N; lo += c->chunk, ++n_chunks) {
        const int b = n_chunks & 1;
        const int64_t n = std::min<int64_t>(c->chunk, N - lo);
        cudaStream_t st = c->st[b];
        HM_HOST_TRY(c, cudaMemcpyAsync(c->d_raw[b], raw_host + (size_t)lo * Dr, (size_t)n * Dr * sizeof(double),
                                    cudaMemcpyHostToDevice, st));
        int rc = hm_encode(space, c->d_raw[b], 0, Dr, n, c->d_x[b], n, nullptr, 0, c->d_invalid, (void*)st);
        if (rc) HM_HOST_FAIL(c, rc);
        rc = hm_rf_eval(forests, M, c->d_x[b], n, n, nullptr, nullptr, nullptr, nullptr, acq, nullptr,
                        c->d_scores + lo, (void*)st);
        if (rc) HM_HOST_FAIL(c, rc);
        if (score_out_host)
            HM_HOST_TRY(c, cudaMemcpyAsync(score_out_host + lo, c->d_scores + lo, (size_t)n * sizeof(double),
                                        cudaMemcpyDeviceToHost, st));
    }
    if (n_chunks > 1) {
        HM_HOST_TRY(c, cudaEventRecord(c->ev[1], c->st[1]));
        HM_HOST_TRY(c, cudaStreamWaitEvent(c->st[0], c->ev[1], 0));
    }
    if (k > 0) {
        int rc = hm_topk(c->d_scores, N, 0, k, c->d_topk_vals, c->d_topk_idx, c->d_topk_ws, (void*)c->st[0]);
        if (rc) HM_HOST_FAIL(c, rc);
        HM_HOST_TRY(c, cudaMemcpyAsync(topk_vals_host, c->d_topk_vals, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost,
                                    c->st[0]));
        HM_HOST_TRY(c, cudaMemcpyAsync(topk_idx_host, c->d_topk_idx, (size_t)k * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                    c->st[0]));
    }
    int32_t bad = 0;
    HM_HOST_TRY(c, cudaMemcpyAsync(&bad, c->d_invalid, sizeof(int32_t), cudaMemcpyDeviceToHost, c->st[0]));
    HM_HOST_TRY(c, cudaStreamSynchronize(c->st[0]));
    HM_HOST_TRY(c, cudaStreamSynchronize(c->st[1]));
    if (n_invalid_out_host) *n_invalid_out_host = bad;
    return 0;
}


// GP path from host buffers: per chunk  H2D of the fp64 column-major candidates -> hm_gp_mean_var per objective ->
// hm_acq_score -> D2H of the scores, double-buffered across the two streams; stable top-k at the end.
extern "C" int hm_gp_score_host(hm_host_ctx_t* c, const hm_gp_t* const* gps, int M, int n_features, const double* X_host,
                                int64_t ldx, int64_t N, const hm_acq_params_t* acq, double* score_out_host, int k,
                                double* topk_vals_host, int64_t* topk_idx_host) {
    HM_REQUIRE(c != nullptr && gps != nullptr && acq != nullptr, "ctx/gps/acq");
    HM_REQUIRE(M >= 1 && M <= HM_MAX_OBJECTIVES && n_features >= 1, "1 <= M <= 8, n_features >= 1");
    HM_REQUIRE(N >= 0 && ldx >= N && (N == 0 || X_host != nullptr), "ldx >= N >= 0");
    HM_REQUIRE(k >= 0 && k <= 1024, "0 <= k <= 1024");
    HM_REQUIRE(k == 0 || (topk_vals_host && topk_idx_host), "top-k outputs");
    const int D = n_features;
    if (D > c->raw_params) {
        for (int i = 0; i < 2; ++i) {
            if (c->d_raw[i]) cudaFree(c->d_raw[i]);
            c->d_raw[i] = nullptr;
        }
        c->raw_params = 0;
        for (int i = 0; i < 2; ++i) HM_CUDA_TRY(cudaMalloc((void**)&c->d_raw[i], (size_t)c->chunk * D * sizeof(double)));
        c->raw_params = D;
    }
    if (M > c->mv_objectives) {
        for (int i = 0; i < 2; ++i) {
            if (c->d_mv[i]) cudaFree(c->d_mv[i]);
            c->d_mv[i] = nullptr;
        }
        c->mv_objectives = 0;
        for (int i = 0; i < 2; ++i) HM_CUDA_TRY(cudaMalloc((void**)&c->d_mv[i], (size_t)2 * M * c->chunk * sizeof(double)));
        c->mv_objectives = M;
    }
    if (N > c->cap_scores) {
        if (c->d_scores) cudaFree(c->d_scores);
        c->d_scores = nullptr;
        c->cap_scores = 0;
        HM_CUDA_TRY(cudaMalloc((void**)&c->d_scores, (size_t)N * sizeof(double)));
        c->cap_scores = N;
    }
    int n_chunks = 0;
    for (int64_t lo = 0; lo < N; lo += c->chunk, ++n_chunks) {
        const int b = n_chunks & 1;
        const int64_t n = std::min<int64_t>(c->chunk, N - lo);
        cudaStream_t st = c->st[b];
        HM_HOST_TRY(c, cudaMemcpy2DAsync(c->d_raw[b], (size_t)n * sizeof(double), X_host + lo, (size_t)ldx * sizeof(double),
                                      (size_t)n * sizeof(double), (size_t)D, cudaMemcpyHostToDevice, st));
        double* mean = c->d_mv[b];
        double* var = c->d_mv[b] + (size_t)M * c->chunk;
        for (int m = 0; m < M; ++m) {
            int rc = hm_gp_mean_var(gps[m], c->d_raw[b], n, n, mean + (size_t)m * c->chunk, var + (size_t)m * c->chunk,
                                    (void*)st);
            if (rc) HM_HOST_FAIL(c, rc);
        }
        int rc = hm_acq_score(acq, mean, var, c->chunk, nullptr, n, c->d_scores + lo, (void*)st);
        if (rc) HM_HOST_FAIL(c, rc);
        if (score_out_host)
            HM_HOST_TRY(c, cudaMemcpyAsync(score_out_host + lo, c->d_scores + lo, (size_t)n * sizeof(double),
                                        cudaMemcpyDeviceToHost, st));
    }
    if (k > 0) {
        if (n_chunks > 1) {
            HM_HOST_TRY(c, cudaEventRecord(c->ev[1], c->st[1]));
            HM_HOST_TRY(c, cudaStreamWaitEvent(c->st[0], c->ev[1], 0));
        }
        int rc = hm_topk(c->d_scores, N, 0, k, c->d_topk_vals, c->d_topk_idx, c->d_topk_ws, (void*)c->st[0]);
        if (rc) HM_HOST_FAIL(c, rc);
        HM_HOST_TRY(c, cudaMemcpyAsync(topk_vals_host, c->d_topk_vals, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost,
                                    c->st[0]));
        HM_HOST_TRY(c, cudaMemcpyAsync(topk_idx_host, c->d_topk_idx, (size_t)k * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                    c->st[0]));
    }
    HM_HOST_TRY(c, cudaStreamSynchronize(c->st[0]));
    HM_HOST_TRY(c, cudaStreamSynchronize(c->st[1]));
    return 0;
}
