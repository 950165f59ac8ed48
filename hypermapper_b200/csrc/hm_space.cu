// K5 / K6 / K7: the search-space side of the hot path on the device.
//   K5 hm_encode      raw parameter values -> the encoded column-major matrices the surrogates read
//                     (models.preprocess_data_array, models.py:941-984), bit-exact
//   K6 hm_sample_pool the two random pools of local_search (space.get_random_configuration, space.py:1263-1326;
//                     per-parameter randomly_select / randomly_select_uniform) drawn from a counter-based generator
//                     (Philox4x32-10 keyed by seed, counter = candidate index x parameter), encoded in the same pass.
//                     Any candidate can be re-generated from (seed, index): the pool itself never has to exist.
//   K7 hm_prior_prob  compute_probability_from_prior (prior_optimization.py:65-95) and the piBO weighting
//                     (pibo.py:101-113)
// One thread = one candidate; the parameter descriptors and their tables sit in shared memory; every output column is
// written 32 candidates x 4/8 bytes per warp (coalesced).  HBM-bound: 8 D bytes in (K5/K7), 4 D_enc (+ 8 D_enc) out.
#include <algorithm>
#include <math.h>
#include <new>
#include <vector>
#include "hm_acq.cuh"
#include "hm_pack.cuh"

#define HM_SPACE_MAX_PARAMS 256
#define HM_SPACE_SMEM_TABLES 4096          // doubles of table data cached in shared memory (32 KB); larger stay in L2

static int hm_check_desc(const hm_param_desc_t& p, long long n_tables, int j) {
    (void)j;
    HM_REQUIRE(p.kind >= HM_P_REAL && p.kind <= HM_P_CATEGORICAL, "parameter kind");
    HM_REQUIRE(p.prior_kind >= HM_PRIOR_BETA && p.prior_kind <= HM_PRIOR_INTERP, "prior kind");
    HM_REQUIRE(p.out_col >= 0, "out_col");
    if (p.kind == HM_P_ORDINAL) {
        HM_REQUIRE(p.n_values >= 1 && p.values_off >= 0 && p.values_off + 2ll * p.n_values <= n_tables,
                   "ordinal value table out of range");
    }
    if (p.kind == HM_P_CATEGORICAL) HM_REQUIRE(p.n_values >= 1, "categorical size");
    if (p.kind == HM_P_INTEGER) {
        // the uniform draw is a 32-bit multiply-high on the range size, the value index an int
        HM_REQUIRE(p.hi >= p.lo && p.hi - p.lo + 1.0 <= 2147483647.0, "integer parameter: at most 2^31 - 1 values");
        if (p.prior_kind == HM_PRIOR_TABLE) HM_REQUIRE(p.n_prior == (int)(p.hi - p.lo) + 1, "integer distribution length");
    }
    if (p.prior_kind == HM_PRIOR_TABLE)
        HM_REQUIRE(p.n_prior >= 1 && p.prior_off >= 0 && p.prior_off + 2ll * p.n_prior <= n_tables,
                   "prior table out of range ([probabilities | cumulative])");
    if (p.prior_kind == HM_PRIOR_INTERP)
        HM_REQUIRE(p.n_prior >= 2 && p.prior_off >= 0 && p.prior_off + 2ll * p.n_prior <= n_tables &&
                       p.n_cdf >= 2 && p.cdf_off >= 0 && p.cdf_off + 2ll * p.n_cdf <= n_tables,
                   "interpolated prior tables out of range");
    return 0;
}

extern "C" int hm_space_create(int n_params, const hm_param_desc_t* params_host, const double* tables_host,
                               int64_t n_tables, hm_space_t** out) {
    HM_REQUIRE(out != nullptr && params_host != nullptr, "out/params");
    HM_REQUIRE(n_params >= 1 && n_params <= HM_SPACE_MAX_PARAMS, "1 <= n_params <= 256");
    HM_REQUIRE(n_tables >= 0 && n_tables < (1ll << 30) && (n_tables == 0 || tables_host != nullptr), "tables");
    int d_enc = 0;
    for (int j = 0; j < n_params; ++j) {
        int rc = hm_check_desc(params_host[j], n_tables, j);
        if (rc) return rc;
        const int width = params_host[j].kind == HM_P_CATEGORICAL ? params_host[j].n_values : 1;
        d_enc = std::max(d_enc, params_host[j].out_col + width);
    }
    hm_space* S = new (std::nothrow) hm_space();
    if (!S) return HM_E_NOMEM;
    S->D = n_params;
    S->D_enc = d_enc;
    S->n_tables = n_tables;
    S->params.assign(params_host, params_host + n_params);
    if (n_tables) S->tables.assign(tables_host, tables_host + n_tables);
    S->pack.resize(n_params);
    S->pack_fast = 0;
    S->pack_bits = hm_pack_layout(params_host, n_params, S->pack.data(), &S->pack_fast);
    if (!S->pack_bits) S->pack.clear();
    S->d_pack_low = nullptr;
    S->d_params = nullptr;
    S->d_tables = nullptr;
#define HM_S_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) {                                                                             \
            hm_set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                                    \
            hm_space_destroy(S);                                                                             \
            return (int)_e;                                                                                  \
        }                                                                                                    \
    } while (0)
    HM_S_TRY(cudaMalloc((void**)&S->d_params, sizeof(hm_param_desc_t) * n_params));
    HM_S_TRY(cudaMemcpy(S->d_params, params_host, sizeof(hm_param_desc_t) * n_params, cudaMemcpyHostToDevice));
    HM_S_TRY(cudaMalloc((void**)&S->d_tables, sizeof(double) * std::max<long long>(n_tables, 1)));
    if (n_tables)
        HM_S_TRY(cudaMemcpy(S->d_tables, tables_host, sizeof(double) * n_tables, cudaMemcpyHostToDevice));
    if (S->pack_bits) {
        std::vector<int> lows(n_params);
        for (int j = 0; j < n_params; ++j) lows[j] = S->pack[j].low;
        HM_S_TRY(cudaMalloc((void**)&S->d_pack_low, sizeof(int) * n_params));
        HM_S_TRY(cudaMemcpy(S->d_pack_low, lows.data(), sizeof(int) * n_params, cudaMemcpyHostToDevice));
    }
    HM_S_TRY(cudaDeviceSynchronize());      // consumed on non-blocking caller streams (see hm_forest_create)
#undef HM_S_TRY
    *out = S;
    return 0;
}

extern "C" void hm_space_destroy(hm_space_t* S) {
    if (!S) return;
    cudaFree(S->d_params);
    cudaFree(S->d_tables);
    cudaFree(S->d_pack_low);
    delete S;
}
extern "C" int hm_space_n_params(const hm_space_t* s) { return s ? s->D : 0; }
extern "C" int hm_space_n_encoded(const hm_space_t* s) { return s ? s->D_enc : 0; }
extern "C" int hm_space_packed_bits(const hm_space_t* s) { return s ? s->pack_bits : 0; }
extern "C" int hm_space_packed_field(const hm_space_t* s, int param, int* low_bit, int* width) {
    HM_REQUIRE(s != nullptr && s->pack_bits > 0, "space is not packable");
    HM_REQUIRE(param >= 0 && param < s->D && low_bit && width, "param");
    *low_bit = s->pack[param].low;
    *width = s->pack[param].width;
    return 0;
}

// ---- Philox4x32-10: one stream per (seed, candidate index, pool kind), consumed sequentially over the parameters ------
struct HmPhilox {
    uint32_t c0, c1, c2, c3, k0, k1, o0, o1, o2, o3;
    int have;
    __device__ __forceinline__ void init(unsigned long long seed, unsigned long long index, uint32_t stream) {
        c0 = (uint32_t)index;
        c1 = (uint32_t)(index >> 32);
        c2 = stream;
        c3 = 0;
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        have = 0;
    }
    __device__ __forceinline__ void block() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t y0 = hi1 ^ x1 ^ ka, y1 = lo1, y2 = hi0 ^ x3 ^ kb, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            ka += 0x9E3779B9u;
            kb += 0xBB67AE85u;
        }
        o0 = x0; o1 = x1; o2 = x2; o3 = x3;
        ++c3;
        have = 4;
    }
    __device__ __forceinline__ uint32_t u32() {
        if (have == 0) block();
        const uint32_t v = have == 4 ? o0 : (have == 3 ? o1 : (have == 2 ? o2 : o3));
        --have;
        return v;
    }
    // uniform index in [0, n): np.random.choice(n) without probabilities
    __device__ __forceinline__ int below(int n) { return (int)__umulhi(u32(), (uint32_t)n); }
    // uniform in [0, 1) with 53 random bits
    __device__ __forceinline__ double u53() {
        const uint32_t hi = u32(), lo = u32();
        return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ double normal() {
        const double u1 = 1.0 - u53();            // (0, 1]
        const double u2 = u53();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    // Marsaglia-Tsang
    __device__ double gamma(double a) {
        double boost = 1.0;
        if (a < 1.0) {
            boost = pow(1.0 - u53(), 1.0 / a);
            a += 1.0;
        }
        const double d = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
        for (int it = 0; it < 64; ++it) {
            const double z = normal();
            double v = 1.0 + cc * z;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = 1.0 - u53();
            if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) return boost * d * v;
        }
        return boost * d;
    }
    __device__ double beta(double a, double b) {
        if (a == 1.0 && b == 1.0) return u53();
        if (b == 1.0) return pow(u53(), 1.0 / a);                 // cdf x^a
        if (a == 1.0) return 1.0 - pow(1.0 - u53(), 1.0 / b);     // cdf 1-(1-x)^b
        const double ga = gamma(a), gb = gamma(b);
        return ga / (ga + gb);
    }
};

// ---- per-parameter device functions ----------------------------------------------------------------------------
// compact copy of hm_param_desc_t in shared memory (32-bit table offsets)
struct HmP {
    int kind, n_values, out_col, flags;          // flags: 1 = z-score, 2 = discrete parameter sampled as Beta(a, b)
    int prior_kind, n_prior, n_cdf, values_off;
    int prior_off, cdf_off, pack_low, pad1;       // pack_low: bit index of the parameter's field in a packed code
    double lo, hi, mean, std, a, b, lbeta, pad2;
};
#define HM_PF_NORMALIZE 1
#define HM_PF_BETA_SAMPLED 2

// index of x in the ascending table tab[0..n): -1 when x is not one of the values (list.index raises in the reference).
// Branch-free lower bound: ceil(log2 n) dependent shared-memory loads.
__device__ __forceinline__ int hm_find_value(const double* __restrict__ tab, int n, double x) {
    int lo = 0, len = n;
    while (len > 1) {
        const int half = len >> 1;
        lo += (tab[lo + half - 1] < x) ? half : 0;
        len -= half;
    }
    return tab[lo] == x ? lo : -1;
}
// np.random.choice(n, p=...): first index whose cumulative probability exceeds u
__device__ __forceinline__ int hm_choice(const double* __restrict__ cum, int n, double u) {
    int lo = 0, len = n;
    while (len > 1) {
        const int half = len >> 1;
        lo += (cum[lo + half - 1] <= u) ? half : 0;
        len -= half;
    }
    return lo;
}
// np.interp(x, xp, fp) with xp ascending (numpy/core/src/multiarray/compiled_base.c arr_interp)
__device__ __forceinline__ double hm_interp(const double* __restrict__ xp, const double* __restrict__ fp, int n, double x) {
    if (x != x) return x;
    if (x < xp[0]) return fp[0];
    if (x > xp[n - 1]) return fp[n - 1];
    // largest j with xp[j] <= x
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (xp[hi] <= x) lo = hi;
    if (lo == n - 1) return fp[lo];
    const double x0 = xp[lo], y0 = fp[lo];
    if (x0 == x) return y0;
    const double slope = hm_sub(fp[lo + 1], y0) / hm_sub(xp[lo + 1], x0);
    return hm_add(hm_mul(slope, hm_sub(x, x0)), y0);
}
// scipy.stats.beta.pdf on [0, 1]
__device__ __forceinline__ double hm_beta_pdf(double x, double a, double b, double lbeta) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    if (x != x) return x;
    if (x < 0.0 || x > 1.0) return 0.0;
    if (x == 0.0) return a < 1.0 ? inf : (a > 1.0 ? 0.0 : exp(-lbeta));
    if (x == 1.0) return b < 1.0 ? inf : (b > 1.0 ? 0.0 : exp(-lbeta));
    return exp((a - 1.0) * log(x) + (b - 1.0) * log1p(-x) - lbeta);
}

// One parameter value drawn as space.py's randomly_select (use_priors) / randomly_select_uniform would; *idx_out = its
// value index for ordinal / categorical parameters (so that the encoder does not search for it again).
__device__ double hm_draw(const double* __restrict__ tab, const HmP& p, HmPhilox& g, bool prior, int* idx_out) {
    if (p.kind == HM_P_REAL) {
        if (!prior) return hm_add(hm_mul(g.u53(), hm_sub(p.hi, p.lo)), p.lo);                   // space.py:201-209
        if (p.prior_kind == HM_PRIOR_INTERP)                                                     // space.py:151-153
            return hm_interp(tab + p.cdf_off, tab + p.cdf_off + p.n_cdf, p.n_cdf, g.u53());
        if (p.prior_kind == HM_PRIOR_GAUSSIAN) {                                                 // space.py:167-193
            double x = p.a;
            for (int it = 0; it < 256; ++it) {
                x = p.a + p.b * g.normal();
                if (x <= p.hi && x >= p.lo) return x;
            }
            return fmin(fmax(x, p.lo), p.hi);
        }
        return hm_add(hm_mul(g.beta(p.a, p.b), hm_sub(p.hi, p.lo)), p.lo);                      // space.py:194-197
    }
    if (p.kind == HM_P_INTEGER) {
        const double size = p.hi - p.lo + 1.0;
        if (!prior) return p.lo + (double)g.below((int)size);                                    // space.py:389-395
        if (p.prior_kind == HM_PRIOR_TABLE && !(p.flags & HM_PF_BETA_SAMPLED))                   // space.py:379-382
            return p.lo + (double)hm_choice(tab + p.prior_off + p.n_prior, p.n_prior, g.u53());
        const double v = rint(g.beta(p.a, p.b) * size - 0.5 + p.lo);                             // space.py:437-453
        return fmin(fmax(v, p.lo), p.hi);
    }
    int idx;
    if (!prior) {
        idx = g.below(p.n_values);                                                               // space.py:505-511, 639-645
    } else if (p.prior_kind == HM_PRIOR_TABLE && !(p.flags & HM_PF_BETA_SAMPLED)) {
        idx = hm_choice(tab + p.prior_off + p.n_prior, p.n_prior, g.u53());                      // space.py:493-494, 634-637
    } else {
        idx = (int)rint(g.beta(p.a, p.b) * (double)p.n_values - 0.5);                            // space.py:578-595
    }
    idx = min(max(idx, 0), p.n_values - 1);
    *idx_out = idx;
    return p.kind == HM_P_ORDINAL ? tab[p.values_off + idx] : (double)idx;
}

// get_x_probability of one parameter (space.py:264-286, 412-417, 513-540, 647-660); idx = value index if known (>= 0)
__device__ double hm_x_probability(const double* __restrict__ tab, const HmP& p, double x, int idx, int* bad) {
    if (p.prior_kind == HM_PRIOR_TABLE) {
        if (idx < 0) {
            if (p.kind == HM_P_ORDINAL) idx = hm_find_value(tab + p.values_off, p.n_values, x);
            else if (p.kind == HM_P_INTEGER) idx = (x >= p.lo && x <= p.hi && x == rint(x)) ? (int)(x - p.lo) : -1;
            else idx = (x >= 0.0 && x < (double)p.n_values && x == rint(x)) ? (int)x : -1;
        }
        if (idx < 0) {
            *bad = 1;
            return __longlong_as_double(0x7ff8000000000000ll);
        }
        return tab[p.prior_off + idx];
    }
    if (p.prior_kind == HM_PRIOR_INTERP) return hm_interp(tab + p.prior_off, tab + p.prior_off + p.n_prior, p.n_prior, x);
    if (p.prior_kind == HM_PRIOR_GAUSSIAN) {
        const double y = hm_sub(x, p.a) / p.b;
        return hm_norm_pdf(y) / p.b;
    }
    const double r = hm_sub(x, p.lo) / hm_sub(p.hi, p.lo);
    double pdf = hm_beta_pdf(r, p.a, p.b, p.lbeta);
    if (p.kind == HM_P_REAL && pdf == __longlong_as_double(0x7ff0000000000000ll)) pdf = 1242.0;   // space.py:283-285
    return pdf;
}

struct HmSpaceArgs {
    const hm_param_desc_t* params;
    const double* tables;
    long long n_tables;
    int D;
    // input (when not generating)
    const double* raw;
    int raw_col_major;
    long long ldr;
    long long N;
    // generation
    unsigned long long seed;
    long long index_base;
    int use_priors;
    const long long* gather;      // hm_sample_rows: candidate indices to re-generate (N of them)
    // outputs (nullable)
    double* raw_out;              // column-major [D][ldro], or row-major [N][D] when raw_out_rows
    long long ldro;
    int raw_out_rows;
    float* x32;
    long long ld32;
    double* x64;
    long long ld64;
    double* prob;
    int n_obj;
    double w[HM_MAX_OBJECTIVES];
    int* n_invalid;
    unsigned long long* codes;    // packed 64-bit code per candidate (hm_pack.cuh), all-discrete spaces only
    const int* pack_low;          // device [D]: bit index of every parameter's field
};

// descriptors (compact form) and the first n_tab table entries into shared memory; ends with a CTA barrier
__device__ __forceinline__ void hm_space_load_shared(const HmSpaceArgs& a, HmP* sp, double* st, int n_tab) {
    for (int j = threadIdx.x; j < a.D; j += blockDim.x) {
        const hm_param_desc_t& d = a.params[j];
        HmP p;
        p.kind = d.kind;
        p.n_values = d.n_values;
        p.out_col = d.out_col;
        p.flags = (d.normalize ? HM_PF_NORMALIZE : 0) | (d.ordinal_beta ? HM_PF_BETA_SAMPLED : 0);
        p.prior_kind = d.prior_kind;
        p.n_prior = d.n_prior;
        p.n_cdf = d.n_cdf;
        p.values_off = (int)d.values_off;
        p.prior_off = (int)d.prior_off;
        p.cdf_off = (int)d.cdf_off;
        p.pack_low = a.pack_low ? a.pack_low[j] : 0;
        p.pad1 = 0;
        p.lo = d.lo; p.hi = d.hi; p.mean = d.mean; p.std = d.std; p.a = d.a; p.b = d.b; p.lbeta = d.lbeta; p.pad2 = 0.0;
        sp[j] = p;
    }
    for (int q = threadIdx.x; q < n_tab; q += blockDim.x) st[q] = __ldg(a.tables + q);
    __syncthreads();
}

#define HM_SPACE_BLOCK 256
// GENERATE: draw the raw values (K6) instead of reading them.  STAGED: the raw matrix is row-major and contiguous
// (ldr == D): a tile of 256 candidates x D values is copied into shared memory with coalesced loads and each thread
// then reads its own row (row stride D|1 doubles: conflict-free).
template <bool GENERATE, bool STAGED>
__global__ void __launch_bounds__(HM_SPACE_BLOCK) hm_space_kernel(const HmSpaceArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    HmP* sp = reinterpret_cast<HmP*>(smem);
    double* st = reinterpret_cast<double*>(smem + sizeof(HmP) * a.D);
    const int tab_s_n = (int)(a.n_tables <= HM_SPACE_SMEM_TABLES ? a.n_tables : 0);   // all tables or none
    double* tile = st + tab_s_n;
    const int Dp = a.D | 1;
    hm_space_load_shared(a, sp, st, tab_s_n);
    // the tables: shared memory when they fit, else global (generic addressing serves both)
    const double* tab = tab_s_n ? st : a.tables;

    const bool want_enc = a.x32 != nullptr || a.x64 != nullptr;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    int bad_any = 0;
    for (long long base = (long long)blockIdx.x * HM_SPACE_BLOCK; base < a.N; base += (long long)gridDim.x * HM_SPACE_BLOCK) {
        if (STAGED) {
            __syncthreads();
            const long long rows = a.N - base < HM_SPACE_BLOCK ? a.N - base : HM_SPACE_BLOCK;
            const int total = (int)rows * a.D;
            const double* src = a.raw + (size_t)base * (size_t)a.D;
            int r = threadIdx.x / a.D, c = threadIdx.x - r * a.D;
            const int dr = HM_SPACE_BLOCK / a.D, dc = HM_SPACE_BLOCK - dr * a.D;
            for (int q = threadIdx.x; q < total; q += HM_SPACE_BLOCK) {
                tile[r * Dp + c] = __ldg(src + q);
                r += dr;
                c += dc;
                if (c >= a.D) {
                    c -= a.D;
                    ++r;
                }
            }
            __syncthreads();
        }
        const long long i = base + threadIdx.x;
        if (i >= a.N) continue;
        double prob = 1.0;
        unsigned long long code = 0ull;
        HmPhilox g;
        if (GENERATE) {
            const unsigned long long gi = (unsigned long long)(a.gather ? __ldg(a.gather + i) : a.index_base + i);
            g.init(a.seed, gi, a.use_priors ? 1u : 0u);
        }
        for (int j = 0; j < a.D; ++j) {
            const HmP& p = sp[j];
            double x;
            int idx = -1;
            if (GENERATE) {
                x = hm_draw(tab, p, g, a.use_priors != 0, &idx);
            } else if (STAGED) {
                x = tile[threadIdx.x * Dp + j];
            } else {
                x = a.raw_col_major ? __ldg(a.raw + (size_t)j * (size_t)a.ldr + (size_t)i)
                                    : __ldg(a.raw + (size_t)i * (size_t)a.ldr + (size_t)j);
            }
            if (a.raw_out) {
                if (a.raw_out_rows) a.raw_out[(size_t)i * (size_t)a.D + j] = x;
                else a.raw_out[(size_t)j * (size_t)a.ldro + (size_t)i] = x;
            }
            if (!GENERATE && p.kind >= HM_P_ORDINAL && (want_enc || a.prob || a.codes)) {
                // value index of a discrete parameter: ordinal -> position in its value list, categorical -> the value
                if (p.kind == HM_P_ORDINAL) idx = hm_find_value(tab + p.values_off, p.n_values, x);
                else idx = (x >= 0.0 && x < (double)p.n_values && x == rint(x)) ? (int)x : -1;
                bad_any |= idx < 0;
            }
            if (a.codes && idx >= 0)      // ordinal: rank; categorical: one-hot over the first n - 1 categories (hm_pack.cuh)
                code |= (p.kind == HM_P_CATEGORICAL ? (idx < p.n_values - 1 ? (1ull << idx) : 0ull)
                                                    : (unsigned long long)idx) << p.pack_low;
            if (want_enc) {
                // models.py:957-984
                if (p.kind == HM_P_CATEGORICAL) {
                    float* o32 = a.x32 ? a.x32 + (size_t)p.out_col * (size_t)a.ld32 + (size_t)i : nullptr;
                    double* o64 = a.x64 ? a.x64 + (size_t)p.out_col * (size_t)a.ld64 + (size_t)i : nullptr;
                    for (int k = 0; k < p.n_values; ++k) {
                        const double v = idx < 0 ? nan : (k == idx ? 1.0 : 0.0);
                        if (o32) o32[(size_t)k * (size_t)a.ld32] = (float)v;
                        if (o64) o64[(size_t)k * (size_t)a.ld64] = v;
                    }
                } else {
                    double v;
                    if (p.kind == HM_P_ORDINAL) v = idx < 0 ? nan : tab[p.values_off + p.n_values + idx];
                    else if (p.flags & HM_PF_NORMALIZE) v = hm_sub(x, p.mean) / p.std;
                    else v = x;
                    if (a.x32) a.x32[(size_t)p.out_col * (size_t)a.ld32 + (size_t)i] = (float)v;
                    if (a.x64) a.x64[(size_t)p.out_col * (size_t)a.ld64 + (size_t)i] = v;
                }
            }
            if (a.prob) {
                // prior_optimization.py:85-93: for every parameter, for every objective: probability *= p ** w
                int bad = 0;
                const double px = hm_x_probability(tab, p, x, idx, &bad);
                bad_any |= bad;
                for (int o = 0; o < a.n_obj; ++o) prob = hm_mul(prob, a.w[o] == 1.0 ? px : pow(px, a.w[o]));
            }
        }
        if (a.prob) a.prob[i] = prob;
        if (a.codes) a.codes[i] = code;
    }
    if (a.n_invalid && bad_any) atomicAdd(a.n_invalid, 1);
}

// ---- K5 fast path: parameter-major over a staged tile --------------------------------------------------------------
// The generic kernel above interprets the descriptor once per (candidate, parameter): ~160 instructions per pair, issue
// bound at 0.4 of HBM.  Here a CTA stages 256*KC rows of the contiguous row-major raw matrix in shared memory (coalesced
// loads) and walks the PARAMETERS in the outer loop: the descriptor is decoded, the kind dispatched and the output
// column addresses formed once per parameter and thread, then applied to the thread's KC candidates (rows tid + 256 c,
// so every column store is still 32 consecutive candidates per warp); the KC value searches of an ordinal are
// independent and overlap.  Tables live in shared memory (32-bit addressing).  Same arithmetic, same outputs.
// MODE: 0 = fp32 columns, 1 = fp64 columns, 2 = both, 3 = prior probability only (hm_prior_prob)
template <int KC, int MODE, bool FULL>
__device__ __forceinline__ void hm_encode_tile_body(const HmSpaceArgs& a, const HmP* __restrict__ sp,
                                                    const double* __restrict__ st, const double* __restrict__ tile, int Dp,
                                                    long long base, int& bad_any) {
    constexpr bool W32 = MODE == 0 || MODE == 2, W64 = MODE == 1 || MODE == 2, PROB = MODE == 3;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const long long i0 = base + threadIdx.x;
    bool live[KC];
    double prob[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        live[c] = FULL || i0 + c * HM_SPACE_BLOCK < a.N;
        prob[c] = 1.0;
    }
    for (int j = 0; j < a.D; ++j) {
        const HmP& p = sp[j];
        const int kind = p.kind, nv = p.n_values;
        double x[KC];
        int idx[KC];
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            x[c] = tile[(threadIdx.x + c * HM_SPACE_BLOCK) * Dp + j];      // rows past N hold stale data; never stored
            idx[c] = -1;
        }
        if (kind == HM_P_ORDINAL) {
            const double* vt = st + p.values_off;
            int lo[KC], len = nv;
#pragma unroll
            for (int c = 0; c < KC; ++c) lo[c] = 0;
            while (len > 1) {                                               // KC independent lower-bound searches
                const int half = len >> 1;
#pragma unroll
                for (int c = 0; c < KC; ++c) lo[c] += (vt[lo[c] + half - 1] < x[c]) ? half : 0;
                len -= half;
            }
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                idx[c] = vt[lo[c]] == x[c] ? lo[c] : -1;
                bad_any |= (live[c] && idx[c] < 0);
            }
        } else if (kind == HM_P_CATEGORICAL) {
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                idx[c] = (x[c] >= 0.0 && x[c] < (double)nv && x[c] == rint(x[c])) ? (int)x[c] : -1;
                bad_any |= (live[c] && idx[c] < 0);
            }
        }
        if (!PROB) {
            // models.py:957-984
            float* o32 = W32 ? a.x32 + (size_t)p.out_col * (size_t)a.ld32 + (size_t)i0 : nullptr;
            double* o64 = W64 ? a.x64 + (size_t)p.out_col * (size_t)a.ld64 + (size_t)i0 : nullptr;
            if (kind == HM_P_CATEGORICAL) {
                for (int k = 0; k < nv; ++k) {
#pragma unroll
                    for (int c = 0; c < KC; ++c) {
                        const float v = idx[c] < 0 ? __int_as_float(0x7fc00000) : (k == idx[c] ? 1.0f : 0.0f);
                        if (FULL || live[c]) {
                            if (W32) o32[c * HM_SPACE_BLOCK] = v;
                            if (W64) o64[c * HM_SPACE_BLOCK] = (double)v;
                        }
                    }
                    if (W32) o32 += a.ld32;
                    if (W64) o64 += a.ld64;
                }
            } else {
                const double* enc = st + p.values_off + nv;
                const bool norm = (p.flags & HM_PF_NORMALIZE) != 0;
                const double mean = p.mean, sd = p.std;
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    double v;
                    if (kind == HM_P_ORDINAL) v = idx[c] < 0 ? nan : enc[idx[c] < 0 ? 0 : idx[c]];
                    else if (norm) v = hm_sub(x[c], mean) / sd;
                    else v = x[c];
                    if (FULL || live[c]) {
                        if (W32) o32[c * HM_SPACE_BLOCK] = (float)v;
                        if (W64) o64[c * HM_SPACE_BLOCK] = v;
                    }
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                int bad = 0;
                const double px = hm_x_probability(st, p, x[c], idx[c], &bad);
                bad_any |= (live[c] && bad);
                for (int o = 0; o < a.n_obj; ++o) prob[c] = hm_mul(prob[c], a.w[o] == 1.0 ? px : pow(px, a.w[o]));
            }
        }
    }
    if (PROB) {
#pragma unroll
        for (int c = 0; c < KC; ++c)
            if (live[c]) a.prob[i0 + c * HM_SPACE_BLOCK] = prob[c];
    }
}

template <int KC, int MODE>
__global__ void __launch_bounds__(HM_SPACE_BLOCK) hm_encode_tile_kernel(const HmSpaceArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    HmP* sp = reinterpret_cast<HmP*>(smem);
    double* st = reinterpret_cast<double*>(smem + sizeof(HmP) * a.D);
    const int tab_n = (int)a.n_tables;
    double* tile = st + tab_n;
    const int Dp = a.D | 1;
    constexpr int TR = HM_SPACE_BLOCK * KC;
    hm_space_load_shared(a, sp, st, tab_n);
    int bad_any = 0;
    for (long long base = (long long)blockIdx.x * TR; base < a.N; base += (long long)gridDim.x * TR) {
        __syncthreads();
        const long long rows = a.N - base < TR ? a.N - base : TR;
        {
            const int total = (int)rows * a.D;
            const double* src = a.raw + (size_t)base * (size_t)a.D;
            int r = threadIdx.x / a.D, c = threadIdx.x - r * a.D;
            const int dr = HM_SPACE_BLOCK / a.D, dc = HM_SPACE_BLOCK - dr * a.D;
#pragma unroll 4
            for (int q = threadIdx.x; q < total; q += HM_SPACE_BLOCK) {
                tile[r * Dp + c] = __ldg(src + q);
                r += dr;
                c += dc;
                if (c >= a.D) {
                    c -= a.D;
                    ++r;
                }
            }
        }
        __syncthreads();
        if (rows == TR) hm_encode_tile_body<KC, MODE, true>(a, sp, st, tile, Dp, base, bad_any);
        else hm_encode_tile_body<KC, MODE, false>(a, sp, st, tile, Dp, base, bad_any);
    }
    if (a.n_invalid && bad_any) atomicAdd(a.n_invalid, 1);
}

// launches the tile kernel when its preconditions hold; *done = 0 otherwise
static int hm_space_launch_tile(const hm_space_t* s, HmSpaceArgs& a, cudaStream_t st, int* done) {
    *done = 0;
    if (a.raw_col_major || a.ldr != s->D || s->n_tables > HM_SPACE_SMEM_TABLES || a.raw_out || a.codes) return 0;
    const bool w32 = a.x32 != nullptr, w64 = a.x64 != nullptr, wp = a.prob != nullptr;
    if (wp == (w32 || w64)) return 0;             // exactly one of {encode, prior}: what hm_encode / hm_prior_prob ask for
    const int mode = wp ? 3 : (w32 && w64 ? 2 : (w64 ? 1 : 0));
    const size_t fixed = sizeof(HmP) * s->D + sizeof(double) * (size_t)s->n_tables;
    const size_t row = sizeof(double) * (size_t)(s->D | 1) * HM_SPACE_BLOCK;
    int kc = 0;
    if (4 * row <= 48 * 1024) kc = 4;
    else if (2 * row <= 48 * 1024) kc = 2;
    else if (row <= 96 * 1024) kc = 1;
    if (!kc) return 0;
    const size_t smem = fixed + kc * row;
    const long long tiles = (a.N + (long long)HM_SPACE_BLOCK * kc - 1) / ((long long)HM_SPACE_BLOCK * kc);
#define HM_TILE_GO(K, MD)                                                                                              \
    do {                                                                                                               \
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_encode_tile_kernel<K, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                         (int)smem));                                                                  \
        int occ = 0;                                                                                                   \
        HM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hm_encode_tile_kernel<K, MD>, HM_SPACE_BLOCK,  \
                                                                  smem));                                              \
        const int grid = (int)std::min<long long>(tiles, (long long)hm_num_sms() * (occ < 1 ? 1 : occ));               \
        hm_encode_tile_kernel<K, MD><<<grid, HM_SPACE_BLOCK, smem, st>>>(a);                                           \
    } while (0)
#define HM_TILE_MODE(K)                                                                                                \
    do {                                                                                                               \
        if (mode == 0) HM_TILE_GO(K, 0);                                                                               \
        else if (mode == 1) HM_TILE_GO(K, 1);                                                                          \
        else if (mode == 2) HM_TILE_GO(K, 2);                                                                          \
        else HM_TILE_GO(K, 3);                                                                                         \
    } while (0)
    if (kc == 4) HM_TILE_MODE(4);
    else if (kc == 2) HM_TILE_MODE(2);
    else HM_TILE_MODE(1);
#undef HM_TILE_MODE
#undef HM_TILE_GO
    HM_CUDA_TRY(cudaGetLastError());
    *done = 1;
    return 0;
}

static int hm_space_launch(const hm_space_t* s, HmSpaceArgs& a, bool generate, void* stream) {
    a.params = s->d_params;
    a.tables = s->d_tables;
    a.n_tables = s->n_tables;
    a.D = s->D;
    if (a.N == 0) return 0;
    if (!generate) {
        int done = 0;
        int rc = hm_space_launch_tile(s, a, (cudaStream_t)stream, &done);
        if (rc || done) return rc;
    }
    const long long blocks = (a.N + HM_SPACE_BLOCK - 1) / HM_SPACE_BLOCK;
    const bool staged = !generate && !a.raw_col_major && a.ldr == a.D && a.D <= 32;
    const size_t smem = sizeof(HmP) * s->D +
                        sizeof(double) * (size_t)(s->n_tables <= HM_SPACE_SMEM_TABLES ? s->n_tables : 0) +
                        (staged ? sizeof(double) * (size_t)HM_SPACE_BLOCK * (size_t)(s->D | 1) : 0);
    cudaStream_t st = (cudaStream_t)stream;
    // grid-stride over uniform work: exactly one resident wave (SMs x CTAs that fit), no low-occupancy tail
#define HM_SPACE_GO(G, S)                                                                                              \
    do {                                                                                                               \
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_space_kernel<G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        int occ = 0;                                                                                                   \
        HM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hm_space_kernel<G, S>, HM_SPACE_BLOCK, smem));  \
        const int grid = (int)std::min<long long>(blocks, (long long)hm_num_sms() * (occ < 1 ? 1 : occ));              \
        hm_space_kernel<G, S><<<grid, HM_SPACE_BLOCK, smem, st>>>(a);                                                  \
    } while (0)
    if (generate) HM_SPACE_GO(true, false);
    else if (staged) HM_SPACE_GO(false, true);
    else HM_SPACE_GO(false, false);
#undef HM_SPACE_GO
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int hm_encode(const hm_space_t* s, const double* raw, int raw_col_major, int64_t ldr, int64_t N,
                         float* X32cm, int64_t ld32, double* X64cm, int64_t ld64, int32_t* n_invalid_dev,
                         void* stream) {
    HM_REQUIRE(s != nullptr && (raw != nullptr || N == 0), "space/raw");
    HM_REQUIRE(N >= 0 && (raw_col_major ? ldr >= N : ldr >= s->D), "leading dimension of raw");
    HM_REQUIRE(X32cm != nullptr || X64cm != nullptr, "at least one encoded output");
    HM_REQUIRE((X32cm == nullptr || ld32 >= N) && (X64cm == nullptr || ld64 >= N), "ld of the encoded outputs");
    HmSpaceArgs a = {};
    a.raw = raw;
    a.raw_col_major = raw_col_major;
    a.ldr = ldr;
    a.N = N;
    a.x32 = X32cm;
    a.ld32 = ld32;
    a.x64 = X64cm;
    a.ld64 = ld64;
    a.n_invalid = n_invalid_dev;
    return hm_space_launch(s, a, false, stream);
}

extern "C" int hm_prior_prob(const hm_space_t* s, const double* raw, int raw_col_major, int64_t ldr, int64_t N,
                             const double* objective_weights_host, int n_objectives, double* prob_out,
                             int32_t* n_invalid_dev, void* stream) {
    HM_REQUIRE(s != nullptr && (raw != nullptr || N == 0) && prob_out != nullptr, "space/raw/prob_out");
    HM_REQUIRE(N >= 0 && (raw_col_major ? ldr >= N : ldr >= s->D), "leading dimension of raw");
    HM_REQUIRE(n_objectives >= 1 && n_objectives <= HM_MAX_OBJECTIVES && objective_weights_host != nullptr, "weights");
    HmSpaceArgs a = {};
    a.raw = raw;
    a.raw_col_major = raw_col_major;
    a.ldr = ldr;
    a.N = N;
    a.prob = prob_out;
    a.n_obj = n_objectives;
    for (int o = 0; o < n_objectives; ++o) a.w[o] = objective_weights_host[o];
    a.n_invalid = n_invalid_dev;
    return hm_space_launch(s, a, false, stream);
}

extern "C" int hm_sample_pool(const hm_space_t* s, uint64_t seed, int64_t index_base, int64_t N, int use_priors,
                              double* raw_cm_out, int64_t ldr, float* X32cm, int64_t ld32, double* X64cm,
                              int64_t ld64, const double* objective_weights_host, int n_objectives,
                              double* prob_out, void* stream) {
    HM_REQUIRE(s != nullptr && N >= 0 && index_base >= 0, "space/N/index_base");
    HM_REQUIRE(raw_cm_out == nullptr || ldr >= N, "ldr >= N");
    HM_REQUIRE((X32cm == nullptr || ld32 >= N) && (X64cm == nullptr || ld64 >= N), "ld of the encoded outputs");
    HM_REQUIRE(prob_out == nullptr || (n_objectives >= 1 && n_objectives <= HM_MAX_OBJECTIVES &&
                                       objective_weights_host != nullptr), "weights for prob_out");
    HmSpaceArgs a = {};
    a.N = N;
    a.seed = seed;
    a.index_base = index_base;
    a.use_priors = use_priors ? 1 : 0;
    a.raw_out = raw_cm_out;
    a.ldro = ldr;
    a.x32 = X32cm;
    a.ld32 = ld32;
    a.x64 = X64cm;
    a.ld64 = ld64;
    a.prob = prob_out;
    a.n_obj = prob_out ? n_objectives : 0;
    for (int o = 0; o < a.n_obj; ++o) a.w[o] = objective_weights_host[o];
    return hm_space_launch(s, a, true, stream);
}

// packed 64-bit codes (hm_pack.cuh) instead of encoded columns: K5 / K6 for all-discrete spaces
extern "C" int hm_encode_packed(const hm_space_t* s, const double* raw, int raw_col_major, int64_t ldr, int64_t N,
                                uint64_t* codes_out, int32_t* n_invalid_dev, void* stream) {
    HM_REQUIRE(s != nullptr && (raw != nullptr || N == 0) && codes_out != nullptr, "space/raw/codes_out");
    HM_REQUIRE(s->pack_bits > 0, "the space has real / integer parameters or needs more than 64 bits: not packable");
    HM_REQUIRE(N >= 0 && (raw_col_major ? ldr >= N : ldr >= s->D), "leading dimension of raw");
    HmSpaceArgs a = {};
    a.raw = raw;
    a.raw_col_major = raw_col_major;
    a.ldr = ldr;
    a.N = N;
    a.codes = reinterpret_cast<unsigned long long*>(codes_out);
    a.pack_low = s->d_pack_low;
    a.n_invalid = n_invalid_dev;
    return hm_space_launch(s, a, false, stream);
}

extern "C" int hm_sample_pool_packed(const hm_space_t* s, uint64_t seed, int64_t index_base, int64_t N, int use_priors,
                                     uint64_t* codes_out, double* raw_cm_out, int64_t ldr, void* stream) {
    HM_REQUIRE(s != nullptr && N >= 0 && index_base >= 0 && codes_out != nullptr, "space/N/index_base/codes_out");
    HM_REQUIRE(s->pack_bits > 0, "the space has real / integer parameters or needs more than 64 bits: not packable");
    HM_REQUIRE(raw_cm_out == nullptr || ldr >= N, "ldr >= N");
    HmSpaceArgs a = {};
    a.N = N;
    a.seed = seed;
    a.index_base = index_base;
    a.use_priors = use_priors ? 1 : 0;
    a.raw_out = raw_cm_out;
    a.ldro = ldr;
    a.codes = reinterpret_cast<unsigned long long*>(codes_out);
    a.pack_low = s->d_pack_low;
    return hm_space_launch(s, a, true, stream);
}

extern "C" int hm_sample_rows(const hm_space_t* s, uint64_t seed, int use_priors, const int64_t* index_dev, int64_t n,
                              double* raw_rows_out, void* stream) {
    HM_REQUIRE(s != nullptr && n >= 0 && (n == 0 || (index_dev != nullptr && raw_rows_out != nullptr)), "arguments");
    HmSpaceArgs a = {};
    a.N = n;
    a.seed = seed;
    a.use_priors = use_priors ? 1 : 0;
    a.gather = reinterpret_cast<const long long*>(index_dev);
    a.raw_out = raw_rows_out;
    a.raw_out_rows = 1;
    return hm_space_launch(s, a, true, stream);
}

// ---- piBO weighting (pibo.py:101-113) ---------------------------------------------------------------------------
//   out = acq * (prob + floor) ** exponent;   TS/UCB: acq is first collapsed to np.min(acq) (pibo.py:103-104)
__device__ __forceinline__ double hm_key_to_double(unsigned long long k) {
    if (k == 0ull) return __longlong_as_double(0x7ff8000000000000ll);
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__global__ void __launch_bounds__(256) hm_min_key_kernel(const double* __restrict__ v, long long N,
                                                         unsigned long long* __restrict__ out) {
    unsigned long long best = ~0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = hm_order_key(__ldg(v + i));
        best = k < best ? k : best;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(out, best);
}
__global__ void __launch_bounds__(256) hm_prior_weight_kernel(const double* __restrict__ acq, const double* __restrict__ prob,
                                                              long long N, double floor_, double exponent,
                                                              const unsigned long long* __restrict__ min_key,
                                                              double* __restrict__ out) {
    const bool collapse = min_key != nullptr;
    const double amin = collapse ? hm_key_to_double(*min_key) : 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double a = collapse ? amin : __ldg(acq + i);
        out[i] = hm_mul(a, pow(hm_add(__ldg(prob + i), floor_), exponent));
    }
}
extern "C" int hm_prior_weight(const double* acq, const double* prob, int64_t N, double prior_floor, double exponent,
                               int collapse_to_min, double* out, void* workspace8, void* stream) {
    HM_REQUIRE(acq != nullptr && prob != nullptr && out != nullptr && N >= 0, "acq/prob/out");
    HM_REQUIRE(!collapse_to_min || workspace8 != nullptr, "8 bytes of device workspace for the minimum");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 256;
    const int grid = (int)std::min<long long>((N + block - 1) / block, (long long)hm_num_sms() * 8);
    unsigned long long* mk = nullptr;
    if (collapse_to_min) {
        mk = reinterpret_cast<unsigned long long*>(workspace8);
        HM_CUDA_TRY(cudaMemsetAsync(mk, 0xff, 8, st));
        hm_min_key_kernel<<<grid, block, 0, st>>>(acq, N, mk);
    }
    hm_prior_weight_kernel<<<grid, block, 0, st>>>(acq, prob, N, prior_floor, exponent, mk, out);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- BOPrO posterior ratio (prior_optimization.compute_EI_from_posteriors, prior_optimization.py:126-374) -------------
// min / max of a vector; ignore_inf: +-inf entries are skipped (prior_optimization.py:318-327); NaNs are skipped.
// out2 = {min key, max key} as hm_order_key values (atomicMin / atomicMax), decoded by hm_minmax_decode.
__global__ void __launch_bounds__(256) hm_minmax_kernel(const double* __restrict__ v, long long N, int ignore_inf,
                                                        unsigned long long* __restrict__ out2) {
    unsigned long long lo = ~0ull, hi = 0ull;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double x = __ldg(v + i);
        if (x != x) continue;
        if (ignore_inf && (x == inf || x == -inf)) continue;
        const unsigned long long k = hm_order_key(x);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo != ~0ull) atomicMin(out2, lo);
        if (hi != 0ull) atomicMax(out2 + 1, hi);
    }
}
__global__ void hm_minmax_decode(const unsigned long long* __restrict__ keys, double* __restrict__ out2) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    out2[0] = keys[0] == ~0ull ? inf : hm_key_to_double(keys[0]);
    out2[1] = keys[1] == 0ull ? -inf : hm_key_to_double(keys[1]);
}
extern "C" int hm_minmax(const double* v, int64_t N, int ignore_inf, double* out2_dev, void* workspace16, void* stream) {
    HM_REQUIRE(out2_dev != nullptr && workspace16 != nullptr && N >= 0 && (N == 0 || v != nullptr), "arguments");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(workspace16);
    HM_CUDA_TRY(cudaMemsetAsync(keys, 0xff, 8, st));
    HM_CUDA_TRY(cudaMemsetAsync(keys + 1, 0x00, 8, st));
    if (N > 0) {
        const int block = 256;
        const int grid = (int)std::min<long long>((N + block - 1) / block, (long long)hm_num_sms() * 8);
        hm_minmax_kernel<<<grid, block, 0, st>>>(v, N, ignore_inf, keys);
    }
    hm_minmax_decode<<<1, 1, 0, st>>>(keys, out2_dev);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

struct HmBoproDev {
    int M, prior_mode, discrete, pad;
    double w[HM_MAX_OBJECTIVES], thr[HM_MAX_OBJECTIVES];
    double prior_lo, prior_hi, floor_, divisor, exponent;
};
// ratio[i] = log posterior good - log posterior bad  (prior_optimization.py:168-313)
__global__ void __launch_bounds__(256) hm_bopro_ratio_kernel(const HmBoproDev p, const double* __restrict__ prior,
                                                             const double* __restrict__ mean,
                                                             const double* __restrict__ var, long long ldm, long long N,
                                                             double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        double pg = __ldg(prior + i);
        if (p.prior_mode == 2) {
            pg = 1.0;                                                           // :181-182
        } else if (p.prior_mode == 1) {                                         // :184-192
            pg = hm_add(p.floor_, hm_mul(hm_sub(1.0, p.floor_), hm_sub(pg, p.prior_lo)) / hm_sub(p.prior_hi, p.prior_lo));
        }
        double pb = hm_sub(1.0, pg);                                            // :195
        if (pb < p.floor_) pb = p.floor_;                                       // :197
        if (p.discrete) pb = pb / p.divisor;                                    // :204-205
        double mg = 1.0, mb = 1.0;
        for (int m = 0; m < p.M; ++m) {                                         // compute_probability_from_model :126-158
            // models.py:762-763 (NaN mean -> sys.maxsize) and the var=False branch np.sqrt(variance) (models.py:790, 1021)
            double mu = __ldg(mean + (size_t)m * (size_t)ldm + (size_t)i);
            if (mu != mu) mu = 9.223372036854775807e18;
            const double sd = sqrt(__ldg(var + (size_t)m * (size_t)ldm + (size_t)i));
            const double c = hm_ndtr(hm_sub(p.thr[m], mu) / sd);
            const double good = c, bad = hm_sub(1.0, c);
            mg = hm_mul(mg, p.w[m] == 1.0 ? good : pow(good, p.w[m]));
            mb = hm_mul(mb, p.w[m] == 1.0 ? bad : pow(bad, p.w[m]));
        }
        const double lg = hm_add(log(pg), hm_mul(p.exponent, log(mg)));         // :302-307
        const double lb = hm_add(log(pb), hm_mul(p.exponent, log(mb)));
        out[i] = hm_sub(lg, lb);
    }
}
extern "C" int hm_bopro_ratio(const hm_bopro_params_t* bp, const double* prior_prob, const double* mean,
                              const double* var, int64_t ldm, int64_t N, double* ratio_out, void* stream) {
    HM_REQUIRE(bp != nullptr && prior_prob != nullptr && mean != nullptr && var != nullptr && ratio_out != nullptr,
               "arguments");
    HM_REQUIRE(bp->n_objectives >= 1 && bp->n_objectives <= HM_MAX_OBJECTIVES && N >= 0 && ldm >= N, "M / N / ldm");
    if (N == 0) return 0;
    HmBoproDev d;
    d.M = bp->n_objectives;
    d.prior_mode = bp->prior_mode;
    d.discrete = bp->discrete_space;
    d.pad = 0;
    for (int m = 0; m < d.M; ++m) {
        d.w[m] = bp->weight[m];
        d.thr[m] = bp->threshold[m];
    }
    d.prior_lo = bp->prior_lo;
    d.prior_hi = bp->prior_hi;
    d.floor_ = bp->posterior_floor;
    d.divisor = bp->discrete_divisor;
    d.exponent = bp->model_exponent;
    const int block = 256;
    const int grid = (int)std::min<long long>((N + block - 1) / block, (long long)hm_num_sms() * 8);
    hm_bopro_ratio_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d, prior_prob, mean, var, ldm, N, ratio_out);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

// out = -(normalised ratio + feasibility indicator), +-inf -> +-sys.maxsize (prior_optimization.py:315-353);
// feas_out (nullable): the feasibility indicator itself (:263-281)
__global__ void __launch_bounds__(256) hm_bopro_finish_kernel(const double* __restrict__ ratio, long long N, int mode,
                                                              double lo, double hi, double floor_,
                                                              const double* __restrict__ feas_prob,
                                                              double* __restrict__ out, double* __restrict__ feas_out) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const double logf = log(floor_);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        double r = __ldg(ratio + i);
        if (mode == 2) r = 1.0;                                                                   // :330-331
        else if (mode == 1) r = hm_add(floor_, hm_mul(hm_sub(1.0, floor_), hm_sub(r, lo)) / hm_sub(hi, lo));   // :334-340
        double f = 1.0;
        if (feas_prob) {
            double q = __ldg(feas_prob + i);
            if (q == 0.0) q = floor_;                                                             // :270
            q = log(q);
            f = hm_add(floor_, hm_mul(hm_sub(1.0, floor_), hm_sub(q, logf)) / hm_sub(0.0, logf));  // :274-279
        }
        double v = hm_mul(-1.0, hm_add(r, f));                                                    // :345-346
        if (v == inf) v = 9223372036854775807.0;
        if (v == -inf) v = -9223372036854775807.0;
        out[i] = v;
        if (feas_out) feas_out[i] = f;
    }
}
extern "C" int hm_bopro_finish(const double* ratio, int64_t N, int normalize_mode, double lo, double hi,
                               double posterior_floor, const double* feas_prob, double* out, double* feas_out,
                               void* stream) {
    HM_REQUIRE(ratio != nullptr && out != nullptr && N >= 0, "ratio/out");
    HM_REQUIRE(normalize_mode >= 0 && normalize_mode <= 2, "normalize_mode in {0 none, 1 rescale, 2 ones}");
    if (N == 0) return 0;
    const int block = 256;
    const int grid = (int)std::min<long long>((N + block - 1) / block, (long long)hm_num_sms() * 8);
    hm_bopro_finish_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(ratio, N, normalize_mode, lo, hi, posterior_floor,
                                                                     feas_prob, out, feas_out);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- N4: neighbourhoods and moves of the multi-start hill climb on the device (discrete spaces) -----------------------
// get_neighbors (local_search.py:83-161) for spaces without real / integer parameters draws no random numbers, and its
// `cand not in neighbors` filter only ever removes repetitions of the unchanged configuration, so a configuration has a
// FIXED number K of neighbours, enumerated parameter by parameter:
//   categorical, or ordinal with <= 5 values : every value in order                    (local_search.py:103-109)
//   ordinal                                   : the window of 5 values around the current index (:110-137)
//   and for every parameter but the first the entry equal to the current value is skipped (it repeats the unchanged
//   configuration, which parameter 0 already emitted).
// plan[q] = {parameter, rank inside that parameter's emitted list} for q < K, built by the host.
__global__ void __launch_bounds__(256) hm_ls_neighbors_kernel(const hm_param_desc_t* __restrict__ params,
                                                              const double* __restrict__ tables, int D,
                                                              const int2* __restrict__ plan, int K,
                                                              const double* __restrict__ current,     // [S][D] raw
                                                              const int* __restrict__ live, int n_live,
                                                              double* __restrict__ rows) {           // [n_live*K][D]
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_live * K) return;
    const int ls = (int)(t / K), q = (int)(t - (long long)ls * K);
    const int s = live[ls];
    const int j = plan[q].x, e = plan[q].y;
    const hm_param_desc_t& p = params[j];
    const double cur = current[(size_t)s * D + j];
    // index of the current value
    int idx;
    if (p.kind == HM_P_ORDINAL) {
        idx = 0;
        for (int v = 0; v < p.n_values; ++v)
            if (tables[p.values_off + v] == cur) idx = v;
    } else {
        idx = (int)cur;
    }
    int pick;                                     // value index of this neighbour
    if (p.kind == HM_P_CATEGORICAL || p.n_values <= 5) {
        pick = (j > 0 && e >= idx) ? e + 1 : e;
    } else {
        const int n = p.n_values;
        const int w0 = idx < 2 ? 0 : ((n - idx) <= 2 ? n - 5 : idx - 2);
        const int rel = idx - w0;                 // position of the current value inside the window
        pick = w0 + ((j > 0 && e >= rel) ? e + 1 : e);
    }
    double* out = rows + (size_t)t * D;
    for (int d = 0; d < D; ++d) out[d] = current[(size_t)s * D + d];
    out[j] = p.kind == HM_P_ORDINAL ? tables[p.values_off + pick] : (double)pick;
}

// One warp per live start: argmin of its K scores with get_min_configurations' order (NaN first, -0.0 == +0.0, first
// index among equals); the start has converged iff the best neighbour is the unchanged configuration, else it moves.
__global__ void __launch_bounds__(32) hm_ls_move_kernel(const double* __restrict__ scores, const double* __restrict__ rows,
                                                        int K, int D, const int* __restrict__ live,
                                                        double* __restrict__ current, int* __restrict__ converged,
                                                        int* __restrict__ best_q) {
    const int ls = blockIdx.x, lane = threadIdx.x;
    const int s = live[ls];
    HmPair best = {~0ull, 0x7fffffffffffffffll};
    for (int q = lane; q < K; q += 32) {
        HmPair e;
        e.key = hm_order_key(scores[(size_t)ls * K + q]);
        e.idx = q;
        if (hm_pair_less(e, best)) best = e;
    }
    for (int o = 16; o > 0; o >>= 1) {
        HmPair other;
        other.key = __shfl_xor_sync(0xffffffffu, best.key, o);
        other.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
        if (hm_pair_less(other, best)) best = other;
    }
    const int q = (int)best.idx;
    const double* row = rows + ((size_t)ls * K + q) * D;
    bool same = true;
    for (int d = lane; d < D; d += 32) same = same && (row[d] == current[(size_t)s * D + d]);
    same = __all_sync(0xffffffffu, same);
    if (!same)
        for (int d = lane; d < D; d += 32) current[(size_t)s * D + d] = row[d];
    if (lane == 0) {
        converged[s] = same ? 1 : 0;
        best_q[s] = q;
    }
}

extern "C" int hm_ls_neighbors(const hm_space_t* s, const int32_t* plan_dev, int K, const double* current_dev,
                               const int32_t* live_dev, int n_live, double* rows_out, void* stream) {
    HM_REQUIRE(s != nullptr && plan_dev != nullptr && current_dev != nullptr && live_dev != nullptr && rows_out != nullptr,
               "arguments");
    HM_REQUIRE(K >= 1 && n_live >= 0, "K >= 1, n_live >= 0");
    for (int j = 0; j < s->D; ++j)
        HM_REQUIRE(s->params[j].kind >= HM_P_ORDINAL, "hm_ls_neighbors: discrete spaces only (ordinal / categorical)");
    if (n_live == 0) return 0;
    const long long total = (long long)n_live * K;
    hm_ls_neighbors_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        s->d_params, s->d_tables, s->D, reinterpret_cast<const int2*>(plan_dev), K, current_dev, live_dev, n_live, rows_out);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int hm_ls_move(const double* scores, const double* rows, int K, int D, const int32_t* live_dev, int n_live,
                          double* current_dev, int32_t* converged_dev, int32_t* best_q_dev, void* stream) {
    HM_REQUIRE(scores && rows && live_dev && current_dev && converged_dev && best_q_dev, "arguments");
    HM_REQUIRE(K >= 1 && D >= 1 && n_live >= 0, "K, D >= 1, n_live >= 0");
    if (n_live == 0) return 0;
    hm_ls_move_kernel<<<n_live, 32, 0, (cudaStream_t)stream>>>(scores, rows, K, D, live_dev, current_dev, converged_dev,
                                                             best_q_dev);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

HM_CHECKED_EXPORT(space)
