// K3 (stand-alone part): scalarised acquisition on precomputed mean/variance, and the stable k-smallest
// selection that replaces utility_functions.get_min_configurations (utility_functions.py:295-316).
//
// Selection = the k lexicographically smallest (order_key(value), global index) pairs, ascending: exactly what
// k rounds of np.argmin + delete produce (ties -> lowest index, NaN first).  Per CTA: a shared-memory candidate
// buffer plus a running threshold tau = current k-th smallest pair; an element is appended only if it beats tau,
// so after the first tile appends are rare (expected k*ln(n/k) per CTA) and the stream is read at HBM speed.
// When the buffer could overflow it is bitonic-sorted, truncated to k and tau is tightened.  Stage 2 merges the
// per-CTA lists with the same routine in one CTA.
#include <algorithm>
#include "hm_acq.cuh"

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hm_acq_kernel(HmAcqDev p, const double* __restrict__ mean,
                                                     const double* __restrict__ var, long long ldm,
                                                     const double* __restrict__ feas, long long N,
                                                     double* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x) {
        double m[HM_MAX_OBJECTIVES], v[HM_MAX_OBJECTIVES];
#pragma unroll
        for (int o = 0; o < HM_MAX_OBJECTIVES; ++o) {
            if (o < p.M) {
                double mm = __ldg(mean + (size_t)o * ldm + i);
                // models.py:762-763 (NaN mean -> sys.maxsize) is applied by the caller of the prediction in the
                // reference; the acquisition sees the patched mean.
                if (p.kind != HM_ACQ_TS && mm != mm) mm = 9.223372036854775807e18;
                m[o] = mm;
                v[o] = (p.kind == HM_ACQ_TS || var == nullptr) ? 0.0 : __ldg(var + (size_t)o * ldm + i);
            } else {
                m[o] = 0.0;
                v[o] = 0.0;
            }
        }
        const double f = feas ? __ldg(feas + i) : 1.0;
        out[i] = hm_acq_value(p, m, v, f);
    }
}

extern "C" int hm_acq_score(const hm_acq_params_t* acq, const double* mean, const double* var, int64_t ldm,
                            const double* feas, int64_t N, double* score_out, void* stream) {
    HM_REQUIRE(acq != nullptr && mean != nullptr && score_out != nullptr, "acq/mean/score_out");
    HM_REQUIRE(N >= 0 && ldm >= N, "ldm >= N >= 0");
    HM_REQUIRE(acq->kind == HM_ACQ_TS || var != nullptr, "var required for EI/UCB");
    if (N == 0) return 0;
    HmAcqDev d;
    int rc = hm_acq_to_dev(acq, &d);
    if (rc) {
        hm_set_error("invalid acquisition parameters");
        return rc;
    }
    const int block = 256;
    long long blocks = (N + block - 1) / block;
    int grid = (int)std::min<long long>(blocks, (long long)hm_num_sms() * 8);
    hm_acq_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d, mean, var, ldm, feas, N, score_out);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
#define HM_TK_BLOCK 512
#define HM_TK_CAP 2048
#define HM_TK_TILE 1024
#define HM_TK_MAXK 1024
#define HM_TK_MAXGRID 512

__device__ __forceinline__ HmPair hm_pad_pair() {
    HmPair p;
    p.key = ~0ull;
    p.idx = 0x7fffffffffffffffll;
    return p;
}

// bitonic sort of buf[0..n2) (n2 power of two, <= HM_TK_CAP) by (key, idx); all threads of the CTA participate
__device__ void hm_bitonic_sort(HmPair* buf, int n2) {
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                HmPair a = buf[lo], b = buf[hi];
                if (hm_pair_less(b, a) == up) {
                    buf[lo] = b;
                    buf[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

struct HmSelState {
    HmPair* buf;      // [HM_TK_CAP] shared
    int* count;       // shared
    HmPair* tau;      // shared
    double* tau_v;    // shared: the value tau's key stands for (+inf while fewer than k elements are held)
};

// keep the k smallest of buf[0..*count); tighten tau.  Called by all threads (uniformly).
__device__ void hm_sel_compact(HmSelState& s, int k) {
    __syncthreads();
    const int c = *s.count;
    int n2 = 2;
    while (n2 < c) n2 <<= 1;
    for (int t = c + threadIdx.x; t < n2; t += blockDim.x) s.buf[t] = hm_pad_pair();
    __syncthreads();
    hm_bitonic_sort(s.buf, n2);
    if (threadIdx.x == 0) {
        const int nc = c < k ? c : k;
        *s.count = nc;
        if (nc >= k) {
            *s.tau = s.buf[k - 1];
            const unsigned long long key = s.buf[k - 1].key;
            // inverse of hm_order_key (key 0 = NaN: then nothing is ever rejected by the value test)
            const unsigned long long b = (key & 0x8000000000000000ull) ? (key ^ 0x8000000000000000ull) : ~key;
            *s.tau_v = key == 0ull ? __longlong_as_double(0x7ff8000000000000ll) : __longlong_as_double((long long)b);
        }
    }
    __syncthreads();
}

// returns the shared buffer whose first min(k, #elements) entries are the selection, ascending.
// The CTA owns the contiguous element range [lo, hi).  It walks it in super-tiles of HM_TK_SUPER elements: every thread
// issues all its loads of the super-tile first (8 independent 8-byte loads: enough bytes in flight for HBM speed), then
// filters them on the raw value -- an element can only enter the buffer if NOT (v > tau_v); NaN, -0.0 and ties fall
// through to the exact (key, index) comparison -- and the CTA synchronises ONCE per super-tile.  A super-tile that
// would overflow the buffer (only while tau is still loose, i.e. the first one or two of a CTA, or adversarial input) is
// rolled back and redone in tiles of HM_TK_TILE elements with a compaction whenever the buffer is more than half full.
#define HM_TK_RS 8
#define HM_TK_SUPER (HM_TK_BLOCK * HM_TK_RS)
template <typename Load>
__device__ const HmPair* hm_select_k(Load load, long long lo, long long hi, int k, int* n_selected) {
    __shared__ HmPair buf[HM_TK_CAP];
    __shared__ int count;
    __shared__ int overflow;
    __shared__ HmPair tau;
    __shared__ double tau_v;
    HmSelState s = {buf, &count, &tau, &tau_v};
    if (threadIdx.x == 0) {
        count = 0;
        overflow = 0;
        tau = hm_pad_pair();
        tau_v = __longlong_as_double(0x7ff0000000000000ll);
    }
    __syncthreads();
    // warm-up: the first HM_TK_BLOCK elements go straight into the buffer and give the first threshold (one small sort),
    // so that the first super-tile does not overflow; later compactions are triggered early (a few hundred entries)
    // so that they sort 512 entries, not 2048 -- the sorts, not the stream, were most of the kernel
    {
        const long long p = lo + threadIdx.x;
        if (p < hi) {
            const int pos = atomicAdd(&count, 1);
            HM_DEV_ASSERT(pos >= 0 && pos < HM_TK_CAP);
            buf[pos] = load.pair(load.item(p), p);
        }
        __syncthreads();
        if (hi - lo > HM_TK_BLOCK) hm_sel_compact(s, k);
    }
    const int compact_at = min(HM_TK_CAP / 2, max(256, 2 * k));
    for (long long e0 = lo + HM_TK_BLOCK; e0 < hi; e0 += HM_TK_SUPER) {
        const long long e1 = e0 + HM_TK_SUPER < hi ? e0 + HM_TK_SUPER : hi;
        const int c0 = count;                       // the roll-back point of this super-tile ...
        const HmPair tl = tau;
        const double tv = tau_v;
        __syncthreads();                            // ... which every thread must have read before anybody appends
        typename Load::Item it[HM_TK_RS];
#pragma unroll
        for (int r = 0; r < HM_TK_RS; ++r) {
            const long long p = e0 + r * HM_TK_BLOCK + threadIdx.x;
            if (p < e1) it[r] = load.item(p);
        }
#pragma unroll
        for (int r = 0; r < HM_TK_RS; ++r) {
            const long long p = e0 + r * HM_TK_BLOCK + threadIdx.x;
            if (p < e1 && load.maybe(it[r], tv, p, tl.idx)) {
                const HmPair e = load.pair(it[r], p);
                if (hm_pair_less(e, tl)) {
                    const int pos = atomicAdd(&count, 1);
                    if (pos < HM_TK_CAP) buf[pos] = e;
                    else overflow = 1;
                }
            }
        }
        __syncthreads();
        const bool redo = overflow != 0;
        const int c = count;
        __syncthreads();
        if (redo) {
            if (threadIdx.x == 0) {
                count = c0;
                overflow = 0;
            }
            __syncthreads();
            for (long long t0 = e0; t0 < e1; t0 += HM_TK_TILE) {
                if (count > HM_TK_CAP - HM_TK_TILE) hm_sel_compact(s, k);      // uniform; leaves room for a whole tile
                const HmPair tl2 = tau;
#pragma unroll
                for (int r = 0; r < HM_TK_TILE / HM_TK_BLOCK; ++r) {
                    const long long p = t0 + r * HM_TK_BLOCK + threadIdx.x;
                    if (p < e1 && p < t0 + HM_TK_TILE) {
                        const HmPair e = load.pair(load.item(p), p);
                        if (hm_pair_less(e, tl2)) {
                            const int pos = atomicAdd(&count, 1);
                            HM_DEV_ASSERT(pos >= 0 && pos < HM_TK_CAP);      // the compaction above left room for a whole tile
                            buf[pos] = e;
                        }
                    }
                }
                __syncthreads();
            }
        } else if (c > compact_at) {
            hm_sel_compact(s, k);
        }
    }
    hm_sel_compact(s, k);
    HM_DEV_ASSERT(count >= 0 && count <= k && count <= HM_TK_CAP && overflow == 0);
    *n_selected = count;
    return buf;
}

// loaders: item(p) = what is fetched early, maybe(item, tau_value) = cheap necessary condition, pair(item, p) = exact key
struct HmLoadVals {
    typedef double Item;
    const double* vals;
    long long base;
    __device__ __forceinline__ double item(long long p) const { return __ldg(vals + p); }
    // exact for non-NaN values: smaller value, or equal value (-0.0 == +0.0) with a smaller global index
    __device__ __forceinline__ bool maybe(double v, double tv, long long p, long long tau_idx) const {
        return v < tv || (v == tv && base + p < tau_idx) || v != v;
    }
    __device__ __forceinline__ HmPair pair(double v, long long p) const {
        HmPair e;
        e.key = hm_order_key(v);
        e.idx = base + p;
        return e;
    }
};
struct HmLoadValIdx {
    typedef double Item;
    const double* vals;
    const long long* idx;
    __device__ __forceinline__ double item(long long p) const { return __ldg(vals + p); }
    __device__ __forceinline__ bool maybe(double v, double tv, long long, long long) const { return !(v > tv); }
    __device__ __forceinline__ HmPair pair(double v, long long p) const {
        HmPair e;
        e.key = hm_order_key(v);
        e.idx = __ldg(idx + p);
        return e;
    }
};
struct HmLoadPairs {
    typedef HmPair Item;
    const HmPair* pairs;
    __device__ __forceinline__ HmPair item(long long p) const { return pairs[p]; }
    __device__ __forceinline__ bool maybe(const HmPair&, double, long long, long long) const { return true; }
    __device__ __forceinline__ HmPair pair(const HmPair& e, long long) const { return e; }
};

template <typename Load>
__global__ void __launch_bounds__(HM_TK_BLOCK) hm_topk_stage1(Load load, long long n, int k, HmPair* ws) {
    int c;
    // contiguous share of this CTA, a multiple of the super-tile
    const long long supers = (n + HM_TK_SUPER - 1) / HM_TK_SUPER;
    const long long per = (supers + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per * HM_TK_SUPER;
    long long hi = lo + per * HM_TK_SUPER;
    if (hi > n) hi = n;
    const HmPair* sel = hm_select_k(load, lo < n ? lo : n, hi, k, &c);
    HmPair* out = ws + (size_t)blockIdx.x * k;
    for (int t = threadIdx.x; t < k; t += blockDim.x) out[t] = (t < c) ? sel[t] : hm_pad_pair();
}

__device__ __forceinline__ double hm_key_to_value(unsigned long long key) {
    if (key == 0ull) return __longlong_as_double(0x7ff8000000000000ll);  // NaN
    unsigned long long b = (key & 0x8000000000000000ull) ? (key ^ 0x8000000000000000ull) : ~key;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(HM_TK_BLOCK) hm_topk_stage2(const HmPair* ws, long long n, int k, double* out_vals,
                                                              long long* out_idx) {
    HmLoadPairs load = {ws};
    int c;
    const HmPair* sel = hm_select_k(load, 0, n, k, &c);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const HmPair e = (t < c) ? sel[t] : hm_pad_pair();
        const bool pad = (e.key == ~0ull && e.idx == 0x7fffffffffffffffll);
        out_vals[t] = pad ? __longlong_as_double(0x7ff0000000000000ll) : hm_key_to_value(e.key);
        out_idx[t] = pad ? -1ll : e.idx;
    }
}

extern "C" int64_t hm_topk_workspace_bytes(int k) {
    if (k < 1) k = 1;
    return (int64_t)HM_TK_MAXGRID * k * (int64_t)sizeof(HmPair) + 256;
}

template <typename Load>
static int hm_topk_run(Load load, int64_t n, int k, double* out_vals, int64_t* out_idx, void* workspace,
                       cudaStream_t st) {
    HM_REQUIRE(k >= 1 && k <= HM_TK_MAXK, "1 <= k <= 1024");
    HM_REQUIRE(out_vals && out_idx && workspace, "outputs/workspace");
    HM_REQUIRE(n >= 0, "n >= 0");
    HmPair* ws = reinterpret_cast<HmPair*>(workspace);
    long long tiles = (n + HM_TK_TILE - 1) / HM_TK_TILE;
    int occ = 0;
    HM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hm_topk_stage1<Load>, HM_TK_BLOCK, 0));
    if (occ < 1) occ = 1;
    int grid = (int)std::max<long long>(1, std::min<long long>(tiles, std::min(HM_TK_MAXGRID, hm_num_sms() * occ)));
    hm_topk_stage1<<<grid, HM_TK_BLOCK, 0, st>>>(load, (long long)n, k, ws);
    HM_CUDA_TRY(cudaGetLastError());
    hm_topk_stage2<<<1, HM_TK_BLOCK, 0, st>>>(ws, (long long)grid * k, k, out_vals, (long long*)out_idx);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int hm_topk(const double* vals, int64_t N, int64_t index_base, int k, double* out_vals, int64_t* out_idx,
                       void* workspace, void* stream) {
    HM_REQUIRE(vals != nullptr || N == 0, "vals");
    HmLoadVals load = {vals, (long long)index_base};
    return hm_topk_run(load, N, k, out_vals, out_idx, workspace, (cudaStream_t)stream);
}

extern "C" int hm_topk_merge(const double* vals, const int64_t* idx, int64_t n, int k, double* out_vals,
                             int64_t* out_idx, void* workspace, void* stream) {
    HM_REQUIRE((vals != nullptr && idx != nullptr) || n == 0, "vals/idx");
    HmLoadValIdx load = {vals, (const long long*)idx};
    return hm_topk_run(load, n, k, out_vals, out_idx, workspace, (cudaStream_t)stream);
}

HM_CHECKED_EXPORT(acq)
