// K4: Pareto dominance filter = compute_pareto.sequential_is_pareto_efficient (compute_pareto.py:89-117).
//
// Pass 1 of the reference sweeps the points in index order; every point that is still efficient when its turn
// comes ("pivot") kills every efficient point j with  NOT any(c_j <= c_pivot).  The sweep is reproduced exactly,
// a batch of B pivot candidates at a time:
//   resolve : one CTA builds the BxB "i kills j" bit matrix of the batch in shared memory (tiled pairwise
//             dominance), then one warp replays the sequential sweep over the bit rows -> which batch members are
//             pivots (efficient at their turn).  Exact even for NaN inputs, where the relation is not transitive.
//   filter  : every efficient point (already-processed pivots, the batch, the not-yet-reached rest) is tested
//             against the <= B pivots held in shared memory; survivors are compacted in index order (block
//             counts -> scan -> scatter), so the next batch is again "the next B efficient points in index order".
// For random data ~(ln N)^M/M! points are ever pivots, so a handful of rounds suffices (10^7 x 3: 5 rounds); the
// first round streams the cost matrix once (HBM bound), later rounds touch only survivors.
// Pass 2 (compute_pareto.py:107-115) is sequential and order dependent by definition; it runs on the survivor set
// in one CTA, one block-wide vote per surviving point.
#include <algorithm>
#include <vector>
#include "hm_common.cuh"

#define HM_PB 1024          // pivot-candidate batch
#define HM_DOM_STRIDE 33    // padded row stride (words) of the bit matrix: conflict-free column writes
#define HM_PF_THREADS 256
#define HM_PF_ITEMS 8
#define HM_PF_CHUNK (HM_PF_THREADS * HM_PF_ITEMS)

// j is killed by p  <=>  for every coordinate NOT (c_j <= c_p)      (compute_pareto.py:102-104)
template <int M>
__device__ __forceinline__ bool hm_kills(const double* __restrict__ cp, const double* __restrict__ cj) {
    bool all = true;
#pragma unroll
    for (int d = 0; d < M; ++d) all = all && !(cj[d] <= cp[d]);
    return all;
}

struct HmParetoRound {
    unsigned n_pivots;
    unsigned new_done;
    unsigned total;
    unsigned pad;
};

template <int M>
__global__ void __launch_bounds__(HM_PB, 1) hm_pareto_resolve(const double* __restrict__ costs,
                                                              const unsigned* __restrict__ alive, unsigned start, int b,
                                                              double* __restrict__ pivots, HmParetoRound* round) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* c = reinterpret_cast<double*>(smem_raw);                          // [HM_PB][M]
    unsigned* dom = reinterpret_cast<unsigned*>(smem_raw + sizeof(double) * HM_PB * M);   // [HM_PB][33]
    __shared__ unsigned pivot_words[HM_PB / 32];
    const int t = threadIdx.x;
    if (t < b) {
        const size_t src = alive ? (size_t)alive[start + t] : (size_t)(start + t);
#pragma unroll
        for (int d = 0; d < M; ++d) c[t * M + d] = costs[src * M + d];
    }
    __syncthreads();
    const int words = (b + 31) >> 5;
    if (t < b) {
        double me[M];
#pragma unroll
        for (int d = 0; d < M; ++d) me[d] = c[t * M + d];
        for (int w = 0; w < words; ++w) {
            unsigned bits = 0;
            const int jn = min(32, b - w * 32);
            for (int jj = 0; jj < jn; ++jj)
                if (hm_kills<M>(me, c + (w * 32 + jj) * M)) bits |= 1u << jj;
            dom[t * HM_DOM_STRIDE + w] = bits;
        }
    }
    __syncthreads();
    if (t < 32) {
        // replay the sequential sweep: lane l owns word l of the "still efficient" and "was a pivot" sets
        unsigned eff = 0, piv = 0;
        if (t < words) {
            const int nb = min(32, b - t * 32);
            eff = (nb == 32) ? 0xffffffffu : ((1u << nb) - 1u);
        }
        for (int i = 0; i < b; ++i) {
            const unsigned w = __shfl_sync(0xffffffffu, eff, i >> 5);
            if ((w >> (i & 31)) & 1u) {
                if (t == (i >> 5)) piv |= 1u << (i & 31);
                if (t < words) eff &= ~dom[i * HM_DOM_STRIDE + t];
            }
        }
        pivot_words[t] = piv;
    }
    __syncthreads();
    // compact the pivots' costs (order irrelevant for the filter)
    unsigned before = 0;
    for (int w = 0; w < (t >> 5); ++w) before += __popc(pivot_words[w]);
    const unsigned myw = pivot_words[t >> 5];
    if (t < b && ((myw >> (t & 31)) & 1u)) {
        const unsigned pos = before + __popc(myw & ((1u << (t & 31)) - 1u));
#pragma unroll
        for (int d = 0; d < M; ++d) pivots[pos * M + d] = c[t * M + d];
    }
    if (t == 0) {
        unsigned n = 0;
        for (int w = 0; w < HM_PB / 32; ++w) n += __popc(pivot_words[w]);
        round->n_pivots = n;
    }
}

// keep flag for every efficient point + per-block survivor counts
template <int M>
__global__ void __launch_bounds__(HM_PF_THREADS) hm_pareto_filter_k(const double* __restrict__ costs,
                                                                    const unsigned* __restrict__ alive, unsigned n_alive,
                                                                    const double* __restrict__ pivots,
                                                                    const HmParetoRound* __restrict__ round,
                                                                    unsigned char* __restrict__ flags,
                                                                    unsigned* __restrict__ block_counts) {
    extern __shared__ __align__(16) unsigned char filter_smem[];
    double* sp = reinterpret_cast<double*>(filter_smem);     // [n_pivots][M]
    const unsigned np = round->n_pivots;
    for (unsigned i = threadIdx.x; i < np * M; i += blockDim.x) sp[i] = pivots[i];
    __syncthreads();
    const unsigned base = blockIdx.x * HM_PF_CHUNK;
    int kept = 0;
#pragma unroll
    for (int r = 0; r < HM_PF_ITEMS; ++r) {
        const unsigned pos = base + r * HM_PF_THREADS + threadIdx.x;
        if (pos < n_alive) {
            const size_t src = alive ? (size_t)alive[pos] : (size_t)pos;
            double cj[M];
#pragma unroll
            for (int d = 0; d < M; ++d) cj[d] = __ldg(costs + src * M + d);
            bool dead = false;
            for (unsigned p = 0; p < np && !dead; ++p) dead = hm_kills<M>(sp + p * M, cj);
            flags[pos] = dead ? 0 : 1;
            kept += dead ? 0 : 1;
        }
    }
    __shared__ unsigned blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
    // warp reduce then one shared atomic per warp
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_down_sync(0xffffffffu, kept, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&blk, (unsigned)kept);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = blk;
}

// exclusive scan of the block counts (single CTA), total -> round->total
__global__ void __launch_bounds__(1024) hm_pareto_scan(unsigned* __restrict__ block_counts, unsigned n_blocks,
                                                       HmParetoRound* round) {
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < n_blocks; base += 1024) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = (i < n_blocks) ? block_counts[i] : 0;
        unsigned x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned s = warp_sums[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, s, o);
                if (threadIdx.x >= o) s += y;
            }
            warp_sums[threadIdx.x] = s;
        }
        __syncthreads();
        const unsigned warp_off = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
        const unsigned incl = x + warp_off + carry;
        if (i < n_blocks) block_counts[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) round->total = carry;
}

// stable scatter of the survivors' original indices; also records how many survivors precede `boundary`
__global__ void __launch_bounds__(HM_PF_THREADS) hm_pareto_scatter(const unsigned* __restrict__ alive, unsigned n_alive,
                                                                   const unsigned char* __restrict__ flags,
                                                                   const unsigned* __restrict__ block_offsets,
                                                                   unsigned* __restrict__ alive_out, unsigned boundary,
                                                                   HmParetoRound* round) {
    __shared__ unsigned warp_sums[HM_PF_THREADS / 32];
    __shared__ unsigned running;
    if (threadIdx.x == 0) running = block_offsets[blockIdx.x];
    __syncthreads();
    const unsigned base = blockIdx.x * HM_PF_CHUNK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 0; r < HM_PF_ITEMS; ++r) {
        const unsigned pos = base + r * HM_PF_THREADS + threadIdx.x;
        const unsigned f = (pos < n_alive) ? flags[pos] : 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, f);
        const unsigned in_warp = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) warp_sums[warp] = __popc(ballot);
        __syncthreads();
        unsigned before = 0;
        for (int w = 0; w < warp; ++w) before += warp_sums[w];
        const unsigned excl = running + before + in_warp;
        if (pos == boundary) round->new_done = excl;
        if (f) alive_out[excl] = alive ? alive[pos] : pos;
        __syncthreads();
        if (threadIdx.x == HM_PF_THREADS - 1) running = excl + f;
        __syncthreads();
    }
}

__global__ void hm_pareto_init_state(unsigned S, unsigned char* __restrict__ state) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) state[i] = 1;
}

// pass 2 (compute_pareto.py:107-115): in index order, an efficient point dies if some currently efficient point has
// one coordinate equal and one coordinate strictly smaller.
template <int M>
__global__ void __launch_bounds__(1024, 1) hm_pareto_pass2_k(const double* __restrict__ costs,
                                                             const unsigned* __restrict__ alive, unsigned S,
                                                             volatile unsigned char* state) {
    for (unsigned i = 0; i < S; ++i) {
        if (!state[i]) continue;                      // uniform: every thread reads the same flag
        double ci[M];
#pragma unroll
        for (int d = 0; d < M; ++d) ci[d] = costs[(size_t)alive[i] * M + d];
        int found = 0;
        for (unsigned j = threadIdx.x; j < S && !found; j += blockDim.x) {
            if (!state[j]) continue;
            bool eq = false, lt = false;
            const size_t src = alive[j];
#pragma unroll
            for (int d = 0; d < M; ++d) {
                const double cj = costs[src * M + d];
                eq = eq || (cj == ci[d]);
                lt = lt || (cj < ci[d]);
            }
            found = (eq && lt) ? 1 : 0;
        }
        const int kill = __syncthreads_or(found);
        if (kill && threadIdx.x == 0) state[i] = 0;
        __syncthreads();
    }
}

__global__ void hm_pareto_write_mask(const unsigned* __restrict__ alive, const unsigned char* __restrict__ state,
                                     unsigned S, unsigned char* __restrict__ mask) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S && (state == nullptr || state[i])) mask[alive[i]] = 1;
}

// ------------------------------------------------------------------------------------------------------------
struct HmParetoWs {
    unsigned* alive[2];
    unsigned char* flags;
    unsigned* block_counts;
    double* pivots;
    unsigned char* state;
    HmParetoRound* round;
};

static size_t hm_align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t hm_pareto_layout(int64_t N, int M, unsigned char* base, HmParetoWs* ws) {
    size_t off = 0;
    const size_t n = (size_t)std::max<int64_t>(N, 1);
    const size_t nblocks = (n + HM_PF_CHUNK - 1) / HM_PF_CHUNK + 1;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? base + off : nullptr;
        off += hm_align256(bytes);
        return p;
    };
    unsigned char* a0 = take(n * 4);
    unsigned char* a1 = take(n * 4);
    unsigned char* fl = take(n);
    unsigned char* bc = take(nblocks * 4);
    unsigned char* pv = take((size_t)HM_PB * M * 8);
    unsigned char* stt = take(n);
    unsigned char* rd = take(sizeof(HmParetoRound));
    if (ws) {
        ws->alive[0] = (unsigned*)a0;
        ws->alive[1] = (unsigned*)a1;
        ws->flags = fl;
        ws->block_counts = (unsigned*)bc;
        ws->pivots = (double*)pv;
        ws->state = stt;
        ws->round = (HmParetoRound*)rd;
    }
    return off;
}

extern "C" int64_t hm_pareto_workspace_bytes(int64_t N, int M) {
    if (M < 1) M = 1;
    return (int64_t)hm_pareto_layout(N, M, nullptr, nullptr) + 256;
}

template <int M>
static int hm_pareto_run(const double* costs, int64_t N, uint8_t* mask, void* workspace, int64_t* n_out, bool pass2,
                         cudaStream_t st) {
    HmParetoWs ws;
    unsigned char* base = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    hm_pareto_layout(N, M, base, &ws);
    const size_t resolve_smem = sizeof(double) * HM_PB * M + sizeof(unsigned) * HM_PB * HM_DOM_STRIDE;
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_pareto_resolve<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resolve_smem));
    const size_t filter_smem = sizeof(double) * HM_PB * M;
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_pareto_filter_k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)filter_smem));

    unsigned n_alive = (unsigned)N, done = 0;
    const unsigned* cur = nullptr;      // nullptr = identity list (first round streams the cost matrix directly)
    int flip = 0;
    HmParetoRound h;
    while (done < n_alive) {
        const unsigned b = std::min<unsigned>(HM_PB, n_alive - done);
        hm_pareto_resolve<M><<<1, HM_PB, resolve_smem, st>>>(costs, cur, done, (int)b, ws.pivots, ws.round);
        const unsigned nblk = (n_alive + HM_PF_CHUNK - 1) / HM_PF_CHUNK;
        hm_pareto_filter_k<M><<<nblk, HM_PF_THREADS, filter_smem, st>>>(costs, cur, n_alive, ws.pivots, ws.round, ws.flags,
                                                             ws.block_counts);
        hm_pareto_scan<<<1, 1024, 0, st>>>(ws.block_counts, nblk, ws.round);
        unsigned* nxt = ws.alive[flip];
        hm_pareto_scatter<<<nblk, HM_PF_THREADS, 0, st>>>(cur, n_alive, ws.flags, ws.block_counts, nxt, done + b, ws.round);
        HM_CUDA_TRY(cudaGetLastError());
        HM_CUDA_TRY(cudaMemcpyAsync(&h, ws.round, sizeof(h), cudaMemcpyDeviceToHost, st));
        HM_CUDA_TRY(cudaStreamSynchronize(st));
        done = (done + b >= n_alive) ? h.total : h.new_done;
        n_alive = h.total;
        cur = nxt;
        flip ^= 1;
    }
    HM_CUDA_TRY(cudaMemsetAsync(mask, 0, (size_t)N, st));
    const unsigned S = n_alive;
    int64_t n_front = S;
    if (S > 0) {
        const unsigned g = (S + 255) / 256;
        if (pass2) {
            hm_pareto_init_state<<<g, 256, 0, st>>>(S, ws.state);
            hm_pareto_pass2_k<M><<<1, 1024, 0, st>>>(costs, cur, S, ws.state);
            hm_pareto_write_mask<<<g, 256, 0, st>>>(cur, ws.state, S, mask);
        } else {
            hm_pareto_write_mask<<<g, 256, 0, st>>>(cur, nullptr, S, mask);
        }
        HM_CUDA_TRY(cudaGetLastError());
    }
    if (n_out) {
        if (pass2 && S > 0) {
            // count the survivors of pass 2 on the host side of the (tiny) state vector
            std::vector<unsigned char> hs(S);
            HM_CUDA_TRY(cudaMemcpyAsync(hs.data(), ws.state, S, cudaMemcpyDeviceToHost, st));
            HM_CUDA_TRY(cudaStreamSynchronize(st));
            n_front = 0;
            for (unsigned i = 0; i < S; ++i) n_front += hs[i];
        }
        *n_out = n_front;
    }
    return 0;
}

static int hm_pareto_dispatch(const double* costs, int64_t N, int M, uint8_t* mask, void* workspace, int64_t* n_out,
                              bool pass2, void* stream) {
    HM_REQUIRE(N >= 0 && N < 2147483647ll, "0 <= N < 2^31");
    HM_REQUIRE(M >= 1 && M <= HM_MAX_OBJECTIVES, "1 <= M <= 8");
    if (N == 0) {
        if (n_out) *n_out = 0;
        return 0;
    }
    HM_REQUIRE(costs && mask && workspace, "costs/mask/workspace");
    cudaStream_t st = (cudaStream_t)stream;
    switch (M) {
        case 1: return hm_pareto_run<1>(costs, N, mask, workspace, n_out, pass2, st);
        case 2: return hm_pareto_run<2>(costs, N, mask, workspace, n_out, pass2, st);
        case 3: return hm_pareto_run<3>(costs, N, mask, workspace, n_out, pass2, st);
        case 4: return hm_pareto_run<4>(costs, N, mask, workspace, n_out, pass2, st);
        case 5: return hm_pareto_run<5>(costs, N, mask, workspace, n_out, pass2, st);
        case 6: return hm_pareto_run<6>(costs, N, mask, workspace, n_out, pass2, st);
        case 7: return hm_pareto_run<7>(costs, N, mask, workspace, n_out, pass2, st);
        default: return hm_pareto_run<8>(costs, N, mask, workspace, n_out, pass2, st);
    }
}

extern "C" int hm_pareto_filter(const double* costs, int64_t N, int M, uint8_t* mask, void* workspace,
                                int64_t* n_front_out_host, void* stream) {
    return hm_pareto_dispatch(costs, N, M, mask, workspace, n_front_out_host, true, stream);
}

extern "C" int hm_pareto_pass1(const double* costs, int64_t N, int M, uint8_t* mask, void* workspace,
                               int64_t* n_alive_out_host, void* stream) {
    return hm_pareto_dispatch(costs, N, M, mask, workspace, n_alive_out_host, false, stream);
}
