// K4: Pareto dominance filter = compute_pareto.sequential_is_pareto_efficient (compute_pareto.py:89-117).
//
// Pass 1 of the reference sweeps the points in index order; every point that is still efficient when its turn
// comes ("pivot") kills every efficient point j with  NOT any(c_j <= c_pivot),  i.e.  all(c_j > c_pivot).
//
// FAST PATH (NaN-free input).  Strict dominance in every coordinate is transitive, so the survivors of the sweep are
// exactly the points that NO other point kills -- independent of the order, and ANY subset of the points may be used
// as pivots to thin the set out first.  Two thinning levels, then an exact all-pairs test on what is left:
//   prep    : the exact front of a strided sample of the (remaining) points, strongest first = <= 1024 pivots
//             (hm_px_gather_sample + hm_px_prep: all-pairs through 1024-point shared-memory tiles, two candidates per
//             warp; the last CTA to finish sorts the front by coordinate sum).
//   stream  : every point is tested against the 8 strongest pivots unconditionally (branch-free), block survivors are
//             compacted in shared memory and tested warp-cooperatively against the remaining pivots; survivors'
//             indices are appended to a list.  Level 1 streams the cost matrix once at HBM speed (8 M + 1 bytes/point
//             algorithmic), level 2 re-filters the ~1 % survivors against pivots sampled from among them.
//   exact   : all-pairs on the level-2 survivors -> mask.
// Everything is sized by device-side counters, the chain is one CUDA graph, and its last node publishes the counters to
// mapped pinned memory: the only host interaction is the graph launch and one stream synchronisation.
// Pass 2 (compute_pareto.py:107-115; sequential and order dependent by definition): at point i's turn every later point
// still has its pass-1 state and every earlier point its final state, so  dead(i) = A(i) or B(i)  with
// A(i) = exists j > i: rel(j, i)  (fully parallel) and B(i) = exists j < i alive: rel(j, i)  (one CTA walks only the
// points that have such a candidate, in index order);  rel(j, i) = any(c_j == c_i) and any(c_j < c_i).  A front of <= 256
// points (the normal case) is sorted, tested and cleaned up by ONE CTA (hm_px_tail_small); up to 8192 points by the
// one-CTA sort + grid-wide A + one-CTA B; beyond that the host runs an ordered mask compaction first.
//
// EXACT PATH (input contains NaN: the kill relation is no longer transitive, the sweep order matters).  The sweep is
// reproduced literally, a batch of B pivot candidates at a time:
//   resolve : one CTA builds the BxB "i kills j" bit matrix of the batch in shared memory (tiled pairwise
//             dominance), then one warp replays the sequential sweep over the bit rows -> which batch members are
//             pivots (efficient at their turn).
//   filter  : every efficient point (already-processed pivots, the batch, the not-yet-reached rest) is tested
//             against the <= B pivots held in shared memory; survivors are compacted in index order (block
//             counts -> scan -> scatter), so the next batch is again "the next B efficient points in index order".
// The fast path raises a device flag when it meets a NaN and the call is redone on the exact path.
#include <stdlib.h>
#include <algorithm>
#include <mutex>
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
                                                       HmParetoRound* round, const unsigned* __restrict__ run_flag) {
    __shared__ unsigned warp_sums[32];
    if (run_flag && !*run_flag) return;
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
                                                                   HmParetoRound* round,
                                                                   const unsigned* __restrict__ run_flag) {
    __shared__ unsigned warp_sums[HM_PF_THREADS / 32];
    if (run_flag && !*run_flag) return;
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

// ============================================================================================================
// fast path kernels
// ============================================================================================================
#define HM_PX_THREADS 256
#define HM_PX_WARPS 8
#define HM_PX_TILE 1024             // points per shared-memory tile
#define HM_PIV_MAX 1024
#define HM_SAMPLE1 4096
#define HM_SAMPLE2 4096
#define HM_SORT_MAX 8192            // longest list the one-CTA sorts take
#define HM_PS_ITEMS1 4              // points per thread, level-1 stream
#define HM_PS_ITEMS2 1              // level 2: few points, spread them over many CTAs

struct HmParetoCtl {
    unsigned n_sample;      // size of the last level's sample
    unsigned n_piv_raw;     // front of the sample (unsorted)
    unsigned n_piv;         // pivots after sort / cap
    unsigned n_list1;       // survivors of level 1
    unsigned n_list2;       // survivors of level 2
    unsigned n_front1;      // pass-1 survivors
    unsigned nan_flag;
    unsigned n_dead2;       // points removed by pass 2
    unsigned need_scan;     // pass-1 front too long for the one-CTA index sort: the host runs the ordered-compaction tail
    unsigned tail_done;     // pass 2 was finished by hm_px_tail_small (front of <= HM_TAIL_SMALL points)
    unsigned prep_ticket;   // CTAs of hm_px_prep that are done (the last one sorts the pivots and resets it)
    unsigned pad[5];
};

__global__ void hm_px_ctl_init(HmParetoCtl* ctl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctl->n_sample = ctl->n_piv_raw = ctl->n_piv = 0;
        ctl->n_list1 = ctl->n_list2 = ctl->n_front1 = 0;
        ctl->nan_flag = ctl->n_dead2 = ctl->need_scan = ctl->tail_done = ctl->prep_ticket = 0;
    }
}

// Exact front of a point list: every candidate is tested against every point of the list (1024-point tiles staged in
// shared memory, SoA, two candidates per warp, lanes stride over the tile; a CTA leaves the tile loop when all its
// candidates are dead).  Every thread of the CTA must call it.
//   TO_MASK = false : survivors' costs are appended to out_costs (count in *out_count)                 [sample -> pivots]
//   TO_MASK = true  : mask[index] = 1 for survivors, their indices appended to (unsigned*)out_costs     [final exact step]
// (Measured alternatives that lost, profiles/r02_pareto_timeline_v3..v5.txt: four candidates per warp with the whole list
// staged at once -- half the CTAs, twice the serial work per warp, 28-32 us against 9-10; a warp-level first phase on 256
// points followed by the whole CTA on each candidate still alive -- three barriers per candidate, 18 us on a uniform
// sample and 36 us on the level-2 sample, whose points are mostly mutually non-dominated.)
#define HM_PX_CPW 2                 // candidates per warp
#define HM_PX_GROUP (HM_PX_WARPS * HM_PX_CPW)
template <int M, bool TO_MASK>
__device__ __forceinline__ void hm_px_front_core(const double* __restrict__ costs, const unsigned* __restrict__ idx,
                                                 unsigned count, double* __restrict__ out_costs,
                                                 unsigned char* __restrict__ mask, unsigned* __restrict__ out_count,
                                                 unsigned char* smem) {
    double (*tile)[HM_PX_TILE] = reinterpret_cast<double (*)[HM_PX_TILE]>(smem);   // [M][HM_PX_TILE]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned g0 = blockIdx.x * HM_PX_GROUP; g0 < count; g0 += gridDim.x * HM_PX_GROUP) {
        double cc[HM_PX_CPW][M];
        unsigned alive = 0;          // bit c: candidate c exists and is not killed yet (warp uniform)
#pragma unroll
        for (int c = 0; c < HM_PX_CPW; ++c) {
            const unsigned q = g0 + warp * HM_PX_CPW + c;
            const bool ok = q < count;
            const size_t src = ok ? (idx ? (size_t)idx[q] : (size_t)q) : 0;
#pragma unroll
            for (int d = 0; d < M; ++d) cc[c][d] = costs[src * M + d];
            if (ok) alive |= 1u << c;
        }
        for (unsigned t0 = 0; t0 < count; t0 += HM_PX_TILE) {
            if (!__syncthreads_or(alive != 0)) break;        // also: everyone is done with the previous tile
            const unsigned tn = min((unsigned)HM_PX_TILE, count - t0);
            for (unsigned t = threadIdx.x; t < tn; t += HM_PX_THREADS) {
                const size_t src = idx ? (size_t)idx[t0 + t] : (size_t)(t0 + t);
#pragma unroll
                for (int d = 0; d < M; ++d) tile[d][t] = costs[src * M + d];
            }
            __syncthreads();
            if (alive) {
                unsigned killed = 0;
                for (unsigned j = lane; j < tn; j += 32) {
                    double p[M];
#pragma unroll
                    for (int d = 0; d < M; ++d) p[d] = tile[d][j];
#pragma unroll
                    for (int c = 0; c < HM_PX_CPW; ++c)
                        if (hm_kills<M>(p, cc[c])) killed |= 1u << c;
                }
#pragma unroll
                for (int c = 0; c < HM_PX_CPW; ++c)
                    if (__any_sync(0xffffffffu, (killed >> c) & 1u)) alive &= ~(1u << c);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < HM_PX_CPW; ++c) {
                if (!((alive >> c) & 1u)) continue;
                const unsigned q = g0 + warp * HM_PX_CPW + c;
                if (TO_MASK) {
                    const unsigned gi = idx ? idx[q] : q;
                    mask[gi] = 1;
                    reinterpret_cast<unsigned*>(out_costs)[atomicAdd(out_count, 1u)] = gi;     // front list (indices)
                } else {
                    const unsigned pos = atomicAdd(out_count, 1u);
#pragma unroll
                    for (int d = 0; d < M; ++d) out_costs[(size_t)pos * M + d] = cc[c][d];
                }
            }
        }
        __syncthreads();
    }
}

// the final exact step: front of the level-2 survivors -> mask, and the front's indices into front_list
template <int M>
__global__ void __launch_bounds__(HM_PX_THREADS) hm_px_exact(const double* __restrict__ costs,
                                                             const unsigned* __restrict__ idx,
                                                             const unsigned* __restrict__ count_ptr,
                                                             unsigned* __restrict__ front_list, unsigned char* __restrict__ mask,
                                                             unsigned* __restrict__ out_count,
                                                             const unsigned* __restrict__ abort_flag) {
    extern __shared__ __align__(16) unsigned char exact_smem[];
    if (*abort_flag) return;        // a NaN was seen: the exact path will redo the call
    hm_px_front_core<M, true>(costs, idx, *count_ptr, reinterpret_cast<double*>(front_list), mask, out_count, exact_smem);
}

// Pivot preparation of one thinning level = the exact front of a strided sample of a point list, strongest first:
//   hm_px_gather_sample  the sample as a compact array [S][M] (point q of the sample = list position q * count / S);
//   hm_px_prep           its exact front (hm_px_front_core) into piv_raw; the last CTA to finish (ticket) sorts the front
//                        by coordinate sum (bitonic, shared memory) and writes the <= piv_max strongest as the level's
//                        pivots -- one launch where there were two (exact, sort_pivots).
template <int M>
__global__ void hm_px_gather_sample(const double* __restrict__ costs, const unsigned* __restrict__ idx,
                                    const unsigned* __restrict__ count_ptr, unsigned n_total, unsigned s_max,
                                    double* __restrict__ sample, HmParetoCtl* ctl) {
    const unsigned count = count_ptr ? *count_ptr : n_total;
    const unsigned S = min(s_max, count);
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) ctl->n_sample = S;
    if (i >= S) return;
    const unsigned pos = (unsigned)(((unsigned long long)i * count) / S);
    const size_t src = idx ? (size_t)idx[pos] : (size_t)pos;
#pragma unroll
    for (int d = 0; d < M; ++d) sample[(size_t)i * M + d] = costs[src * M + d];
}

template <int M>
__global__ void __launch_bounds__(HM_PX_THREADS) hm_px_prep(const double* __restrict__ sample, unsigned piv_max,
                                                            double* __restrict__ piv_raw, double* __restrict__ pivots,
                                                            HmParetoCtl* ctl) {
    extern __shared__ __align__(16) unsigned char prep_smem[];
    __shared__ unsigned s_last;
    if (!*(volatile unsigned*)&ctl->nan_flag)
        hm_px_front_core<M, false>(sample, nullptr, ctl->n_sample, piv_raw, nullptr, &ctl->n_piv_raw, prep_smem);
    // ---- the last CTA sorts ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ctl->prep_ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned n = *(volatile unsigned*)&ctl->n_piv_raw;
    double* key = reinterpret_cast<double*>(prep_smem);                               // [n2]
    unsigned short* ord = reinterpret_cast<unsigned short*>(key + HM_SAMPLE2);        // [n2]
    unsigned n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (unsigned i = threadIdx.x; i < n2; i += HM_PX_THREADS) {
        double sum = INFINITY;
        if (i < n) {
            sum = 0.0;
#pragma unroll
            for (int d = 0; d < M; ++d) sum += __ldcg(piv_raw + (size_t)i * M + d);
            if (sum != sum) sum = INFINITY;       // inf - inf
        }
        key[i] = sum;
    }
    __syncthreads();
    const unsigned P = min(n, piv_max);
    for (unsigned i = threadIdx.x; i < n2; i += HM_PX_THREADS) ord[i] = (unsigned short)i;
    __syncthreads();
    for (unsigned k = 2; k <= n2; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < n2; i += HM_PX_THREADS) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const double a = key[i], b = key[l];
                    const unsigned short oa = ord[i], ob = ord[l];
                    const bool gt = (a > b) || (a == b && oa > ob);
                    if (gt == up) {
                        key[i] = b;
                        key[l] = a;
                        ord[i] = ob;
                        ord[l] = oa;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (unsigned i = threadIdx.x; i < P; i += HM_PX_THREADS) {
        const unsigned src = ord[i];
#pragma unroll
        for (int d = 0; d < M; ++d) pivots[(size_t)i * M + d] = __ldcg(piv_raw + (size_t)src * M + d);
    }
    if (threadIdx.x == 0) {
        ctl->n_piv = P;
        ctl->n_piv_raw = 0;          // ready for the next level
        ctl->prep_ticket = 0;
    }
}

// streaming filter of a point list against the sorted pivots; survivors' indices are appended to out_list.
// Warp-autonomous (no block barrier after the pivot staging), so every warp's DRAM latency overlaps the others' work:
//   stage 1: the K1 (2..8, by M) strongest pivots live in registers; every point of the warp's 32 x ITEMS sub-chunk is
//            tested against all of them (branch-free); all loads of the sub-chunk are issued up front.
//   stage 2: stage-1 survivors queue up in a per-warp shared-memory buffer; whenever 32 are waiting (and at the end)
//            they walk the remaining pivots, one per lane (shared memory, broadcast reads), until killed.
//   output : stage-2 survivors queue up in a second per-warp buffer, flushed 32 at a time with one global atomic.
#define HM_PS_Q1 (32 + 32 * 8)
#define HM_PS_Q2 64
template <int M, int ITEMS>
__global__ void __launch_bounds__(HM_PX_THREADS, 3) hm_px_stream(const double* __restrict__ costs,
                                                              const unsigned* __restrict__ idx,
                                                              const unsigned* __restrict__ count_ptr, unsigned n_total,
                                                              const double* __restrict__ pivots, HmParetoCtl* ctl,
                                                              unsigned* __restrict__ out_list,
                                                              unsigned* __restrict__ out_count) {
    static_assert(ITEMS <= 8, "queue sizing");
    // register-resident pivots: as many as ~36 registers hold
    constexpr int K1 = M <= 2 ? 8 : (M == 3 ? 6 : (M == 4 ? 4 : 2));
    extern __shared__ __align__(16) unsigned char stream_smem[];
    double* sp = reinterpret_cast<double*>(stream_smem);                    // [M][HM_PIV_MAX]  (SoA)
    unsigned* qall = reinterpret_cast<unsigned*>(sp + (size_t)M * HM_PIV_MAX);
    if (*(volatile unsigned*)&ctl->nan_flag) return;      // uniform enough: the whole call is void once the flag is up
    const unsigned count = count_ptr ? *count_ptr : n_total;
    const unsigned P = ctl->n_piv;
    for (unsigned i = threadIdx.x; i < P * M; i += HM_PX_THREADS) {
        const unsigned p = i / M, d = i - p * M;
        sp[(size_t)d * HM_PIV_MAX + p] = pivots[i];
    }
    double pv[K1][M];                            // +inf padding never kills
#pragma unroll
    for (int p = 0; p < K1; ++p)
#pragma unroll
        for (int d = 0; d < M; ++d) pv[p][d] = ((unsigned)p < P) ? __ldg(pivots + (size_t)p * M + d) : INFINITY;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* q1 = qall + warp * (HM_PS_Q1 + HM_PS_Q2);   // stage-1 survivors waiting for stage 2
    unsigned* q2 = q1 + HM_PS_Q1;                          // stage-2 survivors waiting for the flush
    unsigned n1 = 0, n2 = 0;                               // queue fill (warp uniform)
    bool saw_nan = false;

    auto flush = [&](unsigned m) {                         // write the first m (<= 32) entries of q2 to the list
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(out_count, m);
        base = __shfl_sync(0xffffffffu, base, 0);
        HM_DEV_ASSERT(m <= 32u && m <= n2 && n2 <= (unsigned)HM_PS_Q2 && (unsigned long long)base + m <= (unsigned long long)count);
        if ((unsigned)lane < m) out_list[base + lane] = q2[lane];
        __syncwarp();
        if ((unsigned)lane + 32 < n2) {                    // move the tail down (n2 <= 64)
            const unsigned v = q2[lane + 32];
            q2[lane] = v;
        }
        __syncwarp();
        n2 -= m;
    };
    auto stage2 = [&](unsigned m) {                        // the LAST m (<= 32) entries of q1 walk the remaining pivots
        const bool have = (unsigned)lane < m;
        const unsigned sidx = have ? q1[n1 - m + lane] : 0;
        double c2[M];
#pragma unroll
        for (int d = 0; d < M; ++d) c2[d] = have ? __ldg(costs + (size_t)sidx * M + d) : 0.0;
        bool dead = !have;
        if (ITEMS == HM_PS_ITEMS2) {
            // level 2 (few points, long pivot list, most of the batch survives for a long time): the batch's points one
            // after the other, the LANES over the pivots -- a survivor costs P / 32 iterations instead of P, a killed
            // point usually one (the pivots are sorted strongest first)
            unsigned deadmask = 0;                         // warp uniform
            for (unsigned q = 0; q < m; ++q) {
                double cq[M];
#pragma unroll
                for (int d = 0; d < M; ++d) cq[d] = __shfl_sync(0xffffffffu, c2[d], (int)q);
                for (unsigned p0 = K1; p0 < P; p0 += 32) {
                    const unsigned p = p0 + lane;
                    bool all = p < P;
#pragma unroll
                    for (int d = 0; d < M; ++d) all = all && (cq[d] > sp[(size_t)d * HM_PIV_MAX + (p < P ? p : 0)]);
                    if (__any_sync(0xffffffffu, all)) {
                        deadmask |= 1u << q;
                        break;
                    }
                }
            }
            dead = dead || ((deadmask >> lane) & 1u);
        } else {
        for (unsigned p = K1; p < P; ++p) {
            if (__all_sync(0xffffffffu, dead)) break;
            bool all = true;
#pragma unroll
            for (int d = 0; d < M; ++d) all = all && (c2[d] > sp[(size_t)d * HM_PIV_MAX + p]);
            dead = dead || all;
        }
        }
        n1 -= m;
        const unsigned ballot = __ballot_sync(0xffffffffu, !dead);
        HM_DEV_ASSERT(n2 + __popc(ballot) <= (unsigned)HM_PS_Q2 && m <= 32u);
        if (!dead) q2[n2 + __popc(ballot & ((1u << lane) - 1u))] = sidx;
        n2 += __popc(ballot);
        __syncwarp();
        if (n2 >= 32) flush(32);
    };

    const unsigned gw = blockIdx.x * HM_PX_WARPS + warp, nw = gridDim.x * HM_PX_WARPS;
    constexpr unsigned SUB = 32 * ITEMS;
    for (unsigned long long base = (unsigned long long)gw * SUB; base < count; base += (unsigned long long)nw * SUB) {
        double cj[ITEMS][M];
        unsigned src[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const unsigned long long pos = base + r * 32 + lane;
            const unsigned pc = pos < count ? (unsigned)pos : count - 1;
            src[r] = idx ? idx[pc] : pc;
        }
#pragma unroll
        for (int r = 0; r < ITEMS; ++r)
#pragma unroll
            for (int d = 0; d < M; ++d) cj[r][d] = __ldg(costs + (size_t)src[r] * M + d);
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const unsigned long long pos = base + r * 32 + lane;
            double sum = cj[r][0];
#pragma unroll
            for (int d = 1; d < M; ++d) sum += cj[r][d];
            saw_nan = saw_nan || (sum != sum);             // (+inf) + (-inf) also lands here: merely takes the exact path
            bool dead = false;
#pragma unroll
            for (int p = 0; p < K1; ++p) {
                bool all = true;
#pragma unroll
                for (int d = 0; d < M; ++d) all = all && (cj[r][d] > pv[p][d]);
                dead = dead || all;
            }
            const bool keep = !dead && pos < count;
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            HM_DEV_ASSERT(n1 + __popc(ballot) <= (unsigned)HM_PS_Q1);
            if (keep) q1[n1 + __popc(ballot & ((1u << lane) - 1u))] = src[r];
            n1 += __popc(ballot);
        }
        __syncwarp();
        while (n1 >= 32) stage2(32);
    }
    if (n1) stage2(n1);
    while (n2) flush(n2 < 32 ? n2 : 32);
    if (saw_nan) ctl->nan_flag = 1;
}

// sort the (usually tiny) list of pass-1 survivors by index so that pass 2 can walk it in index order.  One CTA.
// Lists longer than HM_SORT_MAX are left to the ordered mask compaction (count -> scan -> scatter).
__global__ void __launch_bounds__(1024, 1) hm_px_sort_front(unsigned* __restrict__ list, HmParetoCtl* ctl) {
    __shared__ unsigned key[HM_SORT_MAX];
    const unsigned n = ctl->n_front1;
    if (ctl->nan_flag || ctl->tail_done) return;
    if (n > HM_SORT_MAX) {
        if (threadIdx.x == 0) ctl->need_scan = 1;
        return;
    }
    unsigned n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (unsigned i = threadIdx.x; i < n2; i += 1024) key[i] = i < n ? list[i] : 0xffffffffu;
    __syncthreads();
    for (unsigned k = 2; k <= n2; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < n2; i += 1024) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned a = key[i], b = key[l];
                    if ((a > b) == up) {
                        key[i] = b;
                        key[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (unsigned i = threadIdx.x; i < n; i += 1024) list[i] = key[i];
}

// per-block counts of set flags (feeds hm_pareto_scan / hm_pareto_scatter: ordered compaction of the mask)
__global__ void __launch_bounds__(HM_PF_THREADS) hm_px_count_flags(const unsigned char* __restrict__ flags, unsigned n,
                                                                   unsigned* __restrict__ block_counts,
                                                                   const unsigned* __restrict__ run_flag) {
    __shared__ unsigned blk;
    if (run_flag && !*run_flag) return;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
    const unsigned base = blockIdx.x * HM_PF_CHUNK;
    unsigned kept = 0;
#pragma unroll
    for (int r = 0; r < HM_PF_ITEMS; ++r) {
        const unsigned pos = base + r * HM_PF_THREADS + threadIdx.x;
        if (pos < n) kept += flags[pos] ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_down_sync(0xffffffffu, kept, o);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&blk, kept);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = blk;
}

template <int M>
__device__ __forceinline__ bool hm_rel2(const double* __restrict__ cj, const double* __restrict__ ci) {
    bool eq = false, lt = false;
#pragma unroll
    for (int d = 0; d < M; ++d) {
        eq = eq || (cj[d] == ci[d]);
        lt = lt || (cj[d] < ci[d]);
    }
    return eq && lt;
}

// pass 2, parallel part: state[i] = 0 (killed by a later point), 2 (a earlier point may kill it), 1 (efficient)
template <int M>
__global__ void __launch_bounds__(HM_PX_THREADS) hm_px_pass2_a(const double* __restrict__ costs,
                                                               const unsigned* __restrict__ sorted,
                                                               const unsigned* __restrict__ count_ptr,
                                                               unsigned char* __restrict__ state,
                                                               const HmParetoCtl* __restrict__ ctl, int forced) {
    // in the graph: nothing to do when the small tail finished pass 2, or when the front was too long to sort in one
    // CTA (the host then runs the ordered compaction and calls this kernel again, forced)
    if (ctl->nan_flag || (!forced && (ctl->tail_done || ctl->need_scan))) return;
    const unsigned S = *count_ptr;
    const int lane = threadIdx.x & 31;
    const unsigned gw = (blockIdx.x * HM_PX_THREADS + threadIdx.x) >> 5, nw = (gridDim.x * HM_PX_THREADS) >> 5;
    for (unsigned i = gw; i < S; i += nw) {
        double ci[M];
        const size_t si = sorted[i];
#pragma unroll
        for (int d = 0; d < M; ++d) ci[d] = __ldg(costs + si * M + d);
        bool later = false, earlier = false;
        for (unsigned j0 = 0; j0 < S; j0 += 32) {
            const unsigned j = j0 + lane;
            if (j < S) {
                double cj[M];
                const size_t sj = sorted[j];
#pragma unroll
                for (int d = 0; d < M; ++d) cj[d] = __ldg(costs + sj * M + d);
                if (hm_rel2<M>(cj, ci)) {
                    later = later || (j > i);
                    earlier = earlier || (j < i);
                }
            }
            if (__any_sync(0xffffffffu, later)) break;
        }
        const bool a = __any_sync(0xffffffffu, later), b = __any_sync(0xffffffffu, earlier);
        if (lane == 0) state[i] = a ? 0 : (b ? 2 : 1);
    }
}

// pass 2, sequential part (one CTA): resolve the state-2 points in index order, clear the mask of the dead.
template <int M>
__global__ void __launch_bounds__(1024, 1) hm_px_pass2_b(const double* __restrict__ costs,
                                                         const unsigned* __restrict__ sorted,
                                                         const unsigned* __restrict__ count_ptr,
                                                         volatile unsigned char* state, unsigned char* __restrict__ mask,
                                                         HmParetoCtl* ctl, int forced) {
    if (ctl->nan_flag || (!forced && (ctl->tail_done || ctl->need_scan))) return;
    const unsigned S = *count_ptr;
    __shared__ unsigned pending[1024];
    __shared__ unsigned n_pending;
    unsigned dead_count = 0;
    for (unsigned c0 = 0; c0 < S; c0 += 1024) {
        if (threadIdx.x == 0) n_pending = 0;
        __syncthreads();
        // ordered list of the undecided points of this chunk
        const unsigned i = c0 + threadIdx.x;
        const bool und = i < S && state[i] == 2;
        const unsigned ballot = __ballot_sync(0xffffffffu, und);
        __shared__ unsigned wcount[32];
        if ((threadIdx.x & 31) == 0) wcount[threadIdx.x >> 5] = __popc(ballot);
        __syncthreads();
        unsigned before = 0;
        for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) before += wcount[w];
        if (und) pending[before + __popc(ballot & ((1u << (threadIdx.x & 31)) - 1u))] = i;
        if (threadIdx.x == 1023) n_pending = before + __popc(ballot);
        __syncthreads();
        const unsigned np = n_pending;
        for (unsigned q = 0; q < np; ++q) {
            const unsigned ii = pending[q];
            double ci[M];
            const size_t si = sorted[ii];
#pragma unroll
            for (int d = 0; d < M; ++d) ci[d] = costs[si * M + d];
            int found = 0;
            for (unsigned j = threadIdx.x; j < ii && !found; j += 1024) {
                if (state[j] != 1) continue;
                double cj[M];
                const size_t sj = sorted[j];
#pragma unroll
                for (int d = 0; d < M; ++d) cj[d] = costs[sj * M + d];
                found = hm_rel2<M>(cj, ci) ? 1 : 0;
            }
            const int kill = __syncthreads_or(found);
            if (threadIdx.x == 0) state[ii] = kill ? 0 : 1;
            __syncthreads();
        }
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < S; i += 1024) {
        if (state[i] == 0) {
            mask[sorted[i]] = 0;
            ++dead_count;
        }
    }
    for (int o = 16; o > 0; o >>= 1) dead_count += __shfl_down_sync(0xffffffffu, dead_count, o);
    if ((threadIdx.x & 31) == 0 && dead_count) atomicAdd(&ctl->n_dead2, dead_count);
}


// Pass 2 on a short pass-1 front (the normal case: a front has a few hundred points), everything in ONE CTA: rank sort
// of the front's indices, the costs staged in shared memory, A(i) / B(i) candidates all-pairs, the rare sequential
// resolution of the undecided points, mask clean-up.  Replaces sort_front + pass2_a + pass2_b (+ the three launches of
// the ordered compaction, which the host now runs only when the front is longer than HM_SORT_MAX).
#define HM_TAIL_SMALL 256
template <int M>
__global__ void __launch_bounds__(HM_TAIL_SMALL, 1) hm_px_tail_small(const double* __restrict__ costs,
                                                                     unsigned* __restrict__ list,
                                                                     unsigned char* __restrict__ mask, HmParetoCtl* ctl) {
    __shared__ unsigned key[HM_TAIL_SMALL], sorted[HM_TAIL_SMALL];
    __shared__ double c[M][HM_TAIL_SMALL];
    __shared__ unsigned char stt[HM_TAIL_SMALL];
    __shared__ unsigned n_dead;
    if (ctl->nan_flag) return;
    const unsigned n = ctl->n_front1;
    if (n > HM_TAIL_SMALL) return;          // sort_front / pass2_a / pass2_b (grid-wide) take it
    const unsigned t = threadIdx.x;
    if (n == 0) {
        if (t == 0) ctl->tail_done = 1;
        return;
    }
    if (t == 0) n_dead = 0;
    if (t < n) key[t] = list[t];
    __syncthreads();
    if (t < n) {                            // rank sort: the indices are distinct (one barrier; a bitonic network's 28
        const unsigned k = key[t];          // barriers for 128 entries cost more than the n reads per thread)
        unsigned r = 0;
        for (unsigned j = 0; j < n; ++j) r += key[j] < k ? 1u : 0u;
        HM_DEV_ASSERT(r < n);
        sorted[r] = k;
    }
    __syncthreads();
    double ci[M];
#pragma unroll
    for (int d = 0; d < M; ++d) ci[d] = 0.0;
    if (t < n) {
        const size_t si = sorted[t];
        list[t] = (unsigned)si;             // the front list is left in index order, as hm_px_sort_front leaves it
#pragma unroll
        for (int d = 0; d < M; ++d) {
            ci[d] = __ldg(costs + si * M + d);
            c[d][t] = ci[d];
        }
    }
    __syncthreads();
    bool later = false, earlier = false;
    if (t < n) {
        for (unsigned j = 0; j < n; ++j) {
            double cj[M];
#pragma unroll
            for (int d = 0; d < M; ++d) cj[d] = c[d][j];
            if (hm_rel2<M>(cj, ci)) {
                later = later || (j > t);
                earlier = earlier || (j < t);
            }
        }
    }
    const unsigned char s0 = later ? 0 : (earlier ? 2 : 1);
    if (t < n) stt[t] = s0;
    // B(i): the undecided points (they share a coordinate value with an earlier point: rare) in index order
    if (__syncthreads_or(t < n && s0 == 2)) {
        for (unsigned i = 0; i < n; ++i) {
            if (stt[i] != 2) continue;      // uniform: shared memory, written behind a barrier
            bool found = false;
            if (t < i && stt[t] == 1) {
                double cI[M];
#pragma unroll
                for (int d = 0; d < M; ++d) cI[d] = c[d][i];
                found = hm_rel2<M>(ci, cI);
            }
            const int kill = __syncthreads_or(found);
            if (t == 0) stt[i] = kill ? 0 : 1;
            __syncthreads();
        }
    }
    if (t < n && stt[t] == 0) {
        mask[sorted[t]] = 0;
        atomicAdd(&n_dead, 1u);
    }
    __syncthreads();
    if (t == 0) {
        ctl->n_dead2 = n_dead;
        ctl->tail_done = 1;
    }
}

// last node of the chain: the control block goes to the host through mapped pinned memory (no D2H copy behind the graph)
__global__ void hm_px_publish(const HmParetoCtl* __restrict__ ctl, volatile unsigned* host_result) {
    const unsigned t = threadIdx.x;
    if (t < sizeof(HmParetoCtl) / 4) host_result[t] = reinterpret_cast<const unsigned*>(ctl)[t];
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------------------
struct HmParetoWs {
    unsigned* alive[2];          // exact path: ping-pong index lists; fast path: level-1 / level-2 survivor lists
    unsigned char* flags;
    unsigned* block_counts;
    double* pivots;              // exact path: batch pivots; fast path: sorted pivots   [HM_PB][M]
    unsigned char* state;
    HmParetoRound* round;
    double* sample;              // [HM_SAMPLE2][M]
    double* piv_raw;             // [HM_SAMPLE2][M]
    HmParetoCtl* ctl;
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
    unsigned char* sm = take((size_t)HM_SAMPLE2 * M * 8);
    unsigned char* pr = take((size_t)HM_SAMPLE2 * M * 8);
    unsigned char* ct = take(sizeof(HmParetoCtl));
    if (ws) {
        ws->alive[0] = (unsigned*)a0;
        ws->alive[1] = (unsigned*)a1;
        ws->flags = fl;
        ws->block_counts = (unsigned*)bc;
        ws->pivots = (double*)pv;
        ws->state = stt;
        ws->round = (HmParetoRound*)rd;
        ws->sample = (double*)sm;
        ws->piv_raw = (double*)pr;
        ws->ctl = (HmParetoCtl*)ct;
    }
    return off;
}

extern "C" int64_t hm_pareto_workspace_bytes(int64_t N, int M) {
    if (M < 1) M = 1;
    return (int64_t)hm_pareto_layout(N, M, nullptr, nullptr) + 256;
}

template <int M>
static int hm_pareto_run_exact(const double* costs, int64_t N, uint8_t* mask, void* workspace, int64_t* n_out, bool pass2,
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
        hm_pareto_scan<<<1, 1024, 0, st>>>(ws.block_counts, nblk, ws.round, nullptr);
        unsigned* nxt = ws.alive[flip];
        hm_pareto_scatter<<<nblk, HM_PF_THREADS, 0, st>>>(cur, n_alive, ws.flags, ws.block_counts, nxt, done + b, ws.round, nullptr);
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

// ---- tiny cache of instantiated launch-chain graphs -----------------------------------------------------------
struct HmGraphKey {
    const void* costs;
    const void* mask;
    const void* workspace;
    int64_t N;
    int M, pass2;
};
#define HM_GRAPH_SLOTS 8
static struct {
    HmGraphKey key;
    cudaGraphExec_t exec;
    int device;
} g_graphs[HM_GRAPH_SLOTS];
static int g_graph_next = 0, g_graph_off = -1;
static cudaStream_t g_graph_stream = nullptr;
static std::mutex g_graph_mutex;
// one mapped pinned control block per graph slot: the chain's last node (hm_px_publish) writes the counters there, the
// host reads them after the stream synchronisation -- no D2H copy node behind the graph
static HmParetoCtl* g_results = nullptr;

static bool hm_graph_enabled() {
    if (g_graph_off < 0) {
        const char* e = getenv("HM_B200_PARETO_GRAPH");
        g_graph_off = (e && atoi(e) == 0) ? 1 : 0;
    }
    return g_graph_off == 0;
}
static void hm_graph_disable() { g_graph_off = 1; }
static cudaStream_t hm_graph_capture_stream() {     // the caller's stream may be the legacy stream, which cannot capture
    std::lock_guard<std::mutex> lk(g_graph_mutex);
    if (!g_graph_stream && cudaStreamCreateWithFlags(&g_graph_stream, cudaStreamNonBlocking) != cudaSuccess)
        g_graph_stream = nullptr;
    return g_graph_stream;
}
static HmParetoCtl* hm_graph_results() {
    std::lock_guard<std::mutex> lk(g_graph_mutex);
    if (!g_results) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, sizeof(HmParetoCtl) * HM_GRAPH_SLOTS, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        g_results = (HmParetoCtl*)p;
    }
    return g_results;
}
// -> slot of the instantiated graph for this key on the current device, or -1
static int hm_graph_lookup(const HmGraphKey& k, cudaGraphExec_t* exec) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_graph_mutex);
    for (int i = 0; i < HM_GRAPH_SLOTS; ++i)
        if (g_graphs[i].exec && g_graphs[i].device == dev && g_graphs[i].key.costs == k.costs &&
            g_graphs[i].key.mask == k.mask && g_graphs[i].key.workspace == k.workspace && g_graphs[i].key.N == k.N &&
            g_graphs[i].key.M == k.M && g_graphs[i].key.pass2 == k.pass2) {
            *exec = g_graphs[i].exec;
            return i;
        }
    return -1;
}
// the slot the next graph will live in (its old graph, if any, is destroyed: the slot's result block is about to be
// baked into a new chain)
static int hm_graph_reserve() {
    std::lock_guard<std::mutex> lk(g_graph_mutex);
    const int i = g_graph_next;
    g_graph_next = (g_graph_next + 1) % HM_GRAPH_SLOTS;
    if (g_graphs[i].exec) cudaGraphExecDestroy(g_graphs[i].exec);
    g_graphs[i].exec = nullptr;
    return i;
}
static void hm_graph_store(int slot, const HmGraphKey& k, cudaGraphExec_t exec) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_graph_mutex);
    g_graphs[slot].key = k;
    g_graphs[slot].exec = exec;
    g_graphs[slot].device = dev;
}

// kernel attributes and occupancy of the fast path: set / queried once per (M, device)
template <int M>
struct HmPxSetup {
    static unsigned done_mask;       // bit per device; a benign race repeats the same calls
    static int occ1[32];
};
template <int M> unsigned HmPxSetup<M>::done_mask = 0;
template <int M> int HmPxSetup<M>::occ1[32] = {0};

// NaN-free fast path; *saw_nan_out = 1 means the result is void and the exact path must run.
template <int M>
static int hm_pareto_run_fast(const double* costs, int64_t N, uint8_t* mask, void* workspace, int64_t* n_out, bool pass2,
                              cudaStream_t st, int* saw_nan_out) {
    HmParetoWs ws;
    constexpr int IT1 = (M <= 4) ? HM_PS_ITEMS1 : 2;      // level-1 points per lane (register budget)
    unsigned char* base = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    hm_pareto_layout(N, M, base, &ws);
    static_assert(HM_PIV_MAX <= HM_PB, "sorted pivots share the exact path's pivot buffer");
    const size_t stream_smem1 = sizeof(double) * M * HM_PIV_MAX + sizeof(unsigned) * HM_PX_WARPS * (HM_PS_Q1 + HM_PS_Q2);
    const size_t stream_smem2 = stream_smem1;
    static_assert(HM_SAMPLE1 == HM_SAMPLE2, "one pivot-preparation geometry for both levels");
    const size_t exact_smem = sizeof(double) * M * HM_PX_TILE;
    // 1024-point tiles; the last CTA's sort needs 10 B per sample point
    const size_t prep_smem = std::max<size_t>(exact_smem, (size_t)HM_SAMPLE1 * 10);
    const unsigned prep_grid = (HM_SAMPLE1 + HM_PX_GROUP - 1) / HM_PX_GROUP;
    // level-1 pivot cap (experiments: HM_B200_PARETO_PIV1).  Measured on C4 (profiles/r02_exp_pareto3.log): 1024 / 128 / 64 /
    // 32 / 16 pivots -> level-1 stream 54.9 / 55.0 / 55.8 / 54.2 / 49.5 us, but level-2 stream 17.1 / 15.5 / 15.3 / 17.3 /
    // 25.6 us: what the weaker pivots leave alive comes back at level 2, so the cap stays at HM_PIV_MAX.
    unsigned piv1_max = HM_PIV_MAX;
    if (const char* e = getenv("HM_B200_PARETO_PIV1")) piv1_max = (unsigned)std::min(std::max(atoi(e), 8), HM_PIV_MAX);
    int dev = 0;
    HM_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 32 || !(HmPxSetup<M>::done_mask >> dev & 1u)) {
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_px_exact<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exact_smem));
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_px_prep<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_px_stream<M, IT1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem1));
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_px_stream<M, HM_PS_ITEMS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem2));
        int occ = 0;
        HM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hm_px_stream<M, IT1>, HM_PX_THREADS, stream_smem1));
        if (dev < 32) {
            HmPxSetup<M>::occ1[dev] = occ < 1 ? 1 : occ;
            HmPxSetup<M>::done_mask |= 1u << dev;
        }
    }
    const unsigned n = (unsigned)N;
    const int sms = hm_num_sms();
    int occ1 = dev < 32 ? HmPxSetup<M>::occ1[dev] : 1;
    if (occ1 < 1) occ1 = 1;
    HmParetoCtl* ctl = ws.ctl;
    unsigned* list1 = ws.alive[0];
    unsigned* list2 = ws.alive[1];

    // the whole launch chain (every size a device-side counter) as a function of the stream; `result` (nullable): mapped
    // host block the last node publishes the counters to
    auto enqueue = [&](cudaStream_t s, HmParetoCtl* result) -> int {
    hm_px_ctl_init<<<1, 32, 0, s>>>(ctl);
    HM_CUDA_TRY(cudaMemsetAsync(mask, 0, (size_t)N, s));
    // ---- level 1: pivots from a strided sample of everything, stream the whole cost matrix ----
    hm_px_gather_sample<M><<<(HM_SAMPLE1 + 255) / 256, 256, 0, s>>>(costs, nullptr, nullptr, n, HM_SAMPLE1, ws.sample, ctl);
    hm_px_prep<M><<<prep_grid, HM_PX_THREADS, prep_smem, s>>>(ws.sample, piv1_max, ws.piv_raw, ws.pivots, ctl);
    {
        const unsigned chunk = HM_PX_THREADS * IT1;
        const unsigned chunks = (n + chunk - 1) / chunk;
        // exactly one resident wave: a 4th CTA per SM would run after the first three at a third of the occupancy (and
        // of the bytes in flight) -- ncu: 1.33 waves, 47 % of HBM
        const unsigned grid = std::min<unsigned>(chunks, (unsigned)(sms * occ1));
        hm_px_stream<M, IT1><<<grid, HM_PX_THREADS, stream_smem1, s>>>(costs, nullptr, nullptr, n, ws.pivots, ctl,
                                                                              list1, &ctl->n_list1);
    }
    // ---- level 2: pivots sampled from the survivors, re-filter the survivors ----
    hm_px_gather_sample<M><<<(HM_SAMPLE2 + 255) / 256, 256, 0, s>>>(costs, list1, &ctl->n_list1, 0, HM_SAMPLE2, ws.sample, ctl);
    hm_px_prep<M><<<prep_grid, HM_PX_THREADS, prep_smem, s>>>(ws.sample, HM_PIV_MAX, ws.piv_raw, ws.pivots, ctl);
    hm_px_stream<M, HM_PS_ITEMS2><<<sms * 4, HM_PX_THREADS, stream_smem2, s>>>(costs, list1, &ctl->n_list1, 0, ws.pivots,
                                                                             ctl, list2, &ctl->n_list2);
    // ---- exact all-pairs on what is left -> mask, and the front's indices into list1 ----
    hm_px_exact<M><<<sms * 4, HM_PX_THREADS, exact_smem, s>>>(costs, list2, &ctl->n_list2, list1, mask,
                                                                     &ctl->n_front1, &ctl->nan_flag);
    HM_CUDA_TRY(cudaGetLastError());
    if (pass2) {
        // the order-dependent clean-up pass.  A front of <= HM_TAIL_SMALL points (the normal case) is finished by one CTA
        // (hm_px_tail_small); up to HM_SORT_MAX points: one-CTA index sort, A(i) grid-wide, B(i) in one CTA (the three
        // kernels return at once when the small tail did the job); longer: sort_front raises need_scan and the HOST runs
        // the ordered mask compaction + pass 2 behind the graph (below).
        hm_px_tail_small<M><<<1, HM_TAIL_SMALL, 0, s>>>(costs, list1, mask, ctl);
        hm_px_sort_front<<<1, 1024, 0, s>>>(list1, ctl);
        hm_px_pass2_a<M><<<sms * 4, HM_PX_THREADS, 0, s>>>(costs, list1, &ctl->n_front1, ws.state, ctl, 0);
        hm_px_pass2_b<M><<<1, 1024, 0, s>>>(costs, list1, &ctl->n_front1, ws.state, mask, ctl, 0);
        HM_CUDA_TRY(cudaGetLastError());
    }
    if (result) {
        hm_px_publish<<<1, 32, 0, s>>>(ctl, reinterpret_cast<volatile unsigned*>(result));
        HM_CUDA_TRY(cudaGetLastError());
    }
        return 0;
    };
    // The chain costs more CPU time to launch than the GPU needs to run it (10^7 x 3: ~0.15 ms of work), so it is
    // captured once per (buffers, N, pass 2) into a CUDA graph and replayed with one launch.
    HmParetoCtl h;
    {
        HmGraphKey key = {costs, mask, workspace, N, M, pass2 ? 1 : 0};
        cudaGraphExec_t exec = nullptr;
        int slot = hm_graph_lookup(key, &exec);
        HmParetoCtl* results = hm_graph_enabled() ? hm_graph_results() : nullptr;
        if (slot < 0 && results) {
            cudaStream_t cap = hm_graph_capture_stream();
            cudaGraph_t graph = nullptr;
            slot = hm_graph_reserve();
            if (cap && cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int rc = enqueue(cap, results + slot);
                const cudaError_t e = cudaStreamEndCapture(cap, &graph);
                if (rc == 0 && e == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess)
                    hm_graph_store(slot, key, exec);
                else
                    exec = nullptr;
                if (graph) cudaGraphDestroy(graph);
            }
            if (!exec) {
                cudaGetLastError();
                hm_graph_disable();
                slot = -1;
            }
        }
        if (exec && slot >= 0) {
            HM_CUDA_TRY(cudaGraphLaunch(exec, st));
            HM_CUDA_TRY(cudaStreamSynchronize(st));
            h = results[slot];
        } else {
            const int rc = enqueue(st, nullptr);
            if (rc) return rc;
            HM_CUDA_TRY(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
            HM_CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    if (pass2 && !h.nan_flag && h.need_scan) {
        // pass-1 front longer than HM_SORT_MAX: survivors in index order by ordered mask compaction, then pass 2
        const unsigned nblk = (n + HM_PF_CHUNK - 1) / HM_PF_CHUNK;
        hm_px_count_flags<<<nblk, HM_PF_THREADS, 0, st>>>(mask, n, ws.block_counts, nullptr);
        hm_pareto_scan<<<1, 1024, 0, st>>>(ws.block_counts, nblk, ws.round, nullptr);
        hm_pareto_scatter<<<nblk, HM_PF_THREADS, 0, st>>>(nullptr, n, mask, ws.block_counts, list1, 0xffffffffu, ws.round,
                                                          nullptr);
        hm_px_pass2_a<M><<<sms * 4, HM_PX_THREADS, 0, st>>>(costs, list1, &ctl->n_front1, ws.state, ctl, 1);
        hm_px_pass2_b<M><<<1, 1024, 0, st>>>(costs, list1, &ctl->n_front1, ws.state, mask, ctl, 1);
        HM_CUDA_TRY(cudaGetLastError());
        HM_CUDA_TRY(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
        HM_CUDA_TRY(cudaStreamSynchronize(st));
    }
    *saw_nan_out = h.nan_flag ? 1 : 0;
    if (n_out) *n_out = (int64_t)h.n_front1 - (int64_t)h.n_dead2;
    return 0;
}

template <int M>
static int hm_pareto_run(const double* costs, int64_t N, uint8_t* mask, void* workspace, int64_t* n_out, bool pass2,
                         cudaStream_t st, int* saw_nan_out) {
    int saw_nan = 0;
    int rc = hm_pareto_run_fast<M>(costs, N, mask, workspace, n_out, pass2, st, &saw_nan);
    if (saw_nan_out) *saw_nan_out = saw_nan;
    if (rc || !saw_nan) return rc;
    return hm_pareto_run_exact<M>(costs, N, mask, workspace, n_out, pass2, st);
}

static int hm_pareto_dispatch(const double* costs, int64_t N, int M, uint8_t* mask, void* workspace, int64_t* n_out,
                              bool pass2, void* stream, int* saw_nan_out = nullptr) {
    HM_REQUIRE(N >= 0 && N < 2147483647ll, "0 <= N < 2^31");
    HM_REQUIRE(M >= 1 && M <= HM_MAX_OBJECTIVES, "1 <= M <= 8");
    if (saw_nan_out) *saw_nan_out = 0;
    if (N == 0) {
        if (n_out) *n_out = 0;
        return 0;
    }
    HM_REQUIRE(costs && mask && workspace, "costs/mask/workspace");
    cudaStream_t st = (cudaStream_t)stream;
    switch (M) {
        case 1: return hm_pareto_run<1>(costs, N, mask, workspace, n_out, pass2, st, saw_nan_out);
        case 2: return hm_pareto_run<2>(costs, N, mask, workspace, n_out, pass2, st, saw_nan_out);
        case 3: return hm_pareto_run<3>(costs, N, mask, workspace, n_out, pass2, st, saw_nan_out);
        case 4: return hm_pareto_run<4>(costs, N, mask, workspace, n_out, pass2, st, saw_nan_out);
        case 5: return hm_pareto_run<5>(costs, N, mask, workspace, n_out, pass2, st, saw_nan_out);
        case 6: return hm_pareto_run<6>(costs, N, mask, workspace, n_out, pass2, st, saw_nan_out);
        case 7: return hm_pareto_run<7>(costs, N, mask, workspace, n_out, pass2, st, saw_nan_out);
        default: return hm_pareto_run<8>(costs, N, mask, workspace, n_out, pass2, st, saw_nan_out);
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

// rows_out[j][M] = costs[list[j]], idx_out[j] = index_base + list[j]   for j < n
__global__ void hm_pareto_gather_rows(const double* __restrict__ costs, int M, const unsigned* __restrict__ list, unsigned n,
                                      long long index_base, double* __restrict__ rows_out, long long* __restrict__ idx_out) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned i = list[j];
    for (int d = 0; d < M; ++d) rows_out[(size_t)j * M + d] = costs[(size_t)i * M + d];
    idx_out[j] = index_base + (long long)i;
}

// Pass 1 followed by the ordered compaction of its survivors, for the sharded path (distributed.pareto_mask_sharded):
// what a rank sends to the others is on the device, in index order, when this returns -- no host-side nonzero / gather
// and no separate NaN scan of the block (the filter itself detects NaNs).
extern "C" int hm_pareto_pass1_compact(const double* costs, int64_t N, int M, uint8_t* mask, void* workspace,
                                       double* rows_out, int64_t* idx_out, int64_t cap, int64_t index_base,
                                       int64_t* n_alive_out_host, int32_t* saw_nan_out_host, void* stream) {
    HM_REQUIRE(rows_out != nullptr && idx_out != nullptr && cap >= 0 && n_alive_out_host != nullptr, "outputs");
    int saw_nan = 0;
    int64_t n_alive = 0;
    int rc = hm_pareto_dispatch(costs, N, M, mask, workspace, &n_alive, false, stream, &saw_nan);
    if (rc) return rc;
    *n_alive_out_host = n_alive;
    if (saw_nan_out_host) *saw_nan_out_host = saw_nan;
    const int64_t take = std::min<int64_t>(n_alive, cap);
    if (N == 0 || take == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    HmParetoWs ws;
    unsigned char* base = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    hm_pareto_layout(N, M, base, &ws);
    const unsigned n = (unsigned)N;
    if (!saw_nan && n_alive <= HM_SORT_MAX) {
        // the fast path left the survivors' indices (unordered) in alive[0]: one-CTA index sort
        hm_px_sort_front<<<1, 1024, 0, st>>>(ws.alive[0], ws.ctl);
    } else {
        const unsigned nblk = (n + HM_PF_CHUNK - 1) / HM_PF_CHUNK;
        hm_px_count_flags<<<nblk, HM_PF_THREADS, 0, st>>>(mask, n, ws.block_counts, nullptr);
        hm_pareto_scan<<<1, 1024, 0, st>>>(ws.block_counts, nblk, ws.round, nullptr);
        hm_pareto_scatter<<<nblk, HM_PF_THREADS, 0, st>>>(nullptr, n, mask, ws.block_counts, ws.alive[0], 0xffffffffu,
                                                          ws.round, nullptr);
    }
    hm_pareto_gather_rows<<<(unsigned)((take + 255) / 256), 256, 0, st>>>(costs, M, ws.alive[0], (unsigned)take,
                                                                          (long long)index_base, rows_out,
                                                                          reinterpret_cast<long long*>(idx_out));
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

HM_CHECKED_EXPORT(pareto)
