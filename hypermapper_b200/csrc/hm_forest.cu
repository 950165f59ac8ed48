// K1: forest traversal + Hutter per-leaf table reduction, fused with the scalarised acquisition (K3).
//
// Replaces (reference luinardi/hypermapper):
//   RFModel.get_leaves_per_sample            models.py:174-182  (sklearn Tree._apply_dense)
//   RFModel.compute_rf_prediction            models.py:225-240
//   RFModel.compute_rf_prediction_variance   models.py:242-275
//   RandomForestRegressor.predict (TS path)  models.py:879-890
//   EI / ucb / thompson_sampling             random_scalarizations.py:160-506 (via hm_acq.cuh)
//
// Layout in HBM (built once per fitted model by hm_forest_create):
//   blob      : tree groups, each 16 B aligned:  [uint32 root[n_trees_g] (padded to 16 B)] [uint2 node[...]]
//               node.x = float32 threshold bits (internal)  |  forest-wide leaf ordinal (leaf)
//               node.y = feature << 22 | left   (left = group-relative index of the left child, right = left+1;
//                        left == 0  <=>  leaf)
//               trees are re-laid out breadth-first so the two children are adjacent and the hot top levels
//               share a few shared-memory lines.
//   groups    : descriptor per group (offset/bytes/first tree/#trees)
//   leaf_tab  : 4 fp64 per leaf  [q = m/T, s = m**2/T, u = v/T, value]  (32 B -> one sector, one 256-bit load per leaf visit)
//   leaf_node : sklearn node id per leaf (for the apply output)
//
// Kernel: one thread = one candidate, trees visited in tree order so the fp64 accumulation order is the
// reference's (sequential over t) and the mean is bit-identical.  The candidate's encoded features live in a
// [feature][thread] shared-memory tile (bank = lane: conflict free).  Tree groups are streamed through two
// shared-memory buffers with 1-D TMA bulk copies (cp.async.bulk + mbarrier); a forest that fits in the two
// buffers is loaded once per CTA and stays resident.  Persistent grid: one 1024-thread CTA per SM.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "hm_acq.cuh"

#define HM_LEFT_BITS 22
#define HM_LEFT_MASK ((1u << HM_LEFT_BITS) - 1u)
#define HM_MAX_FEATURES 1023
#define HM_SMEM_TOTAL (227 * 1024)
#define HM_SMEM_HEADER 128
#define HM_XTILE_MAX (100 * 1024)
#define HM_NODEBUF_MAX (96 * 1024)

struct HmGroupDesc {
    unsigned long long blob_off;
    unsigned int bytes;       // multiple of 16
    unsigned int first_tree;
    unsigned int n_trees;
    unsigned int nodes_off;   // byte offset of the node array inside the group
};

struct HmForestDev {
    const unsigned char* blob;
    const HmGroupDesc* groups;
    const double* leaf_tab;
    const int* leaf_node;
    int n_groups;
    int n_trees;
    int staged;
    int pad;
};

struct hm_forest {
    HmForestDev dev;
    int n_trees, n_features, n_leaves, staged, block, nodebuf;
    size_t blob_bytes;
    void* d_blob;
    void* d_groups;
    void* d_leaf_tab;
    void* d_leaf_node;
};

struct HmEvalArgs {
    HmForestDev f[HM_MAX_OBJECTIVES];
    int* leaf_out[HM_MAX_OBJECTIVES];
    int M, D, nodebuf, steps_per_tile;
    const float* X;
    long long ldx, N;
    double* mean_out;
    double* var_out;
    double* pred_out;
    double* score_out;
    const double* feas;
    HmAcqDev acq;
    int has_acq;
};

// block size / node-buffer size as a function of the encoded feature count (shared by create and launch)
static void hm_forest_geometry(int D, int* block, int* nodebuf) {
    int b = 1024;
    while (b > 128 && (size_t)b * D * 4 > HM_XTILE_MAX) b >>= 1;
    size_t xt = (size_t)b * D * 4;
    long long nb = ((long long)HM_SMEM_TOTAL - HM_SMEM_HEADER - (long long)xt) / 2;
    nb &= ~15ll;
    if (nb > HM_NODEBUF_MAX) nb = HM_NODEBUF_MAX;
    *block = b;
    *nodebuf = (int)nb;
}

// largest float32 <= t  (so that  x32 <= thr32  <=>  (double)x32 <= t  for every finite float32 x32)
static float hm_floor_to_f32(double t) {
    if (t != t) return (float)t;
    float f = (float)t;
    if ((double)f > t) f = nextafterf(f, -INFINITY);
    return f;
}

extern "C" int hm_forest_create(int n_trees, int n_features, const int64_t* off, const int32_t* feature,
                                const double* threshold, const int32_t* left, const int32_t* right,
                                const double* q, const double* s, const double* u, const double* value,
                                hm_forest_t** out) {
    HM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HM_REQUIRE(n_trees > 0, "n_trees > 0");
    HM_REQUIRE(n_features > 0 && n_features <= HM_MAX_FEATURES, "1 <= n_features <= 1023");
    HM_REQUIRE(off && feature && threshold && left && right, "tree arrays");
    int block, nodebuf;
    hm_forest_geometry(n_features, &block, &nodebuf);
    HM_REQUIRE((size_t)block * n_features * 4 <= HM_XTILE_MAX, "n_features too large for the shared-memory tile");

    // pass 1: per-tree BFS relayout (children adjacent), sizes
    struct TreeLay { std::vector<int> order; std::vector<int> newidx; };
    std::vector<TreeLay> lay(n_trees);
    size_t max_tree_bytes = 0;
    long long n_leaves = 0;
    for (int t = 0; t < n_trees; ++t) {
        long long b = off[t], e = off[t + 1];
        HM_REQUIRE(e > b, "empty tree");
        int n = (int)(e - b);
        HM_REQUIRE(n <= (int)HM_LEFT_MASK, "tree too large (>4M nodes)");
        TreeLay& L = lay[t];
        L.order.reserve(n);
        L.newidx.assign(n, -1);
        L.order.push_back(0);
        L.newidx[0] = 0;
        for (size_t h = 0; h < L.order.size(); ++h) {
            int nd = L.order[h];
            int l = left[b + nd], r = right[b + nd];
            if (l >= 0 && l != r) {
                HM_REQUIRE(l < n && r >= 0 && r < n, "child index out of range");
                HM_REQUIRE(L.newidx[l] < 0 && L.newidx[r] < 0, "tree is not a tree");
                int f = feature[b + nd];
                HM_REQUIRE(f >= 0 && f < n_features, "feature index out of range");
                L.newidx[l] = (int)L.order.size();
                L.order.push_back(l);
                L.newidx[r] = (int)L.order.size();
                L.order.push_back(r);
            } else {
                ++n_leaves;
            }
        }
        HM_REQUIRE((int)L.order.size() == n, "unreachable nodes in tree");
        max_tree_bytes = std::max(max_tree_bytes, (size_t)n * 8 + 16);
    }
    const int staged = (max_tree_bytes <= (size_t)nodebuf) ? 1 : 0;

    // pass 2: grouping
    std::vector<HmGroupDesc> groups;
    std::vector<unsigned char> blob;
    std::vector<double> leaf_tab((size_t)n_leaves * 4, 0.0);
    std::vector<int> leaf_node((size_t)n_leaves, 0);
    long long leaf_ord = 0;
    int t = 0;
    while (t < n_trees) {
        int t1 = t;
        size_t nodes = 0;
        // extend while the group fits (staged) / one tree per group (walked in L2)
        while (t1 < n_trees) {
            size_t nn = nodes + (size_t)(off[t1 + 1] - off[t1]);
            size_t hdr = (((size_t)(t1 - t + 1) * 4) + 15) & ~(size_t)15;
            size_t bytes = hdr + nn * 8;
            bytes = (bytes + 15) & ~(size_t)15;
            if (t1 > t && (!staged || bytes > (size_t)nodebuf)) break;
            nodes = nn;
            ++t1;
            if (!staged) break;
        }
        const int ng = t1 - t;
        const size_t hdr = (((size_t)ng * 4) + 15) & ~(size_t)15;
        size_t bytes = (hdr + nodes * 8 + 15) & ~(size_t)15;
        HM_REQUIRE(nodes <= HM_LEFT_MASK, "group too large");
        HmGroupDesc g;
        g.blob_off = blob.size();
        g.bytes = (unsigned)bytes;
        g.first_tree = (unsigned)t;
        g.n_trees = (unsigned)ng;
        g.nodes_off = (unsigned)hdr;
        blob.resize(blob.size() + bytes, 0);
        unsigned char* gp = blob.data() + g.blob_off;
        uint32_t* roots = (uint32_t*)gp;
        uint32_t* nd32 = (uint32_t*)(gp + hdr);
        uint32_t base = 0;
        for (int tt = t; tt < t1; ++tt) {
            const TreeLay& L = lay[tt];
            const long long b = off[tt];
            roots[tt - t] = base;
            for (size_t h = 0; h < L.order.size(); ++h) {
                int nd = L.order[h];
                int l = left[b + nd], r = right[b + nd];
                uint32_t* slot = nd32 + 2 * (size_t)(base + h);
                if (l >= 0 && l != r) {
                    float thr = hm_floor_to_f32(threshold[b + nd]);
                    uint32_t tb;
                    memcpy(&tb, &thr, 4);
                    slot[0] = tb;
                    slot[1] = ((uint32_t)feature[b + nd] << HM_LEFT_BITS) | (base + (uint32_t)L.newidx[l]);
                } else {
                    slot[0] = (uint32_t)leaf_ord;
                    slot[1] = 0u;
                    leaf_tab[4 * leaf_ord + 0] = q ? q[b + nd] : 0.0;
                    leaf_tab[4 * leaf_ord + 1] = s ? s[b + nd] : 0.0;
                    leaf_tab[4 * leaf_ord + 2] = u ? u[b + nd] : 0.0;
                    leaf_tab[4 * leaf_ord + 3] = value ? value[b + nd] : 0.0;
                    leaf_node[leaf_ord] = nd;
                    ++leaf_ord;
                }
            }
            base += (uint32_t)L.order.size();
        }
        groups.push_back(g);
        t = t1;
    }

    hm_forest* F = new (std::nothrow) hm_forest();
    if (!F) return HM_E_NOMEM;
    memset(F, 0, sizeof(*F));
    F->n_trees = n_trees;
    F->n_features = n_features;
    F->n_leaves = (int)n_leaves;
    F->staged = staged;
    F->block = block;
    F->nodebuf = nodebuf;
    F->blob_bytes = blob.size();
    cudaError_t e;
#define HM_F_TRY(x)                                                     \
    if ((e = (x)) != cudaSuccess) {                                     \
        hm_set_error("%s: %s", #x, cudaGetErrorString(e));              \
        hm_forest_destroy(F);                                           \
        return (int)e;                                                  \
    }
    HM_F_TRY(cudaMalloc(&F->d_blob, blob.size()));
    HM_F_TRY(cudaMalloc(&F->d_groups, groups.size() * sizeof(HmGroupDesc)));
    HM_F_TRY(cudaMalloc(&F->d_leaf_tab, leaf_tab.size() * sizeof(double)));
    HM_F_TRY(cudaMalloc(&F->d_leaf_node, leaf_node.size() * sizeof(int)));
    HM_F_TRY(cudaMemcpy(F->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    HM_F_TRY(cudaMemcpy(F->d_groups, groups.data(), groups.size() * sizeof(HmGroupDesc), cudaMemcpyHostToDevice));
    HM_F_TRY(cudaMemcpy(F->d_leaf_tab, leaf_tab.data(), leaf_tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    HM_F_TRY(cudaMemcpy(F->d_leaf_node, leaf_node.data(), leaf_node.size() * sizeof(int), cudaMemcpyHostToDevice));
#undef HM_F_TRY
    F->dev.blob = (const unsigned char*)F->d_blob;
    F->dev.groups = (const HmGroupDesc*)F->d_groups;
    F->dev.leaf_tab = (const double*)F->d_leaf_tab;
    F->dev.leaf_node = (const int*)F->d_leaf_node;
    F->dev.n_groups = (int)groups.size();
    F->dev.n_trees = n_trees;
    F->dev.staged = staged;
    *out = F;
    return 0;
}

extern "C" void hm_forest_destroy(hm_forest_t* F) {
    if (!F) return;
    cudaFree(F->d_blob);
    cudaFree(F->d_groups);
    cudaFree(F->d_leaf_tab);
    cudaFree(F->d_leaf_node);
    delete F;
}
extern "C" int hm_forest_n_trees(const hm_forest_t* f) { return f ? f->n_trees : 0; }
extern "C" int hm_forest_n_features(const hm_forest_t* f) { return f ? f->n_features : 0; }
extern "C" int hm_forest_staged(const hm_forest_t* f) { return f ? f->staged : 0; }

// ------------------------------------------------------------------------------------------------------------
struct HmAcc {
    double q, s, u, v;
};

// Walk `ntrees` trees whose roots/nodes are at the given pointers (shared or global), in tree order.
template <bool APPLY>
__device__ __forceinline__ void hm_walk_group(const uint32_t* __restrict__ roots, const uint2* __restrict__ nodes,
                                              int ntrees, const float* __restrict__ xs_t, int BD,
                                              const double* __restrict__ leaf_tab, const int* __restrict__ leaf_node,
                                              HmAcc& acc, int* __restrict__ leaf_out, long long N, long long i,
                                              bool valid, int first_tree) {
    for (int t = 0; t < ntrees; ++t) {
        uint2 nd = nodes[roots[t]];
        while (nd.y & HM_LEFT_MASK) {
            const float xv = xs_t[(nd.y >> HM_LEFT_BITS) * BD];
            // sklearn: go left iff x <= threshold (NaN -> right)
            const uint32_t nxt = (nd.y & HM_LEFT_MASK) + ((xv <= __uint_as_float(nd.x)) ? 0u : 1u);
            nd = nodes[nxt];
        }
        // one 256-bit load of the 32-byte leaf record (LDG.E.256): one L1 wavefront pass instead of two
        double lq, ls, lu, lv;
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(lq), "=d"(ls), "=d"(lu), "=d"(lv)
                     : "l"(leaf_tab + 4 * (size_t)nd.x));
        acc.q = hm_add(acc.q, lq);
        acc.s = hm_add(acc.s, ls);
        acc.u = hm_add(acc.u, lu);
        acc.v = hm_add(acc.v, lv);
        if (APPLY) {
            if (valid && leaf_out) leaf_out[(size_t)(first_tree + t) * (size_t)N + (size_t)i] = __ldg(leaf_node + nd.x);
        }
    }
}

template <bool APPLY>
__global__ void __launch_bounds__(1024, 1) hm_forest_kernel(const __grid_constant__ HmEvalArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* nbuf0 = smem + HM_SMEM_HEADER;
    float* xs = reinterpret_cast<float*>(smem + HM_SMEM_HEADER + 2 * (size_t)a.nodebuf);

    const int tid = threadIdx.x;
    const int BD = blockDim.x;
    const long long n_tiles = (a.N + BD - 1) / BD;
    if ((long long)blockIdx.x >= n_tiles) return;
    const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

    const int spt = a.steps_per_tile;
    const bool resident = spt <= 2;
    const long long total_steps = resident ? (long long)spt : my_tiles * (long long)spt;

    // prefetch cursor (thread 0 only)
    int pm = 0, pg = 0;
    long long issued = 0;
    if (tid == 0) {
        hm_mbar_init(&bars[0], 1);
        hm_mbar_init(&bars[1], 1);
        hm_fence_barrier_init();
        while (pm < a.M && !a.f[pm].staged) ++pm;
    }
    __syncthreads();

    auto issue = [&](int buf) {
        const HmGroupDesc* gd = a.f[pm].groups + pg;
        const unsigned bytes = gd->bytes;
        hm_mbar_expect_tx(&bars[buf], bytes);
        hm_bulk_g2s(nbuf0 + (size_t)buf * a.nodebuf, a.f[pm].blob + gd->blob_off, bytes, &bars[buf]);
        ++issued;
        if (++pg >= a.f[pm].n_groups) {
            pg = 0;
            do {
                pm = (pm + 1 == a.M) ? 0 : pm + 1;
            } while (!a.f[pm].staged);
        }
    };
    if (tid == 0) {
        if (total_steps > 0) issue(0);
        if (total_steps > 1) issue(1);
    }

    long long step = 0;
    float* xs_t = xs + tid;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i = tile * BD + tid;
        const bool valid = i < a.N;
        const long long ii = valid ? i : a.N - 1;
        for (int f = 0; f < a.D; ++f) xs_t[f * BD] = __ldg(a.X + (size_t)f * (size_t)a.ldx + (size_t)ii);

        double mean_l[HM_MAX_OBJECTIVES], var_l[HM_MAX_OBJECTIVES];
        int step_in_tile = 0;
        for (int m = 0; m < a.M; ++m) {
            const HmForestDev& F = a.f[m];
            HmAcc acc = {0.0, 0.0, 0.0, 0.0};
            int* lo = APPLY ? a.leaf_out[m] : nullptr;
            if (F.staged) {
                for (int g = 0; g < F.n_groups; ++g) {
                    const int buf = resident ? step_in_tile : (int)(step & 1);
                    const uint32_t parity = resident ? 0u : (uint32_t)((step >> 1) & 1);
                    hm_mbar_wait(&bars[buf], parity);
                    const unsigned char* gb = nbuf0 + (size_t)buf * a.nodebuf;
                    const unsigned ntr = __ldg(&F.groups[g].n_trees);
                    const unsigned noff = __ldg(&F.groups[g].nodes_off);
                    const unsigned ft = __ldg(&F.groups[g].first_tree);
                    hm_walk_group<APPLY>(reinterpret_cast<const uint32_t*>(gb),
                                         reinterpret_cast<const uint2*>(gb + noff), (int)ntr, xs_t, BD, F.leaf_tab,
                                         F.leaf_node, acc, lo, a.N, i, valid, (int)ft);
                    if (!resident) {
                        __syncthreads();   // everyone is done reading this buffer
                        if (tid == 0 && issued < total_steps) issue(buf);
                    }
                    ++step;
                    ++step_in_tile;
                }
            } else {
                for (int g = 0; g < F.n_groups; ++g) {
                    const HmGroupDesc gd = F.groups[g];
                    const unsigned char* gb = F.blob + gd.blob_off;
                    hm_walk_group<APPLY>(reinterpret_cast<const uint32_t*>(gb),
                                         reinterpret_cast<const uint2*>(gb + gd.nodes_off), (int)gd.n_trees, xs_t, BD,
                                         F.leaf_tab, F.leaf_node, acc, lo, a.N, i, valid, (int)gd.first_tree);
                }
            }
            // models.py:266-273
            const double mean = acc.q;
            double vom = fabs(hm_sub(acc.s, hm_mul(mean, mean)));
            double var = hm_add(acc.u, vom);
            if (var == 0.0) var = 0.00001;
            if (valid) {
                if (a.mean_out) a.mean_out[(size_t)m * (size_t)a.N + (size_t)i] = mean;
                if (a.var_out) a.var_out[(size_t)m * (size_t)a.N + (size_t)i] = var;
                if (a.pred_out) a.pred_out[(size_t)m * (size_t)a.N + (size_t)i] = acc.v / (double)F.n_trees;
            }
            if (a.has_acq) {
                if (a.acq.kind == HM_ACQ_TS) {
                    mean_l[m] = acc.v / (double)F.n_trees;
                    var_l[m] = 0.0;
                } else {
                    // models.py:762-763: NaN means become sys.maxsize
                    mean_l[m] = (mean != mean) ? 9.223372036854775807e18 : mean;
                    var_l[m] = var;
                }
            }
        }
        if (a.has_acq && valid) {
            const double feas = a.feas ? __ldg(a.feas + i) : 1.0;
            a.score_out[i] = hm_acq_value(a.acq, mean_l, var_l, feas);
        }
    }
}

extern "C" int hm_rf_eval(const hm_forest_t* const* forests, int M, const float* Xcm, int64_t ldx, int64_t N,
                          int32_t* const* leaf_out, double* mean_out, double* var_out, double* pred_out,
                          const hm_acq_params_t* acq, const double* feas, double* score_out, void* stream) {
    HM_REQUIRE(forests != nullptr && M >= 1 && M <= HM_MAX_OBJECTIVES, "1 <= M <= 8");
    HM_REQUIRE(N >= 0 && ldx >= N, "ldx >= N >= 0");
    if (N == 0) return 0;
    HM_REQUIRE(Xcm != nullptr, "Xcm");
    HmEvalArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M;
    a.D = forests[0] ? forests[0]->n_features : 0;
    int steps = 0;
    bool any_leaf = false;
    for (int m = 0; m < M; ++m) {
        HM_REQUIRE(forests[m] != nullptr, "forest handle");
        HM_REQUIRE(forests[m]->n_features == a.D, "all forests must share n_features");
        a.f[m] = forests[m]->dev;
        if (forests[m]->staged) steps += forests[m]->dev.n_groups;
        a.leaf_out[m] = leaf_out ? leaf_out[m] : nullptr;
        any_leaf = any_leaf || (a.leaf_out[m] != nullptr);
    }
    a.nodebuf = forests[0]->nodebuf;
    a.steps_per_tile = steps;
    a.X = Xcm;
    a.ldx = ldx;
    a.N = N;
    a.mean_out = mean_out;
    a.var_out = var_out;
    a.pred_out = pred_out;
    a.score_out = score_out;
    a.feas = feas;
    a.has_acq = 0;
    if (acq) {
        HM_REQUIRE(score_out != nullptr, "score_out required with acq");
        HM_REQUIRE(acq->n_objectives == M, "acq->n_objectives == M");
        int rc = hm_acq_to_dev(acq, &a.acq);
        if (rc) {
            hm_set_error("invalid acquisition parameters");
            return rc;
        }
        a.has_acq = 1;
    } else {
        HM_REQUIRE(score_out == nullptr, "score_out without acq");
    }
    const int block = forests[0]->block;
    const size_t smem = HM_SMEM_HEADER + 2 * (size_t)a.nodebuf + (size_t)block * a.D * 4;
    const long long n_tiles = (N + block - 1) / block;
    const int grid = (int)std::min<long long>(n_tiles, hm_num_sms());
    cudaStream_t st = (cudaStream_t)stream;
    if (any_leaf) {
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_forest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hm_forest_kernel<true><<<grid, block, smem, st>>>(a);
    } else {
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_forest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hm_forest_kernel<false><<<grid, block, smem, st>>>(a);
    }
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}
