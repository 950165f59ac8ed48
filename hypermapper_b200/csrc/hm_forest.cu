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
//   blob      : tree groups, each 16 B aligned:  [uint2 root_node_copy[n_trees_g] (padded to 16 B)] [uint2 node[...]]
//               [+ 32-byte leaf records of the group's trees, 32 B aligned, when the forest stays resident in shared memory]
//               node.x = float32 threshold bits (internal)  |  forest-wide leaf ordinal (leaf)
//               node.y = feature << 22 | left   (left = group-relative index of the left child, right = left+1;
//                        left == 0  <=>  leaf)
//               trees are re-laid out breadth-first (hm_forest_node_order) so the two children are adjacent and the hot
//               top levels share a few shared-memory lines.
//   groups    : descriptor per group (offset/bytes/first tree/#trees)
//   leaf_tab  : 4 fp64 per leaf  [q = m/T, s = m**2/T, u = v/T, value]  (32 B -> one sector, one 256-bit load per leaf visit)
//   leaf_node : sklearn node id per leaf (for the apply output)
//
// Kernel: one thread = one candidate, trees visited in tree order so the fp64 accumulation order is the
// reference's (sequential over t) and the mean is bit-identical.  The candidate's encoded features live in a
// [feature][thread] shared-memory tile (bank = lane: conflict free).  Tree groups are streamed through two
// shared-memory buffers with 1-D TMA bulk copies (cp.async.bulk + mbarrier); a forest that fits in the two
// buffers is loaded once per CTA and stays resident.  Persistent grid: one 1024-thread CTA per SM.
//
// Coherence pre-pass (large forests, large pools).  A forest that streams through shared memory is bound by the
// L1/shared-memory data pipe: with candidates in arrival order the 32 lanes of a warp sit in 12 different nodes per
// visit (2.5-way bank conflicts) and 30 % of the lanes idle behind the deepest path.  The forest itself is a spatial
// index: candidates that land in the same leaves of two "key" trees follow nearly the same path in every other tree.
// So the pool is bucketed by (leaf of key tree A, leaf of key tree B) -- key kernel + histogram, scan, scatter: a
// counting sort on <= 2^22 bins that also leaves the candidates as contiguous rows in bucket order -- and the
// traversal kernel walks that copy (plain coalesced tile loads; results scattered back through the permutation; every candidate's arithmetic is untouched, so results stay bit-identical).  Measured on
// the C3 shape this cuts distinct nodes per warp visit 12.5 -> 2.3 and raises lane utilisation 0.70 -> 0.86.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <utility>
#include <vector>
#include "hm_acq.cuh"
#include "hm_pack.cuh"

#define HM_LEFT_BITS 22
#define HM_LEFT_MASK ((1u << HM_LEFT_BITS) - 1u)
// Forests that stream through shared memory use a second layout of a node's word 1 (word 0 = threshold bits / leaf
// ordinal in both):   (feature * block) << 16 | left      -- "packed"
// i.e. the upper half is the float index of the feature's column inside the [feature][thread] tile, so that the tile
// address is  tile + (word1 >> 14)  (left < 2^14 keeps bits 14-15 clear) and the child address  nodes + left * 8:
// one instruction each, where  feature << 22 | left  costs a shift and a multiply-add for the tile address.
// A group holds < 2^14 nodes (HM_NODEBUF_MAX / 8) and feature * block <= HM_XTILE_MAX / 4 < 2^16.
#define HM_PK_LEFT_MASK 0xffffu
#define HM_MAX_FEATURES 1023
#define HM_SMEM_TOTAL (227 * 1024)
#define HM_SMEM_HEADER 128
// tile budget: 1024 candidates per CTA up to 33 encoded features (32 warps hide the shared-memory latency better than
// larger tree groups do: C5-RF, D_enc = 32, 43.9 -> 40.3 ms when the block went from 512 to 1024 threads)
#define HM_XTILE_MAX (132 * 1024)
#define HM_NODEBUF_MAX (96 * 1024)
// Packed candidates (hm_pack.cuh): a third node layout,
//   word 0 = T (compare value)  |  forest-wide leaf ordinal (leaf)
//   word 1 = (byte offset of the first child slot inside the group's node array) << 6 | s        (0 <=> leaf)
// second slot <=> hi32(code << s) >= T.  There is no feature tile, so the two node buffers take all of shared memory.
// Layout kinds (hm_forest.packed): 1 = FAST (every shift <= 31), leaf word 1 = 0; 2 = wide (64-bit shift), leaf word 1 = 0;
// 3 = FAST with self-looping leaves (hm_pack.cuh: T = bound + leaf ordinal, word 1 = own offset << 6 | 32 | sentinel shift).
#ifndef HM_PACKED_BLOCK
#define HM_PACKED_BLOCK 1024
#endif
// levels walked between two "is any chain still at an internal node" tests of the shared-memory walks (a finished chain's
// visit is predicated off, so extra levels are harmless): the test + branch cost 2 of the 20 instructions of a 3-chain level
#ifndef HM_WALK_LEVELS
#define HM_WALK_LEVELS 2
#endif
#ifndef HM_NODE_ORDER_DEFAULT
#define HM_NODE_ORDER_DEFAULT 0      // 0 = breadth-first node arrays, 1 = depth-first (hm_forest_node_order)
#endif
#ifndef HM_LOOP_LEVELS
#define HM_LOOP_LEVELS 1             // the same for the self-looping-leaf walk (packed kind 3): 1 (C3 51.7 -> 51.1 ms: an
                                     // arrived chain re-loads its leaf, so a level too many costs L1-pipe wavefronts)
#endif
#define HM_PACKED_CTAS (1024 / HM_PACKED_BLOCK)          // CTAs per SM of the packed kernel
#define HM_PACKED_NODEBUF (((HM_SMEM_TOTAL - 1024 * HM_PACKED_CTAS) / HM_PACKED_CTAS - HM_SMEM_HEADER) / 2 / 16 * 16)

struct HmGroupDesc {
    unsigned long long blob_off;
    unsigned int bytes;       // multiple of 16
    unsigned int first_tree;
    unsigned int n_trees;
    unsigned int nodes_off;   // byte offset of the node array inside the group
    unsigned int leaf_off;    // byte offset of the group's own copy of its trees' leaf records (0: the group carries none)
    unsigned int leaf_base;   // forest-wide ordinal of the first of those records
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

// a tree the coherence pre-pass walks in global memory
struct HmKeyTree {
    const uint2* nodes;          // node array of the tree's group
    const unsigned* leaf_rank;   // forest-wide: DFS rank of a leaf among the leaves of its own tree
    unsigned root;               // group-relative index of the root
    unsigned n_leaves;
    unsigned feat_shift;         // feature = word1 >> feat_shift      (22, or 16 + log2(block) for the packed layout)
    unsigned left_mask;          // left    = word1 & left_mask
    unsigned leaf_and, leaf_eq;  // packed trees: leaf <=> (word1 & leaf_and) == leaf_eq;  leaf_rank is biased by the leaf bound
};

// host copy of the trees in breadth-first order (children adjacent), kept so that further node layouts (the packed one)
// can be built after hm_forest_create
struct HmHostNode {
    uint32_t w0;       // float32 threshold bits (internal)  |  forest-wide leaf ordinal (leaf)
    int32_t feat;      // split feature, -1 at leaves (packed trees: the shift s)
    uint32_t left;     // tree-local index of the left child (right = left + 1)
};
typedef std::vector<std::vector<HmHostNode> > HmHostTrees;

struct hm_forest {
    HmForestDev dev;
    int n_trees, n_features, n_leaves, staged, block, nodebuf;
    size_t blob_bytes;
    void* d_blob;
    void* d_groups;
    void* d_leaf_tab;
    void* d_leaf_node;
    void* d_leaf_rank;
    int n_key;
    HmKeyTree key[2];
    HmHostTrees* host_trees;
    std::vector<long long>* tree_leaf_base;
    std::vector<double>* host_leaf_tab;   // kept for small forests only (the packed groups may carry their leaf records too)
    int leaf_smem, leaf_smem_p;           // every group of the fp32 / packed layout carries its leaf records (resident forests)
    int sort_mode;                 // coherence pre-pass: -1 auto, 0 off, 1 on
    // packed layout (hm_forest_attach_packing)
    int packed;                    // once the packed node arrays exist: 1 = FAST layout (every shift <= 31), 2 = wide,
                                   // 3 = FAST with self-looping leaves
    unsigned loop_shift, loop_bound;   // kind 3: the sentinel field's shift and bound (hm_pack.cuh)
    int packed_bits;
    int nodebuf_p;                 // node-buffer size the packed groups were built for
    HmForestDev devp;
    void* d_blob_p;
    void* d_groups_p;
    HmKeyTree keyp[2];
};

// multipliers handed to the kernel as RUN-TIME values: as literals the compiler would strength-reduce the multiply-adds
// that use them into alu-pipe shifts / adds, and the point of those instructions is to run on the fma pipe
struct HmRtConst {
    unsigned k26, k1, k8;
    int nodebuf;          // node-buffer bytes (checked build: staged node addresses are verified against it)
    int xtile_bytes;      // fp32 path: bytes of the [feature][thread] tile
    unsigned leaf_and, leaf_eq;        // packed trees walked in L2: leaf <=> (word1 & leaf_and) == leaf_eq
    unsigned loop_shift, loop_bound;   // packed kind 3: sentinel field of the self-looping leaves
};

struct HmEvalArgs {
    HmForestDev f[HM_MAX_OBJECTIVES];
    int* leaf_out[HM_MAX_OBJECTIVES];
    int M, D, nodebuf, steps_per_tile;
    const unsigned long long* codes;   // packed path: one 64-bit code per candidate (arrival order, or bucket order with perm)
    HmRtConst k;                       // 1 << 26, 1, 8 as run-time values
    const float* X;
    long long ldx, N;
    double* mean_out;
    double* var_out;
    double* pred_out;
    double* score_out;
    const double* feas;
    const unsigned* perm;     // coherence order (nullptr: arrival order)
    const float* Xrows;       // with perm: the candidates physically in coherence order, row-major [N][Dp]
    int Dp;                   // row stride of Xrows (D rounded up to a multiple of 4)
    HmAcqDev acq;
    int has_acq;
    // fused scoring + selection (hm_rf_score_topk): instead of writing every score, a candidate that can still be among
    // the k best -- score < *filter_tau, or == *filter_tau with an index below filter_n0, or NaN -- is appended to a list
    const double* filter_tau;
    double* list_val;
    long long* list_idx;
    unsigned* list_count;
    unsigned list_cap;
    long long filter_n0;
};

struct HmScoreFilter {
    const double* tau;
    double* list_val;
    long long* list_idx;
    unsigned* list_count;
    unsigned list_cap;
    long long n0;
};

// block size / node-buffer size as a function of the encoded feature count (shared by create and launch)
static void hm_forest_geometry(int D, int* block, int* nodebuf) {
    int b = 1024;
    const char* e = getenv("HM_B200_XTILE_MAX");          // experiments: overrides the tile budget (bytes)
    const long long xtile_max = e ? atoll(e) : (long long)HM_XTILE_MAX;
    while (b > 128 && (long long)b * D * 4 > xtile_max) b >>= 1;
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

// coherence pre-pass default of new handles: -1 auto, 0 off, 1 on (env HM_B200_FOREST_SORT); per handle afterwards
static int hm_forest_default_sort_mode() {
    const char* e = getenv("HM_B200_FOREST_SORT");
    const int m = e ? atoi(e) : -1;
    return (m < -1 || m > 1) ? -1 : m;
}

enum { HM_LAYOUT_F32_SMEM = 0, HM_LAYOUT_F32_L2 = 1, HM_LAYOUT_PACKED = 2, HM_LAYOUT_PACKED_LOOP = 3 };

// Groups the trees (as many consecutive trees as fit one node buffer when the forest is staged through shared memory, one
// tree per group when it is walked in L2) and lays the node words out for `layout`.
// Order of a tree's child PAIRS in its node array.  The kernels only need the two children of a node to be adjacent; where a
// pair sits is free.  Breadth-first (0) keeps the nodes of one depth together; depth-first (1) keeps every SUBTREE contiguous:
// its m descendants occupy m consecutive slots right behind its child pair, so the nodes -- and, as leaf ordinals follow the
// same order, the 32-byte leaf records -- that the coherent candidates of a warp touch late in a walk are neighbours in
// memory whatever their depth: distinct shared-memory banks for the node loads and shared 128-byte lines for the record gathers.
static int hm_forest_node_order() {
    static int order = -1;
    if (order < 0) {
        const char* e = getenv("HM_B200_NODE_ORDER");
        order = (e && e[0] == 'b') ? 0 : (e && e[0] == 'd') ? 1 : HM_NODE_ORDER_DEFAULT;
    }
    return order;
}

static int hm_build_blob(const HmHostTrees& trees, int layout, int nodebuf, int block, bool staged,
                         std::vector<unsigned char>& blob, std::vector<HmGroupDesc>& groups,
                         std::vector<unsigned>& tree_root, std::vector<long long>& tree_nodes_off,
                         unsigned loop_shift = 0, unsigned loop_bound = 0, const double* leaf_rows = nullptr,
                         const std::vector<long long>* leaf_base = nullptr) {
    // leaf_rows (4 doubles per forest-wide leaf ordinal) + leaf_base (first ordinal of every tree): every group also
    // carries the records of its trees' leaves, 32-byte aligned behind the node array -- for forests small enough to stay
    // resident in shared memory the kernel then never gathers a leaf record from global memory (mode NVM = 2)
    const bool with_leaves = leaf_rows != nullptr && leaf_base != nullptr && staged;
    auto group_bytes = [&](int t0, int t1x, size_t nn, size_t* leaf_off_out) {
        const size_t hdr = (((size_t)(t1x - t0) * 8) + 15) & ~(size_t)15;
        size_t bytes = (hdr + nn * 8 + 15) & ~(size_t)15;
        if (with_leaves) {
            bytes = (bytes + 31) & ~(size_t)31;
            if (leaf_off_out) *leaf_off_out = bytes;
            bytes += (size_t)((*leaf_base)[t1x] - (*leaf_base)[t0]) * 32;
        }
        return bytes;
    };
    const int n_trees = (int)trees.size();
    tree_root.assign(n_trees, 0);
    tree_nodes_off.assign(n_trees, 0);
    int t = 0;
    while (t < n_trees) {
        int t1 = t;
        size_t nodes = 0;
        while (t1 < n_trees) {
            size_t nn = nodes + trees[t1].size();
            const size_t bytes = group_bytes(t, t1 + 1, nn, nullptr);
            if (t1 > t && (!staged || bytes > (size_t)nodebuf)) break;
            nodes = nn;
            ++t1;
            if (!staged) break;
        }
        const int ng = t1 - t;
        const size_t hdr = (((size_t)ng * 8) + 15) & ~(size_t)15;
        size_t leaf_off = 0;
        const size_t bytes = group_bytes(t, t1, nodes, &leaf_off);
        HM_REQUIRE(nodes <= HM_LEFT_MASK, "group too large");
        HM_REQUIRE(layout != HM_LAYOUT_F32_SMEM || nodes < (1u << 14), "shared-memory group too large for the packed-word node layout");
        HmGroupDesc g;
        g.blob_off = blob.size();
        g.bytes = (unsigned)bytes;
        g.first_tree = (unsigned)t;
        g.n_trees = (unsigned)ng;
        g.nodes_off = (unsigned)hdr;
        g.leaf_off = with_leaves ? (unsigned)leaf_off : 0u;
        g.leaf_base = with_leaves ? (unsigned)(*leaf_base)[t] : 0u;
        blob.resize(blob.size() + bytes, 0);
        unsigned char* gp = blob.data() + g.blob_off;
        if (with_leaves)
            memcpy(gp + leaf_off, leaf_rows + 4 * (size_t)(*leaf_base)[t], (size_t)((*leaf_base)[t1] - (*leaf_base)[t]) * 32);
        // header: a copy of every tree's root node, so a round starts with ONE uniform load per chain (no root index ->
        // node dependency); the key-tree walkers of the pre-pass keep the group-relative root index (tree_root)
        uint32_t* roots = (uint32_t*)gp;
        uint32_t* nd32 = (uint32_t*)(gp + hdr);
        uint32_t base = 0;
        for (int tt = t; tt < t1; ++tt) {
            const std::vector<HmHostNode>& T = trees[tt];
            tree_root[tt] = base;
            tree_nodes_off[tt] = (long long)(g.blob_off + hdr);
            for (size_t h = 0; h < T.size(); ++h) {
                uint32_t* slot = nd32 + 2 * (size_t)(base + h);
                slot[0] = T[h].w0;
                if (T[h].feat < 0) {
                    slot[1] = 0u;
                    if (layout == HM_LAYOUT_PACKED_LOOP) {      // the always-false sentinel test: first slot = the leaf itself
                        slot[0] = loop_bound + T[h].w0;
                        slot[1] = (((base + (uint32_t)h) * 8u) << 6) | 32u | loop_shift;
                    }
                } else if (layout == HM_LAYOUT_F32_SMEM) {
                    slot[1] = (((uint32_t)T[h].feat * (uint32_t)block) << 16) | (base + T[h].left);
                } else if (layout == HM_LAYOUT_F32_L2) {
                    slot[1] = ((uint32_t)T[h].feat << HM_LEFT_BITS) | (base + T[h].left);
                } else {
                    slot[1] = (((base + T[h].left) * 8u) << 6) | (uint32_t)T[h].feat;
                }
            }
            roots[2 * (tt - t)] = nd32[2 * (size_t)base];
            roots[2 * (tt - t) + 1] = nd32[2 * (size_t)base + 1];
            base += (uint32_t)T.size();
        }
        groups.push_back(g);
        t = t1;
    }
    return 0;
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
    HM_REQUIRE((size_t)block * n_features * 4 <= 160 * 1024, "n_features too large for the shared-memory tile");

    // pass 1: per-tree relayout (children adjacent; breadth- or depth-first, hm_forest_node_order), leaf ordinals in
    // (tree, layout) order
    long long n_leaves_total = 0;
    for (int t = 0; t < n_trees; ++t) {
        HM_REQUIRE(off[t + 1] > off[t], "empty tree");
        HM_REQUIRE(off[t + 1] - off[t] <= (long long)HM_LEFT_MASK, "tree too large (>4M nodes)");
        for (long long i = off[t]; i < off[t + 1]; ++i) n_leaves_total += !(left[i] >= 0 && left[i] != right[i]);
    }
    HmHostTrees* trees = new (std::nothrow) HmHostTrees((size_t)n_trees);
    std::vector<long long>* leaf_base = new (std::nothrow) std::vector<long long>((size_t)n_trees + 1, 0);
    if (!trees || !leaf_base) {
        delete trees;
        delete leaf_base;
        return HM_E_NOMEM;
    }
    struct Guard {
        HmHostTrees* a;
        std::vector<long long>* b;
        ~Guard() { delete a; delete b; }
    } guard = {trees, leaf_base};
    std::vector<double> leaf_tab((size_t)n_leaves_total * 4, 0.0);
    std::vector<int> leaf_node((size_t)n_leaves_total, 0);
    std::vector<unsigned> leaf_rank((size_t)n_leaves_total, 0);
    size_t max_tree_bytes = 0;
    long long leaf_ord = 0;
    for (int t = 0; t < n_trees; ++t) {
        const long long b = off[t];
        const int n = (int)(off[t + 1] - b);
        std::vector<int> order, newidx(n, -1);
        order.reserve(n);
        order.push_back(0);
        newidx[0] = 0;
        // work list of placed nodes whose child pair is still to be placed: a queue (breadth-first) or a stack (depth-first)
        const bool dfs = hm_forest_node_order() == 1;
        std::vector<int> work(1, 0);
        size_t head = 0;
        while (dfs ? !work.empty() : head < work.size()) {
            int nd;
            if (dfs) {
                nd = work.back();
                work.pop_back();
            } else {
                nd = work[head++];
            }
            const int l = left[b + nd], r = right[b + nd];
            if (l >= 0 && l != r) {
                HM_REQUIRE(l < n && r >= 0 && r < n, "child index out of range");
                HM_REQUIRE(newidx[l] < 0 && newidx[r] < 0, "tree is not a tree");
                const int f = feature[b + nd];
                HM_REQUIRE(f >= 0 && f < n_features, "feature index out of range");
                newidx[l] = (int)order.size();
                order.push_back(l);
                newidx[r] = (int)order.size();
                order.push_back(r);
                if (dfs) {                 // the left subtree is laid out first
                    work.push_back(r);
                    work.push_back(l);
                } else {
                    work.push_back(l);
                    work.push_back(r);
                }
            }
        }
        HM_REQUIRE((int)order.size() == n, "unreachable nodes in tree");
        max_tree_bytes = std::max(max_tree_bytes, (size_t)n * 8 + 16);
        std::vector<HmHostNode>& T = (*trees)[t];
        T.resize(n);
        (*leaf_base)[t] = leaf_ord;
        for (int h = 0; h < n; ++h) {
            const int nd = order[h];
            const int l = left[b + nd], r = right[b + nd];
            if (l >= 0 && l != r) {
                const float thr = hm_floor_to_f32(threshold[b + nd]);
                memcpy(&T[h].w0, &thr, 4);
                T[h].feat = feature[b + nd];
                T[h].left = (uint32_t)newidx[l];
            } else {
                T[h].w0 = (uint32_t)leaf_ord;
                T[h].feat = -1;
                T[h].left = 0;
                leaf_tab[4 * leaf_ord + 0] = q ? q[b + nd] : 0.0;
                leaf_tab[4 * leaf_ord + 1] = s ? s[b + nd] : 0.0;
                leaf_tab[4 * leaf_ord + 2] = u ? u[b + nd] : 0.0;
                leaf_tab[4 * leaf_ord + 3] = value ? value[b + nd] : 0.0;
                leaf_node[leaf_ord] = nd;
                ++leaf_ord;
            }
        }
        // DFS (= sklearn node id) rank of every leaf of this tree: neighbouring ranks are neighbouring regions
        {
            const long long lb = (*leaf_base)[t], le = leaf_ord;
            std::vector<std::pair<int, long long> > byid;
            byid.reserve((size_t)(le - lb));
            for (long long o = lb; o < le; ++o) byid.push_back(std::make_pair(leaf_node[o], o));
            std::sort(byid.begin(), byid.end());
            for (size_t r = 0; r < byid.size(); ++r) leaf_rank[byid[r].second] = (unsigned)r;
        }
    }
    (*leaf_base)[n_trees] = leaf_ord;
    const long long n_leaves = leaf_ord;
    const int staged = (max_tree_bytes <= (size_t)nodebuf) ? 1 : 0;

    // pass 2: grouping + node words of the fp32 layout
    std::vector<HmGroupDesc> groups;
    std::vector<unsigned char> blob;
    std::vector<unsigned> tree_root;
    std::vector<long long> tree_nodes_off;
    {
        int rc = hm_build_blob(*trees, staged ? HM_LAYOUT_F32_SMEM : HM_LAYOUT_F32_L2, nodebuf, block, staged != 0, blob,
                               groups, tree_root, tree_nodes_off);
        if (rc) return rc;
    }
    // a forest that stays resident in shared memory (<= 2 groups) takes its leaf records along if they fit too
    int leaf_smem = 0;
    if (staged && groups.size() <= 2 && !getenv("HM_B200_NO_LEAF_SMEM")) {
        std::vector<HmGroupDesc> g2;
        std::vector<unsigned char> b2;
        std::vector<unsigned> r2;
        std::vector<long long> o2;
        int rc = hm_build_blob(*trees, HM_LAYOUT_F32_SMEM, nodebuf, block, true, b2, g2, r2, o2, 0, 0, leaf_tab.data(), leaf_base);
        if (rc) return rc;
        if (g2.size() <= 2) {
            groups.swap(g2);
            blob.swap(b2);
            tree_root.swap(r2);
            tree_nodes_off.swap(o2);
            leaf_smem = 1;
        }
    }

    hm_forest* F = new (std::nothrow) hm_forest();
    if (!F) return HM_E_NOMEM;
    memset(F, 0, sizeof(*F));
    F->host_trees = trees;
    F->tree_leaf_base = leaf_base;
    F->leaf_smem = leaf_smem;
    if (leaf_tab.size() * sizeof(double) <= (size_t)HM_SMEM_TOTAL) F->host_leaf_tab = new (std::nothrow) std::vector<double>(leaf_tab);
    guard.a = nullptr;
    guard.b = nullptr;
    F->sort_mode = hm_forest_default_sort_mode();
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
    HM_F_TRY(cudaMalloc(&F->d_leaf_tab, std::max<size_t>(leaf_tab.size(), 1) * sizeof(double)));
    HM_F_TRY(cudaMalloc(&F->d_leaf_node, std::max<size_t>(leaf_node.size(), 1) * sizeof(int)));
    HM_F_TRY(cudaMalloc(&F->d_leaf_rank, std::max<size_t>(leaf_rank.size(), 1) * sizeof(unsigned)));
    HM_F_TRY(cudaMemcpy(F->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    HM_F_TRY(cudaMemcpy(F->d_groups, groups.data(), groups.size() * sizeof(HmGroupDesc), cudaMemcpyHostToDevice));
    HM_F_TRY(cudaMemcpy(F->d_leaf_tab, leaf_tab.data(), leaf_tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    HM_F_TRY(cudaMemcpy(F->d_leaf_node, leaf_node.data(), leaf_node.size() * sizeof(int), cudaMemcpyHostToDevice));
    HM_F_TRY(cudaMemcpy(F->d_leaf_rank, leaf_rank.data(), leaf_rank.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    // the handle is consumed on caller streams (often cudaStreamNonBlocking) that do not order against the legacy
    // stream these pageable-memory uploads went through
    HM_F_TRY(cudaDeviceSynchronize());
#undef HM_F_TRY
    F->dev.blob = (const unsigned char*)F->d_blob;
    F->dev.groups = (const HmGroupDesc*)F->d_groups;
    F->dev.leaf_tab = (const double*)F->d_leaf_tab;
    F->dev.leaf_node = (const int*)F->d_leaf_node;
    F->dev.n_groups = (int)groups.size();
    F->dev.n_trees = n_trees;
    F->dev.staged = staged;
    F->n_key = std::min(n_trees, 2);
    for (int kk = 0; kk < F->n_key; ++kk) {
        F->key[kk].nodes = (const uint2*)((const unsigned char*)F->d_blob + tree_nodes_off[kk]);
        F->key[kk].leaf_rank = (const unsigned*)F->d_leaf_rank;
        F->key[kk].root = tree_root[kk];
        F->key[kk].n_leaves = (unsigned)((*leaf_base)[kk + 1] - (*leaf_base)[kk]);
        unsigned lg = 0;
        while ((1 << lg) < block) ++lg;
        F->key[kk].feat_shift = staged ? 16u + lg : (unsigned)HM_LEFT_BITS;
        F->key[kk].left_mask = staged ? HM_PK_LEFT_MASK : HM_LEFT_MASK;
    }
    *out = F;
    return 0;
}

// Builds the packed node arrays of `F` for the all-discrete space `S` (hm_pack.cuh): every split (feature column,
// float32 threshold) becomes (shift, compare value) on the candidate's 64-bit code; splits that cannot separate any two
// values of their parameter are bypassed.  Decisions are identical to the fp32 path for every candidate of the space.
extern "C" int hm_forest_attach_packing(hm_forest_t* F, const hm_space_t* S) {
    HM_REQUIRE(F != nullptr && S != nullptr, "forest/space");
    HM_REQUIRE(S->pack_bits > 0, "the space has real / integer parameters or needs more than 64 bits: not packable");
    HM_REQUIRE(S->D_enc == F->n_features, "encoded width of the space must equal the forest's n_features");
    if (F->packed) return 0;
    // encoded column -> (parameter, category)
    std::vector<int> col_param(F->n_features, -1), col_cat(F->n_features, -1);
    for (int j = 0; j < S->D; ++j) {
        const HmPackField& f = S->pack[j];
        const int ncol = f.kind == HM_P_CATEGORICAL ? f.n_values : 1;
        for (int c = 0; c < ncol; ++c) {
            HM_REQUIRE(f.out_col + c < F->n_features && col_param[f.out_col + c] < 0, "encoded columns of the space overlap");
            col_param[f.out_col + c] = j;
            col_cat[f.out_col + c] = f.kind == HM_P_CATEGORICAL ? c : -1;
        }
    }
    // float32 encoded values of every ordinal, ascending (models.py:962-965, cast as sklearn's check_array does)
    std::vector<std::vector<float> > enc32(S->D);
    for (int j = 0; j < S->D; ++j) {
        if (S->pack[j].kind != HM_P_ORDINAL) continue;
        const hm_param_desc_t& d = S->params[j];
        enc32[j].resize(d.n_values);
        for (int r = 0; r < d.n_values; ++r) {
            enc32[j][r] = (float)S->tables[(size_t)d.values_off + d.n_values + r];
            HM_REQUIRE(r == 0 || enc32[j][r] >= enc32[j][r - 1], "encoded ordinal values must ascend");
        }
    }
    const HmHostTrees& src = *F->host_trees;
    HmHostTrees dst(src.size());
    size_t max_tree_bytes = 0;
    bool all_fast = true;                 // every surviving split shifts by <= 31: the 32-bit funnel-shift kernel applies
    for (size_t t = 0; t < src.size(); ++t) {
        const std::vector<HmHostNode>& T = src[t];
        // direction of every split on the space's value sets: 0 = both possible, 1 = always left, 2 = always right
        std::vector<unsigned char> fixed(T.size(), 0), swapped(T.size(), 0);
        std::vector<uint32_t> cmp(T.size(), 0), shift(T.size(), 0);
        for (size_t h = 0; h < T.size(); ++h) {
            if (T[h].feat < 0) continue;
            HM_REQUIRE(col_param[T[h].feat] >= 0, "a split uses an encoded column the space does not produce");
            const int j = col_param[T[h].feat];
            const HmPackField& f = S->pack[j];
            float thr;
            memcpy(&thr, &T[h].w0, 4);
            if (col_cat[T[h].feat] >= 0) {
                const int c = col_cat[T[h].feat];
                if (!(thr >= 0.0f)) fixed[h] = 2;            // 0 > thr and 1 > thr (also a NaN threshold: x <= NaN is false)
                else if (thr >= 1.0f) fixed[h] = 1;
                else if (f.n_values == 1) fixed[h] = 2;      // the column of a one-category categorical is always 1.0
                else if (c < f.n_values - 1) {               // right <=> the category's bit is set
                    shift[h] = (uint32_t)(63 - (f.low + c));
                    cmp[h] = 0x80000000u;
                } else {                                     // last category: right <=> the whole field is zero
                    shift[h] = (uint32_t)(63 - f.top);
                    cmp[h] = 1u << (32 - f.width);
                    swapped[h] = 1;
                }
            } else {
                int r1 = 0;
                while (r1 < f.n_values && enc32[j][r1] <= thr) ++r1;
                if (r1 == 0) fixed[h] = 2;
                else if (r1 == f.n_values) fixed[h] = 1;
                else {
                    shift[h] = (uint32_t)(63 - f.top);
                    cmp[h] = (uint32_t)r1 << (32 - f.width);
                }
            }
            if (!fixed[h] && shift[h] > 31) all_fast = false;
        }
        auto resolve = [&](uint32_t h) {
            while (T[h].feat >= 0 && fixed[h]) h = T[h].left + (fixed[h] == 2 ? 1u : 0u);
            return h;
        };
        // placement order of the surviving nodes (hm_forest_node_order): order[pos] = node of the fp32 tree, first_child[pos]
        // = position of its child pair
        std::vector<uint32_t> order, first_child;
        order.push_back(resolve(0));
        first_child.push_back(0);
        {
            const bool dfs = hm_forest_node_order() == 1;
            std::vector<uint32_t> work(1, 0u);
            size_t head = 0;
            while (dfs ? !work.empty() : head < work.size()) {
                uint32_t pos;
                if (dfs) {
                    pos = work.back();
                    work.pop_back();
                } else {
                    pos = work[head++];
                }
                const uint32_t h = order[pos];
                if (T[h].feat < 0) continue;
                const uint32_t at = (uint32_t)order.size();
                first_child[pos] = at;
                // first slot: taken when the test fails = sklearn's left child, unless the test is the swapped one
                order.push_back(resolve(T[h].left + (swapped[h] ? 1u : 0u)));
                order.push_back(resolve(T[h].left + (swapped[h] ? 0u : 1u)));
                first_child.push_back(0);
                first_child.push_back(0);
                if (dfs) {
                    work.push_back(at + 1);
                    work.push_back(at);
                } else {
                    work.push_back(at);
                    work.push_back(at + 1);
                }
            }
        }
        std::vector<HmHostNode>& O = dst[t];
        for (size_t k = 0; k < order.size(); ++k) {
            const uint32_t h = order[k];
            HmHostNode nd;
            if (T[h].feat < 0) {
                nd = T[h];
            } else {
                nd.w0 = cmp[h];
                nd.feat = (int32_t)shift[h];
                nd.left = first_child[k];
            }
            O.push_back(nd);
        }
        max_tree_bytes = std::max(max_tree_bytes, O.size() * 8 + 16);
    }
    int nodebuf_p = HM_PACKED_NODEBUF;      // experiments: a smaller buffer leaves part of the 228 KB to the L1 cache
    if (const char* e = getenv("HM_B200_PACKED_NODEBUF")) nodebuf_p = std::min(HM_PACKED_NODEBUF, std::max(4096, atoi(e) & ~15));
    const int staged = max_tree_bytes <= (size_t)nodebuf_p ? 1 : 0;
    std::vector<HmGroupDesc> groups;
    std::vector<unsigned char> blob;
    std::vector<unsigned> tree_root;
    std::vector<long long> tree_nodes_off;
    // self-looping leaves: a FAST space with a sentinel field (every forest of the space then takes the same kind)
    unsigned loop_shift = 0, loop_bound = 0;
    const bool loop = all_fast && S->pack_fast && hm_pack_sentinel(S->pack.data(), S->D, &loop_shift, &loop_bound) &&
                      (unsigned long long)loop_bound + (unsigned long long)F->n_leaves < 0xFFFFFFFFull && !getenv("HM_B200_PACKED_NOLOOP");
    int rc = hm_build_blob(dst, loop ? HM_LAYOUT_PACKED_LOOP : HM_LAYOUT_PACKED, nodebuf_p, 0, staged != 0, blob, groups,
                           tree_root, tree_nodes_off, loop_shift, loop_bound);
    if (rc) return rc;
    F->leaf_smem_p = 0;
    if (staged && groups.size() <= 2 && F->host_leaf_tab && !getenv("HM_B200_NO_LEAF_SMEM")) {
        std::vector<HmGroupDesc> g2;
        std::vector<unsigned char> b2;
        std::vector<unsigned> r2;
        std::vector<long long> o2;
        rc = hm_build_blob(dst, loop ? HM_LAYOUT_PACKED_LOOP : HM_LAYOUT_PACKED, nodebuf_p, 0, true, b2, g2, r2, o2, loop_shift,
                           loop_bound, F->host_leaf_tab->data(), F->tree_leaf_base);
        if (rc) return rc;
        if (g2.size() <= 2) {
            groups.swap(g2);
            blob.swap(b2);
            tree_root.swap(r2);
            tree_nodes_off.swap(o2);
            F->leaf_smem_p = 1;
        }
    }
    HM_CUDA_TRY(cudaMalloc(&F->d_blob_p, blob.size()));
    HM_CUDA_TRY(cudaMalloc(&F->d_groups_p, groups.size() * sizeof(HmGroupDesc)));
    HM_CUDA_TRY(cudaMemcpy(F->d_blob_p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    HM_CUDA_TRY(cudaMemcpy(F->d_groups_p, groups.data(), groups.size() * sizeof(HmGroupDesc), cudaMemcpyHostToDevice));
    HM_CUDA_TRY(cudaDeviceSynchronize());
    F->devp = F->dev;
    F->devp.blob = (const unsigned char*)F->d_blob_p;
    F->devp.groups = (const HmGroupDesc*)F->d_groups_p;
    F->devp.n_groups = (int)groups.size();
    F->devp.staged = staged;
    if (loop) {
        // word 0 of a leaf is bound + ordinal: the tables are addressed through pointers moved down by `bound` entries
        F->devp.leaf_tab = reinterpret_cast<const double*>(reinterpret_cast<uintptr_t>(F->d_leaf_tab) -
                                                           (uintptr_t)loop_bound * 32u);
        F->devp.leaf_node = reinterpret_cast<const int*>(reinterpret_cast<uintptr_t>(F->d_leaf_node) -
                                                         (uintptr_t)loop_bound * sizeof(int));
    }
    for (int kk = 0; kk < F->n_key; ++kk) {
        F->keyp[kk] = F->key[kk];
        F->keyp[kk].nodes = (const uint2*)((const unsigned char*)F->d_blob_p + tree_nodes_off[kk]);
        F->keyp[kk].root = tree_root[kk];
        F->keyp[kk].leaf_and = loop ? 32u : 0xffffffffu;
        F->keyp[kk].leaf_eq = loop ? 32u : 0u;
        if (loop)
            F->keyp[kk].leaf_rank = reinterpret_cast<const unsigned*>(reinterpret_cast<uintptr_t>(F->d_leaf_rank) -
                                                                      (uintptr_t)loop_bound * sizeof(unsigned));
    }
    F->packed_bits = S->pack_bits;
    F->nodebuf_p = nodebuf_p;
    F->packed = loop ? 3 : (all_fast ? 1 : 2);
    F->loop_shift = loop_shift;
    F->loop_bound = loop_bound;
    return 0;
}

extern "C" void hm_forest_destroy(hm_forest_t* F) {
    if (!F) return;
    cudaFree(F->d_blob);
    cudaFree(F->d_groups);
    cudaFree(F->d_leaf_tab);
    cudaFree(F->d_leaf_node);
    cudaFree(F->d_leaf_rank);
    cudaFree(F->d_blob_p);
    cudaFree(F->d_groups_p);
    delete F->host_trees;
    delete F->tree_leaf_base;
    delete F->host_leaf_tab;
    delete F;
}
extern "C" int hm_forest_n_trees(const hm_forest_t* f) { return f ? f->n_trees : 0; }
extern "C" int hm_forest_n_features(const hm_forest_t* f) { return f ? f->n_features : 0; }
extern "C" int hm_forest_staged(const hm_forest_t* f) { return f ? f->staged : 0; }
extern "C" int hm_forest_n_groups(const hm_forest_t* f) { return f ? f->dev.n_groups : 0; }
extern "C" int hm_forest_packed(const hm_forest_t* f) { return f ? f->packed : 0; }
extern "C" int hm_forest_packed_groups(const hm_forest_t* f) { return (f && f->packed) ? f->devp.n_groups : 0; }
extern "C" int hm_forest_leaf_resident(const hm_forest_t* f, int packed) {
    return f ? (packed ? (f->packed ? f->leaf_smem_p : 0) : f->leaf_smem) : 0;
}

// ------------------------------------------------------------------------------------------------------------
// Running sums of one objective + the leaf records of the previous round of trees, still in flight: a round's leaf
// gathers (L2 latency, ~30 % of a round when waited for on the spot) are issued when its walk ends and added -- in tree
// order -- only after the NEXT round's walk, so their latency hides behind that walk.
#ifndef HM_MAX_CHAINS
#define HM_MAX_CHAINS 3
#endif
// Leaf records are 32 bytes (q, s, u, value), gathered from global memory through L1: 128 + 64 bits when `value` is not
// needed.  HM_LEAF256 = 1 fetches them with ONE 256-bit load whose last column ptxas masks off (LDG.E.ENL2.256, register mask
// 0x3f) -- measured: wins where the kernel is issue-bound (small resident forests; C1 3.04 -> 2.77 ms), loses where the L1
// data pipe is the limit (streamed forests: C3 +0.9 %, C5-RF +1.9 %); a run-time switch between the two forms inside one
// kernel costs the streamed forests 1-4 % (both loads are emitted predicated).  Resident forests therefore got their own
// kernel mode (NVM = 2) that does not gather at all: their groups carry the leaf records into shared memory.
#ifndef HM_LEAF256
#define HM_LEAF256 0
#endif
#ifndef HM_LEAF_LD
#define HM_LEAF_LD "ld.global.nc"      // experiments: "ld.global.cg" (L2 only), "ld.global.nc.L1::evict_last" ...
#endif
struct HmAcc {
    double q, s, u, v;
    double pq[HM_MAX_CHAINS], ps[HM_MAX_CHAINS], pu[HM_MAX_CHAINS], pv[HM_MAX_CHAINS];
};
// All HM_MAX_CHAINS pending slots are added on every flush, in slot (= tree) order; a slot that holds no record holds
// +0.0, and x + 0.0 == x bit for bit (only -0.0 becomes +0.0, which compares equal), so the sums are those of the
// reference's tree-order loop without a data-dependent slot count in the instruction stream.
template <int NVM>
__device__ __forceinline__ void hm_acc_flush(HmAcc& acc) {
#pragma unroll
    for (int c = 0; c < HM_MAX_CHAINS; ++c) {
        acc.q = hm_add(acc.q, acc.pq[c]);
        acc.s = hm_add(acc.s, acc.ps[c]);
        acc.u = hm_add(acc.u, acc.pu[c]);
        if (NVM == 1) acc.v = hm_add(acc.v, acc.pv[c]);
    }
}
__device__ __forceinline__ void hm_acc_clear_pending(HmAcc& acc) {
#pragma unroll
    for (int c = 0; c < HM_MAX_CHAINS; ++c) acc.pq[c] = acc.ps[c] = acc.pu[c] = acc.pv[c] = 0.0;
}

#ifndef HM_FOREST_MAXCH
#define HM_FOREST_MAXCH 3
#endif
// Walk NCH consecutive trees of a group at once.  The walk is a chain of dependent loads (node -> feature -> node
// ...); the chains of different trees are independent, so their loads overlap; a chain that has reached its leaf
// idles (predicated off) until the slowest of the NCH is done.  The leaf records are then added in tree order, so the
// fp64 sums are unchanged.
template <bool APPLY, int NVM, bool SMEM, int NCH, int PACKED>
__device__ __forceinline__ void hm_walk_round(const unsigned long long* __restrict__ roots, const uint2* __restrict__ nodes,
                                              int t0, const float* __restrict__ xs_t, int BD,
                                              unsigned long long code, HmRtConst k,
                                              const double* __restrict__ leaf_tab, const int* __restrict__ leaf_node,
                                              HmAcc& acc, int* __restrict__ leaf_out, long long N, long long i,
                                              bool valid, int first_tree) {
    constexpr bool NEED_V = NVM == 1;      // NVM: 0 = leaf records as 128 + 64 bits, 1 = all four columns, 2 = the groups carry
                                           // their leaf records: 128 + 64 bits from SHARED memory (`leaf_tab` = shared address)
    uint2 nd[NCH];
    if (SMEM && PACKED) {
        // packed candidates: the split is a shift and an unsigned compare on the 64-bit code held in a register pair
        // (hm_pack.cuh); word 1 = (byte offset of the first child slot) << 6 | shift, 0 at a leaf.  One shared-memory load
        // per visit (the node).  The loop is bound by the alu pipe (one warp instruction per two cycles; shift, leaf test,
        // compare, predicated add), so everything that can run elsewhere does: the slot address is an IMAD.HI with its
        // multiplier in a register (fma pipe; as a literal it would be strength-reduced to an alu-pipe LEA.HI), the node
        // stays in one aligned register pair (no moves around the load) and the loop condition is one three-input OR.
        const uint32_t nodes_a = hm_smem_u32(nodes);
        const uint32_t c_lo = (uint32_t)code, c_hi = (uint32_t)(code >> 32);
        const int nodebuf_bytes = k.nodebuf;
        (void)nodebuf_bytes;
        unsigned long long nq[NCH];
        uint32_t yy[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            nq[c] = roots[t0 + c];
            asm volatile("{\n\t.reg .b32 lo;\n\tmov.b64 {lo, %0}, %1;\n\t}" : "=r"(yy[c]) : "l"(nq[c]));
        }
        if (PACKED == 3) {
            // self-looping leaves (hm_pack.cuh): no leaf predicate, five instructions per visit -- funnel shift, compare,
            // slot address, +8, load; a chain that has arrived re-loads its leaf (the sentinel test is false for every
            // candidate of the space) until the slowest chain of the warp is there.  "Everyone has arrived" = bit 5 of
            // every word 1, tested once per HM_LOOP_LEVELS levels.
            uint32_t arrived;
            do {
#pragma unroll
                for (int lvl = 0; lvl < HM_LOOP_LEVELS; ++lvl) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        HM_DEV_ASSERT(((uint32_t)(nq[c] >> 32) >> 6) + 16u <= (unsigned)nodebuf_bytes);
                        asm volatile(
                            "{\n\t"
                            ".reg .pred q;\n\t"
                            ".reg .b32 th, y, hi, na;\n\t"
                            "mov.b64 {th, y}, %0;\n\t"
                            "shf.l.wrap.b32 hi, %2, %3, y;\n\t"        // hi32(code << (y & 31))
                            "mad.hi.u32 na, y, %4, %1;\n\t"            // nodes + (y >> 6)
                            "setp.ge.u32 q, hi, th;\n\t"               // second slot <=> hi32(code << s) >= T
                            "@q add.u32 na, na, 8;\n\t"
                            "ld.shared.b64 %0, [na];\n\t"
                            "}"
                            : "+l"(nq[c])
                            : "r"(nodes_a), "r"(c_lo), "r"(c_hi), "r"(k.k26));
                    }
                }
                arrived = 32u;
#pragma unroll
                for (int c = 0; c < NCH; ++c) arrived &= (uint32_t)(nq[c] >> 32);
            } while (!arrived);
        } else {
        uint32_t live = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) live |= yy[c];
        while (live) {
#pragma unroll
            for (int lvl = 0; lvl < HM_WALK_LEVELS; ++lvl) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                HM_DEV_ASSERT(yy[c] == 0u || (yy[c] >> 6) + 16u <= (unsigned)nodebuf_bytes);     // both child slots inside the staged group
                if (PACKED == 1) {
                    asm volatile(
                        "{\n\t"
                        ".reg .pred p, q;\n\t"
                        ".reg .b32 th, y, hi, na;\n\t"
                        "mov.b64 {th, y}, %0;\n\t"
                        "setp.ne.u32 p, y, 0;\n\t"
                        "shf.l.wrap.b32 hi, %3, %4, y;\n\t"          // hi32(code << (y & 31)): every shift is <= 31
                        "mad.hi.u32 na, y, %5, %2;\n\t"              // nodes + (y >> 6)
                        "setp.ge.u32 q, hi, th;\n\t"                 // second slot <=> hi32(code << s) >= T
                        "@q add.u32 na, na, 8;\n\t"
                        "@p ld.shared.b64 %0, [na];\n\t"
                        "mov.b64 {th, %1}, %0;\n\t"
                        "}"
                        : "+l"(nq[c]), "=r"(yy[c])
                        : "r"(nodes_a), "r"(c_lo), "r"(c_hi), "r"(k.k26));
                } else {
                    asm volatile(
                        "{\n\t"
                        ".reg .pred p, q;\n\t"
                        ".reg .b32 th, y, sft, lo, hi, na;\n\t"
                        ".reg .b64 sh;\n\t"
                        "mov.b64 {th, y}, %0;\n\t"
                        "setp.ne.u32 p, y, 0;\n\t"
                        "and.b32 sft, y, 63;\n\t"
                        "shl.b64 sh, %3, sft;\n\t"
                        "mov.b64 {lo, hi}, sh;\n\t"
                        "mad.hi.u32 na, y, %4, %2;\n\t"
                        "setp.ge.u32 q, hi, th;\n\t"
                        "@q add.u32 na, na, 8;\n\t"
                        "@p ld.shared.b64 %0, [na];\n\t"
                        "mov.b64 {th, %1}, %0;\n\t"
                        "}"
                        : "+l"(nq[c]), "=r"(yy[c])
                        : "r"(nodes_a), "l"(code), "r"(k.k26));
                }
            }
            }
            live = 0;
#pragma unroll
            for (int c = 0; c < NCH; ++c) live |= yy[c];
        }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) nd[c] = make_uint2((uint32_t)nq[c], (uint32_t)(nq[c] >> 32));
    } else if (PACKED) {
        // packed candidates, trees too large for a node buffer: walked in L2
        const unsigned char* nb = reinterpret_cast<const unsigned char*>(nodes);
#pragma unroll
        for (int c = 0; c < NCH; ++c) nd[c] = reinterpret_cast<const uint2*>(roots)[t0 + c];
        bool any = false;
#pragma unroll
        for (int c = 0; c < NCH; ++c) any = any || ((nd[c].y & k.leaf_and) != k.leaf_eq);
        while (any) {
            any = false;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if ((nd[c].y & k.leaf_and) != k.leaf_eq) {
                    const uint32_t hi = (uint32_t)((code << (nd[c].y & 63u)) >> 32);
                    nd[c] = *reinterpret_cast<const uint2*>(nb + (nd[c].y >> 6) + (hi >= nd[c].x ? 8u : 0u));
                    any = any || ((nd[c].y & k.leaf_and) != k.leaf_eq);
                }
            }
        }
    } else if (SMEM) {
        // shared-memory forest.  One visit = one PTX block per chain (ptxas interleaves the chains): both loads are
        // predicated on "this chain is still at an internal node", so a finished chain issues no load and no branch;
        // the node lives in one 64-bit register {threshold bits | leaf ordinal, feature << 22 | left}.
        const uint32_t xs_a = hm_smem_u32(xs_t), nodes_a = hm_smem_u32(nodes);
        unsigned long long nq[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) nq[c] = roots[t0 + c];
        bool any = false;
#pragma unroll
        for (int c = 0; c < NCH; ++c) any = any || ((uint32_t)(nq[c] >> 32) & HM_PK_LEFT_MASK);
        while (any) {
#pragma unroll
            for (int lvl = 0; lvl < HM_WALK_LEVELS; ++lvl) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                // both child slots inside the staged group, feature column inside the tile
                HM_DEV_ASSERT((((uint32_t)(nq[c] >> 32) & HM_PK_LEFT_MASK) * 8u + 16u <= (unsigned)k.nodebuf) &&
                              ((uint32_t)(nq[c] >> 32) >> 14) + 4u <= (unsigned)k.xtile_bytes);
                // packed word 1: tile byte offset = y >> 14 (as the high half of y * 2^18, one IMAD.HI), child = y & 0xffff
                asm volatile(
                    "{\n\t"
                    ".reg .pred p, q;\n\t"
                    ".reg .b32 th, y, left, xa, na;\n\t"
                    ".reg .f32 xv, thf;\n\t"
                    "mov.b64 {th, y}, %0;\n\t"
                    "and.b32 left, y, 0xffff;\n\t"
                    "setp.ne.u32 p, left, 0;\n\t"
                    "mad.hi.u32 xa, y, 0x40000, %1;\n\t"
                    "@p ld.shared.f32 xv, [xa];\n\t"
                    "mov.b32 thf, th;\n\t"
                    "mad.lo.u32 na, left, 8, %2;\n\t"
                    "setp.gtu.f32 q, xv, thf;\n\t"          // sklearn: go left iff x <= threshold (NaN -> right)
                    "@q add.u32 na, na, 8;\n\t"
                    "@p ld.shared.b64 %0, [na];\n\t"
                    "}"
                    : "+l"(nq[c])
                    : "r"(xs_a), "r"(nodes_a));
            }
            }
            any = false;
#pragma unroll
            for (int c = 0; c < NCH; ++c) any = any || ((uint32_t)(nq[c] >> 32) & HM_PK_LEFT_MASK);
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) nd[c] = make_uint2((uint32_t)nq[c], (uint32_t)(nq[c] >> 32));
    } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c) nd[c] = reinterpret_cast<const uint2*>(roots)[t0 + c];
        bool any = false;
#pragma unroll
        for (int c = 0; c < NCH; ++c) any = any || (nd[c].y & HM_LEFT_MASK);
        while (any) {
            any = false;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (nd[c].y & HM_LEFT_MASK) {
                    const float xv = xs_t[(nd[c].y >> HM_LEFT_BITS) * BD];
                    const uint32_t nxt = (nd[c].y & HM_LEFT_MASK) + ((xv <= __uint_as_float(nd[c].x)) ? 0u : 1u);
                    nd[c] = nodes[nxt];
                    any = any || (nd[c].y & HM_LEFT_MASK);
                }
            }
        }
    }
#ifdef HM_CHECKED
#pragma unroll
    for (int c = 0; c < NCH; ++c) HM_DEV_ASSERT(PACKED ? (nd[c].y & k.leaf_and) == k.leaf_eq : (nd[c].y & (SMEM ? HM_PK_LEFT_MASK : HM_LEFT_MASK)) == 0u);
#endif
    // add the previous round's leaf records (they had this whole walk to arrive), then put this round's in flight:
    // one 256-bit load per 32-byte record (LDG.E.256), or 128 + 64 bits when the `value` column is not needed
    hm_acc_flush<NVM>(acc);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if (NVM == 2) {
            const uint32_t ra = (uint32_t)reinterpret_cast<uintptr_t>(leaf_tab) + nd[c].x * 32u;
            HM_DEV_ASSERT(ra >= hm_smem_u32(nodes) && ra + 32u <= hm_smem_u32(nodes) + (unsigned)k.nodebuf);
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(acc.pq[c]), "=d"(acc.ps[c]) : "r"(ra));
            asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(acc.pu[c]) : "r"(ra));
        } else {
        const double* rec;
        asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(rec) : "r"(nd[c].x), "l"(leaf_tab));
        if (NEED_V || HM_LEAF256) {
            asm volatile(HM_LEAF_LD ".v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(acc.pq[c]), "=d"(acc.ps[c]), "=d"(acc.pu[c]), "=d"(acc.pv[c])
                         : "l"(rec));
        } else {
            asm volatile(HM_LEAF_LD ".v2.f64 {%0,%1}, [%2];" : "=d"(acc.pq[c]), "=d"(acc.ps[c]) : "l"(rec));
            asm volatile(HM_LEAF_LD ".f64 %0, [%1];" : "=d"(acc.pu[c]) : "l"(rec + 2));
        }
        }
        if (APPLY) {
            if (valid && leaf_out)
                leaf_out[(size_t)(first_tree + t0 + c) * (size_t)N + (size_t)i] = __ldg(leaf_node + nd[c].x);
        }
    }
    // slots of chains this round does not have go back to +0.0
#pragma unroll
    for (int c = NCH; c < HM_MAX_CHAINS; ++c) acc.pq[c] = acc.ps[c] = acc.pu[c] = acc.pv[c] = 0.0;
}

// Walk `ntrees` trees whose roots/nodes are at the given pointers (shared or global), in tree order: rounds of 3,
// then what is left as a round of 2 and/or 1 (4 = 2 + 2).
template <bool APPLY, int NVM, bool SMEM, int PACKED>
__device__ __forceinline__ void hm_walk_group(const unsigned long long* __restrict__ roots, const uint2* __restrict__ nodes,
                                              int ntrees, const float* __restrict__ xs_t, int BD, unsigned long long code,
                                              HmRtConst k,
                                              const double* __restrict__ leaf_tab, const int* __restrict__ leaf_node,
                                              HmAcc& acc, int* __restrict__ leaf_out, long long N, long long i,
                                              bool valid, int first_tree) {
    int t = 0;
#define HM_ROUND(K) hm_walk_round<APPLY, NVM, SMEM, K, PACKED>(roots, nodes, t, xs_t, BD, code, k, leaf_tab, leaf_node, acc, leaf_out, N, i, valid, first_tree)
#if HM_FOREST_MAXCH >= 5
    while (ntrees - t >= 5) { HM_ROUND(5); t += 5; }
#endif
#if HM_FOREST_MAXCH >= 4
    while (ntrees - t >= 4) { HM_ROUND(4); t += 4; }
#endif
    while (ntrees - t >= 3 && (HM_FOREST_MAXCH >= 4 || ntrees - t != 4)) { HM_ROUND(3); t += 3; }
    while (ntrees - t >= 2) { HM_ROUND(2); t += 2; }
    if (t < ntrees) HM_ROUND(1);
#undef HM_ROUND
}

template <bool APPLY, int NVM, int BDT, int PACKED>
__global__ void __launch_bounds__(BDT, (PACKED && BDT < 1024) ? 1024 / BDT : 1) hm_forest_kernel(const __grid_constant__ HmEvalArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* nbuf0 = smem + HM_SMEM_HEADER;
    float* xs = reinterpret_cast<float*>(smem + HM_SMEM_HEADER + 2 * (size_t)a.nodebuf);     // unused when PACKED

    const int tid = threadIdx.x;
    constexpr int BD = BDT;      // compile-time block size: the x-tile stride folds into the address arithmetic
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
        HM_DEV_ASSERT(bytes <= (unsigned)a.nodebuf && (bytes & 15u) == 0u && pg < a.f[pm].n_groups && pm < a.M);
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
        const long long slot = tile * BD + tid;
        const bool valid = slot < a.N;
        const long long sl = valid ? slot : a.N - 1;
        // i = the candidate this thread evaluates: arrival order, or the coherence order of the pre-pass
        const long long i = a.perm ? (long long)__ldg(a.perm + sl) : sl;
        unsigned long long code = 0ull;
        if (PACKED) {
            code = __ldg(a.codes + sl);          // arrival order, or the bucket-ordered copy of the pre-pass (with perm)
            if (PACKED == 3) {
                // self-looping leaves rely on the sentinel field's bound: a malformed code (not one of the space's
                // candidates) is replaced by the all-zero code instead of walking out of the staged nodes
                const uint32_t hi = (uint32_t)((code << a.k.loop_shift) >> 32);
                if (hi >= a.k.loop_bound) code = 0ull;
            }
        } else if (a.Xrows) {
            // the pre-pass left this slot's candidate as one contiguous row: Dp/4 128-bit loads
            const float4* row = reinterpret_cast<const float4*>(a.Xrows + (size_t)sl * (size_t)a.Dp);
            for (int q = 0; 4 * q < a.D; ++q) {
                const float4 v = __ldg(row + q);
                xs_t[(4 * q) * BD] = v.x;
                if (4 * q + 1 < a.D) xs_t[(4 * q + 1) * BD] = v.y;
                if (4 * q + 2 < a.D) xs_t[(4 * q + 2) * BD] = v.z;
                if (4 * q + 3 < a.D) xs_t[(4 * q + 3) * BD] = v.w;
            }
        } else {
            for (int f = 0; f < a.D; ++f) xs_t[f * BD] = __ldg(a.X + (size_t)f * (size_t)a.ldx + (size_t)i);
        }

        double mean_l[HM_MAX_OBJECTIVES], var_l[HM_MAX_OBJECTIVES];
        int step_in_tile = 0;
        for (int m = 0; m < a.M; ++m) {
            const HmForestDev& F = a.f[m];
            HmAcc acc;
            acc.q = acc.s = acc.u = acc.v = 0.0;
            hm_acc_clear_pending(acc);
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
                    const double* lt = F.leaf_tab;
                    if (NVM == 2) {
                        // the group's own leaf records: shared-memory address of the (virtual) record of ordinal 0
                        // (kind 3: of ordinal -loop_bound, as word 0 of a self-looping leaf is bound + ordinal)
                        const unsigned loff = __ldg(&F.groups[g].leaf_off), lbase = __ldg(&F.groups[g].leaf_base);
                        HM_DEV_ASSERT(loff != 0u);
                        lt = reinterpret_cast<const double*>((uintptr_t)(hm_smem_u32(gb) + loff - (lbase + a.k.loop_bound) * 32u));
                    }
                    hm_walk_group<APPLY, NVM, true, PACKED>(reinterpret_cast<const unsigned long long*>(gb),
                                         reinterpret_cast<const uint2*>(gb + noff), (int)ntr, xs_t, BD, code, a.k, lt,
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
                    hm_walk_group<APPLY, NVM, false, PACKED>(reinterpret_cast<const unsigned long long*>(gb),
                                         reinterpret_cast<const uint2*>(gb + gd.nodes_off), (int)gd.n_trees, xs_t, BD,
                                         code, a.k, F.leaf_tab, F.leaf_node, acc, lo, a.N, i, valid, (int)gd.first_tree);
                }
            }
            hm_acc_flush<NVM>(acc);
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
            const double sc = hm_acq_value(a.acq, mean_l, var_l, feas);
            if (a.filter_tau) {
                const double filter_tau = __ldg(a.filter_tau);
                if (sc < filter_tau || sc != sc || (sc == filter_tau && i < a.filter_n0)) {
                    const unsigned pos = atomicAdd(a.list_count, 1u);
                    if (pos < a.list_cap) {
                        a.list_val[pos] = sc;
                        a.list_idx[pos] = i;
                    }
                }
            } else {
                a.score_out[i] = sc;
            }
        }
    }
}

// ---- coherence pre-pass ----------------------------------------------------------------------------------------
#define HM_SORT_MAX_BITS 22
#define HM_SORT_MIN_N (1 << 17)

struct HmKeyArgs {
    HmKeyTree t[2];
    unsigned shift[2];
    unsigned nB;            // bins of the second component (1 if there is only one key tree)
    int n_key;
    const float* X;
    long long ldx, N;
};

// key[i] = (leaf rank in tree A >> shiftA) * nB + (leaf rank in tree B >> shiftB); histogram of the keys
__global__ void __launch_bounds__(256) hm_forest_sortkey(const HmKeyArgs a, unsigned* __restrict__ key,
                                                         unsigned* __restrict__ hist) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    unsigned k = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (q < a.n_key) {
            const uint2* __restrict__ nodes = a.t[q].nodes;
            uint2 nd = __ldg(nodes + a.t[q].root);
            const unsigned fs = a.t[q].feat_shift, lm = a.t[q].left_mask;
            while (nd.y & lm) {
                const float xv = __ldg(a.X + (size_t)(nd.y >> fs) * (size_t)a.ldx + (size_t)i);
                nd = __ldg(nodes + (nd.y & lm) + ((xv <= __uint_as_float(nd.x)) ? 0u : 1u));
            }
            const unsigned r = __ldg(a.t[q].leaf_rank + nd.x) >> a.shift[q];
            k = (q == 0) ? r : k * a.nB + r;
        }
    }
    key[i] = k;
    atomicAdd(hist + k, 1u);
}

// exclusive scan of `n` counters in place, one CTA (n <= 2^22: a few passes over an L2-resident array)
// Exclusive scan of the bucket histogram, one launch over up to hm_num_sms() CTAs: a CTA takes a ticket (CTAs with a lower
// ticket are therefore already running -- no co-residency assumption), sums its chunk, publishes the sum, adds up the sums
// of the lower tickets as they appear, and scans its chunk with that carry.  `sync` = [ticket, flags[grid], sums[grid]], zeroed.
__global__ void __launch_bounds__(1024, 1) hm_forest_scan(unsigned* __restrict__ v, unsigned n, unsigned* __restrict__ sync) {
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry, ticket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) ticket = atomicAdd(sync, 1u);
    __syncthreads();
    const unsigned b = ticket, nb = gridDim.x;
    volatile unsigned* flags = sync + 1;
    volatile unsigned* sums = sync + 1 + nb;
    const unsigned chunk = ((n + nb - 1) / nb + 4095u) & ~4095u;
    const unsigned lo = min(n, b * chunk), hi = min(n, lo + chunk);
    auto block_sum = [&](unsigned x) {            // sum over the CTA, returned to every thread
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        __syncthreads();
        if (lane == 0) warp_sums[warp] = x;
        __syncthreads();
        unsigned s = warp_sums[lane];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        return s;
    };
    unsigned mine = 0;
    for (unsigned i = lo + threadIdx.x; i < hi; i += 1024) mine += v[i];
    const unsigned total = block_sum(mine);
    if (threadIdx.x == 0) {
        sums[b] = total;
        __threadfence();
        flags[b] = 1u;
    }
    unsigned prev = 0;
    if (threadIdx.x < b) {
        while (flags[threadIdx.x] == 0u) {}
        __threadfence();
        prev = sums[threadIdx.x];
    }
    const unsigned before = block_sum(prev);
    __syncthreads();
    if (threadIdx.x == 0) carry = before;
    __syncthreads();
    for (unsigned base = lo; base < hi; base += 4096) {
        // 4 consecutive counters per thread
        const unsigned i0 = base + threadIdx.x * 4;
        unsigned c[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) c[q] = (i0 + q < hi) ? v[i0 + q] : 0;
        const unsigned own = c[0] + c[1] + c[2] + c[3];
        unsigned x = own;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned s = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        unsigned run = carry + (warp ? warp_sums[warp - 1] : 0) + (x - own);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (i0 + q < hi) v[i0 + q] = run;
            run += c[q];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = run;
        __syncthreads();
    }
}

// pos = offset[key[i]]++;  perm[pos] = i;  rows[pos] = candidate i   (order inside a bucket is arbitrary: it cannot
// influence any result).  The column-major pool is read coalesced; each candidate leaves as Dp/4 128-bit stores into
// its own row, so the traversal kernel's tile load is a plain coalesced row read.
__global__ void __launch_bounds__(256) hm_forest_scatter_rows(const unsigned* __restrict__ key, unsigned* __restrict__ offs,
                                                              const float* __restrict__ X, long long ldx, long long N,
                                                              int D, int Dp, unsigned* __restrict__ perm,
                                                              float* __restrict__ rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned pos = atomicAdd(offs + key[i], 1u);
    perm[pos] = (unsigned)i;
    float4* row = reinterpret_cast<float4*>(rows + (size_t)pos * (size_t)Dp);
    for (int q = 0; 4 * q < D; ++q) {
        float4 v;
        v.x = __ldg(X + (size_t)(4 * q) * (size_t)ldx + (size_t)i);
        v.y = (4 * q + 1 < D) ? __ldg(X + (size_t)(4 * q + 1) * (size_t)ldx + (size_t)i) : 0.0f;
        v.z = (4 * q + 2 < D) ? __ldg(X + (size_t)(4 * q + 2) * (size_t)ldx + (size_t)i) : 0.0f;
        v.w = (4 * q + 3 < D) ? __ldg(X + (size_t)(4 * q + 3) * (size_t)ldx + (size_t)i) : 0.0f;
        row[q] = v;
    }
}

// packed candidates: the same key, the key trees walked with the packed split test
__global__ void __launch_bounds__(256) hm_forest_sortkey_packed(const HmKeyArgs a, const unsigned long long* __restrict__ codes,
                                                                unsigned* __restrict__ key, unsigned* __restrict__ hist) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const unsigned long long code = __ldg(codes + i);
    unsigned k = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (q < a.n_key) {
            const unsigned char* nb = reinterpret_cast<const unsigned char*>(a.t[q].nodes);
            uint2 nd = __ldg(a.t[q].nodes + a.t[q].root);
            while ((nd.y & a.t[q].leaf_and) != a.t[q].leaf_eq) {
                const uint32_t hi = (uint32_t)((code << (nd.y & 63u)) >> 32);
                nd = __ldg(reinterpret_cast<const uint2*>(nb + (nd.y >> 6) + (hi >= nd.x ? 8u : 0u)));
            }
            const unsigned r = __ldg(a.t[q].leaf_rank + nd.x) >> a.shift[q];
            k = (q == 0) ? r : k * a.nB + r;
        }
    }
    key[i] = k;
    atomicAdd(hist + k, 1u);
}

// pos = offset[key[i]]++;  perm[pos] = i;  sorted[pos] = code i
__global__ void __launch_bounds__(256) hm_forest_scatter_codes(const unsigned* __restrict__ key, unsigned* __restrict__ offs,
                                                               const unsigned long long* __restrict__ codes, long long N,
                                                               unsigned* __restrict__ perm,
                                                               unsigned long long* __restrict__ sorted) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned pos = atomicAdd(offs + key[i], 1u);
    perm[pos] = (unsigned)i;
    sorted[pos] = __ldg(codes + i);
}

extern "C" int hm_forest_set_coherence(hm_forest_t* f, int mode) {
    HM_REQUIRE(f != nullptr, "forest");
    HM_REQUIRE(mode >= -1 && mode <= 1, "mode in {-1 auto, 0 off, 1 on}");
    f->sort_mode = mode;
    return 0;
}

// keep freed stream-ordered blocks in the pool: the next call reuses them without a driver round trip (once per device)
static void hm_forest_keep_pool() {
    static unsigned done_mask = 0;                  // bit per device; a benign race sets the same attribute twice
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 32 && (done_mask >> dev & 1u)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    if (dev < 32) done_mask |= 1u << dev;
}

// builds the coherence permutation of the N candidates in *perm_out (stream-ordered allocation, freed by the caller);
// fp32 path: candidates also copied as contiguous rows in bucket order; packed path: the codes in bucket order
static int hm_forest_build_perm(const hm_forest_t* const* forests, int M, const float* Xcm, int64_t ldx,
                                const unsigned long long* codes, int64_t N, cudaStream_t st, unsigned** perm_out,
                                float** rows_out, int* Dp_out, unsigned long long** codes_out, void** scratch_out) {
    const bool packed = codes != nullptr;
    HmKeyArgs ka;
    memset(&ka, 0, sizeof(ka));
    // key trees: one from each of the first two forests, or the first two trees of a single forest
    if (M >= 2) {
        ka.t[0] = packed ? forests[0]->keyp[0] : forests[0]->key[0];
        ka.t[1] = packed ? forests[1]->keyp[0] : forests[1]->key[0];
        ka.n_key = 2;
    } else {
        ka.n_key = forests[0]->n_key;
        for (int q = 0; q < ka.n_key; ++q) ka.t[q] = packed ? forests[0]->keyp[q] : forests[0]->key[q];
    }
    unsigned bits[2] = {0, 0};
    for (int q = 0; q < ka.n_key; ++q)
        while ((1u << bits[q]) < ka.t[q].n_leaves) ++bits[q];
    while (bits[0] + bits[1] > HM_SORT_MAX_BITS) {          // coarsen the larger component first
        const int q = (bits[0] - ka.shift[0] >= bits[1] - ka.shift[1]) ? 0 : 1;
        ++ka.shift[q];
        --bits[q];
    }
    const unsigned nA = ((ka.t[0].n_leaves - 1) >> ka.shift[0]) + 1;
    ka.nB = (ka.n_key == 2) ? ((ka.t[1].n_leaves - 1) >> ka.shift[1]) + 1 : 1;
    const size_t bins = (size_t)nA * ka.nB;
    ka.X = Xcm;
    ka.ldx = ldx;
    ka.N = N;
    hm_forest_keep_pool();
    unsigned* buf = nullptr;
    const int D = forests[0]->n_features, Dp = packed ? 2 : (D + 3) & ~3;      // words per candidate in the sorted copy
    const size_t rows_words = (size_t)N * Dp;                       // 16-byte aligned rows at the front
    const int scan_grid = (int)std::min<size_t>((size_t)std::min(hm_num_sms(), 512), (bins + 4095) / 4096);
    const size_t sync_words = 1 + 2 * (size_t)scan_grid;               // ticket, flags, sums of hm_forest_scan
    const size_t words = rows_words + (size_t)N * 2 + bins + sync_words;
    HM_CUDA_TRY(cudaMallocAsync((void**)&buf, words * sizeof(unsigned), st));
    unsigned* key = buf + rows_words;
    unsigned* perm = key + N;
    unsigned* hist = perm + N;
    unsigned* scan_sync = hist + bins;
    HM_CUDA_TRY(cudaMemsetAsync(hist, 0, (bins + sync_words) * sizeof(unsigned), st));
    const unsigned grid = (unsigned)((N + 255) / 256);
    if (packed) {
        hm_forest_sortkey_packed<<<grid, 256, 0, st>>>(ka, codes, key, hist);
        hm_forest_scan<<<scan_grid, 1024, 0, st>>>(hist, (unsigned)bins, scan_sync);
        hm_forest_scatter_codes<<<grid, 256, 0, st>>>(key, hist, codes, N, perm, reinterpret_cast<unsigned long long*>(buf));
        *codes_out = reinterpret_cast<unsigned long long*>(buf);
    } else {
        hm_forest_sortkey<<<grid, 256, 0, st>>>(ka, key, hist);
        hm_forest_scan<<<scan_grid, 1024, 0, st>>>(hist, (unsigned)bins, scan_sync);
        hm_forest_scatter_rows<<<grid, 256, 0, st>>>(key, hist, Xcm, ldx, N, D, Dp, perm, reinterpret_cast<float*>(buf));
        *rows_out = reinterpret_cast<float*>(buf);
        *Dp_out = Dp;
    }
    {
        const cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) {
            cudaFreeAsync(buf, st);
            HM_CUDA_TRY(le);
        }
    }
    *perm_out = perm;
    *scratch_out = buf;
    return 0;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device) instead of on every call
// (the hill climb issues thousands of small launches)
template <bool AP, int NV, int B, int PK>
static cudaError_t hm_forest_kernel_attr() {
    static unsigned done_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 32 && (done_mask >> dev & 1u)) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(hm_forest_kernel<AP, NV, B, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               HM_SMEM_TOTAL);
    if (e == cudaSuccess && dev < 32) done_mask |= 1u << dev;
    return e;
}

static int hm_rf_eval_impl(const hm_forest_t* const* forests, int M, const float* Xcm, int64_t ldx,
                           const unsigned long long* codes, int64_t N, int32_t* const* leaf_out, double* mean_out,
                           double* var_out, double* pred_out, const hm_acq_params_t* acq, const double* feas,
                           double* score_out, void* stream, const HmScoreFilter* filter = nullptr) {
    const bool packed = codes != nullptr;
    HM_REQUIRE(forests != nullptr && M >= 1 && M <= HM_MAX_OBJECTIVES, "1 <= M <= 8");
    HM_REQUIRE(N >= 0 && (packed || ldx >= N), "ldx >= N >= 0");
    if (N == 0) return 0;
    HM_REQUIRE(packed || Xcm != nullptr, "Xcm");
    HmEvalArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M;
    a.D = forests[0] ? forests[0]->n_features : 0;
    int steps = 0;
    bool any_leaf = false, wide = false;
    int n_loop = 0;
    for (int m = 0; m < M; ++m) {
        HM_REQUIRE(forests[m] != nullptr, "forest handle");
        HM_REQUIRE(forests[m]->n_features == a.D, "all forests must share n_features");
        HM_REQUIRE(!packed || forests[m]->packed, "hm_forest_attach_packing has not been called on this forest");
        if (packed && forests[m]->packed == 2) wide = true;
        if (packed && forests[m]->packed == 3) ++n_loop;
        a.f[m] = packed ? forests[m]->devp : forests[m]->dev;
        if (a.f[m].staged) steps += a.f[m].n_groups;
        a.leaf_out[m] = leaf_out ? leaf_out[m] : nullptr;
        any_leaf = any_leaf || (a.leaf_out[m] != nullptr);
    }
    a.nodebuf = packed ? forests[0]->nodebuf_p : forests[0]->nodebuf;
    for (int m = 1; packed && m < M; ++m) a.nodebuf = std::max(a.nodebuf, forests[m]->nodebuf_p);
    a.steps_per_tile = steps;
    a.X = Xcm;
    a.codes = codes;
    a.k.k26 = 1u << 26;
    a.k.k1 = 1u;
    a.k.k8 = 8u;
    a.k.nodebuf = a.nodebuf;
    a.k.xtile_bytes = packed ? 0 : forests[0]->block * a.D * 4;
    HM_REQUIRE(n_loop == 0 || n_loop == M, "forests of one call must be packed for the same space");
    a.k.leaf_and = n_loop ? 32u : 0xffffffffu;
    a.k.leaf_eq = n_loop ? 32u : 0u;
    a.k.loop_shift = n_loop ? forests[0]->loop_shift : 0u;
    a.k.loop_bound = n_loop ? forests[0]->loop_bound : 0u;
    for (int m = 1; m < n_loop; ++m)
        HM_REQUIRE(forests[m]->loop_shift == a.k.loop_shift && forests[m]->loop_bound == a.k.loop_bound,
                   "forests of one call must be packed for the same space");
    a.ldx = ldx;
    a.N = N;
    a.mean_out = mean_out;
    a.var_out = var_out;
    a.pred_out = pred_out;
    a.score_out = score_out;
    a.feas = feas;
    a.has_acq = 0;
    if (filter) {
        a.filter_tau = filter->tau;
        a.list_val = filter->list_val;
        a.list_idx = filter->list_idx;
        a.list_count = filter->list_count;
        a.list_cap = filter->list_cap;
        a.filter_n0 = filter->n0;
    }
    if (acq) {
        HM_REQUIRE(score_out != nullptr || filter != nullptr, "score_out required with acq");
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
    const int block = packed ? HM_PACKED_BLOCK : forests[0]->block;
    cudaStream_t st = (cudaStream_t)stream;
    // coherence pre-pass: pays when the forest streams through shared memory (not resident) and the pool is large
    void* scratch = nullptr;
    {
        const int mode = forests[0]->sort_mode;
        const bool big = steps > 2 && N >= HM_SORT_MIN_N && N < 4294967295ll;
        if (!any_leaf && N < 4294967295ll && (mode == 1 || (mode == -1 && big))) {
            unsigned* perm = nullptr;
            float* rows = nullptr;
            unsigned long long* sorted = nullptr;
            int rc = hm_forest_build_perm(forests, M, Xcm, ldx, codes, N, st, &perm, &rows, &a.Dp, &sorted, &scratch);
            if (rc) return rc;
            a.perm = perm;
            if (packed) a.codes = sorted;
            else a.Xrows = rows;
        }
    }
    const size_t smem = HM_SMEM_HEADER + 2 * (size_t)a.nodebuf + (packed ? 0 : (size_t)block * a.D * 4);
    const long long n_tiles = (N + block - 1) / block;
    const int grid = (int)std::min<long long>(n_tiles, (long long)hm_num_sms() * (packed ? HM_PACKED_CTAS : 1));
    // the plain-predict column of the leaf records is only read for pred_out and Thompson sampling
#ifndef HM_FORCE_V
#define HM_FORCE_V 0
#endif
    const bool need_v = HM_FORCE_V || pred_out != nullptr || (acq && a.acq.kind == HM_ACQ_TS);
    // forests resident in shared memory whose groups carry their leaf records: no leaf gather from global memory at all
    bool leaf_smem = steps <= 2;
    for (int m = 0; m < M; ++m)
        leaf_smem = leaf_smem && a.f[m].staged && (packed ? forests[m]->leaf_smem_p : forests[m]->leaf_smem);
    cudaError_t attr_err = cudaSuccess;
#define HM_LAUNCH_B(AP, NV, B, PK)                                                                                     \
    do {                                                                                                               \
        attr_err = hm_forest_kernel_attr<AP, NV, B, PK>();                                                             \
        if (attr_err == cudaSuccess) hm_forest_kernel<AP, NV, B, PK><<<grid, B, smem, st>>>(a);                        \
    } while (0)
#define HM_LAUNCH(AP, NV)                                                                                              \
    do {                                                                                                               \
        if (packed && n_loop) HM_LAUNCH_B(AP, NV, HM_PACKED_BLOCK, 3);                                                 \
        else if (packed && !wide) HM_LAUNCH_B(AP, NV, HM_PACKED_BLOCK, 1);                                             \
        else if (packed) HM_LAUNCH_B(AP, NV, HM_PACKED_BLOCK, 2);                                                      \
        else if (block == 1024) HM_LAUNCH_B(AP, NV, 1024, 0);                                                          \
        else if (block == 512) HM_LAUNCH_B(AP, NV, 512, 0);                                                            \
        else if (block == 256) HM_LAUNCH_B(AP, NV, 256, 0);                                                            \
        else HM_LAUNCH_B(AP, NV, 128, 0);                                                                              \
    } while (0)
    if (any_leaf) {
        if (need_v) HM_LAUNCH(true, 1); else HM_LAUNCH(true, 0);
    } else {
        if (need_v) HM_LAUNCH(false, 1); else if (leaf_smem) HM_LAUNCH(false, 2); else HM_LAUNCH(false, 0);
    }
#undef HM_LAUNCH
#undef HM_LAUNCH_B
    const cudaError_t launch_err = attr_err != cudaSuccess ? attr_err : cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, st);          // also on a failed launch: the scratch must not leak
    HM_CUDA_TRY(launch_err);
    return 0;
}

extern "C" int hm_rf_eval(const hm_forest_t* const* forests, int M, const float* Xcm, int64_t ldx, int64_t N,
                          int32_t* const* leaf_out, double* mean_out, double* var_out, double* pred_out,
                          const hm_acq_params_t* acq, const double* feas, double* score_out, void* stream) {
    return hm_rf_eval_impl(forests, M, Xcm, ldx, nullptr, N, leaf_out, mean_out, var_out, pred_out, acq, feas, score_out,
                           stream);
}

extern "C" int hm_rf_eval_packed(const hm_forest_t* const* forests, int M, const uint64_t* codes, int64_t N,
                                 int32_t* const* leaf_out, double* mean_out, double* var_out, double* pred_out,
                                 const hm_acq_params_t* acq, const double* feas, double* score_out, void* stream) {
    HM_REQUIRE(N == 0 || codes != nullptr, "codes");
    return hm_rf_eval_impl(forests, M, nullptr, 0, reinterpret_cast<const unsigned long long*>(codes), N, leaf_out,
                           mean_out, var_out, pred_out, acq, feas, score_out, stream);
}

// ---- fused scoring + selection -------------------------------------------------------------------------------------------
// hm_rf_score_topk: the k best candidates of a pool WITHOUT the score array ever existing in HBM (local_search only needs
// the seeds of its random pools, local_search.py:625-645).  Two passes, both through the forest kernel:
//   1. the first HM_ST_PREFIX candidates are scored the plain way and hm_topk gives tau = their k-th best value;
//   2. the whole pool is scored with the kernel's epilogue in filter mode: a candidate is appended to a short list (one
//      atomic) only if it can still be among the k best -- score < tau, or == tau with an index inside the prefix (k
//      candidates with value <= tau and a smaller index exist already, so a later tie can never be selected: tie classes,
//      the norm on discrete spaces, do not flood the list), or NaN (np.argmin puts NaN first);
//   3. hm_topk_merge on the list (value, global index) = the stable top-k of the whole pool.
// Expected list length ~ k * N / HM_ST_PREFIX; if the list overflows (scores that keep improving along the pool) the call
// falls back to the plain score array + hm_topk.  Results are identical to hm_rf_eval + hm_topk by construction.
#define HM_ST_PREFIX (1 << 16)
#define HM_ST_CAP (1 << 18)
struct HmStWs {
    void* topk_ws;
    double* prefix_scores;   // [HM_ST_PREFIX]
    double* tmp_vals;        // [1024]
    long long* tmp_idx;      // [1024]
    unsigned* count;
    double* list_val;        // [HM_ST_CAP]
    long long* list_idx;     // [HM_ST_CAP]
};
static size_t hm_st_layout(int k, unsigned char* base, HmStWs* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    unsigned char* a = take((size_t)hm_topk_workspace_bytes(k));
    unsigned char* b = take(sizeof(double) * HM_ST_PREFIX);
    unsigned char* c = take(sizeof(double) * 1024);
    unsigned char* d = take(sizeof(long long) * 1024);
    unsigned char* e = take(256);
    unsigned char* f = take(sizeof(double) * HM_ST_CAP);
    unsigned char* g = take(sizeof(long long) * HM_ST_CAP);
    if (ws) {
        ws->topk_ws = a;
        ws->prefix_scores = (double*)b;
        ws->tmp_vals = (double*)c;
        ws->tmp_idx = (long long*)d;
        ws->count = (unsigned*)e;
        ws->list_val = (double*)f;
        ws->list_idx = (long long*)g;
    }
    return off;
}
extern "C" int64_t hm_rf_score_topk_workspace_bytes(int k) {
    if (k < 1) k = 1;
    return (int64_t)hm_st_layout(k, nullptr, nullptr) + 256;
}

__global__ void hm_st_fill(double* __restrict__ val, long long* __restrict__ idx, unsigned n, unsigned* __restrict__ count) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = 0;
    if (i < n) {
        val[i] = __longlong_as_double(0x7ff0000000000000ll);        // +inf with the largest index: behind every real entry
        idx[i] = 0x7fffffffffffffffll;
    }
}

static int hm_rf_score_topk_impl(const hm_forest_t* const* forests, int M, const float* Xcm, int64_t ldx,
                                 const unsigned long long* codes, int64_t N, const hm_acq_params_t* acq, const double* feas,
                                 int k, double* out_vals, int64_t* out_idx, void* workspace, void* stream) {
    HM_REQUIRE(acq != nullptr, "acq");
    HM_REQUIRE(k >= 1 && k <= 1024, "1 <= k <= 1024");
    HM_REQUIRE(out_vals && out_idx && workspace, "outputs/workspace");
    HM_REQUIRE(N >= 0, "N >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    HmStWs ws;
    unsigned char* base = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    hm_st_layout(k, base, &ws);
    if (N <= HM_ST_CAP) {
        // short pool: the list area holds the whole score array
        int rc = hm_rf_eval_impl(forests, M, Xcm, ldx, codes, N, nullptr, nullptr, nullptr, nullptr, acq, feas, ws.list_val,
                                 stream);
        if (rc) return rc;
        return hm_topk(ws.list_val, N, 0, k, out_vals, out_idx, ws.topk_ws, stream);
    }
    const int64_t n0 = HM_ST_PREFIX;
    int rc = hm_rf_eval_impl(forests, M, Xcm, ldx, codes, n0, nullptr, nullptr, nullptr, nullptr, acq, feas,
                             ws.prefix_scores, stream);
    if (rc) return rc;
    rc = hm_topk(ws.prefix_scores, n0, 0, k, ws.tmp_vals, reinterpret_cast<int64_t*>(ws.tmp_idx), ws.topk_ws, stream);
    if (rc) return rc;
    hm_st_fill<<<HM_ST_CAP / 256, 256, 0, st>>>(ws.list_val, ws.list_idx, HM_ST_CAP, ws.count);
    HM_CUDA_TRY(cudaGetLastError());
    HmScoreFilter filt = {ws.tmp_vals + (k - 1), ws.list_val, ws.list_idx, ws.count, HM_ST_CAP, (long long)n0};
    rc = hm_rf_eval_impl(forests, M, Xcm, ldx, codes, N, nullptr, nullptr, nullptr, nullptr, acq, feas, nullptr, stream, &filt);
    if (rc) return rc;
    rc = hm_topk_merge(ws.list_val, reinterpret_cast<const int64_t*>(ws.list_idx), HM_ST_CAP, k, out_vals, out_idx,
                       ws.topk_ws, stream);
    if (rc) return rc;
    unsigned pushed = 0;
    HM_CUDA_TRY(cudaMemcpyAsync(&pushed, ws.count, sizeof(pushed), cudaMemcpyDeviceToHost, st));
    HM_CUDA_TRY(cudaStreamSynchronize(st));
    if (pushed > HM_ST_CAP) {
        // the list overflowed: plain score array + hm_topk (correct for any input, twice the time)
        double* scores = nullptr;
        HM_CUDA_TRY(cudaMallocAsync((void**)&scores, sizeof(double) * (size_t)N, st));
        rc = hm_rf_eval_impl(forests, M, Xcm, ldx, codes, N, nullptr, nullptr, nullptr, nullptr, acq, feas, scores, stream);
        if (!rc) rc = hm_topk(scores, N, 0, k, out_vals, out_idx, ws.topk_ws, stream);
        cudaFreeAsync(scores, st);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int hm_rf_score_topk(const hm_forest_t* const* forests, int M, const float* Xcm, int64_t ldx, int64_t N,
                                const hm_acq_params_t* acq, const double* feas, int k, double* out_vals, int64_t* out_idx,
                                void* workspace, void* stream) {
    HM_REQUIRE(N == 0 || Xcm != nullptr, "Xcm");
    return hm_rf_score_topk_impl(forests, M, Xcm, ldx, nullptr, N, acq, feas, k, out_vals, out_idx, workspace, stream);
}

extern "C" int hm_rf_score_topk_packed(const hm_forest_t* const* forests, int M, const uint64_t* codes, int64_t N,
                                       const hm_acq_params_t* acq, const double* feas, int k, double* out_vals,
                                       int64_t* out_idx, void* workspace, void* stream) {
    HM_REQUIRE(N == 0 || codes != nullptr, "codes");
    return hm_rf_score_topk_impl(forests, M, nullptr, 0, reinterpret_cast<const unsigned long long*>(codes), N, acq, feas, k,
                                 out_vals, out_idx, workspace, stream);
}

HM_CHECKED_EXPORT(forest)
