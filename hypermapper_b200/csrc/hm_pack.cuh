// Packed candidates for all-discrete search spaces (ordinal / categorical parameters only).
//
// A candidate of such a space is a handful of small integers: the rank of every ordinal value inside its ascending value
// list and the index of every category.  They are packed into ONE 64-bit code per candidate; the forest kernel keeps the
// code in a register pair and tests a split with a funnel shift and an unsigned compare -- no shared-memory feature tile,
// no per-visit feature load (38 % of the L1/shared data-pipe traffic of the fp32 path), 8 bytes per candidate over PCIe
// and HBM instead of 4 * D_enc.
//
// Layout (a function of the parameter list alone, so hm_space.cu and hm_forest.cu derive the same one), fields allocated
// from bit 63 DOWNWARDS: first every categorical in input-parameter order, then every ordinal in input-parameter order
// except the (first) widest ordinal, which goes last.
//   categorical, n values  : n - 1 bits, one-hot over the first n - 1 categories (category c < n - 1 sets bit low + c);
//                            the last category is the all-zero field
//   ordinal, n values      : max(1, ceil(log2 n)) bits holding the value's rank (0 = smallest value)
// Split test.  sklearn goes RIGHT iff (double)x32 > threshold (Tree._apply_dense, sklearn/tree/_tree.pyx).  The encoded
// value of an ordinal is monotone in its rank (models.py:962-965: index/size), so with
//        r1 = #{ranks whose float32 encoded value is <= threshold}
// "right" <=> rank >= r1.  A one-hot column is 1.0 for one category, so for 0 <= threshold < 1 "right" <=> the category's
// bit is set -- or, for the last category, <=> the whole field is zero.  All three become ONE test,
//        go to the SECOND child slot  <=>  hi32(code << s) >= T,
// s = 63 - (top bit of the field / the category's bit), T = r1 << (32 - width) (ordinal), 0x80000000 (a category's bit),
// 1 << (32 - width) with the two children stored in swapped order (last category: a non-zero field is sklearn's LEFT).
// The shift moves the field to the top of the word; the bits of lower fields that follow it add less than one unit of
// the field.  Splits that send every possible value the same way (r1 == 0 or r1 == n; threshold outside [0, 1) on a
// one-hot column; a one-category categorical) are bypassed when the packed node array is built.
// Self-looping leaves (FAST layouts with a "sentinel" field).  A field whose shifted value has a provable upper bound B
// below 2^32 -- a categorical with >= 3 categories (one-hot: its two top bits are never both set, B = 0xC0000000 or less),
// an ordinal whose value count is not a power of two (rank <= n - 1, B = n << (32 - width)) -- gives a test that is FALSE
// for every candidate of the space:  hi32(code << s_field) >= B + x  for any x >= 0.  A leaf is then stored as
//        T = B + leaf ordinal,   word 1 = (the leaf's own byte offset) << 6 | 32 | s_field
// and "walks" to its own first slot for ever: the traversal loop needs no per-visit leaf predicate (5 instructions per
// visit instead of 6).  Bit 5 of word 1 marks a leaf (FAST shifts are <= 31; the funnel shift reads bits 0-4 only, the
// slot address bits 6 and up).  hm_pack_sentinel() picks the field; the kernel forces a candidate whose sentinel field
// breaks the bound (a malformed code handed in by a caller) to the all-zero code, so the loop always terminates inside
// the staged nodes.
// FAST layout: when every tested bit / field top lies in the upper word (s <= 31) -- the widest ordinal may hang down
// into the lower word -- the shift is one 32-bit funnel shift of the register pair (SHF.L.W.U32.HI); otherwise the
// kernel uses the 64-bit shift (one more instruction per visit).
#pragma once
#include <vector>
#include "hm_common.cuh"

struct HmPackField {
    int kind;       // HM_P_ORDINAL / HM_P_CATEGORICAL
    int n_values;
    int top;        // bit index (63..0) of the field's most significant bit
    int low;        // bit index of the field's least significant bit
    int width;
    int out_col;    // first encoded column of the parameter
};

// the search-space handle (hm_space.cu); hm_forest.cu reads its host-side copies to build the packed node arrays
struct hm_space {
    int D, D_enc;
    long long n_tables;
    hm_param_desc_t* d_params;
    double* d_tables;
    std::vector<hm_param_desc_t> params;
    std::vector<double> tables;          // host copy of the table blob
    std::vector<HmPackField> pack;       // packed-code layout (empty when the space cannot be packed)
    int pack_bits;                       // 0 when the space cannot be packed
    int pack_fast;                       // every split test needs a shift <= 31 (hm_pack.cuh, FAST layout)
    int* d_pack_low;                     // device [D]: low bit of every parameter's field
};

static inline int hm_pack_width(int kind, int n_values) {
    if (kind == HM_P_CATEGORICAL) return n_values - 1;
    int w = 1;
    while ((1 << w) < n_values) ++w;
    return w;
}

// fills out[n]; returns the number of bits used (>= 1), or 0 when the space cannot be packed (a real / integer parameter,
// or more than 64 bits).  *fast (nullable) = 1 when every split test needs a shift of at most 31.
static inline int hm_pack_layout(const hm_param_desc_t* p, int n, HmPackField* out, int* fast) {
    int widest = -1;
    for (int j = 0; j < n; ++j) {
        if (p[j].kind != HM_P_ORDINAL && p[j].kind != HM_P_CATEGORICAL) return 0;
        if (p[j].n_values < 1) return 0;
        if (hm_pack_width(p[j].kind, p[j].n_values) > 26) return 0;
        if (p[j].kind == HM_P_ORDINAL &&
            (widest < 0 || hm_pack_width(p[j].kind, p[j].n_values) > hm_pack_width(p[widest].kind, p[widest].n_values)))
            widest = j;
    }
    int next = 63;
    bool is_fast = true;
    for (int pass = 0; pass < 3; ++pass) {
        for (int j = 0; j < n; ++j) {
            const bool cat = p[j].kind == HM_P_CATEGORICAL;
            if (pass == 0 ? !cat : (pass == 1 ? (cat || j == widest) : j != widest)) continue;
            const int w = hm_pack_width(p[j].kind, p[j].n_values);
            if (next - w + 1 < 0) return 0;
            out[j].kind = p[j].kind;
            out[j].n_values = p[j].n_values;
            out[j].top = next;
            out[j].low = w ? next - w + 1 : 0;      // a one-category categorical takes no bits
            out[j].width = w;
            out[j].out_col = p[j].out_col;
            if (w > 0 && (cat ? out[j].low < 32 : out[j].top < 32)) is_fast = false;
            next -= w;
        }
    }
    if (fast) *fast = is_fast ? 1 : 0;
    return 63 - next > 0 ? 63 - next : 1;
}

// the sentinel field of a FAST layout (see above): shift and exclusive upper bound of hi32(code << shift); false if none
static inline bool hm_pack_sentinel(const HmPackField* f, int n, unsigned* shift, unsigned* bound) {
    for (int j = 0; j < n; ++j) {
        const int w = f[j].width;
        if (w < 1 || f[j].top < 32) continue;
        unsigned long long b = 0;
        if (f[j].kind == HM_P_CATEGORICAL) {
            if (w >= 2) b = ((1ull << (w - 1)) + 1ull) << (32 - w);      // at most one bit set: never the top two
        } else if (f[j].n_values < (1 << w)) {
            b = (unsigned long long)f[j].n_values << (32 - w);           // rank <= n - 1, lower fields add < 1 unit
        }
        if (b && b <= 0xE0000000ull) {
            *shift = (unsigned)(63 - f[j].top);
            *bound = (unsigned)b;
            return true;
        }
    }
    return false;
}
