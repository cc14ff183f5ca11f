// Shared between match.cu (orchestration, SIMT path) and match_tc.cu (tcgen05 path).
#pragma once
#include <cuda_bf16.h>

#include "mp_common.cuh"

namespace mp {

// best / second-best / third-best similarity key of one descriptor over the other set (larger =
// closer).  The third key has no index: it only tells whether a near-tie is confined to the two
// indexed candidates (then two exact fp64 keys settle it) or needs a full exact rescan.
struct Top2 {
    float best, second, third;
    int best_idx, second_idx;
};

__device__ __forceinline__ Top2 top2_empty() {
    Top2 t;
    t.best = -INFINITY; t.second = -INFINITY; t.third = -INFINITY; t.best_idx = -1; t.second_idx = -1;
    return t;
}

// ascending-index visits + strict '>' => equal keys resolve to the lowest index
__device__ __forceinline__ void top2_update(Top2 &t, float key, int idx) {
    if (key > t.best) {
        t.third = t.second;
        t.second = t.best; t.second_idx = t.best_idx;
        t.best = key; t.best_idx = idx;
    } else if (key > t.second) {
        t.third = t.second;
        t.second = key; t.second_idx = idx;
    } else if (key > t.third) {
        t.third = key;
    }
}

// Bound on |approximate key - exact key| relative to max|a| * max|b| (derivation in DESIGN.md):
//  tensor: split bf16 (hi*hi + hi*mid + mid*hi) drops <= 3*2^-18 of sum|a_k b_k| (1.2e-5), fp32
//          accumulation in the tensor core over <= 48 MMAs adds <= ~1e-5  -> 4e-5 with margin
//  simt  : fp32 FMA chain over D <= 4096 terms                              -> 4e-5 as well
constexpr float MATCH_EPS_TENSOR = 4e-5f;
// the tensor epilogue orders keys by the mantissa of f = key + C, C = 3 * 2^E < 6.1 * bound (KeyScale below): the add
// rounds to ulp(f) / 2 <= 2^-23 * 2^(E+1) < 2^-21 * bound per key.  The flag kernel adds pack_rel * 2 * bound.
constexpr float MATCH_PACK_REL = 4.76837e-7f;  // 2^-21
constexpr float MATCH_EPS_SIMT = 4e-5f;

// Packed 32-bit keys of the tensor epilogue.  With C = 3 * 2^E and 2^E > bound >= |v|, f = v + C lies in the single
// binade [2^(E+1), 2^(E+2)), so bits(f) * 32 + code keeps the whole mantissa above a 5-bit code and compares like f.
struct KeyScale {
    float C;           // 3 * 2^E
    uint32_t expbits;  // exponent field of that binade, in place
};
__device__ __forceinline__ KeyScale key_scale(float bound) {
    const float b = fmaxf(1.01f * bound, 1e-30f);
    const uint32_t e = (__float_as_uint(b) >> 23) + 1u;   // 2^(e-127) > b
    KeyScale k;
    k.C = 3.f * __uint_as_float(e << 23);
    k.expbits = (e + 1u) << 23;
    return k;
}
// the value v a key stands for (low 5 bits = code)
__device__ __forceinline__ float key_value(const KeyScale &k, uint32_t key) {
    return __uint_as_float(k.expbits | ((key >> 5) & 0x7fffffu)) - k.C;
}

struct MatchLayout {
    size_t scalars, row_part_key, row_part_idx, norms1, norms2, top12, top21, idx12, idx21, flagged1, flagged2, pairs1, pairs2, train_tmp, dist_tmp, hi1, mid1, hi2, mid2, colpart, total;
    MatchLayout(int P, int N1, int N2, int D);
};

// match_tc.cu: rows of A (hi/mid bf16 planes, (P,NA,D)) against rows of B; writes top[(P,NA)] and, when top_cols is
// given, the column side top_cols[(P,NB)] (rows of B against rows of A) from the same GEMM.  colpart: workspace of
// match_colpart_bytes(P, NA, NB) for the per-chunk column records.
size_t match_colpart_bytes(int P, int NA, int NB);
int match_top2_tensor(const __nv_bfloat16 *a_hi, const __nv_bfloat16 *a_mid, const int32_t *na, int NA,
                      const __nv_bfloat16 *b_hi, const __nv_bfloat16 *b_mid, const int32_t *nb, int NB, int P, int D,
                      const float *norms_a, const float *norms_b, int use_bias, const unsigned *max_a, const unsigned *max_b,
                      Top2 *top, Top2 *top_cols, void *colpart, cudaStream_t stream);

}  // namespace mp
