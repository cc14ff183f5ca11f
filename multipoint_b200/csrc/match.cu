// Rows 6-8 of the hot path: utils.get_matches (multipoint/utils/matching.py:4-99):
// cv2.BFMatcher(NORM_L2, crossCheck) (:7,31), NNMatcher (:35-72), the knn ratio test (:21-28)
// and ThresholdMatcher (:74-99).  The reference copies both descriptor sets to the host and runs
// OpenCV / numpy there; here the whole matcher stays on the device.
//
// Pipeline (per call, P independent image pairs):
//   prep      row norms (L2 metric) and the largest norm per set (error bound), split of the
//             fp32 descriptors into bf16 hi/mid planes for the tensor path
//   top-2     for every row the best and second-best similarity key over the other set:
//             MP_ALGO_TENSOR -> match_tc.cu (tcgen05 split-bf16 GEMM, fused arg-top-2 epilogue)
//             MP_ALGO_SIMT   -> fp32 CUDA-core tiles below (device-side reference)
//             key = a.b (NN metric) or a.b - |b|^2/2 (L2 metric: argmin |a-b|^2 over b)
//   flag      rows whose top-2 margin is below twice the error bound of the approximate key
//   recheck   flagged rows recomputed over the whole other set in fp64 -> the index is the exact
//             argmin with ties to the lowest index, as np.argmin / OpenCV's strict '<' scan give
//   select    mutual / threshold / ratio test and ordered compaction into (query, train, dist);
//             dist is recomputed per kept pair (fp64 accumulate, rounded once to fp32) in the
//             reference's own formulation.
#include <math.h>

#include "match_internal.cuh"

namespace mp {

// ------------------------------------------------------------------ prep
// One warp per descriptor row: squared norm, per-(pair,side) max norm, optional bf16 split.
__global__ void __launch_bounds__(256)
match_prep_kernel(const float *__restrict__ d, const int32_t *__restrict__ counts, int N, int D, int P,
                  float *__restrict__ norms, unsigned *__restrict__ max_norm_bits,
                  __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ mid) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)P * N) return;
    const int p = (int)(row / N), i = (int)(row - (long long)p * N);
    const bool valid = counts == nullptr || i < counts[p];
    const float *src = d + (size_t)row * D;
    float ss = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float v = valid ? src[c] : 0.f;
        ss += v * v;
        if (hi != nullptr) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const __nv_bfloat16 m = __float2bfloat16_rn(v - __bfloat162float(h));
            hi[(size_t)row * D + c] = h;
            mid[(size_t)row * D + c] = m;
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    if (lane == 0) {
        norms[row] = ss;
        if (valid) atomicMax(max_norm_bits + p, __float_as_uint(sqrtf(ss)));
    }
}

// Vector form for D % 8 == 0: every lane converts 8 consecutive elements (two 128-bit loads, one
// 128-bit store per bf16 plane); 4 / 2 / 1 rows per warp for D <= 64 / 128 / larger.
template <int RPW>
__global__ void __launch_bounds__(256)
match_prep_vec_kernel(const float *__restrict__ d, const int32_t *__restrict__ counts, int N, int D, int P,
                      float *__restrict__ norms, unsigned *__restrict__ max_norm_bits,
                      __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ mid) {
    constexpr int LPR = 32 / RPW;  // lanes per row
    const int lane = threadIdx.x & 31, sub = lane % LPR;
    const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / LPR;
    const bool in_range = row < (long long)P * N;
    const int p = in_range ? (int)(row / N) : 0, i = in_range ? (int)(row - (long long)p * N) : 0;
    const bool valid = in_range && (counts == nullptr || i < counts[p]);
    float ss = 0.f;
    if (in_range) {
        for (int c0 = 8 * sub; c0 < D; c0 += 8 * LPR) {
            float v[8];
            if (valid) {
                const float4 q0 = __ldg(reinterpret_cast<const float4 *>(d + (size_t)row * D + c0));
                const float4 q1 = __ldg(reinterpret_cast<const float4 *>(d + (size_t)row * D + c0 + 4));
                v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.f;
            }
            if (hi != nullptr) {
                __align__(16) __nv_bfloat16 h[8], m[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    h[e] = __float2bfloat16_rn(v[e]);
                    m[e] = __float2bfloat16_rn(v[e] - __bfloat162float(h[e]));
                }
                *reinterpret_cast<uint4 *>(hi + (size_t)row * D + c0) = *reinterpret_cast<const uint4 *>(h);
                *reinterpret_cast<uint4 *>(mid + (size_t)row * D + c0) = *reinterpret_cast<const uint4 *>(m);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
        }
    }
#pragma unroll
    for (int s = LPR / 2; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    if (in_range && sub == 0) norms[row] = ss;
    // largest norm per pair: one atomic per (block, pair) instead of one per row -- 2048 rows of a
    // pair hitting one address serialised in L2 and cost more than the conversion itself
    __shared__ unsigned blk_max[2];
    __shared__ int blk_pair0;
    if (threadIdx.x == 0) {
        blk_max[0] = blk_max[1] = 0;
        const long long first = (long long)blockIdx.x * 8 * RPW;
        blk_pair0 = (int)(first / N);
    }
    __syncthreads();
    if (valid && sub == 0) {
        const int slot = p - blk_pair0;  // a block of <= 32 rows spans at most two pairs when N >= 32
        const unsigned bits = __float_as_uint(sqrtf(ss));
        if (slot == 0 || slot == 1) atomicMax(&blk_max[slot], bits);
        else atomicMax(max_norm_bits + p, bits);
    }
    __syncthreads();
    if (threadIdx.x < 2 && blk_max[threadIdx.x] != 0 && blk_pair0 + (int)threadIdx.x < P)
        atomicMax(max_norm_bits + blk_pair0 + threadIdx.x, blk_max[threadIdx.x]);
}

// ------------------------------------------------------------------ SIMT top-2
// 128 rows per CTA (one per thread), 32 columns of the other set per step, K chunks of 16
// through shared memory.  Columns are visited in ascending order with strict '>' updates, so
// equal keys resolve to the lowest index.
constexpr int ST_ROWS = 128, ST_COLS = 32, ST_K = 16;

__global__ void __launch_bounds__(ST_ROWS)
match_top2_simt_kernel(const float *__restrict__ a, const int32_t *__restrict__ na, int NA,
                       const float *__restrict__ b, const int32_t *__restrict__ nb, int NB, int D,
                       const float *__restrict__ norms_b, int use_bias, Top2 *__restrict__ top) {
    __shared__ float As[ST_ROWS][ST_K + 1];
    __shared__ float Bs[ST_COLS][ST_K + 1];
    const int p = blockIdx.y, tid = threadIdx.x;
    const int row0 = blockIdx.x * ST_ROWS, row = row0 + tid;
    const int n_a = na ? min(na[p], NA) : NA, n_b = nb ? min(nb[p], NB) : NB;
    if (row0 >= n_a) {  // whole CTA beyond this pair's valid rows: empty results
        if (row < NA) {
            top[(size_t)p * NA + row] = top2_empty();
        }
        return;
    }
    const float *ap = a + (size_t)p * NA * D;
    const float *bp = b + (size_t)p * NB * D;
    Top2 t = top2_empty();
    for (int c0 = 0; c0 < n_b; c0 += ST_COLS) {
        float acc[ST_COLS];
#pragma unroll
        for (int j = 0; j < ST_COLS; ++j) acc[j] = 0.f;
        for (int k0 = 0; k0 < D; k0 += ST_K) {
            __syncthreads();
            for (int e = tid; e < ST_ROWS * ST_K; e += ST_ROWS) {
                const int r = e / ST_K, k = e - r * ST_K;
                As[r][k] = (row0 + r < n_a && k0 + k < D) ? ap[(size_t)(row0 + r) * D + k0 + k] : 0.f;
            }
            for (int e = tid; e < ST_COLS * ST_K; e += ST_ROWS) {
                const int c = e / ST_K, k = e - c * ST_K;
                Bs[c][k] = (c0 + c < n_b && k0 + k < D) ? bp[(size_t)(c0 + c) * D + k0 + k] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < ST_K; ++k) {
                const float av = As[tid][k];
#pragma unroll
                for (int j = 0; j < ST_COLS; ++j) acc[j] = fmaf(av, Bs[j][k], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < ST_COLS; ++j) {
            const int col = c0 + j;
            if (col >= n_b) break;
            float key = acc[j];
            if (use_bias) key -= 0.5f * norms_b[(size_t)p * NB + col];
            top2_update(t, key, col);
        }
    }
    if (row < NA) {
        if (row >= n_a) t = top2_empty();
        top[(size_t)p * NA + row] = t;
    }
}

// ------------------------------------------------------------------ flag near-ties
// Rows whose top-2 margin is inside the error bound, per (pair, side) group = 2*pair + side:
//   pair list : the third key is clear of the bound -> only the two indexed candidates can be
//               the true nearest neighbour: two exact keys settle it (match_recheck_pair_kernel)
//   full list : three or more candidates within the bound (or the NN metric's clip plateau) ->
//               exact rescan of the whole other set (match_recheck_kernel)
struct FlagSide {
    const Top2 *top;
    int N;
    const unsigned *max_a, *max_b;
    int32_t *idx_out, *flagged, *pairs;
};
// grid (ceil(max(N1, N2) * P / 256), 2): blockIdx.y = side (0: rows of set 1 against set 2, 1: the other way round)
__global__ void match_flag_kernel(const FlagSide side0, const FlagSide side1, int P, int metric, float eps_rel, float pack_rel,
                                  int *__restrict__ n_flagged, int *__restrict__ n_pairs) {
    const int side = blockIdx.y;
    const FlagSide &S = side == 0 ? side0 : side1;
    const int N = S.N;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)P * N) return;
    const int p = (int)(g / N);
    const Top2 t = S.top[g];
    S.idx_out[g] = t.best_idx;
    if (t.best_idx < 0 || t.second_idx < 0) return;
    const float ma = __uint_as_float(S.max_a[p]), mb = __uint_as_float(S.max_b[p]);
    const float eps = eps_rel * fmaxf(ma * mb, 1e-30f) +
                      pack_rel * 2.f * (1.002f * ma * mb + (metric == MP_METRIC_L2 ? 0.5f * mb * mb : 0.f));
    const bool close2 = !(t.best - t.second >= 2.f * eps);  // also catches NaN keys
    const bool close3 = !(t.best - t.third >= 2.f * eps);
    const bool plateau = metric == MP_METRIC_NN && t.best >= 1.f - eps;  // clip(.,-1,1) can tie many columns
    const int row = (int)(g - (long long)p * N);
    if (plateau || (close2 && close3)) S.flagged[(size_t)p * N + atomicAdd(n_flagged + 2 * p + side, 1)] = row;
    else if (close2) S.pairs[(size_t)p * N + atomicAdd(n_pairs + 2 * p + side, 1)] = row;
}

// exact fp64 key of one (a, b) pair, lanes striding over D; result in every lane
__device__ __forceinline__ double exact_key(const float *a, const float *b, int D, int metric, int lane) {
    double acc = 0.0;
    if (metric == MP_METRIC_NN) {
        for (int c = lane; c < D; c += 32) acc = fma((double)a[c], (double)b[c], acc);
    } else {
        for (int c = lane; c < D; c += 32) { const double df = (double)a[c] - (double)b[c]; acc = fma(df, df, acc); }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    return metric == MP_METRIC_NN ? -fmin(1.0, fmax(-1.0, acc)) : acc;
}

// One warp per pair-flagged row: exact keys of its two candidates, the smaller wins, ties to the
// lower index.  grid (2P groups, RP_Y); 8 warps per CTA.
constexpr int RP_Y = 8;
__global__ void __launch_bounds__(256)
match_recheck_pair_kernel(const float *__restrict__ d1, int N1, const float *__restrict__ d2, int N2, int D, int metric,
                          const Top2 *__restrict__ top12, const Top2 *__restrict__ top21,
                          const int32_t *__restrict__ pairs1, const int32_t *__restrict__ pairs2,
                          const int *__restrict__ n_pairs, int32_t *__restrict__ idx12, int32_t *__restrict__ idx21) {
    const int group = blockIdx.x, p = group >> 1, side = group & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = n_pairs[group];
    const int NA = side == 0 ? N1 : N2, NB = side == 0 ? N2 : N1;
    const float *A = side == 0 ? d1 + (size_t)p * N1 * D : d2 + (size_t)p * N2 * D;
    const float *Bm = side == 0 ? d2 + (size_t)p * N2 * D : d1 + (size_t)p * N1 * D;
    const int32_t *rows = (side == 0 ? pairs1 : pairs2) + (size_t)p * NA;
    const Top2 *top = (side == 0 ? top12 : top21) + (size_t)p * NA;
    int32_t *dst = side == 0 ? idx12 + (size_t)p * N1 : idx21 + (size_t)p * N2;
    (void)NB;
    for (int i = blockIdx.y * 8 + warp; i < count; i += RP_Y * 8) {
        const int row = rows[i];
        const int j1 = top[row].best_idx, j2 = top[row].second_idx;
        const double k1 = exact_key(A + (size_t)row * D, Bm + (size_t)j1 * D, D, metric, lane);
        const double k2 = exact_key(A + (size_t)row * D, Bm + (size_t)j2 * D, D, metric, lane);
        if (lane == 0) dst[row] = (k2 < k1 || (k2 == k1 && j2 < j1)) ? j2 : j1;
    }
}

// ------------------------------------------------------------------ exact recheck
// fp64 key of each flagged row against every row of the other set; first minimum wins.
//   NN: key = -clip(a.b, -1, 1)      L2: key = sum (a-b)^2       (monotone in the distance)
// grid (2P groups, RK_Z): a CTA takes chunks of RK_ROWS flagged rows of its group.  The chunk's
// query rows live in registers as doubles (lane l holds elements l, l+32, ...); each warp walks
// the other set's rows with coalesced 128 B loads and reduces RK_ROWS partial sums per row with a
// halving exchange (18 shuffles instead of 80), so row r's total lands in the lanes with
// ((lane >> 2) & 7) == r.
constexpr int RK_ROWS = 8, RK_Z = 4, RK_WARPS = 8, RK_MAXD = 256;
constexpr int RK_ROW_MAX = 16;  // groups with <= RK_ROW_MAX rows use match_recheck_row_kernel

template <int DPL>  // elements per lane: D <= 32*DPL
__global__ void __launch_bounds__(RK_WARPS * 32)
match_recheck_kernel(const float *__restrict__ d1, const int32_t *__restrict__ n1, int N1,
                     const float *__restrict__ d2, const int32_t *__restrict__ n2, int N2, int D, int metric,
                     const int32_t *__restrict__ flagged1, const int32_t *__restrict__ flagged2,
                     const int *__restrict__ n_flagged, int32_t *__restrict__ idx12, int32_t *__restrict__ idx21) {
    __shared__ double red_key[RK_WARPS][RK_ROWS];
    __shared__ int red_idx[RK_WARPS][RK_ROWS];
    const int group = blockIdx.x, p = group >> 1, side = group & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = n_flagged[group];
    if (count <= RK_ROW_MAX) return;  // handled by match_recheck_row_kernel
    const int NA = side == 0 ? N1 : N2;
    const float *A = side == 0 ? d1 + (size_t)p * N1 * D : d2 + (size_t)p * N2 * D;
    const float *Bm = side == 0 ? d2 + (size_t)p * N2 * D : d1 + (size_t)p * N1 * D;
    const int32_t *rows = (side == 0 ? flagged1 : flagged2) + (size_t)p * NA;
    int32_t *dst = side == 0 ? idx12 + (size_t)p * N1 : idx21 + (size_t)p * N2;
    const int nb = side == 0 ? (n2 ? min(n2[p], N2) : N2) : (n1 ? min(n1[p], N1) : N1);

    for (int c0 = blockIdx.y * RK_ROWS; c0 < count; c0 += RK_Z * RK_ROWS) {
        double a[RK_ROWS][DPL];
#pragma unroll
        for (int r = 0; r < RK_ROWS; ++r) {
            const int row = c0 + r < count ? rows[c0 + r] : -1;
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                const int k = lane + 32 * i;
                a[r][i] = (row >= 0 && k < D) ? (double)A[(size_t)row * D + k] : 0.0;
            }
        }
        double best = INFINITY;
        int bidx = 0x7fffffff;
        // two columns per iteration, the next two prefetched while these are reduced: the first
        // version (one column, load -> convert -> FMA -> shuffle chain) stalled 11.6 cycles per
        // issue on the loads (ncu long_scoreboard)
        float nx[2][DPL];
        auto fetch = [&](int j, float (&dst)[DPL]) {
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                const int k = lane + 32 * i;
                dst[i] = (j < nb && k < D) ? __ldg(Bm + (size_t)j * D + k) : 0.f;
            }
        };
        fetch(warp * 2, nx[0]);
        fetch(warp * 2 + 1, nx[1]);
        for (int j = warp * 2; j < nb; j += RK_WARPS * 2) {
            float cur[2][DPL];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < DPL; ++i) cur[c][i] = nx[c][i];
            fetch(j + RK_WARPS * 2, nx[0]);
            fetch(j + RK_WARPS * 2 + 1, nx[1]);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                double acc[RK_ROWS];
#pragma unroll
                for (int r = 0; r < RK_ROWS; ++r) {
                    double sum = 0.0;
                    if (metric == MP_METRIC_NN) {
#pragma unroll
                        for (int i = 0; i < DPL; ++i) sum = fma(a[r][i], (double)cur[c][i], sum);
                    } else {
#pragma unroll
                        for (int i = 0; i < DPL; ++i) { const double df = a[r][i] - (double)cur[c][i]; sum = fma(df, df, sum); }
                    }
                    acc[r] = sum;
                }
                // halving exchange: after the 3 steps each lane holds the partial of one row
#pragma unroll
                for (int r = 0; r < 4; ++r) {  // step 1: partner lane^16, keep rows [0,4) or [4,8)
                    const bool up = lane & 16;
                    const double send = up ? acc[r] : acc[r + 4], keep = up ? acc[r + 4] : acc[r];
                    acc[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {  // step 2: lane^8
                    const bool up = lane & 8;
                    const double send = up ? acc[r] : acc[r + 2], keep = up ? acc[r + 2] : acc[r];
                    acc[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                {  // step 3: lane^4
                    const bool up = lane & 4;
                    const double send = up ? acc[0] : acc[1], keep = up ? acc[1] : acc[0];
                    acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 2);
                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
                // acc[0] is now the total of row ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1)
                double key = acc[0];
                if (metric == MP_METRIC_NN) key = -fmin(1.0, fmax(-1.0, key));
                if (j + c < nb && key < best) { best = key; bidx = j + c; }  // ascending j per warp: first minimum
            }
        }
        // lanes 4r..4r+3 (in the bit order above) all hold row r's result; publish one per warp
        const int r_of_lane = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        __syncthreads();
        if ((lane & 3) == 0) { red_key[warp][r_of_lane] = best; red_idx[warp][r_of_lane] = bidx; }
        __syncthreads();
        if (threadIdx.x < RK_ROWS && c0 + threadIdx.x < count) {
            double bk = INFINITY;
            int bi = 0x7fffffff;
            for (int w = 0; w < RK_WARPS; ++w) {
                const double k = red_key[w][threadIdx.x];
                const int i = red_idx[w][threadIdx.x];
                if (k < bk || (k == bk && i < bi)) { bk = k; bi = i; }
            }
            dst[rows[c0 + threadIdx.x]] = bi == 0x7fffffff ? -1 : bi;
        }
    }
}

// Few rows on a group's full list (the usual case: a handful per call).  The 8-row kernel above
// amortises the other set's traffic but its single chunk per group runs at the latency of one CTA
// walking 2 MB alone; one CTA per row was no better (58 us for ~26 rows: 32 dependent L2/DRAM round
// trips).  Here every flagged row is cut into RK_SPLIT slices of the other set, one 8-warp CTA per
// (group, slice) walking its slice two rows at a time with the next two prefetched (4 round trips
// for 2048 rows); the slices' (key, index) partials meet in global memory and the last CTA to
// arrive -- a per-row counter that it resets for the next call -- writes the winner.
constexpr int RK_SPLIT = 32;

template <int DPL>
__global__ void __launch_bounds__(256)
match_recheck_row_kernel(const float *__restrict__ d1, const int32_t *__restrict__ n1, int N1,
                         const float *__restrict__ d2, const int32_t *__restrict__ n2, int N2, int D, int metric,
                         const int32_t *__restrict__ flagged1, const int32_t *__restrict__ flagged2,
                         const int *__restrict__ n_flagged, double *__restrict__ part_key, int *__restrict__ part_idx,
                         int *__restrict__ row_done, int32_t *__restrict__ idx12, int32_t *__restrict__ idx21) {
    __shared__ double red_key[8];
    __shared__ int red_idx[8];
    __shared__ int is_last;
    const int group = blockIdx.x, p = group >> 1, side = group & 1, slice = blockIdx.y;
    const int count = n_flagged[group];
    if (count == 0 || count > RK_ROW_MAX) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NA = side == 0 ? N1 : N2;
    const float *A = side == 0 ? d1 + (size_t)p * N1 * D : d2 + (size_t)p * N2 * D;
    const float *Bm = side == 0 ? d2 + (size_t)p * N2 * D : d1 + (size_t)p * N1 * D;
    const int32_t *rows = (side == 0 ? flagged1 : flagged2) + (size_t)p * NA;
    int32_t *dst = side == 0 ? idx12 + (size_t)p * N1 : idx21 + (size_t)p * N2;
    const int nb = side == 0 ? (n2 ? min(n2[p], N2) : N2) : (n1 ? min(n1[p], N1) : N1);
    const int chunk = ((nb + RK_SPLIT - 1) / RK_SPLIT + 15) & ~15;  // multiple of the 16 rows one pass covers
    const int j_lo = slice * chunk, j_hi = min(nb, j_lo + chunk);
    for (int item = 0; item < count; ++item) {
        const int row = rows[item];
        double a[DPL];
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
            const int k = lane + 32 * i;
            a[i] = k < D ? (double)A[(size_t)row * D + k] : 0.0;
        }
        auto fetch = [&](int j, float (&dstv)[DPL]) {
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                const int k = lane + 32 * i;
                dstv[i] = (j < j_hi && k < D) ? __ldg(Bm + (size_t)j * D + k) : 0.f;
            }
        };
        float nx[2][DPL];
        fetch(j_lo + warp * 2, nx[0]);
        fetch(j_lo + warp * 2 + 1, nx[1]);
        double best = INFINITY;
        int bidx = 0x7fffffff;
        for (int j = j_lo + warp * 2; j < j_hi; j += 16) {
            float cur[2][DPL];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < DPL; ++i) cur[c][i] = nx[c][i];
            fetch(j + 16, nx[0]);
            fetch(j + 17, nx[1]);
            double acc[2] = {0.0, 0.0};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (metric == MP_METRIC_NN) {
#pragma unroll
                    for (int i = 0; i < DPL; ++i) acc[c] = fma(a[i], (double)cur[c][i], acc[c]);
                } else {
#pragma unroll
                    for (int i = 0; i < DPL; ++i) { const double df = a[i] - (double)cur[c][i]; acc[c] = fma(df, df, acc[c]); }
                }
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], sft);
                acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], sft);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                double key = acc[c];
                if (metric == MP_METRIC_NN) key = -fmin(1.0, fmax(-1.0, key));
                if (j + c < j_hi && key < best) { best = key; bidx = j + c; }
            }
        }
        __syncthreads();
        if (lane == 0) { red_key[warp] = best; red_idx[warp] = bidx; }
        __syncthreads();
        const size_t slot = ((size_t)group * RK_ROW_MAX + item) * RK_SPLIT;
        if (threadIdx.x == 0) {
            double bk = INFINITY;
            int bi = 0x7fffffff;
            for (int w = 0; w < 8; ++w)
                if (red_key[w] < bk || (red_key[w] == bk && red_idx[w] < bi)) { bk = red_key[w]; bi = red_idx[w]; }
            part_key[slot + slice] = bk;
            part_idx[slot + slice] = bi;
            __threadfence();
            is_last = atomicAdd(row_done + group * RK_ROW_MAX + item, 1) == RK_SPLIT - 1;
        }
        __syncthreads();
        if (is_last && warp == 0) {  // uniform per CTA
            __threadfence();
            double bk = __ldcg(part_key + slot + lane);
            int bi = __ldcg(part_idx + slot + lane);
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                const double ok = __shfl_xor_sync(0xffffffffu, bk, sft);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, sft);
                if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
            }
            if (lane == 0) {
                dst[row] = bi == 0x7fffffff ? -1 : bi;
                row_done[group * RK_ROW_MAX + item] = 0;  // ready for the next call
            }
        }
    }
}

// any D (> 256): one CTA per flagged row, threads stride over the other set's rows
__global__ void __launch_bounds__(256)
match_recheck_generic_kernel(const float *__restrict__ d1, const int32_t *__restrict__ n1, int N1,
                             const float *__restrict__ d2, const int32_t *__restrict__ n2, int N2, int D, int metric,
                             const int32_t *__restrict__ flagged1, const int32_t *__restrict__ flagged2,
                             const int *__restrict__ n_flagged, int32_t *__restrict__ idx12, int32_t *__restrict__ idx21) {
    extern __shared__ double arow[];  // [D]
    __shared__ double red_key[256];
    __shared__ int red_idx[256];
    const int group = blockIdx.x, p = group >> 1, side = group & 1;
    const int count = n_flagged[group];
    const int NA = side == 0 ? N1 : N2;
    const int32_t *rows = (side == 0 ? flagged1 : flagged2) + (size_t)p * NA;
    for (int item = blockIdx.y; item < count; item += gridDim.y) {
        const int row = rows[item];
        const float *a = side == 0 ? d1 + ((size_t)p * N1 + row) * D : d2 + ((size_t)p * N2 + row) * D;
        const float *b = side == 0 ? d2 + (size_t)p * N2 * D : d1 + (size_t)p * N1 * D;
        const int nb = side == 0 ? (n2 ? min(n2[p], N2) : N2) : (n1 ? min(n1[p], N1) : N1);
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += blockDim.x) arow[c] = (double)a[c];
        __syncthreads();
        double best = INFINITY;
        int bidx = 0x7fffffff;
        for (int j = threadIdx.x; j < nb; j += blockDim.x) {
            const float *bj = b + (size_t)j * D;
            double key = 0.0;
            if (metric == MP_METRIC_NN) {
                for (int c = 0; c < D; ++c) key = fma(arow[c], (double)bj[c], key);
                key = -fmin(1.0, fmax(-1.0, key));
            } else {
                for (int c = 0; c < D; ++c) { const double df = arow[c] - (double)bj[c]; key = fma(df, df, key); }
            }
            if (key < best) { best = key; bidx = j; }
        }
        red_key[threadIdx.x] = best;
        red_idx[threadIdx.x] = bidx;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                const double ok = red_key[threadIdx.x + s];
                const int oi = red_idx[threadIdx.x + s];
                if (ok < red_key[threadIdx.x] || (ok == red_key[threadIdx.x] && oi < red_idx[threadIdx.x])) {
                    red_key[threadIdx.x] = ok;
                    red_idx[threadIdx.x] = oi;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            int32_t *dst = side == 0 ? idx12 + (size_t)p * N1 : idx21 + (size_t)p * N2;
            dst[row] = red_idx[0] == 0x7fffffff ? -1 : red_idx[0];
        }
    }
}

// ------------------------------------------------------------------ outputs of mp_nearest_f32
__global__ void match_export_kernel(const Top2 *__restrict__ top, const float *__restrict__ norms_other, int N, int NO,
                                    int P, int use_bias, float *__restrict__ best, float *__restrict__ second) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)P * N) return;
    const int p = (int)(g / N);
    const Top2 t = top[g];
    // undo the L2 bias so the caller sees plain similarities a.b
    float b = t.best, s = t.second;
    if (use_bias) {
        if (t.best_idx >= 0) b += 0.5f * norms_other[(size_t)p * NO + t.best_idx];
        if (t.second_idx >= 0) s += 0.5f * norms_other[(size_t)p * NO + t.second_idx];
    }
    if (best) best[g] = b;
    if (second) second[g] = s;
}

// ------------------------------------------------------------------ select
// the reference's distance for one pair, fp64 accumulate, rounded once to fp32
__device__ __forceinline__ float pair_distance(const float *a, const float *b, int D, int metric, int lane) {
    double acc = 0.0;
    if (metric == MP_METRIC_NN) {
        for (int c = lane; c < D; c += 32) acc += (double)a[c] * (double)b[c];
    } else {
        for (int c = lane; c < D; c += 32) { const double df = (double)a[c] - (double)b[c]; acc += df * df; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (metric == MP_METRIC_NN) {
        float dot = (float)acc;
        dot = fminf(1.f, fmaxf(-1.f, dot));
        return sqrtf(2.f - 2.f * dot);  // matching.py:51
    }
    return sqrtf((float)acc);
}

// one warp per query row: decide keep, compute the distance
__global__ void __launch_bounds__(256)
match_decide_kernel(const float *__restrict__ d1, const int32_t *__restrict__ n1, int N1,
                    const float *__restrict__ d2, int N2, int P, int D, int metric, int kind, int cross_check,
                    float threshold, float ratio, const int32_t *__restrict__ idx12, const int32_t *__restrict__ idx21,
                    const Top2 *__restrict__ top12, int32_t *__restrict__ train_tmp, float *__restrict__ dist_tmp) {
    const int lane = threadIdx.x & 31;
    const long long g = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= (long long)P * N1) return;
    const int p = (int)(g / N1), i = (int)(g - (long long)p * N1);
    const int n_a = n1 ? min(n1[p], N1) : N1;
    int j = -1;
    float dist = 0.f;
    if (i < n_a) {
        j = idx12[g];
        if (j >= 0 && kind == MP_MATCH_MUTUAL && cross_check && idx21[(size_t)p * N2 + j] != i) j = -1;  // not mutual: no distance needed
        if (j >= 0) {
            const float *a = d1 + (size_t)g * D;
            dist = pair_distance(a, d2 + ((size_t)p * N2 + j) * D, D, metric, lane);
            if (kind == MP_MATCH_MUTUAL) {
                if (threshold >= 0.f && !(dist < threshold)) j = -1;
            } else {
                // knnMatch(k=2) + Lowe ratio (matching.py:21-28); needs a second neighbour.  The nearest index is
                // exact (fp64 recheck); the SECOND neighbour is the runner-up of the approximate top-3, i.e. exact only up
                // to the key error bound (~5e-5 |a||b|): when the second and third keys are closer than that, dist2 can be
                // the third neighbour's distance, which differs from the second's by no more than the same bound
                const Top2 t = top12[g];
                int j2 = (t.best_idx == j) ? t.second_idx : t.best_idx;
                if (j2 < 0) {
                    j = -1;
                } else {
                    const float dist2 = pair_distance(a, d2 + ((size_t)p * N2 + j2) * D, D, metric, lane);
                    if (!(dist < ratio * dist2)) j = -1;
                }
            }
        }
    }
    if (lane == 0) {
        train_tmp[g] = j;
        dist_tmp[g] = dist;
    }
}

// Mutual matching (the BFMatcher-crossCheck / NNMatcher case), D <= 256: one thread per query row for the index
// work -- coalesced idx12, one gather from idx21, all rows of the block in flight at once -- then each warp computes
// the exact distance of its mutual rows cooperatively, two rows per pass so that both rows' loads are issued
// before either reduction.  The warp-per-row kernel above serialised three dependent round trips per row
// (68 us for 131 k rows, 44 % of them mutual); arithmetic and results are identical (same per-lane order).
template <int DPL>
__global__ void __launch_bounds__(256)
match_decide_mutual_kernel(const float *__restrict__ d1, const int32_t *__restrict__ n1, int N1,
                           const float *__restrict__ d2, int N2, int P, int D, int metric, int cross_check,
                           float threshold, const int32_t *__restrict__ idx12, const int32_t *__restrict__ idx21,
                           int32_t *__restrict__ train_tmp, float *__restrict__ dist_tmp) {
    const int lane = threadIdx.x & 31;
    const long long total = (long long)P * N1;
    const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long gw = g - lane;  // first row of this warp
    int j = -1, p = 0;
    if (g < total) {
        p = (int)(g / N1);
        const int i = (int)(g - (long long)p * N1);
        const int n_a = n1 ? min(n1[p], N1) : N1;
        if (i < n_a) {
            j = idx12[g];
            if (j >= 0 && cross_check && idx21[(size_t)p * N2 + j] != i) j = -1;
        }
    }
    float dist = 0.f;
    unsigned todo = __ballot_sync(0xffffffffu, j >= 0);
    while (todo) {
        const int l0 = __ffs(todo) - 1;
        todo &= todo - 1;
        const int l1 = todo ? __ffs(todo) - 1 : l0;  // a lone row is simply computed twice
        todo &= todo - 1;
        const int j0 = __shfl_sync(0xffffffffu, j, l0), j1 = __shfl_sync(0xffffffffu, j, l1);
        const int p0 = __shfl_sync(0xffffffffu, p, l0), p1 = __shfl_sync(0xffffffffu, p, l1);
        const float *a0 = d1 + (size_t)(gw + l0) * D, *b0 = d2 + ((size_t)p0 * N2 + j0) * D;
        const float *a1 = d1 + (size_t)(gw + l1) * D, *b1 = d2 + ((size_t)p1 * N2 + j1) * D;
        float va0[DPL], vb0[DPL], va1[DPL], vb1[DPL];
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
            const int c = lane + 32 * i;
            const bool in = c < D;
            va0[i] = in ? __ldg(a0 + c) : 0.f; vb0[i] = in ? __ldg(b0 + c) : 0.f;
            va1[i] = in ? __ldg(a1 + c) : 0.f; vb1[i] = in ? __ldg(b1 + c) : 0.f;
        }
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
            if (lane + 32 * i < D) {  // same terms, same order as pair_distance
                if (metric == MP_METRIC_NN) {
                    acc0 += (double)va0[i] * (double)vb0[i];
                    acc1 += (double)va1[i] * (double)vb1[i];
                } else {
                    const double f0 = (double)va0[i] - (double)vb0[i], f1 = (double)va1[i] - (double)vb1[i];
                    acc0 += f0 * f0;
                    acc1 += f1 * f1;
                }
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, sft);
            acc1 += __shfl_xor_sync(0xffffffffu, acc1, sft);
        }
        float r0, r1;
        if (metric == MP_METRIC_NN) {
            r0 = sqrtf(2.f - 2.f * fminf(1.f, fmaxf(-1.f, (float)acc0)));  // matching.py:51
            r1 = sqrtf(2.f - 2.f * fminf(1.f, fmaxf(-1.f, (float)acc1)));
        } else {
            r0 = sqrtf((float)acc0);
            r1 = sqrtf((float)acc1);
        }
        if (lane == l0) dist = r0;
        if (lane == l1) dist = r1;
    }
    if (j >= 0 && threshold >= 0.f && !(dist < threshold)) j = -1;
    if (g < total) {
        train_tmp[g] = j;
        dist_tmp[g] = dist;
    }
}

// one CTA per pair: ordered compaction (ascending query index)
__global__ void __launch_bounds__(1024)
match_compact_kernel(const int32_t *__restrict__ train_tmp, const float *__restrict__ dist_tmp, int N1,
                     int32_t *__restrict__ query, int32_t *__restrict__ train, float *__restrict__ dist,
                     int32_t *__restrict__ counts) {
    __shared__ int warp_sums[32];
    __shared__ int running;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int i0 = 0; i0 < N1; i0 += 1024) {
        const int i = i0 + tid;
        const int j = i < N1 ? train_tmp[(size_t)p * N1 + i] : -1;
        const int keep = j >= 0;
        int inc = keep;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, s);
            if (lane >= s) inc += t;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        int base = running;
        for (int w = 0; w < warp; ++w) base += warp_sums[w];
        if (keep) {
            const size_t o = (size_t)p * N1 + base + inc - 1;
            query[o] = i;
            train[o] = j;
            dist[o] = dist_tmp[(size_t)p * N1 + i];
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += warp_sums[w];
            running += tot;
        }
        __syncthreads();
    }
    if (tid == 0) counts[p] = running;
}

// ------------------------------------------------------------------ ThresholdMatcher
// fp32 tiles like the SIMT top-2, two passes: count per row, then fill in row-major order.
__global__ void __launch_bounds__(256)
threshold_rows_kernel(const float *__restrict__ d1, int N1, const float *__restrict__ d2, int N2, int D,
                      float threshold, const long long *__restrict__ row_offsets, long long *__restrict__ row_counts,
                      int32_t *__restrict__ query, int32_t *__restrict__ train, float *__restrict__ dist, long long cap) {
    // one warp per query row; lanes stride over train rows; ballot keeps row-major order
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= N1) return;
    const float *a = d1 + (size_t)i * D;
    long long n = 0;
    const long long base = row_offsets ? row_offsets[i] : 0;
    for (int j0 = 0; j0 < N2; j0 += 32) {
        const int j = j0 + lane;
        float dv = INFINITY;
        if (j < N2) {
            const float *b = d2 + (size_t)j * D;
            double acc = 0.0;
            for (int c = 0; c < D; ++c) acc += (double)a[c] * (double)b[c];
            float dot = fminf(1.f, fmaxf(-1.f, (float)acc));
            dv = sqrtf(2.f - 2.f * dot);
        }
        const bool hit = dv < threshold;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (row_offsets && hit) {
            const long long o = base + n + __popc(m & ((1u << lane) - 1));
            if (o < cap) { query[o] = i; train[o] = j; dist[o] = dv; }
        }
        n += __popc(m);
    }
    if (!row_offsets && lane == 0) row_counts[i] = n;
}

__global__ void __launch_bounds__(1024)
exclusive_scan_ll_kernel(const long long *__restrict__ in, long long *__restrict__ out, int n, long long *__restrict__ total) {
    __shared__ long long warp_sums[32];
    __shared__ long long running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        const long long v = i < n ? in[i] : 0;
        long long inc = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, s);
            if (lane >= s) inc += t;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        long long base = running;
        for (int w = 0; w < warp; ++w) base += warp_sums[w];
        if (i < n) out[i] = base + inc - v;
        __syncthreads();
        if (tid == 0) {
            long long tot = 0;
            for (int w = 0; w < 32; ++w) tot += warp_sums[w];
            running += tot;
        }
        __syncthreads();
    }
    if (tid == 0) *total = running;
}

// ------------------------------------------------------------------ host orchestration
MatchLayout::MatchLayout(int P, int N1, int N2, int D) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t r1 = (size_t)P * N1, r2 = (size_t)P * N2;
    scalars = take(sizeof(unsigned) * (6 * (size_t)P + 4 + 2 * (size_t)P * RK_ROW_MAX));  // ... + row_done[2P][RK_ROW_MAX]
    row_part_key = take(sizeof(double) * 2 * (size_t)P * RK_ROW_MAX * RK_SPLIT);
    row_part_idx = take(sizeof(int) * 2 * (size_t)P * RK_ROW_MAX * RK_SPLIT);
    norms1 = take(sizeof(float) * r1);
    norms2 = take(sizeof(float) * r2);
    top12 = take(sizeof(Top2) * r1);
    top21 = take(sizeof(Top2) * r2);
    idx12 = take(sizeof(int32_t) * r1);
    idx21 = take(sizeof(int32_t) * r2);
    flagged1 = take(sizeof(int32_t) * r1);
    flagged2 = take(sizeof(int32_t) * r2);
    pairs1 = take(sizeof(int32_t) * r1);
    pairs2 = take(sizeof(int32_t) * r2);
    train_tmp = take(sizeof(int32_t) * r1);
    dist_tmp = take(sizeof(float) * r1);
    hi1 = take(sizeof(__nv_bfloat16) * r1 * D);
    mid1 = take(sizeof(__nv_bfloat16) * r1 * D);
    hi2 = take(sizeof(__nv_bfloat16) * r2 * D);
    mid2 = take(sizeof(__nv_bfloat16) * r2 * D);
    colpart = take(match_colpart_bytes(P, N1, N2));
    total = off;
}

// Operands of one descriptor set already split for the tensor path (mp_sample_descriptors_split_f32): bf16 planes,
// squared norms, and an upper bound of the largest norm per pair (float bits).
struct SplitSet {
    const __nv_bfloat16 *hi, *mid;
    const float *norms;
    const unsigned *max_norm;
};

// Fills ws.idx12 / ws.idx21 (exact) and ws.top12 / ws.top21 (approximate keys).
static int run_nearest(const float *d1, const int32_t *n1, int N1, const float *d2, const int32_t *n2, int N2,
                       int P, int D, int metric, int algo, const MatchLayout &L, char *ws, cudaStream_t s,
                       const SplitSet *s1 = nullptr, const SplitSet *s2 = nullptr) {
    unsigned *scal = (unsigned *)(ws + L.scalars);
    const unsigned *max1 = s1 ? s1->max_norm : scal, *max2 = s2 ? s2->max_norm : scal + P;
    int *n_flagged = (int *)(scal + 2 * P), *n_pairs = (int *)(scal + 4 * P);
    const float *norms1 = s1 ? s1->norms : (float *)(ws + L.norms1), *norms2 = s2 ? s2->norms : (float *)(ws + L.norms2);
    Top2 *top12 = (Top2 *)(ws + L.top12), *top21 = (Top2 *)(ws + L.top21);
    int32_t *idx12 = (int32_t *)(ws + L.idx12), *idx21 = (int32_t *)(ws + L.idx21);
    int32_t *flagged1 = (int32_t *)(ws + L.flagged1), *flagged2 = (int32_t *)(ws + L.flagged2);
    int32_t *pairs1 = (int32_t *)(ws + L.pairs1), *pairs2 = (int32_t *)(ws + L.pairs2);
    const bool tensor = algo == MP_ALGO_TENSOR;
    const __nv_bfloat16 *hi1 = s1 ? s1->hi : tensor ? (__nv_bfloat16 *)(ws + L.hi1) : nullptr, *mid1 = s1 ? s1->mid : (__nv_bfloat16 *)(ws + L.mid1);
    const __nv_bfloat16 *hi2 = s2 ? s2->hi : tensor ? (__nv_bfloat16 *)(ws + L.hi2) : nullptr, *mid2 = s2 ? s2->mid : (__nv_bfloat16 *)(ws + L.mid2);
    const int use_bias = metric == MP_METRIC_L2;

    MP_CUDA_OK(cudaMemsetAsync(scal, 0, sizeof(unsigned) * (6 * (size_t)P + 4 + 2 * (size_t)P * RK_ROW_MAX), s));
    const long long r1 = (long long)P * N1, r2 = (long long)P * N2;
    auto prep = [&](const float *d, const int32_t *n, int N, long long rows, float *norms, unsigned *mx, __nv_bfloat16 *h,
                    __nv_bfloat16 *m) {
        const bool vec = D % 8 == 0 && (((uintptr_t)d | (uintptr_t)h | (uintptr_t)m) & 15) == 0;
        if (!vec) match_prep_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(d, n, N, D, P, norms, mx, h, m);
        else if (D <= 64) match_prep_vec_kernel<4><<<(unsigned)((rows + 31) / 32), 256, 0, s>>>(d, n, N, D, P, norms, mx, h, m);
        else if (D <= 128) match_prep_vec_kernel<2><<<(unsigned)((rows + 15) / 16), 256, 0, s>>>(d, n, N, D, P, norms, mx, h, m);
        else match_prep_vec_kernel<1><<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(d, n, N, D, P, norms, mx, h, m);
    };
    if (!s1) {
        prep(d1, n1, N1, r1, (float *)(ws + L.norms1), scal, (__nv_bfloat16 *)hi1, (__nv_bfloat16 *)mid1);
        MP_LAUNCH_OK_S("match_prep_vec_kernel", s);
    }
    if (!s2) {
        prep(d2, n2, N2, r2, (float *)(ws + L.norms2), scal + P, (__nv_bfloat16 *)hi2, (__nv_bfloat16 *)mid2);
        MP_LAUNCH_OK_S("match_prep_vec_kernel", s);
    }

    if (tensor) {
        // one GEMM per pair: rows of set 1 against set 2 (top12) and, from the same accumulators, the columns' view
        // = rows of set 2 against set 1 (top21)
        int rc = match_top2_tensor(hi1, mid1, n1, N1, hi2, mid2, n2, N2, P, D, norms1, norms2, use_bias, max1, max2, top12, top21,
                                   ws + L.colpart, s);
        if (rc != MP_OK) return rc;
    } else {
        dim3 g1((N1 + ST_ROWS - 1) / ST_ROWS, P), g2((N2 + ST_ROWS - 1) / ST_ROWS, P);
        match_top2_simt_kernel<<<g1, ST_ROWS, 0, s>>>(d1, n1, N1, d2, n2, N2, D, norms2, use_bias, top12);
        MP_LAUNCH_OK_S("match_top2_simt_kernel", s);
        match_top2_simt_kernel<<<g2, ST_ROWS, 0, s>>>(d2, n2, N2, d1, n1, N1, D, norms1, use_bias, top21);
        MP_LAUNCH_OK_S("match_top2_simt_kernel", s);
    }
    const float eps_rel = tensor ? MATCH_EPS_TENSOR : MATCH_EPS_SIMT, pack_rel = tensor ? MATCH_PACK_REL : 0.f;
    {   // both directions in one launch
        const FlagSide f0{top12, N1, max1, max2, idx12, flagged1, pairs1}, f1{top21, N2, max2, max1, idx21, flagged2, pairs2};
        const long long rmax = r1 > r2 ? r1 : r2;
        match_flag_kernel<<<dim3((unsigned)((rmax + 255) / 256), 2), 256, 0, s>>>(f0, f1, P, metric, eps_rel, pack_rel, n_flagged, n_pairs);
        MP_LAUNCH_OK_S("match_flag_kernel", s);
    }
    {
        dim3 grid(2 * P, RP_Y);
        match_recheck_pair_kernel<<<grid, 256, 0, s>>>(d1, N1, d2, N2, D, metric, top12, top21, pairs1, pairs2, n_pairs, idx12, idx21);
        MP_LAUNCH_OK_S("match_recheck_pair_kernel", s);
    }
    if (D <= RK_MAXD) {
        dim3 grid(2 * P, RK_Z);
#define MP_RECHECK(DPL) match_recheck_kernel<DPL><<<grid, RK_WARPS * 32, 0, s>>>(d1, n1, N1, d2, n2, N2, D, metric, flagged1, flagged2, n_flagged, idx12, idx21)
        if (D <= 64) MP_RECHECK(2);
        else if (D <= 128) MP_RECHECK(4);
        else MP_RECHECK(8);
#undef MP_RECHECK
        MP_LAUNCH_OK_S("match_recheck_kernel", s);
        dim3 grid_row(2 * P, RK_SPLIT);
        double *part_key = (double *)(ws + L.row_part_key);
        int *part_idx = (int *)(ws + L.row_part_idx), *row_done = (int *)(scal + 6 * (size_t)P + 4);
#define MP_RECHECK_ROW(DPL) match_recheck_row_kernel<DPL><<<grid_row, 256, 0, s>>>(d1, n1, N1, d2, n2, N2, D, metric, flagged1, flagged2, n_flagged, part_key, part_idx, row_done, idx12, idx21)
        if (D <= 64) MP_RECHECK_ROW(2);
        else if (D <= 128) MP_RECHECK_ROW(4);
        else MP_RECHECK_ROW(8);
#undef MP_RECHECK_ROW
        MP_LAUNCH_OK_S("match_recheck_row_kernel", s);
    } else {
        dim3 grid(2 * P, 32);
        match_recheck_generic_kernel<<<grid, 256, sizeof(double) * D, s>>>(d1, n1, N1, d2, n2, N2, D, metric, flagged1, flagged2, n_flagged, idx12, idx21);
        MP_LAUNCH_OK_S("match_recheck_generic_kernel", s);
    }
    return MP_OK;
}

static int check_match_args(const char *fn, const float *d1, int N1, const float *d2, int N2, int P, int D, int metric,
                            int algo, void *workspace, size_t workspace_bytes, const MatchLayout &L) {
    MP_CHECK_ARG(P >= 0 && N1 >= 0 && N2 >= 0 && D > 0, "%s: bad shape P=%d N1=%d N2=%d D=%d", fn, P, N1, N2, D);
    MP_CHECK_ARG(metric == MP_METRIC_NN || metric == MP_METRIC_L2, "%s: bad metric %d", fn, metric);
    MP_CHECK_ARG(algo == MP_ALGO_TENSOR || algo == MP_ALGO_SIMT, "%s: bad algo %d", fn, algo);
    MP_CHECK_ARG(D <= 4096, "%s: D=%d too large", fn, D);
    if (algo == MP_ALGO_TENSOR && (D % 64 != 0 || D > 256)) {
        set_error("%s: the tensor path needs D %% 64 == 0 and D <= 256 (got %d); use MP_ALGO_SIMT", fn, D);
        return MP_ERR_UNSUPPORTED;
    }
    if ((long long)P * N1 == 0 || (long long)P * N2 == 0) return MP_OK;
    MP_CHECK_ARG(d1 && d2, "%s: null descriptor pointer", fn);
    if (workspace == nullptr || workspace_bytes < L.total) {
        set_error("%s: workspace %zu B < required %zu B", fn, workspace_bytes, L.total);
        return MP_ERR_WORKSPACE;
    }
    return MP_OK;
}

__global__ void fill_i32_kernel(int32_t *p, long long n, int32_t v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace mp

extern "C" size_t mp_match_workspace_bytes(int P, int N1, int N2, int D) {
    if (P <= 0 || D <= 0) return 256;
    return mp::MatchLayout(P, N1 > 0 ? N1 : 0, N2 > 0 ? N2 : 0, D).total;
}

extern "C" int mp_nearest_f32(const float *d1, const int32_t *n1, int N1, const float *d2,
                              const int32_t *n2, int N2, int P, int D, int metric, int algo,
                              int32_t *idx12, float *best12, float *second12, int32_t *idx21,
                              float *best21, float *second21, void *workspace,
                              size_t workspace_bytes, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    using namespace mp;
    const MatchLayout L(P > 0 ? P : 0, N1 > 0 ? N1 : 0, N2 > 0 ? N2 : 0, D > 0 ? D : 1);
    int rc = check_match_args("mp_nearest_f32", d1, N1, d2, N2, P, D, metric, algo, workspace, workspace_bytes, L);
    if (rc != MP_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const long long r1 = (long long)P * N1, r2 = (long long)P * N2;
    if (r1 == 0 || r2 == 0) {  // one side empty: no neighbours
        if (r1 && idx12) { fill_i32_kernel<<<(unsigned)((r1 + 255) / 256), 256, 0, s>>>(idx12, r1, -1); MP_LAUNCH_OK_S("fill_i32_kernel", s); }
        if (r2 && idx21) { fill_i32_kernel<<<(unsigned)((r2 + 255) / 256), 256, 0, s>>>(idx21, r2, -1); MP_LAUNCH_OK_S("fill_i32_kernel", s); }
        return MP_OK;
    }
    char *ws = (char *)workspace;
    rc = run_nearest(d1, n1, N1, d2, n2, N2, P, D, metric, algo, L, ws, s);
    if (rc != MP_OK) return rc;
    if (idx12) MP_CUDA_OK(cudaMemcpyAsync(idx12, ws + L.idx12, sizeof(int32_t) * r1, cudaMemcpyDeviceToDevice, s));
    if (idx21) MP_CUDA_OK(cudaMemcpyAsync(idx21, ws + L.idx21, sizeof(int32_t) * r2, cudaMemcpyDeviceToDevice, s));
    const int use_bias = metric == MP_METRIC_L2;
    if (best12 || second12) {
        match_export_kernel<<<(unsigned)((r1 + 255) / 256), 256, 0, s>>>((Top2 *)(ws + L.top12), (float *)(ws + L.norms2), N1, N2, P, use_bias, best12, second12);
        MP_LAUNCH_OK_S("match_export_kernel", s);
    }
    if (best21 || second21) {
        match_export_kernel<<<(unsigned)((r2 + 255) / 256), 256, 0, s>>>((Top2 *)(ws + L.top21), (float *)(ws + L.norms1), N2, N1, P, use_bias, best21, second21);
        MP_LAUNCH_OK_S("match_export_kernel", s);
    }
    return MP_OK;
}

static int match_impl(const char *fn, const float *d1, const int32_t *n1, int N1, const float *d2,
                      const int32_t *n2, int N2, int P, int D, int metric, int algo, int kind,
                      int cross_check, double threshold, double ratio, int32_t *query,
                      int32_t *train, float *dist, int32_t *counts, void *workspace,
                      size_t workspace_bytes, mp_stream_t stream, const mp::SplitSet *s1, const mp::SplitSet *s2) {
    using namespace mp;
    const MatchLayout L(P > 0 ? P : 0, N1 > 0 ? N1 : 0, N2 > 0 ? N2 : 0, D > 0 ? D : 1);
    int rc = check_match_args(fn, d1, N1, d2, N2, P, D, metric, algo, workspace, workspace_bytes, L);
    if (rc != MP_OK) return rc;
    MP_CHECK_ARG(kind == MP_MATCH_MUTUAL || kind == MP_MATCH_RATIO, "%s: bad kind %d", fn, kind);
    if (P == 0) return MP_OK;
    MP_CHECK_ARG(counts != nullptr, "%s: counts is required", fn);
    cudaStream_t s = (cudaStream_t)stream;
    if ((long long)P * N1 == 0 || (long long)P * N2 == 0) {
        MP_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * P, s));
        return MP_OK;
    }
    MP_CHECK_ARG(query && train && dist, "%s: null output pointer", fn);
    char *ws = (char *)workspace;
    rc = run_nearest(d1, n1, N1, d2, n2, N2, P, D, metric, algo, L, ws, s, s1, s2);
    if (rc != MP_OK) return rc;
    const long long r1 = (long long)P * N1;
    int32_t *train_tmp = (int32_t *)(ws + L.train_tmp);
    float *dist_tmp = (float *)(ws + L.dist_tmp);
    if (kind == MP_MATCH_MUTUAL && D <= 256) {
        const unsigned grid = (unsigned)((r1 + 255) / 256);
#define MP_DECIDE(DPL) match_decide_mutual_kernel<DPL><<<grid, 256, 0, s>>>(d1, n1, N1, d2, N2, P, D, metric, cross_check, (float)threshold, \
                                                                          (int32_t *)(ws + L.idx12), (int32_t *)(ws + L.idx21), train_tmp, dist_tmp)
        if (D <= 64) MP_DECIDE(2);
        else if (D <= 128) MP_DECIDE(4);
        else MP_DECIDE(8);
#undef MP_DECIDE
        MP_LAUNCH_OK_S("match_decide_mutual_kernel", s);
    } else {
        match_decide_kernel<<<(unsigned)((r1 + 7) / 8), 256, 0, s>>>(d1, n1, N1, d2, N2, P, D, metric, kind, cross_check,
                                                                   (float)threshold, (float)ratio, (int32_t *)(ws + L.idx12),
                                                                   (int32_t *)(ws + L.idx21), (Top2 *)(ws + L.top12), train_tmp, dist_tmp);
        MP_LAUNCH_OK_S("match_decide_kernel", s);
    }
    match_compact_kernel<<<P, 1024, 0, s>>>(train_tmp, dist_tmp, N1, query, train, dist, counts);
    MP_LAUNCH_OK_S("match_compact_kernel", s);
    return MP_OK;
}

extern "C" int mp_match_f32(const float *d1, const int32_t *n1, int N1, const float *d2,
                            const int32_t *n2, int N2, int P, int D, int metric, int algo, int kind,
                            int cross_check, double threshold, double ratio, int32_t *query,
                            int32_t *train, float *dist, int32_t *counts, void *workspace,
                            size_t workspace_bytes, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    return match_impl("mp_match_f32", d1, n1, N1, d2, n2, N2, P, D, metric, algo, kind, cross_check, threshold, ratio, query, train,
                      dist, counts, workspace, workspace_bytes, stream, nullptr, nullptr);
}

extern "C" int mp_match_split_f32(const float *d1, const void *hi1, const void *mid1, const float *sq_norms1,
                                  const uint32_t *max_norm1, const int32_t *n1, int N1, const float *d2,
                                  const void *hi2, const void *mid2, const float *sq_norms2,
                                  const uint32_t *max_norm2, const int32_t *n2, int N2, int P, int D, int metric,
                                  int kind, int cross_check, double threshold, double ratio, int32_t *query,
                                  int32_t *train, float *dist, int32_t *counts, void *workspace,
                                  size_t workspace_bytes, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(hi1 && mid1 && sq_norms1 && max_norm1 && hi2 && mid2 && sq_norms2 && max_norm2,
                 "mp_match_split_f32: null split operand");
    MP_CHECK_ARG((((uintptr_t)hi1 | (uintptr_t)mid1 | (uintptr_t)hi2 | (uintptr_t)mid2) & 15) == 0,
                 "mp_match_split_f32: the bf16 planes must be 16-byte aligned (TMA)");
    const mp::SplitSet s1 = {(const __nv_bfloat16 *)hi1, (const __nv_bfloat16 *)mid1, sq_norms1, max_norm1};
    const mp::SplitSet s2 = {(const __nv_bfloat16 *)hi2, (const __nv_bfloat16 *)mid2, sq_norms2, max_norm2};
    return match_impl("mp_match_split_f32", d1, n1, N1, d2, n2, N2, P, D, metric, MP_ALGO_TENSOR, kind, cross_check, threshold, ratio,
                      query, train, dist, counts, workspace, workspace_bytes, stream, &s1, &s2);
}

extern "C" int mp_match_threshold_f32(const float *d1, int N1, const float *d2, int N2, int D,
                                      double threshold, int32_t *query, int32_t *train, float *dist,
                                      int64_t cap, int64_t *total_host, void *workspace,
                                      size_t workspace_bytes, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    using namespace mp;
    MP_CHECK_ARG(N1 >= 0 && N2 >= 0 && D > 0 && cap >= 0, "mp_match_threshold_f32: bad shape");
    MP_CHECK_ARG(total_host != nullptr, "mp_match_threshold_f32: total_host is required");
    *total_host = 0;
    if (N1 == 0 || N2 == 0) return MP_OK;
    MP_CHECK_ARG(d1 && d2, "mp_match_threshold_f32: null pointer");
    const size_t need = align_up(sizeof(long long) * (2 * (size_t)N1 + 1), 256);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("mp_match_threshold_f32: workspace %zu B < required %zu B", workspace_bytes, need);
        return MP_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    long long *row_counts = (long long *)workspace, *row_offsets = row_counts + N1, *total = row_offsets + N1;
    const unsigned grid = (unsigned)((N1 + 7) / 8);
    threshold_rows_kernel<<<grid, 256, 0, s>>>(d1, N1, d2, N2, D, (float)threshold, nullptr, row_counts, nullptr, nullptr, nullptr, 0);
    MP_LAUNCH_OK_S("threshold_rows_kernel", s);
    exclusive_scan_ll_kernel<<<1, 1024, 0, s>>>(row_counts, row_offsets, N1, total);
    MP_LAUNCH_OK_S("exclusive_scan_ll_kernel", s);
    if (cap > 0) {
        MP_CHECK_ARG(query && train && dist, "mp_match_threshold_f32: null output pointer");
        threshold_rows_kernel<<<grid, 256, 0, s>>>(d1, N1, d2, N2, D, (float)threshold, row_offsets, nullptr, query, train, dist, cap);
        MP_LAUNCH_OK_S("threshold_rows_kernel", s);
    }
    long long t = 0;
    MP_CUDA_OK(cudaMemcpyAsync(&t, total, sizeof(long long), cudaMemcpyDeviceToHost, s));
    MP_CUDA_OK(cudaStreamSynchronize(s));
    *total_host = t;
    return MP_OK;
}
