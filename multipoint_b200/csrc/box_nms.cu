// Row 4 / 4b of the hot path: utils.box_nms (multipoint/utils/utils.py:78-122, on top of
// torchvision.ops.nms / batched_nms) and the torch.nonzero keypoint idiom
// (predict_align_image_pair.py:170-171, evaluation.py:157-158, export_keypoints.py:100).
//
// The reference thresholds, sorts all candidates and runs torchvision's O(N^2) greedy IoU scan
// (seconds per image on the CPU path its configs force).  Boxes are size x size squares centred
// on integer pixels, so "j is suppressed by a kept i" depends only on the offset (dy,dx): the
// host evaluates torchvision's fp32 IoU expression once per offset into a footprint bitmask.
// Greedy NMS is then the unique fixed point of
//     kept(p)       <=> no higher-priority neighbour in the footprint is kept
//     priority      =   (score desc, row-major index asc)      [stable sort of the reference]
// which is computed in parallel with three-valued logic: an undecided pixel becomes suppressed as
// soon as one higher-priority neighbour is kept, and kept once all of them are decided not-kept.
// Decisions are only taken on settled facts, so any evaluation order reaches the same result.
//
// Kernels (algorithmic traffic: 4 B read + 4 B written per pixel, plus 8 B per survivor):
//  dense path (no top-k, footprints that reach beyond 3 px, images the sparse path hands back)
//   1. nms_tile_fast_kernel / nms_tile_kernel   one CTA per TH x TW tile with an E-pixel apron staged in
//                        shared memory; iterates to the local fixed point over candidate lists, writes the
//                        dense result once.  Pixels whose dependency chain leaves the apron (<0.1 % at E=8)
//                        are written as -score and queued on a per-image worklist.
//   2. nms_fixup_select_kernel  one CTA per image; first resolves the worklist against the dense map in L2 (nms_fixup_body),
//                        Exits immediately when the list is empty.
//  sparse top-k path (keep_top_k > 0; the reference's shipped configs use topk: 0 = the dense path above; see the
//  comment above nms_candidates_kernel)
//   1'. nms_candidates_kernel  streams the heatmap once: lists the candidates and histograms their scores (and
//                        zero-fills the dense map when the caller wants it).
//   2'. nms_sparse2_kernel one CTA per image settles only the candidates that can reach the top k, cuts to k and emits.
//  both
//   3. then, in the same launch (nms_select_body): optional top-k by radix select on (score desc, index
//                        asc), then ordered (row-major) compaction of the survivors through a
//                        bitmap into int64 (y,x) keypoints.
#include <math.h>
#include <stdlib.h>

#include "mp_common.cuh"

namespace mp {

struct NmsFootprint {
    int R;
    uint32_t rows[31];  // rows[dy+R] bit (dx+R) set <=> offset (dy,dx) suppresses
};

constexpr int NMS_THREADS = 256;

// ------------------------------------------------------------------------------------------
// block-wide exclusive scan of one int per thread (NMS_THREADS threads); returns the offset of
// this thread and the block total through `total`.
__device__ __forceinline__ int block_exclusive_scan(int val, int *warp_sums, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    int inc = val;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += t;
    }
    __syncthreads();  // warp_sums may still be read from a previous call
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    int base = 0;
    total = 0;
    for (int w = 0; w < nwarps; ++w) {
        const int s = warp_sums[w];
        if (w < warp) base += s;
        total += s;
    }
    return base + inc - val;
}

// Decide one undecided pixel at smem/gmem position given a neighbour fetch functor.
// Returns +score (kept), 0 (suppressed) or the unchanged negative value (still undecided).
template <typename Fetch>
__device__ __forceinline__ float nms_decide(float val, const NmsFootprint &fp, Fetch fetch) {
    const float s = -val;
    const int R = fp.R;
    bool blocked = false;
    for (int dy = -R; dy <= R; ++dy) {
        uint32_t row = fp.rows[dy + R];
        while (row) {
            const int bit = __ffs(row) - 1;
            row &= row - 1;
            const int dx = bit - R;
            const float nv = fetch(dy, dx);
            if (nv == 0.f) continue;
            const float sn = fabsf(nv);
            const bool higher = sn > s || (sn == s && (dy < 0 || (dy == 0 && dx < 0)));
            if (!higher) continue;
            if (nv > 0.f) return 0.f;  // a kept higher-priority neighbour suppresses us
            blocked = true;            // an undecided one: wait
        }
    }
    return blocked ? val : s;
}

// Same decision for a pixel of the staged tile, with a candidate bitmap as pre-filter: the 2R+1
// neighbours of one footprint row are one funnel-shifted 32-bit window of the bitmap row, so only
// actual candidates (~15 % of the footprint) are fetched and compared.
template <int EW, int BW>
__device__ __forceinline__ float nms_decide_tile(float val, int ey, int ex, const float *v, const uint32_t *bm,
                                                 const NmsFootprint &fp) {
    const float s = -val;
    const int R = fp.R;
    const int bitpos = ex - R, w = bitpos >> 5, sh = bitpos & 31;
    bool blocked = false;
    for (int dy = -R; dy <= R; ++dy) {
        const uint32_t *bw = bm + (ey + dy) * BW + w;
        uint32_t win = __funnelshift_r(bw[0], bw[1], sh) & fp.rows[dy + R];
        const float *vr = v + (ey + dy) * EW + bitpos;
        while (win) {
            const int bit = __ffs(win) - 1;
            win &= win - 1;
            const float nv = vr[bit];
            if (nv == 0.f) continue;  // was a candidate, already suppressed
            const float sn = fabsf(nv);
            const int dx = bit - R;
            const bool higher = sn > s || (sn == s && (dy < 0 || (dy == 0 && dx < 0)));
            if (!higher) continue;
            if (nv > 0.f) return 0.f;
            blocked = true;
        }
    }
    return blocked ? val : s;
}

template <int TH, int TW, int E, bool VEC>
__global__ void __launch_bounds__(NMS_THREADS)
nms_tile_kernel(const float *__restrict__ prob, float *__restrict__ out, int H, int W, float thr,
                const NmsFootprint fp, uint2 *__restrict__ survivors, int *__restrict__ surv_count,
                uint32_t *__restrict__ worklist, int *__restrict__ work_count, int cap) {
    constexpr int EH = TH + 2 * E, EW = TW + 2 * E;
    static_assert(EW % 4 == 0 && E % 4 == 0 && TW % 4 == 0, "float4 staging needs 4-px alignment");
    static_assert(EH * EW < 65536, "list entries are uint16");
    constexpr int BW = (EW + 31) / 32 + 1;                          // bitmap words per row (+1: 64-bit windows)
    extern __shared__ __align__(16) float smem[];
    float *v = smem;                                                // [EH][EW] signed state
    uint32_t *bm = reinterpret_cast<uint32_t *>(v + EH * EW);       // [EH][BW] candidate bitmap
    uint16_t *list = reinterpret_cast<uint16_t *>(bm + EH * BW);    // undecided positions (ping)
    uint16_t *list2 = list + EH * EW;                               // (pong)
    __shared__ int n_list;
    __shared__ int n_next[3];
    __shared__ int warp_sums[NMS_THREADS / 32];
    __shared__ int bases[2];

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z;
    const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
    const int gy0 = ty0 - E, gx0 = tx0 - E;
    const float *img = prob + (size_t)b * H * W;
    const int R = fp.R;
    if (tid == 0) { n_list = 0; n_next[0] = n_next[1] = n_next[2] = 0; }
    for (int i = tid; i < EH * BW; i += NMS_THREADS) bm[i] = 0;
    __syncthreads();

    // ---- 1. stage the tile + apron; threshold; encode: 0 = nothing, -s = undecided ----
    constexpr int QW = EW / 4;
    for (int i0 = 0; i0 < EH * QW; i0 += NMS_THREADS) {
        const int i = i0 + tid;
        const bool act = i < EH * QW;
        const int ey = act ? i / QW : 0, q = act ? i - ey * QW : 0;
        const int gy = gy0 + ey, gx = gx0 + 4 * q;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act && gy >= 0 && gy < H) {
            if (VEC) {
                if (gx >= 0 && gx < W) val = ld_stream_f4(reinterpret_cast<const float4 *>(img + (size_t)gy * W + gx));
            } else {
                const float *rowp = img + (size_t)gy * W;
                if (gx + 0 >= 0 && gx + 0 < W) val.x = rowp[gx + 0];
                if (gx + 1 >= 0 && gx + 1 < W) val.y = rowp[gx + 1];
                if (gx + 2 >= 0 && gx + 2 < W) val.z = rowp[gx + 2];
                if (gx + 3 >= 0 && gx + 3 < W) val.w = rowp[gx + 3];
            }
        }
        float c[4] = {val.x, val.y, val.z, val.w};
        const bool rows_ok = act && ey >= R && ey < EH - R;
        uint32_t nib = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool cand = c[j] > thr;  // strict, fp32 (utils.py:97); NaN is not a candidate
            c[j] = cand ? -c[j] : 0.f;
            nib |= (uint32_t)cand << j;
            const int ex = 4 * q + j;
            const bool push = cand && rows_ok && ex >= R && ex < EW - R;
            const unsigned m = __ballot_sync(0xffffffffu, push);
            if (m) {
                int base = 0;
                if (lane == (__ffs(m) - 1)) base = atomicAdd(&n_list, __popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (push) list[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)(ey * EW + ex);
            }
        }
        if (act) {
            *reinterpret_cast<float4 *>(v + ey * EW + 4 * q) = make_float4(c[0], c[1], c[2], c[3]);
            if (nib) atomicOr(&bm[ey * BW + (q >> 3)], nib << ((4 * q) & 31));
        }
    }
    __syncthreads();

    // ---- 2. iterate to the local fixed point ----
    // Every round re-packs the still-undecided pixels into the other list so warps stay dense
    // (after two rounds only ~15 % are left).  Counters rotate over three slots: slot (r+1)%3 is
    // cleared at the top of round r, when every thread has long since read it (round r-2).
    {
        int n = n_list;
        uint16_t *cur = list, *nxt = list2;
        for (int round = 0; n > 0; ++round) {
            int *cnt = &n_next[round % 3];
            if (tid == 0) n_next[(round + 1) % 3] = 0;
            bool changed = false;
            for (int i0 = 0; i0 < n; i0 += NMS_THREADS) {
                const int i = i0 + tid;
                bool still = false;
                int e = 0;
                if (i < n) {
                    e = cur[i];
                    const float val = v[e];
                    const int ey = e / EW, ex = e - ey * EW;
                    const float nv = nms_decide_tile<EW, BW>(val, ey, ex, v, bm, fp);
                    if (nv != val) {
                        v[e] = nv;
                        changed = true;
                    } else {
                        still = true;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, still);
                if (m) {
                    int base = 0;
                    if (lane == (__ffs(m) - 1)) base = atomicAdd(cnt, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                    if (still) nxt[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)e;
                }
            }
            const bool any = __syncthreads_or(changed);
            n = *cnt;
            uint16_t *tmp = cur; cur = nxt; nxt = tmp;
            if (!any) break;  // what is left depends on pixels outside the apron
        }
    }

    // ---- 3. write the interior once; queue survivors and unresolved pixels ----
    constexpr int IQ = TW / 4;
    constexpr int PER_THREAD = (TH * IQ + NMS_THREADS - 1) / NMS_THREADS;
    int kept = 0, unres = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int i = tid + k * NMS_THREADS;
        if (i >= TH * IQ) break;
        const int iy = i / IQ, q = i - iy * IQ;
        const int gy = ty0 + iy, gx = tx0 + 4 * q;
        if (gy >= H || gx >= W) continue;
        const float4 val = *reinterpret_cast<const float4 *>(v + (E + iy) * EW + E + 4 * q);
        const float c[4] = {val.x, val.y, val.z, val.w};
        if (VEC) {
            st_stream_f4(reinterpret_cast<float4 *>(out + ((size_t)b * H + gy) * W + gx), val);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (gx + j < W) out[((size_t)b * H + gy) * W + gx + j] = c[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (gx + j >= W) continue;
            kept += c[j] > 0.f;
            unres += c[j] < 0.f;
        }
    }
    int tot_kept, tot_unres;
    int off_kept = block_exclusive_scan(kept, warp_sums, tot_kept);
    int off_unres = block_exclusive_scan(unres, warp_sums, tot_unres);
    if (tid == 0) {
        bases[0] = tot_kept ? atomicAdd(surv_count + b, tot_kept) : 0;
        bases[1] = tot_unres ? atomicAdd(work_count + b, tot_unres) : 0;
    }
    __syncthreads();
    if (tot_kept == 0 && tot_unres == 0) return;
    off_kept += bases[0];
    off_unres += bases[1];
    uint2 *surv = survivors + (size_t)b * cap;
    uint32_t *work = worklist + (size_t)b * cap;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int i = tid + k * NMS_THREADS;
        if (i >= TH * IQ) break;
        const int iy = i / IQ, q = i - iy * IQ;
        const int gy = ty0 + iy, gx = tx0 + 4 * q;
        if (gy >= H || gx >= W) continue;
        const float *c = v + (E + iy) * EW + E + 4 * q;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (gx + j >= W) continue;
            const uint32_t idx = (uint32_t)(gy * W + gx + j);
            if (c[j] > 0.f) surv[off_kept++] = make_uint2(idx, __float_as_uint(c[j]));
            else if (c[j] < 0.f) work[off_unres++] = idx;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fast tile kernel for footprints that fit a 7x7 window (reach <= 3: box sizes up to 4, the only
// ones the reference's configs use).  Same fixed point as nms_tile_kernel, restructured around
// what the first profile showed (76 k warp-instructions per tile, 19 of 32 lanes active):
//   - staging only thresholds and sets the candidate bitmap (one atomicOr per float4);
//   - the candidate list is generated from the bitmap (one word per thread);
//   - round 0 builds, once per candidate, the 49-bit mask of its higher-priority candidate
//     neighbours (7 funnel-shifted bitmap windows, then one loop over the set bits) -- local
//     maxima are kept right there;
//   - later rounds only revisit the cached mask bits (1-3 per pixel) and re-pack the ids of the
//     still-undecided pixels so warps stay dense.
// A tile with more than CAP candidates in its decidable region (> ~32 % density) skips the lists
// and sweeps its pixels instead (same decisions, no extra memory).
template <int TH, int TW, int E, bool VEC, int CAP>
__device__ __forceinline__ void
nms_tile_fast_body(const float *__restrict__ prob, float *__restrict__ out, int H, int W, float thr,
                   const NmsFootprint &fp, uint2 *__restrict__ survivors, int *__restrict__ surv_count,
                   uint32_t *__restrict__ worklist, int *__restrict__ work_count, int cap,
                   const int tile_x, const int tile_y, const int image) {
    constexpr int EH = TH + 2 * E, EW = TW + 2 * E, QW = EW / 4;
    constexpr int BW = (EW + 31) / 32 + 1;
    constexpr int RM = 3;  // list margin / window geometry
    static_assert(EW % 4 == 0 && E % 4 == 0 && TW % 4 == 0 && E >= RM, "tile geometry");
    static_assert(EH * EW < 65536, "positions are uint16");
    extern __shared__ __align__(16) float smem[];
    float *v = smem;                                                   // [EH][EW] signed state
    uint64_t *mask = reinterpret_cast<uint64_t *>(v + EH * EW);        // [CAP] higher-priority neighbours
    uint32_t *bm = reinterpret_cast<uint32_t *>(mask + CAP);           // [EH][BW] candidate bitmap
    uint16_t *pos = reinterpret_cast<uint16_t *>(bm + EH * BW);        // [CAP] position of candidate id
    uint16_t *ids_a = pos + CAP, *ids_b = ids_a + CAP;                 // undecided ids, ping / pong
    uint16_t *kept_pos = ids_b + CAP;                                  // interior pixels decided "kept"
    __shared__ int n_list, n_keptpos;
    __shared__ int n_next[3];
    __shared__ int bases[2];
    __shared__ uint32_t fp7[7];

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = image;
    const int ty0 = tile_y * TH, tx0 = tile_x * TW;
    const int gy0 = ty0 - E, gx0 = tx0 - E;
    const float *img = prob + (size_t)b * H * W;
    // ---- 1. stage the tile + apron.  All of a thread's global loads are issued first (one round trip instead
    //         of one per loop iteration: with ~20 waves of short-lived CTAs that serial latency was most of the
    //         kernel's fixed cost), the shared-memory setup overlaps them.
    //         Thread layout for the staging: column quad q = tid % QW fixed, rows r, r + RP, r + 2 RP, ... so the
    //         index arithmetic is done once and each pass only moves down RP rows.
    constexpr int RP = NMS_THREADS / QW;                 // rows per pass
    constexpr int LD_ITERS = (EH + RP - 1) / RP;
    const int sq = tid % QW, sr = tid / QW;
    const int sgx = gx0 + 4 * sq;
    const bool col_ok = sr < RP && (VEC ? (sgx >= 0 && sgx < W) : (sgx + 3 >= 0 && sgx < W));
    float4 stage[LD_ITERS];
#pragma unroll
    for (int k = 0; k < LD_ITERS; ++k) {
        const int ey = sr + k * RP;
        const int gy = gy0 + ey;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && ey < EH && gy >= 0 && gy < H) {
            const float *rowp = img + (size_t)gy * W;
            if (VEC) {
                val = ld_stream_f4(reinterpret_cast<const float4 *>(rowp + sgx));
            } else {
                if (sgx + 0 >= 0 && sgx + 0 < W) val.x = rowp[sgx + 0];
                if (sgx + 1 >= 0 && sgx + 1 < W) val.y = rowp[sgx + 1];
                if (sgx + 2 >= 0 && sgx + 2 < W) val.z = rowp[sgx + 2];
                if (sgx + 3 >= 0 && sgx + 3 < W) val.w = rowp[sgx + 3];
            }
        }
        stage[k] = val;
    }
    if (tid == 0) { n_list = 0; n_keptpos = 0; n_next[0] = n_next[1] = n_next[2] = 0; }
    if (tid < 7) {  // footprint rows re-centred in a 7-wide window
        const int dy = tid - RM;
        fp7[tid] = (dy >= -fp.R && dy <= fp.R) ? (fp.rows[dy + fp.R] << (RM - fp.R)) : 0u;
    }
    for (int i = tid; i < EH * BW; i += NMS_THREADS) bm[i] = 0;
    __syncthreads();

    // threshold, encode (0 nothing, -s undecided), set bitmap
    if (sr < RP) {
        float *vdst = v + sr * EW + 4 * sq;
        uint32_t *bdst = bm + sr * BW + (sq >> 3);
        const int bsh = (4 * sq) & 31;
#pragma unroll
        for (int k = 0; k < LD_ITERS; ++k) {
            if (sr + k * RP >= EH) break;
            float4 val = stage[k];
            const uint32_t nib = (uint32_t)(val.x > thr) | ((uint32_t)(val.y > thr) << 1) | ((uint32_t)(val.z > thr) << 2) |
                                 ((uint32_t)(val.w > thr) << 3);  // strict, fp32 (utils.py:97); NaN is not a candidate
            val.x = (nib & 1) ? -val.x : 0.f; val.y = (nib & 2) ? -val.y : 0.f;
            val.z = (nib & 4) ? -val.z : 0.f; val.w = (nib & 8) ? -val.w : 0.f;
            *reinterpret_cast<float4 *>(vdst + k * RP * EW) = val;
            if (nib) atomicOr(bdst + k * RP * BW, nib << bsh);
        }
    }
    __syncthreads();

    // ---- 2. candidate list of the decidable region [RM, EH-RM) x [RM, EW-RM) from the bitmap ----
    for (int w0 = 0; w0 < EH * BW; w0 += NMS_THREADS) {
        const int wi = w0 + tid;
        uint32_t bits = 0;
        int ey = 0, x0 = 0;
        if (wi < EH * BW) {
            ey = wi / BW;
            x0 = (wi - ey * BW) * 32;
            if (ey >= RM && ey < EH - RM) {
                bits = bm[wi];
                // keep columns [RM, EW-RM)
                if (x0 < RM) bits &= ~((1u << (RM - x0)) - 1u);
                const int hi = EW - RM - x0;  // first excluded bit
                if (hi <= 0) bits = 0;
                else if (hi < 32) bits &= (1u << hi) - 1u;
            }
        }
        const int cnt = __popc(bits);
        int inc = cnt;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, sft);
            if (lane >= sft) inc += t;
        }
        int base = 0;
        if (lane == 31 && inc) base = atomicAdd(&n_list, inc);
        base = __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
        while (bits) {
            const int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            if (base < CAP) pos[base] = (uint16_t)(ey * EW + x0 + bpos);
            ++base;
        }
    }
    __syncthreads();
    const int n0 = n_list;

    const int warp = tid >> 5;
    auto note_kept = [&](int e) {  // remember interior survivors as they are decided: no rescan at the end
        const int ey = e / EW, ex = e - ey * EW;
        if (ey >= E && ey < E + TH && ex >= E && ex < E + TW && ty0 + ey - E < H && tx0 + ex - E < W)
            kept_pos[atomicAdd(&n_keptpos, 1)] = (uint16_t)e;
    };
    int n_left = 0;               // undecided ids still listed when the rounds stop
    const uint16_t *left = ids_a;
    if (n0 <= CAP) {
        // ---- 3a. round 0: higher-priority candidate neighbours of every candidate, once ----
        // mask bit 8*r + c  <=>  neighbour at (dy, dx) = (r - 3, c - 3); two 32-bit words (rows 0-3, 4-6)
        for (int base = warp * 32; base < n0; base += NMS_THREADS) {  // warp-strided: idle warps skip
            const int id = base + lane;
            bool still = false;
            if (id < n0) {
                const int e = pos[id];
                const int ey = e / EW, ex = e - ey * EW;
                const float s = -v[e];
                const int bitpos = ex - RM, w = bitpos >> 5, sh = bitpos & 31;
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    const uint32_t *bw = bm + (ey + r - RM) * BW + w;
                    const uint32_t win = __funnelshift_r(bw[0], bw[1], sh) & fp7[r];
                    if (r < 4) lo |= win << (8 * r); else hi |= win << (8 * (r - 4));
                }
                const float *vb = v + e - RM * EW - RM;
                // positive floats order like their bit patterns; neighbours earlier in row-major order
                // (bits 0..26 of lo) win ties (>=), later ones (bits 28..30 of lo, all of hi) need >
                const uint32_t sb = __float_as_uint(s);
                uint32_t hlo = 0, hhi = 0;
                for (uint32_t m = lo & 0x07ffffffu; m;) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t nb_ = __float_as_uint(vb[(k >> 3) * EW + (k & 7)]) & 0x7fffffffu;  // every candidate still carries its score
                    if (nb_ >= sb) hlo |= 1u << k;
                }
                for (uint32_t m = lo & 0x70000000u; m;) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t nb_ = __float_as_uint(vb[3 * EW + (k & 7)]) & 0x7fffffffu;
                    if (nb_ > sb) hlo |= 1u << k;
                }
                for (uint32_t m = hi; m;) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t nb_ = __float_as_uint(vb[(4 + (k >> 3)) * EW + (k & 7)]) & 0x7fffffffu;
                    if (nb_ > sb) hhi |= 1u << k;
                }
                if ((hlo | hhi) == 0) {
                    v[e] = s;  // local maximum: kept
                    if (ey >= E && ey < E + TH && ex >= E && ex < E + TW && ty0 + ey - E < H && tx0 + ex - E < W)
                        kept_pos[atomicAdd(&n_keptpos, 1)] = (uint16_t)e;
                } else {
                    mask[id] = ((uint64_t)hhi << 32) | hlo;
                    still = true;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, still);
            if (bal) {
                int slot = 0;
                if (lane == (__ffs(bal) - 1)) slot = atomicAdd(&n_next[0], __popc(bal));
                slot = __shfl_sync(0xffffffffu, slot, __ffs(bal) - 1);
                if (still) ids_a[slot + __popc(bal & ((1u << lane) - 1))] = (uint16_t)id;
            }
        }
        __syncthreads();
        // ---- 3b. rounds over the cached masks ----
        int n = n_next[0];
        uint16_t *cur = ids_a, *nxt = ids_b;
        n_left = n; left = cur;
        for (int round = 1; n > 0; ++round) {
            int *cnt = &n_next[round % 3];
            if (tid == 0) n_next[(round + 1) % 3] = 0;
            bool changed = false;
            for (int base = warp * 32; base < n; base += NMS_THREADS) {
                const int i = base + lane;
                bool still = false;
                int id = 0;
                if (i < n) {
                    id = cur[i];
                    const int e = pos[id];
                    const float *vb = v + e - RM * EW - RM;
                    const uint64_t old = mask[id];
                    uint32_t hlo = (uint32_t)old, hhi = (uint32_t)(old >> 32);
                    bool sup = false;
                    for (uint32_t m = hlo; m && !sup;) {
                        const int k = __ffs(m) - 1;
                        m &= m - 1;
                        const float nv = vb[(k >> 3) * EW + (k & 7)];
                        if (nv > 0.f) sup = true;                  // kept higher-priority neighbour
                        else if (nv == 0.f) hlo &= ~(1u << k);     // it was suppressed: no longer blocks
                    }
                    for (uint32_t m = hhi; m && !sup;) {
                        const int k = __ffs(m) - 1;
                        m &= m - 1;
                        const float nv = vb[(4 + (k >> 3)) * EW + (k & 7)];
                        if (nv > 0.f) sup = true;
                        else if (nv == 0.f) hhi &= ~(1u << k);
                    }
                    if (sup) { v[e] = 0.f; changed = true; }
                    else if ((hlo | hhi) == 0) { v[e] = -v[e]; changed = true; note_kept(e); }
                    else {
                        const uint64_t nm = ((uint64_t)hhi << 32) | hlo;
                        if (nm != old) mask[id] = nm;
                        still = true;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, still);
                if (bal) {
                    int slot = 0;
                    if (lane == (__ffs(bal) - 1)) slot = atomicAdd(cnt, __popc(bal));
                    slot = __shfl_sync(0xffffffffu, slot, __ffs(bal) - 1);
                    if (still) nxt[slot + __popc(bal & ((1u << lane) - 1))] = (uint16_t)id;
                }
            }
            const bool any = __syncthreads_or(changed);
            n = *cnt;
            uint16_t *tmp = cur; cur = nxt; nxt = tmp;
            n_left = n; left = cur;
            if (!any) break;  // what is left depends on pixels outside the apron
            if (n <= 32) {
                // tail: the last few pixels settle in one warp, round after round, without block-wide
                // barriers (each costs more than the round itself once only a handful are left)
                if (warp == 0) {
                    int id = lane < n ? cur[lane] : -1;
                    while (__any_sync(0xffffffffu, id >= 0)) {
                        bool progressed = false;
                        if (id >= 0) {
                            const int e = pos[id];
                            const float *vb = v + e - RM * EW - RM;
                            const uint64_t old = mask[id];
                            uint32_t hlo = (uint32_t)old, hhi = (uint32_t)(old >> 32);
                            bool sup = false;
                            for (uint32_t m = hlo; m && !sup;) {
                                const int k = __ffs(m) - 1;
                                m &= m - 1;
                                const float nv = vb[(k >> 3) * EW + (k & 7)];
                                if (nv > 0.f) sup = true;
                                else if (nv == 0.f) hlo &= ~(1u << k);
                            }
                            for (uint32_t m = hhi; m && !sup;) {
                                const int k = __ffs(m) - 1;
                                m &= m - 1;
                                const float nv = vb[(4 + (k >> 3)) * EW + (k & 7)];
                                if (nv > 0.f) sup = true;
                                else if (nv == 0.f) hhi &= ~(1u << k);
                            }
                            if (sup) { v[e] = 0.f; id = -1; progressed = true; }
                            else if ((hlo | hhi) == 0) { v[e] = -v[e]; id = -1; progressed = true; note_kept(e); }
                            else mask[id] = ((uint64_t)hhi << 32) | hlo;
                        }
                        __syncwarp();
                        if (!__any_sync(0xffffffffu, progressed)) break;
                    }
                }
                break;
            }
        }
    } else {
        // ---- 3c. dense tile: sweep the pixels of the decidable region until nothing changes ----
        constexpr int DW = EW - 2 * RM, DH = EH - 2 * RM;
        while (true) {
            bool changed = false;
            for (int i = tid; i < DW * DH; i += NMS_THREADS) {
                const int ey = RM + i / DW, ex = RM + i % DW;
                const float val = v[ey * EW + ex];
                if (val >= 0.f) continue;
                const float nv = nms_decide_tile<EW, BW>(val, ey, ex, v, bm, fp);
                if (nv != val) { v[ey * EW + ex] = nv; changed = true; }
            }
            if (!__syncthreads_or(changed)) break;
        }
    }
    __syncthreads();

    constexpr int IQ = TW / 4;
    constexpr int PER_THREAD = (TH * IQ + NMS_THREADS - 1) / NMS_THREADS;
    if (n0 <= CAP) {
        // ---- 4a. list path: plain copy of the interior; survivors come from kept_pos, unresolved
        //          pixels from what is left of the id list (still-undecided entries inside the tile) ----
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            const int i = tid + k * NMS_THREADS;
            const int iy = i / IQ, q = i - iy * IQ;
            const int gy = ty0 + iy, gx = tx0 + 4 * q;
            if (i < TH * IQ && gy < H && gx < W) {
                const float4 val = *reinterpret_cast<const float4 *>(v + (E + iy) * EW + E + 4 * q);
                if (VEC) {
                    st_stream_f4(reinterpret_cast<float4 *>(out + ((size_t)b * H + gy) * W + gx), val);
                } else {
                    const float c[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (gx + j < W) out[((size_t)b * H + gy) * W + gx + j] = c[j];
                }
            }
        }
        // unresolved interior pixels: compact their global indices over the (now free) mask array
        uint32_t *stg_un = reinterpret_cast<uint32_t *>(mask);
        if (tid == 0) n_next[1] = 0;
        __syncthreads();
        for (int base = warp * 32; base < n_left; base += NMS_THREADS) {
            const int i = base + lane;
            uint32_t gidx = 0;
            bool un = false;
            if (i < n_left) {
                const int e = pos[left[i]];
                const int ey = e / EW, ex = e - ey * EW;
                const int gy = ty0 + ey - E, gx = tx0 + ex - E;
                un = v[e] < 0.f && ey >= E && ey < E + TH && ex >= E && ex < E + TW && gy < H && gx < W;
                gidx = (uint32_t)(gy * W + gx);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, un);
            if (bal) {
                int slot = 0;
                if (lane == (__ffs(bal) - 1)) slot = atomicAdd(&n_next[1], __popc(bal));
                slot = __shfl_sync(0xffffffffu, slot, __ffs(bal) - 1);
                if (un) stg_un[slot + __popc(bal & ((1u << lane) - 1))] = gidx;
            }
        }
        __syncthreads();
        const int tk = n_keptpos, tu = n_next[1];
        if (tid == 0) {
            bases[0] = tk ? atomicAdd(surv_count + b, tk) : 0;
            bases[1] = tu ? atomicAdd(work_count + b, tu) : 0;
        }
        __syncthreads();
        uint2 *surv = survivors + (size_t)b * cap + bases[0];
        uint32_t *work = worklist + (size_t)b * cap + bases[1];
        for (int i = tid; i < tk; i += NMS_THREADS) {
            const int e = kept_pos[i];
            const int ey = e / EW, ex = e - ey * EW;
            surv[i] = make_uint2((uint32_t)((ty0 + ey - E) * W + tx0 + ex - E), __float_as_uint(v[e]));
        }
        for (int i = tid; i < tu; i += NMS_THREADS) work[i] = stg_un[i];
        return;
    }

    // ---- 4b. dense path: write the interior once; survivors / unresolved pixels are staged in shared
    //          memory (the mask and id arrays are unused here) and leave as two contiguous runs ----
    constexpr int KCAP = CAP, UCAP = (3 * CAP) / 2;  // uint2 over mask[], uint32 over pos/ids
    uint2 *stg_kept = reinterpret_cast<uint2 *>(mask);
    uint32_t *stg_un = reinterpret_cast<uint32_t *>(pos);
    int *n_kept = &n_next[0], *n_un = &n_next[1];
    if (tid == 0) { n_next[0] = 0; n_next[1] = 0; }
    __syncthreads();
    uint2 *surv = survivors + (size_t)b * cap;
    uint32_t *work = worklist + (size_t)b * cap;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int i = tid + k * NMS_THREADS;  // TH*IQ is a multiple of NMS_THREADS for the shipped tile: no ragged warp
        const int iy = i / IQ, q = i - iy * IQ;
        const int gy = ty0 + iy, gx = tx0 + 4 * q;
        const bool inside = i < TH * IQ && gy < H && gx < W;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        if (inside) {
            const float4 val = *reinterpret_cast<const float4 *>(v + (E + iy) * EW + E + 4 * q);
            c[0] = val.x; c[1] = val.y; c[2] = val.z; c[3] = val.w;
            if (VEC) {
                st_stream_f4(reinterpret_cast<float4 *>(out + ((size_t)b * H + gy) * W + gx), val);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gx + j < W) out[((size_t)b * H + gy) * W + gx + j] = c[j];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gx + j >= W) c[j] = 0.f;
            }
        }
        const int nk = (c[0] > 0.f) + (c[1] > 0.f) + (c[2] > 0.f) + (c[3] > 0.f);
        const int nu = (c[0] < 0.f) + (c[1] < 0.f) + (c[2] < 0.f) + (c[3] < 0.f);
        if (__any_sync(0xffffffffu, (nk | nu) != 0)) {
            int ik = nk, iu = nu;  // inclusive warp scans
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) {
                const int tk = __shfl_up_sync(0xffffffffu, ik, sft), tu = __shfl_up_sync(0xffffffffu, iu, sft);
                if (lane >= sft) { ik += tk; iu += tu; }
            }
            int bk = 0, bu = 0;
            if (lane == 31) {
                if (ik) bk = atomicAdd(n_kept, ik);
                if (iu) bu = atomicAdd(n_un, iu);
            }
            bk = __shfl_sync(0xffffffffu, bk, 31) + ik - nk;
            bu = __shfl_sync(0xffffffffu, bu, 31) + iu - nu;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t idx = (uint32_t)(gy * W + gx + j);
                if (c[j] > 0.f) {
                    const uint2 ent = make_uint2(idx, __float_as_uint(c[j]));
                    if (bk < KCAP) stg_kept[bk] = ent;
                    else surv[atomicAdd(surv_count + b, 1)] = ent;   // staging full (tiny boxes): rare
                    ++bk;
                } else if (c[j] < 0.f) {
                    if (bu < UCAP) stg_un[bu] = idx;
                    else work[atomicAdd(work_count + b, 1)] = idx;
                    ++bu;
                }
            }
        }
    }
    __syncthreads();
    const int tk = min(*n_kept, KCAP), tu = min(*n_un, UCAP);
    if (tid == 0) {
        bases[0] = tk ? atomicAdd(surv_count + b, tk) : 0;
        bases[1] = tu ? atomicAdd(work_count + b, tu) : 0;
    }
    __syncthreads();
    for (int i = tid; i < tk; i += NMS_THREADS) surv[bases[0] + i] = stg_kept[i];
    for (int i = tid; i < tu; i += NMS_THREADS) work[bases[1] + i] = stg_un[i];
}

template <int TH, int TW, int E, bool VEC, int CAP>
__global__ void __launch_bounds__(NMS_THREADS, (TW <= 64 ? 7 : 3))  // 32x64 tiles: 7 CTAs/SM fit in shared memory, keep <= 36 registers
nms_tile_fast_kernel(const float *__restrict__ prob, float *__restrict__ out, int H, int W, float thr,
                     const NmsFootprint fp, uint2 *__restrict__ survivors, int *__restrict__ surv_count,
                     uint32_t *__restrict__ worklist, int *__restrict__ work_count, int cap) {
    nms_tile_fast_body<TH, TW, E, VEC, CAP>(prob, out, H, W, thr, fp, survivors, surv_count, worklist, work_count, cap,
                                            blockIdx.x, blockIdx.y, blockIdx.z);
}

// The same tiles for the images the sparse top-k path flagged, as a persistent grid: when no image is
// flagged (the usual case) each CTA only reads the flags and leaves -- a full grid of 20 k CTAs that
// exit at once still costs 17 us.
template <int TH, int TW, int E, bool VEC, int CAP>
__global__ void __launch_bounds__(NMS_THREADS, (TW <= 64 ? 7 : 3))
nms_tile_fast_redo_kernel(const float *__restrict__ prob, float *__restrict__ out, int H, int W, float thr,
                          const NmsFootprint fp, uint2 *__restrict__ survivors, int *__restrict__ surv_count,
                          uint32_t *__restrict__ worklist, int *__restrict__ work_count, int cap,
                          const int *__restrict__ flagged, int tiles_x, int tiles_y, int B) {
    const int per = tiles_x * tiles_y;
    bool mine = false;
    for (int i = threadIdx.x; i < B; i += NMS_THREADS) mine |= flagged[i] != 0;
    if (!__syncthreads_or(mine)) return;  // nothing to redo: one load per thread and out
    for (int t = blockIdx.x; t < per * B; t += gridDim.x) {
        const int image = t / per;
        if (flagged[image] == 0) continue;  // block-uniform
        const int r = t - image * per;
        nms_tile_fast_body<TH, TW, E, VEC, CAP>(prob, out, H, W, thr, fp, survivors, surv_count, worklist, work_count, cap,
                                                r % tiles_x, r / tiles_x, image);
        __syncthreads();  // shared memory is reused by the next tile
    }
}

// One CTA per image: resolve the pixels whose dependency chain left their tile's apron.
// One warp per queued pixel: the lanes fetch the (2R+1)^2 window from the dense map (L2) in
// parallel and vote, so a round costs one L2 round trip instead of one per neighbour.
__device__ __forceinline__ void
nms_fixup_body(float *__restrict__ out, int H, int W, const NmsFootprint &fp, uint2 *survivors,
               int *surv_count, const uint32_t *__restrict__ worklist,
               const int *__restrict__ work_count, int cap) {
    const int b = blockIdx.x;
    const int n = work_count[b];  // 0 for images the sparse top-k path settled
    if (n == 0) return;
    float *img = out + (size_t)b * H * W;
    const uint32_t *work = worklist + (size_t)b * cap;
    uint2 *surv = survivors + (size_t)b * cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int R = fp.R, S = 2 * R + 1;
    while (true) {
        bool changed = false, pending = false;
        for (int i = warp; i < n; i += nwarps) {
            const int idx = (int)work[i];
            const float val = __ldcg(img + idx);
            if (val >= 0.f) continue;  // warp-uniform
            const float s = -val;
            const int y = idx / W, x = idx - y * W;
            bool kept_higher = false, und_higher = false;
            for (int k = lane; k < S * S; k += 32) {
                const int dy = k / S - R, dx = k - (dy + R) * S - R;
                if (!((fp.rows[dy + R] >> (dx + R)) & 1u)) continue;
                const int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                const float nv = __ldcg(img + yy * W + xx);
                if (nv == 0.f) continue;
                const float sn = fabsf(nv);
                if (sn > s || (sn == s && (dy < 0 || (dy == 0 && dx < 0)))) {
                    if (nv > 0.f) kept_higher = true; else und_higher = true;
                }
            }
            kept_higher = __any_sync(0xffffffffu, kept_higher);
            und_higher = __any_sync(0xffffffffu, und_higher);
            if (kept_higher || !und_higher) {
                const float nv = kept_higher ? 0.f : s;
                if (lane == 0) {
                    __stcg(img + idx, nv);
                    if (nv > 0.f) surv[atomicAdd(surv_count + b, 1)] = make_uint2((uint32_t)idx, __float_as_uint(nv));
                }
                changed = true;
            } else {
                pending = true;
            }
        }
        __threadfence_block();
        const bool any_pending = __syncthreads_or(pending);
        const bool any_changed = __syncthreads_or(changed);
        if (!any_pending || !any_changed) break;  // progress is guaranteed while anything is pending
    }
}

// ------------------------------------------------------------------------------------------
// ordered emission from a per-image bitmap: bit i of word w <=> pixel 32*w+i is a keypoint.
// (no __restrict__ / read-only loads here: the bitmap and the score map were written earlier in
// the same kernel by other threads of this CTA)
__device__ void emit_from_bitmap(const uint32_t *bitmap, int words, int W, const float *score_src,
                                 int64_t *__restrict__ kp, float *__restrict__ kp_scores, int32_t *__restrict__ kp_count,
                                 int kp_cap, int *warp_sums) {
    const int per = (words + blockDim.x - 1) / blockDim.x;
    const int w0 = min(words, (int)threadIdx.x * per), w1 = min(words, w0 + per);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(__ldcg(bitmap + w));
    int total;
    int off = block_exclusive_scan(cnt, warp_sums, total);
    if (threadIdx.x == 0 && kp_count) *kp_count = total;
    if (kp == nullptr) return;
    for (int w = w0; w < w1; ++w) {
        uint32_t bits = __ldcg(bitmap + w);
        while (bits) {
            const int i = __ffs(bits) - 1;
            bits &= bits - 1;
            if (off < kp_cap) {
                const int idx = 32 * w + i;
                kp[2 * (size_t)off] = idx / W;
                kp[2 * (size_t)off + 1] = idx % W;
                if (kp_scores) kp_scores[off] = __ldcg(score_src + idx);
            }
            ++off;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Sparse top-k path (keep_top_k > 0, footprint reach <= 3).
// box_nms(..., keep_top_k=k) returns the k best survivors of greedy NMS.  Whether a pixel survives
// depends only on higher-priority pixels, so the k best survivors are already determined by the
// candidates that score at least as much as the k-th survivor: 2.2 k of 42 k candidates on the
// benchmark's heatmaps (5 %; measured with the oracle).  Instead of settling every candidate,
//  1. nms_candidates_kernel streams the heatmap once (HBM-bound): zero-fills the dense output,
//     appends every candidate (pixel, score bits) to a per-image list and histograms the scores
//     into 128 monotone bins (exponent + 4 mantissa bits),
//  2. nms_sparse2_kernel (one CTA per image) picks the bin threshold that admits ~1.5 k
//     candidates, settles exactly those with the same fixed point as the tile kernel -- whole image
//     in one CTA, so no apron and no fix-up; candidate bitmap, ranks and signed state all in shared
//     memory -- lowers the threshold and repeats while fewer than k survive, then cuts to k and
//     emits the keypoints in row-major order itself (the select step of nms_fixup_select_kernel only serves redone images).
// An image that does not fit (a chunk with more than SP_SEG_CAP candidates, or more than SP2_CAP needed)
// raises its flag and is redone by the tile + fix-up kernels, which otherwise exit at once.
constexpr int SP_BINS = 128;
constexpr int SP_SEG_CAP = 1536;      // candidates listed per SP_CHUNK pixels (37.5 %); a denser chunk sends the image to the tile kernels
constexpr int SP_CHUNK = 4096;        // pixels per CTA of the candidates kernel
constexpr int SP_PAD = 3;             // footprint reach handled here

__device__ __forceinline__ int sp_bin(uint32_t bits) {  // monotone in the (positive) score
    return bits <= 0x3C000000u ? 0 : min(SP_BINS - 1, (int)((bits - 0x3C000000u) >> 19));
}

template <bool VEC, bool ZERO>   // ZERO: also zero-fill the dense output map (only when the caller asked for it)
__global__ void __launch_bounds__(NMS_THREADS)  // 32 registers, 21.5 KB shared: 8 CTAs/SM
nms_candidates_kernel(const float *__restrict__ prob, float *__restrict__ out, int W, int HW, float thr,
                      uint2 *__restrict__ cands, int *__restrict__ seg_count, int *__restrict__ hist) {
    // One histogram per warp: the scores cluster in a few bins just above the threshold, and a single shared histogram
    // serialised its atomics there (ncu: 465 k bank conflicts per launch).  Every CTA owns a fixed segment of the
    // image's candidate list and stores its count: reserving a run of one shared list with a returning atomic put an L2
    // round trip into every CTA's critical path (per-warp reservations were even measured at 109 us against 65: the L2
    // serialises the 640 atomics an image then sends to one address).  List entry: (pixel index, score bits) --
    // splitting the index into (y, x) here costs a division per candidate (86 us against 65); the sparse kernel does it
    // for the admitted 8 %.
    __shared__ int sh_hist[NMS_THREADS / 32][SP_BINS];
    __shared__ int warp_sums[NMS_THREADS / 32];
    __shared__ float sh_val[SP_CHUNK];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int p0 = blockIdx.x * SP_CHUNK;
    const float *img = prob + (size_t)b * HW;
    float *dst = ZERO ? out + (size_t)b * HW : nullptr;
    for (int i = tid; i < (NMS_THREADS / 32) * SP_BINS; i += NMS_THREADS) (&sh_hist[0][0])[i] = 0;
    constexpr int PER = SP_CHUNK / NMS_THREADS;  // 16 pixels per thread
    float v[PER];
    if (VEC) {
#pragma unroll
        for (int k = 0; k < PER / 4; ++k) {
            const int i = p0 + 4 * (tid + k * NMS_THREADS);
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < HW) {  // HW % 4 == 0 on this path
                q = ld_stream_f4(reinterpret_cast<const float4 *>(img + i));
                if (ZERO) st_stream_f4(reinterpret_cast<float4 *>(dst + i), make_float4(0.f, 0.f, 0.f, 0.f));
            }
            v[4 * k + 0] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = p0 + tid + k * NMS_THREADS;
            v[k] = 0.f;
            if (i < HW) { v[k] = img[i]; if (ZERO) dst[i] = 0.f; }
        }
    }
    // Per-thread candidate masks, then a loop over each thread's own candidates.  A pixel slot of a warp almost always
    // holds at least one candidate (12.9 % of the pixels on the benchmark maps), so code that walks the 16 slots and
    // predicates the per-candidate work executes all of it 16 times per thread (first version: 38 instructions per pixel,
    // issue-bound at 35 % of DRAM; with one ballot per slot: ~25).  The loop below runs max-over-the-warp(candidates per
    // thread) ~ 5 times instead.  The values are parked in shared memory because the loop indexes them dynamically.
    uint32_t cm = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        cm |= (v[k] > thr ? 1u : 0u) << k;   // strict, fp32 (utils.py:97); NaN is not a candidate
        sh_val[k * NMS_THREADS + tid] = v[k];
    }
    const int lane = tid & 31, warp = tid >> 5;
    int *my_hist = sh_hist[warp];
    const int cnt = __popc(cm);
    int incl = cnt;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, sft);
        if (lane >= sft) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();   // also orders the histogram zero-fill before the atomics below
    int run = incl - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < NMS_THREADS / 32; ++w) {
        const int t = warp_sums[w];
        if (w < warp) run += t;
        total += t;
    }
    if (tid == 0) seg_count[(size_t)b * gridDim.x + blockIdx.x] = total;   // may exceed SP_SEG_CAP: the sparse kernel then gives the image up
    uint2 *list = cands + ((size_t)b * gridDim.x + blockIdx.x) * SP_SEG_CAP;
    while (cm) {
        const int k = __ffs(cm) - 1;
        cm &= cm - 1;
        const uint32_t bits = __float_as_uint(sh_val[k * NMS_THREADS + tid]);
        const int i = VEC ? p0 + 4 * (tid + (k >> 2) * NMS_THREADS) + (k & 3) : p0 + tid + k * NMS_THREADS;
        if (run < SP_SEG_CAP) list[run] = make_uint2((uint32_t)i, bits);
        ++run;
        atomicAdd(&my_hist[sp_bin(bits)], 1);
    }
    __syncthreads();
    if (tid < SP_BINS) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < NMS_THREADS / 32; ++w) t += sh_hist[w][tid];
        if (t) atomicAdd(hist + b * SP_BINS + tid, t);
    }
}



// One CTA per image: top-k (optional) + ordered compaction.
// Top-k = the k smallest keys (~score_bits, index), i.e. highest scores first, ties to the lower
// row-major index (the stable order of torchvision's sort).  MSB-first radix select with 11-bit
// digits: three passes over the score bits; the three index passes only run when equal scores
// straddle k.  Each thread keeps its (<= 16) survivors in registers, so the list is read once.
constexpr int SEL_BINS = 2048, SEL_CACHE = 16;

// Which bin holds the `remaining`-th entry when bins are visited in descending (DESC) or ascending
// order?  Thread t owns two adjacent bins; one block scan gives the count ahead of them.
template <bool DESC>
__device__ __forceinline__ void select_find_bin(const int *hist, int remaining, int *warp_sums, int *out_bin, int *out_remaining) {
    const int t = threadIdx.x;
    const int d_first = DESC ? SEL_BINS - 1 - 2 * t : 2 * t, d_second = DESC ? d_first - 1 : d_first + 1;
    const int h1 = hist[d_first], h2 = hist[d_second];
    int total;
    const int ahead = block_exclusive_scan(h1 + h2, warp_sums, total);
    if (ahead < remaining && remaining <= ahead + h1) { *out_bin = d_first; *out_remaining = remaining - ahead; }
    else if (ahead + h1 < remaining && remaining <= ahead + h1 + h2) { *out_bin = d_second; *out_remaining = remaining - ahead - h1; }
    __syncthreads();
}

// Sparse top-k NMS, one CTA per image, everything in shared memory (no dense state map, no survivor list, no
// separate select launch):
//   - the admitted candidates (score bin >= the histogram threshold) set bits in a padded candidate bitmap;
//   - a prefix count per bitmap word turns a position into a dense id (row-major rank), so state[id] / mask[id]
//     live in shared memory and a neighbour's state is two shared-memory reads and a popcount away -- the first version
//     kept the state in the dense output map and paid an L2 round trip per neighbour (51 us, latency-bound), and
//     needed the map zero-filled (168 MB per step);
//   - the same three-valued fixed point as the tile kernel (round 0 builds the higher-priority neighbour masks);
//   - top-k cut by radix select over the kept scores, ties to the lower row-major index, and -- ids being in
//     row-major order already -- ordered emission of the keypoints with one block scan.
// `dense` (optional) must be zero-filled (nms_candidates_kernel<.., true>): the kept scores are scattered into it.
// An image that does not fit raises redo_flags[b] and is redone by the tile + fix-up + select kernels.
constexpr int SP2_CAP = 7168;      // candidates settled per image
constexpr int SP2_PER = SP2_CAP / 1024;

struct Sp2Layout {
    int BWR, BROWS, words;
    size_t mask, bm, pos, st, wbase, ids_a, ids_b, total;
    __host__ __device__ Sp2Layout(int H, int W) {
        BWR = (W + 31) / 32 + 2;               // one pad word on each side
        BROWS = H + 2 * SP_PAD;
        words = BROWS * BWR;
        size_t off = 0;
        mask = off;  off += sizeof(uint64_t) * SP2_CAP;
        bm = off;    off += sizeof(uint32_t) * (size_t)((words + 1) & ~1);
        pos = off;   off += sizeof(uint32_t) * SP2_CAP;
        st = off;    off += sizeof(float) * SP2_CAP;
        wbase = off; off += sizeof(uint16_t) * (size_t)((words + 1) & ~1);
        ids_a = off; off += sizeof(uint16_t) * SP2_CAP;
        ids_b = off; off += sizeof(uint16_t) * SP2_CAP;
        total = off;
    }
};

__global__ void __launch_bounds__(1024, 1)
nms_sparse2_kernel(const float *__restrict__ prob, float *__restrict__ dense, int H, int W, int keep_top_k, const NmsFootprint fp,
                   const uint2 *__restrict__ cands, const int *__restrict__ seg_count, const int *__restrict__ hist,
                   int64_t *__restrict__ keypoints, float *__restrict__ kp_scores, int32_t *__restrict__ kp_counts, int kp_cap,
                   int *__restrict__ surv_count, int *__restrict__ redo_flags) {
    extern __shared__ __align__(16) uint8_t sp_smem[];
    const Sp2Layout L(H, W);
    const int BWR = L.BWR;
    uint64_t *mask = reinterpret_cast<uint64_t *>(sp_smem + L.mask);
    uint32_t *bm = reinterpret_cast<uint32_t *>(sp_smem + L.bm);
    uint32_t *pos = reinterpret_cast<uint32_t *>(sp_smem + L.pos);
    volatile float *st = reinterpret_cast<volatile float *>(sp_smem + L.st);   // -s undecided, +s kept, 0 suppressed
    uint16_t *wbase = reinterpret_cast<uint16_t *>(sp_smem + L.wbase);
    uint16_t *ids_a = reinterpret_cast<uint16_t *>(sp_smem + L.ids_a), *ids_b = reinterpret_cast<uint16_t *>(sp_smem + L.ids_b);
    __shared__ int sh_hist[SP_BINS];
    __shared__ int warp_sums[32];
    __shared__ int n_kept, sel_bin, sel_remaining;
    __shared__ int n_next[3];
    __shared__ uint32_t fp7[7];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const float *img = prob + (size_t)b * H * W;
    const int nseg = (H * W + SP_CHUNK - 1) / SP_CHUNK;
    const int *segc = seg_count + (size_t)b * nseg;
    const uint2 *list = cands + (size_t)b * nseg * SP_SEG_CAP;
    {
        bool over = false;
        for (int sg = tid; sg < nseg; sg += 1024) over |= segc[sg] > SP_SEG_CAP;
        if (__syncthreads_or(over)) {  // a chunk's list is truncated: the dense kernels redo this image (block-uniform)
            if (tid == 0) { redo_flags[b] = 1; surv_count[b] = 0; }
            return;
        }
    }
    if (tid < SP_BINS) sh_hist[tid] = hist[b * SP_BINS + tid];
    if (tid < 7) {
        const int dy = tid - SP_PAD;
        fp7[tid] = (dy >= -fp.R && dy <= fp.R) ? (fp.rows[dy + fp.R] << (SP_PAD - fp.R)) : 0u;
    }
    __syncthreads();
    // id of the candidate at padded row r, padded bit position xb (= x + 32)
    auto id_at = [&](int r, int xb) -> int {
        const int wi = r * BWR + (xb >> 5);
        return (int)wbase[wi] + __popc(bm[wi] & ((1u << (xb & 31)) - 1u));
    };

    int target = min(SP2_CAP, keep_top_k + keep_top_k / 2 + 256);
    int n0 = 0;
    const float inv_w = 1.0f / (float)W;
    while (true) {
        // threshold bin: the highest bin t with (number of candidates in bins >= t) >= target, or 0 (one warp searches:
        // four bins per lane, a suffix scan over the lanes, then the four bins of the lane that crosses the target)
        if (warp == 0) {
            const int h0 = sh_hist[4 * lane], h1 = sh_hist[4 * lane + 1], h2 = sh_hist[4 * lane + 2], h3 = sh_hist[4 * lane + 3];
            int suf = h0 + h1 + h2 + h3;      // inclusive suffix sum over lanes >= this one
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) {
                const int t = __shfl_down_sync(0xffffffffu, suf, sft);
                if (lane + sft < 32) suf += t;
            }
            const unsigned reach = __ballot_sync(0xffffffffu, suf >= target);
            int tbv = 0, nsel = __shfl_sync(0xffffffffu, suf, 0);   // not reached: every candidate
            if (reach) {
                const int l = 31 - __clz(reach);                   // highest lane whose suffix reaches the target
                if (lane == l) {
                    int cum = suf - (h0 + h1 + h2 + h3);           // candidates in the bins above this lane's
                    const int hh[4] = {h0, h1, h2, h3};
                    int t = 3;
                    for (; t >= 0; --t) { cum += hh[t]; if (cum >= target) break; }
                    sel_bin = 4 * l + t; sel_remaining = cum;
                }
                __syncwarp();
            } else if (lane == 0) { sel_bin = tbv; sel_remaining = nsel; }
        }
        for (int i = tid; i < L.words; i += 1024) bm[i] = 0;
        if (tid == 0) { n_kept = 0; n_next[0] = n_next[1] = n_next[2] = 0; }
        __syncthreads();
        const int tb = sel_bin, n_sel = sel_remaining;
        if (n_sel > SP2_CAP) {
            if (tid == 0) { redo_flags[b] = 1; surv_count[b] = 0; }
            return;
        }
        // ---- pass 1 over the candidate lists (a warp per chunk segment, sixteen loads per lane in flight: a 4096-pixel segment of the benchmark maps is one pass): set the bitmap bit
        //      of every admitted candidate
        constexpr int SCAN = 16;
        for (int sg = warp; sg < nseg; sg += 32) {
            const int n = segc[sg];
            const uint2 *seg = list + (size_t)sg * SP_SEG_CAP;
            for (int i0 = 0; i0 < n; i0 += 32 * SCAN) {
                uint2 c[SCAN];
#pragma unroll
                for (int u = 0; u < SCAN; ++u) {
                    const int i = i0 + u * 32 + lane;
                    c[u] = i < n ? __ldg(seg + i) : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < SCAN; ++u)
                    if ((i0 + u * 32 + lane) < n && sp_bin(c[u].y) >= tb) {
                        // pixel index -> (y, x) without an integer division: H * W < 2^24 here (the bitmap fits shared memory),
                        // so the float quotient is within one of the row and two compares settle it
                        const int e = (int)c[u].x;
                        int y = (int)((float)e * inv_w);
                        y -= (y * W > e) ? 1 : 0;
                        y += ((y + 1) * W <= e) ? 1 : 0;
                        const int x = e - y * W;
                        atomicOr(&bm[(y + SP_PAD) * BWR + ((x + 32) >> 5)], 1u << ((x + 32) & 31));
                    }
            }
        }
        __syncthreads();
        // ---- ranks: candidates before each bitmap word (row-major), then id -> (pixel, -score)
        const int wper = (L.words + 1023) / 1024;
        const int w0 = min(L.words, tid * wper), w1 = min(L.words, w0 + wper);
        int cnt = 0;
        for (int w = w0; w < w1; ++w) cnt += __popc(bm[w]);
        int total;
        int base = block_exclusive_scan(cnt, warp_sums, total);
        n0 = total;   // == n_sel
        {
            int r = w0 / BWR, wc = w0 - r * BWR;
            for (int w = w0; w < w1; ++w) {
                wbase[w] = (uint16_t)base;
                uint32_t bits = bm[w];
                while (bits) {
                    const int bit = __ffs(bits) - 1;
                    bits &= bits - 1;
                    pos[base++] = ((uint32_t)(r - SP_PAD) << 16) | (uint32_t)(wc * 32 + bit - 32);   // (y, x)
                }
                if (++wc == BWR) { wc = 0; ++r; }
            }
        }
        __syncthreads();
        {   // the scores: every load of a thread is issued before the first result is used (one round trip, not one per
            // load -- the volatile stores kept the simple loop in order: 22 % of this kernel's samples)
            constexpr int SC = SP2_CAP / 1024;
            float sc[SC];
#pragma unroll
            for (int u = 0; u < SC; ++u) {
                const int id = tid + u * 1024;
                sc[u] = 0.f;
                if (id < n0) {
                    const uint32_t yx = pos[id];
                    sc[u] = __ldg(img + (int)(yx >> 16) * W + (int)(yx & 0xffffu));
                }
            }
#pragma unroll
            for (int u = 0; u < SC; ++u) {
                const int id = tid + u * 1024;
                if (id < n0) st[id] = -sc[u];
            }
        }
        __syncthreads();

        // ---- round 0: higher-priority admitted neighbours of every admitted candidate ----
        // mask bit 8*r + c  <=>  neighbour at (dy, dx) = (r - 3, c - 3); scores are |state| (nobody is suppressed yet)
        for (int base0 = warp * 32; base0 < n0; base0 += 1024) {
            const int id = base0 + lane;
            bool still = false;
            if (id < n0) {
                const uint32_t yx = pos[id];
                const uint32_t sb = __float_as_uint(st[id]) & 0x7fffffffu;
                const int y = (int)(yx >> 16), x = (int)(yx & 0xffffu);
                const int bitpos = x + 32 - SP_PAD, w = bitpos >> 5, sh = bitpos & 31;
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    const uint32_t *bw = bm + (y + r) * BWR + w;  // row y + r - 3, padded by 3
                    const uint32_t win = __funnelshift_r(bw[0], bw[1], sh) & fp7[r];
                    if (r < 4) lo |= win << (8 * r); else hi |= win << (8 * (r - 4));
                }
                // earlier in row-major order (bits 0..26) wins ties; later ones need a strictly larger score
                uint64_t higher = 0;
                for (uint64_t m = ((uint64_t)hi << 32) | lo; m;) {
                    const int k = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const uint32_t nb_ = __float_as_uint(st[id_at(y + (k >> 3), bitpos + (k & 7))]) & 0x7fffffffu;
                    if (k < 27 ? nb_ >= sb : nb_ > sb) higher |= 1ull << k;
                }
                if (higher == 0) {
                    st[id] = __uint_as_float(sb);   // local maximum: kept
                    atomicAdd(&n_kept, 1);
                } else {
                    mask[id] = higher;
                    still = true;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, still);
            if (bal) {
                int slot = 0;
                if (lane == (__ffs(bal) - 1)) slot = atomicAdd(&n_next[0], __popc(bal));
                slot = __shfl_sync(0xffffffffu, slot, __ffs(bal) - 1);
                if (still) ids_a[slot + __popc(bal & ((1u << lane) - 1))] = (uint16_t)id;
            }
        }
        __syncthreads();

        // ---- rounds over the cached masks; the whole image is here, so every round makes progress ----
        int n = n_next[0];
        uint16_t *cur = ids_a, *nxt = ids_b;
        for (int round = 1; n > 0; ++round) {
            if (round > 96) {  // a long dependency chain (ramps, plateaus): the dense kernels handle those
                if (tid == 0) { redo_flags[b] = 1; surv_count[b] = 0; }
                return;
            }
            int *cntp = &n_next[round % 3];
            if (tid == 0) n_next[(round + 1) % 3] = 0;
            for (int base0 = warp * 32; base0 < n; base0 += 1024) {
                const int i = base0 + lane;
                bool still = false;
                int id = 0;
                if (i < n) {
                    id = cur[i];
                    const uint32_t yx = pos[id];
                    const int y = (int)(yx >> 16), x = (int)(yx & 0xffffu), bitpos = x + 32 - SP_PAD;
                    const uint64_t old = mask[id];
                    uint64_t left = old;
                    bool sup = false;
                    for (uint64_t m = old; m && !sup;) {
                        const int k = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        const float nv = st[id_at(y + (k >> 3), bitpos + (k & 7))];
                        if (nv > 0.f) sup = true;                      // kept higher-priority neighbour
                        else if (nv == 0.f) left &= ~(1ull << k);      // it was suppressed: no longer blocks
                    }
                    if (sup) st[id] = 0.f;
                    else if (left == 0) { st[id] = -st[id]; atomicAdd(&n_kept, 1); }
                    else {
                        if (left != old) mask[id] = left;
                        still = true;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, still);
                if (bal) {
                    int slot = 0;
                    if (lane == (__ffs(bal) - 1)) slot = atomicAdd(cntp, __popc(bal));
                    slot = __shfl_sync(0xffffffffu, slot, __ffs(bal) - 1);
                    if (still) nxt[slot + __popc(bal & ((1u << lane) - 1))] = (uint16_t)id;
                }
            }
            __syncthreads();
            n = *cntp;
            uint16_t *tmp = cur; cur = nxt; nxt = tmp;
        }
        __syncthreads();
        if (n_kept >= keep_top_k || tb == 0) break;  // enough survivors, or every candidate was admitted
        // too few: admit about twice as many (at least one more bin) and settle again from scratch
        target = min(SP2_CAP, max(2 * n_sel, target));
        if (target <= n_sel) target = n_sel + 1;
        if (n_sel >= SP2_CAP) {
            if (tid == 0) { redo_flags[b] = 1; surv_count[b] = 0; }
            return;
        }
        __syncthreads();
    }

    // ---- top-k cut: the keep_top_k highest scores, ties to the lower row-major index (= the lower id) ----
    // every thread owns SP2_PER consecutive ids, so block scans over the threads run in id (= row-major) order
    const int kept_total = n_kept;
    uint32_t sc[SP2_PER];   // score bits of my kept ids, 0 otherwise
#pragma unroll
    for (int j = 0; j < SP2_PER; ++j) {
        const int id = tid * SP2_PER + j;
        const float v = id < n0 ? st[id] : 0.f;
        sc[j] = v > 0.f ? __float_as_uint(v) : 0u;
    }
    uint32_t cut_score = 0;
    int take_ties = 0x7fffffff;          // how many of the entries that score exactly cut_score are kept
    if (keep_top_k > 0 && kept_total > keep_top_k) {
        int *shist = reinterpret_cast<int *>(mask);   // the neighbour masks are dead: 2048 bins fit in their place
        uint32_t prefix = 0;
        int remaining = keep_top_k;
        const int shifts[3] = {21, 10, 0}, widths[3] = {11, 11, 10};
        for (int pass = 0; pass < 3; ++pass) {
            for (int i = tid; i < SEL_BINS; i += 1024) shist[i] = 0;
            __syncthreads();
            const int sh = shifts[pass], hi_sh = sh + widths[pass];
            const uint32_t dmask = (1u << widths[pass]) - 1u;
#pragma unroll
            for (int j = 0; j < SP2_PER; ++j)
                if (sc[j] && (pass == 0 || (sc[j] >> hi_sh) == (prefix >> hi_sh))) atomicAdd(&shist[(sc[j] >> sh) & dmask], 1);
            __syncthreads();
            select_find_bin<true>(shist, remaining, warp_sums, &sel_bin, &sel_remaining);
            prefix |= (uint32_t)sel_bin << sh;
            remaining = sel_remaining;
            __syncthreads();
        }
        cut_score = prefix;       // the k-th highest score
        take_ties = remaining;    // of the survivors with exactly that score, the first `remaining` in row-major order
    }
    int ties = 0;
#pragma unroll
    for (int j = 0; j < SP2_PER; ++j) ties += (sc[j] != 0 && sc[j] == cut_score) ? 1 : 0;
    int tot_ties;
    int tie_rank = block_exclusive_scan(ties, warp_sums, tot_ties);
    bool keep[SP2_PER];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < SP2_PER; ++j) {
        keep[j] = false;
        if (sc[j] > cut_score) keep[j] = true;
        else if (sc[j] != 0 && sc[j] == cut_score) { keep[j] = tie_rank < take_ties; ++tie_rank; }
        mine += keep[j] ? 1 : 0;
    }
    int total_kept;
    int slot = block_exclusive_scan(mine, warp_sums, total_kept);
    float *dmap = dense ? dense + (size_t)b * H * W : nullptr;
    int64_t *kp = keypoints ? keypoints + (size_t)b * kp_cap * 2 : nullptr;
    float *ks = kp_scores ? kp_scores + (size_t)b * kp_cap : nullptr;
#pragma unroll
    for (int j = 0; j < SP2_PER; ++j) {
        if (!keep[j]) continue;
        const uint32_t yx = pos[tid * SP2_PER + j];
        const int y = (int)(yx >> 16), x = (int)(yx & 0xffffu);
        const float v = __uint_as_float(sc[j]);
        if (dmap) dmap[y * W + x] = v;
        if (slot < kp_cap) {
            if (kp) { kp[2 * (size_t)slot] = y; kp[2 * (size_t)slot + 1] = x; }
            if (ks) ks[slot] = v;
        }
        ++slot;
    }
    if (tid == 0) {
        if (kp_counts) kp_counts[b] = total_kept;
        surv_count[b] = -1;   // finished here: the select kernel skips this image
    }
}

__device__ __forceinline__ void
nms_select_body(float *out, int H, int W, int keep_top_k, const uint2 *survivors,
                const int *surv_count, int cap, uint32_t *__restrict__ bitmaps, int words,
                int64_t *__restrict__ keypoints, float *__restrict__ kp_scores, int32_t *__restrict__ kp_counts,
                int kp_cap) {
    __shared__ int hist[SEL_BINS];
    __shared__ int warp_sums[32];
    __shared__ int sel_bin, sel_remaining;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = *reinterpret_cast<const volatile int *>(surv_count + b);   // the fix-up in front of this may just have raised it
    if (n < 0) return;   // settled, cut and emitted by nms_sparse2_kernel
    const uint2 *surv = survivors + (size_t)b * cap;
    float *img = out + (size_t)b * H * W;
    const bool want_kp = keypoints != nullptr || kp_counts != nullptr;
    uint32_t *bitmap = bitmaps ? bitmaps + (size_t)b * words : nullptr;

    const bool cached = n <= SEL_CACHE * 1024;
    uint2 ent[SEL_CACHE];
    if (cached) {
#pragma unroll
        for (int k = 0; k < SEL_CACHE; ++k) {
            const int i = tid + k * 1024;
            ent[k] = i < n ? surv[i] : make_uint2(0xffffffffu, 0u);  // score bits 0 never match a live prefix
        }
    }
// visit every survivor (x = index, y = score bits) from registers, or from L2 when there are too many
#define MP_FOR_EACH_SURVIVOR(BODY)                                         \
    if (cached) {                                                          \
        _Pragma("unroll") for (int k_ = 0; k_ < SEL_CACHE; ++k_) {         \
            if (tid + k_ * 1024 < n) { const uint2 e = ent[k_]; BODY }      \
        }                                                                  \
    } else {                                                               \
        for (int i_ = tid; i_ < n; i_ += 1024) { const uint2 e = surv[i_]; BODY } \
    }

    uint32_t cut_score = 0, cut_index = 0xffffffffu;  // keep: score > cut_score || (score == cut_score && index <= cut_index)
    if (keep_top_k > 0 && n > keep_top_k) {
        uint32_t prefix = 0;
        int remaining = keep_top_k, in_bin = 0;
        const int shifts[3] = {21, 10, 0}, widths[3] = {11, 11, 10};
        for (int pass = 0; pass < 3; ++pass) {
            for (int i = tid; i < SEL_BINS; i += 1024) hist[i] = 0;
            __syncthreads();
            const int sh = shifts[pass], hi_sh = sh + widths[pass];
            const uint32_t dmask = (1u << widths[pass]) - 1u;
            MP_FOR_EACH_SURVIVOR(if (pass == 0 || (e.y >> hi_sh) == (prefix >> hi_sh)) atomicAdd(&hist[(e.y >> sh) & dmask], 1);)
            __syncthreads();
            select_find_bin<true>(hist, remaining, warp_sums, &sel_bin, &sel_remaining);
            prefix |= (uint32_t)sel_bin << sh;
            remaining = sel_remaining;
            in_bin = hist[sel_bin];
            __syncthreads();
        }
        cut_score = prefix;  // the k-th highest score; `remaining` of the `in_bin` survivors with it are kept
        if (remaining < in_bin) {
            // equal scores straddle k: keep the `remaining` lowest indices among them
            uint32_t iprefix = 0;
            for (int pass = 0; pass < 3; ++pass) {
                for (int i = tid; i < SEL_BINS; i += 1024) hist[i] = 0;
                __syncthreads();
                const int sh = shifts[pass], hi_sh = sh + widths[pass];
                const uint32_t dmask = (1u << widths[pass]) - 1u;
                MP_FOR_EACH_SURVIVOR(if (e.y == cut_score && (pass == 0 || (e.x >> hi_sh) == (iprefix >> hi_sh)))
                                         atomicAdd(&hist[(e.x >> sh) & dmask], 1);)
                __syncthreads();
                select_find_bin<false>(hist, remaining, warp_sums, &sel_bin, &sel_remaining);
                iprefix |= (uint32_t)sel_bin << sh;
                remaining = sel_remaining;
                __syncthreads();
            }
            cut_index = iprefix;
        }
    }
    if (want_kp) {
        for (int w = tid; w < words; w += blockDim.x) bitmap[w] = 0;
        __syncthreads();
    }
    MP_FOR_EACH_SURVIVOR(
        if (e.y > cut_score || (e.y == cut_score && e.x <= cut_index)) {
            if (want_kp) atomicOr(&bitmap[e.x >> 5], 1u << (e.x & 31));
        } else {
            img[e.x] = 0.f;  // cut by top-k (utils.py:109-116)
        })
#undef MP_FOR_EACH_SURVIVOR
    if (!want_kp) return;
    __syncthreads();
    emit_from_bitmap(bitmap, words, W, img, keypoints ? keypoints + (size_t)b * kp_cap * 2 : nullptr,
                     kp_scores ? kp_scores + (size_t)b * kp_cap : nullptr, kp_counts ? kp_counts + b : nullptr,
                     kp_cap, warp_sums);
}

// One CTA per image, one launch for both steps: the fix-up of the pixels whose dependency chain left their tile's apron,
// then (optionally) the top-k cut and the ordered keypoint emission.  Both only touch their own image, so the order
// inside the CTA is all the synchronisation they need; on the sparse top-k path both return at once for every image the
// sparse kernel settled (one ~3 us launch instead of two).
__global__ void __launch_bounds__(1024)
nms_fixup_select_kernel(float *out, int H, int W, const NmsFootprint fp, uint2 *survivors, int *surv_count,
                        const uint32_t *__restrict__ worklist, const int *__restrict__ work_count, int cap, int do_select,
                        int keep_top_k, uint32_t *__restrict__ bitmaps, int words, int64_t *__restrict__ keypoints,
                        float *__restrict__ kp_scores, int32_t *__restrict__ kp_counts, int kp_cap) {
    nms_fixup_body(out, H, W, fp, survivors, surv_count, worklist, work_count, cap);
    if (!do_select) return;
    __threadfence_block();
    __syncthreads();
    nms_select_body(out, H, W, keep_top_k, survivors, surv_count, cap, bitmaps, words, keypoints, kp_scores, kp_counts, kp_cap);
}

// ---- row 4b: torch.nonzero((p > thr).float() [* mask]) ----
__global__ void threshold_bitmap_kernel(const float *__restrict__ prob, const uint8_t *__restrict__ mask, int HW,
                                        float thr, uint32_t *__restrict__ bitmaps, int words) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // blockDim.x is a multiple of 32
    bool on = false;
    if (i < HW) {
        const float p = prob[(size_t)b * HW + i];
        // ((p > thr).float() * mask) != 0
        on = p > thr && (mask == nullptr || mask[(size_t)b * HW + i] != 0);
    }
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < words) bitmaps[(size_t)b * words + (i >> 5)] = m;
}

__global__ void __launch_bounds__(1024)
emit_keypoints_kernel(const float *__restrict__ prob, int H, int W, const uint32_t *__restrict__ bitmaps, int words,
                      int64_t *__restrict__ keypoints, float *__restrict__ kp_scores, int32_t *__restrict__ kp_counts,
                      int kp_cap) {
    __shared__ int warp_sums[32];
    const int b = blockIdx.x;
    emit_from_bitmap(bitmaps + (size_t)b * words, words, W, prob + (size_t)b * H * W,
                     keypoints ? keypoints + (size_t)b * kp_cap * 2 : nullptr,
                     kp_scores ? kp_scores + (size_t)b * kp_cap : nullptr, kp_counts ? kp_counts + b : nullptr, kp_cap,
                     warp_sums);
}

// torchvision's CPU nms test `inter / (area_i + area_j - inter) > iou` in fp32 with the quotient
// compared against the double threshold, for two size x size boxes at offset (dy,dx).
static bool footprint_hit(double size, double iou, int dy, int dx) {
    const float half = (float)(size * 0.5);
    const float a1 = 0.f - half, a2 = 0.f + half;
    const float area = (a2 - a1) * (a2 - a1);
    const float by1 = (float)dy - half, by2 = (float)dy + half;
    const float bx1 = (float)dx - half, bx2 = (float)dx + half;
    const float yy1 = a1 > by1 ? a1 : by1, xx1 = a1 > bx1 ? a1 : bx1;
    const float yy2 = a2 < by2 ? a2 : by2, xx2 = a2 < bx2 ? a2 : bx2;
    float w = yy2 - yy1; if (w < 0.f) w = 0.f;
    float h = xx2 - xx1; if (h < 0.f) h = 0.f;
    const float inter = w * h;
    const float ovr = inter / (area + area - inter);
    return (double)ovr > iou;
}

struct NmsLayout {
    int cap, words;
    size_t survivors, worklist, counts, counts_bytes, bitmaps, cands, segcnt, dense_scratch, total;
    NmsLayout(int B, int H, int W) {
        cap = H * W;
        words = (H * W + 31) / 32;
        size_t off = 0;
        // ints: surv_count[B] work_count[B] (unused)[B] redo_flags[B] hist[B][SP_BINS]; seg_count[B][nseg] lives with the lists
        counts_bytes = sizeof(int) * (4 + SP_BINS) * (size_t)B;
        counts = off;    off = align_up(off + counts_bytes, 256);
        const size_t nseg = ((size_t)H * W + SP_CHUNK - 1) / SP_CHUNK;
        segcnt = off;    off = align_up(off + sizeof(int) * (size_t)B * nseg, 256);
        cands = off;     off = align_up(off + sizeof(uint2) * (size_t)B * nseg * SP_SEG_CAP, 256);
        survivors = off; off = align_up(off + sizeof(uint2) * (size_t)B * cap, 256);
        worklist = off;  off = align_up(off + sizeof(uint32_t) * (size_t)B * cap, 256);
        bitmaps = off;   off = align_up(off + sizeof(uint32_t) * (size_t)B * words, 256);
        // stand-in for the dense output when the caller only wants keypoints (prob_nms == NULL): the tile kernels that
        // redo an image the sparse path gave up on keep their state in a dense map
        dense_scratch = off; off = align_up(off + sizeof(float) * (size_t)B * cap, 256);
        total = off;
    }
};

template <int TH, int TW, int E>
static int launch_tile(const float *prob, float *out, int B, int H, int W, float thr, const NmsFootprint &fp,
                       uint2 *surv, int *surv_count, uint32_t *work, int *work_count, int cap, bool vec,
                       cudaStream_t s) {
    constexpr int EH = TH + 2 * E, EW = TW + 2 * E;
    constexpr int BW = (EW + 31) / 32 + 1;
    constexpr size_t smem = (size_t)EH * EW * (sizeof(float) + 2 * sizeof(uint16_t)) + (size_t)EH * BW * sizeof(uint32_t);
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B);
    if (vec) {
        auto k = nms_tile_kernel<TH, TW, E, true>;
        MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, NMS_THREADS, smem, s>>>(prob, out, H, W, thr, fp, surv, surv_count, work, work_count, cap);
    } else {
        auto k = nms_tile_kernel<TH, TW, E, false>;
        MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, NMS_THREADS, smem, s>>>(prob, out, H, W, thr, fp, surv, surv_count, work, work_count, cap);
    }
    MP_LAUNCH_OK_S("nms_tile_kernel", s);
    return MP_OK;
}

template <int TH, int TW, int E, int CAP>
static int launch_tile_fast(const float *prob, float *out, int B, int H, int W, float thr, const NmsFootprint &fp,
                            uint2 *surv, int *surv_count, uint32_t *work, int *work_count, int cap, bool vec,
                            const int *only_flagged, cudaStream_t s) {
    constexpr int EH = TH + 2 * E, EW = TW + 2 * E;
    constexpr int BW = (EW + 31) / 32 + 1;
    constexpr size_t smem = (size_t)EH * EW * sizeof(float) + (size_t)CAP * (sizeof(uint64_t) + 4 * sizeof(uint16_t)) +
                            (size_t)EH * BW * sizeof(uint32_t);
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B);
    if (only_flagged != nullptr) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long tiles = (long long)grid.x * grid.y * B;
        const unsigned pgrid = (unsigned)(tiles < (long long)sms * 7 ? tiles : (long long)sms * 7);
        if (vec) {
            auto k = nms_tile_fast_redo_kernel<TH, TW, E, true, CAP>;
            MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<pgrid, NMS_THREADS, smem, s>>>(prob, out, H, W, thr, fp, surv, surv_count, work, work_count, cap, only_flagged,
                                              (int)grid.x, (int)grid.y, B);
        } else {
            auto k = nms_tile_fast_redo_kernel<TH, TW, E, false, CAP>;
            MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<pgrid, NMS_THREADS, smem, s>>>(prob, out, H, W, thr, fp, surv, surv_count, work, work_count, cap, only_flagged,
                                              (int)grid.x, (int)grid.y, B);
        }
        MP_LAUNCH_OK_S("nms_tile_fast_redo_kernel", s);
        return MP_OK;
    }
    if (vec) {
        auto k = nms_tile_fast_kernel<TH, TW, E, true, CAP>;
        MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, NMS_THREADS, smem, s>>>(prob, out, H, W, thr, fp, surv, surv_count, work, work_count, cap);
    } else {
        auto k = nms_tile_fast_kernel<TH, TW, E, false, CAP>;
        MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, NMS_THREADS, smem, s>>>(prob, out, H, W, thr, fp, surv, surv_count, work, work_count, cap);
    }
    MP_LAUNCH_OK_S("nms_tile_fast_kernel", s);
    return MP_OK;
}

}  // namespace mp

extern "C" size_t mp_box_nms_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 256;
    return mp::NmsLayout(B, H, W).total;
}

extern "C" int mp_box_nms_f32(const float *prob, int B, int H, int W, double size, double min_prob,
                              double iou, int keep_top_k, float *prob_nms, int64_t *keypoints,
                              float *kp_scores, int32_t *kp_counts, int kp_cap, void *workspace,
                              size_t workspace_bytes, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    using namespace mp;
    MP_CHECK_ARG(B >= 0 && H > 0 && W > 0, "mp_box_nms_f32: bad shape B=%d H=%d W=%d", B, H, W);
    MP_CHECK_ARG((long long)H * W < (1ll << 31), "mp_box_nms_f32: image too large");
    MP_CHECK_ARG(size > 0 && iou >= 0, "mp_box_nms_f32: size and iou must be positive");
    MP_CHECK_ARG(kp_cap >= 0 && keep_top_k >= 0, "mp_box_nms_f32: negative capacity / top-k");
    if (!(min_prob >= 0.0)) {
        set_error("mp_box_nms_f32: min_prob=%g < 0 is not supported (see multipoint_b200.h)", min_prob);
        return MP_ERR_UNSUPPORTED;
    }
    if (B == 0) return MP_OK;
    MP_CHECK_ARG(prob != nullptr, "mp_box_nms_f32: null pointer");
    MP_CHECK_ARG(prob_nms != nullptr || kp_counts != nullptr, "mp_box_nms_f32: no output requested (prob_nms and kp_counts are both NULL)");
    MP_CHECK_ARG(keypoints == nullptr || kp_counts != nullptr, "mp_box_nms_f32: keypoints need kp_counts");

    // footprint: reach and translation invariance
    const float half = (float)(size * 0.5);
    if (half * 2048.f != floorf(half * 2048.f) || H > 8192 || W > 8192) {
        set_error("mp_box_nms_f32: size=%g is not a multiple of 1/1024 (or image > 8192): the fp32 IoU "
                  "of the reference is then position dependent; unsupported", size);
        return MP_ERR_UNSUPPORTED;
    }
    NmsFootprint fp;
    memset(&fp, 0, sizeof(fp));
    const int Rmax = (int)ceil(size);
    int R = 0;
    for (int dy = -Rmax; dy <= Rmax; ++dy)
        for (int dx = -Rmax; dx <= Rmax; ++dx)
            if ((dy || dx) && footprint_hit(size, iou, dy, dx)) R = max(R, max(abs(dy), abs(dx)));
    if (R > 15) {
        set_error("mp_box_nms_f32: box size %g reaches %d px; at most 15 supported", size, R);
        return MP_ERR_UNSUPPORTED;
    }
    fp.R = R;
    for (int dy = -R; dy <= R; ++dy)
        for (int dx = -R; dx <= R; ++dx)
            if ((dy || dx) && footprint_hit(size, iou, dy, dx)) fp.rows[dy + R] |= 1u << (dx + R);

    const NmsLayout L(B, H, W);
    if (workspace == nullptr || workspace_bytes < L.total) {
        set_error("mp_box_nms_f32: workspace %zu B < required %zu B", workspace_bytes, L.total);
        return MP_ERR_WORKSPACE;
    }
    char *ws = (char *)workspace;
    int *counts = (int *)(ws + L.counts);
    int *surv_count = counts, *work_count = counts + B;
    uint2 *surv = (uint2 *)(ws + L.survivors);
    uint32_t *work = (uint32_t *)(ws + L.worklist);
    uint32_t *bitmaps = (uint32_t *)(ws + L.bitmaps);
    cudaStream_t s = (cudaStream_t)stream;
    MP_CUDA_OK(cudaMemsetAsync(counts, 0, L.counts_bytes, s));
    int *seg_count = (int *)(ws + L.segcnt), *redo_flags = counts + 3 * B, *hist = counts + 4 * B;
    uint2 *cands = (uint2 *)(ws + L.cands);

    const float thr = (float)min_prob;
    const bool want_dense = prob_nms != nullptr;
    if (!want_dense) prob_nms = (float *)(ws + L.dense_scratch);   // keypoints only: the map is scratch for the tile kernels
    const bool vec = (W % 4 == 0) && (((uintptr_t)prob & 15) == 0) && (((uintptr_t)prob_nms & 15) == 0);
    int rc;
    static const int tile_variant = getenv("MP_NMS_TILE") ? atoi(getenv("MP_NMS_TILE")) : 0;  // tuning aid
    static const bool no_sparse = getenv("MP_NMS_NO_SPARSE") != nullptr;                      // tuning aid

    // sparse top-k path: settle only the candidates that can reach the top k (see nms_sparse2_kernel)
    const size_t sp_smem = Sp2Layout(H, W).total;
    const int *only_flagged = nullptr;
    if (!no_sparse && keep_top_k > 0 && keep_top_k <= SP2_CAP / 2 && R <= SP_PAD && sp_smem <= 220 * 1024 && (long long)H * W < (1ll << 24)) {
        const int HW = H * W;
        dim3 cgrid((unsigned)((HW + SP_CHUNK - 1) / SP_CHUNK), (unsigned)B);
        const bool cvec = (HW % 4 == 0) && (((uintptr_t)prob & 15) == 0) && (((uintptr_t)prob_nms & 15) == 0);
        // the dense map is zero-filled only when the caller asked for it: the sparse kernel keeps its state in shared memory
        if (cvec && want_dense) nms_candidates_kernel<true, true><<<cgrid, NMS_THREADS, 0, s>>>(prob, prob_nms, W, HW, thr, cands, seg_count, hist);
        else if (cvec) nms_candidates_kernel<true, false><<<cgrid, NMS_THREADS, 0, s>>>(prob, prob_nms, W, HW, thr, cands, seg_count, hist);
        else if (want_dense) nms_candidates_kernel<false, true><<<cgrid, NMS_THREADS, 0, s>>>(prob, prob_nms, W, HW, thr, cands, seg_count, hist);
        else nms_candidates_kernel<false, false><<<cgrid, NMS_THREADS, 0, s>>>(prob, prob_nms, W, HW, thr, cands, seg_count, hist);
        MP_LAUNCH_OK_S("nms_candidates_kernel", s);
        MP_CUDA_OK(cudaFuncSetAttribute(nms_sparse2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp_smem));
        nms_sparse2_kernel<<<B, 1024, sp_smem, s>>>(prob, want_dense ? prob_nms : nullptr, H, W, keep_top_k, fp, cands, seg_count, hist,
                                                    keypoints, kp_scores, kp_counts, kp_cap, surv_count, redo_flags);
        MP_LAUNCH_OK_S("nms_sparse_kernel", s);
        only_flagged = redo_flags;  // the dense kernels below only redo images the sparse path gave up on
    }
    if (R <= 3 && tile_variant == 1)
        rc = launch_tile_fast<32, 128, 8, 1920>(prob, prob_nms, B, H, W, thr, fp, surv, surv_count, work, work_count, L.cap, vec, only_flagged, s);
    else if (R <= 3)
        rc = launch_tile_fast<32, 64, 8, 960>(prob, prob_nms, B, H, W, thr, fp, surv, surv_count, work, work_count, L.cap, vec, only_flagged, s);
    else if (R <= 8)
        rc = launch_tile<32, 128, 8>(prob, prob_nms, B, H, W, thr, fp, surv, surv_count, work, work_count, L.cap, vec, s);
    else
        rc = launch_tile<32, 128, 16>(prob, prob_nms, B, H, W, thr, fp, surv, surv_count, work, work_count, L.cap, vec, s);
    if (rc != MP_OK) return rc;

    const int do_select = (keep_top_k > 0 || keypoints != nullptr || kp_counts != nullptr) ? 1 : 0;
    nms_fixup_select_kernel<<<B, 1024, 0, s>>>(prob_nms, H, W, fp, surv, surv_count, work, work_count, L.cap, do_select, keep_top_k,
                                               bitmaps, L.words, keypoints, kp_scores, kp_counts, kp_cap);
    MP_LAUNCH_OK_S("nms_fixup_select_kernel", s);
    return MP_OK;
}

extern "C" size_t mp_extract_keypoints_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 256;
    return mp::align_up(sizeof(uint32_t) * (size_t)B * (((size_t)H * W + 31) / 32), 256);
}

extern "C" int mp_extract_keypoints_f32(const float *prob, const uint8_t *mask, int B, int H, int W,
                                        double threshold, int64_t *keypoints, float *kp_scores,
                                        int32_t *kp_counts, int kp_cap, void *workspace,
                                        size_t workspace_bytes, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    using namespace mp;
    MP_CHECK_ARG(B >= 0 && H > 0 && W > 0 && kp_cap >= 0, "mp_extract_keypoints_f32: bad shape");
    MP_CHECK_ARG((long long)H * W < (1ll << 31), "mp_extract_keypoints_f32: image too large");
    if (B == 0) return MP_OK;
    MP_CHECK_ARG(prob && kp_counts, "mp_extract_keypoints_f32: null pointer");
    const int HW = H * W, words = (HW + 31) / 32;
    const size_t need = mp_extract_keypoints_workspace_bytes(B, H, W);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("mp_extract_keypoints_f32: workspace %zu B < required %zu B", workspace_bytes, need);
        return MP_ERR_WORKSPACE;
    }
    uint32_t *bitmaps = (uint32_t *)workspace;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((HW + 255) / 256, B);
    threshold_bitmap_kernel<<<grid, 256, 0, s>>>(prob, mask, HW, (float)threshold, bitmaps, words);
    MP_LAUNCH_OK_S("threshold_bitmap_kernel", s);
    emit_keypoints_kernel<<<B, 1024, 0, s>>>(prob, H, W, bitmaps, words, keypoints, kp_scores, kp_counts, kp_cap);
    MP_LAUNCH_OK_S("emit_keypoints_kernel", s);
    return MP_OK;
}
