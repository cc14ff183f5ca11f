// Row 5 of the hot path: utils.interpolate_descriptors (multipoint/utils/utils.py:159-167).
// Bilinear sample of the coarse descriptor map at each keypoint fused with the L2 normalisation,
// instead of the reference's cast / in-place scale / flip / grid_sample / transpose / normalize
// chain (~8 kernels).
//
// The sampling coordinate follows the reference's fp32 operation order exactly:
//   y_n = y / (H*0.5) - 1            (utils.py:162-163; note H/2, not (H-1)/2)
//   iy  = ((y_n + 1) / 2) * (Hc - 1) (ATen grid_sampler unnormalize, align_corners=True; the halving is
//         a multiplication by 0.5 here: the same bits)
// zero padding: a corner outside the map contributes 0.
//
// HBM-bound gather: one warp per keypoint, lanes across channels.  With the channels-last map
// (B,Hc,Wc,D) produced by mp_normalize_descriptors_f32 every corner is one contiguous D*4 B row;
// with NCHW each lane gathers strided words (kept for drop-in use on the reference's layout).
#include <cuda_bf16.h>

#include "mp_common.cuh"

namespace mp {

constexpr int SD_WARPS = 8;
constexpr int SD_MAX_CPL = 16;  // channels per lane held in registers: D <= 512

template <int LAYOUT>
__global__ void __launch_bounds__(SD_WARPS * 32)
sample_descriptors_kernel(const int64_t *__restrict__ kp, const int32_t *__restrict__ counts,
                          const float *__restrict__ desc, float *__restrict__ out, int B, int K, int D,
                          int Hc, int Wc, float half_h, float half_w) {
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * SD_WARPS + (threadIdx.x >> 5);
    if (item >= (long long)B * K) return;
    const int b = (int)(item / K), k = (int)(item - (long long)b * K);
    float *o = out + (size_t)item * D;
    const int n = counts ? min(counts[b], K) : K;
    if (k >= n) {
        for (int c = lane; c < D; c += 32) o[c] = 0.f;
        return;
    }
    const float y = (float)kp[2 * item], x = (float)kp[2 * item + 1];
    const float yn = __fsub_rn(__fdiv_rn(y, half_h), 1.0f);
    const float xn = __fsub_rn(__fdiv_rn(x, half_w), 1.0f);
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(yn, 1.0f), 0.5f), (float)(Hc - 1));
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.0f), 0.5f), (float)(Wc - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float nw = __fmul_rn((float)x1 - ix, (float)y1 - iy);
    const float ne = __fmul_rn(ix - (float)x0, (float)y1 - iy);
    const float sw = __fmul_rn((float)x1 - ix, iy - (float)y0);
    const float se = __fmul_rn(ix - (float)x0, iy - (float)y0);
    const bool vx0 = x0 >= 0 && x0 < Wc, vx1 = x1 >= 0 && x1 < Wc;
    const bool vy0 = y0 >= 0 && y0 < Hc, vy1 = y1 >= 0 && y1 < Hc;

    float v[SD_MAX_CPL];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < SD_MAX_CPL; ++j) {
        const int c = lane + 32 * j;
        float acc = 0.f;
        if (c < D) {
            if (LAYOUT == MP_LAYOUT_NHWC) {
                const float *m = desc + (size_t)b * Hc * Wc * D + c;
                if (vy0 && vx0) acc = __fadd_rn(acc, __fmul_rn(__ldg(m + ((size_t)y0 * Wc + x0) * D), nw));
                if (vy0 && vx1) acc = __fadd_rn(acc, __fmul_rn(__ldg(m + ((size_t)y0 * Wc + x1) * D), ne));
                if (vy1 && vx0) acc = __fadd_rn(acc, __fmul_rn(__ldg(m + ((size_t)y1 * Wc + x0) * D), sw));
                if (vy1 && vx1) acc = __fadd_rn(acc, __fmul_rn(__ldg(m + ((size_t)y1 * Wc + x1) * D), se));
            } else {
                const float *m = desc + ((size_t)b * D + c) * Hc * Wc;
                if (vy0 && vx0) acc = __fadd_rn(acc, __fmul_rn(__ldg(m + y0 * Wc + x0), nw));
                if (vy0 && vx1) acc = __fadd_rn(acc, __fmul_rn(__ldg(m + y0 * Wc + x1), ne));
                if (vy1 && vx0) acc = __fadd_rn(acc, __fmul_rn(__ldg(m + y1 * Wc + x0), sw));
                if (vy1 && vx1) acc = __fadd_rn(acc, __fmul_rn(__ldg(m + y1 * Wc + x1), se));
            }
        }
        v[j] = acc;
        ss += acc * acc;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int j = 0; j < SD_MAX_CPL; ++j) {
        const int c = lane + 32 * j;
        if (c < D) o[c] = v[j] / denom;
    }
}

// Channels-last fast path: a group of LPK lanes per keypoint, every lane moving 128-bit vectors: D = 4 * LPK * NV.
// A corner row of D*4 bytes is NV fully coalesced requests.  D = 64 (the shipped descriptor size) uses 16 lanes per
// keypoint, i.e. two keypoints per warp: with a whole warp per keypoint and 64-bit vectors it spent as many
// instructions per keypoint as D = 256 on a quarter of the bytes (47 % of the HBM roofline).
// SPLIT: also write what the tensor-core matcher consumes -- the rows split into bf16 planes (hi = bf16(v),
// mid = bf16(v - hi)) and their squared norms -- so that mp_match_split_f32 needs no prep pass over the descriptors.
template <int LPK, int NV, bool SPLIT>
__global__ void __launch_bounds__(SD_WARPS * 32)
sample_descriptors_nhwc_vec_kernel(const int64_t *__restrict__ kp, const int32_t *__restrict__ counts,
                                   const float *__restrict__ desc, float *__restrict__ out, int B, int K, int Hc, int Wc,
                                   float half_h, float half_w, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ mid,
                                   float *__restrict__ norms) {
    constexpr int D = 4 * LPK * NV, KPW = 32 / LPK;   // keypoints per warp
    const int lane = threadIdx.x & 31, sub = lane % LPK;
    const long long item = ((long long)blockIdx.x * SD_WARPS + (threadIdx.x >> 5)) * KPW + lane / LPK;
    const bool in_range = item < (long long)B * K;     // ragged last warp: the lanes still take part in the shuffles
    const int b = in_range ? (int)(item / K) : 0, k = in_range ? (int)(item - (long long)b * K) : 0;
    const int n = counts ? min(counts[b], K) : K;
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool live = in_range && k < n;
    if (live) {
        const float y = (float)kp[2 * item], x = (float)kp[2 * item + 1];
        const float yn = __fsub_rn(__fdiv_rn(y, half_h), 1.0f);
        const float xn = __fsub_rn(__fdiv_rn(x, half_w), 1.0f);
        const float iy = __fmul_rn(__fmul_rn(__fadd_rn(yn, 1.0f), 0.5f), (float)(Hc - 1));
        const float ix = __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.0f), 0.5f), (float)(Wc - 1));
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
        const float w4[4] = {__fmul_rn((float)x1 - ix, (float)y1 - iy), __fmul_rn(ix - (float)x0, (float)y1 - iy),
                             __fmul_rn((float)x1 - ix, iy - (float)y0), __fmul_rn(ix - (float)x0, iy - (float)y0)};
        const int cy[4] = {y0, y0, y1, y1}, cx[4] = {x0, x1, x0, x1};
        const float *m = desc + (size_t)b * Hc * Wc * D;
        // all corner rows are requested before the first is used (nw, ne, sw, se = the reference's accumulation order)
        float4 t[4][NV];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool ok = cy[c] >= 0 && cy[c] < Hc && cx[c] >= 0 && cx[c] < Wc;
            const float4 *rowp = reinterpret_cast<const float4 *>(m + ((size_t)(ok ? cy[c] : 0) * Wc + (ok ? cx[c] : 0)) * D);
#pragma unroll
            for (int j = 0; j < NV; ++j) t[c][j] = ok ? __ldg(rowp + sub + LPK * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (cy[c] < 0 || cy[c] >= Hc || cx[c] < 0 || cx[c] >= Wc) continue;   // a corner outside contributes nothing (not +0)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                v[j].x = __fadd_rn(v[j].x, __fmul_rn(t[c][j].x, w4[c]));
                v[j].y = __fadd_rn(v[j].y, __fmul_rn(t[c][j].y, w4[c]));
                v[j].z = __fadd_rn(v[j].z, __fmul_rn(t[c][j].z, w4[c]));
                v[j].w = __fadd_rn(v[j].w, __fmul_rn(t[c][j].w, w4[c]));
            }
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
#pragma unroll
    for (int s = LPK / 2; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    if (!SPLIT && !in_range) return;   // (SPLIT: the lanes still take part in the norm reduction below)
    if (live) {
        float e[4 * NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) { e[4 * j] = v[j].x; e[4 * j + 1] = v[j].y; e[4 * j + 2] = v[j].z; e[4 * j + 3] = v[j].w; }
        divide_all(e, fmaxf(sqrtf(ss), 1e-12f));
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = make_float4(e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3]);
    }
    if (in_range) {
        float4 *o = reinterpret_cast<float4 *>(out + (size_t)item * D);
#pragma unroll
        for (int j = 0; j < NV; ++j) o[sub + LPK * j] = v[j];
    }
    if (SPLIT) {
        float s2 = 0.f;     // squared norm of the row as written (the matcher's L2 term), lanes of the group reduce it
        uint2 *oh = reinterpret_cast<uint2 *>(hi + (size_t)item * D), *om = reinterpret_cast<uint2 *>(mid + (size_t)item * D);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            __align__(8) __nv_bfloat16 h[4], m[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                h[c] = __float2bfloat16_rn(e[c]);
                m[c] = __float2bfloat16_rn(e[c] - __bfloat162float(h[c]));
                s2 += e[c] * e[c];
            }
            if (in_range) {
                oh[sub + LPK * j] = *reinterpret_cast<const uint2 *>(h);
                om[sub + LPK * j] = *reinterpret_cast<const uint2 *>(m);
            }
        }
#pragma unroll
        for (int sft = LPK / 2; sft > 0; sft >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, sft);
        if (sub == 0 && in_range) norms[item] = s2;
    }
}

}  // namespace mp

static int sample_descriptors_impl(const char *fn, const int64_t *keypoints, const int32_t *kp_counts, int B, int K, const float *desc,
                                   int D, int Hc, int Wc, int layout, int H, int W, float *out, __nv_bfloat16 *hi,
                                   __nv_bfloat16 *mid, float *norms, cudaStream_t s) {
    using namespace mp;
    MP_CHECK_ARG(B >= 0 && K >= 0 && D > 0 && Hc > 0 && Wc > 0 && H > 0 && W > 0, "%s: bad shape", fn);
    MP_CHECK_ARG(D <= 32 * SD_MAX_CPL, "%s: D=%d > %d unsupported", fn, D, 32 * SD_MAX_CPL);
    MP_CHECK_ARG(layout == MP_LAYOUT_NCHW || layout == MP_LAYOUT_NHWC, "%s: bad layout %d", fn, layout);
    if ((long long)B * K == 0) return MP_OK;
    MP_CHECK_ARG(keypoints && desc && out, "%s: null pointer", fn);
    const long long items = (long long)B * K;
    const unsigned grid = (unsigned)((items + SD_WARPS - 1) / SD_WARPS);
    const float hh = (float)H * 0.5f, hw = (float)W * 0.5f;
    const bool aligned = (((uintptr_t)desc | (uintptr_t)out) & 15) == 0;
    const bool split = hi != nullptr;
    if (split) {
        if (!(layout == MP_LAYOUT_NHWC && aligned && (D == 64 || D == 128 || D == 256) && mid && norms &&
              (((uintptr_t)hi | (uintptr_t)mid) & 7) == 0)) {
            set_error("%s: the split outputs need the channels-last layout, D in {64,128,256} and aligned buffers", fn);
            return MP_ERR_UNSUPPORTED;
        }
    }
    if (layout == MP_LAYOUT_NHWC && aligned && (D == 64 || D == 128 || D == 256)) {
        const unsigned grid2 = (unsigned)((items + 2 * SD_WARPS - 1) / (2 * SD_WARPS));   // two keypoints per warp
#define MP_SD_LAUNCH(LPK, NV, G)                                                                                                  \
    do {                                                                                                                          \
        if (split) sample_descriptors_nhwc_vec_kernel<LPK, NV, true><<<G, SD_WARPS * 32, 0, s>>>(keypoints, kp_counts, desc, out, B, K, Hc, Wc, hh, hw, hi, mid, norms);   \
        else sample_descriptors_nhwc_vec_kernel<LPK, NV, false><<<G, SD_WARPS * 32, 0, s>>>(keypoints, kp_counts, desc, out, B, K, Hc, Wc, hh, hw, nullptr, nullptr, nullptr); \
    } while (0)
        if (D == 64) MP_SD_LAUNCH(16, 1, grid2);
        else if (D == 128) MP_SD_LAUNCH(32, 1, grid);
        else MP_SD_LAUNCH(32, 2, grid);
#undef MP_SD_LAUNCH
    } else if (layout == MP_LAYOUT_NHWC)
        sample_descriptors_kernel<MP_LAYOUT_NHWC><<<grid, SD_WARPS * 32, 0, s>>>(keypoints, kp_counts, desc, out, B, K, D, Hc, Wc, hh, hw);
    else
        sample_descriptors_kernel<MP_LAYOUT_NCHW><<<grid, SD_WARPS * 32, 0, s>>>(keypoints, kp_counts, desc, out, B, K, D, Hc, Wc, hh, hw);
    MP_LAUNCH_OK_S("sample_descriptors_kernel", s);
    return MP_OK;
}

extern "C" int mp_sample_descriptors_f32(const int64_t *keypoints, const int32_t *kp_counts, int B,
                                         int K, const float *desc, int D, int Hc, int Wc, int layout,
                                         int H, int W, float *out, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    return sample_descriptors_impl("mp_sample_descriptors_f32", keypoints, kp_counts, B, K, desc, D, Hc, Wc, layout, H, W, out,
                                   nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int mp_sample_descriptors_split_f32(const int64_t *keypoints, const int32_t *kp_counts, int B,
                                               int K, const float *desc, int D, int Hc, int Wc, int layout,
                                               int H, int W, float *out, void *hi_bf16, void *mid_bf16,
                                               float *sq_norms, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(hi_bf16 && mid_bf16 && sq_norms, "mp_sample_descriptors_split_f32: null split output");
    return sample_descriptors_impl("mp_sample_descriptors_split_f32", keypoints, kp_counts, B, K, desc, D, Hc, Wc, layout, H, W, out,
                                   (__nv_bfloat16 *)hi_bf16, (__nv_bfloat16 *)mid_bf16, sq_norms, (cudaStream_t)stream);
}
