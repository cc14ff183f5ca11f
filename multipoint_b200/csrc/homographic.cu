// Rows 9-10 of the hot path: the warp / unwarp / aggregate loop of homographic adaptation
// (multipoint/utils/homographies.py:38-189) and warp_perspective_tensor (:404-425).
//
// The reference runs, per sampled homography, kornia's warper three times (matrix inverse,
// meshgrid, matmul, grid_sample), plus ~10 elementwise kernels on full-resolution tensors, and
// re-reads / re-writes the two accumulators every iteration.  Here:
//   mp_warp_f32          one launch warps the image batch for ALL sampled homographies
//   mp_ha_aggregate_f32  one launch per batch: every output pixel loops over the samples with
//                        its two accumulators in registers (sequential order = the reference's
//                        summation order), the spectra product/sum is formed per source pixel,
//                        and the finish (divide, sqrt | *0.5, min_count) is fused.
// HBM-bound: each warped heatmap and mask is read once (gathers hit L1/L2: a homography maps a
// 32x8 pixel block to a compact source quad), the output is written once.
//
// Coordinates follow the kornia-free restatement pinned in the CPU oracle (PARITY UNPINNED
// for kornia itself): destination grid linspace(-1,1) -> A (3x3, fp32, normalised
// space, computed on the host) -> multiply by 1/z (|z| > 1e-8) -> ATen grid_sample with
// align_corners=True.  Every fp32 operation is issued un-fused (__fmul_rn / __fadd_rn) in the
// oracle's order so the sampling position matches bit for bit.
#include "mp_common.cuh"

namespace mp {

__device__ __forceinline__ void src_coord(const float *A, float xs, float ys, int Ws, int Hs, float &ix, float &iy) {
    const float X = __fadd_rn(__fadd_rn(__fmul_rn(A[0], xs), __fmul_rn(A[1], ys)), A[2]);
    const float Y = __fadd_rn(__fadd_rn(__fmul_rn(A[3], xs), __fmul_rn(A[4], ys)), A[5]);
    const float Z = __fadd_rn(__fadd_rn(__fmul_rn(A[6], xs), __fmul_rn(A[7], ys)), A[8]);
    const float sc = fabsf(Z) > 1e-8f ? __fdiv_rn(1.0f, Z) : 1.0f;
    // (.. + 1) / 2 as a multiplication by 0.5: the same bits as the division for every finite input
    ix = __fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(X, sc), 1.f), 0.5f), (float)(Ws - 1));
    iy = __fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(Y, sc), 1.f), 0.5f), (float)(Hs - 1));
}

// ATen reflect_coordinates(in, 0, 2*(size-1)) + clip_coordinates
__device__ __forceinline__ float reflect_coord(float in, int size) {
    if (size <= 1) return 0.f;
    const float span = (float)(size - 1);
    in = fabsf(in);
    if (in <= span) return in;   // inside: zero flips (and in == span reflects onto itself); skips fmodf on the common path
    const float extra = fmodf(in, span);
    const int flips = (int)floorf(__fdiv_rn(in, span));
    float r = (flips % 2 == 0) ? extra : __fsub_rn(span, extra);
    return fminf(fmaxf(r, 0.f), (float)(size - 1));
}

struct Bilinear {
    int x0, y0;
    float nw, ne, sw, se;
    bool vx0, vx1, vy0, vy1, any;
};

__device__ __forceinline__ Bilinear bilinear_setup(float ix, float iy, int Ws, int Hs) {
    Bilinear q;
    const float fx = floorf(ix), fy = floorf(iy);
    q.any = fx >= -1.f && fx <= (float)Ws && fy >= -1.f && fy <= (float)Hs;  // false for NaN too
    q.x0 = q.any ? (int)fx : 0;
    q.y0 = q.any ? (int)fy : 0;
    const float x1 = (float)(q.x0 + 1), y1 = (float)(q.y0 + 1);
    q.nw = __fmul_rn(__fsub_rn(x1, ix), __fsub_rn(y1, iy));
    q.ne = __fmul_rn(__fsub_rn(ix, (float)q.x0), __fsub_rn(y1, iy));
    q.sw = __fmul_rn(__fsub_rn(x1, ix), __fsub_rn(iy, (float)q.y0));
    q.se = __fmul_rn(__fsub_rn(ix, (float)q.x0), __fsub_rn(iy, (float)q.y0));
    q.vx0 = q.x0 >= 0 && q.x0 < Ws; q.vx1 = q.x0 + 1 >= 0 && q.x0 + 1 < Ws;
    q.vy0 = q.y0 >= 0 && q.y0 < Hs; q.vy1 = q.y0 + 1 >= 0 && q.y0 + 1 < Hs;
    return q;
}

template <typename F>
__device__ __forceinline__ float bilinear_apply(const Bilinear &q, int Ws, F value) {
    float v = 0.f;
    if (!q.any) return v;
    if (q.vx0 && q.vx1 && q.vy0 && q.vy1) {   // interior: the four taps are issued together, same summation order
        const int o = q.y0 * Ws + q.x0;
        const float a = value(o), b = value(o + 1), c = value(o + Ws), d = value(o + Ws + 1);
        v = __fadd_rn(v, __fmul_rn(a, q.nw));
        v = __fadd_rn(v, __fmul_rn(b, q.ne));
        v = __fadd_rn(v, __fmul_rn(c, q.sw));
        return __fadd_rn(v, __fmul_rn(d, q.se));
    }
    if (q.vy0 && q.vx0) v = __fadd_rn(v, __fmul_rn(value(q.y0 * Ws + q.x0), q.nw));
    if (q.vy0 && q.vx1) v = __fadd_rn(v, __fmul_rn(value(q.y0 * Ws + q.x0 + 1), q.ne));
    if (q.vy1 && q.vx0) v = __fadd_rn(v, __fmul_rn(value((q.y0 + 1) * Ws + q.x0), q.sw));
    if (q.vy1 && q.vx1) v = __fadd_rn(v, __fmul_rn(value((q.y0 + 1) * Ws + q.x0 + 1), q.se));
    return v;
}

// A CTA owns a 32 x 8 block of destination pixels; each of its 8 warps an 8 x 4 patch of it (not a 32 x 1 row): the
// sampled homographies rotate by up to 90 degrees, and the source footprint of a compact patch stays within a few
// rows whatever the angle, where a 32-pixel row can map onto 32 different source rows = 32 L1 wavefronts per gather.
__device__ __forceinline__ void patch_pixel(int &x, int &y) {
    const int t = threadIdx.y * 32 + threadIdx.x, w = t >> 5, l = t & 31;
    x = blockIdx.x * 32 + (w & 3) * 8 + (l & 7);
    y = blockIdx.y * 8 + (w >> 2) * 4 + (l >> 3);
}

// grid (ceil(W/32), ceil(H/8), n_mats); each thread one destination pixel, all N planes
__global__ void __launch_bounds__(256)
warp_kernel(const float *__restrict__ src, int N, int H, int W, const float *__restrict__ A,
            const float *__restrict__ xs, const float *__restrict__ ys, int mode, int padding,
            float *__restrict__ out) {
    __shared__ float As[9];
    const int m = blockIdx.z;
    if (threadIdx.y == 0 && threadIdx.x < 9) As[threadIdx.x] = A[9 * m + threadIdx.x];
    __syncthreads();
    int x, y;
    patch_pixel(x, y);
    if (x >= W || y >= H) return;
    float ix, iy;
    src_coord(As, xs[x], ys[y], W, H, ix, iy);
    if (padding == MP_PAD_REFLECTION) { ix = reflect_coord(ix, W); iy = reflect_coord(iy, H); }
    const size_t HW = (size_t)H * W;
    float *o = out + (size_t)m * N * HW + (size_t)y * W + x;
    if (mode == MP_NEAREST) {
        const float rx = rintf(ix), ry = rintf(iy);
        const bool ok = rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H;
        const int off = ok ? (int)ry * W + (int)rx : 0;
        for (int n = 0; n < N; ++n) o[n * HW] = ok ? __ldg(src + n * HW + off) : 0.f;
    } else {
        const Bilinear q = bilinear_setup(ix, iy, W, H);
        for (int n = 0; n < N; ++n) {
            const float *pl = src + n * HW;
            o[n * HW] = bilinear_apply(q, W, [&](int off) { return __ldg(pl + off); });
        }
    }
}

// grid (ceil(W/32), ceil(H/8), B); dynamic smem: n*9 floats
template <int AGG, bool TWO>
__global__ void __launch_bounds__(256)
ha_aggregate_kernel(const float *__restrict__ prob0, const float *__restrict__ pa, const float *__restrict__ pb,
                    const uint8_t *__restrict__ masks, const float *__restrict__ Ainv, int n, int B, int H, int W,
                    const float *__restrict__ xs, const float *__restrict__ ys, int min_count, int flags,
                    float *__restrict__ prob_acc, float *__restrict__ count_acc, float *__restrict__ out) {
    extern __shared__ float As[];
    for (int i = threadIdx.y * 32 + threadIdx.x; i < n * 9; i += 256) As[i] = Ainv[i];
    __syncthreads();
    const int b = blockIdx.z;
    int x, y;
    patch_pixel(x, y);
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W;
    const size_t o = (size_t)b * HW + (size_t)y * W + x;
    float prob, count;
    if (flags & MP_HA_INIT) { prob = prob0[o]; count = 1.0f; }
    else { prob = prob_acc[o]; count = count_acc[o]; }
    const float xv = xs[x], yv = ys[y];
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        float ix, iy;
        src_coord(As + 9 * i, xv, yv, W, H, ix, iy);
        // count_sample: nearest / zeros warp of the valid mask (homographies.py:112,180)
        const float rx = rintf(ix), ry = rintf(iy);
        float cs = 0.f;
        if (rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H)
            cs = (float)__ldg(masks + (size_t)i * HW + (int)ry * W + (int)rx);
        const Bilinear q = bilinear_setup(ix, iy, W, H);
        const float *a = pa + ((size_t)i * B + b) * HW;
        float v;
        if (TWO) {
            const float *c = pb + ((size_t)i * B + b) * HW;
            v = bilinear_apply(q, W, [&](int off) {
                return AGG == MP_AGG_PROD ? __fmul_rn(__ldg(a + off), __ldg(c + off)) : __fadd_rn(__ldg(a + off), __ldg(c + off));
            });
        } else {
            v = bilinear_apply(q, W, [&](int off) { return __ldg(a + off); });
        }
        count = __fadd_rn(count, cs);
        prob = __fadd_rn(prob, __fmul_rn(v, cs));
    }
    if (flags & MP_HA_FINISH) {
        float r = __fdiv_rn(prob, count);
        if (AGG == MP_AGG_PROD) r = sqrtf(r);
        else if (AGG == MP_AGG_SUM) r = __fmul_rn(r, 0.5f);
        if (min_count > 0 && count < (float)min_count) r = 0.f;
        out[o] = r;
    } else {
        prob_acc[o] = prob;
        count_acc[o] = count;
    }
}

}  // namespace mp

extern "C" int mp_warp_f32(const float *src, int N, int n_mats, int H, int W, const float *A,
                           const float *xs, const float *ys, int mode, int padding, float *out,
                           mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(N >= 0 && n_mats >= 0 && H > 0 && W > 0, "mp_warp_f32: bad shape");
    MP_CHECK_ARG(mode == MP_BILINEAR || mode == MP_NEAREST, "mp_warp_f32: bad mode %d", mode);
    MP_CHECK_ARG(padding == MP_PAD_ZEROS || padding == MP_PAD_REFLECTION, "mp_warp_f32: bad padding %d", padding);
    MP_CHECK_ARG(n_mats <= 65535, "mp_warp_f32: at most 65535 matrices per call");
    if (N == 0 || n_mats == 0) return MP_OK;
    MP_CHECK_ARG(src && A && xs && ys && out, "mp_warp_f32: null pointer");
    dim3 grid((W + 31) / 32, (H + 7) / 8, n_mats), block(32, 8);
    mp::warp_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, N, H, W, A, xs, ys, mode, padding, out);
    MP_LAUNCH_OK_S("warp_kernel", (cudaStream_t)stream);
    return MP_OK;
}

extern "C" int mp_ha_aggregate_f32(const float *prob0, const float *probw_a, const float *probw_b,
                                   const uint8_t *masks, const float *Ainv, int n, int B, int H, int W,
                                   const float *xs, const float *ys, int aggregation, int min_count,
                                   int flags, float *prob_acc, float *count_acc, float *out,
                                   mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    using namespace mp;
    MP_CHECK_ARG(n >= 0 && B >= 0 && H > 0 && W > 0, "mp_ha_aggregate_f32: bad shape");
    MP_CHECK_ARG(aggregation == MP_AGG_NONE || aggregation == MP_AGG_PROD || aggregation == MP_AGG_SUM,
                 "mp_ha_aggregate_f32: bad aggregation %d", aggregation);
    MP_CHECK_ARG((probw_b != nullptr) == (aggregation != MP_AGG_NONE) || n == 0,
                 "mp_ha_aggregate_f32: a second spectrum needs aggregation prod|sum and vice versa");
    MP_CHECK_ARG(B <= 65535, "mp_ha_aggregate_f32: at most 65535 images per call");
    MP_CHECK_ARG((size_t)n * 9 * sizeof(float) <= 200 * 1024, "mp_ha_aggregate_f32: too many homographies per call");
    if (B == 0) return MP_OK;
    MP_CHECK_ARG(xs && ys, "mp_ha_aggregate_f32: null tables");
    MP_CHECK_ARG(n == 0 || (probw_a && masks && Ainv), "mp_ha_aggregate_f32: null sample pointer");
    MP_CHECK_ARG(!(flags & MP_HA_INIT) || prob0, "mp_ha_aggregate_f32: MP_HA_INIT needs prob0");
    MP_CHECK_ARG((flags & MP_HA_INIT) || (prob_acc && count_acc), "mp_ha_aggregate_f32: continuing needs prob_acc/count_acc");
    MP_CHECK_ARG(!(flags & MP_HA_FINISH) || out, "mp_ha_aggregate_f32: MP_HA_FINISH needs out");
    MP_CHECK_ARG((flags & MP_HA_FINISH) || (prob_acc && count_acc), "mp_ha_aggregate_f32: partial result needs prob_acc/count_acc");
    dim3 grid((W + 31) / 32, (H + 7) / 8, B), block(32, 8);
    const size_t smem = (size_t)n * 9 * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
#define MP_HA_LAUNCH(AGG, TWO)                                                                          \
    do {                                                                                                \
        auto k = ha_aggregate_kernel<AGG, TWO>;                                                         \
        if (smem > 48 * 1024) MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k<<<grid, block, smem, s>>>(prob0, probw_a, probw_b, masks, Ainv, n, B, H, W, xs, ys, min_count, flags, \
                                    prob_acc, count_acc, out);                                          \
    } while (0)
    if (aggregation == MP_AGG_PROD) MP_HA_LAUNCH(MP_AGG_PROD, true);
    else if (aggregation == MP_AGG_SUM) MP_HA_LAUNCH(MP_AGG_SUM, true);
    else MP_HA_LAUNCH(MP_AGG_NONE, false);
#undef MP_HA_LAUNCH
    MP_LAUNCH_OK_S("ha_aggregate_kernel", s);
    return MP_OK;
}
