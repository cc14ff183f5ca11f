// Row 1 of the hot path: MultiPoint.detector_head (multipoint/models/MultiPoint.py:150-158).
// One fused pass: 65-channel softmax, dustbin drop, depth-to-space(8) and the optional
// valid-mask multiply, instead of the reference's softmax + slice + pixel_shuffle kernels.
//
// HBM-bound: 65*4 B read + 64*4 B written per cell (2 641 920 B per 512x640 image).
// Mapping: one thread per coarse cell, one warp per 32 consecutive cells.  Every channel load of
// a warp is one full 128 B line; the 65 loads of a thread are independent, so each warp keeps
// ~8 KB in flight.  The 8x8 output block of a cell is scattered over 8 image rows, so each row
// is transposed through 1 KB of shared memory per warp and leaves as 128-bit stores that cover
// 2 x 512 contiguous bytes per warp.
#include "mp_common.cuh"

namespace mp {

constexpr int DH_WARPS = 8;

// ML = SuperPointMagicLeap.generate_heatmap (multipoint/models/SuperPointMagicLeap.py:68-85): the same
// index map with exp(x) / (sum exp(x) + 1e-5) and no max subtraction (overflows like the reference).
template <bool ML>
__global__ void __launch_bounds__(DH_WARPS * 32, 3)
detector_head_kernel(const float *__restrict__ logits, const uint8_t *__restrict__ mask,
                     float *__restrict__ prob, long long total_cells, int cells, int Wc) {
    __shared__ __align__(16) float tile[DH_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * DH_WARPS + warp;
    const long long g0 = unit * 32;
    if (g0 >= total_cells) return;
    const long long g = g0 + lane;
    const bool live = g < total_cells;

    float x[65];
    if (live) {
        const long long b = g / cells;
        const int p = (int)(g - b * cells);
        const float *src = logits + (b * 65) * cells + p;
#pragma unroll
        for (int c = 0; c < 65; ++c) x[c] = ld_stream_f(src + (size_t)c * cells);
    } else {
#pragma unroll
        for (int c = 0; c < 65; ++c) x[c] = 0.f;
    }
    float m = 0.f;
    if (!ML) {
        m = x[0];
#pragma unroll
        for (int c = 1; c < 65; ++c) m = fmaxf(m, x[c]);
    }
    // exp(d) = 2^(d*log2(e)) with log2(e) split hi/lo so the argument keeps ~2^-30 relative
    // accuracy, then ex2.approx (2^-22.5): ~3e-7 relative overall, 4 instructions instead of ~10.
    // One reciprocal replaces 64 IEEE divisions (<= 1 ulp each).  Both sit well inside the 1e-5
    // contract; the arithmetic, not the memory system, limited the first version (ncu: issue 57 %).
    const float L2E_HI = 1.44269502162933349609375f, L2E_LO = 1.925963033500011e-8f;
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 65; ++c) {
        const float d = x[c] - m;
        const float t = fmaf(d, L2E_LO, d * L2E_HI);
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
        x[c] = e;
        sum += e;  // channel order, like the oracle
    }
    const float inv = __frcp_rn(ML ? __fadd_rn(sum, 1e-5f) : sum);

    // the two float4 this lane stores per output row: float4 index f -> cell f/2, half f&1
    const int W = Wc * 8;
    long long out_off[2];
    bool out_ok[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int f = lane + 32 * k;
        const long long gc = g0 + (f >> 1);
        out_ok[k] = gc < total_cells;
        const long long b = gc / cells;
        const int p = (int)(gc - b * cells);
        const int h = p / Wc, w = p - h * Wc;
        out_off[k] = (b * cells * 64) + (long long)(8 * h) * W + 8 * w + 4 * (f & 1);
    }

    float *t = tile[warp];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 a, c;
        a.x = x[8 * i + 0] * inv; a.y = x[8 * i + 1] * inv; a.z = x[8 * i + 2] * inv; a.w = x[8 * i + 3] * inv;
        c.x = x[8 * i + 4] * inv; c.y = x[8 * i + 5] * inv; c.z = x[8 * i + 6] * inv; c.w = x[8 * i + 7] * inv;
        __syncwarp();
        // the two 16 B halves of a cell swap places in every second group of four cells, which makes
        // both the stores (8 lanes x 16 B per phase) and the loads below bank-conflict free
        const int sw = ((lane >> 2) & 1) * 4;
        *reinterpret_cast<float4 *>(t + 8 * lane + sw) = a;
        *reinterpret_cast<float4 *>(t + 8 * lane + (4 ^ sw)) = c;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!out_ok[k]) continue;
            const int f = lane + 32 * k, cell = f >> 1;
            float4 v = *reinterpret_cast<const float4 *>(t + 8 * cell + (((f & 1) ^ ((cell >> 2) & 1)) * 4));
            const long long o = out_off[k] + (long long)i * W;
            if (mask != nullptr) {
                const uchar4 mk = *reinterpret_cast<const uchar4 *>(mask + o);
                v.x *= (float)mk.x; v.y *= (float)mk.y; v.z *= (float)mk.z; v.w *= (float)mk.w;
            }
            st_stream_f4(reinterpret_cast<float4 *>(prob + o), v);
        }
    }
}

// utils.depth_to_space (multipoint/utils/utils.py:64-69), any block size: pure index map.
__global__ void depth_to_space_kernel(const float *__restrict__ x, float *__restrict__ out,
                                      long long total, int C, int Hc, int Wc, int bs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int W = Wc * bs, H = Hc * bs;
    const int xo = (int)(i % W);
    const int yo = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % C);
    const long long b = i / ((long long)W * H * C);
    const int h = yo / bs, ii = yo - h * bs, w = xo / bs, jj = xo - w * bs;
    // input viewed as (N, bs, bs, C, Hc, Wc)
    const long long src = ((((b * bs + ii) * bs + jj) * C + c) * Hc + h) * Wc + w;
    out[i] = x[src];
}

}  // namespace mp

extern "C" int mp_detector_head_f32(const float *logits, int B, int Hc, int Wc,
                                    const uint8_t *valid_mask, float *prob, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(logits && prob, "mp_detector_head_f32: null pointer");
    MP_CHECK_ARG(B >= 0 && Hc > 0 && Wc > 0, "mp_detector_head_f32: bad shape B=%d Hc=%d Wc=%d", B, Hc, Wc);
    MP_CHECK_ARG(((uintptr_t)prob & 15) == 0, "mp_detector_head_f32: prob must be 16-byte aligned");
    MP_CHECK_ARG(valid_mask == nullptr || ((uintptr_t)valid_mask & 3) == 0,
                 "mp_detector_head_f32: valid_mask must be 4-byte aligned");
    if (B == 0) return MP_OK;
    const long long total = (long long)B * Hc * Wc;
    const long long units = (total + 31) / 32;
    const unsigned grid = (unsigned)((units + mp::DH_WARPS - 1) / mp::DH_WARPS);
    mp::detector_head_kernel<false><<<grid, mp::DH_WARPS * 32, 0, (cudaStream_t)stream>>>(
        logits, valid_mask, prob, total, Hc * Wc, Wc);
    MP_LAUNCH_OK_S("detector_head_kernel", (cudaStream_t)stream);
    return MP_OK;
}

extern "C" int mp_heatmap_magicleap_f32(const float *semi, int B, int Hc, int Wc, float *prob, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(semi && prob, "mp_heatmap_magicleap_f32: null pointer");
    MP_CHECK_ARG(B >= 0 && Hc > 0 && Wc > 0, "mp_heatmap_magicleap_f32: bad shape B=%d Hc=%d Wc=%d", B, Hc, Wc);
    MP_CHECK_ARG(((uintptr_t)prob & 15) == 0, "mp_heatmap_magicleap_f32: prob must be 16-byte aligned");
    if (B == 0) return MP_OK;
    const long long total = (long long)B * Hc * Wc;
    const long long units = (total + 31) / 32;
    const unsigned grid = (unsigned)((units + mp::DH_WARPS - 1) / mp::DH_WARPS);
    mp::detector_head_kernel<true><<<grid, mp::DH_WARPS * 32, 0, (cudaStream_t)stream>>>(
        semi, nullptr, prob, total, Hc * Wc, Wc);
    MP_LAUNCH_OK_S("detector_head_kernel", (cudaStream_t)stream);
    return MP_OK;
}

extern "C" int mp_depth_to_space_f32(const float *x, int B, int C, int Hc, int Wc, int block,
                                     float *out, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(x && out, "mp_depth_to_space_f32: null pointer");
    MP_CHECK_ARG(B >= 0 && C > 0 && Hc > 0 && Wc > 0 && block > 0, "mp_depth_to_space_f32: bad shape");
    const long long total = (long long)B * C * Hc * Wc * block * block;
    if (total == 0) return MP_OK;
    const unsigned grid = (unsigned)((total + 255) / 256);
    mp::depth_to_space_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, total, C, Hc, Wc, block);
    MP_LAUNCH_OK_S("depth_to_space_kernel", (cudaStream_t)stream);
    return MP_OK;
}
