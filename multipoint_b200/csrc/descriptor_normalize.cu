// Row 2 of the hot path: the tail of MultiPoint.descriptor_head
// (multipoint/models/MultiPoint.py:160-166): F.normalize(x, p=2, dim=1) = x / max(||x||_2, 1e-12)
// over the channel dimension of an NCHW map.
//
// HBM-bound: 4*D B read per cell, 4*D B written per requested layout.  One CTA owns 32
// consecutive cells (one 128 B line per channel); its 8 warps split the channels, keep their
// values in registers (single read of the input), exchange partial sums of squares through
// shared memory, and write NCHW straight from registers.  The optional channels-last copy
// (B,HW,D) -- the layout mp_sample_descriptors_f32 gathers contiguous rows from -- is transposed
// through shared memory so its stores are contiguous too.
#include "mp_common.cuh"

namespace mp {

constexpr int DN_WARPS = 8;

// CPT channels per thread (D <= CPT * DN_WARPS), M groups of 32 cells per CTA.  CPT * M = 32 values per thread for every
// supported D, so a CTA always moves 32 KB in and 32 KB out per layout: with one group of 32 cells the D = 64 case (the
// shipped descriptor size) moved 8 KB per CTA and ran at 50-59 % of the HBM roofline on per-CTA overhead alone.
template <int CPT, int M>
__global__ void __launch_bounds__(DN_WARPS * 32, 4)  // <= 64 registers: 4 CTAs/SM (ncu: 80 registers held it at 3, 36 % of the warp slots)
normalize_desc_kernel(const float *__restrict__ x, float *__restrict__ out_nchw,
                      float *__restrict__ out_nhwc, int D, int HW, int tiles_per_image) {
    extern __shared__ float smem[];  // [M][DN_WARPS][32] partials, then optional [32*M][D+1] tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x / tiles_per_image;
    const int p0 = (blockIdx.x - b * tiles_per_image) * (32 * M);
    const float *src = x + (size_t)b * D * HW + p0 + lane;

    float v[M][CPT];
    float ss[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const bool live = p0 + lane + 32 * m < HW;
        ss[m] = 0.f;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int c = warp + k * DN_WARPS;
            v[m][k] = (live && c < D) ? ld_stream_f(src + (size_t)c * HW + 32 * m) : 0.f;
        }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) ss[m] += v[m][k] * v[m][k];
        smem[(m * DN_WARPS + warp) * 32 + lane] = ss[m];
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < DN_WARPS; ++w) tot += smem[(m * DN_WARPS + w) * 32 + lane];
        divide_all(v[m], fmaxf(sqrtf(tot), 1e-12f));   // x / max(|x|, eps), one reciprocal per cell
    }

    if (out_nchw != nullptr) {
        float *dst = out_nchw + (size_t)b * D * HW + p0 + lane;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            if (p0 + lane + 32 * m >= HW) continue;
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int c = warp + k * DN_WARPS;
                if (c < D) dst[(size_t)c * HW + 32 * m] = v[m][k];
            }
        }
    }
    if (out_nhwc != nullptr) {
        float *tile = smem + M * DN_WARPS * 32;  // [32*M][D+1]
        const int ld = D + 1;
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int c = warp + k * DN_WARPS;
                if (c < D) tile[(32 * m + lane) * ld + c] = v[m][k];
            }
        __syncthreads();
        // each warp writes whole cells: D contiguous floats per cell
        for (int cell = warp; cell < 32 * M; cell += DN_WARPS) {
            if (p0 + cell >= HW) break;
            float *dst = out_nhwc + ((size_t)b * HW + p0 + cell) * D;
            for (int c = lane; c < D; c += 32) dst[c] = tile[cell * ld + c];
        }
    }
}

// NCHW (B,D,HW) -> channels-last (B,HW,D) copy of an already normalised map: what utils.interpolate_descriptors does once
// per call so that the sampler gathers contiguous D*4 B rows (the strided NCHW gather moves 4x the sectors)
__global__ void __launch_bounds__(256)
transpose_desc_kernel(const float *__restrict__ x, float *__restrict__ out, int D, int HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, p = p0 + tx;
        tile[ty + 8 * k][tx] = (c < D && p < HW) ? ld_stream_f(x + ((size_t)b * D + c) * HW + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int p = p0 + ty + 8 * k, c = c0 + tx;
        if (p < HW && c < D) out[((size_t)b * HW + p) * D + c] = tile[tx][ty + 8 * k];
    }
}

// any D: one thread per cell, two passes over the channels (second pass hits L2)
__global__ void normalize_desc_generic_kernel(const float *__restrict__ x, float *__restrict__ out_nchw,
                                              float *__restrict__ out_nhwc, int B, int D, int HW) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * HW) return;
    const int b = (int)(i / HW), p = (int)(i - (long long)b * HW);
    const float *src = x + (size_t)b * D * HW + p;
    float ss = 0.f;
    for (int c = 0; c < D; ++c) {
        const float v = src[(size_t)c * HW];
        ss += v * v;
    }
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    for (int c = 0; c < D; ++c) {
        const float v = src[(size_t)c * HW] / denom;
        if (out_nchw) out_nchw[(size_t)b * D * HW + (size_t)c * HW + p] = v;
        if (out_nhwc) out_nhwc[((size_t)b * HW + p) * D + c] = v;
    }
}

}  // namespace mp

extern "C" int mp_normalize_descriptors_f32(const float *x, int B, int D, int HW, float *out_nchw,
                                            float *out_nhwc, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(x != nullptr, "mp_normalize_descriptors_f32: null input");
    MP_CHECK_ARG(out_nchw || out_nhwc, "mp_normalize_descriptors_f32: no output requested");
    MP_CHECK_ARG(B >= 0 && D > 0 && HW > 0, "mp_normalize_descriptors_f32: bad shape B=%d D=%d HW=%d", B, D, HW);
    if (B == 0) return MP_OK;
    cudaStream_t s = (cudaStream_t)stream;
#define MP_DN_LAUNCH(CPT, M)                                                                                         \
    do {                                                                                                             \
        const int tiles = (HW + 32 * (M) - 1) / (32 * (M));                                                          \
        const size_t smem = ((M) * mp::DN_WARPS * 32 + (out_nhwc ? 32 * (M) * (D + 1) : 0)) * sizeof(float);          \
        /* shared memory is the other occupancy limit (~35 KB per CTA with the channels-last tile): full carve-out */ \
        cudaFuncSetAttribute(mp::normalize_desc_kernel<CPT, M>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
        mp::normalize_desc_kernel<CPT, M><<<(unsigned)(B * tiles), mp::DN_WARPS * 32, smem, s>>>(x, out_nchw, out_nhwc, D, HW, tiles); \
    } while (0)
    if (D <= 64) {
        MP_DN_LAUNCH(8, 4);
    } else if (D <= 128) {
        MP_DN_LAUNCH(16, 2);
    } else if (D <= 256) {
        MP_DN_LAUNCH(32, 1);
    } else {
        const long long total = (long long)B * HW;
        mp::normalize_desc_generic_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(x, out_nchw, out_nhwc, B, D, HW);
    }
#undef MP_DN_LAUNCH
    MP_LAUNCH_OK_S("normalize_desc_kernel", s);
    return MP_OK;
}

extern "C" int mp_transpose_descriptors_f32(const float *x, int B, int D, int HW, float *out_nhwc, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(B >= 0 && D > 0 && HW > 0, "mp_transpose_descriptors_f32: bad shape B=%d D=%d HW=%d", B, D, HW);
    if (B == 0) return MP_OK;
    MP_CHECK_ARG(x && out_nhwc, "mp_transpose_descriptors_f32: null pointer");
    MP_CHECK_ARG(B <= 65535 && (D + 31) / 32 <= 65535, "mp_transpose_descriptors_f32: too many images / channels per call");
    dim3 grid((HW + 31) / 32, (D + 31) / 32, B);
    mp::transpose_desc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out_nhwc, D, HW);
    MP_LAUNCH_OK_S("transpose_desc_kernel", (cudaStream_t)stream);
    return MP_OK;
}
