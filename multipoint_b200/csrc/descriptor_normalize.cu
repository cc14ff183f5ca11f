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

template <int CPT>  // channels per thread; D <= CPT * DN_WARPS
__global__ void __launch_bounds__(DN_WARPS * 32, 4)  // <= 64 registers: 4 CTAs/SM (ncu: 80 registers held it at 3, 36 % of the warp slots)
normalize_desc_kernel(const float *__restrict__ x, float *__restrict__ out_nchw,
                      float *__restrict__ out_nhwc, int D, int HW, int tiles_per_image) {
    extern __shared__ float smem[];  // [DN_WARPS][32] partials, then optional [32][D+1] tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x / tiles_per_image;
    const int p0 = (blockIdx.x - b * tiles_per_image) * 32;
    const int p = p0 + lane;
    const bool live = p < HW;
    const float *src = x + (size_t)b * D * HW + p;

    float v[CPT];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = warp + k * DN_WARPS;
        v[k] = (live && c < D) ? ld_stream_f(src + (size_t)c * HW) : 0.f;
        ss += v[k] * v[k];
    }
    smem[warp * 32 + lane] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < DN_WARPS; ++w) tot += smem[w * 32 + lane];
    const float denom = fmaxf(sqrtf(tot), 1e-12f);
#pragma unroll
    for (int k = 0; k < CPT; ++k) v[k] = v[k] / denom;

    if (out_nchw != nullptr && live) {
        float *dst = out_nchw + (size_t)b * D * HW + p;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int c = warp + k * DN_WARPS;
            if (c < D) dst[(size_t)c * HW] = v[k];
        }
    }
    if (out_nhwc != nullptr) {
        float *tile = smem + DN_WARPS * 32;  // [32][D+1]
        const int ld = D + 1;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int c = warp + k * DN_WARPS;
            if (c < D) tile[lane * ld + c] = v[k];
        }
        __syncthreads();
        // each warp writes whole cells: D contiguous floats per cell
        for (int cell = warp; cell < 32; cell += DN_WARPS) {
            if (p0 + cell >= HW) break;
            float *dst = out_nhwc + ((size_t)b * HW + p0 + cell) * D;
            for (int c = lane; c < D; c += 32) dst[c] = tile[cell * ld + c];
        }
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
    const int tiles = (HW + 31) / 32;
    const size_t smem = (mp::DN_WARPS * 32 + (out_nhwc ? 32 * (D + 1) : 0)) * sizeof(float);
    const unsigned grid = (unsigned)(B * tiles);
    if (D <= 64) {
        // shared memory is the other occupancy limit (35 KB per CTA with the channels-last tile): full carve-out
        cudaFuncSetAttribute(mp::normalize_desc_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        mp::normalize_desc_kernel<8><<<grid, mp::DN_WARPS * 32, smem, s>>>(x, out_nchw, out_nhwc, D, HW, tiles);
    } else if (D <= 128) {
        // shared memory is the other occupancy limit (35 KB per CTA with the channels-last tile): full carve-out
        cudaFuncSetAttribute(mp::normalize_desc_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        mp::normalize_desc_kernel<16><<<grid, mp::DN_WARPS * 32, smem, s>>>(x, out_nchw, out_nhwc, D, HW, tiles);
    } else if (D <= 256) {
        // shared memory is the other occupancy limit (35 KB per CTA with the channels-last tile): full carve-out
        cudaFuncSetAttribute(mp::normalize_desc_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        mp::normalize_desc_kernel<32><<<grid, mp::DN_WARPS * 32, smem, s>>>(x, out_nchw, out_nhwc, D, HW, tiles);
    } else {
        const long long total = (long long)B * HW;
        mp::normalize_desc_generic_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(x, out_nchw, out_nhwc, B, D, HW);
    }
    MP_LAUNCH_OK_S("normalize_desc_kernel", s);
    return MP_OK;
}
