// Which start coordinates does a non-swizzled 3-D TMA window load accept?  (diagnostic for homographic.cu)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I multipoint_b200/csrc -I include -o tma_window tma_window.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int z, int bw, int bh, float *out) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bw * bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(s32(sm)), "l"(&map), "r"(s32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    __syncthreads();
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(s32(&bar)) : "memory");
    if (threadIdx.x < 4) out[threadIdx.x] = sm[threadIdx.x == 3 ? bw * bh - 1 : threadIdx.x];
}

int main() {
    const int W = 80, H = 64, P = 4;
    float *d; cudaMalloc(&d, W * H * P * 4);
    float *h = new float[W * H * P];
    for (int i = 0; i < W * H * P; ++i) h[i] = (float)i;
    cudaMemcpy(d, h, W * H * P * 4, cudaMemcpyHostToDevice);
    float *out; cudaMalloc(&out, 16);
    for (int bw : {40, 56}) {
        CUtensorMap map;
        cuuint64_t dims[3] = {W, H, P}; cuuint64_t strides[2] = {W * 4, (cuuint64_t)W * H * 4};
        cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bw, 1}, es[3] = {1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode box %d: %d\n", bw, (int)r);
        const int xs[] = {0, 32, 31, 63, 64, -1, 41, 44, 48, 79}, ys[] = {0, 30, -1};
        for (int x : xs) for (int y : ys) {
            cudaMemset(out, 0, 16);
            k<<<1, 32, bw * bw * 4>>>(map, x, y, 1, bw, bw, out);
            cudaError_t e = cudaDeviceSynchronize();
            float o[4] = {0, 0, 0, 0};
            if (e == cudaSuccess) cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
            printf("box %d start (%3d,%3d): %s  first=%g %g %g last=%g\n", bw, x, y, cudaGetErrorString(e), o[0], o[1], o[2], o[3]);
            if (e != cudaSuccess) { printf("(context lost)\n"); return 0; }
        }
    }
    return 0;
}
