// Microbenchmark: bilinear taps of a rotated sampling grid through tex2Dgather (block-linear CUDA array, one TEX
// instruction per pixel and plane) against four __ldg gathers from pitch-linear memory.  Also prints the component
// order of tex2Dgather at integer-corner coordinates.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tex_gather tools/microbench/tex_gather.cu && /tmp/tex_gather
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int W = 640, H = 512, NP = 2, NM = 99;

__global__ void order_kernel(cudaTextureObject_t tex, float *out) {
    // footprint (x0, y0) = (10, 20): gather at (11.0, 21.0)
    float4 g = tex2Dgather<float4>(tex, 11.0f, 21.0f, 0);
    out[0] = g.x; out[1] = g.y; out[2] = g.z; out[3] = g.w;
    g = tex2Dgather<float4>(tex, 10.5001f, 20.5001f, 0);
    out[4] = g.x; out[5] = g.y; out[6] = g.z; out[7] = g.w;
    g = tex2Dgather<float4>(tex, 11.4999f, 21.4999f, 0);
    out[8] = g.x; out[9] = g.y; out[10] = g.z; out[11] = g.w;
}

__device__ __forceinline__ void coords(const float *A, int x, int y, float &ix, float &iy) {
    ix = A[0] * x + A[1] * y + A[2];
    iy = A[3] * x + A[4] * y + A[5];
}

// quad-expanded source: element (y, x) = (v[y][x], v[y][x+1], v[y+1][x], v[y+1][x+1])
__global__ void expand_kernel(const float *__restrict__ src, float4 *__restrict__ quads) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP * H * W) return;
    const int x = i % W, y = (i / W) % H;
    const bool xr = x + 1 < W, yd = y + 1 < H;
    quads[i] = make_float4(src[i], xr ? src[i + 1] : 0.f, yd ? src[i + W] : 0.f, xr && yd ? src[i + W + 1] : 0.f);
}

// TEX: 0 = four __ldg, 1 = tex2Dgather (CUDA array), 2 = one 128-bit __ldg from the quad-expanded source, 3 = four tex2D point fetches (pitch-linear)
template <int TEX>
__global__ void __launch_bounds__(256) warp_kernel(cudaTextureObject_t tex, const float *__restrict__ src, const float *__restrict__ A, float *__restrict__ out) {
    __shared__ float As[6];
    const int m = blockIdx.z;
    const int t = threadIdx.x;
    if (t < 6) As[t] = A[6 * m + t];
    __syncthreads();
    const int w = t >> 5, l = t & 31;
    const int x = blockIdx.x * 32 + (w & 3) * 8 + (l & 7), y = blockIdx.y * 8 + (w >> 2) * 4 + (l >> 3);
    float ix, iy;
    coords(As, x, y, ix, iy);
    const float fx = floorf(ix), fy = floorf(iy);
    const bool ok = fx >= 0.f && fx < W - 1 && fy >= 0.f && fy < H - 1;
    const int x0 = ok ? (int)fx : 0, y0 = ok ? (int)fy : 0;
    const float ax = ix - fx, ay = iy - fy;
    for (int n = 0; n < NP; ++n) {
        float a, b, c, d;
        if (TEX == 1) {
            const float4 g = tex2Dgather<float4>(tex, (float)(x0 + 1), (float)(y0 + 1 + n * H), 0);
            a = g.w; b = g.z; c = g.x; d = g.y;
        } else if (TEX == 2) {
            const float4 g = __ldg(reinterpret_cast<const float4 *>(src) + (size_t)n * H * W + y0 * W + x0);
            a = g.x; b = g.y; c = g.z; d = g.w;
        } else if (TEX == 3) {
            const float u = x0 + 0.5f, v = y0 + n * H + 0.5f;
            a = tex2D<float>(tex, u, v); b = tex2D<float>(tex, u + 1, v); c = tex2D<float>(tex, u, v + 1); d = tex2D<float>(tex, u + 1, v + 1);
        } else {
            const float *p = src + (size_t)n * H * W + y0 * W + x0;
            a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + W); d = __ldg(p + W + 1);
        }
        float v = ok ? (a * (1 - ax) * (1 - ay) + b * ax * (1 - ay) + c * (1 - ax) * ay + d * ax * ay) : 0.f;
        out[((size_t)m * NP + n) * H * W + (size_t)y * W + x] = v;
    }
}

int main() {
    std::vector<float> h((size_t)NP * H * W);
    for (int n = 0; n < NP; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) h[((size_t)n * H + y) * W + x] = (float)(n * 1000000 + y * 1000 + x);
    float *src, *out, *A, *o12;
    CK(cudaMalloc(&src, h.size() * 4));
    CK(cudaMemcpy(src, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, (size_t)NM * NP * H * W * 4));
    CK(cudaMalloc(&o12, 64));
    std::vector<float> hA(6 * NM);
    for (int m = 0; m < NM; ++m) {
        const float th = (float)(m * 2.0 * M_PI / NM), s = 0.9f + 0.002f * m, cx = W / 2.f, cy = H / 2.f;
        const float c = cosf(th) * s, sn = sinf(th) * s;
        hA[6 * m + 0] = c; hA[6 * m + 1] = -sn; hA[6 * m + 2] = cx - c * cx + sn * cy;
        hA[6 * m + 3] = sn; hA[6 * m + 4] = c; hA[6 * m + 5] = cy - sn * cx - c * cy;
    }
    CK(cudaMalloc(&A, hA.size() * 4));
    CK(cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));

    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaArray_t arr;
    CK(cudaMallocArray(&arr, &cd, W, NP * H, cudaArrayTextureGather));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) CK(cudaMemcpy2DToArrayAsync(arr, 0, 0, src, W * 4, W * 4, NP * H, cudaMemcpyDeviceToDevice, 0));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("copy to array (%d x %d floats): %.2f us\n", W, NP * H, ms * 1000 / 20);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    order_kernel<<<1, 1>>>(tex, o12);
    float ho[12];
    CK(cudaMemcpy(ho, o12, 48, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; ++k) printf("gather %d: x=%.0f y=%.0f z=%.0f w=%.0f   (texel (x0,y0) = 20010, (x0+1,y0) = 20011, (x0,y0+1) = 21010, (x0+1,y0+1) = 21011)\n", k, ho[4 * k], ho[4 * k + 1], ho[4 * k + 2], ho[4 * k + 3]);

    dim3 grid(W / 32, H / 8, NM), block(256);
    std::vector<float> r0((size_t)NM * NP * H * W), r1(r0.size());
    float4 *quads;
    CK(cudaMalloc(&quads, h.size() * 16));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) expand_kernel<<<(NP * H * W + 255) / 256, 256>>>(src, quads);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("quad expansion of %d planes: %.2f us\n", NP, ms * 1000 / 20);
    cudaResourceDesc rp = {};
    rp.resType = cudaResourceTypePitch2D; rp.res.pitch2D.devPtr = src; rp.res.pitch2D.desc = cd;
    rp.res.pitch2D.width = W; rp.res.pitch2D.height = NP * H; rp.res.pitch2D.pitchInBytes = W * 4;
    cudaTextureObject_t texp;
    CK(cudaCreateTextureObject(&texp, &rp, &td, nullptr));
    const char *names[4] = {"4 x ldg", "tex2Dgather", "quad ldg.128", "4 x tex2D pitch2D"};
    for (int variant = 0; variant < 4; ++variant) {
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 10; ++i) {
                if (variant == 1) warp_kernel<1><<<grid, block>>>(tex, src, A, out);
                else if (variant == 2) warp_kernel<2><<<grid, block>>>(tex, reinterpret_cast<const float *>(quads), A, out);
                else if (variant == 3) warp_kernel<3><<<grid, block>>>(texp, src, A, out);
                else warp_kernel<0><<<grid, block>>>(tex, src, A, out);
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        CK(cudaGetLastError());
        printf("%s: %.1f us per launch (%d planes of %dx%d), %.0f GB/s written\n", names[variant], ms * 100, NM * NP, W, H,
               (double)NM * NP * H * W * 4 / (ms * 1e-4) / 1e9);
        CK(cudaMemcpy((variant ? r1 : r0).data(), out, r0.size() * 4, cudaMemcpyDeviceToHost));
        if (variant) {
            size_t diff = 0;
            for (size_t i = 0; i < r0.size(); ++i) diff += r0[i] != r1[i];
            printf("  elements that differ from the ldg variant: %zu of %zu\n", diff, r0.size());
        }
    }
    return 0;
}
