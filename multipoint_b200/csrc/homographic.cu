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
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "mp_common.cuh"
#include "mp_tma.cuh"

namespace mp {

__device__ __forceinline__ void src_coord(const float *A, float xs, float ys, int Ws, int Hs, float &ix, float &iy) {
    const float X = __fadd_rn(__fadd_rn(__fmul_rn(A[0], xs), __fmul_rn(A[1], ys)), A[2]);
    const float Y = __fadd_rn(__fadd_rn(__fmul_rn(A[3], xs), __fmul_rn(A[4], ys)), A[5]);
    const float Z = __fadd_rn(__fadd_rn(__fmul_rn(A[6], xs), __fmul_rn(A[7], ys)), A[8]);
    const float sc = fabsf(Z) > 1e-8f ? __frcp_rn(Z) : 1.0f;   // the correctly rounded 1 / Z, i.e. the same bits as the division
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

// grid (ceil(W/32), ceil(H/8), n_mats); each thread one destination pixel, all N planes.
// GATHER: the N source planes also sit, stacked vertically, in a block-linear CUDA array, and the four taps of an
// interior pixel come back from ONE tex2Dgather (texels (x0,y0+1), (x1,y0+1), (x1,y0), (x0,y0) in .x .y .z .w, measured
// by tools/microbench/tex_gather.cu) fetched at the corner the four texels share -- half a texel away from every
// footprint boundary, so the unit's fixed-point coordinate rounding cannot pick another footprint.  The values are the
// stored fp32 texels (point sampling, no filtering hardware in the arithmetic); the blend is the same un-fused
// sequence, so the result is bit-identical to the direct path.  Four separate gathers of a rotated 8 x 4 patch cost
// ~30 L1 wavefronts per warp and plane, the texture unit 16 cycles (microbenchmark: 230 -> 119 us for 198 planes).
template <bool GATHER>
__global__ void __launch_bounds__(256, GATHER ? 5 : 6)
warp_kernel(const float *__restrict__ src, cudaTextureObject_t tex, int N, int G, int H, int W, const float *__restrict__ A,
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
    // plane n = group n / G, member n % G; out (N / G, n_mats, G, H, W): every group gets its own (n_mats, G, H, W) block.
    // One flat loop over the planes; the output pointer jumps to the next group's block after every G planes.
    float *o = out + (size_t)m * G * HW + (size_t)y * W + x;
    const size_t group_jump = ((size_t)gridDim.z - 1) * G * HW;   // from the end of one group's G planes to the next group's first
    int k = G;
    if (mode == MP_NEAREST) {
        const float rx = rintf(ix), ry = rintf(iy);
        const bool ok = rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H;
        const float *pl = src + (ok ? (int)ry * W + (int)rx : 0);
        for (int n = 0; n < N; ++n, pl += HW, o += HW) {
            *o = ok ? __ldg(pl) : 0.f;
            if (--k == 0) { k = G; o += group_jump; }
        }
    } else {
        const Bilinear q = bilinear_setup(ix, iy, W, H);
        if (GATHER && q.any && q.vx0 && q.vx1 && q.vy0 && q.vy1) {
            const float u = (float)(q.x0 + 1), fh = (float)H;
            float v0 = (float)(q.y0 + 1);   // exact: plane rows stay below 2^24
            for (int n = 0; n < N; ++n, v0 += fh, o += HW) {
                const float4 g = tex2Dgather<float4>(tex, u, v0, 0);
                float v = __fadd_rn(0.f, __fmul_rn(g.w, q.nw));
                v = __fadd_rn(v, __fmul_rn(g.z, q.ne));
                v = __fadd_rn(v, __fmul_rn(g.x, q.sw));
                *o = __fadd_rn(v, __fmul_rn(g.y, q.se));
                if (--k == 0) { k = G; o += group_jump; }
            }
            return;
        }
        const float *pl = src;
        for (int n = 0; n < N; ++n, pl += HW, o += HW) {
            *o = bilinear_apply(q, W, [&](int off) { return __ldg(pl + off); });
            if (--k == 0) { k = G; o += group_jump; }
        }
    }
}

// The library's own gather arrays (not caller-visible memory): one per (device, stream, width), grown on demand and
// kept for the life of the process, so that calls on one stream reuse theirs in stream order and calls on different
// streams or devices (nn.DataParallel: one thread per GPU) never share one.
struct GatherArray {
    int dev;
    cudaStream_t stream;
    int W, rows;
    cudaArray_t arr;
    cudaTextureObject_t tex;
};
static std::mutex g_gather_mutex;
static std::vector<GatherArray> g_gather;

// 0 = ok, 1 = not available for this shape (caller uses the direct path), < 0 = error
static int gather_array_for(int W, int rows, cudaStream_t s, cudaTextureObject_t *tex, cudaArray_t *arr) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    int maxw = 0, maxh = 0;
    cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxTexture2DGatherWidth, dev);
    cudaDeviceGetAttribute(&maxh, cudaDevAttrMaxTexture2DGatherHeight, dev);
    if (W > maxw || rows > maxh) return 1;
    // allocating (or freeing) an array is not a stream operation: while the stream is being captured into a graph only
    // an array that already exists may be used
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &capturing) != cudaSuccess) { cudaGetLastError(); return 1; }
    std::lock_guard<std::mutex> lock(g_gather_mutex);
    for (GatherArray &g : g_gather) {
        if (g.dev != dev || g.stream != s || g.W != W) continue;
        if (g.rows >= rows) { *tex = g.tex; *arr = g.arr; return 0; }
        if (capturing != cudaStreamCaptureStatusNone) return 1;
        // too small: replace (the frees synchronise; this happens once per shape)
        cudaDestroyTextureObject(g.tex);
        cudaFreeArray(g.arr);
        g = g_gather.back();
        g_gather.pop_back();
        break;
    }
    if (capturing != cudaStreamCaptureStatusNone) return 1;
    GatherArray g{dev, s, W, rows, nullptr, 0};
    const cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    if (cudaMallocArray(&g.arr, &cd, (size_t)W, (size_t)rows, cudaArrayTextureGather) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = g.arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    if (cudaCreateTextureObject(&g.tex, &rd, &td, nullptr) != cudaSuccess) {
        cudaGetLastError();
        cudaFreeArray(g.arr);
        return 1;
    }
    g_gather.push_back(g);
    *tex = g.tex;
    *arr = g.arr;
    return 0;
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


// ---------------------------------------------------------------------------------------------------------------
// Staged unwarp + aggregate.  The direct kernel above is bound by L1 wavefronts: a warp's gather touches one cache
// line per source row of its patch's footprint (7-9 lines per request on the reference's +-180 degree homography
// distribution; ncu: l1tex data-pipe wavefronts 86 %, issue active 67 %).  Here a CTA owns a 32 x 32 block of output
// pixels (four per thread), and for every sample
//   1. all threads compute their source coordinates and reduce the bounding box of the taps they will read
//      (warp-collective min / max, then shared-memory atomics),
//   2. one thread has the TMA engine copy that window of the sample's heatmap(s) and mask into shared memory --
//      box sizes 40 / 56 / 72 (the smallest that holds the window; texels outside the image arrive as zeros, which is
//      exactly the zero padding of the unwarp) -- bypassing L1,
//   3. the taps are read from shared memory (box widths are 8 mod 16 words: an 8 x 4 patch is conflict-free).
// A window larger than 72 x 72 (extreme perspective) falls back to the direct gathers for that sample.
// Arithmetic, tap order and accumulation order are those of the direct kernel: results are bit-identical.
constexpr int HS_T = 32;                          // tile side
constexpr int HS_BOX[3] = {40, 56, 72};           // float box sides
constexpr int HS_MBOX[3] = {64, 80, 96};          // mask box widths (bytes: multiples of 16)
constexpr int HS_MAXBOX = 72, HS_MAXMBOX = 96;
// The innermost start coordinate of a TMA window must be 16-byte aligned (measured: an unaligned start traps with an
// illegal instruction, tools/microbench/tma_window.cu), so the float windows start at x & ~3 and the mask windows at
// x & ~15: a box of side S then serves tap windows up to S - 3 wide, and the mask box is S + 24 bytes wide.

struct HaMaps {
    CUtensorMap a[3], b[3], m[3];
};

template <int AGG, bool TWO>
__global__ void __launch_bounds__(256)
ha_aggregate_staged_kernel(const __grid_constant__ HaMaps maps, const float *__restrict__ prob0, const float *__restrict__ pa,
                           const float *__restrict__ pb, const uint8_t *__restrict__ masks, const float *__restrict__ Ainv,
                           int n, int B, int H, int W, const float *__restrict__ xs, const float *__restrict__ ys,
                           int min_count, int flags, float *__restrict__ prob_acc, float *__restrict__ count_acc,
                           float *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t hs_smem_raw[];
    uint8_t *hs_smem = hs_smem_raw + ((128u - (smem_u32(hs_smem_raw) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
    float *sa = reinterpret_cast<float *>(hs_smem);
    float *sb = sa + HS_MAXBOX * HS_MAXBOX;
    uint8_t *smk = reinterpret_cast<uint8_t *>(sb + (TWO ? HS_MAXBOX * HS_MAXBOX : 0));
    float *As = reinterpret_cast<float *>(smk + HS_MAXMBOX * HS_MAXBOX);
    __shared__ int bbox[3][4];                     // per sample (mod 3): xlo, ylo, xhi, yhi
    __shared__ __align__(8) uint64_t bar_mem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const uint32_t bar = smem_u32(&bar_mem);
    for (int i = tid; i < n * 9; i += 256) As[i] = Ainv[i];
    if (tid < 12) (&bbox[0][0])[tid] = (tid & 2) ? INT_MIN : INT_MAX;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();

    // four pixels per thread: patch q = 4 * warp + p of the 8 (across) x 4 (down) patches of the tile
    const size_t HW = (size_t)H * W;
    int px[4], py[4];
    bool live[4];
    float xv[4], yv[4], prob[4], count[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int q = 4 * warp + p;
        px[p] = blockIdx.x * HS_T + (q & 3) * 8 + (lane & 7);
        py[p] = blockIdx.y * HS_T + (q >> 2) * 4 + (lane >> 3);
        live[p] = px[p] < W && py[p] < H;
        xv[p] = live[p] ? xs[px[p]] : 0.f;
        yv[p] = live[p] ? ys[py[p]] : 0.f;
        const size_t o = (size_t)b * HW + (size_t)py[p] * W + px[p];
        if (!live[p]) { prob[p] = 0.f; count[p] = 1.f; }
        else if (flags & MP_HA_INIT) { prob[p] = prob0[o]; count[p] = 1.0f; }
        else { prob[p] = prob_acc[o]; count[p] = count_acc[o]; }
    }

    uint32_t phase = 0;
    for (int i = 0; i < n; ++i) {
        const int slot = i % 3;
        // ---- 1. coordinates and the window of this CTA's taps
        float ix[4], iy[4];
        int lo_x = INT_MAX, lo_y = INT_MAX, hi_x = INT_MIN, hi_y = INT_MIN;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            src_coord(As + 9 * i, xv[p], yv[p], W, H, ix[p], iy[p]);
            if (!live[p]) continue;
            const float fx = floorf(ix[p]), fy = floorf(iy[p]);
            if (fx >= -1.f && fx <= (float)W && fy >= -1.f && fy <= (float)H) {   // bilinear_setup's `any`
                lo_x = min(lo_x, (int)fx); hi_x = max(hi_x, (int)fx + 1);
                lo_y = min(lo_y, (int)fy); hi_y = max(hi_y, (int)fy + 1);
            }
            const float rx = rintf(ix[p]), ry = rintf(iy[p]);
            if (rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H) {
                lo_x = min(lo_x, (int)rx); hi_x = max(hi_x, (int)rx);
                lo_y = min(lo_y, (int)ry); hi_y = max(hi_y, (int)ry);
            }
        }
        lo_x = __reduce_min_sync(0xffffffffu, lo_x); lo_y = __reduce_min_sync(0xffffffffu, lo_y);
        hi_x = __reduce_max_sync(0xffffffffu, hi_x); hi_y = __reduce_max_sync(0xffffffffu, hi_y);
        if (lane == 0 && lo_x <= hi_x) {
            atomicMin(&bbox[slot][0], lo_x); atomicMin(&bbox[slot][1], lo_y);
            atomicMax(&bbox[slot][2], hi_x); atomicMax(&bbox[slot][3], hi_y);
        }
        __syncthreads();
        const int xlo = bbox[slot][0], ylo = bbox[slot][1];
        const int ext = max(bbox[slot][2] - xlo + 3, bbox[slot][3] - ylo) + 1;   // block-uniform; + 3: aligned window start
        const int xf = (xlo >> 2) << 2, xm = (xlo >> 4) << 4;                  // floor to 16 bytes (also for xlo = -1)
        if (tid == 0) {
            const int nslot = (i + 1) % 3;           // next sample's window accumulators (last read two samples ago)
            bbox[nslot][0] = INT_MAX; bbox[nslot][1] = INT_MAX; bbox[nslot][2] = INT_MIN; bbox[nslot][3] = INT_MIN;
        }
        if (xlo == INT_MAX) {                        // no tap of this sample lands in the image
            __syncthreads();                         // (the reset above before anybody's next atomics)
            continue;
        }
        const int v = ext <= 40 ? 0 : ext <= 56 ? 1 : ext <= 72 ? 2 : -1;   // HS_BOX
        const int bw = 40 + 16 * v, mw = 64 + 16 * v;                          // HS_BOX[v], HS_MBOX[v]
        // ---- 2. the window -> shared memory
        if (tid == 0) {
            if (v >= 0) {
                const uint32_t bytes = (uint32_t)(bw * bw * 4 * (TWO ? 2 : 1) + mw * bw);
                mbar_expect_tx(bar, bytes);          // (release: the reset above is visible to whoever sees this phase complete)
                // (constant indices: a dynamically indexed kernel parameter would be copied to local memory, where the
                // TMA unit cannot read a tensor map)
                const CUtensorMap *ma = v == 0 ? &maps.a[0] : v == 1 ? &maps.a[1] : &maps.a[2];
                const CUtensorMap *mb = v == 0 ? &maps.b[0] : v == 1 ? &maps.b[1] : &maps.b[2];
                const CUtensorMap *mm = v == 0 ? &maps.m[0] : v == 1 ? &maps.m[1] : &maps.m[2];
                tma_load_3d(smem_u32(sa), ma, bar, xf, ylo, i * B + b);
                if (TWO) tma_load_3d(smem_u32(sb), mb, bar, xf, ylo, i * B + b);
                tma_load_3d(smem_u32(smk), mm, bar, xm, ylo, i);
            }
        }
        if (v >= 0) {
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            __syncthreads();                         // no TMA hand-shake on this path: order the reset above
        }
        // ---- 3. taps
        const float *ga = pa + ((size_t)i * B + b) * HW;
        const float *gb = TWO ? pb + ((size_t)i * B + b) * HW : nullptr;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (!live[p]) continue;
            const float rx = rintf(ix[p]), ry = rintf(iy[p]);
            const bool near_in = rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H;
            const Bilinear q = bilinear_setup(ix[p], iy[p], W, H);
            float cs = 0.f, val = 0.f;
            if (v >= 0) {
                if (near_in) cs = (float)smk[((int)ry - ylo) * mw + ((int)rx - xm)];
                if (q.any) {
                    // the window holds zeros wherever a tap falls outside the image: adding 0 * w leaves the sum as
                    // skipping the tap does, so no per-tap validity test is needed
                    const int o = (q.y0 - ylo) * bw + (q.x0 - xf);
                    float t0, t1, t2, t3;
                    if (TWO) {
                        if (AGG == MP_AGG_PROD) {
                            t0 = __fmul_rn(sa[o], sb[o]); t1 = __fmul_rn(sa[o + 1], sb[o + 1]);
                            t2 = __fmul_rn(sa[o + bw], sb[o + bw]); t3 = __fmul_rn(sa[o + bw + 1], sb[o + bw + 1]);
                        } else {
                            t0 = __fadd_rn(sa[o], sb[o]); t1 = __fadd_rn(sa[o + 1], sb[o + 1]);
                            t2 = __fadd_rn(sa[o + bw], sb[o + bw]); t3 = __fadd_rn(sa[o + bw + 1], sb[o + bw + 1]);
                        }
                    } else {
                        t0 = sa[o]; t1 = sa[o + 1]; t2 = sa[o + bw]; t3 = sa[o + bw + 1];
                    }
                    val = __fadd_rn(val, __fmul_rn(t0, q.nw));
                    val = __fadd_rn(val, __fmul_rn(t1, q.ne));
                    val = __fadd_rn(val, __fmul_rn(t2, q.sw));
                    val = __fadd_rn(val, __fmul_rn(t3, q.se));
                }
            } else {   // window too large for shared memory: the direct gathers
                if (near_in) cs = (float)__ldg(masks + (size_t)i * HW + (int)ry * W + (int)rx);
                if (TWO) {
                    val = bilinear_apply(q, W, [&](int off) {
                        return AGG == MP_AGG_PROD ? __fmul_rn(__ldg(ga + off), __ldg(gb + off)) : __fadd_rn(__ldg(ga + off), __ldg(gb + off));
                    });
                } else {
                    val = bilinear_apply(q, W, [&](int off) { return __ldg(ga + off); });
                }
            }
            count[p] = __fadd_rn(count[p], cs);
            prob[p] = __fadd_rn(prob[p], __fmul_rn(val, cs));
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (!live[p]) continue;
        const size_t o = (size_t)b * HW + (size_t)py[p] * W + px[p];
        if (flags & MP_HA_FINISH) {
            float r = __fdiv_rn(prob[p], count[p]);
            if (AGG == MP_AGG_PROD) r = sqrtf(r);
            else if (AGG == MP_AGG_SUM) r = __fmul_rn(r, 0.5f);
            if (min_count > 0 && count[p] < (float)min_count) r = 0.f;
            out[o] = r;
        } else {
            prob_acc[o] = prob[p];
            count_acc[o] = count[p];
        }
    }
}

// (W, H, planes) tensor of `elem_bytes`-wide elements -> 3-D tiled map with a (box_w, box_h, 1) box, no swizzle
static int make_window_map(CUtensorMap *map, const void *ptr, CUtensorMapDataType dt, int elem_bytes, int W, int H, long long planes,
                           int box_w, int box_h) {
    PFN_encodeTiled enc = get_encode_fn();
    if (enc == nullptr) {
        set_error("mp_ha_aggregate_f32: cuTensorMapEncodeTiled not available from the driver");
        return MP_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)W * elem_bytes, (cuuint64_t)W * H * elem_bytes};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, dt, 3, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("mp_ha_aggregate_f32: cuTensorMapEncodeTiled failed with CUresult %d (W=%d H=%d planes=%lld box=%dx%d)", (int)r, W, H,
                  planes, box_w, box_h);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

}  // namespace mp

extern "C" int mp_warp_groups_f32(const float *src, int N, int group, int n_mats, int H, int W, const float *A,
                                  const float *xs, const float *ys, int mode, int padding, float *out,
                                  mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(N >= 0 && n_mats >= 0 && H > 0 && W > 0, "mp_warp_f32: bad shape");
    MP_CHECK_ARG(mode == MP_BILINEAR || mode == MP_NEAREST, "mp_warp_f32: bad mode %d", mode);
    MP_CHECK_ARG(padding == MP_PAD_ZEROS || padding == MP_PAD_REFLECTION, "mp_warp_f32: bad padding %d", padding);
    MP_CHECK_ARG(n_mats <= 65535, "mp_warp_f32: at most 65535 matrices per call");
    if (N == 0 || n_mats == 0) return MP_OK;
    MP_CHECK_ARG(group > 0 && N % group == 0, "mp_warp_groups_f32: group %d does not divide the %d planes", group, N);
    MP_CHECK_ARG(src && A && xs && ys && out, "mp_warp_f32: null pointer");
    dim3 grid((W + 31) / 32, (H + 7) / 8, n_mats), block(32, 8);
    cudaStream_t s = (cudaStream_t)stream;
    // many matrices over few planes (the adaptation loop's image warp): stage the planes in a gather array first
    static const bool no_gather = getenv("MP_WARP_NO_GATHER") != nullptr;   // tuning aid
    cudaTextureObject_t tex = 0;
    cudaArray_t arr = nullptr;
    if (!no_gather && mode == MP_BILINEAR && n_mats >= 4 && (long long)N * H < (1ll << 24) &&
        mp::gather_array_for(W, N * H, s, &tex, &arr) == 0) {
        MP_CUDA_OK(cudaMemcpy2DToArrayAsync(arr, 0, 0, src, (size_t)W * sizeof(float), (size_t)W * sizeof(float), (size_t)N * H,
                                            cudaMemcpyDeviceToDevice, s));
        mp::warp_kernel<true><<<grid, block, 0, s>>>(src, tex, N, group, H, W, A, xs, ys, mode, padding, out);
    } else {
        mp::warp_kernel<false><<<grid, block, 0, s>>>(src, 0, N, group, H, W, A, xs, ys, mode, padding, out);
    }
    MP_LAUNCH_OK_S("warp_kernel", s);
    return MP_OK;
}

extern "C" int mp_warp_f32(const float *src, int N, int n_mats, int H, int W, const float *A,
                           const float *xs, const float *ys, int mode, int padding, float *out,
                           mp_stream_t stream) {
    return mp_warp_groups_f32(src, N, N > 0 ? N : 1, n_mats, H, W, A, xs, ys, mode, padding, out, stream);
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
    cudaStream_t s = (cudaStream_t)stream;
    // staged path: TMA needs 16-byte aligned bases and row pitches (floats: W % 4, mask bytes: W % 16)
    static const bool env_staged = getenv("MP_HA_STAGED") != nullptr;   // tuning aid: the staged kernel for every call
    const bool two = probw_b != nullptr;
    const bool staged = (env_staged || (flags & MP_HA_STAGED)) && n > 0 && W % 16 == 0 && (((uintptr_t)probw_a | (uintptr_t)masks | (two ? (uintptr_t)probw_b : 0)) & 15) == 0;
    if (staged) {
        HaMaps maps;
        memset(&maps, 0, sizeof(maps));
        int rc;
        for (int v = 0; v < 3; ++v) {
            if ((rc = make_window_map(&maps.a[v], probw_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W, H, (long long)n * B, HS_BOX[v], HS_BOX[v])) != MP_OK) return rc;
            if (two && (rc = make_window_map(&maps.b[v], probw_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W, H, (long long)n * B, HS_BOX[v], HS_BOX[v])) != MP_OK) return rc;
            if ((rc = make_window_map(&maps.m[v], masks, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, W, H, n, HS_MBOX[v], HS_BOX[v])) != MP_OK) return rc;
        }
        dim3 grid((W + HS_T - 1) / HS_T, (H + HS_T - 1) / HS_T, B);
        const size_t smem = 128 + (size_t)HS_MAXBOX * HS_MAXBOX * 4 * (two ? 2 : 1) + (size_t)HS_MAXMBOX * HS_MAXBOX + (size_t)n * 9 * sizeof(float);
#define MP_HA_LAUNCH_STAGED(AGG, TWO)                                                                          \
    do {                                                                                                \
        auto k = ha_aggregate_staged_kernel<AGG, TWO>;                                                  \
        MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        k<<<grid, 256, smem, s>>>(maps, prob0, probw_a, probw_b, masks, Ainv, n, B, H, W, xs, ys, min_count, flags, \
                                  prob_acc, count_acc, out);                                            \
    } while (0)
        if (aggregation == MP_AGG_PROD) MP_HA_LAUNCH_STAGED(MP_AGG_PROD, true);
        else if (aggregation == MP_AGG_SUM) MP_HA_LAUNCH_STAGED(MP_AGG_SUM, true);
        else MP_HA_LAUNCH_STAGED(MP_AGG_NONE, false);
#undef MP_HA_LAUNCH_STAGED
        MP_LAUNCH_OK_S("ha_aggregate_kernel", s);
        return MP_OK;
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8, B), block(32, 8);
    const size_t smem = (size_t)n * 9 * sizeof(float);
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
