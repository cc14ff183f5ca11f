// compute_valid_mask (multipoint/utils/homographies.py:375-402) for a
// whole batch of homographies in one launch, instead of 99 host-side cv2.warpPerspective + cv2.erode
// calls per adaptation batch (6.5 ms each on the host).
//
//   raw[y,x]  = 1 iff the nearest-neighbour source pixel of (x,y) under the INVERTED homography lies
//               inside the image (cv2.warpPerspective of ones, INTER_NEAREST, constant border 0)
//   mask[y,x] = AND of raw over the (2r+1)^2 window; outside the image counts as 1 (cv2.erode's
//               default border) or as 0 when mask_border adds the one-pixel zero frame.
//
// The coordinate arithmetic is OpenCV's, in double and in its operation order (block origin xb,
// X0 = M0*xb + M1*y + M2, W = 1/(W0 + M6*x1), fX = (X0 + M0*x1)*W, round half to even), written
// with explicit __dmul_rn/__dadd_rn so no FMA contraction changes a boundary pixel; the oracle's
// restatement of the same is pinned bit-exactly against cv2 (tests/golden/valid_mask.npz).
//
// One CTA owns VM_TH output rows of one mask: raw bits for VM_TH + 2r rows go to shared memory as
// one bit per pixel (warp ballot), the erosion is funnel-shift ANDs along the row and ANDs down the
// column, and the result leaves as bytes.  The ~35 double operations (one of them a division) per pixel made the kernel
// the largest of the adaptation batch (0.33 ms for 99 masks, fp64 pipe + the division's instruction sequence); a pixel
// is now classified in fp32 first, with an explicit error bound, and only falls through to the exact sequence when its
// source position is within that bound of the image border.
#include "mp_common.cuh"

namespace mp {

constexpr int VM_TH = 32;        // output rows per CTA (the 2r halo rows are recomputed: 31 % extra at r = 5; 16 rows cost 62 %)
constexpr int VM_THREADS = 256;
constexpr int VM_MAX_R = 31;     // erosion radius limit (one funnel shift per offset)

__device__ __forceinline__ int vm_round_sat(double v) {
    // max(INT_MIN, min(INT_MAX, v)) with OpenCV's std::min/std::max argument order: NaN -> INT_MAX
    if (!(v < 2147483647.0)) v = 2147483647.0;
    if (v < -2147483648.0) v = -2147483648.0;
    return __double2int_rn(v);
}

__global__ void __launch_bounds__(VM_THREADS)
valid_mask_kernel(const double *__restrict__ Minv, int H, int W, int r, int mask_border, int bw0,
                  uint8_t *__restrict__ mask) {
    extern __shared__ uint32_t bits[];  // [(VM_TH + 2r)][nw + 2] raw, then [(VM_TH + 2r)][nw] row-eroded
    const int nw = (W + 31) >> 5;
    const int ld = nw + 2;              // one border word on each side
    const int rows = VM_TH + 2 * r;
    uint32_t *raw = bits;
    uint32_t *hor = bits + rows * ld;
    const int n = blockIdx.y;
    const int y0 = blockIdx.x * VM_TH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t border = mask_border ? 0u : 0xffffffffu;

    double M[9];
    float Mf[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { M[i] = Minv[(size_t)n * 9 + i]; Mf[i] = (float)M[i]; }

    // phase 1: raw bits, one warp per row, word after word
    // Fast classification in fp32 with a per-pixel error bound: only pixels whose source position lies within that bound of
    // the image border (a curve: ~1 % of the warps) pay for OpenCV's exact double sequence.  With c = (size - 1) / 2:
    //   inside  <=> |p - c| <= size / 2 (both axes),  so  surely inside  <=> |p - c| + e <= size / 2
    //                                                     surely outside <=> |p - c| - e >  size / 2 (either axis)
    // e >= 8x the fp32 error (matrix entries rounded to fp32, three fused terms per sum, approximate reciprocal, product:
    // <= 4e-7 * (sum of magnitudes of the numerator + |quotient| * those of the denominator) / |denominator|) + 1e-4 px.
    const float a6 = fabsf(Mf[6]), a0 = fabsf(Mf[0]), a3 = fabsf(Mf[3]);
    const float cxm = 0.5f * (float)(W - 1), cym = 0.5f * (float)(H - 1), hw_ = 0.5f * (float)W, hh_ = 0.5f * (float)H;
    for (int rr = warp; rr < rows; rr += VM_THREADS / 32) {
        const int y = y0 - r + rr;
        uint32_t *dst = raw + rr * ld + 1;
        if (y < 0 || y >= H) {
            for (int w = lane; w < nw; w += 32) dst[w] = border;
            continue;
        }
        const float fy = (float)y;
        const float cden = fmaf(Mf[7], fy, Mf[8]), csd = fmaf(fabsf(Mf[7]), fy, fabsf(Mf[8]));
        const float cnx = fmaf(Mf[1], fy, Mf[2]), csx = fmaf(fabsf(Mf[1]), fy, fabsf(Mf[2]));
        const float cny = fmaf(Mf[4], fy, Mf[5]), csy = fmaf(fabsf(Mf[4]), fy, fabsf(Mf[5]));
        for (int w = 0; w < nw; ++w) {
            const int x = 32 * w + lane;
            bool in = false;
            if (x < W) {
                const float fx = (float)x;
                const float den = fmaf(Mf[6], fx, cden), sd = fmaf(a6, fx, csd);
                const float rd = __fdividef(1.0f, den), ard = fabsf(rd);
                const float px = fmaf(Mf[0], fx, cnx) * rd, py = fmaf(Mf[3], fx, cny) * rd;
                const float ex = fmaf(3.2e-6f * ard, fmaf(fabsf(px), sd, fmaf(a0, fx, csx)), 1e-4f);
                const float ey = fmaf(3.2e-6f * ard, fmaf(fabsf(py), sd, fmaf(a3, fx, csy)), 1e-4f);
                const float dx = fabsf(px - cxm), dyv = fabsf(py - cym);
                const bool sane = fabsf(den) > 1e-3f * sd && dx < 1e6f && dyv < 1e6f;   // false for NaN / infinities
                const bool fast_in = sane && dx + ex <= hw_ && dyv + ey <= hh_;
                const bool fast_out = sane && (dx - ex > hw_ || dyv - ey > hh_);
                in = fast_in;
                if (!fast_in && !fast_out) {
                    const int xb = (x / bw0) * bw0, x1 = x - xb;
                    const double dy = (double)y, dxb = (double)xb, dx1 = (double)x1;
                    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], dxb), __dmul_rn(M[1], dy)), M[2]);
                    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], dxb), __dmul_rn(M[4], dy)), M[5]);
                    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], dxb), __dmul_rn(M[7], dy)), M[8]);
                    double wv = __dadd_rn(W0, __dmul_rn(M[6], dx1));
                    wv = (wv != 0.0) ? __ddiv_rn(1.0, wv) : 0.0;
                    const int sx = vm_round_sat(__dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], dx1)), wv));
                    const int sy = vm_round_sat(__dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], dx1)), wv));
                    in = sx >= 0 && sx < W && sy >= 0 && sy < H;
                }
            }
            uint32_t word = __ballot_sync(0xffffffffu, in);
            const int tail = W - 32 * w;           // pixels of this word that exist
            if (tail < 32) word = (word & ((1u << tail) - 1u)) | (border & ~((1u << tail) - 1u));
            if (lane == 0) dst[w] = word;
        }
    }
    for (int rr = threadIdx.x; rr < rows; rr += VM_THREADS) {
        raw[rr * ld] = border;
        raw[rr * ld + 1 + nw] = border;
    }
    __syncthreads();

    // phase 2: erosion along the row
    for (int t = threadIdx.x; t < rows * nw; t += VM_THREADS) {
        const int rr = t / nw, w = t - rr * nw;
        const uint32_t prev = raw[rr * ld + w], cur = raw[rr * ld + 1 + w], next = raw[rr * ld + 2 + w];
        uint32_t acc = cur;
        for (int d = 1; d <= r; ++d) {
            acc &= __funnelshift_r(cur, next, d);       // bit x <- raw[x + d]
            acc &= __funnelshift_l(prev, cur, d);       // bit x <- raw[x - d]
        }
        hor[rr * nw + w] = acc;
    }
    __syncthreads();

    // phase 3: erosion down the column, bytes out (one thread per 4 pixels)
    const int quads = (W + 3) >> 2;
    const bool vec = (W & 3) == 0;
    for (int t = threadIdx.x; t < VM_TH * quads; t += VM_THREADS) {
        const int ry = t / quads, q = t - ry * quads;
        const int y = y0 + ry;
        if (y >= H) break;
        const int x = 4 * q;
        const int w = x >> 5, sh = x & 31;
        uint32_t acc = 0xfu;
        for (int k = 0; k <= 2 * r; ++k) acc &= hor[(ry + k) * nw + w] >> sh;
        uint8_t *dst = mask + ((size_t)n * H + y) * W + x;
        if (vec) {
            uchar4 o;
            o.x = acc & 1u; o.y = (acc >> 1) & 1u; o.z = (acc >> 2) & 1u; o.w = (acc >> 3) & 1u;
            *reinterpret_cast<uchar4 *>(dst) = o;
        } else {
            for (int j = 0; j < 4 && x + j < W; ++j) dst[j] = (acc >> j) & 1u;
        }
    }
}

}  // namespace mp

extern "C" int mp_valid_mask_u8(const double *Minv, int n, int H, int W, int erosion_radius, int mask_border,
                                uint8_t *mask, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(n >= 0 && H > 0 && W > 0, "mp_valid_mask_u8: bad shape n=%d H=%d W=%d", n, H, W);
    MP_CHECK_ARG(H < 32768 && W < 32768, "mp_valid_mask_u8: image larger than OpenCV's 16-bit remap coordinates");
    if (n == 0) return MP_OK;
    MP_CHECK_ARG(Minv && mask, "mp_valid_mask_u8: null pointer");
    MP_CHECK_ARG(n <= 65535, "mp_valid_mask_u8: at most 65535 masks per call");
    MP_CHECK_ARG(((uintptr_t)mask & 3) == 0 || (W & 3) != 0, "mp_valid_mask_u8: mask must be 4-byte aligned");
    const int r = erosion_radius > 0 ? erosion_radius : 0;
    if (r > mp::VM_MAX_R) {
        mp::set_error("mp_valid_mask_u8: erosion_radius %d > %d is not supported", r, mp::VM_MAX_R);
        return MP_ERR_UNSUPPORTED;
    }
    // OpenCV's destination block width (WarpPerspectiveInvoker): 1024 / min(16, H), capped at W
    const int bh0 = H < 16 ? H : 16;
    int bw0 = 1024 / bh0;
    if (bw0 > W) bw0 = W;
    const int nw = (W + 31) / 32, rows = mp::VM_TH + 2 * r;
    const size_t smem = (size_t)rows * (2 * nw + 2) * sizeof(uint32_t);
    if (smem > 200 * 1024) {
        mp::set_error("mp_valid_mask_u8: W=%d with erosion_radius=%d needs %zu B of shared memory", W, r, smem);
        return MP_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        MP_CUDA_OK(cudaFuncSetAttribute(mp::valid_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((H + mp::VM_TH - 1) / mp::VM_TH), (unsigned)n);
    mp::valid_mask_kernel<<<grid, mp::VM_THREADS, smem, (cudaStream_t)stream>>>(Minv, H, W, r, mask_border ? 1 : 0, bw0, mask);
    MP_LAUNCH_OK_S("valid_mask_kernel", (cudaStream_t)stream);
    return MP_OK;
}
