// The elementwise glue between the backbone's convolutions (row 3 of the hot path, MultiPoint.forward,
// multipoint/models/MultiPoint.py:61-90,99-135): after every 3x3 convolution the reference runs
//     ReLU -> BatchNorm2d (eval) [-> MaxPool2d(2,2)] -> ReflectionPad2d(1) | ZeroPad2d(1)
// (or BatchNorm -> ReLU with bn_first) as separate full-tensor passes: at 512x640 with 64 channels and
// 128 images that is a 10.7 GB activation read and written three to four times per layer, ~20 ms of
// the 156 ms step.  This kernel does the whole chain in one pass: read the convolution output once,
// write the (pooled, padded) input of the next convolution once.  The convolutions stay in cuDNN.
//
// The convolution's bias goes in as well (conv_bias): cuDNN's FFT / Winograd algorithms do not fuse it, so
// torch adds it with one more full-tensor elementwise kernel per layer (24 ms per step in the profile).
//
// HBM-bound: 4 B read per input element + 4 B written per output element.
// BatchNorm in eval mode is the per-channel affine y = x * scale + shift with
// scale = weight / sqrt(running_var + eps), shift = bias - running_mean * scale, folded on the host;
// that differs from torch's (x - mean) * invstd * weight + bias by <= 2 ulp.
#include "mp_common.cuh"

namespace mp {

constexpr int AF_THREADS = 256;
constexpr int AF_ROWS = 8;  // output rows per CTA; a thread owns one column of them, so it has 8 (16 with pooling) loads in flight

__device__ __forceinline__ float af_apply(float x, float cb, float scale, float shift, bool bn_first) {
    x = __fadd_rn(x, cb);  // the convolution bias, rounded like torch's separate add
    if (bn_first) return fmaxf(fmaf(x, scale, shift), 0.f);
    return fmaf(fmaxf(x, 0.f), scale, shift);
}

template <bool POOL>
__global__ void __launch_bounds__(AF_THREADS)
relu_bn_pad_kernel(const float *__restrict__ x, float *__restrict__ out, const float *__restrict__ conv_bias,
                   const float *__restrict__ scale, const float *__restrict__ shift, int C, int H, int W, int pad,
                   int reflect, int bn_first) {
    const int plane = blockIdx.y;
    const int c = plane % C;
    const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
    const int Hp = Ho + 2 * pad, Wp = Wo + 2 * pad;
    const float sc = scale[c], sh = shift[c], cb = conv_bias ? conv_bias[c] : 0.f;
    const float *src = x + (size_t)plane * H * W;
    float *dst = out + (size_t)plane * Hp * Wp;
    const int y0 = blockIdx.x * AF_ROWS;
    // source row of each output row (reflection: -1 -> 1, Ho -> Ho-2); -1 = a zero-padded row
    int ysrc[AF_ROWS];
#pragma unroll
    for (int r = 0; r < AF_ROWS; ++r) {
        int ys = y0 + r - pad;
        if (ys < 0) ys = reflect ? -ys : -1;
        else if (ys >= Ho) ys = reflect ? 2 * Ho - 2 - ys : -1;
        ysrc[r] = (y0 + r < Hp) ? ys : -2;  // -2 = row does not exist
    }
    for (int xo = threadIdx.x; xo < Wp; xo += AF_THREADS) {
        int xs = xo - pad;
        if (xs < 0) xs = reflect ? -xs : -1;
        else if (xs >= Wo) xs = reflect ? 2 * Wo - 2 - xs : -1;
        float v[AF_ROWS];
        if (POOL) {
            float2 a[AF_ROWS], b[AF_ROWS];
#pragma unroll
            for (int r = 0; r < AF_ROWS; ++r) {
                a[r] = b[r] = make_float2(0.f, 0.f);
                if (ysrc[r] >= 0 && xs >= 0) {
                    const float *p = src + (size_t)(2 * ysrc[r]) * W + 2 * xs;
                    a[r] = *reinterpret_cast<const float2 *>(p);
                    b[r] = *reinterpret_cast<const float2 *>(p + W);
                }
            }
#pragma unroll
            for (int r = 0; r < AF_ROWS; ++r)
                v[r] = fmaxf(fmaxf(af_apply(a[r].x, cb, sc, sh, bn_first), af_apply(a[r].y, cb, sc, sh, bn_first)),
                             fmaxf(af_apply(b[r].x, cb, sc, sh, bn_first), af_apply(b[r].y, cb, sc, sh, bn_first)));
        } else {
            float t[AF_ROWS];
#pragma unroll
            for (int r = 0; r < AF_ROWS; ++r) t[r] = (ysrc[r] >= 0 && xs >= 0) ? src[(size_t)ysrc[r] * W + xs] : 0.f;
#pragma unroll
            for (int r = 0; r < AF_ROWS; ++r) v[r] = af_apply(t[r], cb, sc, sh, bn_first);
        }
#pragma unroll
        for (int r = 0; r < AF_ROWS; ++r) {
            if (ysrc[r] == -2) continue;
            dst[(size_t)(y0 + r) * Wp + xo] = (ysrc[r] >= 0 && xs >= 0) ? v[r] : 0.f;  // ZeroPad2d outside
        }
    }
}

// W % 4 == 0 (every layer of the shipped networks): one 128-bit load per thread and row instead of four 32-bit ones.
// Item = (output row, quad of four input columns); with pooling a quad of input columns gives two output columns
// and the thread reads the quad from both input rows.  The padded border columns are written by the threads that
// own the first / last quad (reflection: column -1 mirrors column 1, column Wo mirrors column Wo-2).
template <bool POOL>
__global__ void __launch_bounds__(AF_THREADS)
relu_bn_pad_vec_kernel(const float *__restrict__ x, float *__restrict__ out, const float *__restrict__ conv_bias,
                       const float *__restrict__ scale, const float *__restrict__ shift, int C, int H, int W, int pad,
                       int reflect, int bn_first, int rows_per_cta) {
    const int plane = blockIdx.y;
    const int c = plane % C;
    const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
    const int Hp = Ho + 2 * pad, Wp = Wo + 2 * pad;
    const int Q = W / 4;                         // quads per input row
    const float sc = scale[c], sh = shift[c], cb = conv_bias ? conv_bias[c] : 0.f;
    const float *src = x + (size_t)plane * H * W;
    float *dst = out + (size_t)plane * Hp * Wp;
    const int y0 = blockIdx.x * rows_per_cta;
    const int y1 = min(Hp, y0 + rows_per_cta);
    const int items = (y1 - y0) * Q;
    constexpr int UNROLL = 4;
    for (int it0 = threadIdx.x; it0 < items; it0 += AF_THREADS * UNROLL) {
        float4 a[UNROLL], b[UNROLL];
        int yo[UNROLL], q[UNROLL];
        bool live[UNROLL], zero_row[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int it = it0 + u * AF_THREADS;
            live[u] = it < items;
            const int r = live[u] ? it / Q : 0;
            q[u] = live[u] ? it - r * Q : 0;
            yo[u] = y0 + r;
            int ys = yo[u] - pad;
            zero_row[u] = false;
            if (ys < 0) { if (reflect) ys = -ys; else { zero_row[u] = true; ys = 0; } }
            else if (ys >= Ho) { if (reflect) ys = 2 * Ho - 2 - ys; else { zero_row[u] = true; ys = 0; } }
            a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live[u] && !zero_row[u]) {
                const float *p = src + (size_t)(POOL ? 2 * ys : ys) * W + 4 * q[u];
                a[u] = *reinterpret_cast<const float4 *>(p);
                if (POOL) b[u] = *reinterpret_cast<const float4 *>(p + W);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!live[u]) continue;
            float *row = dst + (size_t)yo[u] * Wp + pad;
            if (POOL) {
                float v0 = fmaxf(fmaxf(af_apply(a[u].x, cb, sc, sh, bn_first), af_apply(a[u].y, cb, sc, sh, bn_first)),
                                 fmaxf(af_apply(b[u].x, cb, sc, sh, bn_first), af_apply(b[u].y, cb, sc, sh, bn_first)));
                float v1 = fmaxf(fmaxf(af_apply(a[u].z, cb, sc, sh, bn_first), af_apply(a[u].w, cb, sc, sh, bn_first)),
                                 fmaxf(af_apply(b[u].z, cb, sc, sh, bn_first), af_apply(b[u].w, cb, sc, sh, bn_first)));
                if (zero_row[u]) v0 = v1 = 0.f;
                row[2 * q[u]] = v0;
                row[2 * q[u] + 1] = v1;
                if (pad) {
                    if (q[u] == 0) row[-1] = reflect ? v1 : 0.f;                 // column -1 mirrors pooled column 1
                    if (q[u] == Q - 1) row[Wo] = reflect ? v0 : 0.f;             // column Wo mirrors pooled column Wo-2
                }
            } else {
                float v[4] = {af_apply(a[u].x, cb, sc, sh, bn_first), af_apply(a[u].y, cb, sc, sh, bn_first),
                              af_apply(a[u].z, cb, sc, sh, bn_first), af_apply(a[u].w, cb, sc, sh, bn_first)};
                if (zero_row[u]) v[0] = v[1] = v[2] = v[3] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) row[4 * q[u] + j] = v[j];
                if (pad) {
                    if (q[u] == 0) row[-1] = reflect ? v[1] : 0.f;
                    if (q[u] == Q - 1) row[Wo] = reflect ? v[2] : 0.f;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// First layer of the encoders: Conv2d(1 -> C, 3x3) on the padded image, + bias, ReLU, eval BatchNorm and the
// padding of the next convolution in one kernel.  With a single input channel the convolution is 9 MACs per
// output value and entirely bound by writing the C-channel activation (5.4 GB per 64 images of 512x640), so
// there is nothing for a GEMM-shaped library kernel to win: cuDNN spends ~2.2 ms on it and the glue pass that
// follows another 1.65 ms, this kernel writes the padded activation once (~0.9 ms).
// One CTA per (output row, image); a thread owns up to four columns T apart (coalesced stores for every
// channel) and keeps their 3x3 input patches in registers; weights and per-channel constants in shared memory.
constexpr int C1_THREADS = 192;
constexpr int C1_PX = 4;

__global__ void __launch_bounds__(C1_THREADS)
conv1_relu_bn_pad_kernel(const float *__restrict__ img, float *__restrict__ out, const float *__restrict__ weight,
                         const float *__restrict__ conv_bias, const float *__restrict__ scale,
                         const float *__restrict__ shift, int C, int H, int W, int in_reflect, int pad, int out_reflect,
                         int bn_first) {
    extern __shared__ __align__(16) float c1_smem[];  // [C][12]: 9 weights, conv bias, scale, shift
    for (int i = threadIdx.x; i < C * 12; i += C1_THREADS) {
        const int c = i / 12, k = i - c * 12;
        c1_smem[i] = k < 9 ? weight[c * 9 + k] : (k == 9 ? (conv_bias ? conv_bias[c] : 0.f) : (k == 10 ? scale[c] : shift[c]));
    }
    __syncthreads();
    const int b = blockIdx.y, yo = blockIdx.x;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const float *src = img + (size_t)b * H * W;
    int ys = yo - pad;
    bool zero_out_row = false;
    if (ys < 0) { if (out_reflect) ys = -ys; else zero_out_row = true; }
    else if (ys >= H) { if (out_reflect) ys = 2 * H - 2 - ys; else zero_out_row = true; }
    float patch[C1_PX][9];
    bool live[C1_PX], zero_px[C1_PX];
#pragma unroll
    for (int j = 0; j < C1_PX; ++j) {
        const int xo = threadIdx.x + j * C1_THREADS;
        live[j] = xo < Wp;
        int xs = xo - pad;
        zero_px[j] = zero_out_row;
        if (xs < 0) { if (out_reflect) xs = -xs; else zero_px[j] = true; }
        else if (xs >= W) { if (out_reflect) xs = 2 * W - 2 - xs; else zero_px[j] = true; }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            int iy = ys + k / 3 - 1, ix = xs + k % 3 - 1;
            bool inside = true;  // the input's own padding (ReflectionPad2d(1) / ZeroPad2d(1) in front of the convolution)
            if (iy < 0) { if (in_reflect) iy = -iy; else inside = false; }
            else if (iy >= H) { if (in_reflect) iy = 2 * H - 2 - iy; else inside = false; }
            if (ix < 0) { if (in_reflect) ix = -ix; else inside = false; }
            else if (ix >= W) { if (in_reflect) ix = 2 * W - 2 - ix; else inside = false; }
            patch[j][k] = (live[j] && !zero_px[j] && inside) ? __ldg(src + (size_t)iy * W + ix) : 0.f;
        }
    }
    float *dst = out + ((size_t)b * C * Hp + yo) * Wp;
    // the four columns as two packed pairs: fma.rn.f32x2 does two of the 9-tap MACs per instruction (the kernel is
    // bound by instruction issue, not by the 5.4 GB it writes); per-lane results are the same IEEE fmas
    float2 p2[C1_PX / 2][9];
#pragma unroll
    for (int h = 0; h < C1_PX / 2; ++h)
#pragma unroll
        for (int k = 0; k < 9; ++k) p2[h][k] = make_float2(patch[2 * h][k], patch[2 * h + 1][k]);
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
        const float4 w0 = *reinterpret_cast<const float4 *>(c1_smem + c * 12);
        const float4 w1 = *reinterpret_cast<const float4 *>(c1_smem + c * 12 + 4);
        const float4 w2 = *reinterpret_cast<const float4 *>(c1_smem + c * 12 + 8);  // w[8], conv bias, scale, shift
        const float wk[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x};
#pragma unroll
        for (int h = 0; h < C1_PX / 2; ++h) {
            float2 acc = make_float2(p2[h][0].x * wk[0], p2[h][0].y * wk[0]);
#pragma unroll
            for (int k = 1; k < 9; ++k) acc = __ffma2_rn(p2[h][k], make_float2(wk[k], wk[k]), acc);
            float v0 = af_apply(acc.x, w2.y, w2.z, w2.w, bn_first);
            float v1 = af_apply(acc.y, w2.y, w2.z, w2.w, bn_first);
            if (zero_px[2 * h]) v0 = 0.f;
            if (zero_px[2 * h + 1]) v1 = 0.f;
            float *d = dst + (size_t)c * Hp * Wp + threadIdx.x;
            if (live[2 * h]) d[(2 * h) * C1_THREADS] = v0;
            if (live[2 * h + 1]) d[(2 * h + 1) * C1_THREADS] = v1;
        }
    }
}

}  // namespace mp

extern "C" int mp_relu_bn_pad_f32(const float *x, int B, int C, int H, int W, const float *conv_bias, const float *scale,
                                  const float *shift, int bn_first, int pool, int pad, int reflect, float *out,
                                  mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0, "mp_relu_bn_pad_f32: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    MP_CHECK_ARG(pad == 0 || pad == 1, "mp_relu_bn_pad_f32: pad must be 0 or 1");
    MP_CHECK_ARG(!pool || (H % 2 == 0 && W % 2 == 0), "mp_relu_bn_pad_f32: pooling needs even H and W (MaxPool2d(2,2) would drop a row)");
    const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
    MP_CHECK_ARG(!(pad && reflect) || (Ho >= 2 && Wo >= 2), "mp_relu_bn_pad_f32: reflection padding needs at least 2 pixels");
    if (B == 0) return MP_OK;
    MP_CHECK_ARG(x && out && scale && shift, "mp_relu_bn_pad_f32: null pointer");
    MP_CHECK_ARG((long long)B * C <= 65535, "mp_relu_bn_pad_f32: B*C must be <= 65535 per call");
    MP_CHECK_ARG(!pool || (((uintptr_t)x & 7) == 0), "mp_relu_bn_pad_f32: input must be 8-byte aligned");
    const int Hp = Ho + 2 * pad;
    if (W % 4 == 0 && (((uintptr_t)x & 15) == 0) && Wo >= 2) {
        // about 4096 quads per CTA: 16 rows at W = 640, more at the lower resolutions
        int rows = 4096 / (W / 4);
        if (rows < 4) rows = 4;
        if (rows > Hp) rows = Hp;
        dim3 vgrid((unsigned)((Hp + rows - 1) / rows), (unsigned)(B * C));
        if (pool)
            mp::relu_bn_pad_vec_kernel<true><<<vgrid, mp::AF_THREADS, 0, (cudaStream_t)stream>>>(x, out, conv_bias, scale, shift, C, H, W, pad, reflect, bn_first, rows);
        else
            mp::relu_bn_pad_vec_kernel<false><<<vgrid, mp::AF_THREADS, 0, (cudaStream_t)stream>>>(x, out, conv_bias, scale, shift, C, H, W, pad, reflect, bn_first, rows);
        MP_LAUNCH_OK_S("relu_bn_pad_kernel", (cudaStream_t)stream);
        return MP_OK;
    }
    dim3 grid((unsigned)((Hp + mp::AF_ROWS - 1) / mp::AF_ROWS), (unsigned)(B * C));
    if (pool)
        mp::relu_bn_pad_kernel<true><<<grid, mp::AF_THREADS, 0, (cudaStream_t)stream>>>(x, out, conv_bias, scale, shift, C, H, W, pad, reflect, bn_first);
    else
        mp::relu_bn_pad_kernel<false><<<grid, mp::AF_THREADS, 0, (cudaStream_t)stream>>>(x, out, conv_bias, scale, shift, C, H, W, pad, reflect, bn_first);
    MP_LAUNCH_OK_S("relu_bn_pad_kernel", (cudaStream_t)stream);
    return MP_OK;
}

extern "C" int mp_conv1_relu_bn_pad_f32(const float *image, int B, int H, int W, const float *weight, const float *conv_bias,
                                        const float *scale, const float *shift, int C, int bn_first, int in_reflect, int pad,
                                        int out_reflect, float *out, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(B >= 0 && C > 0 && H >= 2 && W >= 2, "mp_conv1_relu_bn_pad_f32: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    MP_CHECK_ARG(pad == 0 || pad == 1, "mp_conv1_relu_bn_pad_f32: pad must be 0 or 1");
    MP_CHECK_ARG(C <= 1024, "mp_conv1_relu_bn_pad_f32: at most 1024 output channels");
    MP_CHECK_ARG(W + 2 * pad <= mp::C1_THREADS * mp::C1_PX, "mp_conv1_relu_bn_pad_f32: image wider than %d", mp::C1_THREADS * mp::C1_PX - 2);
    if (B == 0) return MP_OK;
    MP_CHECK_ARG(image && weight && scale && shift && out, "mp_conv1_relu_bn_pad_f32: null pointer");
    MP_CHECK_ARG(B <= 65535, "mp_conv1_relu_bn_pad_f32: at most 65535 images per call");
    dim3 grid((unsigned)(H + 2 * pad), (unsigned)B);
    const size_t smem = (size_t)C * 12 * sizeof(float);
    mp::conv1_relu_bn_pad_kernel<<<grid, mp::C1_THREADS, smem, (cudaStream_t)stream>>>(image, out, weight, conv_bias, scale, shift, C, H, W,
                                                                                      in_reflect, pad, out_reflect, bn_first);
    MP_LAUNCH_OK_S("conv1_relu_bn_pad_kernel", (cudaStream_t)stream);
    return MP_OK;
}
