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
