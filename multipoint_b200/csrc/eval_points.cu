// the point geometry inside the evaluation loops, batched over the samples of a
// batch so the `-e` paths stay on the device:
//   compute_repeatability_multispectral  multipoint/utils/evaluation.py:148-199
//   compute_descriptor_metrics           multipoint/utils/evaluation.py:253-358
// Three small kernels (N <= a few thousand points per sample; nothing here is near a roofline):
//   warp_keypoints_kernel   warp_keypoints (homographies.py:331-346) = cv2.perspectiveTransform in double
//   points_min_dist2_kernel filter_points + the N1 x N2 distance matrix + row minimum (:176-197), exact int64
//   points_correct_kernel   the "correct match" matrix (:294-298) reduced to what the loop reads from it:
//                           row-any (:301-302) and the entries at the matcher's pairs (:306-315)
// All take P problems with a fixed capacity per problem and device-side counts (the layout the
// keypoint kernels produce), so no host synchronisation is needed between the stages.
#include "mp_common.cuh"

namespace mp {

constexpr int EP_THREADS = 256;

__global__ void warp_keypoints_kernel(const int64_t *__restrict__ kp, const int *__restrict__ counts, int cap,
                                      const double *__restrict__ Hm, double *__restrict__ out_f64,
                                      int64_t *__restrict__ out_i64) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = counts ? min(counts[p], cap) : cap;
    if (i >= n) return;
    const double *m = Hm + (size_t)p * 9;
    const size_t o = ((size_t)p * cap + i) * 2;
    const double y = (double)kp[o], x = (double)kp[o + 1];
    // OpenCV's perspectiveTransform_64f as built with FMA contraction (measured against cv2): fma(x, a, y*b) + c
    double w = __dadd_rn(__fma_rn(x, m[6], __dmul_rn(y, m[7])), m[8]);
    double xo = 0., yo = 0.;
    if (fabs(w) > 1.1920928955078125e-07) {
        w = __ddiv_rn(1.0, w);
        xo = __dmul_rn(__dadd_rn(__fma_rn(x, m[0], __dmul_rn(y, m[1])), m[2]), w);
        yo = __dmul_rn(__dadd_rn(__fma_rn(x, m[3], __dmul_rn(y, m[4])), m[5]), w);
    }
    if (out_f64) { out_f64[o] = yo; out_f64[o + 1] = xo; }
    if (out_i64) { out_i64[o] = (int64_t)yo; out_i64[o + 1] = (int64_t)xo; }  // .astype(int): truncation
}

__global__ void __launch_bounds__(EP_THREADS)
points_min_dist2_kernel(const int64_t *__restrict__ q, const int *__restrict__ nq, int capq,
                        const int64_t *__restrict__ t, const int *__restrict__ nt, int capt, int H, int W,
                        int64_t *__restrict__ min_d2) {
    __shared__ longlong2 tile[EP_THREADS];
    const int p = blockIdx.y;
    const int i = blockIdx.x * EP_THREADS + threadIdx.x;
    const int n_q = nq ? min(nq[p], capq) : capq;
    const int n_t = nt ? min(nt[p], capt) : capt;
    if (blockIdx.x * EP_THREADS >= n_q) return;
    const bool live = i < n_q;
    int64_t qy = 0, qx = 0;
    if (live) {
        qy = q[((size_t)p * capq + i) * 2];
        qx = q[((size_t)p * capq + i) * 2 + 1];
    }
    const bool inside = live && qy >= 0 && qx >= 0 && qy < H && qx < W;  // filter_points
    int64_t best = INT64_MAX;
    const longlong2 *tp = reinterpret_cast<const longlong2 *>(t) + (size_t)p * capt;
    for (int j0 = 0; j0 < n_t; j0 += EP_THREADS) {
        __syncthreads();
        if (j0 + threadIdx.x < n_t) tile[threadIdx.x] = tp[j0 + threadIdx.x];
        __syncthreads();
        const int m = min(EP_THREADS, n_t - j0);
        if (inside) {
            for (int j = 0; j < m; ++j) {
                const int64_t dy = qy - tile[j].x, dx = qx - tile[j].y;
                best = min(best, dy * dy + dx * dx);
            }
        }
    }
    if (live) min_d2[(size_t)p * capq + i] = inside ? best : -1;
}

__device__ __forceinline__ bool ep_correct(double qy, double qx, longlong2 t, float thr) {
    const float dy = __double2float_rn(__dsub_rn(qy, (double)t.x)), dx = __double2float_rn(__dsub_rn(qx, (double)t.y));
    return __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx))) <= thr;
}

__global__ void __launch_bounds__(EP_THREADS)
points_correct_kernel(const double *__restrict__ qw, const int *__restrict__ nq, int capq,
                      const int64_t *__restrict__ t, const int *__restrict__ nt, int capt, float thr,
                      uint8_t *__restrict__ row_any) {
    __shared__ longlong2 tile[EP_THREADS];
    const int p = blockIdx.y;
    const int i = blockIdx.x * EP_THREADS + threadIdx.x;
    const int n_q = nq ? min(nq[p], capq) : capq;
    const int n_t = nt ? min(nt[p], capt) : capt;
    if (blockIdx.x * EP_THREADS >= n_q) return;
    const bool live = i < n_q;
    double qy = 0., qx = 0.;
    if (live) {
        qy = qw[((size_t)p * capq + i) * 2];
        qx = qw[((size_t)p * capq + i) * 2 + 1];
    }
    bool any = false;
    const longlong2 *tp = reinterpret_cast<const longlong2 *>(t) + (size_t)p * capt;
    for (int j0 = 0; j0 < n_t; j0 += EP_THREADS) {
        __syncthreads();
        if (j0 + threadIdx.x < n_t) tile[threadIdx.x] = tp[j0 + threadIdx.x];
        __syncthreads();
        const int m = min(EP_THREADS, n_t - j0);
        if (live && !any)
            for (int j = 0; j < m; ++j) any |= ep_correct(qy, qx, tile[j], thr);
    }
    if (live) row_any[(size_t)p * capq + i] = any ? 1 : 0;
}

__global__ void points_correct_pairs_kernel(const double *__restrict__ qw, int capq, const int64_t *__restrict__ t, int capt,
                                            float thr, const int *__restrict__ mq, const int *__restrict__ mt,
                                            const int *__restrict__ nm, int capm, uint8_t *__restrict__ tp) {
    const int p = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = nm ? min(nm[p], capm) : capm;
    if (k >= n) return;
    const int i = mq[(size_t)p * capm + k], j = mt[(size_t)p * capm + k];
    uint8_t v = 0;
    if (i >= 0 && i < capq && j >= 0 && j < capt) {
        const double qy = qw[((size_t)p * capq + i) * 2], qx = qw[((size_t)p * capq + i) * 2 + 1];
        const longlong2 tt = reinterpret_cast<const longlong2 *>(t)[(size_t)p * capt + j];
        v = ep_correct(qy, qx, tt, thr) ? 1 : 0;
    }
    tp[(size_t)p * capm + k] = v;
}

}  // namespace mp

extern "C" int mp_warp_keypoints_i64(const int64_t *kp, const int *counts, int P, int cap, const double *Hm,
                                     double *out_f64, int64_t *out_i64, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(P >= 0 && cap >= 0, "mp_warp_keypoints_i64: bad shape P=%d cap=%d", P, cap);
    MP_CHECK_ARG(out_f64 || out_i64, "mp_warp_keypoints_i64: no output requested");
    if (P == 0 || cap == 0) return MP_OK;
    MP_CHECK_ARG(kp && Hm, "mp_warp_keypoints_i64: null pointer");
    MP_CHECK_ARG(P <= 65535, "mp_warp_keypoints_i64: at most 65535 problems per call");
    dim3 grid((unsigned)((cap + 255) / 256), (unsigned)P);
    mp::warp_keypoints_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kp, counts, cap, Hm, out_f64, out_i64);
    MP_LAUNCH_OK_S("warp_keypoints_kernel", (cudaStream_t)stream);
    return MP_OK;
}

extern "C" int mp_points_min_dist2_i64(const int64_t *q, const int *nq, int capq, const int64_t *t, const int *nt,
                                       int capt, int P, int H, int W, int64_t *min_d2, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(P >= 0 && capq >= 0 && capt >= 0, "mp_points_min_dist2_i64: bad shape");
    if (P == 0 || capq == 0) return MP_OK;
    MP_CHECK_ARG(q && min_d2 && (t || capt == 0), "mp_points_min_dist2_i64: null pointer");
    MP_CHECK_ARG(((uintptr_t)t & 15) == 0, "mp_points_min_dist2_i64: targets must be 16-byte aligned");
    MP_CHECK_ARG(P <= 65535, "mp_points_min_dist2_i64: at most 65535 problems per call");
    dim3 grid((unsigned)((capq + mp::EP_THREADS - 1) / mp::EP_THREADS), (unsigned)P);
    mp::points_min_dist2_kernel<<<grid, mp::EP_THREADS, 0, (cudaStream_t)stream>>>(q, nq, capq, t, nt, capt, H, W, min_d2);
    MP_LAUNCH_OK_S("points_min_dist2_kernel", (cudaStream_t)stream);
    return MP_OK;
}

extern "C" int mp_points_correct_f32(const double *qw, const int *nq, int capq, const int64_t *t, const int *nt, int capt,
                                     int P, float threshold, uint8_t *row_any, const int *mq, const int *mt, const int *nm,
                                     int capm, uint8_t *tp, mp_stream_t stream) {
    mp::prof_entry((cudaStream_t)stream);
    MP_CHECK_ARG(P >= 0 && capq >= 0 && capt >= 0 && capm >= 0, "mp_points_correct_f32: bad shape");
    if (P == 0) return MP_OK;
    MP_CHECK_ARG(P <= 65535, "mp_points_correct_f32: at most 65535 problems per call");
    MP_CHECK_ARG(((uintptr_t)t & 15) == 0, "mp_points_correct_f32: targets must be 16-byte aligned");
    if (row_any && capq > 0) {
        MP_CHECK_ARG(qw && (t || capt == 0), "mp_points_correct_f32: null pointer");
        dim3 grid((unsigned)((capq + mp::EP_THREADS - 1) / mp::EP_THREADS), (unsigned)P);
        mp::points_correct_kernel<<<grid, mp::EP_THREADS, 0, (cudaStream_t)stream>>>(qw, nq, capq, t, nt, capt, threshold, row_any);
        MP_LAUNCH_OK_S("points_correct_kernel", (cudaStream_t)stream);
    }
    if (tp && capm > 0) {
        MP_CHECK_ARG(qw && t && mq && mt, "mp_points_correct_f32: null pointer (match list)");
        dim3 grid((unsigned)((capm + 255) / 256), (unsigned)P);
        mp::points_correct_pairs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(qw, capq, t, capt, threshold, mq, mt, nm, capm, tp);
        MP_LAUNCH_OK_S("points_correct_pairs_kernel", (cudaStream_t)stream);
    }
    return MP_OK;
}
