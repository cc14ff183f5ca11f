/*
 * mp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference algorithms on MultiPoint's keypoint
 * extract-and-match hot path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and
 * only as the checker.  Nothing under multipoint_b200/ imports it.
 *
 * Every function cites the reference site it restates (paths relative to the
 * reference checkout, ethz-asl/multipoint).  Third-party arithmetic that is
 * not in the reference tree is restated from its published algorithm:
 *   - torchvision.ops.nms CPU kernel (requirements.txt:6, unpinned; 0.26.0
 *     installed): stable descending sort + greedy IoU suppression.
 *   - ATen grid_sampler_2d (bilinear / nearest, zeros / reflection padding,
 *     align_corners=True) and F.normalize.
 *   - OpenCV BFMatcher(NORM_L2, crossCheck) (requirements.txt:1 pins 4.2.0.34;
 *     4.13.0 installed): direct sqrt(sum((a-b)^2)) in fp32, first-minimum scan.
 *   - kornia homography_warp / dst_norm_to_dst_norm (NOT installed, no version
 *     pinned anywhere in the reference): PARITY UNPINNED, see DESIGN.md.
 *
 * Pinning: tests/test_oracle_golden.py checks every function here against the
 * fixtures in tests/golden/ that oracle/gen_golden.py produced by running the
 * reference's own Python functions (shimmed import) in the build container.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no fast-math so the
 * fp32 operation order written here is the order executed).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MPO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* Row 1: MultiPoint.detector_head, multipoint/models/MultiPoint.py:150-158  */
/* softmax over 65 channels (nn.Softmax2d :74), drop dustbin channel 64,     */
/* nn.PixelShuffle(8) (:75): prob[b,0,8h+i,8w+j] = softmax[b,8i+j,h,w].      */
/* logits: (B,65,Hc,Wc) fp32 NCHW.  prob: (B,1,8Hc,8Wc).                     */
/* ------------------------------------------------------------------------ */
MPO_API void mpo_detector_head(const float *logits, int B, int Hc, int Wc, float *prob)
{
    const int cells = Hc * Wc, W = 8 * Wc;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < Hc; ++h)
            for (int w = 0; w < Wc; ++w) {
                const float *x = logits + (size_t)b * 65 * cells + (size_t)h * Wc + w;
                float m = x[0];
                for (int c = 1; c < 65; ++c) {
                    float v = x[(size_t)c * cells];
                    if (v > m) m = v;
                }
                float e[65], sum = 0.f;
                for (int c = 0; c < 65; ++c) {
                    e[c] = expf(x[(size_t)c * cells] - m);
                    sum += e[c];
                }
                float *o = prob + (size_t)b * 64 * cells;
                for (int c = 0; c < 64; ++c) {
                    int i = c >> 3, j = c & 7;
                    o[(size_t)(8 * h + i) * W + 8 * w + j] = e[c] / sum;
                }
            }
}

/* ------------------------------------------------------------------------ */
/* SuperPointMagicLeap.generate_heatmap,                    */
/* multipoint/models/SuperPointMagicLeap.py:68-85.  dense = exp(semi) (:72),  */
/* dense / (sum over the 65 channels + 1e-5) (:73), dustbin dropped (:75),    */
/* cells unfolded 8x8 (:79-82) = the PixelShuffle(8) index map of row 1.      */
/* No max subtraction: large logits overflow exactly like the reference.      */
/* ------------------------------------------------------------------------ */
MPO_API void mpo_heatmap_magicleap(const float *semi, int B, int Hc, int Wc, float *prob)
{
    const int cells = Hc * Wc, W = 8 * Wc;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < Hc; ++h)
            for (int w = 0; w < Wc; ++w) {
                const float *x = semi + (size_t)b * 65 * cells + (size_t)h * Wc + w;
                float e[65], sum = 0.f;
                for (int c = 0; c < 65; ++c) {
                    e[c] = expf(x[(size_t)c * cells]);
                    sum += e[c];
                }
                const float den = sum + 1e-5f;
                float *o = prob + (size_t)b * 64 * cells;
                for (int c = 0; c < 64; ++c)
                    o[(size_t)(8 * h + (c >> 3)) * W + 8 * w + (c & 7)] = e[c] / den;
            }
}

/* ------------------------------------------------------------------------ */
/* Row 2: MultiPoint.descriptor_head tail, MultiPoint.py:160-166             */
/* F.normalize(x, p=2, dim=1): x / max(||x||_2, 1e-12) per cell.             */
/* x: (B,D,HW) fp32 (NCHW with the two spatial dims flattened).              */
/* ------------------------------------------------------------------------ */
MPO_API void mpo_normalize_descriptors(const float *x, int B, int D, int HW, float *out)
{
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < HW; ++p) {
            const float *xi = x + (size_t)b * D * HW + p;
            float ss = 0.f;
            for (int c = 0; c < D; ++c) {
                float v = xi[(size_t)c * HW];
                ss += v * v;
            }
            float n = sqrtf(ss);
            if (n < 1e-12f) n = 1e-12f;
            float *oi = out + (size_t)b * D * HW + p;
            for (int c = 0; c < D; ++c) oi[(size_t)c * HW] = xi[(size_t)c * HW] / n;
        }
}

/* ------------------------------------------------------------------------ */
/* Row 4: utils.box_nms, multipoint/utils/utils.py:78-122, on top of         */
/* torchvision.ops.nms / batched_nms (imported utils.py:4-5).                */
/*                                                                           */
/* mpo_box_nms_literal follows the reference step by step:                   */
/*  (i)   candidates = prob > min_prob (fp32 compare) in row-major order :97 */
/*  (ii)  boxes = point -/+ size*0.5 in fp32 :101,105                        */
/*  (iii) torchvision CPU nms: stable sort by descending score, greedy scan, */
/*        j suppressed iff inter/(area_i+area_j-inter) > iou with the fp32   */
/*        quotient compared against the DOUBLE threshold (the CPU kernel     */
/*        takes `double iou_threshold`); per image in 4-D mode :102-103      */
/*  (iv)  keep_top_k > 0: the first k survivors per image in descending      */
/*        score :109-116 (ties broken by lower row-major index = the stable  */
/*        order; the reference's 4-D path ends in a non-stable sort, so      */
/*        equal scores straddling k are undefined there -- see DESIGN.md)    */
/*  (v)   scatter survivor scores into zeros :119-120.                       */
/* O(N^2) like the reference: use on small inputs.                           */
/* ------------------------------------------------------------------------ */
typedef struct { float s; int idx; } mpo_cand;

static int mpo_cand_cmp(const void *a, const void *b)
{
    const mpo_cand *x = (const mpo_cand *)a, *y = (const mpo_cand *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* stable: lower index first */
}

MPO_API int mpo_box_nms_literal(const float *prob, int B, int H, int W, double size,
                                double min_prob, double iou, int keep_top_k, float *out)
{
    const float thr = (float)min_prob;
    const float half = (float)(size * 0.5);
    const size_t HW = (size_t)H * W;
    mpo_cand *c = (mpo_cand *)malloc(sizeof(mpo_cand) * HW);
    unsigned char *sup = (unsigned char *)malloc(HW);
    if (!c || !sup) { free(c); free(sup); return -1; }
    for (int b = 0; b < B; ++b) {
        const float *p = prob + b * HW;
        float *o = out + b * HW;
        memset(o, 0, sizeof(float) * HW);
        int n = 0;
        for (size_t i = 0; i < HW; ++i)
            if (p[i] > thr) { c[n].s = p[i]; c[n].idx = (int)i; ++n; }
        qsort(c, n, sizeof(mpo_cand), mpo_cand_cmp);
        memset(sup, 0, n);
        int kept = 0;
        for (int a = 0; a < n; ++a) {
            if (sup[a]) continue;
            if (keep_top_k <= 0 || kept < keep_top_k) o[c[a].idx] = c[a].s;
            ++kept;
            const float ay = (float)(c[a].idx / W), ax = (float)(c[a].idx % W);
            const float ay1 = ay - half, ax1 = ax - half, ay2 = ay + half, ax2 = ax + half;
            const float aarea = (ay2 - ay1) * (ax2 - ax1);
            for (int d = a + 1; d < n; ++d) {
                if (sup[d]) continue;
                const float by = (float)(c[d].idx / W), bx = (float)(c[d].idx % W);
                const float by1 = by - half, bx1 = bx - half, by2 = by + half, bx2 = bx + half;
                const float barea = (by2 - by1) * (bx2 - bx1);
                float yy1 = ay1 > by1 ? ay1 : by1, xx1 = ax1 > bx1 ? ax1 : bx1;
                float yy2 = ay2 < by2 ? ay2 : by2, xx2 = ax2 < bx2 ? ax2 : bx2;
                float w = yy2 - yy1; if (w < 0.f) w = 0.f;
                float h = xx2 - xx1; if (h < 0.f) h = 0.f;
                float inter = w * h;
                float ovr = inter / (aarea + barea - inter);
                if ((double)ovr > iou) sup[d] = 1;
            }
        }
    }
    free(c); free(sup);
    return 0;
}

/* Footprint of the IoU test for two size x size boxes whose centres differ   */
/* by (dy,dx): the same fp32 expression as above evaluated at the origin.     */
/* Boxes centred on integer pixels with a half-size that is a small dyadic    */
/* rational are translation invariant in fp32, so one table serves the image. */
/* fp[(dy+R)*(2R+1)+(dx+R)] = 1 iff the lower-scored point is suppressed.     */
MPO_API int mpo_nms_footprint(double size, double iou, int R, unsigned char *fp)
{
    const float half = (float)(size * 0.5);
    const float a1 = 0.f - half, a2 = 0.f + half;
    const float area = (a2 - a1) * (a2 - a1);
    int n = 0;
    for (int dy = -R; dy <= R; ++dy)
        for (int dx = -R; dx <= R; ++dx) {
            const float by1 = (float)dy - half, by2 = (float)dy + half;
            const float bx1 = (float)dx - half, bx2 = (float)dx + half;
            float yy1 = a1 > by1 ? a1 : by1, xx1 = a1 > bx1 ? a1 : bx1;
            float yy2 = a2 < by2 ? a2 : by2, xx2 = a2 < bx2 ? a2 : bx2;
            float w = yy2 - yy1; if (w < 0.f) w = 0.f;
            float h = xx2 - xx1; if (h < 0.f) h = 0.f;
            float inter = w * h;
            float ovr = inter / (area + area - inter);
            unsigned char hit = ((double)ovr > iou) && !(dy == 0 && dx == 0);
            fp[(dy + R) * (2 * R + 1) + (dx + R)] = hit;
            n += hit;
        }
    return n;
}

/* Same result as mpo_box_nms_literal in O(N * footprint): walk candidates in */
/* the same stable order; a kept point marks its footprint as suppressed.     */
/* Validated against the literal version and the reference goldens.           */
MPO_API int mpo_box_nms(const float *prob, int B, int H, int W, double size,
                        double min_prob, double iou, int keep_top_k, float *out)
{
    const float thr = (float)min_prob;
    int R = (int)ceil(size);
    if (R < 1) R = 1;
    const int S = 2 * R + 1;
    unsigned char *fp = (unsigned char *)malloc((size_t)S * S);
    const size_t HW = (size_t)H * W;
    mpo_cand *c = (mpo_cand *)malloc(sizeof(mpo_cand) * HW);
    unsigned char *sup = (unsigned char *)malloc(HW);
    if (!fp || !c || !sup) { free(fp); free(c); free(sup); return -1; }
    mpo_nms_footprint(size, iou, R, fp);
    for (int b = 0; b < B; ++b) {
        const float *p = prob + b * HW;
        float *o = out + b * HW;
        memset(o, 0, sizeof(float) * HW);
        memset(sup, 0, HW);
        int n = 0;
        for (size_t i = 0; i < HW; ++i)
            if (p[i] > thr) { c[n].s = p[i]; c[n].idx = (int)i; ++n; }
        qsort(c, n, sizeof(mpo_cand), mpo_cand_cmp);
        int kept = 0;
        for (int a = 0; a < n; ++a) {
            const int idx = c[a].idx;
            if (sup[idx]) continue;
            if (keep_top_k <= 0 || kept < keep_top_k) o[idx] = c[a].s;
            ++kept;
            const int y = idx / W, x = idx % W;
            for (int dy = -R; dy <= R; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int dx = -R; dx <= R; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= W) continue;
                    if (fp[(dy + R) * S + dx + R]) sup[(size_t)yy * W + xx] = 1;
                }
            }
        }
    }
    free(fp); free(c); free(sup);
    return 0;
}

/* Row 4b: torch.nonzero((p > thr).float()) keypoint idiom                    */
/* (predict_align_image_pair.py:170-171, evaluation.py:157-158,262-263,       */
/* export_keypoints.py:100): (y,x) int64 in row-major order. Returns count;   */
/* writes at most cap entries.                                                */
MPO_API int64_t mpo_extract_keypoints(const float *prob, int H, int W, double thr_d,
                                      int64_t *kp, int64_t cap)
{
    const float thr = (float)thr_d;
    int64_t n = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (prob[(size_t)y * W + x] > thr) {
                if (n < cap) { kp[2 * n] = y; kp[2 * n + 1] = x; }
                ++n;
            }
    return n;
}

/* ------------------------------------------------------------------------ */
/* Row 5: utils.interpolate_descriptors, utils.py:159-167.                   */
/* y_n = y/(H*0.5) - 1, x_n = x/(W*0.5) - 1 (:162-163), grid_sample bilinear,*/
/* zero padding, align_corners=True (:166), then F.normalize (:167).         */
/* kp: (K,2) int64 (y,x); desc: (D,Hc,Wc) fp32 CHW; out: (K,D).              */
/* ------------------------------------------------------------------------ */
static inline float mpo_unnormalize_ac(float coord, int size)
{
    return ((coord + 1.f) / 2.f) * (float)(size - 1);
}

MPO_API void mpo_interpolate_descriptors(const int64_t *kp, int64_t K, const float *desc, int D,
                                         int Hc, int Wc, int H, int W, float *out)
{
    const float hh = (float)H * 0.5f, hw = (float)W * 0.5f;
    for (int64_t k = 0; k < K; ++k) {
        const float yn = (float)kp[2 * k] / hh - 1.0f;
        const float xn = (float)kp[2 * k + 1] / hw - 1.0f;
        const float ix = mpo_unnormalize_ac(xn, Wc), iy = mpo_unnormalize_ac(yn, Hc);
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
        const float nw = ((float)x1 - ix) * ((float)y1 - iy);
        const float ne = (ix - (float)x0) * ((float)y1 - iy);
        const float sw = ((float)x1 - ix) * (iy - (float)y0);
        const float se = (ix - (float)x0) * (iy - (float)y0);
        const int vx0 = x0 >= 0 && x0 < Wc, vx1 = x1 >= 0 && x1 < Wc;
        const int vy0 = y0 >= 0 && y0 < Hc, vy1 = y1 >= 0 && y1 < Hc;
        float *o = out + (size_t)k * D;
        float ss = 0.f;
        for (int c = 0; c < D; ++c) {
            const float *m = desc + (size_t)c * Hc * Wc;
            float v = 0.f;
            if (vy0 && vx0) v += m[y0 * Wc + x0] * nw;
            if (vy0 && vx1) v += m[y0 * Wc + x1] * ne;
            if (vy1 && vx0) v += m[y1 * Wc + x0] * sw;
            if (vy1 && vx1) v += m[y1 * Wc + x1] * se;
            o[c] = v;
            ss += v * v;
        }
        float n = sqrtf(ss);
        if (n < 1e-12f) n = 1e-12f;
        for (int c = 0; c < D; ++c) o[c] = o[c] / n;
    }
}

/* ------------------------------------------------------------------------ */
/* Rows 6-8: matching, multipoint/utils/matching.py.                         */
/*                                                                           */
/* mode 0 "nn"    NNMatcher.match :41-72   key = sqrt(2-2*clip(a.b,-1,1))     */
/* mode 1 "bf"    cv2.BFMatcher(NORM_L2)   key = sqrt(sum((a-b)^2))  :7,31    */
/* prec 0: fp32 arithmetic in the order written here                          */
/* prec 1: fp64 arithmetic on the fp32 inputs = the near-tie-free "truth"     */
/*         the CUDA matcher is required to reproduce index for index.         */
/*                                                                           */
/* For every row of desc1: best (first minimum, like np.argmin / OpenCV's     */
/* strict '<' scan) and second-best key over desc2.  best2/second2 likewise   */
/* for every row of desc2 over desc1 (np.argmin(axis=0) :58, crossCheck).     */
/* ------------------------------------------------------------------------ */
static double mpo_key(const float *a, const float *b, int D, int mode, int prec)
{
    if (prec == 0) {
        if (mode == 0) {
            float dot = 0.f;
            for (int k = 0; k < D; ++k) dot += a[k] * b[k];
            if (dot < -1.f) dot = -1.f;
            if (dot > 1.f) dot = 1.f;
            return (double)sqrtf(2.f - 2.f * dot);
        } else {
            float s = 0.f;
            for (int k = 0; k < D; ++k) { float d = a[k] - b[k]; s += d * d; }
            return (double)sqrtf(s);
        }
    } else {
        if (mode == 0) {
            double dot = 0.0;
            for (int k = 0; k < D; ++k) dot += (double)a[k] * (double)b[k];
            if (dot < -1.0) dot = -1.0;
            if (dot > 1.0) dot = 1.0;
            return sqrt(2.0 - 2.0 * dot);
        } else {
            double s = 0.0;
            for (int k = 0; k < D; ++k) { double d = (double)a[k] - (double)b[k]; s += d * d; }
            return sqrt(s);
        }
    }
}

MPO_API void mpo_nearest(const float *d1, int N1, const float *d2, int N2, int D, int mode,
                         int prec, int32_t *idx12, double *best12, double *second12,
                         int32_t *idx21, double *best21, double *second21)
{
    for (int j = 0; j < N2; ++j) { idx21[j] = -1; best21[j] = INFINITY; second21[j] = INFINITY; }
    for (int i = 0; i < N1; ++i) {
        int bi = -1; double b = INFINITY, s = INFINITY;
        for (int j = 0; j < N2; ++j) {
            const double k = mpo_key(d1 + (size_t)i * D, d2 + (size_t)j * D, D, mode, prec);
            if (k < b) { s = b; b = k; bi = j; } else if (k < s) s = k;
            if (k < best21[j]) { second21[j] = best21[j]; best21[j] = k; idx21[j] = i; }
            else if (k < second21[j]) second21[j] = k;
        }
        idx12[i] = bi; best12[i] = b; second12[i] = s;
    }
}

/* Mutual-nearest-neighbour match list in ascending query order.              */
/* nn  (matching.py:53-70): keep i iff best < threshold and idx21[idx12[i]]==i*/
/* bf  crossCheck=True (matching.py:7,31): keep i iff idx21[idx12[i]] == i    */
/* bf  crossCheck=False: every query keeps its nearest train descriptor.      */
/* threshold < 0 disables the distance test.  Returns the match count.        */
MPO_API int mpo_match_mutual(const float *d1, int N1, const float *d2, int N2, int D, int mode,
                             int prec, int cross_check, double threshold, int32_t *q, int32_t *t,
                             float *dist)
{
    if (N1 == 0 || N2 == 0) return 0;
    int32_t *i12 = (int32_t *)malloc(sizeof(int32_t) * N1), *i21 = (int32_t *)malloc(sizeof(int32_t) * N2);
    double *b12 = (double *)malloc(sizeof(double) * N1 * 2), *b21 = (double *)malloc(sizeof(double) * N2 * 2);
    mpo_nearest(d1, N1, d2, N2, D, mode, prec, i12, b12, b12 + N1, i21, b21, b21 + N2);
    int n = 0;
    for (int i = 0; i < N1; ++i) {
        const int j = i12[i];
        if (threshold >= 0.0 && !((float)b12[i] < (float)threshold)) continue;
        if (cross_check && i21[j] != i) continue;
        q[n] = i; t[n] = j;
        /* reported distance is always the fp32 value the reference would print */
        dist[n] = (float)mpo_key(d1 + (size_t)i * D, d2 + (size_t)j * D, D, mode, 0);
        ++n;
    }
    free(i12); free(i21); free(b12); free(b21);
    return n;
}

/* knn_matches=True path of get_matches (matching.py:21-28): knnMatch(k=2)    */
/* then Lowe ratio m.distance < 0.9 * n.distance on L2 (not squared).         */
MPO_API int mpo_match_ratio(const float *d1, int N1, const float *d2, int N2, int D, int mode,
                            int prec, double ratio, int32_t *q, int32_t *t, float *dist)
{
    if (N1 == 0 || N2 < 2) return 0;
    int n = 0;
    for (int i = 0; i < N1; ++i) {
        int bi = -1; double b = INFINITY, s = INFINITY;
        for (int j = 0; j < N2; ++j) {
            const double k = mpo_key(d1 + (size_t)i * D, d2 + (size_t)j * D, D, mode, prec);
            if (k < b) { s = b; b = k; bi = j; } else if (k < s) s = k;
        }
        if ((float)b < (float)ratio * (float)s) {
            q[n] = i; t[n] = bi;
            dist[n] = (float)mpo_key(d1 + (size_t)i * D, d2 + (size_t)bi * D, D, mode, 0);
            ++n;
        }
    }
    return n;
}

/* ThresholdMatcher.match (matching.py:74-99): every pair with key < threshold */
/* in np.argwhere (row-major) order.  Returns the total count; writes at most  */
/* cap entries.                                                                */
MPO_API int64_t mpo_match_threshold(const float *d1, int N1, const float *d2, int N2, int D,
                                    int prec, double threshold, int32_t *q, int32_t *t,
                                    float *dist, int64_t cap)
{
    int64_t n = 0;
    for (int i = 0; i < N1; ++i)
        for (int j = 0; j < N2; ++j) {
            const double k = mpo_key(d1 + (size_t)i * D, d2 + (size_t)j * D, D, 0, prec);
            if ((float)k < (float)threshold) {
                if (n < cap) {
                    q[n] = i; t[n] = j;
                    dist[n] = (float)mpo_key(d1 + (size_t)i * D, d2 + (size_t)j * D, D, 0, 0);
                }
                ++n;
            }
        }
    return n;
}

/* ------------------------------------------------------------------------ */
/* Rows 9-10: warp_perspective_tensor (homographies.py:404-425) as kornia's  */
/* homography_warp does it: base grid linspace(-1,1) over the destination,   */
/* transformed by the normalised matrix A (dst_norm -> src_norm), then       */
/* F.grid_sample(..., align_corners=True).  PARITY UNPINNED (kornia absent). */
/*                                                                           */
/* xs/ys are the linspace tables (length W / H) so the caller controls how   */
/* they were rounded.  A is row-major 3x3 fp32 per homography.  The divide   */
/* is a multiply by the reciprocal guarded by |z| > 1e-8 as in kornia's      */
/* convert_points_from_homogeneous.                                          */
/* mode: 0 bilinear, 1 nearest.  padding: 0 zeros, 1 reflection.             */
/* ------------------------------------------------------------------------ */
static inline void mpo_src_coord(const float *A, float xs, float ys, int Ws, int Hs, float *ix, float *iy)
{
    const float X = A[0] * xs + A[1] * ys + A[2];
    const float Y = A[3] * xs + A[4] * ys + A[5];
    const float Z = A[6] * xs + A[7] * ys + A[8];
    const float sc = fabsf(Z) > 1e-8f ? 1.0f / Z : 1.0f;
    *ix = mpo_unnormalize_ac(X * sc, Ws);
    *iy = mpo_unnormalize_ac(Y * sc, Hs);
}

static inline float mpo_reflect(float in, int size)
{
    /* ATen reflect_coordinates(in, 0, 2*(size-1)) then clip_coordinates */
    if (size <= 1) return 0.f;
    const float span = (float)(size - 1);
    in = fabsf(in);
    const float extra = fmodf(in, span);
    const int flips = (int)floorf(in / span);
    float r = (flips % 2 == 0) ? extra : span - extra;
    if (r < 0.f) r = 0.f;
    if (r > (float)(size - 1)) r = (float)(size - 1);
    return r;
}

static inline float mpo_sample(const float *src, int Hs, int Ws, float ix, float iy, int mode, int padding)
{
    if (padding == 1) { ix = mpo_reflect(ix, Ws); iy = mpo_reflect(iy, Hs); }
    if (mode == 1) {
        const float rx = nearbyintf(ix), ry = nearbyintf(iy);
        if (!(rx >= 0.f && rx < (float)Ws && ry >= 0.f && ry < (float)Hs)) return 0.f;
        return src[(int)ry * Ws + (int)rx];
    }
    const float fx = floorf(ix), fy = floorf(iy);
    if (!(fx >= -1.f && fx <= (float)Ws && fy >= -1.f && fy <= (float)Hs)) return 0.f; /* also NaN */
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float nw = ((float)x1 - ix) * ((float)y1 - iy);
    const float ne = (ix - (float)x0) * ((float)y1 - iy);
    const float sw = ((float)x1 - ix) * (iy - (float)y0);
    const float se = (ix - (float)x0) * (iy - (float)y0);
    const int vx0 = x0 >= 0 && x0 < Ws, vx1 = x1 >= 0 && x1 < Ws;
    const int vy0 = y0 >= 0 && y0 < Hs, vy1 = y1 >= 0 && y1 < Hs;
    float v = 0.f;
    if (vy0 && vx0) v += src[y0 * Ws + x0] * nw;
    if (vy0 && vx1) v += src[y0 * Ws + x1] * ne;
    if (vy1 && vx0) v += src[y1 * Ws + x0] * sw;
    if (vy1 && vx1) v += src[y1 * Ws + x1] * se;
    return v;
}

/* src: (N,H,W); A: (3,3) shared by all N planes; out: (N,H,W) */
MPO_API void mpo_warp(const float *src, int N, int H, int W, const float *A, const float *xs,
                      const float *ys, int mode, int padding, float *out)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float ix, iy;
            mpo_src_coord(A, xs[x], ys[y], W, H, &ix, &iy);
            for (int n = 0; n < N; ++n)
                out[((size_t)n * H + y) * W + x] =
                    mpo_sample(src + (size_t)n * H * W, H, W, ix, iy, mode, padding);
        }
}

/* ------------------------------------------------------------------------ */
/* Row 9: the aggregate half of homographic_adaptation (homographies.py      */
/* :153,162-187) and homographic_adaptation_multispectral (:62-66,77-126).   */
/*  prob0   (B,H,W)      identity-pass heatmap (already o*t or o+t for pairs) */
/*  probw_a (n,B,H,W)    heatmaps of the warped images, n = num-1             */
/*  probw_b same or NULL second spectrum; combined in the WARPED frame        */
/*                       (:105-108) before the unwarp                         */
/*  masks   (n,H,W)      valid masks as fp32 {0,1}                            */
/*  Ainv    (n,3,3)      normalised matrices of the UNWARP (warper(.., H^-1)) */
/*  aggregation 0 none (single spectrum), 1 'prod' (sqrt), 2 'sum' (*0.5)     */
/* count starts at 1 (:62,153); per sample count_sample = nearest/zeros warp  */
/* of the mask (:112,180); prob += bilinear/zeros warp * count_sample         */
/* (:113-114,181-182); out = prob/count (:116,184); zero where count <        */
/* min_count (:125-126,186-187).                                              */
/* ------------------------------------------------------------------------ */
MPO_API void mpo_ha_aggregate(const float *prob0, const float *probw_a, const float *probw_b,
                              const float *masks, const float *Ainv, int n, int B, int H, int W,
                              const float *xs, const float *ys, int aggregation, int min_count,
                              float *out, float *count_out)
{
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float prob = prob0[b * HW + (size_t)y * W + x];
                float count = 1.0f;
                for (int i = 0; i < n; ++i) {
                    float ix, iy;
                    mpo_src_coord(Ainv + 9 * i, xs[x], ys[y], W, H, &ix, &iy);
                    const float cs = mpo_sample(masks + i * HW, H, W, ix, iy, 1, 0);
                    const float *pa = probw_a + ((size_t)i * B + b) * HW;
                    float v;
                    if (probw_b) {
                        /* combine spectra per source pixel, then bilinear */
                        const float *pb = probw_b + ((size_t)i * B + b) * HW;
                        const float fx = floorf(ix), fy = floorf(iy);
                        v = 0.f;
                        if (fx >= -1.f && fx <= (float)W && fy >= -1.f && fy <= (float)H) {
                            const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
                            const float nw = ((float)x1 - ix) * ((float)y1 - iy);
                            const float ne = (ix - (float)x0) * ((float)y1 - iy);
                            const float sw = ((float)x1 - ix) * (iy - (float)y0);
                            const float se = (ix - (float)x0) * (iy - (float)y0);
                            const int vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
                            const int vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
#define MPO_COMB(yy, xx) (aggregation == 1 ? pa[(yy) * W + (xx)] * pb[(yy) * W + (xx)] \
                                           : pa[(yy) * W + (xx)] + pb[(yy) * W + (xx)])
                            if (vy0 && vx0) v += MPO_COMB(y0, x0) * nw;
                            if (vy0 && vx1) v += MPO_COMB(y0, x1) * ne;
                            if (vy1 && vx0) v += MPO_COMB(y1, x0) * sw;
                            if (vy1 && vx1) v += MPO_COMB(y1, x1) * se;
#undef MPO_COMB
                        }
                    } else {
                        v = mpo_sample(pa, H, W, ix, iy, 0, 0);
                    }
                    count += cs;
                    prob += v * cs;
                }
                float o = prob / count;
                if (aggregation == 1) o = sqrtf(o);
                else if (aggregation == 2) o *= 0.5f;
                if (min_count > 0 && count < (float)min_count) o = 0.f;
                out[b * HW + (size_t)y * W + x] = o;
                if (count_out) count_out[b * HW + (size_t)y * W + x] = count;
            }
}

/* ------------------------------------------------------------------------ */
/* compute_valid_mask,                          */
/* multipoint/utils/homographies.py:375-402.                                  */
/*   mask = cv2.warpPerspective(ones(H,W) float64, Hm, (W,H), INTER_NEAREST)  */
/*   optional 1-px zero frame, cv2.erode by a (2r+1)^2 box, frame cropped.    */
/* The arithmetic is OpenCV's (third party, opencv-python pinned 4.2.0.34 in  */
/* requirements.txt:1, 4.13.0 installed), restated from its published         */
/* algorithm and pinned against cv2 itself by oracle/gen_golden.py:           */
/*  - warpPerspective inverts the matrix (cv::invert, closed form for 3x3),   */
/*  - walks the destination in blocks of bw0 columns; for a pixel (x,y) in    */
/*    the block starting at xb: X0 = M0*xb + M1*y + M2 (same for Y0, W0),     */
/*    W = W0 + M6*x1, W = W ? 1/W : 0, fX = (X0 + M0*x1)*W clamped to int     */
/*    range, X = round-half-even(fX) -- all in double,                        */
/*  - nearest remap with constant border 0: ones inside the source, 0 outside */
/*  - erode: minimum over the window, pixels outside the image ignored        */
/*    (border value +max), so only the explicit zero frame erodes the edge.   */
/* Minv: the inverted matrix as cv2.invert returns it (row-major, 9 doubles). */
/* ------------------------------------------------------------------------ */
MPO_API void mpo_invert3x3(const double *S, double *T)
{
    /* cv::invert, 3x3 closed form: cofactors times 1/det, 0 matrix if det == 0 */
    const double d0 = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
                      S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d0 == 0.) {
        for (int i = 0; i < 9; ++i) T[i] = 0.;
        return;
    }
    const double d = 1. / d0;
    T[0] = (S[4] * S[8] - S[5] * S[7]) * d;
    T[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    T[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    T[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    T[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    T[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    T[6] = (S[3] * S[7] - S[4] * S[6]) * d;
    T[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    T[8] = (S[0] * S[4] - S[1] * S[3]) * d;
}

static int mpo_round_even_sat(double v)
{
    if (v < -2147483648.0) v = -2147483648.0;
    if (v > 2147483647.0) v = 2147483647.0;
    return (int)nearbyint(v); /* default rounding mode: half to even, like cvRound */
}

MPO_API void mpo_valid_mask(const double *Minv, int H, int W, int erosion_radius, int mask_border, uint8_t *mask)
{
    int bh0 = H < 16 ? H : 16;
    int bw0 = 1024 / bh0;
    if (bw0 > W) bw0 = W;
    uint8_t *raw = (uint8_t *)malloc((size_t)H * W);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int xb = (x / bw0) * bw0, x1 = x - xb;
            const double X0 = Minv[0] * xb + Minv[1] * y + Minv[2];
            const double Y0 = Minv[3] * xb + Minv[4] * y + Minv[5];
            const double W0 = Minv[6] * xb + Minv[7] * y + Minv[8];
            double w = W0 + Minv[6] * x1;
            w = w ? 1. / w : 0;
            const int sx = mpo_round_even_sat((X0 + Minv[0] * x1) * w);
            const int sy = mpo_round_even_sat((Y0 + Minv[3] * x1) * w);
            raw[(size_t)y * W + x] = (sx >= 0 && sx < W && sy >= 0 && sy < H) ? 1 : 0;
        }
    const int r = erosion_radius;
    if (r <= 0) {
        memcpy(mask, raw, (size_t)H * W);
    } else {
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                uint8_t v = 1;
                for (int dy = -r; dy <= r && v; ++dy)
                    for (int dx = -r; dx <= r; ++dx) {
                        const int yy = y + dy, xx = x + dx;
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) {
                            if (mask_border) { v = 0; break; }
                            continue;
                        }
                        if (!raw[(size_t)yy * W + xx]) { v = 0; break; }
                    }
                mask[(size_t)y * W + x] = v;
            }
    }
    free(raw);
}

/* ------------------------------------------------------------------------ */
/* the point geometry inside compute_repeatability_multi-   */
/* spectral (multipoint/utils/evaluation.py:148-199) and                      */
/* compute_descriptor_metrics (:253-358).                                     */
/* ------------------------------------------------------------------------ */

/* warp_keypoints, multipoint/utils/homographies.py:331-346:                  */
/*   cv2.perspectiveTransform([kp[:, ::-1]] as float64, Hm)[0, :, ::-1]       */
/* OpenCV's arithmetic (third party, see the valid-mask note above), pinned   */
/* against cv2 by oracle/gen_golden.py: w = x*m6 + y*m7 + m8; if |w| >        */
/* FLT_EPSILON: w = 1/w, x' = (x*m0 + y*m1 + m2)*w, y' likewise; else 0.      */
/* The installed build (4.13.0, AVX2 dispatch of matmul.simd.hpp) contracts   */
/* x*a + y*b into fma(x, a, y*b) -- measured: that form is bit-identical on   */
/* 15 000 random points, the uncontracted one differs in the last ulp on      */
/* ~27 % of them.  OpenCV's result is therefore platform dependent at 1 ulp;  */
/* this restates the FMA form.                                                */
/* pts (N,2) int64 (y,x) -> out_f64 (N,2) double (y,x); when out_i64 != NULL  */
/* also the reference's `.astype(int)` (C truncation toward zero).            */
MPO_API void mpo_warp_keypoints(const int64_t *pts, int N, const double *m, double *out_f64, int64_t *out_i64)
{
    for (int i = 0; i < N; ++i) {
        const double y = (double)pts[2 * i], x = (double)pts[2 * i + 1];
        double w = fma(x, m[6], y * m[7]) + m[8];
        double xo = 0., yo = 0.;
        if (fabs(w) > 1.1920928955078125e-07) {
            w = 1. / w;
            xo = (fma(x, m[0], y * m[1]) + m[2]) * w;
            yo = (fma(x, m[3], y * m[4]) + m[5]) * w;
        }
        if (out_f64) { out_f64[2 * i] = yo; out_f64[2 * i + 1] = xo; }
        if (out_i64) { out_i64[2 * i] = (int64_t)yo; out_i64[2 * i + 1] = (int64_t)xo; }
    }
}

/* evaluation.py:176-197 for one direction: queries = warped keypoints (already truncated to int),  */
/* filter_points (homographies.py:358-372) keeps 0 <= y < H, 0 <= x < W; for each kept query the     */
/* minimum over the targets of ||q - t||_2 (np.linalg.norm of an int64 difference = sqrt of an exact  */
/* integer); min_d2[i] = that minimum squared (exact), -1 for filtered-out queries, INT64_MAX when    */
/* there is no target.  The caller counts sqrt(min_d2) <= distance_thresh in double.                  */
MPO_API void mpo_points_min_dist2(const int64_t *q, int Nq, const int64_t *t, int Nt, int H, int W, int64_t *min_d2)
{
    for (int i = 0; i < Nq; ++i) {
        const int64_t qy = q[2 * i], qx = q[2 * i + 1];
        if (qy < 0 || qx < 0 || qy >= H || qx >= W) { min_d2[i] = -1; continue; }
        int64_t best = INT64_MAX;
        for (int j = 0; j < Nt; ++j) {
            const int64_t dy = qy - t[2 * j], dx = qx - t[2 * j + 1];
            const int64_t d2 = dy * dy + dx * dx;
            if (d2 < best) best = d2;
        }
        min_d2[i] = best;
    }
}

/* evaluation.py:294-298,301-315: correct[i,j] = ||float32(warped[i] - kp[j])||_2 <= threshold_keypoints  */
/* with warped in float64 (warp_keypoints(..., np.float)) and kp int64; the difference is taken in double, */
/* rounded to fp32, the norm accumulated in fp32 (dy*dy + dx*dx, no FMA) -- torch.norm's exact reduction  */
/* order for two elements is third-party and unpinned at 1-ulp boundaries.  row_any[i] = any_j correct     */
/* (the rows counted by correct.sum(1).nonzero()); tp[k] = correct[mq[k], mt[k]] for the M given matches.   */
MPO_API void mpo_points_correct(const double *qw, int Nq, const int64_t *t, int Nt, float thr, uint8_t *row_any,
                                const int32_t *mq, const int32_t *mt, int M, uint8_t *tp)
{
    for (int i = 0; i < Nq; ++i) {
        uint8_t any = 0;
        for (int j = 0; j < Nt && !any; ++j) {
            const float dy = (float)(qw[2 * i] - (double)t[2 * j]), dx = (float)(qw[2 * i + 1] - (double)t[2 * j + 1]);
            const float d = sqrtf(dy * dy + dx * dx);
            any = d <= thr;
        }
        row_any[i] = any;
    }
    for (int k = 0; k < M; ++k) {
        const int i = mq[k], j = mt[k];
        const float dy = (float)(qw[2 * i] - (double)t[2 * j]), dx = (float)(qw[2 * i + 1] - (double)t[2 * j + 1]);
        tp[k] = sqrtf(dy * dy + dx * dx) <= thr;
    }
}

/* ------------------------------------------------------------------------ */
/* Row 3 glue: what sits between the backbone's convolutions,                 */
/* multipoint/models/MultiPoint.py:61-90 (getNonlinearity, getConvolutionBlock,*/
/* generate_encoder): nn.ReLU -> nn.BatchNorm2d (eval: (x-mean)/sqrt(var+eps)  */
/* *weight+bias) [-> nn.MaxPool2d(2,2)] [-> nn.ReflectionPad2d(1) |            */
/* nn.ZeroPad2d(1)], or BatchNorm first with bn_first.  conv_bias (may be NULL)*/
/* is the preceding convolution's bias.  x (B,C,H,W) -> out (B,C,Ho+2p,Wo+2p). */
/* ------------------------------------------------------------------------ */
static float mpo_act(float x, float cb, float mean, float var, float eps, float w, float b, int bn_first)
{
    x = x + cb;
    if (bn_first) {
        float y = (x - mean) / sqrtf(var + eps) * w + b;
        return y > 0.f ? y : 0.f;
    }
    x = x > 0.f ? x : 0.f;
    return (x - mean) / sqrtf(var + eps) * w + b;
}

static int mpo_glue_reflect(int i, int n, int reflect, int *outside)
{
    *outside = 0;
    if (i < 0) { *outside = 1; return reflect ? -i : 0; }
    if (i >= n) { *outside = 1; return reflect ? 2 * n - 2 - i : 0; }
    return i;
}

MPO_API void mpo_relu_bn_pad(const float *x, int B, int C, int H, int W, const float *conv_bias, const float *mean,
                             const float *var, float eps, const float *weight, const float *bias, int bn_first, int pool,
                             int pad, int reflect, float *out)
{
    const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W, Hp = Ho + 2 * pad, Wp = Wo + 2 * pad;
    for (int p = 0; p < B * C; ++p) {
        const int c = p % C;
        const float cb = conv_bias ? conv_bias[c] : 0.f;
        for (int yo = 0; yo < Hp; ++yo)
            for (int xo = 0; xo < Wp; ++xo) {
                int oy, ox;
                const int ys = mpo_glue_reflect(yo - pad, Ho, reflect, &oy), xs = mpo_glue_reflect(xo - pad, Wo, reflect, &ox);
                float v;
                if ((oy || ox) && !reflect) {
                    v = 0.f;
                } else if (pool) {
                    v = -INFINITY;
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            const float t = mpo_act(x[((size_t)p * H + 2 * ys + dy) * W + 2 * xs + dx], cb, mean[c], var[c], eps,
                                                    weight[c], bias[c], bn_first);
                            if (t > v) v = t;
                        }
                } else {
                    v = mpo_act(x[((size_t)p * H + ys) * W + xs], cb, mean[c], var[c], eps, weight[c], bias[c], bn_first);
                }
                out[((size_t)p * Hp + yo) * Wp + xo] = v;
            }
    }
}

/* First encoder layer: [pad 1] -> Conv2d(1 -> C, 3x3, cross-correlation like torch) -> the chain above [-> pad 1]. */
MPO_API void mpo_conv1_relu_bn_pad(const float *img, int B, int H, int W, const float *w9, const float *conv_bias, int C,
                                   const float *mean, const float *var, float eps, const float *weight, const float *bias,
                                   int bn_first, int in_reflect, int pad, int out_reflect, float *out)
{
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int yo = 0; yo < Hp; ++yo)
                for (int xo = 0; xo < Wp; ++xo) {
                    int oy, ox;
                    const int ys = mpo_glue_reflect(yo - pad, H, out_reflect, &oy), xs = mpo_glue_reflect(xo - pad, W, out_reflect, &ox);
                    float v = 0.f;
                    if (!((oy || ox) && !out_reflect)) {
                        float acc = 0.f;
                        for (int k = 0; k < 9; ++k) {
                            int iy_out, ix_out;
                            const int iy = mpo_glue_reflect(ys + k / 3 - 1, H, in_reflect, &iy_out);
                            const int ix = mpo_glue_reflect(xs + k % 3 - 1, W, in_reflect, &ix_out);
                            const float t = ((iy_out || ix_out) && !in_reflect) ? 0.f : img[((size_t)b * H + iy) * W + ix];
                            acc += t * w9[c * 9 + k];
                        }
                        v = mpo_act(acc, conv_bias ? conv_bias[c] : 0.f, mean[c], var[c], eps, weight[c], bias[c], bn_first);
                    }
                    out[(((size_t)b * C + c) * Hp + yo) * Wp + xo] = v;
                }
}
