/*
 * multipoint_b200.h -- C ABI of the B200-native keypoint extract-and-match hot path.
 *
 * The reference (ethz-asl/multipoint) is pure Python: it has no plugin / FFI boundary, so the
 * drop-in boundary is a set of Python callables.  This header is what a
 * binding for that path would call: plain pointers and sizes, no torch types.  Each entry point
 * names the reference function it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the library never allocates or frees caller-visible memory: scratch comes from the caller
 *    through (workspace, workspace_bytes); mp_*_workspace_bytes gives the required size;
 *  - stream is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *    unless documented;
 *  - return value 0 = ok, negative = error (mp_status); mp_last_error_string() describes the
 *    last error of the calling thread.  No exceptions cross the ABI, no global mutable state;
 *  - tensors are dense, row-major, fp32 unless stated; images are NCHW like the reference.
 */
#ifndef MULTIPOINT_B200_H
#define MULTIPOINT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP_API __attribute__((visibility("default")))

typedef void *mp_stream_t;

typedef enum {
    MP_OK = 0,
    MP_ERR_INVALID = -1,     /* bad argument (shape, alignment, enum) */
    MP_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
    MP_ERR_WORKSPACE = -3,   /* workspace missing or too small */
    MP_ERR_UNSUPPORTED = -4  /* valid in the reference, not supported here (documented) */
} mp_status;

MP_API int mp_version(void);
MP_API const char *mp_last_error_string(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
MP_API unsigned long long mp_launch_count(void);
/* Per-kernel timing (bench.py's roofline): between mp_profile_begin and mp_profile_end every kernel
 * launch of this library is followed by a CUDA event on its launching stream and every API call
 * starts with one; a kernel's duration is the gap to the previous event on that stream.
 * mp_profile_end synchronises those events, writes {"kernel": {"launches": n, "total_ms": t}, ...}
 * as JSON into buf (truncated to cap) and returns the size needed. */
MP_API int mp_profile_begin(void);
MP_API size_t mp_profile_end(char *buf, size_t cap);

/* ---- row 1: MultiPoint.detector_head, multipoint/models/MultiPoint.py:150-158 ------------
 * softmax over 65 channels, dustbin drop, PixelShuffle(8):
 *   prob[b,0,8h+i,8w+j] = softmax(logits[b,:,h,w])[8i+j]
 * logits (B,65,Hc,Wc); prob (B,1,8Hc,8Wc).  valid_mask (B,1,8Hc,8Wc) uint8 or NULL: when given,
 * prob is multiplied by it (the `prob * valid_mask` of predict_align_image_pair.py:127,132). */
MP_API int mp_detector_head_f32(const float *logits, int B, int Hc, int Wc,
                                const uint8_t *valid_mask, float *prob, mp_stream_t stream);

/* SuperPointMagicLeap.generate_heatmap, multipoint/models/SuperPointMagicLeap.py:68-85:
 * the same index map with the MagicLeap arithmetic, dense = exp(semi) / (sum_c exp(semi) + 1e-5), no
 * max subtraction, dustbin dropped.  semi (B,65,Hc,Wc); prob (B,1,8Hc,8Wc).  Replaces a per-sample
 * device->numpy->device round trip. */
MP_API int mp_heatmap_magicleap_f32(const float *semi, int B, int Hc, int Wc, float *prob, mp_stream_t stream);

/* utils.depth_to_space(x, block) / PixelShuffle, multipoint/utils/utils.py:64-69:
 * x (B, C*block^2, Hc, Wc) -> out (B, C, Hc*block, Wc*block). */
MP_API int mp_depth_to_space_f32(const float *x, int B, int C, int Hc, int Wc, int block,
                                 float *out, mp_stream_t stream);

/* ---- row 2: MultiPoint.descriptor_head tail, MultiPoint.py:160-166 -----------------------
 * F.normalize(x, p=2, dim=1).  x (B,D,HW).  out_nchw (B,D,HW) and/or out_nhwc (B,HW,D); either
 * may be NULL.  The channels-last copy feeds mp_sample_descriptors_f32 with coalesced rows. */
MP_API int mp_normalize_descriptors_f32(const float *x, int B, int D, int HW, float *out_nchw,
                                        float *out_nhwc, mp_stream_t stream);
/* (B,D,HW) -> channels-last (B,HW,D) copy of a descriptor map (no arithmetic): the layout change
 * utils.interpolate_descriptors (utils.py:159-167) makes once per call before sampling, because a
 * bilinear corner is then one contiguous D*4 B row instead of D strided words. */
MP_API int mp_transpose_descriptors_f32(const float *x, int B, int D, int HW, float *out_nhwc,
                                        mp_stream_t stream);

/* ---- row 4: utils.box_nms, multipoint/utils/utils.py:78-122 -------------------------------
 * Greedy IoU NMS of size x size boxes centred on every pixel with prob > min_prob (strict,
 * fp32), per image, priority (score desc, row-major index asc); optional top-k per image.
 * prob (B,H,W) -> prob_nms (B,H,W): surviving scores, zero elsewhere.
 * Optional ordered outputs (the torch.nonzero idiom of predict_align_image_pair.py:170-171):
 *   keypoints (B,kp_cap,2) int64 (y,x) in row-major order, kp_scores (B,kp_cap),
 *   kp_counts (B) = survivors per image (may exceed kp_cap: only kp_cap are written).
 * Pass NULL for the three to skip them.  prob_nms may be NULL when kp_counts is given (keypoints
 * only: the dense map is then neither zero-filled nor written -- the sync-free pipeline's case).
 * min_prob must be >= 0 (MP_ERR_UNSUPPORTED otherwise: probabilities are non-negative, and the
 * dense result cannot represent a kept zero). */
MP_API size_t mp_box_nms_workspace_bytes(int B, int H, int W);
MP_API int mp_box_nms_f32(const float *prob, int B, int H, int W, double size, double min_prob,
                          double iou, int keep_top_k, float *prob_nms, int64_t *keypoints,
                          float *kp_scores, int32_t *kp_counts, int kp_cap, void *workspace,
                          size_t workspace_bytes, mp_stream_t stream);

/* ---- row 4b: torch.nonzero((p > thr).float() [* mask]) -------------------------------------
 * predict_align_image_pair.py:170-171, evaluation.py:157-158,262-263, export_keypoints.py:100.
 * prob (B,H,W); mask (B,H,W) uint8 or NULL.  Outputs as above. */
MP_API size_t mp_extract_keypoints_workspace_bytes(int B, int H, int W);
MP_API int mp_extract_keypoints_f32(const float *prob, const uint8_t *mask, int B, int H, int W,
                                    double threshold, int64_t *keypoints, float *kp_scores,
                                    int32_t *kp_counts, int kp_cap, void *workspace,
                                    size_t workspace_bytes, mp_stream_t stream);

/* ---- row 5: utils.interpolate_descriptors, utils.py:159-167 -------------------------------
 * Bilinear sample (grid_sample, zeros padding, align_corners=True on y/(H/2)-1, x/(W/2)-1)
 * of the coarse descriptor map at each keypoint, then L2 normalisation.
 * keypoints (B,K,2) int64 (y,x); kp_counts (B) or NULL (= K valid per image);
 * desc (B,D,Hc,Wc) when layout==MP_LAYOUT_NCHW, (B,Hc,Wc,D) when MP_LAYOUT_NHWC;
 * out (B,K,D), rows beyond the image's count are zero-filled. */
#define MP_LAYOUT_NCHW 0
#define MP_LAYOUT_NHWC 1
MP_API int mp_sample_descriptors_f32(const int64_t *keypoints, const int32_t *kp_counts, int B,
                                     int K, const float *desc, int D, int Hc, int Wc, int layout,
                                     int H, int W, float *out, mp_stream_t stream);
/* The same, and the rows once more in the form the tensor-core matcher consumes (mp_match_split_f32): bf16 planes
 * hi = bf16(v), mid = bf16(v - hi), both (B,K,D), and the squared norms (B,K) of the rows as written.  Saves the
 * matcher its own pass over the descriptors.  Channels-last layout, D in {64, 128, 256}. */
MP_API int mp_sample_descriptors_split_f32(const int64_t *keypoints, const int32_t *kp_counts, int B,
                                           int K, const float *desc, int D, int Hc, int Wc, int layout,
                                           int H, int W, float *out, void *hi_bf16, void *mid_bf16,
                                           float *sq_norms, mp_stream_t stream);

/* ---- rows 6-8: utils.get_matches, multipoint/utils/matching.py:4-99 -----------------------
 * P independent problems (image pairs).  d1 (P,N1,D), d2 (P,N2,D); n1/n2 (P) device counts of
 * valid rows or NULL (= all).  D must be a multiple of 64 and <= 256 for the tensor-core path.
 *
 * mp_nearest_f32: for every row of d1 the nearest row of d2 and vice versa, with the exact
 * (fp64-verified) argmin and ties resolved to the lowest index like np.argmin / OpenCV:
 *   metric MP_METRIC_NN  sqrt(2 - 2*clip(a.b,-1,1))   NNMatcher, matching.py:50-53
 *   metric MP_METRIC_L2  sqrt(sum((a-b)^2))           cv2.BFMatcher(NORM_L2), matching.py:7
 * idx12 (P,N1) / idx21 (P,N2) int32 (-1 when the other set is empty); best12/second12 and
 * best21/second21 are the fp32 similarity a.b of the nearest and second nearest (may be NULL).
 * algo: MP_ALGO_TENSOR (tcgen05 split-bf16 GEMM + fp64 recheck of near-ties) or MP_ALGO_SIMT
 * (fp32 CUDA-core scan + the same recheck; reference implementation on the device). */
#define MP_METRIC_NN 0
#define MP_METRIC_L2 1
#define MP_ALGO_TENSOR 0
#define MP_ALGO_SIMT 1
MP_API size_t mp_match_workspace_bytes(int P, int N1, int N2, int D);
MP_API int mp_nearest_f32(const float *d1, const int32_t *n1, int N1, const float *d2,
                          const int32_t *n2, int N2, int P, int D, int metric, int algo,
                          int32_t *idx12, float *best12, float *second12, int32_t *idx21,
                          float *best21, float *second21, void *workspace,
                          size_t workspace_bytes, mp_stream_t stream);

/* mp_match_f32: the match list of get_matches in ascending query order.
 *   kind MP_MATCH_MUTUAL : keep i iff (!cross_check || idx21[idx12[i]] == i) and
 *                          (threshold < 0 || dist < threshold)        [bfmatcher / nnmatcher]
 *   kind MP_MATCH_RATIO  : keep i iff dist1 < ratio * dist2 (knn_matches=True, matching.py:21-28)
 * query/train (P,N1) int32, dist (P,N1) fp32 (recomputed in fp32 in the reference's own
 * formulation for the kept pairs), counts (P). */
#define MP_MATCH_MUTUAL 0
#define MP_MATCH_RATIO 1
MP_API int mp_match_f32(const float *d1, const int32_t *n1, int N1, const float *d2,
                        const int32_t *n2, int N2, int P, int D, int metric, int algo, int kind,
                        int cross_check, double threshold, double ratio, int32_t *query,
                        int32_t *train, float *dist, int32_t *counts, void *workspace,
                        size_t workspace_bytes, mp_stream_t stream);
/* mp_match_f32 on the tensor-core path with operands already split by mp_sample_descriptors_split_f32:
 * d (P,N,D) fp32 (exact recheck and distances), hi / mid (P,N,D) bf16, sq_norms (P,N), max_norm (P) = float bits of
 * an upper bound of the largest row norm of the pair's set (1.000001f for unit-norm descriptors). */
MP_API int mp_match_split_f32(const float *d1, const void *hi1, const void *mid1, const float *sq_norms1,
                              const uint32_t *max_norm1, const int32_t *n1, int N1, const float *d2,
                              const void *hi2, const void *mid2, const float *sq_norms2,
                              const uint32_t *max_norm2, const int32_t *n2, int N2, int P, int D, int metric,
                              int kind, int cross_check, double threshold, double ratio, int32_t *query,
                              int32_t *train, float *dist, int32_t *counts, void *workspace,
                              size_t workspace_bytes, mp_stream_t stream);

/* ThresholdMatcher.match, matching.py:74-99: every pair with sqrt(2-2clip(a.b)) < threshold in
 * row-major order.  One problem per call.  total_host receives the number of pairs found (the
 * call synchronises the stream); at most cap are written. */
MP_API int mp_match_threshold_f32(const float *d1, int N1, const float *d2, int N2, int D,
                                  double threshold, int32_t *query, int32_t *train, float *dist,
                                  int64_t cap, int64_t *total_host, void *workspace,
                                  size_t workspace_bytes, mp_stream_t stream);

/* ---- rows 9-10: homographic adaptation, multipoint/utils/homographies.py ------------------
 * mp_warp_f32 = warp_perspective_tensor (:404-425) for matrices already normalised on the host:
 * A (n_mats,3,3) maps destination [-1,1]^2 grid coordinates to source [-1,1]^2 coordinates;
 * xs (W) / ys (H) are the linspace(-1,1) tables.  src (N,H,W) is shared by all matrices,
 * out (n_mats,N,H,W).  mode MP_BILINEAR|MP_NEAREST, padding MP_PAD_ZEROS|MP_PAD_REFLECTION. */
#define MP_BILINEAR 0
#define MP_NEAREST 1
#define MP_PAD_ZEROS 0
#define MP_PAD_REFLECTION 1
MP_API int mp_warp_f32(const float *src, int N, int n_mats, int H, int W, const float *A,
                       const float *xs, const float *ys, int mode, int padding, float *out,
                       mp_stream_t stream);
/* The same warp with the N planes split into N / group groups that each get their own output block:
 * out (N / group, n_mats, group, H, W).  The adaptation of an image PAIR warps both spectra by the same matrices
 * (homographies.py:86 with the `H` of :81 for the optical and the thermal batch): one call with group = B shares the
 * per-pixel coordinate arithmetic between them and still hands each network a contiguous (n_mats * B, 1, H, W) batch.
 * With four or more matrices in bilinear mode the planes are first copied into a block-linear CUDA array owned by the
 * library (one per device, stream and width, kept for the life of the process) and interior footprints are fetched
 * with one tex2Dgather each; the values and the blend are the direct path's, bit for bit. */
MP_API int mp_warp_groups_f32(const float *src, int N, int group, int n_mats, int H, int W, const float *A,
                              const float *xs, const float *ys, int mode, int padding, float *out,
                              mp_stream_t stream);

/* mp_ha_aggregate_f32 = the unwarp + accumulate + finish of homographic_adaptation (:162-187)
 * and homographic_adaptation_multispectral (:77-126) over n pre-computed samples:
 *   count_i = nearest/zeros warp of masks[i] by Ainv[i];  count += count_i
 *   prob   += bilinear/zeros warp of probw[i] (a*b or a+b per source pixel when probw_b != NULL)
 *             * count_i
 * flags MP_HA_INIT  : start from prob0 and count = 1 (else from prob_acc / count_acc)
 *       MP_HA_FINISH: out = prob/count, sqrt ('prod') or *0.5 ('sum'), zero where count<min_count
 *                     (else the partial sums are stored to prob_acc / count_acc for an all-reduce)
 * probw_a/probw_b (n,B,H,W); masks (n,H,W) uint8; Ainv (n,3,3); out/prob_acc/count_acc (B,H,W). */
#define MP_AGG_NONE 0
#define MP_AGG_PROD 1
#define MP_AGG_SUM 2
#define MP_HA_INIT 1
#define MP_HA_FINISH 2
/* MP_HA_STAGED (optional): stage each sample's source window in shared memory with TMA instead of gathering
 * through L1.  Bit-identical results; measured slower than the direct gathers at 512x640,
 * kept as the measured alternative, not the default. */
#define MP_HA_STAGED 4
MP_API int mp_ha_aggregate_f32(const float *prob0, const float *probw_a, const float *probw_b,
                               const uint8_t *masks, const float *Ainv, int n, int B, int H, int W,
                               const float *xs, const float *ys, int aggregation, int min_count,
                               int flags, float *prob_acc, float *count_acc, float *out,
                               mp_stream_t stream);

/* ---- compute_valid_mask, multipoint/utils/homographies.py:375-402 ----
 * n masks in one launch: mask[i] = erode(cv2.warpPerspective(ones(H,W), Hm[i], INTER_NEAREST)).
 * Minv (n,3,3) float64 row-major, device: the INVERTED homographies exactly as cv2.warpPerspective
 * computes them (cv::invert closed form; multipoint_b200.utils.invert_homographies restates it).
 * erosion_radius r: (2r+1)^2 box minimum, r <= 31; outside-image pixels are ignored (cv2.erode's
 * default border) unless mask_border != 0, which is the reference's one-pixel zero frame.
 * mask (n,H,W) uint8 0/1 (4-byte aligned when W % 4 == 0).  Bit-exact with OpenCV's double
 * arithmetic (block origin, reciprocal, round-half-even). */
MP_API int mp_valid_mask_u8(const double *Minv, int n, int H, int W, int erosion_radius, int mask_border,
                            uint8_t *mask, mp_stream_t stream);

/* ---- point geometry of the evaluation loops, multipoint/utils/evaluation.py ----
 * All three take P problems (the samples of a batch) with a fixed capacity per problem and optional
 * device-side counts (NULL = every slot is live), the layout mp_box_nms_f32 / mp_extract_keypoints_f32
 * produce, so the stages chain without a host round trip.  Points are (y,x) like the reference's.
 *
 * warp_keypoints (multipoint/utils/homographies.py:331-346): cv2.perspectiveTransform of the flipped
 * points in double, flipped back.  kp (P,cap,2) int64; Hm (P,3,3) float64; out_f64 (P,cap,2) and/or
 * out_i64 (P,cap,2) = the reference's `.astype(int)` truncation.  Either output may be NULL. */
MP_API int mp_warp_keypoints_i64(const int64_t *kp, const int *counts, int P, int cap, const double *Hm,
                                 double *out_f64, int64_t *out_i64, mp_stream_t stream);

/* evaluation.py:176-197 for one direction: min_d2[p,i] = min_j |q[p,i] - t[p,j]|^2 as an exact int64
 * (np.linalg.norm of an integer difference is the square root of this), -1 where the query fails
 * filter_points (homographies.py:358-372: 0 <= y < H, 0 <= x < W), INT64_MAX where there is no target.
 * The repeatability count is sum(sqrt((double)min_d2) <= distance_thresh) over min_d2 >= 0. */
MP_API int mp_points_min_dist2_i64(const int64_t *q, const int *nq, int capq, const int64_t *t, const int *nt,
                                   int capt, int P, int H, int W, int64_t *min_d2, mp_stream_t stream);

/* evaluation.py:294-315: correct[i,j] = ||float32(qw[i] - t[j])||_2 <= threshold with qw (P,capq,2)
 * float64 warped points and t (P,capt,2) int64, never materialised: row_any[p,i] = any_j correct[i,j]
 * (what correct.sum(1).nonzero() counts, :301-302) and tp[p,k] = correct[mq[p,k], mt[p,k]] for the
 * matcher's pairs (:306-315; nm (P) counts or NULL, capm capacity).  row_any or tp may be NULL. */
MP_API int mp_points_correct_f32(const double *qw, const int *nq, int capq, const int64_t *t, const int *nt, int capt,
                                 int P, float threshold, uint8_t *row_any, const int *mq, const int *mt, const int *nm,
                                 int capm, uint8_t *tp, mp_stream_t stream);

/* ---- row 3: the elementwise glue between the backbone's cuDNN convolutions, MultiPoint.forward,
 * multipoint/models/MultiPoint.py:61-90 (getNonlinearity / getConvolutionBlock / generate_encoder).
 * One pass for  ReLU -> BatchNorm2d(eval) [-> MaxPool2d(2,2)] [-> ReflectionPad2d(1) | ZeroPad2d(1)]
 * (BatchNorm -> ReLU when bn_first): x (B,C,H,W) fp32 -> out (B,C,Ho+2*pad,Wo+2*pad), Ho = H/2 with
 * pool.  scale[c] = weight/sqrt(running_var+eps), shift[c] = bias - running_mean*scale (folded on
 * the host; <= 2 ulp from torch's eval BatchNorm).  conv_bias[c] (or NULL) is added to x first: the
 * convolution is then called without its bias, which torch would add in a separate elementwise
 * kernel.  pad in {0,1}; reflect selects the pad mode.
 * B*C <= 65535 per call. */
MP_API int mp_relu_bn_pad_f32(const float *x, int B, int C, int H, int W, const float *conv_bias, const float *scale,
                              const float *shift, int bn_first, int pool, int pad, int reflect, float *out,
                              mp_stream_t stream);

/* First encoder layer, MultiPoint.py:61-72,85-92: [ReflectionPad2d(1) | ZeroPad2d(1)] -> Conv2d(1 -> C, 3x3) ->
 * ReLU/BatchNorm2d(eval) [-> pad 1 for the next convolution] in one kernel (with one input channel the
 * convolution is bound by writing its output).  image (B,1,H,W); weight (C,1,3,3); conv_bias (C) or NULL;
 * scale/shift as in mp_relu_bn_pad_f32; in_reflect / out_reflect select the pad modes; out
 * (B,C,H+2*pad,W+2*pad).  W + 2*pad <= 768. */
MP_API int mp_conv1_relu_bn_pad_f32(const float *image, int B, int H, int W, const float *weight, const float *conv_bias,
                                    const float *scale, const float *shift, int C, int bn_first, int in_reflect, int pad,
                                    int out_reflect, float *out, mp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MULTIPOINT_B200_H */
