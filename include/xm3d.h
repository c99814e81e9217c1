/* xm3d.h — C ABI of libxm3d.so: the B200-native (sm_100a) cross-modal correspondence path
 * of XMask3D (voxelize -> project+occlusion -> mask gather/pool/scatter -> text logits).
 *
 * The reference (Zifeng-Zhang/XMask3D) has no FFI layer on this path: its boundary is a set
 * of Python call signatures.  Every entry point below names the reference call it replaces
 * (paths relative to the reference tree); the Python shims in xmask3d_b200/ keep those
 * signatures and call this ABI through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends
 *     in `_host`.  The caller owns all buffers (allocate them with torch / cudaMalloc).
 *   - all work is enqueued on `stream` (a cudaStream_t); no call synchronises the device.
 *     Data-dependent sizes (visible counts, voxel counts) are written to device memory.
 *   - return value: XM3D_OK or a negative xm3d_status; xm3d_last_error() has the text.
 *   - `status` (device int32, may be NULL): kernels OR XM3D_FLAG_* bits into it when a
 *     caller-provided capacity was too small; outputs are then truncated, never overrun.
 *   - batched layout: a batch is a list of `segments` (one per (scene, view)); per-point
 *     arrays of all segments are concatenated and addressed through offset arrays.
 *   - no CPU fallback exists: without a CUDA device every compute call fails.
 */
#ifndef XM3D_H_
#define XM3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XM3D_VERSION 100

#if defined(__GNUC__)
#define XM3D_API __attribute__((visibility("default")))
#else
#define XM3D_API
#endif

typedef void *xm3d_stream_t; /* cudaStream_t */

typedef enum {
    XM3D_OK = 0,
    XM3D_ERR_BAD_ARG = -1,
    XM3D_ERR_WORKSPACE = -2, /* workspace smaller than the matching *_ws_bytes() */
    XM3D_ERR_CUDA = -3,
    XM3D_ERR_UNSUPPORTED = -4
} xm3d_status;

#define XM3D_FLAG_VIS_OVERFLOW 1   /* more visible points than cap_vis          */
#define XM3D_FLAG_PAIR_OVERFLOW 2  /* more (point,mask) pairs than cap_pairs    */
#define XM3D_FLAG_GRID_RANGE 4     /* voxel coordinate outside +-2^30           */
#define XM3D_FLAG_KEY_SENTINEL 8   /* a key equal to 2^64-1 was remapped        */
#define XM3D_FLAG_I16_RANGE 16     /* xm3d_pack_i16: a value did not fit int16  */
#define XM3D_FLAG_VOX_FALLBACK 32  /* XM3D_VOX_FAST_ONLY: the batch needs the multi-kernel path */
#define XM3D_FLAG_NONFINITE 64     /* tensor-core pooling met a NaN / Inf feature: pool again with XM3D_POOL_ROWS */

/* depth image element type */
#define XM3D_DEPTH_NONE 0
#define XM3D_DEPTH_U16 1 /* raw PNG units; metres = value / depth_scale (float64 division)  */
#define XM3D_DEPTH_F64 2 /* metres, what the reference loader hands to compute_mapping      */

/* mask threshold modes of the three reference call sites */
#define XM3D_THR_GE_HALF 0         /* m >= 0.5            models/utils/fuser.py:16-17       */
#define XM3D_THR_SIGMOID_GE_HALF 1 /* sigmoid(m) >= 0.5   models/utils/criterion.py:83-85   */
#define XM3D_THR_SIGMOID_GT_HALF 2 /* sigmoid(m) >  0.5   models/xmask3d.py:356-357         */
#define XM3D_MASK_U8 0             /* bool / uint8 masks  */
#define XM3D_MASK_F32 1            /* float32 masks / logits */

XM3D_API int xm3d_version(void);
XM3D_API const char *xm3d_last_error(void);
/* SM count, compute capability of the current device (XM3D_ERR_CUDA without a GPU). */
XM3D_API int xm3d_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor);
/* Number of CUDA kernels this library has launched in this process (monotonic; for benchmarks). */
XM3D_API int64_t xm3d_launch_count(void);
/* Benchmark hook: two cudaEvent_t handles (or NULLs to clear) that the calling thread's next
 * xm3d_pool_batch calls record immediately before / after the dominant kernel (pool_sum_kernel)
 * on the call's stream, so its duration can be read with cudaEventElapsedTime. */
XM3D_API void xm3d_set_pool_events(void *ev_before, void *ev_after);

/* ------------------------------------------------------------------ stage 2: projection
 * Replaces PointCloudToImageMapper.compute_mapping (models/utils/fusion_util.py:46-142) and
 * the caller-side compaction (dataset/data_loader_infer.py:174-182, 263-268), batched over
 * views.  One record per view, 192 bytes, 16-byte aligned (staged into shared memory with a
 * bulk async copy). */
typedef struct xm3d_view {
    double w2c[12];    /* rows 0..2 of world_to_camera = np.linalg.inv(pose), row-major 3x4  */
    double fx, fy, cx, cy;
    int64_t pt_off;    /* first point of this view's scene in `xyz` (points, not floats)      */
    int64_t out_off;   /* first element of this view in the per-(view,point) outputs          */
    int64_t depth_off; /* element offset of this view's depth image in `depth`, <0 = no depth */
    int32_t n_pts;     /* points of the scene                                                 */
    int32_t depth_h, depth_w;
    int32_t reserved0;
    int64_t reserved1[3];
} xm3d_view_t;

XM3D_API size_t xm3d_project_ws_bytes(int32_t n_views, int64_t total_pts, int32_t max_pts_per_view);

/* views_host: HOST array of n_views records, read for launch planning.  views_dev: the same
 * records already in device memory, or NULL — then the library uploads views_host on `stream`
 * (a pageable-memory copy: pass views_dev when the call is captured into a CUDA graph).
 * For every view v and scene point i (i < views[v].n_pts):
 *   vis[out_off+i]        1 if the point projects inside the cut image and passes the depth test
 *   mapping[(out_off+i)*3 + {0,1,2}] = (pixel row, pixel col, vis)   (optional, the drop-in
 *                          int64 [N,3] array of compute_mapping; zero rows when invisible)
 * and the order-preserving compaction of the visible points of each view:
 *   n_vis[v], vis_off[0..V]   counts and their exclusive prefix (vis_off[V] = total)
 *   vis_idx[vis_off[v]+j]     index (within the scene) of the j-th visible point
 *   rowcol[(vis_off[v]+j)*2]  pixel row (x_label), pixel col (y_label)
 *   xyz_vis[(vis_off[v]+j)*3] its coordinates (locals_3d)
 * Compacted outputs hold at most cap_vis points in total (XM3D_FLAG_VIS_OVERFLOW otherwise).
 * vis_idx / rowcol / xyz_vis / mapping may be NULL.  total_pts = sum of n_pts over views
 * (size of vis).  depth: uint16 raw units (metres = value / depth_scale, float64 division — the
 * loader's imread(png)/1000, dataset/data_loader_infer.py:168-171) or float64 metres. */
XM3D_API int xm3d_project_batch(const float *xyz, const xm3d_view_t *views_host, const xm3d_view_t *views_dev,
                       int32_t n_views, int64_t total_pts, const void *depth, int32_t depth_kind, double depth_scale,
                       int32_t img_w, int32_t img_h, int32_t cut_bound, double vis_thres,
                       uint8_t *vis, int64_t *mapping, int32_t *n_vis, int64_t *vis_off,
                       int64_t cap_vis, int32_t *vis_idx, int32_t *rowcol, float *xyz_vis,
                       void *ws, size_t ws_bytes, int32_t *status, xm3d_stream_t stream);

/* ------------------------------------------------------------------ stage 1: voxelization
 * xm3d_unique_batch replaces np.unique(key, return_index, return_inverse, return_counts) as
 * used by sparse_quantize (dataset/voxelization_utils.py:86, :95): per segment, unique keys in
 * ascending order, index of the first occurrence, rank of every element, multiplicity.
 * Segment s owns elements [seg_off[s], seg_off[s+1]) of `keys`; seg_off is a DEVICE int64
 * [n_seg+1] array (seg_off[0] = 0) so a projection can feed it without a host round trip;
 * `cap` is the host-known bound on seg_off[n_seg] (all per-element buffers hold cap entries;
 * if the bound is exceeded XM3D_FLAG_VIS_OVERFLOW is raised and nothing is processed).
 *   m[s], uniq_off[0..n_seg]  unique count per segment and exclusive prefix
 *   first[uniq_off[s]+r]      index within the segment of the first occurrence of rank r
 *   counts[uniq_off[s]+r]     multiplicity (optional)
 *   inverse[seg_off[s]+i]     rank of element i (+ uniq_off[s] if collate != 0, the
 *                             cumulative offset collation_fn adds, dataset/data_loader.py:341-342) */
XM3D_API size_t xm3d_unique_ws_bytes(int32_t n_seg, int64_t cap);
/* `path` (per call, re-entrant — both paths are bit-identical; tests and A/B timings use it): low byte
 * XM3D_VOX_AUTO = shared-memory units whenever every segment fits (<= 224 k points), the multi-kernel
 * global-memory path otherwise; XM3D_VOX_MULTI_KERNEL = multi-kernel path only.  Bits 8.. = points per unit the
 * plan aims at (0 = default 7000; smaller values force several key-range units per segment). */
#define XM3D_VOX_AUTO 0
#define XM3D_VOX_MULTI_KERNEL 1
#define XM3D_VOX_FAST_ONLY 2   /* shared-memory units only: the fallback kernels are not launched (they cost ~5 us each just
                                * to return); a batch that is not eligible or a unit that overflows raises
                                * XM3D_FLAG_VOX_FALLBACK and the outputs are undefined — call again with XM3D_VOX_AUTO */
XM3D_API int xm3d_unique_batch(const uint64_t *keys, const int64_t *seg_off, int32_t n_seg, int64_t cap,
                      int32_t *m, int64_t *uniq_off, int32_t *first, int32_t *counts,
                      int32_t *inverse, int32_t collate, int32_t path, void *ws, size_t ws_bytes,
                      int32_t *status, xm3d_stream_t stream);

/* FNV-1 over whole 64-bit words per row (fnv_hash_vec, dataset/voxelization_utils.py:6-18)
 * and the mixed-radix ravel key (ravel_hash_vec, :21-35) of float64 rows (dim <= 8 for ravel). */
XM3D_API int xm3d_fnv_hash_f64(const double *coords, int64_t n, int32_t dim, uint64_t *keys, xm3d_stream_t stream);
XM3D_API int xm3d_ravel_hash_f64(const double *coords, int64_t n, int32_t dim, uint64_t *keys, void *ws,
                        size_t ws_bytes, xm3d_stream_t stream);
XM3D_API size_t xm3d_ravel_ws_bytes(int32_t dim);

/* Voxelizer.voxelize after the matrix is drawn (dataset/voxelizer.py:110-122) fused with
 * sparse_quantize: grid = floor([x y z 1] @ RT.T[:, :3]); grid -= min; FNV key; unique.
 *   xyz           [cap,3] float32 (xyz_f64 = 0) or float64 (xyz_f64 = 1: what ElasticDistortion hands the
 *                 voxelizer on the augmented training path, dataset/augmentation.py:171), segments as above
 *                 (e.g. xyz_vis / vis_off of the projection)
 *   rt            [n_seg,12] float64: rows 0..2 of rigid_transformation per segment
 *   grid_min      [n_seg,3] int32   column minima that were subtracted (optional)
 *   voxel_xyz     [uniq_off[s]+r, 3] int32 voxel coordinates in unique order (grid[inds])
 * other outputs as xm3d_unique_batch. */
XM3D_API size_t xm3d_voxelize_ws_bytes(int32_t n_seg, int64_t cap);
/* Which path the last call on this workspace took (synchronises the stream; diagnostics / tests):
 * ctl_host[0] = 1 if the batch was not eligible for the shared-memory path, ctl_host[1] = 1 if a
 * unit overflowed and the batch was recomputed by the multi-kernel path. */
XM3D_API int xm3d_voxel_path_info(const void *ws, int32_t n_seg, int64_t cap, int32_t *ctl_host,
                         xm3d_stream_t stream);
XM3D_API int xm3d_voxelize_batch(const void *xyz, int32_t xyz_f64, const int64_t *seg_off, int32_t n_seg, int64_t cap,
                        const double *rt, int32_t *m, int64_t *uniq_off, int32_t *first,
                        int32_t *inverse, int32_t collate, int32_t *voxel_xyz, int32_t *grid_min,
                        int32_t path, void *ws, size_t ws_bytes, int32_t *status, xm3d_stream_t stream);

/* ------------------------------------------------------------------ stage 3: masks at points
 * Mask-at-point gather + threshold (models/utils/fuser.py:16-17, models/utils/criterion.py:83-85,
 * models/xmask3d.py:356-358).  masks: [n_seg, k, h, w] (uint8 or float32); rowcol as produced by
 * xm3d_project_batch (row = x_label, col = y_label); seg_off / cap as in stage 1.
 * member: per point xm3d_mask_words(k) = ceil(k/32) uint32 words, bit m%32 of word m/32 set iff
 * the point is in mask m.  counts (optional): [n_seg,k] int32 points per mask.  k <= 256. */
XM3D_API int32_t xm3d_mask_words(int32_t k);
XM3D_API size_t xm3d_gather_ws_bytes(int32_t n_seg, int32_t k, int32_t h, int32_t w);
XM3D_API int xm3d_gather_masks_batch(const void *masks, int32_t mask_kind, int32_t thr_mode, int32_t n_seg,
                            int32_t k, int32_t h, int32_t w, const int32_t *rowcol,
                            const int64_t *seg_off, int64_t cap, uint32_t *member, int32_t *counts,
                            void *ws, size_t ws_bytes, xm3d_stream_t stream);
/* First half of xm3d_gather_masks_batch alone: masks -> per-pixel membership words
 * pixbits [n_seg, xm3d_mask_words(k), h*w] (the input of xm3d_point_bits_batch).  It does not depend on
 * the projection, so a pipeline can run it NEXT TO the projection (xmask3d_b200/pipeline.py). */
XM3D_API int xm3d_pixel_bits_batch(const void *masks, int32_t mask_kind, int32_t thr_mode, int32_t n_seg, int32_t k,
                          int32_t h, int32_t w, uint32_t *pixbits, xm3d_stream_t stream);

/* Segmented mean pooling of per-point features under each mask (models/utils/criterion.py:148-157;
 * scalar form models/xmask3d.py:362-367 with c = 1).
 *   feat      [rows, c] float32;  row_index (optional int32 [cap]) maps point -> feature row
 *             (e.g. inds_reconstruct, fusing pred_3d[inds_reconstruct], models/xmask3d.py:152)
 *   member    as produced by xm3d_gather_masks_batch (general, overlapping masks) OR
 *   label     int32 [cap], values outside [0,k) = in no mask (partition masks); exactly one of the two
 *   cap_pairs bounds the total number of (point, mask) memberships (= cap for labels / partition
 *   masks); beyond it XM3D_FLAG_PAIR_OVERFLOW is raised and the sums are zero (pair-list path only).
 *   path      XM3D_POOL_AUTO picks the kernel: partition masks / labels -> sorted pair lists + register
 *             accumulation (every row read once, HBM peak); cap_pairs > cap + 1 tells the library that masks may
 *             overlap -> every row is still read exactly once by the TENSOR-CORE kernel (tcgen05 tf32, member bits
 *             as a 0/1 operand; needs member words, no row_index, c % 128 == 0, k <= 128, 16-byte aligned sum / mean; a NaN / Inf
 *             feature would leak
 *             into the other masks of its 64-point tile (0 x Inf), so the kernel raises XM3D_FLAG_NONFINITE when it
 *             meets one and the caller pools that batch again with XM3D_POOL_ROWS), else by
 *             the point-major CUDA-core kernel (c % 128 == 0, k <= 96), else by the pair lists (one row read per
 *             membership).  XM3D_POOL_PAIR_LISTS / _ROWS / _MMA force one path (XM3D_ERR_UNSUPPORTED if not eligible).
 *   sum [n_seg,k,c] float32, cnt [n_seg,k] int32 (optional), mean (optional) [n_seg,k,c] = sum/cnt
 *   (0 where cnt = 0).  Deterministic: every summation order is fixed by the point order. */
#define XM3D_POOL_AUTO 0
#define XM3D_POOL_PAIR_LISTS 1
#define XM3D_POOL_ROWS 2
#define XM3D_POOL_MMA 3
XM3D_API size_t xm3d_pool_ws_bytes(int32_t n_seg, int32_t k, int32_t c, int64_t cap, int64_t cap_pairs);
XM3D_API int xm3d_pool_batch(const float *feat, int32_t c, const int32_t *row_index, const uint32_t *member,
                    const int32_t *label, int32_t n_seg, int32_t k, const int64_t *seg_off, int64_t cap,
                    int64_t cap_pairs, int32_t path, float *sum, int32_t *cnt, float *mean, void *ws,
                    size_t ws_bytes, int32_t *status, xm3d_stream_t stream);

/* Mask -> point scatter-mean (mask_mapper, models/utils/fuser.py:22-34; twin
 * models/xmask3d.py:441-455): out[i,:] = (sum of emb[m,:] over masks m containing i, ascending m)
 * / counter_i, counter 0 -> 1e-5; bit-exact with the reference's float32 op order.
 *   emb [n_seg,k,c], out [cap,c], counter (optional) [cap] float32 */
XM3D_API int xm3d_scatter_batch(const uint32_t *member, const int32_t *label, int32_t n_seg, int32_t k,
                       const int64_t *seg_off, int64_t cap, const float *emb, int32_t c, float *out,
                       float *counter, xm3d_stream_t stream);

/* loss_contra mask selection (models/utils/criterion.py:80-146) for every scene of a batch, no host round trip:
 *   member      [cap, words] = sigmoid(mask[:, x_label, y_label]) >= 0.5 (xm3d_gather_masks_batch with
 *               XM3D_THR_SIGMOID_GE_HALF); binary_gt [cap] float32 (0 = novel, 1 = base, anything else ignored)
 *   mask_logits [n_seg, k, h, w] float32, already up-sampled to cfg.mask_shape (:53-55)
 * The ">= 10 points, else row 0 all True" guard (:87-88) and keep (:90) are applied on the device.  Outputs:
 *   counts [n_seg,k,3] int32   (points, of which binary_gt == 0, of which == 1) per mask AFTER the guard
 *   kind   [n_seg,k]   int8    0 = not pooled, 1 = novel candidate (:109-112), 2 = base candidate (:114-117)
 *   score  [n_seg,k]   float32 mean of sigmoid(mask) over the pixels with sigmoid > 0.5 (candidates; NaN elsewhere)
 *   sel    [n_seg,5]   int32   masks to pool: up to 4 novel by descending score (stable), then up to 1 base; -1 padded
 *   n_sel  [n_seg]     int32
 *   sel_member [cap]   uint32  bit j set iff the point lies in sel[.,j] -> xm3d_pool_batch(member = sel_member, k = 5)
 *               gives feature[mask_3d[fidx]].mean(0) of :148-157 for all selected masks of the batch. */
XM3D_API size_t xm3d_contra_ws_bytes(int32_t n_seg, int32_t k);
XM3D_API int xm3d_contra_select_batch(const uint32_t *member, int32_t k, const float *binary_gt, const int64_t *seg_off,
                             int32_t n_seg, int64_t cap, const float *mask_logits, int32_t h, int32_t w,
                             int32_t *counts, int8_t *kind, float *score, int32_t *sel, int32_t *n_sel,
                             uint32_t *sel_member, void *ws, size_t ws_bytes, xm3d_stream_t stream);

/* ------------------------------------------------------------------ stage 4: text logits
 * XMASK3d.cal_pred_logits (models/xmask3d.py:129-143) + ensemble_logits_with_labels
 * (models/modeling/meta_arch/helper.py:72-97, "max" or "mean"):
 *   out[r, g] = reduce_{t in group g} scale * <mask_embed[r]/|.|, text_embed[t]/|.|>,
 *   out[r, n_groups] = scale * <mask_embed[r]/|.|, null_embed/|.|>
 * rows = B*K mask embeddings [rows,c]; text_embed [n_text,c]; group_off_host: HOST [n_groups+1]
 * column offsets of the synonym groups; out [rows, n_groups+1]; argmax (optional) int32 [rows].
 * n_text + 1 <= 256, c % 4 == 0, every group non-empty.
 * The contraction runs on the tensor cores (tcgen05, 3xTF32 split, fp32 accumulate in TMEM). */
XM3D_API size_t xm3d_logits_ws_bytes(int64_t rows, int32_t n_text, int32_t c, int32_t n_groups);
XM3D_API int xm3d_logits(const float *mask_embed, int64_t rows, int32_t c, const float *text_embed,
                int32_t n_text, const float *null_embed, const int32_t *group_off_host,
                int32_t n_groups, int32_t ensemble_mean, float logit_scale, float *out,
                int32_t *argmax, void *ws, size_t ws_bytes, xm3d_stream_t stream);

/* ------------------------------------------------------------------ after the path: per-point logits
 * logit_scale * (normalize(feat) @ normalize(text_embed).T) for every point, optional base / novel
 * blending with the binary head and argmax (run/infer.py:557, 606-640; models/utils/criterion.py:184-207):
 *   val[i,t] = scale * <feat_i/|feat_i|, text_t/|text_t|>
 *   binary given:  val[i,t] = b_i * (is_base[t] ? val : -1e10) + (1 - b_i) * (is_base[t] ? -1e10 : val)
 *   mask_label given (the FUSED stream, run/infer.py:568-600): p_i = softmax_t(val[i,:]); for a point inside final
 *     mask m = mask_label[i] (>= 0; the final masks are an argmax partition, models/xmask3d.py:418-435) with
 *     q = softmax(logit_scale * normalize(final_pred_open_embedding) @ text.T)[m,:], passed as mask_log_probs = log q:
 *       val[i,t] = log(p^base_ratio * q^(1-base_ratio)) * ov[t] + log(p^novel_ratio * q^(1-novel_ratio)) * (1 - ov[t]),
 *     ov = is_base, evaluated in the log domain (ratio * log p + (1 - ratio) * log q; equal up to float32 rounding, finite
 *     where the reference's p underflows); points in no mask keep val = p.  The binary blend above then applies.
 *   out [rows, n_text] float32 (optional), argmax int32 [rows] (optional, first maximum)
 * feat [rows, c] float32 is read from HBM once (TMA -> in-place TF32 hi/lo split in shared memory ->
 * tcgen05 3xTF32); n_text <= 256, c % 4 == 0, feat 16-byte aligned. */
XM3D_API size_t xm3d_point_logits_ws_bytes(int32_t n_text, int32_t c);
XM3D_API int xm3d_point_logits(const float *feat, int64_t rows, int32_t c, const float *text_embed, int32_t n_text,
                      float logit_scale, const float *binary, const uint8_t *is_base,
                      const int32_t *mask_label, const float *mask_log_probs, int32_t n_masks, float base_ratio,
                      float novel_ratio, float *out, int32_t *argmax, void *ws, size_t ws_bytes,
                      xm3d_stream_t stream);

/* ------------------------------------------------------------------ after the path: votes
 * Cross-view vote accumulation of the inference loop (run/infer.py:642-647, :658), batched:
 *   votes[p, cls[j]] += 1, counter[p] += 1   for every visible pair j, p = view_pt_off[seg(j)] + vis_idx[j]
 * vis_idx / seg_off as produced by xm3d_project_batch; view_pt_off: DEVICE int64 [n_seg], first
 * point of each view's scene in the scene-point arrays; cls: predicted class of every pair
 * (values outside [0,n_classes) only count the visit).  votes int32 [n_scene_pts, n_classes] and
 * counter int32 [n_scene_pts] are accumulated into (zero them once per scene).
 * xm3d_vote_argmax: pred[p] = first maximum of votes[p,:] (torch.max semantics), -1 if counter[p] == 0. */
XM3D_API int xm3d_vote_batch(const int32_t *vis_idx, const int64_t *seg_off, int32_t n_seg, int64_t cap,
                    const int64_t *view_pt_off, const int32_t *cls, int32_t n_classes, int32_t *votes,
                    int32_t *counter, xm3d_stream_t stream);
XM3D_API int xm3d_vote_argmax(const int32_t *votes, const int32_t *counter, int64_t n_pts, int32_t n_classes,
                     int32_t *pred, xm3d_stream_t stream);

/* Nearest SEEN neighbour of every point no view has seen (run/infer.py:651-656, 684-694: the reference
 * builds sklearn KDTree(scene_coords[counter != 0]) and queries scene_coords[counter == 0] with k = 1,
 * then copies the neighbour's prediction: scene_pred[false_idx] = scene_pred[true_idx[indices]]).
 * Batched over scenes: seg_off DEVICE int64 [n_seg+1] point offsets, counter int32 [n_total] (seen iff != 0).
 *   match[seg_off[s]+i] = index inside scene s of the nearest seen point (Euclidean, float64 arithmetic on
 *   the float32 coordinates; ties -> lowest index); i itself for a seen point; -1 if the scene has no
 *   seen point.  Filling is then pred[seg_off[s]+i] = pred[seg_off[s]+match[...]]. */
XM3D_API size_t xm3d_nn_fill_ws_bytes(int32_t n_seg, int64_t n_total);
XM3D_API int xm3d_nn_fill_batch(const float *xyz, const int32_t *counter, const int64_t *seg_off, int32_t n_seg,
                       int64_t n_total, int32_t *match, void *ws, size_t ws_bytes, xm3d_stream_t stream);

/* Per-scene maximum of the sparse bottleneck features (models/xmask3d.py:154-159:
 * torch.max(imp_condition[_idx_ == scene_idx], dim=0)[0] for every scene): feat [rows, c] float32 with the
 * rows of scene s at [seg_off[s], seg_off[s+1]) (collation order), out [n_seg, c].  NaN propagates like
 * torch.max; an empty scene gives -inf. */
XM3D_API int xm3d_segment_max(const float *feat, const int64_t *seg_off, int32_t n_seg, int32_t c, float *out,
                     xm3d_stream_t stream);

/* ------------------------------------------------------------------ after the path: mask preparation
 * The dense torch sequence that turns the mask head's low-resolution logits into the masks the path
 * consumes (models/xmask3d.py:326-331, 356-358, 391-435; models/utils/criterion.py:239-244, 273-320),
 * fused into one pass per output pixel:
 *   up  = F.interpolate(logits [k,hs,ws] -> [k,h,w], "bilinear", align_corners=False)   (torch's CPU
 *         arithmetic bit for bit)
 *   pixbits (optional)  [n_seg, words, h*w] uint32: bit m set iff threshold(up[m]) per thr_mode
 *                       (XM3D_THR_SIGMOID_GT_HALF = `.sigmoid() > 0.5`); feed xm3d_point_bits_batch
 *   label   (optional)  [n_seg, h*w] int16: m = argmax over kept masks of scores[m] * sigmoid(up[m])
 *                       (first maximum) if sigmoid(up[m]) >= 0.5, else -1 — the final partition masks
 *                       (cur_mask_ids == m) & (cur_masks[m] >= 0.5); feed xm3d_gather_labels_batch
 *   areas   (optional)  [n_seg, k, 3] int32: mask_area, original_area, intersection; the reference
 *                       keeps mask m iff all three are > 0 (<=> intersection > 0)
 *   upsampled (optional) [n_seg, k, h*w] float32 (tests / callers that want the dense tensor)
 *   scores [n_seg,k] (null = 1), keep [n_seg,k] uint8 (null = all): `scores > thresh` of the caller. */
XM3D_API int xm3d_mask_prep_batch(const float *logits, int32_t n_seg, int32_t k, int32_t hs, int32_t ws, int32_t h,
                         int32_t w, const float *scores, const uint8_t *keep, int32_t thr_mode,
                         uint32_t *pixbits, int16_t *label, int32_t *areas, float *upsampled,
                         xm3d_stream_t stream);
/* member words of every visible point from per-pixel words (second half of xm3d_gather_masks_batch) */
XM3D_API int xm3d_point_bits_batch(const uint32_t *pixbits, int32_t n_seg, int32_t k, int32_t h, int32_t w,
                          const int32_t *rowcol, const int64_t *seg_off, int64_t cap, uint32_t *member,
                          int32_t *counts, xm3d_stream_t stream);
/* point_label[i] = label image of the point's segment at (row, col); -1 outside the image.  The result
 * is the `label` input of xm3d_pool_batch / xm3d_scatter_batch (mask_3d = mask[:, x_label, y_label]). */
XM3D_API int xm3d_gather_labels_batch(const int16_t *label_img, int32_t n_seg, int32_t h, int32_t w,
                             const int32_t *rowcol, const int64_t *seg_off, int64_t cap,
                             int32_t *point_label, xm3d_stream_t stream);

/* ------------------------------------------------------------------ after the path: batch layout
 * collation_fn (dataset/data_loader.py:319-357) on the device, from the outputs of the stages above:
 *   ori_coords (optional) [total visible, 4] float32: (batch item, x, y, z)   from xyz_vis / vis_off
 *   coords     (optional) [total voxels, 4]  int32:   (batch item, vx, vy, vz) from voxel_xyz / uniq_off
 * cap bounds both row counts; inds_reconstruct with the cumulative voxel offset is the `inverse`
 * output of xm3d_voxelize_batch(collate = 1); x_label / y_label are the columns of rowcol. */
XM3D_API int xm3d_collate_batch(const float *xyz_vis, const int64_t *vis_off, const int32_t *voxel_xyz,
                       const int64_t *uniq_off, int32_t n_seg, int64_t cap, float *ori_coords,
                       int32_t *coords, xm3d_stream_t stream);
/* int32 [rows, width] -> int16 for the host-bound copies of x_label / y_label (rowcol, values < 320) and voxel
 * coordinates (dataset/data_loader.py:319-357 hands them to the host loop as int64 / int32; the end-to-end path is
 * PCIe bound, so bytes matter).  rows_dev: optional DEVICE row count (vis_off[n_seg] / uniq_off[n_seg]), capped by
 * cap_rows.  Out-of-range values are clamped and flagged (XM3D_FLAG_I16_RANGE). */
XM3D_API int xm3d_pack_i16(const int32_t *src, const int64_t *rows_dev, int64_t cap_rows, int32_t width, int16_t *dst,
                  int32_t *status, xm3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* XM3D_H_ */
