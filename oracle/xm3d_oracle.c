/* ORACLE — test infrastructure, NOT product code.
 *
 * Scalar C restatement of the integer-deciding part of the reference's cross-modal
 * correspondence path (projection + occlusion, voxel quantisation, hashing, unique /
 * inverse maps) and exact / high-precision forms of the mask pooling and scatter.
 * It exists so the CUDA kernels can be checked on a box where /root/reference and its
 * BLAS are absent: every float->int decision is written as the exact IEEE operation
 * sequence (FMA chain for the 4-term dot products, individually rounded ops after it).
 *
 * Parity status: pinned against the reference's own functions executed in the build
 * container (tests/golden/make_golden.py -> tests/golden/ fixtures, replayed by
 * tests/test_oracle_golden.py).  The reference has no golden vectors of its own.
 *
 * Compile: see oracle/Makefile  (-O2 -ffp-contract=off -mfma: contraction must be off so
 * that only the explicit fma() calls fuse).
 * Reference lines are cited per function, relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define XO_API __attribute__((visibility("default")))

/* 4-term dot product exactly as numpy's float64 matmul (OpenBLAS dgemm micro-kernel)
 * evaluates it: one accumulator per output, FMA in k order, first term a plain product. */
static inline double dot4(const double *a, double x, double y, double z, double w) {
    double s = a[0] * x;
    s = fma(a[1], y, s);
    s = fma(a[2], z, s);
    s = fma(a[3], w, s);
    return s;
}

/* ---------------------------------------------------------------- stage 2: projection
 * models/utils/fusion_util.py:46-142 (compute_mapping).
 * depth_kind: 0 = no depth (depth=None), 1 = uint16 raw / depth_scale (the loader's
 * `imread(png) / 1000`, dataset/data_loader_infer.py:168-171), 2 = float64 metres.
 * mapping: int64 [n,3] rows (pixel row, pixel col, visible), zero rows when invisible. */
XO_API int xo_project(const float *xyz, int64_t n, const double *w2c,
                      double fx, double fy, double cx, double cy,
                      const void *depth, int depth_kind, int dh, int dw, double depth_scale,
                      int img_w, int img_h, int cut, double vis_thres, int64_t *mapping) {
    uint8_t *inside = (uint8_t *)calloc((size_t)(n ? n : 1), 1);
    uint8_t *okdep = (uint8_t *)calloc((size_t)(n ? n : 1), 1);
    int64_t *ix = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
    int64_t *iy = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
    if (!inside || !okdep || !ix || !iy) return -1;
    int any_inside = 0, any_indepth = 0;
    for (int64_t i = 0; i < n; ++i) {
        double x = (double)xyz[3 * i], y = (double)xyz[3 * i + 1], z = (double)xyz[3 * i + 2];
        double p0 = dot4(w2c + 0, x, y, z, 1.0);                       /* :71 */
        double p1 = dot4(w2c + 4, x, y, z, 1.0);
        double p2 = dot4(w2c + 8, x, y, z, 1.0);
        double zdiv = (fabs(p2) < 1e-8) ? 1.0 : p2;                    /* :75-76 */
        double px = (p0 * fx) / zdiv + cx;                             /* :78 */
        double py = (p1 * fy) / zdiv + cy;                             /* :79 */
        double rx = rint(px), ry = rint(py);                           /* :82-83 half-to-even */
        /* astype(int): out-of-range / NaN become INT64_MIN on x86 (cvttsd2si) */
        ix[i] = (rx >= -9.2233720368547758e18 && rx < 9.2233720368547758e18) ? (int64_t)rx : INT64_MIN;
        iy[i] = (ry >= -9.2233720368547758e18 && ry < 9.2233720368547758e18) ? (int64_t)ry : INT64_MIN;
        int in = (p2 > 0) && ix[i] >= cut && iy[i] >= cut &&           /* :86-95 */
                 ix[i] < (int64_t)img_w - cut && iy[i] < (int64_t)img_h - cut;
        inside[i] = (uint8_t)in;
        if (in) {
            any_inside = 1;
            if (depth_kind != 0) {
                int valid = iy[i] >= 0 && iy[i] < dh && ix[i] >= 0 && ix[i] < dw;   /* :105-110 */
                if (valid) {
                    any_indepth = 1;
                    double d;
                    if (depth_kind == 1) d = (double)((const uint16_t *)depth)[iy[i] * dw + ix[i]] / depth_scale;
                    else d = ((const double *)depth)[iy[i] * dw + ix[i]];
                    okdep[i] = (uint8_t)(fabs(d - p2) <= vis_thres * d);            /* :125 */
                }
            }
        }
    }
    /* :98, :115-135 — the occlusion result replaces `inside` only if at least one inside
     * point fell within the depth image. */
    int use_depth = depth_kind != 0 && any_inside && any_indepth;
    for (int64_t i = 0; i < n; ++i) {
        int v = use_depth ? (inside[i] && okdep[i]) : inside[i];
        mapping[3 * i + 0] = v ? iy[i] : 0;                           /* :138-140 */
        mapping[3 * i + 1] = v ? ix[i] : 0;
        mapping[3 * i + 2] = v ? 1 : 0;
    }
    free(inside); free(okdep); free(ix); free(iy);
    return 0;
}

/* ---------------------------------------------------------------- stage 1: hashing
 * dataset/voxelization_utils.py:6-18 (fnv_hash_vec): FNV-1 over whole uint64 words. */
static inline uint64_t fnv1_words(const uint64_t *w, int dim) {
    uint64_t h = 14695981039346656037ULL;
    for (int j = 0; j < dim; ++j) { h *= 1099511628211ULL; h ^= w[j]; }
    return h;
}

XO_API void xo_fnv(const double *coords, int64_t n, int dim, uint64_t *out) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t w[8];
        for (int j = 0; j < dim && j < 8; ++j) w[j] = (uint64_t)coords[i * dim + j];
        out[i] = fnv1_words(w, dim);
    }
}

typedef struct { uint64_t key; int64_t idx; } xo_pair;
static int cmp_pair(const void *a, const void *b) {
    const xo_pair *p = (const xo_pair *)a, *q = (const xo_pair *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx);
}

/* np.unique(key, return_index=True, return_inverse=True, return_counts=True)
 * (dataset/voxelization_utils.py:86, :95): unique keys ascending, index of the FIRST
 * occurrence, rank of every element's key, multiplicity. */
XO_API int xo_unique_u64(const uint64_t *keys, int64_t n, int64_t *first, int64_t *inverse,
                         int64_t *counts, int64_t *m_out) {
    xo_pair *p = (xo_pair *)malloc(sizeof(xo_pair) * (size_t)(n ? n : 1));
    if (!p) return -1;
    for (int64_t i = 0; i < n; ++i) { p[i].key = keys[i]; p[i].idx = i; }
    qsort(p, (size_t)n, sizeof(xo_pair), cmp_pair);
    int64_t m = 0;
    for (int64_t j = 0; j < n; ++j) {
        if (j == 0 || p[j].key != p[j - 1].key) {
            if (first) first[m] = p[j].idx;
            if (counts) counts[m] = 0;
            ++m;
        }
        if (counts) counts[m - 1] += 1;
        if (inverse) inverse[p[j].idx] = m - 1;
    }
    *m_out = m;
    free(p);
    return 0;
}

/* dataset/voxelizer.py:110-122 (voxelize, after the matrix is drawn):
 *   grid = floor([x y z 1] @ RT.T[:, :3]); grid = floor(grid - grid.min(0));
 *   first, inverse = sparse_quantize(grid)  -> fnv + unique;  voxel_xyz = grid[first].
 * rt: row-major 4x4 float64 (rigid_transformation). Buffers sized for n. */
XO_API int xo_voxelize(const float *xyz, int64_t n, const double *rt,
                       int64_t *first, int64_t *inverse, double *voxel_xyz, int64_t *m_out) {
    double *g = (double *)malloc(sizeof(double) * 3 * (size_t)(n ? n : 1));
    uint64_t *key = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n ? n : 1));
    if (!g || !key) return -1;
    double mn[3] = {INFINITY, INFINITY, INFINITY};
    for (int64_t i = 0; i < n; ++i) {
        double x = (double)xyz[3 * i], y = (double)xyz[3 * i + 1], z = (double)xyz[3 * i + 2];
        for (int j = 0; j < 3; ++j) {
            /* homo (float32, widened) times column j of RT.T == row j of RT */
            const double *r = rt + 4 * j;
            double s = x * r[0];
            s = fma(y, r[1], s);
            s = fma(z, r[2], s);
            s = fma(1.0, r[3], s);
            double f = floor(s);
            g[3 * i + j] = f;
            if (f < mn[j]) mn[j] = f;
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        uint64_t w[3];
        for (int j = 0; j < 3; ++j) {
            g[3 * i + j] = floor(g[3 * i + j] - mn[j]);
            w[j] = (uint64_t)g[3 * i + j];
        }
        key[i] = fnv1_words(w, 3);
    }
    int rc = xo_unique_u64(key, n, first, inverse, NULL, m_out);
    if (rc == 0 && voxel_xyz)
        for (int64_t r = 0; r < *m_out; ++r)
            for (int j = 0; j < 3; ++j) voxel_xyz[3 * r + j] = g[3 * first[r] + j];
    free(g); free(key);
    return rc;
}

/* ---------------------------------------------------------------- stage 3: pooling
 * Truth for models/utils/criterion.py:152-157 (feat[mask_k].mean(0)) in the partition
 * (label-per-point) form: fp64 sums and counts per label; label < 0 or >= K = unpooled. */
XO_API void xo_pool_label_f64(const float *feat, int64_t n, int c, const int32_t *label, int k,
                              double *sum /*[k,c]*/, int64_t *cnt /*[k]*/) {
    memset(sum, 0, sizeof(double) * (size_t)k * (size_t)c);
    memset(cnt, 0, sizeof(int64_t) * (size_t)k);
    for (int64_t i = 0; i < n; ++i) {
        int32_t l = label[i];
        if (l < 0 || l >= k) continue;
        cnt[l] += 1;
        const float *f = feat + i * (int64_t)c;
        double *s = sum + (int64_t)l * c;
        for (int j = 0; j < c; ++j) s[j] += (double)f[j];
    }
}

/* General (overlapping) membership: member is a [k,n] byte matrix. */
XO_API void xo_pool_member_f64(const float *feat, int64_t n, int c, const uint8_t *member, int k,
                               double *sum, int64_t *cnt) {
    memset(sum, 0, sizeof(double) * (size_t)k * (size_t)c);
    memset(cnt, 0, sizeof(int64_t) * (size_t)k);
    for (int m = 0; m < k; ++m)
        for (int64_t i = 0; i < n; ++i)
            if (member[(int64_t)m * n + i]) {
                cnt[m] += 1;
                const float *f = feat + i * (int64_t)c;
                double *s = sum + (int64_t)m * c;
                for (int j = 0; j < c; ++j) s[j] += (double)f[j];
            }
}

/* models/utils/fuser.py:22-34 — mask -> point scatter-mean, fp32, masks in ascending order,
 * then `/ counter` with counter 0 -> 1e-5 (float32).  Bit-exact restatement:
 * out[i,:] = (((0 + e_k1) + e_k2) + ...) / float(count_i).  The "no mask hits any point ->
 * member[0][0] = True" guard (:19-20) is the caller's job. */
XO_API void xo_scatter_member_f32(const uint8_t *member, int k, int64_t n, const float *emb, int c,
                                  float *out /*[n,c]*/, float *counter /*[n]*/) {
    for (int64_t i = 0; i < n; ++i) {
        float *o = out + i * (int64_t)c;
        for (int j = 0; j < c; ++j) o[j] = 0.0f;
        float cnt = 0.0f;
        for (int m = 0; m < k; ++m)
            if (member[(int64_t)m * n + i]) {
                const float *e = emb + (int64_t)m * c;
                for (int j = 0; j < c; ++j) o[j] = o[j] + e[j];
                cnt = cnt + 1.0f;
            }
        if (cnt == 0.0f) cnt = 1e-5f;
        for (int j = 0; j < c; ++j) o[j] = o[j] / cnt;
        counter[i] = cnt;
    }
}
