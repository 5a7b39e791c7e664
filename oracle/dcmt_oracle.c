/*
 * dcmt_oracle.c -- TEST INFRASTRUCTURE ONLY.  Plain-C CPU restatement of the reference hot path.
 *
 * What it is: a scalar, single-threaded C restatement of
 *   /root/reference/src/DC_lidar_only/img_completion.cpp:17-204        (img_completion)
 *   /root/reference/src/DC_lidar_camera/img_completion_lc.cpp:34-203   (interpolate_with_superpixels)
 *   /root/reference/src/DC_stereo_lidar/main_sl.cpp:715-885,1253       (stereo refinement)
 * including the arithmetic of the OpenCV `imgproc` calls those functions make.  OpenCV is a
 * third-party dependency of the reference that is neither vendored nor version-pinned there;
 * the semantics restated here are the published ones of OpenCV 4.13.0 (opencv-python-headless
 * 4.13.0.92) and are checked, function by function, against that very library through
 * oracle/cv2_oracle.py (tests/test_oracle.py) and against the committed vectors in tests/golden/.
 *
 * Parity status: the reference has no tests / golden vectors.  The pin is the reference's own
 * sources compiled against a stand-in OpenCV / Eigen header with the imgproc calls forwarded to
 * cv2 (oracle/refshim, oracle/_ref/libdcmt_ref.so): tests/test_reference_build.py requires
 * "reference build == cv2 transliteration == this file == tests/golden", bit for bit, for
 * img_completion, interpolate_with_superpixels, the stereo chain, SLIC and the evaluation loops.
 *
 * Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs -- as the checker, never as the product.  The product (libdcmt.so) does not link it.
 *
 * Build: oracle/Makefile  ->  oracle/_build/libdcmt_oracle.so
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DCMT_ORACLE_API __attribute__((visibility("default")))

/* img_completion.cpp:59 et al. compare float against the double literal 0.1.
 * (double)x > 0.1  <=>  x >= 0.1f ;  (double)x < 0.1  <=>  x < 0.1f.  Kept literal here. */
static inline int is_valid(float d) { return (double)d > 0.1; }
static inline int is_hole(float d) { return (double)d < 0.1; }

enum { BLUR_NONE = 0, BLUR_GAUSSIAN = 1, BLUR_BILATERAL = 2 };

/* number of per-stage snapshots written by the *_stages entry points */
#define DCMT_ORACLE_N_STAGES 10

/* ---------------------------------------------------------------- elementary operators */

/* img_completion.cpp:55-67 and :191-202 */
static void invert_valid(float *d, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (is_valid(d[i])) d[i] = 100.0f - d[i];
}

/* img_completion.cpp:71-80.  cv::dilate with the int[5][5] "diamond" read as 25 bytes: taps
 * at kernel (row 1,col 3) and (row 4,col 4), anchor (2,2) -> offsets (-1,+1) and (+2,+2).
 * BORDER_CONSTANT with morphologyDefaultBorderValue: absent taps read -FLT_MAX. */
static void dilate_two_tap(const float *src, float *dst, int rows, int cols) {
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float t1 = (y - 1 >= 0 && x + 1 < cols) ? src[(size_t)(y - 1) * cols + x + 1] : -FLT_MAX;
            float t2 = (y + 2 < rows && x + 2 < cols) ? src[(size_t)(y + 2) * cols + x + 2] : -FLT_MAX;
            dst[(size_t)y * cols + x] = t1 > t2 ? t1 : t2;
        }
}

/* cv::dilate / cv::erode with a k x k all-ones kernel, anchor at centre: separable row/column
 * extremum; out-of-image samples are absent (border value = identity of the operator). */
static void box_extremum(const float *src, float *dst, float *tmp, int rows, int cols, int k, int is_max) {
    const int r = k / 2;
    for (int y = 0; y < rows; ++y) {
        const float *s = src + (size_t)y * cols;
        float *t = tmp + (size_t)y * cols;
        for (int x = 0; x < cols; ++x) {
            int lo = x - r < 0 ? 0 : x - r, hi = x + r >= cols ? cols - 1 : x + r;
            float m = s[lo];
            if (is_max) { for (int i = lo + 1; i <= hi; ++i) if (s[i] > m) m = s[i]; }
            else        { for (int i = lo + 1; i <= hi; ++i) if (s[i] < m) m = s[i]; }
            t[x] = m;
        }
    }
    for (int y = 0; y < rows; ++y) {
        int lo = y - r < 0 ? 0 : y - r, hi = y + r >= rows ? rows - 1 : y + r;
        float *o = dst + (size_t)y * cols;
        memcpy(o, tmp + (size_t)lo * cols, sizeof(float) * cols);
        for (int i = lo + 1; i <= hi; ++i) {
            const float *t = tmp + (size_t)i * cols;
            if (is_max) { for (int x = 0; x < cols; ++x) if (t[x] > o[x]) o[x] = t[x]; }
            else        { for (int x = 0; x < cols; ++x) if (t[x] < o[x]) o[x] = t[x]; }
        }
    }
}

/* img_completion.cpp:88-100, :131-144: d = hole(d) ? dilate_k(d) : d.  Returns the number of
 * holes counted BEFORE the fill (img_completion.cpp:149-158). */
static long fill_holes_with_dilate(float *d, float *scratch, float *tmp, int rows, int cols, int k) {
    size_t n = (size_t)rows * cols;
    long holes = 0;
    box_extremum(d, scratch, tmp, rows, cols, k, 1);
    for (size_t i = 0; i < n; ++i)
        if (is_hole(d[i])) { d[i] = scratch[i]; ++holes; }
    return holes;
}

/* img_completion.cpp:103-129 (always on; `extr` is ignored by the reference) */
static void column_extrapolation(float *d, int rows, int cols) {
    for (int j = 0; j < cols; ++j) {
        int max_index = 0, min_index = rows - 1;
        float max_val = -1.0f, min_val = 100.0f;
        for (int i = 0; i < rows; ++i) {
            if (is_valid(d[(size_t)i * cols + j])) { max_index = i; max_val = d[(size_t)i * cols + j]; }
            if (is_valid(d[(size_t)(rows - 1 - i) * cols + j])) {
                min_index = rows - 1 - i; min_val = d[(size_t)(rows - 1 - i) * cols + j];
            }
        }
        for (int i = max_index; i < rows; ++i) d[(size_t)i * cols + j] = max_val;
        for (int i = min_index; i >= 0; --i) d[(size_t)i * cols + j] = min_val;
    }
}

/* cv::medianBlur(ksize 5) on CV_32F: exact rank-12-of-25, BORDER_REPLICATE. */
static void median5(const float *src, float *dst, int rows, int cols) {
    float w[25];
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int n = 0;
            for (int dy = -2; dy <= 2; ++dy) {
                int yy = y + dy; yy = yy < 0 ? 0 : (yy >= rows ? rows - 1 : yy);
                for (int dx = -2; dx <= 2; ++dx) {
                    int xx = x + dx; xx = xx < 0 ? 0 : (xx >= cols ? cols - 1 : xx);
                    w[n++] = src[(size_t)yy * cols + xx];
                }
            }
            /* partial selection sort up to rank 12 */
            for (int i = 0; i <= 12; ++i) {
                int m = i;
                for (int j = i + 1; j < 25; ++j) if (w[j] < w[m]) m = j;
                float t = w[i]; w[i] = w[m]; w[m] = t;
            }
            dst[(size_t)y * cols + x] = w[12];
        }
}

/* cv::borderInterpolate(p, len, BORDER_REFLECT_101) */
static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

/* cv::GaussianBlur(ksize 5x5, sigma 0): getGaussianKernel(5, 0) returns the fixed table
 * [1,4,6,4,1]/16; separable row then column pass in float32, BORDER_REFLECT_101.  The
 * summation order below is the symmetric form; for q8 input every partial sum is exact so the
 * order is immaterial, otherwise results agree with OpenCV to <= 3e-5 abs at |x| <= 100. */
static void gaussian5(const float *src, float *dst, float *tmp, int rows, int cols) {
    const float k0 = 0.375f, k1 = 0.25f, k2 = 0.0625f;
    for (int y = 0; y < rows; ++y) {
        const float *s = src + (size_t)y * cols;
        float *t = tmp + (size_t)y * cols;
        for (int x = 0; x < cols; ++x) {
            float a1 = s[reflect101(x - 1, cols)], b1 = s[reflect101(x + 1, cols)];
            float a2 = s[reflect101(x - 2, cols)], b2 = s[reflect101(x + 2, cols)];
            t[x] = s[x] * k0 + (a1 + b1) * k1 + (a2 + b2) * k2;
        }
    }
    for (int y = 0; y < rows; ++y) {
        const float *c0 = tmp + (size_t)y * cols;
        const float *m1 = tmp + (size_t)reflect101(y - 1, rows) * cols;
        const float *p1 = tmp + (size_t)reflect101(y + 1, rows) * cols;
        const float *m2 = tmp + (size_t)reflect101(y - 2, rows) * cols;
        const float *p2 = tmp + (size_t)reflect101(y + 2, rows) * cols;
        float *o = dst + (size_t)y * cols;
        for (int x = 0; x < cols; ++x) o[x] = c0[x] * k0 + (m1[x] + p1[x]) * k1 + (m2[x] + p2[x]) * k2;
    }
}

/* cv::bilateralFilter(src, dst, d=5, sigmaColor=1.5, sigmaSpace=2.0) for CV_32FC1, restated
 * from OpenCV 4.x bilateral_filter.dispatch.cpp / .simd.hpp (bilateralFilter_32f):
 * radius 2, circular support r<=2 (13 taps incl. centre), BORDER_REFLECT_101, colour weight
 * from a 4096-bin linearly interpolated exp LUT over [0, max-min]. */
static void bilateral5(const float *src, float *dst, int rows, int cols) {
    const double sigma_color = 1.5, sigma_space = 2.0;
    const int radius = 2;
    const double gauss_color_coeff = -0.5 / (sigma_color * sigma_color);
    const double gauss_space_coeff = -0.5 / (sigma_space * sigma_space);
    size_t n = (size_t)rows * cols;
    float mn = src[0], mx = src[0];
    for (size_t i = 1; i < n; ++i) { if (src[i] < mn) mn = src[i]; if (src[i] > mx) mx = src[i]; }
    if (fabs((double)mn - (double)mx) < FLT_EPSILON) { memcpy(dst, src, n * sizeof(float)); return; }
    const int bins = 1 << 12;
    float len = (float)((double)mx - (double)mn);
    float scale_index = (float)bins / len;
    float *lut = (float *)malloc(sizeof(float) * (bins + 2));
    float last = 1.f;
    for (int i = 0; i < bins + 2; ++i) {
        if (last > 0.f) {
            double v = i / scale_index;
            lut[i] = (float)exp(v * v * gauss_color_coeff);
            last = lut[i];
        } else lut[i] = 0.f;
    }
    float sw[25]; int oy[25], ox[25]; int maxk = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > radius || (i == 0 && j == 0)) continue;
            sw[maxk] = (float)exp(r * r * gauss_space_coeff);
            oy[maxk] = i; ox[maxk] = j; ++maxk;
        }
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float val0 = src[(size_t)y * cols + x];
            float sum = val0, wsum = 1.f; /* centre tap: weight 1 */
            for (int k = 0; k < maxk; ++k) {
                int yy = reflect101(y + oy[k], rows), xx = reflect101(x + ox[k], cols);
                float val = src[(size_t)yy * cols + xx];
                float alpha = fabsf(val - val0) * scale_index;
                int idx = (int)floorf(alpha);
                alpha -= (float)idx;
                float w = sw[k] * (lut[idx] + alpha * (lut[idx + 1] - lut[idx]));
                sum += val * w;
                wsum += w;
            }
            dst[(size_t)y * cols + x] = sum / wsum;
        }
    free(lut);
}

/* ---------------------------------------------------------------- shared tail (A4..A10) */

static void snapshot(float *stages, int idx, const float *d, size_t n) {
    if (stages) memcpy(stages + (size_t)idx * n, d, n * sizeof(float));
}

/* img_completion.cpp:88-202 == img_completion_lc.cpp:105-202.  d is modified in place.
 * stats[0] = passes of the while loop (:146-166), stats[1] = holes left after the first 31x31
 * fill (:131-144), stats[2] = holes present after column extrapolation. */
static int completion_tail(float *d, int rows, int cols, int blur_type, int32_t *stats, float *stages) {
    size_t n = (size_t)rows * cols;
    float *scratch = (float *)malloc(n * sizeof(float));
    float *tmp = (float *)malloc(n * sizeof(float));
    if (!scratch || !tmp) { free(scratch); free(tmp); return -1; }

    fill_holes_with_dilate(d, scratch, tmp, rows, cols, 7);            /* :88-100  */
    snapshot(stages, 3, d, n);
    column_extrapolation(d, rows, cols);                               /* :103-129 */
    snapshot(stages, 4, d, n);
    long h0 = fill_holes_with_dilate(d, scratch, tmp, rows, cols, 31); /* :131-144 */
    snapshot(stages, 5, d, n);
    int passes = 0; long first_count = -1;
    for (;;) {                                                         /* :146-166 */
        long c = fill_holes_with_dilate(d, scratch, tmp, rows, cols, 31);
        if (passes == 0) first_count = c;
        ++passes;
        if (c == 0) break;
        if (passes > 100000) { free(scratch); free(tmp); return -2; }
    }
    snapshot(stages, 6, d, n);
    if (stats) { stats[0] = passes; stats[1] = (int32_t)first_count; stats[2] = (int32_t)h0; }

    median5(d, scratch, rows, cols);                                   /* :170 */
    memcpy(d, scratch, n * sizeof(float));
    snapshot(stages, 7, d, n);
    if (blur_type == BLUR_GAUSSIAN) {                                  /* :176-189 */
        gaussian5(d, scratch, tmp, rows, cols);
        for (size_t i = 0; i < n; ++i) if (is_valid(d[i])) d[i] = scratch[i];
    } else if (blur_type == BLUR_BILATERAL) {                          /* :172-175, intent */
        bilateral5(d, scratch, rows, cols);
        memcpy(d, scratch, n * sizeof(float));
    }
    snapshot(stages, 8, d, n);
    invert_valid(d, n);                                                /* :191-202 */
    snapshot(stages, 9, d, n);
    free(scratch); free(tmp);
    return 0;
}

/* 2-tap dilate + 5x5 close on a full frame (img_completion.cpp:71-85) */
static int front_morphology(float *d, int rows, int cols, float *stages) {
    size_t n = (size_t)rows * cols;
    float *a = (float *)malloc(n * sizeof(float));
    float *tmp = (float *)malloc(n * sizeof(float));
    if (!a || !tmp) { free(a); free(tmp); return -1; }
    dilate_two_tap(d, a, rows, cols);
    snapshot(stages, 1, a, n);
    box_extremum(a, d, tmp, rows, cols, 5, 1); /* MORPH_CLOSE = dilate ... */
    box_extremum(d, a, tmp, rows, cols, 5, 0); /* ... then erode           */
    memcpy(d, a, n * sizeof(float));
    snapshot(stages, 2, d, n);
    free(a); free(tmp);
    return 0;
}

/* ---------------------------------------------------------------- public: lidar only */

/* stages (optional): DCMT_ORACLE_N_STAGES x rows x cols floats, snapshots after
 * 0 invert, 1 two-tap, 2 close5, 3 dilate7 fill, 4 column extrapolation, 5 first 31 fill,
 * 6 fill loop, 7 median, 8 blur, 9 final invert. */
DCMT_ORACLE_API int dcmt_oracle_img_completion_stages(const float *sparse, float *dense, int rows, int cols,
                                                      int blur_type, int32_t *stats, float *stages) {
    if (!sparse || !dense || rows < 1 || cols < 1) return -1;
    size_t n = (size_t)rows * cols;
    memcpy(dense, sparse, n * sizeof(float));          /* :27 clone */
    invert_valid(dense, n);                            /* :55-67 */
    snapshot(stages, 0, dense, n);
    if (front_morphology(dense, rows, cols, stages)) return -1;
    return completion_tail(dense, rows, cols, blur_type, stats, stages);
}

DCMT_ORACLE_API int dcmt_oracle_img_completion(const float *sparse, float *dense, int rows, int cols,
                                               int blur_type, int32_t *stats) {
    return dcmt_oracle_img_completion_stages(sparse, dense, rows, cols, blur_type, stats, NULL);
}

/* ---------------------------------------------------------------- public: superpixel guided */

/* img_completion_lc.cpp:34-203.  labels: row-major [row][col] int32 (transposed view of
 * slic.clusters[col][row], :83), values outside [0, n_clusters) are never selected.
 * literal != 0 runs the reference's per-cluster loop (:78-103) verbatim; literal == 0 uses
 * the order-independent per-pixel closed form of SURVEY.md Appendix B.  Both are bit-equal
 * (tests/test_oracle.py). */
DCMT_ORACLE_API int dcmt_oracle_interpolate_with_superpixels(const float *sparse, const int32_t *labels,
                                                             int n_clusters, float *dense, int rows, int cols,
                                                             int blur_type_ignored, int use_superpixel,
                                                             int literal, int32_t *stats, float *stages) {
    (void)blur_type_ignored; /* :183-192 blur is unconditional gaussian */
    if (!sparse || !dense || rows < 1 || cols < 1) return -1;
    if (use_superpixel && !labels) return -1;
    size_t n = (size_t)rows * cols;
    memcpy(dense, sparse, n * sizeof(float));
    invert_valid(dense, n);                            /* :45-52 */
    snapshot(stages, 0, dense, n);
    if (!use_superpixel) {
        if (front_morphology(dense, rows, cols, stages)) return -1;   /* :59-64 */
    } else if (literal) {
        float *region = (float *)malloc(n * sizeof(float));
        float *a = (float *)malloc(n * sizeof(float));
        float *tmp = (float *)malloc(n * sizeof(float));
        if (!region || !a || !tmp) { free(region); free(a); free(tmp); return -1; }
        for (int c = 0; c < n_clusters; ++c) {                         /* :78-103 */
            size_t cnt = 0;
            for (size_t i = 0; i < n; ++i) {
                int in = labels[i] == c;
                region[i] = in ? dense[i] : 0.0f;  /* copyTo(fresh Mat, mask) zero-fills */
                cnt += in;
            }
            if (!cnt) continue;
            dilate_two_tap(region, a, rows, cols);
            box_extremum(a, region, tmp, rows, cols, 5, 1);
            box_extremum(region, a, tmp, rows, cols, 5, 0);
            for (size_t i = 0; i < n; ++i) if (labels[i] == c) dense[i] = a[i];
        }
        free(region); free(a); free(tmp);
        snapshot(stages, 2, dense, n);
    } else {
        float *out = (float *)malloc(n * sizeof(float));
        if (!out) return -1;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                int c = labels[(size_t)y * cols + x];
                if (c < 0 || c >= n_clusters) { out[(size_t)y * cols + x] = dense[(size_t)y * cols + x]; continue; }
                float er = FLT_MAX;
                for (int qy = y - 2; qy <= y + 2; ++qy) {
                    if (qy < 0 || qy >= rows) continue;
                    for (int qx = x - 2; qx <= x + 2; ++qx) {
                        if (qx < 0 || qx >= cols) continue;
                        float dl = -FLT_MAX;
                        for (int ry = qy - 2; ry <= qy + 2; ++ry) {
                            if (ry < 0 || ry >= rows) continue;
                            for (int rx = qx - 2; rx <= qx + 2; ++rx) {
                                if (rx < 0 || rx >= cols) continue;
                                float t1 = -FLT_MAX, t2 = -FLT_MAX;
                                if (ry - 1 >= 0 && rx + 1 < cols) {
                                    size_t s = (size_t)(ry - 1) * cols + rx + 1;
                                    t1 = labels[s] == c ? dense[s] : 0.0f;
                                }
                                if (ry + 2 < rows && rx + 2 < cols) {
                                    size_t s = (size_t)(ry + 2) * cols + rx + 2;
                                    t2 = labels[s] == c ? dense[s] : 0.0f;
                                }
                                float r1 = t1 > t2 ? t1 : t2;
                                if (r1 > dl) dl = r1;
                            }
                        }
                        if (dl < er) er = dl;
                    }
                }
                out[(size_t)y * cols + x] = er;
            }
        memcpy(dense, out, n * sizeof(float));
        free(out);
        snapshot(stages, 2, dense, n);
    }
    return completion_tail(dense, rows, cols, BLUR_GAUSSIAN, stats, stages);
}

/* ---------------------------------------------------------------- public: stereo refinement */

typedef struct {
    float baseline;      /* 0.54       main_sl.cpp:848,866 */
    float focal;         /* 959.791    main_sl.cpp:849,867 */
    float damp_factor;   /* 500 (OFFICIAL 1370)  :808 */
    float err_clip;      /* 255 (OFFICIAL 221)   :820-825 */
    float depth_clip;    /* 100 (OFFICIAL 80)    :875 */
    int32_t num_iterations; /* 4        :805 */
    int32_t final_gauss;    /* 1        :1253 */
} dcmt_oracle_stereo_params;

/* main_sl.cpp:715-745 for one plane: value -> dx (dy is computed by the reference but never
 * read by optimize_IG because dr == 0; it is produced here for the a4 parity test). */
DCMT_ORACLE_API void dcmt_oracle_measurement_derivatives(const float *val, float *dx, float *dy, int rows, int cols) {
    size_t n = (size_t)rows * cols;
    memset(dx, 0, n * sizeof(float));
    if (dy) memset(dy, 0, n * sizeof(float));
    for (int r = 1; r < rows - 1; ++r)
        for (int c = 1; c < cols - 1; ++c) {
            size_t i = (size_t)r * cols + c;
            dx[i] = (float)(.5 * val[i + 1] - .5 * val[i - 1]);
            if (dy) dy[i] = (float)(.5 * val[i + cols] - .5 * val[i - cols]);
        }
}

/* entry read with the reference's undefined out-of-range reads DEFINED as 0 */
static inline float at0(const float *p, int r, int c, int rows, int cols) {
    return (r >= 0 && r < rows && c >= 0 && c < cols) ? p[(size_t)r * cols + c] : 0.0f;
}

/* main_sl.cpp:747-801 */
static int observation_derivatives(const float *val, const float *dx, int rows, int cols, float r, float c,
                                   float *value, float *deriv_x) {
    int r0 = (int)(r + 0.5);
    double cd = (double)c + 0.5;
    int c0 = (cd >= 2147483647.0 || cd <= -2147483648.0 || cd != cd) ? INT32_MIN : (int)cd; /* x86 cvttsd2si */
    if (r0 < 0 || r0 > rows || c0 < 0 || c0 > cols) return 0;
    int r1 = r0 + 1, c1 = c0 + 1;
    if (r1 < 0 || r1 > rows || c1 < 0 || c1 > cols) return 0;
    float p00 = at0(val, r0, c0, rows, cols), p01 = at0(val, r0, c1, rows, cols);
    float p10 = at0(val, r1, c0, rows, cols), p11 = at0(val, r1, c1, rows, cols);
    float g00 = at0(dx, r0, c0, rows, cols), g01 = at0(dx, r0, c1, rows, cols);
    float g10 = at0(dx, r1, c0, rows, cols), g11 = at0(dx, r1, c1, rows, cols);
    const float dr = r - (float)r0;
    const float dc = c - (float)c0;
    const float dr1 = (float)(1. - dr);
    const float dc1 = (float)(1. - dc);
    *value = (p00 * dc1 + p01 * dc) * dr1 + (p10 * dc1 + p11 * dc) * dr;
    *deriv_x = (g00 * dc1 + g01 * dc) * dr1 + (g10 * dc1 + g11 * dc) * dr;
    return 1;
}

/* main_sl.cpp:804-843 */
DCMT_ORACLE_API void dcmt_oracle_optimize_IG(const float *val_l, const float *val_r, const float *dx_r, float *disp,
                                             int rows, int cols, int num_iterations, float damp_factor,
                                             float err_clip) {
    for (int k = 0; k < num_iterations; ++k)
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) {
                size_t p = (size_t)i * cols + j;
                float pixel_right = (float)j - disp[p];
                float value = 0.f, dx = 0.f;
                int ok = observation_derivatives(val_r, dx_r, rows, cols, (float)i, pixel_right, &value, &dx);
                if (ok && disp[p] != 0) {
                    float error = value - val_l[p];
                    if (error > err_clip) error = err_clip;
                    if (error < -err_clip) error = -err_clip;
                    float J = -1;
                    float J_cr = J * dx;
                    float H = J_cr * J_cr + damp_factor;
                    float b = J_cr * error;
                    float dd = -b / H;
                    disp[p] += dd;
                }
            }
}

/* main_sl.cpp:1165-1253 end to end: gray u8 planes + initial depth -> refined depth. */
DCMT_ORACLE_API int dcmt_oracle_stereo_refine(const float *depth_ig, const uint8_t *left_gray,
                                              const uint8_t *right_gray, float *depth_out, float *disp_out,
                                              int rows, int cols, const dcmt_oracle_stereo_params *prm) {
    if (!depth_ig || !left_gray || !right_gray || !depth_out || !prm || rows < 1 || cols < 1) return -1;
    size_t n = (size_t)rows * cols;
    float *vl = (float *)malloc(n * sizeof(float)), *vr = (float *)malloc(n * sizeof(float));
    float *dxr = (float *)malloc(n * sizeof(float)), *disp = (float *)calloc(n, sizeof(float));
    float *tmp = (float *)malloc(n * sizeof(float)), *dep = (float *)calloc(n, sizeof(float));
    if (!vl || !vr || !dxr || !disp || !tmp || !dep) return -1;
    for (size_t i = 0; i < n; ++i) { vl[i] = (float)left_gray[i]; vr[i] = (float)right_gray[i]; } /* :1173-1189 */
    dcmt_oracle_measurement_derivatives(vr, dxr, NULL, rows, cols);                              /* :1193 */
    const float bf = prm->baseline * prm->focal;
    for (size_t i = 0; i < n; ++i) if (depth_ig[i] > 0) disp[i] = bf / depth_ig[i];              /* :846-861 */
    dcmt_oracle_optimize_IG(vl, vr, dxr, disp, rows, cols, prm->num_iterations, prm->damp_factor, prm->err_clip);
    for (size_t i = 0; i < n; ++i)                                                               /* :863-885 */
        if (disp[i] > 0) { float d = bf / disp[i]; if (d > prm->depth_clip) d = prm->depth_clip; dep[i] = d; }
    if (prm->final_gauss) gaussian5(dep, depth_out, tmp, rows, cols);                            /* :1253 */
    else memcpy(depth_out, dep, n * sizeof(float));
    if (disp_out) memcpy(disp_out, disp, n * sizeof(float));
    free(vl); free(vr); free(dxr); free(disp); free(tmp); free(dep);
    return 0;
}

/* ---------------------------------------------------------------- single-operator probes (tests) */
DCMT_ORACLE_API void dcmt_oracle_op_two_tap(const float *s, float *d, int rows, int cols) { dilate_two_tap(s, d, rows, cols); }
DCMT_ORACLE_API void dcmt_oracle_op_box(const float *s, float *d, int rows, int cols, int k, int is_max) {
    float *tmp = (float *)malloc((size_t)rows * cols * sizeof(float));
    box_extremum(s, d, tmp, rows, cols, k, is_max);
    free(tmp);
}
DCMT_ORACLE_API void dcmt_oracle_op_median5(const float *s, float *d, int rows, int cols) { median5(s, d, rows, cols); }
DCMT_ORACLE_API void dcmt_oracle_op_gaussian5(const float *s, float *d, int rows, int cols) {
    float *tmp = (float *)malloc((size_t)rows * cols * sizeof(float));
    gaussian5(s, d, tmp, rows, cols);
    free(tmp);
}
DCMT_ORACLE_API void dcmt_oracle_op_bilateral5(const float *s, float *d, int rows, int cols) { bilateral5(s, d, rows, cols); }
DCMT_ORACLE_API void dcmt_oracle_op_column_extrapolation(float *d, int rows, int cols) { column_extrapolation(d, rows, cols); }
DCMT_ORACLE_API int dcmt_oracle_n_stages(void) { return DCMT_ORACLE_N_STAGES; }

/* ---------------------------------------------------------------- evaluation
 * Literal restatement of the three evaluation loops: float32 accumulators, raster order.
 *   mode 0  src/DC_lidar_only/main.cpp:16-34        mask gt > tol             out[0] = sum(gt - r) / count
 *   mode 1  src/DC_lidar_camera/main_lc.cpp:85-116  mask gt > tol && r > tol  out[1] = sum|d| / count, out[2] = sqrt(sum d^2 / count)
 *           src/DC_stereo_lidar/main_sl.cpp:1031-1061 (tolerance 2)
 * `tolerance` is the reference's int (0, (int)0.1 = 0, 2).  out[3] = count. */
DCMT_ORACLE_API void dcmt_oracle_evaluate(const float *gt, const float *r, int rows, int cols, int tolerance, int mode, float *out) {
    float sum_err = 0, sum_mse = 0, sum_mae = 0;
    int count = 0;
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            const float g = gt[(size_t)i * cols + j], v = r[(size_t)i * cols + j];
            if (mode == 0) {
                if (g > tolerance) { sum_err += (g - v); count++; }
            } else if (g > tolerance && v > tolerance) {
                const float d = fabsf(g - v);
                sum_mse += d * d;
                sum_mae += d;
                count++;
            }
        }
    out[0] = sum_err / count;
    out[1] = sum_mae / count;
    out[2] = sqrtf(sum_mse / count);
    out[3] = (float)count;
}

/* ---------------------------------------------------------------- LiDAR projection
 * Literal restatement of src/DC_stereo_lidar/main_sl.cpp:478-523: transform (:485-487, term by term), z > 0 filter
 * (:488), Eigen P * p.homogeneous() (:500; evaluated like Eigen >= 3.3: (a0 + (a1 + a2)) + P.col(3), see refshim/Eigen/Dense), division (:501-502), bounds test on
 * the float coordinates (:505-506), (int) truncation (:510-511), last writer wins (:515); then cv::normalize
 * (NORM_MINMAX, a, b; OpenCV 4.x norm.cpp: float scale, float shift, dst = src * scale + shift).  T: 4x4 row-major,
 * P: 3x4 row-major.  Returns the number of projected points (:508). */
static float row_dot4(const float *m, float x, float y, float z) { return m[0] * x + m[1] * y + m[2] * z + m[3]; }
static float row_dot4_eigen(const float *m, float x, float y, float z) {
    const float a0 = m[0] * x, a1 = m[1] * y, a2 = m[2] * z;
    float d = a0 + (a1 + a2);
    d += m[3];
    return d;
}
DCMT_ORACLE_API int dcmt_oracle_lidar_project(const float *pts, int n, const float *T, const float *P, int rows, int cols,
                                              float *projected, float *normalized, float a, float b) {
    int count = 0;
    for (size_t i = 0; i < (size_t)rows * cols; ++i) projected[i] = 0.0f;
    for (int i = 0; i < n; ++i) {
        const float x = pts[4 * i], y = pts[4 * i + 1], z = pts[4 * i + 2];
        const float tx = row_dot4(T, x, y, z), ty = row_dot4(T + 4, x, y, z), tz = row_dot4(T + 8, x, y, z);
        if (!(tz > 0)) continue;
        const float X = row_dot4_eigen(P, tx, ty, tz), Y = row_dot4_eigen(P + 4, tx, ty, tz), Z = row_dot4_eigen(P + 8, tx, ty, tz);
        const float u = X / Z, v = Y / Z;
        if (u >= 0 && u < cols && v >= 0 && v < rows) {
            count++;
            projected[(size_t)(int)v * cols + (int)u] = Z;
        }
    }
    if (normalized) {
        double smin = projected[0], smax = projected[0];
        for (size_t i = 1; i < (size_t)rows * cols; ++i) {
            if (projected[i] < smin) smin = projected[i];
            if (projected[i] > smax) smax = projected[i];
        }
        const double dmin = a < b ? a : b, dmax = a < b ? b : a;
        double scale = (dmax - dmin) * (smax - smin > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
        scale = (double)(float)scale;
        const float shift = (float)dmin - (float)(smin * scale);
        const float fs = (float)scale;
        for (size_t i = 0; i < (size_t)rows * cols; ++i) normalized[i] = projected[i] * fs + shift;
    }
    return count;
}

/* ---------------------------------------------------------------- SLIC
 * Literal restatement of Slic::generate_superpixels (src/DC_lidar_camera/slic.cpp:101-182) with init_data (:19-59),
 * find_local_minimum (:72-99) and compute_dist (:61-69): same loops, same double arithmetic, clusters / distances
 * indexed [col][row].  lab: rows x cols x 3 bytes.  labels_out is written row-major [row][col]; centers_out K x 5.
 * Returns K = centers.size(). */
DCMT_ORACLE_API int dcmt_oracle_slic(const uint8_t *lab, int rows, int cols, int step, int nc, int iterations, int32_t *labels_out,
                                     double *centers_out) {
    const int ns = step;
    int K = 0;
    for (int i = step; i < cols - step / 2; i += step)
        for (int j = step; j < rows - step / 2; j += step) ++K;
    int *clusters = (int *)malloc((size_t)rows * cols * sizeof(int));        /* [col][row] */
    double *distances = (double *)malloc((size_t)rows * cols * sizeof(double));
    double *centers = (double *)malloc((size_t)(K ? K : 1) * 5 * sizeof(double));
    int *counts = (int *)malloc((size_t)(K ? K : 1) * sizeof(int));
    for (size_t i = 0; i < (size_t)rows * cols; ++i) { clusters[i] = -1; distances[i] = FLT_MAX; }
#define LAB(y, x, c) lab[((size_t)(y) * cols + (x)) * 3 + (c)]
    int n = 0;
    for (int i = step; i < cols - step / 2; i += step)
        for (int j = step; j < rows - step / 2; j += step) {
            double min_grad = FLT_MAX;
            int lx = i, ly = j;
            for (int a = i - 1; a < i + 2; a++)
                for (int b = j - 1; b < j + 2; b++) {
                    const double i1 = LAB(b + 1, a, 0), i2 = LAB(b, a + 1, 0), i3 = LAB(b, a, 0);
                    if (sqrt(pow(i1 - i3, 2)) + sqrt(pow(i2 - i3, 2)) < min_grad) {
                        min_grad = fabs(i1 - i3) + fabs(i2 - i3);
                        lx = a;
                        ly = b;
                    }
                }
            double *c = centers + (size_t)n * 5;
            c[0] = LAB(ly, lx, 0); c[1] = LAB(ly, lx, 1); c[2] = LAB(ly, lx, 2); c[3] = lx; c[4] = ly;
            counts[n++] = 0;
        }
    for (int it = 0; it < iterations; it++) {
        for (size_t i = 0; i < (size_t)rows * cols; ++i) distances[i] = FLT_MAX;
        for (int j = 0; j < K; j++) {
            const double *c = centers + (size_t)j * 5;
            if (c[3] != c[3] || c[4] != c[4]) continue; /* NaN centre: `int k = NaN` is undefined; x86 gives INT_MIN and the loop body never runs */
            for (int k = c[3] - step; k < c[3] + step; k++)
                for (int l = c[4] - step; l < c[4] + step; l++)
                    if (k >= 0 && k < cols && l >= 0 && l < rows) {
                        const double dc = sqrt(pow(c[0] - LAB(l, k, 0), 2) + pow(c[1] - LAB(l, k, 1), 2) + pow(c[2] - LAB(l, k, 2), 2));
                        const double ds = sqrt(pow(c[3] - k, 2) + pow(c[4] - l, 2));
                        const double d = sqrt(pow(dc / nc, 2) + pow(ds / ns, 2));
                        if (d < distances[(size_t)k * rows + l]) {
                            distances[(size_t)k * rows + l] = d;
                            clusters[(size_t)k * rows + l] = j;
                        }
                    }
        }
        for (int j = 0; j < K; j++) { for (int q = 0; q < 5; q++) centers[(size_t)j * 5 + q] = 0; counts[j] = 0; }
        for (int j = 0; j < cols; j++)
            for (int k = 0; k < rows; k++) {
                const int id = clusters[(size_t)j * rows + k];
                if (id != -1) {
                    double *c = centers + (size_t)id * 5;
                    c[0] += LAB(k, j, 0); c[1] += LAB(k, j, 1); c[2] += LAB(k, j, 2); c[3] += j; c[4] += k;
                    counts[id] += 1;
                }
            }
        for (int j = 0; j < K; j++)
            for (int q = 0; q < 5; q++) centers[(size_t)j * 5 + q] /= counts[j];
    }
#undef LAB
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) labels_out[(size_t)y * cols + x] = clusters[(size_t)x * rows + y];
    if (centers_out) memcpy(centers_out, centers, (size_t)K * 5 * sizeof(double));
    free(clusters); free(distances); free(centers); free(counts);
    return K;
}
