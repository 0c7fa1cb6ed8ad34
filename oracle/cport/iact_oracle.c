/*
 * Oracle, C form: plain-C restatement of the reference's single-mirror render path.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py) -- never linked into the product.
 *
 * Follows (float32, one op per reference op, compiled with -ffp-contract=off):
 *   iactrace/core/render.py:118-157      _trace_single_mirror  (stage 0 only)
 *   iactrace/core/render.py:21-41        _check_occlusions
 *   iactrace/core/intersections.py:6-41  intersect_plane
 *   iactrace/core/intersections.py:44-87 intersect_cylinder
 *   iactrace/core/intersections.py:90-110 intersect_box
 *   iactrace/core/intersections.py:195-226 intersect_sphere
 *   iactrace/core/reflection.py:5-19     reflect
 *   iactrace/sensors/square.py:66-91     SquareSensor.accumulate
 *   iactrace/sensors/hexagonal.py:174-194 HexagonalSensor.accumulate
 * The world-frame sample tables (mirrors.py:64-79) are computed by oracle/trace.py and passed in.
 * Threads split the facet loop (the reference's lax.scan axis); each thread owns a private image.
 */
#include <tgmath.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Arithmetic type of the per-ray chain: float (the reference's dtype; the bit-exact build and the CPU baseline) or,
 * with -DORACLE_REAL=double, float64 -- the same operations in the same order at a precision where shadow and
 * pixel-edge decisions are exact to ~1e-13 (full-size image checks, tools/parity_fullsize.py).  Tables, sources and
 * the sensor description stay float32 in both (they are the inputs). */
#ifndef ORACLE_REAL
#define ORACLE_REAL float
#endif
typedef ORACLE_REAL real;

typedef struct {
    int kind;                 /* 0 square, 1 hexagonal */
    float pos[3];
    float R[9];               /* euler_to_matrix(rotation), row-major */
    int width, height;
    float x0, y0, dx, dy, edge_width;
    float goff[2], cr, sr;    /* grid offset, cos/sin(-grid_rotation) */
    float size, size_sqrt3, size_1p5, inradius, edge_thr;
    int q_min, r_min, tq, tr, n_pixels;
    const int* lookup;
} OracleSensor;

static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static real cyl_t(const real* o, const real* d, const real* p1, const real* ax, real h, real r) {
    const real eps = (real)1e-8;
    real oc[3] = {o[0] - p1[0], o[1] - p1[1], o[2] - p1[2]};
    real oc_ax = dot3(oc, ax), rd_ax = dot3(d, ax);
    real ocp[3], rdp[3];
    for (int i = 0; i < 3; ++i) { ocp[i] = oc[i] - oc_ax * ax[i]; rdp[i] = d[i] - rd_ax * ax[i]; }
    real a = dot3(rdp, rdp), b = 2.0f * dot3(ocp, rdp), c = dot3(ocp, ocp) - r * r;
    real disc = b * b - 4.0f * a * c;
    real sq = sqrt(fmax(disc, 0.0f));
    real t1 = (-b - sq) / (2.0f * a + eps), t2 = (-b + sq) / (2.0f * a + eps);
    real y1 = oc_ax + t1 * rd_ax, y2 = oc_ax + t2 * rd_ax;
    t1 = (t1 > eps && y1 >= 0 && y1 <= h && disc >= 0) ? t1 : INFINITY;
    t2 = (t2 > eps && y2 >= 0 && y2 <= h && disc >= 0) ? t2 : INFINITY;
    real tb = -oc_ax / (rd_ax + eps), tt = (h - oc_ax) / (rd_ax + eps);
    real pb[3], pt[3];
    for (int i = 0; i < 3; ++i) { pb[i] = ocp[i] + tb * rdp[i]; pt[i] = ocp[i] + tt * rdp[i]; }
    tb = (tb > eps && dot3(pb, pb) <= r * r) ? tb : INFINITY;
    tt = (tt > eps && dot3(pt, pt) <= r * r) ? tt : INFINITY;
    return fmin(fmin(t1, t2), fmin(tb, tt));
}

static real box_t(const real* o, const real* d, const real* p1, const real* p2) {
    const real eps = (real)1e-8;
    real tmin = -INFINITY, tmax = INFINITY;
    for (int i = 0; i < 3; ++i) {
        real lo = fmin(p1[i], p2[i]), hi = fmax(p1[i], p2[i]);
        real inv = 1.0f / (d[i] + eps);
        real a = (lo - o[i]) * inv, b = (hi - o[i]) * inv;
        tmin = fmax(tmin, fmin(a, b));
        tmax = fmin(tmax, fmax(a, b));
    }
    int hit = (tmax >= tmin) && (tmax > eps);
    real tr = tmin > eps ? tmin : tmax;
    return hit ? tr : INFINITY;
}

static real sph_t(const real* o, const real* d, const real* c, real r) {
    const real eps = (real)1e-8;
    real oc[3] = {o[0] - c[0], o[1] - c[1], o[2] - c[2]};
    real a = dot3(d, d), b = 2.0f * dot3(oc, d), cc = dot3(oc, oc) - r * r;
    real disc = b * b - 4.0f * a * cc;
    real sq = sqrt(fmax(disc, 0.0f));
    real t1 = (-b - sq) / (2.0f * a + eps), t2 = (-b + sq) / (2.0f * a + eps);
    t1 = (t1 > eps && disc >= 0) ? t1 : INFINITY;
    t2 = (t2 > eps && disc >= 0) ? t2 : INFINITY;
    return fmin(t1, t2);
}

static int square_pixel(const OracleSensor* s, real x, real y) {
    real xc = (x - s->x0) / s->dx, yc = (y - s->y0) / s->dy;
    real xf = floor(xc), yf = floor(yc);
    if (!(xf >= 0 && xf < (real)s->width && yf >= 0 && yf < (real)s->height)) return -1;
    real fx = xc - xf, fy = yc - yf;
    real dist = fmin(fmin(fx, 1.0f - fx) * s->dx, fmin(fy, 1.0f - fy) * s->dy);
    if (dist < s->edge_width) return -1;
    return (int)yf * s->width + (int)xf;
}

static int hex_pixel(const OracleSensor* s, real x, real y) {
    real tx = x - s->goff[0], ty = y - s->goff[1];
    real xg = s->cr * tx - s->sr * ty, yg = s->sr * tx + s->cr * ty;
    real q = ((real)0.5773502691896257 * xg - yg / 3.0f) / s->size, r = (2.0f * yg / 3.0f) / s->size;
    real sc = -q - r;
    real qi = rint(q), ri = rint(r), si = rint(sc);
    real dq = fabs(qi - q), dr = fabs(ri - r), ds = fabs(si - sc);
    if (dq > dr && dq > ds) qi = -ri - si;
    if (dr > dq && dr > ds) ri = -qi - si;
    real qx = qi - (real)s->q_min, rx = ri - (real)s->r_min;
    if (!(qx >= 0 && qx < (real)s->tq && rx >= 0 && rx < (real)s->tr)) return -1;
    int pix = s->lookup[(int)qx * s->tr + (int)rx];
    if (pix < 0) return -1;
    real cx = s->size_sqrt3 * (qi + ri / 2.0f), cy = s->size_1p5 * ri;
    real ax = fabs(xg - cx), ay = fabs(yg - cy);
    real hn = fmax(ax, 0.5f * ax + (real)0.8660254037844386 * ay) / s->inradius;
    if (hn > s->edge_thr) return -1;
    return pix;
}

int oracle_render(int F, int M, const float* tp, const float* tn, const float* tw,
                  int S, const float* sources, const float* values, int source_type /*0 point, 1 parallel*/,
                  int n_cyl, const float* cyl_p1, const float* cyl_p2, const float* cyl_r,
                  int n_box, const float* box_p1, const float* box_p2,
                  int n_sph, const float* sph_c, const float* sph_r,
                  const OracleSensor* sens, float* image, float* dbg_xy, float* dbg_val, int n_threads) {
    const int npix = sens->kind == 0 ? sens->width * sens->height : sens->n_pixels;
    /* obstruction tables in the arithmetic type (a no-op copy for float) */
    real* cp1 = (real*)malloc(sizeof(real) * 3 * (n_cyl > 0 ? n_cyl : 1));
    real* crr = (real*)malloc(sizeof(real) * (n_cyl > 0 ? n_cyl : 1));
    real* bp1 = (real*)malloc(sizeof(real) * 3 * (n_box > 0 ? n_box : 1));
    real* bp2 = (real*)malloc(sizeof(real) * 3 * (n_box > 0 ? n_box : 1));
    real* spc = (real*)malloc(sizeof(real) * 4 * (n_sph > 0 ? n_sph : 1));
    for (int i = 0; i < 3 * n_cyl; ++i) cp1[i] = cyl_p1[i];
    for (int i = 0; i < n_cyl; ++i) crr[i] = cyl_r[i];
    for (int i = 0; i < 3 * n_box; ++i) { bp1[i] = box_p1[i]; bp2[i] = box_p2[i]; }
    for (int i = 0; i < n_sph; ++i) { spc[4 * i] = sph_c[3 * i]; spc[4 * i + 1] = sph_c[3 * i + 1]; spc[4 * i + 2] = sph_c[3 * i + 2]; spc[4 * i + 3] = sph_r[i]; }
    /* intersections.py:46-48 hoisted: same ops, evaluated once per cylinder */
    real* cax = (real*)malloc(sizeof(real) * 4 * (n_cyl > 0 ? n_cyl : 1));
    for (int k = 0; k < n_cyl; ++k) {
        real a[3] = {(real)cyl_p2[3 * k] - cp1[3 * k], (real)cyl_p2[3 * k + 1] - cp1[3 * k + 1], (real)cyl_p2[3 * k + 2] - cp1[3 * k + 2]};
        real h = sqrt(dot3(a, a));
        cax[4 * k] = a[0] / h; cax[4 * k + 1] = a[1] / h; cax[4 * k + 2] = a[2] / h; cax[4 * k + 3] = h;
    }
    const real u1[3] = {sens->R[0], sens->R[3], sens->R[6]}, u2[3] = {sens->R[1], sens->R[4], sens->R[7]};
    const real nrm[3] = {sens->R[2], sens->R[5], sens->R[8]};
    const real spos[3] = {sens->pos[0], sens->pos[1], sens->pos[2]};
    const real ndotp = dot3(nrm, spos);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    int nt = omp_get_max_threads();
#else
    int nt = 1;
#endif
    /* per-thread images in float64: the pixel sums are what is checked; the order of the reference's float32
     * segment_sum + `acc + img` (render.py:216) is not reproducible on any other machine anyway */
    double* priv = image ? (double*)calloc((size_t)nt * npix, sizeof(double)) : NULL;
#pragma omp parallel for schedule(dynamic, 1)
    for (int f = 0; f < F; ++f) {
#ifdef _OPENMP
        double* img = priv ? priv + (size_t)omp_get_thread_num() * npix : NULL;
#else
        double* img = priv;
#endif
        for (int s = 0; s < S; ++s) {
            const real src[3] = {sources[3 * s], sources[3 * s + 1], sources[3 * s + 2]};
            for (int m = 0; m < M; ++m) {
                const size_t row = ((size_t)f * M + m) * 3;
                const real p[3] = {tp[row], tp[row + 1], tp[row + 2]}, n[3] = {tn[row], tn[row + 1], tn[row + 2]};
                real d[3];
                if (source_type == 0) {
                    d[0] = p[0] - src[0]; d[1] = p[1] - src[1]; d[2] = p[2] - src[2];
                    real nr = sqrt(dot3(d, d));
                    d[0] /= nr; d[1] /= nr; d[2] /= nr;
                } else { d[0] = src[0]; d[1] = src[1]; d[2] = src[2]; }
                real u[3] = {-d[0], -d[1], -d[2]};
                real shadow = 1.0f, t;
                if (n_cyl) { t = INFINITY; for (int k = 0; k < n_cyl; ++k) t = fmin(t, cyl_t(p, u, cp1 + 3 * k, cax + 4 * k, cax[4 * k + 3], crr[k])); shadow *= (t < (real)1e10) ? 0.0f : 1.0f; }
                if (n_box) { t = INFINITY; for (int k = 0; k < n_box; ++k) t = fmin(t, box_t(p, u, bp1 + 3 * k, bp2 + 3 * k)); shadow *= (t < (real)1e10) ? 0.0f : 1.0f; }
                if (n_sph) { t = INFINITY; for (int k = 0; k < n_sph; ++k) t = fmin(t, sph_t(p, u, spc + 4 * k, spc[4 * k + 3])); shadow *= (t < (real)1e10) ? 0.0f : 1.0f; }
                real c = dot3(d, n);
                real rf[3] = {d[0] - 2.0f * c * n[0], d[1] - 2.0f * c * n[1], d[2] - 2.0f * c * n[2]};
                real val = (real)values[s] * (-c) / (real)tw[(size_t)f * M + m] * shadow;
                real ndotd = dot3(rf, nrm), ndoto = dot3(p, nrm);
                int parallel = fabs(ndotd) < (real)1e-10;
                real tt = (ndotp - ndoto) / (parallel ? 1.0f : ndotd);
                real op[3] = {p[0] + tt * rf[0] - spos[0], p[1] + tt * rf[1] - spos[1], p[2] + tt * rf[2] - spos[2]};
                real x = dot3(op, u1), y = dot3(op, u2);
                if (parallel || tt <= 0) { x = (real)1e10; y = (real)1e10; }
                if (dbg_xy) {
                    size_t ri = ((size_t)f * S + s) * M + m;
                    dbg_xy[2 * ri] = (float)x; dbg_xy[2 * ri + 1] = (float)y; dbg_val[ri] = (float)val;
                }
                if (img) {
                    int pix = sens->kind == 0 ? square_pixel(sens, x, y) : hex_pixel(sens, x, y);
                    if (pix >= 0) img[pix] += val;
                }
            }
        }
    }
    if (image) {
        for (int i = 0; i < npix; ++i) {
            double a = 0.0;
            for (int t = 0; t < nt; ++t) a += priv[(size_t)t * npix + i];
            image[i] = (float)a;
        }
        free(priv);
    }
    free(cax); free(cp1); free(crr); free(bp1); free(bp2); free(spc);
    return nt;
}
