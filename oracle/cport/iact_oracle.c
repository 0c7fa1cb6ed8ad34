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
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

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

static inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static float cyl_t(const float* o, const float* d, const float* p1, const float* ax, float h, float r) {
    const float eps = 1e-8f;
    float oc[3] = {o[0] - p1[0], o[1] - p1[1], o[2] - p1[2]};
    float oc_ax = dot3(oc, ax), rd_ax = dot3(d, ax);
    float ocp[3], rdp[3];
    for (int i = 0; i < 3; ++i) { ocp[i] = oc[i] - oc_ax * ax[i]; rdp[i] = d[i] - rd_ax * ax[i]; }
    float a = dot3(rdp, rdp), b = 2.0f * dot3(ocp, rdp), c = dot3(ocp, ocp) - r * r;
    float disc = b * b - 4.0f * a * c;
    float sq = sqrtf(fmaxf(disc, 0.0f));
    float t1 = (-b - sq) / (2.0f * a + eps), t2 = (-b + sq) / (2.0f * a + eps);
    float y1 = oc_ax + t1 * rd_ax, y2 = oc_ax + t2 * rd_ax;
    t1 = (t1 > eps && y1 >= 0 && y1 <= h && disc >= 0) ? t1 : INFINITY;
    t2 = (t2 > eps && y2 >= 0 && y2 <= h && disc >= 0) ? t2 : INFINITY;
    float tb = -oc_ax / (rd_ax + eps), tt = (h - oc_ax) / (rd_ax + eps);
    float pb[3], pt[3];
    for (int i = 0; i < 3; ++i) { pb[i] = ocp[i] + tb * rdp[i]; pt[i] = ocp[i] + tt * rdp[i]; }
    tb = (tb > eps && dot3(pb, pb) <= r * r) ? tb : INFINITY;
    tt = (tt > eps && dot3(pt, pt) <= r * r) ? tt : INFINITY;
    return fminf(fminf(t1, t2), fminf(tb, tt));
}

static float box_t(const float* o, const float* d, const float* p1, const float* p2) {
    const float eps = 1e-8f;
    float tmin = -INFINITY, tmax = INFINITY;
    for (int i = 0; i < 3; ++i) {
        float lo = fminf(p1[i], p2[i]), hi = fmaxf(p1[i], p2[i]);
        float inv = 1.0f / (d[i] + eps);
        float a = (lo - o[i]) * inv, b = (hi - o[i]) * inv;
        tmin = fmaxf(tmin, fminf(a, b));
        tmax = fminf(tmax, fmaxf(a, b));
    }
    int hit = (tmax >= tmin) && (tmax > eps);
    float tr = tmin > eps ? tmin : tmax;
    return hit ? tr : INFINITY;
}

static float sph_t(const float* o, const float* d, const float* c, float r) {
    const float eps = 1e-8f;
    float oc[3] = {o[0] - c[0], o[1] - c[1], o[2] - c[2]};
    float a = dot3(d, d), b = 2.0f * dot3(oc, d), cc = dot3(oc, oc) - r * r;
    float disc = b * b - 4.0f * a * cc;
    float sq = sqrtf(fmaxf(disc, 0.0f));
    float t1 = (-b - sq) / (2.0f * a + eps), t2 = (-b + sq) / (2.0f * a + eps);
    t1 = (t1 > eps && disc >= 0) ? t1 : INFINITY;
    t2 = (t2 > eps && disc >= 0) ? t2 : INFINITY;
    return fminf(t1, t2);
}

static int square_pixel(const OracleSensor* s, float x, float y) {
    float xc = (x - s->x0) / s->dx, yc = (y - s->y0) / s->dy;
    float xf = floorf(xc), yf = floorf(yc);
    if (!(xf >= 0 && xf < (float)s->width && yf >= 0 && yf < (float)s->height)) return -1;
    float fx = xc - xf, fy = yc - yf;
    float dist = fminf(fminf(fx, 1.0f - fx) * s->dx, fminf(fy, 1.0f - fy) * s->dy);
    if (dist < s->edge_width) return -1;
    return (int)yf * s->width + (int)xf;
}

static int hex_pixel(const OracleSensor* s, float x, float y) {
    float tx = x - s->goff[0], ty = y - s->goff[1];
    float xg = s->cr * tx - s->sr * ty, yg = s->sr * tx + s->cr * ty;
    float q = (0.5773502691896257f * xg - yg / 3.0f) / s->size, r = (2.0f * yg / 3.0f) / s->size;
    float sc = -q - r;
    float qi = rintf(q), ri = rintf(r), si = rintf(sc);
    float dq = fabsf(qi - q), dr = fabsf(ri - r), ds = fabsf(si - sc);
    if (dq > dr && dq > ds) qi = -ri - si;
    if (dr > dq && dr > ds) ri = -qi - si;
    float qx = qi - (float)s->q_min, rx = ri - (float)s->r_min;
    if (!(qx >= 0 && qx < (float)s->tq && rx >= 0 && rx < (float)s->tr)) return -1;
    int pix = s->lookup[(int)qx * s->tr + (int)rx];
    if (pix < 0) return -1;
    float cx = s->size_sqrt3 * (qi + ri / 2.0f), cy = s->size_1p5 * ri;
    float ax = fabsf(xg - cx), ay = fabsf(yg - cy);
    float hn = fmaxf(ax, 0.5f * ax + 0.8660254037844386f * ay) / s->inradius;
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
    /* intersections.py:46-48 hoisted: same float ops, evaluated once per cylinder */
    float* cax = (float*)malloc(sizeof(float) * 4 * (n_cyl > 0 ? n_cyl : 1));
    for (int k = 0; k < n_cyl; ++k) {
        float a[3] = {cyl_p2[3 * k] - cyl_p1[3 * k], cyl_p2[3 * k + 1] - cyl_p1[3 * k + 1], cyl_p2[3 * k + 2] - cyl_p1[3 * k + 2]};
        float h = sqrtf(dot3(a, a));
        cax[4 * k] = a[0] / h; cax[4 * k + 1] = a[1] / h; cax[4 * k + 2] = a[2] / h; cax[4 * k + 3] = h;
    }
    const float u1[3] = {sens->R[0], sens->R[3], sens->R[6]}, u2[3] = {sens->R[1], sens->R[4], sens->R[7]};
    const float nrm[3] = {sens->R[2], sens->R[5], sens->R[8]};
    const float ndotp = dot3(nrm, sens->pos);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    int nt = omp_get_max_threads();
#else
    int nt = 1;
#endif
    float* priv = image ? (float*)calloc((size_t)nt * npix, sizeof(float)) : NULL;
#pragma omp parallel for schedule(dynamic, 1)
    for (int f = 0; f < F; ++f) {
#ifdef _OPENMP
        float* img = priv ? priv + (size_t)omp_get_thread_num() * npix : NULL;
#else
        float* img = priv;
#endif
        for (int s = 0; s < S; ++s) {
            const float* src = sources + 3 * s;
            for (int m = 0; m < M; ++m) {
                const float* p = tp + ((size_t)f * M + m) * 3;
                const float* n = tn + ((size_t)f * M + m) * 3;
                float d[3];
                if (source_type == 0) {
                    d[0] = p[0] - src[0]; d[1] = p[1] - src[1]; d[2] = p[2] - src[2];
                    float nr = sqrtf(dot3(d, d));
                    d[0] /= nr; d[1] /= nr; d[2] /= nr;
                } else { d[0] = src[0]; d[1] = src[1]; d[2] = src[2]; }
                float u[3] = {-d[0], -d[1], -d[2]};
                float shadow = 1.0f, t;
                if (n_cyl) { t = INFINITY; for (int k = 0; k < n_cyl; ++k) t = fminf(t, cyl_t(p, u, cyl_p1 + 3 * k, cax + 4 * k, cax[4 * k + 3], cyl_r[k])); shadow *= (t < 1e10f) ? 0.0f : 1.0f; }
                if (n_box) { t = INFINITY; for (int k = 0; k < n_box; ++k) t = fminf(t, box_t(p, u, box_p1 + 3 * k, box_p2 + 3 * k)); shadow *= (t < 1e10f) ? 0.0f : 1.0f; }
                if (n_sph) { t = INFINITY; for (int k = 0; k < n_sph; ++k) t = fminf(t, sph_t(p, u, sph_c + 3 * k, sph_r[k])); shadow *= (t < 1e10f) ? 0.0f : 1.0f; }
                float c = dot3(d, n);
                float rf[3] = {d[0] - 2.0f * c * n[0], d[1] - 2.0f * c * n[1], d[2] - 2.0f * c * n[2]};
                float val = values[s] * (-c) / tw[(size_t)f * M + m] * shadow;
                float ndotd = dot3(rf, nrm), ndoto = dot3(p, nrm);
                int parallel = fabsf(ndotd) < 1e-10f;
                float tt = (ndotp - ndoto) / (parallel ? 1.0f : ndotd);
                float op[3] = {p[0] + tt * rf[0] - sens->pos[0], p[1] + tt * rf[1] - sens->pos[1], p[2] + tt * rf[2] - sens->pos[2]};
                float x = dot3(op, u1), y = dot3(op, u2);
                if (parallel || tt <= 0) { x = 1e10f; y = 1e10f; }
                if (dbg_xy) {
                    size_t ri = ((size_t)f * S + s) * M + m;
                    dbg_xy[2 * ri] = x; dbg_xy[2 * ri + 1] = y; dbg_val[ri] = val;
                }
                if (img) {
                    int pix = sens->kind == 0 ? square_pixel(sens, x, y) : hex_pixel(sens, x, y);
                    if (pix >= 0) img[pix] += val;
                }
            }
        }
    }
    if (image) {
        memset(image, 0, sizeof(float) * npix);
        for (int t = 0; t < nt; ++t)
            for (int i = 0; i < npix; ++i) image[i] += priv[(size_t)t * npix + i];
        free(priv);
    }
    free(cax);
    return nt;
}
