// K6: vector-Jacobian product of `render` -- what jax.grad of reference core/render.py:174-220 yields
// (SURVEY.md section 3.5), hand-derived per ray and accumulated with warp reductions.
//
// Differentiable leaves: stage-0 facet positions / rotations (Euler degrees) / perturbation_scale /
// weights, the sources and values, and the sensor position / rotation.  Decisions (shadow mask,
// pixel index, cube rounding, `where` guards) are piecewise constant and carry zero gradient, as in
// JAX; with the hard sensors only d(value) flows, with the soft (Gaussian-splat) sensors
// d(image)/d(x, y) flows too (sensors/square.py:144-172, sensors/hexagonal.py:264-314).
//
// Forward per ray (mirrors.py:64-79, render.py:129-155):
//   p = R p_l + pos;  nw = R (n_l + scale d_l);  n = nw/|nw|;  d = (p - src)/|p - src|  |  src
//   c = d.n;  r = d - 2 c n;  val = v (-c) / w * shadow
//   t = (ns.ps - ns.p)/(ns.r);  h = p + t r - ps;  x = h.u1;  y = h.u2;  image += val * W(x, y)
#include "iact_cull.cuh"

namespace {

// Work queue of the VJP kernel: unit = (facet, sample part, run of `slen` consecutive sources).
struct VjpPlan { int S, slen, sruns, msplit, msize; long long n_units; unsigned long long* counter; };

VjpPlan make_vjp_plan(const SceneDev& d, int S, long long resident_warps) {
    VjpPlan p;
    p.S = S;
    p.slen = (int)std::max(1LL, std::min(32LL, (long long)S * d.F / (resident_warps * 32)));
    p.sruns = (S + p.slen - 1) / p.slen;
    const long long base = (long long)d.F * p.sruns;
    long long ms = 1;                                        // small jobs: split along the samples so every warp gets work
    if (base < resident_warps * 4) ms = std::min<long long>((d.M + 31) / 32, (resident_warps * 4 + base - 1) / std::max(base, 1LL));
    ms = std::max(ms, 1LL);
    p.msize = (int)(((d.M + ms - 1) / ms + 31) / 32 * 32);
    p.msplit = (d.M + p.msize - 1) / std::max(p.msize, 1);
    p.n_units = base * p.msplit;
    p.counter = nullptr;
    return p;
}

struct GradsDev {
    float *weights, *values, *sources;
    float* facc;   // (F,13): [0..2] tau (dL/d rotation as an axial vector, see vjp_kernel), [3..8] unused, [9..11] dL/dpos, [12] dL/dscale
    float* sacc;   // 12: dL/d sensor pos (3), dL/d sensor R (9: u1,u2,n as columns -> row-major R)
    float* macc;   // (N2,MACC) or null: per stage>=1 mirror dL/dR row-major (9), dL/dpos (3), dL/d(c, k, x0, y0) (4)
    float *points, *nq;   // (F,M,3) or null: dL/d(local sample point), dL/d(local normal + scale * delta)
    const float* Gn;      // (P,8) or null: cotangent gathered around every pixel (soft hex sensor, one ring)
};

// Soft hex sensor, one ring of neighbours: the cotangent is pre-gathered per pixel into rows of 8 floats
// (gather_cotangent_kernel: row p = G at the 7 hexagons around pixel p in tap order, 0 where there is no pixel), so a
// hit costs one table lookup (its base hexagon) and two 128-bit loads instead of 7 lookups + 7 loads.
__global__ void __launch_bounds__(256) gather_cotangent_kernel(const SensDev se, const float* __restrict__ G, float* __restrict__ Gn) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= se.tq * se.tr) return;
    const int pix = se.lookup[p];
    if (pix < 0) return;
    const float qb = (float)(p / se.tr + se.qmin), rb = (float)(p % se.tr + se.rmin);
    int t = 0;
    for (int oq = -1; oq <= 1; ++oq)
        for (int orr = -1; orr <= 1; ++orr) {
            if (oq + orr < -1 || oq + orr > 1) continue;
            const int nb = hex_lookup(se, se.lookup, qb + (float)oq, rb + (float)orr);
            Gn[8 * (size_t)pix + t++] = nb >= 0 ? G[nb] : 0.f;
        }
    Gn[8 * (size_t)pix + 7] = 0.f;
}

// d(image . G)/d(val) and /d(x, y) for one hit.  Returns false if the hit contributes nothing.
template <int SENS, typename LUT>
__device__ __forceinline__ bool sensor_adjoint(const SensDev& se, const LUT* lut, const float* __restrict__ G,
                                               float x, float y, float& dval, float& dx, float& dy, const float* __restrict__ Gn) {
    dx = 0.f; dy = 0.f; dval = 0.f;
    if (se.kind == IACT_SENSOR_SQUARE) {
        const int pix = square_pixel(se, x, y);
        if (pix < 0) return false;
        dval = __ldg(G + pix);
        return true;
    }
    if (se.kind == IACT_SENSOR_HEX) {
        const int pix = hex_pixel(se, lut, x, y);
        if (pix < 0) return false;
        dval = __ldg(G + pix);
        return true;
    }
    if (SENS == SENS_SQUARE) {
        // image += val * sum_i g_i w_i / sum_i w_i,  w_i = exp(-((fx-ox)^2 + (fy-oy)^2) / (2 sigma^2))
        const float xp = (x - se.x0) * se.inv_dx, yp = (y - se.y0) * se.inv_dy;
        const float xb = floorf(xp), yb = floorf(yp);
        const int K = se.ksize;
        if (!(xb >= (float)(-K - 1) && xb <= (float)(se.W + K) && yb >= (float)(-K - 1) && yb <= (float)(se.H + K))) return false;
        const float fx = xp - xb, fy = yp - yb;
        const float inv_s2 = 1.0f / (se.sigma * se.sigma);
        float D = 0.f, Nn = 0.f, gx = 0.f, gy = 0.f, wx = 0.f, wy = 0.f;
        for (int oy = -K; oy <= K; ++oy)
            for (int ox = -K; ox <= K; ++ox) {
                const float ddx = fx - (float)ox, ddy = fy - (float)oy;
                const float w = gauss_half((ddx * ddx + ddy * ddy) * inv_s2);
                const float dwx = -w * ddx * inv_s2, dwy = -w * ddy * inv_s2;
                const int xi = (int)xb + ox, yi = (int)yb + oy;
                const float g = (xi >= 0 && xi < se.W && yi >= 0 && yi < se.H) ? __ldg(G + (size_t)yi * se.W + xi) : 0.f;
                D += w; Nn += g * w; gx += g * dwx; gy += g * dwy; wx += dwx; wy += dwy;
            }
        const float invD = frcp_nr(D);
        dval = Nn * invD;
        dx = (gx - dval * wx) * invD * se.inv_dx;
        dy = (gy - dval * wy) * invD * se.inv_dy;
        return true;
    }
    // soft hexagonal
    float xg, yg; hex_grid_coords(se, x, y, xg, yg);
    const float q = se.ax_qx * xg - se.ax_qy * yg, r = se.ax_ry * yg;
    float qb, rb; hex_round(q, r, qb, rb);
    if (!(fabsf(qb) < 1e6f && fabsf(rb) < 1e6f)) return false;
    const float ddx = xg - se.size_sqrt3 * (qb + rb * 0.5f), ddy = yg - se.size_1p5 * rb;
    const int K = se.ksize;
    float D = 0.f, Nn = 0.f, gx = 0.f, gy = 0.f, wx = 0.f, wy = 0.f;
    // w = exp(-(hd / sigma)^2 / 2) with hd = m / inradius, m = max(|a|, |a|/2 + sqrt(3)/2 |b|) (hexagonal.py:42-47,287):
    // w = 2^(soft_nk m^2);  dw/d(m^2) = -w / (2 (inradius sigma)^2);  d(m^2)/da = 2a where m = |a|, else sgn(a) m, and
    // d(m^2)/db = 0 resp. sqrt(3) sgn(b) m  (sgn(0) = 0 as in jnp.abs only matters on a set of measure zero)
    const float dwm = se.soft_nk * 0.69314718055994530942f;          // soft_nk = -log2(e) / (2 (inradius sigma)^2)
    float g7[7];
    if (K == 1) {
        const int pb = Gn ? hex_lookup(se, lut, qb, rb) : -1;
        if (pb >= 0) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(Gn) + 2 * pb), g1 = __ldg(reinterpret_cast<const float4*>(Gn) + 2 * pb + 1);
            g7[0] = g0.x; g7[1] = g0.y; g7[2] = g0.z; g7[3] = g0.w; g7[4] = g1.x; g7[5] = g1.y; g7[6] = g1.z;
        } else {                                            // base hexagon is no pixel (camera rim, holes): tap by tap
            int t = 0;
            for (int oq = -1; oq <= 1; ++oq)
                for (int orr = -1; orr <= 1; ++orr) {
                    if (oq + orr < -1 || oq + orr > 1) continue;
                    const int pix = hex_lookup(se, lut, qb + (float)oq, rb + (float)orr);
                    g7[t++] = pix >= 0 ? __ldg(G + pix) : 0.f;
                }
        }
    }
    auto tap = [&](int oq, int orr, int ti) {
        const float ox = se.size_sqrt3 * ((float)oq + (float)orr * 0.5f), oy = se.size_1p5 * (float)orr;
        const float a = ddx - ox, b = ddy - oy;
        const float aa = fabsf(a), ab = fabsf(b);
        const float alt = 0.5f * aa + 0.8660254037844386f * ab;
        const bool first = aa >= alt;
        const float m = fmaxf(aa, alt);
        float w; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(se.soft_nk * (m * m)));
        const float e = w * dwm;
        const float dwx = e * (first ? a + a : copysignf(alt, a));
        const float dwy = e * (first ? 0.f : copysignf(1.7320508075688772f * alt, b));
        float g;
        if (ti >= 0) g = g7[ti];                            // compile-time: the 7-tap instantiation has no branch here
        else { const int pix = hex_lookup(se, lut, qb + (float)oq, rb + (float)orr); g = pix >= 0 ? __ldg(G + pix) : 0.f; }
        D += w; Nn += g * w; gx += g * dwx; gy += g * dwy; wx += dwx; wy += dwy;
    };
    if (K == 1) {                                           // the common 7-tap case, fully unrolled
        tap(-1, 0, 0); tap(-1, 1, 1); tap(0, -1, 2); tap(0, 0, 3); tap(0, 1, 4); tap(1, -1, 5); tap(1, 0, 6);
    } else {
        for (int oq = -K; oq <= K; ++oq)
            for (int orr = -K; orr <= K; ++orr)
                if (max(max(abs(oq), abs(orr)), abs(oq + orr)) <= K) tap(oq, orr, -1);
    }
    const float invD = frcp_nr(D);
    dval = Nn * invD;
    const float dxg = (gx - dval * wx) * invD, dyg = (gy - dval * wy) * invD;
    // (xg, yg) = Rot(-grid_rotation) (x - off):  xg = cr tx - sr ty, yg = sr tx + cr ty
    dx = se.cr * dxg + se.sr * dyg;
    dy = -se.sr * dxg + se.cr * dyg;
    return true;
}


// ---------------------------------------------------------------- optical stages >= 1
// Forward selection of the mirror a ray hits in one stage (same arithmetic as reflect_at_stage, plus
// the index and ray parameter needed by the backward pass).
__device__ __forceinline__ bool stage_select(int n_mirrors, const float* rec, const float* verts, bool full_scan, V3 o, V3 d,
                                             int& best_mi, float& best_t) {
    best_t = INFINITY; best_mi = -1;
    for (int mi = 0; mi < n_mirrors; ++mi) {
        const float* r = rec + (size_t)mi * STAGE_REC;
        const V3 pos = v3(r[0], r[1], r[2]);
        M33 R;
#pragma unroll
        for (int k = 0; k < 9; ++k) R.m[k] = r[24 + k];
        SurfRef s;
        s.c = r[8]; s.k = r[9]; s.kc2 = r[33]; s.n_asph = (int)r[10]; s.asph = r + 11; s.full_scan = full_scan;
        V3 pl, nl;
        float t = surface_intersect(s, r[6], r[7], r[34], mulT(R, o - pos), mulT(R, d), pl, nl);
        bool inside;
        if (r[19] == 0.f) {
            inside = pl.x * pl.x + pl.y * pl.y <= r[20] * r[20];
        } else {
            const int nv = (int)r[21];
            const float* V = verts + 2 * (size_t)r[22];
            inside = true;
            for (int i = 0; i < nv; ++i) {
                const int j = (i + 1 == nv) ? 0 : i + 1;
                const float cr = (V[2 * j] - V[2 * i]) * (pl.y - V[2 * i + 1]) - (V[2 * j + 1] - V[2 * i + 1]) * (pl.x - V[2 * i]);
                inside = inside && (cr >= 0.f);
            }
        }
        if (!inside) t = INFINITY;
        if (t < best_t) { best_t = t; best_mi = mi; }
    }
    return best_t < IACT_TMAX;
}

// Local geometry of mirror record r at ray parameter t: hit point, unit normal, first and second
// derivatives of the sag (surfaces.py:25-58; d sag/d r2 = c / (2 s), s = sqrt(1 - (1+k) c^2 r2)).
struct StageGeom { M33 R; V3 pos, ol, dl, pl, nl; float sx, sy, sxx, sxy, syy, inv_m;
                   float X, Y, r2, inv_s, c, kc2, x0, y0; const float* rec; };
#define MACC 16   // floats per stage >= 1 mirror accumulator: dL/dR (9), dL/dpos (3), dL/d(curvature, conic, offset x, offset y)

__device__ __forceinline__ StageGeom stage_geometry(const float* r, V3 o, V3 d, float t) {
    StageGeom g;
    g.pos = v3(r[0], r[1], r[2]);
#pragma unroll
    for (int k = 0; k < 9; ++k) g.R.m[k] = r[24 + k];
    SurfRef s;
    s.c = r[8]; s.k = r[9]; s.kc2 = r[33]; s.n_asph = (int)r[10]; s.asph = r + 11; s.full_scan = false;
    g.ol = mulT(g.R, o - g.pos); g.dl = mulT(g.R, d);
    const float x0 = r[6], y0 = r[7];
    const float x = g.ol.x + t * g.dl.x, y = g.ol.y + t * g.dl.y;
    const float X = x + x0, Y = y + y0, r2 = X * X + Y * Y;
    const float ss = 1.0f - s.kc2 * r2;
    const float inv_s = rsqrtf(ss);
    float f1 = 0.5f * s.c * inv_s;                                   // f'(r2)
    float f2 = 0.25f * s.c * s.kc2 * inv_s * inv_s * inv_s;          // f''(r2)
    if (s.n_asph > 0) {
        float r4 = r2 * r2, p1 = r2, p0 = 1.0f;                      // r2^(2i+1), r2^(2i)
        for (int i = 0; i < s.n_asph; ++i) {
            const float e = (float)(2 * i + 2);
            f1 += s.asph[i] * e * p1;
            f2 += s.asph[i] * e * (e - 1.0f) * p0;
            p1 *= r4; p0 *= r4;
        }
    }
    g.sx = 2.0f * X * f1; g.sy = 2.0f * Y * f1;
    g.sxx = 2.0f * f1 + 4.0f * X * X * f2; g.sxy = 4.0f * X * Y * f2; g.syy = 2.0f * f1 + 4.0f * Y * Y * f2;
    g.pl = v3(x, y, sag_fast(s, X, Y) - sag_fast(s, x0, y0));
    g.X = X; g.Y = Y; g.r2 = r2; g.inv_s = inv_s; g.c = s.c; g.kc2 = s.kc2; g.x0 = x0; g.y0 = y0; g.rec = r;
    const V3 m = v3(-g.sx, -g.sy, 1.0f);
    g.inv_m = rsqrtf(dot(m, m));
    g.nl = g.inv_m * m;
    return g;
}

// Reverse pass through one stage.  In: adjoints of the outgoing origin (hit point) g_p, outgoing
// direction g_r and outgoing value g_vout.  Out: adjoints of the incoming origin/direction/value.
// t is differentiated implicitly: g(t) = oz + t dz - sag(x, y) = 0 (the fixed point of the Newton scan).
__device__ __forceinline__ void stage_backward(const StageGeom& g, V3 o, V3 d, float t, float val_in, V3 g_p, V3 g_r, float g_vout,
                                               V3& g_o, V3& g_d, float& g_vin, float* macc) {
    const V3 nw = mul(g.R, g.nl);
    const float c = dot(d, nw);
    // refl = d - 2 c n ; val_out = val_in |c|
    g_d = g_r;
    float g_c = -2.0f * dot(g_r, nw) + g_vout * val_in * (c > 0.f ? 1.f : (c < 0.f ? -1.f : 0.f));
    V3 g_nw = (-2.0f * c) * g_r;
    g_vin = g_vout * fabsf(c);
    g_d = g_d + g_c * nw;
    g_nw = g_nw + g_c * d;
    // world -> local
    const V3 g_pl = mulT(g.R, g_p), g_nl = mulT(g.R, g_nw);
    // n_l = m/|m|, m = (-sx, -sy, 1)
    const V3 g_m = g.inv_m * (g_nl - dot(g_nl, g.nl) * g.nl);
    const float g_sx = -g_m.x, g_sy = -g_m.y;
    // p_l = (x, y, sag(x, y) - z0)
    float g_x = g_pl.x + g_pl.z * g.sx + g_sx * g.sxx + g_sy * g.sxy;
    float g_y = g_pl.y + g_pl.z * g.sy + g_sx * g.sxy + g_sy * g.syy;
    // x = ol.x + t dl.x, y = ol.y + t dl.y
    V3 g_ol = v3(g_x, g_y, 0.f), g_dl = v3(t * g_x, t * g_y, 0.f);
    const float g_t = g_x * g.dl.x + g_y * g.dl.y;
    // implicit: dt = -(dg/d ol . d ol + dg/d dl . d dl) / g',  dg/d ol = (-sx, -sy, 1), dg/d dl = t (-sx, -sy, 1)
    const float gprime = g.dl.z - (g.sx * g.dl.x + g.sy * g.dl.y);
    const float k = -g_t / gprime;
    const V3 q = v3(-g.sx, -g.sy, 1.0f);
    g_ol = g_ol + k * q;
    g_dl = g_dl + (k * t) * q;
    // ol = R^T (o - pos), dl = R^T d
    g_o = mul(g.R, g_ol);
    g_d = g_d + mul(g.R, g_dl);
    if (macc) {
        // mirror pose adjoints: pw = R pl + pos, nw = R nl, ol = R^T (o - pos), dl = R^T d
        const V3 oc = o - g.pos;
        const float a[3] = {g_p.x, g_p.y, g_p.z}, b[3] = {g_nw.x, g_nw.y, g_nw.z};
        const float pl[3] = {g.pl.x, g.pl.y, g.pl.z}, nl[3] = {g.nl.x, g.nl.y, g.nl.z};
        const float oc3[3] = {oc.x, oc.y, oc.z}, d3[3] = {d.x, d.y, d.z};
        const float gol[3] = {g_ol.x, g_ol.y, g_ol.z}, gdl[3] = {g_dl.x, g_dl.y, g_dl.z};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) macc[3 * i + j] += a[i] * pl[j] + b[i] * nl[j] + oc3[i] * gol[j] + d3[i] * gdl[j];
        macc[9] += g_p.x - g_o.x; macc[10] += g_p.y - g_o.y; macc[11] += g_p.z - g_o.z;
        // Surface parameters theta in (curvature c, conic k, offset x0, y0) of surfaces.py:25-45.  With x, y (hence t)
        // held fixed they enter p_l.z = S(X, Y) - S(x0, y0) and the slopes S_X, S_Y; through the root g(t; theta) = 0 they
        // move t by dt/dtheta = P / g', P = d p_l.z / d theta.  Hence
        //   dL/dtheta = (g_pl.z + g_t / g') P_theta + g_sx dS_X/dtheta + g_sy dS_Y/dtheta,
        // with, for the conic part S = c u / (1 + s), u = r^2, s = sqrt(1 - (1 + k) c^2 u):
        //   dS/dc = u / (s (1 + s)),  dS/dk = c^3 u^2 / (2 s (1 + s)^2),  d(S_X)/dc = X / s^3,  d(S_X)/dk = X c^3 u / (2 s^3),
        // and for the offsets (X = x + x0): P = S_X(X, Y) - S_X(x0, y0), dS_X/dx0 = S_XX, dS_Y/dx0 = S_XY (same for y0).
        const float coef = g_pl.z - k;                               // k = -g_t / g'
        const float c1 = g.c, c3 = c1 * c1 * c1;
        auto dS = [&](float u, float inv_s, float& dc, float& dk) {
            const float sq = 1.0f / inv_s, op = 1.0f + sq;
            dc = u * inv_s / op;
            dk = 0.5f * c3 * u * u * inv_s / (op * op);
        };
        float dc_h, dk_h, dc_0, dk_0;
        const float u0 = g.x0 * g.x0 + g.y0 * g.y0;
        const float inv_s0 = rsqrtf(1.0f - g.kc2 * u0);
        dS(g.r2, g.inv_s, dc_h, dk_h);
        dS(u0, inv_s0, dc_0, dk_0);
        const float is3 = g.inv_s * g.inv_s * g.inv_s;
        const float dsl_c = is3, dsl_k = 0.5f * c3 * g.r2 * is3;     // d(S_X)/dtheta = X * these
        macc[12] += coef * (dc_h - dc_0) + (g_sx * g.X + g_sy * g.Y) * dsl_c;
        macc[13] += coef * (dk_h - dk_0) + (g_sx * g.X + g_sy * g.Y) * dsl_k;
        // slopes at the offset point, aspheric terms included
        SurfRef sr;
        sr.c = g.c; sr.k = g.rec[9]; sr.kc2 = g.kc2; sr.n_asph = (int)g.rec[10]; sr.asph = g.rec + 11; sr.full_scan = false;
        const float f10 = dsag_dr2_t(sr, u0);
        macc[14] += coef * (g.sx - 2.0f * g.x0 * f10) + g_sx * g.sxx + g_sy * g.sxy;
        macc[15] += coef * (g.sy - 2.0f * g.y0 * f10) + g_sx * g.sxy + g_sy * g.syy;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

#ifndef IACT_VJP_MIN_BLOCKS
#define IACT_VJP_MIN_BLOCKS 2
#endif
#ifndef IACT_VJP_MIN_BLOCKS_LEAN
#define IACT_VJP_MIN_BLOCKS_LEAN 3
#endif
#ifndef IACT_VJP_MIN_BLOCKS_STAGES
#define IACT_VJP_MIN_BLOCKS_STAGES 2
#endif
// FULL = false: only the facet-pose adjoints (dL/dR, dL/dpos -> rotations, positions) are accumulated -- the
// alignment fit of BASELINE config 5; the sensor-pose, per-source (values, sources), perturbation_scale and
// per-sample weight adjoints and their registers are compiled out.
template <int SRC, int SENS, bool STAGES, bool FULL>
__global__ void __launch_bounds__(256, STAGES ? IACT_VJP_MIN_BLOCKS_STAGES : (FULL ? IACT_VJP_MIN_BLOCKS : IACT_VJP_MIN_BLOCKS_LEAN))
vjp_kernel(const __grid_constant__ SceneDev sc, const IactFacets fa, const float* __restrict__ sources,
           const float* __restrict__ values, const VjpPlan vp, const FacetLists fl,
           const float* __restrict__ G, const GradsDev gr) {
    extern __shared__ __align__(16) float smem_all[];
    float* smem = smem_all;
    ObsSmem ob;
    const bool cull = sc.cull != 0;
    // per-warp cylinder records (as trace_kernel): items whose rays share their direction
    float* wrec = nullptr;
    if (IACT_CYL_RECORDS && cull && sc.n_cyl > 0) {
        wrec = smem + (size_t)(threadIdx.x >> 5) * (CYL_REC_MAX * CYL_REC);
        smem += (size_t)(blockDim.x >> 5) * (CYL_REC_MAX * CYL_REC);
    }
    stage_obstructions(sc, smem, ob, cull);
    const int n_obs = ob.n_cyl + ob.n_rest;
    float* p = smem + obstruction_floats(sc.n_cyl, sc.n_box, sc.n_sph, sc.n_obox, sc.n_tri, cull);
    const float* stage_rec = p;
    if (STAGES) { stage_mirrors(sc, p); p += stage_floats(sc); }
    const short* lut = nullptr;
    if (SENS == SENS_HEX) {
        short* l = reinterpret_cast<short*>(p);
        for (int i = threadIdx.x; i < sc.sens.tq * sc.sens.tr; i += blockDim.x) l[i] = (short)sc.sens.lookup[i];
        lut = l;
        p += (sc.sens.tq * sc.sens.tr + 1) / 2;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned short* list = cull ? reinterpret_cast<unsigned short*>(p) + (size_t)warp * ((n_obs + 1) & ~1) : nullptr;
    __syncthreads();

    const int M = sc.M;
    const SensDev& se = sc.sens;
    const V3 ns = v3(se.nrm[0], se.nrm[1], se.nrm[2]), u1 = v3(se.u1[0], se.u1[1], se.u1[2]), u2 = v3(se.u2[0], se.u2[1], se.u2[2]);
    const V3 ps = v3(se.pos[0], se.pos[1], se.pos[2]);
    // sensor adjoints accumulate per lane over the whole block lifetime
    V3 g_ps = v3(0.f, 0.f, 0.f), g_u1 = g_ps, g_u2 = g_ps, g_ns = g_ps;

    // Work queue: a unit = (facet, sample part, run of sources), pulled by one warp from a global counter.  The
    // facet's pose is set up and its 13 adjoints are warp-reduced once per unit instead of once per (facet, source).
    const bool per_source = FULL && (gr.values != nullptr || gr.sources != nullptr);
    for (;;) {
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(vp.counter, 1ull);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= (unsigned long long)vp.n_units) break;
        const int sr = (int)(u % (unsigned long long)vp.sruns);
        const unsigned long long rest = u / (unsigned long long)vp.sruns;
        const int part = (int)(rest % (unsigned long long)vp.msplit), f = (int)(rest / (unsigned long long)vp.msplit);
        const int s0 = sr * vp.slen, s1 = min(vp.S, s0 + vp.slen);
        const int m0 = part * vp.msize, m1 = min(M, m0 + vp.msize);
        M33 R;
        if (FULL) R = euler_to_matrix(__ldg(fa.rotations + 3 * f), __ldg(fa.rotations + 3 * f + 1), __ldg(fa.rotations + 3 * f + 2));
        const V3 pos = ld3(fa.positions + 3 * f);
        const float scale = __ldg(fa.scale + f);
        const float4 bnd = __ldg(sc.bounds + f);
        // dL/d(rotation) as an axial vector: every derivative of a rotation matrix is dR/dtheta = [a]x R with a world-frame
        // axis a, so dL/dtheta = g_o . (a x (o - pos)) + g_nw . (a x nw) = a . tau,
        // tau = sum (o - pos) x g_o + nw x g_nw = sum (o - pos) x g_o + n x g_n  (three accumulators instead of the nine of dL/dR)
        V3 tau = v3(0.f, 0.f, 0.f);
        V3 g_pos = v3(0.f, 0.f, 0.f);
        float g_scale = 0.f;
        // stage >= 1 mirror adjoints: one register set per lane for the first stage's mirror 0..; rays
        // that use another (stage, mirror) fall back to direct atomics
        float mreg[MACC];
#pragma unroll
        for (int q = 0; q < MACC; ++q) mreg[q] = 0.f;
        int mreg_id = -1;                                      // flat mirror id the register set belongs to

        for (int s = s0; s < s1; ++s) {
            const V3 src = v3(__ldg(sources + 3 * s), __ldg(sources + 3 * s + 1), __ldg(sources + 3 * s + 2));
            const float sval = __ldg(values + s);
            float g_val = 0.f;
            V3 g_src = v3(0.f, 0.f, 0.f);
            // far point source (parallax R / D < 1e-9, as trace_item): one direction for the whole facet, and the
            // adjoint of that direction w.r.t. the ray origin (~1/D) is dropped -- unless d/d(sources) was asked for
            bool uni = SRC != IACT_SOURCE_POINT;
            V3 sd = src;
            if (SRC == IACT_SOURCE_POINT && !(FULL && gr.sources)) {
                const V3 ac = sub_rn(v3(bnd.x, bnd.y, bnd.z), src);
                const float n2 = dot_rn(ac, ac);
                if (bnd.w * bnd.w < 1e-18f * n2 && n2 < 1e37f) { uni = true; sd = scale_rn(frsqrt_nr_rn(n2), ac); }
            }
            int n_list = 0, n_list_cyl = 0, n_rec = 0;
            if (cull) {
                const int2 cnt = fl.count ? __ldg(fl.count + f) : make_int2(-1, -1);
                if (wrec && uni && m1 - m0 > 64 && cnt.x >= 0 && cnt.y <= 32 &&
                    (SRC == IACT_SOURCE_POINT || fabsf(dot_rn(src, src) - 1.0f) < 1e-4f)) {
                    Beam beam;                                   // list and records in one pass (build_list_uni)
                    beam.c = v3(bnd.x, bnd.y, bnd.z); beam.R = bnd.w; beam.spread = 0.f; beam.u = -sd; beam.ok = true;
                    beam.invD = 0.f;
                    if (SRC == IACT_SOURCE_POINT) { const V3 ac = sub_rn(beam.c, src); beam.invD = frsqrt_fast(dot_rn(ac, ac)); }
                    n_list = build_list_uni(ob, beam, fl.ids + (size_t)f * fl.stride, cnt.x, cnt.y, list, wrec, n_list_cyl, n_rec);
                } else {
                    const Beam beam = make_beam<SRC>(bnd, src);
                    if (cnt.x >= 0) n_list = build_list(ob, beam, fl.ids + (size_t)f * fl.stride, cnt.x, cnt.y, list, n_list_cyl);
                    else            n_list = build_list(ob, beam, (const unsigned short*)nullptr, ob.n_cyl, n_obs, list, n_list_cyl);
                }
            }

            for (int m = m0 + lane; m < m1; m += 32) {
                // forward.  The lean kernel reads the packed world table of the render (world point, unit normal, 1/w:
                // the same rays, two 128-bit loads); the full one rebuilds them from the local tables it differentiates.
                V3 o, n, pl, nq, dl, nw;
                float inv_nw = 0.f, inv_w;
                if (FULL) {
                    const size_t li = ((size_t)f * M + m) * 3;
                    pl = ld3(fa.points + li); dl = ld3(fa.delta + li);
                    const V3 nl = ld3(fa.normals + li);
                    o = mul(R, pl) + pos;
                    nq = nl + scale * dl;                               // local perturbed normal
                    nw = mul(R, nq);
                    inv_nw = frsqrt_nr(dot(nw, nw));                    // <= 1 ulp, no IEEE slow path (as the forward)
                    n = inv_nw * nw;
                    inv_w = frcp_nr(__ldg(fa.weights + (size_t)f * M + m));
                } else {
                    const float4 ta = __ldg(sc.world + 2 * ((size_t)f * M + m)), tb = __ldg(sc.world + 2 * ((size_t)f * M + m) + 1);
                    o = v3(ta.x, ta.y, ta.z); n = v3(tb.x, tb.y, tb.z); inv_w = ta.w;
                }
                V3 d = sd; float inv_a = 0.f;
                if (SRC == IACT_SOURCE_POINT && !uni) { d = o - src; inv_a = frsqrt_nr(dot(d, d)); d = inv_a * d; }
                if (occluded(ob, o, -d, list, n_list_cyl, n_list, 0xffffffffu, wrec, n_rec)) continue;
                const float c = dot(d, n);
                const V3 r = d - (2.0f * c) * n;
                const float val0 = (sval * (-c)) * inv_w;
                // optical stages >= 1: forward with the state the reverse pass needs
                V3 so[IACT_MAX_STAGES], sd[IACT_MAX_STAGES];
                float st_t[IACT_MAX_STAGES], sv[IACT_MAX_STAGES];
                int smi[IACT_MAX_STAGES];
                V3 oc = o, dc = r;
                float val = val0;
                bool alive = true;
                if (STAGES) {
                    const float* rec = stage_rec;
                    for (int k = 0; k < sc.n_stages; ++k) {
                        so[k] = oc; sd[k] = dc; sv[k] = val;
                        int mi; float t;
                        if (!stage_select(sc.stages[k].n, rec, sc.stages[k].verts, !cull, oc, dc, mi, t) ||
                            occluded(ob, oc, dc, nullptr, 0, 0)) { alive = false; break; }
                        smi[k] = mi; st_t[k] = t;
                        const StageGeom g = stage_geometry(rec + (size_t)mi * STAGE_REC, oc, dc, t);
                        const V3 nw = mul(g.R, g.nl);
                        const float ck = dot(dc, nw);
                        val *= fabsf(ck);
                        oc = mul(g.R, g.pl) + g.pos;
                        dc = dc - (2.0f * ck) * nw;
                        rec += (size_t)sc.stages[k].n * STAGE_REC;
                    }
                }
                if (!alive) continue;
                const bool aa = se.axis_aligned != 0;                 // u1 = x, u2 = y, ns = z: the zero terms drop out
                const float B = aa ? dc.z : dot(dc, ns), ndoto = aa ? oc.z : dot(oc, ns);
                if (fabsf(B) < 1e-10f) continue;
                const float inv_B = frcp_nr(B);
                const float t = (se.ndotp - ndoto) * inv_B;
                if (t <= 0.f) continue;
                const V3 h = oc + t * dc - ps;
                const float x = aa ? h.x : dot(h, u1), y = aa ? h.y : dot(h, u2);
                float dval, dx, dy;
                if (!sensor_adjoint<SENS>(se, lut, G, x, y, dval, dx, dy, gr.Gn)) continue;
                // backward: sensor plane
                const float xb = val * dx, yb = val * dy;           // dL/dx, dL/dy
                V3 g_o = v3(0.f, 0.f, 0.f), g_r = g_o;
                if (xb != 0.f || yb != 0.f) {
                    const V3 g_h = aa ? v3(xb, yb, 0.f) : xb * u1 + yb * u2;
                    const float g_t = aa ? xb * dc.x + yb * dc.y : dot(g_h, dc);
                    const float gA = g_t * inv_B, gB = -g_t * t * inv_B;
                    if (FULL) {
                        g_u1 = g_u1 + xb * h; g_u2 = g_u2 + yb * h;
                        g_ps = g_ps - g_h;
                        g_ns = g_ns + gA * (ps - oc) + gB * dc;
                        g_ps = g_ps + gA * ns;
                    }
                    if (aa) { g_o = v3(xb, yb, -gA); g_r = v3(t * xb, t * yb, gB); }
                    else    { g_o = g_h - gA * ns; g_r = t * g_h + gB * ns; }
                }
                // backward: stages in reverse order
                if (STAGES) {
                    const float* rec = stage_rec;
                    for (int k = 0; k < sc.n_stages; ++k) rec += (size_t)sc.stages[k].n * STAGE_REC;
                    for (int k = sc.n_stages - 1; k >= 0; --k) {
                        rec -= (size_t)sc.stages[k].n * STAGE_REC;
                        const StageGeom g = stage_geometry(rec + (size_t)smi[k] * STAGE_REC, so[k], sd[k], st_t[k]);
                        V3 go2, gd2; float gv2;
                        if (STAGES && gr.macc) {
                            int flat = smi[k];
                            for (int q = 0; q < k; ++q) flat += sc.stages[q].n;
                            if (mreg_id < 0) mreg_id = flat;
                            if (flat == mreg_id) {
                                stage_backward(g, so[k], sd[k], st_t[k], sv[k], g_o, g_r, dval, go2, gd2, gv2, mreg);
                            } else {
                                float tmp[MACC];
                                for (int q = 0; q < MACC; ++q) tmp[q] = 0.f;
                                stage_backward(g, so[k], sd[k], st_t[k], sv[k], g_o, g_r, dval, go2, gd2, gv2, tmp);
                                for (int q = 0; q < MACC; ++q) if (tmp[q] != 0.f) atomicAdd(gr.macc + (size_t)flat * MACC + q, tmp[q]);
                            }
                        } else {
                            stage_backward(g, so[k], sd[k], st_t[k], sv[k], g_o, g_r, dval, go2, gd2, gv2, nullptr);
                        }
                        g_o = go2; g_r = gd2; dval = gv2;
                    }
                }
                // val = v (-c)/w
                float g_c = -dval * sval * inv_w;
                if (FULL) {
                    g_val += dval * (-c) * inv_w;
                    if (gr.weights) atomicAdd(gr.weights + (size_t)f * M + m, -dval * val0 * inv_w);
                }
                // r = d - 2 c n ; c = d.n
                g_c += -2.0f * dot(g_r, n);
                V3 g_d = g_r + g_c * n;
                V3 g_n = (-2.0f * c) * g_r + g_c * d;
                // d = a/|a| (point) | src (parallel)
                if (SRC == IACT_SOURCE_POINT) {
                    if (!uni) {
                        const V3 g_a = inv_a * (g_d - dot(g_d, d) * d);
                        g_o = g_o + g_a;
                        if (FULL) g_src = g_src - g_a;
                    }
                } else if (FULL) {
                    g_src = g_src + g_d;
                }
                // o = R pl + pos ; n = nw/|nw| ; nw = R (nl + scale dl)
                g_pos = g_pos + g_o;
                V3 g_nw = v3(0.f, 0.f, 0.f);
                if (FULL) {
                    g_nw = inv_nw * (g_n - dot(g_n, n) * n);
                    g_scale += dot(g_nw, mul(R, dl));
                }
                if (FULL && gr.points) {                           // per-sample adjoints (surface fits): one atomic per ray, only when asked for
                    const V3 gl = mulT(R, g_o);
                    float* gp = gr.points + ((size_t)f * M + m) * 3;
                    atomicAdd(gp, gl.x); atomicAdd(gp + 1, gl.y); atomicAdd(gp + 2, gl.z);
                }
                if (FULL && gr.nq) {
                    const V3 gl = mulT(R, g_nw);
                    float* gp = gr.nq + ((size_t)f * M + m) * 3;
                    atomicAdd(gp, gl.x); atomicAdd(gp + 1, gl.y); atomicAdd(gp + 2, gl.z);
                }
                // nw x g_nw = n x g_n (the |nw| cancels and n x n = 0): the world normal is all the rotation adjoint needs
                tau = tau + cross(o - pos, g_o) + cross(n, g_n);
            }
            // per-source adjoints
            if (per_source) {
                g_val = warp_sum(g_val); g_src.x = warp_sum(g_src.x); g_src.y = warp_sum(g_src.y); g_src.z = warp_sum(g_src.z);
                if (lane == 0) {
                    if (gr.values && g_val != 0.f) atomicAdd(gr.values + s, g_val);
                    if (gr.sources) {
                        if (g_src.x != 0.f) atomicAdd(gr.sources + 3 * s, g_src.x);
                        if (g_src.y != 0.f) atomicAdd(gr.sources + 3 * s + 1, g_src.y);
                        if (g_src.z != 0.f) atomicAdd(gr.sources + 3 * s + 2, g_src.z);
                    }
                }
            }
            __syncwarp();                                          // the list is rebuilt for the next source
        }
        // per-facet adjoints: warp reduce, one atomic per component
        float* acc = gr.facc + (size_t)f * 13;
        { const float v = warp_sum(tau.x); if (lane == 0 && v != 0.f) atomicAdd(acc + 0, v); }
        { const float v = warp_sum(tau.y); if (lane == 0 && v != 0.f) atomicAdd(acc + 1, v); }
        { const float v = warp_sum(tau.z); if (lane == 0 && v != 0.f) atomicAdd(acc + 2, v); }
        { const float v = warp_sum(g_pos.x); if (lane == 0 && v != 0.f) atomicAdd(acc + 9, v); }
        { const float v = warp_sum(g_pos.y); if (lane == 0 && v != 0.f) atomicAdd(acc + 10, v); }
        { const float v = warp_sum(g_pos.z); if (lane == 0 && v != 0.f) atomicAdd(acc + 11, v); }
        if (FULL) { const float v = warp_sum(g_scale); if (lane == 0 && v != 0.f) atomicAdd(acc + 12, v); }
        if (STAGES && gr.macc) {
            // lanes may have cached different mirrors: reduce per distinct id
            unsigned todo = __ballot_sync(0xffffffffu, mreg_id >= 0);
            while (todo) {
                const int id = __shfl_sync(0xffffffffu, mreg_id, __ffs(todo) - 1);
                const bool mine = mreg_id == id;
#pragma unroll
                for (int q = 0; q < MACC; ++q) { const float v = warp_sum(mine ? mreg[q] : 0.f); if (lane == 0 && v != 0.f) atomicAdd(gr.macc + (size_t)id * MACC + q, v); }
                todo &= ~__ballot_sync(0xffffffffu, mine);
            }
        }
    }
    if (!FULL) return;
    // sensor adjoints: warp reduce then one atomic per warp and component
    const float sv[12] = {g_ps.x, g_ps.y, g_ps.z,
                          g_u1.x, g_u2.x, g_ns.x, g_u1.y, g_u2.y, g_ns.y, g_u1.z, g_u2.z, g_ns.z};   // dL/dR_s row-major
#pragma unroll
    for (int k = 0; k < 12; ++k) { const float v = warp_sum(sv[k]); if (lane == 0 && v != 0.f) atomicAdd(gr.sacc + k, v); }
}

// dL/d(euler degrees) from dL/dR for R = Rz(rot) Ry(tilt) Rx(tip)  (transforms.py:72-106)
__device__ __forceinline__ void euler_adjoint(float tip, float tilt, float rot, const float* gR, float* out3) {
    const float D2R = 0.017453292519943295f;
    float sx, cx, sy, cy, sz, cz;
    sincosf(tip * D2R, &sx, &cx); sincosf(tilt * D2R, &sy, &cy); sincosf(rot * D2R, &sz, &cz);
    const float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, dRx[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx};
    const float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, dRy[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy};
    const float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1}, dRz[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
    auto mm = [](const float* A, const float* B, float* C) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    };
    auto inner = [](const float* A, const float* B) { float s = 0.f; for (int i = 0; i < 9; ++i) s += A[i] * B[i]; return s; };
    float T[9], U[9];
    mm(Ry, dRx, T); mm(Rz, T, U); out3[0] = inner(gR, U) * D2R;
    mm(dRy, Rx, T); mm(Rz, T, U); out3[1] = inner(gR, U) * D2R;
    mm(Ry, Rx, T);  mm(dRz, T, U); out3[2] = inner(gR, U) * D2R;
}

struct StageList { int n_stages; StageDev st[IACT_MAX_STAGES]; };

__global__ void vjp_finalize_kernel(IactFacets fa, const float* __restrict__ facc, const float* __restrict__ sacc,
                                    const float* __restrict__ macc, StageList sl, IactGrads out, float3 sensor_euler) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (macc && f == 0) {
        int flat = 0;
        for (int k = 0; k < sl.n_stages; ++k)
            for (int i = 0; i < sl.st[k].n; ++i, ++flat) {
                const float* r = sl.st[k].rec + (size_t)i * IACT_MIRROR_REC;
                const float* a = macc + (size_t)flat * MACC;
                if (out.stage_rotations) {
                    float e[3];
                    euler_adjoint(r[3], r[4], r[5], a, e);
                    for (int q = 0; q < 3; ++q) out.stage_rotations[3 * flat + q] += e[q];
                }
                if (out.stage_positions) for (int q = 0; q < 3; ++q) out.stage_positions[3 * flat + q] += a[9 + q];
                if (out.stage_surface) for (int q = 0; q < 4; ++q) out.stage_surface[4 * flat + q] += a[12 + q];
            }
    }
    if (f < fa.n_facets) {
        const float* a = facc + (size_t)f * 13;
        if (out.rotations) {
            // R = Rz(rot) Ry(tilt) Rx(tip) (transforms.py:72-106): axes a_tip = Rz Ry x, a_tilt = Rz y, a_rot = z; per degree
            const float D2R = 0.017453292519943295f;
            float sy, cy, sz, cz;
            sincosf(fa.rotations[3 * f + 1] * D2R, &sy, &cy); sincosf(fa.rotations[3 * f + 2] * D2R, &sz, &cz);
            out.rotations[3 * f + 0] += D2R * (cz * cy * a[0] + sz * cy * a[1] - sy * a[2]);
            out.rotations[3 * f + 1] += D2R * (-sz * a[0] + cz * a[1]);
            out.rotations[3 * f + 2] += D2R * a[2];
        }
        if (out.positions) for (int k = 0; k < 3; ++k) out.positions[3 * f + k] += a[9 + k];
        if (out.scale) out.scale[f] += a[12];
    }
    if (f == 0) {
        if (out.sensor_position) for (int k = 0; k < 3; ++k) out.sensor_position[k] += sacc[k];
        if (out.sensor_euler) {
            float e[3];
            euler_adjoint(sensor_euler.x, sensor_euler.y, sensor_euler.z, sacc + 3, e);
            for (int k = 0; k < 3; ++k) out.sensor_euler[k] += e[k];
        }
    }
}

}  // namespace

extern "C" int iact_render_vjp(const IactScene* scene, const IactFacets* facets, const float* sources, const float* values,
                               int n_sources, int source_type, const float* cotangent, const IactGrads* grads, void* stream) {
    SceneDev d;
    int rc = fill_scene(scene, d);
    if (rc) return rc;
    IACT_REQUIRE(facets && grads && cotangent, "null pointer");
    IACT_REQUIRE(facets->n_facets == d.F && facets->n_samples == d.M, "facet tables do not match the scene");
    IACT_REQUIRE(source_type == IACT_SOURCE_POINT || source_type == IACT_SOURCE_PARALLEL, "bad source_type");
    const int S = n_sources;
    if (S <= 0 || d.F == 0 || d.M == 0) return IACT_OK;
    IACT_REQUIRE(sources && values, "null sources/values");
    IACT_REQUIRE(facets->positions && facets->rotations && facets->scale && facets->points && facets->normals && facets->delta && facets->weights,
                 "null facet table");
    cudaStream_t st = (cudaStream_t)stream;
    const bool hex = d.sens.kind == IACT_SENSOR_HEX || d.sens.kind == IACT_SENSOR_SOFT_HEX;
    Scratch cull_scr, acc_scr;
    FacetLists fl;
    fl.ids = nullptr; fl.count = nullptr; fl.stride = 0; fl.counter = nullptr;
    if (d.cull && S >= 4) { rc = run_facet_cull(d, sources, S, source_type, cull_scr, fl, st); if (rc) return rc; }
    int n2 = 0;
    for (int k = 0; k < d.n_stages; ++k) n2 += d.stages[k].n;
    const bool want_stage = n2 > 0 && (grads->stage_positions || grads->stage_rotations || grads->stage_surface);
    // scratch: the work-queue counter (8 bytes) followed by the adjoint accumulators
    const size_t acc_floats = 2 + (size_t)d.F * 13 + 12 + (want_stage ? (size_t)n2 * MACC : 0);
    rc = acc_scr.alloc(acc_floats * sizeof(float), st);
    if (rc) return rc;
    IACT_CUDA(cudaMemsetAsync(acc_scr.ptr, 0, acc_floats * sizeof(float), st));
    GradsDev gr;
    gr.weights = grads->weights; gr.values = grads->values; gr.sources = grads->sources;
    gr.facc = reinterpret_cast<float*>(acc_scr.ptr) + 2;
    gr.sacc = gr.facc + (size_t)d.F * 13;
    gr.macc = want_stage ? gr.sacc + 12 : nullptr;
    gr.points = grads->points; gr.nq = grads->nq;
    gr.Gn = nullptr;
    Scratch gn_scr;
    if (d.sens.kind == IACT_SENSOR_SOFT_HEX && d.sens.ksize == 1) {
        rc = gn_scr.alloc((size_t)d.sens.npix * 8 * sizeof(float), st);
        if (rc) return rc;
        const int cells = d.sens.tq * d.sens.tr;
        gather_cotangent_kernel<<<(cells + 255) / 256, 256, 0, st>>>(d.sens, cotangent, reinterpret_cast<float*>(gn_scr.ptr));
        iact_count_launch();
        IACT_CUDA(cudaGetLastError());
        gr.Gn = reinterpret_cast<const float*>(gn_scr.ptr);
    }

    const int threads = 256;
    size_t smem = (size_t)obstruction_floats(d.n_cyl, d.n_box, d.n_sph, d.n_obox, d.n_tri, d.cull != 0) * 4 + 16;
    if (hex) smem += (size_t)((d.sens.tq * d.sens.tr + 1) / 2) * 4;
    smem += (size_t)stage_floats(d) * 4;
    if (d.cull) smem += (size_t)(threads / 32) * ((d.n_cyl + d.n_box + d.n_sph + d.n_obox + d.n_tri + 1) & ~1) * 2;
    if (IACT_CYL_RECORDS && d.cull && d.n_cyl > 0) smem += (size_t)(threads / 32) * CYL_REC_MAX * CYL_REC * 4;
    if (smem > 200 * 1024) { iact_set_error("scene needs %zu bytes of shared memory per block (limit 204800)", smem); return IACT_ERR_UNSUPPORTED; }
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024) IACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        IACT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        const long long max_blocks = (long long)sm_count() * std::max(occ, 1);
        VjpPlan vp = make_vjp_plan(d, S, max_blocks * (threads / 32));
        vp.counter = reinterpret_cast<unsigned long long*>(acc_scr.ptr);
        const long long blocks = (vp.n_units + threads / 32 - 1) / (threads / 32);
        const unsigned grid = (unsigned)std::max(1LL, std::min(blocks, max_blocks));
        kern<<<grid, threads, smem, st>>>(d, *facets, sources, values, vp, fl, cotangent, gr);
        iact_count_launch();
        return iact_check_cuda(cudaGetLastError(), "vjp_kernel launch");
    };
    const bool st2 = d.n_stages > 0;
    // lean instantiation: nothing but the facet poses wanted
    const bool full = grads->points || grads->nq || grads->scale || grads->weights || grads->values || grads->sources || grads->sensor_position ||
                      grads->sensor_euler || want_stage;
#define IACT_VJP_PICK(SRC_, SENS_)                                                                              \
    (st2 ? (full ? launch(vjp_kernel<SRC_, SENS_, true, true>) : launch(vjp_kernel<SRC_, SENS_, true, false>))  \
         : (full ? launch(vjp_kernel<SRC_, SENS_, false, true>) : launch(vjp_kernel<SRC_, SENS_, false, false>)))
    if (source_type == IACT_SOURCE_POINT) rc = hex ? IACT_VJP_PICK(IACT_SOURCE_POINT, SENS_HEX) : IACT_VJP_PICK(IACT_SOURCE_POINT, SENS_SQUARE);
    else                                  rc = hex ? IACT_VJP_PICK(IACT_SOURCE_PARALLEL, SENS_HEX) : IACT_VJP_PICK(IACT_SOURCE_PARALLEL, SENS_SQUARE);
#undef IACT_VJP_PICK
    if (rc) return rc;
    const float3 se = make_float3(scene->sensor.euler[0], scene->sensor.euler[1], scene->sensor.euler[2]);
    StageList sl;
    sl.n_stages = d.n_stages;
    for (int k = 0; k < IACT_MAX_STAGES; ++k) sl.st[k] = d.stages[k];
    vjp_finalize_kernel<<<(d.F + 127) / 128, 128, 0, st>>>(*facets, gr.facc, gr.sacc, gr.macc, sl, *grads, se);
    iact_count_launch();
    return iact_check_cuda(cudaGetLastError(), "vjp_finalize_kernel launch");
}

extern "C" int iact_work_plan(int kind, int n_facets, int n_samples, int n_sources, int has_obstructions, long long resident_warps,
                              int* out6, long long* n_units) {
    IACT_REQUIRE(out6 && n_units, "null output");
    IACT_REQUIRE(n_facets > 0 && n_samples > 0 && n_sources > 0 && resident_warps > 0, "sizes must be positive");
    IACT_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (render) or 1 (vjp)");
    SceneDev d;
    memset(&d, 0, sizeof(d));
    d.F = n_facets; d.M = n_samples; d.cull = has_obstructions != 0;
    if (kind == 0) {
        const QueuePlan q = make_queue_plan(d, n_sources, resident_warps);
        out6[0] = q.facets_per_unit; out6[1] = q.runs; out6[2] = q.msplit; out6[3] = q.msize; out6[4] = out6[5] = 0;
        *n_units = q.n_units;
    } else {
        const VjpPlan p = make_vjp_plan(d, n_sources, resident_warps);
        out6[0] = p.slen; out6[1] = p.sruns; out6[2] = p.msplit; out6[3] = p.msize; out6[4] = out6[5] = 0;
        *n_units = p.n_units;
    }
    return IACT_OK;
}
