// Shared between the forward (iact_render.cu) and backward (iact_vjp.cu) kernels: work plan,
// conservative hierarchical obstruction culling, scene packing on the host.
#pragma once
#include "iact_trace.cuh"
#include <algorithm>
#include <cstring>

namespace {

enum { MODE_RENDER = 0, MODE_MATRIX = 1, MODE_DEBUG = 2 };
enum { SENS_SQUARE = 0, SENS_HEX = 1, SENS_SOFT_HEX = 2 };   // hard/soft square share one instantiation

struct LaunchPlan {
    int S, n_chunks, chunk_facets, msplit, msize;
    long long n_items;
};

// Work queue of the forward kernels: a global counter hands out units.  Render / debug: a unit is
// (source, run of `facets_per_unit` facets, sample part) and is pulled by one warp; response matrix: a unit
// is a block item of the LaunchPlan.  counter == nullptr = static grid-stride split.
struct QueuePlan { int facets_per_unit, runs, msplit, msize; long long n_units; unsigned long long* counter; };

// Level-1 culling output: per facet a list of primitive ids (cylinders first) and its two counts;
// count.x < 0 means "no facet-level culling for this facet" (degenerate beam): use every primitive.
struct FacetLists { const unsigned short* ids; const int2* count; int stride;
                    unsigned long long* counter; };   // work-queue counter zeroed by facet_cull_kernel (nullptr: none)

struct Beam { V3 c, u; float R, invD, spread; bool ok; };

// Beam of all rays from the facet's bounding sphere towards one source (and beyond: the reference's
// shadow ray is infinite, render.py:138 + :40).
template <int SRC>
__device__ __forceinline__ Beam make_beam(float4 bnd, V3 src) {
    Beam b;
    b.c = v3(bnd.x, bnd.y, bnd.z); b.R = bnd.w; b.spread = 0.f;
    V3 a = SRC == IACT_SOURCE_POINT ? src - b.c : -src;
    const float n2 = dot(a, a);
    b.ok = n2 > 1e-30f && n2 < 1e37f;
    const float inv = rsqrtf(b.ok ? n2 : 1.f);
    b.u = inv * a;
    b.invD = SRC == IACT_SOURCE_POINT ? inv : 0.f;
    if (b.R * b.invD > 0.1f) b.ok = false;              // source inside/near the facet: no culling
    return b;
}

// Conservative: false only if no ray of the beam can come within r of the segment [p1,p2].
// A ray starts within R of c and its direction is within `spread` (chord) + 1.5708 R/D (point-source
// parallax) of u; rays are half-lines, so everything behind the facet is out of reach.
__device__ __forceinline__ bool beam_keeps_capsule(const Beam& b, V3 p1, V3 p2, float r) {
    const V3 a1 = p1 - b.c, a2 = p2 - b.c;
    const float t1 = dot(a1, b.u), t2 = dot(a2, b.u);
    const float tmx = fmaxf(t1, t2);
    const float marg = 2e-3f;
    if (tmx + r < -(b.R + marg)) return false;          // wholly behind every ray origin
    const float tmax = 1.1f * (fmaxf(tmx, 0.f) + r + b.R);       // 1/cos(max beam half-angle 0.31 rad) < 1.1
    const float Reff = b.R * (1.0f + 1.5708f * tmax * b.invD) + b.spread * tmax + marg + 1e-5f * tmax;
    const V3 q1 = a1 - t1 * b.u, q2 = a2 - t2 * b.u;
    const V3 e = q2 - q1;
    const float ee = dot(e, e);
    const float s = ee > 1e-20f ? fminf(fmaxf(-dot(q1, e) * frcp_fast(ee), 0.f), 1.f) : 0.f;
    const V3 dv = q1 + s * e;
    const float lim = Reff + r;
    return dot(dv, dv) <= lim * lim;
}

__device__ __forceinline__ bool keep_primitive(const ObsSmem& ob, const Beam& b, int id) {
    const float* c = ob.cprox; const int n = ob.n_cyl + ob.n_rest;
    return beam_keeps_capsule(b, v3(c[id], c[n + id], c[2 * n + id]), v3(c[3 * n + id], c[4 * n + id], c[5 * n + id]), c[6 * n + id]);
}

// Level-3 strip test (far or parallel sources, parallax R/D < 1e-7): the rays of a (facet, source) item are
// parallel to the beam axis u, so a ray can touch cylinder e only if its origin lies within r_e of the plane
// through the cylinder axis that contains u: |n_e.(o - p1_e)| <= r_e with n_e = unit(u x axis_e).  A run of 32
// table rows with bounding sphere (c, R) therefore needs cylinder e only if |n_e.c - n_e.p1_e| <= R + r_e + margins
// (2 mm + 1e-5 t for float32 rounding plus the parallax term, as in beam_keeps_capsule; the residual tilt of a
// ray against u, <= 2e-7, moves it by < 0.1 mm over the 50 m of a telescope).  Cylinders within 3 deg of the
// beam axis (n_e ill-conditioned) and the other primitive types are always kept.
// strip_masks: lane j returns the keep mask (bit e = list entry e) of run run0 + j; the candidates' records are
// computed by lane e and broadcast by shuffles, about 10 instructions per candidate for 32 runs.
__device__ __forceinline__ bool strip_applies(const Beam& b) { return b.ok && b.R * b.invD < 1e-7f; }
__device__ __forceinline__ unsigned strip_masks(const ObsSmem& ob, const Beam& beam, const unsigned short* list, int n_list_cyl,
                                                int n_list, const float4* __restrict__ cbs, int n_runs, int run0) {
    const int lane = threadIdx.x & 31;
    V3 sn = v3(0.f, 0.f, 0.f);
    float sk = 0.f, srr = INFINITY;                                         // INFINITY = always kept
    if (lane < n_list_cyl) {
        const float* c = ob.cprox; const int n = ob.n_cyl + ob.n_rest; const int id = list[lane];
        const V3 p1 = v3(c[id], c[n + id], c[2 * n + id]), p2 = v3(c[3 * n + id], c[4 * n + id], c[5 * n + id]);
        const float r = c[6 * n + id];
        const V3 ax = p2 - p1;
        const V3 w = cross(beam.u, ax);
        const float w2 = dot(w, w), a2 = dot(ax, ax);
        if (w2 >= 2.5e-3f * a2 && a2 > 1e-20f) {
            sn = rsqrtf(w2) * w;
            sk = dot(sn, p1);
            const float tfar = 1.1f * (fmaxf(fmaxf(dot(p1 - beam.c, beam.u), dot(p2 - beam.c, beam.u)), 0.f) + r + beam.R);
            srr = r + 2e-3f + 1e-5f * tfar + beam.R * 1.5708f * tfar * beam.invD;
        }
    }
    const float4 cb = __ldg(cbs + 2 * min(run0 + lane, n_runs - 1));
    unsigned mask = 0u;
    for (int e = 0; e < n_list; ++e) {
        const float nx = __shfl_sync(0xffffffffu, sn.x, e), ny = __shfl_sync(0xffffffffu, sn.y, e), nz = __shfl_sync(0xffffffffu, sn.z, e);
        const float k = __shfl_sync(0xffffffffu, sk, e), rr = __shfl_sync(0xffffffffu, srr, e);
        const float sd = nx * cb.x + ny * cb.y + nz * cb.z - k;
        if (!(fabsf(sd) > cb.w + rr)) mask |= 1u << e;
    }
    return mask;
}

// Warp-cooperative compaction of the primitives a beam can reach.  `cand` (may be null = all
// primitives) lists candidate ids, cylinders first (n_cand_cyl of n_cand).  Writes ids to `out`
// (shared or global), returns the total and sets n_cyl_out.  Order is preserved.
__device__ __forceinline__ int build_list(const ObsSmem& ob, const Beam& b, const unsigned short* cand, int n_cand_cyl,
                                          int n_cand, unsigned short* out, int& n_cyl_out) {
    const unsigned lane = threadIdx.x & 31u;
    int n = 0, ncyl = 0;
    for (int base = 0; base < n_cand; base += 32) {
        const int i = base + (int)lane;
        bool keep = false;
        int id = 0;
        if (i < n_cand) {
            id = cand ? (int)cand[i] : i;
            keep = !b.ok || keep_primitive(ob, b, id);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) out[n + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)id;
        n += __popc(mask);
        // cylinders come first in `cand`: count those kept among positions < n_cand_cyl
        const int lim = n_cand_cyl - base;
        ncyl += lim >= 32 ? __popc(mask) : (lim > 0 ? __popc(mask & ((1u << lim) - 1u)) : 0);
    }
    n_cyl_out = ncyl;
    __syncwarp();
    return n;
}

// Level-2 list AND cylinder records of an item whose rays share the unit direction -u (far point / parallel sources)
// in one pass: lane e takes level-1 candidate e (n_cand <= 32, cylinders first), evaluates the direction half of the
// cylinder test (cyl_dir), which the record needs anyway, and runs the capsule test of beam_keeps_capsule on top of it:
// with p2 = p1 + h ax the segment in the plane normal to u is q1 + s h (ax - (ax.u) u), of squared length h^2 a.  Same
// margins as beam_keeps_capsule (the axis of the staged table differs from (p2 - p1) / h by float32 rounding, ~1e-7 h,
// against 2 mm).  Kept cylinders write their record at their compacted position; other primitives go through the
// proxy test.  Returns the list length, sets n_cyl_out and n_rec_out = how many leading list entries the ray loop may
// test through their records (at most CYL_REC_MAX, and none from the first literal-form cylinder on; the rest inline).
__device__ __forceinline__ int build_list_uni(const ObsSmem& ob, const Beam& b, const unsigned short* __restrict__ cand, int n_cand_cyl,
                                              int n_cand, unsigned short* out, float* wrec, int& n_cyl_out, int& n_rec_out) {
    const unsigned lane = threadIdx.x & 31u;
    bool keep = false;
    int id = 0;
    CylDir cd;
    const float* c = nullptr;
    if ((int)lane < n_cand) {
        id = (int)cand[lane];
        if ((int)lane < n_cand_cyl) {
            c = ob.cyl + CYL_STRIDE * id;
            const V3 p1 = v3(c[0], c[1], c[2]), ax = v3(c[3], c[4], c[5]);
            const float h = c[6], r = fabsf(c[7]) * 1.0001f;
            cd = cyl_dir(ax, __fmul_rn(c[7], c[7]), b.u);
            const V3 a1 = p1 - b.c;
            const float t1 = dot(a1, b.u), t2 = t1 + h * cd.rd_ax;
            const float tmx = fmaxf(t1, t2);
            const float marg = 2e-3f;
            if (!(tmx + r < -(b.R + marg))) {                       // else: wholly behind every ray origin
                const float tmax = 1.1f * (fmaxf(tmx, 0.f) + r + b.R);
                const float Reff = b.R * (1.0f + 1.5708f * tmax * b.invD) + marg + 1e-5f * tmax;
                const V3 q1 = a1 - t1 * b.u;
                const V3 e = h * (ax - cd.rd_ax * b.u);
                const float ee = h * h * cd.a;
                const float sp = ee > 1e-20f ? fminf(fmaxf(-dot(q1, e) * frcp_fast(ee), 0.f), 1.f) : 0.f;
                const V3 dv = q1 + sp * e;
                const float lim = Reff + r;
                keep = dot(dv, dv) <= lim * lim;
            }
        } else {
            keep = keep_primitive(ob, b, id);
        }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int pos = __popc(mask & ((1u << lane) - 1u));
    if (keep) {
        out[pos] = (unsigned short)id;
        if (c && pos < CYL_REC_MAX) {
            cyl_record_store(wrec + CYL_REC * pos, c, cd);
        }
    }
    n_cyl_out = n_cand_cyl >= 32 ? __popc(mask) : __popc(mask & ((1u << n_cand_cyl) - 1u));
    // records hold the interval form only: they end before the first kept cylinder the direction is nearly parallel to
    const unsigned lit = __ballot_sync(0xffffffffu, keep && c && !cyl_interval_form(cd));
    const int first_lit = lit ? __popc(mask & ((1u << (__ffs(lit) - 1)) - 1u)) : CYL_REC_MAX;
    n_rec_out = min(min(n_cyl_out, CYL_REC_MAX), first_lit);
    __syncwarp();
    return __popc(mask);
}

// _check_occlusions for the leg towards an optical stage >= 1 (render.py:76), culled per warp from the
// rays themselves: origins lie within R of the leader lane's origin and directions within `spread`
// (chord) of the leader's, so a primitive that beam cannot reach is skipped for all 32 rays.  Exact for
// the same reason as the beam culling of the incoming leg; wide bundles (unsorted samples on a strongly
// curved facet) fall back to testing everything.  `need` = this lane's result matters (rays already at
// value 0 stay 0 whatever the test says).  Must be called by all 32 lanes; needs the culling proxies.
__device__ __forceinline__ bool occluded_leg_culled(const ObsSmem& ob, V3 o, V3 d, bool need) {
    const unsigned am = __ballot_sync(0xffffffffu, need);
    if (am == 0u) return false;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(am) - 1;
    Beam b;
    b.c = v3(__shfl_sync(0xffffffffu, o.x, leader), __shfl_sync(0xffffffffu, o.y, leader), __shfl_sync(0xffffffffu, o.z, leader));
    b.u = v3(__shfl_sync(0xffffffffu, d.x, leader), __shfl_sync(0xffffffffu, d.y, leader), __shfl_sync(0xffffffffu, d.z, leader));
    const V3 eo = o - b.c, ed = d - b.u;
    // non-negative floats order like their bit patterns; a NaN has the largest pattern and switches culling off
    const float r2 = __uint_as_float(__reduce_max_sync(0xffffffffu, need ? __float_as_uint(dot(eo, eo)) : 0u));
    const float s2 = __uint_as_float(__reduce_max_sync(0xffffffffu, need ? __float_as_uint(dot(ed, ed)) : 0u));
    b.R = sqrtf(r2) * 1.001f + 1e-6f;
    b.spread = sqrtf(s2) * 1.001f + 1e-6f;
    b.invD = 0.f;
    const float u2 = dot(b.u, b.u);
    b.ok = (b.spread < 0.15f) && (b.R < 1e6f) && (u2 > 0.99f) && (u2 < 1.01f);
    const int n_obs = ob.n_cyl + ob.n_rest;
    bool blocked = false;
    if (!b.ok) {
        if (need) blocked = occluded(ob, o, d, nullptr, 0, 0);
        return blocked;
    }
    for (int base = 0; base < n_obs; base += 32) {
        const int id = base + (int)lane;
        unsigned mask = __ballot_sync(0xffffffffu, id < n_obs && keep_primitive(ob, b, id));
        while (mask) {
            const int e = __ffs(mask) - 1;
            mask &= mask - 1u;
            if (need) blocked |= hit_primitive(ob, base + e, o, d);
        }
    }
    return blocked;
}

// The same leg culled per 32-row RUN of a binned table instead of per iteration from the rays (first optical stage
// >= 1 only, at most 32 primitives): lane j bounds the reflected rays of run run0 + j without tracing them.  The rows
// of a run start within R of c (bounding sphere) and their normals lie within e of the unit mean normal nb (normal
// cone, transform_binned_kernel).  With the incoming direction d (shared by the item, or towards a point source:
// then it varies by at most dd = 1.5708 R / D over the run), the reflected direction r(d, n) = d - 2 (d.n) n obeys
//   |r(d, n) - r(d0, nb)| <= |I - 2 n n^T| |d - d0| + 2 |d0| |n - nb| (2 |nb| + |n - nb|)
//                         <= dd (1 + 2 (|nb| + e)^2) + 2 |d0| e (2 |nb| + e),
// which is the `spread` of a Beam about u = r(d0, nb): the same conservative capsule test as everywhere else then
// yields, per run, the mask (bit p = primitive p) of what the leg can reach.  About 60 instructions per primitive for
// 32 runs, against about 150 per iteration for the dynamic version.  A degenerate run keeps everything.
template <int SRC>
__device__ __forceinline__ unsigned leg_masks(const ObsSmem& ob, V3 src, bool uni, V3 sd, const float4* __restrict__ cbs,
                                              int n_runs, int run0) {
    const int lane = threadIdx.x & 31;
    const int run = min(run0 + lane, n_runs - 1);
    const float4 cb = __ldg(cbs + 2 * run), cn = __ldg(cbs + 2 * run + 1);
    Beam b;
    b.c = v3(cb.x, cb.y, cb.z); b.R = cb.w; b.invD = 0.f;
    V3 d = sd;
    float dd = 0.f;
    bool ok = true;
    if (SRC == IACT_SOURCE_POINT && !uni) {
        const V3 a = b.c - src;
        const float n2 = dot(a, a);
        ok = n2 > 1e-30f && n2 < 1e37f;
        const float inv = rsqrtf(ok ? n2 : 1.f);
        d = inv * a;
        dd = 1.5708f * b.R * inv * 1.001f;
        ok = ok && b.R * inv < 0.1f;
    }
    const V3 nb = v3(cn.x, cn.y, cn.z);
    const float e = cn.w, dl = sqrtf(dot(d, d)), nl = sqrtf(dot(nb, nb)), nmax = nl + e;
    b.u = d - (2.0f * dot(d, nb)) * nb;
    b.spread = (dd * (1.0f + 2.0f * nmax * nmax) + 2.0f * dl * e * (2.0f * nl + e)) * 1.001f + 1e-6f;
    const float u2 = dot(b.u, b.u);
    ok = ok && (b.spread < 0.15f) && (b.R < 1e6f) && (u2 > 0.99f) && (u2 < 1.01f);
    const int n_obs = ob.n_cyl + ob.n_rest;
    if (!ok) return n_obs >= 32 ? 0xffffffffu : (1u << n_obs) - 1u;
    b.ok = true;
    unsigned mask = 0u;
    for (int p = 0; p < n_obs; ++p)
        if (keep_primitive(ob, b, p)) mask |= 1u << p;
    return mask;
}

// ---------------------------------------------------------------- level-1 culling: facet x all sources
// One warp per facet: bounding cone of the directions towards all sources, then one pass over the
// primitives.  Writes ids[f*stride ..] and count[f] = (n_cyl_kept, n_total_kept) or (-1,-1).
// Also zeroes the work-queue counter of the trace kernel that follows and (render on a hex camera) the output image:
// two memset launches less per render, which matters for the 0.5 ms jobs of an 8-way source split.
template <int SRC>
__global__ void __launch_bounds__(256) facet_cull_kernel(const __grid_constant__ SceneDev sc, const float* __restrict__ sources,
                                                         int S, unsigned short* __restrict__ ids, int2* __restrict__ count, int stride,
                                                         unsigned long long* __restrict__ counter, float* __restrict__ zero, size_t n_zero) {
    extern __shared__ __align__(16) float smem[];
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ull;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += (size_t)gridDim.x * blockDim.x) zero[i] = 0.f;
    ObsSmem ob;
    stage_obstructions(sc, smem, ob, true);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int n_obs = ob.n_cyl + ob.n_rest;
    for (int f = blockIdx.x * nwarps + warp; f < sc.F; f += gridDim.x * nwarps) {
        const float4 bnd = __ldg(sc.bounds + f);
        const V3 c = v3(bnd.x, bnd.y, bnd.z);
        // pass 1: mean unit direction, largest 1/D, degeneracy flag
        V3 sum = v3(0.f, 0.f, 0.f);
        float invDmax = 0.f;
        bool bad = false;
        for (int s = lane; s < S; s += 32) {
            const V3 src = v3(__ldg(sources + 3 * s), __ldg(sources + 3 * s + 1), __ldg(sources + 3 * s + 2));
            const Beam b = make_beam<SRC>(bnd, src);
            bad = bad || !b.ok;
            sum = sum + b.u;
            invDmax = fmaxf(invDmax, b.invD);
        }
        for (int o = 16; o > 0; o >>= 1) {
            sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
            sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
            sum.z += __shfl_xor_sync(0xffffffffu, sum.z, o);
            invDmax = fmaxf(invDmax, __shfl_xor_sync(0xffffffffu, invDmax, o));
        }
        bad = __any_sync(0xffffffffu, bad);
        const float n2 = dot(sum, sum);
        bad = bad || !(n2 > 1e-12f);
        Beam fb;
        fb.c = c; fb.R = bnd.w; fb.invD = invDmax; fb.u = rsqrtf(bad ? 1.f : n2) * sum;
        // pass 2: largest chord between any source direction and the mean direction
        float dmax2 = 0.f;
        if (!bad) {
            for (int s = lane; s < S; s += 32) {
                const V3 src = v3(__ldg(sources + 3 * s), __ldg(sources + 3 * s + 1), __ldg(sources + 3 * s + 2));
                const Beam b = make_beam<SRC>(bnd, src);
                const V3 dd = b.u - fb.u;
                dmax2 = fmaxf(dmax2, dot(dd, dd));
            }
            for (int o = 16; o > 0; o >>= 1) dmax2 = fmaxf(dmax2, __shfl_xor_sync(0xffffffffu, dmax2, o));
        }
        fb.spread = sqrtf(dmax2) * 1.001f + 1e-6f;
        fb.ok = !bad && fb.spread < 0.15f;
        if (!fb.ok) { if (lane == 0) count[f] = make_int2(-1, -1); continue; }
        int ncyl = 0;
        const int n = build_list(ob, fb, (const unsigned short*)nullptr, ob.n_cyl, n_obs, ids + (size_t)f * stride, ncyl);
        if (lane == 0) count[f] = make_int2(ncyl, n);
    }
}

// ---------------------------------------------------------------- host side
static int g_sm_count = 0;

int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_sm_count = 148;
    }
    return g_sm_count;
}

void fill_sensor(const IactSensor& s, SensDev& d) {
    memset(&d, 0, sizeof(d));
    d.kind = s.kind;
    // euler_to_matrix (transforms.py:72-106) in float32
    const float D2R = 0.017453292519943295f;
    const float rx = s.euler[0] * D2R, ry = s.euler[1] * D2R, rz = s.euler[2] * D2R;
    const float cx = cosf(rx), sx = sinf(rx), cy = cosf(ry), sy = sinf(ry), cz = cosf(rz), sz = sinf(rz);
    const float a[3][3] = {{cy, sy * sx, sy * cx}, {0.f, cx, -sx}, {-sy, cy * sx, cy * cx}};
    float R[3][3];
    for (int j = 0; j < 3; ++j) { R[0][j] = cz * a[0][j] - sz * a[1][j]; R[1][j] = sz * a[0][j] + cz * a[1][j]; R[2][j] = a[2][j]; }
    for (int i = 0; i < 3; ++i) { d.pos[i] = s.position[i]; d.u1[i] = R[i][0]; d.u2[i] = R[i][1]; d.nrm[i] = R[i][2]; }
    d.ndotp = d.nrm[0] * d.pos[0] + d.nrm[1] * d.pos[1] + d.nrm[2] * d.pos[2];
    d.axis_aligned = d.u1[0] == 1.f && d.u1[1] == 0.f && d.u1[2] == 0.f && d.u2[0] == 0.f && d.u2[1] == 1.f && d.u2[2] == 0.f &&
                     d.nrm[0] == 0.f && d.nrm[1] == 0.f && d.nrm[2] == 1.f;
    d.W = s.width; d.H = s.height;
    d.x0 = (float)s.x0; d.y0 = (float)s.y0; d.dx = (float)s.dx; d.dy = (float)s.dy; d.edge = (float)s.edge_width;
    d.inv_dx = s.dx != 0.0 ? (float)(1.0 / s.dx) : 0.f; d.inv_dy = s.dy != 0.0 ? (float)(1.0 / s.dy) : 0.f;
    d.goffx = (float)s.grid_offset[0]; d.goffy = (float)s.grid_offset[1];
    const float ang = (float)(-s.grid_rotation);
    d.cr = cosf(ang); d.sr = sinf(ang);
    d.size = (float)s.hex_size; d.size_sqrt3 = (float)(s.hex_size * 1.7320508075688772);
    d.size_1p5 = (float)(s.hex_size * 1.5); d.inradius = (float)s.hex_inradius;
    d.edge_thr = s.hex_inradius != 0.0 ? (float)(1.0 - s.edge_width / s.hex_inradius) : 1.0f;
    if (s.hex_size != 0.0) {
        d.ax_qx = (float)(0.5773502691896257 / s.hex_size); d.ax_qy = (float)(1.0 / (3.0 * s.hex_size));
        d.ax_ry = (float)(2.0 / (3.0 * s.hex_size));
    }
    d.inv_inradius = s.hex_inradius != 0.0 ? (float)(1.0 / s.hex_inradius) : 0.f;
    d.r_out2 = s.hex_outer_radius > 0.0 ? (float)(s.hex_outer_radius * s.hex_outer_radius) : INFINITY;
    d.qmin = s.q_min; d.rmin = s.r_min; d.tq = s.table_q; d.tr = s.table_r; d.npix = s.n_pixels;
    d.lookup = s.lookup; d.sigma = (float)s.sigma; d.ksize = s.kernel_size;
    d.soft_nk = (s.sigma > 0.0 && s.hex_inradius > 0.0) ? (float)(-0.72134752044448170368 / (s.hex_inradius * s.hex_inradius * s.sigma * s.sigma)) : 0.f;
}

int fill_scene(const IactScene* s, SceneDev& d) {
    IACT_REQUIRE(s, "null scene");
    IACT_REQUIRE(s->n_facets >= 0 && s->n_samples >= 0, "negative facet/sample count");
    IACT_REQUIRE(s->n_facets == 0 || s->n_samples == 0 || (s->world && s->bounds), "null world table");
    IACT_REQUIRE(s->n_stages >= 0 && s->n_stages <= IACT_MAX_STAGES, "too many optical stages");
    IACT_REQUIRE(s->n_cyl >= 0 && s->n_box >= 0 && s->n_sph >= 0 && s->n_obox >= 0 && s->n_tri >= 0, "negative obstruction count");
    IACT_REQUIRE((long long)s->n_cyl + s->n_box + s->n_sph + s->n_obox + s->n_tri < 65535, "too many obstructions (max 65534)");
    memset(&d, 0, sizeof(d));
    d.F = s->n_facets; d.M = s->n_samples;
    d.world = reinterpret_cast<const float4*>(s->world); d.bounds = reinterpret_cast<const float4*>(s->bounds);
    d.chunk_bounds = reinterpret_cast<const float4*>(s->chunk_bounds);
    d.n_cyl = s->n_cyl; d.cyl_p1 = s->cyl_p1; d.cyl_p2 = s->cyl_p2; d.cyl_r = s->cyl_r;
    d.n_box = s->n_box; d.box_p1 = s->box_p1; d.box_p2 = s->box_p2;
    d.n_sph = s->n_sph; d.sph_c = s->sph_c; d.sph_r = s->sph_r;
    d.n_obox = s->n_obox; d.obox_c = s->obox_c; d.obox_h = s->obox_h; d.obox_R = s->obox_R;
    d.n_tri = s->n_tri; d.tri_v0 = s->tri_v0; d.tri_v1 = s->tri_v1; d.tri_v2 = s->tri_v2;
    IACT_REQUIRE(!d.n_cyl || (d.cyl_p1 && d.cyl_p2 && d.cyl_r), "null cylinder table");
    IACT_REQUIRE(!d.n_box || (d.box_p1 && d.box_p2), "null box table");
    IACT_REQUIRE(!d.n_sph || (d.sph_c && d.sph_r), "null sphere table");
    IACT_REQUIRE(!d.n_obox || (d.obox_c && d.obox_h && d.obox_R), "null oriented-box table");
    IACT_REQUIRE(!d.n_tri || (d.tri_v0 && d.tri_v1 && d.tri_v2), "null triangle table");
    d.n_stages = s->n_stages;
    for (int i = 0; i < s->n_stages; ++i) {
        d.stages[i].n = s->stages[i].n_mirrors; d.stages[i].rec = s->stages[i].records; d.stages[i].verts = s->stages[i].verts;
        IACT_REQUIRE(d.stages[i].n >= 0 && (d.stages[i].n == 0 || d.stages[i].rec), "bad mirror stage");
    }
    const IactSensor& se = s->sensor;
    IACT_REQUIRE(se.kind >= IACT_SENSOR_SQUARE && se.kind <= IACT_SENSOR_SOFT_HEX, "unknown sensor kind");
    if (se.kind == IACT_SENSOR_SQUARE || se.kind == IACT_SENSOR_SOFT_SQUARE) {
        IACT_REQUIRE(se.width > 0 && se.height > 0 && se.dx != 0.0 && se.dy != 0.0, "bad square sensor");
    } else {
        IACT_REQUIRE(se.n_pixels > 0 && se.n_pixels < 32768 && se.table_q > 0 && se.table_r > 0 && se.lookup && se.hex_size > 0.0,
                     "bad hexagonal sensor (n_pixels must be < 32768)");
    }
    if (se.kind >= IACT_SENSOR_SOFT_SQUARE) IACT_REQUIRE(se.sigma > 0.0 && se.kernel_size >= 0 && se.kernel_size <= 8, "bad soft-sensor parameters");
    fill_sensor(se, d.sens);
    d.cull = s->cull && (d.n_cyl + d.n_box + d.n_sph + d.n_obox + d.n_tri) > 0;
    return IACT_OK;
}

size_t smem_bytes(const SceneDev& d, int sens, int mode, int nwarps) {
    const int n_obs = d.n_cyl + d.n_box + d.n_sph + d.n_obox + d.n_tri;
    size_t fl = obstruction_floats(d.n_cyl, d.n_box, d.n_sph, d.n_obox, d.n_tri, d.cull != 0);
    fl += stage_floats(d);
    if (sens != SENS_SQUARE) {
        if (mode != MODE_DEBUG) fl += d.sens.npix;
        fl += (d.sens.tq * d.sens.tr + 1) / 2;
    }
    size_t bytes = fl * 4;
    if (d.cull) bytes += (size_t)nwarps * ((n_obs + 1) & ~1) * 2;
#if IACT_CYL_RECORDS
    if (d.cull && d.n_cyl > 0) bytes += (size_t)nwarps * CYL_REC_MAX * CYL_REC * 4;        // per-warp CylRec records
#endif
    return bytes + 16;
}

// Scratch for the level-1 lists, stream-ordered (no synchronisation).
struct Scratch {
    void* ptr = nullptr; cudaStream_t st = nullptr;
    int alloc(size_t bytes, cudaStream_t s) {
        st = s;
        // keep freed scratch in the device's default pool across synchronisations (the default
        // release threshold of 0 hands it back to the OS at every sync and re-maps it on the next call)
        static thread_local int pooled_dev = -1;
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev != pooled_dev) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pooled_dev = dev;
        }
        return iact_check_cuda(cudaMallocAsync(&ptr, bytes, s), "cudaMallocAsync");
    }
    ~Scratch() { if (ptr) cudaFreeAsync(ptr, st); }
};

// Launch level-1 culling into `scr`; fills `fl`.
int run_facet_cull(const SceneDev& d, const float* sources, int S, int source_type, Scratch& scr, FacetLists& fl, cudaStream_t st,
                   float* zero = nullptr, size_t n_zero = 0) {
    fl.ids = nullptr; fl.count = nullptr; fl.stride = 0; fl.counter = nullptr;
    if (!d.cull) return IACT_OK;
    const int n_obs = d.n_cyl + d.n_box + d.n_sph + d.n_obox + d.n_tri;
    const int stride = (n_obs + 7) & ~7;
    const size_t count_bytes = 16 + (((size_t)d.F * sizeof(int2) + 15) & ~(size_t)15);   // [0..8) = the work-queue counter
    int rc = scr.alloc(count_bytes + (size_t)d.F * stride * sizeof(unsigned short), st);
    if (rc) return rc;
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(scr.ptr);
    int2* count = reinterpret_cast<int2*>(reinterpret_cast<char*>(scr.ptr) + 16);
    unsigned short* ids = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(scr.ptr) + count_bytes);
    const size_t smem = (size_t)obstruction_floats(d.n_cyl, d.n_box, d.n_sph, d.n_obox, d.n_tri, true) * 4 + 16;
    if (smem > 200 * 1024) { iact_set_error("scene needs %zu bytes of shared memory per block (limit 204800)", smem); return IACT_ERR_UNSUPPORTED; }
    const int blocks = std::max(1, std::min((d.F + 7) / 8, sm_count() * 2));
    auto launch = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<blocks, 256, smem, st>>>(d, sources, S, ids, count, stride, counter, zero, n_zero);
    };
    if (source_type == IACT_SOURCE_POINT) launch(facet_cull_kernel<IACT_SOURCE_POINT>);
    else launch(facet_cull_kernel<IACT_SOURCE_PARALLEL>);
    iact_count_launch();
    fl.ids = ids; fl.count = count; fl.stride = stride; fl.counter = counter;
    return iact_check_cuda(cudaGetLastError(), "facet_cull_kernel launch");
}

// Split S x F x M rays into block items (source, facet chunk) and warp items (facet, sample range).
LaunchPlan make_plan(const SceneDev& d, int S, int mode) {
    LaunchPlan p;
    p.S = S;
    const bool hex = d.sens.kind == IACT_SENSOR_HEX || d.sens.kind == IACT_SENSOR_SOFT_HEX;
    // block items wanted for balance.  Response matrix on a hex camera: with at least 8 rows per SM a block owns whole
    // rows (plain-store flush, no atomics, no memset); with fewer rows (a row shard of a multi-GPU run: 512 rows on 148
    // SMs left 13 % of the block slots empty and a long tail) the rows are cut into facet chunks, ~48 items per SM
    const bool whole_rows = mode == MODE_MATRIX && hex && S >= 8 * sm_count();
    const long long target = (long long)sm_count() * ((mode == MODE_MATRIX && hex) ? 48 : 16);
    int n_chunks = 1;
    if (S < target && !whole_rows) n_chunks = (int)std::min<long long>(d.F, (target + S - 1) / std::max(S, 1));
    n_chunks = std::max(n_chunks, 1);
    p.chunk_facets = (d.F + n_chunks - 1) / n_chunks;
    p.n_chunks = (d.F + p.chunk_facets - 1) / p.chunk_facets;
    // keep >= 16 warp items per block item when the chunk is short
    p.msplit = 1;
    if (p.chunk_facets < 16) p.msplit = std::max(1, std::min((16 + p.chunk_facets - 1) / p.chunk_facets, (d.M + 63) / 64));
    p.msize = ((d.M + p.msplit - 1) / p.msplit + 31) / 32 * 32;
    p.msplit = (d.M + p.msize - 1) / std::max(p.msize, 1);
    p.n_items = (long long)S * p.n_chunks;
    return p;
}

// Units of the render / debug work queue: about 16 per resident warp when the job is large (short tail, one
// atomic per ~1e4 warp instructions); small jobs are split along the samples so that every warp gets work.
#ifndef IACT_QUEUE_FPU
#define IACT_QUEUE_FPU 8
#endif
#ifndef IACT_QUEUE_UNITS
#define IACT_QUEUE_UNITS 16          // 64 made 5920 warps pull 4.5e5 units from one counter in 0.5 ms on a 512-source job
#endif                               // (same-address atomics at 1 per ns): 634 -> 561 us; no change at 4096 sources
QueuePlan make_queue_plan(const SceneDev& d, int S, long long resident_warps) {
    QueuePlan q;
    const long long pairs = (long long)S * d.F;
    q.facets_per_unit = (int)std::max(1LL, std::min((long long)IACT_QUEUE_FPU, pairs / (resident_warps * IACT_QUEUE_UNITS)));
    q.runs = (d.F + q.facets_per_unit - 1) / q.facets_per_unit;
    const long long base = (long long)S * q.runs;
    long long ms = 1;
    if (base < resident_warps * 4) ms = std::min<long long>((d.M + 31) / 32, (resident_warps * 4 + base - 1) / std::max(base, 1LL));
    ms = std::max(ms, 1LL);
    q.msize = (int)(((d.M + ms - 1) / ms + 31) / 32 * 32);
    q.msplit = (d.M + q.msize - 1) / std::max(q.msize, 1);
    q.n_units = base * q.msplit;
    q.counter = nullptr;
    return q;
}

}  // namespace
