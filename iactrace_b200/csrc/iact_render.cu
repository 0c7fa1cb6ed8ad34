// K2/K3/K5/K7: forward trace + sensor binning -- render, render_response_matrix, render_debug
// (reference core/render.py:174-324, _trace_single_mirror :118-157) with exact obstruction culling.
//
// Work decomposition:
//   warp item   = (facet f, source s, sample range)          lane = one ray (sample m of facet f seen from source s)
//   render / render_debug: persistent grid; every warp pulls units (source, run of <= 8 facets, sample part) from a
//                          global counter (QueuePlan) -- no warp idles while another still has a backlog
//   response matrix:       one block owns a source row (its shared histogram is the row); whole rows are pulled
//                          from the same counter, the warps of the block stride over the row's facets
// Culling is hierarchical and conservative (a primitive is dropped only if no ray of a beam can
// touch it, with margins far above float32 rounding):
//   level 1 (facet_cull_kernel, once per call): facet x all sources -> per-facet candidate list in HBM/L2
//   level 2 (per warp item): (facet, source) beam against the facet list -> per-warp list in shared memory
//   level 3 (binned tables, SUB = true): every run of 32 table rows against the warp's list -> per-iteration mask
//            (strip test for far / parallel sources, capsule test otherwise)
// Lanes then test their ray only against the warp's short list (warp-uniform loop).
// Hex cameras are binned into a block-private shared-memory histogram through a per-warp register cache of up to
// three pixels (PixCache, with a fast path that skips rounding and lookup when every lane stays in the cached
// hexagon), flushed once per source (response matrix: plain coalesced stores) or once per block (render: one
// red.global per touched pixel); square cameras use red.global directly.
#include "iact_cull.cuh"

namespace {

// ---------------------------------------------------------------- binning
// Soft (Gaussian-splat) sensors: DifferentiableHexagonalSensor.accumulate (hexagonal.py:264-314)
template <typename LUT>
__device__ __forceinline__ void splat_soft_hex(const SensDev& se, const LUT* lut, float x, float y, float val, float* hist) {
    float xg, yg; hex_grid_coords(se, x, y, xg, yg);
    const float q = se.ax_qx * xg - se.ax_qy * yg, r = se.ax_ry * yg;
    float qb, rb; hex_round(q, r, qb, rb);
    if (!(fabsf(qb) < 1e6f && fabsf(rb) < 1e6f)) return;
    const float ddx = xg - se.size_sqrt3 * (qb + rb * 0.5f), ddy = yg - se.size_1p5 * rb;
    const int K = se.ksize;
    const float inv_sigma = 1.0f / se.sigma;
    float wsum = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
        for (int oq = -K; oq <= K; ++oq)
            for (int orr = -K; orr <= K; ++orr) {
                if (max(max(abs(oq), abs(orr)), abs(oq + orr)) > K) continue;
                const float ox = se.size_sqrt3 * ((float)oq + (float)orr * 0.5f), oy = se.size_1p5 * (float)orr;
                const float ax = fabsf(ddx - ox), ay = fabsf(ddy - oy);
                const float hd = fmaxf(ax, 0.5f * ax + 0.8660254037844386f * ay) * se.inv_inradius;
                const float z = hd * inv_sigma;
                const float w = gauss_half(z * z);
                if (pass == 0) { wsum += w; continue; }
                const int pix = hex_lookup(se, lut, qb + (float)oq, rb + (float)orr);
                if (pix >= 0) atomicAdd(hist + pix, val * (w / wsum));
            }
    }
}

// Square cameras have 1.5-2.1 Mpixel, far too many for a block-private histogram, so rays add straight into global
// memory.  A single image (render, iact_accumulate) is accumulated in FLOAT64 (red.global.add.f64 into a scratch
// image, converted to float32 by square_convert_kernel): a bright pixel receives 1e4-1e5 addends in a launch-dependent
// order, which in float32 wanders by a few 1e-4 from run to run (hazard H6); in float64 the order-dependent part is
// ~1e-12, so the float32 image is the correctly rounded sum and reproducible.  The reference sums per facet with
// segment_sum and adds the facet images (square.py:87, render.py:216).  Response-matrix rows (few addends per pixel,
// S x W x H outputs) stay float32.
#ifndef IACT_SQUARE_F64
#define IACT_SQUARE_F64 1
#endif

// DifferentiableSquareSensor.accumulate (square.py:144-172)
template <typename ACC>
__device__ __forceinline__ void splat_soft_square(const SensDev& se, float x, float y, float val, ACC* img) {
    const float xp = (x - se.x0) * se.inv_dx, yp = (y - se.y0) * se.inv_dy;
    const float xb = floorf(xp), yb = floorf(yp);
    const int K = se.ksize;
    if (!(xb >= (float)(-K - 1) && xb <= (float)(se.W + K) && yb >= (float)(-K - 1) && yb <= (float)(se.H + K))) return;
    const float fx = xp - xb, fy = yp - yb;
    const float inv_s2 = 1.0f / (se.sigma * se.sigma);
    float wsum = 0.f;
    for (int pass = 0; pass < 2; ++pass)
        for (int oy = -K; oy <= K; ++oy)
            for (int ox = -K; ox <= K; ++ox) {
                const float dx = fx - (float)ox, dy = fy - (float)oy;
                const float w = gauss_half((dx * dx + dy * dy) * inv_s2);
                if (pass == 0) { wsum += w; continue; }
                const int xi = (int)xb + ox, yi = (int)yb + oy;
                if (xi >= 0 && xi < se.W && yi >= 0 && yi < se.H) atomicAdd(img + (size_t)yi * se.W + xi, (ACC)(val * (w / wsum)));
            }
}

// Warp-aggregated histogram add: lanes holding the same pixel are summed with shuffles and one lane
// issues the shared-memory atomic (the PSF of one (facet, source) pair covers 1-3 hex pixels, so a
// plain per-lane atomicAdd would serialise 32 ways on the CAS loop).  pix < 0 = nothing to add.
// Must be called by all 32 lanes.
__device__ __forceinline__ void warp_hist_add(float* hist, int pix, float val) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned todo = __ballot_sync(0xffffffffu, pix >= 0);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int lp = __shfl_sync(0xffffffffu, pix, leader);
        const bool mine = pix == lp;
        float v = mine ? val : 0.f;
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (lane == (unsigned)leader) atomicAdd(hist + lp, v);
        todo &= ~__ballot_sync(0xffffffffu, mine);
    }
}

// DifferentiableHexagonalSensor.accumulate for a whole warp at once: every lane walks the same tap
// order, and each tap goes through warp_hist_add (the rays of one (facet, source) pair share their
// base hexagon, so per-lane shared atomics would serialise 32-way on every tap).  `active` = this
// lane has a hit to splat.  Must be called by all 32 lanes.
template <typename LUT>
__device__ __forceinline__ void splat_soft_hex_warp(const SensDev& se, const LUT* lut, bool active, float x, float y, float val,
                                                    float* hist) {
    float xg, yg; hex_grid_coords(se, x, y, xg, yg);
    const float q = se.ax_qx * xg - se.ax_qy * yg, r = se.ax_ry * yg;
    float qb, rb; hex_round(q, r, qb, rb);
    active = active && (fabsf(qb) < 1e6f) && (fabsf(rb) < 1e6f);
    if (!__any_sync(0xffffffffu, active)) return;
    const float ddx = xg - se.size_sqrt3 * (qb + rb * 0.5f), ddy = yg - se.size_1p5 * rb;
    const int K = se.ksize;
    const float inv_sigma = 1.0f / se.sigma;
    float wsum = 0.f;
    for (int oq = -K; oq <= K; ++oq)
        for (int orr = -K; orr <= K; ++orr) {
            if (max(max(abs(oq), abs(orr)), abs(oq + orr)) > K) continue;
            const float ox = se.size_sqrt3 * ((float)oq + (float)orr * 0.5f), oy = se.size_1p5 * (float)orr;
            const float ax = fabsf(ddx - ox), ay = fabsf(ddy - oy);
            const float z = fmaxf(ax, 0.5f * ax + 0.8660254037844386f * ay) * se.inv_inradius * inv_sigma;
            wsum += gauss_half(z * z);
        }
    const float scale = active ? val / wsum : 0.f;
    for (int oq = -K; oq <= K; ++oq)
        for (int orr = -K; orr <= K; ++orr) {
            if (max(max(abs(oq), abs(orr)), abs(oq + orr)) > K) continue;
            const float ox = se.size_sqrt3 * ((float)oq + (float)orr * 0.5f), oy = se.size_1p5 * (float)orr;
            const float ax = fabsf(ddx - ox), ay = fabsf(ddy - oy);
            const float z = fmaxf(ax, 0.5f * ax + 0.8660254037844386f * ay) * se.inv_inradius * inv_sigma;
            const float w = gauss_half(z * z);
            const int pix = active ? hex_lookup(se, lut, qb + (float)oq, rb + (float)orr) : -1;
            warp_hist_add(hist, pix, scale * w);
        }
}

// Soft hex splat with one ring of neighbours (7 taps), cached per warp item: the rays of one
// (facet, source) pair almost always share their base hexagon, so each lane keeps 7 per-tap partial sums
// for the warp's first base hexagon in registers and the shared histogram is touched 7 times per item.
// Rays with another base hexagon take the per-tap warp-aggregated path.
struct SoftHexCache {
    float qb0, rb0, qb1, rb1;          // base hexagons of the two slots (warp-uniform); 1e30 = empty
    float acc[7], acc1[7];
    __device__ __forceinline__ void reset() {
        qb0 = rb0 = qb1 = rb1 = 1e30f;
        for (int i = 0; i < 7; ++i) acc[i] = acc1[i] = 0.f;
    }
    // tap order: the 7 (oq, orr) pairs with max(|oq|, |orr|, |oq + orr|) <= 1, oq outer, orr inner
    template <typename LUT>
    __device__ __forceinline__ void add(const SensDev& se, const LUT* lut, bool active, float x, float y, float val, float* hist) {
        float xg, yg; hex_grid_coords(se, x, y, xg, yg);
        const float q = se.ax_qx * xg - se.ax_qy * yg, r = se.ax_ry * yg;
        float qb, rb; hex_round(q, r, qb, rb);
        active = active && (fabsf(qb) < 1e6f) && (fabsf(rb) < 1e6f);
        unsigned am = __ballot_sync(0xffffffffu, active);
        if (am == 0u) return;
        if (qb0 > 1e29f) {
            const int leader = __ffs(am) - 1;
            qb0 = __shfl_sync(0xffffffffu, qb, leader); rb0 = __shfl_sync(0xffffffffu, rb, leader);
        }
        bool c0 = active && qb == qb0 && rb == rb0;
        // a second base hexagon in this item (the spot straddles two cells in about a third of the items)
        const unsigned other = __ballot_sync(0xffffffffu, active && !c0);
        if (other != 0u && qb1 > 1e29f) {
            const int leader = __ffs(other) - 1;
            qb1 = __shfl_sync(0xffffffffu, qb, leader); rb1 = __shfl_sync(0xffffffffu, rb, leader);
        }
        const bool c1 = active && !c0 && qb == qb1 && rb == rb1;
        const float ddx = xg - se.size_sqrt3 * (qb + rb * 0.5f), ddy = yg - se.size_1p5 * rb;
        float w[7], wsum = 0.f;
        int t = 0;
#pragma unroll
        for (int oq = -1; oq <= 1; ++oq)
#pragma unroll
            for (int orr = -1; orr <= 1; ++orr) {
                if (oq + orr < -1 || oq + orr > 1) continue;
                const float ox = se.size_sqrt3 * ((float)oq + (float)orr * 0.5f), oy = se.size_1p5 * (float)orr;
                const float ax = fabsf(ddx - ox), ay = fabsf(ddy - oy);
                const float m = fmaxf(ax, 0.5f * ax + 0.8660254037844386f * ay);
                float wt; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wt) : "f"(se.soft_nk * (m * m)));   // exp(-z^2 / 2), z = m / (inradius sigma)
                w[t] = wt; wsum += wt; ++t;
            }
        const float scale = active ? val / wsum : 0.f;
        const float s0 = c0 ? scale : 0.f, s1 = c1 ? scale : 0.f;
#pragma unroll
        for (int i = 0; i < 7; ++i) { acc[i] = fmaf(s0, w[i], acc[i]); acc1[i] = fmaf(s1, w[i], acc1[i]); }
        if (__any_sync(0xffffffffu, active && !c0 && !c1)) {        // rare: a third base hexagon in this item
            t = 0;
            for (int oq = -1; oq <= 1; ++oq)
                for (int orr = -1; orr <= 1; ++orr) {
                    if (oq + orr < -1 || oq + orr > 1) continue;
                    const int pix = (active && !c0 && !c1) ? hex_lookup(se, lut, qb + (float)oq, rb + (float)orr) : -1;
                    warp_hist_add(hist, pix, scale * w[t]); ++t;
                }
        }
    }
    // Both slots at once: the 14 per-lane partial sums are reduced over the warp with a packed butterfly (each step
    // halves the values a lane carries: 8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 14 x 5), after which the even lanes
    // hold one total each and add their 14 distinct pixels in a single shared-atomic instruction.
    template <typename LUT>
    __device__ __forceinline__ void flush(const SensDev& se, const LUT* lut, float* hist) {
        if (qb0 > 1e29f) return;
        const unsigned lane = threadIdx.x & 31u;
        float v[16];
#pragma unroll
        for (int i = 0; i < 7; ++i) { v[i] = acc[i]; v[8 + i] = acc1[i]; }
        v[7] = 0.f; v[15] = 0.f;
#pragma unroll
        for (int n = 8, bit = 16; n >= 1; n >>= 1, bit >>= 1) {
            const bool up = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const float send = up ? v[i] : v[i + n], keep = up ? v[i + n] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
        const float tot = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
        const int idx = (int)(lane >> 1);                             // bits (16, 8, 4, 2) of the lane = value index
        const int slot = idx >> 3, tap = idx & 7;
        if ((lane & 1u) == 0u && tap < 7 && tot != 0.f) {
            const float qb = slot ? qb1 : qb0, rb = slot ? rb1 : rb0;
            const int oq = ((0x2211100 >> (4 * tap)) & 0xf) - 1, orr = ((0x1021021 >> (4 * tap)) & 0xf) - 1;
            const int pix = qb > 1e29f ? -1 : hex_lookup(se, lut, qb + (float)oq, rb + (float)orr);
            if (pix >= 0) atomicAdd(hist + pix, tot);
        }
    }
};

// Per-warp-item pixel cache: the rays of one (facet, source) pair land in 1-3 hex pixels, so the warp
// keeps up to three (pixel, per-lane partial sum) slots in registers across all its iterations and
// touches the shared histogram only when a slot is given up.  The cache outlives the warp items (the next
// facet of the same source feeds the same pixels); when a fourth pixel shows up the three slots are flushed
// and reused, and only an iteration that itself holds more than three distinct pixels falls back to direct
// shared atomics.  Slot pixels are warp-uniform.
struct PixCache {
    int p0, p1, p2;
    float a0, a1, a2;
    float cx, cy;            // centre of slot 0's hexagon in grid coordinates (1e30 = slot empty): the fast path of trace_ray
    float cx1, cy1;          // ... and of slot 1's
    __device__ __forceinline__ void reset() { p0 = p1 = p2 = -1; a0 = a1 = a2 = 0.f; cx = cy = cx1 = cy1 = 1e30f; }
    // must be called by all 32 lanes; pix < 0 = nothing to add; (pcx, pcy) = centre of the lane's hexagon
    __device__ __forceinline__ void add(float* hist, int pix, float val, float pcx, float pcy) {
        bool matched = (pix < 0) | (pix == p0) | (pix == p1) | (pix == p2);
        unsigned un = __ballot_sync(0xffffffffu, !matched);
        bool flushed = false;
        while (un != 0u) {
            if (p2 >= 0) {                                  // no free slot
                if (flushed) break;
                flush(hist); reset(); flushed = true;
                matched = pix < 0;
                un = __ballot_sync(0xffffffffu, !matched);
                continue;
            }
            const int leader = __ffs(un) - 1;
            const int lp = __shfl_sync(0xffffffffu, pix, leader);
            if (p0 < 0) {
                p0 = lp;
                cx = __shfl_sync(0xffffffffu, pcx, leader); cy = __shfl_sync(0xffffffffu, pcy, leader);
            } else if (p1 < 0) {
                p1 = lp;
                cx1 = __shfl_sync(0xffffffffu, pcx, leader); cy1 = __shfl_sync(0xffffffffu, pcy, leader);
            } else p2 = lp;
            matched = matched | (pix == lp);
            un = __ballot_sync(0xffffffffu, !matched);
        }
        if (pix >= 0) {                                     // empty slots hold -1: never match them
            if (pix == p0) a0 += val;
            else if (pix == p1) a1 += val;
            else if (pix == p2) a2 += val;
            else atomicAdd(hist + pix, val);
        }
        // the pixel of the first adding lane moves to slot 0, the one the next iteration's fast path tests
        const unsigned am = __ballot_sync(0xffffffffu, pix >= 0);
        if (am != 0u) {
            const int leader = __ffs(am) - 1;
            const int lp = __shfl_sync(0xffffffffu, pix, leader);
            if (lp != p0 && (lp == p1 || lp == p2)) {
                const float ncx = __shfl_sync(0xffffffffu, pcx, leader), ncy = __shfl_sync(0xffffffffu, pcy, leader);
                if (lp == p1) { p1 = p0; const float t = a1; a1 = a0; a0 = t; cx1 = cx; cy1 = cy; }
                else          { p2 = p0; const float t = a2; a2 = a0; a0 = t; }      // slot 2 keeps no centre
                p0 = lp; cx = ncx; cy = ncy;
            }
        }
    }
    __device__ __forceinline__ void flush(float* hist) {
        const unsigned lane = threadIdx.x & 31u;
        if (p0 >= 0) { float v = a0; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); if (lane == 0 && v != 0.f) atomicAdd(hist + p0, v); }
        if (p1 >= 0) { float v = a1; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); if (lane == 0 && v != 0.f) atomicAdd(hist + p1, v); }
        if (p2 >= 0) { float v = a2; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); if (lane == 0 && v != 0.f) atomicAdd(hist + p2, v); }
    }
};

// ---------------------------------------------------------------- the kernel
#ifndef IACT_HEX_FAST
#define IACT_HEX_FAST 1
#endif
#ifndef IACT_HEX_FAST2
#define IACT_HEX_FAST2 1     // second fast path (slot 1's hexagon) + rays outside the camera's bounding circle
#endif
#ifndef IACT_UNI_LIST
#define IACT_UNI_LIST 1      // level-2 list and cylinder records of a shared-direction item in one pass
#endif
#ifndef IACT_REC_MIN_ROWS
#define IACT_REC_MIN_ROWS 33 // rows per warp item from which cylinder records (and the one-pass level-2 list) are used
#endif
#ifndef IACT_STAGE_SIMPLE
#define IACT_STAGE_SIMPLE 1  // lean path for stages that are one conic mirror with a circular aperture (reflect_at_stage_simple)
#endif
#ifndef IACT_LEG_MASKS
#define IACT_LEG_MASKS 1     // leg towards optical stage 1 culled per 32-row run of a binned table (not per iteration)
#endif
#ifndef IACT_FAR_UNIFORM
#define IACT_FAR_UNIFORM 1   // one direction per (facet, source) item for point sources with parallax R / D < 1e-9
#endif

#ifndef IACT_MIN_BLOCKS
#define IACT_MIN_BLOCKS 4
#endif
#ifndef IACT_MIN_BLOCKS_RENDER_HEX
#define IACT_MIN_BLOCKS_RENDER_HEX 5 // render on a hard hex camera: five resident blocks at 48 registers (a few dozen bytes of
#endif                               // spills) beat four at 62 by 2-4 % (r2: also with level-3 culling, M = 4096: 87.3 -> 83.5 ms);
                                     // square cameras and the response matrix lose 2-20 % at five (spills) and stay at four
#ifndef IACT_MIN_BLOCKS_STAGES
#define IACT_MIN_BLOCKS_STAGES 4
#endif
#ifndef IACT_MIN_BLOCKS_SOFT
#define IACT_MIN_BLOCKS_SOFT 4       // soft hex cameras: two 7-tap register caches per warp
#endif
// Everything a warp needs to trace its rays, fixed for the lifetime of the block.
struct TraceCtx {
    ObsSmem ob;
    const float* stage_rec;
    float* hist;
    const short* lut;
    unsigned short* list;
    float* wrec;              // this warp's CylRec records (CYL_REC_MAX x CYL_REC floats) or nullptr
    bool cull, soft;
    bool stage_simple;        // every optical stage >= 1 is one conic mirror with a circular aperture (reflect_at_stage_simple)
};

// Per-block set-up shared by the trace kernels: obstruction tables, stage records, histogram, lookup table
// and the per-warp candidate lists in shared memory.  Ends with a block barrier.
template <int SENS, int MODE, bool STAGES>
__device__ __forceinline__ void trace_setup(const SceneDev& sc, float* smem, TraceCtx& cx) {
    cx.cull = sc.cull != 0;
    // per-warp CylRec records first: the dynamic shared memory is 16-byte aligned and a record is 64 bytes (LDS.128)
    cx.wrec = nullptr;
    if (IACT_CYL_RECORDS && cx.cull && sc.n_cyl > 0) {
        cx.wrec = smem + (size_t)(threadIdx.x >> 5) * (CYL_REC_MAX * CYL_REC);
        smem += (size_t)(blockDim.x >> 5) * (CYL_REC_MAX * CYL_REC);
    }
    stage_obstructions(sc, smem, cx.ob, cx.cull);
    const int n_obs = cx.ob.n_cyl + cx.ob.n_rest;
    float* p = smem + obstruction_floats(sc.n_cyl, sc.n_box, sc.n_sph, sc.n_obox, sc.n_tri, cx.cull);
    cx.stage_rec = p;
    if (STAGES) { stage_mirrors(sc, p); p += stage_floats(sc); }
    cx.hist = nullptr;
    cx.lut = nullptr;
    if (SENS != SENS_SQUARE) {
        if (MODE != MODE_DEBUG) { cx.hist = p; p += sc.sens.npix; }
        short* l = reinterpret_cast<short*>(p);
        for (int i = threadIdx.x; i < sc.sens.tq * sc.sens.tr; i += blockDim.x) l[i] = (short)sc.sens.lookup[i];
        cx.lut = l;
        p += (sc.sens.tq * sc.sens.tr + 1) / 2;
        if (cx.hist) for (int i = threadIdx.x; i < sc.sens.npix; i += blockDim.x) cx.hist[i] = 0.f;
    }
    const int warp = threadIdx.x >> 5;
    cx.list = cx.cull ? reinterpret_cast<unsigned short*>(p) + (size_t)warp * ((n_obs + 1) & ~1) : nullptr;
    cx.soft = SENS == SENS_SQUARE ? sc.sens.kind == IACT_SENSOR_SOFT_SQUARE : SENS == SENS_SOFT_HEX;
    __syncthreads();
    cx.stage_simple = false;
    if (STAGES && IACT_STAGE_SIMPLE) {
        bool simple = sc.n_stages > 0;
        const float* r = cx.stage_rec;
        for (int st = 0; st < sc.n_stages; ++st) { simple = simple && stage_is_simple(sc.stages[st].n, r); r += (size_t)sc.stages[st].n * STAGE_REC; }
        cx.stage_simple = simple;
    }
}

// One ray from table row (a, b) of a facet: shadow of the incoming leg, reflection, optical stages >= 1, sensor
// plane, binning (or the per-ray debug record at index ri).  `sd` is the source position, or -- when `uni` -- the
// direction all rays of this warp item share (parallel sources; far point sources, see trace_item); `n_rec` list
// entries then have a CylRec record.  Called by all 32 lanes; `live` = this lane holds a real ray.
template <int SRC, int SENS, int MODE, bool STAGES, bool SUB>
__device__ __forceinline__ void trace_ray(const SceneDev& sc, const TraceCtx& cx, float4 a, float4 b, V3 sd, bool uni, float sval, bool live,
                                          int n_list_cyl, int n_list, int n_rec, unsigned sub_mask, bool leg_static, unsigned leg_mask,
                                          size_t ri, bool soft7, PixCache& cache, SoftHexCache& scache, float* __restrict__ gout,
                                          float* __restrict__ out_val, int* __restrict__ out_pix) {
    const ObsSmem& ob = cx.ob;
    V3 o = v3(a.x, a.y, a.z);
    const V3 n = v3(b.x, b.y, b.z);
    // render.py:129-133
    V3 d = sd;
    if (SRC == IACT_SOURCE_POINT && !uni) {
        d = sub_rn(o, sd);
        d = scale_rn(frsqrt_nr_rn(dot_rn(d, d)), d);
    }
    // render.py:138 shadow of the incoming leg (infinite ray back towards the source)
    const bool blocked = occluded<SUB>(ob, o, -d, cx.list, n_list_cyl, n_list, sub_mask, cx.wrec, n_rec);
    // render.py:140-141, reflection.py:17-19
    const float c = dot_rn(d, n);
    d = fma_rn(__fmul_rn(-2.0f, c), n, d);
    float val = blocked ? 0.f : __fmul_rn(__fmul_rn(sval, -c), a.w);         // a.w = 1/weight (transform_kernel)
    if (STAGES) {
        const float* rec = cx.stage_rec;
        for (int st = 0; st < sc.n_stages; ++st) {
            bool leg_blocked = false;
            if (!cx.cull) leg_blocked = occluded(ob, o, d, nullptr, 0, 0);
            else if (SUB && leg_static && st == 0) {               // per-run mask of reachable primitives (leg_masks)
                const bool need = val != 0.f;
                for (unsigned mk = leg_mask; mk; mk &= mk - 1u) {
                    const int p = __ffs(mk) - 1;
                    if (need) leg_blocked |= hit_primitive(ob, p, o, d);
                }
            } else leg_blocked = occluded_leg_culled(ob, o, d, val != 0.f);
            if (cx.stage_simple) reflect_at_stage_simple(rec, leg_blocked, !cx.cull, o, d, val);
            else reflect_at_stage(sc.stages[st].n, rec, sc.stages[st].verts, leg_blocked, !cx.cull, o, d, val);
            rec += (size_t)sc.stages[st].n * STAGE_REC;
        }
    }
    // render.py:152-155
    float x, y;
    plane_hit(sc.sens, o, d, x, y);
    if (MODE == MODE_DEBUG) {
        if (live) {
            gout[2 * ri] = x; gout[2 * ri + 1] = y; out_val[ri] = val;
            if (out_pix) {
                int pix = -1;
                if (!cx.soft) pix = SENS != SENS_SQUARE ? hex_pixel(sc.sens, cx.lut, x, y) : square_pixel(sc.sens, x, y);
                out_pix[ri] = pix;
            }
        }
    } else {
        const bool add = live && val != 0.f;
        if (SENS == SENS_SOFT_HEX) {
            if (soft7) scache.add(sc.sens, cx.lut, add, x, y, val, cx.hist);
            else splat_soft_hex_warp(sc.sens, cx.lut, add, x, y, val, cx.hist);
        } else if (SENS == SENS_HEX) {
            // Fast path: the rays of one (facet, source) pair mostly share their hexagon with the previous iteration.
            // A hit whose hex norm about the cached centre is < 0.9999 lies strictly inside that cell, so the cube
            // rounding (hexagonal.py:32-39) would return it, and the edge rejection (:184-190) sees the identical
            // norm; when that holds for every adding lane the rounding, the table lookup and the slot search are skipped.
            float xg, yg; hex_grid_coords(sc.sens, x, y, xg, yg);
            const float fhn = hex_norm_rn(sc.sens, __fsub_rn(xg, cache.cx), __fsub_rn(yg, cache.cy));
            const bool in0 = fhn < 0.9999f;
            if (IACT_HEX_FAST && __all_sync(0xffffffffu, !add || in0)) {
                if (add && !(fhn > sc.sens.edge_thr)) cache.a0 += val;
            } else {
                // second fast path (30 % of the iterations on CT5: the spot of one (facet, source) pair straddles two
                // pixels): same argument against the hexagon of slot 1
                const float fhn1 = hex_norm_rn(sc.sens, __fsub_rn(xg, cache.cx1), __fsub_rn(yg, cache.cy1));
                const bool in1 = fhn1 < 0.9999f;
                // ... and rays that miss the camera altogether (18 % of the iterations: sources at the field edge)
                const bool out = __fmaf_rn(xg, xg, __fmul_rn(yg, yg)) > sc.sens.r_out2;
                if (IACT_HEX_FAST2 && __all_sync(0xffffffffu, !add || in0 || in1 || out)) {
                    if (add && in0) { if (!(fhn > sc.sens.edge_thr)) cache.a0 += val; }
                    else if (add && in1) { if (!(fhn1 > sc.sens.edge_thr)) cache.a1 += val; }
                } else {
                    float pcx, pcy;
                    const int pix = hex_pixel_grid(sc.sens, cx.lut, xg, yg, pcx, pcy);
                    cache.add(cx.hist, add ? pix : -1, val, pcx, pcy);
                }
            }
        } else if (add) {
            if (MODE == MODE_RENDER && IACT_SQUARE_F64) {            // gout is the float64 scratch image (run())
                double* g64 = reinterpret_cast<double*>(gout);
                if (cx.soft) splat_soft_square(sc.sens, x, y, val, g64);
                else { const int pix = square_pixel(sc.sens, x, y); if (pix >= 0) atomicAdd(g64 + pix, (double)val); }
            } else {
                if (cx.soft) splat_soft_square(sc.sens, x, y, val, gout);
                else { const int pix = square_pixel(sc.sens, x, y); if (pix >= 0) atomicAdd(gout + pix, val); }
            }
        }
    }
}

// Level-2 candidate list of one beam into the warp's shared-memory list.
__device__ __forceinline__ int item_list(const TraceCtx& cx, const FacetLists& fl, const Beam& beam, int f, int& n_list_cyl) {
    unsigned short* out = cx.list;
    const int n_obs = cx.ob.n_cyl + cx.ob.n_rest;
    const int2 cnt = fl.count ? __ldg(fl.count + f) : make_int2(-1, -1);
    if (cnt.x >= 0) return build_list(cx.ob, beam, fl.ids + (size_t)f * fl.stride, cnt.x, cnt.y, out, n_list_cyl);
    return build_list(cx.ob, beam, (const unsigned short*)nullptr, cx.ob.n_cyl, n_obs, out, n_list_cyl);
}

// One warp item: rays m0..m1 of facet f seen from source s (level-2 list, optional level-3 masks, per-ray trace).
template <int SRC, int SENS, int MODE, bool STAGES, bool SUB>
__device__ __forceinline__ void trace_item(const SceneDev& sc, const TraceCtx& cx, const FacetLists& fl, int S, int s, V3 src, float sval,
                                           int f, int m0, int m1, PixCache& cache, float* __restrict__ gout,
                                           float* __restrict__ out_val, int* __restrict__ out_pix) {
    const int lane = threadIdx.x & 31;
    const int M = sc.M;
    int n_list = 0, n_list_cyl = 0, n_rec = 0;
    Beam beam;
    beam.ok = false;
    // Rays of one item that share their direction: parallel sources, and point sources so far away that the parallax
    // across the facet (R / D < 1e-9) is below what float32 resolves in `normalize(p - src)` (render.py:130-131: the
    // subtraction itself rounds p away at that distance).  The direction is then evaluated once, from the facet
    // centre, and the direction half of the cylinder tests once per (item, candidate) into the warp's records.
    bool uni = SRC != IACT_SOURCE_POINT;
    V3 sd = src;
    const float4 bnd = __ldg(sc.bounds + f);
    if (SRC == IACT_SOURCE_POINT && IACT_FAR_UNIFORM) {
        const V3 ac = sub_rn(v3(bnd.x, bnd.y, bnd.z), src);
        const float n2 = dot_rn(ac, ac);
        if (bnd.w * bnd.w < 1e-18f * n2 && n2 < 1e37f) { uni = true; sd = scale_rn(frsqrt_nr_rn(n2), ac); }
    }
    // records pay from two 32-ray iterations per item on (CT3 response matrix at M = 64: 1.148 -> 1.067 ms; since the list
    // and the records come out of one pass they cost an item nothing extra) and are not used in the stage >= 1 kernels,
    // which are short of registers
    const bool want_rec = IACT_CYL_RECORDS && !STAGES && uni && cx.wrec && m1 - m0 >= IACT_REC_MIN_ROWS;
    if (cx.cull) {
        const int2 cnt = fl.count ? __ldg(fl.count + f) : make_int2(-1, -1);
        // (parallel directions are not normalised by the library: the one-pass form needs a unit direction)
        if (IACT_UNI_LIST && want_rec && cnt.x >= 0 && cnt.y <= 32 &&
            (SRC == IACT_SOURCE_POINT || fabsf(dot_rn(src, src) - 1.0f) < 1e-4f)) {
            // list and records in one pass (iact_cull.cuh build_list_uni); the beam axis is the shared direction itself
            beam.c = v3(bnd.x, bnd.y, bnd.z); beam.R = bnd.w; beam.spread = 0.f; beam.u = -sd; beam.ok = true;
            beam.invD = 0.f;
            if (SRC == IACT_SOURCE_POINT) { const V3 ac = sub_rn(beam.c, src); beam.invD = frsqrt_fast(dot_rn(ac, ac)); }
            n_list = build_list_uni(cx.ob, beam, fl.ids + (size_t)f * fl.stride, cnt.x, cnt.y, cx.list, cx.wrec, n_list_cyl, n_rec);
        } else {
            beam = make_beam<SRC>(bnd, src);
            n_list = item_list(cx, fl, beam, f, n_list_cyl);
            if (want_rec) {
                n_rec = min(n_list_cyl, CYL_REC_MAX);
                bool literal = false;
                if (lane < n_rec) literal = !cyl_record_write(cx.wrec + CYL_REC * lane, cx.ob.cyl + CYL_STRIDE * cx.list[lane], -sd);
                const unsigned lm = __ballot_sync(0xffffffffu, literal);      // (also orders the record writes before the reads)
                if (lm) n_rec = __ffs(lm) - 1;
            }
        }
    }
    const float4* tab = sc.world + ((size_t)f * M) * 2;
    SoftHexCache scache;
    const bool soft7 = SENS == SENS_SOFT_HEX && MODE != MODE_DEBUG && sc.sens.ksize == 1;
    if (soft7) scache.reset();
    // level-3 culling: with a binned table every run of 32 rows is a compact patch of the facet
    const bool sub_beams = SUB && cx.cull && n_list >= 1 && n_list <= 32;
    const float4* cbs = sub_beams ? sc.chunk_bounds + (size_t)f * ((M + 31) >> 5) * 2 : nullptr;
    // far or parallel sources: strip test, the masks of 32 consecutive runs at once (iact_cull.cuh strip_masks), so
    // nothing but one mask word per lane stays live across the ray loop; nearer sources: capsule test per run
    const bool strip = SUB && sub_beams && strip_applies(beam);
    auto run_strip_masks = [&](int run0) -> unsigned {
        const Beam b = make_beam<SRC>(__ldg(sc.bounds + f), src);           // recomputed: nothing of it stays live in the ray loop
        return strip_masks(cx.ob, b, cx.list, n_list_cyl, n_list, sc.chunk_bounds + (size_t)f * ((M + 31) >> 5) * 2, (M + 31) >> 5, run0);
    };
    unsigned run_masks = 0xffffffffu;                                        // lane j: run ((mb >> 5) & ~31) + j
    if (SUB && strip) run_masks = run_strip_masks((m0 >> 5) & ~31);
    // the leg towards the first optical stage >= 1, culled per run as well (iact_cull.cuh leg_masks)
    const bool leg_static = IACT_LEG_MASKS && STAGES && SUB && cx.cull && sc.n_stages > 0 && cx.ob.n_cyl + cx.ob.n_rest <= 32;
    auto run_leg_masks = [&](int run0) -> unsigned {
        return leg_masks<SRC>(cx.ob, src, uni, sd, sc.chunk_bounds + (size_t)f * ((M + 31) >> 5) * 2, (M + 31) >> 5, run0);
    };
    unsigned leg_run_masks = 0xffffffffu;
    if (STAGES && SUB && leg_static) leg_run_masks = run_leg_masks((m0 >> 5) & ~31);
    for (int mb = m0; mb < m1; mb += 32) {
        const int m = mb + lane;
        const bool live = m < m1;
        const int mm = live ? m : m1 - 1;
        const float4 a = __ldg(tab + 2 * mm), b = __ldg(tab + 2 * mm + 1);
        unsigned sub_mask = 0xffffffffu;
        if (SUB && sub_beams) {
            if (strip) {
                if (((mb >> 5) & 31) == 0 && mb != m0) run_masks = run_strip_masks(mb >> 5);
                sub_mask = __shfl_sync(0xffffffffu, run_masks, (mb >> 5) & 31);
            } else {
                const Beam cb = make_beam<SRC>(__ldg(cbs + 2 * (mb >> 5)), src);
                sub_mask = __ballot_sync(0xffffffffu, lane < n_list && (!cb.ok || keep_primitive(cx.ob, cb, cx.list[lane])));
            }
        }
        unsigned leg_mask = 0xffffffffu;
        if (STAGES && SUB && leg_static) {
            if (((mb >> 5) & 31) == 0 && mb != m0) leg_run_masks = run_leg_masks(mb >> 5);
            leg_mask = __shfl_sync(0xffffffffu, leg_run_masks, (mb >> 5) & 31);
        }
        const size_t ri = ((size_t)f * S + s) * M + __float_as_int(b.w);   // debug: original sample index
        trace_ray<SRC, SENS, MODE, STAGES, SUB>(sc, cx, a, b, sd, uni, sval, live, n_list_cyl, n_list, n_rec, sub_mask, leg_static,
                                                leg_mask, ri, soft7, cache, scache, gout, out_val, out_pix);
    }
    if (SENS == SENS_SOFT_HEX && soft7) scache.flush(sc.sens, cx.lut, cx.hist);
    __syncwarp();
}

// Work distribution.  Render / debug: every warp pulls units (source, run of facets, sample part) from a
// global counter (QueuePlan), so no warp idles while another still has a backlog -- with the static
// grid-stride split 18 % of the resident warp slots were empty (blocks and warps of unequal cost finishing
// early).  Response matrix: one block owns a source row (its histogram is the row), so whole block items are
// pulled from the same counter (queue.counter == nullptr there selects a static grid-stride split).
template <int SRC, int SENS, int MODE, bool STAGES, bool SUB>
__global__ void __launch_bounds__(256, STAGES ? IACT_MIN_BLOCKS_STAGES
                                              : ((MODE == MODE_RENDER && SENS == SENS_HEX) ? IACT_MIN_BLOCKS_RENDER_HEX
                                                 : (SENS == SENS_SOFT_HEX ? IACT_MIN_BLOCKS_SOFT : IACT_MIN_BLOCKS)))
trace_kernel(const __grid_constant__ SceneDev sc, const float* __restrict__ sources, const float* __restrict__ values,
             const LaunchPlan plan, const QueuePlan queue, const FacetLists fl, float* __restrict__ out,
             float* __restrict__ out_val, int* __restrict__ out_pix) {
    extern __shared__ __align__(16) float smem[];
    __shared__ long long s_item[2];
    TraceCtx cx;
    trace_setup<SENS, MODE, STAGES>(sc, smem, cx);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int M = sc.M;
    const size_t npix = SENS != SENS_SQUARE ? (size_t)sc.sens.npix : (size_t)sc.sens.W * sc.sens.H;
    // the pixel cache outlives the warp items: the next facet of the same source feeds the same few pixels,
    // and a new pixel evicts by flushing (PixCache::add); it is emptied before the histogram is read
    PixCache cache;
    cache.reset();

    if (MODE != MODE_MATRIX) {                    // compile-time: render / debug kernels hold the queue path only (code size)
        const unsigned long long per_src = (unsigned long long)queue.runs * queue.msplit;
        for (;;) {
            unsigned long long u = 0;
            if (lane == 0) u = atomicAdd(queue.counter, 1ull);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= (unsigned long long)queue.n_units) break;
            const int s = (int)(u / per_src);
            const int rem = (int)(u - (unsigned long long)s * per_src);
            const int run = rem / queue.msplit, part = rem - run * queue.msplit;
            const V3 src = v3(__ldg(sources + 3 * s), __ldg(sources + 3 * s + 1), __ldg(sources + 3 * s + 2));
            const float sval = __ldg(values + s);
            const int f0 = run * queue.facets_per_unit, f1 = min(sc.F, f0 + queue.facets_per_unit);
            const int m0 = part * queue.msize, m1 = min(M, m0 + queue.msize);
            for (int f = f0; f < f1; ++f)
                trace_item<SRC, SENS, MODE, STAGES, SUB>(sc, cx, fl, plan.S, s, src, sval, f, m0, m1, cache, out, out_val, out_pix);
        }
    } else {
        const bool pull = MODE == MODE_MATRIX && queue.counter != nullptr;
        long long item = blockIdx.x;
        int slot = 0;
        if (pull) {
            if (threadIdx.x == 0) s_item[0] = (long long)atomicAdd(queue.counter, 1ull);
            __syncthreads();
            item = s_item[0];
        }
        while (item < plan.n_items) {
            const int s = (int)(item / plan.n_chunks), ch = (int)(item - (long long)s * plan.n_chunks);
            const int f0 = ch * plan.chunk_facets, f1 = min(sc.F, f0 + plan.chunk_facets);
            const V3 src = v3(__ldg(sources + 3 * s), __ldg(sources + 3 * s + 1), __ldg(sources + 3 * s + 2));
            const float sval = __ldg(values + s);
            float* gout = MODE == MODE_MATRIX ? out + (size_t)s * npix : out;

            const int n_w = (f1 - f0) * plan.msplit;
            for (int wi = warp; wi < n_w; wi += nwarps) {
                int fi = wi, part = 0;
                if (plan.msplit > 1) { fi = wi / plan.msplit; part = wi - fi * plan.msplit; }
                const int m0 = part * plan.msize, m1 = min(M, m0 + plan.msize);
                trace_item<SRC, SENS, MODE, STAGES, SUB>(sc, cx, fl, plan.S, s, src, sval, f0 + fi, m0, m1, cache, gout, out_val, out_pix);
            }
            if (pull) {                                  // next item: slots alternate, so a slow reader never sees an overwrite
                slot ^= 1;
                if (threadIdx.x == 0) s_item[slot] = (long long)atomicAdd(queue.counter, 1ull);
            }
            if (MODE == MODE_MATRIX && SENS != SENS_SQUARE) {
                if (SENS == SENS_HEX) { cache.flush(cx.hist); cache.reset(); }
                __syncthreads();
                if (plan.n_chunks == 1) {
                    for (int i = threadIdx.x; i < sc.sens.npix; i += blockDim.x) { gout[i] = cx.hist[i]; cx.hist[i] = 0.f; }
                } else {
                    for (int i = threadIdx.x; i < sc.sens.npix; i += blockDim.x) {
                        const float v = cx.hist[i];
                        if (v != 0.f) { atomicAdd(gout + i, v); cx.hist[i] = 0.f; }
                    }
                }
            }
            if (pull) { __syncthreads(); item = s_item[slot]; }
            else {
                if (MODE == MODE_MATRIX && SENS != SENS_SQUARE) __syncthreads();
                item += gridDim.x;
            }
        }
    }
    if (MODE == MODE_RENDER && SENS != SENS_SQUARE) {
        if (SENS == SENS_HEX) cache.flush(cx.hist);
        __syncthreads();
        for (int i = threadIdx.x; i < sc.sens.npix; i += blockDim.x) {
            const float v = cx.hist[i];
            if (v != 0.f) atomicAdd(out + i, v);
        }
    }
}

// Culling statistics: level-2 candidate-list lengths summed over all (facet, source) pairs.
template <int SRC>
__global__ void __launch_bounds__(256) cull_stats_kernel(const __grid_constant__ SceneDev sc, const float* __restrict__ sources,
                                                         int S, const FacetLists fl, unsigned long long* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    ObsSmem ob;
    stage_obstructions(sc, smem, ob, true);
    const int n_obs = ob.n_cyl + ob.n_rest;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned short* list = reinterpret_cast<unsigned short*>(smem + obstruction_floats(sc.n_cyl, sc.n_box, sc.n_sph, sc.n_obox, sc.n_tri, true))
                           + (size_t)warp * ((n_obs + 1) & ~1);
    __syncthreads();
    unsigned long long a = 0, b = 0, c = 0, l1 = 0;
    const long long n_pairs = (long long)sc.F * S;
    const int M = sc.M, n_chunks = (M + 31) >> 5;
    for (long long i = (long long)blockIdx.x * nwarps + warp; i < n_pairs; i += (long long)gridDim.x * nwarps) {
        const int f = (int)(i / S), s = (int)(i - (long long)f * S);
        const V3 src = v3(sources[3 * s], sources[3 * s + 1], sources[3 * s + 2]);
        const Beam beam = make_beam<SRC>(__ldg(sc.bounds + f), src);
        const int2 cnt = __ldg(fl.count + f);
        int ncyl = 0, n;
        if (cnt.x >= 0) n = build_list(ob, beam, fl.ids + (size_t)f * fl.stride, cnt.x, cnt.y, list, ncyl);
        else            n = build_list(ob, beam, (const unsigned short*)nullptr, ob.n_cyl, n_obs, list, ncyl);
        if (sc.chunk_bounds && n >= 1 && n <= 32) {
            // level 3: what each 32-row run actually tests (same rule as trace_kernel<..., SUB = true>)
            const bool strip = strip_applies(beam);
            unsigned masks = 0u;
            for (int k = 0; k < n_chunks; ++k) {
                unsigned mask;
                if (strip) {
                    if ((k & 31) == 0) masks = strip_masks(ob, beam, list, ncyl, n, sc.chunk_bounds + (size_t)f * n_chunks * 2, n_chunks, k);
                    mask = __shfl_sync(0xffffffffu, masks, k & 31);
                } else {
                    const Beam cb = make_beam<SRC>(__ldg(sc.chunk_bounds + ((size_t)f * n_chunks + k) * 2), src);
                    mask = __ballot_sync(0xffffffffu, lane < n && (!cb.ok || keep_primitive(ob, cb, list[lane])));
                }
                const unsigned mc = ncyl >= 32 ? mask : (mask & ((1u << ncyl) - 1u));
                const int rays = min(32, M - 32 * k);
                a += (unsigned long long)__popc(mc) * rays; b += (unsigned long long)__popc(mask & ~mc) * rays;
            }
        } else {
            a += (unsigned long long)ncyl * M; b += (unsigned long long)(n - ncyl) * M;
        }
        c += M; l1 += cnt.x >= 0 ? cnt.y : n_obs;
        __syncwarp();
    }
    if (lane == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); atomicAdd(out + 2, c); atomicAdd(out + 3, l1); }
}

// float64 scratch image -> float32 image (square cameras, see IACT_SQUARE_F64)
__global__ void __launch_bounds__(256) square_convert_kernel(const double* __restrict__ acc, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)acc[i];
}
int launch_square_convert(const double* acc, float* out, size_t n, cudaStream_t st) {
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 8);
    square_convert_kernel<<<grid, 256, 0, st>>>(acc, out, n);
    iact_count_launch();
    return iact_check_cuda(cudaGetLastError(), "square_convert_kernel launch");
}

// sensor.accumulate on free-standing hits: one thread per hit, red.global into the image (`out64` = float64 scratch
// image of the square cameras when IACT_SQUARE_F64).
__global__ void __launch_bounds__(256) accumulate_kernel(const SensDev se, const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ v, long long n, float* __restrict__ out,
                                                         double* __restrict__ out64) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float val = v[i];
    switch (se.kind) {
        case IACT_SENSOR_SQUARE: {
            const int p = square_pixel(se, x[i], y[i]);
            if (p >= 0) { if (out64) atomicAdd(out64 + p, (double)val); else atomicAdd(out + p, val); }
            break;
        }
        case IACT_SENSOR_HEX:    { const int p = hex_pixel(se, se.lookup, x[i], y[i]); if (p >= 0) atomicAdd(out + p, val); break; }
        case IACT_SENSOR_SOFT_SQUARE: if (out64) splat_soft_square(se, x[i], y[i], val, out64); else splat_soft_square(se, x[i], y[i], val, out); break;
        default: splat_soft_hex(se, se.lookup, x[i], y[i], val, out); break;
    }
}

template <int MODE, typename K>
int launch_kernel(K kern, const SceneDev& d, const float* sources, const float* values, const LaunchPlan& plan, const FacetLists& fl,
                  float* out, float* out_val, int* out_pix, cudaStream_t stream, size_t smem) {
    const int threads = 256;
    if (smem > 48 * 1024) IACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    IACT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    const long long max_blocks = (long long)sm_count() * occ;
    QueuePlan q = make_queue_plan(d, plan.S, max_blocks * (threads / 32));
    long long blocks = plan.n_items;
    Scratch ctr;
    if (fl.counter) {
        q.counter = fl.counter;                                   // zeroed by facet_cull_kernel
    } else {
        int rc = ctr.alloc(sizeof(unsigned long long), stream);
        if (rc) return rc;
        IACT_CUDA(cudaMemsetAsync(ctr.ptr, 0, sizeof(unsigned long long), stream));
        q.counter = reinterpret_cast<unsigned long long*>(ctr.ptr);
    }
    if (MODE != MODE_MATRIX) blocks = (q.n_units + threads / 32 - 1) / (threads / 32);
    const unsigned grid = (unsigned)std::max(1LL, std::min(blocks, max_blocks));
    kern<<<grid, threads, smem, stream>>>(d, sources, values, plan, q, fl, out, out_val, out_pix);
    iact_count_launch();
    return iact_check_cuda(cudaGetLastError(), "trace_kernel launch");
}

template <int SRC, int SENS, int MODE, bool STAGES>
int launch_variant(const SceneDev& d, const float* sources, const float* values, const LaunchPlan& plan, const FacetLists& fl,
                   float* out, float* out_val, int* out_pix, cudaStream_t stream) {
    const int threads = 256;
    const size_t smem = smem_bytes(d, SENS, MODE, threads / 32);
    if (smem > 200 * 1024) { iact_set_error("scene needs %zu bytes of shared memory per block (limit 204800)", smem); return IACT_ERR_UNSUPPORTED; }
    if (d.chunk_bounds) return launch_kernel<MODE>(trace_kernel<SRC, SENS, MODE, STAGES, true>, d, sources, values, plan, fl, out, out_val, out_pix, stream, smem);
    return launch_kernel<MODE>(trace_kernel<SRC, SENS, MODE, STAGES, false>, d, sources, values, plan, fl, out, out_val, out_pix, stream, smem);
}

#define ARGS const SceneDev& d, const float* a, const float* b, const LaunchPlan& p, const FacetLists& fl, float* o, float* ov, int* op, cudaStream_t st
#define PASS d, a, b, p, fl, o, ov, op, st
template <int SRC, int SENS, int MODE>
int launch_stages(ARGS) {
    return d.n_stages > 0 ? launch_variant<SRC, SENS, MODE, true>(PASS) : launch_variant<SRC, SENS, MODE, false>(PASS);
}
template <int SRC, int MODE>
int launch_sens(ARGS) {
    if (d.sens.kind == IACT_SENSOR_HEX) return launch_stages<SRC, SENS_HEX, MODE>(PASS);
    if (d.sens.kind == IACT_SENSOR_SOFT_HEX) return launch_stages<SRC, SENS_SOFT_HEX, MODE>(PASS);
    return launch_stages<SRC, SENS_SQUARE, MODE>(PASS);
}
template <int MODE>
int launch_src(int source_type, ARGS) {
    return source_type == IACT_SOURCE_POINT ? launch_sens<IACT_SOURCE_POINT, MODE>(PASS) : launch_sens<IACT_SOURCE_PARALLEL, MODE>(PASS);
}
#undef ARGS
#undef PASS

int run(const IactScene* scene, const float* sources, const float* values, int S, int source_type, int mode,
        float* out, float* out_val, int* out_pix, void* stream) {
    SceneDev d;
    int rc = fill_scene(scene, d);
    if (rc) return rc;
    IACT_REQUIRE(S >= 0, "negative source count");
    IACT_REQUIRE(source_type == IACT_SOURCE_POINT || source_type == IACT_SOURCE_PARALLEL, "bad source_type");
    if (mode != MODE_RENDER && S == 0) return IACT_OK;                      // zero-row outputs have no storage
    IACT_REQUIRE(out, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    const bool hex = d.sens.kind == IACT_SENSOR_HEX || d.sens.kind == IACT_SENSOR_SOFT_HEX;
    const size_t npix = hex ? (size_t)d.sens.npix : (size_t)d.sens.W * d.sens.H;
    const bool empty = S == 0 || d.F == 0 || d.M == 0;
    if (!empty) IACT_REQUIRE(sources && values, "null sources/values");
    LaunchPlan plan = make_plan(d, std::max(S, 1), mode);
    Scratch acc64;                                                           // float64 image of a square camera
    float* final_out = out;
    if (mode == MODE_RENDER && !hex && IACT_SQUARE_F64 && !empty) {
        rc = acc64.alloc(npix * sizeof(double), st);
        if (rc) return rc;
        IACT_CUDA(cudaMemsetAsync(acc64.ptr, 0, npix * sizeof(double), st));
        out = reinterpret_cast<float*>(acc64.ptr);
    } else if (mode == MODE_RENDER) {
        if (!(hex && !empty && d.cull && S >= 4))                            // else: zeroed by facet_cull_kernel below
            IACT_CUDA(cudaMemsetAsync(out, 0, npix * sizeof(float), st));    // render.py:198-199,218
    } else if (mode == MODE_MATRIX) {
        if (empty || !(hex && plan.n_chunks == 1)) IACT_CUDA(cudaMemsetAsync(out, 0, (size_t)S * npix * sizeof(float), st));
    } else {
        IACT_REQUIRE(out_val, "null out_val");
    }
    if (empty) return IACT_OK;
    plan.S = S;
    plan.n_items = (long long)S * plan.n_chunks;
    Scratch scr;
    FacetLists fl;
    fl.ids = nullptr; fl.count = nullptr; fl.stride = 0; fl.counter = nullptr;
    // level-1 lists pay off once a facet is seen from several sources
    if (d.cull && S >= 4) {
        const bool zero_image = mode == MODE_RENDER && hex;
        rc = run_facet_cull(d, sources, S, source_type, scr, fl, st, zero_image ? out : nullptr, zero_image ? npix : 0);
        if (rc) return rc;
    }
    switch (mode) {
        case MODE_RENDER:
            rc = launch_src<MODE_RENDER>(source_type, d, sources, values, plan, fl, out, out_val, out_pix, st);
            if (rc || !acc64.ptr) return rc;
            return launch_square_convert(reinterpret_cast<const double*>(acc64.ptr), final_out, npix, st);
        case MODE_MATRIX: return launch_src<MODE_MATRIX>(source_type, d, sources, values, plan, fl, out, out_val, out_pix, st);
        default:          return launch_src<MODE_DEBUG>(source_type, d, sources, values, plan, fl, out, out_val, out_pix, st);
    }
}

}  // namespace

extern "C" int iact_cull_stats(const IactScene* scene, const float* sources, int n_sources, int source_type,
                               unsigned long long* out4, void* stream) {
    SceneDev d;
    IACT_REQUIRE(scene, "null scene");
    IactScene tmp = *scene;
    tmp.cull = 1;
    int rc = fill_scene(&tmp, d);
    if (rc) return rc;
    IACT_REQUIRE(sources && out4 && n_sources > 0 && d.F > 0, "bad arguments");
    IACT_REQUIRE(d.cull, "scene has no obstructions");
    cudaStream_t st = (cudaStream_t)stream;
    IACT_CUDA(cudaMemsetAsync(out4, 0, 4 * sizeof(unsigned long long), st));
    Scratch scr;
    FacetLists fl;
    rc = run_facet_cull(d, sources, n_sources, source_type, scr, fl, st);
    if (rc) return rc;
    const size_t smem = smem_bytes(d, SENS_SQUARE, MODE_DEBUG, 8);
    if (smem > 200 * 1024) { iact_set_error("scene too large for shared memory"); return IACT_ERR_UNSUPPORTED; }
    auto launch = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<sm_count() * 4, 256, smem, st>>>(d, sources, n_sources, fl, out4);
    };
    if (source_type == IACT_SOURCE_POINT) launch(cull_stats_kernel<IACT_SOURCE_POINT>);
    else launch(cull_stats_kernel<IACT_SOURCE_PARALLEL>);
    iact_count_launch();
    return iact_check_cuda(cudaGetLastError(), "cull_stats_kernel launch");
}

extern "C" int iact_accumulate(const IactSensor* sensor, const float* x, const float* y, const float* values,
                               long long n, float* out, void* stream) {
    IACT_REQUIRE(sensor && out && n >= 0, "bad arguments");
    IACT_REQUIRE(n == 0 || (x && y && values), "null input");
    IactScene tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.sensor = *sensor;
    SceneDev d;
    int rc = fill_scene(&tmp, d);
    if (rc) return rc;
    const bool hex = d.sens.kind == IACT_SENSOR_HEX || d.sens.kind == IACT_SENSOR_SOFT_HEX;
    const size_t npix = hex ? (size_t)d.sens.npix : (size_t)d.sens.W * d.sens.H;
    cudaStream_t st = (cudaStream_t)stream;
    Scratch acc64;
    if (!hex && IACT_SQUARE_F64 && n > 0) {
        rc = acc64.alloc(npix * sizeof(double), st);
        if (rc) return rc;
        IACT_CUDA(cudaMemsetAsync(acc64.ptr, 0, npix * sizeof(double), st));
    } else {
        IACT_CUDA(cudaMemsetAsync(out, 0, npix * sizeof(float), st));
    }
    if (n == 0) return IACT_OK;
    accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.sens, x, y, values, n, out, reinterpret_cast<double*>(acc64.ptr));
    iact_count_launch();
    rc = iact_check_cuda(cudaGetLastError(), "accumulate_kernel launch");
    if (rc || !acc64.ptr) return rc;
    return launch_square_convert(reinterpret_cast<const double*>(acc64.ptr), out, npix, st);
}

extern "C" int iact_render(const IactScene* scene, const float* sources, const float* values, int n_sources,
                           int source_type, float* out_image, void* stream) {
    return run(scene, sources, values, n_sources, source_type, MODE_RENDER, out_image, nullptr, nullptr, stream);
}

extern "C" int iact_response_matrix(const IactScene* scene, const float* sources, const float* values, int n_sources,
                                    int source_type, float* out_matrix, void* stream) {
    return run(scene, sources, values, n_sources, source_type, MODE_MATRIX, out_matrix, nullptr, nullptr, stream);
}

extern "C" int iact_render_debug(const IactScene* scene, const float* sources, const float* values, int n_sources,
                                 int source_type, float* out_xy, float* out_val, int32_t* out_pixel, void* stream) {
    return run(scene, sources, values, n_sources, source_type, MODE_DEBUG, out_xy, out_val, out_pixel, stream);
}
