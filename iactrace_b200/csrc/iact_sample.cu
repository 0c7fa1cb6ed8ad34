// Load-time facet sampling and the per-render local->world transform.
//
// K1: MCIntegrator.sample_group            (reference core/integrators.py:68-188)
//     sample_disk / sample_polygon         (reference utils/sampling.py:8-67)
//     AsphericSurface.point_and_normal     (reference core/surfaces.py:41-65)
//     compute_perturbation_delta           (reference core/reflection.py:22-49)
// K8: MirrorGroup.transform_to_world       (reference telescope/mirrors.py:64-79)
//
// One thread per (facet, sample).  The JAX key tree is evaluated with random
// access (threefry2x32 is counter based), so no state is carried between
// threads and any sample can be regenerated in isolation.
#include "iact_common.cuh"

namespace {

struct SampleArgs {
    Key key; int mode; int F, M;        // M = rows written per facet
    int m0, Mtot;                       // ... which are samples m0 .. m0 + M - 1 of a stream of Mtot samples per facet
    SurfDev surf;
    int kind;              // 0 disk, 1 polygon
    int nv;
    const float* radii; const float* verts; const float* offsets;
    float *points, *normals, *delta, *weights;
};

__global__ void __launch_bounds__(256) sample_kernel(const SampleArgs a) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)a.F * a.M) return;
    const int f = (int)(gid / a.M);
    const uint32_t m = (uint32_t)a.m0 + (uint32_t)(gid % a.M);         // index in the facet's full stream (random access)
    const uint32_t M = (uint32_t)a.Mtot;
    const int mode = a.mode;

    // integrators.py:108/153 keys = split(key, n_mirrors); :113/158 key_sample, key_perturb = split(mkey)
    const Key mkey = rng_split(a.key, (uint32_t)f, (uint32_t)a.F, mode);
    const Key ks = rng_split(mkey, 0u, 2u, mode);
    const Key kp = rng_split(mkey, 1u, 2u, mode);

    float x, y, area;
    if (a.kind == 0) {
        // sampling.py:19-27
        const Key k1 = rng_split(ks, 0u, 2u, mode), k2 = rng_split(ks, 1u, 2u, mode);
        const float r = sqrtf(rng_uniform(k1, m, M, mode, 0.f, 1.f));
        const float th = __fmul_rn(__fmul_rn(rng_uniform(k2, m, M, mode, 0.f, 1.f), 2.0f), 3.14159274f);
        float s, c; sincosf(th, &s, &c);
        const float radius = a.radii[f];
        x = __fmul_rn(__fmul_rn(r, c), radius);
        y = __fmul_rn(__fmul_rn(r, s), radius);
        area = 3.14159274f * (radius * radius);                       // integrators.py:126
    } else {
        // sampling.py:42-67: fan triangulation from vertex 0, area-weighted triangle choice
        const float* V = a.verts + (size_t)f * a.nv * 2;
        const int nt = a.nv - 2;
        float areas[IACT_MAX_POLY];
        float tot = 0.f;
        const float v0x = V[0], v0y = V[1];
        for (int i = 0; i < nt; ++i) {
            const float ax = V[2 * (i + 1)] - v0x, ay = V[2 * (i + 1) + 1] - v0y;
            const float bx = V[2 * (i + 2)] - v0x, by = V[2 * (i + 2) + 1] - v0y;
            areas[i] = fabsf(__fsub_rn(__fmul_rn(ax, by), __fmul_rn(bx, ay))) / 2.0f;
            tot = __fadd_rn(tot, areas[i]);
        }
        const Key k1 = rng_split(ks, 0u, 4u, mode), k2 = rng_split(ks, 1u, 4u, mode),
                  k3 = rng_split(ks, 2u, 4u, mode);
        // jax.random.choice(p=...): r = cum[-1] * (1 - U); idx = searchsorted(cum, r, 'left')
        float cum[IACT_MAX_POLY];
        float run = 0.f;
        for (int i = 0; i < nt; ++i) { run = __fadd_rn(run, areas[i] / tot); cum[i] = run; }
        const float rr = __fmul_rn(cum[nt - 1], __fsub_rn(1.0f, rng_uniform(k1, m, M, mode, 0.f, 1.f)));
        int tri = 0;
        while (tri < nt && cum[tri] < rr) ++tri;
        if (tri >= nt) tri = nt - 1;                                   // gather clamps out-of-range
        const float u = sqrtf(rng_uniform(k2, m, M, mode, 0.f, 1.f));
        const float v = rng_uniform(k3, m, M, mode, 0.f, 1.f);
        const float wa = __fsub_rn(1.0f, u), wb = __fmul_rn(u, __fsub_rn(1.0f, v)), wc = __fmul_rn(u, v);
        const float bx = V[2 * (tri + 1)], by = V[2 * (tri + 1) + 1];
        const float cx = V[2 * (tri + 2)], cy = V[2 * (tri + 2) + 1];
        x = __fadd_rn(__fadd_rn(__fmul_rn(wa, v0x), __fmul_rn(wb, bx)), __fmul_rn(wc, cx));
        y = __fadd_rn(__fadd_rn(__fmul_rn(wa, v0y), __fmul_rn(wb, by)), __fmul_rn(wc, cy));
        // integrators.py:172-174 shoelace
        float sh = 0.f;
        for (int i = 0; i < a.nv; ++i) {
            const int j = (i + 1 == a.nv) ? 0 : i + 1;
            sh = __fadd_rn(sh, __fsub_rn(__fmul_rn(V[2 * i], V[2 * j + 1]), __fmul_rn(V[2 * j], V[2 * i + 1])));
        }
        area = 0.5f * fabsf(sh);
    }

    // surfaces.py:41-58: point = (x, y, sag_raw(x+x0, y+y0) - sag_raw(x0, y0)); normal ∝ (-dz/dx, -dz/dy, 1)
    const float x0 = a.offsets[2 * f], y0 = a.offsets[2 * f + 1];
    const float xs = x + x0, ys = y + y0;
    const float z = sag_raw(a.surf, xs, ys) - sag_raw(a.surf, x0, y0);
    const float g = dsag_dr2(a.surf, xs * xs + ys * ys);
    float nx = -(g * (xs + xs)), ny = -(g * (ys + ys)), nz = 1.0f;
    const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
    nx *= inv; ny *= inv; nz *= inv;

    // reflection.py:35-49 tangent-space unit-sigma deltas
    const Key ka = rng_split(kp, 0u, 2u, mode), kb = rng_split(kp, 1u, 2u, mode);
    const float th1 = rng_normal(ka, m, M, mode), th2 = rng_normal(kb, m, M, mode);
    V3 n = v3(nx, ny, nz);
    V3 ref = fabsf(nz) > 0.9f ? v3(1.f, 0.f, 0.f) : v3(0.f, 0.f, 1.f);
    V3 t1 = cross(n, ref);
    t1 = (1.0f / sqrtf(dot(t1, t1))) * t1;
    V3 t2 = cross(n, t1);
    V3 d = th1 * t1 + th2 * t2;

    const size_t o = (size_t)gid * 3;
    a.points[o] = x;  a.points[o + 1] = y;  a.points[o + 2] = z;
    a.normals[o] = nx; a.normals[o + 1] = ny; a.normals[o + 2] = nz;
    a.delta[o] = d.x; a.delta[o + 1] = d.y; a.delta[o + 2] = d.z;
    a.weights[gid] = __fmul_rn(nz / area, (float)a.Mtot);               // integrators.py:127/175
}

__global__ void __launch_bounds__(256) random_kernel(Key key, int mode, int n, int normal, float lo, float hi, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = normal ? rng_normal(key, (uint32_t)i, (uint32_t)n, mode)
                    : rng_uniform(key, (uint32_t)i, (uint32_t)n, mode, lo, hi);
}

// One block per facet: rotate/translate the facet's samples into the world frame, fold the
// roughness perturbation into the normal, and reduce the facet's bounding sphere.
__global__ void __launch_bounds__(256) transform_kernel(const IactFacets fa, int facet_offset, float* world, float* bounds) {
    const int f = blockIdx.x;
    const M33 R = euler_to_matrix(fa.rotations[3 * f], fa.rotations[3 * f + 1], fa.rotations[3 * f + 2]);
    const V3 pos = ld3(fa.positions + 3 * f);
    const float scale = fa.scale[f];
    const int M = fa.n_samples;
    float4* out = reinterpret_cast<float4*>(world) + ((size_t)(facet_offset + f) * M) * 2;
    float maxd2 = 0.f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const size_t i = ((size_t)f * M + m) * 3;
        const V3 pl = ld3(fa.points + i), nl = ld3(fa.normals + i), dl = ld3(fa.delta + i);
        const V3 rp = mul(R, pl);
        const V3 pw = rp + pos;
        V3 nw = mul(R, nl) + scale * mul(R, dl);
        nw = (1.0f / sqrtf(dot(nw, nw))) * nw;
        out[2 * m] = make_float4(pw.x, pw.y, pw.z, 1.0f / fa.weights[(size_t)f * M + m]);   // value = v cos / w (render.py:141)
        out[2 * m + 1] = make_float4(nw.x, nw.y, nw.z, __int_as_float(m));
        const V3 dd = pw - pos;
        maxd2 = fmaxf(maxd2, dot(dd, dd));
    }
    __shared__ float red[256];
    red[threadIdx.x] = maxd2;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float* b = bounds + 4 * (size_t)(facet_offset + f);
        b[0] = pos.x; b[1] = pos.y; b[2] = pos.z;
        b[3] = sqrtf(red[0]) * 1.00001f + 1e-6f;
    }
}

// Binned variant: counting sort of the facet's samples into G x G cells (serpentine order) before the
// transform, so that 32 consecutive rows form a compact patch; then one bounding sphere per run of 32.
__global__ void __launch_bounds__(256) transform_binned_kernel(const IactFacets fa, int facet_offset, int G,
                                                               float* world, float* bounds, float* chunk_bounds) {
    const int f = blockIdx.x;
    const M33 R = euler_to_matrix(fa.rotations[3 * f], fa.rotations[3 * f + 1], fa.rotations[3 * f + 2]);
    const V3 pos = ld3(fa.positions + 3 * f);
    const float scale = fa.scale[f];
    const int M = fa.n_samples;
    float4* out = reinterpret_cast<float4*>(world) + ((size_t)(facet_offset + f) * M) * 2;
    __shared__ float red[4][256];
    __shared__ int count[256], cursor[256];
    // local bounding box of the sample points
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const size_t i = ((size_t)f * M + m) * 3;
        const float x = fa.points[i], y = fa.points[i + 1];
        xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
    }
    red[0][threadIdx.x] = xmin; red[1][threadIdx.x] = -xmax; red[2][threadIdx.x] = ymin; red[3][threadIdx.x] = -ymax;
    count[threadIdx.x] = 0;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) for (int k = 0; k < 4; ++k) red[k][threadIdx.x] = fminf(red[k][threadIdx.x], red[k][threadIdx.x + s]);
        __syncthreads();
    }
    xmin = red[0][0]; xmax = -red[1][0]; ymin = red[2][0]; ymax = -red[3][0];
    const float sx = (float)G / fmaxf(xmax - xmin, 1e-20f), sy = (float)G / fmaxf(ymax - ymin, 1e-20f);
    auto cell_of = [&](float x, float y) {
        const int ix = min(G - 1, max(0, (int)((x - xmin) * sx))), iy = min(G - 1, max(0, (int)((y - ymin) * sy)));
        return iy * G + ((iy & 1) ? G - 1 - ix : ix);               // serpentine: consecutive cells are neighbours
    };
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const size_t i = ((size_t)f * M + m) * 3;
        atomicAdd(&count[cell_of(fa.points[i], fa.points[i + 1])], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c < G * G; ++c) { cursor[c] = run; run += count[c]; }
    }
    __syncthreads();
    float maxd2 = 0.f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const size_t i = ((size_t)f * M + m) * 3;
        const V3 pl = ld3(fa.points + i), nl = ld3(fa.normals + i), dl = ld3(fa.delta + i);
        const int slot = atomicAdd(&cursor[cell_of(pl.x, pl.y)], 1);
        const V3 pw = mul(R, pl) + pos;
        V3 nw = mul(R, nl) + scale * mul(R, dl);
        nw = (1.0f / sqrtf(dot(nw, nw))) * nw;
        out[2 * slot] = make_float4(pw.x, pw.y, pw.z, 1.0f / fa.weights[(size_t)f * M + m]);
        out[2 * slot + 1] = make_float4(nw.x, nw.y, nw.z, __int_as_float(m));
        const V3 dd = pw - pos;
        maxd2 = fmaxf(maxd2, dot(dd, dd));
    }
    red[0][threadIdx.x] = maxd2;
    __syncthreads();                                               // also publishes the rows to the block
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[0][threadIdx.x] = fmaxf(red[0][threadIdx.x], red[0][threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float* b = bounds + 4 * (size_t)(facet_offset + f);
        b[0] = pos.x; b[1] = pos.y; b[2] = pos.z;
        b[3] = sqrtf(red[0][0]) * 1.00001f + 1e-6f;
    }
    // per run of 32 rows: bounding sphere (centre = bounding-box centre, radius = farthest row) and normal cone
    // (unit mean normal, largest distance of a row's normal from it: the leg towards optical stage 1 is culled per
    // run from these, iact_cull.cuh leg_masks)
    const int n_chunks = (M + 31) / 32, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < n_chunks; k += blockDim.x >> 5) {
        const int m = min(32 * k + lane, M - 1);
        const float4 a = out[2 * m], nrow = out[2 * m + 1];
        float lo[3] = {a.x, a.y, a.z}, hi[3] = {a.x, a.y, a.z};
        for (int o = 16; o > 0; o >>= 1)
            for (int j = 0; j < 3; ++j) {
                lo[j] = fminf(lo[j], __shfl_xor_sync(0xffffffffu, lo[j], o));
                hi[j] = fmaxf(hi[j], __shfl_xor_sync(0xffffffffu, hi[j], o));
            }
        const V3 c = v3(0.5f * (lo[0] + hi[0]), 0.5f * (lo[1] + hi[1]), 0.5f * (lo[2] + hi[2]));
        const V3 dd = v3(a.x, a.y, a.z) - c;
        float r2 = dot(dd, dd);
        for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
        V3 ns = v3(nrow.x, nrow.y, nrow.z);
        for (int o = 16; o > 0; o >>= 1) {
            ns.x += __shfl_xor_sync(0xffffffffu, ns.x, o); ns.y += __shfl_xor_sync(0xffffffffu, ns.y, o);
            ns.z += __shfl_xor_sync(0xffffffffu, ns.z, o);
        }
        const float nn = dot(ns, ns);
        const V3 nbar = nn > 1e-20f ? (1.0f / sqrtf(nn)) * ns : v3(0.f, 0.f, 0.f);
        const V3 dn = v3(nrow.x, nrow.y, nrow.z) - nbar;
        float e2 = dot(dn, dn);
        for (int o = 16; o > 0; o >>= 1) e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, o));
        if (lane == 0) {
            float* cb = chunk_bounds + IACT_RUN_BOUND_FLOATS * ((size_t)(facet_offset + f) * n_chunks + k);
            cb[0] = c.x; cb[1] = c.y; cb[2] = c.z; cb[3] = sqrtf(r2) * 1.00001f + 1e-6f;
            cb[4] = nbar.x; cb[5] = nbar.y; cb[6] = nbar.z; cb[7] = sqrtf(e2) * 1.00001f + 1e-7f;
        }
    }
}

SurfDev make_surf(const IactSurface* s, bool jit_fold) {
    SurfDev d;
    d.c = (float)s->curvature; d.k = (float)s->conic;
    d.kc2 = jit_fold ? ((1.0f + d.k) * d.c) * d.c : (float)((1.0 + s->conic) * s->curvature * s->curvature);
    d.n_asph = s->n_aspheric;
    for (int i = 0; i < IACT_MAX_ASPH; ++i) d.asph[i] = i < s->n_aspheric ? s->aspheric[i] : 0.f;
    return d;
}

int launch_sample(const uint32_t key[2], int mode, int F, int M, int m0, int Mtot, const IactSurface* surf, int kind, int nv,
                  const float* radii, const float* verts, const float* offsets,
                  float* points, float* normals, float* delta, float* weights, void* stream) {
    IACT_REQUIRE(key && surf && offsets && points && normals && delta && weights, "null pointer");
    IACT_REQUIRE(F >= 0 && M >= 0, "negative size");
    IACT_REQUIRE(m0 >= 0 && Mtot >= 0 && (long long)m0 + M <= (long long)Mtot, "sample rows outside the stream");
    IACT_REQUIRE(mode == IACT_RNG_PARTITIONABLE || mode == IACT_RNG_LEGACY, "bad rng_mode");
    IACT_REQUIRE(surf->n_aspheric >= 0 && surf->n_aspheric <= IACT_MAX_ASPH, "too many aspheric terms");
    if ((long long)F * M == 0) return IACT_OK;
    SampleArgs a;
    a.key.a = key[0]; a.key.b = key[1]; a.mode = mode; a.F = F; a.M = M; a.m0 = m0; a.Mtot = Mtot;
    a.surf = make_surf(surf, false);
    a.kind = kind; a.nv = nv; a.radii = radii; a.verts = verts; a.offsets = offsets;
    a.points = points; a.normals = normals; a.delta = delta; a.weights = weights;
    const long long n = (long long)F * M;
    const unsigned grid = (unsigned)((n + 255) / 256);
    sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    iact_count_launch();
    return iact_check_cuda(cudaGetLastError(), "sample_kernel launch");
}

}  // namespace

extern "C" int iact_sample_disk_group(const uint32_t key[2], int rng_mode, int n_facets, int n_samples,
                                      const IactSurface* surface, const float* radii, const float* offsets,
                                      float* points, float* normals, float* delta, float* weights, void* stream) {
    IACT_REQUIRE(radii, "null radii");
    return launch_sample(key, rng_mode, n_facets, n_samples, 0, n_samples, surface, 0, 0, radii, nullptr, offsets,
                         points, normals, delta, weights, stream);
}

extern "C" int iact_sample_disk_group_rows(const uint32_t key[2], int rng_mode, int n_facets, int n_rows, int first_sample,
                                           int n_samples_total, const IactSurface* surface, const float* radii,
                                           const float* offsets, float* points, float* normals, float* delta, float* weights,
                                           void* stream) {
    IACT_REQUIRE(radii, "null radii");
    return launch_sample(key, rng_mode, n_facets, n_rows, first_sample, n_samples_total, surface, 0, 0, radii, nullptr, offsets,
                         points, normals, delta, weights, stream);
}

extern "C" int iact_sample_polygon_group(const uint32_t key[2], int rng_mode, int n_facets, int n_samples,
                                         const IactSurface* surface, int n_vertices, const float* vertices,
                                         const float* offsets, float* points, float* normals, float* delta,
                                         float* weights, void* stream) {
    IACT_REQUIRE(vertices, "null vertices");
    IACT_REQUIRE(n_vertices >= 3 && n_vertices <= IACT_MAX_POLY, "polygon vertex count out of range");
    return launch_sample(key, rng_mode, n_facets, n_samples, 0, n_samples, surface, 1, n_vertices, nullptr, vertices, offsets,
                         points, normals, delta, weights, stream);
}

extern "C" int iact_sample_polygon_group_rows(const uint32_t key[2], int rng_mode, int n_facets, int n_rows, int first_sample,
                                              int n_samples_total, const IactSurface* surface, int n_vertices,
                                              const float* vertices, const float* offsets, float* points, float* normals,
                                              float* delta, float* weights, void* stream) {
    IACT_REQUIRE(vertices, "null vertices");
    IACT_REQUIRE(n_vertices >= 3 && n_vertices <= IACT_MAX_POLY, "polygon vertex count out of range");
    return launch_sample(key, rng_mode, n_facets, n_rows, first_sample, n_samples_total, surface, 1, n_vertices, nullptr, vertices,
                         offsets, points, normals, delta, weights, stream);
}

static int launch_random(const uint32_t key[2], int mode, int n, int normal, float lo, float hi, float* out, void* stream) {
    IACT_REQUIRE(key && out, "null pointer");
    IACT_REQUIRE(n >= 0, "negative size");
    IACT_REQUIRE(mode == IACT_RNG_PARTITIONABLE || mode == IACT_RNG_LEGACY, "bad rng_mode");
    if (n == 0) return IACT_OK;
    Key k; k.a = key[0]; k.b = key[1];
    random_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(k, mode, n, normal, lo, hi, out);
    iact_count_launch();
    return iact_check_cuda(cudaGetLastError(), "random_kernel launch");
}

extern "C" int iact_random_normal(const uint32_t key[2], int rng_mode, int n, float* out, void* stream) {
    return launch_random(key, rng_mode, n, 1, 0.f, 1.f, out, stream);
}
extern "C" int iact_random_uniform(const uint32_t key[2], int rng_mode, int n, float lo, float hi, float* out, void* stream) {
    return launch_random(key, rng_mode, n, 0, lo, hi, out, stream);
}

extern "C" int iact_transform_to_world_binned(const IactFacets* fa, int facet_offset, int grid_side, float* world, float* bounds,
                                              float* chunk_bounds, void* stream) {
    IACT_REQUIRE(fa && world && bounds && chunk_bounds, "null pointer");
    IACT_REQUIRE(fa->n_facets >= 0 && fa->n_samples >= 0 && facet_offset >= 0, "negative size");
    IACT_REQUIRE(grid_side >= 1 && grid_side <= 16, "grid_side must be in [1, 16]");
    if (fa->n_facets == 0 || fa->n_samples == 0) return IACT_OK;
    IACT_REQUIRE(fa->positions && fa->rotations && fa->scale && fa->points && fa->normals && fa->delta && fa->weights,
                 "null facet table");
    transform_binned_kernel<<<fa->n_facets, 256, 0, (cudaStream_t)stream>>>(*fa, facet_offset, grid_side, world, bounds, chunk_bounds);
    iact_count_launch();
    return iact_check_cuda(cudaGetLastError(), "transform_binned_kernel launch");
}

extern "C" int iact_transform_to_world(const IactFacets* fa, int facet_offset, float* world, float* bounds, void* stream) {
    IACT_REQUIRE(fa && world && bounds, "null pointer");
    IACT_REQUIRE(fa->n_facets >= 0 && fa->n_samples >= 0 && facet_offset >= 0, "negative size");
    if (fa->n_facets == 0) return IACT_OK;
    IACT_REQUIRE(fa->positions && fa->rotations && fa->scale && fa->points && fa->normals && fa->delta && fa->weights,
                 "null facet table");
    transform_kernel<<<fa->n_facets, 256, 0, (cudaStream_t)stream>>>(*fa, facet_offset, world, bounds);
    iact_count_launch();
    return iact_check_cuda(cudaGetLastError(), "transform_kernel launch");
}
