/*
 * iactrace_b200 -- C ABI of the B200-native Monte-Carlo ray-tracing hot path.
 *
 * The reference (GerritRo/iactrace v0.4.0) has NO native / FFI layer: its
 * boundary is three jit'd Python functions over an Equinox pytree
 * (iactrace/core/render.py:174 render, :223 render_debug,
 * :271 render_response_matrix) plus the load-time sampler
 * (iactrace/core/integrators.py:68 MCIntegrator.sample_group).  This header
 * is the C-ABI an FFI for that path would bind: every entry point names the
 * reference function it replaces.  Binding stubs: INTEGRATION.md.
 *
 * Conventions
 *  - every pointer marked "device" is CUDA device memory owned by the caller;
 *    all arrays are dense, row-major, float32 / int32 / uint32;
 *  - every call is stream-ordered on the cudaStream_t passed as `void* stream`
 *    (NULL = legacy default stream) and never synchronises;
 *  - return value: 0 on success, non-zero error code otherwise;
 *    iact_last_error() returns a thread-local message;
 *  - nothing here falls back to the CPU: without a CUDA device every compute
 *    call fails with IACT_ERR_CUDA.
 */
#ifndef IACTRACE_B200_H
#define IACTRACE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IACT_OK             0
#define IACT_ERR_ARG        1
#define IACT_ERR_CUDA       2
#define IACT_ERR_UNSUPPORTED 3

#define IACT_MAX_ASPH    8   /* aspheric polynomial terms per surface          */
#define IACT_MAX_STAGES  4   /* optical stages >= 1 (secondary, tertiary, ...) */
#define IACT_MAX_POLY   16   /* polygon aperture vertices                      */
#define IACT_MIRROR_REC 24   /* floats per stage>=1 mirror record (below)      */
#define IACT_RUN_BOUND_FLOATS 8 /* floats per 32-row run in IactScene.chunk_bounds */
#define IACT_MAX_TAPS   64   /* soft-sensor neighbourhood size                 */

/* JAX key-derivation mode: jax_threefry_partitionable. */
#define IACT_RNG_PARTITIONABLE 0   /* JAX >= 0.5.0 default */
#define IACT_RNG_LEGACY        1   /* JAX <  0.5.0 default */

#define IACT_SOURCE_POINT    0     /* render.py:129-131 */
#define IACT_SOURCE_PARALLEL 1     /* render.py:132-133 (any string != 'point') */

#define IACT_SENSOR_SQUARE      0  /* sensors/square.py:28    SquareSensor                   */
#define IACT_SENSOR_HEX         1  /* sensors/hexagonal.py:104 HexagonalSensor               */
#define IACT_SENSOR_SOFT_SQUARE 2  /* sensors/square.py:94    DifferentiableSquareSensor     */
#define IACT_SENSOR_SOFT_HEX    3  /* sensors/hexagonal.py:197 DifferentiableHexagonalSensor */

/* Aspheric surface shared by one mirror group (core/surfaces.py:8-23).
 * Doubles: the reference holds these as Python floats and folds them. */
typedef struct IactSurface {
    double  curvature;
    double  conic;
    int32_t n_aspheric;
    float   aspheric[IACT_MAX_ASPH];
} IactSurface;

/* One optical stage >= 1: all mirrors of all of its groups, flattened in
 * group order then mirror order (render.py:61-71 strict '<' keeps the first
 * minimum).  Record layout, IACT_MIRROR_REC floats per mirror:
 *   [0..2] position  [3..5] euler deg (tip,tilt,rot)  [6..7] offset
 *   [8] curvature [9] conic [10] n_aspheric [11..18] aspheric[8]
 *   [19] aperture kind (0 disk, 1 polygon) [20] disk radius
 *   [21] polygon vertex count [22] first vertex index into `verts` [23] pad  */
typedef struct IactMirrorStage {
    int32_t      n_mirrors;
    const float* records;   /* device (n_mirrors, IACT_MIRROR_REC) */
    const float* verts;     /* device (sum_vertices, 2) or NULL    */
} IactMirrorStage;

/* Sensor statics.  Doubles are the reference's Python-float static fields
 * (square.py:36-41, hexagonal.py:112-119); they are folded to float32 exactly
 * where the reference's weak typing folds them. */
typedef struct IactSensor {
    int32_t kind;
    float   position[3];
    float   euler[3];          /* degrees; R = euler_to_matrix (transforms.py:72) */
    /* square */
    int32_t width, height;
    double  x0, y0, dx, dy;
    double  edge_width;        /* both hard sensors */
    /* hexagonal */
    double  hex_size, hex_inradius, grid_rotation, grid_offset[2];
    int32_t q_min, r_min, table_q, table_r, n_pixels;
    const int32_t* lookup;     /* device (table_q, table_r), -1 = no pixel */
    /* soft sensors */
    double  sigma;
    int32_t kernel_size;
    /* hexagonal, optional: radius (about grid_offset) of a circle that contains every hexagon, or 0 = unknown.
     * Hits outside it belong to no pixel (the lookup of hexagonal.py:155-172 would return -1); the kernel uses it
     * to skip the cube rounding for rays that miss the camera. */
    double  hex_outer_radius;
} IactSensor;

/* Everything a render needs.  `world`/`bounds` come from iact_transform_to_world. */
typedef struct IactScene {
    int32_t      n_facets, n_samples;
    const float* world;    /* device (F, M, 8): px,py,pz,1/weight, nx,ny,nz,(original sample index as int bits) */
    const float* bounds;   /* device (F, 4): bounding sphere of the facet's world points */
    const float* chunk_bounds; /* device (F, ceil(M/32), IACT_RUN_BOUND_FLOATS) or NULL: per run of 32 consecutive table
                                  rows its bounding sphere (cx,cy,cz,R) and normal cone (unit mean normal, largest
                                  |n - mean|), as written by iact_transform_to_world_binned (enables per-run culling) */
    /* obstruction groups in the reference's fixed order (obstructions.py:258-278) */
    int32_t n_cyl;  const float *cyl_p1, *cyl_p2, *cyl_r;       /* (K,3)(K,3)(K,)   */
    int32_t n_box;  const float *box_p1, *box_p2;               /* (K,3)(K,3)       */
    int32_t n_sph;  const float *sph_c,  *sph_r;                /* (K,3)(K,)        */
    int32_t n_obox; const float *obox_c, *obox_h, *obox_R;      /* (K,3)(K,3)(K,3,3)*/
    int32_t n_tri;  const float *tri_v0, *tri_v1, *tri_v2;      /* (K,3) x3         */
    int32_t n_stages;                                           /* stages >= 1      */
    IactMirrorStage stages[IACT_MAX_STAGES];
    IactSensor sensor;
    int32_t cull;          /* 1 = conservative beam/obstruction culling + early Newton exit (default; results identical), 0 = the reference's literal brute force */
} IactScene;

/* Stage-0 facet parameters in the LOCAL frame: the differentiable inputs. */
typedef struct IactFacets {
    int32_t      n_facets, n_samples;
    const float* positions;   /* device (F,3)                                  */
    const float* rotations;   /* device (F,3) euler degrees                    */
    const float* scale;       /* device (F,)  perturbation_scale (rad)         */
    const float* points;      /* device (F,M,3) local sample points            */
    const float* normals;     /* device (F,M,3)                                */
    const float* delta;       /* device (F,M,3) perturbation_delta             */
    const float* weights;     /* device (F,M)                                  */
} IactFacets;

/* Gradient outputs of iact_render_vjp; any pointer may be NULL (= not wanted).
 * All are ACCUMULATED into (caller zeroes them). */
typedef struct IactGrads {
    float* positions;      /* device (F,3) */
    float* rotations;      /* device (F,3) */
    float* scale;          /* device (F,)  */
    float* weights;        /* device (F,M) */
    float* values;         /* device (S,)  */
    float* sources;        /* device (S,3) */
    float* sensor_position;/* device (3,)  */
    float* sensor_euler;   /* device (3,)  */
    float* stage_positions;/* device (N2,3): mirrors of optical stages >= 1, flat in IactScene.stages order */
    float* stage_rotations;/* device (N2,3): Euler degrees                                                  */
    /* surface fits (core/surfaces.py:25-65) */
    float* points;         /* device (F,M,3): d/d(local sample point)                        -- one atomic per ray   */
    float* nq;             /* device (F,M,3): d/d(local normal + perturbation_scale * delta)  -- one atomic per ray   */
    float* stage_surface;  /* device (N2,4): d/d(curvature, conic, offset x, offset y) of each stage >= 1 mirror,
                              through the implicit Newton root (intersections.py:290-367)                              */
} IactGrads;

const char* iact_last_error(void);
int  iact_version(void);
/* number of visible CUDA devices, or -1 (with iact_last_error set) */
int  iact_device_count(void);

/* MCIntegrator._sample_disk_group (core/integrators.py:97-140): fills local
 * points/normals/delta (F,M,3) and weights (F,M) from the same threefry key
 * tree as the reference: mkey = split(key,F)[f]; ks,kp = split(mkey); ...   */
int iact_sample_disk_group(const uint32_t key[2], int rng_mode, int n_facets, int n_samples,
                           const IactSurface* surface,
                           const float* radii /*device (F,)*/, const float* offsets /*device (F,2)*/,
                           float* points, float* normals, float* delta, float* weights, void* stream);

/* MCIntegrator._sample_polygon_group (core/integrators.py:142-188). */
int iact_sample_polygon_group(const uint32_t key[2], int rng_mode, int n_facets, int n_samples,
                              const IactSurface* surface, int n_vertices,
                              const float* vertices /*device (F,nv,2)*/, const float* offsets /*device (F,2)*/,
                              float* points, float* normals, float* delta, float* weights, void* stream);

/* The same samplers for a WINDOW of the stream: rows [0, n_rows) of the outputs receive samples
 * first_sample .. first_sample + n_rows - 1 of the n_samples_total the reference would draw per facet
 * (threefry is counter based, so any sample is regenerated in isolation, in both key-derivation modes; weights carry
 * n_samples_total).  This is what lets `render` stream a large MCIntegrator(n_samples) through L2-sized chunks
 * instead of holding (and re-reading from HBM) an (F, n_samples, 8) table: core/integrators.py:97-188. */
int iact_sample_disk_group_rows(const uint32_t key[2], int rng_mode, int n_facets, int n_rows, int first_sample,
                                int n_samples_total, const IactSurface* surface, const float* radii, const float* offsets,
                                float* points, float* normals, float* delta, float* weights, void* stream);
int iact_sample_polygon_group_rows(const uint32_t key[2], int rng_mode, int n_facets, int n_rows, int first_sample,
                                   int n_samples_total, const IactSurface* surface, int n_vertices, const float* vertices,
                                   const float* offsets, float* points, float* normals, float* delta, float* weights,
                                   void* stream);

/* jax.random.normal(key,(n,)) / uniform(key,(n,),lo,hi) on the device: used by the
 * host-side parameter edits (telescope/operations.py:186-188,220) and tests. */
int iact_random_normal(const uint32_t key[2], int rng_mode, int n, float* out, void* stream);
int iact_random_uniform(const uint32_t key[2], int rng_mode, int n, float lo, float hi, float* out, void* stream);

/* MirrorGroup.transform_to_world (telescope/mirrors.py:64-79), writing rows
 * [facet_offset, facet_offset + n_facets) of the packed world table + bounds. */
int iact_transform_to_world(const IactFacets* facets, int facet_offset,
                            float* world /*device (Ftot,M,8): px,py,pz,1/w, nx,ny,nz,0*/, float* bounds /*device (Ftot,4)*/, void* stream);

/* Same transform, but the rows of each facet are written in spatially binned order (counting sort of
 * the samples into a grid_side x grid_side grid of cells over the facet, serpentine cell order), so that
 * every run of 32 consecutive rows covers a small patch; chunk_bounds (Ftot, ceil(M/32), IACT_RUN_BOUND_FLOATS)
 * receives the bounding sphere and the normal cone of each run.  Row slot [7] holds the original sample index (int bits), which
 * iact_render_debug uses to keep the reference's output order.  Summation order aside, rendering from a
 * binned table is identical to rendering from the plain one. */
int iact_transform_to_world_binned(const IactFacets* facets, int facet_offset, int grid_side,
                                   float* world, float* bounds, float* chunk_bounds, void* stream);

/* render (core/render.py:174-220).  out_image: device (H,W) or (P,), OVERWRITTEN. */
int iact_render(const IactScene* scene, const float* sources /*device (S,3)*/, const float* values /*device (S,)*/,
                int n_sources, int source_type, float* out_image, void* stream);

/* render_response_matrix (core/render.py:271-324).  out: device (S, n_pixels), OVERWRITTEN. */
int iact_response_matrix(const IactScene* scene, const float* sources, const float* values,
                         int n_sources, int source_type, float* out_matrix, void* stream);

/* render_debug (core/render.py:223-268).  out_xy: device (F*S*M,2), out_val: (F*S*M,),
 * facet-major, then source, then sample.  out_pixel (optional, may be NULL): int32 flat
 * pixel index the hard sensor assigns, -1 = rejected. */
int iact_render_debug(const IactScene* scene, const float* sources, const float* values,
                      int n_sources, int source_type, float* out_xy, float* out_val, int32_t* out_pixel,
                      void* stream);

/* sensor.accumulate(x, y, values) (sensors/square.py:66,144; sensors/hexagonal.py:174,264) on n
 * free-standing hits.  out: device, sensor-shaped, OVERWRITTEN. */
int iact_accumulate(const IactSensor* sensor, const float* x, const float* y, const float* values,
                    long long n, float* out, void* stream);

/* Vector-Jacobian product of iact_render w.r.t. the stage-0 facet parameters, the
 * sources/values and the sensor pose (what jax.grad of render.py:174 would give;
 * SURVEY.md section 3.5).  cotangent: device, shaped like the image. */
int iact_render_vjp(const IactScene* scene, const IactFacets* facets,
                    const float* sources, const float* values, int n_sources, int source_type,
                    const float* cotangent, const IactGrads* grads, void* stream);

/* Diagnostics for the roofline: candidate-list lengths of the conservative obstruction culling,
 * over all rays.  out4: device uint64[4] = {cylinder tests executed (summed over rays), tests of other
 * primitives, number of rays, sum over (facet, source) pairs of the facet-level (level-1) list length}. */
int iact_cull_stats(const IactScene* scene, const float* sources, int n_sources, int source_type,
                    unsigned long long* out4, void* stream);

/* Roofline probes (bench.py): dependent-free FP32 FMA throughput in FLOP/s written to
 * *out_flops, shared-memory float atomicAdd throughput in atomics/s to *out_atomics. */
int iact_probe_fp32(int iters, double* out_flops, void* stream);
int iact_probe_smem_atomics(int iters, int n_distinct, double* out_atomics, void* stream);

/* Work decomposition of the trace / VJP kernels, host arithmetic only (no device needed): how n_facets x n_samples
 * x n_sources rays are cut into the units the warps pull from the work queue.
 * kind 0 = render / render_debug: out6 = {facets per unit, facet runs, sample parts, rows per part, 0, 0};
 *                                 unit u -> source u / (runs * parts), run (u / parts) % runs, part u % parts.
 * kind 1 = render_vjp:            out6 = {sources per unit, source runs, sample parts, rows per part, 0, 0};
 *                                 unit u -> source run u % sruns, part (u / sruns) % parts, facet u / (sruns * parts).
 * *n_units receives the unit count.  Used by the CPU tests to check that every ray is covered exactly once. */
int iact_work_plan(int kind, int n_facets, int n_samples, int n_sources, int has_obstructions, long long resident_warps,
                   int* out6, long long* n_units);

/* Number of kernel launches issued by this library in this process (bench "gpu_launches"). */
long long iact_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* IACTRACE_B200_H */
