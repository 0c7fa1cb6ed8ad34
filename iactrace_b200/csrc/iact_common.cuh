// Shared device helpers for the iactrace_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/iactrace_b200.h"

// ---------------------------------------------------------------- host-side plumbing
void iact_set_error(const char* fmt, ...);
int  iact_check_cuda(cudaError_t e, const char* what);
void iact_count_launch(int n = 1);

#define IACT_CUDA(call)                                                   \
    do { int _rc = iact_check_cuda((call), #call); if (_rc) return _rc; } while (0)
#define IACT_REQUIRE(cond, msg)                                           \
    do { if (!(cond)) { iact_set_error("%s: %s", __func__, msg); return IACT_ERR_ARG; } } while (0)

// ---------------------------------------------------------------- small vector type
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }

// 3x3 rotation, row-major.
struct M33 { float m[9]; };
__device__ __forceinline__ V3 mul(const M33& R, V3 a) {     // R a
    return v3(R.m[0] * a.x + R.m[1] * a.y + R.m[2] * a.z,
              R.m[3] * a.x + R.m[4] * a.y + R.m[5] * a.z,
              R.m[6] * a.x + R.m[7] * a.y + R.m[8] * a.z);
}
__device__ __forceinline__ V3 mulT(const M33& R, V3 a) {    // R^T a
    return v3(R.m[0] * a.x + R.m[3] * a.y + R.m[6] * a.z,
              R.m[1] * a.x + R.m[4] * a.y + R.m[7] * a.z,
              R.m[2] * a.x + R.m[5] * a.y + R.m[8] * a.z);
}

// euler_to_matrix (reference core/transforms.py:72-106): degrees, R = Rz(rot) Ry(tilt) Rx(tip).
__device__ __forceinline__ M33 euler_to_matrix(float tip, float tilt, float rot) {
    const float D2R = 0.017453292519943295f;
    float sx, cx, sy, cy, sz, cz;
    sincosf(tip * D2R, &sx, &cx);
    sincosf(tilt * D2R, &sy, &cy);
    sincosf(rot * D2R, &sz, &cz);
    // Ry Rx
    float a00 = cy, a01 = sy * sx, a02 = sy * cx;
    float a10 = 0.f, a11 = cx, a12 = -sx;
    float a20 = -sy, a21 = cy * sx, a22 = cy * cx;
    M33 R;
    R.m[0] = cz * a00 - sz * a10; R.m[1] = cz * a01 - sz * a11; R.m[2] = cz * a02 - sz * a12;
    R.m[3] = sz * a00 + cz * a10; R.m[4] = sz * a01 + cz * a11; R.m[5] = sz * a02 + cz * a12;
    R.m[6] = a20;                 R.m[7] = a21;                 R.m[8] = a22;
    return R;
}

// ---------------------------------------------------------------- threefry2x32 (JAX PRNG)
struct Key { uint32_t a, b; };

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

__device__ __forceinline__ Key threefry2x32(Key k, uint32_t c0, uint32_t c1) {
    uint32_t ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
    uint32_t x0 = c0 + ks0, x1 = c1 + ks1;
#define TF_ROUND(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks1; x1 += ks2 + 1u;
    TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
    x0 += ks2; x1 += ks0 + 2u;
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks0; x1 += ks1 + 3u;
    TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
    x0 += ks1; x1 += ks2 + 4u;
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks2; x1 += ks0 + 5u;
#undef TF_ROUND
    Key r; r.a = x0; r.b = x1; return r;
}

// element i of legacy `threefry_2x32(key, iota(n))`: counts cut in halves, outputs concatenated.
__device__ __forceinline__ uint32_t legacy_bits_at(Key k, uint32_t i, uint32_t n) {
    uint32_t h = (n + 1u) >> 1;                        // padded half length
    if (i < h) {
        uint32_t c1 = h + i; if (c1 >= n) c1 = 0u;     // the pad element is a zero count
        return threefry2x32(k, i, c1).a;
    }
    return threefry2x32(k, i - h, i).b;
}

// jax.random.split(key, num)[i]
__device__ __forceinline__ Key rng_split(Key k, uint32_t i, uint32_t num, int mode) {
    if (mode == IACT_RNG_PARTITIONABLE) return threefry2x32(k, 0u, i);
    Key r; r.a = legacy_bits_at(k, 2u * i, 2u * num); r.b = legacy_bits_at(k, 2u * i + 1u, 2u * num); return r;
}

// 32 random bits: element i of random_bits(key, (n,))
__device__ __forceinline__ uint32_t rng_bits(Key k, uint32_t i, uint32_t n, int mode) {
    if (mode == IACT_RNG_PARTITIONABLE) { Key r = threefry2x32(k, 0u, i); return r.a ^ r.b; }
    return legacy_bits_at(k, i, n);
}

// jax.random.uniform(key,(n,),f32,lo,hi)[i]
__device__ __forceinline__ float rng_uniform(Key k, uint32_t i, uint32_t n, int mode, float lo, float hi) {
    uint32_t b = rng_bits(k, i, n, mode);
    float f = __uint_as_float((b >> 9) | 0x3F800000u) - 1.0f;
    return fmaxf(lo, __fadd_rn(__fmul_rn(f, hi - lo), lo));
}

// XLA float32 erf_inv (Giles' polynomial)
__device__ __forceinline__ float erf_inv_f32(float x) {
    float w = -log1pf(-(x * x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w;  p = -0.00125372503f + p * w; p = -0.00417768164f + p * w;
        p = 0.246640727f + p * w;    p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = 0.000100950558f + p * w; p = 0.00134934322f + p * w;  p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w;  p = -0.0076224613f + p * w;  p = 0.00943887047f + p * w;
        p = 1.00167406f + p * w;     p = 2.83297682f + p * w;
    }
    float r = p * x;
    return fabsf(x) == 1.0f ? x * INFINITY : r;
}

// jax.random.normal(key,(n,))[i]
__device__ __forceinline__ float rng_normal(Key k, uint32_t i, uint32_t n, int mode) {
    const float lo = -0.99999994f;   // nextafter(-1, 0)
    float u = rng_uniform(k, i, n, mode, lo, 1.0f);
    return 1.41421356f * erf_inv_f32(u);
}

// ---------------------------------------------------------------- aspheric surface
struct SurfDev { float c, k, kc2; int n_asph; float asph[IACT_MAX_ASPH]; };

// _sag_raw (core/surfaces.py:25-39): z = c r2 / (1 + sqrt(1 - (1+k) c^2 r2)) + sum a_i (r2)^(2i+2)
__device__ __forceinline__ float sag_raw(const SurfDev& s, float x, float y) {
    float r2 = x * x + y * y;
    float z = r2 * s.c / (1.0f + sqrtf(1.0f - s.kc2 * r2));
    if (s.n_asph > 0) {
        float r4 = r2 * r2, p = r4;
        for (int i = 0; i < s.n_asph; ++i) { z += s.asph[i] * p; p *= r4; }
    }
    return z;
}
// d sag_raw / d(r2): conic part simplifies to c / (2 sqrt(1 - (1+k) c^2 r2))
__device__ __forceinline__ float dsag_dr2(const SurfDev& s, float r2) {
    float d = 0.5f * s.c * rsqrtf(1.0f - s.kc2 * r2);
    if (s.n_asph > 0) {
        float r4 = r2 * r2, p = r2;                    // derivative of r2^(2i+2) = (2i+2) r2^(2i+1)
        for (int i = 0; i < s.n_asph; ++i) { d += s.asph[i] * (float)(2 * i + 2) * p; p *= r4; }
    }
    return d;
}
