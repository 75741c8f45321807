// vec.cuh — float3/float2 algebra for the sm_100a kernels.
// Operation order follows the reference's common/basic_types.h (dot :2742, cross :2751-2757,
// v / s == v * (1 / s) :2564-2570, matrix * v = row dots :4264-4271).  dot / cross / length / matrix
// products use EXPLICIT fmaf in the same association order as oracle/vecmath.h; everything else is built
// with -fmad=false (no implicit contraction), so kernels round exactly like the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "detmath.h"

#define GFX_D __device__ __forceinline__
#define GFX_HD __host__ __device__ __forceinline__

namespace gfx {

struct f3 {
    float x, y, z;
    GFX_HD f3() {}
    GFX_HD explicit f3(float v) : x(v), y(v), z(v) {}
    GFX_HD f3(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
};
struct f2 {
    float x, y;
    GFX_HD f2() {}
    GFX_HD f2(float _x, float _y) : x(_x), y(_y) {}
};

GFX_HD f3 operator+(const f3 &a, const f3 &b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
GFX_HD f3 operator-(const f3 &a, const f3 &b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
GFX_HD f3 operator-(const f3 &a) { return f3(-a.x, -a.y, -a.z); }
GFX_HD f3 operator*(const f3 &a, const f3 &b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
GFX_HD f3 operator*(float s, const f3 &a) { return f3(s * a.x, s * a.y, s * a.z); }
GFX_HD f3 operator*(const f3 &a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
GFX_HD f3 operator/(const f3 &a, float s) { const float rr = 1 / s; return f3(a.x * rr, a.y * rr, a.z * rr); }
GFX_HD f3 operator/(const f3 &a, const f3 &b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
GFX_HD f3 &operator+=(f3 &a, const f3 &b) { a = a + b; return a; }
GFX_HD f3 &operator*=(f3 &a, const f3 &b) { a = a * b; return a; }
GFX_HD f3 &operator*=(f3 &a, float s) { a = a * s; return a; }
GFX_HD f3 &operator/=(f3 &a, float s) { a = a / s; return a; }
GFX_HD f2 operator+(const f2 &a, const f2 &b) { return f2(a.x + b.x, a.y + b.y); }
GFX_HD f2 operator-(const f2 &a, const f2 &b) { return f2(a.x - b.x, a.y - b.y); }
GFX_HD f2 operator*(float s, const f2 &a) { return f2(s * a.x, s * a.y); }
GFX_HD f2 operator*(const f2 &a, const f2 &b) { return f2(a.x * b.x, a.y * b.y); }

GFX_HD float dot(const f3 &a, const f3 &b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
GFX_HD f3 cross(const f3 &a, const f3 &b) {
    return f3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
GFX_HD float sqLength(const f3 &v) { return fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x)); }
GFX_HD float length(const f3 &v) { return sqrtf(sqLength(v)); }
GFX_HD f3 normalize(const f3 &v) { const float l = length(v); return v / l; }
GFX_HD f3 min3(const f3 &a, const f3 &b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
GFX_HD f3 max3(const f3 &a, const f3 &b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
GFX_HD bool allFinite(const f3 &v) { return isfinite(v.x) && isfinite(v.y) && isfinite(v.z); }

GFX_HD float pow2f(float x) { return x * x; }
GFX_HD float pow4f(float x) { return pow2f(pow2f(x)); }
GFX_HD float pow5f(float x) { return x * pow4f(x); }
GFX_HD float lerpf(float v0, float v1, float t) { return (1 - t) * v0 + t * v1; }
GFX_HD f3 lerp3(const f3 &v0, const f3 &v1, float t) { return (1 - t) * v0 + t * v1; }
GFX_HD float sRGB_calcLuminance(const f3 &v) { return 0.2126729f * v.x + 0.7151522f * v.y + 0.0721750f * v.z; }

// row-major 3x4 affine / 3x3 linear maps
GFX_HD f3 xfmPoint(const float* m, const f3 &p) {
    return f3(fmaf(m[2], p.z, fmaf(m[1], p.y, m[0] * p.x)) + m[3],
              fmaf(m[6], p.z, fmaf(m[5], p.y, m[4] * p.x)) + m[7],
              fmaf(m[10], p.z, fmaf(m[9], p.y, m[8] * p.x)) + m[11]);
}
GFX_HD f3 xfmVector(const float* m, const f3 &v) {
    return f3(fmaf(m[2], v.z, fmaf(m[1], v.y, m[0] * v.x)),
              fmaf(m[6], v.z, fmaf(m[5], v.y, m[4] * v.x)),
              fmaf(m[10], v.z, fmaf(m[9], v.y, m[8] * v.x)));
}
GFX_HD f3 mul3x3(const float* m, const f3 &v) {
    return f3(fmaf(m[2], v.z, fmaf(m[1], v.y, m[0] * v.x)),
              fmaf(m[5], v.z, fmaf(m[4], v.y, m[3] * v.x)),
              fmaf(m[8], v.z, fmaf(m[7], v.y, m[6] * v.x)));
}

} // namespace gfx
