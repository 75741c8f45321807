// oracle/vecmath.h — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
//
// Minimal float3/RGB/AABB algebra with the *operation order* of the reference's
// common/basic_types.h so that expressions restated from the reference round identically:
//   dot   = a.x*b.x + a.y*b.y + a.z*b.z                (basic_types.h:2742-2747 region)
//   cross = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)   (basic_types.h:2751-2757)
//   normalize(v) = v * (1 / sqrt(dot(v,v)))  (operator/=(F) multiplies by the reciprocal,
//                  basic_types.h:2564-2570, 2590-2593)
//   Matrix * v   = row-dot products                     (basic_types.h:4264-4271, 4760-4768)
// dot / cross / length / matrix-vector products are spelled with EXPLICIT fused multiply-adds in a fixed
// association order (z-term outermost); csrc/vec.cuh spells them identically with fmaf, so host and device
// round the same way.  (The reference itself is built with nvcc's default FMA contraction + fast-math, so
// no particular contraction pattern is canonical.)  Everything else is built with -ffp-contract=off
// (no implicit contraction) — see oracle/Makefile.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <limits>

namespace orc {

struct float3 {
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    explicit float3(float v) : x(v), y(v), z(v) {}
    float3(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float _x, float _y) : x(_x), y(_y) {}
};

inline float3 operator+(const float3 &a, const float3 &b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(const float3 &a, const float3 &b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(const float3 &a) { return float3(-a.x, -a.y, -a.z); }
inline float3 operator*(const float3 &a, const float3 &b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator*(float s, const float3 &a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator*(const float3 &a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
// vector / scalar multiplies by the reciprocal (basic_types.h:2564-2570, 5203-5209)
inline float3 operator/(const float3 &a, float s) { const float rr = 1 / s; return float3(a.x * rr, a.y * rr, a.z * rr); }
inline float3 operator/(const float3 &a, const float3 &b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float3 &operator+=(float3 &a, const float3 &b) { a = a + b; return a; }
inline float3 &operator*=(float3 &a, const float3 &b) { a = a * b; return a; }
inline float3 &operator*=(float3 &a, float s) { a = a * s; return a; }
inline float3 &operator/=(float3 &a, float s) { a = a / s; return a; }
inline float2 operator+(const float2 &a, const float2 &b) { return float2(a.x + b.x, a.y + b.y); }
inline float2 operator-(const float2 &a, const float2 &b) { return float2(a.x - b.x, a.y - b.y); }
inline float2 operator*(float s, const float2 &a) { return float2(s * a.x, s * a.y); }
inline float2 operator*(const float2 &a, const float2 &b) { return float2(a.x * b.x, a.y * b.y); }

inline float dot(const float3 &a, const float3 &b) { return std::fmaf(a.z, b.z, std::fmaf(a.y, b.y, a.x * b.x)); }
inline float3 cross(const float3 &a, const float3 &b) {
    return float3(std::fmaf(a.y, b.z, -(a.z * b.y)), std::fmaf(a.z, b.x, -(a.x * b.z)), std::fmaf(a.x, b.y, -(a.y * b.x)));
}
inline float sqLength(const float3 &v) { return std::fmaf(v.z, v.z, std::fmaf(v.y, v.y, v.x * v.x)); }
inline float length(const float3 &v) { return std::sqrt(sqLength(v)); }
inline float3 normalize(const float3 &v) { const float l = length(v); return v / l; }
inline float3 min3(const float3 &a, const float3 &b) {
    return float3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z));
}
inline float3 max3(const float3 &a, const float3 &b) {
    return float3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z));
}
inline bool allFinite(const float3 &v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }
inline float3 safeDivide(const float3 &a, const float3 &b) { // basic_types.h:2577-2582
    return float3(b.x != 0 ? a.x / b.x : 0.0f, b.y != 0 ? a.y / b.y : 0.0f, b.z != 0 ? a.z / b.z : 0.0f);
}

template <typename T> inline T pow2(T x) { return x * x; }
template <typename T> inline T pow4(T x) { return pow2(pow2(x)); }
template <typename T> inline T pow5(T x) { return x * pow4(x); } // basic_types.h:255-258
inline float lerpf(float v0, float v1, float t) { return (1 - t) * v0 + t * v1; } // basic_types.h:260-263
inline float3 lerp3(const float3 &v0, const float3 &v1, float t) { return (1 - t) * v0 + t * v1; }
inline float sRGB_calcLuminance(const float3 &v) { // basic_types.h:5420-5423
    return 0.2126729f * v.x + 0.7151522f * v.y + 0.0721750f * v.z;
}

// 3x4 row-major affine transform: p' = dot(row, (p,1)) with the Vector4D dot order
// a.x*b.x + a.y*b.y + a.z*b.z + a.w*b.w   (basic_types.h:3349-3352, 4760-4768).
struct Affine {
    float m[12];
    float3 point(const float3 &p) const {
        return float3(
            std::fmaf(m[2], p.z, std::fmaf(m[1], p.y, m[0] * p.x)) + m[3],
            std::fmaf(m[6], p.z, std::fmaf(m[5], p.y, m[4] * p.x)) + m[7],
            std::fmaf(m[10], p.z, std::fmaf(m[9], p.y, m[8] * p.x)) + m[11]);
    }
    float3 vector(const float3 &v) const {
        return float3(
            std::fmaf(m[2], v.z, std::fmaf(m[1], v.y, m[0] * v.x)),
            std::fmaf(m[6], v.z, std::fmaf(m[5], v.y, m[4] * v.x)),
            std::fmaf(m[10], v.z, std::fmaf(m[9], v.y, m[8] * v.x)));
    }
};
struct Mat3 { // row-major
    float m[9];
    float3 mul(const float3 &v) const {
        return float3(
            std::fmaf(m[2], v.z, std::fmaf(m[1], v.y, m[0] * v.x)),
            std::fmaf(m[5], v.z, std::fmaf(m[4], v.y, m[3] * v.x)),
            std::fmaf(m[8], v.z, std::fmaf(m[7], v.y, m[6] * v.x)));
    }
};

struct AABB { // basic_types.h:3358-3465
    float3 minP, maxP;
    AABB() : minP(std::numeric_limits<float>::infinity()), maxP(-std::numeric_limits<float>::infinity()) {}
    AABB(const float3 &a, const float3 &b) : minP(a), maxP(b) {}
    AABB &unify(const float3 &p) { minP = min3(minP, p); maxP = max3(maxP, p); return *this; }
    AABB &unify(const AABB &b) { minP = min3(minP, b.minP); maxP = max3(maxP, b.maxP); return *this; }
    AABB &intersect(const AABB &b) { minP = max3(minP, b.minP); maxP = min3(maxP, b.maxP); return *this; }
    float3 getCenter() const { return 0.5f * (minP + maxP); }
    float calcHalfSurfaceArea() const {
        const float3 d = maxP - minP;
        return d.x * d.y + d.y * d.z + d.z * d.x;
    }
    float3 normalize(const float3 &p) const { return safeDivide(p - minP, maxP - minP); }
    bool isValid() const {
        const float3 d = maxP - minP;
        return d.x >= 0.0f && d.y >= 0.0f && d.z >= 0.0f;
    }
    // basic_types.h:3450-3465
    bool intersectRay(const float3 &org, const float3 &dir, float distMin, float distMax,
                      float* hitDistMin, float* hitDistMax) const {
        if (!isValid())
            return false;
        const float3 invRayDir = float3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
        const float3 tNear = (minP - org) * invRayDir;
        const float3 tFar = (maxP - org) * invRayDir;
        const float3 near = min3(tNear, tFar);
        const float3 far = max3(tNear, tFar);
        *hitDistMin = std::fmax(std::fmax(near.x, near.y), near.z);
        *hitDistMax = std::fmin(std::fmin(far.x, far.y), far.z);
        *hitDistMin = std::fmax(*hitDistMin, distMin);
        *hitDistMax = std::fmin(*hitDistMax, distMax);
        return *hitDistMin <= *hitDistMax && *hitDistMax > 0.0f;
    }
};
inline AABB unify(const AABB &a, const AABB &b) { AABB r = a; r.unify(b); return r; }
inline AABB unify(const AABB &a, const float3 &p) { AABB r = a; r.unify(p); return r; }
inline AABB intersect(const AABB &a, const AABB &b) { AABB r = a; r.intersect(b); return r; }

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t floatToOrderedUInt(float f) { // basic_types.h:429-436
    const uint32_t ui = f2u(f);
    return ui ^ (ui < 0x80000000u ? 0x80000000u : 0xFFFFFFFFu);
}
inline uint32_t popcnt(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
inline uint32_t nextPowOf2Exponent(uint32_t x) { // basic_types.h:344-348
    if (x == 0) return 0;
    return x == 1 ? 0 : 32 - (uint32_t)__builtin_clz(x - 1);
}
inline uint32_t nextPowerOf2(uint32_t x) { // basic_types.h:370-374
    if (x == 0) return 0;
    return 1u << nextPowOf2Exponent(x);
}

} // namespace orc
