// TEST INFRASTRUCTURE -- shim of MathLib's "ml.h" for compiling the reference's OWN host code (oracle/Makefile.ref).
//
// /root/reference/Source/*.cpp include "ml.h" / "ml.hlsli" (Source/InstanceImpl.h:21-22) from NVIDIA-RTX/MathLib, which the
// reference fetches from the network at configure time (CMakeLists.txt:120-129) and which is therefore absent here.  This file
// supplies exactly the subset the host code touches, written from the call sites (cited per item); nothing here is copied from
// MathLib.  Everything the reference computes with these pieces ALONE (rotators, frustum, matrix inverses) is "shim-dependent":
// tests/test_reference_scheduler.py lists those constant-buffer fields separately; every other byte the reference library
// produces -- dispatch order, names, pipelines, resources, ping-pong, clears, grids, and all constants derived from settings --
// is NVIDIA's code talking.
//
// Conventions (deduced from the call sites): float4x4 is stored as four float4 COLUMNS col0..col3 (InstanceImpl.cpp:352-390
// builds it from 4 x 4 consecutive floats of a column-major user matrix; m[3].xyz is the translation, :417), clip = M * v.
// sizeof(float3) == 16 (InstanceImpl.h:84 "do not use float3 constants because of sizeof( ml::float3 ) = 16").
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

struct v4raw { float f[4]; };

struct float2
{
    float x, y;
    float2() = default;
    float2(float a, float b) : x(a), y(b) {}
};
struct int2
{
    int32_t x, y;
    int2() = default;
    int2(int32_t a, int32_t b) : x(a), y(b) {}
};
struct uint2
{
    uint32_t x, y;
    uint2() = default;
    uint2(uint32_t a, uint32_t b) : x(a), y(b) {}
};

struct float4;
struct float3
{
    union
    {
        struct { float x, y, z; };
        v4raw xmm; // Reblur.cpp:345 "consts->gCameraDelta = m_CameraDelta.xmm" (float4 receives all 16 bytes)
    };
    float3() = default;
    float3(float a, float b, float c) { x = a; y = b; z = c; xmm.f[3] = 0.0f; }
    explicit float3(const float4& v);
    static float3 Zero() { return float3(0.0f, 0.0f, 0.0f); }
};
inline float3 operator-(const float3& a, const float3& b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(const float3& a) { return float3(-a.x, -a.y, -a.z); }
inline float3 operator*(const float3& a, float s) { return float3(a.x * s, a.y * s, a.z * s); }

struct float4
{
    union
    {
        struct { float x, y, z, w; };
        float a[4];   // InstanceImpl.cpp:394 "m_Frustum.a"
        float3 xyz;   // InstanceImpl.cpp:417 "m_ViewToWorld[3].xyz"
    };
    float4() = default;
    float4(float a0, float a1, float a2, float a3) { x = a0; y = a1; z = a2; w = a3; }
    float4(const float* p) { x = p[0]; y = p[1]; z = p[2]; w = p[3]; } // InstanceImpl.cpp:354
    float4(const v4raw& r) { memcpy(a, r.f, sizeof(a)); }
    static float4 Zero() { return float4(0.0f, 0.0f, 0.0f, 0.0f); } // InstanceImpl.h:326-330
};
inline float3::float3(const float4& v) { x = v.x; y = v.y; z = v.z; xmm.f[3] = 0.0f; }
inline float4 operator-(const float4& v) { return float4(-v.x, -v.y, -v.z, -v.w); }
inline float4 operator*(const float4& p, const float4& q) { return float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w); }
inline float4 operator+(const float4& p, const float4& q) { return float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w); }

struct float4x4
{
    // aRC = row R, column C (Relax.cpp:69-70 uses the diagonal a00 / a11 only)
    union { float4 col0; struct { float a00, a10, a20, a30; }; };
    union { float4 col1; struct { float a01, a11, a21, a31; }; };
    union { float4 col2; struct { float a02, a12, a22, a32; }; };
    union { float4 col3; struct { float a03, a13, a23, a33; }; };
    float4x4() {}
    float4x4(const float4& c0, const float4& c1, const float4& c2, const float4& c3) { col0 = c0; col1 = c1; col2 = c2; col3 = c3; }
    static float4x4 Identity() { return float4x4(float4(1, 0, 0, 0), float4(0, 1, 0, 0), float4(0, 0, 1, 0), float4(0, 0, 0, 1)); }
    float4& operator[](uint32_t i) { return (&col0)[i]; }
    const float4& operator[](uint32_t i) const { return (&col0)[i]; }
    float at(int r, int c) const { return (&col0)[c].a[r]; }
    float& at(int r, int c) { return (&col0)[c].a[r]; }
    bool operator!=(const float4x4& o) const { return memcmp(this, &o, sizeof(*this)) != 0; } // Denoisers/Reference.hpp:65
    float4 GetRow0() const { return float4(at(0, 0), at(0, 1), at(0, 2), at(0, 3)); }
    float4 GetRow1() const { return float4(at(1, 0), at(1, 1), at(1, 2), at(1, 3)); }
    void Transpose()
    {
        for (int r = 0; r < 4; r++)
            for (int c = r + 1; c < 4; c++) std::swap(at(r, c), at(c, r));
    }
    void SetTranslation(const float3& t) { at(0, 3) = t.x; at(1, 3) = t.y; at(2, 3) = t.z; }
    // inverse of a rigid transform [R | t] = [R^T | -R^T t]   (InstanceImpl.cpp:412-428)
    void InvertOrtho()
    {
        float4x4 r = Identity();
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r.at(i, j) = at(j, i);
        const float tx = at(0, 3), ty = at(1, 3), tz = at(2, 3);
        r.at(0, 3) = -(r.at(0, 0) * tx + r.at(0, 1) * ty + r.at(0, 2) * tz);
        r.at(1, 3) = -(r.at(1, 0) * tx + r.at(1, 1) * ty + r.at(1, 2) * tz);
        r.at(2, 3) = -(r.at(2, 0) * tx + r.at(2, 1) * ty + r.at(2, 2) * tz);
        *this = r;
    }
    // general inverse: Gauss-Jordan with partial pivoting, evaluated in double   (InstanceImpl.cpp:433-443)
    void Invert()
    {
        double a[4][8];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++)
            {
                a[r][c] = at(r, c);
                a[r][c + 4] = r == c ? 1.0 : 0.0;
            }
        for (int col = 0; col < 4; col++)
        {
            int piv = col;
            for (int r = col + 1; r < 4; r++)
                if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
            if (a[piv][col] == 0.0) { *this = Identity(); return; }
            if (piv != col)
                for (int c = 0; c < 8; c++) std::swap(a[col][c], a[piv][c]);
            const double inv = 1.0 / a[col][col];
            for (int c = 0; c < 8; c++) a[col][c] *= inv;
            for (int r = 0; r < 4; r++)
                if (r != col)
                {
                    const double f = a[r][col];
                    if (f != 0.0)
                        for (int c = 0; c < 8; c++) a[r][c] -= f * a[col][c];
                }
        }
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) at(r, c) = (float)a[r][c + 4];
    }
};
inline float4x4 operator*(const float4x4& p, const float4x4& q)
{
    float4x4 r;
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++)
        {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += p.at(row, k) * q.at(k, c);
            r.at(row, c) = s;
        }
    return r;
}
inline float4 operator*(const float4x4& p, const float4& v)
{
    return float4(p.at(0, 0) * v.x + p.at(0, 1) * v.y + p.at(0, 2) * v.z + p.at(0, 3) * v.w, p.at(1, 0) * v.x + p.at(1, 1) * v.y + p.at(1, 2) * v.z + p.at(1, 3) * v.w,
                  p.at(2, 0) * v.x + p.at(2, 1) * v.y + p.at(2, 2) * v.z + p.at(2, 3) * v.w, p.at(3, 0) * v.x + p.at(3, 1) * v.y + p.at(3, 2) * v.z + p.at(3, 3) * v.w);
}
// rotation by the upper 3x3 (Sigma.cpp:107)
inline float3 Rotate(const float4x4& p, const float3& v)
{
    return float3(p.at(0, 0) * v.x + p.at(0, 1) * v.y + p.at(0, 2) * v.z, p.at(1, 0) * v.x + p.at(1, 1) * v.y + p.at(1, 2) * v.z,
                  p.at(2, 0) * v.x + p.at(2, 1) * v.y + p.at(2, 2) * v.z);
}

// ---- scalar helpers the host code calls unqualified -------------------------------------------------------------------
template <class A, class B> inline auto min(A a, B b) -> typename std::common_type<A, B>::type { return a < b ? a : b; }
template <class A, class B> inline auto max(A a, B b) -> typename std::common_type<A, B>::type { return a > b ? a : b; }
template <class T> inline T clamp(T x, T a, T b) { return x < a ? a : (x > b ? b : x); }
template <class T> inline void Swap(T& a, T& b) { T t = a; a = b; b = t; }
inline float saturate(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }
inline float radians(float deg) { return deg * 3.14159265358979323846f / 180.0f; }
using std::abs;
using std::log;

// ---- DecomposeProjection (InstanceImpl.cpp:394,446,451: only flags, frustum and project are requested) -------------------
enum : uint32_t { PROJ_ORTHO = 1, PROJ_LEFT_HANDED = 2 };
enum : uint32_t { STYLE_D3D = 0, STYLE_OGL = 1 };
// flags: left-handed iff clip.w grows with +z (perspective) / depth grows with +z (ortho); ortho iff the last row is (0,0,0,1)
// frustum (x0, y0, dx, dy): viewPos.xy = (uv * frustum.zw + frustum.xy) * viewZ inverts uv = clip.xy / clip.w * (0.5, -0.5) + 0.5
// project[1] = y scale (cot(fovY / 2) for a perspective projection)
inline void DecomposeProjection(uint32_t, uint32_t, const float4x4& p, uint32_t* outFlags, float*, float*, float* frustum, float* project, float*)
{
    uint32_t flags = 0;
    const bool isOrtho = p.at(3, 0) == 0.0f && p.at(3, 1) == 0.0f && p.at(3, 2) == 0.0f && p.at(3, 3) == 1.0f;
    const float m00 = p.at(0, 0), m11 = p.at(1, 1);
    if (isOrtho)
    {
        flags |= PROJ_ORTHO;
        if (p.at(2, 2) > 0.0f) flags |= PROJ_LEFT_HANDED;
        const float m03 = p.at(0, 3), m13 = p.at(1, 3);
        if (frustum)
        {
            frustum[0] = -(-1.0f - m03) / m00;
            frustum[1] = -(1.0f - m13) / m11;
            frustum[2] = -2.0f / m00;
            frustum[3] = 2.0f / m11;
        }
    }
    else
    {
        const float wz = p.at(3, 2);
        if (wz > 0.0f) flags |= PROJ_LEFT_HANDED;
        const float s = wz > 0.0f ? 1.0f : -1.0f;
        const float m02 = p.at(0, 2) * s, m12 = p.at(1, 2) * s;
        if (frustum)
        {
            frustum[0] = (-1.0f - m02) / m00;
            frustum[1] = (1.0f - m12) / m11;
            frustum[2] = 2.0f / m00;
            frustum[3] = -2.0f / m11;
        }
    }
    if (project)
    {
        project[0] = m00;
        project[1] = m11;
        project[2] = p.at(2, 2);
    }
    if (outFlags) *outFlags = flags;
}

// ---- Sequence / Geometry (InstanceImpl.cpp:340-349) -------------------------------------------------------------------------
namespace Sequence
{
// frac(p + n * phi^-1), the golden-ratio step held as a 24-bit fixed-point integer
inline float Weyl1D(float p, uint32_t n)
{
    float v = p + float(n * 10368889u) / 16777216.0f;
    return v - std::floor(v);
}
// 4x4 ordered-dither matrix value in [0,1), advanced by the frame index
inline float Bayer4x4(uint2 pixel, uint32_t frameIndex)
{
    static const uint32_t k[4][4] = {{0, 8, 2, 10}, {12, 4, 14, 6}, {3, 11, 1, 9}, {15, 7, 13, 5}};
    return float((k[pixel.y & 3][pixel.x & 3] + frameIndex) & 0xF) / 16.0f;
}
} // namespace Sequence
namespace Geometry
{
// (cos, sin, -sin, cos)
inline float4 GetRotator(float angle)
{
    float ca = std::cos(angle), sa = std::sin(angle);
    return float4(ca, sa, -sa, ca);
}
// 2x2 product of two rotators
inline float4 CombineRotators(const float4& r1, const float4& r2)
{
    return float4(r1.x * r2.x + r1.z * r2.y, r1.y * r2.x + r1.w * r2.y, r1.x * r2.z + r1.z * r2.w, r1.y * r2.z + r1.w * r2.w);
}
} // namespace Geometry
