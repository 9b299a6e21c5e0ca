// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product library.
// PARITY PINNED against the reference's own shader sources compiled for the CPU (oracle/build_refshaders.py,
// tests/test_reference_shaders.py: bit-identical but for the rounding-sensitive passes listed there); NOT pinned: MathLib (absent from the
// reference tree, restated in mathlib.h), the texture unit (this file / hlsl.h) and GPU float behaviour.
// The reference ships no golden vectors and no GPU API exists here; this file emulates the HLSL / D3D semantics the reference's
// shaders rely on so that they can be restated line by line in C++ -- and so that the shader sources
// themselves can run on the CPU (oracle/refshader/hlsl_cpp.h builds its textures on the Tex below):
//   * tex[p] / Load out of bounds returns 0, UAV stores out of bounds are dropped
//   * SampleLevel(gNearestClamp): texel clamp(floor(uv * size), 0, size-1)
//   * SampleLevel(gLinearClamp): clamp-to-edge bilinear with exact float weights frac(uv * size - 0.5)
//   * Gather*: (x,y,z,w) = texels (0,1),(1,1),(1,0),(0,0) of the bilinear footprint
//   * every store is quantised to the texture's format (UNORM round-half-up of saturate(x) * max, FP16 RTNE, ...)
//   * step(a,x) = x >= a, saturate(NaN) = 0, lerp(a,b,t) = a + (b - a) * t
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <immintrin.h>

namespace hlsl
{
typedef unsigned int uint;

struct float2
{
    float x, y;
    float2() : x(0), y(0) {}
    float2(float v) : x(v), y(v) {}
    float2(float x_, float y_) : x(x_), y(y_) {}
};
struct float3
{
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    float3(float v) : x(v), y(v), z(v) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(float2 a, float z_) : x(a.x), y(a.y), z(z_) {}
    float2 xy() const { return float2(x, y); }
};
struct float4
{
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float v) : x(v), y(v), z(v), w(v) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(float3 a, float w_) : x(a.x), y(a.y), z(a.z), w(w_) {}
    float4(float2 a, float2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    float3 xyz() const { return float3(x, y, z); }
    float2 xy() const { return float2(x, y); }
    float2 zw() const { return float2(z, w); }
    float2 xz() const { return float2(x, z); }
    float2 yw() const { return float2(y, w); }
    void set_xyz(float3 a) { x = a.x; y = a.y; z = a.z; }
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct int2
{
    int x, y;
    int2() : x(0), y(0) {}
    int2(int v) : x(v), y(v) {}
    int2(int x_, int y_) : x(x_), y(y_) {}
};
struct uint4 { uint x, y, z, w; };

#define HLSL_OPS(T, ...)                                                                                             \
    inline T operator+(T a, T b) { return T(__VA_ARGS__(+)); }                                                         \
    inline T operator-(T a, T b) { return T(__VA_ARGS__(-)); }                                                         \
    inline T operator*(T a, T b) { return T(__VA_ARGS__(*)); }                                                         \
    inline T operator/(T a, T b) { return T(__VA_ARGS__(/)); }                                                         \
    inline T& operator+=(T& a, T b) { a = a + b; return a; }                                                           \
    inline T& operator-=(T& a, T b) { a = a - b; return a; }                                                           \
    inline T& operator*=(T& a, T b) { a = a * b; return a; }                                                           \
    inline T& operator/=(T& a, T b) { a = a / b; return a; }
#define HLSL_E2(op) a.x op b.x, a.y op b.y
#define HLSL_E3(op) a.x op b.x, a.y op b.y, a.z op b.z
#define HLSL_E4(op) a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w
HLSL_OPS(float2, HLSL_E2)
HLSL_OPS(float3, HLSL_E3)
HLSL_OPS(float4, HLSL_E4)
inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator-(int2 a, int2 b) { return int2(a.x - b.x, a.y - b.y); }
inline int2 operator*(int2 a, int b) { return int2(a.x * b, a.y * b); }
inline float2 tofloat(int2 a) { return float2(float(a.x), float(a.y)); }

// scalar intrinsics
inline float saturate(float x) { return x > 0.0f ? (x < 1.0f ? x : 1.0f) : 0.0f; } // NaN -> 0
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }
inline float step(float a, float x) { return x >= a ? 1.0f : 0.0f; }
inline float rcp(float x) { return 1.0f / x; }
inline float rsqrt(float x) { return 1.0f / std::sqrt(x); }
inline float frac(float x) { return x - std::floor(x); }
// D3D min/max: if one operand is NaN the other one is returned (IEEE minNum / maxNum), clamp = min(max(x, a), b)
inline float min(float a, float b) { return std::fmin(a, b); }
inline float max(float a, float b) { return std::fmax(a, b); }
inline float clamp(float x, float a, float b) { return std::fmin(std::fmax(x, a), b); }
inline float abs(float a) { return std::fabs(a); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float floor(float a) { return std::floor(a); }
inline float exp2(float a) { return std::exp2(a); }
inline float exp(float a) { return std::exp(a); }
inline float log(float a) { return std::log(a); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float atan(float a) { return std::atan(a); }
inline float sign(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
inline int clamp(int x, int a, int b) { return std::min(std::max(x, a), b); }
inline int min(int a, int b) { return std::min(a, b); }
inline int max(int a, int b) { return std::max(a, b); }

#define HLSL_MAP1(fn)                                                                                                  \
    inline float2 fn(float2 a) { return float2(fn(a.x), fn(a.y)); }                                                    \
    inline float3 fn(float3 a) { return float3(fn(a.x), fn(a.y), fn(a.z)); }                                           \
    inline float4 fn(float4 a) { return float4(fn(a.x), fn(a.y), fn(a.z), fn(a.w)); }
HLSL_MAP1(saturate) HLSL_MAP1(abs) HLSL_MAP1(floor) HLSL_MAP1(sqrt) HLSL_MAP1(frac)
#define HLSL_MAP2(fn)                                                                                                  \
    inline float2 fn(float2 a, float2 b) { return float2(fn(a.x, b.x), fn(a.y, b.y)); }                                \
    inline float3 fn(float3 a, float3 b) { return float3(fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z)); }                  \
    inline float4 fn(float4 a, float4 b) { return float4(fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z), fn(a.w, b.w)); }
HLSL_MAP2(min) HLSL_MAP2(max) HLSL_MAP2(step)
inline float2 lerp(float2 a, float2 b, float t) { return a + (b - a) * float2(t); }
inline float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * float3(t); }
inline float4 lerp(float4 a, float4 b, float t) { return a + (b - a) * float4(t); }
inline float2 lerp(float2 a, float2 b, float2 t) { return a + (b - a) * t; }
inline float4 lerp(float4 a, float4 b, float4 t) { return a + (b - a) * t; }
inline float2 clamp(float2 x, float2 a, float2 b) { return min(max(x, a), b); }
inline int2 clamp(int2 p, int2 a, int2 b) { return int2(clamp(p.x, a.x, b.x), clamp(p.y, a.y, b.y)); }
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(float2 a) { return std::sqrt(dot(a, a)); }
inline float length(float3 a) { return std::sqrt(dot(a, a)); }
inline float3 normalize(float3 a) { return a * float3(rsqrt(dot(a, a))); }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float3 reflect(float3 i, float3 n) { return i - n * float3(2.0f * dot(i, n)); }

// column-major 4x4, mul(M, v): c[k] is column k
struct float4x4
{
    float4 c[4];
};
inline float4 mul(const float4x4& m, float4 v) { return m.c[0] * float4(v.x) + m.c[1] * float4(v.y) + m.c[2] * float4(v.z) + m.c[3] * float4(v.w); }
struct float3x3
{
    float3 r[3]; // rows, as float3x3( T, B, N ) builds them
};
inline float3 mul(const float3x3& m, float3 v) { return float3(dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v)); }

// f32 <-> f16, round to nearest even (F16C)
inline uint16_t f32tof16(float f) { return (uint16_t)_cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT); }
inline float f16tof32(uint16_t h) { return _cvtsh_ss(h); }
inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }

// ---------------------------------------------------------------------------------------------
// Textures.  Format ids are nrd::Format enumerators (reference: Include/NRDDescs.h:264-332).
// ---------------------------------------------------------------------------------------------
enum Fmt : int
{
    R8_UNORM = 0, R8_UINT = 2, RG8_UNORM = 4, RGBA8_UNORM = 8, R16_UNORM = 13, R16_UINT = 15, R16_SFLOAT = 17, RGBA16_SFLOAT = 27,
    R32_UINT = 28, R32_SFLOAT = 30, RGBA32_SFLOAT = 39, R10_G10_B10_A2_UNORM = 40,
};

inline uint unormEncode(float v, float maxv)
{
    float s = saturate(v); // NaN -> 0
    return (uint)(s * maxv + 0.5f);
}

struct Tex
{
    uint8_t* data = nullptr;
    int w = 0, h = 0;
    int pitch = 0;
    int fmt = 0;
    int yoff = 0; // first row physically present (strips); always 0 in the oracle's own tests
    // WithRectOrigin / WithRectOffset (Common.hlsli:200-205): the guide inputs of an application are addressed at rectOrigin + pixelPos.
    // Every access below takes rect-relative coordinates and adds the origin; w / h stay the size of the whole resource (samplers
    // scale uv by it).
    int ox = 0, oy = 0;

    const uint8_t* at(int x, int y) const { return data + size_t(y - yoff) * pitch + size_t(x) * bpp(); }
    uint8_t* at(int x, int y) { return data + size_t(y - yoff) * pitch + size_t(x) * bpp(); }
    int bpp() const
    {
        switch (fmt)
        {
            case R8_UNORM: case R8_UINT: return 1;
            case RG8_UNORM: case R16_UNORM: case R16_UINT: case R16_SFLOAT: return 2;
            case RGBA8_UNORM: case R32_UINT: case R32_SFLOAT: case R10_G10_B10_A2_UNORM: return 4;
            case RGBA16_SFLOAT: return 8;
            case RGBA32_SFLOAT: return 16;
        }
        return 0;
    }
    bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < w && y < h; }

    float4 load(int x, int y) const
    {
        x += ox;
        y += oy;
        if (!inside(x, y)) return float4(0.0f);
        const uint8_t* p = at(x, y);
        switch (fmt)
        {
            case R8_UNORM: return float4(p[0] / 255.0f, 0, 0, 0);
            case RG8_UNORM: return float4(p[0] / 255.0f, p[1] / 255.0f, 0, 0);
            case RGBA8_UNORM: return float4(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f);
            case R16_UNORM: { uint16_t v; memcpy(&v, p, 2); return float4(v / 65535.0f, 0, 0, 0); }
            case R16_SFLOAT: { uint16_t v; memcpy(&v, p, 2); return float4(f16tof32(v), 0, 0, 0); }
            case RGBA16_SFLOAT: { uint16_t v[4]; memcpy(v, p, 8); return float4(f16tof32(v[0]), f16tof32(v[1]), f16tof32(v[2]), f16tof32(v[3])); }
            case R32_SFLOAT: { float v; memcpy(&v, p, 4); return float4(v, 0, 0, 0); }
            case RGBA32_SFLOAT: { float v[4]; memcpy(v, p, 16); return float4(v[0], v[1], v[2], v[3]); }
            case R10_G10_B10_A2_UNORM:
            {
                uint v; memcpy(&v, p, 4);
                return float4((v & 1023) / 1023.0f, ((v >> 10) & 1023) / 1023.0f, ((v >> 20) & 1023) / 1023.0f, (v >> 30) / 3.0f);
            }
        }
        return float4(0.0f);
    }
    float4 load(int2 p) const { return load(p.x, p.y); }
    uint loadu(int x, int y) const
    {
        x += ox;
        y += oy;
        if (!inside(x, y)) return 0;
        const uint8_t* p = at(x, y);
        switch (fmt)
        {
            case R8_UINT: return p[0];
            case R16_UINT: { uint16_t v; memcpy(&v, p, 2); return v; }
            case R32_UINT: { uint v; memcpy(&v, p, 4); return v; }
        }
        return 0;
    }
    uint loadu(int2 p) const { return loadu(p.x, p.y); }

    void store(int x, int y, float4 v)
    {
        x += ox;
        y += oy;
        if (!inside(x, y)) return;
        uint8_t* p = at(x, y);
        switch (fmt)
        {
            case R8_UNORM: p[0] = (uint8_t)unormEncode(v.x, 255.0f); break;
            case RG8_UNORM: p[0] = (uint8_t)unormEncode(v.x, 255.0f); p[1] = (uint8_t)unormEncode(v.y, 255.0f); break;
            case RGBA8_UNORM: for (int i = 0; i < 4; i++) p[i] = (uint8_t)unormEncode(v[i], 255.0f); break;
            case R16_UNORM: { uint16_t q = (uint16_t)unormEncode(v.x, 65535.0f); memcpy(p, &q, 2); break; }
            case R16_SFLOAT: { uint16_t q = f32tof16(v.x); memcpy(p, &q, 2); break; }
            case RGBA16_SFLOAT: { uint16_t q[4] = {f32tof16(v.x), f32tof16(v.y), f32tof16(v.z), f32tof16(v.w)}; memcpy(p, q, 8); break; }
            case R32_SFLOAT: memcpy(p, &v.x, 4); break;
            case RGBA32_SFLOAT: { float q[4] = {v.x, v.y, v.z, v.w}; memcpy(p, q, 16); break; }
            case R10_G10_B10_A2_UNORM:
            {
                uint q = unormEncode(v.x, 1023.0f) | (unormEncode(v.y, 1023.0f) << 10) | (unormEncode(v.z, 1023.0f) << 20) | (unormEncode(v.w, 3.0f) << 30);
                memcpy(p, &q, 4);
                break;
            }
        }
    }
    void store(int2 p, float4 v) { store(p.x, p.y, v); }
    void store(int2 p, float v) { store(p.x, p.y, float4(v, 0, 0, 0)); }
    void storeu(int x, int y, uint v)
    {
        x += ox;
        y += oy;
        if (!inside(x, y)) return;
        uint8_t* p = at(x, y);
        switch (fmt)
        {
            case R8_UINT: p[0] = (uint8_t)std::min(v, 255u); break;
            case R16_UINT: { uint16_t q = (uint16_t)std::min(v, 65535u); memcpy(p, &q, 2); break; }
            case R32_UINT: memcpy(p, &v, 4); break;
        }
    }
    void storeu(int2 p, uint v) { storeu(p.x, p.y, v); }

    // clamp-to-edge texel fetch used by the samplers
    float4 texel(int x, int y) const { return load(clamp(x, -ox, w - 1 - ox), clamp(y, -oy, h - 1 - oy)); }
    uint texelu(int x, int y) const { return loadu(clamp(x, -ox, w - 1 - ox), clamp(y, -oy, h - 1 - oy)); }

    float4 sampleNearest(float2 uv) const { return texel((int)std::floor(uv.x * w), (int)std::floor(uv.y * h)); }
    float4 sampleLinear(float2 uv) const
    {
#ifdef ORACLE_NUDGE_UV
        // noise-floor variant (liboracle_uv.so): the uv of every bilinear fetch moved by one float ulp -- the sub-texel position a
        // shader hands to the sampler is only known to ulp(uv) * size (2.4e-4 texel at 4K), tests/parity.py
        uv.x = std::nextafter(uv.x, 2.0f);
        uv.y = std::nextafter(uv.y, 2.0f);
#endif
        float px = uv.x * w - 0.5f, py = uv.y * h - 0.5f;
        float fx = std::floor(px), fy = std::floor(py);
        float wx = px - fx, wy = py - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float4 a = lerp(texel(x0, y0), texel(x0 + 1, y0), wx);
        float4 b = lerp(texel(x0, y0 + 1), texel(x0 + 1, y0 + 1), wx);
        return lerp(a, b, wy);
    }
    // Gather of channel `ch` (0=red..3=alpha); returns texels (0,1),(1,1),(1,0),(0,0) of the 2x2 footprint
    float4 gather(float2 uv, int ch, int2 offset = int2(0, 0)) const
    {
        int x0 = (int)std::floor(uv.x * w - 0.5f) + offset.x, y0 = (int)std::floor(uv.y * h - 0.5f) + offset.y;
        return float4(texel(x0, y0 + 1)[ch], texel(x0 + 1, y0 + 1)[ch], texel(x0 + 1, y0)[ch], texel(x0, y0)[ch]);
    }
    uint4 gatheru(float2 uv, int2 offset = int2(0, 0)) const
    {
        int x0 = (int)std::floor(uv.x * w - 0.5f) + offset.x, y0 = (int)std::floor(uv.y * h - 0.5f) + offset.y;
        return {texelu(x0, y0 + 1), texelu(x0 + 1, y0 + 1), texelu(x0 + 1, y0), texelu(x0, y0)};
    }
};
} // namespace hlsl
