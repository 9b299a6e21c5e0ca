// ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED at this boundary (the one part of the oracle that the reference-shader pin of
// tests/test_reference_shaders.py cannot reach: both sides of that comparison call these functions).
// Restatement of the shader-side MathLib functions the reference's passes call (ml.hlsli from NVIDIA-RTX/MathLib,
// fetched by the reference's CMakeLists.txt:120-129 with GIT_TAG main, i.e. unpinned, and absent from /root/reference).
// Each function lists its first call site in the reference; in-tree twins (Shaders/Include/NRD.hlsli) win where they exist.
// SURVEY.md Appendix B carries the confidence of every definition.
#pragma once
#include "hlsl.h"

namespace hlsl
{
namespace Math
{
inline float Pi(float x) { return x * 3.14159265358979323846f; }
inline float DegToRad(float x) { return x * 3.14159265358979323846f / 180.0f; }
// REBLUR_Common.hlsli:109 (also used with a > b)
inline float LinearStep(float a, float b, float x) { return saturate((x - a) / (b - a)); }
inline float SmoothStep01(float x) { float t = saturate(x); return t * t * (3.0f - 2.0f * t); }
inline float4 SmoothStep01(float4 x) { return float4(SmoothStep01(x.x), SmoothStep01(x.y), SmoothStep01(x.z), SmoothStep01(x.w)); }
// Common.hlsli:560
inline float SmoothStep(float a, float b, float x) { return SmoothStep01(LinearStep(a, b, x)); }
inline float Sqrt01(float x) { return sqrt(saturate(x)); }
inline float Pow01(float x, float y) { return pow(saturate(x), y); }
inline float2 Pow01(float2 x, float y) { return float2(Pow01(x.x, y), Pow01(x.y, y)); }
inline float Rsqrt(float x) { return rsqrt(x); }
inline float LengthSquared(float3 v) { return dot(v, v); }
inline float LengthSquared(float2 v) { return dot(v, v); }
// REBLUR_Common_DiffuseSpatialFilter.hlsli:168
inline float PositiveRcp(float x) { return 1.0f / max(x, 1e-15f); }
// REBLUR_Common_DiffuseSpatialFilter.hlsli:143 -- small-angle acos
inline float AcosApprox(float x) { return sqrt(2.0f) * sqrt(saturate(1.0f - x)); }
} // namespace Math

namespace Geometry
{
// Common.hlsli:258, InstanceImpl.cpp:341
inline float4 GetRotator(float a) { return float4(std::cos(a), std::sin(a), -std::sin(a), std::cos(a)); }
// Common.hlsli:472
inline float2 RotateVector(float4 r, float2 v) { return float2(v.x) * r.xz() + float2(v.y) * r.yw(); }
// Common.hlsli:264
inline float4 CombineRotators(float4 r1, float4 r2)
{
    return float4(r1.x, r1.y, r1.x, r1.y) * float4(r2.x, r2.x, r2.z, r2.z) + float4(r1.z, r1.w, r1.z, r1.w) * float4(r2.y, r2.y, r2.w, r2.w);
}
// REBLUR_Common_DiffuseSpatialFilter.hlsli:90
inline float4 ScaleRotator(float4 r, float2 s) { return r * float4(s.x, s.x, s.y, s.y); }
// rotate by the upper 3x3 of a column-major matrix / by its transpose
inline float3 RotateVector(const float4x4& m, float3 v) { return m.c[0].xyz() * float3(v.x) + m.c[1].xyz() * float3(v.y) + m.c[2].xyz() * float3(v.z); }
inline float3 RotateVectorInverse(const float4x4& m, float3 v) { return float3(dot(m.c[0].xyz(), v), dot(m.c[1].xyz(), v), dot(m.c[2].xyz(), v)); }
inline float3 RotateVector(const float3x3& m, float3 v) { return mul(m, v); }
// REBLUR_TemporalAccumulation.hlsli:139, Common.hlsli:475
inline float3 AffineTransform(const float4x4& m, float3 p) { return mul(m, float4(p, 1.0f)).xyz(); }
inline float4 ProjectiveTransform(const float4x4& m, float3 p) { return mul(m, float4(p, 1.0f)); }
// Common.hlsli:327
inline float2 GetScreenUv(const float4x4& worldToClip, float3 X, bool killBackprojection = true)
{
    float4 clip = ProjectiveTransform(worldToClip, X);
    float2 uv = (clip.xy() / float2(clip.w)) * float2(0.5f, -0.5f) + float2(0.5f);
    if (killBackprojection && clip.w < 0.0f) uv = float2(99999.0f);
    return uv;
}
// REBLUR_Blur.hlsli:40
inline float3 ReconstructViewPosition(float2 uv, float4 frustum, float viewZ = 1.0f, float orthoMode = 0.0f)
{
    float3 p;
    float2 xy = uv * frustum.zw() + frustum.xy();
    xy = xy * float2(viewZ * (1.0f - abs(orthoMode)) + orthoMode);
    p.x = xy.x; p.y = xy.y; p.z = viewZ;
    return p;
}
// REBLUR_Common.hlsli:280 -- branchless orthonormal basis (Duff et al. 2017); rows T, B, N
inline float3x3 GetBasis(float3 N)
{
    float sz = N.z < 0.0f ? -1.0f : 1.0f;
    float a = 1.0f / (sz + N.z);
    float ya = N.y * a;
    float b = N.x * ya;
    float c = N.x * sz;
    float3x3 m;
    m.r[0] = float3(c * N.x * a - 1.0f, sz * b, c);
    m.r[1] = float3(b, N.y * ya - sz, N.y);
    m.r[2] = N;
    return m;
}
} // namespace Geometry

namespace Filtering
{
struct Bilinear { float2 origin; float2 weights; };
// REBLUR_TemporalAccumulation.hlsli:180
inline Bilinear GetBilinearFilter(float2 uv, float2 texSize)
{
    float2 t = uv * texSize - float2(0.5f);
    Bilinear r;
    r.origin = floor(t);
    r.weights = t - r.origin;
    return r;
}
inline float ApplyBilinearFilter(float s00, float s10, float s01, float s11, Bilinear f)
{
    return lerp(lerp(s00, s10, f.weights.x), lerp(s01, s11, f.weights.x), f.weights.y);
}
inline float4 GetBilinearCustomWeights(Bilinear f, float4 customWeights)
{
    float2 oneMinus = float2(1.0f) - f.weights;
    float4 w = customWeights;
    w.x *= oneMinus.x * oneMinus.y;
    w.y *= f.weights.x * oneMinus.y;
    w.z *= oneMinus.x * f.weights.y;
    w.w *= f.weights.x * f.weights.y;
    return w;
}
// mirrored in-tree at Common.hlsli:645-656
inline float ApplyBilinearCustomWeights(float s00, float s10, float s01, float s11, float4 w)
{
    float sum = dot(w, float4(1.0f));
    float r = s00 * w.x + s10 * w.y + s01 * w.z + s11 * w.w;
    return sum < 0.0001f ? 0.0f : r / sum;
}
struct CatmullRom { float2 origin; };
// REBLUR_TemporalAccumulation.hlsli:166 -- top-left texel centre of the 4x4 footprint
inline CatmullRom GetCatmullRomFilter(float2 uv, float2 texSize)
{
    float2 tc = floor(uv * texSize - float2(0.5f)) + float2(0.5f);
    CatmullRom r;
    r.origin = tc - float2(1.5f);
    return r;
}
// REBLUR_TemporalAccumulation.hlsli:106 -- Toksvig-style roughness widening
inline float GetModifiedRoughnessFromNormalVariance(float linearRoughness, float3 nonNormalizedAverageNormal)
{
    float l = length(nonNormalizedAverageNormal);
    float kappa = saturate(1.0f - l * l) / max(l * (3.0f - l * l), 1e-15f);
    return Math::Sqrt01(linearRoughness * linearRoughness + kappa);
}
} // namespace Filtering

namespace Packing
{
// REBLUR_Common.hlsli:19,26
inline uint RgbaToUint(float4 c, uint rb, uint gb, uint bb, uint ab)
{
    const uint bits[4] = {rb, gb, bb, ab};
    uint p = 0, shift = 0;
    for (int i = 0; i < 4; i++)
    {
        if (bits[i])
        {
            float maxv = float((1u << bits[i]) - 1u);
            p |= uint(saturate(c[i]) * maxv + 0.5f) << shift;
        }
        shift += bits[i];
    }
    return p;
}
inline float4 UintToRgba(uint p, uint rb, uint gb, uint bb, uint ab)
{
    const uint bits[4] = {rb, gb, bb, ab};
    float4 c(0.0f);
    uint shift = 0;
    for (int i = 0; i < 4; i++)
    {
        if (bits[i])
        {
            uint maxv = (1u << bits[i]) - 1u;
            c[i] = float((p >> shift) & maxv) / float(maxv);
        }
        shift += bits[i];
    }
    return c;
}
} // namespace Packing

namespace ImportanceSampling
{
// in-tree twin NRD.hlsli:386-392 (G2 fit)
inline float GetSpecularDominantFactor(float NoV, float roughness)
{
    float a = 0.298475f * log(39.4115f - 39.0029f * roughness);
    float f = pow(saturate(1.0f - NoV), 10.8649f) * (1.0f - a) + a;
    return saturate(f);
}
// Common.hlsli:414 ; twin NRD.hlsli:394-400
inline float4 GetSpecularDominantDirection(float3 N, float3 V, float roughness)
{
    float NoV = abs(dot(N, V));
    float f = GetSpecularDominantFactor(NoV, roughness);
    float3 R = reflect(-V, N);
    float3 D = normalize(lerp(N, R, f));
    return float4(D, f);
}
// Common.hlsli:489 -- tan of the half angle of the cone holding `percentOfVolume` of the GGX lobe
inline float GetSpecularLobeTanHalfAngle(float roughness, float percentOfVolume = 0.75f)
{
    float m = saturate(roughness);
    m = m * m;
    return m * sqrt(percentOfVolume / (1.0f - percentOfVolume + 1e-6f));
}
} // namespace ImportanceSampling

namespace Color
{
inline float Luminance(float3 c) { return dot(c, float3(0.2126f, 0.7152f, 0.0722f)); }
// REBLUR_TemporalStabilization.hlsli:160
inline float Clamp(float m1, float sigma, float x) { return clamp(x, m1 - sigma, m1 + sigma); }
} // namespace Color

namespace BRDF
{
inline float Pow5(float x) { return pow(saturate(1.0f - x), 5.0f); }
// MathLib ml.hlsli (absent from the reference tree; restated from the published MathLib / STL sources, FROZEN CHOICE like the rest of
// this file): BRDF::ConvertBaseColorMetalnessToAlbedoRf0 and BRDF::EnvironmentTerm_Rtg ("Ray Tracing Gems", chapter 32: the
// polynomial fit of the split-sum environment BRDF)
inline void ConvertBaseColorMetalnessToAlbedoRf0(float3 baseColor, float metalness, float3& albedo, float3& Rf0)
{
    albedo = baseColor * float3(saturate(1.0f - metalness));
    Rf0 = lerp(float3(0.04f), baseColor, metalness);
}
inline float3 EnvironmentTerm_Rtg(float3 Rf0, float NoV, float linearRoughness)
{
    float m = linearRoughness * linearRoughness;
    float4 X(1.0f, NoV, NoV * NoV, NoV * NoV * NoV);
    float4 Y(1.0f, m, m * m, m * m * m);
    // mul(M, v) with row-major literals: M1 = {{0.99044, -1.28514}, {1.29678, -0.755907}} etc.
    float2 m1 = float2(0.99044f * X.x + -1.28514f * X.y, 1.29678f * X.x + -0.755907f * X.y);
    float3 m2 = float3(1.0f * X.x + 2.92338f * X.y + 59.4188f * X.w, 20.3225f * X.x + -27.0302f * X.y + 222.592f * X.w, 121.563f * X.x + 626.13f * X.y + 316.627f * X.w);
    float2 m3 = float2(0.0365463f * X.x + 3.32707f * X.y, 9.0632f * X.x + -9.04756f * X.y);
    float3 m4 = float3(1.0f * X.x + 3.59685f * X.z + -1.36772f * X.w, 9.04401f * X.x + -16.3174f * X.z + 9.22949f * X.w, 5.56589f * X.x + 19.7886f * X.z + -20.2123f * X.w);
    float bias = dot(m1, float2(Y.x, Y.y)) * rcp(dot(m2, float3(Y.x, Y.y, Y.w)));
    float scale = dot(m3, float2(Y.x, Y.y)) * rcp(dot(m4, float3(Y.x, Y.y, Y.w)));
    return saturate(Rf0 * float3(scale) + float3(bias));
}
} // namespace BRDF

namespace Sequence
{
// REBLUR_PrePass.hlsli:44
inline uint CheckerBoard(int2 p, uint frameIndex) { return ((uint(p.x) ^ uint(p.y)) ^ frameIndex) & 1u; }
// Common.hlsli:261, REBLUR_TemporalAccumulation.hlsli:408 -- frozen: standard 4x4 ordered-dither matrix
inline float Bayer4x4(int2 p, uint frameIndex)
{
    static const uint k[4][4] = {{0, 8, 2, 10}, {12, 4, 14, 6}, {3, 11, 1, 9}, {15, 7, 13, 5}};
    return float((k[p.y & 3][p.x & 3] + frameIndex) & 0xF) / 16.0f;
}
} // namespace Sequence

// Rng::Hash -- "random by design", unverifiable: frozen here as a PCG-style integer hash stream.
struct RngHash
{
    uint state = 0;
    static uint pcg(uint v)
    {
        uint s = v * 747796405u + 2891336453u;
        uint w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
        return (w >> 22u) ^ w;
    }
    void Initialize(int2 pixelPos, uint frameIndex) { state = pcg(uint(pixelPos.x) + pcg(uint(pixelPos.y) + pcg(frameIndex))); }
    float GetFloat()
    {
        state = pcg(state);
        return float(state >> 8) * (1.0f / 16777216.0f);
    }
    float2 GetFloat2()
    {
        float a = GetFloat();
        float b = GetFloat();
        return float2(a, b);
    }
};
} // namespace hlsl
