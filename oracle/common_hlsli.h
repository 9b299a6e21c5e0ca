// ORACLE -- TEST INFRASTRUCTURE ONLY (see hlsl.h for what the parity of the oracle is pinned against).
// Restatement of the denoiser-independent helpers of Shaders/Include/Common.hlsli and NRD.hlsli used by the SIGMA and
// RELAX restatements (the REBLUR file carries its own copies inside its Pass struct).
#pragma once
#include "mathlib.h"

namespace hlsl
{
namespace common
{
const float NRD_EPS = 1e-6f;
const float NRD_INF = 1e6f;
const float NRD_FP16_MAX = 65504.0f;
const float NRD_NORMAL_ENCODING_ERROR = 0.75f / 255.0f; // Common.hlsli:79-81
const float NRD_DISOCCLUSION_THRESHOLD = 0.02f;         // Common.hlsli:67
const float NRD_CATROM_SHARPNESS = 0.5f;                // Common.hlsli:68

// Common.hlsli:181-192
static const float3 g_Special8[8] = {
    float3(-1.0f, 0.0f, 1.0f), float3(0.0f, 1.0f, 1.0f), float3(1.0f, 0.0f, 1.0f), float3(0.0f, -1.0f, 1.0f),
    float3(-0.25f * 1.41421354f, 0.25f * 1.41421354f, 0.5f), float3(0.25f * 1.41421354f, 0.25f * 1.41421354f, 0.5f),
    float3(0.25f * 1.41421354f, -0.25f * 1.41421354f, 0.5f), float3(-0.25f * 1.41421354f, -0.25f * 1.41421354f, 0.5f)};

inline float3 _NRD_SafeNormalize(float3 v) { return v * float3(rsqrt(dot(v, v) + 1e-9f)); }   // NRD.hlsli:321-324
inline float3 _NRD_DecodeUnitVector(float2 p)                                                 // NRD.hlsli:337-347
{
    p = p * float2(2.0f) - float2(1.0f);
    float3 n = float3(p.x, p.y, 1.0f - abs(p.x) - abs(p.y));
    float t = saturate(-n.z);
    n.x -= t * (step(0.0f, n.x) * 2.0f - 1.0f);
    n.y -= t * (step(0.0f, n.y) * 2.0f - 1.0f);
    return n;
}
inline float4 NRD_FrontEnd_UnpackNormalAndRoughness(float4 p, float& materialID)              // NRD.hlsli:600-628
{
    float4 r;
    r.set_xyz(_NRD_DecodeUnitVector(p.xy()));
    r.w = p.z;
    materialID = p.w * 3.0f;
    r.set_xyz(_NRD_SafeNormalize(r.xyz()));
    return r;
}
inline float4 NRD_FrontEnd_UnpackNormalAndRoughness(float4 p) { float m; return NRD_FrontEnd_UnpackNormalAndRoughness(p, m); }
inline float _NRD_Luminance(float3 c) { return dot(c, float3(0.2126f, 0.7152f, 0.0722f)); }   // NRD.hlsli:350-354

inline float PixelRadiusToWorld(float unproject, float orthoMode, float pixelRadius, float viewZ) { return pixelRadius * unproject * lerp(viewZ, 1.0f, abs(orthoMode)); } // Common.hlsli:237
inline float GetFrustumSize(float minRectDimMulUnproject, float orthoMode, float viewZ) { return minRectDimMulUnproject * lerp(viewZ, 1.0f, abs(orthoMode)); }           // :242
inline float IsInScreenNearest(float2 uv) { return float(uv.x > 0.0f && uv.y > 0.0f && uv.x < 1.0f && uv.y < 1.0f); }                                                 // :280
inline float4 IsInScreenBilinear(float2 footprintOrigin, float2 rectSize)                                                                                                // :287
{
    float4 p = float4(footprintOrigin, footprintOrigin) + float4(0, 0, 1, 1);
    float4 r = float4(float(p.x >= 0.0f), float(p.y >= 0.0f), float(p.z >= 0.0f), float(p.w >= 0.0f));
    r *= float4(float(p.x < rectSize.x), float(p.y < rectSize.y), float(p.z < rectSize.x), float(p.w < rectSize.y));
    return float4(r.x, r.z, r.x, r.z) * float4(r.y, r.y, r.w, r.w);
}
inline float2 GetGeometryWeightParams(float planeDistSensitivity, float frustumSize, float3 Xv, float3 Nv)                                                              // :502
{
    float a = 1.0f / (planeDistSensitivity * frustumSize);
    float b = dot(Nv, Xv) * a;
    return float2(a, -b);
}
inline float ComputeNonExponentialWeight(float x, float px, float py) { return Math::SmoothStep(1.0f, 0.0f, abs(x * px + py)); }                                       // :559
inline float ComputeWeight(float x, float px, float py) { return ComputeNonExponentialWeight(x, px, py); }
inline float ExpApprox(float x) { return rcp(x * x - x + 1.0f); }                                                                                                        // :547
inline float ComputeExponentialWeight(float x, float px, float py) { return ExpApprox(-3.0f * abs(x * px + py)); }
inline float GetGaussianWeight(float r) { return exp(-0.66f * r * r); }                                                                                                  // :571
inline float GetDisocclusionThreshold(float t, float frustumSize, float NoV) { return frustumSize * saturate(t / max(0.01f, NoV)); }                                    // :594
inline float GetStdDev(float m1, float m2) { return sqrt(abs(m2 - m1 * m1)); }                                                                                           // :226
inline float GetSpecMagicCurve(float roughness, float power = 0.25f)                                                                                                     // :311
{
    float f = 1.0f - exp2(-200.0f * roughness * roughness);
    return f * Math::Pow01(roughness, power);
}

// Common.hlsli:602-646: CatRom-12 via 5 bilinear taps with fallback to custom bilinear weights
inline float4 BicubicCustom(float2 samplePos, float2 invResourceSize, float4 bilinearCustomWeights, bool useBicubic, const Tex& tex0)
{
    float2 centerPos = floor(samplePos - float2(0.5f)) + float2(0.5f);
    float2 f = saturate(samplePos - centerPos);
    const float S = NRD_CATROM_SHARPNESS;
    float2 w0 = f * (f * (float2(-S) * f + float2(2.0f * S)) - float2(S));
    float2 w1 = f * (f * (float2(2.0f - S) * f - float2(3.0f - S))) + float2(1.0f);
    float2 w2 = f * (f * (float2(-(2.0f - S)) * f + float2(3.0f - 2.0f * S)) + float2(S));
    float2 w3 = f * (f * (float2(S) * f - float2(S)));
    float2 w12 = w1 + w2;
    float2 tc = w2 / w12;
    float4 w;
    w.x = w12.x * w0.y;
    w.y = w0.x * w12.y;
    w.z = w12.x * w12.y;
    w.w = w3.x * w12.y;
    float w4 = w12.x * w3.y;
    w = useBicubic ? w : bilinearCustomWeights;
    w4 = useBicubic ? w4 : 0.0f;
    float sum = dot(w, float4(1.0f)) + w4;
    float4 cp = float4(centerPos, centerPos);
    float4 uv01 = cp + (useBicubic ? float4(tc.x, -1.0f, -1.0f, tc.y) : float4(0, 0, 1, 0));
    float4 uv23 = cp + (useBicubic ? float4(tc.x, tc.y, 2.0f, tc.y) : float4(0, 1, 1, 1));
    float2 uv4 = centerPos + (useBicubic ? float2(tc.x, 2.0f) : f);
    uv01 *= float4(invResourceSize, invResourceSize);
    uv23 *= float4(invResourceSize, invResourceSize);
    uv4 *= invResourceSize;
    float4 color = tex0.sampleLinear(uv01.xy()) * float4(w.x);
    color += tex0.sampleLinear(uv01.zw()) * float4(w.y);
    color += tex0.sampleLinear(uv23.xy()) * float4(w.z);
    color += tex0.sampleLinear(uv23.zw()) * float4(w.w);
    color += tex0.sampleLinear(uv4) * float4(w4);
    return sum < 0.0001f ? float4(0.0f) : color / float4(sum);
}
// custom-weight bilinear of a second texture at the same footprint (Common.hlsli:648-656)
inline float4 BilinearCustom(float2 samplePos, float4 w, const Tex& tex)
{
    float2 centerPos = floor(samplePos - float2(0.5f)) + float2(0.5f);
    int bx = (int)centerPos.x, by = (int)centerPos.y;
    float4 c = tex.load(bx, by) * float4(w.x);
    c += tex.load(bx + 1, by) * float4(w.y);
    c += tex.load(bx, by + 1) * float4(w.z);
    c += tex.load(bx + 1, by + 1) * float4(w.w);
    float sum = dot(w, float4(1.0f));
    return sum < 0.0001f ? float4(0.0f) : c / float4(sum);
}
} // namespace common
} // namespace hlsl
