// Front-end / back-end helpers of the denoisers' wire formats for CUDA and C++ callers (what a CUDA path tracer includes to feed
// libnrd_b200.so and to read its outputs back).  The reference ships them as HLSL only: Shaders/Include/NRD.hlsli:319-592
// (internals) and :594-1161 (public NRD_FrontEnd_* / REBLUR_* / RELAX_* / SIGMA_* / NRD_SG_* / NRD_SH_* functions); every function
// below cites the lines whose behaviour it restates, with the reference's default encodings (NRD_NORMAL_ENCODING = R10G10B10A2_UNORM,
// NRD_ROUGHNESS_ENCODING = LINEAR, CMakeLists.txt:28-29) -- the ones libnrd_b200.so reports in LibraryDesc.
//
// Plain inline functions on CUDA's float2 / float3 / float4 (vector_types.h), usable from device code (nvcc) and from host C++
// (any compiler, -I<cuda>/include).  Storage formats of the textures themselves (IN_NORMAL_ROUGHNESS R10G10B10A2_UNORM,
// radiance RGBA16F, IN_PENUMBRA R16F ...) are listed in include/nrd_b200.h; nrdPackR10G10B10A2 below does the UNORM quantisation.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector_types.h>

#if defined(__CUDACC__)
#define NRD_HD __host__ __device__ inline
#else
#define NRD_HD inline
#endif

namespace nrd_frontend
{
constexpr float NRD_FP16_MAX = 65504.0f;
constexpr float NRD_PI = 3.14159265358979323846f;
constexpr float NRD_EPS = 1e-6f;
constexpr float NRD_INF = 1e6f;
constexpr float NRD_REJITTER_VIEWZ_THRESHOLD = 0.01f;
constexpr float NRD_MATERIAL_FACTOR_MIN_SCALE = 0.02f;
constexpr float NRD_ROUGHNESS_FACTOR_MIN_SCALE = 0.1f;
constexpr float NRD_ROUGHNESS_EPS = 0.03162277660168379f; // sqrt(sqrt(NRD_EPS))

// ---- small vector vocabulary ---------------------------------------------------------------------------------------------
NRD_HD float2 f2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
NRD_HD float3 f3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
NRD_HD float4 f4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
NRD_HD float3 add(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
NRD_HD float3 sub(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
NRD_HD float3 mul(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
NRD_HD float3 scale(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
NRD_HD float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NRD_HD float length3(float3 a) { return sqrtf(dot3(a, a)); }
NRD_HD float3 normalize3(float3 a) { return scale(a, 1.0f / sqrtf(dot3(a, a))); }
NRD_HD float3 reflect3(float3 i, float3 n) { return sub(i, scale(n, 2.0f * dot3(i, n))); }
NRD_HD float saturatef(float x) { return x > 0.0f ? (x < 1.0f ? x : 1.0f) : 0.0f; } // NaN -> 0, like HLSL saturate
NRD_HD float clampf(float x, float a, float b) { return fminf(fmaxf(x, a), b); }
NRD_HD float lerpf(float a, float b, float t) { return a + (b - a) * t; }
NRD_HD float3 lerp3(float3 a, float3 b, float t) { return f3(lerpf(a.x, b.x, t), lerpf(a.y, b.y, t), lerpf(a.z, b.z, t)); }
NRD_HD float stepf(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }
NRD_HD bool IsInvalid(float x) { return isnan(x) || isinf(x); }                                      // NRD.hlsli:531-534
NRD_HD bool IsInvalid(float3 v) { return IsInvalid(v.x) || IsInvalid(v.y) || IsInvalid(v.z); }        // NRD.hlsli:526-529

// ---- internals (NRD.hlsli:319-523) ----------------------------------------------------------------------------------------
NRD_HD float3 SafeNormalize(float3 v) { return scale(v, 1.0f / sqrtf(dot3(v, v) + 1e-9f)); }          // :321-324
NRD_HD float2 EncodeUnitVector(float3 v, bool isSigned)                                              // :327-335 (octahedral)
{
    const float l1 = fabsf(v.x) + fabsf(v.y) + fabsf(v.z);
    v = scale(v, 1.0f / l1);
    const float wx = (1.0f - fabsf(v.y)) * (stepf(0.0f, v.x) * 2.0f - 1.0f), wy = (1.0f - fabsf(v.x)) * (stepf(0.0f, v.y) * 2.0f - 1.0f);
    const float x = v.z >= 0.0f ? v.x : wx, y = v.z >= 0.0f ? v.y : wy;
    return isSigned ? f2(x, y) : f2(x * 0.5f + 0.5f, y * 0.5f + 0.5f);
}
NRD_HD float3 DecodeUnitVector(float2 p, bool isSigned, bool normalize)                              // :337-347
{
    if (!isSigned) p = f2(p.x * 2.0f - 1.0f, p.y * 2.0f - 1.0f);
    float3 n = f3(p.x, p.y, 1.0f - fabsf(p.x) - fabsf(p.y));
    const float t = saturatef(-n.z);
    n.x -= t * (stepf(0.0f, n.x) * 2.0f - 1.0f);
    n.y -= t * (stepf(0.0f, n.y) * 2.0f - 1.0f);
    return normalize ? normalize3(n) : n;
}
NRD_HD float Luminance(float3 c) { return dot3(c, f3(0.2126f, 0.7152f, 0.0722f)); }                   // :350-354
NRD_HD float3 LinearToYCoCg(float3 c)                                                                // :356-363
{
    return f3(dot3(c, f3(0.25f, 0.5f, 0.25f)), dot3(c, f3(0.5f, 0.0f, -0.5f)), dot3(c, f3(-0.25f, 0.5f, -0.25f)));
}
NRD_HD float3 YCoCgToLinear(float3 c)                                                                // :365-375
{
    const float t = c.x - c.z;
    return f3(fmaxf(t + c.y, 0.0f), fmaxf(c.x + c.z, 0.0f), fmaxf(t - c.y, 0.0f));
}
NRD_HD float3 YCoCgToLinearCorrected(float Y, float Y0, float2 CoCg)                                 // :377-383
{
    Y = fmaxf(Y, 0.0f);
    const float s = (Y + NRD_EPS) / (Y0 + NRD_EPS);
    return YCoCgToLinear(f3(Y, CoCg.x * s, CoCg.y * s));
}
NRD_HD float GetSpecularDominantFactor(float NoV, float roughness)                                   // :386-392
{
    const float a = 0.298475f * logf(39.4115f - 39.0029f * roughness);
    return saturatef(powf(saturatef(1.0f - NoV), 10.8649f) * (1.0f - a) + a);
}
NRD_HD float3 GetSpecularDominantDirection(float3 N, float3 V, float dominantFactor)                 // :394-400
{
    return normalize3(lerp3(N, reflect3(scale(V, -1.0f), N), dominantFactor));
}
NRD_HD float GetSpecMagicCurve(float roughness) { return 1.0f - exp2f(-30.0f * roughness * roughness); } // :402-405
NRD_HD float Pow5(float x) { return powf(saturatef(1.0f - x), 5.0f); }                                // :408-411
NRD_HD float FresnelTerm(float Rf0, float VoNH) { return Rf0 + (1.0f - Rf0) * Pow5(VoNH); }           // :413-416
NRD_HD float DistributionTerm(float roughness, float NoH)                                            // :418-428 (GGX)
{
    const float m = roughness * roughness, m2 = m * m;
    const float t = (NoH * m2 - NoH) * NoH + 1.0f;
    const float a = m / t;
    return a * a / NRD_PI;
}
NRD_HD float GeometryTerm(float roughness, float NoL, float NoV)                                     // :430-439
{
    const float m = roughness * roughness, m2 = m * m;
    const float a = NoL + sqrtf(saturatef((NoL - m2 * NoL) * NoL + m2));
    const float b = NoV + sqrtf(saturatef((NoV - m2 * NoV) * NoV + m2));
    return 1.0f / fmaxf(a * b, NRD_EPS);
}
NRD_HD float DiffuseTerm(float roughness, float NoL, float NoV, float VoH)                           // :441-451 (Burley)
{
    const float m = roughness * roughness;
    const float f = 2.0f * VoH * VoH * m - 0.5f;
    return (f * Pow5(NoV) + 1.0f) * (f * Pow5(NoL) + 1.0f) / NRD_PI;
}
NRD_HD float2 ComputeBrdfs(float3 Ld, float3 Ls, float3 N, float3 V, float Rf0, float roughness)     // :453-488
{
    float2 result;
    const float NoV = fabsf(dot3(N, V));
    {
        const float3 H = normalize3(add(Ld, V));
        const float NoL = saturatef(dot3(N, Ld)), VoH = saturatef(dot3(V, H));
        result.x = (1.0f - FresnelTerm(Rf0, VoH)) * DiffuseTerm(roughness, NoL, NoV, VoH) * NoL;
    }
    {
        float3 H = normalize3(add(Ls, V));
        H = normalize3(lerp3(N, H, roughness));
        const float NoL = saturatef(dot3(N, Ls)), NoH = saturatef(dot3(N, H)), VoH = saturatef(dot3(V, H));
        result.y = FresnelTerm(Rf0, VoH) * DistributionTerm(roughness, NoH) * GeometryTerm(roughness, NoL, NoV) * NoL;
    }
    return result;
}
NRD_HD float3 EnvironmentTermRtg(float3 Rf0, float NoV, float roughness)                             // :490-517 (rational fit)
{
    const float m = saturatef(roughness * roughness);
    const float X[4] = {1.0f, NoV, NoV * NoV, NoV * NoV * NoV}, Y[4] = {1.0f, m, m * m, m * m * m};
    // mul(M, v) of row-major float2x2 / float3x3 literals, then dot with Y
    const float b0 = (0.99044f * X[0] - 1.28514f * X[1]) * Y[0] + (1.29678f * X[0] - 0.755907f * X[1]) * Y[1];
    const float b1 = (1.0f * X[0] + 2.92338f * X[1] + 59.4188f * X[3]) * Y[0] + (20.3225f * X[0] - 27.0302f * X[1] + 222.592f * X[3]) * Y[1] +
                     (121.563f * X[0] + 626.13f * X[1] + 316.627f * X[3]) * Y[3];
    const float s0 = (0.0365463f * X[0] + 3.32707f * X[1]) * Y[0] + (9.0632f * X[0] - 9.04756f * X[1]) * Y[1];
    const float s1 = (1.0f * X[0] + 3.59685f * X[2] - 1.36772f * X[3]) * Y[0] + (9.04401f * X[0] - 16.3174f * X[2] + 9.22949f * X[3]) * Y[1] +
                     (5.56589f * X[0] + 19.7886f * X[2] - 20.2123f * X[3]) * Y[3];
    const float bias = b0 / fmaxf(b1, NRD_EPS), sc = s0 / fmaxf(s1, NRD_EPS);
    return f3(saturatef(Rf0.x * sc + bias), saturatef(Rf0.y * sc + bias), saturatef(Rf0.z * sc + bias));
}
// hitDistParams = nrd::HitDistanceParameters {A, B, C, D}
NRD_HD float ReblurGetHitDistanceNormalization(float viewZ, float4 hitDistParams, float roughness)   // :520-523
{
    return (hitDistParams.x + fabsf(viewZ) * hitDistParams.y) * lerpf(1.0f, hitDistParams.z, saturatef(exp2f(hitDistParams.w * roughness * roughness)));
}

// ---- spherical gaussian / SH carrier (NRD.hlsli:541-590) ----------------------------------------------------------------------
struct NRD_SG
{
    float c0;
    float2 chroma;
    float normHitDist;
    float3 c1;
    float sharpness;
};
NRD_HD NRD_SG SG_Create(float3 radiance, float3 direction, float normHitDist)                        // :551-563
{
    const float3 y = LinearToYCoCg(radiance);
    NRD_SG sg;
    sg.c0 = y.x;
    sg.chroma = f2(y.y, y.z);
    sg.c1 = scale(direction, y.x);
    sg.normHitDist = normHitDist;
    sg.sharpness = 0.0f;
    return sg;
}
NRD_HD float3 SG_ExtractDirectionInternal(const NRD_SG& sg) { return scale(sg.c1, 1.0f / fmaxf(length3(sg.c1), NRD_EPS)); } // :565-568
NRD_HD float SG_IntegralApprox(const NRD_SG& sg) { return 2.0f * NRD_PI * (sg.c0 / sg.sharpness); }   // :570-573
NRD_HD float SG_Integral(const NRD_SG& sg) { return SG_IntegralApprox(sg) * (1.0f - expf(-2.0f * sg.sharpness)); } // :575-580
NRD_HD float SG_InnerProduct(const NRD_SG& a, const NRD_SG& b)                                       // :582-590
{
    const float d = length3(add(scale(SG_ExtractDirectionInternal(a), a.sharpness), scale(SG_ExtractDirectionInternal(b), b.sharpness)));
    float c = expf(d - a.sharpness - b.sharpness);
    c *= 1.0f - expf(-2.0f * d);
    c /= fmaxf(d, NRD_EPS);
    return NRD_PI * saturatef(2.0f * c * a.c0) * b.c0;
}

// ---- FRONT-END, general (NRD.hlsli:594-718) -------------------------------------------------------------------------------------
// p = the four UNORM channels of IN_NORMAL_ROUGHNESS as floats in [0,1];  returns {N.xyz, linear roughness}, materialID in 0..3
NRD_HD float4 NRD_FrontEnd_UnpackNormalAndRoughness(float4 p, float& materialID)                     // :600-628
{
    const float3 n = SafeNormalize(DecodeUnitVector(f2(p.x, p.y), false, false));
    materialID = p.w * 3.0f;
    return f4(n.x, n.y, n.z, p.z);
}
NRD_HD float4 NRD_FrontEnd_UnpackNormalAndRoughness(float4 p) { float unused; return NRD_FrontEnd_UnpackNormalAndRoughness(p, unused); } // :631-635
NRD_HD float4 NRD_FrontEnd_PackNormalAndRoughness(float3 N, float roughness, float materialID)       // :640-667
{
    const float2 e = EncodeUnitVector(N, false);
    return f4(e.x, e.y, roughness, saturatef(materialID / 3.0f));
}
// float4 in [0,1] -> R10G10B10A2_UNORM texel (D3D float -> UNORM: round half up of saturate(x) * max)
NRD_HD uint32_t nrdPackR10G10B10A2(float4 p)
{
    const uint32_t x = (uint32_t)(saturatef(p.x) * 1023.0f + 0.5f), y = (uint32_t)(saturatef(p.y) * 1023.0f + 0.5f), z = (uint32_t)(saturatef(p.z) * 1023.0f + 0.5f),
                   w = (uint32_t)(saturatef(p.w) * 3.0f + 0.5f);
    return x | (y << 10) | (z << 20) | (w << 30);
}
NRD_HD float4 nrdUnpackR10G10B10A2(uint32_t v)
{
    return f4((float)(v & 1023u) / 1023.0f, (float)((v >> 10) & 1023u) / 1023.0f, (float)((v >> 20) & 1023u) / 1023.0f, (float)(v >> 30) / 3.0f);
}
// material de-modulation: irradiance / factor before NRD, radiance * factor after                   :676-688
NRD_HD void NRD_MaterialFactors(float3 N, float3 V, float3 albedo, float3 Rf0, float roughness, float3& diffFactor, float3& specFactor)
{
    const float NoV = fabsf(dot3(N, V));
    const float3 Fenv = EnvironmentTermRtg(Rf0, NoV, roughness);
    diffFactor = mul(f3(1.0f - Fenv.x, 1.0f - Fenv.y, 1.0f - Fenv.z), albedo);
    diffFactor = f3(lerpf(NRD_MATERIAL_FACTOR_MIN_SCALE, 1.0f, diffFactor.x), lerpf(NRD_MATERIAL_FACTOR_MIN_SCALE, 1.0f, diffFactor.y), lerpf(NRD_MATERIAL_FACTOR_MIN_SCALE, 1.0f, diffFactor.z));
    specFactor = scale(Fenv, lerpf(NRD_ROUGHNESS_FACTOR_MIN_SCALE, 1.0f, roughness));
    specFactor = f3(lerpf(NRD_MATERIAL_FACTOR_MIN_SCALE, 1.0f, specFactor.x), lerpf(NRD_MATERIAL_FACTOR_MIN_SCALE, 1.0f, specFactor.y), lerpf(NRD_MATERIAL_FACTOR_MIN_SCALE, 1.0f, specFactor.z));
}
// specular hit distance averaging for rpp > 1                                                       :693-716
NRD_HD float NRD_FrontEnd_SpecHitDistAveraging_Begin() { return NRD_INF; }
NRD_HD float NRD_FrontEnd_TrimHitDistance(float hitDist, float threshold) { return hitDist < threshold ? 0.0f : hitDist; }
NRD_HD void NRD_FrontEnd_SpecHitDistAveraging_Add(float& accumulated, float hitDist) { accumulated = fminf(accumulated, hitDist == 0.0f ? NRD_INF : hitDist); }
NRD_HD void NRD_FrontEnd_SpecHitDistAveraging_End(float& accumulated) { accumulated = accumulated == NRD_INF ? 0.0f : accumulated; }

// ---- FRONT-END, REBLUR (NRD.hlsli:722-783) ----------------------------------------------------------------------------------------
NRD_HD float REBLUR_FrontEnd_GetNormHitDist(float hitDist, float viewZ, float4 hitDistParams, float roughness) // :722-727
{
    return saturatef(hitDist / ReblurGetHitDistanceNormalization(viewZ, hitDistParams, roughness));
}
NRD_HD float3 SanitizeRadiance(float3 r) { return IsInvalid(r) ? f3(0, 0, 0) : f3(clampf(r.x, 0.0f, NRD_FP16_MAX), clampf(r.y, 0.0f, NRD_FP16_MAX), clampf(r.z, 0.0f, NRD_FP16_MAX)); }
NRD_HD float3 SanitizeDirection(float3 d) { return IsInvalid(d) ? f3(0, 0, 0) : f3(clampf(d.x, -1.0f, 1.0f), clampf(d.y, -1.0f, 1.0f), clampf(d.z, -1.0f, 1.0f)); }
// -> IN_DIFF_RADIANCE_HITDIST / IN_SPEC_RADIANCE_HITDIST (YCoCg radiance + normalised hit distance)   :732-743
NRD_HD float4 REBLUR_FrontEnd_PackRadianceAndNormHitDist(float3 radiance, float normHitDist, bool sanitize = true)
{
    if (sanitize)
    {
        radiance = SanitizeRadiance(radiance);
        normHitDist = IsInvalid(normHitDist) ? 0.0f : saturatef(normHitDist);
    }
    const float3 y = LinearToYCoCg(radiance);
    return f4(y.x, y.y, y.z, normHitDist);
}
// -> IN_*_SH0 (returned) and IN_*_SH1 (out1)                                                         :748-765
NRD_HD float4 REBLUR_FrontEnd_PackSh(float3 radiance, float normHitDist, float3 direction, float4& out1, bool sanitize = true)
{
    if (sanitize)
    {
        radiance = SanitizeRadiance(radiance);
        normHitDist = IsInvalid(normHitDist) ? 0.0f : saturatef(normHitDist);
        direction = SanitizeDirection(direction);
    }
    const NRD_SG sg = SG_Create(radiance, direction, normHitDist);
    out1 = f4(sg.c1.x, sg.c1.y, sg.c1.z, sg.sharpness);
    return f4(sg.c0, sg.chroma.x, sg.chroma.y, sg.normHitDist);
}
// -> IN_DIFF_DIRECTION_HITDIST                                                                       :770-783
NRD_HD float4 REBLUR_FrontEnd_PackDirectionalOcclusion(float3 direction, float normHitDist, bool sanitize = true)
{
    if (sanitize)
    {
        direction = SanitizeDirection(direction);
        normHitDist = IsInvalid(normHitDist) ? 0.0f : saturatef(normHitDist);
    }
    const NRD_SG sg = SG_Create(f3(normHitDist, normHitDist, normHitDist), direction, normHitDist);
    return f4(sg.c1.x, sg.c1.y, sg.c1.z, sg.c0);
}

// ---- FRONT-END, RELAX (NRD.hlsli:789-822) -----------------------------------------------------------------------------------------
NRD_HD float4 RELAX_FrontEnd_PackRadianceAndHitDist(float3 radiance, float hitDist, bool sanitize = true) // :789-798
{
    if (sanitize)
    {
        radiance = SanitizeRadiance(radiance);
        hitDist = IsInvalid(hitDist) ? 0.0f : clampf(hitDist, 0.0f, NRD_FP16_MAX);
    }
    return f4(radiance.x, radiance.y, radiance.z, hitDist);
}
NRD_HD float4 RELAX_FrontEnd_PackSh(float3 radiance, float hitDist, float3 direction, float4& out1, bool sanitize = true) // :802-822
{
    if (sanitize)
    {
        radiance = SanitizeRadiance(radiance);
        hitDist = IsInvalid(hitDist) ? 0.0f : clampf(hitDist, 0.0f, NRD_FP16_MAX);
        direction = SanitizeDirection(direction);
    }
    const float3 d = scale(direction, Luminance(radiance));
    out1 = f4(d.x, d.y, d.z, 0.0f);
    return f4(radiance.x, radiance.y, radiance.z, hitDist);
}

// ---- FRONT-END, SIGMA (NRD.hlsli:828-857) -----------------------------------------------------------------------------------------
// infinite (directional) light: -> IN_PENUMBRA.  distanceToOccluder: 0 where NoL <= 0, hit distance, >= NRD_FP16_MAX on a miss
NRD_HD float SIGMA_FrontEnd_PackPenumbra(float distanceToOccluder, float tanOfLightAngularRadius)    // :828-834
{
    const float penumbraRadius = distanceToOccluder * tanOfLightAngularRadius * 0.5f;
    return distanceToOccluder >= NRD_FP16_MAX ? NRD_FP16_MAX : fminf(penumbraRadius, 32768.0f);
}
// local light                                                                                       :839-845
NRD_HD float SIGMA_FrontEnd_PackPenumbra(float distanceToOccluder, float distanceToLight, float lightSize)
{
    const float penumbraSize = lightSize * distanceToOccluder / fmaxf(distanceToLight - distanceToOccluder, NRD_EPS);
    return distanceToOccluder >= NRD_FP16_MAX ? NRD_FP16_MAX : fminf(penumbraSize * 0.5f, 32768.0f);
}
// -> IN_TRANSLUCENCY                                                                                 :848-857
NRD_HD float4 SIGMA_FrontEnd_PackTranslucency(float distanceToOccluder, float3 translucency)
{
    return f4(distanceToOccluder >= NRD_FP16_MAX ? 1.0f : 0.0f, saturatef(translucency.x), saturatef(translucency.y), saturatef(translucency.z));
}

// ---- BACK-END (NRD.hlsli:863-935) -----------------------------------------------------------------------------------------------------
NRD_HD float4 REBLUR_BackEnd_UnpackRadianceAndNormHitDist(float4 data)                               // :863-868
{
    const float3 c = YCoCgToLinear(f3(data.x, data.y, data.z));
    return f4(c.x, c.y, c.z, data.w);
}
NRD_HD NRD_SG REBLUR_BackEnd_UnpackSh(float4 sh0, float4 sh1)                                        // :872-882
{
    NRD_SG sg;
    sg.c0 = sh0.x;
    sg.chroma = f2(sh0.y, sh0.z);
    sg.normHitDist = sh0.w;
    sg.c1 = f3(sh1.x, sh1.y, sh1.z);
    sg.sharpness = sh1.w;
    return sg;
}
NRD_HD NRD_SG REBLUR_BackEnd_UnpackDirectionalOcclusion(float4 data)                                 // :885-896
{
    NRD_SG sg;
    sg.c0 = data.w;
    sg.chroma = f2(0.0f, 0.0f);
    sg.normHitDist = data.w;
    sg.c1 = f3(data.x, data.y, data.z);
    sg.sharpness = 0.0f;
    return sg;
}
NRD_HD float4 RELAX_BackEnd_UnpackRadiance(float4 color) { return color; }                           // :903-906
NRD_HD NRD_SG RELAX_BackEnd_UnpackSh(float4 sh0, float4 sh1) { return REBLUR_BackEnd_UnpackSh(sh0, sh1); } // :910-920
// OUT_SHADOW_TRANSLUCENCY -> shadow (.x) [and translucent shadow .yzw]                               :931
NRD_HD float SIGMA_BackEnd_UnpackShadow(float shadow) { return shadow * shadow; }
NRD_HD float4 SIGMA_BackEnd_UnpackShadow(float4 s) { return f4(s.x * s.x, s.y * s.y, s.z * s.z, s.w * s.w); }

// ---- BACK-END, high-quality resolve of SG / SH outputs (NRD.hlsli:937-1137) -----------------------------------------------------------
NRD_HD float3 NRD_SG_ExtractColor(const NRD_SG& sg) { return YCoCgToLinear(f3(sg.c0, sg.chroma.x, sg.chroma.y)); } // :937-940
NRD_HD float3 NRD_SG_ExtractDirection(const NRD_SG& sg) { return SG_ExtractDirectionInternal(sg); }   // :942-945
NRD_HD float NRD_SG_ExtractRoughnessAA(const NRD_SG& sg) { return sg.sharpness; }                     // :947-950
// rotation: three ROWS of a 3x3 matrix                                                              :952-955
NRD_HD void NRD_SG_Rotate(NRD_SG& sg, const float3 rotation[3]) { sg.c1 = f3(dot3(rotation[0], sg.c1), dot3(rotation[1], sg.c1), dot3(rotation[2], sg.c1)); }
NRD_HD float3 NRD_SG_ResolveDiffuse(NRD_SG sg, float3 N)                                             // :957-1007 (numerically integrated irradiance, sharpness 4)
{
    sg.sharpness = 4.0f;
    const float c0 = 0.36f, c1 = 1.0f / (4.0f * c0);
    const float e = expf(-sg.sharpness), e2 = e * e, r = 1.0f / sg.sharpness;
    const float sc = 1.0f + 2.0f * e2 - r, bias = (e - e2) * r - e2;
    const float NoL = dot3(N, SG_ExtractDirectionInternal(sg));
    const float x = sqrtf(saturatef(1.0f - sc)), x0 = c0 * NoL, x1 = c1 * x, n = x0 + x1;
    float y = saturatef(NoL);
    if (fabsf(x0) <= x1) y = n * n / x;
    const float Y = (sc * y + bias) * SG_IntegralApprox(sg);
    return YCoCgToLinearCorrected(Y, sg.c0, sg.chroma);
}
NRD_HD float3 NRD_SG_ResolveSpecular(NRD_SG sg, float3 N, float3 V, float roughness)                 // :1009-1054
{
    roughness = fmaxf(roughness, NRD_ROUGHNESS_EPS);
    sg.sharpness = 2.0f;
    float3 H = normalize3(add(SG_ExtractDirectionInternal(sg), V));
    H = normalize3(lerp3(N, H, roughness));
    const float m = roughness * roughness, m2 = m * m;
    NRD_SG ndf;
    ndf.c0 = 1.0f / (NRD_PI * m2) * lerpf(1.0f, 0.75f * 2.0f * NRD_PI, m2);
    ndf.c1 = H;
    ndf.sharpness = 2.0f / fmaxf(m2, NRD_EPS);
    NRD_SG warped;
    warped.c0 = ndf.c0;
    warped.c1 = reflect3(scale(V, -1.0f), ndf.c1);
    warped.sharpness = ndf.sharpness / fmaxf(4.0f * fabsf(dot3(ndf.c1, V)), NRD_EPS);
    warped.chroma = f2(0.0f, 0.0f);
    warped.normHitDist = 0.0f;
    const float NoV = fabsf(dot3(N, V)), NoL = saturatef(dot3(N, warped.c1));
    warped.c0 *= NoL * GeometryTerm(roughness, NoL, NoV);
    return YCoCgToLinearCorrected(SG_InnerProduct(warped, sg), sg.c0, sg.chroma);
}
// neighbours: e = (+1, 0), w = (-1, 0), n = (0, +1), s = (0, -1); out-of-screen fetches must return 0  :1064-1112
NRD_HD float2 NRD_SG_ReJitter(const NRD_SG& diffSg, const NRD_SG& specSg, float3 Rf0, float3 V, float roughness, float Z, float Ze, float Zw, float Zn, float Zs, float3 N,
                              float3 Ne, float3 Nw, float3 Nn, float3 Ns)
{
    roughness = fmaxf(roughness, NRD_ROUGHNESS_EPS);
    const float rf0 = Luminance(Rf0);
    const float3 Ld = SG_ExtractDirectionInternal(diffSg);
    float3 Ls = SG_ExtractDirectionInternal(specSg);
    Ls = normalize3(lerp3(V, Ls, GetSpecMagicCurve(roughness)));
    const float2 centre = ComputeBrdfs(Ld, Ls, N, V, rf0, roughness);
    float2 avg = ComputeBrdfs(Ld, Ls, Ne, V, rf0, roughness);
    const float2 bn = ComputeBrdfs(Ld, Ls, Nn, V, rf0, roughness), bw = ComputeBrdfs(Ld, Ls, Nw, V, rf0, roughness), bs = ComputeBrdfs(Ld, Ls, Ns, V, rf0, roughness);
    avg = f2(avg.x + bn.x + bw.x + bs.x, avg.y + bn.y + bw.y + bs.y);
    const float NoV = fabsf(dot3(N, V));
    const float zThreshold = NRD_REJITTER_VIEWZ_THRESHOLD * fabsf(Z) / (NoV * 0.95f + 0.05f);
    int sum = (fabsf(Ze - Z) < zThreshold && dot3(Ne, N) > 0.0f) ? 1 : 0;
    sum += (fabsf(Zn - Z) < zThreshold && dot3(Nn, N) > 0.0f) ? 1 : 0;
    sum += (fabsf(Zw - Z) < zThreshold && dot3(Nw, N) > 0.0f) ? 1 : 0;
    sum += (fabsf(Zs - Z) < zThreshold && dot3(Ns, N) > 0.0f) ? 1 : 0;
    const float2 f = f2((centre.x * 4.0f + NRD_EPS) / (avg.x + NRD_EPS), (centre.y * 4.0f + NRD_EPS) / (avg.y + NRD_EPS));
    return sum != 4 ? f2(1.0f, 1.0f) : f2(clampf(f.x, 1.0f / NRD_PI, NRD_PI), clampf(f.y, 1.0f / NRD_PI, NRD_PI));
}
NRD_HD float3 NRD_SH_ResolveDiffuse(const NRD_SG& sh, float3 N)                                       // :1117-1122
{
    return YCoCgToLinearCorrected(dot3(N, sh.c1) + 0.5f * sh.c0, sh.c0, sh.chroma);
}
NRD_HD float3 NRD_SH_ResolveSpecular(const NRD_SG& sh, float3 N, float3 V, float roughness)           // :1124-1135
{
    const float f = GetSpecularDominantFactor(fabsf(dot3(N, V)), roughness);
    const float3 D = GetSpecularDominantDirection(N, V, f);
    return YCoCgToLinearCorrected(dot3(D, sh.c1) + 0.5f * sh.c0, sh.c0, sh.chroma);
}

// ---- misc (NRD.hlsli:1140-1161) ------------------------------------------------------------------------------------------------------
NRD_HD bool NRD_IsValidRadiance(float3 radiance) { return !IsInvalid(radiance); }
NRD_HD float REBLUR_GetHitDist(float normHitDist, float viewZ, float4 hitDistParams, float roughness) { return normHitDist * ReblurGetHitDistanceNormalization(viewZ, hitDistParams, roughness); }
// pixelSize = gUnproject * (isOrtho ? 1 : |viewZ|)
NRD_HD float NRD_GetNormalizedStrandThickness(float strandThickness, float pixelSize) { return pixelSize / (pixelSize + strandThickness); }
} // namespace nrd_frontend
