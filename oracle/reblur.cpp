// ORACLE -- TEST INFRASTRUCTURE ONLY (see hlsl.h).  Every pass below is checked against the reference's own shader source of that pass,
// compiled for the CPU (oracle/build_refshaders.py), by tests/test_reference_shaders.py; the reference ships no golden vectors.
// CPU restatement, statement by statement, of the reference's REBLUR passes for the DIFFUSE / SPECULAR /
// DIFFUSE_SPECULAR denoisers at the default compile-time switches (Common.hlsli:51-85, REBLUR_Config.hlsli:13-98,
// NRD_NORMAL_ENCODING = 2, NRD_ROUGHNESS_ENCODING = 1, no REBLUR_PERFORMANCE_MODE / OCCLUSION / SH):
//   ClassifyTiles            Shaders/Source/REBLUR_ClassifyTiles.cs.hlsl:19-55
//   PrePass                  Shaders/Include/REBLUR_PrePass.hlsli:11-108
//   TemporalAccumulation     Shaders/Include/REBLUR_TemporalAccumulation.hlsli:11-931
//   HistoryFix               Shaders/Include/REBLUR_HistoryFix.hlsli:11-463
//   Blur / PostBlur          Shaders/Include/REBLUR_Blur.hlsli:11-74, REBLUR_PostBlur.hlsli:11-78
//   spatial filters          REBLUR_Common_DiffuseSpatialFilter.hlsli:23-213, REBLUR_Common_SpecularSpatialFilter.hlsli:23-260
//   TemporalStabilization    Shaders/Include/REBLUR_TemporalStabilization.hlsli:11-367
//   helpers                  Shaders/Include/Common.hlsli, REBLUR_Common.hlsli, NRD.hlsli (cited at each function)
// One "thread" per pixel, group-shared preloads are replaced by the clamped loads they perform.
#include "mathlib.h"
#include "oracle.h"

#include <cstdio>
#include <cstring>

// everything lives inside namespace hlsl so that unqualified abs/floor/sqrt/... resolve to the HLSL intrinsics
namespace hlsl
{
namespace
{
// ---- constant block: REBLUR_Config.hlsli:113-186 -----------------------------------------------
struct CB
{
    float4x4 gWorldToClip, gViewToClip, gViewToWorld, gWorldToViewPrev, gWorldToClipPrev, gWorldPrevToWorld;
    float4 gRotatorPre, gRotator, gRotatorPost, gFrustum, gFrustumPrev, gCameraDelta, gHitDistParams, gViewVectorWorld, gViewVectorWorldPrev, gMvScale;
    float2 gAntilagParams, gResourceSize, gResourceSizeInv, gResourceSizeInvPrev, gRectSize, gRectSizeInv, gRectSizePrev, gResolutionScale,
        gResolutionScalePrev, gRectOffset, gSpecProbabilityThresholdsForMvModification, gJitter;
    uint gPrintfAt[2], gRectOrigin[2];
    int gRectSizeMinusOne[2];
    float gDisocclusionThreshold, gDisocclusionThresholdAlternate, gCameraAttachedReflectionMaterialID, gStrandMaterialID, gStrandThickness,
        gStabilizationStrength, gHitDistStabilizationStrength, gDebug, gOrthoMode, gUnproject, gDenoisingRange, gPlaneDistSensitivity, gFramerateScale,
        gMinBlurRadius, gMaxBlurRadius, gDiffPrepassBlurRadius, gSpecPrepassBlurRadius, gMaxAccumulatedFrameNum, gMaxFastAccumulatedFrameNum,
        gAntiFirefly, gLobeAngleFraction, gRoughnessFraction, gResponsiveAccumulationRoughnessThreshold, gHistoryFixFrameNum, gHistoryFixBasePixelStride,
        gMinRectDimMulUnproject, gUsePrepassNotOnlyForSpecularMotionEstimation, gSplitScreen, gSplitScreenPrev, gCheckerboardResolveAccumSpeed, gViewZScale,
        gFireflySuppressorMinRelativeScale, gMinHitDistanceWeight, gDiffMinMaterial, gSpecMinMaterial;
    uint gHasHistoryConfidence, gHasDisocclusionThresholdMix, gDiffCheckerboard, gSpecCheckerboard, gFrameIndex, gIsRectChanged, gResetHistory;
};
static_assert(sizeof(CB) == 832, "REBLUR_SHARED_CONSTANTS is 832 bytes under HLSL packing");

// ---- compile-time settings ------------------------------------------------------------------------
const float NRD_EPS = 1e-6f;
const float NRD_INF = 1e6f;
const float NRD_NORMAL_ENCODING_ERROR = 0.75f / 255.0f;                 // Common.hlsli:79-81 (R10G10B10A2)
const float NRD_MAX_PERCENT_OF_LOBE_VOLUME = 0.75f;                      // Common.hlsli:75
const float NRD_ROUGHNESS_SENSITIVITY = 0.01f;                           // Common.hlsli:71
const float NRD_EXP_WEIGHT_DEFAULT_SCALE = 3.0f;                         // Common.hlsli:70
const float NRD_CATROM_SHARPNESS = 0.5f;                                 // Common.hlsli:68
const float NRD_CURVATURE_Z_THRESHOLD = 0.1f;                            // Common.hlsli:72
const float NRD_DISOCCLUSION_THRESHOLD = 0.02f;                          // Common.hlsli:67
const float REBLUR_MAX_ACCUM_FRAME_NUM = 63.0f;                          // REBLUR_Config.hlsli:60
const float REBLUR_MAX_MATERIALID_NUM = 15.0f;                           // REBLUR_Config.hlsli:61
const float REBLUR_PRE_BLUR_FRACTION_SCALE = 2.0f;                       // REBLUR_Config.hlsli:71
const float REBLUR_PRE_BLUR_NON_LINEAR_ACCUM_SPEED = 1.0f / (1.0f + 10.0f);
const float REBLUR_BLUR_FRACTION_SCALE = 1.0f;
const float REBLUR_POST_BLUR_FRACTION_SCALE = 0.5f;
const float REBLUR_POST_BLUR_RADIUS_SCALE = 2.0f;
const float REBLUR_NORMAL_ULP = NRD_NORMAL_ENCODING_ERROR;
const float REBLUR_FIREFLY_SUPPRESSOR_MAX_RELATIVE_INTENSITY = 38.0f;
const float REBLUR_FIREFLY_SUPPRESSOR_RADIUS_SCALE = 0.1f;
const float REBLUR_FIREFLY_SUPPRESSOR_FAST_RELATIVE_INTENSITY = 4.0f;
const int REBLUR_ANTI_FIREFLY_FILTER_RADIUS = 4;
const float REBLUR_ANTI_FIREFLY_SIGMA_SCALE = 2.0f;
const float REBLUR_ROUGHNESS_SENSITIVITY_IN_TA = NRD_ROUGHNESS_SENSITIVITY * 0.3f;
const float REBLUR_SAMPLES_PER_FRAME = 1.0f;
const float REBLUR_MAX_PERCENT_OF_LOBE_VOLUME_FOR_PRE_PASS = 0.3f;
const float REBLUR_COLOR_CLAMPING_SIGMA_SCALE = 2.0f;
enum { REBLUR_PRE_BLUR = 0, REBLUR_BLUR = 1, REBLUR_POST_BLUR = 2 };

float REBLUR_ALMOST_ZERO_ANGLE() { return std::cos(Math::DegToRad(89.0f)); }

// Common.hlsli:181-192
const float3 g_Special8[8] = {
    float3(-1.0f, 0.0f, 1.0f), float3(0.0f, 1.0f, 1.0f), float3(1.0f, 0.0f, 1.0f), float3(0.0f, -1.0f, 1.0f),
    float3(-0.25f * std::sqrt(2.0f), 0.25f * std::sqrt(2.0f), 0.5f), float3(0.25f * std::sqrt(2.0f), 0.25f * std::sqrt(2.0f), 0.5f),
    float3(0.25f * std::sqrt(2.0f), -0.25f * std::sqrt(2.0f), 0.5f), float3(-0.25f * std::sqrt(2.0f), -0.25f * std::sqrt(2.0f), 0.5f)};

// Common.hlsli:170-179
const float3 g_Special6[6] = {float3(-0.5f * std::sqrt(3.0f), -0.5f, 1.0f), float3(0.0f, 1.0f, 1.0f), float3(0.5f * std::sqrt(3.0f), -0.5f, 1.0f),
                              float3(0.0f, -0.3f, 0.3f), float3(0.15f * std::sqrt(3.0f), 0.15f, 0.3f), float3(-0.15f * std::sqrt(3.0f), 0.15f, 0.3f)};

// ---- NRD.hlsli helpers ----------------------------------------------------------------------------
float3 _NRD_SafeNormalize(float3 v) { return v * float3(rsqrt(dot(v, v) + 1e-9f)); }              // NRD.hlsli:321-324
float3 _NRD_DecodeUnitVector(float2 p)                                                            // NRD.hlsli:337-347 (unsigned, no normalize)
{
    p = p * float2(2.0f) - float2(1.0f);
    float3 n = float3(p.x, p.y, 1.0f - abs(p.x) - abs(p.y));
    float t = saturate(-n.z);
    n.x -= t * (step(0.0f, n.x) * 2.0f - 1.0f);
    n.y -= t * (step(0.0f, n.y) * 2.0f - 1.0f);
    return n;
}
float3 _NRD_LinearToYCoCg(float3 c)                                                               // NRD.hlsli:356-363
{
    return float3(dot(c, float3(0.25f, 0.5f, 0.25f)), dot(c, float3(0.5f, 0.0f, -0.5f)), dot(c, float3(-0.25f, 0.5f, -0.25f)));
}
float3 _NRD_YCoCgToLinear(float3 c)                                                               // NRD.hlsli:365-375
{
    float t = c.x - c.z;
    float3 r;
    r.y = c.x + c.z;
    r.x = t + c.y;
    r.z = t - c.y;
    return max(r, float3(0.0f));
}
float _REBLUR_GetHitDistanceNormalization(float viewZ, float4 p, float roughness)                  // NRD.hlsli:520-523
{
    return (p.x + abs(viewZ) * p.y) * lerp(1.0f, p.z, saturate(exp2(p.w * roughness * roughness)));
}
float4 NRD_FrontEnd_UnpackNormalAndRoughness(float4 p, float& materialID)                          // NRD.hlsli:600-628
{
    float4 r;
    r.set_xyz(_NRD_DecodeUnitVector(p.xy()));
    r.w = p.z;
    materialID = p.w * 3.0f;
    r.set_xyz(_NRD_SafeNormalize(r.xyz()));
    return r;
}
float4 NRD_FrontEnd_UnpackNormalAndRoughness(float4 p) { float unused; return NRD_FrontEnd_UnpackNormalAndRoughness(p, unused); }
float NRD_GetNormalizedStrandThickness(float strandThickness, float pixelSize) { return pixelSize / (pixelSize + strandThickness); } // NRD.hlsli:1158

// ---- Common.hlsli ------------------------------------------------------------------------------------
struct Pass
{
    const CB& c;
    // REBLUR_PERFORMANCE_MODE (REBLUR_Config.hlsli:196-238, the "REBLUR_Perf_*" shader permutations): no CatRom history filters,
    // screen-space sampling for both signals, 6 taps of g_Special6, anti-firefly radius 3, cheaper history fix / reconstruction
    bool perf = false;
    explicit Pass(const CB& cb, bool performanceMode = false) : c(cb), perf(performanceMode) {}

    float UnpackViewZ(float z) const { return abs(z * c.gViewZScale); }                                                   // :235
    float PixelRadiusToWorld(float unproject, float orthoMode, float pixelRadius, float viewZ) const                     // :237-240
    { return pixelRadius * unproject * lerp(viewZ, 1.0f, abs(orthoMode)); }
    float GetFrustumSize(float minRectDimMulUnproject, float orthoMode, float viewZ) const                               // :242-248
    { return minRectDimMulUnproject * lerp(viewZ, 1.0f, abs(orthoMode)); }
    static float GetHitDistFactor(float hitDist, float frustumSize) { return saturate(hitDist / frustumSize); }          // :250-253
    static float IsInScreenNearest(float2 uv) { return float(uv.x > 0.0f && uv.y > 0.0f && uv.x < 1.0f && uv.y < 1.0f); } // :280-283
    static float4 IsInScreenBilinear(float2 footprintOrigin, float2 rectSize)                                            // :287-295
    {
        float4 p = float4(footprintOrigin, footprintOrigin) + float4(0, 0, 1, 1);
        float4 r = float4(float(p.x >= 0.0f), float(p.y >= 0.0f), float(p.z >= 0.0f), float(p.w >= 0.0f));
        r *= float4(float(p.x < rectSize.x), float(p.y < rectSize.y), float(p.z < rectSize.x), float(p.w < rectSize.y));
        return float4(r.x, r.z, r.x, r.z) * float4(r.y, r.y, r.w, r.w);
    }
    static float2 ApplyCheckerboardShift(float2 pos, uint mode, uint counter, uint frameIndex)                          // :297-307
    {
        float2 posPositive = pos + float2(16384.0f);
        uint checkerboard = Sequence::CheckerBoard(int2((int)posPositive.x, (int)posPositive.y), frameIndex);
        float shift = ((counter & 0x1) == 0) ? -1.0f : 1.0f;
        pos.x += shift * float(checkerboard != mode && mode != 2);
        return pos;
    }
    static float GetSpecMagicCurve(float roughness, float power = 0.25f)                                                 // :311-317
    {
        float f = 1.0f - exp2(-200.0f * roughness * roughness);
        f *= Math::Pow01(roughness, power);
        return f;
    }
    static float ComputeParallaxInPixels(float3 X, float2 uvForZeroParallax, const float4x4& mWorldToClip, float2 rectSize) // :319-332
    {
        float2 uv = Geometry::GetScreenUv(mWorldToClip, X);
        float2 parallaxInUv = uv - uvForZeroParallax;
        return length(parallaxInUv * rectSize);
    }
    float2 ClampUvToViewport(float2 uv) const { return min(uv * c.gResolutionScale, c.gResolutionScale - float2(0.5f) * c.gResourceSizeInv); } // :215
    float2 ClampUvToViewportPrev(float2 uv) const { return uv; }

    // :404-461 thin-lens virtual position (NRD_USE_SPECULAR_MOTION_V2 = 1)
    static float3 GetXvirtual(float hitDist, float curvature, float3 X, float3 Xprev, float3 N, float3 V, float roughness)
    {
        float4 D = ImportanceSampling::GetSpecularDominantDirection(N, V, roughness);
        float3 Iw = V;
        float3 reflectionRay = D.xyz() * float3(hitDist);
        float3x3 reflectorBasis = Geometry::GetBasis(N);
        float3 O = Geometry::RotateVector(reflectorBasis, reflectionRay);
        O.z = -O.z;
        float mag = 1.0f / (2.0f * curvature * O.z - 1.0f);
        float f = length(X);
        f *= 1.0f - abs(dot(N, V));
        f *= max(curvature, 0.0f);
        mag *= 1.0f / (1.0f + f);
        float3 I = O * float3(mag);
        Iw = Iw * float3(length(I));
        float closenessToSurface = saturate(length(Iw) / (hitDist + NRD_EPS));
        float3 origin = lerp(Xprev, X, closenessToSurface * D.w);
        return origin - Iw * float3(D.w);
    }
    // :465-482
    // Texel selection of the Poisson taps.  The reference hands a uv to a nearest-neighbour sampler (fixed-point texel selection in
    // hardware, Common.hlsli:465-482 / REBLUR_Common_*SpatialFilter.hlsli); which texel a tap within rounding distance of a texel
    // border lands on is not defined by the HLSL source.  The oracle fixes ONE evaluation: the tap position is computed directly in
    // texel units, as fused multiply-add chains that start from the exact pixel centre.  The CUDA kernels use the same operations in
    // the same order (device/reblur_spatial.cu), so both sides land on the same texel.
    //   screen-space taps (Common.hlsli:472 RotateVector(rotator, v) = v.x * r.xz + v.y * r.yw, rotator scaled to pixels):
    //     t = pixelPos + 0.5 + RotateVector(rotator * skewInPixels.xxyy, offset)
    static float2 TapTexelScreen(int2 pixelPos, float4 R, float2 o)
    {
        float px = float(pixelPos.x) + 0.5f, py = float(pixelPos.y) + 0.5f;
        return float2(std::fma(o.x, R.x, std::fma(o.y, R.y, px)), std::fma(o.x, R.z, std::fma(o.y, R.w, py)));
    }
    //   world-space taps (Common.hlsli:465-482 GetKernelSampleCoordinates): p = X + T * o.x + B * o.y is affine in the (per-frame
    //   uniform) rotated offset o, so is clip = M * (p, 1): the clip-space images of X, T and B are evaluated once per pixel, with
    //   the x / y rows pre-scaled to texels (uv * rectSize = clip.xy / clip.w * (0.5, -0.5) * rectSize + 0.5 * rectSize)
    struct KernelProjection
    {
        float X0, XT, XB, Y0, YT, YB, W0, WT, WB, hW, hH;
    };
    static KernelProjection ProjectKernel(const float4x4& m, float2 rectSize, float3 X, float3 T, float3 B)
    {
        KernelProjection k;
        k.hW = 0.5f * rectSize.x;
        k.hH = 0.5f * rectSize.y;
        k.X0 = std::fma(m.c[2].x, X.z, std::fma(m.c[1].x, X.y, std::fma(m.c[0].x, X.x, m.c[3].x))) * k.hW;
        k.Y0 = std::fma(m.c[2].y, X.z, std::fma(m.c[1].y, X.y, std::fma(m.c[0].y, X.x, m.c[3].y))) * -k.hH;
        k.W0 = std::fma(m.c[2].w, X.z, std::fma(m.c[1].w, X.y, std::fma(m.c[0].w, X.x, m.c[3].w)));
        k.XT = std::fma(m.c[2].x, T.z, std::fma(m.c[1].x, T.y, m.c[0].x * T.x)) * k.hW;
        k.YT = std::fma(m.c[2].y, T.z, std::fma(m.c[1].y, T.y, m.c[0].y * T.x)) * -k.hH;
        k.WT = std::fma(m.c[2].w, T.z, std::fma(m.c[1].w, T.y, m.c[0].w * T.x));
        k.XB = std::fma(m.c[2].x, B.z, std::fma(m.c[1].x, B.y, m.c[0].x * B.x)) * k.hW;
        k.YB = std::fma(m.c[2].y, B.z, std::fma(m.c[1].y, B.y, m.c[0].y * B.x)) * -k.hH;
        k.WB = std::fma(m.c[2].w, B.z, std::fma(m.c[1].w, B.y, m.c[0].w * B.x));
        return k;
    }
    static float2 TapTexelWorld(const KernelProjection& k, float3 offset, float4 rotator)
    {
        float2 o = Geometry::RotateVector(rotator, offset.xy()); // per-frame uniform (host-evaluated by the kernels' launcher)
        float cx = std::fma(o.y, k.XB, std::fma(o.x, k.XT, k.X0));
        float cy = std::fma(o.y, k.YB, std::fma(o.x, k.YT, k.Y0));
        float cw = std::fma(o.y, k.WB, std::fma(o.x, k.WT, k.W0));
        float rw = 1.0f / cw; // clip.xy / clip.w as one reciprocal and two fused products
        return float2(std::fma(cx, rw, k.hW), std::fma(cy, rw, k.hH));
    }
    // :486-540
    static float GetNormalWeightParam(float nonLinearAccumSpeed, float lobeAngleFraction, float roughness = 1.0f)
    {
        float percentOfVolume = NRD_MAX_PERCENT_OF_LOBE_VOLUME * lerp(lobeAngleFraction, 1.0f, nonLinearAccumSpeed);
        float tanHalfAngle = ImportanceSampling::GetSpecularLobeTanHalfAngle(roughness, percentOfVolume);
        float angle = atan(tanHalfAngle);
        angle = max(angle, NRD_NORMAL_ENCODING_ERROR);
        return 1.0f / angle;
    }
    static float2 GetGeometryWeightParams(float planeDistSensitivity, float frustumSize, float3 Xv, float3 Nv)
    {
        float norm = planeDistSensitivity * frustumSize;
        float a = 1.0f / norm;
        float b = dot(Nv, Xv) * a;
        return float2(a, -b);
    }
    static float2 GetHitDistanceWeightParams(float hitDist, float nonLinearAccumSpeed, float roughness = 1.0f)
    {
        float smc = GetSpecMagicCurve(roughness);
        float norm = lerp(0.0005f, 1.0f, min(nonLinearAccumSpeed, smc));
        float a = 1.0f / norm;
        float b = hitDist * a;
        return float2(a, -b);
    }
    static float2 GetRoughnessWeightParams(float roughness, float fraction, float sensitivity = NRD_ROUGHNESS_SENSITIVITY)
    {
        float a = 1.0f / lerp(sensitivity, 1.0f, saturate(roughness * fraction));
        float b = roughness * a;
        return float2(a, -b);
    }
    static float2 GetRelaxedRoughnessWeightParams(float m, float fraction = 1.0f, float sensitivity = NRD_ROUGHNESS_SENSITIVITY)
    {
        float a = 1.0f / lerp(sensitivity, 1.0f, lerp(m * m, m, fraction));
        float b = m * a;
        return float2(a, -b);
    }
    // :547-569
    static float ExpApprox(float x) { return rcp(x * x - x + 1.0f); }
    static float ComputeExponentialWeight(float x, float px, float py) { return ExpApprox(-NRD_EXP_WEIGHT_DEFAULT_SCALE * abs(x * px + py)); }
    static float ComputeNonExponentialWeight(float x, float px, float py) { return Math::SmoothStep(1.0f, 0.0f, abs(x * px + py)); }
    static float ComputeNonExponentialWeightWithSigma(float x, float px, float py, float sigma) { return Math::SmoothStep(1.0f, 0.0f, abs(x * px + py) - sigma * px); }
    static float ComputeWeight(float x, float px, float py) { return ComputeNonExponentialWeight(x, px, py); }
    static float GetGaussianWeight(float r) { return exp(-0.66f * r * r); }                                                // :571-574
    // :578-590
    static float GetEncodingAwareNormalWeight(float3 Ncurr, float3 Nprev, float maxAngle, float curvatureAngle, float thresholdAngle)
    {
        float cosa = dot(Ncurr, Nprev);
        float angle = Math::AcosApprox(cosa);
        return Math::SmoothStep01(1.0f - (angle - curvatureAngle - thresholdAngle) / maxAngle);
    }
    static float GetDisocclusionThreshold(float disocclusionThreshold, float frustumSize, float NoV)                      // :594-597
    { return frustumSize * saturate(disocclusionThreshold / max(0.01f, NoV)); }
    static float GetStdDev(float m1, float m2) { return sqrt(abs(m2 - m1 * m1)); }                                         // :226

    // ---- REBLUR_Common.hlsli ---------------------------------------------------------------------
    static uint PackInternalData(float diffAccumSpeed, float specAccumSpeed, float materialID)                             // :13-24
    {
        float3 t;
        t.x = diffAccumSpeed / REBLUR_MAX_ACCUM_FRAME_NUM;
        t.y = specAccumSpeed / REBLUR_MAX_ACCUM_FRAME_NUM;
        t.z = materialID / REBLUR_MAX_MATERIALID_NUM;
        return Packing::RgbaToUint(float4(t.x, t.y, t.z, t.z), 6, 6, 4, 0);
    }
    static float3 UnpackInternalData(uint p)                                                                               // :26-33
    {
        float3 t = Packing::UintToRgba(p, 6, 6, 4, 0).xyz();
        t.x *= REBLUR_MAX_ACCUM_FRAME_NUM;
        t.y *= REBLUR_MAX_ACCUM_FRAME_NUM;
        t.z *= REBLUR_MAX_MATERIALID_NUM;
        return t;
    }
    static uint PackData2(float fbits, float curvature, float virtualHistoryAmount)                                        // :59-70
    {
        uint p = uint(fbits + 0.5f);
        p |= uint(saturate(virtualHistoryAmount) * 255.0f + 0.5f) << 8;
        p |= uint(f32tof16(curvature)) << 16;
        return p;
    }
    static float2 UnpackData2(uint p, uint& bits)                                                                          // :72-80
    {
        bits = p & 0xFF;
        float virtualHistoryAmount = float((p >> 8) & 0xFF) / 255.0f;
        float curvature = f16tof32(uint16_t(p >> 16));
        return float2(virtualHistoryAmount, curvature);
    }
    float3 GetViewVector(float3 X, bool isViewSpace = false) const                                                       // :84-87
    { return c.gOrthoMode == 0.0f ? normalize(-X) : (isViewSpace ? float3(0, 0, -1) : c.gViewVectorWorld.xyz()); }
    float3 GetViewVectorPrev(float3 Xprev, float3 cameraDelta) const                                                     // :89-92
    { return c.gOrthoMode == 0.0f ? normalize(cameraDelta - Xprev) : c.gViewVectorWorldPrev.xyz(); }
    float GetMinAllowedLimitForHitDistNonLinearAccumSpeed(float roughness) const                                          // :94-102
    {
        float frameNum = 0.5f * GetSpecMagicCurve(roughness) * c.gMaxAccumulatedFrameNum;
        return 1.0f / (1.0f + frameNum);
    }
    float GetFadeBasedOnAccumulatedFrames(float accumSpeed) const                                                         // :104-110
    {
        float a = c.gHistoryFixFrameNum * 2.0f / 3.0f + 1e-6f;
        float b = c.gHistoryFixFrameNum * 4.0f / 3.0f + 2e-6f;
        return Math::LinearStep(a, b, accumSpeed);
    }
    float GetNonLinearAccumSpeed(float accumSpeed, float maxAccumSpeed, float confidence, bool hasData) const             // :112-124
    {
        float nonLinearAccumSpeed = max(1.0f - confidence, 1.0f / (1.0f + min(accumSpeed, maxAccumSpeed)));
        if (!hasData) nonLinearAccumSpeed *= lerp(1.0f - c.gCheckerboardResolveAccumSpeed, 1.0f, nonLinearAccumSpeed);
        return nonLinearAccumSpeed;
    }
    float RemapRoughnessToResponsiveFactor(float roughness) const                                                         // :126-131
    {
        float amount = (roughness + NRD_EPS) / (c.gResponsiveAccumulationRoughnessThreshold + NRD_EPS);
        return Math::SmoothStep01(amount);
    }
    static float GetLumaScale(float currLuma, float newLuma) { return (newLuma + NRD_EPS) / (currLuma + NRD_EPS); }       // :139-146
    float4 MixHistoryAndCurrent(float4 history, float4 current, float f, float roughness = 1.0f) const                   // :199-206
    {
        float4 r;
        r.set_xyz(lerp(history.xyz(), current.xyz(), f));
        r.w = lerp(history.w, current.w, max(f, GetMinAllowedLimitForHitDistNonLinearAccumSpeed(roughness)));
        return r;
    }
    static float GetLuma(float4 input) { return input.x; }                                                                 // :211-218 (YCoCg)
    static float4 ChangeLuma(float4 input, float newLuma)                                                                  // :220-225
    {
        float s = GetLumaScale(GetLuma(input), newLuma);
        input.x *= s; input.y *= s; input.z *= s;
        return input;
    }
    static float4 ClampNegativeToZero(float4 input)                                                                        // :227-240
    {
        input.set_xyz(_NRD_YCoCgToLinear(input.xyz()));
        input.set_xyz(_NRD_LinearToYCoCg(input.xyz()));
        input.w = saturate(input.w);
        return input;
    }
    float ComputeAntilag(float history, float avg, float sigma, float accumSpeed) const                                   // :244-274 (mode 2)
    {
        float h = history, a = avg;
        float s = sigma * c.gAntilagParams.x;
        float magic = c.gAntilagParams.y * c.gFramerateScale * c.gFramerateScale;
        float hc = Color::Clamp(a, s, h);
        float d = abs(h - hc) / (max(h, hc) + NRD_EPS);
        d = 1.0f / (1.0f + d * accumSpeed / magic);
        return d;
    }
    static void GetKernelBasis(float3 D, float3 N, float3& T, float3& B)                                                   // :278-293
    {
        float3x3 basis = Geometry::GetBasis(N);
        T = basis.r[0];
        B = basis.r[1];
        if (abs(dot(D, N)) < 0.999f)
        {
            float3 R = reflect(-D, N);
            T = normalize(cross(N, R));
            B = cross(R, T);
        }
    }
    float2 GetTemporalAccumulationParams(float isInScreenMulFootprintQuality, float accumSpeed) const                     // :297-306
    {
        accumSpeed *= REBLUR_SAMPLES_PER_FRAME;
        float w = isInScreenMulFootprintQuality;
        w *= accumSpeed / (1.0f + accumSpeed);
        return float2(w, 1.0f + 3.0f * c.gFramerateScale * w);
    }

    // Common.hlsli:602-656 -- CatRom-12 with fallback to custom-weight bilinear.  color0 via bilinear sampler taps, color1
    // (optional) via 4 loads with the custom weights.
    void BicubicFilter(float2 samplePos, float2 invResourceSize, float4 bilinearCustomWeights, bool useBicubic, const Tex& tex0, float4& c0,
                       const Tex* tex1, float* c1) const
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
        int2 bilinearOrigin = int2((int)centerPos.x, (int)centerPos.y);

        float4 color = tex0.sampleLinear(uv01.xy()) * float4(w.x);
        color += tex0.sampleLinear(uv01.zw()) * float4(w.y);
        color += tex0.sampleLinear(uv23.xy()) * float4(w.z);
        color += tex0.sampleLinear(uv23.zw()) * float4(w.w);
        color += tex0.sampleLinear(uv4) * float4(w4);
        c0 = sum < 0.0001f ? float4(0.0f) : color / float4(sum);

        if (tex1)
        {
            float v = tex1->load(bilinearOrigin.x, bilinearOrigin.y).x * bilinearCustomWeights.x;
            v += tex1->load(bilinearOrigin.x + 1, bilinearOrigin.y).x * bilinearCustomWeights.y;
            v += tex1->load(bilinearOrigin.x, bilinearOrigin.y + 1).x * bilinearCustomWeights.z;
            v += tex1->load(bilinearOrigin.x + 1, bilinearOrigin.y + 1).x * bilinearCustomWeights.w;
            float s = dot(bilinearCustomWeights, float4(1.0f));
            *c1 = s < 0.0001f ? 0.0f : v / s;
        }
    }
};

// =====================================================================================================
// Spatial filters (shared by PrePass / Blur / PostBlur)
// =====================================================================================================
struct SpatialCtx
{
    int mode;             // REBLUR_PRE_BLUR / REBLUR_BLUR / REBLUR_POST_BLUR
    bool noTemporalStabilization;
    int2 pixelPos;
    float2 pixelUv;
    float viewZ, roughness, materialID, NoV, frustumSize;
    float3 N, Nv, Xv, Vv;
    float4 rotator;
    float2 data1;
    // pre-pass only
    int checkerboardPosX0, checkerboardPosX1, checkerboardPosY;
    float2 wc;
    const Tex *inViewZ, *inNormalRoughness;
};

// REBLUR_Common_DiffuseSpatialFilter.hlsli:23-213
void DiffuseSpatialFilter(const Pass& P, const SpatialCtx& s, float sum, float4 diff, const Tex& gIn_Diff, Tex& gOut_Diff, Tex* gOut_DiffCopy)
{
    const CB& c = P.c;
    const bool pre = s.mode == REBLUR_PRE_BLUR;
    if (!pre || c.gDiffPrepassBlurRadius != 0.0f)
    {
        float diffNonLinearAccumSpeed = REBLUR_PRE_BLUR_NON_LINEAR_ACCUM_SPEED;
        float fractionScale = 1.0f, radiusScale = 1.0f;
        if (s.mode == REBLUR_PRE_BLUR) fractionScale = REBLUR_PRE_BLUR_FRACTION_SCALE;
        else if (s.mode == REBLUR_BLUR) fractionScale = REBLUR_BLUR_FRACTION_SCALE;
        else { radiusScale = REBLUR_POST_BLUR_RADIUS_SCALE; fractionScale = REBLUR_POST_BLUR_FRACTION_SCALE; }

        float hitDistScale = _REBLUR_GetHitDistanceNormalization(s.viewZ, c.gHitDistParams, 1.0f);
        float hitDist = diff.w * hitDistScale;
        float hitDistFactor = Pass::GetHitDistFactor(hitDist, s.frustumSize);

        float blurRadius, areaFactor;
        if (pre)
        {
            blurRadius = c.gDiffPrepassBlurRadius;
            areaFactor = hitDistFactor;
        }
        else
        {
            float boost = 1.0f - P.GetFadeBasedOnAccumulatedFrames(s.data1.x);
            boost *= 1.0f - BRDF::Pow5(s.NoV);
            diffNonLinearAccumSpeed = 1.0f / (1.0f + REBLUR_SAMPLES_PER_FRAME * (1.0f - boost) * s.data1.x);
            blurRadius = c.gMaxBlurRadius;
            areaFactor = hitDistFactor * diffNonLinearAccumSpeed;
        }
        blurRadius *= Math::Sqrt01(areaFactor);
        blurRadius *= radiusScale;
        blurRadius = max(blurRadius, c.gMinBlurRadius);

        float2 geometryWeightParams = Pass::GetGeometryWeightParams(c.gPlaneDistSensitivity, s.frustumSize, s.Xv, s.Nv);
        float normalWeightParam = Pass::GetNormalWeightParam(diffNonLinearAccumSpeed, c.gLobeAngleFraction) / fractionScale;
        float2 hitDistanceWeightParams = Pass::GetHitDistanceWeightParams(diff.w, diffNonLinearAccumSpeed);
        float minHitDistWeight = c.gMinHitDistanceWeight * fractionScale;
        if (!pre) minHitDistWeight *= sqrt(diffNonLinearAccumSpeed);

        // screen-space sampling (REBLUR_USE_SCREEN_SPACE_SAMPLING_FOR_DIFFUSE = 1)
        float2 skew(1.0f);
        if (!pre)
        {
            skew = lerp(float2(1.0f) - abs(s.Nv.xy()), float2(1.0f), s.NoV);
            skew /= float2(max(skew.x, skew.y));
        }
        skew *= float2(blurRadius); // in pixels: uv * rectSize is evaluated directly (Pass::TapTexelScreen)
        float4 scaledRotator = Geometry::ScaleRotator(s.rotator, skew);

        for (uint n = 0; n < (P.perf ? 6u : 8u); n++)
        {
            float3 offset = P.perf ? g_Special6[n] : g_Special8[n];
            float2 uv = floor(Pass::TapTexelScreen(s.pixelPos, scaledRotator, offset.xy())) + float2(0.5f);
            if (pre) uv = Pass::ApplyCheckerboardShift(uv, c.gDiffCheckerboard, n, c.gFrameIndex);
            uv *= c.gRectSizeInv;

            float2 uvScaled = P.ClampUvToViewport(uv);
            float2 checkerboardUvScaled = uvScaled;
            if (pre && c.gDiffCheckerboard != 2) checkerboardUvScaled.x *= 0.5f;

            float zs = P.UnpackViewZ(s.inViewZ->sampleNearest(uvScaled).x);
            float materialIDs;
            float4 Ns = NRD_FrontEnd_UnpackNormalAndRoughness(s.inNormalRoughness->sampleNearest(uvScaled), materialIDs);

            float angle = Math::AcosApprox(dot(s.N, Ns.xyz()));
            float3 Xvs = Geometry::ReconstructViewPosition(uv, c.gFrustum, zs, c.gOrthoMode);

            float w = Pass::IsInScreenNearest(uv);
            w *= Pass::ComputeWeight(dot(s.Nv, Xvs), geometryWeightParams.x, geometryWeightParams.y);
            w *= float(max(s.materialID, c.gDiffMinMaterial) == max(materialIDs, c.gDiffMinMaterial));
            w *= Pass::ComputeWeight(angle, normalWeightParam, 0.0f);

            float4 sv = gIn_Diff.sampleNearest(checkerboardUvScaled);
            sv = w == 0.0f ? float4(0.0f) : sv;

            w *= lerp(minHitDistWeight, 1.0f, Pass::ComputeExponentialWeight(sv.w, hitDistanceWeightParams.x, hitDistanceWeightParams.y));
            w *= Pass::GetGaussianWeight(offset.z);

            sum += w;
            diff += sv * float4(w);
        }
        float invSum = Math::PositiveRcp(sum);
        diff *= float4(invSum);
    }
    if (pre && sum == 0.0f)
    {
        float4 s0 = gIn_Diff.load(s.checkerboardPosX0, s.checkerboardPosY);
        float4 s1 = gIn_Diff.load(s.checkerboardPosX1, s.checkerboardPosY);
        s0 = s.wc.x == 0.0f ? float4(0.0f) : s0;
        s1 = s.wc.y == 0.0f ? float4(0.0f) : s1;
        diff = s0 * float4(s.wc.x) + s1 * float4(s.wc.y);
    }
    gOut_Diff.store(s.pixelPos, diff);
    if (s.noTemporalStabilization && gOut_DiffCopy) gOut_DiffCopy->store(s.pixelPos, diff);
}

// REBLUR_Common_SpecularSpatialFilter.hlsli:23-260
void SpecularSpatialFilter(const Pass& P, const SpatialCtx& s, float sum, float4 spec, const Tex& gIn_Spec, Tex& gOut_Spec, Tex* gOut_SpecCopy,
                           Tex* gOut_SpecHitDistForTracking)
{
    const CB& c = P.c;
    const bool pre = s.mode == REBLUR_PRE_BLUR;
    float smc = Pass::GetSpecMagicCurve(s.roughness);
    if (!pre || c.gSpecPrepassBlurRadius != 0.0f)
    {
        RngHash rng;
        if (pre) rng.Initialize(s.pixelPos, c.gFrameIndex);
        float specNonLinearAccumSpeed = REBLUR_PRE_BLUR_NON_LINEAR_ACCUM_SPEED;
        float fractionScale = 1.0f, radiusScale = 1.0f;
        if (s.mode == REBLUR_PRE_BLUR) fractionScale = REBLUR_PRE_BLUR_FRACTION_SCALE;
        else if (s.mode == REBLUR_BLUR) fractionScale = REBLUR_BLUR_FRACTION_SCALE;
        else { radiusScale = REBLUR_POST_BLUR_RADIUS_SCALE; fractionScale = REBLUR_POST_BLUR_FRACTION_SCALE; }

        float4 Dv = ImportanceSampling::GetSpecularDominantDirection(s.Nv, s.Vv, s.roughness);
        float NoD = abs(dot(s.Nv, Dv.xyz()));
        float hitDistScale = _REBLUR_GetHitDistanceNormalization(s.viewZ, c.gHitDistParams, s.roughness);
        float hitDist = spec.w * hitDistScale;
        float hitDistFactor = Pass::GetHitDistFactor(hitDist, s.frustumSize);

        float hitDistForTracking = 0.0f;
        float blurRadius, areaFactor;
        if (pre)
        {
            hitDistForTracking = hitDist == 0.0f ? NRD_INF : hitDist;
            blurRadius = c.gSpecPrepassBlurRadius;
            areaFactor = s.roughness * hitDistFactor;
        }
        else
        {
            float boost = 1.0f - P.GetFadeBasedOnAccumulatedFrames(s.data1.y);
            boost *= 1.0f - BRDF::Pow5(s.NoV);
            boost *= smc;
            specNonLinearAccumSpeed = 1.0f / (1.0f + REBLUR_SAMPLES_PER_FRAME * (1.0f - boost) * s.data1.y);
            blurRadius = c.gMaxBlurRadius;
            areaFactor = s.roughness * hitDistFactor * specNonLinearAccumSpeed;
        }
        blurRadius *= Math::Sqrt01(areaFactor);
        if (pre)
        {
            float lobeTanHalfAngle = ImportanceSampling::GetSpecularLobeTanHalfAngle(s.roughness, REBLUR_MAX_PERCENT_OF_LOBE_VOLUME_FOR_PRE_PASS);
            float lobeRadius = hitDist * NoD * lobeTanHalfAngle;
            float minBlurRadius = lobeRadius / P.PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, 1.0f, s.viewZ + hitDist * Dv.w);
            blurRadius = min(blurRadius, minBlurRadius);
        }
        blurRadius *= radiusScale;
        blurRadius = max(blurRadius, c.gMinBlurRadius * smc);

        float roughnessFractionScaled = saturate(c.gRoughnessFraction * fractionScale);
        float2 geometryWeightParams = Pass::GetGeometryWeightParams(c.gPlaneDistSensitivity, s.frustumSize, s.Xv, s.Nv);
        float normalWeightParam = Pass::GetNormalWeightParam(specNonLinearAccumSpeed, c.gLobeAngleFraction, s.roughness) / fractionScale;
        float2 roughnessWeightParams = Pass::GetRoughnessWeightParams(s.roughness, roughnessFractionScaled);
        float2 hitDistanceWeightParams = Pass::GetHitDistanceWeightParams(spec.w, specNonLinearAccumSpeed, s.roughness);
        float minHitDistWeight = c.gMinHitDistanceWeight * fractionScale * smc;
        if (!pre) minHitDistWeight *= sqrt(specNonLinearAccumSpeed);

        float4 scaledRotator(0.0f);
        float3 Tv(0.0f), Bv(0.0f);
        const bool screenSpace = pre || P.perf; // REBLUR_USE_SCREEN_SPACE_SAMPLING_FOR_SPECULAR = 1 in performance mode
        if (screenSpace)
        {
            float2 skew(1.0f);
            skew *= float2(blurRadius); // in pixels (Pass::TapTexelScreen)
            scaledRotator = Geometry::ScaleRotator(s.rotator, skew);
        }
        else
        {
            // world-space sampling (REBLUR_USE_SCREEN_SPACE_SAMPLING_FOR_SPECULAR = 0)
            float bentFactor = sqrt(hitDistFactor);
            float skewFactor = lerp(0.25f + 0.75f * s.roughness, 1.0f, NoD);
            skewFactor = lerp(skewFactor, 1.0f, specNonLinearAccumSpeed);
            skewFactor = lerp(1.0f, skewFactor, bentFactor);
            float3 bentDv = normalize(lerp(s.Nv, Dv.xyz(), bentFactor));
            Pass::GetKernelBasis(bentDv, s.Nv, Tv, Bv);
            float worldRadius = P.PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, blurRadius, s.viewZ);
            Tv *= float3(worldRadius * skewFactor);
            Bv *= float3(worldRadius / skewFactor);
        }
        const Pass::KernelProjection kernelProjection = Pass::ProjectKernel(c.gViewToClip, c.gRectSize, s.Xv, Tv, Bv);

        for (uint n = 0; n < (P.perf ? 6u : 8u); n++)
        {
            float3 offset = P.perf ? g_Special6[n] : g_Special8[n];
            float2 uv;
            if (screenSpace) uv = Pass::TapTexelScreen(s.pixelPos, scaledRotator, offset.xy());
            else uv = Pass::TapTexelWorld(kernelProjection, offset, s.rotator);
            uv = floor(uv) + float2(0.5f);
            if (pre) uv = Pass::ApplyCheckerboardShift(uv, c.gSpecCheckerboard, n, c.gFrameIndex);
            uv *= c.gRectSizeInv;

            float2 uvScaled = P.ClampUvToViewport(uv);
            float2 checkerboardUvScaled = uvScaled;
            if (pre && c.gSpecCheckerboard != 2) checkerboardUvScaled.x *= 0.5f;

            float zs = P.UnpackViewZ(s.inViewZ->sampleNearest(uvScaled).x);
            float materialIDs;
            float4 Ns = NRD_FrontEnd_UnpackNormalAndRoughness(s.inNormalRoughness->sampleNearest(uvScaled), materialIDs);

            float angle = Math::AcosApprox(dot(s.N, Ns.xyz()));
            float3 Xvs = Geometry::ReconstructViewPosition(uv, c.gFrustum, zs, c.gOrthoMode);

            float w = Pass::IsInScreenNearest(uv);
            w *= Pass::ComputeWeight(dot(s.Nv, Xvs), geometryWeightParams.x, geometryWeightParams.y);
            w *= float(max(s.materialID, c.gSpecMinMaterial) == max(materialIDs, c.gSpecMinMaterial));
            w *= Pass::ComputeWeight(angle, normalWeightParam, 0.0f);
            w *= Pass::ComputeWeight(Ns.w, roughnessWeightParams.x, roughnessWeightParams.y);

            float4 sv = gIn_Spec.sampleNearest(checkerboardUvScaled);
            sv = w == 0.0f ? float4(0.0f) : sv;

            if (pre)
            {
                float hs = sv.w * _REBLUR_GetHitDistanceNormalization(zs, c.gHitDistParams, Ns.w);
                float d = length(Xvs - s.Xv) + NRD_EPS;
                float geometryWeight = w * saturate(hs / d);
                if (rng.GetFloat() < geometryWeight) hitDistForTracking = min(hitDistForTracking, hs);
                w *= c.gUsePrepassNotOnlyForSpecularMotionEstimation;
                float t = hs / (d + hitDist);
                w *= lerp(saturate(t), 1.0f, Math::LinearStep(0.5f, 1.0f, s.roughness));
            }
            w *= lerp(minHitDistWeight, 1.0f, Pass::ComputeExponentialWeight(sv.w, hitDistanceWeightParams.x, hitDistanceWeightParams.y));
            w *= Pass::GetGaussianWeight(offset.z);

            sum += w;
            spec += sv * float4(w);
        }
        float invSum = Math::PositiveRcp(sum);
        spec *= float4(invSum);

        if (pre) gOut_SpecHitDistForTracking->store(s.pixelPos, hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking);
    }
    if (pre && sum == 0.0f)
    {
        float4 s0 = gIn_Spec.load(s.checkerboardPosX0, s.checkerboardPosY);
        float4 s1 = gIn_Spec.load(s.checkerboardPosX1, s.checkerboardPosY);
        s0 = s.wc.x == 0.0f ? float4(0.0f) : s0;
        s1 = s.wc.y == 0.0f ? float4(0.0f) : s1;
        spec = s0 * float4(s.wc.x) + s1 * float4(s.wc.y);
    }
    gOut_Spec.store(s.pixelPos, spec);
    if (s.noTemporalStabilization && gOut_SpecCopy) gOut_SpecCopy->store(s.pixelPos, spec);
}

// =====================================================================================================
// Passes.  `t` is the binding list of the dispatch: inputs then outputs, in *.resources.hlsli order.
// =====================================================================================================
struct Signals { bool diff, spec; };

void ClassifyTiles(const Pass& P, Tex* t, int gridW, int gridH)
{
    const Tex& gIn_ViewZ = t[0];
    Tex& gOut_Tiles = t[1];
#pragma omp parallel for schedule(static)
    for (int ty = 0; ty < gridH; ty++)
        for (int tx = 0; tx < gridW; tx++)
        {
            int sum = 0;
            for (int j = 0; j < 16; j++)
                for (int i = 0; i < 16; i++)
                {
                    float viewZ = P.UnpackViewZ(gIn_ViewZ.load(tx * 16 + i, ty * 16 + j).x);
                    sum += viewZ > P.c.gDenoisingRange ? 1 : 0;
                }
            gOut_Tiles.store(tx, ty, float4(sum == 256 ? 1.0f : 0.0f, 0, 0, 0));
        }
}

// common prologue of PrePass / Blur / PostBlur; returns false on early out
bool SpatialPrologue(const Pass& P, SpatialCtx& s, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, float4* normalAndRoughnessPacked,
                     float4 baseRotator)
{
    const CB& c = P.c;
    float isSky = gIn_Tiles.load(s.pixelPos.x >> 4, s.pixelPos.y >> 4).x;
    if (isSky != 0.0f || s.pixelPos.x > c.gRectSizeMinusOne[0] || s.pixelPos.y > c.gRectSizeMinusOne[1]) return false;
    (void)gIn_ViewZ;
    float4 packed = gIn_Normal_Roughness.load(s.pixelPos);
    if (normalAndRoughnessPacked) *normalAndRoughnessPacked = packed;
    float4 normalAndRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(packed, s.materialID);
    s.N = normalAndRoughness.xyz();
    s.Nv = Geometry::RotateVectorInverse(c.gViewToWorld, s.N);
    s.roughness = normalAndRoughness.w;
    s.pixelUv = (tofloat(s.pixelPos) + float2(0.5f)) * c.gRectSizeInv;
    s.Xv = Geometry::ReconstructViewPosition(s.pixelUv, c.gFrustum, s.viewZ, c.gOrthoMode);
    s.Vv = P.GetViewVector(s.Xv, true);
    s.NoV = abs(dot(s.Nv, s.Vv));
    s.frustumSize = P.GetFrustumSize(c.gMinRectDimMulUnproject, c.gOrthoMode, s.viewZ);
    s.rotator = baseRotator; // NRD_FRAME rotator mode: one rotator per frame (Common.hlsli:255-278)
    return true;
}

// REBLUR_HitDistReconstruction.hlsli:10-155 (REBLUR_USE_DECOMPRESSED_HIT_DIST_IN_RECONSTRUCTION = 0, not performance mode):
// a pixel whose ray missed (hit distance 0) borrows the hit distance of its 3x3 / 5x5 neighbourhood
void HitDistReconstruction(const Pass& P, Signals sg, int border, Tex* t, int W, int H)
{
    const CB& c = P.c;
    int k = 0;
    const Tex& gIn_Tiles = t[k++];
    const Tex& gIn_Normal_Roughness = t[k++];
    const Tex& gIn_ViewZ = t[k++];
    const Tex* gIn_Diff = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_Spec = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_Diff = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_Spec = sg.spec ? &t[k++] : nullptr;
    const int2 rectMax(c.gRectSizeMinusOne[0], c.gRectSizeMinusOne[1]);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x > rectMax.x || y > rectMax.y) continue;
            // "shared memory" of the reference: clamped loads (Preload :13-41)
            auto sData = [&](int i, int j, float4& normalAndRoughness) {
                int2 p = clamp(int2(x + i - border, y + j - border), int2(0), rectMax);
                normalAndRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(p));
                float3 d;
                d.x = gIn_Diff ? gIn_Diff->load(p).w : 0.0f;
                d.y = gIn_Spec ? gIn_Spec->load(p).w : 0.0f;
                d.z = P.UnpackViewZ(gIn_ViewZ.load(p).x);
                return d;
            };
            float4 normalAndRoughness;
            float3 center = sData(border, border, normalAndRoughness);
            if (center.z > c.gDenoisingRange) continue;
            float3 N = normalAndRoughness.xyz();
            float roughness = normalAndRoughness.w;
            float2 pixelUv = (tofloat(pixelPos) + float2(0.5f)) * c.gRectSizeInv;
            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, c.gFrustum, center.z, c.gOrthoMode);
            float3 Nv = Geometry::RotateVectorInverse(c.gViewToWorld, N);
            float frustumSize = P.GetFrustumSize(c.gMinRectDimMulUnproject, c.gOrthoMode, center.z);
            float2 geometryWeightParams = Pass::GetGeometryWeightParams(c.gPlaneDistSensitivity, frustumSize, Xv, Nv);
            float2 relaxedRoughnessWeightParams = Pass::GetRelaxedRoughnessWeightParams(roughness * roughness);
            float diffNormalWeightParam = Pass::GetNormalWeightParam(1.0f, 1.0f);
            float specNormalWeightParam = Pass::GetNormalWeightParam(1.0f, 1.0f, roughness);

            float2 sum = float2(1000.0f) * float2(float(center.x != 0.0f), float(center.y != 0.0f));
            float2 acc = float2(center.x, center.y) * sum;
            for (int j = 0; j <= border * 2; j++)
                for (int i = 0; i <= border * 2; i++)
                {
                    float2 o = float2(float(i), float(j)) - float2(float(border));
                    if (o.x == 0.0f && o.y == 0.0f) continue;
                    float4 nr;
                    float3 data = sData(i, j, nr);
                    float2 uv = pixelUv + o * c.gRectSizeInv;
                    float w = Pass::IsInScreenNearest(uv);
                    w *= Pass::GetGaussianWeight(length(o) * 0.5f);
                    float3 Xvs = Geometry::ReconstructViewPosition(uv, c.gFrustum, data.z, c.gOrthoMode);
                    w *= Pass::ComputeWeight(dot(Nv, Xvs), geometryWeightParams.x, geometryWeightParams.y);
                    float2 ww(w);
                    if (!P.perf) // :106-119
                    {
                        float cosa = dot(N, nr.xyz());
                        float angle = Math::AcosApprox(cosa);
                        ww.x *= Pass::ComputeExponentialWeight(angle, diffNormalWeightParam, 0.0f);
                        ww.y *= Pass::ComputeExponentialWeight(angle, specNormalWeightParam, 0.0f);
                        ww.y *= Pass::ComputeExponentialWeight(nr.w * nr.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);
                    }
                    data.x = ww.x == 0.0f ? 0.0f : data.x; // Denanify
                    data.y = ww.y == 0.0f ? 0.0f : data.y;
                    ww *= float2(float(data.x != 0.0f), float(data.y != 0.0f));
                    acc += float2(data.x, data.y) * ww;
                    sum += ww;
                }
            acc /= max(sum, float2(NRD_EPS));
            if (gOut_Diff)
            {
                float4 d = gIn_Diff->load(pixelPos);
                gOut_Diff->store(pixelPos, float4(d.x, d.y, d.z, acc.x));
            }
            if (gOut_Spec)
            {
                float4 sp = gIn_Spec->load(pixelPos);
                gOut_Spec->store(pixelPos, float4(sp.x, sp.y, sp.z, acc.y));
            }
        }
}

void PrePass(const Pass& P, Signals sg, Tex* t, int W, int H)
{
    const CB& c = P.c;
    int k = 0;
    const Tex& gIn_Tiles = t[k++];
    const Tex& gIn_Normal_Roughness = t[k++];
    const Tex& gIn_ViewZ = t[k++];
    const Tex* gIn_Diff = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_Spec = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_Diff = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_Spec = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_SpecHitDistForTracking = sg.spec ? &t[k++] : nullptr;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            SpatialCtx s{};
            s.mode = REBLUR_PRE_BLUR;
            s.pixelPos = int2(x, y);
            s.inViewZ = &gIn_ViewZ;
            s.inNormalRoughness = &gIn_Normal_Roughness;
            // (tile / rect test first, then viewZ: REBLUR_PrePass.hlsli:16-24)
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x > c.gRectSizeMinusOne[0] || y > c.gRectSizeMinusOne[1]) continue;
            s.viewZ = P.UnpackViewZ(gIn_ViewZ.load(x, y).x);
            if (s.viewZ > c.gDenoisingRange) continue;
            if (!SpatialPrologue(P, s, gIn_Tiles, gIn_Normal_Roughness, gIn_ViewZ, nullptr, c.gRotatorPre)) continue;

            uint checkerboard = Sequence::CheckerBoard(s.pixelPos, c.gFrameIndex);
            int cx0 = max(x - 1, 0), cx1 = min(x + 1, c.gRectSizeMinusOne[0]);
            float viewZ0 = P.UnpackViewZ(gIn_ViewZ.load(cx0, y).x);
            float viewZ1 = P.UnpackViewZ(gIn_ViewZ.load(cx1, y).x);
            float thr = Pass::GetDisocclusionThreshold(NRD_DISOCCLUSION_THRESHOLD, s.frustumSize, s.NoV);
            float2 wc = float2(step(abs(viewZ0 - s.viewZ), thr), step(abs(viewZ1 - s.viewZ), thr));
            wc.x = (viewZ0 > c.gDenoisingRange || x < 1) ? 0.0f : wc.x;
            wc.y = (viewZ1 > c.gDenoisingRange || x >= c.gRectSizeMinusOne[0]) ? 0.0f : wc.y;
            wc *= float2(Math::PositiveRcp(wc.x + wc.y));
            s.wc = wc;
            s.checkerboardPosX0 = cx0 >> 1;
            s.checkerboardPosX1 = cx1 >> 1;
            s.checkerboardPosY = y;

            if (sg.diff)
            {
                int px = x >> (c.gDiffCheckerboard == 2 ? 0 : 1);
                float sum = 1.0f;
                float4 diff = gIn_Diff->load(px, y);
                if (c.gDiffCheckerboard != 2 && checkerboard != c.gDiffCheckerboard) { sum = 0; diff = float4(0.0f); }
                DiffuseSpatialFilter(P, s, sum, diff, *gIn_Diff, *gOut_Diff, nullptr);
            }
            if (sg.spec)
            {
                int px = x >> (c.gSpecCheckerboard == 2 ? 0 : 1);
                float sum = 1.0f;
                float4 spec = gIn_Spec->load(px, y);
                if (c.gSpecCheckerboard != 2 && checkerboard != c.gSpecCheckerboard) { sum = 0; spec = float4(0.0f); }
                SpecularSpatialFilter(P, s, sum, spec, *gIn_Spec, *gOut_Spec, nullptr, gOut_SpecHitDistForTracking);
            }
        }
}

void Blur(const Pass& P, Signals sg, Tex* t, int W, int H)
{
    const CB& c = P.c;
    int k = 0;
    const Tex& gIn_Tiles = t[k++];
    const Tex& gIn_Normal_Roughness = t[k++];
    const Tex& gIn_Data1 = t[k++];
    const Tex* gIn_Diff = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_Spec = sg.spec ? &t[k++] : nullptr;
    const Tex& gIn_ViewZ = t[k++];
    Tex* gOut_Diff = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_Spec = sg.spec ? &t[k++] : nullptr;
    Tex& gOut_ViewZ = t[k++];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x > c.gRectSizeMinusOne[0] || y > c.gRectSizeMinusOne[1]) continue;
            float viewZpacked = gIn_ViewZ.load(x, y).x;
            gOut_ViewZ.store(int2(x, y), viewZpacked);
            SpatialCtx s{};
            s.mode = REBLUR_BLUR;
            s.pixelPos = int2(x, y);
            s.inViewZ = &gIn_ViewZ;
            s.inNormalRoughness = &gIn_Normal_Roughness;
            s.viewZ = P.UnpackViewZ(viewZpacked);
            if (s.viewZ > c.gDenoisingRange) continue;
            if (!SpatialPrologue(P, s, gIn_Tiles, gIn_Normal_Roughness, gIn_ViewZ, nullptr, c.gRotator)) continue;
            float4 d1 = gIn_Data1.load(x, y);
            // UnpackData1 (REBLUR_Common.hlsli:49-57): single-signal denoisers alias .y = .x
            s.data1 = (sg.diff && sg.spec) ? float2(d1.x, d1.y) * float2(REBLUR_MAX_ACCUM_FRAME_NUM) : float2(d1.x, d1.x) * float2(REBLUR_MAX_ACCUM_FRAME_NUM);
            if (sg.diff) DiffuseSpatialFilter(P, s, 1.0f, gIn_Diff->load(x, y), *gIn_Diff, *gOut_Diff, nullptr);
            if (sg.spec) SpecularSpatialFilter(P, s, 1.0f, gIn_Spec->load(x, y), *gIn_Spec, *gOut_Spec, nullptr, nullptr);
        }
}

void PostBlur(const Pass& P, Signals sg, bool noTS, Tex* t, int W, int H)
{
    const CB& c = P.c;
    int k = 0;
    const Tex& gIn_Tiles = t[k++];
    const Tex& gIn_Normal_Roughness = t[k++];
    const Tex& gIn_Data1 = t[k++];
    const Tex* gIn_Diff = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_Spec = sg.spec ? &t[k++] : nullptr;
    const Tex& gIn_ViewZ = t[k++];
    Tex& gOut_Normal_Roughness = t[k++];
    Tex* gOut_Diff = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_Spec = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_InternalData = noTS ? &t[k++] : nullptr;
    Tex* gOut_DiffCopy = (noTS && sg.diff) ? &t[k++] : nullptr;
    Tex* gOut_SpecCopy = (noTS && sg.spec) ? &t[k++] : nullptr;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x > c.gRectSizeMinusOne[0] || y > c.gRectSizeMinusOne[1]) continue;
            SpatialCtx s{};
            s.mode = REBLUR_POST_BLUR;
            s.noTemporalStabilization = noTS;
            s.pixelPos = int2(x, y);
            s.inViewZ = &gIn_ViewZ;
            s.inNormalRoughness = &gIn_Normal_Roughness;
            s.viewZ = P.UnpackViewZ(gIn_ViewZ.load(x, y).x);
            if (s.viewZ > c.gDenoisingRange) continue;
            float4 packed;
            if (!SpatialPrologue(P, s, gIn_Tiles, gIn_Normal_Roughness, gIn_ViewZ, &packed, c.gRotatorPost)) continue;
            float4 d1 = gIn_Data1.load(x, y);
            s.data1 = (sg.diff && sg.spec) ? float2(d1.x, d1.y) * float2(REBLUR_MAX_ACCUM_FRAME_NUM) : float2(d1.x, d1.x) * float2(REBLUR_MAX_ACCUM_FRAME_NUM);
            gOut_Normal_Roughness.store(s.pixelPos, packed);
            if (noTS) gOut_InternalData->storeu(s.pixelPos, Pass::PackInternalData(s.data1.x + 1.0f, s.data1.y + 1.0f, s.materialID));
            if (sg.diff) DiffuseSpatialFilter(P, s, 1.0f, gIn_Diff->load(x, y), *gIn_Diff, *gOut_Diff, gOut_DiffCopy);
            if (sg.spec) SpecularSpatialFilter(P, s, 1.0f, gIn_Spec->load(x, y), *gIn_Spec, *gOut_Spec, gOut_SpecCopy, nullptr);
        }
}

// REBLUR_TemporalAccumulation.hlsli
void TemporalAccumulation(const Pass& P, Signals sg, Tex* t, int W, int H)
{
    const CB& c = P.c;
    int k = 0;
    const Tex& gIn_Tiles = t[k++];
    const Tex& gIn_Normal_Roughness = t[k++];
    const Tex& gIn_ViewZ = t[k++];
    const Tex& gIn_Mv = t[k++];
    const Tex& gPrev_ViewZ = t[k++];
    const Tex& gPrev_Normal_Roughness = t[k++];
    const Tex& gPrev_InternalData = t[k++];
    const Tex& gIn_DisocclusionThresholdMix = t[k++];
    const Tex* gIn_DiffConfidence = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_SpecConfidence = sg.spec ? &t[k++] : nullptr;
    const Tex* gIn_Diff = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_Spec = sg.spec ? &t[k++] : nullptr;
    const Tex *gHistory_Diff = nullptr, *gHistory_Spec = nullptr, *gHistory_DiffFast = nullptr, *gHistory_SpecFast = nullptr;
    if (sg.diff && sg.spec) { gHistory_Diff = &t[k++]; gHistory_Spec = &t[k++]; gHistory_DiffFast = &t[k++]; gHistory_SpecFast = &t[k++]; }
    else if (sg.diff) { gHistory_Diff = &t[k++]; gHistory_DiffFast = &t[k++]; }
    else { gHistory_Spec = &t[k++]; gHistory_SpecFast = &t[k++]; }
    const Tex* gPrev_SpecHitDistForTracking = sg.spec ? &t[k++] : nullptr;
    const Tex* gIn_SpecHitDistForTracking = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_Diff = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_Spec = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_DiffFast = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_SpecFast = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_SpecHitDistForTracking = sg.spec ? &t[k++] : nullptr;
    Tex& gOut_Data1 = t[k++];
    Tex& gOut_Data2 = t[k++];

    const int2 rectMax = int2(c.gRectSizeMinusOne[0], c.gRectSizeMinusOne[1]);
    const float almostZeroAngle = REBLUR_ALMOST_ZERO_ANGLE();

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x > rectMax.x || y > rectMax.y) continue;
            float viewZ = P.UnpackViewZ(gIn_ViewZ.load(x, y).x);
            if (viewZ > c.gDenoisingRange) continue;

            // "shared memory" taps (Preload :14-37): clamped loads
            auto sNormalRoughness = [&](int i, int j) { // i, j in [0, 2], centre = (1, 1)
                int2 p = clamp(int2(x + i - 1, y + j - 1), int2(0), rectMax);
                return NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(p));
            };
            auto sHitDistForTracking = [&](int i, int j) {
                int2 p = clamp(int2(x + i - 1, y + j - 1), int2(0), rectMax);
                float4 spec = gIn_Spec->load(p);
                float hitDist = c.gSpecPrepassBlurRadius == 0.0f ? spec.w : gIn_SpecHitDistForTracking->load(p).x;
                return hitDist == 0.0f ? NRD_INF : hitDist;
            };

            float2 pixelUv = (tofloat(pixelPos) + float2(0.5f)) * c.gRectSizeInv;
            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, c.gFrustum, viewZ, c.gOrthoMode);
            float3 X = Geometry::RotateVector(c.gViewToWorld, Xv);

            float3 Navg(0.0f);
            float hitDistForTracking = NRD_INF, roughnessM1 = 0.0f, roughnessM2 = 0.0f;
            for (int j = 0; j <= 2; j++)
                for (int i = 0; i <= 2; i++)
                {
                    float4 nr = sNormalRoughness(i, j);
                    if (i < 2 && j < 2) Navg += nr.xyz();
                    if (sg.spec)
                    {
                        hitDistForTracking = min(hitDistForTracking, sHitDistForTracking(i, j));
                        float roughnessSq = nr.w * nr.w;
                        roughnessM1 += roughnessSq;
                        roughnessM2 += roughnessSq * roughnessSq;
                    }
                }
            Navg /= float3(4.0f);

            float materialID;
            float4 normalAndRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos), materialID);
            float3 N = normalAndRoughness.xyz();
            float roughness = normalAndRoughness.w;

            float roughnessModified = 0.0f, roughnessSigma = 0.0f, hitDistNormalization = 0.0f;
            RngHash rng;
            if (sg.spec)
            {
                roughnessModified = Filtering::GetModifiedRoughnessFromNormalVariance(roughness, Navg);
                roughnessM1 /= 9.0f;
                roughnessM2 /= 9.0f;
                roughnessSigma = Pass::GetStdDev(roughnessM1, roughnessM2);
                rng.Initialize(pixelPos, c.gFrameIndex);
                hitDistForTracking = hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking;
                hitDistNormalization = _REBLUR_GetHitDistanceNormalization(viewZ, c.gHitDistParams, roughness);
                hitDistForTracking *= c.gSpecPrepassBlurRadius == 0.0f ? hitDistNormalization : 1.0f;
                gOut_SpecHitDistForTracking->store(pixelPos, hitDistForTracking);
            }

            // previous position and surface motion uv (:130-150)
            float4 mvRaw = gIn_Mv.load(pixelPos);
            float3 mv = mvRaw.xyz() * c.gMvScale.xyz();
            float3 Xprev = X;
            float2 smbPixelUv = pixelUv + mv.xy();
            if (c.gMvScale.w == 0.0f)
            {
                if (c.gMvScale.z == 0.0f) mv.z = Geometry::AffineTransform(c.gWorldToViewPrev, X).z - viewZ;
                float viewZprev = viewZ + mv.z;
                float3 Xvprevlocal = Geometry::ReconstructViewPosition(smbPixelUv, c.gFrustumPrev, viewZprev, c.gOrthoMode);
                Xprev = Geometry::RotateVectorInverse(c.gWorldToViewPrev, Xvprevlocal) + c.gCameraDelta.xyz();
            }
            else
            {
                Xprev += mv;
                smbPixelUv = Geometry::GetScreenUv(c.gWorldToClipPrev, Xprev);
            }

            // previous viewZ 4x4 (:152-176).  Gather returns (x,y,z,w) = (0,1),(1,1),(1,0),(0,0); .wzxy = 00,10,01,11
            Filtering::CatmullRom smbCatromFilter = Filtering::GetCatmullRomFilter(smbPixelUv, c.gRectSizePrev);
            float2 smbCatromGatherUv = smbCatromFilter.origin * c.gResourceSizeInvPrev;
            auto wzxy = [](float4 g) { return float4(g.w, g.z, g.x, g.y); };
            float4 smbViewZ0 = wzxy(gPrev_ViewZ.gather(smbCatromGatherUv, 0, int2(1, 1)));
            float4 smbViewZ1 = wzxy(gPrev_ViewZ.gather(smbCatromGatherUv, 0, int2(3, 1)));
            float4 smbViewZ2 = wzxy(gPrev_ViewZ.gather(smbCatromGatherUv, 0, int2(1, 3)));
            float4 smbViewZ3 = wzxy(gPrev_ViewZ.gather(smbCatromGatherUv, 0, int2(3, 3)));
            float3 prevViewZ0 = float3(P.UnpackViewZ(smbViewZ0.y), P.UnpackViewZ(smbViewZ0.z), P.UnpackViewZ(smbViewZ0.w));
            float3 prevViewZ1 = float3(P.UnpackViewZ(smbViewZ1.x), P.UnpackViewZ(smbViewZ1.z), P.UnpackViewZ(smbViewZ1.w));
            float3 prevViewZ2 = float3(P.UnpackViewZ(smbViewZ2.x), P.UnpackViewZ(smbViewZ2.y), P.UnpackViewZ(smbViewZ2.w));
            float3 prevViewZ3 = float3(P.UnpackViewZ(smbViewZ3.x), P.UnpackViewZ(smbViewZ3.y), P.UnpackViewZ(smbViewZ3.z));

            // previous normal averaged over the 2x2 footprint (:178-204)
            Filtering::Bilinear smbBilinearFilter = Filtering::GetBilinearFilter(smbPixelUv, c.gRectSizePrev);
            float3 smbNavg(0.0f);
            {
                // uint2( origin ): D3D float->uint conversion saturates, negative -> 0
                int px = (int)max(smbBilinearFilter.origin.x, 0.0f);
                int py = (int)max(smbBilinearFilter.origin.y, 0.0f);
                float sum = 0.0f;
                float w = float(prevViewZ0.z < c.gDenoisingRange);
                smbNavg = NRD_FrontEnd_UnpackNormalAndRoughness(gPrev_Normal_Roughness.load(px, py)).xyz() * float3(w);
                sum += w;
                w = float(prevViewZ1.y < c.gDenoisingRange);
                smbNavg += NRD_FrontEnd_UnpackNormalAndRoughness(gPrev_Normal_Roughness.load(px + 1, py)).xyz() * float3(w);
                sum += w;
                w = float(prevViewZ2.y < c.gDenoisingRange);
                smbNavg += NRD_FrontEnd_UnpackNormalAndRoughness(gPrev_Normal_Roughness.load(px, py + 1)).xyz() * float3(w);
                sum += w;
                w = float(prevViewZ3.x < c.gDenoisingRange);
                smbNavg += NRD_FrontEnd_UnpackNormalAndRoughness(gPrev_Normal_Roughness.load(px + 1, py + 1)).xyz() * float3(w);
                sum += w;
                smbNavg /= float3(sum == 0.0f ? 1.0f : sum);
            }
            smbNavg = Geometry::RotateVector(c.gWorldPrevToWorld, smbNavg);

            // parallax (:206-211)
            float smbParallaxInPixels1 = Pass::ComputeParallaxInPixels(Xprev + c.gCameraDelta.xyz(), c.gOrthoMode == 0.0f ? smbPixelUv : pixelUv, c.gWorldToClipPrev, c.gRectSize);
            float smbParallaxInPixels2 = Pass::ComputeParallaxInPixels(Xprev - c.gCameraDelta.xyz(), c.gOrthoMode == 0.0f ? pixelUv : smbPixelUv, c.gWorldToClip, c.gRectSize);
            float smbParallaxInPixelsMax = max(smbParallaxInPixels1, smbParallaxInPixels2);
            float smbParallaxInPixelsMin = min(smbParallaxInPixels1, smbParallaxInPixels2);

            // disocclusion threshold (:213-234)
            float pixelSize = P.PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, 1.0f, viewZ);
            float frustumSize = P.GetFrustumSize(c.gMinRectDimMulUnproject, c.gOrthoMode, viewZ);
            float disocclusionThresholdMix = 0.0f;
            if (materialID == c.gStrandMaterialID) disocclusionThresholdMix = NRD_GetNormalizedStrandThickness(c.gStrandThickness, pixelSize);
            if (c.gHasDisocclusionThresholdMix) disocclusionThresholdMix = gIn_DisocclusionThresholdMix.load(pixelPos).x;
            float disocclusionThreshold = lerp(c.gDisocclusionThreshold, c.gDisocclusionThresholdAlternate, disocclusionThresholdMix);
            float smallParallax = Math::LinearStep(0.25f, 0.0f, smbParallaxInPixelsMax);
            disocclusionThreshold += 0.05f * smallParallax;

            float3 V = P.GetViewVector(X);
            float NoV = abs(dot(N, V));
            float NoVstrict = lerp(NoV, 1.0f, saturate(smbParallaxInPixelsMax / 30.0f));
            float4 smbDisocclusionThreshold = float4(Pass::GetDisocclusionThreshold(disocclusionThreshold, frustumSize, NoVstrict));
            smbDisocclusionThreshold *= float4(float(dot(smbNavg, Navg) > almostZeroAngle - 0.25f * smallParallax));
            smbDisocclusionThreshold *= Pass::IsInScreenBilinear(smbBilinearFilter.origin, c.gRectSizePrev);
            smbDisocclusionThreshold -= float4(NRD_EPS);

            // plane distance (:236-245)
            float3 Xvprev = Geometry::AffineTransform(c.gWorldToViewPrev, Xprev);
            float3 smbOcclusion0 = step(abs(prevViewZ0 - float3(Xvprev.z)), float3(smbDisocclusionThreshold.x));
            float3 smbOcclusion1 = step(abs(prevViewZ1 - float3(Xvprev.z)), float3(smbDisocclusionThreshold.y));
            float3 smbOcclusion2 = step(abs(prevViewZ2 - float3(Xvprev.z)), float3(smbDisocclusionThreshold.z));
            float3 smbOcclusion3 = step(abs(prevViewZ3 - float3(Xvprev.z)), float3(smbDisocclusionThreshold.w));

            // material ID (:247-269)
            auto wzxyu = [](uint4 g) { return uint4{g.w, g.z, g.x, g.y}; };
            uint4 smbInternalData0 = wzxyu(gPrev_InternalData.gatheru(smbCatromGatherUv, int2(1, 1)));
            uint4 smbInternalData1 = wzxyu(gPrev_InternalData.gatheru(smbCatromGatherUv, int2(3, 1)));
            uint4 smbInternalData2 = wzxyu(gPrev_InternalData.gatheru(smbCatromGatherUv, int2(1, 3)));
            uint4 smbInternalData3 = wzxyu(gPrev_InternalData.gatheru(smbCatromGatherUv, int2(3, 3)));
            auto mid = [](uint p) { return Pass::UnpackInternalData(p).z; };
            float3 smbMaterialID0 = float3(mid(smbInternalData0.y), mid(smbInternalData0.z), mid(smbInternalData0.w));
            float3 smbMaterialID1 = float3(mid(smbInternalData1.x), mid(smbInternalData1.z), mid(smbInternalData1.w));
            float3 smbMaterialID2 = float3(mid(smbInternalData2.x), mid(smbInternalData2.y), mid(smbInternalData2.w));
            float3 smbMaterialID3 = float3(mid(smbInternalData3.x), mid(smbInternalData3.y), mid(smbInternalData3.z));
            float minMaterialID = min(c.gSpecMinMaterial, c.gDiffMinMaterial);
            auto cmp3 = [&](float3 m) {
                return float3(float(max(materialID, minMaterialID) == max(m.x, minMaterialID)), float(max(materialID, minMaterialID) == max(m.y, minMaterialID)),
                              float(max(materialID, minMaterialID) == max(m.z, minMaterialID)));
            };
            smbOcclusion0 *= cmp3(smbMaterialID0);
            smbOcclusion1 *= cmp3(smbMaterialID1);
            smbOcclusion2 *= cmp3(smbMaterialID2);
            smbOcclusion3 *= cmp3(smbMaterialID3);
            uint4 smbInternalData = {smbInternalData0.w, smbInternalData1.z, smbInternalData2.y, smbInternalData3.x};

            // 2x2 occlusion weights (:271-279)
            float4 smbOcclusionWeights = Filtering::GetBilinearCustomWeights(smbBilinearFilter, float4(smbOcclusion0.z, smbOcclusion1.y, smbOcclusion2.y, smbOcclusion3.x));
            bool smbAllowCatRom = dot(smbOcclusion0 + smbOcclusion1 + smbOcclusion2 + smbOcclusion3, float3(1.0f)) > 11.5f && !P.perf;
            float fbits = smbOcclusion0.z * 1.0f;
            fbits += smbOcclusion1.y * 2.0f;
            fbits += smbOcclusion2.y * 4.0f;
            fbits += smbOcclusion3.x * 8.0f;

            // accumulation speed (:281-294)
            float3 id00 = Pass::UnpackInternalData(smbInternalData.x), id10 = Pass::UnpackInternalData(smbInternalData.y);
            float3 id01 = Pass::UnpackInternalData(smbInternalData.z), id11 = Pass::UnpackInternalData(smbInternalData.w);
            float diffAccumSpeed = Filtering::ApplyBilinearCustomWeights(id00.x, id10.x, id01.x, id11.x, smbOcclusionWeights);
            float smbSpecAccumSpeed = Filtering::ApplyBilinearCustomWeights(id00.y, id10.y, id01.y, id11.y, smbOcclusionWeights);

            // footprint quality (:296-305)
            float3 smbVprev = P.GetViewVectorPrev(Xprev, c.gCameraDelta.xyz());
            float NoVprev = abs(dot(N, smbVprev));
            float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f);
            sizeQuality *= sizeQuality;
            sizeQuality = lerp(0.1f, 1.0f, saturate(sizeQuality));
            float smbFootprintQuality = Filtering::ApplyBilinearFilter(smbOcclusion0.z, smbOcclusion1.y, smbOcclusion2.y, smbOcclusion3.x, smbBilinearFilter);
            smbFootprintQuality = Math::Sqrt01(smbFootprintQuality);
            smbFootprintQuality *= sizeQuality;

            uint checkerboard = Sequence::CheckerBoard(pixelPos, c.gFrameIndex);

            float specAccumSpeed = 0.0f, curvature = 0.0f, virtualHistoryAmount = 0.0f;
            if (sg.spec)
            {
                // (:326-352)
                float specHistoryConfidence = smbFootprintQuality;
                if (c.gHasHistoryConfidence) specHistoryConfidence *= gIn_SpecConfidence->load(pixelPos).x;
                smbSpecAccumSpeed *= lerp(specHistoryConfidence, 1.0f, 1.0f / (1.0f + smbSpecAccumSpeed));
                smbSpecAccumSpeed = min(smbSpecAccumSpeed, c.gMaxAccumulatedFrameNum);
                bool specHasData = c.gSpecCheckerboard == 2 || checkerboard == c.gSpecCheckerboard;
                float4 spec = gIn_Spec->load(pixelPos);

                // curvature (:364-447)
                {
                    float2 uvForZeroParallax = c.gOrthoMode == 0.0f ? smbPixelUv : pixelUv;
                    float2 deltaUv = uvForZeroParallax - Geometry::GetScreenUv(c.gWorldToClipPrev, Xprev + c.gCameraDelta.xyz());
                    deltaUv *= c.gRectSize;
                    deltaUv /= float2(max(smbParallaxInPixels1, 1.0f / 256.0f));

                    float3 n10, x10, n01, x01;
                    {
                        float3 xv = Geometry::ReconstructViewPosition(pixelUv + float2(1, 0) * c.gRectSizeInv, c.gFrustum, 1.0f, c.gOrthoMode);
                        float3 xw = Geometry::RotateVector(c.gViewToWorld, xv);
                        float3 v = P.GetViewVector(xw);
                        float3 o = c.gOrthoMode == 0.0f ? float3(0.0f) : xw;
#ifdef ORACLE_REFERENCE_ASSOCIATION
                        x10 = o + v * float3(dot(X - o, N)) / float3(dot(N, v)); // the shader's order: ( v * a ) / b (:379, :389)
#else
                        x10 = o + v * float3(dot(X - o, N) / dot(N, v));         // this restatement and the kernels: v * ( a / b ) -- one rounding apart
#endif
                        n10 = sNormalRoughness(2, 1).xyz();
                    }
                    {
                        float3 xv = Geometry::ReconstructViewPosition(pixelUv + float2(0, 1) * c.gRectSizeInv, c.gFrustum, 1.0f, c.gOrthoMode);
                        float3 xw = Geometry::RotateVector(c.gViewToWorld, xv);
                        float3 v = P.GetViewVector(xw);
                        float3 o = c.gOrthoMode == 0.0f ? float3(0.0f) : xw;
#ifdef ORACLE_REFERENCE_ASSOCIATION
                        x01 = o + v * float3(dot(X - o, N)) / float3(dot(N, v)); // the shader's order: ( v * a ) / b (:379, :389)
#else
                        x01 = o + v * float3(dot(X - o, N) / dot(N, v));         // this restatement and the kernels: v * ( a / b ) -- one rounding apart
#endif
                        n01 = sNormalRoughness(1, 2).xyz();
                    }
                    float2 w = abs(deltaUv) + float2(1.0f / 256.0f);
                    w /= float2(w.x + w.y);
                    float3 xm = x10 * float3(w.x) + x01 * float3(w.y);
                    float3 n = normalize(n10 * float3(w.x) + n01 * float3(w.y));

                    float deltaUvLenFixed = smbParallaxInPixelsMin;
                    deltaUvLenFixed *= 1.0f + c.gFramerateScale * Sequence::Bayer4x4(pixelPos, c.gFrameIndex);
                    float2 motionUvHigh = pixelUv + float2(deltaUvLenFixed) * deltaUv * c.gRectSizeInv;
                    motionUvHigh = (floor(motionUvHigh * c.gRectSize) + float2(0.5f)) * c.gRectSizeInv;
                    if (deltaUvLenFixed > 1.0f && Pass::IsInScreenNearest(motionUvHigh) != 0.0f)
                    {
                        float2 uvScaled = P.ClampUvToViewport(motionUvHigh);
                        float zHigh = P.UnpackViewZ(gIn_ViewZ.sampleNearest(uvScaled).x);
                        float3 xHigh = Geometry::ReconstructViewPosition(motionUvHigh, c.gFrustum, zHigh, c.gOrthoMode);
                        xHigh = Geometry::RotateVector(c.gViewToWorld, xHigh);
                        float3 nHigh = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.sampleNearest(uvScaled)).xyz();
                        float zError = abs(zHigh - viewZ) * rcp(max(zHigh, viewZ));
                        bool cmp = zError < NRD_CURVATURE_Z_THRESHOLD;
                        n = cmp ? nHigh : n;
                        xm = cmp ? xHigh : xm;
                    }
                    float3 edge = xm - X;
                    float edgeLenSq = Math::LengthSquared(edge);
                    curvature = dot(n - N, edge) * Math::PositiveRcp(edgeLenSq);
                }

                // virtual motion (:449-457)
                float3 Xvirtual = Pass::GetXvirtual(hitDistForTracking, curvature, X, Xprev, N, V, roughness);
                float XvirtualLength = length(Xvirtual);
                float2 vmbPixelUv = Geometry::GetScreenUv(c.gWorldToClipPrev, Xvirtual);
                vmbPixelUv = materialID == c.gCameraAttachedReflectionMaterialID ? smbPixelUv : vmbPixelUv;
                float2 vmbDelta = vmbPixelUv - smbPixelUv;
                float vmbPixelsTraveled = length(vmbDelta * c.gRectSize);

                // roughness (:459-470)
                Filtering::Bilinear vmbBilinearFilter = Filtering::GetBilinearFilter(vmbPixelUv, c.gRectSizePrev);
                float2 vmbBilinearGatherUv = (vmbBilinearFilter.origin + float2(1.0f)) * c.gResourceSizeInvPrev;
                float2 relaxedRoughnessWeightParams = Pass::GetRelaxedRoughnessWeightParams(roughness * roughness, c.gRoughnessFraction, REBLUR_ROUGHNESS_SENSITIVITY_IN_TA);
                float4 vmbRoughness = wzxy(gPrev_Normal_Roughness.gather(vmbBilinearGatherUv, 2));
                float4 roughnessWeight;
                for (int i = 0; i < 4; i++)
                    roughnessWeight[i] = Pass::ComputeNonExponentialWeightWithSigma(vmbRoughness[i] * vmbRoughness[i], relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y, roughnessSigma);
                roughnessWeight = lerp(float4(Math::SmoothStep(1.0f, 0.0f, smbParallaxInPixelsMax)), float4(1.0f), roughnessWeight);
                float virtualHistoryRoughnessBasedConfidence = Filtering::ApplyBilinearFilter(roughnessWeight.x, roughnessWeight.y, roughnessWeight.z, roughnessWeight.w, vmbBilinearFilter);

                // normal: parallax (:472-476).  STOCHASTIC_BILINEAR_FILTER = gNearestClamp with a stochastic texel pick
                auto StochasticBilinear = [&](float2 uv, float2 texSize) {
                    Filtering::Bilinear f = Filtering::GetBilinearFilter(uv, texSize);
                    float2 rnd = rng.GetFloat2();
                    f.origin += step(rnd, f.weights);
                    return (f.origin + float2(0.5f)) / texSize;
                };
                float4 vmbNormalAndRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gPrev_Normal_Roughness.sampleNearest(StochasticBilinear(vmbPixelUv, c.gRectSizePrev) * c.gResolutionScalePrev));
                float3 vmbN = Geometry::RotateVector(c.gWorldPrevToWorld, vmbNormalAndRoughness.xyz());
                float Dfactor = ImportanceSampling::GetSpecularDominantFactor(NoV, roughness);
                float virtualHistoryNormalBasedConfidence = 1.0f / (1.0f + 0.5f * Dfactor * saturate(length(N - vmbN) - REBLUR_NORMAL_ULP) * vmbPixelsTraveled);

                smbNavg = smbFootprintQuality == 0.0f ? vmbN : smbNavg;

                // disocclusion: plane distance and roughness (:481-500)
                float4 vmbOcclusion;
                {
                    float4 vmbOcclusionThreshold = float4(disocclusionThreshold * frustumSize);
                    vmbOcclusionThreshold *= float4(lerp(0.25f, 1.0f, NoV));
                    vmbOcclusionThreshold *= float4(float(dot(vmbN, N) > almostZeroAngle));
                    vmbOcclusionThreshold *= float4(float(dot(vmbN, smbNavg) > almostZeroAngle));
                    vmbOcclusionThreshold *= Pass::IsInScreenBilinear(vmbBilinearFilter.origin, c.gRectSizePrev);
                    vmbOcclusionThreshold -= float4(NRD_EPS);

                    float4 g = wzxy(gPrev_ViewZ.gather(vmbBilinearGatherUv, 0));
                    float4 vmbViewZ = float4(P.UnpackViewZ(g.x), P.UnpackViewZ(g.y), P.UnpackViewZ(g.z), P.UnpackViewZ(g.w));
                    float3 vmbVv = Geometry::ReconstructViewPosition(vmbPixelUv, c.gFrustumPrev, 1.0f);
                    float3 vmbV = Geometry::RotateVectorInverse(c.gWorldToViewPrev, vmbVv);
                    float NoXcurr = dot(N, Xprev - c.gCameraDelta.xyz());
                    float4 NoXprev = float4(N.x * vmbV.x + N.y * vmbV.y) * (c.gOrthoMode == 0.0f ? vmbViewZ : float4(c.gOrthoMode)) + float4(N.z * vmbV.z) * vmbViewZ;
                    float4 vmbPlaneDist = abs(NoXprev - float4(NoXcurr));
                    vmbOcclusion = step(vmbPlaneDist, vmbOcclusionThreshold);
                    vmbOcclusion *= step(float4(0.5f), roughnessWeight);
                }

                // material ID (:502-513)
                uint4 vmbInternalData = wzxyu(gPrev_InternalData.gatheru(vmbBilinearGatherUv));
                float3 vmbInternalData00 = Pass::UnpackInternalData(vmbInternalData.x), vmbInternalData10 = Pass::UnpackInternalData(vmbInternalData.y);
                float3 vmbInternalData01 = Pass::UnpackInternalData(vmbInternalData.z), vmbInternalData11 = Pass::UnpackInternalData(vmbInternalData.w);
                float4 vmbMaterialID = float4(vmbInternalData00.z, vmbInternalData10.z, vmbInternalData01.z, vmbInternalData11.z);
                for (int i = 0; i < 4; i++) vmbOcclusion[i] *= float(max(materialID, c.gSpecMinMaterial) == max(vmbMaterialID[i], c.gSpecMinMaterial));

                fbits += vmbOcclusion.x * 16.0f;
                fbits += vmbOcclusion.y * 32.0f;
                fbits += vmbOcclusion.z * 64.0f;
                fbits += vmbOcclusion.w * 128.0f;

                // accumulation speed (:521-531)
                float4 vmbOcclusionWeights = Filtering::GetBilinearCustomWeights(vmbBilinearFilter, vmbOcclusion);
                float vmbSpecAccumSpeed = Filtering::ApplyBilinearCustomWeights(vmbInternalData00.y, vmbInternalData10.y, vmbInternalData01.y, vmbInternalData11.y, vmbOcclusionWeights);
                float vmbFootprintQuality = Filtering::ApplyBilinearFilter(vmbOcclusion.x, vmbOcclusion.y, vmbOcclusion.z, vmbOcclusion.w, vmbBilinearFilter);
                vmbFootprintQuality = Math::Sqrt01(vmbFootprintQuality);
                vmbSpecAccumSpeed *= lerp(vmbFootprintQuality, 1.0f, 1.0f / (1.0f + vmbSpecAccumSpeed));
                bool vmbAllowCatRom = dot(vmbOcclusion, float4(1.0f)) > 3.5f && !P.perf;
                vmbAllowCatRom = vmbAllowCatRom && smbAllowCatRom;

                // (:533-556)
                float curvatureAngleTan = pixelSize * abs(curvature);
                curvatureAngleTan *= max(vmbPixelsTraveled / max(NoV, 0.01f), 1.0f);
                curvatureAngleTan *= 2.0f;
                float curvatureAngle = atan(curvatureAngleTan);
                float percentOfVolume = NRD_MAX_PERCENT_OF_LOBE_VOLUME / (1.0f + vmbSpecAccumSpeed);
                float lobeTanHalfAngle = ImportanceSampling::GetSpecularLobeTanHalfAngle(roughnessModified, percentOfVolume);
                float lobeHalfAngle = atan(lobeTanHalfAngle);
                lobeHalfAngle = max(lobeHalfAngle, NRD_NORMAL_ENCODING_ERROR);
                float normalWeight = Pass::GetEncodingAwareNormalWeight(N, vmbN, lobeHalfAngle, curvatureAngle, REBLUR_NORMAL_ULP);
                normalWeight = lerp(Math::SmoothStep(1.0f, 0.0f, vmbPixelsTraveled), 1.0f, normalWeight);
                virtualHistoryNormalBasedConfidence = min(virtualHistoryNormalBasedConfidence, normalWeight);

                // (:558-561)
                virtualHistoryAmount = Math::SmoothStep(0.05f, 0.95f, Dfactor);
                virtualHistoryAmount *= virtualHistoryNormalBasedConfidence;

                // parallax difference (:563-579)
                float virtualHistoryParallaxBasedConfidence;
                {
                    float hitDistForTrackingPrev = gPrev_SpecHitDistForTracking->sampleLinear(vmbPixelUv * c.gResolutionScalePrev).x;
                    float3 XvirtualPrev = Pass::GetXvirtual(hitDistForTrackingPrev, curvature, X, Xprev, N, V, roughness);
                    float2 vmbPixelUvPrev = Geometry::GetScreenUv(c.gWorldToClipPrev, XvirtualPrev);
                    vmbPixelUvPrev = materialID == c.gCameraAttachedReflectionMaterialID ? smbPixelUv : vmbPixelUvPrev;
                    float pixelSizeAtXvirtual = P.PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, 1.0f, XvirtualLength);
                    float r = (lobeTanHalfAngle + curvatureAngle) * min(hitDistForTracking, hitDistForTrackingPrev) / pixelSizeAtXvirtual;
                    float d = length((vmbPixelUvPrev - vmbPixelUv) * c.gRectSize);
                    r = max(r, 0.1f);
                    virtualHistoryParallaxBasedConfidence = Math::LinearStep(r, 0.0f, d);
                }

                // prev-prev tests (:581-611), 1 iteration
                float stepBetweenTaps = min(vmbPixelsTraveled * c.gFramerateScale, 2.0f) + vmbPixelsTraveled / 1.0f;
                vmbDelta *= float2(Math::Rsqrt(Math::LengthSquared(vmbDelta)));
                vmbDelta /= c.gRectSizePrev;
                relaxedRoughnessWeightParams = Pass::GetRelaxedRoughnessWeightParams(vmbNormalAndRoughness.w * vmbNormalAndRoughness.w, c.gRoughnessFraction, REBLUR_ROUGHNESS_SENSITIVITY_IN_TA);
                for (int i = 1; i <= 1; i++)
                {
                    float2 vmbPixelUvPrev = vmbPixelUv + vmbDelta * float2(float(i) * stepBetweenTaps);
                    float4 vmbNormalAndRoughnessPrev = NRD_FrontEnd_UnpackNormalAndRoughness(gPrev_Normal_Roughness.sampleNearest(StochasticBilinear(vmbPixelUvPrev, c.gRectSizePrev) * c.gResolutionScalePrev));
                    float2 w;
                    w.x = Pass::GetEncodingAwareNormalWeight(vmbNormalAndRoughness.xyz(), vmbNormalAndRoughnessPrev.xyz(), lobeHalfAngle, curvatureAngle * (1.0f + float(i) * stepBetweenTaps), REBLUR_NORMAL_ULP);
                    w.y = Pass::ComputeNonExponentialWeightWithSigma(vmbNormalAndRoughnessPrev.w * vmbNormalAndRoughnessPrev.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y, roughnessSigma);
                    w = lerp(float2(1.0f), w, saturate(stepBetweenTaps));
                    w = Pass::IsInScreenNearest(vmbPixelUvPrev) != 0.0f ? w : float2(1.0f);
                    virtualHistoryNormalBasedConfidence = min(virtualHistoryNormalBasedConfidence, w.x);
                    virtualHistoryRoughnessBasedConfidence = min(virtualHistoryRoughnessBasedConfidence, w.y);
                }

                // (:613-618)
                float virtualHistoryConfidenceForSmbRelaxation = virtualHistoryNormalBasedConfidence * virtualHistoryRoughnessBasedConfidence;
                float virtualHistoryConfidence = virtualHistoryNormalBasedConfidence * virtualHistoryRoughnessBasedConfidence * virtualHistoryParallaxBasedConfidence;
                virtualHistoryAmount *= virtualHistoryRoughnessBasedConfidence;

                // surface history (:620-630)
                float4 smbSpecHistory;
                float smbSpecFastHistory;
                P.BicubicFilter(saturate(smbPixelUv) * c.gRectSizePrev, c.gResourceSizeInvPrev, smbOcclusionWeights, smbAllowCatRom, *gHistory_Spec, smbSpecHistory, gHistory_SpecFast, &smbSpecFastHistory);

                // surface motion confidence (:632-653)
                float surfaceHistoryConfidence = 1.0f;
                {
                    float a = atan(smbParallaxInPixelsMax * pixelSize / length(X));
                    float nonLinearAccumSpeed = 1.0f / (1.0f + smbSpecAccumSpeed);
                    float h = lerp(smbSpecHistory.w, spec.w, nonLinearAccumSpeed) * hitDistNormalization;
                    float tana0 = ImportanceSampling::GetSpecularLobeTanHalfAngle(roughnessModified, NRD_MAX_PERCENT_OF_LOBE_VOLUME);
                    tana0 *= lerp(NoV, 1.0f, roughnessModified);
                    tana0 *= nonLinearAccumSpeed;
                    tana0 /= Pass::GetHitDistFactor(h, frustumSize) + NRD_EPS;
                    float a0 = atan(tana0);
                    a0 = max(a0, NRD_NORMAL_ENCODING_ERROR);
                    float f = Math::LinearStep(a0, 0.0f, a);
                    surfaceHistoryConfidence = Math::Pow01(f, 4.0f);
                }

                // responsive accumulation (:655-667)
                float2 maxResponsiveFrameNum = float2(c.gMaxAccumulatedFrameNum);
                {
                    float responsiveFactor = P.RemapRoughnessToResponsiveFactor(roughness);
                    float smc = Pass::GetSpecMagicCurve(roughnessModified);
                    float2 f;
                    f.x = dot(N, normalize(smbNavg));
                    f.y = dot(N, vmbN);
                    f = float2(lerp(smc, 1.0f, responsiveFactor)) * Math::Pow01(f, lerp(32.0f, 1.0f, smc) * (1.0f - responsiveFactor));
                    maxResponsiveFrameNum = max(float2(c.gMaxAccumulatedFrameNum) * f, float2(c.gHistoryFixFrameNum));
                }

                // (:669-686)
                float smbMaxFrameNum = c.gMaxAccumulatedFrameNum;
                smbMaxFrameNum *= surfaceHistoryConfidence;
                smbMaxFrameNum = min(smbMaxFrameNum, maxResponsiveFrameNum.x);
                float smbBoostedMaxFrameNum = max(smbMaxFrameNum, c.gHistoryFixFrameNum * (1.0f - virtualHistoryConfidenceForSmbRelaxation));
                float smbSpecAccumSpeedBoosted = min(smbSpecAccumSpeed, smbBoostedMaxFrameNum);
                float vmbMaxFrameNum = c.gMaxAccumulatedFrameNum;
                vmbMaxFrameNum *= virtualHistoryConfidence;
                vmbMaxFrameNum = min(vmbMaxFrameNum, maxResponsiveFrameNum.y);
                smbSpecAccumSpeed = min(smbSpecAccumSpeed, smbMaxFrameNum);
                vmbSpecAccumSpeed = min(vmbSpecAccumSpeed, vmbMaxFrameNum);

                // fallback to smb (:688-703)
                float magic = vmbSpecAccumSpeed > smbSpecAccumSpeed ? 8.0f : 0.5f;
                virtualHistoryAmount *= 1.0f + (vmbSpecAccumSpeed - smbSpecAccumSpeed) / (magic * max(vmbSpecAccumSpeed, smbSpecAccumSpeed) + 1.0f);
                virtualHistoryAmount = saturate(virtualHistoryAmount);

                // virtual history (:709-721)
                float4 vmbSpecHistory;
                float vmbSpecFastHistory;
                P.BicubicFilter(saturate(vmbPixelUv) * c.gRectSizePrev, c.gResourceSizeInvPrev, vmbOcclusionWeights, vmbAllowCatRom, *gHistory_Spec, vmbSpecHistory, gHistory_SpecFast, &vmbSpecFastHistory);

                smbSpecHistory = Pass::ClampNegativeToZero(smbSpecHistory);
                vmbSpecHistory = Pass::ClampNegativeToZero(vmbSpecHistory);

                // accumulation (:727-754)
                float smbSpecNonLinearAccumSpeed = 1.0f / (1.0f + smbSpecAccumSpeed);
                float vmbSpecNonLinearAccumSpeed = 1.0f / (1.0f + vmbSpecAccumSpeed);
                if (!specHasData)
                {
                    smbSpecNonLinearAccumSpeed *= lerp(1.0f - c.gCheckerboardResolveAccumSpeed, 1.0f, smbSpecNonLinearAccumSpeed);
                    vmbSpecNonLinearAccumSpeed *= lerp(1.0f - c.gCheckerboardResolveAccumSpeed, 1.0f, vmbSpecNonLinearAccumSpeed);
                }
                float4 smbSpec = P.MixHistoryAndCurrent(smbSpecHistory, spec, smbSpecNonLinearAccumSpeed, roughnessModified);
                float4 vmbSpec = P.MixHistoryAndCurrent(vmbSpecHistory, spec, vmbSpecNonLinearAccumSpeed, roughnessModified);
                float4 specResult = lerp(smbSpec, vmbSpec, virtualHistoryAmount);
                specAccumSpeed = lerp(smbSpecAccumSpeedBoosted, vmbSpecAccumSpeed, virtualHistoryAmount);
                float4 specHistory = lerp(smbSpecHistory, vmbSpecHistory, virtualHistoryAmount);

                // firefly suppressor (:756-771)
                float specMaxRelativeIntensity = c.gFireflySuppressorMinRelativeScale + REBLUR_FIREFLY_SUPPRESSOR_MAX_RELATIVE_INTENSITY / (specAccumSpeed + 1.0f);
                float specAntifireflyFactor = specAccumSpeed * c.gMaxBlurRadius * REBLUR_FIREFLY_SUPPRESSOR_RADIUS_SCALE;
                specAntifireflyFactor /= 1.0f + specAntifireflyFactor;
                float specLumaResult = Pass::GetLuma(specResult);
                float specLumaClamped = min(specLumaResult, Pass::GetLuma(specHistory) * specMaxRelativeIntensity);
                specLumaClamped = lerp(specLumaResult, specLumaClamped, specAntifireflyFactor);
                specResult = Pass::ChangeLuma(specResult, specLumaClamped);
                gOut_Spec->store(pixelPos, specResult);

                // fast history (:779-794)
                float smbSpecFastNonLinearAccumSpeed = P.GetNonLinearAccumSpeed(smbSpecAccumSpeed, c.gMaxFastAccumulatedFrameNum, surfaceHistoryConfidence, specHasData);
                float vmbSpecFastNonLinearAccumSpeed = P.GetNonLinearAccumSpeed(vmbSpecAccumSpeed, c.gMaxFastAccumulatedFrameNum, virtualHistoryConfidence, specHasData);
                float smbSpecFast = lerp(smbSpecFastHistory, Pass::GetLuma(spec), smbSpecFastNonLinearAccumSpeed);
                float vmbSpecFast = lerp(vmbSpecFastHistory, Pass::GetLuma(spec), vmbSpecFastNonLinearAccumSpeed);
                float specFastResult = lerp(smbSpecFast, vmbSpecFast, virtualHistoryAmount);
                float specFastClamped = min(specFastResult, Pass::GetLuma(specHistory) * specMaxRelativeIntensity * REBLUR_FIREFLY_SUPPRESSOR_FAST_RELATIVE_INTENSITY);
                specFastResult = lerp(specFastResult, specFastClamped, specAntifireflyFactor);
                gOut_SpecFast->store(pixelPos, specFastResult);
            }

            gOut_Data2.storeu(pixelPos, Pass::PackData2(fbits, curvature, virtualHistoryAmount));

            if (sg.diff)
            {
                // (:826-927)
                float diffHistoryConfidence = smbFootprintQuality;
                if (c.gHasHistoryConfidence) diffHistoryConfidence *= gIn_DiffConfidence->load(pixelPos).x;
                diffAccumSpeed *= lerp(diffHistoryConfidence, 1.0f, 1.0f / (1.0f + diffAccumSpeed));
                diffAccumSpeed = min(diffAccumSpeed, c.gMaxAccumulatedFrameNum);
                bool diffHasData = c.gDiffCheckerboard == 2 || checkerboard == c.gDiffCheckerboard;
                float4 diff = gIn_Diff->load(pixelPos);

                float4 smbDiffHistory;
                float smbDiffFastHistory;
                P.BicubicFilter(saturate(smbPixelUv) * c.gRectSizePrev, c.gResourceSizeInvPrev, smbOcclusionWeights, smbAllowCatRom, *gHistory_Diff, smbDiffHistory, gHistory_DiffFast, &smbDiffFastHistory);
                smbDiffHistory = Pass::ClampNegativeToZero(smbDiffHistory);

                float diffNonLinearAccumSpeed = 1.0f / (1.0f + diffAccumSpeed);
                if (!diffHasData) diffNonLinearAccumSpeed *= lerp(1.0f - c.gCheckerboardResolveAccumSpeed, 1.0f, diffNonLinearAccumSpeed);
                float4 diffResult = P.MixHistoryAndCurrent(smbDiffHistory, diff, diffNonLinearAccumSpeed);

                float diffMaxRelativeIntensity = c.gFireflySuppressorMinRelativeScale + REBLUR_FIREFLY_SUPPRESSOR_MAX_RELATIVE_INTENSITY / (diffAccumSpeed + 1.0f);
                float diffAntifireflyFactor = diffAccumSpeed * c.gMaxBlurRadius * REBLUR_FIREFLY_SUPPRESSOR_RADIUS_SCALE;
                diffAntifireflyFactor /= 1.0f + diffAntifireflyFactor;
                float diffLumaResult = Pass::GetLuma(diffResult);
                float diffLumaClamped = min(diffLumaResult, Pass::GetLuma(smbDiffHistory) * diffMaxRelativeIntensity);
                diffLumaClamped = lerp(diffLumaResult, diffLumaClamped, diffAntifireflyFactor);
                diffResult = Pass::ChangeLuma(diffResult, diffLumaClamped);
                gOut_Diff->store(pixelPos, diffResult);

                float diffFastAccumSpeed = min(diffAccumSpeed, c.gMaxFastAccumulatedFrameNum);
                float diffFastNonLinearAccumSpeed = 1.0f / (1.0f + diffFastAccumSpeed);
                if (!diffHasData) diffFastNonLinearAccumSpeed *= lerp(1.0f - c.gCheckerboardResolveAccumSpeed, 1.0f, diffFastNonLinearAccumSpeed);
                float diffFastResult = lerp(smbDiffFastHistory, Pass::GetLuma(diff), diffFastNonLinearAccumSpeed);
                float diffFastClamped = min(diffFastResult, Pass::GetLuma(smbDiffHistory) * diffMaxRelativeIntensity * REBLUR_FIREFLY_SUPPRESSOR_FAST_RELATIVE_INTENSITY);
                diffFastResult = lerp(diffFastResult, diffFastClamped, diffAntifireflyFactor);
                gOut_DiffFast->store(pixelPos, diffFastResult);
            }
            else
                diffAccumSpeed = 0.0f;

            // PackData1 (REBLUR_Common.hlsli:35-47)
            float2 r = float2(saturate(diffAccumSpeed / REBLUR_MAX_ACCUM_FRAME_NUM), saturate(specAccumSpeed / REBLUR_MAX_ACCUM_FRAME_NUM));
            if (!sg.diff) r.x = r.y;
            gOut_Data1.store(pixelPos, float4(r.x, r.y, 0, 0));
        }
}

// REBLUR_HistoryFix.hlsli
void HistoryFix(const Pass& P, Signals sg, Tex* t, int W, int H)
{
    const CB& c = P.c;
    int k = 0;
    const Tex& gIn_Tiles = t[k++];
    const Tex& gIn_Normal_Roughness = t[k++];
    const Tex& gIn_Data1 = t[k++];
    const Tex& gIn_ViewZ = t[k++];
    const Tex* gIn_Diff = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_Spec = sg.spec ? &t[k++] : nullptr;
    const Tex* gIn_DiffFast = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_SpecFast = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_Diff = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_Spec = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_DiffFast = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_SpecFast = sg.spec ? &t[k++] : nullptr;
    const int2 rectMax = int2(c.gRectSizeMinusOne[0], c.gRectSizeMinusOne[1]);
    const bool both = sg.diff && sg.spec;
    auto unpackData1 = [&](float4 d) { return both ? float2(d.x, d.y) * float2(REBLUR_MAX_ACCUM_FRAME_NUM) : float2(d.x, d.x) * float2(REBLUR_MAX_ACCUM_FRAME_NUM); };

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x > rectMax.x || y > rectMax.y) continue;
            float viewZ = P.UnpackViewZ(gIn_ViewZ.load(x, y).x);
            if (viewZ > c.gDenoisingRange) continue;

            float materialID;
            float4 normalAndRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos), materialID);
            float3 N = normalAndRoughness.xyz();
            float roughness = normalAndRoughness.w;
            float frustumSize = P.GetFrustumSize(c.gMinRectDimMulUnproject, c.gOrthoMode, viewZ);
            float2 pixelUv = (tofloat(pixelPos) + float2(0.5f)) * c.gRectSizeInv;
            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, c.gFrustum, viewZ, c.gOrthoMode);
            float3 Nv = Geometry::RotateVectorInverse(c.gViewToWorld, N);
            float2 frameNum = unpackData1(gIn_Data1.load(pixelPos));
            float2 stride = float2(c.gHistoryFixBasePixelStride) / (float2(2.0f) + frameNum);

            for (int pass = 0; pass < 2; pass++)
            {
                const bool isSpec = pass == 1;
                if (isSpec ? !sg.spec : !sg.diff) continue;
                const Tex& gIn_Sig = isSpec ? *gIn_Spec : *gIn_Diff;
                const Tex& gIn_Fast = isSpec ? *gIn_SpecFast : *gIn_DiffFast;
                Tex& gOut_Sig = isSpec ? *gOut_Spec : *gOut_Diff;
                Tex& gOut_Fast = isSpec ? *gOut_SpecFast : *gOut_DiffFast;
                const float fn = isSpec ? frameNum.y : frameNum.x;
                const float minMaterial = isSpec ? c.gSpecMinMaterial : c.gDiffMinMaterial;

                float4 sig = gIn_Sig.load(pixelPos);
                float smc = Pass::GetSpecMagicCurve(roughness);
                float sigStride = (isSpec ? stride.y : stride.x) * float(fn < c.gHistoryFixFrameNum);
                if (isSpec) sigStride *= lerp(0.5f, 1.0f, smc);
                sigStride = floor(sigStride);

                if (sigStride != 0.0f)
                {
                    int stridei = int(sigStride + 0.5f);
                    float nonLinearAccumSpeed = 1.0f / (1.0f + fn);
                    float normalWeightParam = Pass::GetNormalWeightParam(nonLinearAccumSpeed, c.gLobeAngleFraction, isSpec ? roughness : 1.0f);
                    float2 geometryWeightParams = Pass::GetGeometryWeightParams(c.gPlaneDistSensitivity, frustumSize, Xv, Nv);
                    float2 relaxedRoughnessWeightParams = Pass::GetRelaxedRoughnessWeightParams(roughness * roughness, sqrt(c.gRoughnessFraction));
                    float hitDistScale = _REBLUR_GetHitDistanceNormalization(viewZ, c.gHitDistParams, isSpec ? roughness : 1.0f);
                    float hitDist = sig.w * hitDistScale;
                    float hitDistFactor = Pass::GetHitDistFactor(hitDist, frustumSize);
                    float2 hitDistanceWeightParams = Pass::GetHitDistanceWeightParams(hitDistFactor, nonLinearAccumSpeed, isSpec ? roughness : 1.0f);
                    float sum = 1.0f + fn;
                    if (P.perf) sum = 1.0f + 1.0f / (1.0f + c.gMaxAccumulatedFrameNum) - nonLinearAccumSpeed; // :88-90
                    sig *= float4(sum);

                    for (int j = -2; j <= 2; j++)
                        for (int i = -2; i <= 2; i++)
                        {
                            if (i == 0 && j == 0) continue;
                            if (std::abs(i) + std::abs(j) == 4) continue;
                            float2 uv = pixelUv + float2(float(i), float(j)) * float2(sigStride) * c.gRectSizeInv;
                            int2 pos = clamp(pixelPos + int2(i, j) * stridei, int2(0), rectMax);
                            float zs = P.UnpackViewZ(gIn_ViewZ.load(pos).x);
                            float materialIDs;
                            float4 Ns = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pos), materialIDs);
                            float angle = Math::AcosApprox(dot(Ns.xyz(), N));
                            float3 Xvs = Geometry::ReconstructViewPosition(uv, c.gFrustum, zs, c.gOrthoMode);
                            float w = Pass::IsInScreenNearest(uv);
                            w *= Pass::ComputeWeight(dot(Nv, Xvs), geometryWeightParams.x, geometryWeightParams.y);
                            w *= float(max(materialID, minMaterial) == max(materialIDs, minMaterial));
                            w *= Pass::ComputeExponentialWeight(angle, normalWeightParam, 0.0f);
                            if (isSpec) w *= Pass::ComputeExponentialWeight(Ns.w * Ns.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);
                            if (!P.perf) // :139-141
                            {
                                float2 d1 = unpackData1(gIn_Data1.load(pos));
                                w *= 1.0f + (isSpec ? d1.y : d1.x);
                            }
                            float4 sv = gIn_Sig.load(pos);
                            sv = w == 0.0f ? float4(0.0f) : sv;
                            float hs = sv.w * hitDistScale;
                            float hsFactor = Pass::GetHitDistFactor(hs, frustumSize);
                            w *= Pass::ComputeExponentialWeight(hsFactor, hitDistanceWeightParams.x, hitDistanceWeightParams.y);
                            if (isSpec)
                            {
                                float d = abs(hitDist - hs) / (max(hitDist, hs) + 0.001f);
                                float b = Math::LinearStep(0.03f, 0.05f, roughness);
                                w *= Math::SmoothStep(0.2f + b, 0.05f + b, d);
                            }
                            sum += w;
                            sig += sv * float4(w);
                        }
                    sum = Math::PositiveRcp(sum);
                    sig *= float4(sum);
                }

                // local variance from the fast history ("shared memory" = clamped loads, Preload :13-25)
                auto sLuma = [&](int i, int j) { return gIn_Fast.load(clamp(int2(x + i - 2, y + j - 2), int2(0), rectMax)).x; };
                float center = sLuma(2, 2);
                float m1 = center;
                float m2 = m1 * m1;
                float f = saturate(fn / (c.gHistoryFixFrameNum + NRD_EPS));
                if (isSpec) f = lerp(1.0f, f, smc);
                center = lerp(Pass::GetLuma(sig), center, f);
                gOut_Fast.store(pixelPos, center);

                for (int j = 0; j <= 4; j++)
                    for (int i = 0; i <= 4; i++)
                    {
                        if (i == 2 && j == 2) continue;
                        float d = sLuma(i, j);
                        m1 += d;
                        m2 += d * d;
                    }
                float luma = Pass::GetLuma(sig);

                if (c.gAntiFirefly != 0.0f)
                {
                    float am1 = 0, am2 = 0;
                    const int REBLUR_ANTI_FIREFLY_FILTER_RADIUS = P.perf ? 3 : 4; // REBLUR_Config.hlsli:87, :236-237
                    for (int j = -REBLUR_ANTI_FIREFLY_FILTER_RADIUS; j <= REBLUR_ANTI_FIREFLY_FILTER_RADIUS; j++)
                        for (int i = -REBLUR_ANTI_FIREFLY_FILTER_RADIUS; i <= REBLUR_ANTI_FIREFLY_FILTER_RADIUS; i++)
                        {
                            if (std::abs(i) <= 1 && std::abs(j) <= 1) continue;
                            float d = gIn_Fast.load(clamp(pixelPos + int2(i, j), int2(0), rectMax)).x;
                            am1 += d;
                            am2 += d * d;
                        }
                    float invNorm = 1.0f / ((REBLUR_ANTI_FIREFLY_FILTER_RADIUS * 2 + 1) * (REBLUR_ANTI_FIREFLY_FILTER_RADIUS * 2 + 1) - 3 * 3);
                    am1 *= invNorm;
                    am2 *= invNorm;
                    float sigma = Pass::GetStdDev(am1, am2) * REBLUR_ANTI_FIREFLY_SIGMA_SCALE;
                    luma = clamp(luma, am1 - sigma, am1 + sigma);
                }

                m1 /= 25.0f;
                m2 /= 25.0f;
                float sigma = Pass::GetStdDev(m1, m2) * REBLUR_COLOR_CLAMPING_SIGMA_SCALE;
                float lumaClamped = clamp(luma, m1 - sigma, m1 + sigma);
                luma = lerp(lumaClamped, luma, 1.0f / (1.0f + float(c.gMaxFastAccumulatedFrameNum < c.gMaxAccumulatedFrameNum) * fn * 2.0f));
                sig = Pass::ChangeLuma(sig, luma);
                gOut_Sig.store(pixelPos, sig);
            }
        }
}

// REBLUR_TemporalStabilization.hlsli
void TemporalStabilization(const Pass& P, Signals sg, Tex* t, int W, int H)
{
    const CB& c = P.c;
    int k = 0;
    const Tex& gIn_Tiles = t[k++];
    const Tex& gIn_Normal_Roughness = t[k++];
    const Tex* gIn_BaseColor_Metalness = sg.spec ? &t[k++] : nullptr; // read only when the MV patch is enabled (isBaseColorMetalnessAvailable, Reblur.cpp:359)
    const Tex& gIn_ViewZ = t[k++];
    const Tex& gIn_Data1 = t[k++];
    const Tex& gIn_Data2 = t[k++];
    const Tex* gIn_Diff = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_Spec = sg.spec ? &t[k++] : nullptr;
    const Tex* gHistory_DiffLumaStabilized = sg.diff ? &t[k++] : nullptr;
    const Tex* gHistory_SpecLumaStabilized = sg.spec ? &t[k++] : nullptr;
    const Tex* gIn_SpecHitDistForTracking = sg.spec ? &t[k++] : nullptr;
    Tex& gInOut_Mv = t[k++];
    Tex& gOut_InternalData = t[k++];
    Tex* gOut_Diff = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_Spec = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_DiffLumaStabilized = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_SpecLumaStabilized = sg.spec ? &t[k++] : nullptr;
    const int2 rectMax = int2(c.gRectSizeMinusOne[0], c.gRectSizeMinusOne[1]);
    const bool both = sg.diff && sg.spec;

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x > rectMax.x || y > rectMax.y) continue;
            float viewZ = P.UnpackViewZ(gIn_ViewZ.load(x, y).x);
            if (viewZ > c.gDenoisingRange) continue;

            float2 pixelUv = (tofloat(pixelPos) + float2(0.5f)) * c.gRectSizeInv;
            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, c.gFrustum, viewZ, c.gOrthoMode);
            float3 X = Geometry::RotateVector(c.gViewToWorld, Xv);

            float4 inMv = gInOut_Mv.load(pixelPos);
            float3 mv = inMv.xyz() * c.gMvScale.xyz();
            float3 Xprev = X;
            float2 smbPixelUv = pixelUv + mv.xy();
            if (c.gMvScale.w == 0.0f)
            {
                if (c.gMvScale.z == 0.0f) mv.z = Geometry::AffineTransform(c.gWorldToViewPrev, X).z - viewZ;
                float viewZprev = viewZ + mv.z;
                float3 Xvprevlocal = Geometry::ReconstructViewPosition(smbPixelUv, c.gFrustumPrev, viewZprev, c.gOrthoMode);
                Xprev = Geometry::RotateVectorInverse(c.gWorldToViewPrev, Xvprevlocal) + c.gCameraDelta.xyz();
            }
            else
            {
                Xprev += mv;
                smbPixelUv = Geometry::GetScreenUv(c.gWorldToClipPrev, Xprev);
            }

            float materialID;
            float4 normalAndRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos), materialID);
            float3 N = normalAndRoughness.xyz();
            float roughness = normalAndRoughness.w;

            uint bits;
            float4 d1 = gIn_Data1.load(pixelPos);
            float2 data1 = both ? float2(d1.x, d1.y) * float2(REBLUR_MAX_ACCUM_FRAME_NUM) : float2(d1.x, d1.x) * float2(REBLUR_MAX_ACCUM_FRAME_NUM);
            float2 data2 = Pass::UnpackData2(gIn_Data2.loadu(pixelPos), bits);

            Filtering::Bilinear smbBilinearFilter = Filtering::GetBilinearFilter(smbPixelUv, c.gRectSizePrev);
            float4 smbOcclusion = float4(float((bits & 1) != 0), float((bits & 2) != 0), float((bits & 4) != 0), float((bits & 8) != 0));
            float4 smbOcclusionWeights = Filtering::GetBilinearCustomWeights(smbBilinearFilter, smbOcclusion);
            bool smbAllowCatRom = dot(smbOcclusion, float4(1.0f)) > 3.5f && !P.perf;
            float smbFootprintQuality = Filtering::ApplyBilinearFilter(smbOcclusion.x, smbOcclusion.y, smbOcclusion.z, smbOcclusion.w, smbBilinearFilter);
            smbFootprintQuality = Math::Sqrt01(smbFootprintQuality);

            // 3x3 luma statistics ("shared memory" = clamped loads of the signal's luma, Preload :13-25)
            auto stats = [&](const Tex& tex, float& luma, float& m1, float& m2, float& mn, float& mx) {
                auto sLuma = [&](int i, int j) { return Pass::GetLuma(tex.load(clamp(int2(x + i - 1, y + j - 1), int2(0), rectMax))); };
                luma = sLuma(1, 1);
                m1 = luma;
                m2 = luma * luma;
                mn = NRD_INF;
                mx = -NRD_INF;
                for (int j = 0; j <= 2; j++)
                    for (int i = 0; i <= 2; i++)
                    {
                        if (i == 1 && j == 1) continue;
                        float d = sLuma(i, j);
                        m1 += d;
                        m2 += d * d;
                        mn = min(mn, d);
                        mx = max(mx, d);
                    }
                m1 /= 9.0f;
                m2 /= 9.0f;
            };

            if (sg.diff)
            {
                float diffLuma, diffLumaM1, diffLumaM2, diffMin, diffMax;
                stats(*gIn_Diff, diffLuma, diffLumaM1, diffLumaM2, diffMin, diffMax);
                float diffLumaSigma = Pass::GetStdDev(diffLumaM1, diffLumaM2);
                if (c.gMaxBlurRadius != 0.0f && !P.perf) diffLuma = clamp(diffLuma, diffMin, diffMax); // RCRS: not in performance mode (:131-135)

                float4 h4;
                P.BicubicFilter(saturate(smbPixelUv) * c.gRectSizePrev, c.gResourceSizeInvPrev, smbOcclusionWeights, smbAllowCatRom, *gHistory_DiffLumaStabilized, h4, nullptr, nullptr);
                float smbDiffLumaHistory = max(h4.x, 0.0f);

                float diffAntilag = P.ComputeAntilag(smbDiffLumaHistory, diffLumaM1, diffLumaSigma, smbFootprintQuality * data1.x);
                float2 tap = P.GetTemporalAccumulationParams(smbFootprintQuality, data1.x);
                float diffHistoryWeight = tap.x;
                diffHistoryWeight *= diffAntilag;
                diffHistoryWeight *= float(pixelUv.x >= c.gSplitScreen);
                diffHistoryWeight *= float(smbPixelUv.x >= c.gSplitScreenPrev);
                smbDiffLumaHistory = Color::Clamp(diffLumaM1, diffLumaSigma * tap.y, smbDiffLumaHistory);
                float diffLumaStabilized = lerp(diffLuma, smbDiffLumaHistory, min(diffHistoryWeight, c.gStabilizationStrength));

                float4 diff = gIn_Diff->load(pixelPos);
                diff = Pass::ChangeLuma(diff, diffLumaStabilized);
                gOut_Diff->store(pixelPos, diff);
                gOut_DiffLumaStabilized->store(pixelPos, diffLumaStabilized);

                data1.x += 1.0f;
                float diffMinAccumSpeed = min(data1.x, c.gHistoryFixFrameNum) * 1.0f;
                data1.x = lerp(diffMinAccumSpeed, data1.x, diffAntilag);
            }

            if (sg.spec)
            {
                float specLuma, specLumaM1, specLumaM2, specMin, specMax;
                stats(*gIn_Spec, specLuma, specLumaM1, specLumaM2, specMin, specMax);
                float specLumaSigma = Pass::GetStdDev(specLumaM1, specLumaM2);
                if (c.gMaxBlurRadius != 0.0f && !P.perf) specLuma = clamp(specLuma, specMin, specMax);

                float virtualHistoryAmount = data2.x;
                float curvature = data2.y;
                float4 spec = gIn_Spec->load(pixelPos);
                float hitDistForTracking = spec.w * _REBLUR_GetHitDistanceNormalization(viewZ, c.gHitDistParams, roughness);
                if (c.gSpecPrepassBlurRadius != 0.0f) hitDistForTracking = min(hitDistForTracking, gIn_SpecHitDistForTracking->load(pixelPos).x);

                float3 V = P.GetViewVector(X);
                float3 Xvirtual = Pass::GetXvirtual(hitDistForTracking, curvature, X, Xprev, N, V, roughness);
                float2 vmbPixelUv = Geometry::GetScreenUv(c.gWorldToClipPrev, Xvirtual);
                vmbPixelUv = materialID == c.gCameraAttachedReflectionMaterialID ? pixelUv : vmbPixelUv;
                // modify MVs if requested (:250-285): specular-dominant pixels get the motion of their reflection
                if (c.gSpecProbabilityThresholdsForMvModification.x < 1.0f)
                {
                    float NoV = abs(dot(N, V));
                    float4 baseColorMetalness = gIn_BaseColor_Metalness->load(pixelPos);
                    float3 albedo, Rf0;
                    BRDF::ConvertBaseColorMetalnessToAlbedoRf0(baseColorMetalness.xyz(), baseColorMetalness.w, albedo, Rf0);
                    float3 Fenv = BRDF::EnvironmentTerm_Rtg(Rf0, NoV, roughness);
                    float lumSpec = Color::Luminance(Fenv);
                    float lumDiff = Color::Luminance(albedo * (float3(1.0f) - Fenv));
                    float specProb = lumSpec / (lumDiff + lumSpec + NRD_EPS);
                    float f = Math::SmoothStep(c.gSpecProbabilityThresholdsForMvModification.x, c.gSpecProbabilityThresholdsForMvModification.y, specProb);
                    f *= 1.0f - Pass::GetSpecMagicCurve(roughness);
                    f *= 1.0f - Math::Sqrt01(abs(curvature));
                    if (f != 0.0f)
                    {
                        float3 specMv = Xvirtual - X;
                        if (c.gMvScale.w == 0.0f)
                        {
                            specMv.x = vmbPixelUv.x - pixelUv.x;
                            specMv.y = vmbPixelUv.y - pixelUv.y;
                            specMv.z = Geometry::AffineTransform(c.gWorldToViewPrev, Xvirtual).z - viewZ;
                        }
                        float3 newMv;
                        newMv.x = specMv.x / c.gMvScale.x;
                        newMv.y = specMv.y / c.gMvScale.y;
                        newMv.z = c.gMvScale.z == 0.0f ? inMv.z : specMv.z / c.gMvScale.z;
                        float3 patched = lerp(inMv.xyz(), newMv, f);
                        gInOut_Mv.store(pixelPos, float4(patched, inMv.w));
                    }
                }

                float4 h4;
                P.BicubicFilter(saturate(smbPixelUv) * c.gRectSizePrev, c.gResourceSizeInvPrev, smbOcclusionWeights, smbAllowCatRom, *gHistory_SpecLumaStabilized, h4, nullptr, nullptr);
                float smbSpecLumaHistory = h4.x;

                Filtering::Bilinear vmbBilinearFilter = Filtering::GetBilinearFilter(vmbPixelUv, c.gRectSizePrev);
                float4 vmbOcclusion = float4(float((bits & 16) != 0), float((bits & 32) != 0), float((bits & 64) != 0), float((bits & 128) != 0));
                float4 vmbOcclusionWeights = Filtering::GetBilinearCustomWeights(vmbBilinearFilter, vmbOcclusion);
                bool vmbAllowCatRom = dot(vmbOcclusion, float4(1.0f)) > 3.5f && !P.perf;
                float vmbFootprintQuality = Filtering::ApplyBilinearFilter(vmbOcclusion.x, vmbOcclusion.y, vmbOcclusion.z, vmbOcclusion.w, vmbBilinearFilter);
                vmbFootprintQuality = Math::Sqrt01(vmbFootprintQuality);

                P.BicubicFilter(saturate(vmbPixelUv) * c.gRectSizePrev, c.gResourceSizeInvPrev, vmbOcclusionWeights, vmbAllowCatRom, *gHistory_SpecLumaStabilized, h4, nullptr, nullptr);
                float vmbSpecLumaHistory = h4.x;

                smbSpecLumaHistory = max(smbSpecLumaHistory, 0.0f);
                vmbSpecLumaHistory = max(vmbSpecLumaHistory, 0.0f);
                float specLumaHistory = lerp(smbSpecLumaHistory, vmbSpecLumaHistory, virtualHistoryAmount);

                float footprintQuality = lerp(smbFootprintQuality, vmbFootprintQuality, virtualHistoryAmount);
                float specAntilag = P.ComputeAntilag(specLumaHistory, specLumaM1, specLumaSigma, footprintQuality * data1.y);
                float2 tap = P.GetTemporalAccumulationParams(footprintQuality, data1.y);
                float specHistoryWeight = tap.x;
                specHistoryWeight *= specAntilag;
                specHistoryWeight *= float(pixelUv.x >= c.gSplitScreen);
                specHistoryWeight *= virtualHistoryAmount != 1.0f ? float(smbPixelUv.x >= c.gSplitScreenPrev) : 1.0f;
                specHistoryWeight *= virtualHistoryAmount != 0.0f ? float(vmbPixelUv.x >= c.gSplitScreenPrev) : 1.0f;

                float responsiveFactor = P.RemapRoughnessToResponsiveFactor(roughness);
                float smc = Pass::GetSpecMagicCurve(roughness);
                float acceleration = lerp(smc, 1.0f, 0.5f + responsiveFactor * 0.5f);
                specHistoryWeight *= materialID == c.gStrandMaterialID ? 0.5f : acceleration;

                specLumaHistory = Color::Clamp(specLumaM1, specLumaSigma * tap.y, specLumaHistory);
                float specLumaStabilized = lerp(specLuma, specLumaHistory, min(specHistoryWeight, c.gStabilizationStrength));
                spec = Pass::ChangeLuma(spec, specLumaStabilized);
                gOut_Spec->store(pixelPos, spec);
                gOut_SpecLumaStabilized->store(pixelPos, specLumaStabilized);

                data1.y += 1.0f;
                float specMinAccumSpeed = min(data1.y, c.gHistoryFixFrameNum) * 1.0f;
                data1.y = lerp(specMinAccumSpeed, data1.y, specAntilag);
            }

            gOut_InternalData.storeu(pixelPos, Pass::PackInternalData(data1.x, data1.y, materialID));
        }
}
// REBLUR_SplitScreen.hlsli:11-47: the part of the screen left of gSplitScreen shows the noisy input (a checkerboarded input is
// stretched: pixel x shows packed column x >> 1)
void SplitScreen(const Pass& P, Signals sg, Tex* t, int W, int H)
{
    const CB& c = P.c;
    int k = 0;
    const Tex& gIn_ViewZ = t[k++];
    const Tex* gIn_Diff = sg.diff ? &t[k++] : nullptr;
    const Tex* gIn_Spec = sg.spec ? &t[k++] : nullptr;
    Tex* gOut_Diff = sg.diff ? &t[k++] : nullptr;
    Tex* gOut_Spec = sg.spec ? &t[k++] : nullptr;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            const int2 pixelPos(x, y);
            float2 pixelUv = (float2(float(x), float(y)) + float2(0.5f)) * c.gRectSizeInv;
            if (pixelUv.x > c.gSplitScreen || x > c.gRectSizeMinusOne[0] || y > c.gRectSizeMinusOne[1]) continue;
            float viewZ = P.UnpackViewZ(gIn_ViewZ.load(pixelPos).x);
            float keep = float(viewZ < c.gDenoisingRange);
            if (sg.diff) gOut_Diff->store(pixelPos, gIn_Diff->load(x >> (c.gDiffCheckerboard != 2 ? 1 : 0), y) * float4(keep));
            if (sg.spec) gOut_Spec->store(pixelPos, gIn_Spec->load(x >> (c.gSpecCheckerboard != 2 ? 1 : 0), y) * float4(keep));
        }
}
} // namespace

int reblur_dispatch_impl(const char* shaderName, const void* constants, int constantsSize, hlsl::Tex* tex, int texNum, int gridW, int gridH);
} // namespace hlsl

// shaderName examples: "REBLUR_ClassifyTiles.cs", "REBLUR_DiffuseSpecular_Blur.cs", "REBLUR_Diffuse_PostBlur_NoTemporalStabilization.cs"
int oracle_reblur_dispatch(const char* shaderName, const void* constants, int constantsSize, hlsl::Tex* tex, int texNum, int gridW, int gridH)
{
    return hlsl::reblur_dispatch_impl(shaderName, constants, constantsSize, tex, texNum, gridW, gridH);
}

int hlsl::reblur_dispatch_impl(const char* shaderName, const void* constants, int constantsSize, hlsl::Tex* tex, int texNum, int gridW, int gridH)
{
    (void)texNum;
    if (constantsSize < (int)sizeof(CB)) return -2;
    CB cb;
    memcpy(&cb, constants, sizeof(CB));
    const bool perf = !strncmp(shaderName, "REBLUR_Perf_", 12);
    Pass P(cb, perf);

    if (!strcmp(shaderName, "REBLUR_ClassifyTiles.cs"))
    {
        ClassifyTiles(P, tex, gridW, gridH);
        return 0;
    }
    if (strncmp(shaderName, "REBLUR_", 7) != 0) return -1;
    const char* p = shaderName + (perf ? 12 : 7);
    Signals sg{false, false};
    if (!strncmp(p, "DiffuseSpecular_", 16)) { sg = {true, true}; p += 16; }
    else if (!strncmp(p, "Diffuse_", 8)) { sg = {true, false}; p += 8; }
    else if (!strncmp(p, "Specular_", 9)) { sg = {false, true}; p += 9; }
    else return -1;

    const int W = gridW * 8, H = gridH * 16; // all these passes run 8x16 groups; overhanging threads early-out / drop stores
    if (!strcmp(p, "HitDistReconstruction.cs")) HitDistReconstruction(P, sg, 1, tex, W, H);
    else if (!strcmp(p, "HitDistReconstruction_5x5.cs")) HitDistReconstruction(P, sg, 2, tex, W, H);
    else if (!strcmp(p, "PrePass.cs")) PrePass(P, sg, tex, W, H);
    else if (!strcmp(p, "TemporalAccumulation.cs")) TemporalAccumulation(P, sg, tex, W, H);
    else if (!strcmp(p, "HistoryFix.cs")) HistoryFix(P, sg, tex, W, H);
    else if (!strcmp(p, "Blur.cs")) Blur(P, sg, tex, W, H);
    else if (!strcmp(p, "PostBlur.cs")) PostBlur(P, sg, false, tex, W, H);
    else if (!strcmp(p, "PostBlur_NoTemporalStabilization.cs")) PostBlur(P, sg, true, tex, W, H);
    else if (!strcmp(p, "TemporalStabilization.cs")) TemporalStabilization(P, sg, tex, W, H);
    else if (!strcmp(p, "SplitScreen.cs")) SplitScreen(P, sg, tex, W, H);
    else return -1;
    return 0;
}
