// ORACLE -- TEST INFRASTRUCTURE ONLY (see hlsl.h).  Every pass below is checked against the reference's own shader source of that pass,
// compiled for the CPU (oracle/build_refshaders.py), by tests/test_reference_shaders.py; the reference ships no golden vectors.
// CPU restatement of the reference's RELAX_DIFFUSE_SPECULAR passes (non-SH) at the default compile-time switches:
//   ClassifyTiles        Shaders/Source/RELAX_ClassifyTiles.cs.hlsl:18-49
//   PrePass              Shaders/Include/RELAX_PrePass.hlsli:13-347
//   TemporalAccumulation Shaders/Include/RELAX_TemporalAccumulation.hlsli:11-930
//   HistoryFix           Shaders/Include/RELAX_HistoryFix.hlsli:10-158
//   HistoryClamping      Shaders/Include/RELAX_HistoryClamping.hlsli:10-364
//   A-trous (SMEM)       Shaders/Include/RELAX_AtrousSmem.hlsli:11-472
//   A-trous              Shaders/Include/RELAX_Atrous.hlsli:11-243
//   helpers              Shaders/Include/RELAX_Common.hlsli:11-185, Common.hlsli
#include "common_hlsli.h"
#include "oracle.h"

#include <cstring>

namespace hlsl
{
using namespace common;
namespace
{
// RELAX_Config.hlsli:21-99 (+ gStepSize / gIsLastPass of the A-trous passes)
struct CB
{
    float4x4 gWorldToClip, gWorldToClipPrev, gWorldToViewPrev, gWorldPrevToWorld;
    float4 gRotatorPre, gFrustumRight, gFrustumUp, gFrustumForward, gPrevFrustumRight, gPrevFrustumUp, gPrevFrustumForward, gCameraDelta, gMvScale;
    float2 gJitter, gResolutionScale, gRectOffset, gResourceSizeInv, gResourceSize, gRectSizeInv, gRectSizePrev, gResourceSizeInvPrev;
    uint gPrintfAt[2], gRectOrigin[2];
    int gRectSize[2];
    float gSpecMaxAccumulatedFrameNum, gSpecMaxFastAccumulatedFrameNum, gDiffMaxAccumulatedFrameNum, gDiffMaxFastAccumulatedFrameNum, gDisocclusionThreshold,
        gDisocclusionThresholdAlternate, gCameraAttachedReflectionMaterialID, gStrandMaterialID, gStrandThickness, gRoughnessFraction, gSpecVarianceBoost, gSplitScreen,
        gDiffBlurRadius, gSpecBlurRadius, gDepthThreshold, gLobeAngleFraction, gSpecLobeAngleSlack, gHistoryFixEdgeStoppingNormalPower, gRoughnessEdgeStoppingRelaxation,
        gNormalEdgeStoppingRelaxation, gColorBoxSigmaScale, gHistoryAccelerationAmount, gHistoryResetTemporalSigmaScale, gHistoryResetSpatialSigmaScale, gHistoryResetAmount,
        gDenoisingRange, gSpecPhiLuminance, gDiffPhiLuminance, gDiffMaxLuminanceRelativeDifference, gSpecMaxLuminanceRelativeDifference, gLuminanceEdgeStoppingRelaxation,
        gConfidenceDrivenRelaxationMultiplier, gConfidenceDrivenLuminanceEdgeStoppingRelaxation, gConfidenceDrivenNormalEdgeStoppingRelaxation, gDebug, gOrthoMode, gUnproject,
        gFramerateScale, gCheckerboardResolveAccumSpeed, gJitterDelta, gHistoryFixFrameNum, gHistoryFixBasePixelStride, gHistoryThreshold, gViewZScale, gMinHitDistanceWeight,
        gDiffMinMaterial, gSpecMinMaterial;
    uint gRoughnessEdgeStoppingEnabled, gFrameIndex, gDiffCheckerboard, gSpecCheckerboard, gHasHistoryConfidence, gHasDisocclusionThresholdMix, gResetHistory;
    uint gStepSize, gIsLastPass;
};
static_assert(sizeof(CB) == 712, "RELAX_SHARED_CONSTANTS is 704 bytes + gStepSize + gIsLastPass");

const float RELAX_NORMAL_ULP = 1.5f / 255.0f;
const float RELAX_MAX_ACCUM_FRAME_NUM = 255.0f;
const float RELAX_ANTILAG_ACCELERATION_AMOUNT_SCALE = 10.0f;
const float NRD_CURVATURE_Z_THRESHOLD = 0.1f;

// Poisson.hlsli:40-50
const float3 g_Poisson8[8] = {float3(-0.4706069f, -0.4427112f, +0.6461146f), float3(-0.9057375f, +0.3003471f, +0.9542373f), float3(-0.3487388f, +0.4037880f, +0.5335386f),
                              float3(+0.1023042f, +0.6439373f, +0.6520134f), float3(+0.5699277f, +0.3513750f, +0.6695386f), float3(+0.2939128f, -0.1131226f, +0.3149309f),
                              float3(+0.7836658f, -0.4208784f, +0.8895339f), float3(+0.1564120f, -0.8198990f, +0.8346850f)};

float3 RgbToYCoCg(float3 c) { return float3(dot(c, float3(0.25f, 0.5f, 0.25f)), dot(c, float3(0.5f, 0.0f, -0.5f)), dot(c, float3(-0.25f, 0.5f, -0.25f))); }
float3 YCoCgToRgb(float3 c)
{
    float t = c.x - c.z;
    float3 r;
    r.y = c.x + c.z;
    r.x = t + c.y;
    r.z = t - c.y;
    return max(r, float3(0.0f));
}
float3 abs3(float3 a) { return float3(abs(a.x), abs(a.y), abs(a.z)); }

struct Pass
{
    const CB& c;
    explicit Pass(const CB& cb) : c(cb) {}
    float UnpackViewZ(float z) const { return abs(z * c.gViewZScale); }
    float2 ClampUvToViewport(float2 uv) const { return min(uv * c.gResolutionScale, c.gResolutionScale - float2(0.5f) * c.gResourceSizeInv); }
    float2 ResolutionScalePrev() const { return c.gRectSizePrev * c.gResourceSizeInvPrev; }
    float2 RectSize() const { return float2(float(c.gRectSize[0]), float(c.gRectSize[1])); }

    // RELAX_Common.hlsli:11-27
    static float4 UnpackPrevNormalRoughness(float4 p)
    {
        float4 r;
        r.set_xyz(_NRD_SafeNormalize(p.xyz() * float3(2.0f) - float3(1.0f)));
        r.w = p.w;
        return r;
    }
    static float4 PackPrevNormalRoughness(float4 nr) { return float4(nr.xyz() * float3(0.5f) + float3(0.5f), nr.w); }
    static float BilinearCustomFloat(float s00, float s10, float s01, float s11, float4 w)                    // :29-39
    {
        float o = s00 * w.x;
        o += s10 * w.y;
        o += s01 * w.z;
        o += s11 * w.w;
        float sum = dot(w, float4(1.0f));
        return sum < 0.0001f ? 0.0f : o * rcp(sum);
    }
    // :66-96
    // RELAX_Common.hlsli:75-80.  The shader sums forward + right * x - up * y left to right; this restatement (and the kernels, whose
    // world positions select history footprints and are pinned to it operation by operation) sums right * x - up * y first.  The two
    // differ by one rounding, which RELAX temporal accumulation amplifies (acos of nearly parallel view vectors, the difference of the
    // surface- and virtual-motion uv): tests/test_reference_shaders.py shows that with ORACLE_REFERENCE_ASSOCIATION (the "src" build of
    // the oracle) every RELAX pass is bit-identical to the reference's own shader, and what the difference costs without it.
    float3 GetCurrentWorldPosFromClipSpaceXY(float2 cs, float viewZ) const
    {
#ifdef ORACLE_REFERENCE_ASSOCIATION
        return c.gOrthoMode == 0.0f ? float3(viewZ) * (c.gFrustumForward.xyz() + c.gFrustumRight.xyz() * float3(cs.x) - c.gFrustumUp.xyz() * float3(cs.y))
                                    : float3(viewZ) * c.gFrustumForward.xyz() + c.gFrustumRight.xyz() * float3(cs.x) - c.gFrustumUp.xyz() * float3(cs.y);
#else
        float3 d = c.gFrustumRight.xyz() * float3(cs.x) - c.gFrustumUp.xyz() * float3(cs.y);
        return c.gOrthoMode == 0.0f ? float3(viewZ) * (c.gFrustumForward.xyz() + d) : float3(viewZ) * c.gFrustumForward.xyz() + d;
#endif
    }
    float3 GetCurrentWorldPosFromPixelPos(int2 p, float viewZ) const
    {
        float2 cs = (tofloat(p) + float2(0.5f)) * c.gRectSizeInv * float2(2.0f) - float2(1.0f);
        return GetCurrentWorldPosFromClipSpaceXY(cs, viewZ);
    }
    float3 GetPreviousWorldPosFromClipSpaceXY(float2 cs, float viewZ) const
    {
#ifdef ORACLE_REFERENCE_ASSOCIATION
        return c.gOrthoMode == 0.0f ? float3(viewZ) * (c.gPrevFrustumForward.xyz() + c.gPrevFrustumRight.xyz() * float3(cs.x) - c.gPrevFrustumUp.xyz() * float3(cs.y))
                                    : float3(viewZ) * c.gPrevFrustumForward.xyz() + c.gPrevFrustumRight.xyz() * float3(cs.x) - c.gPrevFrustumUp.xyz() * float3(cs.y);
#else
        float3 d = c.gPrevFrustumRight.xyz() * float3(cs.x) - c.gPrevFrustumUp.xyz() * float3(cs.y);
        return c.gOrthoMode == 0.0f ? float3(viewZ) * (c.gPrevFrustumForward.xyz() + d) : float3(viewZ) * c.gPrevFrustumForward.xyz() + d;
#endif
    }
    float3 GetPreviousWorldPosFromPixelPos(int2 p, float viewZ) const
    {
        float2 cs = (tofloat(p) + float2(0.5f)) * (float2(1.0f) / c.gRectSizePrev) * float2(2.0f) - float2(1.0f);
        return GetPreviousWorldPosFromClipSpaceXY(cs, viewZ);
    }
    static float GetPlaneDistanceWeight(float3 cw, float3 cn, float cz, float3 sw, float thr) { return abs(dot(sw - cw, cn)) / cz > thr ? 0.0f : 1.0f; }   // :98-103
    static float GetPlaneDistanceWeight_Atrous(float3 cw, float3 cn, float3 sw, float thr) { return abs(dot(sw - cw, cn)) < thr ? 1.0f : 0.0f; }           // :105-110
    static float GetSpecLobeTanHalfAngle(float roughness, float p = 0.75f)                                                                                     // :112-121
    {
        roughness = saturate(roughness);
        p = saturate(p);
        return roughness * roughness * p / (1.0f - p + NRD_EPS);
    }
    static float2 GetNormalWeightParams_ATrous(float roughness, float numFramesInHistory, float specConf, float relaxationK, float lobeFraction, float slack)  // :123-145
    {
        float relaxation = saturate(numFramesInHistory / 5.0f);
        relaxation *= lerp(1.0f, specConf, relaxationK);
        float f = 0.9f + 0.1f * relaxation;
        float angle = atan(GetSpecLobeTanHalfAngle(roughness, lobeFraction));
        angle *= 10.0f - 9.0f * relaxation;
        angle += slack;
        angle = min(Math::Pi(0.5f), angle);
        return float2(angle, f);
    }
    static float GetSpecularNormalWeight_ATrous(float2 p0, float3 n0, float3 n, float3 v0, float3 v)                                                          // :147-156
    {
        float cosa = min(dot(n0, n), dot(v0, v));
        float a = Math::AcosApprox(cosa);
        a = Math::SmoothStep(0.0f, p0.x, a);
        return saturate(1.0f - a * p0.y);
    }
    static float GetNormalWeightParam2(float roughness, float angleFraction)                                                                                   // :158-165
    {
        float angle = atan(GetSpecLobeTanHalfAngle(roughness, angleFraction));
        return 1.0f / max(angle, RELAX_NORMAL_ULP);
    }
    static float GetBilateralWeight(float z, float zc) { return Math::LinearStep(0.03f, 0.0f, abs(z - zc) * rcp(max(z, zc))); }                              // :167-169
    bool CompareMaterials(float m0, float m, float minm) const { return max(m0, minm) == max(m, minm); }

    // Common.hlsli helpers that depend on REBLUR-independent constants
    static float2 GetHitDistanceWeightParams(float hitDist, float nonLinearAccumSpeed, float roughness = 1.0f)
    {
        float smc = GetSpecMagicCurve(roughness);
        float norm = lerp(0.0005f, 1.0f, min(nonLinearAccumSpeed, smc));
        float a = 1.0f / norm;
        return float2(a, -hitDist * a);
    }
    static float2 GetRoughnessWeightParams(float roughness, float fraction, float sensitivity = 0.01f)
    {
        float a = 1.0f / lerp(sensitivity, 1.0f, saturate(roughness * fraction));
        return float2(a, -roughness * a);
    }
    static float GetNormalWeightParam(float nonLinearAccumSpeed, float lobeAngleFraction, float roughness = 1.0f) // Common.hlsli:486-500
    {
        float percentOfVolume = 0.75f * lerp(lobeAngleFraction, 1.0f, nonLinearAccumSpeed); // NRD_MAX_PERCENT_OF_LOBE_VOLUME
        float angle = atan(ImportanceSampling::GetSpecularLobeTanHalfAngle(roughness, percentOfVolume));
        return 1.0f / max(angle, NRD_NORMAL_ENCODING_ERROR);
    }
    static float2 GetRelaxedRoughnessWeightParams(float m, float fraction = 1.0f, float sensitivity = 0.01f)
    {
        float a = 1.0f / lerp(sensitivity, 1.0f, lerp(m * m, m, fraction));
        return float2(a, -m * a);
    }
    static float GetEncodingAwareNormalWeight(float3 Ncurr, float3 Nprev, float maxAngle, float curvatureAngle, float thresholdAngle, bool remap)
    {
        float angle = Math::AcosApprox(dot(Ncurr, Nprev));
        float w = Math::SmoothStep01(1.0f - (angle - curvatureAngle - thresholdAngle) / maxAngle);
        if (remap) w = Math::SmoothStep(0.05f, 0.95f, w);
        return w;
    }
    static float ComputeParallaxInPixels(float3 X, float2 uv0, const float4x4& m, float2 rectSize)
    {
        float2 uv = Geometry::GetScreenUv(m, X);
        return length((uv - uv0) * rectSize);
    }
    static float ApplyThinLensEquation(float O, float curvature) { return O / (2.0f * curvature * O + 1.0f); }
    static float3 GetXvirtual(float hitDist, float curvature, float3 X, float3 Xprev, float3 N, float3 V, float roughness)
    {
        float4 D = ImportanceSampling::GetSpecularDominantDirection(N, V, roughness);
        float3 Iw = V;
        float3 reflectionRay = D.xyz() * float3(hitDist);
        float3x3 basis = Geometry::GetBasis(N);
        float3 O = Geometry::RotateVector(basis, reflectionRay);
        O.z = -O.z;
        float mag = 1.0f / (2.0f * curvature * O.z - 1.0f);
        float f = length(X);
        f *= 1.0f - abs(dot(N, V));
        f *= max(curvature, 0.0f);
        mag *= 1.0f / (1.0f + f);
        float3 I = O * float3(mag);
        Iw = Iw * float3(length(I));
        float closeness = saturate(length(Iw) / (hitDist + NRD_EPS));
        float3 origin = lerp(Xprev, X, closeness * D.w);
        return origin - Iw * float3(D.w);
    }
    static float2 ApplyCheckerboardShift(float2 pos, uint mode, uint counter, uint frameIndex)
    {
        float2 pp = pos + float2(16384.0f);
        uint cb = Sequence::CheckerBoard(int2((int)pp.x, (int)pp.y), frameIndex);
        float shift = ((counter & 1) == 0) ? -1.0f : 1.0f;
        pos.x += shift * float(cb != mode && mode != 2);
        return pos;
    }
};

void ClassifyTiles(const Pass& P, Tex* t, int gridW, int gridH)
{
    const Tex& gIn_ViewZ = t[0];
    Tex& gOut_Tiles = t[1];
#pragma omp parallel for schedule(static)
    for (int ty = 0; ty < gridH; ty++)
        for (int tx = 0; tx < gridW; tx++)
        {
            int n = 0;
            for (int j = 0; j < 16; j++)
                for (int i = 0; i < 16; i++) n += abs(gIn_ViewZ.load(tx * 16 + i, ty * 16 + j).x) > P.c.gDenoisingRange ? 1 : 0;
            gOut_Tiles.store(tx, ty, float4(n == 256 ? 1.0f : 0.0f, 0, 0, 0));
        }
}

// RELAX_HitDistReconstruction.hlsli:10-155 (3x3: border 1, 5x5: border 2); two-signal binding layout: tiles, spec, diff, normal-roughness,
// viewZ | spec, diff
void HitDistReconstruction(const Pass& P, int border, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_Spec = t[1], &gIn_Diff = t[2], &gIn_Normal_Roughness = t[3], &gIn_ViewZ = t[4];
    Tex &gOut_Spec = t[5], &gOut_Diff = t[6];
    const int2 rectMax(c.gRectSize[0] - 1, c.gRectSize[1] - 1);
    const bool hasSpec = gIn_Spec.w != 0, hasDiff = gIn_Diff.w != 0; // absent signal = NULL texture (see kLayouts)
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 8; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float2 pixelUv = (float2(float(x), float(y)) + float2(0.5f)) * c.gRectSizeInv;
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            // the shared-memory tile of the reference holds clamped texels (Preload :13-32); an absent signal reads gDenoisingRange
            auto hitDistViewZ = [&](int2 p) {
                p = clamp(p, int2(0), rectMax);
                return float3(hasSpec ? gIn_Spec.load(p).w : c.gDenoisingRange, hasDiff ? gIn_Diff.load(p).w : c.gDenoisingRange, P.UnpackViewZ(gIn_ViewZ.load(p).x));
            };
            float3 centerHitdistViewZ = hitDistViewZ(pixelPos);
            float centerViewZ = centerHitdistViewZ.z;
            if (centerViewZ > c.gDenoisingRange) continue;
            float4 normalAndRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos));
            float3 centerNormal = normalAndRoughness.xyz();
            float centerRoughness = normalAndRoughness.w;

            float centerSpecularHitDist = centerHitdistViewZ.x;
            float2 relaxedRoughnessWeightParams = Pass::GetRelaxedRoughnessWeightParams(centerRoughness * centerRoughness);
            float specularNormalWeightParam = Pass::GetNormalWeightParam(1.0f, 1.0f, centerRoughness);
            float sumSpecularWeight = 1000.0f * float(centerSpecularHitDist != 0.0f);
            float sumSpecularHitDist = centerSpecularHitDist * sumSpecularWeight;
            float centerDiffuseHitDist = centerHitdistViewZ.y;
            float diffuseNormalWeightParam = Pass::GetNormalWeightParam(1.0f, 1.0f);
            float sumDiffuseWeight = 1000.0f * float(centerDiffuseHitDist != 0.0f);
            float sumDiffuseHitDist = centerDiffuseHitDist * sumDiffuseWeight;

            for (int dy = 0; dy <= border * 2; dy++)
                for (int dx = 0; dx <= border * 2; dx++)
                {
                    int2 o = int2(dx, dy) - int2(border);
                    if (o.x == 0 && o.y == 0) continue;
                    int2 pos = clamp(pixelPos + o, int2(0), rectMax);
                    float3 sampleNormal = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pos)).xyz();
                    float3 sampleHitdistViewZ = hitDistViewZ(pixelPos + o);
                    float sampleViewZ = sampleHitdistViewZ.z;
                    float angle = Math::AcosApprox(dot(centerNormal, sampleNormal));
                    float w = IsInScreenNearest(pixelUv + float2(float(o.x), float(o.y)) * c.gRectSizeInv);
                    w *= float(sampleViewZ < c.gDenoisingRange);
                    w *= GetGaussianWeight(length(float2(float(o.x), float(o.y))) * 0.5f);
                    w *= Pass::GetBilateralWeight(sampleViewZ, centerViewZ);

                    float specularWeight = w;
                    specularWeight *= ComputeExponentialWeight(angle, specularNormalWeightParam, 0.0f);
                    // (sic: the reference feeds the CENTRE roughness here, :117 -- the weight is exp(0) = 1)
                    specularWeight *= ComputeExponentialWeight(normalAndRoughness.w * normalAndRoughness.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);
                    float sampleSpecularHitDist = specularWeight == 0.0f ? 0.0f : sampleHitdistViewZ.x; // Denanify
                    specularWeight *= float(sampleSpecularHitDist != 0.0f);
                    sumSpecularHitDist += sampleSpecularHitDist * specularWeight;
                    sumSpecularWeight += specularWeight;

                    float diffuseWeight = w;
                    diffuseWeight *= ComputeExponentialWeight(angle, diffuseNormalWeightParam, 0.0f);
                    float sampleDiffuseHitDist = diffuseWeight == 0.0f ? 0.0f : sampleHitdistViewZ.y;
                    diffuseWeight *= float(sampleDiffuseHitDist != 0.0f);
                    sumDiffuseHitDist += diffuseWeight == 0.0f ? 0.0f : sampleDiffuseHitDist * diffuseWeight;
                    sumDiffuseWeight += diffuseWeight;
                }
            sumSpecularHitDist /= max(sumSpecularWeight, 1e-6f);
            gOut_Spec.store(pixelPos, float4(gIn_Spec.load(pixelPos).xyz(), sumSpecularHitDist));
            sumDiffuseHitDist /= max(sumDiffuseWeight, 1e-6f);
            gOut_Diff.store(pixelPos, float4(gIn_Diff.load(pixelPos).xyz(), sumDiffuseHitDist));
        }
}

void PrePass(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_Spec = t[1], &gIn_Diff = t[2], &gIn_Normal_Roughness = t[3], &gIn_ViewZ = t[4];
    Tex &gOut_Spec = t[5], &gOut_Diff = t[6];
    const float2 rectSize = P.RectSize();
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 16; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            float centerViewZ = P.UnpackViewZ(gIn_ViewZ.load(pixelPos).x);
            if (centerViewZ > c.gDenoisingRange) continue;

            uint checkerboard = Sequence::CheckerBoard(pixelPos, c.gFrameIndex);
            int cx0 = max(x - 1, 0), cx1 = min(x + 1, c.gRectSize[0] - 1);
            float materialID0 = 0, materialID1 = 0;
            float2 checkerboardResolveWeights(1.0f);
            if (c.gSpecCheckerboard != 2 || c.gDiffCheckerboard != 2)
            {
                float viewZ0 = P.UnpackViewZ(gIn_ViewZ.load(cx0, y).x), viewZ1 = P.UnpackViewZ(gIn_ViewZ.load(cx1, y).x);
                NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(cx0, y), materialID0);
                NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(cx1, y), materialID1);
                checkerboardResolveWeights = float2(Pass::GetBilateralWeight(viewZ0, centerViewZ), Pass::GetBilateralWeight(viewZ1, centerViewZ));
                checkerboardResolveWeights.x = (viewZ0 > c.gDenoisingRange || x < 1) ? 0.0f : checkerboardResolveWeights.x;
                checkerboardResolveWeights.y = (viewZ1 > c.gDenoisingRange || x > c.gRectSize[0] - 2) ? 0.0f : checkerboardResolveWeights.y;
            }
            int cbx0 = cx0 >> 1, cbx1 = cx1 >> 1;

            float centerMaterialID;
            float4 centerNormalRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos), centerMaterialID);
            float3 centerNormal = centerNormalRoughness.xyz();
            float centerRoughness = centerNormalRoughness.w;
            float3 centerWorldPos = P.GetCurrentWorldPosFromPixelPos(pixelPos, centerViewZ);
            float4 rotator = c.gRotatorPre;
            float2 pixelUv = (tofloat(pixelPos) + float2(0.5f)) * c.gRectSizeInv;

            // ---- diffuse
            {
                bool diffHasData = true;
                int2 diffPos = pixelPos;
                if (c.gDiffCheckerboard != 2) { diffHasData = checkerboard == c.gDiffCheckerboard; diffPos.x >>= 1; }
                float4 diffuseIllumination = gIn_Diff.load(diffPos);
                if (!diffHasData)
                {
                    float2 wc = checkerboardResolveWeights;
                    wc.x *= float(P.CompareMaterials(centerMaterialID, materialID0, c.gDiffMinMaterial));
                    wc.y *= float(P.CompareMaterials(centerMaterialID, materialID1, c.gDiffMinMaterial));
                    wc *= float2(Math::PositiveRcp(wc.x + wc.y));
                    float4 d0 = gIn_Diff.load(cbx0, y), d1 = gIn_Diff.load(cbx1, y);
                    d0 = wc.x == 0.0f ? float4(0.0f) : d0;
                    d1 = wc.y == 0.0f ? float4(0.0f) : d1;
                    diffuseIllumination = d0 * float4(wc.x) + d1 * float4(wc.y);
                }
                if (c.gDiffBlurRadius > 0.0f)
                {
                    float frustumSize = PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, float(min(c.gRectSize[0], c.gRectSize[1])), centerViewZ);
                    float hitDist = diffuseIllumination.w == 0.0f ? 1.0f : diffuseIllumination.w;
                    float hitDistFactor = saturate(hitDist / frustumSize);
                    float blurRadius = c.gDiffBlurRadius * hitDistFactor;
                    if (diffuseIllumination.w == 0.0f) blurRadius = max(blurRadius, 1.0f);
                    float normalWeightParam = Pass::GetNormalWeightParam2(1.0f, 0.25f * c.gLobeAngleFraction);
                    float2 hitDistanceWeightParams = Pass::GetHitDistanceWeightParams(diffuseIllumination.w, 1.0f / 9.0f);
                    float weightSum = 1.0f;
                    float diffMinHitDistanceWeight = c.gMinHitDistanceWeight;
                    for (uint i = 0; i < 8; i++)
                    {
                        float3 offset = g_Poisson8[i];
                        float2 uv = pixelUv * rectSize + Geometry::RotateVector(rotator, offset.xy()) * float2(blurRadius);
                        uv = floor(uv) + float2(0.5f);
                        uv = Pass::ApplyCheckerboardShift(uv, c.gDiffCheckerboard, i, c.gFrameIndex) * c.gRectSizeInv;
                        float2 uvScaled = P.ClampUvToViewport(uv);
                        float2 checkerboardUvScaled = float2(uvScaled.x * (c.gDiffCheckerboard != 2 ? 0.5f : 1.0f), uvScaled.y);
                        float sampleMaterialID;
                        float3 sampleNormal = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.sampleNearest(uvScaled), sampleMaterialID).xyz();
                        float sampleViewZ = P.UnpackViewZ(gIn_ViewZ.sampleNearest(uvScaled).x);
                        float3 sampleWorldPos = P.GetCurrentWorldPosFromClipSpaceXY(uv * float2(2.0f) - float2(1.0f), sampleViewZ);
                        float sampleWeight = IsInScreenNearest(uv);
                        sampleWeight *= float(sampleViewZ < c.gDenoisingRange);
                        sampleWeight *= float(P.CompareMaterials(centerMaterialID, sampleMaterialID, c.gDiffMinMaterial));
                        sampleWeight *= Pass::GetPlaneDistanceWeight(centerWorldPos, centerNormal, c.gOrthoMode == 0.0f ? centerViewZ : 1.0f, sampleWorldPos, c.gDepthThreshold);
                        float angle = Math::AcosApprox(dot(centerNormal, sampleNormal));
                        sampleWeight *= ComputeWeight(angle, normalWeightParam, 0.0f);
                        float4 s = gIn_Diff.sampleNearest(checkerboardUvScaled);
                        s = sampleWeight == 0.0f ? float4(0.0f) : s;
                        sampleWeight *= lerp(diffMinHitDistanceWeight, 1.0f, ComputeExponentialWeight(s.w, hitDistanceWeightParams.x, hitDistanceWeightParams.y));
                        sampleWeight *= GetGaussianWeight(offset.z);
                        weightSum += sampleWeight;
                        diffuseIllumination += s * float4(sampleWeight);
                    }
                    diffuseIllumination /= float4(weightSum);
                }
                gOut_Diff.store(pixelPos, min(max(diffuseIllumination, float4(0.0f)), float4(NRD_FP16_MAX)));
            }
            // ---- specular
            {
                bool specHasData = true;
                int2 specPos = pixelPos;
                if (c.gSpecCheckerboard != 2) { specHasData = checkerboard == c.gSpecCheckerboard; specPos.x >>= 1; }
                float4 specularIllumination = gIn_Spec.load(specPos);
                if (!specHasData)
                {
                    float2 wc = checkerboardResolveWeights;
                    wc.x *= float(P.CompareMaterials(centerMaterialID, materialID0, c.gSpecMinMaterial));
                    wc.y *= float(P.CompareMaterials(centerMaterialID, materialID1, c.gSpecMinMaterial));
                    wc *= float2(Math::PositiveRcp(wc.x + wc.y));
                    float4 s0 = gIn_Spec.load(cbx0, y), s1 = gIn_Spec.load(cbx1, y);
                    s0 = wc.x == 0.0f ? float4(0.0f) : s0;
                    s1 = wc.y == 0.0f ? float4(0.0f) : s1;
                    specularIllumination = s0 * float4(wc.x) + s1 * float4(wc.y);
                }
                specularIllumination.w = max(0.0f, min(c.gDenoisingRange, specularIllumination.w));
                if (c.gSpecBlurRadius > 0.0f)
                {
                    float3 viewVector = c.gOrthoMode == 0.0f ? normalize(-centerWorldPos) : c.gFrustumForward.xyz();
                    float4 D = ImportanceSampling::GetSpecularDominantDirection(centerNormal, viewVector, centerRoughness);
                    float NoD = abs(dot(centerNormal, D.xyz()));
                    float frustumSize = PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, float(min(c.gRectSize[0], c.gRectSize[1])), centerViewZ);
                    float hitDist = specularIllumination.w == 0.0f ? 1.0f : specularIllumination.w;
                    float hitDistFactor = saturate(hitDist * NoD / frustumSize);
                    float smc = GetSpecMagicCurve(centerRoughness);
                    float blurRadius = c.gSpecBlurRadius * hitDistFactor * smc;
                    float lobeTanHalfAngle = ImportanceSampling::GetSpecularLobeTanHalfAngle(centerRoughness);
                    float lobeRadius = hitDist * NoD * lobeTanHalfAngle;
                    float minBlurRadius = lobeRadius / PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, 1.0f, centerViewZ + hitDist * D.w);
                    blurRadius = min(blurRadius, minBlurRadius);
                    if (specularIllumination.w == 0.0f) blurRadius = max(blurRadius, 1.0f);
                    float normalWeightParam = Pass::GetNormalWeightParam2(centerRoughness, 0.5f * c.gLobeAngleFraction);
                    float2 hitDistanceWeightParams = Pass::GetHitDistanceWeightParams(specularIllumination.w, 1.0f / 9.0f, centerRoughness);
                    float2 roughnessWeightParams = Pass::GetRoughnessWeightParams(centerRoughness, c.gRoughnessFraction);
                    float specMinHitDistanceWeight = specularIllumination.w == 0.0f ? 1.0f : c.gMinHitDistanceWeight * smc;
                    float specularHitT = specularIllumination.w == 0.0f ? c.gDenoisingRange : specularIllumination.w;
                    float minHitT = specularHitT == 0.0f ? NRD_INF : specularHitT;
                    float weightSum = 1.0f;
                    for (uint i = 0; i < 8; i++)
                    {
                        float3 offset = g_Poisson8[i];
                        float2 uv = pixelUv * rectSize + Geometry::RotateVector(rotator, offset.xy()) * float2(blurRadius);
                        uv = floor(uv) + float2(0.5f);
                        uv = Pass::ApplyCheckerboardShift(uv, c.gSpecCheckerboard, i, c.gFrameIndex) * c.gRectSizeInv;
                        float2 uvScaled = P.ClampUvToViewport(uv);
                        float2 checkerboardUvScaled = float2(uvScaled.x * (c.gSpecCheckerboard != 2 ? 0.5f : 1.0f), uvScaled.y);
                        float sampleMaterialID;
                        float4 snr = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.sampleNearest(uvScaled), sampleMaterialID);
                        float3 sampleNormal = snr.xyz();
                        float sampleRoughness = snr.w;
                        float sampleViewZ = P.UnpackViewZ(gIn_ViewZ.sampleNearest(uvScaled).x);
                        float sampleWeight = IsInScreenNearest(uv);
                        sampleWeight *= float(sampleViewZ < c.gDenoisingRange);
                        sampleWeight *= float(P.CompareMaterials(centerMaterialID, sampleMaterialID, c.gSpecMinMaterial));
                        sampleWeight *= ComputeWeight(sampleRoughness, roughnessWeightParams.x, roughnessWeightParams.y);
                        float angle = Math::AcosApprox(dot(centerNormal, sampleNormal));
                        sampleWeight *= ComputeWeight(angle, normalWeightParam, 0.0f);
                        float3 sampleWorldPos = P.GetCurrentWorldPosFromClipSpaceXY(uv * float2(2.0f) - float2(1.0f), sampleViewZ);
                        sampleWeight *= Pass::GetPlaneDistanceWeight(centerWorldPos, centerNormal, c.gOrthoMode == 0.0f ? centerViewZ : 1.0f, sampleWorldPos, c.gDepthThreshold);
                        float4 s = gIn_Spec.sampleNearest(checkerboardUvScaled);
                        s = sampleWeight == 0.0f ? float4(0.0f) : s;
                        sampleWeight *= lerp(specMinHitDistanceWeight, 1.0f, ComputeExponentialWeight(s.w, hitDistanceWeightParams.x, hitDistanceWeightParams.y));
                        sampleWeight *= GetGaussianWeight(offset.z);
                        float d = length(sampleWorldPos - centerWorldPos);
                        float h = s.w;
                        float tt = h / (specularIllumination.w + d);
                        sampleWeight *= lerp(saturate(tt), 1.0f, Math::LinearStep(0.5f, 1.0f, centerRoughness));
                        weightSum += sampleWeight;
                        specularIllumination.x += s.x * sampleWeight;
                        specularIllumination.y += s.y * sampleWeight;
                        specularIllumination.z += s.z * sampleWeight;
                        if (sampleWeight != 0.0f) minHitT = min(minHitT, s.w == 0.0f ? NRD_INF : s.w);
                    }
                    specularIllumination.x /= weightSum;
                    specularIllumination.y /= weightSum;
                    specularIllumination.z /= weightSum;
                    specularIllumination.w = minHitT == NRD_INF ? 0.0f : minHitT;
                }
                gOut_Spec.store(pixelPos, min(max(specularIllumination, float4(0.0f)), float4(NRD_FP16_MAX)));
            }
        }
}

void TemporalAccumulation(const Pass& P, Tex* t, int gridW, int gridH, bool hasDiff, bool hasSpec)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_Spec = t[1], &gIn_Diff = t[2], &gIn_Mv = t[3], &gIn_Normal_Roughness = t[4], &gIn_ViewZ = t[5], &gHistory_SpecFast = t[6],
              &gHistory_DiffFast = t[7], &gHistory_Spec = t[8], &gHistory_Diff = t[9], &gPrev_Normal_Roughness = t[10], &gPrev_ViewZ = t[11], &gPrev_SpecHitDist = t[12],
              &gPrev_HistoryLength = t[13], &gPrev_MaterialID = t[14], &gIn_SpecConfidence = t[15], &gIn_DiffConfidence = t[16], &gIn_DisocclusionThresholdMix = t[17];
    Tex &gOut_Spec = t[18], &gOut_Diff = t[19], &gOut_SpecFast = t[20], &gOut_DiffFast = t[21], &gOut_SpecHitDist = t[22], &gOut_HistoryLength = t[23], &gOut_SpecReprojectionConfidence = t[24];
    const float2 rectSize = P.RectSize();
    const int2 rectMax(c.gRectSize[0] - 1, c.gRectSize[1] - 1);
    auto wzxy = [](float4 g) { return float4(g.w, g.z, g.x, g.y); };

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            float currentLinearZ = P.UnpackViewZ(gIn_ViewZ.load(pixelPos).x);
            if (currentLinearZ > c.gDenoisingRange) continue;

            auto sNormalSpecHitT = [&](int i, int j) { // offsets relative to the pixel, clamped (Preload :360-374)
                int2 p = clamp(int2(x + i, y + j), int2(0), rectMax);
                float4 nr = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(p));
                nr.w = gIn_Spec.load(p).w;
                return nr;
            };

            float currentMaterialID;
            float4 currentNormalRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos), currentMaterialID);
            float3 currentNormal = currentNormalRoughness.xyz();
            float currentRoughness = currentNormalRoughness.w;
            float3 currentWorldPos = P.GetCurrentWorldPosFromPixelPos(pixelPos, currentLinearZ);
            float3 currentViewVector = c.gOrthoMode == 0.0f ? currentWorldPos : float3(currentLinearZ) * normalize(c.gFrustumForward.xyz());
            float3 V = -normalize(currentViewVector);
            float NoV = abs(dot(currentNormal, V));

            float2 pixelUv = (tofloat(pixelPos) + float2(0.5f)) * c.gRectSizeInv;
            float3 mv = gIn_Mv.load(pixelPos).xyz() * c.gMvScale.xyz();
            float3 prevWorldPos = currentWorldPos;
            float2 prevUVSMB = pixelUv + mv.xy();
            if (c.gMvScale.w == 0.0f)
            {
                if (c.gMvScale.z == 0.0f) mv.z = Geometry::AffineTransform(c.gWorldToViewPrev, currentWorldPos).z - currentLinearZ;
                prevWorldPos = P.GetPreviousWorldPosFromClipSpaceXY(prevUVSMB * float2(2.0f) - float2(1.0f), currentLinearZ + mv.z) + c.gCameraDelta.xyz();
            }
            else
            {
                prevWorldPos += mv;
                prevUVSMB = Geometry::GetScreenUv(c.gWorldToClipPrev, prevWorldPos);
            }

            float3 diffuseIllumination = gIn_Diff.load(pixelPos).xyz();
            float4 specularIllumination = gIn_Spec.load(pixelPos);

            float hitTM1 = sNormalSpecHitT(0, 0).w;
            float minHitDist3x3 = hitTM1 == 0.0f ? NRD_INF : hitTM1;
            float3 currentNormalAveraged = currentNormal;
            for (int i = -1; i <= 1; i++)
                for (int j = -1; j <= 1; j++)
                {
                    if (i == 0 && j == 0) continue;
                    float4 ns = sNormalSpecHitT(i, j);
                    minHitDist3x3 = min(minHitDist3x3, ns.w == 0.0f ? NRD_INF : ns.w);
                    currentNormalAveraged += ns.xyz();
                }
            currentNormalAveraged /= float3(9.0f);
            float currentRoughnessModified = Filtering::GetModifiedRoughnessFromNormalVariance(currentRoughness, currentNormalAveraged);

            float specular1stMoment = Color::Luminance(specularIllumination.xyz());
            float specular2ndMoment = specular1stMoment * specular1stMoment;
            float diffuse1stMoment = Color::Luminance(diffuseIllumination);
            float diffuse2ndMoment = diffuse1stMoment * diffuse1stMoment;

            float smbParallaxInPixels1 = Pass::ComputeParallaxInPixels(prevWorldPos + c.gCameraDelta.xyz(), c.gOrthoMode == 0.0f ? prevUVSMB : pixelUv, c.gWorldToClipPrev, rectSize);
            float smbParallaxInPixels2 = Pass::ComputeParallaxInPixels(prevWorldPos - c.gCameraDelta.xyz(), c.gOrthoMode == 0.0f ? pixelUv : prevUVSMB, c.gWorldToClip, rectSize);
            float smbParallaxInPixelsMax = max(smbParallaxInPixels1, smbParallaxInPixels2);
            float smbParallaxInPixelsMin = min(smbParallaxInPixels1, smbParallaxInPixels2);
            float pixelSize = PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, 1.0f, currentLinearZ);

            float disocclusionThresholdMix = 0.0f;
            if (currentMaterialID == c.gStrandMaterialID) disocclusionThresholdMix = pixelSize / (pixelSize + c.gStrandThickness);
            if (c.gHasDisocclusionThresholdMix) disocclusionThresholdMix = gIn_DisocclusionThresholdMix.load(pixelPos).x;
            float disocclusionThreshold = lerp(c.gDisocclusionThreshold, c.gDisocclusionThresholdAlternate, disocclusionThresholdMix);

            // ================= loadSurfaceMotionBasedPrevData (:35-229)
            float footprintQuality, historyLength, prevReflectionHitTSMB;
            float4 prevDiffSMB, prevSpecSMB;
            float3 prevDiffSMBResponsive, prevSpecSMBResponsive;
            float SMBReprojectionFound;
            {
                float3 currentNormalN = normalize(currentNormalAveraged);
                float2 prevPixelPosFloat = prevUVSMB * c.gRectSizePrev;
                float2 fl = floor(prevPixelPosFloat - float2(0.5f));
                int2 bilinearOrigin((int)fl.x, (int)fl.y);
                float2 bilinearWeights = frac(prevPixelPosFloat - float2(0.5f));
                float2 bo = tofloat(bilinearOrigin);
                float2 g00 = (bo + float2(0.0f, 0.0f)) * c.gResourceSizeInvPrev, g10 = (bo + float2(2.0f, 0.0f)) * c.gResourceSizeInvPrev;
                float2 g01 = (bo + float2(0.0f, 2.0f)) * c.gResourceSizeInvPrev, g11 = (bo + float2(2.0f, 2.0f)) * c.gResourceSizeInvPrev;
                auto unpack4 = [&](float4 v) { return float4(P.UnpackViewZ(v.x), P.UnpackViewZ(v.y), P.UnpackViewZ(v.z), P.UnpackViewZ(v.w)); };
                float4 z00 = unpack4(wzxy(gPrev_ViewZ.gather(g00, 0))), z10 = unpack4(wzxy(gPrev_ViewZ.gather(g10, 0)));
                float4 z01 = unpack4(wzxy(gPrev_ViewZ.gather(g01, 0))), z11 = unpack4(wzxy(gPrev_ViewZ.gather(g11, 0)));
                float4 m00 = wzxy(gPrev_MaterialID.gather(g00, 0)) * float4(255.0f), m10 = wzxy(gPrev_MaterialID.gather(g10, 0)) * float4(255.0f);
                float4 m01 = wzxy(gPrev_MaterialID.gather(g01, 0)) * float4(255.0f), m11 = wzxy(gPrev_MaterialID.gather(g11, 0)) * float4(255.0f);

                float frustumSize = pixelSize * float(min(c.gRectSize[0], c.gRectSize[1]));
                float slopeScale = 1.0f / lerp(lerp(0.05f, 1.0f, NoV), 1.0f, saturate(smbParallaxInPixelsMax / 30.0f));
                float4 thr = float4(saturate(disocclusionThreshold * slopeScale) * frustumSize);
                thr *= IsInScreenBilinear(bo, c.gRectSizePrev);
                thr -= float4(NRD_EPS);

                float3 prevViewPos = Geometry::AffineTransform(c.gWorldToViewPrev, prevWorldPos);
                float3 pz(prevViewPos.z);
                float3 v0 = step(abs(float3(z00.y, z00.z, z00.w) - pz), float3(thr.x)), v1 = step(abs(float3(z10.x, z10.z, z10.w) - pz), float3(thr.y));
                float3 v2 = step(abs(float3(z01.x, z01.y, z01.w) - pz), float3(thr.z)), v3 = step(abs(float3(z11.x, z11.y, z11.z) - pz), float3(thr.w));
                float minMaterialID = min(c.gSpecMinMaterial, c.gDiffMinMaterial);
                auto cm = [&](float m) { return float(P.CompareMaterials(currentMaterialID, m, minMaterialID)); };
                v0 *= float3(cm(m00.y), cm(m00.z), cm(m00.w));
                v1 *= float3(cm(m10.x), cm(m10.z), cm(m10.w));
                v2 *= float3(cm(m01.x), cm(m01.y), cm(m01.w));
                v3 *= float3(cm(m11.x), cm(m11.y), cm(m11.z));
                float bicubicFootprintValid = dot(v0 + v1 + v2 + v3, float3(1.0f)) > 11.5f ? 1.0f : 0.0f;
                float4 bilinearTapsValid(v0.z, v1.y, v2.y, v3.x);

                float2 uvn = (bo + float2(1.0f)) * c.gResourceSizeInvPrev;
                float3 prevNormalFlat = Pass::UnpackPrevNormalRoughness(gPrev_Normal_Roughness.sampleLinear(uvn)).xyz();
                prevNormalFlat = Geometry::RotateVector(c.gWorldPrevToWorld, prevNormalFlat);
                if (dot(currentNormalN, prevNormalFlat) < 0.0f) { bilinearTapsValid = float4(0.0f); bicubicFootprintValid = 0.0f; }

                Filtering::Bilinear bil;
                bil.origin = bo;
                bil.weights = bilinearWeights;
                float4 bcw = Filtering::GetBilinearCustomWeights(bil, bilinearTapsValid);
                bool useBicubic = bicubicFootprintValid > 0.0f;
                prevDiffSMB = max(BicubicCustom(prevPixelPosFloat, c.gResourceSizeInvPrev, bcw, useBicubic, gHistory_Diff), float4(0.0f));
                prevSpecSMB = max(BicubicCustom(prevPixelPosFloat, c.gResourceSizeInvPrev, bcw, useBicubic, gHistory_Spec), float4(0.0f));
                prevDiffSMBResponsive = max(BicubicCustom(prevPixelPosFloat, c.gResourceSizeInvPrev, bcw, useBicubic, gHistory_DiffFast).xyz(), float3(0.0f));
                prevSpecSMBResponsive = max(BicubicCustom(prevPixelPosFloat, c.gResourceSizeInvPrev, bcw, useBicubic, gHistory_SpecFast).xyz(), float3(0.0f));

                float4 hl = wzxy(gPrev_HistoryLength.gather(uvn, 0));
                historyLength = 255.0f * Pass::BilinearCustomFloat(hl.x, hl.y, hl.z, hl.w, bcw);
                float4 ht = wzxy(gPrev_SpecHitDist.gather(uvn, 0));
                prevReflectionHitTSMB = max(0.001f, Pass::BilinearCustomFloat(ht.x, ht.y, ht.z, ht.w, bcw));

                SMBReprojectionFound = bicubicFootprintValid > 0.0f ? 2.0f : 1.0f;
                footprintQuality = bicubicFootprintValid > 0.0f ? 1.0f : dot(bcw, float4(1.0f));
                if (!(bilinearTapsValid.x != 0.0f || bilinearTapsValid.y != 0.0f || bilinearTapsValid.z != 0.0f || bilinearTapsValid.w != 0.0f))
                {
                    SMBReprojectionFound = 0.0f;
                    footprintQuality = 0.0f;
                }
            }

            historyLength = historyLength + 1.0f;
            historyLength = min(RELAX_MAX_ACCUM_FRAME_NUM, historyLength);

            float3 Vprev = c.gOrthoMode == 0.0f ? -normalize(prevWorldPos - c.gCameraDelta.xyz()) : -normalize(c.gPrevFrustumForward.xyz());
            float NoVprev = abs(dot(currentNormal, Vprev));
            float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f);
            sizeQuality *= sizeQuality;
            sizeQuality *= sizeQuality;
            footprintQuality *= lerp(0.1f, 1.0f, saturate(sizeQuality + abs(c.gOrthoMode)));
            if (footprintQuality < 1.0f)
            {
                historyLength *= sqrt(footprintQuality);
                historyLength = max(historyLength, 1.0f);
            }
            historyLength = c.gResetHistory != 0 ? 1.0f : historyLength;
            // :568-574: only the signals the shader was compiled for take part
            float maxAccumulatedFrameNum = 1.0f + (hasDiff && hasSpec ? max(c.gDiffMaxAccumulatedFrameNum, c.gSpecMaxAccumulatedFrameNum)
                                                                     : (hasDiff ? c.gDiffMaxAccumulatedFrameNum : c.gSpecMaxAccumulatedFrameNum));
            historyLength = min(historyLength, maxAccumulatedFrameNum);

            uint checkerboard = Sequence::CheckerBoard(pixelPos, c.gFrameIndex);

            // ---- diffuse (:579-617)
            {
                float diffMaxAccumulatedFrameNum = c.gDiffMaxAccumulatedFrameNum, diffMaxFastAccumulatedFrameNum = c.gDiffMaxFastAccumulatedFrameNum;
                if (c.gHasHistoryConfidence)
                {
                    float conf = gIn_DiffConfidence.load(pixelPos).x;
                    diffMaxAccumulatedFrameNum *= conf;
                    diffMaxFastAccumulatedFrameNum *= conf;
                }
                float diffHistoryLength = historyLength;
                float diffuseAlpha = SMBReprojectionFound > 0.0f ? max(1.0f / (diffMaxAccumulatedFrameNum + 1.0f), 1.0f / diffHistoryLength) : 1.0f;
                float diffuseAlphaResponsive = SMBReprojectionFound > 0.0f ? max(1.0f / (diffMaxFastAccumulatedFrameNum + 1.0f), 1.0f / diffHistoryLength) : 1.0f;
                bool diffHasData = true;
                if (c.gDiffCheckerboard != 2) diffHasData = checkerboard == c.gDiffCheckerboard;
                if (!diffHasData && diffHistoryLength > 1.0f)
                {
                    diffuseAlpha *= 1.0f - c.gCheckerboardResolveAccumSpeed;
                    diffuseAlphaResponsive *= 1.0f - c.gCheckerboardResolveAccumSpeed;
                }
                float4 acc = lerp(prevDiffSMB, float4(diffuseIllumination, diffuse2ndMoment), diffuseAlpha);
                float3 accResp = lerp(prevDiffSMBResponsive, diffuseIllumination, diffuseAlphaResponsive);
                gOut_Diff.store(pixelPos, acc);
                gOut_DiffFast.store(pixelPos, float4(accResp, 0.0f));
            }
            gOut_HistoryLength.store(pixelPos, historyLength / 255.0f);

            // ---- specular (:625-928)
            float specMaxAccumulatedFrameNum = c.gSpecMaxAccumulatedFrameNum, specMaxFastAccumulatedFrameNum = c.gSpecMaxFastAccumulatedFrameNum;
            if (c.gHasHistoryConfidence)
            {
                float conf = gIn_SpecConfidence.load(pixelPos).x;
                specMaxAccumulatedFrameNum *= conf;
                specMaxFastAccumulatedFrameNum *= conf;
            }
            float specHistoryLength = historyLength;
            float specHistoryFrames = min(specMaxAccumulatedFrameNum, specHistoryLength);
            float specHistoryResponsiveFrames = min(specMaxFastAccumulatedFrameNum, specHistoryLength);
            float hitDist = minHitDist3x3 == NRD_INF ? 0.0f : minHitDist3x3;

            float curvature = 0.0f;
            {
                float2 uvForZeroParallax = c.gOrthoMode == 0.0f ? prevUVSMB : pixelUv;
                float2 deltaUv = uvForZeroParallax - Geometry::GetScreenUv(c.gWorldToClipPrev, prevWorldPos + c.gCameraDelta.xyz());
                deltaUv *= rectSize;
                deltaUv /= float2(max(smbParallaxInPixels1, 1.0f / 256.0f));
                float3 n10, x10, n01, x01;
                {
                    float3 xx = P.GetCurrentWorldPosFromClipSpaceXY((pixelUv + float2(1, 0) * c.gRectSizeInv) * float2(2.0f) - float2(1.0f), 1.0f);
                    float3 v = c.gOrthoMode == 0.0f ? normalize(-xx) : c.gFrustumForward.xyz();
                    float3 o = c.gOrthoMode == 0.0f ? float3(0.0f) : xx;
                    x10 = o + v * float3(dot(currentWorldPos - o, currentNormal) / dot(currentNormal, v));
                    n10 = sNormalSpecHitT(1, 0).xyz();
                }
                {
                    float3 xx = P.GetCurrentWorldPosFromClipSpaceXY((pixelUv + float2(0, 1) * c.gRectSizeInv) * float2(2.0f) - float2(1.0f), 1.0f);
                    float3 v = c.gOrthoMode == 0.0f ? normalize(-xx) : c.gFrustumForward.xyz();
                    float3 o = c.gOrthoMode == 0.0f ? float3(0.0f) : xx;
                    x01 = o + v * float3(dot(currentWorldPos - o, currentNormal) / dot(currentNormal, v));
                    n01 = sNormalSpecHitT(0, 1).xyz();
                }
                float2 w = abs(deltaUv) + float2(1.0f / 256.0f);
                w /= float2(w.x + w.y);
                float3 xm = x10 * float3(w.x) + x01 * float3(w.y);
                float3 n = normalize(n10 * float3(w.x) + n01 * float3(w.y));
                float deltaUvLenFixed = smbParallaxInPixelsMin;
                deltaUvLenFixed *= 1.0f + c.gFramerateScale * Sequence::Bayer4x4(pixelPos, c.gFrameIndex);
                float2 motionUvHigh = pixelUv + float2(deltaUvLenFixed) * deltaUv * c.gRectSizeInv;
                motionUvHigh = (floor(motionUvHigh * rectSize) + float2(0.5f)) * c.gRectSizeInv;
                if (deltaUvLenFixed > 1.0f && IsInScreenNearest(motionUvHigh) != 0.0f)
                {
                    float2 uvScaled = P.ClampUvToViewport(motionUvHigh);
                    float zHigh = P.UnpackViewZ(gIn_ViewZ.sampleNearest(uvScaled).x);
                    float3 xHigh = P.GetCurrentWorldPosFromClipSpaceXY(motionUvHigh * float2(2.0f) - float2(1.0f), zHigh);
                    float3 nHigh = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.sampleNearest(uvScaled)).xyz();
                    float zError = abs(zHigh - currentLinearZ) * rcp(max(zHigh, currentLinearZ));
                    bool cmp = zError < NRD_CURVATURE_Z_THRESHOLD;
                    n = cmp ? nHigh : n;
                    xm = cmp ? xHigh : xm;
                }
                float3 edge = xm - currentWorldPos;
                curvature = dot(n - currentNormal, edge) * Math::PositiveRcp(Math::LengthSquared(edge));
            }
            float hitDistFocused = Pass::ApplyThinLensEquation(hitDist, curvature);

            // ================= loadVirtualMotionBasedPrevData (:231-357)
            float4 prevSpecVMB(0.0f), prevSpecVMBResponsive(0.0f);
            float3 prevNormalVMB = currentNormal;
            float2 prevUVVMB;
            float prevRoughnessVMB = 0.0f, prevReflectionHitTVMB = c.gDenoisingRange, VMBReprojectionFound;
            {
                float3 virtualViewVector = normalize(currentViewVector) * float3(hitDistFocused);
                float3 prevVirtualWorldPos = prevWorldPos + virtualViewVector;
                float4 clip = mul(c.gWorldToClipPrev, float4(prevVirtualWorldPos, 1.0f));
                clip.x /= clip.w;
                clip.y /= clip.w;
                prevUVVMB = clip.xy() * float2(0.5f, -0.5f) + float2(0.5f, 0.5f);
                prevUVVMB = currentMaterialID == c.gCameraAttachedReflectionMaterialID ? prevUVSMB : prevUVVMB;
                float2 prevVirtualPixelPosFloat = prevUVVMB * c.gRectSizePrev;
                float2 fl = floor(prevVirtualPixelPosFloat - float2(0.5f));
                int2 bilinearOrigin((int)fl.x, (int)fl.y);
                float2 bilinearWeights = frac(prevVirtualPixelPosFloat - float2(0.5f));
                float2 bo = tofloat(bilinearOrigin);
                float2 gatherOrigin = (bo + float2(1.0f)) * c.gResourceSizeInvPrev;
                float3 cw = currentWorldPos - c.gCameraDelta.xyz();
                float4 thr = float4(disocclusionThreshold * (c.gOrthoMode == 0.0f ? currentLinearZ : 1.0f));
                thr *= IsInScreenBilinear(bo, c.gRectSizePrev);
                thr -= float4(NRD_EPS);
                float4 g = wzxy(gPrev_ViewZ.gather(gatherOrigin, 0));
                float4 prevViewZs(P.UnpackViewZ(g.x), P.UnpackViewZ(g.y), P.UnpackViewZ(g.z), P.UnpackViewZ(g.w));
                float4 prevMaterialIDs = wzxy(gPrev_MaterialID.gather(gatherOrigin, 0)) * float4(255.0f);
                auto tapValid = [&](int2 off, float z, float th) {
                    float3 pw = P.GetPreviousWorldPosFromPixelPos(bilinearOrigin + off, z);
                    float maxPlaneDistance = abs(dot(cw - pw, currentNormal));
                    return maxPlaneDistance > th ? 0.0f : 1.0f;
                };
                float4 bilinearTapsValid(tapValid(int2(0, 0), prevViewZs.x, thr.x), tapValid(int2(1, 0), prevViewZs.y, thr.y), tapValid(int2(0, 1), prevViewZs.z, thr.z),
                                         tapValid(int2(1, 1), prevViewZs.w, thr.w));
                for (int k = 0; k < 4; k++) bilinearTapsValid[k] *= float(P.CompareMaterials(currentMaterialID, prevMaterialIDs[k], c.gSpecMinMaterial));
                bool anyValid = bilinearTapsValid.x != 0.0f || bilinearTapsValid.y != 0.0f || bilinearTapsValid.z != 0.0f || bilinearTapsValid.w != 0.0f;
                bool allValid = bilinearTapsValid.x != 0.0f && bilinearTapsValid.y != 0.0f && bilinearTapsValid.z != 0.0f && bilinearTapsValid.w != 0.0f;
                if (anyValid)
                {
                    Filtering::Bilinear bil;
                    bil.origin = bo;
                    bil.weights = bilinearWeights;
                    float4 bcw = Filtering::GetBilinearCustomWeights(bil, bilinearTapsValid);
                    bool useBicubic = (SMBReprojectionFound == 2.0f) && allValid;
                    prevSpecVMB = max(BicubicCustom(prevVirtualPixelPosFloat, c.gResourceSizeInvPrev, bcw, useBicubic, gHistory_Spec), float4(0.0f));
                    prevSpecVMBResponsive = max(BicubicCustom(prevVirtualPixelPosFloat, c.gResourceSizeInvPrev, bcw, useBicubic, gHistory_SpecFast), float4(0.0f));
                    prevReflectionHitTVMB = max(0.001f, gPrev_SpecHitDist.sampleLinear(prevUVVMB * P.ResolutionScalePrev()).x);
                    float4 pnr = Pass::UnpackPrevNormalRoughness(gPrev_Normal_Roughness.sampleLinear(prevUVVMB * P.ResolutionScalePrev()));
                    prevNormalVMB = Geometry::RotateVector(c.gWorldPrevToWorld, pnr.xyz());
                    prevRoughnessVMB = pnr.w;
                }
                VMBReprojectionFound = allValid ? 1.0f : 0.0f;
            }

            float4 D = ImportanceSampling::GetSpecularDominantDirection(currentNormal, V, currentRoughnessModified);
            float virtualHistoryAmount = VMBReprojectionFound * D.w;
            virtualHistoryAmount *= c.gOrthoMode == 0.0f ? 1.0f : 0.75f;
            virtualHistoryAmount *= float(dot(prevNormalVMB, currentNormalAveraged) > 0.0f);

            float2 uvDiff = prevUVVMB - prevUVSMB;
            float uvDiffLengthInPixels = length(uvDiff * rectSize);
            float tanCurvature = abs(curvature * pixelSize);
            tanCurvature *= max(uvDiffLengthInPixels / max(NoV, 0.01f), 1.0f);
            float curvatureAngle = atan(tanCurvature);

            float lobeHalfAngle = max(atan(Pass::GetSpecLobeTanHalfAngle(currentRoughnessModified)), RELAX_NORMAL_ULP);
            float normalWeight = Pass::GetEncodingAwareNormalWeight(currentNormal, prevNormalVMB, lobeHalfAngle, curvatureAngle, RELAX_NORMAL_ULP, true);
            virtualHistoryAmount *= lerp(1.0f - saturate(uvDiffLengthInPixels), 1.0f, normalWeight);

            float2 rrp = Pass::GetRelaxedRoughnessWeightParams(currentRoughness * currentRoughness, c.gRoughnessFraction);
            float virtualRoughnessWeight = ComputeWeight(prevRoughnessVMB * prevRoughnessVMB, rrp.x, rrp.y);
            virtualRoughnessWeight = lerp(1.0f - saturate(uvDiffLengthInPixels), 1.0f, virtualRoughnessWeight);
            virtualHistoryAmount *= c.gOrthoMode == 0.0f ? virtualRoughnessWeight : 1.0f;
            float specVMBConfidence = virtualRoughnessWeight * 0.9f + 0.1f;

            uvDiff *= float2(Math::Rsqrt(Math::LengthSquared(uvDiff)));
            uvDiff /= c.gRectSizePrev;
            uvDiff *= float2(saturate(uvDiffLengthInPixels / 0.1f) + uvDiffLengthInPixels / 2.0f);
            float2 backUV1 = prevUVVMB + uvDiff * float2(1.0f), backUV2 = prevUVVMB + uvDiff * float2(2.0f);
            float4 back1 = Pass::UnpackPrevNormalRoughness(gPrev_Normal_Roughness.sampleLinear(backUV1 * P.ResolutionScalePrev()));
            float4 back2 = Pass::UnpackPrevNormalRoughness(gPrev_Normal_Roughness.sampleLinear(backUV2 * P.ResolutionScalePrev()));
            back1.set_xyz(Geometry::RotateVector(c.gWorldPrevToWorld, back1.xyz()));
            back2.set_xyz(Geometry::RotateVector(c.gWorldPrevToWorld, back2.xyz()));
            float prevPrevNormalWeight = IsInScreenNearest(backUV1) != 0.0f ? Pass::GetEncodingAwareNormalWeight(prevNormalVMB, back1.xyz(), lobeHalfAngle, curvatureAngle * 2.0f, RELAX_NORMAL_ULP, true) : 1.0f;
            prevPrevNormalWeight *= IsInScreenNearest(backUV2) != 0.0f ? Pass::GetEncodingAwareNormalWeight(prevNormalVMB, back2.xyz(), lobeHalfAngle, curvatureAngle * 3.0f, RELAX_NORMAL_ULP, true) : 1.0f;
            virtualHistoryAmount *= 0.33f + 0.67f * prevPrevNormalWeight;
            specVMBConfidence *= 0.33f + 0.67f * prevPrevNormalWeight;
            float rw = ComputeWeight(back1.w * back1.w, rrp.x, rrp.y);
            rw *= ComputeWeight(back2.w * back2.w, rrp.x, rrp.y);
            virtualHistoryAmount *= c.gOrthoMode == 0.0f ? rw * 0.9f + 0.1f : 1.0f;

            float SMC = GetSpecMagicCurve(currentRoughnessModified);
            float hitDistC = lerp(specularIllumination.w, prevReflectionHitTSMB, SMC);
            float hitDist1 = Pass::ApplyThinLensEquation(hitDistC, curvature);
            float hitDist2 = Pass::ApplyThinLensEquation(prevReflectionHitTVMB, curvature);
            float maxDist = max(hitDist1, hitDist2);
            float dHitT = abs(hitDist1 - hitDist2);
            float dHitTMultiplier = lerp(20.0f, 0.0f, SMC);
            float virtualHistoryHitDistConfidence = 1.0f - saturate(dHitTMultiplier * dHitT / (currentLinearZ + maxDist));
            virtualHistoryHitDistConfidence = lerp(virtualHistoryHitDistConfidence, 1.0f, SMC);

            float3 virtualWorldPos = Pass::GetXvirtual(hitDist, curvature, currentWorldPos, prevWorldPos, currentNormal, V, currentRoughness);
            float virtualWorldPosLength = length(virtualWorldPos);
            float hitDistForTrackingPrev = prevSpecVMBResponsive.w;
            float3 prevVirtualWorldPos2 = Pass::GetXvirtual(hitDistForTrackingPrev, curvature, currentWorldPos, prevWorldPos, currentNormal, V, currentRoughness);
            float virtualWorldPosLengthPrev = length(prevVirtualWorldPos2);
            float2 prevUVVMBTest = Geometry::GetScreenUv(c.gWorldToClipPrev, prevVirtualWorldPos2, false);
            prevUVVMBTest = currentMaterialID == c.gCameraAttachedReflectionMaterialID ? prevUVSMB : prevUVVMBTest;
            float lobeTanHalfAngle = Pass::GetSpecLobeTanHalfAngle(currentRoughness, 0.6f);
            lobeTanHalfAngle = max(lobeTanHalfAngle, 0.5f * c.gRectSizeInv.x);
            float unproj1 = min(hitDist, hitDistForTrackingPrev) / PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, 1.0f, max(virtualWorldPosLength, virtualWorldPosLengthPrev));
            float lobeRadiusInPixels = lobeTanHalfAngle * unproj1;
            float deltaParallaxInPixels = length((prevUVVMBTest - prevUVVMB) * rectSize);
            virtualHistoryHitDistConfidence *= Math::SmoothStep(lobeRadiusInPixels + 0.25f, 0.0f, deltaParallaxInPixels);

            float specSMBConfidence = (SMBReprojectionFound > 0.0f ? 1.0f : 0.0f) * Pass::GetEncodingAwareNormalWeight(V, Vprev, lobeHalfAngle * NoV / c.gFramerateScale, 0.0f, 0.0f, false);
            float specSMBAlpha = 1.0f - specSMBConfidence;
            float specSMBResponsiveAlpha = 1.0f - specSMBConfidence;
            specSMBAlpha = max(specSMBAlpha, 1.0f / (1.0f + specHistoryFrames));
            specSMBResponsiveAlpha = max(specSMBAlpha, 1.0f / (1.0f + specHistoryResponsiveFrames));
            bool specHasData = true;
            if (c.gSpecCheckerboard != 2) specHasData = checkerboard == c.gSpecCheckerboard;
            if (!specHasData && smbParallaxInPixelsMax < 0.5f)
            {
                specSMBAlpha *= 1.0f - c.gCheckerboardResolveAccumSpeed * (SMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
                specSMBResponsiveAlpha *= 1.0f - c.gCheckerboardResolveAccumSpeed * (SMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
            }
            float4 accSMB;
            accSMB.set_xyz(lerp(prevSpecSMB.xyz(), specularIllumination.xyz(), specSMBAlpha));
            accSMB.w = lerp(prevReflectionHitTSMB, specularIllumination.w, max(specSMBAlpha, 0.1f));
            float accM2SMB = lerp(prevSpecSMB.w, specular2ndMoment, specSMBAlpha);
            float3 accSMBResponsive = lerp(prevSpecSMBResponsive, specularIllumination.xyz(), specSMBResponsiveAlpha);

            float specVMBAlpha = 1.0f - specVMBConfidence;
            float specVMBResponsiveAlpha = 1.0f - specVMBConfidence * virtualHistoryHitDistConfidence;
            float specVMBHitTAlpha = specVMBResponsiveAlpha;
            specVMBAlpha = max(specVMBAlpha, 1.0f / (1.0f + specHistoryFrames));
            specVMBResponsiveAlpha = max(specVMBResponsiveAlpha, 1.0f / (1.0f + specHistoryResponsiveFrames));
            specVMBHitTAlpha = max(specVMBHitTAlpha, 1.0f / (1.0f + specHistoryFrames));
            if (!specHasData && smbParallaxInPixelsMax < 0.5f)
            {
                float k = 1.0f - c.gCheckerboardResolveAccumSpeed * (VMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
                specVMBAlpha *= k;
                specVMBResponsiveAlpha *= k;
                specVMBHitTAlpha *= k;
            }
            float4 accVMB;
            accVMB.set_xyz(lerp(prevSpecVMB.xyz(), specularIllumination.xyz(), specVMBAlpha));
            accVMB.w = lerp(prevReflectionHitTVMB, specularIllumination.w, max(specVMBHitTAlpha, 0.1f));
            float accM2VMB = lerp(prevSpecVMB.w, specular2ndMoment, specVMBAlpha);
            float3 accVMBResponsive = lerp(prevSpecVMBResponsive.xyz(), specularIllumination.xyz(), specVMBResponsiveAlpha);

            virtualHistoryAmount *= saturate(specVMBConfidence / (specSMBConfidence + NRD_EPS));
            float accumulatedReflectionHitT = lerp(accSMB.w, accVMB.w, virtualHistoryAmount);
            float3 accSpec = lerp(accSMB.xyz(), accVMB.xyz(), virtualHistoryAmount);
            float3 accSpecResponsive = lerp(accSMBResponsive, accVMBResponsive, virtualHistoryAmount);
            float accSpec2ndMoment = lerp(accM2SMB, accM2VMB, virtualHistoryAmount);
            float specularHistoryConfidence = lerp(specSMBConfidence, specVMBConfidence, virtualHistoryAmount);
            if (accSpec2ndMoment == 0.0f) accSpec2ndMoment = c.gSpecVarianceBoost * (1.0f - specularHistoryConfidence);

            gOut_Spec.store(pixelPos, float4(accSpec, accSpec2ndMoment));
            gOut_SpecFast.store(pixelPos, float4(accSpecResponsive, hitDist));
            gOut_SpecHitDist.store(pixelPos, accumulatedReflectionHitT);
            gOut_SpecReprojectionConfidence.store(pixelPos, specularHistoryConfidence);
        }
}

void HistoryFix(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_Spec = t[1], &gIn_Diff = t[2], &gIn_HistoryLength = t[3], &gIn_Normal_Roughness = t[4], &gIn_ViewZ = t[5];
    Tex &gOut_Spec = t[6], &gOut_Diff = t[7];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 8; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            float centerViewZ = P.UnpackViewZ(gIn_ViewZ.load(pixelPos).x);
            float historyLength = 255.0f * gIn_HistoryLength.load(pixelPos).x;
            if (centerViewZ > c.gDenoisingRange || historyLength > c.gHistoryFixFrameNum || c.gHistoryFixFrameNum == 1.0f) continue;

            float centerMaterialID;
            float4 cnr = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos), centerMaterialID);
            float3 centerNormal = cnr.xyz();
            float centerRoughness = cnr.w;
            float3 centerWorldPos = P.GetCurrentWorldPosFromPixelPos(pixelPos, centerViewZ);
            float3 centerV = -normalize(centerWorldPos);
            float depthThreshold = c.gDepthThreshold * (c.gOrthoMode == 0.0f ? centerViewZ : 1.0f);
            float4 diffSum = gIn_Diff.load(pixelPos), specSum = gIn_Spec.load(pixelPos);
            float diffuseWSum = 1.0f, specularWSum = 1.0f;
            float2 specularNormalWeightParams = Pass::GetNormalWeightParams_ATrous(centerRoughness, 5.0f, 1.0f, 0.0f, c.gLobeAngleFraction, c.gSpecLobeAngleSlack);
            float r = c.gHistoryFixBasePixelStride / (1.0f + historyLength);
            r = floor(r + 0.5f);
            for (int j = -2; j <= 2; j++)
                for (int i = -2; i <= 2; i++)
                {
                    int dx = (int)(float(i) * r), dy = (int)(float(j) * r);
                    int2 sp = pixelPos + int2(dx, dy);
                    bool isInside = sp.x >= 0 && sp.y >= 0 && sp.x < c.gRectSize[0] && sp.y < c.gRectSize[1];
                    if (i == 0 && j == 0) continue;
                    float sampleMaterialID;
                    float3 sampleNormal = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(sp), sampleMaterialID).xyz();
                    float sampleViewZ = P.UnpackViewZ(gIn_ViewZ.load(sp).x);
                    float3 sampleWorldPos = P.GetCurrentWorldPosFromPixelPos(sp, sampleViewZ);
                    float geometryWeight = Pass::GetPlaneDistanceWeight_Atrous(centerWorldPos, centerNormal, sampleWorldPos, depthThreshold);
                    float diffuseW = geometryWeight;
                    diffuseW *= pow(max(0.01f, dot(centerNormal, sampleNormal)), max(c.gHistoryFixEdgeStoppingNormalPower, 0.01f));
                    diffuseW = isInside ? diffuseW : 0.0f;
                    diffuseW *= float(P.CompareMaterials(sampleMaterialID, centerMaterialID, c.gDiffMinMaterial));
                    if (diffuseW > 1e-4f)
                    {
                        diffSum += gIn_Diff.load(sp) * float4(diffuseW);
                        diffuseWSum += diffuseW;
                    }
                    float3 sampleV = -normalize(sampleWorldPos + float3(c.gRoughnessEdgeStoppingRelaxation) * centerWorldPos);
                    float specularW = geometryWeight;
                    specularW *= Pass::GetSpecularNormalWeight_ATrous(specularNormalWeightParams, centerNormal, sampleNormal, centerV, sampleV);
                    specularW = isInside ? specularW : 0.0f;
                    specularW *= float(P.CompareMaterials(sampleMaterialID, centerMaterialID, c.gSpecMinMaterial));
                    if (specularW > 1e-4f)
                    {
                        specSum += gIn_Spec.load(sp) * float4(specularW);
                        specularWSum += specularW;
                    }
                }
            gOut_Diff.store(pixelPos, diffSum / float4(diffuseWSum));
            gOut_Spec.store(pixelPos, specSum / float4(specularWSum));
        }
}

void HistoryClamping(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_ViewZ = t[1], &gIn_SpecNoisy = t[2], &gIn_DiffNoisy = t[3], &gIn_Spec = t[4], &gIn_Diff = t[5], &gIn_SpecFast = t[6], &gIn_DiffFast = t[7],
              &gIn_HistoryLength = t[8];
    Tex &gOut_Spec = t[9], &gOut_Diff = t[10], &gOut_SpecFast = t[11], &gOut_DiffFast = t[12], &gOut_HistoryLength = t[13];
    const int2 rectMax(c.gRectSize[0] - 1, c.gRectSize[1] - 1);

    auto clampSignal = [&](bool isSpec, int2 pixelPos, float historyLength, const Tex& inNoisy, const Tex& inSlow, const Tex& inFast, Tex& outSlow, Tex& outFast) {
        int x = pixelPos.x, y = pixelPos.y;
        float3 m1(0.0f), m2(0.0f), noisyM1(0.0f);
        float noisyM2 = 0.0f, sum = 0.0f;
        for (int dx = -2; dx <= 2; dx++)
            for (int dy = -2; dy <= 2; dy++)
            {
                int2 p = clamp(int2(x + dx, y + dy), int2(0), rectMax);
                float w = float(gIn_ViewZ.load(p).x < c.gDenoisingRange);
                if (w != 0.0f)
                {
                    float3 sy = RgbToYCoCg(inFast.load(p).xyz());
                    m1 += sy;
                    m2 += sy * sy;
                    float3 n = inNoisy.load(p).xyz();
                    float l = Color::Luminance(n);
                    noisyM1 += n;
                    noisyM2 += l * l;
                    sum += w;
                }
            }
        m1 /= float3(sum);
        m2 /= float3(sum);
        noisyM1 /= float3(sum);
        noisyM2 /= sum;
        float3 sigma = sqrt(max(float3(0.0f), m2 - m1 * m1));
        float3 cmin = m1 - float3(c.gColorBoxSigmaScale) * sigma, cmax = m1 + float3(c.gColorBoxSigmaScale) * sigma;
        float4 fastCenter = inFast.load(pixelPos);
        float4 responsiveCenterYCoCg(RgbToYCoCg(fastCenter.xyz()), fastCenter.w);
        cmin = min(cmin, responsiveCenterYCoCg.xyz());
        cmax = max(cmax, responsiveCenterYCoCg.xyz());

        float4 slow = inSlow.load(pixelPos);
        float3 slowYCoCg = RgbToYCoCg(slow.xyz());
        float3 clampedYCoCg = slowYCoCg;
        float maxFast = isSpec ? c.gSpecMaxFastAccumulatedFrameNum : c.gDiffMaxFastAccumulatedFrameNum, maxSlow = isSpec ? c.gSpecMaxAccumulatedFrameNum : c.gDiffMaxAccumulatedFrameNum;
        if (maxFast < maxSlow) clampedYCoCg = min(max(slowYCoCg, cmin), cmax);
        float3 clamped = YCoCgToRgb(clampedYCoCg);

        float4 outS(clamped, slow.w);
        float3 responsiveCenter = YCoCgToRgb(responsiveCenterYCoCg.xyz());
        float4 outR(responsiveCenter, isSpec ? responsiveCenterYCoCg.w : 0.0f);
        if (historyLength <= c.gHistoryFixFrameNum)
        {
            if (isSpec) outS = outR;
            else outS.set_xyz(outR.xyz());
        }
        float clampingFactor = (clampedYCoCg.x - slowYCoCg.x) == 0.0f ? 0.0f : saturate((clampedYCoCg.x - slowYCoCg.x) / (responsiveCenterYCoCg.x - slowYCoCg.x));
        if (historyLength <= c.gHistoryFixFrameNum) clampingFactor = 1.0f;

        float historyDifferenceL = (isSpec ? 0.33f : 1.0f) * RELAX_ANTILAG_ACCELERATION_AMOUNT_SCALE * c.gHistoryAccelerationAmount * Color::Luminance(abs3(responsiveCenter - slow.xyz()));
        historyDifferenceL *= clampingFactor;
        if (historyLength <= c.gHistoryFixFrameNum) historyDifferenceL = 0.0f;

        float3 distToNoisy = noisyM1 - responsiveCenter;
        float distToNoisyL = Color::Luminance(abs3(distToNoisy));
        float3 accel = distToNoisyL == 0.0f ? float3(0.0f) : distToNoisy * float3(historyDifferenceL) / float3(distToNoisyL);
        float accelL = Color::Luminance(abs3(accel));
        float accelRatio = accelL == 0.0f ? 0.0f : distToNoisyL / accelL;
        if (accelRatio < 1.0f) accel *= float3(accelRatio);
        if (accelRatio <= 0.0f) accel = float3(0.0f);
        outS.set_xyz(outS.xyz() + accel);
        outR.set_xyz(outR.xyz() + accel);

        float slowL = Color::Luminance(slow.xyz());
        float noisyL = Color::Luminance(noisyM1);
        float temporalSigma = c.gHistoryResetTemporalSigmaScale * sqrt(max(0.0f, noisyM2 - noisyL * noisyL));
        float spatialSigma = c.gHistoryResetSpatialSigmaScale * sigma.x;
        float resetAmount = (isSpec ? 0.5f : 1.0f) * c.gHistoryResetAmount * max(0.0f, abs(slowL - noisyL) - spatialSigma - temporalSigma) /
                            (1.0e-6f + max(slowL, noisyL) + spatialSigma + temporalSigma);
        resetAmount = saturate(resetAmount);
        float3 noisyCenter = inNoisy.load(pixelPos).xyz();
        outS.set_xyz(lerp(outS.xyz(), noisyCenter, resetAmount));
        outR.set_xyz(lerp(outR.xyz(), noisyCenter, resetAmount));

        float outL = Color::Luminance(outS.xyz());
        outS.w += outL * outL - slowL * slowL;
        outS.w = max(0.0f, outS.w);
        outSlow.store(pixelPos, outS);
        outFast.store(pixelPos, outR);
    };

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 8; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            if (!(gIn_ViewZ.load(pixelPos).x < c.gDenoisingRange)) continue; // s_*Noisy_IsValid.w == 0
            float historyLength = 255.0f * gIn_HistoryLength.load(pixelPos).x;
            clampSignal(true, pixelPos, historyLength, gIn_SpecNoisy, gIn_Spec, gIn_SpecFast, gOut_Spec, gOut_SpecFast);
            clampSignal(false, pixelPos, historyLength, gIn_DiffNoisy, gIn_Diff, gIn_DiffFast, gOut_Diff, gOut_DiffFast);
            gOut_HistoryLength.store(pixelPos, historyLength / 255.0f);
        }
}

// RELAX_Copy.hlsli:11-24
void Copy(const Pass&, Tex* t, int gridW, int gridH)
{
    const Tex &gIn_Spec = t[0], &gIn_Diff = t[1];
    Tex &gOut_Spec = t[2], &gOut_Diff = t[3];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 8; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            gOut_Spec.store(int2(x, y), gIn_Spec.load(x, y));
            gOut_Diff.store(int2(x, y), gIn_Diff.load(x, y));
        }
}

// RELAX_AntiFirefly.hlsli:11-222: cross-bilateral rank-conditioned rank-selection over the 3x3 neighbourhood
void AntiFirefly(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_Spec = t[1], &gIn_Diff = t[2], &gIn_Normal_Roughness = t[3], &gIn_ViewZ = t[4];
    Tex &gOut_Spec = t[5], &gOut_Diff = t[6];
    const int2 rectMax(c.gRectSize[0] - 1, c.gRectSize[1] - 1);

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 8; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            float centerViewZ = P.UnpackViewZ(gIn_ViewZ.load(pixelPos).x);
            if (centerViewZ > c.gDenoisingRange) continue;

            // the shared-memory tile of the reference holds clamped texels (Preload, :22-38)
            auto material = [&](int2 p) {
                float m;
                NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(clamp(p, int2(0), rectMax)), m);
                return m;
            };
            float centerMaterialID = material(pixelPos);
            auto rcrs = [&](const Tex& in, float minMaterial) {
                float4 center = in.load(pixelPos);
                float centerLuminance = Color::Luminance(center.xyz());
                float maxLuminance = -1.0f, minLuminance = 1.0e6f;
                int2 maxCoords = pixelPos, minCoords = pixelPos;
                for (int yy = -1; yy <= 1; yy++)
                    for (int xx = -1; xx <= 1; xx++)
                    {
                        int2 p = pixelPos + int2(xx, yy);
                        if (xx == 0 && yy == 0) continue;
                        if (p.x < 0 || p.y < 0 || p.x >= c.gRectSize[0] || p.y >= c.gRectSize[1]) continue;
                        float luminance = Color::Luminance(in.load(p).xyz());
                        if (P.CompareMaterials(material(p), centerMaterialID, minMaterial))
                        {
                            if (luminance > maxLuminance) { maxLuminance = luminance; maxCoords = p; }
                            if (luminance < minLuminance) { minLuminance = luminance; minCoords = p; }
                        }
                    }
                int2 coords = pixelPos;
                if (centerLuminance > maxLuminance) coords = maxCoords;
                if (centerLuminance < minLuminance) coords = minCoords;
                return float4(in.load(coords).xyz(), center.w);
            };
            gOut_Spec.store(pixelPos, rcrs(gIn_Spec, c.gSpecMinMaterial));
            gOut_Diff.store(pixelPos, rcrs(gIn_Diff, c.gDiffMinMaterial));
        }
}

void AtrousSmem(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_Spec = t[1], &gIn_Diff = t[2], &gIn_HistoryLength = t[3], &gIn_SpecReprojectionConfidence = t[4], &gIn_Normal_Roughness = t[5], &gIn_ViewZ = t[6];
    const Tex &gIn_SpecConfidence = t[7], &gIn_DiffConfidence = t[8];
    Tex &gOut_Spec = t[9], &gOut_Diff = t[10], &gOut_NormalRoughness = t[11], &gOut_MaterialID = t[12], &gOut_ViewZ = t[13];
    const int2 rectMax(c.gRectSize[0] - 1, c.gRectSize[1] - 1);
    const float gk[2] = {0.44198f, 0.27901f};

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 8; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            auto sSpec = [&](int i, int j) { return gIn_Spec.load(clamp(int2(x + i, y + j), int2(0), rectMax)); };
            auto sDiff = [&](int i, int j) { return gIn_Diff.load(clamp(int2(x + i, y + j), int2(0), rectMax)); };
            auto sNormalRoughness = [&](int i, int j, float& mat) { return NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(clamp(int2(x + i, y + j), int2(0), rectMax)), mat); };
            auto sWorldPos = [&](int i, int j) {
                int2 p = clamp(int2(x + i, y + j), int2(0), rectMax);
                return P.GetCurrentWorldPosFromPixelPos(p, P.UnpackViewZ(gIn_ViewZ.load(p).x));
            };

            float viewZpacked = gIn_ViewZ.load(pixelPos).x;
            gOut_ViewZ.store(pixelPos, viewZpacked);
            float centerMaterialID = 0.0f;
            float4 normalRoughness(0.0f);
            float3 centerWorldPos(0.0f);
            // shared memory is only filled for non-sky tiles; for sky tiles the reference reads uninitialised shared memory and then
            // overwrites normal/roughness with 1/255 (all their pixels are out of range), material id stays "whatever": here 0
            if (isSky == 0.0f)
            {
                normalRoughness = sNormalRoughness(0, 0, centerMaterialID);
                centerWorldPos = sWorldPos(0, 0);
            }
            float centerViewZ = P.UnpackViewZ(viewZpacked);
            if (centerViewZ > c.gDenoisingRange) normalRoughness = float4(1.0f / 255.0f);
            gOut_NormalRoughness.store(pixelPos, Pass::PackPrevNormalRoughness(normalRoughness));
            gOut_MaterialID.store(pixelPos, centerMaterialID / 255.0f);

            if (isSky != 0.0f || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            if (centerViewZ > c.gDenoisingRange) continue;
            float3 centerNormal = normalRoughness.xyz();
            float centerRoughness = normalRoughness.w;
            float historyLength = 255.0f * gIn_HistoryLength.load(pixelPos).x;

            if (historyLength >= c.gHistoryThreshold)
            {
                // 3x3 gaussian of the variance (computeVariance :30-83)
                float4 specularSum(0.0f), diffuseSum(0.0f);
                const float kernel[2][2] = {{1.0f / 4.0f, 1.0f / 8.0f}, {1.0f / 8.0f, 1.0f / 16.0f}};
                for (int dx = -1; dx <= 1; dx++)
                    for (int dy = -1; dy <= 1; dy++)
                    {
                        float k = kernel[std::abs(dx)][std::abs(dy)];
                        specularSum += sSpec(dx, dy) * float4(k);
                        diffuseSum += sDiff(dx, dy) * float4(k);
                    }
                float s1 = Color::Luminance(specularSum.xyz()), d1 = Color::Luminance(diffuseSum.xyz());
                float centerSpecularVar = max(0.0f, specularSum.w - s1 * s1), centerDiffuseVar = max(0.0f, diffuseSum.w - d1 * d1);

                float diffuseLobeAngleFraction = c.gLobeAngleFraction;
                float centerSpecularLuminance = Color::Luminance(sSpec(0, 0).xyz());
                float specularPhiLIlluminationInv = 1.0f / max(1.0e-4f, c.gSpecPhiLuminance * sqrt(centerSpecularVar));
                float2 roughnessWeightParams = Pass::GetRoughnessWeightParams(centerRoughness, c.gRoughnessFraction);
                float specularReprojectionConfidence = gIn_SpecReprojectionConfidence.load(pixelPos).x;
                float specularLuminanceWeightRelaxation = lerp(1.0f, specularReprojectionConfidence, c.gLuminanceEdgeStoppingRelaxation);
                // confidence-driven relaxation (:189-201, :226-238)
                float diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = diffuseLobeAngleFraction;
                float specularLobeAngleFraction = c.gLobeAngleFraction;
                if (c.gHasHistoryConfidence)
                {
                    float specConfidenceDrivenRelaxation = saturate(c.gConfidenceDrivenRelaxationMultiplier * (1.0f - gIn_SpecConfidence.load(pixelPos).x));
                    float r = saturate(specConfidenceDrivenRelaxation * c.gConfidenceDrivenNormalEdgeStoppingRelaxation);
                    diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = lerp(diffuseLobeAngleFraction, 1.0f, r);
                    specularLobeAngleFraction = lerp(specularLobeAngleFraction, 1.0f, r);
                    r = saturate(specConfidenceDrivenRelaxation * c.gConfidenceDrivenLuminanceEdgeStoppingRelaxation);
                    specularLuminanceWeightRelaxation *= 1.0f - r;
                }
                float specularNormalWeightParamSimplified = Pass::GetNormalWeightParam2(1.0f, diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight);
                float2 specularNormalWeightParams = Pass::GetNormalWeightParams_ATrous(centerRoughness, historyLength, specularReprojectionConfidence, c.gNormalEdgeStoppingRelaxation,
                                                                                       specularLobeAngleFraction, c.gSpecLobeAngleSlack);
                float sumWSpecular = 0.0f, sumWDiffuse = 0.0f;
                float4 sumSpecular(0.0f), sumDiffuse(0.0f);
                float3 centerV = -normalize(centerWorldPos);
                float centerDiffuseLuminance = Color::Luminance(sDiff(0, 0).xyz());
                float diffusePhiLIlluminationInv = 1.0f / max(1.0e-4f, c.gDiffPhiLuminance * sqrt(centerDiffuseVar));
                float diffuseLuminanceWeightRelaxation = 1.0f;
                if (c.gHasHistoryConfidence)
                {
                    float diffConfidenceDrivenRelaxation = saturate(c.gConfidenceDrivenRelaxationMultiplier * (1.0f - gIn_DiffConfidence.load(pixelPos).x));
                    float r = saturate(diffConfidenceDrivenRelaxation * c.gConfidenceDrivenNormalEdgeStoppingRelaxation);
                    diffuseLobeAngleFraction = lerp(diffuseLobeAngleFraction, 1.0f, r);
                    r = saturate(diffConfidenceDrivenRelaxation * c.gConfidenceDrivenLuminanceEdgeStoppingRelaxation);
                    diffuseLuminanceWeightRelaxation = 1.0f - r;
                }
                float diffuseNormalWeightParam = Pass::GetNormalWeightParam2(1.0f, diffuseLobeAngleFraction);
                float depthThreshold = c.gDepthThreshold * (c.gOrthoMode == 0.0f ? centerViewZ : 1.0f);

                for (int cx = -1; cx <= 1; cx++)
                    for (int cy = -1; cy <= 1; cy++)
                    {
                        int2 p = pixelPos + int2(cx, cy);
                        bool isCenter = cx == 0 && cy == 0;
                        bool isInside = p.x >= 0 && p.y >= 0 && p.x < c.gRectSize[0] && p.y < c.gRectSize[1];
                        float kernelW = isInside ? gk[std::abs(cx)] * gk[std::abs(cy)] : 0.0f;
                        float sampleMaterialID;
                        float4 snr = sNormalRoughness(cx, cy, sampleMaterialID);
                        float3 sampleNormal = snr.xyz();
                        float sampleRoughness = snr.w;
                        float3 sampleWorldPos = sWorldPos(cx, cy);
                        float geometryW = Pass::GetPlaneDistanceWeight_Atrous(centerWorldPos, centerNormal, sampleWorldPos, depthThreshold);
                        geometryW *= kernelW;

                        float angles = Math::AcosApprox(dot(centerNormal, sampleNormal));
                        float3 sampleV = -normalize(sampleWorldPos + float3(c.gRoughnessEdgeStoppingRelaxation) * centerWorldPos);
                        float normalWSpecularSimplified = ComputeWeight(angles, specularNormalWeightParamSimplified, 0.0f);
                        float normalWSpecular = Pass::GetSpecularNormalWeight_ATrous(specularNormalWeightParams, centerNormal, sampleNormal, centerV, sampleV);
                        float roughnessWSpecular = ComputeWeight(sampleRoughness, roughnessWeightParams.x, roughnessWeightParams.y);
                        float4 ss = sSpec(cx, cy);
                        float specularLuminanceW = abs(centerSpecularLuminance - Color::Luminance(ss.xyz())) * specularPhiLIlluminationInv;
                        specularLuminanceW = min(c.gSpecMaxLuminanceRelativeDifference, specularLuminanceW);
                        specularLuminanceW *= specularLuminanceWeightRelaxation;
                        float wSpecular = geometryW * exp(-specularLuminanceW);
                        wSpecular *= c.gRoughnessEdgeStoppingEnabled ? (normalWSpecular * roughnessWSpecular) : normalWSpecularSimplified;
                        wSpecular = isCenter ? kernelW : wSpecular;
                        wSpecular *= float(P.CompareMaterials(sampleMaterialID, centerMaterialID, c.gSpecMinMaterial));
                        sumWSpecular += wSpecular;
                        sumSpecular += float4(wSpecular) * ss;

                        float normalWDiffuse = ComputeWeight(angles, diffuseNormalWeightParam, 0.0f);
                        float4 sd = sDiff(cx, cy);
                        float diffuseLuminanceW = abs(centerDiffuseLuminance - Color::Luminance(sd.xyz())) * diffusePhiLIlluminationInv;
                        diffuseLuminanceW = min(c.gDiffMaxLuminanceRelativeDifference, diffuseLuminanceW);
                        diffuseLuminanceW *= diffuseLuminanceWeightRelaxation;
                        float wDiffuse = geometryW * normalWDiffuse * exp(-diffuseLuminanceW);
                        wDiffuse = isCenter ? kernelW : wDiffuse;
                        wDiffuse *= float(P.CompareMaterials(sampleMaterialID, centerMaterialID, c.gDiffMinMaterial));
                        sumWDiffuse += wDiffuse;
                        sumDiffuse += float4(wDiffuse) * sd;
                    }
                sumWSpecular = max(sumWSpecular, 1e-6f);
                sumSpecular /= float4(sumWSpecular);
                float sp1 = Color::Luminance(sumSpecular.xyz());
                gOut_Spec.store(pixelPos, float4(sumSpecular.xyz(), max(0.0f, sumSpecular.w - sp1 * sp1)));
                sumWDiffuse = max(sumWDiffuse, 1e-6f);
                sumDiffuse /= float4(sumWDiffuse);
                float dp1 = Color::Luminance(sumDiffuse.xyz());
                gOut_Diff.store(pixelPos, float4(sumDiffuse.xyz(), max(0.0f, sumDiffuse.w - dp1 * dp1)));
            }
            else
            {
                float sumWS = 0.0f, sumS1 = 0.0f, sumS2 = 0.0f, sumWD = 0.0f, sumD1 = 0.0f, sumD2 = 0.0f;
                float3 sumS(0.0f), sumD(0.0f);
                float diffuseNormalWeightParam = Pass::GetNormalWeightParam2(1.0f, c.gLobeAngleFraction);
                for (int cx = -2; cx <= 2; cx++)
                    for (int cy = -2; cy <= 2; cy++)
                    {
                        float sampleMaterialID;
                        float3 sampleNormal = sNormalRoughness(cx, cy, sampleMaterialID).xyz();
                        float angle = Math::AcosApprox(dot(centerNormal, sampleNormal));
                        float normalW = ComputeWeight(angle, diffuseNormalWeightParam, 0.0f);
                        float4 ss = sSpec(cx, cy);
                        float specularW = normalW * float(P.CompareMaterials(sampleMaterialID, centerMaterialID, c.gSpecMinMaterial));
                        sumWS += specularW;
                        sumS += ss.xyz() * float3(specularW);
                        sumS1 += Color::Luminance(ss.xyz()) * specularW;
                        sumS2 += ss.w * specularW;
                        float4 sd = sDiff(cx, cy);
                        float diffuseW = normalW * float(P.CompareMaterials(sampleMaterialID, centerMaterialID, c.gDiffMinMaterial));
                        sumWD += diffuseW;
                        sumD += sd.xyz() * float3(diffuseW);
                        sumD1 += Color::Luminance(sd.xyz()) * diffuseW;
                        sumD2 += sd.w * diffuseW;
                    }
                float boost = max(1.0f, 4.0f / (historyLength + 1.0f));
                sumWS = max(sumWS, 1e-6f);
                sumS /= float3(sumWS);
                sumS1 /= sumWS;
                sumS2 /= sumWS;
                gOut_Spec.store(pixelPos, float4(sumS, max(0.0f, sumS2 - sumS1 * sumS1) * boost));
                sumWD = max(sumWD, 1e-6f);
                sumD /= float3(sumWD);
                sumD1 /= sumWD;
                sumD2 /= sumWD;
                gOut_Diff.store(pixelPos, float4(sumD, max(0.0f, sumD2 - sumD1 * sumD1) * boost));
            }
        }
}

void Atrous(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_Spec = t[1], &gIn_Diff = t[2], &gIn_HistoryLength = t[3], &gIn_SpecReprojectionConfidence = t[4], &gIn_Normal_Roughness = t[5], &gIn_ViewZ = t[6];
    const Tex &gIn_SpecConfidence = t[7], &gIn_DiffConfidence = t[8];
    Tex &gOut_Spec = t[9], &gOut_Diff = t[10];
    const float gk[2] = {0.44198f, 0.27901f};
    const int step = (int)c.gStepSize;

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 16; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            float centerViewZ = P.UnpackViewZ(gIn_ViewZ.load(pixelPos).x);
            if (centerViewZ > c.gDenoisingRange) continue;
            float centerMaterialID;
            float4 cnr = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos), centerMaterialID);
            float3 centerNormal = cnr.xyz();
            float centerRoughness = cnr.w;
            float historyLength = 255.0f * gIn_HistoryLength.load(pixelPos).x;

            float diffuseLobeAngleFraction = c.gLobeAngleFraction / sqrt(float(c.gStepSize));
            diffuseLobeAngleFraction = lerp(0.99f, diffuseLobeAngleFraction, saturate(historyLength / 5.0f));

            float4 centerSpec = gIn_Spec.load(pixelPos);
            float centerSpecularLuminance = Color::Luminance(centerSpec.xyz());
            float specularPhiLIlluminationInv = 1.0f / max(1.0e-4f, c.gSpecPhiLuminance * sqrt(centerSpec.w));
            float2 roughnessWeightParams = Pass::GetRoughnessWeightParams(centerRoughness, c.gRoughnessFraction);
            float specularReprojectionConfidence = gIn_SpecReprojectionConfidence.load(pixelPos).x;
            float specularLuminanceWeightRelaxation = 1.0f;
            if (c.gStepSize <= 4) specularLuminanceWeightRelaxation = lerp(1.0f, specularReprojectionConfidence, c.gLuminanceEdgeStoppingRelaxation);
            // confidence-driven relaxation (:55-67, :95-106)
            float diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = diffuseLobeAngleFraction;
            float specularLobeAngleFraction = c.gLobeAngleFraction;
            if (c.gHasHistoryConfidence)
            {
                float specConfidenceDrivenRelaxation = saturate(c.gConfidenceDrivenRelaxationMultiplier * (1.0f - gIn_SpecConfidence.load(pixelPos).x));
                float r = saturate(specConfidenceDrivenRelaxation * c.gConfidenceDrivenNormalEdgeStoppingRelaxation);
                diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = lerp(diffuseLobeAngleFraction, 1.0f, r);
                specularLobeAngleFraction = lerp(specularLobeAngleFraction, 1.0f, r);
                r = saturate(specConfidenceDrivenRelaxation * c.gConfidenceDrivenLuminanceEdgeStoppingRelaxation);
                specularLuminanceWeightRelaxation *= 1.0f - r;
            }
            float specularNormalWeightParamSimplified = Pass::GetNormalWeightParam2(1.0f, diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight);
            float2 specularNormalWeightParams = Pass::GetNormalWeightParams_ATrous(centerRoughness, historyLength, specularReprojectionConfidence, c.gNormalEdgeStoppingRelaxation,
                                                                                   specularLobeAngleFraction, c.gSpecLobeAngleSlack);
            float sumWSpecular = 0.44198f * 0.44198f;
            float4 sumSpecular = centerSpec * float4(sumWSpecular, sumWSpecular, sumWSpecular, sumWSpecular * sumWSpecular);

            float4 centerDiff = gIn_Diff.load(pixelPos);
            float centerDiffuseLuminance = Color::Luminance(centerDiff.xyz());
            float diffusePhiLIlluminationInv = 1.0f / max(1.0e-4f, c.gDiffPhiLuminance * sqrt(centerDiff.w));
            float diffuseLuminanceWeightRelaxation = 1.0f;
            if (c.gHasHistoryConfidence)
            {
                float diffConfidenceDrivenRelaxation = saturate(c.gConfidenceDrivenRelaxationMultiplier * (1.0f - gIn_DiffConfidence.load(pixelPos).x));
                float r = saturate(diffConfidenceDrivenRelaxation * c.gConfidenceDrivenNormalEdgeStoppingRelaxation);
                diffuseLobeAngleFraction = lerp(diffuseLobeAngleFraction, 1.0f, r);
                r = saturate(diffConfidenceDrivenRelaxation * c.gConfidenceDrivenLuminanceEdgeStoppingRelaxation);
                diffuseLuminanceWeightRelaxation = 1.0f - r;
            }
            float diffuseNormalWeightParam = Pass::GetNormalWeightParam2(1.0f, diffuseLobeAngleFraction);
            float sumWDiffuse = 0.44198f * 0.44198f;
            float4 sumDiffuse = centerDiff * float4(sumWDiffuse, sumWDiffuse, sumWDiffuse, sumWDiffuse * sumWDiffuse);

            float3 centerWorldPos = P.GetCurrentWorldPosFromPixelPos(pixelPos, centerViewZ);
            float3 centerV = -normalize(centerWorldPos);
            float depthThreshold = c.gDepthThreshold * (c.gOrthoMode == 0.0f ? centerViewZ : 1.0f);

            int2 offset(0, 0);
            if (c.gStepSize > 4)
            {
                RngHash rng;
                rng.Initialize(pixelPos, c.gFrameIndex);
                float2 r = rng.GetFloat2();
                offset = int2((int)(float(c.gStepSize) * 0.5f * (r.x - 0.5f)), (int)(float(c.gStepSize) * 0.5f * (r.y - 0.5f)));
            }

            for (int yy = -1; yy <= 1; yy++)
                for (int xx = -1; xx <= 1; xx++)
                {
                    int2 p = pixelPos + offset + int2(xx, yy) * step;
                    if (xx == 0 && yy == 0) continue;
                    bool isInside = p.x >= 0 && p.y >= 0 && p.x < c.gRectSize[0] && p.y < c.gRectSize[1];
                    float kernelW = gk[std::abs(xx)] * gk[std::abs(yy)];
                    float sampleMaterialID;
                    float4 snr = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(p), sampleMaterialID);
                    float3 sampleNormal = snr.xyz();
                    float sampleRoughness = snr.w;
                    float sampleViewZ = P.UnpackViewZ(gIn_ViewZ.load(p).x);
                    float3 sampleWorldPos = P.GetCurrentWorldPosFromPixelPos(p, sampleViewZ);
                    float geometryW = Pass::GetPlaneDistanceWeight_Atrous(centerWorldPos, centerNormal, sampleWorldPos, depthThreshold);
                    geometryW *= kernelW;
                    geometryW *= float(isInside && sampleViewZ < c.gDenoisingRange);

                    float3 sampleV = -normalize(sampleWorldPos + float3(c.gRoughnessEdgeStoppingRelaxation) * centerWorldPos);
                    float angles = Math::AcosApprox(dot(centerNormal, sampleNormal));
                    float normalWSpecularSimplified = ComputeWeight(angles, specularNormalWeightParamSimplified, 0.0f);
                    float normalWSpecular = Pass::GetSpecularNormalWeight_ATrous(specularNormalWeightParams, centerNormal, sampleNormal, centerV, sampleV);
                    float roughnessWSpecular = ComputeWeight(sampleRoughness, roughnessWeightParams.x, roughnessWeightParams.y);
                    float wSpecular = geometryW * (c.gRoughnessEdgeStoppingEnabled ? (normalWSpecular * roughnessWSpecular) : normalWSpecularSimplified);
                    wSpecular *= float(P.CompareMaterials(sampleMaterialID, centerMaterialID, c.gSpecMinMaterial));
                    if (wSpecular > 1e-4f)
                    {
                        float4 ss = gIn_Spec.load(p);
                        float lw = abs(centerSpecularLuminance - Color::Luminance(ss.xyz())) * specularPhiLIlluminationInv;
                        lw = min(c.gSpecMaxLuminanceRelativeDifference, lw);
                        lw *= specularLuminanceWeightRelaxation;
                        wSpecular *= exp(-lw);
                        sumWSpecular += wSpecular;
                        sumSpecular += float4(wSpecular, wSpecular, wSpecular, wSpecular * wSpecular) * ss;
                    }
                    float normalWDiffuse = ComputeWeight(angles, diffuseNormalWeightParam, 0.0f);
                    float wDiffuse = geometryW * normalWDiffuse;
                    wDiffuse *= float(P.CompareMaterials(sampleMaterialID, centerMaterialID, c.gDiffMinMaterial));
                    if (wDiffuse > 1e-4f)
                    {
                        float4 sd = gIn_Diff.load(p);
                        float lw = abs(centerDiffuseLuminance - Color::Luminance(sd.xyz())) * diffusePhiLIlluminationInv;
                        lw = min(c.gDiffMaxLuminanceRelativeDifference, lw);
                        lw *= diffuseLuminanceWeightRelaxation;
                        wDiffuse *= exp(-lw);
                        sumWDiffuse += wDiffuse;
                        sumDiffuse += float4(wDiffuse, wDiffuse, wDiffuse, wDiffuse * wDiffuse) * sd;
                    }
                }
            gOut_Spec.store(pixelPos, sumSpecular / float4(sumWSpecular, sumWSpecular, sumWSpecular, sumWSpecular * sumWSpecular));
            gOut_Diff.store(pixelPos, sumDiffuse / float4(sumWDiffuse, sumWDiffuse, sumWDiffuse, sumWDiffuse * sumWDiffuse));
        }
}
// RELAX_SplitScreen.hlsli:11-50 (no SH; a checkerboarded input is stretched: pixel x shows packed column x >> 1): binding layout of the two-signal shader is viewZ, diff, spec | diff, spec
void SplitScreen(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_ViewZ = t[0], &gIn_Diff = t[1], &gIn_Spec = t[2];
    Tex &gOut_Diff = t[3], &gOut_Spec = t[4];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float2 pixelUv = (float2(float(x), float(y)) + float2(0.5f)) * c.gRectSizeInv;
            if (pixelUv.x > c.gSplitScreen || x >= c.gRectSize[0] || y >= c.gRectSize[1]) continue;
            float viewZ = P.UnpackViewZ(gIn_ViewZ.load(pixelPos).x);
            float keep = float(viewZ < c.gDenoisingRange);
            gOut_Diff.store(pixelPos, gIn_Diff.load(x >> (c.gDiffCheckerboard != 2 ? 1 : 0), y) * float4(keep));
            gOut_Spec.store(pixelPos, gIn_Spec.load(x >> (c.gSpecCheckerboard != 2 ? 1 : 0), y) * float4(keep));
        }
}

} // namespace

// RELAX_Diffuse_* / RELAX_Specular_* are the RELAX_DiffuseSpecular_* shaders compiled without the other signal (RELAX_DIFFUSE /
// RELAX_SPECULAR defines): the bindings of the absent signal do not exist (Source/Denoisers/Relax_Diffuse.hpp, Relax_Specular.hpp
// list the same passes minus those resources).  The passes above are written for both signals; a one-signal dispatch is expanded to
// that layout with NULL textures (0 x 0: every load reads 0, every store is dropped) in the places of the absent signal.  Per pass:
// the binding layout of the two-signal shader, c = common, s = specular only, d = diffuse only.
struct PassLayout
{
    const char* pass;
    const char* layout;
};
const PassLayout kLayouts[] = {
    {"HitDistReconstruction.cs", "csdccsd"},
    {"HitDistReconstruction_5x5.cs", "csdccsd"},
    {"PrePass.cs", "csdccsd"},
    {"TemporalAccumulation.cs", "csdcccsdsdccsccsdcsdsdscs"},
    {"HistoryFix.cs", "csdcccsd"},
    {"HistoryClamping.cs", "ccsdsdsdcsdsdc"},
    {"Copy.cs", "sdsd"},
    {"AntiFirefly.cs", "csdccsd"},
    {"AtrousSmem.cs", "csdcsccsdsdccc"},
    {"Atrous.cs", "csdcsccsdsd"},
    {"SplitScreen.cs", "cdsds"},
};

int relax_dispatch_impl(const char* shaderName, const void* constants, int constantsSize, Tex* tex, int texNum, int gridW, int gridH)
{
    if (constantsSize < 704) return -2;
    CB cb;
    memset(&cb, 0, sizeof(cb));
    memcpy(&cb, constants, constantsSize < (int)sizeof(CB) ? constantsSize : (int)sizeof(CB));
    Pass P(cb);
    if (!strcmp(shaderName, "RELAX_ClassifyTiles.cs"))
    {
        ClassifyTiles(P, tex, gridW, gridH);
        return 0;
    }
    if (strncmp(shaderName, "RELAX_", 6) != 0) return -1;
    const char* p = shaderName + 6;
    bool hasDiff = false, hasSpec = false;
    if (!strncmp(p, "DiffuseSpecular_", 16)) { hasDiff = hasSpec = true; p += 16; }
    else if (!strncmp(p, "Diffuse_", 8)) { hasDiff = true; p += 8; }
    else if (!strncmp(p, "Specular_", 9)) { hasSpec = true; p += 9; }
    else return -1;

    const PassLayout* layout = nullptr;
    for (const PassLayout& l : kLayouts)
        if (!strcmp(p, l.pass)) layout = &l;
    if (!layout) return -1;
    Tex t[32];
    int k = 0;
    const int n = (int)strlen(layout->layout);
    for (int i = 0; i < n; i++)
    {
        const char kind = layout->layout[i];
        const bool present = kind == 'c' || (kind == 's' && hasSpec) || (kind == 'd' && hasDiff);
        if (present)
        {
            if (k >= texNum) return -3;
            t[i] = tex[k++];
        }
    }
    if (k != texNum) return -3;

    if (!strcmp(p, "HitDistReconstruction.cs")) HitDistReconstruction(P, 1, t, gridW, gridH);
    else if (!strcmp(p, "HitDistReconstruction_5x5.cs")) HitDistReconstruction(P, 2, t, gridW, gridH);
    else if (!strcmp(p, "PrePass.cs")) PrePass(P, t, gridW, gridH);
    else if (!strcmp(p, "TemporalAccumulation.cs")) TemporalAccumulation(P, t, gridW, gridH, hasDiff, hasSpec);
    else if (!strcmp(p, "HistoryFix.cs")) HistoryFix(P, t, gridW, gridH);
    else if (!strcmp(p, "HistoryClamping.cs")) HistoryClamping(P, t, gridW, gridH);
    else if (!strcmp(p, "Copy.cs")) Copy(P, t, gridW, gridH);
    else if (!strcmp(p, "AntiFirefly.cs")) AntiFirefly(P, t, gridW, gridH);
    else if (!strcmp(p, "AtrousSmem.cs")) AtrousSmem(P, t, gridW, gridH);
    else if (!strcmp(p, "Atrous.cs")) Atrous(P, t, gridW, gridH);
    else if (!strcmp(p, "SplitScreen.cs")) SplitScreen(P, t, gridW, gridH);
    else return -1;
    return 0;
}
} // namespace hlsl

int oracle_relax_dispatch(const char* shaderName, const void* constants, int constantsSize, hlsl::Tex* tex, int texNum, int gridW, int gridH)
{
    return hlsl::relax_dispatch_impl(shaderName, constants, constantsSize, tex, texNum, gridW, gridH);
}
