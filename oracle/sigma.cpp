// ORACLE -- TEST INFRASTRUCTURE ONLY (see hlsl.h).  Every pass below is checked against the reference's own shader source of that pass,
// compiled for the CPU (oracle/build_refshaders.py), by tests/test_reference_shaders.py; the reference ships no golden vectors.
// CPU restatement of the reference's SIGMA_SHADOW / SIGMA_SHADOW_TRANSLUCENCY passes at the default compile-time switches
// (SIGMA_Config.hlsli:13-43; SIGMA_TYPE is float / float4: the signal is carried as float4 here, the scalar variant uses .x only):
//   ClassifyTiles            Shaders/Include/SIGMA_ClassifyTiles.hlsli:10-81
//   SmoothTiles              Shaders/Include/SIGMA_SmoothTiles.hlsli:10-48
//   Copy                     Shaders/Include/SIGMA_Copy.hlsli:10-24
//   Blur / PostBlur          Shaders/Include/SIGMA_Blur.hlsli:11-268 (SIGMA_FIRST_PASS on / off)
//   TemporalStabilization    Shaders/Include/SIGMA_TemporalStabilization.hlsli:10-226
//   helpers                  Shaders/Include/SIGMA_Common.hlsli:13-130
#include "common_hlsli.h"
#include "oracle.h"

#include <cstring>

namespace hlsl
{
using namespace common;
namespace
{
// SIGMA_Config.hlsli:46-80
struct CB
{
    float4x4 gWorldToView, gViewToClip, gWorldToClipPrev, gWorldToViewPrev;
    float4 gRotator, gRotatorPost, gViewVectorWorld, gLightDirectionView, gFrustum, gFrustumPrev, gCameraDelta, gMvScale;
    float2 gResourceSizeInv, gResourceSizeInvPrev, gRectSize, gRectSizeInv, gRectSizePrev, gResolutionScale, gRectOffset;
    uint gPrintfAt[2], gRectOrigin[2];
    int gRectSizeMinusOne[2], gTilesSizeMinusOne[2];
    float gOrthoMode, gUnproject, gDenoisingRange, gPlaneDistSensitivity, gStabilizationStrength, gDebug, gSplitScreen, gViewZScale, gMinRectDimMulUnproject;
    uint gFrameIndex, gIsRectChanged;
};
static_assert(sizeof(CB) == 516, "SIGMA_SHARED_CONSTANTS is 516 bytes before padding");

const float SIGMA_MAX_PIXEL_RADIUS = 32.0f;
const float SIGMA_TS_SIGMA_SCALE = 3.0f;
const float SIGMA_MAX_ACCUM_FRAME_NUM = 7.0f;

float4 PackShadow(float4 s) { return sqrt(saturate(s)); }              // SIGMA_Common.hlsli:13 (Math::Sqrt01 per component)
bool IsLit(float p) { return p >= NRD_FP16_MAX; }                      // :14
float4 UnpackShadow(float4 s) { return s * s; }                        // NRD.hlsli:931
float4 GetStdDev4(float4 m1, float4 m2) { return sqrt(abs(m2 - m1 * m1)); }
float4 clamp4(float4 v, float4 a, float4 b) { return float4(clamp(v.x, a.x, b.x), clamp(v.y, a.y, b.y), clamp(v.z, a.z, b.z), clamp(v.w, a.w, b.w)); }

float GetKernelRadiusInPixels(float hitDist, float unprojectZ, float scale = 1.0f)   // SIGMA_Common.hlsli:21-34
{
    float unclampedRadius = hitDist / unprojectZ;
    unclampedRadius *= scale;
    float minRadius = min(unclampedRadius, 2.0f);
    return clamp(unclampedRadius, minRadius, SIGMA_MAX_PIXEL_RADIUS);
}
float AreBothLitOrUnlit(float p1, float p2) { return float((p1 == 0.0f) == (p2 == 0.0f)); }   // :36-42

// SIGMA_Common.hlsli:46-75
float2 FilterBicubic(float2 size, float2 uv, float4& uv_10_00, float4& uv_11_01)
{
    const float4 c1(3.0f, 0.0f, 1.0f, 4.0f), c2(-1.0f, 3.0f, -3.0f, 1.0f), c3(3.0f, -6.0f, -3.0f, 0.0f);
    const float k = 1.0f / 6.0f;
    float4 dxdy = -float4(c1.z, c1.y, c1.y, c1.z) / float4(size.x, size.y, size.x, size.y);
    float2 f = frac(uv * size - float2(0.5f));
    float2 f2 = f * f;
    float2 f3 = f2 * f;
    float3 xw, yw;
    float4 phi;
    phi = float4(k) * (c2 * float4(f3.x) + float4(c3.x, c3.y, c3.x, c3.w) * float4(f2.x) + float4(c3.z, c3.w, c3.x, c3.w) * float4(f.x) + float4(c1.z, c1.w, c1.z, c1.y));
    xw.x = c2.w + c2.w * f.x + c2.x * phi.y / (phi.x + phi.y);
    xw.y = c2.w + c2.x * f.x + c2.w * phi.w / (phi.z + phi.w);
    xw.z = phi.x + phi.y;
    phi = float4(k) * (c2 * float4(f3.y) + float4(c3.x, c3.y, c3.x, c3.w) * float4(f2.y) + float4(c3.z, c3.w, c3.x, c3.w) * float4(f.y) + float4(c1.z, c1.w, c1.z, c1.y));
    yw.x = c2.w + c2.w * f.y + c2.x * phi.y / (phi.x + phi.y);
    yw.y = c2.w + c2.x * f.y + c2.w * phi.w / (phi.z + phi.w);
    yw.z = phi.x + phi.y;
    uv_10_00 = float4(uv, uv) + float4(c2.w, c2.w, c2.x, c2.x) * float4(xw.x, xw.x, xw.y, xw.y) * float4(dxdy.x, dxdy.y, dxdy.x, dxdy.y);
    uv_11_01 = uv_10_00 + float4(yw.x) * float4(dxdy.z, dxdy.w, dxdy.z, dxdy.w);
    uv_10_00 -= float4(yw.y) * float4(dxdy.z, dxdy.w, dxdy.z, dxdy.w);
    return float2(yw.z, xw.z);
}
// :77-95
float2 TextureCubic(const Tex& tex, float2 uv)
{
    float2 size(float(tex.w), float(tex.h));
    float4 uv_10_00, uv_11_01;
    float2 t = FilterBicubic(size, uv, uv_10_00, uv_11_01);
    float2 c00 = tex.sampleLinear(uv_10_00.zw()).xy();
    float2 c10 = tex.sampleLinear(uv_10_00.xy()).xy();
    float2 c01 = tex.sampleLinear(uv_11_01.zw()).xy();
    float2 c11 = tex.sampleLinear(uv_11_01.xy()).xy();
    c00 = lerp(c00, c01, t.x);
    c10 = lerp(c10, c11, t.x);
    return lerp(c00, c10, t.y);
}

struct Pass
{
    const CB& c;
    explicit Pass(const CB& cb) : c(cb) {}
    float UnpackViewZ(float z) const { return abs(z * c.gViewZScale); }
    float2 ClampUvToViewport(float2 uv) const { return min(uv * c.gResolutionScale, c.gResolutionScale - float2(0.5f) * c.gResourceSizeInv); }
    float3 GetViewVector(float3 X, bool isViewSpace) const { return c.gOrthoMode == 0.0f ? normalize(-X) : (isViewSpace ? float3(0, 0, -1) : c.gViewVectorWorld.xyz()); }
};

void ClassifyTiles(const Pass& P, bool translucent, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_ViewZ = t[0], &gIn_Penumbra = t[1];
    const Tex* gIn_Shadow_Translucency = translucent ? &t[2] : nullptr;
    Tex& gOut_Tiles = t[translucent ? 3 : 2];
#pragma omp parallel for schedule(static)
    for (int ty = 0; ty < gridH; ty++)
        for (int tx = 0; tx < gridW; tx++)
        {
            uint mask = 0;
            float maxRadius = 0.0f;
            for (int j = 0; j < 16; j++)
                for (int i = 0; i < 16; i++)
                {
                    int x = tx * 16 + i, y = ty * 16 + j;
                    float h = gIn_Penumbra.load(x, y).x;
                    float viewZ = P.UnpackViewZ(gIn_ViewZ.load(x, y).x);
                    bool isInf = viewZ > c.gDenoisingRange;
                    bool isShadow = h == 0.0f;
                    bool isLit = IsLit(h);
                    bool isOpaque = true;
                    if (translucent) // SIGMA_ClassifyTiles.hlsli:45-48
                    {
                        float4 tr = gIn_Shadow_Translucency->load(x, y);
                        isOpaque = Color::Luminance(float3(tr.y, tr.z, tr.w)) < 0.003f;
                    }
                    mask += ((isLit || isInf || isShadow) ? 1u : 0u) << 0;
                    mask += (((!isLit && isOpaque) || isInf || isShadow) ? 1u : 0u) << 9;
                    mask += (isInf ? 1u : 0u) << 18;
                    float hitDist = (isLit || isInf) ? 0.0f : h;
                    float pixelSize = PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, 1.0f, viewZ);
                    float pixelRadius = GetKernelRadiusInPixels(hitDist, pixelSize);
                    maxRadius = max(pixelRadius, maxRadius);
                }
            bool isLit = ((mask >> 0) & 511) == 256;
            bool isUmbra = ((mask >> 9) & 511) == 256;
            bool isInf = ((mask >> 18) & 511) == 256;
            float4 result;
            result.x = (isLit || isUmbra) ? 0.0f : 1.0f;
            result.y = saturate(maxRadius / 16.0f);
            result.z = isInf ? 1.0f : 0.0f;
            result.w = 0.0f;
            gOut_Tiles.store(tx, ty, result);
        }
}

void SmoothTiles(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex& gIn_Tiles = t[0];
    Tex& gOut_Tiles = t[1];
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 16; x++)
        {
            float4 center = gIn_Tiles.load(x, y);
            float blurry = 0.0f, sum = 0.0f;
            float k = 1.01f / (center.y + 0.01f);
            for (int j = 0; j <= 2; j++)
                for (int i = 0; i <= 2; i++)
                {
                    float d = length(float2(float(i), float(j)) - float2(1.0f));
                    float w = exp2(-k * d * d);
                    int2 p = clamp(int2(x + i - 1, y + j - 1), int2(0), int2(c.gTilesSizeMinusOne[0], c.gTilesSizeMinusOne[1]));
                    blurry += gIn_Tiles.load(p).x * w;
                    sum += w;
                }
            blurry /= sum;
            gOut_Tiles.store(x, y, float4(center.z, blurry, 0, 0));
        }
}

void Copy(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_Tiles = t[0], &gIn_History = t[1], &gIn_HistoryLength = t[2];
    Tex &gOut_History = t[3], &gOut_HistoryLength = t[4];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f && !c.gIsRectChanged) continue;
            gOut_History.store(x, y, gIn_History.load(x, y));
            gOut_HistoryLength.storeu(x, y, gIn_HistoryLength.loadu(x, y));
        }
}

void Blur(const Pass& P, bool firstPass, bool translucent, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_ViewZ = t[0], &gIn_Normal_Roughness = t[1], &gIn_Penumbra = t[2], &gIn_Tiles = t[3];
    // gIn_Shadow_Translucency is bound unless this is the first pass of the opaque variant (SIGMA_Blur.resources.hlsli:25-27)
    const bool hasShadowInput = !firstPass || translucent;
    const Tex* gIn_Shadow = hasShadowInput ? &t[4] : nullptr;
    Tex& gOut_Penumbra = t[hasShadowInput ? 5 : 4];
    Tex& gOut_Shadow = t[hasShadowInput ? 6 : 5];
    const int2 rectMax(c.gRectSizeMinusOne[0], c.gRectSizeMinusOne[1]);
    const int BORDER = 2;

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f || x > rectMax.x || y > rectMax.y) continue;
            // "shared memory": clamped loads (Preload :14-36)
            auto sPenumbraViewZ = [&](int i, int j) {
                int2 p = clamp(int2(x + i - BORDER, y + j - BORDER), int2(0), rectMax);
                return float2(gIn_Penumbra.load(p).x, P.UnpackViewZ(gIn_ViewZ.load(p).x));
            };
            auto sShadow = [&](int i, int j) {
                int2 p = clamp(int2(x + i - BORDER, y + j - BORDER), int2(0), rectMax);
                float4 s = hasShadowInput ? gIn_Shadow->load(p) : float4(float(IsLit(gIn_Penumbra.load(p).x)));
                if (!firstPass) s = UnpackShadow(s);
                return s;
            };
            float2 centerData = sPenumbraViewZ(BORDER, BORDER);
            float centerPenumbra = centerData.x, viewZ = centerData.y;
            if (viewZ > c.gDenoisingRange) continue;

            float2 pixelUv = (tofloat(pixelPos) + float2(0.5f)) * c.gRectSizeInv;
            float tileValue = TextureCubic(gIn_Tiles, pixelUv * c.gResolutionScale).y;
            if (tileValue == 0.0f || centerPenumbra == 0.0f)
            {
                gOut_Penumbra.store(pixelPos, centerPenumbra);
                gOut_Shadow.store(pixelPos, PackShadow(sShadow(BORDER, BORDER)));
                continue;
            }

            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, c.gFrustum, viewZ, c.gOrthoMode);
            float3 N = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(pixelPos)).xyz();
            float3 Nv = Geometry::RotateVector(c.gWorldToView, N);
            float pixelSize = PixelRadiusToWorld(c.gUnproject, c.gOrthoMode, 1.0f, viewZ);
            float frustumSize = GetFrustumSize(c.gMinRectDimMulUnproject, c.gOrthoMode, viewZ);
            float3 Vv = P.GetViewVector(Xv, true);
            float NoV = abs(dot(Nv, Vv));
            float2 geometryWeightParams = GetGeometryWeightParams(c.gPlaneDistSensitivity, frustumSize, Xv, Nv);

            float2 sum(0.0f);
            float penumbra = 0.0f;
            float4 result(0.0f), centerTap(0.0f);
            for (int j = 0; j <= BORDER * 2; j++)
                for (int i = 0; i <= BORDER * 2; i++)
                {
                    float2 data = sPenumbraViewZ(i, j);
                    float penum = data.x, zs = data.y;
                    float4 s = sShadow(i, j);
                    float w = 1.0f;
                    if (i == BORDER && j == BORDER) centerTap = s;
                    else
                    {
                        float2 uv = pixelUv + float2(float(i - BORDER), float(j - BORDER)) * c.gRectSizeInv;
                        float3 Xvs = Geometry::ReconstructViewPosition(uv, c.gFrustum, zs, c.gOrthoMode);
                        w *= ComputeWeight(dot(Nv, Xvs), geometryWeightParams.x, geometryWeightParams.y);
                        w *= AreBothLitOrUnlit(centerPenumbra, penum);
                        w *= GetGaussianWeight(length(float2(float(i - BORDER), float(j - BORDER)) / float2(float(BORDER))));
                    }
                    result += w == 0.0f ? float4(0.0f) : s * float4(w);
                    sum.x += w;
                    w *= pixelSize / (pixelSize + penum);
                    w *= float(!IsLit(penum));
                    penumbra += w == 0.0f ? 0.0f : penum * w;
                    sum.y += w;
                }
            result /= float4(sum.x);
            sum.x = 1.0f;
            penumbra /= max(sum.y, NRD_EPS);
            sum.y = float(sum.y != 0.0f);

            float penumbraInPixels = penumbra / pixelSize;
            float f = Math::SmoothStep(0.0f, float(BORDER), penumbraInPixels);
            result = lerp(centerTap, result, f);

            f = lerp(4.0f, 1.0f, f);
            result *= float4(f);
            penumbra *= f;
            sum *= float2(f);

            float blurRadius = GetKernelRadiusInPixels(penumbra, pixelSize, tileValue);
            float4 rotator = firstPass ? c.gRotator : c.gRotatorPost; // SIGMA_ROTATOR_MODE = NRD_FRAME
            float2 skew = lerp(float2(1.0f) - abs(Nv.xy()), float2(1.0f), NoV);
            skew /= float2(max(skew.x, skew.y));
            skew *= c.gRectSizeInv * float2(blurRadius);
            float4 scaledRotator = Geometry::ScaleRotator(rotator, skew);

            float invEstimatedPenumbra = 1.0f / max(penumbra, NRD_EPS);
            for (uint n = 0; n < 8; n++)
            {
                float3 offset = g_Special8[n];
                float2 uv = pixelUv + Geometry::RotateVector(scaledRotator, offset.xy());
                uv = (floor(uv * c.gRectSize) + float2(0.5f)) * c.gRectSizeInv;
                float2 uvScaled = P.ClampUvToViewport(uv);
                float penum = gIn_Penumbra.sampleNearest(uvScaled).x;
                float zs = P.UnpackViewZ(gIn_ViewZ.sampleNearest(uvScaled).x);
                float4 s = hasShadowInput ? gIn_Shadow->sampleNearest(uvScaled) : float4(float(IsLit(penum)));
                if (!firstPass) s = UnpackShadow(s);
                float3 Xvs = Geometry::ReconstructViewPosition(uv, c.gFrustum, zs, c.gOrthoMode);
                float w = IsInScreenNearest(uv);
                w *= ComputeWeight(dot(Nv, Xvs), geometryWeightParams.x, geometryWeightParams.y);
                w *= AreBothLitOrUnlit(centerPenumbra, penum);
                w *= GetGaussianWeight(offset.z);
                w *= saturate(penum * invEstimatedPenumbra);
                result += w == 0.0f ? float4(0.0f) : s * float4(w);
                sum.x += w;
                w *= pixelSize / (pixelSize + penum);
                w *= float(!IsLit(penum));
                penumbra += w == 0.0f ? 0.0f : penum * w;
                sum.y += w;
            }
            result /= float4(sum.x);
            penumbra = sum.y == 0.0f ? centerPenumbra : penumbra / sum.y;

            if (firstPass || c.gStabilizationStrength != 0.0f) gOut_Penumbra.store(pixelPos, penumbra);
            gOut_Shadow.store(pixelPos, PackShadow(result));
        }
}

uint PackViewZAndHistoryLength(float viewZ, float historyLength)   // SIGMA_TemporalStabilization.hlsli:25-31
{
    uint p = asuint(viewZ) & ~7u;
    p |= std::min(uint(historyLength + 0.5f), 7u);
    return p;
}

void TemporalStabilization(const Pass& P, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    const Tex &gIn_ViewZ = t[0], &gIn_Mv = t[1], &gIn_Penumbra = t[2], &gIn_Shadow = t[3], &gIn_History = t[4], &gIn_HistoryLength = t[5], &gIn_Tiles = t[6];
    Tex &gOut_Shadow = t[7], &gOut_HistoryLength = t[8];
    const int2 rectMax(c.gRectSizeMinusOne[0], c.gRectSizeMinusOne[1]);
    const int BORDER = 2;

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            auto sShadow = [&](int i, int j) { return UnpackShadow(gIn_Shadow.load(clamp(int2(x + i - BORDER, y + j - BORDER), int2(0), rectMax))); };
            auto sPenumbra = [&](int i, int j) { return gIn_Penumbra.load(clamp(int2(x + i - BORDER, y + j - BORDER), int2(0), rectMax)).x; };
            float viewZ = P.UnpackViewZ(gIn_ViewZ.load(pixelPos).x);
            if (isSky != 0.0f || x > rectMax.x || y > rectMax.y || viewZ > c.gDenoisingRange) continue;
            float centerPenumbra = sPenumbra(BORDER, BORDER);

            float2 pixelUv = (tofloat(pixelPos) + float2(0.5f)) * c.gRectSizeInv;
            float tileValue = TextureCubic(gIn_Tiles, pixelUv * c.gResolutionScale).y;
            bool isHardShadow = tileValue == 0.0f || centerPenumbra == 0.0f;
            if (isHardShadow)
            {
                gOut_Shadow.store(pixelPos, PackShadow(sShadow(BORDER, BORDER)));
                gOut_HistoryLength.storeu(pixelPos, PackViewZAndHistoryLength(viewZ, SIGMA_MAX_ACCUM_FRAME_NUM));
                continue;
            }

            float sum = 0.0f;
            float4 m1(0.0f), m2(0.0f), input(0.0f);
            for (int j = 0; j <= BORDER * 2; j++)
                for (int i = 0; i <= BORDER * 2; i++)
                {
                    float4 s = sShadow(i, j);
                    float w = 1.0f;
                    if (i == BORDER && j == BORDER) input = s;
                    else
                    {
                        float penum = sPenumbra(i, j);
                        w = AreBothLitOrUnlit(centerPenumbra, penum);
                        w *= GetGaussianWeight(length(float2(float(i - BORDER), float(j - BORDER)) / float2(float(BORDER))));
                    }
                    m1 += s * float4(w);
                    m2 += s * s * float4(w);
                    sum += w;
                }
            m1 /= float4(sum);
            m2 /= float4(sum);
            float4 sigma = GetStdDev4(m1, m2);

            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, c.gFrustum, viewZ, c.gOrthoMode);
            float3 X = Geometry::RotateVectorInverse(c.gWorldToView, Xv);
            float3 mv = gIn_Mv.load(pixelPos).xyz() * c.gMvScale.xyz();
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

            Filtering::Bilinear smbBilinearFilter = Filtering::GetBilinearFilter(smbPixelUv, c.gRectSizePrev);
            float2 smbBilinearGatherUv = (smbBilinearFilter.origin + float2(1.0f)) * c.gResourceSizeInvPrev;
            uint4 g = gIn_HistoryLength.gatheru(smbBilinearGatherUv);
            uint4 prevData = {g.w, g.z, g.x, g.y};
            float4 prevViewZ = float4(asfloat(prevData.x & ~7u), asfloat(prevData.y & ~7u), asfloat(prevData.z & ~7u), asfloat(prevData.w & ~7u));
            float4 prevHistoryLength = float4(float(prevData.x & 7u), float(prevData.y & 7u), float(prevData.z & 7u), float(prevData.w & 7u));

            float frustumSize = GetFrustumSize(c.gMinRectDimMulUnproject, c.gOrthoMode, viewZ);
            float disocclusionThreshold = GetDisocclusionThreshold(NRD_DISOCCLUSION_THRESHOLD, frustumSize, 1.0f);
            disocclusionThreshold *= IsInScreenNearest(smbPixelUv);
            disocclusionThreshold -= NRD_EPS;

            float3 Xvprev = Geometry::AffineTransform(c.gWorldToViewPrev, Xprev);
            float4 smbPlaneDist = abs(prevViewZ - float4(Xvprev.z));
            float4 smbOcclusion = step(smbPlaneDist, float4(disocclusionThreshold));
            float4 smbOcclusionWeights = Filtering::GetBilinearCustomWeights(smbBilinearFilter, smbOcclusion);
            float historyLength = Filtering::ApplyBilinearCustomWeights(prevHistoryLength.x, prevHistoryLength.y, prevHistoryLength.z, prevHistoryLength.w, smbOcclusionWeights);

            bool isCatRomAllowed = dot(smbOcclusionWeights, float4(1.0f)) > 3.5f;
            float4 history = BicubicCustom(saturate(smbPixelUv) * c.gRectSizePrev, c.gResourceSizeInvPrev, smbOcclusionWeights, isCatRomAllowed, gIn_History);
            history = saturate(history);
            history = UnpackShadow(history);

            sigma *= float4(lerp(SIGMA_TS_SIGMA_SCALE, 1.0f, 1.0f / (1.0f + historyLength)));
            float4 inputMin = m1 - sigma, inputMax = m1 + sigma;
            float4 historyClamped = clamp4(history, inputMin, inputMax);

            float antilag = abs(historyClamped.x - history.x); // ".x" only, also for the translucent variant (:174)
            antilag = Math::Sqrt01(antilag);
            antilag = saturate(1.0f - antilag);
            historyLength *= antilag;
            float historyWeight = historyLength / (1.0f + historyLength);
            float streetMagic = 0.6f * historyWeight * antilag;
            historyClamped = lerp(historyClamped, history, streetMagic);
            float4 result = lerp(input, historyClamped, min(c.gStabilizationStrength, historyWeight));
            historyLength = min(historyLength + 1.0f, SIGMA_MAX_ACCUM_FRAME_NUM);

            gOut_Shadow.store(pixelPos, PackShadow(result));
            gOut_HistoryLength.storeu(pixelPos, PackViewZAndHistoryLength(viewZ, historyLength));
        }
}
// SIGMA_SplitScreen.hlsli:11-36
void SplitScreen(const Pass& P, bool translucent, Tex* t, int gridW, int gridH)
{
    const CB& c = P.c;
    int k = 0;
    const Tex &gIn_ViewZ = t[k++], &gIn_Penumbra = t[k++];
    const Tex* gIn_Shadow_Translucency = translucent ? &t[k++] : nullptr;
    Tex& gOut_Shadow_Translucency = t[k++];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 8; x++)
        {
            const int2 pixelPos(x, y);
            float2 pixelUv = (float2(float(x), float(y)) + float2(0.5f)) * c.gRectSizeInv;
            if (pixelUv.x > c.gSplitScreen || x > c.gRectSizeMinusOne[0] || y > c.gRectSizeMinusOne[1]) continue;
            float viewZ = abs(gIn_ViewZ.load(pixelPos).x * c.gViewZScale);
            float4 s = translucent ? gIn_Shadow_Translucency->load(pixelPos) : float4(float(IsLit(gIn_Penumbra.load(pixelPos).x)));
            gOut_Shadow_Translucency.store(pixelPos, s * float4(float(viewZ < c.gDenoisingRange)));
        }
}
} // namespace

int sigma_dispatch_impl(const char* shaderName, const void* constants, int constantsSize, Tex* tex, int gridW, int gridH)
{
    if (constantsSize < 516) return -2; // sizeof() of the reference's C++ struct; CB is padded to whole 16-byte registers
    CB cb{};
    memcpy(&cb, constants, constantsSize < (int)sizeof(CB) ? constantsSize : (int)sizeof(CB));
    Pass P(cb);
    if (!strcmp(shaderName, "SIGMA_SmoothTiles.cs")) SmoothTiles(P, tex, gridW, gridH);
    else if (!strcmp(shaderName, "SIGMA_Copy.cs")) Copy(P, tex, gridW, gridH);
    else
    {
        // "SIGMA_Shadow_<pass>.cs" (SIGMA_TYPE float) / "SIGMA_ShadowTranslucency_<pass>.cs" (SIGMA_TRANSLUCENT, SIGMA_TYPE float4)
        const bool translucent = !strncmp(shaderName, "SIGMA_ShadowTranslucency_", 25);
        if (!translucent && strncmp(shaderName, "SIGMA_Shadow_", 13) != 0) return -1;
        const char* pass = shaderName + (translucent ? 25 : 13);
        if (!strcmp(pass, "ClassifyTiles.cs")) ClassifyTiles(P, translucent, tex, gridW, gridH);
        else if (!strcmp(pass, "Blur.cs")) Blur(P, true, translucent, tex, gridW, gridH);
        else if (!strcmp(pass, "PostBlur.cs")) Blur(P, false, translucent, tex, gridW, gridH);
        else if (!strcmp(pass, "TemporalStabilization.cs")) TemporalStabilization(P, tex, gridW, gridH);
        else if (!strcmp(pass, "SplitScreen.cs")) SplitScreen(P, translucent, tex, gridW, gridH);
        else return -1;
    }
    return 0;
}
} // namespace hlsl

int oracle_sigma_dispatch(const char* shaderName, const void* constants, int constantsSize, hlsl::Tex* tex, int, int gridW, int gridH)
{
    return hlsl::sigma_dispatch_impl(shaderName, constants, constantsSize, tex, gridW, gridH);
}
