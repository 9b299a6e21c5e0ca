// Software samplers with D3D clamp-to-edge addressing and the CatRom-12 / custom-weight bilinear history filter
// (Common.hlsli:602-656), shared by the RELAX kernels.
#pragma once
#include "common.cuh"

namespace nrdb200
{
namespace smp
{
__device__ __forceinline__ f4 FetchClamped4(const Surf& s, int x, int y) { return LoadRGBA16F(s, clampi(x, 0, s.w - 1), clampi(y, 0, s.h - 1)); }
__device__ __forceinline__ float FetchClamped1(const Surf& s, int x, int y) { return LoadR16F(s, clampi(x, 0, s.w - 1), clampi(y, 0, s.h - 1)); }

__device__ __forceinline__ f4 SampleLinear4(const Surf& s, float u, float v)
{
    float px = u * (float)s.w - 0.5f, py = v * (float)s.h - 0.5f;
    float fx = floorf(px), fy = floorf(py);
    float wx = px - fx, wy = py - fy;
    int x0 = (int)fx, y0 = (int)fy;
    f4 a = lerp4(FetchClamped4(s, x0, y0), FetchClamped4(s, x0 + 1, y0), wx);
    f4 b = lerp4(FetchClamped4(s, x0, y0 + 1), FetchClamped4(s, x0 + 1, y0 + 1), wx);
    return lerp4(a, b, wy);
}
__device__ __forceinline__ float SampleLinear1(const Surf& s, float u, float v)
{
    float px = u * (float)s.w - 0.5f, py = v * (float)s.h - 0.5f;
    float fx = floorf(px), fy = floorf(py);
    float wx = px - fx, wy = py - fy;
    int x0 = (int)fx, y0 = (int)fy;
    float a = lerpf(FetchClamped1(s, x0, y0), FetchClamped1(s, x0 + 1, y0), wx);
    float b = lerpf(FetchClamped1(s, x0, y0 + 1), FetchClamped1(s, x0 + 1, y0 + 1), wx);
    return lerpf(a, b, wy);
}


// CatRom-12 (no corners) through 5 bilinear taps with fallback to the custom-weight bilinear footprint
// (Common.hlsli:602-656).  HasFast: additionally resolve a scalar texture with the custom weights (4 loads).
struct CatRomSetup
{
    float u01x, u01y, u01z, u01w, u23x, u23y, u23z, u23w, u4x, u4y;
    f4 w;
    float w4, sum;
    int bx, by;
    float tcx, tcy; // position of the merged inner taps between texel bx and bx+1 (by and by+1)
    bool bicubic;
};
__device__ __forceinline__ CatRomSetup SetupCatRom(f2 samplePos, const float* invSize, f4 customWeights, bool useBicubic)
{
    CatRomSetup s;
    float cx = floorf(samplePos.x - 0.5f) + 0.5f, cy = floorf(samplePos.y - 0.5f) + 0.5f;
    float fx = saturate(samplePos.x - cx), fy = saturate(samplePos.y - cy);
    const float S = 0.5f;
    float w0x = fx * (fx * (-S * fx + 2.0f * S) - S), w0y = fy * (fy * (-S * fy + 2.0f * S) - S);
    float w1x = fx * (fx * ((2.0f - S) * fx - (3.0f - S))) + 1.0f, w1y = fy * (fy * ((2.0f - S) * fy - (3.0f - S))) + 1.0f;
    float w2x = fx * (fx * (-(2.0f - S) * fx + (3.0f - 2.0f * S)) + S), w2y = fy * (fy * (-(2.0f - S) * fy + (3.0f - 2.0f * S)) + S);
    float w3x = fx * (fx * (S * fx - S)), w3y = fy * (fy * (S * fy - S));
    float w12x = w1x + w2x, w12y = w1y + w2y;
    float tcx = w2x / w12x, tcy = w2y / w12y;
    f4 w = mk4(w12x * w0y, w0x * w12y, w12x * w12y, w3x * w12y);
    float w4 = w12x * w3y;
    s.w = useBicubic ? w : customWeights;
    s.w4 = useBicubic ? w4 : 0.0f;
    s.sum = s.w.x + s.w.y + s.w.z + s.w.w + s.w4;
    if (useBicubic)
    {
        s.u01x = cx + tcx; s.u01y = cy - 1.0f; s.u01z = cx - 1.0f; s.u01w = cy + tcy;
        s.u23x = cx + tcx; s.u23y = cy + tcy;  s.u23z = cx + 2.0f; s.u23w = cy + tcy;
        s.u4x = cx + tcx;  s.u4y = cy + 2.0f;
    }
    else
    {
        s.u01x = cx;        s.u01y = cy;        s.u01z = cx + 1.0f; s.u01w = cy;
        s.u23x = cx;        s.u23y = cy + 1.0f; s.u23z = cx + 1.0f; s.u23w = cy + 1.0f;
        s.u4x = cx + fx;    s.u4y = cy + fy;
    }
    s.u01x *= invSize[0]; s.u01y *= invSize[1]; s.u01z *= invSize[0]; s.u01w *= invSize[1];
    s.u23x *= invSize[0]; s.u23y *= invSize[1]; s.u23z *= invSize[0]; s.u23w *= invSize[1];
    s.u4x *= invSize[0];  s.u4y *= invSize[1];
    s.bx = (int)cx;
    s.by = (int)cy;
    s.tcx = tcx;
    s.tcy = tcy;
    s.bicubic = useBicubic;
    return s;
}
// The reference takes 5 hardware-bilinear taps; in software that is 20 texel loads for a footprint of 12 distinct texels.
// Every texel is loaded once and weighted with (tap weight) x (its bilinear weight): the same polynomial summed in a different
// order (~1e-6 relative).  Without a valid bicubic footprint: the 2x2 texels with the custom bilinear weights.
__device__ __forceinline__ f4 ResolveCatRom4(const CatRomSetup& s, const Surf& tex)
{
    const int x0 = s.bx, y0 = s.by;
    f4 color;
    if (s.bicubic)
    {
        const float ax = 1.0f - s.tcx, bx = s.tcx, ay = 1.0f - s.tcy, by = s.tcy;
        color = FetchClamped4(tex, x0, y0 - 1) * (s.w.x * ax) + FetchClamped4(tex, x0 + 1, y0 - 1) * (s.w.x * bx);
        color = color + FetchClamped4(tex, x0 - 1, y0) * (s.w.y * ay) + FetchClamped4(tex, x0 - 1, y0 + 1) * (s.w.y * by);
        color = color + FetchClamped4(tex, x0, y0) * (s.w.z * ax * ay) + FetchClamped4(tex, x0 + 1, y0) * (s.w.z * bx * ay);
        color = color + FetchClamped4(tex, x0, y0 + 1) * (s.w.z * ax * by) + FetchClamped4(tex, x0 + 1, y0 + 1) * (s.w.z * bx * by);
        color = color + FetchClamped4(tex, x0 + 2, y0) * (s.w.w * ay) + FetchClamped4(tex, x0 + 2, y0 + 1) * (s.w.w * by);
        color = color + FetchClamped4(tex, x0, y0 + 2) * (s.w4 * ax) + FetchClamped4(tex, x0 + 1, y0 + 2) * (s.w4 * bx);
    }
    else
        color = FetchClamped4(tex, x0, y0) * s.w.x + FetchClamped4(tex, x0 + 1, y0) * s.w.y + FetchClamped4(tex, x0, y0 + 1) * s.w.z + FetchClamped4(tex, x0 + 1, y0 + 1) * s.w.w;
    return s.sum < 0.0001f ? mk4(0.0f) : color * (1.0f / s.sum);
}
__device__ __forceinline__ float ResolveCatRom1(const CatRomSetup& s, const Surf& tex)
{
    float color = SampleLinear1(tex, s.u01x, s.u01y) * s.w.x;
    color += SampleLinear1(tex, s.u01z, s.u01w) * s.w.y;
    color += SampleLinear1(tex, s.u23x, s.u23y) * s.w.z;
    color += SampleLinear1(tex, s.u23z, s.u23w) * s.w.w;
    if (s.w4 != 0.0f) color += SampleLinear1(tex, s.u4x, s.u4y) * s.w4;
    return s.sum < 0.0001f ? 0.0f : color / s.sum;
}
} // namespace smp
} // namespace nrdb200
