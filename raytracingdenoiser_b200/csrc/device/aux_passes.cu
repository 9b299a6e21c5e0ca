// Auxiliary passes on sm_100a: the split-screen passes of the three families (CommonSettings::splitScreen: the left part of the
// screen shows the noisy input) and the two passes of the REFERENCE denoiser.
// What has to be computed: Shaders/Include/REBLUR_SplitScreen.hlsli:11-47, RELAX_SplitScreen.hlsli:11-50, SIGMA_SplitScreen.hlsli:11-36,
// Shaders/Source/REFERENCE_TemporalAccumulation.cs.hlsl:19-30, REFERENCE_Copy.cs.hlsl:19-28.  All of them are pure streaming
// passes (one load, one store per texel, 128-bit accesses where the format allows).
#include "common.cuh"
#include "../constants.h"
#include "launch.h"

#include <cstring>

namespace nrdb200
{
namespace
{
constexpr float kFp16MaxAux = 65504.0f;

// ---- REBLUR / RELAX: out = in * (viewZ < denoisingRange) for uv.x <= splitScreen -------------------------------------------
struct SplitArgs
{
    Surf z, inDiff, inSpec, outDiff, outSpec;
    float rectSizeInvX, splitScreen, viewZScale, denoisingRange;
    int rectW, rectH;
    int hasDiff, hasSpec;
    int diffShift, specShift; // 1: checkerboarded input packed into the left half, pixel x shows column x >> 1 (REBLUR_SplitScreen.hlsli:24, :36; RELAX_SplitScreen.hlsli:24, :38)
    int rowBegin, rowEnd;
};
__global__ void __launch_bounds__(256) RadianceSplitScreenKernel(const __grid_constant__ SplitArgs a)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = a.rowBegin + blockIdx.y * 8 + threadIdx.y;
    if (x >= a.rectW || y >= a.rectH || y >= a.rowEnd) return;
    const float pixelUvX = __fmul_rn(__fadd_rn((float)x, 0.5f), a.rectSizeInvX);
    if (pixelUvX > a.splitScreen) return;
    const float viewZ = fabsf(LoadR32F(Near(a.z), x, y) * a.viewZScale);
    const bool keep = viewZ < a.denoisingRange;
    const uint2 zero = make_uint2(0u, 0u); // x * 0 in half precision: +0 (the sign of a negative input is not preserved; NaN inputs are not expected)
    if (a.hasDiff) *TexelPtrRW<uint2>(a.outDiff, x, y) = keep ? __ldg(TexelPtr<uint2>(Near(a.inDiff), x >> a.diffShift, y)) : zero;
    if (a.hasSpec) *TexelPtrRW<uint2>(a.outSpec, x, y) = keep ? __ldg(TexelPtr<uint2>(Near(a.inSpec), x >> a.specShift, y)) : zero;
}

// ---- SIGMA: out = (translucent ? IN_TRANSLUCENCY : IsLit(penumbra)) * (viewZ < denoisingRange) -------------------------------
struct SigmaSplitArgs
{
    Surf z, penumbra, translucency, out;
    float rectSizeInvX, splitScreen, viewZScale, denoisingRange;
    int rectW, rectH;
    int translucent;
    int rowBegin, rowEnd;
};
__global__ void __launch_bounds__(256) SigmaSplitScreenKernel(const __grid_constant__ SigmaSplitArgs a)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = a.rowBegin + blockIdx.y * 8 + threadIdx.y;
    if (x >= a.rectW || y >= a.rectH || y >= a.rowEnd) return;
    const float pixelUvX = __fmul_rn(__fadd_rn((float)x, 0.5f), a.rectSizeInvX);
    if (pixelUvX > a.splitScreen) return;
    const float viewZ = fabsf(LoadR32F(Near(a.z), x, y) * a.viewZScale);
    const bool keep = viewZ < a.denoisingRange;
    if (a.translucent)
        StoreU32(a.out, x, y, keep ? LoadU32(Near(a.translucency), x, y) : 0u); // RGBA8 -> float4 -> RGBA8 is the identity
    else
        StoreU8(a.out, x, y, keep && LoadR16F(Near(a.penumbra), x, y) >= kFp16MaxAux ? 255u : 0u);
}

// ---- REFERENCE -------------------------------------------------------------------------------------------------------------
struct ReferenceArgs
{
    Surf in, out;
    float accumSpeed, rectSizeInvX, splitScreen;
    int gridW, gridH;
    int rowBegin, rowEnd;
};
// history = lerp(history, input, accumSpeed) on the RGBA32F history (REFERENCE_TemporalAccumulation.cs.hlsl:24-29)
__global__ void __launch_bounds__(256) ReferenceAccumulateKernel(const __grid_constant__ ReferenceArgs a)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = a.rowBegin + blockIdx.y * 8 + threadIdx.y;
    if (x >= a.gridW || y >= a.gridH || y >= a.rowEnd || !Inside(a.out, x, y)) return;
    const f4 input = Inside(a.in, x, y) ? LoadRGBA16F(Near(a.in), x, y) : mk4(0.0f);
    float4* h = TexelPtrRW<float4>(a.out, x, y);
    const float4 history = *h;
    // lerp(a, b, t) = a + (b - a) * t, products and sums individually rounded like the oracle
    *h = make_float4(__fadd_rn(history.x, __fmul_rn(__fadd_rn(input.x, -history.x), a.accumSpeed)), __fadd_rn(history.y, __fmul_rn(__fadd_rn(input.y, -history.y), a.accumSpeed)),
                     __fadd_rn(history.z, __fmul_rn(__fadd_rn(input.z, -history.z), a.accumSpeed)), __fadd_rn(history.w, __fmul_rn(__fadd_rn(input.w, -history.w), a.accumSpeed)));
}
// OUT_SIGNAL = history for uv.x > splitScreen (REFERENCE_Copy.cs.hlsl:24-27)
__global__ void __launch_bounds__(256) ReferenceCopyKernel(const __grid_constant__ ReferenceArgs a)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = a.rowBegin + blockIdx.y * 8 + threadIdx.y;
    if (x >= a.gridW || y >= a.gridH || y >= a.rowEnd || !Inside(a.out, x, y)) return;
    const float pixelUvX = __fmul_rn(__fadd_rn((float)x, 0.5f), a.rectSizeInvX);
    if (!(pixelUvX > a.splitScreen)) return;
    const f4 v = Inside(a.in, x, y) ? LoadRGBA32F(Near(a.in), x, y) : mk4(0.0f);
    StoreRGBA16F(a.out, x, y, v);
}
} // namespace

cudaError_t LaunchAux(const PassLaunch& p, const char* shader)
{
    const int rows = p.rowEnd - p.rowBegin;
    if (rows <= 0) return cudaSuccess;
    const dim3 block(32, 8);
    const bool reblur = !strncmp(shader, "REBLUR_", 7), relax = !strncmp(shader, "RELAX_", 6);
    if ((reblur || relax) && strstr(shader, "_SplitScreen.cs"))
    {
        SplitArgs a{};
        const char* s = shader + (reblur ? 7 : 6);
        a.hasDiff = !strncmp(s, "Diffuse", 7) ? 1 : 0;
        a.hasSpec = (!strncmp(s, "Specular_", 9) || !strncmp(s, "DiffuseSpecular_", 16)) ? 1 : 0;
        unsigned diffCheckerboard, specCheckerboard;
        if (reblur)
        {
            const ReblurConstants& c = *(const ReblurConstants*)p.constants;
            a.rectSizeInvX = c.gRectSizeInv[0]; a.splitScreen = c.gSplitScreen; a.viewZScale = c.gViewZScale; a.denoisingRange = c.gDenoisingRange;
            a.rectW = c.gRectSizeMinusOne[0] + 1; a.rectH = c.gRectSizeMinusOne[1] + 1;
            diffCheckerboard = c.gDiffCheckerboard; specCheckerboard = c.gSpecCheckerboard;
        }
        else
        {
            const RelaxConstants& c = *(const RelaxConstants*)p.constants;
            a.rectSizeInvX = c.gRectSizeInv[0]; a.splitScreen = c.gSplitScreen; a.viewZScale = c.gViewZScale; a.denoisingRange = c.gDenoisingRange;
            a.rectW = c.gRectSize[0]; a.rectH = c.gRectSize[1];
            diffCheckerboard = c.gDiffCheckerboard; specCheckerboard = c.gSpecCheckerboard;
        }
        a.diffShift = diffCheckerboard != 2 ? 1 : 0;
        a.specShift = specCheckerboard != 2 ? 1 : 0;
        int k = 0;
        a.z = p.tex[k++];
        if (a.hasDiff) a.inDiff = p.tex[k++];
        if (a.hasSpec) a.inSpec = p.tex[k++];
        if (a.hasDiff) a.outDiff = p.tex[k++];
        if (a.hasSpec) a.outSpec = p.tex[k++];
        a.rowBegin = p.rowBegin; a.rowEnd = p.rowEnd;
        NRD_B200_LAUNCH(p, dim3((a.rectW + 31) / 32, (rows + 7) / 8), block, a, RadianceSplitScreenKernel);
        return cudaGetLastError();
    }
    if (!strncmp(shader, "SIGMA_", 6) && strstr(shader, "_SplitScreen.cs"))
    {
        const SigmaConstants& c = *(const SigmaConstants*)p.constants;
        SigmaSplitArgs a{};
        a.translucent = !strncmp(shader, "SIGMA_ShadowTranslucency_", 25) ? 1 : 0;
        a.rectSizeInvX = c.gRectSizeInv[0]; a.splitScreen = c.gSplitScreen; a.viewZScale = c.gViewZScale; a.denoisingRange = c.gDenoisingRange;
        a.rectW = c.gRectSizeMinusOne[0] + 1; a.rectH = c.gRectSizeMinusOne[1] + 1;
        int k = 0;
        a.z = p.tex[k++]; a.penumbra = p.tex[k++];
        if (a.translucent) a.translucency = p.tex[k++];
        a.out = p.tex[k++];
        a.rowBegin = p.rowBegin; a.rowEnd = p.rowEnd;
        NRD_B200_LAUNCH(p, dim3((a.rectW + 31) / 32, (rows + 7) / 8), block, a, SigmaSplitScreenKernel);
        return cudaGetLastError();
    }
    if (!strcmp(shader, "REFERENCE_TemporalAccumulation.cs") || !strcmp(shader, "REFERENCE_Copy.cs"))
    {
        ReferenceArgs a{};
        a.in = p.tex[0]; a.out = p.tex[1];
        a.gridW = p.gridW * 16; a.gridH = p.gridH * 16;
        a.rowBegin = p.rowBegin; a.rowEnd = p.rowEnd;
        const int w = p.preloadOnly ? 32 : a.out.w;
        if (!strcmp(shader, "REFERENCE_Copy.cs"))
        {
            const ReferenceCopyConstants& c = *(const ReferenceCopyConstants*)p.constants;
            a.rectSizeInvX = c.gRectSizeInv[0]; a.splitScreen = c.gSplitScreen;
            NRD_B200_LAUNCH(p, dim3((w + 31) / 32, (rows + 7) / 8), block, a, ReferenceCopyKernel);
        }
        else
        {
            a.accumSpeed = ((const ReferenceAccumulateConstants*)p.constants)->gAccumSpeed;
            NRD_B200_LAUNCH(p, dim3((w + 31) / 32, (rows + 7) / 8), block, a, ReferenceAccumulateKernel);
        }
        return cudaGetLastError();
    }
    return cudaErrorNotSupported;
}

#if !defined(NRD_B200_NO_STRIPS)
cudaError_t SetPeerTableAux(int slot, const PeerTable* table) { return SetPeerTableThisTU(slot, table); }
#endif
} // namespace nrdb200
