// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
#include "oracle.h"

#include <omp.h>
#include <cstring>

using namespace hlsl;

// Clear_Float / Clear_Uint (reference: Shaders/Source/Clear_Float.cs.hlsl, Clear_Uint.cs.hlsl): zero the whole texture
static void ClearTexture(Tex& t)
{
    for (int y = 0; y < t.h; y++) memset(t.at(0, y + t.yoff), 0, size_t(t.w) * t.bpp());
}

// REFERENCE denoiser (Shaders/Source/REFERENCE_TemporalAccumulation.cs.hlsl:19-30, REFERENCE_Copy.cs.hlsl:19-28; 16x16 groups)
static int ReferenceDispatch(const char* shaderName, const void* constants, int constantsSize, Tex* t, int gridW, int gridH)
{
    if (constantsSize < 20) return -2;
    if (!strcmp(shaderName, "REFERENCE_TemporalAccumulation.cs"))
    {
        struct { uint gRectOrigin[2]; float gAccumSpeed, gDebug, gViewZScale; } c;
        memcpy(&c, constants, sizeof(c));
        for (int y = 0; y < gridH * 16; y++)
            for (int x = 0; x < gridW * 16; x++)
            {
                float4 input = t[0].load(x, y), history = t[1].load(x, y);
                t[1].store(x, y, lerp(history, input, float4(c.gAccumSpeed)));
            }
        return 0;
    }
    if (!strcmp(shaderName, "REFERENCE_Copy.cs"))
    {
        struct { float gRectSizeInv[2]; float gSplitScreen, gDebug, gViewZScale; } c;
        memcpy(&c, constants, sizeof(c));
        for (int y = 0; y < gridH * 16; y++)
            for (int x = 0; x < gridW * 16; x++)
            {
                float pixelUvX = (float(x) + 0.5f) * c.gRectSizeInv[0];
                if (pixelUvX > c.gSplitScreen) t[1].store(x, y, t[0].load(x, y));
            }
        return 0;
    }
    return -1;
}

extern "C" int oracle_dispatch(const char* shaderName, const void* constants, int constantsSize, const OracleTexture* textures, int texturesNum, int gridW, int gridH)
{
    Tex tex[32];
    if (texturesNum > 32) return -1;
    for (int i = 0; i < texturesNum; i++)
    {
        tex[i].data = (uint8_t*)textures[i].data;
        tex[i].w = textures[i].width;
        tex[i].h = textures[i].height;
        tex[i].pitch = textures[i].pitchBytes;
        tex[i].fmt = textures[i].format;
        tex[i].yoff = textures[i].firstRow;
        tex[i].ox = textures[i].originX;
        tex[i].oy = textures[i].originY;
    }
    if (!strncmp(shaderName, "Clear_", 6))
    {
        ClearTexture(tex[0]);
        return 0;
    }
    if (!strncmp(shaderName, "REFERENCE_", 10)) return ReferenceDispatch(shaderName, constants, constantsSize, tex, gridW, gridH);
    if (!strncmp(shaderName, "REBLUR_", 7)) return oracle_reblur_dispatch(shaderName, constants, constantsSize, tex, texturesNum, gridW, gridH);
    if (!strncmp(shaderName, "SIGMA_", 6)) return oracle_sigma_dispatch(shaderName, constants, constantsSize, tex, texturesNum, gridW, gridH);
    if (!strncmp(shaderName, "RELAX_", 6)) return oracle_relax_dispatch(shaderName, constants, constantsSize, tex, texturesNum, gridW, gridH);
    return -1;
}

extern "C" int oracle_num_threads() { return omp_get_max_threads(); }
// launchers such as torchrun export OMP_NUM_THREADS=1; the timed CPU baseline asks for all host cores explicitly
extern "C" void oracle_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
