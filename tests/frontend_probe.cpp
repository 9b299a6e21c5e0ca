// Host-side probe of include/nrd_b200_frontend.cuh (the header is plain inline C++ outside nvcc): evaluates the helpers on a
// deterministic grid of inputs and prints one line per case; tests/test_frontend_header.py checks the values against the packers
// of raytracingdenoiser_b200/scene.py and independent numpy restatements.
//   g++ -std=c++17 -I include -I /usr/local/cuda/include tests/frontend_probe.cpp
#include "nrd_b200_frontend.cuh"
#include <cstdio>
using namespace nrd_frontend;

static unsigned lcg = 12345u;
static float rnd() { lcg = lcg * 1664525u + 1013904223u; return (float)(lcg >> 8) / 16777216.0f; }

int main()
{
    const float4 hp = f4(3.0f, 0.1f, 20.0f, -25.0f);
    for (int i = 0; i < 400; i++)
    {
        float3 n = f3(rnd() * 2 - 1, rnd() * 2 - 1, rnd() * 2 - 1);
        if (i < 6) n = f3(i == 0, i == 1, i == 2), n = i >= 3 ? f3(-(i == 3), -(i == 4), -(i == 5)) : n; // the six axes
        n = normalize3(n);
        const float rough = i % 7 == 0 ? 0.0f : (i % 7 == 1 ? 1.0f : rnd()), mat = (float)(i % 4);
        const float4 p = NRD_FrontEnd_PackNormalAndRoughness(n, rough, mat);
        const uint32_t bits = nrdPackR10G10B10A2(p);
        float matBack;
        const float4 u = NRD_FrontEnd_UnpackNormalAndRoughness(nrdUnpackR10G10B10A2(bits), matBack);
        const float3 rad = f3(rnd() * 4, rnd() * 4, rnd() * 4);
        const float hit = rnd() * 30, viewz = 0.5f + rnd() * 50;
        const float nh = REBLUR_FrontEnd_GetNormHitDist(hit, viewz, hp, rough);
        const float4 rb = REBLUR_FrontEnd_PackRadianceAndNormHitDist(rad, nh, true);
        const float4 back = REBLUR_BackEnd_UnpackRadianceAndNormHitDist(rb);
        const float4 rx = RELAX_FrontEnd_PackRadianceAndHitDist(rad, hit, true);
        const float dOcc = i % 5 == 0 ? 0.0f : (i % 5 == 1 ? 70000.0f : rnd() * 40);
        const float pen = SIGMA_FrontEnd_PackPenumbra(dOcc, 0.004625f);
        const float pen2 = SIGMA_FrontEnd_PackPenumbra(dOcc, 100.0f, 2.0f);
        float4 sh1;
        const float3 dir = normalize3(f3(rnd() * 2 - 1, rnd() * 2 - 1, rnd() * 2 - 1));
        const float4 sh0 = REBLUR_FrontEnd_PackSh(rad, nh, dir, sh1, true);
        const NRD_SG sg = REBLUR_BackEnd_UnpackSh(sh0, sh1);
        const float3 shd = NRD_SH_ResolveDiffuse(sg, dir), sgc = NRD_SG_ExtractColor(sg), sgd = NRD_SG_ExtractDirection(sg);
        float3 dF, sF;
        NRD_MaterialFactors(n, normalize3(f3(0.2f, 0.3f, 0.9f)), f3(0.5f, 0.4f, 0.3f), f3(0.04f, 0.04f, 0.04f), rough, dF, sF);
        const float4 tr = SIGMA_FrontEnd_PackTranslucency(dOcc, f3(rnd() * 1.5f, rnd(), rnd()));
        printf("%.9g %.9g %.9g %.9g %.9g %u %.9g %.9g %.9g %.9g %.9g "   // 0-10: n, rough, mat, bits, unpacked n + rough + mat
               "%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g "  // 11-21: rad, hit, viewz, nh, packed reblur
               "%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g "            // 22-30: unpacked rgb, relax xyzw, dOcc, pen
               "%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n",
               n.x, n.y, n.z, rough, mat, bits, u.x, u.y, u.z, u.w, matBack, rad.x, rad.y, rad.z, hit, viewz, nh, rb.x, rb.y, rb.z, rb.w, back.x, back.y, back.z, rx.x, rx.y,
               rx.z, rx.w, dOcc, pen, pen2, dir.x, dir.y, dir.z, shd.x, shd.y, shd.z, sgc.x, sgc.y, sgc.z, sgd.x, sgd.y, sgd.z, dF.x, sF.x, tr.x, tr.y, sh1.x, sh1.w,
               REBLUR_GetHitDist(nh, viewz, hp, rough));
    }
    return 0;
}
