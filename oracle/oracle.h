// ORACLE -- TEST INFRASTRUCTURE ONLY.  C entry points of the CPU restatement (liboracle.so).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
// Parity: see hlsl.h (pinned against the reference's shader sources run on the CPU; MathLib restated).
#pragma once
#include "hlsl.h"

extern "C" {
// One texture binding of a dispatch, as an RHI executor would bind it (inputs first, then outputs).
struct OracleTexture
{
    void* data;       // texel (0, firstRow)
    int32_t width, height;
    int32_t pitchBytes;
    int32_t format;   // nrd::Format
    int32_t firstRow; // 0 unless the caller holds a strip
    int32_t originX, originY; // CommonSettings::rectOrigin for the inputs the shaders read through WithRectOrigin, else 0
};
// Executes one DispatchDesc on the CPU.  shaderName = PipelineDesc::shaderFileName of the dispatch's pipeline.
// Returns 0 on success, -1 unknown pass, -2 bad constants.
int oracle_dispatch(const char* shaderName, const void* constants, int constantsSize, const OracleTexture* textures, int texturesNum, int gridW, int gridH);
int oracle_num_threads();
void oracle_set_num_threads(int n);
}

int oracle_reblur_dispatch(const char* shaderName, const void* constants, int constantsSize, hlsl::Tex* tex, int texNum, int gridW, int gridH);
int oracle_sigma_dispatch(const char* shaderName, const void* constants, int constantsSize, hlsl::Tex* tex, int texNum, int gridW, int gridH);
int oracle_relax_dispatch(const char* shaderName, const void* constants, int constantsSize, hlsl::Tex* tex, int texNum, int gridW, int gridH);
