// Layout probe of the public NRD API structs: prints, as JSON, sizeof() and the bytes of a default-initialised object of every
// public struct plus the enum extents.  Compiled once against the reference's Include/NRD.h (tests/golden/make_layout_golden.py ->
// tests/golden/nrd_layout.json) and once against include/nrd_b200.h (tests/test_api.py): both must print the same.
//   g++ -std=c++17 -DPROBE_HEADER='"NRD.h"' -I <include dir> layout_probe.cpp
#include PROBE_HEADER
#include <cstdio>
#include <cstring>
#include <new>

template <class T> void Dump(const char* name, bool last = false)
{
    alignas(16) unsigned char buf[sizeof(T)];
    memset(buf, 0, sizeof(buf));
    new (buf) T; // default member initialisers; padding stays zero
    printf("  \"%s\": {\"sizeof\": %zu, \"alignof\": %zu, \"default\": \"", name, sizeof(T), alignof(T));
    for (size_t i = 0; i < sizeof(T); i++) printf("%02x", buf[i]);
    printf("\"}%s\n", last ? "" : ",");
}
#define DUMP(T) Dump<nrd::T>(#T)

int main()
{
    printf("{\n");
    printf("  \"enums\": {\"Denoiser\": %u, \"ResourceType\": %u, \"Format\": %u, \"Sampler\": %u, \"Result\": %u, \"DescriptorType\": %u, \"REFERENCE\": %u, "
           "\"SIGMA_SHADOW\": %u, \"RELAX_DIFFUSE_SPECULAR\": %u, \"PERMANENT_POOL\": %u, \"OUT_VALIDATION\": %u, \"R9_G9_B9_E5_UFLOAT\": %u, "
           "\"CLEAR_AND_RESTART\": %u, \"AREA_5X5\": %u, \"WHITE\": %u},\n",
           (unsigned)nrd::Denoiser::MAX_NUM, (unsigned)nrd::ResourceType::MAX_NUM, (unsigned)nrd::Format::MAX_NUM, (unsigned)nrd::Sampler::MAX_NUM,
           (unsigned)nrd::Result::MAX_NUM, (unsigned)nrd::DescriptorType::MAX_NUM, (unsigned)nrd::Denoiser::REFERENCE, (unsigned)nrd::Denoiser::SIGMA_SHADOW,
           (unsigned)nrd::Denoiser::RELAX_DIFFUSE_SPECULAR, (unsigned)nrd::ResourceType::PERMANENT_POOL, (unsigned)nrd::ResourceType::OUT_VALIDATION,
           (unsigned)nrd::Format::R9_G9_B9_E5_UFLOAT, (unsigned)nrd::AccumulationMode::CLEAR_AND_RESTART, (unsigned)nrd::HitDistanceReconstructionMode::AREA_5X5,
           (unsigned)nrd::CheckerboardMode::WHITE);
    printf("  \"version\": [%d, %d, %d],\n", NRD_VERSION_MAJOR, NRD_VERSION_MINOR, NRD_VERSION_BUILD);
    DUMP(AllocationCallbacks);
    DUMP(SPIRVBindingOffsets);
    DUMP(LibraryDesc);
    DUMP(DenoiserDesc);
    DUMP(InstanceCreationDesc);
    DUMP(TextureDesc);
    DUMP(ResourceDesc);
    DUMP(ResourceRangeDesc);
    DUMP(ComputeShaderDesc);
    DUMP(PipelineDesc);
    DUMP(DescriptorPoolDesc);
    DUMP(InstanceDesc);
    DUMP(DispatchDesc);
    DUMP(CommonSettings);
    DUMP(HitDistanceParameters);
    DUMP(ReblurAntilagSettings);
    DUMP(ReblurSettings);
    DUMP(RelaxAntilagSettings);
    DUMP(RelaxSettings);
    DUMP(SigmaSettings);
    Dump<nrd::ReferenceSettings>("ReferenceSettings", true);
    printf("}\n");
    return 0;
}
