// The nine exported entry points of the drop-in boundary (reference: Include/NRD.h:51-70, Source/Wrapper.cpp:126-303).
#include "scheduler.h"

#include <cstdlib>

using namespace nrd;
using nrdb200::MemoryHooks;
using nrdb200::Scheduler;

namespace
{
// default allocator = aligned malloc (reference: Source/StdAllocator.h:105-113)
void* DefaultAllocate(void*, size_t size, size_t alignment)
{
    void* p = nullptr;
    if (alignment < sizeof(void*)) alignment = sizeof(void*);
    return posix_memalign(&p, alignment, size ? size : 1) == 0 ? p : nullptr;
}
void* DefaultReallocate(void*, void* memory, size_t size, size_t alignment)
{
    if (!memory) return DefaultAllocate(nullptr, size, alignment);
    void* p = realloc(memory, size);
    if (p && ((uintptr_t)p % (alignment ? alignment : 1)) != 0)
    {
        void* q = DefaultAllocate(nullptr, size, alignment);
        if (q) memcpy(q, p, size);
        free(p);
        p = q;
    }
    return p;
}
void DefaultFree(void*, void* memory) { free(memory); }

// Only the denoisers whose whole pass chain exists as CUDA kernels are advertised; CreateInstance returns
// Result::UNSUPPORTED for the rest exactly like the reference does for an unknown denoiser (InstanceImpl.cpp:110-117).
const Denoiser kSupported[] = {Denoiser::REBLUR_DIFFUSE, Denoiser::REBLUR_SPECULAR, Denoiser::REBLUR_DIFFUSE_SPECULAR,
                               Denoiser::RELAX_DIFFUSE, Denoiser::RELAX_SPECULAR, Denoiser::RELAX_DIFFUSE_SPECULAR, Denoiser::SIGMA_SHADOW, Denoiser::SIGMA_SHADOW_TRANSLUCENCY, Denoiser::REFERENCE};

const LibraryDesc kLibraryDesc = {{100, 200, 300, 400},
                                  kSupported,
                                  (uint32_t)(sizeof(kSupported) / sizeof(kSupported[0])),
                                  NRD_VERSION_MAJOR,
                                  NRD_VERSION_MINOR,
                                  NRD_VERSION_BUILD,
                                  NormalEncoding::R10_G10_B10_A2_UNORM, // NRD_NORMAL_ENCODING = 2 (reference CMakeLists.txt:28)
                                  RoughnessEncoding::LINEAR};           // NRD_ROUGHNESS_ENCODING = 1 (CMakeLists.txt:29)

#define NRD_B200_STR(name, ...) #name,
const char* const kResourceTypeNames[] = {NRD_B200_RESOURCE_TYPES(NRD_B200_STR)};
const char* const kDenoiserNames[] = {NRD_B200_DENOISERS(NRD_B200_STR)};
#undef NRD_B200_STR
} // namespace

NRD_API const LibraryDesc& NRD_CALL nrd::GetLibraryDesc() { return kLibraryDesc; }

NRD_API Result NRD_CALL nrd::CreateInstance(const InstanceCreationDesc& creationDesc, Instance*& instance)
{
    MemoryHooks hooks;
    hooks.cb = creationDesc.allocationCallbacks;
    if (!hooks.cb.Allocate || !hooks.cb.Reallocate || !hooks.cb.Free)
        hooks.cb = {DefaultAllocate, DefaultReallocate, DefaultFree, nullptr};

    void* memory = hooks.alloc(sizeof(Scheduler), alignof(Scheduler) < 16 ? 16 : alignof(Scheduler));
    if (!memory) return Result::FAILURE;
    Scheduler* scheduler = new (memory) Scheduler(hooks);

    Result result = scheduler->Create(creationDesc);
    if (result == Result::SUCCESS)
    {
        instance = (Instance*)scheduler;
        return Result::SUCCESS;
    }
    scheduler->~Scheduler();
    hooks.free(memory);
    return result;
}

NRD_API void NRD_CALL nrd::DestroyInstance(Instance& instance)
{
    Scheduler* scheduler = (Scheduler*)&instance;
    MemoryHooks hooks = scheduler->Hooks();
    scheduler->~Scheduler();
    hooks.free(scheduler);
}

NRD_API const InstanceDesc& NRD_CALL nrd::GetInstanceDesc(const Instance& instance) { return ((const Scheduler&)instance).GetDesc(); }

NRD_API Result NRD_CALL nrd::SetCommonSettings(Instance& instance, const CommonSettings& commonSettings)
{
    return ((Scheduler&)instance).SetCommonSettings(commonSettings);
}

NRD_API Result NRD_CALL nrd::SetDenoiserSettings(Instance& instance, Identifier identifier, const void* denoiserSettings)
{
    return ((Scheduler&)instance).SetDenoiserSettings(identifier, denoiserSettings);
}

NRD_API Result NRD_CALL nrd::GetComputeDispatches(Instance& instance, const Identifier* identifiers, uint32_t identifiersNum,
                                                 const DispatchDesc*& dispatchDescs, uint32_t& dispatchDescsNum)
{
    return ((Scheduler&)instance).GetComputeDispatches(identifiers, identifiersNum, dispatchDescs, dispatchDescsNum);
}

// Note: the reference's name table (Wrapper.cpp:58-95) is not in enum order (e.g. IN_DIFF_CONFIDENCE is enumerator 3 but
// name 12); that is a reference bug -- here the string is always the enumerator's own name.
NRD_API const char* NRD_CALL nrd::GetResourceTypeString(ResourceType resourceType)
{
    uint32_t i = (uint32_t)resourceType;
    return i < (uint32_t)ResourceType::MAX_NUM ? kResourceTypeNames[i] : nullptr;
}

NRD_API const char* NRD_CALL nrd::GetDenoiserString(Denoiser denoiser)
{
    uint32_t i = (uint32_t)denoiser;
    return i < (uint32_t)Denoiser::MAX_NUM ? kDenoiserNames[i] : nullptr;
}
