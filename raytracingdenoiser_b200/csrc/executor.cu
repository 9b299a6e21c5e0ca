// CUDA executor: owns the pool textures in HBM and turns DispatchDesc[] into sm_100a kernel launches.
// Takes the role of the reference's optional NRI integration layer (Integration/NRDIntegration.hpp: pool creation
// :292-363, Denoise :516-623, Dispatch :625-803).  Stream order replaces the SRV/UAV barriers (:667-704); constants
// travel as __grid_constant__ kernel parameters instead of a constant-buffer ring (:721-749).
#include "../../include/nrd_b200.h"
#include "device/launch.h"
#include "scheduler.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace nrd;
using namespace nrdb200;

namespace
{
std::atomic<uint64_t> g_launchCount{0};

uint32_t BytesPerTexel(Format f)
{
#define NRD_B200_BPT(name, bytes, isInt) bytes,
    static const uint32_t table[] = {NRD_B200_FORMATS(NRD_B200_BPT)};
#undef NRD_B200_BPT
    return (uint32_t)f < (uint32_t)Format::MAX_NUM ? table[(uint32_t)f] : 0;
}

struct Texture
{
    void* ptr = nullptr; // texel (0, firstRow)
    size_t pitch = 0;
    Format format = Format::R8_UNORM;
    uint16_t width = 0, height = 0; // virtual size
    uint16_t firstRow = 0, rows = 0;
    bool owned = false;
};
} // namespace

struct NrdCudaContext
{
    Instance* instance = nullptr;
    NrdCudaContextDesc desc{};
    std::vector<Texture> permanent, transient;
    Texture user[(size_t)ResourceType::MAX_NUM];
    std::string lastError;
};

namespace
{
// kernel the pool clears map to (reference: Clear_Float.cs / Clear_Uint.cs -- both write zeros)
Result Fail(NrdCudaContext* ctx, Result r, const std::string& msg)
{
    if (ctx) ctx->lastError = msg;
    return r;
}

void StripRows(const NrdCudaContextDesc& d, uint16_t downsample, uint16_t virtualHeight, uint16_t& first, uint16_t& rows)
{
    int y0 = (int)d.stripY0 - (int)d.haloRows, y1 = (int)d.stripY1 + (int)d.haloRows;
    if (y0 < 0) y0 = 0;
    if (y1 > (int)d.resourceHeight) y1 = d.resourceHeight;
    int f = y0 / downsample, l = (y1 + downsample - 1) / downsample;
    if (l > (int)virtualHeight) l = virtualHeight;
    first = (uint16_t)f;
    rows = (uint16_t)(l - f);
}

Result AllocatePool(NrdCudaContext* ctx, const TextureDesc* descs, uint32_t n, std::vector<Texture>& out)
{
    out.resize(n);
    for (uint32_t i = 0; i < n; i++)
    {
        Texture& t = out[i];
        const uint16_t ds = descs[i].downsampleFactor;
        t.format = descs[i].format;
        t.width = uint16_t((ctx->desc.resourceWidth + ds - 1) / ds);
        t.height = uint16_t((ctx->desc.resourceHeight + ds - 1) / ds);
        StripRows(ctx->desc, ds, t.height, t.firstRow, t.rows);
        size_t rowBytes = (size_t)t.width * BytesPerTexel(t.format);
        t.pitch = (rowBytes + 255) & ~(size_t)255; // 256-B aligned rows: 128-bit vector access and TMA-legal strides
        cudaError_t e = cudaMalloc(&t.ptr, t.pitch * t.rows);
        if (e != cudaSuccess) return Fail(ctx, Result::FAILURE, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        cudaMemset(t.ptr, 0, t.pitch * t.rows);
        t.owned = true;
    }
    return Result::SUCCESS;
}

const Texture* Resolve(NrdCudaContext* ctx, ResourceType type, uint32_t index)
{
    if (type == ResourceType::PERMANENT_POOL) return index < ctx->permanent.size() ? &ctx->permanent[index] : nullptr;
    if (type == ResourceType::TRANSIENT_POOL) return index < ctx->transient.size() ? &ctx->transient[index] : nullptr;
    if ((uint32_t)type < (uint32_t)ResourceType::MAX_NUM && ctx->user[(uint32_t)type].ptr) return &ctx->user[(uint32_t)type];
    return nullptr;
}

Surf ToSurf(const Texture& t)
{
    Surf s;
    s.base = (uint8_t*)t.ptr;
    s.pitch = (int)t.pitch;
    s.w = t.width;
    s.h = t.height;
    s.y0 = t.firstRow;
    s.y1 = t.firstRow + t.rows;
    return s;
}

// "REBLUR_DiffuseSpecular_Blur.cs" -> family REBLUR, signal 2, pass "Blur"
bool ParseReblur(const char* name, int& signal, const char*& pass)
{
    if (strncmp(name, "REBLUR_", 7) != 0) return false;
    const char* p = name + 7;
    if (!strncmp(p, "DiffuseSpecular_", 16)) { signal = 2; p += 16; }
    else if (!strncmp(p, "Diffuse_", 8)) { signal = 0; p += 8; }
    else if (!strncmp(p, "Specular_", 9)) { signal = 1; p += 9; }
    else return false;
    pass = p;
    return true;
}
} // namespace

namespace nrdb200
{
__global__ void ClearKernel(uint4* p, size_t n16)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) p[i] = make_uint4(0, 0, 0, 0);
}

// Clear_Float / Clear_Uint: zero every texel of the (strip of the) texture.  Rows are 256-B multiples, so the whole
// allocation is cleared with 128-bit stores.
cudaError_t LaunchClear(const PassLaunch& p)
{
    const Surf& s = p.tex[0];
    size_t bytes = (size_t)s.pitch * (size_t)(s.y1 - s.y0);
    if ((s.pitch & 15) != 0 || ((uintptr_t)s.base & 15) != 0) return cudaMemsetAsync(s.base, 0, bytes, p.stream); // user textures with odd pitch
    size_t n16 = bytes / 16;
    int blocks = (int)((n16 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    ClearKernel<<<blocks, 256, 0, p.stream>>>((uint4*)s.base, n16);
    return cudaGetLastError();
}
} // namespace nrdb200

extern "C" {

NRD_API Result nrdCudaCreateContext(Instance* instance, const NrdCudaContextDesc* desc, NrdCudaContext** out)
{
    if (!instance || !desc || !out) return Result::INVALID_ARGUMENT;
    if (!desc->resourceWidth || !desc->resourceHeight || desc->stripY1 <= desc->stripY0 || desc->stripY1 > desc->resourceHeight) return Result::INVALID_ARGUMENT;
    int deviceCount = 0;
    if (cudaGetDeviceCount(&deviceCount) != cudaSuccess || deviceCount == 0) return Result::FAILURE; // no silent CPU path: fail loudly
    if (cudaSetDevice(desc->device) != cudaSuccess) return Result::FAILURE;

    NrdCudaContext* ctx = new NrdCudaContext();
    ctx->instance = instance;
    ctx->desc = *desc;
    const InstanceDesc& id = GetInstanceDesc(*instance);
    Result r = AllocatePool(ctx, id.permanentPool, id.permanentPoolSize, ctx->permanent);
    if (r == Result::SUCCESS) r = AllocatePool(ctx, id.transientPool, id.transientPoolSize, ctx->transient);
    if (r != Result::SUCCESS)
    {
        nrdCudaDestroyContext(ctx);
        return r;
    }
    *out = ctx;
    return Result::SUCCESS;
}

NRD_API void nrdCudaDestroyContext(NrdCudaContext* ctx)
{
    if (!ctx) return;
    for (Texture& t : ctx->permanent)
        if (t.owned) cudaFree(t.ptr);
    for (Texture& t : ctx->transient)
        if (t.owned) cudaFree(t.ptr);
    delete ctx;
}

NRD_API Result nrdCudaSetUserTexture(NrdCudaContext* ctx, uint32_t resourceType, void* devicePtr, size_t pitchBytes, uint32_t format)
{
    if (!ctx || resourceType >= (uint32_t)ResourceType::TRANSIENT_POOL || format >= (uint32_t)Format::MAX_NUM) return Result::INVALID_ARGUMENT;
    Format expected;
    switch ((ResourceType)resourceType)
    {
        case ResourceType::IN_MV: expected = Format::RGBA16_SFLOAT; break;
        case ResourceType::IN_NORMAL_ROUGHNESS: expected = Format::R10_G10_B10_A2_UNORM; break;
        case ResourceType::IN_VIEWZ: expected = Format::R32_SFLOAT; break;
        case ResourceType::IN_DIFF_RADIANCE_HITDIST:
        case ResourceType::IN_SPEC_RADIANCE_HITDIST:
        case ResourceType::OUT_DIFF_RADIANCE_HITDIST:
        case ResourceType::OUT_SPEC_RADIANCE_HITDIST: expected = Format::RGBA16_SFLOAT; break;
        case ResourceType::IN_PENUMBRA: expected = Format::R16_SFLOAT; break;
        case ResourceType::OUT_SHADOW_TRANSLUCENCY: expected = Format::R8_UNORM; break;
        default: return Fail(ctx, Result::UNSUPPORTED, "resource type not consumed by the supported denoisers");
    }
    if ((Format)format != expected) return Fail(ctx, Result::UNSUPPORTED, "unsupported format for this resource type (see nrd_b200.h)");
    Texture& t = ctx->user[resourceType];
    t.ptr = devicePtr;
    t.pitch = pitchBytes;
    t.format = (Format)format;
    t.width = ctx->desc.resourceWidth;
    t.height = ctx->desc.resourceHeight;
    StripRows(ctx->desc, 1, t.height, t.firstRow, t.rows);
    t.owned = false;
    return Result::SUCCESS;
}

NRD_API Result nrdCudaGetTexture(NrdCudaContext* ctx, uint32_t resourceType, uint32_t indexInPool, NrdCudaTextureInfo* info)
{
    if (!ctx || !info) return Result::INVALID_ARGUMENT;
    const Texture* t = Resolve(ctx, (ResourceType)resourceType, indexInPool);
    if (!t) return Result::INVALID_ARGUMENT;
    info->devicePtr = t->ptr;
    info->pitchBytes = t->pitch;
    info->format = (uint32_t)t->format;
    info->width = t->width;
    info->height = t->height;
    info->firstRow = t->firstRow;
    info->rowsNum = t->rows;
    return Result::SUCCESS;
}

NRD_API Result nrdCudaExecuteDispatch(NrdCudaContext* ctx, const DispatchDesc* d, void* stream)
{
    if (!ctx || !d) return Result::INVALID_ARGUMENT;
    const InstanceDesc& id = GetInstanceDesc(*ctx->instance);
    if (d->pipelineIndex >= id.pipelinesNum) return Result::INVALID_ARGUMENT;
    const char* shader = id.pipelines[d->pipelineIndex].shaderFileName;
    const CommonSettings& cs = ((Scheduler*)ctx->instance)->Common();
    if (cs.rectSize[0] != cs.resourceSize[0] || cs.rectSize[1] != cs.resourceSize[1] || cs.rectOrigin[0] || cs.rectOrigin[1] ||
        cs.resourceSize[0] != ctx->desc.resourceWidth || cs.resourceSize[1] != ctx->desc.resourceHeight)
        return Fail(ctx, Result::UNSUPPORTED, "dynamic resolution (rectSize != resourceSize) is not implemented by the CUDA executor");
    if (cs.isHistoryConfidenceAvailable || cs.isDisocclusionThresholdMixAvailable || cs.isBaseColorMetalnessAvailable)
        return Fail(ctx, Result::UNSUPPORTED, "confidence / disocclusion-mix / base-colour inputs are not implemented by the CUDA executor");

    PassLaunch p{};
    p.constants = d->constantBufferData;
    p.constantsSize = d->constantBufferDataSize;
    p.texNum = d->resourcesNum;
    p.gridW = d->gridWidth;
    p.gridH = d->gridHeight;
    p.stream = (cudaStream_t)stream;
    if (d->resourcesNum > 32) return Fail(ctx, Result::FAILURE, "too many resources in one dispatch");
    for (uint32_t i = 0; i < d->resourcesNum; i++)
    {
        const ResourceDesc& r = d->resources[i];
        const Texture* t = Resolve(ctx, r.type, r.indexInPool);
        if (!t) return Fail(ctx, Result::INVALID_ARGUMENT, std::string("unbound resource ") + GetResourceTypeString(r.type) + " for " + d->name);
        p.tex[i] = ToSurf(*t);
    }
    // rows to produce: the context's strip plus halo (full frame on one GPU)
    uint16_t first, rows;
    StripRows(ctx->desc, 1, ctx->desc.resourceHeight, first, rows);
    p.rowBegin = first;
    p.rowEnd = first + rows;

    cudaError_t e = cudaErrorNotSupported;
    int signal = 0;
    const char* pass = nullptr;
    if (!strncmp(shader, "Clear_", 6)) e = LaunchClear(p);
    else if (!strcmp(shader, "REBLUR_ClassifyTiles.cs")) e = LaunchReblurClassifyTiles(p);
    else if (ParseReblur(shader, signal, pass))
    {
        if (!strcmp(pass, "PrePass.cs")) e = LaunchReblurPrePass(p, signal);
        else if (!strcmp(pass, "TemporalAccumulation.cs")) e = LaunchReblurTemporalAccumulation(p, signal);
        else if (!strcmp(pass, "HistoryFix.cs")) e = LaunchReblurHistoryFix(p, signal);
        else if (!strcmp(pass, "Blur.cs")) e = LaunchReblurBlur(p, signal);
        else if (!strcmp(pass, "PostBlur.cs")) e = LaunchReblurPostBlur(p, signal, false);
        else if (!strcmp(pass, "PostBlur_NoTemporalStabilization.cs")) e = LaunchReblurPostBlur(p, signal, true);
        else if (!strcmp(pass, "TemporalStabilization.cs")) e = LaunchReblurTemporalStabilization(p, signal);
    }
    else if (!strncmp(shader, "SIGMA_", 6)) e = LaunchSigma(p, shader);
    else if (!strncmp(shader, "RELAX_", 6)) e = LaunchRelax(p, shader);

    if (e == cudaErrorNotSupported) return Fail(ctx, Result::UNSUPPORTED, std::string("no CUDA kernel for pass ") + shader);
    if (e != cudaSuccess) return Fail(ctx, Result::FAILURE, std::string(shader) + ": " + cudaGetErrorString(e));
    g_launchCount.fetch_add(1, std::memory_order_relaxed);
    return Result::SUCCESS;
}

NRD_API Result nrdCudaDenoise(NrdCudaContext* ctx, const Identifier* identifiers, uint32_t identifiersNum, void* stream, uint32_t* launches)
{
    if (!ctx) return Result::INVALID_ARGUMENT;
    const DispatchDesc* dispatches = nullptr;
    uint32_t n = 0;
    Result r = GetComputeDispatches(*ctx->instance, identifiers, identifiersNum, dispatches, n);
    if (r != Result::SUCCESS) return r;
    for (uint32_t i = 0; i < n; i++)
    {
        r = nrdCudaExecuteDispatch(ctx, &dispatches[i], stream);
        if (r != Result::SUCCESS) return r;
    }
    if (launches) *launches = n;
    return Result::SUCCESS;
}

NRD_API Result nrdCudaUploadTexture(NrdCudaContext* ctx, uint32_t resourceType, uint32_t indexInPool, const void* hostPtr, size_t hostPitchBytes)
{
    if (!ctx || !hostPtr) return Result::INVALID_ARGUMENT;
    const Texture* t = Resolve(ctx, (ResourceType)resourceType, indexInPool);
    if (!t) return Result::INVALID_ARGUMENT;
    cudaError_t e = cudaMemcpy2D(t->ptr, t->pitch, hostPtr, hostPitchBytes, (size_t)t->width * BytesPerTexel(t->format), t->rows, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? Result::SUCCESS : Fail(ctx, Result::FAILURE, cudaGetErrorString(e));
}

NRD_API Result nrdCudaDownloadTexture(NrdCudaContext* ctx, uint32_t resourceType, uint32_t indexInPool, void* hostPtr, size_t hostPitchBytes)
{
    if (!ctx || !hostPtr) return Result::INVALID_ARGUMENT;
    const Texture* t = Resolve(ctx, (ResourceType)resourceType, indexInPool);
    if (!t) return Result::INVALID_ARGUMENT;
    cudaError_t e = cudaMemcpy2D(hostPtr, hostPitchBytes, t->ptr, t->pitch, (size_t)t->width * BytesPerTexel(t->format), t->rows, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? Result::SUCCESS : Fail(ctx, Result::FAILURE, cudaGetErrorString(e));
}

NRD_API const char* nrdCudaGetLastError(NrdCudaContext* ctx) { return ctx ? ctx->lastError.c_str() : ""; }
NRD_API uint64_t nrdCudaGetLaunchCount() { return g_launchCount.load(); }
}
