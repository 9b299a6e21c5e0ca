// TMA (cp.async.bulk.tensor) plumbing for the dense-stencil kernels: a 2D tensor map over a pitched surface on the host, and the
// mbarrier / bulk-copy PTX on the device.  A CTA stages its tile + halo of a surface in shared memory with ONE instruction issued by
// one thread (UTMALDG in SASS); texels outside the surface arrive as zeros, kernels that need clamp-to-edge index the tile with
// clamped coordinates (the clamped texel is always inside the staged box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "surf.h"

namespace nrdb200_tma
{
// cuTensorMapEncodeTiled through the runtime's driver entry point query: libnrd_b200.so does not link libcuda
inline CUresult EncodeTiled(CUtensorMap* map, CUtensorMapDataType type, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                            const cuuint32_t* elementStrides)
{
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Fn fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return CUDA_ERROR_NOT_SUPPORTED;
        fn = (Fn)p;
    }
    return fn(map, type, rank, base, dims, strides, box, elementStrides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// Tensor map over the rows a context holds of a surface whose texels are `floatsPerTexel` 32-bit words (RGBA32F: 4).  Coordinates
// of a copy: x in 32-bit words (texel x * floatsPerTexel), y in local rows (row - surf.ly0).  Needs a 16-byte aligned base and pitch.
inline bool MakeSurfaceMap(const nrdb200_abi::Surf& s, int floatsPerTexel, int boxTexelsX, int boxRows, CUtensorMap* map)
{
    if (((uintptr_t)s.base & 15) != 0 || (s.pitch & 15) != 0 || boxTexelsX * floatsPerTexel > 256 || boxRows > 256) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)s.w * floatsPerTexel, (cuuint64_t)s.lrows};
    const cuuint64_t strides[1] = {(cuuint64_t)s.pitch};
    const cuuint32_t box[2] = {(cuuint32_t)(boxTexelsX * floatsPerTexel), (cuuint32_t)boxRows};
    const cuuint32_t es[2] = {1, 1};
    return EncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, s.base, dims, strides, box, es) == CUDA_SUCCESS;
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one thread, once per barrier; followed by __syncthreads()
__device__ __forceinline__ void BarrierInit(uint64_t* bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(SmemAddr(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: announce `bytes` and start the bulk tensor copy of the box whose first element is (x32, yLocal) into `dst`
__device__ __forceinline__ void LoadTile2D(void* dst, const CUtensorMap* map, int x32, int yLocal, uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(SmemAddr(dst)), "l"(map), "r"(x32), "r"(yLocal),
                 "r"(SmemAddr(bar))
                 : "memory");
}
// every thread that reads the tile
__device__ __forceinline__ void BarrierWait(uint64_t* bar, uint32_t phase)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(SmemAddr(bar)), "r"(phase) : "memory");
}
#endif
} // namespace nrdb200_tma
