// TMA (cp.async.bulk.tensor) plumbing for the dense-stencil kernels: a 2D tensor map over a pitched surface on the host, and the
// mbarrier / bulk-copy PTX on the device.  A CTA stages its tile + halo of a surface in shared memory with ONE instruction issued by
// one thread (UTMALDG in SASS); texels outside the surface arrive as zeros, kernels that need clamp-to-edge index the tile with
// clamped coordinates (the clamped texel is always inside the staged box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "surf.h"

namespace nrdb200_tma
{
// NRD_B200_NO_TMA=1 in the environment makes every kernel stage its tiles with plain clamped loads (A/B and triage)
inline bool Enabled()
{
    static const bool on = getenv("NRD_B200_NO_TMA") == nullptr;
    return on;
}
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

// Tensor map over the rows a context holds of a surface of 2-byte texels (R16F).  Coordinates of a copy: x in texels, y in local rows.
// The box width must make whole 16-byte rows in shared memory (a multiple of 8 texels), and the x of every copy must be a multiple
// of 8 texels too: the first byte of a box row has to be 16-byte aligned in global memory (measured: anything else raises
// "illegal instruction" at the UTMALDG, profiles/r2_tma_probe.txt).
inline bool MakeSurfaceMap16(const nrdb200_abi::Surf& s, int boxTexelsX, int boxRows, CUtensorMap* map)
{
    if (((uintptr_t)s.base & 15) != 0 || (s.pitch & 15) != 0 || (boxTexelsX & 7) != 0 || boxTexelsX > 256 || boxRows > 256) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)s.w, (cuuint64_t)s.lrows};
    const cuuint64_t strides[1] = {(cuuint64_t)s.pitch};
    const cuuint32_t box[2] = {(cuuint32_t)boxTexelsX, (cuuint32_t)boxRows};
    const cuuint32_t es[2] = {1, 1};
    return EncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, s.base, dims, strides, box, es) == CUDA_SUCCESS;
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
// one thread: the single arrival of the barrier, announcing the bytes of ALL copies that will complete on it
__device__ __forceinline__ void BarrierExpect(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
// one thread: start the bulk tensor copy of the box whose first element is (x, yLocal) (in elements of the map) into `dst`
__device__ __forceinline__ void IssueTile2D(void* dst, const CUtensorMap* map, int x, int yLocal, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(SmemAddr(dst)), "l"(map), "r"(x), "r"(yLocal),
                 "r"(SmemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void LoadTile2D(void* dst, const CUtensorMap* map, int x, int yLocal, uint64_t* bar, uint32_t bytes)
{
    BarrierExpect(bar, bytes);
    IssueTile2D(dst, map, x, yLocal, bar);
}
// Clamp-to-edge for a staged box whose cell (lx, ly) holds texel (boxX0 + lx, boxY0 + ly): cells outside [0, maxX] x [0, maxY]
// (zero-filled by TMA) take the value of the clamped texel, which is always inside the box.  Only CTAs on the frame edge do
// anything; callers __syncthreads() afterwards.  Readers and writers are disjoint cells, so one pass is enough.
template <class T, int BW, int BH> __device__ __forceinline__ void PatchClampToEdge(T (*tile)[BW], int boxX0, int boxY0, int maxX, int maxY, int tid, int threads)
{
    if (boxX0 >= 0 && boxY0 >= 0 && boxX0 + BW - 1 <= maxX && boxY0 + BH - 1 <= maxY) return;
    for (int i = tid; i < BW * BH; i += threads)
    {
        const int lx = i % BW, ly = i / BW, gx = boxX0 + lx, gy = boxY0 + ly;
        const int cx = min(max(gx, 0), maxX), cy = min(max(gy, 0), maxY);
        // (a box that lies entirely beyond the frame -- a CTA past the rect under dynamic resolution -- has no clamped texel inside it: skipped)
        if ((cx != gx || cy != gy) && (unsigned)(cx - boxX0) < (unsigned)BW && (unsigned)(cy - boxY0) < (unsigned)BH) tile[ly][lx] = tile[cy - boxY0][cx - boxX0];
    }
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
