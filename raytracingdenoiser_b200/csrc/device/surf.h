// Plain data shared by the executor and both builds of the kernels (one-GPU build and strip build, see common.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrdb200_abi
{
constexpr int kMaxPeers = 8;
constexpr int kMaxPeerSlots = 4;

// A pitched HBM surface.  On one GPU the whole texture is local.  In strip mode (several GPUs, one horizontal strip of
// `stripRows` rows each) a surface holds its own strip plus `halo` ghost rows above and below, which the owner of those
// rows refreshes after every pass that writes them (executor.cu, GhostPushKernel); a tap that lands outside even the
// ghost rows is loaded straight from the owner's HBM: every context carves its surfaces out of one arena with the same
// layout, so that address is the local address plus (arena of the owner - local arena).  Stores are always local.
struct Surf
{
    uint8_t* base;       // address of texel (0, ly0)
    int pitch;           // bytes per row
    int w, h;            // full (virtual) texture size
    int y0, y1;          // rows owned by this context: [y0, y1)  (the whole texture on one GPU)
    int ly0;             // first row held locally (y0 - halo in strip mode, may be negative)
    unsigned lrows;      // rows held locally
    unsigned stripRows;  // rows per strip in this texture's own units; 0 = whole frame is local
    unsigned stripMagic; // floor(2^32 / stripRows) + 1: owner(y) = umulhi(y, magic), exact for y, stripRows < 65536
    int halo;            // ghost rows on either side, in this texture's own units
    int peerSlot;
};

// Launch bookkeeping handed from the executor to the launchers
struct PassLaunch
{
    const void* constants; // host pointer to the dispatch's constant block
    uint32_t constantsSize;
    Surf tex[32];          // bindings in DispatchDesc order
    uint32_t texNum;
    int gridW, gridH;      // DispatchDesc grid (reference thread-group counts)
    int rowBegin, rowEnd;  // rows this launch must produce, in the pass's own pixel units
    cudaStream_t stream;
    bool preloadOnly;      // do not launch: only make the driver load the kernel this pass maps to (see NRD_B200_LAUNCH)
};
} // namespace nrdb200_abi
