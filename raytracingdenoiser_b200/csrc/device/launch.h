// Host-visible launchers of the sm_100a kernels (one per reference pass).  `signal`: 0 = diffuse, 1 = specular, 2 = both.
// Every kernel translation unit is compiled twice (common.cuh): NS = nrdb200 is the strip (multi-GPU) build,
// NS = nrdb200_single the one-GPU build.
#pragma once
#include "common.cuh"

#define NRD_B200_DECLARE_LAUNCHERS(NS)                                                                          \
    namespace NS                                                                                                \
    {                                                                                                           \
    cudaError_t LaunchReblurClassifyTiles(const nrdb200_abi::PassLaunch& p);                                    \
    cudaError_t LaunchReblurHitDistReconstruction(const nrdb200_abi::PassLaunch& p, int signal, bool is5x5);   \
    cudaError_t LaunchReblurPrePass(const nrdb200_abi::PassLaunch& p, int signal);                              \
    cudaError_t LaunchReblurTemporalAccumulation(const nrdb200_abi::PassLaunch& p, int signal);                 \
    cudaError_t LaunchReblurHistoryFix(const nrdb200_abi::PassLaunch& p, int signal);                           \
    cudaError_t LaunchReblurBlur(const nrdb200_abi::PassLaunch& p, int signal);                                 \
    cudaError_t LaunchReblurPostBlur(const nrdb200_abi::PassLaunch& p, int signal, bool noTemporalStabilization); \
    cudaError_t LaunchReblurTemporalStabilization(const nrdb200_abi::PassLaunch& p, int signal);                \
    cudaError_t LaunchSigma(const nrdb200_abi::PassLaunch& p, const char* shaderName);                          \
    cudaError_t LaunchRelax(const nrdb200_abi::PassLaunch& p, const char* shaderName);                          \
    cudaError_t LaunchAux(const nrdb200_abi::PassLaunch& p, const char* shaderName);                            \
    }

NRD_B200_DECLARE_LAUNCHERS(nrdb200)

#if !defined(NRD_B200_NO_STRIPS)
namespace nrdb200
{
// peer address table of one context slot, replicated into the constant memory of every kernel translation unit
cudaError_t SetPeerTableReblurSpatial(int slot, const nrdb200_abi::PeerTable* table);
cudaError_t SetPeerTableReblurHitDist(int slot, const nrdb200_abi::PeerTable* table);
cudaError_t SetPeerTableReblurTemporal(int slot, const nrdb200_abi::PeerTable* table);
cudaError_t SetPeerTableSigma(int slot, const nrdb200_abi::PeerTable* table);
cudaError_t SetPeerTableRelax(int slot, const nrdb200_abi::PeerTable* table);
cudaError_t SetPeerTableAux(int slot, const nrdb200_abi::PeerTable* table);
} // namespace nrdb200
#endif
