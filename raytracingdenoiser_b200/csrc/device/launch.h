// Host-visible launchers of the sm_100a kernels (one per reference pass).  `signal`: 0 = diffuse, 1 = specular, 2 = both.
#pragma once
#include "common.cuh"

namespace nrdb200
{
cudaError_t LaunchClear(const PassLaunch& p);

cudaError_t LaunchReblurClassifyTiles(const PassLaunch& p);
cudaError_t LaunchReblurPrePass(const PassLaunch& p, int signal);
cudaError_t LaunchReblurTemporalAccumulation(const PassLaunch& p, int signal);
cudaError_t LaunchReblurHistoryFix(const PassLaunch& p, int signal);
cudaError_t LaunchReblurBlur(const PassLaunch& p, int signal);
cudaError_t LaunchReblurPostBlur(const PassLaunch& p, int signal, bool noTemporalStabilization);
cudaError_t LaunchReblurTemporalStabilization(const PassLaunch& p, int signal);

cudaError_t LaunchSigma(const PassLaunch& p, const char* shaderName);
cudaError_t LaunchRelax(const PassLaunch& p, const char* shaderName);

// peer address table of one context slot, replicated into the constant memory of every kernel translation unit
cudaError_t SetPeerTableReblurSpatial(int slot, const long long* delta);
cudaError_t SetPeerTableReblurTemporal(int slot, const long long* delta);
cudaError_t SetPeerTableSigma(int slot, const long long* delta);
cudaError_t SetPeerTableRelax(int slot, const long long* delta);
} // namespace nrdb200
