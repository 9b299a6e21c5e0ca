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
} // namespace nrdb200
