// placeholder until the SIGMA / RELAX kernels land
#include "launch.h"
namespace nrdb200
{
cudaError_t LaunchSigma(const PassLaunch&, const char*) { return cudaErrorNotSupported; }
cudaError_t LaunchRelax(const PassLaunch&, const char*) { return cudaErrorNotSupported; }
}
