// placeholder until the RELAX kernels land
#include "launch.h"
namespace nrdb200
{
cudaError_t LaunchRelax(const PassLaunch&, const char*) { return cudaErrorNotSupported; }
}
