// Instantiates the TMA-staged marching Dslash kernels (tma_kernel.cuh) for storage precision PrecF32.
#include "tma_kernel.cuh"

namespace b200
{
  template int launch_tma_precision<PrecF32>(const LaunchRequest &);
} // namespace b200
