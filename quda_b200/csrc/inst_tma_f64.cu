// Instantiates the TMA-staged marching Dslash kernels (tma_kernel.cuh) for storage precision PrecF64.
#include "tma_kernel.cuh"

namespace b200
{
  template int launch_tma_precision<PrecF64>(const LaunchRequest &);
} // namespace b200
