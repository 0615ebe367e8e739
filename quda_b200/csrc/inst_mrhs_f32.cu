// Instantiates the multi-RHS Dslash kernels for storage precision PrecF32.
#include "mrhs.cuh"

namespace b200
{
  template int launch_mrhs_precision<PrecF32>(const MrhsRequest &);
} // namespace b200
