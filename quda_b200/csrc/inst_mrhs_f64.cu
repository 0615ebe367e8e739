// Instantiates the multi-RHS Dslash kernels for storage precision PrecF64.
#include "mrhs.cuh"

namespace b200
{
  template int launch_mrhs_precision<PrecF64>(const MrhsRequest &);
} // namespace b200
