// Instantiates the multi-RHS Dslash kernels for storage precision PrecH16.
#include "mrhs.cuh"

namespace b200
{
  template int launch_mrhs_precision<PrecH16>(const MrhsRequest &);
} // namespace b200
