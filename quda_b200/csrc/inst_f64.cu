// Instantiates every Dslash / clover / pack kernel for storage precision PrecF64.
#include "kernels.cuh"

namespace b200
{
  template int launch_precision<PrecF64>(const LaunchRequest &);
  template int launch_clover_precision<PrecF64>(const CloverRequest &);
  template int launch_twist_precision<PrecF64>(const TwistRequest &);
  template int launch_pack_precision<PrecF64>(const PackRequest &);
  template int launch_pack_multi_precision<PrecF64>(const PackRequest &, const PackBatchRequest &);
  template int launch_copy_precision<PrecF64>(const CopyRequest &);
  template int launch_gauge_copy_precision<PrecF64>(const GaugeCopyRequest &);
  template int launch_clover_copy_precision<PrecF64>(const CloverCopyRequest &);
} // namespace b200
