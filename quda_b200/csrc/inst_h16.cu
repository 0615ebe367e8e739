// Instantiates every Dslash / clover / pack kernel for storage precision PrecH16.
#include "kernels.cuh"

namespace b200
{
  template int launch_precision<PrecH16>(const LaunchRequest &);
  template int launch_clover_precision<PrecH16>(const CloverRequest &);
  template int launch_twist_precision<PrecH16>(const TwistRequest &);
  template int launch_pack_precision<PrecH16>(const PackRequest &);
  template int launch_pack_multi_precision<PrecH16>(const PackRequest &, const PackBatchRequest &);
  template int launch_copy_precision<PrecH16>(const CopyRequest &);
  template int launch_gauge_copy_precision<PrecH16>(const GaugeCopyRequest &);
  template int launch_clover_copy_precision<PrecH16>(const CloverCopyRequest &);
} // namespace b200
