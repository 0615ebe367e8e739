// Multi-RHS (batched) Dslash: kernel wrapper + per-precision launcher.  The site code is dslash_site_mrhs
// (dslash_site.h); one translation unit per storage precision includes this header (inst_mrhs_*.cu).
//
// Reference interface: the cvector_ref<ColorSpinorField> form of ApplyWilson / ApplyWilsonClover /
// ApplyWilsonCloverPreconditioned (include/dslash_quda.h:83-234), i.e. WilsonArg::out/in/x[MAX_MULTI_RHS] and the
// source index in the thread grid (include/kernels/dslash_wilson.cuh:37-69).
#pragma once

#include <cuda_runtime.h>
#include <type_traits>

#include "dslash_site.h"
#include "launch.h"

namespace b200
{

#ifndef B2_MAXTILE
#define B2_MAXTILE 128
#endif

  // Register budget: NS accumulators of 24 reals + the NS x 24-real spinor loads in flight.  2 CTAs of <= 128 threads
  // per SM gives ptxas the full 255 registers (fp64 NS = 2, fp32/half NS = 4); NS = 2 in fp32/half fits 168.
#ifndef B2_MRHS_MINBLOCKS_4
#define B2_MRHS_MINBLOCKS_4 2
#endif
#ifndef B2_MRHS_MINBLOCKS_2
#define B2_MRHS_MINBLOCKS_2 3
#endif
  template <class P, int NS> struct MrhsMinBlocks { static constexpr int value = NS >= 4 ? B2_MRHS_MINBLOCKS_4 : B2_MRHS_MINBLOCKS_2; };
  template <int NS> struct MrhsMinBlocks<PrecF64, NS> { static constexpr int value = 2; };

  template <class P, int recon, bool dagger, bool xpay, OpType op, int NS>
  __global__ void __launch_bounds__(B2_MAXTILE, MrhsMinBlocks<P, NS>::value)
    dslash_mrhs_kernel(const __grid_constant__ DslashArgs<P, recon> arg, const __grid_constant__ MrhsFields<P, NS> f,
                       const __grid_constant__ TileMap tm)
  {
    int x[4], x_cb, parity;
    if (!tile_site(x, x_cb, parity, arg.geom, tm, arg.n_parity, arg.parity, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x))
      return;
    dslash_site_mrhs<P, recon, dagger, xpay, op, NS>(arg, f, x, x_cb, parity);
  }

  // ---- CTA flavour: blockDim = (tile sites, sources per CTA); blockIdx.x = tile * n_batch + batch, so the CTAs that
  // share a tile's links are adjacent in launch order (L2 serves what L1 cannot)
  // Occupancy / register configurations of the CTA flavour (B200_MRHS_CTA_CFG): with the links shared through L1 / L2
  // the DRAM stream has head-room, so the kernel is bound by the per-thread load-latency chain and more resident warps
  // (fewer registers, no register-resident link preload) can pay:  cfg 0 = the single-source budget (128 regs, links
  // preloaded), cfg 1 = 6 CTAs of 128 threads (85 regs), cfg 2 = 8 CTAs (64 regs); fp64: 255 / 168 / 128 regs.
  template <class P, int CFG> struct MrhsCtaCfg {
    static constexpr int min_blocks = CFG == 0 ? 4 : (CFG == 1 ? 6 : 8);
    static constexpr bool preload = CFG == 0;
  };
  template <int CFG> struct MrhsCtaCfg<PrecF64, CFG> {
    static constexpr int min_blocks = CFG == 0 ? 2 : (CFG == 1 ? 3 : 4);
    static constexpr bool preload = false;
  };

  template <class P, int recon, bool dagger, bool xpay, OpType op, bool l1_links, int CFG>
  __global__ void __launch_bounds__(B2_MAXTILE, MrhsCtaCfg<P, CFG>::min_blocks)
    dslash_mrhs_cta_kernel(const __grid_constant__ DslashArgs<P, recon> arg, const __grid_constant__ MrhsViews<P> f,
                           const __grid_constant__ TileMap tm, int n_src, int n_batch)
  {
    const unsigned tile = blockIdx.x / (unsigned)n_batch;
    const int s = (int)(blockIdx.x - tile * n_batch) * blockDim.y + threadIdx.y;
    if (s >= n_src) return;
    int x[4], x_cb, parity;
    if (!tile_site(x, x_cb, parity, arg.geom, tm, arg.n_parity, arg.parity, tile, blockIdx.y, blockIdx.z, threadIdx.x)) return;
    dslash_site_src<P, recon, dagger, xpay, op, l1_links ? Cache::REUSE : Cache::STREAM, MrhsCtaCfg<P, CFG>::preload>(
      arg, f.in[s][1 - parity], f.out[s][parity], f.x[s][parity], x, x_cb, parity);
  }

  template <class P, int recon, bool dagger, bool xpay, OpType op>
  int launch_mrhs_cta(const MrhsRequest &rq, const DslashArgs<P, recon> &arg)
  {
    MrhsViews<P> f;
    fill_mrhs_views(f, rq);
    TileMap tm;
    int threads, gx, gy, gz, rc, nsb, n_batch;
    if (int e = make_tile_map(tm, threads, rq.base.tile, arg.geom, B2_MAXTILE)) return e;
    if (!mrhs_box(tm, rq, arg.comm_dim, arg.n_parity, gx, gy, gz, rc)) return rc;
    mrhs_cta_shape(nsb, n_batch, rq.n_src, threads, B2_MAXTILE, rq.cta_sources);
    if ((long long)gx * n_batch >= (1ll << 31)) return set_error(B200_ERR_INVALID, "lattice too large for the multi-RHS grid");
    const dim3 grid(gx * n_batch, gy, gz), block(threads, nsb, 1);
    cudaStream_t st = (cudaStream_t)rq.base.stream;
#define B2_CTA_LAUNCH(L1, CFG) \
  dslash_mrhs_cta_kernel<P, recon, dagger, xpay, op, L1, CFG><<<grid, block, 0, st>>>(arg, f, tm, rq.n_src, n_batch)
    if (!rq.l1_links)
      B2_CTA_LAUNCH(false, 0);
    else if (rq.cta_cfg == 1)
      B2_CTA_LAUNCH(true, 1);
    else if (rq.cta_cfg == 2)
      B2_CTA_LAUNCH(true, 2);
    else
      B2_CTA_LAUNCH(true, 0);
#undef B2_CTA_LAUNCH
    count_launch();
    return check_cuda(cudaGetLastError(), "multi-RHS (CTA) dslash launch");
  }

  template <class P, int recon, bool dagger, bool xpay, OpType op, int NS>
  int launch_mrhs_batch(const MrhsRequest &rq, const DslashArgs<P, recon> &arg, int s0)
  {
    MrhsFields<P, NS> f;
    fill_mrhs_fields(f, rq, s0);
    TileMap tm;
    int threads, gx, gy, gz, rc;
    if (int e = make_tile_map(tm, threads, rq.base.tile, arg.geom, B2_MAXTILE)) return e;
    if (!mrhs_box(tm, rq, arg.comm_dim, arg.n_parity, gx, gy, gz, rc)) return rc;
    dslash_mrhs_kernel<P, recon, dagger, xpay, op, NS><<<dim3(gx, gy, gz), threads, 0, (cudaStream_t)rq.base.stream>>>(arg, f, tm);
    count_launch();
    return check_cuda(cudaGetLastError(), "multi-RHS dslash launch");
  }

  template <class P, int recon, bool dagger, bool xpay, OpType op>
  int launch_mrhs_config(const MrhsRequest &rq, const DslashArgs<P, recon> &arg)
  {
    if (mrhs_mode(rq, P::bytes) == 1) return launch_mrhs_cta<P, recon, dagger, xpay, op>(rq, arg);
    int s0 = 0;
    while (s0 < rq.n_src) {
      const int ns = mrhs_batch<P>(rq.n_src - s0, rq.max_batch);
      int rc;
      if constexpr (!std::is_same<P, PrecF64>::value) {
        if (ns == 4) {
          if ((rc = launch_mrhs_batch<P, recon, dagger, xpay, op, 4>(rq, arg, s0))) return rc;
          s0 += 4;
          continue;
        }
      }
      if (ns >= 2) {
        if ((rc = launch_mrhs_batch<P, recon, dagger, xpay, op, 2>(rq, arg, s0))) return rc;
        s0 += 2;
      } else { // odd one out: the single-source kernel (inst_*.cu)
        LaunchRequest one = rq.base;
        one.out = rq.out[s0];
        one.in = rq.in[s0];
        if (rq.base.xpay) one.x = rq.x[s0];
        if ((rc = launch_precision<P>(one))) return rc;
        s0 += 1;
      }
    }
    return B200_SUCCESS;
  }

  template <class P, int recon> int launch_mrhs_recon(const MrhsRequest &rq)
  {
    DslashArgs<P, recon> arg;
    if (int rc = fill_args(arg, rq.base)) return rc;
    const bool xp = rq.base.xpay, dg = rq.base.dagger;
    switch (rq.base.op) {
    case OP_WILSON:
      if (dg)
        return xp ? launch_mrhs_config<P, recon, true, true, OP_WILSON>(rq, arg) :
                    launch_mrhs_config<P, recon, true, false, OP_WILSON>(rq, arg);
      else
        return xp ? launch_mrhs_config<P, recon, false, true, OP_WILSON>(rq, arg) :
                    launch_mrhs_config<P, recon, false, false, OP_WILSON>(rq, arg);
    case OP_CLOVER:
      if (!xp) return set_error(B200_ERR_INVALID, "ApplyWilsonClover exists in xpay form only (a != 0)");
      return dg ? launch_mrhs_config<P, recon, true, true, OP_CLOVER>(rq, arg) :
                  launch_mrhs_config<P, recon, false, true, OP_CLOVER>(rq, arg);
    case OP_CLOVER_PC:
      if (dg)
        return xp ? launch_mrhs_config<P, recon, true, true, OP_CLOVER_PC>(rq, arg) :
                    launch_mrhs_config<P, recon, true, false, OP_CLOVER_PC>(rq, arg);
      else
        return xp ? launch_mrhs_config<P, recon, false, true, OP_CLOVER_PC>(rq, arg) :
                    launch_mrhs_config<P, recon, false, false, OP_CLOVER_PC>(rq, arg);
    }
    return set_error(B200_ERR_INVALID, "unknown op %d", rq.base.op);
  }

  template <class P> int launch_mrhs_precision(const MrhsRequest &rq)
  {
    switch (rq.base.reconstruct) {
    case 18: return launch_mrhs_recon<P, 18>(rq);
    case 12: return launch_mrhs_recon<P, 12>(rq);
    case 8: return launch_mrhs_recon<P, 8>(rq);
    }
    return set_error(B200_ERR_INVALID, "reconstruct %d not in {18,12,8}", rq.base.reconstruct);
  }

} // namespace b200
