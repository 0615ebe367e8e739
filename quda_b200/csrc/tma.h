// TMA-staged, time-marching interior kernel: plan, tensor-map descriptions, per-thread shared-memory offsets and the
// site arithmetic.  Everything here is `B2_HD` so that the test-only host twin can run the SAME plan / offset / site code
// against an emulated shared memory filled by emulated box loads (tests/hosttwin) -- the CUDA-only parts (mbarrier,
// cp.async.bulk.tensor) live in tma_kernel.cuh.
//
// What it replaces: the reference's interior kernel gathers the 8 neighbour spinors straight from global memory and
// leaves their reuse to L1/L2 (/root/reference/include/kernels/dslash_wilson.cuh:84-161); on B200 that makes the
// SM <- L2 fill path, not HBM, the limiter for fp32 recon-12.  Here a persistent CTA owns an
// (all x) x TY x TZ tile of one parity and marches through t:
//   * the input-spinor slice t+1 (tile rows only) and the y/z halo rows of slice t arrive in shared memory by TMA box
//     loads, one slice ahead of their use; slices t-1, t, t+1 stay resident, so every input site crosses the
//     L2 -> SM path (1 + halo) times instead of ~4 times;
//   * the 8 links of a site arrive the same way, one direction pair (forward + backward link of dimension d) per
//     pipeline stage, so no thread ever waits on a global load: all stencil operands are LDS;
//   * a producer warp issues the box loads, the consumer warps wait on mbarriers -- no __syncthreads in the loop.
// Work is the linearised (tile, t) sequence cut into gridDim.x equal ranges (148 SMs do not divide 2^k tiles).
#pragma once

#include "dslash_site.h"

namespace b200
{

  constexpr int kTmaCenterSlots = 4; // slices t-1, t, t+1 live + one in flight
  constexpr int kTmaHaloSlots = 2;   // halo rows of slice t live + one in flight
  constexpr int kTmaMaxLinkSlots = 4;
  constexpr int kTmaSmemBudget = 227 * 1024;
  constexpr int kTmaMaxConsumers = 256;

  // box shapes; one tensor map per (output parity, shape)
  enum TmaMapId {
    TM_SC = 0,  // spinor tile rows      (TY, TZ)
    TM_SY = 1,  // spinor y-halo rows    (1, TZ)
    TM_SZ = 2,  // spinor z-halo rows    (TY, 1)
    TM_GF = 3,  // forward links, output parity           (TY, TZ)
    TM_GB = 4,  // backward links x / t, other parity     (TY, TZ)
    TM_GYA = 5, // backward y links, row y0-1             (1, TZ)
    TM_GYB = 6, // backward y links, rows y0 .. y0+TY-2   (TY-1, TZ)
    TM_GZA = 7, // backward z links, row z0-1             (TY, 1)
    TM_GZB = 8, // backward z links, rows z0 .. z0+TZ-2   (TY, TZ-1)
    TM_COUNT = 9
  };

  // Description of one tiled tensor map over a native field: 5-d (u32 lane within an x row, plane, y, z, t).
  // The device launcher turns it into a CUtensorMap; the host twin interprets it directly.
  struct TmaDesc {
    const void *base;
    unsigned long long dim[5];    // elements (u32) / planes / sites
    unsigned long long stride[5]; // bytes; stride[0] = 4
    unsigned box[5];
    int valid;                    // 0: shape has a zero extent (TY == 1 or TZ == 1), never issued
  };

  struct TmaPlan {
    int Xh, Y, Z, T;
    int TY, TZ, nty, ntz;
    int n_parity, parity;
    int n_items;      // n_parity * nty * ntz * T
    int n_consumers;  // Xh * TY * TZ
    int n_cwarps;     // consumer warps
    int svec, gvec;   // bytes per spinor / gauge vector
    int SP, GP;       // planes per spinor / per link
    int srow, grow;   // bytes of one (y,z) row record: SP * svec * Xh, GP * gvec * Xh
    int NC, NH;       // tile rows, halo rows
    int center_bytes, halo_bytes, link_bytes; // per slot (128-byte aligned)
    int n_link_slots;
    int off_center, off_halo, off_link, off_bar, smem_bytes;
  };

  B2_HD int tma_align128(int v) { return (v + 127) & ~127; }

  // mbarrier slots inside the barrier block (8 bytes each)
  B2_HD int tma_bar_full_c(int s) { return s; }
  B2_HD int tma_bar_empty_c(int s) { return kTmaCenterSlots + s; }
  B2_HD int tma_bar_full_h(int s) { return 2 * kTmaCenterSlots + s; }
  B2_HD int tma_bar_empty_h(int s) { return 2 * kTmaCenterSlots + kTmaHaloSlots + s; }
  B2_HD int tma_bar_full_l(int s) { return 2 * kTmaCenterSlots + 2 * kTmaHaloSlots + s; }
  B2_HD int tma_bar_empty_l(int s) { return 2 * kTmaCenterSlots + 2 * kTmaHaloSlots + kTmaMaxLinkSlots + s; }
  constexpr int kTmaBarriers = 2 * kTmaCenterSlots + 2 * kTmaHaloSlots + 2 * kTmaMaxLinkSlots;

  // Build the plan for a lattice / precision / reconstruct.  Returns false if this shape is not served by the TMA
  // kernel (the caller falls back to the gather kernel): x rows must fit one box, the tile must divide the lattice and
  // spinor slices + >= 2 link stages must fit in shared memory.
  template <class P, int recon> B2_HD bool tma_make_plan(TmaPlan &p, const Geom &g, int n_parity, int parity, int want_ty, int want_tz)
  {
    using GV = GaugeView<P, recon>;
    p.Xh = g.Xh0;
    p.Y = g.X[1];
    p.Z = g.X[2];
    p.T = g.X[3];
    p.n_parity = n_parity;
    p.parity = parity;
    p.svec = (int)sizeof(typename P::svec);
    p.gvec = (int)sizeof(typename GV::V);
    p.SP = 24 / P::Ns;
    p.GP = GV::M;
    if (P::fixed) return false;                                  // block-float norms would need their own maps
    if (p.svec * p.Xh / 4 > 256 || p.gvec * p.Xh / 4 > 256) return false; // box extent limit
    if ((p.svec * p.Xh) % 16 || (p.gvec * p.Xh) % 16) return false;      // inner box bytes / strides: multiples of 16
    if (p.T < 3) return false;                                   // t-1, t, t+1 must be distinct slices
    // tile: powers of two dividing Y and Z, grown alternately (y first) up to the consumer budget
    const int budget = sizeof(typename P::real) == 8 ? 128 : kTmaMaxConsumers;
    int ty = 1, tz = 1;
    if (want_ty > 0 && want_tz > 0) {
      ty = want_ty;
      tz = want_tz;
      if (p.Y % ty || p.Z % tz) return false;
    } else {
      for (;;) {
        bool grown = false;
        if (ty <= tz && p.Y % (2 * ty) == 0 && p.Xh * 2 * ty * tz <= budget) { ty *= 2; grown = true; }
        else if (p.Z % (2 * tz) == 0 && p.Xh * ty * 2 * tz <= budget) { tz *= 2; grown = true; }
        else if (p.Y % (2 * ty) == 0 && p.Xh * 2 * ty * tz <= budget) { ty *= 2; grown = true; }
        if (!grown) break;
      }
    }
    if (p.Xh * ty * tz > kTmaMaxConsumers) return false;
    p.TY = ty;
    p.TZ = tz;
    p.nty = p.Y / ty;
    p.ntz = p.Z / tz;
    p.n_items = n_parity * p.nty * p.ntz * p.T;
    p.n_consumers = p.Xh * ty * tz;
    p.n_cwarps = (p.n_consumers + 31) / 32;
    p.srow = p.SP * p.svec * p.Xh;
    p.grow = p.GP * p.gvec * p.Xh;
    p.NC = ty * tz;
    p.NH = 2 * tz + 2 * ty;
    p.center_bytes = tma_align128(p.NC * p.srow);
    // halo regions (ym, yp, zm, zp) each start 128-byte aligned
    p.halo_bytes = 2 * tma_align128(tz * p.srow) + 2 * tma_align128(ty * p.srow);
    // link stage: forward block, then the backward block (two boxes for y and z: region A, region B)
    // (region A holds max(TY,TZ) rows, region B the remaining NC - min(TY,TZ); the x / t boxes of NC rows start at A)
    {
      const int mx = ty > tz ? ty : tz, mn = ty > tz ? tz : ty;
      int bwd = tma_align128(mx * p.grow) + tma_align128((p.NC - mn) * p.grow);
      if (bwd < tma_align128(p.NC * p.grow)) bwd = tma_align128(p.NC * p.grow);
      p.link_bytes = tma_align128(p.NC * p.grow) + bwd;
    }
    const int fixed_bytes = kTmaCenterSlots * p.center_bytes + kTmaHaloSlots * p.halo_bytes + 1024;
    int nl = (kTmaSmemBudget - fixed_bytes) / p.link_bytes;
    if (nl > kTmaMaxLinkSlots) nl = kTmaMaxLinkSlots;
    if (nl < 2) return false;
    p.n_link_slots = nl;
    p.off_center = 0;
    p.off_halo = p.off_center + kTmaCenterSlots * p.center_bytes;
    p.off_link = p.off_halo + kTmaHaloSlots * p.halo_bytes;
    p.off_bar = p.off_link + nl * p.link_bytes;
    p.smem_bytes = p.off_bar + 1024;
    return true;
  }

  // byte offsets of the regions inside a halo slot / a link stage
  B2_HD int tma_halo_ym(const TmaPlan &) { return 0; }
  B2_HD int tma_halo_yp(const TmaPlan &p) { return tma_align128(p.TZ * p.srow); }
  B2_HD int tma_halo_zm(const TmaPlan &p) { return 2 * tma_align128(p.TZ * p.srow); }
  B2_HD int tma_halo_zp(const TmaPlan &p) { return 2 * tma_align128(p.TZ * p.srow) + tma_align128(p.TY * p.srow); }
  B2_HD int tma_link_bwd_a(const TmaPlan &p) { return tma_align128(p.NC * p.grow); }
  B2_HD int tma_link_bwd_b(const TmaPlan &p)
  {
    return tma_align128(p.NC * p.grow) + tma_align128((p.TY > p.TZ ? p.TY : p.TZ) * p.grow);
  }

  // ---- tensor-map descriptions ------------------------------------------------------------------------------------
  // spinor field of one parity: (u32 lane, plane, y, z, t)
  B2_HD void tma_desc_field(TmaDesc &d, const void *base, int vec_bytes, int planes, size_t plane_stride_sites, const TmaPlan &p,
                            int box_planes, int by, int bz)
  {
    d.base = base;
    d.dim[0] = (unsigned long long)(vec_bytes / 4) * p.Xh;
    d.dim[1] = planes;
    d.dim[2] = p.Y;
    d.dim[3] = p.Z;
    d.dim[4] = p.T;
    d.stride[0] = 4;
    d.stride[1] = (unsigned long long)vec_bytes * plane_stride_sites;
    d.stride[2] = (unsigned long long)vec_bytes * p.Xh;
    d.stride[3] = d.stride[2] * p.Y;
    d.stride[4] = d.stride[3] * p.Z;
    d.box[0] = (unsigned)d.dim[0];
    d.box[1] = box_planes;
    d.box[2] = by;
    d.box[3] = bz;
    d.box[4] = 1;
    d.valid = (by > 0 && bz > 0) ? 1 : 0;
  }

  // all maps for output parity `par` (input spinor and backward links have parity 1 - par)
  template <class P, int recon> B2_HD void tma_make_descs(TmaDesc *d, const DslashArgs<P, recon> &arg, const TmaPlan &p, int par)
  {
    const void *in = arg.in[1 - par].v;
    const size_t sst = arg.in[1 - par].stride;
    tma_desc_field(d[TM_SC], in, p.svec, p.SP, sst, p, p.SP, p.TY, p.TZ);
    tma_desc_field(d[TM_SY], in, p.svec, p.SP, sst, p, p.SP, 1, p.TZ);
    tma_desc_field(d[TM_SZ], in, p.svec, p.SP, sst, p, p.SP, p.TY, 1);
    const void *gf = arg.U.g[par], *gb = arg.U.g[1 - par];
    const size_t gst = arg.U.stride;
    tma_desc_field(d[TM_GF], gf, p.gvec, 4 * p.GP, gst, p, p.GP, p.TY, p.TZ);
    tma_desc_field(d[TM_GB], gb, p.gvec, 4 * p.GP, gst, p, p.GP, p.TY, p.TZ);
    tma_desc_field(d[TM_GYA], gb, p.gvec, 4 * p.GP, gst, p, p.GP, 1, p.TZ);
    tma_desc_field(d[TM_GYB], gb, p.gvec, 4 * p.GP, gst, p, p.GP, p.TY - 1, p.TZ);
    tma_desc_field(d[TM_GZA], gb, p.gvec, 4 * p.GP, gst, p, p.GP, p.TY, 1);
    tma_desc_field(d[TM_GZB], gb, p.gvec, 4 * p.GP, gst, p, p.GP, p.TY, p.TZ - 1);
  }

  // a tensor map is only usable if the hardware constraints hold (cuTensorMapEncodeTiled): 16-byte aligned base and strides
  B2_HD bool tma_desc_ok(const TmaDesc &d)
  {
    if (!d.valid) return true;
    if (((unsigned long long)(size_t)d.base) & 15) return false;
    for (int i = 1; i < 5; i++)
      if (d.stride[i] % 16 || d.stride[i] >= (1ull << 40)) return false;
    for (int i = 0; i < 5; i++)
      if (d.box[i] == 0 || d.box[i] > 256) return false;
    return true;
  }

  // ---- work items ---------------------------------------------------------------------------------------------------
  struct TmaItem {
    int par, y0, z0, t; // output parity, tile origin, time slice
  };

  B2_HD void tma_item(TmaItem &it, const TmaPlan &p, int w)
  {
    const int tile = w / p.T;
    it.t = w - tile * p.T;
    const int per_par = p.nty * p.ntz;
    const int pi = tile / per_par;
    const int r = tile - pi * per_par;
    const int tzi = r / p.nty;
    it.par = p.n_parity == 2 ? pi : p.parity;
    it.y0 = (r - tzi * p.nty) * p.TY;
    it.z0 = tzi * p.TZ;
  }

  B2_HD void tma_work_range(int &w0, int &w1, const TmaPlan &p, int cta, int n_cta)
  {
    w0 = (int)((long long)p.n_items * cta / n_cta);
    w1 = (int)((long long)p.n_items * (cta + 1) / n_cta);
  }

  // One box load: which map, where (coordinates: lane, plane, y, z, t), destination offset inside the slot.
  struct TmaBox {
    int map, c[5], dst;
  };

  B2_HD int tma_wrap(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }

  // input-spinor tile rows of slice s (already wrapped into [0, T))
  B2_HD void tma_center_box(TmaBox &b, const TmaItem &it, int s)
  {
    b.map = TM_SC;
    b.c[0] = 0, b.c[1] = 0, b.c[2] = it.y0, b.c[3] = it.z0, b.c[4] = s;
    b.dst = 0;
  }

  // the four halo boxes of slice it.t
  B2_HD void tma_halo_boxes(TmaBox *b, const TmaPlan &p, const TmaItem &it)
  {
    for (int k = 0; k < 4; k++) b[k].c[0] = 0, b[k].c[1] = 0, b[k].c[2] = it.y0, b[k].c[3] = it.z0, b[k].c[4] = it.t;
    b[0].map = TM_SY, b[0].c[2] = tma_wrap(it.y0 - 1, p.Y), b[0].dst = tma_halo_ym(p);
    b[1].map = TM_SY, b[1].c[2] = tma_wrap(it.y0 + p.TY, p.Y), b[1].dst = tma_halo_yp(p);
    b[2].map = TM_SZ, b[2].c[3] = tma_wrap(it.z0 - 1, p.Z), b[2].dst = tma_halo_zm(p);
    b[3].map = TM_SZ, b[3].c[3] = tma_wrap(it.z0 + p.TZ, p.Z), b[3].dst = tma_halo_zp(p);
  }

  // boxes of link stage (item, d): returns how many (2 or 3)
  B2_HD int tma_link_boxes(TmaBox *b, const TmaPlan &p, const TmaItem &it, int d)
  {
    for (int k = 0; k < 3; k++) b[k].c[0] = 0, b[k].c[1] = d * p.GP, b[k].c[2] = it.y0, b[k].c[3] = it.z0, b[k].c[4] = it.t;
    b[0].map = TM_GF, b[0].dst = 0;
    int n = 2;
    switch (d) {
    case 0: b[1].map = TM_GB, b[1].dst = tma_link_bwd_a(p); break;
    case 1:
      b[1].map = TM_GYA, b[1].c[2] = tma_wrap(it.y0 - 1, p.Y), b[1].dst = tma_link_bwd_a(p);
      if (p.TY > 1) b[2].map = TM_GYB, b[2].dst = tma_link_bwd_b(p), n = 3;
      break;
    case 2:
      b[1].map = TM_GZA, b[1].c[3] = tma_wrap(it.z0 - 1, p.Z), b[1].dst = tma_link_bwd_a(p);
      if (p.TZ > 1) b[2].map = TM_GZB, b[2].dst = tma_link_bwd_b(p), n = 3;
      break;
    default: b[1].map = TM_GB, b[1].c[4] = tma_wrap(it.t - 1, p.T), b[1].dst = tma_link_bwd_a(p); break;
    }
    return n;
  }

  // The producer's program: the order in which the box loads of the work range [w0, w1) are requested.  Every request
  // may block until its ring slot has been released, so the order must follow the order in which the slots become free:
  //   link stage (w, d)      -- slot of stage (w, d) - NL, released when the consumers are done with that direction pair;
  //   "early" loads of w + 1 -- the slice the item needs first that its predecessor did not already hold (t + 1, or
  //                             t - 1 at the start of a chunk) and its halo rows: slots released at the end of item w - 1,
  //                             i.e. together with link stage (w - 1, 3): requested just before link stage (w, NL);
  //   "late" loads           -- chunk start only: slices t and t + 1, whose slots the previous chunk releases at its very
  //                             end: requested after the last link stage of item w.
  // `Issuer` provides center(n, item, slice), halo(n, item), link(ln, item, d): the CUDA kernel waits on the empty
  // barrier and issues the TMA loads, the host twin records / replays them.
  template <class Issuer> B2_HD void tma_producer_program(const TmaPlan &plan, int w0, int w1, Issuer &is)
  {
    const int NL = plan.n_link_slots;
    int cn = 0, hn = 0;
    long long ln = 0;
    TmaItem it;
    tma_item(it, plan, w0);
    // chunk start: slices t-1 (early), t and t+1 (late)
    is.center(cn, it, tma_wrap(it.t - 1, plan.T));
    is.halo(hn, it);
    is.center(cn + 1, it, it.t);
    is.center(cn + 2, it, tma_wrap(it.t + 1, plan.T));
    for (int w = w0; w < w1; w++) {
      const bool last = (w == w1 - 1) || (it.t == plan.T - 1);
      const int cn_next = cn + (last ? 3 : 1);
      TmaItem nx = it;
      const bool have_next = w + 1 < w1;
      if (have_next) tma_item(nx, plan, w + 1);
      for (int d = 0; d <= 4; d++) {
        if (have_next && (d == NL || (d == 4 && NL >= 4))) { // early loads of the next item
          if (last)
            is.center(cn_next, nx, tma_wrap(nx.t - 1, plan.T));
          else
            is.center(cn_next + 2, nx, tma_wrap(nx.t + 1, plan.T));
          is.halo(hn + 1, nx);
        }
        if (d < 4) is.link(ln++, it, d);
      }
      if (have_next && last) {
        is.center(cn_next + 1, nx, nx.t);
        is.center(cn_next + 2, nx, tma_wrap(nx.t + 1, plan.T));
      }
      cn = cn_next;
      hn++;
      it = nx;
    }
  }

  B2_HD int tma_box_bytes(const TmaDesc &d) { return (int)(4u * d.box[0] * d.box[1] * d.box[2] * d.box[3] * d.box[4]); }

  // ---- per-thread shared-memory offsets -----------------------------------------------------------------------------
  struct TmaThread {
    int active;
    int xh, ly, lz;
    int o_row;                      // own tile row record (centre slots): (lz*TY + ly) * srow
    int o_yp, o_ym, o_zp, o_zm;     // neighbour row records: offset inside the centre slot or the halo slot
    int h_yp, h_ym, h_zp, h_zm;     // 1: the record lives in the halo slot
    int g_f, g_bx, g_by, g_bz, g_bt; // link row records inside a stage (without the x lane)
  };

  B2_HD void tma_thread_init(TmaThread &th, const TmaPlan &p, int tid)
  {
    th.active = tid < p.n_consumers;
    const int l = th.active ? tid : 0;
    th.xh = l % p.Xh;
    const int r = l / p.Xh;
    th.ly = r % p.TY;
    th.lz = r / p.TY;
    const int row = th.lz * p.TY + th.ly;
    th.o_row = row * p.srow;
    th.h_yp = th.ly + 1 >= p.TY;
    th.o_yp = th.h_yp ? tma_halo_yp(p) + th.lz * p.srow : (row + 1) * p.srow;
    th.h_ym = th.ly == 0;
    th.o_ym = th.h_ym ? tma_halo_ym(p) + th.lz * p.srow : (row - 1) * p.srow;
    th.h_zp = th.lz + 1 >= p.TZ;
    th.o_zp = th.h_zp ? tma_halo_zp(p) + th.ly * p.srow : (row + p.TY) * p.srow;
    th.h_zm = th.lz == 0;
    th.o_zm = th.h_zm ? tma_halo_zm(p) + th.ly * p.srow : (row - p.TY) * p.srow;
    th.g_f = row * p.grow;
    th.g_bx = tma_link_bwd_a(p) + row * p.grow;
    th.g_bt = th.g_bx;
    th.g_by = th.ly == 0 ? tma_link_bwd_a(p) + th.lz * p.grow : tma_link_bwd_b(p) + (th.lz * (p.TY - 1) + th.ly - 1) * p.grow;
    th.g_bz = th.lz == 0 ? tma_link_bwd_a(p) + th.ly * p.grow : tma_link_bwd_b(p) + ((th.lz - 1) * p.TY + th.ly) * p.grow;
  }

  // ---- shared-memory loads --------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
  using sptr = unsigned; // shared::cta window address (nvcc's host pass never executes the loads below)
#else
  using sptr = const unsigned char *; // host twin: emulated shared memory
#endif

  template <typename V> B2_HD V lds(sptr p)
  {
#if defined(__CUDA_ARCH__)
    static_assert(sizeof(V) == 16 || sizeof(V) == 8, "vector width");
    if constexpr (sizeof(V) == 16) {
      uint4 r;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(p));
      return *reinterpret_cast<V *>(&r);
    } else {
      uint2 r;
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(p));
      return *reinterpret_cast<V *>(&r);
    }
#elif defined(__CUDACC__)
    (void)p;
    return V {};
#else
    V v;
    memcpy(&v, p, sizeof(V));
    return v;
#endif
  }

  // planes [p0, p0 + np) of the spinor record at `rec` (already offset to this thread's x lane); plane stride = row_bytes
  template <class P, int p0, int np> B2_HD void lds_planes(typename P::real *out, sptr rec, int row_bytes)
  {
    using V = typename P::svec;
#pragma unroll
    for (int i = 0; i < np; i++) {
      const V t = lds<V>(rec + (p0 + i) * row_bytes);
      vec_to_real(out + i * P::Ns, t);
    }
  }

  template <class P, int recon> B2_HD void lds_link(typename GaugeView<P, recon>::Raw &raw, sptr rec, int row_bytes)
  {
    using GV = GaugeView<P, recon>;
#pragma unroll
    for (int i = 0; i < GV::M; i++) raw.w[i] = lds<typename GV::V>(rec + i * row_bytes);
  }

  // One hop with every operand in shared memory; arithmetic identical to hop_from (dslash_site.h) operation for operation.
  template <class P, int recon, bool dagger, bool fwd, int d>
  B2_HD void tma_hop(typename P::real *acc, const GaugeView<P, recon> &U, sptr spinor_rec, int srow_bytes, sptr link_rec,
                     int grow_bytes, int link_idx)
  {
    using real = typename P::real;
    constexpr int sign = fwd ? (dagger ? +1 : -1) : (dagger ? -1 : +1);
    typename GaugeView<P, recon>::Raw raw;
    lds_link<P, recon>(raw, link_rec, grow_bytes);
    real u[18], h[12], r[12];
    U.unpack(u, raw, d, link_idx);
    if constexpr (d == 3) {
      constexpr int np = 12 / P::Ns;
      real t[12];
      lds_planes<P, (sign > 0) ? 0 : np, np>(t, spinor_rec, srow_bytes);
#pragma unroll
      for (int i = 0; i < 12; i++) h[i] = 2 * t[i];
    } else {
      real v[24];
      lds_planes<P, 0, 24 / P::Ns>(v, spinor_rec, srow_bytes);
      project(h, v, d, sign);
    }
    su3_mul<!fwd>(r, u, h);
    reconstruct_add(acc, r, d, sign);
  }

  // Slot bases of the current item
  struct TmaBases {
    sptr cm, c0, cp, halo; // centre slots holding slices t-1, t, t+1; halo slot of slice t
  };

  // forward + backward hop of dimension d.  `rpar` = (y + z + t + parity) & 1 of the output site: x = 2 xh + rpar.
  template <class P, int recon, bool dagger, int d>
  B2_HD void tma_hop_pair(typename P::real *acc, const GaugeView<P, recon> &U, const TmaPlan &p, const TmaThread &th,
                          const TmaBases &b, sptr stage, int rpar, int x_cb, int x_cb_tm)
  {
    const int svec = p.svec, gvec = p.gvec;
    const int srow_b = svec * p.Xh, grow_b = gvec * p.Xh; // plane strides inside a record
    const int xs = th.xh * svec, xg = th.xh * gvec;
    if constexpr (d == 0) {
      const int xhp = rpar ? (th.xh + 1 == p.Xh ? 0 : th.xh + 1) : th.xh;
      const int xhm = rpar ? th.xh : (th.xh == 0 ? p.Xh - 1 : th.xh - 1);
      tma_hop<P, recon, dagger, true, 0>(acc, U, b.c0 + th.o_row + xhp * svec, srow_b, stage + th.g_f + xg, grow_b, x_cb);
      tma_hop<P, recon, dagger, false, 0>(acc, U, b.c0 + th.o_row + xhm * svec, srow_b, stage + th.g_bx + xhm * gvec, grow_b, x_cb);
    } else if constexpr (d == 1) {
      tma_hop<P, recon, dagger, true, 1>(acc, U, (th.h_yp ? b.halo : b.c0) + th.o_yp + xs, srow_b, stage + th.g_f + xg, grow_b, x_cb);
      tma_hop<P, recon, dagger, false, 1>(acc, U, (th.h_ym ? b.halo : b.c0) + th.o_ym + xs, srow_b, stage + th.g_by + xg, grow_b, x_cb);
    } else if constexpr (d == 2) {
      tma_hop<P, recon, dagger, true, 2>(acc, U, (th.h_zp ? b.halo : b.c0) + th.o_zp + xs, srow_b, stage + th.g_f + xg, grow_b, x_cb);
      tma_hop<P, recon, dagger, false, 2>(acc, U, (th.h_zm ? b.halo : b.c0) + th.o_zm + xs, srow_b, stage + th.g_bz + xg, grow_b, x_cb);
    } else {
      tma_hop<P, recon, dagger, true, 3>(acc, U, b.cp + th.o_row + xs, srow_b, stage + th.g_f + xg, grow_b, x_cb);
      tma_hop<P, recon, dagger, false, 3>(acc, U, b.cm + th.o_row + xs, srow_b, stage + th.g_bt + xg, grow_b, x_cb_tm);
    }
  }

  // clover / twist / xpay epilogue and store: the `complete` branch of dslash_site_interior (unpartitioned lattice)
  template <class P, int recon, bool dagger, bool xpay, OpType op>
  B2_HD void tma_epilogue(typename P::real *acc, const DslashArgs<P, recon> &arg, int x_cb, int parity)
  {
    using real = typename P::real;
    if constexpr (op == OP_CLOVER_PC) clover_apply_site<P, true>(acc, arg.A, x_cb, parity);
    if constexpr (op == OP_TM_PC) twist_apply(acc, arg.a, arg.twist_b());
    if constexpr (xpay) {
      real xv[24];
      arg.x[parity].template load<Cache::STREAM>(xv, x_cb);
      if constexpr (op == OP_CLOVER) clover_apply_site<P, false>(xv, arg.A, x_cb, parity);
      if constexpr (op == OP_TM) twist_apply(xv, (real)1, arg.twist_b());
      if constexpr (op == OP_TM_PC) {
#pragma unroll
        for (int i = 0; i < 24; i++) acc[i] = xv[i] + acc[i];
      } else {
#pragma unroll
        for (int i = 0; i < 24; i++) acc[i] = xv[i] + arg.a * acc[i];
      }
    }
    arg.out[parity].save(acc, x_cb);
  }

  // checkerboard index of (xh, y, z, t) and of the same site one slice back (periodic)
  B2_HD void tma_site_index(int &x_cb, int &x_cb_tm, int &rpar, const TmaPlan &p, const TmaThread &th, const TmaItem &it)
  {
    const int y = it.y0 + th.ly, z = it.z0 + th.lz;
    const int slice = p.Xh * p.Y * p.Z;
    const int in_slice = (z * p.Y + y) * p.Xh + th.xh;
    x_cb = it.t * slice + in_slice;
    x_cb_tm = (it.t == 0 ? p.T - 1 : it.t - 1) * slice + in_slice;
    rpar = (y + z + it.t + it.par) & 1;
  }

} // namespace b200
