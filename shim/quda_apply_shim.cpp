// Reference-side binding of libquda_b200.so: a translation unit a QUDA maintainer compiles INSTEAD of
// lib/dslash_wilson.cu, lib/dslash_wilson_clover.cu, lib/dslash_wilson_clover_preconditioned.cu, lib/dslash_twisted_mass.cu,
// lib/dslash_twisted_mass_preconditioned.cu, lib/dslash_clover_helper.cu and lib/dslash_pack2.cu.  It defines the free
// functions of include/dslash_quda.h (ApplyWilson :83, ApplyWilsonClover :137, ApplyWilsonCloverPreconditioned :234,
// ApplyTwistedMass :363, ApplyTwistedMassPreconditioned :403, ApplyClover :811, PackGhost :919) on top of the C ABI of
// include/b200_dslash.h, so that everything above them (Dirac*, the solvers, dslashQuda / MatQuda / invertQuda) keeps
// calling the same symbols.
//
// tests/test_shim_syntax.py compiles this file with `g++ -fsyntax-only` against the reference's own headers (recipe as
// oracle/Makefile), so the QUDA types, accessors and signatures used here are the real ones.
//
// Partitioned lattices: in the reference, ApplyWilson* OWN the halo exchange (lib/dslash_wilson.hpp:18-54 hands the
// kernel to dslash::DslashPolicyTune, lib/dslash_policy.hpp:1471-1650).  The shim does the same with the simplest of the
// reference's schedules (DslashBasic, lib/dslash_policy.hpp:366-450) expressed with ColorSpinorField's own comms calls:
// pack -> PackGhost below (b200_pack_ghost) -> gather / send -> interior kernel meanwhile -> wait / scatter -> fused
// exterior kernel.  QUDA's ghost buffers are used as they are (ColorSpinorField::Ghost2() + GhostOffset(dim, dir),
// lib/color_spinor_field.cpp:321-334).  The single-launch NVLink schedule (b200_dslash_apply_fused) additionally needs the
// arrival-flag words of a b200_comm block, which a port wires through comm_create_neighbor_memory the way
// quda_b200/comm.py does through torch.distributed.
#include <vector>

#include <quda_internal.h>
#include <color_spinor_field.h>
#include <gauge_field.h>
#include <clover_field.h>
#include <comm_quda.h>
#include <device.h>
#include <dslash_quda.h>
#include <quda_cuda_api.h>

#include <b200_dslash.h>

namespace quda
{

  static void *stream_of(const qudaStream_t &s) { return static_cast<void *>(target::cuda::get_stream(s)); }
  static void *default_stream() { return stream_of(device::get_default_stream()); }

  static b200_spinor view(const ColorSpinorField &f)
  {
    b200_spinor s;
    s.v = f.data();
    s.norm = nullptr; // block-float norms follow the data exactly as FloatNOrder lays them out
    s.parity_stride_bytes = f.SiteSubset() == QUDA_FULL_SITE_SUBSET ? f.Bytes() / 2 : 0;
    s.volume_cb = static_cast<int>(f.VolumeCB());
    s.n_parity = f.SiteSubset(); // 1 or 2
    return s;
  }

  static b200_gauge view(const GaugeField &U)
  {
    b200_gauge g;
    g.gauge = U.data();
    g.parity_stride_bytes = U.Bytes() / 2;
    g.stride = U.Stride(); // VolumeCB() + pad, ghost links in the pad
    g.reconstruct = U.Reconstruct() == QUDA_RECONSTRUCT_NO ? 18 : static_cast<int>(U.Reconstruct());
    g.anisotropy = U.Anisotropy();
    g.link_max = U.LinkMax();
    g.t_boundary = U.TBoundary(); // QUDA_PERIODIC_T = 1, QUDA_ANTI_PERIODIC_T = -1
    g.first_time_slice = comm_coord(3) == 0;
    g.last_time_slice = comm_coord(3) == comm_dim(3) - 1;
    return g;
  }

  static b200_clover view(const CloverField &A, bool inverse)
  {
    b200_clover c;
    const bool dyn = clover::dynamic_inverse();
    c.clover = A.data(inverse && !dyn);
    c.parity_stride_bytes = A.Bytes() / 2;
    c.compressed = clover::reconstruct();
    c.dynamic_inverse = dyn;
    c.diagonal = A.Diagonal();
    c.max_element = A.max_element(inverse && !dyn);
    return c;
  }

  static bool partitioned(const int *comm_override)
  {
    for (int d = 0; d < 4; d++)
      if (comm_dim_partitioned(d) && (!comm_override || comm_override[d])) return true;
    return false;
  }

  // receive side: QUDA's ghost zone of `in` as it is (one contiguous buffer, per-(dim, dir) byte offsets)
  static void fill_halo(b200_halo &h, const ColorSpinorField &in, const int *comm_override)
  {
    h = {};
    const char *ghost = static_cast<const char *>(in.Ghost2());
    for (int d = 0; d < 4; d++) {
      h.comm_dim[d] = comm_dim_partitioned(d) && (!comm_override || comm_override[d]);
      if (!h.comm_dim[d]) continue;
      for (int dir = 0; dir < 2; dir++) h.ghost[d][dir] = const_cast<char *>(ghost) + in.GhostOffset(d, dir);
    }
    // wait_flag stays NULL: the exchange below is stream / host ordered, the exterior kernel does not poll
  }

  static void check(int rc)
  {
    if (rc != B200_SUCCESS) errorQuda("%s", b200_last_error());
  }

  // one source on a partitioned lattice: pack, exchange, interior, exterior
  static void apply_partitioned(b200_dslash_args &args, const ColorSpinorField &in, int parity, bool dagger, const int *comm_override)
  {
    const qudaStream_t stream = device::get_default_stream();
    const int in_parity = 1 - parity;
    MemoryLocation loc[2 * QUDA_MAX_DIM];
    for (int i = 0; i < 2 * QUDA_MAX_DIM; i++) loc[i] = Device;
    in.pack(1, in_parity, dagger, stream, loc, Device); // -> quda::PackGhost below
    for (int d = 0; d < 4; d++) {
      if (!(comm_dim_partitioned(d) && (!comm_override || comm_override[d]))) continue;
      for (int dir = 0; dir < 2; dir++) {
        in.recvStart(2 * d + dir, stream);
        in.gather(2 * d + dir, stream);
      }
    }
    qudaStreamSynchronize(stream);
    for (int d = 0; d < 4; d++)
      if (comm_dim_partitioned(d) && (!comm_override || comm_override[d]))
        for (int dir = 0; dir < 2; dir++) in.sendStart(2 * d + dir, stream);
    // halo-independent part while the faces travel
    fill_halo(args.halo, in, comm_override);
    args.kernel = B200_KERNEL_INTERIOR;
    check(b200_dslash_apply(&args));
    for (int d = 0; d < 4; d++)
      if (comm_dim_partitioned(d) && (!comm_override || comm_override[d]))
        for (int dir = 0; dir < 2; dir++) {
          in.commsWait(2 * d + dir, stream);
          in.scatter(2 * d + dir, stream);
        }
    args.kernel = B200_KERNEL_EXTERIOR; // one fused launch over all partitioned faces
    check(b200_dslash_apply(&args));
    in.bufferIndex = (1 - in.bufferIndex);
  }

  static void apply(b200_op op, cvector_ref<ColorSpinorField> &out, cvector_ref<const ColorSpinorField> &in, const GaugeField &U,
                    const CloverField *A, double a, double b, bool with_x, bool asymmetric,
                    cvector_ref<const ColorSpinorField> &x, int parity, bool dagger, const int *comm_override)
  {
    b200_dslash_args args = {};
    args.abi_version = B200_ABI_VERSION;
    args.op = op;
    args.kernel = B200_KERNEL_AUTO;
    args.precision = in[0].Precision(); // QudaPrecision == bytes per real
    for (int d = 0; d < 4; d++) args.X[d] = U.X()[d];
    args.parity = parity == QUDA_INVALID_PARITY ? 0 : parity;
    args.dagger = dagger;
    args.a = a;
    args.b = b;
    args.asymmetric = asymmetric;
    args.U = view(U);
    if (A) args.A = view(*A, op == B200_OP_CLOVER_PC);
    args.stream = default_stream();
    const bool part = partitioned(comm_override);
    if (in.size() > 1 && !part && op <= B200_OP_CLOVER_PC) {
      // multi-RHS batch sharing U (and A): ONE call
      std::vector<b200_spinor> o, s, xs;
      for (auto i = 0u; i < in.size(); i++) {
        o.push_back(view(out[i]));
        s.push_back(view(in[i]));
        if (with_x) xs.push_back(view(x[i]));
      }
      check(b200_dslash_apply_multi(&args, static_cast<int>(in.size()), o.data(), s.data(), with_x ? xs.data() : nullptr));
      return;
    }
    for (auto i = 0u; i < in.size(); i++) {
      args.out = view(out[i]);
      args.in = view(in[i]);
      args.x = with_x ? view(x[i]) : b200_spinor {};
      if (part) {
        apply_partitioned(args, in[i], args.parity, dagger, comm_override);
      } else {
        args.kernel = B200_KERNEL_AUTO;
        check(b200_dslash_apply(&args));
      }
    }
  }

  void ApplyWilson(cvector_ref<ColorSpinorField> &out, cvector_ref<const ColorSpinorField> &in, const GaugeField &U, double a,
                   cvector_ref<const ColorSpinorField> &x, int parity, bool dagger, const int *comm_override, TimeProfile &)
  {
    apply(B200_OP_WILSON, out, in, U, nullptr, a, 0.0, a != 0.0, false, x, parity, dagger, comm_override);
  }

  void ApplyWilsonClover(cvector_ref<ColorSpinorField> &out, cvector_ref<const ColorSpinorField> &in, const GaugeField &U,
                         const CloverField &A, double a, cvector_ref<const ColorSpinorField> &x, int parity, bool dagger,
                         const int *comm_override, TimeProfile &)
  {
    apply(B200_OP_CLOVER, out, in, U, &A, a, 0.0, a != 0.0, false, x, parity, dagger, comm_override);
  }

  void ApplyWilsonCloverPreconditioned(cvector_ref<ColorSpinorField> &out, cvector_ref<const ColorSpinorField> &in,
                                       const GaugeField &U, const CloverField &A, double a,
                                       cvector_ref<const ColorSpinorField> &x, int parity, bool dagger,
                                       const int *comm_override, TimeProfile &)
  {
    apply(B200_OP_CLOVER_PC, out, in, U, &A, a, 0.0, a != 0.0, false, x, parity, dagger, comm_override);
  }

  // degenerate twisted mass: same stencil, a (1 + i b gamma5) rotation in the epilogue; for the preconditioned form `a` is the
  // scale of the rotation (not an xpay switch), so x travels only when xpay is set
  void ApplyTwistedMass(cvector_ref<ColorSpinorField> &out, cvector_ref<const ColorSpinorField> &in, const GaugeField &U,
                        double a, double b, cvector_ref<const ColorSpinorField> &x, int parity, bool dagger,
                        const int *comm_override, TimeProfile &)
  {
    apply(B200_OP_TWISTED_MASS, out, in, U, nullptr, a, b, true, false, x, parity, dagger, comm_override);
  }

  void ApplyTwistedMassPreconditioned(cvector_ref<ColorSpinorField> &out, cvector_ref<const ColorSpinorField> &in,
                                      const GaugeField &U, double a, double b, bool xpay,
                                      cvector_ref<const ColorSpinorField> &x, int parity, bool dagger, bool asymmetric,
                                      const int *comm_override, TimeProfile &)
  {
    apply(B200_OP_TWISTED_MASS_PC, out, in, U, nullptr, a, b, xpay, asymmetric, x, parity, dagger, comm_override);
  }

  void ApplyClover(cvector_ref<ColorSpinorField> &out, cvector_ref<const ColorSpinorField> &in, const CloverField &A,
                   bool inverse, int parity)
  {
    for (auto i = 0u; i < in.size(); i++) {
      b200_spinor o = view(out[i]), s = view(in[i]);
      b200_clover c = view(A, inverse);
      check(b200_clover_apply(&o, &s, &c, in[i].Precision(), inverse, parity, default_stream()));
    }
  }

  // ghost[2*d + f]: where face f of dimension d goes -- QUDA hands over local send buffers or, under
  // QUDA_P2P_REMOTE_WRITE, pointers into the neighbour's ghost zone (include/dslash.h:379-381): both are plain
  // device-addressable destinations for b200_pack_ghost
  void PackGhost(void *ghost[2 * QUDA_MAX_DIM], const ColorSpinorField &, cvector_ref<const ColorSpinorField> &in,
                 MemoryLocation, int nFace, bool dagger, int parity, bool spin_project, double, double, double, int,
                 const qudaStream_t &stream)
  {
    if (nFace != 1 || !spin_project) errorQuda("the B200 engine packs depth-1 spin-projected Wilson halos only");
    for (auto i = 0u; i < in.size(); i++) {
      b200_pack_args p = {};
      p.abi_version = B200_ABI_VERSION;
      p.precision = in[i].Precision();
      for (int d = 0; d < 4; d++) {
        p.X[d] = in[i].X()[d] * (d == 0 && in[i].SiteSubset() == QUDA_PARITY_SITE_SUBSET ? 2 : 1);
        p.comm_dim[d] = comm_dim_partitioned(d);
        for (int f = 0; f < 2; f++) p.dst[d][f] = p.comm_dim[d] ? ghost[2 * d + f] : nullptr;
      }
      p.parity = parity;
      p.dagger = dagger;
      p.in = view(in[i]);
      p.stream = stream_of(stream);
      check(b200_pack_ghost(&p));
    }
  }

} // namespace quda
