// Host-side C++ mirror of the reference's operator layer for the Wilson / Wilson-clover family:
//   ApplyWilson / ApplyWilsonClover / ApplyWilsonCloverPreconditioned / ApplyClover   (include/dslash_quda.h:83-811)
//   DiracWilson, DiracWilsonPC, DiracClover, DiracCloverPC                          (include/dirac_quda.h, lib/dirac_wilson.cpp,
//                                                                                    lib/dirac_clover.cpp)
//   blas:: axpy/xpay/.../norm2/cDotProduct and CG with reliable updates             (lib/blas_quda.cu, lib/reduce_quda.cu,
//                                                                                    lib/inv_cg_quda.cpp)
// Same names, argument meaning and error behaviour (errors throw b200::host::Error, the analogue of errorQuda);
// everything bottoms out in the C ABI of include/b200_dslash.h.  Fields are non-owning views unless created with
// ColorSpinorField::create (device memory from cudaMalloc).
#pragma once

#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/b200_dslash.h"

namespace b200
{
  namespace host
  {

    struct Error : std::runtime_error {
      using std::runtime_error::runtime_error;
    };

    enum QudaParity { QUDA_EVEN_PARITY = 0, QUDA_ODD_PARITY = 1, QUDA_INVALID_PARITY = -1 };
    enum QudaMatPCType {
      QUDA_MATPC_EVEN_EVEN = 0,
      QUDA_MATPC_ODD_ODD = 1,
      QUDA_MATPC_EVEN_EVEN_ASYMMETRIC = 2,
      QUDA_MATPC_ODD_ODD_ASYMMETRIC = 3
    };
    enum QudaSolutionType { QUDA_MAT_SOLUTION, QUDA_MATPC_SOLUTION, QUDA_MATPCDAG_MATPC_SOLUTION };

    // Halo context handed in by whoever bootstrapped the ranks (quda_b200/comm.py over torch.distributed):
    // peer-mapped ghost slabs of the neighbours, local flags, and an allreduce callback for the solver's scalars.
    struct CommContext {
      int comm_dim[4] = {0, 0, 0, 0};
      // [buffer][dim][face]: where our faces go (peer memory) and the flag to raise there
      void *send_dst[2][4][2] = {};
      void *send_signal[2][4][2] = {};
      // [buffer][dim][dir]: local receive buffers and the flags the neighbours raise
      void *recv[2][4][2] = {};
      void *recv_flag[2][4][2] = {};
      int *block_counter = nullptr;
      int *timeout_flag = nullptr;
      unsigned seq = 0;
      void *pack_stream = nullptr; // cudaStream_t for the pack kernels (nullptr: same stream as the Dslash)
      void (*allreduce_sum)(double *data, int n, void *user) = nullptr; // nullptr: single rank
      void *user = nullptr;
      // NVLink mailbox all-reduce (b200_comm::reduce_peer); n_ranks == 0: use the callback
      int rank = 0, n_ranks = 0;
      void *reduce_peer[B200_MAX_RANKS] = {};
      unsigned reduce_seq = 0;
      bool partitioned() const { return comm_dim[0] || comm_dim[1] || comm_dim[2] || comm_dim[3]; }
    };

    class ColorSpinorField
    {
      std::shared_ptr<void> owned;

    public:
      void *v = nullptr;
      int X[4] = {0, 0, 0, 0};
      int precision = 0; // bytes per real: 8, 4, 2
      int n_parity = 1;
      size_t parity_bytes = 0; // bytes of one parity block (data + norms)

      static ColorSpinorField create(const int *X, int precision, int n_parity);
      static ColorSpinorField wrap(void *v, const int *X, int precision, int n_parity);
      int VolumeCB() const { return X[0] * X[1] * X[2] * X[3] / 2; }
      size_t Bytes() const { return parity_bytes * n_parity; }
      size_t Length() const { return (size_t)24 * VolumeCB() * n_parity; } // reals
      ColorSpinorField parity_view(int p) const; // Even()/Odd() of a full field
      ColorSpinorField Even() const { return parity_view(0); }
      ColorSpinorField Odd() const { return parity_view(1); }
      b200_spinor desc() const;
    };

    struct GaugeField {
      b200_gauge g {};
      int X[4] = {0, 0, 0, 0};
      int precision = 0;
    };

    struct CloverField {
      b200_clover c {};     // the direct term A (and the only field for dynamic-inverse builds)
      b200_clover cinv {};  // A^{-1} for static inversion (clover pointer may be null when dynamic)
      int precision = 0;
      bool has_inverse() const { return cinv.clover != nullptr; }
    };

    // ---- the drop-in entry points (reference: include/dslash_quda.h)
    void ApplyWilson(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, double a,
                     const ColorSpinorField &x, int parity, bool dagger, const int *comm_override, CommContext *comm,
                     void *stream = nullptr);
    void ApplyWilsonClover(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, const CloverField &A,
                           double a, const ColorSpinorField &x, int parity, bool dagger, const int *comm_override,
                           CommContext *comm, void *stream = nullptr);
    void ApplyWilsonCloverPreconditioned(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U,
                                         const CloverField &A, double a, const ColorSpinorField &x, int parity,
                                         bool dagger, const int *comm_override, CommContext *comm, void *stream = nullptr);
    void ApplyClover(ColorSpinorField &out, const ColorSpinorField &in, const CloverField &A, bool inverse, int parity,
                     void *stream = nullptr);
    // degenerate twisted mass (include/dslash_quda.h:363-406,883)
    void ApplyTwistedMass(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, double a, double b,
                          const ColorSpinorField &x, int parity, bool dagger, const int *comm_override, CommContext *comm,
                          void *stream = nullptr);
    void ApplyTwistedMassPreconditioned(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, double a,
                                        double b, bool xpay, const ColorSpinorField &x, int parity, bool dagger,
                                        bool asymmetric, const int *comm_override, CommContext *comm, void *stream = nullptr);
    void ApplyTwistGamma(ColorSpinorField &out, const ColorSpinorField &in, double kappa, double mu, bool dagger, bool inverse,
                         void *stream = nullptr);

    // ---- blas on native-order fields of equal precision (fp64 / fp32); reductions accumulate in double
    namespace blas
    {
      void copy(ColorSpinorField &dst, const ColorSpinorField &src); // precision conversion allowed (8 <-> 4)
      void zero(ColorSpinorField &x);
      void ax(double a, ColorSpinorField &x);
      void axpy(double a, const ColorSpinorField &x, ColorSpinorField &y);            // y += a x
      void xpay(const ColorSpinorField &x, double a, ColorSpinorField &y);            // y = x + a y
      void axpby(double a, const ColorSpinorField &x, double b, ColorSpinorField &y); // y = a x + b y
      double norm2(const ColorSpinorField &x, CommContext *comm);
      double reDotProduct(const ColorSpinorField &x, const ColorSpinorField &y, CommContext *comm);
      double axpyNorm(double a, const ColorSpinorField &x, ColorSpinorField &y, CommContext *comm); // y += a x ; |y|^2
      double xmyNorm(const ColorSpinorField &x, ColorSpinorField &y, CommContext *comm);            // y = x - y ; |y|^2
      // p = r + beta p ; x += alpha p_old fused as in the reference's axpyZpbx (lib/inv_cg_quda.cpp:389)
      void axpyZpbx(double a, ColorSpinorField &p, ColorSpinorField &x, const ColorSpinorField &r, double b);
      long long flops();
    } // namespace blas

    struct DiracParam {
      const GaugeField *gauge = nullptr;
      const CloverField *clover = nullptr;
      double kappa = 0.0;
      double mu = 0.0; // twisted mass
      QudaMatPCType matpcType = QUDA_MATPC_EVEN_EVEN;
      bool dagger = false;
      int commDim[4] = {1, 1, 1, 1};
      CommContext *comm = nullptr;
      void *stream = nullptr;
    };

    class Dirac
    {
    protected:
      const GaugeField *gauge;
      double kappa;
      QudaMatPCType matpcType;
      mutable bool dagger;
      int commDim[4];
      CommContext *comm;
      void *stream;
      bool symmetric;
      int this_parity, other_parity;
      mutable long long dslash_applications = 0;

    public:
      explicit Dirac(const DiracParam &p);
      virtual ~Dirac() = default;
      virtual void Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const = 0;
      virtual void DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                              double k) const = 0;
      virtual void M(ColorSpinorField &out, const ColorSpinorField &in) const = 0;
      virtual void MdagM(ColorSpinorField &out, const ColorSpinorField &in) const = 0;
      virtual void prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                           QudaSolutionType) const = 0;
      virtual void reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType) const = 0;
      virtual bool pc() const { return false; }
      void Mdag(ColorSpinorField &out, const ColorSpinorField &in) const;
      void setCommDim(const int *c) { for (int d = 0; d < 4; d++) commDim[d] = c[d]; }
      void flipDagger() const { dagger = !dagger; }
      long long DslashApplications() const { return dslash_applications; }
      CommContext *Comm() const { return comm; }
      // "wilson", "wilsonpc", "clover", "cloverpc", "twistedmass", "twistedmasspc"
      static Dirac *create(const std::string &type, const DiracParam &p);
    };

    class DiracWilson : public Dirac
    {
    public:
      using Dirac::Dirac;
      void Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const override;
      void DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                      double k) const override;
      void M(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void MdagM(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                   QudaSolutionType) const override;
      void reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType) const override;
    };

    class DiracWilsonPC : public DiracWilson
    {
    public:
      using DiracWilson::DiracWilson;
      bool pc() const override { return true; }
      void M(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void MdagM(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                   QudaSolutionType) const override;
      void reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType) const override;
    };

    class DiracClover : public DiracWilson
    {
    protected:
      const CloverField *clover;

    public:
      explicit DiracClover(const DiracParam &p);
      void DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                      double k) const override; // A x + k D in
      void Clover(ColorSpinorField &out, const ColorSpinorField &in, int parity) const;
      void M(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void MdagM(ColorSpinorField &out, const ColorSpinorField &in) const override;
    };

    class DiracCloverPC : public DiracClover
    {
    public:
      explicit DiracCloverPC(const DiracParam &p);
      bool pc() const override { return true; }
      void CloverInv(ColorSpinorField &out, const ColorSpinorField &in, int parity) const;
      void Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const override; // A^-1 D
      void DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                      double k) const override; // x + k A^-1 D in
      void M(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void MdagM(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                   QudaSolutionType) const override;
      void reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType) const override;
    };

    // ---- degenerate (singlet) twisted mass, lib/dirac_twisted_mass.cpp
    class DiracTwistedMass : public DiracWilson
    {
    protected:
      double mu;

    public:
      explicit DiracTwistedMass(const DiracParam &p);
      void setMu(double m) { mu = m; }
      void Twist(ColorSpinorField &out, const ColorSpinorField &in) const; // (1 + i 2 kappa mu gamma5) in
      void Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const override;
      void DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                      double k) const override; // k D in + (1 + i 2 mu kappa gamma5) x
      void M(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void MdagM(ColorSpinorField &out, const ColorSpinorField &in) const override;
    };

    class DiracTwistedMassPC : public DiracTwistedMass
    {
    public:
      using DiracTwistedMass::DiracTwistedMass;
      bool pc() const override { return true; }
      void TwistInv(ColorSpinorField &out, const ColorSpinorField &in) const;
      void Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const override; // A^-1 D / D^dag A^-dag
      void DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                      double k) const override;
      void M(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void MdagM(ColorSpinorField &out, const ColorSpinorField &in) const override;
      void prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                   QudaSolutionType) const override;
      void reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType) const override;
    };

    // ---- CG on MdagM with optional mixed precision + reliable updates (lib/inv_cg_quda.cpp:63-420 restated)
    struct SolverParam {
      double tol = 1e-10;
      int maxiter = 10000;
      double delta = 0.1;       // reliable-update threshold (QudaInvertParam::reliable_delta)
      // results
      int iter = 0;
      double true_res = 0.0;
      double secs = 0.0;
      double gflops = 0.0;
      int reliable_updates = 0;
    };

    // Solve MdagM x = b.  `mat` is the high-precision operator, `matSloppy` the low-precision one (may be the same object).
    void invertCG(const Dirac &mat, const Dirac &matSloppy, ColorSpinorField &x, const ColorSpinorField &b, SolverParam &param);

  } // namespace host
} // namespace b200
