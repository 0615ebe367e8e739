/*
 * TEST INFRASTRUCTURE ONLY -- link glue for oracle/_ref/libquda_hostref.so.
 *
 * oracle/Makefile compiles the reference's own host-side operator sources *where they lie*
 *   /root/reference/tests/host_reference/wilson_dslash_reference.cpp   (wil_dslash, wil_mat, wil_matpc)
 *   /root/reference/tests/host_reference/clover_reference.cpp          (apply_clover, clover_dslash/matpc/mat)
 *   /root/reference/tests/utils/index_utils.cpp, host_blas.cpp, host_utils.cpp (setDims, neighbours,
 *                                                 constructQudaGaugeField / constructQudaCloverField)
 * with g++ (no CUDA device, no libquda, no Eigen needed for these files) and links them with this
 * file.  The glue supplies only what those translation units expect from libquda at link time:
 * error reporting, host malloc wrappers, a single-process "communicator", and inert stand-ins for
 * the GaugeField / ColorSpinorField objects that wil_dslash() builds solely to exchange ghost zones
 * (a no-op on one unpartitioned process: is_multi_gpu() is false without -DMULTI_GPU, and
 * comm_dim_partitioned() is false here, so dslashReference() never dereferences the ghost pointers).
 * The stand-ins zero-fill the object (sizeof taken from the reference's own headers) so that the
 * reference's inline destructors see empty containers.  No reference source text is copied.
 *
 * The exported C entry points (ref_*) forward to the reference functions unchanged.
 */
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <quda.h>
#include <gauge_field.h>
#include <color_spinor_field.h>
#include <host_utils.h>
#include <wilson_dslash_reference.h>

#define EXPORT extern "C" __attribute__((visibility("default")))

/* ---- what the reference sources expect from libquda ---------------------------------------- */

void errorQuda_(const char *func, const char *file, int line, ...)
{
  fprintf(stderr, " (ref_glue: errorQuda in %s, %s:%d)\n", func, file, line);
  abort();
}
FILE *getOutputFile() { return stderr; }
char *getOutputPrefix()
{
  static char prefix[] = "[ref] ";
  return prefix;
}
QudaVerbosity getVerbosity() { return QUDA_SILENT; }

namespace quda
{
  void *safe_malloc_(const char *, const char *, int, size_t size)
  {
    void *p = malloc(size);
    if (!p) abort();
    return p;
  }
  void host_free_(const char *, const char *, int, void *ptr) { free(ptr); }
  int comm_dim_partitioned(int) { return 0; }
  template <> void comm_allreduce_sum<double>(double &) { }
} // namespace quda

/* staggered long-link scaling (tests/utils/staggered_host_utils.cpp) is never reached for QUDA_WILSON_LINKS */
template <> void applyGaugeFieldScaling_long<double>(double **, int, QudaGaugeParam *, QudaDslashType) { abort(); }
template <> void applyGaugeFieldScaling_long<float>(float **, int, QudaGaugeParam *, QudaDslashType) { abort(); }

/* GaugeField's inline destructor destroys its quda_ptr members through their virtual destructor, so the
   stand-in constructor default-constructs exactly those members (protected, hence the derived peek type). */
namespace
{
  struct GaugePeek : quda::GaugeField {
    static void init_members(void *self)
    {
      auto *g = reinterpret_cast<GaugePeek *>(self);
      ::new (static_cast<void *>(&g->gauge)) quda::quda_ptr();
      for (int i = 0; i < 8; i++) ::new (static_cast<void *>(&g->gauge_array[i])) quda::quda_ptr();
      for (int i = 0; i < 2 * QUDA_MAX_DIM; i++) ::new (static_cast<void *>(&g->ghost[i])) quda::quda_ptr();
    }
  };
} // namespace

/* Inert stand-ins, bound to the mangled names the reference objects reference. */
extern "C" {
void ref_stub_gauge_ctor(void *self, const void *) asm("_ZN4quda10GaugeFieldC1ERKNS_15GaugeFieldParamE");
void ref_stub_gauge_ctor(void *self, const void *)
{
  memset(self, 0, sizeof(quda::GaugeField));
  GaugePeek::init_members(self);
}

void ref_stub_lattice_dtor(void *) asm("_ZN4quda12LatticeFieldD2Ev");
void ref_stub_lattice_dtor(void *) { }

void ref_stub_spinor_ctor(void *self, const void *) asm("_ZN4quda16ColorSpinorFieldC1ERKNS_16ColorSpinorParamE");
void ref_stub_spinor_ctor(void *self, const void *) { memset(self, 0, sizeof(quda::ColorSpinorField)); }

void ref_stub_spinor_dtor(void *) asm("_ZN4quda16ColorSpinorFieldD1Ev");
void ref_stub_spinor_dtor(void *) { }

void ref_stub_exchange(const void *, int, int, int, const void *, const void *, bool, bool, int, int, const void *) asm(
  "_ZNK4quda16ColorSpinorField13exchangeGhostE12QudaParity_siiPKNS_14MemoryLocationES4_bb15QudaPrecision_siRKNS_"
  "10vector_refIKS0_EE");
void ref_stub_exchange(const void *, int, int, int, const void *, const void *, bool, bool, int, int, const void *) { }

void ref_stub_ptr_dtor(void *) asm("_ZN4quda8quda_ptrD1Ev");
void ref_stub_ptr_dtor(void *) { }

void *ref_stub_ptr_data(const void *) asm("_ZNK4quda8quda_ptr4dataEv");
void *ref_stub_ptr_data(const void *) { return nullptr; }

/* never dispatched through: wil_dslash only destroys a stack object non-virtually */
void *ref_stub_gauge_vtable[64] asm("_ZTVN4quda10GaugeFieldE") = {nullptr};
/* quda_ptr vtable: {offset-to-top, typeinfo, complete dtor, deleting dtor} -> inert destructor */
void *ref_stub_ptr_vtable[4] asm("_ZTVN4quda8quda_ptrE")
  = {nullptr, nullptr, (void *)ref_stub_ptr_dtor, (void *)ref_stub_ptr_dtor};
}

/* ---- exported entry points ----------------------------------------------------------------- */

static QudaGaugeParam make_gauge_param(const int *X, double anisotropy, int antiperiodic_t)
{
  QudaGaugeParam p;
  memset(&p, 0, sizeof(p));
  for (int d = 0; d < 4; d++) p.X[d] = X[d];
  p.type = QUDA_WILSON_LINKS;
  p.anisotropy = anisotropy;
  p.t_boundary = antiperiodic_t ? QUDA_ANTI_PERIODIC_T : QUDA_PERIODIC_T;
  p.gauge_fix = QUDA_GAUGE_FIXED_NO;
  p.cpu_prec = QUDA_DOUBLE_PRECISION;
  return p;
}

static QudaPrecision prec_of(int bytes) { return bytes == 8 ? QUDA_DOUBLE_PRECISION : QUDA_SINGLE_PRECISION; }

EXPORT void ref_set_dims(const int *X)
{
  int x[4] = {X[0], X[1], X[2], X[3]};
  setDims(x);
}

EXPORT void ref_srand(unsigned seed) { srand(seed); }

/* constructQudaGaugeField(type = 1): random SU(3) + applyGaugeFieldScaling */
EXPORT void ref_random_gauge(void **gauge, int prec_bytes, const int *X, double anisotropy, int antiperiodic_t)
{
  QudaGaugeParam p = make_gauge_param(X, anisotropy, antiperiodic_t);
  constructQudaGaugeField(gauge, 1, prec_of(prec_bytes), &p);
}

EXPORT void ref_random_clover(void *clover, double norm, double diag, int prec_bytes)
{
  constructQudaCloverField(clover, norm, diag, prec_of(prec_bytes));
}

EXPORT void ref_wil_dslash(void *out, void **gauge, void *in, int parity, int dagger, int prec_bytes, const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  wil_dslash(out, gauge, in, parity, dagger, prec_of(prec_bytes), p);
}

EXPORT void ref_wil_mat(void *out, void **gauge, void *in, double kappa, int dagger, int prec_bytes, const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  wil_mat(out, gauge, in, kappa, dagger, prec_of(prec_bytes), p);
}

EXPORT void ref_wil_matpc(void *out, void **gauge, void *in, double kappa, int matpc, int dagger, int prec_bytes,
                          const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  wil_matpc(out, gauge, in, kappa, (QudaMatPCType)matpc, dagger, prec_of(prec_bytes), p);
}

// defined (non-static, but not declared in the header) in tests/host_reference/wilson_dslash_reference.cpp:171
void twist_gamma5(void *out, const void *in, int dagger, double kappa, double mu, QudaTwistFlavorType flavor, int V,
                  QudaTwistGamma5Type twist, QudaPrecision precision);

EXPORT void ref_twist_gamma5(void *out, void *in, int dagger, double kappa, double mu, int V_, int inverse, int prec_bytes)
{
  twist_gamma5(out, in, dagger, kappa, mu, QUDA_TWIST_SINGLET, V_, inverse ? QUDA_TWIST_GAMMA5_INVERSE : QUDA_TWIST_GAMMA5_DIRECT,
               prec_of(prec_bytes));
}

EXPORT void ref_tm_dslash(void *out, void **gauge, void *in, double kappa, double mu, int matpc, int parity, int dagger,
                          int prec_bytes, const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  tm_dslash(out, gauge, in, kappa, mu, QUDA_TWIST_SINGLET, (QudaMatPCType)matpc, parity, dagger, prec_of(prec_bytes), p);
}

EXPORT void ref_tm_mat(void *out, void **gauge, void *in, double kappa, double mu, int dagger, int prec_bytes, const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  tm_mat(out, gauge, in, kappa, mu, QUDA_TWIST_SINGLET, dagger, prec_of(prec_bytes), p);
}

EXPORT void ref_tm_matpc(void *out, void **gauge, void *in, double kappa, double mu, int matpc, int dagger, int prec_bytes,
                         const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  tm_matpc(out, gauge, in, kappa, mu, QUDA_TWIST_SINGLET, (QudaMatPCType)matpc, dagger, prec_of(prec_bytes), p);
}

EXPORT void ref_apply_clover(void *out, void *clover, void *in, int parity, int prec_bytes)
{
  apply_clover(out, clover, in, parity, prec_of(prec_bytes));
}

EXPORT void ref_clover_dslash(void *out, void **gauge, void *clover, void *in, int parity, int dagger, int prec_bytes,
                              const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  clover_dslash(out, gauge, clover, in, parity, dagger, prec_of(prec_bytes), p);
}

EXPORT void ref_clover_matpc(void *out, void **gauge, void *clover, void *clover_inv, void *in, double kappa, int matpc,
                             int dagger, int prec_bytes, const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  clover_matpc(out, gauge, clover, clover_inv, in, kappa, (QudaMatPCType)matpc, dagger, prec_of(prec_bytes), p);
}

EXPORT void ref_clover_mat(void *out, void **gauge, void *clover, void *in, double kappa, int dagger, int prec_bytes,
                           const int *X)
{
  QudaGaugeParam p = make_gauge_param(X, 1.0, 0);
  clover_mat(out, gauge, clover, in, kappa, dagger, prec_of(prec_bytes), p);
}

EXPORT int ref_matpc_enum(int which)
{
  switch (which) {
  case 0: return QUDA_MATPC_EVEN_EVEN;
  case 1: return QUDA_MATPC_ODD_ODD;
  case 2: return QUDA_MATPC_EVEN_EVEN_ASYMMETRIC;
  default: return QUDA_MATPC_ODD_ODD_ASYMMETRIC;
  }
}
