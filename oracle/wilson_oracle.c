/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle (restatement of the reference's host operators).
 * See wilson_oracle_impl.h for the citation list.  Built by oracle/Makefile into
 * oracle/_build/liboracle.so and loaded through ctypes by oracle/__init__.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it; the product (quda_b200/) never does.
 *
 * Parity status: PINNED -- tests/test_oracle_pin.py checks every function here against the
 * reference's own host sources compiled in place (oracle/_ref/libquda_hostref.so, see
 * oracle/ref_glue.cpp); the reference tree stores no golden vectors.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_MATPC_EVEN_EVEN = 0, ORC_MATPC_ODD_ODD = 1, ORC_MATPC_EVEN_EVEN_ASYM = 2, ORC_MATPC_ODD_ODD_ASYM = 3 };

/* DeGrand-Rossi gamma matrices as used by the reference host code
   (tests/host_reference/gamma_reference.h:55-86, `local_gamma`): every row has exactly one
   non-zero entry; orc_gamma_col[mu][row] is its column, orc_gamma_val[mu][row] its (re, im).
   Projector table index 2*mu is 1 + gamma_mu, 2*mu+1 is 1 - gamma_mu (ibid. :4-53). */
static const int orc_gamma_col[4][4] = {{3, 2, 1, 0}, {3, 2, 1, 0}, {2, 3, 0, 1}, {2, 3, 0, 1}};
static const int orc_gamma_val[4][4][2] = {
  {{0, -1}, {0, -1}, {0, 1}, {0, 1}},   /* x */
  {{1, 0}, {-1, 0}, {-1, 0}, {1, 0}},   /* y */
  {{0, -1}, {0, 1}, {0, 1}, {0, -1}},   /* z */
  {{-1, 0}, {-1, 0}, {-1, 0}, {-1, 0}}, /* t */
};

/* checkerboard index -> coordinates, x fastest; X[0] even.  tests/utils/index_utils.cpp:4-29 */
static inline void orc_coords(int *x, const int *X, long cb, int parity)
{
  const long za = cb / (X[0] / 2);
  const long zb = za / X[1];
  x[1] = (int)(za - zb * X[1]);
  x[3] = (int)(zb / X[2]);
  x[2] = (int)(zb - (long)x[3] * X[2]);
  x[0] = (int)(2 * cb + ((x[1] + x[2] + x[3] + parity) & 1) - za * X[0]);
}

static inline long orc_cb_index(const int *x, const int *X)
{
  return ((((long)x[3] * X[2] + x[2]) * X[1] + x[1]) * X[0] + x[0]) >> 1;
}

void orc_coords_from_cb(int *x, const int *X, long cb, int parity) { orc_coords(x, X, cb, parity); }

#define REAL double
#define SUFFIX _f64
#include "wilson_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL float
#define SUFFIX _f32
#include "wilson_oracle_impl.h"
#undef REAL
#undef SUFFIX

/*
 * The reference's pass/fail metric (lib/color_spinor_util.in.cu:191-283): rescale both fields by
 * 1/max|ref component|, count components whose |delta| exceeds 10^-(f+1) for f = 0..15, and return the
 * last decade with zero failures ("accuracy level"; deviation = 10^-level must be <= tolerance,
 * tests/dslash_test_utils.h:1075-1101).  ref and test are double arrays of n reals.
 */
int orc_compare_spinor(const double *ref, const double *test, long n, long *fail16)
{
  double mx = 0;
  for (long i = 0; i < n; i++) {
    const double a = fabs(ref[i]);
    if (a > mx) mx = a;
  }
  const double rescale = 1.0 / mx;
  long fail[16];
  for (int f = 0; f < 16; f++) fail[f] = 0;
  for (long i = 0; i < n; i++) {
    const double d = fabs(rescale * ref[i] - rescale * test[i]);
    for (int f = 0; f < 16; f++)
      if (d > pow(10.0, -(f + 1)) || isnan(d)) fail[f]++;
  }
  int level = 0;
  for (int f = 0; f < 16; f++)
    if (fail[f] == 0) level = f + 1;
  if (fail16) memcpy(fail16, fail, sizeof(fail));
  return level;
}

void orc_srand(unsigned seed) { srand(seed); }
