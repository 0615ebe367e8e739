/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the Wilson / Wilson-clover Dslash hot path.
 *
 * This file is a restatement (not a copy) of the reference's host-side operators; it is
 * included twice by wilson_oracle.c, once with REAL=double (suffix _f64) and once with
 * REAL=float (suffix _f32).  Nothing under quda_b200/ may include, link or call it.
 *
 * Conventions restated from the reference (all paths relative to /root/reference):
 *   - site / checkerboard indexing ........ tests/utils/index_utils.cpp:4-29 (fullLatticeIndex)
 *   - neighbour lookup ..................... tests/host_reference/dslash_reference.h:295-369 (gaugeLink),
 *                                            :419-518 (spinorNeighbor), tests/utils/host_utils.cpp:555-581
 *   - projectors 1 +/- gamma_mu (DeGrand-Rossi, NO factor 1/2)
 *                                            tests/host_reference/gamma_reference.h:4-121
 *   - stencil loop ......................... tests/host_reference/wilson_dslash_reference.cpp:41-82
 *   - M = 1 - kappa D, Mpc = 1 - kappa^2 D D tests/host_reference/wilson_dslash_reference.cpp:202-262
 *   - clover apply / clover operators ...... tests/host_reference/clover_reference.cpp:18-181
 *
 * Host field orders: gauge "QDP" order gauge[mu][(parity*Vh + x_cb)*18 + (row*3+col)*2 + reim],
 * spinor spinor[x_cb*24 + (spin*3+colour)*2 + reim], clover clover[((parity*Vh+x_cb)*2+chi)*36 + ...]
 * (6 real diagonal entries, then 15 complex strictly-lower-triangular entries, column-major).
 *
 * Floating-point operation order follows the reference loops so that, compiled with the same
 * flags, results are bit-identical to oracle/_ref (the reference sources compiled in place).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* y = a*x + y  and  y = x + a*y : tests/utils/host_blas.cpp:7-18,83-95 (the scalar is rounded to the
   field precision first). */
void FN(orc_axpy)(double a_, const REAL *x, REAL *y, long n)
{
  const REAL a = (REAL)a_;
#pragma omp parallel for
  for (long i = 0; i < n; i++) y[i] += a * x[i];
}

void FN(orc_xpay)(const REAL *x, double a_, REAL *y, long n)
{
  const REAL a = (REAL)a_;
#pragma omp parallel for
  for (long i = 0; i < n; i++) y[i] = x[i] + a * y[i];
}

/* One hop: acc += (U or U^dagger) * (1 +/- gamma_mu) * psi, all four spins. */
static inline void FN(hop)(REAL *acc, const REAL *link, const REAL *psi, int proj, int adjoint)
{
  REAL h[24], g[24];
  const int mu = proj >> 1;
  const REAL sgn = (proj & 1) ? (REAL)-1 : (REAL)1; /* even table index: 1 + gamma, odd: 1 - gamma */
  for (int s = 0; s < 4; s++) {
    const int t = orc_gamma_col[mu][s];
    const REAL cr = sgn * (REAL)orc_gamma_val[mu][s][0];
    const REAL ci = sgn * (REAL)orc_gamma_val[mu][s][1];
    for (int c = 0; c < 3; c++) {
      const REAL pr = psi[(t * 3 + c) * 2 + 0], pi = psi[(t * 3 + c) * 2 + 1];
      const REAL dr = psi[(s * 3 + c) * 2 + 0], di = psi[(s * 3 + c) * 2 + 1];
      const REAL or_ = cr * pr - ci * pi, oi = cr * pi + ci * pr;
      /* dense-table order: contributions are added in increasing column index */
      if (t < s) {
        h[(s * 3 + c) * 2 + 0] = or_ + dr;
        h[(s * 3 + c) * 2 + 1] = oi + di;
      } else {
        h[(s * 3 + c) * 2 + 0] = dr + or_;
        h[(s * 3 + c) * 2 + 1] = di + oi;
      }
    }
  }
  for (int s = 0; s < 4; s++) {
    for (int n = 0; n < 3; n++) {
      REAL re = 0, im = 0;
      for (int m = 0; m < 3; m++) {
        REAL ar, ai;
        if (!adjoint) {
          ar = link[(n * 3 + m) * 2 + 0];
          ai = link[(n * 3 + m) * 2 + 1];
        } else {
          ar = link[(m * 3 + n) * 2 + 0];
          ai = -link[(m * 3 + n) * 2 + 1];
        }
        const REAL br = h[(s * 3 + m) * 2 + 0], bi = h[(s * 3 + m) * 2 + 1];
        re += ar * br - ai * bi;
        im += ar * bi + ai * br;
      }
      g[(s * 3 + n) * 2 + 0] = re;
      g[(s * 3 + n) * 2 + 1] = im;
    }
  }
  for (int k = 0; k < 24; k++) acc[k] = acc[k] + g[k];
}

/*
 * out(x) = sum_mu [ P(2mu + dagger) U_mu(x) in(x+mu) + P(2mu + 1 - dagger) U_mu(x-mu)^dagger in(x-mu) ]
 * for all x of parity `parity` (the DESTINATION parity); `in` holds the opposite parity.
 * Periodic wrap in all directions; boundary conditions are baked into the links.
 */
void FN(orc_wil_dslash)(REAL *out, const REAL *const *gauge, const REAL *in, const int *X, int parity, int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
#pragma omp parallel for
  for (long i = 0; i < Vh; i++) {
    int x[4];
    orc_coords(x, X, i, parity);
    REAL acc[24];
    for (int k = 0; k < 24; k++) acc[k] = 0;
    for (int dir = 0; dir < 8; dir++) {
      const int mu = dir >> 1, fwd = !(dir & 1);
      int y[4] = {x[0], x[1], x[2], x[3]};
      y[mu] = (x[mu] + (fwd ? 1 : X[mu] - 1)) % X[mu];
      const long j = orc_cb_index(y, X);
      const REAL *psi = in + j * 24;
      const REAL *link = fwd ? gauge[mu] + ((long)parity * Vh + i) * 18 : gauge[mu] + ((long)(1 - parity) * Vh + j) * 18;
      FN(hop)(acc, link, psi, 2 * mu + (dir + dagger) % 2, !fwd);
    }
    for (int k = 0; k < 24; k++) out[i * 24 + k] = acc[k];
  }
}

/* Full operator out = in - kappa * D in on both parities (even block first). */
void FN(orc_wil_mat)(REAL *out, const REAL *const *gauge, const REAL *in, const int *X, double kappa, int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  FN(orc_wil_dslash)(out + Vh * 24, gauge, in, X, 1, dagger);
  FN(orc_wil_dslash)(out, gauge, in + Vh * 24, X, 0, dagger);
  FN(orc_xpay)(in, -kappa, out, 2 * Vh * 24);
}

/* Even-odd preconditioned operator out = in - kappa^2 D D in; matpc 0/2: even-even, 1/3: odd-odd. */
void FN(orc_wil_matpc)(REAL *out, const REAL *const *gauge, const REAL *in, const int *X, double kappa, int matpc,
                       int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  const int p = (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_EVEN_EVEN_ASYM) ? 0 : 1;
  FN(orc_wil_dslash)(tmp, gauge, in, X, 1 - p, dagger);
  FN(orc_wil_dslash)(out, gauge, tmp, X, p, dagger);
  FN(orc_xpay)(in, -kappa * kappa, out, Vh * 24);
  free(tmp);
}

/* out = A in on one parity; A is block diagonal, two Hermitian 6x6 chiral blocks per site. */
void FN(orc_apply_clover)(REAL *out, const REAL *clover, const REAL *in, const int *X, int parity)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
#pragma omp parallel for
  for (long i = 0; i < Vh; i++) {
    for (int chi = 0; chi < 2; chi++) {
      const REAL *diag = clover + (((long)parity * Vh + i) * 2 + chi) * 36;
      const REAL *tri = diag + 6;
      const REAL *v = in + i * 24 + chi * 12;
      REAL *w = out + i * 24 + chi * 12;
      for (int col = 0; col < 6; col++) {
        REAL re = 0, im = 0;
        for (int row = 0; row < 6; row++) {
          const REAL vr = v[2 * row], vi = v[2 * row + 1];
          if (row == col) {
            re += diag[row] * vr;
            im += diag[row] * vi;
          } else {
            const int lo = row < col ? row : col, hi = row < col ? col : row;
            const int k = 15 - (6 - lo) * (5 - lo) / 2 + hi - lo - 1;
            const REAL lr = tri[2 * k];
            const REAL li = (col < row) ? -tri[2 * k + 1] : tri[2 * k + 1];
            re += lr * vr - li * vi;
            im += lr * vi + li * vr;
          }
        }
        w[2 * col] = re;
        w[2 * col + 1] = im;
      }
    }
  }
}

/* out = A^{-1}(or A) D in : tests/host_reference/clover_reference.cpp:79-88 */
void FN(orc_clover_dslash)(REAL *out, const REAL *const *gauge, const REAL *clover, const REAL *in, const int *X,
                           int parity, int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  FN(orc_wil_dslash)(tmp, gauge, in, X, parity, dagger);
  FN(orc_apply_clover)(out, clover, tmp, X, parity);
  free(tmp);
}

/* tests/host_reference/clover_reference.cpp:91-146 */
void FN(orc_clover_matpc)(REAL *out, const REAL *const *gauge, const REAL *clover, const REAL *clover_inv,
                          const REAL *in, const int *X, double kappa, int matpc, int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  const int p = (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_EVEN_EVEN_ASYM) ? 0 : 1;
  const int q = 1 - p;
  if (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_ODD_ODD) {
    if (!dagger) {
      FN(orc_wil_dslash)(tmp, gauge, in, X, q, dagger);
      FN(orc_apply_clover)(out, clover_inv, tmp, X, q);
      FN(orc_wil_dslash)(tmp, gauge, out, X, p, dagger);
      FN(orc_apply_clover)(out, clover_inv, tmp, X, p);
    } else {
      FN(orc_apply_clover)(tmp, clover_inv, in, X, p);
      FN(orc_wil_dslash)(out, gauge, tmp, X, q, dagger);
      FN(orc_apply_clover)(tmp, clover_inv, out, X, q);
      FN(orc_wil_dslash)(out, gauge, tmp, X, p, dagger);
    }
    FN(orc_xpay)(in, -kappa * kappa, out, Vh * 24);
  } else {
    FN(orc_wil_dslash)(out, gauge, in, X, q, dagger);
    FN(orc_apply_clover)(tmp, clover_inv, out, X, q);
    FN(orc_wil_dslash)(out, gauge, tmp, X, p, dagger);
    FN(orc_apply_clover)(tmp, clover, in, X, p);
    FN(orc_xpay)(tmp, -kappa * kappa, out, Vh * 24);
  }
  free(tmp);
}

/* out = A in - kappa D in on the full lattice: tests/host_reference/clover_reference.cpp:148-181 */
void FN(orc_clover_mat)(REAL *out, const REAL *const *gauge, const REAL *clover, const REAL *in, const int *X,
                        double kappa, int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * 2 * Vh * 24);
  FN(orc_wil_dslash)(out + Vh * 24, gauge, in, X, 1, dagger);
  FN(orc_apply_clover)(tmp + Vh * 24, clover, in + Vh * 24, X, 1);
  FN(orc_wil_dslash)(out, gauge, in + Vh * 24, X, 0, dagger);
  FN(orc_apply_clover)(tmp, clover, in, X, 0);
  FN(orc_xpay)(tmp, -kappa, out, 2 * Vh * 24);
  free(tmp);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Twisted mass (degenerate / singlet flavour), tests/host_reference/wilson_dslash_reference.cpp:139-315.
 * twist_gamma5: out = b (1 + i a gamma5) in with, in the DeGrand-Rossi basis, gamma5 = diag(+1,+1,-1,-1);
 *   direct : a = 2 kappa mu flavor,  b = 1                  inverse: a = -2 kappa mu flavor, b = 1 / (1 + a^2)
 *   dagger flips the sign of a.  The reference evaluates a and b in double from real_t kappa / mu and rounds them to
 *   real_t; the loop itself runs in real_t -- restated literally so that results are bit-identical. */
void FN(orc_twist_gamma5)(REAL *out, const REAL *in, int dagger, double kappa_, double mu_, int flavor, long V, int inverse)
{
  const REAL kappa = (REAL)kappa_, mu = (REAL)mu_;
  REAL a, b;
  if (!inverse) {
    a = (REAL)(2.0 * kappa * mu * flavor);
    b = (REAL)1.0;
  } else {
    a = (REAL)(-2.0 * kappa * mu * flavor);
    b = (REAL)(1.0 / (1.0 + a * a));
  }
  if (dagger) a = (REAL)(a * -1.0);
#pragma omp parallel for
  for (long i = 0; i < V; i++) {
    REAL tmp[24];
    for (int s = 0; s < 4; s++)
      for (int c = 0; c < 3; c++) {
        const REAL a5 = (REAL)(((s / 2) ? -1.0 : +1.0) * a);
        tmp[s * 6 + c * 2 + 0] = b * (in[i * 24 + s * 6 + c * 2 + 0] - a5 * in[i * 24 + s * 6 + c * 2 + 1]);
        tmp[s * 6 + c * 2 + 1] = b * (in[i * 24 + s * 6 + c * 2 + 1] + a5 * in[i * 24 + s * 6 + c * 2 + 0]);
      }
    for (int j = 0; j < 24; j++) out[i * 24 + j] = tmp[j];
  }
}

/* A^{-1} D (not dagger, or dagger with asymmetric preconditioning) or D^dagger A^{-dagger} (dagger, symmetric): ibid. :182-200.
 * The reference twists `in` in place and undoes it afterwards; a private copy gives the same `out`. */
void FN(orc_tm_dslash)(REAL *out, const REAL *const *gauge, const REAL *in, const int *X, double kappa, double mu,
                       int flavor, int matpc, int parity, int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  const int symmetric = (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_ODD_ODD);
  if (dagger && symmetric) {
    REAL *w = (REAL *)malloc(sizeof(REAL) * Vh * 24);
    FN(orc_twist_gamma5)(w, in, dagger, kappa, mu, flavor, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, w, X, parity, dagger);
    free(w);
  } else {
    FN(orc_wil_dslash)(out, gauge, in, X, parity, dagger);
    FN(orc_twist_gamma5)(out, out, dagger, kappa, mu, flavor, Vh, 1);
  }
}

/* out = (1 + i 2 kappa mu gamma5) in - kappa D in on the full lattice: ibid. :217-237 */
void FN(orc_tm_mat)(REAL *out, const REAL *const *gauge, const REAL *in, const int *X, double kappa, double mu, int flavor,
                    int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * 2 * Vh * 24);
  FN(orc_wil_dslash)(out + Vh * 24, gauge, in, X, 1, dagger);
  FN(orc_wil_dslash)(out, gauge, in + Vh * 24, X, 0, dagger);
  FN(orc_twist_gamma5)(tmp, in, dagger, kappa, mu, flavor, 2 * Vh, 0);
  FN(orc_xpay)(tmp, -kappa, out, 2 * Vh * 24);
  free(tmp);
}

/* Even-odd preconditioned twisted-mass operator, all four matpc types x dagger: ibid. :262-315.  In the symmetric dagger
 * case the reference twists its input in place and untwists it again before the final xpay, so that xpay sees the
 * round-tripped (not the pristine) input; `w` reproduces exactly that. */
void FN(orc_tm_matpc)(REAL *out, const REAL *const *gauge, const REAL *in, const int *X, double kappa, double mu, int flavor,
                      int matpc, int dagger)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  REAL *w = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  memcpy(w, in, sizeof(REAL) * Vh * 24);
  const int p = (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_EVEN_EVEN_ASYM) ? 0 : 1;
  const int asym = (matpc == ORC_MATPC_EVEN_EVEN_ASYM || matpc == ORC_MATPC_ODD_ODD_ASYM);
  if (asym) {
    FN(orc_wil_dslash)(tmp, gauge, w, X, 1 - p, dagger);
    FN(orc_twist_gamma5)(tmp, tmp, dagger, kappa, mu, flavor, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, X, p, dagger);
    FN(orc_twist_gamma5)(tmp, w, dagger, kappa, mu, flavor, Vh, 0);
    FN(orc_xpay)(tmp, -kappa * kappa, out, Vh * 24);
  } else if (!dagger) {
    FN(orc_wil_dslash)(tmp, gauge, w, X, 1 - p, dagger);
    FN(orc_twist_gamma5)(tmp, tmp, dagger, kappa, mu, flavor, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, X, p, dagger);
    FN(orc_twist_gamma5)(out, out, dagger, kappa, mu, flavor, Vh, 1);
    FN(orc_xpay)(w, -kappa * kappa, out, Vh * 24);
  } else {
    FN(orc_twist_gamma5)(w, w, dagger, kappa, mu, flavor, Vh, 1);
    FN(orc_wil_dslash)(tmp, gauge, w, X, 1 - p, dagger);
    FN(orc_twist_gamma5)(tmp, tmp, dagger, kappa, mu, flavor, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, X, p, dagger);
    FN(orc_twist_gamma5)(w, w, dagger, kappa, mu, flavor, Vh, 0);
    FN(orc_xpay)(w, -kappa * kappa, out, Vh * 24);
  }
  free(tmp);
  free(w);
}

/*
 * Inverse of every chiral block (same packed order in and out).  The reference inverts on the
 * device (lib/clover_invert.cu, include/kernels/clover_invert.cuh: Cholesky of the Hermitian block);
 * here: Cholesky A = L L^dagger in double, then A^{-1} = L^{-dagger} L^{-1}.
 */
void FN(orc_clover_invert)(REAL *inv, const REAL *clover, long nsites)
{
#pragma omp parallel for
  for (long b = 0; b < 2 * nsites; b++) {
    const REAL *diag = clover + b * 36, *tri = diag + 6;
    double Ar[6][6], Ai[6][6];
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 6; c++) {
        if (r == c) {
          Ar[r][c] = diag[r];
          Ai[r][c] = 0;
        } else {
          const int lo = r < c ? r : c, hi = r < c ? c : r;
          const int k = 15 - (6 - lo) * (5 - lo) / 2 + hi - lo - 1;
          Ar[r][c] = tri[2 * k];
          Ai[r][c] = (r > c) ? tri[2 * k + 1] : -tri[2 * k + 1];
        }
      }
    double Lr[6][6] = {{0}}, Li[6][6] = {{0}};
    for (int j = 0; j < 6; j++) {
      double s = Ar[j][j];
      for (int k = 0; k < j; k++) s -= Lr[j][k] * Lr[j][k] + Li[j][k] * Li[j][k];
      Lr[j][j] = sqrt(s);
      for (int i = j + 1; i < 6; i++) {
        double sr = Ar[i][j], si = Ai[i][j];
        for (int k = 0; k < j; k++) { /* L[i][k] * conj(L[j][k]) */
          sr -= Lr[i][k] * Lr[j][k] + Li[i][k] * Li[j][k];
          si -= Li[i][k] * Lr[j][k] - Lr[i][k] * Li[j][k];
        }
        Lr[i][j] = sr / Lr[j][j];
        Li[i][j] = si / Lr[j][j];
      }
    }
    /* M = L^{-1} (lower triangular) by forward substitution, column by column */
    double Mr[6][6] = {{0}}, Mi[6][6] = {{0}};
    for (int c = 0; c < 6; c++) {
      for (int r = c; r < 6; r++) {
        double sr = (r == c) ? 1.0 : 0.0, si = 0.0;
        for (int k = c; k < r; k++) {
          sr -= Lr[r][k] * Mr[k][c] - Li[r][k] * Mi[k][c];
          si -= Lr[r][k] * Mi[k][c] + Li[r][k] * Mr[k][c];
        }
        Mr[r][c] = sr / Lr[r][r];
        Mi[r][c] = si / Lr[r][r];
      }
    }
    /* Ainv = M^dagger M ; Ainv[r][c] = sum_k conj(M[k][r]) M[k][c] */
    REAL *od = inv + b * 36, *ot = od + 6;
    for (int c = 0; c < 6; c++)
      for (int r = c; r < 6; r++) {
        double sr = 0, si = 0;
        for (int k = r; k < 6; k++) {
          sr += Mr[k][r] * Mr[k][c] + Mi[k][r] * Mi[k][c];
          si += Mr[k][r] * Mi[k][c] - Mi[k][r] * Mr[k][c];
        }
        if (r == c)
          od[r] = (REAL)sr;
        else {
          const int k = 15 - (6 - c) * (5 - c) / 2 + r - c - 1;
          ot[2 * k] = (REAL)sr;
          ot[2 * k + 1] = (REAL)si;
        }
      }
  }
}

/* Random SU(3) links, tests/utils/host_utils.cpp:1022-1098 (libc rand(); caller seeds with srand) followed by
   applyGaugeFieldScaling :940-975 (spatial links / anisotropy; last-time-slice t-links * -1 if antiperiodic). */
static void FN(unit)(REAL *v)
{
  double s = 0;
  for (int k = 0; k < 3; k++) {
    const REAL n2 = v[2 * k] * v[2 * k] + v[2 * k + 1] * v[2 * k + 1]; /* std::norm in field precision */
    s += n2;
  }
  s = sqrt(s);
  for (int k = 0; k < 3; k++) {
    /* complex<Float> /= double  ->  component-wise division by (Float)s in libstdc++ */
    v[2 * k] /= (REAL)s;
    v[2 * k + 1] /= (REAL)s;
  }
}

static void FN(ortho)(const REAL *a, REAL *b)
{
  double dr = 0, di = 0;
  for (int k = 0; k < 3; k++) { /* conj(a) * b, accumulated as complex<double> += complex<Float> */
    const REAL pr = a[2 * k] * b[2 * k] + a[2 * k + 1] * b[2 * k + 1];
    const REAL pi = a[2 * k] * b[2 * k + 1] - a[2 * k + 1] * b[2 * k];
    dr += pr;
    di += pi;
  }
  const REAL fr = (REAL)dr, fi = (REAL)di;
  for (int k = 0; k < 3; k++) {
    const REAL tr = fr * a[2 * k] - fi * a[2 * k + 1];
    const REAL ti = fr * a[2 * k + 1] + fi * a[2 * k];
    b[2 * k] -= tr;
    b[2 * k + 1] -= ti;
  }
}

/* w += sign * conj(u * v) */
static void FN(cprod)(REAL *w, const REAL *u, const REAL *v, int sign)
{
  w[0] += sign * (u[0] * v[0] - u[1] * v[1]);
  w[1] -= sign * (u[0] * v[1] + u[1] * v[0]);
}

static void FN(fill_link)(REAL *m)
{
  FN(unit)(m + 6);
  FN(ortho)(m + 6, m + 12);
  FN(unit)(m + 12);
  REAL *w = m, *u = m + 6, *v = m + 12;
  for (int n = 0; n < 6; n++) w[n] = 0;
  FN(cprod)(w + 0, u + 2, v + 4, +1);
  FN(cprod)(w + 0, u + 4, v + 2, -1);
  FN(cprod)(w + 2, u + 4, v + 0, +1);
  FN(cprod)(w + 2, u + 0, v + 4, -1);
  FN(cprod)(w + 4, u + 0, v + 2, +1);
  FN(cprod)(w + 4, u + 2, v + 0, -1);
}

void FN(orc_random_gauge)(REAL *const *gauge, const int *X, double anisotropy, int antiperiodic_t)
{
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  for (int mu = 0; mu < 4; mu++) {
    REAL *ev = gauge[mu], *od = gauge[mu] + Vh * 18;
    for (long i = 0; i < Vh; i++) {
      for (int m = 1; m < 3; m++)
        for (int n = 0; n < 3; n++) {
          ev[i * 18 + (m * 3 + n) * 2 + 0] = rand() / (REAL)RAND_MAX;
          ev[i * 18 + (m * 3 + n) * 2 + 1] = rand() / (REAL)RAND_MAX;
          od[i * 18 + (m * 3 + n) * 2 + 0] = rand() / (REAL)RAND_MAX;
          od[i * 18 + (m * 3 + n) * 2 + 1] = rand() / (REAL)RAND_MAX;
        }
      FN(fill_link)(ev + i * 18);
      FN(fill_link)(od + i * 18);
    }
  }
  for (int mu = 0; mu < 3; mu++)
    for (long k = 0; k < 2 * Vh * 18; k++) gauge[mu][k] /= anisotropy;
  if (antiperiodic_t) {
    for (long j = (long)(X[0] / 2) * X[1] * X[2] * (X[3] - 1); j < Vh; j++)
      for (int k = 0; k < 18; k++) {
        gauge[3][j * 18 + k] *= -1.0;
        gauge[3][(Vh + j) * 18 + k] *= -1.0;
      }
  }
}

/* tests/utils/host_utils.cpp:1162-1188 */
void FN(orc_random_clover)(REAL *res, long nsites, double norm, double diag)
{
  const REAL c = 2.0 * norm / RAND_MAX;
  static const int dst[9] = {3, 4, 5, 30, 31, 32, 33, 34, 35};
  static const int src[9] = {0, 1, 2, 6, 7, 8, 9, 16, 17};
  for (long i = 0; i < nsites; i++) {
    for (int j = 0; j < 72; j++) res[i * 72 + j] = c * rand() - norm;
    for (int ch = 0; ch < 2; ch++)
      for (int k = 0; k < 9; k++) res[i * 72 + dst[k] + 36 * ch] = -res[i * 72 + src[k] + 36 * ch];
    for (int j = 0; j < 6; j++) {
      res[i * 72 + j] += diag;
      res[i * 72 + j + 36] += diag;
    }
  }
}

/* Uniform [0,1) spinor from the rand48 clone: lib/comm_common.cpp:26-41, lib/color_spinor_util.in.cu:17-24 */
void FN(orc_random_spinor)(REAL *v, long nreal, unsigned long *state)
{
  const double twoneg48 = 0.35527136788005009e-14;
  unsigned long s = *state;
  for (long i = 0; i < nreal; i++) {
    s = (25214903917ul * s + 11ul) & 281474976710655ul;
    v[i] = (REAL)(twoneg48 * s);
  }
  *state = s;
}

#undef FN
#undef CAT
#undef CAT_
