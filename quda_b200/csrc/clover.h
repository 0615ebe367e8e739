// Clover term: native accessor (uncompressed 72-real and compressed 56-real site formats), chiral-basis
// apply, and per-site Cholesky solve for "dynamic" inversion.
//
// Reference behaviour reproduced (/root/reference/include/...):
//   clover_field_order.h:79-172    compressed block: 28 of 36 reals kept; diagonals stored as deviations from a
//                                  global `diagonal`, entries 4,5 and 30..35 implied by symmetry
//   clover_field_order.h:587-720   FloatN order: vector (parity*offset + x + volumeCB*(chirality*M_offset + i)),
//                                  fixed point scale nrm = max_element / (2*32767)
//   clover_field_order.h:206-280   internal block order: 6 real diagonals, 15 complex strictly-lower entries, column major
//   clover_field_order.h:854-866   native values are HALF the host ("packed") values -- the factor 1/2 of the basis change
//   color_spinor.h:602-633         toRel / toNonRel (no 1/sqrt2; product of the two is 2)
//   kernels/dslash_wilson_clover_preconditioned.cuh:74-101  dynamic inverse: 0.25 * Cholesky solve
//   linalg.cuh:40-140, clover_field.h:52-75  Cholesky in double for fp64/fp32 (CLOVER_PROMOTE_CHOLESKY), float for fixed point
#pragma once

#include "core.h"

namespace b200
{

  template <class P> struct CloverView {
    using store = typename P::store;
    const store *c[2]; // base per parity
    int volume_cb;
    int compressed; // 1: 28 reals per chiral block, 0: 36
    int dynamic;    // field holds A (not A^{-1}); inverse apply = Cholesky solve
    typename P::real diagonal;
    float nrm; // fixed-point scale
  };

  template <class P> struct CholT { using type = double; };
  template <> struct CholT<PrecH16> { using type = float; };

  // load one chiral block into the internal 36-real order
  template <class P, int CB> B2_HD void clover_load_block(typename P::real *a, const CloverView<P> &A, int x_cb, int parity, int chi)
  {
    using real = typename P::real;
    using V = typename P::svec;
    constexpr int N = P::Ns;
    constexpr int M = (CB + N - 1) / N;
    constexpr int Moff = CB / N;
    real tmp[M * N];
    const V *base = reinterpret_cast<const V *>(A.c[parity]);
#pragma unroll
    for (int i = 0; i < M; i++) {
      const V w = ld<Cache::STREAM>(base + (size_t)(chi * Moff + i) * A.volume_cb + x_cb);
      vec_to_real(tmp + i * N, w);
      if constexpr (P::fixed) {
#pragma unroll
        for (int j = 0; j < N; j++) tmp[i * N + j] *= A.nrm;
      }
    }
    const int sh = (chi * CB) % N;
    if constexpr (CB == 36) {
#pragma unroll
      for (int i = 0; i < 36; i++) a[i] = tmp[i + sh];
    } else {
      // stored[0..3] = internal[0..3], stored[4..27] = internal[6..29]
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = tmp[i + sh];
#pragma unroll
      for (int i = 6; i < 30; i++) a[i] = tmp[i - 2 + sh];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const real dev = a[i];
        a[i + 3] = A.diagonal - dev;
        a[i] = A.diagonal + dev;
      }
      a[30] = -a[6];
      a[31] = -a[7];
      a[32] = -a[8];
      a[33] = -a[9];
      a[34] = -a[16];
      a[35] = -a[17];
    }
  }

  // index of the (re) slot of strictly-lower element (i > j) in the 36-real block
  B2_HD constexpr int tri_idx(int i, int j) { return 6 + 2 * (15 - (6 - j) * (5 - j) / 2 + i - j - 1); }

  // y = H x for the Hermitian 6x6 block
  template <typename real> B2_HD void hmat_mul(real *y, const real *a, const real *x)
  {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      real re = a[i] * x[2 * i], im = a[i] * x[2 * i + 1];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (j == i) continue;
        const int k = (j < i) ? tri_idx(i, j) : tri_idx(j, i);
        const real ar = a[k];
        const real ai = (j < i) ? a[k + 1] : -a[k + 1];
        re += ar * x[2 * j];
        re -= ai * x[2 * j + 1];
        im += ar * x[2 * j + 1];
        im += ai * x[2 * j];
      }
      y[2 * i] = re;
      y[2 * i + 1] = im;
    }
  }

  // x <- H^{-1} x via Cholesky H = L L^dagger (reciprocal square roots kept on the diagonal)
  template <typename T, typename real> B2_HD void hmat_solve(real *x, const real *a)
  {
    T Lr[6][6], Li[6][6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (j > i) continue;
        T sr = 0, si = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
          if (k >= j) continue;
          sr += Lr[i][k] * Lr[j][k];
          sr += Li[i][k] * Li[j][k];
          si += Li[i][k] * Lr[j][k];
          si -= Lr[i][k] * Li[j][k];
        }
        if (i == j) {
          const T d = (T)a[i] - sr;
#if defined(__CUDA_ARCH__)
          Lr[i][i] = (sizeof(T) == 8) ? (T)::rsqrt((double)d) : (T)::rsqrtf((float)d);
#else
          Lr[i][i] = (T)1 / std::sqrt(d);
#endif
          Li[i][i] = 0;
        } else {
          const int k = tri_idx(i, j);
          Lr[i][j] = ((T)a[k] - sr) * Lr[j][j];
          Li[i][j] = ((T)a[k + 1] - si) * Lr[j][j];
        }
      }
    }
    T yr[6], yi[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { // forward: L y = x
      T r = x[2 * i], m = x[2 * i + 1];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (j >= i) continue;
        r -= Lr[i][j] * yr[j];
        r += Li[i][j] * yi[j];
        m -= Lr[i][j] * yi[j];
        m -= Li[i][j] * yr[j];
      }
      yr[i] = r * Lr[i][i];
      yi[i] = m * Lr[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) { // backward: L^dagger z = y ; (L^dagger)(i,j) = conj(L(j,i)), j > i
      T r = yr[i], m = yi[i];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (j <= i) continue;
        r -= Lr[j][i] * yr[j];
        r -= Li[j][i] * yi[j];
        m -= Lr[j][i] * yi[j];
        m += Li[j][i] * yr[j];
      }
      yr[i] = r * Lr[i][i];
      yi[i] = m * Lr[i][i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      x[2 * i] = (real)yr[i];
      x[2 * i + 1] = (real)yi[i];
    }
  }

  // v <- A v (inverse == false) or v <- A^{-1} v (inverse == true) on a UKQCD-basis spinor of 24 reals.
  template <class P, bool inverse> B2_HD void clover_apply_site(typename P::real *v, const CloverView<P> &A, int x_cb, int parity)
  {
    using real = typename P::real;
    real r[24]; // chiral ("relativistic") basis
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
      for (int z = 0; z < 2; z++) {
        const real v0 = v[(0 * 3 + c) * 2 + z], v1 = v[(1 * 3 + c) * 2 + z];
        const real v2 = v[(2 * 3 + c) * 2 + z], v3 = v[(3 * 3 + c) * 2 + z];
        r[(0 * 3 + c) * 2 + z] = -v1 - v3;
        r[(1 * 3 + c) * 2 + z] = v2 + v0;
        r[(2 * 3 + c) * 2 + z] = v3 - v1;
        r[(3 * 3 + c) * 2 + z] = v0 - v2;
      }
    }
#pragma unroll
    for (int chi = 0; chi < 2; chi++) {
      real a[36];
      if (A.compressed)
        clover_load_block<P, 28>(a, A, x_cb, parity, chi);
      else
        clover_load_block<P, 36>(a, A, x_cb, parity, chi);
      if (inverse && A.dynamic) {
        hmat_solve<typename CholT<P>::type>(r + 12 * chi, a);
#pragma unroll
        for (int i = 0; i < 12; i++) r[12 * chi + i] *= (real)0.25;
      } else {
        real y[12];
        hmat_mul(y, a, r + 12 * chi);
#pragma unroll
        for (int i = 0; i < 12; i++) r[12 * chi + i] = y[i];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
      for (int z = 0; z < 2; z++) {
        const real r0 = r[(0 * 3 + c) * 2 + z], r1 = r[(1 * 3 + c) * 2 + z];
        const real r2 = r[(2 * 3 + c) * 2 + z], r3 = r[(3 * 3 + c) * 2 + z];
        v[(0 * 3 + c) * 2 + z] = r1 + r3;
        v[(1 * 3 + c) * 2 + z] = -r2 - r0;
        v[(2 * 3 + c) * 2 + z] = -r3 + r1;
        v[(3 * 3 + c) * 2 + z] = -r0 + r2;
      }
    }
  }

} // namespace b200
