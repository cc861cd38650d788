/* ecp_math.h - per-thread numerical building blocks of the ECP kernels.
 *
 * Every function is `__host__ __device__` so that the *same source* the sm_100a kernels execute can be
 * compiled by g++ into a test-only harness (tests/hostcheck) and compared with the oracle on the CPU
 * before GPU time is spent.  The product never runs these on the host.
 *
 * The arithmetic follows the reference's operation order (file:line cited per function).  The host harness compiles
 * this header with -ffp-contract=off (bit-identity tests against the oracle); the CUDA translation unit allows FMA
 * contraction except where ECP_MUL_RN / ECP_ADD_RN pin a value an integer decision reads (see below).
 */
#ifndef ECP_MATH_H
#define ECP_MATH_H

#include <math.h>
#include <stdint.h>

#include "ecp_dev.h"

#if defined(__CUDACC__)
#define ECP_HD __host__ __device__ __forceinline__
#else
#define ECP_HD static inline
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* The CUDA translation unit is compiled with FMA contraction enabled (round 2; round 1 used -fmad=false).  Products and
 * sums whose value an INTEGER decision reads are formed with these, so that no contraction can move them off the
 * reference's doubles: the Bessel node index (src/bessel.c:141).  Everything the host decides (screening, windows,
 * triple list) never passes through device arithmetic.  Floating-point gates and convergence tests compare values that
 * already differ from the reference's in the last ulp (device exp / sin, summation order); a contraction there adds
 * nothing new in kind ("stop-level flips"). */
#if defined(__CUDA_ARCH__)
#define ECP_MUL_RN(a, b) __dmul_rn((a), (b))
#define ECP_ADD_RN(a, b) __dadd_rn((a), (b))
#else
#define ECP_MUL_RN(a, b) ((a) * (b))
#define ECP_ADD_RN(a, b) ((a) + (b))
#endif

/* index helpers (reference src/dimensions.h:18-29) */
ECP_HD int ecp_ld(int l) { return (l + 1) * (l + 1); }
ECP_HD int ecp_cd(int l) { return (l + 1) * (l + 2) * (l + 3) / 6; }
ECP_HD int ecp_ijk(int l) { return (l + 1) * (l + 2) / 2; }
ECP_HD int ecp_cidx(int l, int c) { return ecp_cd(l - 1) + c; }

/* ------------------------------------------------------------------------------------------------
 * Weighted modified spherical Bessel functions K_0..K_lmax(z) = e^-z M_l(z)
 * (reference src/bessel.c:101-199).  Table transposed to tabT[index*stride + l].
 * KM is the compile-time bound on lmax; arrays live in registers (all indices static after unroll).
 * Returns the branch taken (0 small-z, 1 Taylor, 2 asymptotic).
 * Taylor branch: the derivative recurrence K'_j = C_j (K_j-1 - K_j+1) - K_j + K_j+1 is evaluated as
 * fma(C_j, K_j-1 - K_j+1, K_j+1 - K_j) and the series accumulated with fma (3 + 1 instead of 4 + 2 FP64 instructions
 * per step).  The correction terms carry factors dz^i / i! <= 5e-3, so this moves K by far less than one ulp of K
 * (the result is the reference's double in all but rare rounding-boundary cases, where it differs by one ulp).
 * ---------------------------------------------------------------------------------------------- */
/* a / B for a small integer constant B, correctly rounded, in three instructions instead of the division routine
 * (Markstein: y = RN(1/B), q = RN(a y) is within one ulp of a / B, r = a - B q is exact in an fma, RN(q + r y) = RN(a / B)
 * for every a whose quotient is a normal number; B = 3, 5, 100 have no all-ones significand).  The node abscissa
 * index / 100 (src/bessel.c:143) and the series factors dz^i / i! (:170) keep the reference's doubles
 * (tests/test_host.py::test_device_math_bessel_rsh_bitwise sweeps all three branches).  Together with the recurrence
 * factors C_j = j / (2j + 1) as immediates instead of 35 loads per evaluation: type-1 1.13 -> 0.99 ms, fallback 0.94 ->
 * 0.83 ms on Au20, results unchanged bit for bit (profiles/r2/README.md). */
template <int B>
ECP_HD double ecp_div_const(double a) {
  const double y = 1.0 / B;
  const double q = a * y;
  const double r = fma(-(double)B, q, a);
  return fma(r, y, q);
}

template <int KM>
ECP_HD int ecp_bessel(const double *__restrict__ tabT, int stride, const double *__restrict__ Cj, int lmax, double z,
                      double (&K)[KM + 1]) {
  if (z < 1.0E-7) {
    if (z <= 0) {
      K[0] = 1.0;
#pragma unroll
      for (int l = 1; l <= KM; l++) K[l] = 0.0;
    } else {
      K[0] = 1 - z;
#pragma unroll
      for (int l = 1; l <= KM; l++)
        if (l <= lmax) K[l] = K[l - 1] * z / (2 * l + 1);
    }
    return 0;
  } else if (z < 16.0) {
    double d[KM + 6];
    const int maxL = lmax + 5;
    const int index = (int)floor(ECP_ADD_RN(ECP_MUL_RN(z, 100.0), 0.5));
    const double dz = z - ecp_div_const<100>((double)index);
    const double *row = tabT + (size_t)index * stride;
    double scale = 1.0;
#pragma unroll
    for (int l = 0; l <= KM + 5; l++) d[l] = (l <= maxL) ? row[l] : 0.0;
#pragma unroll
    for (int l = 0; l <= KM; l++) K[l] = d[l];
#pragma unroll
    for (int i = 1; i <= 5; i++) {
      const int top = maxL - i;
      double prev = d[0];
      d[0] = d[1] - d[0];
#pragma unroll
      for (int j = 1; j <= KM + 5 - i; j++) {
        if (j <= top) {
          const double cur = d[j];
#if defined(ECP_BESSEL_3OP)
          d[j] = fma((double)j / (2.0 * j + 1.0), prev - d[j + 1], d[j + 1] - cur); /* C_j as an immediate */
#else
          /* C_j (prev - next) + (next - cur) = C_j prev + (1 - C_j) next - cur: two fused operations instead of three */
          d[j] = fma((double)j / (2.0 * j + 1.0), prev, fma((double)(j + 1) / (2.0 * j + 1.0), d[j + 1], -cur));
#endif
          prev = cur;
        }
      }
      scale = (i == 1) ? scale * dz : ((i == 2) ? scale * dz * 0.5 : ((i == 4) ? scale * dz * 0.25 : (i == 3 ? ecp_div_const<3>(scale * dz) : ecp_div_const<5>(scale * dz))));
#pragma unroll
      for (int j = 0; j <= KM; j++)
        if (j <= lmax) K[j] = fma(scale, d[j], K[j]);
    }
    return 1;
  } else {
    double A[KM + 1];
    A[0] = 0.5 / z;
#pragma unroll
    for (int l = 0; l <= KM; l++) K[l] = A[0];
#pragma unroll
    for (int l = 1; l <= KM; l++) {
      if (l <= lmax) {
        double f = l * (l + 1);
#pragma unroll
        for (int i = 1; i < l; i++) {
          K[l] += f * A[i];
          f *= (l + i + 1) * (l - i);
        }
        A[l] = -A[0] * A[l - 1] / l;
        K[l] += f * A[l];
      }
    }
    return 2;
  }
}

/* Same function with run-time order bound and caller-provided storage (shared memory in the kernels):
 * K[0..lmax] result, d[0..lmax+5] scratch.  No unrolling waste for small lmax, no register arrays.
 * Operation order identical to ecp_bessel / the reference (the K_j updates are merely interleaved with
 * the derivative recurrence, each K_j sees the same sequence of additions). */
ECP_HD int ecp_bessel_mem(const double *__restrict__ tabT, int stride, const double *__restrict__ Cj, int lmax,
                          double z, double *K, double *d) {
  if (z < 1.0E-7) {
    if (z <= 0) {
      K[0] = 1.0;
      for (int l = 1; l <= lmax; l++) K[l] = 0.0;
    } else {
      double k = 1 - z;
      K[0] = k;
      for (int l = 1; l <= lmax; l++) {
        k = k * z / (2 * l + 1);
        K[l] = k;
      }
    }
    return 0;
  } else if (z < 16.0) {
    const int maxL = lmax + 5;
    const int index = (int)floor(ECP_ADD_RN(ECP_MUL_RN(z, 100.0), 0.5));
    const double dz = z - ecp_div_const<100>((double)index);
    const double *row = tabT + (size_t)index * stride;
    double scale = 1.0;
    for (int l = 0; l <= maxL; l++) {
      const double v = row[l];
      d[l] = v;
      if (l <= lmax) K[l] = v;
    }
    for (int i = 1; i <= 5; i++) {
      const int top = maxL - i;
      double prev = d[0], next = d[1];
      const double d0 = next - prev;
      d[0] = d0;
      scale = (i == 1) ? scale * dz : ((i == 2) ? scale * dz * 0.5 : ((i == 4) ? scale * dz * 0.25 : (i == 3 ? ecp_div_const<3>(scale * dz) : ecp_div_const<5>(scale * dz))));
      K[0] = fma(scale, d0, K[0]);
      for (int j = 1; j <= top; j++) {
        const double cur = next;
        next = d[j + 1];
        const double nd = fma(Cj[j], prev - next, next - cur);
        d[j] = nd;
        prev = cur;
        if (j <= lmax) K[j] = fma(scale, nd, K[j]);
      }
    }
    return 1;
  } else {
    double *A = d;
    A[0] = 0.5 / z;
    for (int l = 0; l <= lmax; l++) K[l] = A[0];
    for (int l = 1; l <= lmax; l++) {
      double f = l * (l + 1);
      double k = K[l];
      for (int i = 1; i < l; i++) {
        k += f * A[i];
        f *= (l + i + 1) * (l - i);
      }
      A[l] = -A[0] * A[l - 1] / l;
      k += f * A[l];
      K[l] = k;
    }
    return 2;
  }
}

/* ------------------------------------------------------------------------------------------------
 * ECP radial channel U_l(r) = sum_k r^n_k d_k exp(-a_k r^2)   (reference src/ecp.c:41-60)
 * Integer powers 0..4 are formed by multiplication (pow(r,2.0) == r*r when correctly rounded).
 * ---------------------------------------------------------------------------------------------- */
ECP_HD double ecp_rpow(double r, double n) {
  if (n == 2.0) return r * r;
  if (n == 1.0) return r;
  if (n == 0.0) return 1.0;
  return pow(r, n);
}
ECP_HD double ecp_pot_eval(const int *__restrict__ gl, const double *__restrict__ gn, const double *__restrict__ gd,
                           const double *__restrict__ ga, int g0, int g1, int l, double r) {
  double v = 0.0;
  const double r2 = r * r;
  for (int i = g0; i < g1; i++)
    if (gl[i] == l) v += ecp_rpow(r, gn[i]) * gd[i] * exp(-ga[i] * r2);
  return v;
}

/* Cartesian -> (r, theta, phi) with the reference's axis rules (src/util.c:74-106) */
ECP_HD void ecp_sphcoord(double x, double y, double z, double *r_, double *theta_, double *phi_) {
  const double eps = 1.0E-14;
  const double r = sqrt(x * x + y * y + z * z);
  double theta, phi;
  theta = (r < eps) ? 0.0 : acos(z / r);
  if (fabs(x) < eps) {
    if (fabs(y) < eps)
      phi = 0.0;
    else if (y < 0.0)
      phi = 1.5 * M_PI;
    else
      phi = 0.5 * M_PI;
  } else {
    phi = (x > 0.0) ? atan(y / x) : atan(y / x) + M_PI;
  }
  *r_ = r;
  *theta_ = theta;
  *phi_ = phi;
}

/* Real spherical harmonics S_lm, l <= lmax <= 2*ECP_MAX_LBS or maxLECP-1+ECP_MAX_LBS (both <= 10),
 * out[l*l + l + m]  (reference src/spherical_harmonics.c:15-114). */
#define ECP_RSH_LMAX 10
ECP_HD void ecp_rsh(int lmax, double theta, double phi, const double *__restrict__ fac,
                    const double *__restrict__ dfac, double *__restrict__ out) {
  double P[(ECP_RSH_LMAX + 1) * (ECP_RSH_LMAX + 1)];
  double s[ECP_RSH_LMAX + 2], c[ECP_RSH_LMAX + 2];
  const int n = (lmax + 1) * (lmax + 1);
#define ECP_RS(l, m) ((l) * (l) + (l) + (m))
  for (int i = 0; i < n; i++) {
    P[i] = 0.0;
    out[i] = 0.0;
  }
  for (int i = 0; i < lmax + 2; i++) s[i] = c[i] = 0.0;
  const double x = cos(theta);
  if (1.0 == x) {
    for (int l = 0; l <= lmax; l++) P[ECP_RS(l, 0)] = 1.0;
  } else if (-1.0 == x) {
    P[ECP_RS(0, 0)] = 1.0;
    for (int l = 1; l <= lmax; l++) P[ECP_RS(l, 0)] = -P[ECP_RS(l - 1, 0)];
  } else {
    s[1] = sqrt(1.0 - x * x);
    for (int l = 2; l <= lmax; l++) s[l] = s[l - 1] * s[1];
    for (int l = 0; l <= lmax; l++) {
      int m = l;
      if (0 == m)
        P[ECP_RS(l, 0)] = 1.0;
      else {
        P[ECP_RS(l, m)] = s[m] * dfac[2 * m - 1];
        m = l - 1;
        P[ECP_RS(l, m)] = x * (2 * m + 1) * P[ECP_RS(l - 1, m)];
        if (l > 1)
          for (m = 0; m <= l - 2; m++)
            P[ECP_RS(l, m)] = (x * (2 * l - 1) * P[ECP_RS(l - 1, m)] - (l + m - 1) * P[ECP_RS(l - 2, m)]) / (l - m);
      }
    }
  }
  for (int l = 0; l <= lmax; l++) {
    const double norm0 = sqrt((2.0 * l + 1.0) / (2.0 * M_PI));
    P[ECP_RS(l, 0)] = norm0 * P[ECP_RS(l, 0)];
    for (int m = 1; m <= l; m++) {
      const double norm = sqrt(fac[l - m] / fac[l + m]) * norm0;
      P[ECP_RS(l, m)] = norm * P[ECP_RS(l, m)];
    }
  }
  if (lmax > 0) {
    if (0.0 == phi) {
      for (int m = 0; m <= lmax; m++) {
        s[m] = 0.0;
        c[m] = 1.0;
      }
    } else {
      s[1] = sin(phi);
      c[1] = cos(phi);
      for (int m = 2; m <= lmax; m++) {
        s[m] = s[1] * c[m - 1] + c[1] * s[m - 1];
        c[m] = c[1] * c[m - 1] - s[1] * s[m - 1];
      }
    }
  }
  for (int l = 0; l <= lmax; l++) {
    out[ECP_RS(l, 0)] = P[ECP_RS(l, 0)] / sqrt(2.0);
    for (int m = 1; m <= l; m++) {
      out[ECP_RS(l, -m)] = P[ECP_RS(l, m)] * s[m];
      out[ECP_RS(l, +m)] = P[ECP_RS(l, m)] * c[m];
    }
  }
#undef ECP_RS
}

/* The same with the order bound fixed at compile time: every loop unrolls and the work arrays stay in registers (the
 * run-time version indexes them dynamically, which puts them into local memory on the device - ncu on k_t1prep showed
 * 2.5 warps per issue throttled by local-memory traffic).  Identical operations in identical order: identical values. */
template <int LM>
ECP_HD void ecp_rsh_t(double theta, double phi, const double *__restrict__ fac, const double *__restrict__ dfac,
                      double *__restrict__ out) {
  double P[(LM + 1) * (LM + 1)];
  double s[LM + 2], c[LM + 2];
#define ECP_RS(l, m) ((l) * (l) + (l) + (m))
#pragma unroll
  for (int i = 0; i < (LM + 1) * (LM + 1); i++) P[i] = 0.0;
#pragma unroll
  for (int i = 0; i < LM + 2; i++) s[i] = c[i] = 0.0;
  const double x = cos(theta);
  if (1.0 == x) {
#pragma unroll
    for (int l = 0; l <= LM; l++) P[ECP_RS(l, 0)] = 1.0;
  } else if (-1.0 == x) {
    P[ECP_RS(0, 0)] = 1.0;
#pragma unroll
    for (int l = 1; l <= LM; l++) P[ECP_RS(l, 0)] = -P[ECP_RS(l - 1, 0)];
  } else {
    if (LM >= 1) s[1] = sqrt(1.0 - x * x);
#pragma unroll
    for (int l = 2; l <= LM; l++) s[l] = s[l - 1] * s[1];
#pragma unroll
    for (int l = 0; l <= LM; l++) {
      if (0 == l)
        P[ECP_RS(l, 0)] = 1.0;
      else {
        P[ECP_RS(l, l)] = s[l] * dfac[2 * l - 1];
        P[ECP_RS(l, l - 1)] = x * (2 * (l - 1) + 1) * P[ECP_RS(l - 1, l - 1)];
        if (l > 1) {
#pragma unroll
          for (int m = 0; m <= l - 2; m++)
            P[ECP_RS(l, m)] = (x * (2 * l - 1) * P[ECP_RS(l - 1, m)] - (l + m - 1) * P[ECP_RS(l - 2, m)]) / (l - m);
        }
      }
    }
  }
#pragma unroll
  for (int l = 0; l <= LM; l++) {
    const double norm0 = sqrt((2.0 * l + 1.0) / (2.0 * M_PI));
    P[ECP_RS(l, 0)] = norm0 * P[ECP_RS(l, 0)];
#pragma unroll
    for (int m = 1; m <= l; m++) {
      const double norm = sqrt(fac[l - m] / fac[l + m]) * norm0;
      P[ECP_RS(l, m)] = norm * P[ECP_RS(l, m)];
    }
  }
  if (LM > 0) {
    if (0.0 == phi) {
#pragma unroll
      for (int m = 0; m <= LM; m++) {
        s[m] = 0.0;
        c[m] = 1.0;
      }
    } else {
      s[1] = sin(phi);
      c[1] = cos(phi);
#pragma unroll
      for (int m = 2; m <= LM; m++) {
        s[m] = s[1] * c[m - 1] + c[1] * s[m - 1];
        c[m] = c[1] * c[m - 1] - s[1] * s[m - 1];
      }
    }
  }
#pragma unroll
  for (int l = 0; l <= LM; l++) {
    out[ECP_RS(l, 0)] = P[ECP_RS(l, 0)] / sqrt(2.0);
#pragma unroll
    for (int m = 1; m <= l; m++) {
      out[ECP_RS(l, -m)] = P[ECP_RS(l, m)] * s[m];
      out[ECP_RS(l, +m)] = P[ECP_RS(l, m)] * c[m];
    }
  }
#undef ECP_RS
}

/* ------------------------------------------------------------------------------------------------
 * Small-grid level structure (PS93), level-major padded slot layout:
 *   slot 0 = centre point, slot 1 = pad, slots 2,3 = first pair, then level `v` occupies slots
 *   [levSlot[v], levSlot[v+1]) as (left,right) pairs in the reference's visiting order.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int levPairs[ECP_SMALL_LEVELS], levJ[ECP_SMALL_LEVELS], levN[ECP_SMALL_LEVELS], levSlot[ECP_SMALL_LEVELS + 1];
  /* screening shortcuts: a level (or an 8-slot chunk) has no in-window point for the window [start,end] iff
   * its largest left index is < start and its smallest right index is > end */
  short levMaxL[ECP_SMALL_LEVELS], levMinR[ECP_SMALL_LEVELS];
  short chMaxL[ECP_SMALL_SLOTS / 8], chMinR[ECP_SMALL_SLOTS / 8];
} EcpSmallMeta;

/* fill the screening shortcuts from the slot -> original-index table */
ECP_HD void ecp_small_meta_bounds(EcpSmallMeta *m, const int16_t *oidx) {
  for (int v = 0; v < ECP_SMALL_LEVELS; v++) {
    int mx = -1, mn = 32767;
    for (int s = m->levSlot[v]; s < m->levSlot[v + 1]; s += 2) {
      if (oidx[s] > mx) mx = oidx[s];
      if (oidx[s + 1] < mn) mn = oidx[s + 1];
    }
    m->levMaxL[v] = (short)mx;
    m->levMinR[v] = (short)mn;
  }
  for (int c = 0; c < ECP_SMALL_SLOTS / 8; c++) {
    int mx = -1, mn = 32767;
    for (int s = 8 * c; s < 8 * c + 8; s += 2) {
      if (s < 4) continue; /* slots 0..3: centre, pad, first pair - never window-tested */
      if (oidx[s] > mx) mx = oidx[s];
      if (oidx[s + 1] < mn) mn = oidx[s + 1];
    }
    m->chMaxL[c] = (short)mx;
    m->chMinR[c] = (short)mn;
  }
}

/* In-window pairs of a level form suffixes of its visiting order: the left indices ascend with the pair number j and
 * a left point counts iff idx >= start; the right indices descend and a right point counts iff idx <= end
 * (src/gc_integrators.c:186-199).  jL[v][start] / jR[v][end] = first pair of level v whose left / right point is inside
 * ([ECP_SMALL_LEVELS][ECP_SMALL_SLOTS] bytes each; a level has at most 64 pairs).  With them the fast-path loop needs
 * no per-point index test.  Returns 0 if the slot table is not monotone (never for the PS93 grid). */
ECP_HD int ecp_small_suffix_tables(const EcpSmallMeta *m, const int16_t *oidx, unsigned char *jL, unsigned char *jR) {
  for (int v = 0; v < ECP_SMALL_LEVELS; v++) {
    const int s0 = m->levSlot[v], np = (m->levSlot[v + 1] - s0) / 2;
    for (int j = 1; j < np; j++)
      if (oidx[s0 + 2 * j] <= oidx[s0 + 2 * j - 2] || oidx[s0 + 2 * j + 1] >= oidx[s0 + 2 * j - 1]) return 0;
    for (int w = 0; w < ECP_SMALL_SLOTS; w++) {
      int a = 0, b = 0;
      while (a < np && oidx[s0 + 2 * a] < w) a++;     /* first left index >= start = w */
      while (b < np && oidx[s0 + 2 * b + 1] > w) b++; /* first right index <= end = w  */
      jL[v * ECP_SMALL_SLOTS + w] = (unsigned char)a;
      jR[v * ECP_SMALL_SLOTS + w] = (unsigned char)b;
    }
  }
  return 1;
}

/* One PS93 level update after the level's points were added to I
 * (reference src/gc_integrators.c:201-214).  Returns 1 when converged (result in *res). */
ECP_HD int ecp_ps93_update(int j, int n, int cnt, double tol, double I, double *p, double *q, double *res) {
  /* j = 1 (two-point stage): p stays, err ~ |I - 2q|, q <- I;  j = 0 (one-point stage): p += I - q, err ~ |q - 3p/2|,
   * q stays.  Same arithmetic as the reference's branch-free form with the factors (1-j), j multiplied out.
   * The test err < tol with err = 16 |x| / (3n) is taken as 16 |x| < tol (3n): no division on the per-level path
   * (the two forms can only differ when err is within one ulp of tol). */
  double x;
  if (j) {
    x = I - 2 * (*q);
    *q = I;
  } else {
    *p += I - *q;
    x = *q - 3 * (*p) / 2;
  }
  if (0 == cnt) return 0;
  if (16 * fabs(x) < tol * (3 * n)) {
    *res = 16 * (*q) / (3 * n);
    return 1;
  }
  return 0;
}

/* One PSM92 level update (reference src/gc_integrators.c:73-83).  pTwoIprev = 2*I before the level,
 * qv = 2*p of the previous level; nNew = 2n+1. */
ECP_HD int ecp_psm92_update(int nNew, int cnt, double tol, double I, double pv, double qv, double *res) {
  const double N = nNew + 1.0;
  const double e = I - pv;
  if (0 == cnt) return 0;
  if (16 * e * e <= 3 * N * fabs(I - qv) * tol) {
    *res = 16 * I / (3 * N);
    return 1;
  }
  return 0;
}

/* Sequential PS93 quadrature of the type-2 fast path: integrand Fa*Fb*U on slot-ordered rows
 * (reference src/type2.c:319-322 + src/gc_integrators.c:156-217).  Summation order identical to the
 * reference's.  The three rows are strided (element of slot s at Fa[s*sa], Fb[s*sb], U[s*su]): on the device
 * the tables are slot-major ([slot][lambda] / [slot][l][N]) so that the quadratures of one triple, which sit in
 * neighbouring lanes and visit the slots in lock step, read neighbouring addresses.
 * Resumable: the levels [v0, v1) are processed on the state (I, p, q) in *st (v0 == 0 initialises it from the three
 * unconditional points), so that a kernel can stop after the first levels and hand the unconverged quadratures to a
 * second, densely packed launch without changing a single operation.
 * rc 0 converged (*result) / 1 all levels done without convergence / 2 reached v1 unconverged; *npts += evaluated
 * points. */
typedef struct {
  double I, p, q;
} EcpPs93State;
template <int UNR>
ECP_HD int ecp_ps93_fastT_levels_u(const double *__restrict__ Fa, int sa, const double *__restrict__ Fb, int sb,
                                 const double *__restrict__ U, int su, const double *__restrict__ w,
                                 const EcpSmallMeta *meta, const unsigned char *__restrict__ jL,
                                 const unsigned char *__restrict__ jR, int start, int end, double tol, int v0, int v1,
                                 EcpPs93State *st, double *result, int *npts) {
  double p, q, I;
  int np = 0;
  if (v0 == 0) {
    p = w[0] * (Fa[0] * Fb[0] * U[0]);
    q = w[2] * (Fa[2 * sa] * Fb[2 * sb] * U[2 * su]) + w[3] * (Fa[3 * sa] * Fb[3 * sb] * U[3 * su]);
    I = p + q;
    np = 3;
  } else {
    p = st->p;
    q = st->q;
    I = st->I;
  }
  const int we = end < 0 ? 0 : (end >= ECP_SMALL_SLOTS ? ECP_SMALL_SLOTS - 1 : end);
  const int ws = start < 0 ? 0 : (start >= ECP_SMALL_SLOTS ? ECP_SMALL_SLOTS - 1 : start);
  for (int v = v0; v < v1; v++) {
    const int s0 = meta->levSlot[v], npair = (meta->levSlot[v + 1] - s0) / 2;
    /* pairs [ja, npair) have their left point inside the window, pairs [jb, npair) their right point; a level without
     * any in-window point only moves the bookkeeping (cnt == 0, src/gc_integrators.c:203-208) */
    const int ja = (start > ECP_SMALL_SLOTS - 1) ? npair : jL[v * ECP_SMALL_SLOTS + ws];
    const int jb = (end < 0) ? npair : jR[v * ECP_SMALL_SLOTS + we];
    const int cnt = (npair - ja) + (npair - jb);
    const int j0 = ja < jb ? ja : jb, j1 = ja < jb ? jb : ja;
    int s = s0 + 2 * j0;
    int oa = s * sa, ob = s * sb, ou = s * su; /* running element offsets of slot s in the three strided rows */
    if (ja < jb) { /* only the left point of these pairs counts: T = 0 + left */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int j = j0; j < j1; j++, s += 2, oa += 2 * sa, ob += 2 * sb, ou += 2 * su) I += w[s] * (Fa[oa] * Fb[ob] * U[ou]);
    } else { /* only the right point */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int j = j0; j < j1; j++, s += 2, oa += 2 * sa, ob += 2 * sb, ou += 2 * su)
        I += w[s + 1] * (Fa[oa + sa] * Fb[ob + sb] * U[ou + su]);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll UNR
#endif
    for (int j = j1; j < npair; j++, s += 2, oa += 2 * sa, ob += 2 * sb, ou += 2 * su) {
      double T = w[s] * (Fa[oa] * Fb[ob] * U[ou]);
      T += w[s + 1] * (Fa[oa + sa] * Fb[ob + sb] * U[ou + su]);
      I += T;
    }
    np += cnt;
    if (ecp_ps93_update(meta->levJ[v], meta->levN[v], cnt, tol, I, &p, &q, result)) {
      if (npts) *npts += np;
      return 0;
    }
  }
  if (npts) *npts += np;
  st->I = I;
  st->p = p;
  st->q = q;
  return (v1 >= ECP_SMALL_LEVELS) ? 1 : 2;
}
ECP_HD int ecp_ps93_fastT_levels(const double *__restrict__ Fa, int sa, const double *__restrict__ Fb, int sb,
                                 const double *__restrict__ U, int su, const double *__restrict__ w,
                                 const EcpSmallMeta *meta, const unsigned char *__restrict__ jL,
                                 const unsigned char *__restrict__ jR, int start, int end, double tol, int v0, int v1,
                                 EcpPs93State *st, double *result, int *npts) {
  return ecp_ps93_fastT_levels_u<1>(Fa, sa, Fb, sb, U, su, w, meta, jL, jR, start, end, tol, v0, v1, st, result, npts);
}
ECP_HD int ecp_ps93_fastT(const double *__restrict__ Fa, int sa, const double *__restrict__ Fb, int sb,
                          const double *__restrict__ U, int su, const double *__restrict__ w, const EcpSmallMeta *meta,
                          const unsigned char *jL, const unsigned char *jR, int start, int end, double tol,
                          double *result, int *npts) {
  EcpPs93State st;
  if (npts) *npts = 0;
  return ecp_ps93_fastT_levels(Fa, sa, Fb, sb, U, su, w, meta, jL, jR, start, end, tol, 0, ECP_SMALL_LEVELS, &st, result,
                               npts);
}

/* Points of small-grid level v that the pair's window [gs, ge) tabulates (src/type1.c:121): the left indices of a level
 * ascend with the pair number j and the right ones descend (src/gc_integrators.c:186-199), so they are the left points
 * of the pairs [jLa, jLa + nLl) and the right points of the pairs [jRa, jRa + nLive - nLl) - read off the suffix tables
 * of the fast path.  cnt = points the PS93 rule counts (left idx >= gs, right idx <= ge, :190-197). */
struct T1Level {
  int s0, jLa, nLl, jRa, nLive, cnt;
};
ECP_HD T1Level t1_level(const EcpSmallMeta *sm, const unsigned char *__restrict__ jL, const unsigned char *__restrict__ jR,
                        int v, int gs, int ge) {
  T1Level L;
  L.s0 = sm->levSlot[v];
  const int npair = (sm->levSlot[v + 1] - L.s0) >> 1;
  const unsigned char *jl = jL + v * ECP_SMALL_SLOTS, *jr = jR + v * ECP_SMALL_SLOTS;
  const int top = ECP_SMALL_SLOTS - 1;
  const int lA = gs > top ? npair : jl[gs];               /* first left idx >= gs            */
  const int lE = ge > top ? npair : jl[ge];               /* first left idx >= ge            */
  const int rA = ge - 1 > top ? 0 : jr[ge - 1];           /* first right idx <= ge - 1       */
  const int rE = gs < 1 ? npair : (gs - 1 > top ? 0 : jr[gs - 1]); /* first right idx <= gs - 1 */
  const int rC = ge > top ? 0 : jr[ge];                   /* first right idx <= ge           */
  L.jLa = lA;
  L.nLl = lE > lA ? lE - lA : 0;
  L.jRa = rA;
  L.nLive = L.nLl + (rE > rA ? rE - rA : 0);
  L.cnt = (npair - lA) + (npair - rC);
  return L;
}


/* ------------------------------------------------------------------------------------------------
 * Large grid (PSM92, level-major slots): which points of a level can pass the exponent gate
 * Both large-grid integrands carry exp(e(r)) with e(r) = a r^2 + b r + c0, a < 0, and a point is tabulated only if
 * e >= ln(acc) (src/type1.c:163, src/type2.c:479-490): the live points are those with r between the two roots of
 * e(r) = ln(acc).  lg_live_range returns a range of ORIGINAL grid indices that contains all of them (roots widened,
 * one index of slack on each side; every point still takes the exact gate).  Within level lev the left points are the
 * original indices (2j+1) off - 1, ascending in j, and the right points their mirror images (src/gc_integrators.c:
 * 58-66), so the candidates of a level are two runs of pairs, in closed form: the persistent groups visit 8 candidates
 * per step instead of 8 consecutive slots (on the FM06-mapped grid [P-7s, P+9s] at most 73 % of the slots are live,
 * typically far fewer).
 * ---------------------------------------------------------------------------------------------- */
struct LgRange {
  int ilo, ihi; /* inclusive; ilo > ihi: no live point */
};
ECP_HD LgRange lg_live_range(const double *__restrict__ xo /* abscissae, original order */, int n, double a, double b,
                             double cmln, double i1, double i2) {
  /* a r^2 + b r + cmln >= 0, cmln = c0 - ln(acc) */
  LgRange R;
  const double bb = b * b, ac4 = 4.0 * a * cmln;
  double disc = bb - ac4;
  if (!(disc >= -1e-9 * (bb + fabs(ac4)))) {
    R.ilo = 1;
    R.ihi = 0;
    return R;
  }
  disc = disc > 0.0 ? disc : 0.0;
  const double sq = sqrt(disc) * (1.0 + 1e-9) + 1e-9 * fabs(b);
  const double r1 = (-b + sq) / (2.0 * a), r2 = (-b - sq) / (2.0 * a); /* a < 0: r1 <= r2 */
  const double m = 1e-9 * (fabs(r1) + fabs(r2) + 1.0);
  const double x1 = (r1 - m - i2) / i1, x2 = (r2 + m - i2) / i1;
  int lo = 0, hi = n; /* first index with xo >= x1 */
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (xo[mid] < x1)
      lo = mid + 1;
    else
      hi = mid;
  }
  R.ilo = lo - 1;
  lo = 0;
  hi = n; /* first index with xo > x2 */
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (xo[mid] <= x2)
      lo = mid + 1;
    else
      hi = mid;
  }
  R.ihi = lo;
  if (R.ilo < 0) R.ilo = 0;
  if (R.ihi > n - 1) R.ihi = n - 1;
  return R;
}
struct LgLevel {
  int jL, nL, jR, nLive; /* left points of the pairs [jL, jL + nL), right points of the pairs [jR, jR + nLive - nL) */
};
ECP_HD void lg_pair_run(int ilo, int ihi, int off, int npair, int *j0, int *n) {
  /* pairs j with ilo <= (2j+1) off - 1 <= ihi */
  const int a = ilo + 1 - off, b = ihi + 1 - off, o2 = 2 * off;
  int lo = a <= 0 ? 0 : (a + o2 - 1) / o2;
  int hi = b < 0 ? -1 : b / o2;
  if (hi > npair - 1) hi = npair - 1;
  *j0 = lo;
  *n = hi >= lo ? hi - lo + 1 : 0;
}
ECP_HD LgLevel lg_level(int largeSlots, int largeOrder, LgRange R, int lev) {
  LgLevel L;
  const int off = largeSlots >> (lev + 1), npair = 1 << (lev - 1);
  int nR;
  if (R.ilo > R.ihi) {
    L.jL = L.jR = L.nL = L.nLive = 0;
    return L;
  }
  lg_pair_run(R.ilo, R.ihi, off, npair, &L.jL, &L.nL);
  lg_pair_run(largeOrder - 1 - R.ihi, largeOrder - 1 - R.ilo, off, npair, &L.jR, &nR);
  L.nLive = L.nL + nR;
  return L;
}
/* slot of candidate m of the level (1 = the pad slot: nothing to do) */
ECP_HD int lg_slot(const LgLevel &L, int lev, int m) {
  return (m < L.nL) ? (1 << lev) + 2 * (L.jL + m) : ((m < L.nLive) ? (1 << lev) + 2 * (L.jR + m - L.nL) + 1 : 1);
}


/* the PSM92 convergence test alone (kernels that keep the one division of the result out of the per-level code) */
ECP_HD int ecp_psm92_test(int nNew, double tol, double I, double pv, double qv) {
  const double N = nNew + 1.0;
  const double e = I - pv;
  return 16 * e * e <= 3 * N * fabs(I - qv) * tol;
}

/* FM06 linear map parameters of the large grid for a primitive pair (reference src/gc_integrators.c:316-331):
 * r = i1*x + i2, w' = w*i1 */
ECP_HD void ecp_fm06_map(double zeta_p, double P, double *i1, double *i2) {
  const double sigma = 1.0 / sqrt(zeta_p);
  const double t = P - 7.0 * sigma;
  const double rmin = (t > 0.0) ? t : 0.0;
  const double rmax = P + 9.0 * sigma;
  *i1 = 0.5 * (rmax - rmin);
  *i2 = 0.5 * (rmax + rmin);
}

#endif
