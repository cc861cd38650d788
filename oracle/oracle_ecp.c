/* oracle_ecp.c - TEST INFRASTRUCTURE ONLY (see oracle_ecp.h).
 *
 * Plain-C restatement of the reference algorithm for the hot path, written from the reference's
 * behaviour, flat arrays instead of its n-D container, derivative order n = 0 only.  Every function
 * names the reference file:line it follows (paths relative to /root/reference).  The arithmetic
 * (operation order, libm calls) is kept identical so that results are bit-for-bit those of the
 * compiled reference (oracle/_ref/libecp_ref.so) - that equality is what tests/test_oracle.py checks.
 *
 * Build: gcc -O2 -fPIC -ffp-contract=off (oracle/Makefile).  Never -march=native / -ffast-math.
 */
#include "oracle_ecp.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- index helpers (src/dimensions.h:18-29) ---- */
static int LD(int l) { return (l + 1) * (l + 1); }
static int CD(int l) { return (l + 1) * (l + 2) * (l + 3) / 6; }
static int IJK(int l) { return (l + 1) * (l + 2) / 2; }
static int CIDX(int l, int c) { return CD(l - 1) + c; }
static int LMI(int l, int m) { return l * l + m; }

typedef struct {
  int l;
  double a, d, n;
} Gauss;
typedef struct {
  int L, N;
  Gauss *g;
} Pot;

typedef struct {
  int order, n;
  double *x, *w;
  double tol, I;
  int start, end;
} Grid;

/* work counters for the algorithmic-flop figure */
typedef struct {
  double triples_exec, callbacks;
  double ps93_calls, ps93_fail, psm92_calls, psm92_fail;
  double P_T_used, P_T_all, P_Q2_used, P_Q2_all, P_Q1;
  double tab2_touched, tab2_tabulated, tab1s_touched, tab1s_tabulated, tab1l_touched, tab1l_tabulated;
  double flops_tab;  /* weighted flops of touched table points                       */
  double flops_Ftab; /* F / U_tab / U_L tables                                        */
  double M_link, M_chi, M_poly;
  double T_quads_used, T_quads_all;
  double fb_pairs; /* type-2 fallback primitive pairs */
  double t1_pairs;
  double stale_center_hits; /* fallback calls whose centre point lay beyond the cut (stale read) */
  double flops_tab2, flops_tab1; /* split of flops_tab: type-2 fallback tables / type-1 tables */
} Counters;
#define NCOUNTERS ((int)(sizeof(Counters) / sizeof(double)))

struct OracleECP {
  int nrAtoms;
  double *geometry;
  Pot **U;
  int *shells, *am, *contraction;
  double *d, *a;
  int nrShells, maxShells;
  int maxLECP, maxLBS, maxAlpha, maxLambda, tmDim, ijkDim;
  int *maxLAtom;
  double *fac, *dfac;
  int nfac;
  int *ijk, *ijkIndex;
  double *cart2sph, *poly2sph, *omega;
  int ncart2sph, npoly2sph, nomega;
  Grid *small1, *small2, *large;
  /* bessel */
  int bLMax, bN, bDim;
  double *bK, *bC;
  double tolerance, accuracy, lnAccuracy1, accuracy2, lnAccuracy2;
  int stale;
  int dims[8];
  Counters cnt;
};

/* ------------------------------------------------------------------------------------------ */
/* factorials: src/util.c:13-56 (note f[0]/f[1] quirk for tiny n, irrelevant for real shapes)    */
static double *fact_table(int n, int inc) {
  double *f = calloc(n + 1, sizeof(double));
  int i;
  if (n > 0) f[0] = 1.0;
  if (n > 1) f[1] = 1.0;
  for (i = 2; i <= n; i++) {
    f[i] = f[i - inc] * i;
    if (isinf(f[i])) abort();
  }
  return f;
}
/* binomial: src/util.c:61-70 */
static double binom(int n, int k, const double *fac) {
  if (k >= 0 && k <= n) return fac[n] / (fac[n - k] * fac[k]);
  return 0.0;
}

/* Cartesian ordering + inverse: src/dimensions.c:17-57 */
static int *shell_order(int am) {
  int *t = calloc(3 * CD(am), sizeof(int));
  int l, i, j;
  for (l = 0; l <= am; l++) {
    int c = 0;
    for (i = 0; i <= l; i++)
      for (j = 0; j <= i; j++) {
        int p = 3 * CIDX(l, c);
        t[p] = l - i;
        t[p + 1] = i - j;
        t[p + 2] = j;
        c++;
      }
  }
  return t;
}
static int *shell_order_index(int am, const int *ijk) {
  int dim = am + 1, l, c;
  int *t = calloc(dim * dim * dim, sizeof(int));
  for (l = 0; l <= am; l++)
    for (c = 0; c < IJK(l); c++) {
      int p = 3 * CIDX(l, c);
      t[ijk[p] * dim * dim + ijk[p + 1] * dim + ijk[p + 2]] = p / 3;
    }
  return t;
}

/* packed cart2sph offsets: src/transformations.h:13-14 (evaluated in double exactly like the macro) */
static int tm_dim(int l, const double *fac) { return (int)((3 * (l) + 2) * fac[l + 3] / (12 * fac[l])); }
static int tm_index(int l, int m, int c, const double *fac) { return l == 0 ? 0 : (tm_dim(l - 1, fac) + m * IJK(l) + c); }

/* Schlegel-Frisch Cartesian -> real spherical: src/transformations.c:28-87 */
static double *make_cart2sph(int lmax, const int *xyz, const double *fac, int *len) {
  int n = tm_dim(lmax, fac);
  double *TM = calloc(n, sizeof(double)), *T = TM;
  int l, m, c, i, k;
  for (l = 0; l <= lmax; l++)
    for (m = -l; m <= l; m++) {
      int mm = abs(m);
      for (c = 0; c < IJK(l); c++) {
        int p = 3 * CIDX(l, c);
        int lx = xyz[p], ly = xyz[p + 1], lz = xyz[p + 2];
        int j = lx + ly - mm;
        if (j < 0 || j % 2 == 1)
          *T = 0.0;
        else {
          double s1 = 0.0;
          j = j / 2;
          for (i = 0; i <= (l - mm) / 2; i++) {
            double s2 = 0.0;
            for (k = 0; k <= j; k++) {
              double s = 0.0;
              if ((m < 0 && abs(mm - lx) % 2 == 1) || (m > 0 && abs(mm - lx) % 2 == 0)) {
                int e = (mm - lx + 2 * k) / 2;
                s = pow(-1.0, e) * sqrt(2.0);
              } else if (m == 0 && lx % 2 == 0) {
                int e = -lx / 2 + k;
                s = pow(-1.0, e);
              }
              s2 += binom(j, k, fac) * binom(mm, (lx - 2 * k), fac) * s;
            }
            s1 += binom(l, i, fac) * binom(i, j, fac) * pow(-1.0, i) * fac[2 * l - 2 * i] / fac[l - mm - 2 * i] * s2;
          }
          *T = sqrt((fac[2 * lx] * fac[2 * ly] * fac[2 * lz] * fac[l] * fac[l - mm]) /
                    (fac[2 * l] * fac[lx] * fac[ly] * fac[lz] * fac[l + mm])) *
               1 / (pow(2.0, l) * fac[l]) * s1;
        }
        T++;
      }
    }
  *len = n;
  return TM;
}

/* unit-sphere monomial -> S_lm expansion (FM06 eq. 36): src/transformations.c:146-207 */
static double *make_poly2sph(const double *M, int lmax, const int *xyz, const double *fac2, int *len) {
  int ldim = LD(lmax), ncd = CD(lmax);
  double *TM = calloc((size_t)ncd * ldim, sizeof(double));
  int l1, c1, l2, m, c2;
  for (l1 = 0; l1 <= lmax; l1++)
    for (c1 = 0; c1 < IJK(l1); c1++) {
      int p = 3 * CIDX(l1, c1);
      int lx1 = xyz[p], ly1 = xyz[p + 1], lz1 = xyz[p + 2];
      double *T = TM + (size_t)CIDX(l1, c1) * ldim;
      const double *TCS = M;
      for (l2 = 0; l2 <= l1; l2++) {
        double s1 = 4.0 * M_PI * fac2[2 * l2 + 1];
        for (m = 0; m < 2 * l2 + 1; m++) {
          double sum = 0.0;
          for (c2 = 0; c2 < IJK(l2); c2++) {
            int q = 3 * CIDX(l2, c2);
            int lx2 = xyz[q], ly2 = xyz[q + 1], lz2 = xyz[q + 2];
            int lx = lx1 + lx2, ly = ly1 + ly2, lz = lz1 + lz2;
            if (lx % 2 == 0 && ly % 2 == 0 && lz % 2 == 0) {
              double s = s1, s2 = 1.0 / fac2[l1 + l2 + 1];
              if (lx > 2) s2 *= fac2[lx - 1];
              if (ly > 2) s2 *= fac2[ly - 1];
              if (lz > 2) s2 *= fac2[lz - 1];
              if (lx2 > 1) s /= fac2[2 * lx2 - 1];
              if (ly2 > 1) s /= fac2[2 * ly2 - 1];
              if (lz2 > 1) s /= fac2[2 * lz2 - 1];
              sum += sqrt(s) * s2 * (*TCS);
            }
            TCS++;
          }
          *T++ = sum;
        }
      }
    }
  *len = ncd * ldim;
  return TM;
}

/* ------------------------------------------------------------------------------------------ */
/* Gauss-Chebyshev grids: src/gc_integrators.c:89-145 (PSM92), :220-283 (PS93), :301-331 (maps)  */
static Grid *grid_psm92(int maxPoints, double tol) {
  int order = pow(2, floor(log(maxPoints + 1) / log(2))) - 1;
  int runs = (int)floor(log(order) / log(2));
  int offset = (int)pow(2, runs);
  int M = (order - 1) / 2, n = 1, i, idx;
  double N = n + 1.0, S0 = 1.0, C0 = 0.0, S1, C1, s, c, t;
  Grid *T = calloc(1, sizeof(Grid));
  double *x = T->x = calloc(order, sizeof(double));
  double *w = T->w = calloc(order, sizeof(double));
  T->n = T->order = order;
  T->start = 0;
  T->end = order - 1;
  T->tol = tol;
  x[M] = 0.0;
  w[M] = 1.0;
  while (n <= M) {
    C1 = C0;
    S1 = S0;
    C0 = sqrt((1 + C1) / 2);
    S0 = S1 / (2 * C0);
    s = S0;
    c = C0;
    offset /= 2;
    for (i = 1; i <= n; i += 2) {
      t = 1 + 2 / (3 * M_PI) * (3 + 2 * s * s) * s * c - i / N;
      idx = i * offset - 1;
      x[order - idx - 1] = t;
      x[idx] = -t;
      w[order - idx - 1] = w[idx] = s * s * s * s;
      t = s;
      s = s * C1 + c * S1;
      c = c * C1 - t * S1;
    }
    n = 2 * n + 1;
    N = n + 1.0;
  }
  return T;
}

static Grid *grid_ps93(int maxPoints, double tol) {
  int runs = (int)floor(log(maxPoints) / log(2));
  int offset = (int)pow(2, runs);
  int i, idx, n, order;
  double C0, S0, S1, C1, s, s2, c, t;
  Grid *T = calloc(1, sizeof(Grid));
  double *x, *w;
  T->order = order = 3 * offset - 1;
  T->start = 0;
  T->end = order - 1;
  T->n = maxPoints;
  T->tol = tol;
  x = T->x = calloc(order, sizeof(double));
  w = T->w = calloc(order, sizeof(double));
  n = 3;
  C0 = sin(M_PI / 3);
  S0 = 0.5;
  C1 = S0;
  S1 = C0;
  c = cos(M_PI / 3);
  s = C0;
  s2 = s * s;
  x[order / 2] = 0.0;
  w[order / 2] = 1.0;
  t = (n - 2.0) / n + 2 / M_PI * (1 + 2 * s2 / 3) * c * s;
  x[offset - 1] = -t;
  x[order - offset] = t;
  w[order - offset] = w[offset - 1] = s2 * s2;
  while ((4 * n / 3 - 1) <= order) {
    c = C0;
    s = S0;
    offset /= 2;
    for (i = 1; i < n; i += 2) {
      s2 = s * s;
      idx = i * offset - 1;
      t = 1 + 2 / (3 * M_PI) * s * c * (3 + 2 * s2) - ((double)i) / n;
      x[idx] = -t;
      x[order - idx - 1] = t;
      w[order - idx - 1] = w[idx] = s2 * s2;
      t = s;
      s = s * C1 + c * S1;
      c = c * C1 - t * S1;
    }
    n *= 2;
    C1 = C0;
    S1 = S0;
    C0 = sqrt((1 + C0) / 2);
    S0 = S0 / (2 * C0);
  }
  return T;
}

static void map_kk(int n, double *x, double *w) {
  double ln2 = log(2.0);
  int i;
  for (i = 0; i < n; i++) {
    double xi = 1.0 - log(1.0 - x[i]) / ln2;
    double wi = w[i] / (ln2 * (1.0 - x[i]));
    x[i] = xi;
    w[i] = wi;
  }
}

static void map_fm06(int n, double *x, double *w, double zeta_P, double P) {
  double sigma = 1.0 / sqrt(zeta_P);
  double t = P - 7.0 * sigma;
  double rmin = (t > 0.0) ? t : 0.0;
  double rmax = P + 9.0 * sigma;
  double i1 = 0.5 * (rmax - rmin), i2 = 0.5 * (rmax + rmin);
  int i;
  for (i = 0; i < n; i++) {
    x[i] = i1 * x[i] + i2;
    w[i] *= i1;
  }
}

static Grid *grid_copy(const Grid *O) {
  Grid *G = calloc(1, sizeof(Grid));
  *G = *O;
  G->x = malloc(O->order * sizeof(double));
  G->w = malloc(O->order * sizeof(double));
  memcpy(G->x, O->x, O->order * sizeof(double));
  memcpy(G->w, O->w, O->order * sizeof(double));
  return G;
}
static void grid_free(Grid *g) {
  if (!g) return;
  free(g->x);
  free(g->w);
  free(g);
}

/* integrand abstraction: value at grid index */
typedef double (*IntegrandFn)(int index, void *params);

/* PSM92 doubling rule: src/gc_integrators.c:38-86.  npts (optional) returns evaluated points. */
static int quad_psm92(IntegrandFn f, void *ps, Grid *t, int *npts) {
  int order = pow(2, floor(log(t->n + 1) / log(2))) - 1;
  double *w = t->w;
  int runs = (int)floor(log(order) / log(2));
  int offset = (int)pow(2, runs);
  const int M = (order - 1) / 2;
  double I = w[M] * f(M, ps);
  int n = 1, idx, i, cnt, np = 1;
  double N = n + 1.0, e, T, q, p = I;
  while (n <= M) {
    q = 2 * p;
    p = 2 * I;
    offset /= 2;
    cnt = 0;
    for (i = 1; i <= n; i += 2) {
      idx = i * offset - 1;
      T = 0.0;
      if (idx >= t->start) {
        T += w[idx] * f(idx, ps);
        cnt++;
      }
      if (order - idx - 1 <= t->end) {
        T += w[order - idx - 1] * f(order - idx - 1, ps);
        cnt++;
      }
      I += T;
    }
    np += cnt;
    n = 2 * n + 1;
    N = n + 1.0;
    e = I - p;
    if (0 == cnt) continue;
    if (16 * e * e <= 3 * N * fabs(I - q) * t->tol) {
      t->I = 16 * I / (3 * N);
      if (npts) *npts = np;
      return 0;
    }
  }
  if (npts) *npts = np;
  return 1;
}

/* PS93 two-point / one-point sequence: src/gc_integrators.c:156-217 */
static int quad_ps93(IntegrandFn f, void *ps, Grid *t, int *npts) {
  int runs = (int)floor(log(t->n) / log(2));
  int offset = (int)pow(2, runs);
  int order = 3 * offset - 1;
  double *w = t->w;
  int n = 3, j = 0, i, cnt, idx = offset - 1, np = 3;
  double p, q, I, err = 0.0, T;
  p = w[(order - 1) / 2] * f((order - 1) / 2, ps);
  q = w[idx] * f(idx, ps) + w[order - offset] * f(order - offset, ps);
  I = p + q;
  offset /= 2;
  while ((2 * n * (1 - j) + j * 4 * n / 3 - 1) <= order) {
    j = 1 - j;
    if (0 == j) offset /= 2;
    cnt = 0;
    for (i = 1; i < n; i += 2) {
      if (3 * ((i + 2 * j) / 3) >= i + j) {
        idx = i * offset - 1;
        T = 0.0;
        if (idx >= t->start) {
          T += w[idx] * f(idx, ps);
          cnt++;
        }
        if (order - idx - 1 <= t->end) {
          T += w[order - idx - 1] * f(order - idx - 1, ps);
          cnt++;
        }
        I += T;
      }
    }
    np += cnt;
    n *= (1 + j);
    p += (1 - j) * (I - q);
    if (0 < cnt) err = 16 * fabs((1 - j) * (q - 3 * p / 2) + j * (I - 2 * q)) / (3 * n);
    q = (1 - j) * q + j * I;
    if (0 == cnt) continue;
    if (err < t->tol) {
      t->I = 16 * q / (3 * n);
      if (npts) *npts = np;
      return 0;
    }
  }
  if (npts) *npts = np;
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Bessel table: src/bessel.c:19-82 */
static int bessel_tabulate(OracleECP *h, int lMax, int N, int cutoff, double accuracy) {
  int i, j, l, m, dim = N + 1;
  double z, f, s;
  double *F = malloc((cutoff + 1) * sizeof(double));
  double *G = malloc((cutoff + lMax + 2) * sizeof(double));
  double *K = calloc((size_t)(lMax + 1) * dim, sizeof(double));
  double *C = calloc(lMax + 1, sizeof(double));
  h->bLMax = lMax;
  h->bN = N;
  h->bDim = dim;
  h->bK = K;
  h->bC = C;
  K[0] = 1.0;
  for (i = 1; i <= N; i++) {
    z = i / (N / 16.0);
    j = 0;
    f = z * z / 2.0;
    F[j] = exp(-z);
    G[j] = 1.0;
    s = F[j] / G[j];
    l = (int)(0.25 * sqrt(1.0 + 16.0 * f));
    while (s > accuracy || j <= l) {
      K[i] += s;
      j++;
      if (j > cutoff) {
        free(F);
        free(G);
        return 1;
      }
      F[j] = F[j - 1] * f / j;
      G[j] = G[j - 1] * (2.0 * j + 1.0);
      s = F[j] / G[j];
    }
    for (l = 1; l <= lMax; l++) G[j + l] = G[j + l - 1] * (2 * j + 2 * l + 1);
    f = z;
    for (l = 1; l <= lMax; l++) {
      s = 0;
      for (m = 0; m < j; m++) s += F[m] / G[l + m];
      K[l * dim + i] = f * s;
      f *= z;
    }
  }
  for (i = 1; i <= lMax; i++) C[i] = i / (2.0 * i + 1.0);
  free(F);
  free(G);
  return 0;
}

/* weighted Bessel K_0..K_lmax(z), output stride dim: src/bessel.c:101-199. returns branch 0/1/2 */
static int bessel_eval(const OracleECP *h, int dim, int lmax, double z, double *K) {
  const double small = 1.0E-7;
  int i, j, l, branch;
  int ddim = h->bDim;
  if (z < small) {
    branch = 0;
    if (z <= 0) {
      K[0] = 1.0;
      for (l = 1; l <= lmax; l++) K[l * dim] = 0.0;
    } else {
      K[0] = 1 - z;
      for (l = 1; l <= lmax; l++) K[l * dim] = K[(l - 1) * dim] * z / (2 * l + 1);
    }
  } else if (z < 16.0) {
    double dKi[64], dKj[64];
    int maxLambda = lmax + 5;
    double scale = h->bN / 16; /* integer division, = 100 */
    int index = (int)floor(z * scale + 0.5);
    double dz = z - index / scale;
    branch = 1;
    scale = 1.0;
    for (l = 0; l <= lmax; l++) K[l * dim] = dKi[l] = h->bK[l * ddim + index];
    for (l = lmax + 1; l <= maxLambda; l++) dKi[l] = h->bK[l * ddim + index];
    for (i = 1; i <= 5; i++) {
      index = maxLambda - i;
      for (j = 0; j <= index + 1; j++) dKj[j] = dKi[j];
      dKi[0] = dKj[1] - dKj[0];
      for (j = 1; j <= index; j++) dKi[j] = h->bC[j] * (dKj[j - 1] - dKj[j + 1]) - dKj[j] + dKj[j + 1];
      scale = scale * dz / i;
      for (j = 0; j <= lmax; j++) K[j * dim] += scale * dKi[j];
    }
  } else {
    double A[64], f;
    branch = 2;
    A[0] = 0.5 / z;
    for (l = 0; l <= lmax; l++) K[l * dim] = A[0];
    for (l = 1; l <= lmax; l++) {
      f = l * (l + 1);
      for (i = 1; i < l; i++) {
        K[l * dim] += f * A[i];
        f *= (l + i + 1) * (l - i);
      }
      A[l] = -A[0] * A[l - 1] / l;
      K[l * dim] += f * A[l];
    }
  }
  return branch;
}
/* flop weights of one Bessel evaluation */
static double wB(int branch, int lmax) {
  if (branch == 0) return 2.0 * lmax + 1;
  if (branch == 1) return 35.0 * lmax + 95;
  return 2.0 * lmax * lmax + 3.0 * lmax + 4;
}

/* ------------------------------------------------------------------------------------------ */
/* ECP radial potential: src/ecp.c:41-60 */
static double pot_eval(const Pot *U, int l, double r) {
  double v = 0.0, r2 = r * r;
  int i;
  for (i = 0; i < U->N; i++)
    if (l == U->g[i].l) v += pow(r, U->g[i].n) * U->g[i].d * exp(-U->g[i].a * r2);
  return v;
}
static int pot_count(const Pot *U, int l) {
  int i, k = 0;
  for (i = 0; i < U->N; i++) k += (U->g[i].l == l);
  return k;
}

/* geometry helpers: src/util.c:74-124 */
void oracle_sphcoord(const double *v, double *s) {
  const double eps = 1.0E-14;
  double x = v[0], y = v[1], z = v[2], r, theta, phi;
  r = sqrt(x * x + y * y + z * z);
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
  s[0] = r;
  s[1] = theta;
  s[2] = phi;
}
static double dist(const double *A, const double *B) {
  double x = A[0] - B[0], y = A[1] - B[1], z = A[2] - B[2];
  return sqrt(x * x + y * y + z * z);
}

/* shell radius by Newton iteration: src/util.c:133-192 */
static double gto_radius(double c, double zeta, int l, double cutoff) {
  int i, converged = 0;
  double delta, dg, guess, zr, r;
  const double t = log(fabs(c) / fabs(cutoff));
  dg = t / zeta;
  r = (dg > cutoff) ? dg : cutoff;
  if (0 != l) {
    if (0 < l) {
      dg = sqrt(0.5 * l / fabs(c));
      guess = (dg > cutoff) ? dg : cutoff;
      if (guess > r) r = 0.5 * (r + guess);
    }
    for (i = 0; i < 40; i++) {
      zr = zeta * r;
      guess = t + l * log(r) - zr * r;
      delta = guess / (l / r - 2 * zr);
      dg = r - delta;
      r = (dg > cutoff) ? dg : cutoff;
      if (fabs(delta) < cutoff) {
        r *= r;
        converged = 1;
        break;
      }
    }
    if (!converged) abort();
  }
  return r;
}
double oracle_shell_radius(int depth, int am, const double *d, const double *a, double zero) {
  double zeta = a[0], c = fabs(d[0]);
  int i;
  for (i = 1; i < depth; i++)
    if (a[i] < zeta && d[i] != 0.0) {
      zeta = a[i];
      c = fabs(d[i]);
    }
  return sqrt(gto_radius(c, zeta, am, zero));
}

/* potential cut-off: src/util.c:198-210 */
static int pot_screen(int start, int end, const double *U, double zero) {
  int i;
  for (i = end; i >= start; i--)
    if (fabs(U[i]) > zero) return i;
  return -1;
}

/* real spherical harmonics: src/spherical_harmonics.c:15-114 ; index l*l + l + m */
static void rsh_eval(int lmax, double theta, double phi, const double *fac, const double *dfac, double *rsh) {
  int l, m, n = LD(lmax);
  double x, norm0, norm;
  double *s = calloc(lmax + 2, sizeof(double)), *c = calloc(lmax + 2, sizeof(double));
  double *P = calloc(n, sizeof(double));
#define RS(l, m) ((l) * (l) + (l) + (m))
  memset(rsh, 0, n * sizeof(double));
  x = cos(theta);
  if (1.0 == x) {
    for (l = 0; l <= lmax; l++) P[RS(l, 0)] = 1.0;
  } else if (-1.0 == x) {
    P[RS(0, 0)] = 1.0;
    for (l = 1; l <= lmax; l++) P[RS(l, 0)] = -P[RS(l - 1, 0)];
  } else {
    s[1] = sqrt(1.0 - x * x);
    for (l = 2; l <= lmax; l++) s[l] = s[l - 1] * s[1];
    for (l = 0; l <= lmax; l++) {
      m = l;
      if (0 == m)
        P[RS(l, 0)] = 1.0;
      else {
        P[RS(l, m)] = s[m] * dfac[2 * m - 1];
        m = l - 1;
        P[RS(l, m)] = x * (2 * m + 1) * P[RS(l - 1, m)];
        if (l > 1)
          for (m = 0; m <= l - 2; m++)
            P[RS(l, m)] = (x * (2 * l - 1) * P[RS(l - 1, m)] - (l + m - 1) * P[RS(l - 2, m)]) / (l - m);
      }
    }
  }
  for (l = 0; l <= lmax; l++) {
    norm0 = sqrt((2.0 * l + 1.0) / (2.0 * M_PI));
    P[RS(l, 0)] = norm0 * P[RS(l, 0)];
    for (m = 1; m <= l; m++) {
      norm = sqrt(fac[l - m] / fac[l + m]) * norm0;
      P[RS(l, m)] = norm * P[RS(l, m)];
    }
  }
  if (lmax > 0) {
    if (0.0 == phi) {
      for (m = 0; m <= lmax; m++) {
        s[m] = 0.0;
        c[m] = 1.0;
      }
    } else {
      s[1] = sin(phi);
      c[1] = cos(phi);
      for (m = 2; m <= lmax; m++) {
        s[m] = s[1] * c[m - 1] + c[1] * s[m - 1];
        c[m] = c[1] * c[m - 1] - s[1] * s[m - 1];
      }
    }
  }
  for (l = 0; l <= lmax; l++) {
    rsh[RS(l, 0)] = P[RS(l, 0)] / sqrt(2.0);
    for (m = 1; m <= l; m++) {
      rsh[RS(l, -m)] = P[RS(l, m)] * s[m];
      rsh[RS(l, +m)] = P[RS(l, m)] * c[m];
    }
  }
#undef RS
  free(P);
  free(s);
  free(c);
}

/* geometry-independent angular table Omega[(l,m)][(lambda,mu)][C_INDEX]: src/angular_integrals.c:15-101 */
static double *make_omega(OracleECP *h, int *len) {
  const int d1 = LD(h->maxLECP), d2 = LD(h->maxLambda), d3 = CD(h->maxAlpha);
  const int inc1 = d3, inc2 = d2 * d3, tmDim = LD(h->tmDim), ijkDim = h->ijkDim;
  double *Om = calloc((size_t)d1 * d2 * d3, sizeof(double));
  const int *ijk = h->ijk, *ijkIndex = h->ijkIndex;
  int lambda, l, alpha, mu, m, c, d;
  for (lambda = 0; lambda <= h->maxLambda; lambda++)
    for (l = 0; l < h->maxLECP; l++) {
      int par = (lambda + l) % 2, dl = lambda - l;
      int minAlpha = (par > dl) ? par : dl;
      for (alpha = minAlpha; alpha <= h->maxAlpha; alpha += 2)
        for (mu = 0; mu < 2 * lambda + 1; mu++)
          for (m = 0; m < 2 * l + 1; m++) {
            double *cijk = Om + LMI(l, m) * inc2 + LMI(lambda, mu) * inc1 + CD(alpha - 1);
            for (c = 0; c < IJK(alpha); c++) {
              int p = 3 * CIDX(alpha, c);
              int i = ijk[p], j = ijk[p + 1], k = ijk[p + 2];
              if (0 == alpha) {
                if (l == lambda && m == mu) cijk[c] = 1.0;
              } else if (lambda <= l + alpha) {
                for (d = 0; d < IJK(l); d++) {
                  int q = 3 * CIDX(l, d);
                  int ax = ijk[q], ay = ijk[q + 1], az = ijk[q + 2];
                  double N = 0.25 * h->dfac[2 * l + 1] / M_PI;
                  if (ax > 1) N /= h->dfac[2 * ax - 1];
                  if (ay > 1) N /= h->dfac[2 * ay - 1];
                  if (az > 1) N /= h->dfac[2 * az - 1];
                  N = sqrt(N);
                  ax += i;
                  ay += j;
                  az += k;
                  q = ijkIndex[ax * ijkDim * ijkDim + ay * ijkDim + az] * tmDim + LMI(lambda, mu);
                  cijk[c] += N * h->cart2sph[tm_index(l, m, d, h->fac)] * h->poly2sph[q];
                }
              }
            }
          }
    }
  *len = d1 * d2 * d3;
  return Om;
}

/* per (C, atom X): Omega_X[lambda][(l,m)][C_INDEX]: src/angular_integrals.c:104-142.
 * dims [lmax+1][LD(maxLambda)][CD(lmax_a)] */
static double *angular_for_atom(OracleECP *h, int lmax_a, const double *rX) {
  const int inc1 = CD(h->maxAlpha), inc2 = LD(h->maxLambda) * CD(h->maxAlpha);
  const int lmax = h->maxLECP - 1 + lmax_a;
  const int incA1 = CD(lmax_a), incA2 = LD(h->maxLambda) * CD(lmax_a);
  double *A = calloc((size_t)(lmax + 1) * incA2, sizeof(double));
  double R[3], *rsph = malloc(LD(lmax) * sizeof(double));
  int l, m, lambda, la, c, mu;
  oracle_sphcoord(rX, R);
  rsh_eval(lmax, R[1], R[2], h->fac, h->dfac, rsph);
  for (l = 0; l < h->maxLECP; l++)
    for (m = 0; m < 2 * l + 1; m++)
      for (lambda = 0; lambda <= lmax; lambda++)
        for (la = 0; la <= lmax_a; la++)
          for (c = 0; c < IJK(la); c++) {
            double v = 0.0;
            for (mu = 0; mu < 2 * lambda + 1; mu++)
              v += rsph[LMI(lambda, mu)] * h->omega[LMI(l, m) * inc2 + LMI(lambda, mu) * inc1 + CIDX(la, c)];
            A[lambda * incA2 + LMI(l, m) * incA1 + CIDX(la, c)] = v;
          }
  free(rsph);
  return A;
}

/* (-x)^i (-y)^j (-z)^k table: src/util.c:214-243, dims [l+1]^3 */
static double *usp_table(int l, const double *r) {
  int dim = l + 1, i, j, k;
  double x = 0, y = 0, z = 0;
  double *t = calloc(dim * dim * dim, sizeof(double));
  for (i = 0; i <= l; i++) {
    x = (0 == i) ? 1.0 : -x * r[0];
    for (j = 0; j <= l - i; j++) {
      y = (0 == j) ? 1.0 : -y * r[1];
      for (k = 0; k <= l - i - j; k++) {
        z = (0 == k) ? 1.0 : -z * r[2];
        t[i * dim * dim + j * dim + k] = x * y * z;
      }
    }
  }
  return t;
}

/* binomial shift C-centred monomials -> A/B-centred Cartesians: src/util.c:246-334.  frees nothing */
static double *shift_polynomials(OracleECP *h, const double *gamma, double N, int la, const double *uspA, int dimA,
                                 int lb, const double *uspB, int dimB) {
  const int incG = CD(lb), incJ = CD(lb), incI = IJK(lb), D = h->ijkDim;
  double *J = calloc((size_t)IJK(la) * CD(lb), sizeof(double));
  double *I = calloc((size_t)IJK(la) * IJK(lb), sizeof(double));
  const double *fac = h->fac;
  int c1, c2, beta, x, y, z, r, q;
  for (c1 = 0; c1 < IJK(la); c1++) {
    int p = 3 * CIDX(la, c1);
    int ax = h->ijk[p], ay = h->ijk[p + 1], az = h->ijk[p + 2];
    for (x = 0; x <= ax; x++) {
      double bx = binom(ax, x, fac);
      for (y = 0; y <= ay; y++) {
        double by = bx * binom(ay, y, fac);
        for (z = 0; z <= az; z++) {
          double bz = by * binom(az, z, fac);
          double factor = bz * uspA[(ax - x) * dimA * dimA + (ay - y) * dimA + (az - z)];
          if (fabs(factor) <= h->accuracy) continue;
          r = h->ijkIndex[x * D * D + y * D + z];
          for (beta = 0; beta <= lb; beta++)
            for (c2 = 0; c2 < IJK(beta); c2++) {
              q = CIDX(beta, c2);
              J[c1 * incJ + q] += factor * gamma[r * incG + q];
              h->cnt.M_poly += 1;
            }
        }
      }
    }
  }
  for (c2 = 0; c2 < IJK(lb); c2++) {
    int p = 3 * CIDX(lb, c2);
    int bx_ = h->ijk[p], by_ = h->ijk[p + 1], bz_ = h->ijk[p + 2];
    for (x = 0; x <= bx_; x++) {
      double bx = binom(bx_, x, fac);
      for (y = 0; y <= by_; y++) {
        double by = bx * binom(by_, y, fac);
        for (z = 0; z <= bz_; z++) {
          double bz = by * binom(bz_, z, fac);
          double factor = bz * uspB[(bx_ - x) * dimB * dimB + (by_ - y) * dimB + (bz_ - z)];
          if (fabs(factor) <= h->accuracy) continue;
          factor *= N;
          r = h->ijkIndex[x * D * D + y * D + z];
          for (c1 = 0; c1 < IJK(la); c1++) {
            I[c1 * incI + c2] += factor * J[c1 * incJ + r];
            h->cnt.M_poly += 1;
          }
        }
      }
    }
  }
  free(J);
  return I;
}

/* ------------------------------------------------------------------------------------------ */
/* init: src/libecp.c:53-201 (+ Type1_init src/type1.c:38-61, Type2_new/init src/type2.c:40-96) */
OracleECP *oracle_libECP_init(int nrAtoms, double *geometry, int *shellsECP, int *lECP, int *KECP, double *nECP,
                              double *dECP, double *aECP, int *shellsBS, int *lBS, int *KBS, double *dBS, double *aBS,
                              int n, int lmax, int *shellOrdering, int largeGridOrder, double tolerance,
                              double accuracy) {
  OracleECP *h = calloc(1, sizeof(OracleECP));
  int i, j, k, L, N, index, si, pi;
  (void)lmax;
  if (n != 0 || shellOrdering != NULL) { /* out of scope */
    free(h);
    return NULL;
  }
  h->nrAtoms = nrAtoms;
  h->geometry = geometry;
  h->shells = shellsBS;
  h->am = lBS;
  h->contraction = KBS;
  h->d = dBS;
  h->a = aBS;
  h->tolerance = tolerance;
  h->accuracy = accuracy;
  h->lnAccuracy1 = log(accuracy) - 2;
  h->accuracy2 = 1.0E-14; /* src/type2.c:58-59: hard-coded */
  h->lnAccuracy2 = log(h->accuracy2) - 2;
  h->stale = 1;
  h->U = calloc(nrAtoms, sizeof(Pot *));
  index = 0;
  for (i = 0; i < nrAtoms; i++)
    if (0 < shellsECP[i]) {
      L = N = 0;
      for (j = 0; j < shellsECP[i]; j++) {
        if (L < lECP[index]) L = lECP[index];
        N += KECP[index];
        index++;
      }
      h->U[i] = calloc(1, sizeof(Pot));
      h->U[i]->L = L;
      h->U[i]->N = N;
      h->U[i]->g = calloc(N, sizeof(Gauss));
      if (L > h->maxLECP) h->maxLECP = L;
    }
  si = pi = 0;
  for (i = 0; i < nrAtoms; i++)
    if (h->U[i]) {
      index = 0;
      for (j = 0; j < shellsECP[i]; j++) {
        for (k = 0; k < KECP[si]; k++) {
          Gauss *g = &h->U[i]->g[index++];
          g->l = lECP[si];
          g->n = nECP[pi];
          g->d = dECP[pi];
          g->a = aECP[pi];
          pi++;
        }
        si++;
      }
    }
  h->maxLAtom = calloc(nrAtoms, sizeof(int));
  index = 0;
  for (i = 0; i < nrAtoms; i++) {
    if (shellsBS[i] > h->maxShells) h->maxShells = shellsBS[i];
    for (j = 0; j < shellsBS[i]; j++) {
      if (h->maxLAtom[i] < lBS[index]) h->maxLAtom[i] = lBS[index];
      if (h->maxLBS < lBS[index]) h->maxLBS = lBS[index];
      index++;
    }
  }
  h->nrShells = index;
  h->maxAlpha = h->maxLBS;
  h->maxLambda = h->maxLECP - 1 + h->maxAlpha;
  h->tmDim = h->maxLambda + h->maxAlpha;
  h->nfac = 2 * h->tmDim + 2;
  h->fac = fact_table(2 * h->tmDim + 1, 1);
  h->dfac = fact_table(2 * h->tmDim + 1, 2);
  h->ijk = shell_order(h->tmDim);
  h->ijkIndex = shell_order_index(h->tmDim, h->ijk);
  h->ijkDim = h->tmDim + 1;
  h->cart2sph = make_cart2sph(h->tmDim, h->ijk, h->fac, &h->ncart2sph);
  h->poly2sph = make_poly2sph(h->cart2sph, h->tmDim, h->ijk, h->dfac, &h->npoly2sph);
  h->large = grid_psm92(largeGridOrder, tolerance);
  if (0 != bessel_tabulate(h, h->maxLECP + h->maxAlpha + 6, 16 * 100, 200, accuracy)) {
    oracle_libECP_free(h);
    return NULL;
  }
  h->small1 = grid_ps93(128, tolerance);
  map_kk(h->small1->order, h->small1->x, h->small1->w);
  h->small2 = grid_ps93(128, tolerance);
  map_kk(h->small2->order, h->small2->x, h->small2->w);
  h->omega = make_omega(h, &h->nomega); /* the reference builds it per call (src/libecp.c:254); same values */
  h->dims[0] = h->maxLECP;
  h->dims[1] = h->maxLBS;
  h->dims[2] = h->maxAlpha;
  h->dims[3] = h->maxLambda;
  h->dims[4] = h->tmDim;
  h->dims[5] = h->bLMax;
  h->dims[6] = h->small1->order;
  h->dims[7] = h->large->order;
  return h;
}

void oracle_libECP_free(OracleECP *h) {
  int i;
  if (!h) return;
  if (h->U) {
    for (i = 0; i < h->nrAtoms; i++)
      if (h->U[i]) {
        free(h->U[i]->g);
        free(h->U[i]);
      }
    free(h->U);
  }
  free(h->maxLAtom);
  free(h->fac);
  free(h->dfac);
  free(h->ijk);
  free(h->ijkIndex);
  free(h->cart2sph);
  free(h->poly2sph);
  free(h->omega);
  grid_free(h->small1);
  grid_free(h->small2);
  grid_free(h->large);
  free(h->bK);
  free(h->bC);
  free(h);
}

/* ------------------------------------------------------------------------------------------ */
/* per-centre tables */

/* U_tab[l][N][n] = r^N U_l(r_n) with cumulative potential cut-off: src/type2.c:184-219.
 * dims [L][lmax+1][order]; narrows grid->end */
static double *tab_potential(OracleECP *h, const Pot *U, Grid *grid, int lmax, double tol, int *end_l) {
  const int order = grid->order, inc1 = order, inc2 = (lmax + 1) * order;
  double *tab = calloc((size_t)U->L * inc2, sizeof(double));
  double *Ul = calloc(order, sizeof(double));
  int l, n, lab;
  for (l = 0; l < U->L; l++) {
    for (n = 0; n < order; n++) Ul[n] = pot_eval(U, l, grid->x[n]);
    h->cnt.flops_Ftab += 24.0 * pot_count(U, l) * order;
    grid->end = pot_screen(grid->start, grid->end, Ul, tol);
    if (end_l) end_l[l] = grid->end;
    for (n = grid->start; n <= grid->end; n++) {
      double *P = tab + l * inc2 + n;
      P[0] = Ul[n];
      for (lab = 1; lab <= lmax; lab++) P[lab * inc1] = P[(lab - 1) * inc1] * grid->x[n];
    }
  }
  free(Ul);
  return tab;
}

typedef struct {
  int start, end, skip;
} Window;

/* src/type2.c:148-180 */
static Window basis_window(const Grid *grid, double R, double d) {
  const double rmin = d - R, rmax = d + R;
  const double *r = grid->x;
  Window s;
  int j;
  for (j = grid->end; j >= grid->start; j--)
    if (r[j] < rmin) break;
  s.start = j + 1;
  for (j = grid->end; j >= grid->start; j--)
    if (r[j] <= rmax) break;
  s.end = (j < 0 || r[j] > rmax) ? -1 : j;
  s.skip = !(s.end >= s.start);
  return s;
}

/* screening windows + contracted radial table F[shell][lambda][n]: src/type2.c:222-310 (n = 0) */
static double *tab_F(OracleECP *h, const double *rC, Window *sg) {
  const int order = h->small2->order, inc1 = order, inc2 = (h->maxLambda + 1) * order;
  double *F = calloc((size_t)h->nrShells * inc2, sizeof(double));
  double *K = calloc(h->maxLambda + 1, sizeof(double));
  const double *r = h->small2->x;
  int *skipAtom = calloc(h->nrAtoms, sizeof(int));
  int *primOff = calloc(h->nrShells, sizeof(int));
  int A, sa, pa, n, i, index = 0, offset = 0;
  const int lmax = h->maxLECP - 1;
  for (A = 0; A < h->nrAtoms; A++) {
    double dAC = dist(rC, &h->geometry[A * 3]);
    skipAtom[A] = 1;
    for (sa = 0; sa < h->shells[A]; sa++) {
      double e = oracle_shell_radius(h->contraction[index], h->am[index], &h->d[offset], &h->a[offset], h->accuracy2);
      sg[index] = basis_window(h->small2, e, dAC);
      if (!sg[index].skip) skipAtom[A] = 0;
      primOff[index] = offset;
      offset += h->contraction[index];
      index++;
    }
  }
  index = -1;
  for (A = 0; A < h->nrAtoms; A++) {
    double dAC;
    if (skipAtom[A]) {
      index += h->shells[A];
      continue;
    }
    dAC = dist(rC, &h->geometry[A * 3]);
    for (sa = 0; sa < h->shells[A]; sa++) {
      int lmaxA;
      index++;
      if (sg[index].skip) continue;
      lmaxA = lmax + h->am[index];
      for (pa = 0; pa < h->contraction[index]; pa++) {
        double zeta = h->a[primOff[index] + pa], da = h->d[primOff[index] + pa];
        for (n = sg[index].start; n < sg[index].end; n++) {
          double e;
          int br = bessel_eval(h, 1, lmaxA, 2.0 * zeta * dAC * r[n], K);
          e = dAC - r[n];
          e = exp(-zeta * e * e);
          for (i = 0; i <= lmaxA; i++) F[index * inc2 + i * inc1 + n] += da * K[i] * e;
          h->cnt.flops_Ftab += wB(br, lmaxA) + 20 + 4 + 2 * (lmaxA + 1);
        }
      }
    }
  }
  free(skipAtom);
  free(primOff);
  free(K);
  return F;
}

/* ------------------------------------------------------------------------------------------ */
/* type 2 */
typedef struct {
  const double *Fa, *Fb, *U;
  double *pts; /* counter */
} TParams;
static double integrand_T(int n, void *v) { /* src/type2.c:319-322 */
  TParams *p = v;
  if (p->pts) *p->pts += 1;
  return p->Fa[n] * p->Fb[n] * p->U[n];
}

typedef struct {
  double C, minExp;
  const double *expo, *Ka, *Kb, *U, *rn;
  unsigned char *touched; /* NULL when the quadrature is not "used" */
  double *pts;
} Q2Params;
static double integrand_Q2(int n, void *v) { /* src/type2.c:397-410 */
  Q2Params *p = v;
  double Q = 0.0, e = p->expo[n];
  if (p->touched) p->touched[n] = 1;
  if (e >= p->minExp) {
    if (p->pts) *p->pts += 1;
    Q = p->C * p->U[n] * p->rn[n] * p->Ka[n] * p->Kb[n] * exp(e);
  }
  return Q;
}

/* is T[l1][l2][l3] ever multiplied by a non-zero angular factor?  (link loops src/type2.c:583-623,
 * zeros of Omega src/angular_integrals.c:40-44,63) */
static int t2_used(int la, int lb, int l, int l1, int l2, int l3) {
  int alpha;
  for (alpha = 0; alpha <= la; alpha++) {
    int beta = l3 - alpha;
    if (beta < 0 || beta > lb) continue;
    if ((alpha + l + l1) % 2 || (beta + l + l2) % 2) continue;
    if (l1 < l - alpha || l2 < l - beta) continue;
    if (l1 > l + alpha || l2 > l + beta) continue;
    return 1;
  }
  return 0;
}

/* fallback buffers live for one calcT_FM06 call (= one (triple,l)) and are reused across primitive
 * pairs without clearing: src/type2.c:443-448 */
typedef struct {
  double *expo, *U, *Ka, *Kb, *rn;
  unsigned char *touched, *brA, *brB;
} FallbackBuf;

/* src/type2.c:417-528; returns 0 ok / 1 large-grid failure */
static int t2_fallback(OracleECP *h, double *T, int Tinc1, int Tinc2, int nFailed, const int *fl1, const int *fl2,
                       const int *fl3, const Pot *U, int l, double dAC, int alpha, int shella, int offa, double dBC,
                       int beta, int shellb, int offb) {
  const int Na = h->contraction[shella], Nb = h->contraction[shellb];
  const double *za = h->a + offa, *ca = h->d + offa, *zb = h->a + offb, *cb = h->d + offb;
  const int laC = alpha + l, lbC = beta + l, lab = alpha + beta, G = h->large->order;
  FallbackBuf b;
  Q2Params ps;
  int pa, pb, n, l3, k, rc = 0;
  b.expo = calloc(G, sizeof(double));
  b.U = calloc(G, sizeof(double));
  b.Ka = calloc((size_t)(laC + 1) * G, sizeof(double));
  b.Kb = calloc((size_t)(lbC + 1) * G, sizeof(double));
  b.rn = calloc((size_t)(lab + 1) * G, sizeof(double));
  b.touched = calloc(G, 1);
  b.brA = calloc(G, 1);
  b.brB = calloc(G, 1);
  ps.minExp = h->lnAccuracy2;
  ps.expo = b.expo;
  ps.U = b.U;
  for (pa = 0; pa < Na && !rc; pa++) {
    double s1 = 2.0 * za[pa] * dAC;
    for (pb = 0; pb < Nb && !rc; pb++) {
      double s2 = 2.0 * zb[pb] * dBC, zp, p;
      Grid *grid = grid_copy(h->large);
      int ntab = 0;
      ps.C = ca[pa] * cb[pb];
      zp = za[pa] + zb[pb];
      p = (za[pa] * dAC + zb[pb] * dBC) / zp;
      map_fm06(grid->order, grid->x, grid->w, zp, p);
      if (!h->stale) { /* "clean" variant: untabulated points contribute exactly 0 */
        memset(b.expo, 0, G * sizeof(double));
        memset(b.U, 0, G * sizeof(double));
      }
      memset(b.touched, 0, G);
      for (n = 0; n < G; n++) {
        double r = grid->x[n], d1 = dAC - r, d2 = dBC - r;
        b.expo[n] = -za[pa] * d1 * d1 - zb[pb] * d2 * d2;
        if (r > dAC && r > dBC && b.expo[n] < ps.minExp) {
          grid->end = n - 1;
          break;
        } else if (b.expo[n] >= ps.minExp) {
          b.U[n] = pot_eval(U, l, r);
          b.brA[n] = bessel_eval(h, G, laC, s1 * r, b.Ka + n);
          b.brB[n] = bessel_eval(h, G, lbC, s2 * r, b.Kb + n);
          b.rn[n] = 1.0;
          for (l3 = 1; l3 <= lab; l3++) b.rn[l3 * G + n] = r * b.rn[(l3 - 1) * G + n];
          ntab++;
        }
      }
      if (grid->end < (G - 1) / 2) h->cnt.stale_center_hits += 1;
      h->cnt.tab2_tabulated += ntab;
      h->cnt.fb_pairs += 1;
      for (k = nFailed - 1; k >= 0; k--) {
        int l1 = fl1[k], l2 = fl2[k], np = 0;
        int used = t2_used(alpha, beta, l, l1, l2, fl3[k]);
        ps.Ka = b.Ka + l1 * G;
        ps.Kb = b.Kb + l2 * G;
        ps.rn = b.rn + fl3[k] * G;
        ps.touched = used ? b.touched : NULL;
        ps.pts = used ? &h->cnt.P_Q2_used : &h->cnt.P_Q2_all;
        h->cnt.psm92_calls += 1;
        if (0 != quad_psm92(integrand_Q2, &ps, grid, &np)) {
          h->cnt.psm92_fail += 1;
          rc = 1;
          break;
        }
        T[l1 * Tinc2 + l2 * Tinc1 + fl3[k]] += grid->I;
      }
      /* flop weight of the table points a used quadrature really read */
      for (n = 0; n < G; n++)
        if (b.touched[n] && b.expo[n] >= ps.minExp) {
          h->cnt.tab2_touched += 1;
          h->cnt.flops_tab += 20 + wB(b.brA[n], laC) + wB(b.brB[n], lbC) + 24.0 * pot_count(U, l) + lab;
          h->cnt.flops_tab2 += 20 + wB(b.brA[n], laC) + wB(b.brB[n], lbC) + 24.0 * pot_count(U, l) + lab;
        }
      grid_free(grid);
    }
  }
  free(b.expo);
  free(b.U);
  free(b.Ka);
  free(b.Kb);
  free(b.rn);
  free(b.touched);
  free(b.brA);
  free(b.brB);
  return rc;
}

/* gamma[C_DIM(la)][C_DIM(lb)]: src/type2.c:532-631 ; returns NULL on large-grid failure */
static double *type2_gamma(OracleECP *h, const double *F, const double *UTab, const Pot *U, double dAC, double dBC,
                           int la, int shella, int offa, const double *omegaA, int lmaxA, int lb, int shellb, int offb,
                           const double *omegaB, int lmaxB) {
  const int order = h->small2->order;
  const int Finc1 = order, Finc2 = (h->maxLambda + 1) * order;
  const int Uinc1 = order, Uinc2 = (h->maxLambda + 1) * order;
  const int incA1 = CD(lmaxA), incA2 = LD(h->maxLambda) * CD(lmaxA);
  const int incB1 = CD(lmaxB), incB2 = LD(h->maxLambda) * CD(lmaxB);
  const int inc = CD(lb);
  double *gamma = calloc((size_t)CD(la) * CD(lb), sizeof(double));
  const double *Fa = F + shella * Finc2, *Fb = F + shellb * Finc2;
  int l, l1, l2, l3, alpha, beta, c1, c2, m;
  for (l = 0; l < U->L; l++) {
    const int laC = la + l, lbC = lb + l, lab = la + lb;
    const int Tinc1 = lab + 1, Tinc2 = (lbC + 1) * (lab + 1);
    const int maxFailed = (laC + 1) * (lbC + 1) * (lab + 1);
    double *T = calloc(maxFailed, sizeof(double));
    int *fl1 = calloc(maxFailed, sizeof(int)), *fl2 = calloc(maxFailed, sizeof(int)),
        *fl3 = calloc(maxFailed, sizeof(int));
    int nFailed = 0, rc = 0;
    TParams ps;
    /* fast path on the small grid: src/type2.c:336-381 */
    for (l1 = 0; l1 <= laC; l1++) {
      ps.Fa = Fa + l1 * Finc1;
      for (l2 = 0; l2 <= lbC; l2++) {
        ps.Fb = Fb + l2 * Finc1;
        for (l3 = 0; l3 <= lab; l3++) {
          int used = t2_used(la, lb, l, l1, l2, l3);
          ps.U = UTab + l * Uinc2 + l3 * Uinc1;
          ps.pts = used ? &h->cnt.P_T_used : &h->cnt.P_T_all;
          h->cnt.ps93_calls += 1;
          if (used) h->cnt.T_quads_used += 1;
          h->cnt.T_quads_all += 1;
          if (0 != quad_ps93(integrand_T, &ps, h->small2, NULL)) {
            h->cnt.ps93_fail += 1;
            fl1[nFailed] = l1;
            fl2[nFailed] = l2;
            fl3[nFailed] = l3;
            nFailed++;
          } else
            T[l1 * Tinc2 + l2 * Tinc1 + l3] = h->small2->I;
        }
      }
    }
    if (nFailed > 0)
      rc = t2_fallback(h, T, Tinc1, Tinc2, nFailed, fl1, fl2, fl3, U, l, dAC, la, shella, offa, dBC, lb, shellb, offb);
    free(fl1);
    free(fl2);
    free(fl3);
    if (rc) {
      free(T);
      free(gamma);
      return NULL;
    }
    /* link angular and radial parts: src/type2.c:583-623 */
    for (alpha = 0; alpha <= la; alpha++) {
      int parity = (alpha + l) % 2, ll1 = l - alpha;
      ll1 = (parity > ll1) ? parity : ll1;
      for (c1 = 0; c1 < IJK(alpha); c1++) {
        int p = CIDX(alpha, c1);
        for (beta = 0; beta <= lb; beta++) {
          int par2 = (beta + l) % 2, ll2 = l - beta;
          ll2 = (par2 > ll2) ? par2 : ll2;
          for (c2 = 0; c2 < IJK(beta); c2++) {
            int q = CIDX(beta, c2);
            double tmp = 0.0;
            for (l1 = ll1; l1 <= laC; l1 += 2)
              for (l2 = ll2; l2 <= lbC; l2 += 2) {
                double factor = 0.0;
                for (m = 0; m < 2 * l + 1; m++) {
                  factor += omegaA[l1 * incA2 + LMI(l, m) * incA1 + p] * omegaB[l2 * incB2 + LMI(l, m) * incB1 + q];
                  h->cnt.M_link += 1;
                }
                tmp += factor * T[l1 * Tinc2 + l2 * Tinc1 + alpha + beta];
              }
            gamma[p * inc + q] += tmp;
          }
        }
      }
    }
    free(T);
  }
  return gamma;
}

/* ------------------------------------------------------------------------------------------ */
/* type 1 */
typedef struct {
  double C, minExp;
  const double *expo, *K, *U, *rn;
  unsigned char *touched;
  double *pts;
} Q1Params;
static double integrand_Q1(int n, void *v) { /* src/type1.c:78-88 */
  Q1Params *p = v;
  double Q = 0.0, e = p->expo[n];
  if (p->touched) p->touched[n] = 1;
  if (e >= p->minExp) {
    if (p->pts) *p->pts += 1;
    Q = p->C * p->rn[n] * p->U[n] * p->K[n] * exp(e);
  }
  return Q;
}

/* radial type-1 integrals Q[N][lambda] for one primitive pair: src/type1.c:94-208. rc 0/1 */
static int type1_Q(OracleECP *h, double *T /* [(lab+1)^2], zeroed */, const double *U_L, const Pot *U, double s,
                   int lab, double dAC, double ca, double za, double dBC, double cb, double zb) {
  const int G = h->large->order, g = h->small1->order, Tinc = lab + 1;
  Grid *sgd = h->small1;
  Q1Params ps;
  double zd2 = -za * dAC * dAC - zb * dBC * dBC, z = -za - zb;
  double *expo = calloc(G, sizeof(double)), *Uv = calloc(G, sizeof(double));
  double *K = calloc((size_t)(lab + 1) * G, sizeof(double)), *rn = calloc((size_t)(lab + 1) * G, sizeof(double));
  unsigned char *touched = calloc(G, 1), *br = calloc(G, 1);
  int *failed = calloc(2 * (lab + 1) * (lab + 1), sizeof(int));
  int nFailed = 0, n, l1, l2, rc = 0;
  ps.C = ca * cb * exp(zd2);
  ps.expo = expo;
  ps.minExp = h->lnAccuracy1;
  ps.U = Uv;
  ps.touched = touched;
  ps.pts = &h->cnt.P_Q1;
  for (n = sgd->start; n < sgd->end; n++) {
    double r = sgd->x[n];
    expo[n] = (z * r + s) * r;
    Uv[n] = U_L[n];
    br[n] = bessel_eval(h, g, lab, s * r, K + n);
    rn[n] = 1.0;
    for (l1 = 1; l1 <= lab; l1++) rn[l1 * g + n] = r * rn[(l1 - 1) * g + n];
    h->cnt.tab1s_tabulated += 1;
  }
  for (l1 = 0; l1 <= lab; l1++) {
    ps.rn = rn + l1 * g;
    for (l2 = l1; l2 >= 0; l2 -= 2) {
      ps.K = K + l2 * g;
      h->cnt.ps93_calls += 1;
      if (0 != quad_ps93(integrand_Q1, &ps, sgd, NULL)) {
        h->cnt.ps93_fail += 1;
        failed[nFailed * 2] = l1;
        failed[nFailed * 2 + 1] = l2;
        nFailed++;
      } else
        T[l1 * Tinc + l2] += sgd->I;
    }
  }
  for (n = sgd->start; n < sgd->end; n++)
    if (touched[n]) {
      h->cnt.tab1s_touched += 1;
      h->cnt.flops_tab += 20 + wB(br[n], lab) + lab;
      h->cnt.flops_tab1 += 20 + wB(br[n], lab) + lab;
    }
  if (nFailed > 0) {
    Grid *grid = grid_copy(h->large);
    double zp = za + zb, p = (za * dAC + zb * dBC) / zp;
    ps.C = ca * cb;
    map_fm06(grid->order, grid->x, grid->w, zp, p);
    memset(touched, 0, G);
    for (n = 0; n < G; n++) {
      double r = grid->x[n];
      expo[n] = (z * r + s) * r + zd2;
      if (r > dAC && r > dBC && expo[n] < ps.minExp) {
        grid->end = n - 1;
        break;
      } else if (expo[n] >= ps.minExp) {
        br[n] = bessel_eval(h, G, lab, s * r, K + n);
        Uv[n] = pot_eval(U, U->L, r);
        rn[n] = 1.0;
        for (l1 = 1; l1 <= lab; l1++) rn[l1 * G + n] = r * rn[(l1 - 1) * G + n];
        h->cnt.tab1l_tabulated += 1;
      }
    }
    for (n = 0; n < nFailed; n++) {
      l1 = failed[n * 2];
      l2 = failed[n * 2 + 1];
      ps.rn = rn + l1 * G;
      ps.K = K + l2 * G;
      h->cnt.psm92_calls += 1;
      if (0 != quad_psm92(integrand_Q1, &ps, grid, NULL)) {
        h->cnt.psm92_fail += 1;
        rc = 1;
        break;
      }
      T[l1 * Tinc + l2] += grid->I;
    }
    for (n = 0; n < G; n++)
      if (touched[n] && expo[n] >= ps.minExp) {
        h->cnt.tab1l_touched += 1;
        h->cnt.flops_tab += 20 + wB(br[n], lab) + 24.0 * pot_count(U, U->L) + lab;
        h->cnt.flops_tab1 += 20 + wB(br[n], lab) + 24.0 * pot_count(U, U->L) + lab;
      }
    grid_free(grid);
  }
  free(expo);
  free(Uv);
  free(K);
  free(rn);
  free(touched);
  free(br);
  free(failed);
  return rc;
}

/* chi[C_DIM(la)][C_DIM(lb)]: src/type1.c:211-302 ; NULL on failure */
static double *type1_chi(OracleECP *h, const double *U_L, const Pot *U, const double *rAC, double dAC,
                         const double *rBC, double dBC, int la, int shella, int offa, int lb, int shellb, int offb) {
  const int Na = h->contraction[shella], Nb = h->contraction[shellb];
  const double *za = h->a + offa, *ca = h->d + offa, *zb = h->a + offb, *cb = h->d + offb;
  const int lab = la + lb, D = h->ijkDim, tmDim = LD(h->tmDim), inc = CD(lb), incQ = lab + 1;
  double *chi = calloc((size_t)CD(la) * CD(lb), sizeof(double));
  double *rsph = malloc(LD(lab) * sizeof(double));
  double *Q = malloc((lab + 1) * (lab + 1) * sizeof(double));
  const int *ix = h->ijkIndex;
  int pa, pb, p, ax, ay, az, bx, by, bz, l, m;
  for (pa = 0; pa < Na; pa++)
    for (pb = 0; pb < Nb; pb++) {
      double P[3], S[3];
      for (p = 0; p < 3; p++) P[p] = 2.0 * (za[pa] * rAC[p] + zb[pb] * rBC[p]);
      oracle_sphcoord(P, S);
      rsh_eval(lab, S[1], S[2], h->fac, h->dfac, rsph);
      memset(Q, 0, (lab + 1) * (lab + 1) * sizeof(double));
      h->cnt.t1_pairs += 1;
      if (0 != type1_Q(h, Q, U_L, U, S[0], lab, dAC, ca[pa], za[pa], dBC, cb[pb], zb[pb])) {
        free(chi);
        free(rsph);
        free(Q);
        return NULL;
      }
      for (ax = 0; ax <= la; ax++)
        for (ay = 0; ay <= la - ax; ay++)
          for (az = 0; az <= la - ax - ay; az++) {
            int i = ix[ax * D * D + ay * D + az];
            for (bx = 0; bx <= lb; bx++)
              for (by = 0; by <= lb - bx; by++)
                for (bz = 0; bz <= lb - bx - by; bz++) {
                  int j = ix[bx * D * D + by * D + bz];
                  int lx = ax + bx, ly = ay + by, lz = az + bz, lmax = lx + ly + lz;
                  p = ix[lx * D * D + ly * D + lz];
                  for (l = lmax; l >= 0; l -= 2) {
                    double factor = 0.0;
                    for (m = 0; m < 2 * l + 1; m++) {
                      factor += rsph[LMI(l, m)] * h->poly2sph[p * tmDim + LMI(l, m)];
                      h->cnt.M_chi += 1;
                    }
                    chi[i * inc + j] += factor * Q[lmax * incQ + l];
                  }
                }
          }
    }
  free(rsph);
  free(Q);
  return chi;
}

/* ------------------------------------------------------------------------------------------ */
/* driver loop nest: src/libecp.c:212-404 (n = 0) */
int oracle_calculateECPIntegrals(OracleECP *h, OracleCallback cb, void *args) {
  const int nrAtoms = h->nrAtoms, order = h->small1->order;
  const double norm1 = 4.0 * M_PI, norm2 = norm1 * norm1;
  double *U_L = calloc(order, sizeof(double));
  Window *sg = calloc(h->nrShells, sizeof(Window));
  int A, B, C, s1, s2, s, result = 0;
  memset(&h->cnt, 0, sizeof(Counters));
  for (C = 0; C < nrAtoms && !result; C++) {
    const double *rC;
    double *UTab, *FTab;
    int sOff1 = 0, pOff1 = 0, sIdx1 = 0, pIdx1 = 0, sOff2, pOff2, sIdx2 = 0, pIdx2 = 0;
    if (!h->U[C]) continue;
    rC = h->geometry + C * 3;
    h->small2->start = 0;
    h->small2->end = order - 1;
    UTab = tab_potential(h, h->U[C], h->small2, h->maxLambda, h->accuracy2, NULL);
    FTab = tab_F(h, rC, sg);
    for (s = 0; s < order; s++) U_L[s] = pot_eval(h->U[C], h->U[C]->L, h->small1->x[s]);
    h->cnt.flops_Ftab += 24.0 * pot_count(h->U[C], h->U[C]->L) * order;
    for (A = 0; A < nrAtoms && !result; A++) {
      const double *rA = h->geometry + A * 3;
      double dAC = dist(rC, rA), rAC[3] = {rA[0] - rC[0], rA[1] - rC[1], rA[2] - rC[2]};
      double *uspA = usp_table(h->maxLAtom[A], rAC);
      double *omegaA = angular_for_atom(h, h->maxLAtom[A], rAC);
      sOff2 = sOff1;
      pOff2 = pOff1;
      for (B = A; B < nrAtoms && !result; B++) {
        const double *rB = h->geometry + B * 3;
        double dBC = dist(rC, rB), rBC[3] = {rB[0] - rC[0], rB[1] - rC[1], rB[2] - rC[2]};
        double *uspB = usp_table(h->maxLAtom[B], rBC);
        double *omegaB = angular_for_atom(h, h->maxLAtom[B], rBC);
        sIdx1 = sOff1;
        pIdx1 = pOff1;
        for (s1 = 0; s1 < h->shells[A] && !result; s1++) {
          int la = h->am[sIdx1];
          sIdx2 = sOff2;
          pIdx2 = pOff2;
          for (s2 = 0; s2 < h->shells[B] && !result; s2++) {
            int lb = h->am[sIdx2], t;
            if (!((A == B && s2 < s1) || sg[sIdx1].skip || sg[sIdx2].skip)) {
              int gs = sg[sIdx1].start > sg[sIdx2].start ? sg[sIdx1].start : sg[sIdx2].start;
              int ge = sg[sIdx1].end > sg[sIdx2].end ? sg[sIdx1].end : sg[sIdx2].end;
              h->small1->start = h->small2->start = gs;
              h->small1->end = h->small2->end = ge;
              if (gs < ge) {
                h->cnt.triples_exec += 1;
                for (t = 1; t <= 2; t++) {
                  double *I, *G;
                  if (t == 1)
                    G = type1_chi(h, U_L, h->U[C], rAC, dAC, rBC, dBC, la, sIdx1, pIdx1, lb, sIdx2, pIdx2);
                  else
                    G = type2_gamma(h, FTab, UTab, h->U[C], dAC, dBC, la, sIdx1, pIdx1, omegaA, h->maxLAtom[A], lb,
                                    sIdx2, pIdx2, omegaB, h->maxLAtom[B]);
                  if (!G) {
                    result = t;
                    break;
                  }
                  I = shift_polynomials(h, G, t == 1 ? norm1 : norm2, la, uspA, h->maxLAtom[A] + 1, lb, uspB,
                                        h->maxLAtom[B] + 1);
                  free(G);
                  h->cnt.callbacks += 1;
                  if (cb) cb(A, s1, la, 0, B, s2, lb, 0, C, I, args);
                  free(I);
                }
              }
            }
            pIdx2 += h->contraction[sIdx2];
            sIdx2++;
          }
          pIdx1 += h->contraction[sIdx1];
          sIdx1++;
        }
        sOff2 = sIdx2;
        pOff2 = pIdx2;
        free(omegaB);
        free(uspB);
      }
      sOff1 = sIdx1;
      pOff1 = pIdx1;
      free(omegaA);
      free(uspA);
    }
    free(UTab);
    free(FTab);
  }
  free(U_L);
  free(sg);
  return result;
}

/* one-call interface: src/getIntegrals.c:22-95 */
typedef struct {
  double *I;
  int dimI, *aoDim, dim;
} ScatterArgs;
static void scatter_cb(int A, int sa, int la, int shifta, int B, int sb, int lb, int shiftb, int C, double *I,
                       void *p) {
  ScatterArgs *P = p;
  const int a0 = P->aoDim[A * P->dim + sa], b0 = P->aoDim[B * P->dim + sb];
  int i, j;
  (void)shifta;
  (void)shiftb;
  (void)C;
  for (i = 0; i < IJK(la); i++)
    for (j = 0; j < IJK(lb); j++) {
      if (a0 + i > b0 + j) continue;
      P->I[(a0 + i) * P->dimI + (b0 + j)] += I[i * IJK(lb) + j];
    }
}
int oracle_getIntegrals(int nrAtoms, double *geometry, int *shellsECP, int *KECP, int *lECP, double *nECP,
                        double *dECP, double *aECP, int *shellsBS, int *lBS, int *KBS, double *dBS, double *aBS,
                        int largeGridOrder, double tolerance, double accuracy, int rowdim, double *I) {
  ScatterArgs P;
  OracleECP *h;
  int i, j, maxShells = 0, lstart = 0, idx = 0;
  for (i = 0; i < nrAtoms; i++)
    if (shellsBS[i] > maxShells) maxShells = shellsBS[i];
  P.aoDim = calloc(nrAtoms * maxShells + 1, sizeof(int));
  for (i = 0; i < nrAtoms; i++)
    for (j = 0; j < shellsBS[i]; j++) {
      P.aoDim[i * maxShells + j] = lstart;
      lstart += IJK(lBS[idx]);
      idx++;
    }
  P.I = I;
  P.dimI = rowdim;
  P.dim = maxShells;
  h = oracle_libECP_init(nrAtoms, geometry, shellsECP, lECP, KECP, nECP, dECP, aECP, shellsBS, lBS, KBS, dBS, aBS, 0,
                         -1, NULL, largeGridOrder, tolerance, accuracy);
  if (!h) {
    printf("error initializing libECP\n");
    free(P.aoDim);
    return 1;
  }
  oracle_calculateECPIntegrals(h, scatter_cb, &P);
  oracle_libECP_free(h);
  free(P.aoDim);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* accessors for unit-level parity tests */
const double *oracle_table(OracleECP *h, const char *name, int *len) {
#define RET(p, n)  \
  do {             \
    *len = (n);    \
    return (p);    \
  } while (0)
  if (!strcmp(name, "fac")) RET(h->fac, h->nfac);
  if (!strcmp(name, "dfac")) RET(h->dfac, h->nfac);
  if (!strcmp(name, "cart2sph")) RET(h->cart2sph, h->ncart2sph);
  if (!strcmp(name, "poly2sph")) RET(h->poly2sph, h->npoly2sph);
  if (!strcmp(name, "omega")) RET(h->omega, h->nomega);
  if (!strcmp(name, "small_x")) RET(h->small1->x, h->small1->order);
  if (!strcmp(name, "small_w")) RET(h->small1->w, h->small1->order);
  if (!strcmp(name, "large_x")) RET(h->large->x, h->large->order);
  if (!strcmp(name, "large_w")) RET(h->large->w, h->large->order);
  if (!strcmp(name, "bessel")) RET(h->bK, (h->bLMax + 1) * h->bDim);
  if (!strcmp(name, "besselC")) RET(h->bC, h->bLMax + 1);
#undef RET
  *len = 0;
  return NULL;
}
const int *oracle_itable(OracleECP *h, const char *name, int *len) {
  if (!strcmp(name, "ijk")) {
    *len = 3 * CD(h->tmDim);
    return h->ijk;
  }
  if (!strcmp(name, "ijkIndex")) {
    *len = h->ijkDim * h->ijkDim * h->ijkDim;
    return h->ijkIndex;
  }
  if (!strcmp(name, "dims")) {
    *len = 8;
    return h->dims;
  }
  *len = 0;
  return NULL;
}
void oracle_bessel(OracleECP *h, int lmax, double z, double *K) { bessel_eval(h, 1, lmax, z, K); }
void oracle_rsh(OracleECP *h, int lmax, double theta, double phi, double *out) {
  rsh_eval(lmax, theta, phi, h->fac, h->dfac, out);
}

typedef struct {
  const double *f;
} TabParams;
static double integrand_tab(int n, void *v) { return ((TabParams *)v)->f[n]; }
int oracle_ps93_table(OracleECP *h, const double *f, int start, int end, double *result, int *npoints) {
  TabParams ps = {f};
  Grid g = *h->small1;
  int rc;
  g.start = start;
  g.end = end;
  rc = quad_ps93(integrand_tab, &ps, &g, npoints);
  *result = g.I;
  return rc;
}
int oracle_psm92_table(OracleECP *h, const double *w, const double *f, int start, int end, double *result,
                       int *npoints) {
  TabParams ps = {f};
  Grid g = *h->large;
  int rc;
  if (w) g.w = (double *)w;
  g.start = start;
  g.end = end;
  rc = quad_psm92(integrand_tab, &ps, &g, npoints);
  *result = g.I;
  return rc;
}
void oracle_screening(OracleECP *h, int C, int *end_l, int *sstart, int *send, int *sskip) {
  Window *sg = calloc(h->nrShells, sizeof(Window));
  double *U, *F;
  int i;
  if (!h->U[C]) {
    free(sg);
    return;
  }
  h->small2->start = 0;
  h->small2->end = h->small2->order - 1;
  U = tab_potential(h, h->U[C], h->small2, h->maxLambda, h->accuracy2, end_l);
  F = tab_F(h, h->geometry + 3 * C, sg);
  for (i = 0; i < h->nrShells; i++) {
    sstart[i] = sg[i].start;
    send[i] = sg[i].end;
    sskip[i] = sg[i].skip;
  }
  free(U);
  free(F);
  free(sg);
}
void oracle_set_stale_buffers(OracleECP *h, int on) { h->stale = on; }
void oracle_counters(OracleECP *h, double *out, int n) {
  int i;
  const double *c = (const double *)&h->cnt;
  for (i = 0; i < n && i < NCOUNTERS; i++) out[i] = c[i];
}

/* callbacks for timing runs (bench.py --impl reference / cpu_baseline): no Python in the loop.
 * args == NULL: ignore the block; else accumulate a checksum into *(double*)args. */
void oracle_sum_callback(int A, int s1, int la, int shifta, int B, int s2, int lb, int shiftb, int C, double *I,
                         void *args) {
  (void)A; (void)s1; (void)B; (void)s2; (void)C;
  if (args) {
    const int n = IJK(la + shifta) * IJK(lb + shiftb);
    double s = 0.0;
    int i;
    for (i = 0; i < n; i++) s += I[i];
    *(double *)args += s;
  }
}
