/* tables.c - geometry-independent host tables of a handle.
 *
 * Replaces the table set-up of libECP_init / Type1_init / Type2_init (reference src/libecp.c:143-198)
 * and tabOmega (src/angular_integrals.c:15-101).  Everything here runs once per handle, in double
 * precision with glibc libm and the reference's operation order, so the uploaded tables are bitwise the
 * reference's.  What is new is the *layout*: level-major padded quadrature grids, a transposed Bessel
 * table, one potential table per distinct ECP parameter set, and per-class lists of the radial
 * quadratures whose angular factor is not identically zero.
 */
#include "tables.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int LD(int l) { return (l + 1) * (l + 1); }
static int CD(int l) { return (l + 1) * (l + 2) * (l + 3) / 6; }
static int IJK(int l) { return (l + 1) * (l + 2) / 2; }
static int CIDX(int l, int c) { return CD(l - 1) + c; }
static int LMI(int l, int m) { return l * l + m; }

/* n! and n!! tables, 0..n (reference src/util.c:13-56) */
static double *factorial_table(int n, int step) {
  double *f = calloc(n + 1, sizeof(double));
  if (n > 0) f[0] = 1.0;
  if (n > 1) f[1] = 1.0;
  for (int i = 2; i <= n; i++) f[i] = f[i - step] * i;
  return f;
}
static double nk(int n, int k, const double *fac) { /* src/util.c:61-70 */
  return (k >= 0 && k <= n) ? fac[n] / (fac[n - k] * fac[k]) : 0.0;
}

/* libint Cartesian order and its inverse (reference src/dimensions.c:17-57) */
static void cartesian_order(int am, int **ijk_, int **inv_) {
  const int dim = am + 1;
  int *ijk = calloc(3 * CD(am), sizeof(int)), *inv = calloc(dim * dim * dim, sizeof(int));
  for (int l = 0; l <= am; l++) {
    int c = 0;
    for (int i = 0; i <= l; i++)
      for (int j = 0; j <= i; j++, c++) {
        int *e = ijk + 3 * CIDX(l, c);
        e[0] = l - i;
        e[1] = i - j;
        e[2] = j;
        inv[e[0] * dim * dim + e[1] * dim + e[2]] = CIDX(l, c);
      }
  }
  *ijk_ = ijk;
  *inv_ = inv;
}

/* a caller-supplied order (layout of cartesianShellOrder(lmaxOrd)): the shells l <= am are copied, the inverse cube has
 * the caller's edge lmaxOrd + 1 (reference src/libecp.c:163-165, src/dimensions.c:42-57).  Returns 1 if a shell's
 * components are not a permutation of its monomials. */
static int custom_order(const int *ord, int lmaxOrd, int am, int **ijk_, int **inv_) {
  const int dim = lmaxOrd + 1;
  int *ijk = calloc(3 * CD(am), sizeof(int)), *inv = calloc((size_t)dim * dim * dim, sizeof(int));
  int bad = 0;
  for (int i = 0; i < dim * dim * dim; i++) inv[i] = -1;
  for (int l = 0; l <= am && !bad; l++)
    for (int c = 0; c < IJK(l); c++) {
      const int *s = ord + 3 * CIDX(l, c);
      int *e = ijk + 3 * CIDX(l, c);
      if (s[0] < 0 || s[1] < 0 || s[2] < 0 || s[0] + s[1] + s[2] != l || inv[s[0] * dim * dim + s[1] * dim + s[2]] >= 0) {
        bad = 1;
        break;
      }
      e[0] = s[0]; e[1] = s[1]; e[2] = s[2];
      inv[e[0] * dim * dim + e[1] * dim + e[2]] = CIDX(l, c);
    }
  for (int i = 0; i < dim * dim * dim; i++)
    if (inv[i] < 0) inv[i] = 0; /* the reference's calloc'ed cube */
  if (bad) {
    free(ijk);
    free(inv);
    return 1;
  }
  *ijk_ = ijk;
  *inv_ = inv;
  return 0;
}

/* packed (l, m, c) offsets of the Cartesian->spherical matrix (reference src/transformations.h:13-14) */
/* ---- process-wide cache of the tables that do not depend on the molecule ----
 * cart2sph / poly2sph (key tmDim), Omega (key maxLECP, maxLambda, maxAlpha, tmDim) and the Bessel table (key lMax,
 * accuracy) are pure functions of a few integers.  A caller of getIntegrals creates a handle per call (reference
 * src/getIntegrals.c:78-89): the first handle computes them as before, later ones copy the stored doubles
 * (bit-identical by construction; 2.6 of 4.6 ms of the host part of libECP_init on the 500-atom config). */
typedef struct {
  int used, key[4];
  double acc;
  size_t n;
  double *data;
} ConstEntry;
enum { CC_C2S, CC_P2S, CC_OMEGA, CC_BESSELK, CC_KINDS };
#define CC_SLOTS 6
static ConstEntry g_cc[CC_KINDS][CC_SLOTS];
static pthread_mutex_t g_ccMu = PTHREAD_MUTEX_INITIALIZER;
static double *cc_get(int kind, int k0, int k1, int k2, int k3, double acc, size_t *n) {
  double *out = NULL;
  pthread_mutex_lock(&g_ccMu);
  for (int i = 0; i < CC_SLOTS; i++) {
    const ConstEntry *e = &g_cc[kind][i];
    if (e->used && e->key[0] == k0 && e->key[1] == k1 && e->key[2] == k2 && e->key[3] == k3 && e->acc == acc) {
      out = malloc((e->n ? e->n : 1) * sizeof(double));
      memcpy(out, e->data, e->n * sizeof(double));
      if (n) *n = e->n;
      break;
    }
  }
  pthread_mutex_unlock(&g_ccMu);
  return out;
}
static void cc_put(int kind, int k0, int k1, int k2, int k3, double acc, const double *data, size_t n) {
  pthread_mutex_lock(&g_ccMu);
  for (int i = 0; i < CC_SLOTS; i++) {
    ConstEntry *e = &g_cc[kind][i];
    if (e->used) continue; /* a full cache simply stops learning */
    e->data = malloc((n ? n : 1) * sizeof(double));
    memcpy(e->data, data, n * sizeof(double));
    e->n = n;
    e->key[0] = k0; e->key[1] = k1; e->key[2] = k2; e->key[3] = k3;
    e->acc = acc;
    e->used = 1;
    break;
  }
  pthread_mutex_unlock(&g_ccMu);
}

static int c2s_size(int l, const double *fac) { return (int)((3 * l + 2) * fac[l + 3] / (12 * fac[l])); }
static int c2s_index(int l, int m, int c, const double *fac) { return l == 0 ? 0 : c2s_size(l - 1, fac) + m * IJK(l) + c; }

/* Schlegel-Frisch coefficients <S_lm | x^lx y^ly z^lz> (reference src/transformations.c:28-87) */
static double *build_cart2sph(int lmax, const int *xyz, const double *fac) {
  double *out = calloc(c2s_size(lmax, fac), sizeof(double)), *T = out;
  for (int l = 0; l <= lmax; l++)
    for (int m = -l; m <= l; m++) {
      const int mm = abs(m);
      for (int c = 0; c < IJK(l); c++, T++) {
        const int *e = xyz + 3 * CIDX(l, c);
        const int lx = e[0], ly = e[1], lz = e[2];
        int j = lx + ly - mm;
        if (j < 0 || j % 2 == 1) continue; /* stays 0.0 */
        j /= 2;
        double s1 = 0.0;
        for (int i = 0; i <= (l - mm) / 2; i++) {
          double s2 = 0.0;
          for (int k = 0; k <= j; k++) {
            double s = 0.0;
            if ((m < 0 && abs(mm - lx) % 2 == 1) || (m > 0 && abs(mm - lx) % 2 == 0))
              s = pow(-1.0, (mm - lx + 2 * k) / 2) * sqrt(2.0);
            else if (m == 0 && lx % 2 == 0)
              s = pow(-1.0, -lx / 2 + k);
            s2 += nk(j, k, fac) * nk(mm, lx - 2 * k, fac) * s;
          }
          s1 += nk(l, i, fac) * nk(i, j, fac) * pow(-1.0, i) * fac[2 * l - 2 * i] / fac[l - mm - 2 * i] * s2;
        }
        *T = sqrt((fac[2 * lx] * fac[2 * ly] * fac[2 * lz] * fac[l] * fac[l - mm]) /
                  (fac[2 * l] * fac[lx] * fac[ly] * fac[lz] * fac[l + mm])) *
             1 / (pow(2.0, l) * fac[l]) * s1;
      }
    }
  return out;
}

/* monomial -> real spherical harmonics on the unit sphere, FM06 eq. 36 (reference src/transformations.c:146-207) */
static double *build_poly2sph(const double *c2s, int lmax, const int *xyz, const double *dfac) {
  const int ldim = LD(lmax);
  double *out = calloc((size_t)CD(lmax) * ldim, sizeof(double));
  for (int l1 = 0; l1 <= lmax; l1++)
    for (int c1 = 0; c1 < IJK(l1); c1++) {
      const int *e1 = xyz + 3 * CIDX(l1, c1);
      double *row = out + (size_t)CIDX(l1, c1) * ldim;
      const double *src = c2s;
      for (int l2 = 0; l2 <= l1; l2++) {
        const double s1 = 4.0 * M_PI * dfac[2 * l2 + 1];
        for (int m = 0; m < 2 * l2 + 1; m++) {
          double sum = 0.0;
          for (int c2 = 0; c2 < IJK(l2); c2++, src++) {
            const int *e2 = xyz + 3 * CIDX(l2, c2);
            const int lx = e1[0] + e2[0], ly = e1[1] + e2[1], lz = e1[2] + e2[2];
            if (lx % 2 || ly % 2 || lz % 2) continue;
            double s = s1, s2 = 1.0 / dfac[l1 + l2 + 1];
            if (lx > 2) s2 *= dfac[lx - 1];
            if (ly > 2) s2 *= dfac[ly - 1];
            if (lz > 2) s2 *= dfac[lz - 1];
            if (e2[0] > 1) s /= dfac[2 * e2[0] - 1];
            if (e2[1] > 1) s /= dfac[2 * e2[1] - 1];
            if (e2[2] > 1) s /= dfac[2 * e2[2] - 1];
            sum += sqrt(s) * s2 * (*src);
          }
          row[LMI(l2, m)] = sum;
        }
      }
    }
  return out;
}

/* Omega[(l,m)][(lambda,mu)][C_INDEX(alpha,c)] (reference src/angular_integrals.c:15-101) */
static double *build_omega(const EcpTables *t, int *len) {
  const EcpHostTables *v = &t->v;
  const int d2 = LD(v->maxLambda), d3 = CD(v->maxAlpha), D = v->ijkDim, pcols = LD(v->tmDim);
  double *om = calloc((size_t)LD(v->maxLECP) * d2 * d3, sizeof(double));
  for (int lam = 0; lam <= v->maxLambda; lam++)
    for (int l = 0; l < v->maxLECP; l++) {
      const int par = (lam + l) % 2, dl = lam - l;
      for (int alpha = (par > dl ? par : dl); alpha <= v->maxAlpha; alpha += 2)
        for (int mu = 0; mu < 2 * lam + 1; mu++)
          for (int m = 0; m < 2 * l + 1; m++) {
            double *dst = om + ((size_t)LMI(l, m) * d2 + LMI(lam, mu)) * d3 + CD(alpha - 1);
            for (int c = 0; c < IJK(alpha); c++) {
              const int *ec = t->ijk + 3 * CIDX(alpha, c);
              if (alpha == 0) {
                if (l == lam && m == mu) dst[c] = 1.0;
                continue;
              }
              if (lam > l + alpha) continue;
              for (int d = 0; d < IJK(l); d++) {
                const int *ed = t->ijk + 3 * CIDX(l, d);
                double N = 0.25 * t->dfac[2 * l + 1] / M_PI;
                if (ed[0] > 1) N /= t->dfac[2 * ed[0] - 1];
                if (ed[1] > 1) N /= t->dfac[2 * ed[1] - 1];
                if (ed[2] > 1) N /= t->dfac[2 * ed[2] - 1];
                N = sqrt(N);
                const int mono = t->ijkIndex[(ed[0] + ec[0]) * D * D + (ed[1] + ec[1]) * D + (ed[2] + ec[2])];
                dst[c] += N * t->cart2sph[c2s_index(l, m, d, t->fac)] * t->poly2sph[(size_t)mono * pcols + LMI(lam, mu)];
              }
            }
          }
    }
  *len = LD(v->maxLECP) * d2 * d3;
  return om;
}

/* ---------------------------------------------------------------------------------------------- */
/* PS93 grid of order 128 -> 383 points, KK-mapped to (0,inf)
 * (reference src/gc_integrators.c:220-283, 301-313) */
static void build_small_grid(EcpTables *t) {
  const int runs = (int)floor(log(128) / log(2));
  int offset = (int)pow(2, runs), n = 3;
  const int order = 3 * offset - 1;
  double *x = calloc(order, sizeof(double)), *w = calloc(order, sizeof(double));
  double C0 = sin(M_PI / 3), S0 = 0.5, C1 = S0, S1 = C0, c = cos(M_PI / 3), s = C0, s2 = s * s, tt;
  x[order / 2] = 0.0;
  w[order / 2] = 1.0;
  tt = (n - 2.0) / n + 2 / M_PI * (1 + 2 * s2 / 3) * c * s;
  x[offset - 1] = -tt;
  x[order - offset] = tt;
  w[order - offset] = w[offset - 1] = s2 * s2;
  while ((4 * n / 3 - 1) <= order) {
    c = C0;
    s = S0;
    offset /= 2;
    for (int i = 1; i < n; i += 2) {
      s2 = s * s;
      const int idx = i * offset - 1;
      tt = 1 + 2 / (3 * M_PI) * s * c * (3 + 2 * s2) - ((double)i) / n;
      x[idx] = -tt;
      x[order - idx - 1] = tt;
      w[order - idx - 1] = w[idx] = s2 * s2;
      tt = s;
      s = s * C1 + c * S1;
      c = c * C1 - tt * S1;
    }
    n *= 2;
    C1 = C0;
    S1 = S0;
    C0 = sqrt((1 + C0) / 2);
    S0 = S0 / (2 * C0);
  }
  const double ln2 = log(2.0);
  for (int i = 0; i < order; i++) {
    const double xi = 1.0 - log(1.0 - x[i]) / ln2, wi = w[i] / (ln2 * (1.0 - x[i]));
    x[i] = xi;
    w[i] = wi;
  }
  t->small_x = x;
  for (int k = 0, j = 0; k <= ECP_WIN_LUT_BINS; k++) { /* abscissae ascend strictly */
    while (j < order && x[j] < k / ECP_WIN_LUT_SCALE) j++;
    t->winLut[k] = j;
  }
  t->small_w = w;

  /* level-major padded slot layout: replay the visiting order of integrateGC_PS93
   * (reference src/gc_integrators.c:156-217) */
  int16_t *oidx = calloc(ECP_SMALL_SLOTS, sizeof(int16_t));
  EcpHostTables *v = &t->v;
  offset = (int)pow(2, runs);
  oidx[0] = (order - 1) / 2;
  oidx[1] = -1;
  oidx[2] = offset - 1;
  oidx[3] = order - offset;
  offset /= 2;
  n = 3;
  int j = 0, lev = 0, slot = 4;
  while ((2 * n * (1 - j) + j * 4 * n / 3 - 1) <= order) {
    j = 1 - j;
    if (0 == j) offset /= 2;
    v->small_levSlot[lev] = slot;
    for (int i = 1; i < n; i += 2)
      if (3 * ((i + 2 * j) / 3) >= i + j) {
        const int idx = i * offset - 1;
        oidx[slot++] = idx;
        oidx[slot++] = order - idx - 1;
      }
    n *= (1 + j);
    v->small_levJ[lev] = j;
    v->small_levN[lev] = n;
    v->small_levPairs[lev] = (slot - v->small_levSlot[lev]) / 2;
    lev++;
    if (lev > ECP_SMALL_LEVELS) abort();
  }
  v->small_levSlot[lev] = slot;
  if (lev != ECP_SMALL_LEVELS || slot != ECP_SMALL_SLOTS || order != ECP_SMALL_ORDER) abort();
  t->small_oidx = oidx;
  t->small_rs = calloc(ECP_SMALL_SLOTS, sizeof(double));
  t->small_ws = calloc(ECP_SMALL_SLOTS, sizeof(double));
  for (int k = 0; k < ECP_SMALL_SLOTS; k++)
    if (oidx[k] >= 0) {
      t->small_rs[k] = x[oidx[k]];
      t->small_ws[k] = w[oidx[k]];
    }
}

/* PSM92 template grid on (-1,1) (reference src/gc_integrators.c:89-145) + level-major slot layout
 * following integrateGC_PSM92's visiting order (src/gc_integrators.c:38-86) */
static void build_large_grid(EcpTables *t, int maxPoints) {
  const int order = pow(2, floor(log(maxPoints + 1) / log(2))) - 1;
  const int runs = (int)floor(log(order) / log(2)), M = (order - 1) / 2;
  int offset = (int)pow(2, runs), n = 1;
  double N = n + 1.0, S0 = 1.0, C0 = 0.0, S1, C1, s, c, tt;
  double *x = calloc(order, sizeof(double)), *w = calloc(order, sizeof(double));
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
    for (int i = 1; i <= n; i += 2) {
      tt = 1 + 2 / (3 * M_PI) * (3 + 2 * s * s) * s * c - i / N;
      const int idx = i * offset - 1;
      x[order - idx - 1] = tt;
      x[idx] = -tt;
      w[order - idx - 1] = w[idx] = s * s * s * s;
      tt = s;
      s = s * C1 + c * S1;
      c = c * C1 - tt * S1;
    }
    n = 2 * n + 1;
    N = n + 1.0;
  }
  t->large_x = x;
  t->large_w = w;
  /* the integrator recomputes its own order from t->n == order (src/gc_integrators.c:41) */
  const int iorder = pow(2, floor(log(order + 1) / log(2))) - 1;
  const int iruns = (int)floor(log(iorder) / log(2)), iM = (iorder - 1) / 2;
  const int slots = iorder + 1;
  int16_t *oidx = calloc(slots, sizeof(int16_t));
  int slot = 2, levels = 0;
  offset = (int)pow(2, iruns);
  oidx[0] = iM;
  oidx[1] = -1;
  n = 1;
  while (n <= iM) {
    offset /= 2;
    if (slot != (1 << (levels + 1))) abort();
    for (int i = 1; i <= n; i += 2) {
      const int idx = i * offset - 1;
      oidx[slot++] = idx;
      oidx[slot++] = iorder - idx - 1;
    }
    n = 2 * n + 1;
    levels++;
  }
  if (slot != slots) abort();
  t->large_oidx = oidx;
  t->large_xs = calloc(slots, sizeof(double));
  t->large_ws = calloc(slots, sizeof(double));
  for (int k = 0; k < slots; k++)
    if (oidx[k] >= 0) {
      t->large_xs[k] = x[oidx[k]];
      t->large_ws[k] = w[oidx[k]];
    }
  t->v.largeOrder = iorder;
  t->v.largeSlots = slots;
  t->v.largeLevels = levels;
}

/* K_l(z_i), z_i = i/100, i <= 1600, by power series (reference src/bessel.c:19-82); also the
 * transposed copy [i][stride] the kernels read */
static int build_bessel(EcpTables *t, int lMax, double accuracy) {
  const int N = 16 * 100, cutoff = 200, dim = N + 1;
  double *F = malloc((cutoff + 1) * sizeof(double)), *G = malloc((cutoff + lMax + 2) * sizeof(double));
  int rc = 0;
  double *K = cc_get(CC_BESSELK, lMax, 0, 0, 0, accuracy, NULL);
  const int cached = K != NULL;
  if (!cached) {
    K = calloc((size_t)(lMax + 1) * dim, sizeof(double));
    K[0] = 1.0;
  }
  for (int i = 1; i <= N && !rc && !cached; i++) {
    const double z = i / (N / 16.0);
    int j = 0;
    double f = z * z / 2.0, s;
    F[0] = exp(-z);
    G[0] = 1.0;
    s = F[0] / G[0];
    const int jmin = (int)(0.25 * sqrt(1.0 + 16.0 * f));
    while (s > accuracy || j <= jmin) {
      K[i] += s;
      if (++j > cutoff) {
        rc = 1;
        break;
      }
      F[j] = F[j - 1] * f / j;
      G[j] = G[j - 1] * (2.0 * j + 1.0);
      s = F[j] / G[j];
    }
    if (rc) break;
    for (int l = 1; l <= lMax; l++) G[j + l] = G[j + l - 1] * (2 * j + 2 * l + 1);
    f = z;
    for (int l = 1; l <= lMax; l++) {
      s = 0;
      for (int m = 0; m < j; m++) s += F[m] / G[l + m];
      K[l * dim + i] = f * s;
      f *= z;
    }
  }
  free(F);
  free(G);
  t->besselK = K;
  if (rc) return rc;
  if (!cached) cc_put(CC_BESSELK, lMax, 0, 0, 0, accuracy, K, (size_t)(lMax + 1) * dim);
  t->besselC = calloc(lMax + 1, sizeof(double));
  for (int i = 1; i <= lMax; i++) t->besselC[i] = i / (2.0 * i + 1.0);
  const int stride = (lMax + 1 + 3) & ~3; /* rows padded to 32 bytes */
  t->besselT = calloc((size_t)dim * stride, sizeof(double));
  for (int i = 0; i < dim; i++)
    for (int l = 0; l <= lMax; l++) t->besselT[(size_t)i * stride + l] = K[l * dim + i];
  t->v.besselLMax = lMax;
  t->v.besselStride = stride;
  return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* U_l(r) (reference src/ecp.c:41-60) */
double ecp_host_pot_eval(const EcpTables *t, int type, int l, double r) {
  const EcpType *T = &t->types[type];
  double v = 0.0;
  const double r2 = r * r;
  for (int i = T->gaussOff; i < T->gaussOff + T->N; i++)
    if (t->gaussL[i] == l) v += pow(r, t->gaussN[i]) * t->gaussD[i] * exp(-t->gaussA[i] * r2);
  return v;
}

/* shell radius: most diffuse primitive, Newton iteration on c r^l exp(-zeta r^2) = cutoff
 * (reference src/util.c:133-192; decides integer windows, so copied operation for operation) */
static double primitive_radius2(double c, double zeta, int l, double cutoff) {
  const double tl = log(fabs(c) / fabs(cutoff));
  double dg = tl / zeta, r = (dg > cutoff) ? dg : cutoff;
  if (l == 0) return r;
  if (l > 0) {
    dg = sqrt(0.5 * l / fabs(c));
    const double guess = (dg > cutoff) ? dg : cutoff;
    if (guess > r) r = 0.5 * (r + guess);
  }
  for (int i = 0; i < 40; i++) {
    const double zr = zeta * r;
    const double g = tl + l * log(r) - zr * r;
    const double delta = g / (l / r - 2 * zr);
    dg = r - delta;
    r = (dg > cutoff) ? dg : cutoff;
    if (fabs(delta) < cutoff) return r * r;
  }
  abort(); /* reference asserts (src/util.c:169) */
}
static double shell_radius(int depth, int am, const double *d, const double *a, double zero) {
  double zeta = a[0], c = fabs(d[0]);
  for (int i = 1; i < depth; i++)
    if (a[i] < zeta && d[i] != 0.0) {
      zeta = a[i];
      c = fabs(d[i]);
    }
  return sqrt(primitive_radius2(c, zeta, am, zero));
}

/* does the link step ever multiply T_l[l1][l2][l3] by an angular factor that is not identically zero?
 * (loop limits reference src/type2.c:590-610; structural zeros of Omega src/angular_integrals.c:40-44,63) */
int ecp_t2_used(int la, int lb, int l, int l1, int l2, int l3) {
  for (int alpha = 0; alpha <= la; alpha++) {
    const int beta = l3 - alpha;
    if (beta < 0 || beta > lb) continue;
    if ((alpha + l + l1) % 2 || (beta + l + l2) % 2) continue;
    if (l1 < l - alpha || l2 < l - beta || l1 > l + alpha || l2 > l + beta) continue;
    return 1;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* order of the shells for the row deal of a sharded run (see rowDeal) */
static const int *g_dealL, *g_dealK, *g_dealP;
static const double *g_dealA;
static uint64_t deal_mix(uint64_t x) {
  x = (x + 0x7F4A7C15ull) * 0x9E3779B97F4A7C15ull;
  x ^= x >> 29;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 32;
  return x;
}
static int deal_kind_cmp(int a, int b) {
  if (g_dealL[a] != g_dealL[b]) return g_dealL[a] < g_dealL[b] ? -1 : 1;
  if (g_dealK[a] != g_dealK[b]) return g_dealK[a] < g_dealK[b] ? -1 : 1;
  for (int k = 0; k < g_dealK[a]; k++) {
    const double x = g_dealA[g_dealP[a] + k], y = g_dealA[g_dealP[b] + k];
    if (x != y) return x > y ? -1 : 1;
  }
  return 0;
}
static int deal_cmp(const void *pa, const void *pb) {
  const int a = *(const int *)pa, b = *(const int *)pb;
  const int c = deal_kind_cmp(a, b);
  if (c) return c;
  const uint64_t ha = deal_mix((uint64_t)a), hb = deal_mix((uint64_t)b);
  if (ha != hb) return ha < hb ? -1 : 1;
  return a < b ? -1 : (a > b);
}

/* Deal of the rows of a sharded run (builder.c: ecp_pair_owner = rowDeal % world).  Shells are grouped by kind -
     * (l, contraction depth, exponents), i.e. "the same shell on another atom": the kind sets the cost of a triple.
     * The shells of a kind are dealt out one by one in a fixed pseudo-random order (atom indices of a lattice are
     * periodic in space: dealing 8 ranks down a column of 8 sites hands one rank the whole surface layer - measured;
     * pairing early with late rows, which have many / few partners b >= a, was measured too and is worse because it
     * halves the number of independent units).  Every rank receives the same number (+-1) of rows of every kind.
     * Executed triples per rank on the 500-atom config, max / mean: 1.001, 1.009, 1.040 at 2, 4, 8 ranks
     * (hash of the row index alone: 1.001, 1.034, -). */
const int *ecp_tables_row_deal(EcpTables *t) {
  static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_mutex_lock(&mu); /* the comparator reads file-scope pointers */
  if (!t->rowDeal) {
    const int nsh = t->v.nrShells;
    int *deal = malloc((nsh + 1) * sizeof(int));
    int *ord = malloc((nsh + 1) * sizeof(int));
    for (int s = 0; s < nsh; s++) ord[s] = s;
    g_dealL = t->dealL; g_dealK = t->dealK; g_dealA = t->dealA; g_dealP = t->shellPrim;
    qsort(ord, nsh, sizeof(int), deal_cmp);
    for (int k = 0; k < nsh; k++) deal[ord[k]] = k;
    free(ord);
    t->rowDeal = deal;
  }
  pthread_mutex_unlock(&mu);
  return t->rowDeal;
}

static _Thread_local char g_taberr[200] = "";
const char *ecp_tables_last_error(void) { return g_taberr; }
#define TAB_FAIL(...)                                     \
  do {                                                    \
    snprintf(g_taberr, sizeof(g_taberr), __VA_ARGS__);    \
    ecp_tables_free(t);                                   \
    return NULL;                                          \
  } while (0)

EcpTables *ecp_tables_build(int nrAtoms, const double *geometry, const int *shellsECP, const int *lECP,
                            const int *KECP, const double *nECP, const double *dECP, const double *aECP,
                            const int *shellsBS, const int *lBS, const int *KBS, const double *dBS, const double *aBS,
                            int largeGridOrder, double tolerance, double accuracy, const EcpBuildOpts *opts) {
  EcpTables *t = calloc(1, sizeof(EcpTables));
  const int *customOrder = opts ? opts->shellOrdering : NULL;
  EcpHostTables *v = &t->v;
  (void)geometry;
  v->nrAtoms = nrAtoms;
  v->tolerance = tolerance;
  v->accuracy = accuracy;
  v->lnAccuracy1 = log(accuracy) - 2;  /* src/libecp.c:90 */
  v->lnAccuracy2 = log(1.0E-14) - 2;   /* src/type2.c:58-59: hard-coded, ignores the argument */

  /* ---- basis set bookkeeping ---- */
  int nsh = 0, nprim = 0, nao = 0;
  for (int i = 0; i < nrAtoms; i++) nsh += shellsBS[i];
  t->shellL = malloc((nsh + 1) * sizeof(int));
  t->shellK = malloc((nsh + 1) * sizeof(int));
  t->shellPrim = malloc((nsh + 1) * sizeof(int));
  t->shellAtom = malloc((nsh + 1) * sizeof(int));
  t->shellAO = malloc((nsh + 1) * sizeof(int));
  t->atomMaxL = calloc(nrAtoms + 1, sizeof(int));
  t->atomFirstShell = calloc(nrAtoms + 1, sizeof(int));
  int present[ECP_MAX_LBS + 1] = {0};
  for (int i = 0, s = 0; i < nrAtoms; i++) {
    t->atomFirstShell[i] = s;
    for (int j = 0; j < shellsBS[i]; j++, s++) {
      const int l = lBS[s];
      if (l < 0 || l > ECP_MAX_LBS || KBS[s] < 1)
        TAB_FAIL("basis shell %d: angular momentum %d outside 0..%d (s-h) or contraction depth %d < 1", s, l, ECP_MAX_LBS, KBS[s]);
      t->shellL[s] = l;
      t->shellK[s] = KBS[s];
      t->shellPrim[s] = nprim;
      t->shellAtom[s] = i;
      t->shellAO[s] = nao;
      nprim += KBS[s];
      nao += IJK(l);
      present[l] = 1;
      if (l > t->atomMaxL[i]) t->atomMaxL[i] = l;
      if (l > v->maxLBS) v->maxLBS = l;
    }
  }
  t->atomFirstShell[nrAtoms] = nsh;
  v->nrShells = nsh;
  v->nrPrims = nprim;
  v->nAO = nao;
  v->shellL = t->shellL;
  v->shellK = t->shellK;
  v->shellPrim = t->shellPrim;
  v->shellAtom = t->shellAtom;
  v->shellAO = t->shellAO;
  v->primD = dBS;
  v->primA = aBS;

  /* ---- ECP parameter sets, de-duplicated (reference builds one ECP struct per atom: src/libecp.c:92-127) ---- */
  int ngTot = 0, nsE = 0;
  for (int i = 0; i < nrAtoms; i++)
    for (int j = 0; j < shellsECP[i]; j++) ngTot += KECP[nsE++];
  t->atomType = malloc((nrAtoms + 1) * sizeof(int));
  t->types = calloc(nrAtoms + 1, sizeof(EcpType));
  t->gaussL = malloc((ngTot + 1) * sizeof(int));
  t->gaussN = malloc((ngTot + 1) * sizeof(double));
  t->gaussD = malloc((ngTot + 1) * sizeof(double));
  t->gaussA = malloc((ngTot + 1) * sizeof(double));
  int ng = 0, si = 0, pi = 0, Lpresent[ECP_MAX_LECP + 1] = {0};
  for (int i = 0; i < nrAtoms; i++) {
    t->atomType[i] = -1;
    if (shellsECP[i] <= 0) continue;
    const int g0 = ng;
    int L = 0;
    for (int j = 0; j < shellsECP[i]; j++, si++) {
      if (lECP[si] > L) L = lECP[si];
      for (int k = 0; k < KECP[si]; k++, pi++, ng++) {
        t->gaussL[ng] = lECP[si];
        t->gaussN[ng] = nECP[pi];
        t->gaussD[ng] = dECP[pi];
        t->gaussA[ng] = aECP[pi];
      }
    }
    if (L > ECP_MAX_LECP) TAB_FAIL("ECP on atom %d: L = %d exceeds the supported maximum %d", i, L, ECP_MAX_LECP);
    const int N = ng - g0;
    int found = -1;
    for (int k = 0; k < v->nTypes && found < 0; k++) {
      const EcpType *T = &t->types[k];
      if (T->L != L || T->N != N) continue;
      found = k;
      for (int q = 0; q < N; q++)
        if (t->gaussL[T->gaussOff + q] != t->gaussL[g0 + q] || t->gaussN[T->gaussOff + q] != t->gaussN[g0 + q] ||
            t->gaussD[T->gaussOff + q] != t->gaussD[g0 + q] || t->gaussA[T->gaussOff + q] != t->gaussA[g0 + q]) {
          found = -1;
          break;
        }
    }
    if (found >= 0) {
      ng = g0; /* drop the duplicate */
      t->atomType[i] = found;
    } else {
      t->types[v->nTypes].L = L;
      t->types[v->nTypes].N = N;
      t->types[v->nTypes].gaussOff = g0;
      t->atomType[i] = v->nTypes++;
      Lpresent[L] = 1;
      if (L > v->maxLECP) v->maxLECP = L;
    }
  }
  /* the row deal of a sharded run only needs the basis bookkeeping: valid on a handle without ECP centres too
   * (libecp_b200_owned_rows / _pair_owner are legal there and every rank of a gather asks for them) */
  if (opts && opts->deriv) {
    t->deriv = opts->deriv;
    t->virtShift = malloc((nsh + 1) * sizeof(int));
    t->virtLocal = malloc((nsh + 1) * sizeof(int));
    memcpy(t->virtShift, opts->virtShift, nsh * sizeof(int));
    memcpy(t->virtLocal, opts->virtLocal, nsh * sizeof(int));
  }
  t->rowDeal = NULL; /* dealt on the first sharded use: ecp_tables_row_deal */
  t->dealL = lBS; t->dealK = KBS; t->dealA = aBS;
  if (v->nTypes == 0) return t; /* no ECP centre: nothing to integrate */

  /* ---- dimensions (reference src/libecp.c:143-171) ---- */
  v->maxAlpha = v->maxLBS;
  v->maxLambda = v->maxLECP - 1 + v->maxAlpha;
  v->tmDim = v->maxLambda + v->maxAlpha;
  v->ijkDim = v->tmDim + 1;
  /* shapes the kernels are built for; maxLBS <= maxLECP+1 keeps every Bessel request inside the table */
  if (v->maxLECP < 1 || v->maxLambda > ECP_KMAX || 2 * v->maxLBS > ECP_KMAX || v->maxLBS > v->maxLECP + 1)
    TAB_FAIL("unsupported shape: max l of the basis %d, max L of the ECPs %d (need L >= 1, L - 1 + l <= %d, l <= L + 1)",
             v->maxLBS, v->maxLECP, ECP_KMAX);
  v->nfac = 2 * v->tmDim + 2;
  t->fac = factorial_table(2 * v->tmDim + 1, 1);
  t->dfac = factorial_table(2 * v->tmDim + 1, 2);
  if (customOrder) {
    /* the caller's component order (reference src/libecp.c:158-166): it must reach one shell beyond tmDim, the index
     * cube has the caller's edge lmax + 1, and the order-dependent tables are built for this handle (no cache) */
    const int lm = opts->lmaxOrd;
    if (lm < v->tmDim + 1)
      TAB_FAIL("shell ordering given up to lmax = %d, needed up to %d (reference src/libecp.c:159-162)", lm, v->tmDim + 1);
    v->ijkDim = lm + 1;
    if (custom_order(customOrder, lm, v->tmDim, &t->ijk, &t->ijkIndex))
      TAB_FAIL("shell ordering: the components of a shell l <= %d are not a permutation of its %s", v->tmDim, "monomials");
    t->cart2sph = build_cart2sph(v->tmDim, t->ijk, t->fac);
    t->poly2sph = build_poly2sph(t->cart2sph, v->tmDim, t->ijk, t->dfac);
    t->omega = build_omega(t, &v->nomega);
  } else {
    cartesian_order(v->tmDim, &t->ijk, &t->ijkIndex);
    if (!(t->cart2sph = cc_get(CC_C2S, v->tmDim, 0, 0, 0, 0.0, NULL))) {
      t->cart2sph = build_cart2sph(v->tmDim, t->ijk, t->fac);
      cc_put(CC_C2S, v->tmDim, 0, 0, 0, 0.0, t->cart2sph, (size_t)c2s_size(v->tmDim, t->fac));
    }
    if (!(t->poly2sph = cc_get(CC_P2S, v->tmDim, 0, 0, 0, 0.0, NULL))) {
      t->poly2sph = build_poly2sph(t->cart2sph, v->tmDim, t->ijk, t->dfac);
      cc_put(CC_P2S, v->tmDim, 0, 0, 0, 0.0, t->poly2sph, (size_t)CD(v->tmDim) * LD(v->tmDim));
    }
  }
  if (!customOrder) {
    size_t nom = 0;
    if ((t->omega = cc_get(CC_OMEGA, v->maxLECP, v->maxLambda, v->maxAlpha, v->tmDim, 0.0, &nom))) {
      v->nomega = (int)nom;
    } else {
      t->omega = build_omega(t, &v->nomega);
      cc_put(CC_OMEGA, v->maxLECP, v->maxLambda, v->maxAlpha, v->tmDim, 0.0, t->omega, (size_t)v->nomega);
    }
  }
  t->binom = calloc((v->maxLBS + 1) * (v->maxLBS + 1), sizeof(double));
  for (int n = 0; n <= v->maxLBS; n++)
    for (int k = 0; k <= n; k++) t->binom[n * (v->maxLBS + 1) + k] = nk(n, k, t->fac);
  { /* term lists of the binomial shift (x-A)^a = sum_alpha binom(a,alpha) (C-A)^(a-alpha) x_C^alpha, in the
     * reference's loop order alpha_x, alpha_y, alpha_z (src/util.c:275-283 / :307-315) */
    const int os = IJK(v->maxLBS) + 1, D = v->ijkDim;
    int nt = 0;
    for (int l = 0; l <= v->maxLBS; l++)
      for (int c = 0; c < IJK(l); c++) {
        const int *e = t->ijk + 3 * CIDX(l, c);
        nt += (e[0] + 1) * (e[1] + 1) * (e[2] + 1);
      }
    t->shTermOff = calloc((size_t)(v->maxLBS + 1) * os, sizeof(int));
    t->shTermP = malloc((nt + 1) * sizeof(int));
    t->shTermD = malloc((nt + 1) * sizeof(int));
    t->shTermBin = malloc((nt + 1) * sizeof(double));
    int k = 0;
    for (int l = 0; l <= v->maxLBS; l++) {
      for (int c = 0; c < IJK(l); c++) {
        const int *e = t->ijk + 3 * CIDX(l, c);
        t->shTermOff[l * os + c] = k;
        for (int x = 0; x <= e[0]; x++) {
          const double bx = nk(e[0], x, t->fac);
          for (int y = 0; y <= e[1]; y++) {
            const double by = bx * nk(e[1], y, t->fac);
            for (int z = 0; z <= e[2]; z++, k++) {
              t->shTermBin[k] = by * nk(e[2], z, t->fac);
              t->shTermP[k] = t->ijkIndex[x * D * D + y * D + z];
              t->shTermD[k] = (e[0] - x) | ((e[1] - y) << 4) | ((e[2] - z) << 8);
            }
          }
        }
      }
      for (int c = IJK(l); c < os; c++) t->shTermOff[l * os + c] = k;
    }
    v->shOffStride = os;
    v->nShTerms = nt;
    v->shTermOff = t->shTermOff;
    v->shTermP = t->shTermP;
    v->shTermD = t->shTermD;
    v->shTermBin = t->shTermBin;
  }
  build_small_grid(t);
  build_large_grid(t, largeGridOrder);
  if (build_bessel(t, v->maxLECP + v->maxAlpha + 6, accuracy))
    TAB_FAIL("Bessel tabulation did not converge (reference src/libecp.c:181-185)");
  v->fac = t->fac;
  v->dfac = t->dfac;
  v->ijk = t->ijk;
  v->ijkIndex = t->ijkIndex;
  v->poly2sph = t->poly2sph;
  v->cart2sph = t->cart2sph;
  v->ncart2sph = c2s_size(v->tmDim, t->fac);
  v->omega = t->omega;
  v->binom = t->binom;
  v->small_r = t->small_rs;
  v->small_w = t->small_ws;
  v->small_oidx = t->small_oidx;
  v->large_x = t->large_xs;
  v->large_w = t->large_ws;
  v->large_xo = t->large_x;
  v->large_oidx = t->large_oidx;
  v->besselT = t->besselT;
  v->besselC = t->besselC;

  /* ---- shell radii (reference src/type2.c:251 recomputes them per centre; they do not depend on it) ---- */
  t->shellRadius = malloc((nsh + 1) * sizeof(double));
  for (int s = 0; s < nsh; s++) {
    const int ps = (opts && opts->screenParent) ? opts->screenParent[s] : s; /* a shifted shell is screened as its parent */
    t->shellRadius[s] = shell_radius(KBS[ps], lBS[ps], dBS + t->shellPrim[ps], aBS + t->shellPrim[ps], 1.0E-14);
  }
  t->atomRmax = calloc(nrAtoms + 1, sizeof(double));
  for (int s = 0; s < nsh; s++)
    if (t->shellRadius[s] > t->atomRmax[t->shellAtom[s]]) t->atomRmax[t->shellAtom[s]] = t->shellRadius[s];

  /* ---- per ECP type: r^N U_l(r) with the cumulative cut-off, and the local channel
   *      (reference src/type2.c:184-219, src/libecp.c:269-270).  Rows N <= max(maxLambda, 2 maxAlpha) so
   *      that la+lb never indexes past the table (the reference does for maxLBS > maxLECP-1). */
  const int nU = (v->maxLambda > 2 * v->maxAlpha ? v->maxLambda : 2 * v->maxAlpha) + 1;
  v->nU = nU;
  t->typeL = malloc(v->nTypes * sizeof(int));
  t->typeGaussOff = malloc((v->nTypes + 1) * sizeof(int));
  t->typeUtab = calloc((size_t)v->nTypes * v->maxLECP * nU * ECP_SMALL_SLOTS, sizeof(double));
  t->typeUL = calloc((size_t)v->nTypes * ECP_SMALL_SLOTS, sizeof(double));
  double *Ul = malloc(ECP_SMALL_ORDER * sizeof(double));
  for (int k = 0; k < v->nTypes; k++) {
    EcpType *T = &t->types[k];
    t->typeL[k] = T->L;
    t->typeGaussOff[k] = T->gaussOff;
    int end = ECP_SMALL_ORDER - 1;
    for (int l = 0; l < T->L; l++) {
      for (int n = 0; n < ECP_SMALL_ORDER; n++) Ul[n] = ecp_host_pot_eval(t, k, l, t->small_x[n]);
      int cut = -1; /* potentialScreening: src/util.c:198-210 */
      for (int i = end; i >= 0; i--)
        if (fabs(Ul[i]) > 1.0E-14) {
          cut = i;
          break;
        }
      end = cut;
      T->endl[l] = end;
      double *base = t->typeUtab + ((size_t)(k * v->maxLECP + l) * nU) * ECP_SMALL_SLOTS;
      for (int s = 0; s < ECP_SMALL_SLOTS; s++) {
        const int n = t->small_oidx[s];
        if (n < 0 || n > end) continue;
        double p = Ul[n];
        base[s] = p;
        for (int lab = 1; lab < nU; lab++) {
          p = p * t->small_x[n];
          base[(size_t)lab * ECP_SMALL_SLOTS + s] = p;
        }
      }
    }
    T->endLast = end;
    for (int s = 0; s < ECP_SMALL_SLOTS; s++)
      if (t->small_oidx[s] >= 0)
        t->typeUL[(size_t)k * ECP_SMALL_SLOTS + s] = ecp_host_pot_eval(t, k, T->L, t->small_x[t->small_oidx[s]]);
  }
  free(Ul);
  t->typeGaussOff[v->nTypes] = ng;
  v->typeL = t->typeL;
  v->typeGaussOff = t->typeGaussOff;
  v->gaussL = t->gaussL;
  v->gaussN = t->gaussN;
  v->gaussD = t->gaussD;
  v->gaussA = t->gaussA;
  v->typeUtab = t->typeUtab;
  v->typeUL = t->typeUL;

  /* ---- classes (la, lb, L) and their used-quadrature lists ---- */
  memset(t->clsLookup, -1, sizeof(t->clsLookup));
  int nc = 0, nq = 0, nqi = 0;
  for (int pass = 0; pass < 2; pass++) {
    nc = nq = nqi = 0;
    for (int L = 0; L <= v->maxLECP; L++) {
      if (!Lpresent[L]) continue;
      for (int la = 0; la <= v->maxLBS; la++) {
        if (!present[la]) continue;
        for (int lb = 0; lb <= v->maxLBS; lb++) {
          if (!present[lb]) continue;
          const int d1 = la + L, d2 = lb + L, d3 = la + lb + 1;
          if (pass) {
            t->clsLookup[la][lb][L] = nc;
            t->clsLa[nc] = la;
            t->clsLb[nc] = lb;
            t->clsL[nc] = L;
            t->clsQOff[nc] = nq;
            t->clsQidxOff[nc] = nqi;
            for (int i = 0; i < L * d1 * d2 * d3; i++) t->qidx[nqi + i] = -1;
          }
          int k = 0;
          for (int l = 0; l < L; l++) {
            const int kstart = k;
            if (pass) t->clsQlOff[nc * (ECP_MAX_LECP + 1) + l] = k;
            for (int l1 = 0; l1 <= la + l; l1++)
              for (int l2 = 0; l2 <= lb + l; l2++)
                for (int l3 = 0; l3 <= la + lb; l3++) {
                  if (!ecp_t2_used(la, lb, l, l1, l2, l3)) continue;
                  if (pass) {
                    t->qlist[nq + k] = l | (l1 << 4) | (l2 << 8) | (l3 << 12);
                    t->qidx[nqi + ((l * d1 + l1) * d2 + l2) * d3 + l3] = (int16_t)k;
                  }
                  k++;
                }
            if (k - kstart > v->maxQPerL) v->maxQPerL = k - kstart;
          }
          if (pass) {
            for (int l = L; l <= ECP_MAX_LECP; l++) t->clsQlOff[nc * (ECP_MAX_LECP + 1) + l] = k;
            t->clsNq[nc] = k;
          }
          nq += k;
          nqi += L * d1 * d2 * d3;
          nc++;
        }
      }
    }
    if (!pass) {
      if (nc > ECP_MAX_CLASSES) TAB_FAIL("more than %d (la, lb, L) classes", ECP_MAX_CLASSES);
      t->clsLa = malloc((nc + 1) * sizeof(int));
      t->clsLb = malloc((nc + 1) * sizeof(int));
      t->clsL = malloc((nc + 1) * sizeof(int));
      t->clsNq = malloc((nc + 1) * sizeof(int));
      t->clsQOff = malloc((nc + 1) * sizeof(int));
      t->clsQidxOff = malloc((nc + 1) * sizeof(int));
      t->clsQlOff = calloc((size_t)(nc + 1) * (ECP_MAX_LECP + 1), sizeof(int));
      t->qlist = malloc((nq + 1) * sizeof(int));
      t->qidx = malloc((nqi + 1) * sizeof(int16_t));
    }
  }
  t->clsQOff[nc] = nq;
  t->clsQidxOff[nc] = nqi;
  v->nClasses = nc;
  v->clsLa = t->clsLa;
  v->clsLb = t->clsLb;
  v->clsL = t->clsL;
  v->clsNq = t->clsNq;
  v->clsQOff = t->clsQOff;
  v->clsQlOff = t->clsQlOff;
  v->qlist = t->qlist;
  v->clsQidxOff = t->clsQidxOff;
  v->qidx = t->qidx;
  v->nqlist = nq;
  v->nqidx = nqi;
  return t;
}

void ecp_tables_free(EcpTables *t) {
  if (!t) return;
  free(t->fac); free(t->dfac); free(t->cart2sph); free(t->poly2sph); free(t->omega); free(t->binom);
  free(t->shTermOff); free(t->shTermP); free(t->shTermD); free(t->shTermBin);
  free(t->ijk); free(t->ijkIndex);
  free(t->virtShift); free(t->virtLocal);
  free(t->small_x); free(t->small_w); free(t->small_rs); free(t->small_ws); free(t->small_oidx);
  free(t->large_x); free(t->large_w); free(t->large_xs); free(t->large_ws); free(t->large_oidx);
  free(t->besselK); free(t->besselT); free(t->besselC);
  free(t->shellL); free(t->shellK); free(t->shellPrim); free(t->shellAtom); free(t->shellAO);
  free(t->shellRadius); free(t->atomRmax); free(t->rowDeal); free(t->atomMaxL); free(t->atomFirstShell);
  free(t->atomType); free(t->types); free(t->typeL); free(t->typeGaussOff); free(t->gaussL);
  free(t->gaussN); free(t->gaussD); free(t->gaussA); free(t->typeUtab); free(t->typeUL);
  free(t->clsLa); free(t->clsLb); free(t->clsL); free(t->clsNq); free(t->clsQOff); free(t->clsQlOff);
  free(t->qlist); free(t->clsQidxOff); free(t->qidx);
  free(t);
}
