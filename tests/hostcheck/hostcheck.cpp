// hostcheck.cpp - TEST-ONLY: compiles the per-thread device math (libecp_b200/csrc/ecp_math.h) with g++
// so the functions the sm_100a kernels execute can be compared with the oracle on the CPU tier.
// Never linked into the product.
#include "ecp_math.h"

extern "C" {
int hc_bessel(const double *tabT, int stride, const double *Cj, int lmax, double z, double *K) {
  double k[ECP_KMAX + 1];
  for (int i = 0; i <= ECP_KMAX; i++) k[i] = 0.0;
  int br = ecp_bessel<ECP_KMAX>(tabT, stride, Cj, lmax, z, k);
  for (int i = 0; i <= lmax; i++) K[i] = k[i];
  return br;
}
int hc_bessel_mem(const double *tabT, int stride, const double *Cj, int lmax, double z, double *K) {
  double d[ECP_KMAX + 8];
  return ecp_bessel_mem(tabT, stride, Cj, lmax, z, K, d);
}
void hc_rsh(int lmax, double theta, double phi, const double *fac, const double *dfac, double *out) {
  ecp_rsh(lmax, theta, phi, fac, dfac, out);
}
void hc_sphcoord(const double *v, double *rtp) { ecp_sphcoord(v[0], v[1], v[2], rtp, rtp + 1, rtp + 2); }
int hc_ps93_fastT(const double *Fa, const double *Fb, const double *U, const double *w, const int *oidx32,
                  const int *meta /* pairs[13] j[13] n[13] slot[14] */, int start, int end, double tol, double *res,
                  int *npts) {
  EcpSmallMeta m;
  int16_t oidx[ECP_SMALL_SLOTS];
  for (int i = 0; i < ECP_SMALL_SLOTS; i++) oidx[i] = (int16_t)oidx32[i];
  for (int i = 0; i < ECP_SMALL_LEVELS; i++) {
    m.levPairs[i] = meta[i];
    m.levJ[i] = meta[13 + i];
    m.levN[i] = meta[26 + i];
  }
  for (int i = 0; i <= ECP_SMALL_LEVELS; i++) m.levSlot[i] = meta[39 + i];
  ecp_small_meta_bounds(&m, oidx);
  static unsigned char jL[ECP_SMALL_LEVELS * ECP_SMALL_SLOTS], jR[ECP_SMALL_LEVELS * ECP_SMALL_SLOTS];
  if (!ecp_small_suffix_tables(&m, oidx, jL, jR)) return -1;
  return ecp_ps93_fastT(Fa, 1, Fb, 1, U, 1, w, &m, jL, jR, start, end, tol, res, npts);
}
// level-wise point enumeration of the type-1 / large-grid kernels (ecp_math.h: t1_level, lg_live_range, lg_level)
static void hc_meta(const int *meta, const int *oidx32, EcpSmallMeta *m, int16_t *oidx) {
  for (int i = 0; i < ECP_SMALL_SLOTS; i++) oidx[i] = (int16_t)oidx32[i];
  for (int i = 0; i < ECP_SMALL_LEVELS; i++) {
    m->levPairs[i] = meta[i];
    m->levJ[i] = meta[13 + i];
    m->levN[i] = meta[26 + i];
  }
  for (int i = 0; i <= ECP_SMALL_LEVELS; i++) m->levSlot[i] = meta[39 + i];
  ecp_small_meta_bounds(m, oidx);
}
// slots the kernel visits on small-grid level v for the window [gs, ge): returns their number, *cnt = PS93 point count
int hc_t1_level_slots(const int *oidx32, const int *meta, int v, int gs, int ge, int *slots, int *cnt) {
  EcpSmallMeta m;
  int16_t oidx[ECP_SMALL_SLOTS];
  hc_meta(meta, oidx32, &m, oidx);
  static unsigned char jL[ECP_SMALL_LEVELS * ECP_SMALL_SLOTS], jR[ECP_SMALL_LEVELS * ECP_SMALL_SLOTS];
  if (!ecp_small_suffix_tables(&m, oidx, jL, jR)) return -1;
  const T1Level L = t1_level(&m, jL, jR, v, gs, ge);
  *cnt = L.cnt;
  for (int k = 0; k < L.nLive; k++) slots[k] = (k < L.nLl) ? L.s0 + 2 * (L.jLa + k) : L.s0 + 2 * (L.jRa + k - L.nLl) + 1;
  return L.nLive;
}
// candidate slots of large-grid level lev for the gate a r^2 + b r + cmln >= 0 on the grid r = i1 x + i2
int hc_lg_level_slots(const double *xo, int order, int slotsTotal, double a, double b, double cmln, double i1, double i2,
                      int lev, int *slots) {
  const LgRange R = lg_live_range(xo, order, a, b, cmln, i1, i2);
  const LgLevel L = lg_level(slotsTotal, order, R, lev);
  for (int k = 0; k < L.nLive; k++) slots[k] = lg_slot(L, lev, k);
  return L.nLive;
}
void hc_fm06_map(double zp, double P, double *i1, double *i2) { ecp_fm06_map(zp, P, i1, i2); }
double hc_pot_eval(const int *gl, const double *gn, const double *gd, const double *ga, int n, int l, double r) {
  return ecp_pot_eval(gl, gn, gd, ga, 0, n, l, r);
}
}
