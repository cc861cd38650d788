/* ecp_cuda.cu - hand-written FP64 sm_100a kernels for the ECP hot path + the thin C-ABI device layer.
 *
 * Compiled with: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
 * (FMA contraction on; values an integer decision reads are pinned with __dmul_rn / __dadd_rn, see ecp_math.h).
 *
 * Kernel map (reference code each one replaces, paths relative to /root/reference):
 *   k_triprep    one record per triple + the pair -> triple map                 (indices only)
 *   k_atomslot   S_lm(r_XC) and (-x)^i(-y)^j(-z)^k per (C, atom)          src/util.c:214-243, spherical_harmonics.c
 *   k_omegaX     Omega_X = sum_mu S_lam,mu Omega per (C, atom)            src/angular_integrals.c:104-142
 *   k_Ftab       contracted radial table F_lambda(r_n) per (C, shell)     src/type2.c:284-303
 *   k_fastT(2)   type-2 fast path, PS93 on Fa*Fb*r^N U_l, two launches    src/type2.c:336-381
 *   k_fallbackG  type-2 large-grid fallback, PSM92 per primitive pair     src/type2.c:417-528   (ecp_fallback.cuh)
 *   k_link       gamma = sum Omega_A Omega_B T, one launch per class      src/type2.c:583-623
 *   k_link4      the same for the large classes: slices in shared memory, compile-time strides, runs of triples
 *   k_t1prep     P, |P|, S_lm(P^), pair record per primitive pair         src/type1.c:235-252
 *   k_type1S/L   radial Q(N,lambda): PS93 small grid, PSM92 fallback      src/type1.c:94-208    (ecp_type1.cuh)
 *   k_chi        chi = sum poly2sph (sum_pairs S Q), 8 lanes per triple   src/type1.c:266-295
 *   k_shift      binomial shift to A/B, x4pi / x16pi^2, block + matrix    src/util.c:246-334, getIntegrals.c:22-43
 *   k_zero_rows, k_pack_rows, k_unpack_rows   result matrix: partial clear, packed rows for D2H / the all-gather
 */
#include <cuda_runtime.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ecp_dev.h"
#include "ecp_math.h"

#define KM ECP_KMAX
#define ECP_MAXDEV 16 /* devices with per-device caches (parked scratch, occupancy figures) */


static char g_err[512] = "";
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      snprintf(g_err, sizeof(g_err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return (int)e_;                                                                              \
    }                                                                                              \
  } while (0)

/* device-side view of all tables of a handle + the current batch */
struct DevT {
  int maxLECP, maxLBS, maxLambda, tmDim, ijkDim, besselStride, nU, pcols;
  int largeSlots, largeLevels, largeOrder, nClasses, omD2, omD3;
  double tolerance, accuracy, lnAcc1, lnAcc2;
  EcpSmallMeta sm;
  const double *fac, *dfac, *poly2sph, *omega, *binom;
  int shOffStride;
  const int *shTermOff, *shTermP, *shTermD;
  const double *shTermBin;
  const int *ijk, *ijkIndex;
  const double *small_r, *small_w, *large_x, *large_w, *large_xo;
  const int16_t *small_oidx;
  const int16_t *small_slotOf; /* original index -> slot (inverse of small_oidx) */
  const unsigned char *small_jL, *small_jR; /* first in-window pair per (level, start) / (level, end), ecp_math.h */
  const double *besselT, *besselC;
  /* non-zero entries of poly2sph per monomial, in the order k_chi adds them (l = N, N-2, ...; m ascending): position
   * in the packed R of the monomial's degree, bit 15 set on the last entry of an l group; ~25 % of the dense rows */
  const int *p2sOff;
  const unsigned short *p2sIdx;
  const double *p2sVal;
  const int *shellL, *shellK, *shellPrim, *shellAtom, *shellAO, *atomMaxL;
  const double *primD, *primA;
  const int *typeL, *typeGaussOff, *gaussL;
  const double *gaussN, *gaussD, *gaussA, *typeUtabT, *typeUL; /* typeUtabT: [type][slot][l][N] */
  const int *clsLa, *clsLb, *clsL, *clsNq, *clsQOff, *clsQlOff, *qlist, *clsQidxOff;
  const int16_t *qidx;
  const int *clsLookup; /* [(la (ECP_MAX_LBS+1) + lb) (ECP_MAX_LECP+1) + L] -> class or -1 */
  int nAO;
};
struct T1Rec;
struct TriRec;
struct DevB {
  int nASlots, nSSlots, nTriples;
  long long nPairs;
  const int *asAtom, *asType;
  const double *asR;
  const long long *asOmOff;
  const int *ssShell, *ssASlot, *ssStart, *ssEnd;
  const long long *ssFOff;
  const int *trA, *trB;
  const long long *trOut, *trPair;
  const int *prTriple;
  const int *clsFirst;
  const long long *clsWork, *clsElem, *clsOutElem, *clsPairBase, *clsQBase;
  /* intermediates */
  double *rshX, *uspX, *omX, *F, *T, *gamma, *chi, *Q, *rshP, *sP, *blocks, *matrix;
  unsigned char *tfail;
  int *tflags, *items;
  T1Rec *t1rec; /* per primitive pair, written by k_t1prep (ecp_type1.cuh) */
  TriRec *trirec; /* per triple, written by k_triprep */
  unsigned long long *dbg; /* LIBECP_B200_TAILS: per-block (start, end) globaltimer of the persistent kernels, else NULL */
  int *counters; /* [0] nItems [1] work counter [2] err1 [3] err2 [4] nFastFail [5] nType1Fail [6] stale [7] work2 */
};
__device__ __forceinline__ unsigned long long ecp_gtimer() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
  return v;
}
#define DBG_STRIDE 8192 /* blocks per probed launch */
#define RSHX_STRIDE 121
#define USPX_STRIDE 216

/* small-grid weights / abscissae / original indices in slot order.  Every lane of a warp visits the slots in lock
 * step, so these reads are warp-uniform: constant memory serves them without touching the LSU pipe.  The small grid
 * is the same for every handle (order 128, KK map; reference src/type1.c:24,56-58, src/type2.c:60,90-92). */
__constant__ double c_small_w[ECP_SMALL_SLOTS];
__constant__ double c_small_r[ECP_SMALL_SLOTS];
__constant__ int16_t c_small_oidx[ECP_SMALL_SLOTS];

/* ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int find_class(const long long *prefix, int nc, long long w) {
  int lo = 0, hi = nc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= w)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ int find_class_i(const int *prefix, int nc, int w) {
  int lo = 0, hi = nc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= w)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}
/* per-triple / per-pair offsets are closed forms of the class prefixes (nothing per triple is uploaded for them) */
__device__ __forceinline__ long long tri_T_off(const DevT &t, const DevB &b, int c, int tri) {
  return b.clsWork[c] + (long long)(tri - b.clsFirst[c]) * t.clsNq[c];
}
__device__ __forceinline__ long long tri_G_off(const DevT &t, const DevB &b, int c, int tri) {
  return b.clsElem[c] + (long long)(tri - b.clsFirst[c]) * (ecp_cd(t.clsLa[c]) * ecp_cd(t.clsLb[c]));
}
__device__ __forceinline__ long long pair_Q_off(const DevT &t, const DevB &b, int c, long long pr) {
  const int lab1 = t.clsLa[c] + t.clsLb[c] + 1;
  return b.clsQBase[c] + (pr - b.clsPairBase[c]) * (lab1 * lab1);
}

/* ---- per (C, atom) : S_lm(r^_XC) and unit-sphere monomials ---- */
__global__ void k_atomslot(DevT t, DevB b) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= b.nASlots) return;
  const double x = b.asR[4 * a], y = b.asR[4 * a + 1], z = b.asR[4 * a + 2];
  const int lX = t.atomMaxL[b.asAtom[a]];
  const int lmax = t.maxLECP - 1 + lX; /* reference src/angular_integrals.c:110 */
  double r, th, ph;
  ecp_sphcoord(x, y, z, &r, &th, &ph);
  ecp_rsh(lmax, th, ph, t.fac, t.dfac, b.rshX + (size_t)a * RSHX_STRIDE);
  /* (-x)^i (-y)^j (-z)^k, reference src/util.c:214-243 */
  double *u = b.uspX + (size_t)a * USPX_STRIDE;
  const int dim = lX + 1;
  double px = 0, py = 0, pz = 0;
  for (int i = 0; i <= lX; i++) {
    px = (i == 0) ? 1.0 : -px * x;
    for (int j = 0; j <= lX - i; j++) {
      py = (j == 0) ? 1.0 : -py * y;
      for (int k = 0; k <= lX - i - j; k++) {
        pz = (k == 0) ? 1.0 : -pz * z;
        u[i * dim * dim + j * dim + k] = px * py * pz;
      }
    }
  }
}

/* ---- Omega_X[lambda][(l,m)][C_INDEX(a,c)], one block per atom slot ---- */
__global__ void k_omegaX(DevT t, DevB b) {
  const int a = blockIdx.x;
  const int lX = t.atomMaxL[b.asAtom[a]];
  const int Lc = t.typeL[b.asType[a]];
  const int nlam = Lc + lX, nlm = Lc * Lc, ncd = ecp_cd(lX);
  const double *rsh = b.rshX + (size_t)a * RSHX_STRIDE;
  double *out = b.omX + b.asOmOff[a];
  const int total = nlam * nlm * ncd;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int c = e % ncd, lm = (e / ncd) % nlm, lam = e / (ncd * nlm);
    double v = 0.0;
    const double *om = t.omega + ((size_t)lm * t.omD2 + lam * lam) * t.omD3 + c;
    for (int mu = 0; mu < 2 * lam + 1; mu++) v += rsh[lam * lam + mu] * om[(size_t)mu * t.omD3];
    out[e] = v;
  }
}

/* ---- F_lambda(r_n) per shell slot; one block of 384 threads per slot ----
 * The Bessel order bound lmaxA = L - 1 + l_shell is uniform over the block: the body is instantiated per value, so the
 * unrolled Taylor recurrence of ecp_bessel carries no masked-off orders (with the single KM = 10 instantiation an s
 * shell under an L = 2 potential, lmaxA = 1, paid for 16 orders). */
template <int LM>
__device__ __forceinline__ void ftab_body(const DevT &t, const DevB &b, int ss, int k, int sh, double dAC) {
  const int oi = t.small_oidx[k];
  double acc[LM + 1];
#pragma unroll
  for (int i = 0; i <= LM; i++) acc[i] = 0.0;
  if (oi >= b.ssStart[ss] && oi < b.ssEnd[ss]) { /* exclusive end: src/type2.c:292 */
    const double r = t.small_r[k];
    const int p0 = t.shellPrim[sh], np = t.shellK[sh];
    for (int p = 0; p < np; p++) {
      const double zeta = t.primA[p0 + p], da = t.primD[p0 + p];
      double K[LM + 1];
      ecp_bessel<LM>(t.besselT, t.besselStride, t.besselC, LM, 2.0 * zeta * dAC * r, K);
      double e = dAC - r;
      e = exp(-zeta * e * e);
#pragma unroll
      for (int i = 0; i <= LM; i++) acc[i] += da * K[i] * e;
    }
  }
  /* slot-major: F[slot][lambda], lambda = 0..lmaxA (the rows a fast-path quadrature multiplies sit side by side) */
  double *F = b.F + (size_t)b.ssFOff[ss] * ECP_SMALL_SLOTS + (size_t)k * (LM + 1);
#pragma unroll
  for (int i = 0; i <= LM; i++) F[i] = acc[i];
}
/* LMAX: highest order bound of the handle (L_max - 1 + l_max).  The usual shapes (s-f shells, L <= 4: LMAX = 6) fit two
 * blocks per SM in registers; the deep-table stress shapes take the one-block variant. */
template <int LMAX>
__global__ void __launch_bounds__(ECP_SMALL_SLOTS, (LMAX <= 6 ? 2 : 1)) k_Ftab(DevT t, DevB b) {
  const int ss = blockIdx.x, k = threadIdx.x;
  const int sh = b.ssShell[ss], as = b.ssASlot[ss];
  const int Lc = t.typeL[b.asType[as]];
  const int lmaxA = Lc - 1 + t.shellL[sh];
  const double dAC = b.asR[4 * as + 3];
  switch (lmaxA) {
    case 0: ftab_body<0>(t, b, ss, k, sh, dAC); break;
    case 1: ftab_body<1>(t, b, ss, k, sh, dAC); break;
    case 2: ftab_body<2>(t, b, ss, k, sh, dAC); break;
    case 3: ftab_body<3>(t, b, ss, k, sh, dAC); break;
    case 4: ftab_body<4>(t, b, ss, k, sh, dAC); break;
    case 5: ftab_body<5>(t, b, ss, k, sh, dAC); break;
    default:
      if (LMAX <= 6 || lmaxA == 6) {
        ftab_body<6>(t, b, ss, k, sh, dAC);
      } else {
        switch (lmaxA) {
          case 7: ftab_body<(LMAX > 6 ? 7 : 6)>(t, b, ss, k, sh, dAC); break;
          case 8: ftab_body<(LMAX > 6 ? 8 : 6)>(t, b, ss, k, sh, dAC); break;
          case 9: ftab_body<(LMAX > 6 ? 9 : 6)>(t, b, ss, k, sh, dAC); break;
          default: ftab_body<(LMAX > 6 ? KM : 6)>(t, b, ss, k, sh, dAC); break;
        }
      }
      break;
  }
}

/* Window-only variant (default since round 2; LIBECP_B200_FTAB=full keeps k_Ftab): a shell's window covers ~30 % of the
 * grid, so 70 % of k_Ftab's threads only write zeros and every warp (32 consecutive level-major slots span the whole
 * radial range) keeps a few live lanes.  Here the F rows are cleared by one memset and a block of 128 threads walks the
 * ORIGINAL indices of the window [start, end) only, writing each point to its slot: the same arithmetic per point (F is
 * bit-identical, test_ftab_variants_are_bit_identical), about a third of the warps; 3.03 -> 2.64 ms per config-5 pass. */
template <int LMAX>
__global__ void __launch_bounds__(128, (LMAX <= 6 ? 7 : 5)) k_Ftab2(DevT t, DevB b) {
  const int ss = blockIdx.x;
  const int sh = b.ssShell[ss], as = b.ssASlot[ss];
  const int Lc = t.typeL[b.asType[as]];
  const int lmaxA = Lc - 1 + t.shellL[sh];
  const double dAC = b.asR[4 * as + 3];
  const int gs = b.ssStart[ss], ge = min(b.ssEnd[ss], ECP_SMALL_ORDER);
  for (int oi = gs + threadIdx.x; oi < ge; oi += blockDim.x) {
    const int k = t.small_slotOf[oi];
    switch (lmaxA) {
      case 0: ftab_body<0>(t, b, ss, k, sh, dAC); break;
      case 1: ftab_body<1>(t, b, ss, k, sh, dAC); break;
      case 2: ftab_body<2>(t, b, ss, k, sh, dAC); break;
      case 3: ftab_body<3>(t, b, ss, k, sh, dAC); break;
      case 4: ftab_body<4>(t, b, ss, k, sh, dAC); break;
      case 5: ftab_body<5>(t, b, ss, k, sh, dAC); break;
      default:
        if (LMAX <= 6 || lmaxA == 6) {
          ftab_body<6>(t, b, ss, k, sh, dAC);
        } else {
          switch (lmaxA) {
            case 7: ftab_body<(LMAX > 6 ? 7 : 6)>(t, b, ss, k, sh, dAC); break;
            case 8: ftab_body<(LMAX > 6 ? 8 : 6)>(t, b, ss, k, sh, dAC); break;
            case 9: ftab_body<(LMAX > 6 ? 9 : 6)>(t, b, ss, k, sh, dAC); break;
            default: ftab_body<(LMAX > 6 ? KM : 6)>(t, b, ss, k, sh, dAC); break;
          }
        }
        break;
    }
  }
}

/* ---- type-2 fast path: one thread per used quadrature, two densely packed launches ----
 * About 7 % of the used quadratures never converge on the small grid and walk their whole window (a few hundred
 * points) while the typical one stops after 11-31 points; they hold more than half of all evaluated points.  With one
 * launch the lanes of a warp idle until its longest quadrature is through (16 of 32 lanes active in round 1).
 * k_fastT runs levels [0, lim) for every used quadrature and appends the unconverged ones - with their (I, p, q) - to a
 * survivor list (warp-aggregated atomic); k_fastT2 continues exactly those, one per thread, from level lim.
 * The operations of a quadrature and their order are unchanged. */
struct FastSurv {
  EcpPs93State st;
  long long w;
};
struct FastQ { /* operands of one used quadrature */
  const double *Fa, *Fb, *U;
  int strA, strB, strU, gs, ge, tri, l;
};
__device__ __forceinline__ FastQ fast_load(const DevT &t, const DevB &b, long long w) {
  FastQ f;
  const int c = find_class(b.clsWork, t.nClasses, w);
  const int nq = t.clsNq[c];
  const long long idx = w - b.clsWork[c];
  int k;
  if (idx < 0x7fffffffLL) { /* (almost always) a 32-bit division instead of two 64-bit ones */
    const unsigned u = (unsigned)idx, tl = u / (unsigned)nq;
    f.tri = b.clsFirst[c] + (int)tl;
    k = (int)(u - tl * (unsigned)nq);
  } else {
    f.tri = b.clsFirst[c] + (int)(idx / nq);
    k = (int)(idx % nq);
  }
  const int q = t.qlist[t.clsQOff[c] + k];
  const int l1 = (q >> 4) & 15, l2 = (q >> 8) & 15, l3 = (q >> 12) & 15;
  f.l = q & 15;
  const int sa = b.trA[f.tri], sb = b.trB[f.tri];
  const int type = b.asType[b.ssASlot[sa]];
  const int Lc = t.clsL[c];
  f.strA = Lc + t.clsLa[c];
  f.strB = Lc + t.clsLb[c];
  f.strU = t.maxLECP * t.nU;
  f.Fa = b.F + (size_t)b.ssFOff[sa] * ECP_SMALL_SLOTS + l1;
  f.Fb = b.F + (size_t)b.ssFOff[sb] * ECP_SMALL_SLOTS + l2;
  f.U = t.typeUtabT + (size_t)type * ECP_SMALL_SLOTS * f.strU + f.l * t.nU + l3;
  f.gs = max(b.ssStart[sa], b.ssStart[sb]); /* src/libecp.c:315-316 */
  f.ge = max(b.ssEnd[sa], b.ssEnd[sb]);
  return f;
}
__device__ __forceinline__ void fast_store(const DevB &b, long long w, int rc, double res, int tri, int l) {
  if (rc) { /* failed on the small grid: the (triple, l) item goes to the large-grid fallback */
    b.T[w] = 0.0;
    b.tfail[w] = 1;
    atomicAdd(&b.counters[4], 1);
    const int old = atomicOr(&b.tflags[tri], 1 << l);
    if (!(old & (1 << l))) b.items[atomicAdd(&b.counters[0], 1)] = tri * 8 + l;
  } else {
    b.T[w] = res; /* position = class T base + (triple - first) * nq + k */
    b.tfail[w] = 0;
  }
}
/* (register bounds for 12 / 16 resident blocks were measured: 0.565 -> 0.66 ms on Au20, spills; left to the compiler) */
__global__ void __launch_bounds__(128) k_fastT(DevT t, DevB b, long long nWork, int lim, FastSurv *surv, int survCap) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = w < nWork;
  int rc = 0, tri = 0, l = 0;
  double res = 0.0;
  EcpPs93State st;
  if (valid) {
    const FastQ f = fast_load(t, b, w);
    tri = f.tri;
    l = f.l;
    rc = ecp_ps93_fastT_levels(f.Fa, f.strA, f.Fb, f.strB, f.U, f.strU, c_small_w, &t.sm, t.small_jL, t.small_jR, f.gs, f.ge,
                               t.tolerance, 0, lim, &st, &res, (int *)0);
  }
  /* unconverged after level lim-1: to the survivor list (one atomic per warp) */
  const bool sv = valid && rc == 2;
  const unsigned m = __ballot_sync(0xffffffffu, sv);
  if (m) {
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&b.counters[8], __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (sv) {
      const int pos = base + __popc(m & ((1u << lane) - 1));
      if (pos < survCap) {
        surv[pos].st = st;
        surv[pos].w = w;
      } else { /* list full: finish here */
        const FastQ f = fast_load(t, b, w);
        rc = ecp_ps93_fastT_levels(f.Fa, f.strA, f.Fb, f.strB, f.U, f.strU, c_small_w, &t.sm, t.small_jL, t.small_jR, f.gs, f.ge,
                                   t.tolerance, lim, ECP_SMALL_LEVELS, &st, &res, (int *)0);
      }
    }
  }
  if (valid && rc != 2) fast_store(b, w, rc, res, tri, l);
}
template <int UNR>
__global__ void __launch_bounds__(128) k_fastT2(DevT t, DevB b, int lim, const FastSurv *surv, int survCap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(b.counters[8], survCap);
  if (i >= n) return;
  EcpPs93State st = surv[i].st;
  const long long w = surv[i].w;
  const FastQ f = fast_load(t, b, w);
  double res = 0.0;
  const int rc = ecp_ps93_fastT_levels_u<UNR>(f.Fa, f.strA, f.Fb, f.strB, f.U, f.strU, c_small_w, &t.sm, t.small_jL, t.small_jR,
                                              f.gs, f.ge, t.tolerance, lim, ECP_SMALL_LEVELS, &st, &res, (int *)0);
  fast_store(b, w, rc, res, f.tri, f.l);
}

#include "ecp_fallback.cuh"
#include "ecp_waves.cuh"
#include "ecp_enum.cuh"

/* ---- per triple: everything the element-parallel kernels (link, chi, shift) would otherwise chase through
 * trA/trB -> ssShell/ssASlot -> asAtom/asOmOff/shellK/... with up to five dependent loads PER ELEMENT.  Those kernels
 * are bound by exactly that latency; one 64-byte record per triple (read as a warp broadcast) cuts the chain to one. */
struct __align__(16) TriRec {
  long long omA, omB; /* offsets of Omega_A / Omega_B in omX                               */
  long long q0;       /* offset of Q / S_lm(P^) of the first primitive pair                */
  int np;             /* primitive pairs Na * Nb                                           */
  int asa, asb;       /* atom slots (unit-sphere monomial tables)                          */
  int incA1, incB1;   /* C_DIM(lmax of atom A / B): stride of Omega_X over (l,m)           */
  int dA, dB;         /* lmax + 1 of atom A / B                                            */
  int rowAO, colAO;   /* first AO of shell a / b                                           */
  int pad;
};
__global__ void k_triprep(DevT t, DevB b) {
  const int tri = blockIdx.x * blockDim.x + threadIdx.x;
  if (tri >= b.nTriples) return;
  const int c = find_class_i(b.clsFirst, t.nClasses, tri);
  const int ssa = b.trA[tri], ssb = b.trB[tri];
  const int sha = b.ssShell[ssa], shb = b.ssShell[ssb];
  const int asa = b.ssASlot[ssa], asb = b.ssASlot[ssb];
  const int lXa = t.atomMaxL[b.asAtom[asa]], lXb = t.atomMaxL[b.asAtom[asb]];
  TriRec r;
  r.omA = b.asOmOff[asa];
  r.omB = b.asOmOff[asb];
  r.q0 = pair_Q_off(t, b, c, b.trPair[tri]);
  r.np = t.shellK[sha] * t.shellK[shb];
  r.asa = asa;
  r.asb = asb;
  r.incA1 = ecp_cd(lXa);
  r.incB1 = ecp_cd(lXb);
  r.dA = lXa + 1;
  r.dB = lXb + 1;
  r.rowAO = t.shellAO[sha];
  r.colAO = t.shellAO[shb];
  r.pad = 0;
  b.trirec[tri] = r;
  int *pt = const_cast<int *>(b.prTriple) + b.trPair[tri]; /* owning triple of each primitive pair */
  for (int k = 0; k < r.np; k++) pt[k] = tri;
}

/* ---- link: gamma[p][q], one thread per element ---- */
__device__ __forceinline__ int deg_of_cindex(int p) {
  int l = 0;
  while (ecp_cd(l) <= p) l++;
  return l;
}
/* One launch per (la, lb, L) class, templated on NA = la + 1 and NB = lb + 1: for one l the admissible
 * lambda1 = ll1, ll1 + 2, ... <= la + l are at most la + 1 values (lambda2: lb + 1).  All their angular factors
 * sum_m Omega_A Omega_B are accumulated together in registers with m outermost, so that every Omega element is loaded
 * once per m instead of once per (lambda1, lambda2, m) and the index arithmetic is paid per m, not per multiply-add
 * (the round-1 kernel issued ~32 instructions per multiply-add).  Each factor is still the sum over m = 0..2l in that
 * order, multiplied by T and added in the reference's (lambda1, lambda2) order (src/type2.c:590-617): gamma is
 * bit-identical to the straightforward loop nest. */
template <int NA, int NB>
__global__ void __launch_bounds__(128) k_link(DevT t, DevB b, int c) {
  constexpr int la = NA - 1, lb = NB - 1;
  constexpr int cda = (la + 1) * (la + 2) * (la + 3) / 6, cdb = (lb + 1) * (lb + 2) * (lb + 3) / 6;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long w = b.clsElem[c] + idx;
  if (w >= b.clsElem[c + 1]) return;
  const int L = t.clsL[c];
  const int tri = b.clsFirst[c] + (int)(idx / (cda * cdb));
  const int pq = (int)(idx % (cda * cdb)), p = pq / cdb, q = pq % cdb;
  const int alpha = deg_of_cindex(p), beta = deg_of_cindex(q);
  const TriRec rec = b.trirec[tri];
  const int incA1 = rec.incA1, incA2 = L * L * incA1;
  const int incB1 = rec.incB1, incB2 = L * L * incB1;
  const double *oA = b.omX + rec.omA + p, *oB = b.omX + rec.omB + q;
  const double *T = b.T + b.clsWork[c] + (long long)(tri - b.clsFirst[c]) * t.clsNq[c];
  const int16_t *qi = t.qidx + t.clsQidxOff[c];
  const int d2 = lb + L, d3 = la + lb + 1;
  double g = 0.0;
  for (int l = 0; l < L; l++) {
    int ll1 = l - alpha, ll2 = l - beta;
    const int par1 = (alpha + l) % 2, par2 = (beta + l) % 2;
    ll1 = (par1 > ll1) ? par1 : ll1;
    ll2 = (par2 > ll2) ? par2 : ll2;
    const int n1 = (la + l - ll1) / 2 + 1, n2 = (lb + l - ll2) / 2 + 1; /* lambda1 = ll1 + 2i, lambda2 = ll2 + 2j */
    double f[NA][NB];
#pragma unroll
    for (int i = 0; i < NA; i++)
#pragma unroll
      for (int j = 0; j < NB; j++) f[i][j] = 0.0;
    const double *pa = oA + (size_t)ll1 * incA2 + (size_t)(l * l) * incA1;
    const double *pb = oB + (size_t)ll2 * incB2 + (size_t)(l * l) * incB1;
#pragma unroll 2
    for (int m = 0; m < 2 * l + 1; m++) {
      double a[NA], bb[NB];
#pragma unroll
      for (int i = 0; i < NA; i++) a[i] = (i < n1) ? pa[(size_t)(2 * i) * incA2] : 0.0;
#pragma unroll
      for (int j = 0; j < NB; j++) bb[j] = (j < n2) ? pb[(size_t)(2 * j) * incB2] : 0.0;
#pragma unroll
      for (int i = 0; i < NA; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) f[i][j] = fma(a[i], bb[j], f[i][j]); /* rows/columns beyond n1/n2 are never read */
      pa += incA1;
      pb += incB1;
    }
    double tmp = 0.0;
    const int16_t *ql = qi + ((l * (la + L) + ll1) * d2 + ll2) * d3 + alpha + beta;
#pragma unroll
    for (int i = 0; i < NA; i++)
#pragma unroll
      for (int j = 0; j < NB; j++)
        if (i < n1 && j < n2) {
          const int k = ql[(2 * i * d2 + 2 * j) * d3];
          if (k >= 0) tmp = fma(f[i][j], T[k], tmp); /* k < 0: angular factor identically zero */
        }
    g += tmp;
  }
  b.gamma[w] = g;
}
typedef void (*LinkKernel)(DevT, DevB, int);
#define LINK_ROW(A) {k_link<A, 1>, k_link<A, 2>, k_link<A, 3>, k_link<A, 4>, k_link<A, 5>, k_link<A, 6>}
static const LinkKernel g_linkKernels[ECP_MAX_LBS + 1][ECP_MAX_LBS + 1] = {LINK_ROW(1), LINK_ROW(2), LINK_ROW(3),
                                                                           LINK_ROW(4), LINK_ROW(5), LINK_ROW(6)};

/* ---- link, specialised shared-memory kernel for the large classes (default; LIBECP_B200_LINK=global keeps k_link) ----
 * ncu on k_link and on two shared-memory rewrites of it (profiles/r2: k_link2 of round 1, staging per triple, and
 * k_link3, staging per run of triples) showed the same picture: 66-71 % of the issue slots busy, 10 % of the executed
 * instructions DFMA - integer multiply-adds for strides that are only known at run time, predicates and zeroing of the
 * (lambda1, lambda2) tile, a division per staged element.  The link is instruction-issue bound.  k_link4 removes the
 * instructions instead of hiding their latency:
 *   - templated on (la+1, lb+1, L): every stride inside the staged slices is a compile-time constant, the l and m
 *     loops are fully unrolled, so the tile update is  NA + NB  LDS with immediate offsets and  NA * NB  DFMA;
 *   - no predication: rows beyond the admissible lambda are simply read (they lie inside the shared-memory slab; a
 *     tile entry they feed is never used because its qidx entry is < 0);
 *   - a thread owns one gamma element for the whole block; its per-l slice offsets and the NA * NB * L positions of
 *     its T values are computed once per block and kept in registers;
 *   - runs: the angular factors depend on the two atoms only, so they are formed once for the consecutive triples of
 *     the class that share both (k_link3's idea, kept), and only the slice that changed is staged again.
 * Per element the multiply-adds and their order are those of k_link (src/type2.c:583-623): gamma is bit-identical
 * (up to the sign of exact zeros: the first m term is a product instead of fma(a, b, +0)). */
#define LINK4_RMAX 4
#define LINK4_MAXL 5
/* one l of the link for one gamma element: the tile is as large as the admissible lambda counts of this l can get,
 * min(la, (la + l) / 2) + 1 by min(lb, (lb + l) / 2) + 1 (compile time) */
template <int NA, int NB, int L, int l>
__device__ __forceinline__ void link4_level(const double *__restrict__ pa, const double *__restrict__ pb,
                                            const double *__restrict__ Ts, const int (&kk)[NA][NB], int rl,
                                            double (&g)[LINK4_RMAX]) {
  constexpr int la = NA - 1, lb = NB - 1, L2 = L * L;
  constexpr int cda = (la + 1) * (la + 2) * (la + 3) / 6, cdb = (lb + 1) * (lb + 2) * (lb + 3) / 6;
  constexpr int incA1 = cda, incA2 = L2 * cda, incB1 = cdb, incB2 = L2 * cdb;
  constexpr int NAl = ((la + l) / 2 + 1 < NA) ? (la + l) / 2 + 1 : NA, NBl = ((lb + l) / 2 + 1 < NB) ? (lb + l) / 2 + 1 : NB;
  double f[NAl][NBl];
#pragma unroll
  for (int m = 0; m < 2 * l + 1; m++) {
    double a[NAl], bb[NBl];
#pragma unroll
    for (int i = 0; i < NAl; i++) a[i] = pa[(2 * i) * incA2 + m * incA1];
#pragma unroll
    for (int j = 0; j < NBl; j++) bb[j] = pb[(2 * j) * incB2 + m * incB1];
#pragma unroll
    for (int i = 0; i < NAl; i++)
#pragma unroll
      for (int j = 0; j < NBl; j++) f[i][j] = (m == 0) ? a[i] * bb[j] : fma(a[i], bb[j], f[i][j]);
  }
#pragma unroll
  for (int r = 0; r < LINK4_RMAX; r++)
    if (r < rl) {
      double tmp = 0.0;
#pragma unroll
      for (int i = 0; i < NAl; i++)
#pragma unroll
        for (int j = 0; j < NBl; j++) tmp = fma(f[i][j], Ts[kk[i][j] + r], tmp);
      g[r] += tmp;
    }
}
template <int NA, int NB, int L, int... Ls>
__device__ __forceinline__ void link4_levels(const double *A, const double *Bm, const double *Ts, const int (&offA)[L],
                                             const int (&offB)[L], const int (&kk)[L][NA][NB], int rl,
                                             double (&g)[LINK4_RMAX], std::integer_sequence<int, Ls...>) {
  (link4_level<NA, NB, L, Ls>(A + offA[Ls], Bm + offB[Ls], Ts, kk[Ls], rl, g), ...);
}
#define LINK4_BD(NA, NB) ((((NA) * ((NA) + 1) * ((NA) + 2) / 6) * ((NB) * ((NB) + 1) * ((NB) + 2) / 6) + 31) / 32 * 32)
/* 8-byte asynchronous global -> shared copy (LDGSTS): the slices of the NEXT run arrive while this run is computed */
__device__ __forceinline__ void cp_async8(double *smemDst, const double *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#define LINK4_TPB_MAX 32
template <int NA, int NB, int L>
__global__ void __launch_bounds__(LINK4_BD(NA, NB), (512 / LINK4_BD(NA, NB) > 0 ? 512 / LINK4_BD(NA, NB) : 1))
    k_link4(DevT t, DevB b, int c, int tpb, int tsz, int smTotal) {
  constexpr int la = NA - 1, lb = NB - 1, L2 = L * L;
  constexpr int cda = (la + 1) * (la + 2) * (la + 3) / 6, cdb = (lb + 1) * (lb + 2) * (lb + 3) / 6, E = cda * cdb;
  constexpr int SA = (la + L) * L2 * cda, SB = (lb + L) * L2 * cdb;
  constexpr int incA2 = L2 * cda, incB2 = L2 * cdb;
  constexpr int d2 = lb + L, d3 = la + lb + 1, BD = LINK4_BD(NA, NB); /* = blockDim.x */
  extern __shared__ __align__(16) double lk_sm[];
  /* A0 A1 | B0 B1 | T0 T1 ([quadrature][triple of the run] + one row of zeros each) | slack.  Two copies of each: the
   * slices of run k+1 are fetched with cp.async while run k is computed.  Rows beyond the admissible lambda are read
   * without a predicate: from A they fall into the next slice, from B into T and the slack, which hold finite numbers
   * (everything is cleared once); the tile entries they feed meet the zero row of T. */
  double *Abuf = lk_sm, *Bbuf = Abuf + 2 * SA, *Tbuf = Bbuf + 2 * SB;
  __shared__ long long mOmA[LINK4_TPB_MAX], mOmB[LINK4_TPB_MAX];
  __shared__ int mIncA[LINK4_TPB_MAX], mIncB[LINK4_TPB_MAX];
  const int nq = t.clsNq[c];
  const int first = b.clsFirst[c], nTri = b.clsFirst[c + 1] - first;
  const int lt0 = blockIdx.x * tpb, nT = min(tpb, nTri - lt0); /* triples of this block */
  if ((int)threadIdx.x < nT) {
    const TriRec *r = &b.trirec[first + lt0 + threadIdx.x];
    mOmA[threadIdx.x] = r->omA;
    mOmB[threadIdx.x] = r->omB;
    mIncA[threadIdx.x] = r->incA1;
    mIncB[threadIdx.x] = r->incB1;
  }
  const int e = threadIdx.x;
  const bool active = e < E;
  const int p = active ? e / cdb : 0, q = active ? e - p * cdb : 0;
  const int alpha = deg_of_cindex(p), beta = deg_of_cindex(q);
  int offA[L], offB[L], kk[L][NA][NB]; /* kk: offset of the quadrature's T row (the zero row if the factor vanishes) */
  {
    const int16_t *qi = t.qidx + t.clsQidxOff[c];
#pragma unroll
    for (int l = 0; l < L; l++) {
      int ll1 = l - alpha, ll2 = l - beta;
      const int par1 = (alpha + l) % 2, par2 = (beta + l) % 2;
      ll1 = (par1 > ll1) ? par1 : ll1;
      ll2 = (par2 > ll2) ? par2 : ll2;
      const int n1 = (la + l - ll1) / 2 + 1, n2 = (lb + l - ll2) / 2 + 1; /* lambda1 = ll1 + 2i, lambda2 = ll2 + 2j */
      offA[l] = p + ll1 * incA2 + (l * l) * cda;
      offB[l] = q + ll2 * incB2 + (l * l) * cdb;
      const int16_t *ql = qi + ((l * (la + L) + ll1) * d2 + ll2) * d3 + alpha + beta;
#pragma unroll
      for (int i = 0; i < NA; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) {
          int k = (i < n1 && j < n2) ? (int)ql[(2 * i * d2 + 2 * j) * d3] : -1;
          kk[l][i][j] = (k >= 0 ? k : nq) * LINK4_RMAX;
        }
    }
  }
  for (int i = threadIdx.x; i < smTotal; i += BD) lk_sm[i] = 0.0; /* unpredicated reads only ever see finite numbers */
  constexpr int RPA = BD / cda, RPB = BD / cdb; /* rows of a slice one sweep of the block copies */
  const int rowA = threadIdx.x / cda, colA = threadIdx.x - rowA * cda, rowB = threadIdx.x / cdb, colB = threadIdx.x - rowB * cdb;
  __syncthreads();
  /* issue the copies of the run that starts at block-local triple `s`; returns its length.  ia / ib: slice copies in use */
  int ia = 0, ib = 0;
  long long curA = -1, curB = -1;
  auto fetch = [&](int s, int tb, int &ja, int &jb, long long &nA, long long &nB) -> int {
    const long long oA = mOmA[s], oB = mOmB[s];
    int rl = 1;
    while (rl < LINK4_RMAX && s + rl < nT && mOmA[s + rl] == oA && mOmB[s + rl] == oB) rl++;
    if (oA != nA) { /* rows (lambda, (l,m)) of Omega_A, cut to the cda columns of the shell: RPA rows per sweep */
      ja ^= 1;
      nA = oA;
      if (rowA < RPA) {
        const double *g = b.omX + oA + rowA * mIncA[s] + colA;
        double *dst = Abuf + ja * SA + rowA * cda + colA;
        const int stepG = RPA * mIncA[s];
#pragma unroll 4
        for (int r = rowA; r < (la + L) * L2; r += RPA, g += stepG, dst += RPA * cda) cp_async8(dst, g);
      }
    }
    if (oB != nB) {
      jb ^= 1;
      nB = oB;
      if (rowB < RPB) {
        const double *g = b.omX + oB + rowB * mIncB[s] + colB;
        double *dst = Bbuf + jb * SB + rowB * cdb + colB;
        const int stepG = RPB * mIncB[s];
#pragma unroll 4
        for (int r = rowB; r < (lb + L) * L2; r += RPB, g += stepG, dst += RPB * cdb) cp_async8(dst, g);
      }
    }
    const double *gT = b.T + b.clsWork[c] + (long long)(lt0 + s) * nq; /* T of consecutive triples is contiguous */
    double *Ts = Tbuf + tb * tsz;
    for (int r = 0; r < rl; r++)
      for (int i = threadIdx.x; i < nq; i += BD) cp_async8(Ts + i * LINK4_RMAX + r, gT + r * nq + i);
    return rl;
  };
  int s = 0, tb = 0;
  int rl = nT > 0 ? fetch(0, 0, ia, ib, curA, curB) : 0;
  while (s < nT) {
    cp_async_wait_all();
    __syncthreads(); /* run s has arrived; every thread is through with the run before it (its slices may be overwritten) */
    const double *A = Abuf + ia * SA, *Bm = Bbuf + ib * SB, *Ts = Tbuf + tb * tsz;
    const int sNext = s + rl;
    int rlNext = 0;
    if (sNext < nT) rlNext = fetch(sNext, tb ^ 1, ia, ib, curA, curB); /* ia / ib now name the NEXT run's slices */
    if (active) {
      double g[LINK4_RMAX];
#pragma unroll
      for (int r = 0; r < LINK4_RMAX; r++) g[r] = 0.0;
      link4_levels<NA, NB, L>(A, Bm, Ts, offA, offB, kk, rl, g, std::make_integer_sequence<int, L>{});
      double *out = b.gamma + b.clsElem[c] + (long long)(lt0 + s) * E + e;
#pragma unroll
      for (int r = 0; r < LINK4_RMAX; r++)
        if (r < rl) out[(size_t)r * E] = g[r];
    }
    s = sNext;
    rl = rlNext;
    tb ^= 1;
  }
}
typedef void (*Link4Kernel)(DevT, DevB, int, int, int, int);
/* classes with at least 36 gamma elements per triple among the s-f shells, L = 1..5; everything else keeps k_link */
#define LINK4_L(A, B) {NULL, k_link4<A, B, 1>, k_link4<A, B, 2>, k_link4<A, B, 3>, k_link4<A, B, 4>, k_link4<A, B, 5>}
#define LINK4_NONE {NULL, NULL, NULL, NULL, NULL, NULL}
static const Link4Kernel g_link4Kernels[4][4][LINK4_MAXL + 1] = {
    {LINK4_NONE, LINK4_NONE, LINK4_NONE, LINK4_NONE},
    {LINK4_NONE, LINK4_NONE, LINK4_L(2, 3), LINK4_L(2, 4)},
    {LINK4_NONE, LINK4_L(3, 2), LINK4_L(3, 3), LINK4_L(3, 4)},
    {LINK4_NONE, LINK4_L(4, 2), LINK4_L(4, 3), LINK4_L(4, 4)}};

#include "ecp_type1.cuh"

/* ---- type 1, per primitive pair: P = 2(za r_AC + zb r_BC), |P|, S_lm(P^), and the pair record of the radial kernels ---- */
__global__ void k_t1prep(DevT t, DevB b) {
  const long long pr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pr >= b.nPairs) return;
  const int tri = b.prTriple[pr];
  const int ssa = b.trA[tri], ssb = b.trB[tri];
  const int sha = b.ssShell[ssa], shb = b.ssShell[ssb];
  const int asa = b.ssASlot[ssa], asb = b.ssASlot[ssb];
  const int Nb = t.shellK[shb];
  const int ip = (int)(pr - b.trPair[tri]), pa = ip / Nb, pb = ip % Nb;
  const double za = t.primA[t.shellPrim[sha] + pa], zb = t.primA[t.shellPrim[shb] + pb];
  const double ca = t.primD[t.shellPrim[sha] + pa], cb = t.primD[t.shellPrim[shb] + pb];
  const double *rA = b.asR + 4 * asa, *rB = b.asR + 4 * asb;
  const double Px = 2.0 * (za * rA[0] + zb * rB[0]);
  const double Py = 2.0 * (za * rA[1] + zb * rB[1]);
  const double Pz = 2.0 * (za * rA[2] + zb * rB[2]);
  double r, th, ph;
  ecp_sphcoord(Px, Py, Pz, &r, &th, &ph);
  const int lab = t.shellL[sha] + t.shellL[shb];
  const long long qoff = pair_Q_off(t, b, find_class(b.clsPairBase, t.nClasses, pr), pr);
  switch (lab) { /* pairs are class-sorted: (almost) uniform over a warp */
    case 0: ecp_rsh_t<0>(th, ph, t.fac, t.dfac, b.rshP + qoff); break;
    case 1: ecp_rsh_t<1>(th, ph, t.fac, t.dfac, b.rshP + qoff); break;
    case 2: ecp_rsh_t<2>(th, ph, t.fac, t.dfac, b.rshP + qoff); break;
    case 3: ecp_rsh_t<3>(th, ph, t.fac, t.dfac, b.rshP + qoff); break;
    case 4: ecp_rsh_t<4>(th, ph, t.fac, t.dfac, b.rshP + qoff); break;
    case 5: ecp_rsh_t<5>(th, ph, t.fac, t.dfac, b.rshP + qoff); break;
    case 6: ecp_rsh_t<6>(th, ph, t.fac, t.dfac, b.rshP + qoff); break;
    default: ecp_rsh(lab, th, ph, t.fac, t.dfac, b.rshP + qoff); break;
  }
  b.sP[pr] = r;
  const double dAC = rA[3], dBC = rB[3];
  T1Rec rec;
  rec.z = -za - zb;
  rec.sS = r;
  rec.zd2 = -za * dAC * dAC - zb * dBC * dBC; /* src/type1.c:103 */
  rec.CcS = ca * cb * exp(rec.zd2);           /* src/type1.c:113 */
  rec.CcL = ca * cb;                          /* src/type1.c:151 */
  const double zp = za + zb;
  ecp_fm06_map(zp, (za * dAC + zb * dBC) / zp, &rec.i1, &rec.i2);
  rec.qoff = qoff;
  rec.type = b.asType[asa];
  rec.gs = max(b.ssStart[ssa], b.ssStart[ssb]); /* src/libecp.c:315-316 */
  rec.ge = max(b.ssEnd[ssa], b.ssEnd[ssb]);
  rec.pad = 0;
  b.t1rec[pr] = rec;
}

/* ---- chi[i][j]: eight lanes per triple, two phases ----
 * The reference forms, per primitive pair, chi[i][j] += sum_{l = N, N-2, ...} (sum_m S_lm(P^) poly2sph[p][(l,m)]) Q[N][l]
 * with N the degree and p the index of the product monomial x^i y^j z^k of the element (src/type1.c:266-295): the
 * pair-dependent factors S_lm Q[N][l] do not depend on the element beyond N.  So the sum over the primitive pairs is
 * taken first,
 *     R[N][(l,m)] = sum_pairs S_lm(P^_pair) Q_pair[N][l]          (l = N, N-2, ...: C_DIM(la+lb) values per triple)
 * by the 8 lanes of the triple into shared memory, and every element is one short dot product
 *     chi[i][j] = sum_{l,m} poly2sph[p][(l,m)] R[N][(l,m)].
 * The multiply-adds per triple drop from  pairs x elements x (N+1)(N+2)/2  to  (pairs + elements) x (N+1)(N+2)/2
 * (p-p shells of the TZ sets have 16 pairs).  Same products as the reference, associated over the pairs first.
 * Measured and rejected in round 2 (profiles/r2/README.md): a class-uniform block version with the index maps and the
 * poly2sph rows staged in shared memory - its per-block tables cost more than the index decoding they replace
 * (0.19 -> 0.26 / 0.65 ms on Au20 with 8 / 128 triples per block, 8.2 -> 19.0 / 15.5 ms per config-5 pass). */
__device__ __forceinline__ int chi_roff(int N, int l) { /* position of (l, m = 0) of level N in the packed R */
  return N * (N + 1) * (N + 2) / 6 + l * (l - 1) / 2;
}
__global__ void __launch_bounds__(128) k_chi(DevT t, DevB b, int rStride) {
  extern __shared__ __align__(16) double chi_R[];
  const int gl = threadIdx.x & 7;
  double *R = chi_R + (size_t)(threadIdx.x >> 3) * rStride;
  const int tri = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3);
  const bool valid = tri < b.nTriples;
  int c = 0, la = 0, lb = 0, lab = 0, np = 0;
  long long q0 = 0;
  if (valid) {
    c = find_class_i(b.clsFirst, t.nClasses, tri);
    la = t.clsLa[c];
    lb = t.clsLb[c];
    lab = la + lb;
    np = b.trirec[tri].np;
    q0 = b.trirec[tri].q0;
  }
  const int ld = (lab + 1) * (lab + 1);
  /* phase 1: R[N][(l,m)] */
  if (valid) {
    const int nR = ecp_cd(lab);
    for (int e = gl; e < nR; e += 8) {
      int N = 0;
      while (ecp_cd(N) <= e) N++;
      int rem = e - N * (N + 1) * (N + 2) / 6, l = N & 1;
      while (rem >= 2 * l + 1) {
        rem -= 2 * l + 1;
        l += 2;
      }
      const double *rsh = b.rshP + q0 + l * l + rem;
      const double *Q = b.Q + q0 + N * (lab + 1) + l;
      double acc = 0.0;
      for (int ip = 0; ip < np; ip++) acc = fma(rsh[(size_t)ip * ld], Q[(size_t)ip * ld], acc);
      R[e] = acc;
    }
  }
  __syncwarp();
  /* phase 2: the elements of the triple, 8 at a time */
  if (valid) {
    const int cdb = ecp_cd(lb), ne = ecp_cd(la) * cdb;
    double *out = b.chi + b.clsElem[c] + (long long)(tri - b.clsFirst[c]) * ne;
    const int D = t.ijkDim;
    for (int pq = gl; pq < ne; pq += 8) {
      const int i = pq / cdb, j = pq - i * cdb;
      const int *ei = t.ijk + 3 * i, *ej = t.ijk + 3 * j;
      const int lx = ei[0] + ej[0], ly = ei[1] + ej[1], lz = ei[2] + ej[2], N = lx + ly + lz;
      (void)N;
      /* only the non-zero poly2sph entries of the monomial (a zero entry adds an exact zero: same value); the sums over
       * m of one l are closed into chi at the flagged entry, as the dense loop nest did (src/type1.c:283-291) */
      const int mono = t.ijkIndex[lx * D * D + ly * D + lz];
      const int k1 = t.p2sOff[mono + 1];
      double chi = 0.0, f = 0.0;
      for (int k = t.p2sOff[mono]; k < k1; k++) {
        const unsigned ix = t.p2sIdx[k];
        f = fma(t.p2sVal[k], R[ix & 0x7fffu], f);
        if (ix & 0x8000u) {
          chi += f;
          f = 0.0;
        }
      }
      out[pq] = chi;
    }
  }
}

/* ---- shift to A/B-centred Cartesians, normalise, write blocks / accumulate matrix ---- */
#include "ecp_shift.cuh"

/* gather M[i][i..n) of the listed rows back to back (one block per row) for a contiguous D2H */
__global__ void k_pack_rows(const double *__restrict__ M, int n, const int *__restrict__ rows,
                            const long long *__restrict__ off, long long base, double *__restrict__ out) {
  const int i = rows[blockIdx.x];
  double *o = out + (off[blockIdx.x] - base) - i;
  const double *src = M + (size_t)i * n;
  for (int j = i + threadIdx.x; j < n; j += blockDim.x) o[j] = src[j];
}

/* inverse of k_pack_rows: packed upper-triangle rows (of another rank's shard) into the full matrix */
__global__ void k_unpack_rows(double *__restrict__ M, int n, const int *__restrict__ rows, const long long *__restrict__ off,
                              const double *__restrict__ in) {
  const int i = rows[blockIdx.x];
  const double *src = in + off[blockIdx.x] - i;
  double *dst = M + (size_t)i * n;
  for (int j = i + threadIdx.x; j < n; j += blockDim.x) dst[j] = src[j];
}

/* ---- sparse download of the result matrix (host consumer, src/getIntegrals.c:36-42) ----
 * The ECP matrix is block sparse: a block is non-zero only if both shells reach a common ECP centre (configuration 5:
 * 15 % of the upper triangle, 24 % of its 128-byte lines).  Instead of the dense upper triangle the host consumer moves
 * only the RUNS of SPR_RUN consecutive elements of a row M[i][i..n) that hold a non-zero, plus one bit per run.
 * k_rows_flag: bitmap and number of such runs per listed row; the host turns the counts into row offsets;
 * k_rows_pack_sparse: the flagged runs of every row back to back (a short last run of a row is padded with zeros). */
#define SPR_RUN 16
#define SPR_PF 6 /* host add: runs prefetched ahead */
__global__ void k_rows_flag(const double *__restrict__ M, int n, const int *__restrict__ rows, int words,
                            unsigned *__restrict__ bitmap, int *__restrict__ count) {
  extern __shared__ unsigned spr_bm[];
  __shared__ int total;
  const int i = rows[blockIdx.x];
  const double *src = M + (size_t)i * n + i;
  const int len = n - i, nruns = (len + SPR_RUN - 1) / SPR_RUN;
  for (int w = threadIdx.x; w < words; w += blockDim.x) spr_bm[w] = 0;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const int hw = threadIdx.x >> 4, nhw = blockDim.x >> 4, l16 = threadIdx.x & 15;
  for (int r0 = 0; r0 < nruns; r0 += nhw) { /* half a warp per run; every thread makes the same number of trips */
    const int r = r0 + hw, j = r * SPR_RUN + l16;
    const bool nz = (r < nruns && j < len) ? (src[j] != 0.0) : false;
    const unsigned b = __ballot_sync(0xffffffffu, nz);
    const unsigned half = (threadIdx.x & 16) ? (b >> 16) : (b & 0xffffu);
    if (l16 == 0 && half) atomicOr(&spr_bm[r >> 5], 1u << (r & 31));
  }
  __syncthreads();
  int c = 0;
  for (int w = threadIdx.x; w < words; w += blockDim.x) {
    const unsigned v = spr_bm[w];
    bitmap[(size_t)blockIdx.x * words + w] = v;
    c += __popc(v);
  }
  if (c) atomicAdd(&total, c);
  __syncthreads();
  if (threadIdx.x == 0) count[blockIdx.x] = total;
}
__global__ void k_rows_pack_sparse(const double *__restrict__ M, int n, const int *__restrict__ rows, int words,
                                   const unsigned *__restrict__ bitmap, const long long *__restrict__ rowBase,
                                   double *__restrict__ out) {
  extern __shared__ unsigned spr_bm[]; /* [words] bitmap, [words] runs before each word */
  unsigned *pre = spr_bm + words;
  const int i = rows[blockIdx.x];
  const double *src = M + (size_t)i * n + i;
  const int len = n - i, nruns = (len + SPR_RUN - 1) / SPR_RUN;
  for (int w = threadIdx.x; w < words; w += blockDim.x) spr_bm[w] = bitmap[(size_t)blockIdx.x * words + w];
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    for (int w = 0; w < words; w++) {
      pre[w] = acc;
      acc += __popc(spr_bm[w]);
    }
  }
  __syncthreads();
  double *o = out + rowBase[blockIdx.x] * SPR_RUN;
  const int hw = threadIdx.x >> 4, nhw = blockDim.x >> 4, l16 = threadIdx.x & 15;
  for (int r = hw; r < nruns; r += nhw) {
    const unsigned word = spr_bm[r >> 5], bit = 1u << (r & 31);
    if (!(word & bit)) continue;
    const int pos = pre[r >> 5] + __popc(word & (bit - 1)), j = r * SPR_RUN + l16;
    o[(size_t)pos * SPR_RUN + l16] = (j < len) ? src[j] : 0.0;
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* FP64 FMA peak probe (roofline denominator when no measured FP64 peak is published) */
__global__ void k_fp64_probe(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

/* ================================================================================================ */
/* host side of the device layer */
struct Buf {
  void *p;
  size_t cap;
};
struct EcpDev {
  int device, nSM;
  cudaStream_t s1, s2, s2real; /* s2 = type-1 stream; aliases s1 in serial (profiling) mode */
  int serial;
  cudaEvent_t ev[12];
  DevT t;
  DevB b;
  int nClasses, maxQPerL, nAO, maxLBS;
  Buf tab[64];
  int ntab;
  /* batch inputs, two sets: the arrays of batch i+1 are copied on the copy stream s3 (from the builder thread, straight
   * out of its page-locked output) while the kernels of batch i run */
  struct UpSet {
    Buf asAtom, asType, asR, asOmOff, ssShell, ssASlot, ssStart, ssEnd, ssFOff, trA, trB, trOut, trPair;
    Buf prTriple, clsFirst, clsWork, clsElem, clsOutElem, clsPairBase, clsQBase;
    /* device enumeration (ecp_enum.cuh): per-centre layout, counters, and the totals / class prefixes read back */
    Buf ceAS0, cePair0, asSS0, ssOwn, ccTri, ccPair, pairCnt, meta;
    EnumIn ein;
    size_t enumSmem;
    /* per-centre tables of the batch in the set (rsh / monomials per atom slot, Omega_X, F): they only need the slot arrays,
     * so the set that is prefetched behind the running batch also computes them there (tablesDone) */
    Buf rshX, uspX, omX, F;
    int tablesDone;
  } up[2];
  void *enumHost[2]; /* page-locked: EnumMeta, clsFirst (as long long), clsWork, clsElem, clsOutElem, clsPairBase, clsQBase */
  int enumHost_nc;
  int enumOff; /* LIBECP_B200_ENUM=host: triples from the host builder also in matrix runs */
  cudaStream_t s3;
  cudaStream_t s4; /* D2H of finished row panels while the next panel computes (ecpdev_matrix_add_to_host, async = 1) */
  cudaEvent_t evUp[2], evTab[2];
  int tabPrefetch; /* LIBECP_B200_TABPREFETCH=1: tables of the prefetched batch on the upload stream (default: with the batch) */
  cudaEvent_t evDone; /* blocking-sync event: the driving thread sleeps while a batch runs (its core goes to the builder) */
  const EcpBatch *upBatch[2]; /* batch whose arrays sit in the set (NULL: none) */
  long long upBytes[2];
  Buf rshX, uspX, omX, F, T, gamma, chi, Q, rshP, sP, blocks, tfail, tflags, items, counters;
  /* spherical-harmonic output (ecpdev_spherical) */
  Buf sphMatrix;
  int nSph;
  const int *sphShell, *shellSph, *c2sOff;
  const double *c2s;
  double *matrix;
  size_t matrixBytes;
  int matrixKnown, dirtyAll, nDirty; /* matrix is zero outside the upper-triangle parts of the dirty rows */
  long long dirtySig;
  Buf dirtyRows, gatherRows, gatherOff;
  /* C-ABI collective (ecpdev_comm_*, ecpdev_allgather): NCCL communicator, shard layout of all ranks, staging */
  void *comm;
  int commOwned, commRank, commWorld;
  Buf agRows, agOff, agBuf;
  Buf dlRows, dlBits, dlCount, dlBase, dlPay; /* sparse download of the host consumer (matrix_add_to_host_sparse); parked with the scratch */
  long long agCap, *agCount, *agFirst; /* per rank: packed doubles, first entry of its rows in agRows/agOff */
  size_t lastSizes[8];
  long long tableBytes, batchH2D;
  int hClsLa[ECP_MAX_CLASSES], hClsLb[ECP_MAX_CLASSES], hClsL[ECP_MAX_CLASSES], hClsNq[ECP_MAX_CLASSES];
  Buf fastSurv;
  int ftabCompact; /* 1 (default): k_Ftab2, only the window of a shell slot is tabulated; LIBECP_B200_FTAB=full: k_Ftab */
  int shift2, shift2Attr; /* 1 (default): k_shift2, both passes in one kernel; LIBECP_B200_SHIFT=two: k_shiftJ + k_shiftI */
  int hShTerms[ECP_MAX_LBS + 1]; /* binomial-shift terms per shell angular momentum */
  int linkMode; /* LIBECP_B200_LINK: 4 (default) specialised shared-memory kernel k_link4 for the large classes; global = k_link
                 * everywhere */
  int linkTpb;  /* LIBECP_B200_LINKTPB: consecutive triples per block of k_link4 (0 = sized to the class) */
  int fastLim; /* levels of the first fast-path launch (LIBECP_B200_FASTLIM, default 4 = 15 points) */
  int fastUnroll; /* LIBECP_B200_FASTUNROLL: point pairs of the survivors' loop whose loads are in flight together (1, 2, 4) */
  long long survCapEnv; /* LIBECP_B200_SURVCAP: survivor-list capacity override (tests of the overflow path) */
  Buf t1list, t1mask, t1count, t1work, t1rec, trirec, clsJ, Jbuf, fbItems, fbList, fbUnits, fbTotals, fbR;
  Buf t1surv, t1smask, t1scount, t1state; /* k_type1A -> k_type1S: survivor list, open masks, per-launch count, (I, p, q) */
  int t1legacy;                           /* LIBECP_B200_T1=legacy: k_type1S alone, from the first point */
  int launchSeq;
  Buf dbgBuf;
  int tails; /* LIBECP_B200_TAILS */
  int fbWaves; /* 1 (default): level waves (ecp_waves.cuh); LIBECP_B200_FB=group: k_fallbackG, persistent 8-lane groups */
  Buf fbwItems, fbwUnits, fbwQd, fbwSI, fbwSP, fbwSQ, fbwRes, fbwOpenFlag, fbwVals, fbwListA, fbwListB, fbwCtr;
  int fbblock, fbocc, fbminb; /* tuning knobs of the fallback kernel: LIBECP_B200_FBBLOCK threads, _FBOCC blocks per SM cap, _FBMINB */
  int t1block;                /* LIBECP_B200_T1BLOCK = 32/64/96/128 threads per block of the type-1 kernels */
};

/* scratch buffers come from the device's stream-ordered pool (release threshold raised in ecpdev_create),
 * so creating / destroying handles back to back does not pay cudaMalloc / cudaFree of gigabytes each time */
static thread_local cudaStream_t g_allocStream = 0;
static int ensure(Buf *b, size_t bytes) {
  if (bytes <= b->cap && b->p) return 0;
  if (b->p) cudaFreeAsync(b->p, g_allocStream);
  b->p = NULL;
  b->cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  CK(cudaMallocAsync(&b->p, want, g_allocStream));
  b->cap = want;
  return 0;
}
template <typename T>
static const T *upload_const(EcpDev *d, const T *h, size_t n) {
  Buf *bf = &d->tab[d->ntab++];
  bf->p = NULL;
  bf->cap = 0;
  size_t bytes = (n ? n : 1) * sizeof(T);
  if (cudaMallocAsync(&bf->p, bytes, d->s1) != cudaSuccess) return NULL;
  bf->cap = bytes;
  d->tableBytes += (long long)(n * sizeof(T));
  if (n && cudaMemcpyAsync(bf->p, h, n * sizeof(T), cudaMemcpyHostToDevice, d->s1) != cudaSuccess) return NULL;
  return (const T *)bf->p;
}

/* Page-locked host memory for the batch arrays the builder fills (H2D straight from the builder's output at PCIe
 * speed instead of through the driver's pageable staging).  Pinning is expensive (~0.3 ms per MB), so blocks are kept
 * in a small process-wide cache and handed out again to later batches / handles.  Without a usable device (CPU test
 * tier: tables-only handles) the blocks are plain malloc memory. */
#include <mutex>
struct PinBlock {
  void *p;
  size_t cap;
  int inUse, pinned;
};
static PinBlock g_pin[128];
static std::mutex g_pinMu;
extern "C" void *ecpdev_pinned_alloc(size_t bytes) {
  std::lock_guard<std::mutex> lk(g_pinMu);
  int best = -1;
  for (int i = 0; i < 128; i++)
    if (g_pin[i].p && !g_pin[i].inUse && g_pin[i].cap >= bytes && g_pin[i].cap <= 2 * bytes + (1 << 20) &&
        (best < 0 || g_pin[i].cap < g_pin[best].cap))
      best = i;
  if (best >= 0) {
    g_pin[best].inUse = 1;
    return g_pin[best].p;
  }
  int slot = -1;
  for (int i = 0; i < 128 && slot < 0; i++)
    if (!g_pin[i].p) slot = i;
  if (slot < 0) /* cache full: evict an idle block */
    for (int i = 0; i < 128 && slot < 0; i++)
      if (!g_pin[i].inUse) {
        if (g_pin[i].pinned) cudaFreeHost(g_pin[i].p); else free(g_pin[i].p);
        g_pin[i].p = NULL;
        slot = i;
      }
  void *p = NULL;
  int pinned = 0;
  if (bytes >= (64 << 10) && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) {
    pinned = 1;
  } else {
    cudaGetLastError(); /* no device / out of lockable memory: pageable memory works too */
    p = malloc(bytes ? bytes : 1);
  }
  if (slot >= 0) {
    g_pin[slot].p = p;
    g_pin[slot].cap = bytes;
    g_pin[slot].inUse = 1;
    g_pin[slot].pinned = pinned;
  }
  return p; /* (a block that found no registry slot is pageable-or-pinned but untracked: freed as malloc'd below) */
}
extern "C" void ecpdev_pinned_free(void *p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pinMu);
  for (int i = 0; i < 128; i++)
    if (g_pin[i].p == p) {
      g_pin[i].inUse = 0; /* stays cached */
      return;
    }
  free(p);
}

extern "C" const char *ecpdev_last_error(void) { return g_err; }

static void adopt_cached(EcpDev *d);
static void comm_release(EcpDev *d);

extern "C" EcpDev *ecpdev_create(const EcpHostTables *h, int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    snprintf(g_err, sizeof(g_err), "libecp_b200: no CUDA device available (this library has no CPU path)");
    return NULL;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "cudaSetDevice(%d) failed", device);
    return NULL;
  }
  EcpDev *d = (EcpDev *)calloc(1, sizeof(EcpDev));
  d->device = device;
  cudaDeviceGetAttribute(&d->nSM, cudaDevAttrMultiProcessorCount, device);
  cudaStreamCreateWithFlags(&d->s1, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&d->s2real, cudaStreamNonBlocking);
  { /* the upload stream also counts / scans the next batch's triples (ecp_enum.cuh): highest priority, so that those small
     * kernels run between the blocks of the current batch instead of behind them */
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&d->s3, cudaStreamNonBlocking, hi) != cudaSuccess) cudaStreamCreateWithFlags(&d->s3, cudaStreamNonBlocking);
  }
  cudaStreamCreateWithFlags(&d->s4, cudaStreamNonBlocking);
  for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&d->evUp[i], cudaEventDisableTiming);
  for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&d->evTab[i], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&d->evDone, cudaEventDisableTiming | cudaEventBlockingSync);
  d->s2 = d->s2real;
  d->serial = getenv("LIBECP_B200_SERIAL") != NULL;
  d->tails = getenv("LIBECP_B200_TAILS") != NULL;
  {
    const char *e = getenv("LIBECP_B200_FASTLIM");
    d->fastLim = e ? atoi(e) : 4;
    {
      const char *lk = getenv("LIBECP_B200_LINK");
      d->linkMode = (lk && !strcmp(lk, "global")) ? 1 : 4;
      {
        const char *tp = getenv("LIBECP_B200_LINKTPB");
        d->linkTpb = tp ? atoi(tp) : 0;
      }
      lk = getenv("LIBECP_B200_SHIFT");
      d->shift2 = !(lk && !strcmp(lk, "two"));
      lk = getenv("LIBECP_B200_FTAB");
      d->ftabCompact = !(lk && !strcmp(lk, "full"));
    }
    if (d->fastLim < 1) d->fastLim = 1;
    e = getenv("LIBECP_B200_TABPREFETCH");
    d->tabPrefetch = e && !strcmp(e, "1"); /* measured: no gain at 1 or 8 GPUs (the table kernels take the same share of
                                             * the device beside the running batch) - off unless asked for */
    e = getenv("LIBECP_B200_ENUM");
    d->enumOff = e && !strcmp(e, "host");
    e = getenv("LIBECP_B200_FASTUNROLL");
    d->fastUnroll = e ? atoi(e) : 1;
    e = getenv("LIBECP_B200_SURVCAP");
    d->survCapEnv = e ? atoll(e) : 0;
    e = getenv("LIBECP_B200_FBBLOCK");
    d->fbblock = e ? atoi(e) : 64;
    if (d->fbblock != 32 && d->fbblock != 64 && d->fbblock != 128) d->fbblock = 64;
    e = getenv("LIBECP_B200_FB");
    d->fbWaves = !(e && !strcmp(e, "group"));
    e = getenv("LIBECP_B200_FBMINB");
    d->fbminb = e ? atoi(e) : 3;
    e = getenv("LIBECP_B200_FBOCC");
    d->fbocc = e ? atoi(e) : 0;
    e = getenv("LIBECP_B200_T1");
    d->t1legacy = (e && !strcmp(e, "legacy"));
    t1_upload_qoff();
    e = getenv("LIBECP_B200_T1BLOCK");
    d->t1block = e ? atoi(e) : 64;
    if (d->t1block != 32 && d->t1block != 64 && d->t1block != 96 && d->t1block != 128) d->t1block = 64;
  }
  { /* keep freed scratch in the pool instead of returning it to the driver */
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  for (int i = 0; i < 12; i++) cudaEventCreate(&d->ev[i]);
  DevT &t = d->t;
  t.maxLECP = h->maxLECP; t.maxLBS = h->maxLBS; t.maxLambda = h->maxLambda; t.tmDim = h->tmDim; t.ijkDim = h->ijkDim;
  t.besselStride = h->besselStride; t.nU = h->nU; t.pcols = (h->tmDim + 1) * (h->tmDim + 1);
  t.largeSlots = h->largeSlots; t.largeLevels = h->largeLevels; t.largeOrder = h->largeOrder; t.nClasses = h->nClasses;
  t.omD2 = (h->maxLambda + 1) * (h->maxLambda + 1);
  t.omD3 = (h->maxAlpha + 1) * (h->maxAlpha + 2) * (h->maxAlpha + 3) / 6;
  t.tolerance = h->tolerance; t.accuracy = h->accuracy; t.lnAcc1 = h->lnAccuracy1; t.lnAcc2 = h->lnAccuracy2;
  for (int i = 0; i < ECP_SMALL_LEVELS; i++) {
    t.sm.levPairs[i] = h->small_levPairs[i];
    t.sm.levJ[i] = h->small_levJ[i];
    t.sm.levN[i] = h->small_levN[i];
    t.sm.levSlot[i] = h->small_levSlot[i];
  }
  t.sm.levSlot[ECP_SMALL_LEVELS] = h->small_levSlot[ECP_SMALL_LEVELS];
  ecp_small_meta_bounds(&t.sm, h->small_oidx);
  const int cdT = (h->tmDim + 1) * (h->tmDim + 2) * (h->tmDim + 3) / 6;
  t.fac = upload_const(d, h->fac, h->nfac);
  t.dfac = upload_const(d, h->dfac, h->nfac);
  t.poly2sph = upload_const(d, h->poly2sph, (size_t)cdT * t.pcols);
  t.omega = upload_const(d, h->omega, h->nomega);
  t.binom = upload_const(d, h->binom, (size_t)(h->maxLBS + 1) * (h->maxLBS + 1));
  t.shOffStride = h->shOffStride;
  t.shTermOff = upload_const(d, h->shTermOff, (size_t)(h->maxLBS + 1) * h->shOffStride);
  t.shTermP = upload_const(d, h->shTermP, h->nShTerms);
  t.shTermD = upload_const(d, h->shTermD, h->nShTerms);
  t.shTermBin = upload_const(d, h->shTermBin, h->nShTerms);
  { /* sparse form of poly2sph for k_chi: a monomial x^a y^b z^c only has components of matching parities */
    const int pc = t.pcols;
    int *off = (int *)malloc((size_t)(cdT + 1) * sizeof(int));
    unsigned short *idx = (unsigned short *)malloc((size_t)cdT * pc * sizeof(unsigned short));
    double *val = (double *)malloc((size_t)cdT * pc * sizeof(double));
    int n = 0;
    for (int p = 0; p < cdT; p++) {
      const int N = h->ijk[3 * p] + h->ijk[3 * p + 1] + h->ijk[3 * p + 2];
      off[p] = n;
      for (int l = N; l >= 0; l -= 2) {
        int last = -1;
        for (int m = 0; m < 2 * l + 1; m++) {
          const double v = h->poly2sph[(size_t)p * pc + l * l + m];
          if (v != 0.0) {
            idx[n] = (unsigned short)(N * (N + 1) * (N + 2) / 6 + l * (l - 1) / 2 + m); /* chi_roff(N, l) + m */
            val[n] = v;
            last = n++;
          }
        }
        if (last >= 0) idx[last] |= 0x8000u;
      }
    }
    off[cdT] = n;
    t.p2sOff = upload_const(d, off, (size_t)cdT + 1);
    t.p2sIdx = upload_const(d, idx, (size_t)(n > 0 ? n : 1));
    t.p2sVal = upload_const(d, val, (size_t)(n > 0 ? n : 1));
    cudaStreamSynchronize(d->s1);
    free(off);
    free(idx);
    free(val);
  }
  t.ijk = upload_const(d, h->ijk, (size_t)3 * cdT);
  t.ijkIndex = upload_const(d, h->ijkIndex, (size_t)h->ijkDim * h->ijkDim * h->ijkDim);
  t.small_r = upload_const(d, h->small_r, ECP_SMALL_SLOTS);
  t.small_w = upload_const(d, h->small_w, ECP_SMALL_SLOTS);
  t.small_oidx = upload_const(d, h->small_oidx, ECP_SMALL_SLOTS);
  {
    static int16_t slotOf[ECP_SMALL_SLOTS]; /* the small grid is the same for every handle */
    for (int i = 0; i < ECP_SMALL_SLOTS; i++) slotOf[i] = 1; /* pad slot */
    for (int k = 0; k < ECP_SMALL_SLOTS; k++)
      if (h->small_oidx[k] >= 0) slotOf[h->small_oidx[k]] = (int16_t)k;
    t.small_slotOf = upload_const(d, slotOf, ECP_SMALL_SLOTS);
  }
  {
    unsigned char *jl = (unsigned char *)malloc(2 * ECP_SMALL_LEVELS * ECP_SMALL_SLOTS), *jr = jl + ECP_SMALL_LEVELS * ECP_SMALL_SLOTS;
    if (!ecp_small_suffix_tables(&t.sm, h->small_oidx, jl, jr)) {
      snprintf(g_err, sizeof(g_err), "libecp_b200: small-grid slot table is not monotone per level");
      free(jl);
      ecpdev_destroy(d);
      return NULL;
    }
    t.small_jL = upload_const(d, jl, (size_t)ECP_SMALL_LEVELS * ECP_SMALL_SLOTS);
    t.small_jR = upload_const(d, jr, (size_t)ECP_SMALL_LEVELS * ECP_SMALL_SLOTS);
    cudaStreamSynchronize(d->s1);
    free(jl);
  }
  t.large_x = upload_const(d, h->large_x, h->largeSlots);
  t.large_w = upload_const(d, h->large_w, h->largeSlots);
  t.large_xo = upload_const(d, h->large_xo, h->largeOrder);
  t.besselT = upload_const(d, h->besselT, (size_t)1601 * h->besselStride);
  t.besselC = upload_const(d, h->besselC, h->besselLMax + 1);
  t.shellL = upload_const(d, h->shellL, h->nrShells);
  t.shellK = upload_const(d, h->shellK, h->nrShells);
  t.shellPrim = upload_const(d, h->shellPrim, h->nrShells);
  t.shellAtom = upload_const(d, h->shellAtom, h->nrShells);
  t.shellAO = upload_const(d, h->shellAO, h->nrShells);
  {
    int *aml = (int *)calloc(h->nrAtoms + 1, sizeof(int));
    for (int s = 0; s < h->nrShells; s++)
      if (h->shellL[s] > aml[h->shellAtom[s]]) aml[h->shellAtom[s]] = h->shellL[s];
    t.atomMaxL = upload_const(d, aml, h->nrAtoms);
    free(aml);
  }
  t.primD = upload_const(d, h->primD, h->nrPrims);
  t.primA = upload_const(d, h->primA, h->nrPrims);
  t.typeL = upload_const(d, h->typeL, h->nTypes);
  t.typeGaussOff = upload_const(d, h->typeGaussOff, h->nTypes + 1);
  const int ng = h->typeGaussOff[h->nTypes];
  t.gaussL = upload_const(d, h->gaussL, ng);
  t.gaussN = upload_const(d, h->gaussN, ng);
  t.gaussD = upload_const(d, h->gaussD, ng);
  t.gaussA = upload_const(d, h->gaussA, ng);
  { /* r^N U_l(r_n): host layout [type][l][N][slot] -> device layout [type][slot][l][N] (see k_fastT) */
    const size_t rows = (size_t)h->maxLECP * h->nU, tot = (size_t)h->nTypes * rows * ECP_SMALL_SLOTS;
    double *tr = (double *)malloc((tot ? tot : 1) * sizeof(double));
    for (int ty = 0; ty < h->nTypes; ty++)
      for (size_t r = 0; r < rows; r++)
        for (int sl = 0; sl < ECP_SMALL_SLOTS; sl++)
          tr[((size_t)ty * ECP_SMALL_SLOTS + sl) * rows + r] = h->typeUtab[((size_t)ty * rows + r) * ECP_SMALL_SLOTS + sl];
    t.typeUtabT = upload_const(d, tr, tot);
    cudaStreamSynchronize(d->s1); /* tr is pageable: the copy has left it when the stream is idle */
    free(tr);
  }
  cudaMemcpyToSymbolAsync(c_small_w, h->small_w, ECP_SMALL_SLOTS * sizeof(double), 0, cudaMemcpyHostToDevice, d->s1);
  cudaMemcpyToSymbolAsync(c_small_r, h->small_r, ECP_SMALL_SLOTS * sizeof(double), 0, cudaMemcpyHostToDevice, d->s1);
  cudaMemcpyToSymbolAsync(c_small_oidx, h->small_oidx, ECP_SMALL_SLOTS * sizeof(int16_t), 0, cudaMemcpyHostToDevice, d->s1);
  t.typeUL = upload_const(d, h->typeUL, (size_t)h->nTypes * ECP_SMALL_SLOTS);
  {
    const int nl = (ECP_MAX_LBS + 1) * (ECP_MAX_LBS + 1) * (ECP_MAX_LECP + 1);
    int *lut = (int *)malloc(nl * sizeof(int));
    for (int k = 0; k < nl; k++) lut[k] = -1;
    for (int c = 0; c < h->nClasses; c++)
      lut[(h->clsLa[c] * (ECP_MAX_LBS + 1) + h->clsLb[c]) * (ECP_MAX_LECP + 1) + h->clsL[c]] = c;
    t.clsLookup = upload_const(d, lut, (size_t)nl);
    cudaStreamSynchronize(d->s1);
    free(lut);
  }
  { /* spherical output: cart2sph of the shells' momenta, spherical row <-> shell maps */
    int off[ECP_MAX_LBS + 2], acc = 0;
    for (int l = 0; l <= ECP_MAX_LBS; l++) {
      off[l] = acc;
      acc += (2 * l + 1) * ((l + 1) * (l + 2) / 2);
    }
    int nSph = 0;
    for (int s2 = 0; s2 < h->nrShells; s2++) nSph += 2 * h->shellL[s2] + 1;
    int *ssph = (int *)malloc(((size_t)h->nrShells + 1) * sizeof(int)), *sof = (int *)malloc(((size_t)nSph + 1) * sizeof(int));
    for (int s2 = 0, k = 0; s2 < h->nrShells; s2++) {
      ssph[s2] = k;
      for (int m = 0; m < 2 * h->shellL[s2] + 1; m++) sof[k++] = s2;
    }
    d->nSph = nSph;
    d->shellSph = upload_const(d, ssph, (size_t)h->nrShells);
    d->sphShell = upload_const(d, sof, (size_t)(nSph > 0 ? nSph : 1));
    d->c2sOff = upload_const(d, off, (size_t)ECP_MAX_LBS + 1);
    const int need = off[h->maxLBS] + (2 * h->maxLBS + 1) * ((h->maxLBS + 1) * (h->maxLBS + 2) / 2);
    d->c2s = upload_const(d, h->cart2sph, (size_t)(need < h->ncart2sph ? need : h->ncart2sph));
    cudaStreamSynchronize(d->s1);
    free(ssph);
    free(sof);
  }
  t.clsLa = upload_const(d, h->clsLa, h->nClasses);
  t.clsLb = upload_const(d, h->clsLb, h->nClasses);
  t.clsL = upload_const(d, h->clsL, h->nClasses);
  t.clsNq = upload_const(d, h->clsNq, h->nClasses);
  t.clsQOff = upload_const(d, h->clsQOff, h->nClasses + 1);
  t.clsQlOff = upload_const(d, h->clsQlOff, (size_t)h->nClasses * (ECP_MAX_LECP + 1));
  t.qlist = upload_const(d, h->qlist, h->nqlist);
  t.clsQidxOff = upload_const(d, h->clsQidxOff, h->nClasses + 1);
  t.qidx = upload_const(d, h->qidx, h->nqidx);
  t.nAO = h->nAO;
  d->nClasses = h->nClasses;
  for (int c = 0; c < h->nClasses; c++) {
    d->hClsLa[c] = h->clsLa[c];
    d->hClsLb[c] = h->clsLb[c];
    d->hClsL[c] = h->clsL[c];
    d->hClsNq[c] = h->clsNq[c];
  }
  d->maxQPerL = h->maxQPerL;
  d->nAO = h->nAO;
  d->maxLBS = h->maxLBS;
  for (int l = 0; l <= h->maxLBS; l++)
    d->hShTerms[l] = h->shTermOff[l * h->shOffStride + (l + 1) * (l + 2) / 2] - h->shTermOff[l * h->shOffStride];
  adopt_cached(d);
  for (int i = 0; i < d->ntab; i++)
    if (!d->tab[i].p) {
      snprintf(g_err, sizeof(g_err), "libecp_b200: table upload %d failed: %s", i, cudaGetErrorString(cudaGetLastError()));
      ecpdev_destroy(d);
      return NULL;
    }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    ecpdev_destroy(d);
    return NULL;
  }
  return d;
}

/* scratch buffers (and the result matrix) of a destroyed handle are parked per device and adopted by the next handle
 * created there: a caller that goes through getIntegrals() creates a handle per call, and growing gigabytes of
 * scratch from the driver costs tens to hundreds of milliseconds each time.  libecp_b200_release_cache() frees them. */
#define ECP_NBUF 128
struct DevCache {
  int valid;
  Buf bufs[ECP_NBUF];
  double *matrix;
  size_t matrixBytes;
  int matrixKnown, dirtyAll, nDirty; /* matrix is zero outside the upper-triangle parts of the dirty rows */
  long long dirtySig;
  Buf dirtyRows;
};
static DevCache g_devCache[ECP_MAXDEV];
static std::mutex g_devCacheMu;
static int collect_bufs(EcpDev *d, Buf **bs) {
  Buf *list[] = {&d->rshX, &d->uspX, &d->omX, &d->F, &d->T, &d->gamma, &d->chi, &d->Q, &d->rshP, &d->sP, &d->blocks,
                 &d->tfail, &d->tflags, &d->items, &d->counters, &d->fastSurv, &d->t1list, &d->t1mask, &d->t1count,
                 &d->t1work, &d->t1rec, &d->trirec, &d->clsJ, &d->Jbuf, &d->fbItems, &d->fbList, &d->fbUnits, &d->fbTotals, &d->fbR,
                 &d->fbwItems, &d->fbwUnits, &d->fbwQd, &d->fbwSI, &d->fbwSP, &d->fbwSQ, &d->fbwRes, &d->fbwOpenFlag, &d->fbwVals,
                 &d->fbwListA, &d->fbwListB, &d->fbwCtr, &d->dlRows, &d->dlBits, &d->dlCount, &d->dlBase, &d->dlPay,
                 &d->t1surv, &d->t1smask, &d->t1scount, &d->t1state,
#define UPSET(i) &d->up[i].asAtom, &d->up[i].asType, &d->up[i].asR, &d->up[i].asOmOff, &d->up[i].ssShell,            \
                 &d->up[i].ssASlot, &d->up[i].ssStart, &d->up[i].ssEnd, &d->up[i].ssFOff, &d->up[i].trA, &d->up[i].trB, \
                 &d->up[i].trOut, &d->up[i].trPair, &d->up[i].prTriple, &d->up[i].clsFirst, &d->up[i].clsWork,        \
                 &d->up[i].clsElem, &d->up[i].clsOutElem, &d->up[i].clsPairBase, &d->up[i].clsQBase,                 \
                 &d->up[i].ceAS0, &d->up[i].cePair0, &d->up[i].asSS0, &d->up[i].ssOwn, &d->up[i].ccTri, &d->up[i].ccPair, \
                 &d->up[i].pairCnt, &d->up[i].meta, &d->up[i].rshX, &d->up[i].uspX, &d->up[i].omX, &d->up[i].F
                 UPSET(0), UPSET(1)};
#undef UPSET
  const int n = (int)(sizeof(list) / sizeof(list[0]));
  static_assert(sizeof(list) / sizeof(list[0]) <= ECP_NBUF, "ECP_NBUF too small");
  for (int i = 0; i < n; i++) bs[i] = list[i];
  return n;
}
static void adopt_cached(EcpDev *d) {
  if (d->device < 0 || d->device >= ECP_MAXDEV) return;
  std::lock_guard<std::mutex> lk(g_devCacheMu);
  DevCache &c = g_devCache[d->device];
  if (!c.valid) return;
  Buf *bs[ECP_NBUF];
  const int n = collect_bufs(d, bs);
  for (int i = 0; i < n; i++) *bs[i] = c.bufs[i];
  const size_t bytes = (size_t)d->nAO * d->nAO * sizeof(double);
  if (c.matrix && c.matrixBytes >= bytes && c.matrixBytes <= 2 * bytes + (1 << 20)) {
    d->matrix = c.matrix;
    d->matrixBytes = c.matrixBytes;
    d->matrixKnown = 0; /* another handle's result */
  } else if (c.matrix) {
    cudaFreeAsync(c.matrix, d->s1);
  }
  memset(&c, 0, sizeof(c));
}
extern "C" void ecpdev_release_cache(void) {
  std::lock_guard<std::mutex> lk(g_devCacheMu);
  for (int dev = 0; dev < ECP_MAXDEV; dev++) {
    DevCache &c = g_devCache[dev];
    if (!c.valid) continue;
    cudaSetDevice(dev);
    for (int i = 0; i < ECP_NBUF; i++)
      if (c.bufs[i].p) cudaFree(c.bufs[i].p);
    if (c.matrix) cudaFree(c.matrix);
    memset(&c, 0, sizeof(c));
  }
}

extern "C" void ecpdev_destroy(EcpDev *d) {
  if (!d) return;
  cudaSetDevice(d->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < d->ntab; i++)
    if (d->tab[i].p) cudaFreeAsync(d->tab[i].p, d->s1);
  Buf *bs[ECP_NBUF];
  const int n = collect_bufs(d, bs);
  bool parked = false;
  /* LIBECP_B200_NO_PARK=1: give the scratch and the result buffer back to the driver with the handle (a caller that
   * shares the GPU with other allocators); default: parked for the next handle on this device, libecp_b200_release_cache
   * frees them */
  if (d->device >= 0 && d->device < ECP_MAXDEV && !getenv("LIBECP_B200_NO_PARK")) {
    std::lock_guard<std::mutex> lk(g_devCacheMu);
    DevCache &c = g_devCache[d->device];
    if (!c.valid) {
      for (int i = 0; i < n; i++) c.bufs[i] = *bs[i];
      c.matrix = d->matrix;
      c.matrixBytes = d->matrixBytes;
      c.valid = 1;
      parked = true;
    }
  }
  if (!parked) {
    for (int i = 0; i < n; i++)
      if (bs[i]->p) cudaFreeAsync(bs[i]->p, d->s1);
    cudaStreamSynchronize(d->s1);
    if (d->matrix) cudaFreeAsync(d->matrix, d->s1);
  }
  comm_release(d);
  if (d->sphMatrix.p) cudaFreeAsync(d->sphMatrix.p, d->s1);
  for (int k = 0; k < 2; k++)
    if (d->enumHost[k]) ecpdev_pinned_free(d->enumHost[k]);
  if (d->agRows.p) cudaFreeAsync(d->agRows.p, d->s1);
  if (d->agOff.p) cudaFreeAsync(d->agOff.p, d->s1);
  if (d->agBuf.p) cudaFreeAsync(d->agBuf.p, d->s1);
  if (d->dirtyRows.p) cudaFreeAsync(d->dirtyRows.p, d->s1);
  if (d->gatherRows.p) cudaFreeAsync(d->gatherRows.p, d->s1);
  if (d->gatherOff.p) cudaFreeAsync(d->gatherOff.p, d->s1);
  cudaStreamSynchronize(d->s1);
  for (int i = 0; i < 12; i++) cudaEventDestroy(d->ev[i]);
  cudaStreamDestroy(d->s1);
  cudaStreamDestroy(d->s2real);
  cudaStreamDestroy(d->s3);
  cudaStreamDestroy(d->s4);
  for (int i = 0; i < 2; i++) cudaEventDestroy(d->evUp[i]);
  for (int i = 0; i < 2; i++) cudaEventDestroy(d->evTab[i]);
  cudaEventDestroy(d->evDone);
  free(d);
}

/* zero M[i][i..n) of the listed rows (all rows when rows == NULL) */
__global__ void k_zero_rows(double *__restrict__ M, int n, const int *__restrict__ rows) {
  const int i = rows ? rows[blockIdx.x] : blockIdx.x;
  double *r = M + (size_t)i * n;
  for (int j = i + threadIdx.x; j < n; j += blockDim.x) r[j] = 0.0;
}
/* Result matrix, zeroed.  A pass only ever writes the upper-triangle part of the AO rows of the shells its rank owns
 * (rowOwned, nAO bytes; NULL = all rows).  After the first full clear only those parts are cleared again - with the
 * rows dealt to 8 ranks that is 1/16 of the 2.9 GB a full memset of the 500-atom matrix touches, a fixed cost that
 * would otherwise not shrink with the number of GPUs.  `sig` identifies the ownership (rank, world). */
extern "C" int ecpdev_matrix_begin(EcpDev *d, const unsigned char *rowOwned, long long sig) {
  CK(cudaSetDevice(d->device));
  const int n = d->nAO;
  const size_t bytes = (size_t)n * n * sizeof(double);
  if (!d->matrix) {
    CK(cudaMallocAsync((void **)&d->matrix, bytes ? bytes : 8, d->s1));
    d->matrixBytes = bytes ? bytes : 8;
    d->matrixKnown = 0;
  }
  if (!d->matrixKnown) {
    CK(cudaMemsetAsync(d->matrix, 0, bytes, d->s1));
  } else if (d->nDirty > 0) {
    k_zero_rows<<<d->nDirty, 256, 0, d->s1>>>(d->matrix, n, d->dirtyAll ? NULL : (const int *)d->dirtyRows.p);
  }
  /* rows the coming pass may write */
  if (!d->matrixKnown || sig != d->dirtySig) {
    if (!rowOwned) {
      d->dirtyAll = 1;
      d->nDirty = n;
    } else {
      int *rows = (int *)malloc((size_t)(n + 1) * sizeof(int));
      int nr = 0;
      for (int i = 0; i < n; i++)
        if (rowOwned[i]) rows[nr++] = i;
      g_allocStream = d->s1;
      int rc_ = ensure(&d->dirtyRows, (size_t)(nr + 1) * sizeof(int));
      if (rc_) {
        free(rows);
        return rc_;
      }
      if (nr) CK(cudaMemcpyAsync(d->dirtyRows.p, rows, (size_t)nr * sizeof(int), cudaMemcpyHostToDevice, d->s1));
      CK(cudaStreamSynchronize(d->s1)); /* rows is pageable */
      free(rows);
      d->dirtyAll = 0;
      d->nDirty = nr;
    }
    d->dirtySig = sig;
  }
  d->matrixKnown = 1;
  return 0;
}
extern "C" int ecpdev_matrix_download(EcpDev *d, double *host) {
  CK(cudaSetDevice(d->device));
  CK(cudaStreamSynchronize(d->s1));
  CK(cudaMemcpy(host, d->matrix, (size_t)d->nAO * d->nAO * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}
/* host I[i*rowdim + j] += M[i][j] for j >= i (what libECP_callback0 does block by block, reference
 * src/getIntegrals.c:36-42): upper-triangle row panels are copied D2H into two pinned staging buffers and
 * added by all host threads while the next panel is in flight. */
/* Device-resident gather of a sharded result: pack the upper-triangle parts M[i][i..n) of the listed rows
 * of the handle's matrix back to back into the caller's DEVICE buffer (dir = 0), or scatter such a packed buffer - e.g.
 * another rank's shard after an NCCL all-gather - into the handle's matrix (dir = 1).  rows: ascending AO rows on the
 * host.  Returns 0 and the number of doubles the rows take in *elems; with devBuf == NULL only counts. */
extern "C" int ecpdev_matrix_rows(EcpDev *d, int dir, const int *rows, long long nrows, void *devBuf, long long cap,
                                  long long *elems) {
  CK(cudaSetDevice(d->device));
  const int n = d->nAO;
  long long *off = (long long *)malloc((size_t)(nrows + 2) * sizeof(long long));
  long long tot = 0;
  for (long long k = 0; k < nrows; k++) {
    if (rows[k] < 0 || rows[k] >= n || (k && rows[k] <= rows[k - 1])) {
      free(off);
      snprintf(g_err, sizeof(g_err), "ecpdev_matrix_rows: row list must be ascending AO rows");
      return -1;
    }
    off[k] = tot;
    tot += n - rows[k];
  }
  if (elems) *elems = tot;
  if (!devBuf || nrows == 0) {
    free(off);
    return 0;
  }
  if (tot > cap || !d->matrix) {
    free(off);
    snprintf(g_err, sizeof(g_err), "ecpdev_matrix_rows: %s", d->matrix ? "buffer too small" : "no result matrix yet");
    return -1;
  }
  g_allocStream = d->s1;
  int rc = ensure(&d->gatherRows, (size_t)(nrows + 1) * sizeof(int));
  if (!rc) rc = ensure(&d->gatherOff, (size_t)(nrows + 2) * sizeof(long long));
  if (rc) {
    free(off);
    return rc;
  }
  CK(cudaMemcpyAsync(d->gatherRows.p, rows, (size_t)nrows * sizeof(int), cudaMemcpyHostToDevice, d->s1));
  CK(cudaMemcpyAsync(d->gatherOff.p, off, (size_t)nrows * sizeof(long long), cudaMemcpyHostToDevice, d->s1));
  if (dir == 0)
    k_pack_rows<<<(unsigned)nrows, 256, 0, d->s1>>>(d->matrix, n, (const int *)d->gatherRows.p,
                                                  (const long long *)d->gatherOff.p, 0, (double *)devBuf);
  else
    k_unpack_rows<<<(unsigned)nrows, 256, 0, d->s1>>>(d->matrix, n, (const int *)d->gatherRows.p,
                                                    (const long long *)d->gatherOff.p, (const double *)devBuf);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(d->s1)); /* off / rows are pageable; the caller's collective runs on its own stream */
  free(off);
  if (dir == 1) d->matrixKnown = 0; /* rows of other ranks are now non-zero: the next pass clears everything */
  return 0;
}

/* Sparse variant of the host consumer (default; LIBECP_B200_D2H=dense keeps the dense panels): flag the non-zero
 * 16-element runs of the listed rows, read the per-row counts and bitmaps back (a few MB), pack the flagged runs on the
 * device in row order, move them in chunks of ~24 MB through two page-locked buffers and add them into the caller's matrix
 * by all host threads while the next chunk is in flight.  The caller's matrix is touched exactly where the dense variant
 * touches it with a non-zero: same result bit for bit. */
static int matrix_add_to_host_sparse(EcpDev *d, double *host, int rowdim, const int *rows, int nR, cudaStream_t st,
                                     long long *bytes) {
  const int n = d->nAO;
  const int maxRuns = (n + SPR_RUN - 1) / SPR_RUN, words = (maxRuns + 31) / 32;
  const size_t chunkRuns = ((size_t)24 << 20) / (SPR_RUN * sizeof(double));
  const bool trace = getenv("LIBECP_B200_TRACE") != NULL;
  const double t0 = omp_get_wtime();
  /* device buffers of the download live in the handle and are parked with its scratch (one download at a time, see
   * api.c): cudaMallocAsync beside a running batch was measured at up to 0.7 s when the pool had to grow - and the batch
   * stalled with it (profiles/r2/session3/e2e_outliers_*.txt).  Sized for all rows once, so that no download of a pass
   * allocates. */
  Buf &dRows = d->dlRows, &dBits = d->dlBits, &dCount = d->dlCount, &dBase = d->dlBase, &dPay = d->dlPay;
  g_allocStream = st;
  int rc = ensure(&dRows, (size_t)n * sizeof(int));
  if (!rc) rc = ensure(&dBits, (size_t)n * words * sizeof(unsigned));
  if (!rc) rc = ensure(&dCount, (size_t)n * sizeof(int));
  if (!rc) rc = ensure(&dBase, (size_t)n * sizeof(long long));
  if (rc) return rc;
  const double tP1 = omp_get_wtime();
  const size_t metaBytes = (size_t)nR * words * sizeof(unsigned) + (size_t)nR * sizeof(int);
  unsigned *hBits = (unsigned *)ecpdev_pinned_alloc(metaBytes);
  double *pin[2] = {(double *)ecpdev_pinned_alloc(chunkRuns * SPR_RUN * sizeof(double) + (size_t)n * sizeof(double)),
                    (double *)ecpdev_pinned_alloc(chunkRuns * SPR_RUN * sizeof(double) + (size_t)n * sizeof(double))};
  long long *base = (long long *)malloc((size_t)(nR + 1) * sizeof(long long));
  int *cb = (int *)malloc((size_t)(nR + 2) * sizeof(int));
  cudaEvent_t done[2] = {NULL, NULL};
  long long moved = 0;
  int nChunks = 0;
  double tWait = 0, tAdd = 0, tMeta = 0;
  int *hCount = NULL;
  const double tP2 = omp_get_wtime();
  double tP3 = 0, tP4 = 0, tP5 = 0;
#define SPR_CK(call)                                                                                       \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) {                                                                               \
      snprintf(g_err, sizeof(g_err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));  \
      rc = (int)e_;                                                                                        \
      goto spr_out;                                                                                        \
    }                                                                                                      \
  } while (0)
  if (!hBits || !pin[0] || !pin[1] || !base || !cb) {
    snprintf(g_err, sizeof(g_err), "ecpdev_matrix_add_to_host: no staging memory");
    rc = -1;
    goto spr_out;
  }
  hCount = (int *)(hBits + (size_t)nR * words);
  for (int k = 0; k < 2; k++) SPR_CK(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
  SPR_CK(cudaMemcpyAsync(dRows.p, rows, (size_t)nR * sizeof(int), cudaMemcpyHostToDevice, st));
  k_rows_flag<<<nR, 256, words * sizeof(unsigned), st>>>(d->matrix, n, (const int *)dRows.p, words, (unsigned *)dBits.p,
                                                        (int *)dCount.p);
  SPR_CK(cudaGetLastError());
  SPR_CK(cudaMemcpyAsync(hBits, dBits.p, (size_t)nR * words * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  SPR_CK(cudaMemcpyAsync(hCount, dCount.p, (size_t)nR * sizeof(int), cudaMemcpyDeviceToHost, st));
  tP3 = omp_get_wtime();
  SPR_CK(cudaStreamSynchronize(st));
  tP4 = omp_get_wtime();
  moved += (long long)metaBytes;
  /* row offsets (in runs) and chunk boundaries at row boundaries */
  base[0] = 0;
  for (int k = 0; k < nR; k++) base[k + 1] = base[k] + hCount[k];
  for (int k = 0; k < nR;) {
    int e = k + 1;
    while (e < nR && (size_t)(base[e + 1] - base[k]) <= chunkRuns) e++;
    cb[nChunks++] = k;
    k = e;
  }
  cb[nChunks] = nR;
  if (base[nR] > 0) {
    { /* grow in large steps: the payload of the whole matrix is a few hundred MB */
      size_t need = (size_t)base[nR] * SPR_RUN * sizeof(double);
      if (need > dPay.cap && need < ((size_t)256 << 20)) need = (size_t)256 << 20;
      rc = ensure(&dPay, need);
    }
    tP5 = omp_get_wtime();
    if (rc) goto spr_out;
    SPR_CK(cudaMemcpyAsync(dBase.p, base, (size_t)nR * sizeof(long long), cudaMemcpyHostToDevice, st));
    k_rows_pack_sparse<<<nR, 256, 2 * words * sizeof(unsigned), st>>>(d->matrix, n, (const int *)dRows.p, words,
                                                                     (const unsigned *)dBits.p, (const long long *)dBase.p,
                                                                     (double *)dPay.p);
    SPR_CK(cudaGetLastError());
  }
  tMeta = omp_get_wtime() - t0;
  {
    auto issue = [&](int p) -> cudaError_t {
      const long long r0 = base[cb[p]], r1 = base[cb[p + 1]];
      moved += (r1 - r0) * SPR_RUN * (long long)sizeof(double);
      if (r1 > r0) {
        cudaError_t e = cudaMemcpyAsync(pin[p & 1], (const double *)dPay.p + r0 * SPR_RUN,
                                        (size_t)(r1 - r0) * SPR_RUN * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
      }
      return cudaEventRecord(done[p & 1], st);
    };
    if (nChunks) SPR_CK(issue(0));
    for (int p = 0; p < nChunks; p++) {
      if (p + 1 < nChunks) SPR_CK(issue(p + 1));
      const double tw0 = omp_get_wtime();
      SPR_CK(cudaEventSynchronize(done[p & 1]));
      const double tw1 = omp_get_wtime();
      tWait += tw1 - tw0;
      const int k0 = cb[p], k1 = cb[p + 1];
      const double *src = pin[p & 1];
      const long long b0 = base[k0];
      /* the runs of a row land at scattered places of the caller's matrix: every run is two or three cache misses the
       * hardware prefetcher does not see coming (measured 1.7 GB/s of payload per thread).  The run positions of a row are
       * decoded from its bitmap first and the destinations of the runs SPR_PF ahead are prefetched for writing. */
#pragma omp parallel
      {
        int *pos = (int *)malloc((size_t)(maxRuns + 1) * sizeof(int));
#pragma omp for schedule(dynamic, 8)
        for (int k = k0; k < k1; k++) {
          if (!hCount[k] || !pos) continue;
          const int i = rows[k];
          const double *pr = src + (base[k] - b0) * SPR_RUN;
          double *dr = host + (size_t)i * rowdim + i;
          const unsigned *bm = hBits + (size_t)k * words;
          const int len = n - i;
          int nr = 0;
          for (int w = 0; w < words; w++) {
            unsigned bits = bm[w];
            while (bits) {
              pos[nr++] = (w * 32 + __builtin_ctz(bits)) * SPR_RUN;
              bits &= bits - 1;
            }
          }
          for (int u = 0; u < nr; u++, pr += SPR_RUN) {
            if (u + SPR_PF < nr) {
              const double *pn = dr + pos[u + SPR_PF];
              __builtin_prefetch(pn, 1, 0);
              __builtin_prefetch(pn + 8, 1, 0);
              __builtin_prefetch(pn + SPR_RUN - 1, 1, 0);
            }
            const int j = pos[u];
            double *dj = dr + j;
            if (j + SPR_RUN <= len)
              for (int q = 0; q < SPR_RUN; q++) dj[q] += pr[q];
            else
              for (int q = 0; j + q < len; q++) dj[q] += pr[q];
          }
        }
        free(pos);
      }
      tAdd += omp_get_wtime() - tw1;
    }
  }
  if (trace) {
    long long allRuns = 0;
    for (int k = 0; k < nR; k++) allRuns += (n - rows[k] + SPR_RUN - 1) / SPR_RUN;
    fprintf(stderr, "[libecp_b200] sparse d2h+add: rows %d runs %lld of %lld chunks %d bytes %.1f MB flag+meta+pack %.1f ms (dev alloc %.1f, pinned %.1f, issue %.1f, sync %.1f, payload alloc %.1f) wait %.1f ms add %.1f ms threads %d\n",
            nR, base[nR], allRuns, nChunks, moved / 1e6, 1e3 * tMeta, 1e3 * (tP1 - t0), 1e3 * (tP2 - tP1), 1e3 * (tP3 - tP2),
            1e3 * (tP4 - tP3), 1e3 * (tP5 - tP4), 1e3 * tWait, 1e3 * tAdd, omp_get_max_threads());
  }
spr_out:
#undef SPR_CK
  for (int k = 0; k < 2; k++) {
    if (done[k]) cudaEventDestroy(done[k]);
    if (pin[k]) ecpdev_pinned_free(pin[k]);
  }
  if (hBits) ecpdev_pinned_free(hBits);
  free(base);
  free(cb);
  if (bytes) *bytes = moved;
  return rc;
}

extern "C" int ecpdev_matrix_add_to_host(EcpDev *d, double *host, int rowdim, const unsigned char *rowOwned,
                                         long long *bytes, int async) {
  /* async = 1: the listed rows are final (their pass has returned) while another pass may be running on the compute
   * stream - use the download stream and never touch the compute stream (called from a helper host thread) */
  CK(cudaSetDevice(d->device));
  cudaStream_t st = async ? d->s4 : d->s1;
  if (!async) CK(cudaStreamSynchronize(d->s1));
  const int n = d->nAO;
  const size_t panelBytes = (size_t)24 << 20;
  /* rows this rank owns (all rows when rowOwned == NULL): a sharded rank's partial matrix is zero outside the AO
   * rows of its shells, so other rows are never transferred.  The upper-triangle parts M[i][i..n) of the owned
   * rows are packed back to back on the device (k_pack_rows), moved in panels of ~24 MB with one contiguous copy
   * each, and added to the caller's matrix by all host threads while the next panel is in flight. */
  int *rows = (int *)malloc((size_t)(n + 1) * sizeof(int));
  long long *off = (long long *)malloc((size_t)(n + 2) * sizeof(long long));
  int nR = 0;
  long long tot = 0;
  for (int i = 0; i < n; i++)
    if (!rowOwned || rowOwned[i]) {
      rows[nR] = i;
      off[nR] = tot;
      tot += n - i;
      nR++;
    }
  off[nR] = tot;
  const long long panelElems = (long long)(panelBytes / sizeof(double)) > n ? (long long)(panelBytes / sizeof(double)) : n;
  {
    const char *e = getenv("LIBECP_B200_D2H");
    if (!(e && !strcmp(e, "dense")) && nR) {
      const int rc = matrix_add_to_host_sparse(d, host, rowdim, rows, nR, st, bytes);
      free(rows);
      free(off);
      return rc;
    }
  }
  Buf dRows = {NULL, 0}, dOff = {NULL, 0}, dStage[2] = {{NULL, 0}, {NULL, 0}};
  g_allocStream = st;
  int rc = ensure(&dRows, (size_t)(nR + 1) * sizeof(int));
  if (!rc) rc = ensure(&dOff, (size_t)(nR + 2) * sizeof(long long));
  if (!rc) rc = ensure(&dStage[0], (size_t)panelElems * sizeof(double));
  if (!rc) rc = ensure(&dStage[1], (size_t)panelElems * sizeof(double));
  if (rc) return rc;
  /* pinned staging buffers come from the process-wide page-locked cache (mutex-protected, blocks marked in use: two
   * handles driven from different host threads never share one; cudaMallocHost of tens of MB costs far more than the
   * transfer itself, so the blocks stay cached after the call) */
  double *pin[2];
  for (int k = 0; k < 2; k++) {
    pin[k] = (double *)ecpdev_pinned_alloc((size_t)panelElems * sizeof(double));
    if (!pin[k]) {
      if (k) ecpdev_pinned_free(pin[0]);
      snprintf(g_err, sizeof(g_err), "ecpdev_matrix_add_to_host: no staging memory");
      return -1;
    }
  }
  cudaEvent_t done[2];
  for (int k = 0; k < 2; k++) CK(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
  if (nR) {
    CK(cudaMemcpyAsync(dRows.p, rows, (size_t)nR * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dOff.p, off, (size_t)(nR + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  }
  /* panel boundaries in row-list positions */
  int *pb = (int *)malloc((size_t)(nR + 2) * sizeof(int));
  int nPanels = 0;
  for (int k = 0; k < nR;) {
    int e = k + 1;
    while (e < nR && off[e + 1] - off[k] <= panelElems) e++;
    pb[nPanels++] = k;
    k = e;
  }
  pb[nPanels] = nR;
  long long moved = 0;
  auto issue = [&](int p) -> cudaError_t {
    const int k0 = pb[p], k1 = pb[p + 1];
    const long long elems = off[k1] - off[k0];
    k_pack_rows<<<k1 - k0, 256, 0, st>>>(d->matrix, n, (const int *)dRows.p + k0, (const long long *)dOff.p + k0,
                                           off[k0], (double *)dStage[p & 1].p);
    moved += elems * (long long)sizeof(double);
    cudaError_t e = cudaMemcpyAsync(pin[p & 1], dStage[p & 1].p, (size_t)elems * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(done[p & 1], st);
  };
  const bool trace = getenv("LIBECP_B200_TRACE") != NULL;
  double tWait = 0, tAdd = 0, tStart = omp_get_wtime();
  if (nPanels) CK(issue(0));
  for (int p = 0; p < nPanels; p++) {
    if (p + 1 < nPanels) CK(issue(p + 1));
    const double tw0 = omp_get_wtime();
    CK(cudaEventSynchronize(done[p & 1]));
    const double tw1 = omp_get_wtime();
    tWait += tw1 - tw0;
    const int k0 = pb[p], k1 = pb[p + 1];
    const double *src = pin[p & 1];
    const long long base = off[k0];
#pragma omp parallel for schedule(dynamic, 8)
    for (int k = k0; k < k1; k++) {
      const int i = rows[k];
      const double *sr = src + (off[k] - base) - i; /* sr[j] = M[i][j] for j >= i */
      double *dr = host + (size_t)i * rowdim;
      /* the ECP matrix is block sparse (a block is non-zero only if both shells reach a common centre):
       * runs of 32 zeros are skipped so that the caller's matrix is only touched where something is added */
      int j = i;
      for (; j + 32 <= n; j += 32) {
        const double *s32 = sr + j;
        int any = 0;
        for (int q = 0; q < 32; q++) any |= (s32[q] != 0.0);
        if (any)
          for (int q = 0; q < 32; q++) dr[j + q] += s32[q];
      }
      for (; j < n; j++) dr[j] += sr[j];
    }
    tAdd += omp_get_wtime() - tw1;
  }
  if (trace)
    fprintf(stderr, "[libecp_b200] d2h+add: rows %d panels %d bytes %.1f MB setup+alloc %.1f ms wait %.1f ms add %.1f ms threads %d\n",
            nR, nPanels, moved / 1e6, 0.0, 1e3 * tWait, 1e3 * tAdd, omp_get_max_threads());
  (void)tStart;
  for (int k = 0; k < 2; k++) {
    cudaEventDestroy(done[k]);
    cudaFreeAsync(dStage[k].p, st);
    ecpdev_pinned_free(pin[k]);
  }
  cudaFreeAsync(dRows.p, st);
  cudaFreeAsync(dOff.p, st);
  free(rows);
  free(off);
  free(pb);
  if (bytes) *bytes = moved;
  return 0;
}
/* diagnostic builds (-DT1_STATS): read and clear the lane / chunk statistics of k_type1S; 0 = not compiled in */
extern "C" int ecpdev_t1stats(unsigned long long *out) {
#ifdef T1_STATS
  unsigned long long z[16] = {0};
  if (cudaMemcpyFromSymbol(out, g_t1stats, sizeof(z)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_t1stats, z, sizeof(z));
  return 1;
#else
  (void)out;
  return 0;
#endif
}
extern "C" void *ecpdev_matrix_ptr(EcpDev *d) { return d->matrix; }

/* ---- spherical-harmonic output (scope row f4, second half: "output to spherical AOs") ----
 * The reference carries the Cartesian -> real-spherical-harmonic matrix (TM_cart2sph, src/transformations.c:28-87; its
 * inverse TM_sph2cart :91-141 is unused) but offers no spherical output; callers with pure (5d / 7f) functions transform
 * the Cartesian matrix themselves.  Here the resident matrix is transformed on the device, per shell pair
 *     S[(a,m)][(b,m')] = sum_{c,c'} cart2sph[la][m][c] cart2sph[lb][m'][c'] M[(a,c)][(b,c')],
 * with the same table (bit-identical to the reference's) and the handle's component order.  A thread per element of the
 * upper triangle of S; the diagonal shell blocks of M hold their upper triangle only (src/getIntegrals.c:36-41) and are
 * read symmetrically. */
__global__ void k_cart2sph_matrix(const double *__restrict__ M, int nAO, double *__restrict__ S, int nSph,
                                  const int *__restrict__ sphShell, const int *__restrict__ shellSph,
                                  const int *__restrict__ shellAO, const int *__restrict__ shellL,
                                  const double *__restrict__ c2s, const int *__restrict__ c2sOff) {
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nSph || j < i) return;
  const int a = sphShell[i], b = sphShell[j];
  const int la = shellL[a], lb = shellL[b], na = (la + 1) * (la + 2) / 2, nb = (lb + 1) * (lb + 2) / 2;
  const double *ca = c2s + c2sOff[la] + (i - shellSph[a]) * na, *cb = c2s + c2sOff[lb] + (j - shellSph[b]) * nb;
  const int r0 = shellAO[a], c0 = shellAO[b];
  double acc = 0.0;
  for (int c = 0; c < na; c++) {
    double row = 0.0;
    for (int cc = 0; cc < nb; cc++) {
      int r = r0 + c, q = c0 + cc;
      if (r > q) { /* diagonal shell block: mirror */
        const int tsw = r;
        r = q;
        q = tsw;
      }
      row = fma(cb[cc], M[(size_t)r * nAO + q], row);
    }
    acc = fma(ca[c], row, acc);
  }
  S[(size_t)i * nSph + j] = acc;
}
extern "C" int ecpdev_spherical(EcpDev *d, void **devS, int *nSphOut) {
  CK(cudaSetDevice(d->device));
  if (!d->matrix) {
    snprintf(g_err, sizeof(g_err), "ecpdev_spherical: no result matrix yet");
    return -1;
  }
  g_allocStream = d->s1;
  const int nSph = d->nSph;
  int rc = ensure(&d->sphMatrix, (size_t)nSph * nSph * sizeof(double));
  if (rc) return rc;
  CK(cudaMemsetAsync(d->sphMatrix.p, 0, (size_t)nSph * nSph * sizeof(double), d->s1));
  dim3 grid((nSph + 127) / 128, nSph);
  k_cart2sph_matrix<<<grid, 128, 0, d->s1>>>(d->matrix, d->nAO, (double *)d->sphMatrix.p, nSph, d->sphShell, d->shellSph,
                                            d->t.shellAO, d->t.shellL, d->c2s, d->c2sOff);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(d->s1));
  if (devS) *devS = d->sphMatrix.p;
  if (nSphOut) *nSphOut = nSph;
  return 0;
}
extern "C" int ecpdev_spherical_add_to_host(EcpDev *d, double *host, int rowdim) {
  void *dS = NULL;
  int n = 0;
  int rc = ecpdev_spherical(d, &dS, &n);
  if (rc) return rc;
  double *tmp = (double *)malloc((size_t)n * n * sizeof(double));
  if (!tmp) return -1;
  CK(cudaMemcpy(tmp, dS, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++)
    for (int j = i; j < n; j++) host[(size_t)i * rowdim + j] += tmp[(size_t)i * n + j];
  free(tmp);
  return 0;
}
/* make the handle's device current on the calling host thread (the builder thread allocates page-locked memory) */
extern "C" void ecpdev_bind_thread(EcpDev *d) {
  if (d) cudaSetDevice(d->device);
}
/* serial = 1: type-1 kernels share the type-2 stream, so the per-kernel event times are not inflated by overlap */
extern "C" void ecpdev_set_serial(EcpDev *d, int on) { d->serial = on; }
extern "C" long long ecpdev_table_bytes(EcpDev *d) { return d->tableBytes; }
extern "C" int ecpdev_sync(EcpDev *d) {
  CK(cudaSetDevice(d->device));
  CK(cudaStreamSynchronize(d->s1));
  CK(cudaStreamSynchronize(d->s2));
  return 0;
}

/* per-centre tables of one batch into the buffers of upload set `slot` on stream st (reference: evalAngularIntegrals,
 * unitarySpherePolynomials, calcF_FM06 per centre, src/libecp.c:262-292) */
static int launch_tables(EcpDev *d, const EcpBatch *h, int slot, cudaStream_t st) {
  EcpDev::UpSet &u = d->up[slot];
  const DevT &t = d->t;
  g_allocStream = st;
  int rc_ = ensure(&u.rshX, ((size_t)h->nASlots * RSHX_STRIDE + 1) * sizeof(double));
  if (!rc_) rc_ = ensure(&u.uspX, ((size_t)h->nASlots * USPX_STRIDE + 1) * sizeof(double));
  if (!rc_) rc_ = ensure(&u.omX, ((size_t)h->omTotal + 1) * sizeof(double));
  if (!rc_) rc_ = ensure(&u.F, ((size_t)h->fRows * ECP_SMALL_SLOTS + 1) * sizeof(double));
  if (rc_) return rc_;
  DevB B;
  memset(&B, 0, sizeof(B));
  B.nASlots = h->nASlots;
  B.nSSlots = h->nSSlots;
  B.asAtom = (const int *)u.asAtom.p; B.asType = (const int *)u.asType.p; B.asR = (const double *)u.asR.p;
  B.asOmOff = (const long long *)u.asOmOff.p; B.ssShell = (const int *)u.ssShell.p; B.ssASlot = (const int *)u.ssASlot.p;
  B.ssStart = (const int *)u.ssStart.p; B.ssEnd = (const int *)u.ssEnd.p; B.ssFOff = (const long long *)u.ssFOff.p;
  B.rshX = (double *)u.rshX.p; B.uspX = (double *)u.uspX.p; B.omX = (double *)u.omX.p; B.F = (double *)u.F.p;
  k_atomslot<<<(unsigned)((h->nASlots + 127) / 128), 128, 0, st>>>(t, B);
  k_omegaX<<<h->nASlots, 256, 0, st>>>(t, B);
  if (d->ftabCompact) { /* window-only tabulation into a cleared table (k_Ftab2) */
    CK(cudaMemsetAsync(B.F, 0, (size_t)h->fRows * ECP_SMALL_SLOTS * sizeof(double), st));
    if (t.maxLECP - 1 + d->maxLBS <= 6)
      k_Ftab2<6><<<h->nSSlots, 128, 0, st>>>(t, B);
    else
      k_Ftab2<KM><<<h->nSSlots, 128, 0, st>>>(t, B);
  } else if (t.maxLECP - 1 + d->maxLBS <= 6)
    k_Ftab<6><<<h->nSSlots, ECP_SMALL_SLOTS, 0, st>>>(t, B);
  else
    k_Ftab<KM><<<h->nSSlots, ECP_SMALL_SLOTS, 0, st>>>(t, B);
  CK(cudaGetLastError());
  return 0;
}

#define UP(buf, src, n, T)                                                                                \
  do {                                                                                                    \
    int rc_ = ensure(&u.buf, ((n) ? (n) : 1) * sizeof(T));                                                \
    if (rc_) return rc_;                                                                                  \
    if (n) CK(cudaMemcpyAsync(u.buf.p, src, (size_t)(n) * sizeof(T), cudaMemcpyHostToDevice, st));        \
    d->upBytes[slot] += (long long)((size_t)(n) * sizeof(T));                                             \
  } while (0)
/* copy the arrays of one batch into upload set `slot` on stream st */
static int upload_set(EcpDev *d, const EcpBatch *h, int flags, int slot, cudaStream_t st) {
  EcpDev::UpSet &u = d->up[slot];
  const int nc = d->nClasses;
  g_allocStream = st;
  d->upBytes[slot] = 0;
  UP(asAtom, h->asAtom, h->nASlots, int);
  UP(asType, h->asType, h->nASlots, int);
  UP(asR, h->asR, (size_t)h->nASlots * 4, double);
  UP(asOmOff, (const long long *)h->asOmOff, h->nASlots, long long);
  UP(ssShell, h->ssShell, h->nSSlots, int);
  UP(ssASlot, h->ssASlot, h->nSSlots, int);
  UP(ssStart, h->ssStart, h->nSSlots, int);
  UP(ssEnd, h->ssEnd, h->nSSlots, int);
  UP(ssFOff, (const long long *)h->ssFOff, h->nSSlots, long long);
  if (h->devEnum) { /* triples are enumerated on the device: count + scan here, fill once the sizes are known */
    UP(ceAS0, h->ceAS0, h->nCentres + 1, int);
    UP(cePair0, (const long long *)h->cePair0, h->nCentres + 1, long long);
    UP(asSS0, h->asSS0, h->nASlots + 1, int);
    UP(ssOwn, h->ssOwn, h->nSSlots, unsigned char);
    const int lb1 = d->maxLBS + 1, nK = lb1 * lb1;
    int rc_ = ensure(&u.ccTri, (size_t)nc * h->nCentres * sizeof(int));
    if (!rc_) rc_ = ensure(&u.ccPair, (size_t)nc * h->nCentres * sizeof(int));
    if (!rc_) rc_ = ensure(&u.pairCnt, ((size_t)h->cePair0[h->nCentres] * nK + 1) * sizeof(int2));
    if (!rc_) rc_ = ensure(&u.meta, sizeof(EnumMeta));
    if (!rc_) rc_ = ensure(&u.clsFirst, (size_t)(nc + 1) * sizeof(int));
    if (!rc_) rc_ = ensure(&u.clsWork, (size_t)(nc + 1) * sizeof(long long));
    if (!rc_) rc_ = ensure(&u.clsElem, (size_t)(nc + 1) * sizeof(long long));
    if (!rc_) rc_ = ensure(&u.clsOutElem, (size_t)(nc + 1) * sizeof(long long));
    if (!rc_) rc_ = ensure(&u.clsPairBase, (size_t)(nc + 1) * sizeof(long long));
    if (!rc_) rc_ = ensure(&u.clsQBase, (size_t)(nc + 1) * sizeof(long long));
    if (rc_) return rc_;
    EnumIn &in = u.ein;
    in.nCentres = h->nCentres; in.nc = nc; in.lb1 = lb1;
    in.ceAS0 = (const int *)u.ceAS0.p; in.cePair0 = (const long long *)u.cePair0.p; in.asSS0 = (const int *)u.asSS0.p;
    in.asType = (const int *)u.asType.p; in.ssShell = (const int *)u.ssShell.p; in.ssStart = (const int *)u.ssStart.p;
    in.ssEnd = (const int *)u.ssEnd.p; in.ssOwn = (const unsigned char *)u.ssOwn.p;
    in.ccTri = (int *)u.ccTri.p; in.ccPair = (int *)u.ccPair.p; in.pairCnt = (int2 *)u.pairCnt.p;
    int maxS = 0; /* most shell slots of one centre: shared memory of the enumeration kernels */
    for (int i = 0; i < h->nCentres; i++) {
      const int n = h->asSS0[h->ceAS0[i + 1]] - h->asSS0[h->ceAS0[i]];
      if (n > maxS) maxS = n;
    }
    u.enumSmem = (size_t)2 * maxS * sizeof(unsigned);
    if (u.enumSmem > 48 * 1024) {
      snprintf(g_err, sizeof(g_err), "libecp_b200: %d shell slots around one centre exceed the enumeration kernels' shared memory", maxS);
      return -1;
    }
    k_enum_count<<<h->nCentres, 256, u.enumSmem, st>>>(d->t, in);
    k_enum_scan<<<1, 128, 0, st>>>(d->t, in, (int *)u.clsFirst.p, (long long *)u.clsWork.p, (long long *)u.clsElem.p,
                                  (long long *)u.clsOutElem.p, (long long *)u.clsPairBase.p, (long long *)u.clsQBase.p,
                                  (EnumMeta *)u.meta.p);
    CK(cudaGetLastError());
    /* read back: EnumMeta | clsFirst (int) | five long long prefix arrays */
    if (!d->enumHost[slot] || d->enumHost_nc < nc) {
      for (int k = 0; k < 2; k++) { /* page-locked blocks from the process-wide cache (pinning costs ~0.3 ms) */
        if (d->enumHost[k]) ecpdev_pinned_free(d->enumHost[k]);
        d->enumHost[k] = ecpdev_pinned_alloc(sizeof(EnumMeta) + (size_t)(nc + 2) * 6 * sizeof(long long));
        if (!d->enumHost[k]) return -1;
      }
      d->enumHost_nc = nc;
    }
    char *hp = (char *)d->enumHost[slot];
    const size_t stride = (size_t)(nc + 2) * sizeof(long long);
    CK(cudaMemcpyAsync(hp, u.meta.p, sizeof(EnumMeta), cudaMemcpyDeviceToHost, st));
    hp += sizeof(EnumMeta);
    CK(cudaMemcpyAsync(hp, u.clsFirst.p, (nc + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hp + stride, u.clsWork.p, (nc + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hp + 2 * stride, u.clsElem.p, (nc + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hp + 3 * stride, u.clsOutElem.p, (nc + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hp + 4 * stride, u.clsPairBase.p, (nc + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hp + 5 * stride, u.clsQBase.p, (nc + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(d->evUp[slot], st));
    d->upBatch[slot] = h;
    return 0;
  }
  UP(trA, h->trA, h->nTriples, int);
  UP(trB, h->trB, h->nTriples, int);
  if (flags & 2) UP(trOut, (const long long *)h->trOut, h->nTriples, long long);
  UP(trPair, (const long long *)h->trPair, h->nTriples, long long);
  { /* pair -> triple map: filled on the device by k_triprep */
    int rc_ = ensure(&u.prTriple, ((size_t)h->nPairs + 1) * sizeof(int));
    if (rc_) return rc_;
  }
  UP(clsFirst, h->clsFirst, nc + 1, int);
  UP(clsWork, (const long long *)h->clsWork, nc + 1, long long);
  UP(clsElem, (const long long *)h->clsElem, nc + 1, long long);
  UP(clsOutElem, (const long long *)h->clsOutElem, nc + 1, long long);
  UP(clsPairBase, (const long long *)h->clsPairBase, nc + 1, long long);
  UP(clsQBase, (const long long *)h->clsQBase, nc + 1, long long);
  CK(cudaEventRecord(d->evUp[slot], st));
  d->upBatch[slot] = h;
  return 0;
}
#undef UP
/* forget prefetched inputs (start of a pass: a batch object of an aborted pass must not be mistaken for the new one) */
extern "C" void ecpdev_invalidate_prefetch(EcpDev *d) {
  cudaSetDevice(d->device);
  cudaStreamSynchronize(d->s3);
  d->upBatch[0] = d->upBatch[1] = NULL;
}
/* called from the builder thread as soon as batch i+1 is built: its H2D overlaps the kernels of batch i */
extern "C" int ecpdev_prefetch_batch(EcpDev *d, const EcpBatch *h, int flags, int slot) {
  CK(cudaSetDevice(d->device));
  if (h->nTriples == 0 && !h->devEnum) return 0;
  slot &= 1;
  d->up[slot].tablesDone = 0;
  int rc = upload_set(d, h, flags, slot, d->s3);
  if (rc) return rc;
  if (d->tabPrefetch && h->nSSlots > 0) {
    /* the next batch's tables beside the running batch: the table kernels are latency bound (k_Ftab2: 21 % of the issue
     * slots) and at 8 GPUs - where every rank tabulates F for nearly all shells - they were 19 % of a rank's pass */
    rc = launch_tables(d, h, slot, d->s3);
    if (rc) return rc;
    d->up[slot].tablesDone = 1;
    CK(cudaEventRecord(d->evTab[slot], d->s3));
  }
  return 0;
}
#define SCRATCH(buf, field, n, T)                                   \
  do {                                                              \
    int rc_ = ensure(&d->buf, ((n) ? (n) : 1) * sizeof(T));         \
    if (rc_) return rc_;                                            \
    B.field = (T *)d->buf.p;                                        \
  } while (0)

static inline unsigned nblk(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

/* small-grid + large-grid type-1 kernels for one LAB value; each launch group gets its own work and failure counters.
 * Persistent groups: the grid is sized to the device (resident blocks of this instantiation), not to the pair count. */
template <int LAB>
static void launch_type1_t(EcpDev *d, const T1Segs &sg, long long listOff, int slot) {
  const long long n = sg.prefix[sg.nseg];
  int *cnt = (int *)d->t1count.p + slot;
  int *work = (int *)d->t1work.p + 2 * slot;
  int *list = (int *)d->t1list.p + listOff;
  unsigned long long *mask = (unsigned long long *)d->t1mask.p;
  const int block = d->t1block;
  const size_t smem = t1_smem_bytes(LAB, block);
  /* resident blocks per SM, per device (the shared-memory attribute is per device too) and block size 32/64/96/128 */
  static int occS_[ECP_MAXDEV][5] = {{0}}, occL_[ECP_MAXDEV][5] = {{0}};
  int *occS = occS_[d->device < ECP_MAXDEV ? d->device : 0], *occL = occL_[d->device < ECP_MAXDEV ? d->device : 0];
  const int bi = block / 32;
  if (!occS[bi] || d->device >= ECP_MAXDEV) {
    cudaFuncSetAttribute(k_type1S<LAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_type1L<LAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occS[bi], k_type1S<LAB>, block, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occL[bi], k_type1L<LAB>, block, smem);
    if (occS[bi] < 1) occS[bi] = 1;
    if (occL[bi] < 1) occL[bi] = 1;
  }
  const long long groupsPerBlock = block / 8;
  const long long wantS = (n + groupsPerBlock - 1) / groupsPerBlock;
  const long long capS = (long long)d->nSM * occS[bi], capL = (long long)d->nSM * occL[bi];
  if (d->t1legacy) {
    k_type1S<LAB><<<(unsigned)(wantS < capS ? wantS : capS), block, smem, d->s2>>>(d->t, d->b, sg, work, cnt, list, mask, NULL,
                                                                               NULL, NULL, NULL);
  } else {
    /* first 16 slots of every pair as a block-wide wave, then the open pairs level-wise (ecp_type1.cuh) */
    static int attrA_[ECP_MAXDEV] = {0};
    const size_t smemA = t1a_smem_bytes(LAB);
    if (d->device >= ECP_MAXDEV || !attrA_[d->device]) {
      cudaFuncSetAttribute(k_type1A<LAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA);
      if (d->device < ECP_MAXDEV) attrA_[d->device] = 1;
    }
    int *scnt = (int *)d->t1scount.p + slot;
    int *slist = (int *)d->t1surv.p + listOff;
    unsigned long long *smask = (unsigned long long *)d->t1smask.p;
    const int PB = T1ACfg<LAB>::PB;
    k_type1A<LAB><<<(unsigned)((n + PB - 1) / PB), 256, smemA, d->s2>>>(d->t, d->b, sg, scnt, slist, smask, (double *)d->t1state.p);
    k_type1S<LAB><<<(unsigned)(wantS < capS ? wantS : capS), block, smem, d->s2>>>(d->t, d->b, sg, work, cnt, list, mask, scnt,
                                                                               slist, smask, (const double *)d->t1state.p);
  }
  /* the number of failed pairs is only known on the device: at most one resident wave, blocks without work leave */
  const long long wantL = (wantS + 3) / 4 > 1 ? (wantS + 3) / 4 : 1;
  k_type1L<LAB><<<(unsigned)(wantL < capL ? wantL : capL), block, smem, d->s2>>>(d->t, d->b, work + 1, cnt, list, mask,
                                                                             d->b.counters + 2);
}
static void launch_type1(EcpDev *d, int lab, const T1Segs &sg, long long listOff) {
  const int slot = d->launchSeq++;
  switch (lab) {
    case 0: launch_type1_t<0>(d, sg, listOff, slot); break;
    case 1: launch_type1_t<1>(d, sg, listOff, slot); break;
    case 2: launch_type1_t<2>(d, sg, listOff, slot); break;
    case 3: launch_type1_t<3>(d, sg, listOff, slot); break;
    case 4: launch_type1_t<4>(d, sg, listOff, slot); break;
    case 5: launch_type1_t<5>(d, sg, listOff, slot); break;
    case 6: launch_type1_t<6>(d, sg, listOff, slot); break;
    case 7: launch_type1_t<7>(d, sg, listOff, slot); break;
    case 8: launch_type1_t<8>(d, sg, listOff, slot); break;
    case 9: launch_type1_t<9>(d, sg, listOff, slot); break;
    default: launch_type1_t<10>(d, sg, listOff, slot); break;
  }
}


/* type-2 large-grid fallback as level waves (ecp_waves.cuh) on stream s1.  The host reads three counters back per wave
 * (items -> units / states, then open units and value space of the next wave) to size the next launches; the type-1
 * kernels of the batch are already queued on the other stream and keep the GPU busy meanwhile. */
static int run_fallback_waves(EcpDev *d, int km, long long *launches) {
  const DevT &t = d->t;
  DevB &B = d->b;
  cudaStream_t s = d->s1;
  int hc[4];
  CK(cudaMemcpyAsync(hc, B.counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int nItems = hc[0];
  if (nItems <= 0) return 0;
  g_allocStream = s;
  int rc = ensure(&d->fbwItems, (size_t)nItems * sizeof(FbwItem));
  if (!rc) rc = ensure(&d->fbwCtr, (FBW_CTR + FBW_NBUCKET) * (sizeof(unsigned long long) + sizeof(int)));
  if (rc) return rc;
  unsigned long long *ctr = (unsigned long long *)d->fbwCtr.p;
  int *bucketOff = (int *)(ctr + FBW_CTR + FBW_NBUCKET);
  CK(cudaMemsetAsync(ctr, 0, (FBW_CTR + FBW_NBUCKET) * sizeof(unsigned long long), s));
  k_fbw_count<<<nblk(nItems, 128), 128, 0, s>>>(t, B, (FbwItem *)d->fbwItems.p, ctr);
  unsigned long long hctr[FBW_CTR + FBW_NBUCKET];
  CK(cudaMemcpyAsync(hctr, ctr, sizeof(hctr), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  {
    int hoff[FBW_NBUCKET], acc = 0;
    for (int k = 0; k < FBW_NBUCKET; k++) {
      hoff[k] = acc;
      acc += (int)hctr[FBW_CTR + k];
    }
    CK(cudaMemcpyAsync(bucketOff, hoff, sizeof(hoff), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s)); /* hoff is on the stack */
  }
  const long long nUnits = (long long)hctr[0], nStates = (long long)hctr[1], nQd = (long long)hctr[2];
  if (nUnits <= 0 || nStates <= 0) return 0;
  if (nUnits > 0x7fffffffLL) {
    snprintf(g_err, sizeof(g_err), "fallback: too many units in one batch");
    return -1;
  }
  rc = ensure(&d->fbwUnits, (size_t)nUnits * sizeof(FbwUnit));
  if (!rc) rc = ensure(&d->fbwQd, (size_t)(nQd + 1) * sizeof(FbwQ));
  if (!rc) rc = ensure(&d->fbwSI, (size_t)nStates * sizeof(double));
  if (!rc) rc = ensure(&d->fbwSP, (size_t)nStates * sizeof(double));
  if (!rc) rc = ensure(&d->fbwSQ, (size_t)nStates * sizeof(double));
  if (!rc) rc = ensure(&d->fbwRes, (size_t)nStates * sizeof(double));
  if (!rc) rc = ensure(&d->fbwOpenFlag, (size_t)nStates);
  if (!rc) rc = ensure(&d->fbwVals, (size_t)nStates * 32 * sizeof(double));
  if (!rc) rc = ensure(&d->fbwListA, (size_t)nUnits * sizeof(FbwOpen));
  if (!rc) rc = ensure(&d->fbwListB, (size_t)nUnits * sizeof(FbwOpen));
  if (rc) return rc;
  const FbwUnit *units = (const FbwUnit *)d->fbwUnits.p;
  const FbwQ *qd = (const FbwQ *)d->fbwQd.p;
  double *sI = (double *)d->fbwSI.p, *sP = (double *)d->fbwSP.p, *sQ = (double *)d->fbwSQ.p, *sRes = (double *)d->fbwRes.p;
  unsigned char *sOpen = (unsigned char *)d->fbwOpenFlag.p;
  k_fbw_units<<<nblk(nItems, 128), 128, 0, s>>>(t, B, (const FbwItem *)d->fbwItems.p, bucketOff, (FbwUnit *)d->fbwUnits.p, (FbwQ *)d->fbwQd.p);
  *launches += 2;
  FbwOpen *cur = NULL, *nxt = (FbwOpen *)d->fbwListA.p, *other = (FbwOpen *)d->fbwListB.p;
  long long nOpen = nUnits;
  for (int lev = 4; lev <= t.largeLevels && nOpen > 0; lev++) {
    const int S = lev == 4 ? 32 : (1 << lev), slot0 = lev == 4 ? 0 : (1 << lev);
    const long long nWarps = nOpen * (S >> 5);
    if (km <= 6)
      k_fbw_eval<6><<<nblk(nWarps, 4), 128, 0, s>>>(t, B, units, qd, cur, nWarps, S, slot0, (double *)d->fbwVals.p);
    else
      k_fbw_eval<10><<<nblk(nWarps, 4), 128, 0, s>>>(t, B, units, qd, cur, nWarps, S, slot0, (double *)d->fbwVals.p);
    CK(cudaMemsetAsync(ctr + 3, 0, 2 * sizeof(unsigned long long), s));
    k_fbw_book<<<nblk(nOpen * 8, 128), 128, 0, s>>>(t, units, cur, (int)nOpen, lev, (const double *)d->fbwVals.p, sI, sP, sQ, sRes, sOpen,
                                                   nxt, ctr, B.counters + 3);
    *launches += 2;
    CK(cudaGetLastError());
    if (lev == t.largeLevels) break;
    CK(cudaMemcpyAsync(hctr, ctr, FBW_CTR * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    nOpen = (long long)hctr[3];
    if (nOpen > 0) {
      rc = ensure(&d->fbwVals, (size_t)hctr[4] * sizeof(double)); /* the values of the wave just booked are no longer needed */
      if (rc) return rc;
    }
    cur = nxt;
    FbwOpen *tmp = nxt == (FbwOpen *)d->fbwListA.p ? other : (FbwOpen *)d->fbwListA.p;
    nxt = tmp;
  }
  k_fbw_final<<<nblk(nItems, 4), 128, 0, s>>>(B, (const FbwItem *)d->fbwItems.p, qd, sRes);
  *launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int ecpdev_run_batch(EcpDev *d, EcpBatch *h, int flags, int slot, double *hostBlocks, EcpDevStats *st) {
  CK(cudaSetDevice(d->device));
  DevB &B = d->b;
  g_allocStream = d->s1;
  d->s2 = d->serial ? d->s1 : d->s2real;
  const int nc = d->nClasses;
  if (st) memset(st, 0, sizeof(*st));
  d->batchH2D = 0;
  if (h->nTriples == 0 && !h->devEnum) return 0;
  const bool trace = getenv("LIBECP_B200_TRACE") != NULL;
  const double tr0 = omp_get_wtime();
  slot &= 1;
  if (d->upBatch[slot] != h) { /* not prefetched (first batch, or no helper thread): copy now, on the compute stream */
    d->up[slot].tablesDone = 0;
    int rc_ = upload_set(d, h, flags, slot, d->s1);
    if (rc_) return rc_;
  } else {
    CK(cudaStreamWaitEvent(d->s1, d->evUp[slot], 0));
  }
  d->upBatch[slot] = NULL; /* consumed: the set may be refilled once this batch is through */
  d->batchH2D = d->upBytes[slot];
  g_allocStream = d->s1;
  if (h->devEnum) {
    /* sizes and class prefixes of the device-enumerated batch (counted and scanned behind the upload, ecp_enum.cuh);
     * they land in the host arrays the batch view points at, and everything below sizes itself from them as before */
    CK(cudaEventSynchronize(d->evUp[slot]));
    const char *hp = (const char *)d->enumHost[slot];
    const EnumMeta *m = (const EnumMeta *)hp;
    hp += sizeof(EnumMeta);
    const size_t stride = (size_t)(nc + 2) * sizeof(long long);
    if (m->nTriples > 0x7fffffffLL || m->nPairs > 0x7fffffffLL) {
      snprintf(g_err, sizeof(g_err), "libecp_b200: batch of %lld triples / %lld primitive pairs exceeds 2^31", m->nTriples, m->nPairs);
      return -1;
    }
    h->nTriples = (int)m->nTriples;
    h->nPairs = m->nPairs;
    h->tTotal = m->tTotal;
    h->gTotal = m->gTotal;
    h->qTotal = m->qTotal;
    h->rshTotal = m->qTotal;
    h->outTotal = m->outTotal;
    memcpy((void *)h->clsFirst, hp, (nc + 1) * sizeof(int));
    memcpy((void *)h->clsWork, hp + stride, (nc + 1) * sizeof(long long));
    memcpy((void *)h->clsElem, hp + 2 * stride, (nc + 1) * sizeof(long long));
    memcpy((void *)h->clsOutElem, hp + 3 * stride, (nc + 1) * sizeof(long long));
    memcpy((void *)h->clsPairBase, hp + 4 * stride, (nc + 1) * sizeof(long long));
    memcpy((void *)h->clsQBase, hp + 5 * stride, (nc + 1) * sizeof(long long));
    if (h->nTriples == 0) return 0;
    EcpDev::UpSet &u = d->up[slot];
    int rc_ = ensure(&u.trA, (size_t)h->nTriples * sizeof(int));
    if (!rc_) rc_ = ensure(&u.trB, (size_t)h->nTriples * sizeof(int));
    if (!rc_) rc_ = ensure(&u.trPair, (size_t)h->nTriples * sizeof(long long));
    if (!rc_) rc_ = ensure(&u.prTriple, ((size_t)h->nPairs + 1) * sizeof(int));
    if (rc_) return rc_;
    CK(cudaEventRecord(d->ev[0], d->s1)); /* the batch's device time includes the fill (count / scan ran behind the previous batch) */
    /* fill + k_triprep on the second stream: the per-centre tables (k_atomslot, k_omegaX, k_Ftab) only need the slots and
     * run beside them on the first */
    cudaStream_t sE = d->serial ? d->s1 : d->s2real;
    if (sE != d->s1) CK(cudaStreamWaitEvent(sE, d->ev[0], 0));
    k_enum_fill<<<h->nCentres, 256, u.enumSmem, sE>>>(d->t, u.ein, (const int *)u.clsFirst.p, (const long long *)u.clsPairBase.p,
                                               (int *)u.trA.p, (int *)u.trB.p, (long long *)u.trPair.p);
    CK(cudaGetLastError());
  }
  B.nASlots = h->nASlots;
  B.nSSlots = h->nSSlots;
  B.nTriples = h->nTriples;
  B.nPairs = h->nPairs;
  {
    const EcpDev::UpSet &u = d->up[slot];
    B.asAtom = (const int *)u.asAtom.p; B.asType = (const int *)u.asType.p; B.asR = (const double *)u.asR.p;
    B.asOmOff = (const long long *)u.asOmOff.p; B.ssShell = (const int *)u.ssShell.p;
    B.ssASlot = (const int *)u.ssASlot.p; B.ssStart = (const int *)u.ssStart.p; B.ssEnd = (const int *)u.ssEnd.p;
    B.ssFOff = (const long long *)u.ssFOff.p; B.trA = (const int *)u.trA.p; B.trB = (const int *)u.trB.p;
    B.trOut = (const long long *)u.trOut.p; B.trPair = (const long long *)u.trPair.p;
    B.prTriple = (const int *)u.prTriple.p; B.clsFirst = (const int *)u.clsFirst.p;
    B.clsWork = (const long long *)u.clsWork.p; B.clsElem = (const long long *)u.clsElem.p;
    B.clsOutElem = (const long long *)u.clsOutElem.p; B.clsPairBase = (const long long *)u.clsPairBase.p;
    B.clsQBase = (const long long *)u.clsQBase.p;
  }
  const bool tablesDone = d->up[slot].tablesDone != 0; /* computed behind the previous batch (ecpdev_prefetch_batch) */
  d->up[slot].tablesDone = 0;
  if (!tablesDone) { /* buffers of the set; the kernels are launched below */
    EcpDev::UpSet &u = d->up[slot];
    int rc_ = ensure(&u.rshX, ((size_t)h->nASlots * RSHX_STRIDE + 1) * sizeof(double));
    if (!rc_) rc_ = ensure(&u.uspX, ((size_t)h->nASlots * USPX_STRIDE + 1) * sizeof(double));
    if (!rc_) rc_ = ensure(&u.omX, ((size_t)h->omTotal + 1) * sizeof(double));
    if (!rc_) rc_ = ensure(&u.F, ((size_t)h->fRows * ECP_SMALL_SLOTS + 1) * sizeof(double));
    if (rc_) return rc_;
  } else {
    CK(cudaStreamWaitEvent(d->s1, d->evTab[slot], 0));
  }
  B.rshX = (double *)d->up[slot].rshX.p;
  B.uspX = (double *)d->up[slot].uspX.p;
  B.omX = (double *)d->up[slot].omX.p;
  B.F = (double *)d->up[slot].F.p;
  SCRATCH(T, T, (size_t)h->tTotal, double);
  SCRATCH(gamma, gamma, (size_t)h->gTotal, double);
  SCRATCH(chi, chi, (size_t)h->gTotal, double);
  SCRATCH(Q, Q, (size_t)h->qTotal, double);
  SCRATCH(rshP, rshP, (size_t)h->rshTotal, double);
  SCRATCH(sP, sP, (size_t)h->nPairs, double);
  SCRATCH(blocks, blocks, (flags & 2) ? (size_t)h->outTotal : 1, double);
  SCRATCH(tfail, tfail, (size_t)h->tTotal, unsigned char);
  SCRATCH(tflags, tflags, (size_t)h->nTriples, int);
  SCRATCH(items, items, (size_t)h->nTriples * 8, int);
  SCRATCH(counters, counters, 16, int);
  {
    int rc_ = ensure(&d->t1list, ((size_t)h->nPairs + 1) * sizeof(int));
    if (rc_) return rc_;
    rc_ = ensure(&d->t1mask, ((size_t)h->nPairs + 1) * sizeof(unsigned long long));
    if (rc_) return rc_;
    rc_ = ensure(&d->t1count, 256 * sizeof(int));
    if (rc_) return rc_;
    CK(cudaMemsetAsync(d->t1count.p, 0, 256 * sizeof(int), d->s1));
    if (!d->t1legacy) {
      rc_ = ensure(&d->t1surv, ((size_t)h->nPairs + 1) * sizeof(int));
      if (rc_) return rc_;
      rc_ = ensure(&d->t1smask, ((size_t)h->nPairs + 1) * sizeof(unsigned long long));
      if (rc_) return rc_;
      rc_ = ensure(&d->t1scount, 256 * sizeof(int));
      if (rc_) return rc_;
      CK(cudaMemsetAsync(d->t1scount.p, 0, 256 * sizeof(int), d->s1));
      /* (I, p, q) of every quadrature of the pairs of ONE launch (launches of a batch follow each other on one stream) */
      size_t need = 0;
      for (int lab = 0; lab <= 2 * d->maxLBS; lab++) {
        long long np = 0;
        for (int c = 0; c < nc; c++)
          if (d->hClsLa[c] + d->hClsLb[c] == lab) np += h->clsPairBase[c + 1] - h->clsPairBase[c];
        const size_t bytes = (size_t)np * T1_NQ(lab) * 3 * sizeof(double);
        if (bytes > need) need = bytes;
      }
      rc_ = ensure(&d->t1state, need + 64);
      if (rc_) return rc_;
    }
    rc_ = ensure(&d->t1work, 512 * sizeof(int));
    if (rc_) return rc_;
    CK(cudaMemsetAsync(d->t1work.p, 0, 512 * sizeof(int), d->s1));
    rc_ = ensure(&d->t1rec, ((size_t)h->nPairs + 1) * sizeof(T1Rec));
    if (rc_) return rc_;
    B.t1rec = (T1Rec *)d->t1rec.p;
    rc_ = ensure(&d->trirec, ((size_t)h->nTriples + 1) * sizeof(TriRec));
    if (rc_) return rc_;
    B.trirec = (TriRec *)d->trirec.p;
    d->launchSeq = 0;
  }
  B.dbg = NULL;
  if (d->tails) {
    int rc_ = ensure(&d->dbgBuf, (size_t)16 * DBG_STRIDE * 2 * sizeof(unsigned long long));
    if (rc_) return rc_;
    CK(cudaMemsetAsync(d->dbgBuf.p, 0, (size_t)16 * DBG_STRIDE * 2 * sizeof(unsigned long long), d->s1));
    B.dbg = (unsigned long long *)d->dbgBuf.p;
  }
  if ((flags & 1) && !d->matrix) {
    int rc = ecpdev_matrix_begin(d, NULL, -1);
    if (rc) return rc;
  }
  B.matrix = d->matrix;
  d->lastSizes[0] = (size_t)h->fRows * ECP_SMALL_SLOTS;
  d->lastSizes[1] = (size_t)h->omTotal;
  d->lastSizes[2] = (size_t)h->tTotal;
  d->lastSizes[3] = (size_t)h->gTotal;
  d->lastSizes[4] = (size_t)h->qTotal;
  d->lastSizes[5] = (size_t)h->outTotal;
  CK(cudaMemsetAsync(B.counters, 0, 16 * sizeof(int), d->s1));
  CK(cudaMemsetAsync(B.tflags, 0, (size_t)h->nTriples * sizeof(int), d->s1));
  CK(cudaMemsetAsync(B.Q, 0, (size_t)h->qTotal * sizeof(double), d->s1));
  const DevT &t = d->t;
  long long launches = 0;
  const double tr1 = omp_get_wtime();
  if (!h->devEnum) CK(cudaEventRecord(d->ev[0], d->s1)); /* device-enumerated batch: recorded before k_enum_fill */
  /* per-centre tables */
  if (h->devEnum && !d->serial) {
    CK(cudaEventRecord(d->ev[11], d->s1)); /* the scratch k_triprep writes was (re)allocated in s1's order */
    CK(cudaStreamWaitEvent(d->s2real, d->ev[11], 0));
    k_triprep<<<nblk(h->nTriples, 128), 128, 0, d->s2real>>>(t, B);
    CK(cudaEventRecord(d->ev[10], d->s2real));
  } else
    k_triprep<<<nblk(h->nTriples, 128), 128, 0, d->s1>>>(t, B);
  if (!tablesDone) k_atomslot<<<nblk(h->nASlots, 128), 128, 0, d->s1>>>(t, B);
  /* the type-1 chain (second stream) needs the triples and the atom slots, not Omega_X or F: it starts here and runs
   * beside the two table kernels, which are latency bound (k_Ftab2: 21 % of the issue slots) - at 8 GPUs the tables of a
   * rank (every rank tabulates F for nearly all shells) were 19 % of its pass */
  if (!d->serial) CK(cudaEventRecord(d->ev[11], d->s1));
  if (!tablesDone) {
    k_omegaX<<<h->nASlots, 256, 0, d->s1>>>(t, B);
    if (d->ftabCompact) { /* window-only tabulation into a cleared table (k_Ftab2) */
      CK(cudaMemsetAsync(B.F, 0, (size_t)h->fRows * ECP_SMALL_SLOTS * sizeof(double), d->s1));
      if (t.maxLECP - 1 + d->maxLBS <= 6)
        k_Ftab2<6><<<h->nSSlots, 128, 0, d->s1>>>(t, B);
      else
        k_Ftab2<KM><<<h->nSSlots, 128, 0, d->s1>>>(t, B);
    } else if (t.maxLECP - 1 + d->maxLBS <= 6)
      k_Ftab<6><<<h->nSSlots, ECP_SMALL_SLOTS, 0, d->s1>>>(t, B);
    else
      k_Ftab<KM><<<h->nSSlots, ECP_SMALL_SLOTS, 0, d->s1>>>(t, B);
  }
  launches += 4;
  if (h->devEnum && !d->serial) CK(cudaStreamWaitEvent(d->s1, d->ev[10], 0)); /* triples and their records are in place */
  CK(cudaEventRecord(d->ev[1], d->s1));
  /* type 1 on the second stream, after the uploads / triples / atom slots (serial mode: after the tables, same stream) */
  CK(cudaStreamWaitEvent(d->s2, d->serial ? d->ev[1] : d->ev[11], 0));
  CK(cudaEventRecord(d->ev[6], d->s2));
  k_t1prep<<<nblk(h->nPairs, 128), 128, 0, d->s2>>>(t, B);
  { /* type-1 radial integrals: per LAB = la+lb, one thread per primitive pair (ecp_type1.cuh) */
    long long listOff = 0;
    for (int lab = 0; lab <= 2 * d->maxLBS; lab++) {
      T1Segs sg;
      sg.nseg = 0;
      sg.prefix[0] = 0;
      for (int c = 0; c < nc; c++) {
        if (d->hClsLa[c] + d->hClsLb[c] != lab || h->clsFirst[c + 1] == h->clsFirst[c]) continue;
        const long long p0 = h->clsPairBase[c], p1 = h->clsPairBase[c + 1]; /* primitive pairs of the class */
        if (sg.nseg == T1_MAXSEG) {
          launch_type1(d, lab, sg, listOff);
          launches += 2;
          listOff += sg.prefix[sg.nseg];
          sg.nseg = 0;
        }
        sg.start[sg.nseg] = p0;
        sg.prefix[sg.nseg + 1] = sg.prefix[sg.nseg] + (p1 - p0);
        sg.nseg++;
      }
      if (sg.nseg) {
        launch_type1(d, lab, sg, listOff);
        launches += 2;
        listOff += sg.prefix[sg.nseg];
      }
    }
  }
  CK(cudaEventRecord(d->ev[7], d->s2));
  { /* 8 lanes per triple; packed R[N][(l,m)] of the largest la+lb per group in shared memory */
    const int labMax = 2 * d->maxLBS, rStride = (labMax + 1) * (labMax + 2) * (labMax + 3) / 6 + 1;
    k_chi<<<nblk((long long)h->nTriples * 8, 128), 128, (size_t)16 * rStride * sizeof(double), d->s2>>>(t, B, rStride);
  }
  CK(cudaEventRecord(d->ev[8], d->s2));
  launches += 2;
  /* type 2 */
  const long long nWork = h->clsWork[nc];
  CK(cudaEventRecord(d->ev[9], d->s1)); /* in serial mode the type-1 kernels above sit on this stream too */
  if (nWork > 0) {
    /* survivors of the first levels: room for a quarter of the quadratures (an overflowing thread finishes in place) */
    const long long cap64 = d->survCapEnv > 0 ? d->survCapEnv : nWork / 4 + 1024;
    const int survCap = cap64 > 0x7fffffff ? 0x7fffffff : (int)cap64;
    const int lim = (d->fastLim >= ECP_SMALL_LEVELS) ? ECP_SMALL_LEVELS : d->fastLim;
    int rc_ = ensure(&d->fastSurv, (size_t)survCap * sizeof(FastSurv));
    if (rc_) return rc_;
    k_fastT<<<nblk(nWork, 128), 128, 0, d->s1>>>(t, B, nWork, lim, (FastSurv *)d->fastSurv.p, survCap);
    launches++;
    if (lim < ECP_SMALL_LEVELS) {
      /* the survivor count lives on the device: the grid covers the list capacity, surplus blocks leave at once */
      if (d->fastUnroll == 4)
        k_fastT2<4><<<nblk(survCap, 128), 128, 0, d->s1>>>(t, B, lim, (const FastSurv *)d->fastSurv.p, survCap);
      else if (d->fastUnroll == 2)
        k_fastT2<2><<<nblk(survCap, 128), 128, 0, d->s1>>>(t, B, lim, (const FastSurv *)d->fastSurv.p, survCap);
      else
        k_fastT2<1><<<nblk(survCap, 128), 128, 0, d->s1>>>(t, B, lim, (const FastSurv *)d->fastSurv.p, survCap);
      launches++;
    }
  }
  CK(cudaEventRecord(d->ev[2], d->s1));
  if (nWork > 0) {
    {
      /* Bessel order bound of the instantiation: max(2 maxLBS, maxLBS + maxLECP - 1) */
      const int km = (2 * d->maxLBS > d->maxLBS + t.maxLECP - 1) ? 2 * d->maxLBS : d->maxLBS + t.maxLECP - 1;
      if (d->fbWaves) {
        int rc_ = run_fallback_waves(d, km, &launches);
        if (rc_) return rc_;
      } else { /* persistent 8-lane groups */
        const int block = d->fbblock;
        void (*kern)(DevT, DevB);
        size_t smem;
        int slot;
        if (km <= 6) {
          smem = fb_smem_bytes<6>(block);
          if (d->fbminb == 4) { kern = k_fallbackG<6, 4>; slot = 0; } else { kern = k_fallbackG<6, 3>; slot = 1; }
        } else {
          smem = fb_smem_bytes<10>(block);
          kern = k_fallbackG<10, 2>;
          slot = 2;
        }
        static int occ_[ECP_MAXDEV][3][5] = {{{0}}};
        int (*occ)[5] = occ_[d->device < ECP_MAXDEV ? d->device : 0];
        const int bi = block / 32;
        if (!occ[slot][bi] || d->device >= ECP_MAXDEV) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(km <= 6 ? fb_smem_bytes<6>(128) : fb_smem_bytes<10>(128)));
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[slot][bi], kern, block, smem);
          if (occ[slot][bi] < 1) occ[slot][bi] = 1;
        }
        const int per = (d->fbocc > 0 && d->fbocc < occ[slot][bi]) ? d->fbocc : occ[slot][bi];
        kern<<<d->nSM * per, block, smem, d->s1>>>(t, B);
      }
    }
    launches++;
  }
  CK(cudaEventRecord(d->ev[3], d->s1));
  for (int c = 0; c < nc; c++) { /* one launch per class: the (lambda1, lambda2) tile is sized by (la, lb) */
    const long long ne = h->clsElem[c + 1] - h->clsElem[c];
    if (ne <= 0) continue;
    const int la_ = d->hClsLa[c], lb_ = d->hClsLb[c], Lc_ = d->hClsL[c];
    const int cda_ = (la_ + 1) * (la_ + 2) * (la_ + 3) / 6, cdb_ = (lb_ + 1) * (lb_ + 2) * (lb_ + 3) / 6;
    const int E_ = cda_ * cdb_;
    const int SA = (la_ + Lc_) * Lc_ * Lc_ * cda_, SB = (lb_ + Lc_) * Lc_ * Lc_ * cdb_;
    const long long ntri = h->clsFirst[c + 1] - h->clsFirst[c];
    const Link4Kernel k4 = (d->linkMode == 4 && la_ < 4 && lb_ < 4 && Lc_ <= LINK4_MAXL) ? g_link4Kernels[la_][lb_][Lc_] : NULL;
    if (k4) { /* specialised shared-memory kernel for the large classes */
      const int slack = (la_ * Lc_ * Lc_ * cda_ > lb_ * Lc_ * Lc_ * cdb_) ? la_ * Lc_ * Lc_ * cda_ : lb_ * Lc_ * Lc_ * cdb_;
      const int tsz = LINK4_RMAX * (d->hClsNq[c] + 1); /* T rows of a run + the zero row */
      const int smTotal = 2 * (SA + SB + tsz) + slack; /* doubles: two copies of the slices and of T, slack */
      const size_t smem = (size_t)smTotal * sizeof(double);
      if (smem <= 200 * 1024) {
        static unsigned char attr_[ECP_MAXDEV][4][4][LINK4_MAXL + 1];
        unsigned char *at = &attr_[d->device < ECP_MAXDEV ? d->device : 0][la_][lb_][Lc_];
        if (!*at || d->device >= ECP_MAXDEV) {
          cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          *at = 1;
        }
        /* triples per block: long enough to reuse a staged Omega_A over the runs that share atom A and to pay for the
         * per-block set-up, short enough that a small class still fills the device */
        long long tpb = d->linkTpb > 0 ? d->linkTpb : ntri / (4LL * d->nSM);
        if (tpb < 1) tpb = 1;
        if (d->linkTpb <= 0 && tpb > 16) tpb = 16;
        if (tpb > LINK4_TPB_MAX) tpb = LINK4_TPB_MAX;
        k4<<<nblk(ntri, (int)tpb), ((E_ + 31) / 32) * 32, smem, d->s1>>>(t, B, c, (int)tpb, tsz, smTotal);
        launches++;
        continue;
      }
    }
    g_linkKernels[la_][lb_]<<<nblk(ne, 128), 128, 0, d->s1>>>(t, B, c);
    launches++;
  }
  CK(cudaEventRecord(d->ev[4], d->s1));
  CK(cudaStreamWaitEvent(d->s1, d->ev[8], 0));
  if (d->shift2) { /* both shift passes in one kernel, blocks of consecutive triples of one class (k_shift2) */
    /* One launch per shared-memory bucket: the dynamic shared memory of a launch is that of its largest class, and with
     * a single launch the f-f classes (tens of KB per block) held every block - also those of the s and p classes that
     * make up most triples - at 4 resident blocks per SM (ncu: 24 % of the warp slots in use, profiles/r2).  Per bucket:
     * first block of every class (classes of other buckets are empty), then the chunks per block of every class. */
    enum { NBK = 3 };
    const size_t bucketCap[NBK] = {14 * 1024, 40 * 1024, 200 * 1024};
    const int stride = 2 * nc + 2;
    int clsBlk[NBK][2 * ECP_MAX_CLASSES + 2];
    size_t smemMax[NBK] = {0, 0, 0};
    for (int k = 0; k < NBK; k++) clsBlk[k][0] = 0;
    for (int c = 0; c < nc; c++) {
      const int la = d->hClsLa[c], lb = d->hClsLb[c];
      const int ntri = h->clsFirst[c + 1] - h->clsFirst[c];
      const int tpb = shift2_tpb(la, lb);
      /* chunks per block: the per-block tables are paid once per block, but a small class must still fill the device */
      long long ch = ntri / ((long long)tpb * d->nSM * 4);
      ch = ch < 1 ? 1 : (ch > SHIFT2_CH ? SHIFT2_CH : ch);
      Shift2Layout L;
      const size_t sm = shift2_smem(la, lb, d->hShTerms[la], d->hShTerms[lb], tpb, &L);
      int bk = 0;
      while (bk < NBK - 1 && sm > bucketCap[bk]) bk++;
      for (int k = 0; k < NBK; k++) {
        clsBlk[k][nc + 1 + c] = (int)ch;
        clsBlk[k][c + 1] = clsBlk[k][c] + ((k == bk) ? (int)((ntri + tpb * ch - 1) / (tpb * ch)) : 0);
      }
      if (ntri > 0 && sm > smemMax[bk]) smemMax[bk] = sm;
    }
    int rc_ = ensure(&d->clsJ, (size_t)NBK * stride * sizeof(int));
    if (rc_) return rc_;
    for (int k = 0; k < NBK; k++)
      CK(cudaMemcpyAsync((int *)d->clsJ.p + k * stride, clsBlk[k], (2 * nc + 1) * sizeof(int), cudaMemcpyHostToDevice, d->s1));
    if (smemMax[NBK - 1] > 48 * 1024 && !d->shift2Attr) {
      CK(cudaFuncSetAttribute(k_shift2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      d->shift2Attr = 1;
    }
    for (int k = 0; k < NBK; k++)
      if (clsBlk[k][nc] > 0) {
        k_shift2<<<clsBlk[k][nc], 128, smemMax[k], d->s1>>>(t, B, (const int *)d->clsJ.p + k * stride, flags);
        launches++;
      }
    launches -= 2; /* counted below */
  } else
  { /* binomial shift in two passes (ecp_shift.cuh); J[type][c1][q] per triple goes through a scratch buffer */
    long long clsJ[ECP_MAX_CLASSES + 1];
    clsJ[0] = 0;
    for (int c = 0; c < nc; c++) {
      const int la = d->hClsLa[c], lb = d->hClsLb[c];
      clsJ[c + 1] = clsJ[c] + (long long)(h->clsFirst[c + 1] - h->clsFirst[c]) * ((la + 1) * (la + 2) / 2) *
                                  ((lb + 1) * (lb + 2) * (lb + 3) / 6);
    }
    int rc_ = ensure(&d->clsJ, (nc + 1) * sizeof(long long));
    if (rc_) return rc_;
    rc_ = ensure(&d->Jbuf, (size_t)(2 * clsJ[nc] + 1) * sizeof(double));
    if (rc_) return rc_;
    CK(cudaMemcpyAsync(d->clsJ.p, clsJ, (nc + 1) * sizeof(long long), cudaMemcpyHostToDevice, d->s1));
    k_shiftJ<<<nblk(clsJ[nc], 128), 128, 0, d->s1>>>(t, B, (const long long *)d->clsJ.p, clsJ[nc], (double *)d->Jbuf.p);
    k_shiftI<<<nblk(h->clsOutElem[nc], 128), 128, 0, d->s1>>>(t, B, (const long long *)d->clsJ.p, h->clsOutElem[nc],
                                                              (const double *)d->Jbuf.p, flags);
  }
  launches += 2;
  CK(cudaEventRecord(d->ev[5], d->s1));
  CK(cudaGetLastError());
  int hc[16], hc1[256];
  CK(cudaMemcpyAsync(hc, B.counters, sizeof(hc), cudaMemcpyDeviceToHost, d->s1));
  CK(cudaMemcpyAsync(hc1, d->t1count.p, sizeof(hc1), cudaMemcpyDeviceToHost, d->s1));
  if ((flags & 2) && hostBlocks)
    CK(cudaMemcpyAsync(hostBlocks, B.blocks, (size_t)h->outTotal * sizeof(double), cudaMemcpyDeviceToHost, d->s1));
  const double tr2 = omp_get_wtime();
  /* s1 joined the type-1 stream before the shift kernels, so the end of s1 is the end of the batch.  Wait on a
   * blocking-sync event instead of spinning in cudaStreamSynchronize: on a box with 4 host cores per GPU the spinning
   * thread took a third of what the batch builder (one batch ahead, other thread) could use. */
  CK(cudaEventRecord(d->evDone, d->s1));
  CK(cudaEventSynchronize(d->evDone));
  CK(cudaStreamSynchronize(d->s2));
  if (d->tails) { /* how long do the persistent kernels run with most of their blocks already gone? */
    unsigned long long *hb = (unsigned long long *)malloc((size_t)16 * DBG_STRIDE * 2 * sizeof(unsigned long long));
    cudaMemcpy(hb, d->dbgBuf.p, (size_t)16 * DBG_STRIDE * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    for (int k = 0; k < 16; k++) {
      unsigned long long t0 = ~0ull, t1 = 0;
      int nb = 0;
      unsigned long long *e = hb + (size_t)k * DBG_STRIDE * 2;
      static unsigned long long ends[DBG_STRIDE];
      for (int i = 0; i < DBG_STRIDE; i++)
        if (e[2 * i + 1]) {
          if (e[2 * i] < t0) t0 = e[2 * i];
          if (e[2 * i + 1] > t1) t1 = e[2 * i + 1];
          ends[nb++] = e[2 * i + 1];
        }
      if (!nb) continue;
      /* time-integrated fraction of blocks still running = mean(end - t0) / (t1 - t0) */
      double sum = 0;
      for (int i = 0; i < nb; i++) sum += (double)(ends[i] - t0);
      fprintf(stderr, "[libecp_b200]   tails: kernel %2d blocks %5d span %8.1f us, mean block residency %.3f of the span\n", k, nb,
              (t1 - t0) * 1e-3, sum / nb / (double)(t1 - t0));
    }
    free(hb);
  }
  if (trace)
    fprintf(stderr, "[libecp_b200]   run_batch: alloc+H2D issue %.2f ms, launches %.2f ms, wait %.2f ms\n", 1e3 * (tr1 - tr0),
            1e3 * (tr2 - tr1), 1e3 * (omp_get_wtime() - tr2));
  if (st) {
    float ms;
    cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]); st->ms_tables = ms;
    cudaEventElapsedTime(&ms, d->ev[9], d->ev[2]); st->ms_fastT = ms;
    cudaEventElapsedTime(&ms, d->ev[2], d->ev[3]); st->ms_fallback = ms;
    cudaEventElapsedTime(&ms, d->ev[3], d->ev[4]); st->ms_link = ms;
    cudaEventElapsedTime(&ms, d->ev[6], d->ev[7]); st->ms_type1 = ms;
    cudaEventElapsedTime(&ms, d->ev[7], d->ev[8]); st->ms_chi = ms;
    cudaEventElapsedTime(&ms, d->ev[4], d->ev[5]); st->ms_shift = ms;
    cudaEventElapsedTime(&ms, d->ev[0], d->ev[5]); st->ms_total = ms;
    st->nFallbackItems = hc[0];
    st->nFastFail = hc[4];
    st->nType1Fail = 0;
    for (int i = 0; i < 256; i++) st->nType1Fail += hc1[i];
    st->nStaleCentre = hc[6];
    st->launches = launches;
    st->h2dBytes = d->batchH2D;
    st->d2hBytes = (long long)sizeof(hc) + (((flags & 2) && hostBlocks) ? (long long)h->outTotal * 8 : 0);
    st->err1 = hc[2];
    st->err2 = hc[3];
  }
  return 0;
}

/* ================================================================================================
 * C-ABI collective for a sharded, device-resident result ("NCCL all-gather over NVLink when the matrix stays
 * on the device").  Every rank holds the upper-triangle parts of its own AO rows; ecpdev_allgather packs them into the
 * rank's slot of one receive buffer, runs ONE in-place ncclAllGather (shards padded to the largest) and scatters the
 * other ranks' rows into the resident matrix - all on the handle's compute stream, no host synchronisation.
 * NCCL is bound with dlopen at first use (a single-GPU caller needs no NCCL at all): libnccl.so.2 is whatever the
 * process already loaded (e.g. the copy bundled with PyTorch) or the system library. */
#include <dlfcn.h>
struct EcpNcclId {
  char internal[128];
};
typedef int (*PfnGetUniqueId)(EcpNcclId *);
typedef int (*PfnCommInitRank)(void **, int, EcpNcclId, int);
typedef int (*PfnAllGather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*PfnCommDestroy)(void *);
typedef const char *(*PfnGetErrorString)(int);
static struct {
  void *lib;
  PfnGetUniqueId getUniqueId;
  PfnCommInitRank commInitRank;
  PfnAllGather allGather;
  PfnCommDestroy commDestroy;
  PfnGetErrorString errorString;
} g_nccl;
static std::mutex g_ncclMu;
static int nccl_bind() {
  std::lock_guard<std::mutex> lk(g_ncclMu);
  if (g_nccl.lib) return 0;
  const char *names[] = {getenv("LIBECP_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
  void *lib = NULL;
  for (int i = 0; i < 3 && !lib; i++)
    if (names[i]) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    snprintf(g_err, sizeof(g_err), "libecp_b200: cannot load NCCL (libnccl.so.2): %s", dlerror());
    return -1;
  }
  g_nccl.getUniqueId = (PfnGetUniqueId)dlsym(lib, "ncclGetUniqueId");
  g_nccl.commInitRank = (PfnCommInitRank)dlsym(lib, "ncclCommInitRank");
  g_nccl.allGather = (PfnAllGather)dlsym(lib, "ncclAllGather");
  g_nccl.commDestroy = (PfnCommDestroy)dlsym(lib, "ncclCommDestroy");
  g_nccl.errorString = (PfnGetErrorString)dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.getUniqueId || !g_nccl.commInitRank || !g_nccl.allGather || !g_nccl.commDestroy) {
    snprintf(g_err, sizeof(g_err), "libecp_b200: NCCL library lacks the expected entry points");
    return -1;
  }
  g_nccl.lib = lib;
  return 0;
}
#define NCCLCK(call)                                                                                   \
  do {                                                                                                 \
    const int r_ = (call);                                                                             \
    if (r_ != 0) {                                                                                     \
      snprintf(g_err, sizeof(g_err), "%s:%d: %s: %s", __FILE__, __LINE__, #call,                       \
               g_nccl.errorString ? g_nccl.errorString(r_) : "NCCL error");                           \
      return -1;                                                                                       \
    }                                                                                                  \
  } while (0)

extern "C" int ecpdev_comm_unique_id(void *id128) {
  if (nccl_bind()) return -1;
  NCCLCK(g_nccl.getUniqueId((EcpNcclId *)id128));
  return 0;
}
static void comm_release(EcpDev *d) {
  if (d->comm && d->commOwned && g_nccl.commDestroy) g_nccl.commDestroy(d->comm);
  d->comm = NULL;
  d->commOwned = 0;
  free(d->agCount);
  d->agCount = NULL;
  d->agFirst = NULL;
}
/* comm == NULL: create a communicator from the unique id (collective call: every rank of `world`); else adopt the
 * caller's ncclComm_t (it stays the caller's) */
extern "C" int ecpdev_comm_init(EcpDev *d, int rank, int world, const void *id128, void *comm) {
  CK(cudaSetDevice(d->device));
  if (nccl_bind()) return -1;
  comm_release(d);
  if (comm) {
    d->comm = comm;
    d->commOwned = 0;
  } else {
    EcpNcclId id;
    memcpy(&id, id128, sizeof(id));
    NCCLCK(g_nccl.commInitRank(&d->comm, world, id, rank));
    d->commOwned = 1;
  }
  d->commRank = rank;
  d->commWorld = world;
  return 0;
}
extern "C" void ecpdev_comm_destroy(EcpDev *d) {
  if (!d) return;
  cudaSetDevice(d->device);
  cudaStreamSynchronize(d->s1);
  comm_release(d);
}
/* shard layout of all ranks: rows[r] = ascending AO rows of rank r (nrows[r] of them); uploaded once per communicator */
extern "C" int ecpdev_allgather_layout(EcpDev *d, const int *const *rows, const long long *nrows) {
  CK(cudaSetDevice(d->device));
  const int world = d->commWorld, n = d->nAO;
  if (!d->comm || world < 1) {
    snprintf(g_err, sizeof(g_err), "ecpdev_allgather_layout: no communicator");
    return -1;
  }
  long long tot = 0;
  for (int r = 0; r < world; r++) tot += nrows[r];
  int *hr = (int *)malloc((size_t)(tot + 1) * sizeof(int));
  long long *ho = (long long *)malloc((size_t)(tot + 1) * sizeof(long long));
  free(d->agCount);
  d->agCount = (long long *)calloc(2 * (size_t)world + 2, sizeof(long long));
  d->agFirst = d->agCount + world + 1;
  long long k = 0, cap = 1;
  for (int r = 0; r < world; r++) {
    d->agFirst[r] = k;
    long long off = 0;
    for (long long i = 0; i < nrows[r]; i++, k++) {
      const int row = rows[r][i];
      if (row < 0 || row >= n || (i && row <= rows[r][i - 1])) {
        free(hr);
        free(ho);
        snprintf(g_err, sizeof(g_err), "ecpdev_allgather_layout: rows of rank %d are not ascending AO rows", r);
        return -1;
      }
      hr[k] = row;
      ho[k] = off; /* offset inside the rank's packed shard */
      off += n - row;
    }
    d->agCount[r] = off;
    if (off > cap) cap = off;
  }
  d->agFirst[world] = k;
  d->agCap = cap;
  g_allocStream = d->s1;
  int rc = ensure(&d->agRows, (size_t)(tot + 1) * sizeof(int));
  if (!rc) rc = ensure(&d->agOff, (size_t)(tot + 1) * sizeof(long long));
  if (!rc) rc = ensure(&d->agBuf, (size_t)cap * world * sizeof(double));
  if (!rc && tot) {
    rc = (int)cudaMemcpyAsync(d->agRows.p, hr, (size_t)tot * sizeof(int), cudaMemcpyHostToDevice, d->s1);
    if (!rc) rc = (int)cudaMemcpyAsync(d->agOff.p, ho, (size_t)tot * sizeof(long long), cudaMemcpyHostToDevice, d->s1);
    if (!rc) rc = (int)cudaStreamSynchronize(d->s1); /* hr / ho are pageable */
  }
  free(hr);
  free(ho);
  if (rc) snprintf(g_err, sizeof(g_err), "ecpdev_allgather_layout: device allocation / copy failed (%d)", rc);
  return rc ? -1 : 0;
}
/* pack own rows -> in-place ncclAllGather -> scatter the other ranks' rows; stream-ordered behind the last batch.
 * bytesRecv: bytes this rank received (shards of the others, unpadded). */
extern "C" int ecpdev_allgather(EcpDev *d, long long *bytesRecv) {
  CK(cudaSetDevice(d->device));
  if (!d->comm || !d->agCount || !d->matrix) {
    snprintf(g_err, sizeof(g_err), "ecpdev_allgather: %s", !d->comm ? "no communicator" : (!d->agCount ? "no shard layout" : "no result matrix yet"));
    return -1;
  }
  const int world = d->commWorld, me = d->commRank, n = d->nAO;
  double *buf = (double *)d->agBuf.p;
  const int *rows = (const int *)d->agRows.p;
  const long long *off = (const long long *)d->agOff.p;
  const long long cap = d->agCap;
  const long long nMine = d->agFirst[me + 1] - d->agFirst[me];
  if (nMine)
    k_pack_rows<<<(unsigned)nMine, 256, 0, d->s1>>>(d->matrix, n, rows + d->agFirst[me], off + d->agFirst[me], 0, buf + (size_t)me * cap);
  NCCLCK(g_nccl.allGather(buf + (size_t)me * cap, buf, (size_t)cap, 8 /* ncclFloat64 */, d->comm, d->s1));
  long long recv = 0;
  for (int r = 0; r < world; r++) {
    if (r == me) continue;
    const long long nr = d->agFirst[r + 1] - d->agFirst[r];
    if (nr)
      k_unpack_rows<<<(unsigned)nr, 256, 0, d->s1>>>(d->matrix, n, rows + d->agFirst[r], off + d->agFirst[r], buf + (size_t)r * cap);
    recv += d->agCount[r] * (long long)sizeof(double);
  }
  CK(cudaGetLastError());
  if (bytesRecv) *bytesRecv = recv;
  d->matrixKnown = 0; /* rows of other ranks are now non-zero: the next pass clears everything */
  return 0;
}

extern "C" int ecpdev_debug_fetch(EcpDev *d, const char *what, double *dst, int64_t n) {
  CK(cudaSetDevice(d->device));
  const void *src = NULL;
  size_t have = 0;
  if (!strcmp(what, "F")) { src = d->b.F; have = d->lastSizes[0]; }
  else if (!strcmp(what, "omegaX")) { src = d->b.omX; have = d->lastSizes[1]; }
  else if (!strcmp(what, "T")) { src = d->b.T; have = d->lastSizes[2]; }
  else if (!strcmp(what, "gamma")) { src = d->b.gamma; have = d->lastSizes[3]; }
  else if (!strcmp(what, "chi")) { src = d->b.chi; have = d->lastSizes[3]; }
  else if (!strcmp(what, "Q")) { src = d->b.Q; have = d->lastSizes[4]; }
  if (!src) return -1;
  if ((size_t)n > have) n = (int64_t)have;
  CK(cudaMemcpy(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

/* ---- device-side unit entry points (tests only): the per-point math of the kernels run on the GPU on caller-supplied
 * arguments, so that the GPU tier can compare primitives - not only end results - with the oracle:
 *   "bessel": in = n x z, ipar[0] = lmax                          -> out = n x (lmax + 1)   weighted Bessel functions
 *   "rsh"   : in = n x (theta, phi), ipar[0] = lmax               -> out = n x (lmax + 1)^2 real spherical harmonics
 *   "ps93"  : in = n x 3 rows of 384 slot-ordered doubles (Fa, Fb, U), ipar = n x (start, end)
 *                                                                  -> out = n x (result, rc, points)
 *   "pot"   : in = n x r, ipar[0] = ECP type, ipar[1] = l         -> out = n x U_l(r)
 * One thread per item; the same device functions the kernels call (ecp_math.h). */
__global__ void k_unit(DevT t, int what, int n, const double *in, const int *ipar, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (what == 0) {
    const int lmax = ipar[0];
    double K[ECP_KMAX + 1];
#pragma unroll
    for (int l = 0; l <= ECP_KMAX; l++) K[l] = 0.0;
    ecp_bessel<ECP_KMAX>(t.besselT, t.besselStride, t.besselC, lmax, in[i], K);
#pragma unroll
    for (int l = 0; l <= ECP_KMAX; l++)
      if (l <= lmax) out[(size_t)i * (lmax + 1) + l] = K[l];
  } else if (what == 1) {
    const int lmax = ipar[0];
    ecp_rsh(lmax, in[2 * i], in[2 * i + 1], t.fac, t.dfac, out + (size_t)i * (lmax + 1) * (lmax + 1));
  } else if (what == 2) {
    const double *Fa = in + (size_t)i * 3 * ECP_SMALL_SLOTS, *Fb = Fa + ECP_SMALL_SLOTS, *U = Fb + ECP_SMALL_SLOTS;
    double res = 0.0;
    int np = 0;
    const int rc = ecp_ps93_fastT(Fa, 1, Fb, 1, U, 1, c_small_w, &t.sm, t.small_jL, t.small_jR, ipar[2 * i], ipar[2 * i + 1],
                                  t.tolerance, &res, &np);
    out[3 * i] = res;
    out[3 * i + 1] = rc;
    out[3 * i + 2] = np;
  } else if (what == 3) {
    const int type = ipar[0], l = ipar[1];
    out[i] = ecp_pot_eval(t.gaussL, t.gaussN, t.gaussD, t.gaussA, t.typeGaussOff[type], t.typeGaussOff[type + 1], l, in[i]);
  }
}
extern "C" int ecpdev_unit(EcpDev *d, const char *what, int n, const double *in, int64_t nin, const int *ipar, int npar,
                           double *out, int64_t nout) {
  CK(cudaSetDevice(d->device));
  const int w = !strcmp(what, "bessel") ? 0 : (!strcmp(what, "rsh") ? 1 : (!strcmp(what, "ps93") ? 2 : (!strcmp(what, "pot") ? 3 : -1)));
  if (w < 0 || n <= 0) return -1;
  double *din = NULL, *dout = NULL;
  int *dpar = NULL;
  CK(cudaMalloc(&din, (size_t)nin * sizeof(double)));
  CK(cudaMalloc(&dout, (size_t)nout * sizeof(double)));
  CK(cudaMalloc(&dpar, (size_t)(npar > 0 ? npar : 1) * sizeof(int)));
  CK(cudaMemcpy(din, in, (size_t)nin * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dpar, ipar, (size_t)npar * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, (size_t)nout * sizeof(double)));
  k_unit<<<(n + 63) / 64, 64, 0, d->s1>>>(d->t, w, n, din, dpar, dout);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(d->s1));
  CK(cudaMemcpy(out, dout, (size_t)nout * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(din);
  cudaFree(dout);
  cudaFree(dpar);
  return 0;
}

extern "C" double ecpdev_fp64_peak_probe(int device, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
  const int blocks = nsm * 8, threads = 256;
  double *out = NULL;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_fp64_probe<<<blocks, threads>>>(out, 1000);
  cudaDeviceSynchronize();
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_fp64_probe<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}
