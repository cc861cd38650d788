/* tables.h - host-side, geometry-independent tables of a libECP handle (built once per handle).
 * All values are computed with the reference's operation order and glibc libm so that what is
 * uploaded to the device equals the reference's tables bit for bit (tests/test_host_tables.py). */
#ifndef ECP_TABLES_H
#define ECP_TABLES_H

#include "ecp_dev.h"

typedef struct {
  int L, N;       /* max angular momentum (local channel), number of Gaussians */
  int gaussOff;   /* offset into the flat gauss arrays                        */
  int endl[ECP_MAX_LECP + 1]; /* cumulative potential cut-off per l < L (src/type2.c:201-203) */
  int endLast;    /* window cap for basis screening = endl[L-1] (382 if L == 0) */
} EcpType;

typedef struct {
  EcpHostTables v; /* view handed to the CUDA layer (pointers into the arrays below) */
  /* owned storage */
  double *fac, *dfac, *cart2sph, *poly2sph, *omega, *binom;
  int *shTermOff, *shTermP, *shTermD;
  double *shTermBin;
  int *ijk, *ijkIndex;
  double *small_x, *small_w;   /* original order [383] (screening uses these)   */
#define ECP_WIN_LUT_SCALE 16.0 /* bins of 1/16 bohr */
#define ECP_WIN_LUT_BINS 640
  int winLut[ECP_WIN_LUT_BINS + 1]; /* winLut[k] = number of small-grid points with r < k/16 (builder.c: ecp_shell_window) */
  double *small_rs, *small_ws; /* slot layout [384]                              */
  int16_t *small_oidx;
  double *large_x, *large_w;   /* original order [1023]                          */
  double *large_xs, *large_ws; /* slot layout [1024]                             */
  int16_t *large_oidx;
  double *besselK, *besselT, *besselC;
  int *shellL, *shellK, *shellPrim, *shellAtom, *shellAO;
  double *shellRadius;
  double *atomRmax;   /* largest shell radius per atom (atom-level screening prune) */
  int *rowDeal;       /* position of every shell in the order rows are dealt to the ranks of a sharded run; NULL until
                         ecp_tables_row_deal() is called (unsharded handles never need it) */
  const int *dealL, *dealK; /* borrowed basis arrays the deal is computed from (the handle borrows them anyway) */
  const double *dealA;
  int *atomMaxL, *atomFirstShell;
  /* derivative runs (scope row f1): the shell list is the expanded one - every shell of the caller's basis is followed
   * by its shifted copies - l + k with coefficients d zeta^k (reference src/type1.c:239-246, src/type2.c:263-269,459-462),
   * l - k with d, second derivatives also l with d zeta (src/libecp.c:246-250,362-369); virtShift = copy number */
  int deriv;
  int *virtShift, *virtLocal;
  /* ECP */
  int *atomType; /* per atom: type index or -1 */
  EcpType *types;
  int *typeL, *typeGaussOff, *gaussL;
  double *gaussN, *gaussD, *gaussA, *typeUtab, *typeUL;
  /* classes */
  int *clsLa, *clsLb, *clsL, *clsNq, *clsQOff, *clsQlOff, *qlist, *clsQidxOff;
  int16_t *qidx;
  int clsLookup[ECP_MAX_LBS + 1][ECP_MAX_LBS + 1][ECP_MAX_LECP + 1];
} EcpTables;

/* optional inputs of ecp_tables_build (NULL = none) */
typedef struct {
  /* caller's Cartesian component order (reference src/libecp.c:152-166): exponent triples of every component of the
   * shells l = 0..lmaxOrd in the layout of cartesianShellOrder(lmaxOrd); NULL = libint order */
  const int *shellOrdering;
  int lmaxOrd;
  /* derivative runs (api.c): the basis handed in is the expanded list of shifted shells; screenParent[s] = the shell
   * whose radius screens shell s (the unshifted one, reference src/type2.c:251); NULL = every shell screens itself */
  const int *screenParent;
  /* derivative order of the run (0, 1 or 2) and, per shell of the expanded list, its copy number (0 = the caller's shell,
   * > 0 = a shifted copy, builder.c: ecp_deriv_copy) and the
   * position of its unshifted shell among the shells of the atom in the caller's basis (callback argument s1 / s2) */
  int deriv;
  const int *virtShift, *virtLocal;
} EcpBuildOpts;

/* returns NULL on unsupported shape / Bessel series failure (reference: src/libecp.c:159-162,181-185) */
EcpTables *ecp_tables_build(int nrAtoms, const double *geometry, const int *shellsECP, const int *lECP,
                            const int *KECP, const double *nECP, const double *dECP, const double *aECP,
                            const int *shellsBS, const int *lBS, const int *KBS, const double *dBS, const double *aBS,
                            int largeGridOrder, double tolerance, double accuracy, const EcpBuildOpts *opts);
void ecp_tables_free(EcpTables *t);
/* why the last ecp_tables_build of this thread returned NULL */
const char *ecp_tables_last_error(void);
/* row deal of a sharded run (see rowDeal); computed on first use, thread-safe */
const int *ecp_tables_row_deal(EcpTables *t);

/* helpers shared with builder.c */
double ecp_host_pot_eval(const EcpTables *t, int type, int l, double r);
int ecp_t2_used(int la, int lb, int l, int l1, int l2, int l3);

#endif
