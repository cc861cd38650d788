/* ecp_dev.h - internal C ABI between the host C layer (tables.c, builder.c, api.c) and the CUDA layer
 * (ecp_cuda.cu).  Plain pointers and sizes only.  Not installed; the public headers are include/*.h.
 *
 * Vocabulary (follows the reference's domain):
 *   centre C      ECP-bearing atom                       (src/libecp.c:256)
 *   atom slot     (C, atom X) with >=1 unskipped shell   -> r_XC, Omega_X, usp_X   (src/libecp.c:278-292)
 *   shell slot    (C, shell) that survives screening     -> window [start,end], F table (src/type2.c:246-303)
 *   triple        (shell slot a, shell slot b) of one C  -> chi, gamma, two integral blocks (src/libecp.c:297-376)
 *   class         (la, lb, L_C): triples of one class share quadrature lists and sizes
 */
#ifndef ECP_DEV_H
#define ECP_DEV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECP_SMALL_ORDER 383  /* PS93 grid, order 128 -> 3*2^7-1 points (src/gc_integrators.c:161)   */
#define ECP_SMALL_SLOTS 384  /* level-major padded layout, see tables.c:build_small_grid             */
#define ECP_SMALL_LEVELS 13
#define ECP_MAX_LBS 5        /* h shells                                                             */
#define ECP_MAX_LECP 6
#define ECP_KMAX 10          /* highest Bessel order ever requested: max(la+l, la+lb) <= 10          */
#define ECP_MAX_CLASSES 256

/* geometry-independent tables, built bit-exactly on the host (tables.c) */
typedef struct {
  int maxLECP, maxLBS, maxAlpha, maxLambda, tmDim, ijkDim;
  int besselLMax;    /* rows in the Bessel table - 1                              */
  int besselStride;  /* doubles per abscissa in the transposed table              */
  int largeOrder;    /* 1023                                                      */
  int largeSlots;    /* 1024                                                      */
  int largeLevels;   /* 9                                                         */
  double tolerance, accuracy, lnAccuracy1, lnAccuracy2;
  const double *fac, *dfac;       int nfac;
  const int *ijk, *ijkIndex;      /* [3*C_DIM(tmDim)], [ijkDim^3]                  */
  const double *poly2sph;         /* [C_DIM(tmDim)][L_DIM(tmDim)]                  */
  const double *cart2sph;         /* packed [l][m][c] (reference src/transformations.h:13-14), l <= tmDim */
  int ncart2sph;
  const double *omega;            /* [L_DIM(maxLECP)][L_DIM(maxLambda)][C_DIM(maxAlpha)] */
  int nomega;
  const double *binom;            /* [(maxLBS+1)^2] n over k                       */
  /* binomial-shift term lists per (l, component c) in the reference's loop order (src/util.c:275-283):
   * terms [shTermOff[l*shOffStride + c], shTermOff[l*shOffStride + c + 1]) */
  int shOffStride, nShTerms;
  const int *shTermOff;           /* [(maxLBS+1)*shOffStride]                      */
  const int *shTermP;             /* C_INDEX of the sub-monomial (alpha_x,alpha_y,alpha_z) */
  const int *shTermD;             /* (a-alpha) packed dx | dy<<4 | dz<<8           */
  const double *shTermBin;        /* binomial product                              */
  /* small grid, level-major padded layout */
  const double *small_r, *small_w;   /* [384]                                      */
  const int16_t *small_oidx;         /* [384] original index, -1 for the pad slot  */
  int small_levPairs[ECP_SMALL_LEVELS], small_levJ[ECP_SMALL_LEVELS], small_levN[ECP_SMALL_LEVELS];
  int small_levSlot[ECP_SMALL_LEVELS + 1];
  /* large grid template on (-1,1), level-major padded layout */
  const double *large_x, *large_w;   /* [1024]                                     */
  const double *large_xo;            /* abscissae in original (ascending) order [largeOrder] */
  const int16_t *large_oidx;         /* [1024]                                     */
  /* Bessel table transposed to [1601][besselStride], and C_j */
  const double *besselT, *besselC;
  /* basis set */
  int nrShells, nrPrims, nrAtoms, nAO;
  const int *shellL, *shellK, *shellPrim, *shellAtom, *shellAO;
  const double *primD, *primA;
  /* ECP types (distinct parameter sets) */
  int nTypes;
  const int *typeL;                  /* [nTypes]                                   */
  const int *typeGaussOff;           /* [nTypes+1] into gauss arrays               */
  const int *gaussL;  const double *gaussN, *gaussD, *gaussA;
  int nU;                            /* rows N of r^N U_l: max(maxLambda, 2 maxAlpha)+1 */
  const double *typeUtab;            /* per type [maxLECP][nU][384] slots          */
  const double *typeUL;              /* per type [384]                             */
  /* classes */
  int nClasses;
  const int *clsLa, *clsLb, *clsL;   /* [nClasses]                                 */
  const int *clsNq;                  /* used type-2 quadratures per triple         */
  const int *clsQOff;                /* [nClasses+1] offsets into qlist            */
  const int *clsQlOff;               /* [nClasses*(ECP_MAX_LECP+1)] start of each l inside the class list */
  const int *qlist;                  /* packed l | l1<<4 | l2<<8 | l3<<12           */
  const int *clsQidxOff;             /* [nClasses+1] offsets into qidx             */
  const int16_t *qidx;               /* [L][la+L][lb+L][la+lb+1] -> position in class list or -1 */
  int nqlist, nqidx;
  int maxQPerL;                      /* max used quadratures of any (class,l)      */
} EcpHostTables;

/* one batch of centres, produced by builder.c */
typedef struct {
  /* atom slots */
  int nASlots;
  const int *asAtom, *asCentre, *asType;
  const double *asR;        /* [nASlots][4]: r_XC (x,y,z), d_XC                    */
  const int64_t *asOmOff;   /* offset (doubles) of Omega_X                         */
  int64_t omTotal;
  /* shell slots */
  int nSSlots;
  const int *ssShell, *ssASlot, *ssStart, *ssEnd;
  const int64_t *ssFOff;    /* offset in rows of 384 doubles                       */
  int64_t fRows;
  /* triples, sorted by class */
  int nTriples;
  const int *trA, *trB;     /* shell slots                                         */
  const int64_t *trOut;     /* offset of the (type1,type2) block pair in the block buffer */
  const int64_t *trPair;    /* first type-1 primitive pair                         */
  int64_t tTotal, gTotal, outTotal, nPairs, qTotal, rshTotal;
  /* per class ranges (triples of class c are [clsFirst[c], clsFirst[c+1])) */
  const int *clsFirst;      /* [nClasses+1]                                        */
  const int64_t *clsWork;   /* [nClasses+1] prefix of ntriples*nq (fast-T threads) */
  const int64_t *clsElem;   /* [nClasses+1] prefix of ntriples*C_DIM(la)*C_DIM(lb) */
  const int64_t *clsOutElem;/* [nClasses+1] prefix of ntriples*IJK(la)*IJK(lb)     */
  const int64_t *clsPairBase; /* [nClasses+1] prefix of primitive pairs per class  */
  const int64_t *clsQBase;    /* [nClasses+1] prefix of pairs*(la+lb+1)^2 (offsets into Q / rsh) */
  /* Device enumeration (matrix runs): the host only screens - per centre the atom slots and shell slots above - and the
   * triples, their class order and every per-class prefix are produced on the device (ecp_enum.cuh).  ecpdev_run_batch
   * fills nTriples, nPairs, the totals and the cls* arrays (which then point at writable host storage) before the
   * kernels of the batch are sized. */
  int devEnum;
  int nCentres;
  const int *ceAS0;           /* [nCentres+1] first atom slot of every centre                     */
  const int64_t *cePair0;     /* [nCentres+1] first atom-slot pair (ka <= kb) of every centre     */
  const int *asSS0;           /* [nASlots+1] first shell slot of every atom slot                  */
  const unsigned char *ssOwn; /* [nSSlots] the row of this shell belongs to the rank              */
  int64_t pairCand;           /* shell pairs (owned a, b >= a) the device tests: for the next batch's size estimate */
} EcpBatch;

typedef struct {
  double ms_tables, ms_fastT, ms_fallback, ms_link, ms_type1, ms_chi, ms_shift, ms_total;
  long long nFallbackItems, nFastFail, nType1Fail, nStaleCentre, launches, h2dBytes, d2hBytes;
  int err1, err2;
} EcpDevStats;

typedef struct EcpDev EcpDev;

/* page-locked (cached, process-wide) host blocks for the builder's batch arrays; plain memory without a device */
void *ecpdev_pinned_alloc(size_t bytes);
void ecpdev_pinned_free(void *p);
/* all return 0 on success, else a cudaError_t value (message via ecpdev_last_error) */
EcpDev *ecpdev_create(const EcpHostTables *t, int device);
void ecpdev_destroy(EcpDev *d);
const char *ecpdev_last_error(void);
/* matrix accumulation target (device resident, nAO x nAO, zeroed) */
int ecpdev_matrix_begin(EcpDev *d, const unsigned char *rowOwned /* [nAO] or NULL = all */, long long ownershipSig);
int ecpdev_matrix_download(EcpDev *d, double *host /* nAO*nAO */);
/* pack (dir 0) / scatter (dir 1) the upper-triangle parts of the listed rows between the matrix and a device buffer */
int ecpdev_matrix_rows(EcpDev *d, int dir, const int *rows, long long nrows, void *devBuf, long long cap, long long *elems);
int ecpdev_matrix_add_to_host(EcpDev *d, double *host, int rowdim, const unsigned char *rowOwned /* [nAO] or NULL */,
                              long long *bytes, int async /* 1: download stream only, rows are final */);
void *ecpdev_matrix_ptr(EcpDev *d);
/* spherical-harmonic form of the resident matrix: S = C^T M C per shell pair with the handle's cart2sph table, upper
 * triangle, nSph = sum over shells of 2l+1; the buffer belongs to the handle */
int ecpdev_spherical(EcpDev *d, void **devS, int *nSph);
int ecpdev_spherical_add_to_host(EcpDev *d, double *host, int rowdim);
/* C-ABI collective of a sharded device-resident result (NCCL bound with dlopen): unique id for the caller to distribute,
 * communicator from that id (comm == NULL) or adopted from the caller, shard layout of all ranks, the all-gather itself */
int ecpdev_comm_unique_id(void *id128);
int ecpdev_comm_init(EcpDev *d, int rank, int world, const void *id128, void *comm);
void ecpdev_comm_destroy(EcpDev *d);
int ecpdev_allgather_layout(EcpDev *d, const int *const *rows, const long long *nrows);
int ecpdev_allgather(EcpDev *d, long long *bytesRecv);
void ecpdev_bind_thread(EcpDev *d);
void ecpdev_release_cache(void);
/* run one batch: flags bit0 = accumulate into matrix, bit1 = keep blocks and copy them to hostBlocks; slot = which of
 * the two input sets to use (alternate between consecutive batches).  ecpdev_prefetch_batch copies the inputs of the
 * NEXT batch into the other set on a copy stream while this one runs (any host thread). */
int ecpdev_run_batch(EcpDev *d, EcpBatch *b, int flags, int slot, double *hostBlocks, EcpDevStats *stats);
int ecpdev_prefetch_batch(EcpDev *d, const EcpBatch *b, int flags, int slot);
void ecpdev_invalidate_prefetch(EcpDev *d);
int ecpdev_sync(EcpDev *d);
long long ecpdev_table_bytes(EcpDev *d);
void ecpdev_set_serial(EcpDev *d, int on);
/* debug access to the intermediates of the last batch (tests only): "F" "omegaX" "T" "gamma" "chi" "Q" "tfail" */
int ecpdev_debug_fetch(EcpDev *d, const char *what, double *dst, int64_t n);
/* tests only: run a per-point device function on the GPU ("bessel" "rsh" "ps93" "pot"; ecp_cuda.cu: k_unit) */
int ecpdev_unit(EcpDev *d, const char *what, int n, const double *in, int64_t nin, const int *ipar, int npar, double *out,
                int64_t nout);
/* FP64 FMA peak probe used by bench.py for the roofline denominator: returns TFLOP/s */
double ecpdev_fp64_peak_probe(int device, int iters);

#ifdef __cplusplus
}
#endif
#endif
