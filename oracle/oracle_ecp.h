/* oracle_ecp.h - TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of libECP's hot path (type-1 + type-2 ECP integrals over every
 * shell-pair x ECP-centre triple).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (libecp_b200/) never does.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this restatement is
 * pinned against the UNMODIFIED reference compiled by oracle/Makefile into oracle/_ref/libecp_ref.so
 * (bit-for-bit on every callback block, tests/test_oracle.py) and against the committed fixtures
 * under tests/golden/ that were generated from that build (tests/golden/make_golden.py).
 *
 * Same public surface as the reference (src/libecp.h:15-29, src/getIntegrals.h:7-13) with an
 * "oracle_" prefix, derivative order n = 0 only.
 */
#ifndef ORACLE_ECP_H
#define ORACLE_ECP_H

typedef struct OracleECP OracleECP;

typedef void (*OracleCallback)(int A, int s1, int la, int shifta, int B, int s2, int lb, int shiftb, int C,
                               double *I, void *args);

OracleECP *oracle_libECP_init(int nrAtoms, double *geometry, int *shellsECP, int *lECP, int *KECP, double *nECP,
                              double *dECP, double *aECP, int *shellsBS, int *lBS, int *KBS, double *dBS, double *aBS,
                              int n, int lmax, int *shellOrdering, int largeGridOrder, double tolerance,
                              double accuracy);
int oracle_calculateECPIntegrals(OracleECP *h, OracleCallback cb, void *args);
void oracle_libECP_free(OracleECP *h);
int oracle_getIntegrals(int nrAtoms, double *geometry, int *shellsECP, int *KECP, int *lECP, double *nECP,
                        double *dECP, double *aECP, int *shellsBS, int *lBS, int *KBS, double *dBS, double *aBS,
                        int largeGridOrder, double tolerance, double accuracy, int rowdim, double *I);

/* --- unit-level accessors used by the parity tests ------------------------------------------- */
/* names: "fac" "dfac" "cart2sph" "poly2sph" "omega" "small_x" "small_w" "large_x" "large_w" "bessel" "besselC" */
const double *oracle_table(OracleECP *h, const char *name, int *len);
const int *oracle_itable(OracleECP *h, const char *name, int *len); /* "ijk" "ijkIndex" "dims" */
void oracle_bessel(OracleECP *h, int lmax, double z, double *K /* [lmax+1] */);
void oracle_rsh(OracleECP *h, int lmax, double theta, double phi, double *out /* [(lmax+1)^2] */);
void oracle_sphcoord(const double *xyz, double *rtp);
double oracle_shell_radius(int depth, int am, const double *d, const double *a, double zero);
/* quadrature drivers on a caller-supplied integrand table f[order] (window [start,end]); rc 0 ok / 1 failed */
int oracle_ps93_table(OracleECP *h, const double *f, int start, int end, double *result, int *npoints);
int oracle_psm92_table(OracleECP *h, const double *x_unused, const double *wf, int start, int end, double *result,
                       int *npoints);
/* per-centre screening: fills end_l[L], and per shell start/end/skip (arrays of nrShells) */
void oracle_screening(OracleECP *h, int C, int *end_l, int *sstart, int *send, int *sskip);
/* behaviour switch for the stale-buffer quirk of src/type2.c:443-448,471-495 (1 = reproduce, default) */
void oracle_set_stale_buffers(OracleECP *h, int on);
/* work counters of the last oracle_calculateECPIntegrals (see oracle_ecp.c: struct Counters) */
void oracle_counters(OracleECP *h, double *out, int n);

#endif
