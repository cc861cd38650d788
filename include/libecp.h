/* libecp.h - drop-in public C API of the B200-native libECP hot path.
 *
 * Same three entry points, argument order and return codes as the reference header
 * (reference src/libecp.h:6-29; definitions src/libecp.c:53-61,212-213,407-408).  Existing callers
 * re-link against libecp_b200.so (or libecp.a + the CUDA runtime) without source changes.
 *
 * Differences a caller can observe:
 *   - the work runs on a CUDA device (sm_100a); libECP_init returns NULL when none is usable - there is
 *     no CPU path;
 *   - derivative order n = 0, 1 or 2 (n >= 1: the callback receives the shifted-momentum blocks of the reference,
 *     src/libecp.c:246-250,322-373; n = 2, shifts (+1,0) / (0,+1): IJK_DIM(la) x IJK_DIM(lb) elements, :362-369);
 *     shellOrdering / lmax as in the reference (src/libecp.c:152-166): NULL = libint order, else the caller's
 *     Cartesian component order for l = 0..lmax, lmax >= maxLambda + maxAlpha + 1;
 *   - shapes must satisfy (with maxLBS + n for maxLBS in a derivative run) maxLBS <= 5, L_ECP <= 6, L_ECP-1+maxLBS <= 10, maxLBS <= L_ECP+1.
 */
#ifndef LIBECP_H
#define LIBECP_H 1

#ifdef __cplusplus
extern "C" {
#endif

typedef struct _libECPHandle libECPHandle;

/* Positional meaning is the reference call site's (src/libecp.c:372):
 *   cb(A, s1, la, shifta, B, s2, lb, shiftb, C, I, args)
 * s1/s2 are shell indices within atom A/B; I is row-major IJK_DIM(la + shifta) x IJK_DIM(lb + shiftb) (the shifts are 0
 * in an n = 0 run), valid only during the call.  Invoked on the caller's thread, in the reference's loop order (C, A, B>=A, s1, s2), type 1
 * then type 2 for every executed triple.  (The parameter names below are the reference typedef's.) */
typedef void (*ECPCallback)(int A, int B, int C,
			    int sa, int sb,
			    int la, int lb,
			    int shifta, int shiftb,
			    double *I, void *p);

/* create a handle: builds all geometry-independent tables on the host (bit-exact with the reference)
 * and uploads them to the device.  Borrows geometry, shellsBS, lBS, KBS, dBS, aBS (must outlive the
 * handle, as in the reference src/libecp.c:68-74); copies the ECP arrays.
 * NOTE positional order: argument 4 is l per ECP shell, argument 5 the number of Gaussians per ECP shell
 * (reference definition src/libecp.c:53-61; the reference header names them the other way round). */
libECPHandle * libECP_init(int nrAtoms, double *geometry,
			   int *shellsECP, int *lECP,
			   int *KECP, double *nECP, double *dECP, double *aECP,
			   int *shellsBS, int *lBS, int *KBS,
			   double *dBS, double *aBS,
			   int n, int lmax, int *shellOrdering,
			   int largeGridOrder, double tolerance,  double accuracy);

/* return value:
   0 - successful integration
   1 - error during type1 integration
   2 - error during type2 integration
   (negative: CUDA failure, message on stderr) */
int calculateECPIntegrals(libECPHandle *h, ECPCallback cb, void *args);

void libECP_free(libECPHandle *h);

#ifdef __cplusplus
}
#endif
#endif
