/* getIntegrals.h - one-call interface, same signature as the reference (src/getIntegrals.h:7-13).
 * Note K comes before l here (the reference swaps them when calling libECP_init, src/getIntegrals.c:78).
 * Accumulates (+=) the upper triangle (incl. diagonal) of the nAO x nAO ECP matrix into I (row stride
 * rowdim); the caller zeroes I.  Returns 1 if the handle could not be created, else 0. */
#ifndef GET_INTEGRALS_H
#define GET_INTEGRALS_H 1

#ifdef __cplusplus
extern "C" {
#endif

int getIntegrals (int nrAtoms, double *geometry,
		  int *shellsECP, int *KECP,
		  int *lECP, double *nECP, double *dECP, double *aECP,
		  int *shellsBS, int *lBS, int *KBS,
		  double *dBS, double *aBS,
		  int largeGridOrder, double tolerance,  double accuracy,
		  int rowdim, double *I);

#ifdef __cplusplus
}
#endif
#endif /* GET_INTEGRALS_H */
