/* ex1.c - the reference's example program (example/ex1.c) against the B200 library: load a structure, an ECP and a
 * basis set from text files, compute the ECP integral matrix with the one-call interface, print it shell block by
 * shell block.  Same command line:   ex1 structure.xyz ECP BS   (+ optional 4th argument "shipped" to read the ECP
 * file in the format of the shipped example/test_c.ecp instead of the format the reference's loader expects).
 *
 *   gcc -Iinclude examples/ex1.c -Llibecp_b200/lib -lecp_b200 -Wl,-rpath,$PWD/libecp_b200/lib -o ex1
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dimensions.h>
#include <getIntegrals.h>
#include <libecp.h>
#include <libecp_b200_io.h>

int main(int argc, char **argv) {
  if (argc != 4 && argc != 5) {
    fprintf(stderr, "usage: %s structure.xyz ECP BS [shipped]\n", argv[0]);
    return 2;
  }
  const int fmt = (argc == 5 && !strcmp(argv[4], "shipped")) ? LIBECP_IO_ECP_SHIPPED : LIBECP_IO_ECP_INDEXED;
  int nat = 0, nsh = 0;
  double *geom = NULL, *aE = NULL, *dE = NULL, *nE = NULL, *aB = NULL, *dB = NULL;
  int *shE = NULL, *lE = NULL, *KE = NULL, *shB = NULL, *lB = NULL, *KB = NULL;
  if (libecp_io_load_xyz(argv[1], &nat, &geom) || libecp_io_load_ecp(argv[2], nat, fmt, &shE, &lE, &KE, &aE, &dE, &nE) ||
      libecp_io_load_bs(argv[3], nat, &shB, &lB, &KB, &aB, &dB, &nsh)) {
    fprintf(stderr, "%s: cannot read the input files\n", argv[0]);
    return 1;
  }
  const int dim = libecp_io_ao_dim(nsh, lB);
  double *I = calloc((size_t)dim * dim, sizeof(double));
  /* note the argument order of getIntegrals: K before l for the ECP (include/getIntegrals.h) */
  if (getIntegrals(nat, geom, shE, KE, lE, nE, dE, aE, shB, lB, KB, dB, aB, 1024, 1.0E-12, 1.0E-14, dim, I)) return 1;

  printf("ECP matrix elements (upper triangle, %d x %d)\n\n", dim, dim);
  int *first = calloc((size_t)nsh + 1, sizeof(int)), *atomOf = calloc((size_t)nsh + 1, sizeof(int));
  for (int a = 0, s = 0; a < nat; a++)
    for (int k = 0; k < shB[a]; k++, s++) {
      atomOf[s] = a;
      first[s + 1] = first[s] + IJK_DIM(lB[s]);
    }
  for (int s1 = 0; s1 < nsh; s1++)
    for (int s2 = s1; s2 < nsh; s2++) {
      printf("shell block (atom %d, shell %d: l=%d) x (atom %d, shell %d: l=%d)\n", atomOf[s1], s1, lB[s1], atomOf[s2], s2,
             lB[s2]);
      for (int i = first[s1]; i < first[s1 + 1]; i++) {
        for (int j = first[s2]; j < first[s2 + 1]; j++) printf(" % 10.6f  ", I[(size_t)i * dim + j]);
        printf("\n");
      }
    }
  free(first); free(atomOf); free(I);
  libecp_io_free(geom); libecp_io_free(shE); libecp_io_free(lE); libecp_io_free(KE); libecp_io_free(aE);
  libecp_io_free(dE); libecp_io_free(nE); libecp_io_free(shB); libecp_io_free(lB); libecp_io_free(KB);
  libecp_io_free(aB); libecp_io_free(dB);
  return 0;
}
