/* libecp_b200_io.h - input loaders for the plain-text formats of the reference's example program
 * (reference example/ex1.c:11-123: loadBS, loadGeometry, loadECP).  Host-only C, no device work.
 *
 * Formats (whitespace separated, one block per atom, atoms in the order of the structure file):
 *   structure (.xyz)  : <nrAtoms> / one comment line / <symbol> <x> <y> <z> per atom.  Coordinates are handed to the
 *                       library unchanged, as the reference example does (it expects bohr).
 *   basis set (.bs)   : <Z> <nshells> ; per shell <l> <K> ; per primitive <index> <exponent> <coefficient>
 *   ECP (.ecp)        : <Z> <ncore> <nshells> ; per shell <l> <K> ; per Gaussian
 *                         LIBECP_IO_ECP_INDEXED : <index> <exponent a> <coefficient d> <power n>   - what the reference's
 *                                                 loadECP reads (example/ex1.c:114);
 *                         LIBECP_IO_ECP_SHIPPED : <coefficient d> <exponent a> <power n>            - what the shipped file
 *                                                 example/test_c.ecp actually contains.
 *                       Reading the shipped file with the INDEXED rule is the reference's behaviour (SURVEY.md App. C-1:
 *                       "37.4565 6.8446 2" becomes index 37, a = 0.4565, d = 6.8446) and is what configuration 1 of the
 *                       parity tests means by "as shipped"; SHIPPED is the corrected reading.
 * All arrays are allocated with malloc and owned by the caller (libecp_io_free).  Returns 0, -1 if the file cannot be
 * opened, -2 if it ends early or a field does not parse (the reference example has no error handling at all).
 */
#ifndef LIBECP_B200_IO_H
#define LIBECP_B200_IO_H 1

#ifdef __cplusplus
extern "C" {
#endif

#define LIBECP_IO_ECP_INDEXED 0
#define LIBECP_IO_ECP_SHIPPED 1

int libecp_io_load_xyz(const char *path, int *nrAtoms, double **geometry);
int libecp_io_load_bs(const char *path, int nrAtoms, int **shellsBS, int **lBS, int **KBS, double **aBS, double **dBS,
                      int *nrShells);
int libecp_io_load_ecp(const char *path, int nrAtoms, int format, int **shellsECP, int **lECP, int **KECP, double **aECP,
                       double **dECP, double **nECP);
/* dimension of the Cartesian AO matrix of a basis (sum of (l+1)(l+2)/2 over the shells) */
int libecp_io_ao_dim(int nrShells, const int *lBS);
void libecp_io_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
