/* libecp_b200.h - extensions of the drop-in C ABI that only make sense on the device build:
 * device selection, multi-GPU sharding by shell-pair ownership, a device-resident result, run
 * statistics.  Everything is plain C (pointers, ints, doubles); no torch / CUDA types.
 *
 * Each entry cites the reference interface it extends or replaces.
 */
#ifndef LIBECP_B200_H
#define LIBECP_B200_H 1

#include "libecp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* CUDA device used by handles created afterwards (default 0, or env LIBECP_B200_DEVICE).
 * The reference has no device notion (no process/device boundary). */
void libecp_b200_set_device(int device);
/* Host threads used by the batch builder and the host-side accumulation (OpenMP); launchers like torchrun
 * export OMP_NUM_THREADS=1, so a multi-GPU caller hands every rank its share of the cores explicitly. */
void libecp_b200_set_host_threads(int n);
/* libECP_free parks the device scratch / result buffers of a handle for the next handle created on that device (a
 * caller of getIntegrals creates one handle per call); this returns them to the driver. */
void libecp_b200_release_cache(void);

/* Restrict a handle to the shell pairs owned by `rank` of `world` (disjoint output blocks per rank,
 * no data-path collective).  Replaces nothing in the reference (single-threaded loop
 * nest src/libecp.c:256-397); the union over ranks of what calculateECPIntegrals / the matrix
 * entry points produce equals the unsharded result block for block. */
void libecp_b200_set_shard(libECPHandle *h, int rank, int world);
/* owner rank of the shell pair (global shell indices) under that partition: the rank that owns the row shell
 * min(shellA, shellB); rows are dealt by a cost model over (l, contraction depth) (csrc/builder.c) */
int libecp_b200_pair_owner(libECPHandle *h, int shellA, int shellB, int world);

/* Upper-triangular ECP matrix accumulated on the device (what getIntegrals' callback
 * src/getIntegrals.c:22-43 builds on the host).  On return *devMatrix is a device pointer to
 * nAO*nAO doubles owned by the handle (valid until the next call or libECP_free), *nAO its dimension.
 * Same return codes as calculateECPIntegrals. */
int libecp_b200_integrals_device(libECPHandle *h, void **devMatrix, int *nAO);
/* device pointer of the handle's resident result matrix (NULL before the first pass); after a gather of a sharded run
 * (libecp_b200_unpack_rows) it holds the full upper-triangular matrix */
void *libecp_b200_matrix_ptr(libECPHandle *h);
/* same, then copied into host memory I (row stride rowdim) with += on the upper triangle */
int libecp_b200_integrals_host(libECPHandle *h, int rowdim, double *I);

/* Device-resident gather of a sharded run ("NCCL all-gather when the matrix stays on the device").
 * A rank's matrix is non-zero only in the upper-triangle parts M[i][i..nAO) of the AO rows of the shells it owns.
 *   libecp_b200_owned_rows : those rows (ascending) for any (rank, world) - every rank can list every other rank's rows;
 *   libecp_b200_pack_rows  : packs the listed rows of the handle's device matrix back to back into a caller-provided
 *                            DEVICE buffer of `cap` doubles (*elems = doubles used; devPacked NULL: only sizes it);
 *   libecp_b200_unpack_rows: scatters such a packed buffer (another rank's shard after the collective) into the
 *                            handle's device matrix.
 * These three are the building blocks for a caller that brings its own collective (libecp_b200/gather.py does it with
 * torch.distributed); libecp_b200_allgather below is the complete collective.  All three return 0 on success, -1 on
 * error (libecp_b200_last_error). */
long long libecp_b200_owned_rows(libECPHandle *h, int rank, int world, int *rows, long long cap);
int libecp_b200_pack_rows(libECPHandle *h, const int *rows, long long nrows, void *devPacked, long long cap, long long *elems);
int libecp_b200_unpack_rows(libECPHandle *h, const int *rows, long long nrows, const void *devPacked, long long cap);

/* The collective itself, behind the C ABI (a C / Fortran caller needs no Python and no NCCL code of its own).  NCCL is
 * bound with dlopen at first use.  Either the library creates the communicator -
 *     rank 0:    libecp_b200_comm_unique_id(id);  ... the caller distributes the 128 bytes (MPI_Bcast, a file, ...) ...
 *     all ranks: libecp_b200_comm_init(h, rank, world, id);          (collective; also does libecp_b200_set_shard)
 * - or it adopts the caller's:  libecp_b200_comm_attach(h, (void *)ncclComm, rank, world).
 * After libecp_b200_integrals_device() on every rank, libecp_b200_allgather(h, &bytes) packs the rank's rows, runs one
 * in-place ncclAllGather over NVLink and scatters the other shards: the handle's device matrix then holds the full
 * upper-triangular result on every GPU.  Stream-ordered on the handle's compute stream (libecp_b200_device_sync waits
 * for it).  All return 0 on success, -1 on error (libecp_b200_last_error). */
int libecp_b200_comm_unique_id(void *id128);
int libecp_b200_comm_init(libECPHandle *h, int rank, int world, const void *id128);
int libecp_b200_comm_attach(libECPHandle *h, void *ncclComm, int rank, int world);
int libecp_b200_allgather(libECPHandle *h, long long *bytesReceived);
int libecp_b200_device_sync(libECPHandle *h);
void libecp_b200_comm_free(libECPHandle *h);

typedef struct {
  long long nominal_triples;   /* centres x nshells(nshells+1)/2 : the reference's loop domain (src/libecp.c:256-312) */
  long long executed_triples;  /* survive screening (src/libecp.c:304-320,344) */
  long long shell_slots, atom_slots, prim_pairs;
  long long fast_quadratures, fast_failed, fallback_items, type1_fallback_pairs, stale_centre_events;
  long long kernel_launches, batches;
  long long h2d_bytes, d2h_bytes;  /* host<->device traffic of the last run (tables_h2d_bytes: once per handle) */
  long long tables_h2d_bytes;
  double ms_build, ms_tables, ms_fastT, ms_fallback, ms_link, ms_type1, ms_chi, ms_shift, ms_device_total;
} libecp_b200_stats_t;
void libecp_b200_get_stats(libECPHandle *h, libecp_b200_stats_t *out);
/* profiling aid: run the type-1 kernels on the same stream as the type-2 kernels (no overlap), so that the per-kernel
 * times in the statistics are each kernel's own duration */
void libecp_b200_set_serial_kernels(libECPHandle *h, int on);

/* screening decisions of one centre, for parity tests against the reference's
 * ScreenedGrid/potentialScreening (src/type2.c:148-180,201-203): arrays of nrShells ints */
int libecp_b200_screening(libECPHandle *h, int centre, int *end_l /* [L] */, int *start, int *end, int *skip);

/* host table access for bit-exactness tests (reference tables: src/libecp.c:147-198):
 * names "fac" "dfac" "poly2sph" "omega" "small_x" "small_w" "large_x" "large_w" "bessel" ; returns length */
int libecp_b200_host_table(libECPHandle *h, const char *name, const double **ptr);
/* integer tables for tests: "small_oidx" "large_oidx" "small_meta" "dims" "ijk" "ijkIndex" "atomType" "lastCentre" (per atom:
 * the last ECP centre whose screening can keep one of its shells, -1 = none - the rows of the atom are final once the pass
 * is beyond it, which is when the host consumer downloads them); returns count */
int libecp_b200_host_itable(libECPHandle *h, const char *name, int *out, int cap);
/* executed triples of the whole job in the reference's loop order, rows (A,s1,la,B,s2,lb,C); host only */
long long libecp_b200_triple_list(libECPHandle *h, int *out, long long cap);
/* Spherical-harmonic output ("output to spherical AOs"; the reference carries TM_cart2sph / TM_sph2cart,
 * src/transformations.c:28-141, but returns Cartesian blocks only).  The pure function (l, m) of a shell is
 * sum_c cart2sph[l][m][c] * (Cartesian component c), with the reference's table and m = 0 .. 2l in its order;
 * S = C^T M C is formed on the device from the resident Cartesian matrix of an n = 0 run.
 *   _spherical_dim    : sum over shells of 2l + 1
 *   _spherical_device : run the integrals, leave S (upper triangle, nSph x nSph, row-major) in device memory owned by the handle
 *   _spherical_host   : the same, then S += into the caller's matrix (upper triangle, leading dimension rowdim)
 * return values as libecp_b200_integrals_device. */
int libecp_b200_spherical_dim(libECPHandle *h);
int libecp_b200_spherical_device(libECPHandle *h, void **devS, int *nSph);
int libecp_b200_spherical_host(libECPHandle *h, int rowdim, double *S);
/* callback keys of the whole job in call order, rows (A,s1,la,shifta,B,s2,lb,shiftb,C), one row per executed (shifted)
 * triple = one type-1 and one type-2 callback (reference src/libecp.c:372); host only */
long long libecp_b200_callback_keys(libECPHandle *h, int *out, long long cap);
/* Wall time (ms) of the host batch builder alone over one pass (no device work); optional totals.  By default this is
 * what a matrix run leaves to the host - screening and slot layout; *triples then counts the shell pairs (owned a, b >= a)
 * handed to the device enumeration.  With LIBECP_B200_ENUM=host: the full host enumeration and the executed triples. */
double libecp_b200_build_only(libECPHandle *h, long long *triples, int *batches);
/* test hook: handles created afterwards build tables + batches only (no device); compute entry points
 * then fail with -1.  Used by the CPU-only test tier; never a compute fallback. */
void libecp_b200_set_tables_only(int on);
/* last batch's device intermediates, tests only: "F" "omegaX" "T" "gamma" "chi" "Q" */
int libecp_b200_debug_fetch(libECPHandle *h, const char *what, double *dst, long long n);
/* tests only: the per-point device functions of the kernels run on the GPU on caller-supplied arguments -
 * "bessel" (weightedBesselFunction, src/bessel.c:101-199), "rsh" (realSphericalHarmonics, src/spherical_harmonics.c:15-114),
 * "ps93" (integrateGC_PS93 on slot-ordered tables, src/gc_integrators.c:156-217), "pot" (evalECP, src/ecp.c:46-60);
 * argument layout in csrc/ecp_cuda.cu (k_unit) */
int libecp_b200_debug_unit(libECPHandle *h, const char *what, int n, const double *in, long long nin, const int *ipar, int npar,
                           double *out, long long nout);

/* measured FP64 FMA throughput of the device in TFLOP/s (roofline denominator for bench.py) */
double libecp_b200_fp64_peak(int device, int iters);
const char *libecp_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
