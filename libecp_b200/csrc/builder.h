/* builder.h - host batch builder: replaces the loop nest of calculateECPIntegrals
 * (reference src/libecp.c:256-397) and the screening of calcF_FM06 (src/type2.c:246-260).
 * All integer screening decisions (skip flags, windows, executed-triple list) are taken here, on the
 * host, in double precision with the reference's comparisons - bit-exact by construction. */
#ifndef ECP_BUILDER_H
#define ECP_BUILDER_H

#include "tables.h"

typedef struct {
  EcpBatch b; /* view for the CUDA layer */
  /* owned storage behind the view */
  int *asAtom, *asCentre, *asType;
  double *asR;
  int64_t *asOmOff;
  int *ssShell, *ssASlot, *ssStart, *ssEnd;
  int64_t *ssFOff;
  int *trA, *trB;
  int64_t *trOut, *trPair;
  int *clsFirst;
  int64_t *clsWork, *clsElem, *clsOutElem, *clsPairBase, *clsQBase;
  int capAS, capSS, capTR, capOut;
  /* device enumeration (EcpBatch.devEnum) */
  int *ceAS0, *asSS0;
  int64_t *cePair0;
  unsigned char *ssOwn;
  int capCE, capAS0, capOwn;
  double triPerPair; /* executed triples per tested shell pair of the batches run so far (batch sizing; api.c feeds it back) */
  /* canonical (reference loop order) list of the executed triples of this batch, for callback replay */
  int nCanon, capCanon;
  int *cnA, *cnS1, *cnB, *cnS2, *cnC, *cnLa, *cnLb; /* cnLa / cnLb: the UNSHIFTED momenta (callback arguments)        */
  int *cnShA, *cnShB;                               /* momentum shifts of a derivative run (0 otherwise)                */
  int64_t *cnOut;
  /* statistics */
  long long nominal, screenedShells;
  void *scratch;      /* per-centre work areas of the parallel builder (builder.c) */
} EcpBatchBuf;

EcpBatchBuf *ecp_batch_new(const EcpTables *t);
void ecp_batch_free(EcpBatchBuf *bb);
/* let dst use the per-centre work areas of src (two buffers that build alternately) */
void ecp_batch_share_scratch(EcpBatchBuf *dst, EcpBatchBuf *src);

/* Build one batch starting at atom index *centre (advanced past the centres consumed).  Stops adding centres
 * once the batch holds >= maxTriples triples (always takes at least one centre).  Only shell pairs owned by
 * (rank, world) are emitted (world = 1: everything).  keepCanon != 0 records the canonical list; wantOut != 0 fills the block offsets trOut.
 * Returns the number of centres consumed (0 = no ECP centre left). */
int ecp_batch_build(const EcpTables *t, const double *geometry, int *centre, long long maxTriples, int rank, int world,
                    int keepCanon, int wantOut, EcpBatchBuf *bb);
/* The same for a device-enumerated batch: screening and slot layout only (O(centres x shells) instead of O(triples));
 * the batch is cut by an estimate of the triples (tested shell pairs x bb->triPerPair). */
int ecp_batch_build_slots(const EcpTables *t, const double *geometry, int *centre, long long maxTriples, int rank, int world,
                          EcpBatchBuf *bb);

/* derivative runs (builder.c): copies of a shell in the expanded list, and shift s of the reference's table - whether the
 * reference evaluates it, which copies it pairs and the shifts the callback receives */
int ecp_deriv_ncopies(int n);
void ecp_deriv_copy(int n, int c, int *dl, int *zpow);
int ecp_deriv_nshifts(int n);
int ecp_deriv_shift(int n, int s, int la, int lb, int aOnC, int bOnC, int *a2, int *b2, int *sa, int *sb);
/* last centre (atom order) whose screening can keep a shell of each atom; -1 = none (builder.c) */
void ecp_atom_last_centre(const EcpTables *t, const double *geometry, int *lastC);
/* exposed for tests: window of one (centre type, shell radius, distance) (reference src/type2.c:148-180) */
void ecp_shell_window(const EcpTables *t, int endLast, double radius, double dist, int *start, int *end, int *skip);
/* owner rank of a shell pair under the multi-GPU partition */
int ecp_pair_owner(const EcpTables *t, int shellA, int shellB, int world);

#endif
