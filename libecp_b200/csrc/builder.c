/* builder.c - host batch builder (see builder.h).
 *
 * Three phases per batch, the two heavy ones parallel over ECP centres (OpenMP):
 *   (a) per centre: screening windows of every shell (atom-level prune first), list of unskipped shells,
 *       count of executed triples / primitive pairs per class                       [parallel]
 *   (b) prefix sums -> slot, table and per-class positions for every centre         [serial, O(centres x classes)]
 *   (c) per centre: write slots and triples straight into their class-sorted place  [parallel]
 * The result is independent of the thread count: within a class, triples are ordered by centre and then in
 * the reference's loop order (A, B>=A, s1, s2; src/libecp.c:278-320).
 */
#include "builder.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int CD(int l) { return (l + 1) * (l + 2) * (l + 3) / 6; }
static int IJK(int l) { return (l + 1) * (l + 2) / 2; }

typedef struct {
  int C, type, Lc, nSS, nAS;
  int *ssShell, *ssStart, *ssEnd, *ssAtom; /* ssAtom: local atom-slot index               */
  unsigned char *ssL, *ssK, *ssOwn;        /* angular momentum, contraction depth, row owned by this rank */
  int *asAtom;
  double *asD;                             /* distance d_XC per local atom slot            */
  int *clsCount;                           /* executed triples per class                   */
  long long *clsPairs;                     /* primitive pairs per class                    */
  long long nTri, outSize, fRows, omSize;
  int cap;
} CentreWork;

struct BuilderScratch {
  CentreWork *cw;
  int ncw;
  int refs;                 /* batch buffers sharing this scratch (ecp_batch_share_scratch)                    */
  /* centres screened by the previous call that did not fit into its batch: cw[0..nleft) */
  int nleft, leftCentre[1024], leftNext, leftRank, leftWorld;
};

EcpBatchBuf *ecp_batch_new(const EcpTables *t) {
  EcpBatchBuf *bb = calloc(1, sizeof(EcpBatchBuf));
  const int nc = t->v.nClasses;
  bb->clsFirst = calloc(nc + 2, sizeof(int));
  bb->clsWork = calloc(nc + 2, sizeof(int64_t));
  bb->clsElem = calloc(nc + 2, sizeof(int64_t));
  bb->clsOutElem = calloc(nc + 2, sizeof(int64_t));
  bb->clsPairBase = calloc(nc + 2, sizeof(int64_t));
  bb->clsQBase = calloc(nc + 2, sizeof(int64_t));
  return bb;
}

static void free_scratch(struct BuilderScratch *s) {
  if (!s) return;
  if (--s->refs > 0) return;
  for (int i = 0; i < s->ncw; i++) {
    CentreWork *w = &s->cw[i];
    free(w->ssShell); free(w->ssStart); free(w->ssEnd); free(w->ssAtom); free(w->asAtom); free(w->asD);
    free(w->ssL); free(w->ssK); free(w->ssOwn);
    free(w->clsCount); free(w->clsPairs);
  }
  free(s->cw);
  free(s);
}

static struct BuilderScratch *get_scratch(EcpBatchBuf *bb) {
  if (!bb->scratch) {
    struct BuilderScratch *S = calloc(1, sizeof(struct BuilderScratch));
    S->refs = 1;
    bb->scratch = S;
  }
  return (struct BuilderScratch *)bb->scratch;
}
/* the two batch buffers of a pipeline build alternately, never at the same time: with one scratch the centres the
 * previous call screened beyond its batch are not screened again */
void ecp_batch_share_scratch(EcpBatchBuf *dst, EcpBatchBuf *src) {
  struct BuilderScratch *S = get_scratch(src);
  if (dst->scratch == S) return;
  free_scratch((struct BuilderScratch *)dst->scratch);
  dst->scratch = S;
  S->refs++;
}

void ecp_batch_free(EcpBatchBuf *bb) {
  if (!bb) return;
  free(bb->asAtom); free(bb->asCentre); free(bb->asType); free(bb->asR); free(bb->asOmOff);
  free(bb->ssShell); free(bb->ssASlot); free(bb->ssStart); free(bb->ssEnd); free(bb->ssFOff);
  ecpdev_pinned_free(bb->trA); ecpdev_pinned_free(bb->trB); ecpdev_pinned_free(bb->trOut); ecpdev_pinned_free(bb->trPair);
 
  free(bb->clsFirst); free(bb->clsWork); free(bb->clsElem); free(bb->clsOutElem); free(bb->clsPairBase); free(bb->clsQBase);
  free(bb->cnA); free(bb->cnS1); free(bb->cnB); free(bb->cnS2); free(bb->cnC); free(bb->cnLa); free(bb->cnLb);
  free(bb->cnOut); free(bb->cnShA); free(bb->cnShB);
  free(bb->ceAS0); free(bb->cePair0); free(bb->asSS0); free(bb->ssOwn);
  free_scratch((struct BuilderScratch *)bb->scratch);
  free(bb);
}

/* basis-set screening of one shell against the (potential-capped) small grid: first grid point with
 * r >= d-R and last with r <= d+R (reference src/type2.c:148-180 scans downwards from the cap; the
 * KK-mapped abscissae increase strictly, so both scans are binary searches with the same result). */
void ecp_shell_window(const EcpTables *t, int endLast, double radius, double dist, int *start, int *end, int *skip) {
  const double rmin = dist - radius, rmax = dist + radius;
  const double *r = t->small_x;
  /* winLut brackets both answers to the few points of one 1/16-bohr bin; the searches inside the bracket compare the
   * same doubles as a search over the whole grid would */
  int lo, hi; /* largest j <= endLast with r[j] < rmin (or -1) */
  if (!(rmin > 0.0)) {
    lo = hi = -1;
  } else if (rmin * ECP_WIN_LUT_SCALE >= ECP_WIN_LUT_BINS) {
    lo = -1;
    hi = endLast;
  } else {
    const int k = (int)(rmin * ECP_WIN_LUT_SCALE);
    lo = t->winLut[k] - 1;
    hi = t->winLut[k + 1] - 1;
    if (hi > endLast) hi = endLast;
    if (lo > endLast) lo = endLast;
  }
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (r[mid] < rmin)
      lo = mid;
    else
      hi = mid - 1;
  }
  *start = lo + 1;
  /* largest j <= endLast with r[j] <= rmax (or -1) */
  if (rmax < 0.0) {
    lo = hi = -1;
  } else if (rmax * ECP_WIN_LUT_SCALE >= ECP_WIN_LUT_BINS) {
    lo = -1;
    hi = endLast;
  } else {
    const int k = (int)(rmax * ECP_WIN_LUT_SCALE);
    lo = t->winLut[k] - 1;
    hi = t->winLut[k + 1] - 1;
    if (hi > endLast) hi = endLast;
    if (lo > endLast) lo = endLast;
  }
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (r[mid] <= rmax)
      lo = mid;
    else
      hi = mid - 1;
  }
  *end = lo;
  *skip = !(*end >= *start);
}

/* Output ownership for the multi-GPU partition: the block of shell pair (a, b), a <= b in global shell order,
 * belongs to the rank that owns ROW shell a.  Owning whole rows means that a rank's builder enumerates only its rows
 * (work / world) and that its partial matrix has non-zeros only in the AO rows of its shells (D2H / world).
 * Cost-model deal: the cost of a triple is set by the kind of its shells (l, contraction depth, exponents).  The rows
 * are dealt to the ranks one by one in the order "by kind, pseudo-random inside a kind" (t->rowDeal, tables.c): every
 * rank receives the same number (+-1) of rows of every kind, from random places of the molecule. */
int ecp_pair_owner(const EcpTables *t, int a, int b, int world) {
  if (world <= 1) return 0;
  const int row = a < b ? a : b;
  const int *deal = t->rowDeal ? t->rowDeal : ecp_tables_row_deal((EcpTables *)t);
  return deal[row] % world;
}

static double dist3(const double *a, const double *b) { /* src/util.c:109-116 */
  const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
  return sqrt(x * x + y * y + z * z);
}

/* Copies of a shell in the expanded list of a derivative run (api.c), behind the caller's shell (copy 0): momentum change
 * and power of zeta in the contraction coefficients (reference src/type1.c:239-246, src/type2.c:263-269).  The lowered
 * copies come last: a shell with l < 1 / l < 2 simply has fewer copies. */
int ecp_deriv_ncopies(int n) { return n == 2 ? 6 : (n == 1 ? 3 : 1); }
void ecp_deriv_copy(int n, int c, int *dl, int *zpow) {
  static const int dl1[3] = {0, +1, -1}, zp1[3] = {0, 1, 0};
  static const int dl2[6] = {0, +1, +2, 0, -1, -2}, zp2[6] = {0, 1, 2, 1, 0, 0};
  if (n == 2) {
    *dl = dl2[c];
    *zpow = zp2[c];
  } else {
    *dl = dl1[c];
    *zpow = zp1[c];
  }
}
int ecp_deriv_nshifts(int n) { return n == 2 ? 10 : (n == 1 ? 4 : 1); }
/* Shift s of a derivative run of order n on (la, lb): the reference's table src/libecp.c:246-250.  Returns 0 when the
 * reference skips it (:325-330): momentum below zero, or - first derivatives only - the shifted function sits on the ECP
 * centre (translational invariance).  *a2 / *b2 = slot offset of the shifted copy behind its unshifted shell
 * (ecp_deriv_copy), *sa / *sb = the shifts handed to the callback.  Second derivatives, shifts (+1,0) and (0,+1)
 * (:362-369): the copy with unchanged momentum and coefficients d zeta. */
int ecp_deriv_shift(int n, int s, int la, int lb, int aOnC, int bOnC, int *a2, int *b2, int *sa, int *sb) {
  static const int sa1[4] = {+1, -1, 0, 0}, sb1[4] = {0, 0, +1, -1};
  static const int sa2[10] = {+2, -2, +1, +1, -1, -1, 0, 0, +1, 0}, sb2[10] = {0, 0, +1, -1, +1, -1, +2, -2, 0, +1};
  const int shifta = n == 2 ? sa2[s] : sa1[s], shiftb = n == 2 ? sb2[s] : sb1[s];
  if (la < -shifta || lb < -shiftb) return 0;
  *sa = shifta;
  *sb = shiftb;
  if (n == 2) {
    static const int slot[5] = {5, 4, 0, 1, 2}; /* shift -2 .. +2 */
    *a2 = slot[shifta + 2];
    *b2 = slot[shiftb + 2];
    if (s == 8) *a2 = 3;
    if (s == 9) *b2 = 3;
    return 1;
  }
  if ((aOnC && shifta) || (bOnC && shiftb)) return 0;
  *a2 = shifta > 0 ? 1 : (shifta < 0 ? 2 : 0);
  *b2 = shiftb > 0 ? 1 : (shiftb < 0 ? 2 : 0);
  return 1;
}

/* Last ECP centre (in processing order = atom order) that can touch the rows of each atom: the atom-level prune of
 * centre_screen below, evaluated for every (atom, centre).  lastC[X] = -1: no centre reaches the atom (its rows stay
 * zero).  Once the pass is beyond lastC[X] the AO rows of X are final - the host consumer streams them out while later
 * centres are still being integrated (api.c). */
void ecp_atom_last_centre(const EcpTables *t, const double *geometry, int *lastC) {
  const int nat = t->v.nrAtoms;
  for (int X = 0; X < nat; X++) lastC[X] = -1;
  for (int C = 0; C < nat; C++) {
    if (t->atomType[C] < 0) continue;
    const int endLast = t->types[t->atomType[C]].endLast;
    if (endLast < 0) continue;
    const double rcap = t->small_x[endLast];
    for (int X = 0; X < nat; X++) {
      if (t->atomFirstShell[X] == t->atomFirstShell[X + 1]) continue;
      if (dist3(geometry + 3 * C, geometry + 3 * X) - t->atomRmax[X] > rcap) continue;
      lastC[X] = C;
    }
  }
}

/* phase (a) for one centre */
static void centre_screen(const EcpTables *t, const double *geometry, int C, int rank, int world, CentreWork *w, int doCount) {
  const EcpHostTables *v = &t->v;
  const int nat = v->nrAtoms, nc = v->nClasses, nsh = v->nrShells;
  if (!w->asAtom) { /* per-atom and per-class areas have fixed sizes; the per-slot ones grow with the slots found */
    w->asAtom = realloc(w->asAtom, (nat + 1) * sizeof(int));
    w->asD = realloc(w->asD, (nat + 1) * sizeof(double));
    w->clsCount = realloc(w->clsCount, (nc + 1) * sizeof(int));
    w->clsPairs = realloc(w->clsPairs, (nc + 1) * sizeof(long long));
  }
  (void)nsh;
  w->C = C;
  w->type = t->atomType[C];
  const EcpType *T = &t->types[w->type];
  const int Lc = T->L, endLast = T->endLast;
  w->Lc = Lc;
  const double *rC = geometry + 3 * C;
  const double rcap = (endLast >= 0) ? t->small_x[endLast] : -1.0;
  int nSS = 0, nAS = 0;
  long long fRows = 0, omSize = 0;
  for (int X = 0; X < nat; X++) {
    const int s0 = t->atomFirstShell[X], s1 = t->atomFirstShell[X + 1];
    if (s0 == s1 || endLast < 0) continue;
    const double d = dist3(rC, geometry + 3 * X);
    /* atom-level prune: d - R > r[cap] for every shell => start = cap+1 > end => skipShell */
    if (d - t->atomRmax[X] > rcap) continue;
    int aslot = -1;
    for (int s = s0; s < s1; s++) {
      int st, en, sk;
      ecp_shell_window(t, endLast, t->shellRadius[s], d, &st, &en, &sk);
      if (sk) continue;
      if (aslot < 0) {
        aslot = nAS++;
        w->asAtom[aslot] = X;
        w->asD[aslot] = d;
        omSize += (long long)(Lc + t->atomMaxL[X]) * Lc * Lc * CD(t->atomMaxL[X]);
      }
      if (nSS == w->cap) {
        w->cap = w->cap ? 2 * w->cap : 512;
        w->ssShell = realloc(w->ssShell, w->cap * sizeof(int));
        w->ssStart = realloc(w->ssStart, w->cap * sizeof(int));
        w->ssEnd = realloc(w->ssEnd, w->cap * sizeof(int));
        w->ssAtom = realloc(w->ssAtom, w->cap * sizeof(int));
        w->ssL = realloc(w->ssL, w->cap);
        w->ssK = realloc(w->ssK, w->cap);
        w->ssOwn = realloc(w->ssOwn, w->cap);
      }
      w->ssShell[nSS] = s;
      w->ssStart[nSS] = st;
      w->ssEnd[nSS] = en;
      w->ssAtom[nSS] = aslot;
      w->ssL[nSS] = (unsigned char)v->shellL[s];
      w->ssK[nSS] = (unsigned char)v->shellK[s];
      w->ssOwn[nSS] = (unsigned char)(world <= 1 || ecp_pair_owner(t, s, s, world) == rank); /* row ownership */
      fRows += Lc + t->shellL[s];
      nSS++;
    }
  }
  w->nSS = nSS;
  w->nAS = nAS;
  w->fRows = fRows;
  w->omSize = omSize;
  if (!doCount) { /* device enumeration: only the number of shell pairs (owned a, b >= a) the device will test */
    long long cand = 0;
    for (int a = 0; a < nSS; a++)
      if (w->ssOwn[a]) cand += nSS - a;
    w->nTri = cand;
    w->outSize = 0;
    return;
  }
  /* count executed triples per class (canonical enumeration, src/libecp.c:297-320,344) */
  memset(w->clsCount, 0, (nc + 1) * sizeof(int));
  memset(w->clsPairs, 0, (nc + 1) * sizeof(long long));
  long long nTri = 0, outSize = 0;
  if (t->deriv) { /* first derivatives: up to four shifted triples per executed shell pair (src/libecp.c:246-250,322-330) */
    const int *st = w->ssStart, *en = w->ssEnd;
    const unsigned char *sl = w->ssL, *sk = w->ssK;
    for (int a = 0; a < nSS; a++) {
      if (t->virtShift[w->ssShell[a]]) continue;
      const int A = w->asAtom[w->ssAtom[a]];
      for (int b = a; b < nSS; b++) {
        if (t->virtShift[w->ssShell[b]]) continue;
        const int B = w->asAtom[w->ssAtom[b]];
        if (A == C && B == C) continue; /* src/libecp.c:305 */
        const int gs = st[a] > st[b] ? st[a] : st[b];
        const int ge = en[a] > en[b] ? en[a] : en[b];
        if (!(gs < ge)) continue;
        for (int sh = 0; sh < ecp_deriv_nshifts(t->deriv); sh++) {
          int a2, b2, sha, shb;
          if (!ecp_deriv_shift(t->deriv, sh, sl[a], sl[b], A == C, B == C, &a2, &b2, &sha, &shb)) continue;
          const int la = sl[a + a2], lb = sl[b + b2];
          const int c = t->clsLookup[la][lb][Lc];
          w->clsCount[c] += 1;
          w->clsPairs[c] += (long long)sk[a] * sk[b];
          outSize += 2LL * IJK(la) * IJK(lb);
          nTri++;
        }
      }
    }
  } else {
    const int *st = w->ssStart, *en = w->ssEnd;
    const unsigned char *sl = w->ssL, *sk = w->ssK;
    for (int a = 0; a < nSS; a++) {
      if (!w->ssOwn[a]) continue;
      const int la = sl[a], Ka = sk[a], sta = st[a], ena = en[a];
      int cnt[ECP_MAX_LBS + 1] = {0}, prs[ECP_MAX_LBS + 1] = {0};
      for (int b = a; b < nSS; b++) {
        const int gs = sta > st[b] ? sta : st[b];
        const int ge = ena > en[b] ? ena : en[b];
        const int ok = gs < ge; /* src/libecp.c:344 */
        cnt[sl[b]] += ok;
        prs[sl[b]] += ok ? sk[b] : 0;
      }
      for (int lb = 0; lb <= v->maxLBS; lb++) {
        if (!cnt[lb]) continue;
        const int c = t->clsLookup[la][lb][Lc];
        w->clsCount[c] += cnt[lb];
        w->clsPairs[c] += (long long)Ka * prs[lb];
        outSize += 2LL * IJK(la) * IJK(lb) * cnt[lb];
        nTri += cnt[lb];
      }
    }
  }
  w->nTri = nTri;
  w->outSize = outSize;
}

#define ENSURE(ptr, cap, need, type)                         \
  do {                                                       \
    if ((long long)(need) > (long long)(cap))                \
      (ptr) = (type *)realloc((ptr), (size_t)((need) + 16) * sizeof(type)); \
  } while (0)

/* arrays that are uploaded every batch live in page-locked memory (contents need not survive a regrow) */
#define ENSURE_PIN(ptr, cap, need, type)                                                        \
  do {                                                                                          \
    if ((long long)(need) > (long long)(cap)) {                                                 \
      ecpdev_pinned_free(ptr);                                                                  \
      (ptr) = (type *)ecpdev_pinned_alloc((size_t)((need) + (need) / 4 + 16) * sizeof(type));   \
    }                                                                                           \
  } while (0)

int ecp_batch_build(const EcpTables *t, const double *geometry, int *centre, long long maxTriples, int rank, int world,
                    int keepCanon, int wantOut, EcpBatchBuf *bb) {
  const int needOut = wantOut || keepCanon; /* block offsets are only needed when the callback blocks are produced */
  const EcpHostTables *v = &t->v;
  const int nat = v->nrAtoms, nc = v->nClasses;
  if (world > 1) ecp_tables_row_deal((EcpTables *)t); /* before the parallel regions */
  struct BuilderScratch *S = get_scratch(bb);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  /* candidate centres are screened in parallel, a round of a few per thread at a time, until the batch is full
   * (a rank that owns 1/8 of the rows needs 8x the centres for the same batch size); then as many as fit into
   * maxTriples are taken.  Screening results of centres that do not fit are recomputed by the next call. */
  const int prof = getenv("LIBECP_B200_BUILD_PROFILE") != NULL;
  const double tp0 = prof ? omp_get_wtime() : 0.0;
  const int round = nthreads * 2 > 8 ? nthreads * 2 : 8;
  int cand[1024], ncand = 0, C = *centre;
  long long screened = 0;
  CentreWork *cw = S->cw;
  if (S->nleft > 0 && S->leftCentre[0] == *centre && S->leftRank == rank && S->leftWorld == world) {
    for (int i = 0; i < S->nleft; i++) { /* screened by the previous call */
      cand[ncand++] = S->leftCentre[i];
      screened += cw[i].nTri;
    }
    C = S->leftNext;
  }
  S->nleft = 0;
  while (C < nat && ncand < 1024 && screened < maxTriples) {
    const int first = ncand;
    for (; C < nat && ncand < first + round && ncand < 1024; C++)
      if (t->atomType[C] >= 0) cand[ncand++] = C;
    if (ncand == first) break;
    if (S->ncw < ncand) {
      S->cw = realloc(S->cw, ncand * sizeof(CentreWork));
      memset(S->cw + S->ncw, 0, (ncand - S->ncw) * sizeof(CentreWork));
      S->ncw = ncand;
    }
    cw = S->cw;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = first; i < ncand; i++) centre_screen(t, geometry, cand[i], rank, world, &cw[i], 1);
    for (int i = first; i < ncand; i++) screened += cw[i].nTri;
  }
  if (ncand == 0) {
    *centre = nat;
    bb->b.nTriples = 0;
    return 0;
  }

  const double tp1 = prof ? omp_get_wtime() : 0.0;
  /* ---- phase (b): how many centres, and where everything goes ---- */
  int ntake = 0;
  long long tri = 0;
  while (ntake < ncand) {
    tri += cw[ntake].nTri;
    ntake++;
    if (tri >= maxTriples) break;
  }
  if (ntake < ncand && C >= nat) { /* every remaining centre is screened: a small rest joins this batch instead of
                                    * paying a batch's launches for a few thousand triples */
    long long rest = 0;
    for (int i = ntake; i < ncand; i++) rest += cw[i].nTri;
    if (4 * rest <= maxTriples) ntake = ncand;
  }
  *centre = (ntake < ncand) ? cand[ntake] : C;
  long long nAS = 0, nSS = 0, omTotal = 0, fRows = 0, outTotal = 0, nTR = 0, nPairs = 0;
  long long *asBase = malloc((ntake + 1) * sizeof(long long)), *ssBase = malloc((ntake + 1) * sizeof(long long));
  long long *omBase = malloc((ntake + 1) * sizeof(long long)), *fBase = malloc((ntake + 1) * sizeof(long long));
  long long *outBase = malloc((ntake + 1) * sizeof(long long)), *cnBase = malloc((ntake + 1) * sizeof(long long));
  long long *posCC = malloc((size_t)(nc + 1) * ntake * sizeof(long long));  /* [class][centre] first triple   */
  long long *pairCC = malloc((size_t)(nc + 1) * ntake * sizeof(long long)); /* [class][centre] first pair     */
  long long *tBase = malloc((nc + 1) * sizeof(long long)), *gBase = malloc((nc + 1) * sizeof(long long));
  long long *pairBase = malloc((nc + 2) * sizeof(long long)), *qBase = malloc((nc + 1) * sizeof(long long));
  bb->nominal = 0;
  for (int i = 0; i < ntake; i++) {
    asBase[i] = nAS; ssBase[i] = nSS; omBase[i] = omTotal; fBase[i] = fRows; outBase[i] = outTotal; cnBase[i] = nTR;
    nAS += cw[i].nAS; nSS += cw[i].nSS; omTotal += cw[i].omSize; fRows += cw[i].fRows; outTotal += cw[i].outSize;
    nTR += cw[i].nTri;
    bb->nominal += (long long)v->nrShells * (v->nrShells + 1) / 2;
  }
  bb->screenedShells = nSS;
  long long tTot = 0, gTot = 0, qTot = 0;
  bb->clsFirst[0] = 0;
  bb->clsWork[0] = bb->clsElem[0] = bb->clsOutElem[0] = 0;
  for (int c = 0; c < nc; c++) {
    const int la = v->clsLa[c], lb = v->clsLb[c], lab = la + lb;
    long long n = 0, np = 0;
    pairBase[c] = nPairs;
    for (int i = 0; i < ntake; i++) {
      posCC[(size_t)c * ntake + i] = bb->clsFirst[c] + n;
      pairCC[(size_t)c * ntake + i] = nPairs + np;
      n += cw[i].clsCount[c];
      np += cw[i].clsPairs[c];
    }
    bb->clsFirst[c + 1] = bb->clsFirst[c] + (int)n;
    bb->clsWork[c + 1] = bb->clsWork[c] + n * v->clsNq[c];
    bb->clsElem[c + 1] = bb->clsElem[c] + n * CD(la) * CD(lb);
    bb->clsOutElem[c + 1] = bb->clsOutElem[c] + n * IJK(la) * IJK(lb);
    tBase[c] = tTot;
    gBase[c] = gTot;
    qBase[c] = qTot;
    tTot += n * v->clsNq[c];
    gTot += n * CD(la) * CD(lb);
    qTot += np * (lab + 1) * (lab + 1);
    nPairs += np;
  }
  pairBase[nc] = nPairs;
  for (int c = 0; c <= nc; c++) {
    bb->clsPairBase[c] = pairBase[c];
    bb->clsQBase[c] = (c < nc) ? qBase[c] : qTot;
  }
  /* storage */
  ENSURE(bb->asAtom, bb->capAS, nAS, int); ENSURE(bb->asCentre, bb->capAS, nAS, int);
  ENSURE(bb->asType, bb->capAS, nAS, int);
  if (nAS > bb->capAS) bb->asR = (double *)realloc(bb->asR, (size_t)(nAS + 16) * 4 * sizeof(double)); /* 4 per slot */
  ENSURE(bb->asOmOff, bb->capAS, nAS, int64_t);
  if (nAS > bb->capAS) bb->capAS = (int)nAS + 16;
  ENSURE(bb->ssShell, bb->capSS, nSS, int); ENSURE(bb->ssASlot, bb->capSS, nSS, int);
  ENSURE(bb->ssStart, bb->capSS, nSS, int); ENSURE(bb->ssEnd, bb->capSS, nSS, int);
  ENSURE(bb->ssFOff, bb->capSS, nSS, int64_t);
  if (nSS > bb->capSS) bb->capSS = (int)nSS + 16;
  ENSURE_PIN(bb->trA, bb->capTR, nTR, int); ENSURE_PIN(bb->trB, bb->capTR, nTR, int);
  ENSURE_PIN(bb->trOut, bb->capOut, needOut ? nTR : 0, int64_t); ENSURE_PIN(bb->trPair, bb->capTR, nTR, int64_t);
  if (needOut && nTR > bb->capOut) bb->capOut = (int)(nTR + nTR / 4) + 16;
  if (nTR > bb->capTR) bb->capTR = (int)(nTR + nTR / 4) + 16;
  if (keepCanon) {
    ENSURE(bb->cnA, bb->capCanon, nTR, int); ENSURE(bb->cnS1, bb->capCanon, nTR, int);
    ENSURE(bb->cnB, bb->capCanon, nTR, int); ENSURE(bb->cnS2, bb->capCanon, nTR, int);
    ENSURE(bb->cnC, bb->capCanon, nTR, int); ENSURE(bb->cnLa, bb->capCanon, nTR, int);
    ENSURE(bb->cnLb, bb->capCanon, nTR, int); ENSURE(bb->cnOut, bb->capCanon, nTR, int64_t);
    ENSURE(bb->cnShA, bb->capCanon, nTR, int); ENSURE(bb->cnShB, bb->capCanon, nTR, int);
    if (nTR > bb->capCanon) bb->capCanon = (int)nTR + 16;
  }
  bb->nCanon = keepCanon ? (int)nTR : 0;

  const double tp2 = prof ? omp_get_wtime() : 0.0;
  /* ---- phase (c): fill, parallel over centres ---- */
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < ntake; i++) {
    const CentreWork *w = &cw[i];
    const double *rC = geometry + 3 * w->C;
    const int Lc = w->Lc;
    long long om = omBase[i], fr = fBase[i];
    for (int a = 0; a < w->nAS; a++) {
      const long long g = asBase[i] + a;
      const int X = w->asAtom[a];
      bb->asAtom[g] = X;
      bb->asCentre[g] = w->C;
      bb->asType[g] = w->type;
      /* r_XC = X - C (reference distanceVector(rAC, rC, rA), src/libecp.c:281) */
      bb->asR[4 * g + 0] = geometry[3 * X + 0] - rC[0];
      bb->asR[4 * g + 1] = geometry[3 * X + 1] - rC[1];
      bb->asR[4 * g + 2] = geometry[3 * X + 2] - rC[2];
      bb->asR[4 * g + 3] = w->asD[a];
      bb->asOmOff[g] = om;
      om += (long long)(Lc + t->atomMaxL[X]) * Lc * Lc * CD(t->atomMaxL[X]);
    }
    for (int s = 0; s < w->nSS; s++) {
      const long long g = ssBase[i] + s;
      bb->ssShell[g] = w->ssShell[s];
      bb->ssASlot[g] = (int)(asBase[i] + w->ssAtom[s]);
      bb->ssStart[g] = w->ssStart[s];
      bb->ssEnd[g] = w->ssEnd[s];
      bb->ssFOff[g] = fr;
      fr += Lc + v->shellL[w->ssShell[s]];
    }
    /* canonical enumeration: A, B>=A, s1, s2 (reference src/libecp.c:278-320); slots of one atom are contiguous */
    long long *lpos = calloc(2 * (nc + 1), sizeof(long long)), *lpair = lpos + nc + 1;
    long long out = outBase[i], cn = cnBase[i];
    const int *st = w->ssStart, *en = w->ssEnd;
    const unsigned char *sl = w->ssL, *sk = w->ssK;
    const int sbase = (int)ssBase[i];
    for (int c = 0; c < nc; c++) {
      lpos[c] = posCC[(size_t)c * ntake + i];
      lpair[c] = pairCC[(size_t)c * ntake + i];
    }
    /* atom blocks [bs[k], bs[k+1]) and, per slot, the end of its run of equal l inside the block (shells of an atom
     * are l-ascending): a run of partners b shares the class, so its triples go to consecutive places */
    const int nSSw = w->nSS;
    int *bs = malloc((size_t)(w->nAS + 2) * sizeof(int)), *runEnd = malloc((size_t)(nSSw + 1) * sizeof(int));
    int nblk = 0;
    for (int s = 0; s < nSSw; s++)
      if (s == 0 || w->ssAtom[s] != w->ssAtom[s - 1]) bs[nblk++] = s;
    bs[nblk] = nSSw;
    for (int k = 0; k < nblk; k++)
      for (int s = bs[k + 1] - 1, e = bs[k + 1]; s >= bs[k]; s--) {
        if (s + 1 < bs[k + 1] && sl[s + 1] != sl[s]) e = s + 1;
        runEnd[s] = e;
      }
    if (t->deriv) { /* shifted triples in the reference's order: A, B, s1, s2, shift (src/libecp.c:278-330) */
      for (int ka = 0; ka < nblk; ka++)
      for (int kb = ka; kb < nblk; kb++)
      for (int a = bs[ka]; a < bs[ka + 1]; a++) {
        if (t->virtShift[w->ssShell[a]]) continue;
        const int A = w->asAtom[w->ssAtom[a]];
        for (int b = (kb == ka ? a : bs[kb]); b < bs[kb + 1]; b++) {
          if (t->virtShift[w->ssShell[b]]) continue;
          const int B = w->asAtom[w->ssAtom[b]];
          if (A == w->C && B == w->C) continue;
          const int gs = st[a] > st[b] ? st[a] : st[b];
          const int ge = en[a] > en[b] ? en[a] : en[b];
          if (!(gs < ge)) continue;
          for (int sh = 0; sh < ecp_deriv_nshifts(t->deriv); sh++) {
            int a2, b2, sha, shb;
            if (!ecp_deriv_shift(t->deriv, sh, sl[a], sl[b], A == w->C, B == w->C, &a2, &b2, &sha, &shb)) continue;
            const int la = sl[a + a2], lb = sl[b + b2];
            const int c = t->clsLookup[la][lb][Lc];
            const long long p = lpos[c]++;
            bb->trA[p] = sbase + a + a2;
            bb->trB[p] = sbase + b + b2;
            bb->trPair[p] = lpair[c];
            lpair[c] += sk[a] * sk[b];
            if (needOut) {
              bb->trOut[p] = out;
              if (keepCanon) {
                bb->cnA[cn] = A;
                bb->cnS1[cn] = t->virtLocal[w->ssShell[a]];
                bb->cnB[cn] = B;
                bb->cnS2[cn] = t->virtLocal[w->ssShell[b]];
                bb->cnC[cn] = w->C;
                bb->cnLa[cn] = sl[a];
                bb->cnLb[cn] = sl[b];
                bb->cnShA[cn] = sha;
                bb->cnShB[cn] = shb;
                bb->cnOut[cn] = out;
                cn++;
              }
              out += 2 * (long long)IJK(la) * IJK(lb);
            }
          }
        }
      }
    } else
    for (int ka = 0; ka < nblk; ka++) {
      const int ia = bs[ka], ia1 = bs[ka + 1];
      int anyOwn = 0;
      for (int a = ia; a < ia1; a++) anyOwn |= w->ssOwn[a];
      if (!anyOwn) continue; /* row ownership: nothing of this atom belongs to the rank */
      for (int kb = ka; kb < nblk; kb++) {
        const int ib = bs[kb], ib1 = bs[kb + 1];
        for (int a = ia; a < ia1; a++) {
          if (!w->ssOwn[a]) continue;
          const int la = sl[a], Ka = sk[a], sta = st[a], ena = en[a];
          const int *lut = &t->clsLookup[la][0][0] + Lc;
          int b = (kb == ka ? a : ib);
          while (b < ib1) {
            const int e = runEnd[b], lb = sl[b];
            const int c = lut[lb * (ECP_MAX_LECP + 1)];
            long long p = lpos[c], pr = lpair[c];
            for (; b < e; b++) {
              const int gs = sta > st[b] ? sta : st[b];
              const int ge = ena > en[b] ? ena : en[b];
              if (!(gs < ge)) continue; /* src/libecp.c:344, identical for both types */
              bb->trA[p] = sbase + a;
              bb->trB[p] = sbase + b;
              bb->trPair[p] = pr;
              pr += Ka * sk[b];
              if (needOut) {
                bb->trOut[p] = out;
                if (keepCanon) {
                  const int sa = w->ssShell[a], sb = w->ssShell[b];
                  const int A = w->asAtom[w->ssAtom[a]], B = w->asAtom[w->ssAtom[b]];
                  bb->cnA[cn] = A;
                  bb->cnS1[cn] = sa - t->atomFirstShell[A];
                  bb->cnB[cn] = B;
                  bb->cnS2[cn] = sb - t->atomFirstShell[B];
                  bb->cnC[cn] = w->C;
                  bb->cnLa[cn] = la;
                  bb->cnLb[cn] = lb;
                  bb->cnShA[cn] = bb->cnShB[cn] = 0;
                  bb->cnOut[cn] = out;
                  cn++;
                }
                out += 2 * (long long)IJK(la) * IJK(lb);
              }
              p++;
            }
            lpos[c] = p;
            lpair[c] = pr;
          }
        }
      }
    }
    free(bs);
    free(runEnd);
    free(lpos);
  }
  free(asBase); free(ssBase); free(omBase); free(fBase); free(outBase); free(cnBase);
  free(posCC); free(pairCC); free(tBase); free(gBase); free(pairBase); free(qBase);

  if (prof)
    fprintf(stderr, "[builder] centres %d/%d triples %lld: screen+count %.2f ms, layout %.2f ms, fill %.2f ms\n", ntake, ncand,
            nTR, 1e3 * (tp1 - tp0), 1e3 * (tp2 - tp1), 1e3 * (omp_get_wtime() - tp2));
  /* the screened centres that did not fit stay at the front of the scratch for the next call */
  for (int i = ntake; i < ncand; i++) {
    const CentreWork tmp = cw[i - ntake];
    cw[i - ntake] = cw[i];
    cw[i] = tmp;
    S->leftCentre[i - ntake] = cand[i];
  }
  S->nleft = ncand - ntake;
  S->leftNext = C;
  S->leftRank = rank;
  S->leftWorld = world;

  EcpBatch *b = &bb->b;
  b->nASlots = (int)nAS;
  b->asAtom = bb->asAtom;
  b->asCentre = bb->asCentre;
  b->asType = bb->asType;
  b->asR = bb->asR;
  b->asOmOff = bb->asOmOff;
  b->omTotal = omTotal;
  b->nSSlots = (int)nSS;
  b->ssShell = bb->ssShell;
  b->ssASlot = bb->ssASlot;
  b->ssStart = bb->ssStart;
  b->ssEnd = bb->ssEnd;
  b->ssFOff = bb->ssFOff;
  b->fRows = fRows;
  b->nTriples = (int)nTR;
  b->trA = bb->trA;
  b->trB = bb->trB;
  b->trOut = bb->trOut;
  b->trPair = bb->trPair;
  b->tTotal = tTot;
  b->gTotal = gTot;
  b->outTotal = outTotal;
  b->nPairs = nPairs;
  b->qTotal = qTot;
  b->rshTotal = qTot;
  b->clsFirst = bb->clsFirst;
  b->clsWork = bb->clsWork;
  b->clsElem = bb->clsElem;
  b->clsOutElem = bb->clsOutElem;
  b->clsPairBase = bb->clsPairBase;
  b->clsQBase = bb->clsQBase;
  return ntake;
}


/* ---- device-enumerated batch: screening + slot layout (builder.h) ---- */
int ecp_batch_build_slots(const EcpTables *t, const double *geometry, int *centre, long long maxTriples, int rank, int world,
                          EcpBatchBuf *bb) {
  const EcpHostTables *v = &t->v;
  const int nat = v->nrAtoms;
  if (world > 1) ecp_tables_row_deal((EcpTables *)t);
  struct BuilderScratch *S = get_scratch(bb);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  const int prof = getenv("LIBECP_B200_BUILD_PROFILE") != NULL;
  const double tp0 = prof ? omp_get_wtime() : 0.0;
  /* executed triples per tested shell pair: measured on the batches run so far; before the first one the upper bound 1
   * (a batch can then only come out smaller than asked for) */
  const double ratio = bb->triPerPair > 0.0 ? bb->triPerPair : 1.0;
  /* the screening of a batch is a few hundred microseconds of work per thread: a small team, so that one thread losing
   * its core to another process (NCCL proxy, a sampler) cannot hold a large team at the loop's barrier for a scheduler
   * quantum - measured at 2 and 8 ranks: builds of 3 ms that took 10-30 ms now and then */
  if (nthreads > 4) nthreads = 4;
  const int round = nthreads * 2 > 8 ? nthreads * 2 : 8;
  int cand[1024], ncand = 0, C = *centre;
  double screened = 0.0; /* estimated triples of the screened centres */
  CentreWork *cw = S->cw;
  if (S->nleft > 0 && S->leftCentre[0] == *centre && S->leftRank == rank && S->leftWorld == world) {
    for (int i = 0; i < S->nleft; i++) {
      cand[ncand++] = S->leftCentre[i];
      screened += ratio * (double)cw[i].nTri;
    }
    C = S->leftNext;
  }
  S->nleft = 0;
  while (C < nat && ncand < 1024 && screened < (double)maxTriples) {
    const int first = ncand;
    for (; C < nat && ncand < first + round && ncand < 1024; C++)
      if (t->atomType[C] >= 0) cand[ncand++] = C;
    if (ncand == first) break;
    if (S->ncw < ncand) {
      S->cw = realloc(S->cw, ncand * sizeof(CentreWork));
      memset(S->cw + S->ncw, 0, (ncand - S->ncw) * sizeof(CentreWork));
      S->ncw = ncand;
    }
    cw = S->cw;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int i = first; i < ncand; i++) centre_screen(t, geometry, cand[i], rank, world, &cw[i], 0);
    for (int i = first; i < ncand; i++) screened += ratio * (double)cw[i].nTri;
  }
  const double tp1 = prof ? omp_get_wtime() : 0.0;
  EcpBatch *b = &bb->b;
  memset(b, 0, sizeof(*b));
  b->devEnum = 1;
  if (ncand == 0) {
    *centre = nat;
    return 0;
  }
  int ntake = 0;
  double est = 0.0;
  while (ntake < ncand) {
    est += ratio * (double)cw[ntake].nTri;
    ntake++;
    if (est >= (double)maxTriples) break;
  }
  if (ntake < ncand && C >= nat) {
    double rest = 0.0;
    for (int i = ntake; i < ncand; i++) rest += ratio * (double)cw[i].nTri;
    if (4.0 * rest <= (double)maxTriples) ntake = ncand;
  }
  *centre = (ntake < ncand) ? cand[ntake] : C;
  long long nAS = 0, nSS = 0, omTotal = 0, fRows = 0, nPairAS = 0, pairCand = 0;
  long long *asBase = malloc((ntake + 1) * sizeof(long long)), *ssBase = malloc((ntake + 1) * sizeof(long long));
  long long *omBase = malloc((ntake + 1) * sizeof(long long)), *fBase = malloc((ntake + 1) * sizeof(long long));
  ENSURE(bb->ceAS0, bb->capCE, ntake + 1, int);
  ENSURE(bb->cePair0, bb->capCE, ntake + 1, int64_t);
  if (ntake + 1 > bb->capCE) bb->capCE = ntake + 1 + 16;
  bb->nominal = 0;
  for (int i = 0; i < ntake; i++) {
    asBase[i] = nAS; ssBase[i] = nSS; omBase[i] = omTotal; fBase[i] = fRows;
    bb->ceAS0[i] = (int)nAS;
    bb->cePair0[i] = nPairAS;
    nAS += cw[i].nAS; nSS += cw[i].nSS; omTotal += cw[i].omSize; fRows += cw[i].fRows;
    nPairAS += (long long)cw[i].nAS * (cw[i].nAS + 1) / 2;
    pairCand += cw[i].nTri;
    bb->nominal += (long long)v->nrShells * (v->nrShells + 1) / 2;
  }
  bb->ceAS0[ntake] = (int)nAS;
  bb->cePair0[ntake] = nPairAS;
  bb->screenedShells = nSS;
  ENSURE(bb->asAtom, bb->capAS, nAS + 1, int); ENSURE(bb->asCentre, bb->capAS, nAS + 1, int);
  ENSURE(bb->asType, bb->capAS, nAS + 1, int);
  ENSURE(bb->asSS0, bb->capAS0, nAS + 1, int);
  if (nAS + 1 > bb->capAS0) bb->capAS0 = (int)nAS + 17;
  if (nAS + 1 > bb->capAS) bb->asR = (double *)realloc(bb->asR, (size_t)(nAS + 17) * 4 * sizeof(double));
  ENSURE(bb->asOmOff, bb->capAS, nAS + 1, int64_t);
  if (nAS + 1 > bb->capAS) bb->capAS = (int)nAS + 17;
  ENSURE(bb->ssShell, bb->capSS, nSS, int); ENSURE(bb->ssASlot, bb->capSS, nSS, int);
  ENSURE(bb->ssStart, bb->capSS, nSS, int); ENSURE(bb->ssEnd, bb->capSS, nSS, int);
  ENSURE(bb->ssFOff, bb->capSS, nSS, int64_t);
  ENSURE(bb->ssOwn, bb->capOwn, nSS, unsigned char);
  if (nSS > bb->capOwn) bb->capOwn = (int)nSS + 16;
  if (nSS > bb->capSS) bb->capSS = (int)nSS + 16;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int i = 0; i < ntake; i++) {
    const CentreWork *w = &cw[i];
    const double *rC = geometry + 3 * w->C;
    const int Lc = w->Lc;
    long long om = omBase[i], fr = fBase[i];
    for (int a = 0; a < w->nAS; a++) {
      const long long g = asBase[i] + a;
      const int X = w->asAtom[a];
      bb->asAtom[g] = X;
      bb->asCentre[g] = w->C;
      bb->asType[g] = w->type;
      bb->asR[4 * g + 0] = geometry[3 * X + 0] - rC[0];
      bb->asR[4 * g + 1] = geometry[3 * X + 1] - rC[1];
      bb->asR[4 * g + 2] = geometry[3 * X + 2] - rC[2];
      bb->asR[4 * g + 3] = w->asD[a];
      bb->asOmOff[g] = om;
      om += (long long)(Lc + t->atomMaxL[X]) * Lc * Lc * CD(t->atomMaxL[X]);
    }
    for (int s = 0; s < w->nSS; s++) {
      const long long g = ssBase[i] + s;
      bb->ssShell[g] = w->ssShell[s];
      bb->ssASlot[g] = (int)(asBase[i] + w->ssAtom[s]);
      bb->ssStart[g] = w->ssStart[s];
      bb->ssEnd[g] = w->ssEnd[s];
      bb->ssOwn[g] = w->ssOwn[s];
      bb->ssFOff[g] = fr;
      fr += Lc + v->shellL[w->ssShell[s]];
      if (s == 0 || w->ssAtom[s] != w->ssAtom[s - 1]) bb->asSS0[asBase[i] + w->ssAtom[s]] = (int)g; /* slots of an atom are contiguous */
    }
  }
  bb->asSS0[nAS] = (int)nSS;
  free(asBase); free(ssBase); free(omBase); free(fBase);
  if (prof)
    fprintf(stderr, "[builder] rank %d slots: centres %d/%d, %lld shell slots, threads %d: screen %.2f ms, layout %.2f ms\n", rank,
            ntake, ncand, nSS, nthreads, 1e3 * (tp1 - tp0), 1e3 * (omp_get_wtime() - tp1));
  for (int i = ntake; i < ncand; i++) {
    const CentreWork tmp = cw[i - ntake];
    cw[i - ntake] = cw[i];
    cw[i] = tmp;
    S->leftCentre[i - ntake] = cand[i];
  }
  S->nleft = ncand - ntake;
  S->leftNext = C;
  S->leftRank = rank;
  S->leftWorld = world;
  bb->nCanon = 0;
  b->nASlots = (int)nAS;
  b->asAtom = bb->asAtom;
  b->asCentre = bb->asCentre;
  b->asType = bb->asType;
  b->asR = bb->asR;
  b->asOmOff = bb->asOmOff;
  b->omTotal = omTotal;
  b->nSSlots = (int)nSS;
  b->ssShell = bb->ssShell;
  b->ssASlot = bb->ssASlot;
  b->ssStart = bb->ssStart;
  b->ssEnd = bb->ssEnd;
  b->ssFOff = bb->ssFOff;
  b->fRows = fRows;
  b->nCentres = ntake;
  b->ceAS0 = bb->ceAS0;
  b->cePair0 = bb->cePair0;
  b->asSS0 = bb->asSS0;
  b->ssOwn = bb->ssOwn;
  b->pairCand = pairCand;
  /* filled by the device layer (ecpdev_run_batch): the triple count, the totals and the class prefixes */
  b->clsFirst = bb->clsFirst;
  b->clsWork = bb->clsWork;
  b->clsElem = bb->clsElem;
  b->clsOutElem = bb->clsOutElem;
  b->clsPairBase = bb->clsPairBase;
  b->clsQBase = bb->clsQBase;
  return ntake;
}
