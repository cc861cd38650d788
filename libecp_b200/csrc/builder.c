/* builder.c - see builder.h */
#include "builder.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static int CD(int l) { return (l + 1) * (l + 2) * (l + 3) / 6; }
static int IJK(int l) { return (l + 1) * (l + 2) / 2; }

static void ensure_aslots(EcpBatchBuf *bb, int need) {
  if (need <= bb->capAS) return;
  const int cap = need * 3 / 2 + 1024;
  bb->asAtom = realloc(bb->asAtom, cap * sizeof(int));
  bb->asCentre = realloc(bb->asCentre, cap * sizeof(int));
  bb->asType = realloc(bb->asType, cap * sizeof(int));
  bb->asR = realloc(bb->asR, (size_t)cap * 4 * sizeof(double));
  bb->asOmOff = realloc(bb->asOmOff, cap * sizeof(int64_t));
  bb->capAS = cap;
}
static void ensure_sslots(EcpBatchBuf *bb, int need) {
  if (need <= bb->capSS) return;
  const int cap = need * 3 / 2 + 1024;
  bb->ssShell = realloc(bb->ssShell, cap * sizeof(int));
  bb->ssASlot = realloc(bb->ssASlot, cap * sizeof(int));
  bb->ssStart = realloc(bb->ssStart, cap * sizeof(int));
  bb->ssEnd = realloc(bb->ssEnd, cap * sizeof(int));
  bb->ssFOff = realloc(bb->ssFOff, cap * sizeof(int64_t));
  bb->capSS = cap;
}

EcpBatchBuf *ecp_batch_new(const EcpTables *t) {
  EcpBatchBuf *bb = calloc(1, sizeof(EcpBatchBuf));
  const int nc = t->v.nClasses;
  bb->clsFirst = calloc(nc + 2, sizeof(int));
  bb->clsWork = calloc(nc + 2, sizeof(int64_t));
  bb->clsElem = calloc(nc + 2, sizeof(int64_t));
  bb->clsOutElem = calloc(nc + 2, sizeof(int64_t));
  bb->scratchSlot = malloc((t->v.nrShells + 1) * sizeof(int));
  bb->scratchList = malloc((t->v.nrShells + 1) * sizeof(int));
  return bb;
}

void ecp_batch_free(EcpBatchBuf *bb) {
  if (!bb) return;
  free(bb->asAtom); free(bb->asCentre); free(bb->asType); free(bb->asR); free(bb->asOmOff);
  free(bb->ssShell); free(bb->ssASlot); free(bb->ssStart); free(bb->ssEnd); free(bb->ssFOff);
  free(bb->trA); free(bb->trB); free(bb->trClass); free(bb->trOut); free(bb->trT); free(bb->trG); free(bb->trPair);
  free(bb->prTriple); free(bb->prQOff); free(bb->prRshOff);
  free(bb->clsFirst); free(bb->clsWork); free(bb->clsElem); free(bb->clsOutElem);
  free(bb->cnA); free(bb->cnS1); free(bb->cnB); free(bb->cnS2); free(bb->cnC); free(bb->cnLa); free(bb->cnLb);
  free(bb->cnOut);
  free(bb->scratchSlot); free(bb->scratchList);
  free(bb);
}

/* basis-set screening of one shell against the (potential-capped) small grid: first grid point with
 * r >= d-R and last with r <= d+R, searched downwards from the cap exactly as the reference does
 * (src/type2.c:148-180).  The KK-mapped abscissae increase strictly, so the two downward scans are
 * binary searches; the linear form is kept for the first/last few points where rounding could matter. */
void ecp_shell_window(const EcpTables *t, int endLast, double radius, double dist, int *start, int *end, int *skip) {
  const double rmin = dist - radius, rmax = dist + radius;
  const double *r = t->small_x;
  int j;
  /* largest j <= endLast with r[j] < rmin  (or -1) */
  {
    int lo = -1, hi = endLast; /* invariant: r[lo] < rmin (or lo == -1), answer in [lo, hi] */
    while (lo < hi) {
      const int mid = (lo + hi + 1) / 2;
      if (r[mid] < rmin)
        lo = mid;
      else
        hi = mid - 1;
    }
    *start = lo + 1;
  }
  /* largest j <= endLast with r[j] <= rmax (or -1) */
  {
    int lo = -1, hi = endLast;
    while (lo < hi) {
      const int mid = (lo + hi + 1) / 2;
      if (r[mid] <= rmax)
        lo = mid;
      else
        hi = mid - 1;
    }
    j = lo;
  }
  *end = j;
  *skip = !(*end >= *start);
}

int ecp_pair_owner(int a, int b, int world) {
  if (world <= 1) return 0;
  uint64_t h = (uint64_t)(uint32_t)a * 0x9E3779B97F4A7C15ull ^ ((uint64_t)(uint32_t)b + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (int)(h % (uint64_t)world);
}

static double dist3(const double *a, const double *b) { /* src/util.c:109-116 */
  const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
  return sqrt(x * x + y * y + z * z);
}

int ecp_batch_build(const EcpTables *t, const double *geometry, int *centre, long long maxTriples, int rank, int world,
                    int keepCanon, EcpBatchBuf *bb) {
  const EcpHostTables *v = &t->v;
  const int nat = v->nrAtoms, nc = v->nClasses;
  int nAS = 0, nSS = 0, nTR = 0, consumed = 0;
  int64_t omTotal = 0, fRows = 0, outTotal = 0;
  bb->nCanon = 0;
  bb->nominal = 0;
  bb->screenedShells = 0;
  /* pass 1: per centre, slots and triples in canonical order (unsorted triple arrays use the tr* buffers
   * temporarily, then a counting sort by class reorders them) */
  int *tmpA = NULL, *tmpB = NULL, *tmpC = NULL;
  int64_t *tmpOut = NULL;
  int capTmp = 0;
  int C = *centre;
  for (; C < nat; C++) {
    const int type = t->atomType[C];
    if (type < 0) continue;
    if (consumed > 0 && nTR >= maxTriples) break;
    consumed++;
    const EcpType *T = &t->types[type];
    const int Lc = T->L, endLast = T->endLast;
    const double *rC = geometry + 3 * C;
    const double rcap = (endLast >= 0) ? t->small_x[endLast] : -1.0;
    bb->nominal += (long long)v->nrShells * (v->nrShells + 1) / 2;
    /* screening: shell slots and atom slots of this centre */
    const int ss0 = nSS;
    for (int X = 0; X < nat; X++) {
      const double d = dist3(rC, geometry + 3 * X);
      const int s0 = t->atomFirstShell[X], s1 = t->atomFirstShell[X + 1];
      int aslot = -1;
      /* atom-level prune: every shell of X lies beyond the potential cut-off */
      double Rmax = 0.0;
      for (int s = s0; s < s1; s++)
        if (t->shellRadius[s] > Rmax) Rmax = t->shellRadius[s];
      if (endLast < 0 || d - Rmax > rcap) {
        for (int s = s0; s < s1; s++) bb->scratchSlot[s] = -1;
        continue;
      }
      for (int s = s0; s < s1; s++) {
        int st, en, sk;
        ecp_shell_window(t, endLast, t->shellRadius[s], d, &st, &en, &sk);
        bb->scratchSlot[s] = -1;
        if (sk) continue;
        if (aslot < 0) {
          ensure_aslots(bb, nAS + 1);
          aslot = nAS++;
          bb->asAtom[aslot] = X;
          bb->asCentre[aslot] = C;
          bb->asType[aslot] = type;
          /* r_XC = X - C (reference distanceVector(rAC, rC, rA), src/libecp.c:281) */
          bb->asR[4 * aslot + 0] = geometry[3 * X + 0] - rC[0];
          bb->asR[4 * aslot + 1] = geometry[3 * X + 1] - rC[1];
          bb->asR[4 * aslot + 2] = geometry[3 * X + 2] - rC[2];
          bb->asR[4 * aslot + 3] = d;
          bb->asOmOff[aslot] = omTotal;
          omTotal += (int64_t)(Lc + t->atomMaxL[X]) * Lc * Lc * CD(t->atomMaxL[X]);
        }
        ensure_sslots(bb, nSS + 1);
        bb->ssShell[nSS] = s;
        bb->ssASlot[nSS] = aslot;
        bb->ssStart[nSS] = st;
        bb->ssEnd[nSS] = en;
        bb->ssFOff[nSS] = fRows;
        fRows += Lc + t->shellL[s];
        bb->scratchSlot[s] = nSS++;
      }
    }
    bb->screenedShells += nSS - ss0;
    /* canonical enumeration: A, B>=A, s1, s2 (reference src/libecp.c:278-320); slots of one atom are contiguous */
    for (int ia = ss0; ia < nSS;) {
      const int A = v->shellAtom[bb->ssShell[ia]];
      int ia1 = ia;
      while (ia1 < nSS && v->shellAtom[bb->ssShell[ia1]] == A) ia1++;
      for (int ib = ia; ib < nSS;) {
        const int B = v->shellAtom[bb->ssShell[ib]];
        int ib1 = ib;
        while (ib1 < nSS && v->shellAtom[bb->ssShell[ib1]] == B) ib1++;
        for (int a = ia; a < ia1; a++)
          for (int b = (A == B ? a : ib); b < ib1; b++) {
            const int gs = bb->ssStart[a] > bb->ssStart[b] ? bb->ssStart[a] : bb->ssStart[b];
            const int ge = bb->ssEnd[a] > bb->ssEnd[b] ? bb->ssEnd[a] : bb->ssEnd[b];
            if (!(gs < ge)) continue; /* src/libecp.c:344, identical for both types */
            const int sa = bb->ssShell[a], sb = bb->ssShell[b];
            if (world > 1 && ecp_pair_owner(sa, sb, world) != rank) continue;
            if (nTR + 1 > capTmp) {
              capTmp = (nTR + 1) * 3 / 2 + 4096;
              tmpA = realloc(tmpA, capTmp * sizeof(int));
              tmpB = realloc(tmpB, capTmp * sizeof(int));
              tmpC = realloc(tmpC, capTmp * sizeof(int));
              tmpOut = realloc(tmpOut, capTmp * sizeof(int64_t));
            }
            const int la = v->shellL[sa], lb = v->shellL[sb];
            tmpA[nTR] = a;
            tmpB[nTR] = b;
            tmpC[nTR] = t->clsLookup[la][lb][Lc];
            tmpOut[nTR] = outTotal;
            if (keepCanon) {
              const int k = bb->nCanon;
              if (k + 1 > bb->capCanon) {
                bb->capCanon = (k + 1) * 3 / 2 + 4096;
                bb->cnA = realloc(bb->cnA, bb->capCanon * sizeof(int));
                bb->cnS1 = realloc(bb->cnS1, bb->capCanon * sizeof(int));
                bb->cnB = realloc(bb->cnB, bb->capCanon * sizeof(int));
                bb->cnS2 = realloc(bb->cnS2, bb->capCanon * sizeof(int));
                bb->cnC = realloc(bb->cnC, bb->capCanon * sizeof(int));
                bb->cnLa = realloc(bb->cnLa, bb->capCanon * sizeof(int));
                bb->cnLb = realloc(bb->cnLb, bb->capCanon * sizeof(int));
                bb->cnOut = realloc(bb->cnOut, bb->capCanon * sizeof(int64_t));
              }
              bb->cnA[k] = A;
              bb->cnS1[k] = sa - t->atomFirstShell[A];
              bb->cnB[k] = B;
              bb->cnS2[k] = sb - t->atomFirstShell[B];
              bb->cnC[k] = C;
              bb->cnLa[k] = la;
              bb->cnLb[k] = lb;
              bb->cnOut[k] = outTotal;
              bb->nCanon = k + 1;
            }
            outTotal += 2 * (int64_t)IJK(la) * IJK(lb);
            nTR++;
          }
        ib = ib1;
      }
      ia = ia1;
    }
  }
  *centre = C;

  /* pass 2: counting sort of the triples by class, then per-triple offsets */
  if (nTR + 1 > bb->capTR) {
    bb->capTR = (nTR + 1) * 3 / 2 + 4096;
    bb->trA = realloc(bb->trA, bb->capTR * sizeof(int));
    bb->trB = realloc(bb->trB, bb->capTR * sizeof(int));
    bb->trClass = realloc(bb->trClass, bb->capTR * sizeof(int));
    bb->trOut = realloc(bb->trOut, bb->capTR * sizeof(int64_t));
    bb->trT = realloc(bb->trT, bb->capTR * sizeof(int64_t));
    bb->trG = realloc(bb->trG, bb->capTR * sizeof(int64_t));
    bb->trPair = realloc(bb->trPair, bb->capTR * sizeof(int64_t));
  }
  int *fill = calloc(nc + 2, sizeof(int));
  memset(bb->clsFirst, 0, (nc + 2) * sizeof(int));
  for (int i = 0; i < nTR; i++) bb->clsFirst[tmpC[i] + 1]++;
  for (int c = 0; c < nc; c++) bb->clsFirst[c + 1] += bb->clsFirst[c];
  for (int i = 0; i < nTR; i++) {
    const int c = tmpC[i], p = bb->clsFirst[c] + fill[c]++;
    bb->trA[p] = tmpA[i];
    bb->trB[p] = tmpB[i];
    bb->trClass[p] = c;
    bb->trOut[p] = tmpOut[i];
  }
  free(fill);
  free(tmpA);
  free(tmpB);
  free(tmpC);
  free(tmpOut);
  int64_t tTot = 0, gTot = 0, nPairs = 0, qTot = 0, rshTot = 0;
  bb->clsWork[0] = bb->clsElem[0] = bb->clsOutElem[0] = 0;
  for (int c = 0; c < nc; c++) {
    const int la = v->clsLa[c], lb = v->clsLb[c], lab = la + lb;
    const int64_t n = bb->clsFirst[c + 1] - bb->clsFirst[c];
    bb->clsWork[c + 1] = bb->clsWork[c] + n * v->clsNq[c];
    bb->clsElem[c + 1] = bb->clsElem[c] + n * CD(la) * CD(lb);
    bb->clsOutElem[c + 1] = bb->clsOutElem[c] + n * IJK(la) * IJK(lb);
    for (int i = bb->clsFirst[c]; i < bb->clsFirst[c + 1]; i++) {
      const int sa = bb->ssShell[bb->trA[i]], sb = bb->ssShell[bb->trB[i]];
      const int np = v->shellK[sa] * v->shellK[sb];
      bb->trT[i] = tTot;
      bb->trG[i] = gTot;
      bb->trPair[i] = nPairs;
      tTot += v->clsNq[c];
      gTot += CD(la) * CD(lb);
      if (nPairs + np > bb->capPR) {
        bb->capPR = (int)((nPairs + np) * 3 / 2 + 4096);
        bb->prTriple = realloc(bb->prTriple, bb->capPR * sizeof(int));
        bb->prQOff = realloc(bb->prQOff, bb->capPR * sizeof(int64_t));
        bb->prRshOff = realloc(bb->prRshOff, bb->capPR * sizeof(int64_t));
      }
      for (int p = 0; p < np; p++) {
        bb->prTriple[nPairs] = i;
        bb->prQOff[nPairs] = qTot;
        bb->prRshOff[nPairs] = rshTot;
        qTot += (lab + 1) * (lab + 1);
        rshTot += (lab + 1) * (lab + 1);
        nPairs++;
      }
    }
  }
  EcpBatch *b = &bb->b;
  b->nASlots = nAS;
  b->asAtom = bb->asAtom;
  b->asCentre = bb->asCentre;
  b->asType = bb->asType;
  b->asR = bb->asR;
  b->asOmOff = bb->asOmOff;
  b->omTotal = omTotal;
  b->nSSlots = nSS;
  b->ssShell = bb->ssShell;
  b->ssASlot = bb->ssASlot;
  b->ssStart = bb->ssStart;
  b->ssEnd = bb->ssEnd;
  b->ssFOff = bb->ssFOff;
  b->fRows = fRows;
  b->nTriples = nTR;
  b->trA = bb->trA;
  b->trB = bb->trB;
  b->trClass = bb->trClass;
  b->trOut = bb->trOut;
  b->trT = bb->trT;
  b->trG = bb->trG;
  b->trPair = bb->trPair;
  b->tTotal = tTot;
  b->gTotal = gTot;
  b->outTotal = outTotal;
  b->nPairs = nPairs;
  b->qTotal = qTot;
  b->rshTotal = rshTot;
  b->prTriple = bb->prTriple;
  b->prQOff = bb->prQOff;
  b->prRshOff = bb->prRshOff;
  b->clsFirst = bb->clsFirst;
  b->clsWork = bb->clsWork;
  b->clsElem = bb->clsElem;
  b->clsOutElem = bb->clsOutElem;
  return consumed;
}
