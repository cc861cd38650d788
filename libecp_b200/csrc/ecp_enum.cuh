/* ecp_enum.cuh - triple enumeration on the device (matrix runs; LIBECP_B200_ENUM=host keeps the host builder).
 *
 * Replaces the O(triples) part of the loop nest of calculateECPIntegrals (reference src/libecp.c:278-344): the pair loops
 * over the shells of two atoms with the window test `max(start) < max(end)` (:315-316,344).  The host keeps what decides
 * integers in double precision - the screening windows of every (centre, shell) (builder.c, bit-identical to the
 * reference) - and uploads, per centre, its atom slots and shell slots.  Measured reason (profiles/r2): at 8 GPUs a rank
 * owns 1/8 of the triples but its builder ran 9-12 ms per pass on the rank's four host cores against 13 ms of kernels, and
 * the ramp of batch sizes made the GPU wait for the builder (20.6 ms per step, efficiency 0.51).
 *
 *   k_enum_count : block per centre, thread per atom-slot pair (ka <= kb): executed triples and primitive pairs of the
 *                  pair per (la, lb), and the centre's totals per class
 *   k_enum_scan  : one block: prefix over the centres per class, then over the classes: clsFirst / clsWork / clsElem /
 *                  clsOutElem / clsPairBase / clsQBase and the totals the host sizes the batch with
 *   k_enum_fill  : block per centre: prefix over the atom-slot pairs per (la, lb) (a warp per class), then every pair
 *                  writes its triples (trA, trB, trPair) to their places
 * Order inside a class: centre, atom pair (ka, kb), a, b - the order of the host builder and of the reference's loop
 * nest, so that consecutive triples share both atoms (k_link4 stages the angular tables per such run).
 */
#ifndef ECP_ENUM_CUH
#define ECP_ENUM_CUH

struct EnumIn {
  int nCentres, nc, lb1; /* lb1 = maxLBS + 1; (la, lb) index k = la * lb1 + lb */
  const int *ceAS0;
  const long long *cePair0;
  const int *asSS0, *asType;
  const int *ssShell, *ssStart, *ssEnd;
  const unsigned char *ssOwn;
  int *ccTri, *ccPair; /* [class][centre]: counts, then (k_enum_scan) exclusive prefixes over the centres */
  int2 *pairCnt;       /* per centre [k][atom pair]: (triples, primitive pairs), then (k_enum_fill) first positions */
};
struct EnumMeta {
  long long nTriples, nPairs, tTotal, gTotal, qTotal, outTotal;
};

/* shell slots of one centre staged in shared memory: l | K << 8 | own << 16, and the window (start | end << 16).  The pair
 * loops below read every slot of an atom pair (13 x 13 on the TZ sets): from global memory each test was a chain of five
 * dependent loads (0.4 ms per kernel and batch, measured as idle time between batches). */
__device__ __forceinline__ void enum_stage(const DevT &t, const EnumIn &in, int ss0, int nS, unsigned *sInfo, unsigned *sWin) {
  for (int s = threadIdx.x; s < nS; s += blockDim.x) {
    const int sh = in.ssShell[ss0 + s];
    sInfo[s] = (unsigned)t.shellL[sh] | ((unsigned)t.shellK[sh] << 8) | ((unsigned)in.ssOwn[ss0 + s] << 16);
    sWin[s] = (unsigned)in.ssStart[ss0 + s] | ((unsigned)in.ssEnd[ss0 + s] << 16);
  }
}
/* the window test of src/libecp.c:315-316,344 on two staged windows (start, end < 2^15) */
__device__ __forceinline__ bool enum_ok(unsigned wa, unsigned wb) {
  const unsigned gs = max(wa & 0xffffu, wb & 0xffffu), ge = max(wa >> 16, wb >> 16);
  return gs < ge;
}

__device__ __forceinline__ void enum_pair_of(int pidx, int nA, int *ka, int *kb) {
  int a = 0, rowLen = nA;
  while (pidx >= rowLen) {
    pidx -= rowLen;
    rowLen--;
    a++;
  }
  *ka = a;
  *kb = a + pidx;
}

__global__ void __launch_bounds__(256) k_enum_count(DevT t, EnumIn in) {
  __shared__ int clsTri[ECP_MAX_CLASSES], clsPair[ECP_MAX_CLASSES];
  const int i = blockIdx.x;
  const int as0 = in.ceAS0[i], nA = in.ceAS0[i + 1] - as0;
  const int nP = nA * (nA + 1) / 2, nK = in.lb1 * in.lb1;
  const int Lc = nA > 0 ? t.typeL[in.asType[as0]] : 0;
  int2 *cnt = in.pairCnt + in.cePair0[i] * nK;
  extern __shared__ unsigned enum_sm[];
  const int ss0 = nA > 0 ? in.asSS0[as0] : 0, nS = nA > 0 ? in.asSS0[as0 + nA] - ss0 : 0;
  unsigned *sInfo = enum_sm, *sWin = enum_sm + nS;
  enum_stage(t, in, ss0, nS, sInfo, sWin);
  for (int c = threadIdx.x; c < in.nc; c += blockDim.x) clsTri[c] = clsPair[c] = 0;
  __syncthreads();
  for (int pidx = threadIdx.x; pidx < nP; pidx += blockDim.x) {
    int ka, kb;
    enum_pair_of(pidx, nA, &ka, &kb);
    int nt[(ECP_MAX_LBS + 1) * (ECP_MAX_LBS + 1)], np[(ECP_MAX_LBS + 1) * (ECP_MAX_LBS + 1)];
    for (int k = 0; k < nK; k++) nt[k] = np[k] = 0;
    const int a0 = in.asSS0[as0 + ka] - ss0, a1 = in.asSS0[as0 + ka + 1] - ss0;
    const int b0 = in.asSS0[as0 + kb] - ss0, b1 = in.asSS0[as0 + kb + 1] - ss0;
    for (int a = a0; a < a1; a++) {
      const unsigned ia = sInfo[a], wa = sWin[a];
      if (!(ia >> 16)) continue; /* row ownership of the multi-GPU partition */
      const int la = ia & 255u, Ka = (ia >> 8) & 255u;
      for (int b = (ka == kb ? a : b0); b < b1; b++) { /* A == B: s2 >= s1 (src/libecp.c:304) */
        if (!enum_ok(wa, sWin[b])) continue;          /* src/libecp.c:315-316,344 */
        const unsigned ib = sInfo[b];
        const int k = la * in.lb1 + (int)(ib & 255u);
        nt[k]++;
        np[k] += Ka * (int)((ib >> 8) & 255u);
      }
    }
    for (int k = 0; k < nK; k++) {
      cnt[(size_t)k * nP + pidx] = make_int2(nt[k], np[k]);
      if (nt[k]) {
        const int c = t.clsLookup[(k / in.lb1 * (ECP_MAX_LBS + 1) + k % in.lb1) * (ECP_MAX_LECP + 1) + Lc];
        atomicAdd(&clsTri[c], nt[k]);
        atomicAdd(&clsPair[c], np[k]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < in.nc; c += blockDim.x) {
    in.ccTri[(size_t)c * in.nCentres + i] = clsTri[c];
    in.ccPair[(size_t)c * in.nCentres + i] = clsPair[c];
  }
}

__global__ void __launch_bounds__(128) k_enum_scan(DevT t, EnumIn in, int *clsFirst, long long *clsWork, long long *clsElem,
                                                  long long *clsOutElem, long long *clsPairBase, long long *clsQBase,
                                                  EnumMeta *meta) {
  __shared__ int nTri[ECP_MAX_CLASSES], nPair[ECP_MAX_CLASSES];
  for (int c = threadIdx.x; c < in.nc; c += blockDim.x) {
    int run = 0, runP = 0;
    int *ct = in.ccTri + (size_t)c * in.nCentres, *cp = in.ccPair + (size_t)c * in.nCentres;
    for (int i = 0; i < in.nCentres; i++) {
      const int a = ct[i], b = cp[i];
      ct[i] = run;
      cp[i] = runP;
      run += a;
      runP += b;
    }
    nTri[c] = run;
    nPair[c] = runP;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long first = 0, work = 0, elem = 0, oelem = 0, pairs = 0, q = 0;
    for (int c = 0; c < in.nc; c++) {
      const int la = t.clsLa[c], lb = t.clsLb[c], lab = la + lb;
      clsFirst[c] = (int)first;
      clsWork[c] = work;
      clsElem[c] = elem;
      clsOutElem[c] = oelem;
      clsPairBase[c] = pairs;
      clsQBase[c] = q;
      const long long n = nTri[c], np = nPair[c];
      first += n;
      work += n * t.clsNq[c];
      elem += n * ecp_cd(la) * ecp_cd(lb);
      oelem += n * ecp_ijk(la) * ecp_ijk(lb);
      pairs += np;
      q += np * (lab + 1) * (lab + 1);
    }
    clsFirst[in.nc] = (int)first;
    clsWork[in.nc] = work;
    clsElem[in.nc] = elem;
    clsOutElem[in.nc] = oelem;
    clsPairBase[in.nc] = pairs;
    clsQBase[in.nc] = q;
    meta->nTriples = first;
    meta->nPairs = pairs;
    meta->tTotal = work;
    meta->gTotal = elem;
    meta->qTotal = q;
    meta->outTotal = 2 * oelem;
  }
}

__global__ void __launch_bounds__(256) k_enum_fill(DevT t, EnumIn in, const int *clsFirst, const long long *clsPairBase,
                                                  int *trA, int *trB, long long *trPair) {
  const int i = blockIdx.x;
  const int as0 = in.ceAS0[i], nA = in.ceAS0[i + 1] - as0;
  const int nP = nA * (nA + 1) / 2, nK = in.lb1 * in.lb1;
  const int Lc = nA > 0 ? t.typeL[in.asType[as0]] : 0;
  int2 *cnt = in.pairCnt + in.cePair0[i] * nK;
  extern __shared__ unsigned enum_sm[];
  const int ss0 = nA > 0 ? in.asSS0[as0] : 0, nS = nA > 0 ? in.asSS0[as0 + nA] - ss0 : 0;
  unsigned *sInfo = enum_sm, *sWin = enum_sm + nS;
  enum_stage(t, in, ss0, nS, sInfo, sWin);
  /* ---- first places per (la, lb) and atom pair: a warp per k, exclusive prefix over the pairs in chunks of 32 ---- */
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = warp; k < nK; k += nw) {
    const int c = t.clsLookup[(k / in.lb1 * (ECP_MAX_LBS + 1) + k % in.lb1) * (ECP_MAX_LECP + 1) + Lc];
    if (c < 0) continue; /* no such class: all counts are zero */
    int run = clsFirst[c] + in.ccTri[(size_t)c * in.nCentres + i];
    int runP = (int)clsPairBase[c] + in.ccPair[(size_t)c * in.nCentres + i];
    int2 *row = cnt + (size_t)k * nP;
    for (int p0 = 0; p0 < nP; p0 += 32) {
      const int p = p0 + lane;
      const int2 v = p < nP ? row[p] : make_int2(0, 0);
      int x = v.x, y = v.y;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ux = __shfl_up_sync(0xffffffffu, x, o), uy = __shfl_up_sync(0xffffffffu, y, o);
        if (lane >= o) {
          x += ux;
          y += uy;
        }
      }
      if (p < nP) row[p] = make_int2(run + x - v.x, runP + y - v.y);
      run += __shfl_sync(0xffffffffu, x, 31);
      runP += __shfl_sync(0xffffffffu, y, 31);
    }
  }
  __syncthreads();
  /* ---- every atom pair writes its triples ---- */
  for (int pidx = threadIdx.x; pidx < nP; pidx += blockDim.x) {
    int ka, kb;
    enum_pair_of(pidx, nA, &ka, &kb);
    const int a0 = in.asSS0[as0 + ka] - ss0, a1 = in.asSS0[as0 + ka + 1] - ss0;
    const int b0 = in.asSS0[as0 + kb] - ss0, b1 = in.asSS0[as0 + kb + 1] - ss0;
    /* first places of the pair per (la, lb): thread-local (a read-modify-write of the global entry per triple was a
     * chain of L2 round trips - 0.5 ms per batch) */
    int px[(ECP_MAX_LBS + 1) * (ECP_MAX_LBS + 1)], py[(ECP_MAX_LBS + 1) * (ECP_MAX_LBS + 1)];
    for (int k = 0; k < nK; k++) {
      const int2 v = cnt[(size_t)k * nP + pidx];
      px[k] = v.x;
      py[k] = v.y;
    }
    for (int a = a0; a < a1; a++) {
      const unsigned ia = sInfo[a], wa = sWin[a];
      if (!(ia >> 16)) continue;
      const int la = ia & 255u, Ka = (ia >> 8) & 255u;
      for (int b = (ka == kb ? a : b0); b < b1; b++) {
        if (!enum_ok(wa, sWin[b])) continue;
        const unsigned ib = sInfo[b];
        const int k = la * in.lb1 + (int)(ib & 255u);
        const int p = px[k]++;
        trA[p] = ss0 + a;
        trB[p] = ss0 + b;
        trPair[p] = py[k];
        py[k] += Ka * (int)((ib >> 8) & 255u);
      }
    }
  }
}

#endif
