/* ecp_waves.cuh - type-2 large-grid fallback as LEVEL WAVES (default; LIBECP_B200_FB=group keeps k_fallbackG).
 *
 * Replaces calcT_FM06 (reference src/type2.c:417-528) + QIntegrand_FM06 (:397-410) + integrateGC_PSM92
 * (src/gc_integrators.c:38-86), like ecp_fallback.cuh.
 *
 * Why: the persistent kernels (k_fallbackG, 8 lanes per item; a one-warp-per-item rewrite, measured in round 2) are
 * latency bound - a chunk is a long serial chain (abscissa -> exponent -> potential -> two Bessel evaluations -> tile ->
 * owner sums -> vote) and at 128-168 registers only 12-16 warps per SM are there to hide it; ncu: 11-17 of 32 lanes
 * active, 23 % of the instructions FP64, 45 % of the issue slots used.  The adaptive rule only needs, per quadrature,
 * the SUM over the new points of a level before it can decide; which points a level has is known up front.  So the
 * work is cut the other way round:
 *   unit        = (item, primitive pair): one PSM92 quadrature family on its own FM06-mapped grid; the reference
 *                 converges every pair on its own and adds the pair integrals (T += grid->I, src/type2.c:513)
 *   k_fbw_eval  = one thread per (open unit, grid point of the wave): potential, exponential, the two Bessel vectors (code
 *                 per order, the warp is one unit: uniform switch), and the integrand of each failed quadrature of the
 *                 unit written to vals[unit][quadrature][point].  No state, no votes, no waiting lanes but the
 *                 points below the exponent gate; thousands of independent warps.
 *   k_fbw_book  = the PSM92 bookkeeping (src/gc_integrators.c:49-83): per (open unit, quadrature) add the wave's values
 *                 pair by pair in slot order (the reference's association: T = left + right, I += T), test, keep the
 *                 state; units with an open quadrature go to the list of the next wave.
 *   waves       : levels 0..4 (slots 0..31 of the level-major grid), then one level per wave (32, 64, ... 512 points).
 *                 The host reads three counters back per wave to size the next launch.
 *   k_fbw_final = T(item, quadrature) = sum over the pairs in the reference's order.
 * Integrand values are those of k_fallbackG (same factors, same association).
 */
#ifndef ECP_WAVES_CUH
#define ECP_WAVES_CUH

struct __align__(16) FbwItem {
  long long sBase; /* first state of the item: states [sBase + ip * nf + j], ip = pair, j = failed quadrature */
  long long tOff;  /* T of the triple */
  int uRank, qBase, nf, npair; /* uRank: first unit of the item inside its bucket */
  int bucket, pad0, pad1, pad2;
};
/* units are laid out bucket by bucket, a bucket = the Bessel orders (la + l, lb + l) of the item: the warps that run side
 * by side on an SM then execute the same instantiation of the Bessel code (ncu on the unsorted form: 5.8 warps per issue
 * stalled on instruction fetch) */
#define FBW_NBUCKET 128
#define FBW_CTR 8 /* counters before the per-bucket unit counts */
struct __align__(16) FbwUnit {
  double dAC, dBC, zA, zB, Cc, i1, i2;
  long long sBase; /* first state of the unit */
  int qBase, nq, g0, g1;
  int lpack, pad; /* laC | lbC << 4 | lab << 8 | l << 12 */
};
struct FbwOpen {
  long long valBase;
  int unit, pad;
};
struct FbwQ { /* one failed quadrature of an item */
  int lll; /* l1 | l2 << 4 | l3 << 8 */
  int k;   /* position in the class list = index into the triple's T */
};
/* counters: [0] units [1] states [2] quadrature descriptors [3] open units of the next wave [4] value space of the next
 * wave; [FBW_CTR + bucket] units per bucket */

/* warp-aggregated reservation: every lane asks for n (may be 0) entries of counter c; returns the lane's first entry */
__device__ __forceinline__ long long fbw_reserve(unsigned long long *ctr, long long n) {
  const int lane = threadIdx.x & 31;
  long long incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const long long tot = __shfl_sync(0xffffffffu, incl, 31);
  long long base = 0;
  if (lane == 31 && tot > 0) base = (long long)atomicAdd(ctr, (unsigned long long)tot);
  base = __shfl_sync(0xffffffffu, base, 31);
  return base + incl - n;
}

/* ---- per item: number of failed quadratures, primitive pairs; reserve units / states / descriptors ---- */
__global__ void k_fbw_count(DevT t, DevB b, FbwItem *items, unsigned long long *ctr) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  const int nItems = b.counters[0];
  int nf = 0, npair = 0, bucket = -1;
  long long tOff = 0;
  if (it < nItems) {
    const int item = b.items[it], tri = item >> 3, l = item & 7;
    const int cl = find_class_i(b.clsFirst, t.nClasses, tri);
    bucket = (t.clsLa[cl] + l) * 11 + (t.clsLb[cl] + l); /* orders <= ECP_KMAX = 10 */
    const int k0 = t.clsQlOff[cl * (ECP_MAX_LECP + 1) + l], k1 = t.clsQlOff[cl * (ECP_MAX_LECP + 1) + l + 1];
    tOff = tri_T_off(t, b, cl, tri);
    for (int k = k0; k < k1; k++) nf += b.tfail[tOff + k] ? 1 : 0;
    npair = t.shellK[b.ssShell[b.trA[tri]]] * t.shellK[b.ssShell[b.trB[tri]]];
  }
  /* rank of the item's units inside its bucket: one atomic per distinct bucket of the warp */
  long long u = 0;
  {
    const unsigned peers = __match_any_sync(0xffffffffu, bucket);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    /* exclusive prefix of npair over the peers, in lane order */
    int before = 0, total = 0;
    for (unsigned m = peers; m; m &= m - 1) {
      const int src = __ffs(m) - 1;
      const int v = __shfl_sync(peers, npair, src);
      if (src < lane) before += v;
      total += v;
    }
    unsigned long long base = 0;
    if (lane == leader && bucket >= 0 && total > 0) base = atomicAdd(ctr + FBW_CTR + bucket, (unsigned long long)total);
    base = __shfl_sync(peers, base, leader);
    u = (long long)base + before;
  }
  fbw_reserve(ctr + 0, npair);
  const long long s = fbw_reserve(ctr + 1, (long long)npair * nf);
  const long long q = fbw_reserve(ctr + 2, nf);
  if (it < nItems) {
    FbwItem r;
    r.sBase = s;
    r.tOff = tOff;
    r.uRank = (int)u;
    r.qBase = (int)q;
    r.nf = nf;
    r.npair = npair;
    r.bucket = bucket;
    r.pad0 = r.pad1 = r.pad2 = 0;
    items[it] = r;
  }
}

/* ---- per item: quadrature descriptors and one record per primitive pair (src/type2.c:452-468) ---- */
__global__ void k_fbw_units(DevT t, DevB b, const FbwItem *items, const int *bucketOff, FbwUnit *units, FbwQ *qd) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= b.counters[0]) return;
  const FbwItem I = items[it];
  const int item = b.items[it], tri = item >> 3, l = item & 7;
  const int cl = find_class_i(b.clsFirst, t.nClasses, tri);
  const int la = t.clsLa[cl], lb = t.clsLb[cl];
  const int k0 = t.clsQlOff[cl * (ECP_MAX_LECP + 1) + l], k1 = t.clsQlOff[cl * (ECP_MAX_LECP + 1) + l + 1];
  const int *ql = t.qlist + t.clsQOff[cl];
  {
    int j = 0;
    for (int k = k0; k < k1; k++)
      if (b.tfail[I.tOff + k]) {
        const int qq = ql[k];
        FbwQ d;
        d.lll = ((qq >> 4) & 15) | (((qq >> 8) & 15) << 4) | (((qq >> 12) & 15) << 8);
        d.k = k;
        qd[I.qBase + j++] = d;
      }
  }
  const int ssa = b.trA[tri], ssb = b.trB[tri];
  const int sha = b.ssShell[ssa], shb = b.ssShell[ssb];
  const int asa = b.ssASlot[ssa], asb = b.ssASlot[ssb];
  const double dAC = b.asR[4 * asa + 3], dBC = b.asR[4 * asb + 3];
  const int type = b.asType[asa];
  int g0 = t.typeGaussOff[type], g1 = t.typeGaussOff[type + 1];
  /* Gaussians of channel l form one run of the type's list (src/ecp.c:47-57 tests and skips the others) */
  while (g0 < g1 && t.gaussL[g0] != l) g0++;
  while (g1 > g0 && t.gaussL[g1 - 1] != l) g1--;
  const int Na = t.shellK[sha], Nb = t.shellK[shb];
  const double *za = t.primA + t.shellPrim[sha], *ca = t.primD + t.shellPrim[sha];
  const double *zb = t.primA + t.shellPrim[shb], *cb = t.primD + t.shellPrim[shb];
  for (int pa = 0, ip = 0; pa < Na; pa++)
    for (int pb = 0; pb < Nb; pb++, ip++) {
      FbwUnit u;
      u.dAC = dAC;
      u.dBC = dBC;
      u.zA = za[pa];
      u.zB = zb[pb];
      u.Cc = ca[pa] * cb[pb];
      const double zp = u.zA + u.zB;
      ecp_fm06_map(zp, (u.zA * dAC + u.zB * dBC) / zp, &u.i1, &u.i2);
      u.sBase = I.sBase + (long long)ip * I.nf;
      u.qBase = I.qBase;
      u.nq = I.nf;
      u.g0 = g0;
      u.g1 = g1;
      u.lpack = (la + l) | ((lb + l) << 4) | ((la + lb) << 8) | (l << 12);
      u.pad = 0;
      units[bucketOff[I.bucket] + I.uRank + ip] = u;
    }
}

/* Bessel vector of exactly the order the unit needs (warp-uniform switch); one out-of-line copy per order */
template <int K>
__device__ __noinline__ void fbw_bessel(const double *__restrict__ tabT, int stride, const double *__restrict__ Cj, double z,
                                        double *dst) {
  double Kv[K + 1];
  ecp_bessel<K>(tabT, stride, Cj, K, z, Kv);
#pragma unroll
  for (int i = 0; i <= K; i++) dst[i] = Kv[i];
}
template <int KO>
__device__ __forceinline__ void fbw_bessel_sel(const DevT &t, int lmax, double z, double *dst) {
#define FBW_CASE(K) \
  case K: fbw_bessel<K>(t.besselT, t.besselStride, t.besselC, z, dst); break;
  switch (lmax) {
    FBW_CASE(0)
    FBW_CASE(1)
    FBW_CASE(2)
    FBW_CASE(3)
    FBW_CASE(4)
    FBW_CASE(5)
    FBW_CASE(6)
    default:
      if (KO > 6) {
        switch (lmax) {
          case 7: fbw_bessel<(KO > 6 ? 7 : 6)>(t.besselT, t.besselStride, t.besselC, z, dst); break;
          case 8: fbw_bessel<(KO > 6 ? 8 : 6)>(t.besselT, t.besselStride, t.besselC, z, dst); break;
          case 9: fbw_bessel<(KO > 6 ? 9 : 6)>(t.besselT, t.besselStride, t.besselC, z, dst); break;
          default: fbw_bessel<(KO > 6 ? 10 : 6)>(t.besselT, t.besselStride, t.besselC, z, dst); break;
        }
      }
      break;
  }
#undef FBW_CASE
}

/* ---- one thread per (open unit, point of the wave); a warp = 32 consecutive slots of one unit ----
 * S = points of the wave per unit (32 for levels 0..4, 2^lev for a later level), slot0 = first slot of the wave */
#ifndef FBW_MINB
#define FBW_MINB 8 /* 64 registers, 132 bytes of spills: 0.82 -> 0.77 ms on Au20 against the unbounded 80-register build */
#endif
template <int KO>
__global__ void __launch_bounds__(128, FBW_MINB) k_fbw_eval(DevT t, DevB b, const FbwUnit *units, const FbwQ *qd, const FbwOpen *open,
                                                 long long nWarps, int S, int slot0, double *vals) {
  constexpr int RS = 3 * (KO + 1); /* odd for even KO: conflict-free rows */
  __shared__ double rows[128 * RS];
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nWarps) return;
  const int cpu = S >> 5; /* warps per unit */
  const long long ou = w / cpu;
  const int ch = (int)(w - ou * cpu);
  const int ui = open ? open[ou].unit : (int)ou;
  const FbwUnit u = units[ui];
  double *out = vals + (open ? open[ou].valBase : u.sBase * 32) + 32 * ch + lane;
  const int laC = u.lpack & 15, lbC = (u.lpack >> 4) & 15, lab = (u.lpack >> 8) & 15, l = (u.lpack >> 12) & 15;
  const int slot = slot0 + 32 * ch + lane;
  double *myrow = rows + threadIdx.x * RS;
  double G = 0.0;
  bool live = false;
  const double r = u.i1 * t.large_x[slot] + u.i2; /* src/gc_integrators.c:326-329 */
  if (slot != 1) {                                 /* slot 1 is the pad of the level-major layout */
    const double d1 = u.dAC - r, d2 = u.dBC - r;
    const double e = -u.zA * d1 * d1 - u.zB * d2 * d2;
    live = e >= t.lnAcc2; /* src/type2.c:479-490 */
    if (slot == 0 && r > u.dAC && r > u.dBC && e < t.lnAcc2) atomicAdd(&b.counters[6], 1);
    if (live) {
      const double U = ecp_pot_eval(t.gaussL, t.gaussN, t.gaussD, t.gaussA, u.g0, u.g1, l, r);
      G = (t.large_w[slot] * u.i1) * ((u.Cc * U) * exp(e));
      double rn = 1.0;
      for (int i = 0; i <= lab; i++) {
        myrow[i] = rn;
        rn = r * rn;
      }
    }
  }
  if (__any_sync(0xffffffffu, live)) {
    if (live) {
      fbw_bessel_sel<KO>(t, laC, (2.0 * u.zA * u.dAC) * r, myrow + (KO + 1));
      fbw_bessel_sel<KO>(t, lbC, (2.0 * u.zB * u.dBC) * r, myrow + 2 * (KO + 1));
    }
    for (int j = 0; j < u.nq; j++) {
      const int lll = qd[u.qBase + j].lll;
      /* c_a c_b U r^N K_l1 K_l2 exp(e), times the mapped weight (src/type2.c:403-405) */
      out[(size_t)j * S] = live ? G * (myrow[(lll >> 8) & 15] * myrow[(KO + 1) + (lll & 15)] * myrow[2 * (KO + 1) + ((lll >> 4) & 15)]) : 0.0;
    }
  } else {
    for (int j = 0; j < u.nq; j++) out[(size_t)j * S] = 0.0;
  }
}

/* ---- PSM92 bookkeeping of one wave: 8 lanes per open unit, lane g takes the quadratures g, g + 8, ... ----
 * lev = 4: the wave of slots 0..31 (centre + levels 1..4, four tests); else the level of 2^lev points just evaluated */
__global__ void __launch_bounds__(128) k_fbw_book(DevT t, const FbwUnit *units, const FbwOpen *open, int nOpen, int lev,
                                                 const double *vals, double *sI, double *sP, double *sQ, double *sRes,
                                                 unsigned char *sOpen, FbwOpen *next, unsigned long long *ctr, int *errFlag) {
  const int g = threadIdx.x & 7;
  const long long ou = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = ou < nOpen;
  bool anyOpen = false;
  int ui = 0, nq = 0;
  if (valid) {
    ui = open ? open[ou].unit : (int)ou;
    const FbwUnit u = units[ui];
    nq = u.nq;
    const int S = lev == 4 ? 32 : (1 << lev);
    const double *vb = vals + (open ? open[ou].valBase : u.sBase * 32);
    for (int j = g; j < nq; j += 8) {
      const long long si = u.sBase + j;
      const double *v = vb + (size_t)j * S;
      if (lev == 4) {
        double2 vv[16]; /* the 32 values of the wave, all loads in flight together */
#pragma unroll
        for (int i = 0; i < 16; i++) vv[i] = __ldcs(reinterpret_cast<const double2 *>(v) + i);
        double I = vv[0].x, P = I, Qv = 0.0, res = 0.0; /* centre (slot 1 is the pad) */
        bool op = true;
        int off = 2, np = 1;
#pragma unroll
        for (int lvl = 1; lvl <= 4; lvl++) { /* q = 2p; p = 2I; I += level; test  (src/gc_integrators.c:56-57, 73-83) */
          Qv = 2 * P;
          P = 2 * I;
#pragma unroll
          for (int i = 0; i < (1 << lvl); i += 2) I += vv[(off + i) >> 1].x + vv[(off + i) >> 1].y;
          off += 1 << lvl;
          np = 2 * np + 1;
          if (op && ecp_psm92_test(np, t.tolerance, I, P, Qv)) {
            res = 16 * I / (3 * (np + 1.0));
            op = false;
          }
        }
        sRes[si] = res;
        sOpen[si] = op;
        if (op) {
          sI[si] = I;
          sP[si] = P;
          sQ[si] = Qv;
          anyOpen = true;
        }
      } else if (sOpen[si]) {
        double I = sI[si], P = sP[si], Qv;
        Qv = 2 * P;
        P = 2 * I;
        for (int i0 = 0; i0 < S; i0 += 32) {
          double2 vv[16];
#pragma unroll
          for (int i = 0; i < 16; i++) vv[i] = __ldcs(reinterpret_cast<const double2 *>(v + i0) + i);
#pragma unroll
          for (int i = 0; i < 16; i++) I += vv[i].x + vv[i].y;
        }
        const int np = 2 * S - 1; /* points including this level */
        if (ecp_psm92_test(np, t.tolerance, I, P, Qv)) {
          sRes[si] = 16 * I / (3 * (np + 1.0));
          sOpen[si] = 0;
        } else {
          sI[si] = I;
          sP[si] = P;
          sQ[si] = Qv;
          anyOpen = true;
        }
      }
    }
  }
  /* unit still open: to the next wave's list, with room for its values there (2^(lev+1) points per quadrature) */
  unsigned m = __ballot_sync(0xffffffffu, anyOpen);
  const int lane = threadIdx.x & 31;
  const bool unitOpen = valid && ((m >> (lane & 24)) & 0xffu) != 0;
  if (unitOpen && lev >= t.largeLevels && g == 0) atomicExch(errFlag, 2); /* large grid did not converge (src/libecp.h:26) */
  const bool lead = unitOpen && g == 0 && lev < t.largeLevels;
  const long long pos = fbw_reserve(ctr + 3, lead ? 1 : 0);
  const long long vs = fbw_reserve(ctr + 4, lead ? (long long)nq * (lev == 4 ? 32 : (2 << lev)) : 0);
  if (lead) {
    FbwOpen o;
    o.valBase = vs;
    o.unit = ui;
    o.pad = 0;
    next[pos] = o;
  }
}

/* ---- T(item, quadrature) = sum of the pair integrals in the reference's order (src/type2.c:513) ---- */
__global__ void k_fbw_final(DevB b, const FbwItem *items, const FbwQ *qd, const double *sRes) {
  const int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); /* one warp per item */
  if (it >= b.counters[0]) return;
  const FbwItem I = items[it];
  for (int j = threadIdx.x & 31; j < I.nf; j += 32) {
    double acc = 0.0;
    for (int ip = 0; ip < I.npair; ip++) acc += sRes[I.sBase + (long long)ip * I.nf + j];
    b.T[I.tOff + qd[I.qBase + j].k] = acc;
  }
}

#endif
