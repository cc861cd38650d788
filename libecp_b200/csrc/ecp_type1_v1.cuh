/* ecp_type1_v1.cuh - the round-1 type-1 kernels (static assignment of four primitive pairs per warp, shuffle
 * reductions).  Kept selectable (LIBECP_B200_T1=v1) for A/B timing against the persistent kernels of
 * ecp_type1.cuh; same arithmetic, bit-identical Q. */
#ifndef ECP_TYPE1_V1_CUH
#define ECP_TYPE1_V1_CUH

struct T1Pair { /* per-pair parameters, identical in the 8 lanes of a group */
  long long pr;
  double za, zb, ca, cb, dAC, dBC, sS, zd2, z;
  int type, gs, ge;
};

__device__ __forceinline__ void t1_load_pair(const DevT &t, const DevB &b, long long pr, T1Pair &P) {
  const int tri = b.prTriple[pr];
  const int ssa = b.trA[tri], ssb = b.trB[tri];
  const int sha = b.ssShell[ssa], shb = b.ssShell[ssb];
  const int asa = b.ssASlot[ssa], asb = b.ssASlot[ssb];
  const int Nb = t.shellK[shb];
  const int ip = (int)(pr - b.trPair[tri]), pa = ip / Nb, pb = ip % Nb;
  P.pr = pr;
  P.za = t.primA[t.shellPrim[sha] + pa];
  P.zb = t.primA[t.shellPrim[shb] + pb];
  P.ca = t.primD[t.shellPrim[sha] + pa];
  P.cb = t.primD[t.shellPrim[shb] + pb];
  P.dAC = b.asR[4 * asa + 3];
  P.dBC = b.asR[4 * asb + 3];
  P.type = b.asType[asa];
  P.sS = b.sP[pr];
  P.gs = max(b.ssStart[ssa], b.ssStart[ssb]); /* src/libecp.c:315-316 */
  P.ge = max(b.ssEnd[ssa], b.ssEnd[ssb]);
  P.zd2 = -P.za * P.dAC * P.dAC - P.zb * P.dBC * P.dBC; /* src/type1.c:103 */
  P.z = -P.za - P.zb;
}

/* ---------------------------------------------------------------------------------------------- */
template <int LAB>
__global__ void __launch_bounds__(128) k_type1S_v1(DevT t, DevB b, T1Segs segs, int *failCount, int *failList,
                                                unsigned long long *failMask) {
  constexpr int NQ = T1_NQ(LAB), NQL = (NQ + 7) / 8;
  const int lane = threadIdx.x & 31, gl = lane & 7, gbase = lane & 24;
  const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  bool active = g < segs.prefix[segs.nseg];
  T1Pair pp;
  double Cc = 0.0;
  const double *UL = t.typeUL;
  double *Qo = b.Q;
  if (active) {
    int sg = 0;
    while (segs.prefix[sg + 1] <= g) sg++;
    t1_load_pair(t, b, segs.start[sg] + (g - segs.prefix[sg]), pp);
    Cc = pp.ca * pp.cb * exp(pp.zd2); /* src/type1.c:113 */
    UL = t.typeUL + (size_t)pp.type * ECP_SMALL_SLOTS;
    Qo = b.Q + pair_Q_off(t, b, find_class(b.clsPairBase, t.nClasses, pp.pr), pp.pr);
  }
  double I[NQL], P[NQL], Qv[NQL];
  unsigned open = 0;
#pragma unroll
  for (int k = 0; k < NQL; k++) {
    I[k] = P[k] = Qv[k] = 0.0;
    if (active && 8 * k + gl < NQ) open |= 1u << k;
  }
  int v = 4;    /* level being accumulated by chunks >= 2 */
  int cnt = 0;  /* in-window points of that level */
  for (int c = 0; c < ECP_SMALL_SLOTS / 8; c++) {
    if (!__any_sync(T1_FULL, active)) break;
    const int slot = 8 * c + gl;
    T1Point<LAB> pt;
    pt.live = false;
    pt.w = 0.0;
    bool inWin = false;
    /* chunks >= 2 whose largest left index is below the window and whose smallest right index is above it hold no
     * in-window point for this pair: nothing to load or tabulate */
    const bool possible = active && (c < 2 || t.sm.chMaxL[c] >= pp.gs || t.sm.chMinR[c] <= pp.ge);
    if (possible && slot != 1) {
      const int oi = t.small_oidx[slot];
      /* the three first points are unconditional (src/gc_integrators.c:175-177); afterwards the left point of a
       * pair needs idx >= start, the right one idx <= end (:190-197) */
      inWin = (slot < 4) ? true : ((slot & 1) ? (oi <= pp.ge) : (oi >= pp.gs));
      if (inWin && oi >= pp.gs && oi < pp.ge) { /* tabulated range [start,end): src/type1.c:121 */
        const double r = t.small_r[slot];
        const double e = (pp.z * r + pp.sS) * r;
        if (e >= t.lnAcc1) {
          pt.live = true;
          pt.w = t.small_w[slot];
          pt.u = UL[slot];
          pt.ex = exp(e);
          t1_fill_point<LAB>(t, r, pp.sS * r, pt);
        }
      }
    }
    const unsigned bal = (__ballot_sync(T1_FULL, inWin) >> gbase) & 0xffu;
    if (c == 0) {
      const int cnt0 = __popc(bal & 0x30u), cnt1 = __popc(bal & 0xc0u);
      int q = 0;
#pragma unroll
      for (int N = 0; N <= LAB; N++)
#pragma unroll
        for (int lam = N; lam >= 0; lam -= 2) {
          const double val = t1_wval<LAB>(pt, Cc, N, lam);
          const double v1 = val + __shfl_xor_sync(T1_FULL, val, 1);
          const double c0 = __shfl_sync(T1_FULL, v1, gbase);
          const double f0 = __shfl_sync(T1_FULL, v1, gbase + 2);
          const double a0 = __shfl_sync(T1_FULL, v1, gbase + 4);
          const double a1 = __shfl_sync(T1_FULL, v1, gbase + 6);
          if ((q & 7) == gl) {
            const int k = q >> 3;
            double res;
            P[k] = c0;
            Qv[k] = f0;
            I[k] = P[k] + Qv[k];
            I[k] += a0;
            if ((open >> k & 1) && ecp_ps93_update(t.sm.levJ[0], t.sm.levN[0], cnt0, t.tolerance, I[k], &P[k], &Qv[k], &res)) {
              Qo[N * (LAB + 1) + lam] = res; /* T[l1][l2] += I  (src/type1.c:143) */
              open &= ~(1u << k);
            }
            I[k] += a1;
            if ((open >> k & 1) && ecp_ps93_update(t.sm.levJ[1], t.sm.levN[1], cnt1, t.tolerance, I[k], &P[k], &Qv[k], &res)) {
              Qo[N * (LAB + 1) + lam] = res;
              open &= ~(1u << k);
            }
          }
          q++;
        }
    } else if (c == 1) {
      const int cnt2 = __popc(bal & 0x0fu), cnt3 = __popc(bal & 0xf0u);
      int q = 0;
#pragma unroll
      for (int N = 0; N <= LAB; N++)
#pragma unroll
        for (int lam = N; lam >= 0; lam -= 2) {
          const double val = t1_wval<LAB>(pt, Cc, N, lam);
          const double v1 = val + __shfl_xor_sync(T1_FULL, val, 1);
          const double v2 = v1 + __shfl_xor_sync(T1_FULL, v1, 2);
          const double a2 = __shfl_sync(T1_FULL, v2, gbase);
          const double a3 = __shfl_sync(T1_FULL, v2, gbase + 4);
          if ((q & 7) == gl) {
            const int k = q >> 3;
            double res;
            I[k] += a2;
            if ((open >> k & 1) && ecp_ps93_update(t.sm.levJ[2], t.sm.levN[2], cnt2, t.tolerance, I[k], &P[k], &Qv[k], &res)) {
              Qo[N * (LAB + 1) + lam] = res;
              open &= ~(1u << k);
            }
            I[k] += a3;
            if ((open >> k & 1) && ecp_ps93_update(t.sm.levJ[3], t.sm.levN[3], cnt3, t.tolerance, I[k], &P[k], &Qv[k], &res)) {
              Qo[N * (LAB + 1) + lam] = res;
              open &= ~(1u << k);
            }
          }
          q++;
        }
    } else {
      cnt += __popc(bal);
      const bool last = (8 * c + 8 == t.sm.levSlot[v + 1]);
      /* screened windows leave whole chunks without a single in-window point for all four pairs of the warp
       * (shells far from the centre only see the fine levels): no products, no shuffles then - only the level
       * bookkeeping, which the reference also performs when cnt == 0 (src/gc_integrators.c:203-208) */
      const bool anyWin = __any_sync(T1_FULL, inWin);
      int q = 0;
      if (anyWin) {
#pragma unroll
        for (int N = 0; N <= LAB; N++)
#pragma unroll
          for (int lam = N; lam >= 0; lam -= 2) {
            const double val = t1_wval<LAB>(pt, Cc, N, lam);
            const double v1 = val + __shfl_xor_sync(T1_FULL, val, 1);
            const double v2 = v1 + __shfl_xor_sync(T1_FULL, v1, 2);
            const double v4 = v2 + __shfl_xor_sync(T1_FULL, v2, 4);
            if ((q & 7) == gl) {
              const int k = q >> 3;
              double res;
              I[k] += v4;
              if (last && (open >> k & 1) &&
                  ecp_ps93_update(t.sm.levJ[v], t.sm.levN[v], cnt, t.tolerance, I[k], &P[k], &Qv[k], &res)) {
                Qo[N * (LAB + 1) + lam] = res;
                open &= ~(1u << k);
              }
            }
            q++;
          }
      } else if (last) {
#pragma unroll
        for (int N = 0; N <= LAB; N++)
#pragma unroll
          for (int lam = N; lam >= 0; lam -= 2) {
            if ((q & 7) == gl) {
              const int k = q >> 3;
              double res;
              if ((open >> k & 1) &&
                  ecp_ps93_update(t.sm.levJ[v], t.sm.levN[v], cnt, t.tolerance, I[k], &P[k], &Qv[k], &res)) {
                Qo[N * (LAB + 1) + lam] = res;
                open &= ~(1u << k);
              }
            }
            q++;
          }
      }
      if (last) {
        v++;
        cnt = 0;
      }
    }
    /* group finished when none of its lanes has an open quadrature */
    const unsigned ob = (__ballot_sync(T1_FULL, open != 0) >> gbase) & 0xffu;
    if (!ob) active = false;
  }
  /* quadratures that never converged on the small grid -> large grid (src/type1.c:149) */
  unsigned long long m = 0;
#pragma unroll
  for (int k = 0; k < NQL; k++)
    if (open >> k & 1) m |= 1ull << (8 * k + gl);
  m |= __shfl_xor_sync(T1_FULL, m, 1);
  m |= __shfl_xor_sync(T1_FULL, m, 2);
  m |= __shfl_xor_sync(T1_FULL, m, 4);
  if (m && gl == 0) {
    failMask[pp.pr] = m;
    failList[atomicAdd(failCount, 1)] = (int)pp.pr;
  }
}

/* ---------------------------------------------------------------------------------------------- */
template <int LAB>
__global__ void __launch_bounds__(128) k_type1L_v1(DevT t, DevB b, const int *failCount, const int *failList,
                                                const unsigned long long *failMask, int *errFlag) {
  constexpr int NQ = T1_NQ(LAB), NQL = (NQ + 7) / 8;
  const int lane = threadIdx.x & 31, gl = lane & 7, gbase = lane & 24;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  bool active = g < *failCount;
  T1Pair pp;
  double Cc = 0.0, i1 = 0.0, i2 = 0.0;
  int Lc = 0, g0 = 0, g1 = 0;
  double *Qo = b.Q;
  unsigned open = 0;
  if (active) {
    t1_load_pair(t, b, failList[g], pp);
    const unsigned long long fm = failMask[pp.pr];
#pragma unroll
    for (int k = 0; k < NQL; k++)
      if (fm >> (8 * k + gl) & 1) open |= 1u << k;
    Cc = pp.ca * pp.cb; /* src/type1.c:151 */
    const double zp = pp.za + pp.zb;
    ecp_fm06_map(zp, (pp.za * pp.dAC + pp.zb * pp.dBC) / zp, &i1, &i2);
    Lc = t.typeL[pp.type];
    g0 = t.typeGaussOff[pp.type];
    g1 = t.typeGaussOff[pp.type + 1];
    Qo = b.Q + pair_Q_off(t, b, find_class(b.clsPairBase, t.nClasses, pp.pr), pp.pr);
  }
  double I[NQL], P[NQL], Qv[NQL];
#pragma unroll
  for (int k = 0; k < NQL; k++) I[k] = P[k] = Qv[k] = 0.0;
  int lev = 3, n = 7; /* level accumulated by chunks >= 1: slots [2^lev, 2^(lev+1)); n = points before it */
  const int nChunks = t.largeSlots / 8;
  for (int c = 0; c < nChunks; c++) {
    if (!__any_sync(T1_FULL, active)) break;
    const int slot = 8 * c + gl;
    T1Point<LAB> pt;
    pt.live = false;
    pt.w = 0.0;
    if (active && slot != 1) {
      const double r = i1 * t.large_x[slot] + i2;           /* src/gc_integrators.c:326-329 */
      const double e = (pp.z * r + pp.sS) * r + pp.zd2;     /* src/type1.c:162 */
      if (e >= t.lnAcc1) {
        pt.live = true;
        pt.w = t.large_w[slot] * i1;
        pt.u = ecp_pot_eval(t.gaussL, t.gaussN, t.gaussD, t.gaussA, g0, g1, Lc, r);
        pt.ex = exp(e);
        t1_fill_point<LAB>(t, r, pp.sS * r, pt);
      }
    }
    const bool first = (c == 0) || (8 * c == (1 << lev));
    const bool last = (c == 0) || (8 * c + 8 == (2 << lev));
    int q = 0;
#pragma unroll
    for (int N = 0; N <= LAB; N++)
#pragma unroll
      for (int lam = N; lam >= 0; lam -= 2) {
        const double val = t1_wval<LAB>(pt, Cc, N, lam);
        const double v1 = val + __shfl_xor_sync(T1_FULL, val, 1);
        const double v2 = v1 + __shfl_xor_sync(T1_FULL, v1, 2);
        if (c == 0) {
          /* slot 0 = centre, slots 2,3 = level 1, slots 4..7 = level 2 */
          const double c0 = __shfl_sync(T1_FULL, v1, gbase);
          const double a1 = __shfl_sync(T1_FULL, v1, gbase + 2);
          const double a2 = __shfl_sync(T1_FULL, v2, gbase + 4);
          if ((q & 7) == gl) {
            const int k = q >> 3;
            double res;
            I[k] = c0; /* I = w[M] f(M); p = I  (src/gc_integrators.c:49-52) */
            P[k] = I[k];
            Qv[k] = 2 * P[k];
            P[k] = 2 * I[k];
            I[k] += a1;
            if ((open >> k & 1) && ecp_psm92_update(3, 1, t.tolerance, I[k], P[k], Qv[k], &res)) {
              Qo[N * (LAB + 1) + lam] = res; /* T[l1][l2] += grid->I  (src/type1.c:193) */
              open &= ~(1u << k);
            }
            Qv[k] = 2 * P[k];
            P[k] = 2 * I[k];
            I[k] += a2;
            if ((open >> k & 1) && ecp_psm92_update(7, 1, t.tolerance, I[k], P[k], Qv[k], &res)) {
              Qo[N * (LAB + 1) + lam] = res;
              open &= ~(1u << k);
            }
          }
        } else {
          const double v4 = v2 + __shfl_xor_sync(T1_FULL, v2, 4);
          if ((q & 7) == gl) {
            const int k = q >> 3;
            double res;
            if (first) { /* q = 2p; p = 2I  (src/gc_integrators.c:56-57) */
              Qv[k] = 2 * P[k];
              P[k] = 2 * I[k];
            }
            I[k] += v4;
            if (last && (open >> k & 1) && ecp_psm92_update(2 * n + 1, 1, t.tolerance, I[k], P[k], Qv[k], &res)) {
              Qo[N * (LAB + 1) + lam] = res;
              open &= ~(1u << k);
            }
          }
        }
        q++;
      }
    if (c > 0 && last) {
      n = 2 * n + 1;
      lev++;
    }
    const unsigned ob = (__ballot_sync(T1_FULL, open != 0) >> gbase) & 0xffu;
    if (!ob) active = false;
  }
  const unsigned ob = (__ballot_sync(T1_FULL, open != 0) >> gbase) & 0xffu;
  if (ob && gl == 0) atomicExch(errFlag, 1); /* large grid failed: rc 1 (src/libecp.h:25) */
}

#endif
