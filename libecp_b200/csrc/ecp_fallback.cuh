/* ecp_fallback.cuh - type-2 large-grid fallback: T(l; lambda1, lambda2, N) of the quadratures whose small-grid
 * product quadrature failed, redone per primitive pair on the FM06-mapped 1023-point grid with PSM92.
 *
 * Replaces calcT_FM06 (reference src/type2.c:417-528) + QIntegrand_FM06 (:397-410) + integrateGC_PSM92
 * (src/gc_integrators.c:38-86).
 *
 * Mapping: EIGHT LANES PER (triple, l) ITEM, four items per warp, persistent groups fed from the item list that
 * k_fastT produced (same structure as the type-1 kernels, ecp_type1.cuh):
 *   - an item has few failed quadratures (median 4, 90 % <= 18 on Au20) but needs ~130 tabulated points per primitive
 *     pair, each costing two Bessel evaluations, the potential and an exponential.  So the points go across the lanes
 *     (one slot of the level-major large-grid layout per lane and chunk, 8-slot granularity: nothing beyond the level
 *     the PSM92 rule stops at is tabulated) and the Bessel functions are evaluated in registers (compile-time order
 *     bound KO) instead of through shared memory;
 *   - per chunk every lane leaves r^n, K_a, K_b of its point in a shared-memory row (the failed quadratures index
 *     them dynamically), forms the integrand of each failed quadrature of the current pass (<= 16 per pass, two per
 *     owner lane) and stores it in a padded tile [quadrature][lane]; the owner lane adds the 8 values as
 *     ((a0+a1)+(a2+a3))+((a4+a5)+(a6+a7)) and runs the PSM92 bookkeeping for its quadratures;
 *   - primitive pairs are visited in the reference's order and their converged integrals added in that order
 *     (T += grid->I, src/type2.c:513); items with more than 16 failed quadratures take several passes.
 */
#ifndef ECP_FALLBACK_CUH
#define ECP_FALLBACK_CUH

#define FB_NQ 16  /* failed quadratures per pass: two per lane */
#define FB_TILE 152 /* 16 rows of 9 doubles, rounded to = 8 mod 16 (see T1Cfg::GS) */

template <int KO>
struct FbCfg {
  static constexpr int RS = 3 * (KO + 1);             /* row: r^0..r^KO, Ka[0..KO], Kb[0..KO]; odd for KO even */
  static constexpr int GROUP = 8 * RS + FB_TILE + 16; /* doubles per group: rows, tile, 32 ints (k index, packed l's) */
};
template <int KO>
static size_t fb_smem_bytes(int block) {
  return (size_t)(block / 8) * FbCfg<KO>::GROUP * sizeof(double);
}

/* K_0..K_KO(z) into a shared-memory row; one copy of the Bessel code for both arguments (the kernel is bound by
 * instruction fetch: keep its footprint small) */
template <int KO>
__device__ __noinline__ void fb_bessel_row(const double *__restrict__ tabT, int stride, const double *__restrict__ Cj,
                                           int lmax, double z, double *dst) {
  double K[KO + 1];
  ecp_bessel<KO>(tabT, stride, Cj, lmax, z, K);
#pragma unroll
  for (int i = 0; i <= KO; i++) dst[i] = K[i];
}

template <int KO, int MINB>
__global__ void __launch_bounds__(128, MINB) k_fallbackG(DevT t, DevB b) {
  using Cfg = FbCfg<KO>;
  constexpr int RS = Cfg::RS;
  extern __shared__ __align__(16) double fb_smem[];
  const int lane = threadIdx.x & 31, gl = lane & 7, gbase = lane & 24;
  double *gsm = fb_smem + (size_t)(threadIdx.x >> 3) * Cfg::GROUP;
  double *myrow = gsm + gl * RS;
  double *tile = gsm + 8 * RS;
  int *qK = (int *)(tile + FB_TILE), *qQ = qK + FB_NQ; /* position in the class list / packed l | l1<<4 | l2<<8 | l3<<12 */
  const unsigned long long dbgT0 = b.dbg ? ecp_gtimer() : 0;
  const int nItems = b.counters[0];
  const int nChunks = t.largeSlots / 8;
  /* item state */
  bool have = false, drained = false;
  int nf = 0, pass = 0, nq = 0;       /* failed quadratures of the item, current pass, quadratures in it */
  int k0 = 0, k1 = 0, l = 0, laC = 0, lbC = 0, lab = 0, Na = 0, Nb = 0, pa = 0, pb = 0, g0 = 0, g1 = 0;
  long long tOff = 0;
  const int *ql = t.qlist;
  const double *za = t.primA, *zb = t.primA, *ca = t.primD, *cb = t.primD;
  double dAC = 0.0, dBC = 0.0;
  /* primitive-pair state */
  double s1 = 0.0, s2 = 0.0, Cc = 0.0, zA = 0.0, zB = 0.0, i1 = 0.0, i2 = 0.0;
  int c = 0, lev = 3, n = 7;
  bool itemFailed = false;
  /* quadrature state of the two quadratures this lane owns in the pass */
  double I[2] = {0.0, 0.0}, P[2] = {0.0, 0.0}, Qv[2] = {0.0, 0.0}, Acc[2] = {0.0, 0.0};
  unsigned open = 0;
  bool newPass = false, newPair = false;
  for (;;) {
    /* ---- next item ---- */
    const bool need = !have && !drained;
    if (__any_sync(0xffffffffu, need)) {
      int it = 0;
      if (need && gl == 0) it = atomicAdd(&b.counters[1], 1);
      it = __shfl_sync(0xffffffffu, it, gbase);
      if (need) {
        if (it < nItems) {
          const int item = b.items[it];
          const int tri = item >> 3;
          l = item & 7;
          const int cl = find_class_i(b.clsFirst, t.nClasses, tri);
          const int la = t.clsLa[cl], lb = t.clsLb[cl];
          laC = la + l;
          lbC = lb + l;
          lab = la + lb;
          k0 = t.clsQlOff[cl * (ECP_MAX_LECP + 1) + l];
          k1 = t.clsQlOff[cl * (ECP_MAX_LECP + 1) + l + 1];
          tOff = tri_T_off(t, b, cl, tri);
          ql = t.qlist + t.clsQOff[cl];
          const int ssa = b.trA[tri], ssb = b.trB[tri];
          const int sha = b.ssShell[ssa], shb = b.ssShell[ssb];
          const int asa = b.ssASlot[ssa], asb = b.ssASlot[ssb];
          dAC = b.asR[4 * asa + 3];
          dBC = b.asR[4 * asb + 3];
          const int type = b.asType[asa];
          g0 = t.typeGaussOff[type];
          g1 = t.typeGaussOff[type + 1];
          /* Gaussians of channel l form one run of the type's list (they are stored shell by shell): evaluate only
           * that run per point (the loop over the others only tests and skips, src/ecp.c:47-57) */
          while (g0 < g1 && t.gaussL[g0] != l) g0++;
          while (g1 > g0 && t.gaussL[g1 - 1] != l) g1--;
          Na = t.shellK[sha];
          Nb = t.shellK[shb];
          za = t.primA + t.shellPrim[sha];
          ca = t.primD + t.shellPrim[sha];
          zb = t.primA + t.shellPrim[shb];
          cb = t.primD + t.shellPrim[shb];
          have = true;
          pass = 0;
          newPass = true;
          itemFailed = false;
        } else {
          drained = true;
        }
      }
      /* number of failed quadratures of a newly fetched item: counted by its 8 lanes, summed over the group */
      int part = 0;
      if (need && have)
        for (int k = k0 + gl; k < k1; k += 8) part += b.tfail[tOff + k] ? 1 : 0;
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (need && have) nf = part;
    }
    if (!__any_sync(0xffffffffu, have)) break;
    /* ---- new pass: the failed quadratures with rank [16 pass, 16 pass + 16) go to qK / qQ ---- */
    if (__any_sync(0xffffffffu, newPass)) {
      int seen = 0; /* failed quadratures before the current scan position */
      const int lo = FB_NQ * pass;
      for (int base = 0;; base += 8) {
        const int k = k0 + base + gl;
        const bool inRange = newPass && have && k < k1;
        if (!__any_sync(0xffffffffu, newPass && have && (k0 + base < k1))) break;
        const bool f = inRange && b.tfail[tOff + k];
        const unsigned m = (__ballot_sync(0xffffffffu, f) >> gbase) & 0xffu;
        if (f) {
          const int rank = seen + __popc(m & ((1u << gl) - 1)) - lo;
          if (rank >= 0 && rank < FB_NQ) {
            qK[rank] = k;
            qQ[rank] = ql[k];
          }
        }
        seen += __popc(m);
      }
      if (newPass && have) {
        nq = nf - lo < FB_NQ ? nf - lo : FB_NQ;
        Acc[0] = Acc[1] = 0.0;
        pa = 0;
        pb = 0;
        newPair = true;
      }
      newPass = false;
      __syncwarp();
    }
    /* ---- new primitive pair (src/type2.c:452-468) ---- */
    if (newPair && have) {
      zA = za[pa];
      zB = zb[pb];
      s1 = 2.0 * zA * dAC;
      s2 = 2.0 * zB * dBC;
      Cc = ca[pa] * cb[pb];
      const double zp = zA + zB;
      ecp_fm06_map(zp, (zA * dAC + zB * dBC) / zp, &i1, &i2);
      open = 0;
      if (gl < nq) open |= 1u;
      if (8 + gl < nq) open |= 2u;
      c = 0;
      lev = 3;
      n = 7;
    }
    newPair = false;
    /* ---- one chunk: every lane tabulates one slot (src/type2.c:471-495) ---- */
    const int slot = 8 * c + gl;
    double W = 0.0, CU = 0.0, EX = 0.0;
    bool live = false;
    if (have && slot != 1) {
      const double r = i1 * t.large_x[slot] + i2; /* src/gc_integrators.c:326-329 */
      const double d1 = dAC - r, d2 = dBC - r;
      const double e = -zA * d1 * d1 - zB * d2 * d2;
      live = e >= t.lnAcc2;
      if (slot == 0 && r > dAC && r > dBC && e < t.lnAcc2) atomicAdd(&b.counters[6], 1);
      if (live) {
        const double U = ecp_pot_eval(t.gaussL, t.gaussN, t.gaussD, t.gaussA, g0, g1, l, r);
        fb_bessel_row<KO>(t.besselT, t.besselStride, t.besselC, laC, s1 * r, myrow + (KO + 1));
        fb_bessel_row<KO>(t.besselT, t.besselStride, t.besselC, lbC, s2 * r, myrow + 2 * (KO + 1));
        W = t.large_w[slot] * i1;
        CU = Cc * U;
        EX = exp(e);
        double rn = 1.0;
#pragma unroll
        for (int i = 0; i <= KO; i++) {
          myrow[i] = rn;
          rn = r * rn;
        }
      }
    }
    /* ---- integrands of the pass's quadratures at my point -> tile[quadrature][lane] ---- */
    {
      const double G = W * (CU * EX); /* common to all quadratures of the point */
      int nqMax = have ? nq : 0;
      nqMax = max(nqMax, __shfl_xor_sync(0xffffffffu, nqMax, 8));
      nqMax = max(nqMax, __shfl_xor_sync(0xffffffffu, nqMax, 16));
      __syncwarp();
      for (int j = 0; j < nqMax; j++) {
        double val = 0.0;
        if (live && j < nq) {
          const int qq = qQ[j];
          const int l1 = (qq >> 4) & 15, l2 = (qq >> 8) & 15, l3 = (qq >> 12) & 15;
          /* c_a c_b U r^N K_l1 K_l2 exp(e), times the mapped weight (src/type2.c:403-405) */
          val = G * (myrow[l3] * myrow[(KO + 1) + l1] * myrow[2 * (KO + 1) + l2]);
        }
        if (j < nq || !have) tile[j * 9 + gl] = val;
      }
      __syncwarp();
    }
    /* ---- owner lanes: PSM92 bookkeeping (src/gc_integrators.c:49-83), same chunk structure as k_type1L ---- */
    if (have) {
      const bool first = (8 * c == (1 << lev));
      const bool last = (8 * c + 8 == (2 << lev));
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int j = 8 * k + gl;
        const double *row = tile + (j < nq ? j : 0) * 9;
        const double A = row[0] + row[1], B = row[2] + row[3], Cq = row[4] + row[5], D = row[6] + row[7];
        double hitI = 0.0;
        int hitN = 0; /* points of the level at which the quadrature converged (0: not in this chunk) */
        if (c == 0) {
          /* slot 0 = centre, slots 2,3 = level 1, slots 4..7 = level 2 */
          I[k] = A;
          P[k] = I[k];
          Qv[k] = 2 * P[k];
          P[k] = 2 * I[k];
          I[k] += B;
          if ((open >> k & 1) && ecp_psm92_test(3, t.tolerance, I[k], P[k], Qv[k])) {
            hitI = I[k];
            hitN = 3;
            open &= ~(1u << k);
          }
          Qv[k] = 2 * P[k];
          P[k] = 2 * I[k];
          I[k] += (Cq + D);
          if ((open >> k & 1) && ecp_psm92_test(7, t.tolerance, I[k], P[k], Qv[k])) {
            hitI = I[k];
            hitN = 7;
            open &= ~(1u << k);
          }
        } else {
          if (first) { /* q = 2p; p = 2I  (src/gc_integrators.c:56-57) */
            Qv[k] = 2 * P[k];
            P[k] = 2 * I[k];
          }
          I[k] += ((A + B) + (Cq + D));
          if (last && (open >> k & 1) && ecp_psm92_test(2 * n + 1, t.tolerance, I[k], P[k], Qv[k])) {
            hitI = I[k];
            hitN = 2 * n + 1;
            open &= ~(1u << k);
          }
        }
        /* T += grid->I = 16 I / (3 N), N = points + 1  (src/gc_integrators.c:80, src/type2.c:513) */
        if (hitN) Acc[k] += 16 * hitI / (3 * (hitN + 1.0));
      }
      if (c > 0 && last) {
        n = 2 * n + 1;
        lev++;
      }
      c++;
    }
    /* ---- pair finished?  -> next pair / next pass / item done ---- */
    const unsigned ob = (__ballot_sync(0xffffffffu, have && open != 0) >> gbase) & 0xffu;
    if (have && (ob == 0 || c == nChunks)) {
      if (ob != 0) itemFailed = true; /* large grid did not converge: rc 2 (src/libecp.h:26) */
      pb++;
      if (pb == Nb) {
        pb = 0;
        pa++;
      }
      if (pa < Na) {
        newPair = true;
      } else {
        /* all primitive pairs of this pass done: store T of the owned quadratures */
        if (gl < nq) b.T[tOff + qK[gl]] = Acc[0];
        if (8 + gl < nq) b.T[tOff + qK[8 + gl]] = Acc[1];
        pass++;
        if (FB_NQ * pass < nf) {
          newPass = true;
        } else {
          if (itemFailed && gl == 0) atomicExch(&b.counters[3], 2);
          have = false;
        }
        open = 0;
      }
    }
  }
  if (b.dbg && threadIdx.x == 0 && blockIdx.x < DBG_STRIDE) {
    b.dbg[2 * blockIdx.x] = dbgT0;
    b.dbg[2 * blockIdx.x + 1] = ecp_gtimer();
  }
}


#endif
