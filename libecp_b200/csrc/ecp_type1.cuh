/* ecp_type1.cuh - type-1 (local) radial integrals Q[N][lambda] per primitive pair.
 *
 * Replaces calcQ (reference src/type1.c:94-208) + QIntegrand (:78-88) + the quadrature drivers it calls
 * (src/gc_integrators.c:156-217 PS93, :38-86 PSM92).
 *
 * Mapping: EIGHT LANES PER PRIMITIVE PAIR (four pairs per warp), templated on LAB = la+lb, PERSISTENT groups.
 *   - k_t1prep writes one 80-byte record per primitive pair (exponent sum, |P|, prefactors, FM06 map, window, Q offset);
 *     a group fetches the next pair of its launch from an atomic work counter as soon as its current pair is finished.
 *     About 15 % of the pairs never converge on the small grid and walk all 48 chunks while the typical pair needs 2-4
 *     chunks: with a static assignment one such pair kept the other three groups of its warp idle (the first kernels of
 *     this round spent ~2x the necessary warp-chunks that way; 2.5 ms -> 1.3 ms on Au20, profiles/r1/ab_kernels.jsonl).
 *   - quadrature points are spread across the 8 lanes: every chunk each lane tabulates ONE grid point of the
 *     level-major slot layout (Bessel K_0..K_LAB, r^0..r^LAB, U_L, exp) in registers - only points a level of
 *     the adaptive rule really needs ("touched" points; the reference tabulates the whole window up front,
 *     src/type1.c:121-130);
 *   - the per-point products w*f of all NQ = sum_N (N/2+1) quadratures go through a padded shared-memory tile
 *     [quadrature][lane]; the lane that owns quadrature q (q mod 8) reads the 8 values back and adds them as
 *     ((a0+a1)+(a2+a3))+((a4+a5)+(a6+a7)): T = left + right as in the reference, then across pairs.  This replaces
 *     3 double shuffles per quadrature per lane by one store + one load per owned quadrature and makes the level
 *     update (PS93 / PSM92 state I,p,q of quadrature q lives in lane q mod 8) run on all lanes at once;
 *   - convergence is a ballot over the group.
 *   Measured and rejected in round 2 (profiles/r2/README.md): a "dense" variant that strings the live points of a pair
 *   (inside the window and below the exponent gate) into one sequence across the levels and lets a chunk span levels
 *   halves the chunks of a typical pair, but its per-pair set-up (gate search, level table) and the point-by-point
 *   owner loop cost more than the chunks saved: 1.17 -> 1.30 ms on Au20, 30.9 -> 38.5 ms per config-5 pass.
 *   Round-2 session 3, tools/t1_stats.py (-DT1_STATS build): on configuration 5 half of all 8-slot chunks are the two fixed
 *   ones of a pair (slots 0..15) with 2.1 live lanes of 8, the level-wise chunks hold 6.6 of 8; 3.6 % of the pairs fail on
 *   the small grid and take 40 % of the chunks.  Evaluating the first 16 slots in ONE chunk when at most 8 are live (pre-test
 *   of window and gate, lane i takes the i-th live slot, owners pick the values out by slot number; Q bit-identical) removes
 *   a quarter of the chunks but costs ~250 instructions in every warp iteration in which any of the four groups starts a
 *   pair (three out of four): type-1 23.4 -> 24.8 ms per configuration-5 pass.  Removed (profiles/r2/session3).
 *   k_type1S<LAB>: small 383-point grid, PS93; writes converged Q, records a mask of failed quadratures
 *   k_type1L<LAB>: failed quadratures on the per-pair FM06-mapped 1023-point grid, PSM92
 */
#ifndef ECP_TYPE1_CUH
#define ECP_TYPE1_CUH

#include <utility>

#define T1_MAXSEG 24
/* -DT1_STATS: lane / chunk statistics of k_type1S (diagnostic builds only, tools/t1_stats.py) */
#ifdef T1_STATS
__device__ unsigned long long g_t1stats[16];
#define T1_STAT(i, v) atomicAdd(&g_t1stats[i], (unsigned long long)(v))
#else
#define T1_STAT(i, v) ((void)0)
#endif
/* resident blocks of 128 threads per SM the type-1 kernels are compiled for (registers <= 65536 / (128 MINB)) */
#ifndef T1_MINB_LO
#define T1_MINB_LO 5
#endif
#ifndef T1_MINB_MID
#define T1_MINB_MID 4
#endif
struct T1Segs {
  int nseg;
  long long start[T1_MAXSEG];      /* first pair of the segment                      */
  long long prefix[T1_MAXSEG + 1]; /* running pair count over the segments           */
};

template <int LAB>
struct T1Point {
  double w, u, ex;
  double rn[LAB + 1], K[LAB + 1];
  bool live; /* tabulated and exponent above the gate */
};

/* number of (N, lambda) quadratures: sum_{N<=LAB} (N/2+1) */
#define T1_NQ(LAB) (((LAB) % 2 == 0) ? ((LAB) / 2 + 1) * ((LAB) / 2 + 1) : ((LAB) / 2 + 1) * ((LAB) / 2 + 2))
#define T1_FULL 0xffffffffu
__constant__ unsigned char c_t1qoff[11][36]; /* [LAB][q] -> N (LAB + 1) + lambda; filled by t1_upload_qoff */

/* quadrature q of the loop nest "for N: for lambda = N, N-2, ..." (src/type1.c:132-146) */
__host__ __device__ constexpr int t1_qN(int q) {
  int N = 0;
  while (q >= N / 2 + 1) {
    q -= N / 2 + 1;
    N++;
  }
  return N;
}
__host__ __device__ constexpr int t1_qLam(int q) {
  int N = 0;
  while (q >= N / 2 + 1) {
    q -= N / 2 + 1;
    N++;
  }
  return N - 2 * q;
}

template <int LAB>
struct T1Cfg {
  static constexpr int NQ = T1_NQ(LAB);
  static constexpr int NQL = (NQ + 7) / 8;
  static constexpr int ROW = 9;                               /* 8 lane values + 1 pad: conflict-free owner reads */
  static constexpr int GS = ((NQ * ROW + 7) / 16) * 16 + 8;   /* doubles per group, = 8 mod 16: the two groups of a
                                                                 half warp store to disjoint banks */
};
static size_t t1_smem_bytes(int lab, int block) {
  const int nq = T1_NQ(lab), gs = ((nq * 9 + 7) / 16) * 16 + 8;
  return (size_t)(block / 32) * 4 * gs * sizeof(double);
}

/* w * C * r^N * U * K_lambda * exp(e)  (integrand of src/type1.c:84).  The factors common to all quadratures of a point
 * are multiplied once: G = w (C U exp(e)), H[N] = G r^N, value = H[N] K_lambda - 3 + (LAB+1) + NQ multiplications per
 * point instead of 5 NQ (the same factors as the reference, associated differently). */
template <int LAB, int... Q>
__device__ __forceinline__ void t1_store_vals(const T1Point<LAB> &pt, double Cc, double *dst,
                                              std::integer_sequence<int, Q...>) {
  double H[LAB + 1];
  const double G = pt.live ? pt.w * ((Cc * pt.u) * pt.ex) : 0.0;
#pragma unroll
  for (int i = 0; i <= LAB; i++) H[i] = pt.live ? G * pt.rn[i] : 0.0;
  ((dst[Q * T1Cfg<LAB>::ROW] = pt.live ? H[t1_qN(Q)] * pt.K[t1_qLam(Q)] : 0.0), ...);
}

/* per primitive pair, written by k_t1prep */
struct __align__(16) T1Rec {
  double z;        /* -(za + zb)                                   src/type1.c:104 */
  double sS;       /* |P|, P = 2 (za r_AC + zb r_BC)               src/type1.c:249-250 */
  double zd2;      /* -za dAC^2 - zb dBC^2                         src/type1.c:103 */
  double CcS;      /* ca cb exp(zd2)   (small grid)                src/type1.c:113 */
  double CcL;      /* ca cb            (large grid)                src/type1.c:151 */
  double i1, i2;   /* FM06 map r = i1 x + i2                       src/gc_integrators.c:316-331 */
  long long qoff;  /* offset of Q[N][lambda] / S_lm(P^) of the pair */
  int type, gs;    /* ECP type of the centre; window start         src/libecp.c:315 */
  int ge, pad;     /* window end                                   src/libecp.c:316 */
};

template <int LAB>
__device__ __forceinline__ void t1_fill_point(const DevT &t, double r, double zarg, T1Point<LAB> &p) {
  ecp_bessel<LAB>(t.besselT, t.besselStride, t.besselC, LAB, zarg, p.K);
  double rn = 1.0;
#pragma unroll
  for (int i = 0; i <= LAB; i++) {
    p.rn[i] = rn;
    rn = r * rn;
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* k_type1A<LAB>: the first 16 slots of every primitive pair (the three unconditional points and levels 0..3 of the PS93
 * rule, src/gc_integrators.c:175-199) as a block-wide wave.  tools/t1_stats.py: in the 8-lanes-per-pair kernel these two
 * fixed chunks are half of all chunks of a configuration-5 pass and hold 2.1 live lanes of 8 (window and exponent gate).
 * Here a block takes PB pairs: (1) a thread per (pair, slot) tests window and gate and appends the live ones to a list in
 * shared memory; (2) a thread per LIVE point evaluates it - potential, exponential, Bessel vector, powers - and stores the
 * products of all NQ quadratures into a zeroed shared-memory table [pair][slot][quadrature] (full warps instead of a third
 * of the lanes); (3) a thread per (pair, quadrature) runs the bookkeeping of levels 0..3 on its 16 values with exactly
 * the sums of k_type1S (pairs first, absent points are exact zeros) and either writes the converged Q or leaves
 * (I, p, q) in `state`; (4) pairs with an open quadrature go to the survivor list, which k_type1S continues level-wise
 * from level 4.  Q is bit-identical to the one-kernel path (LIBECP_B200_T1=legacy; test). */
#ifndef T1A_MINB
#define T1A_MINB 2
#endif
#ifndef T1A_PBNUM
#define T1A_PBNUM 8 /* eighths of the default number of pairs per block (A/B builds) */
#endif
template <int LAB>
struct T1ACfg {
  static constexpr int NQ = T1_NQ(LAB);
  static constexpr int NQP = NQ | 1;                                 /* odd row: conflict-free stores of a point's products */
  static constexpr int PB = (LAB <= 4 ? 64 : (LAB <= 6 ? 32 : 16)) * T1A_PBNUM / 8; /* pairs per block: 72 / 68 / 74 KB of values at most (T1A_PBNUM = 8) */
};
struct T1ARec {
  double z, sS, Cc;
  const double *UL;
  double *Qo;
  long long pr;
  int gs, ge;
  unsigned wm, lm; /* in-window / live masks of the 16 slots */
  unsigned open0, open1;
};
static size_t t1a_smem_bytes(int lab) {
  const int nq = T1_NQ(lab), nqp = nq | 1, pb = (lab <= 4 ? 64 : (lab <= 6 ? 32 : 16)) * T1A_PBNUM / 8;
  return (size_t)pb * 16 * nqp * sizeof(double) + (size_t)pb * sizeof(T1ARec) + (size_t)pb * 16 * sizeof(unsigned short);
}
static void t1_upload_qoff(void) {
  unsigned char h[11][36] = {{0}};
  for (int lab = 0; lab <= 10; lab++)
    for (int q = 0; q < T1_NQ(lab); q++) h[lab][q] = (unsigned char)(t1_qN(q) * (lab + 1) + t1_qLam(q));
  cudaMemcpyToSymbol(c_t1qoff, h, sizeof(h));
}
template <int LAB, int... Q>
__device__ __forceinline__ void t1a_store_vals(const T1Point<LAB> &pt, double Cc, double *dst, std::integer_sequence<int, Q...>) {
  double H[LAB + 1];
  const double G = pt.w * ((Cc * pt.u) * pt.ex);
#pragma unroll
  for (int i = 0; i <= LAB; i++) H[i] = G * pt.rn[i];
  ((dst[Q] = H[t1_qN(Q)] * pt.K[t1_qLam(Q)]), ...);
}
template <int LAB>
__global__ void __launch_bounds__(256, T1A_MINB) k_type1A(DevT t, DevB b, T1Segs segs, int *survCount, int *survList,
                                                   unsigned long long *survMask, double *state) {
  using Cfg = T1ACfg<LAB>;
  constexpr int NQ = Cfg::NQ, NQP = Cfg::NQP, PB = Cfg::PB;
  extern __shared__ __align__(16) unsigned char t1a_raw[];
  double *vals = (double *)t1a_raw; /* [PB][16][NQP] */
  T1ARec *rec = (T1ARec *)(vals + (size_t)PB * 16 * NQP);
  unsigned short *items = (unsigned short *)(rec + PB);
  __shared__ int cnt;
  const int tid = threadIdx.x;
  const int total = (int)segs.prefix[segs.nseg];
  const int g0 = blockIdx.x * PB;
  const int np = min(PB, total - g0);
  if (np <= 0) return;
  if (tid < np) {
    const int g = g0 + tid;
    int sg = 0;
    while (segs.prefix[sg + 1] <= g) sg++;
    const long long pr = segs.start[sg] + (g - segs.prefix[sg]);
    const T1Rec r = b.t1rec[pr];
    T1ARec &a = rec[tid];
    a.z = r.z;
    a.sS = r.sS;
    a.Cc = r.CcS;
    a.UL = t.typeUL + (size_t)r.type * ECP_SMALL_SLOTS;
    a.Qo = b.Q + r.qoff;
    a.pr = pr;
    a.gs = r.gs;
    a.ge = r.ge;
    a.wm = a.lm = 0;
    a.open0 = a.open1 = 0;
  }
  if (tid == 0) cnt = 0;
  for (int i = tid; i < np * 16 * NQP; i += 256) vals[i] = 0.0;
  __syncthreads();
  /* (1) window and gate of every (pair, slot); half a warp per pair */
  for (int i0 = 0; i0 < np * 16; i0 += 256) {
    const int i = i0 + tid;
    const int p = i >> 4, sl = i & 15;
    bool w = false, l = false;
    if (i < np * 16 && sl != 1) {
      const int oi = t.small_oidx[sl], gs = rec[p].gs, ge = rec[p].ge;
      /* the three first points are unconditional (src/gc_integrators.c:175-177); afterwards the left point of a
       * pair needs idx >= start, the right one idx <= end (:190-197); tabulated range [start,end): src/type1.c:121 */
      w = (sl < 4) ? true : ((sl & 1) ? (oi <= ge) : (oi >= gs));
      if (w && oi >= gs && oi < ge) {
        const double r = t.small_r[sl];
        l = __dmul_rn(__fma_rn(rec[p].z, r, rec[p].sS), r) >= t.lnAcc1;
      }
    }
    const unsigned bw = __ballot_sync(T1_FULL, w), bl = __ballot_sync(T1_FULL, l);
    if (i < np * 16 && sl == 0) {
      const int sh = tid & 16;
      rec[p].wm = (bw >> sh) & 0xffffu;
      rec[p].lm = (bl >> sh) & 0xffffu;
    }
    if (l) items[atomicAdd(&cnt, 1)] = (unsigned short)i;
  }
  __syncthreads();
  /* (2) one thread per live point */
  const int nItems = cnt;
  for (int it = tid; it < nItems; it += 256) {
    const int i = items[it], p = i >> 4, sl = i & 15;
    const T1ARec &a = rec[p];
    const double r = t.small_r[sl];
    const double e = __dmul_rn(__fma_rn(a.z, r, a.sS), r);
    T1Point<LAB> pt;
    pt.live = true;
    pt.w = t.small_w[sl];
    pt.u = a.UL[sl];
    pt.ex = exp(e);
    t1_fill_point<LAB>(t, r, a.sS * r, pt);
    t1a_store_vals<LAB>(pt, a.Cc, vals + ((size_t)p * 16 + sl) * NQP, std::make_integer_sequence<int, NQ>{});
  }
  __syncthreads();
  /* (3) bookkeeping of the three first points and levels 0..3, one thread per (pair, quadrature) */
  for (int i = tid; i < np * NQ; i += 256) {
    const int p = i / NQ, q = i - p * NQ;
    const double *v = vals + (size_t)p * 16 * NQP + q;
    const unsigned wm = rec[p].wm;
    /* quadrature q of the loop nest "for N: for lambda = N, N-2, ..." (src/type1.c:132-146) -> offset of Q[N][lambda] */
    double *dst = rec[p].Qo + c_t1qoff[LAB][q];
#define T1A_V(s_) v[(s_) * NQP]
    const double A0 = T1A_V(0) + T1A_V(1), B0 = T1A_V(2) + T1A_V(3), C0 = T1A_V(4) + T1A_V(5), D0 = T1A_V(6) + T1A_V(7);
    const double A1 = T1A_V(8) + T1A_V(9), B1 = T1A_V(10) + T1A_V(11), C1 = T1A_V(12) + T1A_V(13), D1 = T1A_V(14) + T1A_V(15);
#undef T1A_V
    double P = A0, Qv = B0, I = P + Qv, res;
    bool done = false;
    I += C0;
    if (ecp_ps93_update(t.sm.levJ[0], t.sm.levN[0], __popc(wm & 0x30u), t.tolerance, I, &P, &Qv, &res)) done = true;
    if (!done) {
      I += D0;
      if (ecp_ps93_update(t.sm.levJ[1], t.sm.levN[1], __popc(wm & 0xc0u), t.tolerance, I, &P, &Qv, &res)) done = true;
    }
    if (!done) {
      I += (A1 + B1);
      if (ecp_ps93_update(t.sm.levJ[2], t.sm.levN[2], __popc(wm & 0x0f00u), t.tolerance, I, &P, &Qv, &res)) done = true;
    }
    if (!done) {
      I += (C1 + D1);
      if (ecp_ps93_update(t.sm.levJ[3], t.sm.levN[3], __popc(wm & 0xf000u), t.tolerance, I, &P, &Qv, &res)) done = true;
    }
    if (done) {
      *dst = res; /* T[l1][l2] += I  (src/type1.c:143) */
    } else {
      double *st = state + ((size_t)(g0 + p) * NQ + q) * 3;
      st[0] = I;
      st[1] = P;
      st[2] = Qv;
      if (q < 32)
        atomicOr(&rec[p].open0, 1u << q);
      else
        atomicOr(&rec[p].open1, 1u << (q - 32));
    }
  }
  __syncthreads();
  /* (4) pairs with an open quadrature continue level-wise in k_type1S */
  if (tid < np && (rec[tid].open0 | rec[tid].open1)) {
    survList[atomicAdd(survCount, 1)] = g0 + tid;
    survMask[rec[tid].pr] = (unsigned long long)rec[tid].open0 | ((unsigned long long)rec[tid].open1 << 32);
  }
}

/* ---------------------------------------------------------------------------------------------- */
template <int LAB>
__global__ void __launch_bounds__(128, (LAB <= 3 ? T1_MINB_LO : (LAB <= 6 ? T1_MINB_MID : 3))) k_type1S(DevT t, DevB b, T1Segs segs, int *workCtr, int *failCount, int *failList,
                                                unsigned long long *failMask, const int *survCount, const int *survList,
                                                const unsigned long long *survMask, const double *state) {
  using Cfg = T1Cfg<LAB>;
  constexpr int NQ = Cfg::NQ, NQL = Cfg::NQL;
  extern __shared__ __align__(16) double t1_red[];
  const int lane = threadIdx.x & 31, gl = lane & 7, gbase = lane & 24;
  double *red = t1_red + ((threadIdx.x >> 5) * 4 + (lane >> 3)) * Cfg::GS;
  const unsigned long long dbgT0 = b.dbg ? ecp_gtimer() : 0;
  /* survCount != NULL: the pairs k_type1A left open, continued from level 4 with its (I, p, q); NULL: every pair of the
   * launch from its first point (LIBECP_B200_T1=legacy) */
  const int total = survCount ? *survCount : (int)segs.prefix[segs.nseg];
  bool have = false, drained = false;
  double z = 0.0, sS = 0.0, Cc = 0.0;
  int gs = 0, ge = 0;
  long long pr = 0;
  const double *UL = t.typeUL;
  double *Qo = b.Q;
  double I[NQL], P[NQL], Qv[NQL];
  int qo[NQL]; /* offset of Q[N][lambda] of the quadratures this lane owns (q = 8k + lane) */
#pragma unroll
  for (int k = 0; k < NQL; k++) {
    I[k] = P[k] = Qv[k] = 0.0;
    const int q = (8 * k + gl < NQ) ? 8 * k + gl : NQ - 1;
    qo[k] = t1_qN(q) * (LAB + 1) + t1_qLam(q);
  }
  unsigned open = 0;
#ifdef T1_STATS
  int nchunks = 0;
#endif
  int c = 0;   /* 0, 1: the two fixed chunks (slots 0..15 = first points and levels 0..3); 2: level-wise     */
  int v = 4;   /* level being accumulated once c == 2                                                        */
  int ks = 0;  /* 8-point step inside that level: only the points inside the window are visited              */
  T1Level lv = {0, 0, 0, 0, 0, 0};
  for (;;) {
    /* ---- a group without a pair takes the next one of this launch ---- */
    const bool need = !have && !drained;
    if (__any_sync(T1_FULL, need)) {
      int g = 0;
      if (need && gl == 0) g = atomicAdd(workCtr, 1);
      g = __shfl_sync(T1_FULL, g, gbase);
      if (need) {
        if (g < total) {
          if (survCount) g = survList[g];
          int sg = 0;
          while (segs.prefix[sg + 1] <= g) sg++;
          pr = segs.start[sg] + (g - segs.prefix[sg]);
          const T1Rec rec = b.t1rec[pr];
          z = rec.z;
          sS = rec.sS;
          Cc = rec.CcS;
          gs = rec.gs;
          ge = rec.ge;
          UL = t.typeUL + (size_t)rec.type * ECP_SMALL_SLOTS;
          Qo = b.Q + rec.qoff;
          open = 0;
          c = 0;
          if (survCount) {
            const unsigned long long om = survMask[pr];
#pragma unroll
            for (int k = 0; k < NQL; k++) {
              const int q = 8 * k + gl;
              if (q < NQ && (om >> q & 1ull)) {
                const double *st = state + ((size_t)g * NQ + q) * 3;
                open |= 1u << k;
                I[k] = st[0];
                P[k] = st[1];
                Qv[k] = st[2];
              }
            }
            c = 2;
          } else {
#pragma unroll
            for (int k = 0; k < NQL; k++)
              if (8 * k + gl < NQ) open |= 1u << k;
          }
          v = 4;
          ks = 0;
          lv = t1_level(&t.sm, t.small_jL, t.small_jR, 4, gs, ge);
          have = true;
        } else {
          drained = true;
        }
      }
    }
    if (!__any_sync(T1_FULL, have)) break;
    /* ---- one chunk: every lane tabulates one slot; from level 4 on the 8 lanes take the next 8 points of the level
     * that lie inside the window (about 30 % of a level on Au20: the pairs that never converge walked 48 chunks of
     * mostly idle lanes) ---- */
    int slot = 8 * c + gl;
    if (c >= 2) {
      const int m = 8 * ks + gl;
      slot = (m < lv.nLl) ? lv.s0 + 2 * (lv.jLa + m) : ((m < lv.nLive) ? lv.s0 + 2 * (lv.jRa + m - lv.nLl) + 1 : 1);
    }
    T1Point<LAB> pt;
    pt.live = false;
    pt.w = 0.0;
    bool inWin = false;
    if (have && slot != 1) {
      const int oi = t.small_oidx[slot];
      /* the three first points are unconditional (src/gc_integrators.c:175-177); afterwards the left point of a
       * pair needs idx >= start, the right one idx <= end (:190-197) */
      inWin = (slot < 4) ? true : ((slot & 1) ? (oi <= ge) : (oi >= gs));
      if (inWin && oi >= gs && oi < ge) { /* tabulated range [start,end): src/type1.c:121 */
        const double r = t.small_r[slot];
        const double e = (z * r + sS) * r;
        if (e >= t.lnAcc1) {
          pt.live = true;
          pt.w = t.small_w[slot];
          pt.u = UL[slot];
          pt.ex = exp(e);
          t1_fill_point<LAB>(t, r, sS * r, pt);
        }
      }
    }
    const unsigned bal = (__ballot_sync(T1_FULL, inWin) >> gbase) & 0xffu;
#ifdef T1_STATS
    {
      const unsigned bw = __ballot_sync(T1_FULL, inWin), bl = __ballot_sync(T1_FULL, pt.live);
      const unsigned bh = __ballot_sync(T1_FULL, have && gl == 0);
      const unsigned b0 = __ballot_sync(T1_FULL, have && gl == 0 && c == 0), b1 = __ballot_sync(T1_FULL, have && gl == 0 && c == 1);
      const unsigned l0 = __ballot_sync(T1_FULL, pt.live && c == 0), l1 = __ballot_sync(T1_FULL, pt.live && c == 1);
      if (lane == 0) {
        T1_STAT(0, 1);
        T1_STAT(1, __popc(bh));
        T1_STAT(2, __popc(b0));
        T1_STAT(3, __popc(b1));
        T1_STAT(4, __popc(bh) - __popc(b0) - __popc(b1));
        T1_STAT(5, __popc(bw));
        T1_STAT(6, __popc(bl));
        T1_STAT(7, __popc(l0));
        T1_STAT(8, __popc(l1));
        T1_STAT(9, __popc(bl) - __popc(l0) - __popc(l1));
      }
      if (have) nchunks++;
    }
#endif
    /* ---- products of all quadratures -> tile [q][lane]; owner lane q mod 8 reads its rows back ---- */
    __syncwarp();
    t1_store_vals<LAB>(pt, Cc, red + gl, std::make_integer_sequence<int, NQ>{});
    __syncwarp();
    if (have) {
      const bool last = (c >= 2) && (8 * ks + 8 >= lv.nLive);
      const int cnt = lv.cnt;
#pragma unroll
      for (int k = 0; k < NQL; k++) {
        const int q = 8 * k + gl;
        const double *row = red + (q < NQ ? q : NQ - 1) * Cfg::ROW;
        const double A = row[0] + row[1], B = row[2] + row[3], Cq = row[4] + row[5], D = row[6] + row[7];
        double *dst = Qo + qo[k];
        double res;
        if (c == 0) {
          /* slot 0 = centre (p), slots 2,3 = first pair (q), slots 4,5 = level 0, slots 6,7 = level 1 */
          P[k] = A;
          Qv[k] = B;
          I[k] = P[k] + Qv[k];
          I[k] += Cq;
          if ((open >> k & 1) &&
              ecp_ps93_update(t.sm.levJ[0], t.sm.levN[0], __popc(bal & 0x30u), t.tolerance, I[k], &P[k], &Qv[k], &res)) {
            *dst = res; /* T[l1][l2] += I  (src/type1.c:143) */
            open &= ~(1u << k);
          }
          I[k] += D;
          if ((open >> k & 1) &&
              ecp_ps93_update(t.sm.levJ[1], t.sm.levN[1], __popc(bal & 0xc0u), t.tolerance, I[k], &P[k], &Qv[k], &res)) {
            *dst = res;
            open &= ~(1u << k);
          }
        } else if (c == 1) {
          /* slots 8..11 = level 2, slots 12..15 = level 3 */
          I[k] += (A + B);
          if ((open >> k & 1) &&
              ecp_ps93_update(t.sm.levJ[2], t.sm.levN[2], __popc(bal & 0x0fu), t.tolerance, I[k], &P[k], &Qv[k], &res)) {
            *dst = res;
            open &= ~(1u << k);
          }
          I[k] += (Cq + D);
          if ((open >> k & 1) &&
              ecp_ps93_update(t.sm.levJ[3], t.sm.levN[3], __popc(bal & 0xf0u), t.tolerance, I[k], &P[k], &Qv[k], &res)) {
            *dst = res;
            open &= ~(1u << k);
          }
        } else {
          I[k] += ((A + B) + (Cq + D));
          /* a level without in-window point only moves the bookkeeping, as in the reference (cnt == 0,
           * src/gc_integrators.c:203-208) */
          if (last && (open >> k & 1) &&
              ecp_ps93_update(t.sm.levJ[v], t.sm.levN[v], cnt, t.tolerance, I[k], &P[k], &Qv[k], &res)) {
            *dst = res;
            open &= ~(1u << k);
          }
        }
      }
      if (c < 2) {
        c++;
      } else if (last) {
        v++;
        ks = 0;
        if (v < ECP_SMALL_LEVELS) lv = t1_level(&t.sm, t.small_jL, t.small_jR, v, gs, ge);
      } else {
        ks++;
      }
    }
    /* ---- group finished: all its quadratures converged, or the grid is exhausted ---- */
    const unsigned ob = (__ballot_sync(T1_FULL, have && open != 0) >> gbase) & 0xffu;
    const bool fin = have && (ob == 0 || v == ECP_SMALL_LEVELS);
    if (__any_sync(T1_FULL, fin && ob != 0)) {
      /* quadratures that never converged on the small grid -> large grid (src/type1.c:149) */
      unsigned long long m = 0;
      if (fin) {
#pragma unroll
        for (int k = 0; k < NQL; k++)
          if (open >> k & 1) m |= 1ull << (8 * k + gl);
      }
      m |= __shfl_xor_sync(T1_FULL, m, 1);
      m |= __shfl_xor_sync(T1_FULL, m, 2);
      m |= __shfl_xor_sync(T1_FULL, m, 4);
      if (fin && m && gl == 0) {
        failMask[pr] = m;
        failList[atomicAdd(failCount, 1)] = (int)pr;
      }
    }
#ifdef T1_STATS
    if (fin && gl == 0) {
      T1_STAT(10, 1);
      if (ob != 0) {
        T1_STAT(11, 1);
        T1_STAT(12, nchunks);
      } else {
        T1_STAT(13, nchunks);
      }
    }
    if (fin) nchunks = 0;
#endif
    if (fin) {
      have = false;
      open = 0;
    }
  }
  if (b.dbg && threadIdx.x == 0 && blockIdx.x < DBG_STRIDE) {
    unsigned long long *e = b.dbg + (size_t)(1 + LAB) * DBG_STRIDE * 2;
    e[2 * blockIdx.x] = dbgT0;
    e[2 * blockIdx.x + 1] = ecp_gtimer();
  }
}

/* ---------------------------------------------------------------------------------------------- */
template <int LAB>
__global__ void __launch_bounds__(128, (LAB <= 3 ? T1_MINB_LO : (LAB <= 6 ? T1_MINB_MID : 3))) k_type1L(DevT t, DevB b, int *workCtr, const int *failCount, const int *failList,
                                                const unsigned long long *failMask, int *errFlag) {
  using Cfg = T1Cfg<LAB>;
  constexpr int NQ = Cfg::NQ, NQL = Cfg::NQL;
  extern __shared__ __align__(16) double t1_red[];
  const int lane = threadIdx.x & 31, gl = lane & 7, gbase = lane & 24;
  double *red = t1_red + ((threadIdx.x >> 5) * 4 + (lane >> 3)) * Cfg::GS;
  const int total = *failCount;
  bool have = false, drained = false;
  double z = 0.0, sS = 0.0, zd2 = 0.0, Cc = 0.0, i1 = 0.0, i2 = 0.0;
  int Lc = 0, g0 = 0, g1 = 0;
  double *Qo = b.Q;
  double I[NQL], P[NQL], Qv[NQL];
  int qo[NQL];
#pragma unroll
  for (int k = 0; k < NQL; k++) {
    I[k] = P[k] = Qv[k] = 0.0;
    const int q = (8 * k + gl < NQ) ? 8 * k + gl : NQ - 1;
    qo[k] = t1_qN(q) * (LAB + 1) + t1_qLam(q);
  }
  unsigned open = 0;
  int c = 0;          /* 0: the fixed first chunk (slots 0..7 = centre, levels 1 and 2); 1: level-wise            */
  int lev = 3, n = 7; /* level accumulated once c == 1: slots [2^lev, 2^(lev+1)); n = points before it          */
  int ks = 0;         /* 8-candidate step inside the level (lg_level: only points that can pass the gate)        */
  LgRange lr = {1, 0};
  LgLevel lv = {0, 0, 0, 0};
  for (;;) {
    const bool need = !have && !drained;
    if (__any_sync(T1_FULL, need)) {
      int g = 0;
      if (need && gl == 0) g = atomicAdd(workCtr, 1);
      g = __shfl_sync(T1_FULL, g, gbase);
      if (need) {
        if (g < total) {
          const long long pr = failList[g];
          const T1Rec rec = b.t1rec[pr];
          const unsigned long long fm = failMask[pr];
          z = rec.z;
          sS = rec.sS;
          zd2 = rec.zd2;
          Cc = rec.CcL;
          i1 = rec.i1;
          i2 = rec.i2;
          Lc = t.typeL[rec.type];
          g0 = t.typeGaussOff[rec.type];
          g1 = t.typeGaussOff[rec.type + 1];
          Qo = b.Q + rec.qoff;
          open = 0;
#pragma unroll
          for (int k = 0; k < NQL; k++)
            if (fm >> (8 * k + gl) & 1) open |= 1u << k;
          c = 0;
          lev = 3;
          n = 7;
          ks = 0;
          lr = lg_live_range(t.large_xo, t.largeOrder, z, sS, zd2 - t.lnAcc1, i1, i2);
          lv = lg_level(t.largeSlots, t.largeOrder, lr, 3);
          have = true;
        } else {
          drained = true;
        }
      }
    }
    if (!__any_sync(T1_FULL, have)) break;
    const int slot = (c == 0) ? gl : lg_slot(lv, lev, 8 * ks + gl);
    T1Point<LAB> pt;
    pt.live = false;
    pt.w = 0.0;
    if (have && slot != 1) {
      const double r = i1 * t.large_x[slot] + i2;   /* src/gc_integrators.c:326-329 */
      const double e = (z * r + sS) * r + zd2;      /* src/type1.c:162 */
      if (e >= t.lnAcc1) {
        pt.live = true;
        pt.w = t.large_w[slot] * i1;
        pt.u = ecp_pot_eval(t.gaussL, t.gaussN, t.gaussD, t.gaussA, g0, g1, Lc, r);
        pt.ex = exp(e);
        t1_fill_point<LAB>(t, r, sS * r, pt);
      }
    }
    __syncwarp();
    t1_store_vals<LAB>(pt, Cc, red + gl, std::make_integer_sequence<int, NQ>{});
    __syncwarp();
    if (have) {
      const bool first = (ks == 0);
      const bool last = (8 * ks + 8 >= lv.nLive);
#pragma unroll
      for (int k = 0; k < NQL; k++) {
        const int q = 8 * k + gl;
        const double *row = red + (q < NQ ? q : NQ - 1) * Cfg::ROW;
        const double A = row[0] + row[1], B = row[2] + row[3], Cq = row[4] + row[5], D = row[6] + row[7];
        double *dst = Qo + qo[k];
        double res;
        if (c == 0) {
          /* slot 0 = centre, slots 2,3 = level 1, slots 4..7 = level 2 */
          I[k] = A; /* I = w[M] f(M); p = I  (src/gc_integrators.c:49-52) */
          P[k] = I[k];
          Qv[k] = 2 * P[k];
          P[k] = 2 * I[k];
          I[k] += B;
          if ((open >> k & 1) && ecp_psm92_update(3, 1, t.tolerance, I[k], P[k], Qv[k], &res)) {
            *dst = res; /* T[l1][l2] += grid->I  (src/type1.c:193) */
            open &= ~(1u << k);
          }
          Qv[k] = 2 * P[k];
          P[k] = 2 * I[k];
          I[k] += (Cq + D);
          if ((open >> k & 1) && ecp_psm92_update(7, 1, t.tolerance, I[k], P[k], Qv[k], &res)) {
            *dst = res;
            open &= ~(1u << k);
          }
        } else {
          if (first) { /* q = 2p; p = 2I  (src/gc_integrators.c:56-57) */
            Qv[k] = 2 * P[k];
            P[k] = 2 * I[k];
          }
          I[k] += ((A + B) + (Cq + D));
          if (last && (open >> k & 1) && ecp_psm92_update(2 * n + 1, 1, t.tolerance, I[k], P[k], Qv[k], &res)) {
            *dst = res;
            open &= ~(1u << k);
          }
        }
      }
      if (c == 0) {
        c = 1;
      } else if (last) {
        n = 2 * n + 1;
        lev++;
        ks = 0;
        if (lev <= t.largeLevels) lv = lg_level(t.largeSlots, t.largeOrder, lr, lev);
      } else {
        ks++;
      }
    }
    const unsigned ob = (__ballot_sync(T1_FULL, have && open != 0) >> gbase) & 0xffu;
    const bool fin = have && (ob == 0 || (c == 1 && lev > t.largeLevels));
    if (fin) {
      if (ob != 0 && gl == 0) atomicExch(errFlag, 1); /* large grid failed: rc 1 (src/libecp.h:25) */
      have = false;
      open = 0;
    }
  }
}

#endif
