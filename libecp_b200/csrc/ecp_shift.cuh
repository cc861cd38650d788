/* ecp_shift.cuh - binomial shift of the C-centred monomial tables chi / gamma to A- and B-centred Cartesian
 * functions, normalisation (4 pi for type 1, 16 pi^2 for type 2) and output.
 *
 * Replaces calcPolynomials (reference src/util.c:246-334) and the scatter of libECP_callback0
 * (src/getIntegrals.c:22-43).  Same two passes as the reference, in the same order:
 *   k_shiftJ : J[c1][q]  = sum_{alpha <= a(c1)} binom * usp_A[a-alpha] * G[idx(alpha)][q]     (src/util.c:270-299)
 *   k_shiftI : I[c1][c2] = sum_{beta <= b(c2)} (binom * usp_B[b-beta] * N) * J[c1][idx(beta)]  (src/util.c:302-329)
 * Both integral types are carried through the same loops (shared factors).  The term lists (sub-monomial,
 * exponent difference, binomial product) are geometry independent and come from the host (tables.c); terms with
 * |factor| <= accuracy are skipped exactly as the reference does (src/util.c:286,318).
 * One thread per J / I element; lanes of a warp run over q (pass 1: identical trip counts) or c2 (pass 2).
 */
#ifndef ECP_SHIFT_CUH
#define ECP_SHIFT_CUH

__global__ void k_shiftJ(DevT t, DevB b, const long long *__restrict__ clsJ, long long nElem, double *__restrict__ Jbuf) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nElem) return;
  const int c = find_class(clsJ, t.nClasses, w);
  const int la = t.clsLa[c], lb = t.clsLb[c];
  const int na = ecp_ijk(la), cdb = ecp_cd(lb);
  const long long idx = w - clsJ[c];
  const int tri = b.clsFirst[c] + (int)(idx / (na * cdb));
  const int rem = (int)(idx % (na * cdb)), c1 = rem / cdb, q = rem % cdb;
  const int asa = b.trirec[tri].asa;
  const int dA = b.trirec[tri].dA;
  const double *uA = b.uspX + (size_t)asa * USPX_STRIDE;
  const long long gOff = tri_G_off(t, b, c, tri);
  const double *G1 = b.chi + gOff + q, *G2 = b.gamma + gOff + q;
  const int k0 = t.shTermOff[la * t.shOffStride + c1], k1 = t.shTermOff[la * t.shOffStride + c1 + 1];
  double J1 = 0.0, J2 = 0.0;
  for (int k = k0; k < k1; k++) {
    const int dd = t.shTermD[k];
    const double f = t.shTermBin[k] * uA[(dd & 15) * dA * dA + ((dd >> 4) & 15) * dA + (dd >> 8)];
    if (fabs(f) <= t.accuracy) continue; /* src/util.c:286 */
    const int p = t.shTermP[k] * cdb;
    J1 = fma(f, G1[p], J1);
    J2 = fma(f, G2[p], J2);
  }
  double *J = Jbuf + 2 * (clsJ[c] + (long long)(tri - b.clsFirst[c]) * (na * cdb));
  J[rem] = J1;
  J[na * cdb + rem] = J2;
}

__global__ void k_shiftI(DevT t, DevB b, const long long *__restrict__ clsJ, long long nElem,
                         const double *__restrict__ Jbuf, int flags) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nElem) return;
  const int c = find_class(b.clsOutElem, t.nClasses, w);
  const int la = t.clsLa[c], lb = t.clsLb[c];
  const int na = ecp_ijk(la), nb = ecp_ijk(lb), cdb = ecp_cd(lb);
  const long long idx = w - b.clsOutElem[c];
  const int tri = b.clsFirst[c] + (int)(idx / (na * nb));
  const int cc = (int)(idx % (na * nb)), c1 = cc / nb, c2 = cc % nb;
  const TriRec rec = b.trirec[tri];
  const int asb = rec.asb;
  const int dB = rec.dB;
  const double *uB = b.uspX + (size_t)asb * USPX_STRIDE;
  const double *J1 = Jbuf + 2 * (clsJ[c] + (long long)(tri - b.clsFirst[c]) * (na * cdb)) + c1 * cdb;
  const double *J2 = J1 + na * cdb;
  const double n1 = 4.0 * M_PI, n2 = n1 * n1; /* src/libecp.c:234-235 */
  const int k0 = t.shTermOff[lb * t.shOffStride + c2], k1 = t.shTermOff[lb * t.shOffStride + c2 + 1];
  double I1 = 0.0, I2 = 0.0;
  for (int k = k0; k < k1; k++) {
    const int dd = t.shTermD[k];
    const double f = t.shTermBin[k] * uB[(dd & 15) * dB * dB + ((dd >> 4) & 15) * dB + (dd >> 8)];
    if (fabs(f) <= t.accuracy) continue; /* src/util.c:318 */
    const int p = t.shTermP[k];
    I1 = fma(f * n1, J1[p], I1); /* factor *= N; I += factor * J  (src/util.c:321-324) */
    I2 = fma(f * n2, J2[p], I2);
  }
  if (flags & 2) {
    double *o = b.blocks + b.trOut[tri];
    o[cc] = I1;
    o[na * nb + cc] = I2;
  }
  if (flags & 1) {
    const int row = rec.rowAO + c1, col = rec.colAO + c2;
    if (row <= col) atomicAdd(&b.matrix[(size_t)row * t.nAO + col], I1 + I2); /* src/getIntegrals.c:38-40 */
  }
}

/* ----------------------------------------------------------------------------------------------
 * k_shift2: both passes in one kernel, a block per chunk of `tpb` consecutive triples of one class (default;
 * LIBECP_B200_SHIFT=two keeps k_shiftJ / k_shiftI).
 * ncu on the two-kernel form (profiles/r2): 5-9 % of the issue slots FP64, ~10 warps per issue waiting for global loads
 * (L2 hit rate 26-35 %), 150 MB of DRAM traffic per 0.1 ms launch - every element searched its class, divided a 64-bit
 * index, chased its triple record and its unit-sphere factors, and J made a round trip through HBM.  Here
 *   - the class (and with it every size) is uniform over the block; element -> (component, column) maps and the term
 *     lists of the two shells are staged once per block;
 *   - chi / gamma of the chunk arrive as one contiguous copy; the factors binom * usp_X[a - alpha] are formed once per
 *     (triple, term) - not once per element - with the reference's skip of |factor| <= accuracy stored as an exact zero
 *     (a zero factor adds nothing: same result as skipping, src/util.c:286,318);
 *   - J stays in shared memory.
 * Same terms in the same order with the same fma's as k_shiftJ / k_shiftI: bit-identical blocks. */
#define SHIFT2_CH 8 /* at most so many chunks of tpb triples per block (the per-block tables are paid once per block) */
struct Shift2Layout { /* offsets (doubles) inside the dynamic shared memory of one block */
  int G1, G2, J1, J2, fA, fB, ints;
};
__host__ __device__ inline int shift2_tpb(int la, int lb) {
  const int nJ = ((la + 1) * (la + 2) / 2) * ((lb + 1) * (lb + 2) * (lb + 3) / 6);
  int tpb = 512 / nJ;
  return tpb < 1 ? 1 : (tpb > 128 ? 128 : tpb);
}
__host__ __device__ inline size_t shift2_smem(int la, int lb, int TA, int TB, int tpb, Shift2Layout *L) {
  const int na = (la + 1) * (la + 2) / 2, nb = (lb + 1) * (lb + 2) / 2;
  const int cda = (la + 1) * (la + 2) * (la + 3) / 6, cdb = (lb + 1) * (lb + 2) * (lb + 3) / 6;
  const int E = cda * cdb, nJ = na * cdb, nI = na * nb;
  int o = 0;
  L->G1 = o; o += tpb * E;
  L->G2 = o; o += tpb * E;
  L->J1 = o; o += tpb * nJ;
  L->J2 = o; o += tpb * nJ;
  L->fA = o; o += tpb * TA;
  L->fB = o; o += tpb * TB;
  L->ints = o;
  /* ints: pA[TA] pB[TB] dA_[TA] dB_[TB] offA[na+1] offB[nb+1] jmap[nJ] imap[nI] + 8 per triple */
  const int nints = 2 * TA + 2 * TB + na + 1 + nb + 1 + nJ + nI + 8 * tpb;
  return (size_t)o * sizeof(double) + (size_t)nints * sizeof(int);
}
__global__ void __launch_bounds__(128) k_shift2(DevT t, DevB b, const int *__restrict__ clsBlk, int flags) {
  extern __shared__ __align__(16) double sh_sm[];
  /* class of this block */
  int c = 0;
  {
    int lo = 0, hi = t.nClasses - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (clsBlk[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    c = lo;
  }
  const int la = t.clsLa[c], lb = t.clsLb[c];
  const int na = ecp_ijk(la), nb = ecp_ijk(lb), cda = ecp_cd(la), cdb = ecp_cd(lb);
  const int E = cda * cdb, nJ = na * cdb, nI = na * nb;
  const int os = t.shOffStride;
  const int kA0 = t.shTermOff[la * os], TA = t.shTermOff[la * os + na] - kA0;
  const int kB0 = t.shTermOff[lb * os], TB = t.shTermOff[lb * os + nb] - kB0;
  const int tpb = shift2_tpb(la, lb);
  Shift2Layout L;
  shift2_smem(la, lb, TA, TB, tpb, &L);
  double *G1 = sh_sm + L.G1, *G2 = sh_sm + L.G2, *J1 = sh_sm + L.J1, *J2 = sh_sm + L.J2, *fA = sh_sm + L.fA, *fB = sh_sm + L.fB;
  int *pA = (int *)(sh_sm + L.ints), *pB = pA + TA, *dAt = pB + TB, *dBt = dAt + TA, *offA = dBt + TB, *offB = offA + na + 1;
  int *jmap = offB + nb + 1, *imap = jmap + nJ, *tinfo = imap + nI; /* tinfo[8 * tl]: asa asb dA dB rowAO colAO - - */
  const int first = b.clsFirst[c], nTri = b.clsFirst[c + 1] - first;
  const int nCh = clsBlk[t.nClasses + 1 + c];                 /* chunks per block of this class */
  const int ltB = ((int)blockIdx.x - clsBlk[c]) * tpb * nCh; /* first triple of the block */
  const int tid = threadIdx.x;
  for (int k = tid; k < TA; k += 128) {
    pA[k] = t.shTermP[kA0 + k] * cdb;
    dAt[k] = t.shTermD[kA0 + k];
  }
  for (int k = tid; k < TB; k += 128) {
    pB[k] = t.shTermP[kB0 + k];
    dBt[k] = t.shTermD[kB0 + k];
  }
  for (int k = tid; k <= na; k += 128) offA[k] = t.shTermOff[la * os + k] - kA0;
  for (int k = tid; k <= nb; k += 128) offB[k] = t.shTermOff[lb * os + k] - kB0;
  for (int r = tid; r < nJ; r += 128) jmap[r] = (r / cdb) | ((r % cdb) << 16);
  for (int r = tid; r < nI; r += 128) imap[r] = (r / nb) | ((r % nb) << 16);
  const double n1 = 4.0 * M_PI, n2 = n1 * n1; /* src/libecp.c:234-235 */
  for (int ch = 0; ch < nCh; ch++) {
    const int lt0 = ltB + ch * tpb, nT = min(tpb, nTri - lt0);
    if (nT <= 0) break;
    __syncthreads(); /* tables ready / the previous chunk's readers are through */
    for (int tl = tid; tl < nT; tl += 128) {
      const TriRec rec = b.trirec[first + lt0 + tl];
      int *ti = tinfo + 8 * tl;
      ti[0] = rec.asa; ti[1] = rec.asb; ti[2] = rec.dA; ti[3] = rec.dB; ti[4] = rec.rowAO; ti[5] = rec.colAO;
    }
    { /* chi / gamma of the chunk: contiguous */
      const long long g0 = b.clsElem[c] + (long long)lt0 * E;
      for (int i = tid; i < nT * E; i += 128) {
        G1[i] = b.chi[g0 + i];
        G2[i] = b.gamma[g0 + i];
      }
    }
    __syncthreads();
    /* factors binom * usp[a - alpha] per (triple, term); |f| <= accuracy -> 0 */
    for (int w = tid; w < nT * TA; w += 128) {
      const int tl = w / TA, k = w - tl * TA;
      const int *ti = tinfo + 8 * tl;
      const int dd = dAt[k], dA = ti[2];
      const double f = t.shTermBin[kA0 + k] * b.uspX[(size_t)ti[0] * USPX_STRIDE + (dd & 15) * dA * dA + ((dd >> 4) & 15) * dA + (dd >> 8)];
      fA[w] = fabs(f) <= t.accuracy ? 0.0 : f;
    }
    for (int w = tid; w < nT * TB; w += 128) {
      const int tl = w / TB, k = w - tl * TB;
      const int *ti = tinfo + 8 * tl;
      const int dd = dBt[k], dB = ti[3];
      const double f = t.shTermBin[kB0 + k] * b.uspX[(size_t)ti[1] * USPX_STRIDE + (dd & 15) * dB * dB + ((dd >> 4) & 15) * dB + (dd >> 8)];
      fB[w] = fabs(f) <= t.accuracy ? 0.0 : f;
    }
    __syncthreads();
    /* pass 1: J[c1][q] = sum_{alpha <= a(c1)} f G[idx(alpha)][q]   (src/util.c:270-299) */
    for (int w = tid; w < nT * nJ; w += 128) {
      const int tl = w / nJ, r = w - tl * nJ;
      const double *g1 = G1 + tl * E, *g2 = G2 + tl * E, *f = fA + tl * TA;
      const int c1 = jmap[r] & 0xffff, q = jmap[r] >> 16;
      double a1 = 0.0, a2 = 0.0;
      for (int k = offA[c1]; k < offA[c1 + 1]; k++) {
        const double fk = f[k];
        a1 = fma(fk, g1[pA[k] + q], a1);
        a2 = fma(fk, g2[pA[k] + q], a2);
      }
      J1[w] = a1;
      J2[w] = a2;
    }
    __syncthreads();
    /* pass 2: I[c1][c2] = sum_{beta <= b(c2)} (f N) J[c1][idx(beta)]   (src/util.c:302-329), N = 4 pi / 16 pi^2 */
    for (int w = tid; w < nT * nI; w += 128) {
      const int tl = w / nI, r = w - tl * nI;
      const int c1 = imap[r] & 0xffff, c2 = imap[r] >> 16;
      const double *j1 = J1 + tl * nJ + c1 * cdb, *j2 = J2 + tl * nJ + c1 * cdb, *f = fB + tl * TB;
      double I1 = 0.0, I2 = 0.0;
      for (int k = offB[c2]; k < offB[c2 + 1]; k++) {
        const double fk = f[k];
        I1 = fma(fk * n1, j1[pB[k]], I1); /* factor *= N; I += factor * J  (src/util.c:321-324) */
        I2 = fma(fk * n2, j2[pB[k]], I2);
      }
      const int *ti = tinfo + 8 * tl;
      if (flags & 2) {
        double *o = b.blocks + b.trOut[first + lt0 + tl];
        o[r] = I1;
        o[nI + r] = I2;
      }
      if (flags & 1) {
        const int row = ti[4] + c1, col = ti[5] + c2;
        if (row <= col) atomicAdd(&b.matrix[(size_t)row * t.nAO + col], I1 + I2); /* src/getIntegrals.c:38-40 */
      }
    }
  }
}

#endif
