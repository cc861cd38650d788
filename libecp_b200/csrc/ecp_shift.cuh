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

__global__ void k_shiftJ(DevT t, DevB b, const long long *__restrict__ clsJ, long long nElem, double *__restrict__ Jbuf,
                         int fuse) {
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
    if (fuse) { /* experimental (LIBECP_B200_SHIFT=fused, matrix-only runs): the shift is linear, so 4 pi chi + 16 pi^2 gamma
                 * is shifted once instead of chi and gamma separately */
      const double n1 = 4.0 * M_PI;
      J1 = fma(f, fma(n1 * n1, G2[p], n1 * G1[p]), J1);
    } else {
      J1 = fma(f, G1[p], J1);
      J2 = fma(f, G2[p], J2);
    }
  }
  double *J = Jbuf + 2 * (clsJ[c] + (long long)(tri - b.clsFirst[c]) * (na * cdb));
  J[rem] = J1;
  if (!fuse) J[na * cdb + rem] = J2;
}

__global__ void k_shiftI(DevT t, DevB b, const long long *__restrict__ clsJ, long long nElem,
                         const double *__restrict__ Jbuf, int flags, int fuse) {
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
    if (fuse) {
      I1 = fma(f, J1[p], I1); /* the factors 4 pi / 16 pi^2 went into J (k_shiftJ) */
    } else {
      I1 = fma(f * n1, J1[p], I1); /* factor *= N; I += factor * J  (src/util.c:321-324) */
      I2 = fma(f * n2, J2[p], I2);
    }
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

#endif
