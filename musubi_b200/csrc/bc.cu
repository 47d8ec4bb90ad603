// bc.cu -- boundary kernels, applied link-wise on the post-collision state
// before the now/next swap (set_boundary, mus/source/bc/mus_bc_general_module.fpp:179-285).
//
//   fill_bcBuffer        mus_bc_general_module.fpp:1726-1768
//   velocity_bounceback  mus_bc_fluid_module.fpp:1503-1597   (and _incomp :1401-1490)
//   fill_neighBuffer     mus_bc_general_module.fpp:1589-1717
//   pressure_expol       mus_bc_fluid_module.fpp:1165-1362
//   pressure_antiBounceBack  mus_bc_fluid_module.fpp:2161-2353
//   wall                 do_nothing (mus_bc_fluid_wall_module.fpp:407-450): the
//                        bounce-back lives in the neighbour list, nothing to launch.
//
// The two-phase structure of the reference is kept: phase 1 snapshots the QQ
// post-collision PDFs of every boundary element (bcBuffer, AOS), phase 2 writes
// state(links(l)) from the snapshot, so links of different boundaries never see
// each other's writes.  Both phases touch only the boundary surface.
#include "kernels.cuh"

namespace musb200 {

__global__ void fillBcBufferKernel(const double *__restrict__ state, long long S, int QQ,
                                   const int32_t *__restrict__ bcElems,
                                   const int32_t *__restrict__ needed, int nNeeded,
                                   double *__restrict__ bcBuffer) {
  // thread -> (slot, direction) with the slot index fastest: reads of one direction row
  // are as contiguous as the boundary surface allows
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nNeeded * QQ) return;
  const int q = i / nNeeded, k = i % nNeeded;
  const int slot = needed[k] - 1;
  const int e = bcElems[slot] - 1;
  bcBuffer[(long long)slot * QQ + q] = state[(long long)q * S + e];
}

template <int QQ>
__global__ void velocityBounceBackKernel(int incomp, double *__restrict__ state, long long S,
                                         const double *__restrict__ bcBuffer, int nLinks,
                                         const int32_t *__restrict__ links,
                                         const int32_t *__restrict__ outPos,
                                         const int32_t *__restrict__ posInBuffer,
                                         const int32_t *__restrict__ iDir,
                                         const double *__restrict__ velLat) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nLinks) return;
  const double fOut = bcBuffer[outPos[l] - 1];
  const int pib = posInBuffer[l] - 1;
  double rho = 0.0;
#pragma unroll
  for (int q = 0; q < QQ; ++q) rho = rho + bcBuffer[(long long)pib * QQ + q];
  if (incomp) rho = 1.0;  // rho0
  const int d = iDir[l] - 1;
  // weight and cxDir of a run-time direction: small switch-free lookup
  int c0 = 0, c1 = 0, c2 = 0;
  double w = 0.0;
#pragma unroll
  for (int q = 0; q < QQ - 1; ++q)
    if (q == d) { c0 = cx<QQ>(q, 0); c1 = cx<QQ>(q, 1); c2 = cx<QQ>(q, 2); w = weight<QQ>(q); }
  const double eqPlus = w * 6.0 * rho *
                        ((double)c0 * velLat[3 * (long long)l + 0] +
                         (double)c1 * velLat[3 * (long long)l + 1] +
                         (double)c2 * velLat[3 * (long long)l + 2]);
  const int p = links[l] - 1;
  state[(long long)(p % QQ) * S + p / QQ] = fOut + eqPlus;
}

// velocity_bounceback without the bcBuffer snapshot: one thread per boundary ELEMENT reads the
// element's QQ post-collision PDFs (rho = their sequential sum, exactly what the link loop of
// the reference sums out of bcBuffer), then writes each of its links from the element's own
// outgoing slot.  Valid when every link of the level's non-wall boundaries writes a slot of its
// own element and no element belongs to two such boundaries (checked at registration, api.cu):
// then no thread reads a value another thread -- or an earlier link of the same thread --
// has written, and the result equals the two-phase fill_bcBuffer + link loop bit for bit.
template <int QQ>
__global__ void velocityBounceBackFusedKernel(int incomp, double *__restrict__ state, long long S,
                                              int nGroups, const int32_t *__restrict__ groupStart,
                                              const int32_t *__restrict__ groupElem,
                                              const int32_t *__restrict__ links,
                                              const int32_t *__restrict__ outPos,
                                              const int32_t *__restrict__ iDir,
                                              const double *__restrict__ velLat) {
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= nGroups) return;
  const int e = groupElem[gidx];
  double rho = 0.0;
#pragma unroll
  for (int q = 0; q < QQ; ++q) rho = rho + state[(long long)q * S + e];
  if (incomp) rho = 1.0;  // rho0
  const int l1 = groupStart[gidx + 1];
  for (int l = groupStart[gidx]; l < l1; ++l) {
    const int qo = (outPos[l] - 1) % QQ;
    const double fOut = state[(long long)qo * S + e];
    const int d = iDir[l] - 1;
    int c0 = 0, c1 = 0, c2 = 0;
    double w = 0.0;
#pragma unroll
    for (int q = 0; q < QQ - 1; ++q)
      if (q == d) { c0 = cx<QQ>(q, 0); c1 = cx<QQ>(q, 1); c2 = cx<QQ>(q, 2); w = weight<QQ>(q); }
    const double eqPlus = w * 6.0 * rho *
                          ((double)c0 * velLat[3 * (long long)l + 0] +
                           (double)c1 * velLat[3 * (long long)l + 1] +
                           (double)c2 * velLat[3 * (long long)l + 2]);
    const int p = links[l] - 1;
    state[(long long)(p % QQ) * S + p / QQ] = fOut + eqPlus;
  }
}

// run-time inverse direction / lattice vector / weight through a fully unrolled compile-time table
template <int QQ>
__device__ __forceinline__ int invDirRt(int d) {
  int r = QQ - 1;
#pragma unroll
  for (int q = 0; q < QQ - 1; ++q)
    if (q == d) r = invDir<QQ>(q);
  return r;
}
template <int QQ>
__device__ __forceinline__ void dirRt(int d, int &c0, int &c1, int &c2, double &w) {
  c0 = c1 = c2 = 0;
  w = weight<QQ>(QQ - 1);
#pragma unroll
  for (int q = 0; q < QQ - 1; ++q)
    if (q == d) { c0 = cx<QQ>(q, 0); c1 = cx<QQ>(q, 1); c2 = cx<QQ>(q, 2); w = weight<QQ>(q); }
}
// SoA address of the state position FETCH(d, e) = neigh(NGPOS(d, e)) from the encoded list
template <int QQ>
__device__ __forceinline__ long long fetchAddr(const uint32_t *__restrict__ nbr, long long S, int d, int e) {
  if (d == QQ - 1) return (long long)d * S + e;
  const uint32_t n = nbr[(long long)d * S + e];
  return (n & kBounceBit) ? (long long)invDirRt<QQ>(d) * S + e : (long long)d * S + (n & kElemMask);
}

// nb(iNeigh, (iElem-1)*QQ + iDir), stored [nNeighs][nElems*QQ]; post = 0: FETCH (neighBufferPre_nNext),
// post = 1: SAVE (neighBufferPost)
template <int QQ>
__global__ void fillNeighBufferKernel(const double *__restrict__ state, long long S,
                                      const uint32_t *__restrict__ nbr, int nNeighs, int nElems,
                                      const int32_t *__restrict__ neighPos, int post,
                                      double *__restrict__ nb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNeighs * nElems * QQ) return;
  const int q = t / (nNeighs * nElems), r = t % (nNeighs * nElems);
  const int iN = r / nElems, i = r % nElems;
  const int np = neighPos[i * nNeighs + iN] - 1;
  const long long src = post ? (long long)q * S + np : fetchAddr<QQ>(nbr, S, q, np);
  nb[((long long)iN * nElems + i) * QQ + q] = state[src];
}

template <int QQ>
__device__ __forceinline__ void pdfEqAny(int incomp, double rho, double vx, double vy, double vz,
                                         double (&f)[QQ]) {
  if (QQ == 19) {
    double(&g)[19] = reinterpret_cast<double(&)[19]>(f);
    if (incomp) pdfEqIncompD3Q19(rho, vx, vy, vz, g);
    else pdfEqD3Q19(rho, vx, vy, vz, g);
  } else {
    double(&g)[27] = reinterpret_cast<double(&)[27]>(f);
    if (incomp) pdfEqIncompD3Q27(rho, vx, vy, vz, g);
    else pdfEqD3Q27(rho, vx, vy, vz, g);
  }
}

// pressure_expol, link loop (:1280-1312): 1.5 f(1) - 0.5 f(2) of neighBufferPre_nNext
template <int QQ>
__global__ void pressureExpolLinkKernel(double *__restrict__ state, long long S, int nElems, int nLinks,
                                        const int32_t *__restrict__ links,
                                        const int32_t *__restrict__ iElemOfLink,
                                        const int32_t *__restrict__ iDir,
                                        const double *__restrict__ nbPre) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nLinks) return;
  const long long sp = (long long)(iElemOfLink[l] - 1) * QQ + (iDir[l] - 1);   // outletExpol%statePos - 1
  const double fTmp_1 = nbPre[sp], fTmp_2 = nbPre[(long long)nElems * QQ + sp];
  const int p = links[l] - 1;
  state[(long long)(p % QQ) * S + p / QQ] = 1.5 * fTmp_1 - 0.5 * fTmp_2;
}

// pressure_expol, normal direction (:1316-1340), one thread per boundary element
template <int QQ>
__global__ void pressureExpolNormalKernel(int incomp, double *__restrict__ state, long long S,
                                          const uint32_t *__restrict__ nbr,
                                          const double *__restrict__ bcBuffer,
                                          const double *__restrict__ aux, int nElems,
                                          const int32_t *__restrict__ elemPos,
                                          const int32_t *__restrict__ posInBcElemBuf,
                                          const int32_t *__restrict__ normalInd,
                                          const double *__restrict__ rhoDef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nElems) return;
  const int d = normalInd[i] - 1;
  int c0, c1, c2;
  double w;
  dirRt<QQ>(d, c0, c1, c2, w);
  if (abs(c0) + abs(c1) + abs(c2) != 1) return;   // axisNormal
  const int e = elemPos[i] - 1;
  const double rho = aux[e], vx = aux[S + e], vy = aux[2 * S + e], vz = aux[3 * S + e];
  double fEq[QQ], fEq0[QQ];
  pdfEqAny<QQ>(incomp, rho, vx, vy, vz, fEq);
  pdfEqAny<QQ>(incomp, rhoDef[i], vx, vy, vz, fEq0);
  const int invD = invDirRt<QQ>(d);
  double eq0 = 0.0, eqInv = 0.0;
#pragma unroll
  for (int q = 0; q < QQ; ++q) {
    if (q == d) eq0 = fEq0[q];
    if (q == invD) eqInv = fEq[q];
  }
  const double fPostCol = bcBuffer[(long long)(posInBcElemBuf[i] - 1) * QQ + invD];
  state[fetchAddr<QQ>(nbr, S, d, e)] = eq0 + (fPostCol - eqInv);
}

// pressure_antiBounceBack (:2268-2350), one thread per link
template <int QQ>
__global__ void pressureAntiBounceBackKernel(int incomp, double *__restrict__ state, long long S,
                                             const double *__restrict__ bcBuffer, int nLinks,
                                             const int32_t *__restrict__ links,
                                             const int32_t *__restrict__ iElemOfLink,
                                             const int32_t *__restrict__ iDir,
                                             const int32_t *__restrict__ elemPos,
                                             const int32_t *__restrict__ posInBcElemBuf,
                                             const double *__restrict__ rhoDef,
                                             const double *__restrict__ omegaElem,
                                             double omegaUniform,
                                             const double *__restrict__ nbPost) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nLinks) return;
  const int i = iElemOfLink[l] - 1, d = iDir[l] - 1;
  const int invD = invDirRt<QQ>(d);
  double fT[QQ], fN[QQ];
#pragma unroll
  for (int q = 0; q < QQ; ++q) {
    fT[q] = bcBuffer[(long long)(posInBcElemBuf[i] - 1) * QQ + q];
    fN[q] = nbPost[(long long)i * QQ + q];
  }
  double rhoF, rhoN, uF[3], uN[3], uB[3];
  moments<QQ>(fT, rhoF, uF[0], uF[1], uF[2]);
  moments<QQ>(fN, rhoN, uN[0], uN[1], uN[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!incomp) { uF[k] = uF[k] / rhoF; uN[k] = uN[k] / rhoN; }
    uB[k] = 1.5 * uF[k] - 0.5 * uN[k];
  }
  const double usqB = uB[0] * uB[0] + uB[1] * uB[1] + uB[2] * uB[2];
  const double usqF = uF[0] * uF[0] + uF[1] * uF[1] + uF[2] * uF[2];
  const double om = omegaElem ? omegaElem[elemPos[i] - 1] : omegaUniform;
  int c0, c1, c2, b0, b1, b2;
  double w, wInv;
  dirRt<QQ>(d, b0, b1, b2, w);
  dirRt<QQ>(invD, c0, c1, c2, wInv);
  const double cuF = (double)c0 * uF[0] + (double)c1 * uF[1] + (double)c2 * uF[2];
  const double cuB = (double)c0 * uB[0] + (double)c1 * uB[1] + (double)c2 * uB[2];
  constexpr double rho0 = 1.0, div1_3 = 1.0 / 3.0;
  const double fEqPlusFluid = w * rhoF + 4.5 * w * rho0 * (cuF * cuF - div1_3 * usqF);
  const double fEqPlus = w * rhoDef[i] + 4.5 * w * rho0 * (cuB * cuB - div1_3 * usqB);
  double fd = 0.0, fi = 0.0;
#pragma unroll
  for (int q = 0; q < QQ; ++q) {
    if (q == d) fd = fT[q];
    if (q == invD) fi = fT[q];
  }
  const double fPlusFluid = 0.5 * (fd + fi);
  const int p = links[l] - 1;
  state[(long long)(p % QQ) * S + p / QQ] = -fi + 2.0 * fEqPlus + (2.0 - om) * (fPlusFluid - fEqPlusFluid);
}

int launchFillNeighBuffer(int QQ, const double *state, long long S, const uint32_t *nbr, int nNeighs,
                          int nElems, const int32_t *neighPos, int post, double *nb, cudaStream_t st) {
  const long long n = (long long)nNeighs * nElems * QQ;
  if (n <= 0) return 0;
  if (QQ == 19)
    fillNeighBufferKernel<19><<<divUp(n, 256), 256, 0, st>>>(state, S, nbr, nNeighs, nElems, neighPos, post, nb);
  else
    fillNeighBufferKernel<27><<<divUp(n, 256), 256, 0, st>>>(state, S, nbr, nNeighs, nElems, neighPos, post, nb);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchPressureExpol(int QQ, int incomp, double *state, long long S, const uint32_t *nbr,
                        const double *bcBuffer, const double *aux, int nElems, const int32_t *elemPos,
                        const int32_t *posInBcElemBuf, const int32_t *normalInd, const double *rhoDef,
                        int nLinks, const int32_t *links, const int32_t *iElemOfLink,
                        const int32_t *iDir, const double *nbPre, cudaStream_t st) {
  if (nElems <= 0) return 0;
  if (QQ == 19) {
    if (nLinks > 0)
      pressureExpolLinkKernel<19><<<divUp(nLinks, 128), 128, 0, st>>>(state, S, nElems, nLinks, links,
                                                                      iElemOfLink, iDir, nbPre);
    pressureExpolNormalKernel<19><<<divUp(nElems, 128), 128, 0, st>>>(
        incomp, state, S, nbr, bcBuffer, aux, nElems, elemPos, posInBcElemBuf, normalInd, rhoDef);
  } else {
    if (nLinks > 0)
      pressureExpolLinkKernel<27><<<divUp(nLinks, 128), 128, 0, st>>>(state, S, nElems, nLinks, links,
                                                                      iElemOfLink, iDir, nbPre);
    pressureExpolNormalKernel<27><<<divUp(nElems, 128), 128, 0, st>>>(
        incomp, state, S, nbr, bcBuffer, aux, nElems, elemPos, posInBcElemBuf, normalInd, rhoDef);
  }
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchPressureAntiBounceBack(int QQ, int incomp, double *state, long long S,
                                 const double *bcBuffer, int nLinks, const int32_t *links,
                                 const int32_t *iElemOfLink, const int32_t *iDir,
                                 const int32_t *elemPos, const int32_t *posInBcElemBuf,
                                 const double *rhoDef, const double *omegaElem, double omegaUniform,
                                 const double *nbPost, cudaStream_t st) {
  if (nLinks <= 0) return 0;
  if (QQ == 19)
    pressureAntiBounceBackKernel<19><<<divUp(nLinks, 128), 128, 0, st>>>(
        incomp, state, S, bcBuffer, nLinks, links, iElemOfLink, iDir, elemPos, posInBcElemBuf, rhoDef,
        omegaElem, omegaUniform, nbPost);
  else
    pressureAntiBounceBackKernel<27><<<divUp(nLinks, 128), 128, 0, st>>>(
        incomp, state, S, bcBuffer, nLinks, links, iElemOfLink, iDir, elemPos, posInBcElemBuf, rhoDef,
        omegaElem, omegaUniform, nbPost);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchFillBcBuffer(int QQ, const double *state, long long S, const int32_t *bcElems,
                       const int32_t *needed, int nNeeded, double *bcBuffer, cudaStream_t st) {
  if (nNeeded <= 0) return 0;
  fillBcBufferKernel<<<divUp((long long)nNeeded * QQ, 256), 256, 0, st>>>(state, S, QQ, bcElems, needed,
                                                                          nNeeded, bcBuffer);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchVelocityBounceBack(int QQ, int incomp, double *state, long long S,
                             const double *bcBuffer, int nLinks, const int32_t *links,
                             const int32_t *outPos, const int32_t *posInBuffer,
                             const int32_t *iDir, const double *velLat, cudaStream_t st) {
  if (nLinks <= 0) return 0;
  if (QQ == 19)
    velocityBounceBackKernel<19><<<divUp(nLinks, 256), 256, 0, st>>>(
        incomp, state, S, bcBuffer, nLinks, links, outPos, posInBuffer, iDir, velLat);
  else
    velocityBounceBackKernel<27><<<divUp(nLinks, 256), 256, 0, st>>>(
        incomp, state, S, bcBuffer, nLinks, links, outPos, posInBuffer, iDir, velLat);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchVelocityBounceBackFused(int QQ, int incomp, double *state, long long S, int nGroups,
                                  const int32_t *groupStart, const int32_t *groupElem,
                                  const int32_t *links, const int32_t *outPos, const int32_t *iDir,
                                  const double *velLat, cudaStream_t st) {
  if (nGroups <= 0) return 0;
  if (QQ == 19)
    velocityBounceBackFusedKernel<19><<<divUp(nGroups, 128), 128, 0, st>>>(
        incomp, state, S, nGroups, groupStart, groupElem, links, outPos, iDir, velLat);
  else
    velocityBounceBackFusedKernel<27><<<divUp(nGroups, 128), 128, 0, st>>>(
        incomp, state, S, nGroups, groupStart, groupElem, links, outPos, iDir, velLat);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace musb200
