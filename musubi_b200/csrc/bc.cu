// bc.cu -- boundary kernels, applied link-wise on the post-collision state
// before the now/next swap (set_boundary, mus/source/bc/mus_bc_general_module.fpp:179-285).
//
//   fill_bcBuffer        mus_bc_general_module.fpp:1726-1768
//   velocity_bounceback  mus_bc_fluid_module.fpp:1503-1597   (and _incomp :1401-1490)
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

}  // namespace musb200
