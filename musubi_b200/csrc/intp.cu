// intp.cu -- ghost interpolation kernels (see intp.cuh for the reference routines).
#include "intp.cuh"
#include "equilibrium.cuh"
#include <algorithm>
#include <utility>

namespace musb200 {

void IntpSet::release() {
  auto fr = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
  fr(targets); fr(srcOffset); fr(srcSlot); fr(uniqueSrc); fr(weights); fr(posInMat); fr(matOffset);
  fr(matrices); fr(coord); fr(scratch); fr(sel7); fr(sel8); fr(selRest);
  nTargets = 0; nMatrices = 0; nUnique = 0; n7 = 0; n8 = 0; nRest = 0;
}

IntpSet &IntpSet::operator=(IntpSet &&o) noexcept {
  if (this != &o) {
    release();
    order = o.order; nTargets = o.nTargets; nMatrices = o.nMatrices; nUnique = o.nUnique;
    targets = o.targets; srcOffset = o.srcOffset; srcSlot = o.srcSlot; uniqueSrc = o.uniqueSrc;
    weights = o.weights; posInMat = o.posInMat; matOffset = o.matOffset; matrices = o.matrices;
    coord = o.coord; scratch = o.scratch;
    sel7 = o.sel7; sel8 = o.sel8; selRest = o.selRest; n7 = o.n7; n8 = o.n8; nRest = o.nRest;
    o.sel7 = nullptr; o.sel8 = nullptr; o.selRest = nullptr; o.n7 = 0; o.n8 = 0; o.nRest = 0;
    o.targets = nullptr; o.srcOffset = nullptr; o.srcSlot = nullptr; o.uniqueSrc = nullptr;
    o.weights = nullptr; o.posInMat = nullptr; o.matOffset = nullptr; o.matrices = nullptr;
    o.coord = nullptr; o.scratch = nullptr;
    o.nTargets = 0; o.nMatrices = 0; o.nUnique = 0;
  }
  return *this;
}

template <class T>
static int up(T *&dst, const T *src, size_t n, cudaStream_t st) {
  dst = nullptr;
  if (n == 0) return 0;
  MUSB_CUDA(cudaMalloc(&dst, n * sizeof(T)));
  MUSB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

int registerIntp(IntpSet &set, int order, int nTargets, const int32_t *targetList,
                 const int32_t *srcOffset, const int32_t *srcPos, const double *weights,
                 const int32_t *posInMat, int nMatrices, const int32_t *matOffset,
                 const double *matrices, const double *childCoord, cudaStream_t st) {
  set.release();
  set.order = order;
  if (nTargets == 0) return 0;
  if (!targetList || !srcOffset || !srcPos) return setError(1, "interpolation lists missing");
  const int nSrc = srcOffset[nTargets];
  // distinct sources (sourceFromCoarser of the reference) and the CSR re-expressed in slots
  std::vector<int32_t> uniq(srcPos, srcPos + nSrc);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  std::vector<int32_t> slot(nSrc);
  for (int i = 0; i < nSrc; ++i)
    slot[i] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), srcPos[i]) - uniq.begin());
  for (int i = 0; i < nTargets; ++i)
    if (srcOffset[i + 1] - srcOffset[i] < 1 || srcOffset[i + 1] - srcOffset[i] > 27)
      return setError(1, "an interpolation target needs between 1 and 27 sources");
  int rc = 0;
  rc |= up(set.targets, targetList, nTargets, st);
  rc |= up(set.srcOffset, srcOffset, (size_t)nTargets + 1, st);
  rc |= up(set.srcSlot, slot.data(), (size_t)nSrc, st);
  rc |= up(set.uniqueSrc, uniq.data(), uniq.size(), st);
  if (order == 0 && weights) rc |= up(set.weights, weights, (size_t)nSrc, st);
  if (order > 0) {
    if (!posInMat || !matOffset || !matrices || !childCoord || nMatrices < 1)
      return setError(1, "least-square interpolation needs matrices, posInMat and coordinates");
    rc |= up(set.posInMat, posInMat, nTargets, st);
    rc |= up(set.matOffset, matOffset, (size_t)nMatrices + 1, st);
    rc |= up(set.matrices, matrices, (size_t)matOffset[nMatrices], st);
    rc |= up(set.coord, childCoord, (size_t)3 * nTargets, st);
  }
  if (order == 1) {
    std::vector<int32_t> s7, s8, sr;
    for (int i = 0; i < nTargets; ++i) {
      const int n = srcOffset[i + 1] - srcOffset[i];
      (n == 7 ? s7 : (n == 8 ? s8 : sr)).push_back(i);
    }
    rc |= up(set.sel7, s7.data(), s7.size(), st);
    rc |= up(set.sel8, s8.data(), s8.size(), st);
    rc |= up(set.selRest, sr.data(), sr.size(), st);
    set.n7 = (int)s7.size(); set.n8 = (int)s8.size(); set.nRest = (int)sr.size();
  }
  if (rc) return rc;
  set.nTargets = nTargets;
  set.nMatrices = nMatrices;
  set.nUnique = (int)uniq.size();
  MUSB_CUDA(cudaMalloc(&set.scratch, (size_t)2 * 27 * set.nUnique * sizeof(double)));
  MUSB_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
  return 0;
}

// ---------------------------------------------------------------------------
template <int QQ>
__global__ void eqNeqKernel(int incomp, const double *__restrict__ sState,
                            const double *__restrict__ sAux, long long sS,
                            const int32_t *__restrict__ uniqueSrc, int nUnique,
                            double *__restrict__ scratch) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nUnique) return;
  const int e = uniqueSrc[u] - 1;
  const double rho = sAux[e], vx = sAux[sS + e], vy = sAux[2 * sS + e], vz = sAux[3 * sS + e];
  double feq[QQ];
  if (QQ == 19) {
    double(&g)[19] = reinterpret_cast<double(&)[19]>(feq);
    if (incomp) pdfEqIncompD3Q19(rho, vx, vy, vz, g);
    else pdfEqD3Q19(rho, vx, vy, vz, g);
  } else {
    double(&g)[27] = reinterpret_cast<double(&)[27]>(feq);
    if (incomp) pdfEqIncompD3Q27(rho, vx, vy, vz, g);
    else pdfEqD3Q27(rho, vx, vy, vz, g);
  }
#pragma unroll
  for (int q = 0; q < QQ; ++q) {
    const double f = sState[(long long)q * sS + e];
    scratch[(long long)q * nUnique + u] = feq[q];
    scratch[(long long)(27 + q) * nUnique + u] = f - feq[q];
  }
}

__device__ __forceinline__ double omegaFromVisc(double v) { return 1.0 / (3.0 * v + 0.5); }
// getNonEqFac_intp, PULL build (post-collision PDFs), mus_derivedQuantities_module.fpp:601-614
__device__ __forceinline__ double neqFac(double omegaS, double omegaT) {
  return omegaS * (1.0 - omegaT) / ((1.0 - omegaS) * omegaT);
}

// MODE 0: average from finer; 1: weighted average; 2: linear; 3: quadratic.
// One thread per (target, direction), target index fastest (coalesced stores into the ghost
// block).  The sources are visited ONCE, in the host's order, and all polynomial coefficients
// are accumulated side by side in registers: per coefficient the sum runs over the sources in
// ascending order exactly as in the reference's matrix-vector product, so the bits are the same
// as evaluating coefficient after coefficient, with a quarter (linear) or a tenth (quadratic) of
// the gathers.
template <int MODE>
__global__ void __launch_bounds__(128) intpKernel(int QQ, const double *__restrict__ scratch, int nUnique,
                           int nTargets, const int32_t *__restrict__ sel,
                           const int32_t *__restrict__ targets,
                           const int32_t *__restrict__ srcOffset,
                           const int32_t *__restrict__ srcSlot, const double *__restrict__ weights,
                           const int32_t *__restrict__ posInMat,
                           const int32_t *__restrict__ matOffset,
                           const double *__restrict__ matrices, const double *__restrict__ coord,
                           double *__restrict__ tState, long long tS,
                           const double *__restrict__ tVisc, double tViscUniform) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nTargets * QQ) return;
  const int d = idx / nTargets;
  const int i = sel ? sel[idx % nTargets] : idx % nTargets;   // nTargets = list length when sel
  const int tgt = targets[i] - 1;
  const int s0 = srcOffset[i], n = srcOffset[i + 1] - s0;
  const double *eq = scratch + (long long)d * nUnique;
  const double *neq = scratch + (long long)(27 + d) * nUnique;
  const double visc = tVisc ? tVisc[tgt] : tViscUniform;
  double t_eq, t_neq;
  if (MODE == 0) {
    const double inv_n = 1.0 / (double)n;
    double a = 0.0, b = 0.0;
    for (int s = 0; s < n; ++s) {
      const int u = srcSlot[s0 + s];
      a = a + eq[u];
      b = b + neq[u];
    }
    const double fOmega = omegaFromVisc(2.0 * visc), cOmega = omegaFromVisc(visc);
    const double fac = 2.0 * neqFac(fOmega, cOmega);  // getNonEqFac_intp_fine_to_coarse
    t_eq = a * inv_n;
    t_neq = b * inv_n * fac;
    tState[(long long)d * tS + tgt] = t_eq + t_neq;
    return;
  }
  if (MODE == 1) {
    double a = 0.0, b = 0.0;
    for (int s = 0; s < n; ++s) {
      const int u = srcSlot[s0 + s];
      const double w = weights[s0 + s];
      a = a + w * eq[u];
      b = b + w * neq[u];
    }
    t_eq = a;
    t_neq = b;
  } else {
    constexpr int nCoeff = MODE == 2 ? 4 : 10;
    const double *A = matrices + matOffset[posInMat[i]];
    const double x = coord[3 * i + 0], y = coord[3 * i + 1], z = coord[3 * i + 2];
    double ce[nCoeff], cn[nCoeff];
#pragma unroll
    for (int k = 0; k < nCoeff; ++k) { ce[k] = 0.0; cn[k] = 0.0; }
    for (int s = 0; s < n; ++s) {
      const int u = srcSlot[s0 + s];
      const double e = eq[u], ne = neq[u];
#pragma unroll
      for (int k = 0; k < nCoeff; ++k) {
        const double m = A[(long long)k * n + s];
        ce[k] = ce[k] + m * e;
        cn[k] = cn[k] + m * ne;
      }
    }
    t_eq = ce[0] + ce[1] * x + ce[2] * y + ce[3] * z;
    t_neq = cn[0] + cn[1] * x + cn[2] * y + cn[3] * z;
    if (MODE == 3) {
      t_eq = t_eq + ce[4] * x * x + ce[5] * y * y + ce[6] * z * z + ce[7] * x * y + ce[8] * y * z +
             ce[9] * z * x;
      t_neq = t_neq + cn[4] * x * x + cn[5] * y * y + cn[6] * z * z + cn[7] * x * y + cn[8] * y * z +
              cn[9] * z * x;
    }
  }
  const double fOmega = omegaFromVisc(visc), cOmega = omegaFromVisc(0.5 * visc);
  const double fac = 0.5 * neqFac(cOmega, fOmega);  // getNonEqFac_intp_coarse_to_fine
  t_neq = t_neq * fac;
  tState[(long long)d * tS + tgt] = t_neq + t_eq;
}

// Linear interpolation for targets with exactly NS sources: one thread per target keeps the
// source slots and the 4 x NS least-square matrix in registers and walks the directions, so the
// list and matrix loads are paid once per target instead of once per (target, direction).  Per
// (target, direction) the arithmetic is the sequence of intpKernel<2>.
template <int NS>
__global__ void __launch_bounds__(128) intpLinearPerTargetKernel(
    int QQ, const double *__restrict__ scratch, int nUnique, int nSel, const int32_t *__restrict__ sel,
    const int32_t *__restrict__ targets, const int32_t *__restrict__ srcOffset,
    const int32_t *__restrict__ srcSlot, const int32_t *__restrict__ posInMat,
    const int32_t *__restrict__ matOffset, const double *__restrict__ matrices,
    const double *__restrict__ coord, double *__restrict__ tState, long long tS,
    const double *__restrict__ tVisc, double tViscUniform) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nSel) return;
  const int i = sel[t];
  const int tgt = targets[i] - 1;
  const int s0 = srcOffset[i];
  int u[NS];
  double A[4][NS];
  const double *Am = matrices + matOffset[posInMat[i]];
#pragma unroll
  for (int s = 0; s < NS; ++s) u[s] = srcSlot[s0 + s];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int s = 0; s < NS; ++s) A[k][s] = Am[k * NS + s];
  const double x = coord[3 * i + 0], y = coord[3 * i + 1], z = coord[3 * i + 2];
  const double visc = tVisc ? tVisc[tgt] : tViscUniform;
  const double fOmega = omegaFromVisc(visc), cOmega = omegaFromVisc(0.5 * visc);
  const double fac = 0.5 * neqFac(cOmega, fOmega);  // getNonEqFac_intp_coarse_to_fine
  for (int d = 0; d < QQ; ++d) {
    const double *eq = scratch + (long long)d * nUnique;
    const double *neq = scratch + (long long)(27 + d) * nUnique;
    double ce[4] = {0.0, 0.0, 0.0, 0.0}, cn[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const double e = eq[u[s]], ne = neq[u[s]];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ce[k] = ce[k] + A[k][s] * e;
        cn[k] = cn[k] + A[k][s] * ne;
      }
    }
    const double t_eq = ce[0] + ce[1] * x + ce[2] * y + ce[3] * z;
    double t_neq = cn[0] + cn[1] * x + cn[2] * y + cn[3] * z;
    t_neq = t_neq * fac;
    tState[(long long)d * tS + tgt] = t_neq + t_eq;
  }
}

// fillMyGhostsFromFiner_avg_feq_fneq without the scratch pass: one thread per coarse ghost walks
// its (at most 8, Morton-contiguous) children, forms f_eq(rho, u from auxField) and f - f_eq of
// each and accumulates both per direction in the children's order -- the sums of intpKernel<0>.
template <int QQ>
__global__ void __launch_bounds__(64) fromFinerFusedKernel(int incomp, const double *__restrict__ sState,
                                     const double *__restrict__ sAux, long long sS,
                                     const int32_t *__restrict__ uniqueSrc, int nTargets,
                                     const int32_t *__restrict__ targets,
                                     const int32_t *__restrict__ srcOffset,
                                     const int32_t *__restrict__ srcSlot,
                                     double *__restrict__ tState, long long tS,
                                     const double *__restrict__ tVisc, double tViscUniform) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nTargets) return;
  const int tgt = targets[i] - 1;
  const int s0 = srcOffset[i], n = srcOffset[i + 1] - s0;
  double a[QQ], b[QQ];
#pragma unroll
  for (int q = 0; q < QQ; ++q) { a[q] = 0.0; b[q] = 0.0; }
  for (int s = 0; s < n; ++s) {
    const int e = uniqueSrc[srcSlot[s0 + s]] - 1;
    const double rho = sAux[e], vx = sAux[sS + e], vy = sAux[2 * sS + e], vz = sAux[3 * sS + e];
    double feq[QQ];
    if (QQ == 19) {
      double(&g)[19] = reinterpret_cast<double(&)[19]>(feq);
      if (incomp) pdfEqIncompD3Q19(rho, vx, vy, vz, g);
      else pdfEqD3Q19(rho, vx, vy, vz, g);
    } else {
      double(&g)[27] = reinterpret_cast<double(&)[27]>(feq);
      if (incomp) pdfEqIncompD3Q27(rho, vx, vy, vz, g);
      else pdfEqD3Q27(rho, vx, vy, vz, g);
    }
#pragma unroll
    for (int q = 0; q < QQ; ++q) {
      const double f = sState[(long long)q * sS + e];
      a[q] = a[q] + feq[q];
      b[q] = b[q] + (f - feq[q]);
    }
  }
  const double visc = tVisc ? tVisc[tgt] : tViscUniform;
  const double inv_n = 1.0 / (double)n;
  const double fOmega = omegaFromVisc(2.0 * visc), cOmega = omegaFromVisc(visc);
  const double fac = 2.0 * neqFac(fOmega, cOmega);  // getNonEqFac_intp_fine_to_coarse
#pragma unroll
  for (int q = 0; q < QQ; ++q) {
    const double t_eq = a[q] * inv_n;
    const double t_neq = b[q] * inv_n * fac;
    tState[(long long)q * tS + tgt] = t_eq + t_neq;
  }
}

// fillArbiMyGhostsFromFiner_avg for the 4 auxField scalars
__global__ void auxFromFinerKernel(const double *__restrict__ sAux, long long sS,
                                   const int32_t *__restrict__ uniqueSrc, int nTargets,
                                   const int32_t *__restrict__ targets,
                                   const int32_t *__restrict__ srcOffset,
                                   const int32_t *__restrict__ srcSlot, double *__restrict__ tAux,
                                   long long tS) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nTargets * 4) return;
  const int i = idx % nTargets, k = idx / nTargets;
  const int s0 = srcOffset[i], n = srcOffset[i + 1] - s0;
  const double inv_n = 1.0 / (double)n;
  double t = 0.0;
  for (int s = 0; s < n; ++s) {
    const int e = uniqueSrc[srcSlot[s0 + s]] - 1;
    t = sAux[(long long)k * sS + e] + t;
  }
  tAux[(long long)k * tS + targets[i] - 1] = t * inv_n;
}

int launchIntp(const IntpArgs &a, const IntpSet &set, bool fromFiner, cudaStream_t st, int *nLaunch) {
  if (nLaunch) *nLaunch = 0;
  if (set.nTargets == 0) return 0;
  const int B = 128;
  if (fromFiner) {
    if (a.QQ == 19)
      fromFinerFusedKernel<19><<<divUp(set.nTargets, 64), 64, 0, st>>>(
          a.incomp, a.sState, a.sAux, a.sS, set.uniqueSrc, set.nTargets, set.targets, set.srcOffset,
          set.srcSlot, a.tState, a.tS, a.tVisc, a.tViscUniform);
    else
      fromFinerFusedKernel<27><<<divUp(set.nTargets, 64), 64, 0, st>>>(
          a.incomp, a.sState, a.sAux, a.sS, set.uniqueSrc, set.nTargets, set.targets, set.srcOffset,
          set.srcSlot, a.tState, a.tS, a.tVisc, a.tViscUniform);
    MUSB_CUDA(cudaGetLastError());
    if (nLaunch) *nLaunch = 1;
    return 0;
  }
  if (a.QQ == 19)
    eqNeqKernel<19><<<divUp(set.nUnique, B), B, 0, st>>>(a.incomp, a.sState, a.sAux, a.sS,
                                                         set.uniqueSrc, set.nUnique, set.scratch);
  else
    eqNeqKernel<27><<<divUp(set.nUnique, B), B, 0, st>>>(a.incomp, a.sState, a.sAux, a.sS,
                                                         set.uniqueSrc, set.nUnique, set.scratch);
  MUSB_CUDA(cudaGetLastError());
  const int mode = 1 + set.order;
  if (mode == 1 && !set.weights) return setError(1, "weighted-average set without weights");
  int launches = 1;
#define MUSB_INTP(M, N, SEL)                                                                      \
  intpKernel<M><<<divUp((long long)(N) * a.QQ, B), B, 0, st>>>(                                   \
      a.QQ, set.scratch, set.nUnique, (N), (SEL), set.targets, set.srcOffset, set.srcSlot,        \
      set.weights, set.posInMat, set.matOffset, set.matrices, set.coord, a.tState, a.tS, a.tVisc, \
      a.tViscUniform)
#define MUSB_INTP_PT(NS, N, SEL)                                                                  \
  intpLinearPerTargetKernel<NS><<<divUp((N), B), B, 0, st>>>(                                     \
      a.QQ, set.scratch, set.nUnique, (N), (SEL), set.targets, set.srcOffset, set.srcSlot,        \
      set.posInMat, set.matOffset, set.matrices, set.coord, a.tState, a.tS, a.tVisc,              \
      a.tViscUniform)
  if (mode == 1) { MUSB_INTP(1, set.nTargets, nullptr); ++launches; }
  else if (mode == 2) {
    if (set.n7) { MUSB_INTP_PT(7, set.n7, set.sel7); ++launches; }
    if (set.n8) { MUSB_INTP_PT(8, set.n8, set.sel8); ++launches; }
    if (set.nRest) { MUSB_INTP(2, set.nRest, set.selRest); ++launches; }
  } else if (mode == 3) { MUSB_INTP(3, set.nTargets, nullptr); ++launches; }
  else return setError(1, "interpolation order must be 0, 1 or 2");
#undef MUSB_INTP
#undef MUSB_INTP_PT
  MUSB_CUDA(cudaGetLastError());
  if (nLaunch) *nLaunch = launches;
  return 0;
}

int launchAuxFromFiner(const IntpArgs &a, const IntpSet &set, cudaStream_t st) {
  if (set.nTargets == 0) return 0;
  auxFromFinerKernel<<<divUp((long long)set.nTargets * 4, 128), 128, 0, st>>>(
      a.sAux, a.sS, set.uniqueSrc, set.nTargets, set.targets, set.srcOffset, set.srcSlot, a.tAux, a.tS);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace musb200
