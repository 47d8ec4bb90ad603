// intp.cu -- ghost interpolation kernels (see intp.cuh for the reference routines).
#include "intp.cuh"
#include "equilibrium.cuh"
#include <algorithm>
#include <utility>

namespace musb200 {

void IntpSet::release() {
  auto fr = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
  fr(targets); fr(srcOffset); fr(srcSlot); fr(uniqueSrc); fr(weights); fr(posInMat); fr(matOffset);
  fr(matrices); fr(coord); fr(scratch);
  nTargets = 0; nMatrices = 0; nUnique = 0; maxSrc = 0;
}

IntpSet &IntpSet::operator=(IntpSet &&o) noexcept {
  if (this != &o) {
    release();
    order = o.order; nTargets = o.nTargets; nMatrices = o.nMatrices; nUnique = o.nUnique;
    targets = o.targets; srcOffset = o.srcOffset; srcSlot = o.srcSlot; uniqueSrc = o.uniqueSrc;
    weights = o.weights; posInMat = o.posInMat; matOffset = o.matOffset; matrices = o.matrices;
    coord = o.coord; scratch = o.scratch;
    maxSrc = o.maxSrc; o.maxSrc = 0;
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
  if (rc) return rc;
  set.nTargets = nTargets;
  set.maxSrc = 0;
  for (int i = 0; i < nTargets; ++i) set.maxSrc = std::max(set.maxSrc, srcOffset[i + 1] - srcOffset[i]);
  set.nMatrices = nMatrices;
  set.nUnique = (int)uniq.size();
  MUSB_CUDA(cudaMalloc(&set.scratch, (size_t)2 * 27 * set.nUnique * sizeof(double)));
  MUSB_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
  return 0;
}

// ---------------------------------------------------------------------------
// phase A: f_eq / f_neq of every distinct source, written source-major with the pair of a
// direction side by side ([u][d] = {f_eq, f_neq}: one 16-byte load in phase B) through a
// shared-memory tile so that the global
// stores of a CTA are one contiguous, fully coalesced run
template <int QQ, int THREADS>
__global__ void __launch_bounds__(THREADS) eqNeqKernel(int incomp, const double *__restrict__ sState,
                            const double *__restrict__ sAux, long long sS,
                            const int32_t *__restrict__ uniqueSrc, int nUnique,
                            double *__restrict__ scratch) {
  constexpr int W = 2 * QQ, P = W + 1;       // odd pitch: conflict-free column access
  __shared__ double tile[THREADS * P];
  const int u = blockIdx.x * THREADS + threadIdx.x;
  if (u < nUnique) {
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
      tile[threadIdx.x * P + 2 * q] = feq[q];
      tile[threadIdx.x * P + 2 * q + 1] = f - feq[q];
    }
  }
  __syncthreads();
  const int rows = min(THREADS, nUnique - (int)blockIdx.x * THREADS);
  double *out = scratch + (long long)blockIdx.x * THREADS * W;
  for (int t = threadIdx.x; t < rows * W; t += THREADS) out[t] = tile[(t / W) * P + (t % W)];
}

__device__ __forceinline__ double omegaFromVisc(double v) { return 1.0 / (3.0 * v + 0.5); }
// getNonEqFac_intp, PULL build (post-collision PDFs), mus_derivedQuantities_module.fpp:601-614
__device__ __forceinline__ double neqFac(double omegaS, double omegaT) {
  return omegaS * (1.0 - omegaT) / ((1.0 - omegaS) * omegaT);
}

// MODE 1: weighted average; 2: linear; 3: quadratic (the average from finer has its own kernel
// below).  One thread per (target, direction) with the DIRECTION fastest, so the lanes of a warp
// belong to one or two targets: source slots and least-square matrix are warp-uniform loads, and
// the {f_eq, f_neq} pairs of a source are one contiguous run of the source-major scratch, fetched
// as double2.  Per (target, direction) the sources are accumulated in the host's order with all
// polynomial coefficients side by side in registers -- per coefficient the sum over the sources
// of the reference's matrix-vector product, so the bits are the same.  The non-equilibrium factor
// -- three divisions -- is computed once on the host when the level's viscosity is uniform.
// ncu: L1 wavefronts 62 % of peak, FP64 26 %, L2 14 %, DRAM 3 %: a latency / L1 mix; six other
// mappings were measured and none beat this one (profiles/r01_intp_cfg4.md).
template <int MODE>
__global__ void __launch_bounds__(128) intpKernel(int QQ, const double *__restrict__ scratch, int nUnique,
                           int nTargets, const int32_t *__restrict__ targets,
                           const int32_t *__restrict__ srcOffset,
                           const int32_t *__restrict__ srcSlot, const double *__restrict__ weights,
                           const int32_t *__restrict__ posInMat,
                           const int32_t *__restrict__ matOffset,
                           const double *__restrict__ matrices, const double *__restrict__ coord,
                           double *__restrict__ tState, long long tS,
                           const double *__restrict__ tVisc, double tViscUniform) {
  constexpr int nCoeff = MODE == 1 ? 1 : (MODE == 2 ? 4 : 10);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nTargets * QQ) return;
  const int i = (int)(idx / QQ), d = (int)(idx % QQ);
  const int tgt = targets[i] - 1;
  const int s0 = srcOffset[i], n = srcOffset[i + 1] - s0;
  const double2 *pair = reinterpret_cast<const double2 *>(scratch) + d;   // + u * QQ
  const double *A = MODE == 1 ? weights + s0 : matrices + matOffset[posInMat[i]];
  double ce[nCoeff], cn[nCoeff];
#pragma unroll
  for (int k = 0; k < nCoeff; ++k) { ce[k] = 0.0; cn[k] = 0.0; }
  for (int s = 0; s < n; ++s) {
    const double2 v = pair[(long long)srcSlot[s0 + s] * QQ];
#pragma unroll
    for (int k = 0; k < nCoeff; ++k) {
      const double m = A[k * n + s];   // A(k, s), row-major (nCoeff x n); MODE 1: w(s)
      ce[k] = ce[k] + m * v.x;
      cn[k] = cn[k] + m * v.y;
    }
  }
  double t_eq, t_neq;
  if (MODE == 1) {
    t_eq = ce[0];
    t_neq = cn[0];
  } else {
    const double x = coord[3 * i + 0], y = coord[3 * i + 1], z = coord[3 * i + 2];
    t_eq = ce[0] + ce[1] * x + ce[2] * y + ce[3] * z;
    t_neq = cn[0] + cn[1] * x + cn[2] * y + cn[3] * z;
    if (MODE == 3) {
      t_eq = t_eq + ce[4] * x * x + ce[5] * y * y + ce[6] * z * z + ce[7] * x * y + ce[8] * y * z +
             ce[9] * z * x;
      t_neq = t_neq + cn[4] * x * x + cn[5] * y * y + cn[6] * z * z + cn[7] * x * y + cn[8] * y * z +
              cn[9] * z * x;
    }
  }
  double fac;   // 0.5 * getNonEqFac_intp_coarse_to_fine
  if (tVisc) {
    const double visc = tVisc[tgt];
    fac = 0.5 * neqFac(omegaFromVisc(0.5 * visc), omegaFromVisc(visc));
  } else {
    fac = tViscUniform;   // the factor itself, evaluated by the launcher with the same expression
  }
  t_neq = t_neq * fac;
  tState[(long long)d * tS + tgt] = t_neq + t_eq;
}

// fillMyGhostsFromFiner_avg_feq_fneq without a scratch pass: 8 lanes per coarse ghost, one lane per
// child (the children are Morton-contiguous fine elements: coalesced loads).  Each lane forms
// f_eq(rho, u from auxField) and f - f_eq of its child and parks them in shared memory; the lanes
// of the group then share the QQ directions, each summing its directions over the children in
// the children's order, a = (..((c1 + c2) + c3)..), as the reference's loop does.
template <int QQ, int THREADS>
__global__ void __launch_bounds__(THREADS) fromFinerFusedKernel(int incomp, const double *__restrict__ sState,
                                     const double *__restrict__ sAux, long long sS,
                                     const int32_t *__restrict__ uniqueSrc, int nTargets,
                                     const int32_t *__restrict__ targets,
                                     const int32_t *__restrict__ srcOffset,
                                     const int32_t *__restrict__ srcSlot,
                                     double *__restrict__ tState, double *__restrict__ tAux, long long tS,
                                     const double *__restrict__ tVisc, double tViscUniform) {
  // tAux != nullptr: the auxField average of mus_intpAuxFieldCoarserAndExchange
  // (fillArbiMyGhostsFromFiner_avg) is taken in the same pass -- same sources, and nothing
  // between the two calls of the reference's schedule writes them
  constexpr int W = 2 * QQ, P = W + 5;       // + rho, ux, uy, uz; odd pitch
  __shared__ double tile[THREADS * P];
  const int idx = blockIdx.x * THREADS + threadIdx.x;
  const int i = idx >> 3, c = idx & 7;
  const bool valid = i < nTargets;
  int n = 0, s0 = 0, tgt = 0;
  if (valid) { s0 = srcOffset[i]; n = srcOffset[i + 1] - s0; tgt = targets[i] - 1; }
  if (c < n) {
    const int e = uniqueSrc[srcSlot[s0 + c]] - 1;
    const double rho = sAux[e], vx = sAux[sS + e], vy = sAux[2 * sS + e], vz = sAux[3 * sS + e];
    double f[QQ], eq[QQ];
#pragma unroll
    for (int q = 0; q < QQ; ++q) f[q] = sState[(long long)q * sS + e];
    if (QQ == 19) {
      double(&g)[19] = reinterpret_cast<double(&)[19]>(eq);
      if (incomp) pdfEqIncompD3Q19(rho, vx, vy, vz, g);
      else pdfEqD3Q19(rho, vx, vy, vz, g);
    } else {
      double(&g)[27] = reinterpret_cast<double(&)[27]>(eq);
      if (incomp) pdfEqIncompD3Q27(rho, vx, vy, vz, g);
      else pdfEqD3Q27(rho, vx, vy, vz, g);
    }
    double *row = tile + threadIdx.x * P;
#pragma unroll
    for (int q = 0; q < QQ; ++q) {
      row[q] = eq[q];
      row[QQ + q] = f[q] - eq[q];
    }
    row[W] = rho; row[W + 1] = vx; row[W + 2] = vy; row[W + 3] = vz;
  }
  __syncthreads();
  if (!valid) return;
  const double visc = tVisc ? tVisc[tgt] : tViscUniform;
  const double inv_n = 1.0 / (double)n;
  const double fOmega = omegaFromVisc(2.0 * visc), cOmega = omegaFromVisc(visc);
  const double fac = 2.0 * neqFac(fOmega, cOmega);  // getNonEqFac_intp_fine_to_coarse
  const double *grp = tile + (threadIdx.x & ~7) * P;   // rows of my group's children
  for (int q = c; q < QQ; q += 8) {
    double a = 0.0, b = 0.0;
    for (int s = 0; s < n; ++s) {
      a = a + grp[s * P + q];
      b = b + grp[s * P + QQ + q];
    }
    const double t_eq = a * inv_n;
    const double t_neq = b * inv_n * fac;
    tState[(long long)q * tS + tgt] = t_eq + t_neq;
  }
  if (tAux != nullptr && c < 4) {
    double t = 0.0;
    for (int s = 0; s < n; ++s) t = grp[s * P + W + c] + t;
    tAux[(long long)c * tS + tgt] = t * inv_n;
  }
}

// host copies of omegaFromVisc / neqFac: IEEE double on both sides, identical bits
static double hostOmega(double v) { return 1.0 / (3.0 * v + 0.5); }
static double hostNeqFac(double omegaS, double omegaT) { return omegaS * (1.0 - omegaT) / ((1.0 - omegaS) * omegaT); }

int launchIntp(const IntpArgs &a, const IntpSet &set, bool fromFiner, cudaStream_t st, int *nLaunch) {
  if (nLaunch) *nLaunch = 0;
  if (set.nTargets == 0) return 0;
  const int B = 128;
  if (fromFiner) {
    if (set.maxSrc > 8) return setError(1, "a ghostFromFiner element has more than 8 children");
    if (a.QQ == 19)   // 128 threads x 43 doubles = 44 KB of shared memory
      fromFinerFusedKernel<19, 128><<<divUp((long long)set.nTargets * 8, 128), 128, 0, st>>>(
          a.incomp, a.sState, a.sAux, a.sS, set.uniqueSrc, set.nTargets, set.targets, set.srcOffset,
          set.srcSlot, a.tState, a.withAux ? a.tAux : nullptr, a.tS, a.tVisc, a.tViscUniform);
    else              // 64 threads x 59 doubles = 30 KB
      fromFinerFusedKernel<27, 64><<<divUp((long long)set.nTargets * 8, 64), 64, 0, st>>>(
          a.incomp, a.sState, a.sAux, a.sS, set.uniqueSrc, set.nTargets, set.targets, set.srcOffset,
          set.srcSlot, a.tState, a.withAux ? a.tAux : nullptr, a.tS, a.tVisc, a.tViscUniform);
    MUSB_CUDA(cudaGetLastError());
    if (nLaunch) *nLaunch = 1;
    return 0;
  }
  if (a.QQ == 19)
    eqNeqKernel<19, 128><<<divUp(set.nUnique, 128), 128, 0, st>>>(a.incomp, a.sState, a.sAux, a.sS,
                                                                  set.uniqueSrc, set.nUnique, set.scratch);
  else
    eqNeqKernel<27, 64><<<divUp(set.nUnique, 64), 64, 0, st>>>(a.incomp, a.sState, a.sAux, a.sS,
                                                                set.uniqueSrc, set.nUnique, set.scratch);
  MUSB_CUDA(cudaGetLastError());
  const int mode = 1 + set.order;
  if (mode == 1 && !set.weights) return setError(1, "weighted-average set without weights");
  const int grid = divUp((long long)set.nTargets * a.QQ, B);
  // uniform viscosity: hand over 0.5 * getNonEqFac_intp_coarse_to_fine instead of the viscosity
  const double facOrVisc =
      a.tVisc ? 0.0 : 0.5 * hostNeqFac(hostOmega(0.5 * a.tViscUniform), hostOmega(a.tViscUniform));
#define MUSB_INTP(M)                                                                              \
  intpKernel<M><<<grid, B, 0, st>>>(a.QQ, set.scratch, set.nUnique, set.nTargets, set.targets,    \
                                    set.srcOffset, set.srcSlot, set.weights, set.posInMat,        \
                                    set.matOffset, set.matrices, set.coord, a.tState, a.tS,       \
                                    a.tVisc, facOrVisc)
  if (mode == 1) MUSB_INTP(1);
  else if (mode == 2) MUSB_INTP(2);
  else if (mode == 3) MUSB_INTP(3);
  else return setError(1, "interpolation order must be 0, 1 or 2");
#undef MUSB_INTP
  MUSB_CUDA(cudaGetLastError());
  if (nLaunch) *nLaunch = 2;
  return 0;
}

}  // namespace musb200
