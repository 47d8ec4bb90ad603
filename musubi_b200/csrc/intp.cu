// intp.cu -- ghost interpolation kernels (see intp.cuh for the reference routines).
#include "intp.cuh"
#include "equilibrium.cuh"
#include <algorithm>
#include <utility>

namespace musb200 {

void IntpSet::release() {
  auto fr = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
  fr(targets); fr(srcOffset); fr(srcSlot); fr(uniqueSrc); fr(weights); fr(posInMat); fr(matOffset);
  fr(matrices); fr(matricesT); fr(matOffsetT); fr(coord); fr(scratch); fr(tileTarget); fr(tileSrcStart); fr(tileSrc); fr(localSrc); fr(tileMatStart); fr(tileMat); fr(tgtMeta);
  nTargets = 0; nMatrices = 0; nUnique = 0; maxSrc = 0; nTiles = 0;
}

IntpSet &IntpSet::operator=(IntpSet &&o) noexcept {
  if (this != &o) {
    release();
    order = o.order; nTargets = o.nTargets; nMatrices = o.nMatrices; nUnique = o.nUnique;
    targets = o.targets; srcOffset = o.srcOffset; srcSlot = o.srcSlot; uniqueSrc = o.uniqueSrc;
    weights = o.weights; posInMat = o.posInMat; matOffset = o.matOffset; matrices = o.matrices;
    coord = o.coord; scratch = o.scratch; matricesT = o.matricesT; o.matricesT = nullptr;
    matOffsetT = o.matOffsetT; o.matOffsetT = nullptr;
    maxSrc = o.maxSrc; o.maxSrc = 0;
    nTiles = o.nTiles; o.nTiles = 0;
    tileTarget = o.tileTarget; tileSrcStart = o.tileSrcStart; tileSrc = o.tileSrc; localSrc = o.localSrc;
    o.tileTarget = nullptr; o.tileSrcStart = nullptr; o.tileSrc = nullptr; o.localSrc = nullptr;
    tileMatStart = o.tileMatStart; tileMat = o.tileMat; tgtMeta = o.tgtMeta;
    o.tileMatStart = nullptr; o.tileMat = nullptr; o.tgtMeta = nullptr;
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
  std::vector<int32_t> offT;   // offsets of the transposed matrices (least-square orders)
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
    {
      // the same matrices transposed, [s][k]: the nCoeff coefficients of a source side by side,
      // every matrix at an even offset (16-byte aligned pairs)
      const int nc = order == 1 ? 4 : 10;
      std::vector<double> mt;
      offT.assign((size_t)nMatrices + 1, 0);
      for (int mI = 0; mI < nMatrices; ++mI) {
        const int off = matOffset[mI], len = matOffset[mI + 1] - off;
        offT[mI] = (int32_t)mt.size();
        if (len % nc == 0) {                   // else: the 1 x 1 placeholder of a singular set, never referenced
          const int nS = len / nc;
          mt.resize(mt.size() + (size_t)len);
          for (int k = 0; k < nc; ++k)
            for (int sI = 0; sI < nS; ++sI)
              mt[(size_t)offT[mI] + (size_t)sI * nc + k] = matrices[(size_t)off + (size_t)k * nS + sI];
        }
        if (mt.size() & 1u) mt.push_back(0.0);
      }
      offT[nMatrices] = (int32_t)mt.size();
      if (mt.empty()) mt.push_back(0.0);
      rc |= up(set.matricesT, mt.data(), mt.size(), st);
      rc |= up(set.matOffsetT, offT.data(), offT.size(), st);
      MUSB_CUDA(cudaStreamSynchronize(st));
    }
    rc |= up(set.coord, childCoord, (size_t)3 * nTargets, st);
  }
  if (rc) return rc;
  set.nTargets = nTargets;
  set.maxSrc = 0;
  for (int i = 0; i < nTargets; ++i) set.maxSrc = std::max(set.maxSrc, srcOffset[i + 1] - srcOffset[i]);
  set.nMatrices = nMatrices;
  set.nUnique = (int)uniq.size();
  {
    // tiles: greedily pack consecutive targets while the union of their sources stays within
    // kTileSrc, the tile within kTileTgt targets and kTileEnt (target, source) entries, and its
    // distinct least-square matrices within kTileMat doubles
    const int nc = order == 0 ? 1 : (order == 1 ? 4 : 10);
    std::vector<int32_t> tileTarget{0}, tileSrcStart{0}, tileSrc, tileMatStart{0}, tileMat;
    std::vector<int32_t> meta((size_t)nTargets * 4, 0);
    std::vector<uint8_t> local((size_t)nSrc);
    std::vector<int32_t> where((size_t)set.nUnique, -1);   // slot -> index in the current tile
    std::vector<int32_t> cur;
    std::vector<std::pair<int32_t, int32_t>> curMat;        // (matrix id, offset in the staged area)
    int matDoubles = 0;
    auto flush = [&](int nextTarget) {
      for (int32_t sl : cur) where[sl] = -1;
      tileSrc.insert(tileSrc.end(), cur.begin(), cur.end());
      cur.clear();
      for (auto &m : curMat) {
        tileMat.push_back(offT[m.first]);
        tileMat.push_back(offT[m.first + 1] - offT[m.first]);
      }
      curMat.clear();
      matDoubles = 0;
      tileTarget.push_back(nextTarget);
      tileSrcStart.push_back((int32_t)tileSrc.size());
      tileMatStart.push_back((int32_t)tileMat.size() / 2);
    };
    for (int i = 0; i < nTargets; ++i) {
      int add = 0;
      for (int j = srcOffset[i]; j < srcOffset[i + 1]; ++j)
        if (where[slot[j]] < 0) ++add;
      int matAdd = 0, matId = -1;
      if (order > 0) {
        matId = posInMat[i];
        bool have = false;
        for (auto &m : curMat) have = have || m.first == matId;
        if (!have) matAdd = offT[matId + 1] - offT[matId];
      }
      if ((int)cur.size() + add > kTileSrc || i - tileTarget.back() >= kTileTgt ||
          srcOffset[i + 1] - srcOffset[tileTarget.back()] > kTileEnt || matDoubles + matAdd > kTileMat) {
        flush(i);
        if (order > 0) matAdd = offT[matId + 1] - offT[matId];
      }
      for (int j = srcOffset[i]; j < srcOffset[i + 1]; ++j) {
        if (where[slot[j]] < 0) { where[slot[j]] = (int32_t)cur.size(); cur.push_back(slot[j]); }
        local[j] = (uint8_t)where[slot[j]];
      }
      int matOff = 0;
      if (order > 0) {
        bool have = false;
        for (auto &m : curMat)
          if (m.first == matId) { have = true; matOff = m.second; }
        if (!have) { matOff = matDoubles; curMat.push_back({matId, matDoubles}); matDoubles += matAdd; }
      }
      meta[(size_t)i * 4 + 0] = srcOffset[i] - srcOffset[tileTarget.back()];
      meta[(size_t)i * 4 + 1] = srcOffset[i + 1] - srcOffset[i];
      meta[(size_t)i * 4 + 2] = matOff;
      meta[(size_t)i * 4 + 3] = targetList[i] - 1;
    }
    flush(nTargets);
    (void)nc;
    set.nTiles = (int)tileTarget.size() - 1;
    int rc2 = 0;
    rc2 |= up(set.tileTarget, tileTarget.data(), tileTarget.size(), st);
    rc2 |= up(set.tileSrcStart, tileSrcStart.data(), tileSrcStart.size(), st);
    rc2 |= up(set.tileSrc, tileSrc.data(), tileSrc.size(), st);
    rc2 |= up(set.localSrc, local.data(), local.size(), st);
    rc2 |= up(set.tileMatStart, tileMatStart.data(), tileMatStart.size(), st);
    if (tileMat.empty()) tileMat.push_back(0);
    rc2 |= up(set.tileMat, tileMat.data(), tileMat.size(), st);
    rc2 |= up(set.tgtMeta, meta.data(), meta.size(), st);
    if (rc2) return rc2;
    MUSB_CUDA(cudaStreamSynchronize(st));  // the tile vectors go out of scope
  }
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
__global__ void __launch_bounds__(THREADS) eqNeqKernel(int incomp, int passive, const double *__restrict__ sState,
                            const double *__restrict__ sAux, long long sS,
                            const int32_t *__restrict__ uniqueSrc, int nUnique,
                            double *__restrict__ scratch) {
  constexpr int W = 2 * QQ, P = W + 1;       // odd pitch: conflict-free column access
  __shared__ double tile[THREADS * P];
  const int u = blockIdx.x * THREADS + threadIdx.x;
  if (u < nUnique) {
    const int e = uniqueSrc[u] - 1;
    double rho = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
    if (!passive) { rho = sAux[e]; vx = sAux[sS + e]; vy = sAux[2 * sS + e]; vz = sAux[3 * sS + e]; }
    double feq[QQ];
    if (passive) {
      // arbitrary-value interpolation of the PDFs: the pair is {f, 0}
#pragma unroll
      for (int q = 0; q < QQ; ++q) feq[q] = sState[(long long)q * sS + e];
    } else if (QQ == 19) {
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
      tile[threadIdx.x * P + 2 * q + 1] = passive ? 0.0 : f - feq[q];
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

// The same interpolation organised around the SOURCES instead of the targets: the 8 children of a
// coarse parent draw their sources from one <= 19 / 27 element neighbourhood, and neighbouring
// parents share most of theirs.  A CTA takes a tile of consecutive targets (host-packed so that
// the tile has at most kTileSrc distinct sources), copies the {f_eq, f_neq} rows of those sources
// from the scratch into shared memory ONCE -- contiguous rows, fully coalesced -- and evaluates
// every (target, direction) of the tile from there: the per-thread chain of dependent
// slot -> value gathers through L1 (62 % of the L1 data pipe in the target-major kernel above,
// profiles/r01_intp_cfg4.md) becomes shared-memory reads, and the L2 -> SM traffic drops from
// nSrc rows per target to about one row per target.  Per (target, direction) the sources are
// accumulated in the host's order exactly as above, so the bits are unchanged.  Results go
// through shared memory once more so that the stores run along the targets (siblings are
// consecutive in the total list): 64-byte segments instead of 8-byte scattered stores.
template <int MODE, int QQ, int D>
__global__ void __launch_bounds__(128) intpTileKernel(const double *__restrict__ scratch,
                               const int32_t *__restrict__ tileTarget, const int32_t *__restrict__ tileSrcStart,
                               const int32_t *__restrict__ tileSrc, const uint8_t *__restrict__ localSrc,
                               const int32_t *__restrict__ tileMatStart, const int32_t *__restrict__ tileMat,
                               const int4 *__restrict__ tgtMeta, const int32_t *__restrict__ srcOffset,
                               const double *__restrict__ weights, const double *__restrict__ matricesT,
                               const double *__restrict__ coord, double *__restrict__ tState, long long tS,
                               const double *__restrict__ tVisc, double tViscUniform) {
  // Everything the evaluation reads is staged in shared memory first, each piece by independent,
  // contiguous loads (no dependent index chains): the {f_eq, f_neq} rows of the tile's sources,
  // the tile's DISTINCT least-square matrices (transposed, matricesT[s][k]: the coefficients of a
  // source side by side; a refinement surface uses a handful: one per child position), the row
  // index of every (target, source) entry and one int4 of metadata per target.  The inner loops
  // issue no global load; one thread evaluates D consecutive directions of a target.
  constexpr int nCoeff = MODE == 1 ? 1 : (MODE == 2 ? 4 : 10);
  constexpr int nG = (QQ + D - 1) / D;
  extern __shared__ double2 sm2[];                                   // [kTileSrc][QQ] pairs
  double *res = reinterpret_cast<double *>(sm2 + kTileSrc * QQ);      // [QQ][kTileTgt] results
  double *mats = res + kTileTgt * QQ;                                // staged matrices / MODE 1: weights per entry
  int4 *meta = reinterpret_cast<int4 *>(mats + (MODE == 1 ? kTileEnt : kTileMat));   // [kTileTgt]
  double *xyz = reinterpret_cast<double *>(meta + kTileTgt);         // [kTileTgt][4]: child coordinates, factor
  uint8_t *rowOf = reinterpret_cast<uint8_t *>(xyz + 4 * kTileTgt);  // [kTileEnt]
  const int t0 = tileTarget[blockIdx.x], nT = tileTarget[blockIdx.x + 1] - t0;
  const int u0 = tileSrcStart[blockIdx.x], nU = tileSrcStart[blockIdx.x + 1] - u0;
  const int e0 = srcOffset[t0], nE = srcOffset[t0 + nT] - e0;
  const double2 *sc2 = reinterpret_cast<const double2 *>(scratch);
  {
    // all of a thread's loads in flight at once (index, then row piece): the loop form ran its
    // <= 10 iterations back to back, two dependent memory latencies each -- 44 % of the kernel's
    // stall samples sat on the stores of this staging phase (profiles/r02_intp_cfg4.md)
    constexpr int ITER = (kTileSrc * QQ + 127) / 128;
    int srcRow[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int idx = threadIdx.x + it * 128;
      srcRow[it] = idx < nU * QQ ? tileSrc[u0 + idx / QQ] : -1;
    }
    double2 v[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int idx = threadIdx.x + it * 128;
      if (srcRow[it] >= 0) v[it] = sc2[(long long)srcRow[it] * QQ + (idx % QQ)];
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int idx = threadIdx.x + it * 128;
      if (srcRow[it] >= 0) sm2[idx] = v[it];
    }
  }
  {
    constexpr int ITER = kTileEnt / 128;
    uint8_t b[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int e = threadIdx.x + it * 128;
      b[it] = e < nE ? localSrc[e0 + e] : 0;
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int e = threadIdx.x + it * 128;
      if (e < nE) rowOf[e] = b[it];
    }
  }
  if (threadIdx.x < nT) {
    const int4 mt = tgtMeta[t0 + threadIdx.x];
    meta[threadIdx.x] = mt;
    double fac;   // 0.5 * getNonEqFac_intp_coarse_to_fine
    if (tVisc) {
      const double visc = tVisc[mt.w];
      fac = 0.5 * neqFac(omegaFromVisc(0.5 * visc), omegaFromVisc(visc));
    } else {
      fac = tViscUniform;   // the factor itself, evaluated by the launcher with the same expression
    }
    xyz[4 * threadIdx.x + 3] = fac;
  }
  if (MODE != 1)
    for (int k = threadIdx.x; k < 3 * nT; k += blockDim.x) xyz[4 * (k / 3) + k % 3] = coord[3 * t0 + k];
  if (MODE == 1) {
    constexpr int ITER = kTileEnt / 128;
    double w[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int e = threadIdx.x + it * 128;
      w[it] = e < nE ? weights[e0 + e] : 0.0;
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int e = threadIdx.x + it * 128;
      if (e < nE) mats[e] = w[it];
    }
  } else {
    // the tile's distinct matrices lie back to back in tileMat's order: thread k of the staged
    // area finds its matrix by walking the (few) lengths, no dependent loads between matrices
    const int m0 = tileMatStart[blockIdx.x], m1 = tileMatStart[blockIdx.x + 1];
    constexpr int ITER = kTileMat / 128;
    double c[ITER];
    int total = 0;
    for (int m = m0; m < m1; ++m) total += tileMat[2 * m + 1];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int k = threadIdx.x + it * 128;
      c[it] = 0.0;
      if (k < total) {
        int off = 0, m = m0;
        while (k >= off + tileMat[2 * m + 1]) { off += tileMat[2 * m + 1]; ++m; }
        c[it] = matricesT[tileMat[2 * m] + (k - off)];
      }
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int k = threadIdx.x + it * 128;
      if (k < total) mats[k] = c[it];
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < nT * nG; p += blockDim.x) {
    const int il = p / nG, d0 = (p - il * nG) * D;
    const int nd = min(D, QQ - d0);
    const int4 mt = meta[il];                     // entry offset, nSrc, matrix offset, target
    const int s0 = mt.x, n = mt.y;
    const double *A = MODE == 1 ? mats + s0 : mats + mt.z;
    double ce[D][nCoeff], cn[D][nCoeff];
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
      for (int k = 0; k < nCoeff; ++k) { ce[j][k] = 0.0; cn[j][k] = 0.0; }
    for (int s = 0; s < n; ++s) {
      double m[nCoeff];
#pragma unroll
      for (int k = 0; k < nCoeff; ++k) m[k] = A[s * nCoeff + k];
      const double2 *row = sm2 + (int)rowOf[s0 + s] * QQ + d0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        if (j < nd) {
          const double2 v = row[j];
#pragma unroll
          for (int k = 0; k < nCoeff; ++k) {   // per coefficient the sum over the sources, in the host's order
            ce[j][k] = ce[j][k] + m[k] * v.x;
            cn[j][k] = cn[j][k] + m[k] * v.y;
          }
        }
      }
    }
    const double x = xyz[4 * il], y = xyz[4 * il + 1], z = xyz[4 * il + 2], fac = xyz[4 * il + 3];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (j < nd) {
        double t_eq, t_neq;
        if (MODE == 1) {
          t_eq = ce[j][0];
          t_neq = cn[j][0];
        } else {
          t_eq = ce[j][0] + ce[j][1] * x + ce[j][2] * y + ce[j][3] * z;
          t_neq = cn[j][0] + cn[j][1] * x + cn[j][2] * y + cn[j][3] * z;
          if (MODE == 3) {
            t_eq = t_eq + ce[j][4 % nCoeff] * x * x + ce[j][5 % nCoeff] * y * y + ce[j][6 % nCoeff] * z * z +
                   ce[j][7 % nCoeff] * x * y + ce[j][8 % nCoeff] * y * z + ce[j][9 % nCoeff] * z * x;
            t_neq = t_neq + cn[j][4 % nCoeff] * x * x + cn[j][5 % nCoeff] * y * y + cn[j][6 % nCoeff] * z * z +
                    cn[j][7 % nCoeff] * x * y + cn[j][8 % nCoeff] * y * z + cn[j][9 % nCoeff] * z * x;
          }
        }
        t_neq = t_neq * fac;
        res[(d0 + j) * kTileTgt + il] = t_neq + t_eq;
      }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < nT * QQ; p += blockDim.x) {
    const int d = p / nT, il = p - d * nT;
    tState[(long long)d * tS + meta[il].w] = res[d * kTileTgt + il];
  }
}

// fillMyGhostsFromFiner_avg_feq_fneq without a scratch pass: 8 lanes per coarse ghost, one lane per
// child (the children are Morton-contiguous fine elements: coalesced loads).  Each lane forms
// f_eq(rho, u from auxField) and f - f_eq of its child and parks them in shared memory; the lanes
// of the group then share the QQ directions, each summing its directions over the children in
// the children's order, a = (..((c1 + c2) + c3)..), as the reference's loop does.
template <int QQ, int THREADS>
__global__ void __launch_bounds__(THREADS) fromFinerFusedKernel(int incomp, int passive, const double *__restrict__ sState,
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
    double rho = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
    if (!passive) { rho = sAux[e]; vx = sAux[sS + e]; vy = sAux[2 * sS + e]; vz = sAux[3 * sS + e]; }
    double f[QQ], eq[QQ];
#pragma unroll
    for (int q = 0; q < QQ; ++q) f[q] = sState[(long long)q * sS + e];
    if (passive) {
      // fillArbiMyGhostsFromFiner_avg applied to the PDFs: plain average, {f, 0}
#pragma unroll
      for (int q = 0; q < QQ; ++q) eq[q] = f[q];
    } else if (QQ == 19) {
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
      row[QQ + q] = passive ? 0.0 : f[q] - eq[q];
    }
    row[W] = rho; row[W + 1] = vx; row[W + 2] = vy; row[W + 3] = vz;
  }
  __syncthreads();
  if (!valid) return;
  const double visc = tVisc ? tVisc[tgt] : tViscUniform;
  const double inv_n = 1.0 / (double)n;
  const double fOmega = omegaFromVisc(2.0 * visc), cOmega = omegaFromVisc(visc);
  const double fac = passive ? 0.0 : 2.0 * neqFac(fOmega, cOmega);  // getNonEqFac_intp_fine_to_coarse
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

// 1: the target-major kernel (kept for comparison and as the reference of the tiled one)
int g_intpTargetMajor = 0;

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
          a.incomp, a.passive ? 1 : 0, a.sState, a.sAux, a.sS, set.uniqueSrc, set.nTargets, set.targets, set.srcOffset,
          set.srcSlot, a.tState, a.withAux ? a.tAux : nullptr, a.tS, a.tVisc, a.tViscUniform);
    else              // 64 threads x 59 doubles = 30 KB
      fromFinerFusedKernel<27, 64><<<divUp((long long)set.nTargets * 8, 64), 64, 0, st>>>(
          a.incomp, a.passive ? 1 : 0, a.sState, a.sAux, a.sS, set.uniqueSrc, set.nTargets, set.targets, set.srcOffset,
          set.srcSlot, a.tState, a.withAux ? a.tAux : nullptr, a.tS, a.tVisc, a.tViscUniform);
    MUSB_CUDA(cudaGetLastError());
    if (nLaunch) *nLaunch = 1;
    return 0;
  }
  if (a.QQ == 19)
    eqNeqKernel<19, 128><<<divUp(set.nUnique, 128), 128, 0, st>>>(a.incomp, a.passive ? 1 : 0, a.sState, a.sAux, a.sS,
                                                                  set.uniqueSrc, set.nUnique, set.scratch);
  else
    eqNeqKernel<27, 64><<<divUp(set.nUnique, 64), 64, 0, st>>>(a.incomp, a.passive ? 1 : 0, a.sState, a.sAux, a.sS,
                                                                set.uniqueSrc, set.nUnique, set.scratch);
  MUSB_CUDA(cudaGetLastError());
  const int mode = 1 + set.order;
  if (mode == 1 && !set.weights) return setError(1, "weighted-average set without weights");
  const int grid = divUp((long long)set.nTargets * a.QQ, B);
  // uniform viscosity: hand over 0.5 * getNonEqFac_intp_coarse_to_fine instead of the viscosity
  const double facOrVisc = (a.tVisc || a.passive)
      ? 0.0 : 0.5 * hostNeqFac(hostOmega(0.5 * a.tViscUniform), hostOmega(a.tViscUniform));
  const double *tViscArr = a.passive ? nullptr : a.tVisc;   // passive scalar: factor 0 on the zero f_neq
  if (set.nTiles > 0 && !g_intpTargetMajor) {
    const size_t smem = (size_t)kTileSrc * a.QQ * sizeof(double2) + (size_t)kTileTgt * a.QQ * sizeof(double) +
                        (size_t)(mode == 1 ? kTileEnt : kTileMat) * sizeof(double) + (size_t)kTileTgt * sizeof(int4) +
                        (size_t)4 * kTileTgt * sizeof(double) + (size_t)kTileEnt;
#define MUSB_TILE(M, Q, D)                                                                                   \
  {                                                                                                          \
    static bool optIn = false;   /* more than 48 KB of dynamic shared memory needs the attribute, once */     \
    if (!optIn && smem > 48 * 1024) {                                                                        \
      MUSB_CUDA(cudaFuncSetAttribute(intpTileKernel<M, Q, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                     (int)smem));                                                            \
      optIn = true;                                                                                          \
    }                                                                                                        \
    intpTileKernel<M, Q, D><<<set.nTiles, 128, smem, st>>>(                                                  \
        set.scratch, set.tileTarget, set.tileSrcStart, set.tileSrc, set.localSrc, set.tileMatStart,          \
        set.tileMat, reinterpret_cast<const int4 *>(set.tgtMeta), set.srcOffset, set.weights, set.matricesT, \
        set.coord, a.tState, a.tS, tViscArr, facOrVisc);                                                     \
  }
    if (a.QQ == 19) {
      if (mode == 1) MUSB_TILE(1, 19, 4)
      else if (mode == 2) MUSB_TILE(2, 19, 2)
      else MUSB_TILE(3, 19, 2)
    } else {
      if (mode == 1) MUSB_TILE(1, 27, 4)
      else if (mode == 2) MUSB_TILE(2, 27, 2)
      else MUSB_TILE(3, 27, 2)
    }
#undef MUSB_TILE
    MUSB_CUDA(cudaGetLastError());
    if (nLaunch) *nLaunch = 2;
    return 0;
  }
#define MUSB_INTP(M)                                                                              \
  intpKernel<M><<<grid, B, 0, st>>>(a.QQ, set.scratch, set.nUnique, set.nTargets, set.targets,    \
                                    set.srcOffset, set.srcSlot, set.weights, set.posInMat,        \
                                    set.matOffset, set.matrices, set.coord, a.tState, a.tS,       \
                                    tViscArr, facOrVisc)
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
