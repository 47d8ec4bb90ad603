// intp.cuh -- coarse <-> fine ghost interpolation on the device.
//
// Replaces the pointees of intp%fillMineFromFiner%do_intp / do_intpArbiVal and
// intp%fillFinerFromMe(order)%do_intp (mus_interpolate_header_module.f90:103-237):
//   fillMyGhostsFromFiner_avg_feq_fneq       mus_interpolate_average_module.fpp:186-347
//   fillArbiMyGhostsFromFiner_avg            mus_interpolate_average_module.fpp:95-185 (auxField,
//                                            taken inside the from-finer kernel)
//   fillFinerGhostsFromMe_weighAvg_feq_fneq  mus_interpolate_average_module.fpp:854-1038
//   fillFinerGhostsFromMe_linear_feq_fneq    mus_interpolate_linear_module.fpp:315-505
//   fillFinerGhostsFromMe_quad_feq_fneq      mus_interpolate_quadratic_module.fpp:292-...
// The dependency lists, weights and least-square matrices come from the host
// (levelDesc%depFromFiner / depFromCoarser, intpMat_forLSF) unchanged.
//
// Coarse -> fine, two phases per set, both on the refinement surface only:
//   A  one thread per distinct source element: f_eq(rho,u from auxField), f_neq = f - f_eq
//   B  one thread per (target, direction), direction fastest: weighted sum / least-square
//      polynomial over the sources in the host's order, non-equilibrium rescaling, store.
// Fine -> coarse: one fused kernel, 8 lanes per coarse ghost (one per child).
#pragma once
#include "common.cuh"
#include <vector>

namespace musb200 {

constexpr int kTileSrc = 64;   // distinct sources per tile
constexpr int kTileTgt = 64;   // targets per tile
constexpr int kTileEnt = 512;  // (target, source) entries per tile
constexpr int kTileMat = 1024; // doubles of distinct least-square matrices per tile

struct IntpSet {
  int order = 0;
  int nTargets = 0;
  int nMatrices = 0;
  int nUnique = 0;
  int32_t *targets = nullptr;    // [nTargets] 1-based position of the target in its total list
  int32_t *srcOffset = nullptr;  // [nTargets+1] CSR
  int32_t *srcSlot = nullptr;    // CSR: index into the unique-source scratch
  int32_t *uniqueSrc = nullptr;  // [nUnique] 1-based source positions
  double *weights = nullptr;     // CSR (weighted average)
  int32_t *posInMat = nullptr;   // [nTargets]
  int32_t *matOffset = nullptr;  // [nMatrices+1]
  double *matrices = nullptr;    // concatenated row-major (nCoeff x nSrc)
  double *matricesT = nullptr;   // the same, each matrix transposed (nSrc x nCoeff): tile kernel
  int32_t *matOffsetT = nullptr; // [nMatrices+1] into matricesT (even offsets)
  double *coord = nullptr;       // [nTargets][3]
  double *scratch = nullptr;     // [nUnique][2*QQ]  f_eq | f_neq per distinct source
  int maxSrc = 0;                // largest number of sources of a target
  // from-coarser sets: consecutive targets (treeID order: siblings, then neighbouring parents)
  // packed into CTA tiles whose sources -- at most kTileSrc distinct ones -- are staged in shared
  // memory once and shared by all targets of the tile
  int nTiles = 0;
  int32_t *tileTarget = nullptr; // [nTiles+1] first target of each tile
  int32_t *tileSrcStart = nullptr; // [nTiles+1] into tileSrc
  int32_t *tileSrc = nullptr;    // distinct source slots of each tile, concatenated
  uint8_t *localSrc = nullptr;   // CSR like srcSlot: index into the tile's source list
  int32_t *tileMatStart = nullptr; // [nTiles+1] into tileMat
  int32_t *tileMat = nullptr;    // per tile: (offset in matricesT, length) of its distinct matrices
  int32_t *tgtMeta = nullptr;    // [nTargets][4]: entry offset in the tile, nSrc, offset of its matrix in the
                                 // tile's staged matrices (MODE 1: unused), target position - 1
  void release();
  ~IntpSet() { release(); }
  IntpSet() = default;
  IntpSet(const IntpSet &) = delete;
  IntpSet &operator=(const IntpSet &) = delete;
  IntpSet(IntpSet &&o) noexcept { *this = std::move(o); }
  IntpSet &operator=(IntpSet &&o) noexcept;
};

struct IntpArgs {
  int QQ;
  int incomp;
  const double *sState;  // source level state(:, next), SoA
  const double *sAux;    // source level auxField, SoA [4][sS]
  long long sS;
  double *tState;        // target level state(:, next)
  double *tAux;          // target level auxField (aux averaging only)
  long long tS;
  const double *tVisc;   // per-element lattice viscosity of the target level or nullptr
  double tViscUniform;
  bool withAux;          // from-finer: average the auxField in the same kernel
  bool passive;          // passive scalar: interpolate the PDFs themselves (f_eq := f, f_neq := 0)
};

int registerIntp(IntpSet &set, int order, int nTargets, const int32_t *targetList,
                 const int32_t *srcOffset, const int32_t *srcPos, const double *weights,
                 const int32_t *posInMat, int nMatrices, const int32_t *matOffset,
                 const double *matrices, const double *childCoord, cudaStream_t st);
// returns the number of kernels launched through *nLaunch
extern int g_intpTargetMajor;
int launchIntp(const IntpArgs &a, const IntpSet &set, bool fromFiner, cudaStream_t st, int *nLaunch);

}  // namespace musb200
