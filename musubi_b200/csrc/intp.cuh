// intp.cuh -- coarse <-> fine ghost interpolation sets (device side).
// fillMyGhostsFromFiner_avg_feq_fneq   mus/source/intp/mus_interpolate_average_module.fpp:186-347
// fillFinerGhostsFromMe_{weighAvg,linear,quad}_feq_fneq
//     average:854-1038, linear:315-505, quadratic:292-...
#pragma once
#include "common.cuh"

namespace musb200 {

struct IntpSet {
  int order = 0;
  int nTargets = 0;
  int nMatrices = 0;
  int32_t *targets = nullptr;    // [nTargets] 0-based target element
  int32_t *srcOffset = nullptr;  // [nTargets+1]
  int32_t *srcPos = nullptr;     // CSR, 0-based source element
  double *weights = nullptr;     // CSR (average / weighted average)
  int32_t *posInMat = nullptr;   // [nTargets] matrix index (linear / quadratic)
  int32_t *matOffset = nullptr;  // [nMatrices+1] offsets into matrices
  double *matrices = nullptr;    // concatenated (nCoeff x nSrc) row-major LSQ matrices
  double *childCoord = nullptr;  // [nTargets][3] child offset in coarse units (+-0.25)
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
  long long tS;
  const double *tOmega;  // per-element omega of the target level or nullptr
  double tOmegaUniform;
};

int registerIntp(IntpSet &set, int order, int nTargets, const int32_t *targetList,
                 const int32_t *srcOffset, const int32_t *srcPos, const double *weights,
                 const int32_t *posInMat, int nMatrices, const int32_t *matOffset,
                 const double *matrices, const double *childCoord, cudaStream_t st);
int launchIntp(const IntpArgs &a, const IntpSet &set, bool fromFiner, cudaStream_t st);

}  // namespace musb200
