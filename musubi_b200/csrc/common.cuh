// common.cuh -- stencil tables and small helpers shared by the device code.
//
// Direction order and inverse directions: tem/source/tem_stencil_module.fpp:91-168
// (rest direction LAST); weights: mus/source/scheme/mus_scheme_layout_module.f90:699-705.
// Directions are 0-based here (reference: 1-based), so q = iDir - 1.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace musb200 {

// D3Q27 table; D3Q19 is its first 18 rows + rest.
__host__ __device__ constexpr int cxTab(int q, int k) {
  constexpr int t[26][3] = {
      {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
      {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1},
      {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1}, {1, 0, 1},
      {-1, -1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0},
      {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1},
      {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}};
  return t[q][k];
}

template <int QQ>
__host__ __device__ constexpr int cx(int q, int k) {
  return (q == QQ - 1) ? 0 : cxTab(q, k);
}

// inverse direction (0-based) of q
template <int QQ>
__host__ __device__ constexpr int invDir(int q) {
  if (q == QQ - 1) return q;
  for (int j = 0; j < QQ - 1; ++j)
    if (cxTab(j, 0) == -cxTab(q, 0) && cxTab(j, 1) == -cxTab(q, 1) && cxTab(j, 2) == -cxTab(q, 2))
      return j;
  return -1;
}

template <int QQ>
__host__ __device__ constexpr double weight(int q) {
  const int n = cx<QQ>(q, 0) * cx<QQ>(q, 0) + cx<QQ>(q, 1) * cx<QQ>(q, 1) + cx<QQ>(q, 2) * cx<QQ>(q, 2);
  if (QQ == 19) return n == 0 ? 1.0 / 3.0 : (n == 1 ? 1.0 / 18.0 : 1.0 / 36.0);
  return n == 0 ? 8.0 / 27.0 : (n == 1 ? 2.0 / 27.0 : (n == 2 ? 1.0 / 54.0 : 1.0 / 216.0));
}

// 0-based direction names (mus/source/mus_directions_module.f90:10-35 minus one)
enum Dir : int {
  N00 = 0, ZN0, ZZN, P00, ZP0, ZZP, ZNN, ZNP, ZPN, ZPP,
  NZN, PZN, NZP, PZP, NN0, NP0, PN0, PP0,
  NNN, NNP, NPN, NPP, PNN, PNP, PPN, PPP
};

// neighbour-list encoding on the device: bit 31 = bounce-back (pull the element's
// own inverse direction slot), bits 0..30 = 0-based source element.
constexpr uint32_t kBounceBit = 0x80000000u;
constexpr uint32_t kElemMask = 0x7fffffffu;

std::string &lastError();
int setError(int code, const std::string &msg);

#define MUSB_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t err__ = (call);                                                            \
    if (err__ != cudaSuccess)                                                              \
      return ::musb200::setError(2, std::string(#call) + ": " + cudaGetErrorString(err__)); \
  } while (0)

inline int divUp(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace musb200
