// cube.cu -- treelm's predefined cube mesh built on the device, and the equilibrium initial state.
//
//  * generate_treelm_cube / tem_load_internal (tem/source/treelmesh_module.f90:1224-1318): the
//    mesh `predefined = 'cube'` holds all 8^L elements of level L in Morton order (x = bit 0,
//    y = bit 1, z = bit 2 of every octal digit, tem_topology_module.f90:590-638), fully periodic
//    through the wrap at the universe cube, or closed by walls on its six faces.
//  * mus_construct_connectivity (mus/source/mus_connectivity_module.fpp:113-177) for that mesh:
//    direction q pulls from the element at x - c_q; behind a wall from the element's own inverse
//    direction (bounce-back).
//  Why on the device: the host form of the list, neigh(QQ * nSize) as 32-bit state positions,
//  ends at nSize * QQ < 2^31 -- 79 M elements for D3Q27 -- while one B200 holds the 512^3 = 134 M
//  elements of BASELINE config 3 (76 GB).  The encoded device list has no such limit (element
//  index < 2^31), and 14.5 GB of index list never cross PCIe.
//  * mus_init_pdf with zero strain rate (mus/source/mus_flow_module.fpp:484-589): f = f_eq(rho, u)
//    from the auxField rows, written to both state buffers.
#include "equilibrium.cuh"
#include "kernels.cuh"

namespace musb200 {

__device__ __forceinline__ uint32_t compact3(uint32_t v) {   // every third bit of a 30-bit code
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u;
  v = (v | (v >> 4)) & 0x0300F00Fu;
  v = (v | (v >> 8)) & 0x030000FFu;
  v = (v | (v >> 16)) & 0x000003FFu;
  return v;
}
__device__ __forceinline__ uint32_t spread3(uint32_t v) {
  v &= 0x000003FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

template <int QQ>
__global__ void cubeNeighKernel(uint32_t *__restrict__ nbr, int level, int walls, long long S, int nElems) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nElems) return;
  const int n = 1 << level;
  const int x = (int)compact3((uint32_t)e), y = (int)compact3((uint32_t)e >> 1), z = (int)compact3((uint32_t)e >> 2);
#pragma unroll
  for (int q = 0; q < QQ - 1; ++q) {
    int xs = x - cx<QQ>(q, 0), ys = y - cx<QQ>(q, 1), zs = z - cx<QQ>(q, 2);
    uint32_t w;
    if (walls && (xs < 0 || xs >= n || ys < 0 || ys >= n || zs < 0 || zs >= n)) {
      w = (uint32_t)e | kBounceBit;
    } else {
      xs = (xs + n) & (n - 1); ys = (ys + n) & (n - 1); zs = (zs + n) & (n - 1);
      w = spread3((uint32_t)xs) | (spread3((uint32_t)ys) << 1) | (spread3((uint32_t)zs) << 2);
    }
    nbr[(long long)q * S + e] = w;
  }
}

int launchCubeNeigh(int QQ, uint32_t *nbr, int level, int walls, long long S, int nElems, cudaStream_t st) {
  const int grid = divUp(nElems, 256);
  if (QQ == 19) cubeNeighKernel<19><<<grid, 256, 0, st>>>(nbr, level, walls, S, nElems);
  else cubeNeighKernel<27><<<grid, 256, 0, st>>>(nbr, level, walls, S, nElems);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

template <int QQ, bool INCOMP>
__global__ void __launch_bounds__(128) initEquilibriumKernel(const double *__restrict__ aux, double *__restrict__ s0,
                                                             double *__restrict__ s1, long long S, int nElems) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nElems) return;
  const double rho = aux[e], vx = aux[S + e], vy = aux[2 * S + e], vz = aux[3 * S + e];
  double feq[QQ];
  if (QQ == 19) {
    double(&g)[19] = reinterpret_cast<double(&)[19]>(feq);
    if (INCOMP) pdfEqIncompD3Q19(rho, vx, vy, vz, g);
    else pdfEqD3Q19(rho, vx, vy, vz, g);
  } else {
    double(&g)[27] = reinterpret_cast<double(&)[27]>(feq);
    if (INCOMP) pdfEqIncompD3Q27(rho, vx, vy, vz, g);
    else pdfEqD3Q27(rho, vx, vy, vz, g);
  }
#pragma unroll
  for (int q = 0; q < QQ; ++q) {
    s0[(long long)q * S + e] = feq[q];
    s1[(long long)q * S + e] = feq[q];
  }
}

int launchInitEquilibrium(int QQ, int incomp, const double *aux, double *s0, double *s1, long long S, int nElems,
                          cudaStream_t st) {
  if (nElems <= 0) return 0;
  const int grid = divUp(nElems, 128);
  if (QQ == 19 && !incomp) initEquilibriumKernel<19, false><<<grid, 128, 0, st>>>(aux, s0, s1, S, nElems);
  else if (QQ == 19) initEquilibriumKernel<19, true><<<grid, 128, 0, st>>>(aux, s0, s1, S, nElems);
  else if (!incomp) initEquilibriumKernel<27, false><<<grid, 128, 0, st>>>(aux, s0, s1, S, nElems);
  else initEquilibriumKernel<27, true><<<grid, 128, 0, st>>>(aux, s0, s1, S, nElems);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace musb200
