// layout.cu -- conversion between the host's Fortran arrays and the device layout.
//
//  * state / auxField: AOS  IDX(dir,elem) = (elem-1)*nScalars + dir
//    (mus/source/header/lbm_macros.inc:104)  <->  SoA rows of stride S.
//  * neigh: pdf%neigh(NGPOS(dir,elem)) holding 1-based AOS state positions
//    (mus/source/mus_connectivity_module.fpp:113-177)  <->  encoded uint32 list
//    (element index + bounce-back bit, direction implicit).
//  * pack / unpack of halo buffers through the position lists of
//    tem_communication_type (tem/source/tem_comm_module.fpp:549-646).
//  * check_density style reduction (mus/source/mus_tools_module.f90:224-313).
#include "kernels.cuh"

namespace musb200 {

// ---------------------------------------------------------------------------
// AOS <-> SoA through a shared-memory tile so that both sides are coalesced
constexpr int kTileElems = 64;

__global__ void aosToSoaKernel(const double *__restrict__ aos, double *__restrict__ soa, int nComp,
                               int nElems, long long S) {
  extern __shared__ double tile[];  // [kTileElems][nComp+1]
  const int e0 = blockIdx.x * kTileElems;
  const int ne = min(kTileElems, nElems - e0);
  const int pitch = nComp + 1;
  const long long base = (long long)e0 * nComp;
  for (int i = threadIdx.x; i < ne * nComp; i += blockDim.x)
    tile[(i / nComp) * pitch + (i % nComp)] = aos[base + i];
  __syncthreads();
  for (int i = threadIdx.x; i < ne * nComp; i += blockDim.x) {
    const int c = i / ne, e = i % ne;
    soa[(long long)c * S + e0 + e] = tile[e * pitch + c];
  }
}

__global__ void soaToAosKernel(const double *__restrict__ soa, double *__restrict__ aos, int nComp,
                               int nElems, long long S) {
  extern __shared__ double tile[];
  const int e0 = blockIdx.x * kTileElems;
  const int ne = min(kTileElems, nElems - e0);
  const int pitch = nComp + 1;
  for (int i = threadIdx.x; i < ne * nComp; i += blockDim.x) {
    const int c = i / ne, e = i % ne;
    tile[e * pitch + c] = soa[(long long)c * S + e0 + e];
  }
  __syncthreads();
  const long long base = (long long)e0 * nComp;
  for (int i = threadIdx.x; i < ne * nComp; i += blockDim.x)
    aos[base + i] = tile[(i / nComp) * pitch + (i % nComp)];
}

int launchAosToSoa(const double *aos, double *soa, int nComp, int nElems, long long S,
                   cudaStream_t st) {
  if (nElems <= 0) return 0;
  const size_t smem = (size_t)kTileElems * (nComp + 1) * sizeof(double);
  aosToSoaKernel<<<divUp(nElems, kTileElems), 256, smem, st>>>(aos, soa, nComp, nElems, S);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchSoaToAos(const double *soa, double *aos, int nComp, int nElems, long long S,
                   cudaStream_t st) {
  if (nElems <= 0) return 0;
  const size_t smem = (size_t)kTileElems * (nComp + 1) * sizeof(double);
  soaToAosKernel<<<divUp(nElems, kTileElems), 256, smem, st>>>(soa, aos, nComp, nElems, S);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

// per-element source data handed over as a list (force field, transport velocity):
// soa[k][pos[i]-1] = aos[i*nComp + k]
__global__ void scatterRowsKernel(const double *__restrict__ aos, const int32_t *__restrict__ pos, int n,
                                  int nComp, double *__restrict__ soa, long long S) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * nComp) return;
  const int i = t / nComp, k = t % nComp;
  const int e = pos ? pos[i] - 1 : i;
  soa[(long long)k * S + e] = aos[t];
}
int launchScatterRows(const double *aos, const int32_t *pos, int n, int nComp, double *soa, long long S,
                      cudaStream_t st) {
  if (n <= 0) return 0;
  scatterRowsKernel<<<divUp((long long)n * nComp, 256), 256, 0, st>>>(aos, pos, n, nComp, soa, S);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
template <int QQ>
__global__ void encodeNeighKernel(const int32_t *__restrict__ neigh, uint32_t *__restrict__ nbr,
                                  int nSize, int nElems, long long S, int *bad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nElems) return;
  int nbad = 0;
#pragma unroll
  for (int q = 0; q < QQ; ++q) {
    const int pos = neigh[(long long)q * nSize + e] - 1;  // 0-based AOS position
    const int src = pos / QQ, dir = pos % QQ;
    if (q == QQ - 1) {
      if (src != e || dir != q) ++nbad;  // rest direction must pull from itself
    } else if (pos < 0 || src >= nSize) {
      ++nbad;
      nbr[(long long)q * S + e] = (uint32_t)e;
    } else if (dir == q) {
      nbr[(long long)q * S + e] = (uint32_t)src;
    } else if (dir == invDir<QQ>(q) && src == e) {
      nbr[(long long)q * S + e] = (uint32_t)src | kBounceBit;
    } else {
      ++nbad;
      nbr[(long long)q * S + e] = (uint32_t)e;
    }
  }
  if (nbad) atomicAdd(bad, nbad);
}

template <int QQ>
__global__ void decodeNeighKernel(const uint32_t *__restrict__ nbr, int32_t *__restrict__ neigh,
                                  int nSize, int nElems, long long S) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nElems) return;
#pragma unroll
  for (int q = 0; q < QQ - 1; ++q) {
    const uint32_t n = nbr[(long long)q * S + e];
    const int dir = (n & kBounceBit) ? invDir<QQ>(q) : q;
    neigh[(long long)q * nSize + e] = (int)(n & kElemMask) * QQ + dir + 1;
  }
  neigh[(long long)(QQ - 1) * nSize + e] = e * QQ + QQ;
}

int launchEncodeNeigh(int QQ, const int32_t *neigh, uint32_t *nbr, int nSize, int nElems,
                      long long S, int *bad, cudaStream_t st) {
  if (nElems <= 0) return 0;
  if (QQ == 19)
    encodeNeighKernel<19><<<divUp(nElems, 256), 256, 0, st>>>(neigh, nbr, nSize, nElems, S, bad);
  else
    encodeNeighKernel<27><<<divUp(nElems, 256), 256, 0, st>>>(neigh, nbr, nSize, nElems, S, bad);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchDecodeNeigh(int QQ, const uint32_t *nbr, int32_t *neigh, int nSize, int nElems,
                      long long S, cudaStream_t st) {
  if (nElems <= 0) return 0;
  if (QQ == 19)
    decodeNeighKernel<19><<<divUp(nElems, 256), 256, 0, st>>>(nbr, neigh, nSize, nElems, S);
  else
    decodeNeighKernel<27><<<divUp(nElems, 256), 256, 0, st>>>(nbr, neigh, nSize, nElems, S);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// halo buffers: buf(i) = state(pos(i)) / state(pos(i)) = buf(i)
__global__ void packKernel(const double *__restrict__ state, long long S, int QQ,
                           const int32_t *__restrict__ pos, int n, double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = pos[i] - 1;
  buf[i] = state[(long long)(p % QQ) * S + p / QQ];
}

__global__ void unpackKernel(double *__restrict__ state, long long S, int QQ,
                             const int32_t *__restrict__ pos, int n,
                             const double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = pos[i] - 1;
  state[(long long)(p % QQ) * S + p / QQ] = buf[i];
}

int launchPack(int QQ, const double *state, long long S, const int32_t *pos, int n, double *buf,
               cudaStream_t st) {
  if (n <= 0) return 0;
  packKernel<<<divUp(n, 256), 256, 0, st>>>(state, S, QQ, pos, n, buf);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchUnpack(int QQ, double *state, long long S, const int32_t *pos, int n, const double *buf,
                 cudaStream_t st) {
  if (n <= 0) return 0;
  unpackKernel<<<divUp(n, 256), 256, 0, st>>>(state, S, QQ, pos, n, buf);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// restart bridge: mus_pdf_serialize / mus_pdf_unserialize (mus_buffer_module.fpp:80-190) for the
// elements of one level inside a chunk of the global treeID list:
//   buffer((slot(i)-1)*QQ + c) = state(level)%val(IDX(c, levelPointer(i)), nNext)
// slot = position of the element in the chunk (1-based), elemPos = levelPointer
__global__ void serializeKernel(const double *__restrict__ state, long long S, int QQ,
                                const int32_t *__restrict__ slot, const int32_t *__restrict__ elemPos,
                                int n, double *__restrict__ buffer) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * QQ) return;
  const int c = (int)(t / n), i = (int)(t % n);   // element index fastest: coalesced row reads
  buffer[(long long)(slot[i] - 1) * QQ + c] = state[(long long)c * S + (elemPos[i] - 1)];
}
__global__ void unserializeKernel(double *__restrict__ state, long long S, int QQ,
                                  const int32_t *__restrict__ slot, const int32_t *__restrict__ elemPos,
                                  int n, const double *__restrict__ buffer) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * QQ) return;
  const int c = (int)(t / n), i = (int)(t % n);
  state[(long long)c * S + (elemPos[i] - 1)] = buffer[(long long)(slot[i] - 1) * QQ + c];
}
int launchSerialize(int QQ, const double *state, long long S, const int32_t *slot, const int32_t *elemPos,
                    int n, double *buffer, cudaStream_t st) {
  if (n <= 0) return 0;
  serializeKernel<<<divUp((long long)n * QQ, 256), 256, 0, st>>>(state, S, QQ, slot, elemPos, n, buffer);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}
int launchUnserialize(int QQ, double *state, long long S, const int32_t *slot, const int32_t *elemPos,
                      int n, const double *buffer, cudaStream_t st) {
  if (n <= 0) return 0;
  unserializeKernel<<<divUp((long long)n * QQ, 256), 256, 0, st>>>(state, S, QQ, slot, elemPos, n, buffer);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// total mass / max |u|^2 / NaN count over the fluid elements; two-stage and
// deterministic (fixed grid, fixed summation tree).
constexpr int kRedBlocks = 592;  // 4 x 148 SMs
constexpr int kRedThreads = 256;

template <int QQ>
__global__ void reduceStage1(const double *__restrict__ state, long long S, int nFluid,
                             double *__restrict__ scratch) {
  __shared__ double sm[kRedThreads], su[kRedThreads], sn[kRedThreads];
  double mass = 0.0, umax = 0.0, nnan = 0.0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nFluid; e += gridDim.x * blockDim.x) {
    double f[QQ];
#pragma unroll
    for (int q = 0; q < QQ; ++q) f[q] = state[(long long)q * S + e];
    double rho, mx, my, mz;
    moments<QQ>(f, rho, mx, my, mz);
    mass += rho;
    const double u2 = (mx * mx + my * my + mz * mz) / (rho * rho);
    if (u2 > umax) umax = u2;
    if (rho != rho) nnan += 1.0;
  }
  sm[threadIdx.x] = mass; su[threadIdx.x] = umax; sn[threadIdx.x] = nnan;
  __syncthreads();
  for (int s = kRedThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sm[threadIdx.x] += sm[threadIdx.x + s];
      su[threadIdx.x] = fmax(su[threadIdx.x], su[threadIdx.x + s]);
      sn[threadIdx.x] += sn[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    scratch[blockIdx.x] = sm[0];
    scratch[kRedBlocks + blockIdx.x] = su[0];
    scratch[2 * kRedBlocks + blockIdx.x] = sn[0];
  }
}

__global__ void reduceStage2(const double *__restrict__ scratch, double *__restrict__ out) {
  __shared__ double sm[1024], su[1024], sn[1024];
  const int t = threadIdx.x;
  sm[t] = (t < kRedBlocks) ? scratch[t] : 0.0;
  su[t] = (t < kRedBlocks) ? scratch[kRedBlocks + t] : 0.0;
  sn[t] = (t < kRedBlocks) ? scratch[2 * kRedBlocks + t] : 0.0;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (t < s) {
      sm[t] += sm[t + s];
      su[t] = fmax(su[t], su[t + s]);
      sn[t] += sn[t + s];
    }
    __syncthreads();
  }
  if (t == 0) { out[0] = sm[0]; out[1] = su[0]; out[2] = sn[0]; }
}

int launchReduce(int QQ, const double *state, long long S, int nFluid, double *scratch,
                 double *out, cudaStream_t st) {
  if (QQ == 19) reduceStage1<19><<<kRedBlocks, kRedThreads, 0, st>>>(state, S, nFluid, scratch);
  else reduceStage1<27><<<kRedBlocks, kRedThreads, 0, st>>>(state, S, nFluid, scratch);
  MUSB_CUDA(cudaGetLastError());
  reduceStage2<<<1, 1024, 0, st>>>(scratch, out);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace musb200
