// p2p.cu -- halo exchange through peer memory (NVLink 5 / NVSwitch, ranks of one node).
//
// Replaces comm_isend_irecv_real (tem/source/tem_comm_module.fpp:549-646: gather state(pos(i))
// -> MPI_Isend / MPI_Irecv -> MPI_Waitall -> scatter state(pos(i)) = val(i)) and, in multi-level
// runs, the auxField exchange that follows it (auxField%sendBuffer, mus_auxField_module.f90:
// 377-396): every communicated link -- and every auxField entry of the communicated elements --
// is loaded from the local arrays and stored directly into the halo rows of the receiving rank's
// arrays (peer-mapped pointers obtained through CUDA IPC).  No send buffer, no receive buffer, no
// unpack.
//
// The exchange is split into its two halves so that the WAIT can move to where the data is
// needed:
//   pushHaloKernel   (MPI_Isend)  the stores, a system fence, then the CTA that draws the last
//                    ticket bumps this rank's exchange number `exch` (device memory, so the
//                    launch is CUDA-graph capturable) and publishes it into arrived[myRank] of
//                    every receiver.  It does not wait for anything.
//   the wait         (MPI_Waitall) arrived[p] >= exch for every rank p this rank receives from.
//                    Single level: inside the NEXT sweep, and only by the CTAs that pull from a
//                    halo row (a bitmap with one bit per CTA, built once from the neighbour list)
//                    -- every other CTA of that sweep runs while the links are still in flight,
//                    which hides the transfer and the skew between the ranks behind compute
//                    without splitting the sweep.  Elsewhere (multi-level: the interpolation
//                    reads halo rows right away; end of a musb200_step call): waitHaloKernel.
//   Progress: push(n) is stream-ordered after sweep(n) and waits for nothing; the waiting CTAs of
//   sweep(n+1) need the peers' push(n), which needs the peers' sweep(n), whose waiting CTAs need
//   this rank's push(n-1) -- complete before sweep(n) started.  No cycle, and no kernel has to
//   be co-resident with another one.
//   No "ready to receive" handshake is needed either: a peer writes exchange n+1 into the buffer
//   this rank read during sweep(n) only after its own sweep(n+1), whose waiting CTAs needed
//   this rank's push(n), issued after sweep(n) had finished (peer sets are symmetric; checked
//   at connect).
//   Failure handling: every wait gives up after `timeoutNs` (musb200_set_exchange_timeout),
//   records {code, peer, exchange number} in a host-mapped flag and carries on, so the stream
//   drains and the next API call returns MUSB200_ERR_NCCL instead of hanging (the reference
//   aborts all ranks, tem/source/tem_aux_module.f90:457-478).
#include "kernels.cuh"

namespace musb200 {

// one (entry -> remote store); QQ is a template parameter so that position -> (direction, element)
// is a multiply-shift, not a division
template <int QQ>
__device__ __forceinline__ void pushOne(const P2PArgs &a, int i) {
  if (i < a.n) {
    const int p = a.srcPos[i] - 1;
    const double v = a.state[(long long)(p % QQ) * a.S + p / QQ];
    const int k = a.peerOf[i];
    const int r = a.dstPos[i] - 1;
    a.remoteState[k][(long long)(r % QQ) * a.remoteS[k] + r / QQ] = v;
  } else {
    const int j = i - a.n;
    const int p = a.auxSrcPos[j] - 1;
    const double v = a.aux[(long long)(p & 3) * a.S + (p >> 2)];
    const int k = a.auxPeerOf[j];
    const int r = a.auxDstPos[j] - 1;
    a.remoteAux[k][(long long)(r & 3) * a.remoteS[k] + (r >> 2)] = v;
  }
}

template <int QQ>
__global__ void __launch_bounds__(256) pushHaloKernel(P2PArgs a) {
  if (a.handshake) {
    // "ready to receive": this rank has finished every kernel that read exchange n-1 (stream
    // order), so its peers may overwrite the single-buffered auxField halo rows; and it stores
    // only once its receivers have said the same.  Publish first, then wait: no cycle.
    if (threadIdx.x == 0) {
      const unsigned long long want = *reinterpret_cast<volatile unsigned long long *>(a.exch) + 1ull;
      if (blockIdx.x == 0) {
        for (int k = 0; k < a.nSendPeers; ++k) {
          volatile unsigned long long *flag = a.remoteArrived[k] + a.nranks + a.myRank;
          *flag = want;
        }
        __threadfence_system();
      }
      unsigned long long t0 = 0ull;
      for (int k = 0; k < a.nSendPeers; ++k) {
        const volatile unsigned long long *flag = a.ready + a.sendRank[k];
        unsigned int spins = 0u;
        while (*flag < want) {
          __nanosleep(64);
          if (a.timeoutNs != 0ull && (++spins & 1023u) == 0u) {
            const unsigned long long now = globalTimerNs();
            if (t0 == 0ull) t0 = now;
            if (now - t0 > a.timeoutNs) {
              if (atomicCAS(a.errFlag, 0, 2) == 0) {
                a.errFlag[1] = a.sendRank[k];
                a.errFlag[2] = (int)(want & 0xffffffffull);
                a.errFlag[3] = (int)(want >> 32);
              }
              break;
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // the transfer is latency bound (index load -> gather -> remote store): many threads, and four
  // independent entries in flight per thread
  const int stride = gridDim.x * blockDim.x;
  const int total = a.n + a.nAux;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < total; i += 4 * stride) {
    pushOne<QQ>(a, i);
    pushOne<QQ>(a, i + stride);
    pushOne<QQ>(a, i + 2 * stride);
    pushOne<QQ>(a, i + 3 * stride);
  }
  for (; i < total; i += stride) pushOne<QQ>(a, i);
  // one system fence per CTA instead of one per thread: the barrier orders the CTA's stores before
  // thread 0, whose fence (cumulative) orders them before its ticket and the flags at system scope
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence_system();
  const unsigned int ticket = atomicAdd(a.ticket, 1u);
  if (ticket != gridDim.x - 1) return;
  *a.ticket = 0u;                      // ready for the next launch (stream order)
  unsigned long long count;
  if (a.publish != nullptr) {
    count = *reinterpret_cast<const volatile unsigned long long *>(a.publish);
  } else {
    count = *a.exch + 1ull;
    *a.exch = count;
  }
  __threadfence_system();
  for (int k = 0; k < a.nSendPeers; ++k) {
    volatile unsigned long long *flag = a.remoteArrived[k] + a.myRank;
    *flag = count;
  }
  __threadfence_system();
}

__global__ void bumpExchKernel(unsigned long long *exch, unsigned long long *slot) {
  const unsigned long long n = *exch + 1ull;
  *exch = n;
  *slot = n;
}

int launchBumpExch(unsigned long long *exch, unsigned long long *slot, cudaStream_t st) {
  bumpExchKernel<<<1, 1, 0, st>>>(exch, slot);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

// the publishing half alone, for steps whose links were stored by the sweep itself
// (sweep_push.cu): that kernel has completed (stream order), its peer stores are performed; the
// system fence orders them before the arrival flags for every observer
__global__ void signalHaloKernel(P2PArgs a) {
  const int t = threadIdx.x;
  __shared__ unsigned long long count;
  if (t == 0) { count = *a.exch + 1ull; *a.exch = count; }
  __syncthreads();
  __threadfence_system();
  if (t < a.nSendPeers) {
    volatile unsigned long long *flag = a.remoteArrived[t] + a.myRank;
    *flag = count;
  }
  __threadfence_system();
}

__global__ void waitHaloKernel(HaloWait w) {
  if (threadIdx.x == 0) waitHaloArrival(w);
}

// one bit per CTA of the sweep (block threads elements each): set when an element of the CTA
// pulls from a row at or behind haloStart
template <int QQ>
__global__ void haloCtaMaskKernel(const uint32_t *__restrict__ nbr, long long S, int nSolve, int haloStart,
                                  int block, uint32_t *__restrict__ mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nSolve) return;
  bool dep = false;
#pragma unroll
  for (int q = 0; q < QQ - 1; ++q) {
    const uint32_t n = nbr[(long long)q * S + e];
    dep = dep || (!(n & kBounceBit) && (int)(n & kElemMask) >= haloStart);
  }
  if (dep) {
    const int cta = e / block;
    atomicOr(mask + (cta >> 5), 1u << (cta & 31));
  }
}

int launchHaloCtaMask(int QQ, const uint32_t *nbr, long long S, int nSolve, int haloStart, int block,
                      uint32_t *mask, cudaStream_t st) {
  if (nSolve <= 0) return 0;
  if (QQ == 19) haloCtaMaskKernel<19><<<divUp(nSolve, 256), 256, 0, st>>>(nbr, S, nSolve, haloStart, block, mask);
  else haloCtaMaskKernel<27><<<divUp(nSolve, 256), 256, 0, st>>>(nbr, S, nSolve, haloStart, block, mask);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchSignalHalo(const P2PArgs &a, cudaStream_t st) {
  signalHaloKernel<<<1, 32, 0, st>>>(a);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchWaitHalo(const HaloWait &w, cudaStream_t st) {
  waitHaloKernel<<<1, 32, 0, st>>>(w);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchPushHalo(const P2PArgs &a, cudaStream_t st) {
  // enough CTAs to saturate NVLink stores, few enough to keep the ticket cheap
  const int total = a.n + a.nAux;
  int blocks = divUp(total > 0 ? total : 1, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;      // a full wave of 256-thread CTAs
  if (a.QQ == 19) pushHaloKernel<19><<<blocks, 256, 0, st>>>(a);
  else pushHaloKernel<27><<<blocks, 256, 0, st>>>(a);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace musb200
