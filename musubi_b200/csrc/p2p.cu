// p2p.cu -- halo exchange through peer memory (NVLink 5 / NVSwitch, ranks of one node).
//
// Replaces comm_isend_irecv_real (tem/source/tem_comm_module.fpp:549-646: gather state(pos(i))
// -> MPI_Isend / MPI_Irecv -> MPI_Waitall -> scatter state(pos(i)) = val(i)) by ONE kernel per
// level step: every communicated link is loaded from the local state array and stored
// directly into the halo row of the receiving rank's state array (a peer-mapped pointer
// obtained through CUDA IPC), so there is no send buffer, no receive buffer and no unpack.
//
// Synchronisation (all counters are 64-bit exchange numbers that only grow):
//   * after its stores a CTA issues __threadfence_system() and bumps a local ticket; the CTA
//     that draws the last ticket publishes `count` into arrived[myRank] of every receiver
//     (system-scope store behind a system fence: the links are visible before the flag);
//   * that same CTA then waits until arrived[p] >= count for every rank p this rank receives
//     from, so when the kernel has finished the exchange is complete on this rank -- the
//     semantics of MPI_Waitall.
//   * no "ready to receive" handshake is needed: rank A writes exchange n into B's state(:,next)
//     only after A's own sweep n, which waited for B's exchange n-1, which B issued after the
//     sweep that last READ those halo rows (they belonged to B's state(:,now) then).
#include "kernels.cuh"

namespace musb200 {

__global__ void pushHaloKernel(P2PArgs a) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const int p = a.srcPos[i] - 1;
    const double v = a.state[(long long)(p % a.QQ) * a.S + p / a.QQ];
    const int k = a.peerOf[i];
    const int r = a.dstPos[i] - 1;
    a.remoteState[k][(long long)(r % a.QQ) * a.remoteS[k] + r / a.QQ] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned int ticket = atomicAdd(a.ticket, 1u);
  if (ticket != gridDim.x - 1) return;
  *a.ticket = 0u;                      // ready for the next launch (stream order)
  __threadfence_system();
  for (int k = 0; k < a.nSendPeers; ++k) {
    volatile unsigned long long *flag = a.remoteArrived[k] + a.myRank;
    *flag = a.count;
  }
  __threadfence_system();
  for (int k = 0; k < a.nRecvPeers; ++k) {
    volatile unsigned long long *flag = a.arrived + a.recvRank[k];
    while (*flag < a.count) { __nanosleep(100); }
  }
  __threadfence_system();
}

// The handshake alone, for steps whose links were stored by the sweep itself (sweep_push.cu):
// the sweep kernel has completed before this one starts (stream order), its peer stores are
// performed; the system fence orders them before the arrival flags for every observer.
__global__ void signalHaloKernel(P2PArgs a) {
  const int t = threadIdx.x;
  __threadfence_system();
  if (t < a.nSendPeers) {
    volatile unsigned long long *flag = a.remoteArrived[t] + a.myRank;
    *flag = a.count;
  }
  __threadfence_system();
  if (t < a.nRecvPeers) {
    volatile unsigned long long *flag = a.arrived + a.recvRank[t];
    while (*flag < a.count) { __nanosleep(100); }
  }
  __threadfence_system();
}

int launchSignalHalo(const P2PArgs &a, cudaStream_t st) {
  signalHaloKernel<<<1, 32, 0, st>>>(a);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchPushHalo(const P2PArgs &a, cudaStream_t st) {
  // enough CTAs to saturate NVLink stores, few enough to keep the ticket cheap
  int blocks = divUp(a.n > 0 ? a.n : 1, 256);
  if (blocks > 296) blocks = 296;      // 2 x 148 SMs
  pushHaloKernel<<<blocks, 256, 0, st>>>(a);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace musb200
