// kernels.cuh -- launch interface of the device kernels (implemented in *.cu)
#pragma once
#include "collide.cuh"
#include "mrt_tables.cuh"

namespace musb200 {

constexpr int kMaxPeers = 16;

// halo push fused into the sweep (sweep_push.cu): per send element (bit set in `mask`; its rank
// among the send elements = prefix[word] + popc of the lower bits) a CSR row of links
struct PushArgs {
  const uint32_t *mask;      // 1 bit per element, nullptr = no fused push
  const uint32_t *prefix;    // send elements before each 32-element word
  const int32_t *start;      // [nSendElems + 1]
  const uint8_t *entQ;       // local direction (0-based) of the link
  const uint8_t *entPeer;    // index into the peer tables
  const int32_t *entDst;     // receiver's state position (1-based, its recv buffer's pos list)
  double *remoteState[kMaxPeers];
  long long remoteS[kMaxPeers];
};

// the receiving half of the peer-memory halo exchange (p2p.cu): wait until every rank this rank
// receives from has published an exchange number >= this rank's own
struct HaloWait {
  const uint32_t *ctaMask;              // sweep: 1 bit per CTA that pulls from a halo row; nullptr = no wait
  int haloStart;                        // first halo element of the level's rows
  const unsigned long long *arrived;    // my arrived[nranks], written by the senders
  const unsigned long long *exch;       // my exchange number (device memory)
  int nRecvPeers;
  int recvRank[kMaxPeers];
  unsigned long long timeoutNs;         // 0 = wait for ever
  int *errFlag;                         // device memory: {code, peer rank, exchange lo, exchange hi}
};

__device__ __forceinline__ unsigned long long globalTimerNs() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// called by ONE thread; the caller separates it from the readers of the halo rows by a barrier
__device__ __forceinline__ void waitHaloArrival(const HaloWait &w) {
  const unsigned long long want = *reinterpret_cast<const volatile unsigned long long *>(w.exch);
  if (*reinterpret_cast<volatile int *>(w.errFlag) != 0) return;   // an exchange has failed already: drain
  unsigned long long t0 = 0ull;
  for (int k = 0; k < w.nRecvPeers; ++k) {
    const volatile unsigned long long *flag = w.arrived + w.recvRank[k];
    unsigned int spins = 0u;
    while (*flag < want) {
      __nanosleep(64);
      if (w.timeoutNs != 0ull && (++spins & 1023u) == 0u) {
        const unsigned long long now = globalTimerNs();
        if (t0 == 0ull) t0 = now;
        if (now - t0 > w.timeoutNs) {
          if (atomicCAS(w.errFlag, 0, 1) == 0) {           // first reporter keeps its details
            w.errFlag[1] = w.recvRank[k];
            w.errFlag[2] = (int)(want & 0xffffffffull);
            w.errFlag[3] = (int)(want >> 32);
            __threadfence_system();
          }
          return;
        }
      }
    }
  }
  __threadfence_system();
}

// Arguments of one fused "auxField + stream + collide" sweep over a level.
// State and aux are SoA with row stride S (elements): f[q][e] = ptr[q*S + e].
struct SweepArgs {
  const double *in;        // state(:, now)
  double *out;             // state(:, next)
  const uint32_t *nbr;     // [QQ-1][S] encoded pull sources (rest direction is implicit)
  double *aux;             // [4][S] rho, ux, uy, uz (written when write_aux)
  const double *omega;     // per-element omega or nullptr (uniform)
  // re-ordering of the sweep by whole CTAs for the overlapped halo exchange (several ranks, p2p.cu):
  // ctaMode 0: natural order, the CTAs of wait.ctaMask wait for the halo links before they gather;
  // ctaMode 1: the grid is nMain + nCtas CTAs -- the first nMain in natural order, of which the
  // CTAs of wait.ctaMask return at once; they are appended as ctaList[0 .. nCtas) and wait there,
  // i.e. at the END of the launch, when the push that overlapped its beginning has long arrived
  const int32_t *ctaList;
  int ctaMode;
  int nMain, nCtas;
  long long S;
  int first;               // first element (0-based) of the contiguous range
  int count;               // number of elements (range) or list entries
  int write_aux;
  RelaxParams rp;
  // body-force source (0 = none): per-element SoA [3][S] in lattice units, or uniform
  int force_order;
  const double *force;
  double force_uniform[3];
  PushArgs push;
  HaloWait wait;         // single level, several ranks: CTAs that pull from halo rows wait for the exchange
};

int launchSweep(int QQ, int relax, int kind, const SweepArgs &a, cudaStream_t st);
// auxField of the elements [first, first + count) from state(:, now) (fluid kinds only)
int launchAuxOnly(int QQ, int kind, const SweepArgs &a, cudaStream_t st);

// passive scalar (passive_scalar.cu)
struct PsArgs {
  const double *in;
  double *out;
  const uint32_t *nbr;
  double *aux;             // [1][S] zeroth moment (written when write_aux)
  long long S;
  int count;               // nElems_solve
  int write_aux;
  const double *vel;       // transport velocity rows [3][velS] or nullptr (uniform)
  long long velS;
  double vel_uniform[3];
  double d_omega;          // 2 / (1 + 6 diff_coeff)
  double aux_omega;        // trt: 1 / (lambda / (1/d_omega - 1/2) + 1/2)
};
// variant: 1 bgk/first, 2 bgk/second, 3 trt (vStdNoOpt)
int launchPassiveScalar(int QQ, int variant, const PsArgs &a, cudaStream_t st);
// SoA rows [nComp][S] <- AOS list entries: dst[k*S + pos[i]-1] = src[i*nComp + k] (pos == nullptr: i)
int launchScatterRows(const double *aos, const int32_t *pos, int n, int nComp, double *soa, long long S,
                      cudaStream_t st);

// layout conversion
int launchAosToSoa(const double *aos, double *soa, int nComp, int nElems, long long S,
                   cudaStream_t st);
int launchSoaToAos(const double *soa, double *aos, int nComp, int nElems, long long S,
                   cudaStream_t st);
// neigh (Fortran positions, NGPOS layout) -> encoded device list; *bad counts entries
// that are neither plain pulls nor bounce-backs
int launchEncodeNeigh(int QQ, const int32_t *neigh, uint32_t *nbr, int nSize, int nElems,
                      long long S, int *bad, cudaStream_t st);
int launchDecodeNeigh(int QQ, const uint32_t *nbr, int32_t *neigh, int nSize, int nElems,
                      long long S, cudaStream_t st);

// treelm's predefined cube generated on the device + equilibrium initial state (cube.cu)
int launchCubeNeigh(int QQ, uint32_t *nbr, int level, int walls, long long S, int nElems, cudaStream_t st);
int launchInitEquilibrium(int QQ, int incomp, const double *aux, double *s0, double *s1, long long S, int nElems,
                          cudaStream_t st);

// boundary kernels
int launchFillBcBuffer(int QQ, const double *state, long long S, const int32_t *bcElems,
                       const int32_t *needed, int nNeeded, double *bcBuffer, cudaStream_t st);
int launchVelocityBounceBack(int QQ, int incomp, double *state, long long S,
                             const double *bcBuffer, int nLinks, const int32_t *links,
                             const int32_t *outPos, const int32_t *posInBuffer,
                             const int32_t *iDir, const double *velLat, cudaStream_t st);

int launchVelocityBounceBackFused(int QQ, int incomp, double *state, long long S, int nGroups,
                                  const int32_t *groupStart, const int32_t *groupElem,
                                  const int32_t *links, const int32_t *outPos, const int32_t *iDir,
                                  const double *velLat, cudaStream_t st);

// boundaries that read neighbours along the inward normal
int launchFillNeighBuffer(int QQ, const double *state, long long S, const uint32_t *nbr, int nNeighs,
                          int nElems, const int32_t *neighPos, int post, double *nb, cudaStream_t st);
int launchPressureExpol(int QQ, int incomp, double *state, long long S, const uint32_t *nbr,
                        const double *bcBuffer, const double *aux, int nElems, const int32_t *elemPos,
                        const int32_t *posInBcElemBuf, const int32_t *normalInd, const double *rhoDef,
                        int nLinks, const int32_t *links, const int32_t *iElemOfLink,
                        const int32_t *iDir, const double *nbPre, cudaStream_t st);
int launchPressureAntiBounceBack(int QQ, int incomp, double *state, long long S,
                                 const double *bcBuffer, int nLinks, const int32_t *links,
                                 const int32_t *iElemOfLink, const int32_t *iDir,
                                 const int32_t *elemPos, const int32_t *posInBcElemBuf,
                                 const double *rhoDef, const double *omegaElem, double omegaUniform,
                                 const double *nbPost, cudaStream_t st);

// halo exchange pack / unpack (positions are Fortran AOS state positions)
int launchPack(int QQ, const double *state, long long S, const int32_t *pos, int n, double *buf,
               cudaStream_t st);
int launchUnpack(int QQ, double *state, long long S, const int32_t *pos, int n, const double *buf,
                 cudaStream_t st);

// restart bridge (mus_pdf_serialize order): chunk buffer <-> state of one level
int launchSerialize(int QQ, const double *state, long long S, const int32_t *slot, const int32_t *elemPos,
                    int n, double *buffer, cudaStream_t st);
int launchUnserialize(int QQ, double *state, long long S, const int32_t *slot, const int32_t *elemPos,
                      int n, const double *buffer, cudaStream_t st);

// peer-memory halo exchange (p2p.cu)
struct P2PArgs {
  const double *state;        // my state(:, next)
  const double *aux;          // my auxField (multi-level: travels with the halo elements), or nullptr
  long long S;
  int QQ, n, nAux;            // n = all state send entries, nAux = all auxField entries, peers concatenated
  const int32_t *srcPos;      // my state positions (the send buffer's pos list)
  const int32_t *dstPos;      // the receiver's state positions (its recv buffer's pos list)
  const uint8_t *peerOf;      // entry -> index into the peer tables below
  const int32_t *auxSrcPos;   // auxField positions (elem-1)*4 + k, mine / the receiver's
  const int32_t *auxDstPos;
  const uint8_t *auxPeerOf;
  int nSendPeers, myRank;
  double *remoteState[kMaxPeers];              // receiver's state(:, next), peer-mapped
  double *remoteAux[kMaxPeers];                // receiver's auxField, peer-mapped
  long long remoteS[kMaxPeers];
  unsigned long long *remoteArrived[kMaxPeers];  // receiver's arrived[nranks], peer-mapped
  unsigned long long *exch;     // my exchange number, bumped by the kernel ...
  const unsigned long long *publish;   // ... unless set: the number to publish, bumped in the stepping stream
  unsigned int *ticket;
  // auxField rows are single-buffered: before storing, wait until every receiver has announced
  // that it no longer reads the previous exchange (ready[] sits behind arrived[], same mapping)
  int handshake;
  int nranks;
  const unsigned long long *ready;   // my ready[nranks], written by the peers
  int sendRank[kMaxPeers];
  unsigned long long timeoutNs;
  int *errFlag;
};
int launchPushHalo(const P2PArgs &a, cudaStream_t st);
// publish only (the links were stored by the sweep with the fused push)
int launchSignalHalo(const P2PArgs &a, cudaStream_t st);
// MPI_Waitall of the exchange as a one-thread kernel
int launchWaitHalo(const HaloWait &w, cudaStream_t st);
// overlapped exchange: the exchange number advances in the STEPPING stream (the waits read it
// there); the push on the communication stream publishes the value parked in *slot
int launchBumpExch(unsigned long long *exch, unsigned long long *slot, cudaStream_t st);
// bitmap of the sweep CTAs (block elements each) that pull from rows >= haloStart
int launchHaloCtaMask(int QQ, const uint32_t *nbr, long long S, int nSolve, int haloStart, int block,
                      uint32_t *mask, cudaStream_t st);
// threads per CTA of the sweep for this stencil (sweep.cu)
int sweepBlockSize(int QQ);

// reductions: out[0] = total mass, out[1] = max |u|^2, out[2] = nan count
int launchReduce(int QQ, const double *state, long long S, int nFluid, double *scratch,
                 double *out, cudaStream_t st);

}  // namespace musb200
