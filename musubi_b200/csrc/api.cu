// api.cu -- the C ABI of libmusb200.so (include/musb200.h): per-level device
// data, host <-> device conversion, and the time-step schedule of
// mus/source/mus_control_module.f90 (do_fast_singleLevel :507-701,
// do_recursive_multiLevel :242-497, do_intpFinerAndExchange :861-947,
// do_intpCoarserAndExchange :955-1051) issued on CUDA streams.
#include "../../include/musb200.h"
#include "intp.cuh"
#include "kernels.cuh"
#include "nccl_dyn.h"

#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

namespace musb200 {

// ---------------------------------------------------------------------------
std::string &lastError() {
  static std::string s;
  return s;
}
int setError(int code, const std::string &msg) {
  lastError() = msg;
  return code;
}

NcclApi *ncclApi() {
  static NcclApi api;
  static bool tried = false;
  if (api.handle) return &api;
  if (tried) return nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    setError(MUSB200_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
    return nullptr;
  }
#define SYM(field, name)                                                        \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));  \
  if (!api.field) {                                                            \
    setError(MUSB200_ERR_NCCL, std::string("libnccl lacks ") + name);          \
    api.handle = nullptr;                                                      \
    return nullptr;                                                            \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(CommGetAsyncError, "ncclCommGetAsyncError")
  SYM(CommAbort, "ncclCommAbort")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return &api;
}

#define MUSB_NCCL(call)                                                                   \
  do {                                                                                    \
    ncclResult_t r__ = (call);                                                            \
    if (r__ != ncclSuccess)                                                               \
      return setError(MUSB200_ERR_NCCL, std::string(#call) + ": " + g.nccl->GetErrorString(r__)); \
  } while (0)

#define MUSB_TRY(call)            \
  do {                            \
    int rc__ = (call);            \
    if (rc__ != 0) return rc__;   \
  } while (0)

// ---------------------------------------------------------------------------
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return 0;
    MUSB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    return 0;
  }
  int upload(const T *h, size_t count, cudaStream_t st) {
    MUSB_TRY(alloc(count));
    if (count) MUSB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};

struct BcData {
  int id = 0, kind = 0, nLinks = 0;
  DevBuf<int32_t> links, outPos, posInBuffer, iDir;
  DevBuf<double> vals;
  // musb200_bc_set_values copies into valsNext on the copy stream (overlapping the running step);
  // the next set_boundary waits for the copy and swaps the buffers in
  DevBuf<double> valsNext;
  bool pending = false;
  cudaEvent_t copied = nullptr;
  ~BcData() { if (copied) cudaEventDestroy(copied); }
  // velocity_bounceback: links grouped per boundary element (fused kernel, bc.cu)
  bool fusable = false;
  int nGroups = 0;
  DevBuf<int32_t> groupStart, groupElem;
  std::vector<int32_t> slots;   // bc_elemBuffer slots this boundary touches
  // boundaries that read neighbours along the inward normal (musb200_bc_register_elems)
  int nElems = 0, nNeighs = 0;
  DevBuf<int32_t> elemPos, posInBcElemBuf, normalInd, neighPos, iElemOfLink;
  DevBuf<double> neighBuf;  // neighBufferPre_nNext (expol) / neighBufferPost (anti-bounce-back)
};
static bool isPressureBc(int kind) {
  return kind == MUSB200_BC_PRESSURE_EXPOL || kind == MUSB200_BC_PRESSURE_ANTIBOUNCEBACK;
}

struct CommBuf {
  std::vector<int> proc, nVals, offset;
  int total = 0;
  DevBuf<int32_t> pos;
  DevBuf<double> buf;
  // auxField%sendBuffer / recvBuffer (mus_auxField_module.f90:377-396) of the same elements: the
  // four auxField entries of every element that appears in `pos`, in order of first appearance
  std::vector<int> auxNVals, auxOffset;
  int auxTotal = 0;
  DevBuf<int32_t> auxPos;
  std::vector<int32_t> auxPosHost;   // the same list on the host (peer-memory set-up)
};

// peer-memory halo exchange of one level (p2p.cu)
struct PeerLink {
  bool on = false;
  bool pendingWait = false;              // a push has been issued whose arrival nobody has waited for yet
  bool sweepWait = false;                // the next sweep may do that wait (symmetric peers, CTA bitmap built)
  std::vector<int> sendRank, recvRank;
  std::vector<void *> opened;            // IPC-opened pointers, closed at destroy
  double *remoteState[kMaxPeers][2] = {};
  double *remoteAux[kMaxPeers] = {};
  long long remoteS[kMaxPeers] = {};
  unsigned long long *remoteArrived[kMaxPeers] = {};
  DevBuf<int32_t> srcPos, dstPos;        // send entries re-ordered for coalesced remote stores
  DevBuf<uint8_t> peerOf;
  DevBuf<int32_t> auxSrcPos, auxDstPos;  // the auxField entries of the same elements
  DevBuf<uint8_t> auxPeerOf;
  int nAux = 0;
  DevBuf<unsigned long long> arrived;    // [nranks], written by the senders
  DevBuf<unsigned long long> exch;       // [3]: this rank's exchange number, bumped by the push kernel (or, overlapped
                                         // exchange, in the stepping stream) | the number parked for the push in flight, per parity
  DevBuf<uint32_t> ctaMask;              // sweep CTAs that pull from a halo row
  DevBuf<int32_t> haloCtas;              // the same CTAs as a list (second launch of the overlapped exchange)
  int nHaloCtas = 0, nCtas = 0;
  // overlapped exchange: sweep done; push done, one event per buffer parity (the push of step n
  // reads the buffer that the sweep of step n + 2 overwrites)
  cudaEvent_t evSwept = nullptr, evPushed[2] = {nullptr, nullptr};
  bool pushedOnce[2] = {false, false};
  DevBuf<unsigned int> ticket;
  // fused push (sweep_push.cu): the send entries grouped by the element that owns them
  DevBuf<uint32_t> pushMask, pushPrefix;
  DevBuf<int32_t> pushStart, pushDst;
  DevBuf<uint8_t> pushQ, pushPeer;
  ~PeerLink() {
    for (void *p : opened) cudaIpcCloseMemHandle(p);
    if (evSwept) cudaEventDestroy(evSwept);
    for (cudaEvent_t e : evPushed)
      if (e) cudaEventDestroy(e);
  }
};

struct Level {
  int level = 0, QQ = 0, nSize = 0, nFluid = 0, nGFC = 0, nGFF = 0, nHalo = 0;
  int nElems = 0, nSolve = 0;
  long long S = 0;
  int nNow = 0, nNext = 1;  // 0-based buffer indices
  DevBuf<double> state[2], aux, omega, visc, bcBuffer;
  bool elemVisc = false, viscSet = false;
  double viscUniform = 0.0;
  DevBuf<uint32_t> nbr;
  DevBuf<int32_t> bcElems;
  std::vector<int32_t> bcElemsHost;
  bool auxValid = false;            // auxField holds the moments of the last level step
  int bcFused = -1;                 // -1 unknown, 0 two-phase (bcBuffer), 1 fused kernels
  std::vector<char> bcSlotNeeded;   // bcBuffer slots read by a non-wall boundary
  DevBuf<int32_t> bcNeeded;         // those slots (1-based), compact
  int relax = 0, kind = 0;
  bool relaxSet = false, elemOmega = false;
  bool auxForBc = false;            // a boundary reads auxField: materialise it every step
  RelaxParams rp{1.0, 0.25, 1.0};
  std::vector<std::unique_ptr<BcData>> bcs;
  CommBuf send[3], recv[3];
  // elements that own a link of the halo send buffer (prp_sendHalo, set_sendHaloBits
  // mus_construction_module.fpp:2765): swept first so that their exchange overlaps the rest
  PeerLink p2p;
  // source = { force }: order 0 (none) / 1 / 2, uniform or per-element SoA [3][S]
  int forceOrder = 0;
  bool forceElem = false;
  double forceUniform[3] = {0, 0, 0};
  DevBuf<double> force;
  // passive scalar: kernel variant (1 bgk/first, 2 bgk/second, 3 trt), species parameters,
  // transport velocity (uniform / own array / auxField rows of a flow scheme)
  int nAux = 4;
  int psVariant = 0;
  double psDOmega = 0.0, psAuxOmega = 0.0;
  int velMode = 0;             // 0 uniform, 1 own array, 2 coupled to (slot, level)
  double velUniform[3] = {0, 0, 0};
  DevBuf<double> vel;
  int velSlot = 0, velLevel = 0;
  IntpSet fromFiner;                 // fill my ghostFromFiner from level+1
  std::vector<IntpSet> fromCoarser;  // fill my ghostFromCoarser from level-1, per order
};

struct Context {
  bool ready = false;
  int rank = 0, nranks = 1, device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t commStream = nullptr;   // halo exchange, high priority, overlaps the interior sweep
  cudaStream_t copyStream = nullptr;   // boundary values host -> device, overlaps the running step
  cudaEvent_t evBcDone = nullptr;      // the boundary kernels of the last step have read their values
  cudaEvent_t evBoundary = nullptr, evComm = nullptr;
  // CUDA graph of two coarse cycles (after two cycles every level's now/next parity is back where
  // it started, so the captured kernel arguments are valid for every replay)
  bool capturing = false;
  int useGraphs = 1;
  unsigned long long epoch = 0, graphEpoch = 0;
  int graphMin = -1, graphMax = -1, graphSlot = -1, graphParity = -1;
  cudaGraphExec_t graphExec = nullptr;
  long long graphLaunches = 0;
  // peer-memory exchange with the links stored by the sweep itself (musb200_set_fused_push):
  // measured 1.1 % slower than the separate, coalescing push kernel (profiles/r01_multi_gpu.md)
  int fusedPush = 0;
  int noFusedBc = 0; // musb200_set_fused_bc(0): always take the two-phase bcBuffer path
  int overlap = 0;   // measured slower than exchange-after-compute at 256^3 per GPU (profiles/)
  int sweepWait = 0; // peer-memory exchange: 1 = the wait for the halo links moves into the next sweep
  bool commBusy = false;   // a push is in flight on the communication stream
  unsigned long long timeoutNs = 30000000000ull;   // every wait of the exchange gives up after this
  DevBuf<int> errFlag;                             // {code, peer, exchange lo, exchange hi}, set by a wait
  NcclApi *nccl = nullptr;
  ncclComm_t comm = nullptr;
  // one element list per (scheme slot, level); slot 0 is the default scheme, further slots hold
  // schemes living on the same mesh (a passive scalar transported by the flow of slot 0)
  static constexpr int kSlots = 4;
  int slot = 0;
  std::map<int, std::unique_ptr<Level>> levelsOf[kSlots];
  std::map<int, std::unique_ptr<Level>> &levels() { return levelsOf[slot]; }
  DevBuf<double> stage;  // AOS staging for up/download
  DevBuf<double> red;    // reduction scratch + result
  DevBuf<int> flag;
  int auxEveryStep = 0;
  long long launches = 0;
  cudaEvent_t mark[2] = {nullptr, nullptr};
  // section timers (device time, only sampled when profiling is on)
  struct Span { int cat; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> pool;
  int profiling = 0;
  double acc[4] = {0, 0, 0, 0};
};
static Context g;

static void dropGraph() {
  if (g.graphExec) cudaGraphExecDestroy(g.graphExec);
  g.graphExec = nullptr;
}

static int needReady() {
  if (!g.ready) return setError(MUSB200_ERR_STATE, "musb200_init has not been called");
  return 0;
}
static Level *findLevel(int level) {
  auto it = g.levels().find(level);
  return it == g.levels().end() ? nullptr : it->second.get();
}
// read-only queries (downloads, probes, reductions) leave a captured step graph valid
#define GET_LEVEL_RO(L, level)                                                       \
  MUSB_TRY(needReady());                                                             \
  Level *L = findLevel(level);                                                       \
  if (!L) return setError(MUSB200_ERR_ARG, "unknown level " + std::to_string(level))
// calls that may change what a captured step graph would replay bump the epoch
#define GET_LEVEL(L, level)                                                          \
  GET_LEVEL_RO(L, level);                                                            \
  ++g.epoch

static int stageBuf(size_t n) {
  if (g.stage.n < n) MUSB_TRY(g.stage.alloc(n));
  return 0;
}

struct Timed {
  int cat;
  bool on;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  explicit Timed(int c, cudaStream_t stream = nullptr)
      : cat(c), on(g.profiling != 0), st(stream ? stream : g.stream) {
    if (!on) return;
    auto get = [] {
      cudaEvent_t e;
      if (!g.pool.empty()) { e = g.pool.back(); g.pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    a = get(); b = get();
    cudaEventRecord(a, st);
  }
  ~Timed() {
    if (!on) return;
    cudaEventRecord(b, st);
    g.spans.push_back({cat, a, b});
  }
};
enum { T_COMPUTE = 0, T_BC = 1, T_COMM = 2, T_INTP = 3 };

// ---------------------------------------------------------------------------
// the schedule
// boundary values handed over since the last step: wait for their copies, make them current
static int applyPendingBc(Level &L) {
  for (auto &b : L.bcs) {
    if (!b->pending) continue;
    MUSB_CUDA(cudaStreamWaitEvent(g.stream, b->copied, 0));
    std::swap(b->vals.p, b->valsNext.p);
    std::swap(b->vals.n, b->valsNext.n);
    b->pending = false;
  }
  return 0;
}

static int ensureArrived(Level &L, cudaStream_t st = nullptr);

static int setBoundary(Level &L) {
  if (L.bcElems.n == 0) return 0;
  MUSB_TRY(applyPendingBc(L));
  for (auto &b : L.bcs)   // boundaries that read neighbour elements may read halo rows
    if (isPressureBc(b->kind) && b->nLinks > 0) { MUSB_TRY(ensureArrived(L, nullptr)); break; }
  Timed t(T_BC);
  double *st = L.state[L.nNext].p;
  bool any = false;
  for (auto &b : L.bcs) any = any || (b->kind != MUSB200_BC_WALL && b->nLinks > 0);
  if (!any) return 0;
  if (L.bcFused < 0) {
    // fused path: only velocity_bounceback boundaries, each fusable, on disjoint elements
    bool ok = true;
    std::vector<char> used(L.bcSlotNeeded.size(), 0);
    for (auto &b : L.bcs) {
      if (b->kind == MUSB200_BC_WALL || b->nLinks == 0) continue;
      if (b->kind != MUSB200_BC_VELOCITY_BOUNCEBACK || !b->fusable) { ok = false; break; }
      for (int32_t sl : b->slots) {
        if (used[sl - 1]) ok = false;
        used[sl - 1] = 1;
      }
    }
    L.bcFused = ok ? 1 : 0;
  }
  if (L.bcFused == 1 && !g.noFusedBc) {
    for (auto &b : L.bcs) {
      if (b->kind == MUSB200_BC_WALL || b->nLinks == 0) continue;
      if (b->vals.n < (size_t)3 * b->nLinks)
        return setError(MUSB200_ERR_STATE, "velocity_bounceback: musb200_bc_set_values missing");
      MUSB_TRY(launchVelocityBounceBackFused(L.QQ, L.kind == MUSB200_KIND_FLUID_INCOMPRESSIBLE, st, L.S,
                                             b->nGroups, b->groupStart.p, b->groupElem.p, b->links.p,
                                             b->outPos.p, b->iDir.p, b->vals.p, g.stream));
      ++g.launches;
    }
    return 0;
  }
  // fill_bcBuffer restricted to the slots a non-wall boundary reads (walls are do_nothing)
  MUSB_TRY(launchFillBcBuffer(L.QQ, st, L.S, L.bcElems.p, L.bcNeeded.p, (int)L.bcNeeded.n, L.bcBuffer.p,
                              g.stream));
  ++g.launches;
  // fill_neighBuffer for every boundary that needs it, before any boundary writes its links
  for (auto &b : L.bcs) {
    if (!isPressureBc(b->kind) || b->nLinks == 0) continue;
    if (b->nElems == 0) return setError(MUSB200_ERR_STATE, "pressure boundary: musb200_bc_register_elems missing");
    const int post = b->kind == MUSB200_BC_PRESSURE_ANTIBOUNCEBACK ? 1 : 0;
    MUSB_TRY(launchFillNeighBuffer(L.QQ, st, L.S, L.nbr.p, b->nNeighs, b->nElems, b->neighPos.p, post,
                                   b->neighBuf.p, g.stream));
    ++g.launches;
  }
  const int incomp = L.kind == MUSB200_KIND_FLUID_INCOMPRESSIBLE;
  for (auto &b : L.bcs) {
    if (b->kind == MUSB200_BC_WALL || b->nLinks == 0) continue;
    if (isPressureBc(b->kind)) {
      if (b->vals.n < (size_t)b->nElems)
        return setError(MUSB200_ERR_STATE, "pressure boundary: musb200_bc_set_values missing (one density per element)");
      if (b->kind == MUSB200_BC_PRESSURE_EXPOL) {
        MUSB_TRY(launchPressureExpol(L.QQ, incomp, st, L.S, L.nbr.p, L.bcBuffer.p, L.aux.p, b->nElems,
                                     b->elemPos.p, b->posInBcElemBuf.p, b->normalInd.p, b->vals.p,
                                     b->nLinks, b->links.p, b->iElemOfLink.p, b->iDir.p, b->neighBuf.p,
                                     g.stream));
        g.launches += 2;
      } else {
        MUSB_TRY(launchPressureAntiBounceBack(L.QQ, incomp, st, L.S, L.bcBuffer.p, b->nLinks, b->links.p,
                                              b->iElemOfLink.p, b->iDir.p, b->elemPos.p,
                                              b->posInBcElemBuf.p, b->vals.p,
                                              L.elemOmega ? L.omega.p : nullptr, L.rp.omega_uniform,
                                              b->neighBuf.p, g.stream));
        ++g.launches;
      }
      continue;
    }
    if (b->kind == MUSB200_BC_VELOCITY_BOUNCEBACK) {
      if (b->vals.n < (size_t)3 * b->nLinks)
        return setError(MUSB200_ERR_STATE, "velocity_bounceback: musb200_bc_set_values missing");
      MUSB_TRY(launchVelocityBounceBack(L.QQ, L.kind == MUSB200_KIND_FLUID_INCOMPRESSIBLE, st, L.S,
                                        L.bcBuffer.p, b->nLinks, b->links.p, b->outPos.p,
                                        b->posInBuffer.p, b->iDir.p, b->vals.p, g.stream));
      ++g.launches;
    } else {
      return setError(MUSB200_ERR_UNSUPPORTED, "boundary kind not built yet");
    }
  }
  return 0;
}

// the receiving half of the peer-memory exchange (p2p.cu)
static HaloWait haloWait(Level &L, bool forSweep) {
  PeerLink &P = L.p2p;
  HaloWait w{};
  w.ctaMask = forSweep ? P.ctaMask.p : nullptr;
  w.haloStart = L.nFluid + L.nGFC + L.nGFF;
  w.arrived = P.arrived.p;
  w.exch = P.exch.p;
  w.nRecvPeers = (int)P.recvRank.size();
  for (int k = 0; k < w.nRecvPeers; ++k) w.recvRank[k] = P.recvRank[k];
  w.timeoutNs = g.timeoutNs;
  w.errFlag = g.errFlag.p;
  return w;
}
// MPI_Waitall of the last push of this level, unless a sweep has done (or will do) it
static int ensureArrived(Level &L, cudaStream_t st) {
  if (!L.p2p.pendingWait) return 0;
  Timed t(T_COMM, st ? st : g.stream);
  MUSB_TRY(launchWaitHalo(haloWait(L, false), st ? st : g.stream));
  ++g.launches;
  L.p2p.pendingWait = false;
  return 0;
}
// the sending half: one kernel stores every link (and, withAux, every auxField entry of the
// communicated elements) into the receivers' halo rows and publishes the exchange number
static int pushHalo(Level &L, bool withAux, cudaStream_t st, bool pushed, const unsigned long long *publish = nullptr) {
  PeerLink &P = L.p2p;
  CommBuf &s = L.send[MUSB200_BUF_HALO];
  P2PArgs a{};
  a.state = L.state[L.nNext].p; a.aux = L.aux.p; a.S = L.S; a.QQ = L.QQ;
  a.n = s.total; a.nAux = withAux ? P.nAux : 0;
  a.srcPos = P.srcPos.p; a.dstPos = P.dstPos.p; a.peerOf = P.peerOf.p;
  a.auxSrcPos = P.auxSrcPos.p; a.auxDstPos = P.auxDstPos.p; a.auxPeerOf = P.auxPeerOf.p;
  a.nSendPeers = (int)P.sendRank.size(); a.myRank = g.rank;
  for (int k = 0; k < a.nSendPeers; ++k) {
    a.remoteState[k] = P.remoteState[k][L.nNext];   // ranks swap now/next in lockstep
    a.remoteAux[k] = P.remoteAux[k];
    a.remoteS[k] = P.remoteS[k];
    a.remoteArrived[k] = P.remoteArrived[k];
  }
  a.exch = P.exch.p; a.ticket = P.ticket.p; a.publish = publish;
  a.handshake = a.nAux > 0 ? 1 : 0;
  a.nranks = g.nranks; a.ready = P.arrived.p + g.nranks;
  for (int k = 0; k < a.nSendPeers; ++k) a.sendRank[k] = P.sendRank[k];
  a.timeoutNs = g.timeoutNs; a.errFlag = g.errFlag.p;
  // pushed: the sweep stored the links itself (sweep_push.cu), only the publishing is left
  if (pushed) MUSB_TRY(launchSignalHalo(a, st));
  else MUSB_TRY(launchPushHalo(a, st));
  ++g.launches;
  P.pendingWait = true;
  return 0;
}

static int exchange(Level &L, int kind, double *state, int nComp, cudaStream_t st = nullptr,
                    bool pushed = false, bool deferWait = false) {
  if (g.nranks == 1) return 0;
  if (!st) st = g.stream;
  CommBuf &s = L.send[kind], &r = L.recv[kind];
  if (s.total == 0 && r.total == 0) return 0;
  if (kind == MUSB200_BUF_HALO && L.p2p.on && state == L.state[L.nNext].p) {
    {
      Timed t(T_COMM, st);
      MUSB_TRY(pushHalo(L, false, st, pushed));
    }
    // deferWait: the next sweep's halo CTAs wait (or ensureArrived at the end of the call)
    if (!deferWait) MUSB_TRY(ensureArrived(L, st));
    return 0;
  }
  Timed t(T_COMM, st);
  // the auxField travels through its own position lists (four entries per element)
  const bool isAux = (state == L.aux.p);
  const int32_t *sPos = isAux ? s.auxPos.p : s.pos.p, *rPos = isAux ? r.auxPos.p : r.pos.p;
  const int sTotal = isAux ? s.auxTotal : s.total, rTotal = isAux ? r.auxTotal : r.total;
  const std::vector<int> &sN = isAux ? s.auxNVals : s.nVals, &sO = isAux ? s.auxOffset : s.offset;
  const std::vector<int> &rN = isAux ? r.auxNVals : r.nVals, &rO = isAux ? r.auxOffset : r.offset;
  if (sTotal) {
    MUSB_TRY(launchPack(nComp, state, L.S, sPos, sTotal, s.buf.p, st));
    ++g.launches;
  }
  MUSB_NCCL(g.nccl->GroupStart());
  for (size_t i = 0; i < s.proc.size(); ++i)
    MUSB_NCCL(g.nccl->Send(s.buf.p + sO[i], (size_t)sN[i], ncclDouble, s.proc[i], g.comm, st));
  for (size_t i = 0; i < r.proc.size(); ++i)
    MUSB_NCCL(g.nccl->Recv(r.buf.p + rO[i], (size_t)rN[i], ncclDouble, r.proc[i], g.comm, st));
  MUSB_NCCL(g.nccl->GroupEnd());
  if (rTotal) {
    MUSB_TRY(launchUnpack(nComp, state, L.S, rPos, rTotal, r.buf.p, st));
    ++g.launches;
  }
  return 0;
}

// multi-level: state(:, next) and auxField of the halo elements in ONE message pair per peer
// (the reference sends them separately, tags iLevel and iLevel + 100): two pack launches, one NCCL
// group, two unpack launches instead of two complete exchanges
static int exchangeStateAndAux(Level &L, int kind = MUSB200_BUF_HALO) {
  if (g.nranks == 1) return 0;
  CommBuf &s = L.send[kind], &r = L.recv[kind];
  if (s.total == 0 && r.total == 0) return 0;
  cudaStream_t st = g.stream;
  if (kind == MUSB200_BUF_HALO && L.p2p.on) {
    // peer memory: state links and auxField entries in ONE kernel; the interpolation that
    // follows reads halo rows, so the wait comes right behind it
    {
      Timed t(T_COMM, st);
      MUSB_TRY(pushHalo(L, true, st, false));
    }
    return ensureArrived(L, st);
  }
  Timed t(T_COMM, st);
  double *state = L.state[L.nNext].p;
  if (s.total) {
    MUSB_TRY(launchPack(L.QQ, state, L.S, s.pos.p, s.total, s.buf.p, st));
    MUSB_TRY(launchPack(4, L.aux.p, L.S, s.auxPos.p, s.auxTotal, s.buf.p + s.total, st));
    g.launches += 2;
  }
  MUSB_NCCL(g.nccl->GroupStart());
  for (size_t i = 0; i < s.proc.size(); ++i) {
    MUSB_NCCL(g.nccl->Send(s.buf.p + s.offset[i], (size_t)s.nVals[i], ncclDouble, s.proc[i], g.comm, st));
    MUSB_NCCL(g.nccl->Send(s.buf.p + s.total + s.auxOffset[i], (size_t)s.auxNVals[i], ncclDouble, s.proc[i],
                           g.comm, st));
  }
  for (size_t i = 0; i < r.proc.size(); ++i) {
    MUSB_NCCL(g.nccl->Recv(r.buf.p + r.offset[i], (size_t)r.nVals[i], ncclDouble, r.proc[i], g.comm, st));
    MUSB_NCCL(g.nccl->Recv(r.buf.p + r.total + r.auxOffset[i], (size_t)r.auxNVals[i], ncclDouble, r.proc[i],
                           g.comm, st));
  }
  MUSB_NCCL(g.nccl->GroupEnd());
  if (r.total) {
    MUSB_TRY(launchUnpack(L.QQ, state, L.S, r.pos.p, r.total, r.buf.p, st));
    MUSB_TRY(launchUnpack(4, L.aux.p, L.S, r.auxPos.p, r.auxTotal, r.buf.p + r.total, st));
    g.launches += 2;
  }
  return 0;
}

// SWEEP_ALL: every CTA in natural order; SWEEP_HALO_LAST (overlapped exchange): the CTAs that pull
// from a halo row are moved to the end of the launch and wait there for the halo links
enum { SWEEP_ALL = 0, SWEEP_HALO_LAST = 1 };
static int sweep(Level &L, bool writeAux, int part = SWEEP_ALL, bool push = false) {
  if (!L.relaxSet) return setError(MUSB200_ERR_STATE, "musb200_set_relaxation missing");
  Timed t(T_COMPUTE);
  if (L.kind == MUSB200_KIND_PASSIVE_SCALAR) {
    if (part != SWEEP_ALL) return setError(MUSB200_ERR_UNSUPPORTED, "passive scalar: no split sweep");
    MUSB_TRY(ensureArrived(L));
    PsArgs p{};
    p.in = L.state[L.nNow].p; p.out = L.state[L.nNext].p; p.nbr = L.nbr.p; p.aux = L.aux.p;
    p.S = L.S; p.count = L.nSolve; p.write_aux = writeAux ? 1 : 0;
    p.d_omega = L.psDOmega; p.aux_omega = L.psAuxOmega;
    for (int k = 0; k < 3; ++k) p.vel_uniform[k] = L.velUniform[k];
    if (L.velMode == 1) { p.vel = L.vel.p; p.velS = L.S; }
    if (L.velMode == 2) {
      auto it = g.levelsOf[L.velSlot].find(L.velLevel);
      if (it == g.levelsOf[L.velSlot].end())
        return setError(MUSB200_ERR_STATE, "passive scalar: the coupled flow level no longer exists");
      p.vel = it->second->aux.p + it->second->S;   // rows ux, uy, uz of the flow's auxField
      p.velS = it->second->S;
    }
    MUSB_TRY(launchPassiveScalar(L.QQ, L.psVariant, p, g.stream));
    ++g.launches;
    return 0;
  }
  SweepArgs a{};
  a.in = L.state[L.nNow].p;
  a.out = L.state[L.nNext].p;
  a.nbr = L.nbr.p;
  a.aux = L.aux.p;
  a.omega = L.elemOmega ? L.omega.p : nullptr;
  a.ctaList = nullptr;
  a.ctaMode = 0;
  a.nMain = 0;
  a.nCtas = 0;
  a.S = L.S;
  a.first = 0;
  a.count = L.nSolve;
  a.write_aux = writeAux ? 1 : 0;
  a.rp = L.rp;
  a.force_order = L.forceOrder;
  a.force = L.forceElem ? L.force.p : nullptr;
  for (int k = 0; k < 3; ++k) a.force_uniform[k] = L.forceUniform[k];
  if (push) {
    PeerLink &P = L.p2p;
    a.push.mask = P.pushMask.p; a.push.prefix = P.pushPrefix.p; a.push.start = P.pushStart.p;
    a.push.entQ = P.pushQ.p; a.push.entPeer = P.pushPeer.p; a.push.entDst = P.pushDst.p;
    for (size_t k = 0; k < P.sendRank.size(); ++k) {
      a.push.remoteState[k] = P.remoteState[k][L.nNext];   // ranks swap now/next in lockstep
      a.push.remoteS[k] = P.remoteS[k];
    }
  }
  // the previous step's halo links may still be in flight: the CTAs that need them wait
  if (part == SWEEP_HALO_LAST) {
    a.wait = haloWait(L, true);
    a.ctaList = L.p2p.haloCtas.p; a.ctaMode = 1; a.nMain = L.p2p.nCtas; a.nCtas = L.p2p.nHaloCtas;
    L.p2p.pendingWait = false;       // the appended CTAs do the wait
  } else if (L.p2p.pendingWait && L.p2p.sweepWait && g.sweepWait) {
    a.wait = haloWait(L, true);
    L.p2p.pendingWait = false;
  } else {
    MUSB_TRY(ensureArrived(L));
  }
  MUSB_TRY(launchSweep(L.QQ, L.relax, L.kind, a, g.stream));
  ++g.launches;
  return 0;
}

static IntpArgs intpArgs(Level &src, Level &tgt) {
  IntpArgs a{};
  a.QQ = tgt.QQ;
  a.incomp = tgt.kind == MUSB200_KIND_FLUID_INCOMPRESSIBLE;
  a.sState = src.state[src.nNext].p;
  a.sAux = src.aux.p;
  a.sS = src.S;
  a.tState = tgt.state[tgt.nNext].p;
  a.tAux = tgt.aux.p;
  a.tS = tgt.S;
  a.tVisc = tgt.elemVisc ? tgt.visc.p : nullptr;
  // fluid%viscKine%dataOnLvl(level): from musb200_set_viscosity, else derived from omega
  a.tViscUniform = tgt.viscSet ? tgt.viscUniform : (1.0 / tgt.rp.omega_uniform - 0.5) / 3.0;
  return a;
}

static int applyIntp(Level &src, Level &tgt, IntpSet &set, bool fromFiner, bool withAux = false) {
  if (set.nTargets == 0) return 0;
  if (!tgt.viscSet && tgt.elemOmega && tgt.kind != MUSB200_KIND_PASSIVE_SCALAR)
    return setError(MUSB200_ERR_STATE, "per-element omega needs musb200_set_viscosity for interpolation");
  Timed t(T_INTP);
  int n = 0;
  IntpArgs ia = intpArgs(src, tgt);
  ia.withAux = withAux;
  // a passive scalar's ghosts: the reference's interpolation of arbitrary values applied to the
  // PDFs themselves (fillArbiMyGhostsFromFiner_avg / fillArbiFinerGhostsFromMe_*), no f_eq / f_neq split
  ia.passive = tgt.kind == MUSB200_KIND_PASSIVE_SCALAR;
  MUSB_TRY(launchIntp(ia, set, fromFiner, g.stream, &n));
  g.launches += n;
  return 0;
}

// one level step of ONE scheme without the ghost interpolation: set_boundary -> swap -> fused
// auxField + stream + collide -> halo exchange (steps 3-9 of do_fast_singleLevel / the body of
// do_recursive_multiLevel, mus/source/mus_control_module.f90:242-497, 507-701)
static int levelAdvance(int iLevel, int minLevel, int maxLevel, bool lastCycle) {
  Level *Lp = findLevel(iLevel);
  if (!Lp) return setError(MUSB200_ERR_ARG, "level " + std::to_string(iLevel) + " was not created");
  Level &L = *Lp;
  const bool multi = (maxLevel > minLevel);
  const bool passive = L.kind == MUSB200_KIND_PASSIVE_SCALAR;
  MUSB_TRY(setBoundary(L));
  if (!g.capturing && L.bcElems.n) MUSB_CUDA(cudaEventRecord(g.evBcDone, g.stream));
  std::swap(L.nNow, L.nNext);
  // multi-level: the interpolation routines read auxField of their sources every step
  // auxEveryStep: 0 = on the last step of a call, 1 = every step, 2 = lazy (only where the
  // schedule reads it; musb200_aux_probe / _download compute it on demand)
  const bool writeAux = g.auxEveryStep == 1 || multi || (lastCycle && g.auxEveryStep != 2) || L.auxForBc;
  L.auxValid = writeAux;
  if (!multi && g.nranks > 1 && g.overlap && L.p2p.on && L.p2p.nHaloCtas > 0 && !passive &&
      (L.send[MUSB200_BUF_HALO].total > 0 || L.recv[MUSB200_BUF_HALO].total > 0)) {
    // single level, several ranks, exchange overlapped with compute WITHOUT giving up coalescing:
    // the push of step n runs on the communication stream while step n+1 is being swept.  The
    // CTAs that pull from a halo row (3-12 % of them) are moved to the END of that launch --
    // whole CTAs of consecutive elements, appended from a list -- and wait there for the peers'
    // links, which by then have long arrived; every other CTA runs in natural order at once.
    // (Round 1 split by an element list and lost coalescing; a second launch for the halo CTAs
    // costs five latency-bound waves: profiles/r02_multi_gpu.md.)  The reference exchanges strictly
    // after compute (mus_control_module.f90:605-649); the results are identical: the pushed links
    // are final once sweep n has finished, the boundary kernels of step n+1 touch other slots (a
    // link a boundary rewrites points at a wall, no rank pulls it), peers store into halo rows only.
    // No waiting CTA can starve a push: a rank's push n is enqueued (high-priority stream) when
    // its sweep n ends, the first waiting CTA of sweep n+1 is dispatched after all its other CTAs.
    PeerLink &P = L.p2p;
    // this step writes state(:, next): the push that last read that buffer (two steps ago) must be done
    if (P.pushedOnce[L.nNext]) MUSB_CUDA(cudaStreamWaitEvent(g.stream, P.evPushed[L.nNext], 0));
    MUSB_TRY(sweep(L, writeAux, SWEEP_HALO_LAST));
    // the exchange number advances here, in the stepping stream: the next wait reads it in
    // stream order, whenever the push on the other stream gets to run
    MUSB_TRY(launchBumpExch(P.exch.p, P.exch.p + 1 + L.nNext, g.stream));
    ++g.launches;
    MUSB_CUDA(cudaEventRecord(P.evSwept, g.stream));
    MUSB_CUDA(cudaStreamWaitEvent(g.commStream, P.evSwept, 0));
    {
      Timed t(T_COMM, g.commStream);
      MUSB_TRY(pushHalo(L, false, g.commStream, false, P.exch.p + 1 + L.nNext));       // pendingWait = true
    }
    MUSB_CUDA(cudaEventRecord(P.evPushed[L.nNext], g.commStream));
    P.pushedOnce[L.nNext] = true;
    g.commBusy = true;
    return 0;
  }
  // single level on several ranks with the peer-memory exchange: the sweep pushes the halo links
  // itself (the force-source variant of the sweep keeps the separate push kernel)
  const bool pushed = !multi && g.nranks > 1 && g.fusedPush && L.p2p.on && L.forceOrder == 0 &&
                      !passive && L.p2p.pushMask.n > 0;
  MUSB_TRY(sweep(L, writeAux, SWEEP_ALL, pushed));
  // auxField of my ghostFromFiner elements <- average of level+1
  // (mus_intpAuxFieldCoarserAndExchange, mus_auxField_module.f90:404-444) is taken inside the
  // from-finer interpolation kernel below: same sources, and on one rank nothing reads those
  // entries before the from-coarser interpolation, which runs after it
  if (multi && writeAux && !passive) MUSB_TRY(exchangeStateAndAux(L));   // state (tag level) + aux (tag level+100)
  else MUSB_TRY(exchange(L, MUSB200_BUF_HALO, L.state[L.nNext].p, L.QQ, nullptr, pushed, /*deferWait=*/!multi));
  if (iLevel > minLevel) MUSB_TRY(exchange(L, MUSB200_BUF_FROMCOARSER, L.state[L.nNext].p, L.QQ));
  return 0;
}

// the ghost interpolation that closes a level step (do_intpFinerAndExchange,
// do_intpCoarserAndExchange, mus_control_module.f90:861-1051)
static int levelInterpolate(int iLevel, int maxLevel) {
  if (iLevel >= maxLevel) return 0;
  Level &L = *findLevel(iLevel);
  Level *F = findLevel(iLevel + 1);
  const bool passive = L.kind == MUSB200_KIND_PASSIVE_SCALAR;
  // do_intpFinerAndExchange: my ghostFromFiner <- average over the children on level+1
  MUSB_TRY(applyIntp(*F, L, L.fromFiner, true, !passive));
  // ghostFromFiner elements another rank interpolated for me arrive with their auxField entries
  // (the from-coarser interpolation below reads them when such a ghost is one of its sources)
  if (passive) MUSB_TRY(exchange(L, MUSB200_BUF_FROMFINER, L.state[L.nNext].p, L.QQ));
  else MUSB_TRY(exchangeStateAndAux(L, MUSB200_BUF_FROMFINER));
  // do_intpCoarserAndExchange: ghostFromCoarser of level+1 <- me, orders 0..order
  for (auto &set : F->fromCoarser) MUSB_TRY(applyIntp(L, *F, set, false));
  MUSB_TRY(exchange(*F, MUSB200_BUF_FROMCOARSER, F->state[F->nNext].p, F->QQ));
  return 0;
}

// the schemes stepped together (default: the bound slot alone).  Several schemes on one mesh
// -- a passive scalar transported by a flow, BASELINE config 5 -- advance level step by level
// step in the order given, so that the scalar's sweep reads the auxField the flow's sweep of the
// SAME level step has just written; then each scheme interpolates its ghosts.
static int gStepSlots[Context::kSlots] = {0, 0, 0, 0};
static int gNStepSlots = 0;   // 0: the bound slot only

// every scheme stepped by this call, one after the other, with `slot` bound
template <class F>
static int forEachStepSlot(F f) {
  if (gNStepSlots == 0) return f();
  const int keep = g.slot;
  int rc = 0;
  for (int k = 0; k < gNStepSlots && rc == 0; ++k) { g.slot = gStepSlots[k]; rc = f(); }
  g.slot = keep;
  return rc;
}


static int levelStep(int iLevel, int minLevel, int maxLevel, bool lastCycle) {
  if (!findLevel(iLevel)) return setError(MUSB200_ERR_ARG, "level " + std::to_string(iLevel) + " was not created");
  if (iLevel < maxLevel) {
    // nNesting = 2 (acoustic scaling, mus_param_module.f90:191-195)
    for (int n = 0; n < 2; ++n) MUSB_TRY(levelStep(iLevel + 1, minLevel, maxLevel, lastCycle && n == 1));
  }
  if (gNStepSlots == 0) {
    MUSB_TRY(levelAdvance(iLevel, minLevel, maxLevel, lastCycle));
    return levelInterpolate(iLevel, maxLevel);
  }
  const int keep = g.slot;
  int rc = 0;
  for (int k = 0; k < gNStepSlots && rc == 0; ++k) { g.slot = gStepSlots[k]; rc = levelAdvance(iLevel, minLevel, maxLevel, lastCycle); }
  for (int k = 0; k < gNStepSlots && rc == 0; ++k) { g.slot = gStepSlots[k]; rc = levelInterpolate(iLevel, maxLevel); }
  g.slot = keep;
  return rc;
}

// Failure detection at the points where the host synchronises anyway (the reference aborts all
// ranks through tem_abort, tem/source/tem_aux_module.f90:457-478): a wait of the peer-memory
// exchange that gave up (p2p.cu) and NCCL's asynchronous communicator errors.  The stream has
// been synchronised by the caller.
static int checkAsyncErrors() {
  if (g.nranks == 1) return 0;
  int h[4] = {0, 0, 0, 0};
  MUSB_CUDA(cudaMemcpy(h, g.errFlag.p, sizeof(h), cudaMemcpyDeviceToHost));
  if (h[0] != 0) {
    const unsigned long long n = ((unsigned long long)(unsigned int)h[3] << 32) | (unsigned int)h[2];
    return setError(MUSB200_ERR_NCCL, std::string("halo exchange timed out after ") +
                                          std::to_string(g.timeoutNs / 1000000ull) + " ms: rank " +
                                          std::to_string(h[1]) + (h[0] == 1 ? " did not deliver" : " was not ready for") +
                                          " exchange " + std::to_string(n) + " (rank " + std::to_string(g.rank) + ")");
  }
  if (g.comm) {
    ncclResult_t ar = ncclSuccess;
    if (g.nccl->CommGetAsyncError(g.comm, &ar) == ncclSuccess && ar != ncclSuccess && ar != ncclInProgress)
      return setError(MUSB200_ERR_NCCL, std::string("NCCL communicator error: ") + g.nccl->GetErrorString(ar));
  }
  return 0;
}

// mus_init_flow after the state has been filled (mus/source/mus_flow_module.fpp:206-240):
// mus_initAuxField (:1677-1737), fillHelperElementsFineToCoarse (:1517-1588),
// fillHelperElementsCoarseToFine (:1601-1673)
static int fillFineToCoarse(int iLevel, int minLevel, int maxLevel) {
  Level *L = findLevel(iLevel);
  if (!L) return setError(MUSB200_ERR_ARG, "level " + std::to_string(iLevel) + " was not created");
  const bool multi = maxLevel > minLevel;
  if (iLevel < maxLevel) {
    MUSB_TRY(fillFineToCoarse(iLevel + 1, minLevel, maxLevel));
    Level *F = findLevel(iLevel + 1);
    // state and auxField of my ghostFromFiner elements <- average over the children
    // (do_intp of fillMineFromFiner + mus_intpAuxFieldCoarserAndExchange)
    MUSB_TRY(applyIntp(*F, *L, L->fromFiner, true, L->kind != MUSB200_KIND_PASSIVE_SCALAR));
    MUSB_TRY(exchangeStateAndAux(*L, MUSB200_BUF_FROMFINER));
  }
  if (multi || L->nAux == 4) MUSB_TRY(exchangeStateAndAux(*L));
  else MUSB_TRY(exchange(*L, MUSB200_BUF_HALO, L->state[L->nNext].p, L->QQ));
  return 0;
}
static int fillCoarseToFine(int iLevel, int minLevel, int maxLevel) {
  Level *L = findLevel(iLevel);
  if (iLevel > minLevel) MUSB_TRY(exchange(*L, MUSB200_BUF_FROMCOARSER, L->state[L->nNext].p, L->QQ));
  if (iLevel < maxLevel) {
    Level *F = findLevel(iLevel + 1);
    for (auto &set : F->fromCoarser) MUSB_TRY(applyIntp(*L, *F, set, false));
    MUSB_TRY(fillCoarseToFine(iLevel + 1, minLevel, maxLevel));
  }
  return 0;
}

}  // namespace musb200

using namespace musb200;

// ===========================================================================
extern "C" {

int musb200_last_error(char *buf, int buflen) {
  if (!buf || buflen <= 0) return MUSB200_ERR_ARG;
  std::strncpy(buf, lastError().c_str(), (size_t)buflen - 1);
  buf[buflen - 1] = 0;
  return 0;
}

int musb200_device_count(int *n) {
  if (!n) return setError(MUSB200_ERR_ARG, "null argument");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    *n = 0;
    return setError(MUSB200_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *n = c;
  return 0;
}

int musb200_get_unique_id(void *out128) {
  if (!out128) return setError(MUSB200_ERR_ARG, "null argument");
  NcclApi *api = ncclApi();
  if (!api) return MUSB200_ERR_NCCL;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return setError(MUSB200_ERR_NCCL, "ncclGetUniqueId failed");
  std::memcpy(out128, &id, 128);
  return 0;
}

int musb200_init(int rank, int nranks, int local_device, const void *nccl_unique_id) {
  if (g.ready) return setError(MUSB200_ERR_STATE, "musb200_init called twice");
  if (nranks < 1 || rank < 0 || rank >= nranks) return setError(MUSB200_ERR_ARG, "bad rank/nranks");
  int ndev = 0;
  MUSB_TRY(musb200_device_count(&ndev));
  if (ndev == 0) return setError(MUSB200_ERR_CUDA, "no CUDA device: libmusb200 has no CPU fallback");
  if (local_device < 0 || local_device >= ndev) return setError(MUSB200_ERR_ARG, "bad device index");
  MUSB_CUDA(cudaSetDevice(local_device));
  cudaDeviceProp prop;
  MUSB_CUDA(cudaGetDeviceProperties(&prop, local_device));
  if (prop.major < 10)
    return setError(MUSB200_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100 class");
  g.rank = rank; g.nranks = nranks; g.device = local_device;
  MUSB_CUDA(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    MUSB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    MUSB_CUDA(cudaStreamCreateWithPriority(&g.commStream, cudaStreamNonBlocking, hi));
    MUSB_CUDA(cudaEventCreateWithFlags(&g.evBoundary, cudaEventDisableTiming));
    MUSB_CUDA(cudaEventCreateWithFlags(&g.evComm, cudaEventDisableTiming));
    MUSB_CUDA(cudaStreamCreateWithFlags(&g.copyStream, cudaStreamNonBlocking));
    MUSB_CUDA(cudaEventCreateWithFlags(&g.evBcDone, cudaEventDisableTiming));
    MUSB_CUDA(cudaEventRecord(g.evBcDone, g.stream));
  }
  MUSB_CUDA(cudaEventCreate(&g.mark[0]));
  MUSB_CUDA(cudaEventCreate(&g.mark[1]));
  MUSB_TRY(g.red.alloc(3 * 592 + 8));
  MUSB_TRY(g.flag.alloc(1));
  MUSB_TRY(g.errFlag.alloc(4));
  MUSB_CUDA(cudaMemsetAsync(g.errFlag.p, 0, 4 * sizeof(int), g.stream));
  if (nranks > 1) {
    if (!nccl_unique_id) return setError(MUSB200_ERR_ARG, "nranks > 1 needs the NCCL unique id");
    g.nccl = ncclApi();
    if (!g.nccl) return MUSB200_ERR_NCCL;
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, 128);
    MUSB_NCCL(g.nccl->CommInitRank(&g.comm, nranks, id, rank));
  }
  g.ready = true;
  return 0;
}

int musb200_finalize(void) {
  if (!g.ready) return 0;
  cudaStreamSynchronize(g.stream);
  dropGraph();
  for (auto &m : g.levelsOf) m.clear();
  g.slot = 0;
  g.stage.release(); g.red.release(); g.flag.release(); g.errFlag.release();
  for (auto &s : g.spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  g.spans.clear();
  for (auto e : g.pool) cudaEventDestroy(e);
  g.pool.clear();
  if (g.comm) { g.nccl->CommDestroy(g.comm); g.comm = nullptr; }
  cudaEventDestroy(g.mark[0]); cudaEventDestroy(g.mark[1]);
  cudaEventDestroy(g.evBoundary); cudaEventDestroy(g.evComm);
  cudaStreamSynchronize(g.commStream);
  cudaStreamDestroy(g.commStream);
  cudaStreamSynchronize(g.copyStream);
  cudaStreamDestroy(g.copyStream);
  cudaEventDestroy(g.evBcDone);
  cudaStreamDestroy(g.stream);
  g.stream = nullptr; g.commStream = nullptr; g.copyStream = nullptr;
  g.ready = false;
  g.launches = 0;
  return 0;
}

// ---------------------------------------------------------------------------
int musb200_scheme_select(const char *kind, const char *relaxation, const char *variant,
                          const char *layout, int *relax_id, int *kind_id, int *QQ) {
  if (!kind || !relaxation || !layout || !relax_id || !kind_id || !QQ)
    return setError(MUSB200_ERR_ARG, "null argument");
  const std::string k(kind), r(relaxation), l(layout), v(variant ? variant : "standard");
  int kk, rr, qq;
  if (k == "fluid") kk = MUSB200_KIND_FLUID;
  else if (k == "fluid_incompressible") kk = MUSB200_KIND_FLUID_INCOMPRESSIBLE;
  else if (k == "passive_scalar") kk = MUSB200_KIND_PASSIVE_SCALAR;
  else return setError(MUSB200_ERR_UNSUPPORTED, "scheme kind '" + k + "' is outside the B200 hot path");
  if (l == "d3q19") qq = 19;
  else if (l == "d3q27") qq = 27;
  else return setError(MUSB200_ERR_UNSUPPORTED, "layout '" + l + "' is outside the B200 hot path");
  if (r == "bgk") rr = MUSB200_RELAX_BGK;
  else if (r == "trt") rr = MUSB200_RELAX_TRT;
  else if (r == "mrt") rr = MUSB200_RELAX_MRT;
  else return setError(MUSB200_ERR_UNSUPPORTED, "relaxation '" + r + "' is outside the B200 hot path");
  if (kk == MUSB200_KIND_PASSIVE_SCALAR) {
    // mus_init_advRel_lbm_ps (init/mus_initLBMPS_module.f90:59-159): bgk first|second for any
    // layout, trt without a named variant = vStdNoOpt; the E/L-model variants and mrt are d3q19
    // special kernels outside the hot path
    const bool ok = (rr == MUSB200_RELAX_BGK && (v == "first" || v == "second")) ||
                    (rr == MUSB200_RELAX_TRT && v != "Emodel" && v != "EmodelCorr" && v != "Lmodel");
    if (!ok)
      return setError(MUSB200_ERR_UNSUPPORTED, "passive_scalar: relaxation '" + r + "' variant '" + v +
                                                   "' is outside the B200 hot path");
    *relax_id = rr; *kind_id = kk; *QQ = qq;
    return 0;
  }
  if (v != "standard" && v != "b200")
    return setError(MUSB200_ERR_UNSUPPORTED, "relaxation variant '" + v + "' is outside the B200 hot path");
  // mus_init_advRel_fluid_incompressible aborts for trt with a layout other than d3q19
  // (init/mus_initFluidIncomp_module.f90:80-93)
  if (kk == MUSB200_KIND_FLUID_INCOMPRESSIBLE && rr == MUSB200_RELAX_TRT && qq != 19)
    return setError(MUSB200_ERR_UNSUPPORTED,
                    "fluid_incompressible: the reference has no trt kernel for layout '" + l + "'");
  *relax_id = rr; *kind_id = kk; *QQ = qq;
  return 0;
}

// ---------------------------------------------------------------------------
int musb200_level_create(int level, int QQ, int nScalars, int nAuxScalars, int nSize, int nFluid,
                         int nGhostFromCoarser, int nGhostFromFiner, int nHalo,
                         const int32_t *neigh, const int64_t *property, const int64_t *treeID) {
  (void)property; (void)treeID;
  MUSB_TRY(needReady());
  if (QQ != 19 && QQ != 27) return setError(MUSB200_ERR_UNSUPPORTED, "only d3q19 / d3q27");
  if (nScalars != QQ) return setError(MUSB200_ERR_UNSUPPORTED, "multi-field schemes (nScalars != QQ)");
  if (nAuxScalars != 4 && nAuxScalars != 1)
    return setError(MUSB200_ERR_UNSUPPORTED, "nAuxScalars must be 4 (rho, u) or 1 (passive scalar)");
  const long long nElems = (long long)nFluid + nGhostFromCoarser + nGhostFromFiner + nHalo;
  if (nSize < nElems || nSize <= 0 || !neigh) return setError(MUSB200_ERR_ARG, "bad sizes / null neigh");
  if ((long long)nSize > (long long)kElemMask) return setError(MUSB200_ERR_ARG, "nSize exceeds 2^31-1");
  if ((long long)nSize * QQ > 2147483647LL)
    return setError(MUSB200_ERR_ARG, "nSize*QQ exceeds the 32-bit state positions of the host lists");
  auto L = std::make_unique<Level>();
  L->level = level; L->QQ = QQ; L->nSize = nSize; L->nFluid = nFluid;
  L->nGFC = nGhostFromCoarser; L->nGFF = nGhostFromFiner; L->nHalo = nHalo;
  L->nElems = (int)nElems;
  L->nAux = nAuxScalars;
  L->nSolve = nFluid + nGhostFromCoarser;  // mus_pdf_module.f90:125
  L->S = ((long long)nSize + 31) / 32 * 32;
  for (int b = 0; b < 2; ++b) {
    MUSB_TRY(L->state[b].alloc((size_t)L->S * QQ));
    MUSB_CUDA(cudaMemsetAsync(L->state[b].p, 0, (size_t)L->S * QQ * sizeof(double), g.stream));
  }
  MUSB_TRY(L->aux.alloc((size_t)L->S * 4));
  MUSB_CUDA(cudaMemsetAsync(L->aux.p, 0, (size_t)L->S * 4 * sizeof(double), g.stream));
  MUSB_TRY(L->nbr.alloc((size_t)L->S * (QQ - 1)));
  MUSB_CUDA(cudaMemsetAsync(L->nbr.p, 0, (size_t)L->S * (QQ - 1) * sizeof(uint32_t), g.stream));
  // upload the Fortran list and encode it on the device
  DevBuf<int32_t> tmp;
  MUSB_TRY(tmp.upload(neigh, (size_t)nSize * QQ, g.stream));
  MUSB_CUDA(cudaMemsetAsync(g.flag.p, 0, sizeof(int), g.stream));
  MUSB_TRY(launchEncodeNeigh(QQ, tmp.p, L->nbr.p, nSize, L->nElems, L->S, g.flag.p, g.stream));
  int bad = 0;
  MUSB_CUDA(cudaMemcpyAsync(&bad, g.flag.p, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  if (bad)
    return setError(MUSB200_ERR_CONNECTIVITY,
                    std::to_string(bad) + " neigh entries are neither a plain pull nor a bounce-back");
  g.levels()[level] = std::move(L);
  ++g.epoch;
  return 0;
}

// treelm's predefined cube on one rank, connectivity generated on the device (cube.cu)
int musb200_level_create_cube(int level, int treeLevel, int QQ, int walls) {
  MUSB_TRY(needReady());
  if (QQ != 19 && QQ != 27) return setError(MUSB200_ERR_UNSUPPORTED, "only d3q19 / d3q27");
  if (treeLevel < 1 || treeLevel > 10) return setError(MUSB200_ERR_ARG, "cube: tree level must be 1..10");
  const long long n = 1LL << (3 * treeLevel);
  auto L = std::make_unique<Level>();
  L->level = level; L->QQ = QQ; L->nSize = (int)n; L->nFluid = (int)n;
  L->nElems = (int)n; L->nAux = 4; L->nSolve = (int)n;
  L->S = (n + 31) / 32 * 32;
  for (int b = 0; b < 2; ++b) {
    MUSB_TRY(L->state[b].alloc((size_t)L->S * QQ));
    MUSB_CUDA(cudaMemsetAsync(L->state[b].p, 0, (size_t)L->S * QQ * sizeof(double), g.stream));
  }
  MUSB_TRY(L->aux.alloc((size_t)L->S * 4));
  MUSB_CUDA(cudaMemsetAsync(L->aux.p, 0, (size_t)L->S * 4 * sizeof(double), g.stream));
  MUSB_TRY(L->nbr.alloc((size_t)L->S * (QQ - 1)));
  MUSB_CUDA(cudaMemsetAsync(L->nbr.p, 0, (size_t)L->S * (QQ - 1) * sizeof(uint32_t), g.stream));
  MUSB_TRY(launchCubeNeigh(QQ, L->nbr.p, treeLevel, walls ? 1 : 0, L->S, L->nElems, g.stream));
  ++g.launches;
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  g.levels()[level] = std::move(L);
  ++g.epoch;
  return 0;
}

// mus_init_pdf with zero strain rate: both state buffers <- f_eq(auxField)
int musb200_state_init_equilibrium(int level) {
  GET_LEVEL(L, level);
  if (!L->relaxSet || L->kind == MUSB200_KIND_PASSIVE_SCALAR)
    return setError(MUSB200_ERR_STATE, "equilibrium initial state: musb200_set_relaxation of a fluid scheme first");
  MUSB_TRY(launchInitEquilibrium(L->QQ, L->kind == MUSB200_KIND_FLUID_INCOMPRESSIBLE, L->aux.p, L->state[0].p,
                                 L->state[1].p, L->S, L->nElems, g.stream));
  ++g.launches;
  L->auxValid = true;
  return 0;
}

int musb200_level_destroy(int level) {
  MUSB_TRY(needReady());
  ++g.epoch;
  cudaStreamSynchronize(g.stream);
  g.levels().erase(level);
  return 0;
}

int musb200_neigh_download(int level, int32_t *neigh) {
  GET_LEVEL_RO(L, level);
  if (!neigh) return setError(MUSB200_ERR_ARG, "null argument");
  if ((long long)L->nSize * L->QQ > 2147483647LL)
    return setError(MUSB200_ERR_ARG, "nSize*QQ exceeds the 32-bit state positions of the host list");
  DevBuf<int32_t> tmp;
  MUSB_TRY(tmp.alloc((size_t)L->nSize * L->QQ));
  MUSB_CUDA(cudaMemsetAsync(tmp.p, 0, tmp.n * sizeof(int32_t), g.stream));
  MUSB_TRY(launchDecodeNeigh(L->QQ, L->nbr.p, tmp.p, L->nSize, L->nElems, L->S, g.stream));
  MUSB_CUDA(cudaMemcpyAsync(neigh, tmp.p, tmp.n * sizeof(int32_t), cudaMemcpyDeviceToHost, g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}

static int uploadAos(Level *L, const double *host, double *soa, int nComp) {
  const size_t n = (size_t)L->nSize * nComp;
  MUSB_TRY(stageBuf(n));
  MUSB_CUDA(cudaMemcpyAsync(g.stage.p, host, n * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  MUSB_TRY(launchAosToSoa(g.stage.p, soa, nComp, L->nSize, L->S, g.stream));
  ++g.launches;
  return 0;
}
static int downloadAos(Level *L, const double *soa, double *host, int nComp) {
  const size_t n = (size_t)L->nSize * nComp;
  MUSB_TRY(stageBuf(n));
  MUSB_TRY(launchSoaToAos(soa, g.stage.p, nComp, L->nSize, L->S, g.stream));
  ++g.launches;
  MUSB_CUDA(cudaMemcpyAsync(host, g.stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}

int musb200_state_upload(int level, int which, const double *aos_state) {
  GET_LEVEL(L, level);
  if (which < 1 || which > 2 || !aos_state) return setError(MUSB200_ERR_ARG, "which must be 1|2");
  return uploadAos(L, aos_state, L->state[which - 1].p, L->QQ);
}
int musb200_state_download(int level, int which, double *aos_state) {
  GET_LEVEL_RO(L, level);
  if (which < 1 || which > 2 || !aos_state) return setError(MUSB200_ERR_ARG, "which must be 1|2");
  return downloadAos(L, L->state[which - 1].p, aos_state, L->QQ);
}
int musb200_aux_upload(int level, const double *aos_aux) {
  GET_LEVEL(L, level);
  if (!aos_aux) return setError(MUSB200_ERR_ARG, "null argument");
  return uploadAos(L, aos_aux, L->aux.p, L->nAux);
}
// auxField of the elements [first, first + count) of the last level step, from state(:, now)
static int auxOnDemand(Level *L, int first, int count) {
  if (L->auxValid || !L->relaxSet || L->kind == MUSB200_KIND_PASSIVE_SCALAR) return 0;
  SweepArgs a{};
  a.in = L->state[L->nNow].p; a.nbr = L->nbr.p; a.aux = L->aux.p; a.S = L->S;
  a.first = first; a.count = count;
  a.force_order = L->forceOrder;
  a.force = L->forceElem ? L->force.p : nullptr;
  for (int k = 0; k < 3; ++k) a.force_uniform[k] = L->forceUniform[k];
  MUSB_TRY(launchAuxOnly(L->QQ, L->kind, a, g.stream));
  ++g.launches;
  return 0;
}

int musb200_aux_download(int level, double *aos_aux) {
  GET_LEVEL_RO(L, level);
  if (!aos_aux) return setError(MUSB200_ERR_ARG, "null argument");
  MUSB_TRY(auxOnDemand(L, 0, L->nSolve));
  if (L->relaxSet && L->kind != MUSB200_KIND_PASSIVE_SCALAR) L->auxValid = true;
  return downloadAos(L, L->aux.p, aos_aux, L->nAux);
}
int musb200_aux_probe(int level, int elemPos, double *out) {
  GET_LEVEL_RO(L, level);
  if (!out || elemPos < 1 || elemPos > L->nElems) return setError(MUSB200_ERR_ARG, "bad probe element");
  if (elemPos <= L->nSolve) MUSB_TRY(auxOnDemand(L, elemPos - 1, 1));
  MUSB_CUDA(cudaMemcpy2DAsync(out, sizeof(double), L->aux.p + (elemPos - 1), (size_t)L->S * sizeof(double),
                              sizeof(double), (size_t)L->nAux, cudaMemcpyDeviceToHost, g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}
int musb200_state_copy_next_to_now(int level) {
  GET_LEVEL(L, level);
  MUSB_CUDA(cudaMemcpyAsync(L->state[L->nNow].p, L->state[L->nNext].p,
                            (size_t)L->S * L->QQ * sizeof(double), cudaMemcpyDeviceToDevice, g.stream));
  return 0;
}
int musb200_set_now_next(int level, int nNow, int nNext) {
  GET_LEVEL(L, level);
  if (!((nNow == 1 && nNext == 2) || (nNow == 2 && nNext == 1)))
    return setError(MUSB200_ERR_ARG, "nNow/nNext must be a permutation of 1,2");
  L->nNow = nNow - 1; L->nNext = nNext - 1;
  return 0;
}
int musb200_get_now_next(int level, int *nNow, int *nNext) {
  GET_LEVEL_RO(L, level);
  if (nNow) *nNow = L->nNow + 1;
  if (nNext) *nNext = L->nNext + 1;
  return 0;
}

int musb200_set_relaxation(int level, int relax_id, int kind_id, const double *omega,
                           double omega_uniform, double lambda, double omega_bulk) {
  GET_LEVEL(L, level);
  if (relax_id < 0 || relax_id > 2 || kind_id < 0 || kind_id > 1)
    return setError(MUSB200_ERR_ARG, "bad relaxation / kind id (passive scalar: musb200_set_species)");
  if (L->nAux != 4) return setError(MUSB200_ERR_ARG, "a fluid scheme needs nAuxScalars = 4");
  if (kind_id == MUSB200_KIND_FLUID_INCOMPRESSIBLE && relax_id == MUSB200_RELAX_TRT && L->QQ != 19)
    return setError(MUSB200_ERR_UNSUPPORTED, "fluid_incompressible: the reference has no trt kernel for d3q27");
  L->relax = relax_id; L->kind = kind_id;
  L->rp.omega_uniform = omega_uniform; L->rp.lambda = lambda; L->rp.omega_bulk = omega_bulk;
  L->elemOmega = (omega != nullptr);
  if (omega) {
    if (L->omega.n < (size_t)L->S) MUSB_TRY(L->omega.alloc((size_t)L->S));
    MUSB_CUDA(cudaMemcpyAsync(L->omega.p, omega, (size_t)L->nSolve * sizeof(double),
                              cudaMemcpyHostToDevice, g.stream));
  }
  L->relaxSet = true;
  return 0;
}

// ---------------------------------------------------------------------------
// restart: the chunk of the global treeID list goes through the device level by level
static int levelOfTreeID(long long id) {   // tem_LevelOf (tem_topology_module.f90): first id of
  int level = 0;                            // level L is (8^L - 1) / 7
  long long first = 0, count = 1;
  while (id >= first + count) { first += count; count *= 8; ++level; }
  return level;
}
static int serializeChunk(int nElems, const int64_t *treeID, const int32_t *levelPointer, double *out,
                          const double *in) {
  MUSB_TRY(needReady());
  if (nElems < 0 || (nElems > 0 && (!treeID || !levelPointer || (!out && !in))))
    return setError(MUSB200_ERR_ARG, "bad restart chunk");
  if (nElems == 0) return 0;
  std::map<int, std::pair<std::vector<int32_t>, std::vector<int32_t>>> byLevel;  // slot, elemPos
  for (int i = 0; i < nElems; ++i) {
    auto &b = byLevel[levelOfTreeID((long long)treeID[i])];
    b.first.push_back(i + 1);
    b.second.push_back(levelPointer[i]);
  }
  int QQ = 0;
  for (auto &kv : byLevel) {
    Level *L = findLevel(kv.first);
    if (!L) return setError(MUSB200_ERR_ARG, "restart chunk holds a treeID of level " +
                                                 std::to_string(kv.first) + ", which was not created");
    if (QQ && QQ != L->QQ) return setError(MUSB200_ERR_ARG, "levels with different stencils");
    QQ = L->QQ;
    for (int32_t p : kv.second.second)
      if (p < 1 || p > L->nElems) return setError(MUSB200_ERR_ARG, "levelPointer outside the level");
  }
  const size_t nVals = (size_t)nElems * QQ;
  MUSB_TRY(stageBuf(nVals));
  if (in) MUSB_CUDA(cudaMemcpyAsync(g.stage.p, in, nVals * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  for (auto &kv : byLevel) {
    Level *L = findLevel(kv.first);
    const int n = (int)kv.second.first.size();
    DevBuf<int32_t> slot, pos;
    MUSB_TRY(slot.upload(kv.second.first.data(), (size_t)n, g.stream));
    MUSB_TRY(pos.upload(kv.second.second.data(), (size_t)n, g.stream));
    double *st = L->state[L->nNext].p;          // restart reads and writes state(:, nNext)
    if (out) MUSB_TRY(launchSerialize(QQ, st, L->S, slot.p, pos.p, n, g.stage.p, g.stream));
    else MUSB_TRY(launchUnserialize(QQ, st, L->S, slot.p, pos.p, n, g.stage.p, g.stream));
    ++g.launches;
    MUSB_CUDA(cudaStreamSynchronize(g.stream));  // the index buffers are released here
  }
  if (out) {
    MUSB_CUDA(cudaMemcpyAsync(out, g.stage.p, nVals * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    MUSB_CUDA(cudaStreamSynchronize(g.stream));
  }
  return 0;
}
int musb200_pdf_serialize(int nElems, const int64_t *treeID, const int32_t *levelPointer, double *buffer) {
  return serializeChunk(nElems, treeID, levelPointer, buffer, nullptr);
}
int musb200_pdf_unserialize(int nElems, const int64_t *treeID, const int32_t *levelPointer,
                            const double *buffer) {
  return serializeChunk(nElems, treeID, levelPointer, nullptr, buffer);
}

// ---------------------------------------------------------------------------
int musb200_scheme_bind(int slot) {
  MUSB_TRY(needReady());
  if (slot < 0 || slot >= Context::kSlots) return setError(MUSB200_ERR_ARG, "scheme slot out of range");
  g.slot = slot;
  return 0;
}

int musb200_source_force(int level, int order, int nElems, const int32_t *posInTotal,
                         const double *force, int uniform) {
  GET_LEVEL(L, level);
  if (order == 0 || nElems == 0) { L->forceOrder = 0; return 0; }
  if ((order != 1 && order != 2) || nElems < 0 || !force) return setError(MUSB200_ERR_ARG, "bad force source");
  if (L->kind == MUSB200_KIND_PASSIVE_SCALAR)
    return setError(MUSB200_ERR_UNSUPPORTED, "force source on a passive-scalar scheme");
  if (nElems > L->nSolve) return setError(MUSB200_ERR_ARG, "force source: more elements than nElems_solve");
  if (uniform && !posInTotal && nElems == L->nSolve) {
    // the common case (global shape, constant force): no array, no extra traffic
    for (int k = 0; k < 3; ++k) L->forceUniform[k] = force[k];
    L->forceElem = false;
    L->forceOrder = order;
    return 0;
  }
  if (!posInTotal && nElems != L->nSolve)
    return setError(MUSB200_ERR_ARG, "force source: posInTotal missing for a partial element list");
  // dense SoA field, zero outside the source's elements (adding a zero force changes nothing)
  if (L->force.n < (size_t)L->S * 3) MUSB_TRY(L->force.alloc((size_t)L->S * 3));
  MUSB_CUDA(cudaMemsetAsync(L->force.p, 0, (size_t)L->S * 3 * sizeof(double), g.stream));
  std::vector<double> rep;
  const double *src = force;
  if (uniform) {
    rep.resize((size_t)nElems * 3);
    for (int i = 0; i < nElems; ++i)
      for (int k = 0; k < 3; ++k) rep[(size_t)i * 3 + k] = force[k];
    src = rep.data();
  }
  MUSB_TRY(stageBuf((size_t)nElems * 3));
  MUSB_CUDA(cudaMemcpyAsync(g.stage.p, src, (size_t)nElems * 3 * sizeof(double), cudaMemcpyHostToDevice,
                            g.stream));
  DevBuf<int32_t> pos;
  if (posInTotal) {
    for (int i = 0; i < nElems; ++i)
      if (posInTotal[i] < 1 || posInTotal[i] > L->nSolve)
        return setError(MUSB200_ERR_ARG, "force source: posInTotal outside 1..nElems_solve");
    MUSB_TRY(pos.upload(posInTotal, (size_t)nElems, g.stream));
  }
  MUSB_TRY(launchScatterRows(g.stage.p, posInTotal ? pos.p : nullptr, nElems, 3, L->force.p, L->S, g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.stream));   // rep / pos are released on return
  L->forceElem = true;
  L->forceOrder = order;
  return 0;
}

int musb200_set_species(int level, int relax_id, int variant, double diff_coeff, double lambda) {
  GET_LEVEL(L, level);
  if (L->nAux != 1) return setError(MUSB200_ERR_ARG, "passive scalar needs nAuxScalars = 1");
  int v;
  if (relax_id == MUSB200_RELAX_BGK && (variant == 1 || variant == 2)) v = variant;
  else if (relax_id == MUSB200_RELAX_TRT) v = 3;
  else return setError(MUSB200_ERR_UNSUPPORTED, "passive_scalar: bgk first|second or trt");
  L->kind = MUSB200_KIND_PASSIVE_SCALAR;
  L->relax = relax_id;
  L->psVariant = v;
  L->psDOmega = 2.0 / (1.0 + 6.0 * diff_coeff);
  L->psAuxOmega = 1.0 / (lambda / (1.0 / L->psDOmega - 0.5) + 0.5);
  L->relaxSet = true;
  return 0;
}

int musb200_set_transport_velocity(int level, int nElems, const double *vel, int uniform) {
  GET_LEVEL(L, level);
  if (!vel) return setError(MUSB200_ERR_ARG, "null argument");
  if (uniform) {
    for (int k = 0; k < 3; ++k) L->velUniform[k] = vel[k];
    L->velMode = 0;
    return 0;
  }
  if (nElems != L->nSolve) return setError(MUSB200_ERR_ARG, "transport velocity: one triple per solved element");
  if (L->vel.n < (size_t)L->S * 3) {
    MUSB_TRY(L->vel.alloc((size_t)L->S * 3));
    MUSB_CUDA(cudaMemsetAsync(L->vel.p, 0, (size_t)L->S * 3 * sizeof(double), g.stream));
  }
  MUSB_TRY(stageBuf((size_t)nElems * 3));
  MUSB_CUDA(cudaMemcpyAsync(g.stage.p, vel, (size_t)nElems * 3 * sizeof(double), cudaMemcpyHostToDevice,
                            g.stream));
  MUSB_TRY(launchAosToSoa(g.stage.p, L->vel.p, 3, nElems, L->S, g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.stream));   // vel is borrowed for the call only
  L->velMode = 1;
  return 0;
}

int musb200_couple_transport_velocity(int level, int flow_slot, int flow_level) {
  GET_LEVEL(L, level);
  if (flow_slot < 0 || flow_slot >= Context::kSlots || flow_slot == g.slot)
    return setError(MUSB200_ERR_ARG, "coupling: flow scheme slot out of range / same as the scalar's");
  auto it = g.levelsOf[flow_slot].find(flow_level);
  if (it == g.levelsOf[flow_slot].end()) return setError(MUSB200_ERR_ARG, "coupling: unknown flow level");
  if (it->second->nAux != 4 || it->second->nSolve < L->nSolve)
    return setError(MUSB200_ERR_ARG, "coupling: the flow level must hold rho,u for the same element list");
  L->velMode = 2; L->velSlot = flow_slot; L->velLevel = flow_level;
  return 0;
}

// ---------------------------------------------------------------------------
int musb200_set_viscosity(int level, const double *visc, double visc_uniform) {
  GET_LEVEL(L, level);
  L->elemVisc = (visc != nullptr);
  L->viscUniform = visc_uniform;
  if (visc) {
    if (L->visc.n < (size_t)L->S) MUSB_TRY(L->visc.alloc((size_t)L->S));
    MUSB_CUDA(cudaMemcpyAsync(L->visc.p, visc, (size_t)L->nElems * sizeof(double), cudaMemcpyHostToDevice,
                              g.stream));
  }
  L->viscSet = true;
  return 0;
}

int musb200_bc_elembuffer(int level, int nBcElems, const int32_t *bc_elemBuffer) {
  GET_LEVEL(L, level);
  if (nBcElems < 0 || (nBcElems > 0 && !bc_elemBuffer)) return setError(MUSB200_ERR_ARG, "bad BC element buffer");
  MUSB_TRY(L->bcElems.upload(bc_elemBuffer, (size_t)nBcElems, g.stream));
  L->bcElemsHost.assign(bc_elemBuffer, bc_elemBuffer + nBcElems);
  L->bcSlotNeeded.assign((size_t)nBcElems, 0);
  MUSB_TRY(L->bcBuffer.alloc((size_t)std::max(1, nBcElems) * L->QQ));
  MUSB_CUDA(cudaMemsetAsync(L->bcBuffer.p, 0, L->bcBuffer.n * sizeof(double), g.stream));
  return 0;
}

int musb200_bc_register(int level, int bc_id, int bc_kind, int nLinks, const int32_t *links,
                        const int32_t *outPos, const int32_t *posInBuffer, const int32_t *iDir) {
  GET_LEVEL(L, level);
  if (bc_kind < MUSB200_BC_WALL || bc_kind > MUSB200_BC_PRESSURE_EXPOL)
    return setError(MUSB200_ERR_UNSUPPORTED,
                    "boundary kind " + std::to_string(bc_kind) + " is outside the B200 hot path");
  if (nLinks < 0) return setError(MUSB200_ERR_ARG, "nLinks < 0");
  auto b = std::make_unique<BcData>();
  b->id = bc_id; b->kind = bc_kind; b->nLinks = nLinks;
  if (bc_kind != MUSB200_BC_WALL && nLinks > 0) {
    if (!links || !outPos || !posInBuffer || !iDir) return setError(MUSB200_ERR_ARG, "null link list");
    if (L->bcElems.n == 0) return setError(MUSB200_ERR_STATE, "musb200_bc_elembuffer must come first");
    MUSB_TRY(b->links.upload(links, nLinks, g.stream));
    MUSB_TRY(b->outPos.upload(outPos, nLinks, g.stream));
    MUSB_TRY(b->posInBuffer.upload(posInBuffer, nLinks, g.stream));
    MUSB_TRY(b->iDir.upload(iDir, nLinks, g.stream));
    for (int l = 0; l < nLinks; ++l) {
      const int pib = posInBuffer[l], op = (outPos[l] - 1) / L->QQ + 1;
      if (pib < 1 || pib > (int)L->bcSlotNeeded.size() || op < 1 || op > (int)L->bcSlotNeeded.size())
        return setError(MUSB200_ERR_ARG, "boundary link refers to a slot outside bc_elemBuffer");
      L->bcSlotNeeded[pib - 1] = 1;
      L->bcSlotNeeded[op - 1] = 1;
    }
    if (bc_kind == MUSB200_BC_VELOCITY_BOUNCEBACK) {
      // group the links by boundary element; the fused kernel needs every link to read and
      // write slots of its own element and every element to form one contiguous group
      std::vector<int32_t> gs, ge;
      std::vector<char> seen(L->bcSlotNeeded.size(), 0);
      bool ok = true;
      for (int l = 0; l < nLinks && ok; ++l) {
        const int pib = posInBuffer[l];
        const int e = L->bcElemsHost[pib - 1] - 1;
        if ((outPos[l] - 1) / L->QQ + 1 != pib || (links[l] - 1) / L->QQ != e) ok = false;
        if (l == 0 || posInBuffer[l - 1] != pib) {
          if (seen[pib - 1]) ok = false;
          seen[pib - 1] = 1;
          gs.push_back(l);
          ge.push_back(e);
          b->slots.push_back(pib);
        }
      }
      gs.push_back(nLinks);
      b->fusable = ok;
      if (ok) {
        b->nGroups = (int)ge.size();
        MUSB_TRY(b->groupStart.upload(gs.data(), gs.size(), g.stream));
        MUSB_TRY(b->groupElem.upload(ge.data(), ge.size(), g.stream));
      }
    }
    L->bcFused = -1;
    std::vector<int32_t> needed;
    for (size_t i = 0; i < L->bcSlotNeeded.size(); ++i)
      if (L->bcSlotNeeded[i]) needed.push_back((int32_t)i + 1);
    MUSB_TRY(L->bcNeeded.upload(needed.data(), needed.size(), g.stream));
    MUSB_CUDA(cudaStreamSynchronize(g.stream));
  }
  for (auto &o : L->bcs)
    if (o->id == bc_id) { o = std::move(b); return 0; }
  L->bcs.push_back(std::move(b));
  return 0;
}

int musb200_bc_register_elems(int level, int bc_id, int nElems, const int32_t *elemPos,
                              const int32_t *posInBcElemBuf, const int32_t *normalInd, int nNeighs,
                              const int32_t *neighPos, const int32_t *iElemOfLink) {
  GET_LEVEL(L, level);
  for (auto &b : L->bcs) {
    if (b->id != bc_id) continue;
    if (!isPressureBc(b->kind)) return setError(MUSB200_ERR_ARG, "boundary kind reads no neighbours");
    const int need = b->kind == MUSB200_BC_PRESSURE_EXPOL ? 2 : 1;   // me%nNeighs, mus_bc_header_module.fpp
    if (nElems <= 0 || nNeighs < need || !elemPos || !posInBcElemBuf || !normalInd || !neighPos || !iElemOfLink)
      return setError(MUSB200_ERR_ARG, "bad boundary element lists");
    for (int i = 0; i < nElems; ++i) {
      if (elemPos[i] < 1 || elemPos[i] > L->nElems || posInBcElemBuf[i] < 1 ||
          posInBcElemBuf[i] > (int)L->bcSlotNeeded.size() || normalInd[i] < 1 || normalInd[i] > L->QQ)
        return setError(MUSB200_ERR_ARG, "boundary element list entry out of range");
      for (int k = 0; k < nNeighs; ++k)
        if (neighPos[(size_t)i * nNeighs + k] < 1 || neighPos[(size_t)i * nNeighs + k] > L->nElems)
          return setError(MUSB200_ERR_ARG, "boundary neighbour position out of range");
    }
    // keep the first `need` neighbours of every element (posInState(1:need, iElem))
    std::vector<int32_t> np((size_t)nElems * need);
    for (int i = 0; i < nElems; ++i)
      for (int k = 0; k < need; ++k) np[(size_t)i * need + k] = neighPos[(size_t)i * nNeighs + k];
    b->nElems = nElems; b->nNeighs = need;
    MUSB_TRY(b->elemPos.upload(elemPos, nElems, g.stream));
    MUSB_TRY(b->posInBcElemBuf.upload(posInBcElemBuf, nElems, g.stream));
    MUSB_TRY(b->normalInd.upload(normalInd, nElems, g.stream));
    MUSB_TRY(b->neighPos.upload(np.data(), np.size(), g.stream));
    MUSB_TRY(b->iElemOfLink.upload(iElemOfLink, b->nLinks, g.stream));
    MUSB_TRY(b->neighBuf.alloc((size_t)need * nElems * L->QQ));
    MUSB_CUDA(cudaStreamSynchronize(g.stream));
    if (b->kind == MUSB200_BC_PRESSURE_EXPOL) L->auxForBc = true;   // reads auxField of the previous step
    return 0;
  }
  return setError(MUSB200_ERR_ARG, "unknown boundary id " + std::to_string(bc_id));
}

int musb200_bc_set_values(int level, int bc_id, int nVals, const double *vals) {
  GET_LEVEL(L, level);
  for (auto &b : L->bcs) {
    if (b->id != bc_id) continue;
    if (nVals < 0 || (nVals > 0 && !vals)) return setError(MUSB200_ERR_ARG, "bad values");
    if (!b->copied) MUSB_CUDA(cudaEventCreateWithFlags(&b->copied, cudaEventDisableTiming));
    // valsNext was the active buffer until the last swap: its last readers are the boundary
    // kernels of an earlier step
    MUSB_CUDA(cudaStreamWaitEvent(g.copyStream, g.evBcDone, 0));
    if (b->valsNext.n != (size_t)nVals) {
      MUSB_CUDA(cudaStreamSynchronize(g.copyStream));   // a pending copy into the old allocation
      MUSB_TRY(b->valsNext.alloc((size_t)nVals));
    }
    if (nVals)
      MUSB_CUDA(cudaMemcpyAsync(b->valsNext.p, vals, (size_t)nVals * sizeof(double), cudaMemcpyHostToDevice,
                                g.copyStream));
    MUSB_CUDA(cudaEventRecord(b->copied, g.copyStream));
    b->pending = true;
    return 0;
  }
  return setError(MUSB200_ERR_ARG, "unknown boundary id " + std::to_string(bc_id));
}

// ---------------------------------------------------------------------------
int musb200_comm_register(int level, int buf_kind, int dir, int nProcs, const int32_t *proc,
                          const int32_t *nVals, const int32_t *pos) {
  GET_LEVEL(L, level);
  if (buf_kind < 0 || buf_kind > 2 || dir < 0 || dir > 1 || nProcs < 0)
    return setError(MUSB200_ERR_ARG, "bad buffer kind / direction");
  CommBuf &c = (dir == MUSB200_DIR_SEND) ? L->send[buf_kind] : L->recv[buf_kind];
  c.proc.clear(); c.nVals.clear(); c.offset.clear(); c.total = 0;
  for (int i = 0; i < nProcs; ++i) {
    if (proc[i] < 0 || proc[i] >= g.nranks || proc[i] == g.rank || nVals[i] < 0)
      return setError(MUSB200_ERR_ARG, "bad peer rank in comm buffer");
    c.proc.push_back(proc[i]); c.nVals.push_back(nVals[i]); c.offset.push_back(c.total);
    c.total += nVals[i];
  }
  MUSB_TRY(c.pos.upload(pos, (size_t)c.total, g.stream));
  MUSB_TRY(c.buf.alloc((size_t)std::max(1, c.total)));
  c.auxNVals.clear(); c.auxOffset.clear(); c.auxTotal = 0; c.auxPosHost.clear();
  // halo elements and ghostFromFiner elements travel with their auxField entries
  // (auxField%sendBuffer / sendBufferFromFiner, mus_auxField_module.f90:377-444)
  if (buf_kind == MUSB200_BUF_HALO || buf_kind == MUSB200_BUF_FROMFINER) {
    std::vector<int32_t> apos;
    std::vector<char> seen;
    for (int i = 0; i < nProcs; ++i) {
      seen.assign((size_t)L->nSize, 0);
      const int before = (int)apos.size();
      for (int j = c.offset[i]; j < c.offset[i] + c.nVals[i]; ++j) {
        const int e = (pos[j] - 1) / L->QQ;
        if (e < 0 || e >= L->nSize) return setError(MUSB200_ERR_ARG, "comm position outside the level");
        if (seen[e]) continue;
        seen[e] = 1;
        for (int k = 0; k < 4; ++k) apos.push_back(e * 4 + k + 1);
      }
      c.auxOffset.push_back(before);
      c.auxNVals.push_back((int)apos.size() - before);
    }
    c.auxTotal = (int)apos.size();
    c.auxPosHost = apos;
    MUSB_TRY(c.auxPos.upload(apos.data(), apos.size(), g.stream));
    MUSB_TRY(c.buf.alloc((size_t)std::max(1, c.total + c.auxTotal)));   // state links | auxField entries
    MUSB_CUDA(cudaStreamSynchronize(g.stream));
  }
  return 0;
}

// ---------------------------------------------------------------------------
// peer-memory halo exchange: export / connect
namespace {
struct P2PBlob {                 // MUSB200_P2P_BLOB bytes, plain data
  cudaIpcMemHandle_t state[2];
  cudaIpcMemHandle_t arrived;    // arrived[nranks] | ready[nranks]
  cudaIpcMemHandle_t aux;
  long long S;
  int rank, QQ, nSize, device;
  char pad[MUSB200_P2P_BLOB - 4 * (int)sizeof(cudaIpcMemHandle_t) - (int)sizeof(long long) - 4 * (int)sizeof(int)];
};
static_assert(sizeof(P2PBlob) == MUSB200_P2P_BLOB, "blob layout");
}  // namespace

int musb200_p2p_export(int level, void *blob) {
  GET_LEVEL(L, level);
  if (!blob) return setError(MUSB200_ERR_ARG, "null argument");
  PeerLink &P = L->p2p;
  if (P.arrived.n == 0) {
    MUSB_TRY(P.arrived.alloc((size_t)2 * std::max(g.nranks, 1)));
    MUSB_TRY(P.ticket.alloc(1));
    MUSB_TRY(P.exch.alloc(3));
    MUSB_CUDA(cudaMemsetAsync(P.arrived.p, 0, P.arrived.n * sizeof(unsigned long long), g.stream));
    MUSB_CUDA(cudaMemsetAsync(P.ticket.p, 0, sizeof(unsigned int), g.stream));
    MUSB_CUDA(cudaMemsetAsync(P.exch.p, 0, 3 * sizeof(unsigned long long), g.stream));
    MUSB_CUDA(cudaStreamSynchronize(g.stream));
  }
  P2PBlob b;
  std::memset(&b, 0, sizeof(b));
  MUSB_CUDA(cudaIpcGetMemHandle(&b.state[0], L->state[0].p));
  MUSB_CUDA(cudaIpcGetMemHandle(&b.state[1], L->state[1].p));
  MUSB_CUDA(cudaIpcGetMemHandle(&b.arrived, P.arrived.p));
  MUSB_CUDA(cudaIpcGetMemHandle(&b.aux, L->aux.p));
  b.S = L->S; b.rank = g.rank; b.QQ = L->QQ; b.nSize = L->nSize; b.device = g.device;
  std::memcpy(blob, &b, sizeof(b));
  return 0;
}

int musb200_p2p_connect(int level, int nProcs, const int32_t *proc, const void *blobs,
                        const int32_t *nVals, const int32_t *remotePos) {
  GET_LEVEL(L, level);
  PeerLink &P = L->p2p;
  CommBuf &s = L->send[MUSB200_BUF_HALO], &r = L->recv[MUSB200_BUF_HALO];
  if (P.arrived.n == 0) return setError(MUSB200_ERR_STATE, "musb200_p2p_export must come first");
  if (nProcs != (int)s.proc.size() || nProcs > kMaxPeers || (int)r.proc.size() > kMaxPeers)
    return setError(MUSB200_ERR_ARG, "peer list does not match the halo send buffer");
  if (nProcs > 0 && (!proc || !blobs || !nVals || !remotePos)) return setError(MUSB200_ERR_ARG, "null argument");
  std::vector<uint8_t> peerOf((size_t)s.total);
  const int QQ = L->QQ;
  // auxField entries of the communicated elements on the receiving side: the elements of the
  // receiver's list in order of first appearance, exactly how both sides build auxField%recvBuffer /
  // sendBuffer from their own lists (musb200_comm_register), so entry j here pairs with entry j of
  // this rank's s.auxPosHost
  std::vector<int32_t> auxRemote;
  std::vector<uint8_t> auxPeer;
  for (int k = 0; k < nProcs; ++k) {
    if (proc[k] != s.proc[k] || nVals[k] != s.nVals[k])
      return setError(MUSB200_ERR_ARG, "peer order / message length differs from musb200_comm_register");
    P2PBlob b;
    std::memcpy(&b, static_cast<const char *>(blobs) + (size_t)k * MUSB200_P2P_BLOB, sizeof(b));
    if (b.rank != proc[k] || b.QQ != L->QQ) return setError(MUSB200_ERR_ARG, "blob belongs to another rank / stencil");
    int can = 0;
    MUSB_CUDA(cudaDeviceCanAccessPeer(&can, g.device, b.device));
    if (!can) return setError(MUSB200_ERR_CUDA, "no peer access between devices " + std::to_string(g.device) +
                                                    " and " + std::to_string(b.device));
    void *p0 = nullptr, *p1 = nullptr, *pf = nullptr, *pa = nullptr;
    MUSB_CUDA(cudaIpcOpenMemHandle(&p0, b.state[0], cudaIpcMemLazyEnablePeerAccess));
    P.opened.push_back(p0);
    MUSB_CUDA(cudaIpcOpenMemHandle(&p1, b.state[1], cudaIpcMemLazyEnablePeerAccess));
    P.opened.push_back(p1);
    MUSB_CUDA(cudaIpcOpenMemHandle(&pf, b.arrived, cudaIpcMemLazyEnablePeerAccess));
    P.opened.push_back(pf);
    MUSB_CUDA(cudaIpcOpenMemHandle(&pa, b.aux, cudaIpcMemLazyEnablePeerAccess));
    P.opened.push_back(pa);
    P.remoteState[k][0] = static_cast<double *>(p0);
    P.remoteState[k][1] = static_cast<double *>(p1);
    P.remoteArrived[k] = static_cast<unsigned long long *>(pf);
    P.remoteAux[k] = static_cast<double *>(pa);
    P.remoteS[k] = b.S;
    std::vector<char> seen((size_t)b.nSize, 0);
    for (int i = 0; i < nVals[k]; ++i) {
      const int rp = remotePos[s.offset[k] + i];
      if (rp < 1 || rp > b.nSize * b.QQ) return setError(MUSB200_ERR_ARG, "remote position out of range");
      peerOf[(size_t)s.offset[k] + i] = (uint8_t)k;
      const int e = (rp - 1) / QQ;
      if (!seen[e]) {
        seen[e] = 1;
        for (int c = 0; c < 4; ++c) { auxRemote.push_back(e * 4 + c + 1); auxPeer.push_back((uint8_t)k); }
      }
    }
  }
  if (auxRemote.size() != s.auxPosHost.size())
    return setError(MUSB200_ERR_ARG, "the receivers' element lists do not pair with this rank's send elements");
  P.sendRank.assign(s.proc.begin(), s.proc.end());
  P.recvRank.assign(r.proc.begin(), r.proc.end());
  // The lists come elem-major / direction-minor (the reference's AOS order).  The receiver's
  // rows are SoA and its halo block is sorted by treeID, so ordering the entries by (peer,
  // remote direction, remote element) makes consecutive threads store consecutive addresses:
  // full 128-byte NVLink writes instead of scattered 8-byte ones.
  std::vector<int32_t> srcHost((size_t)s.total);
  MUSB_CUDA(cudaMemcpyAsync(srcHost.data(), s.pos.p, (size_t)s.total * sizeof(int32_t), cudaMemcpyDeviceToHost,
                            g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  std::vector<int32_t> order((size_t)s.total);
  for (int i = 0; i < s.total; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
    if (peerOf[x] != peerOf[y]) return peerOf[x] < peerOf[y];
    const int dx = (remotePos[x] - 1) % QQ, dy = (remotePos[y] - 1) % QQ;
    if (dx != dy) return dx < dy;
    return remotePos[x] < remotePos[y];
  });
  std::vector<int32_t> srcSorted((size_t)s.total), dstSorted((size_t)s.total);
  std::vector<uint8_t> peerSorted((size_t)s.total);
  for (int i = 0; i < s.total; ++i) {
    srcSorted[i] = srcHost[order[i]];
    dstSorted[i] = remotePos[order[i]];
    peerSorted[i] = peerOf[order[i]];
  }
  MUSB_TRY(P.srcPos.upload(srcSorted.data(), srcSorted.size(), g.stream));
  MUSB_TRY(P.dstPos.upload(dstSorted.data(), dstSorted.size(), g.stream));
  MUSB_TRY(P.peerOf.upload(peerSorted.data(), peerSorted.size(), g.stream));
  {
    // auxField entries, ordered the same way: (peer, component, remote element)
    const int nA = (int)auxRemote.size();
    std::vector<int32_t> ao((size_t)nA);
    for (int i = 0; i < nA; ++i) ao[i] = i;
    std::stable_sort(ao.begin(), ao.end(), [&](int32_t x, int32_t y) {
      if (auxPeer[x] != auxPeer[y]) return auxPeer[x] < auxPeer[y];
      const int cx_ = (auxRemote[x] - 1) & 3, cy_ = (auxRemote[y] - 1) & 3;
      if (cx_ != cy_) return cx_ < cy_;
      return auxRemote[x] < auxRemote[y];
    });
    std::vector<int32_t> as((size_t)nA), ad((size_t)nA);
    std::vector<uint8_t> ap((size_t)nA);
    for (int i = 0; i < nA; ++i) { as[i] = s.auxPosHost[ao[i]]; ad[i] = auxRemote[ao[i]]; ap[i] = auxPeer[ao[i]]; }
    MUSB_TRY(P.auxSrcPos.upload(as.data(), as.size(), g.stream));
    MUSB_TRY(P.auxDstPos.upload(ad.data(), ad.size(), g.stream));
    MUSB_TRY(P.auxPeerOf.upload(ap.data(), ap.size(), g.stream));
    P.nAux = nA;
  }
  {
    // the same entries grouped by owning element, for the push fused into the sweep
    std::vector<int32_t> byElem((size_t)s.total);
    for (int i = 0; i < s.total; ++i) byElem[i] = i;
    std::stable_sort(byElem.begin(), byElem.end(), [&](int32_t x, int32_t y) {
      return (srcHost[x] - 1) / QQ < (srcHost[y] - 1) / QQ;
    });
    const size_t nWords = ((size_t)L->S + 31) / 32;
    std::vector<uint32_t> mask(nWords, 0u), prefix(nWords, 0u);
    std::vector<int32_t> start, dst((size_t)s.total);
    std::vector<uint8_t> eq((size_t)s.total), ep((size_t)s.total);
    int last = -1;
    for (int j = 0; j < s.total; ++j) {
      const int i = byElem[j];
      const int e = (srcHost[i] - 1) / QQ;
      if (e >= L->nSolve) return setError(MUSB200_ERR_ARG, "halo send buffer refers to an element that is not solved here");
      if (e != last) { start.push_back(j); mask[e >> 5] |= 1u << (e & 31); last = e; }
      eq[j] = (uint8_t)((srcHost[i] - 1) % QQ);
      ep[j] = peerOf[i];
      dst[j] = remotePos[i];
    }
    start.push_back(s.total);
    uint32_t run = 0;
    for (size_t w = 0; w < nWords; ++w) { prefix[w] = run; run += (uint32_t)__builtin_popcount(mask[w]); }
    MUSB_TRY(P.pushMask.upload(mask.data(), mask.size(), g.stream));
    MUSB_TRY(P.pushPrefix.upload(prefix.data(), prefix.size(), g.stream));
    MUSB_TRY(P.pushStart.upload(start.data(), start.size(), g.stream));
    MUSB_TRY(P.pushDst.upload(dst.data(), dst.size(), g.stream));
    MUSB_TRY(P.pushQ.upload(eq.data(), eq.size(), g.stream));
    MUSB_TRY(P.pushPeer.upload(ep.data(), ep.size(), g.stream));
  }
  {
    // which CTAs of the sweep pull from a halo row: they do the exchange's wait (p2p.cu).  The
    // protocol without a "ready to receive" handshake needs every rank this rank sends to to
    // send to it as well.
    const int block = sweepBlockSize(QQ);
    const size_t nCtaWords = ((size_t)divUp(std::max(L->nSolve, 1), block) + 31) / 32;
    MUSB_TRY(P.ctaMask.alloc(nCtaWords));
    MUSB_CUDA(cudaMemsetAsync(P.ctaMask.p, 0, nCtaWords * sizeof(uint32_t), g.stream));
    MUSB_TRY(launchHaloCtaMask(QQ, L->nbr.p, L->S, L->nSolve, L->nFluid + L->nGFC + L->nGFF, block,
                               P.ctaMask.p, g.stream));
    {
      std::vector<uint32_t> hm(nCtaWords);
      MUSB_CUDA(cudaMemcpyAsync(hm.data(), P.ctaMask.p, nCtaWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, g.stream));
      MUSB_CUDA(cudaStreamSynchronize(g.stream));
      std::vector<int32_t> list;
      P.nCtas = divUp(std::max(L->nSolve, 1), block);
      for (int c = 0; c < P.nCtas; ++c)
        if ((hm[c >> 5] >> (c & 31)) & 1u) list.push_back(c);
      P.nHaloCtas = (int)list.size();
      if (list.empty()) list.push_back(0);
      MUSB_TRY(P.haloCtas.upload(list.data(), list.size(), g.stream));
      if (!P.evSwept) MUSB_CUDA(cudaEventCreateWithFlags(&P.evSwept, cudaEventDisableTiming));
      for (int b = 0; b < 2; ++b)
        if (!P.evPushed[b]) MUSB_CUDA(cudaEventCreateWithFlags(&P.evPushed[b], cudaEventDisableTiming));
    }
    std::vector<int> a1(P.sendRank), a2(P.recvRank);
    std::sort(a1.begin(), a1.end());
    std::sort(a2.begin(), a2.end());
    if (a1 != a2)
      return setError(MUSB200_ERR_UNSUPPORTED, "peer-memory halo exchange needs symmetric peers (every rank this "
                                               "rank sends to must send to it): stay on the NCCL path");
    P.sweepWait = true;
  }
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  P.on = true;
  return 0;
}

int musb200_p2p_enable(int level, int flag) {
  GET_LEVEL(L, level);
  if (flag && L->p2p.dstPos.n == 0 && L->send[MUSB200_BUF_HALO].total > 0)
    return setError(MUSB200_ERR_STATE, "musb200_p2p_connect must come first");
  L->p2p.on = flag != 0;
  return 0;
}

// ---------------------------------------------------------------------------
int musb200_intp_register(int tgtLevel, int direction, int order, int nTargets,
                          const int32_t *targetList, const int32_t *srcOffset, const int32_t *srcPos,
                          const double *weights, const int32_t *posInMat, int nMatrices,
                          const int32_t *matOffset, const double *matrices, const double *childCoord) {
  GET_LEVEL(L, tgtLevel);
  if (nTargets < 0 || order < 0 || order > 2) return setError(MUSB200_ERR_ARG, "bad interpolation set");
  IntpSet *set;
  if (direction == MUSB200_INTP_FROMFINER) {
    set = &L->fromFiner;
  } else if (direction == MUSB200_INTP_FROMCOARSER) {
    if ((int)L->fromCoarser.size() <= order) L->fromCoarser.resize(order + 1);
    set = &L->fromCoarser[order];
  } else {
    return setError(MUSB200_ERR_ARG, "bad interpolation direction");
  }
  return registerIntp(*set, order, nTargets, targetList, srcOffset, srcPos, weights, posInMat,
                      nMatrices, matOffset, matrices, childCoord, g.stream);
}

// ---------------------------------------------------------------------------
int musb200_set_aux_every_step(int flag) {
  ++g.epoch;
  g.auxEveryStep = (flag == 2) ? 2 : (flag ? 1 : 0);
  return 0;
}

// Long single-rank runs without per-stage timing replay a CUDA graph of two coarse cycles: the
// stepping loop of a small or multi-level mesh is launch-bound (64^3: 19 us of kernel per 22 us
// step; a two-level cycle is 7 launches).  Several ranks stay on direct launches: the peer-memory
// exchange carries a running exchange number as a kernel argument.
// identifies what a captured graph steps: the bound slot, or the list of coupled slots
static int stepSlotKey() {
  if (gNStepSlots == 0) return g.slot;
  int key = 1000;
  for (int k = 0; k < gNStepSlots; ++k) key = key * 8 + gStepSlots[k] + 1;
  return key;
}

static int stepGraphed(int minLevel, int maxLevel, int nPairs) {
  // the captured kernel arguments hold the now/next buffers of the capture: a replay must start
  // from the same parity of the coarsest level (finer levels toggle twice per cycle; an explicit
  // musb200_set_now_next bumps the epoch)
  int parity = 0, nth = 0;
  forEachStepSlot([&] { parity |= findLevel(minLevel)->nNow << nth++; return 0; });
  if (!g.graphExec || g.graphEpoch != g.epoch || g.graphMin != minLevel || g.graphMax != maxLevel ||
      g.graphSlot != stepSlotKey() || g.graphParity != parity) {
    dropGraph();
    const long long before = g.launches;
    const int auxSave = g.auxEveryStep;
    cudaGraph_t graph = nullptr;
    MUSB_CUDA(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    g.capturing = true;
    // single level on several ranks: a replay follows a replay whose last push nobody has waited
    // for, so the FIRST sweep of the pair must carry the wait as well (it passes at once when
    // nothing is in flight)
    if (g.nranks > 1 && maxLevel == minLevel)
      forEachStepSlot([&] {
        Level *L0 = findLevel(minLevel);
        if (L0->p2p.on && (L0->send[MUSB200_BUF_HALO].total > 0 || L0->recv[MUSB200_BUF_HALO].total > 0))
          L0->p2p.pendingWait = true;
        return 0;
      });
    for (int it = 0; it < 2 && rc == 0; ++it) rc = levelStep(minLevel, minLevel, maxLevel, false);
    g.capturing = false;
    cudaError_t ce = cudaStreamEndCapture(g.stream, &graph);
    g.auxEveryStep = auxSave;
    g.graphLaunches = g.launches - before;
    g.launches = before;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    MUSB_CUDA(ce);
    ce = cudaGraphInstantiate(&g.graphExec, graph, 0);
    cudaGraphDestroy(graph);
    MUSB_CUDA(ce);
    // capturing ran the host side of two cycles: every level toggled now/next an even number of
    // times, so the indices are where they were and nothing was executed yet
    g.graphEpoch = g.epoch; g.graphMin = minLevel; g.graphMax = maxLevel; g.graphSlot = stepSlotKey();
    g.graphParity = parity;
  }
  for (int p = 0; p < nPairs; ++p) {
    MUSB_CUDA(cudaGraphLaunch(g.graphExec, g.stream));
    g.launches += g.graphLaunches;
  }
  if (nPairs > 0 && g.nranks > 1)
    forEachStepSlot([&] {
      for (int l = minLevel; l <= maxLevel; ++l) {
        Level *Lp = findLevel(l);
        if (Lp->p2p.on && (Lp->send[MUSB200_BUF_HALO].total > 0 || Lp->recv[MUSB200_BUF_HALO].total > 0))
          Lp->p2p.pendingWait = (maxLevel == minLevel);   // multi-level pushes are followed by their wait
      }
      return 0;
    });
  return 0;   // two cycles leave every level's now/next parity unchanged
}

static int stepImpl(int minLevel, int maxLevel, int nCoarseCycles);

int musb200_step(int minLevel, int maxLevel, int nCoarseCycles) {
  MUSB_TRY(needReady());
  gNStepSlots = 0;
  return stepImpl(minLevel, maxLevel, nCoarseCycles);
}

int musb200_step_schemes(int nSlots, const int *slots, int minLevel, int maxLevel, int nCoarseCycles) {
  MUSB_TRY(needReady());
  if (nSlots < 1 || nSlots > Context::kSlots || !slots) return setError(MUSB200_ERR_ARG, "bad scheme slot list");
  for (int k = 0; k < nSlots; ++k) {
    if (slots[k] < 0 || slots[k] >= Context::kSlots) return setError(MUSB200_ERR_ARG, "scheme slot out of range");
    for (int j = 0; j < k; ++j)
      if (slots[j] == slots[k]) return setError(MUSB200_ERR_ARG, "scheme slot listed twice");
    gStepSlots[k] = slots[k];
  }
  gNStepSlots = nSlots;
  const int keep = g.slot;
  g.slot = slots[0];
  const int rc = stepImpl(minLevel, maxLevel, nCoarseCycles);
  g.slot = keep;
  gNStepSlots = 0;
  return rc;
}

static int stepImpl(int minLevel, int maxLevel, int nCoarseCycles) {
  if (maxLevel < minLevel || nCoarseCycles < 0) return setError(MUSB200_ERR_ARG, "bad level range / cycles");
  int it = 0;
  // several ranks: capturable when every exchange of the range goes through peer memory (its
  // kernels keep the exchange number on the device; NCCL calls stay outside graphs)
  bool capturable = true;
  MUSB_TRY(forEachStepSlot([&] {
    for (int l = minLevel; l <= maxLevel; ++l)
      if (!findLevel(l))
        return setError(MUSB200_ERR_ARG, "level " + std::to_string(l) + " was not created (scheme slot " +
                                             std::to_string(g.slot) + ")");
    return 0;
  }));
  if (g.nranks > 1) {
    if (g.overlap) capturable = false;
    forEachStepSlot([&] {
      for (int l = minLevel; l <= maxLevel; ++l) {
        Level *Lc = findLevel(l);
        for (int k = 0; k < 3; ++k) {
          const bool any = Lc->send[k].total > 0 || Lc->recv[k].total > 0;
          if (any && !(k == MUSB200_BUF_HALO && Lc->p2p.on)) capturable = false;
        }
      }
      return 0;
    });
  }
  if (g.useGraphs && capturable && !g.profiling && nCoarseCycles >= 8) {
    // all but the last cycles (the last one materialises auxField) in pairs through the graph
    const int nPairs = (nCoarseCycles - 1) / 2;
    MUSB_TRY(forEachStepSlot([&] {
      for (int l = minLevel; l <= maxLevel; ++l) MUSB_TRY(applyPendingBc(*findLevel(l)));   // cross-stream waits stay outside the capture
      return 0;
    }));
    MUSB_TRY(stepGraphed(minLevel, maxLevel, nPairs));
    MUSB_TRY(forEachStepSlot([&] {
      for (int l = minLevel; l <= maxLevel; ++l) {
        Level *Lg = findLevel(l);
        if (Lg->bcElems.n) MUSB_CUDA(cudaEventRecord(g.evBcDone, g.stream));
        Lg->auxValid = (maxLevel > minLevel) || g.auxEveryStep == 1 || Lg->auxForBc;
      }
      return 0;
    }));
    it = 2 * nPairs;
  }
  for (; it < nCoarseCycles; ++it)
    MUSB_TRY(levelStep(minLevel, minLevel, maxLevel, it == nCoarseCycles - 1));
  // MPI_Waitall of the last exchange: whatever follows this call sees complete halo rows
  return forEachStepSlot([&] {
    for (int l = minLevel; l <= maxLevel; ++l) {
      Level *Lw = findLevel(l);
      MUSB_TRY(ensureArrived(*Lw));
      for (int b = 0; b < 2; ++b)
        if (g.commBusy && Lw->p2p.pushedOnce[b]) MUSB_CUDA(cudaStreamWaitEvent(g.stream, Lw->p2p.evPushed[b], 0));
    }
    g.commBusy = false;
    return 0;
  });
}

int musb200_fill_helper_elements(int minLevel, int maxLevel) {
  MUSB_TRY(needReady());
  if (maxLevel < minLevel) return setError(MUSB200_ERR_ARG, "bad level range");
  ++g.epoch;
  for (int l = minLevel; l <= maxLevel; ++l) {
    Level *L = findLevel(l);
    if (!L) return setError(MUSB200_ERR_ARG, "level " + std::to_string(l) + " was not created");
    if (!L->relaxSet) return setError(MUSB200_ERR_STATE, "musb200_set_relaxation missing");
    if (L->kind == MUSB200_KIND_PASSIVE_SCALAR) continue;   // its auxField is rebuilt by the first step
    // mus_initAuxFieldFluidAndExchange: auxField of the fluid elements from their own PDFs
    SweepArgs a{};
    a.in = L->state[L->nNext].p; a.nbr = nullptr; a.aux = L->aux.p; a.S = L->S;
    a.first = 0; a.count = L->nFluid;
    MUSB_TRY(launchAuxOnly(L->QQ, L->kind, a, g.stream));
    ++g.launches;
    L->auxValid = true;
  }
  MUSB_TRY(fillFineToCoarse(minLevel, minLevel, maxLevel));
  MUSB_TRY(fillCoarseToFine(minLevel, minLevel, maxLevel));
  for (int l = minLevel; l <= maxLevel; ++l) MUSB_TRY(ensureArrived(*findLevel(l)));
  return 0;
}

int musb200_set_graphs(int flag) {
  g.useGraphs = flag ? 1 : 0;
  if (!flag) dropGraph();
  return 0;
}

int musb200_set_fused_push(int flag) {
  g.fusedPush = flag ? 1 : 0;
  return 0;
}

int musb200_set_fused_bc(int flag) {
  ++g.epoch;
  g.noFusedBc = flag ? 0 : 1;
  return 0;
}

int musb200_set_exchange_timeout(double seconds) {
  if (!(seconds >= 0.0)) return setError(MUSB200_ERR_ARG, "timeout must be >= 0 (0 = wait for ever)");
  ++g.epoch;
  g.timeoutNs = (unsigned long long)(seconds * 1.0e9);
  return 0;
}

int musb200_set_intp_tiled(int flag) {
  ++g.epoch;
  g_intpTargetMajor = flag ? 0 : 1;
  return 0;
}

int musb200_set_sweep_wait(int flag) {
  ++g.epoch;
  g.sweepWait = flag ? 1 : 0;
  return 0;
}

int musb200_set_overlap(int flag) {
  g.overlap = flag ? 1 : 0;
  return 0;
}

int musb200_synchronize(void) {
  MUSB_TRY(needReady());
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.commStream));
  return checkAsyncErrors();
}

int musb200_reduce(int level, double *total_mass, double *max_vel, int *any_nan) {
  GET_LEVEL_RO(L, level);
  double *out = g.red.p + 3 * 592;
  MUSB_TRY(launchReduce(L->QQ, L->state[L->nNext].p, L->S, L->nFluid, g.red.p, out, g.stream));
  g.launches += 2;
  if (g.nranks > 1) {
    // check_density: mpi_reduce of the total density (mus_tools_module.f90:298)
    MUSB_NCCL(g.nccl->AllReduce(out, out + 4, 1, ncclDouble, ncclSum, g.comm, g.stream));
    MUSB_NCCL(g.nccl->AllReduce(out + 1, out + 5, 1, ncclDouble, ncclMax, g.comm, g.stream));
    MUSB_NCCL(g.nccl->AllReduce(out + 2, out + 6, 1, ncclDouble, ncclSum, g.comm, g.stream));
    out += 4;
  }
  double h[3];
  MUSB_CUDA(cudaMemcpyAsync(h, out, 3 * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  MUSB_TRY(checkAsyncErrors());
  if (total_mass) *total_mass = h[0];
  if (max_vel) *max_vel = sqrt(h[1]);
  if (any_nan) *any_nan = (h[2] > 0.0 || h[0] != h[0]) ? 1 : 0;
  return 0;
}

// ---------------------------------------------------------------------------
int musb200_compute_host(int relax_id, int kind_id, int QQ, const double *inState, double *outState,
                         double *auxField, const int32_t *neigh, int nElems, int nSolve,
                         const double *omega, double lambda, double omega_bulk) {
  MUSB_TRY(needReady());
  if (!inState || !outState || !neigh || !omega || nSolve > nElems)
    return setError(MUSB200_ERR_ARG, "bad arguments");
  const int tmpLevel = -4711;
  MUSB_TRY(musb200_level_create(tmpLevel, QQ, QQ, 4, nElems, nSolve, 0, 0, nElems - nSolve, neigh,
                                nullptr, nullptr));
  int rc = 0;
  do {
    if ((rc = musb200_state_upload(tmpLevel, 1, inState))) break;
    if ((rc = musb200_set_now_next(tmpLevel, 1, 2))) break;
    if ((rc = musb200_set_relaxation(tmpLevel, relax_id, kind_id, omega, 0.0, lambda, omega_bulk))) break;
    Level *L = findLevel(tmpLevel);
    if ((rc = sweep(*L, true))) break;
    if ((rc = musb200_state_download(tmpLevel, 2, outState))) break;
    if (auxField) rc = musb200_aux_download(tmpLevel, auxField);
  } while (0);
  musb200_level_destroy(tmpLevel);
  return rc;
}

// ---------------------------------------------------------------------------
int musb200_set_profiling(int flag) {
  g.profiling = flag ? 1 : 0;
  return 0;
}

int musb200_timers(double *compute_ms, double *bc_ms, double *comm_ms, double *intp_ms) {
  MUSB_TRY(needReady());
  MUSB_CUDA(cudaStreamSynchronize(g.stream));
  for (auto &s : g.spans) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s.a, s.b);
    g.acc[s.cat] += ms;
    g.pool.push_back(s.a);
    g.pool.push_back(s.b);
  }
  g.spans.clear();
  if (compute_ms) *compute_ms = g.acc[T_COMPUTE];
  if (bc_ms) *bc_ms = g.acc[T_BC];
  if (comm_ms) *comm_ms = g.acc[T_COMM];
  if (intp_ms) *intp_ms = g.acc[T_INTP];
  return 0;
}

int musb200_timers_reset(void) {
  MUSB_TRY(musb200_timers(nullptr, nullptr, nullptr, nullptr));
  for (double &a : g.acc) a = 0.0;
  g.launches = 0;
  return 0;
}

int musb200_launch_count(long long *n) {
  if (!n) return setError(MUSB200_ERR_ARG, "null argument");
  *n = g.launches;
  return 0;
}

int musb200_event_mark(int which) {
  MUSB_TRY(needReady());
  if (which < 0 || which > 1) return setError(MUSB200_ERR_ARG, "which must be 0|1");
  MUSB_CUDA(cudaEventRecord(g.mark[which], g.stream));
  return 0;
}

int musb200_event_elapsed(double *ms) {
  MUSB_TRY(needReady());
  if (!ms) return setError(MUSB200_ERR_ARG, "null argument");
  MUSB_CUDA(cudaEventSynchronize(g.mark[1]));
  float f = 0.f;
  MUSB_CUDA(cudaEventElapsedTime(&f, g.mark[0], g.mark[1]));
  *ms = f;
  return 0;
}

int musb200_host_alloc(size_t bytes, void **ptr) {
  if (!ptr) return setError(MUSB200_ERR_ARG, "null argument");
  MUSB_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return 0;
}

int musb200_host_free(void *ptr) {
  if (ptr) MUSB_CUDA(cudaFreeHost(ptr));
  return 0;
}

}  // extern "C"
