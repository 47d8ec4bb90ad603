// sweep.cu -- the fused per-level sweep: pull-stream through the encoded
// neighbour list, pre-collision moments (the reference's separate
// mus_calcAuxField pass, mus_auxFieldVar_module.fpp:605-817), relaxation
// parameter (mus_update_relaxParamKine) and collision, one pass over HBM.
//
// Replaces the pointee of scheme%compute (kernel interface,
// mus/source/scheme/mus_scheme_type_module.f90:204-235) plus steps 5-7 of
// do_fast_singleLevel (mus/source/mus_control_module.f90:564-617).
//
// Data layout: SoA, f[q][e] = state[q*S + e]; one thread per element, so every
// store and the rest-direction load are fully coalesced 8-byte accesses; the
// QQ-1 pulls go through the 4-byte encoded neighbour list (coalesced) and gather
// from the source rows (Morton order keeps them within a few sectors per warp).
// Algorithmic HBM traffic per lattice update: 2*QQ*8 B (PDF read + write) +
// (QQ-1)*4 B (neighbour list) = 376 B (D3Q19) / 536 B (D3Q27).
#include "sweep_kernel.cuh"

namespace musb200 {

int launchSweepForce(int QQ, int relax, int kind, const SweepArgs &a, cudaStream_t st);  // sweep_force.cu
int launchSweepPush(int QQ, int relax, int kind, const SweepArgs &a, cudaStream_t st);   // sweep_push.cu

int launchSweep(int QQ, int relax, int kind, const SweepArgs &a, cudaStream_t st) {
  if (a.force_order != 0) return launchSweepForce(QQ, relax, kind, a, st);
  if (a.push.mask != nullptr) return launchSweepPush(QQ, relax, kind, a, st);
  return dispatchSweep<0>(QQ, relax, kind, a, st);
}

int sweepBlockSize(int QQ) { return QQ == 27 ? sweepThreads<27>() : sweepThreads<19>(); }

int launchAuxOnly(int QQ, int kind, const SweepArgs &a, cudaStream_t st) {
  if (a.count <= 0) return 0;
  const int grid = divUp(a.count, 128);
  if (QQ == 19 && kind == 0) auxOnlyKernel<19, false><<<grid, 128, 0, st>>>(a);
  else if (QQ == 19) auxOnlyKernel<19, true><<<grid, 128, 0, st>>>(a);
  else if (kind == 0) auxOnlyKernel<27, false><<<grid, 128, 0, st>>>(a);
  else auxOnlyKernel<27, true><<<grid, 128, 0, st>>>(a);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace musb200
