// sweep_force.cu -- the fused sweep with the body-force source term
// (mus_apply_sourceTerms, mus/source/mus_source_module.f90:430-512, and the addSrcToAuxField
// step of mus_calcAuxFieldAndExchange, mus/source/mus_auxField_module.f90:341-375) folded in:
// steps 5-8 of do_fast_singleLevel in one pass over HBM.  Extra traffic: 24 B per update for a
// per-element force field, none for a uniform force.
#include "sweep_kernel.cuh"

namespace musb200 {

int launchSweepForce(int QQ, int relax, int kind, const SweepArgs &a, cudaStream_t st) {
  return dispatchSweep<1>(QQ, relax, kind, a, st);
}

}  // namespace musb200
