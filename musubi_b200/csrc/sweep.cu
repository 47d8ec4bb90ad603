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
#include "kernels.cuh"

namespace musb200 {

// Launch shape: 128 threads per CTA.  D3Q19 is capped at 80 registers (__launch_bounds__(128, 6):
// 6 CTAs = 24 warps per SM, at most 24 B of spills in the MRT variants); uncapped the compiler
// takes 94-110 registers, only 16 warps fit and the TRT sweep of 256^3 drops from 0.961 ms
// (1.01 of the measured HBM peak) to 1.079 ms (0.90).  D3Q27 needs up to 128 registers
// (4 CTAs/SM, no spills).  A warp lives long here (26 index loads -> 27 gathers -> 800-1300 FP64
// instructions -> 27 stores) and a CTA's registers are only released when its last warp retires,
// so small CTAs keep more loads in flight.  Measured on B200 (profiles/r01_launch_shape.md):
// D3Q19 TRT 256^3, 80 registers: 64 / 128 / 192 / 256 threads = 0.966 / 0.961 / 0.966 / 0.964 ms,
// 94 registers (5 CTAs) 0.984 ms, 72 registers (7 CTAs, 52 B spills) 0.987 ms;
// D3Q27 MRT 256^3: 256 threads 2.30 ms, 128 threads 1.55 ms, 64 threads 1.57 ms, 512 threads
// 1.81 ms.  Capping D3Q27 at 96 or 80 registers (5-6 CTAs/SM) spills 270-570 B per thread and
// is slower (1.98 / 2.64 ms).
#ifndef SWEEP27_THREADS
#define SWEEP27_THREADS 128
#endif
#ifndef SWEEP27_MINBLOCKS
#define SWEEP27_MINBLOCKS 4
#endif
#ifndef SWEEP19_THREADS
#define SWEEP19_THREADS 128
#endif
#ifndef SWEEP19_MINBLOCKS
#define SWEEP19_MINBLOCKS 6
#endif
template <int QQ>
constexpr int sweepThreads() { return QQ == 27 ? SWEEP27_THREADS : SWEEP19_THREADS; }
template <int QQ>
constexpr int sweepMinBlocks() { return QQ == 27 ? SWEEP27_MINBLOCKS : SWEEP19_MINBLOCKS; }

template <int QQ, int RELAX, bool INCOMP>
__global__ void __launch_bounds__(sweepThreads<QQ>(), sweepMinBlocks<QQ>()) sweepKernel(const SweepArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  int e;
  if (a.list != nullptr) {
    e = a.list[i];
  } else {
    e = a.first + i;
    if (a.skip != nullptr && ((a.skip[e >> 5] >> (e & 31)) & 1u)) return;
  }
  const long long S = a.S;

  double f[QQ];
  {
    uint32_t n[QQ - 1];
#pragma unroll
    for (int q = 0; q < QQ - 1; ++q) n[q] = __ldcs(a.nbr + q * S + e);
#pragma unroll
    for (int q = 0; q < QQ - 1; ++q) {
      const long long row = (n[q] & kBounceBit) ? (long long)invDir<QQ>(q) * S : (long long)q * S;
      f[q] = __ldg(a.in + row + (n[q] & kElemMask));
    }
    f[QQ - 1] = __ldg(a.in + (long long)(QQ - 1) * S + e);
  }

  double rho, ux, uy, uz;
  moments<QQ>(f, rho, ux, uy, uz);
  if (!INCOMP) {
    ux = ux / rho;
    uy = uy / rho;
    uz = uz / rho;
  }
  if (a.write_aux) {
    __stcs(a.aux + e, rho);
    __stcs(a.aux + S + e, ux);
    __stcs(a.aux + 2 * S + e, uy);
    __stcs(a.aux + 3 * S + e, uz);
  }
  const double omega = (a.omega != nullptr) ? __ldcs(a.omega + e) : a.rp.omega_uniform;

  double *out = a.out + e;
  auto st = [&](int q, double v) { __stcs(out + (long long)q * S, v); };
  if (QQ == 19) {
    const double(&g)[19] = reinterpret_cast<const double(&)[19]>(f);
    if (RELAX == 0) collide_bgk_d3q19<INCOMP>(g, rho, ux, uy, uz, omega, st);
    if (RELAX == 1 && !INCOMP) collide_trt_d3q19(g, rho, ux, uy, uz, omega, a.rp.lambda, st);
    if (RELAX == 1 && INCOMP) collide_trt_d3q19_incomp(g, rho, ux, uy, uz, omega, a.rp.lambda, st);
    if (RELAX == 2) collide_mrt_d3q19<INCOMP>(g, rho, ux, uy, uz, omega, a.rp.omega_bulk, st);
  } else {
    const double(&g)[27] = reinterpret_cast<const double(&)[27]>(f);
    if (RELAX == 0) collide_bgk_d3q27<INCOMP>(g, rho, ux, uy, uz, omega, st);
    if (RELAX == 1) collide_trt_d3q27(g, rho, ux, uy, uz, omega, a.rp.lambda, st);
    if (RELAX == 2) collide_mrt_d3q27<INCOMP>(g, rho, ux, uy, uz, omega, a.rp.omega_bulk, st);
  }
}

template <int QQ, int RELAX, bool INCOMP>
static int launchT(const SweepArgs &a, cudaStream_t st) {
  if (a.count <= 0) return 0;
  const int block = sweepThreads<QQ>();
  sweepKernel<QQ, RELAX, INCOMP><<<divUp(a.count, block), block, 0, st>>>(a);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchSweep(int QQ, int relax, int kind, const SweepArgs &a, cudaStream_t st) {
  if (kind == 1) {
    // mus_init_advRel_fluid_incompressible (init/mus_initFluidIncomp_module.f90:73-218):
    // trt exists for d3q19 only
    if (QQ == 19 && relax == 0) return launchT<19, 0, true>(a, st);
    if (QQ == 19 && relax == 1) return launchT<19, 1, true>(a, st);
    if (QQ == 19 && relax == 2) return launchT<19, 2, true>(a, st);
    if (QQ == 27 && relax == 0) return launchT<27, 0, true>(a, st);
    if (QQ == 27 && relax == 2) return launchT<27, 2, true>(a, st);
    return setError(4, "fluid_incompressible: the reference has no trt kernel for this layout");
  }
  if (QQ == 19) {
    if (relax == 0) return launchT<19, 0, false>(a, st);
    if (relax == 1) return launchT<19, 1, false>(a, st);
    if (relax == 2) return launchT<19, 2, false>(a, st);
  } else if (QQ == 27) {
    if (relax == 0) return launchT<27, 0, false>(a, st);
    if (relax == 1) return launchT<27, 1, false>(a, st);
    if (relax == 2) return launchT<27, 2, false>(a, st);
  }
  return setError(4, "no kernel for this (layout, relaxation)");
}

}  // namespace musb200
