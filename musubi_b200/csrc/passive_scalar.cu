// passive_scalar.cu -- fused sweep of the passive-scalar (advection-diffusion) scheme kind:
// pull-stream, zeroth moment (mus_calcAuxField_zerothMoment, mus_auxFieldVar_module.fpp:825-869)
// and the collision of
//   mus_advRel_kPS_rBGK_v1st_l        mus/source/compute/mus_compute_passiveScalar_module.fpp:77-169
//   mus_advRel_kPS_rBGK_v2nd_l        ...:183-279
//   mus_advRel_kPS_rTRT_vStdNoOpt_l   ...:293-398
// selected as mus_init_advRel_lbm_ps does (init/mus_initLBMPS_module.f90:59-159).
//
// The transport velocity (scheme%transVar%method(1), lattice units) is either uniform, a
// per-element SoA array [3][S] uploaded by the host, or -- device-side coupling -- rows 1..3 of
// the auxField of a flow scheme living on the same element list.
// Algorithmic HBM traffic: the fluid sweep's 2*QQ*8 + (QQ-1)*4 B plus 24 B of velocity.
#include "kernels.cuh"

namespace musb200 {

template <int QQ, int VARIANT>
__global__ void __launch_bounds__(128, QQ == 19 ? 6 : 4) passiveScalarKernel(const PsArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.count) return;
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
  double rho = 0.0;  // sum(pdfTmp), sequential
#pragma unroll
  for (int q = 0; q < QQ; ++q) rho = rho + f[q];
  if (a.write_aux) __stcs(a.aux + e, rho);

  double ux, uy, uz;
  if (a.vel != nullptr) {
    ux = __ldcs(a.vel + e);
    uy = __ldcs(a.vel + a.velS + e);
    uz = __ldcs(a.vel + 2 * a.velS + e);
  } else {
    ux = a.vel_uniform[0]; uy = a.vel_uniform[1]; uz = a.vel_uniform[2];
  }
  const double usq = ux * ux + uy * uy + uz * uz;
  double *out = a.out + e;
#pragma unroll
  for (int q = 0; q < QQ; ++q) {
    const int c0 = cx<QQ>(q, 0), c1 = cx<QQ>(q, 1), c2 = cx<QQ>(q, 2);
    const double w = weight<QQ>(q);
    // dble(cxDir) * u: exact for components in {-1, 0, 1}
    const double uc = (c0 == 0 ? 0.0 : (c0 > 0 ? ux : -ux)) + (c1 == 0 ? 0.0 : (c1 > 0 ? uy : -uy)) +
                      (c2 == 0 ? 0.0 : (c2 > 0 ? uz : -uz));
    double v;
    if (VARIANT == 1) {
      const double feq = rho * w * (1.0 + 3.0 * uc);
      v = f[q] + a.d_omega * (feq - f[q]);
    } else if (VARIANT == 2) {
      const double feq = rho * w * (1.0 + 3.0 * uc + 9.0 * uc * uc * 0.5 - usq * 0.5 * 3.0);
      v = f[q] + a.d_omega * (feq - f[q]);
    } else {
      const double feqPlus = rho * w * (1.0 + 9.0 * uc * uc * 0.5 - usq * 0.5 * 3.0);
      const double feqMinus = rho * w * 3.0 * uc;
      const int qi = invDir<QQ>(q);
      const double fPlus = 0.5 * (f[q] + f[qi]);
      const double fMinus = 0.5 * (f[q] - f[qi]);
      v = f[q] + a.d_omega * (feqMinus - fMinus) + a.aux_omega * (feqPlus - fPlus);
    }
    __stcs(out + (long long)q * S, v);
  }
}

template <int QQ, int VARIANT>
static int launchPsT(const PsArgs &a, cudaStream_t st) {
  if (a.count <= 0) return 0;
  passiveScalarKernel<QQ, VARIANT><<<divUp(a.count, 128), 128, 0, st>>>(a);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}

int launchPassiveScalar(int QQ, int variant, const PsArgs &a, cudaStream_t st) {
  if (QQ == 19) {
    if (variant == 1) return launchPsT<19, 1>(a, st);
    if (variant == 2) return launchPsT<19, 2>(a, st);
    if (variant == 3) return launchPsT<19, 3>(a, st);
  } else if (QQ == 27) {
    if (variant == 1) return launchPsT<27, 1>(a, st);
    if (variant == 2) return launchPsT<27, 2>(a, st);
    if (variant == 3) return launchPsT<27, 3>(a, st);
  }
  return setError(4, "passive_scalar: no kernel for this (layout, relaxation, variant)");
}

}  // namespace musb200
