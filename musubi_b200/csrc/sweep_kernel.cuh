// sweep_kernel.cuh -- the fused sweep kernel template, instantiated without the force source in
// sweep.cu and with it in sweep_force.cu (two translation units so they compile in parallel).
#pragma once
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

// exact product with a lattice component c in {-1, 0, 1}
__device__ __forceinline__ double mulc(int c, double x) { return c == 0 ? 0.0 : (c > 0 ? x : -x); }

// Body-force source term fused into the sweep (source = { force = ... }):
//   order 2: velocity shift F/(2 rho) of mus_addForceToAuxField_fluid / _fluidIncomp
//            (mus_auxFieldVar_module.fpp:1032-1214) before the collision, then per direction
//            applySrc_force (bgk, trt; mus_derQuan_module.fpp:3129-3243),
//            applySrc_force_MRT_d3q19 (:3716-3862) or applySrc_force_MRT_d3q27 (:3555-3693)
//   order 1: applySrc_force1stOrd (:4043-4143), no velocity shift
// added to the post-collision value in the same operation order as the reference; matrix
// entries and lattice components that are zero are skipped (x + 0*y = x).
template <int QQ, int RELAX>
struct ForceSrc {
  double Fx, Fy, Fz, ux, uy, uz;
  double ofac;   // bgk / trt: 1 - omega/2
  double sK, sB; // mrt: 1 - omegaKine/2, 1 - omegaBulk/2
  double m[9];   // mrt: the nine non-zero force moments, ascending moment index
  int order;

  __device__ __forceinline__ void prepare(double omega, double omegaBulk) {
    if (order != 2) return;
    if (RELAX != 2) { ofac = 1.0 - omega * 0.5; return; }
    sK = 1.0 - 0.5 * omega;
    sB = 1.0 - 0.5 * omegaBulk;
    const double fu2 = 2.0 * (Fx * ux + Fy * uy + Fz * uz);
    const double m10 = -2.0 * (Fy * uy - 2.0 * Fx * ux + Fz * uz);
    const double m12 = 2.0 * (Fy * uy - Fz * uz);
    const double mxy = Fx * uy + Fy * ux, myz = Fy * uz + Fz * uy, mxz = Fx * uz + Fz * ux;
    if (QQ == 19) {  // momForce(2,4,6,8,10,12,14,15,16)
      m[0] = fu2; m[1] = Fx; m[2] = Fy; m[3] = Fz; m[4] = m10; m[5] = m12; m[6] = mxy; m[7] = myz; m[8] = mxz;
    } else {         // momForce(2..10)
      m[0] = Fx; m[1] = Fy; m[2] = Fz; m[3] = mxy; m[4] = myz; m[5] = mxz; m[6] = m10; m[7] = m12; m[8] = fu2;
    }
  }
  // 0-based moment index and relaxation factor of the j-th non-zero force moment
  static __device__ __forceinline__ constexpr int momIdx(int j) {
    if (QQ == 19) { constexpr int t[9] = {1, 3, 5, 7, 9, 11, 13, 14, 15}; return t[j]; }
    return j + 1;
  }
  __device__ __forceinline__ double sOf(int j) const {
    if (QQ == 19) return j == 0 ? sB : (j < 4 ? 1.0 : sK);
    return j < 3 ? 1.0 : (j < 8 ? sK : sB);
  }
  static __device__ __forceinline__ constexpr double toPdf(int q, int k) {
    return QQ == 19 ? mmIvD3Q19(q, k) : wmmIvD3Q27(q, k);
  }

  __device__ __forceinline__ double term(int q) const {
    const int c0 = cx<QQ>(q, 0), c1 = cx<QQ>(q, 1), c2 = cx<QQ>(q, 2);
    const double w = weight<QQ>(q);
    if (order == 1) {
      const double ft = mulc(c0, Fx) + mulc(c1, Fy) + mulc(c2, Fz);
      return w * 3.0 * ft;
    }
    if (RELAX != 2) {
      const double ucx = mulc(c0, ux) + mulc(c1, uy) + mulc(c2, uz);
      const double t0 = ((double)c0 - ux) * 3.0 + mulc(c0, ucx) * 9.0;
      const double t1 = ((double)c1 - uy) * 3.0 + mulc(c1, ucx) * 9.0;
      const double t2 = ((double)c2 - uz) * 3.0 + mulc(c2, ucx) * 9.0;
      const double ft = t0 * Fx + t1 * Fy + t2 * Fz;
      return ofac * w * ft;
    }
    double disc = 0.0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const double A = toPdf(q, momIdx(j));
      if (A != 0.0) disc = disc + (A * sOf(j)) * m[j];
    }
    return disc;
  }
};

// VAR: 0 = plain sweep, 1 = with the body-force source (sweep_force.cu), 2 = with the halo push
// fused in (sweep_push.cu): an element that owns links of the halo send buffer stores them
// straight into the halo rows of the receiving ranks' state arrays (peer-mapped) right after
// its collision, so the transfer rides on the sweep and only the arrival handshake is left
// for after it (signalHaloKernel, p2p.cu).
template <int QQ, int RELAX, bool INCOMP, int VAR>
__global__ void __launch_bounds__(sweepThreads<QQ>(), sweepMinBlocks<QQ>()) sweepKernel(const SweepArgs a) {
  constexpr bool FORCE = VAR == 1, PUSH = VAR == 2;
  // a CTA always sweeps blockDim.x CONSECUTIVE elements (full coalescing); which block of them is
  // blockIdx.x, or -- second launch of the overlapped exchange -- an entry of the halo-CTA list
  const bool appended = a.ctaMode == 1 && (int)blockIdx.x >= a.nMain;
  const int cta = appended ? a.ctaList[(int)blockIdx.x - a.nMain] : (int)blockIdx.x;
  const int i = cta * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  const int e = a.first + i;
  const long long S = a.S;

  double f[QQ];
  {
    // several ranks, peer-memory halo exchange: a CTA that pulls from a halo row waits until the
    // peers' links of the previous step have arrived (p2p.cu); all other CTAs run right away.
    // The CTA's mask word is fetched together with the index loads, so the check costs no extra
    // memory round trip at the head of every CTA.
    uint32_t maskWord = 0u;
    if (a.wait.ctaMask != nullptr && !appended) maskWord = __ldg(a.wait.ctaMask + (cta >> 5));
    uint32_t n[QQ - 1];
#pragma unroll
    for (int q = 0; q < QQ - 1; ++q) n[q] = __ldg(a.nbr + q * S + e);
    const bool masked = (maskWord >> (cta & 31)) & 1u;   // uniform over the CTA
    if (masked && a.ctaMode == 1) return;                // swept by its appended twin at the end of the launch
    const bool halo = masked || appended;
    if (halo) {
      if (threadIdx.x == 0) waitHaloArrival(a.wait);   // thread 0 of a CTA is always in range
      __syncthreads();                                 // threads out of range have exited
    }
    if (!halo) {
#pragma unroll
      for (int q = 0; q < QQ - 1; ++q) {
        const long long row = (n[q] & kBounceBit) ? (long long)invDir<QQ>(q) * S : (long long)q * S;
        f[q] = __ldg(a.in + row + (n[q] & kElemMask));
      }
    } else {
      // halo rows were written by a peer DURING this kernel: not read-only data, so no
      // ld.global.nc and no L1 (a line shared with own elements may sit there stale) -- they
      // come from L2; the CTA's pulls from own elements keep the read-only path
      const uint32_t h0 = (uint32_t)a.wait.haloStart;
#pragma unroll
      for (int q = 0; q < QQ - 1; ++q) {
        const long long row = (n[q] & kBounceBit) ? (long long)invDir<QQ>(q) * S : (long long)q * S;
        const uint32_t src = n[q] & kElemMask;
        const double *ptr = a.in + row + src;
        f[q] = src >= h0 ? __ldcg(ptr) : __ldg(ptr);
      }
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
  ForceSrc<QQ, RELAX> fs;
  if (FORCE) {
    if (a.force != nullptr) {
      fs.Fx = __ldg(a.force + e);
      fs.Fy = __ldg(a.force + S + e);
      fs.Fz = __ldg(a.force + 2 * S + e);
    } else {
      fs.Fx = a.force_uniform[0]; fs.Fy = a.force_uniform[1]; fs.Fz = a.force_uniform[2];
    }
    fs.order = a.force_order;
    if (fs.order == 2) {
      const double inv_rho = INCOMP ? 1.0 : 1.0 / rho;
      ux = ux + fs.Fx * 0.5 * inv_rho;
      uy = uy + fs.Fy * 0.5 * inv_rho;
      uz = uz + fs.Fz * 0.5 * inv_rho;
    }
    fs.ux = ux; fs.uy = uy; fs.uz = uz;
  }
  if (a.write_aux) {
    a.aux[e] = rho;
    a.aux[S + e] = ux;
    a.aux[2 * S + e] = uy;
    a.aux[3 * S + e] = uz;
  }
  const double omega = (a.omega != nullptr) ? __ldg(a.omega + e) : a.rp.omega_uniform;

  double *out = a.out + e;
  if (FORCE) fs.prepare(omega, a.rp.omega_bulk);
  auto st = [&](int q, double v) {
    if (FORCE) v = v + fs.term(q);
    out[(long long)q * S] = v;
  };
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
  if (PUSH) {
    const uint32_t w = a.push.mask[e >> 5];
    if ((w >> (e & 31)) & 1u) {
      const int k = (int)a.push.prefix[e >> 5] + __popc(w & ((1u << (e & 31)) - 1u));
      const int j1 = a.push.start[k + 1];
      for (int j = a.push.start[k]; j < j1; ++j) {
        const double v = out[(long long)a.push.entQ[j] * S];   // stored above by this thread
        const int pk = a.push.entPeer[j];
        const int r = a.push.entDst[j] - 1;
        a.push.remoteState[pk][(long long)(r % QQ) * a.push.remoteS[pk] + r / QQ] = v;
      }
    }
  }
}

// auxField on demand (lazy auxField, tracking of single elements): the moments of the PDFs an
// element pulls from state(:, now) -- exactly what the sweep of the same step computed, because
// state(:, now) is not touched until the next step -- including the half-force velocity shift.
template <int QQ, bool INCOMP>
__global__ void __launch_bounds__(128) auxOnlyKernel(const SweepArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  const int e = a.first + i;
  const long long S = a.S;
  double f[QQ];
  if (a.nbr == nullptr) {
    // the element's own PDFs (SAVE access): auxFieldFromState of mus_initAuxField
#pragma unroll
    for (int q = 0; q < QQ - 1; ++q) f[q] = a.in[(long long)q * S + e];
  } else {
#pragma unroll
    for (int q = 0; q < QQ - 1; ++q) {
      const uint32_t n = a.nbr[q * S + e];
      const long long row = (n & kBounceBit) ? (long long)invDir<QQ>(q) * S : (long long)q * S;
      f[q] = a.in[row + (n & kElemMask)];
    }
  }
  f[QQ - 1] = a.in[(long long)(QQ - 1) * S + e];
  double rho, ux, uy, uz;
  moments<QQ>(f, rho, ux, uy, uz);
  if (!INCOMP) {
    ux = ux / rho;
    uy = uy / rho;
    uz = uz / rho;
  }
  if (a.force_order == 2) {
    double Fx, Fy, Fz;
    if (a.force != nullptr) { Fx = a.force[e]; Fy = a.force[S + e]; Fz = a.force[2 * S + e]; }
    else { Fx = a.force_uniform[0]; Fy = a.force_uniform[1]; Fz = a.force_uniform[2]; }
    const double inv_rho = INCOMP ? 1.0 : 1.0 / rho;
    ux = ux + Fx * 0.5 * inv_rho;
    uy = uy + Fy * 0.5 * inv_rho;
    uz = uz + Fz * 0.5 * inv_rho;
  }
  a.aux[e] = rho;
  a.aux[S + e] = ux;
  a.aux[2 * S + e] = uy;
  a.aux[3 * S + e] = uz;
}

template <int QQ, int RELAX, bool INCOMP, int VAR>
static int launchT(const SweepArgs &a, cudaStream_t st) {
  if (a.count <= 0) return 0;
  const int block = sweepThreads<QQ>();
  const int grid = a.ctaMode == 1 ? a.nMain + a.nCtas : divUp(a.count, block);
  if (grid <= 0) return 0;
  sweepKernel<QQ, RELAX, INCOMP, VAR><<<grid, block, 0, st>>>(a);
  MUSB_CUDA(cudaGetLastError());
  return 0;
}


// (layout, relaxation, kind) -> instantiation; VAR selects the translation unit
template <int VAR>
static int dispatchSweep(int QQ, int relax, int kind, const SweepArgs &a, cudaStream_t st) {
  if (kind == 1) {
    // mus_init_advRel_fluid_incompressible (init/mus_initFluidIncomp_module.f90:73-218):
    // trt exists for d3q19 only
    if (QQ == 19 && relax == 0) return launchT<19, 0, true, VAR>(a, st);
    if (QQ == 19 && relax == 1) return launchT<19, 1, true, VAR>(a, st);
    if (QQ == 19 && relax == 2) return launchT<19, 2, true, VAR>(a, st);
    if (QQ == 27 && relax == 0) return launchT<27, 0, true, VAR>(a, st);
    if (QQ == 27 && relax == 2) return launchT<27, 2, true, VAR>(a, st);
    return setError(4, "fluid_incompressible: the reference has no trt kernel for this layout");
  }
  if (QQ == 19) {
    if (relax == 0) return launchT<19, 0, false, VAR>(a, st);
    if (relax == 1) return launchT<19, 1, false, VAR>(a, st);
    if (relax == 2) return launchT<19, 2, false, VAR>(a, st);
  } else if (QQ == 27) {
    if (relax == 0) return launchT<27, 0, false, VAR>(a, st);
    if (relax == 1) return launchT<27, 1, false, VAR>(a, st);
    if (relax == 2) return launchT<27, 2, false, VAR>(a, st);
  }
  return setError(4, "no kernel for this (layout, relaxation)");
}

}  // namespace musb200
