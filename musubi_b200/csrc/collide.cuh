// collide.cuh -- the collision operators of the six in-scope kernels as device
// functions acting on the QQ pulled PDFs held in registers.
//
// Arithmetic (operation order kept so that, compiled with -fmad=false, results
// are bit-identical to the reference algorithm evaluated in IEEE double):
//   BGK D3Q19   mus/source/compute/mus_compute_d3q19_module.fpp:483-640
//   fluid_incompressible: BGK D3Q19 ...:1596-1700, TRT D3Q19 ...:2782-2959,
//               MRT D3Q19 mus_compute_mrt_d3q19_module.fpp:466-739,
//               MRT D3Q27 mus_compute_mrt_d3q27_module.fpp:374-546,
//               BGK D3Q27 = the kCFD kernel with get_pdfEq_incomp_d3q27
//   TRT D3Q19   mus/source/compute/mus_compute_d3q19_module.fpp:2644-2763
//   MRT D3Q19   mus/source/compute/mus_compute_mrt_d3q19_module.fpp:238-450
//   BGK D3Q27   mus/source/compute/mus_compute_d3q27_module.fpp:398-517
//               + get_pdfEq_d3q27 (mus_scheme_derived_quantities_type_module.f90:688-745)
//   TRT D3Q27   mus/source/compute/mus_compute_d3q27_module.fpp:601-740
//   MRT D3Q27   mus/source/compute/mus_compute_mrt_d3q27_module.fpp:255-361
//               + WMMIvD3Q27 (mus/source/init/mus_mrtInit_module.f90:225-301)
//   rho, u      mus/source/derived/mus_auxFieldVar_module.fpp:655-697 and
//               get_vel_from_pdf_d3q19/_d3q27 (...derived_quantities...:1011-1018, 1109-1119)
//
// `St` is a functor  void operator()(int q, double v)  that stores the
// post-collision PDF of 0-based direction q (a coalesced SoA store).
#pragma once
#include "common.cuh"
#include "equilibrium.cuh"
#include "mrt_tables.cuh"
#include <utility>

namespace musb200 {

struct RelaxParams {
  double omega_uniform;
  double lambda;      // TRT magic parameter (fluid%lambda)
  double omega_bulk;  // MRT: fluid%omegaBulkLvl(level)
};

// ---------------------------------------------------------------------------
// moments: rho = sum(pdf) sequential; momentum in the literal +/- order
template <int QQ>
__device__ __forceinline__ void moments(const double (&f)[QQ], double &rho, double &mx,
                                        double &my, double &mz) {
  double r = 0.0;
#pragma unroll
  for (int q = 0; q < QQ; ++q) r = r + f[q];
  rho = r;
  // p(k) of the reference is f[k-1]
  mx = f[3] - f[0] - f[10] + f[11] - f[12] + f[13] - f[14] - f[15] + f[16] + f[17];
  my = f[4] - f[1] - f[6] - f[7] + f[8] + f[9] - f[14] + f[15] - f[16] + f[17];
  mz = f[5] - f[2] - f[6] + f[7] - f[8] + f[9] - f[10] - f[11] + f[12] + f[13];
  if (QQ == 27) {
    mx = mx - f[18] - f[19] - f[20] - f[21] + f[22] + f[23] + f[24] + f[25];
    my = my - f[18] - f[19] + f[20] + f[21] - f[22] - f[23] + f[24] + f[25];
    mz = mz - f[18] + f[19] - f[20] + f[21] - f[22] + f[23] - f[24] + f[25];
  }
}

// ---------------------------------------------------------------------------
template <bool INCOMP, class St>
__device__ __forceinline__ void collide_bgk_d3q19(const double (&f)[19], double rho, double u_x,
                                                  double u_y, double u_z, double omega, St st) {
  constexpr double div1_3 = 1.0 / 3.0, div1_8 = 1.0 / 8.0, div1_36 = 1.0 / 36.0;
  constexpr double div3_4h = 3.0 / 4.5;
  const double usq = (u_x * u_x) + (u_y * u_y) + (u_z * u_z);
  const double cmpl_o = 1.0 - omega;
  double coeff_1, coeff_2, usqn_o1, usqn_o2;
  if (!INCOMP) {
    const double usqn = div1_36 * (1.0 - 1.5 * usq) * rho;
    st(18, f[18] * cmpl_o + omega * rho * (div1_3 - 0.5 * usq));
    coeff_1 = div1_8 * omega * rho;
    usqn_o1 = omega * usqn;
    const double omega_2 = 2.0 * omega;
    coeff_2 = div1_8 * omega_2 * rho;
    usqn_o2 = omega_2 * usqn;
  } else {
    usqn_o1 = omega * div1_36 * (rho - 1.5 * usq);
    st(18, f[18] * cmpl_o + 12.0 * usqn_o1);
    coeff_1 = div1_8 * omega;
    coeff_2 = div1_8 * omega * 2.0;
    usqn_o2 = 2.0 * usqn_o1;
  }
  auto diag = [&](int qp, int qm, double ui) {
    const double fac = coeff_1 * ui;
    const double s1 = fac * div3_4h;
    const double s2 = fac * ui + usqn_o1;
    st(qp, f[qp] * cmpl_o + s1 + s2);
    st(qm, f[qm] * cmpl_o - s1 + s2);
  };
  diag(PP0, NN0, u_x + u_y);
  diag(NP0, PN0, -u_x + u_y);
  diag(PZP, NZN, u_x + u_z);
  diag(NZP, PZN, -u_x + u_z);
  diag(ZPP, ZNN, u_y + u_z);
  diag(ZNP, ZPN, -u_y + u_z);
  auto axis = [&](int qp, int qm, double u) {
    const double fac = coeff_2 * u;
    const double s1 = fac * div3_4h;
    const double s2 = fac * u + usqn_o2;
    st(qp, f[qp] * cmpl_o + s1 + s2);
    st(qm, f[qm] * cmpl_o - s1 + s2);
  };
  axis(ZP0, ZN0, u_y);
  axis(P00, N00, u_x);
  axis(ZZP, ZZN, u_z);
}

// ---------------------------------------------------------------------------
template <class St>
__device__ __forceinline__ void collide_trt_d3q19(const double (&f)[19], double rho, double u_x,
                                                  double u_y, double u_z, double omega,
                                                  double lambda, St st) {
  constexpr double div1_3 = 1.0 / 3.0, t2cs4inv = 4.5;
  constexpr double t1x2_0 = 1.0 / 18.0 * 2.0, t2x2_0 = 1.0 / 36.0 * 2.0;
  const double usq = (u_x * u_x) + (u_y * u_y) + (u_z * u_z);
  const double feq_common = 1.0 - 1.5 * usq;
  const double omega_h = 0.5 * omega;
  const double asym_omega = 1.0 / (0.5 + lambda / (1.0 / omega - 0.5));
  const double asym_omega_h = 0.5 * asym_omega;
  st(18, f[18] * (1.0 - omega) + omega * div1_3 * rho * feq_common);
  auto link = [&](double tx2, double fc, int qp, int qm, double ui) {
    const double sym = omega_h * (f[qp] + f[qm] - fc * ui * ui - tx2 * feq_common);
    const double asym = asym_omega_h * (f[qp] - f[qm] - 3.0 * tx2 * ui);
    st(qp, f[qp] - sym - asym);
    st(qm, f[qm] - sym + asym);
  };
  const double t2x2 = t2x2_0 * rho;
  const double fac2 = t2x2 * t2cs4inv;
  link(t2x2, fac2, PP0, NN0, u_x + u_y);
  link(t2x2, fac2, PN0, NP0, u_x - u_y);
  link(t2x2, fac2, PZP, NZN, u_x + u_z);
  link(t2x2, fac2, PZN, NZP, u_x - u_z);
  link(t2x2, fac2, ZPP, ZNN, u_y + u_z);
  link(t2x2, fac2, ZPN, ZNP, u_y - u_z);
  const double t1x2 = t1x2_0 * rho;
  const double fac1 = t1x2 * t2cs4inv;
  link(t1x2, fac1, P00, N00, u_x);
  link(t1x2, fac1, ZP0, ZN0, u_y);
  link(t1x2, fac1, ZZP, ZZN, u_z);
}

// ---------------------------------------------------------------------------
// mus_advRel_kFluidIncomp_rTRT_vStd_lD3Q19 (mus_compute_d3q19_module.fpp:2782-2959)
template <class St>
__device__ __forceinline__ void collide_trt_d3q19_incomp(const double (&f)[19], double rho,
                                                         double u_x, double u_y, double u_z,
                                                         double omega, double lambda, St st) {
  constexpr double div1_3 = 1.0 / 3.0, div1_6 = 1.0 / 6.0, t2cs4inv = 4.5;
  constexpr double t1x2 = 1.0 / 9.0, t2x2 = 1.0 / 18.0;
  constexpr double fac1 = t1x2 * t2cs4inv, fac2 = t2x2 * t2cs4inv;
  const double usq = (u_x * u_x) + (u_y * u_y) + (u_z * u_z);
  const double feq_common = rho - 1.5 * usq;
  const double omega_h = 0.5 * omega;
  const double asym_omega = 1.0 / (0.5 + lambda / (1.0 / omega - 0.5));
  const double asym_omega_h = 0.5 * asym_omega;
  st(18, f[18] * (1.0 - omega) + omega * div1_3 * feq_common);
  auto link = [&](double fc, double tfeq, double dv, int qp, int qm, double ui) {
    const double sym = omega_h * (f[qp] + f[qm] - fc * ui * ui - tfeq);
    const double asym = asym_omega_h * (f[qp] - f[qm] - dv * ui);
    st(qp, f[qp] - sym - asym);
    st(qm, f[qm] - sym + asym);
  };
  const double t2_feq = t2x2 * feq_common;
  link(fac2, t2_feq, div1_6, PP0, NN0, u_x + u_y);
  link(fac2, t2_feq, div1_6, PN0, NP0, u_x - u_y);
  link(fac2, t2_feq, div1_6, PZP, NZN, u_x + u_z);
  link(fac2, t2_feq, div1_6, PZN, NZP, u_x - u_z);
  link(fac2, t2_feq, div1_6, ZPP, ZNN, u_y + u_z);
  link(fac2, t2_feq, div1_6, ZPN, ZNP, u_y - u_z);
  const double t1_feq = t1x2 * feq_common;
  link(fac1, t1_feq, div1_3, ZP0, ZN0, u_y);
  link(fac1, t1_feq, div1_3, P00, N00, u_x);
  link(fac1, t1_feq, div1_3, ZZP, ZZN, u_z);
}

// ---------------------------------------------------------------------------
template <bool INCOMP, class St>
__device__ __forceinline__ void collide_mrt_d3q19(const double (&f)[19], double rho, double u_x,
                                                  double u_y, double u_z, double omegaKine,
                                                  double omegaBulk, St st) {
  constexpr double div1_4 = 1.0 / 4.0, div1_8 = 1.0 / 8.0, div1_12 = 1.0 / 12.0;
  constexpr double div1_16 = 1.0 / 16.0, div1_24 = 1.0 / 24.0, div1_48 = 1.0 / 48.0;
  constexpr double div1_72 = 1.0 / 72.0;
  // s_mrt of mrt_d3q19 (mus_mrtRelaxation_module.fpp:238-262) pre-scaled as in :244-251
  const double s2 = omegaBulk * div1_24;
  constexpr double s3 = 1.40 * div1_72;
  constexpr double s5 = 1.20 * div1_24, s7 = 1.20 * div1_24, s9 = 1.20 * div1_24;
  constexpr double s11 = 1.40, s13 = 1.40;
  constexpr double s17 = 1.98 * div1_8, s18 = 1.98 * div1_8, s19 = 1.98 * div1_8;
  const double s10 = omegaKine, s12 = omegaKine;
  const double s14 = div1_4 * omegaKine, s15 = div1_4 * omegaKine, s16 = div1_4 * omegaKine;

  const double fN00 = f[N00], f0N0 = f[ZN0], f00N = f[ZZN], f100 = f[P00], f010 = f[ZP0],
               f001 = f[ZZP], f0NN = f[ZNN], f0N1 = f[ZNP], f01N = f[ZPN], f011 = f[ZPP],
               fN0N = f[NZN], f10N = f[PZN], fN01 = f[NZP], f101 = f[PZP], fNN0 = f[NN0],
               fN10 = f[NP0], f1N0 = f[PN0], f110 = f[PP0], f000 = f[18];

  const double m6 = f101 + fN0N + f10N + fN01;
  const double m8 = f011 + f0NN + f01N + f0N1;
  const double sum1 = f110 + fNN0 + f1N0 + fN10;
  const double m2 = -f000 + sum1 + m6 + m8;
  const double sum2 = f010 + f0N0;
  const double sum3 = f001 + f00N;
  const double sum4 = 2.0 * (f100 + fN00);
  const double sum5 = sum2 + sum3;
  const double mout3 = (2.0 * (f000 - sum5) - sum4 + m2) * s3;

  // incompressible (mus_compute_mrt_d3q19_module.fpp:578-603): rho0 = 1 replaces rho
  const double meq2 = INCOMP ? u_x * u_x + u_y * u_y + u_z * u_z
                             : rho * (u_x * u_x + u_y * u_y + u_z * u_z);
  const double meq10 = INCOMP ? 3.0 * u_x * u_x - meq2 : rho * 3.0 * u_x * u_x - meq2;
  const double meq12 = INCOMP ? u_y * u_y - u_z * u_z : rho * (u_y * u_y - u_z * u_z);
  const double mout2 = s2 * (m2 - meq2);
  const double m14 = f110 + fNN0 - f1N0 - fN10;
  const double mout14 = s14 * (m14 - (INCOMP ? u_x * u_y : rho * u_x * u_y));
  const double m15 = f011 + f0NN - f01N - f0N1;
  const double mout15 = s15 * (m15 - (INCOMP ? u_y * u_z : rho * u_y * u_z));
  const double m16 = f101 + fN0N - f10N - fN01;
  const double mout16 = s16 * (m16 - (INCOMP ? u_x * u_z : rho * u_x * u_z));

  const double sum6 = sum1 + m6 - m8 * 2.0;
  const double sum7 = sum4 - sum5;
  const double mout10 = (sum7 + sum6 - meq10) * s10;
  const double mout11 = (-sum7 + sum6) * s11;
  const double sum8 = sum1 - m6;
  const double sum9 = sum2 - sum3;
  const double mout12 = (sum8 + sum9 - meq12) * s12;
  const double mout13 = (sum8 - sum9) * s13;

  double c1 = f110 - fNN0, c2 = f1N0 - fN10;
  double c3 = f101 - fN0N, c4 = f10N - fN01;
  const double sum10 = c1 + c2, sum11 = c3 + c4;
  const double mout5 = (sum10 + sum11 - 2.0 * (f100 - fN00)) * s5;
  const double mout17 = (sum10 - sum11) * s17;
  double c5 = f011 - f0NN;
  const double c6 = f01N - f0N1;
  const double sum12 = c1 - c2, sum13 = c5 + c6;
  const double mout7 = (sum12 + sum13 - 2.0 * (f010 - f0N0)) * s7;
  const double mout18 = (-sum12 + sum13) * s18;
  const double sum14 = c3 - c4, sum15 = c5 - c6;
  const double mout9 = (sum14 + sum15 - 2.0 * (f001 - f00N)) * s9;
  const double mout19 = (sum14 - sum15) * s19;

  st(18, f000 + 12.0 * (mout2 - mout3));

  const double c0 = -4.0 * mout3 + div1_12 * (mout10 - mout11);
  const double mout5_4 = mout5 * 4.0;
  st(P00, f100 - (c0 - mout5_4));
  st(N00, fN00 - (c0 + mout5_4));

  c1 = -4.0 * mout3 - div1_24 * (mout10 - mout11);
  c2 = div1_8 * (mout12 - mout13);
  const double sum_c1_c2 = c1 + c2;
  const double mout7_4 = mout7 * 4.0;
  st(ZP0, f010 - (sum_c1_c2 - mout7_4));
  st(ZN0, f0N0 - (sum_c1_c2 + mout7_4));
  const double sub_c1_c2 = c1 - c2;
  const double mout9_4 = mout9 * 4.0;
  st(ZZP, f001 - (sub_c1_c2 - mout9_4));
  st(ZZN, f00N - (sub_c1_c2 + mout9_4));

  const double mout1 = mout2 + mout3;
  c3 = mout1 + div1_48 * (mout10 + mout11) + div1_16 * (mout12 + mout13);
  const double sum_5_17 = mout5 + mout17, sub_7_18 = mout7 - mout18;
  const double d1 = c3 + mout14, d2 = sum_5_17 + sub_7_18;
  st(PP0, f110 - (d1 + d2));
  st(NN0, fNN0 - (d1 - d2));
  const double d3 = c3 - mout14, d4 = sum_5_17 - sub_7_18;
  st(PN0, f1N0 - (d3 + d4));
  st(NP0, fN10 - (d3 - d4));

  c4 = c3 - div1_8 * (mout12 + mout13);
  const double sum_9_19 = mout9 + mout19, sub_5_17 = mout5 - mout17;
  const double e1 = c4 + mout16, e2 = sum_9_19 + sub_5_17;
  st(PZP, f101 - (e1 + e2));
  st(NZN, fN0N - (e1 - e2));
  const double e3 = c4 - mout16, e4 = sum_9_19 - sub_5_17;
  st(PZN, f10N - (e3 - e4));
  st(NZP, fN01 - (e3 + e4));

  c5 = mout1 - div1_24 * (mout10 + mout11);
  const double sum_7_18 = mout7 + mout18, sub_9_19 = mout9 - mout19;
  const double g1 = c5 + mout15, g2 = sum_7_18 + sub_9_19;
  st(ZPP, f011 - (g1 + g2));
  st(ZNN, f0NN - (g1 - g2));
  const double g3 = c5 - mout15, g4 = sum_7_18 - sub_9_19;
  st(ZPN, f01N - (g3 + g4));
  st(ZNP, f0N1 - (g3 - g4));
}

// ---------------------------------------------------------------------------
// mus_advRel_kCFD_rBGK_vStd_lD3Q27: out = f - omega*(f - fEq), fEq = pdfEq_ptr(rho, vel):
// get_pdfEq_d3q27 (fluid) or get_pdfEq_incomp_d3q27 (fluid_incompressible)
template <bool INCOMP, class St>
__device__ __forceinline__ void collide_bgk_d3q27(const double (&f)[27], double rho, double vx,
                                                  double vy, double vz, double omega, St st) {
  double feq[27];
  if (INCOMP) pdfEqIncompD3Q27(rho, vx, vy, vz, feq);
  else pdfEqD3Q27(rho, vx, vy, vz, feq);
#pragma unroll
  for (int q = 0; q < 27; ++q) st(q, f[q] - omega * (f[q] - feq[q]));
}

// ---------------------------------------------------------------------------
// product-form equilibrium  fEq = -rho * X * Y * Z
template <class St>
__device__ __forceinline__ void collide_trt_d3q27(const double (&f)[27], double rho, double u,
                                                  double v, double w, double wP, double lambda,
                                                  St st) {
  constexpr double div2_3 = 2.0 / 3.0, div1_2 = 1.0 / 2.0;
  const double u2 = u * u, v2 = v * v, w2 = w * w;
  // index c+1: [0] = XN (c=-1), [1] = X0, [2] = X1 (c=+1)
  double X[3], Y[3], Z[3];
  X[1] = -div2_3 + u2; X[2] = -(X[1] + 1.0 + u) * 0.5; X[0] = X[2] + u;
  Y[1] = -div2_3 + v2; Y[2] = -(Y[1] + 1.0 + v) * 0.5; Y[0] = Y[2] + v;
  Z[1] = -div2_3 + w2; Z[2] = -(Z[1] + 1.0 + w) * 0.5; Z[0] = Z[2] + w;
  const double wN = 1.0 / (0.5 + lambda / (1.0 / wP - 0.5));
  st(26, (1.0 - wP) * f[26] - rho * wP * X[1] * Y[1] * Z[1]);
  // the 13 (+c,-c) pairs of :656-740
  constexpr int pairs[13][2] = {{P00, N00}, {ZP0, ZN0}, {ZZP, ZZN}, {ZPP, ZNN}, {ZPN, ZNP},
                                {PZP, NZN}, {PZN, NZP}, {PP0, NN0}, {PN0, NP0}, {PNN, NPP},
                                {PPN, NNP}, {PNP, NPN}, {PPP, NNN}};
#pragma unroll
  for (int k = 0; k < 13; ++k) {
    const int dp = pairs[k][0], dm = pairs[k][1];
    const int c0 = cx<27>(dp, 0), c1 = cx<27>(dp, 1), c2 = cx<27>(dp, 2);
    const double Xp = X[c0 + 1], Yp = Y[c1 + 1], Zp = Z[c2 + 1];
    const double Xm = X[-c0 + 1], Ym = Y[-c1 + 1], Zm = Z[-c2 + 1];
    const double p_part =
        wP * ((f[dp] + f[dm]) - (-rho * Xp * Yp * Zp - rho * Xm * Ym * Zm)) * div1_2;
    const double n_part =
        wN * ((f[dp] - f[dm]) - (-rho * Xp * Yp * Zp + rho * Xm * Ym * Zm)) * div1_2;
    st(dp, f[dp] - p_part - n_part);
    st(dm, f[dm] - p_part + n_part);
  }
}

// ---------------------------------------------------------------------------
// sparse, fully unrolled  M^-1 * mneq  of D3Q27 (WMMIvD3Q27), entries folded at compile time
template <int D, int J>
__device__ __forceinline__ void mrt27Acc(double &acc, const double (&m)[27]) {
  constexpr double w = wmmIvD3Q27(D, J);
#ifdef MRT27_FMA   // experiment only: fused multiply-add changes the rounding sequence
  if constexpr (w != 0.0) acc = __fma_rn(w, m[J], acc);
#else
  if constexpr (w != 0.0) acc = acc + w * m[J];
#endif
}
template <int D, int... J>
__device__ __forceinline__ double mrt27Row(const double (&m)[27], std::integer_sequence<int, J...>) {
  double acc = 0.0;
  (mrt27Acc<D, J + 4>(acc, m), ...);
  return acc;
}
template <class St, int... D>
__device__ __forceinline__ void mrt27BackTransform(const double (&g)[27], const double (&m)[27],
                                                   St st, std::integer_sequence<int, D...>) {
  ((st(D, g[D] - mrt27Row<D>(m, std::make_integer_sequence<int, 23>{}))), ...);
}

// ---------------------------------------------------------------------------
template <bool INCOMP, class St>
__device__ __forceinline__ void collide_mrt_d3q27(const double (&g)[27], double rho, double u_x,
                                                  double u_y, double u_z, double omegaKine,
                                                  double omegaBulk, St st) {
  // f(k) of the reference is g[k-1]; written with a 1-based view for legibility
  auto f = [&](int k) -> double { return g[k - 1]; };
  double mneq[27];  // 0-based moment index; entries 0..3 are identically zero (s = 0)

  const double sum_19_22 = f(19) + f(20) + f(21) + f(22);
  const double sum_23_26 = f(23) + f(24) + f(25) + f(26);
  const double sum_19_26 = sum_19_22 + sum_23_26;
  const double sum_21_22_23_24 = f(21) + f(22) + f(23) + f(24);
  const double mom5 = f(15) - f(16) - f(17) + f(18) + sum_19_26 - 2.0 * sum_21_22_23_24;
  const double sum_20_21_24_25 = f(20) + f(21) + f(24) + f(25);
  const double mom6 = f(7) - f(8) - f(9) + f(10) + sum_19_26 - 2.0 * sum_20_21_24_25;
  const double sum_20_22_23_25 = f(20) + f(22) + f(23) + f(25);
  const double mom7 = f(11) - f(12) - f(13) + f(14) + sum_19_26 - 2.0 * sum_20_22_23_25;
  const double sum_7_10 = f(7) + f(8) + f(9) + f(10);
  const double sum_11_14 = f(11) + f(12) + f(13) + f(14);
  const double sum_15_18 = f(15) + f(16) + f(17) + f(18);
  const double sum_11_18 = sum_11_14 + sum_15_18;
  const double mom8 = 2.0 * (f(1) + f(4) - sum_7_10) - f(2) - f(3) - f(5) - f(6) + sum_11_18;
  const double mom9 = f(2) - f(3) + f(5) - f(6) - sum_11_14 + sum_15_18;
  const double mom10 = sum_7_10 + sum_11_18 + 2.0 * (sum_19_26) - f(27);
  const double mom11 = 2.0 * (f(1) - f(4)) - f(11) + f(12) - f(13) + f(14) - f(15) - f(16) +
                       f(17) + f(18) + 4.0 * (-sum_19_22 + sum_23_26);
  const double sum_19_20_23_24 = f(19) + f(20) + f(23) + f(24);
  const double mom12 = 2.0 * (f(2) - f(5)) - f(7) - f(8) + f(9) + f(10) - f(15) + f(16) - f(17) +
                       f(18) + 4.0 * (sum_19_26 - 2.0 * sum_19_20_23_24);
  const double sum_19_21_23_25 = f(19) + f(21) + f(23) + f(25);
  const double mom13 = 2.0 * (f(3) - f(6)) - f(7) + f(8) - f(9) + f(10) - f(11) - f(12) + f(13) +
                       f(14) + 4.0 * (sum_19_26 - 2.0 * sum_19_21_23_25);
  const double mom14 = f(11) - f(12) + f(13) - f(14) - f(15) - f(16) + f(17) + f(18);
  const double mom15 = -f(7) - f(8) + f(9) + f(10) + f(15) - f(16) + f(17) - f(18);
  const double mom16 = f(7) - f(8) + f(9) - f(10) - f(11) - f(12) + f(13) + f(14);
  const double mom17 = -f(19) + f(20) + f(21) - f(22) + f(23) - f(24) - f(25) + f(26);
  const double mom18 = -f(1) - f(2) - f(3) - f(4) - f(5) - f(6) + 4.0 * (sum_19_26) + f(27);
  const double mom19 =
      2.0 * (-f(1) - f(4)) + f(2) + f(3) + f(5) + f(6) - 4.0 * sum_7_10 + 2.0 * (sum_11_18);
  const double mom20 = -f(2) + f(3) - f(5) + f(6) + 2.0 * (-sum_11_14 + sum_15_18);
  const double mom21 = -f(15) + f(16) + f(17) - f(18) + 2.0 * (sum_19_26 - 2.0 * sum_21_22_23_24);
  const double mom22 = -f(7) + f(8) + f(9) - f(10) + 2.0 * (sum_19_26 - 2.0 * sum_20_21_24_25);
  const double mom23 = -f(11) + f(12) + f(13) - f(14) + 2.0 * (sum_19_26 - 2.0 * sum_20_22_23_25);
  const double mom24 = -f(1) + f(4) +
                       2.0 * (f(11) - f(12) + f(13) - f(14) + f(15) + f(16) - f(17) - f(18)) +
                       4.0 * (-sum_19_22 + sum_23_26);
  const double mom25 = -f(2) + f(5) +
                       2.0 * (f(7) + f(8) - f(9) - f(10) + f(15) - f(16) + f(17) - f(18)) +
                       4.0 * (sum_19_26 - 2.0 * sum_19_20_23_24);
  const double mom26 = -f(3) + f(6) +
                       2.0 * (f(7) - f(8) + f(9) - f(10) + f(11) + f(12) - f(13) - f(14)) +
                       4.0 * (sum_19_26 - 2.0 * sum_19_21_23_25);
  const double mom27 = 2.0 * (f(1) + f(2) + f(3) + f(4) + f(5) + f(6)) +
                       4.0 * (-sum_7_10 - sum_11_18) + 8.0 * (sum_19_26) - f(27);

  // incompressible (mus_compute_mrt_d3q27_module.fpp:516-526): rho0 = 1 in meq(2:10)
  const double rq = INCOMP ? 1.0 : rho;
  const double meq2 = rq * u_x, meq3 = rq * u_y, meq4 = rq * u_z;
  const double meq5 = meq2 * u_y, meq6 = meq3 * u_z, meq7 = meq4 * u_x;
  const double meq8 = rq * (2.0 * u_x * u_x - u_y * u_y - u_z * u_z);
  const double meq9 = rq * (u_y * u_y - u_z * u_z);
  const double meq10 = rq * (u_x * u_x + u_y * u_y + u_z * u_z);

  // s_mrt of mrt_d3q27 (mus_mrtRelaxation_module.fpp:267-288), s(5:9) = omegaKine
  mneq[0] = 0.0; mneq[1] = 0.0; mneq[2] = 0.0; mneq[3] = 0.0;
  mneq[4] = omegaKine * (mom5 - meq5);
  mneq[5] = omegaKine * (mom6 - meq6);
  mneq[6] = omegaKine * (mom7 - meq7);
  mneq[7] = omegaKine * (mom8 - meq8);
  mneq[8] = omegaKine * (mom9 - meq9);
  mneq[9] = omegaBulk * (mom10 - meq10);
  mneq[10] = 1.50 * (mom11 - 0.0);
  mneq[11] = 1.50 * (mom12 - 0.0);
  mneq[12] = 1.50 * (mom13 - 0.0);
  mneq[13] = 1.74 * (mom14 - 0.0);
  mneq[14] = 1.74 * (mom15 - 0.0);
  mneq[15] = 1.74 * (mom16 - 0.0);
  mneq[16] = 1.74 * (mom17 - 0.0);
  mneq[17] = 1.4 * (mom18 - 0.0);
  mneq[18] = 1.98 * (mom19 - 0.0);
  mneq[19] = 1.98 * (mom20 - 0.0);
  mneq[20] = 1.98 * (mom21 - 0.0);
  mneq[21] = 1.98 * (mom22 - 0.0);
  mneq[22] = 1.98 * (mom23 - 0.0);
  mneq[23] = 1.83 * (mom24 - 0.0);
  mneq[24] = 1.83 * (mom25 - 0.0);
  mneq[25] = 1.83 * (mom26 - 0.0);
  mneq[26] = 1.61 * (mom27 - 0.0);

  // f(iDir) - sum( WMMIvD3Q27(iDir,:) * mneq(:) ): the zero entries of the matrix and
  // the four zero moments drop out without changing any rounding (x + 0*y = x).
  mrt27BackTransform(g, mneq, st, std::make_integer_sequence<int, 27>{});
}

}  // namespace musb200
