/* kernels.c -- ORACLE (test infrastructure): the stream+collide kernels of the
 * reference restated in C with the same operation order (AOS + PULL).
 *
 *   BGK D3Q19  mus/source/compute/mus_compute_d3q19_module.fpp:483-640
 *   TRT D3Q19  mus/source/compute/mus_compute_d3q19_module.fpp:2644-2763
 *   fluid_incompressible: BGK D3Q19 :1215-1705, TRT D3Q19 :2782-2959,
 *              MRT D3Q19 mus_compute_mrt_d3q19_module.fpp:466-739,
 *              MRT D3Q27 mus_compute_mrt_d3q27_module.fpp:374-546,
 *              BGK D3Q27 = the kCFD kernel with get_pdfEq_incomp_d3q27
 *   MRT D3Q19  mus/source/compute/mus_compute_mrt_d3q19_module.fpp:238-450
 *   BGK D3Q27  mus/source/compute/mus_compute_d3q27_module.fpp:398-517
 *   TRT D3Q27  mus/source/compute/mus_compute_d3q27_module.fpp:601-740
 *   MRT D3Q27  mus/source/compute/mus_compute_mrt_d3q27_module.fpp:255-361
 *   NoOpt BGK  mus/source/compute/mus_compute_bgk_module.fpp:126-157
 *   NoOpt MRT  mus/source/compute/mus_compute_mrt_d3q19_module.fpp:1054-1097
 *   s_mrt      mus/source/mus_mrtRelaxation_module.fpp:238-288
 *
 * Build with -ffp-contract=off so that every product and sum is rounded as
 * written (the reference itself is only reproducible to FMA-contraction noise).
 */
#include "mus_oracle.h"
#include "mrt_tables.h"
#include <stddef.h>

/* 1-based direction names (mus/source/mus_directions_module.f90:10-35) */
enum {
  qN00 = 1, q0N0, q00N, q100, q010, q001, q0NN, q0N1, q01N, q011,
  qN0N, q10N, qN01, q101, qNN0, qN10, q1N0, q110,
  qNNN, qNN1, qN1N, qN11, q1NN, q1N1, q11N, q111
};

#define PULL(d) in[neigh[(size_t)((d) - 1) * nSize + (e - 1)] - 1]   /* FETCH */
#define SAVE(d) out[(size_t)(e - 1) * QQ + (d) - 1]                   /* IDX   */
#define AUX(k)  aux[(size_t)(e - 1) * 4 + (k)]

void ora_mrt_diag(int QQ, double omegaKine, double omegaBulk, double *s /*0-based*/) {
  for (int i = 0; i < QQ; ++i) s[i] = 0.0;
  double *m = s - 1;
  if (QQ == 19) {
    m[2] = omegaBulk; m[3] = 1.40;
    m[5] = 1.20; m[7] = 1.20; m[9] = 1.20;
    m[10] = omegaKine; m[11] = 1.40; m[12] = omegaKine; m[13] = 1.40;
    m[14] = omegaKine; m[15] = omegaKine; m[16] = omegaKine;
    m[17] = 1.98; m[18] = 1.98; m[19] = 1.98;
  } else {
    for (int i = 5; i <= 9; ++i) m[i] = omegaKine;
    m[10] = omegaBulk;
    for (int i = 11; i <= 13; ++i) m[i] = 1.50;
    for (int i = 14; i <= 17; ++i) m[i] = 1.74;
    m[18] = 1.4;
    for (int i = 19; i <= 23; ++i) m[i] = 1.98;
    for (int i = 24; i <= 26; ++i) m[i] = 1.83;
    m[27] = 1.61;
  }
}

/* MMtrD3Q19 / MMivD3Q19 / WMMtrD3Q27 / WMMIvD3Q27 (mus_mrtInit_module.f90:66-301), row-major */
const double *ora_mrt_matrix(int QQ, int inverse) {
  if (QQ == 19) return inverse ? &ORA_MMivD3Q19[0][0] : &ORA_MMtrD3Q19[0][0];
  if (QQ == 27) return inverse ? &ORA_WMMIvD3Q27[0][0] : &ORA_WMMtrD3Q27[0][0];
  return 0;
}

/* ======================================================================== */
static void bgk_d3q19(int incomp, const double *in, double *out, const double *aux,
                      const int32_t *neigh, const double *omg, int nSize, int nSolve) {
  const int QQ = 19;
  const double div1_3 = 1.0 / 3.0, div1_8 = 1.0 / 8.0, div1_36 = 1.0 / 36.0;
  const double div3_4h = 3.0 / 4.5;
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    const double fN00 = PULL(qN00), f0N0 = PULL(q0N0), f00N = PULL(q00N);
    const double f100 = PULL(q100), f010 = PULL(q010), f001 = PULL(q001);
    const double f0NN = PULL(q0NN), f0N1 = PULL(q0N1), f01N = PULL(q01N), f011 = PULL(q011);
    const double fN0N = PULL(qN0N), f10N = PULL(q10N), fN01 = PULL(qN01), f101 = PULL(q101);
    const double fNN0 = PULL(qNN0), fN10 = PULL(qN10), f1N0 = PULL(q1N0), f110 = PULL(q110);
    const double f000 = PULL(19);

    const double rho = AUX(0), u_x = AUX(1), u_y = AUX(2), u_z = AUX(3);
    const double usq = (u_x * u_x) + (u_y * u_y) + (u_z * u_z);
    const double omega = omg[e - 1];
    const double cmpl_o = 1.0 - omega;
    double coeff_1, coeff_2, usqn_o1, usqn_o2;
    if (!incomp) {
      const double usqn = div1_36 * (1.0 - 1.5 * usq) * rho;
      SAVE(19) = f000 * cmpl_o + omega * rho * (div1_3 - 0.5 * usq);
      coeff_1 = div1_8 * omega * rho;
      usqn_o1 = omega * usqn;
      const double omega_2 = 2.0 * omega;
      coeff_2 = div1_8 * omega_2 * rho;
      usqn_o2 = omega_2 * usqn;
    } else {
      /* mus_advRel_kFluidIncomp_rBGK_vStd_lD3Q19, mus_compute_d3q19_module.fpp:1596-1672:
       * rho0 = 1 replaces rho in every momentum term                          */
      usqn_o1 = omega * div1_36 * (rho - 1.5 * usq);
      SAVE(19) = f000 * cmpl_o + 12.0 * usqn_o1;
      coeff_1 = div1_8 * omega;
      coeff_2 = div1_8 * omega * 2.0;
      usqn_o2 = 2.0 * usqn_o1;
    }

    double ui, fac, s1, s2;
#define PAIR(plus, minus, fp, fm)                 \
    fac = coeff_1 * ui; s1 = fac * div3_4h; s2 = fac * ui + usqn_o1; \
    SAVE(plus) = fp * cmpl_o + s1 + s2;           \
    SAVE(minus) = fm * cmpl_o - s1 + s2;
    ui = u_x + u_y;  PAIR(q110, qNN0, f110, fNN0)
    ui = -u_x + u_y; PAIR(qN10, q1N0, fN10, f1N0)
    ui = u_x + u_z;  PAIR(q101, qN0N, f101, fN0N)
    ui = -u_x + u_z; PAIR(qN01, q10N, fN01, f10N)
    ui = u_y + u_z;  PAIR(q011, q0NN, f011, f0NN)
    ui = -u_y + u_z; PAIR(q0N1, q01N, f0N1, f01N)
#undef PAIR
#define AXIS(plus, minus, fp, fm, u)              \
    fac = coeff_2 * u; s1 = fac * div3_4h; s2 = fac * u + usqn_o2; \
    SAVE(plus) = fp * cmpl_o + s1 + s2;           \
    SAVE(minus) = fm * cmpl_o - s1 + s2;
    AXIS(q010, q0N0, f010, f0N0, u_y)
    AXIS(q100, qN00, f100, fN00, u_x)
    AXIS(q001, q00N, f001, f00N, u_z)
#undef AXIS
  }
}

/* ======================================================================== */
static void trt_d3q19(const double *in, double *out, const double *aux,
                      const int32_t *neigh, const double *omg, int nSize, int nSolve,
                      double lambda) {
  const int QQ = 19;
  const double div1_3 = 1.0 / 3.0, t2cs4inv = 4.5;
  const double t1x2_0 = 1.0 / 18.0 * 2.0, t2x2_0 = 1.0 / 36.0 * 2.0;
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    const double fN00 = PULL(qN00), f0N0 = PULL(q0N0), f00N = PULL(q00N);
    const double f100 = PULL(q100), f010 = PULL(q010), f001 = PULL(q001);
    const double f0NN = PULL(q0NN), f0N1 = PULL(q0N1), f01N = PULL(q01N), f011 = PULL(q011);
    const double fN0N = PULL(qN0N), f10N = PULL(q10N), fN01 = PULL(qN01), f101 = PULL(q101);
    const double fNN0 = PULL(qNN0), fN10 = PULL(qN10), f1N0 = PULL(q1N0), f110 = PULL(q110);
    const double f000 = PULL(19);

    const double rho = AUX(0), u_x = AUX(1), u_y = AUX(2), u_z = AUX(3);
    const double usq = (u_x * u_x) + (u_y * u_y) + (u_z * u_z);
    const double feq_common = 1.0 - 1.5 * usq;
    const double omega = omg[e - 1];
    const double omega_h = 0.5 * omega;
    const double asym_omega = 1.0 / (0.5 + lambda / (1.0 / omega - 0.5));
    const double asym_omega_h = 0.5 * asym_omega;

    SAVE(19) = f000 * (1.0 - omega) + omega * div1_3 * rho * feq_common;

    double ui, sym, asym;
    const double t2x2 = t2x2_0 * rho;
    const double fac2 = t2x2 * t2cs4inv;
#define LINK(tx2, fc, plus, minus, fp, fm)                                  \
    sym = omega_h * (fp + fm - fc * ui * ui - tx2 * feq_common);             \
    asym = asym_omega_h * (fp - fm - 3.0 * tx2 * ui);                        \
    SAVE(plus) = fp - sym - asym;                                            \
    SAVE(minus) = fm - sym + asym;
    ui = u_x + u_y; LINK(t2x2, fac2, q110, qNN0, f110, fNN0)
    ui = u_x - u_y; LINK(t2x2, fac2, q1N0, qN10, f1N0, fN10)
    ui = u_x + u_z; LINK(t2x2, fac2, q101, qN0N, f101, fN0N)
    ui = u_x - u_z; LINK(t2x2, fac2, q10N, qN01, f10N, fN01)
    ui = u_y + u_z; LINK(t2x2, fac2, q011, q0NN, f011, f0NN)
    ui = u_y - u_z; LINK(t2x2, fac2, q01N, q0N1, f01N, f0N1)
    const double t1x2 = t1x2_0 * rho;
    const double fac1 = t1x2 * t2cs4inv;
    ui = u_x; LINK(t1x2, fac1, q100, qN00, f100, fN00)
    ui = u_y; LINK(t1x2, fac1, q010, q0N0, f010, f0N0)
    ui = u_z; LINK(t1x2, fac1, q001, q00N, f001, f00N)
#undef LINK
  }
}

/* ======================================================================== */
/* mus_advRel_kFluidIncomp_rTRT_vStd_lD3Q19, mus_compute_d3q19_module.fpp:2782-2959 */
static void trt_d3q19_incomp(const double *in, double *out, const double *aux,
                             const int32_t *neigh, const double *omg, int nSize, int nSolve,
                             double lambda) {
  const int QQ = 19;
  const double div1_3 = 1.0 / 3.0, div1_6 = 1.0 / 6.0, t2cs4inv = 4.5;
  const double t1x2 = 1.0 / 9.0, t2x2 = 1.0 / 18.0;
  const double fac1 = t1x2 * t2cs4inv, fac2 = t2x2 * t2cs4inv;
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    const double fN00 = PULL(qN00), f0N0 = PULL(q0N0), f00N = PULL(q00N);
    const double f100 = PULL(q100), f010 = PULL(q010), f001 = PULL(q001);
    const double f0NN = PULL(q0NN), f0N1 = PULL(q0N1), f01N = PULL(q01N), f011 = PULL(q011);
    const double fN0N = PULL(qN0N), f10N = PULL(q10N), fN01 = PULL(qN01), f101 = PULL(q101);
    const double fNN0 = PULL(qNN0), fN10 = PULL(qN10), f1N0 = PULL(q1N0), f110 = PULL(q110);
    const double f000 = PULL(19);

    const double rho = AUX(0), u_x = AUX(1), u_y = AUX(2), u_z = AUX(3);
    const double usq = (u_x * u_x) + (u_y * u_y) + (u_z * u_z);
    const double feq_common = rho - 1.5 * usq;
    const double omega = omg[e - 1];
    const double omega_h = 0.5 * omega;
    const double asym_omega = 1.0 / (0.5 + lambda / (1.0 / omega - 0.5));
    const double asym_omega_h = 0.5 * asym_omega;

    SAVE(19) = f000 * (1.0 - omega) + omega * div1_3 * feq_common;

    double ui, sym, asym;
    const double t2_feq = t2x2 * feq_common;
#define LINK(fc, tfeq, dv, plus, minus, fp, fm)                 \
    sym = omega_h * (fp + fm - fc * ui * ui - tfeq);            \
    asym = asym_omega_h * (fp - fm - dv * ui);                  \
    SAVE(plus) = fp - sym - asym;                               \
    SAVE(minus) = fm - sym + asym;
    ui = u_x + u_y; LINK(fac2, t2_feq, div1_6, q110, qNN0, f110, fNN0)
    ui = u_x - u_y; LINK(fac2, t2_feq, div1_6, q1N0, qN10, f1N0, fN10)
    ui = u_x + u_z; LINK(fac2, t2_feq, div1_6, q101, qN0N, f101, fN0N)
    ui = u_x - u_z; LINK(fac2, t2_feq, div1_6, q10N, qN01, f10N, fN01)
    ui = u_y + u_z; LINK(fac2, t2_feq, div1_6, q011, q0NN, f011, f0NN)
    ui = u_y - u_z; LINK(fac2, t2_feq, div1_6, q01N, q0N1, f01N, f0N1)
    const double t1_feq = t1x2 * feq_common;
    ui = u_y; LINK(fac1, t1_feq, div1_3, q010, q0N0, f010, f0N0)
    ui = u_x; LINK(fac1, t1_feq, div1_3, q100, qN00, f100, fN00)
    ui = u_z; LINK(fac1, t1_feq, div1_3, q001, q00N, f001, f00N)
#undef LINK
  }
}

/* ======================================================================== */
static void mrt_d3q19(int incomp, const double *in, double *out, const double *aux,
                      const int32_t *neigh, const double *omg, int nSize, int nSolve,
                      double omegaBulk) {
  const int QQ = 19;
  const double div1_4 = 1.0 / 4.0, div1_8 = 1.0 / 8.0, div1_12 = 1.0 / 12.0;
  const double div1_16 = 1.0 / 16.0, div1_24 = 1.0 / 24.0, div1_48 = 1.0 / 48.0;
  const double div1_72 = 1.0 / 72.0;
  double s0[19];
  ora_mrt_diag(19, 1.0, omegaBulk, s0);
  double *sc = s0 - 1; /* 1-based */
  sc[2] *= div1_24; sc[3] *= div1_72; sc[5] *= div1_24; sc[7] *= div1_24; sc[9] *= div1_24;
  sc[17] *= div1_8; sc[18] *= div1_8; sc[19] *= div1_8;
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    double sl[19];
    for (int i = 0; i < 19; ++i) sl[i] = s0[i];
    double *s = sl - 1;
    const double fN00 = PULL(qN00), f0N0 = PULL(q0N0), f00N = PULL(q00N);
    const double f100 = PULL(q100), f010 = PULL(q010), f001 = PULL(q001);
    const double f0NN = PULL(q0NN), f0N1 = PULL(q0N1), f01N = PULL(q01N), f011 = PULL(q011);
    const double fN0N = PULL(qN0N), f10N = PULL(q10N), fN01 = PULL(qN01), f101 = PULL(q101);
    const double fNN0 = PULL(qNN0), fN10 = PULL(qN10), f1N0 = PULL(q1N0), f110 = PULL(q110);
    const double f000 = PULL(19);

    const double m6 = f101 + fN0N + f10N + fN01;
    const double m8 = f011 + f0NN + f01N + f0N1;
    const double sum1 = f110 + fNN0 + f1N0 + fN10;
    const double m2 = -f000 + sum1 + m6 + m8;
    const double sum2 = f010 + f0N0;
    const double sum3 = f001 + f00N;
    const double sum4 = 2.0 * (f100 + fN00);
    const double sum5 = sum2 + sum3;
    const double mout3 = (2.0 * (f000 - sum5) - sum4 + m2) * s[3];

    const double rho = AUX(0), u_x = AUX(1), u_y = AUX(2), u_z = AUX(3);
    const double omegaKine = omg[e - 1];
    s[10] = omegaKine; s[12] = omegaKine;
    s[14] = div1_4 * omegaKine; s[15] = div1_4 * omegaKine; s[16] = div1_4 * omegaKine;

    /* incompressible (mus_compute_mrt_d3q19_module.fpp:578-603): rho0 = 1 replaces rho */
    const double meq2 = incomp ? u_x * u_x + u_y * u_y + u_z * u_z
                               : rho * (u_x * u_x + u_y * u_y + u_z * u_z);
    const double meq10 = incomp ? 3.0 * u_x * u_x - meq2 : rho * 3.0 * u_x * u_x - meq2;
    const double meq12 = incomp ? u_y * u_y - u_z * u_z : rho * (u_y * u_y - u_z * u_z);
    const double mout2 = s[2] * (m2 - meq2);
    const double m14 = f110 + fNN0 - f1N0 - fN10;
    const double mout14 = s[14] * (m14 - (incomp ? u_x * u_y : rho * u_x * u_y));
    const double m15 = f011 + f0NN - f01N - f0N1;
    const double mout15 = s[15] * (m15 - (incomp ? u_y * u_z : rho * u_y * u_z));
    const double m16 = f101 + fN0N - f10N - fN01;
    const double mout16 = s[16] * (m16 - (incomp ? u_x * u_z : rho * u_x * u_z));

    const double sum6 = sum1 + m6 - m8 * 2.0;
    const double sum7 = sum4 - sum5;
    const double mout10 = (sum7 + sum6 - meq10) * s[10];
    const double mout11 = (-sum7 + sum6) * s[11];
    const double sum8 = sum1 - m6;
    const double sum9 = sum2 - sum3;
    const double mout12 = (sum8 + sum9 - meq12) * s[12];
    const double mout13 = (sum8 - sum9) * s[13];

    double c1 = f110 - fNN0, c2 = f1N0 - fN10;
    double c3 = f101 - fN0N, c4 = f10N - fN01;
    const double sum10 = c1 + c2, sum11 = c3 + c4;
    const double mout5 = (sum10 + sum11 - 2.0 * (f100 - fN00)) * s[5];
    const double mout17 = (sum10 - sum11) * s[17];
    double c5 = f011 - f0NN, c6 = f01N - f0N1;
    const double sum12 = c1 - c2, sum13 = c5 + c6;
    const double mout7 = (sum12 + sum13 - 2.0 * (f010 - f0N0)) * s[7];
    const double mout18 = (-sum12 + sum13) * s[18];
    const double sum14 = c3 - c4, sum15 = c5 - c6;
    const double mout9 = (sum14 + sum15 - 2.0 * (f001 - f00N)) * s[9];
    const double mout19 = (sum14 - sum15) * s[19];

    SAVE(19) = f000 + 12.0 * (mout2 - mout3);

    const double c0 = -4.0 * mout3 + div1_12 * (mout10 - mout11);
    const double mout5_4 = mout5 * 4.0;
    SAVE(4) = f100 - (c0 - mout5_4);
    SAVE(1) = fN00 - (c0 + mout5_4);

    c1 = -4.0 * mout3 - div1_24 * (mout10 - mout11);
    c2 = div1_8 * (mout12 - mout13);
    const double sum_c1_c2 = c1 + c2;
    const double mout7_4 = mout7 * 4.0;
    SAVE(5) = f010 - (sum_c1_c2 - mout7_4);
    SAVE(2) = f0N0 - (sum_c1_c2 + mout7_4);
    const double sub_c1_c2 = c1 - c2;
    const double mout9_4 = mout9 * 4.0;
    SAVE(6) = f001 - (sub_c1_c2 - mout9_4);
    SAVE(3) = f00N - (sub_c1_c2 + mout9_4);

    const double mout1 = mout2 + mout3;
    c3 = mout1 + div1_48 * (mout10 + mout11) + div1_16 * (mout12 + mout13);
    const double sum_5_17 = mout5 + mout17, sub_7_18 = mout7 - mout18;
    const double d1 = c3 + mout14, d2 = sum_5_17 + sub_7_18;
    SAVE(18) = f110 - (d1 + d2);
    SAVE(15) = fNN0 - (d1 - d2);
    const double d3 = c3 - mout14, d4 = sum_5_17 - sub_7_18;
    SAVE(17) = f1N0 - (d3 + d4);
    SAVE(16) = fN10 - (d3 - d4);

    c4 = c3 - div1_8 * (mout12 + mout13);
    const double sum_9_19 = mout9 + mout19, sub_5_17 = mout5 - mout17;
    const double e1 = c4 + mout16, e2 = sum_9_19 + sub_5_17;
    SAVE(14) = f101 - (e1 + e2);
    SAVE(11) = fN0N - (e1 - e2);
    const double e3 = c4 - mout16, e4 = sum_9_19 - sub_5_17;
    SAVE(12) = f10N - (e3 - e4);
    SAVE(13) = fN01 - (e3 + e4);

    c5 = mout1 - div1_24 * (mout10 + mout11);
    const double sum_7_18 = mout7 + mout18, sub_9_19 = mout9 - mout19;
    const double g1 = c5 + mout15, g2 = sum_7_18 + sub_9_19;
    SAVE(10) = f011 - (g1 + g2);
    SAVE(7) = f0NN - (g1 - g2);
    const double g3 = c5 - mout15, g4 = sum_7_18 - sub_9_19;
    SAVE(9) = f01N - (g3 + g4);
    SAVE(8) = f0N1 - (g3 - g4);
    (void)c6;
  }
}

/* ======================================================================== */
static void bgk_generic(int QQ, int incomp, const double *in, double *out, const double *aux,
                        const int32_t *neigh, const double *omg, int nSize, int nSolve) {
  /* D3Q27 BGK (mus_compute_d3q27_module.fpp:398-517) and the NoOpt BGK
   * (mus_compute_bgk_module.fpp:126-157) are the same arithmetic:
   * out = f - omega*(f - fEq) with fEq = pdfEq_ptr(rho, vel); for
   * fluid_incompressible the pointer is get_pdfEq_incomp_d3q27
   * (mus_scheme_derived_quantities_type_module.f90:751-811).                  */
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    double f[27], fEq[27];
    for (int d = 1; d <= QQ; ++d) f[d - 1] = PULL(d);
    const double rho = AUX(0);
    const double vel[3] = {AUX(1), AUX(2), AUX(3)};
    if (incomp) ora_pdfEq_incomp(QQ, rho, vel, fEq); else ora_pdfEq(QQ, rho, vel, fEq);
    const double omega = omg[e - 1];
    for (int d = 1; d <= QQ; ++d) SAVE(d) = f[d - 1] - omega * (f[d - 1] - fEq[d - 1]);
  }
}

/* ======================================================================== */
static void trt_d3q27(const double *in, double *out, const double *aux,
                      const int32_t *neigh, const double *omg, int nSize, int nSolve,
                      double lambda) {
  const int QQ = 27;
  const double div2_3 = 2.0 / 3.0, div1_2 = 1.0 / 2.0;
  /* the 13 (+c, -c) pairs in the reference's order (:656-740) */
  static const int pairs[13][2] = {
    {q100, qN00}, {q010, q0N0}, {q001, q00N}, {q011, q0NN}, {q01N, q0N1},
    {q101, qN0N}, {q10N, qN01}, {q110, qNN0}, {q1N0, qN10}, {q1NN, qN11},
    {q11N, qNN1}, {q1N1, qN1N}, {q111, qNNN}};
  const int *cx = ora_cxDir(27);
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    double f[28];
    for (int d = 1; d <= QQ; ++d) f[d] = PULL(d);
    const double rho = AUX(0), u = AUX(1), v = AUX(2), w = AUX(3);
    const double u2 = u * u, v2 = v * v, w2 = w * w;
    /* X[c+1]: c=-1 -> XN, 0 -> X0, +1 -> X1 (eq. A.19-A.21 of the cited paper) */
    double X[3], Y[3], Z[3];
    X[1] = -div2_3 + u2; X[2] = -(X[1] + 1.0 + u) * 0.5; X[0] = X[2] + u;
    Y[1] = -div2_3 + v2; Y[2] = -(Y[1] + 1.0 + v) * 0.5; Y[0] = Y[2] + v;
    Z[1] = -div2_3 + w2; Z[2] = -(Z[1] + 1.0 + w) * 0.5; Z[0] = Z[2] + w;
    const double wP = omg[e - 1];
    const double wN = 1.0 / (0.5 + lambda / (1.0 / wP - 0.5));

    SAVE(27) = (1.0 - wP) * f[27] - rho * wP * X[1] * Y[1] * Z[1];
    for (int k = 0; k < 13; ++k) {
      const int dp = pairs[k][0], dm = pairs[k][1];
      const int *c = cx + 3 * (dp - 1);
      const double Xp = X[c[0] + 1], Yp = Y[c[1] + 1], Zp = Z[c[2] + 1];
      const double Xm = X[-c[0] + 1], Ym = Y[-c[1] + 1], Zm = Z[-c[2] + 1];
      const double p_part = wP * ((f[dp] + f[dm]) - (-rho * Xp * Yp * Zp - rho * Xm * Ym * Zm)) * div1_2;
      const double n_part = wN * ((f[dp] - f[dm]) - (-rho * Xp * Yp * Zp + rho * Xm * Ym * Zm)) * div1_2;
      SAVE(dp) = f[dp] - p_part - n_part;
      SAVE(dm) = f[dm] - p_part + n_part;
    }
  }
}

/* ======================================================================== */
static void mrt_d3q27(int incomp, const double *in, double *out, const double *aux,
                      const int32_t *neigh, const double *omg, int nSize, int nSolve,
                      double omegaBulk) {
  const int QQ = 27;
  double s0[27];
  ora_mrt_diag(27, 1.0, omegaBulk, s0);
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    double f[28], mom[28], meq[28], mneq[28], s[28];
    for (int d = 1; d <= QQ; ++d) { f[d] = PULL(d); s[d] = s0[d - 1]; meq[d] = 0.0; }
    const double rho = AUX(0), u_x = AUX(1), u_y = AUX(2), u_z = AUX(3);

    mom[1] = rho;
    mom[2] = rho * u_x; mom[3] = rho * u_y; mom[4] = rho * u_z;
    const double sum_19_22 = f[19] + f[20] + f[21] + f[22];
    const double sum_23_26 = f[23] + f[24] + f[25] + f[26];
    const double sum_19_26 = sum_19_22 + sum_23_26;
    const double sum_21_22_23_24 = f[21] + f[22] + f[23] + f[24];
    mom[5] = f[15] - f[16] - f[17] + f[18] + sum_19_26 - 2.0 * sum_21_22_23_24;
    const double sum_20_21_24_25 = f[20] + f[21] + f[24] + f[25];
    mom[6] = f[7] - f[8] - f[9] + f[10] + sum_19_26 - 2.0 * sum_20_21_24_25;
    const double sum_20_22_23_25 = f[20] + f[22] + f[23] + f[25];
    mom[7] = f[11] - f[12] - f[13] + f[14] + sum_19_26 - 2.0 * sum_20_22_23_25;
    const double sum_7_10 = f[7] + f[8] + f[9] + f[10];
    const double sum_11_14 = f[11] + f[12] + f[13] + f[14];
    const double sum_15_18 = f[15] + f[16] + f[17] + f[18];
    const double sum_11_18 = sum_11_14 + sum_15_18;
    mom[8] = 2.0 * (f[1] + f[4] - sum_7_10) - f[2] - f[3] - f[5] - f[6] + sum_11_18;
    mom[9] = f[2] - f[3] + f[5] - f[6] - sum_11_14 + sum_15_18;
    mom[10] = sum_7_10 + sum_11_18 + 2.0 * (sum_19_26) - f[27];
    mom[11] = 2.0 * (f[1] - f[4]) - f[11] + f[12] - f[13] + f[14] - f[15] - f[16] + f[17] + f[18]
            + 4.0 * (-sum_19_22 + sum_23_26);
    const double sum_19_20_23_24 = f[19] + f[20] + f[23] + f[24];
    mom[12] = 2.0 * (f[2] - f[5]) - f[7] - f[8] + f[9] + f[10] - f[15] + f[16] - f[17] + f[18]
            + 4.0 * (sum_19_26 - 2.0 * sum_19_20_23_24);
    const double sum_19_21_23_25 = f[19] + f[21] + f[23] + f[25];
    mom[13] = 2.0 * (f[3] - f[6]) - f[7] + f[8] - f[9] + f[10] - f[11] - f[12] + f[13] + f[14]
            + 4.0 * (sum_19_26 - 2.0 * sum_19_21_23_25);
    mom[14] = f[11] - f[12] + f[13] - f[14] - f[15] - f[16] + f[17] + f[18];
    mom[15] = -f[7] - f[8] + f[9] + f[10] + f[15] - f[16] + f[17] - f[18];
    mom[16] = f[7] - f[8] + f[9] - f[10] - f[11] - f[12] + f[13] + f[14];
    mom[17] = -f[19] + f[20] + f[21] - f[22] + f[23] - f[24] - f[25] + f[26];
    mom[18] = -f[1] - f[2] - f[3] - f[4] - f[5] - f[6] + 4.0 * (sum_19_26) + f[27];
    mom[19] = 2.0 * (-f[1] - f[4]) + f[2] + f[3] + f[5] + f[6] - 4.0 * sum_7_10 + 2.0 * (sum_11_18);
    mom[20] = -f[2] + f[3] - f[5] + f[6] + 2.0 * (-sum_11_14 + sum_15_18);
    mom[21] = -f[15] + f[16] + f[17] - f[18] + 2.0 * (sum_19_26 - 2.0 * sum_21_22_23_24);
    mom[22] = -f[7] + f[8] + f[9] - f[10] + 2.0 * (sum_19_26 - 2.0 * sum_20_21_24_25);
    mom[23] = -f[11] + f[12] + f[13] - f[14] + 2.0 * (sum_19_26 - 2.0 * sum_20_22_23_25);
    mom[24] = -f[1] + f[4] + 2.0 * (f[11] - f[12] + f[13] - f[14] + f[15] + f[16] - f[17] - f[18])
            + 4.0 * (-sum_19_22 + sum_23_26);
    mom[25] = -f[2] + f[5] + 2.0 * (f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18])
            + 4.0 * (sum_19_26 - 2.0 * sum_19_20_23_24);
    mom[26] = -f[3] + f[6] + 2.0 * (f[7] - f[8] + f[9] - f[10] + f[11] + f[12] - f[13] - f[14])
            + 4.0 * (sum_19_26 - 2.0 * sum_19_21_23_25);
    mom[27] = 2.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6]) + 4.0 * (-sum_7_10 - sum_11_18)
            + 8.0 * (sum_19_26) - f[27];

    /* incompressible (mus_compute_mrt_d3q27_module.fpp:516-526): rho0 = 1 in meq(2:10) */
    const double rq = incomp ? 1.0 : rho;
    meq[1] = rho;
    meq[2] = rq * u_x; meq[3] = rq * u_y; meq[4] = rq * u_z;
    meq[5] = meq[2] * u_y;
    meq[6] = meq[3] * u_z;
    meq[7] = meq[4] * u_x;
    meq[8] = rq * (2.0 * u_x * u_x - u_y * u_y - u_z * u_z);
    meq[9] = rq * (u_y * u_y - u_z * u_z);
    meq[10] = rq * (u_x * u_x + u_y * u_y + u_z * u_z);

    for (int i = 5; i <= 9; ++i) s[i] = omg[e - 1];
    for (int i = 1; i <= QQ; ++i) mneq[i] = s[i] * (mom[i] - meq[i]);
    for (int d = 1; d <= QQ; ++d) {
      double acc = 0.0; /* Fortran sum(W(iDir,:)*mneq(:)) */
      for (int j = 1; j <= QQ; ++j) acc = acc + ORA_WMMIvD3Q27[d - 1][j - 1] * mneq[j];
      f[d] = f[d] - acc;
    }
    for (int d = 1; d <= QQ; ++d) SAVE(d) = f[d];
  }
}

/* ======================================================================== */
static void mrt_noopt(int QQ, int incomp, const double *in, double *out, const double *aux,
                      const int32_t *neigh, const double *omg, int nSize, int nSolve,
                      double omegaBulk) {
  /* M^-1 S M (f - fEq); the D3Q27 NoOpt variant (mrt_d3q27:88-185) has the
   * same structure with the weighted matrices.                               */
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    double f[27], fEq[27], fneq[27], mneq[27], s[27];
    for (int d = 1; d <= QQ; ++d) f[d - 1] = PULL(d);
    const double rho = AUX(0);
    const double vel[3] = {AUX(1), AUX(2), AUX(3)};
    if (incomp) ora_pdfEq_incomp(QQ, rho, vel, fEq); else ora_pdfEq(QQ, rho, vel, fEq);
    for (int d = 0; d < QQ; ++d) fneq[d] = f[d] - fEq[d];
    for (int i = 0; i < QQ; ++i) {
      double acc = 0.0;
      for (int j = 0; j < QQ; ++j)
        acc += (QQ == 19 ? ORA_MMtrD3Q19[i][j] : ORA_WMMtrD3Q27[i][j]) * fneq[j];
      mneq[i] = acc;
    }
    ora_mrt_diag(QQ, omg[e - 1], omegaBulk, s);
    for (int d = 0; d < QQ; ++d) {
      double acc = 0.0;
      for (int j = 0; j < QQ; ++j)
        acc += ((QQ == 19 ? ORA_MMivD3Q19[d][j] : ORA_WMMIvD3Q27[d][j]) * s[j]) * mneq[j];
      out[(size_t)(e - 1) * QQ + d] = f[d] - acc;
    }
  }
}

/* ======================================================================== */
/* Generic formulations: second implementations of BGK and TRT written from the textbook
 * definitions, NOT from the reference's optimised kernels -- table-driven equilibria and an
 * explicit symmetric / antisymmetric split.  They are the comparison partners of the optimised
 * restatements above where the reference's own utests offer none (TRT D3Q19 / D3Q27, BGK D3Q27),
 * at the utests' tolerance of 2500 eps (tests/test_oracle_properties.py).
 *   feq_kind 0: f_i = w_i rho (1 + 3 c.u + 9/2 (c.u)^2 - 3/2 u^2)     (get_pdfEq_d3q19 / _d3q27)
 *   feq_kind 1: f_i = rho Phi_cx(ux) Phi_cy(uy) Phi_cz(uz), Phi_0(a) = 2/3 - a^2,
 *               Phi_+-1(a) = (1/3 + a^2 +- a) / 2                      (the product form of
 *               mus_advRel_kFluid_rTRT_vStd_lD3Q27, mus_compute_d3q27_module.fpp:601-655)
 *   feq_kind 2: f_i = w_i (rho + rho0 (3 c.u + 9/2 (c.u)^2 - 3/2 u^2)), rho0 = 1 (incompressible) */
static void feq_generic(int QQ, int feq_kind, double rho, const double u[3], double *fEq) {
  const int *cx = ora_cxDir(QQ);
  const double *w = ora_weights(QQ);
  if (feq_kind == 1) {
    for (int d = 0; d < QQ; ++d) {
      double prod = rho;
      for (int k = 0; k < 3; ++k) {
        const int c = cx[3 * d + k];
        const double a = u[k];
        prod *= c == 0 ? (2.0 / 3.0 - a * a) : 0.5 * (1.0 / 3.0 + a * a + (double)c * a);
      }
      fEq[d] = prod;
    }
    return;
  }
  const double usq = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  for (int d = 0; d < QQ; ++d) {
    const double cu = (double)cx[3 * d] * u[0] + (double)cx[3 * d + 1] * u[1] + (double)cx[3 * d + 2] * u[2];
    const double poly = 3.0 * cu + 4.5 * cu * cu - 1.5 * usq;
    fEq[d] = feq_kind == 2 ? w[d] * (rho + poly) : w[d] * rho * (1.0 + poly);
  }
}

int ora_compute_generic(int relax, int QQ, int feq_kind, const double *in, double *out,
                        const double *aux, const int32_t *neigh, const double *omg,
                        int nSize, int nSolve, const ora_relax_t *rp) {
  if ((QQ != 19 && QQ != 27) || (relax != ORA_BGK && relax != ORA_TRT)) return -1;
  const int *inv = ora_cxDirInv(QQ);
  for (int e = 1; e <= nSolve; ++e) {
    double f[27], fEq[27];
    for (int d = 1; d <= QQ; ++d) f[d - 1] = PULL(d);
    const double u[3] = {AUX(1), AUX(2), AUX(3)};
    feq_generic(QQ, feq_kind, AUX(0), u, fEq);
    const double wP = omg[e - 1];
    if (relax == ORA_BGK) {
      for (int d = 0; d < QQ; ++d) SAVE(d + 1) = f[d] + wP * (fEq[d] - f[d]);
      continue;
    }
    /* TRT: omega^- from the magic parameter, Lambda = (1/omega^+ - 1/2)(1/omega^- - 1/2) */
    const double wN = 1.0 / (rp->lambda / (1.0 / wP - 0.5) + 0.5);
    for (int d = 0; d < QQ; ++d) {
      const int b = inv[d] - 1;
      const double fs = 0.5 * (f[d] + f[b]), fa = 0.5 * (f[d] - f[b]);
      const double es = 0.5 * (fEq[d] + fEq[b]), ea = 0.5 * (fEq[d] - fEq[b]);
      SAVE(d + 1) = f[d] - wP * (fs - es) - wN * (fa - ea);
    }
  }
  return 0;
}

/* ======================================================================== */
int ora_compute(int relax, int QQ, int incomp, const double *in, double *out,
                const double *aux, const int32_t *neigh, const double *omega,
                int nSize, int nSolve, const ora_relax_t *rp) {
  /* dispatch of mus_init_advRel_fluid / _fluid_incompressible
   * (mus/source/init/mus_initFluid_module.f90, mus_initFluidIncomp_module.f90:73-218);
   * fluid_incompressible + trt exists for d3q19 only (the reference aborts otherwise) */
  if (QQ == 19 && relax == ORA_BGK) { bgk_d3q19(incomp, in, out, aux, neigh, omega, nSize, nSolve); return 0; }
  if (QQ == 19 && relax == ORA_TRT) {
    if (incomp) trt_d3q19_incomp(in, out, aux, neigh, omega, nSize, nSolve, rp->lambda);
    else trt_d3q19(in, out, aux, neigh, omega, nSize, nSolve, rp->lambda);
    return 0;
  }
  if (QQ == 19 && relax == ORA_MRT) { mrt_d3q19(incomp, in, out, aux, neigh, omega, nSize, nSolve, rp->omegaBulk); return 0; }
  if (QQ == 27 && relax == ORA_BGK) { bgk_generic(27, incomp, in, out, aux, neigh, omega, nSize, nSolve); return 0; }
  if (QQ == 27 && relax == ORA_TRT && !incomp) { trt_d3q27(in, out, aux, neigh, omega, nSize, nSolve, rp->lambda); return 0; }
  if (QQ == 27 && relax == ORA_MRT) { mrt_d3q27(incomp, in, out, aux, neigh, omega, nSize, nSolve, rp->omegaBulk); return 0; }
  return -1;
}

int ora_compute_noopt_kind(int relax, int QQ, int incomp, const double *in, double *out,
                           const double *aux, const int32_t *neigh, const double *omega,
                           int nSize, int nSolve, const ora_relax_t *rp) {
  if (QQ != 19 && QQ != 27) return -1;
  if (relax == ORA_BGK) { bgk_generic(QQ, incomp, in, out, aux, neigh, omega, nSize, nSolve); return 0; }
  if (relax == ORA_MRT) { mrt_noopt(QQ, incomp, in, out, aux, neigh, omega, nSize, nSolve, rp->omegaBulk); return 0; }
  return -1;
}

int ora_compute_noopt(int relax, int QQ, const double *in, double *out,
                      const double *aux, const int32_t *neigh, const double *omega,
                      int nSize, int nSolve, const ora_relax_t *rp) {
  return ora_compute_noopt_kind(relax, QQ, 0, in, out, aux, neigh, omega, nSize, nSolve, rp);
}
