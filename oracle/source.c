/* source.c -- ORACLE (test infrastructure): body-force source terms and the
 * passive-scalar kernels of the reference restated in C, same operation order.
 *
 *   mus_addForceToAuxField_fluid        mus/source/derived/mus_auxFieldVar_module.fpp:1032-1089
 *   mus_addForceToAuxField_fluidIncomp  mus/source/derived/mus_auxFieldVar_module.fpp:1160-1214
 *   applySrc_force (BGK/TRT, 2nd order) mus/source/derived/mus_derQuan_module.fpp:3129-3243
 *   applySrc_force_MRT_d3q27            mus/source/derived/mus_derQuan_module.fpp:3555-3693
 *   applySrc_force_MRT_d3q19            mus/source/derived/mus_derQuan_module.fpp:3716-3862
 *   applySrc_force1stOrd                mus/source/derived/mus_derQuan_module.fpp:4043-4143
 *   selection                           mus/source/derived/mus_variable_module.f90:1043-1081
 *   mus_advRel_kPS_rBGK_v1st_l / _v2nd_l / rTRT_vStdNoOpt_l
 *                                       mus/source/compute/mus_compute_passiveScalar_module.fpp:77-398
 *   mus_calcAuxField_zerothMoment       mus/source/derived/mus_auxFieldVar_module.fpp:825-869
 *
 * The force is handed over in LATTICE units (forceField / fac%body_force), one
 * triple per source element; posInTotal is fun%elemLvl(iLevel)%posInTotal (1-based).
 * "Parity unpinned by reference fixtures": no golden of the reference exercises
 * these routines on a mesh reproducible here; they are pinned by analytic
 * checks (momentum balance, Poiseuille profile, MRT -> BGK limit) in tests/.
 */
#include "mus_oracle.h"
#include "mrt_tables.h"
#include <stddef.h>

static const double cs2inv = 3.0, cs4inv = 9.0;

void ora_add_force_to_aux(double *aux, int incompressible, int nElems, const int32_t *posInTotal,
                          const double *force) {
  for (int i = 0; i < nElems; ++i) {
    const size_t off = (size_t)(posInTotal[i] - 1) * 4;
    /* forceTerm = forceField * 0.5 * inv_rho (rho0Inv = 1 for the incompressible model) */
    const double inv_rho = incompressible ? 1.0 : 1.0 / aux[off + 0];
    for (int k = 0; k < 3; ++k) {
      const double forceTerm = force[3 * (size_t)i + k] * 0.5 * inv_rho;
      aux[off + 1 + k] = aux[off + 1 + k] + forceTerm;
    }
  }
}

/* order 2, relaxation bgk / trt: Guo forcing with the BGK prefactor (1 - omega/2) */
static void apply_force_2nd(int QQ, double *out, const double *aux, const double *omega, int nElems,
                            const int32_t *posInTotal, const double *force) {
  const int *cx = ora_cxDir(QQ);
  const double *w = ora_weights(QQ);
  for (int i = 0; i < nElems; ++i) {
    const int e = posInTotal[i];
    const double *vel = aux + (size_t)(e - 1) * 4 + 1;
    const double *G = force + 3 * (size_t)i;
    const double omega_fac = 1.0 - omega[e - 1] * 0.5;
    for (int d = 0; d < QQ; ++d) {
      const double c[3] = {(double)cx[3 * d], (double)cx[3 * d + 1], (double)cx[3 * d + 2]};
      const double ucx = c[0] * vel[0] + c[1] * vel[1] + c[2] * vel[2];
      double t[3];
      for (int k = 0; k < 3; ++k) t[k] = (c[k] - vel[k]) * cs2inv + ucx * c[k] * cs4inv;
      const double forceTerm = t[0] * G[0] + t[1] * G[1] + t[2] * G[2];
      double *o = out + (size_t)(e - 1) * QQ + d;
      *o = *o + omega_fac * w[d] * forceTerm;
    }
  }
}

static void apply_force_1st(int QQ, double *out, int nElems, const int32_t *posInTotal,
                            const double *force) {
  const int *cx = ora_cxDir(QQ);
  const double *w = ora_weights(QQ);
  for (int i = 0; i < nElems; ++i) {
    const int e = posInTotal[i];
    const double *F = force + 3 * (size_t)i;
    for (int d = 0; d < QQ; ++d) {
      const double forceTerm = (double)cx[3 * d] * F[0] + (double)cx[3 * d + 1] * F[1] +
                               (double)cx[3 * d + 2] * F[2];
      double *o = out + (size_t)(e - 1) * QQ + d;
      *o = *o + w[d] * cs2inv * forceTerm;
    }
  }
}

static void apply_force_mrt_d3q19(double *out, const double *aux, const double *omega,
                                  double omegaBulk, int nElems, const int32_t *posInTotal,
                                  const double *force) {
  enum { QQ = 19 };
  const double *A = ora_mrt_matrix(19, 1); /* toPDF%A, row = direction */
  double s0[QQ];
  ora_mrt_diag(19, 1.0, omegaBulk, s0);
  for (int k = 0; k < QQ; ++k) s0[k] = 1.0 - 0.5 * s0[k];
  for (int i = 0; i < nElems; ++i) {
    const int e = posInTotal[i];
    const double *v = aux + (size_t)(e - 1) * 4 + 1;
    const double *F = force + 3 * (size_t)i;
    double sl[QQ], mom[QQ];
    for (int k = 0; k < QQ; ++k) { sl[k] = s0[k]; mom[k] = 0.0; }
    sl[10 - 1] = 1.0 - 0.5 * omega[e - 1];
    sl[12 - 1] = sl[10 - 1]; sl[14 - 1] = sl[10 - 1]; sl[15 - 1] = sl[10 - 1]; sl[16 - 1] = sl[10 - 1];
    mom[2 - 1] = 2.0 * (F[0] * v[0] + F[1] * v[1] + F[2] * v[2]);
    mom[4 - 1] = F[0];
    mom[6 - 1] = F[1];
    mom[8 - 1] = F[2];
    mom[10 - 1] = -2.0 * (F[1] * v[1] - 2.0 * F[0] * v[0] + F[2] * v[2]);
    mom[12 - 1] = 2.0 * (F[1] * v[1] - F[2] * v[2]);
    mom[14 - 1] = F[0] * v[1] + F[1] * v[0];
    mom[15 - 1] = F[1] * v[2] + F[2] * v[1];
    mom[16 - 1] = F[0] * v[2] + F[2] * v[0];
    for (int d = 0; d < QQ; ++d) {
      double disc = 0.0; /* sum(mInvXOmega(iDir,1:QQ) * momForce(1:QQ)) */
      for (int k = 0; k < QQ; ++k) disc = disc + (A[d * QQ + k] * sl[k]) * mom[k];
      double *o = out + (size_t)(e - 1) * QQ + d;
      *o = *o + disc;
    }
  }
}

static void apply_force_mrt_d3q27(double *out, const double *aux, const double *omega,
                                  double omegaBulk, int nElems, const int32_t *posInTotal,
                                  const double *force) {
  enum { QQ = 27 };
  const double *A = ora_mrt_matrix(27, 1);
  double s0[QQ];
  ora_mrt_diag(27, 1.0, omegaBulk, s0);
  for (int k = 1; k <= 3; ++k) s0[k] = 1.0 - 0.5 * s0[k]; /* s_mrt(2:4)  */
  s0[9] = 1.0 - 0.5 * s0[9];                                /* s_mrt(10)   */
  for (int i = 0; i < nElems; ++i) {
    const int e = posInTotal[i];
    const double *v = aux + (size_t)(e - 1) * 4 + 1;
    const double *F = force + 3 * (size_t)i;
    double sl[QQ], mom[QQ];
    for (int k = 0; k < QQ; ++k) { sl[k] = s0[k]; mom[k] = 0.0; }
    for (int k = 5; k <= 9; ++k) sl[k - 1] = 1.0 - 0.5 * omega[e - 1];
    mom[2 - 1] = F[0]; mom[3 - 1] = F[1]; mom[4 - 1] = F[2];
    mom[5 - 1] = F[0] * v[1] + F[1] * v[0];
    mom[6 - 1] = F[1] * v[2] + F[2] * v[1];
    mom[7 - 1] = F[0] * v[2] + F[2] * v[0];
    mom[8 - 1] = -2.0 * (F[1] * v[1] - 2.0 * F[0] * v[0] + F[2] * v[2]);
    mom[9 - 1] = 2.0 * (F[1] * v[1] - F[2] * v[2]);
    mom[10 - 1] = 2.0 * (F[0] * v[0] + F[1] * v[1] + F[2] * v[2]);
    for (int d = 0; d < QQ; ++d) {
      double disc = 0.0; /* dot_product(mInvXOmega(iDir,2:10), momForce(2:10)) */
      for (int k = 1; k <= 9; ++k) disc = disc + (A[d * QQ + k] * sl[k]) * mom[k];
      double *o = out + (size_t)(e - 1) * QQ + d;
      *o = *o + disc;
    }
  }
}

int ora_apply_src_force(int relax, int QQ, int order, double *out, const double *aux,
                        const double *omega, double omegaBulk, int nElems,
                        const int32_t *posInTotal, const double *force) {
  if (QQ != 19 && QQ != 27) return -1;
  if (order == 1) { apply_force_1st(QQ, out, nElems, posInTotal, force); return 0; }
  if (order != 2) return -1;
  if (relax == ORA_MRT) {
    if (QQ == 19) apply_force_mrt_d3q19(out, aux, omega, omegaBulk, nElems, posInTotal, force);
    else apply_force_mrt_d3q27(out, aux, omega, omegaBulk, nElems, posInTotal, force);
    return 0;
  }
  apply_force_2nd(QQ, out, aux, omega, nElems, posInTotal, force);
  return 0;
}

/* ---- passive scalar ---------------------------------------------------- */
void ora_calc_aux_zeroth(int QQ, double *aux, const double *state, const int32_t *neigh,
                         int nSize, int nSolve) {
  for (int e = 1; e <= nSolve; ++e) {
    double rho = 0.0;
    for (int d = 1; d <= QQ; ++d)
      rho = rho + state[neigh[(size_t)(d - 1) * nSize + (e - 1)] - 1];
    aux[e - 1] = rho;
  }
}

/* variant: 1 = bgk/first, 2 = bgk/second, 3 = trt (vStdNoOpt); transVel in lattice units,
 * [nSolve][3]; d_omega = 2/(1 + 6 diff_coeff); lambda = species%lambda (trt) */
int ora_compute_passive_scalar(int variant, int QQ, const double *in, double *out,
                               const int32_t *neigh, int nSize, int nSolve,
                               const double *transVel, double diff_coeff, double lambda) {
  if ((QQ != 19 && QQ != 27) || variant < 1 || variant > 3) return -1;
  const int *cx = ora_cxDir(QQ);
  const int *inv = ora_cxDirInv(QQ);
  const double *w = ora_weights(QQ);
  const double d_omega = 2.0 / (1.0 + 6.0 * diff_coeff);
  const double aux_omega = 1.0 / (lambda / (1.0 / d_omega - 0.5) + 0.5);
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    const double *u = transVel + 3 * (size_t)(e - 1);
    double pdf[27], rho = 0.0;
    for (int d = 0; d < QQ; ++d) {
      pdf[d] = in[neigh[(size_t)d * nSize + (e - 1)] - 1];
      rho = rho + pdf[d];
    }
    for (int d = 0; d < QQ; ++d) {
      const double uc = (double)cx[3 * d] * u[0] + (double)cx[3 * d + 1] * u[1] +
                        (double)cx[3 * d + 2] * u[2];
      double *o = out + (size_t)(e - 1) * QQ + d;
      if (variant == 1) {
        const double feq = rho * w[d] * (1.0 + 3.0 * uc);
        *o = pdf[d] + d_omega * (feq - pdf[d]);
      } else if (variant == 2) {
        const double usq = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
        const double feq = rho * w[d] * (1.0 + 3.0 * uc + 9.0 * uc * uc * 0.5 - usq * 0.5 * 3.0);
        *o = pdf[d] + d_omega * (feq - pdf[d]);
      } else {
        const double usq = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
        const double feqPlus = rho * w[d] * (1.0 + 9.0 * uc * uc * 0.5 - usq * 0.5 * 3.0);
        const double feqMinus = rho * w[d] * 3.0 * uc;
        const int id = inv[d] - 1;
        const double fPlus = 0.5 * (pdf[d] + pdf[id]);
        const double fMinus = 0.5 * (pdf[d] - pdf[id]);
        *o = pdf[d] + d_omega * (feqMinus - fMinus) + aux_omega * (feqPlus - fPlus);
      }
    }
  }
  return 0;
}
