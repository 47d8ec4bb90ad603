/* intp.c -- ORACLE (test infrastructure): coarse <-> fine ghost interpolation,
 * restated from
 *   fillMyGhostsFromFiner_avg_feq_fneq       mus/source/intp/mus_interpolate_average_module.fpp:274-347
 *   fillArbiMyGhostsFromFiner_avg            mus_interpolate_average_module.fpp:95-185
 *   fillFinerGhostsFromMe_weighAvg_feq_fneq  mus_interpolate_average_module.fpp:905-1038
 *   fillFinerGhostsFromMe_linear_feq_fneq    mus/source/intp/mus_interpolate_linear_module.fpp:399-505
 *   mus_interpolate_linear3D_leastSq         mus_interpolate_linear_module.fpp:1010-1049
 *   fillFinerGhostsFromMe_quad_feq_fneq + mus_interpolate_quad3D_leastSq
 *                                            mus/source/intp/mus_interpolate_quadratic_module.fpp:292-..., 989-1034
 *   getNonEqFac_intp_*                       mus/source/mus_derivedQuantities_module.fpp:601-643
 * "parity unpinned by reference fixtures": the reference's interpolation utests are
 * deactivated and its multi-level goldens need Seeder meshes; pinned by analytic
 * properties (constant and linear fields are reproduced, conservation of the mean).
 * All positions are 1-based positions in the level's total list; state is AOS.
 */
#include "mus_oracle.h"
#include <stddef.h>

static double omega_from_visc(double v) { return 1.0 / (3.0 * v + 0.5); }
/* PULL build: factor for post-collision PDFs */
static double neq_fac(double omegaS, double omegaT) {
  return omegaS * (1.0 - omegaT) / ((1.0 - omegaS) * omegaT);
}

static void eq_neq(int QQ, int incomp, const double *sState, const double *sAux, int src,
                   double *f_eq, double *f_neq) {
  const double *f = sState + (size_t)(src - 1) * QQ;
  const double *a = sAux + (size_t)(src - 1) * 4;
  const double vel[3] = {a[1], a[2], a[3]};
  if (incomp) ora_pdfEq_incomp(QQ, a[0], vel, f_eq);
  else ora_pdfEq(QQ, a[0], vel, f_eq);
  for (int d = 0; d < QQ; ++d) f_neq[d] = f[d] - f_eq[d];
}

void ora_fill_my_ghosts_from_finer_avg(int QQ, int incomp, const double *sState,
                                       const double *sAux, double *tState, int nTargets,
                                       const int32_t *targetPos, const int32_t *srcOffset,
                                       const int32_t *srcPos, const double *tVisc) {
  for (int i = 0; i < nTargets; ++i) {
    const int tgt = targetPos[i];
    const int n = srcOffset[i + 1] - srcOffset[i];
    const double inv_n = 1.0 / (double)n;
    double t_eq[27], t_neq[27], fe[27], fn[27];
    for (int d = 0; d < QQ; ++d) { t_eq[d] = 0.0; t_neq[d] = 0.0; }
    for (int s = 0; s < n; ++s) {
      eq_neq(QQ, incomp, sState, sAux, srcPos[srcOffset[i] + s], fe, fn);
      for (int d = 0; d < QQ; ++d) { t_eq[d] = t_eq[d] + fe[d]; t_neq[d] = t_neq[d] + fn[d]; }
    }
    const double cVisc = tVisc[tgt - 1];
    const double fOmega = omega_from_visc(2.0 * cVisc);
    const double cOmega = omega_from_visc(cVisc);
    const double fac = 2.0 * neq_fac(fOmega, cOmega); /* getNonEqFac_intp_fine_to_coarse */
    for (int d = 0; d < QQ; ++d) {
      t_eq[d] = t_eq[d] * inv_n;
      t_neq[d] = t_neq[d] * inv_n * fac;
      tState[(size_t)(tgt - 1) * QQ + d] = t_eq[d] + t_neq[d];
    }
  }
}

void ora_fill_arbi_from_finer_avg(int nScalars, const double *sVal, double *tVal, int nTargets,
                                  const int32_t *targetPos, const int32_t *srcOffset,
                                  const int32_t *srcPos) {
  for (int i = 0; i < nTargets; ++i) {
    const int tgt = targetPos[i];
    const int n = srcOffset[i + 1] - srcOffset[i];
    const double inv_n = 1.0 / (double)n;
    double t[27];
    for (int k = 0; k < nScalars; ++k) t[k] = 0.0;
    for (int s = 0; s < n; ++s) {
      const int src = srcPos[srcOffset[i] + s];
      for (int k = 0; k < nScalars; ++k) t[k] = sVal[(size_t)(src - 1) * nScalars + k] + t[k];
    }
    for (int k = 0; k < nScalars; ++k) tVal[(size_t)(tgt - 1) * nScalars + k] = t[k] * inv_n;
  }
}

/* order 0: weights[]; order 1/2: least-square matrices (row-major nCoeff x nSrc) */
void ora_fill_finer_ghosts_from_me(int order, int QQ, int incomp, const double *sState,
                                   const double *sAux, double *tState, int nTargets,
                                   const int32_t *targetPos, const int32_t *srcOffset,
                                   const int32_t *srcPos, const double *weights,
                                   const int32_t *posInMat, const int32_t *matOffset,
                                   const double *matrices, const double *coord,
                                   const double *tVisc) {
  const int nCoeff = order == 1 ? 4 : 10;
  for (int i = 0; i < nTargets; ++i) {
    const int tgt = targetPos[i];
    const int n = srcOffset[i + 1] - srcOffset[i];
    double fe[27][27], fn[27][27], t_eq[27], t_neq[27];
    for (int s = 0; s < n; ++s) {
      double e[27], q[27];
      eq_neq(QQ, incomp, sState, sAux, srcPos[srcOffset[i] + s], e, q);
      for (int d = 0; d < QQ; ++d) { fe[d][s] = e[d]; fn[d][s] = q[d]; }
    }
    if (order == 0) {
      const double *w = weights + srcOffset[i];
      for (int d = 0; d < QQ; ++d) {
        double a = 0.0, b = 0.0;
        for (int s = 0; s < n; ++s) { a = a + w[s] * fe[d][s]; b = b + w[s] * fn[d][s]; }
        t_eq[d] = a; t_neq[d] = b;
      }
    } else {
      const double *A = matrices + matOffset[posInMat[i]];
      const double x = coord[3 * i + 0], y = coord[3 * i + 1], z = coord[3 * i + 2];
      for (int pass = 0; pass < 2; ++pass) {
        for (int d = 0; d < QQ; ++d) {
          const double *src = pass == 0 ? fe[d] : fn[d];
          double a[10];
          for (int k = 0; k < nCoeff; ++k) {
            double acc = 0.0;
            for (int s = 0; s < n; ++s) acc = acc + A[(size_t)k * n + s] * src[s];
            a[k] = acc;
          }
          double phi = a[0] + a[1] * x + a[2] * y + a[3] * z;
          if (order == 2)
            phi = phi + a[4] * x * x + a[5] * y * y + a[6] * z * z + a[7] * x * y + a[8] * y * z +
                  a[9] * z * x;
          if (pass == 0) t_eq[d] = phi; else t_neq[d] = phi;
        }
      }
    }
    const double fVisc = tVisc[tgt - 1];
    const double fOmega = omega_from_visc(fVisc);
    const double cOmega = omega_from_visc(0.5 * fVisc);
    const double fac = 0.5 * neq_fac(cOmega, fOmega); /* getNonEqFac_intp_coarse_to_fine */
    for (int d = 0; d < QQ; ++d) {
      t_neq[d] = t_neq[d] * fac;
      tState[(size_t)(tgt - 1) * QQ + d] = t_neq[d] + t_eq[d];
    }
  }
}

/* fillArbiFinerGhostsFromMe_weighAvg / _linear / _quad
 *   mus/source/intp/mus_interpolate_average_module.fpp:762-850 (sum(weight * sArbi) per variable),
 *   mus_interpolate_linear_module.fpp:124-205 + mus_interpolate_linear3D_leastSq :1010-1049,
 *   mus_interpolate_quadratic_module.fpp:102-... + mus_interpolate_quad3D_leastSq :989-1034
 * the reference's interpolation of ARBITRARY per-element values (it applies them to the
 * auxField); applied here to the nScalars = QQ PDFs of a passive scalar, whose ghosts have no
 * f_eq / f_neq rescaling.  sVal / tVal: AOS, nScalars per element. */
void ora_fill_arbi_finer_from_me(int order, int nScalars, const double *sVal, double *tVal,
                                 int nTargets, const int32_t *targetPos, const int32_t *srcOffset,
                                 const int32_t *srcPos, const double *weights,
                                 const int32_t *posInMat, const int32_t *matOffset,
                                 const double *matrices, const double *coord) {
  const int nCoeff = order == 1 ? 4 : 10;
  for (int i = 0; i < nTargets; ++i) {
    const int tgt = targetPos[i];
    const int n = srcOffset[i + 1] - srcOffset[i];
    for (int v = 0; v < nScalars; ++v) {
      double phi;
      if (order == 0) {
        const double *w = weights + srcOffset[i];
        phi = 0.0;
        for (int s = 0; s < n; ++s)
          phi = phi + w[s] * sVal[(size_t)(srcPos[srcOffset[i] + s] - 1) * nScalars + v];
      } else {
        const double *A = matrices + matOffset[posInMat[i]];
        const double x = coord[3 * i + 0], y = coord[3 * i + 1], z = coord[3 * i + 2];
        double a[10];
        for (int k = 0; k < nCoeff; ++k) {
          double acc = 0.0;
          for (int s = 0; s < n; ++s)
            acc = acc + A[(size_t)k * n + s] * sVal[(size_t)(srcPos[srcOffset[i] + s] - 1) * nScalars + v];
          a[k] = acc;
        }
        phi = a[0] + a[1] * x + a[2] * y + a[3] * z;
        if (order == 2)
          phi = phi + a[4] * x * x + a[5] * y * y + a[6] * z * z + a[7] * x * y + a[8] * y * z +
                a[9] * z * x;
      }
      tVal[(size_t)(tgt - 1) * nScalars + v] = phi;
    }
  }
}
