/* bc.c -- ORACLE (test infrastructure): boundary routines and the initial
 * condition, restated from
 *   fill_bcBuffer        mus/source/bc/mus_bc_general_module.fpp:1726-1768
 *   velocity_bounceback  mus/source/bc/mus_bc_fluid_module.fpp:1503-1597
 *   velocity_bounceback_incomp  (same file; rho0 = 1 replaces rho in eqPlus)
 *   mus_init_pdf         mus/source/mus_flow_module.fpp:484-589 (fEq + fNeq, S = 0)
 * "parity unpinned by reference fixtures": no golden of the reference that is
 * reproducible without Seeder exercises these BCs in 3-D.
 */
#include "mus_oracle.h"
#include <stddef.h>

void ora_fill_bcBuffer(double *bcBuffer, const double *state, int QQ,
                       const int32_t *bcElems, int nBcElems) {
  /* bcBuffer(varPos + (iElem-1)*nScalars) = state(IDX(iDir, posInTotal(iElem))) */
  for (int i = 1; i <= nBcElems; ++i) {
    const int e = bcElems[i - 1];
    for (int d = 1; d <= QQ; ++d)
      bcBuffer[(size_t)(i - 1) * QQ + d - 1] = state[(size_t)(e - 1) * QQ + d - 1];
  }
}

void ora_velocity_bounceback(double *state, const double *bcBuffer, int QQ,
                             int nLinks, const int32_t *links, const int32_t *outPos,
                             const int32_t *iDirLink, const int32_t *posInBuffer,
                             const double *velLat /* [nLinks][3] = vel_b*inv_vel */,
                             int incompressible) {
  const int *cx = ora_cxDir(QQ);
  const double *w = ora_weights(QQ);
  for (int l = 1; l <= nLinks; ++l) {
    const double fOut = bcBuffer[outPos[l - 1] - 1];
    const int pib = posInBuffer[l - 1];
    double rho = 0.0;
    for (int d = 1; d <= QQ; ++d) rho = rho + bcBuffer[(size_t)(pib - 1) * QQ + d - 1];
    if (incompressible) rho = 1.0; /* velocity_bounceback_incomp uses rho0 */
    const int iDir = iDirLink[l - 1];
    const double *c = velLat + (size_t)(l - 1) * 3;
    const double eqPlus = w[iDir - 1] * 6.0 * rho *
        ((double)cx[3 * (iDir - 1) + 0] * c[0] + (double)cx[3 * (iDir - 1) + 1] * c[1] +
         (double)cx[3 * (iDir - 1) + 2] * c[2]);
    state[links[l - 1] - 1] = fOut + eqPlus;
  }
}

/* mus_init_pdf (mus_flow_module.fpp:422-601), acoustic scaling: state = fEq(rho, vel) +
 * fNeq(omega, S); S6 = (Sxx, Syy, Szz, Sxy, Syz, Sxz) per element in LATTICE units, omega per
 * element.  The 3x3 tensor is assembled as at :548-552.                       */
void ora_init_pdf(int QQ, int incompressible, int nElems, const double *rho, const double *vel,
                  const double *S6, const double *omega, double *state) {
  for (int e = 0; e < nElems; ++e) {
    double fEq[27], fNeq[27];
    if (incompressible) ora_pdfEq_incomp(QQ, rho[e], vel + 3 * (size_t)e, fEq);
    else ora_pdfEq(QQ, rho[e], vel + 3 * (size_t)e, fEq);
    const double *s = S6 + 6 * (size_t)e;
    const double S[9] = {s[0], s[3], s[5], s[3], s[1], s[4], s[5], s[4], s[2]};
    ora_nEq_acoustic(QQ, omega[e], S, fNeq);
    for (int d = 0; d < QQ; ++d) state[(size_t)e * QQ + d] = fEq[d] + fNeq[d];
  }
}

void ora_init_equilibrium(int QQ, int incompressible, int nElems, const double *rho,
                          const double *vel /* [nElems][3] */, double *state) {
  for (int e = 0; e < nElems; ++e) {
    double fEq[27];
    if (incompressible) ora_pdfEq_incomp(QQ, rho[e], vel + 3 * (size_t)e, fEq);
    else ora_pdfEq(QQ, rho[e], vel + 3 * (size_t)e, fEq);
    for (int d = 0; d < QQ; ++d) state[(size_t)e * QQ + d] = fEq[d] + 0.0; /* + fNeq(S=0) */
  }
}
