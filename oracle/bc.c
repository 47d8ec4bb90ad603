/* bc.c -- ORACLE (test infrastructure): boundary routines and the initial
 * condition, restated from
 *   fill_bcBuffer        mus/source/bc/mus_bc_general_module.fpp:1726-1768
 *   velocity_bounceback  mus/source/bc/mus_bc_fluid_module.fpp:1503-1597
 *   velocity_bounceback_incomp  (same file; rho0 = 1 replaces rho in eqPlus)
 *   fill_neighBuffer     mus/source/bc/mus_bc_general_module.fpp:1589-1717
 *   pressure_expol       mus/source/bc/mus_bc_fluid_module.fpp:1165-1362
 *   pressure_antiBounceBack  mus/source/bc/mus_bc_fluid_module.fpp:2161-2353
 *   mus_init_pdf         mus/source/mus_flow_module.fpp:484-589 (fEq + fNeq)
 * "parity unpinned by reference fixtures": no golden of the reference that is
 * reproducible without Seeder exercises these BCs in 3-D.
 */
#include "mus_oracle.h"
#include <stddef.h>
#include <stdlib.h>

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

/* fill_neighBuffer: nb(iNeigh, (iElem-1)*QQ + iDir), stored [nNeighs][nElems*QQ].
 * post = 0: neighBufferPre_nNext = currstate(FETCH(iDir, neighPos))  (:1647-1667)
 * post = 1: neighBufferPost      = currstate(SAVE (iDir, neighPos))  (:1670-1693)
 * neighPos[iElem][iNeigh] = fieldBC%neigh(level)%posInState(iNeigh, iElem)        */
void ora_fill_neighBuffer(double *nb, const double *state, const int32_t *neigh, int nSize,
                          int QQ, int nNeighs, int nElems, const int32_t *neighPos, int post) {
  for (int iN = 1; iN <= nNeighs; ++iN)
    for (int i = 1; i <= nElems; ++i) {
      const int np = neighPos[(size_t)(i - 1) * nNeighs + iN - 1];
      for (int d = 1; d <= QQ; ++d) {
        const size_t src = post ? (size_t)(np - 1) * QQ + d
                                : (size_t)neigh[(size_t)(d - 1) * nSize + np - 1];
        nb[(size_t)(iN - 1) * nElems * QQ + (size_t)(i - 1) * QQ + d - 1] = state[src - 1];
      }
    }
}

/* pressure_expol: the links of the boundary are extrapolated from the two neighbours along
 * the inward normal, 1.5 f(1) - 0.5 f(2) of neighBufferPre_nNext; the axis-normal link gets
 * fEq0(rho_bc, u) + (f_post(inv) - fEq(rho, u)(inv)) (:1316-1340).  rhoDef = boundary pressure
 * converted to lattice density by the host (:1270).  No element here has prp_hasQVal.       */
void ora_pressure_expol(double *state, const double *bcBuffer, const double *aux,
                        const int32_t *neigh, int nSize, int QQ, int incompressible, int nElems,
                        const int32_t *elemPos, const int32_t *posInBcElemBuf,
                        const int32_t *normalInd, const double *rhoDef, int nLinks,
                        const int32_t *links, const int32_t *statePos, const double *nbPre) {
  const int *cx = ora_cxDir(QQ);
  const int *inv = ora_cxDirInv(QQ);
  for (int l = 1; l <= nLinks; ++l) {
    const double fTmp_1 = nbPre[statePos[l - 1] - 1];
    const double fTmp_2 = nbPre[(size_t)nElems * QQ + statePos[l - 1] - 1];
    state[links[l - 1] - 1] = 1.5 * fTmp_1 - 0.5 * fTmp_2;
  }
  for (int i = 1; i <= nElems; ++i) {
    const int e = elemPos[i - 1];
    const int iDir = normalInd[i - 1];
    const int axisNormal = (abs(cx[3 * (iDir - 1)]) + abs(cx[3 * (iDir - 1) + 1]) +
                            abs(cx[3 * (iDir - 1) + 2])) == 1;
    if (!axisNormal) continue;
    const double *a = aux + (size_t)(e - 1) * 4;
    double fEq[27], fEq0[27];
    if (incompressible) { ora_pdfEq_incomp(QQ, a[0], a + 1, fEq); ora_pdfEq_incomp(QQ, rhoDef[i - 1], a + 1, fEq0); }
    else { ora_pdfEq(QQ, a[0], a + 1, fEq); ora_pdfEq(QQ, rhoDef[i - 1], a + 1, fEq0); }
    const int invDir = inv[iDir - 1];
    const double fPostCol = bcBuffer[(size_t)(posInBcElemBuf[i - 1] - 1) * QQ + invDir - 1];
    state[neigh[(size_t)(iDir - 1) * nSize + e - 1] - 1] = fEq0[iDir - 1] + (fPostCol - fEq[invDir - 1]);
  }
}

/* pressure_antiBounceBack: anti-bounce-back with the velocity extrapolated to the boundary,
 * uxB = 1.5 uxF - 0.5 uxN (uxN from neighBufferPost(1,:)); bitmask links elem-major.        */
void ora_pressure_antibounceback(double *state, const double *bcBuffer, int QQ, int incompressible,
                                 int nElems, const int32_t *elemPos, const int32_t *posInBcElemBuf,
                                 const double *rhoDef, const double *omega /* per total elem */,
                                 int nLinks, const int32_t *links, const int32_t *iElemOfLink,
                                 const int32_t *iDirOfLink, const double *nbPost) {
  const int *cx = ora_cxDir(QQ);
  const int *inv = ora_cxDirInv(QQ);
  const double *w = ora_weights(QQ);
  const double rho0 = 1.0, div1_3 = 1.0 / 3.0;
  for (int l = 1; l <= nLinks; ++l) {
    const int i = iElemOfLink[l - 1], iDir = iDirOfLink[l - 1], invDir = inv[iDir - 1];
    const double *fTmp = bcBuffer + (size_t)(posInBcElemBuf[i - 1] - 1) * QQ - 1; /* 1-based */
    const double *fN = nbPost + (size_t)(i - 1) * QQ - 1;
    double rhoF = 0.0, rhoN = 0.0;
    for (int d = 1; d <= QQ; ++d) { rhoF = rhoF + fTmp[d]; rhoN = rhoN + fN[d]; }
    double uxF[3], uxN[3], uxB[3];
    ora_first_moment(QQ, fTmp, uxF);
    ora_first_moment(QQ, fN, uxN);
    for (int k = 0; k < 3; ++k) {
      if (!incompressible) { uxF[k] = uxF[k] / rhoF; uxN[k] = uxN[k] / rhoN; }
      uxB[k] = 1.5 * uxF[k] - 0.5 * uxN[k];
    }
    const double usqB = uxB[0] * uxB[0] + uxB[1] * uxB[1] + uxB[2] * uxB[2];
    const double usqF = uxF[0] * uxF[0] + uxF[1] * uxF[1] + uxF[2] * uxF[2];
    const double om = omega[elemPos[i - 1] - 1];
    const double c0 = (double)cx[3 * (invDir - 1)], c1 = (double)cx[3 * (invDir - 1) + 1],
                 c2 = (double)cx[3 * (invDir - 1) + 2];
    const double cuF = c0 * uxF[0] + c1 * uxF[1] + c2 * uxF[2];
    const double cuB = c0 * uxB[0] + c1 * uxB[1] + c2 * uxB[2];
    const double fEqPlusFluid = w[iDir - 1] * rhoF + 4.5 * w[iDir - 1] * rho0 * (cuF * cuF - div1_3 * usqF);
    const double fEqPlus = w[iDir - 1] * rhoDef[i - 1] + 4.5 * w[iDir - 1] * rho0 * (cuB * cuB - div1_3 * usqB);
    const double fPlusFluid = 0.5 * (fTmp[iDir] + fTmp[invDir]);
    state[links[l - 1] - 1] = -fTmp[invDir] + 2.0 * fEqPlus + (2.0 - om) * (fPlusFluid - fEqPlusFluid);
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
