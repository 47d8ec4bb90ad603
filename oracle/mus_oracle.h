/* mus_oracle.h -- CPU ORACLE for the Musubi per-level LBM time step.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * CPU algorithm (AOS state, PULL streaming through the `neigh` position list,
 * separate auxField sweep, per-element omega) used as the checker by tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py.  Nothing under musubi_b200/ may include, link or call it.
 *
 * Parity status: PINNED for fluid/BGK/D3Q19 + IC + unit conversion + periodic
 * connectivity by the reference's own golden
 *   mus/examples/fluid/benchmark/gaussianPulse/reference/ (the .res files)
 * (tests/test_oracle_golden.py) and by restated utest properties
 * (optimised-vs-NoOpt, rest-state fixed point, M*Minv = I).  TRT D3Q19,
 * BGK/TRT/MRT D3Q27, the boundary and interpolation routines are
 * "parity unpinned by reference fixtures" (no reproducible golden exists
 * without Seeder / a Fortran toolchain); they are pinned by analytic checks.
 *
 * Conventions follow the reference's default build (AOS + PULL,
 * mus/source/header/lbm_macros.inc:70-110), all positions 1-based as in
 * Fortran:
 *   state position  IDX(dir,elem)   = (elem-1)*QQ + dir
 *   neigh position  NGPOS(dir,elem) = (dir-1)*nSize + elem
 *   aux position    (elem-1)*4 + {1:rho, 2:ux, 3:uy, 4:uz}
 */
#ifndef MUS_ORACLE_H
#define MUS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORA_BGK = 0, ORA_TRT = 1, ORA_MRT = 2 };

/* relaxation parameters the kernels read besides omega(elem) */
typedef struct {
  double lambda;     /* TRT magic parameter, fluid%lambda (mus_fluid_module.f90:104) */
  double omegaBulk;  /* fluid%omegaBulkLvl(level) (mus_fluid_module.f90:482-484)      */
} ora_relax_t;

/* ---- stencil tables (tem_stencil_module.fpp:91-168) -------------------- */
const int *ora_cxDir(int QQ);      /* [QQ][3] */
const int *ora_cxDirInv(int QQ);   /* [QQ], 1-based values */
const double *ora_weights(int QQ); /* mus_scheme_layout_module.f90:699-705 */

/* ---- equilibrium + moments (mus_scheme_derived_quantities_type_module.f90) */
void ora_pdfEq(int QQ, double rho, const double vel[3], double *fEq);
void ora_pdfEq_incomp(int QQ, double rho, const double vel[3], double *fEq);

/* mus_calcAuxField_fluid_d3q19/_d3q27 (mus_auxFieldVar_module.fpp:605-817) */
void ora_calc_aux(int QQ, double *aux, const double *state, const int32_t *neigh,
                  int nSize, int nSolve);
void ora_calc_aux_incomp(int QQ, double *aux, const double *state,
                         const int32_t *neigh, int nSize, int nSolve);

/* mus_update_relaxParamKine (mus_relaxationParam_module.f90:276-280) */
void ora_update_omega(double *omega, const double *visc, int nSolve);
double ora_omega_bulk(double viscBulkLat); /* mus_fluid_module.f90:482-484 */

/* ---- the compute kernels (scheme%compute pointees) ---------------------- */
/* relax = ORA_BGK/TRT/MRT, QQ = 19/27; returns 0 or -1 for an unknown combo  */
int ora_compute(int relax, int QQ, int incompressible, const double *in, double *out,
                const double *aux, const int32_t *neigh, const double *omega,
                int nSize, int nSolve, const ora_relax_t *rp);
/* generic unoptimised kernels used by the reference's utests as the
 * comparison partner (mus_compute_bgk_module.fpp:77-159,
 * mus_compute_mrt_d3q19_module.fpp:999-1100, mus_compute_mrt_d3q27_module.fpp:88-185) */
int ora_compute_noopt(int relax, int QQ, const double *in, double *out,
                      const double *aux, const int32_t *neigh, const double *omega,
                      int nSize, int nSolve, const ora_relax_t *rp);
/* the same with the incompressible equilibrium (mus_advRel_kFluidIncomp_rMRT_vStdNoOpt_lD3Q19,
 * mus_compute_mrt_d3q19_module.fpp:751-866, written there with explicit m_eq) */
int ora_compute_noopt_kind(int relax, int QQ, int incompressible, const double *in, double *out,
                           const double *aux, const int32_t *neigh, const double *omega,
                           int nSize, int nSolve, const ora_relax_t *rp);
/* generic formulations written from the textbook definitions (table-driven equilibrium, explicit
 * symmetric / antisymmetric split), independent of the optimised restatements: relax = ORA_BGK |
 * ORA_TRT; feq_kind 0 second-order polynomial, 1 product form (TRT D3Q27), 2 incompressible */
int ora_compute_generic(int relax, int QQ, int feq_kind, const double *in, double *out,
                        const double *aux, const int32_t *neigh, const double *omega,
                        int nSize, int nSolve, const ora_relax_t *rp);
const double *ora_mrt_matrix(int QQ, int inverse); /* [QQ][QQ] row-major */
void ora_mrt_diag(int QQ, double omegaKine, double omegaBulk, double *s_mrt);

/* ---- connectivity (mus_connectivity_module.fpp:73-179) ------------------ */
/* nghElems: [nElems][QQN] (row per element, 1-based neighbour positions, <=0
 * = missing / BC id); property: prp bits per element; offsets as levelDesc. */
void ora_construct_connectivity(int32_t *neigh, int nSize, int nElems, int QQ,
                                const int32_t *nghElems, const int64_t *property,
                                int nFluid, int haloOffset /* offset(1,eT_halo) */);

/* ---- boundary conditions (bc/mus_bc_fluid_module.fpp) ------------------- */
void ora_fill_bcBuffer(double *bcBuffer, const double *state, int QQ,
                       const int32_t *bcElems, int nBcElems);
void ora_velocity_bounceback(double *state, const double *bcBuffer, int QQ,
                             int nLinks, const int32_t *links, const int32_t *outPos,
                             const int32_t *iDirLink, const int32_t *posInBuffer,
                             const double *velLat /* [nLinks][3] lattice units */,
                             int incompressible);
void ora_first_moment(int QQ, const double *pdf_1based, double m[3]);
/* fill_neighBuffer (mus_bc_general_module.fpp:1589-1717), pressure_expol
 * (mus_bc_fluid_module.fpp:1165-1362), pressure_antiBounceBack (:2161-2353) */
void ora_fill_neighBuffer(double *nb, const double *state, const int32_t *neigh, int nSize,
                          int QQ, int nNeighs, int nElems, const int32_t *neighPos, int post);
void ora_pressure_expol(double *state, const double *bcBuffer, const double *aux,
                        const int32_t *neigh, int nSize, int QQ, int incompressible, int nElems,
                        const int32_t *elemPos, const int32_t *posInBcElemBuf,
                        const int32_t *normalInd, const double *rhoDef, int nLinks,
                        const int32_t *links, const int32_t *statePos, const double *nbPre);
void ora_pressure_antibounceback(double *state, const double *bcBuffer, int QQ, int incompressible,
                                 int nElems, const int32_t *elemPos, const int32_t *posInBcElemBuf,
                                 const double *rhoDef, const double *omega, int nLinks,
                                 const int32_t *links, const int32_t *iElemOfLink,
                                 const int32_t *iDirOfLink, const double *nbPost);
/* mus_init_pdf with zero strain rate: state = fEq(rho, vel) */
void ora_init_equilibrium(int QQ, int incompressible, int nElems, const double *rho,
                          const double *vel, double *state);

/* mus_init_pdf with the acoustic non-equilibrium part getNEq_acoustic
 * (mus_flow_module.fpp:422-601, mus_derivedQuantities_module.fpp:441-478) */
void ora_nEq_acoustic(int QQ, double omega, const double *S /* 3x3 column-major */, double *nEq);
void ora_init_pdf(int QQ, int incompressible, int nElems, const double *rho, const double *vel,
                  const double *S6 /* [nElems][6] Sxx,Syy,Szz,Sxy,Syz,Sxz */,
                  const double *omega /* [nElems] */, double *state);

/* ---- ghost interpolation (mus/source/intp) ----------------------------- */
void ora_fill_my_ghosts_from_finer_avg(int QQ, int incomp, const double *sState,
                                       const double *sAux, double *tState, int nTargets,
                                       const int32_t *targetPos, const int32_t *srcOffset,
                                       const int32_t *srcPos, const double *tVisc);
void ora_fill_arbi_from_finer_avg(int nScalars, const double *sVal, double *tVal, int nTargets,
                                  const int32_t *targetPos, const int32_t *srcOffset,
                                  const int32_t *srcPos);
void ora_fill_finer_ghosts_from_me(int order, int QQ, int incomp, const double *sState,
                                   const double *sAux, double *tState, int nTargets,
                                   const int32_t *targetPos, const int32_t *srcOffset,
                                   const int32_t *srcPos, const double *weights,
                                   const int32_t *posInMat, const int32_t *matOffset,
                                   const double *matrices, const double *coord,
                                   const double *tVisc);

/* the reference's interpolation of arbitrary per-element values (fillArbiFinerGhostsFromMe_*),
 * applied to the PDFs of a passive scalar on a multi-level mesh */
void ora_fill_arbi_finer_from_me(int order, int nScalars, const double *sVal, double *tVal,
                                 int nTargets, const int32_t *targetPos, const int32_t *srcOffset,
                                 const int32_t *srcPos, const double *weights,
                                 const int32_t *posInMat, const int32_t *matOffset,
                                 const double *matrices, const double *coord);

/* ---- ghost dependency build (dependencies.c): an implementation independent of the
 * product's host-side generator, compared with it list by list ------------------------------- */
/* tem_build_verticalDependencies (tem_construction_module.f90:2894-2985); blockStart[0..4]:
 * 0-based start of the fluid / ghostFromCoarser / ghostFromFiner / halo blocks of `total`, end */
void ora_vertical_dep_from_coarser(int nGhost, const int64_t *ghostID, const int64_t *cTotal,
                                   const int32_t *cBlockStart, int32_t *parentPos, int32_t *childNum,
                                   double *coord);
void ora_vertical_dep_from_finer(int nGhost, const int64_t *ghostID, const int64_t *fTotal,
                                 const int32_t *fBlockStart, int32_t *srcOffset, int32_t *srcPos);
/* tem_intpMatrixLSF_type store of one interpolation order (tem_matrix_module.fpp:161-425) */
void *ora_lsf_new(int order);
void ora_lsf_delete(void *store);
int ora_lsf_count(const void *store);
int ora_lsf_get(const void *store, int pos, int32_t *hashID, int32_t *invertible, int32_t *rows,
                int32_t *cols, double *A);
int ora_lsf_append(void *store, int QQ, int nSources, const int32_t *neighDir, int32_t *pos);
/* mus_intp_update_depFromCoarser (mus_interpolate_module.fpp:544-833) for one target level */
int ora_update_dep_from_coarser(int QQ, int orderMax, int nGhost, const int32_t *parentPos,
                                const int32_t *childNum, const double *coord,
                                const int32_t *cNghElems, int32_t *order, int32_t *nSrc,
                                int32_t *src, int32_t *dir, int32_t *posInMat, double *weights,
                                void *lin, void *quad);

/* ---- source terms and passive scalar (source.c) ------------------------ */
/* force in lattice units, [nElems][3]; posInTotal = fun%elemLvl(iLevel)%posInTotal (1-based) */
void ora_add_force_to_aux(double *aux, int incompressible, int nElems, const int32_t *posInTotal,
                          const double *force);
/* order 2: applySrc_force (bgk, trt) / applySrc_force_MRT_d3q19 / _d3q27; order 1:
 * applySrc_force1stOrd.  omega = omLvl(iLevel)%val, indexed by posInTotal */
int ora_apply_src_force(int relax, int QQ, int order, double *out, const double *aux,
                        const double *omega, double omegaBulk, int nElems,
                        const int32_t *posInTotal, const double *force);
/* mus_calcAuxField_zerothMoment: aux(iElem) = sum of the pulled PDFs (nAuxScalars = 1) */
void ora_calc_aux_zeroth(int QQ, double *aux, const double *state, const int32_t *neigh,
                         int nSize, int nSolve);
/* variant 1 bgk/first, 2 bgk/second, 3 trt (vStdNoOpt) */
int ora_compute_passive_scalar(int variant, int QQ, const double *in, double *out,
                               const int32_t *neigh, int nSize, int nSolve,
                               const double *transVel, double diff_coeff, double lambda);

/* ---- halo exchange (tem_comm_module.fpp:549-646) ------------------------ */
void ora_comm_gather(double *buf, const double *state, const int32_t *pos, int n);
void ora_comm_scatter(double *state, const double *buf, const int32_t *pos, int n);

/* ---- diagnostics -------------------------------------------------------- */
double ora_total_mass(const double *state, int QQ, int nFluid);

#ifdef __cplusplus
}
#endif
#endif
