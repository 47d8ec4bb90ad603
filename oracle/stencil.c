/* stencil.c -- ORACLE (test infrastructure): direction tables, weights,
 * equilibrium functions and moment evaluation, restated from
 *   tem/source/tem_stencil_module.fpp:91-168       (cxDir, rest direction LAST)
 *   mus/source/scheme/mus_scheme_layout_module.f90:699-705 (weights)
 *   mus/source/scheme/mus_scheme_derived_quantities_type_module.f90
 *       get_sigma_d3q19 :495-526, get_pdfEq_d3q19 :532-580,
 *       get_sigma_d3q27 :640-682, get_pdfEq_d3q27 :688-745,
 *       get_vel_from_pdf_d3q19 :1000-1019, get_vel_from_pdf_d3q27 :1098-1120
 *   mus/source/derived/mus_auxFieldVar_module.fpp:605-817 (calcAuxField)
 */
#include "mus_oracle.h"
#include <stddef.h>

static const int cx19[19][3] = {
  {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1}, { 1, 0, 0}, { 0, 1, 0}, { 0, 0, 1},
  { 0,-1,-1}, { 0,-1, 1}, { 0, 1,-1}, { 0, 1, 1},
  {-1, 0,-1}, { 1, 0,-1}, {-1, 0, 1}, { 1, 0, 1},
  {-1,-1, 0}, {-1, 1, 0}, { 1,-1, 0}, { 1, 1, 0},
  { 0, 0, 0}};

static const int cx27[27][3] = {
  {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1}, { 1, 0, 0}, { 0, 1, 0}, { 0, 0, 1},
  { 0,-1,-1}, { 0,-1, 1}, { 0, 1,-1}, { 0, 1, 1},
  {-1, 0,-1}, { 1, 0,-1}, {-1, 0, 1}, { 1, 0, 1},
  {-1,-1, 0}, {-1, 1, 0}, { 1,-1, 0}, { 1, 1, 0},
  {-1,-1,-1}, {-1,-1, 1}, {-1, 1,-1}, {-1, 1, 1},
  { 1,-1,-1}, { 1,-1, 1}, { 1, 1,-1}, { 1, 1, 1},
  { 0, 0, 0}};

static int inv19[19], inv27[27];
static double w19[19], w27[27];
static int tables_ready = 0;

static void build_tables(void) {
  if (tables_ready) return;
  for (int i = 0; i < 19; ++i) {
    for (int j = 0; j < 19; ++j)
      if (cx19[i][0] == -cx19[j][0] && cx19[i][1] == -cx19[j][1] && cx19[i][2] == -cx19[j][2])
        inv19[i] = j + 1;
    int n = cx19[i][0] * cx19[i][0] + cx19[i][1] * cx19[i][1] + cx19[i][2] * cx19[i][2];
    w19[i] = (n == 0) ? 1.0 / 3.0 : (n == 1) ? 1.0 / 18.0 : 1.0 / 36.0;
  }
  for (int i = 0; i < 27; ++i) {
    for (int j = 0; j < 27; ++j)
      if (cx27[i][0] == -cx27[j][0] && cx27[i][1] == -cx27[j][1] && cx27[i][2] == -cx27[j][2])
        inv27[i] = j + 1;
    int n = cx27[i][0] * cx27[i][0] + cx27[i][1] * cx27[i][1] + cx27[i][2] * cx27[i][2];
    w27[i] = (n == 0) ? 8.0 / 27.0 : (n == 1) ? 2.0 / 27.0 : (n == 2) ? 1.0 / 54.0 : 1.0 / 216.0;
  }
  tables_ready = 1;
}

const int *ora_cxDir(int QQ) { return QQ == 19 ? &cx19[0][0] : QQ == 27 ? &cx27[0][0] : NULL; }
const int *ora_cxDirInv(int QQ) {
  build_tables();
  return QQ == 19 ? inv19 : QQ == 27 ? inv27 : NULL;
}
const double *ora_weights(int QQ) {
  build_tables();
  return QQ == 19 ? w19 : QQ == 27 ? w27 : NULL;
}

/* ------------------------------------------------------------------------ */
/* sigma vectors, 1-based like the reference (index 0 unused)                 */
static void sigma_d3q19(const double v[3], double s[23]) {
  s[22] = v[0] + v[1];
  s[21] = v[0] - v[1];
  s[20] = v[0] + v[2];
  s[19] = v[0] - v[2];
  s[18] = v[1] + v[2];
  s[17] = v[1] - v[2];
  s[16] = 3.0 * s[22];
  s[15] = 3.0 * s[21];
  s[14] = 3.0 * s[20];
  s[13] = 3.0 * s[19];
  s[12] = 3.0 * s[18];
  s[11] = 3.0 * s[17];
  s[10] = 4.5 * (s[22] * s[22]);
  s[9] = 4.5 * (s[21] * s[21]);
  s[8] = 4.5 * (s[20] * s[20]);
  s[7] = 4.5 * (s[19] * s[19]);
  s[6] = 4.5 * (s[18] * s[18]);
  s[5] = 4.5 * (s[17] * s[17]);
  s[4] = 4.5 * (v[0] * v[0]);
  s[3] = 4.5 * (v[1] * v[1]);
  s[2] = 4.5 * (v[2] * v[2]);
  s[1] = (1.0 / 3.0) * (s[2] + s[3] + s[4]);
}

static void pdfEq_d3q19(double rho, const double v[3], double *fEq /*0-based*/) {
  double s[23];
  sigma_d3q19(v, s);
  const double r18 = (1.0 / 18.0) * rho, r36 = (1.0 / 36.0) * rho;
  double *f = fEq - 1; /* 1-based view */
  f[1] = -r18 * (3.0 * v[0] - s[4] + s[1] - 1.0);
  f[2] = -r18 * (3.0 * v[1] - s[3] + s[1] - 1.0);
  f[3] = -r18 * (3.0 * v[2] - s[2] + s[1] - 1.0);
  f[4] = r18 * (3.0 * v[0] + s[4] - s[1] + 1.0);
  f[5] = r18 * (3.0 * v[1] + s[3] - s[1] + 1.0);
  f[6] = r18 * (3.0 * v[2] + s[2] - s[1] + 1.0);
  f[7] = r36 * (s[6] - s[12] - s[1] + 1.0);
  f[8] = r36 * (s[5] - s[11] - s[1] + 1.0);
  f[9] = r36 * (s[5] + s[11] - s[1] + 1.0);
  f[10] = r36 * (s[6] + s[12] - s[1] + 1.0);
  f[11] = r36 * (s[8] - s[14] - s[1] + 1.0);
  f[12] = r36 * (s[7] + s[13] - s[1] + 1.0);
  f[13] = r36 * (s[7] - s[13] - s[1] + 1.0);
  f[14] = r36 * (s[8] + s[14] - s[1] + 1.0);
  f[15] = r36 * (s[10] - s[16] - s[1] + 1.0);
  f[16] = r36 * (s[9] - s[15] - s[1] + 1.0);
  f[17] = r36 * (s[9] + s[15] - s[1] + 1.0);
  f[18] = r36 * (s[10] + s[16] - s[1] + 1.0);
  f[19] = -(1.0 / 3.0) * rho * (s[1] - 1.0);
}

static void sigma_d3q27(const double v[3], double s[35]) {
  s[34] = v[0] + v[1];
  s[33] = v[0] - v[1];
  s[32] = v[0] + v[2];
  s[31] = v[0] - v[2];
  s[30] = v[1] + v[2];
  s[29] = v[1] - v[2];
  s[28] = v[0] + v[1] + v[2];
  s[27] = v[0] + v[1] - v[2];
  s[26] = v[0] - v[1] + v[2];
  s[25] = v[1] - v[0] + v[2];
  for (int k = 0; k < 10; ++k) s[24 - k] = 3.0 * s[34 - k];
  for (int k = 0; k < 10; ++k) s[14 - k] = 4.5 * (s[34 - k] * s[34 - k]);
  s[4] = 4.5 * (v[0] * v[0]);
  s[3] = 4.5 * (v[1] * v[1]);
  s[2] = 4.5 * (v[2] * v[2]);
  s[1] = (1.0 / 3.0) * (s[2] + s[3] + s[4]);
}

static void pdfEq_d3q27(double rho, const double v[3], double *fEq) {
  double s[35];
  sigma_d3q27(v, s);
  const double r27 = (2.0 / 27.0) * rho, r54 = (1.0 / 54.0) * rho, r216 = (1.0 / 216.0) * rho;
  double *f = fEq - 1;
  f[1] = -r27 * (3.0 * v[0] - s[4] + s[1] - 1.0);
  f[2] = -r27 * (3.0 * v[1] - s[3] + s[1] - 1.0);
  f[3] = -r27 * (3.0 * v[2] - s[2] + s[1] - 1.0);
  f[4] = r27 * (3.0 * v[0] + s[4] - s[1] + 1.0);
  f[5] = r27 * (3.0 * v[1] + s[3] - s[1] + 1.0);
  f[6] = r27 * (3.0 * v[2] + s[2] - s[1] + 1.0);
  f[7] = r54 * (s[10] - s[20] - s[1] + 1.0);
  f[8] = r54 * (s[9] - s[19] - s[1] + 1.0);
  f[9] = r54 * (s[9] + s[19] - s[1] + 1.0);
  f[10] = r54 * (s[10] + s[20] - s[1] + 1.0);
  f[11] = r54 * (s[12] - s[22] - s[1] + 1.0);
  f[12] = r54 * (s[11] + s[21] - s[1] + 1.0);
  f[13] = r54 * (s[11] - s[21] - s[1] + 1.0);
  f[14] = r54 * (s[12] + s[22] - s[1] + 1.0);
  f[15] = r54 * (s[14] - s[24] - s[1] + 1.0);
  f[16] = r54 * (s[13] - s[23] - s[1] + 1.0);
  f[17] = r54 * (s[13] + s[23] - s[1] + 1.0);
  f[18] = r54 * (s[14] + s[24] - s[1] + 1.0);
  f[19] = -r216 * (s[18] - s[8] + s[1] - 1.0);
  f[20] = -r216 * (s[17] - s[7] + s[1] - 1.0);
  f[21] = -r216 * (s[16] - s[6] + s[1] - 1.0);
  f[22] = r216 * (s[15] + s[5] - s[1] + 1.0);
  f[23] = -r216 * (s[15] - s[5] + s[1] - 1.0);
  f[24] = r216 * (s[16] + s[6] - s[1] + 1.0);
  f[25] = r216 * (s[17] + s[7] - s[1] + 1.0);
  f[26] = r216 * (s[18] + s[8] - s[1] + 1.0);
  f[27] = -(8.0 / 27.0) * rho * (s[1] - 1.0);
}

void ora_pdfEq(int QQ, double rho, const double vel[3], double *fEq) {
  if (QQ == 19) pdfEq_d3q19(rho, vel, fEq);
  else pdfEq_d3q27(rho, vel, fEq);
}

/* ------------------------------------------------------------------------ */
/* incompressible equilibria, get_pdfEq_incomp_d3q19 :586-632 / _d3q27 :751-… :
 * fEq_i = w_i*rho + w_i*rho0*( 3 c.u + 4.5 (c.u)^2 - 1.5 u^2 ) in the
 * reference's sigma form: rho_div_w +/- rho0_div_w*( ... )                   */
static void pdfEq_incomp_d3q19(double rho, const double v[3], double *fEq) {
  double s[23];
  sigma_d3q19(v, s);
  const double rho0 = 1.0;
  const double r18 = (1.0 / 18.0) * rho, r36 = (1.0 / 36.0) * rho;
  const double z18 = (1.0 / 18.0) * rho0, z36 = (1.0 / 36.0) * rho0;
  double *f = fEq - 1;
  f[1] = r18 - z18 * (3.0 * v[0] - s[4] + s[1]);
  f[2] = r18 - z18 * (3.0 * v[1] - s[3] + s[1]);
  f[3] = r18 - z18 * (3.0 * v[2] - s[2] + s[1]);
  f[4] = r18 + z18 * (3.0 * v[0] + s[4] - s[1]);
  f[5] = r18 + z18 * (3.0 * v[1] + s[3] - s[1]);
  f[6] = r18 + z18 * (3.0 * v[2] + s[2] - s[1]);
  f[7] = r36 + z36 * (s[6] - s[12] - s[1]);
  f[8] = r36 + z36 * (s[5] - s[11] - s[1]);
  f[9] = r36 + z36 * (s[5] + s[11] - s[1]);
  f[10] = r36 + z36 * (s[6] + s[12] - s[1]);
  f[11] = r36 + z36 * (s[8] - s[14] - s[1]);
  f[12] = r36 + z36 * (s[7] + s[13] - s[1]);
  f[13] = r36 + z36 * (s[7] - s[13] - s[1]);
  f[14] = r36 + z36 * (s[8] + s[14] - s[1]);
  f[15] = r36 + z36 * (s[10] - s[16] - s[1]);
  f[16] = r36 + z36 * (s[9] - s[15] - s[1]);
  f[17] = r36 + z36 * (s[9] + s[15] - s[1]);
  f[18] = r36 + z36 * (s[10] + s[16] - s[1]);
  f[19] = (1.0 / 3.0) * rho - (1.0 / 3.0) * rho0 * s[1];
}

static void pdfEq_incomp_d3q27(double rho, const double v[3], double *fEq) {
  double s[35];
  sigma_d3q27(v, s);
  const double rho0 = 1.0;
  const double r27 = (2.0 / 27.0) * rho, r54 = (1.0 / 54.0) * rho, r216 = (1.0 / 216.0) * rho;
  const double z27 = (2.0 / 27.0) * rho0, z54 = (1.0 / 54.0) * rho0, z216 = (1.0 / 216.0) * rho0;
  double *f = fEq - 1;
  f[1] = r27 - z27 * (3.0 * v[0] - s[4] + s[1]);
  f[2] = r27 - z27 * (3.0 * v[1] - s[3] + s[1]);
  f[3] = r27 - z27 * (3.0 * v[2] - s[2] + s[1]);
  f[4] = r27 + z27 * (3.0 * v[0] + s[4] - s[1]);
  f[5] = r27 + z27 * (3.0 * v[1] + s[3] - s[1]);
  f[6] = r27 + z27 * (3.0 * v[2] + s[2] - s[1]);
  f[7] = r54 + z54 * (s[10] - s[20] - s[1]);
  f[8] = r54 + z54 * (s[9] - s[19] - s[1]);
  f[9] = r54 + z54 * (s[9] + s[19] - s[1]);
  f[10] = r54 + z54 * (s[10] + s[20] - s[1]);
  f[11] = r54 + z54 * (s[12] - s[22] - s[1]);
  f[12] = r54 + z54 * (s[11] + s[21] - s[1]);
  f[13] = r54 + z54 * (s[11] - s[21] - s[1]);
  f[14] = r54 + z54 * (s[12] + s[22] - s[1]);
  f[15] = r54 + z54 * (s[14] - s[24] - s[1]);
  f[16] = r54 + z54 * (s[13] - s[23] - s[1]);
  f[17] = r54 + z54 * (s[13] + s[23] - s[1]);
  f[18] = r54 + z54 * (s[14] + s[24] - s[1]);
  f[19] = r216 - z216 * (s[18] - s[8] + s[1]);
  f[20] = r216 - z216 * (s[17] - s[7] + s[1]);
  f[21] = r216 - z216 * (s[16] - s[6] + s[1]);
  f[22] = r216 + z216 * (s[15] + s[5] - s[1]);
  f[23] = r216 - z216 * (s[15] - s[5] + s[1]);
  f[24] = r216 + z216 * (s[16] + s[6] - s[1]);
  f[25] = r216 + z216 * (s[17] + s[7] - s[1]);
  f[26] = r216 + z216 * (s[18] + s[8] - s[1]);
  f[27] = (8.0 / 27.0) * rho - (8.0 / 27.0) * rho0 * s[1];
}

void ora_pdfEq_incomp(int QQ, double rho, const double vel[3], double *fEq) {
  if (QQ == 19) pdfEq_incomp_d3q19(rho, vel, fEq);
  else pdfEq_incomp_d3q27(rho, vel, fEq);
}

/* getNEq_acoustic (mus_derivedQuantities_module.fpp:441-478): non-equilibrium part from the
 * strain rate S(3,3) (Fortran column-major: S[j*3+i] = Sxx(i,j)), converted to the
 * post-collision form of the PULL build (convPrePost :584-595).                */
void ora_nEq_acoustic(int QQ, double omega, const double *S, double *nEq) {
  const double cs2 = 1.0 / 3.0, cs4inv = 9.0;
  const int *cx = ora_cxDir(QQ);
  const double *w = ora_weights(QQ);
  const double nu = cs2 * (1.0 / omega - 0.5);
  double tau[9];
  for (int k = 0; k < 9; ++k) tau[k] = 2.0 * nu * S[k];
  const double coeff = cs4inv / (2.0 - omega);
  for (int d = 0; d < QQ; ++d) {
    double acc = 0.0;
    for (int j = 0; j < 3; ++j) {
      for (int i = 0; i < 3; ++i)
        acc = acc + tau[j * 3 + i] * (double)cx[3 * d + i] * (double)cx[3 * d + j];
      acc = acc - cs2 * tau[j * 3 + j];
    }
    nEq[d] = -w[d] * acc * coeff;
  }
  if (omega != 1.0) {
    const double conv = 1.0 / (1.0 - omega);
    for (int d = 0; d < QQ; ++d) nEq[d] = nEq[d] / conv;
  }
}

/* ------------------------------------------------------------------------ */
/* velocity sums in the literal +/- order of get_vel_from_pdf_d3q19/_d3q27.   */
static void mom1_d3q19(const double *p /*1-based*/, double m[3]) {
  m[0] = p[4] - p[1] - p[11] + p[12] - p[13] + p[14] - p[15] - p[16] + p[17] + p[18];
  m[1] = p[5] - p[2] - p[7] - p[8] + p[9] + p[10] - p[15] + p[16] - p[17] + p[18];
  m[2] = p[6] - p[3] - p[7] + p[8] - p[9] + p[10] - p[11] - p[12] + p[13] + p[14];
}
static void mom1_d3q27(const double *p, double m[3]) {
  m[0] = p[4] - p[1] - p[11] + p[12] - p[13] + p[14] - p[15] - p[16] + p[17] + p[18]
       - p[19] - p[20] - p[21] - p[22] + p[23] + p[24] + p[25] + p[26];
  m[1] = p[5] - p[2] - p[7] - p[8] + p[9] + p[10] - p[15] + p[16] - p[17] + p[18]
       - p[19] - p[20] + p[21] + p[22] - p[23] - p[24] + p[25] + p[26];
  m[2] = p[6] - p[3] - p[7] + p[8] - p[9] + p[10] - p[11] - p[12] + p[13] + p[14]
       - p[19] + p[20] - p[21] + p[22] - p[23] + p[24] - p[25] + p[26];
}

/* sum_i c_i f_i in the literal order of get_vel_from_pdf_d3q19/_d3q27; p is 1-based */
void ora_first_moment(int QQ, const double *p, double m[3]) {
  if (QQ == 19) mom1_d3q19(p, m); else mom1_d3q27(p, m);
}

static void calc_aux(int QQ, int incomp, double *aux, const double *state,
                     const int32_t *neigh, int nSize, int nSolve) {
#pragma omp parallel for schedule(static) if (nSolve >= 20000)
  for (int e = 1; e <= nSolve; ++e) {
    double pdf[28];
    for (int d = 1; d <= QQ; ++d) pdf[d] = state[neigh[(size_t)(d - 1) * nSize + (e - 1)] - 1];
    double rho = 0.0;
    for (int d = 1; d <= QQ; ++d) rho = rho + pdf[d]; /* Fortran sum(): sequential */
    double m[3];
    if (QQ == 19) mom1_d3q19(pdf, m); else mom1_d3q27(pdf, m);
    double *a = aux + (size_t)(e - 1) * 4;
    a[0] = rho;
    if (incomp) { /* get_vel_from_pdf_*_incompressible: no division by rho (rho0 = 1) */
      a[1] = m[0]; a[2] = m[1]; a[3] = m[2];
    } else {
      a[1] = m[0] / rho; a[2] = m[1] / rho; a[3] = m[2] / rho;
    }
  }
}

void ora_calc_aux(int QQ, double *aux, const double *state, const int32_t *neigh,
                  int nSize, int nSolve) {
  calc_aux(QQ, 0, aux, state, neigh, nSize, nSolve);
}
void ora_calc_aux_incomp(int QQ, double *aux, const double *state, const int32_t *neigh,
                         int nSize, int nSolve) {
  calc_aux(QQ, 1, aux, state, neigh, nSize, nSolve);
}

void ora_update_omega(double *omega, const double *visc, int nSolve) {
  for (int e = 0; e < nSolve; ++e) omega[e] = 1.0 / (3.0 * visc[e] + 0.5);
}

double ora_omega_bulk(double viscBulkLat) {
  const double cs2 = 1.0 / 3.0;
  return 1.0 / (9.0 * viscBulkLat / (5.0 - 9.0 * cs2) + 0.5);
}

double ora_total_mass(const double *state, int QQ, int nFluid) {
  double tot = 0.0;
  for (int e = 0; e < nFluid; ++e)
    for (int d = 0; d < QQ; ++d) tot += state[(size_t)e * QQ + d];
  return tot;
}
