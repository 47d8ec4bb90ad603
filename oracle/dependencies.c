/* dependencies.c -- ORACLE (test infrastructure): the ghost dependency build, restated from
 *   tem_build_verticalDependencies          tem/source/tem_construction_module.f90:2894-2985
 *   tem_treeIDinTotal                       tem/source/tem_construction_module.f90:2317-2346
 *   mus_intp_update_depFromCoarser          mus/source/intp/mus_interpolate_module.fpp:544-833
 *   find_possIntpOrderAndUpdateMySources    mus/source/intp/mus_interpolate_module.fpp:841-947
 *   compute_weight ('linear_distance')      mus/source/intp/mus_interpolate_module.fpp:952-1011
 *   mus_set_nSources, init_cxDirWeightedAvg mus/source/intp/mus_interpolate_header_module.f90:390-607
 *   append_intpMatrixLSF, build_matrixLSF_* tem/source/tem_matrix_module.fpp:161-425, 464-523
 *   invert_matrix = DGETRF + DGETRI         tem/source/tem_matrix_module.fpp:610-660, with the
 *       unblocked paths of the vendored LAPACK (tem/external/lapack/dgetf2.f, dtrti2.f,
 *       dgetri.f: n = 4 or 10 is below every block size)
 * It is an implementation independent of the product's host-side generator
 * (musubi_b200/treelm_multilevel.py): the source lists, source directions, interpolation orders,
 * posInIntpMatLSF and child coordinates of the two are compared bit for bit, the weights and
 * least-square matrices to rounding (tests/test_oracle_dependencies.py).
 * "parity unpinned by reference fixtures": no golden of these lists exists in the reference tree.
 * All positions are 1-based as in Fortran.
 */
#include "mus_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- direction names, mus/source/mus_directions_module.f90:10-35 ---------------------------- */
enum {
  qN00 = 1, q0N0, q00N, q100, q010, q001, q0NN, q0N1, q01N, q011, qN0N, q10N, qN01, q101,
  qNN0, qN10, q1N0, q110, qNNN, qNN1, qN1N, qN11, q1NN, q1N1, q11N, q111
};
#define q000_19 19
#define q000_27 27

/* init_cxDirWeightedAvg, mus_interpolate_header_module.f90:568-607: [child][source] */
static const int wavg19[8][7] = {
    {q000_19, qN00, q0N0, q00N, qNN0, qN0N, q0NN}, {q000_19, q100, q0N0, q00N, q1N0, q10N, q0NN},
    {q000_19, qN00, q010, q00N, qN10, qN0N, q01N}, {q000_19, q100, q010, q00N, q110, q10N, q01N},
    {q000_19, qN00, q0N0, q001, qNN0, qN01, q0N1}, {q000_19, q100, q0N0, q001, q1N0, q101, q0N1},
    {q000_19, qN00, q010, q001, qN10, qN01, q011}, {q000_19, q100, q010, q001, q110, q101, q011}};
static const int wavg27[8][8] = {
    {q000_27, qN00, q0N0, q00N, qNN0, qN0N, q0NN, qNNN}, {q000_27, q100, q0N0, q00N, q1N0, q10N, q0NN, q1NN},
    {q000_27, qN00, q010, q00N, qN10, qN0N, q01N, qN1N}, {q000_27, q100, q010, q00N, q110, q10N, q01N, q11N},
    {q000_27, qN00, q0N0, q001, qNN0, qN01, q0N1, qNN1}, {q000_27, q100, q0N0, q001, q1N0, q101, q0N1, q1N1},
    {q000_27, qN00, q010, q001, qN10, qN01, q011, qN11}, {q000_27, q100, q010, q001, q110, q101, q011, q111}};

/* childPosition, tem/source/tem_param_module.f90:170-173 */
static const int childPosition[8][3] = {{-1, -1, -1}, {1, -1, -1}, {-1, 1, -1}, {1, 1, -1},
                                        {-1, -1, 1},  {1, -1, 1},  {-1, 1, 1},  {1, 1, 1}};

/* ---- tem_treeIDinTotal: the four blocks of the total list are sorted each ---------------------- */
/* blockStart[0..4]: 0-based start of fluid, ghostFromCoarser, ghostFromFiner, halo, end */
static int pos_in_total(int64_t tID, const int64_t *total, const int32_t *blockStart) {
  for (int b = 0; b < 4; ++b) {
    int lo = blockStart[b], hi = blockStart[b + 1] - 1;
    while (lo <= hi) {
      const int mid = lo + (hi - lo) / 2;
      if (total[mid] == tID) return mid + 1;
      if (total[mid] < tID) lo = mid + 1;
      else hi = mid - 1;
    }
  }
  return 0;
}

/* tem_build_verticalDependencies, first loop: parent, child number, relative coordinates */
void ora_vertical_dep_from_coarser(int nGhost, const int64_t *ghostID, const int64_t *cTotal,
                                   const int32_t *cBlockStart, int32_t *parentPos, int32_t *childNum,
                                   double *coord) {
  for (int i = 0; i < nGhost; ++i) {
    const int64_t parentID = (ghostID[i] - 1) / 8;             /* tem_parentOf */
    const int c = (int)(ghostID[i] - 8 * parentID);            /* tem_ChildNumber: 1..8 */
    childNum[i] = c;
    for (int k = 0; k < 3; ++k) coord[3 * i + k] = 0.25 * (double)childPosition[c - 1][k];
    parentPos[i] = pos_in_total(parentID, cTotal, cBlockStart);
  }
}

/* second loop: the children of a ghostFromFiner that exist on the finer level, in child order */
void ora_vertical_dep_from_finer(int nGhost, const int64_t *ghostID, const int64_t *fTotal,
                                 const int32_t *fBlockStart, int32_t *srcOffset, int32_t *srcPos) {
  srcOffset[0] = 0;
  for (int i = 0; i < nGhost; ++i) {
    int n = srcOffset[i];
    for (int c = 1; c <= 8; ++c) {                              /* tem_directChildren */
      const int p = pos_in_total(ghostID[i] * 8 + c, fTotal, fBlockStart);
      if (p > 0) srcPos[n++] = p;
    }
    srcOffset[i + 1] = n;
  }
}

/* ---- least-square-fit matrices ---------------------------------------------------------------- */
typedef struct {
  int order, nCoeffs, n, cap;
  int32_t *hashID, *invertible, *cols;
  double **A;                  /* nCoeffs x cols[i], row-major; 1 x 1 when not invertible */
} lsf_store;

void *ora_lsf_new(int order) {
  lsf_store *s = (lsf_store *)calloc(1, sizeof(lsf_store));
  s->order = order;
  s->nCoeffs = order == 1 ? 4 : 10;
  return s;
}
void ora_lsf_delete(void *h) {
  lsf_store *s = (lsf_store *)h;
  if (!s) return;
  for (int i = 0; i < s->n; ++i) free(s->A[i]);
  free(s->A); free(s->hashID); free(s->invertible); free(s->cols); free(s);
}
int ora_lsf_count(const void *h) { return ((const lsf_store *)h)->n; }
int ora_lsf_get(const void *h, int pos /* 1-based */, int32_t *hashID, int32_t *invertible,
                int32_t *rows, int32_t *cols, double *A) {
  const lsf_store *s = (const lsf_store *)h;
  if (pos < 1 || pos > s->n) return -1;
  const int i = pos - 1;
  *hashID = s->hashID[i]; *invertible = s->invertible[i];
  *rows = s->invertible[i] ? s->nCoeffs : 1;
  *cols = s->invertible[i] ? s->cols[i] : 1;
  if (A) memcpy(A, s->A[i], sizeof(double) * (size_t)(*rows) * (size_t)(*cols));
  return 0;
}

/* DGETF2: LU with partial pivoting, column-major a(n,n); returns info */
static int dgetf2(int n, double *a, int *ipiv) {
  int info = 0;
#define A_(i, j) a[(size_t)(j) * n + (i)]
  for (int j = 0; j < n; ++j) {
    int jp = j;                                                 /* IDAMAX: first maximum */
    double amax = fabs(A_(j, j));
    for (int i = j + 1; i < n; ++i)
      if (fabs(A_(i, j)) > amax) { amax = fabs(A_(i, j)); jp = i; }
    ipiv[j] = jp;
    if (A_(jp, j) != 0.0) {
      if (jp != j)
        for (int k = 0; k < n; ++k) { const double t = A_(j, k); A_(j, k) = A_(jp, k); A_(jp, k) = t; }
      if (j < n - 1) {
        if (fabs(A_(j, j)) >= 2.2250738585072014e-308) {        /* DLAMCH('S') */
          const double r = 1.0 / A_(j, j);
          for (int i = j + 1; i < n; ++i) A_(i, j) = A_(i, j) * r;
        } else {
          for (int i = j + 1; i < n; ++i) A_(i, j) = A_(i, j) / A_(j, j);
        }
      }
    } else if (info == 0) {
      info = j + 1;
    }
    if (j < n - 1)                                              /* DGER, alpha = -1 */
      for (int k = j + 1; k < n; ++k) {
        if (A_(j, k) == 0.0) continue;
        const double t = -1.0 * A_(j, k);
        for (int i = j + 1; i < n; ++i) A_(i, k) = A_(i, k) + A_(i, j) * t;
      }
  }
  return info;
}

/* DGETRI, unblocked path: DTRTRI('Upper','Non-unit') = singularity check + DTRTI2, then
 * inv(A)*L = inv(U) column by column (DGEMV), then the column interchanges */
static int dgetri(int n, double *a, const int *ipiv) {
  for (int j = 0; j < n; ++j)
    if (A_(j, j) == 0.0) return j + 1;
  for (int j = 0; j < n; ++j) {                                 /* DTRTI2, upper */
    A_(j, j) = 1.0 / A_(j, j);
    const double ajj = -A_(j, j);
    /* DTRMV('Upper','No transpose','Non-unit', j, A, x = A(1:j, j+1)) */
    for (int k = 0; k < j; ++k) {
      if (A_(k, j) == 0.0) continue;
      const double t = A_(k, j);
      for (int i = 0; i < k; ++i) A_(i, j) = A_(i, j) + t * A_(i, k);
      A_(k, j) = A_(k, j) * A_(k, k);
    }
    for (int i = 0; i < j; ++i) A_(i, j) = ajj * A_(i, j);      /* DSCAL */
  }
  double work[16];
  for (int j = n - 1; j >= 0; --j) {
    for (int i = j + 1; i < n; ++i) { work[i] = A_(i, j); A_(i, j) = 0.0; }
    if (j < n - 1)                                              /* DGEMV, alpha = -1, beta = 1 */
      for (int k = j + 1; k < n; ++k) {
        if (work[k] == 0.0) continue;
        const double t = -1.0 * work[k];
        for (int i = 0; i < n; ++i) A_(i, j) = A_(i, j) + t * A_(i, k);
      }
  }
  for (int j = n - 2; j >= 0; --j) {
    const int jp = ipiv[j];
    if (jp != j)
      for (int i = 0; i < n; ++i) { const double t = A_(i, j); A_(i, j) = A_(i, jp); A_(i, jp) = t; }
  }
#undef A_
  return 0;
}

/* append_intpMatrixLSF: returns success (1/0), *pos = 1-based position of the hash in the store */
int ora_lsf_append(void *h, int QQ, int nSources, const int32_t *neighDir, int32_t *pos) {
  lsf_store *s = (lsf_store *)h;
  const int *cx = ora_cxDir(QQ);
  int32_t hashID = 0;
  for (int i = 0; i < nSources; ++i) hashID |= (int32_t)(1u << neighDir[i]);   /* ibset(hashID, iNeigh) */
  for (int i = 0; i < s->n; ++i)
    if (s->hashID[i] == hashID) { *pos = i + 1; return s->invertible[i] ? 1 : 0; }
  if (s->n == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 16;
    s->hashID = (int32_t *)realloc(s->hashID, sizeof(int32_t) * (size_t)s->cap);
    s->invertible = (int32_t *)realloc(s->invertible, sizeof(int32_t) * (size_t)s->cap);
    s->cols = (int32_t *)realloc(s->cols, sizeof(int32_t) * (size_t)s->cap);
    s->A = (double **)realloc(s->A, sizeof(double *) * (size_t)s->cap);
  }
  const int nc = s->nCoeffs;
  /* tmp(iSrc, :) = polyLinear_3D / polyQuadratic_3D(cxDirRK(:, iDir)) */
  double *tmp = (double *)malloc(sizeof(double) * (size_t)nSources * (size_t)nc);
  for (int i = 0; i < nSources; ++i) {
    const int d = neighDir[i] - 1;
    const double x = (double)cx[3 * d], y = (double)cx[3 * d + 1], z = (double)cx[3 * d + 2];
    double *phi = tmp + (size_t)i * nc;
    phi[0] = 1.0; phi[1] = x; phi[2] = y; phi[3] = z;
    if (nc == 10) {
      phi[4] = x * x; phi[5] = y * y; phi[6] = z * z; phi[7] = x * y; phi[8] = y * z; phi[9] = z * x;
    }
  }
  double AtA[100];                                              /* column-major nc x nc */
  int ipiv[10];
  for (int c = 0; c < nc; ++c)
    for (int r = 0; r < nc; ++r) {
      double acc = 0.0;
      for (int i = 0; i < nSources; ++i) acc = acc + tmp[(size_t)i * nc + r] * tmp[(size_t)i * nc + c];
      AtA[c * nc + r] = acc;
    }
  int info = dgetf2(nc, AtA, ipiv);
  if (info == 0) info = dgetri(nc, AtA, ipiv);
  const int k = s->n++;
  s->hashID[k] = hashID;
  s->invertible[k] = info == 0;
  s->cols[k] = nSources;
  if (info == 0) {                                              /* me%A = matmul(inv_AtA, transpose(tmp)) */
    s->A[k] = (double *)malloc(sizeof(double) * (size_t)nc * (size_t)nSources);
    for (int r = 0; r < nc; ++r)
      for (int i = 0; i < nSources; ++i) {
        double acc = 0.0;
        for (int c = 0; c < nc; ++c) acc = acc + AtA[c * nc + r] * tmp[(size_t)i * nc + c];
        s->A[k][(size_t)r * nSources + i] = acc;
      }
  } else {
    s->A[k] = (double *)calloc(1, sizeof(double));
  }
  free(tmp);
  *pos = k + 1;
  return info == 0;
}

/* ---- mus_intp_update_depFromCoarser for one target level --------------------------------------- */
/* cNghElems: levelDesc(sourceLevel)%neigh(1)%nghElems as [cnElems][QQ-1] (row per element).
 * Per ghost i: order[i], nSrc[i], src[i][27], dir[i][27], posInMat[i] (1-based, 0 = none),
 * weights[i][27] (weighted average only).  lin / quad: the intpMat_forLSF stores of
 * fillFinerFromMe(linear / quadratic), shared by all levels.  Returns 0, or -1 when a ghost has
 * no source at all (the reference aborts). */
int ora_update_dep_from_coarser(int QQ, int orderMax, int nGhost, const int32_t *parentPos,
                                const int32_t *childNum, const double *coord,
                                const int32_t *cNghElems, int32_t *order, int32_t *nSrc,
                                int32_t *src, int32_t *dir, int32_t *posInMat, double *weights,
                                void *lin, void *quad) {
  const int QQN = QQ - 1;
  const int *cx = ora_cxDir(QQ);
  const int nMaxWavg = QQ == 19 ? 7 : 8;                       /* mus_set_nSources, 3-D d3q19 / d3q27 */
  const int nMin[3] = {1, 4, 10};                              /* nMinSources = nCoeffs */
  for (int i = 0; i < nGhost; ++i) {
    int32_t *mySources = src + (size_t)i * 27, *myNeighDir = dir + (size_t)i * 27;
    double *w = weights + (size_t)i * 27;
    int nFound = 0;
    for (int k = 0; k < 27; ++k) { mySources[k] = 0; myNeighDir[k] = -1; w[k] = 0.0; }
    posInMat[i] = 0;
    for (int iNeigh = 1; iNeigh <= QQ; ++iNeigh) {
      const int p = iNeigh == QQ ? parentPos[i]                /* restPosition = QQ */
                                 : cNghElems[(size_t)(parentPos[i] - 1) * QQN + (iNeigh - 1)];
      if (p > 0) { mySources[nFound] = p; myNeighDir[nFound] = iNeigh; ++nFound; }
    }
    if (nFound == 0) return -1;
    /* find_possIntpOrderAndUpdateMySources */
    int intpOrder = 0;
    for (int o = orderMax; o >= 1; --o)
      if (nFound >= nMin[o]) { intpOrder = o; break; }
    if (intpOrder == 0 || intpOrder == 1) {                     /* weightedAvgStencil%isActive */
      int32_t s2[27], d2[27];
      int n2 = 0;
      for (int k = 0; k < nFound; ++k) {
        int in = 0;
        for (int m = 0; m < nMaxWavg; ++m) {
          const int want = QQ == 19 ? wavg19[childNum[i] - 1][m] : wavg27[childNum[i] - 1][m];
          if (want == myNeighDir[k]) in = 1;
        }
        if (in) { s2[n2] = mySources[k]; d2[n2] = myNeighDir[k]; ++n2; }
      }
      if (n2 == nMaxWavg) {
        for (int k = 0; k < 27; ++k) { mySources[k] = 0; myNeighDir[k] = -1; }
        for (int k = 0; k < n2; ++k) { mySources[k] = s2[k]; myNeighDir[k] = d2[k]; }
        nFound = n2;
      }
    }
    int32_t pos = 0;
    if (intpOrder == 2) {
      if (ora_lsf_append(quad, QQ, nFound, myNeighDir, &pos)) posInMat[i] = pos;
      else intpOrder = 1;
    }
    if (intpOrder == 1) {
      if (ora_lsf_append(lin, QQ, nFound, myNeighDir, &pos)) posInMat[i] = pos;
      else intpOrder = 0;
    }
    if (intpOrder == 0) {                                       /* compute_weight, 'linear_distance' */
      double sum = 0.0;
      for (int k = 0; k < nFound; ++k) {
        const int d = myNeighDir[k] - 1;
        double dist[3];
        for (int c = 0; c < 3; ++c) dist[c] = fabs((double)cx[3 * d + c] - coord[3 * i + c]);
        w[k] = (1.0 - dist[0]) * (1.0 - dist[1]) * (1.0 - dist[2]);
      }
      for (int k = 0; k < nFound; ++k) sum = sum + w[k];        /* Fortran sum: sequential */
      for (int k = 0; k < nFound; ++k) w[k] = w[k] / sum;
    }
    order[i] = intpOrder;
    nSrc[i] = nFound;
  }
  return 0;
}
