/* connectivity.c -- ORACLE (test infrastructure): restatement of
 * mus_construct_connectivity (mus/source/mus_connectivity_module.fpp:73-179)
 * for the default AOS + PULL build, and of the gather/scatter halves of
 * comm_isend_irecv_real (tem/source/tem_comm_module.fpp:549-646).
 */
#include "mus_oracle.h"
#include <stddef.h>

#define PRP_SOLID 2 /* tem/source/tem_property_module.f90: prp_solid = 2 */

void ora_construct_connectivity(int32_t *neigh, int nSize, int nElems, int QQ,
                                const int32_t *nghElems, const int64_t *property,
                                int nFluid, int haloOffset) {
  const int QQN = QQ - 1;
  const int *inv = ora_cxDirInv(QQ);
  /* rest direction (restPosition = QQ) pulls from itself, :118-124 */
  for (int e = 1; e <= nElems; ++e)
    neigh[(size_t)(QQ - 1) * nSize + (e - 1)] = (e - 1) * QQ + QQ;

  for (int e = 1; e <= nElems; ++e) {
    const int64_t elemProp = property[e - 1];
    for (int d = 1; d <= QQN; ++d) {
      const int nghDir = inv[d - 1]; /* NgDir for PULL = cxDirInv(iDir) */
      const int neighPos = nghElems[(size_t)(e - 1) * QQN + (nghDir - 1)];
      const int64_t neighProp = (neighPos > 0) ? property[neighPos - 1] : 0;
      const int missing_for_nonghost = (neighPos <= 0) && ((e <= nFluid) || (e > haloOffset));
      const int solidified = (int)(((neighProp >> PRP_SOLID) & 1) | ((elemProp >> PRP_SOLID) & 1));
      int sourceDir = d;
      if (missing_for_nonghost || solidified) sourceDir = inv[d - 1];
      int getFromPos = neighPos;
      if (neighPos <= 0 || solidified) getFromPos = e;
      neigh[(size_t)(d - 1) * nSize + (e - 1)] = (getFromPos - 1) * QQ + sourceDir;
    }
  }
}

void ora_comm_gather(double *buf, const double *state, const int32_t *pos, int n) {
  for (int i = 0; i < n; ++i) buf[i] = state[pos[i] - 1];
}

void ora_comm_scatter(double *state, const double *buf, const int32_t *pos, int n) {
  for (int i = 0; i < n; ++i) state[pos[i] - 1] = buf[i];
}
