/* treelm_box.h -- C interface of the synthetic treelm generator (libmusb200_mesh.so,
 * host/treelm_box.cpp): the arrays treelm + mus_construct hand to the solver for box meshes
 * (total list, property, nghElems, pdf%neigh, halo send/recv lists, boundary link lists), all
 * 1-based as the Fortran arrays hold them.  Host-side only, no CUDA. */
#ifndef MUSB200_TREELM_BOX_H
#define MUSB200_TREELM_BOX_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* kind: 0 periodic cube, 1 cavity (walls + moving lid), 2 channel (walls, velocity inlet,
 * pressure outlet); octants: the first 1, 2, 4 or 8 octants of the level's universe cube */
void *musb200_mesh_box_create(int level, int QQ, int kind, int rank, int nranks, int comm_reduced,
                              int octants);
void musb200_mesh_destroy(void *h);
/* info[12]: nFluid, nHalo, nElems, nSize, nBcElems, nBCs, nRecvProcs, nSendProcs, nRecvVals,
 * nSendVals, nRecvElems, nSendElems */
int musb200_mesh_info(void *h, int64_t *info);
int musb200_mesh_total(void *h, int64_t *out);
int musb200_mesh_property(void *h, int64_t *out);
int musb200_mesh_nghelems(void *h, int32_t *out);
int musb200_mesh_neigh(void *h, int32_t *out);
int musb200_mesh_bc_elembuffer(void *h, int32_t *out);
int musb200_mesh_comm(void *h, int dir, int32_t *proc, int32_t *nVals, int32_t *pos,
                      int32_t *elemCount, int32_t *elemPos);
int musb200_mesh_bc_info(void *h, int i, int32_t *sizes /* id, kind, nElems, nLinks */);
int musb200_mesh_bc_lists(void *h, int i, int32_t *elems, int32_t *links, int32_t *outPos,
                          int32_t *posInBuffer, int32_t *iDir);
int musb200_mesh_bc_elem_lists(void *h, int i, int32_t *normalInd, int32_t *posInBcElemBuf,
                               int32_t *neighPos, int32_t *iElemOfLink, int32_t *statePos);
int musb200_mesh_bary(void *h, double ox, double oy, double oz, double length, double *out);
#ifdef __cplusplus
}
#endif
#endif
