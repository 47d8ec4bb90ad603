/* musb200.h -- C ABI of libmusb200.so: the B200-native per-level LBM time step
 * behind Musubi's plugin surface.
 *
 * The reference (apes-suite/musubi, Fortran 2003 + MPI) has NO C ABI on this
 * path; its plugin points are Fortran procedure pointers with derived-type
 * arguments.  Every entry point below names the reference interface whose
 * pointee it replaces; the Fortran shim that unwraps the derived types and
 * calls these functions through ISO_C_BINDING is
 * musubi_b200/fortran/mus_b200_module.f90.
 *
 * Conventions
 *  - every function returns 0 on success and a non-zero code on failure; the
 *    message is available through musb200_last_error().  The shim calls
 *    tem_abort() on non-zero (tem/source/tem_aux_module.f90:457-478).
 *  - host arrays are BORROWED for the duration of the call (copied to the
 *    device); index lists are passed exactly as the Fortran arrays hold them:
 *    1-based, default-integer (int32), AOS state positions
 *        IDX(dir,elem)   = (elem-1)*nScalars + dir
 *        NGPOS(dir,elem) = (dir-1)*nSize + elem        (lbm_macros.inc:82,104)
 *  - one process = one MPI rank = one GPU; all calls come from the rank's main
 *    thread.
 *  - there is no CPU fallback: without a CUDA device every compute entry
 *    point fails with MUSB200_ERR_CUDA.
 */
#ifndef MUSB200_H
#define MUSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUSB200_OK            0
#define MUSB200_ERR_ARG       1   /* bad argument / unknown handle            */
#define MUSB200_ERR_CUDA      2   /* CUDA runtime error (message has detail)  */
#define MUSB200_ERR_NCCL      3
#define MUSB200_ERR_UNSUPPORTED 4 /* valid Musubi configuration outside the hot path */
#define MUSB200_ERR_CONNECTIVITY 5 /* neigh entry that is neither a plain pull nor a bounce-back */
#define MUSB200_ERR_STATE     6   /* call order violated (e.g. step before level_create) */

/* relaxation / kind / layout identifiers: mus_scheme_header_type
 * (mus/source/scheme/mus_scheme_header_module.f90) identify{kind,relaxation,layout} */
#define MUSB200_RELAX_BGK 0
#define MUSB200_RELAX_TRT 1
#define MUSB200_RELAX_MRT 2
#define MUSB200_KIND_FLUID 0
#define MUSB200_KIND_FLUID_INCOMPRESSIBLE 1
#define MUSB200_KIND_PASSIVE_SCALAR 2

/* boundary kinds: field%bc(i)%BC_kind (mus/source/bc/mus_bc_header_module.fpp) */
#define MUSB200_BC_WALL                 0  /* do_nothing, mus_bc_fluid_wall_module.fpp:407-450 */
#define MUSB200_BC_VELOCITY_BOUNCEBACK  1  /* mus_bc_fluid_module.fpp:1503-1597 */
#define MUSB200_BC_PRESSURE_ANTIBOUNCEBACK 2 /* mus_bc_fluid_module.fpp:2161-2353 */
#define MUSB200_BC_PRESSURE_EXPOL       3  /* mus_bc_fluid_module.fpp:1165-1362 */

/* comm buffer kinds: tem_levelDesc_type%{send,recv}buffer[FromCoarser|FromFiner]
 * (tem/source/tem_construction_module.f90:186-306) */
#define MUSB200_BUF_HALO        0
#define MUSB200_BUF_FROMCOARSER 1
#define MUSB200_BUF_FROMFINER   2
#define MUSB200_DIR_SEND 0
#define MUSB200_DIR_RECV 1

/* interpolation direction / order: mus_interpolation_type
 * (mus/source/intp/mus_interpolate_header_module.f90:103-128) */
#define MUSB200_INTP_FROMFINER   0  /* fillMineFromFiner   (average)            */
#define MUSB200_INTP_FROMCOARSER 1  /* fillFinerFromMe(order): 0 weighted avg, 1 linear, 2 quadratic */

/* ---- library life cycle ------------------------------------------------- */
/* Replaces nothing in the reference (MPI is initialised by tem_start); binds
 * this rank to `local_device` and, for nranks > 1, joins the NCCL communicator
 * described by the 128-byte unique id (rank 0 obtains it from
 * musb200_get_unique_id and broadcasts it with MPI_Bcast).                   */
int musb200_init(int rank, int nranks, int local_device, const void *nccl_unique_id);
int musb200_get_unique_id(void *out128);
int musb200_finalize(void);
int musb200_last_error(char *buf, int buflen);
int musb200_device_count(int *n);

/* ---- kernel selection ---------------------------------------------------
 * Mirrors mus_init_advRel_fluid / _fluid_incompressible
 * (mus/source/init/mus_initFluid_module.f90:102-232, 288-409): the Lua
 * `identify` strings select the pointee of scheme%compute.  Unknown or
 * out-of-scope combinations return MUSB200_ERR_UNSUPPORTED (the reference
 * calls tem_abort for unknown ones).                                          */
int musb200_scheme_select(const char *kind, const char *relaxation, const char *variant,
                          const char *layout, int *relax_id, int *kind_id, int *QQ);

/* ---- per-level data: pdf_data_type + tem_levelDesc_type ------------------
 * mus/source/mus_pdf_module.f90:55-102, tem_construction_module.f90:186-306.
 * neigh: pdf(level)%neigh(1:QQ*nSize); property / treeID: levelDesc%property,
 * levelDesc%total (may be NULL).                                              */
int musb200_level_create(int level, int QQ, int nScalars, int nAuxScalars, int nSize,
                         int nFluid, int nGhostFromCoarser, int nGhostFromFiner, int nHalo,
                         const int32_t *neigh, const int64_t *property, const int64_t *treeID);
/* treelm's predefined cube (`mesh = { predefined = 'cube', refinementLevel = treeLevel }`,
 * generate_treelm_cube, tem/source/treelmesh_module.f90:1224-1318) on ONE rank with the
 * connectivity of mus_construct_connectivity (mus_connectivity_module.fpp:113-177) generated on the
 * device: all 8^treeLevel elements in Morton order, fully periodic (walls = 0) or closed by walls
 * on the six faces (walls = 1).  The host form of the list ends at nSize*QQ < 2^31 (79 M elements
 * for d3q27); this entry point has no such limit, so one B200 holds BASELINE config 3 (512^3,
 * d3q27: 76 GB).  Equivalent to musb200_level_create with the list the host would build
 * (musb200_neigh_download returns exactly that list where it is representable).            */
int musb200_level_create_cube(int level, int treeLevel, int QQ, int walls);
/* mus_init_pdf with zero strain rate (mus/source/mus_flow_module.fpp:484-589): both state
 * buffers of every element <- f_eq(rho, u) of the auxField handed over by musb200_aux_upload   */
int musb200_state_init_equilibrium(int level);
int musb200_level_destroy(int level);
/* reads back the device neighbour list re-encoded as the Fortran positions
 * (bit-exact parity check of the index lists) */
int musb200_neigh_download(int level, int32_t *neigh);

/* state(level)%val(:, which), which = 1|2 (array2D_type, mus_scheme_type_module.f90:75-80) */
int musb200_state_upload(int level, int which, const double *aos_state);
int musb200_state_download(int level, int which, double *aos_state);
/* pdf%nNow / pdf%nNext (mus_pdf_module.f90:167-175) */
int musb200_set_now_next(int level, int nNow, int nNext);
int musb200_get_now_next(int level, int *nNow, int *nNext);
/* auxField(level)%val(1:nSize*nAuxScalars), AOS (elem-1)*nAuxScalars + {rho,ux,uy,uz} */
int musb200_aux_upload(int level, const double *aos_aux);
int musb200_aux_download(int level, double *aos_aux);
/* tracking of one element (tem_tracking with a point shape, interval {iter=1}):
 * reads auxField((elemPos-1)*4 + 1:4) of one element, 32 bytes device -> host */
int musb200_aux_probe(int level, int elemPos, double *rho_ux_uy_uz);
/* state(:, nNow) = state(:, nNext) on the device (mus_flow_module.fpp:181-185) */
int musb200_state_copy_next_to_now(int level);

/* fluid%viscKine%omLvl(level)%val, fluid%lambda, fluid%omegaBulkLvl(level)
 * (mus_relaxationParam_module.f90:249-283, mus_fluid_module.f90:104, 468-485).
 * omega == NULL: every element uses omega_uniform.                            */
int musb200_set_relaxation(int level, int relax_id, int kind_id, const double *omega,
                           double omega_uniform, double lambda, double omega_bulk);

/* fluid%viscKine%dataOnLvl(level)%val: lattice kinematic viscosity per element (or one
 * value); read by the ghost interpolation for the non-equilibrium rescaling
 * (mus_interpolate_average_module.fpp:328-337, mus_interpolate_linear_module.fpp:471-480) */
int musb200_set_viscosity(int level, const double *visc, double visc_uniform);

/* ---- restart bridge: mus_pdf_serialize / mus_pdf_unserialize ---------------
 * mus/source/mus_buffer_module.fpp:80-190, called chunk-wise by tem_restart_writeData /
 * tem_restart_readData (mus_restart_module.f90:89-248).  treeID / levelPointer: the chunk of
 * tree%treeID and tree%levelPointer (position of each element in its level's total list);
 * buffer(nElems*QQ): QQ PDFs of state(:, nNext) per element in treeID (space-filling-curve)
 * order -- byte for byte the payload of the reference's restart *.lsb file.                    */
int musb200_pdf_serialize(int nElems, const int64_t *treeID, const int32_t *levelPointer, double *buffer);
int musb200_pdf_unserialize(int nElems, const int64_t *treeID, const int32_t *levelPointer,
                            const double *buffer);

/* ---- source terms: field%source / globSrc with varname 'force' ------------
 * mus_source_module.f90:83-310 (element lists), :430-512 (mus_apply_sourceTerms);
 * order 2 (default): mus_addForceToAuxField_fluid / _fluidIncomp
 * (mus_auxFieldVar_module.fpp:1032-1214) + applySrc_force | applySrc_force_MRT_d3q19 |
 * applySrc_force_MRT_d3q27 (mus_derQuan_module.fpp:3129-3862, chosen by relaxation and layout
 * as mus_variable_module.f90:1043-1081 does); order 1: applySrc_force1stOrd (:4043-4143).
 * posInTotal  fun%elemLvl(level)%posInTotal(1:nElems)  (NULL: elements 1..nElems_solve)
 * force       forceField / fac%body_force: LATTICE units, 3 per element as the Fortran array
 *             holds them, or 3 values when uniform != 0
 * Both halves are fused into the sweep; a time-dependent force is handed over again before
 * the step it applies to.  order = 0 or nElems = 0 removes the source.                        */
int musb200_source_force(int level, int order, int nElems, const int32_t *posInTotal,
                         const double *force, int uniform);

/* ---- passive scalar (scheme kind 'passive_scalar', nAuxScalars = 1) --------
 * mus_init_advRel_lbm_ps (init/mus_initLBMPS_module.f90:59-159): relaxation bgk with variant
 * 1 = 'first' | 2 = 'second' (mus_compute_passiveScalar_module.fpp:77-279), trt = vStdNoOpt
 * (:293-398); diff_coeff = species%diff_coeff(1), lambda = species%lambda.                   */
int musb200_set_species(int level, int relax_id, int variant, double diff_coeff, double lambda);
/* scheme%transVar%method(1): transport velocity in LATTICE units (transVel * 1/fac%vel), one
 * triple per element 1..nElems_solve as get_valOfIndex returns them, or 3 values (uniform)   */
int musb200_set_transport_velocity(int level, int nElems, const double *vel, int uniform);
/* Several schemes on one mesh: every call addresses the levels of the bound slot (default 0).
 * musb200_couple_transport_velocity lets the passive scalar of the bound slot read its
 * transport velocity from the auxField of the flow scheme in `flow_slot` on the device
 * (the flow must have been stepped with auxField materialised: musb200_set_aux_every_step). */
int musb200_scheme_bind(int slot);
int musb200_couple_transport_velocity(int level, int flow_slot, int flow_level);

/* ---- boundaries: boundary_type + glob_boundary_type per level -------------
 * links      me%links(level)%val            (mus_bc_header_module.fpp:1702-1739)
 * outPos, posInBuffer, iDir: me%inletUbbQVal(level)  (:1876-1967)
 * bc_elemBuffer: levelDesc%bc_elemBuffer%val  (all BC elements of the level)   */
int musb200_bc_elembuffer(int level, int nBcElems, const int32_t *bc_elemBuffer);
int musb200_bc_register(int level, int bc_id, int bc_kind, int nLinks, const int32_t *links,
                        const int32_t *outPos, const int32_t *posInBuffer, const int32_t *iDir);
/* boundaries that read neighbours along the inward normal (pressure_expol: nNeighs = 2,
 * neighBufferPre_nNext; pressure_antibounceback: nNeighs = 1, neighBufferPost;
 * mus_bc_header_module.fpp:1036-1060), to be called after musb200_bc_register:
 * elemPos        globBC%elemLvl(level)%elem%val            (position in the total list)
 * posInBcElemBuf globBC%elemLvl(level)%posInBcElemBuf%val
 * normalInd      globBC%elemLvl(level)%normalInd%val       (mus_construction_module.fpp:2393-2440)
 * neighPos       fieldBC%neigh(level)%posInState(nNeighs, nElems), Fortran column-major
 *                (setFieldBCNeigh, mus_construction_module.fpp:1733-1900)
 * iElemOfLink    me%outletExpol(level)%iElem(nLinks)       (mus_bc_header_module.fpp:2256-2319);
 *                statePos(l) = iDir(l) + (iElem(l)-1)*QQ is implied                         */
int musb200_bc_register_elems(int level, int bc_id, int nElems, const int32_t *elemPos,
                              const int32_t *posInBcElemBuf, const int32_t *normalInd, int nNeighs,
                              const int32_t *neighPos, const int32_t *iElemOfLink);
/* boundary values in LATTICE units, evaluated by the host's spacetime function:
 * velocity boundaries 3 per link; pressure boundaries 1 per boundary element as lattice
 * density (pressure * cs2inv / fac%press, mus_bc_fluid_module.fpp:1270) */
int musb200_bc_set_values(int level, int bc_id, int nVals, const double *vals);
/* The copy runs on a dedicated stream and overlaps the level step in flight; the values become
 * current at the next musb200_step.  Pageable host memory is staged before the call returns;
 * pinned memory (musb200_host_alloc) must stay unchanged until that step has been issued and
 * the copy has completed (musb200_synchronize, or any call that reads back). */

/* ---- halo exchange: tem_communication_type --------------------------------
 * tem/source/tem_comm_module.fpp:93-177; pos(iProc) = buf_real(iProc)%pos.
 * proc: 0-based ranks; nVals[iProc]; pos: concatenated position lists.
 * In multi-level runs the elements of the HALO buffer travel with their four auxField entries
 * (auxField%sendBuffer, mus_auxField_module.f90:377-396), and so do the ghostFromFiner
 * elements of the FROMFINER buffer (mus_intpAuxFieldCoarserAndExchange, :404-444): the entries
 * are derived from the elements the positions belong to, no separate list is needed.
 * FROMCOARSER is exchanged after the level's own step and after do_intpCoarserAndExchange,
 * FROMFINER after do_intpFinerAndExchange (mus_control_module.f90:434-465, 861-1051).        */
int musb200_comm_register(int level, int buf_kind, int dir, int nProcs, const int32_t *proc,
                          const int32_t *nVals, const int32_t *pos);

/* ---- halo exchange through peer memory (ranks of one node, NVLink / NVSwitch) ----------
 * Optional replacement of pack -> ncclSend/ncclRecv -> unpack for the state halo buffer
 * (buf_kind HALO): ONE kernel per level step stores every communicated link directly into
 * the halo rows of the receiver's state array (CUDA IPC peer mapping) and completes the
 * exchange with arrival counters -- the MPI_Isend/Irecv/Waitall of comm_isend_irecv_real
 * (tem_comm_module.fpp:549-646) in one launch.  Set-up, once per level after
 * musb200_comm_register:
 *   1. every rank: musb200_p2p_export(level, blob)        -> MUSB200_P2P_BLOB bytes
 *   2. host: all-gather the blobs (MPI_Allgather); every receiver sends its recv-buffer
 *      position list buf_real(iProc)%pos to the rank it receives from (MPI_Sendrecv)
 *   3. every rank: musb200_p2p_connect(level, nProcs, proc, blobs of those procs in the
 *      order of the SEND buffer, nVals, remotePos = the receivers' position lists,
 *      concatenated in the same order)
 * Ranks must keep pdf%nNow/nNext in lockstep (they do: one swap per level step) and must
 * synchronise + barrier before musb200_level_destroy.  The NCCL path stays available
 * (musb200_p2p_enable(level, 0)) and is the one used across nodes.                       */
#define MUSB200_P2P_BLOB 512
int musb200_p2p_export(int level, void *blob);
int musb200_p2p_connect(int level, int nProcs, const int32_t *proc, const void *blobs,
                        const int32_t *nVals, const int32_t *remotePos);
int musb200_p2p_enable(int level, int flag);
/* 1: with the peer-memory exchange on, the sweep kernel itself stores the halo links of the
 * elements it has just collided into the receivers' state arrays (compute and transfer in one
 * kernel) and the exchange shrinks to the arrival handshake; 0 (default): a separate push kernel
 * after the sweep, whose entries are ordered for full 128-byte NVLink writes.  Identical results;
 * on 2 x B200 the fused form hides 16 us of exchange but its scattered 8-byte peer stores cost
 * the sweep 25 us (profiles/r01_multi_gpu.md), so it is opt-in. */
int musb200_set_fused_push(int flag);

/* ---- ghost interpolation: levelDesc%intpFromFiner / intpFromCoarser(order) -
 * targetList: positions of the target ghosts in the target level's total list;
 * per target a CSR row of source positions on the source level with weights
 * (average / weighted average) or an index into the least-square matrices
 * (mus_interpolate_header_module.f90:103-237, tem_matrix_module.fpp:161-425).  */
int musb200_intp_register(int tgtLevel, int direction, int order, int nTargets,
                          const int32_t *targetList, const int32_t *srcOffset /* nTargets+1 */,
                          const int32_t *srcPos, const double *weights,
                          const int32_t *posInMat, int nMatrices, const int32_t *matOffset,
                          const double *matrices, const double *childCoord /* 3 per target */);

/* ---- the time step: control%do_computation(minLevel) ----------------------
 * Runs nCoarseCycles iterations of do_fast_singleLevel / do_recursive_multiLevel
 * (mus/source/mus_control_module.f90:242-701) on the device: set_boundary, swap,
 * fused auxField + stream-collide, halo exchange, ghost interpolation.         */
int musb200_step(int minLevel, int maxLevel, int nCoarseCycles);
/* Several schemes on one mesh stepped TOGETHER (BASELINE config 5: a passive scalar transported by
 * the flow of another slot, musb200_couple_transport_velocity): within every level step of the
 * recursive schedule the schemes advance in the order given -- slots[0] first -- so that a scalar's
 * sweep reads the auxField the flow's sweep of the same level step has just written, also on the
 * finer levels' sub-steps; then each scheme fills its ghosts.  A passive scalar's ghosts are filled
 * by the reference's interpolation of arbitrary values applied to its PDFs
 * (fillArbiMyGhostsFromFiner_avg, fillArbiFinerGhostsFromMe_weighAvg / _linear / _quad,
 * mus_interpolate_average_module.fpp:95-185, 762-850, mus_interpolate_linear_module.fpp:124-205,
 * mus_interpolate_quadratic_module.fpp:102-190): the reference itself aborts for a passive scalar on
 * a multi-level mesh (mus_scheme_module.f90:166-190), this is the documented extension.         */
int musb200_step_schemes(int nSlots, const int *slots, int minLevel, int maxLevel, int nCoarseCycles);
/* 1: auxField is written by every level step;
 * 0 (default): only where the schedule reads it and on the last step of a call;
 * 2 (lazy): only where the schedule reads it -- musb200_aux_probe and musb200_aux_download
 *    compute the requested entries on demand from state(:, nNow), which still holds what the
 *    last step pulled from (tracking of a few elements every step then costs a one-thread
 *    kernel instead of 32 B of HBM writes per element and step) */
int musb200_set_aux_every_step(int flag);
/* mus_init_flow once the fluid PDFs are in state(:, nNext) -- after an initial condition or
 * mus_readRestart (mus/source/mus_flow_module.fpp:206-240): mus_initAuxField (auxField of the
 * fluid elements from their own PDFs, auxField halos, auxField of the ghostFromFiner elements),
 * fillHelperElementsFineToCoarse (:1517-1588: ghostFromFiner <- finer level, FromFiner and halo
 * exchange, finest level first) and fillHelperElementsCoarseToFine (:1601-1673: FromCoarser
 * exchange, ghostFromCoarser <- coarser level for every order).  Restart files and initial
 * conditions hold fluid elements only; without this call halo, ghost and auxField rows of a
 * freshly created level are zero.  Collective over the ranks.                               */
int musb200_fill_helper_elements(int minLevel, int maxLevel);

/* Single level on several ranks with the peer-memory exchange: 1 = the push of step n runs on a
 * second, high-priority stream WHILE step n+1 is swept; the CTAs of that sweep that pull from a halo
 * row (a bitmap built from the neighbour list: 3-12 % of the CTAs at 256^3 per GPU) are moved to the
 * END of the launch -- whole CTAs of consecutive elements, so nothing is lost in coalescing -- and
 * wait there for the peers' links.  0 (default): exchange strictly after compute as
 * comm_isend_irecv_real is called in do_fast_singleLevel (mus_control_module.f90:644-649).
 * Identical results either way; measured on 2 x B200 the overlapped form is 2 % slower (the
 * re-ordered CTAs lose the L2 locality of their neighbours), so it is opt-in. */
int musb200_set_overlap(int flag);
/* Peer-memory halo exchange, single level: 1 = the wait for the links of step n moves into the
 * sweep of step n+1, where only the CTAs that pull from a halo row wait (the same bitmap); 0
 * (default) = MPI_Waitall right after the push.  Identical results; measured on 2 and 8 B200 the
 * in-sweep wait is 1-2 % slower than the wait kernel (the push itself is what costs). */
int musb200_set_sweep_wait(int flag);
/* Every wait of the peer-memory exchange gives up after `seconds` (default 30; 0 = never): the
 * stream drains, and the next synchronising call (musb200_synchronize, musb200_reduce, ...) returns
 * MUSB200_ERR_NCCL naming the rank that did not deliver -- the shim then calls tem_abort as the
 * reference does when a rank fails (tem/source/tem_aux_module.f90:457-478).  The same calls poll
 * ncclCommGetAsyncError on the NCCL path. */
int musb200_set_exchange_timeout(double seconds);
/* 1 (default): a single-rank musb200_step call of 8 or more coarse cycles without per-stage
 * timers replays a CUDA graph of two coarse cycles (captured on first use, re-captured after
 * any call that changes a level); 0: every kernel is launched directly.  Same kernels, same
 * order, same results. */
int musb200_set_graphs(int flag);
/* 1 (default): when the only non-wall boundaries of a level are velocity_bounceback ones whose
 * links stay inside their own elements (the lists mus_set_inletUbb builds), fill_bcBuffer and the
 * link loop run as ONE kernel per boundary, one thread per boundary element, without the
 * bcBuffer snapshot; 0: always the reference's two phases (fill_bcBuffer, then the link loops).
 * Identical results; the fused form saves two launches and the snapshot traffic per step. */
int musb200_set_fused_bc(int flag);
/* 1 (default): the coarse -> fine interpolation (fillFinerGhostsFromMe_*, all orders) runs tile by
 * tile: consecutive targets whose sources -- the children of a parent share one neighbourhood --
 * fit a shared-memory tile are evaluated from one staged copy of those sources; 0: one thread per
 * (target, direction) gathering its sources from global memory.  Identical results. */
int musb200_set_intp_tiled(int flag);
int musb200_synchronize(void);

/* check_density / check_flow_status (mus_tools_module.f90:224-313):
 * sum of all PDFs of the fluid elements, max |u| and a NaN flag               */
int musb200_reduce(int level, double *total_mass, double *max_vel, int *any_nan);

/* ---- strict drop-in of the `kernel` interface -----------------------------
 * scheme%compute (mus_scheme_type_module.f90:204-235) with HOST arrays, one
 * call = H2D of inState, the fused kernel, D2H of outState and auxField.
 * Used by the unit-test style parity checks (mus/utests/mus_bgk_d3q19_compare_test.f90). */
int musb200_compute_host(int relax_id, int kind_id, int QQ, const double *inState,
                         double *outState, double *auxField, const int32_t *neigh,
                         int nElems /* = nSize */, int nSolve, const double *omega,
                         double lambda, double omega_bulk);

/* ---- timers: mus_timerHandles (mus_timer_module.f90) ---------------------- */
/* device time in ms accumulated since the last reset: compute, bc, comm, intp */
int musb200_timers(double *compute_ms, double *bc_ms, double *comm_ms, double *intp_ms);
int musb200_timers_reset(void);
/* number of kernel launches issued by this library since the last reset */
int musb200_launch_count(long long *n);
/* elapsed device time (ms) of the stepping stream between two marks */
int musb200_event_mark(int which /*0 start, 1 stop*/);
int musb200_event_elapsed(double *ms);
/* 1: bracket every stage with CUDA events (the reference's per-stage timers) */
int musb200_set_profiling(int flag);

/* pinned host memory for the state mirrors (the shim may use it for state/auxField) */
int musb200_host_alloc(size_t bytes, void **ptr);
int musb200_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* MUSB200_H */
