!> mus_b200_module -- ISO_C_BINDING shim between Musubi's plugin surface and
!! libmusb200.so (include/musb200.h).
!!
!! NOT compiled in this repository's image (no Fortran compiler / MPI / CoCo);
!! delivered as the source a maintainer adds to mus/source/ (see INTEGRATION.md
!! for the three registration hunks).  Every routine conforms to one of the
!! reference's abstract interfaces and only unwraps derived types into the raw
!! arrays the C ABI takes:
!!
!!   mus_b200_compute      kernel interface      mus_scheme_type_module.f90:204-235
!!   do_b200               control routine       mus_control_module.f90:149-221
!!   mus_b200_upload/...   hand-over of state    mus_construction_module.fpp:500-660
!!
!! State ownership: after mus_b200_upload the device copy is authoritative;
!! state/auxField on the host are stale mirrors refreshed by mus_b200_download
!! at tracking / restart / check intervals.
module mus_b200_module
  use, intrinsic :: iso_c_binding

  use env_module,               only: rk, long_k
  use tem_aux_module,           only: tem_abort
  use tem_logging_module,       only: logUnit
  use tem_construction_module,  only: tem_levelDesc_type
  use tem_comm_module,          only: tem_communication_type
  use mus_scheme_type_module,   only: mus_scheme_type
  use mus_param_module,         only: mus_param_type
  use mus_pdf_module,           only: pdf_data_type

  implicit none
  private

  public :: mus_b200_init, mus_b200_finalize
  public :: mus_b200_upload, mus_b200_download
  public :: mus_b200_step, mus_b200_step_schemes, mus_b200_compute
  public :: mus_b200_check
  public :: mus_b200_upload_intp, mus_b200_set_force, mus_b200_p2p_connect
  public :: mus_b200_pdf_serialize, mus_b200_pdf_unserialize, mus_b200_fill_helper_elements
  public :: mus_b200_probe, mus_b200_track_every_step, mus_b200_cleanup, mus_b200_timers
  public :: mus_b200_bind_scheme, mus_b200_couple_transport_velocity
  public :: mus_b200_set_bc_values, mus_b200_set_species, mus_b200_set_transport_velocity

  integer(c_int), parameter :: buf_halo = 0, buf_fromCoarser = 1, buf_fromFiner = 2
  integer(c_int), parameter :: dir_send = 0, dir_recv = 1

  interface
    function musb200_init(rank, nranks, device, id) bind(C, name='musb200_init') result(rc)
      import :: c_int, c_ptr
      integer(c_int), value :: rank, nranks, device
      type(c_ptr), value :: id
      integer(c_int) :: rc
    end function
    function musb200_get_unique_id(id) bind(C, name='musb200_get_unique_id') result(rc)
      import :: c_int, c_char
      character(kind=c_char) :: id(128)
      integer(c_int) :: rc
    end function
    function musb200_finalize() bind(C, name='musb200_finalize') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function
    function musb200_last_error(buf, n) bind(C, name='musb200_last_error') result(rc)
      import :: c_int, c_char
      character(kind=c_char) :: buf(*)
      integer(c_int), value :: n
      integer(c_int) :: rc
    end function
    function musb200_scheme_select(kind, relaxation, variant, layout, relax_id, kind_id, QQ) &
      & bind(C, name='musb200_scheme_select') result(rc)
      import :: c_int, c_char
      character(kind=c_char) :: kind(*), relaxation(*), variant(*), layout(*)
      integer(c_int) :: relax_id, kind_id, QQ
      integer(c_int) :: rc
    end function
    function musb200_level_create(level, QQ, nScalars, nAuxScalars, nSize, nFluid, nGFC, nGFF, &
      & nHalo, neigh, property, treeID) bind(C, name='musb200_level_create') result(rc)
      import :: c_int, c_int32_t, c_int64_t
      integer(c_int), value :: level, QQ, nScalars, nAuxScalars, nSize, nFluid, nGFC, nGFF, nHalo
      integer(c_int32_t) :: neigh(*)
      integer(c_int64_t) :: property(*), treeID(*)
      integer(c_int) :: rc
    end function
    function musb200_state_upload(level, which, state) bind(C, name='musb200_state_upload') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level, which
      real(c_double) :: state(*)
      integer(c_int) :: rc
    end function
    function musb200_state_download(level, which, state) bind(C, name='musb200_state_download') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level, which
      real(c_double) :: state(*)
      integer(c_int) :: rc
    end function
    function musb200_aux_upload(level, aux) bind(C, name='musb200_aux_upload') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level
      real(c_double) :: aux(*)
      integer(c_int) :: rc
    end function
    function musb200_aux_download(level, aux) bind(C, name='musb200_aux_download') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level
      real(c_double) :: aux(*)
      integer(c_int) :: rc
    end function
    function musb200_set_now_next(level, nNow, nNext) bind(C, name='musb200_set_now_next') result(rc)
      import :: c_int
      integer(c_int), value :: level, nNow, nNext
      integer(c_int) :: rc
    end function
    function musb200_get_now_next(level, nNow, nNext) bind(C, name='musb200_get_now_next') result(rc)
      import :: c_int
      integer(c_int), value :: level
      integer(c_int) :: nNow, nNext
      integer(c_int) :: rc
    end function
    function musb200_set_relaxation(level, relax_id, kind_id, omega, omega_uniform, lambda, &
      & omega_bulk) bind(C, name='musb200_set_relaxation') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level, relax_id, kind_id
      real(c_double) :: omega(*)
      real(c_double), value :: omega_uniform, lambda, omega_bulk
      integer(c_int) :: rc
    end function
    function musb200_bc_elembuffer(level, n, elems) bind(C, name='musb200_bc_elembuffer') result(rc)
      import :: c_int, c_int32_t
      integer(c_int), value :: level, n
      integer(c_int32_t) :: elems(*)
      integer(c_int) :: rc
    end function
    function musb200_bc_register(level, bc_id, bc_kind, nLinks, links, outPos, posInBuffer, iDir) &
      & bind(C, name='musb200_bc_register') result(rc)
      import :: c_int, c_int32_t
      integer(c_int), value :: level, bc_id, bc_kind, nLinks
      integer(c_int32_t) :: links(*), outPos(*), posInBuffer(*), iDir(*)
      integer(c_int) :: rc
    end function
    function musb200_bc_set_values(level, bc_id, nVals, vals) bind(C, name='musb200_bc_set_values') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level, bc_id, nVals
      real(c_double) :: vals(*)
      integer(c_int) :: rc
    end function
    function musb200_comm_register(level, buf_kind, dir, nProcs, proc, nVals, pos) &
      & bind(C, name='musb200_comm_register') result(rc)
      import :: c_int, c_int32_t
      integer(c_int), value :: level, buf_kind, dir, nProcs
      integer(c_int32_t) :: proc(*), nVals(*), pos(*)
      integer(c_int) :: rc
    end function
    function musb200_step(minLevel, maxLevel, nCycles) bind(C, name='musb200_step') result(rc)
      import :: c_int
      integer(c_int), value :: minLevel, maxLevel, nCycles
      integer(c_int) :: rc
    end function
    function musb200_step_schemes(nSlots, slots, minLevel, maxLevel, nCycles) &
      & bind(C, name='musb200_step_schemes') result(rc)
      import :: c_int
      integer(c_int), value :: nSlots, minLevel, maxLevel, nCycles
      integer(c_int) :: slots(*)
      integer(c_int) :: rc
    end function
    function musb200_set_exchange_timeout(seconds) bind(C, name='musb200_set_exchange_timeout') result(rc)
      import :: c_int, c_double
      real(c_double), value :: seconds
      integer(c_int) :: rc
    end function
    function musb200_reduce(level, mass, maxvel, anynan) bind(C, name='musb200_reduce') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level
      real(c_double) :: mass, maxvel
      integer(c_int) :: anynan
      integer(c_int) :: rc
    end function
    function musb200_set_viscosity(level, visc, visc_uniform) bind(C, name='musb200_set_viscosity') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level
      real(c_double) :: visc(*)
      real(c_double), value :: visc_uniform
      integer(c_int) :: rc
    end function
    function musb200_bc_register_elems(level, bc_id, nElems, elemPos, posInBcElemBuf, normalInd, nNeighs, &
      & neighPos, iElemOfLink) bind(C, name='musb200_bc_register_elems') result(rc)
      import :: c_int, c_int32_t
      integer(c_int), value :: level, bc_id, nElems, nNeighs
      integer(c_int32_t) :: elemPos(*), posInBcElemBuf(*), normalInd(*), neighPos(*), iElemOfLink(*)
      integer(c_int) :: rc
    end function
    function musb200_intp_register(tgtLevel, direction, order, nTargets, targetList, srcOffset, srcPos, &
      & weights, posInMat, nMatrices, matOffset, matrices, childCoord) &
      & bind(C, name='musb200_intp_register') result(rc)
      import :: c_int, c_int32_t, c_double
      integer(c_int), value :: tgtLevel, direction, order, nTargets, nMatrices
      integer(c_int32_t) :: targetList(*), srcOffset(*), srcPos(*), posInMat(*), matOffset(*)
      real(c_double) :: weights(*), matrices(*), childCoord(*)
      integer(c_int) :: rc
    end function
    function musb200_source_force(level, order, nElems, posInTotal, force, uniform) &
      & bind(C, name='musb200_source_force') result(rc)
      import :: c_int, c_int32_t, c_double
      integer(c_int), value :: level, order, nElems, uniform
      integer(c_int32_t) :: posInTotal(*)
      real(c_double) :: force(*)
      integer(c_int) :: rc
    end function
    function musb200_set_species(level, relax_id, variant, diff_coeff, lambda) &
      & bind(C, name='musb200_set_species') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level, relax_id, variant
      real(c_double), value :: diff_coeff, lambda
      integer(c_int) :: rc
    end function
    function musb200_set_transport_velocity(level, nElems, vel, uniform) &
      & bind(C, name='musb200_set_transport_velocity') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level, nElems, uniform
      real(c_double) :: vel(*)
      integer(c_int) :: rc
    end function
    function musb200_level_destroy(level) bind(C, name='musb200_level_destroy') result(rc)
      import :: c_int
      integer(c_int), value :: level
      integer(c_int) :: rc
    end function
    function musb200_aux_probe(level, elemPos, rho_u) bind(C, name='musb200_aux_probe') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: level, elemPos
      real(c_double) :: rho_u(4)
      integer(c_int) :: rc
    end function
    function musb200_set_aux_every_step(flag) bind(C, name='musb200_set_aux_every_step') result(rc)
      import :: c_int
      integer(c_int), value :: flag
      integer(c_int) :: rc
    end function
    function musb200_scheme_bind(slot) bind(C, name='musb200_scheme_bind') result(rc)
      import :: c_int
      integer(c_int), value :: slot
      integer(c_int) :: rc
    end function
    function musb200_couple_transport_velocity(level, flow_slot, flow_level) &
      & bind(C, name='musb200_couple_transport_velocity') result(rc)
      import :: c_int
      integer(c_int), value :: level, flow_slot, flow_level
      integer(c_int) :: rc
    end function
    function musb200_synchronize() bind(C, name='musb200_synchronize') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function
    function musb200_timers(compute_ms, bc_ms, comm_ms, intp_ms) bind(C, name='musb200_timers') result(rc)
      import :: c_int, c_double
      real(c_double) :: compute_ms, bc_ms, comm_ms, intp_ms
      integer(c_int) :: rc
    end function
    function musb200_timers_reset() bind(C, name='musb200_timers_reset') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function
    function musb200_pdf_serialize(nElems, treeID, levelPointer, buffer) &
      & bind(C, name='musb200_pdf_serialize') result(rc)
      import :: c_int, c_int32_t, c_int64_t, c_double
      integer(c_int), value :: nElems
      integer(c_int64_t) :: treeID(*)
      integer(c_int32_t) :: levelPointer(*)
      real(c_double) :: buffer(*)
      integer(c_int) :: rc
    end function
    function musb200_pdf_unserialize(nElems, treeID, levelPointer, buffer) &
      & bind(C, name='musb200_pdf_unserialize') result(rc)
      import :: c_int, c_int32_t, c_int64_t, c_double
      integer(c_int), value :: nElems
      integer(c_int64_t) :: treeID(*)
      integer(c_int32_t) :: levelPointer(*)
      real(c_double) :: buffer(*)
      integer(c_int) :: rc
    end function
    function musb200_fill_helper_elements(minLevel, maxLevel) &
      & bind(C, name='musb200_fill_helper_elements') result(rc)
      import :: c_int
      integer(c_int), value :: minLevel, maxLevel
      integer(c_int) :: rc
    end function
    function musb200_p2p_export(level, blob) bind(C, name='musb200_p2p_export') result(rc)
      import :: c_int, c_char
      integer(c_int), value :: level
      character(kind=c_char) :: blob(*)
      integer(c_int) :: rc
    end function
    function musb200_p2p_connect(level, nProcs, proc, blobs, nVals, remotePos) &
      & bind(C, name='musb200_p2p_connect') result(rc)
      import :: c_int, c_int32_t, c_char
      integer(c_int), value :: level, nProcs
      integer(c_int32_t) :: proc(*), nVals(*), remotePos(*)
      character(kind=c_char) :: blobs(*)
      integer(c_int) :: rc
    end function
    function musb200_compute_host(relax_id, kind_id, QQ, inState, outState, auxField, neigh, &
      & nElems, nSolve, omega, lambda, omega_bulk) bind(C, name='musb200_compute_host') result(rc)
      import :: c_int, c_int32_t, c_double
      integer(c_int), value :: relax_id, kind_id, QQ, nElems, nSolve
      real(c_double) :: inState(*), outState(*), auxField(*), omega(*)
      integer(c_int32_t) :: neigh(*)
      real(c_double), value :: lambda, omega_bulk
      integer(c_int) :: rc
    end function
  end interface

  integer(c_int), save :: relax_id = 0, kind_id = 0, QQ_id = 19
  !> MUSB200_P2P_BLOB of include/musb200.h
  integer, parameter :: p2p_blob = 512

contains

  !> tem_abort with the library's message on a non-zero return code
  subroutine chk(rc, where)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char) :: buf(512)
    character(len=512) :: msg
    integer :: i
    if (rc == 0) return
    i = musb200_last_error(buf, 512_c_int)
    msg = ''
    do i = 1, 512
      if (buf(i) == c_null_char) exit
      msg(i:i) = buf(i)
    end do
    call tem_abort('libmusb200 ('//trim(where)//'): '//trim(msg))
  end subroutine chk

  !> one rank = one GPU; rank 0 creates the NCCL id and broadcasts it with MPI
  subroutine mus_b200_init(params)
    use mpi
    type(mus_param_type), intent(in) :: params
    character(kind=c_char), target :: id(128)
    integer :: iError, localComm, localRank
    if (params%general%proc%rank == 0) call chk(musb200_get_unique_id(id), 'get_unique_id')
    call mpi_bcast(id, 128, mpi_character, 0, params%general%proc%comm, iError)
    call mpi_comm_split_type(params%general%proc%comm, mpi_comm_type_shared, 0, mpi_info_null, &
      &                      localComm, iError)
    call mpi_comm_rank(localComm, localRank, iError)
    call chk(musb200_init(int(params%general%proc%rank, c_int),      &
      &                   int(params%general%proc%comm_size, c_int), &
      &                   int(localRank, c_int), c_loc(id)), 'init')
    ! a rank that dies must not hang the others: the waits of the halo exchange give up after this
    ! many seconds and the next synchronising call reports the silent rank (chk -> tem_abort)
    call chk(musb200_set_exchange_timeout(60.0_c_double), 'set_exchange_timeout')
  end subroutine mus_b200_init

  subroutine mus_b200_finalize()
    call chk(musb200_finalize(), 'finalize')
  end subroutine mus_b200_finalize

  !> hand pdf(level), levelDesc(level), state, auxField, omega, BC and comm lists over
  !! (called once after mus_init_flow / mus_init_boundary, mus_program_module.fpp:118-238)
  subroutine mus_b200_upload(scheme, params, minLevel, maxLevel)
    type(mus_scheme_type), intent(inout) :: scheme
    type(mus_param_type), intent(in) :: params
    integer, intent(in) :: minLevel, maxLevel
    integer :: iLevel, iBnd, nBCs
    character(len=64) :: variant
    variant = trim(scheme%header%relaxHeader%variant)
    if (variant == 'b200') variant = 'standard'
    call chk(musb200_scheme_select(trim(scheme%header%kind)//c_null_char,       &
      &        trim(scheme%header%relaxation)//c_null_char,                      &
      &        trim(variant)//c_null_char, trim(scheme%header%layout)//c_null_char, &
      &        relax_id, kind_id, QQ_id), 'scheme_select')
    do iLevel = minLevel, maxLevel
      associate(pdf => scheme%pdf(iLevel), ld => scheme%levelDesc(iLevel), &
        &       fluid => scheme%field(1)%fieldProp%fluid)
        call chk(musb200_level_create(int(iLevel, c_int), int(QQ_id, c_int),                 &
          &   int(scheme%varSys%nScalars, c_int), int(scheme%varSys%nAuxScalars, c_int),      &
          &   int(pdf%nSize, c_int), int(pdf%nElems_fluid, c_int),                            &
          &   int(pdf%nElems_ghostFromCoarser, c_int), int(pdf%nElems_ghostFromFiner, c_int), &
          &   int(pdf%nElems_halo, c_int), pdf%neigh, ld%property, ld%total), 'level_create')
        call chk(musb200_state_upload(int(iLevel, c_int), 1_c_int, scheme%state(iLevel)%val(:,1)), 'state')
        call chk(musb200_state_upload(int(iLevel, c_int), 2_c_int, scheme%state(iLevel)%val(:,2)), 'state')
        call chk(musb200_set_now_next(int(iLevel, c_int), int(pdf%nNow, c_int), int(pdf%nNext, c_int)), 'now')
        call chk(musb200_aux_upload(int(iLevel, c_int), scheme%auxField(iLevel)%val), 'aux')
        call chk(musb200_set_relaxation(int(iLevel, c_int), relax_id, kind_id,            &
          &   fluid%viscKine%omLvl(iLevel)%val, 0.0_c_double, real(fluid%lambda, c_double), &
          &   real(fluid%omegaBulkLvl(iLevel), c_double)), 'relaxation')
        call upload_comm(iLevel, buf_halo, ld%sendBuffer, ld%recvBuffer)
        call upload_comm(iLevel, buf_fromCoarser, ld%sendBufferFromCoarser, ld%recvBufferFromCoarser)
        call upload_comm(iLevel, buf_fromFiner, ld%sendBufferFromFiner, ld%recvBufferFromFiner)
        if (ld%bc_elemBuffer%nVals > 0) then
          call chk(musb200_bc_elembuffer(int(iLevel, c_int), int(ld%bc_elemBuffer%nVals, c_int), &
            &      ld%bc_elemBuffer%val), 'bc_elembuffer')
        end if
      end associate
      nBCs = size(scheme%field(1)%bc)
      do iBnd = 1, nBCs
        call upload_bc(scheme, iLevel, iBnd)
      end do
    end do
  end subroutine mus_b200_upload

  subroutine upload_comm(iLevel, bufKind, send, recv)
    integer, intent(in) :: iLevel
    integer(c_int), intent(in) :: bufKind
    type(tem_communication_type), intent(in) :: send, recv
    call one(send, dir_send)
    call one(recv, dir_recv)
  contains
    subroutine one(c, dir)
      type(tem_communication_type), intent(in) :: c
      integer(c_int), intent(in) :: dir
      integer(c_int32_t), allocatable :: proc(:), nVals(:), pos(:)
      integer :: iProc, n
      if (c%nProcs == 0) return
      allocate(proc(c%nProcs), nVals(c%nProcs))
      n = 0
      do iProc = 1, c%nProcs
        proc(iProc) = c%proc(iProc)
        nVals(iProc) = c%buf_real(iProc)%nVals
        n = n + nVals(iProc)
      end do
      allocate(pos(n))
      n = 0
      do iProc = 1, c%nProcs
        pos(n+1:n+nVals(iProc)) = c%buf_real(iProc)%pos(1:nVals(iProc))
        n = n + nVals(iProc)
      end do
      call chk(musb200_comm_register(int(iLevel, c_int), bufKind, dir, int(c%nProcs, c_int), &
        &                            proc, nVals, pos), 'comm_register')
    end subroutine one
  end subroutine upload_comm

  subroutine upload_bc(scheme, iLevel, iBnd)
    type(mus_scheme_type), intent(inout) :: scheme
    integer, intent(in) :: iLevel, iBnd
    integer(c_int) :: kind
    integer(c_int32_t), allocatable :: iDir(:)
    integer :: iLink, nLinks
    associate(bc => scheme%field(1)%bc(iBnd))
      select case (trim(bc%BC_kind))
      case ('wall');                kind = 0
      case ('velocity_bounceback'); kind = 1
      case ('pressure_antibounceback'); kind = 2
      case ('pressure_expol');      kind = 3
      case default
        call tem_abort('boundary kind "'//trim(bc%BC_kind)//'" is outside the B200 hot path')
      end select
      if (kind == 0) then
        call chk(musb200_bc_register(int(iLevel, c_int), int(iBnd, c_int), kind, 0_c_int, &
          &      bc%links(iLevel)%val, bc%links(iLevel)%val, bc%links(iLevel)%val,        &
          &      bc%links(iLevel)%val), 'bc_register')
      else if (kind == 1) then
        ! velocity_bounceback: the lists of mus_set_inletUbb (mus_bc_header_module.fpp:1876-1967)
        call chk(musb200_bc_register(int(iLevel, c_int), int(iBnd, c_int), kind,            &
          &      int(bc%links(iLevel)%nVals, c_int), bc%links(iLevel)%val,                   &
          &      bc%inletUbbQVal(iLevel)%outPos, bc%inletUbbQVal(iLevel)%posInBuffer,        &
          &      bc%inletUbbQVal(iLevel)%iDir), 'bc_register')
      else
        ! pressure boundaries carry mus_set_outletExpol's lists instead (:2256-2319):
        ! statePos(l) = iDir(l) + (iElem(l)-1)*QQ; outPos / posInBuffer are not read for these kinds
        nLinks = bc%links(iLevel)%nVals
        allocate(iDir(max(nLinks, 1)))
        do iLink = 1, nLinks
          iDir(iLink) = bc%outletExpol(iLevel)%statePos(iLink) &
            &         - (bc%outletExpol(iLevel)%iElem(iLink) - 1) * scheme%layout%fStencil%QQ
        end do
        call chk(musb200_bc_register(int(iLevel, c_int), int(iBnd, c_int), kind, int(nLinks, c_int), &
          &      bc%links(iLevel)%val, bc%links(iLevel)%val, bc%links(iLevel)%val, iDir), 'bc_register')
        deallocate(iDir)
        ! boundaries reading neighbours along the inward normal (mus_bc_header_module.fpp:1036-1060)
        associate(gbc => scheme%globBC(iBnd))
          call chk(musb200_bc_register_elems(int(iLevel, c_int), int(iBnd, c_int),              &
            &      int(gbc%nElems(iLevel), c_int), gbc%elemLvl(iLevel)%elem%val,                 &
            &      gbc%elemLvl(iLevel)%posInBcElemBuf%val, gbc%elemLvl(iLevel)%normalInd%val,    &
            &      int(bc%nNeighs, c_int), bc%neigh(iLevel)%posInState,                          &
            &      bc%outletExpol(iLevel)%iElem), 'bc_register_elems')
        end associate
      end if
    end associate
  end subroutine upload_bc

  !> boundary values of the non-wall boundaries of one level, evaluated exactly as the
  !! reference's boundary routines do at the top of every call -- the boundary's space-time
  !! function through get_valOfIndex on its pntIndex list -- converted to lattice units and
  !! handed to the device (double-buffered on a copy stream; current at the next mus_b200_step):
  !!   velocity_bounceback      vel_b(3 per link) * 1/fac%vel    mus_bc_fluid_module.fpp:1536-1575
  !!   pressure_expol / _antiBB rho_b(per BC element) * cs2inv / fac%press        :1249-1261
  !! Call once after mus_b200_upload for constant boundaries, before every mus_b200_step
  !! (nCycles = 1) for time-dependent ones.
  subroutine mus_b200_set_bc_values(scheme, iLevel, params, time)
    use tem_time_module, only: tem_time_type
    use tem_param_module, only: cs2inv
    type(mus_scheme_type), intent(in) :: scheme
    integer, intent(in) :: iLevel
    type(mus_param_type), intent(in) :: params
    type(tem_time_type), intent(in) :: time
    real(kind=rk), allocatable :: vals(:)
    integer :: iBnd, nVals, varPos
    do iBnd = 1, size(scheme%field(1)%bc)
      associate(bc => scheme%field(1)%bc(iBnd))
        select case (trim(bc%BC_kind))
        case ('velocity_bounceback')
          nVals = bc%links(iLevel)%nVals
          if (nVals == 0) cycle
          allocate(vals(3*nVals))
          varPos = bc%BC_states%velocity%varPos
          call scheme%varSys%method%val(varPos)%get_valOfIndex(                  &
            & varSys = scheme%varSys, time = time, iLevel = iLevel,               &
            & idx    = bc%BC_states%velocity%pntIndex%indexLvl(iLevel)%val(1:nVals), &
            & nVals  = nVals, res = vals )
          vals = vals * (1.0_rk / params%physics%fac(iLevel)%vel)
          call chk(musb200_bc_set_values(int(iLevel, c_int), int(iBnd, c_int),    &
            &      int(3*nVals, c_int), vals), 'bc_set_values')
          deallocate(vals)
        case ('pressure_expol', 'pressure_antibounceback')
          nVals = scheme%globBC(iBnd)%nElems(iLevel)
          if (nVals == 0) cycle
          allocate(vals(nVals))
          varPos = bc%BC_states%pressure%varPos
          call scheme%varSys%method%val(varPos)%get_valOfIndex(                  &
            & varSys = scheme%varSys, time = time, iLevel = iLevel,               &
            & idx    = bc%BC_states%pressure%pntIndex%indexLvl(iLevel)%val(1:nVals), &
            & nVals  = nVals, res = vals )
          vals = vals * (1.0_rk / params%physics%fac(iLevel)%press * cs2inv)
          call chk(musb200_bc_set_values(int(iLevel, c_int), int(iBnd, c_int),    &
            &      int(nVals, c_int), vals), 'bc_set_values')
          deallocate(vals)
        end select
      end associate
    end do
  end subroutine mus_b200_set_bc_values

  !> ghost interpolation: flatten depFromFiner / depFromCoarser of the target level into the CSR
  !! lists the library takes (tem_construction_module.f90:160-276; least-square matrices
  !! tem_matrix_module.fpp:75-96).  direction 0: fillMineFromFiner, 1: fillFinerFromMe(order)
  subroutine mus_b200_upload_intp(scheme, iLevel)
    type(mus_scheme_type), intent(in) :: scheme
    integer, intent(in) :: iLevel    !< TARGET level
    integer :: iOrder
    associate(ld => scheme%levelDesc(iLevel), intp => scheme%intp)
      if (ld%intpFromFiner%nVals > 0) call one(ld%intpFromFiner%val, ld%intpFromFiner%nVals, &
        &                                     0_c_int, 0_c_int, .true.)
      if (allocated(ld%intpFromCoarser)) then
        do iOrder = 0, intp%config%order
          if (ld%intpFromCoarser(iOrder)%nVals > 0) &
            & call one(ld%intpFromCoarser(iOrder)%val, ld%intpFromCoarser(iOrder)%nVals, &
            &          1_c_int, int(iOrder, c_int), .false.)
        end do
      end if
      ! fluid%viscKine%dataOnLvl(iLevel)%val: target viscosity for the f_neq rescaling
      call chk(musb200_set_viscosity(int(iLevel, c_int),                                  &
        &      scheme%field(1)%fieldProp%fluid%viscKine%dataOnLvl(iLevel)%val, 0.0_c_double), 'viscosity')
    end associate
  contains
    subroutine one(list, nTargets, direction, order, fromFiner)
      integer, intent(in) :: list(:), nTargets
      integer(c_int), intent(in) :: direction, order
      logical, intent(in) :: fromFiner
      integer(c_int32_t), allocatable :: tgt(:), off(:), src(:), pim(:), moff(:)
      real(c_double), allocatable :: wgt(:), mats(:), coord(:)
      integer :: i, k, n, nMat, nTot, offs
      associate(ld => scheme%levelDesc(iLevel))
        allocate(tgt(nTargets), off(nTargets+1), pim(nTargets), coord(3*nTargets))
        n = 0
        do i = 1, nTargets
          if (fromFiner) then
            n = n + ld%depFromFiner(list(i))%elem%nVals
          else
            n = n + ld%depFromCoarser(list(i))%elem%nVals
          end if
        end do
        allocate(src(n), wgt(n))
        n = 0; off(1) = 0
        do i = 1, nTargets
          if (fromFiner) then
            associate(dep => ld%depFromFiner(list(i)))
              tgt(i) = ld%offset(1, 3) + list(i)      ! eT_ghostFromFiner
              do k = 1, dep%elem%nVals
                src(n+k) = dep%elem%val(k); wgt(n+k) = 1.0_c_double / dep%elem%nVals
              end do
              n = n + dep%elem%nVals; pim(i) = 0; coord(3*i-2:3*i) = 0.0_c_double
            end associate
          else
            associate(dep => ld%depFromCoarser(list(i)))
              tgt(i) = ld%offset(1, 2) + list(i)      ! eT_ghostFromCoarser
              do k = 1, dep%elem%nVals
                src(n+k) = dep%elem%val(k)
                if (allocated(dep%weight)) wgt(n+k) = dep%weight(k)
              end do
              n = n + dep%elem%nVals; pim(i) = dep%posInIntpMatLSF; coord(3*i-2:3*i) = dep%coord
            end associate
          end if
          off(i+1) = n
        end do
        nMat = 0; nTot = 0
        if (.not. fromFiner .and. order > 0) then
          associate(lsf => scheme%intp%fillFinerFromME(order)%intpMat_forLSF)
            nMat = lsf%matArray%nVals
            allocate(moff(nMat+1)); moff(1) = 0
            do i = 1, nMat
              moff(i+1) = moff(i) + size(lsf%matArray%val(i)%A)
            end do
            allocate(mats(moff(nMat+1)))
            do i = 1, nMat      ! row-major (nCoeffs, nSources) as the kernels read it
              offs = moff(i)
              mats(offs+1:moff(i+1)) = reshape(transpose(lsf%matArray%val(i)%A), [size(lsf%matArray%val(i)%A)])
            end do
          end associate
        else
          allocate(moff(1), mats(1)); moff(1) = 0
        end if
        call chk(musb200_intp_register(int(iLevel, c_int), direction, order, int(nTargets, c_int), &
          &      tgt, off, src, wgt, pim, int(nMat, c_int), moff, mats, coord), 'intp_register')
      end associate
    end subroutine one
  end subroutine mus_b200_upload_intp

  !> source = { force = ... }: evaluated by the host exactly as applySrc_force does
  !! (get_valOfIndex + division by fac%body_force), handed over in lattice units; call again
  !! before a step whose force differs (mus_update_sourceVars position of do_fast_singleLevel)
  subroutine mus_b200_set_force(iLevel, order, nElems, posInTotal, forceLattice)
    integer, intent(in) :: iLevel, order, nElems
    integer, intent(in) :: posInTotal(:)
    real(kind=rk), intent(in) :: forceLattice(:)     !< (iElem-1)*3 + 1:3
    call chk(musb200_source_force(int(iLevel, c_int), int(order, c_int), int(nElems, c_int), &
      &      posInTotal, forceLattice, 0_c_int), 'source_force')
  end subroutine mus_b200_set_force

  !> peer-memory halo exchange among the ranks of one node: all-gather the export blobs and ship
  !! every recv position list to the rank that sends into it (include/musb200.h, p2p section)
  subroutine mus_b200_p2p_connect(iLevel, send, recv, comm)
    use mpi
    integer, intent(in) :: iLevel, comm
    type(tem_communication_type), intent(in) :: send, recv
    character(kind=c_char), allocatable :: blob(:), allBlobs(:), peerBlobs(:)
    integer(c_int32_t), allocatable :: proc(:), nVals(:), remotePos(:)
    integer, allocatable :: req(:)
    integer :: iProc, nRanks, iError, n, off
    call mpi_comm_size(comm, nRanks, iError)
    allocate(blob(p2p_blob), allBlobs(p2p_blob*nRanks), peerBlobs(p2p_blob*max(1, send%nProcs)))
    call chk(musb200_p2p_export(int(iLevel, c_int), blob), 'p2p_export')
    call mpi_allgather(blob, p2p_blob, mpi_character, allBlobs, p2p_blob, mpi_character, comm, iError)
    allocate(proc(send%nProcs), nVals(send%nProcs), req(send%nProcs + recv%nProcs))
    n = 0
    do iProc = 1, send%nProcs
      proc(iProc) = send%proc(iProc); nVals(iProc) = send%buf_real(iProc)%nVals
      n = n + nVals(iProc)
      peerBlobs(p2p_blob*(iProc-1)+1:p2p_blob*iProc) = &
        & allBlobs(p2p_blob*send%proc(iProc)+1:p2p_blob*(send%proc(iProc)+1))
    end do
    allocate(remotePos(max(1, n)))
    off = 0
    do iProc = 1, send%nProcs       ! the receiver's list for my message
      call mpi_irecv(remotePos(off+1), nVals(iProc), mpi_integer, send%proc(iProc), 4200 + iLevel, &
        &            comm, req(iProc), iError)
      off = off + nVals(iProc)
    end do
    do iProc = 1, recv%nProcs       ! my recv list goes to its sender
      call mpi_isend(recv%buf_real(iProc)%pos, recv%buf_real(iProc)%nVals, mpi_integer, &
        &            recv%proc(iProc), 4200 + iLevel, comm, req(send%nProcs + iProc), iError)
    end do
    call mpi_waitall(size(req), req, mpi_statuses_ignore, iError)
    call chk(musb200_p2p_connect(int(iLevel, c_int), int(send%nProcs, c_int), proc, peerBlobs, &
      &      nVals, remotePos), 'p2p_connect')
    call mpi_barrier(comm, iError)
  end subroutine mus_b200_p2p_connect

  !> drop-ins of mus_pdf_serialize / mus_pdf_unserialize (mus_buffer_module.fpp:80-190): the
  !! restart chunk is gathered from / scattered to the device state in treeID order
  subroutine mus_b200_pdf_serialize(treeID, levelPointer, nElems, buffer)
    integer, intent(in) :: nElems
    integer(kind=long_k), intent(in) :: treeID(nElems)
    integer, intent(in) :: levelPointer(nElems)
    real(kind=rk), intent(inout) :: buffer(:)
    call chk(musb200_pdf_serialize(int(nElems, c_int), treeID, levelPointer, buffer), 'pdf_serialize')
  end subroutine mus_b200_pdf_serialize

  subroutine mus_b200_pdf_unserialize(treeID, levelPointer, nElems, buffer)
    integer, intent(in) :: nElems
    integer(kind=long_k), intent(in) :: treeID(nElems)
    integer, intent(in) :: levelPointer(nElems)
    real(kind=rk), intent(in) :: buffer(:)
    call chk(musb200_pdf_unserialize(int(nElems, c_int), treeID, levelPointer, buffer), 'pdf_unserialize')
  end subroutine mus_b200_pdf_unserialize

  !> several schemes on one mesh stepped together (a passive scalar in slot 1 transported by the
  !! flow in slot 0): one C call per coarse cycle for all of them, the flow advancing first inside
  !! every level step
  subroutine mus_b200_step_schemes(slots, minLevel, maxLevel, nCycles)
    integer, intent(in) :: slots(:), minLevel, maxLevel, nCycles
    integer(c_int) :: cslots(size(slots))
    cslots = int(slots, c_int)
    call chk(musb200_step_schemes(int(size(slots), c_int), cslots, int(minLevel, c_int), &
      &                           int(maxLevel, c_int), int(nCycles, c_int)), 'step_schemes')
  end subroutine mus_b200_step_schemes

  !> after mus_b200_pdf_unserialize filled state(:, nNext) of the fluid elements on the device
  !! (a restart read straight into the device state): what mus_init_flow does next on the host,
  !! mus_initAuxField + fillHelperElementsFineToCoarse + fillHelperElementsCoarseToFine
  !! (mus_flow_module.fpp:206-240), on the device
  subroutine mus_b200_fill_helper_elements(minLevel, maxLevel)
    integer, intent(in) :: minLevel, maxLevel
    call chk(musb200_fill_helper_elements(int(minLevel, c_int), int(maxLevel, c_int)), 'fill_helper_elements')
  end subroutine mus_b200_fill_helper_elements

  !> control routine: one C call per coarse cycle instead of steps 1-9 of do_fast_singleLevel
  !! (registered in mus_init_control for control_routine = 'b200')
  subroutine mus_b200_step(minLevel, maxLevel, nCycles)
    integer, intent(in) :: minLevel, maxLevel, nCycles
    call chk(musb200_step(int(minLevel, c_int), int(maxLevel, c_int), int(nCycles, c_int)), 'step')
  end subroutine mus_b200_step

  !> refresh the host mirrors (tracking, restart, check_flow_status)
  subroutine mus_b200_download(scheme, minLevel, maxLevel)
    type(mus_scheme_type), intent(inout) :: scheme
    integer, intent(in) :: minLevel, maxLevel
    integer :: iLevel
    integer(c_int) :: nNow, nNext
    do iLevel = minLevel, maxLevel
      call chk(musb200_get_now_next(int(iLevel, c_int), nNow, nNext), 'now_next')
      scheme%pdf(iLevel)%nNow = nNow
      scheme%pdf(iLevel)%nNext = nNext
      call chk(musb200_state_download(int(iLevel, c_int), nNext, scheme%state(iLevel)%val(:,nNext)), 'state')
      call chk(musb200_aux_download(int(iLevel, c_int), scheme%auxField(iLevel)%val), 'aux')
    end do
  end subroutine mus_b200_download

  !> check_density replacement (mus_tools_module.f90:224-313)
  subroutine mus_b200_check(iLevel, totalDens, maxVel, hasNaN)
    integer, intent(in) :: iLevel
    real(kind=rk), intent(out) :: totalDens, maxVel
    logical, intent(out) :: hasNaN
    integer(c_int) :: flag
    real(c_double) :: m, v
    call chk(musb200_reduce(int(iLevel, c_int), m, v, flag), 'reduce')
    totalDens = m; maxVel = v; hasNaN = (flag /= 0)
  end subroutine mus_b200_check

  !> point tracking (tem_tracking with a canoND point, interval = {iter = 1}): density and
  !! velocity of one element of levelDesc(iLevel)%total without downloading auxField --
  !! what mus_derVarPos ... get_element reads from scheme%auxField(iLevel)%val((elemPos-1)*4+1:4)
  subroutine mus_b200_probe(iLevel, elemPos, dens, vel)
    integer, intent(in) :: iLevel, elemPos
    real(kind=rk), intent(out) :: dens, vel(3)
    real(c_double) :: r(4)
    call chk(musb200_aux_probe(int(iLevel, c_int), int(elemPos, c_int), r), 'aux_probe')
    dens = r(1); vel = r(2:4)
  end subroutine mus_b200_probe

  !> tracking objects that are active every iteration: 1 = auxField written by every level step,
  !! 2 = lazy (mus_b200_probe computes the element's moments on demand), 0 = default
  subroutine mus_b200_track_every_step(mode)
    integer, intent(in) :: mode
    call chk(musb200_set_aux_every_step(int(mode, c_int)), 'set_aux_every_step')
  end subroutine mus_b200_track_every_step

  !> mus_scheme_cleanup (mus_scheme_module.f90:441-480) before dynamic load balancing rebuilds the
  !! level descriptors (mus_dynLoadBal_module.f90:116): download first (mus_b200_download), then
  !! drop the device levels; mus_b200_upload of the re-partitioned scheme follows.  Every rank
  !! synchronises before a peer's halo rows disappear (the caller adds the MPI barrier).
  subroutine mus_b200_cleanup(minLevel, maxLevel)
    integer, intent(in) :: minLevel, maxLevel
    integer :: iLevel
    call chk(musb200_synchronize(), 'synchronize')
    do iLevel = minLevel, maxLevel
      call chk(musb200_level_destroy(int(iLevel, c_int)), 'level_destroy')
    end do
  end subroutine mus_b200_cleanup

  !> device time per stage since the last call, in seconds, for mus_timerHandles
  !! (mus_timer_module.f90: compute, setBnd, comm, intp) and the MLUPS report of mus_perf_measure
  subroutine mus_b200_timers(tCompute, tBC, tComm, tIntp)
    real(kind=rk), intent(out) :: tCompute, tBC, tComm, tIntp
    real(c_double) :: c, b, m, i
    call chk(musb200_timers(c, b, m, i), 'timers')
    call chk(musb200_timers_reset(), 'timers_reset')
    tCompute = c * 1.e-3_rk; tBC = b * 1.e-3_rk; tComm = m * 1.e-3_rk; tIntp = i * 1.e-3_rk
  end subroutine mus_b200_timers

  !> passive scalar (scheme kind 'passive_scalar'): the species' relaxation after the level has
  !! been created by mus_b200_upload (mus_init_advRel_lbm_ps, init/mus_initLBMPS_module.f90:59-159:
  !! relax_id bgk with variant 1 = 'first' | 2 = 'second', trt = vStdNoOpt)
  subroutine mus_b200_set_species(scheme, iLevel, relaxId, variant)
    type(mus_scheme_type), intent(in) :: scheme
    integer, intent(in) :: iLevel, relaxId, variant
    call chk(musb200_set_species(int(iLevel, c_int), int(relaxId, c_int), int(variant, c_int),    &
      &      real(scheme%field(1)%fieldProp%species%diff_coeff(1), c_double),                      &
      &      real(scheme%field(1)%fieldProp%species%lambda, c_double)), 'set_species')
  end subroutine mus_b200_set_species

  !> transport velocity of a passive scalar from its space-time function, as the kernels fetch it
  !! at the top of every call (mus_compute_passiveScalar_module.fpp:113-126): get_valOfIndex on
  !! scheme%transVar%method(1) for the elements 1..nElems_solve, times 1/fac%vel.  Once for a
  !! constant field, before every step for a time-dependent one; a velocity that IS another
  !! scheme's flow field is coupled on the device instead (mus_b200_couple_transport_velocity).
  subroutine mus_b200_set_transport_velocity(scheme, iLevel, params, time)
    use tem_time_module, only: tem_time_type
    type(mus_scheme_type), intent(in) :: scheme
    integer, intent(in) :: iLevel
    type(mus_param_type), intent(in) :: params
    type(tem_time_type), intent(in) :: time
    real(kind=rk), allocatable :: transVel(:)
    integer :: nSolve, varPos
    nSolve = scheme%pdf(iLevel)%nElems_solve
    allocate(transVel(3*nSolve))
    varPos = scheme%transVar%method(1)%data_varPos
    call scheme%varSys%method%val(varPos)%get_valOfIndex(                            &
      & varSys = scheme%varSys, time = time, iLevel = iLevel,                         &
      & idx    = scheme%transVar%method(1)%pntIndex%indexLvl(iLevel)%val(1:nSolve),   &
      & nVals  = nSolve, res = transVel )
    transVel = transVel * (1.0_rk / params%physics%fac(iLevel)%vel)
    call chk(musb200_set_transport_velocity(int(iLevel, c_int), int(nSolve, c_int), transVel, &
      &      0_c_int), 'set_transport_velocity')
    deallocate(transVel)
  end subroutine mus_b200_set_transport_velocity

  !> several schemes on one mesh (scheme slots): every following call addresses `slot`
  subroutine mus_b200_bind_scheme(slot)
    integer, intent(in) :: slot
    call chk(musb200_scheme_bind(int(slot, c_int)), 'scheme_bind')
  end subroutine mus_b200_bind_scheme

  !> passive scalar whose transport_velocity is the flow scheme's velocity: read on the device
  !! from the flow's auxField instead of get_valOfIndex + upload per step
  subroutine mus_b200_couple_transport_velocity(iLevel, flowSlot)
    integer, intent(in) :: iLevel, flowSlot
    call chk(musb200_couple_transport_velocity(int(iLevel, c_int), int(flowSlot, c_int), &
      &                                        int(iLevel, c_int)), 'couple_transport_velocity')
  end subroutine mus_b200_couple_transport_velocity

  !> strict drop-in of the `kernel` interface (host arrays in, host arrays out); registered by
  !! mus_init_advRel_fluid for relaxation variant 'b200'.  One H2D + kernel + D2H per call:
  !! meant for verification runs, the production path is control_routine = 'b200'.
  subroutine mus_b200_compute(fieldProp, inState, outState, auxField, neigh, nElems, &
    &                         nSolve, level, layout, params, varSys, derVarPos)
    use mus_field_prop_module,        only: mus_field_prop_type
    use mus_scheme_layout_module,     only: mus_scheme_layout_type
    use mus_derVarPos_module,         only: mus_derVarPos_type
    use tem_varSys_module,            only: tem_varSys_type
    type(mus_field_prop_type), intent(in) :: fieldProp(:)
    type(tem_varSys_type), intent(in) :: varSys
    type(mus_scheme_layout_type), intent(in) :: layout
    integer, intent(in) :: nElems
    real(kind=rk), intent(in)  ::  inState(nElems * varSys%nScalars)
    real(kind=rk), intent(out) :: outState(nElems * varSys%nScalars)
    real(kind=rk), intent(inout) :: auxField(nElems * varSys%nAuxScalars)
    integer, intent(in) :: neigh(nElems * layout%fStencil%QQ)
    integer, intent(in) :: nSolve
    integer, intent(in) :: level
    type(mus_param_type), intent(in) :: params
    type(mus_derVarPos_type), intent(in) :: derVarPos(:)
    !$omp master
    call chk(musb200_compute_host(relax_id, kind_id, int(layout%fStencil%QQ, c_int), inState,      &
      &   outState, auxField, neigh, int(nElems, c_int), int(nSolve, c_int),                        &
      &   fieldProp(1)%fluid%viscKine%omLvl(level)%val, real(fieldProp(1)%fluid%lambda, c_double), &
      &   real(fieldProp(1)%fluid%omegaBulkLvl(level), c_double)), 'compute_host')
    !$omp end master
    !$omp barrier
  end subroutine mus_b200_compute

end module mus_b200_module
