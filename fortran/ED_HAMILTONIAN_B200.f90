! ED_HAMILTONIAN_B200 -- ISO_C_BINDING shim: the reference's ED_HAMILTONIAN interface on top of
! libcdmft_b200.so (include/cdmft_b200.h).  SOURCE ONLY: there is no Fortran compiler in the build
! image, so this file is not compiled or tested there; it is kept purely declarative (bind(C)
! interfaces + thin wrappers) and mirrored 1:1 by cdmft_lanc_ed_b200/ed_hamiltonian.py, which IS tested.
!
! Drop-in use (see INTEGRATION.md): in ED_DIAG.f90 / ED_GF_NORMAL.f90 replace `USE ED_HAMILTONIAN`
! by `USE ED_HAMILTONIAN_B200`.  build_Hv_sector / delete_Hv_sector / vecDim_Hv_sector keep their
! signatures (ED_HAMILTONIAN.f90:39,149,197) and spHtimesV_p (ED_VARS_GLOBAL.f90:146) is bound to
! b200_HxV, which conforms to the abstract interface cc_sparse_HxV (ED_VARS_GLOBAL.f90:72-78).
MODULE ED_HAMILTONIAN_B200
  USE, INTRINSIC :: ISO_C_BINDING
  USE ED_INPUT_VARS   ! Nlat,Norb,Nspin,Nbath,Uloc,Ust,Jh,Jx,Jp,xmu,hfmode,ed_sparse_H
  USE ED_VARS_GLOBAL  ! impHloc, dmft_bath, spHtimesV_p, MpiStatus, MpiRank, MpiSize
  USE ED_BATH         ! Hbath_build
  USE ED_AUX_FUNX     ! index_stride_lso
  implicit none
  private

  public :: b200_init, b200_finalize
  public :: build_Hv_sector, delete_Hv_sector, vecDim_Hv_sector
  public :: b200_HxV
  public :: b200_lanc_eigh, b200_lanc_tridiag, b200_sp_eigh
  public :: b200_scatter_vector, b200_gather_vector
  public :: b200_imp_weights, b200_imp_kinetic, b200_density_matrices

  type, bind(C) :: cdmft_b200_model
     integer(c_int32_t) :: nlat, norb, nspin, nbath
     real(c_double)     :: uloc(5)
     real(c_double)     :: ust, jh, jx, jp, xmu
     integer(c_int32_t) :: hfmode
     integer(c_int32_t) :: quirk_direct_bathdiag
     type(c_ptr)        :: imphloc, hbath, vbath
  end type cdmft_b200_model

  interface
     function c_last_error() bind(C, name="cdmft_b200_last_error") result(p)
       import :: c_ptr
       type(c_ptr) :: p
     end function c_last_error
     function c_init(device) bind(C, name="cdmft_b200_init") result(rc)
       import :: c_int, c_int32_t
       integer(c_int32_t), value :: device
       integer(c_int) :: rc
     end function c_init
     function c_unique_id(uid) bind(C, name="cdmft_b200_nccl_unique_id") result(rc)
       import :: c_int, c_char
       character(kind=c_char) :: uid(128)
       integer(c_int) :: rc
     end function c_unique_id
     function c_init_rank(device, nranks, rank, uid) bind(C, name="cdmft_b200_init_rank") result(rc)
       import :: c_int, c_int32_t, c_char
       integer(c_int32_t), value :: device, nranks, rank
       character(kind=c_char) :: uid(128)
       integer(c_int) :: rc
     end function c_init_rank
     function c_finalize() bind(C, name="cdmft_b200_finalize") result(rc)
       import :: c_int
       integer(c_int) :: rc
     end function c_finalize
     function c_set_model(m) bind(C, name="cdmft_b200_set_model") result(rc)
       import :: c_int, cdmft_b200_model
       type(cdmft_b200_model) :: m
       integer(c_int) :: rc
     end function c_set_model
     function c_vecdim(isector, n) bind(C, name="cdmft_b200_vecdim_hv_sector") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int32_t), value :: isector
       integer(c_int64_t) :: n
       integer(c_int) :: rc
     end function c_vecdim
     function c_build(isector, mode, nloc) bind(C, name="cdmft_b200_build_hv_sector") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int32_t), value :: isector, mode
       integer(c_int64_t) :: nloc
       integer(c_int) :: rc
     end function c_build
     function c_delete() bind(C, name="cdmft_b200_delete_hv_sector") result(rc)
       import :: c_int
       integer(c_int) :: rc
     end function c_delete
     function c_hxv(nloc, v, hv) bind(C, name="cdmft_b200_hxv") result(rc)
       import :: c_int, c_int32_t, c_double_complex
       integer(c_int32_t), value :: nloc
       complex(c_double_complex) :: v(*), hv(*)
       integer(c_int) :: rc
     end function c_hxv
     function c_lanc_gs(nloc, vect, nitermax, threshold, ncheck, egs, niter, alanc, blanc) &
          bind(C, name="cdmft_b200_lanczos_gs") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_double, c_double_complex, c_ptr
       integer(c_int64_t), value :: nloc
       complex(c_double_complex) :: vect(*)
       integer(c_int32_t), value :: nitermax, ncheck
       real(c_double), value :: threshold
       real(c_double) :: egs
       integer(c_int32_t) :: niter
       type(c_ptr), value :: alanc, blanc
       integer(c_int) :: rc
     end function c_lanc_gs
     function c_eigh(nloc, neigen, nblock, nitermax, tol, eig_values, eig_basis, nconv, nmatvec) &
          bind(C, name="cdmft_b200_eigh") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_double, c_double_complex
       integer(c_int64_t), value :: nloc
       integer(c_int32_t), value :: neigen, nblock, nitermax
       real(c_double), value :: tol
       real(c_double) :: eig_values(*)
       complex(c_double_complex) :: eig_basis(*)
       integer(c_int32_t) :: nconv, nmatvec
       integer(c_int) :: rc
     end function c_eigh
     function c_lanc_tridiag(nloc, v0, nitermax, threshold, alanc, blanc, ndone) &
          bind(C, name="cdmft_b200_lanczos_tridiag") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_double, c_double_complex
       integer(c_int64_t), value :: nloc
       complex(c_double_complex) :: v0(*)
       integer(c_int32_t), value :: nitermax
       real(c_double), value :: threshold
       real(c_double) :: alanc(*), blanc(*)
       integer(c_int32_t) :: ndone
       integer(c_int) :: rc
     end function c_lanc_tridiag
     function c_build_hmat(hmat) bind(C, name="cdmft_b200_build_hmat") result(rc)
       import :: c_int, c_double_complex
       complex(c_double_complex) :: hmat(*)
       integer(c_int) :: rc
     end function c_build_hmat
     function c_scatter(vfull, vloc, root) bind(C, name="cdmft_b200_scatter_vector") result(rc)
       import :: c_int, c_int32_t, c_double_complex
       complex(c_double_complex) :: vfull(*), vloc(*)
       integer(c_int32_t), value :: root
       integer(c_int) :: rc
     end function c_scatter
     function c_gather(vloc, vfull, root) bind(C, name="cdmft_b200_gather_vector") result(rc)
       import :: c_int, c_int32_t, c_double_complex
       complex(c_double_complex) :: vloc(*), vfull(*)
       integer(c_int32_t), value :: root
       integer(c_int) :: rc
     end function c_gather
     function c_ipc_export(h) bind(C, name="cdmft_b200_ipc_export") result(rc)
       import :: c_int, c_char
       character(kind=c_char) :: h(128)
       integer(c_int) :: rc
     end function c_ipc_export
     function c_ipc_import(all, nranks) bind(C, name="cdmft_b200_ipc_import") result(rc)
       import :: c_int, c_int32_t, c_char
       character(kind=c_char) :: all(*)
       integer(c_int32_t), value :: nranks
       integer(c_int) :: rc
     end function c_ipc_import
     function c_imp_kinetic(nloc, vec, ek) bind(C, name="cdmft_b200_imp_kinetic") result(rc)
       import :: c_int, c_int64_t, c_double, c_double_complex
       integer(c_int64_t), value :: nloc
       complex(c_double_complex) :: vec(*)
       real(c_double) :: ek(2)
       integer(c_int) :: rc
     end function c_imp_kinetic
     function c_density_matrices(nloc, vec, peso, cdm, spdm) bind(C, name="cdmft_b200_density_matrices") result(rc)
       import :: c_int, c_int64_t, c_double, c_double_complex
       integer(c_int64_t), value :: nloc
       complex(c_double_complex) :: vec(*), cdm(*), spdm(*)
       real(c_double), value :: peso
       integer(c_int) :: rc
     end function c_density_matrices
     function c_imp_weights(nloc, vec, w) bind(C, name="cdmft_b200_imp_weights") result(rc)
       import :: c_int, c_int64_t, c_double, c_double_complex
       integer(c_int64_t), value :: nloc
       complex(c_double_complex) :: vec(*)
       real(c_double) :: w(*)
       integer(c_int) :: rc
     end function c_imp_weights
  end interface

  ! contiguous copies handed to C (must outlive set_model only: the library copies them)
  complex(8), allocatable, target :: c_imphloc(:,:,:,:,:,:), c_hbath(:,:,:,:,:,:,:)
  real(8),    allocatable, target :: c_vbath(:,:)

contains

  subroutine check(rc, where)
    integer(c_int) :: rc
    character(len=*) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: i
    if (rc == 0) return
    call c_f_pointer(c_last_error(), msg, [1024])
    write(*,"(A)",advance="no") trim(where)//": "
    do i = 1, 1024
       if (msg(i) == c_null_char) exit
       write(*,"(A)",advance="no") msg(i)
    end do
    write(*,*)
    stop "ED_HAMILTONIAN_B200 error"      ! the reference's only error mechanism is `stop`
  end subroutine check

  !> once per program, after ed_set_MpiComm: one MPI rank per GPU; rank 0 creates the NCCL id
  subroutine b200_init(device)
    integer :: device
    character(kind=c_char) :: uid(128)
#ifdef _MPI
    integer :: ierr
    if (MpiStatus) then
       if (MpiRank == 0) call check(c_unique_id(uid), "nccl_unique_id")
       call MPI_Bcast(uid, 128, MPI_CHARACTER, 0, MpiComm_Global, ierr)
       call check(c_init_rank(int(device, c_int32_t), int(MpiSize, c_int32_t), int(MpiRank, c_int32_t), uid), "init_rank")
       return
    end if
#endif
    call check(c_init(int(device, c_int32_t)), "init")
  end subroutine b200_init

  subroutine b200_finalize()
    call check(c_finalize(), "finalize")
  end subroutine b200_finalize

  !> pack impHloc, Hbath_build(lambda) and the hybridisations exactly as
  !> ED_HAMILTONIAN_SPARSE_HxV.f90:63-75 / ED_HAMILTONIAN_DIRECT_HxV.f90:59-69 read them
  subroutine push_model()
    type(cdmft_b200_model) :: m
    integer :: ibath, ilat, ispin, iorb
    if (allocated(c_imphloc)) deallocate(c_imphloc, c_hbath, c_vbath)
    allocate(c_imphloc(Nlat,Nlat,Nspin,Nspin,Norb,Norb))
    allocate(c_hbath(Nlat,Nlat,Nspin,Nspin,Norb,Norb,Nbath))
    allocate(c_vbath(Nlat*Nspin*Norb,Nbath))
    c_imphloc = impHloc
    do ibath = 1, Nbath
       c_hbath(:,:,:,:,:,:,ibath) = Hbath_build(dmft_bath%item(ibath)%lambda)
       do ilat = 1, Nlat
          do ispin = 1, Nspin
             do iorb = 1, Norb
                c_vbath(index_stride_lso(ilat,ispin,iorb), ibath) = dmft_bath%item(ibath)%v(index_stride_lso(ilat,ispin,iorb))
             end do
          end do
       end do
    end do
    m%nlat = Nlat; m%norb = Norb; m%nspin = Nspin; m%nbath = Nbath
    m%uloc = 0d0; m%uloc(1:min(5,size(Uloc))) = Uloc(1:min(5,size(Uloc)))
    m%ust = Ust; m%jh = Jh; m%jx = Jx; m%jp = Jp; m%xmu = xmu
    m%hfmode = merge(1, 0, hfmode)
    m%quirk_direct_bathdiag = 0
    m%imphloc = c_loc(c_imphloc); m%hbath = c_loc(c_hbath); m%vbath = c_loc(c_vbath)
    call check(c_set_model(m), "set_model")
  end subroutine push_model

  !> ED_HAMILTONIAN.f90:39-143, same signature: the optional Hmat (caller ED_DIAG.f90:199, LAPACK branch) is the
  !> dense matrix of the sector, assembled on the device (ED_HAMILTONIAN_SPARSE_HxV.f90:112-148)
  subroutine build_Hv_sector(isector, Hmat)
    integer                            :: isector
    complex(8), dimension(:,:), optional :: Hmat
    integer(c_int64_t) :: nloc
#ifdef _MPI
    character(kind=c_char) :: mine(128)
    character(kind=c_char), allocatable :: allh(:)
    integer :: ierr
#endif
    call push_model()      ! the reference re-reads the bath on every direct H*v (:59-69); once per sector here
    call check(c_build(int(isector, c_int32_t), merge(1_c_int32_t, 0_c_int32_t, ed_sparse_H), nloc), "build_Hv_sector")
#ifdef _MPI
    if (MpiStatus .and. MpiSize > 1) then   ! CUDA-IPC windows for the copy-engine exchange (collective)
       allocate(allh(128*MpiSize))
       call check(c_ipc_export(mine), "ipc_export")
       call MPI_Allgather(mine, 128, MPI_CHARACTER, allh, 128, MPI_CHARACTER, MpiComm_Global, ierr)
       call check(c_ipc_import(allh, int(MpiSize, c_int32_t)), "ipc_import")
       deallocate(allh)
    end if
#endif
    if (present(Hmat)) call check(c_build_hmat(Hmat), "build_Hv_sector(Hmat)")
    spHtimesV_p => b200_HxV
  end subroutine build_Hv_sector

  !> ED_HAMILTONIAN.f90:149-190
  subroutine delete_Hv_sector()
    call check(c_delete(), "delete_Hv_sector")
    spHtimesV_p => null()
  end subroutine delete_Hv_sector

  !> ED_HAMILTONIAN.f90:197-221.  The reference's interface is default integer: sectors whose local dimension
  !> does not fit (Ns = 18 on fewer than 2 ranks) stop here instead of wrapping around (ED_SETUP.f90:321 would).
  function vecDim_Hv_sector(isector) result(vecDim)
    integer :: isector, vecDim
    integer(c_int64_t) :: n
    call check(c_vecdim(int(isector, c_int32_t), n), "vecDim_Hv_sector")
    if (n > int(huge(vecDim), c_int64_t)) stop "vecDim_Hv_sector: local dimension exceeds default integer; use more ranks"
    vecDim = int(n)
  end function vecDim_Hv_sector

  !> scatter_vector_MPI / gather_vector_MPI (ED_SETUP.f90:575-668) for the active sector, root = master
  subroutine b200_scatter_vector(v, vloc)
    complex(8), dimension(:) :: v, vloc
    call check(c_scatter(v, vloc, 0_c_int32_t), "scatter_vector_MPI")
  end subroutine b200_scatter_vector
  subroutine b200_gather_vector(vloc, v)
    complex(8), dimension(:) :: vloc, v
    call check(c_gather(vloc, v, 0_c_int32_t), "gather_vector_MPI")
  end subroutine b200_gather_vector

  !> conforms to cc_sparse_HxV (ED_VARS_GLOBAL.f90:72-78): host arrays in, host arrays out
  subroutine b200_HxV(Nloc, v, Hv)
    integer                    :: Nloc
    complex(8), dimension(Nloc) :: v, Hv
    call check(c_hxv(int(Nloc, c_int32_t), v, Hv), "spHtimesV_p")
  end subroutine b200_HxV

  !> replaces `call sp_lanc_eigh(MpiComm,spHtimesV_p,eig_values(1),eig_basis(:,1),Nitermax,threshold=lanc_tolerance)`
  !> at ED_DIAG.f90:176-184: the Krylov vectors never leave the GPU
  subroutine b200_lanc_eigh(egs, vect, Nitermax, threshold, ncheck)
    real(8)                 :: egs
    complex(8), dimension(:) :: vect
    integer                 :: Nitermax
    real(8), optional       :: threshold
    integer, optional       :: ncheck
    integer(c_int32_t)      :: niter
    real(8) :: thr
    integer :: nck
    thr = 1d-12 ; if (present(threshold)) thr = threshold
    nck = 10    ; if (present(ncheck)) nck = ncheck
    call check(c_lanc_gs(int(size(vect), c_int64_t), vect, int(Nitermax, c_int32_t), thr, int(nck, c_int32_t), &
         egs, niter, c_null_ptr, c_null_ptr), "sp_lanc_eigh")
  end subroutine b200_lanc_eigh

  !> replaces `call sp_eigh([MpiComm,]spHtimesV_p,eig_values,eig_basis,Nblock,Nitermax,tol=lanc_tolerance)` at
  !> ED_DIAG.f90:152-169 (the default LANC_METHOD, (P)ARPACK): device-resident thick-restart Lanczos, collective over the
  !> ranks; the Krylov basis never leaves the GPUs.  eig_basis(vecDim,Neigen) as the reference allocates it.
  subroutine b200_sp_eigh(eig_values, eig_basis, Nblock, Nitermax, tol)
    real(8), dimension(:)      :: eig_values
    complex(8), dimension(:,:) :: eig_basis
    integer                    :: Nblock, Nitermax
    real(8), optional          :: tol
    integer(c_int32_t)         :: nconv, nmatvec
    real(8) :: tl
    tl = 0d0 ; if (present(tol)) tl = tol
    if (size(eig_basis, 2) /= size(eig_values)) stop "b200_sp_eigh: eig_basis(:,Neigen) does not match eig_values(Neigen)"
    call check(c_eigh(int(size(eig_basis, 1), c_int64_t), int(size(eig_values), c_int32_t), int(Nblock, c_int32_t), &
         int(Nitermax, c_int32_t), tl, eig_values, eig_basis, nconv, nmatvec), "sp_eigh")
  end subroutine b200_sp_eigh

  !> replaces `call sp_lanc_tridiag(MpiComm,spHtimesV_p,vvloc,alfa_,beta_)` at ED_GF_NORMAL.f90:215
  subroutine b200_lanc_tridiag(vin, alanc, blanc, threshold)
    complex(8), dimension(:) :: vin
    real(8), dimension(:)    :: alanc
    real(8), dimension(size(alanc)) :: blanc
    real(8), optional        :: threshold
    integer(c_int32_t)       :: ndone
    real(8) :: thr
    thr = 1d-12 ; if (present(threshold)) thr = threshold
    call check(c_lanc_tridiag(int(size(vin), c_int64_t), vin, int(size(alanc), c_int32_t), thr, alanc, blanc, ndone), &
         "sp_lanc_tridiag")
  end subroutine b200_lanc_tridiag

  !> replaces the master-only O(Dim) loop of lanc_observables (ED_OBSERVABLES.f90:120-192): the device reduces
  !> gs_weight = |state_cvec|^2 onto the impurity configurations (mu,md) of the up / dw Fock states and hands the
  !> 4**Nimp table back.  The accumulation formulas stay where they are, in ED_OBSERVABLES: its state loop runs over
  !> (mu,md) with gs_weight = peso*W(mu,md) instead of over the Dim basis states (INTEGRATION.md shows the edit).
  !> vec = the local shard of an eigenvector of the ACTIVE sector (build_Hv_sector(isector) first).
  subroutine b200_imp_weights(vec, W)
    complex(8), dimension(:) :: vec
    real(8), dimension(0:,0:) :: W   ! (0:2**Nimp-1, 0:2**Nimp-1) = (mu, md)
    call check(c_imp_weights(int(size(vec), c_int64_t), vec, W), "lanc_observables")
  end subroutine b200_imp_weights

  !> the hopping part of ed_Eknot in lanc_local_energy (ED_OBSERVABLES.f90:305-345): <vec| K |vec>, K = impurity block of
  !> the off-diagonal impHloc of both spins; every other piece of that routine is a sum over the table of b200_imp_weights
  function b200_imp_kinetic(vec) result(ek)
    complex(8), dimension(:) :: vec
    real(8) :: ek, tmp(2)
    call check(c_imp_kinetic(int(size(vec), c_int64_t), vec, tmp), "lanc_local_energy")
    ek = tmp(1)
  end function b200_imp_kinetic

  !> the two accumulations of density_matrix_impurity (ED_OBSERVABLES.f90:465-686) for ONE eigenvector of the active
  !> sector: cluster_density_matrix and single_particle_density_matrix are the module arrays of ED_OBSERVABLES, += peso * ...
  subroutine b200_density_matrices(vec, peso, cluster_dm, sp_dm)
    complex(8), dimension(:)           :: vec
    real(8)                            :: peso
    complex(8), dimension(:,:)         :: cluster_dm
    complex(8), dimension(:,:,:,:,:,:) :: sp_dm
    call check(c_density_matrices(int(size(vec), c_int64_t), vec, peso, cluster_dm, sp_dm), "density_matrix_impurity")
  end subroutine b200_density_matrices

END MODULE ED_HAMILTONIAN_B200
