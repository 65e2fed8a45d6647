! -*- mode: F90 -*-
!=============================================================================!
!                       E N E R G Y   (GPU back end)                          !
!=============================================================================!
! Drop-in replacement for molint.F90 of keb721/mc_water_ls_mw: same module    !
! name, same public procedures and variables (molint.F90:22-48), bodies       !
! forwarded to libmwgpu.so through the iso_c_binding module `mwgpu`.          !
! mc_moves.F90 / main.f90 compile against it unchanged.                       !
!                                                                             !
! Two levels of use (INTEGRATION.md):                                         !
!  * fine-grained: every call below uploads the host model (ljr, ref_ljr,     !
!    hmatrix) and runs one small kernel -- bit-compatible, slow (one launch   !
!    per CPU call); meant for validation of the boundary;                     !
!  * coarse: mc_cycle hands the whole move loop to mwgpu_mc_run (see          !
!    mc_cycle_gpu.F90) and the host arrays are refreshed with                 !
!    energy_pull_from_device() before the periodic bookkeeping.               !
!                                                                             !
! NOT COMPILED in this repository (no Fortran compiler in the build image).   !
!=============================================================================!
module energy

  use iso_c_binding, only : c_ptr,c_null_ptr,c_int,c_double,c_associated
  use constants,     only : dp,int32,ang_to_bohr
  use mwgpu

  implicit none
  private

  public :: energy_init
  public :: energy_deinit
  public :: compute_ivects
  public :: compute_model_energy
  public :: compute_local_real_energy
  public :: compute_neighbours

  public :: model_energy
  public :: nivect,ivect
  public :: maxneigh
  public :: nn,jn,vn

  public :: mw_sigma,mw_epsilon,mw_lambda
  public :: sw_bigA,sw_B,sw_gamma,sw_a,sw_p,sw_q,cos0

  ! additions for the coarse path
  public :: gpu_ctx
  public :: energy_push_to_device,energy_pull_from_device

  real(kind=dp),allocatable,dimension(:),save :: model_energy
  integer,allocatable,dimension(:) :: nivect
  real(kind=dp),allocatable,dimension(:,:,:) :: ivect

  ! same values as molint.F90:64-74 (compiled into the kernels as well)
  real(kind=dp),parameter :: mw_sigma   = 2.3925_dp*ang_to_bohr
  real(kind=dp),parameter :: mw_epsilon = 6.189_dp/627.509469_dp
  real(kind=dp),parameter :: mw_lambda  = 23.15_dp
  real(kind=dp),parameter :: sw_bigA = 7.049556277_dp
  real(kind=dp),parameter :: sw_B = 0.6022245584_dp
  real(kind=dp),parameter :: sw_gamma = 1.2_dp
  real(kind=dp),parameter :: sw_a = 1.8_dp
  integer,parameter :: sw_p=4,sw_q=0
  real(kind=dp),parameter :: cos0 = -0.33331324756

  integer,parameter :: maxneigh = 50
  integer,allocatable,dimension(:,:),save :: nn
  integer,allocatable,dimension(:,:,:),save ::jn,vn

  type(c_ptr),save :: gpu_ctx = c_null_ptr      ! one walker per rank, as in the reference

contains

  subroutine energy_push_to_device()
    ! host model -> device (after read_xmol, checkpoint load, chain synchronisation)
    use model, only : ljr,ref_ljr,hmatrix
    implicit none
    call mwgpu_check(mwgpu_upload(gpu_ctx,0_c_int,ljr,ref_ljr,hmatrix),'energy_push_to_device')
  end subroutine energy_push_to_device

  subroutine energy_pull_from_device()
    ! device -> host model (before checkpoint write, dcd snapshot, monitor, therm output)
    use model,      only : ljr,ref_ljr,hmatrix,volume,ls
    use userparams, only : num_lattices
    implicit none
    type(mwgpu_walker_state) :: st
    call mwgpu_check(mwgpu_download(gpu_ctx,0_c_int,ljr,ref_ljr,hmatrix),'energy_pull_from_device')
    call mwgpu_check(mwgpu_mc_get_state(gpu_ctx,0_c_int,st),'energy_pull_from_device')
    model_energy(1:num_lattices) = st%model_energy(1:num_lattices)
    volume(1:num_lattices)       = st%volume(1:num_lattices)
    ls = st%ls
  end subroutine energy_pull_from_device

  subroutine energy_init
    !------------------------------------------------------------------------------!
    ! molint.F90:91-153                                                            !
    !------------------------------------------------------------------------------!
    use userparams, only : num_lattices,nwater
    use model,      only : volume
    use comms,      only : myrank
    implicit none
    integer :: ils,ierr,ndev
    type(mwgpu_walker_state) :: st

    allocate(model_energy(1:num_lattices),stat=ierr)
    if (ierr/=0) stop 'Error allocating model and recip energy arrays'
    allocate(nivect(1:num_lattices),stat=ierr)
    if (ierr/=0) stop 'Error allocating nivect'
    allocate(ivect(1:3,1:MWGPU_MAXIVECT,1:num_lattices),stat=ierr)
    if (ierr/=0) stop 'Error allocating ivect'
    allocate(nn(1:nwater,1:num_lattices),stat=ierr)
    if (ierr/=0) stop 'Error allocating nn array in molint.F90'
    allocate(vn(1:maxneigh,1:nwater,1:num_lattices),stat=ierr)
    if (ierr/=0) stop 'Error allocating vn array in molint.F90'
    allocate(jn(1:maxneigh,1:nwater,1:num_lattices),stat=ierr)
    if (ierr/=0) stop 'Error allocating jn array in molint.F90'

    ndev = mwgpu_device_count()
    if (ndev<1) stop 'mwgpu: no CUDA device (there is no CPU fallback)'
    call mwgpu_check(mwgpu_create(int(nwater,c_int),int(num_lattices,c_int),1_c_int, &
                                  int(mod(myrank,ndev),c_int),gpu_ctx),'energy_init')
    call energy_push_to_device()
    call mwgpu_check(mwgpu_energy_init(gpu_ctx),'energy_init')
    call mwgpu_check(mwgpu_mc_get_state(gpu_ctx,0_c_int,st),'energy_init')
    do ils = 1,num_lattices
       volume(ils)       = st%volume(ils)
       model_energy(ils) = st%model_energy(ils)
       call mwgpu_check(mwgpu_compute_ivects(gpu_ctx,0_c_int,int(ils,c_int),nivect(ils),ivect(:,:,ils)),'energy_init')
       call mwgpu_check(mwgpu_get_neighbours(gpu_ctx,0_c_int,int(ils,c_int),nn(:,ils),jn(:,:,ils),vn(:,:,ils)),'energy_init')
    end do
    return
  end subroutine energy_init

  subroutine energy_deinit()
    implicit none
    integer :: ierr
    call mwgpu_destroy(gpu_ctx)
    gpu_ctx = c_null_ptr
    deallocate(ivect,stat=ierr)
    if (ierr/=0) stop 'Error deallocating ivect'
    return
  end subroutine energy_deinit

  subroutine compute_ivects(ils)
    ! molint.F90:174-217
    implicit none
    integer,intent(in) :: ils
    call energy_push_to_device()
    call mwgpu_check(mwgpu_compute_ivects(gpu_ctx,0_c_int,int(ils,c_int),nivect(ils),ivect(:,:,ils)),'compute_ivects')
  end subroutine compute_ivects

  subroutine compute_neighbours(ils)
    ! molint.F90:501-559
    implicit none
    integer,intent(in) :: ils
    call energy_push_to_device()
    call mwgpu_check(mwgpu_compute_neighbours(gpu_ctx,0_c_int,int(ils,c_int),nn(:,ils),jn(:,:,ils),vn(:,:,ils)), &
                     'compute_neighbours')
  end subroutine compute_neighbours

  subroutine compute_model_energy(ils)
    ! molint.F90:407-499 ; result in model_energy(ils)
    implicit none
    integer,intent(in) :: ils
    real(c_double) :: e
    call energy_push_to_device()
    call mwgpu_check(mwgpu_compute_model_energy(gpu_ctx,0_c_int,int(ils,c_int),e),'compute_model_energy')
    model_energy(ils) = e
  end subroutine compute_model_energy

  real(kind=dp) function compute_local_real_energy(imol,ils)
    ! molint.F90:220-404
    implicit none
    integer,intent(in) :: imol,ils
    real(c_double) :: e
    call energy_push_to_device()
    call mwgpu_check(mwgpu_compute_local_real_energy(gpu_ctx,0_c_int,int(imol,c_int),int(ils,c_int),e), &
                     'compute_local_real_energy')
    compute_local_real_energy = e
  end function compute_local_real_energy

end module energy
