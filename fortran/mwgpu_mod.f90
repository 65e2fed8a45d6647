! -*- mode: F90 -*-
!=============================================================================!
!                               M W G P U                                     !
!=============================================================================!
! iso_c_binding interface to libmwgpu.so (include/mwgpu.h), the B200-native   !
! implementation of the energy module (molint.F90) and of the move loop of    !
! mc_cycle (mc_moves.F90:217-255).  One interface per C entry point; derived  !
! types mirror the C structs member by member.                                !
!                                                                             !
! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran        !
! compiler.  Shipped as source for hosts that have gfortran/ifort + the       !
! reference sources; see INTEGRATION.md and fortran/Makefile.                 !
!=============================================================================!
module mwgpu

  use iso_c_binding, only : c_int,c_double,c_ptr,c_char,c_int64_t,c_int32_t,c_float,c_null_ptr
  implicit none
  public

  integer(c_int),parameter :: MWGPU_MAXNEIGH = 50     ! molint.F90:79
  integer(c_int),parameter :: MWGPU_MAXIVECT = 32

  ! struct mwgpu_mc_params (include/mwgpu.h) : userparams.f90:14-79 in internal units
  type,bind(C) :: mwgpu_mc_params
     real(c_double)  :: temperature
     real(c_double)  :: pressure
     integer(c_int)  :: npt
     real(c_double)  :: mc_max_trans
     real(c_double)  :: mc_dv_max
     real(c_double)  :: mc_target_ratio
     real(c_double)  :: wl_factor
     integer(c_int)  :: wl_swetnam
     real(c_double)  :: wl_alpha
     integer(c_int)  :: eta_interp
     integer(c_int)  :: samplerun
     integer(c_int)  :: leshift
     integer(c_int)  :: nbins
     real(c_double)  :: mu_min,mu_max
     integer(c_int)  :: allow_switch,allow_vol,allow_trans
     real(c_double)  :: mc_trans_prob,mc_vol_prob,mc_switch_prob
     integer(c_int)  :: mc_always_switch
     integer(c_int)  :: list_update_int
     integer(c_int)  :: eq_mc_cycles
     integer(c_int)  :: max_mc_cycles
     integer(c_int)  :: eq_adjust_mc
     integer(c_int)  :: monitor_int
     integer(c_int)  :: dd
     integer(c_int)  :: window_overlap
     real(c_double)  :: input_ref_enthalpy(2)
     integer(c_int)  :: ls
  end type mwgpu_mc_params

  ! struct mwgpu_walker_state
  type,bind(C) :: mwgpu_walker_state
     real(c_double)     :: model_energy(2)
     real(c_double)     :: volume(2)
     real(c_double)     :: ls_mu
     real(c_double)     :: mc_max_trans,mc_dv_max
     real(c_double)     :: wl_factor
     real(c_double)     :: my_mu_min,my_mu_max
     real(c_double)     :: average_energy(2)
     real(c_double)     :: min_dmu,max_dmu
     real(c_double)     :: ref_enthalpy(2)
     integer(c_int64_t) :: rng_index
     integer(c_int)     :: ls
     integer(c_int)     :: mc_cycle_num
     integer(c_int)     :: accepted(3)
     integer(c_int)     :: attempted(3)
     integer(c_int)     :: my_start_bin,my_end_bin
     integer(c_int)     :: walker_in_window
     integer(c_int)     :: error
     integer(c_int)     :: wl_invt_active
  end type mwgpu_walker_state

  ! struct mwgpu_therm_row: the values of one row of <seed>RRR_therm.dat (main.f90:200-223)
  type,bind(C) :: mwgpu_therm_row
     integer(c_int64_t) :: icyc,ls
     real(c_double)     :: model_energy(2)
     real(c_double)     :: ls_mu
     real(c_double)     :: volume(2)
     real(c_double)     :: hmatrix1(9)
  end type mwgpu_therm_row

  ! struct mwgpu_flat_params (userparams.f90:33-36)
  type,bind(C) :: mwgpu_flat_params
     integer(c_int) :: wl_schedule
     integer(c_int) :: wl_minhist
     real(c_double) :: wl_flattol
     integer(c_int) :: wl_useinvt
  end type mwgpu_flat_params

  ! struct mwgpu_flat_report: what mc_check_flatness writes to the log (walker 0 of the context)
  type,bind(C) :: mwgpu_flat_report
     integer(c_int) :: checked,hist_reset,flat,invt_switched
     real(c_double) :: mean,max_pct,min_pct
     real(c_double) :: wl_factor
  end type mwgpu_flat_report

  interface

     !---------------- lifecycle ----------------!
     integer(c_int) function mwgpu_create(nwater,nlat,nwalkers,device,ctx) bind(C,name='mwgpu_create')
       import :: c_int,c_ptr
       integer(c_int),value :: nwater,nlat,nwalkers,device
       type(c_ptr)          :: ctx                          ! mwgpu_ctx** (out)
     end function mwgpu_create

     subroutine mwgpu_destroy(ctx) bind(C,name='mwgpu_destroy')
       import :: c_ptr
       type(c_ptr),value :: ctx
     end subroutine mwgpu_destroy

     type(c_ptr) function mwgpu_last_error() bind(C,name='mwgpu_last_error')
       import :: c_ptr
     end function mwgpu_last_error

     integer(c_int) function mwgpu_device_count() bind(C,name='mwgpu_device_count')
       import :: c_int
     end function mwgpu_device_count

     !---------------- model state ----------------!
     integer(c_int) function mwgpu_upload(ctx,walker,ljr,ref_ljr,hmatrix) bind(C,name='mwgpu_upload')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value     :: ctx
       integer(c_int),value  :: walker
       real(c_double),intent(in) :: ljr(*),ref_ljr(*),hmatrix(*)
     end function mwgpu_upload

     integer(c_int) function mwgpu_download(ctx,walker,ljr,ref_ljr,hmatrix) bind(C,name='mwgpu_download')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value     :: ctx
       integer(c_int),value  :: walker
       real(c_double),intent(out) :: ljr(*),ref_ljr(*),hmatrix(*)
     end function mwgpu_download

     !---------------- module energy ----------------!
     integer(c_int) function mwgpu_energy_init(ctx) bind(C,name='mwgpu_energy_init')
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function mwgpu_energy_init

     integer(c_int) function mwgpu_compute_ivects(ctx,walker,ils,nivect,ivect) bind(C,name='mwgpu_compute_ivects')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker,ils
       integer(c_int),intent(out) :: nivect
       real(c_double),intent(out) :: ivect(3,*)
     end function mwgpu_compute_ivects

     integer(c_int) function mwgpu_compute_neighbours(ctx,walker,ils,nn,jn,vn) bind(C,name='mwgpu_compute_neighbours')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker,ils
       integer(c_int),intent(out) :: nn(*),jn(*),vn(*)
     end function mwgpu_compute_neighbours

     integer(c_int) function mwgpu_get_neighbours(ctx,walker,ils,nn,jn,vn) bind(C,name='mwgpu_get_neighbours')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker,ils
       integer(c_int),intent(out) :: nn(*),jn(*),vn(*)
     end function mwgpu_get_neighbours

     integer(c_int) function mwgpu_compute_model_energy(ctx,walker,ils,energy) bind(C,name='mwgpu_compute_model_energy')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker,ils
       real(c_double),intent(out) :: energy
     end function mwgpu_compute_model_energy

     integer(c_int) function mwgpu_compute_local_real_energy(ctx,walker,imol,ils,energy) &
          bind(C,name='mwgpu_compute_local_real_energy')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker,imol,ils
       real(c_double),intent(out) :: energy
     end function mwgpu_compute_local_real_energy

     !---------------- mc_moves ----------------!
     integer(c_int) function mwgpu_mc_init(ctx,p,first_rank,size,file_weights,n_file_weights,file_wl_factor) &
          bind(C,name='mwgpu_mc_init')
       import :: c_int,c_ptr,c_double,mwgpu_mc_params
       type(c_ptr),value         :: ctx
       type(mwgpu_mc_params),intent(in) :: p
       integer(c_int),value      :: first_rank,size,n_file_weights
       real(c_double),intent(in) :: file_weights(*)
       real(c_double),value      :: file_wl_factor
     end function mwgpu_mc_init

     integer(c_int) function mwgpu_mc_set_rng_philox(ctx,seed,first_stream,start_index) &
          bind(C,name='mwgpu_mc_set_rng_philox')
       import :: c_int,c_ptr,c_int64_t,c_int32_t
       type(c_ptr),value         :: ctx
       integer(c_int64_t),value  :: seed,start_index
       integer(c_int32_t),value  :: first_stream
     end function mwgpu_mc_set_rng_philox

     ! next draw index of one walker (-1: all): a restart must advance the stream (checkpoints carry no generator state)
     integer(c_int) function mwgpu_mc_set_rng_index(ctx,walker,index) bind(C,name='mwgpu_mc_set_rng_index')
       import :: c_int,c_ptr,c_int64_t
       type(c_ptr),value         :: ctx
       integer(c_int),value      :: walker
       integer(c_int64_t),value  :: index
     end function mwgpu_mc_set_rng_index

     ! which walker kernel mwgpu_mc_run uses: 0 automatic, 1 one warp per walker, 2 one warp per lattice,
     ! 4 two warps per lattice
     integer(c_int) function mwgpu_mc_set_kernel(ctx,warps_per_walker) bind(C,name='mwgpu_mc_set_kernel')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: warps_per_walker
     end function mwgpu_mc_set_kernel

     ! scheduling of a launch: MC cycles per unit of work, persistent blocks (0 = automatic); never changes a result
     integer(c_int) function mwgpu_mc_set_schedule(ctx,chunk_cycles,max_blocks) bind(C,name='mwgpu_mc_set_schedule')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: chunk_cycles,max_blocks
     end function mwgpu_mc_set_schedule

     ! %globaltimer (ns) at the start / end of every walker's part of the last mwgpu_mc_run: start_end_ns(2,nwalkers)
     integer(c_int) function mwgpu_mc_get_walker_times(ctx,start_end_ns) bind(C,name='mwgpu_mc_get_walker_times')
       import :: c_int,c_ptr,c_int64_t
       type(c_ptr),value  :: ctx
       integer(c_int64_t) :: start_end_ns(*)
     end function mwgpu_mc_get_walker_times

     integer(c_int) function mwgpu_mc_set_rng_fifo(ctx,u,n) bind(C,name='mwgpu_mc_set_rng_fifo')
       import :: c_int,c_ptr,c_int64_t,c_double
       type(c_ptr),value         :: ctx
       real(c_double),intent(in) :: u(*)
       integer(c_int64_t),value  :: n
     end function mwgpu_mc_set_rng_fifo

     integer(c_int) function mwgpu_mc_run(ctx,ncycles) bind(C,name='mwgpu_mc_run')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: ncycles
     end function mwgpu_mc_run

     integer(c_int) function mwgpu_mc_get_state(ctx,walker,state) bind(C,name='mwgpu_mc_get_state')
       import :: c_int,c_ptr,mwgpu_walker_state
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker
       type(mwgpu_walker_state),intent(out) :: state
     end function mwgpu_mc_get_state

     integer(c_int) function mwgpu_mc_get_translations(ctx,walker,mc_translations) &
          bind(C,name='mwgpu_mc_get_translations')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker
       integer(c_int),intent(out) :: mc_translations(*)
     end function mwgpu_mc_get_translations

     integer(c_int) function mwgpu_mc_get_bins(ctx,walker,weight,histogram,unbiased_hist) &
          bind(C,name='mwgpu_mc_get_bins')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker
       real(c_double),intent(out) :: weight(*),histogram(*),unbiased_hist(*)
     end function mwgpu_mc_get_bins

     integer(c_int) function mwgpu_mc_set_bins(ctx,walker,weight,histogram,unbiased_hist) &
          bind(C,name='mwgpu_mc_set_bins')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker
       real(c_double),intent(in) :: weight(*),histogram(*),unbiased_hist(*)
     end function mwgpu_mc_set_bins

     integer(c_int) function mwgpu_mc_set_wl_factor(ctx,walker,wl_factor,wl_invt_active) &
          bind(C,name='mwgpu_mc_set_wl_factor')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker,wl_invt_active
       real(c_double),value :: wl_factor
     end function mwgpu_mc_set_wl_factor

     integer(c_int) function mwgpu_mc_set_active_lattice(ctx,walker,ls) bind(C,name='mwgpu_mc_set_active_lattice')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: walker,ls
     end function mwgpu_mc_set_active_lattice

     integer(c_int) function mwgpu_mc_monitor(ctx) bind(C,name='mwgpu_mc_monitor')
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function mwgpu_mc_monitor

     integer(c_int) function mwgpu_mc_chain_sync(ctx) bind(C,name='mwgpu_mc_chain_sync')
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function mwgpu_mc_chain_sync

     ! mc_check_flatness (mc_moves.F90:1936-2185) on the device: state effects for every walker
     integer(c_int) function mwgpu_mc_check_flatness(ctx,p,report) bind(C,name='mwgpu_mc_check_flatness')
       import :: c_int,c_ptr,mwgpu_flat_params,mwgpu_flat_report
       type(c_ptr),value                   :: ctx
       type(mwgpu_flat_params),intent(in)  :: p
       type(mwgpu_flat_report),intent(out) :: report
     end function mwgpu_mc_check_flatness

     ! mc_compute_deltaG_from_hist (mc_moves.F90:2498-2621) on the device
     integer(c_int) function mwgpu_mc_deltag_from_hist(ctx,deltaG,normP) bind(C,name='mwgpu_mc_deltag_from_hist')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value          :: ctx
       real(c_double),intent(out) :: deltaG
       real(c_double),intent(out) :: normP(*)
     end function mwgpu_mc_deltag_from_hist

     ! mc_checkpoint_load (mc_moves.F90:403-501) + restart refresh (:842-862) for one uploaded walker
     integer(c_int) function mwgpu_mc_restore(ctx,walker,mc_cycle_num,mc_max_trans,mc_dv_max,wl_factor, &
          wl_invt_active,ls,histogram,weight,unbiased_hist,hmatrix,ref_ljr,ljr) bind(C,name='mwgpu_mc_restore')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value         :: ctx
       integer(c_int),value      :: walker,mc_cycle_num,wl_invt_active,ls
       real(c_double),value      :: mc_max_trans,mc_dv_max,wl_factor
       real(c_double),intent(in) :: histogram(*),weight(*),unbiased_hist(*)
       real(c_double),intent(in) :: hmatrix(*),ref_ljr(*),ljr(*)
     end function mwgpu_mc_restore

     ! therm rows (main.f90:200-223) recorded by the walker kernel
     integer(c_int) function mwgpu_mc_set_therm(ctx,file_output_int,capacity) bind(C,name='mwgpu_mc_set_therm')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: file_output_int,capacity
     end function mwgpu_mc_set_therm

     integer(c_int) function mwgpu_mc_get_therm(ctx,walker,rows,max_rows,nrows,ndropped) bind(C,name='mwgpu_mc_get_therm')
       import :: c_int,c_ptr,mwgpu_therm_row
       type(c_ptr),value                 :: ctx
       integer(c_int),value              :: walker,max_rows
       type(mwgpu_therm_row),intent(out) :: rows(*)
       integer(c_int),intent(out)        :: nrows,ndropped
     end function mwgpu_mc_get_therm

     !---------------- many walkers per context (one MPI rank drives a batch) ----------------!
     integer(c_int) function mwgpu_num_walkers(ctx) bind(C,name='mwgpu_num_walkers')
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function mwgpu_num_walkers

     ! arrays carry a trailing walker dimension: ljr(3,1,nwater,nlat,nwalkers), hmatrix(3,3,nlat,nwalkers)
     integer(c_int) function mwgpu_upload_all(ctx,ljr,ref_ljr,hmatrix) bind(C,name='mwgpu_upload_all')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value         :: ctx
       real(c_double),intent(in) :: ljr(*),ref_ljr(*),hmatrix(*)
     end function mwgpu_upload_all

     integer(c_int) function mwgpu_download_all(ctx,ljr,ref_ljr,hmatrix) bind(C,name='mwgpu_download_all')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value          :: ctx
       real(c_double),intent(out) :: ljr(*),ref_ljr(*),hmatrix(*)
     end function mwgpu_download_all

     ! compute_local_real_energy for every molecule of a lattice in one launch: energy(nwater)
     integer(c_int) function mwgpu_compute_local_real_energy_all(ctx,walker,ils,energy) &
          bind(C,name='mwgpu_compute_local_real_energy_all')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value          :: ctx
       integer(c_int),value       :: walker,ils
       real(c_double),intent(out) :: energy(*)
     end function mwgpu_compute_local_real_energy_all

     integer(c_int) function mwgpu_compute_neighbours_all(ctx) bind(C,name='mwgpu_compute_neighbours_all')
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function mwgpu_compute_neighbours_all

     ! compute_model_energy for every lattice of every walker: energies(nlat,nwalkers)
     integer(c_int) function mwgpu_compute_model_energy_all(ctx,energies) bind(C,name='mwgpu_compute_model_energy_all')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value          :: ctx
       real(c_double),intent(out) :: energies(*)
     end function mwgpu_compute_model_energy_all

     ! mwgpu_mc_run without the host synchronisation; mwgpu_synchronize waits and reports the walkers' error flags
     integer(c_int) function mwgpu_mc_run_async(ctx,ncycles) bind(C,name='mwgpu_mc_run_async')
       import :: c_int,c_ptr
       type(c_ptr),value    :: ctx
       integer(c_int),value :: ncycles
     end function mwgpu_mc_run_async

     integer(c_int) function mwgpu_synchronize(ctx) bind(C,name='mwgpu_synchronize')
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function mwgpu_synchronize

     integer(c_int) function mwgpu_mc_get_states(ctx,states) bind(C,name='mwgpu_mc_get_states')
       import :: c_int,c_ptr,mwgpu_walker_state
       type(c_ptr),value                    :: ctx
       type(mwgpu_walker_state),intent(out) :: states(*)
     end function mwgpu_mc_get_states

     ! the bin grid of mc_init (mc_moves.F90:557-656): mu_bin(nbins), binwidth(nbins),
     ! scalars(4) = r_pos, r_neg, av_binwidth, log_unbiased_norm
     integer(c_int) function mwgpu_mc_get_grid(ctx,mu_bin,binwidth,scalars) bind(C,name='mwgpu_mc_get_grid')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value          :: ctx
       real(c_double),intent(out) :: mu_bin(*),binwidth(*),scalars(*)
     end function mwgpu_mc_get_grid

     !---------------- measurement helpers ----------------!
     integer(c_int) function mwgpu_last_kernel_ms(ctx,ms) bind(C,name='mwgpu_last_kernel_ms')
       import :: c_int,c_ptr,c_float
       type(c_ptr),value        :: ctx
       real(c_float),intent(out) :: ms
     end function mwgpu_last_kernel_ms

     integer(c_int) function mwgpu_timer_start(ctx) bind(C,name='mwgpu_timer_start')
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function mwgpu_timer_start

     integer(c_int) function mwgpu_timer_stop(ctx,ms) bind(C,name='mwgpu_timer_stop')
       import :: c_int,c_ptr,c_float
       type(c_ptr),value        :: ctx
       real(c_float),intent(out) :: ms
     end function mwgpu_timer_stop

     integer(c_int) function mwgpu_measure_fp64_peak(device,tflops) bind(C,name='mwgpu_measure_fp64_peak')
       import :: c_int,c_double
       integer(c_int),value       :: device
       real(c_double),intent(out) :: tflops
     end function mwgpu_measure_fp64_peak

     integer(c_int) function mwgpu_kernel_launches(ctx,count) bind(C,name='mwgpu_kernel_launches')
       import :: c_int,c_ptr,c_int64_t
       type(c_ptr),value              :: ctx
       integer(c_int64_t),intent(out) :: count
     end function mwgpu_kernel_launches

     !---------------- comms ----------------!
     ! (mwgpu_comms_reduce_local / _apply -- the in-process merge over several contexts of one process -- are not
     !  bound here: a Fortran host has one context per MPI rank and calls mwgpu_comms_allreduce_bins)
     ! comms_join_uhist / comms_join_eta (comms_mpi.f90:299-375, :377-459)
     integer(c_int) function mwgpu_comms_join_uhist(ctx,overlap,joined) bind(C,name='mwgpu_comms_join_uhist')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value          :: ctx
       integer(c_int),value       :: overlap
       real(c_double),intent(out) :: joined(*)
     end function mwgpu_comms_join_uhist

     integer(c_int) function mwgpu_comms_join_eta(ctx,overlap,joined) bind(C,name='mwgpu_comms_join_eta')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value          :: ctx
       integer(c_int),value       :: overlap
       real(c_double),intent(out) :: joined(*)
     end function mwgpu_comms_join_eta

     integer(c_int) function mwgpu_comms_get_unique_id(id128) bind(C,name='mwgpu_comms_get_unique_id')
       import :: c_int,c_char
       character(kind=c_char),intent(out) :: id128(128)
     end function mwgpu_comms_get_unique_id

     integer(c_int) function mwgpu_comms_init(ctx,nranks,rank,id128) bind(C,name='mwgpu_comms_init')
       import :: c_int,c_ptr,c_char
       type(c_ptr),value    :: ctx
       integer(c_int),value :: nranks,rank
       character(kind=c_char),intent(in) :: id128(128)
     end function mwgpu_comms_init

     integer(c_int) function mwgpu_comms_allreduce_bins(ctx) bind(C,name='mwgpu_comms_allreduce_bins')
       import :: c_int,c_ptr
       type(c_ptr),value :: ctx
     end function mwgpu_comms_allreduce_bins

     integer(c_int) function mwgpu_comms_set_hist_base(ctx,histogram,unbiased_hist) &
          bind(C,name='mwgpu_comms_set_hist_base')
       import :: c_int,c_ptr,c_double
       type(c_ptr),value         :: ctx
       real(c_double),intent(in) :: histogram(*),unbiased_hist(*)
     end function mwgpu_comms_set_hist_base

  end interface

contains

  subroutine mwgpu_check(ierr,where)
    !--------------------------------------------------------------------------!
    ! The reference reports every failure with `stop 'text'`; keep that.        !
    !--------------------------------------------------------------------------!
    use iso_c_binding, only : c_f_pointer,c_associated
    integer(c_int),intent(in)   :: ierr
    character(len=*),intent(in) :: where
    character(kind=c_char),pointer :: msg(:)
    type(c_ptr) :: p
    integer :: i
    if (ierr==0) return
    p = mwgpu_last_error()
    write(0,'("mwgpu error ",I6," in ",A)')ierr,where
    if (c_associated(p)) then
       call c_f_pointer(p,msg,[512])
       do i = 1,512
          if (msg(i)==achar(0)) exit
          write(0,'(A1)',advance='no')msg(i)
       end do
       write(0,*)
    end if
    stop 'mwgpu call failed'
  end subroutine mwgpu_check

end module mwgpu
